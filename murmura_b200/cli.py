"""Command-line interface: ``murmura run | run-node | list-components``.

Parity: reference ``murmura/cli.py:22-308`` (same commands, options, exit code 1 + ``Error: …``
on any exception, Rich results table).  Additions: ``backend: b200`` routing, and
``murmura bench`` / ``murmura build`` helpers.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch
import typer
from rich.console import Console
from rich.table import Table

from murmura_b200.config import load_config

app = typer.Typer(name="murmura", help="Murmura-B200: decentralized federated learning on Blackwell",
                  add_completion=False)
console = Console()


def _banner(config, backend_label: str) -> None:
    console.print(f"\n[bold]Experiment:[/bold] {config.experiment.name}")
    console.print(f"  Backend:    {backend_label}")
    console.print(f"  Rounds:     {config.experiment.rounds}")
    console.print(f"  Topology:   {config.topology.type} ({config.topology.num_nodes} nodes)")
    console.print(f"  Aggregation:{config.aggregation.algorithm}")
    if config.attack.enabled:
        console.print(f"  [red]Attack: {config.attack.type} ({config.attack.percentage * 100:.0f}% nodes)[/red]")


@app.command()
def run(config_path: Path = typer.Argument(..., help="Path to configuration file (YAML/JSON)"),
        device_override: Optional[str] = typer.Option(None, "--device", help="Override device (cpu/cuda/mps)"),
        verbose: bool = typer.Option(True, "--verbose/--quiet", help="Enable verbose output")):
    """Run an experiment; the backend (simulation / distributed / b200) comes from the config."""
    try:
        console.print(f"[bold blue]Loading config:[/bold blue] {config_path}")
        config = load_config(config_path)
        if config.backend == "distributed":
            _run_distributed(config_path, verbose)
        else:
            _run_in_process(config, device_override, verbose)
    except Exception as exc:
        console.print(f"\n[bold red]Error:[/bold red] {exc}")
        raise typer.Exit(1)


def _run_in_process(config, device_override: Optional[str], verbose: bool) -> None:
    from murmura_b200.core.network import Network
    from murmura_b200.utils.device import get_device
    from murmura_b200.utils.factories import (build_aggregator_factory, build_criterion, build_dataset_adapter,
                                              build_model_factory)
    from murmura_b200.utils.seed import set_seed

    set_seed(config.experiment.seed)
    device = torch.device(device_override) if device_override else get_device()
    console.print(f"[bold green]Device:[/bold green] {device}")
    _banner(config, "b200 (fused sm_100a exchange+aggregate)" if config.backend == "b200" else "simulation")
    console.print(f"\n[bold]Loading dataset:[/bold] {config.data.adapter}")
    adapter = build_dataset_adapter(config)
    console.print(f"[bold]Loading model:[/bold] {config.model.factory}")
    model_factory = build_model_factory(config)
    aggregator_factory = build_aggregator_factory(config, model_factory, device)
    criterion, evidential = build_criterion(config)
    if evidential:
        console.print("[bold cyan]Using evidential deep learning (EDL)[/bold cyan]")
    console.print("\n[bold]Creating network…[/bold]")
    network = Network.from_config(config=config, model_factory=model_factory, dataset_adapter=adapter,
                                  aggregator_factory=aggregator_factory, device=device, criterion=criterion,
                                  evidential=evidential)
    console.print("\n[bold green]Starting training…[/bold green]\n")
    history = network.train(rounds=config.experiment.rounds, local_epochs=config.training.local_epochs,
                            lr=config.training.lr, verbose=verbose or config.experiment.verbose)
    if getattr(network, "is_primary", True):
        _display_results(history)
        perf = getattr(network, "perf_summary", None)
        if perf:
            console.print(f"\n[bold]Throughput:[/bold] {perf()}")
        console.print("\n[bold green]✓ Training complete[/bold green]")


def _run_distributed(config_path: Path, verbose: bool) -> None:
    from murmura_b200.distributed.runner import DistributedRunner
    config = load_config(config_path)
    _banner(config, f"distributed (ZMQ {config.distributed.transport.upper()})")
    console.print("\n[bold green]Launching distributed processes…[/bold green]\n")
    history = DistributedRunner(config_path).run(verbose=verbose or config.experiment.verbose)
    _display_results(history)
    console.print("\n[bold green]✓ Distributed training complete[/bold green]")


@app.command(name="run-node")
def run_node(config_path: Path = typer.Argument(..., help="Path to the shared config file"),
             node_id: int = typer.Option(..., "--node-id", "-n", help="This node's ID (0-indexed)"),
             t_start: float = typer.Option(..., "--t-start", help="Monotonic start time of round 0 printed by the head node"),
             run_id: str = typer.Option("default", "--run-id", help="Run identifier printed by the head node"),
             host_override: Optional[str] = typer.Option(None, "--host", help="Override host for TCP transport")):
    """Launch one ZeroMQ node process (multi-machine deployments)."""
    try:
        config = load_config(config_path)
        if config.backend != "distributed":
            console.print("[yellow]Warning:[/yellow] config.backend is not 'distributed'. "
                          "Proceeding anyway with distributed node.")
        dist_cfg = config.distributed
        if host_override:
            from murmura_b200.config.schema import DistributedConfig
            dist_cfg = DistributedConfig(**{**config.distributed.model_dump(), "host": host_override})
        from murmura_b200.distributed.endpoints import Endpoints
        from murmura_b200.distributed.runner import node_process_class
        endpoints = Endpoints(dist_cfg, config.topology.num_nodes, run_id)
        console.print(f"[bold green]Starting node {node_id}[/bold green]  run_id={run_id}")
        node_process_class(config).from_config_path(node_id=node_id, config_path=str(config_path),
                                                    endpoints=endpoints, t_start=t_start).run()
    except Exception as exc:
        console.print(f"\n[bold red]Error:[/bold red] {exc}")
        raise typer.Exit(1)


_COMPONENTS = {
    "topologies": ("Available Topologies:", [
        "ring      — each node connects to 2 neighbours", "fully     — all-to-all",
        "erdos     — Erdős-Rényi random graph (requires p)", "k-regular — k-regular ring lattice (requires k)",
        "mobility  — time-varying G^t from the random-walk model (add a mobility: block)"]),
    "aggregators": ("Available Aggregators:", [
        "fedavg           — simple decentralised averaging", "krum             — Krum Byzantine-resilient selection",
        "balance          — distance-based with adaptive thresholds",
        "sketchguard      — Count-Sketch compression filtering",
        "ubar             — two-stage (distance + loss) Byzantine-resilient",
        "evidential_trust — uncertainty-aware trust aggregation (EDL)"]),
    "attacks": ("Available Attacks:", [
        "gaussian           — Gaussian noise injection", "directed_deviation — directional parameter scaling",
        "topology_liar      — falsified DMTT topology claims (optionally wrapping a model attack)"]),
    "backends": ("Available Backends:", [
        "simulation  — single-process in-memory (default)",
        "distributed — multi-process ZMQ (set backend: distributed in config)",
        "b200        — one process per GPU, fused sm_100a exchange+aggregate kernels over NVLink"]),
}


@app.command()
def list_components(component_type: str = typer.Argument(..., help="topologies/aggregators/attacks/backends")):
    """List available components."""
    if component_type not in _COMPONENTS:
        console.print(f"[red]Unknown component type: {component_type}[/red]")
        console.print("Available: topologies, aggregators, attacks, backends")
        return
    title, rows = _COMPONENTS[component_type]
    console.print(f"[bold]{title}[/bold]")
    for row in rows:
        console.print(f"  • {row}")
    if component_type == "backends":
        console.print("\n[bold]Distributed transport options:[/bold]")
        console.print("  • ipc — IPC sockets, single machine (default)")
        console.print("  • tcp — TCP sockets, multi-machine")


@app.command()
def build():
    """Compile the sm_100a CUDA extension in-tree."""
    from murmura_b200.ops import build_extension
    console.print(f"built {build_extension(verbose=True)}")


def _display_results(history: dict) -> None:
    evid = len(history.get("mean_vacuity", [])) > 0
    table = Table(title="Training Results")
    for col, style in (("Round", "cyan"), ("Mean Acc", "green"), ("Std Acc", "yellow"), ("Honest Acc", "blue"),
                       ("Comp. Acc", "red")):
        table.add_column(col, style=style)
    if evid:
        for col, style in (("Vacuity", "magenta"), ("Entropy", "cyan"), ("Strength", "white")):
            table.add_column(col, style=style)

    def cell(series, i):
        return f"{series[i]:.4f}" if i < len(series) else "-"

    for i, rnd in enumerate(history["round"]):
        row = [str(rnd), f"{history['mean_accuracy'][i]:.4f}", f"{history['std_accuracy'][i]:.4f}",
               cell(history["honest_accuracy"], i), cell(history["compromised_accuracy"], i)]
        if evid:
            row += [f"{history['mean_vacuity'][i]:.4f}", f"{history['mean_entropy'][i]:.4f}",
                    f"{history['mean_strength'][i]:.2f}"]
        table.add_row(*row)
    console.print(table)
    if evid:
        console.print("\n[bold]Uncertainty Metrics:[/bold]")
        console.print("  • [magenta]Vacuity[/magenta]: epistemic uncertainty — lower is more confident")
        console.print("  • [cyan]Entropy[/cyan]: aleatoric uncertainty — lower is more decisive")
        console.print("  • [white]Strength[/white]: Dirichlet strength — higher means more evidence")


if __name__ == "__main__":
    app()
