"""Single-host launcher of the ZeroMQ backend: one monitor + N node processes.

Parity: reference ``murmura/distributed/runner.py:33-213`` (``run_id``, shared
``t_start = monotonic()+grace``, DMTT process class when ``config.dmtt`` is set, 5 s straggler
join, history through an ``mp.Queue``).  Processes are started with the ``spawn`` context so a
parent that already initialised CUDA stays usable.
"""
from __future__ import annotations

import multiprocessing as mp
import time
import traceback
import uuid
from pathlib import Path
from typing import Any, Dict, List, Set

from murmura_b200.config.loader import load_config
from murmura_b200.config.schema import Config, DistributedConfig
from murmura_b200.distributed.endpoints import Endpoints
from murmura_b200.utils.factories import build_attack


def _monitor_main(num_nodes: int, dist_cfg: dict, run_id: str, rounds: int, t_start: float,
                  compromised: Set[int], verbose: bool, out) -> None:
    try:
        from murmura_b200.distributed.monitor import Monitor
        cfg = DistributedConfig(**dist_cfg)
        history = Monitor(num_nodes=num_nodes, endpoints=Endpoints(cfg, num_nodes, run_id), rounds=rounds,
                          t_start=t_start, round_duration_s=cfg.round_duration_s,
                          compromised_nodes=compromised, verbose=verbose).run()
        out.put(("ok", history))
    except Exception:
        out.put(("error", traceback.format_exc()))


def node_process_class(config: Config):
    if config.dmtt is not None:
        from murmura_b200.dmtt.node_process import DMTTNodeProcess
        return DMTTNodeProcess
    from murmura_b200.distributed.node_process import NodeProcess
    return NodeProcess


def _node_main(node_id: int, config_path: str, dist_cfg: dict, num_nodes: int, run_id: str, t_start: float) -> None:
    try:
        endpoints = Endpoints(DistributedConfig(**dist_cfg), num_nodes, run_id)
        cls = node_process_class(load_config(config_path))
        cls.from_config_path(node_id=node_id, config_path=config_path, endpoints=endpoints, t_start=t_start).run()
    except Exception:
        print(f"[Node {node_id}] FATAL:\n{traceback.format_exc()}", flush=True)


class DistributedRunner:
    def __init__(self, config_path: Path):
        self.config_path = Path(config_path)
        self.config: Config = load_config(self.config_path)

    def run(self, verbose: bool = False) -> Dict[str, List[Any]]:
        cfg = self.config
        n = cfg.topology.num_nodes
        run_id = uuid.uuid4().hex[:8]
        attack = build_attack(cfg)
        compromised: Set[int] = set(attack.get_compromised_nodes()) if attack else set()
        Endpoints(cfg.distributed, n, run_id).ensure_dirs()
        dist_cfg = cfg.distributed.model_dump()
        t_start = time.monotonic() + cfg.distributed.startup_grace_s
        print(f"[DistributedRunner] run_id={run_id}  t_start={t_start:.3f}  "
              f"(startup_grace={cfg.distributed.startup_grace_s}s)", flush=True)
        ctx = mp.get_context("spawn")
        out = ctx.Queue()
        monitor = ctx.Process(target=_monitor_main, name="murmura-monitor", daemon=True,
                              args=(n, dist_cfg, run_id, cfg.experiment.rounds, t_start, compromised, verbose, out))
        monitor.start()
        time.sleep(0.2)
        workers = [ctx.Process(target=_node_main, name=f"murmura-node-{i}", daemon=True,
                               args=(i, str(self.config_path), dist_cfg, n, run_id, t_start)) for i in range(n)]
        for w in workers:
            w.start()
        status, value = None, None
        while monitor.is_alive() or not out.empty():
            try:
                status, value = out.get(timeout=0.5)
                break
            except Exception:
                continue
        monitor.join(timeout=5.0)
        for w in workers:
            w.join(timeout=5.0)
            if w.is_alive():
                w.terminate()
        if status is None:
            raise RuntimeError("Monitor exited without producing a result.")
        if status == "error":
            raise RuntimeError(f"Monitor failed:\n{value}")
        return value
