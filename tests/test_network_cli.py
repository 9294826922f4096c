"""End-to-end CPU tests: simulation backend, stdout contract, CLI, ZeroMQ backend, DMTT state."""
import math
import os
import re
import subprocess
import sys

import pytest
import torch
from typer.testing import CliRunner

from murmura_b200 import Network, create_topology
from murmura_b200.config import Config, load_config
from murmura_b200.config.schema import DMTTConfig
from murmura_b200.dmtt import DMTTNodeState
from murmura_b200.dmtt.node_process import verify_claim
from murmura_b200.utils.factories import (build_aggregator_factory, build_attack, build_criterion, build_dataset_adapter,
                                          build_mobility_model, build_model_factory)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "murmura_b200", "examples", "configs")


def _cfg(**over):
    base = {"experiment": {"name": "t", "rounds": 2, "seed": 1}, "topology": {"type": "ring", "num_nodes": 4},
            "aggregation": {"algorithm": "fedavg"}, "training": {"batch_size": 16, "lr": 0.05},
            "data": {"adapter": "synthetic.mnist", "params": {"samples_per_node": 48, "partition_method": "iid"}},
            "model": {"factory": "models.mlp", "params": {"hidden_dims": [16]}}}
    base.update(over)
    return Config(**base)


def _build(cfg):
    adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    crit, evid = build_criterion(cfg)
    return Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf, torch.device("cpu")),
                               device=torch.device("cpu"), criterion=crit, evidential=evid)


def test_simple_programmatic_example(capsys):
    from murmura_b200.examples.simple_programmatic import main
    hist = main(rounds=3)
    out = capsys.readouterr().out
    assert len(hist["round"]) == 3 and hist["mean_accuracy"][-1] > 0.6
    assert re.search(r"Round 3: Mean Accuracy = \d\.\d{4} ± \d\.\d{4}", out)
    assert "=== Round 1/3 ===" in out


def test_two_node_ring_fedavg_nodes_identical():
    torch.manual_seed(0)
    net = _build(_cfg(topology={"type": "ring", "num_nodes": 2}))
    net.train(rounds=1, lr=0.05)
    a, b = net.nodes[0].get_state(), net.nodes[1].get_state()
    assert all(torch.allclose(a[k], b[k]) for k in a)              # each averages {own, other}
    with pytest.raises(ValueError):
        Network(net.nodes, create_topology("ring", 3))


def test_attack_history_and_contract(capsys):
    torch.manual_seed(0)
    cfg = _cfg(attack={"enabled": True, "type": "gaussian", "percentage": 0.25, "params": {"noise_std": 1.0}},
               aggregation={"algorithm": "balance", "params": {"gamma": 0.5}})
    net = _build(cfg)
    hist = net.train(rounds=2, verbose=True, eval_every=1)
    out = capsys.readouterr().out
    assert len(hist["honest_accuracy"]) == 2 and len(hist["compromised_accuracy"]) == 2
    assert re.search(r"  Honest: \d\.\d{4}, Compromised: \d\.\d{4}", out)
    stats = net.get_node_statistics()
    assert set(stats) == {0, 1, 2, 3} and "mean_acceptance_rate" in stats[0]
    byz = next(iter(net.attack.get_compromised_nodes()))
    assert net.attack.is_compromised(byz)
    h2 = _build(cfg).train(rounds=2, eval_every=2)
    assert h2["round"] == [2]


def test_evidential_pipeline_and_uncertainty_line(capsys):
    torch.manual_seed(0)
    cfg = _cfg(model={"factory": "examples.wearables.uci_har", "params": {"input_dim": 561, "num_classes": 6, "hidden_dims": [16]}},
               data={"adapter": "wearables.uci_har", "params": {"data_path": "synthetic", "samples_per_node": 40}},
               aggregation={"algorithm": "evidential_trust", "params": {"trust_threshold": 0.1}})
    crit, evid = build_criterion(cfg)
    assert evid and crit.annealing_epochs == 1 and crit.lambda_weight == 0.1
    hist = _build(cfg).train(rounds=2, verbose=True)
    out = capsys.readouterr().out
    assert len(hist["mean_vacuity"]) == 2
    assert re.search(r"  Uncertainty: Vacuity=\d+\.\d{4}, Entropy=\d+\.\d{4}, Strength=\d+\.\d{2}", out)


@pytest.mark.parametrize("algo,params", [("krum", {"num_compromised": 1}), ("sketchguard", {"sketch_size": 64}),
                                         ("ubar", {"rho": 0.5}), ("balance", {})])
def test_all_aggregators_run(algo, params):
    torch.manual_seed(0)
    cfg = _cfg(topology={"type": "fully", "num_nodes": 5}, aggregation={"algorithm": algo, "params": params},
               attack={"enabled": True, "type": "directed_deviation", "percentage": 0.2})
    hist = _build(cfg).train(rounds=1)
    assert 0.0 <= hist["mean_accuracy"][0] <= 1.0


def test_factories_misc():
    cfg = _cfg(attack={"enabled": True, "type": "topology_liar", "percentage": 0.5, "params": {"model_attack_type": "gaussian"}},
               mobility={"comm_range": 50}, dmtt={})
    atk = build_attack(cfg)
    assert hasattr(atk, "get_false_claims") and atk._model_attack is not None
    assert build_mobility_model(cfg).comm_range == 50 and build_mobility_model(_cfg()) is None
    assert build_attack(_cfg()) is None
    cfg2 = _cfg(model={"factory": "murmura.models.mlp", "params": {"hidden_dims": [4]}})        # reference-style dotted path
    assert build_model_factory(cfg2)().net[0].out_features == 4
    sk = _cfg(aggregation={"algorithm": "sketchguard"})
    agg = build_aggregator_factory(sk, build_model_factory(sk))(0)
    assert agg.model_dim == 784 * 16 + 16 + 16 * 10 + 10 and agg.total_rounds == 2


def test_cli(tmp_path):
    from murmura_b200.cli import app
    runner = CliRunner()
    res = runner.invoke(app, ["list-components", "aggregators"])
    assert res.exit_code == 0 and "evidential_trust" in res.stdout
    assert "b200" in runner.invoke(app, ["list-components", "backends"]).stdout
    assert "Unknown component type" in runner.invoke(app, ["list-components", "nope"]).stdout
    res = runner.invoke(app, ["run", os.path.join(CFG, "basic_fedavg.yaml"), "--device", "cpu", "--quiet"])
    assert res.exit_code == 0 and "Training complete" in res.stdout and "Training Results" in res.stdout
    bad = tmp_path / "bad.yaml"; bad.write_text("experiment: {name: x}\n")
    res = runner.invoke(app, ["run", str(bad)])
    assert res.exit_code == 1 and "Error:" in res.stdout


def test_dmtt_state_math():
    s = DMTTNodeState(0, DMTTConfig(), num_nodes=4)
    assert s.link_reliability(2) == 0.5 and s.topo_trust(2) == pytest.approx(0.5)      # U=sqrt(1/12)<tau_U
    s.update_link_reliability(2, True); assert s.link_reliability(2) == pytest.approx(0.55)
    s.update_link_reliability(2, False); assert s.link_reliability(2) == pytest.approx(0.495)
    s.update_trust(1, d=3, x=0); a, b = 0.9 + 3, 0.9
    assert s._alpha[1] == pytest.approx(a) and s._beta[1] == pytest.approx(b)
    U = math.sqrt(a * b / ((a + b) ** 2 * (a + b + 1)))
    assert s.topo_trust(1) == pytest.approx(a / (a + b) * math.exp(-5.0 * max(0, U - 0.3)))
    s.update_trust(3, d=0, x=50)
    assert s.topo_trust(3) < 0.05
    assert s.model_score(1.0, 0.0) == 1.0 and s.model_score(0.5, 0.8) == pytest.approx(0.2 * 0.65 * math.exp(-0.3))
    assert s.collab_score(1, 1.0) == pytest.approx(0.4 + 0.3 * s.topo_trust(1) + 0.2 * 0.5)
    assert s.top_b([1, 2, 3], {1: 0.9, 3: 0.9}, 2) == [1, 3]
    assert s.top_b([], {}, 3) == []
    assert set(s.state_summary()) == {1, 2, 3}
    assert verify_claim([1, 2, 5], {1, 2}) == (2.0, 1.0)
    big = DMTTNodeState(0, DMTTConfig()); big.update_trust(7, 1, 0); assert big._alpha[7] == pytest.approx(1.9)


def test_messaging_roundtrip(monkeypatch):
    from murmura_b200.distributed.messaging import MsgType, decode, decode_full, encode, pack_obj, pack_state, unpack_obj, unpack_state
    import struct
    frames = encode(MsgType.TOPO_CLAIM, 5, b"xyz", round_idx=9)
    assert decode(frames) == (MsgType.TOPO_CLAIM, 5, b"xyz") and decode_full(frames)[2] == 9
    legacy = [struct.pack("!Bi", 0, 3), b"p"]
    assert decode_full(legacy) == (MsgType.MODEL_STATE, 3, -1, b"p")
    st = {"w": torch.arange(4.0)}
    assert torch.equal(unpack_state(pack_state(st))["w"], st["w"]) and unpack_obj(pack_obj({"a": 1})) == {"a": 1}
    from murmura_b200.distributed.endpoints import Endpoints
    from murmura_b200.config.schema import DistributedConfig
    ep = Endpoints(DistributedConfig(transport="tcp", node_hosts={2: "10.0.0.2"}), 3, "r1")
    assert ep.node_pull_connect(2) == "tcp://10.0.0.2:5552" and ep.node_pull_bind(1) == "tcp://127.0.0.1:5551"
    assert len(frames[0]) == 5 and struct.unpack("!Bi", frames[0]) == (2, 5)          # frame 0 stays the reference's 5-byte header
    assert decode_full([struct.pack("!Bii", 1, 4, 7), b"q"]) == (MsgType.METRICS, 4, 7, b"q")   # round-1 header still accepted
    monkeypatch.setenv("MURMURA_BIND_HOST", "0.0.0.0")
    assert Endpoints(DistributedConfig(transport="tcp"), 3, "r1").node_pull_bind(1) == "tcp://0.0.0.0:5551"
    assert Endpoints(DistributedConfig(), 3, "r1").monitor_pull_bind() == "ipc:///tmp/murmura/r1/monitor_pull"


@pytest.mark.timeout(180)
def test_zmq_distributed_backend(tmp_path):
    """3 node processes + monitor over ipc:// — the reference's 'multi-node without a cluster' test bed."""
    cfg = load_config(os.path.join(CFG, "distributed_fedavg.yaml"))
    data = cfg.model_dump(); data["distributed"]["ipc_dir"] = str(tmp_path / "ipc")
    import yaml
    p = tmp_path / "dist.yaml"; p.write_text(yaml.safe_dump(data))
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    code = ("from murmura_b200.distributed import DistributedRunner; import sys;"
            f"h = DistributedRunner('{p}').run(verbose=True); print('ROUNDS', h['round'], h['mean_accuracy'])")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=170)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ROUNDS [1, 2]" in res.stdout and "[Monitor] Round 2 (3/3 nodes)" in res.stdout


@pytest.mark.timeout(240)
def test_zmq_dmtt_backend(tmp_path):
    cfg = load_config(os.path.join(CFG, "distributed_fedavg.yaml")).model_dump()
    cfg["distributed"].update(ipc_dir=str(tmp_path / "ipc"), round_duration_s=6.0)
    cfg["topology"]["num_nodes"] = 4
    cfg["mobility"] = {"comm_range": 60.0, "seed": 1}
    cfg["dmtt"] = {"budget_B": 2}
    cfg["attack"] = {"enabled": True, "type": "topology_liar", "percentage": 0.25, "params": {"model_attack_type": "gaussian", "noise_std": 0.1}}
    import yaml
    p = tmp_path / "dmtt.yaml"; p.write_text(yaml.safe_dump(cfg))
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    code = f"from murmura_b200.distributed import DistributedRunner; h = DistributedRunner('{p}').run(verbose=True); print('ROUNDS', h['round'])"
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=230)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ROUNDS [1, 2]" in res.stdout


def test_experiment_harness(tmp_path):
    """Config generator + resumable suite runner that scrapes the stdout contract."""
    import json
    env = dict(os.environ, PYTHONPATH=ROOT)
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "generate_configs.py"), "--out", str(tmp_path / "cfgs"),
                          "--backend", "simulation", "--rounds", "2"], capture_output=True, text=True, env=env)
    assert gen.returncode == 0 and "wrote" in gen.stdout and "covering 314 experiment slots" in gen.stdout
    files = sorted(f for f in os.listdir(tmp_path / "cfgs") if f.endswith(".yaml"))
    assert len(files) > 250 and all(load_config(tmp_path / "cfgs" / f) for f in files)
    index = json.load(open(tmp_path / "cfgs" / "index.json"))
    assert {k: len(v) for k, v in index.items()} == {"baseline": 18, "heterogeneity": 54, "attacks": 108, "topologies": 48,
                                                      "ablation": 51, "dmtt": 3, "scenarios": 32}
    assert all(f in files for fam in index.values() for f in fam.values())
    one = tmp_path / "one"; one.mkdir()
    (one / "a.yaml").write_text((tmp_path / "cfgs" / "ppg_dalia__fedavg__none0__ring__a0.5.yaml").read_text())
    res = tmp_path / "res.json"
    run = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "run_suite.py"), str(one), "--results", str(res), "--device", "cpu"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert run.returncode == 0, run.stderr[-1000:]
    rec = json.load(open(res))["a"]
    assert rec["status"] == "ok" and len(rec["rounds"]) == 2 and "vacuity" in rec["rounds"][-1]
    again = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "run_suite.py"), str(one), "--results", str(res), "--device", "cpu"],
                           capture_output=True, text=True, env=env, timeout=300)
    assert "1/1 experiments have results" in again.stdout and "ok " not in again.stdout.split("1/1")[0]      # resumed: nothing re-run


def test_compat_alias_and_extra_configs():
    code = ("import murmura_b200.compat as c; c.install_alias(); import murmura; from murmura import Network, Config; "
            "from murmura.aggregation import KrumAggregator; from murmura.topology import create_topology; "
            "from murmura.utils import set_seed; print('ALIAS_OK', murmura.__name__)")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT))
    assert res.returncode == 0 and "ALIAS_OK murmura_b200" in res.stdout, res.stderr[-500:]
    for name in ("uci_har_dirichlet", "pamap2_dirichlet", "ubar_attack"):
        cfg = load_config(os.path.join(CFG, name + ".yaml"))
        cfg.experiment.rounds = 1
        cfg.data.params["samples_per_node"] = 40
        if "hidden_dims" not in cfg.model.params:
            cfg.model.params["hidden_dims"] = [16]
        torch.manual_seed(0)
        hist = _build(cfg).train(rounds=1, lr=0.01)
        assert len(hist["round"]) == 1


def test_generate_figures(tmp_path):
    import json
    res = {"uci_har__fedavg__none0__ring__a0.5": {"status": "ok", "final_accuracy": 0.9, "final_std": 0.01, "convergence_round": 3,
                                                   "rounds": [{"round": 1, "mean_accuracy": 0.5, "std_accuracy": 0.1}, {"round": 2, "mean_accuracy": 0.9, "std_accuracy": 0.01}]},
           "bad": {"status": "failed", "rounds": []}}
    (tmp_path / "r.json").write_text(json.dumps(res))
    run = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "generate_figures.py"), str(tmp_path / "r.json"), "--out", str(tmp_path / "fig")],
                         capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT))
    assert run.returncode == 0, run.stderr
    md = (tmp_path / "fig" / "summary.md").read_text()
    (tmp_path / "index.json").write_text(json.dumps({"topologies": {"uci_har/fedavg_ring": "uci_har__fedavg__none0__ring__a0.5.yaml",
                                                                     "uci_har/fedavg_fully": "missing.yaml"}}))
    run2 = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "generate_figures.py"), str(tmp_path / "r.json"), "--out", str(tmp_path / "fig"),
                           "--index", str(tmp_path / "index.json")], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=ROOT))
    assert run2.returncode == 0, run2.stderr
    fam = (tmp_path / "fig" / "families.md").read_text()
    assert "## topologies (2 experiments)" in fam and "0.9000 ± 0.0100" in fam and "1/2 done" in fam
    assert "uci_har__fedavg__none0__ring__a0.5" in md and "0.9000" in md and "bad" not in md


def test_public_api_parity_with_reference_package():
    """Every public name / method / parameter of the unmodified reference (baseline/_ref) exists in the same-named module here."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "baseline", "_ref", "murmura")):
        import pytest
        pytest.skip("reference install (baseline/_ref) not present")
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "api_parity.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "## Missing (0)" in out.stdout, out.stdout[-3000:]


def test_paper_figures_from_suite_results(tmp_path):
    """The six paper figures (reference experiments/paper/generate_figures.py:109-513) from a run_suite results file."""
    import importlib.util
    import xml.dom.minidom
    spec = importlib.util.spec_from_file_location("paper_figures", os.path.join(ROOT, "experiments", "paper_figures.py"))
    pf = importlib.util.module_from_spec(spec); spec.loader.exec_module(pf)
    index = {"heterogeneity": {}, "ablation": {}}
    res = {}
    for ds, base in (("uci_har", 0.80), ("pamap2", 0.70)):
        for ai, algo in enumerate(pf.ALGORITHMS):
            for al, bump in (("01", 0.0), ("05", 0.06), ("10", 0.10)):
                key = f"{ds}__{algo}__a{al}"
                index["heterogeneity"][f"{ds}/{algo}_alpha{al}"] = key + ".yaml"
                res[key] = {"status": "ok", "final_accuracy": base + bump + 0.01 * ai, "final_std": 0.05, "convergence_round": 30 - 4 * ai, "rounds": []}
        for p in ("accuracy_weight_03", "self_weight_05", "trust_threshold_01", "vacuity_threshold_05"):
            key = f"{ds}__evidential_trust__abl_{p}"
            index["ablation"][f"{ds}/{p}"] = key + ".yaml"
            res[key] = {"status": "ok", "final_accuracy": 0.9, "final_std": 0.02, "convergence_round": 9, "rounds": []}
    res["uci_har__ubar__a10"]["status"] = "error"                      # a failed run must not break the averages
    data = pf.generate(res, index, str(tmp_path))
    for name in ("fig1_noniid_robustness", "fig2_degradation", "fig3_personalization", "fig4_convergence", "fig5_ablation", "fig6_combined_summary"):
        xml.dom.minidom.parse(str(tmp_path / f"{name}.svg"))
    assert data["datasets"] == ["pamap2", "uci_har"]
    assert abs(data["fig2_degradation"]["fedavg"] - 10.0) < 1e-6                     # (α=1.0) − (α=0.1) in points
    assert data["fig4_convergence"]["evidential_trust"] == 14.0 and data["fig4_convergence"]["fedavg"] == 30.0
    assert abs(data["fig1_noniid_robustness"]["fedavg"]["01"]["mean_acc"] - 75.0) < 1e-6
    assert data["fig5_ablation"]["self_weight"]["runs"] == 2
