"""BASELINE config 1 on the CPU: 2-node ring FedAvg, 784-200-10 MLP, synthetic MNIST-shaped shards, simulation backend.

    python scripts/cpu_config1.py [--rounds 20] [--arm ours|reference|both] [--nodes 2]

Both arms run the same config through their own public API on CPU tensors; wall clock, 2 warm-up rounds excluded.
"""
import argparse, contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(arm, rounds, nodes, algo, topo="ring"):
    import torch
    d = {"experiment": {"name": "cfg1", "seed": 42, "rounds": rounds + 2}, "topology": {"type": topo, "num_nodes": nodes, "p": 0.3, "k": 4, "seed": 7},
         "aggregation": {"algorithm": algo, "params": {}}, "training": {"local_epochs": 1, "batch_size": 64, "lr": 0.01}}
    if arm == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        from murmura.config import Config
        from murmura.core.network import Network
        from murmura.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
        from murmura.utils.seed import set_seed
        d["data"] = {"adapter": "baseline.ref_workloads.SyntheticRefAdapter",
                     "params": {"name": "mnist", "num_nodes": nodes, "samples_per_node": 512, "alpha": 0.5, "seed": 42}}
        d["model"] = {"factory": "baseline.ref_workloads.mlp", "params": {}}
    else:
        from murmura_b200.config import Config
        from murmura_b200.core.network import Network
        from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
        from murmura_b200.utils.seed import set_seed
        d["data"] = {"adapter": "synthetic.mnist", "params": {"samples_per_node": 512, "partition_method": "dirichlet", "alpha": 0.5}}
        d["model"] = {"factory": "models.mlp", "params": {"input_dim": 784, "hidden_dims": [200], "num_classes": 10}}
        d["backend"] = "simulation"
    cfg = Config(**d)
    set_seed(42)
    dev = torch.device("cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
        net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf, dev), device=dev)
        net.train(rounds=2, local_epochs=1, lr=0.01)
        t0 = time.perf_counter()
        h = net.train(rounds=rounds, local_epochs=1, lr=0.01)
        dt = time.perf_counter() - t0
    return {"arm": arm, "nodes": nodes, "algo": algo, "topology": topo, "rounds": rounds, "rounds_per_s": round(rounds / dt, 2), "ms_per_round": round(dt / rounds * 1e3, 1),
            "final_acc": round(float(h["mean_accuracy"][-1]), 4), "threads": torch.get_num_threads()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20); ap.add_argument("--arm", default="both"); ap.add_argument("--nodes", type=int, default=2)
    ap.add_argument("--algo", default="fedavg"); ap.add_argument("--topo", default="ring")
    a = ap.parse_args()
    for arm in (("ours", "reference") if a.arm == "both" else (a.arm,)):
        if a.arm == "both":            # separate interpreters: both packages patch global RNG / thread state
            import subprocess
            print(subprocess.run([sys.executable, __file__, "--rounds", str(a.rounds), "--arm", arm, "--nodes", str(a.nodes), "--algo", a.algo, "--topo", a.topo],
                                 capture_output=True, text=True).stdout.strip())
        else:
            print(json.dumps(run(arm, a.rounds, a.nodes, a.algo, a.topo)))
