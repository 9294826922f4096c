"""Accuracy-parity study: the same attack × aggregator grid on (a) the UNMODIFIED reference package, (b) this repo's
simulation backend and (c) this repo's B200 backend — same synthetic UCI-HAR-shaped shards (bit-identical tensors and
Dirichlet partitions), same evidential HAR MLP, same hyper-parameters as the reference's paper configs
(10 nodes, fully connected, 2 local epochs, batch 32, lr 0.01, seed 42).

    python experiments/parity_study.py --arm reference|simulation|b200 --rounds 30 --out experiments/parity_<arm>.json
"""
from __future__ import annotations

import argparse, contextlib, io, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

AGG = {"fedavg": {}, "krum": {"num_compromised": 1}, "balance": {"gamma": 0.5, "kappa": 1.0, "alpha": 0.5, "min_neighbors": 1},
       "sketchguard": {"sketch_size": 1000, "gamma": 0.5, "kappa": 1.0, "alpha": 0.5}, "ubar": {"rho": 0.5, "alpha": 0.5, "min_neighbors": 1},
       "evidential_trust": {"vacuity_threshold": 0.5, "accuracy_weight": 0.7, "trust_threshold": 0.1, "self_weight": 0.6}}
ATTACKS = {"none": None, "gaussian30": {"type": "gaussian", "percentage": 0.3, "params": {"noise_std": 10.0}},
           "directed30": {"type": "directed_deviation", "percentage": 0.3, "params": {"lambda_param": -5.0}}}
N, SAMPLES, ALPHA = 10, 256, 0.5


def config_dict(arm, algo, attack, rounds):
    d = {"experiment": {"name": f"{algo}-{attack}", "seed": 42, "rounds": rounds},
         "topology": {"type": "fully", "num_nodes": N}, "aggregation": {"algorithm": algo, "params": AGG[algo]},
         "training": {"local_epochs": 2, "batch_size": 32, "lr": 0.01},
         "model": {"factory": "examples.wearables.uci_har", "params": {"input_dim": 561, "num_classes": 6}}}
    if ATTACKS[attack]:
        d["attack"] = {"enabled": True, **ATTACKS[attack]}
    if arm == "reference":
        d["data"] = {"adapter": "baseline.ref_workloads.SyntheticRefAdapter",
                     "params": {"name": "uci_har", "num_nodes": N, "samples_per_node": SAMPLES, "alpha": ALPHA, "seed": 42}}
    else:
        d["data"] = {"adapter": "synthetic.uci_har", "params": {"samples_per_node": SAMPLES, "partition_method": "dirichlet", "alpha": ALPHA}}
        d["backend"] = "b200" if arm == "b200" else "simulation"
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", required=True, choices=["reference", "simulation", "b200"])
    ap.add_argument("--rounds", type=int, default=30); ap.add_argument("--out", default=None); ap.add_argument("--device", default=None)
    args = ap.parse_args()
    import torch
    if args.arm == "reference":
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import murmura as pkg
        from murmura import Network
        from murmura.config import Config
        from murmura.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory
        from murmura.utils.seed import set_seed
    else:
        import murmura_b200 as pkg
        from murmura_b200 import Network
        from murmura_b200.config import Config
        from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory
        from murmura_b200.utils.seed import set_seed
    device = torch.device(args.device or ("cuda" if args.arm == "b200" else "cpu"))
    out = args.out or os.path.join(ROOT, "experiments", f"parity_{args.arm}.json")
    results = json.load(open(out)) if os.path.exists(out) else {}
    for attack in ATTACKS:
        for algo in AGG:
            key = f"{algo}__{attack}"
            if key in results:
                continue
            cfg = Config(**config_dict(args.arm, algo, attack, args.rounds))
            set_seed(42)
            t0 = time.time()
            with contextlib.redirect_stdout(io.StringIO()):
                adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg); crit, evid = build_criterion(cfg)
                net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf, device), device=device,
                                          criterion=crit, evidential=evid)
                h = net.train(rounds=args.rounds, local_epochs=2, lr=0.01)
            if hasattr(net, "close"):
                net.close()
            results[key] = {"final_acc": float(h["mean_accuracy"][-1]), "std": float(h["std_accuracy"][-1]),
                            "honest": float(h["honest_accuracy"][-1]) if h["honest_accuracy"] else None,
                            "compromised": float(h["compromised_accuracy"][-1]) if h["compromised_accuracy"] else None,
                            "vacuity": float(h["mean_vacuity"][-1]) if h["mean_vacuity"] else None,
                            "conv_round": next((r for r, a in zip(h["round"], h["mean_accuracy"]) if a >= 0.8), None),
                            "seconds": round(time.time() - t0, 2)}
            json.dump(results, open(out, "w"), indent=1)
            print(f"{args.arm:10s} {key:32s} acc={results[key]['final_acc']:.4f} honest={results[key]['honest']} {results[key]['seconds']}s", flush=True)


if __name__ == "__main__":
    main()
