"""Generate the experiment grid (aggregators × attacks × topologies × heterogeneity) as YAML files.

Counterpart of the reference's ``experiments/paper/generate_all_configs.py`` (its 282-experiment grid: six aggregators on
three wearable datasets under no attack / Gaussian / directed deviation at 10-30 %, four topologies, Dirichlet α sweeps).
Datasets are the synthetic generators of the same shapes unless ``--data-root`` points at real data.

    python experiments/generate_configs.py --out experiments/configs [--backend b200|simulation] [--rounds 50]
"""
from __future__ import annotations

import argparse
import os

import yaml

DATASETS = {
    "uci_har": {"nodes": 10, "input_dim": 561, "num_classes": 6},
    "pamap2": {"nodes": 9, "input_dim": 4000, "num_classes": 12},
    "ppg_dalia": {"nodes": 15, "input_dim": 192, "num_classes": 7},
}
AGGREGATORS = {
    "fedavg": {},
    "krum": {"num_compromised": 1},
    "balance": {"gamma": 0.5, "kappa": 1.0, "alpha": 0.5, "min_neighbors": 1},
    "sketchguard": {"sketch_size": 1000, "gamma": 0.5, "kappa": 1.0, "alpha": 0.5},
    "ubar": {"rho": 0.5, "alpha": 0.5, "min_neighbors": 1},
    "evidential_trust": {"vacuity_threshold": 0.5, "accuracy_weight": 0.7, "trust_threshold": 0.1, "self_weight": 0.6},
}
ATTACKS = {"none": None, "gaussian": {"type": "gaussian", "params": {"noise_std": 10.0}},
           "directed_deviation": {"type": "directed_deviation", "params": {"lambda_param": -5.0}}}
TOPOLOGIES = {"fully": {"type": "fully"}, "ring": {"type": "ring"}, "erdos": {"type": "erdos", "p": 0.3},
              "k_regular": {"type": "k-regular", "k": 4}}


def make(dataset: str, algo: str, attack: str, pct: float, topo: str, alpha: float, backend: str, rounds: int, data_root: str | None):
    d = DATASETS[dataset]
    data_params = {"partition_method": "dirichlet", "alpha": alpha}
    data_params["data_path"] = os.path.join(data_root, dataset) if data_root else "synthetic"
    if not data_root:
        data_params["samples_per_node"] = 512
    cfg = {
        "experiment": {"name": f"{dataset}__{algo}__{attack}{int(pct * 100)}__{topo}__a{alpha}", "seed": 42, "rounds": rounds},
        "topology": {**TOPOLOGIES[topo], "num_nodes": d["nodes"], "seed": 12345},
        "aggregation": {"algorithm": algo, "params": AGGREGATORS[algo]},
        "training": {"local_epochs": 2, "batch_size": 32, "lr": 0.01},
        "data": {"adapter": f"wearables.{dataset}", "params": data_params},
        "model": {"factory": f"examples.wearables.{dataset}", "params": {"input_dim": d["input_dim"], "num_classes": d["num_classes"]}},
        "backend": backend,
    }
    if ATTACKS[attack]:
        cfg["attack"] = {"enabled": True, "percentage": pct, **ATTACKS[attack]}
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="experiments/configs"); ap.add_argument("--backend", default="b200")
    ap.add_argument("--rounds", type=int, default=50); ap.add_argument("--data-root", default=None)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    grid = []
    for ds in DATASETS:
        for algo in AGGREGATORS:
            grid.append((ds, algo, "none", 0.0, "fully", 0.1))                       # heterogeneity α=0.1 (paper table)
            grid.append((ds, algo, "none", 0.0, "fully", 0.5))
            for atk in ("gaussian", "directed_deviation"):
                for pct in (0.1, 0.2, 0.3):
                    grid.append((ds, algo, atk, pct, "fully", 0.5))
        for topo in TOPOLOGIES:
            grid.append((ds, "fedavg", "none", 0.0, topo, 0.5))
    seen = set()
    for g in grid:
        cfg = make(*g, args.backend, args.rounds, args.data_root)
        name = cfg["experiment"]["name"]
        if name in seen:
            continue
        seen.add(name)
        with open(os.path.join(args.out, name + ".yaml"), "w") as fh:
            yaml.safe_dump(cfg, fh, sort_keys=False)
    print(f"wrote {len(seen)} configs to {args.out}")


if __name__ == "__main__":
    main()
