"""Generate the whole experiment universe of the reference as YAML files for this framework.

The reference ships 314 hand-expanded YAMLs (``experiments/paper/{uci_har,pamap2,ppg_dalia,heterogeneity,attacks,topologies,
ablation,dmtt}`` written by ``experiments/paper/generate_all_configs.py`` plus the 32 scenario files of ``experiments/configs``).
Here the universe is *declared* as seven families of experiment slots; slots with identical settings (e.g. "FedAvg, α = 0.5,
fully connected" appears in the heterogeneity, attack and topology families) share ONE config file, and ``index.json`` maps
every slot of every family to its file so result tables / figures can be produced per family.

    python experiments/generate_configs.py --out experiments/configs [--backend b200|simulation|distributed] [--rounds 50]
                                           [--data-root wearables_datasets] [--families baseline attacks …]

Datasets are the synthetic generators of the same shapes unless ``--data-root`` points at the real data.
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Any, Dict, List, Optional, Tuple

import yaml

DATASETS = {
    "uci_har": {"nodes": 10, "input_dim": 561, "num_classes": 6},
    "pamap2": {"nodes": 9, "input_dim": 4000, "num_classes": 12},
    "ppg_dalia": {"nodes": 15, "input_dim": 192, "num_classes": 7},
}
EVIDENTIAL = {"vacuity_threshold": 0.5, "accuracy_weight": 0.7, "trust_threshold": 0.1, "self_weight": 0.6}
AGGREGATORS: Dict[str, Dict[str, Any]] = {
    "fedavg": {},
    "krum": {"num_compromised": 1},
    "balance": {"gamma": 0.5, "kappa": 1.0, "alpha": 0.5, "min_neighbors": 1},
    "sketchguard": {"sketch_size": 1000, "gamma": 0.5, "kappa": 1.0, "alpha": 0.5},
    "ubar": {"rho": 0.5, "alpha": 0.5, "min_neighbors": 1},
    "evidential_trust": dict(EVIDENTIAL),
}
ATTACKS = {"none": None, "gaussian": {"type": "gaussian", "params": {"noise_std": 10.0}},
           "directed_deviation": {"type": "directed_deviation", "params": {"lambda_param": -5.0}}}
TOPOLOGIES = {"fully": {"type": "fully"}, "ring": {"type": "ring"}, "erdos": {"type": "erdos", "p": 0.3},
              "k_regular": {"type": "k-regular", "k": 4}}
ABLATION = {"accuracy_weight": (0.3, 0.5, 0.7, 0.9), "self_weight": (0.3, 0.5, 0.6, 0.7, 0.9),
            "trust_threshold": (0.05, 0.1, 0.2, 0.3), "vacuity_threshold": (0.3, 0.5, 0.7, 0.9)}
MOBILITY = {"area_size": 100.0, "comm_range": 40.0, "max_speed": 8.0, "seed": 42, "ensure_connected": True}
DMTT = {"budget_B": 5, "rho": 0.1, "lambda_forget": 0.9, "w_d": 1.0, "w_c": 0.5, "w_x": 1.0, "tau_U": 0.3, "eta": 5.0, "w_a": 0.7,
        "tau_u": 0.5, "lambda1": 0.4, "lambda2": 0.3, "lambda3": 0.2, "lambda4": 0.1}

Slot = Tuple[str, str, Dict[str, Any]]          # (family, slot name, config dict)


class Builder:
    def __init__(self, backend: str, rounds: int, data_root: Optional[str]):
        self.backend, self.rounds, self.data_root = backend, rounds, data_root

    def config(self, name: str, dataset: str, algo: str, *, agg_params: Optional[Dict[str, Any]] = None, attack: str = "none",
               pct: float = 0.0, attack_params: Optional[Dict[str, Any]] = None, topo: str = "fully", alpha: float = 0.5,
               lr: float = 0.01, extra: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        d = DATASETS[dataset]
        data_params: Dict[str, Any] = {"partition_method": "dirichlet", "alpha": alpha}
        data_params["data_path"] = os.path.join(self.data_root, dataset) if self.data_root else "synthetic"
        if not self.data_root:
            data_params["samples_per_node"] = 512
        cfg: Dict[str, Any] = {
            "experiment": {"name": name, "seed": 42, "rounds": self.rounds},
            "topology": {**TOPOLOGIES[topo], "num_nodes": d["nodes"], "seed": 12345},
            "aggregation": {"algorithm": algo, "params": dict(AGGREGATORS[algo] if agg_params is None else agg_params)},
            "training": {"local_epochs": 2, "batch_size": 32, "lr": lr},
            "data": {"adapter": f"wearables.{dataset}", "params": data_params},
            "model": {"factory": f"examples.wearables.{dataset}", "params": {"input_dim": d["input_dim"], "num_classes": d["num_classes"]}},
            "backend": self.backend,
        }
        if ATTACKS.get(attack):
            a = ATTACKS[attack]
            cfg["attack"] = {"enabled": True, "type": a["type"], "percentage": pct, "params": dict(attack_params or a["params"])}
        if extra:
            cfg.update(extra)
        return cfg

    def grid_name(self, dataset, algo, attack, pct, topo, alpha) -> str:
        return f"{dataset}__{algo}__{attack}{int(round(pct * 100))}__{topo}__a{alpha}"

    def grid(self, dataset, algo, attack="none", pct=0.0, topo="fully", alpha=0.5) -> Dict[str, Any]:
        return self.config(self.grid_name(dataset, algo, attack, pct, topo, alpha), dataset, algo, attack=attack, pct=pct, topo=topo, alpha=alpha)

    # ---- the seven families ------------------------------------------------------------------------------------------------
    def baseline(self) -> List[Slot]:                      # paper/{dataset}: six aggregators, α = 0.1
        return [("baseline", f"{ds}/{algo}", self.grid(ds, algo, alpha=0.1)) for ds in DATASETS for algo in AGGREGATORS]

    def heterogeneity(self) -> List[Slot]:                 # paper/heterogeneity: α ∈ {0.1, 0.5, 1.0}
        return [("heterogeneity", f"{ds}/{algo}_alpha{str(a).replace('.', '')}", self.grid(ds, algo, alpha=a))
                for ds in DATASETS for algo in AGGREGATORS for a in (0.1, 0.5, 1.0)]

    def attacks(self) -> List[Slot]:                       # paper/attacks: 2 attacks × {10, 20, 30} %
        return [("attacks", f"{ds}/{algo}_{atk}_{int(p * 100)}pct", self.grid(ds, algo, atk, p))
                for ds in DATASETS for algo in AGGREGATORS for atk in ("gaussian", "directed_deviation") for p in (0.1, 0.2, 0.3)]

    def topologies(self) -> List[Slot]:                    # paper/topologies: 4 aggregators × 4 graphs
        return [("topologies", f"{ds}/{algo}_{t}", self.grid(ds, algo, topo=t))
                for ds in DATASETS for algo in ("fedavg", "krum", "sketchguard", "evidential_trust") for t in TOPOLOGIES]

    def ablation(self) -> List[Slot]:                      # paper/ablation: EvidentialTrust hyper-parameters, one at a time
        out = []
        for ds in DATASETS:
            for key, values in ABLATION.items():
                for v in values:
                    tag = f"{key}_{str(v).replace('.', '')}"
                    params = dict(EVIDENTIAL, **{key: v})
                    name = f"{ds}__evidential_trust__abl_{tag}"
                    out.append(("ablation", f"{ds}/{tag}", self.config(name, ds, "evidential_trust", agg_params=params)))
        return out

    def dmtt(self) -> List[Slot]:                          # paper/dmtt: static vs dynamic vs dynamic + trust protocol, 30 % liars
        liar = {"model_attack_type": "gaussian", "noise_std": 10.0}
        atk = {"enabled": True, "type": "topology_liar", "percentage": 0.3, "params": liar}
        dist = {"transport": "ipc", "round_duration_s": 120.0, "startup_grace_s": 8.0}
        base = dict(dataset="uci_har", algo="fedavg")
        out = []
        for tag, extra in (("01_baseline_static", {}), ("02_dynamic_no_trust", {"mobility": MOBILITY}),
                           ("03_dmtt", {"mobility": MOBILITY, "dmtt": DMTT})):
            cfg = self.config(f"dmtt__{tag}", **base, extra=dict(extra))
            cfg["attack"] = dict(atk)
            if self.backend == "distributed":
                cfg["distributed"] = dict(dist)
            elif extra and self.backend == "simulation":
                cfg["backend"] = "distributed"             # like the reference, mobility / DMTT need a backend that honours them
                cfg["distributed"] = dict(dist)
            out.append(("dmtt", tag, cfg))
        return out

    def scenarios(self) -> List[Slot]:                     # experiments/configs: exp1 … exp4 on UCI-HAR
        ds, out = "uci_har", []
        short = {"evidential_trust": "evidential"}

        def add(tag, cfg):
            out.append(("scenarios", tag, cfg))

        for algo in AGGREGATORS:
            s = short.get(algo, algo)
            add(f"exp1_baseline_{s}", self.grid(ds, algo))
            add(f"exp2_attack20_{s}", self.grid(ds, algo, "gaussian", 0.2))
            add(f"exp4_personalization_{s}", self.grid(ds, algo, alpha=0.1))
        mild = dict(EVIDENTIAL, accuracy_weight=0.5, self_weight=0.5, use_adaptive_trust=True, trust_momentum=0.7,
                    use_tightening_threshold=True, gamma=0.5, kappa=1.0, max_eval_samples=100)
        add("exp2_attack20_mild_evidential", self.config(f"{ds}__scenario__attack20_mild_evidential", ds, "evidential_trust", agg_params=mild,
                                                         attack="gaussian", pct=0.2, attack_params={"noise_std": 1.0}, lr=0.001))
        for algo in ("evidential_trust", "krum"):
            s = short.get(algo, algo)
            for pct in (0.3, 0.4):
                params = dict(AGGREGATORS[algo], num_compromised=int(pct * 10)) if algo == "krum" else None
                cfg = self.config(f"{ds}__{algo}__gaussian{int(pct * 100)}__fully__a0.5" + ("__f" + str(int(pct * 10)) if algo == "krum" else ""),
                                  ds, algo, agg_params=params, attack="gaussian", pct=pct)
                add(f"exp2_attack{int(pct * 100)}_{s}", cfg)
            add(f"exp2_directed_{s}", self.grid(ds, algo, "directed_deviation", 0.2))
            add(f"exp3_heterog_extreme_attack_{s}", self.config(f"{ds}__{algo}__gaussian20__fully__a0.1", ds, algo, attack="gaussian", pct=0.2, alpha=0.1))
        for algo in ("evidential_trust", "fedavg"):
            s = short.get(algo, algo)
            add(f"exp3_heterog_extreme_{s}", self.grid(ds, algo, alpha=0.1))
            add(f"exp3_heterog_mild_{s}", self.grid(ds, algo, alpha=1.0))
        local = dict(EVIDENTIAL, accuracy_weight=0.5, trust_threshold=1.0, self_weight=1.0, use_adaptive_trust=False, use_tightening_threshold=False)
        add("exp4_personalization_local_only", self.config(f"{ds}__scenario__local_only", ds, "evidential_trust", agg_params=local, topo="ring", alpha=0.1))
        return out


FAMILIES = ("baseline", "heterogeneity", "attacks", "topologies", "ablation", "dmtt", "scenarios")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="experiments/configs"); ap.add_argument("--backend", default="b200")
    ap.add_argument("--rounds", type=int, default=50); ap.add_argument("--data-root", default=None)
    ap.add_argument("--families", nargs="*", default=list(FAMILIES), choices=FAMILIES)
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    b = Builder(args.backend, args.rounds, args.data_root)
    index: Dict[str, Dict[str, str]] = {}
    written: Dict[str, Dict[str, Any]] = {}
    slots = 0
    for fam in args.families:
        for family, slot, cfg in getattr(b, fam)():
            slots += 1
            name = cfg["experiment"]["name"]
            if name in written and written[name] != cfg:
                raise SystemExit(f"two different experiments share the name {name}")
            if name not in written:
                written[name] = cfg
                with open(os.path.join(args.out, name + ".yaml"), "w") as fh:
                    yaml.safe_dump(cfg, fh, sort_keys=False)
            index.setdefault(family, {})[slot] = name + ".yaml"
    with open(os.path.join(args.out, "index.json"), "w") as fh:
        json.dump(index, fh, indent=1)
    per = ", ".join(f"{k} {len(v)}" for k, v in index.items())
    print(f"wrote {len(written)} configs to {args.out} covering {slots} experiment slots ({per}); slot → file map in index.json")


if __name__ == "__main__":
    main()
