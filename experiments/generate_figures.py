"""Tables / figures from a results JSON written by ``run_suite.py`` (counterpart of the reference's
``experiments/paper/generate_figures.py``).  Always writes Markdown tables; PNG plots only when matplotlib is importable.

    python experiments/generate_figures.py experiments/results.json --out experiments/figures [--index experiments/configs/index.json]

With ``--index`` (the slot → file map written by ``generate_configs.py``) it also writes ``families.md``: one table per experiment
family (baseline, heterogeneity, attacks, topologies, ablation, dmtt, scenarios) in the layout of the reference's RESULTS_SUMMARY.
"""
from __future__ import annotations

import argparse
import json
import os
from collections import defaultdict


def family_tables(res: dict, index: dict, path: str) -> None:
    """One Markdown table per family: rows = slots (``dataset/variant``), columns = final accuracy ± std, honest accuracy,
    convergence round.  Slots whose experiment has not run yet are listed as ``-`` so coverage is visible."""
    with open(path, "w") as fh:
        for family, slots in index.items():
            done = 0
            fh.write(f"## {family} ({len(slots)} experiments)\n\n| slot | config | final acc ± std | honest | convergence round |\n|---|---|---|---|---|\n")
            for slot, fname in sorted(slots.items()):
                r = res.get(os.path.splitext(fname)[0])
                if r and r.get("status") == "ok":
                    done += 1
                    honest = r.get("final_honest_accuracy")
                    fh.write(f"| {slot} | {fname} | {r['final_accuracy']:.4f} ± {r['final_std']:.4f} | "
                             f"{'-' if honest is None else format(honest, '.4f')} | {r.get('convergence_round') or 'never'} |\n")
                else:
                    fh.write(f"| {slot} | {fname} | - | - | - |\n")
            fh.write(f"\n{done}/{len(slots)} done\n\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("results"); ap.add_argument("--out", default="experiments/figures"); ap.add_argument("--index", default=None)
    args = ap.parse_args()
    res = json.load(open(args.results))
    os.makedirs(args.out, exist_ok=True)
    groups = defaultdict(list)
    for key, rec in sorted(res.items()):
        if rec.get("status") != "ok":
            continue
        parts = key.split("__")
        groups[parts[0] if len(parts) > 1 else "all"].append((key, rec))
    with open(os.path.join(args.out, "summary.md"), "w") as fh:
        for ds, items in groups.items():
            fh.write(f"## {ds}\n\n| experiment | final acc | std | honest | compromised | convergence round (≥80 %) | rounds/s |\n|---|---|---|---|---|---|---|\n")
            for key, r in items:
                fmt = lambda v: "-" if v is None else f"{v:.4f}"
                fh.write(f"| {key} | {fmt(r.get('final_accuracy'))} | {fmt(r.get('final_std'))} | {fmt(r.get('final_honest_accuracy'))} | "
                         f"{fmt(r.get('final_compromised_accuracy'))} | {r.get('convergence_round') or 'never'} | {r.get('rounds_per_s', '-')} |\n")
            fh.write("\n")
    if args.index and os.path.exists(args.index):
        family_tables(res, json.load(open(args.index)), os.path.join(args.out, "families.md"))
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        print(f"matplotlib not available: wrote {args.out}/summary.md only")
        return
    for ds, items in groups.items():
        plt.figure(figsize=(7, 4))
        for key, r in items:
            plt.plot([x["round"] for x in r["rounds"]], [x["mean_accuracy"] for x in r["rounds"]], label=key.split("__", 1)[-1][:40])
        plt.xlabel("round"); plt.ylabel("mean accuracy"); plt.title(ds); plt.legend(fontsize=5); plt.tight_layout()
        plt.savefig(os.path.join(args.out, f"{ds}_accuracy.png"), dpi=150); plt.close()
    print(f"wrote figures to {args.out}")


if __name__ == "__main__":
    main()
