"""Tables / figures from a results JSON written by ``run_suite.py`` (counterpart of the reference's
``experiments/paper/generate_figures.py``).  Always writes Markdown tables; PNG plots only when matplotlib is importable.

    python experiments/generate_figures.py experiments/results.json --out experiments/figures
"""
from __future__ import annotations

import argparse
import json
import os
from collections import defaultdict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("results"); ap.add_argument("--out", default="experiments/figures")
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
