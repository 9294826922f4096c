"""Resumable experiment-suite runner (counterpart of the reference's ``experiments/paper/run_comprehensive.py``).

Shells out to the CLI for every YAML under a directory, scrapes the stdout contract
(``Round k: Mean Accuracy = m ± s`` / ``Honest: …`` / ``Uncertainty: …``), records final accuracy, the convergence round
(first round with mean accuracy ≥ 80 %) and wall time, and skips configs that already have a result (resume).

    python experiments/run_suite.py experiments/configs --results experiments/results.json [--device cuda] [--timeout 1800]
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
import subprocess
import sys
import time

ROUND_RE = re.compile(r"Round (\d+): Mean Accuracy = ([\d.]+) ± ([\d.]+)")
HONEST_RE = re.compile(r"Honest: ([\d.]+), Compromised: ([\d.]+)")
UNC_RE = re.compile(r"Uncertainty: Vacuity=([\d.]+), Entropy=([\d.]+), Strength=([\d.]+)")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse(stdout: str) -> dict:
    rounds = []
    for line in stdout.splitlines():
        m = ROUND_RE.search(line)
        if m:
            rounds.append({"round": int(m.group(1)), "mean_accuracy": float(m.group(2)), "std_accuracy": float(m.group(3))})
            continue
        m = HONEST_RE.search(line)
        if m and rounds:
            rounds[-1].update(honest_accuracy=float(m.group(1)), compromised_accuracy=float(m.group(2)))
            continue
        m = UNC_RE.search(line)
        if m and rounds:
            rounds[-1].update(vacuity=float(m.group(1)), entropy=float(m.group(2)), strength=float(m.group(3)))
    out = {"rounds": rounds}
    if rounds:
        out["final_accuracy"] = rounds[-1]["mean_accuracy"]
        out["final_std"] = rounds[-1]["std_accuracy"]
        out["final_honest_accuracy"] = rounds[-1].get("honest_accuracy")
        out["final_compromised_accuracy"] = rounds[-1].get("compromised_accuracy")
        out["convergence_round"] = next((r["round"] for r in rounds if r["mean_accuracy"] >= 0.8), None)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config_dir"); ap.add_argument("--results", default="experiments/results.json")
    ap.add_argument("--device", default=None); ap.add_argument("--timeout", type=int, default=1800)
    ap.add_argument("--limit", type=int, default=0)
    ap.add_argument("--families", nargs="*", default=None,
                    help="run only these experiment families (keys of <config_dir>/index.json written by generate_configs.py)")
    args = ap.parse_args()
    results = json.load(open(args.results)) if os.path.exists(args.results) else {}
    todo = sorted(glob.glob(os.path.join(args.config_dir, "*.yaml")))
    if args.families:
        index = json.load(open(os.path.join(args.config_dir, "index.json")))
        unknown = [f for f in args.families if f not in index]
        if unknown:
            raise SystemExit(f"unknown families {unknown}; available: {sorted(index)}")
        keep = {name for f in args.families for name in index[f].values()}
        todo = [p for p in todo if os.path.basename(p) in keep]
    done = 0
    for path in todo:
        key = os.path.splitext(os.path.basename(path))[0]
        if key in results and results[key].get("status") == "ok":
            continue
        cmd = [sys.executable, "-m", "murmura_b200", "run", path] + (["--device", args.device] if args.device else [])
        t0 = time.time()
        try:
            proc = subprocess.run(cmd, capture_output=True, text=True, timeout=args.timeout, cwd=ROOT)
            rec = parse(proc.stdout)
            rec["status"] = "ok" if proc.returncode == 0 and rec["rounds"] else "failed"
            if rec["status"] == "failed":
                rec["stderr_tail"] = proc.stderr[-500:] + proc.stdout[-500:]
        except subprocess.TimeoutExpired:
            rec = {"status": "timeout", "rounds": []}
        rec["wall_s"] = round(time.time() - t0, 2)
        if rec.get("rounds"):
            rec["rounds_per_s"] = round(len(rec["rounds"]) / rec["wall_s"], 3)
        results[key] = rec
        json.dump(results, open(args.results, "w"), indent=1)
        print(f"{key:70s} {rec['status']:8s} acc={rec.get('final_accuracy')} conv={rec.get('convergence_round')} {rec['wall_s']}s", flush=True)
        done += 1
        if args.limit and done >= args.limit:
            break
    ok = [k for k, v in results.items() if v.get("status") == "ok"]
    print(f"{len(ok)}/{len(todo)} experiments have results → {args.results}")


if __name__ == "__main__":
    main()
