"""The six figures of the paper's evaluation, drawn from a ``run_suite.py`` results file.

Counterpart of the reference's ``experiments/paper/generate_figures.py:109-513`` (fig1 non-IID robustness, fig2 IID→non-IID
degradation, fig3 personalisation at α = 0.1, fig4 rounds to convergence, fig5 hyper-parameter ablation, fig6 the 2×2 summary).
The figures are written as SVG by a small built-in chart writer, so the harness has no plotting dependency (matplotlib is not
part of the runtime image); the numbers behind every bar are also written to ``figure_data.json`` for the paper tables.

    python experiments/paper_figures.py experiments/results/simulation_all.json --index experiments/configs/index.json \
        --out experiments/figures
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
from typing import Dict, List, Optional, Sequence, Tuple

ALGORITHMS = ["fedavg", "balance", "sketchguard", "ubar", "evidential_trust"]
ALGO_NAMES = {"fedavg": "FedAvg", "krum": "Krum", "balance": "BALANCE", "sketchguard": "Sketchguard", "ubar": "UBAR",
              "evidential_trust": "Evidential Trust"}
COLORS = {"fedavg": "#4DBBD5", "krum": "#8491B4", "balance": "#00A087", "sketchguard": "#3C5488", "ubar": "#F39B7F",
          "evidential_trust": "#E64B35"}
ALPHAS = [("01", "α=0.1"), ("05", "α=0.5"), ("10", "α=1.0")]
ABLATION = [("accuracy_weight", "Accuracy weight (λ)"), ("self_weight", "Self weight (ω)"), ("trust_threshold", "Trust threshold (τ)"),
            ("vacuity_threshold", "Vacuity threshold (ν)")]


# ---------------------------------------------------------------------------------------------------------------------
# results access: family slot → record
# ---------------------------------------------------------------------------------------------------------------------
class Results:
    def __init__(self, results: Dict, index: Dict):
        self.res, self.index = results, index
        het = index.get("heterogeneity", {})
        self.datasets = sorted({slot.split("/")[0] for slot in het})

    def slot(self, family: str, slot: str) -> Optional[Dict]:
        fname = self.index.get(family, {}).get(slot)
        rec = self.res.get(os.path.splitext(fname)[0]) if fname else None
        return rec if rec and rec.get("status") == "ok" else None

    def het(self, dataset: str, algo: str, alpha: str) -> Optional[Dict]:
        return self.slot("heterogeneity", f"{dataset}/{algo}_alpha{alpha}")

    def avg_over_datasets(self, algo: str, alpha: str) -> Tuple[float, float, float]:
        """(mean accuracy %, std of that mean across datasets, mean per-node std %) — the three numbers fig1 / fig3 plot."""
        accs, stds = [], []
        for ds in self.datasets:
            r = self.het(ds, algo, alpha)
            if r is not None:
                accs.append(100.0 * r["final_accuracy"]); stds.append(100.0 * (r.get("final_std") or 0.0))
        if not accs:
            return 0.0, 0.0, 0.0
        return statistics.fmean(accs), statistics.pstdev(accs), statistics.fmean(stds)

    def degradation(self, algo: str) -> float:
        degs = []
        for ds in self.datasets:
            lo, hi = self.het(ds, algo, "01"), self.het(ds, algo, "10")
            if lo and hi:
                degs.append(100.0 * (hi["final_accuracy"] - lo["final_accuracy"]))
        return statistics.fmean(degs) if degs else 0.0

    def convergence(self, algo: str) -> float:
        rounds = [r["convergence_round"] for ds in self.datasets if (r := self.het(ds, algo, "01")) and r.get("convergence_round")]
        return statistics.fmean(rounds) if rounds else 0.0

    def ablation(self, param: str) -> List[float]:
        out = []
        for slot in self.index.get("ablation", {}):
            if param in slot:
                r = self.slot("ablation", slot)
                if r is not None:
                    out.append(100.0 * r["final_accuracy"])
        return out


# ---------------------------------------------------------------------------------------------------------------------
# a very small SVG bar-chart writer
# ---------------------------------------------------------------------------------------------------------------------
def _esc(s: str) -> str:
    return s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;")


def bar_panel(title: str, xlabel: str, ylabel: str, groups: Sequence[str], series: Sequence[Dict], *, ylim: Tuple[float, float],
              width: int = 560, height: int = 380, hline: Optional[float] = None, legend: bool = True, annotate: str = "",
              notes: Sequence[Tuple[int, str]] = ()) -> str:
    """One chart as an SVG ``<g>``.  ``series`` = [{name, color, values[len(groups)], errs?, bold?}]; with one series the bars take
    their colours from ``series[0]['colors']``.  ``annotate`` is a format string for the value labels ('' = none)."""
    L, R, T, B = 70, 16, 46, 64
    pw, ph = width - L - R, height - T - B
    y0, y1 = ylim
    sy = lambda v: T + ph * (1.0 - (min(max(v, y0), y1) - y0) / (y1 - y0))
    out = [f'<rect x="0" y="0" width="{width}" height="{height}" fill="white"/>',
           f'<text x="{width / 2}" y="20" text-anchor="middle" font-size="14" font-weight="bold">{_esc(title)}</text>']
    ticks = 5
    for t in range(ticks + 1):
        v = y0 + (y1 - y0) * t / ticks
        y = sy(v)
        out.append(f'<line x1="{L}" y1="{y:.1f}" x2="{L + pw}" y2="{y:.1f}" stroke="#dddddd" stroke-width="0.6"/>')
        out.append(f'<text x="{L - 6}" y="{y + 4:.1f}" text-anchor="end" font-size="11">{v:g}</text>')
    if hline is not None:
        out.append(f'<line x1="{L}" y1="{sy(hline):.1f}" x2="{L + pw}" y2="{sy(hline):.1f}" stroke="gray" stroke-dasharray="5,4" stroke-width="0.9"/>')
    ng, ns = len(groups), len(series)
    gw = pw / max(ng, 1)
    bw = gw * 0.78 / max(ns, 1)
    base = sy(max(y0, 0.0) if y0 <= 0.0 <= y1 else y0)
    for gi, gname in enumerate(groups):
        gx = L + gi * gw
        for si, s in enumerate(series):
            v = s["values"][gi]
            color = s["colors"][gi] if "colors" in s else s["color"]
            bold = (s.get("bolds") or [s.get("bold", False)] * ng)[gi]
            x = gx + gw * 0.11 + si * bw
            top, bot = min(sy(v), base), max(sy(v), base)
            out.append(f'<rect x="{x:.1f}" y="{top:.1f}" width="{bw * 0.94:.1f}" height="{max(bot - top, 0.5):.1f}" fill="{color}" '
                       f'stroke="{"black" if bold else "none"}" stroke-width="{1.6 if bold else 0}"/>')
            err = (s.get("errs") or [0.0] * ng)[gi]
            cx = x + bw * 0.47
            if err:
                out.append(f'<line x1="{cx:.1f}" y1="{sy(v - err):.1f}" x2="{cx:.1f}" y2="{sy(v + err):.1f}" stroke="black" stroke-width="1"/>')
                for e in (v - err, v + err):
                    out.append(f'<line x1="{cx - 3:.1f}" y1="{sy(e):.1f}" x2="{cx + 3:.1f}" y2="{sy(e):.1f}" stroke="black" stroke-width="1"/>')
            if annotate:
                out.append(f'<text x="{cx:.1f}" y="{sy(v + err) - 4 if v >= 0 else sy(v - err) + 12:.1f}" text-anchor="middle" font-size="10" '
                           f'font-weight="bold">{_esc(annotate.format(v))}</text>')
        for li, line in enumerate(gname.split("\n")):
            out.append(f'<text x="{gx + gw / 2:.1f}" y="{T + ph + 16 + 12 * li}" text-anchor="middle" font-size="11">{_esc(line)}</text>')
    for gi, text in notes:
        out.append(f'<text x="{L + gi * gw + gw / 2:.1f}" y="{T + 12}" text-anchor="middle" font-size="10" fill="#E64B35" font-weight="bold">{_esc(text)}</text>')
    out.append(f'<line x1="{L}" y1="{T}" x2="{L}" y2="{T + ph}" stroke="black"/><line x1="{L}" y1="{base:.1f}" x2="{L + pw}" y2="{base:.1f}" stroke="black"/>')
    out.append(f'<text x="{L + pw / 2}" y="{height - 8}" text-anchor="middle" font-size="12">{_esc(xlabel)}</text>')
    out.append(f'<text transform="translate(16,{T + ph / 2}) rotate(-90)" text-anchor="middle" font-size="12">{_esc(ylabel)}</text>')
    if legend and ns > 1:
        for si, s in enumerate(series):
            ly = T + ph - 14 * (ns - si) - 4
            out.append(f'<rect x="{L + pw - 130}" y="{ly}" width="10" height="10" fill="{s["color"]}"/>'
                       f'<text x="{L + pw - 116}" y="{ly + 9}" font-size="10">{_esc(s["name"])}</text>')
    return "\n".join(out)


def write_svg(path: str, panels: Sequence[Tuple[int, int, str]], width: int, height: int) -> None:
    body = "\n".join(f'<g transform="translate({x},{y})">{p}</g>' for x, y, p in panels)
    with open(path, "w") as fh:
        fh.write(f'<svg xmlns="http://www.w3.org/2000/svg" width="{width}" height="{height}" viewBox="0 0 {width} {height}" '
                 f'font-family="Helvetica, Arial, sans-serif">\n{body}\n</svg>\n')


# ---------------------------------------------------------------------------------------------------------------------
# the six figures
# ---------------------------------------------------------------------------------------------------------------------
def _single(values, errs=None):
    return [{"name": "", "color": "#888", "colors": [COLORS[a] for a in ALGORITHMS], "values": values, "errs": errs,
             "bolds": [a == "evidential_trust" for a in ALGORITHMS]}]


def panel_accuracy_vs_alpha(R: Results, title: str, with_err: bool) -> Tuple[str, Dict]:
    series, data = [], {}
    for a in ALGORITHMS:
        stats = [R.avg_over_datasets(a, al) for al, _ in ALPHAS]
        series.append({"name": ALGO_NAMES[a], "color": COLORS[a], "values": [s[0] for s in stats],
                       "errs": [s[1] for s in stats] if with_err else None, "bold": a == "evidential_trust"})
        data[a] = {al: {"mean_acc": round(s[0], 2), "std_across_datasets": round(s[1], 2)} for (al, _), s in zip(ALPHAS, stats)}
    svg = bar_panel(title, "Data heterogeneity level (Dirichlet α)", "Average accuracy (%)", [lbl for _, lbl in ALPHAS], series,
                    ylim=(0, 105), hline=90)
    return svg, data


def panel_degradation(R: Results, title: str) -> Tuple[str, Dict]:
    vals = [R.degradation(a) for a in ALGORITHMS]
    lo, hi = min(0.0, min(vals)) - 2, max(vals + [1.0]) * 1.25 + 2
    svg = bar_panel(title, "Algorithm", "Degradation (%)  IID α=1.0 → non-IID α=0.1", [ALGO_NAMES[a].replace(" ", "\n") for a in ALGORITHMS],
                    _single(vals), ylim=(lo, hi), annotate="{:.1f}%", legend=False)
    return svg, {a: round(v, 2) for a, v in zip(ALGORITHMS, vals)}


def panel_personalization(R: Results, title: str) -> Tuple[str, Dict]:
    stats = [R.avg_over_datasets(a, "01") for a in ALGORITHMS]
    svg = bar_panel(title, "Algorithm", "Accuracy (%)", [ALGO_NAMES[a].replace(" ", "\n") for a in ALGORITHMS],
                    _single([s[0] for s in stats], [s[2] for s in stats]), ylim=(0, 110), hline=90, annotate="{:.1f}%", legend=False)
    return svg, {a: {"mean_acc": round(s[0], 2), "mean_node_std": round(s[2], 2)} for a, s in zip(ALGORITHMS, stats)}


def panel_convergence(R: Results, title: str) -> Tuple[str, Dict]:
    vals = [R.convergence(a) for a in ALGORITHMS]
    notes = []
    live = [v for v in vals if v > 0]
    if live:
        notes.append((vals.index(min(live)), "fastest"))
        if vals[0] > 0 and vals[-1] > 0:
            notes.append((len(vals) - 1, f"{vals[0] / vals[-1]:.1f}x vs FedAvg") if vals.index(min(live)) != len(vals) - 1 else
                         (len(vals) - 1, f"fastest · {vals[0] / vals[-1]:.1f}x vs FedAvg"))
            if vals.index(min(live)) == len(vals) - 1:
                notes = notes[1:]
    svg = bar_panel(title, "Algorithm", "Rounds to convergence (≥ 80 %)", [ALGO_NAMES[a].replace(" ", "\n") for a in ALGORITHMS], _single(vals),
                    ylim=(0, max(vals + [1.0]) * 1.3), annotate="{:.1f}", legend=False, notes=notes)
    return svg, {a: round(v, 2) for a, v in zip(ALGORITHMS, vals)}


def panel_ablation(R: Results, title: str) -> Tuple[str, Dict]:
    labels, means, stds, data = [], [], [], {}
    for key, label in ABLATION:
        accs = R.ablation(key)
        if accs:
            labels.append(label.replace(" (", "\n(")); means.append(statistics.fmean(accs)); stds.append(statistics.pstdev(accs))
            data[key] = {"mean": round(means[-1], 2), "std": round(stds[-1], 2), "min": round(min(accs), 2), "max": round(max(accs), 2), "runs": len(accs)}
    if not labels:
        labels, means, stds = ["(no ablation runs)"], [0.0], [0.0]
    series = [{"name": "", "color": COLORS["evidential_trust"], "colors": [COLORS["evidential_trust"]] * len(labels), "values": means,
               "errs": stds, "bolds": [True] * len(labels)}]
    lo = max(0.0, min(means) - max(stds + [0.0]) - 10.0)
    svg = bar_panel(title, "Hyper-parameter", "Accuracy (%)", labels, series, ylim=(5 * int(lo // 5), 105), hline=90, annotate="{:.1f}%", legend=False)
    return svg, data


def generate(results: Dict, index: Dict, out_dir: str) -> Dict:
    os.makedirs(out_dir, exist_ok=True)
    R = Results(results, index)
    data: Dict = {"datasets": R.datasets}
    W, H = 560, 380
    nds = ", ".join(R.datasets)
    p, data["fig1_noniid_robustness"] = panel_accuracy_vs_alpha(R, f"Accuracy across heterogeneity levels (mean over {nds})", True)
    write_svg(os.path.join(out_dir, "fig1_noniid_robustness.svg"), [(0, 0, p)], W, H)
    p, data["fig2_degradation"] = panel_degradation(R, "Robustness to data heterogeneity (lower is better)")
    write_svg(os.path.join(out_dir, "fig2_degradation.svg"), [(0, 0, p)], W, H)
    p, data["fig3_personalization"] = panel_personalization(R, "Personalisation at α=0.1 (error bars: std over nodes)")
    write_svg(os.path.join(out_dir, "fig3_personalization.svg"), [(0, 0, p)], W, H)
    p, data["fig4_convergence"] = panel_convergence(R, "Convergence speed at α=0.1 (lower is better)")
    write_svg(os.path.join(out_dir, "fig4_convergence.svg"), [(0, 0, p)], W, H)
    p, data["fig5_ablation"] = panel_ablation(R, "Ablation: hyper-parameter sensitivity of Evidential Trust")
    write_svg(os.path.join(out_dir, "fig5_ablation.svg"), [(0, 0, p)], W, H)
    quad = [panel_accuracy_vs_alpha(R, "(a) Accuracy vs. heterogeneity", False)[0], panel_degradation(R, "(b) IID → non-IID degradation")[0],
            panel_personalization(R, "(c) Personalisation at α=0.1")[0], panel_convergence(R, "(d) Convergence speed")[0]]
    write_svg(os.path.join(out_dir, "fig6_combined_summary.svg"), [(0, 0, quad[0]), (W, 0, quad[1]), (0, H, quad[2]), (W, H, quad[3])], 2 * W, 2 * H)
    with open(os.path.join(out_dir, "figure_data.json"), "w") as fh:
        json.dump(data, fh, indent=1)
    return data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("results")
    ap.add_argument("--index", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs", "index.json"))
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "figures"))
    args = ap.parse_args()
    data = generate(json.load(open(args.results)), json.load(open(args.index)), args.out)
    print(f"wrote fig1..fig6 (*.svg) + figure_data.json to {args.out}; datasets: {data['datasets']}")


if __name__ == "__main__":
    main()
