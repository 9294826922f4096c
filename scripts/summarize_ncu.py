"""Summarise gpurun_out/ncu/*.ncu-rep into profiles/ncu/<kernel>.md (+ raw csv extracts)."""
import csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ncu"); DST = os.path.join(ROOT, "profiles", "ncu")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]
os.makedirs(DST, exist_ok=True)
for rep in sorted(os.listdir(SRC)):
    if not rep.endswith(".ncu-rep"):
        continue
    name = rep[:-8]
    raw = subprocess.run(["ncu", "-i", os.path.join(SRC, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        print("skip", rep); continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    table = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = [f"# ncu --set full: `{table.get('Kernel Name', (name,))[0]}`", "",
             f"capture: `ncu --set full --clock-control none --import-source on -k regex:{name} -s 3 -c 1 python scripts/kernel_bench.py …` "
             f"on one B200 (report `gpurun_out/ncu/{rep}`, not committed: binary).", "", "| metric | value | unit |", "|---|---|---|"]
    for k in KEYS:
        if k in table:
            lines.append(f"| {k} | {table[k][0]} | {table[k][1]} |")
    extra = [h for h in hdr if re.search(r"tensor|tmem|utc|multimem", h, re.I) and h not in KEYS][:12]
    for k in extra:
        lines.append(f"| {k} | {table[k][0]} | {table[k][1]} |")
    # hottest source lines (needs -lineinfo)
    src = subprocess.run(["ncu", "-i", os.path.join(SRC, rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 2:
        h = srows[0]
        def col(name):
            for i, c in enumerate(h):
                if c.strip().lower() == name.lower():
                    return i
            return None
        ci, cs, cstall = col("Source"), col("# Samples") or col("Samples"), col("Warp Stall Sampling (All Samples)")
        key = cstall if cstall is not None else cs
        if ci is not None and key is not None:
            def num(x):
                try: return float(x.replace(",", ""))
                except Exception: return 0.0
            top = sorted(srows[1:], key=lambda r: -num(r[key]) if len(r) > key else 0)[:10]
            lines += ["", "Hottest SASS/source lines by warp-stall samples:", "", "```"]
            for r in top:
                lines.append(f"{num(r[key]):10.0f}  {r[ci][:150]}")
            lines.append("```")
    open(os.path.join(DST, name + ".md"), "w").write("\n".join(lines) + "\n")
    print("wrote", name)
