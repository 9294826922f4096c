"""Device-timed roofline of the hand-written kernels at flagship sizes (ResNet-18 rows, P = 11.19 M floats).

CUDA events on the launching stream, >= 3 warm-ups, L2 flushed (256 MiB write) before every timed launch,
median of `reps`.  Fractions are reported against the MEASURED copy bandwidth / cuBLAS figures in
MEASURED_PEAKS.json ("of measured").  Writes profiles/kernel_roofline.json when run with --write.
"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ext = ops.ext()
dev = "cuda"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM = peaks["hbm_gbs"]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def edges(S, deg):
    row_ptr, slot = [0], []
    for v in range(S):
        slot += [v] + [(v + k + 1) % S for k in range(deg)]; row_ptr.append(len(slot))
    E = len(slot)
    t = lambda x, dt: torch.tensor(x, dtype=dt, device=dev)
    return {"row_ptr": t(row_ptr, torch.int32), "src_rank": torch.zeros(E, dtype=torch.int32, device=dev), "src_slot": t(slot, torch.int32),
            "mask": torch.ones(E, device=dev), "w": torch.full((E,), 1.0 / (deg + 1), device=dev), "E": E}


results = []
Pf = 11_191_242
Pf_pad = (Pf + 255) // 256 * 256
stride = Pf_pad + 256
only = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else None


def record(name, ms, bytes_moved, note=""):
    gbs = bytes_moved / ms / 1e6
    results.append({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(bytes_moved / 1e9, 4), "GBps": round(gbs, 1),
                    "frac_of_measured_hbm": round(gbs / HBM, 3), "note": note})
    print(f"{name:42s} {ms:8.4f} ms  {bytes_moved/1e9:7.3f} GB  {gbs:8.1f} GB/s  {gbs/HBM:5.2f} of measured HBM  {note}", flush=True)


for S, deg in ((8, 7), (20, 4), (32, 8)):
    live = torch.randn(S, stride, device=dev); pub = torch.randn(2, S, stride, device=dev)
    tbl = torch.tensor([pub.data_ptr()], dtype=torch.int64, device=dev)
    et = edges(S, deg)
    ints = torch.zeros(S, 20, dtype=torch.int64, device=dev)
    scale = torch.ones(S, device=dev); noise = torch.zeros(S, device=dev); gid = torch.arange(S, dtype=torch.int32, device=dev)
    ticket = torch.zeros(1, dtype=torch.int32, device=dev)
    if only in (None, "publish"):
        ms = timeit(lambda: ext.publish(live, pub[0].data_ptr(), stride, S, Pf, Pf_pad, ints, scale, noise, gid, 1, 0, 0, 1, 0, 1, ticket))
        record(f"publish S={S}", ms, 2 * S * Pf_pad * 4, "copy live->published (+epoch flag)")
        noise2 = torch.full((S,), 10.0, device=dev)
        ms = timeit(lambda: ext.publish(live, pub[0].data_ptr(), stride, S, Pf, Pf_pad, ints, scale, noise2, gid, 1, 0, 0, 1, 0, 1, ticket))
        record(f"publish+gaussian S={S}", ms, 2 * S * Pf_pad * 4, "Philox-4x32-10 + Box-Muller fused in the copy")
    if only in (None, "gather"):
        ms = timeit(lambda: ext.weighted_gather(live, tbl.data_ptr(), 0, stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                                                et["w"], Pf_pad, False, 0, 1, 0, 0.0, 0))
        record(f"weighted_gather S={S} deg={deg}", ms, (S * (deg + 1) + S) * Pf_pad * 4, "reads (deg+1) rows + writes 1 per node; neighbours may hit L2")
        ms = timeit(lambda: ext.weighted_gather(live, tbl.data_ptr(), 0, stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                                                et["w"], Pf_pad, False, 0, 1, 0, 0.0, 0, True))
        record(f"weighted_gather_tma S={S} deg={deg}", ms, (S * (deg + 1) + S) * Pf_pad * 4, "cp.async.bulk (UBLKCP) 8-stage ring through smem")
    if only in (None, "dist"):
        d2 = torch.zeros(et["E"], device=dev); n2 = torch.zeros(S, device=dev)
        ms = timeit(lambda: ext.edge_distances(live, tbl.data_ptr(), 0, stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                                               stride, d2, n2, 0, 1, 0, 0.0, 0))
        record(f"edge_distances S={S} deg={deg}", ms, S * (deg + 1) * stride * 4, "own + deg neighbour rows per node")
    if only in (None, "pairwise") and deg + 1 <= 32:
        D = torch.zeros(S, 32, 32, device=dev)
        ms = timeit(lambda: ext.pairwise_distances(live, tbl.data_ptr(), 0, stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                                                   Pf_pad, D, deg + 1, 0, 1, 0, 0.0, 0))
        record(f"pairwise_fp32 S={S} m={deg+1}", ms, S * (deg + 1) * Pf_pad * 4, "exact fp32 all-pairs per node (Krum fallback)")
    if only in (None, "gram"):
        # Gram over live+published planes of all S nodes, as the Krum plan does
        X = torch.randn(3 * S, stride, device=dev)
        gpr = (S + 7) // 8
        gy = [8 * j for j in range(gpr)] + [S + 8 * j for j in range(gpr)]
        if len(gy) <= 16:
            kbs = ext.gram_kb_per_stage(len(gy))
            maps = ext.gram_make_maps([X.data_ptr()], 3 * S, stride, Pf_pad, kbs)
            out = torch.zeros(128 * 128, device=dev)
            R = 8 * len(gy)
            ms = timeit(lambda: ext.gram_tf32(maps, [0] * len(gy), gy, 0, Pf_pad // 32, R, out, True, 0))
            record(f"gram_tcgen05_tf32 rows={R}", ms, R * Pf_pad * 4, f"TMA(3-D,{kbs} k-blocks/op)+tcgen05 split-K; {2*R*R*Pf_pad/ms/1e9:.1f} TFLOP/s tf32")
    if only in (None, "sketch"):
        from murmura_b200.aggregation.sketchguard import count_sketch_tables, pack_sketch_tables
        import numpy as np
        if S == 8:
            b, s = count_sketch_tables(Pf, 1000, 42)
            table = np.zeros((Pf + 3) // 4 * 4, dtype=np.uint16); table[:Pf] = pack_sketch_tables(b, s)
            tt = torch.from_numpy(table.view(np.int16)).to(dev)
            sk = torch.zeros(S, 1000, device=dev); slots = torch.arange(S, dtype=torch.int32, device=dev)
            ms = timeit(lambda: ext.count_sketch(live.data_ptr(), stride, slots, tt, Pf, 1000, sk))
            record(f"count_sketch S={S} K=1000", ms, S * Pf * 4 + Pf * 2 * S, "smem-privatised histogram; 4B value + 2B packed table per element")
    if only in (None, "sgd") and S == 8:
        params = [torch.randn(Pf // 8, device=dev) for _ in range(8)]; grads = [torch.randn_like(p) for p in params]
        ms = timeit(lambda: ext.sgd_multi(params, grads, 0.01))
        record("sgd_multi 8 tensors (one node)", ms, 3 * (Pf // 8) * 8 * 4, "read p,g write p")
    del live, pub

if "--write" in sys.argv:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"peaks": peaks, "results": results}, open(os.path.join(ROOT, "gpurun_out", "kernel_roofline.json"), "w"), indent=1)
