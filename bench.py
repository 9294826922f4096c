#!/usr/bin/env python
"""Headline benchmark: FL rounds/sec on BASELINE.json config 2 —
8-node fully-connected FedAvg, ResNet-18 (11.19 M params), CIFAR-10-shaped synthetic non-IID shards
(Dirichlet α=0.5, 512 samples/node, batch 64, 1 local epoch, lr 0.01, evaluation every round).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference|nccl] [--config 2|3|4|5]

``--impl nccl`` is the comparison baseline BASELINE.json names: the same engine and training, but the neighbour exchange goes
through ``torch.distributed`` NCCL and the aggregation through stock PyTorch ops (``parallel/nccl_baseline.py``).

One "step" = one complete federated round (local SGD over every node's shard + neighbour exchange +
aggregation + evaluation of every node).  The 8 nodes are placed on N GPUs (8/N virtual nodes per GPU),
so total work is fixed → ``"scaling": "strong"``.  Every round is timed with its own CUDA-event pair on the
device, L2 is flushed (256 MiB write) between rounds outside the timed pairs, the per-rank sums are
max-reduced over ranks, and ``value`` = K / that time.  ``e2e`` repeats the measurement through the public
API (``Network.from_config(...).train(rounds=K)``) with the shards held in pinned host memory and copied
H2D every round plus the per-round D2H of the metrics, timed by wall clock.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fl_rounds_per_sec"

# BASELINE.json configs 2-5 (config 1 is the CPU simulation backend: scripts/cpu_config1.py).  The flagship (default) is 2.
CONFIGS = {
    2: {"yaml": "resnet18_fully_fedavg_b200.yaml", "model": "resnet18", "ref_model": ("baseline.ref_workloads.resnet18", {"num_classes": 10}),
        "ref_data": "cifar10"},
    3: {"yaml": "cnn_kregular_krum_gaussian_b200.yaml", "model": "cifar_cnn", "ref_model": ("baseline.ref_workloads.cifar_cnn", {"num_classes": 10}),
        "ref_data": "cifar10"},
    4: {"yaml": "er16_sketchguard_fp8_directed_b200.yaml", "model": "cifar_cnn", "ref_model": ("baseline.ref_workloads.cifar_cnn", {"num_classes": 10}),
        "ref_data": "cifar10"},
    5: {"yaml": "mobility32_ubar_dmtt_liar_b200.yaml", "model": "leaf_femnist_cnn", "ref_model": None, "ref_data": "femnist"},
}


def load_bench_config(idx: int):
    """The bundled YAML of BASELINE config ``idx`` (murmura_b200/examples/configs)."""
    from murmura_b200.config import load_config
    path = os.path.join(ROOT, "murmura_b200", "examples", "configs", CONFIGS[idx]["yaml"])
    return load_config(path)


def describe(cfg, idx: int, world: int):
    """The ``config`` block of the JSON line — identical for every arm (implementation details go to ``impl_detail``)."""
    d = cfg.data.params
    atk = cfg.attack
    return {"baseline_config": idx, "model": CONFIGS[idx]["model"], "nodes": cfg.topology.num_nodes,
            "topology": cfg.topology.type if cfg.mobility is None else "mobility G^t",
            "aggregation": cfg.aggregation.algorithm + ("+dmtt" if cfg.dmtt is not None else ""),
            "attack": f"{atk.type} {atk.percentage:g}" if atk.enabled else "none",
            "global_batch": cfg.training.batch_size * cfg.topology.num_nodes, "samples_per_node": d.get("samples_per_node"),
            "partition": f"{d.get('partition_method', 'dirichlet')} alpha={d.get('alpha', 0.5)}", "batch_size": cfg.training.batch_size,
            "local_epochs": cfg.training.local_epochs, "seq_len": None, "lr": cfg.training.lr, "eval_every": 1,
            "parallelism": f"{cfg.topology.num_nodes} federated nodes over {world} GPU(s) of one box",
            "l2": "flushed between rounds (256 MiB write), one CUDA-event pair per round"}


class ClockSampler:
    """SM clock + throttle reasons sampled in a background thread DURING the timed region (recipe's clocks line).

    NVML is polled directly (``pynvml``, ~10 ms period: the timed region of a short run is only a few hundred ms, less than
    one ``nvidia-smi`` start-up); ``nvidia-smi -lms`` is the fallback when the binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
            ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index: int, uuid: str = None):
        self.rows, self.proc, self.gpu, self.uuid = [], None, gpu_index, uuid
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid if self.uuid.startswith("GPU-") else "GPU-" + self.uuid)
            except Exception:
                pass
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu)

    def _poll(self, nv, h):
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(h))
                self.reasons.update(name for name, bit in self.BITS if mask & bit)
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            self._thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            self.source = "nvidia-smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) > 8 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        self.reasons.add(name)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self.proc is not None:
            self.proc.terminate()
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source,
                "power_w_max": round(max(self.power), 1) if self.power else None}


def _dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def _max_over_ranks(x: float, device) -> float:
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def _barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def _timed_rounds(run_round, steps: int, device, flush):
    """Per-round CUDA-event pairs; L2 flush between rounds is outside the pairs. Returns summed ms."""
    import torch
    pairs = []
    _barrier(); torch.cuda.synchronize(device)
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run_round(); b.record()
        pairs.append((a, b))
    torch.cuda.synchronize(device); _barrier()
    return sum(a.elapsed_time(b) for a, b in pairs)


def _b200_overrides(args):
    """``--b200 key=value …`` → extra ``b200:`` options (ablations: split_backward=false, fused_bn=false, gather_impl=ldg …)."""
    out = {}
    for kv in args.b200 or []:
        k, v = kv.split("=", 1)
        out[k] = {"true": True, "false": False}.get(v.lower(), int(v) if v.lstrip("-").isdigit() else v)
    return out


def _criterion(cfg):
    from murmura_b200.utils.factories import build_criterion
    return build_criterion(cfg)


def run_ours(args):
    """``--impl ours`` (fused kernels, in-kernel P2P / NVLS exchange) and ``--impl nccl`` (same engine, NCCL + PyTorch exchange/aggregation)."""
    import torch
    from murmura_b200 import Network
    from murmura_b200.parallel.engine import init_distributed
    from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
    from murmura_b200.utils.seed import set_seed

    rank, world, local_rank = init_distributed()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    transport = "nccl" if args.impl == "nccl" else args.transport

    def make(stream_inputs: bool, profile: bool = False):
        cfg = load_bench_config(args.config)
        cfg.backend = "b200"
        cfg.experiment.rounds = args.steps + args.warmup
        over = {"stream_inputs": stream_inputs, "transport": transport, "profile": profile, **_b200_overrides(args)}
        for k, v in over.items():
            setattr(cfg.b200, k, v)
        set_seed(cfg.experiment.seed)
        adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
        crit, evid = _criterion(cfg)
        import contextlib, io
        with (contextlib.redirect_stdout(io.StringIO()) if rank else contextlib.nullcontext()):
            net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf), device=device, criterion=crit, evidential=evid)
        return net, cfg

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flush = lambda: flush_buf.fill_(1)
    sampler = ClockSampler(local_rank, str(getattr(torch.cuda.get_device_properties(device), 'uuid', '') or ''))

    # ---- device-timed number (shards resident in HBM) -----------------------------------------
    net, cfg = make(stream_inputs=False)
    T = cfg.training
    net.train(rounds=args.warmup, local_epochs=T.local_epochs, lr=T.lr)
    launches0 = net.kernel_launches
    if rank == 0:
        sampler.start()
    ms = _timed_rounds(lambda: net.train(rounds=1, local_epochs=T.local_epochs, lr=T.lr), args.steps, device, flush)
    clocks = sampler.stop() if rank == 0 else None
    launches = net.kernel_launches - launches0
    ms = _max_over_ranks(ms, device)
    final_acc = float(net.history["mean_accuracy"][-1])
    params = net.layout.P_float_real
    fused = getattr(net, "fused", None)
    detail = {"transport": net.opt.transport, "fused_train": fused is not None, "nodes_per_gpu": net.V, "params_per_node": params,
              "tma_conv_launches_per_round": (fused.be.tma_launches if fused is not None else 0)}
    # ---- phase split + roofline of the exchange+aggregate path (extra profiled rounds, outside the timed region) ----------
    net.opt.profile = True
    net.reset_timers()
    net.train(rounds=5, local_epochs=T.local_epochs, lr=T.lr)
    t = dict(net.timers)
    r = max(t["rounds"], 1)
    agg_ms = _max_over_ranks(t["aggregate_ms"] / r, device)
    peaks = _peaks()
    hbm_gbs = t.get("hbm_bytes", 0.0) / r / max(agg_ms, 1e-6) / 1e6
    nvl_gbs = t.get("nvlink_bytes", 0.0) / r / max(agg_ms, 1e-6) / 1e6
    phases = {"train_ms": round(_max_over_ranks(t["train_ms"] / r, device), 3), "aggregate_ms": round(agg_ms, 3),
              "eval_ms": round(_max_over_ranks(t["eval_ms"] / r, device), 3)}
    exch = {"ms_per_round": round(agg_ms, 4), "hbm_gbs": round(hbm_gbs, 1), "nvlink_gbs": round(nvl_gbs, 1),
            "roofline_frac": round(max(hbm_gbs / peaks["hbm_gbs"], nvl_gbs / peaks["nvlink_gbs"]), 3),
            "of": f"measured copy bandwidth {peaks['hbm_gbs']:.0f} GB/s (HBM) / {peaks['nvlink_gbs']:.0f} GB/s (NVLink peer), algorithmic bytes of rank 0"}
    net.close()

    # ---- end-to-end number through the public API (pinned host shards → H2D every round, metrics D2H) ----
    net, cfg = make(stream_inputs=True)
    net.train(rounds=args.warmup, local_epochs=T.local_epochs, lr=T.lr)
    _barrier(); torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    net.train(rounds=args.steps, local_epochs=T.local_epochs, lr=T.lr)
    torch.cuda.synchronize(device); _barrier()
    wall = _max_over_ranks(time.perf_counter() - t0, device)
    h2d = _max_over_ranks(float(net.h2d_bytes_per_round), device)
    d2h = float(net.d2h_bytes_per_round)
    net.close()

    if rank == 0:
        value = args.steps / (ms / 1e3)
        out = {"metric": METRIC, "value": value, "unit": "rounds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "fp32 parameters, TF32 tensor-core math (tcgen05 kind::tf32; the reference's cuDNN default)", "data": "synthetic",
               "impl": args.impl, "final_acc": final_acc, "config": describe(cfg, args.config, world), "clocks": clocks,
               "e2e": {"value": args.steps / wall, "unit": "rounds/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "timing": "wall clock around Network.train(rounds=K), barrier+synchronize both sides, max over ranks"},
               "gpu_launches": int(launches), "phase_ms": phases, "exchange_aggregate": exch, "roofline_frac": exch["roofline_frac"],
               "impl_detail": detail}
        print(json.dumps(out), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return {"hbm_gbs": float(p["hbm_gbs"]), "nvlink_gbs": 770.0}
    except Exception:  # noqa: BLE001 - the profiling recipe's stated fallback
        return {"hbm_gbs": 6650.0, "nvlink_gbs": 770.0}


def run_reference(args):
    """UNMODIFIED reference (baseline/_ref) through its own public API and stock simulation code path."""
    rank, world, local_rank = _dist_env()
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    reason = None
    if not os.path.isdir(os.path.join(ref_root, "murmura")):
        reason = "baseline/_ref/murmura not installed (see DESIGN.md)"
    elif CONFIGS[args.config]["ref_model"] is None:
        reason = "config 5 needs the reference's distributed backend (wall-clock rounds of 120 s by construction)"
    if reason:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": reason}))
        return
    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    ours = load_bench_config(args.config)                       # only read for the workload description (no engine code runs)
    out = None
    if rank == 0:       # the reference has no multi-GPU path (every node uses get_device() → cuda:0); other ranks idle
        sys.path.insert(0, ref_root)
        import murmura
        from murmura import Network
        from murmura.config import Config
        from murmura.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
        from murmura.utils.seed import set_seed
        assert os.path.realpath(murmura.__file__).startswith(os.path.realpath(ref_root))
        dp = ours.data.params
        ref_model, ref_params = CONFIGS[args.config]["ref_model"]
        topo = {"type": ours.topology.type, "num_nodes": ours.topology.num_nodes, "seed": ours.topology.seed}
        if ours.topology.p is not None:
            topo["p"] = ours.topology.p
        if ours.topology.k is not None:
            topo["k"] = ours.topology.k
        cfg = Config(**{
            "experiment": {"name": "bench-reference", "rounds": args.steps + args.warmup, "seed": ours.experiment.seed},
            "topology": topo,
            "aggregation": {"algorithm": ours.aggregation.algorithm, "params": dict(ours.aggregation.params)},
            "attack": {"enabled": ours.attack.enabled, "type": ours.attack.type, "percentage": ours.attack.percentage, "params": dict(ours.attack.params)},
            "training": {"batch_size": ours.training.batch_size, "lr": ours.training.lr, "local_epochs": ours.training.local_epochs},
            "data": {"adapter": "baseline.ref_workloads.SyntheticRefAdapter",
                     "params": {"name": CONFIGS[args.config]["ref_data"], "num_nodes": ours.topology.num_nodes,
                                "samples_per_node": dp.get("samples_per_node", 512), "alpha": dp.get("alpha", 0.5), "seed": ours.experiment.seed}},
            "model": {"factory": ref_model, "params": ref_params},
        })
        T = cfg.training
        set_seed(cfg.experiment.seed)
        adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
        net = Network.from_config(config=cfg, model_factory=mf, dataset_adapter=adapter,
                                  aggregator_factory=build_aggregator_factory(cfg, mf, device), device=device)
        sample_bytes = int(torch.tensor(ours_sample_shape(CONFIGS[args.config]["ref_data"])).prod()) * 4 + 8
        shard_bytes = sum(len(p) for p in adapter.get_client_partitions()) * sample_bytes
        model_bytes = sum(v.numel() * v.element_size() for v in mf().state_dict().values())
        net.train(rounds=args.warmup, local_epochs=T.local_epochs, lr=T.lr)
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        sampler = ClockSampler(local_rank, str(getattr(torch.cuda.get_device_properties(device), 'uuid', '') or '')); sampler.start()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pairs = []
        for _ in range(args.steps):
            flush_buf.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net.train(rounds=1, local_epochs=T.local_epochs, lr=T.lr); b.record()
            pairs.append((a, b))
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        clocks = sampler.stop()
        N = ours.topology.num_nodes
        out = {"metric": METRIC, "value": args.steps / (ms / 1e3), "unit": "rounds/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "fp32 parameters, TF32 tensor-core math (cuDNN / cuBLAS torch defaults)", "data": "synthetic",
               "impl": "reference", "final_acc": float(net.history["mean_accuracy"][-1]), "config": describe(ours, args.config, world),
               "clocks": clocks,
               "e2e": {"value": args.steps / wall, "unit": "rounds/s", "h2d_bytes_per_step": int(shard_bytes * 2),
                       "d2h_bytes_per_step": int(2 * N * model_bytes),
                       "timing": "wall clock; the reference's DataLoader copies every batch H2D and get_state() copies every model D2H"},
               "gpu_launches": 0,
               "impl_detail": {"path": "reference simulation backend: sequential nodes on cuda:0 (it has no multi-GPU path); ranks > 0 idle",
                               "model": ref_model}}
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def ours_sample_shape(name: str):
    return {"cifar10": (3, 32, 32), "femnist": (1, 28, 28), "mnist": (784,)}[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference", "nccl"], default="ours")
    ap.add_argument("--config", type=int, choices=sorted(CONFIGS), default=2, help="BASELINE.json config (default 2 = the flagship)")
    ap.add_argument("--b200", nargs="*", default=[], help="extra b200 engine options as key=value (ablations)")
    ap.add_argument("--transport", choices=["auto", "p2p", "nvls"], default="auto",
                    help="auto: NVLS multimem reduce for full-mesh FedAvg over several GPUs, in-kernel peer loads otherwise")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
