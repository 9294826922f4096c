#!/usr/bin/env python
"""Headline benchmark: FL rounds/sec on BASELINE.json config 2 —
8-node fully-connected FedAvg, ResNet-18 (11.19 M params), CIFAR-10-shaped synthetic non-IID shards
(Dirichlet α=0.5, 512 samples/node, batch 64, 1 local epoch, lr 0.01, evaluation every round).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]

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

WORKLOAD = {"nodes": 8, "samples_per_node": 512, "batch": 64, "local_epochs": 1, "lr": 0.01, "alpha": 0.5, "seed": 42}
METRIC = "fl_rounds_per_sec"


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


def run_ours(args):
    import torch
    from murmura_b200 import Network
    from murmura_b200.config import Config
    from murmura_b200.parallel.engine import init_distributed
    from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

    rank, world, local_rank = init_distributed()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    W = WORKLOAD

    def make(stream_inputs: bool):
        cfg = Config(**{
            "experiment": {"name": "bench-resnet18-fully8-fedavg", "rounds": args.steps + args.warmup, "seed": W["seed"]},
            "topology": {"type": "fully", "num_nodes": W["nodes"]},
            "aggregation": {"algorithm": "fedavg"},
            "training": {"batch_size": W["batch"], "lr": W["lr"], "local_epochs": W["local_epochs"]},
            "data": {"adapter": "synthetic.cifar10", "params": {"samples_per_node": W["samples_per_node"],
                                                                "partition_method": "dirichlet", "alpha": W["alpha"]}},
            "model": {"factory": "models.resnet18", "params": {"num_classes": 10}},
            "backend": "b200", "b200": {"stream_inputs": stream_inputs, "streams": 8, "transport": args.transport, **_b200_overrides(args)},
        })
        adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
        return Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf), device=device)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flush = lambda: flush_buf.fill_(1)
    sampler = ClockSampler(local_rank, str(getattr(torch.cuda.get_device_properties(device), 'uuid', '') or ''))

    # ---- device-timed number (shards resident in HBM) -----------------------------------------
    net = make(stream_inputs=False)
    net.train(rounds=args.warmup, local_epochs=W["local_epochs"], lr=W["lr"])
    net.reset_timers()
    launches0 = net.kernel_launches
    if rank == 0:
        sampler.start()
    ms = _timed_rounds(lambda: net.train(rounds=1, local_epochs=W["local_epochs"], lr=W["lr"]), args.steps, device, flush)
    clocks = sampler.stop() if rank == 0 else None
    launches = net.kernel_launches - launches0
    ms = _max_over_ranks(ms, device)
    final_acc = float(net.history["mean_accuracy"][-1])
    params = net.layout.P_float_real
    phases = None
    if net.opt.profile and net.timers["rounds"]:
        r = net.timers["rounds"]
        phases = {k: round(net.timers[k] / r, 3) for k in ("train_ms", "aggregate_ms", "eval_ms")}
        fused = getattr(net, "fused", None)
        phases["fused_train"] = fused is not None
        if fused is not None:
            phases["steps_per_round"] = fused.max_steps
            phases["active_nodes_per_step"] = fused.active
    net.close()

    # ---- end-to-end number through the public API (pinned host shards → H2D every round, metrics D2H) ----
    net = make(stream_inputs=True)
    net.train(rounds=args.warmup, local_epochs=W["local_epochs"], lr=W["lr"])
    _barrier(); torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    net.train(rounds=args.steps, local_epochs=W["local_epochs"], lr=W["lr"])
    torch.cuda.synchronize(device); _barrier()
    wall = _max_over_ranks(time.perf_counter() - t0, device)
    h2d = _max_over_ranks(float(net.h2d_bytes_per_round), device)
    d2h = float(net.metrics_host.numel() * 4)
    net.close()

    if rank == 0:
        value = args.steps / (ms / 1e3)
        out = {"metric": METRIC, "value": value, "unit": "rounds/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "fp32 (cuDNN TF32 conv math = reference default)", "data": "synthetic", "impl": "ours",
               "final_acc": final_acc,
               "config": {"model": "resnet18 (11,191,242 float state elems)", "nodes": W["nodes"], "topology": "fully-connected",
                          "aggregation": "fedavg", "global_batch": W["batch"] * W["nodes"], "samples_per_node": W["samples_per_node"],
                          "batch_size": W["batch"], "local_epochs": W["local_epochs"], "seq_len": None, "lr": W["lr"],
                          "parallelism": f"{W['nodes']} federated nodes on {world} GPU(s) ({W['nodes'] // world}/GPU), fused exchange+aggregate ({args.transport})",
                          "l2": "flushed between rounds (256 MiB write), one CUDA-event pair per round",
                          "params_per_node": params},
               "clocks": clocks,
               "e2e": {"value": args.steps / wall, "unit": "rounds/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "timing": "wall clock around Network.train(rounds=K), barrier+synchronize both sides, max over ranks"},
               "gpu_launches": int(launches)}
        if phases is not None:
            out["phase_ms"] = phases
        print(json.dumps(out), flush=True)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def run_reference(args):
    """UNMODIFIED reference (baseline/_ref) through its own public API and stock simulation code path."""
    rank, world, local_rank = _dist_env()
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_root, "murmura")):
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/murmura not installed (see DESIGN.md)"}))
        return
    import torch
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    out = None
    if rank == 0:       # the reference has no multi-GPU path (every node uses get_device() → cuda:0); other ranks idle
        sys.path.insert(0, ref_root)
        import murmura
        from murmura import Network
        from murmura.config import Config
        from murmura.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
        from murmura.utils.seed import set_seed
        assert os.path.realpath(murmura.__file__).startswith(os.path.realpath(ref_root))
        W = WORKLOAD
        cfg = Config(**{
            "experiment": {"name": "bench-reference", "rounds": args.steps + args.warmup, "seed": W["seed"]},
            "topology": {"type": "fully", "num_nodes": W["nodes"]},
            "aggregation": {"algorithm": "fedavg"},
            "training": {"batch_size": W["batch"], "lr": W["lr"], "local_epochs": W["local_epochs"]},
            "data": {"adapter": "baseline.ref_workloads.SyntheticRefAdapter",
                     "params": {"name": "cifar10", "num_nodes": W["nodes"], "samples_per_node": W["samples_per_node"],
                                "alpha": W["alpha"], "seed": W["seed"]}},
            "model": {"factory": "baseline.ref_workloads.resnet18", "params": {"num_classes": 10}},
        })
        set_seed(W["seed"])
        adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
        net = Network.from_config(config=cfg, model_factory=mf, dataset_adapter=adapter,
                                  aggregator_factory=build_aggregator_factory(cfg, mf, device), device=device)
        shard_bytes = sum(len(p) for p in adapter.get_client_partitions()) * (3 * 32 * 32 * 4 + 8)
        net.train(rounds=args.warmup, local_epochs=W["local_epochs"], lr=W["lr"])
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=device)
        sampler = ClockSampler(local_rank, str(getattr(torch.cuda.get_device_properties(device), 'uuid', '') or '')); sampler.start()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pairs = []
        for _ in range(args.steps):
            flush_buf.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); net.train(rounds=1, local_epochs=W["local_epochs"], lr=W["lr"]); b.record()
            pairs.append((a, b))
        torch.cuda.synchronize(device)
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in pairs)
        clocks = sampler.stop()
        out = {"metric": METRIC, "value": args.steps / (ms / 1e3), "unit": "rounds/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "fp32 (cuDNN TF32 conv math, torch defaults)", "data": "synthetic", "impl": "reference",
               "final_acc": float(net.history["mean_accuracy"][-1]),
               "config": {"model": "torchvision resnet18(num_classes=10)", "nodes": W["nodes"], "topology": "fully-connected",
                          "aggregation": "fedavg", "global_batch": W["batch"] * W["nodes"], "samples_per_node": W["samples_per_node"],
                          "batch_size": W["batch"], "local_epochs": W["local_epochs"], "seq_len": None, "lr": W["lr"],
                          "parallelism": "reference simulation backend: sequential nodes on cuda:0 (it has no multi-GPU path); "
                                         "ranks > 0 idle", "l2": "flushed between rounds (256 MiB write), one CUDA-event pair per round"},
               "clocks": clocks,
               "e2e": {"value": args.steps / wall, "unit": "rounds/s", "h2d_bytes_per_step": int(shard_bytes * 2),
                       "d2h_bytes_per_step": int(2 * 8 * 11_200_000 * 4),
                       "timing": "wall clock; the reference's DataLoader copies every batch H2D and get_state() copies every model D2H"},
               "gpu_launches": 0}
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--b200", nargs="*", default=[], help="extra b200 engine options as key=value (ablations)")
    ap.add_argument("--transport", choices=["p2p", "nvls", "nccl"], default="p2p",
                    help="p2p: in-kernel peer loads (default); nvls: full-mesh FedAvg through multimem.ld_reduce; nccl: baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
