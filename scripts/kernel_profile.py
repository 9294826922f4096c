"""Per-kernel device-time table of whole federated rounds (torch.profiler / CUPTI; numbers are for ranking only, never a bench value).

    python scripts/kernel_profile.py --config 5 --rounds 3 [--out profiles/kernels_config5.md]
"""
import argparse, collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="2", help="BASELINE config index (2..5) or a YAML path"); ap.add_argument("--rounds", type=int, default=3); ap.add_argument("--out", default="")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--phase", default="round", choices=["round", "train", "aggregate", "eval"])
args = ap.parse_args()
import bench
if str(args.config).isdigit():
    cfg = bench.load_bench_config(int(args.config))
else:
    from murmura_b200.config import load_config
    cfg = load_config(args.config); cfg.backend = "b200"
net = bench.build_network(cfg, torch.device("cuda:0")) if hasattr(bench, "build_network") else None
if net is None:
    from murmura_b200 import Network
    from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory
    ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg); crit, evid = build_criterion(cfg)
    net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf), device=torch.device("cuda:0"), criterion=crit, evidential=evid)
net.train(rounds=4, local_epochs=cfg.training.local_epochs, lr=cfg.training.lr)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    if args.phase == "round":
        net.train(rounds=args.rounds, local_epochs=cfg.training.local_epochs, lr=cfg.training.lr)
    else:
        for r in range(args.rounds):
            if args.phase == "train":
                net._local_training(cfg.training.local_epochs, cfg.training.lr)
            elif args.phase == "aggregate":
                net._aggregate(parity=r & 1)
            else:
                net._evaluate()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name
        tot[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot[name][1] += 1
rows = sorted(tot.items(), key=lambda kv: -kv[1][0])
total = sum(v[0] for _, v in rows)
lines = [f"# config {args.config} ({args.phase}): device time by kernel over {args.rounds} rounds (sum {total / args.rounds / 1e3:.2f} ms/round, profiler overhead included)",
         "", "| kernel | launches/round | µs/round | % | avg µs |", "|---|---|---|---|---|"]
for name, (us, n) in rows[: args.top]:
    lines.append(f"| `{name[:110]}` | {n / args.rounds:.1f} | {us / args.rounds:.1f} | {100 * us / total:.1f} | {us / n:.2f} |")
text = "\n".join(lines)
print(text)
if args.out:
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True); open(args.out, "w").write(text + "\n")
