"""Run one bundled YAML on the b200 backend and print device-timed rounds/s + final accuracy as JSON.

    python scripts/run_config.py CONFIG.yaml [--rounds K] [--warmup W] [--transport p2p|nccl] [--set b200.key=value ...]
    (multi-GPU: python -m torch.distributed.run --nproc-per-node N scripts/run_config.py …)
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import load_config
from murmura_b200.parallel.engine import init_distributed
from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory
from murmura_b200.utils.seed import set_seed

ap = argparse.ArgumentParser()
ap.add_argument("config"); ap.add_argument("--rounds", type=int, default=10); ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--transport", default=None); ap.add_argument("--set", nargs="*", default=[])
args = ap.parse_args()
rank, world, local = init_distributed()
cfg = load_config(args.config)
cfg.backend = "b200"
if args.transport:
    cfg.b200.transport = args.transport
for kv in args.set:
    k, v = kv.split("=", 1)
    obj = cfg
    parts = k.split(".")
    for p in parts[:-1]:
        obj = getattr(obj, p)
    cur = getattr(obj, parts[-1])
    setattr(obj, parts[-1], type(cur)(v) if not isinstance(cur, bool) else v.lower() in ("1", "true", "yes"))
cfg.experiment.rounds = args.rounds + args.warmup
set_seed(cfg.experiment.seed)
adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg); crit, evid = build_criterion(cfg)
import contextlib, io
with (contextlib.redirect_stdout(io.StringIO()) if rank else contextlib.nullcontext()):
    net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf), criterion=crit, evidential=evid)
T = cfg.training
net.train(rounds=args.warmup, local_epochs=T.local_epochs, lr=T.lr)
if world > 1:
    torch.distributed.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
net.opt.profile = True
for k in ("train_ms", "aggregate_ms", "eval_ms", "rounds"):
    net.timers[k] = 0
a.record(); net.train(rounds=args.rounds, local_epochs=T.local_epochs, lr=T.lr); b.record(); torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], device="cuda")
if world > 1:
    torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
if rank == 0:
    h = net.history
    r = max(net.timers["rounds"], 1)
    print(json.dumps({"config": os.path.basename(args.config), "gpus": world, "transport": cfg.b200.transport, "rounds": args.rounds,
                      "rounds_per_s": round(args.rounds / (ms.item() / 1e3), 3), "ms_per_round": round(ms.item() / args.rounds, 3),
                      "train_ms": round(net.timers["train_ms"] / r, 3), "aggregate_ms": round(net.timers["aggregate_ms"] / r, 3),
                      "eval_ms": round(net.timers["eval_ms"] / r, 3), "final_acc": round(float(h["mean_accuracy"][-1]), 4),
                      "honest_acc": round(float(h["honest_accuracy"][-1]), 4) if h["honest_accuracy"] else None,
                      "nodes": cfg.topology.num_nodes, "aggregation": cfg.aggregation.algorithm, "params_per_node": net.layout.P_float_real}), flush=True)
net.close()
if world > 1:
    torch.distributed.destroy_process_group()
