"""Host-vs-device split of a bundled config: wall time per round, device time per phase, and a cProfile of the host side."""
import argparse, cProfile, io, json, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import load_config
from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory
from murmura_b200.utils.seed import set_seed

ap = argparse.ArgumentParser()
ap.add_argument("config"); ap.add_argument("--rounds", type=int, default=30); ap.add_argument("--top", type=int, default=25)
ap.add_argument("--eval-every", type=int, default=1)
args = ap.parse_args()
cfg = load_config(args.config); cfg.backend = "b200"; cfg.b200.profile = False
set_seed(cfg.experiment.seed)
ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg); crit, evid = build_criterion(cfg)
net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf), criterion=crit, evidential=evid)
T = cfg.training
net.train(rounds=5, local_epochs=T.local_epochs, lr=T.lr, eval_every=args.eval_every)
torch.cuda.synchronize()
t0 = time.perf_counter()
net.train(rounds=args.rounds, local_epochs=T.local_epochs, lr=T.lr, eval_every=args.eval_every)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.rounds
pr = cProfile.Profile(); pr.enable()
net.train(rounds=args.rounds, local_epochs=T.local_epochs, lr=T.lr, eval_every=args.eval_every)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(args.top)
print(json.dumps({"config": os.path.basename(args.config), "wall_ms_per_round": round(wall * 1e3, 3), "rounds_per_s": round(1 / wall, 1),
                  "final_acc": float(net.history["mean_accuracy"][-1])}))
print(s.getvalue()[:6000])
