"""Host-vs-device time split of the config-5 aggregate phase (mobility + UBAR + DMTT on FEMNIST CNN)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import load_config
from murmura_b200.parallel import engine as E
from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory

cfg = load_config("murmura_b200/examples/configs/mobility32_ubar_dmtt_liar_b200.yaml")
ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf))
acc = {}
def wrap(name):
    fn = getattr(E.B200Network, name)
    def inner(self, *a, **k):
        t0 = time.perf_counter(); r = fn(self, *a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
    setattr(E.B200Network, name, inner)
for n in ("_neighbors_for_round", "_edge_table", "_publish", "_dmtt_score_and_update", "_agg_ubar", "_local_training", "_evaluate", "_aggregate"):
    wrap(n)
net.train(rounds=3, lr=0.01)
acc.clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
net.train(rounds=5, lr=0.01)
torch.cuda.synchronize(); wall = time.perf_counter() - t0
print(json.dumps({"wall_ms_per_round": wall / 5 * 1e3, "host_ms_per_round": {k: round(v / 5 * 1e3, 2) for k, v in acc.items()},
                  "edges": int(sum(len(r) for r in net._received.tolist() if True) if hasattr(net, "_received") else -1),
                  "received_edges": int(net._received.sum())}))
