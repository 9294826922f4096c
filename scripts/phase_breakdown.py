import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.parallel.engine import init_distributed
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory
rank, world, local = init_distributed()
extra = dict(kv.split("=") for kv in sys.argv[1:])
b200 = {"profile": True, "streams": int(extra.get("streams", 8)), "channels_last": extra.get("cl", "1") == "1",
        "compute_dtype": extra.get("dtype", "fp32")}
cfg = Config(**{"experiment": {"name": "p", "rounds": 10, "seed": 42}, "topology": {"type": "fully", "num_nodes": 8}, "aggregation": {"algorithm": "fedavg"},
  "training": {"batch_size": 64, "lr": 0.01}, "data": {"adapter": "synthetic.cifar10", "params": {"samples_per_node": 512, "alpha": 0.5}},
  "model": {"factory": "models.resnet18"}, "backend": "b200", "b200": b200})
ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf))
net.train(rounds=3, lr=0.01)
for k in ("train_ms","aggregate_ms","eval_ms","rounds"): net.timers[k] = 0
net.train(rounds=10, lr=0.01)
if rank == 0:
    print(json.dumps({"world": world, **b200, "summary": net.perf_summary(), "acc": float(net.history["mean_accuracy"][-1]),
                      "nb": [vn.nb for vn in net.nodes]}))
net.close()
