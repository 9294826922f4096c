"""Tiny eager (no CUDA graph) run for ncu launch lists: 2 ResNet-18 nodes, 2 steps each, N rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

algo = sys.argv[1] if len(sys.argv) > 1 else "fedavg"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nodes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = sys.argv[4] if len(sys.argv) > 4 else "models.resnet18"
params = {"krum": {"num_compromised": 1}, "sketchguard": {"sketch_size": 1000}}.get(algo, {})
cfg = Config(**{"experiment": {"name": "prof", "rounds": rounds, "seed": 1}, "topology": {"type": "fully", "num_nodes": nodes},
                "aggregation": {"algorithm": algo, "params": params},
                "attack": {"enabled": algo != "fedavg", "type": "gaussian", "percentage": 0.2, "params": {"noise_std": 1.0}},
                "training": {"batch_size": 64, "lr": 0.01}, "data": {"adapter": "synthetic.cifar10", "params": {"samples_per_node": 128, "partition_method": "iid"}},
                "model": {"factory": model}, "backend": "b200",
                "b200": {"cuda_graphs": False, "streams": 1, "krum_gram": os.environ.get("KRUM_GRAM", "auto"), "sketch_dtype": os.environ.get("SKETCH", "fp32")}})
ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf), device=torch.device("cuda"))
torch.cuda.nvtx.range_push("rounds")
net.train(rounds=rounds, lr=0.01)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", net.history["mean_accuracy"])
