import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

def mk(graphs, streams, n=4):
    cfg = Config(**{"experiment": {"name": "d", "rounds": 3, "seed": 3}, "topology": {"type": "fully", "num_nodes": n},
        "aggregation": {"algorithm": "fedavg"}, "training": {"batch_size": 32, "lr": 0.05},
        "data": {"adapter": "synthetic.mnist", "params": {"samples_per_node": 128, "partition_method": "iid"}},
        "model": {"factory": "models.mlp", "params": {"hidden_dims": [32]}}, "backend": "b200", "b200": {"cuda_graphs": graphs, "streams": streams}})
    ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    return Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf), device=torch.device("cuda"))

for graphs, streams in ((False, 4), (True, 1), (True, 4)):
    net = mk(graphs, streams)
    h = net.train(rounds=6, lr=0.05)
    print("graphs", graphs, "streams", streams, [round(float(a), 3) for a in h["mean_accuracy"]])
    net.close()
# per-node delta check: after one round of training only (no aggregation), how much did each node move?
for graphs, streams in ((False, 4), (True, 4)):
    net = mk(graphs, streams)
    net._prepare_training(1, 0.05)
    before = net.live.clone()
    net._local_training(1, 0.05); torch.cuda.synchronize()
    print("graphs", graphs, "per-node delta", [(net.live[i] - before[i]).abs().sum().item() for i in range(4)], "steps", [vn.step.item() for vn in net.nodes])
    net.close()
