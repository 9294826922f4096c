"""Zero-dependency programmatic example: 10-node ring, tiny MLP, random data, FedAvg.

Mirrors what reference ``murmura/examples/simple_programmatic.py:15-100`` demonstrates (building
``Node``s by hand and calling ``Network.train``) and doubles as the API smoke test.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.utils.data import DataLoader, TensorDataset

from murmura_b200 import FedAvgAggregator, Network, Node, create_topology


class TinyNet(nn.Module):
    def __init__(self, in_dim: int = 10, hidden: int = 50, classes: int = 2):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, classes))

    def forward(self, x):
        return self.body(x)


def make_network(num_nodes: int = 10, samples: int = 100, seed: int = 0) -> Network:
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(10, generator=g)
    nodes = []
    for nid in range(num_nodes):
        x = torch.randn(samples, 10, generator=g)
        y = (x @ w > 0).long()
        loader = DataLoader(TensorDataset(x, y), batch_size=32, shuffle=True)
        nodes.append(Node(node_id=nid, model=TinyNet(), train_loader=loader, test_loader=loader,
                          aggregator=FedAvgAggregator(), device=torch.device("cpu")))
    return Network(nodes=nodes, topology=create_topology("ring", num_nodes=num_nodes))


def main(rounds: int = 10) -> dict:
    net = make_network()
    print(f"ring of {net.topology.num_nodes} nodes, avg degree {net.topology.avg_degree():.1f}")
    history = net.train(rounds=rounds, local_epochs=2, lr=0.05, verbose=True)
    print(f"final mean accuracy: {history['mean_accuracy'][-1]:.4f}")
    return history


if __name__ == "__main__":
    main()
