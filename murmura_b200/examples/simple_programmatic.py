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


def create_simple_model() -> nn.Module:
    """Same entry point name as the reference example (``simple_programmatic.py:15``): a fresh 10→50→2 MLP."""
    return TinyNet()


def create_synthetic_dataset(num_clients: int = 5, samples_per_client: int = 100, input_dim: int = 10, seed: int = 0):
    """One linearly-separable ``TensorDataset`` per client (reference ``simple_programmatic.py:24`` builds random data the same
    way); all clients share the separating direction so federation helps."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(input_dim, generator=g)
    out = []
    for _ in range(num_clients):
        x = torch.randn(samples_per_client, input_dim, generator=g)
        out.append(TensorDataset(x, (x @ w > 0).long()))
    return out


def make_network(num_nodes: int = 10, samples: int = 100, seed: int = 0) -> Network:
    nodes = []
    for nid, ds in enumerate(create_synthetic_dataset(num_nodes, samples, 10, seed)):
        loader = DataLoader(ds, batch_size=32, shuffle=True)
        nodes.append(Node(node_id=nid, model=create_simple_model(), train_loader=loader, test_loader=loader,
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
