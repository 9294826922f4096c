"""Epoch-flag race stress (SURVEY §4 test plan, item 3): many rounds of publish → in-kernel flag wait → P2P gather with
RANDOMISED per-rank delays injected before every publish, verifying every round against a closed-form oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/flag_stress.py [rounds]

Each node's row is a constant (value = f(node, round)), so the FedAvg result over any neighbourhood is known exactly; a reader
that runs ahead of a slow publisher (stale parity buffer) or a publisher that overwrites a buffer still being read would
produce a wrong mean immediately.
"""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.parallel.engine import init_distributed
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rank, world, _ = init_distributed()
n = 3 * world
cfg = Config(**{"experiment": {"name": "stress", "rounds": rounds, "seed": 1}, "topology": {"type": "k-regular", "num_nodes": n, "k": 4},
                "aggregation": {"algorithm": "fedavg"}, "training": {"batch_size": 8, "lr": 0.0},
                "data": {"adapter": "synthetic.mnist", "params": {"samples_per_node": 8, "partition_method": "iid"}},
                "model": {"factory": "models.mlp", "params": {"hidden_dims": [64]}}, "backend": "b200", "b200": {"placement": "contiguous"}})
ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
net = Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf))
L = net.layout
rng = random.Random(1234 + rank)
val = lambda g, r: float((g * 7 + r * 13) % 101) / 10.0
bad = 0
for r in range(rounds):
    net.round_idx = r
    for vn in net.nodes:
        net.live[vn.slot, : L.Pf].fill_(val(vn.gid, r))
    if rng.random() < 0.5:                                   # random device-side delay: this rank publishes late
        torch.cuda._sleep(int(rng.random() * 3e6))
    net._aggregate(parity=r & 1)
    if r % 50 == 49 or r == rounds - 1:                      # check in batches (the check itself synchronises)
        torch.cuda.synchronize()
        for vn in net.nodes:
            ids = [vn.gid] + net.topology.neighbors[vn.gid]
            want = sum(val(g, r) for g in ids) / len(ids)
            got = net.live[vn.slot, : L.Pp]
            if not torch.allclose(got, torch.full_like(got, want), atol=1e-4):
                bad += 1
        if int(net.arena.timed_out.item()) != 0:
            bad += 1000
t = torch.tensor([bad], device="cuda"); dist.all_reduce(t)
if rank == 0:
    print(f"flag stress: {rounds} rounds on {world} GPUs, {n} nodes, mismatches={int(t.item())}", flush=True)
net.close()
dist.destroy_process_group()
sys.exit(1 if int(t.item()) else 0)
