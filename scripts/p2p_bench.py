"""NVLink roofline of the fused exchange+aggregate kernels (run under torchrun on >= 2 GPUs).

Times `weighted_gather` (in-kernel P2P loads), the NVLS `nvls_fedavg` kernel and `edge_distances` on ResNet-18-sized rows where
the neighbours live on PEER GPUs, with CUDA events on every rank (max over ranks), and reports the bytes that must cross
NVLink per GPU ÷ time against the measured peer-copy bandwidth (770 GB/s per direction per GPU, B200_PROFILING.md).
"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.parallel.engine import init_distributed
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

NVLINK_GBS = 770.0
rank, world, _ = init_distributed()
n = 8 if world <= 8 else world


def build(transport, topo, gather="ldg", rank_sum=True, two_shot=True):
    cfg = Config(**{"experiment": {"name": "p2p", "rounds": 4, "seed": 1}, "topology": topo, "aggregation": {"algorithm": "fedavg"},
                    "training": {"batch_size": 16, "lr": 0.01}, "data": {"adapter": "synthetic.cifar10", "params": {"samples_per_node": 16, "partition_method": "iid"}},
                    "model": {"factory": "models.resnet18"}, "backend": "b200", "b200": {"transport": transport, "placement": "contiguous", "cuda_graphs": False, "gather_impl": gather, "fullmesh_rank_sum": rank_sum, "fullmesh_two_shot": two_shot}})
    ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    return Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf))


def timed(net, reps=12):
    ts = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for r in range(reps + 3):
        net.round_idx = r
        flush.fill_(1)
        # publish first (not timed), then make sure every rank's publish is visible so the timed region is the pure exchange+aggregate
        parity = r & 1
        neighbors, key = net._neighbors_for_round(r)
        et = net._edge_table(neighbors, key)
        et["rank_sum"] = net._fullmesh_fedavg(et)
        net._liveness_frozen = False
        net._publish(parity, with_sum=et["rank_sum"])
        net._freeze_liveness(); torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net._agg_fedavg(et, parity); b.record(); torch.cuda.synchronize()
        if r >= 3:
            ts.append(a.elapsed_time(b))
    ms = torch.tensor([statistics.median(ts)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


out = []
for transport, topo_name, gather, rank_sum, two_shot in (("p2p", "fully", "ldg", True, True), ("nvls", "fully", "ldg", True, True),
                                                        ("p2p", "fully", "ldg", True, False), ("nvls", "fully", "ldg", True, False),
                                                        ("p2p", "fully", "ldg", False, False), ("p2p", "fully", "tma", False, False),
                                                        ("nvls", "fully", "ldg", False, False), ("p2p", "ring", "ldg", False, False),
                                                        ("p2p", "ring", "tma", False, False)):
    topo = {"type": topo_name, "num_nodes": n}
    net = build(transport, topo, gather, rank_sum, two_shot)
    ms = timed(net)
    L, pl = net.layout, net.placement
    remote = sum(1 for vn in net.nodes for j in net.topology.neighbors[vn.gid] if pl.rank_of[j] != net.rank)
    local = sum(1 for vn in net.nodes for j in net.topology.neighbors[vn.gid] if pl.rank_of[j] == net.rank) + len(net.nodes)
    row = L.Pf_pad * 4
    if rank_sum and two_shot and topo_name == "fully":
        if transport == "nvls":
            link_bytes = row * (1.0 / world + (world - 1.0) / world)
            note = "two-shot on rank sums: multimem.ld_reduce of the own slice + multimem.st of it to every GPU"
        else:
            link_bytes = row * 2.0 * (world - 1.0) / world
            note = "two-shot on rank sums: own slice from every peer + reduced slices scattered by peer stores"
    elif rank_sum and topo_name == "fully":
        if transport == "nvls":
            link_bytes = row * (world - 1) / world
            note = "per-rank sum rows, multimem.ld_reduce: ONE reduced row per GPU"
        else:
            link_bytes = row * (world - 1)
            note = f"per-rank sum rows: {world - 1} peer rows read per GPU"
    elif transport == "nvls" and topo_name == "fully":
        link_bytes = net.S * row * (world - 1) / world            # in-switch reduction: each GPU ingests S reduced rows (own share stays local)
        note = "multimem.ld_reduce: the switch sums the ranks' copies"
    else:
        link_bytes = remote * row
        note = f"{remote} remote + {local} local row reads per GPU"
    rec = {"kernel": f"fedavg exchange+aggregate ({transport}, {topo_name}, gather={gather}, rank_sum={rank_sum and topo_name == 'fully'}, two_shot={rank_sum and two_shot and topo_name == 'fully'})", "gpus": world, "nodes": n, "ms": round(ms, 4),
           "nvlink_GB_per_gpu": round(link_bytes / 1e9, 4), "nvlink_GBps": round(link_bytes / ms / 1e6, 1),
           "frac_of_measured_nvlink": round(link_bytes / ms / 1e6 / NVLINK_GBS, 3), "note": note}
    out.append(rec)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    net.close()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/p2p_roofline_{world}gpu.json", "w"), indent=1)
dist.barrier(); dist.destroy_process_group()
