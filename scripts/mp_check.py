"""Multi-GPU correctness check (run under torchrun): fused P2P exchange+aggregate vs a CPU oracle on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/mp_check.py

Every node's state is a deterministic function of its global id, so each rank can rebuild *all* states on the
CPU, run the reference-parity aggregator classes for the nodes it hosts and compare with what the kernels wrote
after reading the neighbours' rows over NVLink.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.parallel.engine import init_distributed
from murmura_b200.utils.factories import build_aggregator_factory, build_criterion, build_dataset_adapter, build_model_factory

HAR = {"factory": "examples.wearables.uci_har", "params": {"input_dim": 561, "num_classes": 6, "hidden_dims": [32, 16]}}
HAR_DATA = {"adapter": "wearables.uci_har", "params": {"data_path": "synthetic", "samples_per_node": 24, "partition_method": "iid"}}
ATTACK = {"enabled": True, "type": "directed_deviation", "percentage": 0.3, "params": {"lambda_param": -5.0}}


def det_state(layout, gid, rnd=0):
    g = torch.Generator().manual_seed(1234 + gid + 1000 * rnd)
    row = torch.zeros(layout.stride)
    for e in layout.float_entries():
        base = torch.linspace(-1, 1, e.numel) * 0.1
        row[e.offset:e.offset + e.numel] = base + 0.02 * (gid + 1) * torch.randn(e.numel, generator=g)
        if e.name.endswith("running_var"):
            row[e.offset:e.offset + e.numel] = row[e.offset:e.offset + e.numel].abs() + 0.5
    ints = torch.arange(max(layout.Pi, 1)) + 3 * gid
    return row, ints


def state_dict_of(layout, row, ints):
    st = {k: v.clone() for k, v in layout.row_views(row, ints if layout.Pi else None).items()}
    return st


def check(algo, params, n, topo, b200=None, attack=ATTACK, rounds=2, tol=5e-5):
    rank = dist.get_rank()
    cfg = Config(**{"experiment": {"name": "mp", "rounds": 4, "seed": 5}, "topology": topo,
                    "aggregation": {"algorithm": algo, "params": params}, "attack": attack or {},
                    "training": {"batch_size": 32, "lr": 0.05}, "data": HAR_DATA, "model": HAR, "backend": "b200", "b200": b200 or {}})
    adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    crit, evid = build_criterion(cfg)
    net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf), device=None, criterion=crit, evidential=evid)
    L = net.layout
    agg_factory = build_aggregator_factory(cfg, mf, torch.device("cpu"))
    cpu_aggs = {vn.gid: agg_factory(vn.gid) for vn in net.nodes}
    worst = 0.0
    for r in range(rounds):
        net.round_idx = r
        rows = {g: det_state(L, g, r) for g in range(n)}
        for vn in net.nodes:
            net.live[vn.slot].copy_(rows[vn.gid][0]); net.ints[vn.slot].copy_(rows[vn.gid][1][: net.ints.shape[1]])
        net._aggregate(parity=r & 1)
        torch.cuda.synchronize()
        if int(net.arena.timed_out.item()) != 0:
            raise RuntimeError(f"rank {rank}: peers timed out: mask={net.arena.timed_out.item()}")
        own = {g: state_dict_of(L, rows[g][0], rows[g][1]) for g in range(n)}
        pub = {}
        for g in range(n):
            st = {k: v.clone() for k, v in own[g].items()}
            if g in net.compromised:
                st = {k: (v * -5.0 if v.is_floating_point() else v) for k, v in st.items()}
            pub[g] = st
        for vn in net.nodes:
            template = mf()
            out = cpu_aggs[vn.gid].aggregate(node_id=vn.gid, own_state=own[vn.gid],
                                             neighbor_states={j: pub[j] for j in net.topology.neighbors[vn.gid]}, round_num=r,
                                             train_loader=[(vn.X.cpu(), vn.y.cpu())], model_template=template, device=torch.device("cpu"))
            template.load_state_dict(out)
            want = template.state_dict()
            got = L.row_views(net.live[vn.slot], net.ints[vn.slot])
            for k, w in want.items():
                gk = got[k].detach().cpu()
                if w.is_floating_point():
                    err = (gk.float() - w.float()).abs().max().item()
                    worst = max(worst, err)
                    if err > tol * max(1.0, w.abs().max().item()):
                        et = net._last_et; vi = vn.slot; e0, e1 = et["host_rows"][vi], et["host_rows"][vi + 1]
                        dbg = {"gids": et["host_gid"][e0:e1], "w": et["w"][e0:e1].tolist(), "dist": et["dist"][e0:e1].tolist(),
                               "stats": et["stats"][vi].tolist(), "cpu_thr": getattr(cpu_aggs[vn.gid], "threshold_history", [None])[-1:],
                               "cpu_scores": {j: v[-1] for j, v in getattr(cpu_aggs[vn.gid], "neighbor_scores", {}).items()},
                               "byz": sorted(net.compromised),
                               "gpu_cand": et["aux"][e0:e1].tolist() if "aux" in et else None, "gpu_loss": et["aux3"][e0:e1].tolist() if "aux3" in et else None,
                               "gpu_own_loss": et["n2"][vi].item() if "n2" in et else None, "gpu_d2": et["d2"][e0:e1].tolist() if "d2" in et else None,
                               "cpu_losses": {j: v[-1] for j, v in getattr(cpu_aggs[vn.gid], "neighbor_losses", {}).items() if v},
                               "cpu_dist": {j: v[-1] for j, v in getattr(cpu_aggs[vn.gid], "neighbor_distances", {}).items() if v}}
                        raise AssertionError(f"[{algo}] rank {rank} node {vn.gid} key {k} round {r}: max err {err}\n  debug: {dbg}")
                elif not torch.equal(gk.long(), w.long()):
                    raise AssertionError(f"[{algo}] rank {rank} node {vn.gid} int key {k}: {gk} vs {w}")
    net.close()
    t = torch.tensor([worst], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"OK {algo:17s} n={n:2d} topo={topo['type']:9s} b200={b200 or {}} max_err={t.item():.2e}", flush=True)


def train_check():
    """Short real training across ranks: accuracy must improve and histories agree on every rank."""
    nodes = max(6, dist.get_world_size())            # every rank hosts at least one node
    cfg = Config(**{"experiment": {"name": "mp-train", "rounds": 4, "seed": 5}, "topology": {"type": "ring", "num_nodes": nodes},
                    "aggregation": {"algorithm": "fedavg"}, "training": {"batch_size": 32, "lr": 0.05},
                    "data": {"adapter": "synthetic.mnist", "params": {"samples_per_node": 128, "partition_method": "iid"}},
                    "model": {"factory": "models.mlp", "params": {"hidden_dims": [32]}}, "backend": "b200"})
    adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf))
    hist = net.train(rounds=6, lr=0.05)
    acc = torch.tensor(hist["mean_accuracy"], device="cuda", dtype=torch.float64)
    ref = acc.clone(); dist.broadcast(ref, 0)
    assert torch.equal(acc, ref), "ranks disagree on history"
    assert acc[-1] > acc[0] + 0.1, acc
    if dist.get_rank() == 0:
        print(f"OK training ring{nodes} fedavg acc", [round(float(a), 3) for a in acc], flush=True)
    net.close()


def main():
    init_distributed()
    G = dist.get_world_size()
    n = int(os.environ.get("MP_NODES", max(6, 3 * G)))
    only = os.environ.get("MP_ONLY")
    if only == "nvls":
        check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"transport": "nvls"})
        check("fedavg", {}, n, {"type": "ring", "num_nodes": n}, b200={"transport": "nvls"})
        check("krum", {"num_compromised": 1}, n, {"type": "k-regular", "num_nodes": n, "k": 4}, b200={"transport": "nvls", "krum_gram": "tcgen05"})
        dist.barrier(); dist.destroy_process_group(); return
    if only == "ubar":
        check("ubar", {"rho": 0.6, "alpha": 0.5}, n, {"type": "k-regular", "num_nodes": n, "k": 4}, b200={"grouped_mlp": False})
        check("ubar", {"rho": 0.6, "alpha": 0.5}, n, {"type": "k-regular", "num_nodes": n, "k": 4}, b200={"grouped_mlp": False, "fused_train": False})
        dist.barrier(); dist.destroy_process_group(); return
    if only == "sketch":
        kreg = {"type": "k-regular", "num_nodes": n, "k": 4}
        check("sketchguard", {"sketch_size": 256, "gamma": 0.6, "alpha": 0.5}, n, kreg)
        check("sketchguard", {"sketch_size": 256, "gamma": 0.6, "alpha": 0.5}, n, kreg, b200={"sketch_dtype": "fp8"})
        dist.barrier(); dist.destroy_process_group(); return
    kreg = {"type": "k-regular", "num_nodes": n, "k": 4}
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n})
    check("fedavg", {}, n, {"type": "ring", "num_nodes": n})
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"transport": "nvls"})       # multimem.ld_reduce in the switch
    check("fedavg", {}, n, {"type": "ring", "num_nodes": n}, b200={"transport": "nvls"})        # symm-memory arena, P2P gather
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"fullmesh_rank_sum": False})                         # edge-list gather
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"transport": "nvls", "fullmesh_rank_sum": False})    # per-slot multimem
    check("fedavg", {}, n + 1, {"type": "fully", "num_nodes": n + 1}, b200={"transport": "nvls"})                        # uneven slots per rank
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"fullmesh_two_shot": True})                          # fused reduce-scatter + all-gather
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"transport": "nvls", "fullmesh_two_shot": True})     # … through the switch
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"fullmesh_two_shot": False})                         # one-shot rank sums (p2p)
    check("fedavg", {}, n, {"type": "fully", "num_nodes": n}, b200={"transport": "nvls", "fullmesh_two_shot": False})    # one-shot, multimem
    check("balance", {"gamma": 0.6, "alpha": 0.5}, n, kreg)
    check("krum", {"num_compromised": 1}, n, kreg, b200={"krum_gram": "fp32"})
    check("krum", {"num_compromised": 1}, n, kreg, b200={"krum_gram": "tcgen05"})
    check("sketchguard", {"sketch_size": 256, "gamma": 0.6, "alpha": 0.5}, n, kreg)
    check("sketchguard", {"sketch_size": 256, "gamma": 0.6, "alpha": 0.5}, n, kreg, b200={"sketch_dtype": "fp8"})
    check("ubar", {"rho": 0.6, "alpha": 0.5}, n, kreg, b200={"grouped_mlp": False})
    check("evidential_trust", {"trust_threshold": 0.05, "self_weight": 0.6}, n, {"type": "fully", "num_nodes": n}, attack=None,
          b200={"grouped_mlp": False}, tol=3e-3)                   # fused scoring tape = tcgen05 TF32, like the grouped MLP below
    # grouped tcgen05 forward reading candidate weights in place from peer arenas (TF32 → looser tolerance on the trust weights)
    check("evidential_trust", {"trust_threshold": 0.05, "self_weight": 0.6}, n, {"type": "fully", "num_nodes": n}, attack=None,
          b200={"grouped_mlp": True}, tol=3e-3)
    train_check()
    dist.barrier()
    if dist.get_rank() == 0:
        print("ALL MULTI-GPU CHECKS PASSED", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
