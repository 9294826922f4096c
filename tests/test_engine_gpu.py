"""B200 engine on one GPU: fused aggregation vs the CPU oracle classes, end-to-end training, checkpoints."""
import copy
import math
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.utils.factories import (build_aggregator_factory, build_criterion, build_dataset_adapter,
                                          build_model_factory)


def _cfg(algo="fedavg", params=None, n=6, topo=None, attack=None, model=None, data=None, b200=None, rounds=3, **extra):
    base = {"experiment": {"name": "t", "rounds": rounds, "seed": 3},
            "topology": topo or {"type": "k-regular", "num_nodes": n, "k": 2},
            "aggregation": {"algorithm": algo, "params": params or {}},
            "training": {"batch_size": 32, "lr": 0.05, "local_epochs": 1},
            "data": data or {"adapter": "synthetic.mnist", "params": {"samples_per_node": 24, "partition_method": "iid"}},
            "model": model or {"factory": "models.mlp", "params": {"hidden_dims": [32]}},
            "backend": "b200", "b200": b200 or {}}
    if attack:
        base["attack"] = attack
    base.update(extra)
    return Config(**base)


def _build(cfg):
    adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    crit, evid = build_criterion(cfg)
    net = Network.from_config(cfg, mf, adapter, build_aggregator_factory(cfg, mf, torch.device("cuda")),
                              device=torch.device("cuda"), criterion=crit, evidential=evid)
    return net, adapter, mf


HAR = {"factory": "examples.wearables.uci_har", "params": {"input_dim": 561, "num_classes": 6, "hidden_dims": [32, 16]}}
HAR_DATA = {"adapter": "wearables.uci_har", "params": {"data_path": "synthetic", "samples_per_node": 24, "partition_method": "iid"}}


def _oracle_round(net, cfg, adapter, mf, parity=0):
    """Run the fused aggregation once and the CPU aggregator classes on the same inputs; return both results."""
    L = net.layout
    for vn in net.nodes:                                             # decorrelate node states + non-trivial int buffers
        for e in L.float_entries():                                  # (only real entries: padding must stay zero)
            net.live[vn.slot, e.offset:e.offset + e.numel] += 0.05 * (vn.gid + 1) * torch.randn(e.numel, device=net.device)
            if e.name.endswith("running_var"):                       # honest running variances stay positive, as after real training
                net.live[vn.slot, e.offset:e.offset + e.numel] = net.live[vn.slot, e.offset:e.offset + e.numel].abs() + 0.5
        if L.Pi:
            net.ints[vn.slot] = torch.arange(L.Pi, device=net.device) + 3 * vn.gid
    own = {vn.gid: {k: v.detach().cpu().clone() for k, v in L.row_views(net.live[vn.slot], net.ints[vn.slot]).items()} for vn in net.nodes}
    net._aggregate(parity=parity)
    torch.cuda.synchronize()
    pub = {}
    for vn in net.nodes:
        row = net.arena.pub[parity, vn.slot]
        st = {k: v.detach().cpu().clone() for k, v in L.row_views(row, None).items()}
        for e in L.int_entries():
            st[e.name] = row[L.Pf_pad + e.offset: L.Pf_pad + e.offset + e.numel].round().long().view(e.shape).cpu()
        pub[vn.gid] = st
    got = {vn.gid: {k: v.detach().cpu().clone() for k, v in L.row_views(net.live[vn.slot], net.ints[vn.slot]).items()} for vn in net.nodes}
    agg_factory = build_aggregator_factory(cfg, mf, torch.device("cpu"))
    crit, evid = build_criterion(cfg)
    want = {}
    for vn in net.nodes:
        agg = agg_factory(vn.gid)
        template = mf()
        loader = [(vn.X.cpu(), vn.y.cpu())]
        nbrs = {j: pub[j] for j in net.topology.neighbors[vn.gid]}
        out = agg.aggregate(node_id=vn.gid, own_state=own[vn.gid], neighbor_states=nbrs, round_num=net.round_idx,
                            train_loader=loader, model_template=template, device=torch.device("cpu"))
        template.load_state_dict(out)                                # int buffers are cast back like load_state_dict does
        want[vn.gid] = {k: v.clone() for k, v in template.state_dict().items()}
    return got, want, own, pub


def _fill(net, values):
    for e in net.layout.float_entries():
        for slot, val in enumerate(values):
            net.live[slot, e.offset:e.offset + e.numel] = val


def _assert_states_close(got, want, atol=2e-5):
    for gid in want:
        for k in want[gid]:
            a, b = got[gid][k], want[gid][k]
            if b.is_floating_point():
                assert torch.allclose(a.float(), b.float(), rtol=1e-4, atol=atol), (gid, k, (a.float() - b.float()).abs().max())
            else:
                assert torch.equal(a.long(), b.long()), (gid, k, a, b)


@pytest.mark.parametrize("algo,params,attack", [
    ("fedavg", {}, None),
    ("fedavg", {}, {"enabled": True, "type": "directed_deviation", "percentage": 0.34, "params": {"lambda_param": -5.0}}),
    ("balance", {"gamma": 0.3, "kappa": 1.0, "alpha": 0.5}, {"enabled": True, "type": "directed_deviation", "percentage": 0.34}),
    ("krum", {"num_compromised": 1}, {"enabled": True, "type": "gaussian", "percentage": 0.34, "params": {"noise_std": 1.0}}),
    ("sketchguard", {"sketch_size": 256, "gamma": 0.3, "alpha": 0.5}, {"enabled": True, "type": "directed_deviation", "percentage": 0.34}),
    ("ubar", {"rho": 0.6, "alpha": 0.5}, {"enabled": True, "type": "directed_deviation", "percentage": 0.34}),
])
def test_fused_aggregation_matches_cpu_oracle_bn_model(algo, params, attack):
    cfg = _cfg(algo, params, n=6, topo={"type": "k-regular", "num_nodes": 6, "k": 4}, attack=attack, model=HAR, data=HAR_DATA,
               b200={"krum_gram": "fp32", "grouped_mlp": False})
    net, adapter, mf = _build(cfg)
    try:
        got, want, own, pub = _oracle_round(net, cfg, adapter, mf)
        _assert_states_close(got, want)
        if attack and attack["type"] == "directed_deviation":
            byz = sorted(net.compromised)[0]
            k = next(iter(own[byz]))
            assert torch.allclose(pub[byz][k], -5.0 * own[byz][k], atol=1e-5)
    finally:
        net.close()


@pytest.mark.parametrize("attack", [None, {"enabled": True, "type": "directed_deviation", "percentage": 0.34, "params": {"lambda_param": -5.0}},
                                    {"enabled": True, "type": "gaussian", "percentage": 0.34, "params": {"noise_std": 0.5}}])
def test_fullmesh_rank_sum_fedavg_matches_cpu_oracle(attack):
    """Fully connected FedAvg takes the publish_sum → fedavg_fullmesh path (one per-rank sum row); same result as the reference
    aggregator on the published states, and as the general edge-list gather."""
    outs = []
    for rank_sum in (True, False):
        cfg = _cfg("fedavg", {}, n=6, topo={"type": "fully", "num_nodes": 6}, attack=attack, model=HAR, data=HAR_DATA,
                   b200={"fullmesh_rank_sum": rank_sum})
        net, adapter, mf = _build(cfg)
        try:
            got, want, own, pub = _oracle_round(net, cfg, adapter, mf)
            _assert_states_close(got, want)
            assert bool(net._last_et.get("rank_sum")) == rank_sum
            outs.append(got)
        finally:
            net.close()
    _assert_states_close(outs[0], outs[1])


def test_ubar_rejects_nan_scoring_candidates_like_reference():
    """A directed-deviation attacker (λ = −5) also flips the sign of the BatchNorm running variance, so its model evaluates to NaN
    in the reference and UBAR's stage 2 rejects it (NaN ≤ own is false).  The fused scoring tape must propagate that NaN (ReLU /
    max-pool written NaN-preserving), not launder it into a finite loss."""
    cfg = _cfg("ubar", {"rho": 0.6, "alpha": 0.5}, n=24, topo={"type": "k-regular", "num_nodes": 24, "k": 4},
               attack={"enabled": True, "type": "directed_deviation", "percentage": 0.3, "params": {"lambda_param": -5.0}}, model=HAR, data=HAR_DATA,
               b200={"grouped_mlp": False})
    net, adapter, mf = _build(cfg)
    try:
        got, want, own, pub = _oracle_round(net, cfg, adapter, mf)
        _assert_states_close(got, want)
    finally:
        net.close()


def test_fused_evidential_trust_matches_cpu_oracle():
    cfg = _cfg("evidential_trust", {"trust_threshold": 0.05, "self_weight": 0.6, "accuracy_weight": 0.7}, n=5,
               topo={"type": "fully", "num_nodes": 5}, model=HAR, data=HAR_DATA, b200={"grouped_mlp": False})
    net, adapter, mf = _build(cfg)
    try:
        got, want, _, _ = _oracle_round(net, cfg, adapter, mf)
        _assert_states_close(got, want, atol=5e-5)
    finally:
        net.close()


@pytest.mark.parametrize("algo,params", [("evidential_trust", {"trust_threshold": 0.05, "self_weight": 0.6, "accuracy_weight": 0.7}),
                                         ("ubar", {"rho": 0.6, "alpha": 0.5})])
def test_grouped_tcgen05_mlp_scoring_path(algo, params):
    """Same aggregation with the foreign-model scores computed by the grouped tcgen05 (TF32) forward."""
    cfg = _cfg(algo, params, n=5, topo={"type": "fully", "num_nodes": 5}, model=HAR, data=HAR_DATA, b200={"grouped_mlp": True})
    net, adapter, mf = _build(cfg)
    try:
        assert net._mlp_plan is not None and [l["act"] for l in net._mlp_plan] == [1, 1, 2]
        got, want, _, _ = _oracle_round(net, cfg, adapter, mf)
        _assert_states_close(got, want, atol=3e-3)            # trust weights carry TF32 (10-bit mantissa) forward error
    finally:
        net.close()


def test_krum_tcgen05_gram_path_agrees_with_fp32_path():
    res = {}
    for mode in ("fp32", "tcgen05"):
        cfg = _cfg("krum", {"num_compromised": 1}, n=10, topo={"type": "k-regular", "num_nodes": 10, "k": 4},
                   attack={"enabled": True, "type": "gaussian", "percentage": 0.3, "params": {"noise_std": 5.0}},
                   model={"factory": "models.mlp", "params": {"hidden_dims": [64]}}, b200={"krum_gram": mode})
        torch.manual_seed(0)
        net, adapter, mf = _build(cfg)
        try:
            got, want, _, _ = _oracle_round(net, cfg, adapter, mf)
            _assert_states_close(got, want)
            res[mode] = got
        finally:
            net.close()
    for gid in res["fp32"]:
        for k in res["fp32"][gid]:
            assert torch.equal(res["fp32"][gid][k], res["tcgen05"][gid][k])


def test_end_to_end_training_and_contract(capsys):
    cfg = _cfg("fedavg", n=8, topo={"type": "ring", "num_nodes": 8}, rounds=4,
               data={"adapter": "synthetic.mnist", "params": {"samples_per_node": 128, "partition_method": "dirichlet", "alpha": 0.5}})
    net, _, _ = _build(cfg)
    try:
        hist = net.train(rounds=4, local_epochs=1, lr=0.05, verbose=True)
        out = capsys.readouterr().out
        assert hist["round"] == [1, 2, 3, 4] and hist["mean_accuracy"][-1] > hist["mean_accuracy"][0]
        assert hist["mean_accuracy"][-1] > 0.5
        assert re.search(r"Round 4: Mean Accuracy = \d\.\d{4} ± \d\.\d{4}", out) and "=== Round 1/4 ===" in out
        assert set(net.get_node_statistics()) == set(range(8))
    finally:
        net.close()


def test_profile_mode_reports_phase_split_and_roofline_fraction():
    cfg = _cfg("balance", n=6, rounds=3, b200={"profile": True})
    net, _, _ = _build(cfg)
    try:
        net.train(rounds=3, local_epochs=1, lr=0.05)
        line = net.perf_summary()
        assert re.search(r"train \d+\.\d+ ms  aggregate \d+\.\d+ ms  eval \d+\.\d+ ms", line)
        assert "of" in line and "GB/s measured" in line
        row = net.layout.Pf_pad * 4                       # k-regular(2): 3 edges per node incl. self; balance streams neighbours twice
        assert net.timers["hbm_bytes"] == pytest.approx(3 * row * (3 * 6 + 2.0 * 18))
        assert net.timers.get("nvlink_bytes", 0.0) == 0.0
    finally:
        net.close()


@pytest.mark.parametrize("algo,params", [("fedavg", {}), ("krum", {"num_compromised": 1}), ("balance", {"gamma": 0.5, "kappa": 1.0, "alpha": 0.5})])
def test_seed_parity_mode_tracks_the_simulation_backend_round_by_round(algo, params):
    """``b200.seed_parity``: model init and per-round shuffles come from the same host RNG stream as the simulation backend, so an
    attack-free run is comparable round by round.  Autograd training path (same fp32 PyTorch kernels as the simulation): the
    accuracy histories agree to 1e-3; fused tcgen05 path (TF32 operands): to a few samples."""
    from murmura_b200.utils.seed import set_seed
    data = {"adapter": "synthetic.mnist", "params": {"samples_per_node": 192, "partition_method": "dirichlet", "alpha": 0.5}}
    topo = {"type": "k-regular", "num_nodes": 6, "k": 4}
    hist = {}
    for name, backend, b200 in (("sim", "simulation", {}), ("autograd", "b200", {"seed_parity": True, "fused_train": False, "krum_gram": "fp32"}),
                                ("fused", "b200", {"seed_parity": True, "krum_gram": "fp32"})):
        set_seed(11)
        cfg = _cfg(algo, params, n=6, topo=topo, data=data, b200=b200, backend=backend, rounds=10)
        net, _, _ = _build(cfg)
        try:
            hist[name] = net.train(rounds=10, local_epochs=2, lr=0.05)["mean_accuracy"]
        finally:
            if hasattr(net, "close"):
                net.close()
    assert hist["sim"][-1] > hist["sim"][1] and (algo != "fedavg" or hist["sim"][-1] > 0.8)
    assert max(abs(a - b) for a, b in zip(hist["sim"], hist["autograd"])) <= 1e-3, (hist["sim"], hist["autograd"])
    assert max(abs(a - b) for a, b in zip(hist["sim"], hist["fused"])) <= 2e-2, (hist["sim"], hist["fused"])


def test_graphs_match_eager_and_simulation_statistically():
    accs = {}
    data = {"adapter": "synthetic.mnist", "params": {"samples_per_node": 128, "partition_method": "iid"}}
    for graphs in (True, False):
        cfg = _cfg("fedavg", n=4, topo={"type": "fully", "num_nodes": 4}, b200={"cuda_graphs": graphs, "streams": 4, "fused_train": False},
                   data=data)
        net, _, _ = _build(cfg)
        try:
            accs[graphs] = net.train(rounds=8, lr=0.05)["mean_accuracy"]
        finally:
            net.close()
    assert accs[True] == accs[False]            # same RNG streams, same kernels: CUDA graphs on 4 streams are bit-identical to eager
    cfg = _cfg("fedavg", n=4, topo={"type": "fully", "num_nodes": 4}, data=data)        # default: the fused tcgen05 program
    net, _, _ = _build(cfg)
    try:
        accs["fused"] = net.train(rounds=8, lr=0.05)["mean_accuracy"]
        assert getattr(net, "fused", None) is not None and net.fused.be.tma_launches > 0
    finally:
        net.close()
    assert abs(accs["fused"][-1] - accs[True][-1]) < 0.12 and accs["fused"][-1] > 0.6, accs
    sim_cfg = _cfg("fedavg", n=4, topo={"type": "fully", "num_nodes": 4}, data=data, backend="simulation")
    adapter = build_dataset_adapter(sim_cfg); mf = build_model_factory(sim_cfg)
    torch.manual_seed(3)
    sim = Network.from_config(sim_cfg, mf, adapter, build_aggregator_factory(sim_cfg, mf), device=torch.device("cpu"))
    accs["sim"] = sim.train(rounds=8, lr=0.05)["mean_accuracy"]
    assert abs(accs[True][-1] - accs["sim"][-1]) < 0.12 and accs[True][-1] > 0.6, accs


def test_evidential_training_with_device_annealing():
    """BN + dropout + fused evidential loss (device-side annealing) inside CUDA graphs tracks the CPU simulation."""
    data = {"adapter": "wearables.uci_har", "params": {"data_path": "synthetic", "samples_per_node": 256, "partition_method": "iid"}}
    kw = dict(n=4, topo={"type": "ring", "num_nodes": 4}, model=HAR, data=data, rounds=6)
    cfg = _cfg("fedavg", **kw)
    cfg.training.local_epochs = 2
    net, _, _ = _build(cfg)
    try:
        hist = net.train(rounds=6, local_epochs=2, lr=0.05)
    finally:
        net.close()
    assert len(hist["mean_vacuity"]) == 6 and 0 < hist["mean_vacuity"][-1] <= 1.0 and hist["mean_strength"][-1] >= 6.0
    sim_cfg = _cfg("fedavg", backend="simulation", **kw)
    adapter = build_dataset_adapter(sim_cfg); mf = build_model_factory(sim_cfg)
    crit, evid = build_criterion(sim_cfg)
    torch.manual_seed(3)
    sim = Network.from_config(sim_cfg, mf, adapter, build_aggregator_factory(sim_cfg, mf), device=torch.device("cpu"),
                              criterion=crit, evidential=evid)
    ref = sim.train(rounds=6, local_epochs=2, lr=0.05)
    assert hist["mean_accuracy"][-1] > 1.0 / 6 + 0.1, hist["mean_accuracy"]
    assert abs(hist["mean_accuracy"][-1] - ref["mean_accuracy"][-1]) < 0.15, (hist["mean_accuracy"], ref["mean_accuracy"])
    assert abs(hist["mean_vacuity"][-1] - ref["mean_vacuity"][-1]) < 0.05


def test_byzantine_nodes_frozen_and_checkpoint_roundtrip(tmp_path):
    cfg = _cfg("balance", {"gamma": 0.5}, n=6, attack={"enabled": True, "type": "gaussian", "percentage": 0.34, "params": {"noise_std": 10.0}})
    net, _, _ = _build(cfg)
    try:
        hist = net.train(rounds=2, lr=0.05)
        assert len(hist["honest_accuracy"]) == 2 and len(hist["compromised_accuracy"]) == 2
        net.save_checkpoint(str(tmp_path / "ck"))
        snap = net.live.clone(); r = net.round_idx
        net.train(rounds=1, lr=0.05)
        assert not torch.equal(snap, net.live)
        net.load_checkpoint(str(tmp_path / "ck"))
        assert torch.equal(snap[: net.V], net.live[: net.V]) and net.round_idx == r
    finally:
        net.close()


def test_fault_injection_drops_edge():
    cfg = _cfg("fedavg", n=3, topo={"type": "fully", "num_nodes": 3}, b200={"fault_drop_edges": {0: [[1, 0]]}})
    net, adapter, mf = _build(cfg)
    try:
        L = net.layout
        _fill(net, [1.0, 2.0, 6.0])
        net._aggregate(parity=0); torch.cuda.synchronize()
        assert net.live[0, 0].item() == pytest.approx(3.5)           # node 0 lost the edge from node 1 → mean(1, 6)
        assert net.live[1, 0].item() == pytest.approx(3.0) and net.live[2, 0].item() == pytest.approx(3.0)
    finally:
        net.close()


def test_mobility_and_dmtt_on_device():
    cfg = _cfg("fedavg", n=8, topo={"type": "ring", "num_nodes": 8}, rounds=3,
               attack={"enabled": True, "type": "topology_liar", "percentage": 0.25, "params": {"model_attack_type": "gaussian", "noise_std": 1.0}},
               mobility={"comm_range": 45.0, "seed": 1}, dmtt={"budget_B": 2})
    net, _, _ = _build(cfg)
    try:
        hist = net.train(rounds=3, lr=0.05)
        assert hist["round"] == [1, 2, 3]
        collab = net.collab.cpu().numpy()
        assert collab.shape == (8, 8) and (collab.sum(1) <= 2).all() and collab.diagonal().sum() == 0
        honest = [i for i in range(8) if i not in net.compromised]
        liars = sorted(net.compromised)
        # trust in liars' topology claims decays relative to honest peers
        a, b = net.t_alpha.cpu().numpy(), net.t_beta.cpu().numpy()
        R = a / (a + b)
        seen_liar = [R[i, j] for i in honest for j in liars if b[i, j] != 1.0 or a[i, j] != 1.0]
        seen_honest = [R[i, j] for i in honest for j in honest if i != j and (b[i, j] != 1.0 or a[i, j] != 1.0)]
        if seen_liar and seen_honest:
            assert np.mean(seen_liar) <= np.mean(seen_honest) + 1e-6
    finally:
        net.close()


def test_custom_aggregator_and_attack_fall_back_to_generic_path():
    from murmura_b200.aggregation import Aggregator

    class TakeMax(Aggregator):
        def aggregate(self, node_id, own_state, neighbor_states, round_num, **kw):
            out = {}
            for k, v in own_state.items():
                out[k] = torch.stack([v] + [s[k] for s in neighbor_states.values()]).max(0).values if v.is_floating_point() else v
            return out

    cfg = _cfg("fedavg", n=3, topo={"type": "fully", "num_nodes": 3})
    adapter = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    net = Network.from_config(cfg, mf, adapter, lambda nid: TakeMax(), device=torch.device("cuda"))
    try:
        assert net.family == "generic"
        L = net.layout
        _fill(net, [1.0, 2.0, 6.0])
        net._aggregate(parity=0); torch.cuda.synchronize()
        for e in L.float_entries():
            assert (net.live[:, e.offset:e.offset + e.numel] == 6.0).all()
    finally:
        net.close()


def test_split_backward_gradients_match_autograd_eager_and_graph():
    """Weight gradients computed on the side stream (SplitBackward) equal stock autograd's, eagerly and inside a CUDA graph."""
    import torch.nn as nn
    import torch.nn.functional as F
    from murmura_b200.models.resnet import ResNet18
    from murmura_b200.parallel.split_backward import SplitBackward
    torch.manual_seed(11)
    dev = torch.device("cuda")
    for make, shape in ((lambda: ResNet18(), (16, 3, 32, 32)),
                        (lambda: nn.Sequential(nn.Conv2d(1, 8, 5, padding=2), nn.ReLU(), nn.MaxPool2d(2), nn.Flatten(),
                                               nn.Linear(8 * 14 * 14, 32), nn.ReLU(), nn.Linear(32, 10)), (16, 1, 28, 28))):
        model = make().to(dev).to(memory_format=torch.channels_last).train()
        ref = copy.deepcopy(model)
        x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, 10, (shape[0],), device=dev)
        F.cross_entropy(ref(x), y).backward()
        want = [p.grad.clone() for p in ref.parameters()]
        params = list(model.parameters())
        sb = SplitBackward(dev)
        c0, l0 = F.conv2d, F.linear

        def step():
            for p in params:
                p.grad = None
            with sb:
                out = model(x)
            F.cross_entropy(out, y).backward()
            return sb.join(params)

        got = step()
        assert F.conv2d is c0 and F.linear is l0
        torch.cuda.synchronize()
        for g, w, p in zip(got, want, params):
            assert g.shape == p.shape and g.stride() == p.stride()
            assert float((g - w).abs().max()) <= 2e-3 * (float(w.abs().max()) + 1e-6)
        # graph capture: fork/join of the side stream become graph edges; replay must reproduce the gradients
        model.load_state_dict(ref.state_dict())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
        torch.cuda.current_stream().wait_stream(s)
        model.load_state_dict(ref.state_dict())
        static = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            static = step()
        for t in static:
            t.zero_()
        model.load_state_dict(ref.state_dict())
        g.replay(); torch.cuda.synchronize()
        for gt, w in zip(static, want):
            assert float((gt - w).abs().max()) <= 2e-3 * (float(w.abs().max()) + 1e-6)


@pytest.mark.parametrize("split", [True, False, "auto"])
def test_training_with_split_backward_modes(split):
    cfg = _cfg("fedavg", n=4, topo={"type": "ring", "num_nodes": 4}, rounds=3, b200={"split_backward": split, "fused_train": False},
               data={"adapter": "synthetic.mnist", "params": {"samples_per_node": 96, "partition_method": "iid"}})
    net, _, _ = _build(cfg)
    try:
        hist = net.train(rounds=3, local_epochs=1, lr=0.05)
        assert len(hist["round"]) == 3 and all(math.isfinite(a) for a in hist["mean_accuracy"])
        assert hist["mean_accuracy"][-1] > 0.25                # 9 SGD steps, 10 classes; identical in all three modes
        used = any(vn.split_bwd is not None for vn in net.nodes)
        assert used == (split is True)                     # 4 nodes on one GPU: "auto" keeps the single-stream backward
    finally:
        net.close()


@pytest.mark.parametrize("gram", ["fp32", "tcgen05"])
def test_krum_on_converged_models_matches_fp64_oracle(gram):
    """Honest models that have (almost) converged + two far outliers: the distance table and the Krum winner of every node must
    match an fp64 oracle.  The TF32 Gram alone cannot resolve ‖a−b‖² ≪ ‖a‖² (cancellation): the tcgen05 path recomputes those
    pairs exactly (``krum_refine_kernel``), ``auto`` / ``fp32`` use exact differences throughout."""
    n = 8
    cfg = _cfg("krum", params={"num_compromised": 1}, n=n, topo={"type": "fully", "num_nodes": n},
               model={"factory": "models.mlp", "params": {"hidden_dims": [256]}}, b200={"krum_gram": gram})
    net, _, _ = _build(cfg)
    try:
        L = net.layout
        g = torch.Generator(device="cuda").manual_seed(5)
        base = torch.randn(L.Pf, device=net.device, generator=g)
        for vn in net.nodes:
            noise = 1e-3 if vn.gid not in (2, 5) else 0.5
            net.live[vn.slot, : L.Pf] = base + noise * torch.randn(L.Pf, device=net.device, generator=g)
        X = net.live[: net.V, : L.Pf].double().cpu()
        net._aggregate(parity=0)
        torch.cuda.synchronize()
        et = net._last_et
        rows, gids = et["host_rows"], et["host_gid"]
        D = et["krum_D"].cpu().double()
        win = et["krum_win"].cpu()
        for vi, vn in enumerate(net.nodes):
            ids = gids[rows[vi]:rows[vi + 1]]
            A = X[[net.placement.slot_of[j] for j in ids]]
            D64 = torch.cdist(A, A) ** 2
            m = len(ids)
            got = D[vi, :m, :m]
            assert torch.allclose(got, D64, rtol=2e-3, atol=1e-6), (gram, vi, (got - D64).abs().max())
            k = max(1, m - 1 - 2)
            score = (D64 + torch.diag(torch.full((m,), float("inf")))).sqrt().sort(dim=1).values[:, :k].sum(1)
            assert int(win[vi]) == int(score.argmin()), (gram, vi)
        if gram == "tcgen05":
            assert float(et["krum_refined"].sum()) > 0          # the near-identical honest pairs went through the exact recomputation
    finally:
        net.close()
