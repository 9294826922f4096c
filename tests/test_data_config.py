import json

import numpy as np
import pytest
import torch
from pydantic import ValidationError

from murmura_b200.config import Config, load_config, save_config
from murmura_b200.data import (DatasetAdapter, SyntheticAdapter, combine_partitions_with_dirichlet, dirichlet_partition,
                               iid_partition, natural_partition)


def _base():
    return {"experiment": {"name": "t"}, "topology": {"type": "ring", "num_nodes": 3}, "aggregation": {"algorithm": "fedavg"},
            "training": {}, "data": {"adapter": "synthetic.mnist"}, "model": {"factory": "models.mlp"}}


def test_schema_defaults_and_rejects():
    c = Config(**_base())
    assert c.experiment.seed == 42 and c.experiment.rounds == 20 and c.topology.seed == 12345
    assert c.training.batch_size == 64 and c.training.lr == 0.01 and c.backend == "simulation"
    assert c.distributed.round_duration_s == 60.0 and c.attack.enabled is False and c.mobility is None
    assert c.b200.transport == "auto" and c.b200.fused_train == "auto"
    with pytest.raises(ValidationError):
        Config(**{**_base(), "bogus": 1})                                  # unknown top-level key
    ok = _base(); ok["topology"]["whatever"] = 3; Config(**ok)              # unknown sub-key ignored
    bad = _base(); bad["aggregation"] = {"algorithm": "median"}
    with pytest.raises(ValidationError):
        Config(**bad)
    c2 = Config(**{**_base(), "backend": "b200", "dmtt": {}, "mobility": {"comm_range": 40}})
    assert c2.dmtt.budget_B == 5 and c2.mobility.comm_range == 40 and c2.dmtt.lambda4 == 0.1


def test_loader_roundtrip(tmp_path):
    c = Config(**_base())
    for name in ("a.yaml", "b.yml", "c.json"):
        p = tmp_path / name
        save_config(c, p)
        assert load_config(p).model_dump() == c.model_dump()
    with pytest.raises(ValueError):
        save_config(c, tmp_path / "x.toml")
    with pytest.raises(FileNotFoundError):
        load_config(tmp_path / "missing.yaml")
    (tmp_path / "d.txt").write_text("x")
    with pytest.raises(ValueError):
        load_config(tmp_path / "d.txt")


def test_bundled_configs_load():
    import glob, os
    root = os.path.join(os.path.dirname(os.path.dirname(__file__)), "murmura_b200", "examples", "configs")
    files = glob.glob(os.path.join(root, "*.yaml"))
    assert len(files) >= 8
    for f in files:
        load_config(f)


def test_dirichlet_partition_properties():
    labels = np.repeat(np.arange(6), 100)
    parts = dirichlet_partition(labels, 5, alpha=0.1, seed=7)
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(600))                        # exact cover
    assert parts == dirichlet_partition(labels, 5, alpha=0.1, seed=7)
    assert all(len(p) >= 1 for p in parts)
    skew = [np.bincount(labels[p], minlength=6).max() / max(len(p), 1) for p in parts]
    assert np.mean(skew) > 0.4                             # alpha=0.1 is strongly non-IID


def test_min_samples_rebalance():
    labels = np.zeros(40, dtype=int)
    parts = dirichlet_partition(labels, 8, alpha=0.01, min_samples_per_client=3, seed=0)
    assert all(len(p) >= 3 for p in parts) and sum(map(len, parts)) == 40


def test_iid_natural_combine():
    parts = iid_partition(103, 4, seed=1)
    assert sorted(map(len, parts)) == [25, 26, 26, 26] and sorted(i for p in parts for i in p) == list(range(103))
    ids = np.array([5, 5, 2, 9, 2, 9, 9])
    nat, k = natural_partition(ids)
    assert k == 3 and nat == [[2, 4], [0, 1], [3, 5, 6]]
    nat2, k2 = natural_partition(ids, num_clients=2)
    assert k2 == 2 and nat2 == [[2, 4], [0, 1]]
    labels = np.array([0, 1, 0, 1, 0, 1, 0])
    comb = combine_partitions_with_dirichlet(nat, labels, 2, alpha=1.0, seed=3)
    assert sorted(i for p in comb for i in p) == list(range(7))


def test_adapter_and_synthetic():
    ad = SyntheticAdapter(name="cifar10", num_nodes=4, samples_per_node=20, seed=1)
    assert ad.get_num_clients() == 4 and sum(len(p) for p in ad.get_client_partitions()) == 80
    x, y = ad.get_client_data(0)[0]
    assert x.shape == (3, 32, 32)
    X, Y = ad.client_tensors(1)
    assert X.shape[0] == len(ad.get_client_partitions()[1]) and Y.dtype == torch.long
    with pytest.raises(ValueError):
        ad.get_client_data(9)
    a2 = SyntheticAdapter(name="cifar10", num_nodes=4, samples_per_node=20, seed=1)
    assert torch.equal(a2.dataset.tensors[0], ad.dataset.tensors[0])
    with pytest.raises(ValueError):
        SyntheticAdapter(name="nope")
    generic = DatasetAdapter([(torch.zeros(2), 1), (torch.ones(2), 0)], [[0], [1]])
    X, Y = generic.client_tensors(1)
    assert X.tolist() == [[1.0, 1.0]] and Y.tolist() == [0]


def test_wearables_window_helper_and_synthetic_route():
    from murmura_b200.examples.wearables.datasets import majority_windows
    feats = np.arange(20, dtype=np.float32).reshape(10, 2)
    acts = np.array([1, 1, 1, 2, 2, 2, 2, 3, 3, 3])
    w, l = majority_windows(feats, acts, 4, 2, {1: 0, 2: 1})
    assert w.shape == (3, 8) and l.tolist() == [0, 1, 1][: len(l)]
    np.testing.assert_array_equal(w[0], feats[0:4].ravel())
    from murmura_b200.examples.wearables import load_wearable_adapter, get_wearable_dataset_info
    ad = load_wearable_adapter("uci_har", "synthetic", num_nodes=3, samples_per_node=10)
    assert ad.client_tensors(0)[0].shape[1] == 561
    assert get_wearable_dataset_info("PPG-DaLiA")["num_classes"] == 7
    with pytest.raises(ValueError):
        get_wearable_dataset_info("nope")


def test_leaf_partitions_and_json_loader(tmp_path):
    from murmura_b200.examples.leaf.datasets import LEAFFEMNISTDataset, create_leaf_client_partitions
    for split, k in (("train", 3), ("test", 1)):
        d = tmp_path / split; d.mkdir()
        users = [f"u{i}" for i in range(5)]
        blob = {"users": users, "num_samples": [k + i for i in range(5)],
                "user_data": {u: {"x": [[0.5] * 784] * (k + i), "y": [i] * (k + i)} for i, u in enumerate(users)}}
        (d / "all.json").write_text(json.dumps(blob))
    tr, te = LEAFFEMNISTDataset(str(tmp_path), "train"), LEAFFEMNISTDataset(str(tmp_path), "test")
    assert len(tr) == 3 + 4 + 5 + 6 + 7 and tr[0][0].shape == (1, 28, 28)
    a, b = create_leaf_client_partitions(tr, te, num_nodes=2, seed=42)
    assert sorted(i for p in a for i in p) == list(range(len(tr))) and sorted(i for p in b for i in p) == list(range(len(te)))
    assert create_leaf_client_partitions(tr, te, 2, 42)[0] == a
    from murmura_b200.examples.leaf import load_leaf_adapter
    ad = load_leaf_adapter("femnist", data_path=str(tmp_path), num_nodes=2, seed=42, max_samples=4)
    assert all(len(p) <= 4 for p in ad.get_client_partitions())
    syn = load_leaf_adapter("femnist", data_path="synthetic", num_nodes=2, samples_per_node=8)
    assert syn.client_tensors(0)[0].shape[1:] == (1, 28, 28)


def test_state_layout_and_channels_last_permutation():
    import torch.nn as nn
    from murmura_b200.models import ResNet18, CIFARCNN
    from murmura_b200.parallel.arena import Placement, StateLayout
    m = ResNet18()
    lay = StateLayout.from_model(m)
    assert lay.P_float_real == 11191242 and lay.Pi == 20 and lay.Pp == 11181642
    assert lay.Pf_pad % 256 == 0 and lay.stride == lay.Pf_pad + 256 and lay.Pp4 % 4 == 0
    names = [e.name for e in lay.entries]
    assert names.index("fc.bias") < names.index("bn1.running_mean") < names.index("bn1.num_batches_tracked")
    small = CIFARCNN()
    for cl in (False, True):
        lay = StateLayout.from_model(small, channels_last=cl)
        row = torch.zeros(lay.stride); ints = torch.zeros(max(lay.Pi, 1), dtype=torch.long)
        ref_flat = torch.cat([t.flatten() for t in small.state_dict().values() if t.is_floating_point()])
        lay.bind(small, row, None, ints if lay.Pi else None)
        perm = lay.ref_permutation()
        ok = perm >= 0
        assert ok.sum() == lay.P_float_real and sorted(perm[ok].tolist()) == list(range(lay.P_float_real))
        assert torch.equal(row[: lay.Pf][torch.from_numpy(ok)], ref_flat[torch.from_numpy(perm[ok])])
        views = lay.row_views(row, None)
        assert all(torch.equal(views[k], v) for k, v in small.state_dict().items())       # logical values unchanged
        w = dict(small.named_parameters())["conv1.weight"]
        assert w.data_ptr() == row.data_ptr() and w.is_contiguous(memory_format=torch.channels_last) == cl or not cl
        assert small(torch.zeros(2, 3, 32, 32)).shape == (2, 10)
    p = Placement(20, 8)
    assert p.counts == [3, 3, 3, 3, 2, 2, 2, 2] and p.slots_per_rank == 3 and p.local_nodes(4) == [12, 13]
    assert int(p.rank_of[13]) == 4 and int(p.slot_of[13]) == 1 and Placement(4, 8).counts == [1, 1, 1, 1, 0, 0, 0, 0]
    w = [11, 12, 7, 8, 9, 7, 3, 3]                         # steps per node of the flagship shards
    b = Placement(8, 2, w)
    loads = [sum(w[g] for g in b.local_nodes(r)) for r in range(2)]
    assert sorted(loads) == [30, 30] and sorted(b.local_nodes(0) + b.local_nodes(1)) == list(range(8))
    assert all(int(b.slot_of[g]) == b.local_nodes(int(b.rank_of[g])).index(g) for g in range(8))
    c = Placement(8, 2)
    assert [sum(w[g] for g in c.local_nodes(r)) for r in range(2)] == [38, 22]


def test_evidential_mlp_forward_equals_its_sequential_definition():
    """forward() walks the Sequential itself (BatchNorm→ReLU pairs go through ops.bn_act); on CPU it must equal the plain
    Sequential composition, in train mode (batch statistics, running-stat updates) and eval mode."""
    import copy
    import torch
    from murmura_b200.models.mlp import EvidentialHARClassifier
    torch.manual_seed(0)
    a = EvidentialHARClassifier(input_dim=24, hidden_dims=(16, 8), num_classes=5, dropout=0.0)
    b = copy.deepcopy(a)
    x = torch.randn(12, 24)
    ya = a(x)
    yb = b.evidential_head(b.feature_extractor(x))
    assert torch.allclose(ya, yb, atol=1e-6)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.allclose(va.float(), vb.float(), atol=1e-6), ka
    a.eval(); b.eval()
    assert torch.allclose(a(x), b.evidential_head(b.feature_extractor(x)), atol=1e-6)


def test_bn_act_cpu_fallback_equals_stock_composition():
    """ops.bn_act on CPU tensors (or any non-fusable layout) is the stock BatchNorm → (+residual) → ReLU composition."""
    import copy
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from murmura_b200 import ops
    torch.manual_seed(3)
    a = nn.BatchNorm2d(8); b = copy.deepcopy(a)
    x = torch.randn(5, 8, 4, 4); r = torch.randn_like(x)
    assert not ops.bn_act_fusable(x, a)
    ya = ops.bn_act(x, a, residual=r, relu=True)
    yb = F.relu(b(x) + r)
    assert torch.equal(ya, yb)
    assert torch.equal(a.running_mean, b.running_mean) and int(a.num_batches_tracked) == int(b.num_batches_tracked) == 1
    a.eval(); b.eval()
    assert torch.equal(ops.bn_act(x, a, relu=False), b(x))


def test_split_backward_is_transparent_for_cpu_tensors_and_restores_functional():
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    from murmura_b200.parallel.split_backward import SplitBackward
    conv0, lin0 = F.conv2d, F.linear
    sb = SplitBackward(torch.device("cpu"), side=object())
    model = nn.Sequential(nn.Conv2d(2, 3, 3), nn.Flatten(), nn.Linear(3 * 4 * 4, 2))
    with sb:
        assert F.conv2d is not conv0 and F.linear is not lin0
        model(torch.randn(2, 2, 6, 6)).sum().backward()
    assert F.conv2d is conv0 and F.linear is lin0
    assert all(p.grad is not None for p in model.parameters()) and not sb.stash
    grads = sb.join(list(model.parameters()))                       # nothing stashed: falls back to p.grad
    assert all(g is p.grad for g, p in zip(grads, model.parameters()))


def test_b200_block_new_options_validate():
    import pytest
    from pydantic import ValidationError
    from murmura_b200.config.schema import B200Config
    c = B200Config()
    assert c.split_backward == "auto" and c.fused_bn is True and c.gather_impl == "auto"
    assert B200Config(split_backward=False).split_backward is False
    assert B200Config(split_backward=True, gather_impl="tma", fused_bn=False).gather_impl == "tma"
    with pytest.raises(ValidationError):
        B200Config(gather_impl="dma")
    with pytest.raises(ValidationError):
        B200Config(split_backward="sometimes")


@pytest.mark.parametrize("family", ["evidential", "plain"])
def test_batched_mlp_trainer_equals_per_node_sgd(family):
    """K8b: one batched step over arena-row views == every node's own forward/backward/SGD step (incl. BatchNorm running
    statistics, num_batches_tracked and nodes that sit a step out)."""
    import copy
    import torch
    import torch.nn.functional as F
    from murmura_b200.models.mlp import MLP, EvidentialHARClassifier, evidential_loss_reference
    from murmura_b200.parallel.arena import StateLayout
    from murmura_b200.parallel.batched_mlp import BatchedMLPTrainer
    torch.manual_seed(0)
    V, B, Fin, C, lr, lam = 4, 16, 20, 5, 0.05, 0.3
    make = (lambda: EvidentialHARClassifier(input_dim=Fin, hidden_dims=(12, 8), num_classes=C, dropout=0.0)) if family == "evidential" \
        else (lambda: MLP(input_dim=Fin, hidden_dims=(12,), num_classes=C))
    models = [make() for _ in range(V)]
    layout = StateLayout.from_model(models[0])
    rows = torch.zeros(V, layout.stride)
    ints = torch.zeros(V, max(layout.Pi, 1), dtype=torch.int64)
    for v, m in enumerate(models):                                    # flatten every node's state into its arena row
        sd = m.state_dict()
        for e in layout.float_entries():
            rows[v, e.offset:e.offset + e.numel] = sd[e.name].reshape(-1)
        for e in layout.int_entries():
            ints[v, e.offset:e.offset + e.numel] = sd[e.name].reshape(-1)
    trainer = BatchedMLPTrainer(models[0], layout, rows, ints if layout.Pi else None)
    refs = [copy.deepcopy(m).train() for m in models]
    for step in range(4):
        x = torch.randn(V, B, Fin); y = torch.randint(0, C, (V, B))
        active = torch.tensor([1.0, 1.0, 0.0 if step >= 2 else 1.0, 1.0 if step != 1 else 0.0])
        trainer.step(x, y, active, lr, lam)
        for v, m in enumerate(refs):
            if active[v] == 0:
                continue
            for p in m.parameters():
                p.grad = None
            out = m(x[v])
            loss = evidential_loss_reference(out, y[v], lam) if family == "evidential" else F.cross_entropy(out, y[v])
            loss.backward()
            with torch.no_grad():
                for p in m.parameters():
                    p.add_(p.grad, alpha=-lr)
    for v, m in enumerate(refs):
        got = layout.row_views(rows[v], ints[v] if layout.Pi else None)
        for k, want in m.state_dict().items():
            assert torch.allclose(got[k].double(), want.double(), atol=2e-5, rtol=1e-4), (v, k)
    xp, yp = torch.randn(V, 9, Fin), torch.randint(0, C, (V, 9))
    idx = torch.randint(0, 9, (V, 3))
    gx, gy = BatchedMLPTrainer.gather(xp, yp, idx)
    assert all(torch.equal(gx[v], xp[v][idx[v]]) and torch.equal(gy[v], yp[v][idx[v]]) for v in range(V))


@pytest.mark.parametrize("n,bs,shuffle,drop_last", [(100, 32, True, True), (100, 32, False, False), (64, 64, True, False), (5, 2, True, True)])
def test_fast_tensor_loader_is_bit_identical_to_dataloader(n, bs, shuffle, drop_last):
    """Same batches AND same global-RNG consumption as the stock DataLoader, for several epochs in a row."""
    import torch
    from torch.utils.data import DataLoader, TensorDataset
    from murmura_b200.data.fast_loader import FastTensorLoader
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, 7, generator=g); y = torch.randint(0, 5, (n,), generator=g)
    ref = DataLoader(TensorDataset(x, y), batch_size=bs, shuffle=shuffle, drop_last=drop_last)
    fast = FastTensorLoader(x, y, bs, shuffle=shuffle, drop_last=drop_last)
    assert len(ref) == len(fast)
    torch.manual_seed(123); a = [[(xb.clone(), yb.clone()) for xb, yb in ref] for _ in range(3)]; ra = torch.rand(3)
    torch.manual_seed(123); b = [[(xb.clone(), yb.clone()) for xb, yb in fast] for _ in range(3)]; rb = torch.rand(3)
    assert torch.equal(ra, rb)                                          # the global generator is in the same state afterwards
    for ea, eb in zip(a, b):
        assert len(ea) == len(eb)
        for (xa, ya), (xb_, yb_) in zip(ea, eb):
            assert torch.equal(xa, xb_) and torch.equal(ya, yb_) and ya.dtype == yb_.dtype
    torch.manual_seed(9); first_ref = next(iter(ref)); torch.manual_seed(9); first_fast = next(iter(fast))       # UBAR's next(iter(loader))
    assert torch.equal(first_ref[0], first_fast[0])


def test_replay_round_orders_matches_the_simulation_loaders_rng_consumption():
    """Seed-parity mode of the B200 engine: ``replay_round_orders`` must draw from the global torch stream exactly like iterating
    the simulation backend's shuffling loaders node by node (reference ``murmura/core/network.py:80-92`` + ``DataLoader(shuffle=True)``),
    including drop_last truncation, skipped (compromised) clients and two local epochs — and leave the stream in the same state."""
    import torch
    from murmura_b200.data.fast_loader import FastTensorLoader, replay_round_orders
    sizes, bs, epochs, skip = [70, 33, 5, 64, 1, 200], 32, 2, {3}
    shards = [(torch.arange(n, dtype=torch.float32).view(n, 1), torch.arange(n)) for n in sizes]
    torch.manual_seed(1234)
    want = {}
    for cid, (x, y) in enumerate(shards):
        if cid in skip:
            continue
        n = len(x)
        eb = min(bs, max(2, n))
        loader = FastTensorLoader(x, y, eb, shuffle=True, drop_last=n > eb)
        rows = []
        for _ in range(epochs):
            got = [yb for xb, yb in loader if xb.size(0) >= 2]
            if got:
                rows.append(torch.cat(got))
        if rows:
            want[cid] = torch.stack(rows)
    tail_want = torch.rand(3)
    torch.manual_seed(1234)
    got = replay_round_orders(sizes, bs, epochs, skip=skip)
    tail_got = torch.rand(3)
    assert set(got) == set(want) == {0, 1, 2, 5}
    for cid in want:
        assert torch.equal(got[cid], want[cid]), cid
    assert torch.equal(tail_got, tail_want)


def _store_barrier_worker(rank, world, port, delays, out):
    import time
    import torch.distributed as dist
    from datetime import timedelta
    from murmura_b200.parallel.hostsync import store_barrier
    store = dist.TCPStore("127.0.0.1", port, world, is_master=(rank == 0), timeout=timedelta(seconds=60))
    time.sleep(delays[rank])
    arrive = time.time()
    ok = store_barrier(store, "hb/1/1", world, timeout_s=40.0)
    out.put((rank, "first", ok, arrive, time.time()))
    if rank == 1:                                                   # second barrier: the other ranks never arrive → timeout path
        out.put((rank, "second", store_barrier(store, "hb/1/2", world, timeout_s=0.3), 0.0, 0.0))
    store_barrier(store, "hb/1/done", world, timeout_s=40.0)        # keep the master alive until everybody is done


def test_store_barrier_rendezvous_and_timeout():
    """Host-side rendezvous used between publish and the flag wait (parallel/hostsync.py): nobody passes before the last rank
    arrived, a missing rank times out instead of hanging."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    delays = [0.0, 0.0, 0.8]
    procs = [ctx.Process(target=_store_barrier_worker, args=(r, 3, port, delays, out)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = [out.get(timeout=5) for _ in range(4)]
    first = [g for g in got if g[1] == "first"]
    assert len(first) == 3 and all(g[2] for g in first)
    last_arrival = max(g[3] for g in first)
    assert min(g[4] for g in first) >= last_arrival - 1e-3          # the barrier property (same-host clocks)
    assert [g[2] for g in got if g[1] == "second"] == [False]


def test_bench_only_uses_engine_attributes_that_exist():
    """bench.py runs on the GPU box only; guard its use of the engine's surface on the CPU (a renamed / removed attribute would
    otherwise surface at round end)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bench = open(os.path.join(root, "bench.py")).read()
    engine = open(os.path.join(root, "murmura_b200", "parallel", "engine.py")).read()
    arena = open(os.path.join(root, "murmura_b200", "parallel", "arena.py")).read()
    used = set(re.findall(r"\bnet\.([A-Za-z_]\w*)", bench))
    assert {"train", "close", "timers", "history", "d2h_bytes_per_round", "h2d_bytes_per_round"} <= used
    for name in used:
        assert re.search(rf"def {name}\b|self\.{name}\b\s*[:=]|self\.{name}\b", engine), name
    for sub in re.findall(r"\bnet\.layout\.([A-Za-z_]\w*)", bench):
        assert re.search(rf"\b{sub}\b", arena), sub
    for sub in re.findall(r"\bnet\.opt\.([A-Za-z_]\w*)", bench):
        assert re.search(rf"^\s+{sub}\s*:", open(os.path.join(root, "murmura_b200", "config", "schema.py")).read(), re.M), sub
