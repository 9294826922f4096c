"""Property tests (hypothesis) for the pure-function layer: topologies, partitioners, aggregators, layout, placement."""

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from murmura_b200.aggregation import BALANCEAggregator, FedAvgAggregator, KrumAggregator, UBARAggregator
from murmura_b200.aggregation.base import average_states, compute_model_distance
from murmura_b200.data.partitioners import dirichlet_partition, iid_partition
from murmura_b200.parallel.arena import Placement, StateLayout
from murmura_b200.topology import MobilityModel, create_topology

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(n=st.integers(2, 40), p=st.floats(0.0, 1.0), seed=st.integers(0, 10_000))
def test_erdos_renyi_invariants(n, p, seed):
    t = create_topology("erdos", n, p=p, seed=seed)
    adj = t.adjacency()
    assert (adj == adj.T).all() and not adj.diagonal().any()
    assert all(t.degree(i) >= 1 for i in range(n))                       # isolated nodes are repaired
    assert sorted(t.edges) == t.edges and all(a < b for a, b in t.edges)
    assert sum(map(len, t.neighbors)) == 2 * len(t.edges)
    assert create_topology("erdos", n, p=p, seed=seed).edges == t.edges  # deterministic in the seed


@FAST
@given(n=st.integers(3, 40), k=st.integers(1, 12))
def test_k_regular_invariants(n, k):
    t = create_topology("k-regular", n, k=k)
    k_eff = k + (k % 2)
    if k_eff >= n:
        assert all(t.degree(i) == n - 1 for i in range(n))
    else:
        assert all(t.degree(i) == k_eff for i in range(n)) and t.is_connected()
    row_ptr, cols = t.to_csr()
    assert row_ptr[-1] == len(cols) == sum(map(len, t.neighbors)) + n and all(cols[row_ptr[i]] == i for i in range(n))


@FAST
@given(n=st.integers(2, 24), rng=st.floats(1.0, 80.0), seed=st.integers(0, 1000), r=st.integers(0, 6))
def test_mobility_invariants(n, rng, seed, r):
    m = MobilityModel(n, 100.0, rng, 5.0, seed=seed)
    pos = m.positions_at(r)
    assert pos.shape == (n, 2) and (pos >= 0).all() and (pos < 100.0).all()
    nb = m.neighbors_at(r)
    adj = m.adjacency_at(r)
    assert (adj == adj.T).all()
    assert all(sorted(nb[i]) == np.flatnonzero(adj[i]).tolist() for i in range(n))
    assert all(len(nb[i]) >= 1 for i in range(n))                        # ensure_connected repairs isolated nodes
    step = np.abs(m.positions_at(r + 1) - pos)
    step = np.minimum(step, 100.0 - step)
    assert (step <= 5.0 + 1e-9).all()


@FAST
@given(classes=st.integers(2, 8), per_class=st.integers(5, 40), clients=st.integers(2, 9), alpha=st.floats(0.05, 5.0), seed=st.integers(0, 999))
def test_dirichlet_partition_is_exact_cover(classes, per_class, clients, alpha, seed):
    labels = np.repeat(np.arange(classes), per_class)
    parts = dirichlet_partition(labels, clients, alpha=alpha, min_samples_per_client=1, seed=seed)
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(labels))) and len(parts) == clients
    if len(labels) >= clients:
        assert all(len(p) >= 1 for p in parts)
    parts_iid = iid_partition(len(labels), clients, seed=seed)
    assert max(map(len, parts_iid)) - min(map(len, parts_iid)) <= 1


def _states(k, dim, seed):
    g = torch.Generator().manual_seed(seed)
    return [{"w": torch.randn(dim, generator=g), "n": torch.tensor(i + 1)} for i in range(k)]


@FAST
@given(k=st.integers(1, 6), dim=st.integers(1, 30), seed=st.integers(0, 999))
def test_fedavg_is_permutation_invariant_mean(k, dim, seed):
    own, *others = _states(k + 1, dim, seed)
    nb = {i: s for i, s in enumerate(others)}
    out = FedAvgAggregator().aggregate(0, own, nb, 0)
    ref = torch.stack([own["w"]] + [s["w"] for s in others]).mean(0)
    assert torch.allclose(out["w"], ref, atol=1e-6) and out["n"].item() == own["n"].item()
    rev = {i: s for i, s in reversed(list(nb.items()))}
    assert torch.allclose(FedAvgAggregator().aggregate(0, own, rev, 0)["w"], out["w"], atol=1e-6)


@FAST
@given(k=st.integers(3, 7), dim=st.integers(2, 20), seed=st.integers(0, 999), c=st.integers(0, 2))
def test_krum_returns_one_of_the_candidates(k, dim, seed, c):
    own, *others = _states(k + 1, dim, seed)
    out = KrumAggregator(num_compromised=c).aggregate(0, own, {i: s for i, s in enumerate(others)}, 0)
    assert any(out is s for s in [own, *others])
    if c >= (k + 1 - 2) / 2:
        assert out is own


@FAST
@given(k=st.integers(1, 6), dim=st.integers(2, 20), seed=st.integers(0, 999), alpha=st.floats(0.0, 1.0))
def test_balance_and_ubar_outputs_are_convex_blends(k, dim, seed, alpha):
    own, *others = _states(k + 1, dim, seed)
    nb = {i: s for i, s in enumerate(others)}
    lo = torch.stack([own["w"]] + [s["w"] for s in others]).min(0).values - 1e-5
    hi = torch.stack([own["w"]] + [s["w"] for s in others]).max(0).values + 1e-5
    for agg in (BALANCEAggregator(gamma=1.0, alpha=alpha), UBARAggregator(rho=0.5, alpha=alpha)):
        out = agg.aggregate(0, own, nb, 0)["w"]
        assert ((out >= lo) & (out <= hi)).all()                         # convex combination stays inside the hull (per coordinate)
    assert BALANCEAggregator(alpha=1.0).aggregate(0, own, nb, 0)["w"].allclose(own["w"])


@FAST
@given(ws=st.lists(st.floats(0.01, 1.0), min_size=1, max_size=5), dim=st.integers(1, 10))
def test_average_states_weighted(ws, dim):
    w = np.array(ws) / sum(ws)
    states = _states(len(ws), dim, 7)
    out = average_states(states, list(w))
    ref = sum(float(wi) * s["w"] for wi, s in zip(w, states))
    assert torch.allclose(out["w"], ref, atol=1e-5)
    assert compute_model_distance(states[0], states[0]) == 0.0


@FAST
@given(n=st.integers(1, 40), g=st.integers(1, 8), seed=st.integers(0, 99))
def test_placement_is_a_bijection_and_balanced(n, g, seed):
    rng = np.random.RandomState(seed)
    w = rng.randint(1, 20, size=n).tolist()
    for weights in (None, w):
        p = Placement(n, g, weights)
        pairs = {(int(p.rank_of[i]), int(p.slot_of[i])) for i in range(n)}
        assert len(pairs) == n and max(p.counts) - min(p.counts) <= 1 and p.slots_per_rank == max(p.counts)
        assert sorted(sum((p.local_nodes(r) for r in range(g)), [])) == list(range(n))
        assert all(len(p.local_nodes(r)) == p.counts[r] for r in range(g))
    if n >= g > 1:
        bal = Placement(n, g, w); con = Placement(n, g)
        load = lambda pl: max(sum(w[i] for i in pl.local_nodes(r)) for r in range(g))
        assert load(bal) <= load(con) + max(w)                           # LPT is never much worse than contiguous packing


@FAST
@given(hidden=st.lists(st.integers(1, 9), min_size=0, max_size=3), inp=st.integers(1, 9), out=st.integers(1, 5), bn=st.booleans())
def test_state_layout_roundtrip(hidden, inp, out, bn):
    import torch.nn as nn
    layers, prev = [], inp
    for h in hidden:
        layers += [nn.Linear(prev, h)] + ([nn.BatchNorm1d(h)] if bn else []) + [nn.ReLU()]
        prev = h
    model = nn.Sequential(*layers, nn.Linear(prev, out))
    ref = {k: v.clone() for k, v in model.state_dict().items()}
    lay = StateLayout.from_model(model)
    row = torch.zeros(lay.stride); ints = torch.zeros(max(lay.Pi, 1), dtype=torch.long)
    lay.bind(model, row, None, ints if lay.Pi else None)
    views = lay.row_views(row, ints if lay.Pi else None)
    assert list(views) == list(ref) and all(torch.equal(views[k], ref[k]) for k in ref)
    assert lay.stride % 256 == 0 and lay.Pp4 % 4 == 0 and lay.P_float_real == sum(v.numel() for v in ref.values() if v.is_floating_point())
    assert (row[lay.Pf:lay.Pf_pad] == 0).all() and (row[lay.Pp:lay.Pp4] == 0).all()          # padding is zero
    x = torch.randn(4, inp)
    model.eval()
    y0 = model(x)
    row[: lay.Pp] += 1.0                                                     # parameters really alias the arena row
    assert not torch.allclose(model(x), y0)
