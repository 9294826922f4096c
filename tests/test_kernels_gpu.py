"""Numerics of every sm_100a kernel against plain-PyTorch fp32/fp64 oracles (run with -m gpu on a B200)."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from murmura_b200.ops import reference as R


@pytest.fixture(scope="module")
def ext():
    from murmura_b200 import ops
    assert ops.available(), f"CUDA extension not loaded: {ops.load_error()}"
    return ops.ext()


DEV = "cuda"


class FakeArena:
    """Single-GPU stand-in for the symmetric region: live [S,stride], pub [2,S,stride], 1-entry pointer table."""

    def __init__(self, S, Pf, Pi=0, seed=0):
        g = torch.Generator(device="cpu").manual_seed(seed)
        self.S, self.Pf, self.Pi = S, Pf, Pi
        self.Pf_pad = (max(Pf, 4) + 255) // 256 * 256
        self.stride = self.Pf_pad + ((Pi + 255) // 256 * 256 if Pi else 0)
        self.live = torch.zeros(S, self.stride, device=DEV)
        self.pub = torch.zeros(2, S, self.stride, device=DEV)
        self.live[:, :Pf] = torch.randn(S, Pf, generator=g).to(DEV)
        self.pub[:, :, :Pf] = torch.randn(2, S, Pf, generator=g).to(DEV)
        self.tbl = torch.tensor([self.pub.data_ptr()], dtype=torch.int64, device=DEV)

    def parity_off(self, p):
        return p * self.S * self.stride

    def edges(self, neighbors):
        row_ptr, slot = [0], []
        for v, nb in enumerate(neighbors):
            slot += [v] + list(nb); row_ptr.append(len(slot))
        E = len(slot)
        t = lambda x, dt: torch.tensor(x, dtype=dt, device=DEV)
        return {"row_ptr": t(row_ptr, torch.int32), "src_rank": torch.zeros(E, dtype=torch.int32, device=DEV),
                "src_slot": t(slot, torch.int32), "mask": torch.ones(E, device=DEV), "w": torch.zeros(E, device=DEV),
                "w_tail": torch.zeros(E, device=DEV), "stats": torch.zeros(len(neighbors), 4, device=DEV), "E": E,
                "host_rows": row_ptr, "host_slot": slot}


def _et(et, V):
    return (V, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], et["w"], et["w_tail"], et["stats"])


@pytest.mark.parametrize("Pf", [1000, 4099, 70001])
def test_publish_copy_scale_noise_tail(ext, Pf):
    A = FakeArena(4, Pf, Pi=3)
    ints = torch.tensor([[1, 2, 3], [4, 5, 6], [7, 8, 9], [10, 11, 12]], dtype=torch.int64, device=DEV)
    scale = torch.tensor([1.0, -5.0, 1.0, 1.0], device=DEV); noise = torch.tensor([0.0, 0.0, 10.0, 0.0], device=DEV)
    gid = torch.arange(4, dtype=torch.int32, device=DEV); ticket = torch.zeros(1, dtype=torch.int32, device=DEV)
    A.pub.zero_()
    live0 = A.live.clone()
    ext.publish(A.live, A.pub[1].data_ptr(), A.stride, 4, Pf, A.Pf_pad, ints, scale, noise, gid, 42, 3, 0, 1, 0, 1, ticket)
    torch.cuda.synchronize()
    p = A.pub[1]
    assert torch.equal(p[0, :Pf], live0[0, :Pf]) and torch.equal(p[3, :Pf], live0[3, :Pf])
    assert torch.allclose(p[1, :Pf], -5.0 * live0[1, :Pf])
    z = (p[2, :Pf] - live0[2, :Pf]) / 10.0
    assert abs(z.mean().item()) < 5.0 / math.sqrt(Pf) and abs(z.std().item() - 1.0) < 0.1
    assert (p[:, Pf:A.Pf_pad] == 0).all()                                        # padding stays exactly zero
    assert torch.equal(p[:, A.Pf_pad:A.Pf_pad + 3], ints.float()) and torch.equal(A.live[:, A.Pf_pad:A.Pf_pad + 3], ints.float())
    assert ticket.item() == 0 and (A.pub[0] == 0).all()
    p2 = torch.zeros_like(A.pub[0])
    ext.publish(A.live, p2.data_ptr(), A.stride, 4, Pf, A.Pf_pad, ints, scale, noise, gid, 42, 3, 0, 1, 0, 2, ticket)
    assert torch.equal(p2[2], p[2])                                             # Philox noise is a pure function of (seed, round, node, index)
    ext.publish(A.live, p2.data_ptr(), A.stride, 4, Pf, A.Pf_pad, ints, scale, noise, gid, 42, 4, 0, 1, 0, 3, ticket)
    assert not torch.equal(p2[2], p[2])


@pytest.mark.parametrize("tma", [False, True])
@pytest.mark.parametrize("Pf,deg", [(777, 1), (4096, 3), (100003, 7), (50000, 0)])
def test_weighted_gather_matches_oracle(ext, Pf, deg, tma):
    S = 8
    A = FakeArena(S, Pf, seed=1)
    nbrs = [[(v + k + 1) % S for k in range(deg)] for v in range(S)]
    et = A.edges(nbrs)
    g = torch.Generator().manual_seed(5)
    w = torch.rand(et["E"], generator=g).to(DEV); et["w"].copy_(w)
    live0 = A.live.clone()
    ext.weighted_gather(A.live, A.tbl.data_ptr(), A.parity_off(1), A.stride, S, et["row_ptr"], et["src_rank"], et["src_slot"],
                        et["mask"], et["w"], A.Pf_pad, False, 0, 1, 0, 0.0, 0, tma)
    rows, selfw = [], []
    wh = w.cpu().tolist()
    for v in range(S):
        e0 = et["host_rows"][v]
        selfw.append(wh[e0]); rows.append([(et["host_slot"][e], wh[e]) for e in range(e0 + 1, et["host_rows"][v + 1])])
    ref = R.weighted_gather(live0, A.pub[1], rows, selfw, A.Pf_pad)
    assert torch.allclose(A.live, ref, rtol=1e-5, atol=1e-5)


def test_weighted_gather_renorm_mask_and_keep_own(ext):
    S, Pf = 4, 5000
    A = FakeArena(S, Pf, seed=2)
    et = A.edges([[1, 2, 3], [0], [], [0, 1]])
    ext.fedavg_weights(*_et(et, S))
    et["mask"][2] = 0.0                                   # drop edge 0 <- 2 (fault injection)
    live0 = A.live.clone()
    ext.weighted_gather(A.live, A.tbl.data_ptr(), 0, A.stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                        et["w"], A.Pf_pad, True, 0, 1, 0, 0.0, 0)
    pub = A.pub[0]
    assert torch.allclose(A.live[0], (live0[0] + pub[1] + pub[3]) / 3, atol=1e-5)
    assert torch.allclose(A.live[1], (live0[1] + pub[0]) / 2, atol=1e-5)
    assert torch.equal(A.live[2], live0[2])                # isolated node keeps its own state bit-exactly
    assert torch.equal(et["w_tail"].cpu(), torch.tensor([1.0, 0, 0, 0, 1, 0, 1, 1, 0, 0]))


def test_tail_blend_truncates(ext):
    S, Pf = 2, 64
    A = FakeArena(S, Pf, Pi=2, seed=3)
    A.live[:, A.Pf_pad:A.Pf_pad + 2] = torch.tensor([[10.0, 20.0], [0.0, 0.0]], device=DEV)
    A.pub[0, :, A.Pf_pad:A.Pf_pad + 2] = torch.tensor([[10.0, 20.0], [13.0, 27.0]], device=DEV)
    et = A.edges([[1], [0]])
    et["w_tail"].copy_(torch.tensor([0.5, 0.5, 1.0, 0.0], device=DEV))
    ints = torch.zeros(S, 2, dtype=torch.int64, device=DEV)
    ext.tail_blend(A.live, A.tbl.data_ptr(), 0, A.stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                   et["w_tail"], A.Pf_pad, ints, 0)
    assert ints.tolist() == [[11, 23], [0, 0]]             # trunc(0.5*10+0.5*13)=11, trunc(23.5)=23


@pytest.mark.parametrize("Pf", [513, 20000, 300001])
def test_edge_distances(ext, Pf):
    S = 6
    A = FakeArena(S, Pf, Pi=5, seed=4)
    A.live[:, A.Pf_pad:A.Pf_pad + 5] = 3.0; A.pub[:, :, A.Pf_pad:A.Pf_pad + 5] = 7.0
    nbrs = [[1, 2, 3, 4, 5], [0], [0, 1], [], [5], [4, 0]]
    et = A.edges(nbrs)
    d2 = torch.zeros(et["E"], device=DEV); n2 = torch.zeros(S, device=DEV)
    for length in (A.stride, A.Pf_pad):
        ext.edge_distances(A.live, A.tbl.data_ptr(), A.parity_off(1), A.stride, S, et["row_ptr"], et["src_rank"], et["src_slot"],
                           et["mask"], length, d2, n2, 0, 1, 0, 0.0, 0)
        for v in range(S):
            e0 = et["host_rows"][v]
            if nbrs[v]:
                ref = R.edge_sq_distances(A.live, A.pub[1], v, nbrs[v], length)
                assert torch.allclose(d2[e0 + 1:e0 + 1 + len(nbrs[v])].cpu(), ref.cpu(), rtol=2e-4)
            assert d2[e0].item() == 0.0
            assert n2[v].item() == pytest.approx(float((A.live[v, :length].double() ** 2).sum()), rel=2e-4)


@pytest.mark.parametrize("m,Pf", [(2, 1000), (5, 40000), (9, 12345), (32, 9000)])
def test_pairwise_distances(ext, m, Pf):
    S = m
    A = FakeArena(S, Pf, seed=6)
    nbrs = [[j for j in range(S) if j != v] for v in range(S)]
    et = A.edges(nbrs)
    D = torch.zeros(S, 32, 32, device=DEV)
    ext.pairwise_distances(A.live, A.tbl.data_ptr(), 0, A.stride, S, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"],
                           A.Pf_pad, D, m, 0, 1, 0, 0.0, 0)
    for v in (0, S - 1):
        cands = torch.stack([A.live[v, :A.Pf_pad]] + [A.pub[0, j, :A.Pf_pad] for j in nbrs[v]])
        assert torch.allclose(D[v, :m, :m].cpu(), R.pairwise_sq(cands).cpu(), rtol=3e-4, atol=1e-2)


def test_krum_select_matches_cpu_aggregator(ext):
    from murmura_b200.aggregation.krum import krum_scores
    g = torch.Generator().manual_seed(7)
    V, m = 5, 6
    pts = torch.randn(V, m, 16, generator=g)
    pts[:, 3] += 30.0
    D = torch.zeros(V, 32, 32)
    for v in range(V):
        D[v, :m, :m] = torch.cdist(pts[v], pts[v]) ** 2
    row_ptr = torch.arange(0, (V + 1) * m, m, dtype=torch.int32, device=DEV)
    E = V * m
    et = {"row_ptr": row_ptr, "src_rank": torch.zeros(E, dtype=torch.int32, device=DEV), "src_slot": torch.zeros(E, dtype=torch.int32, device=DEV),
          "mask": torch.ones(E, device=DEV), "w": torch.zeros(E, device=DEV), "w_tail": torch.zeros(E, device=DEV),
          "stats": torch.zeros(V, 4, device=DEV)}
    win = torch.zeros(V, dtype=torch.int32, device=DEV)
    for c in (0, 1, 2):
        ext.krum_select(*_et(et, V), D.to(DEV), c, win, 0)
        for v in range(V):
            dist = torch.cdist(pts[v], pts[v]).tolist()
            expect = 0 if c >= (m - 2) / 2 else int(np.argmin(krum_scores(dist, c)))
            assert win[v].item() == expect
            w = et["w"][v * m:(v + 1) * m].cpu()
            assert w.sum().item() == 1.0 and w[expect].item() == 1.0


def test_count_sketch_and_mxfp8(ext):
    from murmura_b200.aggregation.sketchguard import count_sketch_tables, pack_sketch_tables
    Pf, K, S = 54321, 1000, 3
    A = FakeArena(S, Pf, seed=8)
    buckets, signs = count_sketch_tables(Pf, K, 42)
    table = np.zeros((Pf + 3) // 4 * 4, dtype=np.uint16); table[:Pf] = pack_sketch_tables(buckets, signs)
    t = torch.from_numpy(table.view(np.int16)).to(DEV)
    out = torch.zeros(S, K, device=DEV)
    ext.count_sketch(A.live.data_ptr(), A.stride, torch.arange(S, dtype=torch.int32, device=DEV), t, Pf, K, out)
    for v in range(S):
        ref = R.count_sketch(A.live[v, :Pf], buckets, signs, K)
        assert torch.allclose(out[v].cpu(), ref, rtol=1e-3, atol=2e-3)
    Kpad = 1024
    q = torch.zeros(S, Kpad, dtype=torch.uint8, device=DEV); sc = torch.zeros(S, Kpad // 32, dtype=torch.uint8, device=DEV)
    ext.sketch_quant_mxfp8(out, q.data_ptr(), sc.data_ptr(), Kpad)
    deq = q.view(torch.float8_e4m3fn).float().view(S, -1, 32) * torch.exp2(sc.float() - 127).unsqueeze(-1)
    deq = deq.view(S, Kpad)[:, :K]
    assert torch.allclose(deq, R.mxfp8_roundtrip(out), rtol=0, atol=1e-6)
    rel = (deq - out).norm() / out.norm()
    assert rel < 0.05                                      # e4m3 (3 mantissa bits) block-scaled: ~3% rms


def test_filters_match_cpu_aggregators(ext):
    """balance / ubar stage 1+2 / trust kernels reproduce the CPU classes' decisions and weights."""
    from murmura_b200.aggregation import BALANCEAggregator, EvidentialTrustAggregator, UBARAggregator
    V, deg = 4, 5
    m = deg + 1
    g = torch.Generator().manual_seed(11)
    d = torch.rand(V, deg, generator=g) * 4
    own_norm = torch.tensor([3.0, 0.5, 10.0, 1.0])
    E = V * m
    et = {"row_ptr": torch.arange(0, E + 1, m, dtype=torch.int32, device=DEV), "src_rank": torch.zeros(E, dtype=torch.int32, device=DEV),
          "src_slot": torch.zeros(E, dtype=torch.int32, device=DEV), "mask": torch.ones(E, device=DEV), "w": torch.zeros(E, device=DEV),
          "w_tail": torch.zeros(E, device=DEV), "stats": torch.zeros(V, 4, device=DEV)}
    d2 = torch.zeros(V, m); d2[:, 1:] = d ** 2
    d2 = d2.reshape(-1).to(DEV); n2 = (own_norm ** 2).to(DEV); dist = torch.zeros(E, device=DEV)
    agg = BALANCEAggregator(gamma=0.6, kappa=1.0, alpha=0.3, min_neighbors=1, total_rounds=10)
    from murmura_b200.aggregation.balance import decayed_factor
    ext.balance_filter(*_et(et, V), d2, n2, dist, decayed_factor(0.6, 1.0, 4, 10), 0.3, 1, 0)
    w = et["w"].view(V, m).cpu()
    for v in range(V):
        thr = agg.threshold(own_norm[v].item(), 4)
        acc = agg.select({j: d[v, j].item() for j in range(deg)}, thr)
        expect = torch.zeros(m); expect[0] = 0.3
        for j in acc:
            expect[1 + j] = 0.7 / len(acc)
        assert torch.allclose(w[v], expect, atol=1e-6), (v, w[v], expect)
    assert torch.allclose(et["w_tail"].view(V, m).cpu(), w)

    ub = UBARAggregator(rho=0.5, alpha=0.4, min_neighbors=1)
    cand = torch.zeros(E, device=DEV); rank = torch.zeros(E, device=DEV)
    ext.ubar_stage1(*_et(et, V), d2, 0.5, 1, cand, rank, 0)
    loss = torch.rand(V, m, generator=g); own_loss = torch.tensor([0.5, 0.0, 0.9, 0.3])
    ext.ubar_stage2(*_et(et, V), cand, rank, loss.reshape(-1).to(DEV), own_loss.to(DEV), 0.4, True)
    w = et["w"].view(V, m).cpu(); wt = et["w_tail"].view(V, m).cpu()
    for v in range(V):
        short = ub.shortlist({j: d[v, j].item() for j in range(deg)})
        assert sorted(short) == sorted((cand.view(V, m)[v, 1:].nonzero().flatten()).tolist())
        kept = ub.loss_filter(own_loss[v].item(), {j: loss[v, 1 + j].item() for j in short})
        expect = torch.zeros(m); expect[0] = 0.4
        for j in kept:
            expect[1 + j] = 0.6 / len(kept)
        assert torch.allclose(w[v], expect, atol=1e-6)
        assert wt[v, 0].item() == pytest.approx(0.4) and wt[v, 1 + kept[0]].item() == pytest.approx(0.6)

    et_agg = EvidentialTrustAggregator(accuracy_weight=0.7, vacuity_threshold=0.5, trust_threshold=0.2, self_weight=0.6,
                                       use_tightening_threshold=False)
    vac = torch.rand(V, m, generator=g); acc = torch.rand(V, m, generator=g)
    gid = torch.arange(m, dtype=torch.int32).repeat(V).to(DEV)
    ema = torch.zeros(V, m, device=DEV); valid = torch.zeros(V, m, device=DEV); trust = torch.zeros(E, device=DEV)
    for rnd in range(2):
        ext.trust_filter(*_et(et, V), vac.reshape(-1).to(DEV), acc.reshape(-1).to(DEV), gid, m, ema, valid, 0.7, 0.5, 0.7, True, 0.2,
                         0.6, trust, 0)
        w = et["w"].view(V, m).cpu()
        for v in range(1):
            scores = {}
            for j in range(1, m):
                s = et_agg.score_from_metrics({"vacuity": vac[v, j].item(), "accuracy": acc[v, j].item()})
                scores[j] = et_agg.smooth(j, s)
            accd = {j: s for j, s in scores.items() if s >= 0.2}
            expect = torch.zeros(m)
            if accd:
                expect[0] = 0.6
                for j, s in accd.items():
                    expect[j] = 0.4 * s / sum(accd.values())
            else:
                expect[0] = 1.0
            assert torch.allclose(w[v], expect, atol=1e-5)


def test_sketchguard_filter_kernel(ext):
    V, K = 3, 200
    g = torch.Generator().manual_seed(13)
    base = torch.randn(K, generator=g)
    own = base + 0.05 * torch.randn(V, K, generator=g)               # all honest nodes sit near a common model
    pubsk = torch.zeros(2, V, K)
    pubsk[1] = own + 0.01 * torch.randn(V, K, generator=g)
    pubsk[1, 2] = own[2] * -5.0                                       # node 2 publishes a directed-deviation sketch
    own_d = own.to(DEV); pub_d = pubsk.to(DEV)
    et_n = [[1, 2], [0, 2], [0, 1]]
    rows, slots = [0], []
    for v, nb in enumerate(et_n):
        slots += [v] + nb; rows.append(len(slots))
    E = len(slots)
    et = {"row_ptr": torch.tensor(rows, dtype=torch.int32, device=DEV), "src_rank": torch.zeros(E, dtype=torch.int32, device=DEV),
          "src_slot": torch.tensor(slots, dtype=torch.int32, device=DEV), "mask": torch.ones(E, device=DEV), "w": torch.zeros(E, device=DEV),
          "w_tail": torch.zeros(E, device=DEV), "stats": torch.zeros(V, 4, device=DEV)}
    tbl = torch.tensor([pub_d.data_ptr()], dtype=torch.int64, device=DEV)
    hist = torch.zeros(V, 4, device=DEV); dist = torch.zeros(E, device=DEV)
    ext.sketchguard_filter(*_et(et, V), own_d, tbl.data_ptr(), 0, 0, 1 * V, K, 224, False, 0.5, 0.5, 1, hist, dist, 0, 1, 0, 0.0, 0)
    w = et["w"].cpu()
    # node 0: neighbour 1 close (accepted), neighbour 2 = -5x (rejected)
    assert w[0].item() == 0.5 and w[1].item() == 0.5 and w[2].item() == 0.0
    assert dist[1].item() == pytest.approx((own[0] - pubsk[1, 1]).norm().item(), rel=1e-4)
    assert hist[0, 0].item() == 0.5 and hist[0, 3].item() == 1.0
    # fp8 path gives the same decisions
    Kpad = 224
    q = torch.zeros(2, V, Kpad, dtype=torch.uint8, device=DEV); sc = torch.zeros(2, V, Kpad // 32, dtype=torch.uint8, device=DEV)
    ext.sketch_quant_mxfp8(pub_d[1].contiguous(), q[1].data_ptr(), sc[1].data_ptr(), Kpad)
    tq = torch.tensor([q.data_ptr()], dtype=torch.int64, device=DEV); ts = torch.tensor([sc.data_ptr()], dtype=torch.int64, device=DEV)
    w_before = et["w"].clone()
    ext.sketchguard_filter(*_et(et, V), own_d, tbl.data_ptr(), tq.data_ptr(), ts.data_ptr(), 1 * V, K, Kpad, True, 0.5, 0.5, 1, hist,
                           dist, 0, 1, 0, 0.0, 0)
    assert torch.equal(et["w"], w_before)


def test_sgd_and_eval_kernels(ext):
    S, P = 3, 10000
    live = torch.randn(S, 12288, device=DEV); grad = torch.randn(S, P, device=DEV)
    l0, g0 = live.clone(), grad.clone()
    ext.sgd_step(live, 12288, grad, P, 1, 2, P, 0.1)
    assert torch.equal(live[0], l0[0]) and torch.allclose(live[1:, :P], l0[1:, :P] - 0.1 * g0[1:], atol=1e-6)
    assert torch.equal(live[:, P:], l0[:, P:]) and (grad[1:] == 0).all() and torch.equal(grad[0], g0[0])
    g = torch.Generator().manual_seed(3)
    for C in (2, 10, 62, 100):
        logits = torch.randn(257, C, generator=g).to(DEV) * 3; y = torch.randint(0, C, (257,), generator=g).to(DEV)
        stats = torch.zeros(8, device=DEV)
        ext.ce_eval(logits, y, None, stats)
        assert torch.allclose(stats[:3].cpu(), R.ce_stats(logits.cpu(), y.cpu()), rtol=1e-4)
        nv = torch.tensor([100], dtype=torch.int32, device=DEV); stats.zero_()
        ext.ce_eval(logits, y, nv, stats)
        assert torch.allclose(stats[:3].cpu(), R.ce_stats(logits[:100].cpu(), y[:100].cpu()), rtol=1e-4)
        alpha = torch.nn.functional.softplus(logits) + 1; stats.zero_()
        ext.dirichlet_eval(alpha.contiguous(), y, None, stats)
        assert torch.allclose(stats[:6].cpu(), R.dirichlet_stats(alpha.cpu(), y.cpu()).float(), rtol=2e-4)


@pytest.mark.parametrize("C,lam", [(6, 0.0), (6, 0.1), (12, 1.0), (62, 0.05)])
def test_evidential_loss_fused(ext, C, lam):
    from murmura_b200 import ops
    from murmura_b200.models.mlp import evidential_loss_reference
    g = torch.Generator().manual_seed(C)
    raw = (torch.randn(33, C, generator=g) * 2).to(DEV)
    y = torch.randint(0, C, (33,), generator=g).to(DEV)
    a1 = (torch.nn.functional.softplus(raw) + 1).requires_grad_(True)
    a2 = a1.detach().clone().double().requires_grad_(True)
    l1 = ops.evidential_loss(a1, y, lam); l1.backward()
    l2 = evidential_loss_reference(a2, y, lam); l2.backward()
    assert l1.item() == pytest.approx(l2.item(), rel=2e-4, abs=1e-5)
    assert torch.allclose(a1.grad, a2.grad.float(), rtol=2e-3, atol=2e-5)
    lam_t = torch.tensor(lam, device=DEV)
    a3 = a1.detach().clone().requires_grad_(True)
    l3 = ops.evidential_loss(a3, y, lam_t); l3.backward()
    assert torch.allclose(l3, l1) and torch.allclose(a3.grad, a1.grad)


def test_mobility_and_dmtt_kernels(ext):
    from murmura_b200.attacks import TopologyLiarAttack
    from murmura_b200.config.schema import DMTTConfig
    from murmura_b200.dmtt import DMTTNodeState
    from murmura_b200.dmtt.node_process import verify_claim
    from murmura_b200.topology import MobilityModel
    N, R = 32, 6
    mm = MobilityModel(N, 100.0, 30.0, 5.0, seed=42)
    pos = torch.from_numpy(mm.positions_tensor(R)).to(DEV)
    adj = torch.zeros(N, N, dtype=torch.uint8, device=DEV)
    for r in range(R):
        ext.mobility_adjacency(pos, r, 100.0, 30.0, True, adj)
        assert np.array_equal(adj.cpu().numpy().astype(bool), mm.adjacency_at(r))
    sparse = MobilityModel(8, 1000.0, 1.0, 1.0, seed=0)
    adj8 = torch.zeros(8, 8, dtype=torch.uint8, device=DEV)
    ext.mobility_adjacency(torch.from_numpy(sparse.positions_tensor(1)).to(DEV), 0, 1000.0, 1.0, True, adj8)
    assert np.array_equal(adj8.cpu().numpy().astype(bool), sparse.adjacency_at(0))
    liar = TopologyLiarAttack(N, 0.3, seed=42)
    is_liar = torch.zeros(N, dtype=torch.uint8); is_liar[sorted(liar.get_compromised_nodes())] = 1
    claims = torch.zeros(N, N, dtype=torch.uint8, device=DEV)
    ext.liar_claims(adj, is_liar.to(DEV), claims)
    truth = mm.adjacency_at(R - 1)
    assert np.array_equal(claims.cpu().numpy().astype(bool), liar.claim_bitmask(truth))
    # trust update + Top-B vs the scalar reference-parity implementation
    cfg = DMTTConfig(budget_B=3)
    V, node0 = 4, 8
    g = torch.Generator().manual_seed(1)
    collab = torch.from_numpy(truth.astype(np.uint8)).clone()
    received = torch.zeros(V, N, dtype=torch.uint8)
    score = torch.rand(V, N, generator=g); valid = torch.zeros(V, N, dtype=torch.uint8)
    states = [DMTTNodeState(node0 + v, cfg, N) for v in range(V)]
    for v in range(V):
        i = node0 + v
        nb = np.flatnonzero(truth[i])
        got = nb[::2]
        received[v, got] = 1; valid[v, got] = 1
    c_hat = torch.full((V, N), 0.5, device=DEV); al = torch.ones(V, N, device=DEV); be = torch.ones(V, N, device=DEV)
    nxt = torch.zeros(V, N, dtype=torch.uint8, device=DEV); q = torch.zeros(V, N, device=DEV)
    claims_np = claims.cpu().numpy().astype(bool)
    for rnd in range(2):
        ext.dmtt_update(adj, claims, collab.to(DEV), received.to(DEV), score.to(DEV), valid.to(DEV), c_hat, al, be, nxt, q, cfg.rho,
                        cfg.lambda_forget, cfg.w_d, cfg.w_x, cfg.tau_U, cfg.eta, cfg.lambda1, cfg.lambda2, cfg.lambda3, cfg.budget_B,
                        torch.arange(node0, node0 + V, dtype=torch.int32, device=DEV))
        for v in range(V):
            i = node0 + v
            st = states[v]
            for j in np.flatnonzero(collab[i].numpy()):
                st.update_link_reliability(int(j), bool(received[v, j]))
            for j in np.flatnonzero(received[v].numpy()):
                d, x = verify_claim(np.flatnonzero(claims_np[j]).tolist(), set(np.flatnonzero(truth[j]).tolist()))
                st.update_trust(int(j), d=d, x=x)
            ms = {int(j): float(score[v, j]) for j in np.flatnonzero(valid[v].numpy())}
            expect = st.top_b(np.flatnonzero(truth[i]).tolist(), ms, cfg.budget_B)
            assert sorted(np.flatnonzero(nxt[v].cpu().numpy()).tolist()) == sorted(expect)
            for j in expect:
                assert q[v, j].item() == pytest.approx(st.collab_score(j, ms.get(j, 0.5)), rel=1e-4, abs=1e-5)


@pytest.mark.parametrize("rows,P", [(8, 4096), (24, 100000 // 32 * 32), (48, 1 << 20), (20, 11191296)])
def test_gram_tcgen05_tf32(ext, rows, P):
    """tcgen05/TMEM/TMA Gram kernel vs an fp64 X·Xᵀ (TF32 inputs → ~1e-3 relative)."""
    g = torch.Generator().manual_seed(rows)
    stride = P + 256
    X = torch.zeros(3 * rows, stride, device=DEV)
    X[:, :P] = torch.randn(3 * rows, P, generator=g).to(DEV)
    gpr = (rows + 7) // 8
    # tile = 8-row groups of rows [0, rows) ("live" plane) followed by groups of [2*rows, 3*rows) ("published parity 1")
    gy = [8 * j for j in range(gpr)] + [2 * rows + 8 * j for j in range(gpr)]
    gm = [0] * len(gy)
    kbs = ext.gram_kb_per_stage(len(gy))
    maps = ext.gram_make_maps([X.data_ptr()], 3 * rows, stride, P, kbs)
    out = torch.zeros(128 * 128, device=DEV)
    R = 8 * len(gy)
    ext.gram_tf32(maps, gm, gy, 0, P // 32, R, out, True, 0)
    torch.cuda.synchronize()
    G = out.view(128, 128)
    sel_rows = [min(y + i, 3 * rows - 1) for y in gy for i in range(8)]
    valid = torch.tensor([y + i < 3 * rows for y in gy for i in range(8)])
    sel = X[sel_rows, :P].double() * valid.to(DEV).double().unsqueeze(1)         # rows past the tensor end are zero-filled by TMA
    ref = (sel @ sel.T).float()
    got = G[:R, :R]
    err = (got - ref).abs().max().item() / ref.diagonal().max().item()
    assert err < 3e-3, err
    # split-K accumulation across two launches (stage-aligned split) equals one launch
    out2 = torch.zeros(128 * 128, device=DEV)
    half = (P // 32) // 2 // kbs * kbs
    ext.gram_tf32(maps, gm, gy, 0, half, R, out2, True, 0)
    ext.gram_tf32(maps, gm, gy, half, P // 32, R, out2, False, 0)
    scale = ref.diagonal().max().item()
    assert ((out2.view(128, 128)[:R, :R] - got).abs().max().item() / scale) < 1e-3
    # an unaligned tail (kb1 not a multiple of the stage depth) only accumulates the valid k-blocks
    out3 = torch.zeros(128 * 128, device=DEV)
    kb_tail = max(1, P // 32 - 3)
    ext.gram_tf32(maps, gm, gy, 0, kb_tail, R, out3, True, 0)
    ref3 = (sel[:, :kb_tail * 32] @ sel[:, :kb_tail * 32].T).float()
    assert ((out3.view(128, 128)[:R, :R] - ref3).abs().max().item() / scale) < 3e-3
    # distances from the Gram agree with exact fp32 distances to TF32 accuracy
    d_ref = torch.cdist(sel, sel).pow(2).float()
    dg = got.diagonal()[:, None] + got.diagonal()[None, :] - 2 * got
    assert ((dg - d_ref).abs().max() / d_ref.max()).item() < 5e-3

def test_sgd_multi(ext):
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 3, 7, 7), (10,), (512, 512, 3, 3), (5130,), (1,), (3, 5)]
    params = [torch.randn(*s, generator=g).to(DEV) for s in shapes]
    params[0] = params[0].contiguous(memory_format=torch.channels_last)
    grads = [torch.randn_like(p) for p in params]                       # randn_like preserves the channels_last strides
    assert grads[0].stride() == params[0].stride()
    ref = [p - 0.05 * gr for p, gr in zip(params, grads)]
    g0 = [gr.clone() for gr in grads]
    ext.sgd_multi(params, grads, 0.05)
    for p, r, gr, gr0 in zip(params, ref, grads, g0):
        assert torch.allclose(p, r, atol=1e-6) and torch.equal(gr, gr0)
    big = torch.zeros(22_000_000, device=DEV); gb = torch.ones_like(big)   # > 320 chunks → split across launches
    ext.sgd_multi([big], [gb], 2.0)
    assert (big == -2.0).all()


@pytest.mark.parametrize("shape,cl", [((37, 64, 8, 8), True), ((37, 64, 8, 8), False), ((5, 512, 1, 1), True), ((9, 30, 7, 5), True),
                                      ((33, 16), False)])
def test_bn_eval_kernel_and_patch(ext, shape, cl):
    import torch.nn as nn
    from murmura_b200.ops import fast_eval_batchnorm
    g = torch.Generator().manual_seed(len(shape))
    C = shape[1]
    x = torch.randn(*shape, generator=g).to(DEV)
    if cl and len(shape) == 4:
        x = x.contiguous(memory_format=torch.channels_last)
    bn = (nn.BatchNorm2d(C) if len(shape) == 4 else nn.BatchNorm1d(C)).to(DEV).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(C, generator=g)); bn.running_var.copy_(torch.rand(C, generator=g) + 0.3)
        bn.weight.copy_(torch.randn(C, generator=g)); bn.bias.copy_(torch.randn(C, generator=g))
        ref = bn(x)
        with fast_eval_batchnorm():
            got = bn(x)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)
        assert torch.allclose(ext.bn_eval(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, 1e-5, True), ref.relu(), rtol=1e-4, atol=1e-5)
    bn.train()
    with fast_eval_batchnorm():
        assert torch.allclose(bn(x), torch.nn.functional.batch_norm(x, None, None, bn.weight, bn.bias, True), atol=1e-4)   # training path untouched


@pytest.mark.parametrize("K,N,act,bn", [(561, 256, 1, True), (256, 128, 1, True), (128, 6, 2, False), (4000, 512, 1, True), (784, 200, 1, False),
                                        (200, 10, 0, False)])
def test_grouped_linear_tcgen05(ext, K, N, act, bn):
    """Grouped tf32 linear + fused bias/BN/activation epilogue vs fp64 torch, ragged M per group, unaligned rows (K=561)."""
    g = torch.Generator().manual_seed(K + N)
    Ms = [100, 128, 37, 200]
    G = len(Ms)
    maxm = max(Ms)
    X = [torch.randn(m, K, generator=g).to(DEV) for m in Ms]
    blob = torch.randn(G, N * K + 5 * N + 16, generator=g).to(DEV)          # weights at odd offsets like an arena row
    Y = torch.zeros(G, maxm, N, device=DEV)
    host = np.zeros((G, 9), dtype=np.int64)
    refs = []
    for i in range(G):
        off = 1 + i                                                        # deliberately misaligned
        W = blob[i, off:off + N * K].view(N, K); b = blob[i, off + N * K: off + N * K + N]
        mean = blob[i, off + N * K + N: off + N * K + 2 * N]
        var = blob[i, off + N * K + 2 * N: off + N * K + 3 * N]
        var.copy_(var.abs() + 0.5)                                         # in place: the kernel reads this very memory
        gam = blob[i, off + N * K + 3 * N: off + N * K + 4 * N]; bet = blob[i, off + N * K + 4 * N: off + N * K + 5 * N]
        host[i] = (X[i].data_ptr(), W.data_ptr(), b.data_ptr(), mean.data_ptr() if bn else 0, var.data_ptr() if bn else 0,
                   gam.data_ptr() if bn else 0, bet.data_ptr() if bn else 0, Y[i].data_ptr(), Ms[i])
        z = X[i].double() @ W.double().T + b.double()
        if bn:
            z = (z - mean.double()) / torch.sqrt(var.double() + 1e-5) * gam.double() + bet.double()
        z = z.relu() if act == 1 else (torch.nn.functional.softplus(z) + 1 if act == 2 else z)
        refs.append(z.float())
    ext.grouped_linear_tf32(torch.from_numpy(host).to(DEV), G, maxm, K, N, K, N, act, 1e-5)
    torch.cuda.synchronize()
    for i in range(G):
        got = Y[i, :Ms[i]]
        scale = refs[i].abs().max().item() + 1e-6
        assert ((got - refs[i]).abs().max().item() / scale) < 4e-3, (i, (got - refs[i]).abs().max().item(), scale)
        assert (Y[i, Ms[i]:] == 0).all()                                    # rows past M are never written
    # grouped metrics on the outputs
    if act == 2 or act == 0:
        C = N
        tg = [torch.randint(0, C, (m,), generator=g).to(DEV) for m in Ms]
        ev = np.array([[Y[i].data_ptr(), tg[i].data_ptr(), Ms[i]] for i in range(G)], dtype=np.int64)
        stats = torch.zeros(G, 8, device=DEV)
        ext.grouped_eval(torch.from_numpy(ev).to(DEV), G, maxm, C, N, act == 2, stats)
        for i in range(G):
            ref = R.dirichlet_stats(Y[i, :Ms[i]].cpu(), tg[i].cpu()) if act == 2 else R.ce_stats(Y[i, :Ms[i]].cpu(), tg[i].cpu())
            assert torch.allclose(stats[i, :len(ref)].cpu(), ref.float(), rtol=3e-4, atol=1e-3)


# ---- fused BatchNorm (+residual) (+ReLU) training kernels (bn_train.cu) -----------------------------------------------
@pytest.mark.parametrize("shape", [(64, 64, 16, 16), (64, 64, 8, 8), (64, 128, 4, 4), (64, 512, 1, 1), (7, 20, 5, 3), (32, 256), (3, 36)])
@pytest.mark.parametrize("relu,with_res", [(True, False), (True, True), (False, False)])
def test_bn_act_training_matches_stock_batchnorm(ext, shape, relu, with_res):
    import torch.nn as nn
    import torch.nn.functional as F
    from murmura_b200 import ops
    torch.manual_seed(5)
    C = shape[1]
    bn_cls = nn.BatchNorm2d if len(shape) == 4 else nn.BatchNorm1d
    fmt = torch.channels_last if len(shape) == 4 else torch.contiguous_format
    ref, fused = bn_cls(C).cuda(), bn_cls(C).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.normal_(); ref.running_mean.normal_(); ref.running_var.uniform_(0.5, 2.0)
    fused.load_state_dict(ref.state_dict())
    x0 = (torch.randn(*shape, device="cuda") * 2.0 + 3.0).contiguous(memory_format=fmt)        # non-zero mean: exercises the shifted sums
    r0 = torch.randn(*shape, device="cuda").contiguous(memory_format=fmt) if with_res else None
    g = torch.randn(*shape, device="cuda").contiguous(memory_format=fmt)
    outs = []
    for mod, use_fused in ((ref, False), (fused, True)):
        x = x0.clone().requires_grad_(True)
        r = r0.clone().requires_grad_(True) if with_res else None
        if use_fused:
            assert ops.bn_act_fusable(x, mod)
            y = ops.bn_act(x, mod, residual=r, relu=relu)
        else:
            y = mod(x)
            if with_res:
                y = y + r
            y = F.relu(y) if relu else y
        y.backward(g)
        outs.append((y.detach(), x.grad, mod.weight.grad, mod.bias.grad, r.grad if with_res else None,
                     mod.running_mean.clone(), mod.running_var.clone(), int(mod.num_batches_tracked)))
    a, b = outs
    names = ["y", "dx", "dgamma", "dbeta", "dres", "running_mean", "running_var"]
    for name, u, v in zip(names, a[:7], b[:7]):
        if u is None:
            continue
        scale = float(u.abs().max()) + 1e-6
        assert float((u - v).abs().max()) / scale < 2e-4, name
    assert a[7] == b[7] == 1
    assert b[0].stride() == x0.stride()


def test_bn_act_eval_and_fallbacks(ext):
    import torch.nn as nn
    import torch.nn.functional as F
    from murmura_b200 import ops
    torch.manual_seed(6)
    bn = nn.BatchNorm2d(32).cuda()
    with torch.no_grad():
        bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_()
    bn.eval()
    x = torch.randn(9, 32, 6, 6, device="cuda").contiguous(memory_format=torch.channels_last)
    r = torch.randn_like(x)
    with torch.no_grad():
        got = ops.bn_act(x, bn, residual=r, relu=True)
        want = F.relu(bn(x) + r)
    assert float((got - want).abs().max()) < 1e-5
    # NCHW-contiguous input is not fusable: identical semantics through the stock ops
    xn = torch.randn(9, 32, 6, 6, device="cuda")
    assert not ops.bn_act_fusable(xn, bn)
    with torch.no_grad():
        assert torch.equal(ops.bn_act(xn, bn, relu=False), bn(xn))
    assert int(bn.num_batches_tracked) == 0


@pytest.mark.parametrize("B,C", [(64, 10), (7, 62), (130, 6), (1, 3)])
def test_ce_loss_fwd_bwd_matches_autograd(ext, B, C):
    import torch.nn.functional as F
    torch.manual_seed(B + C)
    z = (torch.randn(B, C, device="cuda") * 3).requires_grad_(True)
    t = torch.randint(0, C, (B,), device="cuda")
    want = F.cross_entropy(z, t); want.backward()
    acc = torch.full((), 2.5, device="cuda")
    loss, grad = ext.ce_loss_fwd_bwd(z.detach(), t, acc)
    assert abs(float(loss) - float(want)) < 1e-5 * max(1.0, abs(float(want)))
    assert float((grad - z.grad).abs().max()) < 1e-6
    assert abs(float(acc) - 2.5 - float(want)) < 1e-4
    from murmura_b200 import ops
    z2 = z.detach().clone().requires_grad_(True)
    (ops.ce_loss(z2, t) * 3.0).backward()
    assert float((z2.grad - 3.0 * z.grad).abs().max()) < 1e-5


@pytest.mark.parametrize("shape", [(50, 32, 32, 3), (33, 561), (29, 7)])
def test_gather_batch_walks_permutation_and_advances_step(ext, shape):
    torch.manual_seed(1)
    n, eb = shape[0], 8
    X = torch.randn(*shape, device="cuda"); y = torch.randint(0, 9, (n,), device="cuda")
    perm = torch.randperm(n, device="cuda")[: 3 * eb].contiguous()
    step = torch.zeros((), dtype=torch.int64, device="cuda"); ticket = torch.zeros((), dtype=torch.int32, device="cuda")
    for k in range(3):
        xb, yb = ext.gather_batch(X, y, perm, step, ticket, eb)
        idx = perm[k * eb:(k + 1) * eb]
        assert torch.equal(xb, X.index_select(0, idx)) and torch.equal(yb, y.index_select(0, idx))
        assert int(step) == k + 1 and int(ticket) == 0
