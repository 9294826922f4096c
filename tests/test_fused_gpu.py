"""GPU checks of the fused training program: one SGD step of every supported family vs stock autograd (exact fp32),
CUDA-graph replay of an unrolled round, dropout statistics, side-stream weight gradients."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from murmura_b200.models import CIFARCNN, MLP, EvidentialMLP, LEAFFEMNISTModel, ResNet18
from murmura_b200.models.mlp import evidential_loss_reference
from murmura_b200.parallel.arena import StateLayout
from murmura_b200.parallel.fused_trainer import FusedTrainer

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _classes(m):
    return [x for x in m.modules() if isinstance(x, nn.Linear)][-1].out_features


def _setup(factory, shape, ns, batch, steps, evidential=False, side=False, seed=0):
    torch.manual_seed(seed)
    probe = factory()
    layout = StateLayout.from_model(probe, channels_last=True)
    S = len(ns)
    live = torch.zeros(S, layout.stride, device=DEV)
    ints = torch.zeros(S, max(layout.Pi, 1), dtype=torch.int64, device=DEV)
    models, refs, shards = [], [], []
    for s in range(S):
        m = factory().to(DEV)
        layout.bind(m, live[s], None, ints[s] if layout.Pi else None)
        models.append(m); refs.append(copy.deepcopy(m))
        x = torch.randn(ns[s], *shape, device=DEV)
        y = torch.randint(0, _classes(probe), (ns[s],), device=DEV)
        shards.append((x.permute(0, 2, 3, 1).contiguous() if len(shape) == 3 else x.contiguous(), y))
    tr = FusedTrainer(models[0], layout, live, ints if layout.Pi else None, shards, steps, batch, shape, evidential=evidential, seed=3,
                      side_stream=side)
    return tr, layout, live, ints, refs, shards


def _ref_steps(ref, shard, idxs, lr, image, evidential=False, lam=0.0):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    try:
        opt = torch.optim.SGD(ref.parameters(), lr=lr)
        ref.train()
        for idx in idxs:
            xb = shard[0][idx]
            if image:
                xb = xb.permute(0, 3, 1, 2)
            out = ref(xb)
            loss = evidential_loss_reference(out, shard[1][idx], lam) if evidential else F.cross_entropy(out, shard[1][idx])
            opt.zero_grad(); loss.backward(); opt.step()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _update_agreement(layout, live, ints, before, ref, slot):
    """cosine / norm ratio of the parameter update (fused vs autograd) + worst running-stat mismatch."""
    views = layout.row_views(live[slot], ints[slot] if layout.Pi else None)
    old = layout.row_views(before[slot], None)
    du, dr, stat = [], [], 0.0
    for k, v in ref.state_dict().items():
        if not v.is_floating_point():
            assert int(views[k]) == int(v), k
            continue
        if "running_" in k:
            stat = max(stat, float((views[k] - v).abs().max() / v.abs().max().clamp_min(1e-3)))
            continue
        du.append((views[k] - old[k]).reshape(-1)); dr.append((v - old[k]).reshape(-1))
    du, dr = torch.cat(du), torch.cat(dr)
    cos = float(F.cosine_similarity(du, dr, dim=0))
    return cos, float(du.norm() / dr.norm()), stat


FAMILIES = [
    ("resnet18", lambda: ResNet18(10), (3, 32, 32), 16, False),
    ("femnist-cnn", lambda: LEAFFEMNISTModel(62), (1, 28, 28), 16, False),
    ("cifar-cnn", lambda: CIFARCNN(10), (3, 32, 32), 16, False),
    ("mlp", lambda: MLP(784, (200,), 10), (784,), 32, False),
    ("har-mlp", lambda: EvidentialMLP(561, (256, 128), 6, dropout=0.0), (561,), 32, True),
]


@pytest.mark.parametrize("name,factory,shape,batch,evidential", FAMILIES, ids=[f[0] for f in FAMILIES])
def test_fused_step_matches_autograd(name, factory, shape, batch, evidential):
    lr, lam = 0.02, 0.2
    ns = [3 * batch + 5, batch, 2 * batch]
    # ResNet-18 at 32×32 ends on 1×1 maps: BatchNorm over `batch` values amplifies TF32 round-off chaotically over several
    # steps (stock fp32 vs fp64 autograd already disagree), so it is compared after ONE step with a larger batch
    chaotic = name == "resnet18"
    if chaotic:
        batch = 64
        ns = [batch + 9, batch, batch + 3]
    steps = [1, 0, 1] if chaotic else [2, 0, 1]
    tr, layout, live, ints, refs, shards = _setup(factory, shape, ns, batch, steps, evidential)
    assert tr.supported
    tr.lam_t.fill_(lam)
    tr.perm[0, :steps[0] * batch] = torch.randperm(ns[0], device=DEV)[:steps[0] * batch]
    tr.perm[2, :batch] = torch.randperm(ns[2], device=DEV)[:batch]
    before = live.clone()
    tr.run_steps(lr)
    torch.cuda.synchronize()
    image = len(shape) == 3
    _ref_steps(refs[0], shards[0], [tr.perm[0, i * batch:(i + 1) * batch] for i in range(steps[0])], lr, image, evidential, lam)
    _ref_steps(refs[2], shards[2], [tr.perm[2, :batch]], lr, image, evidential, lam)
    for slot in (0, 2):
        cos, ratio, stat = _update_agreement(layout, live, ints, before, refs[slot], slot)
        lo = 0.97 if chaotic else 0.995
        assert cos > lo and 0.95 < ratio < 1.05 and stat < 3e-2, (name, slot, cos, ratio, stat)
    assert torch.equal(live[1], before[1])
    assert torch.isfinite(live).all()
    assert tr.be.tma_launches > 0                                 # the TMA-fed kernels ran (not only the cp.async fallback)


@pytest.mark.parametrize("side", [False, True])
def test_graph_replay_equals_eager(side):
    batch = 16
    fac = lambda: LEAFFEMNISTModel(62)
    a = _setup(fac, (1, 28, 28), [40, 20], batch, [2, 1], side=side, seed=5)
    b = _setup(fac, (1, 28, 28), [40, 20], batch, [2, 1], side=False, seed=5)
    tra, trb = a[0], b[0]
    gen = torch.Generator(device=DEV).manual_seed(11)
    tra.refresh_permutations(1, gen)
    trb.perm.copy_(tra.perm)
    before = a[2].clone()
    tra._capture(0.02); tra.graph.replay()
    trb.run_steps(0.02)
    torch.cuda.synchronize()
    # atomics make the reduction order free: agreement to fp32 round-off, not bit equality
    da, db = a[2] - before, b[2] - before
    assert float((da - db).abs().max()) <= 1e-3 * float(db.abs().max()) + 1e-7
    assert float(db.abs().max()) > 0


def test_dropout_is_unbiased_and_reproduced_in_backward():
    tr, layout, live, ints, refs, shards = _setup(lambda: EvidentialMLP(64, (128,), 4, dropout=0.3), (64,), [64], 32, [1], evidential=True)
    assert tr.supported
    tr.perm[0, :32] = torch.arange(32, device=DEV)
    tr.step(1, 0, 0.0)                                           # lr = 0: a forward/backward pass that leaves the weights alone
    torch.cuda.synchronize()
    bn = tr.bn_ops[0]
    y = bn.y.t[0].view(32, 128)
    kept = (y != 0).float().mean().item()
    assert 0.2 < kept < 0.5                                      # relu halves, dropout keeps 70 %
    dx = bn.x.grad.t[0].view(32, 128)
    assert torch.isfinite(dx).all()
