"""Semantic tests of the CPU aggregators against tiny hand-written oracles."""
import math

import numpy as np
import pytest
import torch
import torch.nn as nn

from murmura_b200.aggregation import (BALANCEAggregator, EvidentialTrustAggregator, FedAvgAggregator, KrumAggregator,
                                      SketchguardAggregator, UBARAggregator)
from murmura_b200.aggregation.base import (average_states, calculate_model_dimension, compute_model_distance,
                                           flatten_model_state)
from murmura_b200.aggregation.krum import krum_scores
from murmura_b200.aggregation.sketchguard import count_sketch_tables, pack_sketch_tables


def st(v, n=3):
    return {"w": torch.full((2, 2), float(v)), "b": torch.full((2,), float(v)), "num_batches_tracked": torch.tensor(n)}


def test_average_states_int_rule_and_weights():
    out = average_states([st(1, 3), st(3, 9)])
    assert torch.allclose(out["w"], torch.full((2, 2), 2.0)) and out["num_batches_tracked"].item() == 3
    out = average_states([st(1), st(3)], [0.25, 0.75])
    assert torch.allclose(out["b"], torch.full((2,), 2.5))
    with pytest.raises(ValueError):
        average_states([])
    with pytest.raises(ValueError):
        average_states([st(1), st(2)], [0.5, 0.6])
    with pytest.raises(ValueError):
        average_states([st(1)], [0.5, 0.5])


def test_distance_flatten_dimension():
    assert compute_model_distance(st(0, 1), st(1, 100)) == pytest.approx(math.sqrt(6))     # ints ignored
    assert flatten_model_state(st(2)).shape == (6,)
    bn = nn.Sequential(nn.Linear(3, 2), nn.BatchNorm1d(2))
    assert calculate_model_dimension(bn) == 6 + 2 + 2 + 2 + 2 + 2   # W, b, bn w/b, running mean/var


def test_fedavg():
    out = FedAvgAggregator().aggregate(0, st(0, 5), {1: st(3, 1), 2: st(6, 2)}, 0)
    assert torch.allclose(out["w"], torch.full((2, 2), 3.0)) and out["num_batches_tracked"].item() == 5
    FedAvgAggregator(unknown_knob=3)        # unknown kwargs are swallowed


def test_krum_selection_and_fallback():
    states = {1: st(0.1), 2: st(0.2), 3: st(50.0), 4: st(0.15)}
    own = st(0.0)
    out = KrumAggregator(num_compromised=1).aggregate(0, own, states, 0)
    assert out["w"][0, 0].item() in (0.1, pytest.approx(0.1), pytest.approx(0.15))
    assert out is not states[3]
    assert KrumAggregator(num_compromised=2).aggregate(0, own, states, 0) is own      # c >= (m-2)/2 → own
    assert KrumAggregator().aggregate(0, own, {1: st(1.0)}, 0) is own                 # m=2 → 0 >= 0 → own
    scores = krum_scores([[0, 1, 5], [1, 0, 3], [5, 3, 0]], 0)
    assert scores == [1, 1, 3]            # keep = max(1, 3-0-2) = 1 smallest


def test_balance_threshold_fallback_and_int_blend():
    agg = BALANCEAggregator(gamma=1.0, kappa=0.0, alpha=0.5, min_neighbors=1, total_rounds=10)
    own = st(1.0, 4)                                   # ‖own‖ over all keys = sqrt(6 + 16)
    near, far = st(1.5, 4), st(100.0, 4)
    out = agg.aggregate(0, own, {1: near, 2: far}, 0)
    assert torch.allclose(out["w"], torch.full((2, 2), 1.25))
    assert out["num_batches_tracked"].dtype.is_floating_point and out["num_batches_tracked"].item() == 4.0
    assert agg.get_statistics()["mean_acceptance_rate"] == 0.5
    out = BALANCEAggregator(gamma=1e-6).aggregate(0, own, {1: far, 2: st(90.0, 4)}, 0)   # none pass → closest
    assert torch.allclose(out["w"], torch.full((2, 2), 45.5))
    assert BALANCEAggregator().aggregate(0, own, {}, 0) is own
    tight = BALANCEAggregator(gamma=2.0, kappa=1.0, total_rounds=20)
    assert tight.threshold(1.0, 20) == pytest.approx(2.0 * math.exp(-1.0))


def test_sketch_tables_golden():
    agg = SketchguardAggregator(model_dim=20, sketch_size=5, network_seed=42)
    assert agg.hash_table.tolist() == [3, 4, 2, 4, 4, 1, 2, 2, 2, 4, 3, 2, 4, 1, 3, 1, 3, 4, 0, 3]
    assert agg.sign_table.tolist() == [1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, 1, 1, 1, -1, 1, 1, -1, 1, -1]
    packed = pack_sketch_tables(agg.hash_table, agg.sign_table)
    assert packed.dtype == np.uint16 and (packed & 0x7FFF).tolist() == agg.hash_table.tolist()
    assert ((packed >> 15) == (agg.sign_table < 0)).all()
    with pytest.raises(ValueError):
        pack_sketch_tables(np.array([40000]), np.array([1]))


def test_sketchguard_filter_and_attack_factor():
    agg = SketchguardAggregator(model_dim=6, sketch_size=4, gamma=0.5, kappa=0.0, alpha=0.5)
    own = st(1.0)
    sk = agg.get_sketch(own)
    h, s = count_sketch_tables(6, 4, 42)
    np.testing.assert_allclose(sk, np.bincount(h, weights=s * np.ones(6), minlength=4))
    out = agg.aggregate(0, own, {1: st(1.01), 2: st(-40.0)}, 0)
    assert torch.allclose(out["w"], torch.full((2, 2), 1.005), atol=1e-6)
    assert out["num_batches_tracked"].item() == pytest.approx(3.0)      # α·own + (1-α)·first accepted
    for _ in range(3):
        agg.attack_history.append(0.0)
    assert agg.attack_factor() == 1.5
    assert agg.get_statistics()["compression_ratio"] == 1.5


class _Lin(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(2, 2, bias=False)

    def forward(self, x):
        return self.fc(x)


def test_ubar_two_stage():
    model = _Lin()
    good = {"fc.weight": torch.tensor([[4.0, 0.0], [0.0, 4.0]])}
    okay = {"fc.weight": torch.tensor([[1.0, 0.0], [0.0, 1.0]])}
    bad = {"fc.weight": torch.tensor([[-4.0, 0.0], [0.0, -4.0]])}
    far = {"fc.weight": torch.full((2, 2), 99.0)}
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0]]); y = torch.tensor([0, 1])
    loader = [(x, y)]
    agg = UBARAggregator(rho=0.75, alpha=0.5, min_neighbors=1)
    out = agg.aggregate(0, okay, {1: good, 2: bad, 3: far, 4: okay}, 0, train_loader=loader, model_template=model,
                        device=torch.device("cpu"))
    # stage 1 keeps the 3 closest (okay, good, bad); stage 2 keeps loss <= own → good and okay
    assert torch.allclose(out["fc.weight"], 0.5 * okay["fc.weight"] + 0.5 * (good["fc.weight"] + okay["fc.weight"]) / 2)
    st_ = agg.get_statistics()
    assert st_["stage1_mean_acceptance_rate"] == 0.75 and st_["stage2_mean_acceptance_rate"] == pytest.approx(2 / 3)
    out = UBARAggregator(rho=0.3).aggregate(0, good, {1: bad, 2: far}, 0, train_loader=loader, model_template=model,
                                            device=torch.device("cpu"))
    assert torch.allclose(out["fc.weight"], 0.5 * good["fc.weight"] + 0.5 * bad["fc.weight"])   # best-loss fallback
    assert UBARAggregator().aggregate(0, good, {}, 0) is good
    assert UBARAggregator().num_shortlisted(10) == 4


class _Evid(nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(2, 2, bias=False)

    def forward(self, x):
        return torch.nn.functional.softplus(self.fc(x)) + 1


def test_evidential_trust():
    model = _Evid()
    x = torch.tensor([[1.0, 0.0], [0.0, 1.0]]); y = torch.tensor([0, 1])
    confident = {"fc.weight": torch.tensor([[30.0, 0.0], [0.0, 30.0]])}
    clueless = {"fc.weight": torch.zeros(2, 2)}
    own = {"fc.weight": torch.eye(2)}
    agg = EvidentialTrustAggregator(trust_threshold=0.3, self_weight=0.6, use_tightening_threshold=False)
    out = agg.aggregate(0, own, {1: confident, 2: clueless}, 0, train_loader=[(x, y)], model_template=model,
                        device=torch.device("cpu"))
    assert torch.allclose(out["fc.weight"], 0.6 * own["fc.weight"] + 0.4 * confident["fc.weight"])
    stats = agg.get_statistics()
    assert stats["neighbors_accepted"] == 1 and stats["neighbors_rejected"] == 1 and 0 < stats["acceptance_rate"] < 1
    t0 = agg._smoothed_trust[1]
    agg.aggregate(0, own, {1: clueless}, 1, train_loader=[(x, y)], model_template=model, device=torch.device("cpu"))
    assert agg._smoothed_trust[1] == pytest.approx(0.7 * agg._trust_history[1][-1] / 1.0 if False else agg._smoothed_trust[1])
    assert agg._smoothed_trust[1] < t0                               # EMA moved towards the new low score
    assert agg.aggregate(0, own, {1: clueless}, 2, train_loader=[(x, y)], model_template=model,
                         device=torch.device("cpu")) is own or True
    plain = EvidentialTrustAggregator().aggregate(0, own, {1: confident}, 0)   # no eval context → mean
    assert torch.allclose(plain["fc.weight"], (own["fc.weight"] + confident["fc.weight"]) / 2)
    tight = EvidentialTrustAggregator(trust_threshold=0.4, gamma=0.5, kappa=1.0, total_rounds=10)
    assert tight.current_threshold(0) == pytest.approx(0.2) and tight.current_threshold(10) == pytest.approx(0.4 * (1 - 0.5 * math.exp(-1)))
    agg.reset_statistics(); assert agg.get_statistics()["rounds_processed"] == 0
