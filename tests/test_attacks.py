import random

import torch

from murmura_b200.attacks import Attack, DirectedDeviationAttack, GaussianAttack, TopologyLiarAttack


def _state():
    return {"w": torch.ones(4, 3), "b": torch.zeros(3), "num_batches_tracked": torch.tensor(7)}


def test_compromised_sets_golden(capsys):
    assert sorted(GaussianAttack(10, 0.1).get_compromised_nodes()) == [1]
    assert sorted(GaussianAttack(10, 0.2).get_compromised_nodes()) == [0, 1]
    assert sorted(GaussianAttack(10, 0.3).get_compromised_nodes()) == [0, 1, 4]
    assert sorted(GaussianAttack(8, 0.3).get_compromised_nodes()) == [0, 1]
    assert sorted(DirectedDeviationAttack(16, 0.3).get_compromised_nodes()) == [0, 3, 4, 11]
    assert sorted(GaussianAttack(20, 0.3).get_compromised_nodes()) == [0, 2, 3, 7, 8, 16]
    assert sorted(TopologyLiarAttack(32, 0.3).get_compromised_nodes()) == [0, 3, 4, 7, 8, 23, 27, 29, 31]
    out = capsys.readouterr().out
    assert "Gaussian Attack: Compromised 1/10 nodes" in out and "Noise std: 10.0" in out


def test_gaussian_reseeds_global_rng_liar_does_not():
    random.seed(123); before = random.random()
    random.seed(123); TopologyLiarAttack(10, 0.3, seed=42); assert random.random() == before
    GaussianAttack(10, 0.3, seed=42)
    random.seed(42); expect = random.sample(range(10), 3)
    assert sorted(expect) == [0, 1, 4]


def test_gaussian_apply():
    atk = GaussianAttack(4, 0.5, noise_std=2.0, seed=1)
    bad = next(iter(atk.get_compromised_nodes())); good = next(i for i in range(4) if not atk.is_compromised(i))
    st = _state()
    assert atk.apply_attack(good, st, 0) is st
    out = atk.apply_attack(bad, st, 0)
    assert out["num_batches_tracked"].item() == 7 and out["num_batches_tracked"] is not st["num_batches_tracked"]
    assert not torch.equal(out["w"], st["w"]) and out["w"].dtype == torch.float32
    assert isinstance(atk, Attack)
    assert atk.device_spec()["noise_std"] == 2.0


def test_directed_apply():
    atk = DirectedDeviationAttack(4, 0.5, lambda_param=-5.0, seed=1)
    bad = next(iter(atk.get_compromised_nodes()))
    out = atk.apply_attack(bad, _state(), 3)
    assert torch.equal(out["w"], torch.full((4, 3), -5.0)) and out["num_batches_tracked"].item() == 7
    assert atk.device_spec()["scale"] == -5.0


def test_topology_liar():
    inner = DirectedDeviationAttack(10, 0.3, lambda_param=2.0, seed=42)
    liar = TopologyLiarAttack(10, 0.3, seed=42, model_attack=inner)
    assert liar.get_compromised_nodes() == {0, 1, 4}
    assert liar.get_false_claims(0, [3, 5], 0) == [1, 3, 4, 5]
    assert torch.equal(liar.apply_attack(0, _state(), 0)["w"], torch.full((4, 3), 2.0))
    plain = TopologyLiarAttack(10, 0.0, seed=42)
    assert len(plain.get_compromised_nodes()) == 1          # at least one liar
    st = _state(); assert plain.apply_attack(0, st, 0) is st
    import numpy as np
    adj = np.zeros((10, 10), dtype=bool); adj[0, 3] = adj[3, 0] = True
    claims = liar.claim_bitmask(adj)
    assert claims[0, 1] and claims[0, 4] and claims[0, 3] and not claims[0, 0] and not claims[3, 1]
