"""Golden-value tests for topology generators and the mobility model."""
import numpy as np
import pytest

from murmura_b200.topology import MobilityModel, Topology, create_topology


def test_ring_small():
    t = create_topology("ring", 2)
    assert t.neighbors == [[1], [0]] and t.edges == [(0, 1)]
    t3 = create_topology("ring", 3)
    assert t3.neighbors == [[1, 2], [0, 2], [0, 1]]
    t10 = create_topology("ring", 10)
    assert all(t10.degree(i) == 2 for i in range(10)) and t10.is_connected()


def test_fully_and_aliases():
    t = create_topology("full", 5)
    assert all(t.degree(i) == 4 for i in range(5)) and len(t.edges) == 10
    assert create_topology("FULLY", 4).edges == create_topology("fully", 4).edges
    with pytest.raises(ValueError):
        create_topology("star", 4)


def test_k_regular_golden(capsys):
    t = create_topology("k-regular", 20, k=4)
    assert t.neighbors[0] == [1, 2, 18, 19] and len(t.edges) == 40
    t = create_topology("kregular", 5, k=3)
    out = capsys.readouterr().out
    assert "odd" in out
    assert len(t.edges) == 10       # k→4 on 5 nodes is the complete graph
    create_topology("k-regular", 4, k=6)
    assert ">= n" in capsys.readouterr().out


def test_erdos_golden():
    t = create_topology("erdos", 10, p=0.3, seed=12345)
    assert t.edges == [(0, 2), (0, 4), (0, 6), (0, 8), (0, 9), (1, 4), (1, 8), (2, 5), (2, 7), (2, 9), (3, 8),
                       (4, 7), (4, 8), (5, 9), (7, 9)]
    t16 = create_topology("er", 16, p=0.3, seed=12345)
    assert len(t16.edges) == 43
    assert [t16.degree(i) for i in range(16)] == [6, 5, 5, 5, 6, 2, 6, 8, 5, 4, 7, 2, 6, 6, 5, 8]
    assert t16.is_connected()
    with pytest.raises(ValueError):
        create_topology("erdos", 4, p=1.5)
    with pytest.raises(TypeError):
        create_topology("erdos", 4, p=None)


def test_erdos_isolated_repair():
    t = create_topology("erdos", 6, p=0.0, seed=1)
    assert all(t.degree(i) >= 1 for i in range(6))


def test_csr_and_adjacency():
    t = create_topology("ring", 4)
    row_ptr, cols = t.to_csr()
    assert row_ptr.tolist() == [0, 3, 6, 9, 12]
    assert cols[:3].tolist() == [0, 1, 3]
    adj = t.adjacency()
    assert adj.sum() == 8 and (adj == adj.T).all()


def test_topology_validation():
    with pytest.raises(AssertionError):
        Topology(num_nodes=3, neighbors=[[1], [0]], edges=[(0, 1)])


def test_mobility_golden():
    m = MobilityModel(10, 100, 40, 8, seed=42)
    np.testing.assert_allclose(m.positions_at(0)[:3], [[77.39560486, 43.88784398], [85.85979199, 69.73680291],
                                                       [9.41773479, 97.56223516]], rtol=1e-8)
    assert m.neighbors_at(0) == {0: [1, 3, 4, 7, 9], 1: [0, 2, 3, 4, 6, 9], 2: [1, 3, 5], 3: [0, 1, 2, 6, 8, 9],
                                 4: [0, 1, 7, 9], 5: [2, 6, 7, 8], 6: [1, 3, 5, 8, 9], 7: [0, 4, 5, 8],
                                 8: [3, 5, 6, 7], 9: [0, 1, 3, 4, 6]}
    m32 = MobilityModel(32, 100, 30, 5, seed=42)
    assert [sum(len(v) for v in m32.neighbors_at(r).values()) // 2 for r in (0, 1, 5)] == [130, 128, 135]


def test_mobility_sequential_and_symmetric():
    a, b = MobilityModel(8, seed=3), MobilityModel(8, seed=3)
    b.positions_at(5)
    np.testing.assert_array_equal(a.positions_at(5), b.positions_at(5))       # order of queries does not matter
    adj = a.adjacency_at(2)
    assert (adj == adj.T).all() and not adj.diagonal().any()
    assert a.torus_dist(0, 1, 2) == pytest.approx(a.torus_dist(1, 0, 2))
    assert a.positions_tensor(4).shape == (4, 8, 2)


def test_mobility_isolated_nodes_connected():
    m = MobilityModel(6, area_size=1000.0, comm_range=1.0, max_speed=1.0, seed=0, ensure_connected=True)
    assert all(len(v) >= 1 for v in m.neighbors_at(0).values())
    m2 = MobilityModel(6, area_size=1000.0, comm_range=1.0, max_speed=1.0, seed=0, ensure_connected=False)
    assert all(len(v) == 0 for v in m2.neighbors_at(0).values())
