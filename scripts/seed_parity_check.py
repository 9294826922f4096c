"""Round-by-round accuracy of the simulation backend vs the B200 engine in seed-parity mode (autograd path, eager, fused tape).

    python scripts/seed_parity_check.py [ALGO] [JSON params]      e.g.  krum '{"num_compromised": 1}'
"""
import sys; import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import torch, json
ALGO = sys.argv[1] if len(sys.argv) > 1 else "fedavg"
PARAMS = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
from test_engine_gpu import _cfg, _build
from murmura_b200.utils.seed import set_seed
data = {"adapter": "synthetic.mnist", "params": {"samples_per_node": 192, "partition_method": "dirichlet", "alpha": 0.5}}
topo = {"type": "k-regular", "num_nodes": 6, "k": 4}
for name, backend, b200 in (("sim", "simulation", {}), ("autograd", "b200", {"seed_parity": True, "fused_train": False, "krum_gram": "fp32"}),
                            ("autograd_nographs", "b200", {"seed_parity": True, "fused_train": False, "cuda_graphs": False}),
                            ("fused", "b200", {"seed_parity": True})):
    set_seed(11)
    cfg = _cfg(ALGO, PARAMS, n=6, topo=topo, data=data, b200=b200, backend=backend, rounds=10)
    net, _, _ = _build(cfg)
    if name == "sim":
        w0 = [p.detach().flatten()[:3].tolist() for p in net.nodes[1].model.parameters()][:1]
    else:
        w0 = [p.detach().flatten()[:3].tolist() for p in net.nodes[1].model.parameters()][:1]
    h = net.train(rounds=6, local_epochs=2, lr=0.05)
    print(name, w0, [round(float(a),4) for a in h["mean_accuracy"]], [round(float(a),4) for a in h.get("mean_loss", [])][:3])
    if hasattr(net, "close"): net.close()
