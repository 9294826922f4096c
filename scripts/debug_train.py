import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from murmura_b200 import Network
from murmura_b200.config import Config
from murmura_b200.utils.factories import build_aggregator_factory, build_dataset_adapter, build_model_factory

def mk(graphs, n=1):
    cfg = Config(**{"experiment": {"name": "d", "rounds": 3, "seed": 3}, "topology": {"type": "fully", "num_nodes": max(n,2)},
        "aggregation": {"algorithm": "fedavg"}, "training": {"batch_size": 32, "lr": 0.05},
        "data": {"adapter": "synthetic.mnist", "params": {"samples_per_node": 128, "partition_method": "iid"}},
        "model": {"factory": "models.mlp", "params": {"hidden_dims": [32]}}, "backend": "b200", "b200": {"cuda_graphs": graphs}})
    ad = build_dataset_adapter(cfg); mf = build_model_factory(cfg)
    return Network.from_config(cfg, mf, ad, build_aggregator_factory(cfg, mf), device=torch.device("cuda")), mf

res = {}
for graphs in (False, True):
    net, mf = mk(graphs)
    vn = net.nodes[0]
    init = net.live[0].clone()
    net._prepare_training(1, 0.05)
    assert torch.equal(init, net.live[0]), "capture changed params"
    # deterministic perm
    vn.perm_buf.copy_(torch.arange(vn.perm_buf.numel(), device=net.device) % vn.n); vn.step.zero_()
    vn.model.train()
    for s in range(4):
        if vn.train_graph is not None: vn.train_graph.replay()
        else: net._train_step(vn, 0.05)
    torch.cuda.synchronize()
    res[graphs] = net.live[0].clone()
    print("graphs", graphs, "step", vn.step.item(), "delta", (res[graphs]-init).abs().sum().item())
    if not graphs:
        # manual torch reference
        m = mf().cuda()
        sd = {k: v.clone() for k, v in net.layout.row_views(init, None).items()}
        m.load_state_dict(sd)
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        for s in range(4):
            idx = torch.arange(s*32, (s+1)*32, device="cuda") % vn.n
            opt.zero_grad(); loss = torch.nn.functional.cross_entropy(m(vn.X[idx]), vn.y[idx]); loss.backward(); opt.step()
        ref = torch.cat([p.detach().flatten() for p in m.parameters()])
        mine = torch.cat([net.layout.row_views(res[False], None)[k].flatten() for k, _ in m.named_parameters()])
        print("eager vs manual max diff", (ref - mine).abs().max().item())
    net.close()
print("eager vs graph max diff", (res[True]-res[False]).abs().max().item())
for graphs in (False, True):
    net, _ = mk(graphs, n=4)
    h = net.train(rounds=8, lr=0.05)
    print("graphs", graphs, [round(float(a),3) for a in h["mean_accuracy"]])
    net.close()

# ---- sketch filter hist debug ----
from murmura_b200 import ops
ext = ops.ext()
DEV = "cuda"
V, K = 3, 200
g = torch.Generator().manual_seed(13)
own = torch.randn(V, K, generator=g)
pubsk = torch.zeros(2, V, K); pubsk[1] = own + 0.05 * torch.randn(V, K, generator=g); pubsk[1, 2] = own[2] * -5.0
own_d = own.to(DEV); pub_d = pubsk.to(DEV)
rows, slots = [0], []
for v, nb in enumerate([[1, 2], [0, 2], [0, 1]]):
    slots += [v] + nb; rows.append(len(slots))
E = len(slots)
et = {"row_ptr": torch.tensor(rows, dtype=torch.int32, device=DEV), "src_rank": torch.zeros(E, dtype=torch.int32, device=DEV),
      "src_slot": torch.tensor(slots, dtype=torch.int32, device=DEV), "mask": torch.ones(E, device=DEV), "w": torch.zeros(E, device=DEV),
      "w_tail": torch.zeros(E, device=DEV), "stats": torch.zeros(V, 4, device=DEV)}
tbl = torch.tensor([pub_d.data_ptr()], dtype=torch.int64, device=DEV)
hist = torch.zeros(V, 4, device=DEV); dist = torch.zeros(E, device=DEV)
for it in range(2):
    ext.sketchguard_filter(V, et["row_ptr"], et["src_rank"], et["src_slot"], et["mask"], et["w"], et["w_tail"], et["stats"], own_d, tbl.data_ptr(), 0, 0, 1 * V, K, 224, False, 0.5, 0.5, 1, hist, dist, 0, 1, 0, 0.0, 0)
    torch.cuda.synchronize()
    print("hist", hist.cpu().tolist()); print("stats", et["stats"].cpu().tolist()); print("w", et["w"].cpu().tolist()); print("dist", dist.cpu().tolist())
