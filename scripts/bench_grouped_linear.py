import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from murmura_b200 import ops
ext = ops.ext()
G, M, K, N = 90, 128, 561, 256
X = torch.randn(10, M, K, device="cuda"); W = torch.randn(10, N * K + 4 * N + 8, device="cuda"); Y = torch.zeros(G, M, N, device="cuda")
host = np.zeros((G, 9), dtype=np.int64)
for g in range(G):
    d, c = g // 9, g % 10
    host[g] = (X[d].data_ptr(), W[c].data_ptr() + 4, W[c].data_ptr() + 4 + N * K * 4, 0, 0, 0, 0, Y[g].data_ptr(), M)
desc = torch.from_numpy(host).cuda()
for _ in range(5):
    ext.grouped_linear_tf32(desc, G, M, K, N, K, N, 1, 1e-5)
torch.cuda.synchronize()
