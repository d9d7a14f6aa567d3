import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from spearmint_b200.backend import DeviceBackend
N, D, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
b = DeviceBackend()
rs = np.random.RandomState(0); X = rs.rand(N, D); y = rs.randn(N)
ll = b.loglik("Matern52", X, y)
hs = [(0.0, 1e-3, 1.0 + 0.01 * i, np.ones(D)) for i in range(B)]
for _ in range(3):
    ll.batch(hs)
torch.cuda.synchronize()
