import os, sys, torch, numpy as np
ROOT = os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_b2nn
b2nn = load_b2nn()
net = b2nn.Net()
M, N, K = 128, 16, 32
A = torch.zeros(M, K + 4, device="cuda"); B = torch.zeros(N, K + 4, device="cuda")
vals = [1 + 2**-11 + 2**-20, -(1 + 2**-11 + 2**-20), 1 + 2**-11, 1 + 2**-11 - 2**-20, 1 + 3 * 2**-11, 1 + 2**-10 + 2**-11 + 2**-20, 1 + 2**-12, 2**-11 + 2**-13]
for i, v in enumerate(vals):
    A[i, 0] = v
B[0, 0] = 1.0
D = torch.zeros(N, M, device="cuda")
net.gemm(1, A.data_ptr(), K + 4, 0, B.data_ptr(), K + 4, 0, D.data_ptr(), M, M, N, K)
torch.cuda.synchronize()
for i, v in enumerate(vals):
    f = np.float32(v)
    bits = f.view(np.uint32)
    trunc = np.uint32(bits & 0xFFFFE000).view(np.float32)
    rn_bits = np.uint32((int(bits) + 0x1000) & 0xFFFFE000)
    rn = rn_bits.view(np.float32)
    print(f"x={float(f)!r:>22} got={float(D[0, i])!r:>22} trunc={float(trunc)!r:>22} rn_half_up={float(rn)!r}")
# B side rounding
A.zero_(); B.zero_(); A[0, 0] = 1.0
for i, v in enumerate(vals[:N]):
    B[i, 0] = v
net.gemm(1, A.data_ptr(), K + 4, 0, B.data_ptr(), K + 4, 0, D.data_ptr(), M, M, N, K)
torch.cuda.synchronize()
print("B side:", [float(D[i, 0]) for i in range(len(vals))])
