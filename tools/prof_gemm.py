"""Run ONE config-C3 contraction a few times (for ncu).  python tools/prof_gemm.py <fwd0|fwd1|bwd1|bwd2|wgrad1|...> <single|pair> [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_b2nn  # noqa: E402

b2nn = load_b2nn()
B, H, I, O = 8192, 4096, 784, 10
GEMMS = {"fwd0": (B, H, I + 1, 1, 0, 2), "fwd1": (B, H, H + 1, 1, 0, 2), "fwd2": (B, O, H + 1, 1, 0, 0),
         "bwd2": (B, H, O, 1, 1, 4), "bwd1": (B, H, H, 1, 1, 4), "wgrad0": (I + 1, H, B, 0, 0, 0),
         "wgrad1": (H + 1, H, B, 0, 0, 0), "wgrad2": (H + 1, O, B, 0, 0, 0)}


def main():
    name, mode = sys.argv[1], sys.argv[2]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    os.environ["B2NN_TC_NCTA"] = "2" if mode == "pair" else "1"
    M, N, K, am, bm, epi = GEMMS[name]
    net = b2nn.Net()

    def buf(mn, k, major):
        if major == 0:
            ld = (k + 31) // 32 * 32 + 32
            return torch.randn(mn + 1, ld, device="cuda") * 0.05, ld
        ld = (mn + 31) // 32 * 32 + 32
        return torch.randn(k + 1, ld, device="cuda") * 0.05, ld

    A, lda = buf(M, K, am)
    Bm, ldb = buf(N, K, bm)
    ldd = (M + 31) // 32 * 32
    D = torch.empty(N, ldd, device="cuda")
    aux = torch.rand(N, ldd, device="cuda")
    for _ in range(reps):
        net.gemm(3, A.data_ptr(), lda, am, Bm.data_ptr(), ldb, bm, D.data_ptr(), ldd, M, N, K, epi, aux.data_ptr(), ldd)
    torch.cuda.synchronize()
    print("done", name, mode)


if __name__ == "__main__":
    main()
