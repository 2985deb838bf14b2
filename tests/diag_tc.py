"""Bring-up diagnostic for the tcgen05 GEMM (not a test): prints error statistics per operand-major combination and
shape so one GPU call tells which descriptor path is wrong.  python tests/diag_tc.py [--sweep]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_b2nn  # noqa: E402

b2nn = load_b2nn()
MAJORS = {"wgrad(K,K)": (0, 0), "fwd(MN,K)": (1, 0), "kmn(K,MN)": (0, 1), "bwd(MN,MN)": (1, 1)}
SHAPES = [(128, 256, 32), (128, 256, 64), (128, 64, 32), (128, 16, 32), (256, 256, 128), (300, 70, 45),
          (4200, 50, 785), (785, 50, 4200), (1000, 513, 130)]


def operand(mn, k, major, gen):
    if major == 0:
        ld = (k + 3) // 4 * 4 + 4
        t = torch.randn(mn, ld, device="cuda", generator=gen)
        return t, ld, t[:, :k].double()
    ld = (mn + 3) // 4 * 4 + 4
    t = torch.randn(k, ld, device="cuda", generator=gen)
    return t, ld, t[:, :mn].double().t()


def run(net, M, N, K, majors):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    A, lda, Ad = operand(M, K, majors[0], gen)
    B, ldb, Bd = operand(N, K, majors[1], gen)
    ldd = (M + 3) // 4 * 4
    D = torch.zeros(N, ldd, device="cuda")
    net.gemm(1, A.data_ptr(), lda, majors[0], B.data_ptr(), ldb, majors[1], D.data_ptr(), ldd, M, N, K)
    torch.cuda.synchronize()
    ref = (Ad @ Bd.t()).t()
    bound = (Ad.abs() @ Bd.abs().t()).t() * 2.0 ** -9 * 1.25 + 1e-5
    err = (D[:, :M].double() - ref).abs()
    bad = err > bound
    return float(err.max()), float((err / bound).max()), float(bad.float().mean()), float(ref.abs().mean())


def main():
    net = b2nn.Net()
    ok_all = True
    for name, mj in MAJORS.items():
        for shp in SHAPES:
            try:
                e, ratio, frac, mag = run(net, *shp, mj)
                status = "OK " if frac == 0 else "BAD"
                ok_all &= frac == 0
                print(f"{status} {name:12s} M,N,K={shp} max_err={e:.3e} err/tol={ratio:.2f} bad_frac={frac:.4f} |ref|={mag:.2f}", flush=True)
            except Exception as ex:  # noqa: BLE001
                ok_all = False
                print(f"EXC {name} {shp}: {ex}", flush=True)
                return 1
    if "--sweep" in sys.argv and not ok_all:
        for lbo, sbo in ((1024, 4096), (4096, 128), (128, 1024), (1024, 128), (128, 4096)):
            os.environ["B2NN_TC_MN_LBO"], os.environ["B2NN_TC_MN_SBO"] = str(lbo), str(sbo)
            for name, mj in MAJORS.items():
                if mj == (0, 0):
                    continue
                e, ratio, frac, mag = run(net, 128, 256, 64, mj)
                print(f"sweep LBO={lbo} SBO={sbo} {name:12s} max_err={e:.3e} bad_frac={frac:.4f}", flush=True)
    print("ALL OK" if ok_all else "SOME BAD")
    return 0


if __name__ == "__main__":
    sys.exit(main())
