"""Bring-up diagnostic for the tcgen05 GEMM (not a test): error statistics per operand-major combination, CTA-pair
mode and TMA box mode, then timings of the config-C3 contractions.  python tools/diag_tc.py [--perf-only]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_b2nn  # noqa: E402

b2nn = load_b2nn()
MAJORS = {"wgrad(K,K)": (0, 0), "fwd(MN,K)": (1, 0), "kmn(K,MN)": (0, 1), "bwd(MN,MN)": (1, 1)}
SHAPES = [(128, 256, 32), (256, 256, 64), (128, 16, 32), (384, 64, 96), (300, 70, 45), (4200, 50, 785),
          (785, 50, 4200), (1000, 513, 130), (2048, 768, 512)]


def operand(mn, k, major, gen, pad):
    if major == 0:
        ld = (k + 3) // 4 * 4 + 4
        t = torch.randn(mn, ld, device="cuda", generator=gen)
        return t, ld, t[:, :k].double()
    ld = (mn + pad - 1) // pad * pad + pad
    t = torch.randn(k + 1, ld, device="cuda", generator=gen)
    return t, ld, t[:k, :mn].double().t()


def run(net, backend, M, N, K, majors):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    pad = 32 if backend == 3 else 4
    A, lda, Ad = operand(M, K, majors[0], gen, pad)
    B, ldb, Bd = operand(N, K, majors[1], gen, pad)
    ldd = (M + 3) // 4 * 4
    D = torch.zeros(N, ldd, device="cuda")
    net.gemm(backend, A.data_ptr(), lda, majors[0], B.data_ptr(), ldb, majors[1], D.data_ptr(), ldd, M, N, K)
    torch.cuda.synchronize()
    ref = (Ad @ Bd.t()).t()
    bound = (Ad.abs() @ Bd.abs().t()).t() * 2.0 ** -9 * 1.25 + 1e-5
    err = (D[:, :M].double() - ref).abs()
    bad = err > bound
    return float(err.max()), float((err / bound).max()), float(bad.float().mean())


def perf(net):
    B, H, I, O = 8192, 4096, 784, 10
    # (name, M, N, K, a_major, b_major, epi)
    gemms = [("fwd0", B, H, I + 1, 1, 0, 2), ("fwd1", B, H, H + 1, 1, 0, 2), ("fwd2", B, O, H + 1, 1, 0, 0),
             ("bwd2", B, H, O, 1, 1, 4), ("bwd1", B, H, H, 1, 1, 4),
             ("wgrad0", I + 1, H, B, 0, 0, 0), ("wgrad1", H + 1, H, B, 0, 0, 0), ("wgrad2", H + 1, O, B, 0, 0, 0)]
    stream = torch.cuda.ExternalStream(net.stream())
    for name, M, N, K, am, bm, epi in gemms:
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
        for mode in ("pair", "single"):
            os.environ["B2NN_TC_NCTA"] = "1" if mode == "single" else "2"
            try:
                for _ in range(2):
                    net.gemm(3, A.data_ptr(), lda, am, Bm.data_ptr(), ldb, bm, D.data_ptr(), ldd, M, N, K, epi, aux.data_ptr(), ldd)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                reps = 5
                for _ in range(reps):
                    net.gemm(3, A.data_ptr(), lda, am, Bm.data_ptr(), ldb, bm, D.data_ptr(), ldd, M, N, K, epi, aux.data_ptr(), ldd)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"perf {name:7s} {mode:6s} M,N,K=({M},{N},{K}) {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"perf {name} {mode} EXC {ex}", flush=True)
                return
    os.environ["B2NN_TC_NCTA"] = "1"
    # yardstick only (never on the product path): cuBLAS TF32 on the same shapes through torch.matmul
    torch.backends.cuda.matmul.allow_tf32 = True
    for name, M, N, K in (("fwd1", B, H, H + 1), ("square", 8192, 8192, 8192), ("cublas-friendly", B, H, H)):
        a = torch.randn(M, K, device="cuda") * 0.05
        b = torch.randn(K, N, device="cuda") * 0.05
        for _ in range(3):
            torch.matmul(a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"yardstick cuBLAS tf32 {name:16s} M,N,K=({M},{N},{K}) {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
        a16, b16 = a.bfloat16(), b.bfloat16()
        for _ in range(3):
            torch.matmul(a16, b16)
        e0.record()
        for _ in range(10):
            torch.matmul(a16, b16)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"yardstick cuBLAS bf16 {name:16s} M,N,K=({M},{N},{K}) {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)


def sustained(net, seconds=3.0):
    """Power-limited regime: the same 8192 x 4096 x 4096 contraction back to back for `seconds`, ours and the cuBLAS
    yardsticks, with the SM clock / power seen at the end of each run (the tensor pipe is capped by board power)."""
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    M = 8192; N = 4096; K = 4096
    stream = torch.cuda.ExternalStream(net.stream())
    ld = K + 32
    A = torch.randn(M + 1, ld, device="cuda") * 0.05          # K-major both: the wgrad form
    Bm = torch.randn(N + 1, ld, device="cuda") * 0.05
    D = torch.empty(N, M, device="cuda")
    a = torch.randn(M, K, device="cuda") * 0.05
    b = torch.randn(K, N, device="cuda") * 0.05
    a16, b16 = a.bfloat16(), b.bfloat16()
    torch.backends.cuda.matmul.allow_tf32 = True
    os.environ.pop("B2NN_TC_NCTA", None)

    def ours():
        net.gemm(3, A.data_ptr(), ld, 0, Bm.data_ptr(), ld, 0, D.data_ptr(), M, M, N, K, 0, 0, M)
    cases = (("ours tcgen05 tf32 (pair)", ours, stream), ("cuBLAS tf32", lambda: torch.matmul(a, b), None),
             ("cuBLAS bf16", lambda: torch.matmul(a16, b16), None))
    for name, fn, st in cases:
        torch.cuda.synchronize()
        time.sleep(2.0)                                         # let the board cool back to boost clocks
        t_end = time.time() + seconds
        total = 0
        last_ms = None
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st) if st is not None else e0.record()
            for _ in range(100):
                fn()
            e1.record(st) if st is not None else e1.record()
            torch.cuda.synchronize()
            last_ms = e0.elapsed_time(e1) / 100
            total += 100
            if total == 100:
                first_ms = last_ms
        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        watts = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
        fl = 2.0 * M * N * K
        print(f"sustained {name:26s} first 100: {fl / first_ms / 1e9:7.1f} TFLOP/s   last 100 of {total}: {fl / last_ms / 1e9:7.1f} TFLOP/s"
              f"   {mhz} MHz {watts:.0f} W", flush=True)


def main():
    net = b2nn.Net()
    ok_all = True
    if "--sustained" in sys.argv:
        sustained(net)
        return
    if "--perf-only" not in sys.argv:
        for mode, backend, ncta in (("single/2d", 1, "1"), ("single/3d", 3, "1"), ("pair/2d", 1, "2"), ("pair/3d", 3, "2")):
            os.environ["B2NN_TC_NCTA"] = ncta
            for name, mj in MAJORS.items():
                for shp in SHAPES:
                    try:
                        e, ratio, frac = run(net, backend, *shp, mj)
                        status = "OK " if frac == 0 else "BAD"
                        ok_all &= frac == 0
                        print(f"{status} {mode:9s} {name:12s} M,N,K={shp} max_err={e:.3e} err/tol={ratio:.2f} bad_frac={frac:.4f}", flush=True)
                    except Exception as ex:  # noqa: BLE001
                        print(f"EXC {mode} {name} {shp}: {ex}", flush=True)
                        return 1
        print("ALL OK" if ok_all else "SOME BAD", flush=True)
    os.environ["B2NN_TC_NCTA"] = "1"
    perf(net)
    return 0


if __name__ == "__main__":
    sys.exit(main())
