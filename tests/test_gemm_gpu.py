"""GPU unit tests of the contraction kernels through the C ABI (b2nn_gemm), both backends:
backend 0 = fp32 FFMA, backend 1 = tcgen05 kind::tf32.  Reference = the same contraction in fp64 (torch on the GPU is
only plumbing here: device buffers and the fp64 check).  Shapes cover every operand major-ness used by the train
step, tails in M, N and K, and the fused epilogues."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

MAJORS = {"fwd": (1, 0), "bwd": (1, 1), "wgrad": (0, 0), "kmn": (0, 1)}
SHAPES = [(128, 256, 64), (256, 128, 96), (300, 70, 45), (4200, 50, 785), (785, 50, 4200), (200, 10, 4097),
          (129, 17, 33), (64, 300, 8), (1000, 513, 130)]


def _operand(mn, k, major, gen, pad=4):
    """Returns (tensor, ld, logical fp64 matrix [mn, k]); storage has a padded leading dimension full of junk."""
    if major == 0:   # K contiguous: [mn][ld]
        ld = (k + pad - 1) // pad * pad + pad
        t = torch.randn(mn, ld, device="cuda", generator=gen, dtype=torch.float32)
        return t, ld, t[:, :k].double()
    ld = (mn + pad - 1) // pad * pad + pad
    t = torch.randn(k + 1, ld, device="cuda", generator=gen, dtype=torch.float32)   # +1 row: slack for backend 3
    if pad >= 32:
        t[:, mn:] = float("nan")     # the slack may be read (backend 3) but must never reach a stored output
    return t, ld, t[:k, :mn].double().t()


def _run(net, b2nn, backend, M, N, K, majors, epi=0, seed=0):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    a_major, b_major = majors
    pad = {3: 32, 7: 32, 9: 8, 11: 64}.get(backend, 4)      # bf16 rows must start 16-byte aligned: ld % 8 == 0
    A, lda, Ad = _operand(M, K, a_major, gen, pad)
    B, ldb, Bd = _operand(N, K, b_major, gen, pad)
    ldd = (M + 3) // 4 * 4 + 8
    D = torch.full((N, ldd), 7.0, device="cuda", dtype=torch.float32)
    aux = torch.tanh(torch.randn(N, ldd, device="cuda", generator=gen, dtype=torch.float32)) * 0.9
    net.gemm(backend, A.data_ptr(), lda, a_major, B.data_ptr(), ldb, b_major, D.data_ptr(), ldd, M, N, K, epi,
             aux.data_ptr() if epi >= 3 else 0, ldd)
    torch.cuda.synchronize()
    ref = (Ad @ Bd.t()).t().contiguous()          # [N, M] like D
    bound = (Ad.abs() @ Bd.abs().t()).t()
    if epi == 1: ref = torch.sigmoid(ref)
    elif epi == 2: ref = torch.tanh(ref)
    elif epi == 3: a = aux[:, :M].double(); ref = ref * a * (1 - a); bound = bound * (a * (1 - a)).abs()
    elif epi == 4: a = aux[:, :M].double(); ref = ref * (1 - a * a); bound = bound * (1 - a * a).abs()
    got = D[:, :M].double()
    assert torch.all(D[:, M:] == 7.0), "kernel wrote outside the M extent"
    return got, ref, bound


@pytest.fixture(scope="module")
def net(b2nn):
    n = b2nn.Net()
    yield n
    n.close()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", list(MAJORS))
def test_ffma_gemm_matches_fp64(net, b2nn, kind, shape):
    M, N, K = shape
    got, ref, bound = _run(net, b2nn, 0, M, N, K, MAJORS[kind])
    err = (got - ref).abs()
    assert torch.all(err <= 4e-6 * bound + 1e-6), (kind, shape, float(err.max()))


@pytest.mark.parametrize("epi", [0, 2, 4])
@pytest.mark.parametrize("shape", [(2000, 1300, 70), (1921, 1283, 45)])
@pytest.mark.parametrize("kind", list(MAJORS))
def test_ffma_gemm_throughput_kernel(net, b2nn, kind, shape, epi):
    """Shapes with at least one 128 x 128 block per SM take the 8 x 8-per-thread FFMA kernel (ragged edges in M, N, K)."""
    M, N, K = shape
    got, ref, bound = _run(net, b2nn, 0, M, N, K, MAJORS[kind], epi=epi, seed=epi)
    err = (got - ref).abs()
    assert torch.all(err <= 4e-6 * bound + 2e-6), (kind, shape, epi, float(err.max()))


@pytest.mark.parametrize("backend", [1, 3])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", list(MAJORS))
def test_tcgen05_gemm_matches_fp64_within_tf32(net, b2nn, kind, shape, backend):
    """TF32 keeps 10 mantissa bits of each input: |err| <= ~2^-9 * sum|a||b| per output element.
    backend 3 = one 3-D TMA box per M/N-contiguous tile (padded operands, NaN in the padding)."""
    M, N, K = shape
    got, ref, bound = _run(net, b2nn, backend, M, N, K, MAJORS[kind])
    err = (got - ref).abs()
    tol = 1.25 * 2.0 ** -9 * bound + 1e-5
    bad = err > tol
    assert not bool(bad.any()), (kind, shape, float(err.max()), int(bad.sum()), bad.nonzero()[:4].tolist())


@pytest.mark.parametrize("backend", [5, 7])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", list(MAJORS))
def test_tcgen05_3xtf32_gemm_is_fp32_grade(net, b2nn, kind, shape, backend):
    """fp32-grade tensor-core mode: hi*hi + hi*lo + lo*hi with hi = trunc_tf32(x), lo = x - hi (lo is itself truncated
    to 11 bits by the MMA, the lo*lo term is dropped): per product error <= ~3 * 2^-21 |a||b|, i.e. >= 2^11 times tighter
    than plain tf32; the bound below also covers fp32 accumulation over K."""
    M, N, K = shape
    got, ref, bound = _run(net, b2nn, backend, M, N, K, MAJORS[kind])
    err = (got - ref).abs()
    tol = 2.0 ** -18 * bound + 1e-6
    bad = err > tol
    assert not bool(bad.any()), (kind, shape, float(err.max()), float((err / bound).max()), int(bad.sum()))


@pytest.mark.parametrize("backend", [9, 11])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", list(MAJORS))
def test_tcgen05_bf16_gemm_matches_fp64_within_bf16(net, b2nn, kind, shape, backend):
    """bf16 mode: both operands rounded to nearest-even bf16 (8 significant bits: relative error <= 2^-8 each), fp32
    accumulation in TMEM: |err| <= ~2^-7 * sum|a||b| per output element (reached only by short K: errors average out).  backend 11 = one 3-D TMA box per
    M/N-contiguous tile (64-element boxes; NaN in the padding must never reach a stored output)."""
    M, N, K = shape
    got, ref, bound = _run(net, b2nn, backend, M, N, K, MAJORS[kind])
    err = (got - ref).abs()
    tol = 1.05 * 2.0 ** -7 * bound + 1e-5
    bad = err > tol
    assert not bool(bad.any()), (kind, shape, float(err.max()), int(bad.sum()), bad.nonzero()[:4].tolist())


@pytest.mark.parametrize("epi", [1, 2, 3, 4])
@pytest.mark.parametrize("backend", [0, 1, 3])
def test_fused_epilogues(net, b2nn, backend, epi):
    M, N, K = 520, 96, 200
    got, ref, bound = _run(net, b2nn, backend, M, N, K, MAJORS["fwd" if epi < 3 else "bwd"], epi=epi, seed=epi)
    err = (got - ref).abs()
    rel = 4e-6 if backend == 0 else 1.25 * 2.0 ** -9
    # activations are 1-Lipschitz (sigmoid 1/4): the pre-activation bound carries over
    assert torch.all(err <= rel * bound + 2e-6), (backend, epi, float(err.max()))


def test_momentum_update_kernel(net, b2nn):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    for n in (1, 3, 4, 1023, 4096 * 33 + 2):
        th = torch.randn(n, device="cuda", generator=gen)
        v = torch.randn(n, device="cuda", generator=gen)
        g = torch.randn(n, device="cuda", generator=gen)
        v_ref = 0.9 * v.double() + (-0.01) * g.double()
        th_ref = th.double() + v_ref
        net.momentum_update(th.data_ptr(), v.data_ptr(), g.data_ptr(), n, 0.9, -0.01)
        torch.cuda.synchronize()
        assert torch.allclose(v.double(), v_ref, rtol=0, atol=1e-6)
        assert torch.allclose(th.double(), th_ref, rtol=0, atol=1e-6)
