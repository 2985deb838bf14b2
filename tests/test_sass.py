"""CPU checks of the BUILT device code (no GPU needed: cuobjdump reads the sm_100a cubin inside libb2nn.so).

What they pin: the hot contractions really are tcgen05 / TMEM / TMA code (SASS mnemonics, B200_PROFILING.md), CTA pairs
and bulk stores are compiled in, no kernel on the hot path spills, the owner-side kernels of the peer exchange keep the
register budget that lets them share an SM with a GEMM CTA, and no GEMM library is linked (a library fallback would show
up as a NEEDED entry)."""
import re
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")


@pytest.fixture(scope="module")
def res_usage(b2nn):
    out = subprocess.run(["cuobjdump", "-res-usage", b2nn.LIB_PATH], capture_output=True, text=True, check=True).stdout
    table = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out):
        table[m.group(1)] = {"reg": int(m.group(2)), "stack": int(m.group(3)), "shared": int(m.group(4)), "local": int(m.group(5))}
    assert table, "no kernels found in libb2nn.so"
    return table


@pytest.fixture(scope="module")
def sass(b2nn):
    return subprocess.run(["cuobjdump", "-sass", "-arch", "sm_100a", b2nn.LIB_PATH], capture_output=True, text=True, check=True).stdout


def _kernels(table, needle):
    return {k: v for k, v in table.items() if needle in k}


def test_only_sm100a_code_is_embedded(b2nn):
    out = subprocess.run(["cuobjdump", "-lelf", b2nn.LIB_PATH], capture_output=True, text=True, check=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_tensor_core_contraction_is_tcgen05_tma(sass):
    counts = {m: sass.count(m) for m in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTCBAR", "LDTM", "UBLKCP", "SYNCS")}
    # tcgen05.mma (single CTA and cta_group::2), TMA tensor loads, tcgen05.commit, tcgen05.ld, cp.async.bulk stores of the
    # fused reduce-scatter epilogue, mbarrier waits
    for m, n in counts.items():
        assert n > 0, f"{m} missing from the SASS: {counts}"
    assert counts["UTCHMMA.2CTA"] >= counts["UTCHMMA"] // 4, counts


def test_gemm_tc_kernels_fit_the_register_budget_without_spills(res_usage):
    ks = _kernels(res_usage, "gemm_tc_kernel")
    assert len(ks) >= 60, len(ks)          # BN x majors x {1, 2} CTAs x {tf32, tf32x3, bf16}
    for name, r in ks.items():
        # 576 threads per CTA: 65536 / 576 = 113 registers at most; no local memory, no stack frame
        assert r["reg"] <= 112 and r["stack"] == 0 and r["local"] == 0, (name, r)


def test_peer_exchange_kernels_stay_co_resident_with_a_gemm_cta(res_usage):
    # 256-thread blocks, 8 per SM beside a 576-thread GEMM CTA -> at most 32 registers, no shared memory
    for needle in ("reduce_update_allgather_kernel", "small_reduce_update_kernel", "gate_layers_kernel", "gate_kernel",
                   "signal_kernel", "small_scatter_kernel"):
        ks = _kernels(res_usage, needle)
        assert ks, needle
        for name, r in ks.items():
            assert r["reg"] <= 32 and r["shared"] == 0 and r["local"] == 0, (name, r)


def test_streaming_kernels_do_not_spill(res_usage):
    for needle in ("momentum_update", "skinny_bwd_kernel", "tail_kernel", "output_layer_kernel", "normalise_rows_kernel",
                   "feature_stats_kernel", "micro_mlp_kernel"):
        ks = _kernels(res_usage, needle)
        assert ks, needle
        for name, r in ks.items():
            assert r["local"] == 0, (name, r)


def test_no_gemm_library_is_linked(b2nn):
    out = subprocess.run(["readelf", "-d", b2nn.LIB_PATH], capture_output=True, text=True, check=True).stdout
    needed = re.findall(r"\(NEEDED\)\s+Shared library: \[([^\]]+)\]", out)
    assert needed, out
    for lib in needed:
        assert not re.search(r"cublas|cudnn|cutlass|nccl", lib), needed        # NCCL is dlopen'ed only for N > 1
