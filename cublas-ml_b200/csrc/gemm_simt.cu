// gemm_simt.cu — fp32 FFMA contraction with the fused epilogues of engine.h.
//
// This is the "exact fp32" backend: same arithmetic class as the reference's default-math cublasSgemm
// (cublasNN.h:100-111, cublasNN.cpp:25 never changes the math mode), differing only in summation order.  It takes
// every shape (any M, N, K, any leading dimension, both operand major-nesses) and is what small or TMA-hostile
// layers run on; large layers go to the tcgen05 kernel in gemm_tc.cu when the precision mode allows it.
#include "engine.h"
#include <cstdlib>
#include "pdl.cuh"
#include "epilogue.cuh"

namespace b2 {
namespace {

constexpr int BM = 64, BN = 64, NT = 256;      // BK (16 or 32) is a template parameter of the kernel
constexpr int PAD = 4;

// Loads one BK x 64 tile of an operand into registers (BK/4 values per thread), zero-filled outside [0,mn_ext)x[k_lo,k_hi).
template <bool MN_MAJOR, int BK>
__device__ __forceinline__ void load_tile(const float* __restrict__ base, int64_t ld, int mn0, int mn_ext, int kt,
                                          int k_hi, float (&r)[BK / 4]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < BK / 4; ++i) {
        const int e = t + NT * i;
        int mn, k;
        if (MN_MAJOR) { k = e >> 6; mn = e & 63; } else { mn = e / BK; k = e % BK; }
        const int gm = mn0 + mn, gk = kt + k;
        float v = 0.f;
        if (gm < mn_ext && gk < k_hi)
            v = MN_MAJOR ? __ldg(base + int64_t(gk) * ld + gm) : __ldg(base + int64_t(gm) * ld + gk);
        r[i] = v;
    }
}
template <bool MN_MAJOR, int BK>
__device__ __forceinline__ void store_tile(float (*s)[64 + PAD], const float (&r)[BK / 4]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < BK / 4; ++i) {
        const int e = t + NT * i;
        int mn, k;
        if (MN_MAJOR) { k = e >> 6; mn = e & 63; } else { mn = e / BK; k = e % BK; }
        s[k][mn] = r[i];
    }
}

// BK = 32 halves the number of load -> sync -> compute rounds: small contractions (one block per SM or fewer) are a
// chain of global-load latencies, not of FMAs.
template <bool A_MN, bool B_MN, int BK>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmDesc g, int k_per_split) {
    pdl_enter();
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int k_lo = blockIdx.z * k_per_split;
    const int k_hi = min(g.K, k_lo + k_per_split);
    const float* A = g.A + (A_MN ? g.a_k0 * g.lda + g.a_mn0 : g.a_mn0 * g.lda + g.a_k0);
    const float* B = g.B + (B_MN ? g.b_k0 * g.ldb + g.b_mn0 : g.b_mn0 * g.ldb + g.b_k0);

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

    float ra[BK / 4], rb[BK / 4];
    int buf = 0;
    if (k_lo < k_hi) {
        load_tile<A_MN, BK>(A, g.lda, m0, g.M, k_lo, k_hi, ra);
        load_tile<B_MN, BK>(B, g.ldb, n0, g.N, k_lo, k_hi, rb);
        store_tile<A_MN, BK>(As[0], ra);
        store_tile<B_MN, BK>(Bs[0], rb);
    }
    __syncthreads();
    for (int kt = k_lo; kt < k_hi; kt += BK) {
        const bool more = kt + BK < k_hi;
        if (more) {
            load_tile<A_MN, BK>(A, g.lda, m0, g.M, kt + BK, k_hi, ra);
            load_tile<B_MN, BK>(B, g.ldb, n0, g.N, kt + BK, k_hi, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][tx * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
        }
        if (more) {
            store_tile<A_MN, BK>(As[buf ^ 1], ra);
            store_tile<B_MN, BK>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + ty * 4 + j;
        if (n >= g.N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + tx * 4 + i;
            if (m >= g.M) continue;
            float v = acc[j][i];
            if (g.epi != EPI_STORE) {
                const float a = (g.epi >= EPI_DSIGMOID) ? __ldg(g.aux + int64_t(n) * g.ldaux + m) : 0.f;
                v = apply_epilogue(g.epi, v, a);
            }
            float* dst = g.D + int64_t(n) * g.ldd + m;
            if (atomic) atomicAdd(dst, v); else *dst = v;
        }
    }
}

// Throughput regime: 128 x 128 x 16 tiles, 8 x 8 outputs per thread (two 4-wide strips 64 apart in each direction, so a
// half-warp's 16-byte shared-memory reads cover 256 contiguous bytes), 4 shared-memory reads per 64 FMA.  Same
// interface, epilogues and split-K as the 64 x 64 kernel above, which stays for everything that does not fill the SMs.
constexpr int GM = 128, GN = 128, GK = 16;
template <bool MN_MAJOR>
__device__ __forceinline__ void load_tile_big(const float* __restrict__ base, int64_t ld, int mn0, int mn_ext, int kt,
                                              int k_hi, float (&r)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = t + NT * i;
        int mn, k;
        if (MN_MAJOR) { k = e >> 7; mn = e & 127; } else { mn = e >> 4; k = e & 15; }
        const int gm = mn0 + mn, gk = kt + k;
        float v = 0.f;
        if (gm < mn_ext && gk < k_hi)
            v = MN_MAJOR ? __ldg(base + int64_t(gk) * ld + gm) : __ldg(base + int64_t(gm) * ld + gk);
        r[i] = v;
    }
}
template <bool MN_MAJOR>
__device__ __forceinline__ void store_tile_big(float (*s)[GM + PAD], const float (&r)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = t + NT * i;
        int mn, k;
        if (MN_MAJOR) { k = e >> 7; mn = e & 127; } else { mn = e >> 4; k = e & 15; }
        s[k][mn] = r[i];
    }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NT, 2) gemm_simt_big_kernel(GemmDesc g, int k_per_split) {
    pdl_enter();
    __shared__ __align__(16) float As[2][GK][GM + PAD];
    __shared__ __align__(16) float Bs[2][GK][GN + PAD];
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
    const int k_lo = blockIdx.z * k_per_split;
    const int k_hi = min(g.K, k_lo + k_per_split);
    const float* A = g.A + (A_MN ? g.a_k0 * g.lda + g.a_mn0 : g.a_mn0 * g.lda + g.a_k0);
    const float* B = g.B + (B_MN ? g.b_k0 * g.ldb + g.b_mn0 : g.b_mn0 * g.ldb + g.b_k0);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
    float ra[8], rb[8];
    int buf = 0;
    if (k_lo < k_hi) {
        load_tile_big<A_MN>(A, g.lda, m0, g.M, k_lo, k_hi, ra);
        load_tile_big<B_MN>(B, g.ldb, n0, g.N, k_lo, k_hi, rb);
        store_tile_big<A_MN>(As[0], ra);
        store_tile_big<B_MN>(Bs[0], rb);
    }
    __syncthreads();
    for (int kt = k_lo; kt < k_hi; kt += GK) {
        const bool more = kt + GK < k_hi;
        if (more) {
            load_tile_big<A_MN>(A, g.lda, m0, g.M, kt + GK, k_hi, ra);
            load_tile_big<B_MN>(B, g.ldb, n0, g.N, kt + GK, k_hi, rb);
        }
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][ty * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + ty * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(av[i], bv[j], acc[j][i]);
        }
        if (more) {
            store_tile_big<A_MN>(As[buf ^ 1], ra);
            store_tile_big<B_MN>(Bs[buf ^ 1], rb);
        }
        __syncthreads();
        buf ^= 1;
    }
    const bool atomic = gridDim.z > 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = n0 + (j < 4 ? ty * 4 + j : 64 + ty * 4 + (j - 4));
        if (n >= g.N) continue;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
            if (m >= g.M) continue;
            float v = acc[j][i];
            if (g.epi != EPI_STORE) {
                const float a = (g.epi >= EPI_DSIGMOID) ? __ldg(g.aux + int64_t(n) * g.ldaux + m) : 0.f;
                v = apply_epilogue(g.epi, v, a);
            }
            float* dst = g.D + int64_t(n) * g.ldd + m;
            if (atomic) atomicAdd(dst, v); else *dst = v;
        }
    }
}

cudaError_t launch_big(const GemmDesc& g, int splits, cudaStream_t stream) {
    int k_per = (g.K + splits - 1) / splits;
    k_per = ((k_per + GK - 1) / GK) * GK;
    if (k_per <= 0) k_per = GK;
    splits = (g.K + k_per - 1) / k_per;
    if (splits < 1) splits = 1;
    dim3 grid((g.M + GM - 1) / GM, (g.N + GN - 1) / GN, splits);
    if (g.a_major && g.b_major)       return launch_pdl(gemm_simt_big_kernel<true, true>, grid, dim3(NT), 0, stream, g, k_per);
    else if (g.a_major && !g.b_major) return launch_pdl(gemm_simt_big_kernel<true, false>, grid, dim3(NT), 0, stream, g, k_per);
    else if (!g.a_major && g.b_major) return launch_pdl(gemm_simt_big_kernel<false, true>, grid, dim3(NT), 0, stream, g, k_per);
    return launch_pdl(gemm_simt_big_kernel<false, false>, grid, dim3(NT), 0, stream, g, k_per);
}

template <int BK>
cudaError_t launch_bk(const GemmDesc& g, int splits, cudaStream_t stream) {
    int k_per = (g.K + splits - 1) / splits;
    k_per = ((k_per + BK - 1) / BK) * BK;
    if (k_per <= 0) k_per = BK;
    splits = (g.K + k_per - 1) / k_per;
    if (splits < 1) splits = 1;
    dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN, splits);
    if (g.a_major && g.b_major)       return launch_pdl(gemm_simt_kernel<true, true, BK>, grid, dim3(NT), 0, stream, g, k_per);
    else if (g.a_major && !g.b_major) return launch_pdl(gemm_simt_kernel<true, false, BK>, grid, dim3(NT), 0, stream, g, k_per);
    else if (!g.a_major && g.b_major) return launch_pdl(gemm_simt_kernel<false, true, BK>, grid, dim3(NT), 0, stream, g, k_per);
    return launch_pdl(gemm_simt_kernel<false, false, BK>, grid, dim3(NT), 0, stream, g, k_per);
}

}  // namespace

cudaError_t gemm_simt_launch(const GemmDesc& g, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    const int splits = g.split_k > 1 && g.epi == EPI_STORE ? g.split_k : 1;
    const int64_t blocks = int64_t((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN) * splits;
    // throughput regime: at least one 128 x 128 block per SM
    const int64_t big_blocks = int64_t((g.M + GM - 1) / GM) * ((g.N + GN - 1) / GN) * splits;
    if (big_blocks >= 148 && !std::getenv("B2NN_SIMT_NO_BIG")) return launch_big(g, splits, stream);
    // latency-bound regime (at most two blocks per SM): deeper K tiles; otherwise the 16-deep tile
    if (blocks <= 2 * 148 && g.K / splits > 16) return launch_bk<32>(g, splits, stream);
    return launch_bk<16>(g, splits, stream);
}

}  // namespace b2
