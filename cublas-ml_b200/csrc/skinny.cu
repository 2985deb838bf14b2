// skinny.cu — backward step of a NARROW layer behind a WIDE hidden layer (config C3/C4: 4096 or 8192 units -> 10).
//
// The reference runs it as two cuBLAS calls plus two elementwise kernels (cublasNN.cpp:673-676, 680-684):
//     delta_prev = (delta W)[:, no bias] .* f'(a)          K-deep contraction, K = 10
//     G          = delta^T [a, 1]                          10 x (H+1) outputs, contraction over the batch
// Both are pure streaming over the hidden activation a (rows x (H+1), 134 MB for C3): on tensor-core tiles they are
// > 90 % padding and each re-reads a.  This kernel reads every tile of a ONCE from HBM and produces both results:
//
//   block  = 32 hidden units x one slice of the samples, 256 threads; grid = (ceil((H+1)/32), sample splits)
//   tile   = 32 units x 128 samples staged in shared memory by coalesced 512-byte row reads (16 B per lane),
//            delta's K x 128 tile next to it
//   thread = one unit (lane) x 16 samples (warp): delta reads are warp-wide broadcasts, a reads are conflict-free
//            (row pitch 132 floats), per element 2 K FMA: K for (delta W), K for the gradient
//   out    = delta_prev written back into the tile, then stored with the same coalesced pattern;
//            G partials live in registers across the block's tiles, are summed over the 8 warps in shared memory and
//            added to the (pre-zeroed) gradient with one atomic per (k, unit) per block
// Exact fp32 arithmetic in every precision mode.  HBM traffic: read a, write delta_prev = 8 B per element.
#include "engine.h"
#include <cuda_bf16.h>
#include <cstdlib>
#include "pdl.cuh"
#include "../../include/b2nn.h"

namespace b2 {
namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_ROWS = 32;            // hidden units per block
constexpr int SK_COLS = 128;           // samples per tile
constexpr int SK_PITCH = SK_COLS + 4;  // floats; 16-byte aligned rows, conflict-free for lane = row

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    // 16-byte asynchronous copy global -> shared; src-size 0 zero-fills the destination without touching memory
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// Blackwell packed fp32 FMA: two independent multiply-adds per issued instruction (the kernel is issue-bound).
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// EXACT: K == KMAX, so the per-k guards (a branch each in the unrolled loops) disappear.
template <int KMAX, bool EXACT>
__global__ void __launch_bounds__(SK_THREADS, 4) skinny_bwd_kernel(const SkinnyArgs p) {
    pdl_enter();
    extern __shared__ __align__(16) float sk_smem[];
    constexpr int A_FLOATS = SK_ROWS * SK_PITCH, D_FLOATS = KMAX * SK_COLS, BUF_FLOATS = A_FLOATS + D_FLOATS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = EXACT ? KMAX : p.K;
    const int h0 = blockIdx.x * SK_ROWS;
    const int h = h0 + lane;                       // this thread's unit; h == H is the ones row (gradient of the bias only)
    unsigned long long wreg[KMAX], g[KMAX];        // (w, w) pairs; gradient partials of the even / odd samples
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        const float w = ((EXACT || k < K) && h < p.H) ? __ldg(p.W + int64_t(k) * p.ldw + h) : 0.f;
        wreg[k] = pack2(w, w);
        g[k] = 0ull;
    }
    const int64_t n_tiles = (p.rows + SK_COLS - 1) / SK_COLS;
    // the next tile streams into the other buffer (cp.async) while this one is computed and stored
    auto issue = [&](int64_t tile, int buf) {
        float* a_s = sk_smem + buf * BUF_FLOATS;
        float* d_s = a_s + A_FLOATS;
        const int64_t sc = tile * SK_COLS + lane * 4;          // this lane's 4 samples in the coalesced pattern
        const bool s_ok = sc < p.ld;                           // ld is a multiple of 32: a float4 is entirely inside or outside
        const int64_t scc = s_ok ? sc : 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int hh = h0 + warp * 4 + r;
            const bool ok = hh <= p.H && s_ok;
            cp_async16(a_s + (warp * 4 + r) * SK_PITCH + lane * 4, p.a + int64_t(ok ? hh : 0) * p.ld + scc, ok);
        }
        for (int k = warp; k < K; k += SK_THREADS / 32)
            cp_async16(d_s + k * SK_COLS + lane * 4, p.delta + int64_t(k) * p.ld + scc, s_ok);
        cp_async_commit();
    };
    int buf = 0;
    if (int64_t(blockIdx.y) < n_tiles) issue(blockIdx.y, 0);
    for (int64_t tile = blockIdx.y; tile < n_tiles; tile += gridDim.y, buf ^= 1) {
        float* a_s = sk_smem + buf * BUF_FLOATS;
        float* d_s = a_s + A_FLOATS;
        const int64_t next = tile + gridDim.y;
        if (next < n_tiles) { issue(next, buf ^ 1); cp_async_wait<1>(); } else cp_async_wait<0>();
        __syncthreads();
        const int64_t s0 = tile * SK_COLS;
        // samples past the batch end may hold a previous, longer batch's deltas: they count as zero.  Only the last tile
        // is affected, so it is fixed up in shared memory instead of masking inside the FMA loop (that cost 320 predicated
        // instructions per warp and tile: the kernel is bound by instruction issue, not by HBM).
        if (s0 + SK_COLS > p.rows) {
            for (int idx = threadIdx.x; idx < K * SK_COLS; idx += SK_THREADS)
                if (s0 + (idx & (SK_COLS - 1)) >= p.rows) d_s[idx] = 0.f;
            __syncthreads();
        }
        // ---- compute: unit = lane, samples [16 warp, 16 warp + 16) ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int so = warp * 16 + q * 4;
            const ulonglong2 a2 = *reinterpret_cast<const ulonglong2*>(a_s + lane * SK_PITCH + so);   // (a0,a1), (a2,a3)
            unsigned long long t01 = 0ull, t23 = 0ull;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (EXACT || k < K) {
                    const ulonglong2 d2 = *reinterpret_cast<const ulonglong2*>(d_s + k * SK_COLS + so);   // warp-wide broadcast
                    t01 = ffma2(d2.x, wreg[k], t01);
                    t23 = ffma2(d2.y, wreg[k], t23);
                    g[k] = ffma2(d2.x, a2.x, g[k]);
                    g[k] = ffma2(d2.y, a2.y, g[k]);
                }
            }
            float4 a4, t;
            unpack2(a2.x, a4.x, a4.y); unpack2(a2.y, a4.z, a4.w);
            unpack2(t01, t.x, t.y); unpack2(t23, t.z, t.w);
            if (p.act_kind == B2NN_ACT_SIGMOID) {
                a4.x = a4.x * (1.0f - a4.x); a4.y = a4.y * (1.0f - a4.y); a4.z = a4.z * (1.0f - a4.z); a4.w = a4.w * (1.0f - a4.w);
            } else {
                a4.x = 1.0f - a4.x * a4.x; a4.y = 1.0f - a4.y * a4.y; a4.z = 1.0f - a4.z * a4.z; a4.w = 1.0f - a4.w * a4.w;
            }
            *reinterpret_cast<float4*>(a_s + lane * SK_PITCH + so) = make_float4(t.x * a4.x, t.y * a4.y, t.z * a4.z, t.w * a4.w);
        }
        __syncthreads();
        // ---- store delta_prev with the load pattern ----
        const int64_t sc = s0 + lane * 4;
        if (sc < p.ld) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int hh = h0 + warp * 4 + r;
                if (hh < p.H) {
                    const float4 v = *reinterpret_cast<const float4*>(a_s + (warp * 4 + r) * SK_PITCH + lane * 4);
                    *reinterpret_cast<float4*>(p.delta_prev + int64_t(hh) * p.ld + sc) = v;
                    if (p.delta_prev_bf16) {       // bf16 mode: the copy the next contractions read, written in the same pass
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.delta_prev_bf16) + int64_t(hh) * p.ld + sc) =
                            make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                    }
                }
            }
        }
        __syncthreads();                               // the tile after next streams into this buffer
    }
    // ---- gradient: sum the 8 warps' partials, one atomic per (k, unit) ----
    float* red = sk_smem;                          // [8][KMAX][32] floats <= 32 * 132
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
        float g0, g1;
        unpack2(g[k], g0, g1);
        red[(warp * KMAX + k) * 32 + lane] = g0 + g1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < K * 32; idx += SK_THREADS) {
        const int k = idx >> 5, l = idx & 31;
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < SK_THREADS / 32; ++w) sum += red[(w * KMAX + k) * 32 + l];
        if (h0 + l <= p.H) atomicAdd(p.G + int64_t(k) * p.ldw + h0 + l, sum);
    }
}

template <int KMAX, bool EXACT>
cudaError_t launch_k(const SkinnyArgs& a, dim3 grid, cudaStream_t s) {
    constexpr size_t smem = sizeof(float) * 2 * (SK_ROWS * SK_PITCH + KMAX * SK_COLS);
    static std::atomic<unsigned long long> done{0};
    if (cudaError_t e = ensure_dyn_smem(skinny_bwd_kernel<KMAX, EXACT>, int(smem), done)) return e;
    return launch_pdl(skinny_bwd_kernel<KMAX, EXACT>, grid, dim3(SK_THREADS), smem, s, a);
}

}  // namespace

bool skinny_bwd_supported(int K, int H, int64_t ld) {
    if (std::getenv("B2NN_NO_SKINNY")) return false;
    return K >= 1 && K <= 16 && H >= 1024 && ld % 4 == 0;
}

// The caller has zeroed G (ldw * K floats).
cudaError_t skinny_bwd_launch(const SkinnyArgs& a, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    const int row_groups = (a.H + 1 + SK_ROWS - 1) / SK_ROWS;
    const int64_t n_tiles = (a.rows + SK_COLS - 1) / SK_COLS;
    // ~8 blocks per SM's worth of work (four are resident: 64 registers x 256 threads, 44 KB of shared memory).  Measured on
    // C3 (129 row groups): 4 / 5 / 9 / 10 splits -> 72 / 70 / 64 / 64 us; shorter blocks balance the SMs better.
    int64_t splits = (148 * 8 + row_groups - 1) / row_groups;
    if (const char* e = std::getenv("B2NN_SKINNY_SPLITS")) splits = std::atoi(e);
    if (splits > n_tiles) splits = n_tiles;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    const dim3 grid(unsigned(row_groups), unsigned(splits), 1);
    if (a.K == 10) return launch_k<10, true>(a, grid, s);       // the ten classes of MNIST-shaped problems
    if (a.K == 1) return launch_k<1, true>(a, grid, s);         // one regression target
    if (a.K <= 4) return launch_k<4, false>(a, grid, s);
    if (a.K <= 8) return launch_k<8, false>(a, grid, s);
    return launch_k<16, false>(a, grid, s);
}

}  // namespace b2
