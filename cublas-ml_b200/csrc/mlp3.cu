// mlp3.cu — a whole training step of a ONE-HIDDEN-LAYER network (config C1: MNIST 784-50-10, main.cpp:126-157) in one
// kernel plus a reduce/update kernel; the same kernel, forward only, is config C5's inference path in the fp32-grade mode.
//
// The reference spends 15 launches per step on this net (cublasNN.cpp:597-629), round 1 of this engine four (tcgen05
// forward, fused tail, tcgen05 split-K gradient, update: 37-43 us for ~1 us of tensor work).  Here a block owns 32 samples:
//   load     X tile (l0+1 features x 32 samples, 113 KB for MNIST) -> shared memory with cp.async, ONCE; it feeds both the
//            forward contraction and the input-layer gradient
//   phase 1  hidden = f([x,1] W0^T): warp-level tensor-core MMA (mma.sync m16n8k8 tf32), the eight warps split the feature
//            dimension (split-K), partial sums meet in shared memory; W0 fragments come straight from L2 as 16-byte loads
//   phase 2  output layer, output activation, delta, cost / argmax / error count, delta backprop, output-layer gradient:
//            fp32 FFMA out of shared memory (a few thousand multiply-adds per sample tile)
//   phase 3  input-layer gradient of the tile, G0 += delta1^T x: mma.sync again, A = delta1 (registers, reused by every
//            feature tile), B = the resident X tile; written to this block's slice of a partial-gradient buffer
// then mlp3_reduce_update sums the blocks' partial gradients and applies the momentum update (cublasNN.cpp:619-627) in the
// same pass — no atomics, no gradient arena round trip.
// fp32-grade mode: every MMA operand is split in REGISTERS into its tf32 head and the exact remainder and each product is
// issued as hi*hi + hi*lo + lo*hi (the dropped lo*lo term is 2^-22 relative): no low-part arrays in memory at all.
//
// Shared memory.  X tile: row = feature, 128 B (32 samples) per row, no padding; the 16-byte slot index is XOR-swizzled
// with the row (2 ((row >> 2) & 3) ^ 4 (row & 1)) so that BOTH fragment patterns are conflict-free 16-byte loads: phase 1
// (lanes walk 4 feature rows x 8 slots) and phase 3 (8 feature rows x 4 slots).  The MMA's k index and the sample index of
// its rows are PERMUTATIONS of features / samples chosen so that every lane's fragment values are 4 consecutive floats:
//   phase 1  k-slot t   of k-step j  <->  feature kb + 4t + j,  k-slot t+4 <-> feature kb + 16 + 4t + j   (A and B alike)
//            row g / g+8 of row tile mt  <->  samples 4g + 2mt, 4g + 2mt + 1
//   phase 3  k-slot t / t+4 of k-step j  <->  samples 4t + j / 16 + 4t + j
// W0 travels through a ring of stages of 128 features (4 chunks of 32, one per K group of warps), loaded with cp.async
// together with the X rows of the same features — the first MMAs start when the first 128 feature rows have landed, and
// the L2 latency of W0 (every block reads all of it) is paid once, not once per chunk.  Blocks start at different rounds
// so that 132 blocks do not ask L2 for the same lines at the same moment.
#include "engine.h"
#include "pdl.cuh"
#include "epilogue.cuh"
#include "../../include/b2nn.h"

#include <cstdlib>

namespace b2 {
namespace {

constexpr int M3_THREADS = 256;
constexpr int M3_WARPS = M3_THREADS / 32;
constexpr int M3_S = 32;                  // samples per tile
constexpr int M3_KG = 4;                  // K groups of warps (chunks of a round processed concurrently)
constexpr int M3_NG = 2;                  // column groups of warps
constexpr int M3_RF = 32 * M3_KG;         // features per round
constexpr int M3_AP = 36;                 // pitch of the small activation arrays (a1s)
constexpr uint32_t TF32_MASK = 0xFFFFE000u;

// float index of (row, 16-byte slot) in a 32-float-per-row array with the X-tile swizzle
__device__ __forceinline__ int xs_slot(int row, int slot) { return row * 32 + ((slot ^ ((((row >> 2) & 3) << 1) ^ ((row & 1) << 2))) << 2); }
// float index of (unit row, slot) in a W0 stage (128 floats per row) / in d1s (32 floats per row): slot ^ 4 on odd rows
__device__ __forceinline__ int ws_slot(int row, int slot) { return row * M3_RF + ((slot ^ ((row & 1) << 2)) << 2); }
__device__ __forceinline__ int ds_slot(int row, int slot) { return row * 32 + ((slot ^ ((row & 1) << 2)) << 2); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// D(16x8) += A(16x8) B(8x8), tf32 inputs (fp32 bit patterns), fp32 accumulate.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hi_of(float x) { return __float_as_uint(x) & TF32_MASK; }
__device__ __forceinline__ uint32_t lo_of(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & TF32_MASK)); }

// One product tile.  Plain tf32: the fp32 bit patterns go to the tensor core as they are (mma.sync .tf32 reads the upper
// 19 bits, i.e. truncates like tcgen05 kind::tf32).  fp32-grade: operands arrive pre-split into head and exact remainder
// and the product is issued as hi*hi + hi*lo + lo*hi.
struct Frag4 { uint32_t hi[4], lo[4]; };
template <bool X3>
__device__ __forceinline__ Frag4 split4(float a0, float a1, float a2, float a3) {
    Frag4 f;
    const float v[4] = {a0, a1, a2, a3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (X3) { f.hi[i] = hi_of(v[i]); f.lo[i] = lo_of(v[i]); }
        else    { f.hi[i] = __float_as_uint(v[i]); f.lo[i] = 0u; }
    }
    return f;
}
template <bool X3>
__device__ __forceinline__ void mma_prod(float (&d)[4], const Frag4& a, float b0, float b1) {
    if (X3) {
        const uint32_t bh0 = hi_of(b0), bh1 = hi_of(b1);
        mma_tf32(d, a.hi, bh0, bh1);
        mma_tf32(d, a.hi, lo_of(b0), lo_of(b1));
        mma_tf32(d, a.lo, bh0, bh1);
    } else {
        mma_tf32(d, a.hi, __float_as_uint(b0), __float_as_uint(b1));
    }
}
__device__ __forceinline__ float f4(const float4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

__device__ __forceinline__ float act_fwd3(int kind, float z) { return kind == B2NN_ACT_SIGMOID ? sigmoid_f(z) : tanhf(z); }
__device__ __forceinline__ float act_grad3(int kind, float a) { return kind == B2NN_ACT_SIGMOID ? a * (1.0f - a) : 1.0f - a * a; }

__host__ __device__ constexpr int m3_stages(int nt) { return nt <= 7 ? 3 : 2; }

// NT = 8-unit column tiles of the hidden layer (H <= 8 NT), X3 = fp32-grade arithmetic, TRAIN = backward + gradients.
template <int NT, bool X3, bool TRAIN>
__global__ void __launch_bounds__(M3_THREADS, 1) mlp3_kernel(const Mlp3Args p) {
    constexpr int NTG = (NT + M3_NG - 1) / M3_NG;    // column tiles per column group of warps
    constexpr int MT = (NT + 1) / 2;                 // 16-unit row tiles of the hidden layer in phase 3
    constexpr int NSTG = m3_stages(NT);
    constexpr int WROWS = 8 * NT;
    constexpr int STAGE = WROWS * M3_RF;             // floats per W0 stage
    constexpr int RED = M3_NG * 2 * NTG * 4 * 32;    // floats one K group contributes to the phase-1 reduction
    static_assert(M3_KG * RED <= NSTG * STAGE, "the reduction buffer aliases the W0 ring");
    extern __shared__ __align__(16) float m3_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int kg = warp & (M3_KG - 1), ng = warp / M3_KG;
    const int n1 = p.n1, H = p.H, K = p.K;
    const int xrows = (n1 + 31) / 32 * 32;           // rows n1 .. xrows-1 stay zero
    const int rounds = (n1 + M3_RF - 1) / M3_RF;
    float* Xs = m3_smem;                                   // [xrows][32] swizzled
    float* ring = Xs + size_t(xrows) * 32;                 // [NSTG][WROWS][128] swizzled; after phase 1: the K-group partial sums
    float* a1s = ring + NSTG * STAGE;                      // [H+1][36] hidden activations, ones row last
    float* d1s = a1s + (WROWS + 1) * M3_AP;                // [16 MT][32] swizzled: delta of the hidden layer (rows >= H zero)
    float* zs = d1s + 16 * MT * 32;                        // [K][32] z, then exp(z - max)
    float* d2s = zs + 16 * M3_S;                           // [K][32] output delta
    float* ys = d2s + 16 * M3_S;                           // [K][32] targets of the tile
    float* w1s = ys + 16 * M3_S;                           // [K][H+1]

    // ---- once per block: constants in shared memory (nothing here reads what the previous kernel wrote) ----
    for (int i = tid; i < (xrows - n1) * 32; i += M3_THREADS) Xs[size_t(n1) * 32 + i] = 0.f;
    for (int i = tid; i < NSTG * (WROWS - H) * M3_RF; i += M3_THREADS)        // unit rows >= H are never loaded: zeros
        ring[(i / ((WROWS - H) * M3_RF)) * STAGE + H * M3_RF + i % ((WROWS - H) * M3_RF)] = 0.f;
    for (int i = tid; i < 16 * MT * 32; i += M3_THREADS) d1s[i] = 0.f;
    if (tid < M3_AP) a1s[H * M3_AP + tid] = 1.0f;
    pdl_enter();
    for (int i = tid; i < K * (H + 1); i += M3_THREADS) w1s[i] = __ldg(p.W1 + (i / (H + 1)) * p.ldw1 + i % (H + 1));

    const int64_t n_tiles = (p.rows + M3_S - 1) / M3_S;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) && (p.ldx % 4 == 0);
    const bool yvec_ok = p.Y && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) && (p.ldy % 4 == 0);
    const int r0 = int(blockIdx.x % unsigned(rounds));      // blocks walk the rounds from different starting points
    double cost = 0.0, errs = 0.0;
    bool first_tile = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first_tile = false) {
        const int64_t s0 = tile * M3_S;
        const int valid = int(p.rows - s0 < M3_S ? p.rows - s0 : M3_S);
        const bool full = valid == M3_S;
        __syncthreads();                                    // the previous tile is done with Xs / ring / d1s / ys
        if (!first_tile) {                                  // the ring held the K-group sums: rows >= H must read as zeros again
            for (int i = tid; i < NSTG * (WROWS - H) * M3_RF; i += M3_THREADS)
                ring[(i / ((WROWS - H) * M3_RF)) * STAGE + H * M3_RF + i % ((WROWS - H) * M3_RF)] = 0.f;
            __syncthreads();
        }

        // One load group = the X rows and the W0 columns of one round of 128 features (+ the targets with the first).
        auto issue_round = [&](int i) {
            const int rd = (r0 + i) % rounds, f0 = rd * M3_RF;
            const int frows = min(M3_RF, n1 - f0);
            if (vec_ok && full) {
                for (int e = tid; e < frows * 8; e += M3_THREADS) {
                    const int r = f0 + (e >> 3), sl = e & 7;
                    cp_async16(Xs + xs_slot(r, sl), p.X + int64_t(r) * p.ldx + s0 + 4 * sl);
                }
            } else {
                for (int e = tid; e < frows * M3_S; e += M3_THREADS) {
                    const int r = f0 + (e >> 5), c = e & 31;
                    float* dst = Xs + xs_slot(r, c >> 2) + (c & 3);
                    if (c < valid) cp_async4(dst, p.X + int64_t(r) * p.ldx + s0 + c); else *dst = 0.f;
                }
            }
            float* stage = ring + (i % NSTG) * STAGE;
            for (int e = tid; e < H * 32; e += M3_THREADS) {
                const int u = e >> 5, sl = e & 31, f = f0 + 4 * sl;
                float* dst = stage + ws_slot(u, sl);
                if (f + 3 < p.ldw0) cp_async16(dst, p.W0 + int64_t(u) * p.ldw0 + f);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (i == 0 && p.Y) {
                if (yvec_ok && full) {
                    for (int e = tid; e < K * 8; e += M3_THREADS) cp_async16(ys + (e >> 3) * M3_S + 4 * (e & 7), p.Y + int64_t(e >> 3) * p.ldy + s0 + 4 * (e & 7));
                } else {
                    for (int e = tid; e < K * M3_S; e += M3_THREADS) {
                        const int c = e & 31;
                        if (c < valid) cp_async4(ys + e, p.Y + int64_t(e >> 5) * p.ldy + s0 + c); else ys[e] = 0.f;
                    }
                }
            }
            cp_async_commit();
        };

        // ---- phase 1: hidden pre-activations.  Warp (kg, ng): chunk kg of every round, column tiles ng*NTG .. ----
        {
            float acc[2][NTG][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTG; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
#pragma unroll 1
            for (int i = 0; i < NSTG - 1; ++i) { if (i < rounds) issue_round(i); else cp_async_commit(); }
#pragma unroll 1
            for (int i = 0; i < rounds; ++i) {
                cp_async_wait<NSTG - 2>();                   // this thread's copies of round i have landed ...
                __syncthreads();                             // ... everyone's have, and round i-1's stage is free again
                if (i + NSTG - 1 < rounds) issue_round(i + NSTG - 1); else cp_async_commit();
                const int rd = (r0 + i) % rounds, kb = rd * M3_RF + kg * 32;
                if (kb < n1) {                               // warp-uniform
                    const float* stage = ring + (i % NSTG) * STAGE;
                    Frag4 fa[4][2];                           // A fragments of the four k-steps, both row tiles
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 x1 = *reinterpret_cast<const float4*>(Xs + xs_slot(kb + 4 * t + j, g));
                        const float4 x2 = *reinterpret_cast<const float4*>(Xs + xs_slot(kb + 16 + 4 * t + j, g));
                        fa[j][0] = split4<X3>(x1.x, x1.y, x2.x, x2.y);      // samples 4g, 4g+1
                        fa[j][1] = split4<X3>(x1.z, x1.w, x2.z, x2.w);      // samples 4g+2, 4g+3
                    }
#pragma unroll
                    for (int nt = 0; nt < NTG; ++nt) {
                        if (ng * NTG + nt < NT) {
                            const int u = (ng * NTG + nt) * 8 + g;
                            const float4 w1v = *reinterpret_cast<const float4*>(stage + ws_slot(u, kg * 8 + t));
                            const float4 w2v = *reinterpret_cast<const float4*>(stage + ws_slot(u, kg * 8 + 4 + t));
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                mma_prod<X3>(acc[0][nt], fa[j][0], f4(w1v, j), f4(w2v, j));
                                mma_prod<X3>(acc[1][nt], fa[j][1], f4(w1v, j), f4(w2v, j));
                            }
                        }
                    }
                }
            }
            cp_async_wait<0>();
            __syncthreads();                                 // every warp is done with the ring: it becomes the reduction buffer
            float* mine = ring + kg * RED + ng * (2 * NTG * 4 * 32);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NTG; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) mine[((mt * NTG + nt) * 4 + q) * 32 + lane] = acc[mt][nt][q];
        }
        __syncthreads();
        for (int o = tid; o < RED; o += M3_THREADS) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < M3_KG; ++w) sum += ring[w * RED + o];
            const int l = o & 31, q = (o >> 5) & 3, tl = (o >> 7) % (2 * NTG), grp = (o >> 7) / (2 * NTG);
            const int mt = tl / NTG, nt = grp * NTG + tl % NTG;
            const int sample = 4 * (l >> 2) + 2 * mt + (q >> 1), unit = nt * 8 + 2 * (l & 3) + (q & 1);
            if (unit < H) a1s[unit * M3_AP + sample] = act_fwd3(p.act_kind, sum);
        }
        __syncthreads();

        // ---- phase 2: output layer (fp32 FFMA), output activation, delta, cost, argmax ----
        for (int i = tid; i < K * M3_S; i += M3_THREADS) {
            const int k = i >> 5, s = i & 31;
            const float* w = w1s + k * (H + 1);
            float z = w[H];
            for (int u = 0; u < H; ++u) z = fmaf(w[u], a1s[u * M3_AP + s], z);
            zs[k * M3_S + s] = z;
        }
        __syncthreads();
        if (warp == 0) {
            const int s = lane;
            const bool ok = s < valid;
            const int64_t i = s0 + (ok ? s : 0);
            float zmax = -INFINITY, total = 1.f;
            if (p.out_kind == B2NN_OUT_SOFTMAX) {
                for (int k = 0; k < K; ++k) zmax = fmaxf(zmax, zs[k * M3_S + s]);
                total = 0.f;
                for (int k = 0; k < K; ++k) { const float e = expf(zs[k * M3_S + s] - zmax); zs[k * M3_S + s] = e; total += e; }
            }
            float best_h = 0.f, best_y = 0.f; int arg_h = 0, arg_y = 0;
            for (int k = 0; k < K; ++k) {
                const float zv = zs[k * M3_S + s];
                float hv;
                if (p.out_kind == B2NN_OUT_SOFTMAX) hv = zv / total;
                else if (p.out_kind == B2NN_OUT_SIGMOID) hv = sigmoid_f(zv);
                else hv = zv;
                if (hv > best_h) { best_h = hv; arg_h = k; }
                if (ok && p.h_out) p.h_out[int64_t(k) * p.ldh + i] = hv;
                float dv = 0.f;
                if (p.Y) {
                    const float yv = ys[k * M3_S + s];
                    if (yv > best_y) { best_y = yv; arg_y = k; }
                    dv = hv - yv;
                    if (ok && p.cost_sum) {
                        float c = 0.f;
                        if (p.cost_kind == B2NN_COST_SQUARED) c = dv * dv;
                        else {
                            if (yv != 0.f) c += -yv * logf(hv);
                            if (p.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) c += -(1.f - yv) * logf(1.f - hv);
                            c = fabsf(c);
                        }
                        cost += double(c);
                    }
                }
                d2s[k * M3_S + s] = ok ? dv : 0.f;            // samples past the batch end contribute nothing
            }
            if (ok && p.pred) p.pred[i] = float(arg_h);
            if (ok && p.err_sum) {
                if (p.y_labels) errs += (float(arg_h) != __ldg(p.y_labels + i)) ? 1.0 : 0.0;
                else if (p.Y)   errs += (arg_h != arg_y) ? 1.0 : 0.0;
            }
        }
        if (!TRAIN) continue;
        __syncthreads();
        float* part = p.partial + int64_t(blockIdx.x) * p.total;
        // delta of the hidden layer  d1 = (d2 W1) .* f'(a1)   (cublasNN.cpp:673-676), and the output-layer gradient
        for (int i = tid; i < H * M3_S; i += M3_THREADS) {
            const int u = i >> 5, s = i & 31;
            float v = 0.f;
            for (int k = 0; k < K; ++k) v = fmaf(d2s[k * M3_S + s], w1s[k * (H + 1) + u], v);
            d1s[ds_slot(u, s >> 2) + (s & 3)] = v * act_grad3(p.act_kind, a1s[u * M3_AP + s]);
        }
        for (int i = tid; i < K * p.ldw1; i += M3_THREADS) {       // pad columns (u > H) of the slice are written too: zeros
            const int k = i / p.ldw1, u = i - k * p.ldw1;
            float gsum = 0.f;
            if (u <= H)
                for (int s = 0; s < M3_S; ++s) gsum = fmaf(d2s[k * M3_S + s], a1s[u * M3_AP + s], gsum);
            float* dst = part + p.w1_off + i;
            *dst = first_tile ? gsum : *dst + gsum;
        }
        __syncthreads();

        // ---- phase 3: input-layer gradient of this tile, G0[unit][feature] = sum_s d1[unit][s] x[feature][s] ----
        {
            float4 af[MT][4];                                // A = d1 rows g, g+8 of every row tile: slots t and 4+t
            float* prow[MT][2];                              // this lane's output rows in the block's gradient slice
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                af[mt][0] = *reinterpret_cast<const float4*>(d1s + ds_slot(mt * 16 + g, t));
                af[mt][1] = *reinterpret_cast<const float4*>(d1s + ds_slot(mt * 16 + g + 8, t));
                af[mt][2] = *reinterpret_cast<const float4*>(d1s + ds_slot(mt * 16 + g, 4 + t));
                af[mt][3] = *reinterpret_cast<const float4*>(d1s + ds_slot(mt * 16 + g + 8, 4 + t));
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int u = mt * 16 + g + 8 * hrow;
                    prow[mt][hrow] = u < H ? part + int64_t(u) * p.ldw0 + 2 * t : nullptr;
                }
            }
            const int f_tiles = (n1 + 7) / 8;
            for (int ft = warp; ft < f_tiles; ft += M3_WARPS) {
                float acc[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mt][q] = 0.f;
                const float4 b1v = *reinterpret_cast<const float4*>(Xs + xs_slot(ft * 8 + g, t));
                const float4 b2v = *reinterpret_cast<const float4*>(Xs + xs_slot(ft * 8 + g, 4 + t));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        const Frag4 a = split4<X3>(f4(af[mt][0], j), f4(af[mt][1], j), f4(af[mt][2], j), f4(af[mt][3], j));
                        mma_prod<X3>(acc[mt], a, f4(b1v, j), f4(b2v, j));
                    }
                }
                // this lane's two features: ft*8 + 2t, + 1 (even, ldw0 % 4 == 0: an aligned pair inside the row)
                if (ft * 8 + 2 * t < p.ldw0) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            if (prow[mt][hrow]) {
                                float2* dst = reinterpret_cast<float2*>(prow[mt][hrow] + ft * 8);
                                float2 v = make_float2(acc[mt][2 * hrow], acc[mt][2 * hrow + 1]);
                                if (!first_tile) { const float2 o = *dst; v.x += o.x; v.y += o.y; }
                                *dst = v;
                            }
                        }
                }
            }
        }
    }
    if (warp == 0) {
        for (int o = 16; o > 0; o >>= 1) { cost += __shfl_xor_sync(0xffffffffu, cost, o); errs += __shfl_xor_sync(0xffffffffu, errs, o); }
        if (lane == 0) {
            if (p.cost_sum && cost != 0.0) atomicAdd(p.cost_sum, cost);
            if (p.err_sum && errs != 0.0) atomicAdd(p.err_sum, errs);
        }
    }
}

// g = sum over the blocks' partial gradients; v <- mu v + alpha g; theta <- theta + v   (cublasNN.cpp:619-627).
// Block = 64 float4 columns x 4 groups of partials; the groups meet in shared memory.
__global__ void __launch_bounds__(256) mlp3_reduce_update_kernel(const float4* __restrict__ partial, int nparts, int64_t total4,
                                                                 float4* __restrict__ theta, float4* __restrict__ vel, float mu,
                                                                 float alpha) {
    pdl_enter();
    __shared__ float4 sh[4][64];
    const int x = threadIdx.x & 63, y = threadIdx.x >> 6;
    const int64_t i = int64_t(blockIdx.x) * 64 + x;
    float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total4) {
#pragma unroll 4
        for (int q = y; q < nparts; q += 4) {
            const float4 v = __ldcs(partial + int64_t(q) * total4 + i);
            gsum.x += v.x; gsum.y += v.y; gsum.z += v.z; gsum.w += v.w;
        }
    }
    sh[y][x] = gsum;
    __syncthreads();
    if (y == 0 && i < total4) {
        float4 gq = sh[0][x];
#pragma unroll
        for (int q = 1; q < 4; ++q) { gq.x += sh[q][x].x; gq.y += sh[q][x].y; gq.z += sh[q][x].z; gq.w += sh[q][x].w; }
        float4 v = vel[i], th = theta[i];
        v.x = fmaf(mu, v.x, alpha * gq.x); v.y = fmaf(mu, v.y, alpha * gq.y);
        v.z = fmaf(mu, v.z, alpha * gq.z); v.w = fmaf(mu, v.w, alpha * gq.w);
        th.x += v.x; th.y += v.y; th.z += v.z; th.w += v.w;
        vel[i] = v;
        theta[i] = th;
    }
}

size_t mlp3_smem_bytes(int n1, int nt) {
    const int mt = (nt + 1) / 2, xrows = (n1 + 31) / 32 * 32;
    const size_t floats = size_t(xrows) * 32 + size_t(m3_stages(nt)) * (8 * nt) * M3_RF + size_t(8 * nt + 1) * M3_AP +
                          size_t(16 * mt) * 32 + 3 * 16 * M3_S + 16 * (8 * nt + 1);
    return floats * sizeof(float);
}

int pick_nt(int H) { const int need = (H + 7) / 8; return need <= 2 ? 2 : need <= 4 ? 4 : need <= 7 ? 7 : 8; }

template <int NT, bool X3, bool TRAIN>
cudaError_t launch_inst(const Mlp3Args& a, unsigned grid, cudaStream_t s) {
    auto kern = mlp3_kernel<NT, X3, TRAIN>;
    static std::atomic<unsigned long long> done{0};
    const size_t smem = mlp3_smem_bytes(a.n1, NT);
    if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024, done)) return e;
    return launch_pdl(kern, dim3(grid), dim3(M3_THREADS), smem, s, a);
}
template <int NT>
cudaError_t launch_nt(const Mlp3Args& a, unsigned grid, cudaStream_t s) {
    if (a.x3) return a.train ? launch_inst<NT, true, true>(a, grid, s) : launch_inst<NT, true, false>(a, grid, s);
    return a.train ? launch_inst<NT, false, true>(a, grid, s) : launch_inst<NT, false, false>(a, grid, s);
}

}  // namespace

bool mlp3_supported(int n1, int H, int K) {
    if (std::getenv("B2NN_NO_MLP3")) return false;
    if (H < 1 || H > 64 || K < 1 || K > 16 || n1 < 2) return false;
    return mlp3_smem_bytes(n1, pick_nt(H)) <= size_t(227 * 1024);
}

int mlp3_grid(int64_t rows) {
    const int64_t tiles = (rows + M3_S - 1) / M3_S;
    return int(tiles < 148 ? tiles : 148);
}

cudaError_t mlp3_launch(const Mlp3Args& a, cudaStream_t s) {
    if (a.rows <= 0) return cudaSuccess;
    const unsigned grid = unsigned(mlp3_grid(a.rows));
    switch (pick_nt(a.H)) {
        case 2: return launch_nt<2>(a, grid, s);
        case 4: return launch_nt<4>(a, grid, s);
        case 7: return launch_nt<7>(a, grid, s);
        default: return launch_nt<8>(a, grid, s);
    }
}

cudaError_t mlp3_reduce_update_launch(const float* partial, int nparts, int64_t total, float* theta, float* vel, float mu, float alpha,
                                      cudaStream_t s) {
    if (total <= 0 || (total & 3)) return cudaErrorInvalidValue;
    const int64_t total4 = total / 4;
    return launch_pdl(mlp3_reduce_update_kernel, dim3(unsigned((total4 + 63) / 64)), dim3(256), 0, s,
                      reinterpret_cast<const float4*>(partial), nparts, total4, reinterpret_cast<float4*>(theta),
                      reinterpret_cast<float4*>(vel), mu, alpha);
}

}  // namespace b2
