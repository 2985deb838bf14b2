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
// Shared-memory X tile: row = feature, 36-float pitch, column (sample) index XOR 8 on odd 8-row groups.  Phase 1 reads it
// "transposed" (lanes walk 4 features x 8 samples), phase 3 row-wise (8 features x 4 samples); with this pitch and swizzle
// both fragment loads touch 32 distinct banks.
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
constexpr int M3_XP = 36;                 // floats per shared-memory row
constexpr uint32_t TF32_MASK = 0xFFFFE000u;

__device__ __forceinline__ int xs_index(int row, int col) { return row * M3_XP + (col ^ (((row >> 3) & 1) << 3)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// D(16x8) += A(16x8) B(8x8), tf32 inputs (fp32 bit patterns), fp32 accumulate.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hi_of(float x) { return __float_as_uint(x) & TF32_MASK; }
__device__ __forceinline__ uint32_t lo_of(float x) { return __float_as_uint(x - __uint_as_float(__float_as_uint(x) & TF32_MASK)); }

// One product tile: plain tf32 (operands truncated like tcgen05 kind::tf32 does) or fp32-grade (three MMAs).
template <bool X3>
__device__ __forceinline__ void mma_prod(float (&d)[4], const float (&a)[4], float b0, float b1) {
    uint32_t ah[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ah[i] = hi_of(a[i]);
    const uint32_t bh0 = hi_of(b0), bh1 = hi_of(b1);
    mma_tf32(d, ah, bh0, bh1);
    if (X3) {
        uint32_t al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) al[i] = lo_of(a[i]);
        mma_tf32(d, ah, lo_of(b0), lo_of(b1));
        mma_tf32(d, al, bh0, bh1);
    }
}

__device__ __forceinline__ float act_fwd3(int kind, float z) { return kind == B2NN_ACT_SIGMOID ? sigmoid_f(z) : tanhf(z); }
__device__ __forceinline__ float act_grad3(int kind, float a) { return kind == B2NN_ACT_SIGMOID ? a * (1.0f - a) : 1.0f - a * a; }

// NT = 8-unit column tiles of the hidden layer (H <= 8 NT), X3 = fp32-grade arithmetic, TRAIN = backward + gradients.
template <int NT, bool X3, bool TRAIN>
__global__ void __launch_bounds__(M3_THREADS, 1) mlp3_kernel(const Mlp3Args p) {
    constexpr int MT = (NT + 1) / 2;             // 16-unit row tiles of the hidden layer in phase 3
    constexpr int RED = 2 * NT * 4 * 32;         // floats one warp contributes to the phase-1 reduction
    extern __shared__ __align__(16) float m3_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int n1 = p.n1, H = p.H, K = p.K;
    const int xrows = (n1 + 31) / 32 * 32;       // rows n1 .. xrows-1 stay zero: the feature loop runs in blocks of 32
    float* Xs = m3_smem;                                   // [xrows][36]
    float* red = Xs + size_t(xrows) * M3_XP;               // [8 warps][RED]
    float* a1s = red + M3_WARPS * RED;                     // [H+1][36] hidden activations, ones row last
    float* d1s = a1s + (8 * NT + 1) * M3_XP;               // [16 MT][36] delta of the hidden layer (rows >= H zero)
    float* zs = d1s + 16 * MT * M3_XP;                     // [K][32] z, then h
    float* d2s = zs + 16 * M3_S;                           // [K][32] output delta
    float* w1s = d2s + 16 * M3_S;                          // [K][H+1]

    // ---- once per block: constants in shared memory (nothing here reads what the previous kernel wrote) ----
    for (int i = tid; i < (xrows - n1) * M3_XP; i += M3_THREADS) Xs[size_t(n1) * M3_XP + i] = 0.f;
    for (int i = tid; i < 16 * MT * M3_XP; i += M3_THREADS) d1s[i] = 0.f;
    if (tid < M3_XP) a1s[H * M3_XP + tid] = 1.0f;
    pdl_enter();
    for (int i = tid; i < K * (H + 1); i += M3_THREADS) w1s[i] = __ldg(p.W1 + (i / (H + 1)) * p.ldw1 + i % (H + 1));

    const int64_t n_tiles = (p.rows + M3_S - 1) / M3_S;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) && (p.ldx % 4 == 0);
    double cost = 0.0, errs = 0.0;
    bool first_tile = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, first_tile = false) {
        const int64_t s0 = tile * M3_S;
        const int valid = int(p.rows - s0 < M3_S ? p.rows - s0 : M3_S);
        __syncthreads();                                    // the previous tile's phase 3 is done with Xs / d1s
        // ---- X tile -> shared memory (features x samples, zero beyond the batch end) ----
        if (vec_ok && valid == M3_S) {
            for (int i = tid; i < n1 * 8; i += M3_THREADS) {
                const int r = i >> 3, c4 = (i & 7) << 2;
                cp_async16(Xs + xs_index(r, c4), p.X + int64_t(r) * p.ldx + s0 + c4);
            }
        } else {
            for (int i = tid; i < n1 * M3_S; i += M3_THREADS) {
                const int r = i >> 5, c = i & 31;
                if (c < valid) cp_async4(Xs + xs_index(r, c), p.X + int64_t(r) * p.ldx + s0 + c);
                else Xs[xs_index(r, c)] = 0.f;
            }
        }
        cp_async_wait_all();
        __syncthreads();

        // ---- phase 1: hidden pre-activations, split-K over the warps (blocks of 32 features) ----
        {
            float acc[2][NT][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
            for (int kb = warp * 32; kb < n1; kb += M3_WARPS * 32) {
                // B fragments: lane (g, t) holds W0[unit 8nt+g][kb + 4t + j] and [kb + 16 + 4t + j], j = 0..3 — the k index
                // of the MMA is a permutation of the features, the same one on the A side
                float4 w_lo4[NT], w_hi4[NT];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int u = nt * 8 + g;
                    const float* wr = p.W0 + int64_t(u) * p.ldw0 + kb + 4 * t;
                    const bool row_ok = u < H;
                    w_lo4[nt] = (row_ok && kb + 4 * t + 3 < p.ldw0) ? __ldg(reinterpret_cast<const float4*>(wr)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    w_hi4[nt] = (row_ok && kb + 16 + 4 * t + 3 < p.ldw0) ? __ldg(reinterpret_cast<const float4*>(wr + 16)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a[2][4];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        a[mt][0] = Xs[xs_index(kb + 4 * t + j, mt * 16 + g)];
                        a[mt][1] = Xs[xs_index(kb + 4 * t + j, mt * 16 + g + 8)];
                        a[mt][2] = Xs[xs_index(kb + 16 + 4 * t + j, mt * 16 + g)];
                        a[mt][3] = Xs[xs_index(kb + 16 + 4 * t + j, mt * 16 + g + 8)];
                    }
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const float b0 = j == 0 ? w_lo4[nt].x : j == 1 ? w_lo4[nt].y : j == 2 ? w_lo4[nt].z : w_lo4[nt].w;
                        const float b1 = j == 0 ? w_hi4[nt].x : j == 1 ? w_hi4[nt].y : j == 2 ? w_hi4[nt].z : w_hi4[nt].w;
                        mma_prod<X3>(acc[0][nt], a[0], b0, b1);
                        mma_prod<X3>(acc[1][nt], a[1], b0, b1);
                    }
                }
            }
            float* mine = red + warp * RED;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) mine[((mt * NT + nt) * 4 + q) * 32 + lane] = acc[mt][nt][q];
        }
        __syncthreads();
        for (int o = tid; o < RED; o += M3_THREADS) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < M3_WARPS; ++w) sum += red[w * RED + o];
            const int l = o & 31, q = (o >> 5) & 3, tl = o >> 7, mt = tl / NT, nt = tl - mt * NT;
            const int sample = mt * 16 + (l >> 2) + 8 * (q >> 1), unit = nt * 8 + 2 * (l & 3) + (q & 1);
            if (unit < H) a1s[unit * M3_XP + sample] = act_fwd3(p.act_kind, sum);
        }
        __syncthreads();

        // ---- phase 2: output layer (fp32 FFMA), output activation, delta, cost, argmax ----
        for (int i = tid; i < K * M3_S; i += M3_THREADS) {
            const int k = i >> 5, s = i & 31;
            const float* w = w1s + k * (H + 1);
            float z = w[H];
            for (int u = 0; u < H; ++u) z = fmaf(w[u], a1s[u * M3_XP + s], z);
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
                    const float yv = __ldg(p.Y + int64_t(k) * p.ldy + i);
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
            d1s[u * M3_XP + s] = v * act_grad3(p.act_kind, a1s[u * M3_XP + s]);
        }
        for (int i = tid; i < K * p.ldw1; i += M3_THREADS) {       // pad columns (u > H) of the slice are written too: zeros
            const int k = i / p.ldw1, u = i - k * p.ldw1;
            float gsum = 0.f;
            if (u <= H)
                for (int s = 0; s < M3_S; ++s) gsum = fmaf(d2s[k * M3_S + s], a1s[u * M3_XP + s], gsum);
            float* dst = part + p.w1_off + i;
            *dst = first_tile ? gsum : *dst + gsum;
        }
        __syncthreads();

        // ---- phase 3: input-layer gradient of this tile, G0[unit][feature] = sum_s d1[unit][s] x[feature][s] ----
        {
            float af[4][MT][4];                              // A = d1: every feature tile reuses it
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    af[ks][mt][0] = d1s[(mt * 16 + g) * M3_XP + ks * 8 + t];
                    af[ks][mt][1] = d1s[(mt * 16 + g + 8) * M3_XP + ks * 8 + t];
                    af[ks][mt][2] = d1s[(mt * 16 + g) * M3_XP + ks * 8 + t + 4];
                    af[ks][mt][3] = d1s[(mt * 16 + g + 8) * M3_XP + ks * 8 + t + 4];
                }
            const int f_tiles = (n1 + 7) / 8;
            for (int ft = warp; ft < f_tiles; ft += M3_WARPS) {
                float acc[MT][4];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[mt][q] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const float b0 = Xs[xs_index(ft * 8 + g, ks * 8 + t)];
                    const float b1 = Xs[xs_index(ft * 8 + g, ks * 8 + t + 4)];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) mma_prod<X3>(acc[mt], af[ks][mt], b0, b1);
                }
                const int f = ft * 8 + 2 * t;                // this lane's two features: f, f + 1 (f even, ldw0 % 4 == 0)
                if (f < p.ldw0) {
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                        for (int hrow = 0; hrow < 2; ++hrow) {
                            const int u = mt * 16 + g + 8 * hrow;
                            if (u < H) {
                                float2* dst = reinterpret_cast<float2*>(part + int64_t(u) * p.ldw0 + f);
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
    const size_t floats = size_t(xrows) * M3_XP + size_t(M3_WARPS) * (2 * nt * 4 * 32) + size_t(8 * nt + 1) * M3_XP +
                          size_t(16 * mt) * M3_XP + 2 * 16 * M3_S + 16 * (8 * nt + 1);
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
