// gemm_tc.cu — tcgen05 / TMEM / TMA contraction for sm_100a, TF32 inputs with FP32 accumulation.
//
//   D(m,n) = sum_k A(m,k) B(n,k),   D stored m-contiguous (D[n*ldd + m]).
//
// One kernel template covers the three contractions of the train step; they differ only in which operand index is
// contiguous in memory (engine layout: every matrix is "unit-major", i.e. [unit][sample] or [out][in]):
//   forward   z = a W^T    : A = a_{j-1}  MN-major ([k = unit][m = sample]),  B = W_j      K-major ([n = out][k = in])
//   backprop  d = dn W     : A = delta    MN-major ([k = out ][m = sample]),  B = W_{j+1}  MN-major ([k = out][n = in])
//   wgrad     G = d^T a    : A = a_{j-1}  K-major  ([m = in  ][k = sample]),  B = delta_j  K-major  ([n = out][k = sample])
// replacing the three cublasSgemm op combinations of the reference (cublasNN.cpp:645-659 NN, :674-675 NT, :680-684 TN).
//
// CTA = 18 warps (persistent, one CTA per SM, normally paired into clusters of two that share a 256 x BN tile), K
// streamed in blocks of 32 floats (= one 128-byte swizzle row):
//   warp 0       TMA producer: cp.async.bulk.tensor into a STAGES-deep shared-memory ring, mbarrier complete_tx
//   warp 1       TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (128 or 256 x BN x 8),
//                tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2..17  epilogue: tcgen05.ld 32 lanes x 16 columns -> registers, fused activation / derivative, coalesced
//                stores (lane = sample index, so each warp store is one 128-byte line); two TMEM accumulator buffers
// Tails in every dimension are handled by TMA zero fill on the load side and predication on the store side.
// Shared-memory tiles use the 128-byte swizzle; the UMMA shared-memory descriptors are "stage 0 + uniform offset":
//   K-major  tile: rows of 128 B, 8-row groups 1024 B apart (SBO); K step of 8 floats = +32 B on the start address
//   MN-major tile: 32-float x 32-k boxes of 4096 B written with the 32-byte-chunk 128 B swizzle (the only one
//                  tcgen05 accepts for MN-major 4-byte operands); atoms along M/N are 4096 B apart (LBO), 4-k groups
//                  512 B apart (SBO); K step of 8 = +1024 B on the start address
#include "engine.h"
#include "epilogue.cuh"
#include "ptx.cuh"
#include "pdl.cuh"
#include "../../include/b2nn.h"

#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace b2 {
namespace {

constexpr int BM = 128;                // rows of the output tile owned by ONE CTA (TMEM lanes)
constexpr int BK = 32;                 // floats per K block = 128 bytes
constexpr int UMMA_K = 8;              // tf32: 32 bytes per instruction
constexpr int EPI_WARPS = 16;          // 4 warps per TMEM lane quarter, each takes a quarter of the tile's columns
constexpr int EPI_PARTS = EPI_WARPS / 4;
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int ATOM_BYTES = 32 * BK * 4;   // one MN-major box: 32 (mn) x 32 (k) floats
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int MAX_STAGES = 8;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 16 * 32 * 4;   // scatter mode: one 16-column x 32-row chunk per epilogue warp

constexpr int ATOM_BYTES_BF = 64 * 64 * 2;   // bf16 mode: one MN-major box is 64 (mn) x 64 (k) two-byte elements

// bytes of the B tile ONE CTA stages: `rows` = N rows this CTA loads (BN, or BN/2 in a CTA pair).  A K block is 128
// bytes of K in every mode (32 floats or 64 bf16), so K-major tiles have the same size in all of them.
__host__ __device__ constexpr int b_tile_bytes(int rows, bool b_mn, bool bf = false) {
    return b_mn ? (bf ? ((rows + 63) / 64) * ATOM_BYTES_BF : ((rows + 31) / 32) * ATOM_BYTES) : rows * BK * 4;
}

struct TcArgs {
    float* D; int64_t ldd;
    float* D_lo;                    // X3 only: x - trunc_tf32(x) of every stored output (may be null)
    const float* aux; int64_t ldaux;
    int M, N;
    int tiles_m, tiles_n, splits;   // work item = (tile_m, tile_n, split); tile_m counts (128 * NCTA)-row tiles
    int group_m;                    // rasterisation group (see decode_tile)
    int serp;                       // 1: odd groups walk the column tiles backwards (see decode_tile)
    int k_blocks;                   // K blocks per split
    int k_blocks_total;
    int a_mn0, a_k0, b_mn0, b_k0;
    int epi;
    int atomic;                     // split-K: accumulate with red.global.add
    int d_trans;                    // store D[m * ldd + n] (row of this thread's TMEM lane, 32 consecutive columns)
    int stages;
    uint32_t mn_lbo, mn_sbo, mn_layout;   // MN-major descriptor: 4096 / 512 / type 1 (tf32), 8192 / 1024 / type 2 (bf16)
    uint32_t mn_kstep;                    // bytes one MMA K step advances an MN-major operand: 1024 (8 rows) / 2048 (16 rows)
    int a_3d, b_3d;                 // MN-major operand fetched with one 3-D TMA box per tile
    float* scatter_base[8];         // fused reduce-scatter: owner arenas (peer mappings); see GemmDesc
    int scatter_world, scatter_rank, scatter_rows;
    int n_rot;                      // scatter mode: column-tile rotation, so every rank starts with a different owner's tiles
    int scatter_bulk;               // scatter mode: tiles leave through bulk asynchronous stores staged in shared memory
    const int* wait_flags;          // data parallel over peer memory: the weights this contraction reads are complete once
    int wait_count, wait_value;     // wait_flags[0 .. wait_count) >= wait_value (one flag per owner rank, written by the owners)
    TcOutput out;                   // fused output layer (out.counters != nullptr): see engine.h
};

// Work item -> output tile.  Tiles are walked in groups of `group_m` row-tiles: inside a group the row-tile index is
// fastest, then the column-tile index, so the ~148 tiles in flight at any time touch about group_m panels of A and
// 148 / group_m panels of B.  group_m = 1 sweeps all column tiles of one row-tile first (B stays L2 resident when it is
// the small operand, e.g. the weights in forward / backprop); large group_m approaches row-fastest order.
// serp: odd groups walk the column tiles backwards, so the B panels one group ended on are the ones the next starts on.
// n_rot (fused reduce-scatter): the column-tile index is rotated by (rank + 1) owners' worth of tiles, so at any moment
// the ranks are storing to DIFFERENT owners (no incast on one GPU's NVLink ingress) and each rank's own tiles — plain
// local stores — come last, where the epilogue of the final wave is no longer hidden behind MMAs.
__device__ __forceinline__ void decode_tile(int rem, int tiles_m, int tiles_n, int group_m, int n_rot, int serp, int& tile_m, int& tile_n) {
    const int per_group = group_m * tiles_n;
    const int g = rem / per_group;
    const int first_m = g * group_m;
    const int gsize = min(group_m, tiles_m - first_m);
    const int in_g = rem - g * per_group;
    tile_m = first_m + in_g % gsize;
    int tn = in_g / gsize;
    if (serp && (g & 1)) tn = tiles_n - 1 - tn;
    tile_n = tn + n_rot;
    if (tile_n >= tiles_n) tile_n -= tiles_n;
}

// Fast forms for the tensor-core epilogue (tf32 mode): MUFU.EX2 + MUFU.RCP, relative error ~1e-6 — three orders below
// the tf32 rounding of the contraction they follow.  The fp32 FFMA path keeps expf / tanhf (epilogue.cuh).
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

// Output layer of one 32-sample slice (lane = sample), called by the epilogue warp that completed the slice's z: output
// activation (max-shifted softmax like output_layer_kernel), delta = h - y, cost summed in double.  Kept out of line: its
// registers must not count against the 576-thread GEMM kernel's budget; it runs once per slice.
__device__ __noinline__ void fused_output_slice(const float* __restrict__ y, int64_t ldy, float* __restrict__ delta, int64_t ldd,
                                                double* cost_sum, int out_kind, int cost_kind, const float* __restrict__ Z,
                                                int64_t ldz, int K, int m, bool m_ok, int lane) {
    struct { const float* y; int64_t ldy; float* delta; int64_t ldd; double* cost_sum; int out_kind, cost_kind; } o{y, ldy, delta, ldd, cost_sum, out_kind, cost_kind};
    double cost = 0.0;
    if (m_ok) {
        float zv[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) zv[k] = k < K ? __ldcg(Z + int64_t(k) * ldz + m) : -INFINITY;
        float zmax = -INFINITY, total = 0.f;
        if (o.out_kind == B2NN_OUT_SOFTMAX) {
#pragma unroll
            for (int k = 0; k < 16; ++k) zmax = fmaxf(zmax, zv[k]);
#pragma unroll
            for (int k = 0; k < 16; ++k) if (k < K) { zv[k] = expf(zv[k] - zmax); total += zv[k]; }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k < K) {
                float h;
                if (o.out_kind == B2NN_OUT_SOFTMAX) h = zv[k] / total;
                else if (o.out_kind == B2NN_OUT_SIGMOID) h = 1.0f / (1.0f + expf(-zv[k]));
                else h = zv[k];
                const float yv = __ldg(o.y + int64_t(k) * o.ldy + m);
                const float dv = h - yv;
                o.delta[int64_t(k) * o.ldd + m] = dv;
                if (o.cost_sum) {
                    float cst = 0.f;
                    if (o.cost_kind == B2NN_COST_SQUARED) cst = dv * dv;
                    else {
                        if (yv != 0.f) cst += -yv * logf(h);
                        if (o.cost_kind == B2NN_COST_NEGLNMAX && yv != 1.f) cst += -(1.f - yv) * logf(1.f - h);
                        cst = fabsf(cst);
                    }
                    cost += double(cst);
                }
            }
        }
    }
    if (o.cost_sum) {
        for (int s = 16; s > 0; s >>= 1) cost += __shfl_xor_sync(0xffffffffu, cost, s);
        if (lane == 0 && cost != 0.0) atomicAdd(o.cost_sum, cost);
    }
}

// Persistent, warp-specialised tcgen05 GEMM.  NCTA = 1: one CTA per 128 x BN tile.  NCTA = 2: a CTA pair (cluster of
// two, cta_group::2) per 256 x BN tile — each CTA stages its own 128 rows of A and HALF of B, so the L2 -> shared
// traffic per multiply-add drops by a third; the leader CTA issues the MMAs for both.
//   warp 0      TMA producer (ring of `stages` slots, runs ahead across tiles)
//   warp 1      TMEM allocator; in the leader CTA also the MMA issuer (one lane)
//   warps 2..17 epilogue: TMEM -> registers -> fused activation / derivative -> coalesced global stores.  Two
//               accumulator buffers in TMEM (2 x BN columns) let the epilogue of tile i overlap the MMAs of tile i+1.
//
// X3 = true is the fp32-grade mode ("3xTF32"): tcgen05 TRUNCATES fp32 inputs to tf32 (measured, tools/
// probe_tf32_rounding.py), so an fp32 array already is its own high part hi = x & 0xFFFFE000; the producers additionally
// keep lo = x - hi (exact).  Each K step then issues A*B + A*B_lo + A_lo*B into the same fp32 accumulator; the dropped
// lo*lo term is 2^-22 relative.  Four tiles per stage instead of two.
// MODE 0: kind::tf32 on the fp32 arrays.  MODE 1: fp32-grade 3xTF32 (fp32 arrays + their low-part companions).
// MODE 2: kind::f16 with bf16 inputs — the operands are the bf16 COMPANION arrays (map_a / map_b describe those), a K block
// is 64 elements, an MMA K step 16; accumulation and the stored output stay fp32, D_lo receives the bf16 copy of the output.
template <int BN, bool A_MN, bool B_MN, int NCTA, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_lo, const TcArgs p) {
    constexpr bool X3 = MODE == 1, BF = MODE == 2;
    constexpr int BKE = BF ? 64 : 32;                       // elements per K block (128 bytes)
    constexpr int MN_ATOM = BF ? 64 : 32;                   // contiguous M/N elements per MN-major box
    constexpr int MN_BOX = BF ? ATOM_BYTES_BF : ATOM_BYTES; // bytes per MN-major box
    constexpr int B_ROWS = BN / NCTA;                       // N rows of B this CTA stages
    constexpr int B_TILE = b_tile_bytes(B_ROWS, B_MN, BF);
    constexpr int HALF_STAGE = A_TILE_BYTES + B_TILE;       // one (A, B) tile pair
    constexpr int STAGE_BYTES = X3 ? 2 * HALF_STAGE : HALF_STAGE;      // per CTA
    constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;        // TMEM columns of one accumulator buffer
    constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;            // 64 .. 512, always a power of two
    constexpr uint32_t IDESC = BF ? umma_idesc_bf16(BM * NCTA, BN, A_MN, B_MN) : umma_idesc_tf32(BM * NCTA, BN, A_MN, B_MN);
    static_assert(!(B_MN && NCTA == 2 && (BN / 2) % MN_ATOM != 0), "a CTA of a pair must stage whole MN-major boxes of B");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    uint64_t* full_bar   = reinterpret_cast<uint64_t*>(smem + size_t(stages) * STAGE_BYTES);
    uint64_t* empty_bar  = full_bar + MAX_STAGES;
    uint64_t* tmem_full  = empty_bar + MAX_STAGES;          // [2]
    uint64_t* tmem_empty = tmem_full + 2;                   // [2]
    uint32_t* tmem_slot  = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* epi_stage = reinterpret_cast<float*>(smem + size_t(stages) * STAGE_BYTES + 256);       // only allocated in scatter mode

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int worker = NCTA == 2 ? int(cluster_id_x()) : int(blockIdx.x);
    const int n_workers = NCTA == 2 ? int(num_clusters_x()) : int(gridDim.x);
    const int tiles_mn = p.tiles_m * p.tiles_n;
    const int total_work = tiles_mn * p.splits;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (X3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], NCTA); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS * NCTA); }
        mbar_fence_init();
    }
    if (warp == 1) {
        if (NCTA == 2) { tmem_alloc_pair(tmem_slot, TMEM_COLS); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
    }
    __syncwarp();                 // warp 0 diverged for the barrier init; cluster barriers need converged warps
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above overlapped the previous kernel's tail (pdl.cuh); from here on its results are visible
    pdl_enter();

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        // Gate fused into the consumer: every owner's rows of this layer's weights (previous step's update, stored into this
        // GPU's arena over NVLink) must have landed before the first TMA load reads them.  Lanes poll one owner flag each.
        if (p.wait_count > 0) {
            if (lane < p.wait_count) {
                const long long t0 = clock64();
                int v;
                do {
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p.wait_flags + lane) : "memory");
                    if (v < p.wait_value && clock64() - t0 > 20000000000LL) __trap();      // a peer died: fail loudly
                } while (v < p.wait_value);
            }
            __syncwarp();
            asm volatile("fence.proxy.async;\n" ::: "memory");       // the TMA (async proxy) loads below must observe those rows
        }
        // the whole warp walks the loop (warp-uniform control flow and operands); one elected lane issues
        const bool issuer = elect_one();
        int s = 0;
        uint32_t ph = 0;
        for (int w = worker; w < total_work; w += n_workers) {
            const int split = w / tiles_mn, rem = w - split * tiles_mn;
            int tile_m, tile_n; decode_tile(rem, p.tiles_m, p.tiles_n, p.group_m, p.n_rot, p.serp, tile_m, tile_n);
            const int m0 = (tile_m * NCTA + int(rank)) * BM;          // this CTA's 128 rows of A
            const int n0 = tile_n * BN + int(rank) * B_ROWS;          // this CTA's share of B
            const int kb0 = split * p.k_blocks;
            int nkb = p.k_blocks_total - kb0;
            if (nkb > p.k_blocks) nkb = p.k_blocks;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* sa = smem + size_t(s) * STAGE_BYTES;
                uint8_t* sb = sa + A_TILE_BYTES;
                uint64_t* bar = &full_bar[s];
                const int kc = (kb0 + kb) * BKE;
                if (issuer) {
                    if (NCTA == 1) mbar_expect_tx(bar, STAGE_BYTES);
                    auto load_pair = [&](uint8_t* ta, uint8_t* tb, const CUtensorMap* ma, const CUtensorMap* mb) {
                        if (A_MN) {
                            if (p.a_3d) {
                                if (NCTA == 2) tma_load_3d_pair(ta, ma, bar, 0, p.a_k0 + kc, (p.a_mn0 + m0) / MN_ATOM);
                                else tma_load_3d(ta, ma, bar, 0, p.a_k0 + kc, (p.a_mn0 + m0) / MN_ATOM);
                            } else {
#pragma unroll
                                for (int i = 0; i < BM / MN_ATOM; ++i) {
                                    if (NCTA == 2) tma_load_2d_pair(ta + i * MN_BOX, ma, bar, p.a_mn0 + m0 + i * MN_ATOM, p.a_k0 + kc);
                                    else tma_load_2d(ta + i * MN_BOX, ma, bar, p.a_mn0 + m0 + i * MN_ATOM, p.a_k0 + kc);
                                }
                            }
                        } else {
                            if (NCTA == 2) tma_load_2d_pair(ta, ma, bar, p.a_k0 + kc, p.a_mn0 + m0);
                            else tma_load_2d(ta, ma, bar, p.a_k0 + kc, p.a_mn0 + m0);
                        }
                        if (B_MN) {
                            if (p.b_3d) {
                                if (NCTA == 2) tma_load_3d_pair(tb, mb, bar, 0, p.b_k0 + kc, (p.b_mn0 + n0) / MN_ATOM);
                                else tma_load_3d(tb, mb, bar, 0, p.b_k0 + kc, (p.b_mn0 + n0) / MN_ATOM);
                            } else {
#pragma unroll
                                for (int i = 0; i < (B_ROWS + MN_ATOM - 1) / MN_ATOM; ++i) {
                                    if (NCTA == 2) tma_load_2d_pair(tb + i * MN_BOX, mb, bar, p.b_mn0 + n0 + i * MN_ATOM, p.b_k0 + kc);
                                    else tma_load_2d(tb + i * MN_BOX, mb, bar, p.b_mn0 + n0 + i * MN_ATOM, p.b_k0 + kc);
                                }
                            }
                        } else {
                            if (NCTA == 2) tma_load_2d_pair(tb, mb, bar, p.b_k0 + kc, p.b_mn0 + n0);
                            else tma_load_2d(tb, mb, bar, p.b_k0 + kc, p.b_mn0 + n0);
                        }
                    };
                    load_pair(sa, sb, &map_a, &map_b);
                    if (X3) load_pair(sa + HALF_STAGE, sb + HALF_STAGE, &map_a_lo, &map_b_lo);
                    if (NCTA == 2) {
                        // the leader's barrier counts one arrival per CTA and the bytes of BOTH CTAs' boxes
                        if (leader) mbar_expect_tx(bar, 2 * STAGE_BYTES);
                        else mbar_arrive_remote(bar, 0);
                    }
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only) ----------------
        if (leader) {
            // Everything the issue loop needs is warp-uniform and kept in registers as (descriptor of stage 0) + offset, so
            // the loop body is a handful of uniform adds around the four tcgen05.mma: with the descriptors rebuilt per
            // instruction inside a lane-0 branch the issuing thread itself (~150 SASS instructions per k-block, each
            // operand moved to the uniform file through a vote loop) was the bottleneck at ~62 % tensor-pipe activity.
            const uint32_t smem_base = smem_u32(smem);
            const uint64_t a_desc0 = A_MN ? umma_smem_desc(smem_base, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                          : umma_smem_desc(smem_base, 16, 1024);
            const uint64_t b_desc0 = B_MN ? umma_smem_desc(smem_base + A_TILE_BYTES, p.mn_lbo, p.mn_sbo, p.mn_layout)
                                          : umma_smem_desc(smem_base + A_TILE_BYTES, 16, 1024);
            // one MMA K step = 32 bytes of K (8 floats / 16 bf16) on a K-major operand, 8 / 16 box rows on an MN-major one
            const uint32_t A_KSTEP = (A_MN ? p.mn_kstep : uint32_t(UMMA_K * 4)) >> 4;      // descriptor units of 16 bytes
            const uint32_t B_KSTEP = (B_MN ? p.mn_kstep : uint32_t(UMMA_K * 4)) >> 4;
            constexpr uint32_t STAGE_STEP = STAGE_BYTES >> 4;
            constexpr uint32_t LO_STEP = HALF_STAGE >> 4;
            const bool issuer = elect_one();
            int s = 0, t = 0;
            uint32_t ph = 0;
            for (int w = worker; w < total_work; w += n_workers, ++t) {
                const int split = w / tiles_mn;
                const int kb0 = split * p.k_blocks;
                int nkb = p.k_blocks_total - kb0;
                if (nkb > p.k_blocks) nkb = p.k_blocks;
                const int acc = t & 1;
                const uint32_t acc_ph = (t >> 1) & 1;
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1);          // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + uint32_t(acc) * ACC_COLS;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    if (issuer) {
                        const uint64_t da0 = a_desc0 + uint64_t(uint32_t(s) * STAGE_STEP);
                        const uint64_t db0 = b_desc0 + uint64_t(uint32_t(s) * STAGE_STEP);
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            auto mma = [&](uint64_t da, uint64_t db, uint32_t accumulate) {
                                if (BF) { if (NCTA == 2) umma_f16_pair(tmem_d, da, db, IDESC, accumulate); else umma_f16(tmem_d, da, db, IDESC, accumulate); }
                                else    { if (NCTA == 2) umma_tf32_pair(tmem_d, da, db, IDESC, accumulate); else umma_tf32(tmem_d, da, db, IDESC, accumulate); }
                            };
                            const uint64_t da = da0 + kk * A_KSTEP, db = db0 + kk * B_KSTEP;
                            mma(da, db, (kb | kk) != 0 ? 1u : 0u);
                            if (X3) {
                                mma(da, db + LO_STEP, 1u);                       // hi(A) * lo(B)
                                mma(da + LO_STEP, db, 1u);                       // lo(A) * hi(B)
                            }
                        }
                        // slot reusable (in both CTAs) once these MMAs have read it; accumulator complete after the last
                        if (NCTA == 2) umma_commit_pair(&empty_bar[s], 3); else umma_commit(&empty_bar[s]);
                        if (kb == nkb - 1) { if (NCTA == 2) umma_commit_pair(&tmem_full[acc], 3); else umma_commit(&tmem_full[acc]); }
                    }
                    __syncwarp();
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ---------------- epilogue ----------------
        const int q = warp & 3;                          // TMEM lane quarter this warp may read
        const int part = (warp - 2) >> 2;                // which slice of the tile's columns
        constexpr int CH = 16;                           // columns per TMEM read: 16 keeps the 576-thread kernel inside 112 registers
        constexpr int COLS_PER_WARP = (BN / EPI_PARTS >= CH) ? BN / EPI_PARTS : CH;
        const int c_begin = part * COLS_PER_WARP;        // >= BN for the surplus warps of narrow tiles: they only arrive
        const int c_end = (c_begin + COLS_PER_WARP < BN) ? c_begin + COLS_PER_WARP : BN;
        const float* __restrict__ aux = p.aux;
        float* __restrict__ Dp = p.D;
        int t = 0;
        for (int w = worker; w < total_work; w += n_workers, ++t) {
            const int split = w / tiles_mn, rem = w - split * tiles_mn;
            int tile_m, tile_n; decode_tile(rem, p.tiles_m, p.tiles_n, p.group_m, p.n_rot, p.serp, tile_m, tile_n);
            const int m = (tile_m * NCTA + int(rank)) * BM + q * 32 + lane;
            const int n0 = tile_n * BN;
            const bool m_ok = m < p.M;
            const int acc = t & 1;
            const uint32_t acc_ph = (t >> 1) & 1;
            // Derivative epilogues read one activation per output.  Those loads are software-pipelined one chunk ahead:
            // the first chunk is requested before the accumulator is even complete, chunk c+1 before chunk c is stored,
            // so reads stay in flight through the store phases (the K = 10 backprop of the output layer is all epilogue).
            const bool deriv = p.epi >= EPI_DSIGMOID;
            // Addressing is incremental (one 64-bit add per element) with an unpredicated path for interior chunks: with
            // a 64-bit multiply and two predicates per access the epilogue ran ~25 instructions per output element and
            // the K = 10 backprop of the output layer was bound by instruction issue (ncu: 26 M warp instructions).
            auto load_aux = [&](int c0, float (&dst)[CH]) {
                const float* ap = aux + int64_t(n0 + c0) * p.ldaux + m;
                if (m_ok && n0 + c0 + CH <= p.N) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) { dst[j] = __ldg(ap); ap += p.ldaux; }
                } else {
#pragma unroll
                    for (int j = 0; j < CH; ++j) { dst[j] = (n0 + c0 + j < p.N && m_ok) ? __ldg(ap) : 0.f; ap += p.ldaux; }
                }
            };
            float av[CH];
            if (deriv && n0 + c_begin < p.N) load_aux(c_begin, av);
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            const uint32_t lane_addr = tmem_base + uint32_t(acc) * ACC_COLS + (uint32_t(q * 32) << 16);
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += CH) {
                if (n0 + c0 >= p.N) break;                // warp-uniform
                float av_next[CH];
                const bool more = deriv && c0 + CH < c_end && n0 + c0 + CH < p.N;
                if (more) load_aux(c0 + CH, av_next);
                uint32_t r[CH];
                tmem_ld_32x16(lane_addr + c0, r);
                tmem_ld_wait();
                float v[CH];
                switch (p.epi) {                          // warp-uniform
                    case EPI_SIGMOID:
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = sigmoid_fast(__uint_as_float(r[j]));
                        break;
                    case EPI_TANH:
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = tanh_fast(__uint_as_float(r[j]));
                        break;
                    case EPI_DSIGMOID:
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) * (av[j] * (1.0f - av[j]));
                        break;
                    case EPI_DTANH:
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]) * (1.0f - av[j] * av[j]);
                        break;
                    default:
#pragma unroll
                        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(r[j]);
                        break;
                }
                if (X3 && p.D_lo != nullptr && m_ok) {
                    // low part of the output for the next contraction that consumes it: exact remainder of the truncation
                    float* dlo = p.D_lo + int64_t(n0 + c0) * p.ldd + m;
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        if (n0 + c0 + j < p.N) *dlo = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xFFFFE000u);
                        dlo += p.ldd;
                    }
                }
                if (BF && p.D_lo != nullptr && m_ok) {
                    // bf16 copy of the output (round to nearest even): the operand of the next contraction that consumes it
                    __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(p.D_lo) + int64_t(n0 + c0) * p.ldd + m;
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        if (n0 + c0 + j < p.N) *dh = __float2bfloat16_rn(v[j]);
                        dh += p.ldd;
                    }
                }
                if (p.d_trans) {
                    // swapped plan: lane = unit, columns = samples -> each thread owns CH consecutive floats of one row
                    if (m_ok) {
                        float* dst = Dp + int64_t(m) * p.ldd + n0 + c0;
                        if (p.atomic) {
#pragma unroll
                            for (int j = 0; j < CH; ++j) if (n0 + c0 + j < p.N) atomicAdd(dst + j, v[j]);
                        } else if (n0 + c0 + CH <= p.N) {
#pragma unroll
                            for (int j = 0; j < CH; j += 4)
                                *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < CH; ++j) if (n0 + c0 + j < p.N) dst[j] = v[j];
                        }
                    }
                } else if (p.scatter_bulk) {
                    // Fused reduce-scatter over NVLink: the four warps that share this column slice (one per TMEM lane
                    // quarter) stage their 16 x 32 chunks side by side (row = output column, 128 consecutive m), then every
                    // column leaves as ONE 512-byte bulk asynchronous store into the owner's arena.  Ordinary stores to a
                    // peer stall the epilogue warps (an SM tracks only a handful of outstanding remote writes: measured
                    // ~1.2 GB/s per SM, the epilogue then outlasts the next tile's MMAs); bulk stores are queued to the TMA
                    // unit and the warps move on.
                    float* stg = epi_stage + part * (CH * BM);
                    const uint32_t bar_id = 1u + uint32_t(part);       // named barrier of this slice's four warps
                    if (lane < CH / 4) bulk_wait_read_all();            // the columns this lane sent last time have left
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
#pragma unroll
                    for (int j = 0; j < CH; ++j) stg[j * BM + q * 32 + lane] = v[j];
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;\n" ::"r"(bar_id) : "memory");
                    const int m0c = (tile_m * NCTA + int(rank)) * BM;   // this CTA's first output row index m
                    const int col = q * (CH / 4) + lane;                // each warp sends a quarter of the chunk's columns
                    const int n = n0 + c0 + col;
                    if (lane < CH / 4 && n < p.N && m0c < int(p.ldd)) {
                        const int owner = n0 / p.scatter_rows;
                        float* dst = p.scatter_base[owner] + (int64_t(p.scatter_rank - owner) * p.scatter_rows + n) * p.ldd + m0c;
                        const int run = int(p.ldd) - m0c < BM ? int(p.ldd) - m0c : BM;      // pad columns of the row receive zeros
                        bulk_store(dst, stg + col * BM, uint32_t(run) * 4u);
                    }
                    if (lane < CH / 4) bulk_commit();
                } else if (m_ok) {
                    float* dst = Dp + int64_t(n0 + c0) * p.ldd + m;
                    if (p.scatter_world > 0) {
                        // this tile's rows belong to rank `owner`: store into its arena over NVLink, slice = my rank
                        const int owner = n0 / p.scatter_rows;
                        dst = p.scatter_base[owner] + (int64_t(p.scatter_rank - owner) * p.scatter_rows + n0 + c0) * p.ldd + m;
                    }
                    if (p.atomic) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) { if (n0 + c0 + j < p.N) atomicAdd(dst, v[j]); dst += p.ldd; }
                    } else if (n0 + c0 + CH <= p.N) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) { *dst = v[j]; dst += p.ldd; }
                    } else {
#pragma unroll
                        for (int j = 0; j < CH; ++j) { if (n0 + c0 + j < p.N) *dst = v[j]; dst += p.ldd; }
                    }
                }
                if (more) {
#pragma unroll
                    for (int j = 0; j < CH; ++j) av[j] = av_next[j];
                }
            }
            // this warp is done reading the accumulator buffer: hand it back to the MMA issuer (leader CTA)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (NCTA == 2) mbar_arrive_remote(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }

            // ---- fused output layer: the warp whose partial sums complete this 32-sample slice finishes it ----
            if (p.out.counters != nullptr && c_begin == 0) {
                const int slice = (tile_m * NCTA + int(rank)) * 4 + q;
                __threadfence();                                   // this warp's red.adds / stores of z are performed
                int last = 0;
                if (lane == 0) last = atomicAdd(p.out.counters + slice, 1) == p.splits - 1;
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence();
                    fused_output_slice(p.out.y, p.out.ldy, p.out.delta, p.out.ldd, p.out.cost_sum, p.out.out_kind, p.out.cost_kind, Dp, p.ldd,
                                       p.N, m, m_ok, lane);
                    if (lane == 0) p.out.counters[slice] = 0;      // ready for the next launch
                }
            }
        }
        if (p.scatter_bulk) bulk_wait_all();       // every queued store has been performed before the kernel retires
    }

    __syncwarp();                 // producer / issuer lanes rejoin their warps
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) { if (NCTA == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ---- host side ----------------------------------------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;
std::string g_encode_err;

void load_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
        g_encode_err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
        (void)cudaGetLastError();
        return;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

// Tensor map over the allocation that holds the operand.  The tensor extent ends where the sub-matrix ends, so
// TMA's out-of-bounds zero fill supplies the K tail (wgrad: nothing past the batch is summed) and the M/N tails.
bool encode_operand(CUtensorMap* map, const void* base, int64_t ld, bool mn_major, int64_t mn_end, int64_t k_end,
                    int box_rows, bool as_3d, std::string& err, bool bf = false) {
    const int64_t esz = bf ? 2 : 4;                // element size
    const cuuint32_t atom = bf ? 64u : 32u;        // elements per 128-byte row of a box
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) { err = "operand base not 16-byte aligned"; return false; }
    if ((ld * esz) % 16 != 0) { err = "leading dimension not a multiple of 16 bytes"; return false; }
    const int64_t inner = mn_major ? mn_end : k_end;
    const int64_t outer = mn_major ? k_end : mn_end;
    if (inner <= 0 || outer <= 0 || inner > ld) { err = "operand extent inconsistent with leading dimension"; return false; }
    if (inner >= (int64_t(1) << 32) || outer >= (int64_t(1) << 32)) { err = "operand extent exceeds TMA limits"; return false; }
    // MN-major fp32 tiles must use the 32-byte-chunk flavour of the 128-byte swizzle (see umma_smem_desc); two-byte
    // elements take the ordinary 128-byte swizzle in both major-nesses.
    CUtensorMapSwizzle swz = (mn_major && !bf) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (mn_major) if (const char* e = std::getenv("B2NN_TC_MN_TMA_SWIZZLE")) swz = CUtensorMapSwizzle(std::atoi(e));
    CUresult r;
    if (mn_major && as_3d) {
        // The contiguous M/N index is folded into (32 floats, atom): {32, k, atoms} with strides {ld*4, 128} bytes, so a
        // single box {32, 32, box_rows/32} lands in shared memory exactly as box_rows/32 consecutive 2-D boxes would.
        // Elements of the last atom past mn_end alias the start of the next k-row (readable: rows are padded, arenas carry
        // slack); they only feed output rows / columns >= M / N, which the epilogue never stores.
        cuuint64_t gdim[3] = {atom, cuuint64_t(outer), cuuint64_t((inner + atom - 1) / atom)};
        cuuint64_t gstride[2] = {cuuint64_t(ld * esz), 128u};
        cuuint32_t box[3] = {atom, atom, cuuint32_t((box_rows + atom - 1) / atom)};
        cuuint32_t estr[3] = {1u, 1u, 1u};
        r = g_encode(map, dt, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t gdim[2] = {cuuint64_t(inner), cuuint64_t(outer)};
        cuuint64_t gstride[1] = {cuuint64_t(ld * esz)};
        cuuint32_t box[2] = {atom, cuuint32_t(mn_major ? atom : box_rows)};
        cuuint32_t estr[2] = {1u, 1u};
        r = g_encode(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { err = "cuTensorMapEncodeTiled failed, CUresult " + std::to_string(int(r)); return false; }
    return true;
}

int g_num_sms = 0;
int g_reserved_sms = 0;      // SMs left free for a concurrent communication kernel (data parallel)

template <int BN, bool A_MN, bool B_MN, int NCTA, int MODE>
cudaError_t launch_one(const TcPlan& pl, const TcArgs& a, cudaStream_t stream) {
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, NCTA, MODE>;
    static std::atomic<unsigned long long> done{0};   // per instantiation, one bit per device
    if (cudaError_t e = ensure_dyn_smem(kern, SMEM_LIMIT, done)) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.grid_x, 1, 1);
    cfg.blockDim = dim3(NUM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (NCTA == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = NCTA; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, pl.map_a, pl.map_b, MODE == 1 ? pl.map_a_lo : pl.map_a, MODE == 1 ? pl.map_b_lo : pl.map_b, a);
}

// A CTA of a pair stages BN/2 columns of B; MN-major B comes in whole boxes of 32 (fp32) or 64 (bf16) columns.
template <int BN, int NCTA, int MODE>
constexpr bool b_mn_ok() { return NCTA == 1 ? (MODE == 2 ? BN >= 64 : BN >= 32) : (BN / 2) % (MODE == 2 ? 64 : 32) == 0; }

template <int BN, int NCTA, int MODE>
cudaError_t launch_majors(const TcPlan& pl, const TcArgs& a, cudaStream_t stream) {
    const bool am = (pl.swapped ? pl.g.b_major : pl.g.a_major) != 0, bm = (pl.swapped ? pl.g.a_major : pl.g.b_major) != 0;
    if (am && bm) {
        if constexpr (!b_mn_ok<BN, NCTA, MODE>()) return cudaErrorInvalidValue; else return launch_one<BN, true, true, NCTA, MODE>(pl, a, stream);
    }
    if (am && !bm) return launch_one<BN, true, false, NCTA, MODE>(pl, a, stream);
    if (!am && bm) {
        if constexpr (!b_mn_ok<BN, NCTA, MODE>()) return cudaErrorInvalidValue; else return launch_one<BN, false, true, NCTA, MODE>(pl, a, stream);
    }
    return launch_one<BN, false, false, NCTA, MODE>(pl, a, stream);
}

template <int BN>
cudaError_t launch_bn(const TcPlan& pl, const TcArgs& a, cudaStream_t stream) {
    if (pl.x3) {
        // the fp32-grade mode keeps 4 tiles per stage: a single CTA takes tiles up to 128 columns, a pair up to 256
        if (pl.ncta == 2) return launch_majors<BN, 2, 1>(pl, a, stream);
        if constexpr (BN <= 128) return launch_majors<BN, 1, 1>(pl, a, stream); else return cudaErrorInvalidValue;
    }
    if (pl.bf16) return pl.ncta == 2 ? launch_majors<BN, 2, 2>(pl, a, stream) : launch_majors<BN, 1, 2>(pl, a, stream);
    return pl.ncta == 2 ? launch_majors<BN, 2, 0>(pl, a, stream) : launch_majors<BN, 1, 0>(pl, a, stream);
}

}  // namespace

void tc_set_reserved_sms(int n) { g_reserved_sms = n < 0 ? 0 : n; }

bool tc_runtime_ok(std::string& err) {
    std::call_once(g_encode_once, load_encode);
    if (!g_encode) { err = g_encode_err; return false; }
    return true;
}

// The description the kernel actually runs: the caller's, or (swapped plan) D^T = B A^T with a transposed store.
static GemmDesc effective_desc(const GemmDesc& u, bool swapped) {
    if (!swapped) return u;
    GemmDesc e = u;
    e.A = u.B; e.lda = u.ldb; e.a_major = u.b_major; e.a_mn0 = u.b_mn0; e.a_k0 = u.b_k0; e.a_mn_end = u.b_mn_end; e.a_k_end = u.b_k_end;
    e.B = u.A; e.ldb = u.lda; e.b_major = u.a_major; e.b_mn0 = u.a_mn0; e.b_k0 = u.a_k0; e.b_mn_end = u.a_mn_end; e.b_k_end = u.a_k_end;
    e.M = u.N; e.N = u.M;
    e.d_trans = true;
    return e;
}

bool tc_plan_build(TcPlan& plan, const GemmDesc& gu, std::string& err) {
    plan.valid = false;
    plan.swapped = false;
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_num_sms = n;
    }
    // Skinny forward contraction (activations MN-major x weights K-major, <= 128 output units) that would leave most of
    // the machine idle as ceil(rows / 256) pair tiles: run it swapped, as one 128-lane tile of units x many small tiles
    // of samples (config C1's 4200 x 50 x 785: 17 pair tiles -> 132 tiles of 32 samples).
    // Not when split-K is available (plain store, split_k == 0): a long K then spreads over the SMs as splits, which
    // keeps 16 KB per stage in flight instead of the swapped plan's 4 KB sample tiles (C3's 8192 x 10 x 4097 output layer:
    // 27 us split against ~75 us swapped).  Measured gains: small-batch inference on the MNIST net 29 -> 20 us.
    const bool may_split = gu.epi == EPI_STORE && gu.split_k == 0;
    if (gu.a_major == 1 && gu.b_major == 0 && gu.N <= BM && gu.M > 0 && gu.epi <= EPI_TANH && !gu.A_lo && !gu.D_lo && !may_split &&
        gu.scatter_world == 0 && gu.split_k <= 1 && !gu.d_trans && gu.ldd % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(gu.D) & 15) == 0 && !std::getenv("B2NN_TC_NO_SWAP")) {
        const int64_t pair_tiles = (int64_t(gu.M) + 2 * BM - 1) / (2 * BM);
        if (pair_tiles * 2 < g_num_sms) plan.swapped = true;
    }
    const GemmDesc g = effective_desc(gu, plan.swapped);
    if (!tc_runtime_ok(err)) return false;
    if (g.M <= 0 || g.N <= 0 || g.K <= 0) { err = "empty contraction"; return false; }
    if (g.a_mn0 + g.M >= (int64_t(1) << 31) || g.a_k0 + g.K >= (int64_t(1) << 31) ||
        g.b_mn0 + g.N >= (int64_t(1) << 31) || g.b_k0 + g.K >= (int64_t(1) << 31)) { err = "coordinates exceed int32"; return false; }
    if (g.epi >= EPI_DSIGMOID && !g.aux) { err = "derivative epilogue without aux"; return false; }

    int bn;
    if (g.N > 128) bn = 256; else if (g.N > 64) bn = 128; else if (g.N > 32) bn = 64; else if (g.N > 16) bn = 32; else bn = 16;
    const bool bf = g.bf16 && g.A_lo != nullptr && g.B_lo != nullptr;       // bf16 mode: the companions ARE the operands
    const int atom = bf ? 64 : 32;         // an MN-major box is 32 floats / 64 bf16 wide
    if (g.b_major && bn < atom) bn = atom;
    if (plan.swapped) {
        // largest sample tile that still gives every SM one: the weights are re-read per tile, so wider is cheaper
        bn = 32;
        for (int cand : {256, 128, 64}) if ((int64_t(g.N) + cand - 1) / cand >= g_num_sms) { bn = cand; break; }
    }
    plan.bf16 = bf;
    plan.x3 = !bf && g.A_lo != nullptr && g.B_lo != nullptr;
    // CTA pairs (cta_group::2, 256-row tiles, each CTA stages half of B) whenever there are two 128-row tiles to pair:
    // 745 vs 692 TFLOP/s on 8192 x 4096 x 4097 (profiles/r1_gemm_pair.md).  B2NN_TC_NCTA=1 forces single CTAs.
    int ncta = g.M > BM ? 2 : 1;
    if (const char* e = std::getenv("B2NN_TC_NCTA")) { if (std::atoi(e) == 1) ncta = 1; }
    if (g.b_major && bn < 2 * atom) ncta = 1;    // a CTA of a pair must stage whole MN-major boxes of B
    // 3xTF32 stages four tiles per k-block: 128 x 128 alone (3 stages), 256 x 256 as a pair (3 stages)
    if (plan.x3 && ncta == 1 && bn > 128) bn = 128;
    if (g.scatter_world > 0 && g.scatter_rows % bn != 0) { err = "scatter rows not a multiple of the tile"; return false; }
    plan.block_n = bn;
    plan.g = gu;
    plan.ncta = ncta;

    const int stage_bytes = (A_TILE_BYTES + b_tile_bytes(bn / ncta, g.b_major != 0, bf)) * (plan.x3 ? 2 : 1);
    // scatter mode stages its output chunks for bulk stores: 32 KB less for the operand ring (7 -> 6 stages of 32 KB)
    plan.scatter_bulk = g.scatter_world > 0 && !g.d_trans && (g.ldd % 4) == 0 && !std::getenv("B2NN_TC_NO_BULK");
    const int epi_bytes = plan.scatter_bulk ? EPI_STAGE_BYTES : 0;
    int stages = (SMEM_LIMIT - 1024 - 256 - epi_bytes) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) { err = "tile does not fit shared memory"; return false; }
    plan.stages = stages;
    plan.smem_bytes = size_t(stages) * stage_bytes + 1024 + 256 + epi_bytes;

    const int bke = bf ? 64 : BK;          // elements per 128-byte K block
    plan.k_blocks_total = (g.K + bke - 1) / bke;
    plan.tiles_m = (g.M + BM * ncta - 1) / (BM * ncta);
    plan.tiles_n = (g.N + bn - 1) / bn;
    int usable_sms = g_num_sms - g_reserved_sms;
    if (const char* e = std::getenv("B2NN_TC_RESERVE_SMS")) usable_sms = g_num_sms - std::atoi(e);
    if (usable_sms < 2 * ncta) usable_sms = 2 * ncta;
    const int64_t max_workers = usable_sms / ncta;
    int splits = 1;
    if (g.scatter_world > 0) splits = 1;        // partial tiles travel as plain stores: no split-K
    else if (g.epi == EPI_STORE && g.split_k > 1) splits = g.split_k;
    else if (g.epi == EPI_STORE && g.split_k == 0) {
        // auto split-K: few output tiles but a long K (weight gradients, the 10-column output layer).  Model: time ~
        // waves * (K blocks per item + fixed per-item cost); items beyond one wave only pay when they shorten it.
        const int64_t tiles = int64_t(plan.tiles_m) * plan.tiles_n;
        if (tiles < max_workers) {
            double best = 1e30;
            const int smax = std::min(64, std::max(1, plan.k_blocks_total / 4));
            for (int sp = 1; sp <= smax; ++sp) {
                const int64_t items = tiles * sp;
                const int64_t waves = (items + max_workers - 1) / max_workers;
                const int kb = (plan.k_blocks_total + sp - 1) / sp;
                const double cost = double(waves) * (kb + 6.0) + 0.02 * sp;       // tie-break towards fewer atomics
                if (cost < best) { best = cost; splits = sp; }
            }
        }
    }
    if (splits > plan.k_blocks_total) splits = plan.k_blocks_total;
    plan.k_blocks = (plan.k_blocks_total + splits - 1) / splits;
    splits = (plan.k_blocks_total + plan.k_blocks - 1) / plan.k_blocks;
    plan.splits = splits;
    const int64_t work = int64_t(plan.tiles_m) * plan.tiles_n * splits;
    if (work >= (int64_t(1) << 30)) { err = "too many tiles"; return false; }
    plan.grid_x = unsigned(std::min<int64_t>(work, max_workers) * ncta);      // persistent: one CTA (pair) per SM (pair)

    const int64_t a_mn_end = g.a_mn_end > 0 ? g.a_mn_end : g.a_mn0 + g.M;
    const int64_t a_k_end = g.a_k_end > 0 ? g.a_k_end : g.a_k0 + g.K;
    // One 3-D box per MN-major tile needs the operand origin on a 32-float atom boundary (tile offsets always are) and
    // the folded view to stay inside memory the caller owns (mn_slack: the operand allocation is padded accordingly).
    const bool no3d = std::getenv("B2NN_TC_NO_3D") != nullptr;
    plan.a_3d = g.a_major && g.mn_slack && !no3d && (g.a_mn0 % atom == 0);
    plan.b_3d = g.b_major && g.mn_slack && !no3d && (g.b_mn0 % atom == 0);
    const void* a_src = bf ? static_cast<const void*>(g.A_lo) : static_cast<const void*>(g.A);
    const void* b_src = bf ? static_cast<const void*>(g.B_lo) : static_cast<const void*>(g.B);
    if (!encode_operand(&plan.map_a, a_src, g.lda, g.a_major != 0, a_mn_end, a_k_end, BM, plan.a_3d, err, bf)) return false;
    const int64_t b_mn_end = g.b_mn_end > 0 ? g.b_mn_end : g.b_mn0 + g.N;
    const int64_t b_k_end = g.b_k_end > 0 ? g.b_k_end : g.b_k0 + g.K;
    if (!encode_operand(&plan.map_b, b_src, g.ldb, g.b_major != 0, b_mn_end, b_k_end, bn / ncta, plan.b_3d, err, bf)) return false;   // each CTA of a pair stages half of B
    if (plan.x3) {
        if (!encode_operand(&plan.map_a_lo, g.A_lo, g.lda, g.a_major != 0, a_mn_end, a_k_end, BM, plan.a_3d, err)) return false;
        if (!encode_operand(&plan.map_b_lo, g.B_lo, g.ldb, g.b_major != 0, b_mn_end, b_k_end, bn / ncta, plan.b_3d, err)) return false;
    }
    plan.valid = true;
    return true;
}

cudaError_t tc_plan_launch(const TcPlan& pl, cudaStream_t stream, const TcWait* wait, int64_t a_shift, const TcOutput* out) {
    if (!pl.valid) return cudaErrorInvalidValue;
    GemmDesc g = effective_desc(pl.g, pl.swapped);
    // a_shift: the caller's A operand (the training matrix) starts `a_shift` samples further on — only its TMA coordinates
    // move, the tensor maps span the whole matrix.  (In the swapped plan the caller's A is this description's B.)
    if (a_shift != 0) {
        if (pl.swapped) { if (g.b_major) g.b_mn0 += a_shift; else g.b_k0 += a_shift; }
        else            { if (g.a_major) g.a_mn0 += a_shift; else g.a_k0 += a_shift; }
    }
    TcArgs a;
    a.D = g.D; a.ldd = g.ldd; a.aux = g.aux; a.ldaux = g.ldaux;
    a.D_lo = (pl.x3 || pl.bf16) ? g.D_lo : nullptr;
    for (int i = 0; i < 8; ++i) a.scatter_base[i] = g.scatter_base[i];
    a.scatter_world = g.scatter_world; a.scatter_rank = g.scatter_rank; a.scatter_rows = g.scatter_rows;
    a.scatter_bulk = pl.scatter_bulk ? 1 : 0;
    a.out = TcOutput{};
    if (out) {
        // one column tile, plain store / split-K epilogue, samples on the TMEM lanes
        if (pl.swapped || pl.tiles_n != 1 || g.N > 16 || g.epi != EPI_STORE || g.d_trans) return cudaErrorInvalidValue;
        a.out = *out;
    }
    a.wait_flags = wait ? wait->flags : nullptr; a.wait_count = wait ? wait->count : 0; a.wait_value = wait ? wait->value : 0;
    a.n_rot = 0;
    if (g.scatter_world > 1 && !std::getenv("B2NN_TC_NO_ROT"))
        a.n_rot = (((g.scatter_rank + 1) % g.scatter_world) * (g.scatter_rows / pl.block_n)) % pl.tiles_n;
    a.M = g.M; a.N = g.N; a.d_trans = g.d_trans ? 1 : 0;
    a.tiles_m = pl.tiles_m; a.tiles_n = pl.tiles_n; a.splits = pl.splits;
    {
        // Groups of 8 row-tiles walked serpentine (profiles/r2_tile_walk.md): against the round-1 walk (groups of 16, column
        // tiles always ascending) the three 8192 x 4096 x 4096 contractions of config C3 read 11 % less DRAM and the
        // sustained step is ~1 % shorter; row-fastest order (one group) costs 582 vs 611 TFLOP/s.
        int gm = 8;
        if (const char* e = std::getenv("B2NN_TC_GROUP_M")) gm = std::atoi(e);
        if (gm < 1) gm = 1;
        if (gm > pl.tiles_m) gm = pl.tiles_m;
        a.group_m = gm;
        a.serp = 1;
        if (const char* e = std::getenv("B2NN_TC_SERP")) a.serp = std::atoi(e) ? 1 : 0;
        if (a.n_rot) a.serp = 0;    // the reduce-scatter rotation fixes the column order (own tiles last)
    }
    a.k_blocks = pl.k_blocks; a.k_blocks_total = pl.k_blocks_total;
    a.a_mn0 = int(g.a_mn0); a.a_k0 = int(g.a_k0); a.b_mn0 = int(g.b_mn0); a.b_k0 = int(g.b_k0);
    a.epi = g.epi; a.atomic = pl.splits > 1 ? 1 : 0; a.stages = pl.stages;
    a.a_3d = pl.a_3d ? 1 : 0; a.b_3d = pl.b_3d ? 1 : 0;
    a.mn_lbo = ATOM_BYTES; a.mn_sbo = 512; a.mn_layout = 1; a.mn_kstep = 1024;
    // bf16 MN-major: boxes of 64 (mn) x 64 (k), ordinary 128-byte swizzle; 8-row groups 1024 B apart, boxes 8192 B apart
    if (pl.bf16) { a.mn_lbo = ATOM_BYTES_BF; a.mn_sbo = 1024; a.mn_layout = 2; a.mn_kstep = 2048; }
    if (const char* e = std::getenv("B2NN_TC_MN_KSTEP")) a.mn_kstep = uint32_t(std::atoi(e));
    if (const char* e = std::getenv("B2NN_TC_MN_LBO")) a.mn_lbo = uint32_t(std::atoi(e));   // bring-up knobs
    if (const char* e = std::getenv("B2NN_TC_MN_SBO")) a.mn_sbo = uint32_t(std::atoi(e));
    if (const char* e = std::getenv("B2NN_TC_MN_LAYOUT")) a.mn_layout = uint32_t(std::atoi(e));
    switch (pl.block_n) {
        case 16:  return launch_bn<16>(pl, a, stream);
        case 32:  return launch_bn<32>(pl, a, stream);
        case 64:  return launch_bn<64>(pl, a, stream);
        case 128: return launch_bn<128>(pl, a, stream);
        case 256: return launch_bn<256>(pl, a, stream);
        default:  return cudaErrorInvalidValue;
    }
}

}  // namespace b2
