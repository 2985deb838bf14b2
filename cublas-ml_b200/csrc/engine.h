// engine.h — internal declarations shared by the CUDA translation units of libb2nn.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>

namespace b2 {

// Epilogues fused into the contraction kernels.
enum Epi : int {
    EPI_STORE = 0,     // D = acc
    EPI_SIGMOID = 1,   // D = 1/(1+exp(-acc))                 (forward, activations.cu:34-39 fused)
    EPI_TANH = 2,      // D = tanh(acc)                       (forward, activations.cu:51-56 fused)
    EPI_DSIGMOID = 3,  // D = acc * a(1-a), a = aux           (delta backprop, activations.cu:41-49 + kernels.cu:53-58 fused)
    EPI_DTANH = 4      // D = acc * (1-a^2), a = aux          (delta backprop, activations.cu:58-66 + kernels.cu:53-58 fused)
};

// D(m,n) = sum_k A(m,k) * B(n,k);  D stored "m contiguous": D[n*ldd + m].
// An operand is K-major when k is its contiguous index (X[mn*ld + k]) and MN-major otherwise (X[k*ld + mn]).
// (mn0, k0) is the origin of the operand inside its allocation: TMA needs the 16-byte aligned allocation base and
// addresses the sub-matrix by coordinates, the FFMA kernel just offsets the pointer.
struct GemmDesc {
    const float* A = nullptr; int64_t lda = 0; int a_major = 0; int64_t a_mn0 = 0, a_k0 = 0;
    const float* B = nullptr; int64_t ldb = 0; int b_major = 0; int64_t b_mn0 = 0, b_k0 = 0;
    float* D = nullptr; int64_t ldd = 0;
    int M = 0, N = 0, K = 0;
    int epi = EPI_STORE;
    const float* aux = nullptr; int64_t ldaux = 0;
    int split_k = 1;          // > 1: partial sums are added atomically into a pre-zeroed D (EPI_STORE only); 0: the tcgen05 plan picks
    // Extent of A's allocation along each index when it is larger than origin + size (0 = origin + size).  Lets one
    // tensor map over the whole training matrix serve every batch.
    int64_t a_mn_end = 0, a_k_end = 0;
    int64_t b_mn_end = 0, b_k_end = 0;     // same for B (only set by the swapped plan below)
    // D(m,n) stored n-contiguous (D[m * ldd + n]) instead of m-contiguous.  Internal to the tcgen05 plan: a skinny
    // forward contraction is run with the operands swapped (units on the TMEM lanes, samples on the columns) so its
    // sample dimension spreads over many small tiles; the transposed store restores the unit-major layout.
    bool d_trans = false;
    // The caller guarantees that reading up to 31 floats past the M/N extent of an MN-major operand (in every k row,
    // including the last) stays inside its allocation, which the one-box-per-tile 3-D TMA view relies on.
    bool mn_slack = false;
    // fp32-grade tensor-core mode (3xTF32): low parts x - trunc_tf32(x) of both operands, same layout and leading
    // dimension as A / B; D_lo (optional) receives the low part of the output.  Null = plain tf32.
    const float* A_lo = nullptr; const float* B_lo = nullptr; float* D_lo = nullptr;
    // bf16 mode: the three companion pointers above hold BF16 copies (two bytes per element, same indexing and leading
    // dimensions) and the contraction runs on them alone with kind::f16; D stays fp32, D_lo receives its bf16 copy.
    bool bf16 = false;
    // Reduce-scatter fused into the epilogue (data parallel over peer memory): output rows n in [o*R, (o+1)*R) are stored
    // into scatter_base[o] + ((rank - o) * R + n) * ldd + m, i.e. into owner o's arena, slice `rank`.  R % tile_n == 0.
    float* scatter_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int scatter_world = 0, scatter_rank = 0, scatter_rows = 0;
};

// fp32 FFMA backend (gemm_simt.cu)
cudaError_t gemm_simt_launch(const GemmDesc& g, cudaStream_t stream);

// tcgen05 / TMA backend (gemm_tc.cu).  The plan owns the two tensor maps; building it is host-only work.
struct TcPlan {
    alignas(64) CUtensorMap map_a;
    alignas(64) CUtensorMap map_b;
    alignas(64) CUtensorMap map_a_lo;
    alignas(64) CUtensorMap map_b_lo;
    bool x3 = false;
    bool bf16 = false;
    GemmDesc g;
    int block_n = 0;          // tile N: 16, 32, 64, 128 or 256
    int stages = 0;
    int k_blocks = 0;         // per split
    int k_blocks_total = 0;
    int ncta = 1;             // 1: one CTA per 128-row tile; 2: CTA pair (cta_group::2) per 256-row tile
    int tiles_m = 0, tiles_n = 0, splits = 1;
    unsigned grid_x = 0;      // persistent grid: min(work items, SMs / ncta) * ncta
    size_t smem_bytes = 0;
    bool a_3d = false, b_3d = false;
    bool scatter_bulk = false;   // scatter mode: output chunks leave through bulk asynchronous stores (32 KB of staging)
    bool swapped = false;     // the kernel runs D^T = B A^T (see GemmDesc::d_trans); `g` stays the caller's description
    bool valid = false;
};
// Returns false (and sets err) when the operands violate a TMA constraint (alignment, stride) so the caller can
// route the contraction to the FFMA kernel instead.
bool tc_plan_build(TcPlan& plan, const GemmDesc& g, std::string& err);
// Optional gate fused into the kernel: its TMA producer spins until flags[0 .. count) >= value before the first load.
struct TcWait { const int* flags; int count; int value; };
// Output layer fused into the last forward contraction (N <= 16 output units, training): the epilogue warp that completes
// a 32-sample slice — with split-K the LAST of the splits to add its partial sums, found through a self-resetting counter —
// applies the output activation, writes delta = h - y and sums the cost (activations.cu:68-88, kernels.cu:46-51,
// costfunctions.cu:14-26 in the epilogue of cublasNN.cpp:658-659's contraction).
struct TcOutput {
    const float* y; int64_t ldy;      // targets, laid out like D (unit-major)
    float* delta; int64_t ldd;        // h - y
    double* cost_sum;                 // += per-element cost in double (may be null)
    int* counters;                    // one per 32-sample slice, zero between launches
    int out_kind, cost_kind;
};
cudaError_t tc_plan_launch(const TcPlan& plan, cudaStream_t stream, const TcWait* wait = nullptr, int64_t a_shift = 0,
                           const TcOutput* out = nullptr);
bool tc_runtime_ok(std::string& err);      // driver entry point for cuTensorMapEncodeTiled available?
// Persistent GEMM grids leave `n` SMs unoccupied so a concurrent NCCL kernel is never queued behind a whole GEMM.
void tc_set_reserved_sms(int n);

// element-wise / bandwidth-bound kernels (elementwise.cu)
struct OutputArgs {
    const float* z; int64_t ldz;      // rows x K, unit-major: z[k*ldz + i]
    const float* y; int64_t ldy;      // targets, same layout
    float* h;       int64_t ldh;      // activated output (may be null)
    float* delta;   int64_t ldd;      // h - y (may be null)
    double* cost_sum;                 // += sum of per-element cost, double accumulator (may be null)
    float* pred;                      // argmax class per row as float (may be null)
    double* err_sum;                  // += #(argmax h != argmax y), double accumulator (may be null)
    const float* y_labels;            // optional class ids per row (validate path) used for the error count
    int64_t rows; int K; int out_kind; int cost_kind;
};
cudaError_t output_layer_launch(const OutputArgs& a, cudaStream_t stream);
cudaError_t momentum_update_launch(float* theta, float* vel, const float* grad, int64_t count, float mu, float alpha,
                                   cudaStream_t stream);
cudaError_t momentum_update_zero_launch(float* theta, float* vel, float* grad, int64_t count, float mu, float alpha,
                                        cudaStream_t stream);   // ... and leaves grad all-zero
// reference arena (Theta_j (in+1) x out col-major, bias row 0)  <->  internal (out rows of ldw floats, bias at col in)
cudaError_t theta_ref_to_internal_launch(const float* ref, float* w, int in, int out, int64_t ldw, cudaStream_t s);
cudaError_t theta_internal_to_ref_launch(const float* w, float* ref, int in, int out, int64_t ldw, cudaStream_t s);
// reference X (m x (n+1) col-major, ones column first) -> internal ((n+1) rows of ld floats, ones row LAST)
cudaError_t stage_x_launch(const float* x_ref, int64_t ld_ref, float* x_int, int64_t ld_int, int64_t rows, int n,
                           cudaStream_t s);
cudaError_t fill_launch(float* p, float v, int64_t count, cudaStream_t s);
// *out += sum |J[k*ld + i]| over i < rows, k < K   (cublasSasum over the valid region, double accumulator)
cudaError_t abs_sum_launch(const float* J, int64_t ld, int64_t rows, int K, double* out, cudaStream_t s);
// lo[i] = x[i] - trunc_tf32(x[i])  (exact; what tcgen05 kind::tf32 drops from x)
cudaError_t split_lo_launch(const float* x, float* lo, int64_t count, cudaStream_t s);
// h[i] = bf16(x[i]), round to nearest even (h: two-byte elements)
cudaError_t to_bf16_launch(const float* x, void* h, int64_t count, cudaStream_t s);
// momentum update that also refreshes the bf16 copy of the weights
cudaError_t momentum_update_bf16_launch(float* theta, float* vel, const float* grad, void* theta_h, int64_t count, float mu,
                                        float alpha, cudaStream_t stream);
// momentum update that also refreshes the low parts of the weights
cudaError_t momentum_update_lo_launch(float* theta, float* vel, const float* grad, float* theta_lo, int64_t count, float mu,
                                      float alpha, cudaStream_t stream);
// dst[f][b*bs_pad + r] = src[f][b*bs + r] for the nb contiguous batches of allocVarGPU's table (last absorbs the rest)
cudaError_t split_batches_launch(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int64_t bs,
                                 int64_t bs_pad, int nb, int n_rows, cudaStream_t s);
cudaError_t multiply_launch(const float* a, const float* b, float* out, int64_t count, cudaStream_t s);  // out may alias b
cudaError_t init_scale_launch(float* theta, int in, int out, int64_t count, int kind, cudaStream_t s);
// device-side normaliseData: per-feature mean / (m-1) standard deviation of an engine-layout matrix, and its application
cudaError_t feature_stats_launch(const float* x, int64_t ld, int64_t rows, int n, float* mean, float* stddev, cudaStream_t s);
cudaError_t normalise_rows_launch(float* x, int64_t ld, int64_t rows, int n, const float* mean, const float* stddev, cudaStream_t s);
cudaError_t labels_to_onehot_launch(const float* labels, float* y, int64_t ldy, int64_t rows, int K, cudaStream_t s);

// Fused skinny last layer (tail.cu): forward + output activation + delta + cost/argmax [+ delta backprop + gradient].
struct TailArgs {
    const float* a; int64_t lda;          // last hidden activation, (H+1) rows (ones row last)
    const float* W; int64_t ldw;          // K rows of H+1 weights (bias last)
    const float* Y; int64_t ldy;          // targets, K rows (may be null)
    float* delta_last; int64_t ldd;       // K rows (may be null)
    float* delta_prev; int64_t ldp;       // H rows (training)
    float* G_inline;                      // K rows of ldw: gradient of the last layer accumulated in-kernel (or null)
    float* h_out; int64_t ldh;            // activated output, K rows (may be null)
    float* pred;                          // argmax class per row as float (may be null)
    const float* y_labels;                // class ids for the error count (validate path, may be null)
    double* cost_sum; double* err_sum;    // double accumulators (may be null)
    int64_t rows; int H, K; int out_kind, cost_kind, act_kind;
};
bool tail_supported(int K, int H, bool want_inline_g);
cudaError_t tail_launch(const TailArgs& a, bool train, cudaStream_t s);

// Backward step of a narrow layer behind a wide one (skinny.cu): delta_prev = (delta W) .* f'(a) and G = delta^T [a,1]
// in one pass over a.  a, delta, delta_prev share the leading dimension ld (samples contiguous); the caller zeroes G.
struct SkinnyArgs {
    const float* a;            // (H+1) rows, ones row last
    const float* delta;        // K rows
    const float* W;            // K rows of ldw
    float* delta_prev;         // H rows
    void* delta_prev_bf16 = nullptr;   // optional bf16 copy of delta_prev (same indexing, two-byte elements)
    float* G;                  // K rows of ldw
    int64_t ld, ldw, rows;
    int H, K, act_kind;
};
bool skinny_bwd_supported(int K, int H, int64_t ld);
cudaError_t skinny_bwd_launch(const SkinnyArgs& a, cudaStream_t s);

// A whole step (forward, output, backward, gradients of every layer) of a network whose weights fit shared memory
// (micro.cu).  W / G are the internal arenas (layer l at w_off[l], out[l] rows of ldw[l]); X, Y point at the batch's first
// sample; the caller zeroes G.
constexpr int MICRO_MAX_LAYERS = 8;
constexpr int MICRO_PITCH = 36;        // floats per shared-memory activation row (act_off / act_total are in these units)
struct MicroArgs {
    const float* W; float* G;
    const float* X; int64_t ldx;
    const float* Y; int64_t ldy;
    double* cost_sum;              // may be null
    int64_t rows;
    int nw, total, act_total, max_width;
    int in[MICRO_MAX_LAYERS], out[MICRO_MAX_LAYERS], ldw[MICRO_MAX_LAYERS], w_off[MICRO_MAX_LAYERS], act_off[MICRO_MAX_LAYERS];
    int act_kind, out_kind, cost_kind;
};
bool micro_supported(const MicroArgs& a);
cudaError_t micro_launch(const MicroArgs& a, cudaStream_t s);

// One-hidden-layer network, whole step in one kernel (mlp3.cu).  X / Y point at the batch's first sample; `partial` holds
// one gradient arena (internal layout, `total` floats) per block; the reduce/update kernel sums them into the update.
struct Mlp3Args {
    const float* X; int64_t ldx;           // (l0+1) feature rows, ones row last
    const float* Y; int64_t ldy;           // K target rows (null: forward only without cost)
    const float* W0; int ldw0;             // H rows of ldw0 floats
    const float* W1; int ldw1;             // K rows of ldw1 floats
    float* partial; int total; int w1_off; // training: [grid][total] partial gradients; w1_off = offset of layer 1 in the arena
    float* h_out; int64_t ldh;             // activated output, K rows (may be null)
    float* pred;                           // argmax class per row as float (may be null)
    const float* y_labels;                 // class ids for the error count (validate path, may be null)
    double* cost_sum; double* err_sum;     // double accumulators (may be null)
    int64_t rows; int n1, H, K;
    int act_kind, out_kind, cost_kind;
    int x3;                                // fp32-grade arithmetic (hi/lo split in registers), else plain tf32
    int train;
};
bool mlp3_supported(int n1, int H, int K);
int mlp3_grid(int64_t rows);               // blocks mlp3_launch will use (= partial gradient slices)
cudaError_t mlp3_launch(const Mlp3Args& a, cudaStream_t s);
cudaError_t mlp3_reduce_update_launch(const float* partial, int nparts, int64_t total, float* theta, float* vel, float mu, float alpha,
                                      cudaStream_t s);

// ---- peer-memory data parallelism (p2p.cu) ---------------------------------------------------------------------------
constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_MAX_LAYERS = 64;
struct P2pPeers {
    float* W[P2P_MAX_WORLD];       // every rank's weight arena (own entry = local pointer)
    float* G[P2P_MAX_WORLD];       // every rank's gradient arena
    int*   flags[P2P_MAX_WORLD];   // every rank's flag block: [0, L) wg_done per (layer, src) ... see engine.cu
    float* S[P2P_MAX_WORLD];       // every rank's slot area for small layers (behind its flag block, same allocation)
    int world, rank;
};
constexpr int P2P_FLAG_INTS = 2 * P2P_MAX_LAYERS * P2P_MAX_WORLD;       // [kind][layer][rank]
cudaError_t p2p_gate_launch(const int* flags, int count, int value, cudaStream_t s);
// one launch for several layers: flags_kind = base of a [P2P_MAX_LAYERS][P2P_MAX_WORLD] block, bit j of mask = wait for layer j
cudaError_t p2p_gate_layers_launch(const int* flags_kind, uint64_t layer_mask, int world, int value, cudaStream_t s);
// small layers (rows do not split into whole tiles per rank): every rank copies its whole gradient into slot `rank` of
// every peer, then each rank sums the slots in rank order and updates its own replica (identical bits everywhere)
cudaError_t p2p_small_scatter_launch(const P2pPeers& peers, const float* src, int64_t slot_off, int64_t count, cudaStream_t s);
cudaError_t p2p_small_reduce_update_launch(const float* slots, int64_t slot_stride, int world, float* theta, float* vel,
                                           int64_t count, float mu, float alpha, cudaStream_t s);
cudaError_t p2p_signal_launch(const P2pPeers& peers, int index, int value, cudaStream_t s);
cudaError_t p2p_reduce_update_launch(const P2pPeers& peers, const float* partial, int64_t slice_floats, float* vel,
                                     int64_t w_off, int64_t count, float mu, float alpha, cudaStream_t s);

void set_error(const std::string& msg);

}  // namespace b2

#define B2_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            b2::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +  \
                          std::to_string(__LINE__) + ")");                                              \
            return B2NN_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)
