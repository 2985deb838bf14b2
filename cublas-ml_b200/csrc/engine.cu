// engine.cu — context, arenas, the train loop and the C ABI of include/b2nn.h.
//
// Engine layout ("unit-major"): every matrix is stored [unit][sample] or [out][in] with the LAST index contiguous.
//   X_int   (l_0+1) rows of ldx floats : feature rows 0..l_0-1, ones row at index l_0   (reference: ones COLUMN first)
//   a_j     (l_{j+1}+1) rows of ld     : f(z_j) rows, ones row last, written once at allocation
//   delta_j l_{j+1} rows of ld
//   W_j     l_{j+1} rows of ldw_j = roundup4(l_j+1) floats : input weights, bias weight at column l_j
//   V, G    same arena layout as W (velocity, weight gradient) — the reference's Delta is the transpose (F3); here the
//           gradient is produced directly in W's layout so the update is a pure streaming pass.
// Putting the bias unit LAST keeps every bias-free sub-matrix (W without bias column, a without ones row) at the
// 16-byte aligned allocation base, which TMA requires; the reference's "+mBatch" pointer bump (cublasNN.cpp:676) has
// no aligned equivalent with the bias first.
#include "engine.h"
#include "../../include/b2nn.h"

#include <curand.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace b2 {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

namespace {

inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

struct LayerInfo {
    int in = 0, out = 0;       // l_j, l_{j+1} (bias not counted)
    int64_t ldw = 0;           // roundup4(in + 1)
    int64_t w_off = 0;         // offset (floats) inside the internal arenas W / V / G
    int64_t ref_off = 0;       // thetaPos[j] of the reference arena (cublasNN.cpp:293-301)
    int64_t ref_size = 0;      // thetaSize[j]
};

// One contraction of the step, bound to a backend at plan time.
struct GemmOp {
    GemmDesc g;
    bool tc = false;
    bool zero_first = false;   // split-K accumulates atomically into a zeroed gradient block
    bool a_is_x = false;       // A is the staged training matrix: the batch start shifts its sample coordinate
    TcPlan plan;
};

// Everything one (forward, backward) pass over `rows` samples needs; cached per distinct batch size.
struct StepPlan {
    int64_t rows = 0;
    std::vector<GemmOp> fwd;     // L-1 entries
    std::vector<GemmOp> wgrad;   // L-1 entries
    std::vector<GemmOp> bwd;     // entry j computes delta_{j-1} from delta_j (index 0 unused)
    std::vector<char> fused;     // data parallel over peer memory: layer's gradient is scattered to its owners by the GEMM
    std::vector<char> fwd_waits; // ... and its forward tcgen05 kernel waits for the owners' weight rows itself (no gate launch)
    bool use_tail = false;       // skinny last layer: forward + output + delta backprop fused in tail.cu
    bool tail_inline_g = false;  // ... and its weight gradient too
    bool skinny_bwd = false;     // narrow last layer behind a wide hidden layer: delta backprop + gradient in one pass (skinny.cu)
    bool micro = false;          // the whole net fits shared memory: forward + backward + gradients in one kernel (micro.cu)
    bool mlp3 = false;           // one hidden layer (config C1): X tile resident in shared memory, whole step in one kernel (mlp3.cu)
    MicroArgs micro_args{};
};

struct Workspace {
    int64_t cap = 0, ld = 0;
    std::vector<float*> a;       // hidden activations (L-2 buffers)
    std::vector<float*> delta;   // L-1 buffers
    std::vector<float*> z;       // only for custom activations (L-2 buffers)
    float* z_last = nullptr;     // K rows
    float* scratch = nullptr;    // custom path: act'(z) / product
    std::vector<float*> a_lo, delta_lo;   // fp32-grade tensor-core mode: x - trunc_tf32(x) of a / delta (same shapes)
    float *hbuf = nullptr, *ybuf = nullptr, *jbuf = nullptr;   // custom output activation / cost: h, compact y, per-element J
    void release() {
        for (float* p : a) cudaFree(p);
        for (float* p : delta) cudaFree(p);
        for (float* p : z) cudaFree(p);
        for (float* p : a_lo) cudaFree(p);
        for (float* p : delta_lo) cudaFree(p);
        a_lo.clear(); delta_lo.clear();
        cudaFree(z_last); cudaFree(scratch); cudaFree(hbuf); cudaFree(ybuf); cudaFree(jbuf);
        hbuf = ybuf = jbuf = nullptr;
        a.clear(); delta.clear(); z.clear(); z_last = nullptr; scratch = nullptr; cap = ld = 0;
    }
};

// ---- NCCL, loaded lazily so the library has no link-time dependency on it ------------------------------------------
struct Id128 { char bytes[128]; };   // ncclUniqueId, passed by value
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;
std::string g_nccl_err;

void load_nccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { g_nccl_err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return; }
    auto sym = [&](const char* s) { return dlsym(g_nccl.handle, s); };
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(sym("ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(sym("ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(sym("ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
    g_nccl.GroupStart = reinterpret_cast<int (*)()>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<int (*)()>(sym("ncclGroupEnd"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) g_nccl_err = "NCCL symbols missing";
}
constexpr int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_INT64 = 4, NCCL_SUM = 0;

}  // namespace
}  // namespace b2

using namespace b2;

struct b2nn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;
    curandGenerator_t gen = nullptr;

    std::vector<int> layers;
    std::vector<LayerInfo> L;
    int64_t total_ref = 0, total_int = 0;
    float *W = nullptr, *V = nullptr, *G = nullptr;
    float* W_lo = nullptr; bool w_lo_valid = false;      // companion of the weights: low parts (fp32-grade mode) or bf16 copies
    bool comp_bf16 = false;                              // which of the two the companion buffers currently hold
    std::map<const float*, std::pair<float*, int64_t>> lo_cache;   // low-part companions of input matrices, by base pointer

    // training set, engine layout
    float* X = nullptr; int64_t ldx = 0;
    float* Y = nullptr; int64_t ldy = 0;
    int64_t m = 0; int n = 0, K = 0;

    // batch-aligned copy of the training set (splitData, cublasNN.cpp:498-524): every batch starts on a 32-sample
    // boundary so MN-major TMA boxes can address it as whole 32-float atoms.  Built per train call when needed.
    float* Xsp = nullptr; float* Ysp = nullptr; int64_t ldsp = 0; int sp_nb = 0; int64_t sp_stride = 0;

    // double-buffered staging of b2nn_train_host (streaming trainer)
    float* Xst[2] = {nullptr, nullptr}; float* Yst[2] = {nullptr, nullptr}; int64_t ldst = 0, capst = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t loss_stream = nullptr;        // per-step loss read-back of b2nn_train_host (keeps the DMA off the compute stream)
    cudaEvent_t st_copied[2] = {nullptr, nullptr}, st_consumed[2] = {nullptr, nullptr}, st_loss = nullptr;
    double* d_jsteps = nullptr; double* h_jsteps = nullptr; int64_t jsteps_cap = 0;

    // staging of b2nn_predict (kept between calls: small-batch inference is otherwise dominated by cudaMalloc/cudaFree)
    // Two of everything: chunk c+1 is packed into pinned memory and travels H2D while chunk c computes and chunk c-1's
    // results travel back (b2nn_predict).
    float* pr_x[2] = {nullptr, nullptr}; float* pr_p[2] = {nullptr, nullptr}; float* pr_h[2] = {nullptr, nullptr};
    int64_t pr_ld = 0; int pr_n = 0, pr_k = 0;
    float* pr_hx[2] = {nullptr, nullptr};      // pinned host staging of one input chunk (n rows of pr_ld floats)
    float* pr_hout[2] = {nullptr, nullptr};    // pinned host staging of one chunk's results (pred: pr_ld, probs: K * pr_ld)
    cudaEvent_t pr_copied[2] = {nullptr, nullptr}, pr_computed[2] = {nullptr, nullptr}, pr_out[2] = {nullptr, nullptr};
    cudaStream_t pr_d2h_stream = nullptr;

    // host-step staging (b2nn_train_step_host)
    float* Xs = nullptr; float* Ys = nullptr; int64_t lds = 0, caps = 0;

    // device-side normaliseData (cublasNN.cpp:191-268): per-feature mean / stddev of the training set, applied to raw rows
    // while they are staged (b2nn_set_train_data_raw, b2nn_evaluate_raw, b2nn_predict_raw)
    float* d_mean = nullptr; float* d_std = nullptr; int norm_n = 0;

    int* out_counters = nullptr; int out_counters_cap = 0;      // fused output layer: one self-resetting counter per 32-sample slice
    float* m3_partial = nullptr; int64_t m3_partial_cap = 0;    // per-block partial gradients of the one-hidden-layer kernel (mlp3.cu)

    Workspace ws;
    std::map<std::pair<int64_t, const float*>, StepPlan> plans;   // keyed by (rows, X base); valid for the (ws, ldx, precision, ...) below
    int plans_prec = -1;
    int64_t plans_ldx = 0, plans_xtotal = 0;
    bool plans_xaligned = false, plans_notail = false;
    int plans_act = -1;

    double* d_acc = nullptr;             // device accumulators: [0] cost, [1] errors, [2..] display slots
    int acc_slots = 0;
    double* h_acc = nullptr;             // pinned mirror
    std::vector<cudaEvent_t> log_events;

    // data parallel
    void* comm = nullptr; int rank = 0, nranks = 1;
    std::vector<cudaEvent_t> grad_ready, grad_reduced;
    // ... over NVLink peer memory (p2p.cu): fused reduce-scatter / update / all-gather for the large layers
    bool p2p = false;
    P2pPeers peers{};
    int* p2p_flags = nullptr;                 // local flag block: [kind][layer][rank], kind 0 = wg_done, 1 = w_done
    int p2p_step = 0;
    std::vector<int> grad_order;              // layers in the order this step produced their gradients
    bool g_zero = false;                      // the whole gradient arena is known to be zero (small nets: the update leaves it so)
    std::vector<int64_t> p2p_small_off;       // per layer: float offset of its slot block in the small-layer area, -1 = none
    int64_t p2p_small_floats = 0;
    std::vector<int64_t> p2p_small_cnt;       // floats per slot the area was sized for (guards against a later set_layers)
    cudaStream_t upd_stream = nullptr;
    std::vector<cudaEvent_t> p2p_wg_ev;
    cudaEvent_t p2p_upd_done = nullptr;
    std::vector<void*> p2p_opened;

    // single GPU, large nets: per-layer momentum updates on the exchange stream, under the remaining backward GEMMs
    bool ov_active = false; float ov_mu = 0.f, ov_alpha = 0.f; int ov_prec = 0;
    std::vector<cudaEvent_t> ov_ev; cudaEvent_t ov_done = nullptr;

    int64_t launches = 0, tc_launches = 0;

    // B2NN_TRACE=1: CUDA-event timeline of the last steps of a train call (both streams), printed to stderr when it ends
    bool trace = false; int64_t trace_step = 0, trace_from = 0;
    std::vector<std::pair<std::string, cudaEvent_t>> trace_ev;

    // optional per-launch CUDA-event timing (b2nn_profile_*): kind 0 forward, 1 delta backprop, 2 weight gradient,
    // 3 output layer, 4 momentum update
    bool profile = false;
    struct ProfRec { int kind; bool tc; double flops; double bytes; cudaEvent_t e0, e1; bool own_e0; };
    bool prof_chain = false;           // the previous scope's end event can serve as this scope's start
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
};

namespace {

int check_device(int device, std::string& why) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        (void)cudaGetLastError();
        why = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return B2NN_ERR_NO_DEVICE;
    }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= n) { why = "device index out of range"; return B2NN_ERR_ARG; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { why = "cudaGetDeviceProperties failed"; return B2NN_ERR_CUDA; }
    if (prop.major != 10) {
        why = "device is compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor) +
              "; this library only contains sm_100a code";
        return B2NN_ERR_NO_DEVICE;
    }
    return B2NN_OK;
}

int resolve_precision(int prec) {
    if (prec == B2NN_PREC_FP32 || prec == B2NN_PREC_TF32 || prec == B2NN_PREC_TF32X3 || prec == B2NN_PREC_BF16) return prec;
    const char* e = std::getenv("B2NN_PRECISION");
    if (e && (std::strcmp(e, "fp32") == 0 || std::strcmp(e, "FP32") == 0)) return B2NN_PREC_FP32;
    if (e && (std::strcmp(e, "tf32") == 0 || std::strcmp(e, "TF32") == 0)) return B2NN_PREC_TF32;
    if (e && (std::strcmp(e, "bf16") == 0 || std::strcmp(e, "BF16") == 0)) return B2NN_PREC_BF16;
    // Default = fp32-grade (3xTF32 on the tensor cores): a drop-in for the reference's default-math cublasSgemm
    // (cublasNN.cpp:25 never changes the math mode) must not silently narrow the arithmetic; plain tf32 is an opt-in.
    return B2NN_PREC_TF32X3;
}

// Tensor-core modes whose contractions read a COMPANION of every operand: the fp32-grade mode its low part
// x - trunc_tf32(x) (fp32), the bf16 mode its bf16 copy (two-byte elements at the start of the same buffers).
inline bool is_tc_prec(int prec) { return prec == B2NN_PREC_TF32 || prec == B2NN_PREC_TF32X3 || prec == B2NN_PREC_BF16; }
inline bool is_comp_prec(int prec) { return prec == B2NN_PREC_TF32X3 || prec == B2NN_PREC_BF16; }
// companion of element `off` of an array whose companion buffer starts at `base`
inline float* comp_at(float* base, int64_t off, bool bf) {
    return bf ? reinterpret_cast<float*>(reinterpret_cast<uint16_t*>(base) + off) : base + off;
}
inline cudaError_t make_companion(const float* x, float* comp, int64_t count, bool bf, cudaStream_t s) {
    return bf ? to_bf16_launch(x, comp, count, s) : split_lo_launch(x, comp, count, s);
}

// Companion buffer of an input matrix (allocated on demand, content refreshed by the caller).
float* lo_buffer(b2nn_ctx* c, const float* x, int64_t count) {
    auto it = c->lo_cache.find(x);
    if (it != c->lo_cache.end() && it->second.second >= count) return it->second.first;
    if (it != c->lo_cache.end()) { cudaFree(it->second.first); c->lo_cache.erase(it); }
    float* lo = nullptr;
    if (cudaMalloc(&lo, sizeof(float) * size_t(count)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    c->lo_cache[x] = std::make_pair(lo, count);
    c->plans.clear();            // plans embed tensor maps over the low-part buffers
    return lo;
}
void drop_lo(b2nn_ctx* c, const float* x) {
    auto it = c->lo_cache.find(x);
    if (it != c->lo_cache.end()) { cudaFree(it->second.first); c->lo_cache.erase(it); c->plans.clear(); }
}
const float* find_lo(const b2nn_ctx* c, const float* x) {
    auto it = c->lo_cache.find(x);
    return it == c->lo_cache.end() ? nullptr : it->second.first;
}

int ensure_w_lo(b2nn_ctx* c) {
    if (!c->W_lo) {
        B2_CUDA(cudaMalloc(&c->W_lo, sizeof(float) * (c->total_int + 64)));
        B2_CUDA(cudaMemsetAsync(c->W_lo, 0, sizeof(float) * (c->total_int + 64), c->stream));
        c->w_lo_valid = false;
    }
    if (!c->w_lo_valid) {
        B2_CUDA(make_companion(c->W, c->W_lo, c->total_int, c->comp_bf16, c->stream));
        c->w_lo_valid = true;
    }
    return B2NN_OK;
}

// Scope guards: every error exit of the entry points below releases what it created.
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};
struct GraphExecs {
    cudaGraphExec_t g[2] = {nullptr, nullptr};
    void destroy() { for (auto& e : g) if (e) { cudaGraphExecDestroy(e); e = nullptr; } }
    ~GraphExecs() { destroy(); }
};

int epi_forward(int act_kind) { return act_kind == B2NN_ACT_SIGMOID ? EPI_SIGMOID : act_kind == B2NN_ACT_TANH ? EPI_TANH : EPI_STORE; }
int epi_backward(int act_kind) { return act_kind == B2NN_ACT_SIGMOID ? EPI_DSIGMOID : act_kind == B2NN_ACT_TANH ? EPI_DTANH : EPI_STORE; }

int ensure_acc(b2nn_ctx* c, int slots) {
    if (slots <= c->acc_slots) return B2NN_OK;
    if (c->d_acc) cudaFree(c->d_acc);
    if (c->h_acc) cudaFreeHost(c->h_acc);
    c->d_acc = nullptr; c->h_acc = nullptr; c->acc_slots = 0;
    B2_CUDA(cudaMalloc(&c->d_acc, sizeof(double) * slots));
    B2_CUDA(cudaMallocHost(&c->h_acc, sizeof(double) * slots));
    c->acc_slots = slots;
    return B2NN_OK;
}

static int ensure_workspace_impl(b2nn_ctx* c, int64_t rows, bool custom, bool want_lo);
int ensure_workspace(b2nn_ctx* c, int64_t rows, bool custom, bool want_lo = false) {
    const int rc = ensure_workspace_impl(c, rows, custom, want_lo);
    if (rc) { c->ws.release(); c->plans.clear(); }     // a half-built workspace must not claim its capacity
    return rc;
}
static int ensure_workspace_impl(b2nn_ctx* c, int64_t rows, bool custom, bool want_lo) {
    Workspace& w = c->ws;
    const int nl = int(c->layers.size());
    const bool need_z = custom && w.z.empty();
    const bool need_lo = want_lo && w.a_lo.empty();
    if (rows <= w.cap && !need_z && !need_lo) return B2NN_OK;
    if (rows > w.cap) {
        w.release();
        c->plans.clear();
        w.cap = rows;
        w.ld = round_up(rows, 32);
        w.a.assign(nl - 2, nullptr);
        w.delta.assign(nl - 1, nullptr);
        for (int j = 0; j < nl - 2; ++j) {
            const int64_t units = c->layers[j + 1];
            B2_CUDA(cudaMalloc(&w.a[j], sizeof(float) * (units + 1) * w.ld));
            B2_CUDA(cudaMemsetAsync(w.a[j], 0, sizeof(float) * units * w.ld, c->stream));
            B2_CUDA(fill_launch(w.a[j] + units * w.ld, 1.0f, w.ld, c->stream));      // addBiasMatGPU, once (kernels.cu:67-72)
        }
        for (int j = 0; j < nl - 1; ++j) {
            B2_CUDA(cudaMalloc(&w.delta[j], sizeof(float) * int64_t(c->layers[j + 1]) * w.ld));
            B2_CUDA(cudaMemsetAsync(w.delta[j], 0, sizeof(float) * int64_t(c->layers[j + 1]) * w.ld, c->stream));
        }
        B2_CUDA(cudaMalloc(&w.z_last, sizeof(float) * int64_t(c->layers[nl - 1]) * w.ld));
    }
    if (custom && w.z.empty()) {
        w.z.assign(nl - 2, nullptr);
        int64_t zmax = 0;
        for (int j = 0; j < nl - 2; ++j) {
            const int64_t cnt = int64_t(c->layers[j + 1]) * w.ld;
            B2_CUDA(cudaMalloc(&w.z[j], sizeof(float) * cnt));
            zmax = std::max(zmax, cnt);
        }
        B2_CUDA(cudaMalloc(&w.scratch, sizeof(float) * zmax));
    }
    if (want_lo && w.a_lo.empty()) {
        c->plans.clear();
        w.a_lo.assign(nl - 2, nullptr);
        w.delta_lo.assign(nl - 1, nullptr);
        for (int j = 0; j < nl - 2; ++j) {
            const size_t bytes = sizeof(float) * size_t(c->layers[j + 1] + 1) * w.ld;       // low part of the ones row is 0
            B2_CUDA(cudaMalloc(&w.a_lo[j], bytes));
            B2_CUDA(cudaMemsetAsync(w.a_lo[j], 0, bytes, c->stream));
        }
        for (int j = 0; j < nl - 1; ++j) {
            const size_t bytes = sizeof(float) * size_t(c->layers[j + 1]) * w.ld;
            B2_CUDA(cudaMalloc(&w.delta_lo[j], bytes));
            B2_CUDA(cudaMemsetAsync(w.delta_lo[j], 0, bytes, c->stream));
        }
    }
    return B2NN_OK;
}

// Everything the fp32-grade tensor-core mode needs around one input matrix X (count floats): workspace low parts,
// weight low parts, and X's own low part, freshly split.
int prepare_x3(b2nn_ctx* c, const float* X, int64_t count, int64_t rows, bool custom, bool bf = false) {
    if (c->comp_bf16 != bf) {            // the companion buffers change their meaning: nothing in them is valid any more
        c->comp_bf16 = bf;
        c->w_lo_valid = false;
        c->plans.clear();
    }
    int rc = ensure_workspace(c, std::max<int64_t>(rows, c->ws.cap), custom, true);
    if (rc) return rc;
    // the ones row of every hidden activation is written once and never by a contraction: its companion (0 as a low
    // part, bf16(1) as a copy) is set here
    for (size_t j = 0; j < c->ws.a.size() && j < c->ws.a_lo.size(); ++j) {
        const int64_t off = int64_t(c->layers[j + 1]) * c->ws.ld;
        B2_CUDA(make_companion(c->ws.a[j] + off, comp_at(c->ws.a_lo[j], off, bf), c->ws.ld, bf, c->stream));
    }
    rc = ensure_w_lo(c);
    if (rc) return rc;
    float* lo = lo_buffer(c, X, count);
    if (!lo) { set_error("cudaMalloc Failed"); return B2NN_ERR_ALLOC; }
    B2_CUDA(make_companion(X, lo, count, bf, c->stream));
    return B2NN_OK;
}

constexpr double kTcMinWork = double(1 << 26);     // multiply-adds below which a contraction stays on the fp32 FFMA kernel

// A tensor-core mode on a network none of whose contractions would reach the tensor cores (config C2's bikeshare net) is
// plain fp32: skip the low-part bookkeeping of the fp32-grade mode altogether.
int effective_precision(const b2nn_ctx* c, int prec, int64_t rows) {
    if (!is_tc_prec(prec)) return prec;
    for (size_t j = 0; j < c->L.size(); ++j) {
        const LayerInfo& li = c->L[j];
        const double work = double(rows) * double(li.in + 1) * double(li.out);
        if (work >= kTcMinWork || (li.in + 1 >= 512 && work >= double(1 << 15))) return prec;
    }
    return B2NN_PREC_FP32;
}

// Decide the backend of one contraction and, for tcgen05, build its tensor maps.
void bind_backend(b2nn_ctx* c, GemmOp& op, int prec, bool allow_split) {
    GemmDesc& g = op.g;
    op.tc = false; op.zero_first = false; g.split_k = 1;
    const double work = double(g.M) * double(g.N) * double(g.K);
    if (allow_split) {
        // Weight gradient: K = batch rows is long while M x N can be a handful of tiles (50 x 785, 10 x 4097).  Split K so
        // the grid covers the 148 SMs; partial sums meet in a zeroed gradient block via red.global.add.f32.
        const bool tc_mode = is_tc_prec(prec);
        const int tile_m = tc_mode ? 128 : 64;
        const int tile_n = tc_mode ? (g.N > 128 ? 256 : g.N > 64 ? 128 : g.N > 32 ? 64 : g.N > 16 ? 32 : 16) : 64;
        const int64_t tiles = int64_t((g.M + tile_m - 1) / tile_m) * ((g.N + tile_n - 1) / tile_n);
        const int kblk = (g.K + 31) / 32;
        if (tiles < 74 && kblk >= 16) {
            int s = int(std::min<int64_t>(kblk / 8, (296 + tiles - 1) / tiles));
            if (s > 1) { g.split_k = s; op.zero_first = true; }
        }
    }
    // tcgen05 only where tensor cores pay: below ~2^26 multiply-adds a contraction is launch / HBM bound and stays on
    // the exact fp32 FFMA kernel (the bikeshare-sized nets of config C2 never leave fp32).
    // ... except long-K contractions with few rows (small-batch inference on a wide input, config C5): the FFMA kernel
    // would walk K serially in a single block, the tcgen05 pipeline streams it in a few microseconds.
    const bool long_k_few_rows = !allow_split && g.a_major == 1 && g.b_major == 0 && g.K >= 512 && work >= double(1 << 15);
    if (is_tc_prec(prec) && (work >= kTcMinWork || long_k_few_rows)) {
        std::string err;
        const int requested = g.split_k;
        if (g.epi == EPI_STORE && (allow_split || g.N <= 32)) g.split_k = 0;     // let the tcgen05 plan pick the split
        if (tc_plan_build(op.plan, g, err)) {
            op.tc = true;
            op.zero_first = op.plan.splits > 1;
        } else {
            g.split_k = requested;
        }
    }
    if (!op.tc && g.split_k > 1) {
        // FFMA tiles are 64 x 64: re-derive the split for that grid
        const int64_t tiles = int64_t((g.M + 63) / 64) * ((g.N + 63) / 64);
        const int kblk = (g.K + 15) / 16;
        int s = (tiles < 148 && kblk >= 16) ? int(std::min<int64_t>(kblk / 8, (296 + tiles - 1) / tiles)) : 1;
        g.split_k = std::max(1, s);
        op.zero_first = g.split_k > 1;
    }
}

// One-hidden-layer nets in a tensor-core mode on a single GPU with built-in functions CAN take the one-kernel step of
// mlp3.cu (B2NN_MLP3=1).  Opt-in: measured on config C1 it is no faster than the four-launch tcgen05 path (38-41 vs 37 us
// per step) — its contractions run on the legacy mma.sync path, whose tf32 rate on sm_100 is ~136 multiply-adds per
// clock per SM (ncu: HMMA sub-pipe busy 21 K cycles for 2 800 m16n8k8 per SM), one sixteenth of tcgen05's.
bool mlp3_ok(const b2nn_ctx* c, int prec, int act_kind) {
    const char* e = std::getenv("B2NN_MLP3");
    return e && std::atoi(e) != 0 && c->L.size() == 2 && !c->comm && act_kind != B2NN_ACT_CUSTOM && (prec == B2NN_PREC_TF32 || prec == B2NN_PREC_TF32X3) &&
           c->total_int > 16384 && (c->total_int & 3) == 0 && mlp3_supported(c->L[0].in + 1, c->L[0].out, c->L[1].out);
}

// x_total: number of sample columns of X that hold data.  The tensor maps over X span all of them so one plan serves
// every batch of the same size; the batch start only moves the TMA coordinates (run_gemm).
// x_aligned: every batch start passed to run_gemm is a multiple of 32 samples (one 3-D TMA box per X tile).
StepPlan& get_plan(b2nn_ctx* c, int64_t rows, const float* X, int64_t ldx, int64_t x_total, bool x_aligned, int prec,
                   int act_kind, bool no_tail_arg = false) {
    if (c->plans_prec != prec || c->plans_ldx != ldx || c->plans_act != act_kind || c->plans_xtotal != x_total ||
        c->plans_xaligned != x_aligned || c->plans_notail != no_tail_arg) {
        c->plans.clear();
        c->plans_prec = prec; c->plans_ldx = ldx; c->plans_act = act_kind; c->plans_xtotal = x_total;
        c->plans_xaligned = x_aligned; c->plans_notail = no_tail_arg;
    }
    const auto key = std::make_pair(rows, X);
    auto it = c->plans.find(key);
    if (it != c->plans.end()) return it->second;
    if (c->plans.size() > 64) c->plans.clear();        // bounded: a few batch sizes x a few input buffers
    StepPlan& sp = c->plans[key];
    sp.rows = rows;
    const int nl = int(c->layers.size());
    const int nw = nl - 1;
    Workspace& w = c->ws;
    const bool custom = act_kind == B2NN_ACT_CUSTOM;
    // x3: the contractions carry companions of their operands (low parts, or bf16 copies when bf)
    const bool bf = prec == B2NN_PREC_BF16;
    const bool x3 = is_comp_prec(prec) && !custom && !w.a_lo.empty() && c->W_lo != nullptr && c->comp_bf16 == bf;
    const float* X_lo = x3 ? find_lo(c, X) : nullptr;
    sp.fwd.resize(nw); sp.wgrad.resize(nw); sp.bwd.resize(nw);
    sp.fused.assign(nw, 0);
    sp.fwd_waits.assign(nw, 0);
    {
        const LayerInfo& last = c->L[nw - 1];
        const bool no_tail = no_tail_arg || std::getenv("B2NN_NO_TAIL") != nullptr;
        sp.use_tail = !custom && !no_tail && tail_supported(last.out, last.in, false);
        sp.tail_inline_g = sp.use_tail && tail_supported(last.out, last.in, true);
        sp.skinny_bwd = !sp.use_tail && !custom && nw >= 2 && skinny_bwd_supported(last.out, last.in, w.ld);
        // tiny nets (every weight in shared memory, single GPU, built-in functions): one kernel for the whole step
        if (!custom && !no_tail && !c->comm && nw <= MICRO_MAX_LAYERS && c->total_int <= 16384) {
            MicroArgs& m = sp.micro_args;
            m = MicroArgs{};
            m.W = c->W; m.G = c->G; m.nw = nw; m.total = int(c->total_int);
            int off = 0, width = 0;
            for (int j = 0; j < nw; ++j) {
                m.in[j] = c->L[j].in; m.out[j] = c->L[j].out; m.ldw[j] = int(c->L[j].ldw); m.w_off[j] = int(c->L[j].w_off);
                m.act_off[j] = off;
                off += (c->L[j].in + 1) * MICRO_PITCH;
                width = std::max(width, c->L[j].out);
            }
            m.act_total = off; m.max_width = width; m.act_kind = act_kind;
            sp.micro = micro_supported(m);
        }
        sp.mlp3 = !no_tail && !sp.micro && mlp3_ok(c, prec, act_kind);
    }
    for (int j = 0; j < nw; ++j) {
        const LayerInfo& li = c->L[j];
        const bool tail_layer = sp.use_tail && j == nw - 1;
        // forward  z_j = [a_{j-1}, 1] W_j^T     (cublasNN.cpp:645-646, 653-654, 658-659)
        GemmOp& f = sp.fwd[j];
        f.g = GemmDesc();
        // (bf16 boxes are 64 samples wide while batch starts are only 32-aligned: X then takes one 2-D box per 64 samples)
        if (j == 0) { f.g.A = X; f.g.lda = ldx; f.a_is_x = true; f.g.a_mn_end = x_total; f.g.mn_slack = x_aligned && !bf; }
        else { f.g.A = w.a[j - 1]; f.g.lda = w.ld; f.g.mn_slack = true; }       // workspace rows are padded to 32 samples
        f.g.a_major = 1;
        f.g.B = c->W + li.w_off; f.g.ldb = li.ldw; f.g.b_major = 0;
        f.g.M = int(rows); f.g.N = li.out; f.g.K = li.in + 1;
        if (j < nw - 1) {
            f.g.D = custom ? w.z[j] : w.a[j]; f.g.ldd = w.ld; f.g.epi = custom ? EPI_STORE : epi_forward(act_kind);
        } else {
            f.g.D = w.z_last; f.g.ldd = w.ld; f.g.epi = EPI_STORE;
        }
        if (x3) {
            f.g.A_lo = (j == 0) ? X_lo : w.a_lo[j - 1];
            f.g.B_lo = comp_at(c->W_lo, li.w_off, bf);
            f.g.D_lo = (j < nw - 1) ? w.a_lo[j] : nullptr;
            f.g.bf16 = bf;
            if (!f.g.A_lo) { f.g.B_lo = nullptr; f.g.D_lo = nullptr; }
        }
        if (!tail_layer) bind_backend(c, f, prec, false);

        // weight gradient  G_j = delta_j^T [a_{j-1}, 1]  in W_j's layout   (cublasNN.cpp:680-684)
        GemmOp& g = sp.wgrad[j];
        g.g = GemmDesc();
        // Past the batch end X holds the next batch's samples, not zeros; the products vanish anyway because delta_j's
        // tensor map ends exactly at `rows` and TMA zero-fills it beyond.
        if (j == 0) { g.g.A = X; g.g.lda = ldx; g.a_is_x = true; g.g.a_k_end = x_total; } else { g.g.A = w.a[j - 1]; g.g.lda = w.ld; }
        g.g.a_major = 0;
        g.g.B = w.delta[j]; g.g.ldb = w.ld; g.g.b_major = 0;
        g.g.D = c->G + li.w_off; g.g.ldd = li.ldw;
        g.g.M = li.in + 1; g.g.N = li.out; g.g.K = int(rows);
        g.g.epi = EPI_STORE;
        if (x3) {
            g.g.A_lo = (j == 0) ? X_lo : w.a_lo[j - 1];
            g.g.B_lo = w.delta_lo[j];
            g.g.bf16 = bf;
            if (!g.g.A_lo) g.g.B_lo = nullptr;
        }
        // Peer-memory exchange: the layer's rows split evenly over the ranks in whole 256-column tiles of the gradient GEMM
        const bool want_fused = c->p2p && (prec == B2NN_PREC_TF32 || x3) && !custom && !tail_layer && nw <= P2P_MAX_LAYERS &&
                                li.out % (256 * c->nranks) == 0;
        if (want_fused) {
            // B2NN_P2P_FAKE_LOCAL (timing experiments only, results are wrong): every tile goes to local memory
            static const bool fake_local = std::getenv("B2NN_P2P_FAKE_LOCAL") != nullptr;
            for (int o = 0; o < c->nranks; ++o) g.g.scatter_base[o] = (fake_local ? c->G : c->peers.G[o]) + li.w_off;
            g.g.scatter_world = c->nranks; g.g.scatter_rank = c->rank; g.g.scatter_rows = li.out / c->nranks;
            bind_backend(c, g, prec, false);
            if (g.tc) sp.fused[j] = 1;
            else { g.g.scatter_world = 0; bind_backend(c, g, prec, true); }
        } else if (!(tail_layer && sp.tail_inline_g)) bind_backend(c, g, prec, true);
        // ... layers whose rows do not split into whole tiles per rank (the 10-unit output layer) exchange through slot
        // buffers in peer memory instead of NCCL: no foreign kernel competes with the persistent GEMMs for SMs
        if (c->p2p && !sp.fused[j] && !custom && (!is_comp_prec(prec) || x3) && nw <= P2P_MAX_LAYERS &&
            j < int(c->p2p_small_off.size()) && c->p2p_small_off[j] >= 0 && c->p2p_small_cnt[j] == li.ldw * li.out) sp.fused[j] = 2;
        // 1: a one-warp gate kernel in front of the layer's forward contraction; 2: the wait inside the tcgen05 kernel's TMA
        // producer (B2NN_P2P_FWD_WAIT=1; only safe when the owners' kernels can become resident next to the GEMM CTAs:
        // a GEMM that holds every SM while it spins on a flag whose writer needs an SM deadlocks the pair of GPUs)
        sp.fwd_waits[j] = (sp.fused[j] == 1 && prec == B2NN_PREC_TF32 && !tail_layer) ? ((f.tc && std::getenv("B2NN_P2P_FWD_WAIT")) ? 2 : 1) : 0;

        // delta backprop  delta_{j-1} = (delta_j W_j)[:, no bias] .* f'(z_{j-1})   (cublasNN.cpp:673-676)
        if (j > 0) {
            GemmOp& b = sp.bwd[j];
            b.g = GemmDesc();
            b.g.A = w.delta[j]; b.g.lda = w.ld; b.g.a_major = 1;
            b.g.mn_slack = true;      // delta rows are padded to 32 samples; the weight arena carries 64 floats of slack
            b.g.B = c->W + li.w_off; b.g.ldb = li.ldw; b.g.b_major = 1;
            b.g.M = int(rows); b.g.N = li.in; b.g.K = li.out;
            if (custom) { b.g.D = w.scratch; b.g.ldd = w.ld; b.g.epi = EPI_STORE; }
            else { b.g.D = w.delta[j - 1]; b.g.ldd = w.ld; b.g.epi = epi_backward(act_kind); b.g.aux = w.a[j - 1]; b.g.ldaux = w.ld; }
            if (x3 && !custom) { b.g.A_lo = w.delta_lo[j]; b.g.B_lo = comp_at(c->W_lo, li.w_off, bf); b.g.D_lo = w.delta_lo[j - 1]; b.g.bf16 = bf; }
            if (!tail_layer) bind_backend(c, b, prec, false);
        }
    }
    return sp;
}

void trace_mark(b2nn_ctx* c, const char* name, int layer, cudaStream_t s) {
    if (!c->trace || c->trace_step < c->trace_from) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    c->trace_ev.emplace_back(std::string(name) + (layer >= 0 ? std::to_string(layer) : "") + "@" + std::to_string(c->trace_step), e);
}
void trace_dump(b2nn_ctx* c) {
    if (c->trace_ev.empty()) return;
    cudaDeviceSynchronize();
    std::string out = "[b2nn trace rank " + std::to_string(c->rank) + "] us since first mark:";
    for (auto& kv : c->trace_ev) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->trace_ev.front().second, kv.second);
        char buf[96];
        std::snprintf(buf, sizeof(buf), " %s=%.1f", kv.first.c_str(), ms * 1e3f);
        out += buf;
    }
    std::fprintf(stderr, "%s\n", out.c_str());
    for (auto& kv : c->trace_ev) cudaEventDestroy(kv.second);
    c->trace_ev.clear();
}

cudaEvent_t prof_event(b2nn_ctx* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    b2nn_ctx* c; bool on;
    ProfScope(b2nn_ctx* ctx, int kind, bool tc, double flops, double bytes) : c(ctx), on(ctx->profile) {
        if (!on) return;
        // Back-to-back scopes on the engine stream share one event (end of the previous = start of this one): a timing
        // event drains the stream, so two per launch inflated every interval by several microseconds.
        const bool chain = c->prof_chain && !c->prof.empty();
        b2nn_ctx::ProfRec r{kind, tc, flops, bytes, chain ? c->prof.back().e1 : prof_event(c), prof_event(c), !chain};
        if (!chain) cudaEventRecord(r.e0, c->stream);
        c->prof.push_back(r);
    }
    ~ProfScope() { if (on) { cudaEventRecord(c->prof.back().e1, c->stream); c->prof_chain = true; } }
};

int run_gemm(b2nn_ctx* c, const GemmOp& op, int64_t x_start, int kind, const TcWait* wait = nullptr, const TcOutput* out = nullptr) {
    ProfScope scope(c, kind, op.tc, 2.0 * double(op.g.M) * double(op.g.N) * double(op.g.K), 0.0);
    const bool d_is_grad = op.g.D >= c->G && op.g.D < c->G + c->total_int;
    if (op.zero_first && !(d_is_grad && c->g_zero))
        B2_CUDA(cudaMemsetAsync(op.g.D, 0, sizeof(float) * size_t(op.g.ldd) * op.g.N, c->stream));
    if (op.tc) {
        // tensor maps span the whole training matrix; a batch start only moves the TMA coordinates
        B2_CUDA(tc_plan_launch(op.plan, c->stream, wait, op.a_is_x ? x_start : 0, out));
        ++c->tc_launches;
    } else {
        GemmDesc g = op.g;
        if (op.a_is_x) { if (g.a_major) g.a_mn0 += x_start; else g.a_k0 += x_start; }
        B2_CUDA(gemm_simt_launch(g, c->stream));
        // fp32-grade mode: a later tensor-core contraction may consume this output, so it needs its low part too
        if (g.D_lo) B2_CUDA(make_companion(g.D, g.D_lo, int64_t(g.N) * g.ldd, g.bf16, c->stream));
    }
    ++c->launches;
    return B2NN_OK;
}

// forwardPropagate (cublasNN.cpp:642-661) over rows [x_start, x_start+rows) of X; leaves z_{L-2} in ws.z_last.
int forward_pass(b2nn_ctx* c, StepPlan& sp, int64_t x_start, const b2nn_train_desc* d, int gate_step = 0, const TcOutput* out = nullptr) {
    const int nw = int(sp.fwd.size()) - (sp.use_tail ? 1 : 0);      // the tail kernel owns the last layer
    const bool custom = d->act_kind == B2NN_ACT_CUSTOM;
    for (int j = 0; j < nw; ++j) {
        // data parallel over peer memory: this layer's weights are complete once every owner's flag shows the previous step
        TcWait wait{c->p2p_flags + (1 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD, c->nranks, gate_step};
        if (gate_step > 0 && sp.fwd_waits[j] == 1)
            B2_CUDA(p2p_gate_launch(wait.flags, wait.count, wait.value, c->stream));
        int rc = run_gemm(c, sp.fwd[j], x_start, 0, (gate_step > 0 && sp.fwd_waits[j] == 2) ? &wait : nullptr,
                          j == int(sp.fwd.size()) - 1 ? out : nullptr);      // the output layer rides in the last contraction's epilogue
        if (rc) return rc;
        if (custom && j < int(sp.fwd.size()) - 1) {
            // User activation: a host function that launches on the legacy default stream (cublasNN.cpp:648, 655).  The
            // engine stream is a blocking stream, so legacy-stream work is ordered after and before it implicitly.
            const int64_t cnt = int64_t(c->layers[j + 1]) * c->ws.ld;
            if (cnt > INT32_MAX) { set_error("custom activation: element count exceeds int"); return B2NN_ERR_ARG; }
            d->act_fn(c->ws.z[j], c->ws.a[j], int(cnt));
        }
    }
    return B2NN_OK;
}

// backwardPropagate (cublasNN.cpp:663-685), re-ordered so G_j is produced right after delta_j: with data parallelism
// its all-reduce then overlaps the backprop of the earlier layers.
int backward_pass(b2nn_ctx* c, StepPlan& sp, int64_t x_start, const b2nn_train_desc* d) {
    const int nw = int(sp.fwd.size());
    const bool custom = d->act_kind == B2NN_ACT_CUSTOM;
    c->grad_order.clear();
    auto is_skinny = [&](int j) { return sp.skinny_bwd && j == nw - 1 && j > 0; };
    auto is_tail = [&](int j) { return sp.use_tail && j == nw - 1; };

    // G_j exists: hand it to the exchange (peer memory: slot copy / signal; NCCL: event for the communication stream)
    auto post_gradient = [&](int j) -> int {
        c->grad_order.push_back(j);
        if (sp.fused[j]) {
            if (sp.fused[j] == 2) {
                const LayerInfo& li = c->L[j];
                const int64_t cnt = li.ldw * li.out;
                B2_CUDA(p2p_small_scatter_launch(c->peers, c->G + li.w_off, c->p2p_small_off[j] + int64_t(c->rank) * cnt, cnt, c->stream));
                ++c->launches;
            }
            trace_mark(c, "wgrad_done", j, c->stream);
            B2_CUDA(p2p_signal_launch(c->peers, (0 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD + c->rank, c->p2p_step, c->stream));
            B2_CUDA(cudaEventRecord(c->p2p_wg_ev[j], c->stream));
        } else if (c->comm && !c->grad_ready.empty()) B2_CUDA(cudaEventRecord(c->grad_ready[j], c->stream));
        return B2NN_OK;
    };
    // delta_{j-1} from delta_j (cublasNN.cpp:673-676); the skinny kernel produces G_j in the same pass over a_{j-1}
    auto delta_step = [&](int j) -> int {
        if (j <= 0 || is_tail(j)) return B2NN_OK;              // the tail kernel already produced delta_{j-1}
        if (is_skinny(j)) {
            const LayerInfo& li = c->L[j];
            {
                ProfScope scope(c, 1, false, 4.0 * double(sp.rows) * li.in * li.out, 8.0 * double(sp.rows) * li.in);
                SkinnyArgs a{};
                a.a = c->ws.a[j - 1]; a.delta = c->ws.delta[j]; a.W = c->W + li.w_off; a.delta_prev = c->ws.delta[j - 1];
                a.G = c->G + li.w_off; a.ld = c->ws.ld; a.ldw = li.ldw; a.rows = sp.rows; a.H = li.in; a.K = li.out; a.act_kind = d->act_kind;
                if (!c->g_zero) B2_CUDA(cudaMemsetAsync(a.G, 0, sizeof(float) * size_t(li.ldw) * li.out, c->stream));
                const bool have_comp = !c->ws.delta_lo.empty() && c->ws.delta_lo[j - 1];
                if (have_comp && c->comp_bf16) a.delta_prev_bf16 = c->ws.delta_lo[j - 1];      // bf16 copy written in the same pass
                B2_CUDA(skinny_bwd_launch(a, c->stream));
            }
            ++c->launches;
            if (!c->ws.delta_lo.empty() && c->ws.delta_lo[j - 1] && !c->comp_bf16)      // fp32-grade mode: the next contraction wants the low part too
                B2_CUDA(make_companion(c->ws.delta[j - 1], c->ws.delta_lo[j - 1], int64_t(li.in) * c->ws.ld, false, c->stream));
            return post_gradient(j);
        }
        int rc = run_gemm(c, sp.bwd[j], 0, 1);
        if (rc) return rc;
        if (custom) {
            // user activation: sigGrad = f'(z_{j-1}) by the user's function, then delta = product .* sigGrad (cublasNN.cpp:673, 676)
            const int64_t cnt = int64_t(c->layers[j]) * c->ws.ld;
            if (cnt > INT32_MAX) { set_error("custom activation: element count exceeds int"); return B2NN_ERR_ARG; }
            float* tmp = c->ws.delta[j - 1];
            d->act_grad_fn(c->ws.z[j - 1], tmp, int(cnt));
            B2_CUDA(multiply_launch(c->ws.scratch, tmp, tmp, cnt, c->stream));
            ++c->launches;
        }
        return B2NN_OK;
    };
    // G_j = delta_j^T [a_{j-1}, 1]  (cublasNN.cpp:680-684)
    auto grad_step = [&](int j) -> int {
        if (is_skinny(j)) return B2NN_OK;                        // produced by delta_step
        if (!(is_tail(j) && sp.tail_inline_g)) { int rc = run_gemm(c, sp.wgrad[j], x_start, 2); if (rc) return rc; }
        return post_gradient(j);
    };

    if (c->p2p) {
        // Peer-memory exchange.  (1) An owner overwrites W_j on this rank as soon as EVERY rank has signalled its G_j, so
        // the whole delta chain (which reads W_j) runs before any signal.  (2) The gradients then come input layer FIRST:
        // the exchange of layer j hides behind the gradient GEMM of layer j+1, and the exchange of the LAST layer behind
        // the next step's forward pass through the earlier layers (each forward contraction waits for its own layer's
        // weights only, see forward_pass) — nothing is left exposed at the step boundary.
        for (int j = nw - 1; j >= 1; --j) { int rc = delta_step(j); if (rc) return rc; }
        for (int j = 0; j < nw; ++j) { int rc = grad_step(j); if (rc) return rc; }
        return B2NN_OK;
    }
    // Single GPU / NCCL: G_j right after delta_j (the reference computes all deltas, then all gradients, :670-684), so its
    // all-reduce overlaps the backprop of the earlier layers.
    // ... and (single GPU, large nets) the momentum update of layer j leaves for the exchange stream the moment G_j exists
    // and W_j has been read for the last time: a <= 32-register streaming kernel that becomes resident next to the
    // persistent GEMM CTAs and uses the HBM bandwidth the tensor-bound contractions leave idle.  Only the input layer's
    // (small) update is still exposed at the end of the step.
    auto overlapped_update = [&](int j) -> int {
        if (!c->ov_active) return B2NN_OK;
        const LayerInfo& li = c->L[j];
        const int64_t cnt = li.ldw * li.out;
        B2_CUDA(cudaEventRecord(c->ov_ev[j], c->stream));
        B2_CUDA(cudaStreamWaitEvent(c->upd_stream, c->ov_ev[j], 0));
        if (c->ov_prec == B2NN_PREC_BF16 && c->W_lo)
            B2_CUDA(momentum_update_bf16_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, comp_at(c->W_lo, li.w_off, true), cnt,
                                                c->ov_mu, c->ov_alpha, c->upd_stream));
        else if (c->ov_prec == B2NN_PREC_TF32X3 && c->W_lo)
            B2_CUDA(momentum_update_lo_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, c->W_lo + li.w_off, cnt, c->ov_mu,
                                              c->ov_alpha, c->upd_stream));
        else
            B2_CUDA(momentum_update_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, cnt, c->ov_mu, c->ov_alpha, c->upd_stream));
        ++c->launches;
        return B2NN_OK;
    };
    for (int j = nw - 1; j >= 0; --j) {
        int rc = grad_step(j);
        if (!rc) rc = delta_step(j);
        if (!rc) rc = overlapped_update(j);
        if (rc) return rc;
    }
    return B2NN_OK;
}

int output_and_delta(b2nn_ctx* c, const b2nn_train_desc* d, const float* Y, int64_t ldy, int64_t rows, double* cost_slot) {
    const int nl = int(c->layers.size());
    Workspace& w = c->ws;
    const int K = c->layers[nl - 1];
    const int out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
    const int cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
    if (out_kind == B2NN_OUT_CUSTOM || (cost_kind == B2NN_COST_CUSTOM && cost_slot)) {
        // callers route custom function pointers through output_custom(); reaching this is an engine bug, not a user error
        set_error("internal: custom output activation / cost reached a fused output kernel");
        return B2NN_ERR_ARG;
    }
    OutputArgs a{};
    a.z = w.z_last; a.ldz = w.ld;
    a.y = Y; a.ldy = ldy;
    a.h = nullptr; a.ldh = 0;
    a.delta = w.delta[nl - 2]; a.ldd = w.ld;
    a.cost_sum = cost_slot; a.pred = nullptr; a.err_sum = nullptr; a.y_labels = nullptr;
    a.rows = rows; a.K = K; a.out_kind = out_kind; a.cost_kind = cost_kind;
    {
        ProfScope scope(c, 3, false, 0.0, 12.0 * double(rows) * K);
        B2_CUDA(output_layer_launch(a, c->stream));
    }
    ++c->launches;
    return B2NN_OK;
}

bool custom_outcost(const b2nn_train_desc* d) {
    return d->classify && (d->out_kind == B2NN_OUT_CUSTOM || d->cost_kind == B2NN_COST_CUSTOM);
}

// User-supplied output activation and / or cost (function pointers of cublasNN.h:49-50): z_last is materialised, the
// user's host functions are called on device buffers exactly like the reference calls them (cublasNN.cpp:606, 613, 731-733;
// legacy default stream, ordered with the blocking engine stream), the rest stays in the fused output kernel.
// The matrices handed to the user functions have `ld` rows (>= rows, padding rows hold garbage that is never read back).
int output_custom(b2nn_ctx* c, const b2nn_train_desc* d, const float* Y, int64_t ldy, int64_t rows, bool train,
                  double* cost_slot, double* err_slot, const float* labels, float* pred, float* probs, int64_t ldp) {
    const int nl = int(c->layers.size());
    Workspace& w = c->ws;
    const int K = c->layers[nl - 1];
    const int64_t cnt = int64_t(K) * w.ld;
    if (cnt > INT32_MAX) { set_error("custom output / cost: element count exceeds int"); return B2NN_ERR_ARG; }
    if (d->out_kind == B2NN_OUT_CUSTOM && !d->out_fn) { set_error("custom output activation needs out_fn"); return B2NN_ERR_ARG; }
    if (d->cost_kind == B2NN_COST_CUSTOM && cost_slot && !d->cost_fn) { set_error("custom cost needs cost_fn"); return B2NN_ERR_ARG; }
    if (!w.hbuf) {
        B2_CUDA(cudaMalloc(&w.hbuf, sizeof(float) * cnt));
        B2_CUDA(cudaMalloc(&w.ybuf, sizeof(float) * cnt));
        B2_CUDA(cudaMalloc(&w.jbuf, sizeof(float) * cnt));
        B2_CUDA(cudaMemsetAsync(w.ybuf, 0, sizeof(float) * cnt, c->stream));
    }
    OutputArgs a{};
    if (d->out_kind == B2NN_OUT_CUSTOM) {
        d->out_fn(w.z_last, w.hbuf, int(w.ld), K);
    } else {
        a.z = w.z_last; a.ldz = w.ld; a.h = w.hbuf; a.ldh = w.ld; a.rows = rows; a.K = K;
        a.out_kind = d->out_kind; a.cost_kind = B2NN_COST_CROSSENTROPY;
        B2_CUDA(output_layer_launch(a, c->stream));
        ++c->launches;
    }
    const bool builtin_cost = d->cost_kind != B2NN_COST_CUSTOM;
    a = OutputArgs{};
    a.z = w.hbuf; a.ldz = w.ld;                     // identity "activation" over the user's h
    a.y = Y; a.ldy = ldy;
    a.h = probs; a.ldh = ldp;
    a.delta = train ? w.delta[nl - 2] : nullptr; a.ldd = w.ld;
    a.cost_sum = (builtin_cost && Y) ? cost_slot : nullptr;
    a.pred = pred; a.err_sum = err_slot; a.y_labels = labels;
    a.rows = rows; a.K = K; a.out_kind = B2NN_OUT_LINEAR; a.cost_kind = builtin_cost ? d->cost_kind : B2NN_COST_CROSSENTROPY;
    B2_CUDA(output_layer_launch(a, c->stream));
    ++c->launches;
    if (!builtin_cost && cost_slot && Y) {
        B2_CUDA(cudaMemcpy2DAsync(w.ybuf, sizeof(float) * w.ld, Y, sizeof(float) * ldy, sizeof(float) * rows, size_t(K),
                                  cudaMemcpyDeviceToDevice, c->stream));
        d->cost_fn(w.hbuf, w.ybuf, w.jbuf, int(cnt));
        B2_CUDA(abs_sum_launch(w.jbuf, w.ld, rows, K, cost_slot, c->stream));      // cublasSasum over the valid rows (:614, :734)
        ++c->launches;
    }
    return B2NN_OK;
}

// The fused last layer (tail.cu).  Training: z, h, delta_last, cost, delta_prev [, G_last].  Evaluation: z, h, cost,
// argmax, error count.
int run_tail(b2nn_ctx* c, const StepPlan& sp, const b2nn_train_desc* d, const float* Y, int64_t ldy, int64_t rows, bool train,
             double* cost_slot, double* err_slot, const float* labels, float* pred, float* probs, int64_t ldp) {
    const int nw = int(c->L.size());
    const LayerInfo& li = c->L[nw - 1];
    Workspace& w = c->ws;
    const int out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
    const int cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
    if (out_kind == B2NN_OUT_CUSTOM || (cost_kind == B2NN_COST_CUSTOM && cost_slot)) {
        // callers route custom function pointers through output_custom(); reaching this is an engine bug, not a user error
        set_error("internal: custom output activation / cost reached a fused output kernel");
        return B2NN_ERR_ARG;
    }
    TailArgs a{};
    a.a = w.a[nw - 2]; a.lda = w.ld;
    a.W = c->W + li.w_off; a.ldw = li.ldw;
    a.Y = Y; a.ldy = ldy;
    a.delta_last = train ? w.delta[nw - 1] : nullptr; a.ldd = w.ld;
    a.delta_prev = train ? w.delta[nw - 2] : nullptr; a.ldp = w.ld;
    a.G_inline = (train && sp.tail_inline_g) ? c->G + li.w_off : nullptr;
    a.h_out = probs; a.ldh = ldp;
    a.pred = pred; a.y_labels = labels;
    a.cost_sum = cost_slot; a.err_sum = err_slot;
    a.rows = rows; a.H = li.in; a.K = li.out;
    a.out_kind = out_kind; a.cost_kind = cost_kind; a.act_kind = d->act_kind;
    if (a.G_inline && !c->g_zero) B2_CUDA(cudaMemsetAsync(a.G_inline, 0, sizeof(float) * size_t(li.ldw) * li.out, c->stream));
    {
        ProfScope scope(c, 3, false, 2.0 * double(rows) * (li.in + 1) * li.out * (train ? 3 : 1),
                        4.0 * double(rows) * (li.in + 1) * (train ? 2 : 1));
        B2_CUDA(tail_launch(a, train, c->stream));
    }
    ++c->launches;
    return B2NN_OK;
}

bool p2p_layer_was_fused(b2nn_ctx* c, int j) {
    for (auto& kv : c->plans) if (int(kv.second.fused.size()) > j && kv.second.fused[j]) return true;
    return false;
}

// One mini-batch step on device-resident data: forward, output, backward, momentum update.
int train_step(b2nn_ctx* c, const b2nn_train_desc* d, int prec, const float* X, int64_t ldx, int64_t x_total, bool x_aligned,
               const float* Y, int64_t ldy, int64_t x_start, int64_t rows, int64_t global_rows, double* cost_slot) {
    const bool custom_oc = custom_outcost(d);
    StepPlan& sp = get_plan(c, rows, X, ldx, x_total, x_aligned, prec, d->act_kind, custom_oc);
    // "gradient arena is all-zero" survives this step only if it ends in the zeroing update (any other exit clears it)
    struct GZeroGuard { b2nn_ctx* c; bool keep; ~GZeroGuard() { if (!keep) c->g_zero = false; } } gz{c, false};
    if (c->comm) c->prof_chain = false;      // waits on the exchange streams sit between steps: not part of any launch
    bool any_fused = false;
    for (char f : sp.fused) any_fused = any_fused || f;
    ++c->trace_step;
    trace_mark(c, "step_begin", -1, c->stream);
    if (any_fused) {
        ++c->p2p_step;
        if (c->p2p_step > 1) {     // every owner's rows of the previous step have landed in this rank's weights (and every
            uint64_t mask = 0;     // rank has consumed the small-layer slots this step will overwrite)
            // Layers whose forward contraction is a tcgen05 launch wait for their own weights INSIDE that kernel (its TMA
            // producer polls the owners' flags): the end-of-step exchange of the late layers then overlaps the forward pass
            // through the early ones.  Everything else (slot layers, FFMA / tail consumers, the fp32-grade mode whose low
            // parts are re-derived below) is gated here.
            for (size_t j = 0; j < sp.fused.size(); ++j)
                if (sp.fused[j] && !(sp.fwd_waits[j])) mask |= uint64_t(1) << j;
            B2_CUDA(p2p_gate_layers_launch(c->p2p_flags + 1 * P2P_MAX_LAYERS * P2P_MAX_WORLD, mask, c->nranks, c->p2p_step - 1, c->stream));
        }
        // fp32-grade mode: the owners' kernels store new weight rows only; their low parts are re-derived here, once the
        // gate above has seen every owner's rows of the previous step land (one streaming pass, 8 B per weight)
        if (is_comp_prec(prec) && c->W_lo) { int rl = ensure_w_lo(c); if (rl) return rl; }
    }
    trace_mark(c, "gated", -1, c->stream);
    int rc = B2NN_OK;
    const bool micro = sp.micro && !custom_oc && !c->comm;
    if (micro) {
        MicroArgs a = sp.micro_args;
        a.X = X + x_start; a.ldx = ldx; a.Y = Y + x_start; a.ldy = ldy; a.rows = rows; a.cost_sum = cost_slot;
        a.out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
        a.cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
        if (!c->g_zero) B2_CUDA(cudaMemsetAsync(c->G, 0, sizeof(float) * size_t(c->total_int), c->stream));
        ProfScope scope(c, 0, false, 6.0 * double(rows) * double(c->total_int), 4.0 * double(rows) * (c->layers[0] + 1));
        B2_CUDA(micro_launch(a, c->stream));
        ++c->launches;
    } else if (sp.mlp3 && !custom_oc) {
        // config C1's shape: forward, output, backward and both gradients in one kernel over 32-sample tiles, then one
        // reduce + momentum-update kernel (2 launches; the reference: 15, cublasNN.cpp:597-629)
        const int grid = mlp3_grid(rows);
        const int64_t need = int64_t(grid) * c->total_int;
        if (need > c->m3_partial_cap) {
            B2_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->m3_partial); c->m3_partial = nullptr; c->m3_partial_cap = 0;
            B2_CUDA(cudaMalloc(&c->m3_partial, sizeof(float) * size_t(need)));
            c->m3_partial_cap = need;
        }
        Mlp3Args a{};
        a.X = X + x_start; a.ldx = ldx; a.Y = Y + x_start; a.ldy = ldy;
        a.W0 = c->W + c->L[0].w_off; a.ldw0 = int(c->L[0].ldw); a.W1 = c->W + c->L[1].w_off; a.ldw1 = int(c->L[1].ldw);
        a.partial = c->m3_partial; a.total = int(c->total_int); a.w1_off = int(c->L[1].w_off);
        a.cost_sum = cost_slot; a.rows = rows; a.n1 = c->L[0].in + 1; a.H = c->L[0].out; a.K = c->L[1].out;
        a.act_kind = d->act_kind; a.out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
        a.cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
        a.x3 = prec == B2NN_PREC_TF32X3 ? 1 : 0; a.train = 1;
        {
            ProfScope scope(c, 0, false, 6.0 * double(rows) * double(c->total_ref), 4.0 * double(rows) * (c->layers[0] + 1));
            B2_CUDA(mlp3_launch(a, c->stream));
        }
        {
            ProfScope scope(c, 4, false, 0.0, (4.0 * grid + 16.0) * double(c->total_int));
            B2_CUDA(mlp3_reduce_update_launch(c->m3_partial, grid, c->total_int, c->W, c->V, d->momentum,
                                              -d->rate / float(global_rows), c->stream));
        }
        c->launches += 2;
        c->w_lo_valid = false;
        return B2NN_OK;
    } else {
    // Output layer fused into the epilogue of the last forward contraction (north_star: "the softmax+NLL gradient fused into
    // the epilogues") when that contraction is a tcgen05 launch with a single column tile of <= 16 units (configs C3, C4):
    // with split-K the warp that adds the last partial sums of a 32-sample slice finishes the slice.
    TcOutput fused_out{};
    bool fuse_out = false;
    {
        const GemmOp& lastf = sp.fwd.back();
        const int Kout = c->layers.back();
        fuse_out = !custom_oc && !sp.use_tail && lastf.tc && !lastf.plan.swapped && lastf.plan.tiles_n == 1 && Kout <= 16 &&
                   lastf.g.epi == EPI_STORE && !std::getenv("B2NN_NO_FUSED_OUTPUT");
        if (fuse_out) {
            const int need = lastf.plan.tiles_m * lastf.plan.ncta * 4;
            if (need > c->out_counters_cap) {
                B2_CUDA(cudaStreamSynchronize(c->stream));
                cudaFree(c->out_counters); c->out_counters = nullptr; c->out_counters_cap = 0;
                B2_CUDA(cudaMalloc(&c->out_counters, sizeof(int) * size_t(need)));
                B2_CUDA(cudaMemset(c->out_counters, 0, sizeof(int) * size_t(need)));
                c->out_counters_cap = need;
            }
            fused_out.y = Y + x_start; fused_out.ldy = ldy;
            fused_out.delta = c->ws.delta[c->L.size() - 1]; fused_out.ldd = c->ws.ld;
            fused_out.cost_sum = cost_slot; fused_out.counters = c->out_counters;
            fused_out.out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
            fused_out.cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
        }
    }
    rc = forward_pass(c, sp, x_start, d, (any_fused && c->p2p_step > 1) ? c->p2p_step - 1 : 0, fuse_out ? &fused_out : nullptr);
    if (rc) return rc;
    trace_mark(c, "fwd_done", -1, c->stream);
    if (fuse_out) rc = B2NN_OK;
    else if (custom_oc) rc = output_custom(c, d, Y + x_start, ldy, rows, true, cost_slot, nullptr, nullptr, nullptr, nullptr, 0);
    else if (sp.use_tail) rc = run_tail(c, sp, d, Y + x_start, ldy, rows, true, cost_slot, nullptr, nullptr, nullptr, nullptr, 0);
    else rc = output_and_delta(c, d, Y + x_start, ldy, rows, cost_slot);
    if (rc) return rc;
    if (is_comp_prec(prec) && !c->ws.delta_lo.empty()) {
        const int nwl = int(c->L.size());
        B2_CUDA(make_companion(c->ws.delta[nwl - 1], c->ws.delta_lo[nwl - 1], int64_t(c->L[nwl - 1].out) * c->ws.ld, c->comp_bf16, c->stream));
        if (sp.use_tail)
            B2_CUDA(make_companion(c->ws.delta[nwl - 2], c->ws.delta_lo[nwl - 2], int64_t(c->L[nwl - 1].in) * c->ws.ld, c->comp_bf16, c->stream));
    }
    {
        // overlapped per-layer updates (see backward_pass): single GPU, built-in functions, nets too large for the
        // graph-replayed small-net path, not while per-launch profiling or capturing
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        (void)cudaStreamIsCapturing(c->stream, &cap);
        c->ov_active = !c->comm && d->act_kind != B2NN_ACT_CUSTOM && c->total_int > (int64_t(1) << 20) && !c->profile &&
                       cap == cudaStreamCaptureStatusNone && !std::getenv("B2NN_NO_OVERLAP_UPDATE");
        if (c->ov_active) {
            if (!c->upd_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->upd_stream, cudaStreamNonBlocking));
            if (!c->ov_done) B2_CUDA(cudaEventCreateWithFlags(&c->ov_done, cudaEventDisableTiming));
            while (c->ov_ev.size() < c->L.size()) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ov_ev.push_back(e); }
            c->ov_mu = d->momentum; c->ov_alpha = -d->rate / float(global_rows); c->ov_prec = prec;
        }
    }
    rc = backward_pass(c, sp, x_start, d);
    if (rc) { c->ov_active = false; return rc; }
    trace_mark(c, "bwd_done", -1, c->stream);
    }
    const float alpha = -d->rate / float(global_rows);      // cublasNN.cpp:544, 601
    const int nw = int(c->L.size());
    if (c->comm) {
        // gradient exchange per layer, issued in the order the gradients were produced
        for (int j : c->grad_order) {
            const LayerInfo& li = c->L[j];
            if (sp.fused[j] == 2) {
                // small layer: once every rank's copy sits in this rank's slots, sum them and update the local replica
                const int64_t cnt = li.ldw * li.out;
                B2_CUDA(cudaStreamWaitEvent(c->upd_stream, c->p2p_wg_ev[j], 0));
                B2_CUDA(p2p_gate_launch(c->p2p_flags + (0 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD, c->nranks, c->p2p_step, c->upd_stream));
                trace_mark(c, "U_gate", j, c->upd_stream);
                B2_CUDA(p2p_small_reduce_update_launch(c->peers.S[c->rank] + c->p2p_small_off[j], cnt, c->nranks, c->W + li.w_off,
                                                       c->V + li.w_off, cnt, d->momentum, alpha, c->upd_stream));
                trace_mark(c, "U_reduced", j, c->upd_stream);
                B2_CUDA(p2p_signal_launch(c->peers, (1 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD + c->rank, c->p2p_step, c->upd_stream));
                c->launches += 3;
                continue;
            }
            if (sp.fused[j]) {
                // owner side of the fused exchange, on its own stream: wait for every rank's partial rows, then reduce +
                // update + all-gather this rank's R rows and tell everyone
                const int64_t R = li.out / c->nranks, slice = R * li.ldw, off = li.w_off + int64_t(c->rank) * slice;
                B2_CUDA(cudaStreamWaitEvent(c->upd_stream, c->p2p_wg_ev[j], 0));
                B2_CUDA(p2p_gate_launch(c->p2p_flags + (0 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD, c->nranks, c->p2p_step, c->upd_stream));
                trace_mark(c, "U_gate", j, c->upd_stream);
                B2_CUDA(p2p_reduce_update_launch(c->peers, c->G + li.w_off, slice, c->V + off, off, slice, d->momentum, alpha, c->upd_stream));
                trace_mark(c, "U_reduced", j, c->upd_stream);
                B2_CUDA(p2p_signal_launch(c->peers, (1 * P2P_MAX_LAYERS + j) * P2P_MAX_WORLD + c->rank, c->p2p_step, c->upd_stream));
                c->launches += 3;
                continue;
            }
            B2_CUDA(cudaStreamWaitEvent(c->comm_stream, c->grad_ready[j], 0));
            float* gptr = c->G + li.w_off;
            int r = g_nccl.AllReduce(gptr, gptr, size_t(li.ldw) * li.out, NCCL_FLOAT32, NCCL_SUM, c->comm, c->comm_stream);
            if (r != 0) { set_error(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?")); return B2NN_ERR_CUDA; }
            B2_CUDA(cudaEventRecord(c->grad_reduced[j], c->comm_stream));
        }
        if (any_fused) {
            // No join of the two streams here: the next step orders itself against the owners' kernels (this rank's included)
            // through the w_done flags — per layer, inside the forward kernels — so the exchange of the last layers keeps
            // running under the next forward pass.  b2nn_train / b2nn_train_host end with a gate on every layer.
            if (std::getenv("B2NN_P2P_SYNC_STEP")) {
                B2_CUDA(cudaEventRecord(c->p2p_upd_done, c->upd_stream));
                B2_CUDA(cudaStreamWaitEvent(c->stream, c->p2p_upd_done, 0));
            }
            c->w_lo_valid = false;      // weights changed behind the low-part arena's back (refreshed after the next gate)
        }
        for (int j = nw - 1; j >= 0; --j) {
            const LayerInfo& li = c->L[j];
            if (sp.fused[j]) continue;
            B2_CUDA(cudaStreamWaitEvent(c->stream, c->grad_reduced[j], 0));
            {
                ProfScope scope(c, 4, false, 0.0, 20.0 * double(li.ldw) * li.out);
                if (prec == B2NN_PREC_BF16 && c->W_lo)
                    B2_CUDA(momentum_update_bf16_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, comp_at(c->W_lo, li.w_off, true),
                                                        li.ldw * li.out, d->momentum, alpha, c->stream));
                else if (prec == B2NN_PREC_TF32X3 && c->W_lo)
                    B2_CUDA(momentum_update_lo_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, c->W_lo + li.w_off,
                                                      li.ldw * li.out, d->momentum, alpha, c->stream));
                else
                    B2_CUDA(momentum_update_launch(c->W + li.w_off, c->V + li.w_off, c->G + li.w_off, li.ldw * li.out,
                                                   d->momentum, alpha, c->stream));
            }
            ++c->launches;
        }
    } else if (c->ov_active) {
        // every layer's update is already running (or done) on the exchange stream: the next step starts behind the last one
        c->ov_active = false;
        B2_CUDA(cudaEventRecord(c->ov_done, c->upd_stream));
        B2_CUDA(cudaStreamWaitEvent(c->stream, c->ov_done, 0));
    } else {
        // all layers in one pass: W, V, G share one arena layout (cublasNN.cpp:558-566 is 2 launches per layer)
        {
            ProfScope scope(c, 4, false, 0.0, 20.0 * double(c->total_int));
            // small nets: the update also zeroes the gradient arena, so the atomically accumulating producers of the next
            // step (split-K, in-kernel gradient sums) need no memset node in front of them — two to three fewer links in
            // a chain of few-microsecond launches.  Large nets keep the memsets (the extra 4 B per weight would cost more).
            const bool zeroing = !(is_comp_prec(prec) && c->W_lo) && c->total_int <= (int64_t(1) << 20) &&
                                 (c->total_int & 3) == 0 && !std::getenv("B2NN_NO_ZERO_UPDATE");
            if (prec == B2NN_PREC_BF16 && c->W_lo)
                B2_CUDA(momentum_update_bf16_launch(c->W, c->V, c->G, c->W_lo, c->total_int, d->momentum, alpha, c->stream));
            else if (prec == B2NN_PREC_TF32X3 && c->W_lo)
                B2_CUDA(momentum_update_lo_launch(c->W, c->V, c->G, c->W_lo, c->total_int, d->momentum, alpha, c->stream));
            else if (zeroing) {
                B2_CUDA(momentum_update_zero_launch(c->W, c->V, c->G, c->total_int, d->momentum, alpha, c->stream));
                c->g_zero = true;
                gz.keep = true;
            } else
                B2_CUDA(momentum_update_launch(c->W, c->V, c->G, c->total_int, d->momentum, alpha, c->stream));
        }
        ++c->launches;
    }
    return B2NN_OK;
}

// calcFinalCost / validate (cublasNN.cpp:687-881) over device-resident rows, chunked through the workspace.
int evaluate_device(b2nn_ctx* c, const b2nn_train_desc* d, int prec, const float* X, int64_t ldx, const float* Y, int64_t ldy,
                    const float* labels, int64_t rows, float* pred_out, float* probs_out, int64_t ldp, double* cost,
                    double* errors, int64_t x_extent = 0) {
    const int nl = int(c->layers.size());
    const int K = c->layers[nl - 1];
    int rc = ensure_acc(c, 2);
    if (rc) return rc;
    if (cost || errors) B2_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double) * 2, c->stream));
    // Chunk size: at least 2^16 rows, more while the per-chunk workspace (activations + deltas of every layer) stays
    // under 2 GiB — small nets then evaluate 2^20 rows in one pass (full occupancy for the streaming kernels) instead of
    // sixteen short launches.
    int64_t row_bytes = 0;
    for (int j = 1; j < nl; ++j) row_bytes += int64_t(sizeof(float)) * (2 * int64_t(c->layers[j]) + 1);
    const int64_t budget_rows = ((int64_t(2) << 30) / std::max<int64_t>(row_bytes, 1)) / 32 * 32;
    const int64_t chunk_cap = std::max<int64_t>(c->ws.cap, std::min<int64_t>(rows, std::max<int64_t>(1 << 16, budget_rows)));
    rc = ensure_workspace(c, chunk_cap, d->act_kind == B2NN_ACT_CUSTOM);
    if (rc) return rc;
    // One-hidden-layer nets in the fp32-grade mode: the forward half of mlp3.cu (operands split in registers, X read once)
    // instead of the tcgen05 path, which would first write a low-part copy of X (3x the HBM traffic of this HBM-bound pass).
    // Plain tf32 inference keeps the tcgen05 path (B2NN_MLP3_FWD=1 forces the one-kernel forward there too).
    const bool m3 = mlp3_ok(c, prec, d->act_kind) && !custom_outcost(d) && !std::getenv("B2NN_NO_TAIL") &&
                    (prec == B2NN_PREC_TF32X3 || std::getenv("B2NN_MLP3_FWD") != nullptr);
    if (is_comp_prec(prec) && d->act_kind != B2NN_ACT_CUSTOM && !m3) {
        rc = prepare_x3(c, X, ldx * int64_t(c->layers[0] + 1), c->ws.cap, false, prec == B2NN_PREC_BF16);
        if (rc) return rc;
    }
    // chunk starts stay on 32-sample boundaries so the forward contractions keep their one-box TMA loads
    const int64_t chunk = rows <= c->ws.cap ? c->ws.cap : std::max<int64_t>(32, c->ws.cap / 32 * 32);
    const int out_kind = d->classify ? d->out_kind : B2NN_OUT_LINEAR;
    const int cost_kind = d->classify ? d->cost_kind : B2NN_COST_SQUARED;
    for (int64_t s = 0; s < rows; s += chunk) {
        const int64_t r = std::min<int64_t>(chunk, rows - s);
        if (m3) {
            Mlp3Args a{};
            a.X = X + s; a.ldx = ldx; a.Y = Y ? Y + s : nullptr; a.ldy = ldy;
            a.W0 = c->W + c->L[0].w_off; a.ldw0 = int(c->L[0].ldw); a.W1 = c->W + c->L[1].w_off; a.ldw1 = int(c->L[1].ldw);
            a.h_out = probs_out ? probs_out + s : nullptr; a.ldh = ldp;
            a.pred = pred_out ? pred_out + s : nullptr; a.y_labels = labels ? labels + s : nullptr;
            a.cost_sum = (Y && cost) ? c->d_acc : nullptr;
            a.err_sum = (d->classify && errors && (Y || labels)) ? c->d_acc + 1 : nullptr;
            a.rows = r; a.n1 = c->L[0].in + 1; a.H = c->L[0].out; a.K = c->L[1].out;
            a.act_kind = d->act_kind; a.out_kind = out_kind; a.cost_kind = cost_kind;
            a.x3 = prec == B2NN_PREC_TF32X3 ? 1 : 0; a.train = 0;
            B2_CUDA(mlp3_launch(a, c->stream));
            ++c->launches;
            continue;
        }
        // x_extent: sample columns of X the tensor maps may span (staging buffers: the whole padded buffer, so plans stay
        // valid from call to call; rows past `r` only feed accumulator rows that are never stored)
        StepPlan& sp = get_plan(c, r, X, ldx, x_extent > 0 ? x_extent : rows, true, prec, d->act_kind, custom_outcost(d));
        rc = forward_pass(c, sp, s, d);
        if (rc) return rc;
        if (custom_outcost(d)) {
            rc = output_custom(c, d, Y ? Y + s : nullptr, ldy, r, false, (Y && cost) ? c->d_acc : nullptr,
                               (d->classify && errors && (Y || labels)) ? c->d_acc + 1 : nullptr, labels ? labels + s : nullptr,
                               pred_out ? pred_out + s : nullptr, probs_out ? probs_out + s : nullptr, ldp);
            if (rc) return rc;
            continue;
        }
        if (sp.use_tail) {
            rc = run_tail(c, sp, d, Y ? Y + s : nullptr, ldy, r, false, (Y && cost) ? c->d_acc : nullptr,
                          (d->classify && errors && (Y || labels)) ? c->d_acc + 1 : nullptr, labels ? labels + s : nullptr,
                          pred_out ? pred_out + s : nullptr, probs_out ? probs_out + s : nullptr, ldp);
            if (rc) return rc;
            continue;
        }
        OutputArgs a{};
        a.z = c->ws.z_last; a.ldz = c->ws.ld;
        a.y = Y ? Y + s : nullptr; a.ldy = ldy;
        a.h = probs_out ? probs_out + s : nullptr; a.ldh = ldp;
        a.delta = nullptr; a.ldd = 0;
        a.cost_sum = (Y && cost) ? c->d_acc : nullptr;
        a.pred = pred_out ? pred_out + s : nullptr;
        a.err_sum = (d->classify && errors && (Y || labels)) ? c->d_acc + 1 : nullptr;
        a.y_labels = labels ? labels + s : nullptr;
        a.rows = r; a.K = K; a.out_kind = out_kind; a.cost_kind = cost_kind;
        B2_CUDA(output_layer_launch(a, c->stream));
        ++c->launches;
    }
    if (cost || errors) {      // forward-only callers (predict) stay asynchronous on the engine stream
        B2_CUDA(cudaMemcpyAsync(c->h_acc, c->d_acc, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
        if (cost) *cost = d->classify ? c->h_acc[0] / double(rows) : c->h_acc[0] / (2.0 * double(rows));   // :735, :741
        if (errors) *errors = c->h_acc[1];
    }
    return B2NN_OK;
}

// Stage a reference-layout host matrix (rows x (n+1) col-major, ones column FIRST, ld = ld_host) into engine layout.
int stage_x_from_host(b2nn_ctx* c, const float* x_host, int64_t ld_host, int64_t row0, int64_t rows, int n, float* x_dev,
                      int64_t ld_dev, bool fill_ones, bool raw = false) {
    // feature f of the engine matrix is column f+1 of the reference matrix (column f of a RAW matrix, which has no ones
    // column yet): one strided 2-D copy, no temp buffer
    B2_CUDA(cudaMemcpy2DAsync(x_dev, sizeof(float) * ld_dev, x_host + (raw ? 0 : ld_host) + row0, sizeof(float) * ld_host,
                              sizeof(float) * rows, size_t(n), cudaMemcpyHostToDevice, c->stream));
    if (fill_ones) B2_CUDA(fill_launch(x_dev + int64_t(n) * ld_dev, 1.0f, ld_dev, c->stream));
    return B2NN_OK;
}

// `training`: the entry point back-propagates, so a custom hidden activation needs its derivative too; forward-only entry
// points (evaluate / predict, which the reference calls with activationHidden alone: cublasNN.h:51-69) need act_fn only.
int validate_desc(const b2nn_ctx* c, const b2nn_train_desc* d, bool training) {
    if (!c || !d) { set_error("null context or descriptor"); return B2NN_ERR_ARG; }
    if (c->layers.size() < 3) { set_error("set_layers must be called first with >= 3 layers (cublasNN.cpp:657 needs a hidden layer)"); return B2NN_ERR_ARG; }
    if (d->act_kind < 0 || d->act_kind > B2NN_ACT_CUSTOM) { set_error("bad act_kind"); return B2NN_ERR_ARG; }
    if (d->act_kind == B2NN_ACT_CUSTOM && !d->act_fn) { set_error("custom activation needs act_fn"); return B2NN_ERR_ARG; }
    if (d->act_kind == B2NN_ACT_CUSTOM && training && !d->act_grad_fn) { set_error("custom activation needs act_fn and act_grad_fn"); return B2NN_ERR_ARG; }
    if (d->classify && (d->out_kind < 0 || d->out_kind > B2NN_OUT_CUSTOM)) { set_error("bad out_kind"); return B2NN_ERR_ARG; }
    return B2NN_OK;
}

// Host-streaming staging buffers (b2nn_train_step_host, b2nn_train_host, b2nn_predict) are sized for one (n, K); drop
// them whenever the layer widths may have changed.
void release_staging(b2nn_ctx* c) {
    drop_lo(c, c->Xs);
    cudaFree(c->Xs); cudaFree(c->Ys); c->Xs = c->Ys = nullptr; c->caps = 0; c->lds = 0;
    for (int i = 0; i < 2; ++i) {
        drop_lo(c, c->Xst[i]);
        cudaFree(c->Xst[i]); cudaFree(c->Yst[i]); c->Xst[i] = c->Yst[i] = nullptr;
    }
    c->capst = 0; c->ldst = 0;
    for (int i = 0; i < 2; ++i) {
        drop_lo(c, c->pr_x[i]);
        cudaFree(c->pr_x[i]); cudaFree(c->pr_p[i]); cudaFree(c->pr_h[i]);
        c->pr_x[i] = c->pr_p[i] = c->pr_h[i] = nullptr;
        if (c->pr_hx[i]) cudaFreeHost(c->pr_hx[i]);
        if (c->pr_hout[i]) cudaFreeHost(c->pr_hout[i]);
        c->pr_hx[i] = c->pr_hout[i] = nullptr;
    }
    c->pr_ld = 0; c->pr_n = c->pr_k = 0;
}

int comm_barrier_impl(b2nn_ctx* c);

// Undo b2nn_comm_p2p_export / _open.  Collective when a communicator exists: the barrier separates "every rank has
// closed its imports" from "the exporter frees the memory" (freeing exported memory under a live import is undefined).
int p2p_teardown(b2nn_ctx* c) {
    if (!c->p2p && !c->p2p_flags && c->p2p_opened.empty()) return B2NN_OK;
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->upd_stream) cudaStreamSynchronize(c->upd_stream);
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    c->p2p_opened.clear();
    c->p2p = false;
    c->peers = P2pPeers{};
    if (c->comm) { int rc = comm_barrier_impl(c); if (rc) return rc; }
    cudaFree(c->p2p_flags); c->p2p_flags = nullptr;
    c->p2p_step = 0;
    c->p2p_small_off.clear(); c->p2p_small_cnt.clear(); c->p2p_small_floats = 0;
    c->plans.clear();
    return B2NN_OK;
}

int comm_barrier_impl(b2nn_ctx* c) {
    B2_CUDA(cudaSetDevice(c->device));
    if (c->comm) {
        int rc = ensure_acc(c, 3);
        if (rc) return rc;
        int r = g_nccl.AllReduce(c->d_acc, c->d_acc, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
        if (r != 0) { set_error("ncclAllReduce(barrier) failed"); return B2NN_ERR_CUDA; }
    }
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2NN_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char* b2nn_last_error(void) { return g_last_error.c_str(); }
const char* b2nn_version(void) { return "b2nn 0.1 (sm_100a)"; }

int b2nn_device_ok(void) {
    std::string why;
    return check_device(-1, why) == B2NN_OK ? 1 : 0;
}

int b2nn_create(b2nn_ctx** out, int device) {
    if (!out) { set_error("null out pointer"); return B2NN_ERR_ARG; }
    *out = nullptr;
    std::string why;
    int rc = check_device(device, why);
    if (rc) { set_error(why); return rc; }
    if (device < 0) cudaGetDevice(&device);
    B2_CUDA(cudaSetDevice(device));
    struct CtxGuard { b2nn_ctx* p; ~CtxGuard() { if (p) b2nn_destroy(p); } b2nn_ctx* operator->() const { return p; } } c{new b2nn_ctx};
    c->device = device;
    B2_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault));          // blocking: orders with user ops on stream 0
    B2_CUDA(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    if (curandCreateGenerator(&c->gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) {   // cublasNN.cpp:26
        set_error("curandCreateGenerator failed"); return B2NN_ERR_CUDA;
    }
    curandSetStream(c->gen, c->stream);
    *out = c.p;
    c.p = nullptr;
    return B2NN_OK;
}

void b2nn_destroy(b2nn_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
    cudaFree(c->p2p_flags);
    for (auto e : c->p2p_wg_ev) cudaEventDestroy(e);
    if (c->p2p_upd_done) cudaEventDestroy(c->p2p_upd_done);
    if (c->upd_stream) cudaStreamDestroy(c->upd_stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    c->ws.release();
    cudaFree(c->W); cudaFree(c->V); cudaFree(c->G); cudaFree(c->W_lo);
    for (auto& kv : c->lo_cache) cudaFree(kv.second.first);
    cudaFree(c->X); cudaFree(c->Y); cudaFree(c->Xs); cudaFree(c->Ys); cudaFree(c->Xsp); cudaFree(c->Ysp);
    cudaFree(c->d_acc);
    cudaFree(c->d_mean); cudaFree(c->d_std);
    cudaFree(c->m3_partial); cudaFree(c->out_counters);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->pr_x[i]); cudaFree(c->pr_p[i]); cudaFree(c->pr_h[i]);
        if (c->pr_hx[i]) cudaFreeHost(c->pr_hx[i]);
        if (c->pr_hout[i]) cudaFreeHost(c->pr_hout[i]);
        if (c->pr_copied[i]) cudaEventDestroy(c->pr_copied[i]);
        if (c->pr_computed[i]) cudaEventDestroy(c->pr_computed[i]);
        if (c->pr_out[i]) cudaEventDestroy(c->pr_out[i]);
    }
    if (c->pr_d2h_stream) cudaStreamDestroy(c->pr_d2h_stream);
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->Xst[i]); cudaFree(c->Yst[i]);
        if (c->st_copied[i]) cudaEventDestroy(c->st_copied[i]);
        if (c->st_consumed[i]) cudaEventDestroy(c->st_consumed[i]);
    }
    cudaFree(c->d_jsteps);
    if (c->h_jsteps) cudaFreeHost(c->h_jsteps);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->loss_stream) cudaStreamDestroy(c->loss_stream);
    if (c->st_loss) cudaEventDestroy(c->st_loss);
    if (c->h_acc) cudaFreeHost(c->h_acc);
    for (auto e : c->log_events) cudaEventDestroy(e);
    for (auto& r : c->prof) { if (r.own_e0) cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    for (auto e : c->ov_ev) cudaEventDestroy(e);
    if (c->ov_done) cudaEventDestroy(c->ov_done);
    for (auto e : c->grad_ready) cudaEventDestroy(e);
    for (auto e : c->grad_reduced) cudaEventDestroy(e);
    if (c->gen) curandDestroyGenerator(c->gen);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    delete c;
}

int b2nn_set_layers(b2nn_ctx* c, const int* layers, int n_layers, int init_kind, b2nn_init_fn custom_init, uint64_t seed,
                    int seed_is_clock) {
    if (!c || !layers) { set_error("null argument"); return B2NN_ERR_ARG; }
    if (n_layers < 3) { set_error("need at least one hidden layer (cublasNN.cpp:657)"); return B2NN_ERR_ARG; }
    for (int i = 0; i < n_layers; ++i) if (layers[i] < 1) { set_error("layer width must be >= 1"); return B2NN_ERR_ARG; }
    if (init_kind != B2NN_INIT_1 && init_kind != B2NN_INIT_2 && !(init_kind == B2NN_INIT_CUSTOM && custom_init)) {
        set_error("bad init_kind"); return B2NN_ERR_ARG;
    }
    B2_CUDA(cudaSetDevice(c->device));
    // Peer-memory exchange: peers hold IPC mappings of the arenas about to be freed.  Collective teardown (every rank
    // calls set_layers): close what this rank imported, meet at a barrier, only then free what it exported.
    if (int rc = p2p_teardown(c)) return rc;
    c->layers.assign(layers, layers + n_layers);
    c->L.assign(n_layers - 1, LayerInfo());
    c->total_ref = 0; c->total_int = 0;
    for (int j = 0; j < n_layers - 1; ++j) {
        LayerInfo& li = c->L[j];
        li.in = layers[j]; li.out = layers[j + 1];
        li.ldw = round_up(int64_t(li.in) + 1, 8);      // 8: rows of the bf16 copy of the weights start 16-byte aligned too (TMA)
        li.ref_off = c->total_ref; li.ref_size = (int64_t(li.in) + 1) * li.out;
        li.w_off = c->total_int;
        c->total_ref += li.ref_size;
        c->total_int += li.ldw * li.out;      // multiple of 4 floats: every layer block starts 16-byte aligned
    }
    release_staging(c);       // sized for the old input / output widths (their ones row sits at the old index n)
    cudaFree(c->W); cudaFree(c->V); cudaFree(c->G); cudaFree(c->W_lo);
    c->W = c->V = c->G = c->W_lo = nullptr; c->w_lo_valid = false;
    c->plans.clear(); c->ws.release();
    // + 64 floats: the folded 3-D TMA view of the last weight rows may read one 32-float atom past the arena's end
    if (cudaMalloc(&c->W, sizeof(float) * (c->total_int + 64)) != cudaSuccess ||
        cudaMalloc(&c->V, sizeof(float) * c->total_int) != cudaSuccess ||
        cudaMalloc(&c->G, sizeof(float) * c->total_int) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc Failed");     // reference text, cublasNN.cpp:309
        return B2NN_ERR_ALLOC;
    }
    B2_CUDA(cudaMemsetAsync(c->W, 0, sizeof(float) * (c->total_int + 64), c->stream));
    B2_CUDA(cudaMemsetAsync(c->V, 0, sizeof(float) * c->total_int, c->stream));
    B2_CUDA(cudaMemsetAsync(c->G, 0, sizeof(float) * c->total_int, c->stream));
    c->g_zero = false;        // a fresh arena: the first step zeroes what it accumulates into

    // One uniform draw over the whole reference-layout arena, then the per-layer rescale (cublasNN.cpp:314-326).
    DevBuf refbuf;
    B2_CUDA(cudaMalloc(&refbuf.p, sizeof(float) * c->total_ref));
    float* const ref = refbuf.as<float>();
    const unsigned long long s = seed_is_clock ? (unsigned long long)clock() : (unsigned long long)seed;   // :27
    curandSetPseudoRandomGeneratorSeed(c->gen, s);
    curandSetGeneratorOffset(c->gen, 0ULL);
    if (curandGenerateUniform(c->gen, ref, size_t(c->total_ref)) != CURAND_STATUS_SUCCESS) {
        set_error("curandGenerateUniform failed"); return B2NN_ERR_CUDA;
    }
    for (int j = 0; j < n_layers - 1; ++j) {
        const LayerInfo& li = c->L[j];
        if (init_kind == B2NN_INIT_CUSTOM) {
            if (li.ref_size > INT32_MAX) { set_error("custom init: layer too large for int count"); return B2NN_ERR_ARG; }
            custom_init(ref + li.ref_off, li.in + 1, li.out, int(li.ref_size));       // legacy stream, ordered by the blocking engine stream
        } else {
            B2_CUDA(init_scale_launch(ref + li.ref_off, li.in + 1, li.out, li.ref_size, init_kind, c->stream));
        }
    }
    if (init_kind == B2NN_INIT_CUSTOM) B2_CUDA(cudaDeviceSynchronize());
    for (int j = 0; j < n_layers - 1; ++j) {
        const LayerInfo& li = c->L[j];
        B2_CUDA(theta_ref_to_internal_launch(ref + li.ref_off, c->W + li.w_off, li.in, li.out, li.ldw, c->stream));
    }
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2NN_OK;
}

int64_t b2nn_num_weights(const b2nn_ctx* c) { return c ? c->total_ref : 0; }

// One arena (weights or velocity) between the engine layout and the reference arena layout on the host.
static int arena_to_host(b2nn_ctx* c, const float* arena, float* host) {
    DevBuf ref;
    B2_CUDA(cudaMalloc(&ref.p, sizeof(float) * c->total_ref));
    for (const LayerInfo& li : c->L)
        B2_CUDA(theta_internal_to_ref_launch(arena + li.w_off, ref.as<float>() + li.ref_off, li.in, li.out, li.ldw, c->stream));
    B2_CUDA(cudaMemcpyAsync(host, ref.p, sizeof(float) * c->total_ref, cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2NN_OK;
}
static int arena_from_host(b2nn_ctx* c, float* arena, const float* host) {
    DevBuf ref;
    B2_CUDA(cudaMalloc(&ref.p, sizeof(float) * c->total_ref));
    B2_CUDA(cudaMemcpyAsync(ref.p, host, sizeof(float) * c->total_ref, cudaMemcpyHostToDevice, c->stream));
    for (const LayerInfo& li : c->L)
        B2_CUDA(theta_ref_to_internal_launch(ref.as<float>() + li.ref_off, arena + li.w_off, li.in, li.out, li.ldw, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2NN_OK;
}

int b2nn_get_weights(b2nn_ctx* c, float* host, int64_t count) {
    if (!c || !host || count != c->total_ref || !c->W) { set_error("get_weights: bad arguments"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    return arena_to_host(c, c->W, host);
}

int b2nn_set_weights(b2nn_ctx* c, const float* host, int64_t count) {
    if (!c || !host || count != c->total_ref || !c->W) { set_error("set_weights: bad arguments"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    c->w_lo_valid = false;
    return arena_from_host(c, c->W, host);
}

// ---- checkpoint (not in the reference: its weights die with the process) ---------------------------------------------
// File: "B2NNCKP1" | int32 n_layers | int32 layers[] | int32 has_velocity | int32 norm_n | int64 n_weights |
//       float theta[] (reference arena layout) | float velocity[] | float mean[norm_n] | float stddev[norm_n] |
//       uint64 FNV-1a of everything before it.  Little-endian, no padding.
static uint64_t fnv1a(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int b2nn_save_checkpoint(b2nn_ctx* c, const char* path) {
    if (!c || !path || !c->W) { set_error("save_checkpoint: no network"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    std::vector<float> th(static_cast<size_t>(c->total_ref)), vel(static_cast<size_t>(c->total_ref)), mean(static_cast<size_t>(c->norm_n)),
        sd(static_cast<size_t>(c->norm_n));
    int rc = arena_to_host(c, c->W, th.data());
    if (!rc) rc = arena_to_host(c, c->V, vel.data());
    if (rc) return rc;
    if (c->norm_n > 0) {
        B2_CUDA(cudaMemcpy(mean.data(), c->d_mean, sizeof(float) * c->norm_n, cudaMemcpyDeviceToHost));
        B2_CUDA(cudaMemcpy(sd.data(), c->d_std, sizeof(float) * c->norm_n, cudaMemcpyDeviceToHost));
    }
    FILE* f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("save_checkpoint: cannot open ") + path); return B2NN_ERR_ARG; }
    uint64_t h = 1469598103934665603ull;
    bool ok = true;
    auto put = [&](const void* p, size_t n) { if (n && std::fwrite(p, 1, n, f) != n) ok = false; h = fnv1a(h, p, n); };
    const int32_t nl = int32_t(c->layers.size()), has_v = 1, nn = int32_t(c->norm_n);
    const int64_t nw = c->total_ref;
    put("B2NNCKP1", 8); put(&nl, 4);
    for (int v : c->layers) { const int32_t w = v; put(&w, 4); }
    put(&has_v, 4); put(&nn, 4); put(&nw, 8);
    put(th.data(), sizeof(float) * th.size()); put(vel.data(), sizeof(float) * vel.size());
    put(mean.data(), sizeof(float) * mean.size()); put(sd.data(), sizeof(float) * sd.size());
    const uint64_t sum = h;
    if (std::fwrite(&sum, 1, 8, f) != 8) ok = false;
    if (std::fclose(f) != 0) ok = false;
    if (!ok) { set_error(std::string("save_checkpoint: write failed: ") + path); return B2NN_ERR_ARG; }
    return B2NN_OK;
}

int b2nn_load_checkpoint(b2nn_ctx* c, const char* path) {
    if (!c || !path) { set_error("load_checkpoint: null argument"); return B2NN_ERR_ARG; }
    FILE* f = std::fopen(path, "rb");
    if (!f) { set_error(std::string("load_checkpoint: cannot open ") + path); return B2NN_ERR_ARG; }
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{f};
    uint64_t h = 1469598103934665603ull;
    bool ok = true;
    auto get = [&](void* p, size_t n) { if (n && std::fread(p, 1, n, f) != n) ok = false; else h = fnv1a(h, p, n); };
    char magic[8]; int32_t nl = 0, has_v = 0, nn = 0; int64_t nw = 0;
    get(magic, 8); get(&nl, 4);
    if (!ok || std::memcmp(magic, "B2NNCKP1", 8) != 0 || nl < 3 || nl > 4096) { set_error("load_checkpoint: not a b2nn checkpoint"); return B2NN_ERR_ARG; }
    std::vector<int> layers(size_t(nl), 0);
    for (int i = 0; i < nl; ++i) { int32_t w = 0; get(&w, 4); layers[size_t(i)] = w; }
    get(&has_v, 4); get(&nn, 4); get(&nw, 8);
    int64_t expect = 0;
    for (int i = 0; ok && i + 1 < nl; ++i) {
        if (layers[size_t(i)] < 1 || layers[size_t(i + 1)] < 1) { ok = false; break; }
        expect += (int64_t(layers[size_t(i)]) + 1) * layers[size_t(i + 1)];
        if (expect > (int64_t(1) << 40)) ok = false;
    }
    if (!ok || nw != expect || (has_v != 0 && has_v != 1) || nn < 0 || (nn != 0 && nn != layers[0])) {
        set_error("load_checkpoint: inconsistent header"); return B2NN_ERR_ARG;
    }
    // the header must agree with the file's length before anything is sized by it
    const long body0 = std::ftell(f);
    if (body0 < 0 || std::fseek(f, 0, SEEK_END) != 0) { set_error("load_checkpoint: cannot seek"); return B2NN_ERR_ARG; }
    const long flen = std::ftell(f);
    const int64_t need = int64_t(sizeof(float)) * (nw * (1 + has_v) + 2 * int64_t(nn)) + 8;
    if (flen < 0 || int64_t(flen) - int64_t(body0) != need || std::fseek(f, body0, SEEK_SET) != 0) {
        set_error("load_checkpoint: truncated file or checksum mismatch"); return B2NN_ERR_ARG;
    }
    std::vector<float> th, vel, mean, sd;
    try {
        th.resize(static_cast<size_t>(nw)); vel.resize(has_v ? static_cast<size_t>(nw) : 0);
        mean.resize(static_cast<size_t>(nn)); sd.resize(static_cast<size_t>(nn));
    } catch (const std::bad_alloc&) { set_error("load_checkpoint: out of host memory"); return B2NN_ERR_ALLOC; }
    get(th.data(), sizeof(float) * th.size()); get(vel.data(), sizeof(float) * vel.size());
    get(mean.data(), sizeof(float) * mean.size()); get(sd.data(), sizeof(float) * sd.size());
    const uint64_t sum = h;
    uint64_t stored = 0;
    if (!ok || std::fread(&stored, 1, 8, f) != 8 || stored != sum) { set_error("load_checkpoint: truncated file or checksum mismatch"); return B2NN_ERR_ARG; }
    const bool same = c->W && c->layers == layers;
    if (!same) {        // (re)build the arenas for the stored architecture; the random draw is overwritten below
        int rc = b2nn_set_layers(c, layers.data(), nl, B2NN_INIT_1, nullptr, 0, 0);
        if (rc) return rc;
    }
    B2_CUDA(cudaSetDevice(c->device));
    c->w_lo_valid = false;
    int rc = arena_from_host(c, c->W, th.data());
    if (!rc && has_v) rc = arena_from_host(c, c->V, vel.data());
    if (!rc) rc = b2nn_set_normalisation(c, mean.data(), sd.data(), nn);
    return rc;
}

static int set_train_data_impl(b2nn_ctx* c, const float* x, const float* y, int64_t m, int n_plus_1, int k, bool raw, bool normalise,
                               float* mean_out, float* std_out);

int b2nn_set_train_data(b2nn_ctx* c, const float* x, const float* y, int64_t m, int n_plus_1, int k) {
    return set_train_data_impl(c, x, y, m, n_plus_1, k, false, false, nullptr, nullptr);
}

int b2nn_set_train_data_raw(b2nn_ctx* c, const float* x_raw, const float* y, int64_t m, int n, int k, int normalise, float* mean_out,
                            float* std_out) {
    return set_train_data_impl(c, x_raw, y, m, n + 1, k, true, normalise != 0, mean_out, std_out);
}

int b2nn_set_normalisation(b2nn_ctx* c, const float* mean, const float* stddev, int n) {
    if (!c || n < 0 || (n > 0 && (!mean || !stddev))) { set_error("set_normalisation: bad arguments"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_mean); cudaFree(c->d_std); c->d_mean = c->d_std = nullptr; c->norm_n = 0;
    if (n == 0) return B2NN_OK;
    B2_CUDA(cudaMalloc(&c->d_mean, sizeof(float) * n));
    B2_CUDA(cudaMalloc(&c->d_std, sizeof(float) * n));
    B2_CUDA(cudaMemcpy(c->d_mean, mean, sizeof(float) * n, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(c->d_std, stddev, sizeof(float) * n, cudaMemcpyHostToDevice));
    c->norm_n = n;
    return B2NN_OK;
}

int b2nn_get_normalisation(b2nn_ctx* c, float* mean, float* stddev, int n) {
    if (!c || !mean || !stddev || n != c->norm_n || n == 0) { set_error("get_normalisation: no statistics of that width"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    B2_CUDA(cudaMemcpy(mean, c->d_mean, sizeof(float) * n, cudaMemcpyDeviceToHost));
    B2_CUDA(cudaMemcpy(stddev, c->d_std, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return B2NN_OK;
}

static int set_train_data_impl(b2nn_ctx* c, const float* x, const float* y, int64_t m, int n_plus_1, int k, bool raw, bool normalise,
                               float* mean_out, float* std_out) {
    if (!c || !x || !y || m < 1) { set_error("set_train_data: bad arguments"); return B2NN_ERR_ARG; }
    if (normalise && m < 2) { set_error("set_train_data_raw: normalisation needs at least 2 rows"); return B2NN_ERR_ARG; }
    if (c->layers.empty()) { set_error("set_layers must precede set_train_data (main.cpp:125)"); return B2NN_ERR_ARG; }
    if (n_plus_1 != c->layers[0] + 1 || k != c->layers.back()) { set_error("data shape does not match the layers"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    drop_lo(c, c->X); drop_lo(c, c->Xsp);
    cudaFree(c->X); cudaFree(c->Y); c->X = c->Y = nullptr;
    cudaFree(c->Xsp); cudaFree(c->Ysp); c->Xsp = c->Ysp = nullptr; c->sp_nb = 0;
    c->plans.clear();
    c->m = m; c->n = n_plus_1 - 1; c->K = k;
    c->ldx = round_up(m, 32); c->ldy = c->ldx;
    if (cudaMalloc(&c->X, sizeof(float) * c->ldx * n_plus_1) != cudaSuccess ||
        cudaMalloc(&c->Y, sizeof(float) * c->ldy * k) != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("cudaMalloc Failed");      // cublasNN.cpp:366
        return B2NN_ERR_ALLOC;
    }
    B2_CUDA(cudaMemsetAsync(c->X, 0, sizeof(float) * c->ldx * n_plus_1, c->stream));
    B2_CUDA(cudaMemsetAsync(c->Y, 0, sizeof(float) * c->ldy * k, c->stream));
    int rc = stage_x_from_host(c, x, m, 0, m, c->n, c->X, c->ldx, true, raw);
    if (rc) return rc;
    if (normalise) {
        // normaliseData on the device (cublasNN.cpp:214-268): statistics of the training set, kept for validate / predict
        if (c->norm_n != c->n) {
            cudaFree(c->d_mean); cudaFree(c->d_std); c->d_mean = c->d_std = nullptr; c->norm_n = 0;
            B2_CUDA(cudaMalloc(&c->d_mean, sizeof(float) * c->n));
            B2_CUDA(cudaMalloc(&c->d_std, sizeof(float) * c->n));
            c->norm_n = c->n;
        }
        B2_CUDA(feature_stats_launch(c->X, c->ldx, m, c->n, c->d_mean, c->d_std, c->stream));
        B2_CUDA(normalise_rows_launch(c->X, c->ldx, m, c->n, c->d_mean, c->d_std, c->stream));
        if (mean_out) B2_CUDA(cudaMemcpyAsync(mean_out, c->d_mean, sizeof(float) * c->n, cudaMemcpyDeviceToHost, c->stream));
        if (std_out) B2_CUDA(cudaMemcpyAsync(std_out, c->d_std, sizeof(float) * c->n, cudaMemcpyDeviceToHost, c->stream));
    }
    B2_CUDA(cudaMemcpy2DAsync(c->Y, sizeof(float) * c->ldy, y, sizeof(float) * m, sizeof(float) * m, size_t(k),
                              cudaMemcpyHostToDevice, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    return B2NN_OK;
}

int b2nn_reset_velocity(b2nn_ctx* c) {
    if (!c || !c->V) { set_error("reset_velocity: no network"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(cudaMemsetAsync(c->V, 0, sizeof(float) * c->total_int, c->stream));
    return B2NN_OK;
}

int b2nn_train(b2nn_ctx* c, const b2nn_train_desc* d, b2nn_log_cb log, void* log_user, b2nn_train_result* result) {
    auto t_call0 = std::chrono::steady_clock::now();                    // cublasNN.cpp:529, 586
    int rc = validate_desc(c, d, true);
    if (rc) return rc;
    if (!c->X || !c->Y) { set_error("set_train_data must precede train (copyDataGPU, main.cpp:87)"); return B2NN_ERR_ARG; }
    if (d->batch_num < 1 || int64_t(d->batch_num) > c->m) { set_error("batch_num must be in [1, m]"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int display = d->display > 0 ? d->display : 1;               // Q5
    const int nb = d->batch_num;
    const int64_t bs = c->m / nb;                                      // cublasNN.cpp:401
    const int64_t last_rows = c->m - int64_t(nb - 1) * bs;             // last batch absorbs the remainder (:429)
    const int prec = effective_precision(c, resolve_precision(d->precision), last_rows);

    B2_CUDA(cudaMemsetAsync(c->V, 0, sizeof(float) * c->total_int, c->stream));   // :533-534, :590-591
    rc = ensure_workspace(c, last_rows, d->act_kind == B2NN_ACT_CUSTOM);
    if (rc) return rc;
    const int n_disp = d->iters / display;
    rc = ensure_acc(c, 2 + std::max(n_disp, 1));
    if (rc) return rc;
    B2_CUDA(cudaMemsetAsync(c->d_acc, 0, sizeof(double) * c->acc_slots, c->stream));
    while (int(c->log_events.size()) < 32) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->log_events.push_back(e); }

    // data parallel: the update uses the GLOBAL batch size (sum over ranks of this batch's rows)
    std::vector<int64_t> global_rows(nb);
    for (int b = 0; b < nb; ++b) global_rows[b] = (b == nb - 1) ? last_rows : bs;
    if (c->comm) {
        DevBuf tmp;
        B2_CUDA(cudaMalloc(&tmp.p, sizeof(int64_t) * nb));
        B2_CUDA(cudaMemcpyAsync(tmp.p, global_rows.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, c->stream));
        int r = g_nccl.AllReduce(tmp.p, tmp.p, size_t(nb), NCCL_INT64, NCCL_SUM, c->comm, c->stream);
        if (r != 0) { set_error("ncclAllReduce(batch rows) failed"); return B2NN_ERR_CUDA; }
        B2_CUDA(cudaMemcpyAsync(global_rows.data(), tmp.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }

    if (c->comm) {
        while (c->grad_ready.size() < c->L.size()) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_ready.push_back(e); }
        while (c->grad_reduced.size() < c->L.size()) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_reduced.push_back(e); }
    }
    // splitData (cublasNN.cpp:498-524): the reference re-lays every batch as a contiguous block.  Here batches are already
    // contiguous column ranges of X; a copy is only made when the batch starts would not fall on 32-sample boundaries.
    const float* Xtrain = c->X; const float* Ytrain = c->Y;
    int64_t ld_train = c->ldx, x_total = c->m, b_stride = bs;
    bool x_aligned = (nb == 1) || (bs % 32 == 0);
    if (!x_aligned && bs >= 64 && prec != B2NN_PREC_FP32) {
        const int64_t bs_pad = round_up(bs, 32);
        const int64_t ldsp = int64_t(nb - 1) * bs_pad + round_up(last_rows, 32);
        if (!c->Xsp || c->sp_nb != nb || c->ldsp != ldsp) {
            drop_lo(c, c->Xsp);
            cudaFree(c->Xsp); cudaFree(c->Ysp); c->Xsp = c->Ysp = nullptr;
            B2_CUDA(cudaMalloc(&c->Xsp, sizeof(float) * ldsp * (c->n + 1)));
            B2_CUDA(cudaMalloc(&c->Ysp, sizeof(float) * ldsp * c->K));
            c->ldsp = ldsp; c->sp_nb = nb; c->sp_stride = bs_pad;
        }
        B2_CUDA(cudaMemsetAsync(c->Xsp, 0, sizeof(float) * ldsp * (c->n + 1), c->stream));
        B2_CUDA(cudaMemsetAsync(c->Ysp, 0, sizeof(float) * ldsp * c->K, c->stream));
        B2_CUDA(split_batches_launch(c->X, c->ldx, c->Xsp, ldsp, c->m, bs, bs_pad, nb, c->n + 1, c->stream));
        B2_CUDA(split_batches_launch(c->Y, c->ldy, c->Ysp, ldsp, c->m, bs, bs_pad, nb, c->K, c->stream));
        Xtrain = c->Xsp; Ytrain = c->Ysp; ld_train = ldsp; x_total = ldsp; b_stride = bs_pad; x_aligned = true;
    }

    if (is_comp_prec(prec) && d->act_kind != B2NN_ACT_CUSTOM) {
        rc = prepare_x3(c, Xtrain, ld_train * int64_t(c->n + 1), last_rows, false, prec == B2NN_PREC_BF16);
        if (rc) return rc;
    } else {
        c->w_lo_valid = false;       // the weights are about to change without their companions being refreshed
    }

    const int64_t launches0 = c->launches, tc0 = c->tc_launches;
    EventPair evs;
    B2_CUDA(cudaEventCreate(&evs.a)); B2_CUDA(cudaEventCreate(&evs.b));
    cudaEvent_t const ev0 = evs.a, ev1 = evs.b;
    B2_CUDA(cudaEventRecord(ev0, c->stream));
    c->trace = std::getenv("B2NN_TRACE") != nullptr;
    c->trace_step = 0;
    c->trace_from = int64_t(d->iters) * nb - 2;           // the last three steps
    struct TraceDump { b2nn_ctx* c; ~TraceDump() { trace_dump(c); c->trace = false; } } trace_guard{c};

    int issued = 0, reported = 0;
    std::vector<int> slot_iter;
    auto flush_logs = [&](bool block) -> int {
        while (reported < issued) {
            cudaEvent_t e = c->log_events[reported % c->log_events.size()];
            cudaError_t q = block ? cudaEventSynchronize(e) : cudaEventQuery(e);
            if (q == cudaErrorNotReady) { (void)cudaGetLastError(); break; }
            if (q != cudaSuccess) { set_error(std::string("log event: ") + cudaGetErrorString(q)); return B2NN_ERR_CUDA; }
            const int64_t rows0 = global_rows[0];
            double J = c->h_acc[2 + reported];
            J = d->classify ? J / double(rows0) : J / (2.0 * double(rows0));       // :554, :615
            if (log) log(log_user, slot_iter[reported], float(J));
            ++reported;
        }
        return B2NN_OK;
    };

    // Small networks are launch bound (config C1: ~8 kernels of a few microseconds per step): after one eager epoch the
    // remaining epochs replay a captured CUDA graph of the whole epoch (one variant with, one without the cost of batch 0).
    double step_work = 0.0;
    for (const LayerInfo& li : c->L) step_work += 6.0 * double(last_rows) * double(li.in + 1) * double(li.out);
    bool use_graph = !c->comm && d->act_kind != B2NN_ACT_CUSTOM && !custom_outcost(d) && !c->profile && nb <= 128 && d->iters > 2 &&
                     step_work < 2e10;
    if (const char* e = std::getenv("B2NN_GRAPH")) use_graph = use_graph && std::atoi(e) != 0;
    GraphExecs graphs;                                      // [0] no cost, [1] batch 0 reports its cost
    cudaGraphExec_t (&graph_exec)[2] = graphs.g;
    int64_t graph_launches[2] = {0, 0}, graph_tc[2] = {0, 0};
    auto destroy_graphs = [&]() { graphs.destroy(); };
    double* const graph_slot = c->d_acc + 1;                // fixed accumulator the captured batch-0 step writes
    auto run_epoch = [&](double* slot0) -> int {
        for (int b = 0; b < nb; ++b) {
            const int64_t rows = (b == nb - 1) ? last_rows : bs;
            int r = train_step(c, d, prec, Xtrain, ld_train, x_total, x_aligned, Ytrain, ld_train, int64_t(b) * b_stride, rows,
                               global_rows[b], b == 0 ? slot0 : nullptr);
            if (r) return r;
        }
        return B2NN_OK;
    };

    for (int it = 0; it < d->iters; ++it) {                                         // :540-568, :597-629
        const bool show = ((it + 1) % display == 0) && issued < n_disp;             // :550, :610 (batch 0 only)
        double* slot = nullptr;
        if (use_graph && it >= 1) {
            const int v = show ? 1 : 0;
            if (!graph_exec[v]) {
                const int64_t l0 = c->launches, t0 = c->tc_launches;
                cudaGraph_t graph = nullptr;
                B2_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
                if (show) cudaMemsetAsync(graph_slot, 0, sizeof(double), c->stream);
                rc = run_epoch(show ? graph_slot : nullptr);
                cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
                if (rc || ce != cudaSuccess) {
                    if (graph) cudaGraphDestroy(graph);
                    destroy_graphs();
                    if (!rc) { set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce)); rc = B2NN_ERR_CUDA; }
                    return rc;
                }
                ce = cudaGraphInstantiate(&graph_exec[v], graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { destroy_graphs(); set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); return B2NN_ERR_CUDA; }
                graph_launches[v] = c->launches - l0; graph_tc[v] = c->tc_launches - t0;
                c->launches = l0; c->tc_launches = t0;           // capturing is not launching
            }
            B2_CUDA(cudaGraphLaunch(graph_exec[v], c->stream));
            c->launches += graph_launches[v]; c->tc_launches += graph_tc[v];
            slot = show ? graph_slot : nullptr;
        } else {
            slot = show ? c->d_acc + 2 + issued : nullptr;
            rc = run_epoch(slot);
            if (rc) return rc;
        }
        if (show) {
            if (issued - reported >= int(c->log_events.size())) { rc = flush_logs(true); if (rc) return rc; }
            if (c->comm) {
                // the printed cost is the global batch-0 cost: sum the per-rank partial sums
                int r = g_nccl.AllReduce(slot, slot, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
                if (r != 0) { set_error("ncclAllReduce(cost) failed"); return B2NN_ERR_CUDA; }
            }
            B2_CUDA(cudaMemcpyAsync(c->h_acc + 2 + issued, slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA(cudaEventRecord(c->log_events[issued % c->log_events.size()], c->stream));
            slot_iter.push_back(it);
            ++issued;
        }
        rc = flush_logs(false);
        if (rc) return rc;
    }
    if (c->p2p && c->p2p_step > 0) {
        uint64_t mask = 0;       // gate only layers that went through the peer-memory exchange; other flags stay 0
        for (int j = 0; j < int(c->L.size()) && j < P2P_MAX_LAYERS; ++j) if (p2p_layer_was_fused(c, j)) mask |= uint64_t(1) << j;
        B2_CUDA(p2p_gate_layers_launch(c->p2p_flags + 1 * P2P_MAX_LAYERS * P2P_MAX_WORLD, mask, c->nranks, c->p2p_step, c->stream));
    }
    B2_CUDA(cudaEventRecord(ev1, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    if (c->p2p && c->upd_stream) B2_CUDA(cudaStreamSynchronize(c->upd_stream));
    destroy_graphs();
    rc = flush_logs(true);
    if (rc) return rc;
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    const int64_t loop_launches = c->launches - launches0, loop_tc = c->tc_launches - tc0;

    double final_cost = 0.0, errors = 0.0;
    if (!d->skip_final_cost) {                                                       // :574, :635
        rc = evaluate_device(c, d, prec, c->X, c->ldx, c->Y, c->ldy, nullptr, c->m, nullptr, nullptr, 0, &final_cost,
                             &errors);
        if (rc) return rc;
        if (c->comm) {
            // data-parallel: rank-local sums -> global cost / error count
            double loc[3] = {d->classify ? final_cost * double(c->m) : final_cost * 2.0 * double(c->m), errors, double(c->m)};
            DevBuf t;
            B2_CUDA(cudaMalloc(&t.p, sizeof(double) * 3));
            B2_CUDA(cudaMemcpyAsync(t.p, loc, sizeof(loc), cudaMemcpyHostToDevice, c->stream));
            int r = g_nccl.AllReduce(t.p, t.p, 3, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
            if (r != 0) { set_error("ncclAllReduce(final cost) failed"); return B2NN_ERR_CUDA; }
            B2_CUDA(cudaMemcpyAsync(loc, t.p, sizeof(loc), cudaMemcpyDeviceToHost, c->stream));
            B2_CUDA(cudaStreamSynchronize(c->stream));
            final_cost = d->classify ? loc[0] / loc[2] : loc[0] / (2.0 * loc[2]);
            errors = loc[1];
        }
    }
    auto t_call1 = std::chrono::steady_clock::now();
    if (result) {
        result->call_seconds = std::chrono::duration<double>(t_call1 - t_call0).count();
        result->loop_seconds = double(ms) * 1e-3;
        result->final_cost = final_cost;
        result->error_count = errors;
        result->steps = int64_t(d->iters) * nb;
        result->kernel_launches = loop_launches;
        result->gemm_tc_launches = loop_tc;
    }
    return B2NN_OK;
}

int b2nn_train_step_host(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, const float* y, int64_t rows,
                         int64_t global_rows, float* J_out) {
    int rc = validate_desc(c, d, true);
    if (rc) return rc;
    if (!x || !y || rows < 1) { set_error("train_step_host: bad arguments"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int prec = effective_precision(c, resolve_precision(d->precision), rows);
    const int n = c->layers[0], K = c->layers.back();
    if (rows > c->caps) {
        drop_lo(c, c->Xs);
        cudaFree(c->Xs); cudaFree(c->Ys); c->Xs = c->Ys = nullptr; c->caps = 0;
        c->lds = round_up(rows, 32);
        B2_CUDA(cudaMalloc(&c->Xs, sizeof(float) * c->lds * (n + 1)));
        B2_CUDA(cudaMalloc(&c->Ys, sizeof(float) * c->lds * K));
        B2_CUDA(cudaMemsetAsync(c->Xs, 0, sizeof(float) * c->lds * (n + 1), c->stream));
        B2_CUDA(cudaMemsetAsync(c->Ys, 0, sizeof(float) * c->lds * K, c->stream));
        B2_CUDA(fill_launch(c->Xs + int64_t(n) * c->lds, 1.0f, c->lds, c->stream));
        c->caps = rows;           // only a fully built pair of buffers claims its capacity
    }
    rc = ensure_workspace(c, rows, d->act_kind == B2NN_ACT_CUSTOM);
    if (rc) return rc;
    rc = ensure_acc(c, 3);
    if (rc) return rc;
    // H2D of this step's inputs (skipping the reference's ones column: the engine keeps its own ones row)
    rc = stage_x_from_host(c, x, rows, 0, rows, n, c->Xs, c->lds, false);
    if (rc) return rc;
    B2_CUDA(cudaMemcpy2DAsync(c->Ys, sizeof(float) * c->lds, y, sizeof(float) * rows, sizeof(float) * rows, size_t(K),
                              cudaMemcpyHostToDevice, c->stream));
    if (is_comp_prec(prec) && d->act_kind != B2NN_ACT_CUSTOM) {
        rc = prepare_x3(c, c->Xs, c->lds * int64_t(n + 1), rows, false, prec == B2NN_PREC_BF16);
        if (rc) return rc;
    } else {
        c->w_lo_valid = false;
    }
    double* slot = J_out ? c->d_acc + 2 : nullptr;
    if (slot) B2_CUDA(cudaMemsetAsync(slot, 0, sizeof(double), c->stream));
    const int64_t g = global_rows > 0 ? global_rows : rows;
    rc = train_step(c, d, prec, c->Xs, c->lds, rows, true, c->Ys, c->lds, 0, rows, g, slot);
    if (rc) return rc;
    if (slot) {
        if (c->comm) {
            int r = g_nccl.AllReduce(slot, slot, 1, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
            if (r != 0) { set_error("ncclAllReduce(cost) failed"); return B2NN_ERR_CUDA; }
        }
        B2_CUDA(cudaMemcpyAsync(c->h_acc + 2, slot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
        const double J = c->h_acc[2];
        *J_out = float(d->classify ? J / double(g) : J / (2.0 * double(g)));
    }
    return B2NN_OK;
}

int b2nn_train_host(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, const float* y, int64_t m, int n_plus_1, int k,
                    float* J_steps, b2nn_train_result* result) {
    auto t_call0 = std::chrono::steady_clock::now();
    int rc = validate_desc(c, d, true);
    if (rc) return rc;
    if (!x || !y || m < 1) { set_error("train_host: bad arguments"); return B2NN_ERR_ARG; }
    if (n_plus_1 != c->layers[0] + 1 || k != c->layers.back()) { set_error("data shape does not match the layers"); return B2NN_ERR_ARG; }
    if (d->batch_num < 1 || int64_t(d->batch_num) > m) { set_error("batch_num must be in [1, m]"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int n = n_plus_1 - 1, nb = d->batch_num;
    const int64_t bs = m / nb, last_rows = m - int64_t(nb - 1) * bs;
    const int prec = effective_precision(c, resolve_precision(d->precision), last_rows);
    const int64_t steps = int64_t(d->iters) * nb;

    if (!c->copy_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->loss_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->loss_stream, cudaStreamNonBlocking));
    if (!c->st_loss) B2_CUDA(cudaEventCreateWithFlags(&c->st_loss, cudaEventDisableTiming));
    // the 8-byte loss copies ride their own stream: a DMA between two steps on the compute stream costs a bubble per step
    const bool loss_inline = std::getenv("B2NN_LOSS_D2H_INLINE") != nullptr;
    for (int i = 0; i < 2; ++i) {
        if (!c->st_copied[i]) B2_CUDA(cudaEventCreateWithFlags(&c->st_copied[i], cudaEventDisableTiming));
        if (!c->st_consumed[i]) B2_CUDA(cudaEventCreateWithFlags(&c->st_consumed[i], cudaEventDisableTiming));
    }
    if (last_rows > c->capst) {
        for (int i = 0; i < 2; ++i) {
            drop_lo(c, c->Xst[i]);
            cudaFree(c->Xst[i]); cudaFree(c->Yst[i]); c->Xst[i] = c->Yst[i] = nullptr;
        }
        c->capst = 0;
        c->ldst = round_up(last_rows, 32);
        for (int i = 0; i < 2; ++i) {
            B2_CUDA(cudaMalloc(&c->Xst[i], sizeof(float) * c->ldst * n_plus_1));
            B2_CUDA(cudaMalloc(&c->Yst[i], sizeof(float) * c->ldst * k));
            B2_CUDA(cudaMemsetAsync(c->Xst[i], 0, sizeof(float) * c->ldst * n_plus_1, c->stream));
            B2_CUDA(cudaMemsetAsync(c->Yst[i], 0, sizeof(float) * c->ldst * k, c->stream));
            B2_CUDA(fill_launch(c->Xst[i] + int64_t(n) * c->ldst, 1.0f, c->ldst, c->stream));
        }
        c->capst = last_rows;     // only fully built buffers claim their capacity
    }
    if (steps > c->jsteps_cap) {
        cudaFree(c->d_jsteps); if (c->h_jsteps) cudaFreeHost(c->h_jsteps);
        c->d_jsteps = nullptr; c->h_jsteps = nullptr; c->jsteps_cap = 0;
        B2_CUDA(cudaMalloc(&c->d_jsteps, sizeof(double) * steps));
        B2_CUDA(cudaMallocHost(&c->h_jsteps, sizeof(double) * steps));
        c->jsteps_cap = steps;
    }
    rc = ensure_workspace(c, last_rows, d->act_kind == B2NN_ACT_CUSTOM);
    if (rc) return rc;
    const bool x3 = is_comp_prec(prec) && d->act_kind != B2NN_ACT_CUSTOM;
    if (x3) {
        for (int i = 0; i < 2; ++i) { rc = prepare_x3(c, c->Xst[i], c->ldst * int64_t(n_plus_1), last_rows, false, prec == B2NN_PREC_BF16); if (rc) return rc; }
    } else {
        c->w_lo_valid = false;
    }
    B2_CUDA(cudaMemsetAsync(c->V, 0, sizeof(float) * c->total_int, c->stream));
    B2_CUDA(cudaMemsetAsync(c->d_jsteps, 0, sizeof(double) * steps, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));                 // staging buffers initialised before the copy stream touches them

    auto batch_rows = [&](int b) { return b == nb - 1 ? last_rows : bs; };
    // data parallel: every rank streams its own shard; the update uses the GLOBAL rows of each batch
    std::vector<int64_t> global_rows(nb);
    for (int b = 0; b < nb; ++b) global_rows[b] = batch_rows(b);
    if (c->comm) {
        DevBuf tmp;
        B2_CUDA(cudaMalloc(&tmp.p, sizeof(int64_t) * nb));
        B2_CUDA(cudaMemcpyAsync(tmp.p, global_rows.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, c->stream));
        int r = g_nccl.AllReduce(tmp.p, tmp.p, size_t(nb), NCCL_INT64, NCCL_SUM, c->comm, c->stream);
        if (r != 0) { set_error("ncclAllReduce(batch rows) failed"); return B2NN_ERR_CUDA; }
        B2_CUDA(cudaMemcpyAsync(global_rows.data(), tmp.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
        while (c->grad_ready.size() < c->L.size()) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_ready.push_back(e); }
        while (c->grad_reduced.size() < c->L.size()) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_reduced.push_back(e); }
    }
    auto issue_copy = [&](int64_t s) -> int {                   // H2D of step s's batch into staging buffer s & 1
        const int buf = int(s & 1), b = int(s % nb);
        const int64_t rows = batch_rows(b), start = int64_t(b) * bs;
        if (s >= 2) B2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->st_consumed[buf], 0));
        // column f+1 of the reference matrix is feature row f of the engine matrix (its ones column is not copied)
        B2_CUDA(cudaMemcpy2DAsync(c->Xst[buf], sizeof(float) * c->ldst, x + m + start, sizeof(float) * m, sizeof(float) * rows,
                                  size_t(n), cudaMemcpyHostToDevice, c->copy_stream));
        B2_CUDA(cudaMemcpy2DAsync(c->Yst[buf], sizeof(float) * c->ldst, y + start, sizeof(float) * m, sizeof(float) * rows,
                                  size_t(k), cudaMemcpyHostToDevice, c->copy_stream));
        B2_CUDA(cudaEventRecord(c->st_copied[buf], c->copy_stream));
        return B2NN_OK;
    };

    const int64_t launches0 = c->launches, tc0 = c->tc_launches;
    EventPair evs;
    B2_CUDA(cudaEventCreate(&evs.a)); B2_CUDA(cudaEventCreate(&evs.b));
    cudaEvent_t const ev0 = evs.a, ev1 = evs.b;
    B2_CUDA(cudaEventRecord(ev0, c->stream));
    rc = issue_copy(0);
    if (rc) return rc;
    for (int64_t s = 0; s < steps; ++s) {
        const int buf = int(s & 1), b = int(s % nb);
        if (s + 1 < steps) { rc = issue_copy(s + 1); if (rc) return rc; }      // next batch travels while this one computes
        B2_CUDA(cudaStreamWaitEvent(c->stream, c->st_copied[buf], 0));
        if (x3) B2_CUDA(make_companion(c->Xst[buf], lo_buffer(c, c->Xst[buf], c->ldst * int64_t(n_plus_1)), c->ldst * int64_t(n_plus_1), c->comp_bf16, c->stream));
        rc = train_step(c, d, prec, c->Xst[buf], c->ldst, c->ldst, true, c->Yst[buf], c->ldst, 0, batch_rows(b), global_rows[b],
                        c->d_jsteps + s);
        if (rc) return rc;
        B2_CUDA(cudaEventRecord(c->st_consumed[buf], c->stream));
        cudaStream_t ls = loss_inline ? c->stream : c->loss_stream;
        if (!loss_inline) B2_CUDA(cudaStreamWaitEvent(ls, c->st_consumed[buf], 0));
        B2_CUDA(cudaMemcpyAsync(c->h_jsteps + s, c->d_jsteps + s, sizeof(double), cudaMemcpyDeviceToHost, ls));   // this step's loss
    }
    if (!loss_inline) {          // the timed region (and the caller) ends after the last loss has landed
        B2_CUDA(cudaEventRecord(c->st_loss, c->loss_stream));
        B2_CUDA(cudaStreamWaitEvent(c->stream, c->st_loss, 0));
    }
    if (c->p2p && c->p2p_step > 0) {
        uint64_t mask = 0;       // gate only layers that went through the peer-memory exchange; other flags stay 0
        for (int j = 0; j < int(c->L.size()) && j < P2P_MAX_LAYERS; ++j) if (p2p_layer_was_fused(c, j)) mask |= uint64_t(1) << j;
        B2_CUDA(p2p_gate_layers_launch(c->p2p_flags + 1 * P2P_MAX_LAYERS * P2P_MAX_WORLD, mask, c->nranks, c->p2p_step, c->stream));
    }
    B2_CUDA(cudaEventRecord(ev1, c->stream));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    if (c->p2p && c->upd_stream) B2_CUDA(cudaStreamSynchronize(c->upd_stream));
    float ms = 0.f;
    B2_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (J_steps)
        for (int64_t s = 0; s < steps; ++s) {
            const double rows = double(batch_rows(int(s % nb)));
            J_steps[s] = float(d->classify ? c->h_jsteps[s] / rows : c->h_jsteps[s] / (2.0 * rows));
        }
    auto t_call1 = std::chrono::steady_clock::now();
    if (result) {
        std::memset(result, 0, sizeof(*result));
        result->call_seconds = std::chrono::duration<double>(t_call1 - t_call0).count();
        result->loop_seconds = double(ms) * 1e-3;
        result->steps = steps;
        result->kernel_launches = c->launches - launches0;
        result->gemm_tc_launches = c->tc_launches - tc0;
    }
    return B2NN_OK;
}

static int evaluate_impl(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, const float* y, int64_t rows, int y_is_labels,
                         double* cost_out, double* errors_out, bool raw, bool normalise);

int b2nn_evaluate(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, const float* y, int64_t rows, int y_is_labels,
                  double* cost_out, double* errors_out) {
    return evaluate_impl(c, d, x, y, rows, y_is_labels, cost_out, errors_out, false, false);
}
int b2nn_evaluate_raw(b2nn_ctx* c, const b2nn_train_desc* d, const float* x_raw, const float* y, int64_t rows, int y_is_labels,
                      int normalise, double* cost_out, double* errors_out) {
    return evaluate_impl(c, d, x_raw, y, rows, y_is_labels, cost_out, errors_out, true, normalise != 0);
}

static int evaluate_impl(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, const float* y, int64_t rows, int y_is_labels,
                         double* cost_out, double* errors_out, bool raw, bool normalise) {
    int rc = validate_desc(c, d, false);
    if (rc) return rc;
    if (!x || !y || rows < 1) { set_error("evaluate: bad arguments"); return B2NN_ERR_ARG; }
    if (normalise && c->norm_n != c->layers[0]) { set_error("evaluate_raw: no normalisation statistics (b2nn_set_train_data_raw / b2nn_set_normalisation first)"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int prec = effective_precision(c, resolve_precision(d->precision), rows);
    const int n = c->layers[0], K = c->layers.back();
    const int64_t ld = round_up(rows, 32);
    DevBuf xbuf, ybuf, labbuf;
    B2_CUDA(cudaMalloc(&xbuf.p, sizeof(float) * ld * (n + 1)));
    B2_CUDA(cudaMalloc(&ybuf.p, sizeof(float) * ld * K));
    float *xd = xbuf.as<float>(), *yd = ybuf.as<float>(), *lab = nullptr;
    B2_CUDA(cudaMemsetAsync(xd, 0, sizeof(float) * ld * (n + 1), c->stream));
    rc = stage_x_from_host(c, x, rows, 0, rows, n, xd, ld, true, raw);
    if (!rc && normalise) rc = normalise_rows_launch(xd, ld, rows, n, c->d_mean, c->d_std, c->stream) == cudaSuccess ? 0 : B2NN_ERR_CUDA;
    if (!rc) {
        if (y_is_labels && d->classify) {
            // validate(): labels -> one-hot for the cost, raw labels for the error count (cublasNN.cpp:788-790, 859)
            rc = cudaMalloc(&labbuf.p, sizeof(float) * rows) == cudaSuccess ? 0 : B2NN_ERR_ALLOC;
            lab = labbuf.as<float>();
            if (!rc) rc = cudaMemcpyAsync(lab, y, sizeof(float) * rows, cudaMemcpyHostToDevice, c->stream) == cudaSuccess ? 0 : B2NN_ERR_CUDA;
            if (!rc) rc = labels_to_onehot_launch(lab, yd, ld, rows, K, c->stream) == cudaSuccess ? 0 : B2NN_ERR_CUDA;
        } else {
            rc = cudaMemcpy2DAsync(yd, sizeof(float) * ld, y, sizeof(float) * rows, sizeof(float) * rows, size_t(K),
                                   cudaMemcpyHostToDevice, c->stream) == cudaSuccess ? 0 : B2NN_ERR_CUDA;
        }
    }
    if (!rc) rc = evaluate_device(c, d, prec, xd, ld, yd, ld, lab, rows, nullptr, nullptr, 0, cost_out, errors_out);
    cudaStreamSynchronize(c->stream);
    c->plans.clear();      // plans reference xd
    drop_lo(c, xd);
    return rc;
}

// Copy a [rows x cols] block of floats between two pitched host matrices, split over a few threads when it is large
// (one thread packs ~10 GB/s; PCIe gen5 moves ~50).
static void host_copy_2d(float* dst, int64_t dst_pitch, const float* src, int64_t src_pitch, int64_t rows, int64_t cols) {
    auto work = [=](int64_t c0, int64_t c1) {
        for (int64_t j = c0; j < c1; ++j) std::memcpy(dst + j * dst_pitch, src + j * src_pitch, sizeof(float) * size_t(rows));
    };
    const int64_t bytes = rows * cols * int64_t(sizeof(float));
    int nt = int(std::min<int64_t>(std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency())), cols));
    if (bytes < (int64_t(2) << 20) || nt <= 1) { work(0, cols); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, cols * t / nt, cols * (t + 1) / nt);
    work(0, cols / nt);
    for (auto& t : th) t.join();
}

// predict (cublasNN.cpp:906-977: cudaMalloc, one pageable H2D of the whole matrix, forward, one D2H).  Here the rows
// travel in chunks through two pinned staging buffers and two device buffers: chunk c+1 is packed and copied H2D while
// chunk c computes and chunk c-1's results come back, so a large batch runs at the PCIe rate and never needs more than
// two chunks of device memory.  A source that is already pinned (cudaMallocHost / cudaHostRegister) is DMA'd directly.
static int predict_impl(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, int64_t rows, float* out, float* probs, bool raw,
                        bool normalise);

int b2nn_predict(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, int64_t rows, float* out, float* probs) {
    return predict_impl(c, d, x, rows, out, probs, false, false);
}
int b2nn_predict_raw(b2nn_ctx* c, const b2nn_train_desc* d, const float* x_raw, int64_t rows, int normalise, float* out, float* probs) {
    return predict_impl(c, d, x_raw, rows, out, probs, true, normalise != 0);
}

static int predict_impl(b2nn_ctx* c, const b2nn_train_desc* d, const float* x, int64_t rows, float* out, float* probs, bool raw,
                        bool normalise) {
    int rc = validate_desc(c, d, false);
    if (rc) return rc;
    if (!x || !out || rows < 1) { set_error("predict: bad arguments"); return B2NN_ERR_ARG; }
    if (normalise && c->norm_n != c->layers[0]) { set_error("predict_raw: no normalisation statistics (b2nn_set_train_data_raw / b2nn_set_normalisation first)"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int prec = effective_precision(c, resolve_precision(d->precision), rows);
    const int n = c->layers[0], K = c->layers.back();
    // ~48 MB of input per chunk: long enough for the full PCIe rate, short enough that the pipeline fills quickly
    int64_t chunk = std::min<int64_t>(int64_t(1) << 17, std::max<int64_t>(1024, ((int64_t(48) << 20) / (4 * int64_t(n))) / 32 * 32));
    if (const char* e = std::getenv("B2NN_PREDICT_CHUNK")) chunk = std::max<int64_t>(32, std::atoll(e) / 32 * 32);
    if (rows <= chunk + chunk / 4) chunk = round_up(rows, 32);
    const int64_t ld = chunk;
    cudaPointerAttributes pa{};
    const bool src_pinned = cudaPointerGetAttributes(&pa, x) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (ld > c->pr_ld || n != c->pr_n || K != c->pr_k) {
        B2_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 2; ++i) {
            drop_lo(c, c->pr_x[i]);
            cudaFree(c->pr_x[i]); cudaFree(c->pr_p[i]); cudaFree(c->pr_h[i]); c->pr_x[i] = c->pr_p[i] = c->pr_h[i] = nullptr;
            if (c->pr_hx[i]) cudaFreeHost(c->pr_hx[i]);
            if (c->pr_hout[i]) cudaFreeHost(c->pr_hout[i]);
            c->pr_hx[i] = c->pr_hout[i] = nullptr;
        }
        c->pr_ld = 0;
        c->plans.clear();
        const int nbuf = rows > chunk ? 2 : 1;
        for (int i = 0; i < nbuf; ++i) {
            B2_CUDA(cudaMalloc(&c->pr_x[i], sizeof(float) * ld * (n + 1)));
            B2_CUDA(cudaMalloc(&c->pr_p[i], sizeof(float) * ld));
            B2_CUDA(cudaMalloc(&c->pr_h[i], sizeof(float) * ld * K));
            B2_CUDA(cudaMallocHost(&c->pr_hout[i], sizeof(float) * ld * (K + 1)));
            B2_CUDA(cudaMemsetAsync(c->pr_x[i], 0, sizeof(float) * ld * (n + 1), c->stream));
            B2_CUDA(fill_launch(c->pr_x[i] + int64_t(n) * ld, 1.0f, ld, c->stream));
        }
        c->pr_ld = ld; c->pr_n = n; c->pr_k = K;
        B2_CUDA(cudaStreamSynchronize(c->stream));           // the copy stream may touch the buffers next
    }
    const int64_t ldp = c->pr_ld;
    const int64_t nchunks = (rows + chunk - 1) / chunk;
    if (nchunks > 1 && !c->pr_x[1]) {                         // the buffers were sized by a single-chunk call
        B2_CUDA(cudaMalloc(&c->pr_x[1], sizeof(float) * ldp * (n + 1)));
        B2_CUDA(cudaMalloc(&c->pr_p[1], sizeof(float) * ldp));
        B2_CUDA(cudaMalloc(&c->pr_h[1], sizeof(float) * ldp * K));
        B2_CUDA(cudaMallocHost(&c->pr_hout[1], sizeof(float) * ldp * (K + 1)));
        B2_CUDA(cudaMemsetAsync(c->pr_x[1], 0, sizeof(float) * ldp * (n + 1), c->stream));
        B2_CUDA(fill_launch(c->pr_x[1] + int64_t(n) * ldp, 1.0f, ldp, c->stream));
        B2_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (!src_pinned)
        for (int i = 0; i < (nchunks > 1 ? 2 : 1); ++i)
            if (!c->pr_hx[i]) B2_CUDA(cudaMallocHost(&c->pr_hx[i], sizeof(float) * ldp * n));
    if (!c->copy_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->pr_d2h_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->pr_d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        if (!c->pr_copied[i]) B2_CUDA(cudaEventCreateWithFlags(&c->pr_copied[i], cudaEventDisableTiming));
        if (!c->pr_computed[i]) B2_CUDA(cudaEventCreateWithFlags(&c->pr_computed[i], cudaEventDisableTiming));
        if (!c->pr_out[i]) B2_CUDA(cudaEventCreateWithFlags(&c->pr_out[i], cudaEventDisableTiming));
    }
    const bool want_h = !d->classify || probs != nullptr;
    auto drain = [&](int64_t ci) -> int {                     // chunk ci's results: pinned staging -> the caller's arrays
        const int buf = int(ci & 1);
        const int64_t s0 = ci * chunk, r = std::min<int64_t>(chunk, rows - s0);
        B2_CUDA(cudaEventSynchronize(c->pr_out[buf]));
        const float* ho = c->pr_hout[buf];
        if (d->classify) std::memcpy(out + s0, ho, sizeof(float) * size_t(r));
        else host_copy_2d(out + s0, rows, ho + ldp, ldp, r, K);
        if (probs) host_copy_2d(probs + s0, rows, ho + ldp, ldp, r, K);
        return B2NN_OK;
    };
    for (int64_t ci = 0; ci < nchunks; ++ci) {
        const int buf = int(ci & 1);
        const int64_t s0 = ci * chunk, r = std::min<int64_t>(chunk, rows - s0);
        // column f+1 of the reference matrix (its ones column is skipped) is feature row f of the engine matrix
        const float* src = x + (raw ? 0 : rows) + s0;
        if (ci >= 2) B2_CUDA(cudaStreamWaitEvent(c->copy_stream, c->pr_computed[buf], 0));    // device buffer free again
        if (src_pinned) {
            B2_CUDA(cudaMemcpy2DAsync(c->pr_x[buf], sizeof(float) * ldp, src, sizeof(float) * rows, sizeof(float) * r, size_t(n),
                                      cudaMemcpyHostToDevice, c->copy_stream));
        } else {
            if (ci >= 2) B2_CUDA(cudaEventSynchronize(c->pr_copied[buf]));                      // pinned staging free again
            host_copy_2d(c->pr_hx[buf], r, src, rows, r, n);
            B2_CUDA(cudaMemcpy2DAsync(c->pr_x[buf], sizeof(float) * ldp, c->pr_hx[buf], sizeof(float) * r, sizeof(float) * r, size_t(n),
                                      cudaMemcpyHostToDevice, c->copy_stream));
        }
        B2_CUDA(cudaEventRecord(c->pr_copied[buf], c->copy_stream));
        B2_CUDA(cudaStreamWaitEvent(c->stream, c->pr_copied[buf], 0));
        if (ci >= 2) B2_CUDA(cudaStreamWaitEvent(c->stream, c->pr_out[buf], 0));               // previous results have left
        if (normalise) B2_CUDA(normalise_rows_launch(c->pr_x[buf], ldp, r, n, c->d_mean, c->d_std, c->stream));
        rc = evaluate_device(c, d, prec, c->pr_x[buf], ldp, nullptr, 0, nullptr, r, d->classify ? c->pr_p[buf] : nullptr,
                             want_h ? c->pr_h[buf] : nullptr, ldp, nullptr, nullptr, ldp);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        B2_CUDA(cudaEventRecord(c->pr_computed[buf], c->stream));
        B2_CUDA(cudaStreamWaitEvent(c->pr_d2h_stream, c->pr_computed[buf], 0));
        if (d->classify)
            B2_CUDA(cudaMemcpyAsync(c->pr_hout[buf], c->pr_p[buf], sizeof(float) * r, cudaMemcpyDeviceToHost, c->pr_d2h_stream));
        if (want_h)
            B2_CUDA(cudaMemcpy2DAsync(c->pr_hout[buf] + ldp, sizeof(float) * ldp, c->pr_h[buf], sizeof(float) * ldp, sizeof(float) * r,
                                      size_t(K), cudaMemcpyDeviceToHost, c->pr_d2h_stream));
        B2_CUDA(cudaEventRecord(c->pr_out[buf], c->pr_d2h_stream));
        if (ci >= 1) { rc = drain(ci - 1); if (rc) return rc; }
    }
    return drain(nchunks - 1);
}

int b2nn_predict_device(b2nn_ctx* c, const b2nn_train_desc* d, const float* x_dev, int64_t rows, int64_t ld, float* out_dev) {
    int rc = validate_desc(c, d, false);
    if (rc) return rc;
    if (!x_dev || !out_dev || rows < 1 || ld < rows) { set_error("predict_device: bad arguments"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    const int prec = effective_precision(c, resolve_precision(d->precision), rows);
    // x_dev is already in ENGINE layout: (l_0+1) rows of ld floats, ones row last.
    if (d->classify) return evaluate_device(c, d, prec, x_dev, ld, nullptr, 0, nullptr, rows, out_dev, nullptr, 0, nullptr, nullptr);
    return evaluate_device(c, d, prec, x_dev, ld, nullptr, 0, nullptr, rows, nullptr, out_dev, ld, nullptr, nullptr);
}

int b2nn_gemm(b2nn_ctx* c, int backend, const float* A, int64_t lda, int a_major, const float* B, int64_t ldb, int b_major,
              float* D, int64_t ldd, int M, int N, int K, int epi, const float* aux, int64_t ldaux) {
    if (!c || !A || !B || !D) { set_error("gemm: null argument"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    GemmDesc g;
    g.A = A; g.lda = lda; g.a_major = a_major; g.B = B; g.ldb = ldb; g.b_major = b_major;
    g.D = D; g.ldd = ldd; g.M = M; g.N = N; g.K = K; g.epi = epi; g.aux = aux; g.ldaux = ldaux;
    if (backend == 1 || backend == 3 || backend == 5 || backend == 7 || backend == 9 || backend == 11) {
        g.mn_slack = (backend & 2) != 0;
        const bool bf = (backend & 8) != 0;
        DevBuf alo, blo;
        if (backend & 12) {         // fp32-grade / bf16 mode: the operands' companions go into temporaries first
            // the caller's slack (32 floats past the last row) is converted too; the temporaries carry a little more,
            // since the folded bf16 view may touch one 64-element box past the end (feeding rows that are never stored)
            const int64_t na = (a_major ? int64_t(K) : int64_t(M)) * lda + (g.mn_slack ? 32 : 0);
            const int64_t nbb = (b_major ? int64_t(K) : int64_t(N)) * ldb + (g.mn_slack ? 32 : 0);
            B2_CUDA(cudaMalloc(&alo.p, sizeof(float) * (na + 64)));
            B2_CUDA(cudaMalloc(&blo.p, sizeof(float) * (nbb + 64)));
            B2_CUDA(cudaMemsetAsync(alo.p, 0, sizeof(float) * (na + 64), c->stream));
            B2_CUDA(cudaMemsetAsync(blo.p, 0, sizeof(float) * (nbb + 64), c->stream));
            B2_CUDA(make_companion(A, alo.as<float>(), na, bf, c->stream));
            B2_CUDA(make_companion(B, blo.as<float>(), nbb, bf, c->stream));
            g.A_lo = alo.as<float>(); g.B_lo = blo.as<float>(); g.bf16 = bf;
        }
        TcPlan plan;
        std::string err;
        if (!tc_plan_build(plan, g, err)) { set_error("tcgen05 plan: " + err); return B2NN_ERR_ARG; }
        cudaError_t le = tc_plan_launch(plan, c->stream);
        if (backend & 12) cudaStreamSynchronize(c->stream);     // the temporaries are freed when this scope ends
        B2_CUDA(le);
        ++c->tc_launches;
    } else {
        B2_CUDA(gemm_simt_launch(g, c->stream));
    }
    ++c->launches;
    return B2NN_OK;
}

int b2nn_momentum_update(b2nn_ctx* c, float* theta, float* vel, const float* grad, int64_t count, float mu, float alpha) {
    if (!c || !theta || !vel || !grad) { set_error("momentum_update: null argument"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(momentum_update_launch(theta, vel, grad, count, mu, alpha, c->stream));
    ++c->launches;
    return B2NN_OK;
}

void* b2nn_stream(b2nn_ctx* c) { return c ? (void*)c->stream : nullptr; }

int b2nn_profile_enable(b2nn_ctx* c, int on) {
    if (!c) { set_error("null ctx"); return B2NN_ERR_ARG; }
    c->profile = on != 0;
    c->prof_chain = false;
    return B2NN_OK;
}

int b2nn_profile_read(b2nn_ctx* c, double* ms, double* flops, double* bytes, int64_t* launches, int64_t* tc_launches) {
    if (!c || !ms || !flops || !bytes || !launches || !tc_launches) { set_error("profile_read: null argument"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    B2_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < B2NN_PROFILE_KINDS; ++k) { ms[k] = 0; flops[k] = 0; bytes[k] = 0; launches[k] = 0; tc_launches[k] = 0; }
    for (auto& r : c->prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess && r.kind >= 0 && r.kind < B2NN_PROFILE_KINDS) {
            ms[r.kind] += t; flops[r.kind] += r.flops; bytes[r.kind] += r.bytes; ++launches[r.kind];
            if (r.tc) ++tc_launches[r.kind];
        }
        if (r.own_e0) c->prof_pool.push_back(r.e0);
        c->prof_pool.push_back(r.e1);
    }
    (void)cudaGetLastError();
    c->prof.clear();
    c->prof_chain = false;
    return B2NN_OK;
}

int b2nn_comm_unique_id(void* id128) {
    if (!id128) { set_error("null id"); return B2NN_ERR_ARG; }
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl_err.empty()) { set_error(g_nccl_err); return B2NN_ERR_CUDA; }
    int r = g_nccl.GetUniqueId(id128);
    if (r != 0) { set_error("ncclGetUniqueId failed"); return B2NN_ERR_CUDA; }
    return B2NN_OK;
}

int b2nn_comm_init(b2nn_ctx* c, int rank, int nranks, const void* id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_error("comm_init: bad arguments"); return B2NN_ERR_ARG; }
    if (nranks == 1) return B2NN_OK;
    if (c->comm) { set_error("comm_init: this context already has a communicator"); return B2NN_ERR_ARG; }
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl_err.empty()) { set_error(g_nccl_err); return B2NN_ERR_CUDA; }
    B2_CUDA(cudaSetDevice(c->device));
    Id128 id;
    std::memcpy(id.bytes, id128, 128);
    // The GEMMs are persistent kernels that fill every SM they are given; the gradient all-reduce runs beside them on
    // its own stream.  Bound NCCL to a few CTAs and keep that many SMs out of the GEMM grids so neither waits for the other.
    int reserve = 0;            // measured at N = 4 (2.23-2.27 ms/step for 0, 8 and 16 reserved SMs): no gain, so nothing is reserved by default
    if (const char* e = std::getenv("B2NN_COMM_SMS")) reserve = std::atoi(e);
    if (reserve > 0) {
        setenv("NCCL_MAX_CTAS", std::to_string(reserve).c_str(), 0);
        tc_set_reserved_sms(reserve);
        c->plans.clear();
    }
    void* comm = nullptr;
    int r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != 0) { set_error(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?")); return B2NN_ERR_CUDA; }
    c->comm = comm; c->rank = rank; c->nranks = nranks;
    const size_t nw = c->L.size() ? c->L.size() : 16;
    while (c->grad_ready.size() < nw) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_ready.push_back(e); }
    while (c->grad_reduced.size() < nw) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->grad_reduced.push_back(e); }
    return B2NN_OK;
}

int b2nn_comm_p2p_export(b2nn_ctx* c, void* blob192) {
    if (!c || !blob192) { set_error("p2p_export: null argument"); return B2NN_ERR_ARG; }
    if (!c->W || !c->G) { set_error("p2p_export: set_layers must come first"); return B2NN_ERR_ARG; }
    if (c->nranks < 2 || c->nranks > P2P_MAX_WORLD) { set_error("p2p_export: needs 2..8 ranks (b2nn_comm_init first)"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    if (!c->p2p_flags) {
        // one peer-visible allocation: the flag block, then `world` slots for every small layer (a layer whose rows do
        // not split into whole 256-column tiles per rank, up to 4 Mi floats)
        const int nw = int(c->L.size());
        c->p2p_small_off.assign(nw, -1);
        c->p2p_small_cnt.assign(nw, 0);
        c->p2p_small_floats = 0;
        for (int j = 0; j < nw && j < P2P_MAX_LAYERS; ++j) {
            const int64_t cnt = c->L[j].ldw * c->L[j].out;
            if (c->L[j].out % (256 * c->nranks) == 0 || cnt > (int64_t(4) << 20) || (c->L[j].w_off & 3)) continue;
            c->p2p_small_off[j] = c->p2p_small_floats;
            c->p2p_small_cnt[j] = cnt;
            c->p2p_small_floats += cnt * c->nranks;
        }
        const size_t bytes = sizeof(int) * P2P_FLAG_INTS + sizeof(float) * size_t(c->p2p_small_floats);
        B2_CUDA(cudaMalloc(&c->p2p_flags, bytes));
        B2_CUDA(cudaMemset(c->p2p_flags, 0, bytes));
    }
    cudaIpcMemHandle_t h[3];
    B2_CUDA(cudaIpcGetMemHandle(&h[0], c->W));
    B2_CUDA(cudaIpcGetMemHandle(&h[1], c->G));
    B2_CUDA(cudaIpcGetMemHandle(&h[2], c->p2p_flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::memcpy(blob192, h, 192);
    return B2NN_OK;
}

int b2nn_comm_p2p_open(b2nn_ctx* c, const void* blobs) {
    if (!c) { set_error("p2p_open: null context"); return B2NN_ERR_ARG; }
    if (!blobs) { c->p2p = false; c->plans.clear(); return B2NN_OK; }      // switch back to the NCCL path
    if (!c->p2p_flags) { set_error("p2p_open: call b2nn_comm_p2p_export first"); return B2NN_ERR_ARG; }
    B2_CUDA(cudaSetDevice(c->device));
    if (!c->p2p_opened.empty()) {       // opened before: an IPC handle maps once per process, close the old imports first
        B2_CUDA(cudaStreamSynchronize(c->stream));
        if (c->upd_stream) B2_CUDA(cudaStreamSynchronize(c->upd_stream));
        for (void* p : c->p2p_opened) cudaIpcCloseMemHandle(p);
        c->p2p_opened.clear();
        c->p2p = false;
    }
    const char* b = static_cast<const char*>(blobs);
    c->peers.world = c->nranks; c->peers.rank = c->rank;
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) {
            c->peers.W[r] = c->W; c->peers.G[r] = c->G; c->peers.flags[r] = c->p2p_flags;
            c->peers.S[r] = reinterpret_cast<float*>(c->p2p_flags + P2P_FLAG_INTS);
            continue;
        }
        cudaIpcMemHandle_t h[3];
        std::memcpy(h, b + size_t(r) * 192, 192);
        void* ptr[3] = {nullptr, nullptr, nullptr};
        for (int i = 0; i < 3; ++i) {
            cudaError_t e = cudaIpcOpenMemHandle(&ptr[i], h[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); (void)cudaGetLastError(); return B2NN_ERR_CUDA; }
            c->p2p_opened.push_back(ptr[i]);
        }
        c->peers.W[r] = static_cast<float*>(ptr[0]); c->peers.G[r] = static_cast<float*>(ptr[1]); c->peers.flags[r] = static_cast<int*>(ptr[2]);
        c->peers.S[r] = reinterpret_cast<float*>(c->peers.flags[r] + P2P_FLAG_INTS);
    }
    if (!c->upd_stream) B2_CUDA(cudaStreamCreateWithFlags(&c->upd_stream, cudaStreamNonBlocking));
    if (!c->p2p_upd_done) B2_CUDA(cudaEventCreateWithFlags(&c->p2p_upd_done, cudaEventDisableTiming));
    while (c->p2p_wg_ev.size() < size_t(P2P_MAX_LAYERS)) { cudaEvent_t e; B2_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->p2p_wg_ev.push_back(e); }
    c->p2p = true;
    c->plans.clear();
    return B2NN_OK;
}

int b2nn_comm_barrier(b2nn_ctx* c) {
    if (!c) { set_error("null ctx"); return B2NN_ERR_ARG; }
    return comm_barrier_impl(c);
}

}  // extern "C"
