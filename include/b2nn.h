/* b2nn.h — C ABI of the B200-native MLP training engine.
 *
 * This is the drop-in boundary for the ONE hot path of HenryJia/CUBLAS-ML: the mini-batch forward / delta-backprop /
 * momentum-update loop of class cublasNN (reference cublasNN.cpp:526-685) plus the setup and evaluation calls that
 * bracket it.  The reference has no FFI layer; its boundary is the C++ class in cublasNN.h.  The source-compatible
 * façade in include/cublasNN.h is a thin C++ wrapper over exactly these entry points; any other host language binds
 * them the same way (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, non-zero on failure; b2nn_last_error() explains.
 *   - matrices crossing the ABI use the REFERENCE host layout: column-major, sample index fastest
 *     (IDX2C(i,j,ld) = j*ld+i, cublasNN.h:17).  X carries the leading ones column written by addBias
 *     (cublasNN.cpp:330-346).  Weights cross in the reference arena layout: Theta_j is (l_j+1) x l_{j+1} column-major,
 *     row 0 = bias weights, layers concatenated (cublasNN.cpp:293-301).
 *   - there is NO CPU fallback: every compute entry point fails with B2NN_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef B2NN_H
#define B2NN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2nn_ctx b2nn_ctx;

enum {
    B2NN_OK = 0,
    B2NN_ERR_ARG = 1,        /* bad argument / call order (reference: undefined behaviour) */
    B2NN_ERR_NO_DEVICE = 2,  /* no CUDA device, or not compute capability 10.x */
    B2NN_ERR_CUDA = 3,       /* a CUDA / cuRAND / NCCL call failed */
    B2NN_ERR_ALLOC = 4       /* reference prints "cudaMalloc Failed" (cublasNN.cpp:307-311) */
};

/* Hidden activation pair = (activationHidden, activationDerivative) of train*Momentum (cublasNN.h:45-50). */
enum { B2NN_ACT_SIGMOID = 0 /* sigmoidGPU + sigmoidGradGPU, activations.cu:34-49 */,
       B2NN_ACT_TANH = 1    /* tanhGPU + sechSqGPU,        activations.cu:51-66 */,
       B2NN_ACT_CUSTOM = 2  /* user function pointers, generic unfused path       */ };
/* Output activation (activationOutput, cublasNN.h:49). */
enum { B2NN_OUT_SIGMOID = 0 /* sigmoidOutputGPU, activations.cu:9-12 */,
       B2NN_OUT_SOFTMAX = 1 /* softmaxGPU, activations.cu:68-88      */,
       B2NN_OUT_LINEAR = 2  /* trainFuncApproxMomentum: raw z (cublasNN.cpp:548) */,
       B2NN_OUT_CUSTOM = 3 };
/* Reported cost (costFunction, cublasNN.h:50).  Never influences gradients (cublasNN.cpp:666-667). */
enum { B2NN_COST_NEGLNMAX = 0     /* negLnMaxCostGPU, costfunctions.cu:14-19     */,
       B2NN_COST_CROSSENTROPY = 1 /* crossEntropyCostGPU, costfunctions.cu:21-26 */,
       B2NN_COST_SQUARED = 2      /* cublasSdot(delta,delta)/(2 mB), cublasNN.cpp:552-554 */,
       B2NN_COST_CUSTOM = 3 };
/* Weight initialisers (randinitweights.cu:13-31). */
enum { B2NN_INIT_1 = 1 /* eps = 1/sqrt(in)          */, B2NN_INIT_2 = 2 /* eps = sqrt(6)/sqrt(in+out) */,
       B2NN_INIT_CUSTOM = 3 };
/* Arithmetic of the three contractions (forward, delta backprop, weight gradient). */
enum { B2NN_PREC_FP32 = 0 /* fp32 FFMA, matches cublasSgemm default math up to summation order          */,
       B2NN_PREC_TF32 = 1 /* tcgen05 kind::tf32 (inputs truncated to 10 mantissa bits), fp32 accumulate in TMEM; small layers
                             stay on fp32 FFMA.  OPT-IN: narrower than the reference's arithmetic (stated bound 2e-2)  */,
       B2NN_PREC_AUTO = 2 /* B2NN_PRECISION env var ("fp32" | "tf32" | "tf32x3" | "bf16"); default tf32x3 = fp32-grade, so a
                             drop-in caller keeps the reference's fp32 tolerance without asking                        */,
       B2NN_PREC_TF32X3 = 3 /* fp32-grade on tensor cores: every operand carries its tf32 remainder and each K step
                               issues hi*hi + hi*lo + lo*hi (3 tcgen05.mma) into the fp32 TMEM accumulator            */,
       B2NN_PREC_BF16 = 4   /* tcgen05 kind::f16 on bf16 COPIES of the operands (8 mantissa bits, round to nearest even),
                               fp32 accumulate; weights, velocity, gradients, activations and deltas stay fp32 — only the
                               contractions read the bf16 copies.  OPT-IN, stated bound 6e-2 on the weights              */ };

/* Function-pointer shapes of the reference's pluggable device ops (cublasNN.h:33, 45-50).  All pointers are DEVICE
 * pointers; the functions are HOST functions that enqueue work on the legacy default stream. */
typedef void (*b2nn_act_fn)(const float* in, float* out, int count);
typedef void (*b2nn_out_fn)(const float* in, float* out, int rows, int cols);
typedef void (*b2nn_cost_fn)(float* h, float* y, float* J, int count);
typedef void (*b2nn_init_fn)(float* theta, int in, int out, int count);

/* "Iteration: <iter>\tCost: <J>" events (cublasNN.cpp:555, 616); the façade prints them. */
typedef void (*b2nn_log_cb)(void* user, int iter, float J);

typedef struct b2nn_train_desc {
    float momentum;            /* mu  (cublasNN.cpp:560, 621) */
    float rate;                /* alpha = -rate / mB (cublasNN.cpp:544, 601); mB = GLOBAL batch rows when data-parallel */
    int   iters;               /* epochs (setIters) */
    int   display;             /* print period in epochs (setDisplay); <= 0 means 1 (reference: UB, cublasNN.h:118) */
    int   batch_num;           /* NUMBER of batches, not batch size (cublasNN.cpp:401) */
    int   classify;            /* 1: trainClassifyMomentum, 0: trainFuncApproxMomentum */
    int   act_kind;            /* B2NN_ACT_*  */
    int   out_kind;            /* B2NN_OUT_*  (ignored when classify == 0) */
    int   cost_kind;           /* B2NN_COST_* (ignored when classify == 0) */
    int   precision;           /* B2NN_PREC_* */
    int   skip_final_cost;     /* 1: do not run calcFinalCost at the end (benchmarks time the loop only) */
    int   reserved0;
    b2nn_act_fn  act_fn;       /* used when act_kind  == B2NN_ACT_CUSTOM */
    b2nn_act_fn  act_grad_fn;
    b2nn_out_fn  out_fn;       /* used when out_kind  == B2NN_OUT_CUSTOM */
    b2nn_cost_fn cost_fn;      /* used when cost_kind == B2NN_COST_CUSTOM */
} b2nn_train_desc;

typedef struct b2nn_train_result {
    double call_seconds;       /* what train*Momentum returns: setup + loop + final cost (cublasNN.cpp:529-578) */
    double loop_seconds;       /* CUDA-event time of the epoch loop only */
    double final_cost;         /* "Final Cost" (cublasNN.cpp:745) */
    double error_count;        /* "Total errors" (cublasNN.cpp:760), classification only */
    int64_t steps;             /* iters * batch_num */
    int64_t kernel_launches;   /* kernels of this library launched inside the loop */
    int64_t gemm_tc_launches;  /* of those, tcgen05 GEMMs */
} b2nn_train_result;

/* ---- lifetime ---------------------------------------------------------------------------------------------- */
/* replaces cublasNN::cublasNN / ~cublasNN (cublasNN.cpp:12-60).  device < 0 => current device. */
int  b2nn_create(b2nn_ctx** out, int device);
void b2nn_destroy(b2nn_ctx* ctx);
const char* b2nn_last_error(void);
/* 1 when a usable sm_100 device is present.  Never throws; no side effects beyond CUDA runtime init.  (The reference
 * checks nothing: its constructor creates handles on whatever device 0 is, cublasNN.cpp:12-30.) */
int  b2nn_device_ok(void);
const char* b2nn_version(void);

/* ---- network construction ---------------------------------------------------------------------------------- */
/* replaces cublasNN::setLayers (cublasNN.cpp:270-328): one cuRAND XORWOW uniform call over the whole arena, then the
 * per-layer rescale.  seed_is_clock != 0 reproduces the reference's clock() seed (cublasNN.cpp:27). */
int b2nn_set_layers(b2nn_ctx* ctx, const int* layers, int n_layers, int init_kind, b2nn_init_fn custom_init,
                    uint64_t seed, int seed_is_clock);
int64_t b2nn_num_weights(const b2nn_ctx* ctx);
/* Not in the reference (weights are private, cublasNN.h:135): needed for parity tests and checkpoints. */
int b2nn_get_weights(b2nn_ctx* ctx, float* host_theta, int64_t count);
int b2nn_set_weights(b2nn_ctx* ctx, const float* host_theta, int64_t count);

/* Checkpoint file (not in the reference: its weights die with the process): layer widths, weights and velocity in the
 * reference arena layout, the normalisation statistics when present, and a checksum.  load (re)creates the layers. */
int b2nn_save_checkpoint(b2nn_ctx* ctx, const char* path);
int b2nn_load_checkpoint(b2nn_ctx* ctx, const char* path);

/* ---- data --------------------------------------------------------------------------------------------------- */
/* replaces cublasNN::copyDataGPU + splitData (cublasNN.cpp:348-358, 498-524).  x: m x n_plus_1 (ones column first),
 * y: m x k, both column-major HOST buffers; the engine keeps its own device copy. */
int b2nn_set_train_data(b2nn_ctx* ctx, const float* x, const float* y, int64_t m, int n_plus_1, int k);

/* Device-side data staging: replaces normaliseData + addBiasData + copyDataGPU (cublasNN.cpp:191-268, 330-358) in one
 * call.  x_raw: m x n column-major HOST matrix WITHOUT the ones column.  normalise != 0: the per-feature mean and (m-1)
 * standard deviation are computed on the device (the reference's two passes, cublasNN.cpp:214-268, accumulated in double),
 * applied in place ((x - mean) / stddev, 0 where the feature is constant, cublasNN.cpp:199-212) and kept in the context
 * for b2nn_evaluate_raw / b2nn_predict_raw; mean_out / std_out (optional, n floats) receive them. */
int b2nn_set_train_data_raw(b2nn_ctx* ctx, const float* x_raw, const float* y, int64_t m, int n, int k, int normalise,
                            float* mean_out, float* std_out);
/* Install / read the statistics directly (n = 0 clears them): the reference keeps them as host members filled by
 * mean / stddev (cublasNN.cpp:214-268) and reuses them for the validation and prediction sets (cublasNN.cpp:191-197). */
int b2nn_set_normalisation(b2nn_ctx* ctx, const float* mean, const float* stddev, int n);
int b2nn_get_normalisation(b2nn_ctx* ctx, float* mean, float* stddev, int n);

/* ---- the hot path ------------------------------------------------------------------------------------------- */
/* replaces trainFuncApproxMomentum / trainClassifyMomentum (cublasNN.cpp:526-640). */
int b2nn_train(b2nn_ctx* ctx, const b2nn_train_desc* desc, b2nn_log_cb log, void* log_user,
               b2nn_train_result* result);

/* One iteration of the reference's batch loop (cublasNN.cpp:597-629 / 540-568) from HOST buffers (the end-to-end
 * path: H2D of the batch, forward, backward, update, J back).
 * x: rows x n_plus_1, y: rows x k column-major (ld = rows).  Velocity persists across calls until
 * b2nn_reset_velocity.  global_rows: batch rows summed over all data-parallel ranks (0 => rows). */
int b2nn_train_step_host(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x, const float* y, int64_t rows,
                         int64_t global_rows, float* J_out);
/* velocity = 0, what every train* call of the reference starts with (cublasNN.cpp:533-534, 590-591). */
int b2nn_reset_velocity(b2nn_ctx* ctx);

/* Streaming trainer for a HOST-resident training set (it never has to fit on the device): same loop as b2nn_train,
 * every batch travels host -> device through two staging buffers on a copy stream, so the copy of batch b+1 overlaps
 * the compute of batch b.  x: m x n_plus_1 (ones column first), y: m x k, column-major HOST buffers (pinned memory
 * gives the full PCIe rate).  J_steps (optional, iters*batch_num floats): the cost of EVERY step, each read back with
 * its own asynchronous D2H.  No final cost pass (the data is not resident); velocity is zeroed at entry like
 * train*Momentum (cublasNN.cpp:533-534).  Data parallel: every rank streams its own shard (its rows of every batch);
 * J_steps then holds the mean cost of the rank's rows. */
int b2nn_train_host(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x, const float* y, int64_t m, int n_plus_1,
                    int k, float* J_steps, b2nn_train_result* result);

/* ---- evaluation / inference --------------------------------------------------------------------------------- */
/* replaces calcFinalCost / validate (cublasNN.cpp:687-881) on caller data.  y_is_labels: y is rows x 1 class ids
 * (validate, cublasNN.cpp:788-790) instead of rows x k one-hot. */
int b2nn_evaluate(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x, const float* y, int64_t rows,
                  int y_is_labels, double* cost_out, double* errors_out);
/* replaces predict (cublasNN.cpp:906-977).  out: rows x 1 class ids when classify, rows x k column-major otherwise.
 * probs (optional): rows x k column-major activations of the last layer.  Large batches travel in chunks through two
 * pinned staging buffers (H2D of chunk c+1, compute of chunk c and D2H of chunk c-1 overlap); a pinned x is DMA'd directly. */
int b2nn_predict(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x, int64_t rows, float* out, float* probs);
/* b2nn_evaluate / b2nn_predict on RAW rows (rows x n column-major, no ones column): normalise*Data + addBias* of the
 * reference happen on the device while the rows are staged (normalise != 0 applies the stored statistics). */
int b2nn_evaluate_raw(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x_raw, const float* y, int64_t rows,
                      int y_is_labels, int normalise, double* cost_out, double* errors_out);
int b2nn_predict_raw(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x_raw, int64_t rows, int normalise, float* out,
                     float* probs);
/* Same forward (forwardPropagate + the output activation, cublasNN.cpp:642-661, 937-960) with x ALREADY on the device
 * (config 5's data-resident inference sweep). */
int b2nn_predict_device(b2nn_ctx* ctx, const b2nn_train_desc* desc, const float* x_dev, int64_t rows, int64_t ld,
                        float* out_dev);

/* ---- built-in recognition ----------------------------------------------------------------------------------- */
/* The reference passes its ops as function pointers (cublasNN.h:42-69; activations.h, costfunctions.h,
 * randinitweights.h).  Maps the address of a built-in op (tanhGPU, softmaxGPU, ...) to its B2NN_* enum, or -1 for user code.
 * family: 0 hidden activation, 1 its derivative, 2 output activation, 3 cost, 4 init. */
int b2nn_builtin_kind(const void* fn, int family);

/* ---- data parallel (new; the reference is single-GPU) --------------------------------------------------------- */
/* One process per GPU.  Rank 0 calls b2nn_comm_unique_id, the host plumbing (torch.distributed, MPI, ...) broadcasts
 * the 128 bytes, every rank calls b2nn_comm_init.  Afterwards b2nn_train all-reduces the weight gradients with NCCL
 * over NVLink, overlapped with the backprop of earlier layers; each rank's train data is its slice of every batch. */
int b2nn_comm_unique_id(void* id128);
int b2nn_comm_init(b2nn_ctx* ctx, int rank, int nranks, const void* id128);
int b2nn_comm_barrier(b2nn_ctx* ctx);
/* Optional: exchange the gradients of the large layers over NVLink peer memory instead of NCCL.  The weight-gradient
 * GEMM stores each output tile directly into the arena of the rank that owns those rows (reduce-scatter fused into the
 * epilogue); the owner's update kernel sums the partial slices, applies the momentum step and stores the new rows into
 * every rank's weights (update fused with the all-gather).  Call after b2nn_comm_init and b2nn_set_layers, on every
 * rank: export 192 bytes (three CUDA IPC handles), all-gather them in rank order with the host's own plumbing, open. */
int b2nn_comm_p2p_export(b2nn_ctx* ctx, void* blob192);
int b2nn_comm_p2p_open(b2nn_ctx* ctx, const void* blobs_world_x_192);   /* NULL: back to the NCCL path */

/* ---- introspection ------------------------------------------------------------------------------------------ */
/* Raw GEMM entry (what the reference's cublasSgemm call sites compute, cublasNN.cpp:645-684) used by the kernel unit
 * tests and the roofline benchmark: D(m,n) = sum_k A(m,k) B(n,k), D stored
 * at D[n*ldd + m].  a_major / b_major: 0 = K contiguous (A[m*lda+k]), 1 = M/N contiguous (A[k*lda+m]).
 * epi: 0 store, 1 sigmoid, 2 tanh, 3 multiply by sigmoid'(aux), 4 multiply by tanh'(aux) with aux = activation values
 * laid out like D.  All pointers are device pointers.  backend: 0 fp32 FFMA, 1 tcgen05 tf32, 3 tcgen05 tf32 where
 * the caller guarantees 32 floats of readable slack after every row of an M/N-contiguous operand (one TMA box per
 * tile instead of one per 32-float atom); 5 / 7 = the same two with the fp32-grade 3xTF32 arithmetic (operand low
 * parts are split into temporaries inside the call); 9 / 11 = the same two on bf16 copies of the operands (kind::f16,
 * fp32 accumulation). */
int b2nn_gemm(b2nn_ctx* ctx, int backend, const float* A, int64_t lda, int a_major, const float* B, int64_t ldb,
              int b_major, float* D, int64_t ldd, int M, int N, int K, int epi, const float* aux, int64_t ldaux);
/* Momentum update kernel alone (the reference's two cublasSgeam per layer, cublasNN.cpp:558-566, 619-627):
 * v = mu*v + alpha*g ; theta = theta + v over count elements (device pointers). */
int b2nn_momentum_update(b2nn_ctx* ctx, float* theta, float* vel, const float* grad, int64_t count, float mu,
                         float alpha);
void* b2nn_stream(b2nn_ctx* ctx);   /* cudaStream_t the engine launches on */
/* Per-launch CUDA-event timing on the engine's stream (what bench.py's roofline leg reads).  Kinds: 0 forward
 * contraction, 1 delta backprop, 2 weight gradient, 3 output layer, 4 momentum update.  Arrays have
 * B2NN_PROFILE_KINDS entries; read() synchronises, sums the events recorded since the last read and clears them. */
#define B2NN_PROFILE_KINDS 5
int b2nn_profile_enable(b2nn_ctx* ctx, int on);
int b2nn_profile_read(b2nn_ctx* ctx, double* ms, double* flops, double* bytes, int64_t* launches,
                      int64_t* tc_launches);

#ifdef __cplusplus
}
#endif
#endif /* B2NN_H */
