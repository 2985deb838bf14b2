// oracle/mlp_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (fp32 and fp64) of the one hot path of HenryJia/CUBLAS-ML: the mini-batch
// forward / delta-backprop / momentum-update loop of class cublasNN.  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load this library, and only as
// the checker.  The product (cublas-ml_b200/csrc) never links or calls it.
//
// Parity pin: the reference ships no tests, golden vectors or CPU path.  This oracle
// is therefore pinned against outputs of the reference's own cuBLAS build (oracle/_ref, built from the
// unmodified sources under /root/reference by oracle/Makefile) run on a B200; those outputs are committed
// under tests/golden/ together with the script that made them (tests/golden/make_golden.py).
//
// All matrices are column-major with the SAMPLE index fastest, exactly as the reference keeps them on the
// host (IDX2C(i,j,ld) = j*ld+i, cublasNN.h:17).  X carries the leading ones column written by addBias
// (cublasNN.cpp:330-346).  Theta_j is (l_j+1) x l_{j+1}, row 0 = bias weights (cublasNN.cpp:293-301).
//
// Each function cites the reference lines it restates.  Quirks that are NOT reproduced:
// Q2 (element-wise work over mBatchMax rows, stale tail in printed J) — the oracle uses the exact batch.

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_WITH_CURAND
#include <curand.h>
#endif

extern "C" {

// Mirrors the arguments of trainClassifyMomentum / trainFuncApproxMomentum (cublasNN.h:45-50) plus the
// state the class holds when they are called (layers, iters, display, x, y, m).
struct oracle_problem {
    const int*   layers;      // l_0 .. l_{L-1}, bias units NOT included (cublasNN.cpp:281-282)
    int          n_layers;    // L (layerNum), >= 3 (cublasNN.cpp:657 needs one hidden layer)
    const float* x;           // m x (l_0+1), column-major, column 0 == 1.0f
    const float* y;           // m x l_{L-1}, column-major (one-hot for classification, cublasNN.cpp:135-146)
    int64_t      m;           // rows
    int          batch_num;   // number of batches (NOT batch size), cublasNN.cpp:401
    int          iters;       // epochs (cublasNN.h:34)
    int          display;     // print period in epochs (cublasNN.h:35)
    float        momentum;
    float        rate;
    int          hidden_act;  // 0 = sigmoidGPU/sigmoidGradGPU, 1 = tanhGPU/sechSqGPU
    int          classify;    // 1 = trainClassifyMomentum, 0 = trainFuncApproxMomentum
    int          out_act;     // classify only: 0 = sigmoidOutputGPU, 1 = softmaxGPU
    int          cost;        // classify only: 0 = negLnMaxCostGPU, 1 = crossEntropyCostGPU
};

struct oracle_result {
    double  final_cost;       // "Final Cost" of calcFinalCost (cublasNN.cpp:725-745)
    double  error_count;      // "Total errors" (cublasNN.cpp:756-760), classify only
    int     n_logged;         // number of (iter, J) pairs written
    double  loop_seconds;     // wall time of the epoch loop only
};

}  // extern "C"

namespace {

inline int64_t cm(int64_t i, int64_t j, int64_t ld) { return j * ld + i; }  // cublasNN.h:17

// ---- element-wise math, one template per reference kernel ------------------------------------------

template <class T> inline T sigmoid_fn(T v) { return T(1) / (T(1) + std::exp(-v)); }     // activations.cu:34-39
template <class T> inline T sigmoid_grad_fn(T v) { T a = sigmoid_fn(v); return a * (T(1) - a); }  // :41-49
template <class T> inline T tanh_fn(T v) { return std::tanh(v); }                          // activations.cu:51-56
template <class T> inline T sech_sq_fn(T v) { T c = std::cosh(v); return T(1) / (c * c); }  // activations.cu:58-66

template <class T>
inline T hidden_fn(int kind, T v) { return kind == 0 ? sigmoid_fn(v) : tanh_fn(v); }
template <class T>
inline T hidden_grad_fn(int kind, T v) { return kind == 0 ? sigmoid_grad_fn(v) : sech_sq_fn(v); }

// ---- three GEMM shapes of matMatMultiplyGPU (cublasNN.h:100-111), alpha = 1, beta = 0 --------------

// C(m x n, ldc) = A(m x k, lda) * B(k x n, ldb)                 [cublasSgemm N,N  cublasNN.cpp:645,653,658]
template <class T>
void gemm_nn(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, int64_t m, int64_t n, int64_t k) {
    const int64_t MB = 512;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i0 = 0; i0 < m; i0 += MB) {
            const int64_t i1 = std::min(m, i0 + MB);
            T* c = C + cm(0, j, ldc);
            for (int64_t i = i0; i < i1; ++i) c[i] = T(0);
            for (int64_t p = 0; p < k; ++p) {
                const T b = B[cm(p, j, ldb)];
                const T* a = A + cm(0, p, lda);
#pragma omp simd
                for (int64_t i = i0; i < i1; ++i) c[i] += a[i] * b;
            }
        }
}

// C(m x n, ldc) = A(m x k, lda) * B(n x k, ldb)^T               [cublasSgemm N,T  cublasNN.cpp:674-675]
template <class T>
void gemm_nt(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, int64_t m, int64_t n, int64_t k) {
    const int64_t MB = 512;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i0 = 0; i0 < m; i0 += MB) {
            const int64_t i1 = std::min(m, i0 + MB);
            T* c = C + cm(0, j, ldc);
            for (int64_t i = i0; i < i1; ++i) c[i] = T(0);
            for (int64_t p = 0; p < k; ++p) {
                const T b = B[cm(j, p, ldb)];
                const T* a = A + cm(0, p, lda);
#pragma omp simd
                for (int64_t i = i0; i < i1; ++i) c[i] += a[i] * b;
            }
        }
}

// C(m x n, ldc) = A(k x m, lda)^T * B(k x n, ldb)               [cublasSgemm T,N  cublasNN.cpp:680-684]
template <class T>
void gemm_tn(const T* A, int64_t lda, const T* B, int64_t ldb, T* C, int64_t ldc, int64_t m, int64_t n, int64_t k) {
    // Tile so a block of A columns and B columns stays cache-resident while k streams.
    const int64_t TI = 8, TJ = 8;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t j0 = 0; j0 < n; j0 += TJ)
        for (int64_t i0 = 0; i0 < m; i0 += TI) {
            const int64_t j1 = std::min(n, j0 + TJ), i1 = std::min(m, i0 + TI);
            for (int64_t j = j0; j < j1; ++j)
                for (int64_t i = i0; i < i1; ++i) {
                    const T* a = A + cm(0, i, lda);
                    const T* b = B + cm(0, j, ldb);
                    T s = T(0);
#pragma omp simd reduction(+ : s)
                    for (int64_t p = 0; p < k; ++p) s += a[p] * b[p];
                    C[cm(i, j, ldc)] = s;
                }
        }
}

// ---- the network state the reference keeps in its device arenas (cublasNN.h:135-182) ---------------

template <class T>
struct Net {
    int L = 0;
    std::vector<int> l;                    // layers
    std::vector<int64_t> theta_pos;        // thetaPos  (cublasNN.cpp:293-301)
    std::vector<int64_t> theta_size;       // thetaSize
    int64_t total_theta = 0;
    std::vector<T> theta;                  // thetaBaseGPU contents

    void shape(const int* layers, int n_layers) {
        L = n_layers;
        l.assign(layers, layers + n_layers);
        theta_pos.assign(L - 1, 0);
        theta_size.assign(L - 1, 0);
        total_theta = 0;
        for (int j = 0; j < L - 1; ++j) {
            theta_pos[j] = total_theta;
            theta_size[j] = int64_t(l[j] + 1) * l[j + 1];
            total_theta += theta_size[j];
        }
        theta.assign(total_theta, T(0));
    }
};

// Work buffers for one forward/backward over `rows` samples (allocVarGPU, cublasNN.cpp:398-496).
template <class T>
struct Work {
    int64_t rows = 0;
    std::vector<std::vector<T>> z;      // z_j : rows x l_{j+1}
    std::vector<std::vector<T>> a;      // a_j : rows x (l_{j+1}+1) for hidden j, rows x l_{L-1} for the last
    std::vector<std::vector<T>> delta;  // delta_j : rows x l_{j+1}
    std::vector<std::vector<T>> Delta;  // Delta_j : l_{j+1} x (l_j+1)  (transposed vs theta_j, cublasNN.cpp:680-684)
    std::vector<T> product, sig_grad;

    void alloc(const Net<T>& net, int64_t r, bool need_backward) {
        rows = r;
        const int L = net.L;
        z.resize(L - 1); a.resize(L - 1);
        int64_t zmax = 0;
        for (int j = 0; j < L - 1; ++j) {
            z[j].assign(size_t(r) * net.l[j + 1], T(0));
            const int64_t acols = (j < L - 2) ? net.l[j + 1] + 1 : net.l[j + 1];
            a[j].assign(size_t(r) * acols, T(0));
            zmax = std::max<int64_t>(zmax, r * net.l[j + 1]);
        }
        if (need_backward) {
            delta.resize(L - 1); Delta.resize(L - 1);
            for (int j = 0; j < L - 1; ++j) {
                delta[j].assign(size_t(r) * net.l[j + 1], T(0));
                Delta[j].assign(size_t(net.theta_size[j]), T(0));
            }
            product.assign(size_t(zmax + r), T(0));
            sig_grad.assign(size_t(zmax), T(0));
        }
    }
};

// forwardPropagate (cublasNN.cpp:642-661): leaves z_{L-2} un-activated; a_j = [1 | f(z_j)].
// X is rows x (l_0+1) with leading dimension ldx.
template <class T>
void forward(const Net<T>& net, Work<T>& w, const T* X, int64_t ldx, int64_t rows, int hidden_act) {
    const int L = net.L;
    gemm_nn(X, ldx, net.theta.data() + net.theta_pos[0], int64_t(net.l[0] + 1), w.z[0].data(), rows,
            rows, int64_t(net.l[1]), int64_t(net.l[0] + 1));                                   // :645-646
    for (int j = 0; j < L - 2; ++j) {
        T* aj = w.a[j].data();
        const T* zj = w.z[j].data();
        const int64_t cnt = rows * net.l[j + 1];
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) aj[rows + i] = hidden_fn(hidden_act, zj[i]);        // :648, :655
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < rows; ++i) aj[i] = T(1);                                       // addBiasMatGPU :652,:657
        gemm_nn(aj, rows, net.theta.data() + net.theta_pos[j + 1], int64_t(net.l[j + 1] + 1),
                w.z[j + 1].data(), rows, rows, int64_t(net.l[j + 2]), int64_t(net.l[j + 1] + 1));  // :653-654, :658-659
    }
}

// Output activation (activations.cu:9-12 sigmoidOutputGPU, :68-88 softmax WITHOUT max-shift, exp twice).
template <class T>
void output_activation(const T* z, T* h, int64_t rows, int K, int out_act) {
    if (out_act == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < rows * K; ++i) h[i] = sigmoid_fn(z[i]);
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < rows; ++i) {
            T total = T(0);
            for (int j = 0; j < K; ++j) total += std::exp(z[cm(i, j, rows)]);
            for (int j = 0; j < K; ++j) h[cm(i, j, rows)] = std::exp(z[cm(i, j, rows)]) / total;
        }
    }
}

// backwardPropagate (cublasNN.cpp:663-685).  `output` is h (classification) or z_{L-2} (regression).
// Y is rows x l_{L-1}, X is rows x (l_0+1).
template <class T>
void backward(const Net<T>& net, Work<T>& w, const T* output, const T* X, int64_t ldx, const T* Y, int64_t ldy,
              int64_t rows, int hidden_act) {
    const int L = net.L;
    const int K = net.l[L - 1];
    T* dl = w.delta[L - 2].data();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < rows; ++i)
        for (int j = 0; j < K; ++j) dl[cm(i, j, rows)] = output[cm(i, j, rows)] - Y[cm(i, j, ldy)];   // :666-667

    for (int j = L - 3; j >= 0; --j) {                                                           // :670-677
        const int64_t cnt = rows * net.l[j + 1];
        const T* zj = w.z[j].data();
        T* sg = w.sig_grad.data();
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) sg[i] = hidden_grad_fn(hidden_act, zj[i]);            // :673
        gemm_nt(w.delta[j + 1].data(), rows, net.theta.data() + net.theta_pos[j + 1], int64_t(net.l[j + 1] + 1),
                w.product.data(), rows, rows, int64_t(net.l[j + 1] + 1), int64_t(net.l[j + 2]));  // :674-675
        const T* pr = w.product.data() + rows;   // skip the bias column (:676 "product + mBatch[b]")
        T* dj = w.delta[j].data();
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < cnt; ++i) dj[i] = sg[i] * pr[i];                                 // kernels.cu:53-58
    }

    // Delta_j = delta_j^T * [1, a_{j-1}]  (l_{j+1} x (l_j+1), ld = l_{j+1})                    :680-684
    gemm_tn(w.delta[0].data(), rows, X, ldx, w.Delta[0].data(), int64_t(net.l[1]),
            int64_t(net.l[1]), int64_t(net.l[0] + 1), rows);
    for (int j = 1; j < L - 1; ++j)
        gemm_tn(w.delta[j].data(), rows, w.a[j - 1].data(), rows, w.Delta[j].data(), int64_t(net.l[j + 1]),
                int64_t(net.l[j + 1]), int64_t(net.l[j] + 1), rows);
}

// Per-element costs (costfunctions.cu:14-26) summed with |.| like cublasSasum (cublasNN.cpp:614, 734).
template <class T>
double classify_cost(const T* h, const T* Y, int64_t ldy, int64_t rows, int K, int cost) {
    double s = 0.0;
    // Accumulate in T-typed partials per column, then in double across columns: the reference's Sasum
    // order is unspecified, so this only has to be a faithful sum.
    for (int j = 0; j < K; ++j) {
        T cs = T(0);
        for (int64_t i = 0; i < rows; ++i) {
            const T hv = h[cm(i, j, rows)], yv = Y[cm(i, j, ldy)];
            T c;
            if (cost == 0) c = -yv * std::log(hv) - (T(1) - yv) * std::log(T(1) - hv);
            else           c = -yv * std::log(hv);
            cs += std::fabs(c);
        }
        s += double(cs);
    }
    return s;
}

// Momentum update, two cublasSgeam per layer (cublasNN.cpp:558-566, 619-627):
//   v <- mu*v + alpha*Delta^T ; theta <- theta + v      with alpha = -rate/mB (:544, :601)
template <class T>
void momentum_update(Net<T>& net, const Work<T>& w, std::vector<T>& vel, T mu, T alpha) {
    for (int j = 0; j < net.L - 1; ++j) {
        const int64_t rin = net.l[j] + 1, cout = net.l[j + 1];
        T* th = net.theta.data() + net.theta_pos[j];
        T* v = vel.data() + net.theta_pos[j];
        const T* D = w.Delta[j].data();
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < cout; ++c)
            for (int64_t r = 0; r < rin; ++r) {
                const int64_t t = cm(r, c, rin);
                v[t] = mu * v[t] + alpha * D[cm(c, r, cout)];
                th[t] = th[t] + v[t];
            }
    }
}

template <class T>
int argmax_first(const T* p, int64_t i, int64_t ld, int K) {
    // kernelProbToNum (kernels.cu:74-86): max seeded with 0, strict '>' => first maximum wins.
    T best = T(0); int idx = 0;
    for (int j = 0; j < K; ++j) if (p[cm(i, j, ld)] > best) { best = p[cm(i, j, ld)]; idx = j; }
    return idx;
}

template <class T>
int train_impl(const oracle_problem* p, const float* theta0, float* theta_out, int* log_iter, double* log_J,
               int log_cap, oracle_result* res) {
    if (!p || p->n_layers < 3 || p->batch_num < 1 || p->m < p->batch_num) return 1;
    Net<T> net;
    net.shape(p->layers, p->n_layers);
    for (int64_t i = 0; i < net.total_theta; ++i) net.theta[i] = T(theta0[i]);
    const int L = net.L, K = net.l[L - 1];
    const int64_t m = p->m, ncol = net.l[0] + 1;

    std::vector<T> X(size_t(m) * ncol), Y(size_t(m) * K);
    for (size_t i = 0; i < X.size(); ++i) X[i] = T(p->x[i]);
    for (size_t i = 0; i < Y.size(); ++i) Y[i] = T(p->y[i]);

    // Batch table (allocVarGPU, cublasNN.cpp:401, 421-432): contiguous row ranges, last absorbs the rest.
    const int nb = p->batch_num;
    const int64_t bs = m / nb;
    std::vector<int64_t> b_start(nb), b_rows(nb);
    for (int b = 0; b < nb; ++b) { b_start[b] = b * bs; b_rows[b] = (b == nb - 1) ? m - b * bs : bs; }
    const int64_t rows_max = b_rows[nb - 1];

    Work<T> w;
    w.alloc(net, rows_max, true);
    std::vector<T> vel(size_t(net.total_theta), T(0));                                      // :533-534, :590-591
    std::vector<T> h(size_t(rows_max) * K);
    const int display = p->display > 0 ? p->display : 1;

    int n_logged = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < p->iters; ++it) {                                                   // :540-568, :597-629
        for (int b = 0; b < nb; ++b) {
            const int64_t rows = b_rows[b];
            // splitData (cublasNN.cpp:498-524) only re-lays each batch as a contiguous col-major block; a
            // (pointer, ld = m) view of the same rows is the identical matrix.
            const T* Xb = X.data() + b_start[b];
            const T* Yb = Y.data() + b_start[b];
            // Work buffers are sized for rows_max but used with ld = rows, like the reference (ld = mBatch[b]).
            forward(net, w, Xb, m, rows, p->hidden_act);
            const T* out;
            if (p->classify) { output_activation(w.z[L - 2].data(), h.data(), rows, K, p->out_act); out = h.data(); }  // :606
            else out = w.z[L - 2].data();                                                      // :548
            backward(net, w, out, Xb, m, Yb, m, rows, p->hidden_act);
            if ((it + 1) % display == 0 && b == 0) {                                           // :550-556, :610-617
                double J;
                if (p->classify) J = classify_cost(h.data(), Yb, m, rows, K, p->cost) / double(rows);
                else {
                    double s = 0.0; const T* d = w.delta[L - 2].data();
                    for (int64_t i = 0; i < rows * K; ++i) s += double(d[i]) * double(d[i]);
                    J = s / (2.0 * double(rows));
                }
                if (n_logged < log_cap) { log_iter[n_logged] = it; log_J[n_logged] = J; }
                ++n_logged;
            }
            const T alpha = -T(p->rate) / T(rows);                                             // :544, :601
            momentum_update(net, w, vel, T(p->momentum), alpha);
        }
    }
    auto t1 = std::chrono::steady_clock::now();

    // calcFinalCost (cublasNN.cpp:687-777): one forward over all m rows with the final theta.
    Work<T> wf;
    wf.alloc(net, m, false);
    forward(net, wf, X.data(), m, m, p->hidden_act);
    double Jf, errs = 0.0;
    if (p->classify) {
        std::vector<T> hf(size_t(m) * K);
        output_activation(wf.z[L - 2].data(), hf.data(), m, K, p->out_act);
        Jf = classify_cost(hf.data(), Y.data(), m, m, K, p->cost) / double(m);
        for (int64_t i = 0; i < m; ++i)
            if (argmax_first(hf.data(), i, m, K) != argmax_first(Y.data(), i, m, K)) errs += 1.0;   // :756-759
    } else {
        double s = 0.0; const T* zl = wf.z[L - 2].data();
        for (int64_t i = 0; i < m * K; ++i) { double d = double(zl[i]) - double(Y[i]); s += d * d; }  // :739-741
        Jf = s / (2.0 * double(m));
    }
    if (theta_out) for (int64_t i = 0; i < net.total_theta; ++i) theta_out[i] = float(net.theta[i]);
    if (res) {
        res->final_cost = Jf; res->error_count = errs; res->n_logged = std::min(n_logged, log_cap);
        res->loop_seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    return 0;
}

// predict (cublasNN.cpp:906-977): forward only; classification returns the argmax class as a float
// (rows x 1), regression returns z_{L-2} (rows x l_{L-1}).  If probs != nullptr it also receives h / z.
template <class T>
int predict_impl(const int* layers, int n_layers, const float* theta, const float* x, int64_t rows, int hidden_act,
                 int classify, int out_act, float* out, float* probs) {
    if (n_layers < 3 || rows < 1) return 1;
    Net<T> net; net.shape(layers, n_layers);
    for (int64_t i = 0; i < net.total_theta; ++i) net.theta[i] = T(theta[i]);
    const int L = net.L, K = net.l[L - 1];
    std::vector<T> X(size_t(rows) * (net.l[0] + 1));
    for (size_t i = 0; i < X.size(); ++i) X[i] = T(x[i]);
    Work<T> w; w.alloc(net, rows, false);
    forward(net, w, X.data(), rows, rows, hidden_act);
    if (classify) {
        std::vector<T> h(size_t(rows) * K);
        output_activation(w.z[L - 2].data(), h.data(), rows, K, out_act);
        for (int64_t i = 0; i < rows; ++i) out[i] = float(argmax_first(h.data(), i, rows, K));
        if (probs) for (size_t i = 0; i < h.size(); ++i) probs[i] = float(h[i]);
    } else {
        for (int64_t i = 0; i < rows * K; ++i) out[i] = float(w.z[L - 2][i]);
        if (probs) for (int64_t i = 0; i < rows * K; ++i) probs[i] = float(w.z[L - 2][i]);
    }
    return 0;
}

}  // namespace

extern "C" {

int oracle_train_f32(const oracle_problem* p, const float* theta0, float* theta_out, int* log_iter, double* log_J,
                     int log_cap, oracle_result* res) {
    return train_impl<float>(p, theta0, theta_out, log_iter, log_J, log_cap, res);
}
int oracle_train_f64(const oracle_problem* p, const float* theta0, float* theta_out, int* log_iter, double* log_J,
                     int log_cap, oracle_result* res) {
    return train_impl<double>(p, theta0, theta_out, log_iter, log_J, log_cap, res);
}
int oracle_predict_f32(const int* layers, int n_layers, const float* theta, const float* x, int64_t rows,
                       int hidden_act, int classify, int out_act, float* out, float* probs) {
    return predict_impl<float>(layers, n_layers, theta, x, rows, hidden_act, classify, out_act, out, probs);
}
int oracle_predict_f64(const int* layers, int n_layers, const float* theta, const float* x, int64_t rows,
                       int hidden_act, int classify, int out_act, float* out, float* probs) {
    return predict_impl<double>(layers, n_layers, theta, x, rows, hidden_act, classify, out_act, out, probs);
}

// Total number of weights and the per-layer offsets (cublasNN.cpp:293-301).
int64_t oracle_theta_size(const int* layers, int n_layers) {
    int64_t t = 0;
    for (int j = 0; j + 1 < n_layers; ++j) t += int64_t(layers[j] + 1) * layers[j + 1];
    return t;
}

// Scale u in (0,1] to [-eps, +eps]: theta = u*2*eps - eps  (randinitweights.cu:13-31).
// init_kind 1: eps = 1/sqrt(in); 2: eps = sqrt(6)/sqrt(in+out); `in` INCLUDES the bias unit (cublasNN.cpp:317).
// nvcc contracts u*2*eps - eps into one FMA (checked in the reference's SASS: FFMA), so fmaf is used here.
void oracle_scale_uniform(float* theta, const int* layers, int n_layers, int init_kind) {
    int64_t pos = 0;
    for (int j = 0; j + 1 < n_layers; ++j) {
        const int in = layers[j] + 1, out = layers[j + 1];
        const float eps = (init_kind == 1) ? (1.0f / sqrtf(float(in))) : (sqrtf(6.0f) / sqrtf(float(in + out)));
        const int64_t cnt = int64_t(in) * out;
        for (int64_t i = 0; i < cnt; ++i) theta[pos + i] = fmaf(theta[pos + i] * 2.0f, eps, -eps);
        pos += cnt;
    }
}

// setLayers (cublasNN.cpp:270-328): curandGenerateUniform (XORWOW, default ordering) over the WHOLE arena in
// one call, then the per-layer rescale.  cuRAND is a closed-source dependency of the reference (Makefile:9,
// version unpinned; here 10.3.10); its host generator yields the device generator's sequence, so this is the
// same library the reference calls, not a re-derivation.  Returns non-zero when built without cuRAND.
int oracle_init_weights(float* theta, const int* layers, int n_layers, int init_kind, unsigned long long seed) {
#ifdef ORACLE_WITH_CURAND
    curandGenerator_t g;
    if (curandCreateGeneratorHost(&g, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) return 2;
    curandSetPseudoRandomGeneratorSeed(g, seed);
    curandSetGeneratorOffset(g, 0ULL);
    const int64_t n = oracle_theta_size(layers, n_layers);
    int rc = curandGenerateUniform(g, theta, size_t(n)) == CURAND_STATUS_SUCCESS ? 0 : 3;
    curandDestroyGenerator(g);
    if (rc == 0) oracle_scale_uniform(theta, layers, n_layers, init_kind);
    return rc;
#else
    (void)theta; (void)layers; (void)n_layers; (void)init_kind; (void)seed;
    return 1;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
