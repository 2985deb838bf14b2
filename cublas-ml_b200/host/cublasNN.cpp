// cublasNN.cpp — host-side mirror of the reference's class cublasNN over the C ABI of include/b2nn.h.
//
// Host work kept here (one-shot, not on the hot path): CSV I/O, vector -> column-major conversion, one-hot encoding.
// normalise*Data and addBias* only record what was asked for: the z-score statistics, their application and the ones
// row happen on the device while the rows are staged (b2nn_set_train_data_raw / b2nn_evaluate_raw / b2nn_predict_raw;
// reference: host loops, cublasNN.cpp:191-268, 330-346).  Everything that touches the GPU goes through b2nn_* calls;
// this file does not include a CUDA header.  Printed text follows the reference byte for byte where the reference prints
// (cublasNN.cpp:555, 616, 745, 760, 849, 861-862); failures print the reference's messages and return early.
#include "cublasNN.h"
#include "b2nn.h"
#include "activations.h"
#include "costfunctions.h"
#include "randinitweights.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace {

// A host matrix in the reference's layout: column-major, rows = samples (vector2dToMat, cublasNN.cpp:116-133).
struct HostMat {
    std::vector<float> v;
    long long rows = 0;
    int cols = 0;
    bool empty() const { return v.empty(); }
    float& at(long long i, int j) { return v[size_t(j) * size_t(rows) + size_t(i)]; }
    float at(long long i, int j) const { return v[size_t(j) * size_t(rows) + size_t(i)]; }
};

HostMat to_colmajor(const std::vector<std::vector<float>>& data) {
    HostMat m;
    if (data.empty() || data[0].empty()) return m;
    m.rows = (long long)data.size();
    m.cols = int(data[0].size());
    m.v.resize(size_t(m.rows) * size_t(m.cols));
    for (long long i = 0; i < m.rows; ++i) {
        const std::vector<float>& row = data[size_t(i)];
        const int c = int(std::min<size_t>(row.size(), size_t(m.cols)));
        for (int j = 0; j < c; ++j) m.at(i, j) = row[size_t(j)];
    }
    return m;
}

// classToBin (cublasNN.cpp:135-146) with integer indexing (Q10).
HostMat one_hot(const HostMat& labels, int K) {
    HostMat y;
    y.rows = labels.rows; y.cols = K;
    y.v.assign(size_t(y.rows) * size_t(K), 0.f);
    for (long long i = 0; i < labels.rows; ++i) {
        const long long c = (long long)labels.at(i, 0);
        if (c >= 0 && c < K) y.at(i, int(c)) = 1.f;
    }
    return y;
}

void print_log(void*, int iter, float J) { std::cout << "Iteration: " << iter << "\t" << "Cost: " << J << std::endl; }

}  // namespace

struct cublasNN::Impl {
    b2nn_ctx* ctx = nullptr;
    bool ctx_failed = false;
    std::vector<int> layers;
    int iters = 100;             // cublasNN.cpp:21
    int display = 1;             // uninitialised in the reference (Q5)
    float lambda = 0.f;
    int precision = B2NN_PREC_AUTO;
    bool seed_set = false;
    unsigned long long seed = 0;

    HostMat x, y, xValidate, yValidate, xPredict;
    bool classify_data = false;
    int n = 0;                   // features without bias (cublasNN.cpp:151)
    bool x_has_bias = false, xv_has_bias = false, xp_has_bias = false;      // addBias* was called (the ones row is the engine's)
    bool x_norm = false, xv_norm = false, xp_norm = false;                  // normalise*Data was called
    b2nn_train_result last{};

    // No usable sm_100 device: there is no CPU path and nothing after this point can work.  The reference would print
    // "cudaMalloc Failed" and carry on with null pointers (cublasNN.cpp:309, 366); here the failure is an exception, so an
    // unmodified main.cpp stops with a message and a non-zero status instead of reporting results it never computed.
    bool ensure_ctx() {
        if (ctx) return true;
        if (ctx_failed || b2nn_create(&ctx, -1) != B2NN_OK) {
            ctx_failed = true;
            const std::string why = std::string("b2nn: ") + b2nn_last_error();
            std::cerr << why << std::endl;
            throw std::runtime_error(why);
        }
        return true;
    }

    void fill_desc(b2nn_train_desc& d, bool classify, HiddenFn act, HiddenFn grad, OutputFn out, CostFn cost) const {
        std::memset(&d, 0, sizeof(d));
        d.iters = iters; d.display = display; d.batch_num = 1; d.classify = classify ? 1 : 0; d.precision = precision;
        const int ka = b2nn_builtin_kind((const void*)act, 0);
        const int kg = grad ? b2nn_builtin_kind((const void*)grad, 1) : ka;
        if (ka >= 0 && ka == kg) d.act_kind = ka;
        else { d.act_kind = B2NN_ACT_CUSTOM; d.act_fn = act; d.act_grad_fn = grad; }
        if (classify) {
            const int ko = b2nn_builtin_kind((const void*)out, 2);
            if (ko >= 0) d.out_kind = ko; else { d.out_kind = B2NN_OUT_CUSTOM; d.out_fn = out; }
            const int kc = cost ? b2nn_builtin_kind((const void*)cost, 3) : B2NN_COST_CROSSENTROPY;
            if (kc >= 0) d.cost_kind = kc; else { d.cost_kind = B2NN_COST_CUSTOM; d.cost_fn = cost; }
        } else {
            d.out_kind = B2NN_OUT_LINEAR; d.cost_kind = B2NN_COST_SQUARED;
        }
    }

    double train(bool classify, float momentum, float rate, HiddenFn act, HiddenFn grad, OutputFn out, CostFn cost,
                 int batchNum) {
        auto t0 = std::chrono::steady_clock::now();
        if (!ctx) { std::cout << "copyDataGPU must be called before training" << std::endl; return 0.0; }
        b2nn_train_desc d;
        fill_desc(d, classify, act, grad, out, cost);
        d.momentum = momentum; d.rate = rate; d.batch_num = batchNum;
        std::memset(&last, 0, sizeof(last));
        if (b2nn_train(ctx, &d, print_log, nullptr, &last) != B2NN_OK) {
            std::cout << "b2nn_train failed: " << b2nn_last_error() << std::endl;
            return 0.0;
        }
        std::cout << "Final Cost: " << float(last.final_cost) << std::endl;                         // cublasNN.cpp:745
        if (classify) {
            const float errs = float(last.error_count);
            std::cout << "Total errors " << '\t' << errs << std::endl
                      << "Error frequency" << '\t' << (errs / float(x.rows)) << std::endl;          // cublasNN.cpp:760
        }
        auto t1 = std::chrono::steady_clock::now();
        return std::chrono::duration<double>(t1 - t0).count();
    }
};

cublasNN::cublasNN() : impl(new Impl) {}

cublasNN::~cublasNN() {
    if (impl->ctx) b2nn_destroy(impl->ctx);
    delete impl;
}

// One pass over the file with strtold (the reference calls stold per field on a shrinking copy of the line:
// cublasNN.cpp:76-87, quadratic per line).  Same result: a vector of rows of floats; `header` skips the first line.
vector<vector<float>> cublasNN::readCSV(string fileName, bool header, float& time) {
    auto start = std::chrono::steady_clock::now();
    vector<vector<float>> result;
    std::ifstream in(fileName, std::ios::binary);
    if (!in.is_open()) {
        std::cerr << "failed to open file\n";                                                       // cublasNN.cpp:71-72
    } else {
        std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        const char* p = text.c_str();
        const char* end = p + text.size();
        if (header) { while (p < end && *p != '\n') ++p; if (p < end) ++p; }
        size_t width = 0;
        while (p < end) {
            const char* eol = static_cast<const char*>(std::memchr(p, '\n', size_t(end - p)));
            if (!eol) eol = end;
            if (eol > p && !(eol - p == 1 && *p == '\r')) {
                vector<float> row;
                row.reserve(width);
                const char* q = p;
                while (q < eol) {
                    char* next = nullptr;
                    const long double v = std::strtold(q, &next);
                    if (next == q) break;
                    row.push_back(float(v));
                    q = next;
                    while (q < eol && *q != ',') ++q;
                    if (q < eol) ++q;
                }
                if (!row.empty()) { width = row.size(); result.push_back(std::move(row)); }
            }
            p = eol < end ? eol + 1 : end;
        }
    }
    auto stop = std::chrono::steady_clock::now();
    time = std::chrono::duration<float>(stop - start).count();
    return result;
}

// Same text format as the reference (default ostream float formatting, ',' separators, no trailing newline).  The
// reference writes only the LAST element of the last row (cublasNN.cpp:108); here the whole last row is written —
// identical for the single-column results both examples produce.
double cublasNN::writeCSV(vector<vector<float>> data, string fileName) {
    auto start = std::chrono::steady_clock::now();
    std::ofstream out(fileName);
    for (size_t i = 0; i < data.size(); ++i) {
        for (size_t j = 0; j < data[i].size(); ++j) {
            if (j) out << ',';
            out << data[i][j];
        }
        if (i + 1 < data.size()) out << '\n';
    }
    out.flush();
    auto stop = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(stop - start).count();
}

void cublasNN::setData(vector<vector<float>> xVec, vector<vector<float>> yVec, bool classify) {
    if (xVec.empty() || yVec.empty()) { std::cout << "setData: empty input" << std::endl; return; }
    impl->x = to_colmajor(xVec);
    impl->n = impl->x.cols;
    impl->x_has_bias = false; impl->x_norm = false;
    impl->classify_data = classify;
    HostMat yRaw = to_colmajor(yVec);
    if (classify) {
        if (impl->layers.empty()) { std::cout << "setLayers must be called before setData for classification" << std::endl; return; }
        impl->y = one_hot(yRaw, impl->layers.back());                                               // cublasNN.cpp:162-164
    } else {
        impl->y = std::move(yRaw);
    }
}

void cublasNN::setValidateData(vector<vector<float>> xVec, vector<vector<float>> yVec, bool) {
    impl->xValidate = to_colmajor(xVec);
    impl->yValidate = to_colmajor(yVec);          // kept raw: labels for classification (cublasNN.cpp:177-178)
    impl->xv_has_bias = false; impl->xv_norm = false;
}

void cublasNN::setPredictData(vector<vector<float>> xVec) {
    impl->xPredict = to_colmajor(xVec);
    impl->xp_has_bias = false; impl->xp_norm = false;
}

void cublasNN::setLayers(int* l, int lNum, InitFn randInitWeights) {
    if (!l || lNum < 3) { std::cout << "setLayers: need at least one hidden layer" << std::endl; return; }
    impl->layers.assign(l, l + lNum);
    if (!impl->ensure_ctx()) return;
    int kind = b2nn_builtin_kind((const void*)randInitWeights, 4);
    b2nn_init_fn custom = nullptr;
    if (kind < 0) { kind = B2NN_INIT_CUSTOM; custom = randInitWeights; }
    if (b2nn_set_layers(impl->ctx, l, lNum, kind, custom, impl->seed, impl->seed_set ? 0 : 1) != B2NN_OK) {
        const char* e = b2nn_last_error();
        std::cout << (std::strcmp(e, "cudaMalloc Failed") == 0 ? "cudaMalloc Failed" : e) << std::endl;   // cublasNN.cpp:309
    }
}

void cublasNN::setIters(int i) { impl->iters = i; }
void cublasNN::setDisplay(int i) { impl->display = i; }
void cublasNN::setLambda(int l) { impl->lambda = float(l); }

// normaliseData (cublasNN.cpp:191-268): mean / (m-1) standard deviation of the training features, then (x - mean) / stddev.
// Recorded here, done on the device in copyDataGPU (statistics) and while validate / predict rows are staged.
void cublasNN::normaliseData() { if (!impl->x.empty()) impl->x_norm = true; }
void cublasNN::normaliseValidateData() { if (!impl->xValidate.empty()) impl->xv_norm = true; }
void cublasNN::normalisePredictData() { if (!impl->xPredict.empty()) impl->xp_norm = true; }

// addBias (cublasNN.cpp:330-346): the ones column.  The engine keeps its own ones row; the call is only recorded.
void cublasNN::addBiasData() { if (!impl->x.empty()) impl->x_has_bias = true; }
void cublasNN::addBiasDataValidate() { if (!impl->xValidate.empty()) impl->xv_has_bias = true; }
void cublasNN::addBiasDataPredict() { if (!impl->xPredict.empty()) impl->xp_has_bias = true; }

void cublasNN::copyDataGPU() {
    if (!impl->ensure_ctx()) return;
    if (impl->x.empty() || impl->y.empty() || !impl->x_has_bias) {
        std::cout << "copyDataGPU: setData and addBiasData must be called first" << std::endl;
        return;
    }
    if (b2nn_set_train_data_raw(impl->ctx, impl->x.v.data(), impl->y.v.data(), impl->x.rows, impl->x.cols, impl->y.cols,
                                impl->x_norm ? 1 : 0, nullptr, nullptr) != B2NN_OK) {
        const char* e = b2nn_last_error();
        std::cout << e << std::endl;            // "cudaMalloc Failed" / "cudaMemcpy Failed" in the reference (cublasNN.cpp:366, 373)
    }
}

double cublasNN::trainFuncApproxMomentum(float momentum, float rate, HiddenFn activationHidden, HiddenFn activationDerivative,
                                         int batchNum) {
    return impl->train(false, momentum, rate, activationHidden, activationDerivative, nullptr, nullptr, batchNum);
}

double cublasNN::trainClassifyMomentum(float momentum, float rate, HiddenFn activationHidden, HiddenFn activationDerivative,
                                       OutputFn activationOutput, CostFn costFunction, int batchNum) {
    return impl->train(true, momentum, rate, activationHidden, activationDerivative, activationOutput, costFunction, batchNum);
}

// validate (cublasNN.cpp:779-881).  The regression branch of the reference reads a_last, which that path never
// writes (Q7); here the cost is computed from z like calcFinalCost does.
void cublasNN::validateFuncApprox(HiddenFn activationHidden) {
    if (!impl->ctx || impl->xValidate.empty() || !impl->xv_has_bias) { std::cout << "validate: no validation data" << std::endl; return; }
    b2nn_train_desc d;
    impl->fill_desc(d, false, activationHidden, nullptr, nullptr, nullptr);
    double J = 0, errs = 0;
    if (b2nn_evaluate_raw(impl->ctx, &d, impl->xValidate.v.data(), impl->yValidate.v.data(), impl->xValidate.rows, 0,
                          impl->xv_norm ? 1 : 0, &J, &errs) != B2NN_OK) {
        std::cout << "b2nn_evaluate failed: " << b2nn_last_error() << std::endl;
        return;
    }
    std::cout << "Final Validation Cost: " << float(J) << std::endl;                                  // cublasNN.cpp:849
}

void cublasNN::validateClassify(HiddenFn activationHidden, OutputFn activationOutput, CostFn costFunction) {
    if (!impl->ctx || impl->xValidate.empty() || !impl->xv_has_bias) { std::cout << "validate: no validation data" << std::endl; return; }
    b2nn_train_desc d;
    impl->fill_desc(d, true, activationHidden, nullptr, activationOutput, costFunction);
    double J = 0, errs = 0;
    if (b2nn_evaluate_raw(impl->ctx, &d, impl->xValidate.v.data(), impl->yValidate.v.data(), impl->xValidate.rows, 1,
                          impl->xv_norm ? 1 : 0, &J, &errs) != B2NN_OK) {
        std::cout << "b2nn_evaluate failed: " << b2nn_last_error() << std::endl;
        return;
    }
    std::cout << "Final Validation Cost: " << float(J) << std::endl;
    std::cout << "Total validation errors " << '\t' << float(errs) << std::endl;                       // cublasNN.cpp:861
    std::cout << "Validation error frequency" << '\t' << (float(errs) / float(impl->xValidate.rows)) << std::endl;
}

static vector<vector<float>> run_predict(b2nn_ctx* ctx, const b2nn_train_desc& d, const std::vector<float>& x, long long rows, int K,
                                         bool normalise) {
    vector<vector<float>> result;
    const int outputs = d.classify ? 1 : K;                                                           // cublasNN.cpp:968
    std::vector<float> out(size_t(rows) * size_t(outputs));
    if (b2nn_predict_raw(ctx, &d, x.data(), rows, normalise ? 1 : 0, out.data(), nullptr) != B2NN_OK) {
        std::cout << "b2nn_predict failed: " << b2nn_last_error() << std::endl;
        return result;
    }
    result.resize(size_t(rows));
    for (long long i = 0; i < rows; ++i) {
        result[size_t(i)].resize(size_t(outputs));
        for (int j = 0; j < outputs; ++j) result[size_t(i)][size_t(j)] = out[size_t(j) * size_t(rows) + size_t(i)];
    }
    return result;
}

vector<vector<float>> cublasNN::predictFuncApprox(HiddenFn activationHidden) {
    if (!impl->ctx || impl->xPredict.empty() || !impl->xp_has_bias) { std::cout << "predict: no prediction data" << std::endl; return {}; }
    b2nn_train_desc d;
    impl->fill_desc(d, false, activationHidden, nullptr, nullptr, nullptr);
    return run_predict(impl->ctx, d, impl->xPredict.v, impl->xPredict.rows, impl->layers.back(), impl->xp_norm);
}

vector<vector<float>> cublasNN::predictClassify(HiddenFn activationHidden, OutputFn activationOutput) {
    if (!impl->ctx || impl->xPredict.empty() || !impl->xp_has_bias) { std::cout << "predict: no prediction data" << std::endl; return {}; }
    b2nn_train_desc d;
    impl->fill_desc(d, true, activationHidden, nullptr, activationOutput, nullptr);
    return run_predict(impl->ctx, d, impl->xPredict.v, impl->xPredict.rows, impl->layers.back(), impl->xp_norm);
}

void cublasNN::setSeed(unsigned long long seed) { impl->seed = seed; impl->seed_set = true; }
void cublasNN::setPrecision(int p) { impl->precision = p; }

vector<float> cublasNN::getWeights() {
    vector<float> th;
    if (!impl->ctx) return th;
    th.resize(size_t(b2nn_num_weights(impl->ctx)));
    if (b2nn_get_weights(impl->ctx, th.data(), (long long)th.size()) != B2NN_OK) th.clear();
    return th;
}
void cublasNN::setWeights(const vector<float>& theta) {
    if (impl->ctx && b2nn_set_weights(impl->ctx, theta.data(), (long long)theta.size()) != B2NN_OK)
        std::cout << "b2nn_set_weights failed: " << b2nn_last_error() << std::endl;
}
bool cublasNN::saveCheckpoint(const string& path) {
    if (!impl->ctx || b2nn_save_checkpoint(impl->ctx, path.c_str()) != B2NN_OK) {
        std::cout << "b2nn_save_checkpoint failed: " << b2nn_last_error() << std::endl;
        return false;
    }
    return true;
}
bool cublasNN::loadCheckpoint(const string& path) {
    if (!impl->ensure_ctx()) return false;
    if (b2nn_load_checkpoint(impl->ctx, path.c_str()) != B2NN_OK) {
        std::cout << "b2nn_load_checkpoint failed: " << b2nn_last_error() << std::endl;
        return false;
    }
    return true;
}
double cublasNN::lastFinalCost() const { return impl->last.final_cost; }
double cublasNN::lastErrorCount() const { return impl->last.error_count; }
double cublasNN::lastLoopSeconds() const { return impl->last.loop_seconds; }
