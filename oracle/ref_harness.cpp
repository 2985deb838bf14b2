// oracle/ref_harness.cpp — TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference (sources compiled where they
// lie under /root/reference by oracle/Makefile; nothing is copied into this repo) so that its own cuBLAS path
// can be (1) the parity anchor of oracle/mlp_oracle.cpp and of the CUDA engine, and (2) the GPU baseline that
// bench.py --impl reference times.  It needs a GPU: it is built here and runs on the B200 box only.
//
// The reference is non-reproducible and non-observable as shipped (seed = clock(), weights private:
// cublasNN.cpp:26-27, cublasNN.h:71,135).  This translation unit reaches the privates with the usual
// `#define private public` trick (standard headers are included first so only cublasNN.h is affected), pins the
// cuRAND seed before setLayers, optionally overwrites theta_0, and captures the "Iteration: i\tCost: J" lines
// the class prints to std::cout (cublasNN.cpp:555, 616, 745, 760).

#include <vector>
#include <string>
#include <thread>
#include <iostream>
#include <sstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <new>
#include <chrono>

#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuda.h>
#include <curand.h>

#define private public
#include "cublasNN.h"
#undef private
#include "activations.h"
#include "costfunctions.h"
#include "randinitweights.h"

static void trace(const char* what) {
    if (std::getenv("REF_TRACE")) { std::fprintf(stderr, "[ref_harness] %s\n", what); std::fflush(stderr); }
}

extern "C" {

struct ref_problem {
    const int*   layers;
    int          n_layers;
    const float* x;          // m x n, ROW-major rows of features, NO bias column (what readCSV would return)
    const float* y;          // m x ycols row-major: class label (ycols = 1) when classify, targets otherwise
    int64_t      m;
    int          n;
    int          ycols;
    int          batch_num;
    int          iters;
    int          display;
    float        momentum;
    float        rate;
    int          hidden_act; // 0 sigmoid, 1 tanh
    int          classify;
    int          out_act;    // 0 sigmoidOutputGPU, 1 softmaxGPU
    int          cost;       // 0 negLnMaxCostGPU, 1 crossEntropyCostGPU
    int          init_kind;  // 1 randInitWeights1GPU, 2 randInitWeights2GPU
};

struct ref_result {
    double final_cost;
    double error_count;
    int    n_logged;
    double call_seconds;     // what train*Momentum returns: alloc + split + loop + final cost (cublasNN.cpp:529-578)
    double wall_seconds;     // steady_clock around the same call, device synchronised
};

static void parse_log(const std::string& text, int* log_iter, double* log_J, int cap, ref_result* res) {
    std::istringstream in(text);
    std::string line;
    int n = 0;
    res->final_cost = 0; res->error_count = 0;
    while (std::getline(in, line)) {
        if (line.rfind("Iteration: ", 0) == 0) {
            size_t tab = line.find('\t');
            size_t c = line.find("Cost: ");
            if (tab != std::string::npos && c != std::string::npos && n < cap) {
                log_iter[n] = std::atoi(line.substr(11, tab - 11).c_str());
                log_J[n] = std::strtod(line.c_str() + c + 6, nullptr);
                ++n;
            }
        } else if (line.rfind("Final Cost: ", 0) == 0) {
            res->final_cost = std::strtod(line.c_str() + 12, nullptr);
        } else if (line.rfind("Total errors ", 0) == 0) {
            size_t tab = line.find('\t');
            if (tab != std::string::npos) res->error_count = std::strtod(line.c_str() + tab + 1, nullptr);
        }
    }
    res->n_logged = n;
}

// theta0: optional (total weights, reference layout); when null the cuRAND stream seeded with `seed` is used.
// theta0_out / theta_out: optional outputs (weights before / after training).
int ref_train(const ref_problem* p, const float* theta0, unsigned long long seed, float* theta0_out,
              float* theta_out, int* log_iter, double* log_J, int log_cap, ref_result* res) {
    if (!p || !res) return 1;
    // The destructor frees members the constructor never initialises (cublasNN.cpp:40-58); zeroed storage
    // makes those free(nullptr).
    void* mem = std::calloc(1, sizeof(cublasNN));
    if (!mem) return 2;
    trace("construct");
    cublasNN* nn = new (mem) cublasNN;
    trace("seed");
    curandSetPseudoRandomGeneratorSeed(nn->gen, seed);
    curandSetGeneratorOffset(nn->gen, 0ULL);

    std::vector<int> layers(p->layers, p->layers + p->n_layers);
    trace("setLayers");
    nn->setLayers(layers.data(), p->n_layers, p->init_kind == 1 ? randInitWeights1GPU : randInitWeights2GPU);
    trace("theta io");
    if (theta0)
        cudaMemcpy(nn->thetaBaseGPU, theta0, sizeof(float) * size_t(nn->totalThetaSize), cudaMemcpyHostToDevice);
    if (theta0_out)
        cudaMemcpy(theta0_out, nn->thetaBaseGPU, sizeof(float) * size_t(nn->totalThetaSize), cudaMemcpyDeviceToHost);

    {
        std::vector<std::vector<float>> xv(size_t(p->m)), yv(size_t(p->m));
        for (int64_t i = 0; i < p->m; ++i) {
            xv[size_t(i)].assign(p->x + i * p->n, p->x + (i + 1) * p->n);
            yv[size_t(i)].assign(p->y + i * p->ycols, p->y + (i + 1) * p->ycols);
        }
        trace("setData");
        nn->setData(xv, yv, p->classify != 0);
    }
    trace("bias+copy");
    nn->setIters(p->iters);
    nn->setDisplay(p->display > 0 ? p->display : 1);
    nn->addBiasData();
    nn->copyDataGPU();

    trace("train");
    std::ostringstream captured;
    std::streambuf* old = std::cout.rdbuf(captured.rdbuf());
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    double secs;
    void (*act)(const float*, float*, int)  = p->hidden_act == 0 ? sigmoidGPU : tanhGPU;
    void (*dact)(const float*, float*, int) = p->hidden_act == 0 ? sigmoidGradGPU : sechSqGPU;
    if (p->classify)
        secs = nn->trainClassifyMomentum(p->momentum, p->rate, act, dact,
                                         p->out_act == 0 ? sigmoidOutputGPU : softmaxGPU,
                                         p->cost == 0 ? negLnMaxCostGPU : crossEntropyCostGPU, p->batch_num);
    else
        secs = nn->trainFuncApproxMomentum(p->momentum, p->rate, act, dact, p->batch_num);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    std::cout.rdbuf(old);
    trace("trained");

    parse_log(captured.str(), log_iter, log_J, log_cap, res);
    res->call_seconds = secs;
    res->wall_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (theta_out)
        cudaMemcpy(theta_out, nn->thetaBaseGPU, sizeof(float) * size_t(nn->totalThetaSize), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();   // the reference double-frees some device buffers (cublasNN.cpp:742 vs 897)
    (void)e;

    trace("destroy");
    nn->~cublasNN();
    std::free(mem);
    trace("done");
    return 0;
}

// predictClassify / predictFuncApprox (cublasNN.cpp:906-977) with a given theta.  out: m x 1 (class id) when
// classify, m x l_{L-1} row-major otherwise.
int ref_predict(const ref_problem* p, const float* theta, float* out) {
    if (!p || !theta || !out) return 1;
    void* mem = std::calloc(1, sizeof(cublasNN));
    if (!mem) return 2;
    cublasNN* nn = new (mem) cublasNN;
    std::vector<int> layers(p->layers, p->layers + p->n_layers);
    nn->setLayers(layers.data(), p->n_layers, randInitWeights1GPU);
    cudaMemcpy(nn->thetaBaseGPU, theta, sizeof(float) * size_t(nn->totalThetaSize), cudaMemcpyHostToDevice);
    {
        std::vector<std::vector<float>> xv(size_t(p->m));
        for (int64_t i = 0; i < p->m; ++i) xv[size_t(i)].assign(p->x + i * p->n, p->x + (i + 1) * p->n);
        nn->n = p->n;                    // predict() reads n, normally set by setData (cublasNN.cpp:151)
        nn->setPredictData(xv);
    }
    nn->addBiasDataPredict();
    void (*act)(const float*, float*, int) = p->hidden_act == 0 ? sigmoidGPU : tanhGPU;
    std::vector<std::vector<float>> r =
        p->classify ? nn->predictClassify(act, p->out_act == 0 ? sigmoidOutputGPU : softmaxGPU)
                    : nn->predictFuncApprox(act);
    size_t cols = r.empty() ? 0 : r[0].size();
    for (size_t i = 0; i < r.size(); ++i)
        for (size_t j = 0; j < cols; ++j) out[i * cols + j] = r[i][j];
    nn->x = nullptr; nn->y = nullptr;
    nn->~cublasNN();
    std::free(mem);
    return 0;
}

int ref_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
