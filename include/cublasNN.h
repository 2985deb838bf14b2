// cublasNN.h — source-compatible façade of the reference's class cublasNN (reference cublasNN.h:22-69) over the B200
// engine's C ABI (include/b2nn.h).
//
// The reference's main.cpp compiles against this header unchanged: same class name, same public methods with the same
// argument meaning and call protocol (setLayers -> setData -> normalise* -> setIters/setDisplay -> addBias* ->
// copyDataGPU -> train* -> validate* / predict*; main.cpp:74-103, 126-166), same printed text, and the header still
// leaks `using namespace std;` because main.cpp relies on it (cublasNN.h:20).  Everything private is gone: the device
// state lives behind the opaque engine handle, host state behind a pimpl.  No CUDA headers are needed to include this.
//
// Additions (not in the reference): setSeed / setPrecision / getWeights / setWeights / save/loadCheckpoint / last*.
#ifndef CUBLASNN_B200_H
#define CUBLASNN_B200_H

#include <iostream>
#include <string>
#include <thread>
#include <vector>

#ifndef IDX2C
#define IDX2C(i, j, ld) (((j) * (ld)) + (i))  // column-major: i = row (sample), j = column, ld = rows
#endif

using namespace std;

class cublasNN {
public:
    typedef void (*HiddenFn)(const float*, float*, int);        // activation / derivative: (in, out, count)
    typedef void (*OutputFn)(const float*, float*, int, int);   // output activation: (in, out, rows, cols)
    typedef void (*CostFn)(float*, float*, float*, int);        // cost: (h, y, J, count)
    typedef void (*InitFn)(float*, int, int, int);              // initialiser: (theta, in, out, count)

    cublasNN();
    ~cublasNN();
    cublasNN(const cublasNN&) = delete;
    cublasNN& operator=(const cublasNN&) = delete;

    // CSV I/O (reference cublasNN.cpp:62-114)
    vector<vector<float>> readCSV(string fileName, bool header, float& time);
    double writeCSV(vector<vector<float>> data, string filename);

    // data (reference cublasNN.cpp:148-189); by value like the reference, the class keeps its own copies
    void setData(vector<vector<float>> xVec, vector<vector<float>> yVec, bool classify);
    void setValidateData(vector<vector<float>> xVec, vector<vector<float>> yVec, bool classify);
    void setPredictData(vector<vector<float>> xVec);

    // network (reference cublasNN.cpp:270-328); must precede setData(..., true) (main.cpp:125)
    void setLayers(int* l, int lNum, InitFn randInitWeights);
    void setIters(int i);
    void setDisplay(int i);
    void setLambda(int l);          // stored, never read — as in the reference (cublasNN.h:36, 115)

    // preprocessing (reference cublasNN.cpp:191-268, 330-346)
    void normaliseData();
    void normaliseValidateData();
    void normalisePredictData();    // uses the predict row count (the reference passes mValidate: cublasNN.h:39, Q6)
    void addBiasData();
    void addBiasDataValidate();
    void addBiasDataPredict();
    void copyDataGPU();

    // the hot path (reference cublasNN.cpp:526-640); returns seconds of the whole call like the reference
    double trainFuncApproxMomentum(float momentum, float rate, HiddenFn activationHidden, HiddenFn activationDerivative,
                                   int batchNum = 1);
    double trainClassifyMomentum(float momentum, float rate, HiddenFn activationHidden, HiddenFn activationDerivative,
                                 OutputFn activationOutput, CostFn costFunction, int batchNum = 1);

    // evaluation / inference (reference cublasNN.cpp:779-881, 906-977)
    void validateFuncApprox(HiddenFn activationHidden);
    void validateClassify(HiddenFn activationHidden, OutputFn activationOutput, CostFn costFunction);
    vector<vector<float>> predictFuncApprox(HiddenFn activationHidden);
    vector<vector<float>> predictClassify(HiddenFn activationHidden, OutputFn activationOutput);

    // ---- additions -------------------------------------------------------------------------------------------------
    void setSeed(unsigned long long seed);      // pin the cuRAND seed (default: clock(), cublasNN.cpp:27); call before setLayers
    void setPrecision(int b2nn_prec);           // B2NN_PREC_AUTO (default: fp32-grade tf32x3) / _FP32 / _TF32X3 / _TF32 (opt-in, 10-bit inputs)
    vector<float> getWeights();                 // reference arena layout
    void setWeights(const vector<float>& theta);
    bool saveCheckpoint(const string& path);    // layers + weights + velocity + normalisation statistics (b2nn_save_checkpoint)
    bool loadCheckpoint(const string& path);    // call after setLayers (or instead of it: the file carries the layer widths)
    double lastFinalCost() const;
    double lastErrorCount() const;
    double lastLoopSeconds() const;

private:
    struct Impl;
    Impl* impl;
};

#endif  // CUBLASNN_B200_H
