// Config C2 — the reference's function-approximation example (main.cpp:49-104, commented out there), written against
// the SAME class API.  Usage: bikeshare_example <dir with trainP7_1.csv ...> [iters] [seed]
// Builds with any C++11 compiler against include/ and links libcublasNN_b200.so + libb2nn.so; no CUDA headers needed.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "cublasNN.h"
#include "randinitweights.h"
#include "activations.h"
#include "costfunctions.h"

int main(int argc, char** argv) {
    const std::string dir = argc > 1 ? argv[1] : "../bikeshare";
    const int iters = argc > 2 ? std::atoi(argv[2]) : 30000;
    cout << "Neural Network Function Approximation Example" << endl;
    cublasNN* nn = new cublasNN;
    if (argc > 3) nn->setSeed(std::strtoull(argv[3], nullptr, 10));
    if (const char* p = std::getenv("B2NN_EXAMPLE_FP32")) { (void)p; nn->setPrecision(0); }
    cout << "Read CSV: " << endl;
    float csvTime = 0, t = 0;
    vector<vector<float>> xVec = nn->readCSV(dir + "/trainP7_1.csv", false, t); csvTime += t;
    vector<vector<float>> yVec = nn->readCSV(dir + "/trainYP2_1.csv", false, t); csvTime += t;
    vector<vector<float>> xVecValidate = nn->readCSV(dir + "/validateP7_1.csv", false, t); csvTime += t;
    vector<vector<float>> yVecValidate = nn->readCSV(dir + "/validateY1.csv", false, t); csvTime += t;
    vector<vector<float>> xVecPredict = nn->readCSV(dir + "/testPF1_1.csv", false, t); csvTime += t;
    cout << csvTime << " s" << endl;
    if (xVec.empty() || yVec.empty()) return 2;

    for (size_t i = 0; i < xVec.size(); i++) xVec[i].erase(xVec[i].begin() + 1);                   // main.cpp:66-71
    for (size_t i = 0; i < xVecValidate.size(); i++) xVecValidate[i].erase(xVecValidate[i].begin() + 1);
    for (size_t i = 0; i < xVecPredict.size(); i++) xVecPredict[i].erase(xVecPredict[i].begin() + 1);

    int layers[4] = {10, 40, 160, 1};
    nn->setLayers(layers, 4, randInitWeights1GPU);
    nn->setData(xVec, yVec, false);
    nn->setValidateData(xVecValidate, yVecValidate, false);
    nn->setPredictData(xVecPredict);
    nn->normaliseData();
    nn->normaliseValidateData();
    nn->normalisePredictData();
    nn->setIters(iters);
    nn->setDisplay(iters >= 10 ? iters / 10 : 1);
    nn->addBiasData();
    nn->addBiasDataValidate();
    nn->addBiasDataPredict();
    nn->copyDataGPU();

    double gpuTime = nn->trainFuncApproxMomentum(0.9, 0.003, sigmoidGPU, sigmoidGradGPU, 1);        // main.cpp:96
    cout << "GPU Training " << gpuTime << " s" << endl;
    nn->validateFuncApprox(sigmoidGPU);
    cout << "Write CSV" << endl;
    vector<vector<float>> result = nn->predictFuncApprox(sigmoidGPU);
    cout << nn->writeCSV(result, "result.csv") << " s" << endl;
    cout << "rows predicted: " << result.size() << endl;
    delete nn;
    return 0;
}
