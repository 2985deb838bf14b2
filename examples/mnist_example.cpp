// Config C1 — the reference's classification example (main.cpp:108-167, commented out there), same class API.
// Usage: mnist_example <trainX.csv> <trainY.csv> [iters] [batchNum] [seed]
// The reference checkout ships only the label CSVs; any 784-column pixel CSV works.
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "cublasNN.h"
#include "randinitweights.h"
#include "activations.h"
#include "costfunctions.h"

int main(int argc, char** argv) {
    if (argc < 3) { cout << "usage: mnist_example trainX.csv trainY.csv [iters] [batchNum] [seed]" << endl; return 2; }
    const int iters = argc > 3 ? std::atoi(argv[3]) : 150;
    const int batchNum = argc > 4 ? std::atoi(argv[4]) : 10;
    cout << "Neural Network Classification Example" << endl;
    cublasNN* nn = new cublasNN;
    if (argc > 5) nn->setSeed(std::strtoull(argv[5], nullptr, 10));
    if (const char* p = std::getenv("B2NN_EXAMPLE_FP32")) { (void)p; nn->setPrecision(0); }
    float csvTime = 0, t = 0;
    vector<vector<float>> xVec = nn->readCSV(argv[1], false, t); csvTime += t;
    vector<vector<float>> yVec = nn->readCSV(argv[2], false, t); csvTime += t;
    cout << "Read CSV: " << csvTime << " s" << endl;
    if (xVec.empty() || yVec.empty()) return 2;

    int layers[3] = {int(xVec[0].size()), 50, 10};
    nn->setLayers(layers, 3, randInitWeights2GPU);          // must precede setData(..., true) (main.cpp:125)
    nn->setData(xVec, yVec, true);
    nn->setPredictData(xVec);
    nn->normaliseData();
    nn->normalisePredictData();
    nn->setIters(iters);
    nn->setDisplay(1);
    nn->addBiasData();
    nn->addBiasDataPredict();
    nn->copyDataGPU();

    double gpuTime = nn->trainClassifyMomentum(0.9, 0.075, tanhGPU, sechSqGPU, softmaxGPU, crossEntropyCostGPU, batchNum);  // main.cpp:157
    cout << "GPU Training " << gpuTime << " s" << endl;
    vector<vector<float>> result = nn->predictClassify(tanhGPU, softmaxGPU);
    size_t wrong = 0;
    for (size_t i = 0; i < result.size(); ++i) wrong += (result[i][0] != yVec[i][0]);
    cout << "predict errors on the training set: " << wrong << endl;
    delete nn;
    return 0;
}
