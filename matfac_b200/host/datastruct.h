// Params and Data: the hyper-parameter bag and the three rating matrices, with the field names
// and constructor signatures of the reference (datastruct.h:12-69 Params, :72-136 Data).
#pragma once

#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "GKlib.h"
#include "const.h"

class Params {
 public:
  // problem and model size, iteration budget, seeds
  int nUsers, nItems, facDim, maxIter, svdFacDim, seed;
  // regularisation, step size, and the two knobs of the frequency-aware models (IFWMF / TMF)
  float uReg, iReg, learnRate, rhoRMS, alpha;
  // inputs and output prefix (borrowed C strings, see the constructor)
  const char *trainMatFile, *testMatFile, *valMatFile, *graphMatFile;
  const char *origUFacFile, *origIFacFile, *initUFacFile, *initIFacFile, *prefix;

  // The strings are borrowed (datastruct.h:43-50): they must outlive the Params object.
  Params(int facDim, int maxIter, int svdFacDim, int seed, float uReg, float iReg, float learnRate, float rhoRMS,
         float alpha, std::string &trainMatFile, std::string &testMatFile, std::string &valMatFile,
         std::string &graphMatFile, std::string &origUFacFile, std::string &origIFacFile, std::string &initUFacFile,
         std::string &initIFacFile, std::string &prefix);

  void display();
};

class Data {
 public:
  const char *prefix;
  gk_csr_t *trainMat, *testMat, *valMat, *graphMat;        // owned; freed by the destructor
  std::vector<std::vector<double>> origUFac, origIFac;      // ground-truth factors of synthetic inputs, if given
  int facDim, trainNNZ, nUsers, nItems;

  Data(gk_csr_t *p_trainMat, gk_csr_t *p_testMat);
  // Reads the three text-CSR files and builds their column indices (datastruct.cpp:3-120).
  Data(const Params &params);
  // Takes ownership of three in-memory matrices (column indices are built here).
  Data(gk_csr_t *train, gk_csr_t *val, gk_csr_t *test, int facDim, const char *prefix);
  ~Data();

 private:
  Data(const Data &);
  Data &operator=(const Data &);
  void finish();
};

