// Params and Data: the hyper-parameter bag and the three rating matrices, with the field names
// and constructor signatures of the reference (datastruct.h:12-69 Params, :72-136 Data).
#ifndef _DATASTRUCT_H_
#define _DATASTRUCT_H_

#include <iostream>
#include <numeric>
#include <string>
#include <vector>

#include "GKlib.h"
#include "const.h"

class Params {
 public:
  int nUsers;
  int nItems;
  int facDim;
  int maxIter;
  int svdFacDim;
  int seed;
  float uReg;
  float iReg;
  float learnRate;
  float rhoRMS;
  float alpha;
  const char *trainMatFile;
  const char *testMatFile;
  const char *valMatFile;
  const char *graphMatFile;
  const char *origUFacFile;
  const char *origIFacFile;
  const char *initUFacFile;
  const char *initIFacFile;
  const char *prefix;

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
  gk_csr_t *trainMat;
  gk_csr_t *testMat;
  gk_csr_t *valMat;
  gk_csr_t *graphMat;
  std::vector<std::vector<double>> origUFac;
  std::vector<std::vector<double>> origIFac;
  int facDim;
  int trainNNZ;
  int nUsers;
  int nItems;

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

#endif
