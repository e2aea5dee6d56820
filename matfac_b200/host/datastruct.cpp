#include "datastruct.h"

#include <algorithm>

#include "device_session.h"
#include "io.h"

static const char *orNull(std::string &s) { return s.empty() ? NULL : s.c_str(); }

Params::Params(int facDim, int maxIter, int svdFacDim, int seed, float uReg, float iReg, float learnRate,
               float rhoRMS, float alpha, std::string &trainMatFile, std::string &testMatFile,
               std::string &valMatFile, std::string &graphMatFile, std::string &origUFacFile,
               std::string &origIFacFile, std::string &initUFacFile, std::string &initIFacFile, std::string &prefix)
    : nUsers(-1), nItems(-1), facDim(facDim), maxIter(maxIter), svdFacDim(svdFacDim), seed(seed), uReg(uReg),
      iReg(iReg), learnRate(learnRate), rhoRMS(rhoRMS), alpha(alpha), trainMatFile(trainMatFile.c_str()),
      testMatFile(testMatFile.c_str()), valMatFile(valMatFile.c_str()), graphMatFile(orNull(graphMatFile)),
      origUFacFile(orNull(origUFacFile)), origIFacFile(orNull(origIFacFile)), initUFacFile(orNull(initUFacFile)),
      initIFacFile(orNull(initIFacFile)), prefix(prefix.c_str()) {}

void Params::display() {
  auto s = [](const char *p) { return p ? p : " "; };
  std::cout << "*** PARAMETERS ***" << std::endl;
  std::cout << "nUsers: " << nUsers << " nItems: " << nItems << std::endl;
  std::cout << "facDim: " << facDim << " svdFacDim: " << svdFacDim << std::endl;
  std::cout << "maxIter: " << maxIter << std::endl;
  std::cout << "uReg: " << uReg << " iReg: " << iReg << std::endl;
  std::cout << "rhoRMS: " << rhoRMS << " alpha: " << alpha << std::endl;
  std::cout << "learnRate: " << learnRate << std::endl;
  std::cout << "trainMat: " << trainMatFile << std::endl;
  std::cout << "testMat: " << testMatFile << std::endl;
  std::cout << "valMat: " << valMatFile << std::endl;
  std::cout << "graphMat: " << s(graphMatFile) << std::endl;
  std::cout << "origUFac: " << s(origUFacFile) << std::endl;
  std::cout << "origIFac: " << s(origIFacFile) << std::endl;
  std::cout << "initUFac: " << s(initUFacFile) << std::endl;
  std::cout << "initIFac: " << s(initIFacFile) << std::endl;
}

Data::Data(gk_csr_t *p_trainMat, gk_csr_t *p_testMat)
    : prefix(""), trainMat(p_trainMat), testMat(p_testMat), valMat(NULL), graphMat(NULL), facDim(0), trainNNZ(0) {
  nUsers = trainMat->nrows;
  nItems = trainMat->ncols;
}

static gk_csr_t *readIndexed(const char *file, const char *what) {
  std::cout << "Reading " << what << " matrix 0-indexed... " << file << std::endl;
  gk_csr_t *m = gk_csr_Read((char *)file, GK_CSR_FMT_CSR, GK_CSR_IS_VAL, 0);
  gk_csr_CreateIndex(m, GK_CSR_COL);
  return m;
}

// nUsers = rows of the train matrix, nItems = largest column id of the three matrices + 1
// (datastruct.cpp:23,91); ids beyond the train matrix become "invalid" in the trainers.
void Data::finish() {
  nUsers = trainMat->nrows;
  trainNNZ = (int)trainMat->rowptr[trainMat->nrows];
  int maxItemInd = trainMat->ncols - 1;
  if (testMat) maxItemInd = std::max(maxItemInd, testMat->ncols - 1);
  if (valMat) maxItemInd = std::max(maxItemInd, valMat->ncols - 1);
  nItems = maxItemInd + 1;
  std::cout << "\ntrain nnz = " << trainNNZ << std::endl;
  std::cout << "train nrows: " << trainMat->nrows << " ncols: " << trainMat->ncols << std::endl;
  std::cout << "minItemInd: 0 maxItemInd: " << maxItemInd << std::endl;
}

Data::Data(const Params &params)
    : prefix(params.prefix), trainMat(NULL), testMat(NULL), valMat(NULL), graphMat(NULL), facDim(params.facDim),
      trainNNZ(0), nUsers(-1), nItems(-1) {
  if (params.trainMatFile) trainMat = readIndexed(params.trainMatFile, "partial train");
  if (params.testMatFile) testMat = readIndexed(params.testMatFile, "test");
  if (params.valMatFile) valMat = readIndexed(params.valMatFile, "val");
  if (!trainMat) {
    std::cerr << "No train matrix" << std::endl;
    exit(-1);
  }
  finish();
  if (params.graphMatFile && isFileExist(params.graphMatFile))
    graphMat = gk_csr_Read((char *)params.graphMatFile, GK_CSR_FMT_CSR, 1, 0);
  if (facDim > 0) {
    if (params.origUFacFile) {
      origUFac.assign(nUsers, std::vector<double>(facDim, 0));
      readMat(origUFac, nUsers, facDim, params.origUFacFile);
    }
    if (params.origIFacFile) {
      origIFac.assign(nItems, std::vector<double>(facDim, 0));
      readMat(origIFac, nItems, facDim, params.origIFacFile);
    }
  }
}

Data::Data(gk_csr_t *train, gk_csr_t *val, gk_csr_t *test, int facDim, const char *prefix)
    : prefix(prefix), trainMat(train), testMat(test), valMat(val), graphMat(NULL), facDim(facDim), trainNNZ(0),
      nUsers(-1), nItems(-1) {
  gk_csr_t *all[3] = {trainMat, valMat, testMat};
  for (gk_csr_t *m : all)
    if (m && !m->colptr) gk_csr_CreateIndex(m, GK_CSR_COL);
  finish();
}

Data::~Data() {
  matfac::DeviceSession::dropFor(this);  // device copies of the matrices die with the host ones
  if (trainMat) gk_csr_Free(&trainMat);
  if (testMat) gk_csr_Free(&testMat);
  if (valMat) gk_csr_Free(&valMat);
  if (graphMat) gk_csr_Free(&graphMat);
}
