// Model: parameter state, evaluation, early stopping and persistence — the reference's public
// class (model.h:22-265) with the same field names, constructor and method signatures for
// everything on the training path.  The method bodies run on the GPU through the C ABI of
// include/mfb.h; host-visible uFac/iFac keep (row = user|item, col = latent dim) semantics and
// are synchronised with the device before any trainer returns.
//
// Not carried over (off the training path): ranking metrics (hitRate/arHR/NDCG),
// sub-matrix errors, mean/variance helpers, the SVD-regularised and incremental trainers.
#ifndef _MODEL_H_
#define _MODEL_H_

#include <Eigen/Cholesky>
#include <Eigen/Dense>
#include <Eigen/LU>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <string>
#include <tuple>
#include <unordered_set>
#include <vector>

#include "GKlib.h"
#include "const.h"
#include "datastruct.h"
#include "defs.h"
#include "util.h"

namespace matfac {
class DeviceSession;
}

class Model {
 public:
  int nUsers;
  int nItems;
  int facDim;
  int trainSeed;
  float origLearnRate;
  float learnRate;
  float rhoRMS;  // IFWMF: weight scale; TMF: sigmoid steepness
  float alpha;   // TMF: sigmoid centre
  int maxIter;
  float uReg;
  float iReg;
  float sing_a, sing_b;
  Eigen::MatrixXf uFac;
  Eigen::MatrixXf iFac;
  Eigen::VectorXf uBias;
  Eigen::VectorXf iBias;
  Eigen::VectorXf singularVals;
  double mu;

  Model(const Params &params);
  Model(int nUsers, int nItems, int facDim);
  Model(int nUsers, int nItems, const Params &params);
  Model(const Params &params, int seed);
  Model(const Params &params, const char *uFacName, const char *iFacName, int seed);
  Model(const Params &params, const char *uFacName, const char *iFacName, const char *uBFName, const char *iBFName,
        const char *gBFName, int seed);
  virtual ~Model() {}

#define MATFAC_TRAINER(name)                                                                                  \
  virtual void name(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,                \
                    std::unordered_set<int> &invalidItems) {                                                  \
    (void)data; (void)bestModel; (void)invalidUsers; (void)invalidItems;                                      \
    std::cerr << "\n" #name ": method not in base class" << std::endl;                                        \
  }
  MATFAC_TRAINER(train)
  MATFAC_TRAINER(trainSGDPar)
  MATFAC_TRAINER(trainUShuffle)
  MATFAC_TRAINER(trainALS)
  MATFAC_TRAINER(trainCCDPP)
  MATFAC_TRAINER(trainCCDPPFreqAdap)
  MATFAC_TRAINER(trainCCD)
  MATFAC_TRAINER(hogTrain)
#undef MATFAC_TRAINER

  // sum of squared errors over the train matrix + uReg |U|^2 + iReg |V|^2 (model.cpp:1694, :1770)
  virtual double objective(const Data &data);
  virtual double objective(const Data &data, std::unordered_set<int> &invalidUsers,
                           std::unordered_set<int> &invalidItems);
  // early stopping on validation RMSE, LR halving, NaN recovery, best snapshot (model.cpp:1471-1540)
  virtual bool isTerminateModel(Model &bestModel, const Data &data, int iter, int &bestIter, double &bestObj,
                                double &prevObj, double &bestValRMSE, double &prevValRMSE,
                                std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems);
  // objective-only variant (model.cpp:1421-1468 without the rating list)
  bool isTerminateModel(Model &bestModel, const Data &data, int iter, int &bestIter, double &bestObj, double &prevObj,
                        std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems);
  double RMSE(gk_csr_t *mat);
  double RMSE(gk_csr_t *mat, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems);
  // filtered variants (model.cpp:348-486): {count, RMSE} / {count, squared error} over the ratings whose item
  // (RMSE, SE) resp. user (RMSEU) is in the filter set
  std::pair<int, double> RMSE(gk_csr_t *mat, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                              std::unordered_set<int> &invalidItems);
  std::pair<int, double> SE(gk_csr_t *mat, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                            std::unordered_set<int> &invalidItems);
  std::pair<int, double> RMSEU(gk_csr_t *mat, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                               std::unordered_set<int> &invalidItems);
  // all parts of quartileRMSEs (main.cpp:700-768) in ONE device pass: userGroup / itemGroup give every id a part
  // 0..7 or 255; out[((side * 8) + part) * 2 + {0, 1}] = squared error, count (side 0 = items, 1 = users)
  void groupSE(gk_csr_t *mat, const std::vector<uint8_t> &userGroup, const std::vector<uint8_t> &itemGroup,
               std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, double out[32]);
  // ranking metrics (model.cpp:760-1332).  The candidate scan — every item scored for every user — runs on the device
  // (mfb_rank_positions: dense U V^T on the tensor cores where the model allows), the heaps of the reference reduce to
  // one position per user; NDCG ranks the device's predictions of the test ratings (mfb_predict).
  double hitRate(const Data &data, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<int, double> hitRateU(const Data &data, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                  std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<int, double> hitRateI(const Data &data, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                  std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  double arHR(const Data &data, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<double, double> arHRU(const Data &data, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                  std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<double, double> arHRI(const Data &data, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                  std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  double NDCG(std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<int, double> NDCGU(std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                               std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  std::pair<int, double> NDCGI(std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                               std::unordered_set<int> &invalidItems, gk_csr_t *testMat);
  virtual double estRating(int user, int item);
  std::string modelSignature();
  void display();
  void save(std::string prefix);
  void saveFacs(std::string prefix);
  void load(std::string prefix);
  void loadFacs(std::string prefix);
  void load(const char *uFacName, const char *iFacName);
  void saveBinFacs(std::string prefix);
  void loadBinFacs(std::string prefix);

 protected:
  // ---- device plumbing -------------------------------------------------------------------
  matfac::DeviceSession *dev_ = nullptr;  // non-null while a trainer of this object is running
  bool bestOnDevice_ = false;             // the device holds a newer best snapshot than bestModel's host copy
  int evalCounter_ = 0;

  virtual int deviceVariant() const;  // MFB_MF / MFB_IFWMF / MFB_TMF / MFB_TMFDROPOUT
  // upload the per-user / per-item auxiliaries this model's update and prediction rules need
  virtual void uploadAux(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                         std::unordered_set<int> &invalidItems);
  void uploadAuxAll(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                    std::unordered_set<int> &invalidItems);
  void copyScalarsFrom(const Model &o);
  void uploadFactors(matfac::DeviceSession &s);
  double deviceEval(matfac::DeviceSession &s, int which, bool objective);
  // per user: position of its test item among the candidates (mfb_rank_positions) and the test item
  void rankPositions(const Data &data, gk_csr_t *testMat, std::unordered_set<int> &invalidUsers,
                     std::unordered_set<int> &invalidItems, std::vector<int32_t> &pos, std::vector<int32_t> &testItem);
  // {hits, counted users} of the hit-rate family: N = list length, reciprocal = arHR's 1 / (pos + 1) instead of 1
  std::pair<double, double> hitStats(const Data &data, gk_csr_t *testMat, std::unordered_set<int> &invalidUsers,
                                     std::unordered_set<int> &invalidItems, const std::unordered_set<int> *filtUsers,
                                     const std::unordered_set<int> *filtItems, int N, bool reciprocal);
  std::pair<int, double> ndcgStats(gk_csr_t *testMat, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems,
                                   const std::unordered_set<int> *filtUsers, const std::unordered_set<int> *filtItems);

  // shared trainer skeletons
  struct Stop {
    int bestIter = -1;
    double bestObj = 0, prevObj = 0, bestValRMSE = 0, prevValRMSE = 0;
  };
  // groupMode: matfac::GroupMode — how the session's engines (one per visible GPU) share this trainer's work
  void beginTraining(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                     std::unordered_set<int> &invalidItems, Stop &st, const char *tag, int groupMode = 0);
  void endTraining(Model &bestModel);
  void syncBest(Model &bestModel);
  // returns true when training must stop; prints the reference's progress line
  bool afterEpoch(const Data &data, Model &bestModel, int iter, Stop &st, std::unordered_set<int> &invalidUsers,
                  std::unordered_set<int> &invalidItems, double subIterDuration, const char *tag, bool saves);
  void runFlatSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                  std::unordered_set<int> &invalidItems, const char *tag, bool saves);
  void runStratifiedSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                        std::unordered_set<int> &invalidItems, const char *tag, bool saves);
  void runUserMajorSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems, const char *tag, bool saves);
};

#endif
