#include "modelMF.h"

#include "device_session.h"

using matfac::DeviceSession;

void ModelMF::train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                    std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelMF::train trainSeed: " << trainSeed;
  std::cout << "\nObj b4 svd: " << objective(data) << " Train RMSE: " << RMSE(data.trainMat)
            << " Train nnz: " << data.trainNNZ << std::endl;
  runFlatSgd(data, bestModel, invalidUsers, invalidItems, "ModelMF::train", true);
}

void ModelMF::hogTrain(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelMF::hogTrain trainSeed: " << trainSeed;
  runFlatSgd(data, bestModel, invalidUsers, invalidItems, "ModelMF::hogTrain", true);
}

// --mf_method sgdu (modelMF.cpp:560-706): the valid users in shuffled order, each user's ratings back to back.  On the
// device this is the user-major kernel of the stratified trainers over the whole matrix as one block: a sub-warp owns a
// user's run and keeps u in registers for all of it (exactly the reference's inner loop for that user), many users
// run at once and item rows are updated by reductions.
void ModelMF::trainUShuffle(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                            std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelMF::train trainSeed: " << trainSeed;
  std::cout << "\nObj b4 svd: " << objective(data) << " Train RMSE: " << RMSE(data.trainMat)
            << " Train nnz: " << data.trainNNZ << std::endl;
  runUserMajorSgd(data, bestModel, invalidUsers, invalidItems, "ModelMF::trainUShuffle", true);
}

void ModelMF::trainSGDPar(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                          std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelMF::trainSGDPar trainSeed: " << trainSeed;
  std::cout << "\nObj b4 svd: " << objective(data) << " Train RMSE: " << RMSE(data.trainMat)
            << " Train nnz: " << data.trainNNZ << std::endl;
  // the reference never writes factor files from this trainer (modelMF.cpp:337,346)
  runStratifiedSgd(data, bestModel, invalidUsers, invalidItems, "ModelMF::trainSGDPar", false);
}

// One epoch = user half-step over the CSR, then item half-step over the CSC with the fresh U
// (modelMF.cpp:795-882).
void ModelMF::trainALS(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelMF::trainALS trainSeed: " << trainSeed;
  Stop st;
  // rows sharded over every visible GPU: each engine solves its row range and stores the solved rows into all peers
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, "ModelMF::trainALS", matfac::GROUP_ROWS);
  DeviceSession &s = *dev_;
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    for (int r = 0; r < s.world(); r++) s.check(mfb_als_half_step(s.engineOf(r), MFB_USER, uReg));
    for (int r = 0; r < s.world(); r++) s.check(mfb_als_half_step(s.engineOf(r), MFB_ITEM, iReg));
    s.check(mfb_event_record(s.eng, 1));
    float ms = 0;
    s.check(mfb_event_elapsed_ms(s.eng, 0, 1, &ms));
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, ms * 1e-3, "ModelMF::trainALS", true)) break;
  }
  endTraining(bestModel);
  bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}

void ModelMF::runCcdpp(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems, bool freqAdap, const char *tag) {
  std::cout << "\n" << tag << " trainSeed: " << trainSeed;
  Stop st;
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, tag, matfac::GROUP_ROWS);
  DeviceSession &s = *dev_;
  if (freqAdap) {
    // itemFreq of the train matrix feeds the "fewer than 75 ratings" rule (modelMF.cpp:1204-1206,1336)
    auto freq = getRowColFreq(data.trainMat);
    std::vector<int32_t> uf(nUsers, 0), itf(nItems, 0);
    for (size_t u = 0; u < freq.first.size(); u++) uf[u] = (int32_t)freq.first[u];
    for (size_t i = 0; i < freq.second.size(); i++) itf[i] = (int32_t)freq.second[i];
    for (int r = 0; r < s.world(); r++)
      s.check(mfb_set_aux(s.engineOf(r), MFB_MF, uf.data(), itf.data(), nullptr, nullptr, nullptr, nullptr, nullptr));
  }
  std::mt19937 mt(trainSeed);
  std::vector<int> dims(facDim);
  std::iota(dims.begin(), dims.end(), 0);
  for (int r = 0; r < s.world(); r++) s.check(mfb_ccdpp_begin(s.engineOf(r)));  // residual = ratings, U = 0 (modelMF.cpp:1013,1020)
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    if (!freqAdap) std::shuffle(dims.begin(), dims.end(), mt);  // commented out in the FreqAdap twin (:1271)
    for (int k : dims)  // step by step for all engines: the barriers between the passes run on the devices
      for (int r = 0; r < s.world(); r++)
        s.check(mfb_ccdpp_rank1(s.engineOf(r), k, iter == 0, 5, uReg, iReg, freqAdap ? 75 : 0));
    s.check(mfb_event_record(s.eng, 1));
    float ms = 0;
    s.check(mfb_event_elapsed_ms(s.eng, 0, 1, &ms));
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, ms * 1e-3, tag, true)) break;
  }
  for (int r = 0; r < s.world(); r++) s.check(mfb_ccdpp_end(s.engineOf(r)));
  endTraining(bestModel);
  bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}

void ModelMF::trainCCDPP(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                         std::unordered_set<int> &invalidItems) {
  runCcdpp(data, bestModel, invalidUsers, invalidItems, false, "ModelMF::trainCCDPP");
}

void ModelMF::trainCCDPPFreqAdap(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                                 std::unordered_set<int> &invalidItems) {
  runCcdpp(data, bestModel, invalidUsers, invalidItems, true, "ModelMF::trainCCDPPFreqAdap");
}

// --mf_method ccd (modelMF.cpp:1426-1653): coordinate descent row by row over the residual matrix.  The visiting
// order of every row's dims comes from the reference's single mt19937(trainSeed) — one std::shuffle per valid user in
// index order, then one per valid item, every epoch (:1495,1537,1577; defined for one thread there) — drawn here and
// handed to the device with the half-step.  One engine: the two sides exchange residuals.
void ModelMF::trainCCD(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems) {
  const char *tag = "ModelMF::trainCCD";
  std::cout << "\n" << tag << " trainSeed: " << trainSeed;
  if (facDim > 256) {
    std::cerr << "\ntrainCCD: facDim > 256 not supported" << std::endl;
    return;
  }
  Stop st;
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, tag, matfac::GROUP_NONE);
  DeviceSession &s = *dev_;
  std::mt19937 mt(trainSeed);
  std::vector<int> dims(facDim);
  std::iota(dims.begin(), dims.end(), 0);
  std::vector<uint8_t> uOrder((size_t)nUsers * facDim), iOrder((size_t)nItems * facDim);
  const int nCols = data.trainMat->ncols;
  s.check(mfb_ccdpp_begin(s.eng));  // residual = gk_csr_Dup(trainMat), U = 0 (:1511,1518)
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    for (int u = 0; u < nUsers; u++) {
      if (invalidUsers.count(u) > 0) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (int k = 0; k < facDim; k++) uOrder[(size_t)u * facDim + k] = (uint8_t)udims[k];
    }
    s.check(mfb_ccd_half_step(s.eng, MFB_USER, uReg, uOrder.data()));
    for (int item = 0; item < nItems; item++) {
      if (invalidItems.count(item) > 0 || item >= nCols) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (int k = 0; k < facDim; k++) iOrder[(size_t)item * facDim + k] = (uint8_t)udims[k];
    }
    s.check(mfb_ccd_half_step(s.eng, MFB_ITEM, iReg, iOrder.data()));
    s.check(mfb_event_record(s.eng, 1));
    float ms = 0;
    s.check(mfb_event_elapsed_ms(s.eng, 0, 1, &ms));
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, ms * 1e-3, tag, true)) break;
  }
  s.check(mfb_ccdpp_end(s.eng));
  endTraining(bestModel);
  bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}
