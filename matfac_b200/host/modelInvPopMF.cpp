#include "modelInvPopMF.h"

#include <cstring>

#include "device_session.h"

int ModelInvPopMF::deviceVariant() const { return MFB_IFWMF; }

double ModelInvPopMF::objective(const Data &data, std::unordered_set<int> &invalidUsers,
                                std::unordered_set<int> &invalidItems) {
  return Model::objective(data, invalidUsers, invalidItems);
}

// Popularity scores normalised to sum 1 over the valid ids (modelInvPopMF.cpp:98-114) and, per
// id, the weight the update uses when that side is the rarer one:
//   float wt = invPop; wt = 1.0/(1.0 + rhoRMS*wt)          (modelInvPopMF.cpp:163-168)
void ModelInvPopMF::uploadAux(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                              std::unordered_set<int> &invalidItems) {
  const int nTrainRows = (int)userFreq.size(), nTrainCols = (int)itemFreq.size();
  if (data || invPopU.empty()) {
    std::vector<int> trainUsers = matfac::validIds(nTrainRows, invalidUsers);
    std::vector<int> trainItems = matfac::validIds(nTrainCols, invalidItems);
    nTrainUsers = (int)trainUsers.size();
    nTrainItems = (int)trainItems.size();
    invPopU.clear();
    invPopI.clear();
    double sumPopScore = 0;
    for (int u : trainUsers) {
      invPopU[u] = userFreq[u] / ((double)nTrainItems);
      sumPopScore += invPopU[u];
    }
    for (int u : trainUsers) invPopU[u] = invPopU[u] / sumPopScore;
    sumPopScore = 0;
    for (int item : trainItems) {
      invPopI[item] = itemFreq[item] / ((double)nTrainUsers);
      sumPopScore += invPopI[item];
    }
    for (int item : trainItems) invPopI[item] = invPopI[item] / sumPopScore;
  }
  std::vector<int32_t> uf(nUsers, 0), itf(nItems, 0);
  std::vector<float> uw(nUsers, 1.0f), iw(nItems, 1.0f);
  for (int u = 0; u < nUsers && u < nTrainRows; u++) uf[u] = (int32_t)userFreq[u];
  for (int i = 0; i < nItems && i < nTrainCols; i++) itf[i] = (int32_t)itemFreq[i];
  for (auto &kv : invPopU) {
    float wt = kv.second;
    wt = (1.0 / (1.0 + rhoRMS * wt));
    if (kv.first < nUsers) uw[kv.first] = wt;
  }
  for (auto &kv : invPopI) {
    float wt = kv.second;
    wt = (1.0 / (1.0 + rhoRMS * wt));
    if (kv.first < nItems) iw[kv.first] = wt;
  }
  s.check(mfb_set_aux(s.eng, MFB_IFWMF, uf.data(), itf.data(), uw.data(), iw.data(), nullptr, nullptr, nullptr));
}

void ModelInvPopMF::train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                          std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelInvPopMF::train trainSeed: " << trainSeed;
  runFlatSgd(data, bestModel, invalidUsers, invalidItems, "ModelInvPopMF::train", true);
}

void ModelInvPopMF::trainSGDPar(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                                std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelInvPopMF::trainSGDPar trainSeed: " << trainSeed;
  runStratifiedSgd(data, bestModel, invalidUsers, invalidItems, "ModelInvPopMF::trainSGDPar", false);
}
