#include "modelDropoutSigmoid.h"

#include <cassert>
#include <cmath>

#include "device_session.h"

// factorial table, min/max and mean/std of the concatenated user+item frequency vector
// (modelDropoutSigmoid.h:75-95)
void ModelDropoutSigmoid::initFreqStats(int r) {
  factorial.push_back(1);
  for (int i = 1; i <= r + 1; i++) factorial.push_back(factorial.back() * ((double)i));
  fDimWt = std::vector<double>(r, 0);
  minFreq = std::min(minVec(userFreq), minVec(itemFreq));
  maxFreq = std::max(maxVec(userFreq), maxVec(itemFreq));
  std::vector<double> all(userFreq.begin(), userFreq.end());
  all.insert(all.end(), itemFreq.begin(), itemFreq.end());
  auto ms = meanStdDev(all);
  meanFreq = ms.first;
  stdFreq = ms.second;
}

int ModelDropoutSigmoid::sigmoidRank(double freq) const {
  double scaleFreq = (freq - meanFreq) / stdFreq;
  double sigmPc = 1.0 / (1.0 + exp(-rhoRMS * (scaleFreq - alpha)));
  return (int)std::ceil(sigmPc * ((double)facDim));
}

double ModelDropoutSigmoid::estRating(int user, int item) {
  const bool isUMinFreq = userFreq[user] < itemFreq[item];
  int updMinRank = sigmoidRank(isUMinFreq ? userFreq[user] : itemFreq[item]);
  assert(updMinRank > 0);
  if (updMinRank > facDim) updMinRank = facDim;
  double rat = 0;
  for (int k = 0; k < updMinRank; k++) rat += uFac(user, k) * iFac(item, k);
  return rat;
}

int ModelDropoutSigmoid::deviceVariant() const { return MFB_TMF; }

void ModelDropoutSigmoid::uploadAux(matfac::DeviceSession &s, const Data *, std::unordered_set<int> &,
                                    std::unordered_set<int> &) {
  std::vector<int32_t> uf(nUsers, 0), itf(nItems, 0), ur(nUsers, 1), ir(nItems, 1);
  auto clampRank = [&](int k) {
    if (k < EPS) k = 1;  // modelDropoutSigmoid.cpp:165-170
    if (k > facDim) k = facDim;
    return k;
  };
  for (int u = 0; u < nUsers && u < (int)userFreq.size(); u++) {
    uf[u] = (int32_t)userFreq[u];
    ur[u] = clampRank(sigmoidRank(userFreq[u]));
  }
  for (int i = 0; i < nItems && i < (int)itemFreq.size(); i++) {
    itf[i] = (int32_t)itemFreq[i];
    ir[i] = clampRank(sigmoidRank(itemFreq[i]));
  }
  // prediction truncates at the same rank (modelDropoutSigmoid.cpp:12-18)
  s.check(mfb_set_aux(s.eng, MFB_TMF, uf.data(), itf.data(), ur.data(), ir.data(), ur.data(), ir.data(), nullptr));
}

void ModelDropoutSigmoid::train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                                std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelDropoutSigmoid ::train trainSeed: " << trainSeed;
  std::cout << "\nrhoRMS: " << rhoRMS << " alpha: " << alpha << " minFreq: " << minFreq << " maxFreq: " << maxFreq
            << std::endl;
  runStratifiedSgd(data, bestModel, invalidUsers, invalidItems, "trainSigmoid", false);
}
