#include "modelPoissonDropout.h"

#include <cassert>
#include <cmath>

#include "device_session.h"

// For every lambda in 1..facDim: the last dimension index k such that the Poisson(lambda) CDF up
// to k+1 events reaches 0.99 (capped at facDim-1).
void ModelPoissonDropout::initCDFRanks() {
  cdfRanks = std::vector<int>(facDim, 0);
  for (int lambda = 1; lambda <= facDim; lambda++) {
    double cdf = std::exp(-lambda) * (std::pow(lambda, 0) / factorial[0]);
    int k = 0;
    for (k = 0; k < facDim; k++) {
      double wt = std::exp(-lambda) * (std::pow(lambda, k + 1) / factorial[k + 1]);
      cdf += wt;
      if (cdf >= 0.99) break;
    }
    cdfRanks[lambda - 1] = (k == facDim) ? k - 1 : k;
    std::cout << "cdfRank: " << lambda - 1 << " " << cdfRanks[lambda - 1] << std::endl;
  }
}

double ModelPoissonDropout::estRating(int user, int item) {
  const bool isUMinFreq = userFreq[user] < itemFreq[item];
  const int lambda = sigmoidRank(isUMinFreq ? userFreq[user] : itemFreq[item]);
  assert(lambda > 0);
  double rat = 0;
  for (int k = 0; k <= cdfRanks[lambda - 1] && k < facDim; k++) rat += uFac(user, k) * iFac(item, k);
  return rat;
}

int ModelPoissonDropout::deviceVariant() const { return MFB_TMFDROPOUT; }

void ModelPoissonDropout::uploadAux(matfac::DeviceSession &s, const Data *, std::unordered_set<int> &,
                                    std::unordered_set<int> &) {
  const int r = facDim;
  std::vector<int32_t> uf(nUsers, 0), itf(nItems, 0), ul(nUsers, 1), il(nItems, 1), up(nUsers, 1), ip(nItems, 1);
  auto lam = [&](double f) { return std::min(std::max(sigmoidRank(f), 1), r); };
  auto pred = [&](int lambda) { return std::min(cdfRanks[lambda - 1] + 1, r); };
  for (int u = 0; u < nUsers && u < (int)userFreq.size(); u++) {
    uf[u] = (int32_t)userFreq[u];
    ul[u] = lam(userFreq[u]);
    up[u] = pred(ul[u]);
  }
  for (int i = 0; i < nItems && i < (int)itemFreq.size(); i++) {
    itf[i] = (int32_t)itemFreq[i];
    il[i] = lam(itemFreq[i]);
    ip[i] = pred(il[i]);
  }
  // row l-1: P(Poisson(l) <= k), k = 0..r-1 — the device draws the update rank by CDF inversion
  // from a counter-based generator (the reference draws std::poisson_distribution on one
  // mt19937(seed+t) per OpenMP thread, modelPoissonDropout.cpp:118-121,200-201; the streams are
  // thread-count dependent there, so parity is distributional)
  std::vector<float> cdf((size_t)r * r);
  for (int l = 1; l <= r; l++) {
    double p = std::exp(-(double)l), c = p;
    for (int k = 0; k < r; k++) {
      cdf[(size_t)(l - 1) * r + k] = (float)std::min(c, 1.0);
      p = p * l / (k + 1);
      c += p;
    }
  }
  s.check(mfb_set_aux(s.eng, MFB_TMFDROPOUT, uf.data(), itf.data(), ul.data(), il.data(), up.data(), ip.data(),
                      cdf.data()));
}

void ModelPoissonDropout::train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                                std::unordered_set<int> &invalidItems) {
  std::cout << "\nModelPoissonDropout ::train trainSeed: " << trainSeed;
  std::cout << "\nrhoRMS: " << rhoRMS << " alpha: " << alpha << std::endl;
  runStratifiedSgd(data, bestModel, invalidUsers, invalidItems, "ModelPoissonDropout::train", false);
}
