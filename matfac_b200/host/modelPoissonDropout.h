// ModelPoissonDropout (TMF + Dropout): as TMF, but the update rank of a rating is drawn from
// Poisson(lambda), lambda = the TMF rank, and prediction uses the smallest rank whose Poisson CDF
// reaches 0.99.  Class shape of modelPoissonDropout.h:23-158.
#ifndef _MODEL_POISSON_DROPOUT_H_
#define _MODEL_POISSON_DROPOUT_H_

#include "modelDropoutSigmoid.h"

class ModelPoissonDropout : public ModelDropoutSigmoid {
 public:
  ModelPoissonDropout(const Params &params, std::vector<double> &userRankMap, std::vector<double> &itemRankMap,
                      std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : ModelDropoutSigmoid(params, userRankMap, itemRankMap, userFreq, itemFreq) {
    initCDFRanks();
  }
  ModelPoissonDropout(const Params &params, int seed, std::vector<double> &userRankMap,
                      std::vector<double> &itemRankMap, std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : ModelDropoutSigmoid(params, seed, userRankMap, itemRankMap, userFreq, itemFreq) {
    initCDFRanks();
  }

  void initCDFRanks();  // modelPoissonDropout.cpp:25-47
  double estRating(int user, int item) override;
  void train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
             std::unordered_set<int> &invalidItems) override;

 protected:
  int deviceVariant() const override;
  void uploadAux(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                 std::unordered_set<int> &invalidItems) override;
};

#endif
