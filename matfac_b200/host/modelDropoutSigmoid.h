// ModelDropoutSigmoid (TMF): truncated MF — a rating is predicted and updated with the first
// k = ceil(facDim * sigmoid(rhoRMS * (z - alpha))) dimensions, z the z-scored frequency of the
// rarer of (user, item).  Class shape of modelDropoutSigmoid.h:18-157.
#ifndef _MODEL_DROPOUT_SIGMOID_H_
#define _MODEL_DROPOUT_SIGMOID_H_

#include "model.h"

class ModelDropoutSigmoid : public Model {
 public:
  std::vector<double> userRankMap;  // percentile ranks (main.cpp:1187-1201); not used by the arithmetic
  std::vector<double> itemRankMap;
  std::vector<double> userFreq;
  std::vector<double> itemFreq;
  std::vector<double> factorial;
  std::vector<double> fDimWt;
  std::vector<int> cdfRanks;
  double minFreq;
  double maxFreq;
  double meanFreq;
  double stdFreq;

  ModelDropoutSigmoid(const Params &params, std::vector<double> &userRankMap, std::vector<double> &itemRankMap,
                      std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(params), userRankMap(userRankMap), itemRankMap(itemRankMap), userFreq(userFreq), itemFreq(itemFreq) {
    initFreqStats(params.facDim);
  }
  ModelDropoutSigmoid(const Params &params, int seed, std::vector<double> &userRankMap,
                      std::vector<double> &itemRankMap, std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(params, seed), userRankMap(userRankMap), itemRankMap(itemRankMap), userFreq(userFreq),
        itemFreq(itemFreq) {
    initFreqStats(params.facDim);
  }

  double estRating(int user, int item) override;
  void train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
             std::unordered_set<int> &invalidItems) override;

 protected:
  // ceil(facDim * sigmoid(rhoRMS * ((freq - mean)/std - alpha))), the expression of
  // modelDropoutSigmoid.cpp:158-163 evaluated for one side's frequency
  int sigmoidRank(double freq) const;
  void initFreqStats(int facDim);
  int deviceVariant() const override;
  void uploadAux(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                 std::unordered_set<int> &invalidItems) override;
};

#endif
