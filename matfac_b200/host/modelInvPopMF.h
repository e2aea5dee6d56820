// ModelInvPopMF (IFWMF): inverse-frequency-weighted MF.  Class shape of modelInvPopMF.h:17-52.
#ifndef _MODEL_INV_POP_MF_H_
#define _MODEL_INV_POP_MF_H_

#include <map>

#include "model.h"

class ModelInvPopMF : public Model {
 public:
  std::map<int, double> invPopU;
  std::map<int, double> invPopI;
  std::vector<double> userFreq;
  std::vector<double> itemFreq;
  int nTrainUsers;
  int nTrainItems;

  ModelInvPopMF(int nUsers, int nItems, int facDim, std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(nUsers, nItems, facDim), userFreq(userFreq), itemFreq(itemFreq), nTrainUsers(0), nTrainItems(0) {}
  ModelInvPopMF(const Params &params, std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(params), userFreq(userFreq), itemFreq(itemFreq), nTrainUsers(0), nTrainItems(0) {}
  ModelInvPopMF(const Params &params, int seed, std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(params, seed), userFreq(userFreq), itemFreq(itemFreq), nTrainUsers(0), nTrainItems(0) {}
  ModelInvPopMF(const Params &params, const char *uFacName, const char *iFacName, int seed,
                std::vector<double> &userFreq, std::vector<double> &itemFreq)
      : Model(params, uFacName, iFacName, seed), userFreq(userFreq), itemFreq(itemFreq), nTrainUsers(0),
        nTrainItems(0) {}

  void train(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
             std::unordered_set<int> &invalidItems) override;
  void trainSGDPar(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                   std::unordered_set<int> &invalidItems) override;
  // weighted objective (modelInvPopMF.cpp:3-55) — the base implementation dispatches on
  // deviceVariant(), this override only exists to keep the reference's signature visible
  double objective(const Data &data, std::unordered_set<int> &invalidUsers,
                   std::unordered_set<int> &invalidItems) override;

 protected:
  int deviceVariant() const override;
  void uploadAux(matfac::DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                 std::unordered_set<int> &invalidItems) override;
};

#endif
