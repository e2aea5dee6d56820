// ModelMF: plain matrix factorisation (squared loss + L2).  Same class, constructors and trainer
// names as the reference's modelMF.h:21-57; the trainers run on the GPU.
#ifndef _MODEL_MF_H_
#define _MODEL_MF_H_

#include "model.h"

class ModelMF : public Model {
 public:
  ModelMF(int nUsers, int nItems, int facDim) : Model(nUsers, nItems, facDim) {}
  ModelMF(const Params &params) : Model(params) {}
  ModelMF(const Params &params, int seed) : Model(params, seed) {}
  ModelMF(const Params &params, const char *uFacName, const char *iFacName, int seed)
      : Model(params, uFacName, iFacName, seed) {}

#define MATFAC_DECL(name)                                                                         \
  void name(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,           \
            std::unordered_set<int> &invalidItems) override;
  MATFAC_DECL(train)               // serial SGD            modelMF.cpp:4
  MATFAC_DECL(trainSGDPar)         // stratified SGD        modelMF.cpp:154
  MATFAC_DECL(hogTrain)            // Hogwild SGD           modelMF.cpp:1656
  MATFAC_DECL(trainUShuffle)       // user-major SGD        modelMF.cpp:560
  MATFAC_DECL(trainALS)            // alternating LS        modelMF.cpp:709
  MATFAC_DECL(trainCCDPP)          // CCD++                 modelMF.cpp:931
  MATFAC_DECL(trainCCDPPFreqAdap)  // CCD++, freq-adaptive  modelMF.cpp:1172
  MATFAC_DECL(trainCCD)            // CCD, row by row       modelMF.cpp:1426
#undef MATFAC_DECL

 private:
  void runCcdpp(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                std::unordered_set<int> &invalidItems, bool freqAdap, const char *tag);
};

#endif
