// TEST INFRASTRUCTURE — command-line driver for the UNMODIFIED reference translation units
// (model.cpp modelMF.cpp modelInvPopMF.cpp modelDropoutSigmoid.cpp modelPoissonDropout.cpp
// util.cpp io.cpp datastruct.cpp, compiled where they lie under /root/reference against the
// shims in oracle/shim/).  It stands in for the reference's main.cpp, which cannot be
// compiled here because it drags in five analysis headers plus gflags; the statements below
// follow main.cpp:1233-1382 (Params/Data construction, frequency vectors, the algo x
// mf_method dispatch table with its "ccd++ -> trainCCDPPFreqAdap" quirk, and the final RMSE
// prints) and add binary dumps so the CPU restatement in oracle/mf_oracle.cpp and the CUDA
// engine can be checked against the reference's own arithmetic.
//
// Built into oracle/_ref/mf_ref by oracle/Makefile.  Never shipped, never on the product path.
#include "datastruct.h"
#include "io.h"
#include "modelDropoutSigmoid.h"
#include "modelInvPopMF.h"
#include "modelMF.h"
#include "modelPoissonDropout.h"
#include "util.h"

#include <cstdint>
#include <cstring>
#include <map>
#include <memory>

static std::map<std::string, std::string> parseFlags(int argc, char **argv) {
  std::map<std::string, std::string> kv;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    while (!a.empty() && a[0] == '-') a.erase(0, 1);
    size_t eq = a.find('=');
    if (eq != std::string::npos) {
      kv[a.substr(0, eq)] = a.substr(eq + 1);
    } else if (i + 1 < argc) {
      kv[a] = argv[++i];
    }
  }
  return kv;
}

static std::string flag(std::map<std::string, std::string> &kv, const char *name,
                        const char *dflt) {
  auto it = kv.find(name);
  return it == kv.end() ? std::string(dflt) : it->second;
}

static void dumpMat(Eigen::MatrixXf &m, int nrows, int ncols, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { std::cerr << "cannot write " << path << std::endl; exit(-1); }
  int32_t hdr[2] = {nrows, ncols};
  fwrite(hdr, sizeof(int32_t), 2, fp);
  std::vector<float> row(ncols);
  for (int i = 0; i < nrows; i++) {
    for (int j = 0; j < ncols; j++) row[j] = m(i, j);
    fwrite(row.data(), sizeof(float), ncols, fp);
  }
  fclose(fp);
}

static void dumpCsr(gk_csr_t *mat, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { std::cerr << "cannot write " << path << std::endl; exit(-1); }
  int64_t nnz = mat->rowptr[mat->nrows];
  int64_t hdr[3] = {mat->nrows, mat->ncols, nnz};
  fwrite(hdr, sizeof(int64_t), 3, fp);
  std::vector<int64_t> ptr(mat->rowptr, mat->rowptr + mat->nrows + 1);
  fwrite(ptr.data(), sizeof(int64_t), ptr.size(), fp);
  fwrite(mat->rowind, sizeof(int32_t), nnz, fp);
  fwrite(mat->rowval, sizeof(float), nnz, fp);
  ptr.assign(mat->colptr, mat->colptr + mat->ncols + 1);
  fwrite(ptr.data(), sizeof(int64_t), ptr.size(), fp);
  fwrite(mat->colind, sizeof(int32_t), nnz, fp);
  fwrite(mat->colval, sizeof(float), nnz, fp);
  fclose(fp);
}

static void dumpSet(std::unordered_set<int> &s, const std::string &path) {
  std::vector<int32_t> v(s.begin(), s.end());
  std::sort(v.begin(), v.end());
  FILE *fp = fopen(path.c_str(), "wb");
  int64_t n = v.size();
  fwrite(&n, sizeof(int64_t), 1, fp);
  fwrite(v.data(), sizeof(int32_t), v.size(), fp);
  fclose(fp);
}

// main.cpp:1170-1201 (percentile rank maps; dead inputs of the TMF constructors)
static void setPcLocal(std::vector<double> &indFreq, std::vector<double> &indRank) {
  std::vector<std::pair<int, double>> pairs;
  for (int i = 0; i < (int)indFreq.size(); i++) pairs.push_back(std::make_pair(i, indFreq[i]));
  std::sort(pairs.begin(), pairs.end(), descComp);
  for (int i = 0; i < (int)indFreq.size(); i++) {
    indRank[pairs[i].first] = double(indFreq.size() - i) / double(indFreq.size());
  }
}

int main(int argc, char **argv) {
  auto kv = parseFlags(argc, argv);
  std::string trainmat = flag(kv, "trainmat", ""), testmat = flag(kv, "testmat", ""),
              valmat = flag(kv, "valmat", ""), prefix = flag(kv, "prefix", ""),
              graphmat, origufac, origifac, initufac, initifac;
  std::string mf_method = flag(kv, "mf_method", "sgd"), algo = flag(kv, "algo", "mf");
  std::string dumpDir = flag(kv, "dump", "");
  if (trainmat.empty() || testmat.empty() || valmat.empty() || prefix.empty()) {
    std::cerr << "Missing train/test/val matrix or prefix" << std::endl;
    return -1;
  }
  // defaults of main.cpp:26-34
  Params params(std::stoi(flag(kv, "facdim", "5")), std::stoi(flag(kv, "maxiter", "5000")),
                std::stoi(flag(kv, "svdfacdim", "5")), std::stoi(flag(kv, "seed", "1")),
                std::stod(flag(kv, "ureg", "0.01")), std::stod(flag(kv, "ireg", "0.01")),
                std::stod(flag(kv, "learnrate", "0.005")), std::stod(flag(kv, "rhorms", "0.0")),
                std::stod(flag(kv, "alpha", "0.0")), trainmat, testmat, valmat, graphmat,
                origufac, origifac, initufac, initifac, prefix);

  Data data(params);
  params.nUsers = data.nUsers;
  params.nItems = data.nItems;
  params.display();
  std::srand(params.seed);

  auto rowColFreq = getRowColFreq(data.trainMat);
  auto userFreq = rowColFreq.first;
  auto itemFreq = rowColFreq.second;
  std::vector<double> userRankPc(userFreq.size(), 0), itemRankPc(itemFreq.size(), 0);
  setPcLocal(userFreq, userRankPc);
  setPcLocal(itemFreq, itemRankPc);

  std::unique_ptr<Model> mfModel, bestModel;
  std::unordered_set<int> invalidUsers, invalidItems;

  if (!dumpDir.empty()) {
    dumpCsr(data.trainMat, dumpDir + "/train.csr.bin");
    dumpCsr(data.valMat, dumpDir + "/val.csr.bin");
    dumpCsr(data.testMat, dumpDir + "/test.csr.bin");
  }

  if (algo == "mf") {
    mfModel = std::make_unique<ModelMF>(params, params.seed);
    bestModel = std::make_unique<ModelMF>(params, params.seed);
  } else if (algo == "TMF") {
    mfModel = std::make_unique<ModelDropoutSigmoid>(params, params.seed, userRankPc, itemRankPc,
                                                    userFreq, itemFreq);
    bestModel = std::make_unique<ModelDropoutSigmoid>(params, params.seed, userRankPc,
                                                      itemRankPc, userFreq, itemFreq);
  } else if (algo == "TMFDropout") {
    mfModel = std::make_unique<ModelPoissonDropout>(params, params.seed, userRankPc, itemRankPc,
                                                    userFreq, itemFreq);
    bestModel = std::make_unique<ModelPoissonDropout>(params, params.seed, userRankPc,
                                                      itemRankPc, userFreq, itemFreq);
  } else if (algo == "IFWMF") {
    mfModel = std::make_unique<ModelInvPopMF>(params, params.seed, userFreq, itemFreq);
    bestModel = std::make_unique<ModelInvPopMF>(params, params.seed, userFreq, itemFreq);
  } else {
    std::cerr << "Invalid algo input: " << algo << std::endl;
    return 0;
  }
  if (!dumpDir.empty()) {
    dumpMat(mfModel->uFac, mfModel->nUsers, mfModel->facDim, dumpDir + "/init_uFac.bin");
    dumpMat(mfModel->iFac, mfModel->nItems, mfModel->facDim, dumpDir + "/init_iFac.bin");
  }

  // dispatch table of main.cpp:1325-1370; "sgdpar_ifw" additionally reaches the API-only
  // ModelInvPopMF::trainSGDPar, "ccdpp_plain" the API-only ModelMF::trainCCDPP.
  if (algo == "mf") {
    if (mf_method == "ccd++") {
      mfModel->trainCCDPPFreqAdap(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "ccdpp_plain") {
      mfModel->trainCCDPP(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "ccd") {
      mfModel->trainCCD(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "als") {
      mfModel->trainALS(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "hogsgd") {
      mfModel->hogTrain(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "sgdu") {
      mfModel->trainUShuffle(data, *bestModel, invalidUsers, invalidItems);
    } else if (mf_method == "sgdpar") {
      mfModel->trainSGDPar(data, *bestModel, invalidUsers, invalidItems);
    } else {
      mfModel->train(data, *bestModel, invalidUsers, invalidItems);
    }
  } else if (algo == "IFWMF" && mf_method == "sgdpar") {
    mfModel->trainSGDPar(data, *bestModel, invalidUsers, invalidItems);
  } else {
    mfModel->train(data, *bestModel, invalidUsers, invalidItems);
  }

  double trainRMSE = bestModel->RMSE(data.trainMat, invalidUsers, invalidItems);
  double testRMSE = bestModel->RMSE(data.testMat, invalidUsers, invalidItems);
  double valRMSE = bestModel->RMSE(data.valMat, invalidUsers, invalidItems);
  std::cout << "\nTrain RMSE: " << trainRMSE;
  std::cout << "\nTest RMSE: " << testRMSE;
  std::cout << "\nValidation RMSE: " << valRMSE << std::endl;

  if (!dumpDir.empty()) {
    dumpMat(mfModel->uFac, mfModel->nUsers, mfModel->facDim, dumpDir + "/last_uFac.bin");
    dumpMat(mfModel->iFac, mfModel->nItems, mfModel->facDim, dumpDir + "/last_iFac.bin");
    dumpMat(bestModel->uFac, bestModel->nUsers, bestModel->facDim, dumpDir + "/best_uFac.bin");
    dumpMat(bestModel->iFac, bestModel->nItems, bestModel->facDim, dumpDir + "/best_iFac.bin");
    dumpSet(invalidUsers, dumpDir + "/invalidUsers.bin");
    dumpSet(invalidItems, dumpDir + "/invalidItems.bin");
    FILE *fp = fopen((dumpDir + "/result.txt").c_str(), "w");
    fprintf(fp, "best_train_rmse %.17g\nbest_test_rmse %.17g\nbest_val_rmse %.17g\n", trainRMSE,
            testRMSE, valRMSE);
    fprintf(fp, "last_val_rmse %.17g\n", mfModel->RMSE(data.valMat, invalidUsers, invalidItems));
    fprintf(fp, "last_test_rmse %.17g\n", mfModel->RMSE(data.testMat, invalidUsers, invalidItems));
    fprintf(fp, "last_objective %.17g\n", mfModel->objective(data, invalidUsers, invalidItems));
    fprintf(fp, "learn_rate %.9g\n", (double)mfModel->learnRate);
    if (flag(kv, "rank_metrics", "0") == "1") {
      // the reference's own ranking metrics (model.cpp:760-1332) of the best model; filters: users u % 3 == 0, items i % 2 == 0
      std::unordered_set<int> fu, fi;
      for (int u = 0; u < data.nUsers; u += 3) fu.insert(u);
      for (int i = 0; i < data.nItems; i += 2) fi.insert(i);
      const char *names[2] = {"val", "test"};
      gk_csr_t *mats[2] = {data.valMat, data.testMat};
      for (int w = 0; w < 2; w++) {
        fprintf(fp, "%s_hr %.17g\n", names[w], bestModel->hitRate(data, invalidUsers, invalidItems, mats[w]));
        fprintf(fp, "%s_arhr %.17g\n", names[w], bestModel->arHR(data, invalidUsers, invalidItems, mats[w]));
        fprintf(fp, "%s_ndcg %.17g\n", names[w], bestModel->NDCG(invalidUsers, invalidItems, mats[w]));
        auto a = bestModel->hitRateU(data, fu, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_hru_first %.17g\n%s_hru %.17g\n", names[w], (double)a.first, names[w], a.second);
        auto b = bestModel->arHRU(data, fu, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_arhru_first %.17g\n%s_arhru %.17g\n", names[w], b.first, names[w], b.second);
        auto c = bestModel->NDCGU(fu, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_ndcgu_first %.17g\n%s_ndcgu %.17g\n", names[w], (double)c.first, names[w], c.second);
        auto d2 = bestModel->hitRateI(data, fi, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_hri_first %.17g\n%s_hri %.17g\n", names[w], (double)d2.first, names[w], d2.second);
        auto e2 = bestModel->arHRI(data, fi, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_arhri_first %.17g\n%s_arhri %.17g\n", names[w], e2.first, names[w], e2.second);
        auto f2 = bestModel->NDCGI(fi, invalidUsers, invalidItems, mats[w]);
        fprintf(fp, "%s_ndcgi_first %.17g\n%s_ndcgi %.17g\n", names[w], (double)f2.first, names[w], f2.second);
      }
    }
    fprintf(fp, "signature %s\n", bestModel->modelSignature().c_str());
    fclose(fp);
  }
  return 0;
}
