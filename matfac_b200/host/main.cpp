// `mf` — command-line driver with the reference's flags, defaults and dispatch table
// (main.cpp:26-46 flags, :48-73 validation, :1233-1382 main).  gflags is not available here, so a
// small parser accepts the same `--flag value` / `--flag=value` / `-flag value` spellings.
//
// Extra, engine-only switches (default to reference behaviour):
//   --dump DIR      binary dumps of the CSR arrays, initial / last / best factors, invalid sets
//                   and a result.txt (the layout the test suite reads back), used by the
//                   parity tests
//   --dry_run 1     stop before training after dumping the host-side plan (initial factors,
//                   invalid sets, stratum partitions, the first schedules, per-id aux values);
//                   needs no GPU
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>

#include "datastruct.h"
#include "device_session.h"
#include "io.h"
#include "modelDropoutSigmoid.h"
#include "modelInvPopMF.h"
#include "modelMF.h"
#include "modelPoissonDropout.h"
#include "util.h"

namespace {

struct Flags {
  std::map<std::string, std::string> kv;
  std::string get(const char *name, const char *dflt) const {
    auto it = kv.find(name);
    return it == kv.end() ? std::string(dflt) : it->second;
  }
};

Flags parseFlags(int argc, char **argv) {
  Flags f;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    size_t dash = 0;
    while (dash < a.size() && a[dash] == '-') dash++;
    if (dash == 0) continue;
    a = a.substr(dash);
    size_t eq = a.find('=');
    if (eq != std::string::npos) f.kv[a.substr(0, eq)] = a.substr(eq + 1);
    else if (i + 1 < argc) f.kv[a] = argv[++i];
    else f.kv[a] = "true";
  }
  return f;
}

void dumpMat(Eigen::MatrixXf &m, int nrows, int ncols, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { std::cerr << "cannot write " << path << std::endl; exit(-1); }
  int32_t hdr[2] = {nrows, ncols};
  fwrite(hdr, sizeof(int32_t), 2, fp);
  for (int i = 0; i < nrows; i++) fwrite(&m(i, 0), sizeof(float), ncols, fp);
  fclose(fp);
}

template <typename T>
void dumpVec(const std::vector<T> &v, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { std::cerr << "cannot write " << path << std::endl; exit(-1); }
  int64_t n = (int64_t)v.size();
  fwrite(&n, sizeof(int64_t), 1, fp);
  fwrite(v.data(), sizeof(T), v.size(), fp);
  fclose(fp);
}

void dumpSet(std::unordered_set<int> &s, const std::string &path) {
  std::vector<int32_t> v(s.begin(), s.end());
  std::sort(v.begin(), v.end());
  dumpVec(v, path);
}

void dumpCsr(gk_csr_t *mat, const std::string &path) {
  FILE *fp = fopen(path.c_str(), "wb");
  if (!fp) { std::cerr << "cannot write " << path << std::endl; exit(-1); }
  int64_t nnz = mat->rowptr[mat->nrows];
  int64_t hdr[3] = {mat->nrows, mat->ncols, nnz};
  fwrite(hdr, sizeof(int64_t), 3, fp);
  fwrite(mat->rowptr, sizeof(int64_t), (size_t)mat->nrows + 1, fp);
  fwrite(mat->rowind, sizeof(int32_t), nnz, fp);
  fwrite(mat->rowval, sizeof(float), nnz, fp);
  fwrite(mat->colptr, sizeof(int64_t), (size_t)mat->ncols + 1, fp);
  fwrite(mat->colind, sizeof(int32_t), nnz, fp);
  fwrite(mat->colval, sizeof(float), nnz, fp);
  fclose(fp);
}

// percentile rank of every id by frequency (main.cpp:1170-1201)
std::vector<double> percentileRanks(const std::vector<double> &freq) {
  std::vector<std::pair<int, double>> pairs;
  for (int i = 0; i < (int)freq.size(); i++) pairs.push_back(std::make_pair(i, freq[i]));
  std::sort(pairs.begin(), pairs.end(),
            [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.second > b.second; });
  std::vector<double> rank(freq.size(), 0);
  for (int i = 0; i < (int)freq.size(); i++) rank[pairs[i].first] = double(freq.size() - i) / double(freq.size());
  return rank;
}

// Frequency quartiles of main.cpp:1109-1169 (setAdapRank / getUserItemRankMap): ids sorted by decreasing
// training frequency with the same std::sort + descComp call, cut into four parts of 25 % of the ids.
typedef std::vector<std::pair<int, std::vector<int>>> Parts;

static bool descCompPair(const std::pair<int, double> a, const std::pair<int, double> b) { return a.second > b.second; }

static Parts frequencyQuartiles(const std::vector<double> &freq) {
  std::vector<std::pair<int, double>> pairs;
  for (int i = 0; i < (int)freq.size(); i++) pairs.push_back(std::make_pair(i, freq[i]));
  std::sort(pairs.begin(), pairs.end(), descCompPair);
  Parts parts;
  const int n = (int)pairs.size();
  int i = 0, partInd = 0;
  while (i < n) {
    int end = i + 0.25 * ((float)n);
    if (end > n || partInd == 3) end = n;
    std::vector<int> ids;
    for (int k = i; k < end; k++) ids.push_back(pairs[k].first);
    parts.push_back(std::make_pair(partInd, ids));
    i = end;
    partInd++;
  }
  return parts;
}

// itemPartition.txt / userPartition.txt (main.cpp:1091-1107, :1412-1413)
static void writePartition(const Parts &parts, const std::unordered_set<int> &invalid, const char *name) {
  std::ofstream f(name);
  if (!f.is_open()) return;
  for (auto &part : parts)
    for (int id : part.second)
      if (invalid.count(id) == 0) f << part.first << " " << id << std::endl;
}

// quartileRMSEs (main.cpp:700-768): same lines on stdout; the eight filtered passes per matrix are one
// grouped device pass (Model::groupSE)
static void quartileRMSEs(Model &bestModel, const Data &data, const Parts &partItems, const Parts &partUsers,
                          std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems) {
  std::vector<uint8_t> ug(data.nUsers, 255), ig(data.nItems, 255);
  for (auto &p : partUsers)
    for (int u : p.second)
      if (u < data.nUsers) ug[u] = (uint8_t)p.first;
  for (auto &p : partItems)
    for (int i : p.second)
      if (i < data.nItems) ig[i] = (uint8_t)p.first;
  std::cout << std::endl;
  std::cout << "Train RMSE: " << bestModel.RMSE(data.trainMat, invalidUsers, invalidItems) << std::endl;
  std::cout << "Test RMSE: " << bestModel.RMSE(data.testMat, invalidUsers, invalidItems) << std::endl;
  std::cout << "Val RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
  const char *names[2] = {"Test RMSE: ", "Validation RMSE: "};
  gk_csr_t *mats[2] = {data.testMat, data.valMat};
  for (int m = 0; m < 2; m++) {
    double out[32];
    bestModel.groupSE(mats[m], ug, ig, invalidUsers, invalidItems, out);
    std::cout << names[m] << std::endl;
    std::cout << "Items Part: ";
    for (auto &p : partItems) std::cout << (int)out[(0 * 8 + p.first) * 2 + 1] << " "
                                        << sqrt(out[(0 * 8 + p.first) * 2] / out[(0 * 8 + p.first) * 2 + 1]) << " ";
    std::cout << std::endl;
    std::cout << "Users Part: ";
    for (auto &p : partUsers) std::cout << (int)out[(1 * 8 + p.first) * 2 + 1] << " "
                                        << sqrt(out[(1 * 8 + p.first) * 2] / out[(1 * 8 + p.first) * 2 + 1]) << " ";
    std::cout << std::endl;
  }
}

}  // namespace

int main(int argc, char **argv) {
  Flags fl = parseFlags(argc, argv);
  std::string trainmat = fl.get("trainmat", ""), testmat = fl.get("testmat", ""), valmat = fl.get("valmat", ""),
              prefix = fl.get("prefix", ""), graphmat = fl.get("graphmat", ""), origufac = fl.get("origufac", ""),
              origifac = fl.get("origifac", ""), initufac = fl.get("initufac", ""), initifac = fl.get("initifac", "");
  const std::string mf_method = fl.get("mf_method", "sgd"), algo = fl.get("algo", "mf");
  const std::string dumpDir = fl.get("dump", "");
  const bool dryRun = fl.get("dry_run", "0") != "0";
  if (trainmat.empty() || testmat.empty() || valmat.empty()) {
    std::cerr << "Missing either train, test or val matrix" << std::endl;
    exit(-1);
  }
  if (prefix.empty()) {
    std::cerr << "Missing model prefix" << std::endl;
    exit(-1);
  }
  Params params(std::stoi(fl.get("facdim", "5")), std::stoi(fl.get("maxiter", "5000")),
                std::stoi(fl.get("svdfacdim", "5")), std::stoi(fl.get("seed", "1")), std::stod(fl.get("ureg", "0.01")),
                std::stod(fl.get("ireg", "0.01")), std::stod(fl.get("learnrate", "0.005")),
                std::stod(fl.get("rhorms", "0.0")), std::stod(fl.get("alpha", "0.0")), trainmat, testmat, valmat,
                graphmat, origufac, origifac, initufac, initifac, prefix);
  Data data(params);
  params.nUsers = data.nUsers;
  params.nItems = data.nItems;
  params.display();
  std::srand(params.seed);
  std::cout << "train sorted by item: " << checkIfUISorted(data.trainMat) << std::endl;

  auto rowColFreq = getRowColFreq(data.trainMat);
  std::vector<double> userFreq = rowColFreq.first, itemFreq = rowColFreq.second;
  std::vector<double> userRankPc = percentileRanks(userFreq), itemRankPc = percentileRanks(itemFreq);

  std::unique_ptr<Model> mfModel, bestModel;
  std::unordered_set<int> invalidUsers, invalidItems;
  if (algo == "mf") {
    mfModel.reset(new ModelMF(params, params.seed));
    bestModel.reset(new ModelMF(params, params.seed));
  } else if (algo == "TMF") {
    mfModel.reset(new ModelDropoutSigmoid(params, params.seed, userRankPc, itemRankPc, userFreq, itemFreq));
    bestModel.reset(new ModelDropoutSigmoid(params, params.seed, userRankPc, itemRankPc, userFreq, itemFreq));
  } else if (algo == "TMFDropout") {
    mfModel.reset(new ModelPoissonDropout(params, params.seed, userRankPc, itemRankPc, userFreq, itemFreq));
    bestModel.reset(new ModelPoissonDropout(params, params.seed, userRankPc, itemRankPc, userFreq, itemFreq));
  } else if (algo == "IFWMF") {
    mfModel.reset(new ModelInvPopMF(params, params.seed, userFreq, itemFreq));
    bestModel.reset(new ModelInvPopMF(params, params.seed, userFreq, itemFreq));
  } else {
    std::cerr << "Invalid algo input: " << algo << std::endl;
    return 0;
  }

  if (!dumpDir.empty()) {
    dumpCsr(data.trainMat, dumpDir + "/train.csr.bin");
    dumpCsr(data.valMat, dumpDir + "/val.csr.bin");
    dumpCsr(data.testMat, dumpDir + "/test.csr.bin");
    dumpMat(mfModel->uFac, mfModel->nUsers, mfModel->facDim, dumpDir + "/init_uFac.bin");
    dumpMat(mfModel->iFac, mfModel->nItems, mfModel->facDim, dumpDir + "/init_iFac.bin");
  }

  if (dryRun) {
    // host-side plan only: what the stratified trainers would hand to the engine
    std::vector<std::unordered_set<int>> none;
    getInvalidUsersItems(data.trainMat, none, invalidUsers, invalidItems);
    for (int u = data.trainMat->nrows; u < data.nUsers; u++) invalidUsers.insert(u);
    for (int i = data.trainMat->ncols; i < data.nItems; i++) invalidItems.insert(i);
    std::vector<int> trainUsers = matfac::validIds(data.trainMat->nrows, invalidUsers);
    std::vector<int> trainItems = matfac::validIds(data.trainMat->ncols, invalidItems);
    std::mt19937 mt(params.seed);
    std::shuffle(trainUsers.begin(), trainUsers.end(), mt);
    std::shuffle(trainItems.begin(), trainItems.end(), mt);
    const int P = omp_get_max_threads();
    std::vector<int> up = matfac::partitionIds(trainUsers, P, data.nUsers), ip = matfac::partitionIds(trainItems, P, data.nItems);
    const int nSched = std::stoi(fl.get("dry_schedules", "24"));
    std::vector<int32_t> sched;
    std::vector<std::pair<int, int>> seq;
    for (int s = 0; s < nSched; s++) {
      sgdUpdateBlockSeq(P, seq, mt);
      for (auto &pr : seq) { sched.push_back(pr.first); sched.push_back(pr.second); }
    }
    if (!dumpDir.empty()) {
      dumpSet(invalidUsers, dumpDir + "/invalidUsers.bin");
      dumpSet(invalidItems, dumpDir + "/invalidItems.bin");
      dumpVec(up, dumpDir + "/user_part.bin");
      dumpVec(ip, dumpDir + "/item_part.bin");
      dumpVec(sched, dumpDir + "/schedule.bin");
      std::vector<int32_t> dimOrder;
      std::mt19937 mt2(params.seed);
      std::vector<int> dims(params.facDim);
      std::iota(dims.begin(), dims.end(), 0);
      for (int e = 0; e < 3; e++) {
        std::shuffle(dims.begin(), dims.end(), mt2);
        dimOrder.insert(dimOrder.end(), dims.begin(), dims.end());
      }
      dumpVec(dimOrder, dumpDir + "/ccdpp_dims.bin");
    }
    std::cout << "dry run: P = " << P << std::endl;
    return 0;
  }

  // algo x mf_method dispatch of main.cpp:1325-1370 (ccd++ runs the frequency-adaptive trainer;
  // TMF / TMFDropout / IFWMF always call train()).  "sgdpar" with IFWMF and "ccdpp_plain" reach
  // the two trainers that the reference only exposes through its C++ API.
  if (algo == "mf") {
    if (mf_method == "ccd++") mfModel->trainCCDPPFreqAdap(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "ccdpp_plain") mfModel->trainCCDPP(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "als") mfModel->trainALS(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "hogsgd") mfModel->hogTrain(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "sgdpar") mfModel->trainSGDPar(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "sgdu") mfModel->trainUShuffle(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "ccd") mfModel->trainCCD(data, *bestModel, invalidUsers, invalidItems);
    else if (mf_method == "sgdparsvd") {
      std::cerr << "--mf_method " << mf_method << " is not provided by the GPU engine" << std::endl;
      return -1;
    } else mfModel->train(data, *bestModel, invalidUsers, invalidItems);
  } else if (algo == "IFWMF" && mf_method == "sgdpar") {
    mfModel->trainSGDPar(data, *bestModel, invalidUsers, invalidItems);
  } else {
    mfModel->train(data, *bestModel, invalidUsers, invalidItems);
  }

  const double trainRMSE = bestModel->RMSE(data.trainMat, invalidUsers, invalidItems);
  const double testRMSE = bestModel->RMSE(data.testMat, invalidUsers, invalidItems);
  const double valRMSE = bestModel->RMSE(data.valMat, invalidUsers, invalidItems);
  std::cout << "\nTrain RMSE: " << trainRMSE;
  std::cout << "\nTest RMSE: " << testRMSE;
  std::cout << "\nValidation RMSE: " << valRMSE << std::endl;
  mfModel->display();
  std::cout << std::endl;

  // tail / head report of main.cpp:1407-1413
  std::cout << "invalid users: " << invalidUsers.size() << " invalid items: " << invalidItems.size() << std::endl;
  const Parts partItems = frequencyQuartiles(itemFreq), partUsers = frequencyQuartiles(userFreq);
  quartileRMSEs(*bestModel, data, partItems, partUsers, invalidUsers, invalidItems);
  writePartition(partItems, invalidItems, "itemPartition.txt");
  writePartition(partUsers, invalidUsers, "userPartition.txt");

  if (!dumpDir.empty()) {
    dumpMat(mfModel->uFac, mfModel->nUsers, mfModel->facDim, dumpDir + "/last_uFac.bin");
    dumpMat(mfModel->iFac, mfModel->nItems, mfModel->facDim, dumpDir + "/last_iFac.bin");
    dumpMat(bestModel->uFac, bestModel->nUsers, bestModel->facDim, dumpDir + "/best_uFac.bin");
    dumpMat(bestModel->iFac, bestModel->nItems, bestModel->facDim, dumpDir + "/best_iFac.bin");
    dumpSet(invalidUsers, dumpDir + "/invalidUsers.bin");
    dumpSet(invalidItems, dumpDir + "/invalidItems.bin");
    FILE *fp = fopen((dumpDir + "/result.txt").c_str(), "w");
    fprintf(fp, "best_train_rmse %.17g\nbest_test_rmse %.17g\nbest_val_rmse %.17g\n", trainRMSE, testRMSE, valRMSE);
    fprintf(fp, "last_val_rmse %.17g\n", mfModel->RMSE(data.valMat, invalidUsers, invalidItems));
    fprintf(fp, "last_test_rmse %.17g\n", mfModel->RMSE(data.testMat, invalidUsers, invalidItems));
    fprintf(fp, "last_objective %.17g\n", mfModel->objective(data, invalidUsers, invalidItems));
    fprintf(fp, "learn_rate %.9g\n", (double)mfModel->learnRate);
    fprintf(fp, "signature %s\n", bestModel->modelSignature().c_str());
    fclose(fp);
  }
  matfac::DeviceSession::dropAll();
  return 0;
}
