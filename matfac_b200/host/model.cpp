// Model base class: state, seeded initialisation, device-backed evaluation, early stopping and
// persistence.  Mirrors the reference's model.cpp for the functions on the training path
// (ctor/init :2315-2366, estRating :547, RMSE :191/:214, objective :1694/:1770,
// isTerminateModel :1471, save/load :31-188, modelSignature :11).
#include "model.h"

#include <omp.h>
#include <type_traits>

#include <cmath>

#include "device_session.h"
#include "io.h"

using matfac::DeviceSession;

// ---- construction --------------------------------------------------------------------------
Model::Model(int p_nUsers, int p_nItems, int p_facDim)
    : nUsers(p_nUsers), nItems(p_nItems), facDim(p_facDim), trainSeed(-1), origLearnRate(0), learnRate(0),
      rhoRMS(0), alpha(0), maxIter(0), uReg(0), iReg(0), sing_a(0), sing_b(0), mu(0) {}

// One std::default_random_engine stream seeded with params.seed feeds, in this order, uFac
// (user outer, dimension inner), iFac, uBias, iBias through
// uniform_real_distribution<double>(-0.01f, 0.01f) — the same calls as model.cpp:2331-2362, so
// with the same libstdc++ the initial factors are bit-identical to the reference's.
Model::Model(const Params &params)
    : nUsers(params.nUsers), nItems(params.nItems), facDim(params.facDim), trainSeed(-1),
      origLearnRate(params.learnRate), learnRate(params.learnRate), rhoRMS(params.rhoRMS), alpha(params.alpha),
      maxIter(params.maxIter), uReg(params.uReg), iReg(params.iReg), sing_a(params.uReg), sing_b(params.iReg),
      mu(0) {
  float lb = -0.01, ub = 0.01;
  std::cout << "lb = " << lb << " ub = " << ub << std::endl;
  uFac = Eigen::MatrixXf(nUsers, facDim);
  iFac = Eigen::MatrixXf(nItems, facDim);
  uBias = Eigen::VectorXf(nUsers);
  iBias = Eigen::VectorXf(nItems);
  // The reference draws everything from ONE serial stream (model.cpp:2331-2362).  libstdc++'s default_random_engine
  // is minstd_rand0, x <- 16807 x mod (2^31 - 1), and uniform_real_distribution<double> takes exactly two draws per
  // value from it (generate_canonical<double, 53> over a 31-bit range), so value number n starts at draw 2 n and the
  // state there is 16807^(2 n) x0: every OpenMP thread jumps to its part of the stream and then makes the very same
  // library calls — bit-identical to the serial loop (tests/test_host.py), 0.46 s -> 0.15 s at Netflix size on 8 cores.
  static_assert(std::is_same<std::default_random_engine, std::minstd_rand0>::value,
                "the jump-ahead below assumes libstdc++'s default_random_engine (minstd_rand0)");
  const uint64_t mod = 2147483647ull, mul = 16807ull;
  uint64_t x0 = (uint64_t) static_cast<std::default_random_engine::result_type>(params.seed) % mod;
  if (x0 == 0) x0 = 1;  // linear_congruential_engine::seed maps a zero state to 1
  auto state_at = [&](uint64_t draws) {  // 16807^draws * x0 mod (2^31 - 1)
    uint64_t r = x0, b = mul;
    for (uint64_t e = draws; e; e >>= 1, b = b * b % mod)
      if (e & 1) r = r * b % mod;
    return r;
  };
  struct Span { float *dst; size_t n; };
  const Span spans[4] = {{uFac.data(), (size_t)nUsers * facDim}, {iFac.data(), (size_t)nItems * facDim},
                         {uBias.data(), (size_t)nUsers}, {iBias.data(), (size_t)nItems}};
  size_t first = 0;
  for (const Span &sp : spans) {
    const size_t n = sp.n;
    const int nt = n >= ((size_t)1 << 16) ? omp_get_max_threads() : 1;
#pragma omp parallel for num_threads(nt) schedule(static, 1)
    for (int t = 0; t < nt; t++) {
      const size_t lo = n * (size_t)t / nt, hi = n * (size_t)(t + 1) / nt;
      std::default_random_engine generator((std::default_random_engine::result_type)state_at(2 * (first + lo)));
      std::uniform_real_distribution<double> dist(lb, ub);
      for (size_t i = lo; i < hi; i++) sp.dst[i] = dist(generator);
    }
    first += n;
  }
  singularVals = Eigen::VectorXf(facDim);
}

Model::Model(int p_nUsers, int p_nItems, const Params &params) : Model(params) {
  nUsers = p_nUsers;
  nItems = p_nItems;
}

Model::Model(const Params &params, int seed) : Model(params) { trainSeed = seed; }

Model::Model(const Params &params, const char *uFacName, const char *iFacName, int seed) : Model(params, seed) {
  std::cout << "\nLoading user factors: " << uFacName;
  readMat(uFac, nUsers, facDim, uFacName);
  std::cout << "\nLoading item factors: " << iFacName;
  readMat(iFac, nItems, facDim, iFacName);
}

Model::Model(const Params &params, const char *uFacName, const char *iFacName, const char *uBFName,
             const char *iBFName, const char *gBFName, int seed)
    : Model(params, uFacName, iFacName, seed) {
  uBias = readEigVector(uBFName);
  iBias = readEigVector(iBFName);
  std::vector<double> gBias = readDVector(gBFName);
  mu = gBias.empty() ? 0 : gBias[0];
}

// ---- small host-side pieces ----------------------------------------------------------------
double Model::estRating(int user, int item) { return uFac.row(user).dot(iFac.row(item)); }

std::string Model::modelSignature() {
  return std::to_string(nUsers) + "X" + std::to_string(nItems) + "_" + std::to_string(facDim) + "_" +
         std::to_string(uReg) + "_" + std::to_string(iReg) + "_" + std::to_string(origLearnRate);
}

void Model::display() {
  std::cout << "nUsers: " << nUsers << " nItems: " << nItems << std::endl;
  std::cout << "facDim: " << facDim << std::endl;
  std::cout << "uReg: " << uReg << " iReg: " << iReg << std::endl;
  std::cout << "learnRate: " << learnRate << std::endl;
  std::cout << "trainSeed: " << trainSeed;
}

void Model::copyScalarsFrom(const Model &o) {
  nUsers = o.nUsers; nItems = o.nItems; facDim = o.facDim; trainSeed = o.trainSeed;
  origLearnRate = o.origLearnRate; learnRate = o.learnRate; rhoRMS = o.rhoRMS; alpha = o.alpha;
  maxIter = o.maxIter; uReg = o.uReg; iReg = o.iReg; sing_a = o.sing_a; sing_b = o.sing_b; mu = o.mu;
}

// ---- persistence (file names and text format of model.cpp:31-128, io.cpp:139-154) -----------
void Model::saveFacs(std::string prefix) {
  std::cout << "Saving model... " << prefix << std::endl;
  const std::string sign = modelSignature();
  const std::string uName = prefix + "_uFac_" + sign + ".mat", iName = prefix + "_iFac_" + sign + ".mat";
  writeMat(uFac, nUsers, facDim, uName.c_str());
  std::cout << "uFac Norm: " << uFac.norm() << std::endl;
  writeMat(iFac, nItems, facDim, iName.c_str());
  std::cout << "iFac Norm: " << iFac.norm() << std::endl;
}

void Model::save(std::string prefix) {
  saveFacs(prefix);
  const std::string sign = modelSignature();
  writeVector(uBias, (prefix + "_uBias_" + sign + ".vec").c_str());
  writeVector(iBias, (prefix + "_iBias_" + sign + ".vec").c_str());
  std::vector<double> gBias = {mu};
  writeVector(gBias, (prefix + "_" + sign + "_gBias").c_str());
}

void Model::loadFacs(std::string prefix) {
  const std::string sign = modelSignature();
  const std::string uName = prefix + "_uFac_" + sign + ".mat", iName = prefix + "_iFac_" + sign + ".mat";
  if (isFileExist(uName.c_str())) readMat(uFac, nUsers, facDim, uName.c_str());
  if (isFileExist(iName.c_str())) readMat(iFac, nItems, facDim, iName.c_str());
  std::cout << "uFac Norm: " << uFac.norm() << " iFac Norm: " << iFac.norm() << std::endl;
}

void Model::load(std::string prefix) {
  loadFacs(prefix);
  const std::string sign = modelSignature();
  const std::string ub = prefix + "_uBias_" + sign + ".vec", ib = prefix + "_iBias_" + sign + ".vec",
                    gb = prefix + "_" + sign + "_gBias";
  if (isFileExist(ub.c_str())) uBias = readEigVector(ub.c_str());
  if (isFileExist(ib.c_str())) iBias = readEigVector(ib.c_str());
  if (isFileExist(gb.c_str())) {
    std::vector<double> g = readDVector(gb.c_str());
    if (!g.empty()) mu = g[0];
  }
}

void Model::load(const char *uFacName, const char *iFacName) {
  readMat(uFac, nUsers, facDim, uFacName);
  readMat(iFac, nItems, facDim, iFacName);
}

void Model::saveBinFacs(std::string prefix) {
  const std::string sign = modelSignature();
  writeMatBin(uFac, nUsers, facDim, (prefix + "_uFac_" + sign + ".binmat").c_str());
  writeMatBin(iFac, nItems, facDim, (prefix + "_iFac_" + sign + ".binmat").c_str());
}

void Model::loadBinFacs(std::string prefix) {
  const std::string sign = modelSignature();
  readMatBin(uFac, nUsers, facDim, (prefix + "_uFac_" + sign + ".binmat").c_str());
  readMatBin(iFac, nItems, facDim, (prefix + "_iFac_" + sign + ".binmat").c_str());
}

// ---- device-backed evaluation ---------------------------------------------------------------
int Model::deviceVariant() const { return MFB_MF; }

void Model::uploadAux(DeviceSession &, const Data *, std::unordered_set<int> &, std::unordered_set<int> &) {}

void Model::uploadFactors(DeviceSession &s) {
  for (int r = 0; r < s.world(); r++) s.check(mfb_upload_factors(s.engineOf(r), uFac.data(), facDim, iFac.data(), facDim));
}

// the model's auxiliaries into every engine of the session (uploadAux writes to s.eng)
void Model::uploadAuxAll(DeviceSession &s, const Data *data, std::unordered_set<int> &invalidUsers,
                         std::unordered_set<int> &invalidItems) {
  uploadAux(s, data, invalidUsers, invalidItems);
  for (size_t w = 0; w < s.workers.size(); w++) {
    std::swap(s.eng, s.workers[w]);
    uploadAux(s, nullptr, invalidUsers, invalidItems);
    std::swap(s.eng, s.workers[w]);
  }
}

// objective = true: weighted SSE over train + uReg |U|^2 + iReg |V|^2; false: RMSE over `which`
double Model::deviceEval(DeviceSession &s, int which, bool objective) {
  double out[4] = {0, 0, 0, 0};
  const int variant = deviceVariant();
  // row-sharded sessions: every engine evaluates its own rows (and the norms of its own ranges), the sums add up
  const int nEval = s.mode == matfac::GROUP_ROWS ? s.world() : 1;
  for (int r = 0; r < nEval; r++) {
    double part[4] = {0, 0, 0, 0};
    s.check(mfb_eval(s.engineOf(r), which, MFB_CURRENT, variant, objective && variant == MFB_IFWMF, objective, part));
    for (int k = 0; k < 4; k++) out[k] += part[k];
  }
  if (objective) return out[0] + out[2] * uReg + out[3] * iReg;
  return sqrt(out[0] / out[1]);
}

double Model::RMSE(gk_csr_t *mat, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems) {
  if (dev_ && dev_->slotOf(mat) >= 0) return deviceEval(*dev_, dev_->slotOf(mat), false);  // inside a trainer
  int which = 0;
  DeviceSession &s = DeviceSession::forMatrix(mat, nUsers, nItems, facDim, &which);
  s.setMasks(invalidUsers, invalidItems);
  uploadFactors(s);
  uploadAuxAll(s, nullptr, invalidUsers, invalidItems);
  return deviceEval(s, which, false);
}

double Model::RMSE(gk_csr_t *mat) {
  std::unordered_set<int> none;
  return RMSE(mat, none, none);
}

void Model::groupSE(gk_csr_t *mat, const std::vector<uint8_t> &userGroup, const std::vector<uint8_t> &itemGroup,
                    std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, double out[32]) {
  std::vector<uint8_t> ug(nUsers, 255), ig(nItems, 255);
  for (size_t i = 0; i < userGroup.size() && i < ug.size(); i++) ug[i] = userGroup[i];
  for (size_t i = 0; i < itemGroup.size() && i < ig.size(); i++) ig[i] = itemGroup[i];
  int which = 0;
  DeviceSession *s = dev_ && dev_->slotOf(mat) >= 0 ? dev_ : nullptr;
  if (s) {
    which = s->slotOf(mat);
  } else {
    s = &DeviceSession::forMatrix(mat, nUsers, nItems, facDim, &which);
    s->setMasks(invalidUsers, invalidItems);
    uploadFactors(*s);
    uploadAuxAll(*s, nullptr, invalidUsers, invalidItems);
  }
  if (s->mode == matfac::GROUP_ROWS && s->world() > 1) {  // every engine its own rows
    for (int k = 0; k < 32; k++) out[k] = 0;
    for (int r = 0; r < s->world(); r++) {
      double part[32];
      s->check(mfb_eval_groups(s->engineOf(r), which, MFB_CURRENT, deviceVariant(), ug.data(), ig.data(), part));
      for (int k = 0; k < 32; k++) out[k] += part[k];
    }
    return;
  }
  s->check(mfb_eval_groups(s->eng, which, MFB_CURRENT, deviceVariant(), ug.data(), ig.data(), out));
}

// ---- ranking metrics (model.cpp:760-1332) ------------------------------------------------------------------------
void Model::rankPositions(const Data &data, gk_csr_t *testMat, std::unordered_set<int> &invalidUsers,
                          std::unordered_set<int> &invalidItems, std::vector<int32_t> &pos, std::vector<int32_t> &testItem) {
  DeviceSession &s = DeviceSession::forData(data, facDim);
  const int which = s.slotOf(testMat);
  if (which != MFB_VAL && which != MFB_TEST) matfac::fatal("ranking metrics: testMat must be the validation or the test matrix of data");
  s.setMasks(invalidUsers, invalidItems);
  uploadFactors(s);
  uploadAuxAll(s, &data, invalidUsers, invalidItems);
  pos.assign(nUsers, -1);
  testItem.assign(nUsers, -1);
  s.check(mfb_rank_positions(s.eng, which, MFB_CURRENT, deviceVariant(), pos.data(), testItem.data()));
}

// hitRate / arHR and their U / I variants differ in who counts (model.cpp:1000, :1050-1057, :1113) and in what a hit is
// worth (:1024 vs :1196); the position of the test item in the sorted list of the N best candidates is the device's count
std::pair<double, double> Model::hitStats(const Data &data, gk_csr_t *testMat, std::unordered_set<int> &invalidUsers,
                                          std::unordered_set<int> &invalidItems, const std::unordered_set<int> *filtUsers,
                                          const std::unordered_set<int> *filtItems, int N, bool reciprocal) {
  std::vector<int32_t> pos, tst;
  rankPositions(data, testMat, invalidUsers, invalidItems, pos, tst);
  double hits = 0, counted = 0;
  for (int u = 0; u < data.trainMat->nrows && u < nUsers; u++) {
    if (pos[u] == -1) continue;  // invalid user (or no rating to test against)
    if (filtUsers && filtUsers->count(u) == 0) continue;
    if (filtItems && filtItems->count(tst[u]) == 0) continue;
    if (pos[u] >= 0 && pos[u] < N) hits += reciprocal ? 1.0 / (pos[u] + 1) : 1.0;
    counted += 1;
  }
  return std::make_pair(hits, counted);
}

double Model::hitRate(const Data &data, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, nullptr, nullptr, 10, false);
  return hc.first / hc.second;
}
std::pair<int, double> Model::hitRateU(const Data &data, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                       std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, &filtUsers, nullptr, 10, false);
  return std::make_pair((int)hc.first, hc.first / hc.second);
}
std::pair<int, double> Model::hitRateI(const Data &data, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                       std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, nullptr, &filtItems, 10, false);
  return std::make_pair((int)hc.first, hc.first / hc.second);
}
double Model::arHR(const Data &data, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, nullptr, nullptr, 1000, true);
  return hc.first / hc.second;
}
std::pair<double, double> Model::arHRU(const Data &data, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                       std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, &filtUsers, nullptr, 1000, true);
  return std::make_pair(hc.first, hc.first / hc.second);
}
std::pair<double, double> Model::arHRI(const Data &data, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                       std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  auto hc = hitStats(data, testMat, invalidUsers, invalidItems, nullptr, &filtItems, 1000, true);
  return std::make_pair(hc.first, hc.first / hc.second);
}

// NDCG / NDCGU / NDCGI (model.cpp:760-978): per user the N = 10 best predicted of its test ratings (heap ordered by the
// prediction), DCG in that order over the ideal DCG of those same ratings; users with fewer than two ratings left or a
// zero ideal DCG do not count.  The predictions are the device's (mfb_predict); the per-user heaps are a few entries.
std::pair<int, double> Model::ndcgStats(gk_csr_t *testMat, std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems,
                                        const std::unordered_set<int> *filtUsers, const std::unordered_set<int> *filtItems) {
  int which = 0;
  DeviceSession &s = DeviceSession::forMatrix(testMat, nUsers, nItems, facDim, &which);
  s.setMasks(invalidUsers, invalidItems);
  uploadFactors(s);
  uploadAuxAll(s, nullptr, invalidUsers, invalidItems);
  const int64_t nnz = testMat->rowptr[testMat->nrows];
  std::vector<float> pred((size_t)std::max<int64_t>(nnz, 1));
  s.check(mfb_predict(s.eng, which, MFB_CURRENT, deviceVariant(), pred.data()));
  const int N = 10;
  typedef std::tuple<int, float, float> Triplet;  // item, actual, predicted
  auto byPred = [](const Triplet &a, const Triplet &b) { return std::get<2>(a) > std::get<2>(b); };
  auto byAct = [](const Triplet &a, const Triplet &b) { return std::get<1>(a) > std::get<1>(b); };
  int nValUsers = 0;
  double ndcg = 0;
  for (int u = 0; u < testMat->nrows; u++) {
    if (invalidUsers.count(u) > 0 || (filtUsers && filtUsers->count(u) == 0)) continue;
    std::vector<Triplet> rs;
    for (int64_t ii = testMat->rowptr[u]; ii < testMat->rowptr[u + 1]; ii++) {
      const int item = testMat->rowind[ii];
      if (invalidItems.count(item) > 0 || (filtItems && filtItems->count(item) == 0)) continue;
      rs.push_back(Triplet(item, testMat->rowval[ii], pred[ii]));
      std::push_heap(rs.begin(), rs.end(), byPred);
      if ((int)rs.size() > N) {
        std::pop_heap(rs.begin(), rs.end(), byPred);
        rs.pop_back();
      }
    }
    if (rs.size() < 2) continue;
    std::sort(rs.begin(), rs.end(), byPred);
    float u_ndcg = 0.0, u_dcg_max = 0.0;
    for (int i = 0; i < N && i < (int)rs.size(); i++) u_ndcg += (std::pow(2.0, std::get<1>(rs[i])) - 1) / std::log2((i + 1) + 1);
    std::sort(rs.begin(), rs.end(), byAct);
    for (int i = 0; i < N && i < (int)rs.size(); i++) u_dcg_max += (std::pow(2.0, std::get<1>(rs[i])) - 1) / std::log2((i + 1) + 1);
    if (!(u_dcg_max > EPS)) continue;
    ndcg += u_ndcg / u_dcg_max;
    nValUsers++;
  }
  return std::make_pair(nValUsers, ndcg / nValUsers);
}
double Model::NDCG(std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  return ndcgStats(testMat, invalidUsers, invalidItems, nullptr, nullptr).second;
}
std::pair<int, double> Model::NDCGU(std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                    std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  return ndcgStats(testMat, invalidUsers, invalidItems, &filtUsers, nullptr);
}
std::pair<int, double> Model::NDCGI(std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                    std::unordered_set<int> &invalidItems, gk_csr_t *testMat) {
  return ndcgStats(testMat, invalidUsers, invalidItems, nullptr, &filtItems);
}

static std::vector<uint8_t> groupOf(const std::unordered_set<int> &ids, int n) {
  std::vector<uint8_t> g(n, 255);
  for (int id : ids)
    if (id >= 0 && id < n) g[id] = 0;
  return g;
}

std::pair<int, double> Model::SE(gk_csr_t *mat, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                 std::unordered_set<int> &invalidItems) {
  double out[32];
  groupSE(mat, std::vector<uint8_t>(nUsers, 255), groupOf(filtItems, nItems), invalidUsers, invalidItems, out);
  return std::make_pair((int)out[1], out[0]);
}

std::pair<int, double> Model::RMSE(gk_csr_t *mat, std::unordered_set<int> &filtItems, std::unordered_set<int> &invalidUsers,
                                   std::unordered_set<int> &invalidItems) {
  auto se = SE(mat, filtItems, invalidUsers, invalidItems);
  return std::make_pair(se.first, sqrt(se.second / se.first));
}

std::pair<int, double> Model::RMSEU(gk_csr_t *mat, std::unordered_set<int> &filtUsers, std::unordered_set<int> &invalidUsers,
                                    std::unordered_set<int> &invalidItems) {
  double out[32];
  groupSE(mat, groupOf(filtUsers, nUsers), std::vector<uint8_t>(nItems, 255), invalidUsers, invalidItems, out);
  return std::make_pair((int)out[16 + 1], sqrt(out[16] / out[16 + 1]));
}

double Model::objective(const Data &data, std::unordered_set<int> &invalidUsers,
                        std::unordered_set<int> &invalidItems) {
  if (dev_ && dev_->slotOf(data.trainMat) >= 0) return deviceEval(*dev_, MFB_TRAIN, true);
  DeviceSession &s = DeviceSession::forData(data, facDim);
  s.setMasks(invalidUsers, invalidItems);
  uploadFactors(s);
  uploadAuxAll(s, &data, invalidUsers, invalidItems);
  return deviceEval(s, MFB_TRAIN, true);
}

double Model::objective(const Data &data) {
  std::unordered_set<int> none;
  return objective(data, none, none);
}

// ---- early stopping --------------------------------------------------------------------------
bool Model::isTerminateModel(Model &bestModel, const Data &data, int iter, int &bestIter, double &bestObj,
                             double &prevObj, double &bestValRMSE, double &prevValRMSE,
                             std::unordered_set<int> &invalidUsers, std::unordered_set<int> &invalidItems) {
  bool ret = false;
  const double currObj = objective(data, invalidUsers, invalidItems);
  double currValRMSE = -1;
  if (data.valMat) {
    currValRMSE = RMSE(data.valMat, invalidUsers, invalidItems);
  } else {
    std::cerr << "\nNo validation data" << std::endl;
    exit(0);
  }
  const bool onDevice = dev_ != nullptr;
  if (currObj != currObj || currValRMSE != currValRMSE) {
    std::cout << "Found nan " << std::endl;
    if (learnRate > 1e-5) {
      // *this = bestModel; learnRate /= 2   (model.cpp:1490-1493)
      copyScalarsFrom(bestModel);
      if (onDevice) {
        if (bestOnDevice_) {
          dev_->check(mfb_restore_best(dev_->eng));
        } else {
          dev_->check(mfb_upload_factors(dev_->eng, bestModel.uFac.data(), facDim, bestModel.iFac.data(), facDim));
        }
        dev_->broadcastFactors();  // the other engines of the session restart from the same factors
      } else {
        uFac = bestModel.uFac;
        iFac = bestModel.iFac;
      }
      learnRate = learnRate / 2;
      return false;
    }
    return true;
  }
  if (currValRMSE < bestValRMSE) {
    // bestModel = *this (model.cpp:1500-1504): scalars now, factors as a device snapshot
    bestModel.copyScalarsFrom(*this);
    if (onDevice) {
      dev_->check(mfb_snapshot_best(dev_->eng));
      bestOnDevice_ = true;
    } else {
      bestModel.uFac = uFac;
      bestModel.iFac = iFac;
    }
    bestValRMSE = currValRMSE;
    bestIter = iter;
  }
  if (iter - bestIter >= 100) {
    if (learnRate > 1e-5) learnRate = learnRate / 2;
  }
  if (iter - bestIter >= CHANCE_ITER) {
    printf("\nNOT CONVERGED: bestIter:%d bestObj: %.10e bestValRMSE: %.10e currIter:%d currObj: %.10e currValRMSE: %.10e",
           bestIter, bestObj, bestValRMSE, iter, currObj, currValRMSE);
    ret = true;
  }
  if (fabs(prevObj - currObj) < EPS) {
    printf("\nConverged in iteration: %d prevObj: %.10e currObj: %.10e bestValRMSE: %.10e", iter, prevObj, currObj,
           bestValRMSE);
    ret = true;
  }
  prevObj = currObj;
  prevValRMSE = currValRMSE;
  return ret;
}

bool Model::isTerminateModel(Model &bestModel, const Data &data, int iter, int &bestIter, double &bestObj,
                             double &prevObj, std::unordered_set<int> &invalidUsers,
                             std::unordered_set<int> &invalidItems) {
  bool ret = false;
  const double currObj = objective(data, invalidUsers, invalidItems);
  if (iter > 0) {
    if (currObj < bestObj) {
      bestModel.copyScalarsFrom(*this);
      if (dev_) {
        dev_->check(mfb_snapshot_best(dev_->eng));
        bestOnDevice_ = true;
      } else {
        bestModel.uFac = uFac;
        bestModel.iFac = iFac;
      }
      bestObj = currObj;
      bestIter = iter;
    }
    if (iter - bestIter >= 100) {
      if (learnRate > 1e-5) learnRate = learnRate / 2;
      else if (learnRate < 1e-5) learnRate = 1e-5;
    }
    if (iter - bestIter >= 500) ret = true;
    if (fabs(prevObj - currObj) < EPS) ret = true;
  }
  if (iter == 0) {
    bestObj = currObj;
    bestIter = iter;
  }
  prevObj = currObj;
  return ret;
}

// ---- trainer skeletons ------------------------------------------------------------------------
// Common preamble of every trainer (modelMF.cpp:34-61): invalid ids, initial objective and
// validation RMSE — plus the device set-up that replaces the host arrays.
void Model::beginTraining(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                          std::unordered_set<int> &invalidItems, Stop &st, const char *tag, int groupMode) {
  (void)bestModel;
  gk_csr_t *trainMat = data.trainMat;
  std::vector<std::unordered_set<int>> uISet;
  genStats(trainMat, uISet, std::to_string(trainSeed));
  getInvalidUsersItems(trainMat, uISet, invalidUsers, invalidItems);
  for (int u = trainMat->nrows; u < data.nUsers; u++) invalidUsers.insert(u);
  for (int item = trainMat->ncols; item < data.nItems; item++) invalidItems.insert(item);

  DeviceSession &s = DeviceSession::forData(data, facDim);
  s.ensureGroup((matfac::GroupMode)groupMode);  // every visible GPU for the trainers that shard
  s.setMasks(invalidUsers, invalidItems);
  uploadFactors(s);
  uploadAuxAll(s, &data, invalidUsers, invalidItems);
  dev_ = &s;
  bestOnDevice_ = false;

  st.prevObj = objective(data, invalidUsers, invalidItems);
  st.bestObj = st.prevObj;
  st.bestValRMSE = st.prevValRMSE = RMSE(data.valMat, invalidUsers, invalidItems);
  std::cout << "\nObj aftr svd: " << st.prevObj << " Train RMSE: " << RMSE(data.trainMat, invalidUsers, invalidItems)
            << " Val RMSE: " << st.bestValRMSE;
  std::cout << "\n" << tag << " trainSeed: " << trainSeed << " invalidUsers: " << invalidUsers.size()
            << " invalidItems: " << invalidItems.size() << std::endl;
}

void Model::syncBest(Model &bestModel) {
  if (dev_ && bestOnDevice_) {
    dev_->check(mfb_download_factors(dev_->eng, MFB_BEST, bestModel.uFac.data(), facDim, bestModel.iFac.data(), facDim));
    bestOnDevice_ = false;
  }
}

void Model::endTraining(Model &bestModel) {
  if (!dev_) return;
  syncBest(bestModel);
  dev_->check(mfb_download_factors(dev_->eng, MFB_CURRENT, uFac.data(), facDim, iFac.data(), facDim));
  dev_ = nullptr;
}

bool Model::afterEpoch(const Data &data, Model &bestModel, int iter, Stop &st, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems, double subIterDuration, const char *tag, bool saves) {
  if (!(iter % OBJ_ITER == 0 || iter == maxIter - 1)) return false;
  if (isTerminateModel(bestModel, data, iter, st.bestIter, st.bestObj, st.prevObj, st.bestValRMSE, st.prevValRMSE,
                       invalidUsers, invalidItems))
    return true;
  if (iter % DISP_ITER == 0) {
    std::cout << tag << " trainSeed: " << trainSeed << " Iter: " << iter << " Objective: " << std::scientific
              << st.prevObj << " Train RMSE: " << RMSE(data.trainMat, invalidUsers, invalidItems)
              << " Val RMSE: " << st.prevValRMSE << " subIterDuration: " << subIterDuration << std::endl;
  }
  if (saves && (iter % SAVE_ITER == 0 || iter == maxIter - 1)) {
    syncBest(bestModel);
    bestModel.saveFacs(std::string(data.prefix));
  }
  return false;
}

static double elapsedSeconds(DeviceSession &s) {
  float ms = 0;
  s.check(mfb_event_elapsed_ms(s.eng, 0, 1, &ms));
  return ms * 1e-3;
}

// Serial / Hogwild trainers on the device: every valid rating once per epoch in a fresh
// pseudo-random order (the reference reshuffles each epoch: modelMF.cpp:76-81).
void Model::runFlatSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                       std::unordered_set<int> &invalidItems, const char *tag, bool saves) {
  Stop st;
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, tag);
  DeviceSession &s = *dev_;
  s.check(mfb_set_option(s.eng, "sgd_shuffle_seed", (double)(uint32_t)trainSeed));
  s.check(mfb_sgd_plan(s.eng, 1, nullptr, nullptr));
  const int variant = deviceVariant();
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    s.check(mfb_sgd_epoch_flat(s.eng, variant, learnRate, uReg, iReg, (uint64_t)(uint32_t)trainSeed, (uint64_t)iter));
    s.check(mfb_event_record(s.eng, 1));
    const double dur = elapsedSeconds(s);
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, dur, tag, saves)) break;
  }
  endTraining(bestModel);
  if (saves) bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}

// User-major SGD over the whole matrix (trainUShuffle, modelMF.cpp:626-660): one block, user runs in the queue order of
// the kernel (the reference reshuffles the users every epoch — with thousands of runs in flight the order of the queue
// is immaterial), every run from its own pseudo-random start (see runStratifiedSgd).
void Model::runUserMajorSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                            std::unordered_set<int> &invalidItems, const char *tag, bool saves) {
  Stop st;
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, tag);
  DeviceSession &s = *dev_;
  s.check(mfb_set_option(s.eng, "sgd_shuffle_seed", (double)(uint32_t)trainSeed));
  s.check(mfb_set_option(s.eng, "sgd_rotate", 1));
  s.check(mfb_set_option(s.eng, "sgd_block_order", 0));
  s.check(mfb_sgd_plan(s.eng, 1, nullptr, nullptr));
  const int variant = deviceVariant();
  const int32_t whole[2] = {0, 0};
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    s.check(mfb_sgd_subepoch(s.eng, whole, 1, variant, learnRate, uReg, iReg, (uint64_t)(uint32_t)trainSeed, (uint64_t)iter));
    s.check(mfb_event_record(s.eng, 1));
    const double dur = elapsedSeconds(s);
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, dur, tag, saves)) break;
  }
  endTraining(bestModel);
  if (saves) bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}

// Stratified SGD (modelMF.cpp:154-350 and the IFWMF / TMF / TMF+Dropout twins): users and items
// are shuffled by mt19937(trainSeed) and cut into P = omp_get_max_threads() parts with the
// reference's boundary rule; every epoch draws P random permutation schedules from the same
// engine.  All of that is host code calling libstdc++ exactly like the reference (bit-exact
// partitions and schedules); the P conflict-free blocks of a schedule run as one device launch.
void Model::runStratifiedSgd(const Data &data, Model &bestModel, std::unordered_set<int> &invalidUsers,
                             std::unordered_set<int> &invalidItems, const char *tag, bool saves) {
  Stop st;
  beginTraining(data, bestModel, invalidUsers, invalidItems, st, tag, matfac::GROUP_STRATA);
  DeviceSession &s = *dev_;
  gk_csr_t *trainMat = data.trainMat;
  std::vector<int> trainUsers = matfac::validIds(trainMat->nrows, invalidUsers);
  std::vector<int> trainItems = matfac::validIds(trainMat->ncols, invalidItems);
  std::mt19937 mt(trainSeed);
  std::shuffle(trainUsers.begin(), trainUsers.end(), mt);
  std::shuffle(trainItems.begin(), trainItems.end(), mt);
  int P = omp_get_max_threads();
  if (P > 64) P = 64;  // one launch carries at most 64 blocks
  std::cout << "maxThreads: " << P << std::endl;
  std::cout << "train users: " << trainUsers.size() << " usersPerPart: " << trainUsers.size() / P << std::endl;
  std::cout << "train items: " << trainItems.size() << " itemsPerPart: " << trainItems.size() / P << std::endl;
  std::vector<int> userPart = matfac::partitionIds(trainUsers, P, nUsers);
  std::vector<int> itemPart = matfac::partitionIds(trainItems, P, nItems);
  // Several GPUs: user part p is pinned to engine p mod N — that engine plans and trains only the ratings
  // of its own user parts — and an item part travels to the engine that needs it in the next sub-epoch (peer-memory
  // push + sequence flag, csrc/comm.cu).  With one engine this is the plain P x P plan.
  const int N = s.world();
  std::vector<std::vector<int32_t>> ownUsers(N);
  for (int r = 0; r < N; r++) {
    std::vector<int> mine(userPart);
    if (N > 1)
      for (int u = 0; u < nUsers; u++) {
        if (mine[u] >= 0 && mine[u] % N != r) mine[u] = -1;
        else if (mine[u] >= 0) ownUsers[r].push_back(u);
      }
    s.check(mfb_set_option(s.engineOf(r), "sgd_shuffle_seed", (double)(uint32_t)trainSeed));
    // every user's run starts at its own pseudo-random item and wraps: thousands of runs that all begin at the lowest
    // item ids would hit the same few item rows at once (measured: the 1/20-scale bench matrix diverges in epoch 1
    // without it at learnrate 0.002, where the reference's own trainSGDPar does not; profiles/r2_dsgd_parity.md)
    s.check(mfb_set_option(s.engineOf(r), "sgd_rotate", 1));
    s.check(mfb_sgd_plan(s.engineOf(r), P, mine.data(), itemPart.data()));
  }
  const int variant = deviceVariant();
  std::vector<std::pair<int, int>> updateSeq, nextSeq;
  std::vector<int> holder(P, -1);                              // engine that holds the current rows of an item part (-1: all)
  std::vector<std::vector<uint64_t>> pushed(N, std::vector<uint64_t>(N, 0));  // pushes issued src -> dst so far
  sgdUpdateBlockSeq(P, updateSeq, mt);
  for (int iter = 0; iter < maxIter; iter++) {
    s.check(mfb_event_record(s.eng, 0));
    for (int k = 0; k < P; k++) {
      const bool lastOfEpoch = k == P - 1;
      // the reference draws one sequence per sub-epoch from the same engine (modelMF.cpp:274); drawing the next one
      // before this sub-epoch is launched (to know where every item part goes) leaves the stream unchanged
      if (!lastOfEpoch || iter + 1 < maxIter) sgdUpdateBlockSeq(P, nextSeq, mt);
      std::vector<std::vector<int32_t>> blocks(N);
      for (int t = 0; t < P; t++) {
        const int r = N > 1 ? updateSeq[t].first % N : 0;
        blocks[r].push_back(updateSeq[t].first);
        blocks[r].push_back(updateSeq[t].second);
      }
      for (int r = 0; r < N; r++) {
        if (blocks[r].empty()) continue;
        mfb_engine *e = s.engineOf(r);
        if (N > 1) {  // wait until every item part of this sub-epoch has arrived from its previous holder
          std::vector<char> seen(N, 0);
          for (size_t b = 1; b < blocks[r].size(); b += 2) {
            const int h = holder[blocks[r][b]];
            if (h >= 0 && h != r && !seen[h]) {
              seen[h] = 1;
              s.check(mfb_comm_wait_block(e, h, pushed[h][r]));
            }
          }
        }
        s.check(mfb_sgd_subepoch(e, blocks[r].data(), (int32_t)(blocks[r].size() / 2), variant, learnRate, uReg, iReg,
                                 (uint64_t)(uint32_t)trainSeed, (uint64_t)iter * P + k));
      }
      if (N > 1) {
        for (int t = 0; t < P; t++) holder[updateSeq[t].second] = updateSeq[t].first % N;
        if (!lastOfEpoch) {
          // hand every item part to the engine that trains it next; holder[] keeps naming the source of the newest
          // rows, which is what the receiver waits on (pushes of one source complete in issue order)
          for (int t = 0; t < P; t++) {
            const int part = nextSeq[t].second, dst = nextSeq[t].first % N, src = holder[part];
            if (src >= 0 && src != dst) s.check(mfb_dsgd_push_block(s.engineOf(src), part, dst, ++pushed[src][dst]));
          }
        }
      }
      updateSeq.swap(nextSeq);
    }
    if (N > 1) {
      // end of the epoch: every engine publishes the item parts it holds and its own users' rows to all peers, so
      // that the evaluation (engine 0) and the next epoch start from complete, identical factors everywhere
      for (int b = 0; b < P; b++)
        if (holder[b] >= 0) s.check(mfb_dsgd_push_block(s.engineOf(holder[b]), b, -1, 0));
      for (int r = 0; r < N; r++)
        s.check(mfb_comm_allgather_rows(s.engineOf(r), MFB_USER, ownUsers[r].empty() ? nullptr : ownUsers[r].data(), 0,
                                        (int32_t)ownUsers[r].size()));
      std::fill(holder.begin(), holder.end(), -1);
    }
    s.check(mfb_event_record(s.eng, 1));
    const double dur = elapsedSeconds(s);
    if (afterEpoch(data, bestModel, iter, st, invalidUsers, invalidItems, dur, tag, saves)) break;
  }
  endTraining(bestModel);
  if (saves) bestModel.saveFacs(std::string(data.prefix));
  std::cout << "\nBest model validation RMSE: " << bestModel.RMSE(data.valMat, invalidUsers, invalidItems) << std::endl;
}
