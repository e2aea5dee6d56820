// TEST INFRASTRUCTURE — CPU restatement of mohit-shrma/matfac's training hot path.
//
// Nothing under oracle/ is on the product path: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library (see mf_oracle.h).
//
// Every function cites the reference statements it follows (paths relative to
// /root/reference).  The arithmetic types of the reference are kept exactly: fp32 factor
// storage, fp32 sequential dot products (what Eigen emits for strided rows), the bracketed
// gradient evaluated in double and rounded once on the "-=" store, double num/denom in
// CCD++, fp32 Gram + fp32 pivoted LDL^T in ALS.  libstdc++'s <random>/<algorithm> are called
// the same way the reference calls them, so schedules, partitions and factor
// initialisation are bit-identical to the reference built with the same GCC.
//
// Parity pinning: tests/test_oracle_vs_ref.py runs this library next to oracle/_ref/mf_ref
// (the reference's own translation units compiled unmodified) and requires bit-identical
// factors; the outputs of that run are committed under tests/golden/.
#include "mf_oracle.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>

#include <omp.h>

// const.h:4-12
static const int kObjIter = 1;
static const int kChanceIter = 500;
static const double kEps = 1e-5;

struct Csr {
  int nrows = 0, ncols = 0;
  std::vector<int64_t> rowptr, colptr;
  std::vector<int32_t> rowind, colind;
  std::vector<float> rowval, colval;
  int64_t nnz() const { return rowptr.empty() ? 0 : rowptr.back(); }
};

struct mfo_data {
  Csr mat[3];  // train, val, test
  int nUsers = 0, nItems = 0;
};

// gk_csr_CreateIndex(mat, GK_CSR_COL): stable counting sort (rows ascend inside a column).
static void buildCsc(Csr &m) {
  int64_t nnz = m.nnz();
  m.colptr.assign((size_t)m.ncols + 1, 0);
  m.colind.resize(nnz);
  m.colval.resize(nnz);
  for (int64_t j = 0; j < nnz; j++) m.colptr[m.rowind[j] + 1]++;
  for (int c = 0; c < m.ncols; c++) m.colptr[c + 1] += m.colptr[c];
  std::vector<int64_t> next(m.colptr.begin(), m.colptr.end());
  for (int r = 0; r < m.nrows; r++)
    for (int64_t j = m.rowptr[r]; j < m.rowptr[r + 1]; j++) {
      int64_t d = next[m.rowind[j]]++;
      m.colind[d] = r;
      m.colval[d] = m.rowval[j];
    }
}

// gk_csr_Read(file, GK_CSR_FMT_CSR, readvals=1, numbering=0) as datastruct.cpp:16 calls it:
// one row per line, "col val" pairs, ncols = max column index + 1.
static bool readTextCsr(const char *path, Csr &m) {
  FILE *fp = fopen(path, "rb");
  if (!fp) return false;
  fseek(fp, 0, SEEK_END);
  long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  std::vector<char> buf((size_t)sz + 2);
  if (fread(buf.data(), 1, (size_t)sz, fp) != (size_t)sz) { fclose(fp); return false; }
  fclose(fp);
  if (sz > 0 && buf[sz - 1] != '\n') buf[sz++] = '\n';
  buf[sz] = '\0';
  m.rowptr.assign(1, 0);
  m.rowind.clear();
  m.rowval.clear();
  int maxcol = -1;
  char *p = buf.data(), *endbuf = buf.data() + sz;
  while (p < endbuf) {
    char *eol = (char *)memchr(p, '\n', endbuf - p);
    *eol = '\0';
    char *q = p;
    for (;;) {
      char *e;
      long col = strtol(q, &e, 10);
      if (e == q) break;
      q = e;
      float v = strtof(q, &e);
      if (e == q) return false;
      q = e;
      m.rowind.push_back((int32_t)col);
      m.rowval.push_back(v);
      if ((int)col > maxcol) maxcol = (int)col;
    }
    m.rowptr.push_back((int64_t)m.rowind.size());
    p = eol + 1;
  }
  m.nrows = (int)m.rowptr.size() - 1;
  m.ncols = maxcol + 1;
  buildCsc(m);
  return true;
}

// datastruct.cpp:23,91: nUsers = train rows; nItems = max column index over all three + 1.
static void finishData(mfo_data *d) {
  d->nUsers = d->mat[0].nrows;
  int maxItem = d->mat[0].ncols - 1;
  for (int w = 1; w < 3; w++) maxItem = std::max(maxItem, d->mat[w].ncols - 1);
  d->nItems = maxItem + 1;
}

extern "C" mfo_data *mfo_data_read(const char *train, const char *val, const char *test) {
  mfo_data *d = new mfo_data();
  const char *paths[3] = {train, val, test};
  for (int w = 0; w < 3; w++)
    if (!readTextCsr(paths[w], d->mat[w])) { delete d; return nullptr; }
  finishData(d);
  return d;
}

static void fromArrays(Csr &m, int64_t rows, const int64_t *ptr, const int32_t *ind, const float *val) {
  m.nrows = (int)rows;
  m.rowptr.assign(ptr, ptr + rows + 1);
  int64_t nnz = ptr[rows];
  m.rowind.assign(ind, ind + nnz);
  m.rowval.assign(val, val + nnz);
  int maxcol = -1;
  for (int64_t j = 0; j < nnz; j++) maxcol = std::max(maxcol, (int)ind[j]);
  m.ncols = maxcol + 1;
  buildCsc(m);
}

extern "C" mfo_data *mfo_data_from_arrays(int64_t tr_rows, const int64_t *tr_ptr, const int32_t *tr_ind,
                                          const float *tr_val, int64_t va_rows, const int64_t *va_ptr,
                                          const int32_t *va_ind, const float *va_val, int64_t te_rows,
                                          const int64_t *te_ptr, const int32_t *te_ind, const float *te_val) {
  mfo_data *d = new mfo_data();
  fromArrays(d->mat[0], tr_rows, tr_ptr, tr_ind, tr_val);
  fromArrays(d->mat[1], va_rows, va_ptr, va_ind, va_val);
  fromArrays(d->mat[2], te_rows, te_ptr, te_ind, te_val);
  finishData(d);
  return d;
}

extern "C" void mfo_data_free(mfo_data *d) { delete d; }
extern "C" int mfo_data_nusers(const mfo_data *d) { return d->nUsers; }
extern "C" int mfo_data_nitems(const mfo_data *d) { return d->nItems; }
extern "C" void mfo_data_dims(const mfo_data *d, int which, int64_t dims[3]) {
  dims[0] = d->mat[which].nrows;
  dims[1] = d->mat[which].ncols;
  dims[2] = d->mat[which].nnz();
}
extern "C" void mfo_data_csr(const mfo_data *d, int which, int64_t *rowptr, int32_t *rowind, float *rowval) {
  const Csr &m = d->mat[which];
  std::copy(m.rowptr.begin(), m.rowptr.end(), rowptr);
  std::copy(m.rowind.begin(), m.rowind.end(), rowind);
  std::copy(m.rowval.begin(), m.rowval.end(), rowval);
}
extern "C" void mfo_data_csc(const mfo_data *d, int which, int64_t *colptr, int32_t *colind, float *colval) {
  const Csr &m = d->mat[which];
  std::copy(m.colptr.begin(), m.colptr.end(), colptr);
  std::copy(m.colind.begin(), m.colind.end(), colind);
  std::copy(m.colval.begin(), m.colval.end(), colval);
}

// ------------------------------------------------------------------------------------------
// Model state.  Factors are kept row-major [n][facDim]; the reference's column-major Eigen
// storage (model.h:37-38) only changes addresses, not arithmetic.
struct Facs {
  std::vector<float> U, V;
  float learnRate = 0;  // copied by "bestModel = *this" (model.cpp:1501)
};

struct HistEntry {
  std::vector<float> U, V;
  double obj, valRmse;
};

struct mfo_model {
  int algo;
  int nUsers, nItems, facDim, maxIter, seed, nThreads;
  float uReg, iReg, origLearnRate, rhoRMS, alpha;
  Facs cur, best;
  std::unordered_set<int> invalidUsers, invalidItems;
  // frequency-derived state (main.cpp:1262-1264 -> ctor arguments of the derived models)
  std::vector<double> userFreq, itemFreq;
  double meanFreq = 0, stdFreq = 1;
  std::vector<int> cdfRanks;
  std::vector<double> invPopU, invPopI;  // std::map<int,double> in the reference; 0 when absent
  std::vector<HistEntry> hist;
  int keepHistory = 0;
  std::vector<double> epochSecs;  // the reference's subIterDuration (modelMF.cpp:75,106-109)

  float &u(int uu, int k) { return cur.U[(size_t)uu * facDim + k]; }
  float &v(int ii, int k) { return cur.V[(size_t)ii * facDim + k]; }
};

// util.cpp:278-294
static std::pair<double, double> meanStdDev(const std::vector<double> &v) {
  double sum = 0;
  for (size_t i = 0; i < v.size(); i++) sum += v[i];
  double mean = sum / v.size();
  double sq = 0;
  for (size_t i = 0; i < v.size(); i++) sq += (v[i] - mean) * (v[i] - mean);
  return std::make_pair(mean, sqrt(sq / v.size()));
}

// modelPoissonDropout.cpp:25-47 (factorial table: modelPoissonDropout.h:46-49)
static void initCdfRanks(mfo_model *m) {
  int r = m->facDim;
  std::vector<double> factorial;
  factorial.push_back(1);
  for (int i = 1; i <= r + 1; i++) factorial.push_back(factorial.back() * ((double)i));
  m->cdfRanks = std::vector<int>(r, 0);
  double cdf = 0, wt = 0;
  for (int lambda = 1; lambda <= r; lambda++) {
    cdf = std::exp(-lambda) * (std::pow(lambda, 0) / factorial[0]);
    int k = 0;
    for (k = 0; k < r; k++) {
      wt = std::exp(-lambda) * (std::pow(lambda, k + 1) / factorial[k + 1]);
      cdf += wt;
      if (cdf >= 0.99) break;
    }
    m->cdfRanks[lambda - 1] = k;
    if (k == r) m->cdfRanks[lambda - 1] = k - 1;
  }
}

extern "C" mfo_model *mfo_model_create(const mfo_data *d, const mfo_params *p, int algo) {
  mfo_model *m = new mfo_model();
  m->algo = algo;
  m->nUsers = d->nUsers;
  m->nItems = d->nItems;
  m->facDim = p->facDim;
  m->maxIter = p->maxIter;
  m->seed = p->seed;
  m->nThreads = p->nThreads > 0 ? p->nThreads : 1;
  m->uReg = p->uReg;
  m->iReg = p->iReg;
  m->origLearnRate = p->learnRate;
  m->rhoRMS = p->rhoRMS;
  m->alpha = p->alpha;
  m->cur.learnRate = p->learnRate;

  // model.cpp:2331-2350: one minstd_rand0 stream, users first then items, k inner.
  std::default_random_engine generator(p->seed);
  float lb = -0.01, ub = 0.01;
  std::uniform_real_distribution<double> dist(lb, ub);
  m->cur.U.resize((size_t)m->nUsers * m->facDim);
  m->cur.V.resize((size_t)m->nItems * m->facDim);
  for (int u = 0; u < m->nUsers; u++)
    for (int k = 0; k < m->facDim; k++) m->u(u, k) = dist(generator);
  for (int i = 0; i < m->nItems; i++)
    for (int k = 0; k < m->facDim; k++) m->v(i, k) = dist(generator);
  m->best = m->cur;  // bestModel is built from the same params and seed (main.cpp:1326-1327)

  // util.cpp:555-569 (getRowColFreq over the train matrix, empty rows/cols included)
  const Csr &tr = d->mat[0];
  m->userFreq.assign(tr.nrows, 0);
  m->itemFreq.assign(tr.ncols, 0);
  for (int u = 0; u < tr.nrows; u++)
    for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
      m->userFreq[u] += 1;
      m->itemFreq[tr.rowind[ii]] += 1;
    }
  if (algo == MFO_ALGO_TMF || algo == MFO_ALGO_TMFDROPOUT) {
    // modelDropoutSigmoid.h:88-94
    std::vector<double> concatVec(m->userFreq.begin(), m->userFreq.end());
    concatVec.insert(concatVec.end(), m->itemFreq.begin(), m->itemFreq.end());
    auto ms = meanStdDev(concatVec);
    m->meanFreq = ms.first;
    m->stdFreq = ms.second;
  }
  if (algo == MFO_ALGO_TMFDROPOUT) initCdfRanks(m);
  return m;
}

extern "C" void mfo_model_free(mfo_model *m) { delete m; }

// TMF effective rank (modelDropoutSigmoid.cpp:158-170 / :7-18; modelPoissonDropout.cpp:7-12,189-196)
static inline int tmfLambda(const mfo_model *m, int u, int item) {
  bool isUMinFreq = m->userFreq[u] < m->itemFreq[item];
  double scaleFreq = isUMinFreq ? (m->userFreq[u] - m->meanFreq) / m->stdFreq
                                : (m->itemFreq[item] - m->meanFreq) / m->stdFreq;
  double sigmPc = 1.0 / (1.0 + exp(-m->rhoRMS * (scaleFreq - m->alpha)));
  return (int)std::ceil(sigmPc * ((double)m->facDim));
}

// virtual estRating: model.cpp:547; modelDropoutSigmoid.cpp:5-24; modelPoissonDropout.cpp:5-23
static inline double estRating(const mfo_model *m, const Facs &f, int u, int item) {
  const int r = m->facDim;
  const float *pu = &f.U[(size_t)u * r], *pv = &f.V[(size_t)item * r];
  if (m->algo == MFO_ALGO_MF || m->algo == MFO_ALGO_IFWMF) {
    float s = 0;
    for (int k = 0; k < r; k++) s += pu[k] * pv[k];
    return s;
  }
  int lambda = tmfLambda(m, u, item);
  double rat = 0;
  if (m->algo == MFO_ALGO_TMF) {
    int updMinRank = lambda;
    if (updMinRank > r) updMinRank = r;
    for (int k = 0; k < updMinRank; k++) rat += pu[k] * pv[k];
    return rat;
  }
  for (int k = 0; k <= m->cdfRanks[lambda - 1] && k < r; k++) rat += pu[k] * pv[k];
  return rat;
}

// model.cpp:214-251
static double rmseMasked(const mfo_model *m, const Facs &f, const Csr &mat) {
  int nnz = 0;
  double rmse = 0;
#pragma omp parallel for reduction(+ : rmse, nnz) schedule(static)
  for (int u = 0; u < m->nUsers; u++) {
    if (m->invalidUsers.count(u) > 0) continue;
    for (int64_t ii = mat.rowptr[u]; ii < mat.rowptr[u + 1]; ii++) {
      int item = mat.rowind[ii];
      if (m->invalidItems.count(item) > 0 || item >= m->nItems) continue;
      double r_ui = mat.rowval[ii];
      double r_ui_est = estRating(m, f, u, item);
      double diff = r_ui - r_ui_est;
      rmse += diff * diff;
      nnz++;
    }
  }
  return sqrt(rmse / nnz);
}

static inline float ifwWeight(const mfo_model *m, int u, int item) {
  // modelInvPopMF.cpp:23-28,163-168
  float wt = m->invPopI[item];
  if (m->itemFreq[item] > m->userFreq[u]) wt = m->invPopU[u];
  wt = (1.0 / (1.0 + m->rhoRMS * wt));
  return wt;
}

// model.cpp:1770-1815; IFWMF override modelInvPopMF.cpp:3-55
static double objectiveMasked(const mfo_model *m, const Facs &f, const Csr &tr) {
  double rmse = 0, uRegErr = 0, iRegErr = 0;
  const int r = m->facDim;
#pragma omp parallel for reduction(+ : rmse, uRegErr) schedule(static)
  for (int u = 0; u < m->nUsers; u++) {
    if (m->invalidUsers.count(u) > 0) continue;
    for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
      int item = tr.rowind[ii];
      if (m->invalidItems.count(item) > 0) continue;
      float itemRat = tr.rowval[ii];
      double diff = itemRat - estRating(m, f, u, item);
      if (m->algo == MFO_ALGO_IFWMF) {
        float wt = ifwWeight(m, u, item);
        rmse += wt * diff * diff;
      } else {
        rmse += diff * diff;
      }
    }
    float s = 0;
    const float *pu = &f.U[(size_t)u * r];
    for (int k = 0; k < r; k++) s += pu[k] * pu[k];
    uRegErr += s;
  }
  uRegErr = uRegErr * m->uReg;
#pragma omp parallel for reduction(+ : iRegErr) schedule(static)
  for (int item = 0; item < m->nItems; item++) {
    if (m->invalidItems.count(item) > 0) continue;
    float s = 0;
    const float *pv = &f.V[(size_t)item * r];
    for (int k = 0; k < r; k++) s += pv[k] * pv[k];
    iRegErr += s;
  }
  iRegErr = iRegErr * m->iReg;
  return rmse + uRegErr + iRegErr;
}

// util.cpp:511-544 with empty ignore sets, plus the tail loops modelMF.cpp:40-45
static void computeInvalid(mfo_model *m, const mfo_data *d) {
  const Csr &tr = d->mat[0];
  m->invalidUsers.clear();
  m->invalidItems.clear();
  std::vector<int> uItemCount(tr.nrows, 0), iUserCount(tr.ncols, 0);
  for (int u = 0; u < tr.nrows; u++)
    for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
      uItemCount[u] += 1;
      iUserCount[tr.rowind[ii]] += 1;
    }
  for (int u = 0; u < tr.nrows; u++)
    if (0 == uItemCount[u]) m->invalidUsers.insert(u);
  for (int item = 0; item < tr.ncols; item++)
    if (0 == iUserCount[item]) m->invalidItems.insert(item);
  for (int u = tr.nrows; u < d->nUsers; u++) m->invalidUsers.insert(u);
  for (int item = tr.ncols; item < d->nItems; item++) m->invalidItems.insert(item);
}

extern "C" void mfo_compute_invalid(mfo_model *m, const mfo_data *d) { computeInvalid(m, d); }

// modelInvPopMF.cpp:86-114
static void initIfw(mfo_model *m, const Csr &tr, const std::vector<int> &trainUsers,
                    const std::vector<int> &trainItems) {
  int nTrainUsers = (int)trainUsers.size(), nTrainItems = (int)trainItems.size();
  m->invPopU.assign(std::max(m->nUsers, tr.nrows), 0.0);
  m->invPopI.assign(std::max(m->nItems, tr.ncols), 0.0);
  double sumPopScore = 0;
  for (auto &u : trainUsers) {
    m->invPopU[u] = m->userFreq[u] / ((double)nTrainItems);
    sumPopScore += m->invPopU[u];
  }
  for (auto &u : trainUsers) m->invPopU[u] = m->invPopU[u] / sumPopScore;
  sumPopScore = 0;
  for (auto &item : trainItems) {
    m->invPopI[item] = m->itemFreq[item] / ((double)nTrainUsers);
    sumPopScore += m->invPopI[item];
  }
  for (auto &item : trainItems) m->invPopI[item] = m->invPopI[item] / sumPopScore;
}

static void validIds(const mfo_model *m, const Csr &tr, std::vector<int> &trainUsers,
                     std::vector<int> &trainItems) {
  for (int u = 0; u < tr.nrows; u++)
    if (m->invalidUsers.count(u) == 0) trainUsers.push_back(u);
  for (int item = 0; item < tr.ncols; item++)
    if (m->invalidItems.count(item) == 0) trainItems.push_back(item);
}

struct EpochTimer {
  mfo_model *m;
  std::chrono::steady_clock::time_point t0;
  explicit EpochTimer(mfo_model *m) : m(m), t0(std::chrono::steady_clock::now()) {}
  void stop() { m->epochSecs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count()); }
};

// model.cpp:1471-1540 (validation variant).  Returns true to stop.
struct StopState {
  int bestIter = -1;
  double bestObj = 0, prevObj = 0, bestValRMSE = 0, prevValRMSE = 0;
};

static bool isTerminateModel(mfo_model *m, const mfo_data *d, int iter, StopState &s) {
  bool ret = false;
  double currObj = objectiveMasked(m, m->cur, d->mat[0]);
  double currValRMSE = rmseMasked(m, m->cur, d->mat[1]);
  if (m->keepHistory) {
    HistEntry h;
    h.U = m->cur.U;
    h.V = m->cur.V;
    h.obj = currObj;
    h.valRmse = currValRMSE;
    m->hist.push_back(std::move(h));
  }
  if (currObj != currObj || currValRMSE != currValRMSE) {
    if (m->cur.learnRate > 1e-5) {
      m->cur = m->best;
      m->cur.learnRate = m->cur.learnRate / 2;
      return false;
    } else {
      return true;
    }
  }
  if (currValRMSE < s.bestValRMSE) {
    m->best = m->cur;
    s.bestValRMSE = currValRMSE;
    s.bestIter = iter;
  }
  if (iter - s.bestIter >= 100) {
    if (m->cur.learnRate > 1e-5) m->cur.learnRate = m->cur.learnRate / 2;
  }
  if (iter - s.bestIter >= kChanceIter) ret = true;
  if (fabs(s.prevObj - currObj) < kEps) ret = true;
  s.prevObj = currObj;
  s.prevValRMSE = currValRMSE;
  return ret;
}

// util.cpp:1077-1107
static void sgdUpdateBlockSeq(int dim, std::vector<std::pair<int, int>> &updateSeq, std::mt19937 &mt) {
  updateSeq.clear();
  std::vector<bool> colMask(dim, false);
  std::vector<int> rowInds(dim);
  std::iota(rowInds.begin(), rowInds.end(), 0);
  std::shuffle(rowInds.begin(), rowInds.end(), mt);
  for (int ind = 0; ind < dim; ind++) {
    int currRow = rowInds[ind];
    std::vector<int> leftCols;
    for (int k = 0; k < dim; k++)
      if (!colMask[k]) leftCols.push_back(k);
    std::uniform_int_distribution<int> dis(0, leftCols.size() - 1);
    int currCol = leftCols[dis(mt)];
    updateSeq.push_back(std::make_pair(currRow, currCol));
    colMask[currCol] = true;
  }
}

// modelMF.cpp:233-265: part 0 receives perPart+1 ids (boundary quirk), the last part the rest.
static void makeParts(const std::vector<int> &ids, int P, std::vector<std::unordered_set<int>> &parts) {
  parts.assign(P, std::unordered_set<int>());
  int perPart = ids.size() / P;
  int currPart = 0;
  for (int i = 0; i < (int)ids.size(); i++) {
    parts[currPart].insert(ids[i]);
    if (i != 0 && perPart != 0 && i % perPart == 0) {
      if (currPart != P - 1) currPart++;
    }
  }
}

struct Triple {
  int u, item;
  float r;
};

// util.cpp:722-747
static std::vector<Triple> getUIRatings(const mfo_model *m, const Csr &tr) {
  std::vector<Triple> out;
  for (int u = 0; u < tr.nrows; u++) {
    if (m->invalidUsers.count(u)) continue;
    for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
      int item = tr.rowind[ii];
      if (m->invalidItems.count(item)) continue;
      out.push_back(Triple{u, item, tr.rowval[ii]});
    }
  }
  return out;
}

static void preamble(mfo_model *m, const mfo_data *d, StopState &s, std::vector<int> *trainUsers,
                     std::vector<int> *trainItems) {
  computeInvalid(m, d);
  std::vector<int> tu, ti;
  validIds(m, d->mat[0], tu, ti);
  if (m->algo == MFO_ALGO_IFWMF) initIfw(m, d->mat[0], tu, ti);
  // modelMF.cpp:48-50
  s.prevObj = objectiveMasked(m, m->cur, d->mat[0]);
  s.bestObj = s.prevObj;
  s.bestValRMSE = s.prevValRMSE = rmseMasked(m, m->cur, d->mat[1]);
  if (trainUsers) *trainUsers = tu;
  if (trainItems) *trainItems = ti;
}

// ModelMF::train modelMF.cpp:4-151; ModelInvPopMF::train modelInvPopMF.cpp:58-226
// (parBlockShuffle, util.cpp:1047-1064, degenerates to a full std::shuffle with one thread)
static int trainSerialSgd(mfo_model *m, const mfo_data *d) {
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  const int r = m->facDim;
  const bool ifw = m->algo == MFO_ALGO_IFWMF;
  std::mt19937 mt(m->seed);
  const auto uiRatings = getUIRatings(m, d->mat[0]);
  std::vector<size_t> inds(uiRatings.size());
  std::iota(inds.begin(), inds.end(), 0);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    std::shuffle(inds.begin(), inds.end(), mt);
    const float learnRate = m->cur.learnRate, uReg = m->uReg, iReg = m->iReg;
    for (const auto &ind : inds) {
      int u = uiRatings[ind].u, item = uiRatings[ind].item;
      float itemRat = uiRatings[ind].r;
      float *pu = &m->cur.U[(size_t)u * r], *pv = &m->cur.V[(size_t)item * r];
      float dotp = 0;
      for (int k = 0; k < r; k++) dotp += pu[k] * pv[k];
      double r_ui_est = dotp;
      double diff = itemRat - r_ui_est;
      if (ifw) {
        float wt = ifwWeight(m, u, item);
        for (int i = 0; i < r; i++) pu[i] -= learnRate * (-2.0 * wt * diff * pv[i] + 2.0 * uReg * pu[i]);
        for (int i = 0; i < r; i++) pv[i] -= learnRate * (-2.0 * wt * diff * pu[i] + 2.0 * iReg * pv[i]);
      } else {
        for (int i = 0; i < r; i++) pu[i] -= learnRate * (-2.0 * diff * pv[i] + 2.0 * uReg * pu[i]);
        for (int i = 0; i < r; i++) pv[i] -= learnRate * (-2.0 * diff * pu[i] + 2.0 * iReg * pv[i]);
      }
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// ModelMF::trainUShuffle modelMF.cpp:560-706 (--mf_method sgdu): the valid users in a fresh shuffled order every
// epoch (mt19937(trainSeed), :617,633), each user's ratings in CSR order (:637), the per-rating step of the serial
// trainer with diff / r_ui_est kept in double (:583,644-646).  The inner loop does not test invalidItems (an item
// without training ratings cannot occur in the training matrix).
static int trainUserShuffle(mfo_model *m, const mfo_data *d) {
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  const Csr &tr = d->mat[0];
  const int r = m->facDim;
  std::mt19937 mt(m->seed);
  std::vector<size_t> validUsers;
  for (int u = 0; u < m->nUsers; u++)
    if (!m->invalidUsers.count(u)) validUsers.push_back(u);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    std::shuffle(validUsers.begin(), validUsers.end(), mt);
    const float learnRate = m->cur.learnRate, uReg = m->uReg, iReg = m->iReg;
    for (const auto &u : validUsers) {
      for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
        const int item = tr.rowind[ii];
        const float itemRat = tr.rowval[ii];
        float *pu = &m->cur.U[(size_t)u * r], *pv = &m->cur.V[(size_t)item * r];
        float dotp = 0;
        for (int k = 0; k < r; k++) dotp += pu[k] * pv[k];
        const double r_ui_est = dotp;
        const double diff = itemRat - r_ui_est;
        for (int i = 0; i < r; i++) pu[i] -= learnRate * (-2.0 * diff * pv[i] + 2.0 * uReg * pu[i]);
        for (int i = 0; i < r; i++) pv[i] -= learnRate * (-2.0 * diff * pu[i] + 2.0 * iReg * pv[i]);
      }
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// ModelMF::hogTrain modelMF.cpp:1656-1810 executed by ONE thread (Eigen row expressions:
// foreign scalars are converted to float before they multiply a row, :1759-1762).
static int trainHogwildSerial(mfo_model *m, const mfo_data *d) {
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  const int r = m->facDim;
  std::mt19937 mt(m->seed);
  const auto uiRatings = getUIRatings(m, d->mat[0]);
  std::vector<size_t> inds(uiRatings.size());
  std::iota(inds.begin(), inds.end(), 0);
  std::vector<float> tmp(r);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    std::shuffle(inds.begin(), inds.end(), mt);
    const float learnRate = m->cur.learnRate, uReg = m->uReg, iReg = m->iReg;
    for (size_t kk = 0; kk < inds.size(); kk++) {
      const Triple &t = uiRatings[inds[kk]];
      float *pu = &m->cur.U[(size_t)t.u * r], *pv = &m->cur.V[(size_t)t.item * r];
      float dotp = 0;
      for (int k = 0; k < r; k++) dotp += pu[k] * pv[k];
      double r_ui_est = dotp;
      const double diff = t.r - r_ui_est;
      float a = static_cast<float>(-2.0 * diff), bu = static_cast<float>(2.0 * uReg),
            bi = static_cast<float>(2.0 * iReg);
      for (int k = 0; k < r; k++) tmp[k] = learnRate * (a * pv[k] + bu * pu[k]);
      for (int k = 0; k < r; k++) pu[k] -= tmp[k];
      for (int k = 0; k < r; k++) tmp[k] = learnRate * (a * pu[k] + bi * pv[k]);
      for (int k = 0; k < r; k++) pv[k] -= tmp[k];
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// Stratified SGD: ModelMF::trainSGDPar modelMF.cpp:154-350; ModelInvPopMF::trainSGDPar
// modelInvPopMF.cpp:229-442; ModelDropoutSigmoid::train modelDropoutSigmoid.cpp:26-246;
// ModelPoissonDropout::train modelPoissonDropout.cpp:50-288.
static int trainStratified(mfo_model *m, const mfo_data *d) {
  const Csr &tr = d->mat[0];
  const int r = m->facDim;
  const int P = m->nThreads;
  StopState s;
  std::vector<int> trainUsers, trainItems;
  preamble(m, d, s, &trainUsers, &trainItems);

  // modelPoissonDropout.cpp:118-123: one engine per thread, engine 0 also drives shuffles/schedule
  std::vector<std::mt19937> rEngines;
  for (int t = 0; t < P; t++) rEngines.push_back(std::mt19937(m->seed + t));
  std::mt19937 &mt = rEngines[0];  // == std::mt19937 mt(trainSeed) for the other trainers

  std::shuffle(trainUsers.begin(), trainUsers.end(), mt);
  std::shuffle(trainItems.begin(), trainItems.end(), mt);
  std::vector<std::unordered_set<int>> usersPart, itemsPart;
  makeParts(trainUsers, P, usersPart);
  makeParts(trainItems, P, itemsPart);
  std::vector<std::pair<int, int>> updateSeq;
  const int algo = m->algo;
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    const float learnRate = m->cur.learnRate, uReg = m->uReg, iReg = m->iReg;
    for (int k = 0; k < P; k++) {
      sgdUpdateBlockSeq(P, updateSeq, mt);
#pragma omp parallel for num_threads(P) schedule(static, 1)
      for (int t = 0; t < P; t++) {
        const auto &users = usersPart[updateSeq[t].first];
        const auto &items = itemsPart[updateSeq[t].second];
        for (const auto &u : users) {
          for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
            int item = tr.rowind[ii];
            if (items.count(item) == 0) continue;
            float *pu = &m->cur.U[(size_t)u * r], *pv = &m->cur.V[(size_t)item * r];
            int rank = r;
            if (algo == MFO_ALGO_TMF) {
              int updMinRank = tmfLambda(m, u, item);
              if (updMinRank < kEps) updMinRank = 1;
              if (updMinRank > r) updMinRank = r;
              rank = updMinRank;
            } else if (algo == MFO_ALGO_TMFDROPOUT) {
              int lambda = tmfLambda(m, u, item);
              std::poisson_distribution<> pdis(lambda);
              int updRank = pdis(rEngines[t]);
              if (updRank > r) updRank = r;
              if (updRank < kEps) updRank = 1;
              rank = updRank;
            }
            float itemRat = tr.rowval[ii];
            float r_ui_est = 0;
            for (int kk = 0; kk < rank; kk++) r_ui_est += pu[kk] * pv[kk];
            float diff = itemRat - r_ui_est;
            if (algo == MFO_ALGO_IFWMF) {
              float wt = ifwWeight(m, u, item);
              for (int i = 0; i < r; i++) pu[i] -= learnRate * (-2.0 * wt * diff * pv[i] + 2.0 * uReg * pu[i]);
              for (int i = 0; i < r; i++) pv[i] -= learnRate * (-2.0 * wt * diff * pu[i] + 2.0 * iReg * pv[i]);
            } else {
              for (int i = 0; i < rank; i++) pu[i] -= learnRate * (-2.0 * diff * pv[i] + 2.0 * uReg * pu[i]);
              for (int i = 0; i < rank; i++) pv[i] -= learnRate * (-2.0 * diff * pu[i] + 2.0 * iReg * pv[i]);
            }
          }
        }
      }
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// fp32 diagonally pivoted LDL^T of the lower triangle + solve: the algorithm Eigen::LDLT
// documents, as called at modelMF.cpp:836,874 (Eigen itself is absent; same arithmetic as the
// stand-in oracle/shim/Eigen/Dense that oracle/_ref is built with).
static void ldltSolve(int n, const float *Acolmajor, const float *b, float *xout) {
  std::vector<float> a(Acolmajor, Acolmajor + (size_t)n * n);
  auto at = [&](int i, int j) -> float & { return a[(size_t)j * n + i]; };
  for (int j = 0; j < n; j++)
    for (int i = j + 1; i < n; i++) at(j, i) = at(i, j);
  std::vector<int> perm(n);
  for (int k = 0; k < n; k++) {
    int piv = k;
    float best = std::abs(at(k, k));
    for (int i = k + 1; i < n; i++) {
      float v = std::abs(at(i, i));
      if (v > best) { best = v; piv = i; }
    }
    perm[k] = piv;
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(at(k, j), at(piv, j));
      for (int i = 0; i < n; i++) std::swap(at(i, k), at(i, piv));
    }
    float dd = at(k, k);
    if (dd == 0.0f) continue;
    for (int i = k + 1; i < n; i++) at(i, k) /= dd;
    for (int j = k + 1; j < n; j++) {
      float ljk_d = at(j, k) * dd;
      for (int i = j; i < n; i++) at(i, j) -= at(i, k) * ljk_d;
    }
    for (int j = k + 1; j < n; j++)
      for (int i = j + 1; i < n; i++) at(j, i) = at(i, j);
  }
  std::vector<float> x(b, b + n);
  for (int k = 0; k < n; k++) std::swap(x[k], x[perm[k]]);
  for (int j = 0; j < n; j++)
    for (int i = j + 1; i < n; i++) x[i] -= at(i, j) * x[j];
  const float tol = std::numeric_limits<float>::min();
  for (int i = 0; i < n; i++) {
    float dd = at(i, i);
    x[i] = (std::abs(dd) > tol) ? x[i] / dd : 0.0f;
  }
  for (int j = n - 1; j >= 0; j--)
    for (int i = j + 1; i < n; i++) x[j] -= at(i, j) * x[i];
  for (int k = n - 1; k >= 0; k--) std::swap(x[k], x[perm[k]]);
  for (int i = 0; i < n; i++) xout[i] = x[i];
}

extern "C" void mfo_ldlt_solve(int n, int r, const float *A, const float *b, float *x) {
  // A symmetric: row-major == column-major
  for (int i = 0; i < n; i++) ldltSolve(r, A + (size_t)i * r * r, b + (size_t)i * r, x + (size_t)i * r);
}

// ModelMF::trainALS modelMF.cpp:709-928
static int trainAls(mfo_model *m, const mfo_data *d) {
  const Csr &tr = d->mat[0];
  const int r = m->facDim;
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
#pragma omp parallel
    {
      std::vector<float> YTY((size_t)r * r), b(r), x(r);
#pragma omp for schedule(dynamic, 64)
      for (int u = 0; u < m->nUsers; u++) {
        if (m->invalidUsers.count(u) > 0) continue;
        std::fill(YTY.begin(), YTY.end(), 0.0f);
        std::fill(b.begin(), b.end(), 0.0f);
        for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
          const float *pv = &m->cur.V[(size_t)tr.rowind[ii] * r];
          float rating = tr.rowval[ii];
          if (rating > 0) {
            for (int j = 0; j < r; j++) {
              for (int k = 0; k < r; k++) YTY[(size_t)k * r + j] += pv[j] * pv[k];
              b[j] += rating * pv[j];
            }
          }
        }
        for (int j = 0; j < r; j++) YTY[(size_t)j * r + j] += m->uReg;
        ldltSolve(r, YTY.data(), b.data(), x.data());
        for (int j = 0; j < r; j++) m->u(u, j) = x[j];
      }
#pragma omp for schedule(dynamic, 16)
      for (int item = 0; item < m->nItems; item++) {
        if (m->invalidItems.count(item) > 0) continue;
        std::fill(YTY.begin(), YTY.end(), 0.0f);
        std::fill(b.begin(), b.end(), 0.0f);
        for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++) {
          const float *pu = &m->cur.U[(size_t)tr.colind[uu] * r];
          float rating = tr.colval[uu];
          if (rating > 0) {
            for (int j = 0; j < r; j++) {
              for (int k = 0; k < r; k++) YTY[(size_t)k * r + j] += pu[j] * pu[k];
              b[j] += rating * pu[j];
            }
          }
        }
        for (int j = 0; j < r; j++) YTY[(size_t)j * r + j] += m->iReg;
        ldltSolve(r, YTY.data(), b.data(), x.data());
        for (int j = 0; j < r; j++) m->v(item, j) = x[j];
      }
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// ModelMF::trainCCDPP modelMF.cpp:931-1169 and ::trainCCDPPFreqAdap :1172-1423
static int trainCcdpp(mfo_model *m, const mfo_data *d, bool freqAdap) {
  const Csr &tr = d->mat[0];
  const int r = m->facDim;
  const int nUsers = m->nUsers, nItems = m->nItems;
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  std::mt19937 mt(m->seed);
  std::vector<int> dims(r);
  std::iota(dims.begin(), dims.end(), 0);
  // residual copy: gk_csr_Dup(trainMat) — CSR and CSC value arrays both maintained
  std::vector<float> resRow(tr.rowval), resCol(tr.colval);
  std::fill(m->cur.U.begin(), m->cur.U.end(), 0.0f);  // modelMF.cpp:1020
  std::vector<float> u_k(nUsers), v_k(nItems);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    if (!freqAdap) std::shuffle(dims.begin(), dims.end(), mt);  // commented out at :1271
    for (const auto &k : dims) {
      for (int u = 0; u < nUsers; u++) u_k[u] = m->u(u, k);
      for (int i = 0; i < nItems; i++) v_k[i] = m->v(i, k);
      if (iter > 0) {
#pragma omp parallel for schedule(static)
        for (int u = 0; u < nUsers; u++) {
          if (m->invalidUsers.count(u) > 0) continue;
          for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++)
            resRow[ii] += m->u(u, k) * m->v(tr.rowind[ii], k);
        }
#pragma omp parallel for schedule(static)
        for (int item = 0; item < nItems; item++) {
          if (m->invalidItems.count(item) > 0 || item >= tr.ncols) continue;
          for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++)
            resCol[uu] += m->u(tr.colind[uu], k) * m->v(item, k);
        }
      }
      for (int subIter = 0; subIter < 5; subIter++) {
#pragma omp parallel for schedule(static)
        for (int u = 0; u < nUsers; u++) {
          if (m->invalidUsers.count(u) > 0) continue;
          double num = 0, denom = m->uReg, newV;
          for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
            int item = tr.rowind[ii];
            num += resRow[ii] * v_k[item];
            denom += v_k[item] * v_k[item];
          }
          newV = num / denom;
          u_k[u] = newV;
        }
#pragma omp parallel for schedule(static)
        for (int item = 0; item < nItems; item++) {
          if (m->invalidItems.count(item) > 0 || item >= tr.ncols) continue;
          double num = 0, denom = m->iReg, newV;
          for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++) {
            int u = tr.colind[uu];
            num += resCol[uu] * u_k[u];
            denom += u_k[u] * u_k[u];
          }
          newV = num / denom;
          v_k[item] = newV;
          if (freqAdap && m->itemFreq[item] < 75) {  // modelMF.cpp:1336-1342
            if (k > 0) v_k[item] = 0;
          }
        }
      }
#pragma omp parallel for schedule(static)
      for (int u = 0; u < nUsers; u++) {
        if (m->invalidUsers.count(u) > 0) continue;
        for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++)
          resRow[ii] -= u_k[u] * v_k[tr.rowind[ii]];
      }
#pragma omp parallel for schedule(static)
      for (int item = 0; item < nItems; item++) {
        if (m->invalidItems.count(item) > 0 || item >= tr.ncols) continue;
        for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++)
          resCol[uu] -= u_k[tr.colind[uu]] * v_k[item];
      }
      for (int u = 0; u < nUsers; u++) m->u(u, k) = u_k[u];
      for (int i = 0; i < nItems; i++) m->v(i, k) = v_k[i];
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// util.cpp:847-864 binSearch: position of key in sortedArr[lb..ub], -1 when absent
static int64_t binSearch(const std::vector<int32_t> &sortedArr, int key, int64_t ub, int64_t lb) {
  int64_t ind = -1;
  while (ub >= lb) {
    int64_t midP = (ub + lb) / 2;
    if (sortedArr[midP] == key) { ind = midP; break; }
    else if (sortedArr[midP] < key) lb = midP + 1;
    else ub = midP - 1;
  }
  return ind;
}

// ModelMF::trainCCD modelMF.cpp:1426-1653 (--mf_method ccd): cyclic coordinate descent one ROW at a time.  residual =
// gk_csr_Dup(trainMat) with both views kept (:1511); U = 0 (:1518-1523).  Per epoch every valid user visits its dims in
// a fresh std::shuffle of 0..r-1 drawn from the ONE mt19937(trainSeed) (:1495,1537-1538), num / denom / newV / upd in
// double over float products (:1541-1565), each residual entry patched in the other view through binSearch; then the
// items the same way over the CSC view (:1569-1606).  The reference draws the shuffles inside an OpenMP loop from the
// shared engine — defined only for one thread, which is what this restates (rows in index order).
static int trainCcd(mfo_model *m, const mfo_data *d) {
  const Csr &tr = d->mat[0];
  const int r = m->facDim;
  const int nUsers = m->nUsers, nItems = m->nItems;
  StopState s;
  preamble(m, d, s, nullptr, nullptr);
  std::mt19937 mt(m->seed);
  std::vector<int> dims(r);
  std::iota(dims.begin(), dims.end(), 0);
  std::vector<float> resRow(tr.rowval), resCol(tr.colval);
  std::fill(m->cur.U.begin(), m->cur.U.end(), 0.0f);
  int iter;
  for (iter = 0; iter < m->maxIter; iter++) {
    EpochTimer tm(m);
    for (int u = 0; u < nUsers; u++) {
      if (m->invalidUsers.count(u) > 0) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (const auto &k : udims) {
        double num = 0, denom = m->uReg, newV;
        for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
          int item = tr.rowind[ii];
          num += (resRow[ii] + m->u(u, k) * m->v(item, k)) * m->v(item, k);
          denom += m->v(item, k) * m->v(item, k);
        }
        newV = num / denom;
        for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) {
          int item = tr.rowind[ii];
          double upd = (newV - m->u(u, k)) * m->v(item, k);
          resRow[ii] -= upd;
          int64_t pos = binSearch(tr.colind, u, tr.colptr[item + 1] - 1, tr.colptr[item]);
          if (pos != -1) resCol[pos] -= upd;
        }
        m->u(u, k) = newV;
      }
    }
    for (int item = 0; item < nItems; item++) {
      if (m->invalidItems.count(item) > 0 || item >= tr.ncols) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (const auto &k : udims) {
        double num = 0, denom = m->iReg, newV;
        for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++) {
          int u = tr.colind[uu];
          num += (resCol[uu] + m->u(u, k) * m->v(item, k)) * m->u(u, k);
          denom += m->u(u, k) * m->u(u, k);
        }
        newV = num / denom;
        for (int64_t uu = tr.colptr[item]; uu < tr.colptr[item + 1]; uu++) {
          int u = tr.colind[uu];
          double upd = (newV - m->v(item, k)) * m->u(u, k);
          resCol[uu] -= upd;
          int64_t pos = binSearch(tr.rowind, item, tr.rowptr[u + 1] - 1, tr.rowptr[u]);
          if (pos != -1) resRow[pos] -= upd;
        }
        m->v(item, k) = newV;
      }
    }
    tm.stop();
    if (iter % kObjIter == 0 || iter == m->maxIter - 1)
      if (isTerminateModel(m, d, iter, s)) { iter++; break; }
  }
  return iter;
}

// the dims order every valid row of trainCCD draws (one shared mt19937(seed); per epoch the valid users in index order, then
// the valid items): out = [n_epochs][n_valid_users + n_valid_items][r]
extern "C" void mfo_ccd_dim_orders(const mfo_model *m, const mfo_data *d, int n_epochs, uint8_t *out) {
  const int r = m->facDim;
  std::mt19937 mt(m->seed);
  std::vector<int> dims(r);
  std::iota(dims.begin(), dims.end(), 0);
  size_t o = 0;
  for (int e = 0; e < n_epochs; e++) {
    for (int u = 0; u < m->nUsers; u++) {
      if (m->invalidUsers.count(u) > 0) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (int k = 0; k < r; k++) out[o++] = (uint8_t)udims[k];
    }
    for (int item = 0; item < m->nItems; item++) {
      if (m->invalidItems.count(item) > 0 || item >= d->mat[0].ncols) continue;
      std::vector<int> udims(dims);
      std::shuffle(udims.begin(), udims.end(), mt);
      for (int k = 0; k < r; k++) out[o++] = (uint8_t)udims[k];
    }
  }
}

extern "C" int mfo_train(mfo_model *m, const mfo_data *d, int method, int keep_history) {
  m->keepHistory = keep_history;
  m->hist.clear();
  m->epochSecs.clear();
  int saved = omp_get_max_threads();
  int iters = 0;
  // dispatch table of main.cpp:1325-1370: TMF / TMFDropout always run their stratified train()
  if (m->algo == MFO_ALGO_TMF || m->algo == MFO_ALGO_TMFDROPOUT) method = MFO_SGDPAR;
  switch (method) {
    case MFO_SGD: iters = trainSerialSgd(m, d); break;
    case MFO_SGDPAR: iters = trainStratified(m, d); break;
    case MFO_ALS: iters = trainAls(m, d); break;
    case MFO_CCDPP: iters = trainCcdpp(m, d, false); break;
    case MFO_CCDPP_FREQ: iters = trainCcdpp(m, d, true); break;
    case MFO_HOGWILD: iters = trainHogwildSerial(m, d); break;
    case MFO_SGDU: iters = trainUserShuffle(m, d); break;
    case MFO_CCD: iters = trainCcd(m, d); break;
    default: iters = -1;
  }
  omp_set_num_threads(saved);
  return iters;
}

extern "C" void mfo_get_factors(const mfo_model *m, int which, float *U, float *V) {
  const Facs &f = which ? m->best : m->cur;
  if (U) std::copy(f.U.begin(), f.U.end(), U);
  if (V) std::copy(f.V.begin(), f.V.end(), V);
}
extern "C" void mfo_set_factors(mfo_model *m, const float *U, const float *V) {
  if (U) std::copy(U, U + m->cur.U.size(), m->cur.U.begin());
  if (V) std::copy(V, V + m->cur.V.size(), m->cur.V.begin());
  m->best = m->cur;
}
extern "C" int mfo_history_len(const mfo_model *m) { return (int)m->hist.size(); }
extern "C" void mfo_get_history(const mfo_model *m, int epoch, float *U, float *V, double *objective,
                                double *val_rmse) {
  const HistEntry &h = m->hist[epoch];
  if (U) std::copy(h.U.begin(), h.U.end(), U);
  if (V) std::copy(h.V.begin(), h.V.end(), V);
  if (objective) *objective = h.obj;
  if (val_rmse) *val_rmse = h.valRmse;
}
extern "C" float mfo_learn_rate(const mfo_model *m) { return m->cur.learnRate; }
extern "C" void mfo_get_invalid(const mfo_model *m, uint8_t *users, uint8_t *items) {
  for (int u = 0; u < m->nUsers; u++) users[u] = m->invalidUsers.count(u) ? 1 : 0;
  for (int i = 0; i < m->nItems; i++) items[i] = m->invalidItems.count(i) ? 1 : 0;
}
extern "C" double mfo_rmse(const mfo_model *m, const mfo_data *d, int which, int best) {
  return rmseMasked(m, best ? m->best : m->cur, d->mat[which]);
}
extern "C" double mfo_objective(const mfo_model *m, const mfo_data *d) {
  return objectiveMasked(m, m->cur, d->mat[0]);
}

// ---- ranking metrics (model.cpp:760-1332) ------------------------------------------------------------------
// hitRate / arHR family (model.cpp:981-1332): the user's test item is the FIRST rating of its row in testMat; all items
// that are neither rated by the user in the training matrix nor invalid are scored with estRating and the N best kept
// in a heap ordered by descComp (util.cpp:760: a.second > b.second, i.e. the smallest kept score on top), sorted, and
// searched for the test item.  Returns its position or -1.
static int topNPosition(const mfo_model *m, const Facs &f, const Csr &tr, int u, int testItem, int N) {
  std::unordered_set<int> uTrItems;
  for (int64_t ii = tr.rowptr[u]; ii < tr.rowptr[u + 1]; ii++) uTrItems.insert(tr.rowind[ii]);
  auto descComp = [](const std::pair<int, double> &a, const std::pair<int, double> &b) { return a.second > b.second; };
  std::vector<std::pair<int, double>> topN;
  std::make_heap(topN.begin(), topN.end(), descComp);
  for (int item = 0; item < tr.ncols; item++) {
    if (uTrItems.count(item) || m->invalidItems.count(item) > 0) continue;
    topN.push_back(std::make_pair(item, estRating(m, f, u, item)));
    std::push_heap(topN.begin(), topN.end(), descComp);
    if ((int)topN.size() > N) {
      std::pop_heap(topN.begin(), topN.end(), descComp);
      topN.pop_back();
    }
  }
  std::sort(topN.begin(), topN.end(), descComp);
  for (size_t pos = 0; pos < topN.size(); pos++)
    if (testItem == topN[pos].first) return (int)pos;
  return -1;
}

// one user's term of NDCG / NDCGU / NDCGI (model.cpp:776-826): the N = 10 best predicted of the user's test ratings,
// DCG in prediction order over the ideal DCG of THOSE ratings; false when the user does not count
static bool ndcgUser(const mfo_model *m, const Facs &f, const Csr &te, int u, const uint8_t *filtItems, double *term) {
  const int N = 10;
  typedef std::tuple<int, float, float> Triplet;  // item, actual, predicted
  auto byPred = [](const Triplet &a, const Triplet &b) { return std::get<2>(a) > std::get<2>(b); };
  auto byAct = [](const Triplet &a, const Triplet &b) { return std::get<1>(a) > std::get<1>(b); };
  std::vector<Triplet> rs;
  for (int64_t ii = te.rowptr[u]; ii < te.rowptr[u + 1]; ii++) {
    int item = te.rowind[ii];
    if (m->invalidItems.count(item) > 0) continue;
    if (filtItems && !filtItems[item]) continue;
    float rat = te.rowval[ii];
    float pred = estRating(m, f, u, item);
    rs.push_back(Triplet(item, rat, pred));
    std::push_heap(rs.begin(), rs.end(), byPred);
    if ((int)rs.size() > N) {
      std::pop_heap(rs.begin(), rs.end(), byPred);
      rs.pop_back();
    }
  }
  if (rs.size() < 2) return false;
  std::sort(rs.begin(), rs.end(), byPred);
  float u_ndcg = 0.0;
  for (int i = 0; i < N && i < (int)rs.size(); i++) u_ndcg += (std::pow(2.0, std::get<1>(rs[i])) - 1) / std::log2((i + 1) + 1);
  std::sort(rs.begin(), rs.end(), byAct);
  float u_dcg_max = 0.0;
  for (int i = 0; i < N && i < (int)rs.size(); i++) u_dcg_max += (std::pow(2.0, std::get<1>(rs[i])) - 1) / std::log2((i + 1) + 1);
  if (!(u_dcg_max > kEps)) return false;
  *term = u_ndcg / u_dcg_max;
  return true;
}

// out[0..2] = hitRate, arHR, NDCG (model.cpp:1158, :981, :760);
// out[3..8] = hitRateU {first, second}, arHRU {first, second}, NDCGU {first, second} for filt_users (:1277, :1100, :833);
// out[9..14] = the I variants for filt_items (:1214, :1037, :907).  Filters: uint8 per id, 1 = in the set; NULL = skip.
extern "C" void mfo_rank_metrics(const mfo_model *m, const mfo_data *d, int which, int best, const uint8_t *filt_users,
                                 const uint8_t *filt_items, double out[15]) {
  const Facs &f = best ? m->best : m->cur;
  const Csr &tr = d->mat[0], &te = d->mat[which];
  for (int k = 0; k < 15; k++) out[k] = 0;
  // three passes of the hitRate family: no filter, users, items
  for (int pass = 0; pass < 3; pass++) {
    if ((pass == 1 && !filt_users) || (pass == 2 && !filt_items)) continue;
    double hits10 = 0, hits1000 = 0, nVal = 0;
    for (int u = 0; u < tr.nrows; u++) {
      if (m->invalidUsers.count(u) > 0) continue;
      if (pass == 1 && !filt_users[u]) continue;
      const int testItem = te.rowind[te.rowptr[u]];
      if (pass == 2 && !filt_items[testItem]) continue;
      if (topNPosition(m, f, tr, u, testItem, 10) >= 0) hits10 += 1;  // N = 10 (:1164)
      const int pos = topNPosition(m, f, tr, u, testItem, 1000);       // N = 1000 (:985)
      if (pos >= 0) hits1000 += 1.0 / (pos + 1);
      nVal += 1;
    }
    if (pass == 0) { out[0] = hits10 / nVal; out[1] = hits1000 / nVal; }
    else { double *o = out + (pass == 1 ? 3 : 9); o[0] = hits10; o[1] = hits10 / nVal; o[2] = hits1000; o[3] = hits1000 / nVal; }
  }
  for (int pass = 0; pass < 3; pass++) {
    if ((pass == 1 && !filt_users) || (pass == 2 && !filt_items)) continue;
    double ndcg = 0;
    int nVal = 0;
    for (int u = 0; u < te.nrows; u++) {
      if (m->invalidUsers.count(u) > 0) continue;
      if (pass == 1 && !filt_users[u]) continue;
      double term;
      if (!ndcgUser(m, f, te, u, pass == 2 ? filt_items : nullptr, &term)) continue;
      ndcg += term;
      nVal++;
    }
    if (pass == 0) out[2] = ndcg / nVal;
    else { double *o = out + (pass == 1 ? 7 : 13); o[0] = nVal; o[1] = ndcg / nVal; }
  }
}

extern "C" void mfo_dsgd_plan(const mfo_model *mc, const mfo_data *d, int P, int n_subepochs,
                              int32_t *user_part, int32_t *item_part, int32_t *schedule) {
  mfo_model *m = const_cast<mfo_model *>(mc);
  computeInvalid(m, d);
  std::vector<int> trainUsers, trainItems;
  validIds(m, d->mat[0], trainUsers, trainItems);
  std::mt19937 mt(m->seed);
  std::shuffle(trainUsers.begin(), trainUsers.end(), mt);
  std::shuffle(trainItems.begin(), trainItems.end(), mt);
  std::vector<std::unordered_set<int>> usersPart, itemsPart;
  makeParts(trainUsers, P, usersPart);
  makeParts(trainItems, P, itemsPart);
  for (int u = 0; u < m->nUsers; u++) user_part[u] = -1;
  for (int i = 0; i < m->nItems; i++) item_part[i] = -1;
  for (int p = 0; p < P; p++) {
    for (int u : usersPart[p]) user_part[u] = p;
    for (int i : itemsPart[p]) item_part[i] = p;
  }
  std::vector<std::pair<int, int>> seq;
  for (int s = 0; s < n_subepochs; s++) {
    sgdUpdateBlockSeq(P, seq, mt);
    for (int t = 0; t < P; t++) {
      schedule[((size_t)s * P + t) * 2 + 0] = seq[t].first;
      schedule[((size_t)s * P + t) * 2 + 1] = seq[t].second;
    }
  }
}

extern "C" void mfo_tmf_ranks(const mfo_model *m, int32_t *user_rank, int32_t *item_rank, int for_prediction) {
  // rank as a function of one side's frequency (the reference picks the side per rating)
  const int r = m->facDim;
  auto rankOf = [&](double freq) {
    double scaleFreq = (freq - m->meanFreq) / m->stdFreq;
    double sigmPc = 1.0 / (1.0 + exp(-m->rhoRMS * (scaleFreq - m->alpha)));
    int lambda = (int)std::ceil(sigmPc * ((double)r));
    if (m->algo == MFO_ALGO_TMFDROPOUT && for_prediction) {
      int k = m->cdfRanks[lambda - 1] + 1;
      return k > r ? r : k;
    }
    if (lambda < kEps) lambda = 1;
    if (lambda > r) lambda = r;
    return lambda;
  };
  for (size_t u = 0; u < m->userFreq.size(); u++) user_rank[u] = rankOf(m->userFreq[u]);
  for (size_t i = 0; i < m->itemFreq.size(); i++) item_rank[i] = rankOf(m->itemFreq[i]);
}

extern "C" void mfo_ifw_weights(const mfo_model *mc, const mfo_data *d, double *inv_pop_u, double *inv_pop_i) {
  mfo_model *m = const_cast<mfo_model *>(mc);
  computeInvalid(m, d);
  std::vector<int> tu, ti;
  validIds(m, d->mat[0], tu, ti);
  initIfw(m, d->mat[0], tu, ti);
  for (int u = 0; u < m->nUsers; u++) inv_pop_u[u] = m->invPopU[u];
  for (int i = 0; i < m->nItems; i++) inv_pop_i[i] = m->invPopI[i];
}

// dims visiting order of trainCCDPP: std::shuffle(dims, mt) once per epoch (modelMF.cpp:1000,1026)
extern "C" void mfo_ccdpp_dim_order(int seed, int r, int n_epochs, int32_t *out) {
  std::mt19937 mt(seed);
  std::vector<int> dims(r);
  std::iota(dims.begin(), dims.end(), 0);
  for (int e = 0; e < n_epochs; e++) {
    std::shuffle(dims.begin(), dims.end(), mt);
    for (int k = 0; k < r; k++) out[(size_t)e * r + k] = dims[k];
  }
}

extern "C" int mfo_epoch_seconds(const mfo_model *m, double *out, int cap) {
  int n = (int)m->epochSecs.size();
  for (int i = 0; i < n && i < cap; i++) out[i] = m->epochSecs[i];
  return n;
}
