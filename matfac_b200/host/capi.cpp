// C entry point over the host classes for callers that hold the matrices in memory (bench.py,
// the Python tests): builds a Data object from arrays, constructs the requested model, runs the
// requested trainer — the same objects and calls as `mf` — and copies results out.
#include <omp.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <random>

#include "device_session.h"
#include "matfac_host.h"
#include "modelDropoutSigmoid.h"
#include "modelInvPopMF.h"
#include "modelMF.h"
#include "modelPoissonDropout.h"

static gk_csr_t *fromArrays(const mfh_csr &m) { return gk_csr_FromArrays(m.nrows, m.rowptr, m.rowind, m.rowval); }

extern "C" int mfh_train(const mfh_problem *p, mfh_result *out) {
  if (!p || !out) return 1;
  std::string empty, trainName = "<memory>", prefix = p->prefix ? p->prefix : "/tmp/matfac";
  Params params(p->facdim, p->maxiter, 5, p->seed, p->ureg, p->ireg, p->learnrate, p->rhorms, p->alpha, trainName,
                trainName, trainName, empty, empty, empty, empty, empty, prefix);
  Data data(fromArrays(p->train), fromArrays(p->val), fromArrays(p->test), p->facdim, prefix.c_str());
  params.nUsers = data.nUsers;
  params.nItems = data.nItems;
  const int savedThreads = omp_get_max_threads();
  if (p->num_parts > 0) omp_set_num_threads(p->num_parts);  // P of the stratified trainers
  auto freq = getRowColFreq(data.trainMat);
  std::vector<double> userFreq = freq.first, itemFreq = freq.second, noRank;
  std::unique_ptr<Model> model, best;
  const std::string algo = p->algo ? p->algo : "mf", method = p->mf_method ? p->mf_method : "sgd";
  if (algo == "mf") {
    model.reset(new ModelMF(params, params.seed));
    best.reset(new ModelMF(params, params.seed));
  } else if (algo == "IFWMF") {
    model.reset(new ModelInvPopMF(params, params.seed, userFreq, itemFreq));
    best.reset(new ModelInvPopMF(params, params.seed, userFreq, itemFreq));
  } else if (algo == "TMF") {
    model.reset(new ModelDropoutSigmoid(params, params.seed, noRank, noRank, userFreq, itemFreq));
    best.reset(new ModelDropoutSigmoid(params, params.seed, noRank, noRank, userFreq, itemFreq));
  } else if (algo == "TMFDropout") {
    model.reset(new ModelPoissonDropout(params, params.seed, noRank, noRank, userFreq, itemFreq));
    best.reset(new ModelPoissonDropout(params, params.seed, noRank, noRank, userFreq, itemFreq));
  } else {
    return 2;
  }
  if (p->init_U && p->init_V) {
    memcpy(model->uFac.data(), p->init_U, sizeof(float) * (size_t)data.nUsers * p->facdim);
    memcpy(model->iFac.data(), p->init_V, sizeof(float) * (size_t)data.nItems * p->facdim);
    best->uFac = model->uFac;
    best->iFac = model->iFac;
  }
  std::unordered_set<int> invalidUsers, invalidItems;
  if (algo == "mf" && method == "ccd++") model->trainCCDPPFreqAdap(data, *best, invalidUsers, invalidItems);
  else if (algo == "mf" && method == "ccdpp_plain") model->trainCCDPP(data, *best, invalidUsers, invalidItems);
  else if (algo == "mf" && method == "ccd") model->trainCCD(data, *best, invalidUsers, invalidItems);
  else if (algo == "mf" && method == "als") model->trainALS(data, *best, invalidUsers, invalidItems);
  else if (algo == "mf" && method == "hogsgd") model->hogTrain(data, *best, invalidUsers, invalidItems);
  else if (algo == "mf" && method == "sgdu") model->trainUShuffle(data, *best, invalidUsers, invalidItems);
  else if ((algo == "mf" || algo == "IFWMF") && method == "sgdpar") model->trainSGDPar(data, *best, invalidUsers, invalidItems);
  else model->train(data, *best, invalidUsers, invalidItems);

  out->n_users = data.nUsers;
  out->n_items = data.nItems;
  out->learn_rate = model->learnRate;
  out->best_val_rmse = best->RMSE(data.valMat, invalidUsers, invalidItems);
  out->best_test_rmse = best->RMSE(data.testMat, invalidUsers, invalidItems);
  out->last_val_rmse = model->RMSE(data.valMat, invalidUsers, invalidItems);
  out->last_objective = model->objective(data, invalidUsers, invalidItems);
  const size_t ub = sizeof(float) * (size_t)data.nUsers * p->facdim, vb = sizeof(float) * (size_t)data.nItems * p->facdim;
  if (out->last_U) memcpy(out->last_U, model->uFac.data(), ub);
  if (out->last_V) memcpy(out->last_V, model->iFac.data(), vb);
  if (out->best_U) memcpy(out->best_U, best->uFac.data(), ub);
  if (out->best_V) memcpy(out->best_V, best->iFac.data(), vb);
  omp_set_num_threads(savedThreads);
  return 0;
}

// The stratified trainers' host plan, exactly as Model::runStratifiedSgd draws it: valid ids shuffled by
// mt19937(seed) (users, then items), cut into P parts with the reference's boundary rule (modelMF.cpp:229-265), then
// n_subepochs update sequences from the same engine (util.cpp:1077-1107).  invalid_* are one byte per id (1 = not
// trained, part -1).  schedule_out = [n_subepochs][P][2] (user part, item part).
extern "C" int mfh_sgd_plan(int32_t n_users, int32_t n_items, const uint8_t *invalid_users, const uint8_t *invalid_items,
                            int32_t seed, int32_t P, int32_t n_subepochs, int32_t *user_part_out, int32_t *item_part_out,
                            int32_t *schedule_out) {
  if (n_users <= 0 || n_items <= 0 || P < 1 || n_subepochs < 0 || !user_part_out || !item_part_out) return 1;
  std::vector<int> users, items;
  for (int u = 0; u < n_users; u++)
    if (!invalid_users || !invalid_users[u]) users.push_back(u);
  for (int i = 0; i < n_items; i++)
    if (!invalid_items || !invalid_items[i]) items.push_back(i);
  std::mt19937 mt(seed);
  std::shuffle(users.begin(), users.end(), mt);
  std::shuffle(items.begin(), items.end(), mt);
  const std::vector<int> up = matfac::partitionIds(users, P, n_users), ip = matfac::partitionIds(items, P, n_items);
  memcpy(user_part_out, up.data(), sizeof(int32_t) * (size_t)n_users);
  memcpy(item_part_out, ip.data(), sizeof(int32_t) * (size_t)n_items);
  std::vector<std::pair<int, int>> seq;
  for (int s = 0; s < n_subepochs && schedule_out; s++) {
    sgdUpdateBlockSeq(P, seq, mt);
    for (int t = 0; t < P; t++) {
      schedule_out[((size_t)s * P + t) * 2] = seq[t].first;
      schedule_out[((size_t)s * P + t) * 2 + 1] = seq[t].second;
    }
  }
  return 0;
}

// Ranking metrics of a given factor pair through the host classes (Model::hitRate / arHR / NDCG and their U / I
// variants): out[0..2] = hitRate, arHR, NDCG; out[3..8] = {first, second} of hitRateU, arHRU, NDCGU; out[9..14] = the I
// variants (zeros where the filter is NULL) — the layout of the oracle's mfo_rank_metrics.
extern "C" int mfh_rank_metrics(const mfh_problem *p, const float *U, const float *V, const uint8_t *invalid_users,
                                const uint8_t *invalid_items, int which, const uint8_t *filt_users, const uint8_t *filt_items,
                                double out[15]) {
  if (!p || !U || !V || !out || (which != 1 && which != 2)) return 1;
  std::string empty, trainName = "<memory>", prefix = p->prefix ? p->prefix : "/tmp/matfac";
  Params params(p->facdim, p->maxiter, 5, p->seed, p->ureg, p->ireg, p->learnrate, p->rhorms, p->alpha, trainName,
                trainName, trainName, empty, empty, empty, empty, empty, prefix);
  Data data(fromArrays(p->train), fromArrays(p->val), fromArrays(p->test), p->facdim, prefix.c_str());
  params.nUsers = data.nUsers;
  params.nItems = data.nItems;
  auto freq = getRowColFreq(data.trainMat);
  std::vector<double> userFreq = freq.first, itemFreq = freq.second, noRank;
  std::unique_ptr<Model> model;
  const std::string algo = p->algo ? p->algo : "mf";
  if (algo == "mf") model.reset(new ModelMF(params, params.seed));
  else if (algo == "IFWMF") model.reset(new ModelInvPopMF(params, params.seed, userFreq, itemFreq));
  else if (algo == "TMF") model.reset(new ModelDropoutSigmoid(params, params.seed, noRank, noRank, userFreq, itemFreq));
  else if (algo == "TMFDropout") model.reset(new ModelPoissonDropout(params, params.seed, noRank, noRank, userFreq, itemFreq));
  else return 2;
  memcpy(model->uFac.data(), U, sizeof(float) * (size_t)data.nUsers * p->facdim);
  memcpy(model->iFac.data(), V, sizeof(float) * (size_t)data.nItems * p->facdim);
  std::unordered_set<int> invalidUsers, invalidItems, fu, fi;
  for (int u = 0; u < data.nUsers; u++) {
    if (invalid_users && invalid_users[u]) invalidUsers.insert(u);
    if (filt_users && filt_users[u]) fu.insert(u);
  }
  for (int i = 0; i < data.nItems; i++) {
    if (invalid_items && invalid_items[i]) invalidItems.insert(i);
    if (filt_items && filt_items[i]) fi.insert(i);
  }
  gk_csr_t *mat = which == 1 ? data.valMat : data.testMat;
  for (int k = 0; k < 15; k++) out[k] = 0;
  out[0] = model->hitRate(data, invalidUsers, invalidItems, mat);
  out[1] = model->arHR(data, invalidUsers, invalidItems, mat);
  out[2] = model->NDCG(invalidUsers, invalidItems, mat);
  if (filt_users) {
    auto a = model->hitRateU(data, fu, invalidUsers, invalidItems, mat);
    auto b = model->arHRU(data, fu, invalidUsers, invalidItems, mat);
    auto c = model->NDCGU(fu, invalidUsers, invalidItems, mat);
    out[3] = a.first; out[4] = a.second; out[5] = b.first; out[6] = b.second; out[7] = c.first; out[8] = c.second;
  }
  if (filt_items) {
    auto a = model->hitRateI(data, fi, invalidUsers, invalidItems, mat);
    auto b = model->arHRI(data, fi, invalidUsers, invalidItems, mat);
    auto c = model->NDCGI(fi, invalidUsers, invalidItems, mat);
    out[9] = a.first; out[10] = a.second; out[11] = b.first; out[12] = b.second; out[13] = c.first; out[14] = c.second;
  }
  return 0;
}

extern "C" void mfh_release_device(void) { matfac::DeviceSession::dropAll(); }
