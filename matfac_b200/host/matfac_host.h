/* C view of the host library (libmatfac_host.so) for callers that are not C++: one call that
 * runs a whole training job through the same classes as the `mf` binary.  All pointers are
 * host pointers; matrices are CSR with int64 row pointers. */
#ifndef MATFAC_HOST_H
#define MATFAC_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfh_csr {
  int32_t nrows;
  const int64_t *rowptr;
  const int32_t *rowind;
  const float *rowval;
} mfh_csr;

typedef struct mfh_problem {
  mfh_csr train, val, test;
  const char *algo;      /* mf | IFWMF | TMF | TMFDropout           (main.cpp:45) */
  const char *mf_method; /* sgd | sgdpar | hogsgd | als | ccd++ | ccdpp_plain | ccd | sgdu (main.cpp:44) */
  int32_t facdim, maxiter, seed;
  int32_t num_parts;     /* P of the stratified trainers; 0 = omp_get_max_threads() like the reference */
  float ureg, ireg, learnrate, rhorms, alpha;
  const float *init_U, *init_V; /* optional [n][facdim] starting factors (NULL = seeded init) */
  const char *prefix;
} mfh_problem;

typedef struct mfh_result {
  int32_t n_users, n_items;
  float learn_rate;
  double best_val_rmse, best_test_rmse, last_val_rmse, last_objective;
  float *last_U, *last_V, *best_U, *best_V; /* caller-allocated [n][facdim], may be NULL */
} mfh_result;

int mfh_train(const mfh_problem *p, mfh_result *out);
/* Host plan of the stratified trainers (Model::trainSGDPar and twins): reference partitions (modelMF.cpp:229-265) and
 * n_subepochs update sequences (util.cpp:1077-1107) drawn from mt19937(seed) in the reference's order.  invalid_* are
 * one byte per id (may be NULL); schedule_out = [n_subepochs][P][2] (may be NULL). */
int mfh_sgd_plan(int32_t n_users, int32_t n_items, const uint8_t *invalid_users, const uint8_t *invalid_items,
                 int32_t seed, int32_t P, int32_t n_subepochs, int32_t *user_part_out, int32_t *item_part_out,
                 int32_t *schedule_out);
/* Ranking metrics (model.cpp:760-1332) of the factor pair (U, V: [n][facdim]) through the host classes, against matrix
 * `which` (1 = val, 2 = test) of the problem; invalid_* / filt_*: one byte per id (may be NULL).  out[0..2] = hitRate,
 * arHR, NDCG; out[3..8] = hitRateU, arHRU, NDCGU as {first, second}; out[9..14] = hitRateI, arHRI, NDCGI. */
int mfh_rank_metrics(const mfh_problem *p, const float *U, const float *V, const uint8_t *invalid_users,
                     const uint8_t *invalid_items, int which, const uint8_t *filt_users, const uint8_t *filt_items,
                     double out[15]);
void mfh_release_device(void);

#ifdef __cplusplus
}
#endif
#endif
