/*
 * TEST INFRASTRUCTURE — C API of the CPU restatement of mohit-shrma/matfac's training hot
 * path (oracle/mf_oracle.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path never does.
 *
 * Parity pinning: the reference ships no golden vectors.  The restatement is
 * pinned against the reference's own translation units compiled unmodified into
 * oracle/_ref/mf_ref (see oracle/Makefile, oracle/ref_driver.cpp) — tests/test_oracle_vs_ref.py
 * runs both here, and tests/golden/ holds the vectors that run produced so the GPU box
 * (where /root/reference does not exist) can re-check the oracle.
 */
#ifndef MF_ORACLE_H
#define MF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfo_data mfo_data;
typedef struct mfo_model mfo_model;

enum mfo_algo { MFO_ALGO_MF = 0, MFO_ALGO_IFWMF = 1, MFO_ALGO_TMF = 2, MFO_ALGO_TMFDROPOUT = 3 };
enum mfo_method {
  MFO_SGD = 0,          /* ModelMF::train / ModelInvPopMF::train (serial SGD)            */
  MFO_SGDPAR = 1,       /* trainSGDPar and the TMF / TMFDropout train() (stratified SGD) */
  MFO_ALS = 2,          /* ModelMF::trainALS                                             */
  MFO_CCDPP = 3,        /* ModelMF::trainCCDPP                                           */
  MFO_CCDPP_FREQ = 4,   /* ModelMF::trainCCDPPFreqAdap (what --mf_method ccd++ runs)     */
  MFO_HOGWILD = 5,      /* ModelMF::hogTrain executed by one thread                      */
  MFO_SGDU = 6,         /* ModelMF::trainUShuffle (--mf_method sgdu, modelMF.cpp:560-706) */
  MFO_CCD = 7           /* ModelMF::trainCCD (--mf_method ccd, modelMF.cpp:1426-1653) by one thread */
};

/* --- data (datastruct.cpp:3-120) ------------------------------------------------------ */
/* Parse a text CSR file the way gk_csr_Read(..., GK_CSR_FMT_CSR, 1, 0) does. Returns NULL on error. */
mfo_data *mfo_data_read(const char *train, const char *val, const char *test);
/* Build from arrays (rowptr int64[nrows+1], rowind int32[nnz], rowval float[nnz]); ncols = max index + 1. */
mfo_data *mfo_data_from_arrays(int64_t tr_rows, const int64_t *tr_ptr, const int32_t *tr_ind,
                               const float *tr_val, int64_t va_rows, const int64_t *va_ptr,
                               const int32_t *va_ind, const float *va_val, int64_t te_rows,
                               const int64_t *te_ptr, const int32_t *te_ind, const float *te_val);
void mfo_data_free(mfo_data *d);
int mfo_data_nusers(const mfo_data *d);
int mfo_data_nitems(const mfo_data *d);
/* which: 0 train, 1 val, 2 test.  dims = {nrows, ncols, nnz}. */
void mfo_data_dims(const mfo_data *d, int which, int64_t dims[3]);
void mfo_data_csr(const mfo_data *d, int which, int64_t *rowptr, int32_t *rowind, float *rowval);
void mfo_data_csc(const mfo_data *d, int which, int64_t *colptr, int32_t *colind, float *colval);

/* --- model ----------------------------------------------------------------------------- */
typedef struct mfo_params {
  int facDim, maxIter, seed, nThreads; /* nThreads plays omp_get_max_threads() */
  float uReg, iReg, learnRate, rhoRMS, alpha;
} mfo_params;

mfo_model *mfo_model_create(const mfo_data *d, const mfo_params *p, int algo);
void mfo_model_free(mfo_model *m);
/* Run the trainer exactly as the reference's method would (early stopping included).
 * keep_history != 0 stores the factors after every epoch. Returns epochs executed. */
int mfo_train(mfo_model *m, const mfo_data *d, int method, int keep_history);

/* which: 0 current (last epoch), 1 best-validation; row-major [n][facDim]. */
void mfo_get_factors(const mfo_model *m, int which, float *U, float *V);
void mfo_set_factors(mfo_model *m, const float *U, const float *V);
int mfo_history_len(const mfo_model *m);
void mfo_get_history(const mfo_model *m, int epoch, float *U, float *V, double *objective,
                     double *val_rmse);
float mfo_learn_rate(const mfo_model *m);
/* wall seconds of each epoch body of the last mfo_train (the reference's subIterDuration); returns count */
int mfo_epoch_seconds(const mfo_model *m, double *out, int cap);
/* invalid masks as filled by the trainer (uint8 per id, 1 = invalid). */
void mfo_get_invalid(const mfo_model *m, uint8_t *users, uint8_t *items);
void mfo_compute_invalid(mfo_model *m, const mfo_data *d);

/* masked RMSE (model.cpp:214-251) through the model's virtual estRating; which: 0/1/2; best: use best model */
double mfo_rmse(const mfo_model *m, const mfo_data *d, int which, int best);
/* objective (model.cpp:1770-1815; IFWMF modelInvPopMF.cpp:3-55) */
double mfo_objective(const mfo_model *m, const mfo_data *d);

/* ranking metrics (model.cpp:760-1332) of the current (best = 0) or best model against matrix `which` (1 val, 2 test);
 * every valid user must hold at least one rating in that matrix (the reference reads the first one unconditionally).
 * out[0..2] = hitRate, arHR, NDCG; out[3..8] = hitRateU, arHRU, NDCGU as {first, second} pairs for filt_users;
 * out[9..14] = hitRateI, arHRI, NDCGI for filt_items (uint8 per id, 1 = in the filter set; NULL skips the variants) */
void mfo_rank_metrics(const mfo_model *m, const mfo_data *d, int which, int best, const uint8_t *filt_users,
                      const uint8_t *filt_items, double out[15]);

/* DSGD bookkeeping exposed for bit-exact checks of the host-side schedule code:
 * user_part/item_part: part id per id, -1 for invalid ids; schedule: n_subepochs * P pairs (row,col)
 * drawn from the same mt19937 stream the trainer uses (shuffles first, then schedules). */
void mfo_dsgd_plan(const mfo_model *m, const mfo_data *d, int P, int n_subepochs,
                   int32_t *user_part, int32_t *item_part, int32_t *schedule /* [n][P][2] */);
/* per-user / per-item auxiliaries as the reference derives them */
void mfo_tmf_ranks(const mfo_model *m, int32_t *user_rank, int32_t *item_rank, int for_prediction);
/* normalised popularity scores invPopU / invPopI (modelInvPopMF.cpp:98-114), 0 for invalid ids */
void mfo_ifw_weights(const mfo_model *m, const mfo_data *d, double *inv_pop_u, double *inv_pop_i);
/* dims order of trainCCDPP for n_epochs epochs: out[n_epochs][r] */
void mfo_ccdpp_dim_order(int seed, int r, int n_epochs, int32_t *out);
/* trainCCD: the shuffled dims of every valid row, [n_epochs][valid users then valid items][r] (modelMF.cpp:1537,1577) */
void mfo_ccd_dim_orders(const mfo_model *m, const mfo_data *d, int n_epochs, uint8_t *out);
/* batched fp32 solve used by ALS (pivoted LDL^T), for unit checks: A is [n][r][r] row-major */
void mfo_ldlt_solve(int n, int r, const float *A, const float *b, float *x);

#ifdef __cplusplus
}
#endif
#endif
