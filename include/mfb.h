/*
 * mfb.h — C ABI of the B200 training engine for mohit-shrma/matfac's hot path.
 *
 * The reference has no FFI layer: its boundary for this path is the C++ class API
 * (model.h:56-216) called from main.cpp:1325-1382.  This header is the thin C boundary that
 * sits directly under those method bodies; matfac_b200/host/ holds the C++ classes
 * (Params / Data / Model / ModelMF / ModelInvPopMF / ModelDropoutSigmoid /
 * ModelPoissonDropout) that keep the reference's signatures and call into this library.
 * Each entry point names the reference statements it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer borrowed for the duration of the
 *     call unless the name says "device"; the engine owns all device memory;
 *   - every function returns 0 on success, non-zero on error with mfb_last_error() set;
 *     CUDA errors are sticky and fatal for the engine (the C++ wrapper prints to stderr and
 *     exit(-1)s, the reference's convention: model.cpp:1482-1483, main.cpp:62-64);
 *   - an engine is bound to one CUDA device and one stream; calls are asynchronous on that
 *     stream unless they return host data; not re-entrant (the reference's Model methods
 *     are not either);
 *   - there is NO CPU fallback: without a CUDA device mfb_create fails.
 *
 * Factor matrices cross the ABI row-major, [n][rank] fp32 with a leading dimension in
 * floats; (row = user|item, col = latent dim) as in Model::uFac/iFac (model.h:37-38).
 */
#ifndef MFB_H
#define MFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfb_engine mfb_engine;

/* which rating matrix (Data::trainMat / valMat / testMat, datastruct.h:77-79) */
enum { MFB_TRAIN = 0, MFB_VAL = 1, MFB_TEST = 2 };
/* which factor set: the model being trained, or the best-validation snapshot (bestModel) */
enum { MFB_CURRENT = 0, MFB_BEST = 1 };
/* user side / item side */
enum { MFB_USER = 0, MFB_ITEM = 1 };
/* update / prediction rule */
enum {
  MFB_MF = 0,         /* ModelMF             modelMF.cpp:83-105,275-303     */
  MFB_IFWMF = 1,      /* ModelInvPopMF       modelInvPopMF.cpp:152-180,356-391 */
  MFB_TMF = 2,        /* ModelDropoutSigmoid modelDropoutSigmoid.cpp:145-194 */
  MFB_TMFDROPOUT = 3  /* ModelPoissonDropout modelPoissonDropout.cpp:176-229 */
};

typedef struct mfb_config {
  int32_t device;   /* CUDA device ordinal */
  int32_t n_users;  /* Data::nUsers  (datastruct.cpp:23) */
  int32_t n_items;  /* Data::nItems  (datastruct.cpp:91) */
  int32_t rank;     /* Params::facDim */
  int32_t reserved[4];
} mfb_config;

const char *mfb_last_error(void);
/* number of kernels launched by all engines of this process since load (bench.py "gpu_launches") */
uint64_t mfb_launch_count(void);

/* number of visible CUDA devices (0 when there is none or the driver is missing) */
int32_t mfb_device_count(void);
int mfb_create(const mfb_config *cfg, mfb_engine **out);
void mfb_destroy(mfb_engine *e);
int mfb_sync(mfb_engine *e);

/* Page-lock a host range so uploads run at full PCIe rate (optional). */
int mfb_pin_host(void *ptr, uint64_t bytes);
int mfb_unpin_host(void *ptr);

/* ---- data (replaces the host-resident gk_csr_t of datastruct.cpp:16-18,49-51,72-74) ------
 * CSR is required; CSC (gk_csr_CreateIndex(…, GK_CSR_COL)) is required for MFB_TRAIN when
 * ALS or CCD++ will run, and may be NULL otherwise.  rowptr/colptr are int64 (ssize_t in
 * gk_csr_t), indices int32, values fp32.  nrows may be smaller than n_users for val/test
 * files that stop early; rows beyond nrows are empty. */
int mfb_upload_csr(mfb_engine *e, int which, int32_t nrows, int32_t ncols, int64_t nnz,
                   const int64_t *rowptr, const int32_t *rowind, const float *rowval,
                   const int64_t *colptr, const int32_t *colind, const float *colval);

/* Build the column index of an uploaded matrix on the device: gk_csr_CreateIndex(mat, GK_CSR_COL)
 * (datastruct.cpp:18,51,74) as a stable radix sort by column — colind ascends inside every column, exactly
 * the arrays the reference's counting sort produces.  n_items + 1 pointers (columns beyond ncols are empty).
 * mfb_download_csc copies them to the host (colptr int64 [n_items + 1], colind / colval [nnz]). */
int mfb_build_csc(mfb_engine *e, int which);
int mfb_download_csc(mfb_engine *e, int which, int64_t *colptr, int32_t *colind, float *colval);

/* invalidUsers / invalidItems (util.cpp:511-544 + modelMF.cpp:40-45) as one byte per id
 * (1 = invalid), n_users resp. n_items entries.  Applied by mfb_eval, ALS and CCD++. */
int mfb_set_masks(mfb_engine *e, const uint8_t *invalid_users, const uint8_t *invalid_items);

/* Model::uFac / iFac in, e.g. the seeded initialisation of model.cpp:2331-2350. */
int mfb_upload_factors(mfb_engine *e, const float *U, int64_t ldU, const float *V, int64_t ldV);
/* which = MFB_CURRENT (*this) or MFB_BEST (bestModel); either pointer may be NULL. */
int mfb_download_factors(mfb_engine *e, int which, float *U, int64_t ldU, float *V, int64_t ldV);

/* Per-user / per-item auxiliaries of the frequency-aware models, all n_users / n_items long:
 *   freq        userFreq / itemFreq (util.cpp:555-569) as int32
 *   train       MFB_IFWMF: fp32 weight 1/(1+rho*invPop) of that side (modelInvPopMF.cpp:163-168);
 *               MFB_TMF: int32 update rank of that side (modelDropoutSigmoid.cpp:158-170);
 *               MFB_TMFDROPOUT: int32 Poisson mean lambda (modelPoissonDropout.cpp:189-196)
 *   pred        int32 prediction rank of that side used by estRating
 *               (modelDropoutSigmoid.cpp:5-24; modelPoissonDropout.cpp:5-23); ignored for IFWMF
 * The side is picked per rating exactly as the reference does: the user's entry when
 * userFreq[u] < itemFreq[i] (TMF) resp. itemFreq[i] > userFreq[u] (IFWMF), else the item's.
 * variant = MFB_MF stores the frequencies only (the FreqAdap rule of CCD++ needs itemFreq).
 * poisson_cdf (MFB_TMFDROPOUT only): [rank][rank] fp32, row l-1 = P(Poisson(l) <= k), k=0..rank-1. */
int mfb_set_aux(mfb_engine *e, int variant, const int32_t *user_freq, const int32_t *item_freq,
                const void *user_train, const void *item_train, const int32_t *user_pred,
                const int32_t *item_pred, const float *poisson_cdf);

/* ---- SGD (modelMF.cpp:4-151 train, :154-350 trainSGDPar, :1656-1810 hogTrain and the
 *      IFWMF / TMF / TMF+Dropout trainers) --------------------------------------------------
 * mfb_sgd_plan buckets the valid training ratings into the P x P stratum grid of
 * modelMF.cpp:233-265 once (the reference rescans every rating in every sub-epoch, :283).
 * user_part / item_part give the part of every id (-1 = not trained); P = 1 with NULL
 * arrays makes the whole matrix one block (serial-SGD / Hogwild trainers).
 * Inside a block the visiting order is the reference's: user-major, CSR order inside a row
 * (modelMF.cpp:279-281); one sub-warp owns a user's run of ratings and keeps u in registers. */
int mfb_sgd_plan(mfb_engine *e, int32_t P, const int32_t *user_part, const int32_t *item_part);

/* One sub-epoch: the nb conflict-free blocks (user_part, item_part) of one updateSeq
 * (util.cpp:1077-1107) run concurrently.  blocks = nb pairs.  seed/counter feed the
 * counter-based Poisson draws of MFB_TMFDROPOUT. */
int mfb_sgd_subepoch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float learn_rate,
                     float ureg, float ireg, uint64_t seed, uint64_t counter);
/* One epoch of the serial / Hogwild trainers (modelMF.cpp:83-105 train, :1747-1763 hogTrain;
 * modelInvPopMF.cpp:152-180): every valid rating once, in a fresh pseudo-random order per
 * (seed, counter); a sub-warp owns one rating at a time, both factor rows are read with 128-bit
 * loads and updated with vector reductions.  Needs mfb_sgd_plan(e, 1, NULL, NULL). */
int mfb_sgd_epoch_flat(mfb_engine *e, int variant, float learn_rate, float ureg, float ireg, uint64_t seed,
                       uint64_t counter);
/* Tuning knobs: "sgd_shuffle_seed" (key of the plan-time physical shuffle of the rating records; the host trainers
 * pass Model::trainSeed before mfb_sgd_plan, as the reference seeds its shuffles with mt19937(trainSeed),
 * modelMF.cpp:63,78), "sgd_workers" (concurrent sub-warps, 0 = automatic), "sgd_warps_per_sm",
 * "sgd_max_hot_inflight" (bound on concurrent updates of the hottest item row, default 8),
 * "sgd_flat_hot_lr" (shuffled kernel: the hottest row's concurrency is capped at value / learn_rate,
 * default 0.15), "sgd_flat_inflight_frac" (shuffled kernel: ratings in flight <= this fraction of
 * the epoch, default 2e-4), "sgd_flat_inflight_steady" (the same bound once the user rows have stopped growing — their
 * rating-weighted mean squared norm within 0.8 .. 1.25 of its value at the previous launch; whole-matrix plans; default
 * 1e-3, set it to sgd_flat_inflight_frac to switch the relaxation off), "sgd_flat_launch_lr" (shuffled kernel: ratings in flight <= value / learn_rate x the
 * ratings of the launch, default 1.2e-5: a short launch — one stratum block — must not be in flight all at once),
 * "sgd_flat_band_mb" (shuffled kernel, whole-matrix plans: the epoch visits
 * the ratings in bands of users whose rows take this many MB, so that a band of U stays in L2; default
 * 0 = one band, a uniformly shuffled epoch as in modelMF.cpp:76-81 (banding converges at a different rate
 * per epoch than the reference's order); takes effect at the next mfb_sgd_plan), "ccd_fuse" (CCD++: 1 = the residual add-back rides on the first u_k / v_k update pass and the column subtract on the
 * last v_k update pass of a rank-one step — same statements in the same order, 11 instead of 14 passes; 0 = one pass each),
 * "ccd_stream" (CCD++ passes: 0 = one warp per row segment, segments sorted by length, the default; 1 = the same kernels over
 * the segments in memory order; 2 = warp-streamed chunks of consecutive segments, the next batch of ratings always in flight;
 * "ccd_cap" = ratings per chunk, default 4096; "ccd_stage" = 1 keeps the gathered vector in shared memory when it fits —
 * all measured equal or slower than the default, profiles/r2_ccdpp.md; the plans are rebuilt at the next mfb_ccdpp_begin),
 * "copy_overlap" (1 = mfb_upload_csr / mfb_upload_factors / mfb_download_factors return without synchronising:
 * factor copies run on a copy stream — an upload behind the rating upload and next to mfb_sgd_plan, a download next to
 * the evaluations queued after it; the caller keeps the host buffers (pinned) valid and untouched until mfb_sync; every
 * call that touches the factors is ordered behind a pending copy on the device; default 0 = copies complete before the
 * call returns),
 * "als_tensor_cores" (rank > 32: 1 = tcgen05 3xTF32 Gram in the warp-specialised persistent kernel, default; 0 = fp32
 * CUDA-core Gram; 2 = one CTA per row, rank > 64 only), "als_ws_split" (warp-specialised kernel: 0 = the split between
 * converter teams and solver groups is picked per half-step from the mean row length, default; 1 = the
 * many-short-rows split, 2 = the few-long-rows split), "als_dual" (1 = rows with fewer ratings than half the padded rank are solved through the len x len
 * dual system F (F F^T + reg I)^-1 r, default; 0 = always the rank x rank normal equations), "als_chunk" (ratings
 * one CTA accumulates before a row is split over several CTAs that add into a workspace, default 16384), "sgd_block_order" (stratified trainers: 0 = user-major runs, 1 = shuffled inside the
 * blocks), "sgd_atomic" (1 = item rows updated by reductions, 0 = plain stores), "sgd_rotate" (1 = every
 * user run of the stratified kernel starts at a pseudo-random offset and wraps around, 0 = CSR
 * order as in modelMF.cpp:280). */
int mfb_set_option(mfb_engine *e, const char *name, double value);
/* number of ratings the given blocks hold (for updates/s accounting) */
int mfb_sgd_block_nnz(mfb_engine *e, const int32_t *blocks, int32_t nb, int64_t *nnz);
/* Diagnostics: the rating records of one stratum block as the shuffled kernel visits them.  records =
 * [mfb_sgd_block_nnz of the block][4] int32 {user, item, rating bits, 0}: first *cold_records records in the
 * block's shuffled order, then one contiguous list per hot item (options "sgd_hot", "sgd_hot_min_count",
 * "sgd_hot_inflight", "sgd_hot_max_lists": an item of which the shuffled kernel would keep more than sgd_hot_inflight
 * (default 16) updates in flight when it runs the block on the whole machine, and that holds at least
 * sgd_hot_min_count (default 1024) of the block's ratings, is trained by one CTA that keeps the item row in shared
 * memory — the sgd_hot_max_lists (default and maximum 127) most rated ones when there are more; "sgd_hot_batch" ratings per mini-batch round of such a CTA, 0 = automatic; "sgd_hot_stages" 8 (default) or 4 rounds
 * of user rows staged ahead; "sgd_hot_pace" 1 = a list advances in step with the shuffled kernel (default); "sgd_hot_stab"
 * see mfb_debug_sgd_hot_batch).
 * lists = [*n_lists][3] int32 {item, first record relative to the block, records}; lists may be NULL or hold
 * sgd_hot_max_lists (<= 127) entries. */
int mfb_debug_sgd_records(mfb_engine *e, int32_t user_part, int32_t item_part, int32_t *records,
                          int64_t *cold_records, int32_t *lists, int32_t *n_lists);
/* Diagnostics: out = {sum over users of degree x |u|^2, sum of degrees, ratings per round the hot CTAs last used}.
 * The batch of a hot CTA is bounded on the device by sgd_hot_stab / (learn_rate x out[0] / out[1]) (option
 * "sgd_hot_stab", default 0.5: a mini-batch of T ratings of one item is only stable while learn_rate x T x |u|^2 < 1)
 * and on the host by 64 and by sgd_flat_hot_lr / learn_rate; out[2] is zero when the plan has no hot lists. */
int mfb_debug_sgd_hot_batch(mfb_engine *e, double out[3]);

/* ---- ALS (modelMF.cpp:795-882) -------------------------------------------------------------
 * side = MFB_USER: for every valid user solve (sum_{i in row, r>0} v v^T + reg I) x = sum r v
 * over the train CSR and overwrite U; MFB_ITEM: same over the CSC with the current U. */
int mfb_als_half_step(mfb_engine *e, int side, float reg);

/* Diagnostics: the Gram matrix and right-hand side of ONE row as the production ALS kernels form
 * them (no regulariser).  out = [R*R + R] floats with R = *padded_rank (64 or 128; rank > 32 only). */
int mfb_debug_als_gram(mfb_engine *e, int side, int32_t row, float *out, int32_t *padded_rank);

/* Diagnostics: the batched rank-64 solver of the ALS half-step (csrc/als_mn.cu) on n host records.  A record is
 * 2440 floats: the lower triangle of the Gram matrix row by row — row r holds r / 4 + 1 units of four floats and starts at
 * the first unit at or after the end of row r - 1 whose index mod 8 is not taken by an earlier row of its group of
 * eight rows (a bank-conflict-free placement; 594 units) — then the 64 floats of the right-hand side.  x = [n][64]: the
 * solutions of (G + reg I) x = b with rows / columns >= rank replaced by the identity. */
int mfb_debug_chol64(mfb_engine *e, int32_t n, const float *records, float *x, int32_t rank, float reg);

/* ---- CCD++ (modelMF.cpp:1013-1121 trainCCDPP; :1258-1375 trainCCDPPFreqAdap) ---------------
 * begin: residual := train values (gk_csr_Dup), U := 0 (:1020).  rank1: one pass of the k loop
 * body (:1028-1120): add-back unless first_iter, `inner` alternations of the u_k / v_k
 * closed-form updates, subtract, write column k.  item_freq_thresh > 0 applies the FreqAdap
 * rule (:1336-1342): v_k(i) = 0 when itemFreq[i] < thresh and k > 0.  end: frees the residual. */
int mfb_ccdpp_begin(mfb_engine *e);
int mfb_ccdpp_rank1(mfb_engine *e, int32_t k, int first_iter, int32_t inner, float ureg, float ireg,
                    int32_t item_freq_thresh);
int mfb_ccdpp_end(mfb_engine *e);

/* ---- CCD (modelMF.cpp:1426-1653 trainCCD, --mf_method ccd) ---------------------------------
 * Between mfb_ccdpp_begin and mfb_ccdpp_end (same residual state: ratings in both views, U = 0).
 * One half of an epoch: side = MFB_USER: every valid user visits its dims in the order
 * dim_order[u][0..rank) and sets u_k = sum((res + u_k v_k) v_k) / (reg + sum v_k^2) over its row,
 * then subtracts (new - old) v_k from the row's residuals (:1527-1566); MFB_ITEM: the same over
 * the columns (:1569-1606).  dim_order = host [n_users | n_items][rank] bytes (the reference draws
 * one std::shuffle per valid row from a single mt19937(trainSeed), :1537,1577; entries of invalid
 * rows are ignored), NULL = 0 .. rank-1 for every row.  The train matrix needs both views with
 * sorted indices (the residual of the other view is patched through the reference's binary
 * search, util.cpp:847).  One engine only (rows of the two sides exchange residuals). */
int mfb_ccd_half_step(mfb_engine *e, int side, float reg, const uint8_t *dim_order);

/* ---- evaluation (model.cpp:214-251 RMSE, :1770-1815 objective, modelInvPopMF.cpp:3-55) ----
 * One fused pass over `which`: out[0] = sum of (weighted) squared errors over ratings whose
 * user and item are valid, out[1] = their count, out[2] = sum_u |U_u|^2 and out[3] =
 * sum_i |V_i|^2 over valid ids (the two norm terms are computed only when want_norms != 0).
 * variant selects estRating (MFB_TMF / MFB_TMFDROPOUT truncate) and, for MFB_IFWMF with
 * weighted != 0, the weighted error of the IFWMF objective.  factors = MFB_CURRENT | MFB_BEST. */
int mfb_eval(mfb_engine *e, int which, int factors, int variant, int weighted, int want_norms,
             double out[4]);

/* Ranking positions — the candidate scan of Model::hitRate / arHR and their U / I variants (model.cpp:981-1332).  For
 * every user: the number of candidate items (not invalid, inside the training matrix, not rated by the user in the
 * TRAIN matrix, not the test item itself) whose estRating exceeds that of the user's test item, which is the first
 * rating of its row in matrix `which` (MFB_VAL or MFB_TEST, model.cpp:992).  That count is the test item's 0-based
 * position in the reference's sorted top-N list (scores tie with probability zero), so pos[u] < 10 is a hit of
 * hitRate, 1 / (pos[u] + 1) with pos[u] < 1000 the term of arHR, and the U / I variants are host-side filters over
 * the same array.  pos[u] = -1: the user does not count (invalid, or no rating in `which` — the reference reads the
 * row's first entry unconditionally); -2: counted, but the test item is not a candidate (never a hit).
 * test_item (optional, [n_users]) receives the test items.  MFB_MF / MFB_IFWMF with rank <= 64: dense U V^T on
 * tcgen05 (3xTF32 split, option "rank_tensor_cores" = 1, default) with the count fused into the epilogue; otherwise
 * (and for the pair-dependent prediction ranks of MFB_TMF / MFB_TMFDROPOUT) one CTA per user on CUDA cores, rounded
 * exactly as the reference's loops. */
int mfb_rank_positions(mfb_engine *e, int which, int factors, int variant, int32_t *pos, int32_t *test_item);
/* estRating of every rating of matrix `which`, CSR order (NaN where the user or the item is masked): the
 * predictions Model::NDCG / NDCGU / NDCGI (model.cpp:760-978) rank, [nnz of the matrix] floats. */
int mfb_predict(mfb_engine *e, int which, int factors, int variant, float *pred);

/* Filtered evaluation in one pass (quartileRMSEs, main.cpp:700-768, which calls Model::RMSE(mat, filtItems, ...)
 * model.cpp:348-394, ::SE :397-443 and ::RMSEU :446-486 once per part): user_group[n_users] / item_group[n_items]
 * give every id a group 0..7 or 255 (in no group).  out[((side * 8) + g) * 2 + {0,1}] = sum of squared errors and
 * count of the ratings (valid user and item) whose item (side 0) resp. user (side 1) is in group g. */
int mfb_eval_groups(mfb_engine *e, int which, int factors, int variant, const uint8_t *user_group,
                    const uint8_t *item_group, double out[32]);

/* bestModel = *this (model.cpp:1500-1504) and *this = bestModel (:1492) as device copies */
int mfb_snapshot_best(mfb_engine *e);
int mfb_restore_best(mfb_engine *e);

/* ---- timing on the engine's stream (bench / per-epoch log line, modelMF.cpp:75,106-109) -- */
int mfb_event_record(mfb_engine *e, int32_t slot /* 0..15 */);
int mfb_event_elapsed_ms(mfb_engine *e, int32_t slot_a, int32_t slot_b, float *ms);

/* ---- multi-GPU plumbing (one engine per GPU; the exchange itself is peer memory, see below)
 * Device pointer and leading dimension (floats) of a factor matrix, and the engine's stream
 * (a cudaStream_t) so a communicator can order its work after the engine's kernels. */
int mfb_device_factors(mfb_engine *e, int side, void **dev_ptr, int64_t *ld);
void *mfb_stream(mfb_engine *e);
/* Gather / scatter `n` factor rows (ids = host int32 array) of `side` to / from a packed
 * device buffer [n][ld] — moves a stratum's item block between ranks. */
int mfb_pack_rows(mfb_engine *e, int side, const int32_t *ids, int32_t n, void *dev_buf);
int mfb_unpack_rows(mfb_engine *e, int side, const int32_t *ids, int32_t n, const void *dev_buf);
/* Restrict the rows this engine's ALS / CCD++ / eval kernels own to [begin, end) of `side`
 * (row sharding across ranks); default is everything. */
int mfb_set_row_range(mfb_engine *e, int side, int32_t begin, int32_t end);

/* ---- peer-memory exchange between the engines of one node (one process per GPU, <= 8 ranks) ----
 * The reference has one shared-memory address space (OpenMP threads write uFac / iFac in place,
 * modelMF.cpp:275-303, :806-880); across GPUs the same effect is obtained by mapping every rank's
 * factor matrices into every peer (CUDA IPC over NVLink / NVSwitch) and storing produced rows into
 * the peers from inside the kernels.  Ordering uses 64-bit sequence flags in peer memory; nothing on
 * this path synchronises with the host.
 * init: allocates the flag words and returns this rank's handle blob (320 bytes) to be swapped by the
 * caller's process group; connect: takes the `world` blobs in rank order and maps the peers.
 * After connect, mfb_als_half_step and mfb_ccdpp_rank1 store the rows of this rank's row range
 * (mfb_set_row_range) into all peers and end with a flag barrier. */
int mfb_comm_init(mfb_engine *e, int32_t rank, int32_t world, uint8_t *handles_out, int64_t *handles_bytes);
int mfb_comm_connect(mfb_engine *e, const uint8_t *all_handles, int64_t bytes);
/* The same for engines that live in ONE process (one host thread driving all visible GPUs, as the host classes do;
 * several engines on one device are allowed): engines[r] becomes rank r, peer access between their devices is enabled
 * and the peers' buffers are addressed directly.  The caller issues each step for all ranks before the next step (all
 * calls are asynchronous; flag waits run on the device).  mfb_comm_disconnect undoes either kind of connection. */
int mfb_comm_connect_local(mfb_engine **engines, int32_t world);
int mfb_comm_disconnect(mfb_engine *e);
/* device-side barrier over all ranks on the engines' streams (every rank must call it) */
int mfb_comm_barrier(mfb_engine *e);
/* *timed_out = 1 when a device-side wait gave up (20 s): a peer died or the schedule is inconsistent */
int mfb_comm_error(mfb_engine *e, int32_t *timed_out);
/* DSGD exchange step: push the rows of item part `item_part` of the stratified plan
 * into rank dst_rank's V and publish sequence number `seq` there (dst_rank = -1: into every peer, no
 * flag); wait_block makes the stream wait until rank src_rank has published a sequence >= seq here.
 * Pushes from one source must use increasing sequence numbers. */
int mfb_dsgd_push_block(mfb_engine *e, int32_t item_part, int32_t dst_rank, uint64_t seq);
int mfb_comm_wait_block(mfb_engine *e, int32_t src_rank, uint64_t seq);
/* store n rows of `side` (host id list, or the range [first, first + n) when ids is NULL) into every
 * peer, then barrier: assembles sharded results on all ranks */
int mfb_comm_allgather_rows(mfb_engine *e, int side, const int32_t *ids, int32_t first, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* MFB_H */
