// Internal declarations shared by the engine's translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/mfb.h"

namespace mfb {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

// Device allocations of the engine go through a small per-process cache of whole cudaMalloc blocks: a plan or an
// upload that is repeated with the same sizes (every step of the bench's end-to-end arm, every epoch's evaluation
// plan after a re-upload) gets its buffers back without a cudaMalloc / cudaFree pair — each costs a host
// round trip and a device synchronisation, together more than the kernels of a plan on a busy host.  Blocks are never
// split (CUDA IPC handles stay valid), at most 16 GB are kept, mfb_destroy releases everything cached on its device.
cudaError_t dev_alloc_bytes(void **p, size_t bytes);
cudaError_t dev_free(void *p);
void dev_cache_release(int device);
}  // namespace mfb
struct mfb_engine;
namespace mfb {
// every C-ABI entry: makes the engine's device current and names its stream as the one this thread works on; with
// option "copy_overlap" it also orders the engine's stream behind factor copies still running on the copy stream —
// kJoinUpload: behind a pending factor upload (everything that touches U / V), kJoinDownload: behind a pending factor
// download (everything that WRITES U / V; evaluations may run next to it)
constexpr int kJoinUpload = 1, kJoinDownload = 2, kJoinAll = 3;
cudaError_t enter(mfb_engine *e, int join = kJoinAll);
void leave();
template <class T>
inline cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc_bytes(reinterpret_cast<void **>(p), bytes); }

int fail(const char *what, const char *file, int line);
int fail_cuda(cudaError_t err, const char *expr, const char *file, int line);

#define MFB_CUDA(expr)                                                     \
  do {                                                                     \
    cudaError_t _e = (expr);                                               \
    if (_e != cudaSuccess) return mfb::fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)
#define MFB_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)
#define MFB_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) return mfb::fail(msg, __FILE__, __LINE__);      \
  } while (0)
// every kernel launch goes through this so that mfb_launch_count() is exact
#define MFB_LAUNCH(kernel, grid, block, smem, stream, ...)                \
  do {                                                                    \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);           \
    mfb::g_launches.fetch_add(1, std::memory_order_relaxed);              \
    MFB_CUDA(cudaGetLastError());                                         \
  } while (0)

// A list of row segments of a CSR/CSC matrix, sorted by length (longest first) so that the
// block scheduler starts the long ones first.  Rows longer than `chunk` are split; such rows
// get a slot in a per-row accumulator workspace (slot >= 0), single-segment rows slot = -1.
struct SegPlan {
  int32_t n_seg = 0;
  int32_t n_multi = 0;         // rows split over several segments
  int32_t max_len = 0;
  int32_t n_longer[3] = {0, 0, 0};  // segments longer than 64 / 32 / 16 ratings (they sort first)
  int32_t *row = nullptr;      // [n_seg]
  int32_t *start = nullptr;    // [n_seg] offset into the index/value arrays
  int32_t *len = nullptr;      // [n_seg]
  int32_t *slot = nullptr;     // [n_seg]
  int32_t *multi_row = nullptr;  // [n_multi] row id of each slot
  // CCD++ (memory-ordered plans): consecutive segments grouped into chunks that span at most ccd_cap ratings of
  // the index / residual arrays (warp-streamed kernels, option ccd_stream = 2) — chunk c holds segments [chunk_seg[c], chunk_seg[c + 1])
  int32_t *chunk_seg = nullptr;  // [n_chunk + 1]
  int32_t n_chunk = 0;
  bool built = false;
  void release();
};

struct DevCsr {
  int32_t nrows = 0, ncols = 0;
  int64_t nnz = 0;
  int64_t *rowptr = nullptr, *colptr = nullptr;
  int32_t *rowind = nullptr, *colind = nullptr;
  float *rowval = nullptr, *colval = nullptr;
  SegPlan eval_rows;   // chunked, for mfb_eval
  SegPlan als_rows, als_cols;
  SegPlan ccd_rows, ccd_cols;
  // CCD++ passes with the gathered vector staged in shared memory (csrc/ccdpp.cu): mode 0 = plain kernels, 1 = whole
  // vector staged over the plain plan, 2 = plan split at the block boundaries of the gathered index range
  SegPlan ccd_rows_blk, ccd_cols_blk;
  std::vector<int32_t> ccd_rows_blk_off, ccd_cols_blk_off;  // [blocks + 1] first segment of every block
  int32_t *ccd_rows_doff = nullptr, *ccd_cols_doff = nullptr;  // device copies
  int ccd_rows_mode = 0, ccd_cols_mode = 0;
  void release_ccd();
  void release();
};

// 16-byte per-id record of the frequency-aware models (see mfb_set_aux)
struct __align__(16) Aux {
  int32_t freq;
  int32_t train;  // IFWMF: float bits; TMF: rank; TMFDROPOUT: lambda
  int32_t pred;
  int32_t pad;
};

constexpr int kMaxBlocks = 64;  // blocks per sub-epoch launch (P <= 64)
constexpr int kMaxRanks = 8;    // engines of one node that exchange through peer memory
constexpr int kFlagSlots = 64;  // 64-bit flag words per engine: [0,8) barrier arrivals, [8,16) block arrivals

// Peer-memory view of the other ranks' engines (same node, one process per GPU; CUDA IPC mappings
// over NVLink / NVSwitch).  Index = rank; the own entry points at the local buffers.
struct Comm {
  int rank = 0, world = 1;
  bool connected = false;
  bool ipc = false;  // peers mapped through CUDA IPC (other processes) rather than addressed directly (same process)
  float *U[kMaxRanks] = {}, *V[kMaxRanks] = {}, *uk[kMaxRanks] = {}, *vk[kMaxRanks] = {};
  unsigned long long *flags[kMaxRanks] = {};
  unsigned long long *own_flags = nullptr;  // [kFlagSlots] + ticket + error word
  uint64_t barrier_seq = 0;
};

struct SgdPlan {
  int32_t P = 0;
  int64_t nnz = 0;
  int32_t n_seg = 0;
  // ratings reordered block-major / user-major / CSR order; for P == 1 these alias the CSR
  int32_t *item = nullptr;
  float *val = nullptr;
  bool owns_ratings = false;
  int32_t *seg_user = nullptr, *seg_start = nullptr, *seg_len = nullptr;
  int32_t *rat_user = nullptr;   // user of every rating
  void *recs = nullptr;          // int4 {user, item, rating bits, 0} per rating, shuffled inside every block range
  int *work_counter = nullptr;   // device work-queue head of the persistent kernel
  double hot_item_share = 0.0;   // largest item count / nnz: bounds useful concurrency
  double collision_mass = 0.0;   // sum over items of (count / nnz)^2: P(two random ratings share an item)
  std::vector<int32_t> blk_seg_off, blk_seg_cnt;  // [P*P]
  std::vector<int64_t> blk_nnz;                   // [P*P]
  std::vector<int64_t> blk_rat_off;               // [P*P] first rating of every block
  std::vector<int32_t> item_count;                // ratings per item over the uploaded rows (host copy)
  std::vector<double> blk_hot_share;              // [P*P] hottest item's share of the block's ratings
  std::vector<int64_t> band_rat_off;              // P == 1: rating offsets of the user bands of the shuffled kernel
  int32_t *part_items = nullptr;                  // item ids grouped by item part (device)
  std::vector<int32_t> part_item_off;             // [P+1]
  // Hot item rows (sgd_hot_kernel): inside every block range the records of the items that hold at least
  // sgd_hot_min_count of the block's ratings are moved behind the block's cold records, one contiguous list per
  // item; a list is trained by one CTA that keeps the item row in shared memory.
  std::vector<int64_t> blk_cold_nnz;              // [P*P] records of the block's cold range (starts at blk_rat_off)
  std::vector<int32_t> blk_hot_off, blk_hot_cnt;  // [P*P] the block's lists inside hot_lists
  std::vector<double> blk_cold_share;             // [P*P] hottest cold item's share of the block's cold ratings
  void *hot_lists = nullptr;                      // device int4 {item, first record, records, 0} per list
  int32_t n_hot = 0;
  int64_t hot_nnz = 0;                            // records in hot lists
  double *hot_stat = nullptr;                     // device [3]: sum degree x |u|^2, sum degree, last batch used
  int hot_stat_age = 0;                           // launches since the statistic was refreshed
  double last_norm = 0.0;                         // rating-weighted mean |u|^2 at the previous whole-matrix launch
  bool built = false;
  bool runs_built = true;  // user runs present (P == 1 plans build them lazily)
  void release();
};

}  // namespace mfb

struct mfb_engine {
  int device = 0;
  int n_users = 0, n_items = 0, rank = 0;
  int ld = 0;  // leading dimension of the device factor matrices in floats (rank rounded up to 4)
  int sm_count = 148;
  // tuning knobs (mfb_set_option)
  int opt_sgd_workers = 0;            // 0 = automatic
  int opt_sgd_warps_per_sm = 32;      // automatic mode: persistent warps per SM
  double opt_sgd_max_hot_inflight = 8.0;  // bound on concurrent updates of the hottest item row
  double opt_sgd_flat_hot_lr = 0.15;  // shuffled kernel: cap on (hot-row concurrency x learning rate)
  double opt_sgd_flat_inflight_frac = 2e-4;  // shuffled kernel: ratings in flight <= this fraction of the epoch
  double opt_sgd_flat_inflight_steady = 1e-3;  // the same bound once the user rows have stopped growing (whole-matrix plans)
  double opt_sgd_flat_launch_lr = 1.2e-5;  // shuffled kernel: ratings in flight <= value / learnrate x ratings of the launch
  double opt_sgd_flat_band_mb = 0.0;  // shuffled kernel: user rows per band (MB of U), 0 = one band (the reference's order)
  int opt_sgd_flat_user_store = 0;    // shuffled kernel: 1 = user rows by plain stores (Hogwild on U), 0 = reductions
  int opt_sgd_flat_debug = 0;         // timing diagnostics of the shuffled kernel (results are wrong when set)
  int opt_sgd_atomic = 1;             // item rows updated by vector reductions (no lost updates)
  int opt_sgd_block_order = 0;        // stratified trainers: 0 = user-major runs (reference order), 1 = shuffled inside the blocks
  int opt_sgd_rotate = 0;             // user runs start at a pseudo-random offset (de-correlates heavy users)
  uint64_t opt_sgd_shuffle_seed = 0;  // key of the plan-time physical shuffle of the rating records (the host passes trainSeed)
  int opt_sgd_hot = 1;                // shuffled kernel: hot item rows trained by dedicated CTAs (row in shared memory)
  int opt_sgd_hot_min_count = 1024;   // a hot list holds at least this many ratings
  double opt_sgd_hot_inflight = 16.0; // an item is hot inside a block when the shuffled kernel would keep more updates of its row in flight
  int opt_sgd_hot_max_lists = 127;    // hot items per block (<= 127)
  int opt_sgd_hot_stages = 8;         // rounds a hot CTA stages ahead (4 or 8; rank > 128: 4)
  double opt_sgd_hot_stab = 0.5;      // hot CTAs: mini-batch <= value / (learnrate x rating-weighted mean |u|^2)
  int opt_sgd_hot_pace = 1;           // hot CTAs advance through their list in step with the shuffled kernel
  int opt_sgd_hot_batch = 0;          // ratings per round of a hot CTA, 0 = automatic (<= 64 and <= sgd_flat_hot_lr / learnrate)
  int opt_ccd_smem = 0;               // CCD++: 1 = gathered u_k / v_k staged in shared memory where the shape allows (measured slower than the L1/L2 gather: profiles/r2_ccdpp.md), 2 / 3 = row / column side only
  int opt_ccd_stream = 0;             // CCD++ passes: 0 = one warp per row segment, segments sorted by length (default); 1 = the same kernels over the segments in memory order; 2 = warp-streamed chunks of consecutive segments with the next batch always in flight (ccd_update_flat_kernel).  All three measure 3.87 - 3.89 ms per rank-one step on the Netflix shape (profiles/r2_ccdpp.md)
  int opt_ccd_stage = 0;              // CCD++ warp-streamed kernels: 1 = the gathered vector staged in shared memory by persistent CTAs when it fits 200 KB (measured slower: the pass turns issue-bound)
  int opt_ccd_cap = 4096;             // CCD++ warp-streamed kernels: ratings per chunk (one warp per chunk)
  int opt_ccd_fuse = 1;               // CCD++: add-back / column subtract ride on the first / last update passes
  int opt_als_chunk = 16384;          // ratings per CTA before a row is split over several CTAs
  int opt_als_dual = 1;               // short rows: solve the len x len dual system instead of rank x rank
  int opt_als_tensor_cores = 1;       // rank > 32: Gram on tcgen05 (3xTF32), warp-specialised persistent kernel; 0 = fp32 CUDA-core Gram; 2 = the round-1 one-CTA-per-row kernel (rank > 64); rank 33 .. 64 with 1: MN-major operands + batched solve (csrc/als_mn.cu), 3 = the converter kernel there too
  int opt_rank_tensor_cores = 1;      // ranking positions: 1 = dense U V^T on tcgen05 where the model allows (rank <= 64, plain dot), 0 = CUDA cores
  int opt_als_chol_warps = 211;       // batched rank-64 solver: 211 (default) / 208: two matrices per warp (half-warps), 11 / 8 warps per CTA; 11: one matrix per warp, two record buffers; 16 / 20 / 22: one matrix, one buffer
  int opt_als_debug = 0;              // timing experiments of csrc/als_mn.cu (results are wrong when set)
  int opt_als_ws_split = 0;           // warp-specialised kernel: 0 = converter teams / solver groups picked per side from the mean row length, 1 = the many-short-rows split, 2 = the few-long-rows split
  cudaStream_t stream = nullptr;
  cudaStream_t stream_hot = nullptr;  // hot-row CTAs run next to the shuffled kernel (forked from / joined into `stream`)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // option "copy_overlap": factor uploads / downloads go to their own stream and return without synchronising
  int opt_copy_overlap = 0;
  cudaStream_t stream_copy = nullptr;
  cudaEvent_t ev_copy_mark = nullptr, ev_upload = nullptr, ev_download = nullptr;
  bool upload_pending = false, download_pending = false;
  cudaEvent_t events[16] = {};

  float *U = nullptr, *V = nullptr;          // [n][ld] fp32, padding columns are zero
  float *bestU = nullptr, *bestV = nullptr;  // bestModel snapshot
  uint8_t *bad_user = nullptr, *bad_item = nullptr;  // 1 = invalid
  mfb::Aux *aux_u = nullptr, *aux_i = nullptr;
  float *poisson_cdf = nullptr;              // [rank][rank]
  int aux_variant = -1;

  mfb::DevCsr mat[3];
  mfb::SgdPlan sgd;
  int32_t row_begin[2] = {0, 0}, row_end[2] = {0, 0};

  // evaluation scratch
  double *eval_partial = nullptr;  // [eval_partial_cap][2]
  int32_t eval_partial_cap = 0;
  double *eval_out = nullptr;      // [4] device
  double *eval_out_host = nullptr; // [4] pinned

  // ALS scratch: per split row, an rp x rp Gram + rp rhs accumulator
  float *als_ws = nullptr;
  size_t als_ws_bytes = 0;

  // CCD++ state
  float *res_row = nullptr, *res_col = nullptr;  // residual (CSR order / CSC order)
  // u_k is allocated twice over: [0, n_users + 4) the live vector, [uk_old_offset(), ...) its value from before the
  // first update of the rank-one step (fused column add-back).  One allocation so that a peer's copy of both is
  // reachable from the one exchanged pointer: every rank stores its extracted rows into both halves on all ranks.
  float *uk = nullptr, *vk = nullptr;
  double *ccd_acc = nullptr;  // [slots][2]
  size_t ccd_acc_slots = 0;

  mfb::Comm comm;
  size_t uk_old_offset() const { return ((size_t)n_users + 4 + 31) & ~(size_t)31; }

  // generic scratch for plan building (grown on demand)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
};

namespace mfb {

// ---- ALS (csrc/als.cu, csrc/als_mn.cu) ----
struct AlsArgs {
  const float *Fin;  // opposite side's factors [.][ld]
  float *Fout;       // this side's factors
  int ld, rank;
  const int32_t *ind;
  const float *val;
  const int32_t *seg_row, *seg_start, *seg_len, *seg_slot;
  const int32_t *multi_row;
  float *ws;
  float reg;
  // row-sharded runs: the other ranks' copies of this side's factors (peer memory); every solved row
  // is stored into all of them, i.e. the all-gather is fused into the solve epilogue
  float *Fpeer[kMaxRanks - 1];
  int n_peer;
  int seg0, nseg;  // segments [seg0, seg0 + nseg) of the plan belong to this launch (sorted longest first)
};

__device__ __forceinline__ void store_solution(const AlsArgs &a, int row, int tid, const float *bv) {
  if (tid < a.ld) {
    const float x = tid < a.rank ? bv[tid] : 0.f;
    a.Fout[(size_t)row * a.ld + tid] = x;
    for (int p = 0; p < a.n_peer; p++) a.Fpeer[p][(size_t)row * a.ld + tid] = x;
  }
}
// the rank-64 path on MN-major tensor-core operands (csrc/als_mn.cu): Gram records to global memory, then a batched
// warp-per-matrix Cholesky; n_primal = leading segments of the plan that go through the rank x rank normal equations
int als_mn_half_step(mfb_engine *e, const AlsArgs &a, const SegPlan &sp, int n_primal);
int als_debug_chol64(mfb_engine *e, int32_t n, const float *rec_host, float *x_host, int32_t rank, float reg);

int ensure_scratch(mfb_engine *e, size_t bytes);
// Builds a SegPlan over rows [row_lo,row_hi) of ptr (device int64 [nrows+1]), sorted longest first
// (by_length) or left in memory order; rows whose mask
// byte is set (mask may be null) or that are empty produce no segment.
int build_seg_plan(mfb_engine *e, const int64_t *ptr, int32_t nrows, const uint8_t *mask, int32_t row_lo,
                   int32_t row_hi, int32_t chunk, SegPlan *out, bool by_length = true);

int sgd_plan_build(mfb_engine *e, int32_t P, const int32_t *user_part, const int32_t *item_part);
int sgd_subepoch_launch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float lr, float ureg,
                        float ireg, uint64_t seed, uint64_t counter);
int sgd_flat_launch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float lr, float ureg, float ireg,
                    uint64_t seed, uint64_t counter);
int sgd_debug_hot_batch(mfb_engine *e, double out[3]);
int sgd_debug_records(mfb_engine *e, int32_t a, int32_t b, int32_t *recs_out, int64_t *cold_nnz, int32_t *lists_out,
                      int32_t *n_lists);
int eval_launch(mfb_engine *e, int which, int factors, int variant, int weighted, int want_norms, double out[4]);
int eval_groups_launch(mfb_engine *e, int which, int factors, int variant, const uint8_t *user_group,
                       const uint8_t *item_group, double *out);
int rank_positions_launch(mfb_engine *e, int which, int factors, int variant, int32_t *pos_host, int32_t *test_item_host);
int rank_predict_launch(mfb_engine *e, int which, int factors, int variant, float *pred_host);
int als_half_step_launch(mfb_engine *e, int side, float reg);
int als_debug_gram(mfb_engine *e, int side, int32_t row, float *out, int32_t *rp_out);
int ccdpp_begin_impl(mfb_engine *e);
int ccdpp_rank1_impl(mfb_engine *e, int32_t k, int first_iter, int32_t inner, float ureg, float ireg,
                     int32_t item_freq_thresh);
int ccdpp_end_impl(mfb_engine *e);
int ccd_half_step_impl(mfb_engine *e, int side, float reg, const uint8_t *dims_host);
inline size_t uk_alloc_bytes(const mfb_engine *e) { return sizeof(float) * (e->uk_old_offset() + (size_t)e->n_users + 4); }
int comm_barrier_launch(mfb_engine *e);
// non-zero (mfb_last_error set) when a device-side flag wait of this engine has timed out; syncs the stream
int comm_check_error(mfb_engine *e);

}  // namespace mfb
