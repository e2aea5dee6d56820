// SGD on the device: plan building (ratings bucketed once into the P x P stratum grid, shuffled rating records, hot
// item lists) and the three update kernels.
//
// Reference statements replaced: the per-rating step of modelMF.cpp:83-105 (serial),
// :275-303 (stratified), :1747-1763 (Hogwild) and its weighted / truncated / Poisson-truncated
// forms modelInvPopMF.cpp:356-391, modelDropoutSigmoid.cpp:145-194,
// modelPoissonDropout.cpp:176-229.
//
// Kernels (HBM/L2-bound gather-update, not GEMM-shaped):
//   * sgd_flat_kernel — serial / Hogwild trainers and the shuffled-inside-blocks order of the multi-GPU path: a warp
//     takes 32 consecutive records of the physically shuffled record array with one coalesced load, a sub-warp of G
//     lanes owns one rating, both factor rows are read with 128-bit ld.global.cg and updated with
//     red.global.add.v4.f32; the groups are visited in a freshly keyed pseudo-random order per epoch, pulled from a
//     chunk queue;
//   * sgd_hot_kernel — the most rated item rows of a block: one CTA per item keeps the row in shared memory and trains
//     its rating list in paced mini-batch rounds over cp.async-staged user rows;
//   * sgd_run_kernel — the stratified trainers in the reference's own visiting order (user-major, CSR order in the
//     row, modelMF.cpp:279-281): a persistent sub-warp pulls user runs (longest first) from a work queue and keeps u
//     in registers for the whole run; (item, rating) pairs are fetched G at a time and broadcast by shuffle, the next
//     item row is prefetched during the current reduction.
// The IFWMF weight, the TMF rank and the Poisson-drawn rank are resolved per rating in the same pass in all three.
#include "engine.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>

namespace mfb {

// ---- plan ------------------------------------------------------------------------------------
__global__ void sgd_key_kernel(const int64_t *__restrict__ rowptr, int32_t nrows, const int32_t *__restrict__ rowind,
                               int64_t nnz, const int32_t *__restrict__ user_part,
                               const int32_t *__restrict__ item_part, int32_t P, uint64_t *__restrict__ keys,
                               int32_t *__restrict__ idx) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  // row of nnz j: largest r with rowptr[r] <= j
  int lo = 0, hi = nrows;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= j) lo = mid; else hi = mid;
  }
  int u = lo, item = rowind[j];
  int pu = user_part[u], pi = item_part[item];
  uint32_t blk = (pu < 0 || pi < 0) ? (uint32_t)(P * P) : (uint32_t)(pu * P + pi);
  keys[j] = ((uint64_t)blk << 32) | (uint32_t)u;
  idx[j] = (int32_t)j;
}

__global__ void sgd_gather_kernel(const int32_t *__restrict__ perm, int64_t n, const int32_t *__restrict__ ind,
                                  const float *__restrict__ val, int32_t *__restrict__ oind, float *__restrict__ oval) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int p = perm[j];
  oind[j] = ind[p];
  oval[j] = val[p];
}

__global__ void key_user_kernel(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ out) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[j] = (int32_t)(keys[j] & 0xFFFFFFFFu);
}

__global__ void sgd_seg_key_kernel(const uint64_t *__restrict__ run_key, const int32_t *__restrict__ run_len, int32_t n,
                                   uint64_t *__restrict__ key2, int32_t *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t blk = run_key[i] >> 32;
  key2[i] = (blk << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)run_len[i]);
  idx[i] = i;
}

__global__ void sgd_seg_gather_kernel(const int32_t *__restrict__ perm, int32_t n, const uint64_t *__restrict__ run_key,
                                      const int32_t *__restrict__ run_start, const int32_t *__restrict__ run_len,
                                      int32_t *__restrict__ seg_user, int32_t *__restrict__ seg_start,
                                      int32_t *__restrict__ seg_len) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int p = perm[i];
  seg_user[i] = (int32_t)(run_key[p] & 0xFFFFFFFFu);
  seg_start[i] = run_start[p];
  seg_len[i] = run_len[p];
}

// first run index whose block id >= b, for b = 0..nblk (runs are sorted by block)
__global__ void sgd_blk_bounds_kernel(const uint64_t *__restrict__ run_key, int32_t n, int32_t nblk,
                                      int32_t *__restrict__ bounds) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nblk) return;
  int lo = 0, hi = n;  // first i with (key>>32) >= b
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((uint32_t)(run_key[mid] >> 32) >= (uint32_t)b) hi = mid; else lo = mid + 1;
  }
  bounds[b] = lo;
}

// item counts with the bins in shared memory (n_bins x 4 bytes must fit; one global atomic per touched bin and CTA
// instead of one per rating: the most rated items serialise 10^5 global atomics on one address otherwise)
__global__ void __launch_bounds__(1024) item_hist_smem_kernel(const int32_t *__restrict__ ind, int64_t n, int32_t n_bins,
                                                              int32_t *__restrict__ hist) {
  extern __shared__ int32_t bins[];
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = per * blockIdx.x, hi = min(n, lo + per);
  for (int64_t j = lo + threadIdx.x; j < hi; j += blockDim.x) atomicAdd(bins + __ldg(ind + j), 1);
  __syncthreads();
  for (int i = threadIdx.x; i < n_bins; i += blockDim.x)
    if (bins[i]) atomicAdd(hist + i, bins[i]);
}

__global__ void item_hist_kernel(const int32_t *__restrict__ ind, int64_t n, int32_t *__restrict__ hist) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) atomicAdd(hist + ind[j], 1);
}

__global__ void sq_count_kernel(const int32_t *__restrict__ hist, int32_t n, double *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)hist[i] * (double)hist[i];
}

// ---- physically shuffled rating records for the shuffled kernel -------------------------------
// The epoch order of the serial trainers is a random permutation (modelMF.cpp:76-81).  Fetching
// (user, item, rating) through a permuted index costs three random 4-byte DRAM reads per rating —
// measured: 11.3 of the kernel's 17.2 ms (tools/flat_limits.py).  So the records are shuffled ONCE, at
// plan time, inside every block range (whole matrix, user band or P x P stratum block) into an array of
// 16-byte records; an epoch then visits groups of 32 consecutive records (one coalesced 512-byte load per
// warp) in a freshly keyed pseudo-random order of the groups.  Every rating is visited once per epoch, the
// neighbours inside a group are random ratings, and the order of the groups changes every epoch.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x);

__device__ __forceinline__ uint32_t keyed_bijection(uint32_t x, uint32_t n, uint64_t key) {
  // multiply / xor-shift rounds on the next power of two, cycle-walking back into [0, n)
  uint32_t bits = 1;
  while ((1ull << bits) < n) bits++;
  const uint32_t mask = bits >= 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u);
  const uint32_t sh = (bits + 1) / 2;
  const uint64_t h0 = mix64(key), h1 = mix64(key ^ 0x9E3779B97F4A7C15ull), h2 = mix64(key + 0x632BE59BD9B4E019ull);
  const uint32_t m0 = (uint32_t)(h0 >> 7) | 1u, m1 = (uint32_t)(h1 >> 7) | 1u, m2 = (uint32_t)(h2 >> 7) | 1u;
  const uint32_t a0 = (uint32_t)(h0 >> 40), a1 = (uint32_t)(h1 >> 40), a2 = (uint32_t)(h2 >> 40);
  do {
    x = (x * m0 + a0) & mask; x ^= x >> sh;
    x = (x * m1 + a1) & mask; x ^= x >> sh;
    x = (x * m2 + a2) & mask; x ^= x >> sh;
  } while (x >= n);
  return x;
}

// (user, item, rating) of every rating position as one 16-byte record, in memory order: the shuffle below then
// gathers ONE sector per rating instead of three (6.3 -> 2.5 ms at 100 M ratings).  user = rat_user[j], or the row
// of position j by binary search in rowptr when rat_user is null (whole-matrix plans).
__global__ void sgd_pack_records_kernel(const int64_t *__restrict__ rowptr, int32_t nrows, const int32_t *__restrict__ rat_user,
                                        const int32_t *__restrict__ item, const float *__restrict__ val, int64_t n,
                                        int4 *__restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int user;
  if (rat_user) {
    user = rat_user[j];
  } else {
    int lo = 0, hi = nrows;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rowptr[mid] <= j) lo = mid; else hi = mid;
    }
    user = lo;
  }
  out[j] = make_int4(user, item[j], __float_as_int(val[j]), 0);
}

__global__ void sgd_shuffle_records_kernel(const int4 *__restrict__ packed, int64_t n, const int64_t *__restrict__ blk_off,
                                           int nblk, uint64_t seed, int4 *__restrict__ recs,
                                           const int32_t *__restrict__ range_row, int32_t n_items,
                                           const uint8_t *__restrict__ cls, uint32_t *__restrict__ keys) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int lo = 0, hi = nblk;  // block b with blk_off[b] <= q < blk_off[b + 1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (blk_off[mid] <= q) lo = mid; else hi = mid;
  }
  const int64_t off = blk_off[lo];
  const uint32_t nb = (uint32_t)(blk_off[lo + 1] - off);
  const int64_t j = off + keyed_bijection((uint32_t)(q - off), nb, seed * 0x100000001B3ull + (uint64_t)lo);
  const int4 rec = __ldg(packed + j);
  recs[q] = rec;
  // sort key of the hot / cold split: range in the high bits, 0 = cold or 1 + the item's hot slot in the low 7
  if (keys) keys[q] = ((uint32_t)lo << 7) | (uint32_t)cls[(size_t)range_row[lo] * n_items + rec.y];
}

// ratings per (user part of the range, item) over the block ranges of (pre-shuffle) rating positions
__global__ void hot_hist_kernel(const int32_t *__restrict__ item, int64_t n, const int64_t *__restrict__ blk_off, int nblk,
                                const int32_t *__restrict__ range_row, int32_t n_items, int32_t *__restrict__ hist) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  int lo = 0, hi = nblk;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (blk_off[mid] <= q) lo = mid; else hi = mid;
  }
  atomicAdd(hist + (size_t)range_row[lo] * n_items + item[q], 1);
}

// out[i] = first position whose key is >= queries[i]
__global__ void key_lower_bound_kernel(const uint32_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ queries,
                                       int nq, int64_t *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const uint32_t want = queries[i];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] >= want) hi = mid; else lo = mid + 1;
  }
  out[i] = lo;
}

// Builds pl.recs from (rat_user, item, val).  ranges = starts of the shuffle ranges (ascending) + end;
// range_bid[r] = stratum block (a * P + b) of range r, or empty when the ranges are user bands (no hot split).
//
// Hot / cold split.  Reductions on one factor row serialise in the L2 atomic unit and the number of updates of a
// row that may be in flight is bounded by the learning rate (sgd_flat_launch), so the items that hold many of a
// block's ratings limit how fast the block can be trained (profiles/r1_dsgd_scaling.md).  Their records are sorted
// out of the block's shuffled range into one list per item; sgd_hot_kernel trains a list with the item row in
// shared memory.  item_part may be null for P == 1.
static int sgd_build_records(mfb_engine *e, const std::vector<int64_t> &ranges, const std::vector<int32_t> &range_bid,
                             const int32_t *item_part) {
  SgdPlan &pl = e->sgd;
  cudaStream_t st = e->stream;
  const int64_t n = ranges.back();
  const int nrng = (int)ranges.size() - 1;
  const int P = pl.P;
  const size_t nblk = (size_t)P * P;
  pl.blk_cold_nnz.assign(nblk, 0);
  pl.blk_hot_off.assign(nblk, 0);
  pl.blk_hot_cnt.assign(nblk, 0);
  pl.blk_cold_share.assign(nblk, 0.0);
  pl.n_hot = 0;
  pl.hot_nnz = 0;
  for (size_t r = 0; r < range_bid.size(); r++) pl.blk_cold_nnz[range_bid[r]] = ranges[r + 1] - ranges[r];
  if (range_bid.empty() && nblk == 1) pl.blk_cold_nnz[0] = n;  // user bands: the launch cuts the range itself
  MFB_CUDA(dev_alloc(&pl.recs, sizeof(int4) * (size_t)(n > 0 ? n : 1)));
  if (n == 0) return 0;
  int64_t *d_off;
  MFB_CUDA(dev_alloc(&d_off, sizeof(int64_t) * ranges.size()));
  MFB_CUDA(cudaMemcpyAsync(d_off, ranges.data(), sizeof(int64_t) * ranges.size(), cudaMemcpyHostToDevice, st));
  const unsigned grid = (unsigned)((n + 255) / 256);

  // ---- which items are hot in which block ----
  const bool try_hot = e->opt_sgd_hot && !range_bid.empty() && (P == 1 || item_part) && !pl.item_count.empty();
  struct Cand { int32_t item, count; };
  std::vector<std::vector<Cand>> cand(nblk);
  std::vector<int32_t> cold_max(nblk, 0);
  std::vector<int32_t> range_row((size_t)nrng, 0);
  std::vector<int32_t> hist;  // [P][n_items]
  int32_t *d_row = nullptr, *d_hist = nullptr;
  size_t n_cand = 0;
  if (try_hot) {
    std::vector<char> have(nblk, 0);
    for (int r = 0; r < nrng; r++) { range_row[r] = range_bid[r] / P; have[range_bid[r]] = 1; }
    MFB_CUDA(dev_alloc(&d_row, sizeof(int32_t) * (size_t)nrng));
    MFB_CUDA(cudaMemcpyAsync(d_row, range_row.data(), sizeof(int32_t) * (size_t)nrng, cudaMemcpyHostToDevice, st));
    const int32_t *counts = pl.item_count.data();
    if (P > 1) {
      const size_t hn = (size_t)P * e->n_items;
      MFB_CUDA(dev_alloc(&d_hist, sizeof(int32_t) * hn));
      MFB_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * hn, st));
      MFB_LAUNCH(hot_hist_kernel, grid, 256, 0, st, pl.item, n, d_off, nrng, d_row, e->n_items, d_hist);
      hist.resize(hn);
      MFB_CUDA(cudaMemcpyAsync(hist.data(), d_hist, sizeof(int32_t) * hn, cudaMemcpyDeviceToHost, st));
      MFB_CUDA(cudaStreamSynchronize(st));
      dev_free(d_hist);
      counts = hist.data();
    }
    // An item is hot inside a block when the shuffled kernel, running the block with the whole machine
    // (2 ratings in flight per sub-warp), would keep more than sgd_hot_inflight (default 16) updates of its row in
    // flight — rows near the stability bound of sgd_flat_launch (c x lr <= 0.15) are fine one at a time, a block
    // whose few hundred most rated rows all sit there is not (measured: sporadic divergence of 8 x 8 strata at
    // c = 70, lr = 0.002) — and it holds at least sgd_hot_min_count ratings (a list shorter than that is not worth
    // a CTA).
    const double machine_inflight = 2.0 * (double)e->sm_count * e->opt_sgd_warps_per_sm * 2.0;
    std::vector<int32_t> min_count(nblk, std::max(e->opt_sgd_hot_min_count, 1));
    for (int r = 0; r < nrng; r++) {
      const double share = e->opt_sgd_hot_inflight / machine_inflight;
      const double c = share * (double)(ranges[r + 1] - ranges[r]);
      min_count[range_bid[r]] = (int32_t)std::min<double>(std::max<double>(min_count[range_bid[r]], c), 2e9);
    }
    for (int a = 0; a < P; a++)
      for (int i = 0; i < e->n_items; i++) {
        const int32_t c = counts[(size_t)a * e->n_items + i];
        if (c <= 0) continue;
        const int b = P == 1 ? 0 : item_part[i];
        if (b < 0 || b >= P) continue;
        const size_t bid = (size_t)a * P + b;
        if (!have[bid]) continue;
        if (c >= min_count[bid]) cand[bid].push_back({i, c});
        else cold_max[bid] = std::max(cold_max[bid], c);
      }
    // every list keeps a mini-batch and its staged user rows in flight: a block gets at most one list per 128
    // ratings of the in-flight budget of sgd_flat_launch (small matrices get none: their epochs are too short
    // for any extra staleness, and they have no hot-row problem to solve)
    const size_t budget_lists = e->opt_sgd_hot_inflight > 0 ? (size_t)(e->opt_sgd_flat_inflight_frac * (double)n / 128.0) : 127;
    const size_t max_lists = std::min<size_t>((size_t)std::min(std::max(e->opt_sgd_hot_max_lists, 1), 127), budget_lists);
    for (size_t bid = 0; bid < nblk; bid++) {
      auto &cv = cand[bid];
      std::sort(cv.begin(), cv.end(), [](const Cand &x, const Cand &y) { return x.count != y.count ? x.count > y.count : x.item < y.item; });
      if (cv.size() > max_lists) {
        cold_max[bid] = std::max(cold_max[bid], cv[max_lists].count);
        cv.resize(max_lists);
      }
      if (max_lists == 0) cv.clear();
      n_cand += cv.size();
    }
  }
  // key of the physical shuffle: derived from the training seed (option sgd_shuffle_seed, set by the host trainers from
  // Model::trainSeed before the plan is built — the reference seeds its shuffles with mt19937(trainSeed),
  // modelMF.cpp:63,78), so that group membership and the order of the hot lists differ from seed to seed
  const uint64_t shuffle_key = 0x5EEDULL ^ mix64(e->opt_sgd_shuffle_seed);
  int4 *packed;
  MFB_CUDA(dev_alloc(&packed, sizeof(int4) * (size_t)n));
  MFB_LAUNCH(sgd_pack_records_kernel, grid, 256, 0, st, e->mat[MFB_TRAIN].rowptr, e->n_users, pl.rat_user, pl.item, pl.val, n, packed);
  uint8_t *d_cls = nullptr;
  uint32_t *d_keys = nullptr;
  if (n_cand > 0) {
    std::vector<uint8_t> cls((size_t)P * e->n_items, 0);
    for (size_t bid = 0; bid < nblk; bid++)
      for (size_t k = 0; k < cand[bid].size(); k++) cls[(bid / P) * (size_t)e->n_items + cand[bid][k].item] = (uint8_t)(k + 1);
    MFB_CUDA(dev_alloc(&d_cls, cls.size()));
    MFB_CUDA(cudaMemcpyAsync(d_cls, cls.data(), cls.size(), cudaMemcpyHostToDevice, st));
    MFB_CUDA(dev_alloc(&d_keys, sizeof(uint32_t) * (size_t)n));
    MFB_LAUNCH(sgd_shuffle_records_kernel, grid, 256, 0, st, packed, n, d_off, nrng, shuffle_key,
               reinterpret_cast<int4 *>(pl.recs), d_row, e->n_items, d_cls, d_keys);
    MFB_CUDA(cudaStreamSynchronize(st));  // cls (host vector) is copied
  } else {
    MFB_LAUNCH(sgd_shuffle_records_kernel, grid, 256, 0, st, packed, n, d_off, nrng, shuffle_key,
               reinterpret_cast<int4 *>(pl.recs), nullptr, e->n_items, nullptr, nullptr);
    MFB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_off); dev_free(d_row); dev_free(packed);
    for (size_t bid = 0; bid < nblk; bid++)
      if (pl.blk_cold_nnz[bid] > 0 && try_hot) pl.blk_cold_share[bid] = (double)cold_max[bid] / (double)pl.blk_cold_nnz[bid];
    return 0;
  }
  // ---- stable sort by (range, hot slot): cold records stay in shuffled order at the front of the range ----
  {
    uint32_t *d_keys2;
    int4 *recs2 = packed;  // the packed records have been consumed by the shuffle (same stream): reuse the buffer
    MFB_CUDA(dev_alloc(&d_keys2, sizeof(uint32_t) * (size_t)n));
    int end_bit = 7;
    while ((1 << (end_bit - 7)) < nrng) end_bit++;
    size_t tmp_bytes = 0;
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, reinterpret_cast<int4 *>(pl.recs), recs2,
                                             (int)n, 0, end_bit, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(e->scratch, tmp_bytes, d_keys, d_keys2, reinterpret_cast<int4 *>(pl.recs), recs2,
                                             (int)n, 0, end_bit, st));
    // list bounds
    std::vector<uint32_t> queries;
    for (int r = 0; r < nrng; r++)
      for (size_t k = 0; k < cand[range_bid[r]].size(); k++) queries.push_back(((uint32_t)r << 7) | (uint32_t)(k + 1));
    uint32_t *d_q;
    int64_t *d_pos;
    const int nq = (int)queries.size();
    MFB_CUDA(dev_alloc(&d_q, sizeof(uint32_t) * (size_t)nq));
    MFB_CUDA(dev_alloc(&d_pos, sizeof(int64_t) * (size_t)nq));
    MFB_CUDA(cudaMemcpyAsync(d_q, queries.data(), sizeof(uint32_t) * (size_t)nq, cudaMemcpyHostToDevice, st));
    MFB_LAUNCH(key_lower_bound_kernel, (nq + 127) / 128, 128, 0, st, d_keys2, n, d_q, nq, d_pos);
    std::vector<int64_t> pos((size_t)nq);
    MFB_CUDA(cudaMemcpyAsync(pos.data(), d_pos, sizeof(int64_t) * (size_t)nq, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_q); dev_free(d_pos); dev_free(d_keys); dev_free(d_keys2); dev_free(d_cls); dev_free(d_off); dev_free(d_row);
    dev_free(pl.recs);
    pl.recs = recs2;
    std::vector<int4> lists;
    size_t qi = 0;
    for (int r = 0; r < nrng; r++) {
      const size_t bid = (size_t)range_bid[r];
      const size_t h = cand[bid].size();
      pl.blk_hot_off[bid] = (int32_t)lists.size();
      pl.blk_hot_cnt[bid] = (int32_t)h;
      for (size_t k = 0; k < h; k++) {
        const int64_t lo = pos[qi + k], hi = (k + 1 < h) ? pos[qi + k + 1] : ranges[r + 1];
        lists.push_back(make_int4(cand[bid][k].item, (int32_t)lo, (int32_t)(hi - lo), 0));
        pl.hot_nnz += hi - lo;
      }
      if (h > 0) pl.blk_cold_nnz[bid] = pos[qi] - ranges[r];
      qi += h;
      if (pl.blk_cold_nnz[bid] > 0) pl.blk_cold_share[bid] = (double)cold_max[bid] / (double)pl.blk_cold_nnz[bid];
    }
    pl.n_hot = (int32_t)lists.size();
    MFB_CUDA(dev_alloc(&pl.hot_lists, sizeof(int4) * lists.size()));
    MFB_CUDA(cudaMemcpyAsync(pl.hot_lists, lists.data(), sizeof(int4) * lists.size(), cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

static int sgd_plan_common(mfb_engine *e) {
  SgdPlan &pl = e->sgd;
  const DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  MFB_CUDA(dev_alloc(&pl.work_counter, sizeof(int)));
  MFB_CUDA(cudaMemsetAsync(pl.work_counter, 0, sizeof(int), st));
  MFB_CUDA(dev_alloc(&pl.hot_stat, sizeof(double) * 3));
  MFB_CUDA(cudaMemsetAsync(pl.hot_stat, 0, sizeof(double) * 3, st));
  pl.hot_stat_age = 0;
  pl.last_norm = 0.0;
  pl.hot_item_share = 0.0;
  if (m.nnz > 0) {
    int32_t *hist, *d_max;
    MFB_CUDA(dev_alloc(&hist, sizeof(int32_t) * ((size_t)e->n_items + 1)));
    d_max = hist + e->n_items;
    MFB_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * ((size_t)e->n_items + 1), st));
    const size_t bin_bytes = sizeof(int32_t) * (size_t)e->n_items;
    if (bin_bytes <= 200 * 1024) {
      MFB_CUDA(cudaFuncSetAttribute(item_hist_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_bytes));
      MFB_LAUNCH(item_hist_smem_kernel, e->sm_count, 1024, bin_bytes, st, m.rowind, m.nnz, e->n_items, hist);
    } else {
      MFB_LAUNCH(item_hist_kernel, (unsigned)((m.nnz + 255) / 256), 256, 0, st, m.rowind, m.nnz, hist);
    }
    size_t tmp_bytes = 0;
    MFB_CUDA(cub::DeviceReduce::Max(nullptr, tmp_bytes, hist, d_max, e->n_items, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceReduce::Max(e->scratch, tmp_bytes, hist, d_max, e->n_items, st));
    int32_t mx = 0;
    MFB_CUDA(cudaMemcpyAsync(&mx, d_max, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    double *sq, *d_sum, sum = 0;
    MFB_CUDA(dev_alloc(&sq, sizeof(double) * ((size_t)e->n_items + 1)));
    d_sum = sq + e->n_items;
    MFB_LAUNCH(sq_count_kernel, (e->n_items + 255) / 256, 256, 0, st, hist, e->n_items, sq);
    MFB_CUDA(cub::DeviceReduce::Sum(nullptr, tmp_bytes, sq, d_sum, e->n_items, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceReduce::Sum(e->scratch, tmp_bytes, sq, d_sum, e->n_items, st));
    MFB_CUDA(cudaMemcpyAsync(&sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost, st));
    pl.item_count.resize((size_t)e->n_items);
    MFB_CUDA(cudaMemcpyAsync(pl.item_count.data(), hist, sizeof(int32_t) * (size_t)e->n_items, cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    dev_free(hist);
    dev_free(sq);
    pl.hot_item_share = (double)mx / (double)m.nnz;
    pl.collision_mass = sum / ((double)m.nnz * (double)m.nnz);
  }
  return 0;
}

int sgd_plan_build(mfb_engine *e, int32_t P, const int32_t *user_part, const int32_t *item_part) {
  SgdPlan &pl = e->sgd;
  pl.release();
  const DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  MFB_TRY(sgd_plan_common(e));
  pl.P = P;
  pl.blk_seg_off.assign((size_t)P * P, 0);
  pl.blk_seg_cnt.assign((size_t)P * P, 0);
  pl.blk_nnz.assign((size_t)P * P, 0);
  pl.blk_rat_off.assign((size_t)P * P, 0);
  if (P == 1 && user_part == nullptr) {
    // whole matrix = one block; the user runs of the user-major kernel (the CSR rows themselves) are built
    // on first use (sgd_ensure_runs): the shuffled kernel does not need them
    pl.item = m.rowind;
    pl.val = m.rowval;
    pl.owns_ratings = false;
    pl.nnz = m.nnz;
    pl.blk_nnz[0] = m.nnz;
    pl.runs_built = false;
    // (the user of every rating is resolved by sgd_pack_records_kernel straight from rowptr)
    // user bands of the shuffled kernel: equal user counts, rating offsets read back from rowptr
    int nbands_built = 1;
    {
      const double row_bytes = sizeof(float) * (double)e->ld;
      int nbands = 1;
      if (e->opt_sgd_flat_band_mb > 0)
        nbands = (int)std::ceil((double)e->n_users * row_bytes / (e->opt_sgd_flat_band_mb * 1048576.0));
      nbands = std::max(1, std::min(nbands, std::min(64, e->n_users)));
      nbands_built = nbands;
      pl.band_rat_off.assign((size_t)nbands + 1, 0);
      for (int b = 0; b <= nbands; b++) {
        const int64_t u = (int64_t)e->n_users * b / nbands;
        MFB_CUDA(cudaMemcpyAsync(&pl.band_rat_off[b], m.rowptr + u, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
      }
      MFB_CUDA(cudaStreamSynchronize(st));
    }
    {
      std::vector<int32_t> bids;
      if (nbands_built == 1) bids.push_back(0);  // user bands (an option) keep the plain shuffled ranges
      MFB_TRY(sgd_build_records(e, pl.band_rat_off, bids, nullptr));
    }
    pl.built = true;
    return 0;
  }
  int64_t nnz = m.nnz;
  {  // item ids grouped by item part: the rows an exchange step moves between ranks
    pl.part_item_off.assign((size_t)P + 1, 0);
    for (int i = 0; i < e->n_items; i++)
      if (item_part[i] >= 0 && item_part[i] < P) pl.part_item_off[item_part[i] + 1]++;
    for (int b = 0; b < P; b++) pl.part_item_off[b + 1] += pl.part_item_off[b];
    std::vector<int32_t> ids((size_t)std::max(pl.part_item_off[P], 1)), fill(pl.part_item_off.begin(), pl.part_item_off.end() - 1);
    for (int i = 0; i < e->n_items; i++)
      if (item_part[i] >= 0 && item_part[i] < P) ids[fill[item_part[i]]++] = i;
    MFB_CUDA(dev_alloc(&pl.part_items, sizeof(int32_t) * ids.size()));
    MFB_CUDA(cudaMemcpyAsync(pl.part_items, ids.data(), sizeof(int32_t) * ids.size(), cudaMemcpyHostToDevice, st));
    MFB_CUDA(cudaStreamSynchronize(st));
  }
  int32_t *d_up, *d_ip;
  MFB_CUDA(dev_alloc(&d_up, sizeof(int32_t) * e->n_users));
  MFB_CUDA(dev_alloc(&d_ip, sizeof(int32_t) * e->n_items));
  MFB_CUDA(cudaMemcpyAsync(d_up, user_part, sizeof(int32_t) * e->n_users, cudaMemcpyHostToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(d_ip, item_part, sizeof(int32_t) * e->n_items, cudaMemcpyHostToDevice, st));
  uint64_t *keys, *keys2;
  int32_t *idx, *idx2;
  size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  MFB_CUDA(dev_alloc(&keys, sizeof(uint64_t) * nn));
  MFB_CUDA(dev_alloc(&keys2, sizeof(uint64_t) * nn));
  MFB_CUDA(dev_alloc(&idx, sizeof(int32_t) * nn));
  MFB_CUDA(dev_alloc(&idx2, sizeof(int32_t) * nn));
  int tb = 256;
  unsigned gb = (unsigned)((nnz + tb - 1) / tb);
  if (nnz > 0) MFB_LAUNCH(sgd_key_kernel, gb, tb, 0, st, m.rowptr, e->n_users, m.rowind, nnz, d_up, d_ip, P, keys, idx);
  int end_bit = 32;
  while ((1u << (end_bit - 32)) <= (unsigned)(P * P)) end_bit++;
  size_t tmp_bytes = 0;
  MFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int)nnz, 0, end_bit, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceRadixSort::SortPairs(e->scratch, tmp_bytes, keys, keys2, idx, idx2, (int)nnz, 0, end_bit, st));
  // reordered ratings
  MFB_CUDA(dev_alloc(&pl.item, sizeof(int32_t) * nn));
  MFB_CUDA(dev_alloc(&pl.val, sizeof(float) * nn));
  pl.owns_ratings = true;
  if (nnz > 0) MFB_LAUNCH(sgd_gather_kernel, gb, tb, 0, st, idx2, nnz, m.rowind, m.rowval, pl.item, pl.val);
  // user of every reordered rating (low word of the sorted key) for the shuffled-in-block kernel
  MFB_CUDA(dev_alloc(&pl.rat_user, sizeof(int32_t) * nn));
  if (nnz > 0) MFB_LAUNCH(key_user_kernel, gb, tb, 0, st, keys2, nnz, pl.rat_user);
  // runs of equal (block, user): reuse `keys` for the unique keys, `idx` for the run lengths
  int32_t *d_nruns;
  MFB_CUDA(dev_alloc(&d_nruns, sizeof(int32_t)));
  MFB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, keys2, keys, idx, d_nruns, (int)nnz, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceRunLengthEncode::Encode(e->scratch, tmp_bytes, keys2, keys, idx, d_nruns, (int)nnz, st));
  int32_t nruns = 0;
  MFB_CUDA(cudaMemcpyAsync(&nruns, d_nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  dev_free(d_nruns);
  if (nnz == 0) nruns = 0;
  // run starts = exclusive scan of the run lengths (into idx2)
  int32_t *run_len = idx, *run_start = idx2;
  uint64_t *run_key = keys;
  if (nruns > 0) {
    MFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, run_len, run_start, nruns, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceScan::ExclusiveSum(e->scratch, tmp_bytes, run_len, run_start, nruns, st));
  }
  // block bounds over the runs (the dropped bucket P*P, if any, sorts last)
  int nblk = P * P;
  std::vector<int32_t> bounds(nblk + 1, 0), nnz_at(nblk + 1, 0);
  if (nruns > 0) {
    int32_t *d_bounds;
    MFB_CUDA(dev_alloc(&d_bounds, sizeof(int32_t) * (nblk + 1)));
    MFB_LAUNCH(sgd_blk_bounds_kernel, (nblk + 1 + 255) / 256, 256, 0, st, run_key, nruns, nblk, d_bounds);
    MFB_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds, sizeof(int32_t) * (nblk + 1), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    dev_free(d_bounds);
    // rating offset of each block's first run
    for (int b = 0; b <= nblk; b++) {
      if (bounds[b] >= nruns) {
        nnz_at[b] = -1;  // resolved below
      } else {
        MFB_CUDA(cudaMemcpyAsync(&nnz_at[b], run_start + bounds[b], sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      }
    }
    MFB_CUDA(cudaStreamSynchronize(st));
  }
  int32_t nseg = bounds[nblk];
  // total ratings kept = start of the dropped bucket, or nnz when nothing was dropped
  int64_t kept = nnz;
  if (nruns > 0 && nnz_at[nblk] >= 0) kept = nnz_at[nblk];
  for (int b = nblk; b >= 0; b--)
    if (nnz_at[b] < 0) nnz_at[b] = (int32_t)kept;
  pl.n_seg = nseg;
  pl.nnz = kept;
  for (int b = 0; b < nblk; b++) {
    pl.blk_seg_off[b] = bounds[b];
    pl.blk_seg_cnt[b] = bounds[b + 1] - bounds[b];
    pl.blk_nnz[b] = (int64_t)nnz_at[b + 1] - nnz_at[b];
    pl.blk_rat_off[b] = nnz_at[b];
  }
  {  // hottest item row of every block: its share of the block's ratings bounds the useful concurrency.
     // Item counts are over the locally uploaded rows; they are spread over the user parts present here.
    std::vector<int32_t> part_max((size_t)P, 0);
    for (int i = 0; i < e->n_items; i++)
      if (item_part[i] >= 0 && item_part[i] < P && !pl.item_count.empty())
        part_max[item_part[i]] = std::max(part_max[item_part[i]], pl.item_count[i]);
    int local_parts = 0;
    for (int a = 0; a < P; a++) {
      int64_t t = 0;
      for (int b = 0; b < P; b++) t += pl.blk_nnz[(size_t)a * P + b];
      if (t > 0) local_parts++;
    }
    pl.blk_hot_share.assign((size_t)P * P, 0.0);
    for (int a = 0; a < P; a++)
      for (int b = 0; b < P; b++) {
        const int64_t bn = pl.blk_nnz[(size_t)a * P + b];
        if (bn > 0)
          pl.blk_hot_share[(size_t)a * P + b] = std::min(1.0, (double)part_max[b] / std::max(local_parts, 1) / (double)bn);
      }
  }
  size_t ns = (size_t)(nseg > 0 ? nseg : 1);
  MFB_CUDA(dev_alloc(&pl.seg_user, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&pl.seg_start, sizeof(int32_t) * ns));
  MFB_CUDA(dev_alloc(&pl.seg_len, sizeof(int32_t) * ns));
  if (nseg > 0) {
    // longest-first inside every block
    uint64_t *key2 = keys2;  // sorted rating keys no longer needed
    uint64_t *key2_out = keys2 + nseg;  // nnz >= 2 * nseg is not guaranteed: allocate separately if short
    int32_t *sidx, *sidx_out;
    bool own_key_out = (size_t)nnz < 2 * (size_t)nseg;
    if (own_key_out) MFB_CUDA(dev_alloc(&key2_out, sizeof(uint64_t) * ns));
    MFB_CUDA(dev_alloc(&sidx, sizeof(int32_t) * 2 * ns));
    sidx_out = sidx + nseg;
    int gs = (nseg + tb - 1) / tb;
    MFB_LAUNCH(sgd_seg_key_kernel, gs, tb, 0, st, run_key, run_len, nseg, key2, sidx);
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key2, key2_out, sidx, sidx_out, nseg, 0, end_bit, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(e->scratch, tmp_bytes, key2, key2_out, sidx, sidx_out, nseg, 0, end_bit, st));
    MFB_LAUNCH(sgd_seg_gather_kernel, gs, tb, 0, st, sidx_out, nseg, run_key, run_start, run_len, pl.seg_user,
               pl.seg_start, pl.seg_len);
    MFB_CUDA(cudaStreamSynchronize(st));
    if (own_key_out) dev_free(key2_out);
    dev_free(sidx);
  }
  MFB_CUDA(cudaStreamSynchronize(st));
  dev_free(keys); dev_free(keys2); dev_free(idx); dev_free(idx2); dev_free(d_up); dev_free(d_ip);
  {  // shuffle blocks = the P x P stratum blocks (block-major order, the dropped bucket excluded)
    std::vector<int64_t> ranges;
    std::vector<int32_t> bids;
    for (int b = 0; b < nblk; b++)
      if (pl.blk_nnz[b] > 0) { ranges.push_back(pl.blk_rat_off[b]); bids.push_back(b); }
    ranges.push_back(kept);
    if (ranges.size() == 1) { ranges.insert(ranges.begin(), 0); bids.clear(); }
    MFB_TRY(sgd_build_records(e, ranges, bids, item_part));
  }
  pl.built = true;
  return 0;
}

// ---- update kernels ----------------------------------------------------------------------------
struct SgdArgs {
  float *U, *V;
  int nq;  // float4 words per factor row (ld / 4)
  int rank;
  const int32_t *item;
  const float *val;
  const int32_t *seg_user, *seg_start, *seg_len;
  int nb, max_cnt, total;  // total = nb * max_cnt segment slots
  int rotate;              // start every run at a pseudo-random offset
  int user_store;          // shuffled kernel: user rows written back with plain stores instead of reductions
  int debug;               // shuffled kernel, timing diagnostics: 1 skip U write, 2 skip V write, 4 skip u load, 8 skip v load,
                           // 16 skip the hot-row CTAs, 32 skip the shuffled kernel when hot lists exist, 64 swap their launch order
  int32_t off[kMaxBlocks], cnt[kMaxBlocks];
  int *counter;  // dynamic work queue head
  float lr, ureg, ireg;
  const Aux *aux_u, *aux_i;
  const float *cdf;
  uint64_t seed, counter_id;
  // shuffled kernel: the scheduled ranges of the shuffled record array are cut into groups of 32
  // records; the groups are visited in the order p(t) = a keyed pseudo-random permutation of [0, n)
  // (3-round multiply / xor-shift bijection on the next power of two with cycle walking).
  int64_t n;  // number of 32-record groups of this launch
  uint32_t perm_bits, perm_mul[3], perm_add[3];
  const int4 *recs;                  // shuffled records (see sgd_shuffle_records_kernel)
  int32_t rat_off[kMaxBlocks];       // first record of every scheduled range
  int32_t rat_len[kMaxBlocks];       // records in the range
  int32_t grp_cum[kMaxBlocks + 1];   // cumulative group counts of the ranges
  // hot-row kernel: the lists of scheduled block i are hot_lists[hot_base[i] .. ), CTA c serves list c - hot_cum[i]
  const int4 *hot_lists;
  int32_t hot_base[kMaxBlocks], hot_cum[kMaxBlocks + 1];
  // pacing: the shuffled kernel publishes the number of 32-record groups it has finished (every 32 groups of a
  // warp), a hot CTA does not run ahead of that share of its own list (see sgd_hot_kernel)
  // adaptive mini-batch of the hot-row CTAs: stat[0] = sum over users of degree x |u|^2, stat[1] = sum of degrees
  // (sgd_hot_stat_kernel), stat[2] = the batch the last CTA used (diagnostics); batch <= hot_stab / (lr x mean |u|^2)
  double *hot_stat;
  float hot_stab;
  unsigned int *progress;
  int pace, chunk;
  unsigned pace_lead;  // groups the queue head runs ahead of the finished work (two chunks per warp)
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Inverse-CDF Poisson(lambda) draw clamped to [1, rank] (modelPoissonDropout.cpp:200-207)
__device__ __forceinline__ int poisson_rank(const float *__restrict__ cdf, int rank, int lambda, uint64_t seed,
                                            uint64_t counter, uint32_t rating_idx) {
  uint64_t h = mix64(seed ^ mix64(counter * 0x100000001B3ull + rating_idx));
  float uf = (float)(h >> 40) * (1.0f / 16777216.0f);  // [0,1)
  const float *row = cdf + (size_t)(lambda - 1) * rank;
  int lo = 0, hi = rank;  // smallest k with row[k] >= uf, or rank
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(row + mid) >= uf) hi = mid; else lo = mid + 1;
  }
  int k = lo;
  if (k > rank) k = rank;
  if (k < 1) k = 1;
  return k;
}

// 128-bit vector reduction into global memory (sm_90+): no lost update when two workers hit the
// same factor row, at the cost of a slightly stale read (mini-batch semantics on hot rows).
__device__ __forceinline__ void red_add_v4(float4 *addr, float4 d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w)
               : "memory");
}

// One SGD step on the G-lane slice of (u, v) held by this lane.  Returns the new u in place and
// either the new v (ATOMIC = false) or the increment of v (ATOMIC = true) in vo.
//   u -= lr * (-2 g v + 2 ureg u);  v -= lr * (-2 g u_new + 2 ireg v)      (modelMF.cpp:95-103)
template <int G, int VPL, bool TRUNC>
__device__ __forceinline__ float dot_slice(const float4 (&u)[VPL], const float4 (&v)[VPL], int sl, int k) {
  float p = 0.f;
#pragma unroll
  for (int c = 0; c < VPL; c++) {
    if (TRUNC) {
      const int base = (c * G + sl) * 4;
      p += (base + 0 < k ? u[c].x * v[c].x : 0.f) + (base + 1 < k ? u[c].y * v[c].y : 0.f) +
           (base + 2 < k ? u[c].z * v[c].z : 0.f) + (base + 3 < k ? u[c].w * v[c].w : 0.f);
    } else {
      p = fmaf(u[c].x, v[c].x, p);
      p = fmaf(u[c].y, v[c].y, p);
      p = fmaf(u[c].z, v[c].z, p);
      p = fmaf(u[c].w, v[c].w, p);
    }
  }
#pragma unroll
  for (int m = G / 2; m >= 1; m >>= 1) p += __shfl_xor_sync(0xFFFFFFFFu, p, m);
  return p;
}

// Stratified trainers: a persistent sub-warp pulls user runs (longest first) from a work queue.
template <int G, int VPL, int VARIANT, bool ATOMIC>
__global__ void __launch_bounds__(128) sgd_run_kernel(const SgdArgs a) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  constexpr bool TRUNC = (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT);
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1);
  const float4 *Vq = reinterpret_cast<const float4 *>(a.V);
  float4 *Vw = reinterpret_cast<float4 *>(a.V);
  float4 *Uw = reinterpret_cast<float4 *>(a.U);
  const float lr = a.lr, two_ureg = 2.0f * a.ureg, two_ireg = 2.0f * a.ireg;

  bool done = false, have = false, fresh = false;
  // a run is visited from a per-run pseudo-random offset, wrapping around: heavy users all
  // start first (longest-first queue) and would otherwise sweep the item ids in lockstep and
  // collide on the same item rows
  int user = 0, pos = 0, seg_end = 0, cbase = 0, base = 0, off = 0;
  int c_it = 0, c_pay = 0, n_it = 0, n_pay = 0;
  float c_rt = 0.f, n_rt = 0.f;
  int ufreq = 0, upay = 0;
  float4 u[VPL], vn[VPL];
  bool own[VPL];
#pragma unroll
  for (int c = 0; c < VPL; c++) {
    own[c] = (c * G + sl) < a.nq;
    u[c] = vn[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  auto fetch = [&](int q, int &it, float &rt, int &pay) {  // q-th rating of the rotated run
    it = 0; rt = 0.f; pay = 0;
    if (q < seg_end) {
      int j = q + off;
      if (j >= seg_end) j -= seg_end;
      j += base;
      it = __ldg(a.item + j);
      rt = __ldg(a.val + j);
      if (VARIANT != MFB_MF) {
        Aux ai = a.aux_i[it];
        // the rarer side decides (modelInvPopMF.cpp:164-166, modelDropoutSigmoid.cpp:158)
        pay = (ufreq < ai.freq) ? upay : ai.train;
        if (VARIANT == MFB_TMFDROPOUT) pay = poisson_rank(a.cdf, a.rank, pay, a.seed, a.counter_id, (uint32_t)j);
      }
    }
  };

  for (;;) {
    const bool need = !done && pos >= seg_end;
    if (need && have) {
#pragma unroll
      for (int c = 0; c < VPL; c++)
        if (own[c]) __stcg(Uw + (size_t)user * a.nq + c * G + sl, u[c]);
    }
    int s = -1;
    if (need && sl == 0) {
      for (;;) {
        const int idx = atomicAdd(a.counter, 1);
        if (idx >= a.total) break;
        const int b = idx % a.nb, local = idx / a.nb;
        if (local < a.cnt[b]) { s = a.off[b] + local; break; }
      }
    }
    s = __shfl_sync(kFull, s, 0, G);
    if (need) {
      if (s < 0) {
        done = true;
        have = false;
      } else {
        user = a.seg_user[s];
        base = a.seg_start[s];
        seg_end = a.seg_len[s];
        pos = 0;
        off = a.rotate ? (int)(mix64(a.seed ^ (a.counter_id * 0x9E3779B97F4A7C15ull + (uint32_t)user)) % (uint32_t)seg_end) : 0;
        have = true;
        fresh = true;
#pragma unroll
        for (int c = 0; c < VPL; c++)
          if (own[c]) u[c] = __ldcg(Uw + (size_t)user * a.nq + c * G + sl);
        if (VARIANT != MFB_MF) {
          Aux au = a.aux_u[user];
          ufreq = au.freq;
          upay = au.train;
        }
        cbase = pos;
        fetch(pos + sl, c_it, c_rt, c_pay);
        fetch(pos + G + sl, n_it, n_rt, n_pay);
      }
    }
    if (__all_sync(kFull, done)) break;
    if (!done && pos >= cbase + G) {
      c_it = n_it; c_rt = n_rt; c_pay = n_pay;
      cbase += G;
      fetch(cbase + G + sl, n_it, n_rt, n_pay);
    }
    const int t = (pos - cbase) & (G - 1);
    const int it = __shfl_sync(kFull, c_it, t, G);
    const float rt = __shfl_sync(kFull, c_rt, t, G);
    int pay = 0;
    if (VARIANT != MFB_MF) pay = __shfl_sync(kFull, c_pay, t, G);
    const int itn_same = __shfl_sync(kFull, c_it, (t + 1) & (G - 1), G);
    const int itn_next = __shfl_sync(kFull, n_it, 0, G);
    const int itn = (t + 1 < G) ? itn_same : itn_next;
    float4 v[VPL];
#pragma unroll
    for (int c = 0; c < VPL; c++) v[c] = vn[c];
    if (!done && fresh) {
#pragma unroll
      for (int c = 0; c < VPL; c++)
        if (own[c]) v[c] = __ldcg(Vq + (size_t)it * a.nq + c * G + sl);
    }
    if (!done && pos + 1 < seg_end) {
#pragma unroll
      for (int c = 0; c < VPL; c++)
        if (own[c]) vn[c] = __ldcg(Vq + (size_t)itn * a.nq + c * G + sl);
    }
    fresh = false;
    const int k = TRUNC ? pay : a.rank;  // leading dimensions this update touches
    const float p = dot_slice<G, VPL, TRUNC>(u, v, sl, k);
    float g = rt - p;  // diff
    if (VARIANT == MFB_IFWMF) g *= __int_as_float(pay);
    const float m2g = -2.0f * g;
    if (!done) {
#pragma unroll
      for (int c = 0; c < VPL; c++) {
        if (!own[c]) continue;
        const int base = (c * G + sl) * 4;
        if (TRUNC && base >= k) continue;  // nothing of this word is touched
        const float4 uu = u[c], vv = v[c];
        float4 un, dv;
        un.x = fmaf(-lr, fmaf(two_ureg, uu.x, m2g * vv.x), uu.x);
        un.y = fmaf(-lr, fmaf(two_ureg, uu.y, m2g * vv.y), uu.y);
        un.z = fmaf(-lr, fmaf(two_ureg, uu.z, m2g * vv.z), uu.z);
        un.w = fmaf(-lr, fmaf(two_ureg, uu.w, m2g * vv.w), uu.w);
        dv.x = -lr * fmaf(two_ireg, vv.x, m2g * un.x);
        dv.y = -lr * fmaf(two_ireg, vv.y, m2g * un.y);
        dv.z = -lr * fmaf(two_ireg, vv.z, m2g * un.z);
        dv.w = -lr * fmaf(two_ireg, vv.w, m2g * un.w);
        if (TRUNC) {
          if (base + 1 >= k) { un.y = uu.y; dv.y = 0.f; }
          if (base + 2 >= k) { un.z = uu.z; dv.z = 0.f; }
          if (base + 3 >= k) { un.w = uu.w; dv.w = 0.f; }
        }
        u[c] = un;
        float4 *dst = Vw + (size_t)it * a.nq + c * G + sl;
        if (ATOMIC) red_add_v4(dst, dv);
        else __stcg(dst, make_float4(vv.x + dv.x, vv.y + dv.y, vv.z + dv.z, vv.w + dv.w));
      }
      pos++;
    }
  }
}

// Keyed bijection of [0, n): each round (odd multiply + add, xor-shift right) is a bijection of the
// perm_bits-bit integers; values >= n are walked through again (at most ~2 rounds on average).
__device__ __forceinline__ uint32_t permute_index(const SgdArgs &a, uint32_t x) {
  const uint32_t mask = a.perm_bits >= 32 ? 0xFFFFFFFFu : ((1u << a.perm_bits) - 1u);
  const uint32_t sh = (a.perm_bits + 1) / 2;
  do {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      x = (x * a.perm_mul[r] + a.perm_add[r]) & mask;
      x ^= x >> sh;
    }
  } while (x >= (uint32_t)a.n);
  return x;
}

// Serial / Hogwild trainers (modelMF.cpp:83-105, :1747-1763): a warp takes a group of 32 shuffled
// records with one coalesced 512-byte load (the next group is fetched while the current one is
// processed) and works through it 32 / G ratings at a time — every sub-warp of G lanes owns one rating:
// both factor rows are read with 128-bit loads and updated with vector reductions.
template <int G, int VPL, int VARIANT>
__global__ void __launch_bounds__(128) sgd_flat_kernel(const SgdArgs a) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  constexpr bool TRUNC = (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT);
  constexpr int PER = 32 / G;  // ratings a warp has in flight
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1), sub = lane / G;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 *Uw = reinterpret_cast<float4 *>(a.U), *Vw = reinterpret_cast<float4 *>(a.V);
  const float lr = a.lr, two_ureg = 2.0f * a.ureg, two_ireg = 2.0f * a.ireg;
  bool own[VPL];
#pragma unroll
  for (int c = 0; c < VPL; c++) own[c] = (c * G + sl) < a.nq;
  // lane `lane` holds record `lane` of the group; q0 = its position in the record array (RNG counter)
  auto fetch_group = [&](int64_t t, int4 &rec, int &cnt, uint32_t &q0) {
    rec = make_int4(0, 0, 0, 0); cnt = 0; q0 = 0;
    if (t < a.n) {
      uint32_t g = permute_index(a, (uint32_t)t);
      int b = 0;
      while (b + 1 < a.nb && g >= (uint32_t)a.grp_cum[b + 1]) b++;
      g -= (uint32_t)a.grp_cum[b];
      cnt = min(32, a.rat_len[b] - (int)g * 32);
      q0 = (uint32_t)a.rat_off[b] + g * 32u;
      if (lane < cnt) rec = __ldcs(a.recs + q0 + lane);
    }
  };
  int4 n_rec;
  int n_cnt;
  uint32_t n_q0;
  // Work distribution.  Long launches (a.chunk > 0): warps pull chunks of a.chunk consecutive positions of the
  // keyed group order from a queue (the head doubles as the progress the hot-row CTAs pace themselves against), so
  // that SMs which also host a hot CTA simply take fewer chunks; the next chunk is requested one chunk ahead.
  // Short launches (a.chunk == 0, a few groups per warp): static round-robin, no atomics.
  const unsigned CH = (unsigned)a.chunk;
  int64_t t = wid;
  unsigned left = 0, raw_next = 0;
  if (CH > 0) {
    unsigned c0 = 0;
    if (lane == 0) { c0 = atomicAdd(a.progress, 1u); raw_next = atomicAdd(a.progress, 1u); }
    t = (int64_t)__shfl_sync(kFull, c0, 0) * CH;
    left = CH - 1;
  }
  fetch_group(t, n_rec, n_cnt, n_q0);
  while (t < a.n) {  // warp-uniform
    int64_t t_next;
    if (CH == 0) {
      t_next = t + n_warps;
    } else if (left > 0) {
      t_next = t + 1;
      left--;
    } else {
      t_next = (int64_t)__shfl_sync(kFull, raw_next, 0) * CH;
      if (lane == 0) raw_next = atomicAdd(a.progress, 1u);
      left = CH - 1;
    }
    const int4 rec = n_rec;
    const int cnt = n_cnt;
    const uint32_t q0 = n_q0;
    fetch_group(t_next, n_rec, n_cnt, n_q0);
    t = t_next;
#pragma unroll 2
    for (int i = 0; i < G; i++) {
      if (i * PER >= cnt) break;  // warp-uniform
      const int r = i * PER + sub;
      const bool on = r < cnt;
      const int user = __shfl_sync(kFull, rec.x, r), it = __shfl_sync(kFull, rec.y, r);
      const float rt = __int_as_float(__shfl_sync(kFull, rec.z, r));
      int pay = 0;
      float4 u[VPL], v[VPL];
#pragma unroll
      for (int c = 0; c < VPL; c++) u[c] = v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on) {
#pragma unroll
        for (int c = 0; c < VPL; c++)
          if (own[c]) {
            if (!(a.debug & 4)) u[c] = __ldcg(Uw + (size_t)user * a.nq + c * G + sl);
            if (!(a.debug & 8)) v[c] = __ldcg(Vw + (size_t)it * a.nq + c * G + sl);
          }
        if (VARIANT != MFB_MF) {
          const Aux au = a.aux_u[user], ai = a.aux_i[it];
          pay = (au.freq < ai.freq) ? au.train : ai.train;
          if (VARIANT == MFB_TMFDROPOUT) pay = poisson_rank(a.cdf, a.rank, pay, a.seed, a.counter_id, q0 + (uint32_t)r);
        }
      }
      const int k = TRUNC ? pay : a.rank;
      const float pdot = dot_slice<G, VPL, TRUNC>(u, v, sl, k);
      float g = rt - pdot;
      if (VARIANT == MFB_IFWMF) g *= __int_as_float(pay);
      const float m2g = -2.0f * g;
      if (on) {
#pragma unroll
        for (int c = 0; c < VPL; c++) {
          if (!own[c]) continue;
          const int base = (c * G + sl) * 4;
          if (TRUNC && base >= k) continue;
          const float4 uu = u[c], vv = v[c];
          float4 du, dv;
          du.x = -lr * fmaf(two_ureg, uu.x, m2g * vv.x);
          du.y = -lr * fmaf(two_ureg, uu.y, m2g * vv.y);
          du.z = -lr * fmaf(two_ureg, uu.z, m2g * vv.z);
          du.w = -lr * fmaf(two_ureg, uu.w, m2g * vv.w);
          dv.x = -lr * fmaf(two_ireg, vv.x, m2g * (uu.x + du.x));
          dv.y = -lr * fmaf(two_ireg, vv.y, m2g * (uu.y + du.y));
          dv.z = -lr * fmaf(two_ireg, vv.z, m2g * (uu.z + du.z));
          dv.w = -lr * fmaf(two_ireg, vv.w, m2g * (uu.w + du.w));
          if (TRUNC) {
            if (base + 1 >= k) { du.y = 0.f; dv.y = 0.f; }
            if (base + 2 >= k) { du.z = 0.f; dv.z = 0.f; }
            if (base + 3 >= k) { du.w = 0.f; dv.w = 0.f; }
          }
          if (a.debug & 1) {  // diagnostics only (option sgd_flat_debug): keep the value alive without writing
            if (du.x == 12345.678f) Uw[0] = du;
          } else if (a.user_store) {
            __stcg(Uw + (size_t)user * a.nq + c * G + sl, make_float4(uu.x + du.x, uu.y + du.y, uu.z + du.z, uu.w + du.w));
          } else {
            red_add_v4(Uw + (size_t)user * a.nq + c * G + sl, du);
          }
          if (a.debug & 2) {
            if (dv.x == 12345.678f) Vw[0] = dv;
          } else {
            red_add_v4(Vw + (size_t)it * a.nq + c * G + sl, dv);
          }
        }
      }
    }
  }
}

// Hot item rows (see sgd_build_records): one CTA trains one list = all the ratings of one item inside one block,
// as a sequence of small dense mini-batch steps.  The item row v lives in shared memory for the whole list.  A round
// takes T ratings (T <= 64, bounded like the concurrency of the shuffled kernel: sgd_flat_hot_lr / learnrate):
//   * the T user rows are staged in a shared-memory tile by cp.async, S rounds ahead of their use (records 2 S
//     rounds ahead: the loaded-memory latency next to the shuffled kernel is ~2 us, a round lasts ~0.3 us);
//   * Q adjacent threads own one rating, each E 128-bit words of the row: partial dot products, a log2(Q)-step
//     shuffle, then the user row is updated by vector reductions exactly as in sgd_flat_kernel (the users of one
//     item's ratings are distinct, the rows are shared with the other kernels) and the rating's increment of v
//     is written over its tile row;
//   * the column sums of the tile are added to v.
// Against a sub-warp-per-rating layout this needs a quarter of the instructions per rating (two shuffles instead of
// nine, no per-rating cross-warp reduction), which is what bounds a CTA that must finish a 230 k-rating list alone.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kHotMaxParts = 16;  // row groups of the column-sum pass

__host__ __device__ inline int hot_row_quads(int nq, int Q) {  // smallest row stride >= nq, congruent to Q mod 8 (bank-conflict free)
  return nq + (((Q - nq) % 8) + 8) % 8;
}

// Rating-weighted mean squared norm of the user rows (the raters of an item are drawn in proportion to their
// degree).  A mini-batch of T ratings of one item moves v by 2 lr sum_j u_j (r_j - u_j . v): along a direction the
// raters share this contracts by 1 - 2 lr T |u|^2, stable only while lr T |u|^2 < 1 — measured at 1/20 of the
// bench matrix: lr = 0.005 is fine with T = 8 and diverges with T = 16.  The hot CTAs therefore bound their batch by
// hot_stab / (lr x mean |u|^2) (sgd_hot_stab, default 0.5) with the mean taken on the device, no host round trip.
__global__ void sgd_hot_stat_kernel(const float *__restrict__ U, int ld, const int64_t *__restrict__ rowptr, int n_users,
                                    double *__restrict__ stat) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  double s0 = 0.0, s1 = 0.0;
  for (int u = warp * 4; u < n_users; u += n_warps * 4) {  // every fourth user: a sample is enough
    const int64_t d = rowptr[u + 1] - rowptr[u];
    if (d <= 0) continue;
    float p = 0.f;
    for (int c = lane; c < ld; c += 32) { const float x = __ldcg(U + (size_t)u * ld + c); p = fmaf(x, x, p); }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(0xFFFFFFFFu, p, m);
    s0 += (double)d * (double)p;
    s1 += (double)d;
  }
  if (lane == 0 && s1 > 0.0) { atomicAdd(stat, s0); atomicAdd(stat + 1, s1); }
}

template <int Q, int E, int VARIANT, bool FULL, int S>  // FULL: nq == Q * E, every thread owns E words; S ring stages
__global__ void __launch_bounds__(512) sgd_hot_kernel(const SgdArgs a, const int T) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  constexpr bool TRUNC = (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT);
  extern __shared__ float4 hot_sm[];
  const int tid = threadIdx.x, jl = tid / Q, q = tid % Q;
  int b = 0;
  while (b + 1 < a.nb && (int)blockIdx.x >= a.hot_cum[b + 1]) b++;
  const int4 hl = a.hot_lists[a.hot_base[b] + ((int)blockIdx.x - a.hot_cum[b])];
  const int it = hl.x, off = hl.y, len = hl.z;
  const int nq = FULL ? Q * E : a.nq, rowq = hot_row_quads(nq, Q);
  const int np = min(kHotMaxParts, max((int)blockDim.x / nq, 1));
  float4 *vs = hot_sm;                                                 // [nq] the item row
  float4 *part = vs + nq;                                              // [np][nq] partial column sums
  int4 *rec_ring = reinterpret_cast<int4 *>(part + np * nq);           // [2 S][T]
  int4 *aux_ring = rec_ring + 2 * S * T;                               // [S][T]
  float4 *tile = reinterpret_cast<float4 *>(aux_ring + S * T);         // [S][T][rowq]
  float4 *Uw = reinterpret_cast<float4 *>(a.U), *Vw = reinterpret_cast<float4 *>(a.V);
  const float c1 = -a.lr * 2.0f * a.ureg, c3 = -a.lr * 2.0f * a.ireg, lr2 = a.lr * 2.0f;
  for (int t = tid; t < nq; t += blockDim.x) vs[t] = __ldcg(Vw + (size_t)it * nq + t);
  bool own[E];
#pragma unroll
  for (int c = 0; c < E; c++) own[c] = FULL || (q + Q * c) < nq;
  // effective batch (see sgd_hot_stat_kernel): Te <= T rows of the tile are used per round
  int Te = T;
  {
    const double s0 = a.hot_stat[0], s1 = a.hot_stat[1];
    if (s1 > 0.0 && s0 > 0.0) {
      const double bound = (double)a.hot_stab * s1 / ((double)a.lr * s0);
      if (bound < (double)T) Te = max(1, (int)bound);
    }
    if (blockIdx.x == 0 && tid == 0) a.hot_stat[2] = (double)Te;
  }
  const bool in_tile = jl < T;       // owns a tile row (writes zeros when it has no rating)
  const bool lane_on = jl < Te;
  Aux ai = {0, 0, 0, 0};
  if (VARIANT != MFB_MF) ai = a.aux_i[it];
  // loop-invariant addresses: this thread's words of its tile row / of v, its column-sum slice
  const float4 *vq = vs + q;
  const int sum_f = tid % nq, sum_pg = tid / nq;
  const bool fast = FULL && (Q * T) % 32 == 0 && np * E == T;  // every thread sums E rows of one column
  bool pacing = a.pace != 0;
  const uint32_t row_off = (uint32_t)jl * rowq + q;   // + stage * T * rowq
  const uint32_t stage_quads = (uint32_t)T * rowq;
  const int4 *rec_src = a.recs + off + jl;            // + x * Te
  __syncthreads();
  const int rounds = (len + Te - 1) / Te;
  // the list is walked from a fresh starting round every epoch (round x of the epoch is chunk (x + r0) mod rounds of
  // the list): the reference reshuffles its whole rating order per epoch (modelMF.cpp:76-81), a list replayed in the
  // same order would give the most rated rows the same update sequence every time
  const int r0 = (int)(mix64(a.seed ^ mix64(a.counter_id * 0x9E3779B97F4A7C15ull + (uint64_t)(uint32_t)it)) % (uint64_t)rounds);
  auto chunk_of = [&](int x) { const int c = x + r0; return c >= rounds ? c - rounds : c; };
  for (int r = -(2 * S - 1); r < rounds; r++) {
    {  // record of round r + 2S - 1
      const int x = r + 2 * S - 1;
      if (x < rounds) {
        const int cx = chunk_of(x), j = cx * Te + jl;
        if (lane_on && q == 0 && j < len) cp_async16(rec_ring + (x & (2 * S - 1)) * T + jl, rec_src + cx * Te);
      }
    }
    {  // user row (and aux record) of round r + S - 1: its record was committed S iterations ago
      const int x = r + S - 1, j = (x >= 0 && x < rounds) ? chunk_of(x) * Te + jl : len;
      if (x >= 0 && lane_on && j < len) {
        const int user = rec_ring[(x & (2 * S - 1)) * T + jl].x;
        float4 *dst = tile + (x & (S - 1)) * stage_quads + row_off;
        const float4 *src = Uw + (size_t)user * nq + q;
#pragma unroll
        for (int c = 0; c < E; c++)
          if (own[c]) cp_async16(dst + Q * c, src + Q * c);
        if (VARIANT != MFB_MF && q == 0) cp_async16(aux_ring + (x & (S - 1)) * T + jl, a.aux_u + user);
      }
    }
    cp_async_commit();
    cp_async_wait<S - 1>();
    __syncwarp();
    if (r < 0) continue;
    const int st = r & (S - 1);
    float4 *row = tile + st * stage_quads + row_off;
    const int j = chunk_of(r) * Te + jl;
    const bool on = lane_on && j < len;
    int4 rec = make_int4(0, 0, 0, 0);
    int pay = 0;
    float4 u[E];
    if (on) {
      rec = rec_ring[(r & (2 * S - 1)) * T + jl];
#pragma unroll
      for (int c = 0; c < E; c++)
        if (own[c]) u[c] = row[Q * c];
      if (VARIANT != MFB_MF) {
        const Aux au = *reinterpret_cast<const Aux *>(aux_ring + st * T + jl);
        pay = (au.freq < ai.freq) ? au.train : ai.train;
        if (VARIANT == MFB_TMFDROPOUT) pay = poisson_rank(a.cdf, a.rank, pay, a.seed, a.counter_id, (uint32_t)(off + j));
      }
    }
#pragma unroll
    for (int c = 0; c < E; c++)
      if (!on || !own[c]) u[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int k = TRUNC ? pay : a.rank;
    if (pacing && tid == 0 && (r & 7) == 0 && r > 0) {
      // Pacing.  A list alone would be finished in a fraction of the epoch: its item would take all its updates
      // while the user rows have hardly moved, which is not the uniformly shuffled epoch of modelMF.cpp:76-81
      // (visible in epoch 0, when the factors grow from their 0.01-scale start).  So the list advances with the
      // shuffled kernel: round r waits until that kernel has finished r / rounds of its groups (5 % slack).  If the
      // queue head stands still for 2 ms (a chunk of a concurrency-capped launch takes ~0.3 ms) nothing else is
      // running — a profiler serialises the kernels — and the pacing is switched off.
      const uint64_t target = ((uint64_t)a.n * (uint64_t)r) / (uint64_t)rounds;
      const uint64_t slack = (uint64_t)(a.n / 20) + 64u;
      unsigned last = 0xFFFFFFFFu;
      int idle = 0;
      for (;;) {
        const unsigned c = *reinterpret_cast<volatile unsigned int *>(a.progress);
        uint64_t done = (uint64_t)c * (uint64_t)a.chunk;
        done = done > a.pace_lead ? done - a.pace_lead : 0;
        if (done + slack >= target) break;
        if (c == last) {
          if (++idle >= 8192) { pacing = false; break; }
        } else {
          idle = 0;
          last = c;
        }
        __nanosleep(256);
      }
    }
    __syncthreads();  // v of the previous round is complete (its final sum ran next to the copies above)
    float4 v[E];
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
    for (int c = 0; c < E; c++) {
      v[c] = own[c] ? vq[Q * c] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (TRUNC) {
        const int base = (q + Q * c) * 4;
        p0 += base + 0 < k ? u[c].x * v[c].x : 0.f;
        p1 += base + 1 < k ? u[c].y * v[c].y : 0.f;
        p2 += base + 2 < k ? u[c].z * v[c].z : 0.f;
        p3 += base + 3 < k ? u[c].w * v[c].w : 0.f;
      } else {
        p0 = fmaf(u[c].x, v[c].x, p0);
        p1 = fmaf(u[c].y, v[c].y, p1);
        p2 = fmaf(u[c].z, v[c].z, p2);
        p3 = fmaf(u[c].w, v[c].w, p3);
      }
    }
    float p = (p0 + p1) + (p2 + p3);
#pragma unroll
    for (int m = Q / 2; m >= 1; m >>= 1) p += __shfl_xor_sync(kFull, p, m);
    float g = __int_as_float(rec.z) - p;
    if (VARIANT == MFB_IFWMF) g *= __int_as_float(pay);
    // u += -lr (2 ureg u - 2 g v);  v += -lr (2 ireg v - 2 g u_new)      (modelMF.cpp:95-103)
    const float c2 = on ? lr2 * g : 0.f, c3j = on ? c3 : 0.f;
    float4 *udst = Uw + (size_t)rec.x * nq + q;
#pragma unroll
    for (int c = 0; c < E; c++) {
      if (!own[c]) continue;
      const float4 uu = u[c], vv = v[c];
      float4 du, d;
      du.x = fmaf(c1, uu.x, c2 * vv.x);
      du.y = fmaf(c1, uu.y, c2 * vv.y);
      du.z = fmaf(c1, uu.z, c2 * vv.z);
      du.w = fmaf(c1, uu.w, c2 * vv.w);
      d.x = fmaf(c3j, vv.x, c2 * (uu.x + du.x));
      d.y = fmaf(c3j, vv.y, c2 * (uu.y + du.y));
      d.z = fmaf(c3j, vv.z, c2 * (uu.z + du.z));
      d.w = fmaf(c3j, vv.w, c2 * (uu.w + du.w));
      if (TRUNC) {
        const int base = (q + Q * c) * 4;
        if (base + 0 >= k) { du.x = 0.f; d.x = 0.f; }
        if (base + 1 >= k) { du.y = 0.f; d.y = 0.f; }
        if (base + 2 >= k) { du.z = 0.f; d.z = 0.f; }
        if (base + 3 >= k) { du.w = 0.f; d.w = 0.f; }
        if (on && base < k) red_add_v4(udst + Q * c, du);
      } else {
        if (on) red_add_v4(udst + Q * c, du);
      }
      if (in_tile) row[Q * c] = d;
    }
    __syncthreads();
    if (fast) {
      // column sums, every thread: E consecutive rows of one 128-bit column, then the column's other row groups of
      // the same warp by shuffle; one partial per warp (nq < 32) or per row group
      constexpr int NQ = Q * E;
      const float4 *col = tile + st * stage_quads + (uint32_t)(sum_pg * E) * rowq + sum_f;
      float4 d[E];
#pragma unroll
      for (int i = 0; i < E; i++) d[i] = col[i * rowq];
#pragma unroll
      for (int w = 1; w < E; w <<= 1)
#pragma unroll
        for (int i = 0; i + w < E; i += 2 * w) {
          d[i].x += d[i + w].x; d[i].y += d[i + w].y; d[i].z += d[i + w].z; d[i].w += d[i + w].w;
        }
      float4 acc = d[0];
#pragma unroll
      for (int m = NQ; m < 32; m <<= 1) {
        acc.x += __shfl_xor_sync(kFull, acc.x, m);
        acc.y += __shfl_xor_sync(kFull, acc.y, m);
        acc.z += __shfl_xor_sync(kFull, acc.z, m);
        acc.w += __shfl_xor_sync(kFull, acc.w, m);
      }
      if (NQ >= 32) part[sum_pg * NQ + sum_f] = acc;
      else if ((tid & 31) < NQ) part[(tid >> 5) * NQ + sum_f] = acc;
      __syncthreads();
      if (tid < NQ) {  // runs next to the other warps' copies of the next iteration
        const int nparts = NQ >= 32 ? np : (int)(blockDim.x >> 5);
        float4 acc0 = vs[tid], acc1 = make_float4(0.f, 0.f, 0.f, 0.f);
        int pg = 0;
#pragma unroll 4
        for (; pg + 1 < nparts; pg += 2) {
          const float4 d0 = part[pg * NQ + tid], d1 = part[(pg + 1) * NQ + tid];
          acc0.x += d0.x; acc0.y += d0.y; acc0.z += d0.z; acc0.w += d0.w;
          acc1.x += d1.x; acc1.y += d1.y; acc1.z += d1.z; acc1.w += d1.w;
        }
        if (pg < nparts) {
          const float4 d0 = part[pg * NQ + tid];
          acc0.x += d0.x; acc0.y += d0.y; acc0.z += d0.z; acc0.w += d0.w;
        }
        vs[tid] = make_float4(acc0.x + acc1.x, acc0.y + acc1.y, acc0.z + acc1.z, acc0.w + acc1.w);
      }
      continue;
    }
    // generic ranks / tiny batches: np row groups x nq words, strided over the CTA (it may have fewer threads than
    // the row has words: rank 256 with one rating per round is 8 threads wide)
    for (int idx = tid; idx < np * nq; idx += blockDim.x) {
      const int f = idx % nq, pg = idx / nq;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 *col = tile + st * stage_quads + f;
      for (int jj = pg; jj < T; jj += np) {
        const float4 d = col[jj * rowq];
        acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
      }
      part[pg * nq + f] = acc;
    }
    __syncthreads();
    for (int f = tid; f < nq; f += blockDim.x) {
      float4 acc0 = vs[f];
      for (int pg = 0; pg < np; pg++) {
        const float4 d0 = part[pg * nq + f];
        acc0.x += d0.x; acc0.y += d0.y; acc0.z += d0.z; acc0.w += d0.w;
      }
      vs[f] = acc0;
    }
  }
  __syncthreads();
  for (int t = tid; t < nq; t += blockDim.x) __stcg(Vw + (size_t)it * nq + t, vs[t]);
}

template <int G, int VPL>
static int launch_run(mfb_engine *e, const SgdArgs &a, int variant, int workers, bool atomic) {
  const int tb = 128;
  const unsigned grid = (unsigned)(((int64_t)workers * G + tb - 1) / tb);
#define MFB_RUN_CASE(V)                                                                           \
  if (atomic) MFB_LAUNCH((sgd_run_kernel<G, VPL, V, true>), grid, tb, 0, e->stream, a);          \
  else MFB_LAUNCH((sgd_run_kernel<G, VPL, V, false>), grid, tb, 0, e->stream, a);
  switch (variant) {
    case MFB_MF: MFB_RUN_CASE(MFB_MF) break;
    case MFB_IFWMF: MFB_RUN_CASE(MFB_IFWMF) break;
    case MFB_TMF: MFB_RUN_CASE(MFB_TMF) break;
    default: MFB_RUN_CASE(MFB_TMFDROPOUT) break;
  }
#undef MFB_RUN_CASE
  return 0;
}

template <int G, int VPL>
static int launch_flat(mfb_engine *e, const SgdArgs &a, int variant, int workers) {
  const int tb = 128;
  const unsigned grid = (unsigned)(((int64_t)workers * G + tb - 1) / tb);
  // The kernel reads its factor rows past L1 (ld.global.cg) and streams its records (ld.global.cs): it has no use
  // for L1, and an SM whose L1 / shared-memory split was set by a resident CTA keeps it until the SM drains — the
  // hot-row CTAs (up to 164 KB of shared memory each) could then only start after this kernel has finished.
#define MFB_FLAT_CASE(V)                                                                                          \
  MFB_CUDA(cudaFuncSetAttribute(sgd_flat_kernel<G, VPL, V>, cudaFuncAttributePreferredSharedMemoryCarveout,        \
                                (int)cudaSharedmemCarveoutMaxShared));                                              \
  MFB_LAUNCH((sgd_flat_kernel<G, VPL, V>), grid, tb, 0, e->stream, a);
  switch (variant) {
    case MFB_MF: MFB_FLAT_CASE(MFB_MF) break;
    case MFB_IFWMF: MFB_FLAT_CASE(MFB_IFWMF) break;
    case MFB_TMF: MFB_FLAT_CASE(MFB_TMF) break;
    default: MFB_FLAT_CASE(MFB_TMFDROPOUT) break;
  }
#undef MFB_FLAT_CASE
  return 0;
}

template <int Q, int E, int S>
static int launch_hot(mfb_engine *e, const SgdArgs &a, int variant, int n_lists, int T) {
  const size_t nq = (size_t)a.nq;
  unsigned tb;
  size_t smem;
  for (;; T /= 2) {  // the staged tile must fit the SM's shared memory (rank > 128: 32 ratings per round)
    tb = (unsigned)((Q * T + 31) / 32 * 32);
    const size_t np = std::min<size_t>(kHotMaxParts, std::max<size_t>(tb / nq, 1));
    smem = sizeof(float4) * (nq + np * nq + 2 * S * T + S * T + (size_t)S * T * (size_t)hot_row_quads(a.nq, Q));
    if (smem <= 200 * 1024 || T == 1) break;
  }
  cudaStream_t st = e->stream_hot;
#define MFB_HOT_CASE(V)                                                                                                        \
  if (a.nq == Q * E) {                                                                                                           \
    MFB_CUDA(cudaFuncSetAttribute(sgd_hot_kernel<Q, E, V, true, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    MFB_LAUNCH((sgd_hot_kernel<Q, E, V, true, S>), n_lists, tb, smem, st, a, T);                                                 \
  } else {                                                                                                                       \
    MFB_CUDA(cudaFuncSetAttribute(sgd_hot_kernel<Q, E, V, false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MFB_LAUNCH((sgd_hot_kernel<Q, E, V, false, S>), n_lists, tb, smem, st, a, T);                                                \
  }
  switch (variant) {
    case MFB_MF: MFB_HOT_CASE(MFB_MF) break;
    case MFB_IFWMF: MFB_HOT_CASE(MFB_IFWMF) break;
    case MFB_TMF: MFB_HOT_CASE(MFB_TMF) break;
    default: MFB_HOT_CASE(MFB_TMFDROPOUT) break;
  }
#undef MFB_HOT_CASE
  return 0;
}

// Number of concurrent sub-warp workers.  Concurrency is what makes parallel SGD differ from the
// serial loop: the expected number of in-flight updates that hit the hottest item row is
// (ratings in flight) x (that item's share of the ratings).  It is kept at or below
// `sgd_max_hot_inflight` (default 8), and never above what fills the machine.
static int pick_workers(const mfb_engine *e, int G, int64_t units, double hot_share, int inflight_per_worker) {
  if (e->opt_sgd_workers > 0) return (int)std::min<int64_t>(e->opt_sgd_workers, std::max<int64_t>(units, 1));
  const int per_warp = 32 / G;
  int64_t hw = (int64_t)e->sm_count * e->opt_sgd_warps_per_sm * per_warp;
  double cap = e->opt_sgd_max_hot_inflight / (std::max(hot_share, 1e-9) * inflight_per_worker);
  int64_t w = std::min<int64_t>(hw, (int64_t)std::max(cap, 1.0));
  w = std::min<int64_t>(w, std::max<int64_t>(units, 1));
  w = (w + per_warp - 1) / per_warp * per_warp;
  return (int)std::max<int64_t>(w, per_warp);
}

static void fill_common(mfb_engine *e, SgdArgs &a, float lr, float ureg, float ireg, uint64_t seed, uint64_t counter) {
  const SgdPlan &pl = e->sgd;
  a.U = e->U; a.V = e->V;
  a.nq = e->ld / 4;
  a.rank = e->rank;
  a.item = pl.item; a.val = pl.val;
  a.seg_user = pl.seg_user; a.seg_start = pl.seg_start; a.seg_len = pl.seg_len;
  a.lr = lr; a.ureg = ureg; a.ireg = ireg;
  a.aux_u = e->aux_u; a.aux_i = e->aux_i; a.cdf = e->poisson_cdf;
  a.seed = seed; a.counter_id = counter;
  a.counter = e->sgd.work_counter;
  a.n = pl.nnz;
  a.perm_bits = 1;
  for (int r = 0; r < 3; r++) { a.perm_mul[r] = 1; a.perm_add[r] = 0; }
  a.rotate = e->opt_sgd_rotate;
  a.progress = nullptr;
  a.pace = 0; a.chunk = 0; a.pace_lead = 0;
  a.hot_stat = nullptr; a.hot_stab = 0.f;
  a.user_store = e->opt_sgd_flat_user_store;
  a.debug = e->opt_sgd_flat_debug;
}

static int sgd_ensure_runs(mfb_engine *e) {
  SgdPlan &pl = e->sgd;
  if (pl.runs_built) return 0;
  const DevCsr &m = e->mat[MFB_TRAIN];
  SegPlan sp;
  MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, 0, e->n_users, 0, &sp));
  pl.n_seg = sp.n_seg;
  pl.seg_user = sp.row; pl.seg_start = sp.start; pl.seg_len = sp.len;
  sp.row = sp.start = sp.len = nullptr;
  sp.release();
  pl.blk_seg_cnt[0] = pl.n_seg;
  pl.runs_built = true;
  return 0;
}

int sgd_subepoch_launch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float lr, float ureg,
                        float ireg, uint64_t seed, uint64_t counter) {
  MFB_TRY(sgd_ensure_runs(e));
  const SgdPlan &pl = e->sgd;
  SgdArgs a;
  fill_common(e, a, lr, ureg, ireg, seed, counter);
  a.nb = nb;
  a.max_cnt = 0;
  int64_t nseg = 0;
  for (int i = 0; i < nb; i++) {
    size_t bid = (size_t)blocks[2 * i] * pl.P + blocks[2 * i + 1];
    a.off[i] = pl.blk_seg_off[bid];
    a.cnt[i] = pl.blk_seg_cnt[bid];
    nseg += a.cnt[i];
    if (a.cnt[i] > a.max_cnt) a.max_cnt = a.cnt[i];
  }
  if (a.max_cnt == 0) return 0;
  a.total = a.max_cnt * nb;
  MFB_CUDA(cudaMemsetAsync(pl.work_counter, 0, sizeof(int), e->stream));
  const int nq = a.nq;
  const bool atomic = e->opt_sgd_atomic != 0;
#define MFB_PICK(G, VPL) return launch_run<G, VPL>(e, a, variant, pick_workers(e, G, nseg, pl.hot_item_share, 2), atomic)
  if (nq <= 2) MFB_PICK(2, 1);
  if (nq <= 4) MFB_PICK(4, 1);
  if (nq <= 8) MFB_PICK(8, 1);
  if (nq <= 16) MFB_PICK(16, 1);
  if (nq <= 32) MFB_PICK(32, 1);
  MFB_PICK(32, 2);
#undef MFB_PICK
}

// keyed permutation parameters of one launch: a fresh order of the groups per (seed, epoch, band)
static void fill_perm(SgdArgs &a, int64_t n, uint64_t seed, uint64_t counter, uint64_t salt) {
  a.n = n;
  a.perm_bits = 1;
  while (((int64_t)1 << a.perm_bits) < n) a.perm_bits++;
  uint64_t h = seed * 0x9E3779B97F4A7C15ull + counter * 0xD1B54A32D192ED03ull + salt * 0xA24BAED4963EE407ull + 0x2545F4914F6CDD1Dull;
  for (int r = 0; r < 3; r++) {
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    a.perm_mul[r] = (uint32_t)(h >> 7) | 1u;
    a.perm_add[r] = (uint32_t)(h >> 33);
  }
}

int sgd_flat_launch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float lr, float ureg, float ireg,
                    uint64_t seed, uint64_t counter) {
  SgdPlan &pl = e->sgd;
  SgdArgs a;
  fill_common(e, a, lr, ureg, ireg, seed, counter);
  a.recs = reinterpret_cast<const int4 *>(pl.recs);
  a.nb = nb; a.max_cnt = 0; a.total = 0;
  a.grp_cum[0] = 0;
  a.hot_lists = reinterpret_cast<const int4 *>(pl.hot_lists);
  a.hot_cum[0] = 0;
  int64_t n_all = 0, n_sched = 0;  // cold records / all records of the scheduled blocks
  for (int i = 0; i < nb; i++) {
    size_t bid = (size_t)blocks[2 * i] * pl.P + blocks[2 * i + 1];
    a.rat_off[i] = (int32_t)pl.blk_rat_off[bid];
    a.rat_len[i] = (int32_t)pl.blk_cold_nnz[bid];
    a.grp_cum[i + 1] = a.grp_cum[i] + (int32_t)((pl.blk_cold_nnz[bid] + 31) / 32);
    a.hot_base[i] = pl.blk_hot_off[bid];
    a.hot_cum[i + 1] = a.hot_cum[i] + pl.blk_hot_cnt[bid];
    n_all += pl.blk_cold_nnz[bid];
    n_sched += pl.blk_nnz[bid];
  }
  const int n_lists = a.hot_cum[nb];
  if (n_sched == 0) return 0;
  const int nq = a.nq;
  // The shuffled kernel reads both rows right before it adds its increments, so concurrency only
  // turns the updates of a hot row into a mini-batch of c = (ratings in flight) x (the row's share
  // of the ratings).  That is harmless while c * 2 * lr * |u|^2 stays well below 1 (measured: the
  // validation curve is identical to four digits from c = 1 to c = 90 at lr = 0.002) and diverges
  // beyond it, which only tiny matrices can reach; c is capped at sgd_flat_hot_lr / lr.
  // A second bound keeps the ratings in flight below a fixed fraction of the epoch
  // (sgd_flat_inflight_frac, default 2e-4): during the first epochs the factors grow exponentially
  // from their 0.01-scale start, and on a small matrix a few hundred stale updates are a visible
  // share of that growth (measured: +10 % RMSE after epoch 0 with 0.1 % of a 240 k-rating epoch in
  // flight, < 0.5 % with 0.02 %).  On the bench matrix this bound is 20 k workers, i.e. inactive.
  // The hottest row's share is taken over the cold records of the scheduled blocks (the rows in hot lists are
  // trained by sgd_hot_kernel), the in-flight bound against all the ratings this engine trains.
  double hot_share = 0.0;
  for (int i = 0; i < nb; i++) {
    const size_t bid = (size_t)blocks[2 * i] * pl.P + blocks[2 * i + 1];
    double sh = pl.blk_cold_share[bid];
    if (sh <= 0.0) sh = pl.blk_hot_share.empty() ? pl.hot_item_share : pl.blk_hot_share[bid];  // no hot split in this plan
    if (n_all > 0) hot_share = std::max(hot_share, sh * (double)pl.blk_cold_nnz[bid] / (double)n_all);
  }
  // A third bound is on the share of ONE launch that is in flight: a single stratum block of an 8 x 8 grid is short
  // enough for the whole machine to hold 1.7 % of it in flight, and there the run diverges although no single row
  // is near the c x lr bound (many rows are stale at once; tools/dsgd_sweep.py: fine at 0.9 %, NaN at 1.7 %,
  // lr = 0.002).  In flight <= sgd_flat_launch_lr / lr x (ratings of the launch), default 1.2e-5: 0.6 % at lr = 0.002.
  // Bound (ii) is about the growth phase: once the user rows have stopped growing (rating-weighted mean |u|^2 within
  // 0.8 .. 1.25 of its value at the previous launch) the budget is sgd_flat_inflight_steady (default 1e-3) —
  // tools/inflight_phase.py, ML-20M shape: strict for the first epoch only gives the curve of strict throughout
  // (within the 0.6 % run-to-run spread) at 2.6 instead of 6.5 ms per epoch; relaxed from the start is off by
  // 1.7 - 4.5 % in epochs 1 - 3.  Whole-matrix plans only: the statistic is read back here (one 16-byte copy).
  double inflight_frac = e->opt_sgd_flat_inflight_frac;
  bool stat_fresh = false;
  if (pl.P == 1 && nb == 1) {
    const DevCsr &m = e->mat[MFB_TRAIN];
    MFB_CUDA(cudaMemsetAsync(pl.hot_stat, 0, sizeof(double) * 2, e->stream));
    MFB_LAUNCH(sgd_hot_stat_kernel, e->sm_count * 8, 256, 0, e->stream, e->U, e->ld, m.rowptr, e->n_users, pl.hot_stat);
    stat_fresh = true;
    if (e->opt_sgd_flat_inflight_steady > inflight_frac) {
      double s[2] = {0.0, 0.0};
      MFB_CUDA(cudaMemcpyAsync(s, pl.hot_stat, sizeof(double) * 2, cudaMemcpyDeviceToHost, e->stream));
      MFB_CUDA(cudaStreamSynchronize(e->stream));
      const double norm = s[1] > 0.0 ? s[0] / s[1] : 0.0;
      if (pl.last_norm > 0.0 && norm > 0.8 * pl.last_norm && norm < 1.25 * pl.last_norm) inflight_frac = e->opt_sgd_flat_inflight_steady;
      pl.last_norm = std::isfinite(norm) ? norm : 0.0;
    }
  }
  const double inflight_cap =
      std::max(std::min(inflight_frac * (double)std::max<int64_t>(pl.nnz, n_sched),
                        e->opt_sgd_flat_launch_lr / std::max((double)lr, 1e-12) * (double)n_sched), 8.0);
  double hot_cap = e->opt_sgd_flat_hot_lr / std::max((double)lr, 1e-12);
  const double lr_cap = hot_cap;
  if (hot_share > 0) hot_cap = std::min(hot_cap, inflight_cap * hot_share);
  const double saved_cap = e->opt_sgd_max_hot_inflight;
  e->opt_sgd_max_hot_inflight = hot_cap;
  struct Restore { mfb_engine *e; double v; ~Restore() { e->opt_sgd_max_hot_inflight = v; } } restore{e, saved_cap};
  const int G_flat = nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
  // Sets the work distribution of one launch of the shuffled kernel over b.n groups: a chunk queue for long
  // launches (>= 64 groups per warp; the queue head is zeroed on the stream), static round-robin otherwise.
  auto prep = [&](SgdArgs &b, int64_t n_ratings, int &workers) -> int {
    workers = pick_workers(e, G_flat, n_ratings, hot_share, 2);
    const int64_t n_warps = (((int64_t)workers * G_flat + 127) / 128) * 4;
    b.chunk = (b.n >= n_warps * 64) ? 16 : 0;
    b.progress = reinterpret_cast<unsigned int *>(pl.work_counter);
    b.pace_lead = (unsigned)std::min<int64_t>(2 * n_warps * b.chunk, 0x7FFFFFFF);
    if (b.chunk > 0) MFB_CUDA(cudaMemsetAsync(pl.work_counter, 0, sizeof(int), e->stream));
    return 0;
  };
  auto launch = [&](const SgdArgs &b, int workers) -> int {
#define MFB_PICK(G, VPL) return launch_flat<G, VPL>(e, b, variant, workers)
    if (nq <= 2) MFB_PICK(2, 1);
    if (nq <= 4) MFB_PICK(4, 1);
    if (nq <= 8) MFB_PICK(8, 1);
    if (nq <= 16) MFB_PICK(16, 1);
    if (nq <= 32) MFB_PICK(32, 1);
    MFB_PICK(32, 2);
#undef MFB_PICK
  };
  // User bands (option sgd_flat_band_mb, off by default): the epoch visits the ratings band by band, in a
  // fresh random order inside each band and a rotated band order per epoch, so that a band's user rows
  // stay in the 126 MB L2.  Measured on the bench matrix (gpurun_out/band_sweep.log, profiles/): only
  // 3-6 % faster and the validation curve leaves the reference's (a uniformly shuffled epoch,
  // modelMF.cpp:76-81) by up to 10 % at equal epochs — hence not the default.
  const int nbands = (nb == 1 && pl.P == 1) ? (int)pl.band_rat_off.size() - 1 : 0;
  if (nbands > 1) {
    const int first = (int)(mix64(seed ^ (counter * 0x9E3779B97F4A7C15ull)) % (uint64_t)nbands);
    for (int k = 0; k < nbands; k++) {
      const int b = (first + k) % nbands;
      const int64_t lo = pl.band_rat_off[b], n = pl.band_rat_off[b + 1] - lo;
      if (n <= 0) continue;
      SgdArgs c = a;
      c.rat_off[0] = (int32_t)lo;
      c.rat_len[0] = (int32_t)n;
      c.grp_cum[1] = (int32_t)((n + 31) / 32);
      fill_perm(c, c.grp_cum[1], seed, counter, (uint64_t)b + 1);
      int workers = 0;
      MFB_TRY(prep(c, n, workers));
      MFB_TRY(launch(c, workers));
    }
    return 0;
  }
  fill_perm(a, a.grp_cum[nb], seed, counter, 0);
  if (n_lists > 0) {
    // hot lists first, on their own (high-priority) stream next to the shuffled kernel.  Mini-batch per list:
    // T <= sgd_flat_hot_lr / lr, and all lists together stay inside the in-flight bound of the epoch.
    const double w_max = std::min(lr_cap, inflight_cap / (double)n_lists);
    int T = 1;
    while (T * 2 <= 64 && (double)(T * 2) <= w_max) T *= 2;
    if (e->opt_sgd_hot_batch > 0) T = std::min(e->opt_sgd_hot_batch, 64);
    const bool run_cold = n_all > 0 && !(a.debug & 32);
    // rating-weighted mean |u|^2 for the device-side batch bound: every launch of a whole-matrix plan (above), every
    // second launch of a stratum-block plan (a quarter of the rows with local ratings is sampled: ~10 us)
    if (!stat_fresh && (pl.hot_stat_age++ & 1) == 0) {
      const DevCsr &m = e->mat[MFB_TRAIN];
      MFB_CUDA(cudaMemsetAsync(pl.hot_stat, 0, sizeof(double) * 2, e->stream));
      MFB_LAUNCH(sgd_hot_stat_kernel, e->sm_count * 8, 256, 0, e->stream, e->U, e->ld, m.rowptr, e->n_users, pl.hot_stat);
    }
    a.hot_stat = pl.hot_stat;
    a.hot_stab = (float)e->opt_sgd_hot_stab;
    int workers = 0;
    if (n_all > 0) MFB_TRY(prep(a, n_all, workers));
    // pacing and "shuffled kernel first" go with the long launches (chunk queue); a short launch of one stratum
    // block starts its hot CTAs first: the longest list is its critical path
    a.pace = (run_cold && a.chunk > 0 && e->opt_sgd_hot_pace) ? 1 : 0;
    MFB_CUDA(cudaEventRecord(e->ev_fork, e->stream));
    MFB_CUDA(cudaStreamWaitEvent(e->stream_hot, e->ev_fork, 0));
#define MFB_PICK(Q, E, S) do { if (!(a.debug & 16)) MFB_TRY((launch_hot<Q, E, S>(e, a, variant, n_lists, T))); } while (0)
    const bool deep = e->opt_sgd_hot_stages >= 8;
    // Launch order (measured, tools/hot_probe.py): long launches — shuffled kernel first; the short launch of a
    // single stratum block — hot CTAs first (the longest list is the critical path).  Bit 64 swaps the order.
    bool cold_first = a.chunk > 0;
    if (a.debug & 64) cold_first = !cold_first;
    if (cold_first && run_cold) MFB_TRY(launch(a, workers));
    if (nq <= 4) { if (deep) MFB_PICK(1, 4, 8); else MFB_PICK(1, 4, 4); }
    else if (nq <= 8) { if (deep) MFB_PICK(2, 4, 8); else MFB_PICK(2, 4, 4); }
    else if (nq <= 16) { if (deep) MFB_PICK(4, 4, 8); else MFB_PICK(4, 4, 4); }
    else if (nq <= 32) { if (deep) MFB_PICK(8, 4, 8); else MFB_PICK(8, 4, 4); }
    else MFB_PICK(8, 8, 4);
#undef MFB_PICK
    if (!cold_first && run_cold) MFB_TRY(launch(a, workers));
    MFB_CUDA(cudaEventRecord(e->ev_join, e->stream_hot));
    MFB_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    return 0;
  }
  if (n_all == 0) return 0;
  int workers = 0;
  MFB_TRY(prep(a, n_all, workers));
  return launch(a, workers);
}

int sgd_debug_hot_batch(mfb_engine *e, double out[3]) {
  out[0] = out[1] = out[2] = 0.0;
  if (!e->sgd.hot_stat) return 0;
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  MFB_CUDA(cudaMemcpy(out, e->sgd.hot_stat, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  return 0;
}

int sgd_debug_records(mfb_engine *e, int32_t a, int32_t b, int32_t *recs_out, int64_t *cold_nnz, int32_t *lists_out,
                      int32_t *n_lists) {
  const SgdPlan &pl = e->sgd;
  const size_t bid = (size_t)a * pl.P + b;
  const int64_t off = pl.blk_rat_off[bid], n = pl.blk_nnz[bid];
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  if (recs_out && n > 0)
    MFB_CUDA(cudaMemcpy(recs_out, reinterpret_cast<const int4 *>(pl.recs) + off, sizeof(int4) * (size_t)n, cudaMemcpyDeviceToHost));
  if (cold_nnz) *cold_nnz = pl.blk_cold_nnz.empty() ? n : pl.blk_cold_nnz[bid];
  const int32_t h = pl.blk_hot_cnt.empty() ? 0 : pl.blk_hot_cnt[bid];
  if (n_lists) *n_lists = h;
  if (lists_out && h > 0) {
    std::vector<int4> l((size_t)h);
    MFB_CUDA(cudaMemcpy(l.data(), reinterpret_cast<const int4 *>(pl.hot_lists) + pl.blk_hot_off[bid], sizeof(int4) * (size_t)h,
                        cudaMemcpyDeviceToHost));
    for (int32_t k = 0; k < h; k++) {
      lists_out[3 * k] = l[k].x;
      lists_out[3 * k + 1] = (int32_t)(l[k].y - off);
      lists_out[3 * k + 2] = l[k].z;
    }
  }
  return 0;
}

}  // namespace mfb
