// Stratified SGD on the device: plan building (ratings bucketed once into the P x P stratum
// grid) and the update kernel.
//
// Reference statements replaced: the per-rating step of modelMF.cpp:83-105 (serial),
// :275-303 (stratified), :1747-1763 (Hogwild) and its weighted / truncated / Poisson-truncated
// forms modelInvPopMF.cpp:356-391, modelDropoutSigmoid.cpp:145-194,
// modelPoissonDropout.cpp:176-229.
//
// Kernel design (HBM/L2-bound gather-update, not GEMM-shaped):
//   * a sub-warp of G lanes owns one user's run of ratings inside a block (the reference's own
//     visiting order: user-major, CSR order inside the row, modelMF.cpp:279-281) and keeps the
//     user vector in registers for the whole run — u is read and written once per run instead
//     of once per rating;
//   * per rating the item vector is one coalesced 128-bit-per-lane load and store that
//     bypasses L1 (ld/st.global.cg) so other SMs' updates are seen at L2;
//   * (item, rating) pairs are fetched G at a time, one per lane, and broadcast by shuffle;
//     the next item vector is prefetched while the current one is reduced (shuffle tree);
//   * the IFWMF weight, the TMF rank and the Poisson-drawn rank are resolved per lane at fetch
//     time (off the dependent chain) and applied in the same pass;
//   * runs are sorted longest-first so that the serial chain of a heavy user starts at t = 0.
#include "engine.h"

#include <cub/cub.cuh>

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

int sgd_plan_build(mfb_engine *e, int32_t P, const int32_t *user_part, const int32_t *item_part) {
  SgdPlan &pl = e->sgd;
  pl.release();
  const DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  pl.P = P;
  pl.blk_seg_off.assign((size_t)P * P, 0);
  pl.blk_seg_cnt.assign((size_t)P * P, 0);
  pl.blk_nnz.assign((size_t)P * P, 0);
  if (P == 1 && user_part == nullptr) {
    // whole matrix = one block; runs are the CSR rows themselves
    SegPlan sp;
    MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, 0, e->n_users, 0, &sp));
    pl.item = m.rowind;
    pl.val = m.rowval;
    pl.owns_ratings = false;
    pl.n_seg = sp.n_seg;
    pl.seg_user = sp.row; pl.seg_start = sp.start; pl.seg_len = sp.len;
    sp.row = sp.start = sp.len = nullptr;
    sp.release();
    pl.nnz = m.nnz;
    pl.blk_seg_cnt[0] = pl.n_seg;
    pl.blk_nnz[0] = m.nnz;
    pl.built = true;
    return 0;
  }
  int64_t nnz = m.nnz;
  int32_t *d_up, *d_ip;
  MFB_CUDA(cudaMalloc(&d_up, sizeof(int32_t) * e->n_users));
  MFB_CUDA(cudaMalloc(&d_ip, sizeof(int32_t) * e->n_items));
  MFB_CUDA(cudaMemcpyAsync(d_up, user_part, sizeof(int32_t) * e->n_users, cudaMemcpyHostToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(d_ip, item_part, sizeof(int32_t) * e->n_items, cudaMemcpyHostToDevice, st));
  uint64_t *keys, *keys2;
  int32_t *idx, *idx2;
  size_t nn = (size_t)(nnz > 0 ? nnz : 1);
  MFB_CUDA(cudaMalloc(&keys, sizeof(uint64_t) * nn));
  MFB_CUDA(cudaMalloc(&keys2, sizeof(uint64_t) * nn));
  MFB_CUDA(cudaMalloc(&idx, sizeof(int32_t) * nn));
  MFB_CUDA(cudaMalloc(&idx2, sizeof(int32_t) * nn));
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
  MFB_CUDA(cudaMalloc(&pl.item, sizeof(int32_t) * nn));
  MFB_CUDA(cudaMalloc(&pl.val, sizeof(float) * nn));
  pl.owns_ratings = true;
  if (nnz > 0) MFB_LAUNCH(sgd_gather_kernel, gb, tb, 0, st, idx2, nnz, m.rowind, m.rowval, pl.item, pl.val);
  // runs of equal (block, user): reuse `keys` for the unique keys, `idx` for the run lengths
  int32_t *d_nruns;
  MFB_CUDA(cudaMalloc(&d_nruns, sizeof(int32_t)));
  MFB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, keys2, keys, idx, d_nruns, (int)nnz, st));
  MFB_TRY(ensure_scratch(e, tmp_bytes));
  MFB_CUDA(cub::DeviceRunLengthEncode::Encode(e->scratch, tmp_bytes, keys2, keys, idx, d_nruns, (int)nnz, st));
  int32_t nruns = 0;
  MFB_CUDA(cudaMemcpyAsync(&nruns, d_nruns, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_nruns);
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
    MFB_CUDA(cudaMalloc(&d_bounds, sizeof(int32_t) * (nblk + 1)));
    MFB_LAUNCH(sgd_blk_bounds_kernel, (nblk + 1 + 255) / 256, 256, 0, st, run_key, nruns, nblk, d_bounds);
    MFB_CUDA(cudaMemcpyAsync(bounds.data(), d_bounds, sizeof(int32_t) * (nblk + 1), cudaMemcpyDeviceToHost, st));
    MFB_CUDA(cudaStreamSynchronize(st));
    cudaFree(d_bounds);
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
  }
  size_t ns = (size_t)(nseg > 0 ? nseg : 1);
  MFB_CUDA(cudaMalloc(&pl.seg_user, sizeof(int32_t) * ns));
  MFB_CUDA(cudaMalloc(&pl.seg_start, sizeof(int32_t) * ns));
  MFB_CUDA(cudaMalloc(&pl.seg_len, sizeof(int32_t) * ns));
  if (nseg > 0) {
    // longest-first inside every block
    uint64_t *key2 = keys2;  // sorted rating keys no longer needed
    uint64_t *key2_out = keys2 + nseg;  // nnz >= 2 * nseg is not guaranteed: allocate separately if short
    int32_t *sidx, *sidx_out;
    bool own_key_out = (size_t)nnz < 2 * (size_t)nseg;
    if (own_key_out) MFB_CUDA(cudaMalloc(&key2_out, sizeof(uint64_t) * ns));
    MFB_CUDA(cudaMalloc(&sidx, sizeof(int32_t) * 2 * ns));
    sidx_out = sidx + nseg;
    int gs = (nseg + tb - 1) / tb;
    MFB_LAUNCH(sgd_seg_key_kernel, gs, tb, 0, st, run_key, run_len, nseg, key2, sidx);
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key2, key2_out, sidx, sidx_out, nseg, 0, end_bit, st));
    MFB_TRY(ensure_scratch(e, tmp_bytes));
    MFB_CUDA(cub::DeviceRadixSort::SortPairs(e->scratch, tmp_bytes, key2, key2_out, sidx, sidx_out, nseg, 0, end_bit, st));
    MFB_LAUNCH(sgd_seg_gather_kernel, gs, tb, 0, st, sidx_out, nseg, run_key, run_start, run_len, pl.seg_user,
               pl.seg_start, pl.seg_len);
    MFB_CUDA(cudaStreamSynchronize(st));
    if (own_key_out) cudaFree(key2_out);
    cudaFree(sidx);
  }
  MFB_CUDA(cudaStreamSynchronize(st));
  cudaFree(keys); cudaFree(keys2); cudaFree(idx); cudaFree(idx2); cudaFree(d_up); cudaFree(d_ip);
  pl.built = true;
  return 0;
}

// ---- update kernel ---------------------------------------------------------------------------
struct SgdArgs {
  float *U, *V;
  int nq;  // float4 words per factor row (ld / 4)
  int rank;
  const int32_t *item;
  const float *val;
  const int32_t *seg_user, *seg_start, *seg_len;
  int nb, max_cnt;
  int32_t off[kMaxBlocks], cnt[kMaxBlocks];
  float lr, ureg, ireg;
  const Aux *aux_u, *aux_i;
  const float *cdf;
  uint64_t seed, counter;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
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

template <int G, int VPL, int VARIANT>
__global__ void __launch_bounds__(128) sgd_update_kernel(const SgdArgs a) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1);  // lane inside the sub-warp
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int b = (int)(gid % a.nb);
  const int sidx = (int)(gid / a.nb);
  const bool active = sidx < a.cnt[b];
  int user = 0, start = 0, len = 0;
  if (active) {
    int s = a.off[b] + sidx;
    user = a.seg_user[s];
    start = a.seg_start[s];
    len = a.seg_len[s];
  }
  int maxlen = len;
#pragma unroll
  for (int m = 16; m >= G; m >>= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, m));
  if (maxlen == 0) return;  // warp-uniform

  float4 u[VPL];
  bool own[VPL];  // this lane holds a real float4 word of the row
  float4 *urow = reinterpret_cast<float4 *>(a.U) + (size_t)user * a.nq;
#pragma unroll
  for (int c = 0; c < VPL; c++) {
    own[c] = (c * G + sl) < a.nq;
    u[c] = (active && own[c]) ? __ldcg(urow + c * G + sl) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int ufreq = 0, upay = 0;
  if (VARIANT != MFB_MF && active) {
    Aux au = a.aux_u[user];
    ufreq = au.freq;
    upay = au.train;
  }
  const float4 *Vq = reinterpret_cast<const float4 *>(a.V);
  float4 *Vw = reinterpret_cast<float4 *>(a.V);

  // fetch one (item, rating[, payload]) per lane
  auto fetch = [&](int j, int &it, float &rt, int &pay) {
    it = 0; rt = 0.f; pay = 0;
    if (j < len) {
      it = __ldg(a.item + start + j);
      rt = __ldg(a.val + start + j);
      if (VARIANT != MFB_MF) {
        Aux ai = a.aux_i[it];
        // the rarer side decides (modelInvPopMF.cpp:164-166, modelDropoutSigmoid.cpp:158)
        pay = (ufreq < ai.freq) ? upay : ai.train;
        if (VARIANT == MFB_TMFDROPOUT) pay = poisson_rank(a.cdf, a.rank, pay, a.seed, a.counter, (uint32_t)(start + j));
      }
    }
  };

  int n_it, n_pay;
  float n_rt;
  fetch(sl, n_it, n_rt, n_pay);
  float4 vn[VPL];
  {
    int it0 = __shfl_sync(kFull, n_it, 0, G);
#pragma unroll
    for (int c = 0; c < VPL; c++)
      vn[c] = (len > 0 && own[c]) ? __ldcg(Vq + (size_t)it0 * a.nq + c * G + sl) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float lr = a.lr, two_ureg = 2.0f * a.ureg, two_ireg = 2.0f * a.ireg;

  for (int j0 = 0; j0 < maxlen; j0 += G) {
    const int c_it = n_it, c_pay = n_pay;
    const float c_rt = n_rt;
    fetch(j0 + G + sl, n_it, n_rt, n_pay);
#pragma unroll 4
    for (int t = 0; t < G; t++) {
      const int j = j0 + t;
      if (j >= maxlen) break;  // warp-uniform
      const int it = __shfl_sync(kFull, c_it, t, G);
      const float rt = __shfl_sync(kFull, c_rt, t, G);
      int pay = 0;
      if (VARIANT != MFB_MF) pay = __shfl_sync(kFull, c_pay, t, G);
      const int itn_same = __shfl_sync(kFull, c_it, (t + 1) & (G - 1), G);
      const int itn_next = __shfl_sync(kFull, n_it, 0, G);
      const int itn = (t + 1 < G) ? itn_same : itn_next;
      const bool on = j < len;
      float4 v[VPL];
#pragma unroll
      for (int c = 0; c < VPL; c++) v[c] = vn[c];
      if (j + 1 < len) {
#pragma unroll
        for (int c = 0; c < VPL; c++)
          if (own[c]) vn[c] = __ldcg(Vq + (size_t)itn * a.nq + c * G + sl);
      }
      int k = a.rank;  // number of leading dimensions this update touches
      if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) k = pay;
      float p = 0.f;
#pragma unroll
      for (int c = 0; c < VPL; c++) {
        const int base = (c * G + sl) * 4;
        if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
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
      for (int m = G / 2; m >= 1; m >>= 1) p += __shfl_xor_sync(kFull, p, m);
      float g = rt - p;  // diff
      if (VARIANT == MFB_IFWMF) g *= __int_as_float(pay);
      const float m2g = -2.0f * g;
      if (on) {
#pragma unroll
        for (int c = 0; c < VPL; c++) {
          if (!own[c]) continue;
          const int base = (c * G + sl) * 4;
          float4 un, vv = v[c];
          // u -= lr * (-2 g v + 2 ureg u);  v -= lr * (-2 g u_new + 2 ireg v)   (modelMF.cpp:95-103)
          un.x = fmaf(-lr, fmaf(two_ureg, u[c].x, m2g * vv.x), u[c].x);
          un.y = fmaf(-lr, fmaf(two_ureg, u[c].y, m2g * vv.y), u[c].y);
          un.z = fmaf(-lr, fmaf(two_ureg, u[c].z, m2g * vv.z), u[c].z);
          un.w = fmaf(-lr, fmaf(two_ureg, u[c].w, m2g * vv.w), u[c].w);
          float4 vo;
          vo.x = fmaf(-lr, fmaf(two_ireg, vv.x, m2g * un.x), vv.x);
          vo.y = fmaf(-lr, fmaf(two_ireg, vv.y, m2g * un.y), vv.y);
          vo.z = fmaf(-lr, fmaf(two_ireg, vv.z, m2g * un.z), vv.z);
          vo.w = fmaf(-lr, fmaf(two_ireg, vv.w, m2g * un.w), vv.w);
          if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
            if (base >= k) continue;  // nothing of this word is touched
            if (base + 1 >= k) { un.y = u[c].y; vo.y = vv.y; }
            if (base + 2 >= k) { un.z = u[c].z; vo.z = vv.z; }
            if (base + 3 >= k) { un.w = u[c].w; vo.w = vv.w; }
          }
          u[c] = un;
          __stcg(Vw + (size_t)it * a.nq + c * G + sl, vo);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int c = 0; c < VPL; c++)
      if (own[c]) __stcg(urow + c * G + sl, u[c]);
  }
}

template <int G, int VPL>
static int launch_variant(mfb_engine *e, const SgdArgs &a, int variant, int64_t n_groups) {
  const int tb = 128;
  const int64_t threads = n_groups * G;
  const unsigned grid = (unsigned)((threads + tb - 1) / tb);
  switch (variant) {
    case MFB_MF: MFB_LAUNCH((sgd_update_kernel<G, VPL, MFB_MF>), grid, tb, 0, e->stream, a); break;
    case MFB_IFWMF: MFB_LAUNCH((sgd_update_kernel<G, VPL, MFB_IFWMF>), grid, tb, 0, e->stream, a); break;
    case MFB_TMF: MFB_LAUNCH((sgd_update_kernel<G, VPL, MFB_TMF>), grid, tb, 0, e->stream, a); break;
    default: MFB_LAUNCH((sgd_update_kernel<G, VPL, MFB_TMFDROPOUT>), grid, tb, 0, e->stream, a); break;
  }
  return 0;
}

int sgd_subepoch_launch(mfb_engine *e, const int32_t *blocks, int32_t nb, int variant, float lr, float ureg,
                        float ireg, uint64_t seed, uint64_t counter) {
  const SgdPlan &pl = e->sgd;
  SgdArgs a;
  a.U = e->U; a.V = e->V;
  a.nq = e->ld / 4;
  a.rank = e->rank;
  a.item = pl.item; a.val = pl.val;
  a.seg_user = pl.seg_user; a.seg_start = pl.seg_start; a.seg_len = pl.seg_len;
  a.nb = nb;
  a.max_cnt = 0;
  for (int i = 0; i < nb; i++) {
    size_t bid = (size_t)blocks[2 * i] * pl.P + blocks[2 * i + 1];
    a.off[i] = pl.blk_seg_off[bid];
    a.cnt[i] = pl.blk_seg_cnt[bid];
    if (a.cnt[i] > a.max_cnt) a.max_cnt = a.cnt[i];
  }
  a.lr = lr; a.ureg = ureg; a.ireg = ireg;
  a.aux_u = e->aux_u; a.aux_i = e->aux_i; a.cdf = e->poisson_cdf;
  a.seed = seed; a.counter = counter;
  if (a.max_cnt == 0) return 0;
  const int64_t n_groups = (int64_t)a.max_cnt * nb;
  // sub-warp width: the smallest power of two of lanes that covers the row with <= 2 words per lane
  const int nq = a.nq;
  if (nq <= 2) return launch_variant<2, 1>(e, a, variant, n_groups);
  if (nq <= 4) return launch_variant<4, 1>(e, a, variant, n_groups);
  if (nq <= 8) return launch_variant<8, 1>(e, a, variant, n_groups);
  if (nq <= 16) return launch_variant<16, 1>(e, a, variant, n_groups);
  if (nq <= 32) return launch_variant<32, 1>(e, a, variant, n_groups);
  return launch_variant<32, 2>(e, a, variant, n_groups);
}

}  // namespace mfb
