// CCD++ rank-one sweeps over the CSR and CSC residual copies.
//
// Replaces the k-loop body of ModelMF::trainCCDPP (modelMF.cpp:1027-1121) and
// ::trainCCDPPFreqAdap (:1272-1375): residual add-back (:1034-1055), five alternations of the
// closed-form u_k (:1062-1074) and v_k (:1078-1090) updates with fp64 numerator/denominator,
// residual subtract (:1096-1116) and the column write-back (:1119-1120).  Both residual copies
// (CSR order and CSC order) are kept, as gk_csr_Dup does (:1013), so that every pass streams.
//
// Pure streaming, HBM-bound: one warp per row segment reads indices and residuals with
// coalesced loads and gathers the dense u_k / v_k vectors (a few MB, L2 resident).  Rows
// longer than a chunk are split over warps that combine through fp64 atomics.
#include "engine.h"

namespace mfb {

constexpr int kCcdChunk = 1024;
constexpr int kCcdDepth = 8;  // independent loads per lane in flight

// row-sharded runs: the other ranks' copies of the dense u_k / v_k vector being produced (peer memory);
// every new entry is stored into all of them, so the all-gather rides on the update pass itself
struct CcdPeers {
  float *p[kMaxRanks - 1];
  int n;
};
__device__ __forceinline__ void store_all(float *own, const CcdPeers &pe, int row, float v) {
  own[row] = v;
  for (int i = 0; i < pe.n; i++) pe.p[i][row] = v;
}

struct CcdPass {
  const int32_t *ind;
  float *res;
  const int32_t *seg_row, *seg_start, *seg_len, *seg_slot;
  int n_seg;
};

// res[j] += sign * own[row] * other[ind[j]]; only_multi: segments of split rows only (the fused update pass below
// has already subtracted the single-segment rows)
__global__ void __launch_bounds__(256) ccd_resid_kernel(const CcdPass p, const float *__restrict__ own,
                                                        const float *__restrict__ other, float sign, int only_multi) {
  const int lane = threadIdx.x & 31;
  const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= p.n_seg) return;
  if (only_multi && p.seg_slot[seg] < 0) return;
  const int row = p.seg_row[seg], start = p.seg_start[seg], len = p.seg_len[seg];
  const float a = sign * __ldg(own + row);
  // kCcdDepth independent (index, residual) loads per lane are issued before the first dependent gather: a pass is
  // bound by the latency of its short segments, not by bandwidth
  for (int base = 0; base < len; base += 32 * kCcdDepth) {
    int c[kCcdDepth];
    float r[kCcdDepth];
#pragma unroll
    for (int d = 0; d < kCcdDepth; d++) {
      const int j = base + d * 32 + lane;
      c[d] = j < len ? __ldg(p.ind + start + j) : -1;
      r[d] = j < len ? p.res[start + j] : 0.f;
    }
#pragma unroll
    for (int d = 0; d < kCcdDepth; d++) {
      if (c[d] < 0) continue;
      // own*other is rounded to fp32 before it is added (modelMF.cpp:1041), hence no fma here
      p.res[start + base + d * 32 + lane] = __fadd_rn(r[d], __fmul_rn(a, __ldg(other + c[d])));
    }
  }
}

// own[row] = sum res*other / (reg + sum other^2), fp64 accumulation of fp32 products.
// Fused forms (same statements in the same order, fewer trips through memory — a pass is bound by the latency of
// its short segments, 209 ratings per user row on the bench matrix, so every pass saved counts in full):
//   ADDBACK  : the residual add-back of modelMF.cpp:1034-1055, res += own_old[row] * other_old[ind], is applied on
//              the way in and written back; the update then uses the current `other` (for the column pass u_k has
//              already been updated once, other_old = its copy from before);
//   SUBTRACT : after the last update of v_k the subtraction of :1096-1116, res -= own_new[row] * other[ind], runs as a
//              second loop over the segment the warp has just read (L1 / L2 hits) — single-segment rows only, the
//              segments of split rows are subtracted by ccd_resid_kernel(only_multi) after ccd_finalize_kernel.
template <bool ADDBACK, bool SUBTRACT>
__global__ void __launch_bounds__(256) ccd_update_kernel(const CcdPass p, float *__restrict__ own,
                                                         const float *__restrict__ other,
                                                         const float *__restrict__ other_old, float reg,
                                                         double *__restrict__ acc, const Aux *__restrict__ aux_freq,
                                                         int freq_thresh, const CcdPeers pe) {
  const int lane = threadIdx.x & 31;
  const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (seg >= p.n_seg) return;
  const int row = p.seg_row[seg], start = p.seg_start[seg], len = p.seg_len[seg], slot = p.seg_slot[seg];
  double num = 0.0, den = 0.0;
  const float a_old = ADDBACK ? own[row] : 0.f;
  for (int base = 0; base < len; base += 32 * kCcdDepth) {
    int c[kCcdDepth];
    float r[kCcdDepth], o[kCcdDepth];
#pragma unroll
    for (int d = 0; d < kCcdDepth; d++) {
      const int j = base + d * 32 + lane;
      c[d] = j < len ? __ldg(p.ind + start + j) : -1;
      r[d] = j < len ? p.res[start + j] : 0.f;
    }
#pragma unroll
    for (int d = 0; d < kCcdDepth; d++) {
      o[d] = c[d] >= 0 ? __ldg(other + c[d]) : 0.f;
      if (ADDBACK && c[d] >= 0) {
        r[d] = __fadd_rn(r[d], __fmul_rn(a_old, __ldg(other_old + c[d])));
        p.res[start + base + d * 32 + lane] = r[d];
      }
    }
#pragma unroll
    for (int d = 0; d < kCcdDepth; d++) {
      num += (double)__fmul_rn(r[d], o[d]);
      den += (double)__fmul_rn(o[d], o[d]);
    }
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    num += __shfl_xor_sync(0xFFFFFFFFu, num, m);
    den += __shfl_xor_sync(0xFFFFFFFFu, den, m);
  }
  if (slot < 0) {
    float nv = (float)(num / ((double)reg + den));
    if (freq_thresh > 0 && aux_freq[row].freq < freq_thresh) nv = 0.f;
    if (lane == 0) store_all(own, pe, row, nv);
    if (SUBTRACT) {
      const float a = -nv;
      for (int base = 0; base < len; base += 32 * kCcdDepth) {
        int c[kCcdDepth];
        float r[kCcdDepth];
#pragma unroll
        for (int d = 0; d < kCcdDepth; d++) {
          const int j = base + d * 32 + lane;
          c[d] = j < len ? __ldg(p.ind + start + j) : -1;
          r[d] = j < len ? p.res[start + j] : 0.f;
        }
#pragma unroll
        for (int d = 0; d < kCcdDepth; d++)
          if (c[d] >= 0) p.res[start + base + d * 32 + lane] = __fadd_rn(r[d], __fmul_rn(a, __ldg(other + c[d])));
      }
    }
  } else if (lane == 0) {
    atomicAdd(acc + 2 * (size_t)slot, num);
    atomicAdd(acc + 2 * (size_t)slot + 1, den);
  }
}

__global__ void ccd_finalize_kernel(const int32_t *__restrict__ multi_row, int n_multi, double *__restrict__ acc,
                                    float *__restrict__ own, float reg, const Aux *__restrict__ aux_freq,
                                    int freq_thresh, const CcdPeers pe) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_multi) return;
  const int row = multi_row[s];
  float nv = (float)(acc[2 * (size_t)s] / ((double)reg + acc[2 * (size_t)s + 1]));
  if (freq_thresh > 0 && aux_freq[row].freq < freq_thresh) nv = 0.f;
  store_all(own, pe, row, nv);
  acc[2 * (size_t)s] = 0.0;
  acc[2 * (size_t)s + 1] = 0.0;
}

__global__ void col_extract_kernel(const float *__restrict__ F, int ld, int k, int lo, int hi, float *__restrict__ out,
                                   const CcdPeers pe) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) store_all(out, pe, i, F[(size_t)i * ld + k]);
}
__global__ void col_insert_kernel(float *__restrict__ F, int ld, int k, int lo, int hi, const float *__restrict__ in) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) F[(size_t)i * ld + k] = in[i];
}

int ccdpp_begin_impl(mfb_engine *e) {
  DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  size_t nn = (size_t)(m.nnz > 0 ? m.nnz : 1);
  if (!e->res_row) MFB_CUDA(dev_alloc(&e->res_row, sizeof(float) * nn));
  if (!e->res_col) MFB_CUDA(dev_alloc(&e->res_col, sizeof(float) * nn));
  if (!e->uk) MFB_CUDA(dev_alloc(&e->uk, sizeof(float) * e->n_users));
  if (!e->vk) MFB_CUDA(dev_alloc(&e->vk, sizeof(float) * e->n_items));
  if (!e->uk_old) MFB_CUDA(dev_alloc(&e->uk_old, sizeof(float) * e->n_users));
  // res = gk_csr_Dup(trainMat) (modelMF.cpp:1013); uFac.fill(0) (:1020)
  MFB_CUDA(cudaMemcpyAsync(e->res_row, m.rowval, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(e->res_col, m.colval, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToDevice, st));
  MFB_CUDA(cudaMemsetAsync(e->U, 0, sizeof(float) * (size_t)e->n_users * e->ld, st));
  if (!m.ccd_rows.built)
    MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, e->row_begin[MFB_USER], e->row_end[MFB_USER],
                           kCcdChunk, &m.ccd_rows));
  if (!m.ccd_cols.built)
    MFB_TRY(build_seg_plan(e, m.colptr, e->n_items, e->bad_item, e->row_begin[MFB_ITEM], e->row_end[MFB_ITEM],
                           kCcdChunk, &m.ccd_cols));
  size_t slots = (size_t)max(m.ccd_rows.n_multi, m.ccd_cols.n_multi);
  if (slots > e->ccd_acc_slots) {
    if (e->ccd_acc) MFB_CUDA(dev_free(e->ccd_acc));
    e->ccd_acc = nullptr;
    MFB_CUDA(dev_alloc(&e->ccd_acc, sizeof(double) * 2 * slots));
    e->ccd_acc_slots = slots;
  }
  if (e->ccd_acc) MFB_CUDA(cudaMemsetAsync(e->ccd_acc, 0, sizeof(double) * 2 * e->ccd_acc_slots, st));
  return 0;
}

int ccdpp_rank1_impl(mfb_engine *e, int32_t k, int first_iter, int32_t inner, float ureg, float ireg,
                     int32_t item_freq_thresh) {
  DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  const SegPlan &rp = m.ccd_rows, &cp = m.ccd_cols;
  CcdPass rows{m.rowind, e->res_row, rp.row, rp.start, rp.len, rp.slot, rp.n_seg};
  CcdPass cols{m.colind, e->res_col, cp.row, cp.start, cp.len, cp.slot, cp.n_seg};
  const int tb = 256, wpb = tb / 32;
  const int g_rows = (rp.n_seg + wpb - 1) / wpb, g_cols = (cp.n_seg + wpb - 1) / wpb;
  // row-sharded: this rank owns users [ulo, uhi) and items [ilo, ihi); every new u_k / v_k entry is also
  // stored into the peers' vectors and a flag barrier closes each pass
  const int ulo = e->row_begin[MFB_USER], uhi = e->row_end[MFB_USER], ilo = e->row_begin[MFB_ITEM], ihi = e->row_end[MFB_ITEM];
  CcdPeers pu, pv;
  pu.n = pv.n = 0;
  const Comm &c = e->comm;
  if (c.connected)
    for (int p = 0; p < c.world; p++)
      if (p != c.rank) { pu.p[pu.n++] = c.uk[p]; pv.p[pv.n++] = c.vk[p]; }
  // u_k = uFac.col(k); v_k = iFac.col(k)   (modelMF.cpp:1028-1029)
  if (uhi > ulo) MFB_LAUNCH(col_extract_kernel, (uhi - ulo + 255) / 256, 256, 0, st, e->U, e->ld, k, ulo, uhi, e->uk, pu);
  if (ihi > ilo) MFB_LAUNCH(col_extract_kernel, (ihi - ilo + 255) / 256, 256, 0, st, e->V, e->ld, k, ilo, ihi, e->vk, pv);
  MFB_TRY(comm_barrier_launch(e));
  // the FreqAdap rule zeroes v_k of infrequent items for k > 0 (modelMF.cpp:1336-1342)
  const int thresh = (item_freq_thresh > 0 && k > 0) ? item_freq_thresh : 0;
  const bool fuse_add = !first_iter && inner >= 1 && e->opt_ccd_fuse;          // add-back rides on the first updates
  const bool fuse_sub = inner >= (fuse_add ? 2 : 1) && e->opt_ccd_fuse;        // column subtract rides on the last v_k update
  if (!first_iter && !fuse_add) {
    if (g_rows) MFB_LAUNCH(ccd_resid_kernel, g_rows, tb, 0, st, rows, e->uk, e->vk, 1.0f, 0);
    if (g_cols) MFB_LAUNCH(ccd_resid_kernel, g_cols, tb, 0, st, cols, e->vk, e->uk, 1.0f, 0);
  }
  if (fuse_add)  // the column add-back needs u_k as it was before its first update
    MFB_CUDA(cudaMemcpyAsync(e->uk_old, e->uk, sizeof(float) * (size_t)e->n_users, cudaMemcpyDeviceToDevice, st));
  for (int s = 0; s < inner; s++) {
    if (g_rows) {
      if (s == 0 && fuse_add)
        MFB_LAUNCH((ccd_update_kernel<true, false>), g_rows, tb, 0, st, rows, e->uk, e->vk, e->vk, ureg, e->ccd_acc, e->aux_u, 0, pu);
      else
        MFB_LAUNCH((ccd_update_kernel<false, false>), g_rows, tb, 0, st, rows, e->uk, e->vk, e->vk, ureg, e->ccd_acc, e->aux_u, 0, pu);
    }
    if (rp.n_multi)
      MFB_LAUNCH(ccd_finalize_kernel, (rp.n_multi + 255) / 256, 256, 0, st, rp.multi_row, rp.n_multi, e->ccd_acc, e->uk,
                 ureg, e->aux_u, 0, pu);
    MFB_TRY(comm_barrier_launch(e));
    if (g_cols) {
      if (s == 0 && fuse_add)
        MFB_LAUNCH((ccd_update_kernel<true, false>), g_cols, tb, 0, st, cols, e->vk, e->uk, e->uk_old, ireg, e->ccd_acc, e->aux_i, thresh, pv);
      else if (s == inner - 1 && fuse_sub)
        MFB_LAUNCH((ccd_update_kernel<false, true>), g_cols, tb, 0, st, cols, e->vk, e->uk, e->uk, ireg, e->ccd_acc, e->aux_i, thresh, pv);
      else
        MFB_LAUNCH((ccd_update_kernel<false, false>), g_cols, tb, 0, st, cols, e->vk, e->uk, e->uk, ireg, e->ccd_acc, e->aux_i, thresh, pv);
    }
    if (cp.n_multi)
      MFB_LAUNCH(ccd_finalize_kernel, (cp.n_multi + 255) / 256, 256, 0, st, cp.multi_row, cp.n_multi, e->ccd_acc, e->vk,
                 ireg, e->aux_i, thresh, pv);
    MFB_TRY(comm_barrier_launch(e));
  }
  if (g_rows) MFB_LAUNCH(ccd_resid_kernel, g_rows, tb, 0, st, rows, e->uk, e->vk, -1.0f, 0);
  if (fuse_sub) {
    if (g_cols && cp.n_multi) MFB_LAUNCH(ccd_resid_kernel, g_cols, tb, 0, st, cols, e->vk, e->uk, -1.0f, 1);
  } else if (g_cols) {
    MFB_LAUNCH(ccd_resid_kernel, g_cols, tb, 0, st, cols, e->vk, e->uk, -1.0f, 0);
  }
  if (uhi > ulo) MFB_LAUNCH(col_insert_kernel, (uhi - ulo + 255) / 256, 256, 0, st, e->U, e->ld, k, ulo, uhi, e->uk);
  if (ihi > ilo) MFB_LAUNCH(col_insert_kernel, (ihi - ilo + 255) / 256, 256, 0, st, e->V, e->ld, k, ilo, ihi, e->vk);
  // the subtract passes above gather ALL of u_k / v_k; the next rank-one step starts by storing its column into every
  // peer's u_k / v_k — no rank may get there while a peer is still reading (write-after-read across ranks)
  MFB_TRY(comm_barrier_launch(e));
  return 0;
}

int comm_allgather_range(mfb_engine *e, int side, int first, int n);

int ccdpp_end_impl(mfb_engine *e) {
  if (e->comm.connected) {  // assemble the sharded factors on every rank
    MFB_TRY(comm_allgather_range(e, MFB_USER, e->row_begin[MFB_USER], e->row_end[MFB_USER] - e->row_begin[MFB_USER]));
    MFB_TRY(comm_allgather_range(e, MFB_ITEM, e->row_begin[MFB_ITEM], e->row_end[MFB_ITEM] - e->row_begin[MFB_ITEM]));
  }
  // no host synchronisation here: with several engines in one process the caller issues every step for all ranks
  // before any of them can finish (the barriers above wait for peers on the device); the cached allocator orders
  // the reuse of the freed blocks behind this stream's work
  dev_free(e->res_row); dev_free(e->res_col);
  e->res_row = e->res_col = nullptr;
  return 0;
}

}  // namespace mfb
