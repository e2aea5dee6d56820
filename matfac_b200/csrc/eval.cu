// Fused per-epoch evaluation: masked (weighted) squared error, count and both factor norms.
//
// Replaces Model::RMSE (model.cpp:214-251), Model::objective (model.cpp:1770-1815) and
// ModelInvPopMF::objective (modelInvPopMF.cpp:3-55), which the reference runs after every
// epoch from isTerminateModel (model.cpp:1476-1480).  estRating is the virtual of
// model.cpp:547 / modelDropoutSigmoid.cpp:5-24 / modelPoissonDropout.cpp:5-23.
//
// HBM/L2-bound gather: a sub-warp owns a chunk of one user's ratings, keeps u in registers,
// gathers v with 128-bit loads, reduces the dot by shuffles and accumulates the squared error
// in double; per-CTA partials are combined in a fixed order (deterministic result).
#include "engine.h"

namespace mfb {

struct EvalArgs {
  const float *U, *V;
  int nq, rank;
  const int32_t *ind;
  const float *val;
  const int32_t *seg_row, *seg_start, *seg_len;
  int n_seg;
  const uint8_t *bad_item;
  const Aux *aux_u, *aux_i;
  int weighted;
  double *partial;  // [grid][2]
};

template <int G, int VPL, int VARIANT>
__global__ void __launch_bounds__(128) eval_sse_kernel(const EvalArgs a) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1);
  const int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
  const bool active = seg < a.n_seg;
  int user = 0, start = 0, len = 0;
  if (active) {
    user = a.seg_row[seg];
    start = a.seg_start[seg];
    len = a.seg_len[seg];
  }
  int maxlen = len;
#pragma unroll
  for (int m = 16; m >= G; m >>= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, m));
  double sse = 0.0, cnt = 0.0;
  if (maxlen > 0) {
    float4 u[VPL];
    bool own[VPL];
    const float4 *urow = reinterpret_cast<const float4 *>(a.U) + (size_t)user * a.nq;
#pragma unroll
    for (int c = 0; c < VPL; c++) {
      own[c] = (c * G + sl) < a.nq;
      u[c] = (active && own[c]) ? __ldcg(urow + c * G + sl) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int ufreq = 0, upred = 0, utrain = 0;
    if (VARIANT != MFB_MF && active) {
      Aux au = a.aux_u[user];
      ufreq = au.freq; upred = au.pred; utrain = au.train;
    }
    const float4 *Vq = reinterpret_cast<const float4 *>(a.V);
    for (int j0 = 0; j0 < maxlen; j0 += G) {
      int c_it = -1, c_pay = 0;
      float c_rt = 0.f;
      if (j0 + sl < len) {
        int it = __ldg(a.ind + start + j0 + sl);
        c_rt = __ldg(a.val + start + j0 + sl);
        if (!a.bad_item[it]) {  // masks are applied before any aux gather (model.cpp:235)
          c_it = it;
          if (VARIANT != MFB_MF) {
            Aux ai = a.aux_i[it];
            const bool user_side = ufreq < ai.freq;
            if (VARIANT == MFB_IFWMF) c_pay = user_side ? utrain : ai.train;
            else c_pay = user_side ? upred : ai.pred;
          }
        }
      }
#pragma unroll 4
      for (int t = 0; t < G; t++) {
        if (j0 + t >= maxlen) break;
        const int it = __shfl_sync(kFull, c_it, t, G);
        const float rt = __shfl_sync(kFull, c_rt, t, G);
        int pay = 0;
        if (VARIANT != MFB_MF) pay = __shfl_sync(kFull, c_pay, t, G);
        const bool on = it >= 0;
        int k = a.rank;
        if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) k = pay;
        float p = 0.f;
        if (on) {
#pragma unroll
          for (int c = 0; c < VPL; c++) {
            if (!own[c]) continue;
            const float4 v = __ldcg(Vq + (size_t)it * a.nq + c * G + sl);
            const int base = (c * G + sl) * 4;
            if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
              p += (base + 0 < k ? u[c].x * v.x : 0.f) + (base + 1 < k ? u[c].y * v.y : 0.f) +
                   (base + 2 < k ? u[c].z * v.z : 0.f) + (base + 3 < k ? u[c].w * v.w : 0.f);
            } else {
              p = fmaf(u[c].x, v.x, p);
              p = fmaf(u[c].y, v.y, p);
              p = fmaf(u[c].z, v.z, p);
              p = fmaf(u[c].w, v.w, p);
            }
          }
        }
#pragma unroll
        for (int m = G / 2; m >= 1; m >>= 1) p += __shfl_xor_sync(kFull, p, m);
        if (on && sl == 0) {
          const double diff = (double)rt - (double)p;
          double w = 1.0;
          if (VARIANT == MFB_IFWMF && a.weighted) w = (double)__int_as_float(pay);
          sse += w * diff * diff;
          cnt += 1.0;
        }
      }
    }
  }
  // CTA reduction (fixed order)
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    sse += __shfl_xor_sync(kFull, sse, m);
    cnt += __shfl_xor_sync(kFull, cnt, m);
  }
  __shared__ double s_sse[4], s_cnt[4];
  const int w = threadIdx.x >> 5;
  if (lane == 0) { s_sse[w] = sse; s_cnt[w] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, c = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) { s += s_sse[i]; c += s_cnt[i]; }
    a.partial[2 * (size_t)blockIdx.x] = s;
    a.partial[2 * (size_t)blockIdx.x + 1] = c;
  }
}

// sum over valid rows in [lo,hi) of the fp32 row dot (model.cpp:1792,1805), accumulated in double
__global__ void __launch_bounds__(256) row_norm_kernel(const float *__restrict__ F, int ld, int lo, int hi,
                                                       const uint8_t *__restrict__ bad, double *__restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  double acc = 0.0;
  for (int r = lo + wid; r < hi; r += warps) {
    if (bad[r]) continue;
    float s = 0.f;
    for (int k = lane; k < ld; k += 32) {
      float x = __ldcg(F + (size_t)r * ld + k);
      s = fmaf(x, x, s);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, m);
    acc += (double)s;
  }
  __shared__ double sh[8];
  if (lane == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sh[i];
    partial[blockIdx.x] = t;
  }
}

// out[0..1] = sum of the sse/count partials; out[2], out[3] = sums of the two norm partial arrays
__global__ void __launch_bounds__(256) eval_final_kernel(const double *__restrict__ p_sse, int n_sse,
                                                         const double *__restrict__ p_un, int n_un,
                                                         const double *__restrict__ p_in, int n_in,
                                                         double *__restrict__ out) {
  __shared__ double sh[4][256];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int i = threadIdx.x; i < n_sse; i += 256) { a0 += p_sse[2 * (size_t)i]; a1 += p_sse[2 * (size_t)i + 1]; }
  for (int i = threadIdx.x; i < n_un; i += 256) a2 += p_un[i];
  for (int i = threadIdx.x; i < n_in; i += 256) a3 += p_in[i];
  sh[0][threadIdx.x] = a0; sh[1][threadIdx.x] = a1; sh[2][threadIdx.x] = a2; sh[3][threadIdx.x] = a3;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int q = 0; q < 4; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x < 4) out[threadIdx.x] = sh[threadIdx.x][0];
}

template <int G, int VPL>
static int launch_eval(mfb_engine *e, const EvalArgs &a, int variant, unsigned grid) {
  switch (variant) {
    case MFB_MF: MFB_LAUNCH((eval_sse_kernel<G, VPL, MFB_MF>), grid, 128, 0, e->stream, a); break;
    case MFB_IFWMF: MFB_LAUNCH((eval_sse_kernel<G, VPL, MFB_IFWMF>), grid, 128, 0, e->stream, a); break;
    case MFB_TMF: MFB_LAUNCH((eval_sse_kernel<G, VPL, MFB_TMF>), grid, 128, 0, e->stream, a); break;
    default: MFB_LAUNCH((eval_sse_kernel<G, VPL, MFB_TMFDROPOUT>), grid, 128, 0, e->stream, a); break;
  }
  return 0;
}

int eval_launch(mfb_engine *e, int which, int factors, int variant, int weighted, int want_norms, double out[4]) {
  DevCsr &m = e->mat[which];
  if (!m.eval_rows.built)
    MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, e->row_begin[MFB_USER], e->row_end[MFB_USER], 512,
                           &m.eval_rows));
  const SegPlan &sp = m.eval_rows;
  const int nq = e->ld / 4;
  int G = 2;
  while (G < 32 && G < nq) G <<= 1;
  const int segs_per_cta = 128 / G;
  const int grid_sse = sp.n_seg > 0 ? (sp.n_seg + segs_per_cta - 1) / segs_per_cta : 0;
  const int grid_norm = want_norms ? 2 * e->sm_count : 0;
  const int need = 2 * grid_sse + 2 * grid_norm + 8;
  if (need > e->eval_partial_cap) {
    if (e->eval_partial) MFB_CUDA(dev_free(e->eval_partial));
    e->eval_partial = nullptr;
    MFB_CUDA(dev_alloc(&e->eval_partial, sizeof(double) * (size_t)need));
    e->eval_partial_cap = need;
  }
  double *p_sse = e->eval_partial, *p_un = p_sse + 2 * (size_t)grid_sse, *p_in = p_un + grid_norm;
  EvalArgs a;
  a.U = factors == MFB_BEST ? e->bestU : e->U;
  a.V = factors == MFB_BEST ? e->bestV : e->V;
  a.nq = nq; a.rank = e->rank;
  a.ind = m.rowind; a.val = m.rowval;
  a.seg_row = sp.row; a.seg_start = sp.start; a.seg_len = sp.len; a.n_seg = sp.n_seg;
  a.bad_item = e->bad_item;
  a.aux_u = e->aux_u; a.aux_i = e->aux_i;
  a.weighted = weighted;
  a.partial = p_sse;
  if (grid_sse > 0) {
    if (nq <= 2) MFB_TRY((launch_eval<2, 1>(e, a, variant, grid_sse)));
    else if (nq <= 4) MFB_TRY((launch_eval<4, 1>(e, a, variant, grid_sse)));
    else if (nq <= 8) MFB_TRY((launch_eval<8, 1>(e, a, variant, grid_sse)));
    else if (nq <= 16) MFB_TRY((launch_eval<16, 1>(e, a, variant, grid_sse)));
    else if (nq <= 32) MFB_TRY((launch_eval<32, 1>(e, a, variant, grid_sse)));
    else MFB_TRY((launch_eval<32, 2>(e, a, variant, grid_sse)));
  }
  if (want_norms) {
    MFB_LAUNCH(row_norm_kernel, grid_norm, 256, 0, e->stream, a.U, e->ld, e->row_begin[MFB_USER], e->row_end[MFB_USER],
               e->bad_user, p_un);
    MFB_LAUNCH(row_norm_kernel, grid_norm, 256, 0, e->stream, a.V, e->ld, e->row_begin[MFB_ITEM], e->row_end[MFB_ITEM],
               e->bad_item, p_in);
  }
  MFB_LAUNCH(eval_final_kernel, 1, 256, 0, e->stream, p_sse, grid_sse, p_un, grid_norm, p_in, grid_norm, e->eval_out);
  MFB_CUDA(cudaMemcpyAsync(e->eval_out_host, e->eval_out, sizeof(double) * 4, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < 4; i++) out[i] = e->eval_out_host[i];
  MFB_TRY(comm_check_error(e));  // an evaluation over item rows that never arrived must not be reported
  return 0;
}


// ---- grouped evaluation: squared error and count per item group and per user group in ONE pass --------
// Replaces the eight filtered passes of quartileRMSEs (main.cpp:700-768): Model::RMSE(mat, filtItems, ...)
// (model.cpp:348-394), ::SE (:397-443) and ::RMSEU (:446-486), which probe an unordered_set per rating.
// Every id carries a group byte (255 = in no group); a rating adds its squared error to the bin of its
// item's group and to the bin of its user's group.  Bins live in registers (compare-select, no indexing).
constexpr int kEvalGroups = 8;

struct EvalGroupArgs {
  EvalArgs e;
  const uint8_t *user_group, *item_group;
  double *partial;  // [grid][2 sides][kEvalGroups][2]
};

template <int G, int VPL, int VARIANT>
__global__ void __launch_bounds__(128) eval_group_kernel(const EvalGroupArgs ga) {
  constexpr unsigned kFull = 0xFFFFFFFFu;
  const EvalArgs &a = ga.e;
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1);
  const int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
  const bool active = seg < a.n_seg;
  int user = 0, start = 0, len = 0;
  if (active) {
    user = a.seg_row[seg];
    start = a.seg_start[seg];
    len = a.seg_len[seg];
  }
  int maxlen = len;
#pragma unroll
  for (int m = 16; m >= G; m >>= 1) maxlen = max(maxlen, __shfl_xor_sync(kFull, maxlen, m));
  double isse[kEvalGroups], icnt[kEvalGroups], usse = 0.0, ucnt = 0.0;
#pragma unroll
  for (int g = 0; g < kEvalGroups; g++) isse[g] = icnt[g] = 0.0;
  const int ugrp = active ? ga.user_group[user] : 255;
  if (maxlen > 0) {
    float4 u[VPL];
    bool own[VPL];
    const float4 *urow = reinterpret_cast<const float4 *>(a.U) + (size_t)user * a.nq;
#pragma unroll
    for (int c = 0; c < VPL; c++) {
      own[c] = (c * G + sl) < a.nq;
      u[c] = (active && own[c]) ? __ldcg(urow + c * G + sl) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int ufreq = 0, upred = 0;
    if (VARIANT != MFB_MF && active) {
      Aux au = a.aux_u[user];
      ufreq = au.freq; upred = au.pred;
    }
    const float4 *Vq = reinterpret_cast<const float4 *>(a.V);
    for (int j0 = 0; j0 < maxlen; j0 += G) {
      int c_it = -1, c_pay = 0, c_grp = 255;
      float c_rt = 0.f;
      if (j0 + sl < len) {
        int it = __ldg(a.ind + start + j0 + sl);
        c_rt = __ldg(a.val + start + j0 + sl);
        if (!a.bad_item[it]) {
          c_it = it;
          c_grp = ga.item_group[it];
          if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
            Aux ai = a.aux_i[it];
            c_pay = (ufreq < ai.freq) ? upred : ai.pred;
          }
        }
      }
#pragma unroll 4
      for (int t = 0; t < G; t++) {
        if (j0 + t >= maxlen) break;
        const int it = __shfl_sync(kFull, c_it, t, G);
        const float rt = __shfl_sync(kFull, c_rt, t, G);
        const int grp = __shfl_sync(kFull, c_grp, t, G);
        int k = a.rank;
        if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) k = __shfl_sync(kFull, c_pay, t, G);
        const bool on = it >= 0;
        float p = 0.f;
        if (on) {
#pragma unroll
          for (int c = 0; c < VPL; c++) {
            if (!own[c]) continue;
            const float4 v = __ldcg(Vq + (size_t)it * a.nq + c * G + sl);
            const int base = (c * G + sl) * 4;
            if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
              p += (base + 0 < k ? u[c].x * v.x : 0.f) + (base + 1 < k ? u[c].y * v.y : 0.f) +
                   (base + 2 < k ? u[c].z * v.z : 0.f) + (base + 3 < k ? u[c].w * v.w : 0.f);
            } else {
              p = fmaf(u[c].x, v.x, p);
              p = fmaf(u[c].y, v.y, p);
              p = fmaf(u[c].z, v.z, p);
              p = fmaf(u[c].w, v.w, p);
            }
          }
        }
#pragma unroll
        for (int m = G / 2; m >= 1; m >>= 1) p += __shfl_xor_sync(kFull, p, m);
        if (on && sl == 0) {
          const double diff = (double)rt - (double)p;
          const double d2 = diff * diff;
          usse += d2;
          ucnt += 1.0;
#pragma unroll
          for (int g = 0; g < kEvalGroups; g++)
            if (grp == g) { isse[g] += d2; icnt[g] += 1.0; }
        }
      }
    }
  }
  // CTA reduction through shared memory (fixed order): bins [side][group][sse|count]
  __shared__ double flat[2 * kEvalGroups * 2];
  for (int i = threadIdx.x; i < 2 * kEvalGroups * 2; i += blockDim.x) flat[i] = 0.0;
  __syncthreads();
  // one sub-warp leader at a time adds its bins: serialised per CTA, but only 128 / G leaders exist
  for (int turn = 0; turn < (int)blockDim.x / G; turn++) {
    if (sl == 0 && (int)(threadIdx.x / G) == turn) {
#pragma unroll
      for (int g = 0; g < kEvalGroups; g++) {
        flat[(0 * kEvalGroups + g) * 2 + 0] += isse[g];
        flat[(0 * kEvalGroups + g) * 2 + 1] += icnt[g];
      }
      if (ugrp < kEvalGroups) {
        flat[(1 * kEvalGroups + ugrp) * 2 + 0] += usse;
        flat[(1 * kEvalGroups + ugrp) * 2 + 1] += ucnt;
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < 2 * kEvalGroups * 2; i += blockDim.x)
    ga.partial[(size_t)blockIdx.x * (2 * kEvalGroups * 2) + i] = flat[i];
}

// out[i] = sum over CTAs of partial[cta][i], i < 32 (fixed order: deterministic)
__global__ void __launch_bounds__(256) eval_group_final_kernel(const double *__restrict__ partial, int n_cta, double *__restrict__ out) {
  __shared__ double sh[8][32];
  const int bin = threadIdx.x & 31, stripe = threadIdx.x >> 5;
  double acc = 0.0;
  for (int c = stripe; c < n_cta; c += 8) acc += partial[(size_t)c * 32 + bin];
  sh[stripe][bin] = acc;
  __syncthreads();
  if (stripe == 0) {
    double t = 0.0;
    for (int s2 = 0; s2 < 8; s2++) t += sh[s2][bin];
    out[bin] = t;
  }
}

template <int G, int VPL>
static int launch_eval_group(mfb_engine *e, const EvalGroupArgs &a, int variant, unsigned grid) {
  switch (variant) {
    case MFB_MF: MFB_LAUNCH((eval_group_kernel<G, VPL, MFB_MF>), grid, 128, 0, e->stream, a); break;
    case MFB_IFWMF: MFB_LAUNCH((eval_group_kernel<G, VPL, MFB_IFWMF>), grid, 128, 0, e->stream, a); break;
    case MFB_TMF: MFB_LAUNCH((eval_group_kernel<G, VPL, MFB_TMF>), grid, 128, 0, e->stream, a); break;
    default: MFB_LAUNCH((eval_group_kernel<G, VPL, MFB_TMFDROPOUT>), grid, 128, 0, e->stream, a); break;
  }
  return 0;
}

int eval_groups_launch(mfb_engine *e, int which, int factors, int variant, const uint8_t *user_group,
                       const uint8_t *item_group, double *out /* [2][kEvalGroups][2] */) {
  DevCsr &m = e->mat[which];
  if (!m.eval_rows.built)
    MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, e->row_begin[MFB_USER], e->row_end[MFB_USER], 512,
                           &m.eval_rows));
  const SegPlan &sp = m.eval_rows;
  for (int i = 0; i < 2 * kEvalGroups * 2; i++) out[i] = 0.0;
  if (sp.n_seg == 0) return 0;
  const int nq = e->ld / 4;
  int G = 2;
  while (G < 32 && G < nq) G <<= 1;
  const int segs_per_cta = 128 / G;
  const int grid = (sp.n_seg + segs_per_cta - 1) / segs_per_cta;
  uint8_t *d_groups;
  double *d_partial;
  MFB_CUDA(dev_alloc(&d_groups, (size_t)e->n_users + e->n_items));
  MFB_CUDA(dev_alloc(&d_partial, sizeof(double) * 32 * ((size_t)grid + 1)));
  MFB_CUDA(cudaMemcpyAsync(d_groups, user_group, e->n_users, cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(d_groups + e->n_users, item_group, e->n_items, cudaMemcpyHostToDevice, e->stream));
  EvalGroupArgs ga;
  EvalArgs &a = ga.e;
  a.U = factors == MFB_BEST ? e->bestU : e->U;
  a.V = factors == MFB_BEST ? e->bestV : e->V;
  a.nq = nq; a.rank = e->rank;
  a.ind = m.rowind; a.val = m.rowval;
  a.seg_row = sp.row; a.seg_start = sp.start; a.seg_len = sp.len; a.n_seg = sp.n_seg;
  a.bad_item = e->bad_item;
  a.aux_u = e->aux_u; a.aux_i = e->aux_i;
  a.weighted = 0;
  a.partial = nullptr;
  ga.user_group = d_groups;
  ga.item_group = d_groups + e->n_users;
  ga.partial = d_partial;
  if (nq <= 2) MFB_TRY((launch_eval_group<2, 1>(e, ga, variant, grid)));
  else if (nq <= 4) MFB_TRY((launch_eval_group<4, 1>(e, ga, variant, grid)));
  else if (nq <= 8) MFB_TRY((launch_eval_group<8, 1>(e, ga, variant, grid)));
  else if (nq <= 16) MFB_TRY((launch_eval_group<16, 1>(e, ga, variant, grid)));
  else if (nq <= 32) MFB_TRY((launch_eval_group<32, 1>(e, ga, variant, grid)));
  else MFB_TRY((launch_eval_group<32, 2>(e, ga, variant, grid)));
  double *d_out = d_partial + (size_t)grid * 32;
  MFB_LAUNCH(eval_group_final_kernel, 1, 256, 0, e->stream, d_partial, grid, d_out);
  MFB_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * 32, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  dev_free(d_groups);
  dev_free(d_partial);
  return 0;
}

}  // namespace mfb
