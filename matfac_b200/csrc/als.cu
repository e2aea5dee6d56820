// ALS half-step: per-row Gram + right-hand side, then a Cholesky solve, fused in one CTA.
//
// Replaces the OpenMP row loops of ModelMF::trainALS (modelMF.cpp:806-841 users, :845-880
// items): A = sum_{rated, r>0} f f^T + reg I, b = sum r f, x = A^-1 b, where f are the
// opposite side's factor rows.  The reference factorises with Eigen's fp32 pivoted LDL^T
// (:836); A + reg I is symmetric positive definite so an unpivoted fp32 LL^T gives the same
// solution to rounding.
//
// This translation unit is the CUDA-core (fp32 FMA) form of the Gram: one 256-thread CTA per
// row segment, a 16 x 16 thread grid in which every thread keeps a TR x TR register tile of
// A, factor rows staged through shared memory by cp.async (double buffered).  The Cholesky
// runs on the same register tiles (one shared-memory column broadcast per step), the two
// triangular solves by one warp.  Rows longer than a chunk are split over CTAs that add
// their partial Gram into a workspace; a second kernel solves those.
#include "engine.h"

#include <cuda_pipeline.h>

namespace mfb {

constexpr int kAlsTile = 32;    // factor rows staged per pipeline stage
constexpr int kAlsChunk = 2048; // ratings per CTA before a row is split

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
};

template <int TR>
__device__ __forceinline__ int tile_idx(int t, int i) {
  // element owned by thread coordinate t, local index i; for TR = 8 the tile is two strided
  // float4 groups so that a half-warp's 128-bit shared loads are contiguous (no bank conflict)
  if (TR <= 4) return t * TR + i;
  return (i >> 2) * 64 + t * 4 + (i & 3);
}

template <int TR>
struct AlsSmem {
  static constexpr int RP = 16 * TR;
  static constexpr int LDL = RP + 1;
  // layout (floats): tile[2][kAlsTile][RP] | L[RP][LDL] | col[2][RP] | b[RP] | rate[3][kAlsTile] ; then int item[3][kAlsTile]
  static constexpr int off_tile = 0;
  static constexpr int off_L = off_tile + 2 * kAlsTile * RP;
  static constexpr int off_col = off_L + RP * LDL;
  static constexpr int off_b = off_col + 2 * RP;
  static constexpr int off_rate = off_b + RP;
  static constexpr int off_item = off_rate + 3 * kAlsTile;
  static constexpr int total_floats = off_item + 3 * kAlsTile;
  static constexpr size_t bytes = sizeof(float) * total_floats;
};

// Cholesky of the register-tiled matrix (lower triangle meaningful), L written to shared
// memory, then L L^T x = b solved by warp 0; x left in sm_b.
template <int TR>
__device__ __forceinline__ void chol_solve(float (&acc)[TR][TR], float *sm, int tx, int ty) {
  using S = AlsSmem<TR>;
  constexpr int RP = S::RP, LDL = S::LDL;
  float *Lm = sm + S::off_L, *colbuf = sm + S::off_col, *bv = sm + S::off_b;
  const int tid = ty * 16 + tx;
  constexpr int NH = TR <= 4 ? 1 : TR / 4;  // strided groups
  constexpr int IW = TR <= 4 ? TR : 4;      // local columns per group
  int step = 0;
#pragma unroll
  for (int h = 0; h < NH; h++) {
    for (int kb = 0; kb < 16; kb++) {
#pragma unroll
      for (int i4 = 0; i4 < IW; i4++, step++) {
        const int ic = h * 4 + i4;  // static local column (for TR <= 4: h = 0, ic = i4)
        const int k = tile_idx<TR>(kb, ic);
        float *cb = colbuf + (step & 1) * RP;
        if (tx == kb) {
#pragma unroll
          for (int i = 0; i < TR; i++) cb[tile_idx<TR>(ty, i)] = acc[i][ic];
        }
        __syncthreads();
        const float d = cb[k];
        const float inv = d > 0.f ? 1.0f / d : 0.f;
        if (tid < RP && tid >= k) Lm[tid * LDL + k] = d > 0.f ? cb[tid] / sqrtf(d) : 0.f;
        float cj[TR];
#pragma unroll
        for (int j = 0; j < TR; j++) cj[j] = cb[tile_idx<TR>(tx, j)];
#pragma unroll
        for (int i = 0; i < TR; i++) {
          const float s = cb[tile_idx<TR>(ty, i)] * inv;
#pragma unroll
          for (int j = 0; j < TR; j++) acc[i][j] = fmaf(-s, cj[j], acc[i][j]);
        }
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    constexpr int PER = (RP + 31) / 32;
    // forward: L y = b (column oriented)
    float bl[PER];
#pragma unroll
    for (int q = 0; q < PER; q++) bl[q] = (q * 32 + lane) < RP ? bv[q * 32 + lane] : 0.f;
    for (int k = 0; k < RP; k++) {
      const float diag = Lm[k * LDL + k];
      float yk = 0.f;
#pragma unroll
      for (int q = 0; q < PER; q++)
        if ((k >> 5) == q) yk = bl[q];
      yk = __shfl_sync(0xFFFFFFFFu, yk, k & 31);
      yk = diag != 0.f ? yk / diag : 0.f;
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int i = q * 32 + lane;
        if (i == k) bl[q] = yk;
        else if (i > k && i < RP) bl[q] = fmaf(-Lm[i * LDL + k], yk, bl[q]);
      }
    }
    // backward: L^T x = y (row oriented)
    for (int k = RP - 1; k >= 0; k--) {
      const float diag = Lm[k * LDL + k];
      float xk = 0.f;
#pragma unroll
      for (int q = 0; q < PER; q++)
        if ((k >> 5) == q) xk = bl[q];
      xk = __shfl_sync(0xFFFFFFFFu, xk, k & 31);
      xk = diag != 0.f ? xk / diag : 0.f;
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int i = q * 32 + lane;
        if (i == k) bl[q] = xk;
        else if (i < k) bl[q] = fmaf(-Lm[k * LDL + i], xk, bl[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < PER; q++)
      if ((q * 32 + lane) < RP) bv[q * 32 + lane] = bl[q];
  }
  __syncthreads();
}

template <int TR>
__device__ __forceinline__ void add_reg_diag(float (&acc)[TR][TR], int tx, int ty, int rank, float reg) {
#pragma unroll
  for (int i = 0; i < TR; i++)
#pragma unroll
    for (int j = 0; j < TR; j++) {
      const int r = tile_idx<TR>(ty, i), c = tile_idx<TR>(tx, j);
      if (r == c) acc[i][j] = r < rank ? acc[i][j] + reg : 1.0f;  // padded dims: identity
    }
}

template <int TR>
__global__ void __launch_bounds__(256, (TR == 8 ? 2 : 3)) als_gram_solve_kernel(const AlsArgs a) {
  using S = AlsSmem<TR>;
  constexpr int RP = S::RP;
  extern __shared__ __align__(16) float sm[];
  float *tile = sm + S::off_tile, *bv = sm + S::off_b, *rate = sm + S::off_rate;
  int *item = reinterpret_cast<int *>(sm + S::off_item);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int seg = blockIdx.x;
  const int row = a.seg_row[seg], start = a.seg_start[seg], len = a.seg_len[seg], slot = a.seg_slot[seg];
  const int nq = a.ld >> 2;

  // zero both stages once: columns >= ld are never written afterwards
  for (int i = tid; i < 2 * kAlsTile * RP; i += 256) tile[i] = 0.f;
  float acc[TR][TR], bacc[TR];
#pragma unroll
  for (int i = 0; i < TR; i++) {
    bacc[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TR; j++) acc[i][j] = 0.f;
  }
  const int ntiles = (len + kAlsTile - 1) / kAlsTile;
  auto load_idx = [&](int t) {
    if (tid < kAlsTile) {
      const int j = t * kAlsTile + tid;
      int it = 0;
      float rt = 0.f;
      if (j < len) {
        it = __ldg(a.ind + start + j);
        rt = __ldg(a.val + start + j);
      }
      item[(t % 3) * kAlsTile + tid] = it;
      rate[(t % 3) * kAlsTile + tid] = (j < len && rt > 0.f) ? rt : 0.f;  // rating > 0 filter (modelMF.cpp:819)
    }
  };
  auto issue_rows = [&](int t) {
    float *dst = tile + (t & 1) * kAlsTile * RP;
    const int *its = item + (t % 3) * kAlsTile;
    for (int i = tid; i < kAlsTile * nq; i += 256) {
      const int r = i / nq, q = i - r * nq;
      __pipeline_memcpy_async(dst + r * RP + q * 4, a.Fin + (size_t)its[r] * a.ld + q * 4, 16);
    }
    __pipeline_commit();
  };
  load_idx(0);
  __syncthreads();
  issue_rows(0);
  if (ntiles > 1) load_idx(1);
  for (int t = 0; t < ntiles; t++) {
    __pipeline_wait_prior(0);
    __syncthreads();
    if (t + 1 < ntiles) issue_rows(t + 1);
    if (t + 2 < ntiles) load_idx(t + 2);
    const float *tb = tile + (t & 1) * kAlsTile * RP;
    const float *rt = rate + (t % 3) * kAlsTile;
    const int rows = min(kAlsTile, len - t * kAlsTile);
    for (int j = 0; j < rows; j++) {
      const float r = rt[j];
      const float flag = r > 0.f ? 1.f : 0.f;
      const float *fr = tb + j * RP;
      float av[TR], bw[TR];
      if (TR == 8) {
        const float4 a0 = *reinterpret_cast<const float4 *>(fr + ty * 4), a1 = *reinterpret_cast<const float4 *>(fr + 64 + ty * 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(fr + tx * 4), b1 = *reinterpret_cast<const float4 *>(fr + 64 + tx * 4);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4 % TR] = a1.x; av[5 % TR] = a1.y; av[6 % TR] = a1.z; av[7 % TR] = a1.w;
        bw[0] = b0.x; bw[1] = b0.y; bw[2] = b0.z; bw[3] = b0.w; bw[4 % TR] = b1.x; bw[5 % TR] = b1.y; bw[6 % TR] = b1.z; bw[7 % TR] = b1.w;
      } else {
#pragma unroll
        for (int i = 0; i < TR; i++) {
          av[i] = fr[ty * TR + i];
          bw[i] = fr[tx * TR + i];
        }
      }
#pragma unroll
      for (int i = 0; i < TR; i++) {
        const float ai = av[i] * flag;
#pragma unroll
        for (int jj = 0; jj < TR; jj++) acc[i][jj] = fmaf(ai, bw[jj], acc[i][jj]);
      }
      if (ty == 0) {
#pragma unroll
        for (int i = 0; i < TR; i++) bacc[i] = fmaf(r, bw[i], bacc[i]);
      }
    }
  }
  if (slot >= 0) {
    // partial Gram of a split row
    float *w = a.ws + (size_t)slot * (RP * RP + RP);
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TR; j++) atomicAdd(w + tile_idx<TR>(ty, i) * RP + tile_idx<TR>(tx, j), acc[i][j]);
    if (ty == 0) {
#pragma unroll
      for (int i = 0; i < TR; i++) atomicAdd(w + RP * RP + tile_idx<TR>(tx, i), bacc[i]);
    }
    return;
  }
  if (ty == 0) {
#pragma unroll
    for (int i = 0; i < TR; i++) bv[tile_idx<TR>(tx, i)] = bacc[i];
  }
  add_reg_diag<TR>(acc, tx, ty, a.rank, a.reg);
  chol_solve<TR>(acc, sm, tx, ty);
  if (tid < a.ld) a.Fout[(size_t)row * a.ld + tid] = tid < a.rank ? bv[tid] : 0.f;
}

template <int TR>
__global__ void __launch_bounds__(256, (TR == 8 ? 2 : 3)) als_solve_ws_kernel(const AlsArgs a) {
  using S = AlsSmem<TR>;
  constexpr int RP = S::RP;
  extern __shared__ __align__(16) float sm[];
  float *bv = sm + S::off_b;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int slot = blockIdx.x;
  const int row = a.multi_row[slot];
  const float *w = a.ws + (size_t)slot * (RP * RP + RP);
  float acc[TR][TR];
#pragma unroll
  for (int i = 0; i < TR; i++)
#pragma unroll
    for (int j = 0; j < TR; j++) acc[i][j] = w[tile_idx<TR>(ty, i) * RP + tile_idx<TR>(tx, j)];
  if (tid < RP) bv[tid] = w[RP * RP + tid];
  add_reg_diag<TR>(acc, tx, ty, a.rank, a.reg);
  chol_solve<TR>(acc, sm, tx, ty);
  if (tid < a.ld) a.Fout[(size_t)row * a.ld + tid] = tid < a.rank ? bv[tid] : 0.f;
}

template <int TR>
static int launch_als(mfb_engine *e, const AlsArgs &a, const SegPlan &sp) {
  using S = AlsSmem<TR>;
  MFB_CUDA(cudaFuncSetAttribute(als_gram_solve_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
  MFB_CUDA(cudaFuncSetAttribute(als_solve_ws_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
  if (sp.n_multi > 0) {
    size_t need = sizeof(float) * (size_t)sp.n_multi * (S::RP * S::RP + S::RP);
    if (need > e->als_ws_bytes) {
      if (e->als_ws) MFB_CUDA(cudaFree(e->als_ws));
      e->als_ws = nullptr;
      e->als_ws_bytes = 0;
      MFB_CUDA(cudaMalloc(&e->als_ws, need));
      e->als_ws_bytes = need;
    }
    MFB_CUDA(cudaMemsetAsync(e->als_ws, 0, need, e->stream));
  }
  AlsArgs b = a;
  b.ws = e->als_ws;
  if (sp.n_seg > 0) MFB_LAUNCH((als_gram_solve_kernel<TR>), sp.n_seg, 256, S::bytes, e->stream, b);
  if (sp.n_multi > 0) MFB_LAUNCH((als_solve_ws_kernel<TR>), sp.n_multi, 256, S::bytes, e->stream, b);
  return 0;
}

int als_half_step_launch(mfb_engine *e, int side, float reg) {
  DevCsr &m = e->mat[MFB_TRAIN];
  SegPlan &sp = side == MFB_USER ? m.als_rows : m.als_cols;
  if (!sp.built) {
    if (side == MFB_USER)
      MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, e->row_begin[MFB_USER], e->row_end[MFB_USER],
                             kAlsChunk, &sp));
    else
      MFB_TRY(build_seg_plan(e, m.colptr, e->n_items, e->bad_item, e->row_begin[MFB_ITEM], e->row_end[MFB_ITEM],
                             kAlsChunk, &sp));
  }
  AlsArgs a;
  a.Fin = side == MFB_USER ? e->V : e->U;
  a.Fout = side == MFB_USER ? e->U : e->V;
  a.ld = e->ld;
  a.rank = e->rank;
  a.ind = side == MFB_USER ? m.rowind : m.colind;
  a.val = side == MFB_USER ? m.rowval : m.colval;
  a.seg_row = sp.row; a.seg_start = sp.start; a.seg_len = sp.len; a.seg_slot = sp.slot;
  a.multi_row = sp.multi_row;
  a.ws = nullptr;
  a.reg = reg;
  const int r = e->rank;
  if (r <= 16) return launch_als<1>(e, a, sp);
  if (r <= 32) return launch_als<2>(e, a, sp);
  if (r <= 64) return launch_als<4>(e, a, sp);
  return launch_als<8>(e, a, sp);
}

}  // namespace mfb
