// ModelMF::trainCCD (modelMF.cpp:1426-1653, --mf_method ccd): cyclic coordinate descent one ROW at a time over the
// residual matrix that mfb_ccdpp_begin sets up (residual = ratings in both views, U = 0).
//
// A row's dims are visited one after the other (every update of u_k changes the residual the next dim reads), rows of
// one side are independent: one warp per row, the row's residuals and item ids in registers while it has at most
// 32 x kCcdRegs ratings, streamed from memory otherwise; rows longer than kCcdLong go to a second launch that gives
// every such row a whole CTA.  Arithmetic follows the reference line by line: float products (no contraction) summed
// in double (:1541-1550), newV / upd in double, `res -= upd` rounded once to float (:1553-1556).  The patch of the
// other view (:1557-1563) happens once per rating after the last dim: both views receive the same sequence of
// subtractions from the same start value in the reference, so the final value is what the other view must hold; its
// position is found with the reference's own binary search (util.cpp:847-864), absent entries are skipped as there.
#include <algorithm>

#include "engine.h"

namespace mfb {
namespace {

constexpr int kCcdRegs = 8;
constexpr int kCcdLong = 2048;
constexpr int kCcdThreads = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// util.cpp:847-864
__device__ __forceinline__ int64_t bin_search(const int32_t *__restrict__ sorted, int key, int64_t ub, int64_t lb) {
  while (ub >= lb) {
    const int64_t mid = (ub + lb) / 2;
    const int v = sorted[mid];
    if (v == key) return mid;
    if (v < key) lb = mid + 1;
    else ub = mid - 1;
  }
  return -1;
}

struct CcdArgs {
  const int64_t *ptr;     // this view: row pointers
  const int32_t *ind;
  float *res;
  const int64_t *optr;    // the other view
  const int32_t *oind;
  float *ores;
  float *fself;           // factor matrix of the rows being updated
  const float *fother;
  const uint8_t *mask;    // 1 = invalid row (may be null)
  const uint8_t *dims;    // [n][rank] visiting order of every row, null = 0..rank-1
  int32_t n, rank, ld;
  float reg;
  int32_t *counters;      // [0] next row, [1] long rows found, [2] next long row
  int32_t *long_rows;
};

// one warp, row of at most 32 * E ratings kept in registers
template <int E>
__device__ __forceinline__ void ccd_row_regs(const CcdArgs &a, int row, int64_t start, int len, int lane) {
  int it[E];
  float rs[E];
#pragma unroll
  for (int e = 0; e < E; e++) {
    const int idx = lane + 32 * e;
    const bool ok = idx < len;
    it[e] = ok ? a.ind[start + idx] : -1;
    rs[e] = ok ? a.res[start + idx] : 0.0f;
  }
  float *fs = a.fself + (size_t)row * a.ld;
  const uint8_t *dm = a.dims ? a.dims + (size_t)row * a.rank : nullptr;
  for (int kk = 0; kk < a.rank; kk++) {
    const int k = dm ? dm[kk] : kk;
    const float uk = fs[k];
    float f[E];
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int e = 0; e < E; e++) {
      f[e] = 0.0f;
      if (it[e] >= 0) {
        f[e] = a.fother[(size_t)it[e] * a.ld + k];
        num += (double)__fmul_rn(__fadd_rn(rs[e], __fmul_rn(uk, f[e])), f[e]);
        den += (double)__fmul_rn(f[e], f[e]);
      }
    }
    num = warp_sum(num);
    den = (double)a.reg + warp_sum(den);
    const double newV = num / den;
    const double dlt = newV - (double)uk;
#pragma unroll
    for (int e = 0; e < E; e++)
      if (it[e] >= 0) rs[e] = (float)((double)rs[e] - dlt * (double)f[e]);
    __syncwarp();
    if (lane == 0) fs[k] = (float)newV;
  }
#pragma unroll
  for (int e = 0; e < E; e++) {
    if (it[e] < 0) continue;
    a.res[start + lane + 32 * e] = rs[e];
    const int64_t pos = bin_search(a.oind, row, a.optr[it[e] + 1] - 1, a.optr[it[e]]);
    if (pos >= 0) a.ores[pos] = rs[e];
  }
}

// TEAM threads (one warp or the whole CTA), residuals streamed from memory; every thread owns the entries tid, tid + TEAM, ...
template <int TEAM>
__device__ __forceinline__ void ccd_row_mem(const CcdArgs &a, int row, int64_t start, int len, int tid, double *red) {
  float *fs = a.fself + (size_t)row * a.ld;
  const uint8_t *dm = a.dims ? a.dims + (size_t)row * a.rank : nullptr;
  const int32_t *ind = a.ind + start;
  float *res = a.res + start;
  for (int kk = 0; kk < a.rank; kk++) {
    const int k = dm ? dm[kk] : kk;
    const float uk = fs[k];
    double num = 0.0, den = 0.0;
    for (int idx = tid; idx < len; idx += TEAM) {
      const float f = a.fother[(size_t)ind[idx] * a.ld + k];
      num += (double)__fmul_rn(__fadd_rn(res[idx], __fmul_rn(uk, f)), f);
      den += (double)__fmul_rn(f, f);
    }
    num = warp_sum(num);
    den = warp_sum(den);
    if (TEAM > 32) {  // every thread adds the warp sums in the same order; two buffers, one barrier per dim
      double *buf = red + (kk & 1) * 2 * (TEAM / 32);
      if ((tid & 31) == 0) { buf[2 * (tid >> 5)] = num; buf[2 * (tid >> 5) + 1] = den; }
      __syncthreads();
      num = den = 0.0;
#pragma unroll
      for (int w = 0; w < TEAM / 32; w++) { num += buf[2 * w]; den += buf[2 * w + 1]; }
    }
    den += (double)a.reg;
    const double newV = num / den;
    const double dlt = newV - (double)uk;
    for (int idx = tid; idx < len; idx += TEAM) {
      const float f = a.fother[(size_t)ind[idx] * a.ld + k];
      res[idx] = (float)((double)res[idx] - dlt * (double)f);
    }
    if (tid == 0) fs[k] = (float)newV;
  }
  for (int idx = tid; idx < len; idx += TEAM) {
    const int it = ind[idx];
    const int64_t pos = bin_search(a.oind, row, a.optr[it + 1] - 1, a.optr[it]);
    if (pos >= 0) a.ores[pos] = res[idx];
  }
}

__global__ void __launch_bounds__(kCcdThreads) ccd_warp_rows_kernel(CcdArgs a) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int row = 0;
    if (lane == 0) row = atomicAdd(&a.counters[0], 1);
    row = __shfl_sync(0xffffffffu, row, 0);
    if (row >= a.n) break;
    if (a.mask && a.mask[row]) continue;
    const int64_t start = a.ptr[row];
    const int64_t len64 = a.ptr[row + 1] - start;
    if (len64 > kCcdLong) {
      if (lane == 0) a.long_rows[atomicAdd(&a.counters[1], 1)] = row;
      continue;
    }
    const int len = (int)len64;
    if (len <= 32) ccd_row_regs<1>(a, row, start, len, lane);
    else if (len <= 64) ccd_row_regs<2>(a, row, start, len, lane);
    else if (len <= 128) ccd_row_regs<4>(a, row, start, len, lane);
    else if (len <= 32 * kCcdRegs) ccd_row_regs<kCcdRegs>(a, row, start, len, lane);
    else ccd_row_mem<32>(a, row, start, len, lane, nullptr);
  }
}

__global__ void __launch_bounds__(kCcdThreads) ccd_long_rows_kernel(CcdArgs a) {
  __shared__ double red[2 * 2 * (kCcdThreads / 32)];
  __shared__ int s_next;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_next = atomicAdd(&a.counters[2], 1);
    __syncthreads();
    const int i = s_next;
    if (i >= a.counters[1]) break;
    const int row = a.long_rows[i];
    const int64_t start = a.ptr[row];
    ccd_row_mem<kCcdThreads>(a, row, start, (int)(a.ptr[row + 1] - start), threadIdx.x, red);
  }
}

}  // namespace

// One half of a trainCCD epoch: side = MFB_USER walks the CSR view and updates U (:1527-1566), MFB_ITEM the CSC view and V
// (:1569-1606).  dims_host = [n rows of that side][rank] bytes, the order in which every row visits its dims (the
// caller draws them as the reference does, std::shuffle from one mt19937(trainSeed)); null = 0 .. rank-1 for every row.
int ccd_half_step_impl(mfb_engine *e, int side, float reg, const uint8_t *dims_host) {
  DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  const bool user = side == MFB_USER;
  const int32_t n = user ? e->n_users : e->n_items;
  CcdArgs a;
  a.ptr = user ? m.rowptr : m.colptr;
  a.ind = user ? m.rowind : m.colind;
  a.res = user ? e->res_row : e->res_col;
  a.optr = user ? m.colptr : m.rowptr;
  a.oind = user ? m.colind : m.rowind;
  a.ores = user ? e->res_col : e->res_row;
  a.fself = user ? e->U : e->V;
  a.fother = user ? e->V : e->U;
  a.mask = user ? e->bad_user : e->bad_item;
  a.n = n;
  a.rank = e->rank;
  a.ld = e->ld;
  a.reg = reg;
  uint8_t *dims = nullptr;
  if (dims_host) {
    MFB_CUDA(dev_alloc(&dims, (size_t)n * e->rank));
    MFB_CUDA(cudaMemcpyAsync(dims, dims_host, (size_t)n * e->rank, cudaMemcpyHostToDevice, st));
  }
  a.dims = dims;
  int32_t *work = nullptr;
  MFB_CUDA(dev_alloc(&work, sizeof(int32_t) * ((size_t)n + 4)));
  MFB_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t) * 4, st));
  a.counters = work;
  a.long_rows = work + 4;
  const int grid = e->sm_count * 4;
  MFB_LAUNCH(ccd_warp_rows_kernel, grid, kCcdThreads, 0, st, a);
  MFB_LAUNCH(ccd_long_rows_kernel, grid, kCcdThreads, 0, st, a);
  if (dims_host) MFB_CUDA(cudaStreamSynchronize(st));  // the caller may overwrite its order buffer once this returns
  if (dims) dev_free(dims);
  dev_free(work);
  return 0;
}

}  // namespace mfb
