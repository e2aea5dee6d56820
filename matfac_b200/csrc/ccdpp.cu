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

#include <cub/cub.cuh>

#include <algorithm>

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

// old_off > 0: a second copy at out[old_off + i] on every rank — u_k as it was before the step's updates.  A copy made
// by each rank after the barrier would race with a faster peer's first update pass storing new u_k entries.
__global__ void col_extract_kernel(const float *__restrict__ F, int ld, int k, int lo, int hi, float *__restrict__ out,
                                   const CcdPeers pe, size_t old_off) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) {
    const float v = F[(size_t)i * ld + k];
    store_all(out, pe, i, v);
    if (old_off) store_all(out, pe, (int)old_off + i, v);
  }
}
// writes column k of this rank's rows back into the factor matrix — its own and every peer's, so that all ranks hold
// the complete factors after every rank-one step (the per-epoch evaluation and the best-model snapshot read them)
struct CcdPeerMats {
  float *p[kMaxRanks - 1];
  int n;
};
__global__ void col_insert_kernel(float *__restrict__ F, int ld, int k, int lo, int hi, const float *__restrict__ in,
                                  const CcdPeerMats pm) {
  const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < hi) {
    const float v = in[i];
    F[(size_t)i * ld + k] = v;
    for (int q = 0; q < pm.n; q++) pm.p[q][(size_t)i * ld + k] = v;
  }
}


// ---- warp-streamed passes (option ccd_stream = 2; measured equal to the warp-per-segment kernels, profiles/r2_ccdpp.md) -----
// One warp per chunk of consecutive segments.  The warp streams the chunk's contiguous range of the index / residual
// arrays in batches of 256 ratings (8 per lane, coalesced) and ALWAYS has the next batch in flight while it works on the
// current one, whatever the rows look like; the segments are then only boundaries inside the stream: per batch the warp
// walks the segments that overlap it, adds the ratings that fall into [start, start + len) to the per-lane fp64 partial
// sums and closes a row when its last rating has passed.  Segment descriptors are fetched 32 at a time (one per lane)
// and handed round by shuffle.  Streaming loads carry the evict-first hint so that L1 keeps the gathered vector.
constexpr int kCcdBatch = 256;

struct CcdDesc {
  int row, start, len, slot;
  float a;
};
// lane l holds the descriptor of segment sb + l (a = own[row] when WITH_A)
template <bool WITH_A>
__device__ __forceinline__ CcdDesc ccd_load_descs(const CcdPass &p, int sb, int s1, int lane, const float *own) {
  CcdDesc d{0, 0, 0, -1, 0.f};
  const int s = sb + lane;
  if (s < s1) {
    d.row = __ldg(p.seg_row + s); d.start = __ldg(p.seg_start + s); d.len = __ldg(p.seg_len + s); d.slot = __ldg(p.seg_slot + s);
    if (WITH_A) d.a = own[d.row];
  }
  return d;
}
__device__ __forceinline__ CcdDesc ccd_pick_desc(const CcdDesc &d, int src) {
  CcdDesc o;
  o.row = __shfl_sync(0xFFFFFFFFu, d.row, src); o.start = __shfl_sync(0xFFFFFFFFu, d.start, src);
  o.len = __shfl_sync(0xFFFFFFFFu, d.len, src); o.slot = __shfl_sync(0xFFFFFFFFu, d.slot, src);
  o.a = __shfl_sync(0xFFFFFFFFu, d.a, src);
  return o;
}

// STAGED (option ccd_stage): persistent CTAs (one per SM) keep the whole gathered vector in shared memory and their warps
// take chunks in turn; the add-back of a row pass reads the same vector (v_k has not moved yet).  Measured on the Netflix
// shape (profiles/r2_ccdpp.md): the gather leaves L2 (152 M -> 54 M sectors per pass, L2 at 19 %) and the pass becomes
// issue-bound instead — 194 M warp instructions for 503 k rows, 54 % of the issue slots at 24 warps per SM — 374 us against
// ~310 us with the gather through L1: off by default.
constexpr int kCcdStagedThreads = 768;
__device__ __forceinline__ void ccd_stage(uint64_t *bar_mem, float *dst0, const float *src0, float *dst1, const float *src1, int n);

template <bool ADDBACK, bool SUBTRACT, bool STAGED>
__global__ void __launch_bounds__(STAGED ? kCcdStagedThreads : 256, STAGED ? 1 : 3)
    ccd_update_flat_kernel(const CcdPass p, const int32_t *__restrict__ chunk_seg, int n_chunk, float *__restrict__ own,
                           const float *__restrict__ other, const float *__restrict__ other_old, float reg, double *__restrict__ acc,
                           const Aux *__restrict__ aux_freq, int freq_thresh, const CcdPeers pe, int gather_n) {
  constexpr int D = kCcdBatch / 32;
  extern __shared__ __align__(16) float ccd_sm[];
  __shared__ uint64_t stage_bar;
  if (STAGED) ccd_stage(&stage_bar, ccd_sm, other, nullptr, nullptr, gather_n);
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_chunk; w += n_warps) {
  const int s0 = chunk_seg[w], s1 = chunk_seg[w + 1];
  int sb = s0;
  CcdDesc dl = ccd_load_descs<ADDBACK>(p, sb, s1, lane, own);
  CcdDesc cur = ccd_pick_desc(dl, 0);
  const int g0 = cur.start;
  const int g1 = __ldg(p.seg_start + s1 - 1) + __ldg(p.seg_len + s1 - 1);
  int seg = s0;
  int c[D], cn[D];
  float r[D], rn[D], o[D], q[D];
#pragma unroll
  for (int d = 0; d < D; d++) {
    const int pos = g0 + d * 32 + lane;
    c[d] = pos < g1 ? __ldcs(p.ind + pos) : -1;
    r[d] = pos < g1 ? __ldcs(p.res + pos) : 0.f;
  }
  double num = 0.0, den = 0.0;
  for (int base = g0; base < g1; base += kCcdBatch) {
#pragma unroll
    for (int d = 0; d < D; d++) {
      if (STAGED) {
        o[d] = c[d] >= 0 ? ccd_sm[c[d]] : 0.f;
        if (ADDBACK) q[d] = o[d];
      } else {
        o[d] = c[d] >= 0 ? __ldg(other + c[d]) : 0.f;
        if (ADDBACK) q[d] = c[d] >= 0 ? __ldg(other_old + c[d]) : 0.f;
      }
    }
#pragma unroll
    for (int d = 0; d < D; d++) {  // the next batch travels while this one is reduced
      const int pos = base + kCcdBatch + d * 32 + lane;
      cn[d] = pos < g1 ? __ldcs(p.ind + pos) : -1;
      rn[d] = pos < g1 ? __ldcs(p.res + pos) : 0.f;
    }
    const int bend = min(base + kCcdBatch, g1);
    while (seg < s1 && cur.start < bend) {
      const int send = cur.start + cur.len;
#pragma unroll
      for (int d = 0; d < D; d++) {
        const int pos = base + d * 32 + lane;
        if (pos >= cur.start && pos < send) {
          float rr = r[d];
          if (ADDBACK) {
            rr = __fadd_rn(rr, __fmul_rn(cur.a, q[d]));
            p.res[pos] = rr;
          }
          num += (double)__fmul_rn(rr, o[d]);
          den += (double)__fmul_rn(o[d], o[d]);
        }
      }
      if (send > bend) break;  // the row continues in the next batch
#pragma unroll
      for (int m = 16; m >= 1; m >>= 1) {
        num += __shfl_xor_sync(0xFFFFFFFFu, num, m);
        den += __shfl_xor_sync(0xFFFFFFFFu, den, m);
      }
      if (cur.slot < 0) {
        float nv = (float)(num / ((double)reg + den));
        if (freq_thresh > 0 && aux_freq[cur.row].freq < freq_thresh) nv = 0.f;
        if (lane == 0) store_all(own, pe, cur.row, nv);
        if (SUBTRACT) {  // the row's ratings have just passed through L1 / L2
          const float a = -nv;
          for (int j = cur.start + lane; j < send; j += 32)
            p.res[j] = __fadd_rn(p.res[j], __fmul_rn(a, __ldg(other + __ldg(p.ind + j))));
        }
      } else if (lane == 0) {
        atomicAdd(acc + 2 * (size_t)cur.slot, num);
        atomicAdd(acc + 2 * (size_t)cur.slot + 1, den);
      }
      num = 0.0; den = 0.0;
      seg++;
      if (seg - sb == 32) {
        sb = seg;
        dl = ccd_load_descs<ADDBACK>(p, sb, s1, lane, own);
      }
      cur = ccd_pick_desc(dl, seg - sb);
    }
#pragma unroll
    for (int d = 0; d < D; d++) { c[d] = cn[d]; r[d] = rn[d]; }
  }
  }
}

// res[j] += sign * own[row] * other[ind[j]], the same stream
template <bool STAGED>
__global__ void __launch_bounds__(STAGED ? kCcdStagedThreads : 256, STAGED ? 1 : 4)
    ccd_resid_flat_kernel(const CcdPass p, const int32_t *__restrict__ chunk_seg, int n_chunk, const float *__restrict__ own,
                          const float *__restrict__ other, float sign, int gather_n) {
  constexpr int D = kCcdBatch / 32;
  extern __shared__ __align__(16) float ccd_sm[];
  __shared__ uint64_t stage_bar;
  if (STAGED) ccd_stage(&stage_bar, ccd_sm, other, nullptr, nullptr, gather_n);
  const int lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_chunk; w += n_warps) {
  const int s0 = chunk_seg[w], s1 = chunk_seg[w + 1];
  int sb = s0;
  CcdDesc dl = ccd_load_descs<true>(p, sb, s1, lane, own);
  CcdDesc cur = ccd_pick_desc(dl, 0);
  const int g0 = cur.start;
  const int g1 = __ldg(p.seg_start + s1 - 1) + __ldg(p.seg_len + s1 - 1);
  int seg = s0;
  int c[D], cn[D];
  float r[D], rn[D], o[D];
#pragma unroll
  for (int d = 0; d < D; d++) {
    const int pos = g0 + d * 32 + lane;
    c[d] = pos < g1 ? __ldcs(p.ind + pos) : -1;
    r[d] = pos < g1 ? __ldcs(p.res + pos) : 0.f;
  }
  for (int base = g0; base < g1; base += kCcdBatch) {
#pragma unroll
    for (int d = 0; d < D; d++) o[d] = c[d] >= 0 ? (STAGED ? ccd_sm[c[d]] : __ldg(other + c[d])) : 0.f;
#pragma unroll
    for (int d = 0; d < D; d++) {
      const int pos = base + kCcdBatch + d * 32 + lane;
      cn[d] = pos < g1 ? __ldcs(p.ind + pos) : -1;
      rn[d] = pos < g1 ? __ldcs(p.res + pos) : 0.f;
    }
    const int bend = min(base + kCcdBatch, g1);
    while (seg < s1 && cur.start < bend) {
      const int send = cur.start + cur.len;
      const float a = sign * cur.a;
#pragma unroll
      for (int d = 0; d < D; d++) {
        const int pos = base + d * 32 + lane;
        if (pos >= cur.start && pos < send) __stcs(p.res + pos, __fadd_rn(r[d], __fmul_rn(a, o[d])));
      }
      if (send > bend) break;
      seg++;
      if (seg - sb == 32) {
        sb = seg;
        dl = ccd_load_descs<true>(p, sb, s1, lane, own);
      }
      cur = ccd_pick_desc(dl, seg - sb);
    }
#pragma unroll
    for (int d = 0; d < D; d++) { c[d] = cn[d]; r[d] = rn[d]; }
  }
  }
}

constexpr size_t kCcdStagedMaxBytes = 200 * 1024;  // of the 227 KB a CTA may have
static size_t ccd_flat_smem(int gather_n) { return sizeof(float) * (size_t)((gather_n + 3) & ~3); }
static bool ccd_flat_staged(const mfb_engine *e, int gather_n) {
  return e->opt_ccd_stage && gather_n > 0 && ccd_flat_smem(gather_n) <= kCcdStagedMaxBytes;
}
constexpr int kCcdCapMax = 8192;  // ratings per chunk of the warp-streamed kernels: engine option ccd_cap (default 4096), >= kCcdChunk so that every segment fits a chunk
static int ccd_cap(const mfb_engine *e) { return std::min(kCcdCapMax, std::max(kCcdChunk, e->opt_ccd_cap & ~3)); }
// groups the (memory-ordered) segments of a plan into chunks, one per warp: a chunk's segments span at most ccd_cap ratings
static int build_chunk_plan(mfb_engine *e, SegPlan *sp) {
  dev_free(sp->chunk_seg);
  sp->chunk_seg = nullptr;
  sp->n_chunk = 0;
  const int ns = sp->n_seg;
  if (ns <= 0) return 0;
  const int cap = ccd_cap(e);
  std::vector<int32_t> start((size_t)ns), len((size_t)ns), cut;
  MFB_CUDA(cudaMemcpyAsync(start.data(), sp->start, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaMemcpyAsync(len.data(), sp->len, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  cut.reserve((size_t)ns / 8 + 2);
  cut.push_back(0);
  int32_t first = start[0];
  for (int s = 0; s < ns; s++) {
    MFB_REQUIRE(len[s] <= cap && (s == 0 || start[s] >= start[s - 1] + len[s - 1]), "ccd chunk plan: segments not in memory order");
    if (start[s] + len[s] - first > cap) {
      cut.push_back(s);
      first = start[s];
    }
  }
  cut.push_back(ns);
  sp->n_chunk = (int32_t)cut.size() - 1;
  MFB_CUDA(dev_alloc(&sp->chunk_seg, sizeof(int32_t) * cut.size()));
  MFB_CUDA(cudaMemcpyAsync(sp->chunk_seg, cut.data(), sizeof(int32_t) * cut.size(), cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));  // `cut` leaves scope
  return 0;
}


// ---- shared-memory staged passes ------------------------------------------------------------------------------
// A pass gathers one entry of the dense vector `other` per rating; every lane of a warp hits a different 128-byte line
// and a 32-byte sector moves per 4-byte entry (profiles/r1_ncu_ccdpp.txt: L2 at 69 % in the column pass).  Here the
// gathered vector sits in shared memory instead (random 4-byte reads: a few bank conflicts per warp) and the index /
// residual streams are the only global traffic.  Measured slower than the plain kernels (option ccd_smem, off by
// default; profiles/r2_ccdpp.md has the numbers of every kernel family).  A vector that does not fit is cut into blocks of kCcdBlock entries;
// every row is split at the block boundaries (its indices ascend: datastruct.cpp:18 / util.cpp:919) and a CTA serves
// the sub-segments of ONE block with that block of `other` (and of `other_old` for the fused column add-back) staged.
// Partial sums of a split row meet in the fp64 accumulators, ccd_finalize_kernel closes the row — the same arithmetic
// as the plain kernels' split rows.
constexpr int kCcdBlock = 24576;  // entries per staged block: 96 KB, two arrays = 192 KB

struct CcdBlk {
  const int32_t *off;  // [nb + 1] first segment of every block (device)
  int nb, block, gather_n, ctas_per_blk;
};

// One block of the gathered vector global -> shared through the bulk-copy engine (TMA, cp.async.bulk): contiguous, up to
// 96 KB, issued by one thread in 16 KB pieces and counted in bytes on an mbarrier — the case the engine is built for
// (a per-thread load loop spends ~20 us of latency-bound iterations on the same block).  The vectors are allocated with
// 4 floats of slack so that the copy may round the block up to 16 bytes.
__device__ __forceinline__ void ccd_stage_issue(uint32_t bar, float *dst, const float *src, int n) {
  const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
  for (uint32_t off = 0; off < bytes; off += 16384u) {
    const uint32_t piece = bytes - off < 16384u ? bytes - off : 16384u;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d0 + off),
                 "l"(reinterpret_cast<const char *>(src) + off), "r"(piece), "r"(bar)
                 : "memory");
  }
}
// all threads: initialise the barrier (thread 0), issue the copies (thread 0) and wait for them
__device__ __forceinline__ void ccd_stage(uint64_t *bar_mem, float *dst0, const float *src0, float *dst1, const float *src1, int n) {
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(bar_mem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(dst1 ? 2u : 1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ccd_stage_issue(bar, dst0, src0, n);
    if (dst1) ccd_stage_issue(bar, dst1, src1, n);
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar)
        : "memory");
  }
}

template <bool ADDBACK, bool TWO>
__global__ void __launch_bounds__(512) ccd_update_sm_kernel(const CcdPass p, float *__restrict__ own, const float *__restrict__ other,
                                                            const float *__restrict__ other_old, float reg, double *__restrict__ acc,
                                                            const Aux *__restrict__ aux_freq, int freq_thresh, const CcdPeers pe,
                                                            const CcdBlk bk) {
  extern __shared__ __align__(16) float ccd_sm[];
  const int b = blockIdx.x / bk.ctas_per_blk, ci = blockIdx.x % bk.ctas_per_blk;
  const int g_lo = b * bk.block, g_n = min(bk.block, bk.gather_n - g_lo);
  __shared__ uint64_t stage_bar;
  float *so = ccd_sm, *sold = TWO ? ccd_sm + bk.block : ccd_sm;
  ccd_stage(&stage_bar, so, other + g_lo, TWO ? sold : nullptr, other_old + g_lo, g_n);
  const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int seg_hi = bk.off[b + 1];
  for (int seg = bk.off[b] + ci * wpc + (threadIdx.x >> 5); seg < seg_hi; seg += bk.ctas_per_blk * wpc) {
    const int row = p.seg_row[seg], start = p.seg_start[seg], len = p.seg_len[seg], slot = p.seg_slot[seg];
    double num = 0.0, den = 0.0;
    const float a_old = ADDBACK ? own[row] : 0.f;
    for (int base = 0; base < len; base += 32 * kCcdDepth) {
      int c[kCcdDepth];
      float r[kCcdDepth], o[kCcdDepth];
#pragma unroll
      for (int d = 0; d < kCcdDepth; d++) {
        const int j = base + d * 32 + lane;
        c[d] = j < len ? __ldg(p.ind + start + j) - g_lo : -1;
        r[d] = j < len ? p.res[start + j] : 0.f;
      }
#pragma unroll
      for (int d = 0; d < kCcdDepth; d++) {
        o[d] = c[d] >= 0 ? so[c[d]] : 0.f;
        if (ADDBACK && c[d] >= 0) {
          r[d] = __fadd_rn(r[d], __fmul_rn(a_old, sold[c[d]]));
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
    } else if (lane == 0) {
      atomicAdd(acc + 2 * (size_t)slot, num);
      atomicAdd(acc + 2 * (size_t)slot + 1, den);
    }
  }
}

__global__ void __launch_bounds__(512) ccd_resid_sm_kernel(const CcdPass p, const float *__restrict__ own, const float *__restrict__ other,
                                                           float sign, const CcdBlk bk) {
  extern __shared__ __align__(16) float ccd_sm[];
  const int b = blockIdx.x / bk.ctas_per_blk, ci = blockIdx.x % bk.ctas_per_blk;
  const int g_lo = b * bk.block, g_n = min(bk.block, bk.gather_n - g_lo);
  __shared__ uint64_t stage_bar;
  ccd_stage(&stage_bar, ccd_sm, other + g_lo, nullptr, nullptr, g_n);
  const int lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int seg_hi = bk.off[b + 1];
  for (int seg = bk.off[b] + ci * wpc + (threadIdx.x >> 5); seg < seg_hi; seg += bk.ctas_per_blk * wpc) {
    const int row = p.seg_row[seg], start = p.seg_start[seg], len = p.seg_len[seg];
    const float a = sign * __ldg(own + row);
    for (int base = 0; base < len; base += 32 * kCcdDepth) {
      int c[kCcdDepth];
      float r[kCcdDepth];
#pragma unroll
      for (int d = 0; d < kCcdDepth; d++) {
        const int j = base + d * 32 + lane;
        c[d] = j < len ? __ldg(p.ind + start + j) - g_lo : -1;
        r[d] = j < len ? p.res[start + j] : 0.f;
      }
#pragma unroll
      for (int d = 0; d < kCcdDepth; d++)
        if (c[d] >= 0) p.res[start + base + d * 32 + lane] = __fadd_rn(r[d], __fmul_rn(a, ccd_sm[c[d]]));
    }
  }
}

// plan of the sub-segments (row x block of the gathered index range), block-major
__global__ void ccd_blk_count_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ ind, const uint8_t *__restrict__ mask,
                                     int32_t row_lo, int32_t n, int nb, int block, int chunk, int32_t *__restrict__ nch,
                                     int32_t *__restrict__ valid) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * nb) return;
  const int b = (int)(t / n), i = (int)(t - (int64_t)b * n), r = row_lo + i;
  const int64_t s = ptr[r], len = ptr[r + 1] - s;
  const bool ok = len > 0 && !(mask && mask[r]);
  if (b == 0) valid[i] = ok ? 1 : 0;
  int cnt = 0;
  if (ok) {
    auto lower = [&](int key) {
      int64_t lo = 0, hi = len;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ind[s + mid] < key) lo = mid + 1; else hi = mid;
      }
      return lo;
    };
    const int64_t a0 = lower(b * block), a1 = b + 1 < nb ? lower((b + 1) * block) : len;
    cnt = (int)((a1 - a0 + chunk - 1) / chunk);
  }
  nch[t] = cnt;
}

__global__ void ccd_blk_fill_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ ind, int32_t row_lo, int32_t n, int nb,
                                    int block, int chunk, const int32_t *__restrict__ nch, const int32_t *__restrict__ off,
                                    const int32_t *__restrict__ slot_of, int32_t *__restrict__ seg_row, int32_t *__restrict__ seg_start,
                                    int32_t *__restrict__ seg_len, int32_t *__restrict__ seg_slot, int32_t *__restrict__ multi_row) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * nb) return;
  const int b = (int)(t / n), i = (int)(t - (int64_t)b * n), r = row_lo + i;
  const int cnt = nch[t];
  if (b == 0 && (cnt > 0 || slot_of[i + 1] > slot_of[i])) multi_row[slot_of[i]] = r;
  if (cnt == 0) return;
  const int64_t s = ptr[r], len = ptr[r + 1] - s;
  auto lower = [&](int key) {
    int64_t lo = 0, hi = len;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (ind[s + mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const int64_t a0 = lower(b * block), a1 = b + 1 < nb ? lower((b + 1) * block) : len;
  const int64_t per = (a1 - a0 + cnt - 1) / cnt;
  for (int k = 0; k < cnt; k++) {
    const int64_t lo = a0 + k * per, hi = lo + per < a1 ? lo + per : a1;
    const int q = off[t] + k;
    seg_row[q] = r;
    seg_start[q] = (int32_t)(s + lo);
    seg_len[q] = (int32_t)(hi - lo);
    seg_slot[q] = slot_of[i];
  }
}

// block_off (host, [nb + 1]) and d_off (device copy) describe the segment range of every block
static int build_block_seg_plan(mfb_engine *e, const int64_t *ptr, const int32_t *ind, const uint8_t *mask, int32_t row_lo,
                                int32_t row_hi, int nb, SegPlan *out, std::vector<int32_t> *block_off, int32_t **d_off) {
  out->release();
  out->built = true;
  block_off->assign((size_t)nb + 1, 0);
  const int32_t n = row_hi - row_lo;
  cudaStream_t st = e->stream;
  if (!*d_off) MFB_CUDA(dev_alloc(d_off, sizeof(int32_t) * ((size_t)nb + 1)));
  MFB_CUDA(cudaMemsetAsync(*d_off, 0, sizeof(int32_t) * ((size_t)nb + 1), st));
  if (n <= 0) return 0;
  const size_t total = (size_t)n * nb;
  int32_t *nch, *off, *valid, *slot_of;
  MFB_CUDA(dev_alloc(&nch, sizeof(int32_t) * (2 * (total + 1) + 2 * ((size_t)n + 1))));
  off = nch + total + 1;
  valid = off + total + 1;
  slot_of = valid + n + 1;
  MFB_CUDA(cudaMemsetAsync(nch, 0, sizeof(int32_t) * (2 * (total + 1) + 2 * ((size_t)n + 1)), st));
  const unsigned grid = (unsigned)((total + 255) / 256);
  MFB_LAUNCH(ccd_blk_count_kernel, grid, 256, 0, st, ptr, ind, mask, row_lo, n, nb, kCcdBlock, kCcdChunk, nch, valid);
  size_t tmp_bytes = 0, tmp2 = 0;
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, nch, off, (int)(total + 1), st));
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, valid, slot_of, n + 1, st));
  MFB_TRY(ensure_scratch(e, std::max(tmp_bytes, tmp2)));
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(e->scratch, tmp_bytes, nch, off, (int)(total + 1), st));
  MFB_CUDA(cub::DeviceScan::ExclusiveSum(e->scratch, tmp2, valid, slot_of, n + 1, st));
  std::vector<int32_t> h_off(total + 1);
  int32_t n_valid = 0;
  MFB_CUDA(cudaMemcpyAsync(h_off.data(), off, sizeof(int32_t) * (total + 1), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaMemcpyAsync(&n_valid, slot_of + n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  for (int b = 0; b <= nb; b++) (*block_off)[b] = h_off[(size_t)b * n];
  const int32_t ns = h_off[total];
  out->n_seg = ns;
  out->n_multi = n_valid;
  out->max_len = kCcdChunk;
  MFB_CUDA(cudaMemcpyAsync(*d_off, block_off->data(), sizeof(int32_t) * ((size_t)nb + 1), cudaMemcpyHostToDevice, st));
  const size_t nsz = (size_t)std::max(ns, 1), nmz = (size_t)std::max(n_valid, 1);
  MFB_CUDA(dev_alloc(&out->row, sizeof(int32_t) * nsz));
  MFB_CUDA(dev_alloc(&out->start, sizeof(int32_t) * nsz));
  MFB_CUDA(dev_alloc(&out->len, sizeof(int32_t) * nsz));
  MFB_CUDA(dev_alloc(&out->slot, sizeof(int32_t) * nsz));
  MFB_CUDA(dev_alloc(&out->multi_row, sizeof(int32_t) * nmz));
  if (ns > 0)
    MFB_LAUNCH(ccd_blk_fill_kernel, grid, 256, 0, st, ptr, ind, row_lo, n, nb, kCcdBlock, kCcdChunk, nch, off, slot_of, out->row,
               out->start, out->len, out->slot, out->multi_row);
  MFB_CUDA(cudaStreamSynchronize(st));  // block_off (host vector) has been copied
  dev_free(nch);
  return 0;
}

// Which plan a side's passes use: 0 = plain kernels (gather through L1 / L2), 1 = the whole gathered vector staged
// (one block, the ordinary row plan), 2 = blocked plan.  Blocking pays while the sub-segments stay long enough for a
// warp (>= 64 ratings on average).
static int ccd_side_mode(const mfb_engine *e, int64_t nnz_side, int n_rows_side, int gather_n, bool row_side) {
  if (!e->opt_ccd_smem || gather_n <= 0) return 0;
  if ((e->opt_ccd_smem == 2 && !row_side) || (e->opt_ccd_smem == 3 && row_side)) return 0;
  const int nb = (gather_n + kCcdBlock - 1) / kCcdBlock;
  if (nb == 1) return 1;
  if (n_rows_side <= 0 || (double)nnz_side / ((double)n_rows_side * nb) < 64.0) return 0;
  return 2;
}

// With lazy module loading (the CUDA 12 default) the first launch of a kernel loads it, and that load waits for the
// device.  Several engines of ONE process exchange through spinning barrier kernels: a variant launched for the first
// time while a peer's barrier is already spinning (the add-back forms first run in the second sweep) would then block
// the host thread that still has to issue that peer's partner — a deadlock until the flag wait times out.  So every
// kernel of the rank-one step is loaded here, before the first barrier exists.
static int ccd_preload_kernels(const mfb_engine *e) {
  cudaFuncAttributes fa;
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_resid_kernel));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_kernel<false, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_kernel<true, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_kernel<false, true>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_sm_kernel<false, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_sm_kernel<true, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_sm_kernel<true, true>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_resid_sm_kernel));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_flat_kernel<false, false, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_flat_kernel<true, false, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_update_flat_kernel<false, true, false>));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_resid_flat_kernel<false>));
  MFB_CUDA(cudaFuncSetAttribute(ccd_update_flat_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCcdStagedMaxBytes));
  MFB_CUDA(cudaFuncSetAttribute(ccd_update_flat_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCcdStagedMaxBytes));
  MFB_CUDA(cudaFuncSetAttribute(ccd_resid_flat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCcdStagedMaxBytes));
  MFB_CUDA(cudaFuncGetAttributes(&fa, ccd_finalize_kernel));
  MFB_CUDA(cudaFuncGetAttributes(&fa, col_extract_kernel));
  MFB_CUDA(cudaFuncGetAttributes(&fa, col_insert_kernel));
  return 0;
}

int ccdpp_begin_impl(mfb_engine *e) {
  DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  MFB_TRY(ccd_preload_kernels(e));
  size_t nn = (size_t)(m.nnz > 0 ? m.nnz : 1);
  if (!e->res_row) MFB_CUDA(dev_alloc(&e->res_row, sizeof(float) * nn));
  if (!e->res_col) MFB_CUDA(dev_alloc(&e->res_col, sizeof(float) * nn));
  if (!e->uk) MFB_CUDA(dev_alloc(&e->uk, uk_alloc_bytes(e)));
  if (!e->vk) MFB_CUDA(dev_alloc(&e->vk, sizeof(float) * ((size_t)e->n_items + 4)));
  // res = gk_csr_Dup(trainMat) (modelMF.cpp:1013); uFac.fill(0) (:1020)
  MFB_CUDA(cudaMemcpyAsync(e->res_row, m.rowval, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToDevice, st));
  MFB_CUDA(cudaMemcpyAsync(e->res_col, m.colval, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToDevice, st));
  MFB_CUDA(cudaMemsetAsync(e->U, 0, sizeof(float) * (size_t)e->n_users * e->ld, st));
  const int ulo = e->row_begin[MFB_USER], uhi = e->row_end[MFB_USER], ilo = e->row_begin[MFB_ITEM], ihi = e->row_end[MFB_ITEM];
  if (!m.ccd_rows.built) {
    MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, ulo, uhi, kCcdChunk, &m.ccd_rows, e->opt_ccd_stream == 0));
    if (e->opt_ccd_stream == 2) MFB_TRY(build_chunk_plan(e, &m.ccd_rows));
    const int64_t nnz_side = (int64_t)((double)m.nnz * (double)(uhi - ulo) / std::max(e->n_users, 1));
    m.ccd_rows_mode = ccd_side_mode(e, nnz_side, uhi - ulo, e->n_items, true);  // the row passes gather v_k
    if (m.ccd_rows_mode == 2) {
      MFB_TRY(build_block_seg_plan(e, m.rowptr, m.rowind, e->bad_user, ulo, uhi, (e->n_items + kCcdBlock - 1) / kCcdBlock,
                                   &m.ccd_rows_blk, &m.ccd_rows_blk_off, &m.ccd_rows_doff));
    } else if (m.ccd_rows_mode == 1) {
      const int32_t off2[2] = {0, m.ccd_rows.n_seg};
      if (!m.ccd_rows_doff) MFB_CUDA(dev_alloc(&m.ccd_rows_doff, sizeof(off2)));
      MFB_CUDA(cudaMemcpy(m.ccd_rows_doff, off2, sizeof(off2), cudaMemcpyHostToDevice));
    }
  }
  if (!m.ccd_cols.built) {
    MFB_TRY(build_seg_plan(e, m.colptr, e->n_items, e->bad_item, ilo, ihi, kCcdChunk, &m.ccd_cols, e->opt_ccd_stream == 0));
    if (e->opt_ccd_stream == 2) MFB_TRY(build_chunk_plan(e, &m.ccd_cols));
    const int64_t nnz_side = (int64_t)((double)m.nnz * (double)(ihi - ilo) / std::max(e->n_items, 1));
    m.ccd_cols_mode = ccd_side_mode(e, nnz_side, ihi - ilo, e->n_users, false);  // the column passes gather u_k
    if (m.ccd_cols_mode == 2) {
      MFB_TRY(build_block_seg_plan(e, m.colptr, m.colind, e->bad_item, ilo, ihi, (e->n_users + kCcdBlock - 1) / kCcdBlock,
                                   &m.ccd_cols_blk, &m.ccd_cols_blk_off, &m.ccd_cols_doff));
    } else if (m.ccd_cols_mode == 1) {
      const int32_t off2[2] = {0, m.ccd_cols.n_seg};
      if (!m.ccd_cols_doff) MFB_CUDA(dev_alloc(&m.ccd_cols_doff, sizeof(off2)));
      MFB_CUDA(cudaMemcpy(m.ccd_cols_doff, off2, sizeof(off2), cudaMemcpyHostToDevice));
    }
  }
  size_t slots = (size_t)std::max(std::max(m.ccd_rows.n_multi, m.ccd_cols.n_multi), std::max(m.ccd_rows_blk.n_multi, m.ccd_cols_blk.n_multi));
  if (slots > e->ccd_acc_slots) {
    if (e->ccd_acc) MFB_CUDA(dev_free(e->ccd_acc));
    e->ccd_acc = nullptr;
    MFB_CUDA(dev_alloc(&e->ccd_acc, sizeof(double) * 2 * slots));
    e->ccd_acc_slots = slots;
  }
  if (e->ccd_acc) MFB_CUDA(cudaMemsetAsync(e->ccd_acc, 0, sizeof(double) * 2 * e->ccd_acc_slots, st));
  return 0;
}

namespace {
// one side (rows = users over the CSR residual, columns = items over the CSC residual) of the rank-one step
struct CcdSide {
  CcdPass plain, blk;
  int mode, nb, gather_n;
  const int32_t *doff;
  const SegPlan *sp_plain, *sp_blk;
};

template <bool ADDBACK, bool TWO>
int ccd_launch_update_sm(mfb_engine *e, const CcdPass &p, float *own, const float *other, const float *other_old, float reg,
                         const Aux *aux, int thresh, const CcdPeers &pe, const CcdSide &sd) {
  const int n_stage = (std::min(kCcdBlock, sd.gather_n) + 3) & ~3;
  const size_t smem = sizeof(float) * (size_t)(TWO ? kCcdBlock + n_stage : n_stage);
  CcdBlk bk{sd.doff, sd.nb, kCcdBlock, sd.gather_n, std::max(1, e->sm_count * (smem > 100 * 1024 ? 1 : 2) / sd.nb)};
  MFB_CUDA(cudaFuncSetAttribute(ccd_update_sm_kernel<ADDBACK, TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MFB_LAUNCH((ccd_update_sm_kernel<ADDBACK, TWO>), sd.nb * bk.ctas_per_blk, 512, smem, e->stream, p, own, other, other_old, reg,
             e->ccd_acc, aux, thresh, pe, bk);
  return 0;
}

int ccd_launch_resid_sm(mfb_engine *e, const CcdPass &p, const float *own, const float *other, float sign, const CcdSide &sd) {
  const size_t smem = sizeof(float) * (size_t)((std::min(kCcdBlock, sd.gather_n) + 3) & ~3);
  CcdBlk bk{sd.doff, sd.nb, kCcdBlock, sd.gather_n, std::max(1, e->sm_count * 2 / sd.nb)};
  MFB_CUDA(cudaFuncSetAttribute(ccd_resid_sm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MFB_LAUNCH(ccd_resid_sm_kernel, sd.nb * bk.ctas_per_blk, 512, smem, e->stream, p, own, other, sign, bk);
  return 0;
}

// own[row] = sum res * other / (reg + sum other^2) over the side's rows; addback / subtract: the fused forms of
// ccd_update_kernel (subtract only exists for the plain kernels; the staged passes subtract in a pass of their own)
int ccd_update_side(mfb_engine *e, const CcdSide &sd, float *own, const float *other, const float *other_old, float reg,
                    const Aux *aux, int thresh, const CcdPeers &pe, bool addback, bool subtract) {
  cudaStream_t st = e->stream;
  const int tb = 256, wpb = tb / 32;
  const SegPlan *sp = sd.mode == 2 ? sd.sp_blk : sd.sp_plain;
  const CcdPass &p = sd.mode == 2 ? sd.blk : sd.plain;
  if (p.n_seg > 0) {
    if (sd.mode == 0 && e->opt_ccd_stream == 2 && sp->n_chunk > 0) {
      const int grid = (sp->n_chunk + wpb - 1) / wpb;
      // the whole gathered vector staged in shared memory when it fits and the add-back reads the same vector
      const bool staged = ccd_flat_staged(e, sd.gather_n) && !subtract && (!addback || other_old == other);
      const size_t sm = ccd_flat_smem(sd.gather_n);
      if (staged && addback) MFB_LAUNCH((ccd_update_flat_kernel<true, false, true>), e->sm_count, kCcdStagedThreads, sm, st, p, sp->chunk_seg, sp->n_chunk, own, other, other_old, reg, e->ccd_acc, aux, thresh, pe, sd.gather_n);
      else if (staged) MFB_LAUNCH((ccd_update_flat_kernel<false, false, true>), e->sm_count, kCcdStagedThreads, sm, st, p, sp->chunk_seg, sp->n_chunk, own, other, other, reg, e->ccd_acc, aux, thresh, pe, sd.gather_n);
      else if (addback) MFB_LAUNCH((ccd_update_flat_kernel<true, false, false>), grid, tb, 0, st, p, sp->chunk_seg, sp->n_chunk, own, other, other_old, reg, e->ccd_acc, aux, thresh, pe, sd.gather_n);
      else if (subtract) MFB_LAUNCH((ccd_update_flat_kernel<false, true, false>), grid, tb, 0, st, p, sp->chunk_seg, sp->n_chunk, own, other, other, reg, e->ccd_acc, aux, thresh, pe, sd.gather_n);
      else MFB_LAUNCH((ccd_update_flat_kernel<false, false, false>), grid, tb, 0, st, p, sp->chunk_seg, sp->n_chunk, own, other, other, reg, e->ccd_acc, aux, thresh, pe, sd.gather_n);
    } else if (sd.mode == 0) {
      const int grid = (p.n_seg + wpb - 1) / wpb;
      if (addback) MFB_LAUNCH((ccd_update_kernel<true, false>), grid, tb, 0, st, p, own, other, other_old, reg, e->ccd_acc, aux, thresh, pe);
      else if (subtract) MFB_LAUNCH((ccd_update_kernel<false, true>), grid, tb, 0, st, p, own, other, other, reg, e->ccd_acc, aux, thresh, pe);
      else MFB_LAUNCH((ccd_update_kernel<false, false>), grid, tb, 0, st, p, own, other, other, reg, e->ccd_acc, aux, thresh, pe);
    } else if (addback && other_old != other) {
      MFB_TRY((ccd_launch_update_sm<true, true>(e, p, own, other, other_old, reg, aux, thresh, pe, sd)));
    } else if (addback) {
      MFB_TRY((ccd_launch_update_sm<true, false>(e, p, own, other, other, reg, aux, thresh, pe, sd)));
    } else {
      MFB_TRY((ccd_launch_update_sm<false, false>(e, p, own, other, other, reg, aux, thresh, pe, sd)));
    }
  }
  if (sp->n_multi)
    MFB_LAUNCH(ccd_finalize_kernel, (sp->n_multi + 255) / 256, 256, 0, st, sp->multi_row, sp->n_multi, e->ccd_acc, own, reg, aux,
               thresh, pe);
  return 0;
}

int ccd_resid_side(mfb_engine *e, const CcdSide &sd, const float *own, const float *other, float sign, int only_multi) {
  const CcdPass &p = sd.mode == 2 ? sd.blk : sd.plain;
  if (p.n_seg <= 0) return 0;
  if (sd.mode == 0 && e->opt_ccd_stream == 2 && !only_multi && sd.sp_plain->n_chunk > 0) {
    if (ccd_flat_staged(e, sd.gather_n))
      MFB_LAUNCH(ccd_resid_flat_kernel<true>, e->sm_count, kCcdStagedThreads, ccd_flat_smem(sd.gather_n), e->stream, p, sd.sp_plain->chunk_seg,
                 sd.sp_plain->n_chunk, own, other, sign, sd.gather_n);
    else
      MFB_LAUNCH(ccd_resid_flat_kernel<false>, (sd.sp_plain->n_chunk + 7) / 8, 256, 0, e->stream, p, sd.sp_plain->chunk_seg, sd.sp_plain->n_chunk,
                 own, other, sign, sd.gather_n);
    return 0;
  }
  if (sd.mode == 0) {
    const int tb = 256, wpb = tb / 32;
    MFB_LAUNCH(ccd_resid_kernel, (p.n_seg + wpb - 1) / wpb, tb, 0, e->stream, p, own, other, sign, only_multi);
    return 0;
  }
  return ccd_launch_resid_sm(e, p, own, other, sign, sd);
}
}  // namespace

int ccdpp_rank1_impl(mfb_engine *e, int32_t k, int first_iter, int32_t inner, float ureg, float ireg,
                     int32_t item_freq_thresh) {
  DevCsr &m = e->mat[MFB_TRAIN];
  cudaStream_t st = e->stream;
  const SegPlan &rp = m.ccd_rows, &cp = m.ccd_cols, &rb = m.ccd_rows_blk, &cb = m.ccd_cols_blk;
  CcdSide rows{{m.rowind, e->res_row, rp.row, rp.start, rp.len, rp.slot, rp.n_seg},
               {m.rowind, e->res_row, rb.row, rb.start, rb.len, rb.slot, rb.n_seg},
               m.ccd_rows_mode, m.ccd_rows_mode == 2 ? (int)m.ccd_rows_blk_off.size() - 1 : 1, e->n_items, m.ccd_rows_doff, &rp, &rb};
  CcdSide cols{{m.colind, e->res_col, cp.row, cp.start, cp.len, cp.slot, cp.n_seg},
               {m.colind, e->res_col, cb.row, cb.start, cb.len, cb.slot, cb.n_seg},
               m.ccd_cols_mode, m.ccd_cols_mode == 2 ? (int)m.ccd_cols_blk_off.size() - 1 : 1, e->n_users, m.ccd_cols_doff, &cp, &cb};
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
  const bool fuse_add = !first_iter && inner >= 1 && e->opt_ccd_fuse;          // add-back rides on the first updates
  float *uk_old = e->uk + e->uk_old_offset();  // the column add-back needs u_k as it was before its first update
  if (uhi > ulo)
    MFB_LAUNCH(col_extract_kernel, (uhi - ulo + 255) / 256, 256, 0, st, e->U, e->ld, k, ulo, uhi, e->uk, pu,
               fuse_add ? e->uk_old_offset() : (size_t)0);
  if (ihi > ilo) MFB_LAUNCH(col_extract_kernel, (ihi - ilo + 255) / 256, 256, 0, st, e->V, e->ld, k, ilo, ihi, e->vk, pv, (size_t)0);
  MFB_TRY(comm_barrier_launch(e));
  // the FreqAdap rule zeroes v_k of infrequent items for k > 0 (modelMF.cpp:1336-1342)
  const int thresh = (item_freq_thresh > 0 && k > 0) ? item_freq_thresh : 0;
  // the column subtract rides on the last v_k update pass (plain kernels only)
  const bool fuse_sub = inner >= (fuse_add ? 2 : 1) && e->opt_ccd_fuse && cols.mode == 0;
  if (!first_iter && !fuse_add) {
    MFB_TRY(ccd_resid_side(e, rows, e->uk, e->vk, 1.0f, 0));
    MFB_TRY(ccd_resid_side(e, cols, e->vk, e->uk, 1.0f, 0));
  }
  for (int s = 0; s < inner; s++) {
    MFB_TRY(ccd_update_side(e, rows, e->uk, e->vk, e->vk, ureg, e->aux_u, 0, pu, s == 0 && fuse_add, false));
    MFB_TRY(comm_barrier_launch(e));
    MFB_TRY(ccd_update_side(e, cols, e->vk, e->uk, uk_old, ireg, e->aux_i, thresh, pv, s == 0 && fuse_add,
                            s == inner - 1 && fuse_sub && !(s == 0 && fuse_add)));
    MFB_TRY(comm_barrier_launch(e));
  }
  MFB_TRY(ccd_resid_side(e, rows, e->uk, e->vk, -1.0f, 0));
  if (fuse_sub) {
    if (cp.n_multi) MFB_TRY(ccd_resid_side(e, cols, e->vk, e->uk, -1.0f, 1));
  } else {
    MFB_TRY(ccd_resid_side(e, cols, e->vk, e->uk, -1.0f, 0));
  }
  CcdPeerMats mu, mv;
  mu.n = mv.n = 0;
  if (c.connected)
    for (int p = 0; p < c.world; p++)
      if (p != c.rank) { mu.p[mu.n++] = c.U[p]; mv.p[mv.n++] = c.V[p]; }
  if (uhi > ulo) MFB_LAUNCH(col_insert_kernel, (uhi - ulo + 255) / 256, 256, 0, st, e->U, e->ld, k, ulo, uhi, e->uk, mu);
  if (ihi > ilo) MFB_LAUNCH(col_insert_kernel, (ihi - ilo + 255) / 256, 256, 0, st, e->V, e->ld, k, ilo, ihi, e->vk, mv);
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
