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

#include <algorithm>

namespace mfb {

constexpr int kAlsTile = 32;    // factor rows staged per pipeline stage

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

// ---- blocked Cholesky + both triangular solves on the register-tiled matrix ---------------------
// The 16 x 16 thread grid holds A as TR x TR tiles: thread (tx, ty) owns rows tile_idx(tx, .) and
// columns tile_idx(ty, .) (A is symmetric, so the accumulation's acc[i][j] is read as T(a, b) =
// acc[b][a]).  Right-looking block algorithm over the 16 tile columns j (a symmetric permutation of
// the matrix, which an SPD factorisation does not mind):
//   1. thread (j, j) factors its diagonal tile in registers and forward-substitutes its slice of the
//      right-hand side (the rhs rides along as one more row of the matrix);
//   2. threads (R > j, C = j) — half a warp — solve their tile against L_jj^T and publish it;
//   3. threads (R >= C > j) subtract P_R P_C^T from their tile, diagonal threads also update the rhs.
// Two barriers per tile column instead of one per scalar column; the strictly-lower tiles keep L in
// registers, so the backward substitution L^T x = y needs only the 16 x TR vector in shared memory.
template <int N>
__device__ __forceinline__ void lds_vec(const float *p, float (&o)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; q++) {
      const float4 v = *reinterpret_cast<const float4 *>(p + 4 * q);
      o[4 * q] = v.x; o[4 * q + 1] = v.y; o[4 * q + 2] = v.z; o[4 * q + 3] = v.w;
    }
  } else if constexpr (N == 2) {
    const float2 v = *reinterpret_cast<const float2 *>(p);
    o[0] = v.x; o[1] = v.y;
  } else {
#pragma unroll
    for (int q = 0; q < N; q++) o[q] = p[q];
  }
}
template <int N>
__device__ __forceinline__ void sts_vec(float *p, const float (&o)[N]) {
  if constexpr (N % 4 == 0) {
#pragma unroll
    for (int q = 0; q < N / 4; q++)
      *reinterpret_cast<float4 *>(p + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
  } else if constexpr (N == 2) {
    *reinterpret_cast<float2 *>(p) = make_float2(o[0], o[1]);
  } else {
#pragma unroll
    for (int q = 0; q < N; q++) p[q] = o[q];
  }
}

template <int TR>
struct CholScratch {  // floats, carved from the (by then free) L / tile region
  static constexpr int T2 = TR * TR;
  static constexpr int TS = T2 + 4;                       // padded tile stride: conflict-free 128-bit loads
  static constexpr int off_pan = 0;                       // [2][16][TS] panel tiles, transposed
  static constexpr int off_dg = 2 * 16 * TS;              // [2][T2 + TR] L_jj and 1 / diag
  static constexpr int off_y = off_dg + 2 * (T2 + TR);    // [16][TR] y, then z, then x
  static constexpr int total = off_y + 16 * TR;
};

#define MFB_T(a, b) acc[b][a]
struct CtaBarrier {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
// barrier over a sub-group of warps of the CTA (bar.sync id, count): the solver groups of the warp-specialised kernel
struct NamedBarrier {
  int id, count;
  __device__ __forceinline__ void operator()() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
};
// (R, C) = tile row / tile column of the thread (R >= C works, everything else idles through the barriers; threads
// without a tile pass R = -1, C = -2); pan / dg / yb = CholScratch regions, bv = right-hand side in, solution out.
template <int TR, class Bar>
__device__ __forceinline__ void chol_solve_core(float (&acc)[TR][TR], float *pan, float *dg, float *yb, float *bv, int R, int C,
                                                const Bar &bar) {
  using K = CholScratch<TR>;
  const bool on_diag = R == C;
  float rhs[TR], dinv[TR];
  bar();  // the right-hand side is complete, the scratch region is free
#pragma unroll
  for (int a = 0; a < TR; a++) {
    rhs[a] = on_diag ? bv[tile_idx<TR>(R, a)] : 0.f;
    dinv[a] = 0.f;
  }
  for (int j = 0; j < 16; j++) {
    float *pj = pan + (j & 1) * 16 * K::TS;
    float *dj = dg + (j & 1) * (K::T2 + TR);
    if (on_diag && R == j) {
#pragma unroll
      for (int c = 0; c < TR; c++) {
        const float d = MFB_T(c, c);
        float r = d > 0.f ? rsqrtf(d) : 0.f;
        r = r * fmaf(-0.5f * d * r, r, 1.5f);  // one Newton step on the hardware approximation
        dinv[c] = r;
        MFB_T(c, c) = d * r;
        rhs[c] *= r;
#pragma unroll
        for (int a = c + 1; a < TR; a++) {
          MFB_T(a, c) *= r;
          rhs[a] = fmaf(-MFB_T(a, c), rhs[c], rhs[a]);
        }
#pragma unroll
        for (int a = c + 1; a < TR; a++)
#pragma unroll
          for (int b = c + 1; b <= a; b++) MFB_T(a, b) = fmaf(-MFB_T(a, c), MFB_T(b, c), MFB_T(a, b));
      }
#pragma unroll
      for (int a = 0; a < TR; a++) {
        float row[TR];
#pragma unroll
        for (int b = 0; b < TR; b++) row[b] = MFB_T(a, b);
        sts_vec<TR>(dj + a * TR, row);
      }
      sts_vec<TR>(dj + K::T2, dinv);
      sts_vec<TR>(yb + j * TR, rhs);
    }
    bar();
    if (C == j && R > j) {
      float di[TR];
      lds_vec<TR>(dj + K::T2, di);
#pragma unroll
      for (int c = 0; c < TR; c++) {
        float Lc[TR];  // row c of L_jj, one row at a time keeps the register peak below the tile itself
        lds_vec<TR>(dj + c * TR, Lc);
#pragma unroll
        for (int a = 0; a < TR; a++) {
          float sacc = MFB_T(a, c);
#pragma unroll
          for (int k = 0; k < c; k++) sacc = fmaf(-MFB_T(a, k), Lc[k], sacc);
          MFB_T(a, c) = sacc * di[c];
        }
      }
#pragma unroll
      for (int c = 0; c < TR; c++) {
        float col[TR];
#pragma unroll
        for (int a = 0; a < TR; a++) col[a] = MFB_T(a, c);
        sts_vec<TR>(pj + R * K::TS + c * TR, col);
      }
    }
    bar();
    if (C > j && R >= C) {
      float yj[TR];
#pragma unroll
      for (int c = 0; c < TR; c++) yj[c] = 0.f;
      if (on_diag) lds_vec<TR>(yb + j * TR, yj);
#pragma unroll
      for (int c = 0; c < TR; c++) {
        float pr[TR], pc[TR];
        lds_vec<TR>(pj + R * K::TS + c * TR, pr);
        lds_vec<TR>(pj + C * K::TS + c * TR, pc);
#pragma unroll
        for (int a = 0; a < TR; a++) {
#pragma unroll
          for (int b = 0; b < TR; b++) MFB_T(a, b) = fmaf(-pr[a], pc[b], MFB_T(a, b));
          rhs[a] = fmaf(-pr[a], yj[c], rhs[a]);
        }
      }
    }
  }
  // L^T x = y, from the last tile column backwards
  for (int i = 15; i >= 0; i--) {
    if (on_diag && R == i) {
      float z[TR], x[TR];
      lds_vec<TR>(yb + i * TR, z);
#pragma unroll
      for (int c = TR - 1; c >= 0; c--) {
        float sacc = z[c];
#pragma unroll
        for (int a = c + 1; a < TR; a++) sacc = fmaf(-MFB_T(a, c), x[a], sacc);
        x[c] = sacc * dinv[c];
      }
      sts_vec<TR>(yb + i * TR, x);
#pragma unroll
      for (int c = 0; c < TR; c++) bv[tile_idx<TR>(i, c)] = x[c];
    }
    bar();
    if (R == i && C < i) {
      float x[TR], z[TR];
      lds_vec<TR>(yb + i * TR, x);
      lds_vec<TR>(yb + C * TR, z);
#pragma unroll
      for (int c = 0; c < TR; c++) {
        float sacc = z[c];
#pragma unroll
        for (int a = 0; a < TR; a++) sacc = fmaf(-MFB_T(a, c), x[a], sacc);
        z[c] = sacc;
      }
      sts_vec<TR>(yb + C * TR, z);
    }
    bar();
  }
}
#undef MFB_T

template <int TR, class S>
__device__ __forceinline__ void chol_solve(float (&acc)[TR][TR], float *sm, int tx, int ty) {
  using K = CholScratch<TR>;
  static_assert(K::total <= S::RP * S::LDL, "Cholesky scratch must fit the L region");
  chol_solve_core<TR>(acc, sm + S::off_L + K::off_pan, sm + S::off_L + K::off_dg, sm + S::off_L + K::off_y, sm + S::off_b, tx, ty,
                      CtaBarrier());
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
  const int seg = a.seg0 + blockIdx.x;
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
  chol_solve<TR, AlsSmem<TR>>(acc, sm, tx, ty);
  store_solution(a, row, tid, bv);
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
  chol_solve<TR, AlsSmem<TR>>(acc, sm, tx, ty);
  store_solution(a, row, tid, bv);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core Gram for rank 65..128 (padded to 128): tcgen05.mma kind::tf32, fp32 accumulator in
// TMEM, 3xTF32 split so that the result keeps fp32-level accuracy (the 1e-4 parity bar of the
// reference's fp32 ALS cannot be met by a single tf32 product: 10-bit mantissa).
//
//   G = sum_k f_k f_k^T over the gathered factor rows f_k = Fin[ind[k]]   (K = ratings)
//   D[128 x 128] += A[128 x 8] * B[8 x 128] per instruction with A(m,k) = B(n,k) = f_k[m]: both
//   operands are the SAME shared-memory tile.  kind::tf32 only multiplies K-major operands (an
//   MN-major tile — the layout the gathered rows arrive in — returns zeros, measured with
//   tools/tc_probe.cu), so the tile is transposed on the way into shared memory: every thread
//   loads the same 16-byte unit of four consecutive ratings and stores four 16-byte K-vectors.
//   Layout: no-swizzle K-major core matrices (8 factor dims x 4 ratings = 128 B), core stride along
//   the factor dim SBO = 144 B (16 B of padding makes the transposing stores conflict free), core
//   stride along the ratings LBO = 16 * 144 B.
//   f = big + small with big = f rounded to tf32, small = f - big (exact in fp32) rounded to tf32
//   (|small| <= 2^-11 |f|, its rounding error <= 2^-22 |f|); G ~= big big^T + big small^T + small big^T
//   (the dropped small small^T term is 2^-22 relative).
//
// 256 threads: all of them gather / split / store the tiles (and accumulate the right-hand side
// b = sum r f on CUDA cores), thread 0 issues the MMAs; tcgen05.commit on an mbarrier frees a
// stage for the next gather, so tile t+1 is loaded while tile t multiplies.  The epilogue pulls
// the accumulator out of TMEM (tcgen05.ld 32x32b), and the Cholesky solve runs as in the CUDA-core
// kernel.
constexpr int kTcKT = 32;       // ratings per stage
constexpr int kTcStages = 2;
constexpr uint32_t kTcSbo = 144;                       // bytes between 8-dim core matrices
constexpr uint32_t kTcLbo = 16 * kTcSbo;               // bytes between 4-rating core matrices
constexpr uint32_t kTcPartBytes = (kTcKT / 4) * kTcLbo;  // one (stage, big|small) tile: 18432 B

struct AlsTcSmem {
  static constexpr int RP = 128;
  static constexpr int LDL = RP + 1;
  // byte layout; the tile region is reused for the L matrix once the MMAs are done
  static constexpr uint32_t tiles_bytes = kTcStages * 2 * kTcPartBytes;  // 73728
  static constexpr uint32_t L_bytes = RP * LDL * 4;                       // 66048
  static constexpr uint32_t region0 = tiles_bytes;                        // max of both
  static constexpr int off_tile = 0;
  static constexpr int off_L = 0;
  static constexpr int off_col = region0 / 4;
  static constexpr int off_b = off_col + 2 * RP;
  static constexpr int off_bred = off_b + RP;            // [8][RP] partial right-hand sides
  static constexpr int off_misc = off_bred + 8 * RP;     // mbarriers (3 x 8 B), tmem address
  static constexpr size_t bytes = sizeof(float) * (off_misc + 16) + 1024;  // + slack for 1024-byte alignment
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Waits for the phase of `parity` to complete.  Bounded: a protocol bug must not hang the GPU — after ~4 s without
// progress the kernel traps (the launch fails with an error the host reports) instead of spinning for ever.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// the same with a suspend-time hint: the hardware parks the warp until the phase completes or `ns` have passed
// (NANOSLEEP.SYNCS), so a waiting warp leaves the issue slots to the warps that work — most warps of the
// warp-specialised kernel wait most of the time
__device__ __forceinline__ bool mbar_try_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0 = 0;
  for (uint32_t spins = 1;; spins++) {
    if (mbar_try_wait_parked(bar, parity, 20000u)) return;
    if ((spins & 0x3Fu) == 0) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) asm volatile("trap;");
    }
  }
}
// SM100 shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4),
// descriptor version 1, no swizzle.  K-major: LBO = stride between core matrices along K,
// SBO = stride between 8-row core matrices along M/N (verified with tools/tc_probe.cu).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// round-to-nearest fp32 -> tf32 (the tensor core itself truncates the low 13 mantissa bits)
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// round to nearest (ties away from zero) on the 10-bit tf32 mantissa for finite values: two integer instructions
// where cvt.rna.tf32.f32 is expanded into four (it also keeps inf / NaN intact)
__device__ __forceinline__ float round_tf32_fast(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 2) als_gram_tc_kernel(const AlsArgs a) {
  using S = AlsTcSmem;
  constexpr int RP = S::RP, TR = 8;
  extern __shared__ uint8_t sm_raw[];
  // keep the tiles on a 1024-byte boundary (not required without swizzle, harmless)
  uint8_t *smb = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  float *sm = reinterpret_cast<float *>(smb);
  float *bv = sm + S::off_b, *bred = sm + S::off_bred;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + S::off_misc);  // [0..1] stage free, [2] all done
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int seg = a.seg0 + blockIdx.x;
  const int row = a.seg_row[seg], start = a.seg_start[seg], len = a.seg_len[seg], slot = a.seg_slot[seg];
  const int nq = a.ld >> 2;  // 16-byte units per factor row (<= 32)

  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t tiles = smem_u32(smb);
  const int ntiles = (len + kTcKT - 1) / kTcKT;
  const int q = lane;        // 16-byte unit of the factor row handled by this thread
  float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t phase[kTcStages] = {0, 0};

  // gather of one 32-rating tile: warp w takes ratings 4w..4w+3, lane q the q-th 16-byte unit of the row
  float4 f[4];
  float rt[4];
  auto gather = [&](int t) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int j = t * kTcKT + warp * 4 + i;
      f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      rt[i] = 0.f;
      if (j < len) {
        const int it = __ldg(a.ind + start + j);
        const float r = __ldg(a.val + start + j);
        if (r > 0.f) {  // modelMF.cpp:819
          rt[i] = r;
          if (q < nq) f[i] = __ldg(reinterpret_cast<const float4 *>(a.Fin + (size_t)it * a.ld) + q);
        }
      }
    }
  };
  gather(0);
  for (int t = 0; t < ntiles; t++) {
    const int s = t & 1;
    if (t >= kTcStages) {  // the MMAs that read this stage two tiles ago must have finished
      mbar_wait(smem_u32(&bars[s]), phase[s]);
      phase[s] ^= 1;
    }
    const uint32_t big0 = tiles + (uint32_t)(s * 2) * kTcPartBytes, small0 = big0 + kTcPartBytes;
    // transpose in registers: for factor dim 4q+e the four ratings form one 16-byte K-vector
    const float fe[4][4] = {{f[0].x, f[1].x, f[2].x, f[3].x}, {f[0].y, f[1].y, f[2].y, f[3].y},
                            {f[0].z, f[1].z, f[2].z, f[3].z}, {f[0].w, f[1].w, f[2].w, f[3].w}};
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int mdim = 4 * q + e;
      float bg[4], sm4[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        bg[i] = round_tf32(fe[e][i]);
        sm4[i] = round_tf32(fe[e][i] - bg[i]);
      }
      const uint32_t off = (uint32_t)warp * kTcLbo + (uint32_t)(mdim >> 3) * kTcSbo + (uint32_t)(mdim & 7) * 16;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(big0 + off), "f"(bg[0]), "f"(bg[1]), "f"(bg[2]), "f"(bg[3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(small0 + off), "f"(sm4[0]), "f"(sm4[1]), "f"(sm4[2]), "f"(sm4[3]) : "memory");
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      bacc.x = fmaf(rt[i], f[i].x, bacc.x);
      bacc.y = fmaf(rt[i], f[i].y, bacc.y);
      bacc.z = fmaf(rt[i], f[i].z, bacc.z);
      bacc.w = fmaf(rt[i], f[i].w, bacc.w);
    }
    if (t + 1 < ntiles) gather(t + 1);  // in flight while this tile is fenced and multiplied
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int g = 0; g < kTcKT / 8; g++) {
        const uint64_t db = umma_desc_kmajor(big0 + g * 2 * kTcLbo, kTcLbo, kTcSbo);
        const uint64_t ds = umma_desc_kmajor(small0 + g * 2 * kTcLbo, kTcLbo, kTcSbo);
        umma_tf32(tmem_d, db, db, idesc, (t > 0 || g > 0) ? 1u : 0u);
        umma_tf32(tmem_d, db, ds, idesc, 1u);
        umma_tf32(tmem_d, ds, db, idesc, 1u);
      }
      umma_commit(smem_u32(&bars[s]));
      if (t == ntiles - 1) umma_commit(smem_u32(&bars[2]));
    }
  }
  // right-hand side: reduce the 8 row slots
  *reinterpret_cast<float4 *>(bred + warp * RP + q * 4) = bacc;
  mbar_wait(smem_u32(&bars[2]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();  // every thread is past the MMAs: the tile region may be overwritten
  if (tid < RP) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) sacc += bred[w * RP + tid];
    bv[tid] = sacc;
  }
  // accumulator: TMEM lane = Gram row, column = Gram column; warps 0..3 own lanes 32w..32w+31
  float *Lm = sm + S::off_L;
  if (warp < 4) {
    const int grow = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < RP; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 32; c++) Lm[grow * S::LDL + c0 + c] = __uint_as_float(v[c]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");

  float acc[TR][TR];
#pragma unroll
  for (int i = 0; i < TR; i++)
#pragma unroll
    for (int j = 0; j < TR; j++) acc[i][j] = Lm[tile_idx<TR>(ty, i) * S::LDL + tile_idx<TR>(tx, j)];
  if (slot >= 0) {
    float *w = a.ws + (size_t)slot * (RP * RP + RP);
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TR; j++) atomicAdd(w + tile_idx<TR>(ty, i) * RP + tile_idx<TR>(tx, j), acc[i][j]);
    if (tid < RP) atomicAdd(w + RP * RP + tid, bv[tid]);
    return;
  }
  __syncthreads();  // everyone holds its tile: L may now be overwritten by the factorisation
  add_reg_diag<TR>(acc, tx, ty, a.rank, a.reg);
  chol_solve<TR, S>(acc, sm, tx, ty);
  store_solution(a, row, tid, bv);
}

// ---------------------------------------------------------------------------------------------
// Short rows (fewer ratings than half the padded rank): the same solution through the dual system.
//   x = (F^T F + reg I)^-1 F^T r  =  F^T (F F^T + reg I)^-1 r          (push-through identity)
// with F the len x rank matrix of gathered factor rows.  The dual Gram F F^T is len x len, so a user
// with 20 ratings costs a 32 x 32 factorisation instead of a 128 x 128 one — and most rows of a
// ratings matrix are short (power-law degrees).  Same tile Cholesky as above on a 16*TR-padded system
// (padding rows are zero: their diagonal is reg, their right-hand side 0, their alpha 0).
template <int TR>
struct AlsDualSmem {
  static constexpr int RP = 16 * TR;
  static constexpr int LDL = RP + 1;
  static constexpr int LDF = 132;  // row stride of the staged factor rows: 128 + 4 keeps 128-bit loads conflict free
  static constexpr int off_F = 0;
  static constexpr int off_L = off_F + RP * LDF;
  static constexpr int off_b = off_L + (RP * LDL + 3) / 4 * 4;
  static constexpr int total_floats = off_b + RP;
  static constexpr size_t bytes = sizeof(float) * total_floats;
};

template <int TR>
__global__ void __launch_bounds__(256, 4) als_dual_kernel(const AlsArgs a) {
  using S = AlsDualSmem<TR>;
  constexpr int RP = S::RP, LDF = S::LDF;
  extern __shared__ __align__(16) float sm[];
  float *F = sm + S::off_F, *bv = sm + S::off_b;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int seg = a.seg0 + blockIdx.x;
  const int row = a.seg_row[seg], start = a.seg_start[seg], len = a.seg_len[seg];
  const int nq = a.ld >> 2;
  // stage the rated factor rows (rating > 0, modelMF.cpp:819); staged row m sits in tile (m % 16),
  // element m / 16, i.e. at position (m % 16) * TR + m / 16 of the right-hand side
  for (int i = tid; i < RP * 32; i += 256) {
    const int m = i >> 5, q = i & 31;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m < len && q < nq && __ldg(a.val + start + m) > 0.f)
      v = __ldg(reinterpret_cast<const float4 *>(a.Fin + (size_t)__ldg(a.ind + start + m) * a.ld) + q);
    *reinterpret_cast<float4 *>(F + m * LDF + q * 4) = v;
  }
  if (tid < RP) {
    float r = 0.f;
    if (tid < len) r = fmaxf(__ldg(a.val + start + tid), 0.f);
    bv[(tid & 15) * TR + (tid >> 4)] = r;
  }
  __syncthreads();
  float acc[TR][TR];
#pragma unroll
  for (int i = 0; i < TR; i++)
#pragma unroll
    for (int j = 0; j < TR; j++) acc[i][j] = 0.f;
  for (int q = 0; q < nq; q++) {
    float4 fa[TR], fb[TR];
#pragma unroll
    for (int i = 0; i < TR; i++) {
      fa[i] = *reinterpret_cast<const float4 *>(F + (i * 16 + ty) * LDF + q * 4);
      fb[i] = *reinterpret_cast<const float4 *>(F + (i * 16 + tx) * LDF + q * 4);
    }
#pragma unroll
    for (int i = 0; i < TR; i++)
#pragma unroll
      for (int j = 0; j < TR; j++) {
        acc[i][j] = fmaf(fa[i].x, fb[j].x, acc[i][j]);
        acc[i][j] = fmaf(fa[i].y, fb[j].y, acc[i][j]);
        acc[i][j] = fmaf(fa[i].z, fb[j].z, acc[i][j]);
        acc[i][j] = fmaf(fa[i].w, fb[j].w, acc[i][j]);
      }
  }
  if (tx == ty) {
#pragma unroll
    for (int i = 0; i < TR; i++) acc[i][i] += a.reg;
  }
  chol_solve<TR, S>(acc, sm, tx, ty);  // alpha, in the right-hand side's layout
  if (tid < a.ld) {
    float x = 0.f;
    for (int m = 0; m < RP; m++) x = fmaf(bv[(m & 15) * TR + (m >> 4)], F[m * LDF + tid], x);
    if (tid >= a.rank) x = 0.f;
    a.Fout[(size_t)row * a.ld + tid] = x;
    for (int p = 0; p < a.n_peer; p++) a.Fpeer[p][(size_t)row * a.ld + tid] = x;
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised persistent ALS half-step for rank 33..128 (padded rank RP = 64 or 128): one CTA per SM, TMEM and the
// mbarriers set up once per CTA, rows taken round-robin from the longest-first segment list.
//
//   producers (4 warps)   fetch the rated factor rows of a 32-rating tile with 16-byte cp.async copies counted on an
//                         mbarrier (a ring of raw stages keeps several tiles in flight across row boundaries; the
//                         (item, rating) pairs are requested two tiles earlier still).  One cp.async.bulk per row was
//                         measured first: a 256 .. 512-byte bulk copy costs ~100 cycles of the SM's TMA unit, 3 - 4 k
//                         cycles per 32-row tile — profiles/r2_als.md — so scattered rows go through LDGSTS and TMA is
//                         kept for contiguous blocks (csrc/ccdpp.cu).  Then every value is split into tf32 big + small, transposed into the K-major
//                         operand layout of the tensor core and accumulate the right-hand side b = sum r f on CUDA
//                         cores; thread 0 issues the MMAs of the tile and commits them to mbarriers.
//   tensor core           ONE tcgen05.mma per 8 ratings gives both products of the 3xTF32 scheme: the operand tile holds
//                         [big dims; small dims] as 2 RP rows.  RP = 128: A = big (M = 128), B = [big; small] (N = 256):
//                         D = [big big^T | big small^T].  RP = 64: A = [big; small] (M = 128), B = big (N = 64):
//                         D = [big big^T ; small big^T].  G = BB + X + X^T is formed when the accumulator is drained
//                         (the third product of the scheme is the transpose of the second).  One accumulator per solver group in
//                         TMEM: the next rows multiply while earlier ones are drained.
//   solver groups         G groups of 5 warps; group g takes rows g, g + G, ...: four of its warps (one per TMEM lane
//                         quadrant) drain the accumulator into lower-triangle tiles in shared memory, 136 threads — one
//                         per tile (R >= C) of the 16 x 16 tile grid — hold a tile in registers and run the tile
//                         Cholesky + both triangular solves (chol_solve_core) with barriers over the group only, so G
//                         rows are being solved while the next Gram is accumulated.
constexpr int kWsTeam = 128;         // threads of a converter team
constexpr int kWsGroupThreads = 160;
constexpr uint32_t kWsSbo = 144;     // bytes between 8-dim core matrices (16 B of padding: conflict-free transposing stores)
// named barriers (hardware ids 1 .. 15): converter team t -> 1 + t (<= 4 teams), drain warps of solver group g ->
// kWsDrainBar0 + g, whole solver group g -> kWsGroupBar0 + g (<= 4 groups); the three ranges must not overlap
constexpr int kWsDrainBar0 = 5, kWsGroupBar0 = 9;
constexpr int kWsFirst = 0x100, kWsLast = 0x200;  // tile descriptor flags: first / last tile of its row
constexpr int kWsClassShift = 12;                 // tile descriptor: (tile index within the row) mod NT

// NT converter teams and G solver groups per CTA: a half-step over many short rows (users) is bound by the solves, one
// over few long rows (items) by the conversion of the gathered rows — the launcher picks the split per side.
template <int TR, int NT_, int G_>
struct AlsWs {
  static constexpr int RP = 16 * TR;
  static constexpr int QP = RP / 4;                               // 16-byte units per factor row
  static constexpr int KT = 512 / QP;                             // ratings per tile: 32 (RP = 64) / 16 (RP = 128) — a tile is
                                                                  // 128 units of (4 ratings x 16 bytes), one per team thread
  static constexpr int NGRP = KT / 4;                             // 4-rating K groups of a tile (8 / 4)
  static constexpr int NT = NT_;                                  // converter teams (tile g belongs to team g mod NT)
  static constexpr int G = G_;                                    // solver groups
  static constexpr int MCORES = 2 * RP / 8;                       // core matrices along M of [big; small]
  static constexpr uint32_t LBO = MCORES * kWsSbo;                // bytes between 4-rating K groups
  static constexpr uint32_t OP_BYTES = (KT / 4) * LBO;            // one operand stage (18 KB)
  static constexpr int NS = NT_ + (NT_ >= 4 ? 1 : 2);             // operand stages; >= NT: a team moves NT tiles ahead per step and
                                                                  // may be at most one phase of a stage's barrier ahead of the MMAs;
                                                                  // > NT: a team does not wait for the MMAs of its own previous tile
  static constexpr int NRT = NT_ >= 3 ? 3 : 4;                    // raw stages per team (tiles of gathered rows in flight); shared-memory budget
  static constexpr int MR = 8;                                    // (item, rating, descriptor) ring per team: the scheduler warp
                                                                  // runs up to MR team tiles ahead of the converters
  static constexpr uint32_t RAW_ROW = RP * 4;                     // bytes per staged factor row
  static constexpr uint32_t RAW_BYTES = KT * RAW_ROW;             // 8 KB
  // thread map: NT converter teams | NT scheduler warps | the warp that issues the tensor-core instructions | G solver groups
  static constexpr int SCHED0 = NT * kWsTeam;
  static constexpr int MMA0 = SCHED0 + NT * 32;
  static constexpr int SOLVER0 = MMA0 + 32;
  static constexpr int NTHREADS = SOLVER0 + G * kWsGroupThreads;
  static constexpr int ACC_COLS = TR == 8 ? 256 : 64;             // TMEM columns of one accumulator
  static constexpr int NACC = G == 1 ? 2 : G;                     // accumulators: row i multiplies into accumulator i mod NACC and
                                                                  // is drained by group i mod G — NACC is a multiple of G, so a
                                                                  // group sees the phases of an accumulator in order; a single
                                                                  // group gets two (the next row multiplies while it drains)
  static constexpr int TMEM_USED = NACC * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;  // allocations are powers of two
  static constexpr int T2 = TR * TR, TS = T2 + 4;                 // tile stride (floats): conflict-free 128-bit loads
  static constexpr int GS_FLOATS = 136 * TS;                      // lower-triangle tiles of one Gram
  // byte offsets from the 1024-aligned base
  static constexpr uint32_t off_op = 0;
  static constexpr uint32_t off_raw = off_op + NS * OP_BYTES;                 // [NT][NRT] stages
  static constexpr uint32_t off_gs = off_raw + NT * NRT * RAW_BYTES;
  static constexpr uint32_t off_bpart = off_gs + G * GS_FLOATS * 4;           // [NACC][NT * NGRP][RP] partial right-hand sides
  static constexpr uint32_t off_gbv = off_bpart + NACC * NT * NGRP * RP * 4;  // [G][RP]
  static constexpr uint32_t off_meta = off_gbv + G * RP * 4;                  // [NT][MR][32] item, then [NT][MR][32] rate
  static constexpr uint32_t off_desc = off_meta + NT * MR * 32 * 8;           // [NT][MR] {ratings in the tile | kWsFirst | kWsLast, row}
  static constexpr uint32_t off_opinfo = off_desc + NT * MR * 8;              // [NS] descriptor of the tile in each operand stage
  static constexpr uint32_t off_bars = off_opinfo + NS * 8;                   // mbarriers
  static constexpr int n_bars = 2 * NT * MR + 2 * NS + 3 * NACC;
  static constexpr uint32_t off_tmem = off_bars + n_bars * 8;
  static constexpr size_t bytes = off_tmem + 16 + 1024;
  static_assert(CholScratch<TR>::total <= GS_FLOATS, "Cholesky scratch is carved from the drained tile buffer");
  static_assert(bytes <= 227 * 1024, "shared memory budget");
  static_assert(NT_ >= 1 && NT_ <= 4 && G_ >= 1 && G_ <= 4, "named-barrier id ranges (kWsDrainBar0 / kWsGroupBar0)");
  static_assert(NTHREADS <= 1024, "threads per CTA");
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one factor row global -> shared through the bulk-copy engine (TMA, non-tensor form); completion is counted in bytes
// on the mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tile coordinate / element index of matrix index x under tile_idx<TR> (its inverse)
template <int TR>
__device__ __forceinline__ void tile_of(int x, int &t, int &e) {
  if (TR <= 4) { t = x / TR; e = x % TR; }
  else { t = (x & 63) >> 2; e = ((x >> 6) << 2) | (x & 3); }
}
__device__ __forceinline__ int tri_index(int R, int C) { return R * (R + 1) / 2 + C; }

template <int TR, int NT_, int G_>
__global__ void __launch_bounds__(AlsWs<TR, NT_, G_>::NTHREADS, 1) als_ws_kernel(const AlsArgs a) {
  using W = AlsWs<TR, NT_, G_>;
  constexpr int RP = W::RP, NS = W::NS, NRT = W::NRT, G = W::G, NT = W::NT, KT = W::KT, QP = W::QP, MR = W::MR, NACC = W::NACC,
                NGRP = W::NGRP;
  extern __shared__ uint8_t sm_raw[];
  uint8_t *smb = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smb);
  float *bpart = reinterpret_cast<float *>(smb + W::off_bpart);
  int2 *opinfo = reinterpret_cast<int2 *>(smb + W::off_opinfo);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smb + W::off_bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + W::off_tmem);
  // mbarriers: meta_full[NT][MR] | meta_empty[NT][MR] | op_full[NS] | op_empty[NS] | acc_full | rhs_full | acc_empty [NACC]
  const uint32_t bar_meta = smem_u32(bars), bar_mempty = bar_meta + NT * MR * 8,
                 bar_opf = bar_mempty + NT * MR * 8, bar_op = bar_opf + NS * 8, bar_accf = bar_op + NS * 8,
                 bar_rhs = bar_accf + NACC * 8, bar_acce = bar_rhs + NACC * 8;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_rows = (int)blockIdx.x < a.nseg ? (a.nseg - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int nq = a.ld >> 2;

  if (tid == 0) {
    for (int i = 0; i < NT * MR; i++) {
      mbar_init(bar_meta + i * 8, 33);                // 32 cp.async arrivals + the descriptor's release
      mbar_init(bar_mempty + i * 8, kWsTeam / 32);    // one arrival per converter warp
    }
    for (int i = 0; i < NS; i++) {
      mbar_init(bar_opf + i * 8, kWsTeam / 32);
      mbar_init(bar_op + i * 8, 1);
    }
    for (int i = 0; i < NACC; i++) {
      mbar_init(bar_accf + i * 8, 1);
      mbar_init(bar_rhs + i * 8, NT * kWsTeam / 32);
      mbar_init(bar_acce + i * 8, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(W::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // A cursor walks the tiles of this CTA's rows (row i = segment seg0 + blockIdx.x + i * gridDim.x); the descriptor of
  // the next row is fetched when a row is entered, so that the dependent loads are off the critical path.  Only the
  // scheduler warps and the tensor-core thread walk cursors; the converters are told what a tile is.
  struct Cursor { int i, t, ntiles, start, len, nstart, nlen; };
  auto fetch_next = [&](Cursor &c) {
    c.nstart = c.nlen = 0;
    if (c.i + 1 < n_rows) {
      const int seg = a.seg0 + (int)blockIdx.x + (c.i + 1) * (int)gridDim.x;
      c.nstart = __ldg(a.seg_start + seg);
      c.nlen = __ldg(a.seg_len + seg);
    }
  };
  auto init_cursor = [&](Cursor &c) {
    c.i = 0; c.t = 0; c.ntiles = 1; c.start = c.len = 0;
    if (n_rows > 0) {
      const int seg = a.seg0 + (int)blockIdx.x;
      c.start = __ldg(a.seg_start + seg);
      c.len = __ldg(a.seg_len + seg);
      c.ntiles = (c.len + KT - 1) / KT;
    }
    fetch_next(c);
  };
  auto advance = [&](Cursor &c) {
    if (c.i >= n_rows) return;
    if (++c.t == c.ntiles) {
      c.i++; c.t = 0;
      c.start = c.nstart; c.len = c.nlen;
      c.ntiles = (c.len + KT - 1) / KT;
      fetch_next(c);
    }
  };

  if (tid < W::SCHED0) {
    // ================================= converter teams =================================
    // Team tile x (global tile x * NT + team): wait for its gathered rows, accumulate the right-hand side, split every
    // value into tf32 big + small and store both transposed into the tile's operand stage.
    const int team = tid / kWsTeam, tt = tid % kWsTeam;
    const int *meta_item = reinterpret_cast<const int *>(smb + W::off_meta) + team * MR * 32;
    const float *meta_rate = reinterpret_cast<const float *>(smb + W::off_meta + NT * MR * 32 * 4) + team * MR * 32;
    const int2 *desc = reinterpret_cast<const int2 *>(smb + W::off_desc) + team * MR;
    const uint32_t raw0 = sbase + W::off_raw + team * NRT * W::RAW_BYTES;
    const uint32_t bmeta = bar_meta + team * MR * 8, bmempty = bar_mempty + team * MR * 8;
    constexpr int D = NRT - 1;     // team tiles whose factor rows are in flight
    const int q = tt % QP;         // this thread's 16-byte unit of the factor row
    const int grp = tt / QP;       // its K group: ratings 4 grp .. 4 grp + 3 of the tile
    const int j0 = grp * 4;
    bool issue_done = false;
    // the factor rows of team tile y, 16 bytes per cp.async: thread (grp, q) copies unit q of ratings 4 grp .. 4 grp + 3 —
    // exactly what it converts later, so a raw stage needs no barrier at all (its own cp.async group tells the thread
    // when its 64 bytes have landed) and the warps of a team run independently; a warp instruction still covers two
    // whole factor rows (coalesced).  One commit group per tile, also after the stream has ended (empty groups keep
    // the wait_group arithmetic fixed).
    auto issue_rows = [&](int y) {
      if (!issue_done) {
        const int st = y % NRT, ms = y % MR;
        mbar_wait(bmeta + ms * 8, (uint32_t)(y / MR) & 1u);
        const int n = desc[ms].x & 0xFF;  // ratings of the tile, 0 = the stream has ended
        if (n == 0) issue_done = true;
        else if (q < nq) {
          const uint32_t dst0 = raw0 + st * W::RAW_BYTES + q * 16;
          const int4 it4 = *reinterpret_cast<const int4 *>(meta_item + ms * 32 + j0);
          const float4 rt4 = *reinterpret_cast<const float4 *>(meta_rate + ms * 32 + j0);
          const int its[4] = {it4.x, it4.y, it4.z, it4.w};
          const float rts[4] = {rt4.x, rt4.y, rt4.z, rt4.w};
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (j0 + k < n && rts[k] > 0.f)  // rating > 0 filter (modelMF.cpp:819); beyond the row's end the ring is stale
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (j0 + k) * W::RAW_ROW),
                           "l"(a.Fin + (size_t)its[k] * a.ld + q * 4)
                           : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // right-hand side b = sum r f: a thread keeps the partial sum of its unit over its ratings of the row and writes it
    // into bpart[row mod NACC] when it leaves the row.  The slot is picked by the CLASS of the team's tiles in this row
    // (tile index within the row mod NT — one class per team and row), not by the team: which team gets a row's first
    // tile depends on how many tiles the CTA has seen before, and a row must not be summed in a different order
    // because of that (row-sharded runs reproduce a single engine bit for bit).  The solver group adds the classes the
    // row has; a team arrives on rhs_full for every row, also for rows it had no tile of (fixed arrival count).
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    int brow = -1, bcls = 0, flushed = 0;  // row / class bacc belongs to; rows [0, flushed) have been flushed by this thread
    auto flush_to = [&](int upto) {
      for (; flushed < upto; flushed++) {
        const int acc = flushed % NACC;
        if (flushed >= NACC) mbar_wait(bar_acce + acc * 8, (uint32_t)(flushed / NACC - 1) & 1u);  // bpart[acc] consumed
        if (flushed == brow) {
          *reinterpret_cast<float4 *>(bpart + ((size_t)(acc * NT + bcls) * NGRP + grp) * RP + q * 4) = bacc;
          bacc = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_rhs + acc * 8);
      }
    };
    for (int k = 0; k < D; k++) issue_rows(k);
    for (int x = 0;; x++) {
      const int g = x * NT + team;  // global tile index: operand stage and order of the tensor-core instructions
      issue_rows(x + D);  // into the stage this thread read tile x - 1 from
      const int st = x % NRT, ms = x % MR;
      const int2 d = desc[ms];  // published before meta_full[ms], which issue_rows(x) has waited for
      const int n = d.x & 0xFF;
      if (n == 0) break;
      asm volatile("cp.async.wait_group %0;" ::"n"(D) : "memory");  // this thread's copies of tile x have landed
      if (d.y != brow) {
        flush_to(d.y);
        brow = d.y;
        bcls = (d.x >> kWsClassShift) & 3;
      }
      const float4 *raw4 = reinterpret_cast<const float4 *>(smb + W::off_raw + (team * NRT + st) * W::RAW_BYTES);
      const float4 rt4 = *reinterpret_cast<const float4 *>(meta_rate + ms * 32 + j0);
      const float rts[4] = {rt4.x, rt4.y, rt4.z, rt4.w};
      float4 f[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool on = j0 + i < n && rts[i] > 0.f && q < nq;  // beyond the row's end the ring holds stale values
        const float rt = on ? rts[i] : 0.f;
        f[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (on) f[i] = raw4[(j0 + i) * QP + q];
        bacc.x = fmaf(rt, f[i].x, bacc.x);
        bacc.y = fmaf(rt, f[i].y, bacc.y);
        bacc.z = fmaf(rt, f[i].z, bacc.z);
        bacc.w = fmaf(rt, f[i].w, bacc.w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bmempty + ms * 8);  // the scheduler may refill this ring slot
      const int os = g % NS;
      if (g >= NS) mbar_wait(bar_op + os * 8, (uint32_t)(g / NS - 1) & 1u);  // the MMAs that read this stage are done
      const uint32_t op0 = sbase + W::off_op + os * W::OP_BYTES;
      {
        const float fe[4][4] = {{f[0].x, f[1].x, f[2].x, f[3].x}, {f[0].y, f[1].y, f[2].y, f[3].y},
                                {f[0].z, f[1].z, f[2].z, f[3].z}, {f[0].w, f[1].w, f[2].w, f[3].w}};
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int mdim = 4 * q + e;
          float bg[4], sl[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            bg[i] = round_tf32_fast(fe[e][i]);
            sl[i] = fe[e][i] - bg[i];  // exact in fp32; the tensor core reads its upper 19 bits (truncation: 2^-21 of f)
          }
          const uint32_t off = (uint32_t)grp * W::LBO + (uint32_t)(mdim >> 3) * kWsSbo + (uint32_t)(mdim & 7) * 16;
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(op0 + off), "f"(bg[0]), "f"(bg[1]), "f"(bg[2]), "f"(bg[3]) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(op0 + off + (RP / 8) * kWsSbo), "f"(sl[0]), "f"(sl[1]), "f"(sl[2]), "f"(sl[3]) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the tensor core
      if (tt == 0) opinfo[os] = d;  // what the tensor-core thread needs to know about this stage's tile
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_opf + os * 8);
    }
    flush_to(n_rows);
  } else if (tid < W::MMA0) {
    // ================================= scheduler warps (one per team) =================================
    // Walks the team's tiles MR ahead of the converters: lane l requests (item, rating) of rating l of the tile with
    // 4-byte cp.async copies straight into the ring, lane 0 publishes the descriptor {ratings in the tile, row}.
    const int team = (tid - W::SCHED0) >> 5;
    int *meta_item = reinterpret_cast<int *>(smb + W::off_meta) + team * MR * 32;
    float *meta_rate = reinterpret_cast<float *>(smb + W::off_meta + NT * MR * 32 * 4) + team * MR * 32;
    int2 *desc = reinterpret_cast<int2 *>(smb + W::off_desc) + team * MR;
    const uint32_t bmeta = bar_meta + team * MR * 8, bmempty = bar_mempty + team * MR * 8;
    Cursor c;
    init_cursor(c);
    for (int k = 0; k < team; k++) advance(c);  // the team's first tile is global tile `team`
    for (int x = 0;; x++) {
      const int slot = x % MR;
      if (x >= MR) mbar_wait(bmempty + slot * 8, (uint32_t)(x / MR - 1) & 1u);  // the converters have left this slot
      const bool live = c.i < n_rows;
      if (live) {
        const int j = c.t * KT + lane;
        if (lane < KT && j < c.len) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(meta_item + slot * 32 + lane)), "l"(a.ind + c.start + j) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(meta_rate + slot * 32 + lane)), "l"(a.val + c.start + j) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bmeta + slot * 8) : "memory");
      if (lane == 0) {
        desc[slot] = live ? make_int2(min(KT, c.len - c.t * KT) | (c.t == 0 ? kWsFirst : 0) | (c.t == c.ntiles - 1 ? kWsLast : 0) |
                                          ((c.t % NT) << kWsClassShift), c.i)
                          : make_int2(0, n_rows);
        mbar_arrive(bmeta + slot * 8);  // release: the descriptor is visible to whoever sees the phase complete
      }
      if (!live) break;
#pragma unroll
      for (int k = 0; k < NT; k++) advance(c);
    }
  } else if (tid < W::SOLVER0) {
    // ================================= tensor-core issue (one warp, one elected lane) =================================
    // Takes the operand stages in global tile order.  The whole warp runs the loop converged and every operand of the
    // instructions is warp-uniform by construction (stage, accumulator and first-tile flag are loop-carried, the
    // last-tile flag comes out of a vote): the compiler keeps descriptors and addresses in uniform registers.  With the
    // loop under `if (lane == 0)` each tcgen05.mma cost ~20 instructions of register -> uniform-register moves and
    // the issuing thread, not the tensor core, bounded the kernel (1180 cycles per 32-rating tile, tensor pipe 12 %
    // busy: profiles/r2_als.md).  All K steps of a stage are multiplied: the converters zero-fill short tiles.
    constexpr uint32_t NDIM = TR == 8 ? 256u : 64u;
    // instruction descriptor: D = F32, A = B = TF32, both K-major; RP = 128: M = 128, N = 256; RP = 64: M = 128, N = 64
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((NDIM >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc0 = umma_desc_kmajor(sbase + W::off_op, W::LBO, kWsSbo);
    int os = 0, acc = 0, gen = 0;  // operand stage, accumulator (= row mod NACC) and its use count (= row / NACC)
    uint32_t ph = 0;
    bool first = true;
    for (int rows_done = 0; rows_done < n_rows;) {
      mbar_wait(bar_opf + os * 8, ph);  // the tile's operand stage is written
      const bool last = __any_sync(0xFFFFFFFFu, (opinfo[os].x & kWsLast) != 0);
      if (first && gen > 0) mbar_wait(bar_acce + acc * 8, (uint32_t)(gen - 1) & 1u);  // accumulator drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one_sync()) {
        const uint32_t dt = tmem_base + (uint32_t)acc * W::ACC_COLS;
        const uint64_t descs = desc0 + (uint64_t)((uint32_t)os * (W::OP_BYTES >> 4));
#pragma unroll
        for (int k8 = 0; k8 < KT / 8; k8++) {
          const uint64_t dk = descs + (uint64_t)(k8 * ((2 * W::LBO) >> 4));
          umma_tf32(dt, dk, dk, idesc, (first && k8 == 0) ? 0u : 1u);
        }
        umma_commit(bar_op + os * 8);
        if (last) umma_commit(bar_accf + acc * 8);
      }
      __syncwarp();
      first = last;
      if (last) {
        rows_done++;
        if (++acc == NACC) { acc = 0; gen++; }
      }
      if (++os == NS) { os = 0; ph ^= 1u; }
    }
  } else {
    // ================================= solver groups =================================
    const int st = tid - W::SOLVER0, g = st / kWsGroupThreads, t = st % kWsGroupThreads;
    float *Gs = reinterpret_cast<float *>(smb + W::off_gs) + (size_t)g * W::GS_FLOATS;
    float *gbv = reinterpret_cast<float *>(smb + W::off_gbv) + g * RP;
    using K = CholScratch<TR>;
    const NamedBarrier group_bar{kWsGroupBar0 + g, kWsGroupThreads};
    int R = -1, C = -2;
    if (t < 136) {
      R = 0;
      while ((R + 1) * (R + 2) / 2 <= t) R++;
      C = t - R * (R + 1) / 2;
    }
    const int p = 32 * ((tid >> 5) & 3) + lane;  // TMEM lane this thread may read (its warp's quadrant)
    for (int i = g; i < n_rows; i += G) {
      const int acc = i % NACC;
      const uint32_t par = (uint32_t)(i / NACC) & 1u;
      const int seg = a.seg0 + (int)blockIdx.x + i * (int)gridDim.x;
      const int row = a.seg_row[seg], slot = a.seg_slot[seg];
      if (t < 128) {
        mbar_wait(bar_accf + acc * 8, par);
        mbar_wait(bar_rhs + acc * 8, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tad = tmem_base + (uint32_t)acc * W::ACC_COLS + ((uint32_t)(p & ~31) << 16);
        uint32_t v[32], x[32];
        if (TR == 8) {
          int ta, ea;
          tile_of<TR>(p, ta, ea);
          for (int c0 = 0; c0 < RP; c0 += 32) {  // G(a, b) = BB + BS for tile(a) <= tile(b)
            tmem_ld32(tad + c0, v);
            tmem_ld32(tad + RP + c0, x);
            tmem_ld_wait();
#pragma unroll
            for (int k4 = 0; k4 < 8; k4++) {
              int tb, eb;
              tile_of<TR>(c0 + 4 * k4, tb, eb);
              if (ta <= tb)
                *reinterpret_cast<float4 *>(Gs + tri_index(tb, ta) * W::TS + ea * TR + eb) =
                    make_float4(__uint_as_float(v[4 * k4]) + __uint_as_float(x[4 * k4]), __uint_as_float(v[4 * k4 + 1]) + __uint_as_float(x[4 * k4 + 1]),
                                __uint_as_float(v[4 * k4 + 2]) + __uint_as_float(x[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3]) + __uint_as_float(x[4 * k4 + 3]));
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(kWsDrainBar0 + g) : "memory");
          for (int c0 = 0; c0 < RP; c0 += 32) {  // + BS^T: G(c, a) += BS[a][c] for tile(c) <= tile(a)
            tmem_ld32(tad + RP + c0, x);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; k++) {
              int tc, ec;
              tile_of<TR>(c0 + k, tc, ec);
              if (tc <= ta) Gs[tri_index(ta, tc) * W::TS + ec * TR + ea] += __uint_as_float(x[k]);
            }
          }
        } else {
          // lanes 0..63: BB rows, lanes 64..127: SB rows (small big^T)
          const int arow = p & 63;
          int ta, ea;
          tile_of<TR>(arow, ta, ea);
          if (p < 64) {
            for (int c0 = 0; c0 < RP; c0 += 32) {
              tmem_ld32(tad + c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int k4 = 0; k4 < 8; k4++) {
                const int tb = (c0 + 4 * k4) / TR;
                if (ta <= tb)
                  *reinterpret_cast<float4 *>(Gs + tri_index(tb, ta) * W::TS + ea * TR) =
                      make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]), __uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3]));
              }
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(kWsDrainBar0 + g) : "memory");
          if (p >= 64) {
            for (int c0 = 0; c0 < RP; c0 += 32) {  // + SB: G(a, b) += SB[a][b]
              tmem_ld32(tad + c0, x);
              tmem_ld_wait();
#pragma unroll
              for (int k4 = 0; k4 < 8; k4++) {
                const int tb = (c0 + 4 * k4) / TR;
                if (ta <= tb) {
                  float4 *dst = reinterpret_cast<float4 *>(Gs + tri_index(tb, ta) * W::TS + ea * TR);
                  float4 o = *dst;
                  o.x += __uint_as_float(x[4 * k4]); o.y += __uint_as_float(x[4 * k4 + 1]);
                  o.z += __uint_as_float(x[4 * k4 + 2]); o.w += __uint_as_float(x[4 * k4 + 3]);
                  *dst = o;
                }
              }
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(kWsDrainBar0 + g) : "memory");
          if (p >= 64) {
            for (int c0 = 0; c0 < RP; c0 += 32) {  // + SB^T: G(c, a) += SB[a][c]
              tmem_ld32(tad + c0, x);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 32; k++) {
                int tc, ec;
                tile_of<TR>(c0 + k, tc, ec);
                if (tc <= ta) Gs[tri_index(ta, tc) * W::TS + ec * TR + ea] += __uint_as_float(x[k]);
              }
            }
          }
        }
        if (t < RP) {  // right-hand side: the converters' partial sums, one class per tile-in-row index mod NT
          const int ncls = min(NT, (a.seg_len[seg] + KT - 1) / KT);
          float sacc = 0.f;
          for (int k = 0; k < ncls * NGRP; k++) sacc += bpart[(size_t)(acc * NT * NGRP + k) * RP + t];
          gbv[t] = sacc;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_acce + acc * 8);  // accumulator and bpart[acc] may be reused
      }
      group_bar();
      float tile[TR][TR];
#pragma unroll
      for (int ii = 0; ii < TR; ii++) {
        float rowv[TR];
        if (t < 136) lds_vec<TR>(Gs + tri_index(R, C) * W::TS + ii * TR, rowv);
#pragma unroll
        for (int jj = 0; jj < TR; jj++) tile[ii][jj] = t < 136 ? rowv[jj] : 0.f;
      }
      group_bar();  // every tile is in registers: the buffer becomes the Cholesky scratch
      if (slot >= 0) {
        // segment of a split row: add the partial Gram / right-hand side into the row's workspace (als_solve_ws_kernel)
        float *wsp = a.ws + (size_t)slot * (RP * RP + RP);
        if (t < 136) {
#pragma unroll
          for (int ii = 0; ii < TR; ii++)
#pragma unroll
            for (int jj = 0; jj < TR; jj++) atomicAdd(wsp + tile_idx<TR>(C, ii) * RP + tile_idx<TR>(R, jj), tile[ii][jj]);
        }
        if (t < RP) atomicAdd(wsp + RP * RP + t, gbv[t]);
        group_bar();
        continue;
      }
      if (t < 136) add_reg_diag<TR>(tile, R, C, a.rank, a.reg);
      chol_solve_core<TR>(tile, Gs + K::off_pan, Gs + K::off_dg, Gs + K::off_y, gbv, R, C, group_bar);
      store_solution(a, row, t, gbv);
      group_bar();  // the solution has been read: the buffers may be overwritten by the next row
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(W::TMEM_COLS) : "memory");
}

template <int TR, int NT, int G>
static int launch_als_ws_cfg(mfb_engine *e, AlsArgs b, int n_primal) {
  using W = AlsWs<TR, NT, G>;
  MFB_CUDA(cudaFuncSetAttribute((als_ws_kernel<TR, NT, G>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::bytes));
  b.seg0 = 0;
  b.nseg = n_primal;
  const int grid = std::min(e->sm_count, n_primal);
  MFB_LAUNCH((als_ws_kernel<TR, NT, G>), grid, W::NTHREADS, W::bytes, e->stream, b);
  return 0;
}

// long_rows: few long rows (the item side of a ratings matrix): conversion-bound -> more converter teams, one solver group;
// otherwise many short rows: solve-bound -> more solver groups
template <int TR>
static int launch_als_ws(mfb_engine *e, const AlsArgs &b, int n_primal, bool long_rows) {
  if (n_primal <= 0) return 0;
  if (TR == 8) return long_rows ? launch_als_ws_cfg<8, 3, 1>(e, b, n_primal) : launch_als_ws_cfg<8, 1, 2>(e, b, n_primal);
  return long_rows ? launch_als_ws_cfg<4, 4, 1>(e, b, n_primal) : launch_als_ws_cfg<4, 2, 4>(e, b, n_primal);
}

template <int TRD>
static int launch_dual(mfb_engine *e, AlsArgs b, int seg0, int count) {
  if (count <= 0) return 0;
  using S = AlsDualSmem<TRD>;
  MFB_CUDA(cudaFuncSetAttribute(als_dual_kernel<TRD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
  b.seg0 = seg0;
  MFB_LAUNCH((als_dual_kernel<TRD>), count, 256, S::bytes, e->stream, b);
  return 0;
}

template <int TR>
static int launch_als(mfb_engine *e, const AlsArgs &a, const SegPlan &sp) {
  using S = AlsSmem<TR>;
  MFB_CUDA(cudaFuncSetAttribute(als_gram_solve_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
  MFB_CUDA(cudaFuncSetAttribute(als_solve_ws_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
  if (sp.n_multi > 0) {
    size_t need = sizeof(float) * (size_t)sp.n_multi * (S::RP * S::RP + S::RP);
    if (need > e->als_ws_bytes) {
      if (e->als_ws) MFB_CUDA(dev_free(e->als_ws));
      e->als_ws = nullptr;
      e->als_ws_bytes = 0;
      MFB_CUDA(dev_alloc(&e->als_ws, need));
      e->als_ws_bytes = need;
    }
    MFB_CUDA(cudaMemsetAsync(e->als_ws, 0, need, e->stream));
  }
  AlsArgs b = a;
  b.ws = e->als_ws;
  // segments are sorted longest first: [0, n_primal) go through the rank x rank normal equations, the
  // shorter ones through the dual system whose padded size is at most half of that
  const int n64 = sp.n_longer[0], n32 = sp.n_longer[1], n16 = sp.n_longer[2];
  int n_primal = sp.n_seg;
  bool n_primal_done = false;
  if (e->opt_als_dual) n_primal = TR == 8 ? n64 : TR == 4 ? n32 : TR == 2 ? n16 : sp.n_seg;
  if constexpr (TR == 4) {
    if (e->opt_als_tensor_cores == 1) {
      // rank 33 .. 64: Gram records on MN-major tensor-core operands + batched warp-per-matrix solve (csrc/als_mn.cu);
      // split rows are summed into their own record there, the short rows below go through the dual kernels as before
      MFB_TRY(als_mn_half_step(e, b, sp, n_primal));
      if (n_primal < sp.n_seg) {
        MFB_TRY(launch_dual<2>(e, b, n32, n16 - n32));
        MFB_TRY(launch_dual<1>(e, b, n16, sp.n_seg - n16));
      }
      return 0;
    }
  }
  if constexpr (TR == 8 || TR == 4) {
    if (e->opt_als_tensor_cores == 1 || (TR == 4 && e->opt_als_tensor_cores == 3)) {  // 3: the round-2 converter kernel at rank <= 64
      // mean length of the rows this launch takes: the plan is sorted longest first and [0, n_primal) are primal
      // the split is picked from the mean row length of the WHOLE side, not of this rank's shard: every rank of a
      // row-sharded run then sums a row's right-hand side in the same order as a single engine would (bit-identical rows)
      const int64_t side_rows = a.ind == e->mat[MFB_TRAIN].rowind ? e->n_users : e->n_items;
      const bool long_rows = e->opt_als_ws_split ? e->opt_als_ws_split == 2 : e->mat[MFB_TRAIN].nnz / std::max<int64_t>(side_rows, 1) >= 1024;
      MFB_TRY(launch_als_ws<TR>(e, b, n_primal, long_rows));
      n_primal_done = true;
    }
  }
  if (n_primal_done) {
  } else if (TR == 8 && e->opt_als_tensor_cores) {  // als_tensor_cores = 2: the round-1 kernel (one CTA per row)
    MFB_CUDA(cudaFuncSetAttribute(als_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AlsTcSmem::bytes));
    if (n_primal > 0) MFB_LAUNCH(als_gram_tc_kernel, n_primal, 256, AlsTcSmem::bytes, e->stream, b);
  } else if (n_primal > 0) {
    MFB_LAUNCH((als_gram_solve_kernel<TR>), n_primal, 256, S::bytes, e->stream, b);
  }
  if (n_primal < sp.n_seg) {
    if (TR == 8) MFB_TRY(launch_dual<4>(e, b, n64, n32 - n64));
    if (TR >= 4) MFB_TRY(launch_dual<2>(e, b, n32, n16 - n32));
    if (TR >= 2) MFB_TRY(launch_dual<1>(e, b, n16, sp.n_seg - n16));
  }
  if (sp.n_multi > 0) MFB_LAUNCH((als_solve_ws_kernel<TR>), sp.n_multi, 256, S::bytes, e->stream, b);
  return 0;
}

// Diagnostics: Gram matrix and right-hand side of ONE row through the production kernels (split-row
// path: partial sums land in the workspace).  out = [RP*RP + RP] floats, RP = padded rank.
int als_debug_gram(mfb_engine *e, int side, int32_t row, float *out, int32_t *rp_out) {
  DevCsr &m = e->mat[MFB_TRAIN];
  const int64_t *ptr = side == MFB_USER ? m.rowptr : m.colptr;
  int64_t be[2];
  MFB_CUDA(cudaMemcpy(be, ptr + row, sizeof(int64_t) * 2, cudaMemcpyDeviceToHost));
  const int r = e->rank;
  const int RP = r <= 16 ? 16 : r <= 32 ? 32 : r <= 64 ? 64 : 128;
  *rp_out = RP;
  int32_t h[5] = {row, (int32_t)be[0], (int32_t)(be[1] - be[0]), 0, row};
  int32_t *d;
  MFB_CUDA(dev_alloc(&d, sizeof(h)));
  MFB_CUDA(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
  float *ws;
  const size_t wsb = sizeof(float) * ((size_t)RP * RP + RP);
  MFB_CUDA(dev_alloc(&ws, wsb));
  MFB_CUDA(cudaMemset(ws, 0, wsb));
  AlsArgs a;
  a.Fin = side == MFB_USER ? e->V : e->U;
  a.Fout = side == MFB_USER ? e->U : e->V;
  a.ld = e->ld; a.rank = r;
  a.ind = side == MFB_USER ? m.rowind : m.colind;
  a.val = side == MFB_USER ? m.rowval : m.colval;
  a.seg_row = d; a.seg_start = d + 1; a.seg_len = d + 2; a.seg_slot = d + 3; a.multi_row = d + 4;
  a.ws = ws; a.reg = 0.f; a.n_peer = 0; a.seg0 = 0; a.nseg = 1;
  if (RP == 128 && e->opt_als_tensor_cores) {
    MFB_CUDA(cudaFuncSetAttribute(als_gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AlsTcSmem::bytes));
    MFB_LAUNCH(als_gram_tc_kernel, 1, 256, AlsTcSmem::bytes, e->stream, a);
  } else if (RP == 128) {
    MFB_CUDA(cudaFuncSetAttribute(als_gram_solve_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AlsSmem<8>::bytes));
    MFB_LAUNCH((als_gram_solve_kernel<8>), 1, 256, AlsSmem<8>::bytes, e->stream, a);
  } else if (RP == 64) {
    MFB_CUDA(cudaFuncSetAttribute(als_gram_solve_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AlsSmem<4>::bytes));
    MFB_LAUNCH((als_gram_solve_kernel<4>), 1, 256, AlsSmem<4>::bytes, e->stream, a);
  } else {
    return fail("als_debug_gram: rank <= 32 not supported", __FILE__, __LINE__);
  }
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  MFB_CUDA(cudaMemcpy(out, ws, wsb, cudaMemcpyDeviceToHost));
  dev_free(ws);
  dev_free(d);
  return 0;
}

int als_half_step_launch(mfb_engine *e, int side, float reg) {
  DevCsr &m = e->mat[MFB_TRAIN];
  SegPlan &sp = side == MFB_USER ? m.als_rows : m.als_cols;
  if (!sp.built) {
    if (side == MFB_USER)
      MFB_TRY(build_seg_plan(e, m.rowptr, e->n_users, e->bad_user, e->row_begin[MFB_USER], e->row_end[MFB_USER],
                             e->opt_als_chunk, &sp));
    else
      MFB_TRY(build_seg_plan(e, m.colptr, e->n_items, e->bad_item, e->row_begin[MFB_ITEM], e->row_end[MFB_ITEM],
                             e->opt_als_chunk, &sp));
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
  a.n_peer = 0;
  a.seg0 = 0;
  a.nseg = 0;
  const Comm &c = e->comm;
  if (c.connected)
    for (int p = 0; p < c.world; p++)
      if (p != c.rank) a.Fpeer[a.n_peer++] = side == MFB_USER ? c.U[p] : c.V[p];
  const int r = e->rank;
  if (r <= 16) MFB_TRY(launch_als<1>(e, a, sp));
  else if (r <= 32) MFB_TRY(launch_als<2>(e, a, sp));
  else if (r <= 64) MFB_TRY(launch_als<4>(e, a, sp));
  else MFB_TRY(launch_als<8>(e, a, sp));
  return comm_barrier_launch(e);  // every rank's rows have landed before the next half-step reads them
}

}  // namespace mfb
