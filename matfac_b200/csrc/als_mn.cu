// ALS half-step at padded rank 64 (rank 33 .. 64) in two launches (replaces the OpenMP row loops of ModelMF::trainALS,
// modelMF.cpp:806-841 users, :845-880 items, for this rank class):
//
//   als_split_kernel    the opposite side's factor rows split once per half-step into tf32 "big" + fp32 remainder
//                       "small": Fs[n + 1][128] = [big(64) | small(64)], row n all zero (ratings that are filtered out
//                       or beyond a row's end gather it).
//   als_gram_mn_kernel  persistent, warp-specialised: producer warps copy the rated rows of a 32-rating tile with 16-byte
//                       cp.async straight into the MN-major operand layout of tcgen05.mma kind::tf32 (no register pass,
//                       no conversion: a factor row IS a K-row of the operand); one warp issues the MMAs
//                       D[128 x 64] = [big; small] big^T (3xTF32: G = BB + SB + SB^T) and D2[128 x 16] = [big; small] [r_big r_small]^T
//                       (the right-hand side b = sum r f through the tensor core as well); a drain team assembles the
//                       lower triangle of G and b in shared memory and stores one 8960-byte record per row with a bulk
//                       copy (split rows: reductions into the row's record).
//   als_chol64_kernel   batched solve, one warp per matrix: the record arrives by cp.async.bulk (double buffered), a
//                       blocked left-looking Cholesky (4 columns at a time, 128-bit shared loads, the 4 x 4 diagonal
//                       block factored redundantly by every lane after ten shuffles) and both substitutions run without
//                       any CTA-level barrier; the solution goes to this side's factor matrix (and to the peers').
//
// MN-major tf32 operands need the shared-memory layout type SWIZZLE_128B_BASE32B (tools/mn_probe.cu on this pod: every
// other layout type returns zeros for kind::tf32 with a_major = MN): element (m, k) of an operand lives at
//   (m / 32) * LBO + (k / 4) * SBO + swz((k % 4) * 128 + (m % 32) * 4),  swz(x) = x ^ (((x >> 7) & 3) << 5)
// i.e. 32 dims of one rating are 128 contiguous bytes whose 32-byte pieces are XOR-ed with k % 4.
#include <algorithm>

#include "engine.h"

namespace mfb {
namespace {

// Record of one row: the lower triangle of G row by row, then the right-hand side.  Row r holds r / 4 + 1 units of 16 bytes and
// starts at the first unit at or after the end of row r - 1 whose index mod 8 differs from those of the earlier rows of its
// group of eight rows (r / 8): eight neighbouring rows then sit in eight different 16-byte bank groups at every unit index, so
// a quarter warp's LDS.128 / STS.128 over eight rows is conflict free (the packed triangle cost the solver 47 % of its
// shared-memory wavefronts in bank conflicts; profiles/r2_als_mn.md).  594 units instead of 544.
constexpr int kRecG = 2376;            // floats of the triangle
constexpr int kRecFloats = kRecG + 64; // + right-hand side
constexpr uint32_t kRecBytes = kRecFloats * 4;  // 9760
__constant__ int c_row_off[64] = {0, 4, 8, 12, 16, 24, 52, 60, 68, 80, 92, 104, 116, 140, 160, 184, 200, 220, 240, 260, 280, 308, 332, 384, 408, 436, 464, 492, 520, 572, 608, 644, 676, 712, 748, 784, 820, 860, 920, 960, 1000, 1044, 1088, 1132, 1176, 1232, 1284, 1340, 1388, 1440, 1492, 1544, 1596, 1656, 1712, 1796, 1852, 1912, 1972, 2032, 2092, 2176, 2244, 2312};  // float offset of row r
[[maybe_unused]] static const int h_row_off[64] = {0, 4, 8, 12, 16, 24, 52, 60, 68, 80, 92, 104, 116, 140, 160, 184, 200, 220, 240, 260, 280, 308, 332, 384, 408, 436, 464, 492, 520, 572, 608, 644, 676, 712, 748, 784, 820, 860, 920, 960, 1000, 1044, 1088, 1132, 1176, 1232, 1284, 1340, 1388, 1440, 1492, 1544, 1596, 1656, 1712, 1796, 1852, 1912, 1972, 2032, 2092, 2176, 2244, 2312};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "MNW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MND_%=;\n\t"
      "bra MNW_%=;\n\t"
      "MND_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// the same with a suspend-time hint: a warp that expects to wait long is parked by the hardware instead of spinning
// through the issue slots of the warps that have work
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "MPW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra MPD_%=;\n\t"
      "bra MPW_%=;\n\t"
      "MPD_%=:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(ns)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// batched Cholesky solve, one warp per 64 x 64 system
__device__ __forceinline__ int chol_row_off(int r) { return c_row_off[r]; }
__device__ __forceinline__ float rsqrt_nr(float d) {
  const float r = d > 0.f ? rsqrtf(d) : 0.f;
  return r * fmaf(-0.5f * d * r, r, 1.5f);  // one Newton step on the hardware approximation
}
__device__ __forceinline__ float sub_dot4(const float4 a, const float4 b, float acc) {
  acc = fmaf(-a.x, b.x, acc);
  acc = fmaf(-a.y, b.y, acc);
  acc = fmaf(-a.z, b.z, acc);
  return fmaf(-a.w, b.w, acc);
}
__device__ __forceinline__ float pick4(const float4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

struct Blk4 {  // Cholesky factor of a 4 x 4 diagonal block: strictly lower entries and the inverse diagonal
  float l10, l20, l21, l30, l31, l32, i0, i1, i2, i3;
};
__device__ __forceinline__ float4 blk_solve(const Blk4 &b, const float4 t) {  // x L^T = t
  float4 x;
  x.x = t.x * b.i0;
  x.y = fmaf(-x.x, b.l10, t.y) * b.i1;
  x.z = fmaf(-x.y, b.l21, fmaf(-x.x, b.l20, t.z)) * b.i2;
  x.w = fmaf(-x.z, b.l32, fmaf(-x.y, b.l31, fmaf(-x.x, b.l30, t.w))) * b.i3;
  return x;
}

// M: the record in shared memory (compact lower triangle of G, then b); dinv: 64 floats of scratch.  Lane l owns rows l
// and l + 32.  Solves (G + reg I) x = b (padded dims >= rank: identity rows) and returns x[lane], x[lane + 32].
__device__ __forceinline__ void warp_chol64(float *M, float *dinv, int lane, int rank, float reg, float &x0, float &x1) {
  const int r0 = lane, r1 = lane + 32;
  float *row0 = M + chol_row_off(r0), *row1 = M + chol_row_off(r1);
  row0[r0] = r0 < rank ? row0[r0] + reg : 1.0f;
  row1[r1] = r1 < rank ? row1[r1] + reg : 1.0f;
  __syncwarp();
  float4 *R0 = reinterpret_cast<float4 *>(row0), *R1 = reinterpret_cast<float4 *>(row1);
#pragma unroll 1
  for (int J = 0; J < 16; J++) {
    // rows 4J .. 4J + 3 hold J + 1 units of 16 bytes each
    const float4 *B0 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J)), *B1 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 1)),
                 *B2 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 2)), *B3 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 3));
    const bool two = J < 8;  // rows 0 .. 31 are still below or inside the block column
    float4 t1 = R1[J], t0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (two) {
      t0 = R0[J];
#pragma unroll 1
      for (int k = 0; k < J; k++) {
        const float4 b0 = B0[k], b1 = B1[k], b2 = B2[k], b3 = B3[k], a0 = R0[k], a1 = R1[k];
        t0.x = sub_dot4(a0, b0, t0.x); t0.y = sub_dot4(a0, b1, t0.y); t0.z = sub_dot4(a0, b2, t0.z); t0.w = sub_dot4(a0, b3, t0.w);
        t1.x = sub_dot4(a1, b0, t1.x); t1.y = sub_dot4(a1, b1, t1.y); t1.z = sub_dot4(a1, b2, t1.z); t1.w = sub_dot4(a1, b3, t1.w);
      }
    } else {
#pragma unroll 2
      for (int k = 0; k < J; k++) {
        const float4 b0 = B0[k], b1 = B1[k], b2 = B2[k], b3 = B3[k], a1 = R1[k];
        t1.x = sub_dot4(a1, b0, t1.x); t1.y = sub_dot4(a1, b1, t1.y); t1.z = sub_dot4(a1, b2, t1.z); t1.w = sub_dot4(a1, b3, t1.w);
      }
    }
    // the diagonal block sits in the lanes that own rows 4J .. 4J + 3
    const int src = (4 * J) & 31;
    const float4 ts = two ? t0 : t1;
    const float d00 = __shfl_sync(0xffffffffu, ts.x, src);
    const float d10 = __shfl_sync(0xffffffffu, ts.x, src + 1), d11 = __shfl_sync(0xffffffffu, ts.y, src + 1);
    const float d20 = __shfl_sync(0xffffffffu, ts.x, src + 2), d21 = __shfl_sync(0xffffffffu, ts.y, src + 2),
                d22 = __shfl_sync(0xffffffffu, ts.z, src + 2);
    const float d30 = __shfl_sync(0xffffffffu, ts.x, src + 3), d31 = __shfl_sync(0xffffffffu, ts.y, src + 3),
                d32 = __shfl_sync(0xffffffffu, ts.z, src + 3), d33 = __shfl_sync(0xffffffffu, ts.w, src + 3);
    Blk4 b;
    b.i0 = rsqrt_nr(d00);
    b.l10 = d10 * b.i0; b.l20 = d20 * b.i0; b.l30 = d30 * b.i0;
    b.i1 = rsqrt_nr(fmaf(-b.l10, b.l10, d11));
    b.l21 = fmaf(-b.l20, b.l10, d21) * b.i1;
    b.l31 = fmaf(-b.l30, b.l10, d31) * b.i1;
    b.i2 = rsqrt_nr(fmaf(-b.l21, b.l21, fmaf(-b.l20, b.l20, d22)));
    b.l32 = fmaf(-b.l31, b.l21, fmaf(-b.l30, b.l20, d32)) * b.i2;
    b.i3 = rsqrt_nr(fmaf(-b.l32, b.l32, fmaf(-b.l31, b.l31, fmaf(-b.l30, b.l30, d33))));
    // rows inside the block get their own entries of L from the same formula (entries right of the diagonal are padding)
    if (r1 >= 4 * J) R1[J] = blk_solve(b, t1);
    if (two && r0 >= 4 * J) R0[J] = blk_solve(b, t0);
    if (lane == 0) reinterpret_cast<float4 *>(dinv)[J] = make_float4(b.i0, b.i1, b.i2, b.i3);
    __syncwarp();
  }
  // L y = b, block column by block column; y0 / y1 carry the running right-hand side of rows lane / lane + 32
  const float *bv = M + kRecG;
  float y0 = bv[r0], y1 = bv[r1];
#pragma unroll 1
  for (int J = 0; J < 16; J++) {
    const float4 q1 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 1))[J], q2 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 2))[J],
                 q3 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 3))[J];  // unit J of rows 4J + 1 .. 4J + 3
    const float4 di = reinterpret_cast<const float4 *>(dinv)[J];
    const bool two = J < 8;
    const int src = (4 * J) & 31;
    const float sel = two ? y0 : y1;
    const float c0 = __shfl_sync(0xffffffffu, sel, src), c1 = __shfl_sync(0xffffffffu, sel, src + 1),
                c2 = __shfl_sync(0xffffffffu, sel, src + 2), c3 = __shfl_sync(0xffffffffu, sel, src + 3);
    float4 z;
    z.x = c0 * di.x;
    z.y = fmaf(-q1.x, z.x, c1) * di.y;
    z.z = fmaf(-q2.y, z.y, fmaf(-q2.x, z.x, c2)) * di.z;
    z.w = fmaf(-q3.z, z.z, fmaf(-q3.y, z.y, fmaf(-q3.x, z.x, c3))) * di.w;
    if (r1 > 4 * J + 3) y1 = sub_dot4(R1[J], z, y1);
    if (two && r0 > 4 * J + 3) y0 = sub_dot4(R0[J], z, y0);
    if (two) { if ((lane >> 2) == J) y0 = pick4(z, lane & 3); }
    else { if ((lane >> 2) == J - 8) y1 = pick4(z, lane & 3); }
  }
  // L^T x = y, from the last block column backwards
#pragma unroll 1
  for (int J = 15; J >= 0; J--) {
    const float *S0 = M + chol_row_off(4 * J), *S1 = M + chol_row_off(4 * J + 1), *S2 = M + chol_row_off(4 * J + 2), *S3 = M + chol_row_off(4 * J + 3);
    const float4 q1 = reinterpret_cast<const float4 *>(S1)[J], q2 = reinterpret_cast<const float4 *>(S2)[J],
                 q3 = reinterpret_cast<const float4 *>(S3)[J];
    const float4 di = reinterpret_cast<const float4 *>(dinv)[J];
    const bool two = J < 8;
    const int src = (4 * J) & 31;
    const float sel = two ? y0 : y1;
    const float c0 = __shfl_sync(0xffffffffu, sel, src), c1 = __shfl_sync(0xffffffffu, sel, src + 1),
                c2 = __shfl_sync(0xffffffffu, sel, src + 2), c3 = __shfl_sync(0xffffffffu, sel, src + 3);
    float4 x;
    x.w = c3 * di.w;
    x.z = fmaf(-q3.z, x.w, c2) * di.z;
    x.y = fmaf(-q3.y, x.w, fmaf(-q2.y, x.z, c1)) * di.y;
    x.x = fmaf(-q3.x, x.w, fmaf(-q2.x, x.z, fmaf(-q1.x, x.y, c0))) * di.x;
    if (r0 < 4 * J) y0 = fmaf(-S3[r0], x.w, fmaf(-S2[r0], x.z, fmaf(-S1[r0], x.y, fmaf(-S0[r0], x.x, y0))));
    if (J > 8 && r1 < 4 * J) y1 = fmaf(-S3[r1], x.w, fmaf(-S2[r1], x.z, fmaf(-S1[r1], x.y, fmaf(-S0[r1], x.x, y1))));
    if (two) { if ((lane >> 2) == J) y0 = pick4(x, lane & 3); }
    else { if ((lane >> 2) == J - 8) y1 = pick4(x, lane & 3); }
  }
  x0 = y0;
  x1 = y1;
}

template <int P0>
__device__ __forceinline__ void hw_kloop(const float4 *B0, const float4 *B1, const float4 *B2, const float4 *B3, float *const (&rowp)[4], int J,
                                         float4 (&t)[4]) {
#pragma unroll 2
  for (int k = 0; k < J; k++) {
    const float4 b0 = B0[k], b1 = B1[k], b2 = B2[k], b3 = B3[k];
#pragma unroll
    for (int p = P0; p < 4; p++) {
      const float4 a = reinterpret_cast<const float4 *>(rowp[p])[k];
      t[p].x = sub_dot4(a, b0, t[p].x); t[p].y = sub_dot4(a, b1, t[p].y);
      t[p].z = sub_dot4(a, b2, t[p].z); t[p].w = sub_dot4(a, b3, t[p].w);
    }
  }
}

// Two matrices per warp: lanes 0-15 solve the record at M0, lanes 16-31 the one at M1 (the caller passes each lane ITS
// matrix); lane q of a half owns rows q, q + 16, q + 32, q + 48.  The four rows of a block column that every lane needs
// ("broadcast" loads: a warp-wide LDS.128 costs four shared-memory wavefronts whatever its addresses) now serve two
// matrices and up to 64 FMAs per lane, and the granularity of idle rows drops from 32 to 16.  Shuffles stay inside a half.
__device__ __forceinline__ void halfwarp_chol64(float *M, float *dinv, int q, int rank, float reg, float (&x)[4]) {
  float *rowp[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int r = q + 16 * p;
    rowp[p] = M + chol_row_off(r);
    rowp[p][r] = r < rank ? rowp[p][r] + reg : 1.0f;
  }
  __syncwarp();
#pragma unroll 1
  for (int J = 0; J < 16; J++) {
    const float4 *B0 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J)), *B1 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 1)),
                 *B2 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 2)), *B3 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 3));
    const int p0 = J >> 2;  // passes below p0 hold finished rows only
    float4 t[4];
#pragma unroll
    for (int p = 0; p < 4; p++) t[p] = p >= p0 ? reinterpret_cast<const float4 *>(rowp[p])[J] : make_float4(0.f, 0.f, 0.f, 0.f);
    switch (p0) {  // warp-uniform: the loop body of each case is straight-line code
      case 0: hw_kloop<0>(B0, B1, B2, B3, rowp, J, t); break;
      case 1: hw_kloop<1>(B0, B1, B2, B3, rowp, J, t); break;
      case 2: hw_kloop<2>(B0, B1, B2, B3, rowp, J, t); break;
      default: hw_kloop<3>(B0, B1, B2, B3, rowp, J, t); break;
    }
    const int src = (4 * J) & 15;
    const float4 ts = p0 == 0 ? t[0] : p0 == 1 ? t[1] : p0 == 2 ? t[2] : t[3];
    const float d00 = __shfl_sync(0xffffffffu, ts.x, src, 16);
    const float d10 = __shfl_sync(0xffffffffu, ts.x, src + 1, 16), d11 = __shfl_sync(0xffffffffu, ts.y, src + 1, 16);
    const float d20 = __shfl_sync(0xffffffffu, ts.x, src + 2, 16), d21 = __shfl_sync(0xffffffffu, ts.y, src + 2, 16),
                d22 = __shfl_sync(0xffffffffu, ts.z, src + 2, 16);
    const float d30 = __shfl_sync(0xffffffffu, ts.x, src + 3, 16), d31 = __shfl_sync(0xffffffffu, ts.y, src + 3, 16),
                d32 = __shfl_sync(0xffffffffu, ts.z, src + 3, 16), d33 = __shfl_sync(0xffffffffu, ts.w, src + 3, 16);
    Blk4 b;
    b.i0 = rsqrt_nr(d00);
    b.l10 = d10 * b.i0; b.l20 = d20 * b.i0; b.l30 = d30 * b.i0;
    b.i1 = rsqrt_nr(fmaf(-b.l10, b.l10, d11));
    b.l21 = fmaf(-b.l20, b.l10, d21) * b.i1;
    b.l31 = fmaf(-b.l30, b.l10, d31) * b.i1;
    b.i2 = rsqrt_nr(fmaf(-b.l21, b.l21, fmaf(-b.l20, b.l20, d22)));
    b.l32 = fmaf(-b.l31, b.l21, fmaf(-b.l30, b.l20, d32)) * b.i2;
    b.i3 = rsqrt_nr(fmaf(-b.l32, b.l32, fmaf(-b.l31, b.l31, fmaf(-b.l30, b.l30, d33))));
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p >= p0 && q + 16 * p >= 4 * J) reinterpret_cast<float4 *>(rowp[p])[J] = blk_solve(b, t[p]);
    if (q == 0) reinterpret_cast<float4 *>(dinv)[J] = make_float4(b.i0, b.i1, b.i2, b.i3);
    __syncwarp();
  }
  const float *bv = M + kRecG;
  float y[4];
#pragma unroll
  for (int p = 0; p < 4; p++) y[p] = bv[q + 16 * p];
#pragma unroll 1
  for (int J = 0; J < 16; J++) {
    const float4 q1 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 1))[J], q2 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 2))[J],
                 q3 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J + 3))[J];
    const float4 di = reinterpret_cast<const float4 *>(dinv)[J];
    const int p0 = J >> 2, src = (4 * J) & 15;
    const float sel = p0 == 0 ? y[0] : p0 == 1 ? y[1] : p0 == 2 ? y[2] : y[3];
    const float c0 = __shfl_sync(0xffffffffu, sel, src, 16), c1 = __shfl_sync(0xffffffffu, sel, src + 1, 16),
                c2 = __shfl_sync(0xffffffffu, sel, src + 2, 16), c3 = __shfl_sync(0xffffffffu, sel, src + 3, 16);
    float4 z;
    z.x = c0 * di.x;
    z.y = fmaf(-q1.x, z.x, c1) * di.y;
    z.z = fmaf(-q2.y, z.y, fmaf(-q2.x, z.x, c2)) * di.z;
    z.w = fmaf(-q3.z, z.z, fmaf(-q3.y, z.y, fmaf(-q3.x, z.x, c3))) * di.w;
    const bool owner = (q >> 2) == (J & 3);
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p >= p0) {
        if (q + 16 * p > 4 * J + 3) y[p] = sub_dot4(reinterpret_cast<const float4 *>(rowp[p])[J], z, y[p]);
        if (p == p0 && owner) y[p] = pick4(z, q & 3);
      }
  }
#pragma unroll 1
  for (int J = 15; J >= 0; J--) {
    const float *S0 = M + chol_row_off(4 * J), *S1 = M + chol_row_off(4 * J + 1), *S2 = M + chol_row_off(4 * J + 2), *S3 = M + chol_row_off(4 * J + 3);
    const float4 q1 = reinterpret_cast<const float4 *>(S1)[J], q2 = reinterpret_cast<const float4 *>(S2)[J],
                 q3 = reinterpret_cast<const float4 *>(S3)[J];
    const float4 di = reinterpret_cast<const float4 *>(dinv)[J];
    const int p0 = J >> 2, src = (4 * J) & 15;
    const float sel = p0 == 0 ? y[0] : p0 == 1 ? y[1] : p0 == 2 ? y[2] : y[3];
    const float c0 = __shfl_sync(0xffffffffu, sel, src, 16), c1 = __shfl_sync(0xffffffffu, sel, src + 1, 16),
                c2 = __shfl_sync(0xffffffffu, sel, src + 2, 16), c3 = __shfl_sync(0xffffffffu, sel, src + 3, 16);
    float4 v;
    v.w = c3 * di.w;
    v.z = fmaf(-q3.z, v.w, c2) * di.z;
    v.y = fmaf(-q3.y, v.w, fmaf(-q2.y, v.z, c1)) * di.y;
    v.x = fmaf(-q3.x, v.w, fmaf(-q2.x, v.z, fmaf(-q1.x, v.y, c0))) * di.x;
    const bool owner = (q >> 2) == (J & 3);
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p <= p0) {
        const int r = q + 16 * p;
        if (r < 4 * J) y[p] = fmaf(-S3[r], v.w, fmaf(-S2[r], v.z, fmaf(-S1[r], v.y, fmaf(-S0[r], v.x, y[p]))));
        if (p == p0 && owner) y[p] = pick4(v, q & 3);
      }
  }
#pragma unroll
  for (int p = 0; p < 4; p++) x[p] = y[p];
}

// NW warps per CTA, NB record buffers per warp (2: the next record is fetched while the current one is solved; 1: more
// warps fit the shared memory and hide each other's fetches)
template <int NW, int NB>
struct CholCfg {
  static constexpr uint32_t warp_bytes = NB * kRecBytes + 256;  // record buffers + the inverse diagonal
  static constexpr uint32_t smem = NW * warp_bytes + NW * 16 + 128;
  static_assert(smem <= 227 * 1024, "shared memory budget");
};

struct CholArgs {
  const float *rec;     // [n_jobs][kRecFloats]
  int n_jobs, n_seg;    // jobs [0, n_seg): records of single-segment rows (segments of split rows are skipped);
                        // jobs [n_seg, n_jobs): the accumulated records of the split rows
  const int32_t *seg_row, *seg_slot, *multi_row;
  float *Fout;
  int ld, rank;
  float reg;
  float *Fpeer[kMaxRanks - 1];
  int n_peer;
};

template <int NW, int NB>
__global__ void __launch_bounds__(NW * 32, 1) als_chol64_kernel(const CholArgs a) {
  using C = CholCfg<NW, NB>;
  extern __shared__ __align__(128) uint8_t smraw[];
  // offset arithmetic on the shared array itself (not through an integer): the compiler keeps the shared address space
  // and emits LDS / STS — through a uintptr_t round trip every access of the solver became a generic LD.E / ST.E
  uint8_t *smb = smraw + ((128u - (smem_u32(smraw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *mine = smb + (size_t)warp * C::warp_bytes;
  float *dinv = reinterpret_cast<float *>(mine + NB * kRecBytes);
  const uint32_t bars = smem_u32(smb + (size_t)NW * C::warp_bytes) + warp * 16;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int stride = (int)gridDim.x * NW;
  auto next_job = [&](int j) {  // first job >= j (on this warp's stride) that has a matrix to solve
    while (j < a.n_seg && a.seg_slot[j] >= 0) j += stride;
    return j;
  };
  auto fetch = [&](int j, int b) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was last touched by ordinary loads / stores
      mbar_arrive_expect_tx(bars + b * 8, kRecBytes);
      bulk_g2s(smem_u32(mine + b * kRecBytes), a.rec + (size_t)j * kRecFloats, kRecBytes, bars + b * 8);
    }
  };
  int j = next_job((int)blockIdx.x * NW + warp);
  if (j < a.n_jobs) fetch(j, 0);
  uint32_t ph = 0u;  // bit b: phase parity of buffer b's barrier
  for (int b = 0; j < a.n_jobs; b = (b + 1) % NB) {
    const int jn = next_job(j + stride);
    __syncwarp();
    if (NB > 1 && jn < a.n_jobs) fetch(jn, b ^ 1);
    mbar_wait(bars + b * 8, (ph >> b) & 1u);
    ph ^= 1u << b;
    float x0, x1;
    warp_chol64(reinterpret_cast<float *>(mine + b * kRecBytes), dinv, lane, a.rank, a.reg, x0, x1);
    const int row = j < a.n_seg ? a.seg_row[j] : a.multi_row[j - a.n_seg];
    if (lane >= a.rank) x0 = 0.f;
    if (lane + 32 >= a.rank) x1 = 0.f;
    float *o = a.Fout + (size_t)row * a.ld;
    if (lane < a.ld) o[lane] = x0;
    if (lane + 32 < a.ld) o[lane + 32] = x1;
    for (int p = 0; p < a.n_peer; p++) {
      float *op = a.Fpeer[p] + (size_t)row * a.ld;
      if (lane < a.ld) op[lane] = x0;
      if (lane + 32 < a.ld) op[lane + 32] = x1;
    }
    j = jn;
    if (NB == 1 && j < a.n_jobs) {
      __syncwarp();
      fetch(j, 0);
    }
  }
}

// Two jobs per warp and round (halfwarp_chol64): jobs 2i and 2i + 1 of the warp's stride share a warp; a half whose
// job has nothing to solve (segment of a split row, or past the end) idles through the round on whatever its buffer holds.
template <int NW>
struct Chol2Cfg {
  static constexpr uint32_t warp_bytes = 2 * kRecBytes + 512;  // two records + two inverse diagonals
  static constexpr uint32_t smem = NW * warp_bytes + NW * 16 + 128;
  static_assert(smem <= 227 * 1024, "shared memory budget");
};
template <int NW>
__global__ void __launch_bounds__(NW * 32, 1) als_chol64x2_kernel(const CholArgs a) {
  using C = Chol2Cfg<NW>;
  extern __shared__ __align__(128) uint8_t smraw[];
  uint8_t *smb = smraw + ((128u - (smem_u32(smraw) & 127u)) & 127u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4, q = lane & 15;
  uint8_t *mine = smb + (size_t)warp * C::warp_bytes;
  float *M = reinterpret_cast<float *>(mine + h * kRecBytes);
  float *dinv = reinterpret_cast<float *>(mine + 2 * kRecBytes + h * 256);
  const uint32_t bar = smem_u32(smb + (size_t)NW * C::warp_bytes) + warp * 16;
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int n_pairs = (a.n_jobs + 1) / 2;
  uint32_t ph = 0u;
  for (int pr = (int)blockIdx.x * NW + warp; pr < n_pairs; pr += (int)gridDim.x * NW) {
    const int j = 2 * pr + h;
    const bool valid = j < a.n_jobs && !(j < a.n_seg && a.seg_slot[j] >= 0);
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const bool v0 = (vm & 1u) != 0, v1 = (vm & 0x10000u) != 0;
    if (!v0 && !v1) continue;
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_expect_tx(bar, (v0 ? kRecBytes : 0u) + (v1 ? kRecBytes : 0u));
      if (v0) bulk_g2s(smem_u32(mine), a.rec + (size_t)(2 * pr) * kRecFloats, kRecBytes, bar);
      if (v1) bulk_g2s(smem_u32(mine + kRecBytes), a.rec + (size_t)(2 * pr + 1) * kRecFloats, kRecBytes, bar);
    }
    mbar_wait(bar, ph);
    ph ^= 1u;
    float x[4];
    halfwarp_chol64(M, dinv, q, a.rank, a.reg, x);
    if (valid) {
      const int row = j < a.n_seg ? a.seg_row[j] : a.multi_row[j - a.n_seg];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const int d = q + 16 * p;
        const float v = d < a.rank ? x[p] : 0.f;
        if (d < a.ld) {
          a.Fout[(size_t)row * a.ld + d] = v;
          for (int pp = 0; pp < a.n_peer; pp++) a.Fpeer[pp][(size_t)row * a.ld + d] = v;
        }
      }
    }
    __syncwarp();
  }
}

template <int NW>
int launch_chol64x2_cfg(mfb_engine *e, const CholArgs &c) {
  using C = Chol2Cfg<NW>;
  MFB_CUDA(cudaFuncSetAttribute((als_chol64x2_kernel<NW>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
  const int n_pairs = (c.n_jobs + 1) / 2;
  const int grid = std::min(e->sm_count, (n_pairs + NW - 1) / NW);
  MFB_LAUNCH((als_chol64x2_kernel<NW>), grid, NW * 32, C::smem, e->stream, c);
  return 0;
}

template <int NW, int NB>
int launch_chol64_cfg(mfb_engine *e, const CholArgs &c) {
  using C = CholCfg<NW, NB>;
  MFB_CUDA(cudaFuncSetAttribute((als_chol64_kernel<NW, NB>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem));
  const int grid = std::min(e->sm_count, (c.n_jobs + NW - 1) / NW);
  MFB_LAUNCH((als_chol64_kernel<NW, NB>), grid, NW * 32, C::smem, e->stream, c);
  return 0;
}

int launch_chol64(mfb_engine *e, const CholArgs &c) {
  if (c.n_jobs <= 0) return 0;
  switch (e->opt_als_chol_warps) {
    case 11: return launch_chol64_cfg<11, 2>(e, c);
    case 16: return launch_chol64_cfg<16, 1>(e, c);
    case 20: return launch_chol64_cfg<20, 1>(e, c);
    case 22: return launch_chol64_cfg<22, 1>(e, c);
    case 208: return launch_chol64x2_cfg<8>(e, c);
    default: return launch_chol64x2_cfg<11>(e, c);  // 211: two matrices per warp, 11 warps
  }
}

// ---------------------------------------------------------------------------------------------
// tf32 split of the gathered side's factors
__device__ __forceinline__ float round_tf32_fast(float x) {  // nearest, ties away, finite values (two integer instructions)
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__global__ void __launch_bounds__(256) als_split_kernel(const float *__restrict__ F, int ld, int n, float *__restrict__ Fs) {
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t row = t >> 4;
  const int q = (int)(t & 15);
  if (row > n) return;
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n && 4 * q < ld) f = *reinterpret_cast<const float4 *>(F + row * ld + 4 * q);
  const float4 b = make_float4(round_tf32_fast(f.x), round_tf32_fast(f.y), round_tf32_fast(f.z), round_tf32_fast(f.w));
  const float4 sm = make_float4(f.x - b.x, f.y - b.y, f.z - b.z, f.w - b.w);  // exact in fp32; the tensor core reads its upper 19 bits
  *reinterpret_cast<float4 *>(Fs + row * 128 + 4 * q) = b;
  *reinterpret_cast<float4 *>(Fs + row * 128 + 64 + 4 * q) = sm;
}

// ---------------------------------------------------------------------------------------------
// Gram records on MN-major operands
constexpr int kGnKT = 32;                     // ratings per tile
constexpr uint32_t kGnBlock = 4096;           // one operand block: 32 dims x 32 ratings (LBO: stride between 32-dim blocks)
constexpr uint32_t kGnSbo = 512;              // 4 ratings x 128 bytes (stride between groups of 4 ratings)
constexpr uint32_t kGnStage = 5 * kGnBlock;   // small 0-31 | small 32-63 | big 0-31 | big 32-63 | ratings (n = 0: r_big, n = 1: r_small).
                                              // A = all four blocks (M = 128: rows 0-63 small, 64-127 big); B starts at the big
                                              // blocks and runs into the ratings block (N = 80): ONE tcgen05.mma per 8 ratings gives
                                              // small big^T, big big^T and both products of the right-hand side — the kernel is bound
                                              // by shared-memory bandwidth (profiles/r2_als_mn.md) and a second MMA for the
                                              // right-hand side re-read the 4 KB of A
constexpr int kGnStages = 9;
constexpr int kGnProducers = 128;             // 4 producer warps, each owns every 4th tile (8 warps: same kernel time, the MMAs set the pace)
constexpr int kGnAhead = 4;                   // a warp's ring holds the (item, rating) pairs of its next four tiles; the pairs of the fifth
                                              // are on their way in registers
constexpr int kGnAcc = 4;                     // accumulators in TMEM: row i -> i mod 4
constexpr int kGnAccCols = 128;               // D in columns [0, 64), D2 in [64, 80)
constexpr int kGnDrainTeams = 2;              // row i is drained by team i mod 2 (4 warps, one per TMEM lane quadrant)
constexpr int kGnMma0 = kGnProducers, kGnDrain0 = kGnMma0 + 32;
constexpr int kGnThreads = kGnDrain0 + kGnDrainTeams * 128;
constexpr int kGnFirst = 1, kGnLast = 2;

struct GramSmem {
  static constexpr uint32_t off_stage = 0;
  static constexpr uint32_t off_rec = off_stage + kGnStages * kGnStage;               // [teams] record staging
  static constexpr uint32_t off_tmp = off_rec + kGnDrainTeams * kRecBytes;        // [teams][64] small x r_big
  static constexpr uint32_t off_meta = off_tmp + kGnDrainTeams * 64 * 4;              // [producer warps][kGnAhead][32] {item, rating}: warp-private rings
  static constexpr uint32_t off_tflags = off_meta + (kGnProducers / 32) * kGnAhead * 256;  // [producer warps][kGnAhead] flags of the tiles a cursor has visited
  static constexpr uint32_t off_info = off_tflags + (kGnProducers / 32) * kGnAhead * 4;                    // [stages] tile flags
  static constexpr uint32_t off_bars = (off_info + kGnStages * 4 + 15u) & ~15u;                     // op_full | op_empty | acc_full | acc_empty
  static constexpr int n_bars = 2 * kGnStages + 2 * kGnAcc;
  static constexpr uint32_t off_tmem = off_bars + n_bars * 8;
  static constexpr size_t bytes = off_tmem + 16 + 1024;
  static_assert(bytes <= 227 * 1024, "shared memory budget");
  static_assert(off_bars % 8 == 0 && off_meta % 16 == 0 && off_rec % 16 == 0, "alignment of mbarriers / rings / bulk-copy source");
  static_assert(kGnAhead >= 2, "pipeline depths");
};

struct GramArgs {
  const float *Fs;      // [zero_row + 1][128] split factors of the gathered side
  int zero_row;
  const int32_t *ind;
  const float *val;
  const int32_t *seg_start, *seg_len, *seg_slot;
  int nseg;             // segments [0, nseg) of the plan (longest first); CTA b takes b, b + grid, ...
  int debug_mode;       // timing experiments (results are wrong): 1 = every rating gathers row (item & 1023): L2-hot rows;
                        // 2 = no factor-row copies at all (the MMAs read stale stages); 3 = copies, but no MMAs (commits only);
                        // 4 = unused; 5 / 6 = MMAs with N = 16 / M = 64
  float *rec;           // [nseg + split rows][kRecFloats]
};

__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr) {  // SWIZZLE_128B_BASE32B, version 1
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((kGnBlock >> 4) & 0x3FFF) << 16;  // leading byte offset: next block of 32 dims
  d |= (uint64_t)((kGnSbo >> 4) & 0x3FFF) << 32;    // stride byte offset: next group of 4 ratings
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
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
__device__ __forceinline__ void umma_tf32_acc(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {  // D += A B
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t &a, uint32_t &b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(kGnThreads, 1) als_gram_mn_kernel(const GramArgs a) {
  using S = GramSmem;
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t *smb = sm_raw + ((1024u - (smem_u32(sm_raw) & 1023u)) & 1023u);  // stays a shared-space pointer (LDS / STS)
  const uint32_t sbase = smem_u32(smb);
  int *opinfo = reinterpret_cast<int *>(smb + S::off_info);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + S::off_tmem);
  const uint32_t bar_opf = sbase + S::off_bars, bar_ope = bar_opf + kGnStages * 8, bar_accf = bar_ope + kGnStages * 8,
                 bar_acce = bar_accf + kGnAcc * 8;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_rows = (int)blockIdx.x < a.nseg ? (a.nseg - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int i = 0; i < kGnStages; i++) {
      mbar_init(bar_opf + i * 8, 33);  // the owning producer warp: 32 cp.async completions + lane 0
      mbar_init(bar_ope + i * 8, 1);
    }
    for (int i = 0; i < kGnAcc; i++) {
      mbar_init(bar_accf + i * 8, 1);
      mbar_init(bar_acce + i * 8, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the ratings block of every stage is zero except the two floats a producer writes per rating
  for (int i = tid; i < kGnStages * (int)(kGnBlock / 16); i += kGnThreads) {
    const int stg = i / (int)(kGnBlock / 16), u = i % (int)(kGnBlock / 16);
    *reinterpret_cast<float4 *>(smb + S::off_stage + stg * kGnStage + 4 * kGnBlock + u * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kGnAcc * kGnAccCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (tid < kGnProducers) {
    // ================================= producers =================================
    // Every producer WARP owns whole tiles: warp w copies global tiles w, w + NPW, ... — the warps run independently of
    // each other.  (With all warps co-operating on every tile the kernel ran at the latency of ONE warp's ~170-instruction
    // loop per tile: removing the copies or the MMAs altogether did not change its time, profiles/r2_als_mn.md.)
    // Lane l = (o8 = l / 8, c = l % 8): chunk c of each of the four 128-byte blocks of the split rows of ratings o8, o8 + 4, ...,
    // o8 + 28 of the tile — eight neighbouring lanes copy one contiguous 128-byte block of one rating, its swizzled destination
    // covers all 32 banks.  A warp-private cursor walks the warp's tiles kGnAhead of them ahead of the copies and requests
    // their (item, rating) pairs (lane l: pair l of the tile) into a warp-private ring.
    constexpr int NPW = kGnProducers / 32;
    const int w = tid >> 5;
    struct Cursor { int i, t, ntiles, start, len, s1, l1, s2, l2; };
    auto load_seg = [&](int i, int &st_, int &ln_) {
      st_ = ln_ = 0;
      if (i < n_rows) {
        const int seg = (int)blockIdx.x + i * (int)gridDim.x;
        st_ = __ldg(a.seg_start + seg);
        ln_ = __ldg(a.seg_len + seg);
      }
    };
    auto advance_n = [&](Cursor &c, int n) {  // n tiles further (whole rows at a time)
      while (n > 0 && c.i < n_rows) {
        const int rem = c.ntiles - c.t;
        if (n < rem) { c.t += n; return; }
        n -= rem;
        c.i++; c.t = 0;
        c.start = c.s1; c.len = c.l1;
        c.ntiles = (c.len + kGnKT - 1) / kGnKT;
        c.s1 = c.s2; c.l1 = c.l2;
        load_seg(c.i + 2, c.s2, c.l2);
      }
    };
    const int o8 = lane >> 3, cg = lane & 7;
    const uint32_t base_off = (uint32_t)o8 * 128u + ((((uint32_t)cg >> 1) ^ (uint32_t)o8) << 5) + ((uint32_t)cg & 1u) * 16u;
    const uint32_t roff = 4u * kGnBlock + (uint32_t)o8 * 128u + ((uint32_t)o8 << 5);  // n = 0, 1 of rating o8 (+ 4 m: + m * kGnSbo)
    const uint32_t meta0 = sbase + S::off_meta + (uint32_t)w * (kGnAhead * 256u);     // [kGnAhead][32] {item, rating}
    const int2 *meta_ptr = reinterpret_cast<const int2 *>(smb + S::off_meta) + w * (kGnAhead * 32);
    int *tflags = reinterpret_cast<int *>(smb + S::off_tflags) + w * kGnAhead;
    int ny = -1;  // number of this warp's tiles, known once the cursor has run off the end
    // lane l loads pair l of the cursor's tile with ordinary loads (NOT cp.async: the stage-full arrival below waits for every
    // cp.async the lane has in flight, and these come from DRAM) and stores it into the ring one iteration later
    auto load_pair = [&](const Cursor &c, int y, int &it, int &rt, int &fl) {  // y = index among this warp's tiles
      const bool alive = c.i < n_rows;
      if (!alive && ny < 0) ny = y;
      it = 0; rt = 0;  // rating 0: filtered out
      const int jj = c.t * kGnKT + lane;
      if (alive && jj < c.len) {
        it = __ldg(a.ind + c.start + jj);
        rt = __float_as_int(__ldg(a.val + c.start + jj));
      }
      fl = alive ? ((c.t == 0 ? kGnFirst : 0) | (c.t == c.ntiles - 1 ? kGnLast : 0)) : 0;
    };
    auto store_pair = [&](int y, int it, int rt, int fl) {
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(meta0 + (uint32_t)(y % kGnAhead) * 256u + (uint32_t)lane * 8u), "r"(it), "r"(rt) : "memory");
      if (lane == 0) tflags[y % kGnAhead] = fl;
    };
    Cursor nxt;
    nxt.i = 0; nxt.t = 0;
    load_seg(0, nxt.start, nxt.len);
    nxt.ntiles = (nxt.len + kGnKT - 1) / kGnKT;
    load_seg(1, nxt.s1, nxt.l1);
    load_seg(2, nxt.s2, nxt.l2);
    advance_n(nxt, w);  // the warp's first tile is global tile w
    int p_it, p_rt, p_fl;  // the pair / flags of tile y + kGnAhead while tile y is copied
    for (int y = 0; y < kGnAhead; y++) {
      load_pair(nxt, y, p_it, p_rt, p_fl);
      store_pair(y, p_it, p_rt, p_fl);
      advance_n(nxt, NPW);
    }
    load_pair(nxt, kGnAhead, p_it, p_rt, p_fl);
    __syncwarp();
    for (int y = 0;; y++) {
      const bool live = ny < 0 || y < ny;
      if (!live) break;
      const int g = w + NPW * y;  // global tile: its operand stage and its place in the tensor-core warp's order
      const int stg = g % kGnStages;
      if (g >= kGnStages) mbar_wait(bar_ope + stg * 8, (uint32_t)(g / kGnStages - 1) & 1u);  // the MMAs that read this stage are done
      if (lane == 0) opinfo[stg] = tflags[y % kGnAhead];  // only now: the tensor-core warp has read the previous tile's flags
      const uint32_t dst0 = sbase + S::off_stage + (uint32_t)stg * kGnStage;
      const int2 *mp = meta_ptr + (y % kGnAhead) * 32 + o8;
#pragma unroll
      for (int m = 0; m < 8; m++) {
        const int2 mt = mp[4 * m];  // rating 4 m + o8 of the tile
        const float rt = __int_as_float(mt.y);
        const bool on = rt > 0.f;  // rating > 0 filter (modelMF.cpp:819); beyond the row's end the ring holds 0
        const int row = on ? (a.debug_mode == 1 ? (mt.x & 1023) : mt.x) : a.zero_row;
        const float *src = a.Fs + (size_t)row * 128 + cg * 4;
        if (a.debug_mode != 2) {
#pragma unroll
          for (int i = 0; i < 4; i++)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + base_off + m * kGnSbo + ((i + 2) & 3) * kGnBlock), "l"(src + 32 * i) : "memory");  // big -> blocks 2, 3; small -> 0, 1
        }
        if (cg == 0) {
          const float rb = on ? round_tf32_fast(rt) : 0.f;
          const float rs = on ? rt - rb : 0.f;
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(dst0 + roff + m * kGnSbo), "f"(rb), "f"(rs) : "memory");
        }
      }
      // the stage is full when every lane's copies have landed (the hardware arrives for the lane then: the warp does not
      // wait for its data) and lane 0 has arrived for the ratings / flags the lanes stored themselves
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_opf + stg * 8) : "memory");
      __syncwarp();                        // every lane has read the ring slot that is refilled now, and stored its ratings
      if (lane == 0) mbar_arrive(bar_opf + stg * 8);
      store_pair(y + kGnAhead, p_it, p_rt, p_fl);            // loaded one iteration ago
      advance_n(nxt, NPW);
      load_pair(nxt, y + kGnAhead + 1, p_it, p_rt, p_fl);    // in flight during the next tile's copies
      __syncwarp();
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (tid < kGnDrain0) {
    // ================================= tensor-core issue =================================
    // instruction descriptor: D = F32, A = B = TF32, both MN-major; M = 128 ([small; big]), N = 80 ([big | r_big r_small 0 ..])
    // This warp's loop is the pace of the kernel once the producers run ahead (profiles/r2_als_mn.md: a handful of extra
    // instructions per MMA cost 10 %): the descriptors are {low word = address field that moves, high word = constant},
    // advanced by 32-bit adds; the first MMA of a tile carries the accumulate predicate, the other three are unconditional.
    constexpr uint32_t idesc_g = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((80u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc0 = umma_desc_mn(sbase + S::off_stage);
    const uint32_t desc_hi = (uint32_t)(desc0 >> 32), lo0 = (uint32_t)desc0;
    constexpr uint32_t kStep = (2 * kGnSbo) >> 4, kBoff = (2 * kGnBlock) >> 4, kStageStep = kGnStage >> 4;
    const bool no_mma = a.debug_mode == 3;
    // timing experiments: 5 = N = 16 instead of 80 (B and the math shrink, A stays), 6 = M = 64 (half of A)
    const uint32_t idesc_run = a.debug_mode == 5 ? ((idesc_g & ~(0x3Fu << 17)) | ((16u >> 3) << 17))
                             : a.debug_mode == 6 ? ((idesc_g & ~(0x1Fu << 24)) | ((64u >> 4) << 24)) : idesc_g;
    int os = 0, acc = 0, gen = 0;
    uint32_t ph = 0, lo = lo0;
    bool first = true;
    for (int rows_done = 0; rows_done < n_rows;) {
      mbar_wait(bar_opf + os * 8, ph);
      const bool last = __any_sync(0xFFFFFFFFu, (opinfo[os] & kGnLast) != 0);
      if (first && gen > 0) mbar_wait(bar_acce + acc * 8, (uint32_t)(gen - 1) & 1u);  // accumulator drained
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was written through the generic proxy (cp.async, st.shared)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one_sync()) {
        const uint32_t dt = tmem_base + (uint32_t)acc * kGnAccCols;
        if (!no_mma) {
          auto dsc = [&](uint32_t l) { return ((uint64_t)desc_hi << 32) | l; };
          umma_tf32(dt, dsc(lo), dsc(lo + kBoff), idesc_run, first ? 0u : 1u);
          umma_tf32_acc(dt, dsc(lo + kStep), dsc(lo + kStep + kBoff), idesc_run);
          umma_tf32_acc(dt, dsc(lo + 2 * kStep), dsc(lo + 2 * kStep + kBoff), idesc_run);
          umma_tf32_acc(dt, dsc(lo + 3 * kStep), dsc(lo + 3 * kStep + kBoff), idesc_run);
        }
        umma_commit(bar_ope + os * 8);
        if (last) umma_commit(bar_accf + acc * 8);
      }
      __syncwarp();
      first = last;
      if (last) {
        rows_done++;
        if (++acc == kGnAcc) { acc = 0; gen++; }
      }
      lo += kStageStep;
      if (++os == kGnStages) { os = 0; ph ^= 1u; lo = lo0; }
    }
  } else {
    // ================================= drain teams =================================
    const int dt = tid - kGnDrain0, team = dt >> 7, t = dt & 127;
    const int p = 32 * ((tid >> 5) & 3) + lane;  // TMEM lane this thread may read (its warp's quadrant)
    const int arow = p & 63;
    const bool bb = p >= 64;  // accumulator rows 64 .. 127 = big x [big | ratings], rows 0 .. 63 = small x [big | ratings]
    float *tmp = reinterpret_cast<float *>(smb + S::off_tmp) + team * 64;
    const int bar_id = 1 + team;
    auto team_bar = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
    int n = 0;
    for (int i = team; i < n_rows; i += kGnDrainTeams, n++) {
      const int acc = i % kGnAcc;
      const uint32_t par = (uint32_t)(i / kGnAcc) & 1u;
      const int seg = (int)blockIdx.x + i * (int)gridDim.x;
      const int slot = a.seg_slot[seg];
      float *Sr = reinterpret_cast<float *>(smb + S::off_rec + (uint32_t)team * kRecBytes);
      if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous row's store has read the buffer
      mbar_wait_parked(bar_accf + acc * 8, par, 20000u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tad = tmem_base + (uint32_t)acc * kGnAccCols + ((uint32_t)(p & ~31) << 16);
      uint32_t v[64], r0, r1;
      {
        uint32_t lo[32], hi[32];
        tmem_ld32(tad, lo);
        tmem_ld32(tad + 32, hi);
        tmem_ld2(tad + 64, r0, r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 32; k++) { v[k] = lo[k]; v[32 + k] = hi[k]; }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acce + acc * 8);  // everything of this accumulator is in registers
      team_bar();                        // the staging buffer is free (thread 0 has waited for its last store)
      float4 *rowp = reinterpret_cast<float4 *>(Sr + chol_row_off(arow));
      if (bb) {  // big big^T: units 0 .. arow / 4 of row arow
#pragma unroll
        for (int u = 0; u < 16; u++)
          if (4 * u <= arow)
            rowp[u] = make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
      } else {
        tmp[arow] = __uint_as_float(r0);  // small x r_big
      }
      team_bar();
      if (!bb) {  // + small big^T, lower part of row arow (the diagonal once here, once below)
#pragma unroll
        for (int u = 0; u < 16; u++)
          if (4 * u <= arow) {
            float4 o = rowp[u];
            o.x += __uint_as_float(v[4 * u]); o.y += __uint_as_float(v[4 * u + 1]);
            o.z += __uint_as_float(v[4 * u + 2]); o.w += __uint_as_float(v[4 * u + 3]);
            rowp[u] = o;
          }
      } else {
        Sr[kRecG + arow] = __uint_as_float(r0) + __uint_as_float(r1) + tmp[arow];  // big r_big + big r_small + small r_big
      }
      team_bar();
      if (!bb) {  // + (small big^T)^T: G(c, arow) += SB[arow][c] for c >= arow
#pragma unroll
        for (int c = 0; c < 64; c++)
          if (c >= arow) Sr[chol_row_off(c) + arow] += __uint_as_float(v[c]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      team_bar();
      if (slot < 0) {
        if (t == 0) {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.rec + (size_t)seg * kRecFloats),
                       "r"(smem_u32(Sr)), "r"(kRecBytes)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {  // segment of a split row: add into the row's record
        float *dst = a.rec + ((size_t)a.nseg + slot) * kRecFloats;
        for (int k = t; k < kRecFloats; k += 128) atomicAdd(dst + k, Sr[k]);
        if (t == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // keeps the wait_group arithmetic fixed
      }
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last stores have completed
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kGnAcc * kGnAccCols) : "memory");
}

}  // namespace

// Diagnostics: n records (compact lower triangle + right-hand side, host) through the batched solver; x_host = [n][64].
int als_debug_chol64(mfb_engine *e, int32_t n, const float *rec_host, float *x_host, int32_t rank, float reg) {
  float *rec = nullptr, *x = nullptr;
  int32_t *rows = nullptr;
  MFB_CUDA(dev_alloc(&rec, (size_t)n * kRecBytes));
  MFB_CUDA(dev_alloc(&x, sizeof(float) * 64 * (size_t)n));
  MFB_CUDA(dev_alloc(&rows, sizeof(int32_t) * 2 * (size_t)n));
  std::vector<int32_t> h(2 * (size_t)n);
  for (int i = 0; i < n; i++) { h[i] = i; h[n + i] = -1; }
  MFB_CUDA(cudaMemcpyAsync(rec, rec_host, (size_t)n * kRecBytes, cudaMemcpyHostToDevice, e->stream));
  MFB_CUDA(cudaMemcpyAsync(rows, h.data(), sizeof(int32_t) * 2 * (size_t)n, cudaMemcpyHostToDevice, e->stream));
  CholArgs c;
  c.rec = rec; c.n_jobs = n; c.n_seg = n; c.seg_row = rows; c.seg_slot = rows + n; c.multi_row = rows;
  c.Fout = x; c.ld = 64; c.rank = rank; c.reg = reg; c.n_peer = 0;
  MFB_TRY(launch_chol64(e, c));
  MFB_CUDA(cudaMemcpyAsync(x_host, x, sizeof(float) * 64 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  dev_free(rec); dev_free(x); dev_free(rows);
  return 0;
}

int als_mn_half_step(mfb_engine *e, const AlsArgs &a, const SegPlan &sp, int n_primal) {
  if (n_primal <= 0) return 0;
  cudaStream_t st = e->stream;
  const int n_in = a.Fin == e->V ? e->n_items : e->n_users;
  float *Fs = nullptr, *rec = nullptr;
  MFB_CUDA(dev_alloc(&Fs, sizeof(float) * 128 * ((size_t)n_in + 1)));
  const size_t n_rec = (size_t)n_primal + (size_t)sp.n_multi;
  MFB_CUDA(dev_alloc(&rec, n_rec * kRecBytes));
  if (sp.n_multi > 0) MFB_CUDA(cudaMemsetAsync(rec + (size_t)n_primal * kRecFloats, 0, (size_t)sp.n_multi * kRecBytes, st));
  {
    const int64_t total = ((int64_t)n_in + 1) * 16;  // one thread per 4 dims
    MFB_LAUNCH(als_split_kernel, (unsigned)((total + 255) / 256), 256, 0, st, a.Fin, a.ld, n_in, Fs);
  }
  GramArgs g;
  g.Fs = Fs; g.zero_row = n_in;
  g.ind = a.ind; g.val = a.val;
  g.seg_start = a.seg_start; g.seg_len = a.seg_len; g.seg_slot = a.seg_slot;
  g.nseg = n_primal;
  g.debug_mode = e->opt_als_debug;
  g.rec = rec;
  MFB_CUDA(cudaFuncSetAttribute(als_gram_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GramSmem::bytes));
  const int grid = std::min(e->sm_count, n_primal);
  MFB_LAUNCH(als_gram_mn_kernel, grid, kGnThreads, GramSmem::bytes, st, g);
  CholArgs c;
  c.rec = rec; c.n_jobs = (int)n_rec; c.n_seg = n_primal;
  c.seg_row = a.seg_row; c.seg_slot = a.seg_slot; c.multi_row = a.multi_row;
  c.Fout = a.Fout; c.ld = a.ld; c.rank = a.rank; c.reg = a.reg;
  c.n_peer = a.n_peer;
  for (int p = 0; p < a.n_peer; p++) c.Fpeer[p] = a.Fpeer[p];
  MFB_TRY(launch_chol64(e, c));
  dev_free(Fs);
  dev_free(rec);
  return 0;
}

}  // namespace mfb
