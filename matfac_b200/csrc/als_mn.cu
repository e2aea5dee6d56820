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

constexpr int kRecG = 2176;            // floats of the compact lower triangle: row r starts at chol_row_off(r), padded to 4
constexpr int kRecFloats = kRecG + 64; // + right-hand side
constexpr uint32_t kRecBytes = kRecFloats * 4;  // 8960

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
__host__ __device__ __forceinline__ int chol_row_off(int r) {
  const int m = r >> 2, s = r & 3;
  return 4 * (m + 1) * (2 * m + s);
}
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
  for (int J = 0; J < 16; J++) {
    // rows 4J .. 4J + 3 hold J + 1 units of 16 bytes each
    const float4 *B0 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J));
    const float4 *B1 = B0 + (J + 1), *B2 = B1 + (J + 1), *B3 = B2 + (J + 1);
    const bool two = J < 8;  // rows 0 .. 31 are still below or inside the block column
    float4 t1 = R1[J], t0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (two) {
      t0 = R0[J];
      for (int k = 0; k < J; k++) {
        const float4 b0 = B0[k], b1 = B1[k], b2 = B2[k], b3 = B3[k], a0 = R0[k], a1 = R1[k];
        t0.x = sub_dot4(a0, b0, t0.x); t0.y = sub_dot4(a0, b1, t0.y); t0.z = sub_dot4(a0, b2, t0.z); t0.w = sub_dot4(a0, b3, t0.w);
        t1.x = sub_dot4(a1, b0, t1.x); t1.y = sub_dot4(a1, b1, t1.y); t1.z = sub_dot4(a1, b2, t1.z); t1.w = sub_dot4(a1, b3, t1.w);
      }
    } else {
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
  for (int J = 0; J < 16; J++) {
    const float4 *B0 = reinterpret_cast<const float4 *>(M + chol_row_off(4 * J));
    const float4 q1 = B0[2 * J + 1], q2 = B0[3 * J + 2], q3 = B0[4 * J + 3];  // unit J of rows 4J + 1 .. 4J + 3
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
  for (int J = 15; J >= 0; J--) {
    const float *S0 = M + chol_row_off(4 * J);
    const float *S1 = S0 + 4 * (J + 1), *S2 = S1 + 4 * (J + 1), *S3 = S2 + 4 * (J + 1);
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

constexpr int kCholWarps = 12;
constexpr uint32_t kCholWarpBytes = 2 * kRecBytes + 256;  // two record buffers + the inverse diagonal
constexpr uint32_t kCholSmem = kCholWarps * kCholWarpBytes + kCholWarps * 16 + 128;

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

__global__ void __launch_bounds__(kCholWarps * 32, 1) als_chol64_kernel(const CholArgs a) {
  extern __shared__ __align__(128) uint8_t smraw[];
  uint8_t *smb = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smraw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t *mine = smb + (size_t)warp * kCholWarpBytes;
  float *dinv = reinterpret_cast<float *>(mine + 2 * kRecBytes);
  const uint32_t bars = smem_u32(smb + (size_t)kCholWarps * kCholWarpBytes) + warp * 16;
  if (lane == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int stride = (int)gridDim.x * kCholWarps;
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
  int j = next_job((int)blockIdx.x * kCholWarps + warp);
  if (j < a.n_jobs) fetch(j, 0);
  uint32_t ph = 0u;  // bit b: phase parity of buffer b's barrier
  for (int b = 0; j < a.n_jobs; b ^= 1) {
    const int jn = next_job(j + stride);
    __syncwarp();
    if (jn < a.n_jobs) fetch(jn, b ^ 1);
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
  }
}

int launch_chol64(mfb_engine *e, const CholArgs &c) {
  if (c.n_jobs <= 0) return 0;
  MFB_CUDA(cudaFuncSetAttribute(als_chol64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholSmem));
  const int grid = std::min(e->sm_count, (c.n_jobs + kCholWarps - 1) / kCholWarps);
  MFB_LAUNCH(als_chol64_kernel, grid, kCholWarps * 32, kCholSmem, e->stream, c);
  return 0;
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
  (void)e; (void)a; (void)sp; (void)n_primal;
  return fail("als_mn_half_step: not built yet", __FILE__, __LINE__);
}

}  // namespace mfb
