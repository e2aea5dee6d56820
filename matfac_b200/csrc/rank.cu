// Ranking positions for the hit-rate family of metrics and per-rating predictions for NDCG.
//
// Replaces the candidate scan of Model::hitRate / arHR / hitRateU / hitRateI / arHRU / arHRI (model.cpp:981-1332):
// for every user the reference scores ALL items that are neither invalid nor rated by the user in the training matrix
// with estRating, keeps the N best in a heap and looks for the user's test item (the first rating of its row in the
// validation / test matrix).  The test item is at position p of that list exactly when p candidates score higher, so
// one number per user — the count of better candidates — serves N = 10 (hitRate) and N = 1000 (arHR) and every
// user / item filter at once; the host turns the counts into the metrics (matfac_b200/host/model.cpp).
//
//   * rank <= 64, plain-dot models (MF, IFWMF): the scores are a dense U V^T — a real GEMM (O(users x items x rank)),
//     the one place of the path where the tensor cores are fed straight from the factor matrices: 128-user x 128-item
//     tiles on tcgen05 (kind::tf32, fp32 accumulators in TMEM, 3xTF32 split prepared once per call), the count is the
//     fused epilogue — scores never leave the SM.  Items the user rated in training are counted by a sparse pass with
//     the same split arithmetic and subtracted.
//   * every other case (rank > 64, TMF / TMF+Dropout whose prediction rank depends on the (user, item) pair): one CTA per
//     user on CUDA cores, products and sums rounded exactly as the reference's float loop (model.cpp:547) does.
#include "engine.h"

namespace mfb {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait (a protocol bug traps after ~4 s instead of hanging the GPU); the warp is parked while it waits
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned long long t0 = 0;
  for (uint32_t spins = 0;; spins++) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
    if ((spins & 0x3Fu) == 0x3Fu) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) asm volatile("trap;");
    }
  }
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
// SM100 shared-memory matrix descriptor, K-major, no swizzle: core matrices of 8 rows x 16 bytes;
// LBO = bytes between core matrices along K, SBO = bytes between 8-row groups (see als.cu / tools/tc_probe.cu)
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
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
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

constexpr int kRkK = 64;                       // padded rank of the tensor-core path
constexpr int kRkTile = 128;                   // users / items per tile
constexpr uint32_t kRkLbo = 128, kRkSbo = 2048;  // 16 K-chunks of a row group adjacent, 8-row groups 2 KB apart
constexpr uint32_t kRkABytes = kRkTile * kRkK * 4;      // 32 KB: one 128-row operand
constexpr uint32_t kRkBBytes = 2 * kRkABytes;           // 64 KB: [big rows; small rows] of an item tile
constexpr int kRkThreads = 192;                // warps 0-3: count epilogue, warp 4: loader, warp 5: tensor-core issue
struct RkSmem {
  static constexpr uint32_t off_abig = 0, off_asmall = kRkABytes, off_b = 2 * kRkABytes, off_bars = off_b + 2 * kRkBBytes;
  static constexpr int n_bars = 10;  // a_full, a_empty, b_full[2], b_empty[2], t_full[2], t_empty[2]
  static constexpr uint32_t off_tmem = off_bars + n_bars * 8;
  static constexpr size_t bytes = off_tmem + 16 + 1024;
};

// ---- preparation ------------------------------------------------------------------------------------------------
// test item of every user = first rating of its row in the evaluated matrix (model.cpp:992); state: >= 0 the item,
// -1 invalid user or empty row, -2 the item is no candidate (invalid, outside the training matrix, or rated in training)
__global__ void rank_test_item_kernel(const int64_t *__restrict__ te_ptr, const int32_t *__restrict__ te_ind,
                                      const int64_t *__restrict__ tr_ptr, const int32_t *__restrict__ tr_ind, int tr_ncols,
                                      const uint8_t *__restrict__ bad_user, const uint8_t *__restrict__ bad_item, int n_users,
                                      int32_t *__restrict__ tst, int32_t *__restrict__ state) {
  const int lane = threadIdx.x & 31;
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_users) return;
  int item = -1, st = -1;
  if (!bad_user[u] && te_ptr[u + 1] > te_ptr[u]) {
    item = te_ind[te_ptr[u]];
    st = item;
    if (item >= tr_ncols || bad_item[item]) st = -2;
    bool rated = false;
    for (int64_t j = tr_ptr[u] + lane; j < tr_ptr[u + 1]; j += 32) rated |= tr_ind[j] == item;
    if (__any_sync(0xFFFFFFFFu, rated)) st = -2;
  }
  if (lane == 0) {
    tst[u] = item;
    state[u] = st;
  }
}

// F -> tf32-rounded big part and fp32 remainder, rows padded to kRkK floats (zeros), row count padded by the caller
__global__ void rank_split_kernel(const float *__restrict__ F, int n, int ld, int rank, int n_pad, float *__restrict__ big,
                                  float *__restrict__ small) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_pad * kRkK) return;
  const int row = (int)(i / kRkK), d = (int)(i % kRkK);
  float f = 0.f;
  if (row < n && d < rank) f = F[(size_t)row * ld + d];
  const float b = __uint_as_float((__float_as_uint(f) + 0x1000u) & 0xFFFFE000u);
  big[i] = b;
  small[i] = f - b;  // exact; the tensor core reads its upper 19 bits
}

// score of (u, item) in the arithmetic of the tensor-core path: big.big + big.small + small.big, fp32
__device__ __forceinline__ float split_score(const float *ub, const float *us, const float *vb, const float *vs, int lane) {
  // two dims per lane (kRkK = 64)
  const float2 a = reinterpret_cast<const float2 *>(ub)[lane], as = reinterpret_cast<const float2 *>(us)[lane];
  const float2 b = reinterpret_cast<const float2 *>(vb)[lane], bs = reinterpret_cast<const float2 *>(vs)[lane];
  float s = a.x * b.x + a.y * b.y;
  s += a.x * bs.x + a.y * bs.y;
  s += as.x * b.x + as.y * b.y;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, m);
  return s;
}

// one warp per user: score of the test item, and the number of TRAINING items of the user that are valid, are not the
// test item and score higher (they are part of the dense count and are no candidates)
__global__ void rank_sparse_kernel(const float *__restrict__ Ub, const float *__restrict__ Us, const float *__restrict__ Vb,
                                   const float *__restrict__ Vs, const int64_t *__restrict__ tr_ptr,
                                   const int32_t *__restrict__ tr_ind, int tr_ncols, const uint8_t *__restrict__ bad_item,
                                   const int32_t *__restrict__ state, int n_users, float *__restrict__ s_test,
                                   int32_t *__restrict__ train_better, int phase) {
  const int lane = threadIdx.x & 31;
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_users) return;
  const int item = state[u];
  if (item < 0) {
    if (lane == 0 && phase == 0) s_test[u] = 0.f;
    return;
  }
  const float *ub = Ub + (size_t)u * kRkK, *us = Us + (size_t)u * kRkK;
  if (phase == 0) {
    const float s = split_score(ub, us, Vb + (size_t)item * kRkK, Vs + (size_t)item * kRkK, lane);
    if (lane == 0) s_test[u] = s;
    return;
  }
  const float st = s_test[u];
  int cnt = 0;
  for (int64_t j = tr_ptr[u]; j < tr_ptr[u + 1]; j++) {
    const int it = tr_ind[j];
    if (it >= tr_ncols || bad_item[it] || it == item) continue;
    const float s = split_score(ub, us, Vb + (size_t)it * kRkK, Vs + (size_t)it * kRkK, lane);
    cnt += s > st;
  }
  if (lane == 0) train_better[u] = cnt;
}

// bit i of word w = item 32 w + i may be a candidate (valid, inside the training matrix)
__global__ void rank_item_mask_kernel(const uint8_t *__restrict__ bad_item, int tr_ncols, int n_words, uint32_t *__restrict__ mask) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) {
    const int it = 32 * w + i;
    if (it < tr_ncols && !bad_item[it]) m |= 1u << i;
  }
  mask[w] = m;
}

struct RkArgs {
  const float *Ub, *Us, *Vb, *Vs;  // [rows padded to 128][64]
  const uint32_t *item_mask;       // [item tiles * 4]
  const int32_t *state;            // test item or < 0
  const float *s_test;
  int32_t *better;                 // out: valid items (test item excluded) that score higher
  int n_users, n_user_tiles, n_item_tiles;
};

// ---- dense count on the tensor cores -----------------------------------------------------------------------------
// Persistent CTAs, one 128-user tile at a time against all item tiles.  Per item tile and K step of 8 dims two
// instructions:  D[:, 0:256] (+)= Ubig x [Vbig; Vsmall]^T  (M 128, N 256)  and  D[:, 0:128] += Usmall x Vbig^T (N 128);
// score(u, i) = D[u][i] + D[u][128 + i].  Two TMEM accumulators (2 x 256 columns): the epilogue warps — one thread per
// user: TMEM lane = row — compare and count tile t while tile t + 1 is multiplied.
__global__ void __launch_bounds__(kRkThreads, 1) rank_count_tc_kernel(const RkArgs a) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t *smb = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smb);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smb + RkSmem::off_tmem);
  const uint32_t bar0 = sbase + RkSmem::off_bars;
  const uint32_t a_full = bar0, a_empty = bar0 + 8, b_full = bar0 + 16, b_empty = bar0 + 32, t_full = bar0 + 48, t_empty = bar0 + 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(a_full, 32);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; i++) {
      mbar_init(b_full + i * 8, 32);
      mbar_init(b_empty + i * 8, 1);
      mbar_init(t_full + i * 8, 1);
      mbar_init(t_empty + i * 8, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int n_my = (int)blockIdx.x < a.n_user_tiles ? (a.n_user_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 4) {
    // ---------------- loader: 16-byte cp.async copies into the core-matrix layout ----------------
    // unit (row m, chunk c) of a 128-row operand -> (m / 8) * SBO + c * LBO + (m % 8) * 16; a warp instruction reads two
    // whole rows (2 x 256 contiguous bytes).  A tile is published (fence.proxy.async + arrive) once its copy group has
    // landed — one tile behind the copies being issued.
    auto copy_rows = [&](uint32_t dst, const float *src) {  // 128 rows x 16 chunks
#pragma unroll 4
      for (int k = 0; k < kRkTile * 16 / 32; k++) {
        const int idx = k * 32 + lane, m = idx >> 4, c = idx & 15;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(m >> 3) * kRkSbo + (uint32_t)c * kRkLbo + (uint32_t)(m & 7) * 16),
                     "l"(src + (size_t)m * kRkK + c * 4)
                     : "memory");
      }
    };
    int nb = 0;           // item tiles issued so far (all user tiles)
    int pending_bar = 0;  // barrier to arrive on when the previous group has landed (0 = none)
    bool pending_a = false;
    auto publish_pending = [&](int groups_in_flight) {
      if (!pending_bar) return;
      if (groups_in_flight) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group 0;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (pending_a) mbar_arrive(a_full);
      mbar_arrive((uint32_t)pending_bar);
      pending_bar = 0;
    };
    for (int i = 0; i < n_my; i++) {
      const int ut = (int)blockIdx.x + i * (int)gridDim.x;
      if (i > 0) {
        publish_pending(0);  // the previous user tile's last item tile: its MMAs are what frees the A tiles
        mbar_wait(a_empty, (uint32_t)(i - 1) & 1u);
      }
      copy_rows(sbase + RkSmem::off_abig, a.Ub + (size_t)ut * kRkTile * kRkK);
      copy_rows(sbase + RkSmem::off_asmall, a.Us + (size_t)ut * kRkTile * kRkK);
      for (int t = 0; t < a.n_item_tiles; t++, nb++) {
        const int st = nb & 1;
        if (nb >= 2) mbar_wait(b_empty + st * 8, (uint32_t)(nb / 2 - 1) & 1u);
        const uint32_t dst = sbase + RkSmem::off_b + st * kRkBBytes;
        copy_rows(dst, a.Vb + (size_t)t * kRkTile * kRkK);
        copy_rows(dst + kRkABytes, a.Vs + (size_t)t * kRkTile * kRkK);
        asm volatile("cp.async.commit_group;" ::: "memory");
        publish_pending(1);
        pending_a = t == 0;  // the A tiles travel in the group of the user tile's first item tile
        pending_bar = (int)(b_full + st * 8);
      }
    }
    publish_pending(0);
  } else if (warp == 5) {
    // ---------------- tensor-core issue (converged warp, one elected lane) ----------------
    constexpr uint32_t idesc256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
    constexpr uint32_t idesc128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t d_abig = umma_desc_kmajor(sbase + RkSmem::off_abig, kRkLbo, kRkSbo);
    const uint64_t d_asmall = umma_desc_kmajor(sbase + RkSmem::off_asmall, kRkLbo, kRkSbo);
    const uint64_t d_b0 = umma_desc_kmajor(sbase + RkSmem::off_b, kRkLbo, kRkSbo);
    int nb = 0;
    for (int i = 0; i < n_my; i++) {
      mbar_wait(a_full, (uint32_t)i & 1u);
      for (int t = 0; t < a.n_item_tiles; t++, nb++) {
        const int st = nb & 1;  // operand stage and accumulator alternate together
        mbar_wait(b_full + st * 8, (uint32_t)(nb / 2) & 1u);
        if (nb >= 2) mbar_wait(t_empty + st * 8, (uint32_t)(nb / 2 - 1) & 1u);  // the epilogue has read this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one_sync()) {
          const uint32_t dt = tmem_base + (uint32_t)st * 256u;
          const uint64_t d_b = d_b0 + (uint64_t)((uint32_t)st * (kRkBBytes >> 4));
#pragma unroll
          for (int k8 = 0; k8 < kRkK / 8; k8++) {
            const uint64_t ko = (uint64_t)(k8 * ((2 * kRkLbo) >> 4));
            umma_tf32(dt, d_abig + ko, d_b + ko, idesc256, k8 == 0 ? 0u : 1u);
            umma_tf32(dt, d_asmall + ko, d_b + ko, idesc128, 1u);
          }
          umma_commit(b_empty + st * 8);
          umma_commit(t_full + st * 8);
          if (t == a.n_item_tiles - 1) umma_commit(a_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- count epilogue: thread = user (TMEM lane 32 warp + lane) ----------------
    int nb = 0;
    for (int i = 0; i < n_my; i++) {
      const int ut = (int)blockIdx.x + i * (int)gridDim.x;
      const int u = ut * kRkTile + warp * 32 + lane;
      const int item = u < a.n_users ? a.state[u] : -1;
      const float st_score = u < a.n_users ? a.s_test[u] : 0.f;
      int cnt = 0;
      for (int t = 0; t < a.n_item_tiles; t++, nb++) {
        const int st = nb & 1;
        mbar_wait(t_full + st * 8, (uint32_t)(nb / 2) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tad = tmem_base + (uint32_t)st * 256u + ((uint32_t)(warp * 32) << 16);
        const int rel = item - t * kRkTile;  // the test item's column in this tile, if any
#pragma unroll 1
        for (int c0 = 0; c0 < kRkTile; c0 += 32) {
          uint32_t v[32], x[32];
          tmem_ld32(tad + c0, v);
          tmem_ld32(tad + 128 + c0, x);
          tmem_ld_wait();
          uint32_t m = __ldg(a.item_mask + t * 4 + (c0 >> 5));
          if (rel >= c0 && rel < c0 + 32) m &= ~(1u << (rel - c0));
#pragma unroll
          for (int k = 0; k < 32; k++) cnt += ((m >> k) & 1u) && (__uint_as_float(v[k]) + __uint_as_float(x[k]) > st_score);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + st * 8);
      }
      if (u < a.n_users) a.better[u] = cnt;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

__global__ void rank_combine_kernel(const int32_t *__restrict__ state, const int32_t *__restrict__ better,
                                    const int32_t *__restrict__ train_better, int n_users, int32_t *__restrict__ pos) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  pos[u] = state[u] >= 0 ? better[u] - train_better[u] : state[u];
}

// ---- general path: one CTA per user on CUDA cores ------------------------------------------------------------------
// estRating of the variant, rounded like the reference's loops: model.cpp:547 (float products, float sum, in order) for
// MF / IFWMF; modelDropoutSigmoid.cpp:5-24 / modelPoissonDropout.cpp:5-23 accumulate the float products in double over
// the first k dims, k chosen per (user, item) pair from the prediction ranks (user side when the user is the rarer one).
template <int VARIANT>
__device__ __forceinline__ double est_rating(const float *__restrict__ u, const float *__restrict__ v, int rank, int k) {
  if (VARIANT == MFB_MF || VARIANT == MFB_IFWMF) {
    float s = 0.f;
    for (int d = 0; d < rank; d++) s = __fadd_rn(s, __fmul_rn(u[d], v[d]));
    return (double)s;
  }
  double s = 0.0;
  for (int d = 0; d < k && d < rank; d++) s += (double)__fmul_rn(u[d], v[d]);
  return s;
}

template <int VARIANT>
__global__ void __launch_bounds__(256) rank_count_generic_kernel(const float *__restrict__ U, const float *__restrict__ V, int ld,
                                                                 int rank, const int64_t *__restrict__ tr_ptr,
                                                                 const int32_t *__restrict__ tr_ind, int tr_ncols,
                                                                 const uint8_t *__restrict__ bad_item, const Aux *__restrict__ aux_u,
                                                                 const Aux *__restrict__ aux_i, const int32_t *__restrict__ state,
                                                                 int32_t *__restrict__ pos) {
  extern __shared__ float su[];
  __shared__ int s_cnt[8];
  const int u = blockIdx.x, tid = threadIdx.x;
  const int item = state[u];
  if (item < 0) {
    if (tid == 0) pos[u] = item;
    return;
  }
  for (int d = tid; d < rank; d += blockDim.x) su[d] = U[(size_t)u * ld + d];
  __syncthreads();
  int ufreq = 0, upred = 0;
  if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
    const Aux au = aux_u[u];
    ufreq = au.freq; upred = au.pred;
  }
  auto score = [&](int it) {
    int k = rank;
    if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
      const Aux ai = aux_i[it];
      k = ufreq < ai.freq ? upred : ai.pred;
    }
    return est_rating<VARIANT>(su, V + (size_t)it * ld, rank, k);
  };
  const double st = score(item);
  int cnt = 0;
  for (int it = tid; it < tr_ncols; it += blockDim.x)
    if (!bad_item[it] && it != item) cnt += score(it) > st;
  for (int64_t j = tr_ptr[u] + tid; j < tr_ptr[u + 1]; j += blockDim.x) {  // training items are no candidates
    const int it = tr_ind[j];
    if (it < tr_ncols && !bad_item[it] && it != item) cnt -= score(it) > st;
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, m);
  if ((tid & 31) == 0) s_cnt[tid >> 5] = cnt;
  __syncthreads();
  if (tid == 0) {
    int s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_cnt[w];
    pos[u] = s;
  }
}

// predicted rating of every rating of a matrix, CSR order (NaN where the user or the item is masked)
template <int VARIANT>
__global__ void rank_predict_kernel(const float *__restrict__ U, const float *__restrict__ V, int ld, int rank,
                                    const int64_t *__restrict__ ptr, const int32_t *__restrict__ ind, int n_users, int n_items,
                                    const uint8_t *__restrict__ bad_user, const uint8_t *__restrict__ bad_item,
                                    const Aux *__restrict__ aux_u, const Aux *__restrict__ aux_i, float *__restrict__ pred) {
  const int lane = threadIdx.x & 31;
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (u >= n_users) return;
  int ufreq = 0, upred = 0;
  if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
    const Aux au = aux_u[u];
    ufreq = au.freq; upred = au.pred;
  }
  for (int64_t j = ptr[u] + lane; j < ptr[u + 1]; j += 32) {
    const int it = ind[j];
    float p = __int_as_float(0x7FC00000);
    if (!bad_user[u] && it < n_items && !bad_item[it]) {
      int k = rank;
      if (VARIANT == MFB_TMF || VARIANT == MFB_TMFDROPOUT) {
        const Aux ai = aux_i[it];
        k = ufreq < ai.freq ? upred : ai.pred;
      }
      p = (float)est_rating<VARIANT>(U + (size_t)u * ld, V + (size_t)it * ld, rank, k);
    }
    pred[j] = p;
  }
}

}  // namespace

int rank_positions_launch(mfb_engine *e, int which, int factors, int variant, int32_t *pos_host, int32_t *test_item_host) {
  DevCsr &tr = e->mat[MFB_TRAIN], &te = e->mat[which];
  const float *U = factors == MFB_BEST ? e->bestU : e->U, *V = factors == MFB_BEST ? e->bestV : e->V;
  cudaStream_t st = e->stream;
  const int nu = e->n_users;
  int32_t *tst = nullptr, *state = nullptr, *pos = nullptr;
  MFB_CUDA(dev_alloc(&tst, sizeof(int32_t) * nu));
  MFB_CUDA(dev_alloc(&state, sizeof(int32_t) * nu));
  MFB_CUDA(dev_alloc(&pos, sizeof(int32_t) * nu));
  MFB_LAUNCH(rank_test_item_kernel, (nu * 32 + 255) / 256, 256, 0, st, te.rowptr, te.rowind, tr.rowptr, tr.rowind, tr.ncols,
             e->bad_user, e->bad_item, nu, tst, state);
  const bool plain = variant == MFB_MF || variant == MFB_IFWMF;
  if (plain && e->rank <= kRkK && e->opt_rank_tensor_cores) {
    const int ut = (nu + kRkTile - 1) / kRkTile, it = (tr.ncols + kRkTile - 1) / kRkTile;
    const int nup = ut * kRkTile, nip = it * kRkTile;
    float *Ub, *Us, *Vb, *Vs, *s_test;
    uint32_t *mask;
    int32_t *better, *train_better;
    MFB_CUDA(dev_alloc(&Ub, sizeof(float) * (size_t)nup * kRkK));
    MFB_CUDA(dev_alloc(&Us, sizeof(float) * (size_t)nup * kRkK));
    MFB_CUDA(dev_alloc(&Vb, sizeof(float) * (size_t)nip * kRkK));
    MFB_CUDA(dev_alloc(&Vs, sizeof(float) * (size_t)nip * kRkK));
    MFB_CUDA(dev_alloc(&s_test, sizeof(float) * nu));
    MFB_CUDA(dev_alloc(&mask, sizeof(uint32_t) * it * 4));
    MFB_CUDA(dev_alloc(&better, sizeof(int32_t) * nu));
    MFB_CUDA(dev_alloc(&train_better, sizeof(int32_t) * nu));
    MFB_LAUNCH(rank_split_kernel, (int)(((int64_t)nup * kRkK + 255) / 256), 256, 0, st, U, nu, e->ld, e->rank, nup, Ub, Us);
    MFB_LAUNCH(rank_split_kernel, (int)(((int64_t)nip * kRkK + 255) / 256), 256, 0, st, V, std::min(e->n_items, tr.ncols), e->ld, e->rank, nip, Vb, Vs);
    MFB_LAUNCH(rank_item_mask_kernel, (it * 4 + 255) / 256, 256, 0, st, e->bad_item, tr.ncols, it * 4, mask);
    const int wgrid = (nu * 32 + 255) / 256;
    MFB_LAUNCH(rank_sparse_kernel, wgrid, 256, 0, st, Ub, Us, Vb, Vs, tr.rowptr, tr.rowind, tr.ncols, e->bad_item, state, nu, s_test,
               train_better, 0);
    MFB_LAUNCH(rank_sparse_kernel, wgrid, 256, 0, st, Ub, Us, Vb, Vs, tr.rowptr, tr.rowind, tr.ncols, e->bad_item, state, nu, s_test,
               train_better, 1);
    RkArgs a{Ub, Us, Vb, Vs, mask, state, s_test, better, nu, ut, it};
    MFB_CUDA(cudaFuncSetAttribute(rank_count_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RkSmem::bytes));
    MFB_LAUNCH(rank_count_tc_kernel, std::min(e->sm_count, ut), kRkThreads, RkSmem::bytes, st, a);
    MFB_LAUNCH(rank_combine_kernel, (nu + 255) / 256, 256, 0, st, state, better, train_better, nu, pos);
    dev_free(Ub); dev_free(Us); dev_free(Vb); dev_free(Vs); dev_free(s_test); dev_free(mask); dev_free(better); dev_free(train_better);
  } else {
    const size_t smem = sizeof(float) * (size_t)e->rank;
    if (variant == MFB_TMF || variant == MFB_TMFDROPOUT)
      MFB_LAUNCH((rank_count_generic_kernel<MFB_TMF>), nu, 256, smem, st, U, V, e->ld, e->rank, tr.rowptr, tr.rowind, tr.ncols, e->bad_item,
                 e->aux_u, e->aux_i, state, pos);
    else
      MFB_LAUNCH((rank_count_generic_kernel<MFB_MF>), nu, 256, smem, st, U, V, e->ld, e->rank, tr.rowptr, tr.rowind, tr.ncols, e->bad_item,
                 e->aux_u, e->aux_i, state, pos);
  }
  MFB_CUDA(cudaMemcpyAsync(pos_host, pos, sizeof(int32_t) * nu, cudaMemcpyDeviceToHost, st));
  if (test_item_host) MFB_CUDA(cudaMemcpyAsync(test_item_host, tst, sizeof(int32_t) * nu, cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  dev_free(tst); dev_free(state); dev_free(pos);
  return 0;
}

int rank_predict_launch(mfb_engine *e, int which, int factors, int variant, float *pred_host) {
  DevCsr &m = e->mat[which];
  const float *U = factors == MFB_BEST ? e->bestU : e->U, *V = factors == MFB_BEST ? e->bestV : e->V;
  if (m.nnz == 0) return 0;
  float *pred = nullptr;
  MFB_CUDA(dev_alloc(&pred, sizeof(float) * (size_t)m.nnz));
  const int grid = (e->n_users * 32 + 255) / 256;
  if (variant == MFB_TMF || variant == MFB_TMFDROPOUT)
    MFB_LAUNCH((rank_predict_kernel<MFB_TMF>), grid, 256, 0, e->stream, U, V, e->ld, e->rank, m.rowptr, m.rowind, e->n_users, e->n_items,
               e->bad_user, e->bad_item, e->aux_u, e->aux_i, pred);
  else
    MFB_LAUNCH((rank_predict_kernel<MFB_MF>), grid, 256, 0, e->stream, U, V, e->ld, e->rank, m.rowptr, m.rowind, e->n_users, e->n_items,
               e->bad_user, e->bad_item, e->aux_u, e->aux_i, pred);
  MFB_CUDA(cudaMemcpyAsync(pred_host, pred, sizeof(float) * (size_t)m.nnz, cudaMemcpyDeviceToHost, e->stream));
  MFB_CUDA(cudaStreamSynchronize(e->stream));
  dev_free(pred);
  return 0;
}

}  // namespace mfb
