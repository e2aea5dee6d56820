// Stand-alone probe of tcgen05.mma kind::tf32 operand layouts (diagnostic, not product code).
// One CTA, one MMA M=128 N=128 K=8; host prepares the shared-memory images for each variant.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Args { uint32_t a_off, b_off, lbo, sbo, layout_type, idesc, img_bytes; int do_st_test; };

__device__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(const uint8_t *img, float *out, Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t *smb = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid; i < a.img_bytes / 4; i += 128) ((uint32_t *)smb)[i] = ((const uint32_t *)img)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (a.do_st_test) {
    // write lane*1000 + column into TMEM with tcgen05.st, then read back below
    for (int c0 = 0; c0 < 128; c0 += 4) {
      uint32_t v0 = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0)), v1 = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + 1)),
               v2 = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + 2)), v3 = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + 3));
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v0), "r"(v1), "r"(v2), "r"(v3) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  } else if (tid == 0) {
    uint64_t da = make_desc(smem_u32(smb) + a.a_off, a.lbo, a.sbo, a.layout_type);
    uint64_t db = make_desc(smem_u32(smb) + a.b_off, a.lbo, a.sbo, a.layout_type);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(a.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  if (!a.do_st_test) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(&bar)), "r"(0u)
        : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();
  for (int c0 = 0; c0 < 128; c0 += 4) {
    uint32_t v0, v1, v2, v3;
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float *o = out + (size_t)(warp * 32 + lane) * 128 + c0;
    o[0] = __uint_as_float(v0); o[1] = __uint_as_float(v1); o[2] = __uint_as_float(v2); o[3] = __uint_as_float(v3);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int M = 128, N = 128, K = 8;
  std::vector<float> A(M * K), B(N * K);
  srand(1);
  for (auto &x : A) x = tf32((float)(rand() % 2000 - 1000) / 256.f);
  for (auto &x : B) x = tf32((float)(rand() % 2000 - 1000) / 256.f);
  std::vector<float> D(M * N);
  for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { float s = 0; for (int k = 0; k < K; k++) s += A[m * K + k] * B[n * K + k]; D[m * N + n] = s; }
  uint8_t *d_img; float *d_out;
  const uint32_t IMG = 32768;
  CK(cudaMalloc(&d_img, IMG)); CK(cudaMalloc(&d_out, sizeof(float) * M * N));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 2048));
  std::vector<float> out(M * N);
  auto run = [&](const char *name, std::vector<uint8_t> &img, Args a) {
    CK(cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xFF, sizeof(float) * M * N));
    a.img_bytes = IMG;
    probe<<<1, 128, IMG + 2048>>>(d_img, d_out, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s CUDA error %s\n", name, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(out.data(), d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    double err = 0, nrm = 0, errT = 0; int nz = 0;
    for (int i = 0; i < M * N; i++) { err += (out[i] - D[i]) * (double)(out[i] - D[i]); nrm += (double)D[i] * D[i]; nz += out[i] != 0.f; }
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { double d = out[m * N + n] - D[n * N + m]; errT += d * d; }
    printf("%-40s rel err %.3e  (vs transposed %.3e)  nonzeros %d  out[0..3] %g %g %g %g  want %g %g %g %g\n", name, sqrt(err / nrm), sqrt(errT / nrm), nz,
           out[0], out[1], out[2], out[3], D[0], D[1], D[2], D[3]);
  };
  const uint32_t TF32 = 2, F32 = 1;
  auto idesc = [&](int a_major, int b_major) {
    return (F32 << 4) | (TF32 << 7) | (TF32 << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  };
  // ---- 0. TMEM st/ld round trip
  {
    std::vector<uint8_t> img(IMG, 0);
    Args a = {0, 0, 0, 0, 0, 0, IMG, 1};
    CK(cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice));
    probe<<<1, 128, IMG + 2048>>>(d_img, d_out, a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) bad += out[m * N + n] != (float)(m * 1000 + n);
    printf("tmem st/ld round trip: %d mismatches (out[5][7] = %g)\n", bad, out[5 * N + 7]);
  }
  const uint32_t AOFF = 0, BOFF = 16384;
  auto put = [&](std::vector<uint8_t> &img, uint32_t off, float v) { memcpy(&img[off], &v, 4); };
  // ---- 1. K-major, no swizzle: core matrix = 8 rows (MN) x 16 B (4 K-elements), K = 8 -> 2 cores along K
  for (int lbo_first = 0; lbo_first < 2; lbo_first++) {
    std::vector<uint8_t> img(IMG, 0);
    // cores along K adjacent (128 B apart), 8-row groups 256 B apart
    const uint32_t kstride = 128, mstride = 256;
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) {
      uint32_t o = (m % 8) * 16 + (m / 8) * mstride + (k % 4) * 4 + (k / 4) * kstride;
      put(img, AOFF + o, A[m * K + k]); put(img, BOFF + o, B[m * K + k]);
    }
    Args a = {AOFF, BOFF, lbo_first ? kstride : mstride, lbo_first ? mstride : kstride, 0, idesc(0, 0), IMG, 0};
    run(lbo_first ? "K-major noswz LBO=Kstride SBO=MNstride" : "K-major noswz LBO=MNstride SBO=Kstride", img, a);
  }
  // ---- 2. MN-major, no swizzle: core matrix = 8 K-rows x 16 B (4 MN-elements)
  for (int lbo_first = 0; lbo_first < 2; lbo_first++) {
    std::vector<uint8_t> img(IMG, 0);
    const uint32_t mnstride = 128;  // next group of 4 MN elements
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) {
      uint32_t o = (m / 4) * mnstride + k * 16 + (m % 4) * 4;
      put(img, AOFF + o, A[m * K + k]); put(img, BOFF + o, B[m * K + k]);
    }
    Args a = {AOFF, BOFF, lbo_first ? mnstride : 4096u, lbo_first ? 4096u : mnstride, 0, idesc(1, 1), IMG, 0};
    run(lbo_first ? "MN-major noswz LBO=MNstride" : "MN-major noswz SBO=MNstride", img, a);
  }
  // ---- 3. MN-major, 128B swizzle: chunk (32 floats) stride CS, K row stride 128 B, unit ^= k
  for (uint32_t CS : {1024u, 4096u}) for (int lbo_first = 0; lbo_first < 2; lbo_first++) {
    std::vector<uint8_t> img(IMG, 0);
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) {
      uint32_t c = m / 32, unit = (m % 32) / 4, e = m % 4;
      uint32_t o = c * CS + k * 128 + ((unit ^ (uint32_t)k) << 4) + e * 4;
      put(img, AOFF + o, A[m * K + k]); put(img, BOFF + o, B[m * K + k]);
    }
    Args a = {AOFF, BOFF, lbo_first ? CS : 1024u, lbo_first ? 1024u : CS, 2, idesc(1, 1), IMG, 0};
    char name[96];
    snprintf(name, sizeof(name), "MN-major SW128 chunk=%u %s", CS, lbo_first ? "LBO=chunk" : "SBO=chunk");
    run(name, img, a);
  }
  // ---- 4. K-major 128B swizzle is not applicable for K = 8 tf32 (32 B rows): use SW32: row = 32 B
  {
    std::vector<uint8_t> img(IMG, 0);
    // K-major SW32: 8 rows x 32 B atom (256 B), unit(16 B) index ^= (row >> 2)&1 ... probe identity placement
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) {
      uint32_t row = m % 8, grp = m / 8, unit = k / 4, e = k % 4;
      uint32_t o = grp * 256 + row * 32 + ((unit ^ ((row >> 2) & 1)) << 4) + e * 4;
      put(img, AOFF + o, A[m * K + k]); put(img, BOFF + o, B[m * K + k]);
    }
    Args a = {AOFF, BOFF, 16, 256, 6, idesc(0, 0), IMG, 0};
    run("K-major SW32 SBO=256", img, a);
  }
  return 0;
}
