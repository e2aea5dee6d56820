// Stand-alone probe (diagnostic, not product code): which shared-memory floats does tcgen05.mma kind::tf32 read for an
// MN-major A operand?  B is K-major (layout verified by tools/tc_probe.cu) and selects one k: B[n][k] = (k == k0), so
// D[m][n] = A[m][k0] as the hardware read it; the A image holds its own float index at every position.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Args { uint32_t a_off, b_off, a_lbo, a_sbo, a_type, b_lbo, b_sbo, b_type, idesc, img_bytes; };
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
  // poison the accumulator so that a dropped instruction is visible
  for (int c0 = 0; c0 < 128; c0 += 4) {
    uint32_t v = __float_as_uint(-777.f);
    uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v), "r"(v), "r"(v), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    uint64_t da = make_desc(smem_u32(smb) + a.a_off, a.a_lbo, a.a_sbo, a.a_type);
    uint64_t db = make_desc(smem_u32(smb) + a.b_off, a.b_lbo, a.b_sbo, a.b_type);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(a.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(&bar)), "r"(0u)
      : "memory");
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

int main() {
  const int M = 128, N = 64, K = 8;
  const uint32_t IMG = 32768, AOFF = 0, BOFF = 16384, A_FLOATS = 2048;
  uint8_t *d_img; float *d_out;
  CK(cudaMalloc(&d_img, IMG)); CK(cudaMalloc(&d_out, sizeof(float) * 128 * 128));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, IMG + 2048));
  std::vector<float> out(128 * 128);
  const uint32_t TF32 = 2, F32 = 1;
  auto idesc = [&](int a_major, int b_major) {
    return (F32 << 4) | (TF32 << 7) | (TF32 << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  };
  struct V { const char *name; uint32_t lbo, sbo, type; int a_major; };
  const V vs[] = {
      {"A K-major noswz lbo=128 sbo=256 (control)", 128, 256, 0, 0},
      {"A MN noswz lbo=1024 sbo=128", 1024, 128, 0, 1},
      {"A MN noswz lbo=128 sbo=1024", 128, 1024, 0, 1},
      {"A MN noswz lbo=2048 sbo=128", 2048, 128, 0, 1},
      {"A MN noswz lbo=128 sbo=2048", 128, 2048, 0, 1},
      {"A MN sw128 lbo=1024 sbo=4096", 1024, 4096, 2, 1},
      {"A MN sw128 lbo=4096 sbo=1024", 4096, 1024, 2, 1},
      {"A MN sw32 lbo=256 sbo=1024", 256, 1024, 6, 1},
      {"A MN sw32 lbo=1024 sbo=256", 1024, 256, 6, 1},
      {"A MN sw64 lbo=512 sbo=2048", 512, 2048, 4, 1},
      {"A MN base32b lbo=1024 sbo=4096", 1024, 4096, 1, 1},
  };
  for (const V &v : vs) {
    printf("== %s\n", v.name);
    std::vector<int> idx(M * K, -1);
    bool dropped = false;
    for (int k0 = 0; k0 < K; k0++) {
      std::vector<uint8_t> img(IMG, 0);
      for (uint32_t i = 0; i < A_FLOATS; i++) { float f = (float)i; memcpy(&img[AOFF + 4 * i], &f, 4); }
      for (int n = 0; n < N; n++) {  // B K-major, no swizzle: adjacent K cores 128 B apart, 8-row groups 256 B apart
        float one = 1.f;
        uint32_t o = (n % 8) * 16 + (n / 8) * 256 + (k0 % 4) * 4 + (k0 / 4) * 128;
        memcpy(&img[BOFF + o], &one, 4);
      }
      Args a = {AOFF, BOFF, v.lbo, v.sbo, v.type, 128, 256, 0, idesc(v.a_major, 0), IMG};
      CK(cudaMemcpy(d_img, img.data(), IMG, cudaMemcpyHostToDevice));
      CK(cudaMemset(d_out, 0xFF, sizeof(float) * 128 * 128));
      probe<<<1, 128, IMG + 2048>>>(d_img, d_out, a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("   CUDA error %s\n", cudaGetErrorString(e)); return 2; }
      CK(cudaMemcpy(out.data(), d_out, sizeof(float) * 128 * 128, cudaMemcpyDeviceToHost));
      for (int m = 0; m < M; m++) {
        idx[m * K + k0] = (int)out[m * 128 + 0];
        if (out[m * 128] == -777.f) dropped = true;
        if (out[m * 128 + 0] != out[m * 128 + 5]) idx[m * K + k0] = -2;  // columns differ: B was not read as intended
      }
    }
    if (dropped) printf("   accumulator untouched (instruction dropped)\n");
    for (int m : {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 63, 64, 65, 127}) {
      printf("   m=%3d: float index read for k=0..7:", m);
      for (int k = 0; k < K; k++) printf(" %5d", idx[m * K + k]);
      printf("\n");
    }
  }
  return 0;
}
