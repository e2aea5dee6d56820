// Stand-alone probe (diagnostic, not product code): cycles per tcgen05.mma kind::tf32 (K = 8) issued back to back by one thread,
// for the K-major no-swizzle operand layout (the converter kernel's) and the MN-major SWIZZLE_128B_BASE32B layout (csrc/als_mn.cu),
// for several shapes, accumulating into one accumulator or rotating over four.  Operands are zeros: only time is measured.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
struct Args { uint32_t lbo, sbo, type, idesc, iters, rotate, kadv; };
__device__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
__global__ void __launch_bounds__(128) probe(long long *out, Args a) {
  extern __shared__ uint8_t raw[];
  uint8_t *smb = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < 49152 / 4; i += 128) ((uint32_t *)smb)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint64_t d0 = make_desc(smem_u32(smb), a.lbo, a.sbo, a.type);
    const long long t0 = clock64();
    for (uint32_t i = 0; i < a.iters; i++) {
      const uint64_t da = d0 + (uint64_t)((i & 3) * (a.kadv >> 4));  // four K steps of a 32-rating tile
      const uint32_t dt = tmem + (a.rotate ? (i & 3) * 128u : 0u);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(dt), "l"(da), "l"(da), "r"(a.idesc), "r"(1u)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(&bar)), "r"(0u)
        : "memory");
    out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  long long *d_out;
  CK(cudaMalloc(&d_out, sizeof(long long) * 148));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 2048));
  const uint32_t TF32 = 2, F32 = 1, ITERS = 4096;
  auto idesc = [&](int major, int M, int N) {
    return (F32 << 4) | (TF32 << 7) | (TF32 << 10) | ((uint32_t)major << 15) | ((uint32_t)major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  };
  struct V { const char *name; uint32_t lbo, sbo, type; int major, M, N; uint32_t kadv; };
  const V vs[] = {
      {"K-major  no swizzle        M128 N64 ", 2304, 144, 0, 0, 128, 64, 4608},
      {"K-major  no swizzle        M128 N128", 2304, 144, 0, 0, 128, 128, 4608},
      {"K-major  no swizzle        M128 N16 ", 2304, 144, 0, 0, 128, 16, 4608},
      {"MN-major 128B_BASE32B      M128 N64 ", 4096, 512, 1, 1, 128, 64, 1024},
      {"MN-major 128B_BASE32B      M128 N80 ", 4096, 512, 1, 1, 128, 80, 1024},
      {"MN-major 128B_BASE32B      M128 N16 ", 4096, 512, 1, 1, 128, 16, 1024},
      {"MN-major 128B_BASE32B      M64  N80 ", 4096, 512, 1, 1, 64, 80, 1024},
      {"MN-major 128B_BASE32B      M128 N256", 4096, 512, 1, 1, 128, 256, 1024},
  };
  for (int grid : {1, 148})
    for (const V &v : vs)
      for (int rotate = 0; rotate < 2; rotate++) {
        if (rotate && v.N > 128) continue;
        Args a = {v.lbo, v.sbo, v.type, idesc(v.major, v.M, v.N), ITERS, (uint32_t)rotate, v.kadv};
        probe<<<grid, 128, 49152 + 2048>>>(d_out, a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s CUDA error %s\n", v.name, cudaGetErrorString(e)); return 2; }
        long long h[148];
        CK(cudaMemcpy(h, d_out, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  %s  %s  %.1f cycles per MMA\n", grid, v.name, rotate ? "4 accumulators" : "1 accumulator ", (double)mx / ITERS);
      }
  return 0;
}
