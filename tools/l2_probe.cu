// L2 micro-benchmarks for the roofline of the SGD / evaluation kernels: on the Netflix shape the factor
// matrices (U 123 MB + V 4.5 MB) live in the 126 MB L2, so the HBM figure says nothing about those kernels.  Measures, on
// the GPU it runs on, with CUDA events:
//   l2_read_gbs      coalesced 128-bit ld.global.cg over a 48 MB buffer, repeated (pure L2 read bandwidth)
//   gather_gbs       random 256-byte row gathers (16 lanes x float4) from a 123 MB matrix
//   red_v4_gbs       red.global.add.v4.f32 into random 256-byte rows of the same matrix
//   gather_red_gbs   the SGD kernel's own pattern per "rating": gather one 256 B row of a 123 MB matrix and one of a
//                    4.5 MB matrix, then reduce 256 B into each; counted like the algorithmic bytes of SURVEY 8d
//                    (4 x 256 B + 12 B per rating) — the speed of light of that access pattern without any arithmetic
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_probe tools/l2_probe.cu ; prints one JSON line.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ void red_add_v4(float4 *addr, float4 d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w) : "memory");
}

__global__ void read_kernel(const float4 *__restrict__ buf, size_t n, int reps, float *sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; r++)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float4 v = __ldcg(buf + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;
}

// mode 1: gather a row of A; 2: red into a row of A; 3: gather A-row + B-row, red into both
__global__ void row_kernel(float4 *A, uint32_t rowsA, float4 *B, uint32_t rowsB, int iters, int mode, float *sink) {
  const int lane = threadIdx.x & 31, sl = lane & 15, sub = lane >> 4;
  const uint64_t wid = (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2 + sub;
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    const uint64_t h = mix64(wid * 0x100000001B3ull + (uint64_t)it);
    const uint32_t ra = (uint32_t)(h % rowsA), rb = (uint32_t)((h >> 32) % rowsB);
    float4 *pa = A + (size_t)ra * 16 + sl, *pb = B + (size_t)rb * 16 + sl;
    if (mode == 1) {
      const float4 v = __ldcg(pa);
      acc += v.x + v.w;
    } else if (mode == 2) {
      red_add_v4(pa, make_float4(1e-9f, 0.f, 0.f, 0.f));
    } else {
      const float4 u = __ldcg(pa), v = __ldcg(pb);
      acc += u.x * v.x;
      red_add_v4(pa, make_float4(v.x * 1e-9f, v.y * 1e-9f, v.z * 1e-9f, v.w * 1e-9f));
      red_add_v4(pb, make_float4(u.x * 1e-9f, u.y * 1e-9f, u.z * 1e-9f, u.w * 1e-9f));
    }
  }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const uint32_t rowsA = 480189, rowsB = 17770;  // the bench shape at rank 64: 256-byte rows
  float4 *A, *B, *R;
  float *sink;
  const size_t nread = (size_t)48 << 20 >> 4;  // 48 MB of float4
  CK(cudaMalloc(&A, (size_t)rowsA * 256));
  CK(cudaMalloc(&B, (size_t)rowsB * 256));
  CK(cudaMalloc(&R, nread * 16));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(A, 0, (size_t)rowsA * 256));
  CK(cudaMemset(B, 0, (size_t)rowsB * 256));
  CK(cudaMemset(R, 0, nread * 16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;
  // 1. coalesced L2 reads
  const int reps = 40;
  read_kernel<<<sms * 8, 512>>>(R, nread, 2, sink);
  CK(cudaEventRecord(e0));
  read_kernel<<<sms * 8, 512>>>(R, nread, reps, sink);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double l2_read = (double)nread * 16 * reps / (ms * 1e-3) / 1e9;
  double res[4] = {0, 0, 0, 0};
  const int iters = 4096;
  const int grid = sms * 8, tb = 512;  // 64 warps per SM, two rows in flight per warp
  const double rows_total = (double)grid * tb / 16 * iters;
  for (int mode = 1; mode <= 3; mode++) {
    row_kernel<<<grid, tb>>>(A, rowsA, B, rowsB, 256, mode, sink);
    CK(cudaEventRecord(e0));
    row_kernel<<<grid, tb>>>(A, rowsA, B, rowsB, iters, mode, sink);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = mode == 3 ? rows_total * (4 * 256 + 12) : rows_total * 256;
    res[mode] = bytes / (ms * 1e-3) / 1e9;
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"l2_read_gbs\": %.1f, \"gather_gbs\": %.1f, \"red_v4_gbs\": %.1f, "
         "\"gather_red_gbs\": %.1f, \"gather_red_ratings_per_s\": %.4g, \"rows_a\": %u, \"rows_b\": %u, \"row_bytes\": 256}\n",
         prop.name, sms, prop.l2CacheSize, l2_read, res[1], res[2], res[3], res[3] * 1e9 / (4 * 256 + 12), rowsA, rowsB);
  return 0;
}
