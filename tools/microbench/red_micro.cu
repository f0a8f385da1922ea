// K3 question: 6 scalar red.global.add.f64 per 48-byte row of a 6x6 block, or ONE cp.reduce.async.bulk (TMA reduce-add, FP64,
// SASS UBLKRED.G.S.ADD.F64) of the same row staged in shared memory? Rows scattered over a 290 MB matrix like S. GPU box only.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ size_t row_addr(unsigned long long k, int n_blocks, int ld) {
  // pseudo-random 6x6 block (bi > bj) of an n_blocks x n_blocks grid, row a of it
  unsigned long long h = k * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  const int bi = 1 + (int)(h % (unsigned)(n_blocks - 1)), bj = (int)((h >> 20) % (unsigned)bi), a = (int)((h >> 50) % 6);
  return (size_t)(6 * bi + a) * ld + 6 * bj;
}

__global__ void __launch_bounds__(128) k_scalar(double* S, int ld, int n_blocks, int rows_per_thread) {
  const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  for (int it = 0; it < rows_per_thread; ++it) {
    double* dst = S + row_addr(t * rows_per_thread + it, n_blocks, ld);
    const double v = -1e-9 * (double)(it + 1);
#pragma unroll
    for (int b = 0; b < 6; ++b) asm volatile("red.global.add.f64 [%0], %1;" ::"l"(dst + b), "d"(v * (b + 1)) : "memory");
  }
}

__global__ void __launch_bounds__(128) k_bulk(double* S, int ld, int n_blocks, int rows_per_thread) {
  __shared__ __align__(16) double stage[2][128][6];
  const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  for (int it = 0; it < rows_per_thread; ++it) {
    double* dst = S + row_addr(t * rows_per_thread + it, n_blocks, ld);
    double* slot = stage[it & 1][threadIdx.x];
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the op that read this slot two rows ago is done with it
    const double v = -1e-9 * (double)(it + 1);
#pragma unroll
    for (int b = 0; b < 6; ++b) slot[b] = v * (b + 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned sa = (unsigned)__cvta_generic_to_shared(slot);
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(sa), "r"(48) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int n = 6016, ld = 6016, n_blocks = 1000;
  double* S;
  cudaMalloc(&S, sizeof(double) * (size_t)n * ld);
  cudaMemset(S, 0, sizeof(double) * (size_t)n * ld);
  const long long rows = 27000000;  // 4.5 M camera pairs x 6 rows (C2)
  const int rpt = 64, threads = 128;
  const int grid = (int)((rows / rpt + threads - 1) / threads);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 2; ++mode)
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a);
      if (mode == 0) k_scalar<<<grid, threads>>>(S, ld, n_blocks, rpt); else k_bulk<<<grid, threads>>>(S, ld, n_blocks, rpt);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      printf("%s: %lld rows of 48 B in %.3f ms = %.1f M rows/s (%s)\n", mode == 0 ? "6 x red.global.add.f64  " : "cp.reduce.async.bulk f64", (long long)grid * threads * rpt, ms,
             (double)grid * threads * rpt / ms * 1e-3, cudaGetErrorString(cudaGetLastError()));
    }
  double h[12];
  cudaMemcpy(h, S + (size_t)6 * ld, sizeof(h), cudaMemcpyDeviceToHost);
  printf("sample of S: %g %g %g\n", h[0], h[1], h[5]);
  return 0;
}
