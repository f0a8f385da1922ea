// Latency probes for the serial pivot chain of the diagonal-block factorisation (sm_100a). GPU box only.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double fast_rsqrt(double d) {
  const float f = rsqrtf((float)d);
  double y = (double)f;
  const double h = 0.5 * d;
  y = y * fma(-h, y * y, 1.5);
  y = y * fma(-h, y * y, 1.5);
  return y;
}
__device__ __forceinline__ double rsqrt64h(double d) {  // MUFU.RSQ64H seed + one third-order step
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);          // 1 - d y^2
  const double p = fma(0.375, e, 0.5);           // 1/2 + 3/8 e
  return fma(y * e, p, y);                       // y (1 + e/2 + 3/8 e^2)
}
__device__ __forceinline__ double rsqrt64h_2n(double d) {  // MUFU.RSQ64H seed + two Newton steps
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double h = 0.5 * d;
  y = y * fma(-h, y * y, 1.5);
  y = y * fma(-h, y * y, 1.5);
  return y;
}

template <int MODE>
__global__ void k_pivot(double* out, int iters) {
  __shared__ double cb[2][64];
  const int t = threadIdx.x, i = t & 63;
  double a0 = 1.0 + 1e-3 * i, acc = 0.0;
  cb[0][i] = a0; cb[1][i] = a0;
  __syncthreads();
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
    if ((t >> 6) == (k & 3) || blockDim.x <= 64) cb[k & 1][i] = a0;
    __syncthreads();
    const double d = cb[k & 1][k & 63];
    double rs;
    if (MODE == 0) rs = d;
    else if (MODE == 1) rs = fast_rsqrt(d);
    else if (MODE == 2) rs = rsqrt(d);
    else if (MODE == 3) rs = rsqrt64h(d);
    else rs = rsqrt64h_2n(d);
    const double ci = cb[k & 1][i];
    const double cid = ci * (rs * rs);
    a0 = fma(-cid, ci * 1e-3, a0 + 1e-3);
    acc += rs;
  }
  const long long t1 = clock64();
  if (t == 0) { out[0] = (double)(t1 - t0) / iters; out[1] = a0 + acc; }
}

__global__ void k_rsqrt_check(double* out) {
  double worst1 = 0, worst3 = 0, worst4 = 0;
  for (int e = -300; e <= 300; e += 7)
    for (int m = 0; m < 1000; ++m) {
      const double d = ldexp(1.0 + m * 0.000999 + threadIdx.x * 1e-7, e);
      const double ref = 1.0 / sqrt(d);
      worst1 = fmax(worst1, fabs(fast_rsqrt(d) - ref) / ref);
      worst3 = fmax(worst3, fabs(rsqrt64h(d) - ref) / ref);
      worst4 = fmax(worst4, fabs(rsqrt64h_2n(d) - ref) / ref);
    }
  if (threadIdx.x == 0) { out[0] = worst1; out[1] = worst3; out[2] = worst4; }
}

__global__ void k_panel_step(double* out, int iters) {
  const int lane = threadIdx.x & 31, q = lane & 3;
  double b0 = 1.0 + lane * 1e-3, b1 = 2.0;
  const long long t0 = clock64();
  for (int k = 0; k < iters; ++k) {
    const int qk = k & 3;
    double xv = b0 * 0.999;
    xv = __shfl_sync(0xffffffffu, xv, (lane & ~3) | qk);
    if (q == qk) b0 = xv;
    if (q > qk) b0 -= xv * 1e-3;
    b1 -= xv * 1e-4;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / iters; out[1] = b0 + b1; }
}

int main() {
  double* d; cudaMalloc(&d, 64);
  double h[4];
#define RUN(label, kern, threads)                                     \
  kern<<<1, threads>>>(d, 4096); kern<<<1, threads>>>(d, 4096);       \
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);                       \
  printf("%-56s %8.1f cycles per pivot\n", label, h[0]);
  RUN("publish + barrier + read, 256 threads", k_pivot<0>, 256);
  RUN("publish + barrier + read, 128 threads", k_pivot<0>, 128);
  RUN("publish + barrier + read,  64 threads", k_pivot<0>, 64);
  RUN("publish + barrier + read,  32 threads", k_pivot<0>, 32);
  RUN("+ fast_rsqrt (cvt, MUFU.RSQ f32, cvt, 2 Newton), 256", k_pivot<1>, 256);
  RUN("+ fast_rsqrt,  64 threads", k_pivot<1>, 64);
  RUN("+ rsqrt() builtin, 256 threads", k_pivot<2>, 256);
  RUN("+ MUFU.RSQ64H + 3rd-order step, 256 threads", k_pivot<3>, 256);
  RUN("+ MUFU.RSQ64H + 3rd-order step,  64 threads", k_pivot<3>, 64);
  RUN("+ MUFU.RSQ64H + 2 Newton, 256 threads", k_pivot<4>, 256);
  RUN("panel step (DMUL, quad SHFL, select, DFMA), 1 warp", k_panel_step, 32);
  k_rsqrt_check<<<1, 32>>>(d);
  cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("max rel err: fast_rsqrt %.3e, rsqrt64h+3rd %.3e, rsqrt64h+2N %.3e  [%s]\n", h[0], h[1], h[2], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
