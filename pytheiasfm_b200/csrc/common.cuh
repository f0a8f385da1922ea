// Shared helpers for the sm_100a kernels of libtheia_b200.so.
#ifndef THB_COMMON_CUH_
#define THB_COMMON_CUH_

#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <string>

#include "../../include/theia_b200.h"

// The per-element camera math (Dual numbers, project_*) is plain C++: under nvcc it is __host__ __device__, and the
// host adapter (csrc/adapter/pt_module.cc, compiled by g++) includes the same headers for Camera::ProjectPoint.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define THB_HD __host__ __device__ __forceinline__
#else
#define THB_HD inline
#endif

namespace thb {

#if defined(__CUDACC__)
// thread-local error text behind thb_last_error()
void SetLastError(const std::string& s);
const char* GetLastError();

#define THB_CUDA_CHECK(expr)                                                               \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      char _b[512];                                                                        \
      snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      ::thb::SetLastError(_b);                                                             \
      return THB_E_CUDA;                                                                   \
    }                                                                                      \
  } while (0)

#define THB_FAIL(code, msg)          \
  do {                               \
    ::thb::SetLastError(msg);        \
    return (code);                   \
  } while (0)

#endif  // __CUDACC__

constexpr int KS = THB_INTR_STRIDE;

// Forward-mode dual number with N tangent directions, fully unrolled into registers.
template <int N>
struct Dual {
  double a;
  double v[N];
  THB_HD Dual() {}
  THB_HD Dual(double s) : a(s) {  // NOLINT
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = 0.0;
  }
};

template <int N> THB_HD Dual<N> seed(double s, int k) {
  Dual<N> d(s);
#pragma unroll
  for (int i = 0; i < N; ++i) d.v[i] = (i == k) ? 1.0 : 0.0;
  return d;
}
template <int N> THB_HD Dual<N> operator+(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h; h.a = f.a + g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i];
  return h; }
template <int N> THB_HD Dual<N> operator-(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h; h.a = f.a - g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i];
  return h; }
template <int N> THB_HD Dual<N> operator-(const Dual<N>& f) {
  Dual<N> h; h.a = -f.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i];
  return h; }
template <int N> THB_HD Dual<N> operator*(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h; h.a = f.a * g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return h; }
template <int N> THB_HD Dual<N> operator/(const Dual<N>& f, const Dual<N>& g) {
  Dual<N> h; const double gi = 1.0 / g.a; h.a = f.a * gi;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi;
  return h; }
template <int N> THB_HD Dual<N> operator+(const Dual<N>& f, double s) { Dual<N> h = f; h.a += s; return h; }
template <int N> THB_HD Dual<N> operator+(double s, const Dual<N>& f) { Dual<N> h = f; h.a += s; return h; }
template <int N> THB_HD Dual<N> operator-(const Dual<N>& f, double s) { Dual<N> h = f; h.a -= s; return h; }
template <int N> THB_HD Dual<N> operator-(double s, const Dual<N>& f) { Dual<N> h = -f; h.a += s; return h; }
template <int N> THB_HD Dual<N> operator*(const Dual<N>& f, double s) {
  Dual<N> h; h.a = f.a * s;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s;
  return h; }
template <int N> THB_HD Dual<N> operator*(double s, const Dual<N>& f) { return f * s; }
template <int N> THB_HD Dual<N> operator/(const Dual<N>& f, double s) { return f * (1.0 / s); }
template <int N> THB_HD Dual<N> operator/(double s, const Dual<N>& g) { return Dual<N>(s) / g; }

template <int N> THB_HD Dual<N> chain(const Dual<N>& f, double val, double dval) {
  Dual<N> h; h.a = val;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = dval * f.v[i];
  return h; }

THB_HD double val(double x) { return x; }
template <int N> THB_HD double val(const Dual<N>& f) { return f.a; }

THB_HD double d_sqrt(double x) { return sqrt(x); }
THB_HD double d_tan(double x) { return tan(x); }
THB_HD double d_atan(double x) { return atan(x); }
THB_HD double d_abs(double x) { return fabs(x); }
THB_HD double d_atan2(double y, double x) { return atan2(y, x); }
THB_HD double d_sin(double x) { return sin(x); }
THB_HD double d_cos(double x) { return cos(x); }
template <int N> THB_HD Dual<N> d_sqrt(const Dual<N>& f) { const double s = sqrt(f.a); return chain(f, s, 0.5 / s); }
template <int N> THB_HD Dual<N> d_sin(const Dual<N>& f) { return chain(f, sin(f.a), cos(f.a)); }
template <int N> THB_HD Dual<N> d_cos(const Dual<N>& f) { return chain(f, cos(f.a), -sin(f.a)); }
template <int N> THB_HD Dual<N> d_tan(const Dual<N>& f) { const double t = tan(f.a); return chain(f, t, 1.0 + t * t); }
template <int N> THB_HD Dual<N> d_atan(const Dual<N>& f) { return chain(f, atan(f.a), 1.0 / (1.0 + f.a * f.a)); }
template <int N> THB_HD Dual<N> d_abs(const Dual<N>& f) { return chain(f, fabs(f.a), copysign(1.0, f.a)); }
template <int N> THB_HD Dual<N> d_atan2(const Dual<N>& y, const Dual<N>& x) {
  Dual<N> h; const double t = 1.0 / (x.a * x.a + y.a * y.a); h.a = atan2(y.a, x.a);
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = t * (x.a * y.v[i] - y.a * x.v[i]);
  return h; }

#if defined(__CUDACC__)
// Block-wide sum of one double; result valid in thread 0. blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* smem32) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? smem32[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

#endif  // __CUDACC__

}  // namespace thb
#endif  // THB_COMMON_CUH_
