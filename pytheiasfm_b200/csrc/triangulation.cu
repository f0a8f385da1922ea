// Batched theia::TriangulateMidpoint (/root/reference/src/theia/sfm/triangulation/triangulation.cc:130-157): one thread per
// track accumulates A = sum (I - d d^T), b = sum (I - d d^T) o over the track's rays and solves the 4x4 SPD system by
// Cholesky (what Eigen::LLT<Matrix4d> does). The 4th row / column of A is (0, 0, 0, n): w = 1 up to rounding.
// HBM-bound: 48 B per ray in, 33 B per track out; rays of a track are contiguous, tracks of a warp are neighbours.
#include <vector>

#include "common.cuh"

namespace thb {
namespace {

__global__ void __launch_bounds__(128) k_triangulate_midpoint(int nt, const long long* __restrict__ off, const double* __restrict__ org,
                                                              const double* __restrict__ dir, double* __restrict__ out,
                                                              uint8_t* __restrict__ ok) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long long q0 = off[t], q1 = off[t + 1];
  double A[4][4] = {}, b[4] = {};
  for (long long q = q0; q < q1; ++q) {
    const double d[4] = {dir[3 * q], dir[3 * q + 1], dir[3 * q + 2], 0.0};
    const double o[4] = {org[3 * q], org[3 * q + 1], org[3 * q + 2], 1.0};
    double T[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) T[r][c] = (r == c ? 1.0 : 0.0) - d[r] * d[c];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) { A[r][c] += T[r][c]; s += T[r][c] * o[c]; }
      b[r] += s;
    }
  }
  bool good = (q1 - q0) >= 2;
  double L[4][4] = {};
#pragma unroll
  for (int k = 0; k < 4; ++k) {  // unblocked lower Cholesky, column by column
    double x = A[k][k];
#pragma unroll
    for (int j = 0; j < k; ++j) x -= L[k][j] * L[k][j];
    if (!(x > 0.0)) good = false;
    x = sqrt(x);
    L[k][k] = x;
#pragma unroll
    for (int r = k + 1; r < 4; ++r) {
      double v = A[r][k];
#pragma unroll
      for (int j = 0; j < k; ++j) v -= L[r][j] * L[k][j];
      L[r][k] = v / x;
    }
  }
  double y[4], z[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { double v = b[r]; for (int j = 0; j < r; ++j) v -= L[r][j] * y[j]; y[r] = v / L[r][r]; }
#pragma unroll
  for (int r = 3; r >= 0; --r) { double v = y[r]; for (int j = r + 1; j < 4; ++j) v -= L[j][r] * z[j]; z[r] = v / L[r][r]; }
#pragma unroll
  for (int r = 0; r < 4; ++r) out[4 * (size_t)t + r] = good ? z[r] : 0.0;
  if (ok) ok[t] = good ? 1 : 0;
}

}  // namespace
}  // namespace thb

extern "C" int thb_triangulate_midpoint_batch(const double* ray_origins, const double* ray_directions, const int64_t* ray_offset,
                                              int32_t num_tracks, int32_t memory_space, double* points_out, uint8_t* ok,
                                              void* cuda_stream) {
  using namespace thb;
  if (num_tracks < 0 || (memory_space != THB_MEM_HOST && memory_space != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  if (num_tracks == 0) return THB_OK;
  if (!ray_origins || !ray_directions || !ray_offset || !points_out) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  int n = 0, dev = 0, major = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) THB_FAIL(THB_E_NO_DEVICE, "no CUDA device visible; libtheia_b200 has no CPU path");
  cudaGetDevice(&dev);
  THB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) THB_FAIL(THB_E_NO_DEVICE, "device is not sm_100 (B200); kernels are built for sm_100a only");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int grid = (num_tracks + 127) / 128;
  if (memory_space == THB_MEM_DEVICE) {
    k_triangulate_midpoint<<<grid, 128, 0, st>>>(num_tracks, reinterpret_cast<const long long*>(ray_offset), ray_origins, ray_directions, points_out, ok);
    THB_CUDA_CHECK(cudaGetLastError());
    return THB_OK;
  }
  const long long total = ray_offset[num_tracks];
  if (ray_offset[0] != 0 || total < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "ray_offset must start at 0 and be non-decreasing");
  for (int t = 0; t < num_tracks; ++t)  // a malformed interior offset would send the kernel out of bounds (the device path trusts its caller)
    if (ray_offset[t + 1] < ray_offset[t] || ray_offset[t + 1] > total) THB_FAIL(THB_E_INVALID_ARGUMENT, "ray_offset must start at 0 and be non-decreasing");
  double *d_org = nullptr, *d_dir = nullptr, *d_out = nullptr; long long* d_off = nullptr; uint8_t* d_ok = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&d_org), sizeof(double) * 3 * (total ? total : 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&d_dir), sizeof(double) * 3 * (total ? total : 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&d_off), sizeof(long long) * (num_tracks + 1), st);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&d_out), sizeof(double) * 4 * num_tracks, st);
  if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&d_ok), num_tracks, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_org, ray_origins, sizeof(double) * 3 * total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_dir, ray_directions, sizeof(double) * 3 * total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_off, ray_offset, sizeof(long long) * (num_tracks + 1), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    k_triangulate_midpoint<<<grid, 128, 0, st>>>(num_tracks, d_off, d_org, d_dir, d_out, d_ok);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(points_out, d_out, sizeof(double) * 4 * num_tracks, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && ok) e = cudaMemcpyAsync(ok, d_ok, num_tracks, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFreeAsync(d_org, st); cudaFreeAsync(d_dir, st); cudaFreeAsync(d_off, st); cudaFreeAsync(d_out, st); cudaFreeAsync(d_ok, st);
  if (e != cudaSuccess) THB_FAIL(THB_E_CUDA, cudaGetErrorString(e));
  return THB_OK;
}
