// How fast can B200 absorb K1's output pattern (20 planes x 8 B per observation, 1M observations) and its input
// stream, without any arithmetic? Floors for the K1 roofline discussion. GPU box only.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_write_planes(double* __restrict__ out, int no, int nplanes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= no) return;
  const double v = i * 1e-9;
  for (int k = 0; k < nplanes; ++k) __stcs(out + (size_t)k * no + i, v + k);
}
__global__ void k_write_planes_v2(double2* __restrict__ out, int no, int npairs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= no) return;
  const double v = i * 1e-9;
  for (int k = 0; k < npairs; ++k) __stcs(out + (size_t)k * no + i, make_double2(v + k, v - k));
}
__global__ void k_rw(const int* __restrict__ a, const int* __restrict__ b, const double2* __restrict__ xy, const double2* __restrict__ si,
                     double* __restrict__ out, int no, int nplanes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= no) return;
  const double2 x = __ldcs(xy + i), s = __ldcs(si + i);
  const double v = x.x * s.x + x.y * s.y + __ldcs(a + i) + __ldcs(b + i);
  for (int k = 0; k < nplanes; ++k) __stcs(out + (size_t)k * no + i, v + k);
}
__global__ void k_flush(const double2* buf, size_t n, double* sink) {
  double acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { double2 v = __ldcg(buf + i); acc += v.x + v.y; }
  if (acc == 123.456) *sink = acc;
}
int main() {
  const int no = 1000000;
  double* out; cudaMalloc(&out, (size_t)no * 20 * 8);
  int *a, *b; double2 *xy, *si; cudaMalloc(&a, no * 4); cudaMalloc(&b, no * 4); cudaMalloc(&xy, no * 16); cudaMalloc(&si, no * 16);
  cudaMemset(a, 0, no * 4); cudaMemset(b, 0, no * 4); cudaMemset(xy, 0, no * 16); cudaMemset(si, 0, no * 16);
  void* fl; cudaMalloc(&fl, 256 << 20); cudaMemset(fl, 0, 256 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* label, auto launch, double bytes) {
    float tot = 0;
    for (int it = 0; it < 12; ++it) {
      k_flush<<<148 * 8, 256>>>((const double2*)fl, (256 << 20) / 16, out);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2) tot += ms;
    }
    printf("%-52s %7.2f us  %7.1f GB/s\n", label, 1e3 * tot / 10, bytes / (tot / 10 * 1e-3) / 1e9);
  };
  run("write 20 planes, STG.64 (160 MB)", [&] { k_write_planes<<<(no + 127) / 128, 128>>>(out, no, 20); }, 160e6);
  run("write 10 plane pairs, STG.128 (160 MB)", [&] { k_write_planes_v2<<<(no + 127) / 128, 128>>>((double2*)out, no, 10); }, 160e6);
  run("write 20 planes, 256-thread CTAs", [&] { k_write_planes<<<(no + 255) / 256, 256>>>(out, no, 20); }, 160e6);
  run("read obs stream (40 MB) + write 20 planes (160 MB)", [&] { k_rw<<<(no + 127) / 128, 128>>>(a, b, xy, si, out, no, 20); }, 200e6);
  printf("[%s]\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
