// See ba_setup.cuh. CUB (part of the CUDA toolkit) provides the scan and the radix sort: setup plumbing, not a hot kernel.
#include "ba_setup.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace thb {

namespace {
__global__ void k_iota(int n, int* __restrict__ a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}
}  // namespace

int GroupByKey(const int* d_key, int no, int nkeys, int* d_count, int* d_perm, cudaStream_t st) {
  if (nkeys <= 0) return THB_OK;
  size_t scan_bytes = 0, sort_bytes = 0;
  int bits = 1;
  while ((1LL << bits) < nkeys) ++bits;
  int *d_keys_out = nullptr, *d_iota = nullptr;
  THB_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_count, d_count, nkeys + 1, st));
  THB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_key, d_keys_out, d_iota, d_perm, no, 0, bits, st));
  const size_t tmp_bytes = std::max(scan_bytes, sort_bytes) + 256;
  void* d_tmp = nullptr;
  THB_CUDA_CHECK(cudaMallocAsync(&d_tmp, tmp_bytes, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_keys_out), sizeof(int) * std::max(no, 1), st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&d_iota), sizeof(int) * std::max(no, 1), st));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(d_tmp, scan_bytes, d_count, d_count, nkeys + 1, st);
  if (e == cudaSuccess && no > 0) {
    k_iota<<<(no + 255) / 256, 256, 0, st>>>(no, d_iota);
    e = cub::DeviceRadixSort::SortPairs(d_tmp, sort_bytes, d_key, d_keys_out, d_iota, d_perm, no, 0, bits, st);
  }
  cudaFreeAsync(d_tmp, st); cudaFreeAsync(d_keys_out, st); cudaFreeAsync(d_iota, st);
  THB_CUDA_CHECK(e);
  return THB_OK;
}

int SortPairsU64(const unsigned long long* d_key_in, unsigned long long* d_key_out, const int* d_val_in, int* d_val_out, int n, cudaStream_t st) {
  if (n <= 0) return THB_OK;
  size_t bytes = 0;
  THB_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_key_in, d_key_out, d_val_in, d_val_out, n, 0, 64, st));
  void* d_tmp = nullptr;
  THB_CUDA_CHECK(cudaMallocAsync(&d_tmp, bytes + 256, st));
  const cudaError_t e = cub::DeviceRadixSort::SortPairs(d_tmp, bytes, d_key_in, d_key_out, d_val_in, d_val_out, n, 0, 64, st);
  cudaFreeAsync(d_tmp, st);
  THB_CUDA_CHECK(e);
  return THB_OK;
}

}  // namespace thb
