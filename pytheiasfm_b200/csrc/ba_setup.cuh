// Device-side problem setup of the BA path: the two observation orders (point-major for the Schur elimination,
// camera-major for the camera blocks). Stands where the reference walks its hash maps to add one residual block per
// observation (bundle_adjuster.cc:116-173); here it is a histogram, a scan and a stable radix sort on the GPU.
#ifndef THB_BA_SETUP_CUH_
#define THB_BA_SETUP_CUH_

#include "common.cuh"

namespace thb {

// d_count: nkeys + 1 ints; on entry d_count[k] = number of observations with key k (d_count[nkeys] ignored), on exit the
// exclusive prefix sum (d_count[nkeys] = no). d_perm[q] = caller index of the q-th observation in key-major order; equal
// keys keep the caller's order (stable), so the summation order of every block is the caller's observation order.
int GroupByKey(const int* d_key, int no, int nkeys, int* d_count, int* d_perm, cudaStream_t st);
// Stable sort of (key, value) pairs by a 64-bit key.
int SortPairsU64(const unsigned long long* d_key_in, unsigned long long* d_key_out, const int* d_val_in, int* d_val_out, int n, cudaStream_t st);

}  // namespace thb
#endif
