// Dense FP64 Cholesky solve of the reduced camera system (K4). See dense_chol.cu.
#ifndef THB_DENSE_CHOL_CUH_
#define THB_DENSE_CHOL_CUH_

#include "common.cuh"

namespace thb {

struct DenseChol {
  int n = 0, n_pad = 0, ld = 0, nblk = 0;
  double* A = nullptr;     // (n_pad + 64) x ld, row-major; lower triangle = S, row n_pad = rhs
  double* dinv = nullptr;  // nblk inverted 64x64 diagonal factors
  double* x = nullptr;     // n_pad solution

  static size_t WorkspaceDoubles(int n);
  int Init(int n);
  void Free();
  int Clear(cudaStream_t st);  // zero A, identity on the padding
  double* RhsRow() { return A + (size_t)n_pad * ld; }
  int FactorAndSolve(cudaStream_t st, int* fail_flag, int* launches);
};

}  // namespace thb
#endif
