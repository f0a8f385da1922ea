// Dense FP64 Cholesky solve of the reduced camera system (K4). See dense_chol.cu.
#ifndef THB_DENSE_CHOL_CUH_
#define THB_DENSE_CHOL_CUH_

#include <vector>

#include "common.cuh"

namespace thb {

struct DenseChol {
  int n = 0, n_pad = 0, ld = 0, nblk = 0, rows_total = 0, num_sms = 148;
  double* A = nullptr;     // (n_pad + 1) x ld, row-major; lower triangle = S, row n_pad = rhs
  double* dinv = nullptr;  // nblk inverted 64x64 diagonal factors
  double* rdiag = nullptr; // 1 / L_kk
  double* dscr = nullptr;  // two 64 x 64 copies of the next diagonal blocks (input of the fused diag + panel kernel)
  double* x = nullptr;     // n_pad solution
  ulonglong2* xtag = nullptr;  // backward substitution: x published as tagged words (value halves + the solve's epoch)
  unsigned epoch = 0;
  int* ll_sync = nullptr;  // task counter + progress counters of the left-looking tile kernel
  int* ll_cols = nullptr;  // first task index of every 64-column block (+ total)
  int ll_grid = 0, ll_sync_ints = 0;
  bool legacy = false;     // THB_K4_MODE=legacy: the r01 right-looking multi-launch schedule (kept for A/B timing)

  int Init(int n, cudaStream_t st);  // stream-ordered allocations
  void Free(cudaStream_t st);
  int Clear(cudaStream_t st);  // zero A, identity on the padding
  double* RhsRow() { return A + (size_t)n_pad * ld; }
  int FactorAndSolve(cudaStream_t st, int* fail_flag, int* launches);

 private:
  void PanelPair(cudaStream_t q, int ob, int* fail_flag, int* launches);
  int FactorLegacy(cudaStream_t st, int* fail_flag, int* launches);
  std::vector<int> h_cols;
  cudaStream_t s2 = nullptr;  // lookahead stream for the diag/panel chain
  cudaEvent_t ev_start = nullptr, ev_pp[4] = {nullptr, nullptr, nullptr, nullptr}, ev_c2[4] = {nullptr, nullptr, nullptr, nullptr};
};

}  // namespace thb
#endif
