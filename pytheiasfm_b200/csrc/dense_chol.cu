// K4: FP64 dense Cholesky factor + solve of the reduced camera system S x = b on one B200.
//
// The reference hands S to Ceres (SPARSE_SCHUR / DENSE_SCHUR, bundle_adjuster.cc:65-88); for the
// headline configuration (1k cameras, every camera pair shares points) S is a dense 6000x6000 SPD
// matrix, i.e. the one GEMM-shaped piece of the path. Right-looking blocked factorisation, NB = 64:
//   diag   : one CTA factors the 64x64 diagonal block in shared memory and inverts the factor
//   panel  : L21 = A21 * inv(L11)^T as a small GEMM per 64-row tile
//   update : A22 -= L21 L21^T on the lower triangle with FP64 tensor-core MMAs (mma.sync m8n8k4;
//            tcgen05 has no FP64 kind), 64x64 tile per CTA
// The right-hand side rides along as one extra row of the matrix, so the forward substitution
// is a by-product of the panel/update steps; the backward substitution is one launch per block.
#include "dense_chol.cuh"

namespace thb {

namespace {

constexpr int NB = 64;
constexpr int LDS = NB + 4;  // padded shared-memory row stride (doubles): conflict-free fragments
constexpr int kSmem2 = 2 * NB * (NB + 1) * sizeof(double);

// ---- diag: factor A[k0:k0+64, k0:k0+64] in place (lower), write inv(L11) to dinv -----------------
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int ld, int k0,
                                                        double* __restrict__ dinv, int* __restrict__ fail) {
  extern __shared__ double dsm[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm);
  double (*inv)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + NB * (NB + 1));
  const int t = threadIdx.x;
  for (int e = t; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    s[i][j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  const int i = t & 63, jg = t >> 6;  // row, column group (4 groups)
  for (int k = 0; k < NB; ++k) {
    const double d = s[k][k];
    if (!(d > 0.0)) {  // not positive definite (or NaN)
      if (t == 0) atomicExch(fail, 1);
      return;
    }
    const double rd = 1.0 / sqrt(d);
    __syncthreads();
    if (t < NB && t >= k) s[t][k] = (t == k) ? sqrt(d) : s[t][k] * rd;
    __syncthreads();
    // trailing update of the lower triangle: s[i][j] -= s[i][k] s[j][k], k < j <= i
    const double lik = s[i][k];
    for (int j = k + 1 + jg; j <= i; j += 4) s[i][j] -= lik * s[j][k];
    __syncthreads();
  }
  // inverse of the lower-triangular factor: thread j (< 64) owns column j
  for (int e = t; e < NB * NB; e += 256) inv[e / NB][e % NB] = 0.0;
  __syncthreads();
  if (t < NB) {
    const int j = t;
    inv[j][j] = 1.0 / s[j][j];
    for (int r = j + 1; r < NB; ++r) {
      double acc = 0.0;
      for (int k = j; k < r; ++k) acc += s[r][k] * inv[k][j];
      inv[r][j] = -acc / s[r][r];
    }
  }
  __syncthreads();
  for (int e = t; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = s[r][c];
    dinv[e] = inv[r][c];
  }
}

// ---- panel: for each 64-row tile below the diagonal block, X = A21 * inv(L11)^T ------------------
__global__ void __launch_bounds__(256) chol_panel_kernel(double* __restrict__ A, int ld, int k0,
                                                         const double* __restrict__ dinv) {
  extern __shared__ double dsm[];
  double (*sa)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm);
  double (*si)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + NB * (NB + 1));
  const int r0 = k0 + NB + blockIdx.x * NB;
  const int t = threadIdx.x;
  for (int e = t; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    sa[i][j] = A[(size_t)(r0 + i) * ld + k0 + j];
    si[i][j] = dinv[e];
  }
  __syncthreads();
  // thread computes a 4x4 patch: rows 4*(t/16).., cols 4*(t%16)..
  const int pi = (t >> 4) * 4, pj = (t & 15) * 4;
  double acc[4][4] = {};
  for (int k = 0; k < NB; ++k) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { a[u] = sa[pi + u][k]; b[u] = si[pj + u][k]; }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] += a[u] * b[v];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) A[(size_t)(r0 + pi + u) * ld + k0 + pj + v] = acc[u][v];
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// ---- update: C[bi][bj] -= P[bi] * P[bj]^T over the lower-triangular tile set -------------------
// P = the freshly computed panel (columns k0..k0+63). One 64x64 tile per CTA, 4 warps, each warp a
// 32x32 quadrant = 4x4 m8n8k4 accumulators.
__global__ void __launch_bounds__(128) chol_update_kernel(double* __restrict__ A, int ld, int k0, int nt) {
  extern __shared__ double smem[];
  double* sa = smem;              // [64][LDS] rows of tile bi
  double* sb = smem + NB * LDS;   // [64][LDS] rows of tile bj
  // linear index -> (bi, bj), bj <= bi < nt
  const int lin = blockIdx.x;
  int bi = (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= lin) ++bi;
  while (bi * (bi + 1) / 2 > lin) --bi;
  const int bj = lin - bi * (bi + 1) / 2;
  const int r0 = k0 + NB + bi * NB, c0 = k0 + NB + bj * NB;
  const int t = threadIdx.x;
  for (int e = t; e < NB * (NB / 2); e += 128) {
    const int i = e / (NB / 2), j2 = (e % (NB / 2)) * 2;
    const double2 va = *reinterpret_cast<const double2*>(&A[(size_t)(r0 + i) * ld + k0 + j2]);
    sa[i * LDS + j2] = va.x; sa[i * LDS + j2 + 1] = va.y;
    const double2 vb = *reinterpret_cast<const double2*>(&A[(size_t)(c0 + i) * ld + k0 + j2]);
    sb[i * LDS + j2] = vb.x; sb[i * LDS + j2 + 1] = vb.y;
  }
  __syncthreads();
  const int w = t >> 5, lane = t & 31;
  const int wr = (w >> 1) * 32, wc = (w & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  double acc[4][4][2] = {};
#pragma unroll 4
  for (int k = 0; k < NB; k += 4) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a[u] = sa[(wr + u * 8 + fr) * LDS + k + fk];
      b[u] = sb[(wc + u * 8 + fr) * LDS + k + fk];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
  }
  // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = r0 + wr + u * 8 + fr, c = c0 + wc + v * 8 + 2 * fk;
      double2* p = reinterpret_cast<double2*>(&A[(size_t)r * ld + c]);
      double2 cv = *p;
      cv.x -= acc[u][v][0]; cv.y -= acc[u][v][1];
      *p = cv;
    }
}

// ---- backward substitution, one launch per 64-block from the bottom ---------------------------
// x_k = inv(L_kk)^T y_k (every CTA recomputes it; CTA kb stores it); CTA b < kb: y_b -= L[k, b]^T x_k
__global__ void __launch_bounds__(64) chol_backsolve_kernel(const double* __restrict__ A, int ld, int kb,
                                                            const double* __restrict__ dinv,
                                                            double* __restrict__ y) {
  __shared__ double xk[NB];
  __shared__ double yk[NB];
  const int t = threadIdx.x;
  const int k0 = kb * NB;
  yk[t] = y[k0 + t];
  __syncthreads();
  const double* di = dinv + (size_t)kb * NB * NB;
  double acc = 0.0;
  for (int r = t; r < NB; ++r) acc += di[r * NB + t] * yk[r];  // (inv^T y)_t = sum_r inv[r][t] y_r
  xk[t] = acc;
  __syncthreads();
  const int b = blockIdx.x;
  if (b == kb) { y[k0 + t] = acc; return; }
  double s = 0.0;
  for (int r = 0; r < NB; ++r) s += A[(size_t)(k0 + r) * ld + b * NB + t] * xk[r];
  y[b * NB + t] -= s;
}

__global__ void chol_pad_kernel(double* __restrict__ A, int ld, int n, int n_pad) {
  // rows/cols n..n_pad-1: identity; rhs row (n_pad) beyond n: zero
  const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) A[(size_t)i * ld + i] = 1.0;
}

__global__ void chol_copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace

size_t DenseChol::WorkspaceDoubles(int n) {
  const int n_pad = (n + NB - 1) / NB * NB;
  return (size_t)(n_pad + NB) * n_pad;
}

int DenseChol::Init(int n_) {
  n = n_;
  n_pad = (n + NB - 1) / NB * NB;
  ld = n_pad;
  nblk = n_pad / NB;
  THB_CUDA_CHECK(cudaMalloc(&A, sizeof(double) * (size_t)(n_pad + NB) * ld));
  THB_CUDA_CHECK(cudaMalloc(&dinv, sizeof(double) * (size_t)nblk * NB * NB));
  THB_CUDA_CHECK(cudaMalloc(&x, sizeof(double) * n_pad));
  THB_CUDA_CHECK(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
  THB_CUDA_CHECK(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem2));
  THB_CUDA_CHECK(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(2 * NB * LDS * sizeof(double))));
  return THB_OK;
}

void DenseChol::Free() {
  cudaFree(A); cudaFree(dinv); cudaFree(x);
  A = dinv = x = nullptr;
}

int DenseChol::Clear(cudaStream_t st) {
  THB_CUDA_CHECK(cudaMemsetAsync(A, 0, sizeof(double) * (size_t)(n_pad + NB) * ld, st));
  if (n_pad > n) chol_pad_kernel<<<(n_pad - n + 63) / 64, 64, 0, st>>>(A, ld, n, n_pad);
  return THB_OK;
}

// Factor the lower triangle of A (n_pad x n_pad, row n_pad = rhs) and leave the solution in x.
int DenseChol::FactorAndSolve(cudaStream_t st, int* fail_flag, int* launches) {
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb * NB;
    chol_diag_kernel<<<1, 256, kSmem2, st>>>(A, ld, k0, dinv + (size_t)kb * NB * NB, fail_flag);
    const int rows_below = nblk - kb - 1 + 1;  // + the rhs tile-row
    chol_panel_kernel<<<rows_below, 256, kSmem2, st>>>(A, ld, k0, dinv + (size_t)kb * NB * NB);
    *launches += 2;
    const int nt = nblk - kb - 1;  // square trailing tiles; the rhs tile-row is handled below
    if (nt > 0) {
      // tiles (bi, bj), bj <= bi < nt, plus the rhs tile-row bi = nt with bj < nt
      const int tiles = nt * (nt + 1) / 2 + nt;
      chol_update_kernel<<<tiles, 128, 2 * NB * LDS * sizeof(double), st>>>(A, ld, k0, nt + 1);
      *launches += 1;
    }
  }
  chol_copy_row_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(A + (size_t)n_pad * ld, x, n_pad);
  *launches += 1;
  for (int kb = nblk - 1; kb >= 0; --kb) {
    chol_backsolve_kernel<<<kb + 1, 64, 0, st>>>(A, ld, kb, dinv, x);
    *launches += 1;
  }
  THB_CUDA_CHECK(cudaGetLastError());
  return THB_OK;
}

}  // namespace thb
