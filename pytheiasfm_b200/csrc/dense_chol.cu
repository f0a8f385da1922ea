// K4: FP64 dense Cholesky factor + solve of the reduced camera system S x = b on one B200.
//
// The reference hands S to Ceres (SPARSE_SCHUR / DENSE_SCHUR, bundle_adjuster.cc:65-88); for the
// headline configuration (1k cameras, every camera pair shares points) S is a dense 6000x6000 SPD
// matrix, i.e. the one GEMM-shaped piece of the path. FP64 has no tcgen05 kind, so the GEMM part
// runs on the FP64 tensor-core path that exists on sm_100a: mma.sync m8n8k4 (DMMA).
//
// Right-looking, two-level blocking (inner panel 64, outer block 128):
//   diag   : one CTA factors a 64x64 diagonal block in shared memory (16x16 sub-blocks) and
//            builds inv(L11) block-wise
//   panel  : L21 = A21 * inv(L11)^T, a small GEMM per 64-row tile
//   strip  : after the first inner panel, only the next 64 columns are updated (K = 64)
//   update : after the second inner panel the whole trailing lower triangle is updated with
//            K = 128: 128x128 tile per CTA, 8 warps of 64x32 DMMA accumulators, operands streamed
//            through a 3-stage cp.async ring; every C element is read/written once per 128 columns
//            (16 flop per byte of C traffic)
// The right-hand side rides along as one extra (padded) tile-row of the matrix, so the forward
// substitution is a by-product; the backward substitution is ONE kernel: a CTA per 64-column
// block, waiting on per-block ready flags, x_b = inv(L_bb)^T (y_b - sum_{k>b} L_kb^T x_k).
#include "dense_chol.cuh"

#include <algorithm>
#include <mutex>

namespace thb {

namespace {

constexpr int NB = 64;    // inner panel width
constexpr int OB = 128;   // outer block / update tile
constexpr int SB = 16;    // sub-block of the diagonal kernel
constexpr int KC = 32;    // K chunk of the update pipeline
constexpr int LDK = KC + 4;  // padded smem row stride (doubles) -> conflict-free DMMA fragments
constexpr int STAGES = 3;
constexpr int kInvSmem = 2 * NB * (NB + 1) * sizeof(double);
constexpr int kBackSmem = NB * (NB + 1) * sizeof(double);

// 1/sqrt(d) with 7 FP64 instructions: fp32 MUFU seed + two Newton steps (46 -> 92 bits). FP64 SIMT issue is
// the bottleneck of the diagonal kernel (one warp-wide FP64 instruction occupies an SMSP for 8 cycles on B200).
__device__ __forceinline__ double fast_rsqrt(double d) {
  const float f = rsqrtf((float)d);
  if (!(f > 0.0f) || !(f < 3.0e38f)) return rsqrt(d);
  double y = (double)f;
  const double h = 0.5 * d;
  y = y * fma(-h, y * y, 1.5);
  y = y * fma(-h, y * y, 1.5);
  return y;
}

// ---- diag: factor A[k0:k0+64, k0:k0+64] in place (lower); rd[k] = 1 / L_kk ------------------------
// Thread (i = t % 64, g = t / 64) owns row i and the 16 columns j = 4*jj + g in registers. One pivot = one
// shared-memory publish of the raw pivot column + ONE barrier; every thread derives 1/sqrt(d) itself (cheaper
// than a second barrier, measured) and applies a[i][j] -= (c_i / d) c_j with one FMA per live column.
// Fully unrolled: static register indices, and the column loop of pivot k only covers the live columns.
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int ld, int k0,
                                                        double* __restrict__ rd, int* __restrict__ fail) {
  __shared__ double S[NB][NB + 1];
  __shared__ double colbuf[2][NB];
  const int t = threadIdx.x, i = t & 63, g = t >> 6;
  for (int e = t; e < NB * NB; e += 256) {  // coalesced tile load
    const int r = e >> 6, c = e & 63;
    S[r][c] = (c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
  }
  __syncthreads();
  double a[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) a[jj] = S[i][4 * jj + g];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int gk = k & 3, jk = k >> 2;
    if (g == gk) colbuf[k & 1][i] = a[jk];
    __syncthreads();
    const double d = colbuf[k & 1][k];
    if (!(d > 0.0)) {  // not positive definite (or NaN): uniform across the CTA
      if (t == 0) atomicExch(fail, 1);
      return;
    }
    const double rs = fast_rsqrt(d);
    const double ci = colbuf[k & 1][i];
    if (g == gk) {
      a[jk] = ci * rs;  // i == k: d * rsqrt(d) = sqrt(d)
      if (i == k) rd[k] = rs;
    }
    const double cid = ci * (rs * rs);  // c_i / d
#pragma unroll
    for (int jj = jk; jj < 16; ++jj) {
      const int j = 4 * jj + g;
      if (j > k && j <= i) a[jj] -= cid * colbuf[k & 1][j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) S[i][4 * jj + g] = a[jj];
  __syncthreads();
  for (int e = t; e < NB * NB; e += 256) {  // coalesced store of the lower triangle
    const int r = e >> 6, c = e & 63;
    if (c <= r) A[(size_t)(k0 + r) * ld + k0 + c] = S[r][c];
  }
}

// ---- panel: for each 64-row tile below the diagonal block, X L11^T = A21 -------------------------------
// 128 threads: all four warps move the tiles (coalesced), warps 0-1 solve with one row per lane in registers.
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ A, int ld, int k0,
                                                         const double* __restrict__ rd, int nrows_total) {
  extern __shared__ double psm[];
  double (*Lt)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm);                   // Lt[k][j] = L11[j][k]
  double (*Bs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm + NB * (NB + 1));
  __shared__ double rdg[NB];
  const int r0 = k0 + NB + blockIdx.x * NB;
  const int t = threadIdx.x;
  for (int e = t; e < NB * NB; e += 128) {
    const int i = e >> 6, j = e & 63;
    Lt[j][i] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
    Bs[i][j] = (r0 + i < nrows_total) ? A[(size_t)(r0 + i) * ld + k0 + j] : 0.0;
  }
  if (t < NB) rdg[t] = rd[t];
  __syncthreads();
  if (t < NB) {
    double b[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) b[j] = Bs[t][j];
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const double xv = b[k] * rdg[k];
      b[k] = xv;
#pragma unroll
      for (int j = k + 1; j < NB; ++j) b[j] -= xv * Lt[k][j];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) Bs[t][j] = b[j];
  }
  __syncthreads();
  for (int e = t; e < NB * NB; e += 128) {
    const int i = e >> 6, j = e & 63;
    if (r0 + i < nrows_total) A[(size_t)(r0 + i) * ld + k0 + j] = Bs[i][j];
  }
}

// ---- batched inverse of the 64x64 diagonal factors (off the critical path; used by the backward solve) --
__global__ void __launch_bounds__(256) chol_inverse_kernel(const double* __restrict__ A, int ld, double* __restrict__ dinv) {
  extern __shared__ double dsm[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm);
  double (*inv)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + NB * (NB + 1));
  const int t = threadIdx.x;
  const int k0 = blockIdx.x * NB;
  for (int e = t; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    s[i][j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
    inv[i][j] = 0.0;
  }
  __syncthreads();
  // diagonal 16x16 blocks: thread (block b = t/16, column j = t%16) solves L_bb x = e_j
  if (t < NB) {
    const int b = (t >> 4) * SB, j = t & 15;
    double x[SB];
#pragma unroll
    for (int r = 0; r < SB; ++r) {
      double acc = (r == j) ? 1.0 : 0.0;
#pragma unroll
      for (int u = 0; u < r; ++u) acc -= s[b + r][b + u] * x[u];
      x[r] = (r >= j) ? acc / s[b + r][b + r] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < SB; ++r) inv[b + r][b + j] = x[r];
  }
  __syncthreads();
  // off-diagonal blocks by increasing distance d: X_ij = -X_ii * (sum_{k=j}^{i-1} L_ik X_kj);
  // the strictly upper triangle of s is free and holds the intermediate sums (transposed)
  double (*tmp)[NB + 1] = s;
  for (int d = 1; d < NB / SB; ++d) {
    const int nb = NB / SB - d;
    for (int e = t; e < nb * SB * SB; e += 256) {
      const int bj = e / (SB * SB), r = (e / SB) % SB, c = e % SB;
      const int bi = bj + d;
      double acc = 0.0;
      for (int k = bj; k < bi; ++k)
#pragma unroll
        for (int u = 0; u < SB; ++u) acc += s[bi * SB + r][k * SB + u] * inv[k * SB + u][bj * SB + c];
      tmp[bj * SB + c][bi * SB + r] = acc;
    }
    __syncthreads();
    for (int e = t; e < nb * SB * SB; e += 256) {
      const int bj = e / (SB * SB), r = (e / SB) % SB, c = e % SB;
      const int bi = bj + d;
      double acc = 0.0;
#pragma unroll
      for (int u = 0; u < SB; ++u) acc += inv[bi * SB + r][bi * SB + u] * tmp[bj * SB + c][bi * SB + u];
      inv[bi * SB + r][bj * SB + c] = -acc;
    }
    __syncthreads();
  }
  double* out = dinv + (size_t)blockIdx.x * NB * NB;
  for (int e = t; e < NB * NB; e += 256) {
    const int r = e / NB, c = e % NB;
    out[e] = (c <= r) ? inv[r][c] : 0.0;
  }
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- update: C[r0:r0+128, c0:c0+BN] -= P[r0.., kc0:kc0+KT] * P[c0.., kc0:kc0+KT]^T ------------------
// BN = 128: square tiles over the lower-triangular tile set (linear block index -> (bi, bj <= bi)).
// BN = 64 : the one 64-column strip after the first inner panel (blockIdx.x = bi, bj = 0).
// 256 threads = 8 warps in a 2 x 4 grid, warp tile 64 x (BN/4).
template <int BN, int BM>
__global__ void __launch_bounds__(256, 1) chol_update_kernel(double* __restrict__ A, int ld, int kc0, int KT,
                                                            int row_base, int col_base, int nrows_total, int col_only) {
  extern __shared__ __align__(16) double smem[];
  constexpr int A_ELEMS = BM * LDK, B_ELEMS = BN * LDK, STAGE = A_ELEMS + B_ELEMS;
  constexpr int WGR = BM / 64;      // warp grid rows (each warp: 64 rows)
  constexpr int WGC = 8 / WGR;      // warp grid columns
  constexpr int WN = BN / WGC;      // warp tile columns
  constexpr int NV = WN / 8;        // b fragments per warp
  int bi, bj;
  if (!col_only) {
    const int lin = blockIdx.x;
    bi = (int)((sqrt(8.0 * lin + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= lin) ++bi;
    while (bi * (bi + 1) / 2 > lin) --bi;
    bj = lin - bi * (bi + 1) / 2;
  } else {
    bi = blockIdx.x; bj = 0;
  }
  const int r0 = row_base + bi * BM, c0 = col_base + bj * BN;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int wr = (w / WGC) * 64, wc = (w % WGC) * WN;
  const int fr = lane >> 2, fk = lane & 3;
  const int nchunks = KT / KC;

  auto load_stage = [&](int chunk, int stage) {
    double* sa = smem + stage * STAGE;
    double* sb = sa + A_ELEMS;
    const int kc = kc0 + chunk * KC;
    // rows of A tile: OB rows x KC doubles = OB*16 16-byte pieces; B tile: BN rows
    for (int e = t; e < (BM + BN) * (KC / 2); e += 256) {
      const int row = e / (KC / 2), piece = e % (KC / 2);
      if (row < BM) {
        int gr = r0 + row; if (gr >= nrows_total) gr = nrows_total - 1;
        cp_async16(sa + row * LDK + piece * 2, A + (size_t)gr * ld + kc + piece * 2);
      } else {
        const int rb = row - BM;
        cp_async16(sb + rb * LDK + piece * 2, A + (size_t)(c0 + rb) * ld + kc + piece * 2);
      }
    }
  };

  // accumulators start from C (loads overlap the first operand chunks), D = C + (-A) B^T
  double acc[8][NV][2];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    int r = r0 + wr + u * 8 + fr; if (r >= nrows_total) r = nrows_total - 1;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = c0 + wc + v * 8 + 2 * fk;
      const double2 cv = *reinterpret_cast<const double2*>(&A[(size_t)r * ld + c]);
      acc[u][v][0] = cv.x; acc[u][v][1] = cv.y;
    }
  }

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = ch + STAGES - 1;
    if (nxt < nchunks) load_stage(nxt, nxt % STAGES);
    cp_async_commit();
    const double* sa = smem + (ch % STAGES) * STAGE;
    const double* sb = sa + A_ELEMS;
#pragma unroll
    for (int k = 0; k < KC; k += 4) {
      double a[8], b[NV];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = -sa[(wr + u * 8 + fr) * LDK + k + fk];
#pragma unroll
      for (int v = 0; v < NV; ++v) b[v] = sb[(wc + v * 8 + fr) * LDK + k + fk];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    }
  }
  cp_async_wait<0>();
  // C fragment: row = lane/4, cols = 2*(lane%4) + {0,1}
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int r = r0 + wr + u * 8 + fr;
    if (r >= nrows_total) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = c0 + wc + v * 8 + 2 * fk;
      *reinterpret_cast<double2*>(&A[(size_t)r * ld + c]) = make_double2(acc[u][v][0], acc[u][v][1]);
    }
  }
}

// ---- backward substitution in one launch ---------------------------------------------------------
// CTA handles 64-column blocks b = nblk-1-blockIdx.x, then b - gridDim.x, ... (descending, so that a
// CTA never waits on a block owned by a CTA that is not yet resident).
//   x_b = inv(L_bb)^T ( y_b - sum_{k>b} L[k-block rows, b-block cols]^T x_k )
// ready[k] is set (release) once x_k has been written to y[k*64 ..).
__global__ void __launch_bounds__(256) chol_backsolve_kernel(const double* __restrict__ A, int ld, int nblk,
                                                             const double* __restrict__ dinv, double* __restrict__ y,
                                                             int* __restrict__ ready) {
  extern __shared__ double sdi[];  // inv(L_bb), 64 x 65
  __shared__ double xk[NB];
  __shared__ double part[4][NB];
  const int t = threadIdx.x, col = t & 63, rg = t >> 6;  // 4 row groups of 16 rows
  for (int b = nblk - 1 - blockIdx.x; b >= 0; b -= gridDim.x) {
    for (int e = t; e < NB * NB; e += 256) sdi[(e / NB) * (NB + 1) + (e % NB)] = dinv[(size_t)b * NB * NB + e];
    double acc = 0.0;
    for (int k = nblk - 1; k > b; --k) {
      // the L block does not depend on x: issue its loads before waiting for x_k
      const double* Lkb = A + (size_t)(k * NB + rg * 16) * ld + b * NB + col;
      double l[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) l[r] = Lkb[(size_t)r * ld];
      if (t == 0) {
        while (atomicAdd(&ready[k], 0) == 0) { __nanosleep(20); }
        __threadfence();
      }
      __syncthreads();
      const double* xs = y + k * NB + rg * 16;
#pragma unroll
      for (int r = 0; r < 16; ++r) acc += l[r] * __ldcg(xs + r);
    }
    part[rg][col] = acc;
    __syncthreads();
    if (t < NB) xk[t] = y[b * NB + t] - (part[0][t] + part[1][t] + part[2][t] + part[3][t]);
    __syncthreads();
    // (inv^T y)_c = sum_r inv[r][c] y_r, rows split over the 4 groups
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) v += sdi[(rg * 16 + r) * (NB + 1) + col] * xk[rg * 16 + r];
    part[rg][col] = v;
    __syncthreads();
    if (t < NB) y[b * NB + t] = part[0][t] + part[1][t] + part[2][t] + part[3][t];
    __threadfence();
    __syncthreads();
    if (t == 0) atomicExch(&ready[b], 1);
  }
}

__global__ void chol_pad_kernel(double* __restrict__ A, int ld, int n, int n_pad) {
  const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) A[(size_t)i * ld + i] = 1.0;
}

__global__ void chol_copy_row_kernel(const double* __restrict__ src, double* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace

int DenseChol::Init(int n_, cudaStream_t st) {
  n = n_;
  n_pad = (n + OB - 1) / OB * OB;
  ld = n_pad;
  nblk = n_pad / NB;
  rows_total = n_pad + 1;  // + the rhs row
  // stream-ordered allocations out of the device's default pool (kept warm by ConfigurePoolOnce in ba_solver.cu)
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&A), sizeof(double) * (size_t)rows_total * ld, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&dinv), sizeof(double) * (size_t)nblk * NB * NB, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&x), sizeof(double) * n_pad, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&ready), sizeof(int) * nblk, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&rdiag), sizeof(double) * n_pad, st));
  static std::once_flag attr_once[64];
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once[dev & 63], [&attr_err] {
    auto set = [&attr_err](const void* f, int bytes) {
      const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e != cudaSuccess) attr_err = e;
    };
    set((const void*)chol_inverse_kernel, kInvSmem);
    set((const void*)chol_panel_kernel, kInvSmem);
    set((const void*)chol_update_kernel<OB, OB>, (int)(STAGES * (OB + OB) * LDK * sizeof(double)));
    set((const void*)chol_update_kernel<OB, NB>, (int)(STAGES * (NB + OB) * LDK * sizeof(double)));
    set((const void*)chol_update_kernel<NB, NB>, (int)(STAGES * (NB + NB) * LDK * sizeof(double)));
  });
  THB_CUDA_CHECK(attr_err);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  num_sms = sms > 0 ? sms : 148;
  THB_CUDA_CHECK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming));
    THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_col[i], cudaEventDisableTiming));
  }
  return THB_OK;
}

void DenseChol::Free(cudaStream_t st) {
  if (A) cudaFreeAsync(A, st);
  if (dinv) cudaFreeAsync(dinv, st);
  if (x) cudaFreeAsync(x, st);
  if (ready) cudaFreeAsync(ready, st);
  if (rdiag) cudaFreeAsync(rdiag, st);
  A = dinv = x = rdiag = nullptr; ready = nullptr;
  if (s2) { cudaStreamDestroy(s2); s2 = nullptr; }
  if (ev_start) { cudaEventDestroy(ev_start); ev_start = nullptr; }
  for (int i = 0; i < 2; ++i) {
    if (ev_ready[i]) { cudaEventDestroy(ev_ready[i]); ev_ready[i] = nullptr; }
    if (ev_col[i]) { cudaEventDestroy(ev_col[i]); ev_col[i] = nullptr; }
  }
}

int DenseChol::Clear(cudaStream_t st) {
  THB_CUDA_CHECK(cudaMemsetAsync(A, 0, sizeof(double) * (size_t)rows_total * ld, st));
  if (n_pad > n) chol_pad_kernel<<<(n_pad - n + 63) / 64, 64, 0, st>>>(A, ld, n, n_pad);
  return THB_OK;
}

// One outer block (two inner panels) of panel work on stream `q`.
void DenseChol::PanelPair(cudaStream_t q, int ob, int* fail_flag, int* launches) {
  const int k0 = ob * OB;
  chol_diag_kernel<<<1, 256, 0, q>>>(A, ld, k0, rdiag + k0, fail_flag);
  // 64-row tiles below the diagonal block; the last tile holds only the rhs row (guarded)
  chol_panel_kernel<<<(n_pad - k0 - NB) / NB + 1, 128, kInvSmem, q>>>(A, ld, k0, rdiag + k0, rows_total);
  {  // strip: columns k0+64 .. k0+127, rows k0+64 .. end, K = 64
    const int rows = rows_total - (k0 + NB);
    const int tr = (rows + NB - 1) / NB;
    chol_update_kernel<NB, NB><<<tr, 256, STAGES * (NB + NB) * LDK * sizeof(double), q>>>(A, ld, k0, NB, k0 + NB, k0 + NB, rows_total, 1);
  }
  chol_diag_kernel<<<1, 256, 0, q>>>(A, ld, k0 + NB, rdiag + k0 + NB, fail_flag);
  chol_panel_kernel<<<(n_pad - k0 - OB) / NB + 1, 128, kInvSmem, q>>>(A, ld, k0 + NB, rdiag + k0 + NB, rows_total);
  *launches += 5;
}

// Factor the lower triangle of A (n_pad x n_pad, row n_pad = rhs) and leave the solution in x.
// Lookahead: the trailing update of outer block ob first updates only the next block's 128 columns; the
// next block's (latency-bound) diag/panel chain then runs on a second stream while the rest of the
// trailing matrix is updated on the caller's stream.
int DenseChol::FactorAndSolve(cudaStream_t st, int* fail_flag, int* launches) {
  const int nob = n_pad / OB;
  const size_t upd_smem = STAGES * (OB + OB) * LDK * sizeof(double);
  THB_CUDA_CHECK(cudaEventRecord(ev_start, st));
  THB_CUDA_CHECK(cudaStreamWaitEvent(s2, ev_start, 0));
  PanelPair(s2, 0, fail_flag, launches);
  THB_CUDA_CHECK(cudaEventRecord(ev_ready[0], s2));
  for (int ob = 0; ob < nob; ++ob) {
    const int k0 = ob * OB;
    THB_CUDA_CHECK(cudaStreamWaitEvent(st, ev_ready[ob & 1], 0));
    const int nt = nob - ob - 1;
    if (nt <= 0) break;
    // (A) next block's columns only: tiles (bi, 0), bi = 0..nt (bi = nt is the rhs row)
    chol_update_kernel<OB, NB><<<2 * nt + 1, 256, STAGES * (NB + OB) * LDK * sizeof(double), st>>>(A, ld, k0, OB, k0 + OB, k0 + OB, rows_total, 1);
    THB_CUDA_CHECK(cudaEventRecord(ev_col[ob & 1], st));
    // (B) next block's panel chain on the second stream
    THB_CUDA_CHECK(cudaStreamWaitEvent(s2, ev_col[ob & 1], 0));
    PanelPair(s2, ob + 1, fail_flag, launches);
    THB_CUDA_CHECK(cudaEventRecord(ev_ready[(ob + 1) & 1], s2));
    // (C) the rest of the trailing matrix: tiles (bi, bj) with bj >= 1
    const int ntr = nt - 1;
    if (ntr > 0) {
      const int tiles = ntr * (ntr + 1) / 2 + ntr;  // + the rhs tile-row
      chol_update_kernel<OB, OB><<<tiles, 256, upd_smem, st>>>(A, ld, k0, OB, k0 + 2 * OB, k0 + 2 * OB, rows_total, 0);
      *launches += 1;
    }
    *launches += 1;
  }
  chol_inverse_kernel<<<nblk, 256, kInvSmem, st>>>(A, ld, dinv);
  chol_copy_row_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(A + (size_t)n_pad * ld, x, n_pad);
  THB_CUDA_CHECK(cudaMemsetAsync(ready, 0, sizeof(int) * nblk, st));
  chol_backsolve_kernel<<<std::min(nblk, num_sms), 256, kBackSmem, st>>>(A, ld, nblk, dinv, x, ready);
  *launches += 3;
  THB_CUDA_CHECK(cudaGetLastError());
  return THB_OK;
}

}  // namespace thb
