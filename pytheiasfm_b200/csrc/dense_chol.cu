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
// block, polling the tagged words x_k is published in, x_b = inv(L_bb)^T (y_b - sum_{k>b} L_kb^T x_k).
#include "dense_chol.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace thb {

namespace {

#ifdef THB_K4_PROBE  // tools/microbench/k4_micro.cu: phase time stamps of CTA 0
__device__ long long g_probe[16];
#define THB_PROBE(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_probe[slot] = clock64(); } while (0)
#else
#define THB_PROBE(slot) do { } while (0)
#endif

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
    // rows 0..31 are finished once k >= 32: their two warps only keep the barrier count (FP64 issue is the limit)
    if (k < 32 || i >= 32) {
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

// ---- panel: for each 32-row tile below the diagonal block, X L11^T = A21 -------------------------------
// FP64 SIMT issue bounds this kernel (16 lanes/clk/SM), so a row is split over FOUR lanes (columns j = 4*jj + q) and a
// tile is only 32 rows: up to 188 CTAs spread the substitution over all SMs. x_k is produced by the lane that owns
// column k and broadcast inside the quad with one shuffle; everything else is thread-local registers.
constexpr int PR = 32;
constexpr int kPanelSmem = (NB * (NB + 1) + PR * (NB + 1)) * sizeof(double);
__global__ void __launch_bounds__(128) chol_panel_kernel(double* __restrict__ A, int ld, int k0,
                                                         const double* __restrict__ rd, int nrows_total) {
  extern __shared__ double psm[];
  double (*Lt)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm);                   // Lt[k][j] = L11[j][k]
  double (*Bs)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(psm + NB * (NB + 1));
  __shared__ double rdg[NB];
  const int r0 = k0 + NB + blockIdx.x * PR;
  const int t = threadIdx.x;
  for (int e = t; e < NB * NB; e += 128) {
    const int i = e >> 6, j = e & 63;
    Lt[j][i] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
  }
  for (int e = t; e < PR * NB; e += 128) {
    const int i = e >> 6, j = e & 63;
    Bs[i][j] = (r0 + i < nrows_total) ? A[(size_t)(r0 + i) * ld + k0 + j] : 0.0;
  }
  if (t < NB) rdg[t] = rd[t];
  __syncthreads();
  const int row = t >> 2, q = t & 3, lane = t & 31;
  double b[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) b[jj] = Bs[row][4 * jj + q];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    const int qk = k & 3, jk = k >> 2;
    double xv = b[jk] * rdg[k];  // meaningful in the owner lane (q == qk) only
    xv = __shfl_sync(0xffffffffu, xv, (lane & ~3) | qk);
    if (q == qk) b[jk] = xv;
    if (q > qk) b[jk] -= xv * Lt[k][4 * jk + q];
#pragma unroll
    for (int jj = jk + 1; jj < 16; ++jj) b[jj] -= xv * Lt[k][4 * jj + q];
  }
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) Bs[row][4 * jj + q] = b[jj];
  __syncthreads();
  for (int e = t; e < PR * NB; e += 128) {
    const int i = e >> 6, j = e & 63;
    if (r0 + i < nrows_total) A[(size_t)(r0 + i) * ld + k0 + j] = Bs[i][j];
  }
}

// ---- thread-per-row variants (the ones FactorAndSolve launches) ------------------------------------------------------
// Measured on B200 (tools/microbench/k4_micro.cu, tools/microbench/lat_micro.cu): a warp-wide DFMA issues every 2.1 cycles per SM
// sub-partition with 8.8 cycles latency, i.e. FP64 SIMT runs at the full 64 FMA/clk/SM; a publish + __syncthreads + read
// round trip costs 75-95 cycles, a quad shuffle 30, a branch ~25, and straight-line code that is executed once is
// instruction-fetch bound when the I-cache is cold. The r01-start kernels above spend 350-650 cycles per pivot on exactly
// those latencies (a loop-compacted variant of them, 950; removed). Here one THREAD owns one matrix row in registers, columns are handled in blocks of 8 (2 barriers per block
// instead of 1 per pivot, no shuffles), the register row is rotated by 8 per block so the loop body has static register
// indices and stays a few KB, and everything inside a block is independent FMAs that pipeline at full rate.

// MUFU.RSQ64H seed + one third-order step: 2.2e-16 relative error over the whole exponent range, ~65 cycles of latency
// (the fp32-seed version above: cvt + MUFU + cvt + two Newton steps = ~120).
__device__ __forceinline__ double rsqrt_f64(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}
__device__ __forceinline__ void ldg256(const double* p, double* o) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void stg256(double* p, const double* o) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(o[0]), "d"(o[1]), "d"(o[2]), "d"(o[3]) : "memory");
}

// diag: factor A[k0:k0+64, k0:k0+64] in place (lower); rd[k] = 1 / L_kk. 64 threads, thread i owns row i.
__global__ void __launch_bounds__(64) chol_diag3_kernel(double* __restrict__ A, int ld, int k0,
                                                        double* __restrict__ rd, int* __restrict__ fail) {
  __shared__ __align__(32) double Dblk[8][8];   // raw diagonal block of the current block column
  __shared__ __align__(32) double Lb[NB][8];    // the 8 finished L values of every row for the current block column
  const int i = threadIdx.x;
  THB_PROBE(0);
  double* grow = A + (size_t)(k0 + i) * ld + k0;
  double a[NB];
#pragma unroll
  for (int c = 0; c < NB; c += 4) ldg256(grow + c, a + c);  // entries right of the diagonal are never used
  THB_PROBE(1);
#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int k = 8 * kb;
    if ((i >> 3) == kb) {
#pragma unroll
      for (int u = 0; u < 8; ++u) Dblk[i & 7][u] = a[u];
    }
    __syncthreads();
    if (i >= k) {
      // every live thread factors the 8x8 block itself: no communication inside the block
      double D[8][8], rs[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double4 lo = *reinterpret_cast<const double4*>(&Dblk[v][0]);
        const double4 hi = *reinterpret_cast<const double4*>(&Dblk[v][4]);
        D[v][0] = lo.x; D[v][1] = lo.y; D[v][2] = lo.z; D[v][3] = lo.w; D[v][4] = hi.x; D[v][5] = hi.y; D[v][6] = hi.z; D[v][7] = hi.w;
      }
      bool bad = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double d = D[u][u];
        bad |= !(d > 0.0);
        rs[u] = rsqrt_f64(d);
#pragma unroll
        for (int v = u + 1; v < 8; ++v) D[v][u] *= rs[u];
#pragma unroll
        for (int v = u + 1; v < 8; ++v)
#pragma unroll
          for (int w = u + 1; w <= v; ++w) D[v][w] -= D[v][u] * D[w][u];
      }
      if (bad) {  // not positive definite (or NaN); identical in every thread, the factor is garbage from here on
        if (i == k) atomicExch(fail, 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) rs[u] = 0.0;
      }
      // own row of the block column: x L_kk^T = a[0..7]
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double x = a[u] * rs[u];
        a[u] = x;
#pragma unroll
        for (int v = u + 1; v < 8; ++v) a[v] -= x * D[v][u];
      }
      *reinterpret_cast<double4*>(&Lb[i][0]) = make_double4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<double4*>(&Lb[i][4]) = make_double4(a[4], a[5], a[6], a[7]);
      if (i == k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) rd[k + u] = rs[u];
      }
      // finished block column of this row -> global (only the lower triangle is ever written)
      if (i >= k + 8) {
        stg256(grow + k, a); stg256(grow + k + 4, a + 4);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) if (k + u <= i) grow[k + u] = a[u];
      }
    }
    __syncthreads();
    // trailing columns of this row: a[c] -= sum_u L_i,u * L_j,u for k + 8 <= j = k + c <= i
#pragma unroll
    for (int cg = 1; cg < 8; ++cg) {
      if (cg < 8 - kb && k + 8 * cg <= (i | 31)) {  // uniform per warp
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int j = k + 8 * cg + v;
          const double4 lo = *reinterpret_cast<const double4*>(&Lb[j][0]);
          const double4 hi = *reinterpret_cast<const double4*>(&Lb[j][4]);
          double acc = a[8 * cg + v];
          acc = fma(-a[0], lo.x, acc); acc = fma(-a[1], lo.y, acc); acc = fma(-a[2], lo.z, acc); acc = fma(-a[3], lo.w, acc);
          acc = fma(-a[4], hi.x, acc); acc = fma(-a[5], hi.y, acc); acc = fma(-a[6], hi.z, acc); acc = fma(-a[7], hi.w, acc);
          if (j <= i) a[8 * cg + v] = acc;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NB - 8; ++c) a[c] = a[c + 8];
    THB_PROBE(2 + kb);
  }
}

// panel: for each 64-row tile below the diagonal block, X L11^T = A21; thread-per-row, rows straight from / to global.
constexpr int PR3 = 64;
constexpr int LP = NB + 2;  // L11 row pitch in shared memory (16-byte aligned rows)
constexpr int kPanel3Smem = NB * LP * sizeof(double);
__global__ void __launch_bounds__(64) chol_panel3_kernel(double* __restrict__ A, int ld, int k0,
                                                         const double* __restrict__ rd, int nrows_total) {
  extern __shared__ __align__(16) double psm[];
  double (*L)[LP] = reinterpret_cast<double (*)[LP]>(psm);  // L[j][k] = L11[j][k]
  __shared__ double rdg[NB];
  const int t = threadIdx.x;
  THB_PROBE(0);
  const int r = k0 + NB + blockIdx.x * PR3 + t;
  const bool valid = r < nrows_total;
  // L11 -> shared memory: 2048 16-byte pieces, all in flight at once
#pragma unroll
  for (int m = 0; m < 32; ++m) {
    const int e = t + 64 * m, row = e >> 5, piece = e & 31;
    cp_async16(&L[row][2 * piece], A + (size_t)(k0 + row) * ld + k0 + 2 * piece);
  }
  cp_async_commit();
  rdg[t] = rd[t];
  double* grow = A + (size_t)(valid ? r : k0 + NB) * ld + k0;
  double b[NB];
#pragma unroll
  for (int c = 0; c < NB; c += 4) ldg256(grow + c, b + c);
  cp_async_wait<0>();
  __syncthreads();
  THB_PROBE(1);
#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int k = 8 * kb;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double x = b[u] * rdg[k + u];
      b[u] = x;
#pragma unroll
      for (int v = u + 1; v < 8; ++v) b[v] -= x * L[k + v][k + u];
    }
    if (valid) { stg256(grow + k, b); stg256(grow + k + 4, b + 4); }
#pragma unroll
    for (int cg = 1; cg < 8; ++cg) {
      if (cg < 8 - kb) {  // uniform
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const double* Lj = &L[k + 8 * cg + v][k];
          const double2 l0 = *reinterpret_cast<const double2*>(Lj), l1 = *reinterpret_cast<const double2*>(Lj + 2);
          const double2 l2 = *reinterpret_cast<const double2*>(Lj + 4), l3 = *reinterpret_cast<const double2*>(Lj + 6);
          double acc = b[8 * cg + v];
          acc = fma(-b[0], l0.x, acc); acc = fma(-b[1], l0.y, acc); acc = fma(-b[2], l1.x, acc); acc = fma(-b[3], l1.y, acc);
          acc = fma(-b[4], l2.x, acc); acc = fma(-b[5], l2.y, acc); acc = fma(-b[6], l3.x, acc); acc = fma(-b[7], l3.y, acc);
          b[8 * cg + v] = acc;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NB - 8; ++c) b[c] = b[c + 8];
    THB_PROBE(2 + kb);
  }
}

// ---- diag + panel in ONE launch (what FactorAndSolve uses) ---------------------------------------------------------
// Every CTA factors the 64 x 64 diagonal block itself (the same latency-bound 8-column-block loop as chol_diag3_kernel;
// redundant FP64 work is free here) while cp.async brings its own 64 panel rows into shared memory, then runs the
// chol_panel3_kernel substitution against the factor it holds in shared memory. CTA 0 alone stores the factored block
// and 1 / L_kk. Saves a dependent launch and an L11 round trip through global memory per inner panel.
constexpr int kDpSmem = 2 * NB * LP * sizeof(double);
// panel block step kb for one row held in b[] (rotated by 8 per step): in-block substitution, store, trailing update
__device__ __forceinline__ void dp_panel_step(int kb, double (&b)[NB], const double (*L)[LP], const double* rdg, double* prow,
                                              bool valid) {
  const int k = 8 * kb;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const double x = b[u] * rdg[k + u];
    b[u] = x;
#pragma unroll
    for (int v = u + 1; v < 8; ++v) b[v] -= x * L[k + v][k + u];
  }
  if (valid) { stg256(prow + k, b); stg256(prow + k + 4, b + 4); }
#pragma unroll
  for (int cg = 1; cg < 8; ++cg) {
    if (cg < 8 - kb) {  // uniform
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double* Lj = &L[k + 8 * cg + v][k];
        const double2 l0 = *reinterpret_cast<const double2*>(Lj), l1 = *reinterpret_cast<const double2*>(Lj + 2);
        const double2 l2 = *reinterpret_cast<const double2*>(Lj + 4), l3 = *reinterpret_cast<const double2*>(Lj + 6);
        double acc = b[8 * cg + v];
        acc = fma(-b[0], l0.x, acc); acc = fma(-b[1], l0.y, acc); acc = fma(-b[2], l1.x, acc); acc = fma(-b[3], l1.y, acc);
        acc = fma(-b[4], l2.x, acc); acc = fma(-b[5], l2.y, acc); acc = fma(-b[6], l3.x, acc); acc = fma(-b[7], l3.y, acc);
        b[8 * cg + v] = acc;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NB - 8; ++c) b[c] = b[c + 8];
}

// 128 threads, two roles on separate warps (one warp per SM sub-partition): threads 0-63 factor the diagonal block (one
// row each), threads 64-127 own one panel row each and run ONE BLOCK STEP BEHIND the factorisation - block column kb-1 of
// L11 is complete in shared memory after the second barrier of step kb-1 - so the panel costs one extra block step of
// latency instead of a second pass.
__global__ void __launch_bounds__(128) chol_dp_kernel(double* __restrict__ A, int ld, int k0, double* __restrict__ rd,
                                                      int* __restrict__ fail, int nrows_total,
                                                      const double* __restrict__ diag_src) {
  extern __shared__ __align__(16) double psm[];
  double (*L)[LP] = reinterpret_cast<double (*)[LP]>(psm);             // L11, filled block column by block column
  double (*Bt)[LP] = reinterpret_cast<double (*)[LP]>(psm + NB * LP);  // this CTA's panel rows (prefetched)
  __shared__ __align__(32) double Dblk[8][8];
  __shared__ double rdg[NB];
  const int t = threadIdx.x;
  const bool diag_role = t < NB;
  const int i = t & (NB - 1);  // diagonal-block row (diag role) or panel row within the tile (panel role)
  const bool writer = blockIdx.x == 0;
  const int r = k0 + NB + blockIdx.x * PR3 + i;
  const bool valid = r < nrows_total;
  {  // panel rows -> shared memory, in flight during the first block step
    const int r0 = k0 + NB + blockIdx.x * PR3;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int e = t + 128 * m, row = e >> 5, piece = e & 31;
      const int gr = min(r0 + row, nrows_total - 1);
      cp_async16(&Bt[row][2 * piece], A + (size_t)gr * ld + k0 + 2 * piece);
    }
    cp_async_commit();
  }
  double* grow = A + (size_t)(k0 + i) * ld + k0;
  double* prow = A + (size_t)(valid ? r : k0 + NB) * ld + k0;
  double a[NB];  // diag role: row i of the diagonal block; panel role: panel row i. Rotated by 8 per block step.
  if (diag_role) {
#pragma unroll
    for (int c = 0; c < NB; c += 4) ldg256(diag_src + i * NB + c, a + c);  // the copy, not A (see chol_update64_kernel)
  }
#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int k = 8 * kb;
    if (diag_role && (i >> 3) == kb) {
#pragma unroll
      for (int u = 0; u < 8; ++u) Dblk[i & 7][u] = a[u];
    }
    __syncthreads();
    if (!diag_role) {
      if (kb == 1) {  // the panel rows landed before the second barrier of step 0
#pragma unroll
        for (int c = 0; c < NB; c += 2) {
          const double2 v = *reinterpret_cast<const double2*>(&Bt[i][c]);
          a[c] = v.x; a[c + 1] = v.y;
        }
      }
      if (kb >= 1) dp_panel_step(kb - 1, a, L, rdg, prow, valid);
    } else if (i >= k) {
      double D[8][8], rs[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double4 lo = *reinterpret_cast<const double4*>(&Dblk[v][0]);
        const double4 hi = *reinterpret_cast<const double4*>(&Dblk[v][4]);
        D[v][0] = lo.x; D[v][1] = lo.y; D[v][2] = lo.z; D[v][3] = lo.w; D[v][4] = hi.x; D[v][5] = hi.y; D[v][6] = hi.z; D[v][7] = hi.w;
      }
      bool bad = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double d = D[u][u];
        bad |= !(d > 0.0);
        rs[u] = rsqrt_f64(d);
#pragma unroll
        for (int v = u + 1; v < 8; ++v) D[v][u] *= rs[u];
#pragma unroll
        for (int v = u + 1; v < 8; ++v)
#pragma unroll
          for (int w = u + 1; w <= v; ++w) D[v][w] -= D[v][u] * D[w][u];
      }
      if (bad) {  // not positive definite (or NaN); identical in every thread and CTA, the factor is garbage from here on
        if (writer && i == k) atomicExch(fail, 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) rs[u] = 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double x = a[u] * rs[u];
        a[u] = x;
#pragma unroll
        for (int v = u + 1; v < 8; ++v) a[v] -= x * D[v][u];
      }
#pragma unroll
      for (int u = 0; u < 8; u += 2) *reinterpret_cast<double2*>(&L[i][k + u]) = make_double2(a[u], a[u + 1]);
      if (i == k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { rdg[k + u] = rs[u]; if (writer) rd[k + u] = rs[u]; }
      }
      if (writer) {  // finished block column of this row -> global (only the lower triangle is ever written)
        if (i >= k + 8) {
          stg256(grow + k, a); stg256(grow + k + 4, a + 4);
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) if (k + u <= i) grow[k + u] = a[u];
        }
      }
    }
    if (kb == 0) cp_async_wait<0>();
    __syncthreads();
    if (diag_role) {
#pragma unroll
      for (int cg = 1; cg < 8; ++cg) {
        if (cg < 8 - kb && k + 8 * cg <= (i | 31)) {  // uniform per warp
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int j = k + 8 * cg + v;
            const double* Lj = &L[j][k];
            const double2 l0 = *reinterpret_cast<const double2*>(Lj), l1 = *reinterpret_cast<const double2*>(Lj + 2);
            const double2 l2 = *reinterpret_cast<const double2*>(Lj + 4), l3 = *reinterpret_cast<const double2*>(Lj + 6);
            double acc = a[8 * cg + v];
            acc = fma(-a[0], l0.x, acc); acc = fma(-a[1], l0.y, acc); acc = fma(-a[2], l1.x, acc); acc = fma(-a[3], l1.y, acc);
            acc = fma(-a[4], l2.x, acc); acc = fma(-a[5], l2.y, acc); acc = fma(-a[6], l3.x, acc); acc = fma(-a[7], l3.y, acc);
            if (j <= i) a[8 * cg + v] = acc;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NB - 8; ++c) a[c] = a[c + 8];
    }
  }
  if (!diag_role) dp_panel_step(7, a, L, rdg, prow, valid);  // last block column: no further barrier needed, L is complete
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
    bi = blockIdx.x; bj = blockIdx.y;
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

// ---- trailing update, main variant: 128 x 64 tile, 128 threads (2 x 2 warps of 64 x 32), TWO CTAs per SM ---------
// A lone 128x128 CTA per SM (above) spends a third of its life loading and storing its C tile with the tensor pipe idle
// (measured r01: 26.6 us per tile against 16.7 us of DMMA time). Two smaller co-resident CTAs overlap one's C traffic
// with the other's K loop: K chunk 16, 3-stage cp.async ring = 90 KB of shared memory. (NT = 256, four warps per SM
// sub-partition with 32 x 32 warp tiles, is 8 % faster in isolation - 23.9 vs 22.1 TFLOP/s - but fills the register file,
// so the critical-path kernels cannot slip in next to it: the whole factorisation gets 0.3 ms SLOWER. Measured.)
// Tiles cover the lower triangle in 128-row x 64-column units: row block bi owns column tiles bj = 0 .. 2*bi+1, the
// rhs tile-row (bi == row_tiles) the tiles 0 .. 2*row_tiles-1.
constexpr int KC2 = 16, LDK2 = KC2 + 4, U_BM = 128, U_BN = 64;
constexpr int kUpd2Smem = STAGES * (U_BM + U_BN) * LDK2 * sizeof(double);
template <bool LOAD_C = true, bool STORE_C = true, int NT = 128>
__global__ void __launch_bounds__(NT, 2) chol_update2_kernel(double* __restrict__ A, int ld, int kc0, int KT,
                                                            int base, int nrows_total) {
  // NT = 128: 2 x 2 warps of 64 x 32; NT = 256: 4 x 2 warps of 32 x 32 (half the accumulators per warp, twice the warps)
  constexpr int MU = NT == 128 ? 8 : 4;  // m8 fragments per warp
  extern __shared__ __align__(16) double smem[];
  constexpr int A_ELEMS = U_BM * LDK2, B_ELEMS = U_BN * LDK2, STAGE = A_ELEMS + B_ELEMS;
  constexpr int NV = 4;
  const int lin = blockIdx.x;
  int bi = (int)((sqrt(4.0 * lin + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) <= lin) ++bi;
  while (bi * (bi + 1) > lin) --bi;
  const int bj = lin - bi * (bi + 1);
  const int r0 = base + bi * U_BM, c0 = base + bj * U_BN;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int wr = (w >> 1) * (8 * MU), wc = (w & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const int nchunks = KT / KC2;

  auto load_stage = [&](int chunk, int stage) {
    double* sa = smem + stage * STAGE;
    double* sb = sa + A_ELEMS;
    const int kc = kc0 + chunk * KC2;
#pragma unroll
    for (int e = t; e < (U_BM + U_BN) * (KC2 / 2); e += NT) {
      const int row = e / (KC2 / 2), piece = e % (KC2 / 2);
      if (row < U_BM) {
        int gr = r0 + row; if (gr >= nrows_total) gr = nrows_total - 1;
        cp_async16(sa + row * LDK2 + piece * 2, A + (size_t)gr * ld + kc + piece * 2);
      } else {
        const int rb = row - U_BM;
        cp_async16(sb + rb * LDK2 + piece * 2, A + (size_t)(c0 + rb) * ld + kc + piece * 2);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  // accumulators start from C (the loads overlap the first operand chunks), D = C + (-A) B^T
  double acc[MU][NV][2];
#pragma unroll
  for (int u = 0; u < MU; ++u) {
    int r = r0 + wr + u * 8 + fr; if (r >= nrows_total) r = nrows_total - 1;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = c0 + wc + v * 8 + 2 * fk;
      if (LOAD_C) {
        const double2 cv = *reinterpret_cast<const double2*>(&A[(size_t)r * ld + c]);
        acc[u][v][0] = cv.x; acc[u][v][1] = cv.y;
      } else {
        acc[u][v][0] = 0.0; acc[u][v][1] = 0.0;
      }
    }
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
    for (int k = 0; k < KC2; k += 4) {
      double a[MU], b[NV];
#pragma unroll
      for (int u = 0; u < MU; ++u) a[u] = -sa[(wr + u * 8 + fr) * LDK2 + k + fk];
#pragma unroll
      for (int v = 0; v < NV; ++v) b[v] = sb[(wc + v * 8 + fr) * LDK2 + k + fk];
#pragma unroll
      for (int u = 0; u < MU; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int u = 0; u < MU; ++u) {
    const int r = r0 + wr + u * 8 + fr;
    if (r >= nrows_total) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = c0 + wc + v * 8 + 2 * fk;
      if (STORE_C || acc[u][v][0] == 1.2345e300) *reinterpret_cast<double2*>(&A[(size_t)r * ld + c]) = make_double2(acc[u][v][0], acc[u][v][1]);
    }
  }
}

// ---- block-column updates on the critical path (L(b) and the strip): 64 x 64 tiles, 8 warps of 32 x 16 ------------
// The generic kernel above gives a 64 x 64 tile to 8 warps of 64 x 8: nine shared-memory fragment loads per eight DMMAs.
// Here a warp owns 32 x 16 (six loads per eight DMMAs), the K chunk is 16 and three CTAs fit an SM.
// TM = 32 halves the tile (and its latency) for the second half of the factorisation, where a block column has fewer
// 64-row tiles than the GPU has SMs.
constexpr int S64 = 3;  // cp.async stages (5 stages measured no faster per launch and 0.2 ms slower overall: less co-residency)
constexpr int kUpd64Smem = S64 * (NB + NB) * LDK2 * sizeof(double);
template <int TM>
__global__ void __launch_bounds__(256, 3) chol_update64_kernel(double* __restrict__ A, int ld, int kc0, int KT,
                                                              int row_base, int col_base, int nrows_total,
                                                              double* __restrict__ diag_copy) {
  // diag_copy (64 x 64, row-major): second copy of the updated diagonal block A[row_base.., col_base..]. The fused
  // diag + panel kernel that follows reads the block from there, because its CTA 0 overwrites the block in A with the
  // factor while later-scheduled CTAs may not have read it yet.
  extern __shared__ __align__(16) double smem[];
  constexpr int A_ELEMS = TM * LDK2, STAGE = A_ELEMS + NB * LDK2;
  constexpr int MU = TM / 16, NV = 2;
  const int r0 = row_base + blockIdx.x * TM, c0 = col_base + blockIdx.y * NB;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int wr = (w >> 2) * (TM / 2), wc = (w & 3) * 16;
  const int fr = lane >> 2, fk = lane & 3;
  const int nchunks = KT / KC2;
  auto load_stage = [&](int chunk, int stage) {
    double* sa = smem + stage * STAGE;
    double* sb = sa + A_ELEMS;
    const int kc = kc0 + chunk * KC2;
#pragma unroll
    for (int e = t; e < (TM + NB) * (KC2 / 2); e += 256) {
      const int row = e / (KC2 / 2), piece = e % (KC2 / 2);
      if (row < TM) {
        int gr = r0 + row; if (gr >= nrows_total) gr = nrows_total - 1;
        cp_async16(sa + row * LDK2 + piece * 2, A + (size_t)gr * ld + kc + piece * 2);
      } else {
        const int rb = row - TM;
        cp_async16(sb + rb * LDK2 + piece * 2, A + (size_t)(c0 + rb) * ld + kc + piece * 2);
      }
    }
  };
#pragma unroll
  for (int s = 0; s < S64 - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  double acc[MU][NV][2];
#pragma unroll
  for (int u = 0; u < MU; ++u) {
    int r = r0 + wr + u * 8 + fr; if (r >= nrows_total) r = nrows_total - 1;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double2 cv = *reinterpret_cast<const double2*>(&A[(size_t)r * ld + c0 + wc + v * 8 + 2 * fk]);
      acc[u][v][0] = cv.x; acc[u][v][1] = cv.y;
    }
  }
  for (int ch = 0; ch < nchunks; ++ch) {
    cp_async_wait<S64 - 2>();
    __syncthreads();
    const int nxt = ch + S64 - 1;
    if (nxt < nchunks) load_stage(nxt, nxt % S64);
    cp_async_commit();
    const double* sa = smem + (ch % S64) * STAGE;
    const double* sb = sa + A_ELEMS;
#pragma unroll
    for (int k = 0; k < KC2; k += 4) {
      double a[MU], b[NV];
#pragma unroll
      for (int u = 0; u < MU; ++u) a[u] = -sa[(wr + u * 8 + fr) * LDK2 + k + fk];
#pragma unroll
      for (int v = 0; v < NV; ++v) b[v] = sb[(wc + v * 8 + fr) * LDK2 + k + fk];
#pragma unroll
      for (int u = 0; u < MU; ++u)
#pragma unroll
        for (int v = 0; v < NV; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int u = 0; u < MU; ++u) {
    const int r = r0 + wr + u * 8 + fr;
    if (r >= nrows_total) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = c0 + wc + v * 8 + 2 * fk;
      *reinterpret_cast<double2*>(&A[(size_t)r * ld + c]) = make_double2(acc[u][v][0], acc[u][v][1]);
      if (diag_copy && r < row_base + NB && c < col_base + NB)
        *reinterpret_cast<double2*>(&diag_copy[(r - row_base) * NB + (c - col_base)]) = make_double2(acc[u][v][0], acc[u][v][1]);
    }
  }
}

// ---- left-looking tile Cholesky in ONE persistent kernel (what FactorAndSolve runs) ------------------------------------
// The r01 schedule (above) is right-looking: ~230 dependent launches on two streams, every 128-column step re-reads and
// re-writes the whole trailing matrix, and for the last 55 % of the block columns the factorisation waits on the serial
// diag -> panel -> strip chain (65 us per 128 columns) while the bulk kernels have almost nothing left to do.
// Here the unit of work is a TILE TASK (J, m): 128 rows x 64 columns of block column J, rows starting at the diagonal
// (row blocks J + 2m, J + 2m + 1 of 64 rows). A task keeps its tile in DMMA accumulators and runs the whole left-looking
// update  C -= L[rows, 0:64J] L[J-block rows, 0:64J]^T  (one long K loop: the tile is read once and written once, the
// operands stream through a 3-stage cp.async ring), then turns it into a tile of L:
//   m == 0 : the top 64 x 64 is the diagonal block: Cholesky factor + inverse in shared memory (thread-per-row, 8-column
//            blocks; the inverse rows are produced one block step behind by the second half of the CTA), inv(L_JJ) is
//            published (dinv, also what the backward substitution uses), then the lower 64 rows are multiplied by inv^T
//   m  > 0 : wait for inv(L_JJ), X = C inv(L_JJ)^T with DMMA from shared memory (a GEMM instead of a substitution)
// Tasks are handed out by one atomic counter in column-major order to a grid that is exactly co-resident (2 CTAs per SM),
// and every dependency of a task is an EARLIER task, so spinning on the progress counters cannot deadlock:
//   rb_done[rb]  = number of leading block columns in which row block rb is finished (single writer at a time, monotone)
//   diag_done    = number of diagonal blocks factored and inverted
// A task polls the three counters of its row blocks (its own two and row block J, the B operand) a few chunks ahead of
// the loads that need them, so columns J+1, J+2, ... run their K loops while the chain of column J completes: the chain
// (last update of the diagonal tile, factor + inverse, multiply, publish: ~20 us per 64 columns) only shows where a column
// has less than that much work, i.e. in the first and last ~12 of the 94 block columns.
constexpr int TP = NB + 4;                 // shared-memory pitch of the tile / inverse copies: conflict-free DMMA fragments
constexpr int LL_T = U_BM * TP;            // tile copy (doubles)
constexpr int LL_LI = NB * TP;             // inverse of the diagonal factor
constexpr int kLlSmem = (LL_T + LL_LI) * sizeof(double);
static_assert(kLlSmem >= kUpd2Smem, "the operand ring aliases the tile copy");
constexpr int LL_TASK = 0, LL_DIAG = 1, LL_TIMEOUT = 2, LL_RB = 8;  // int slots of the sync block

__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// inverse role, block step kb: row i of X = L^-T (X L^T = I) held in b[] (rotated by 8 per step); writes inv(L)[k+u][i]
__device__ __forceinline__ void ll_inv_step(int kb, double (&b)[NB], const double* __restrict__ L, const double* rdg,
                                            double* __restrict__ Li, double* __restrict__ dinv_g, int i) {
  const int k = 8 * kb;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const double x = b[u] * rdg[k + u];
    b[u] = x;
#pragma unroll
    for (int v = u + 1; v < 8; ++v) b[v] -= x * L[(k + v) * TP + k + u];
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) { Li[(k + u) * TP + i] = b[u]; dinv_g[(k + u) * NB + i] = b[u]; }
#pragma unroll
  for (int cg = 1; cg < 8; ++cg) {
    if (cg < 8 - kb) {  // uniform
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double* Lj = &L[(k + 8 * cg + v) * TP + k];
        const double2 l0 = *reinterpret_cast<const double2*>(Lj), l1 = *reinterpret_cast<const double2*>(Lj + 2);
        const double2 l2 = *reinterpret_cast<const double2*>(Lj + 4), l3 = *reinterpret_cast<const double2*>(Lj + 6);
        double acc = b[8 * cg + v];
        acc = fma(-b[0], l0.x, acc); acc = fma(-b[1], l0.y, acc); acc = fma(-b[2], l1.x, acc); acc = fma(-b[3], l1.y, acc);
        acc = fma(-b[4], l2.x, acc); acc = fma(-b[5], l2.y, acc); acc = fma(-b[6], l3.x, acc); acc = fma(-b[7], l3.y, acc);
        b[8 * cg + v] = acc;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NB - 8; ++c) b[c] = b[c + 8];
}

// Factor the 64 x 64 block in rows 0..63 of the tile copy T (pitch TP; only its lower triangle is read) IN PLACE and
// write inv(L) to Li (pitch TP, full 64 x 64 with explicit zeros) and dinv_g (64 x 64 row-major). 128 threads:
// t < 64 own one row of the block each (the loop of chol_diag3_kernel), t >= 64 own one row of L^-T each and run one
// block step behind. Called by all 128 threads of the CTA.
__device__ void ll_factor_inverse(double* __restrict__ T, double* __restrict__ Li, double* __restrict__ dinv_g, int* fail,
                                  double (*Dblk)[8], double* rdg) {
  const int t = threadIdx.x;
  const bool diag_role = t < NB;
  const int i = t & (NB - 1);
  double a[NB];
  if (diag_role) {
#pragma unroll
    for (int c = 0; c < NB; c += 2) { const double2 v = *reinterpret_cast<const double2*>(&T[i * TP + c]); a[c] = v.x; a[c + 1] = v.y; }
  }
  __syncthreads();  // every row is in registers: rows 0..63 of T now hold L, block column by block column
#pragma unroll 1
  for (int kb = 0; kb < 8; ++kb) {
    const int k = 8 * kb;
    if (diag_role && (i >> 3) == kb) {
#pragma unroll
      for (int u = 0; u < 8; ++u) Dblk[i & 7][u] = a[u];
    }
    __syncthreads();
    if (!diag_role) {
      if (kb == 1) {
#pragma unroll
        for (int c = 0; c < NB; ++c) a[c] = (c == i) ? 1.0 : 0.0;
      }
      if (kb >= 1) ll_inv_step(kb - 1, a, T, rdg, Li, dinv_g, i);
    } else if (i >= k) {
      double D[8][8], rs[8];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const double4 lo = *reinterpret_cast<const double4*>(&Dblk[v][0]);
        const double4 hi = *reinterpret_cast<const double4*>(&Dblk[v][4]);
        D[v][0] = lo.x; D[v][1] = lo.y; D[v][2] = lo.z; D[v][3] = lo.w; D[v][4] = hi.x; D[v][5] = hi.y; D[v][6] = hi.z; D[v][7] = hi.w;
      }
      bool bad = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double d = D[u][u];
        bad |= !(d > 0.0);
        rs[u] = rsqrt_f64(d);
#pragma unroll
        for (int v = u + 1; v < 8; ++v) D[v][u] *= rs[u];
#pragma unroll
        for (int v = u + 1; v < 8; ++v)
#pragma unroll
          for (int w = u + 1; w <= v; ++w) D[v][w] -= D[v][u] * D[w][u];
      }
      if (bad) {  // not positive definite (or NaN); identical in every thread, the factor is garbage from here on
        if (i == k) atomicExch(fail, 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) rs[u] = 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double x = a[u] * rs[u];
        a[u] = x;
#pragma unroll
        for (int v = u + 1; v < 8; ++v) a[v] -= x * D[v][u];
      }
#pragma unroll
      for (int u = 0; u < 8; u += 2) *reinterpret_cast<double2*>(&T[i * TP + k + u]) = make_double2(a[u], a[u + 1]);
      if (i == k) {
#pragma unroll
        for (int u = 0; u < 8; ++u) rdg[k + u] = rs[u];
      }
    }
    __syncthreads();
    if (diag_role) {
#pragma unroll
      for (int cg = 1; cg < 8; ++cg) {
        if (cg < 8 - kb && k + 8 * cg <= (i | 31)) {  // uniform per warp
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            const int j = k + 8 * cg + v;
            const double* Lj = &T[j * TP + k];
            const double2 l0 = *reinterpret_cast<const double2*>(Lj), l1 = *reinterpret_cast<const double2*>(Lj + 2);
            const double2 l2 = *reinterpret_cast<const double2*>(Lj + 4), l3 = *reinterpret_cast<const double2*>(Lj + 6);
            double acc = a[8 * cg + v];
            acc = fma(-a[0], l0.x, acc); acc = fma(-a[1], l0.y, acc); acc = fma(-a[2], l1.x, acc); acc = fma(-a[3], l1.y, acc);
            acc = fma(-a[4], l2.x, acc); acc = fma(-a[5], l2.y, acc); acc = fma(-a[6], l3.x, acc); acc = fma(-a[7], l3.y, acc);
            if (j <= i) a[8 * cg + v] = acc;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NB - 8; ++c) a[c] = a[c + 8];
    }
  }
  if (!diag_role) ll_inv_step(7, a, T, rdg, Li, dinv_g, i);
}

// X = T[rows] inv(L)^T for 8 * MF rows per warp (all 64 columns), straight from shared memory; stores the valid rows to A.
template <int MF>
__device__ __forceinline__ void ll_multiply_store(const double* __restrict__ T, const double* __restrict__ Li, int tile_row0,
                                                  double* __restrict__ A, int ld, int grow0, int c0, int rows_total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int fr = lane >> 2, fk = lane & 3;
  const int rw = tile_row0 + w * (8 * MF);
  double x[MF][8][2];
#pragma unroll
  for (int u = 0; u < MF; ++u)
#pragma unroll
    for (int v = 0; v < 8; ++v) { x[u][v][0] = 0.0; x[u][v][1] = 0.0; }
#pragma unroll
  for (int k = 0; k < NB; k += 4) {
    double a[MF];
#pragma unroll
    for (int u = 0; u < MF; ++u) a[u] = T[(rw + u * 8 + fr) * TP + k + fk];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      if (8 * v + 7 >= k) {  // inv(L) is lower triangular: column block v only sees k <= 8 v + 7
        const double b = Li[(v * 8 + fr) * TP + k + fk];
#pragma unroll
        for (int u = 0; u < MF; ++u) dmma_m8n8k4(x[u][v][0], x[u][v][1], a[u], b);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < MF; ++u) {
    const int r = grow0 + rw + u * 8 + fr;
    if (r >= rows_total) continue;
#pragma unroll
    for (int v = 0; v < 8; ++v)
      *reinterpret_cast<double2*>(&A[(size_t)r * ld + c0 + v * 8 + 2 * fk]) = make_double2(x[u][v][0], x[u][v][1]);
  }
}

__global__ void __launch_bounds__(128, 2) chol_ll_kernel(double* __restrict__ A, int ld, int nb64, int rows_total,
                                                         const int* __restrict__ col_task_start, int* __restrict__ sync,
                                                         double* __restrict__ dinv, int* __restrict__ fail,
                                                         long long* __restrict__ prof) {
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(32) double Dblk[8][8];
  __shared__ double rdg[NB];
  __shared__ int s_task;
  constexpr int A_ELEMS = U_BM * LDK2, B_ELEMS = U_BN * LDK2, STAGE = A_ELEMS + B_ELEMS;
  constexpr int MU = 8, NV = 4;
  double* T = smem;
  double* Li = smem + LL_T;
  int* rb_done = sync + LL_RB;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int wr = (w >> 1) * 64, wc = (w & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const int nrb = (rows_total + NB - 1) / NB;
  const int total_tasks = col_task_start[nb64];
  constexpr int kSpinLimit = 1 << 22;

  auto stamp = [&](int J_, int m_, int slot) {  // THB_K4_PROF: global-timer stamps of the first two tiles of every column
    if (prof && t == 0 && m_ < 2) { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); prof[(J_ * 2 + m_) * 8 + slot] = g; }
  };
  while (true) {
    __syncthreads();  // the previous task is done with shared memory
    if (t == 0) s_task = atomicAdd(&sync[LL_TASK], 1);
    __syncthreads();
    const int task = s_task;
    if (task >= total_tasks) break;
    int J = 0;
    {
      int lo = 0, hi = nb64;  // col_task_start[lo] <= task < col_task_start[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (col_task_start[mid] <= task) lo = mid; else hi = mid; }
      J = lo;
    }
    const int m = task - col_task_start[J];
    const int rb0 = J + 2 * m;
    const int r0 = rb0 * NB, c0 = J * NB;
    const int nch = 4 * J;  // K chunks of 16 columns

    auto load_stage = [&](int chunk, int stage) {
      double* sa = smem + stage * STAGE;
      double* sb = sa + A_ELEMS;
      const int kc = chunk * KC2;
#pragma unroll
      for (int e = t; e < (U_BM + U_BN) * (KC2 / 2); e += 128) {
        const int row = e / (KC2 / 2), piece = e % (KC2 / 2);
        if (row < U_BM) {
          int gr = r0 + row; if (gr >= rows_total) gr = rows_total - 1;
          cp_async16(sa + row * LDK2 + piece * 2, A + (size_t)gr * ld + kc + piece * 2);
        } else {
          const int rb = row - U_BM;
          cp_async16(sb + rb * LDK2 + piece * 2, A + (size_t)(c0 + rb) * ld + kc + piece * 2);
        }
      }
    };
    // readiness of block columns [0, known) for the three row blocks this task reads (thread 0 only)
    int known = 0, fenced = 0, p0 = 0, p1 = 0, p2 = 0, spins = 0;
    bool pf = false;
    auto poll = [&]() {
      const int a = ld_relaxed(&rb_done[J]), b = ld_relaxed(&rb_done[rb0]);
      const int c = rb0 + 1 < nrb ? ld_relaxed(&rb_done[rb0 + 1]) : J;
      return min(min(a, b), min(c, J));
    };
    auto wait_for = [&](int kb_needed) {  // CTA-uniform: returns when block column kb_needed is ready
      bool stall = false;
      if (t == 0) {
        if (kb_needed >= known) known = max(known, poll());
        stall = kb_needed >= known;
        if (!stall && known != fenced) { __threadfence(); fenced = known; }
      }
      while (__syncthreads_or(stall)) {
        if (t == 0) {
          __nanosleep(m <= 1 ? 40 : 400);
          known = max(known, poll());
          stall = kb_needed >= known;
          if (stall && ++spins > kSpinLimit) { atomicExch(&sync[LL_TIMEOUT], 1); known = J; stall = false; }  // never hang the GPU
          if (!stall) { __threadfence(); fenced = known; }
        }
      }
    };

    stamp(J, m, 0);
    if (nch > 0) wait_for(0);
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < nch) load_stage(s, s);
      cp_async_commit();
    }
    // accumulators start from the tile itself (the loads overlap the first operand chunks)
    double acc[MU][NV][2];
#pragma unroll
    for (int u = 0; u < MU; ++u) {
      int r = r0 + wr + u * 8 + fr; if (r >= rows_total) r = rows_total - 1;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double2 cv = *reinterpret_cast<const double2*>(&A[(size_t)r * ld + c0 + wc + v * 8 + 2 * fk]);
        acc[u][v][0] = cv.x; acc[u][v][1] = cv.y;
      }
    }
    auto compute_chunk = [&](int ch) {
      const double* sa = smem + (ch % STAGES) * STAGE;
      const double* sb = sa + A_ELEMS;
#pragma unroll
      for (int k = 0; k < KC2; k += 4) {
        double a[MU], b[NV];
#pragma unroll
        for (int u = 0; u < MU; ++u) a[u] = -sa[(wr + u * 8 + fr) * LDK2 + k + fk];
#pragma unroll
        for (int v = 0; v < NV; ++v) b[v] = sb[(wc + v * 8 + fr) * LDK2 + k + fk];
#pragma unroll
        for (int u = 0; u < MU; ++u)
#pragma unroll
          for (int v = 0; v < NV; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
      }
    };
    for (int ch = 0; ch < nch; ++ch) {
      cp_async_wait<STAGES - 2>();
      const int nxt = ch + STAGES - 1;
      if (t == 0 && pf && (ch & 3) == 2) { known = max(known, min(min(p0, p1), min(p2, J))); pf = false; }
      bool late = false;  // CTA-uniform: the block column of chunk nxt is not ready yet
      if (nxt < nch && (nxt & 3) == 0) {
        bool stall = false;
        if (t == 0) {
          const int kb = nxt >> 2;
          if (kb >= known) known = max(known, poll());
          stall = kb >= known;
          if (!stall && known != fenced) { __threadfence(); fenced = known; }
        }
        late = __syncthreads_or(stall);  // also the barrier of this chunk
      } else {
        __syncthreads();
      }
      if (late) {  // use the wait: the chunks already in the ring do not depend on it
        compute_chunk(ch);
        wait_for(nxt >> 2);
        load_stage(nxt, nxt % STAGES);
        cp_async_commit();
        continue;
      }
      if (nxt < nch) load_stage(nxt, nxt % STAGES);
      cp_async_commit();
      if (t == 0 && (ch & 3) == 0 && known < J) {  // look ahead: consumed two chunks later
        p0 = ld_relaxed(&rb_done[J]); p1 = ld_relaxed(&rb_done[rb0]); p2 = rb0 + 1 < nrb ? ld_relaxed(&rb_done[rb0 + 1]) : J;
        pf = true;
      }
      compute_chunk(ch);
    }
    cp_async_wait<0>();
    __syncthreads();  // the ring is free: the tile copy aliases it
    stamp(J, m, 1);
#pragma unroll
    for (int u = 0; u < MU; ++u)
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<double2*>(&T[(wr + u * 8 + fr) * TP + wc + v * 8 + 2 * fk]) = make_double2(acc[u][v][0], acc[u][v][1]);
    __syncthreads();
    double* dinv_g = dinv + (size_t)J * NB * NB;
    if (m == 0) {
      stamp(J, m, 2);
      ll_factor_inverse(T, Li, dinv_g, fail, Dblk, rdg);
      stamp(J, m, 3);
      __threadfence();
      __syncthreads();
      if (t == 0) st_release(&sync[LL_DIAG], J + 1);  // inv(L_JJ) is published: the rest of the column can finish
      stamp(J, m, 4);
      ll_multiply_store<2>(T, Li, NB, A, ld, r0, c0, rows_total);
      stamp(J, m, 5);
    } else {
      bool stall = false;
      if (t == 0) { stall = ld_relaxed(&sync[LL_DIAG]) <= J; if (!stall) __threadfence(); }
      while (__syncthreads_or(stall)) {
        if (t == 0) {
          __nanosleep(m <= 1 ? 40 : 400);
          stall = ld_relaxed(&sync[LL_DIAG]) <= J;
          if (stall && ++spins > kSpinLimit) { atomicExch(&sync[LL_TIMEOUT], 1); stall = false; }
          if (!stall) __threadfence();
        }
      }
#pragma unroll
      for (int q = 0; q < (NB * NB / 2) / 128; ++q) {
        const int e = t + 128 * q, row = e >> 5, piece = e & 31;
        cp_async16(&Li[row * TP + 2 * piece], dinv_g + row * NB + 2 * piece);
      }
      stamp(J, m, 2);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      stamp(J, m, 3);
      // top row block first: the next column's diagonal tile is waiting for it when this is tile m = 1
      ll_multiply_store<2>(T, Li, 0, A, ld, r0, c0, rows_total);
      __threadfence();
      __syncthreads();
      if (t == 0) st_release(&rb_done[rb0], J + 1);
      ll_multiply_store<2>(T, Li, NB, A, ld, r0, c0, rows_total);
      stamp(J, m, 5);
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
      if (m == 0) st_release(&rb_done[rb0], J + 1);
      if (rb0 + 1 < nrb) st_release(&rb_done[rb0 + 1], J + 1);
    }
    stamp(J, m, 6);
  }
}

// ---- backward substitution in one launch ---------------------------------------------------------
// CTA handles 64-column blocks b = nblk-1-blockIdx.x, then b - gridDim.x, ... (descending, so that a
// CTA never waits on a block owned by a CTA that is not yet resident).
//   x_b = inv(L_bb)^T ( y_b - sum_{k>b} L[k-block rows, b-block cols]^T x_k )
// The 94 blocks are a serial chain, so the hand-off is what is timed: x_k is published as TAGGED words (each 32-bit half of
// a value travels in its own 8-byte word next to the tag of this solve - single-copy atomic, the NCCL LL idea), and a consumer
// polls the data itself: one L2 round trip per link instead of flag poll + barrier + data load, and no fence / flag store on
// the producer side (r02: 0.285 -> see DESIGN 3.1). Tags are the solve's epoch, so the buffer is never cleared.
__device__ __forceinline__ ulonglong2 ld_tagged(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_tagged(ulonglong2* p, double value, unsigned tag) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(value), tg = (unsigned long long)tag << 32;
  const unsigned long long lo = (bits & 0xffffffffull) | tg, hi = (bits >> 32) | tg;
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__global__ void __launch_bounds__(256) chol_backsolve_kernel(const double* __restrict__ A, int ld, int nblk,
                                                             const double* __restrict__ dinv, double* __restrict__ y,
                                                             ulonglong2* __restrict__ xt, unsigned tag, int* __restrict__ fail) {
  extern __shared__ double sdi[];  // inv(L_bb), 64 x 65
  __shared__ double xk[NB];
  __shared__ double part[4][NB];
  const int t = threadIdx.x, col = t & 63, rg = t >> 6;  // 4 row groups of 16 rows
  for (int b = nblk - 1 - blockIdx.x; b >= 0; b -= gridDim.x) {
    for (int e = t; e < NB * NB; e += 256) sdi[(e / NB) * (NB + 1) + (e % NB)] = dinv[(size_t)b * NB * NB + e];
    const double yb = t < NB ? y[b * NB + t] : 0.0;
    double acc = 0.0;
    for (int k = nblk - 1; k > b; --k) {
      // the L block does not depend on x: its loads are in flight while x_k is polled
      const double* Lkb = A + (size_t)(k * NB + rg * 16) * ld + b * NB + col;
      double l[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) l[r] = Lkb[(size_t)r * ld];
      const ulonglong2* xs = xt + k * NB + rg * 16;
      ulonglong2 w[16];
      int spins = 0;
      for (;;) {
#pragma unroll
        for (int r = 0; r < 16; ++r) w[r] = ld_tagged(xs + r);
        bool ok = true;
#pragma unroll
        for (int r = 0; r < 16; ++r) ok = ok && (unsigned)(w[r].x >> 32) == tag && (unsigned)(w[r].y >> 32) == tag;
        if (ok) break;
        if (++spins > (1 << 22)) { atomicExch(fail, 2); break; }  // a lost producer: report it like a failed factorisation
      }
#pragma unroll
      for (int r = 0; r < 16; ++r)
        acc += l[r] * __longlong_as_double((long long)((w[r].x & 0xffffffffull) | (w[r].y << 32)));
    }
    part[rg][col] = acc;
    __syncthreads();
    if (t < NB) xk[t] = yb - (part[0][t] + part[1][t] + part[2][t] + part[3][t]);
    __syncthreads();
    // (inv^T y)_c = sum_r inv[r][c] y_r, rows split over the 4 groups
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < 16; ++r) v += sdi[(rg * 16 + r) * (NB + 1) + col] * xk[rg * 16 + r];
    part[rg][col] = v;
    __syncthreads();
    if (t < NB) {
      const double xb = part[0][t] + part[1][t] + part[2][t] + part[3][t];
      st_tagged(xt + b * NB + t, xb, tag);
      y[b * NB + t] = xb;
    }
    __syncthreads();  // part / sdi are rewritten by the next block
  }
}

// a spin that ran into its limit means the factor is garbage: report it like a failed factorisation
__global__ void chol_ll_check_kernel(const int* __restrict__ sync, int* __restrict__ fail) {
  if (sync[LL_TIMEOUT]) atomicExch(fail, 2);
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
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&xtag), sizeof(ulonglong2) * n_pad, st));
  THB_CUDA_CHECK(cudaMemsetAsync(xtag, 0, sizeof(ulonglong2) * n_pad, st));
  epoch = 0;
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&rdiag), sizeof(double) * n_pad, st));
  THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&dscr), sizeof(double) * 2 * NB * NB, st));
  {  // tile tasks of the left-looking kernel: block column J has ceil((row blocks from the diagonal down) / 2) tiles
    const int nrb = (rows_total + NB - 1) / NB;
    h_cols.assign(nblk + 1, 0);
    for (int J = 0; J < nblk; ++J) h_cols[J + 1] = h_cols[J] + (nrb - J + 1) / 2;
    ll_sync_ints = LL_RB + nrb + 8;
    THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&ll_sync), sizeof(int) * ll_sync_ints, st));
    THB_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&ll_cols), sizeof(int) * (nblk + 1), st));
    THB_CUDA_CHECK(cudaMemcpyAsync(ll_cols, h_cols.data(), sizeof(int) * (nblk + 1), cudaMemcpyHostToDevice, st));
    const char* mode = getenv("THB_K4_MODE");
    legacy = mode && !strcmp(mode, "legacy");
  }
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
    set((const void*)chol_panel_kernel, kPanelSmem);
    set((const void*)chol_panel3_kernel, kPanel3Smem);
    set((const void*)chol_dp_kernel, kDpSmem);
    set((const void*)chol_ll_kernel, kLlSmem);
    set((const void*)chol_update_kernel<OB, OB>, (int)(STAGES * (OB + OB) * LDK * sizeof(double)));
    set((const void*)chol_update2_kernel<true, true, 128>, kUpd2Smem);
    set((const void*)chol_update64_kernel<64>, kUpd64Smem);
    set((const void*)chol_update64_kernel<32>, kUpd64Smem);
    set((const void*)chol_update_kernel<OB, NB>, (int)(STAGES * (NB + OB) * LDK * sizeof(double)));
    set((const void*)chol_update_kernel<NB, NB>, (int)(STAGES * (NB + NB) * LDK * sizeof(double)));
  });
  THB_CUDA_CHECK(attr_err);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  num_sms = sms > 0 ? sms : 148;
  {  // the persistent grid must be co-resident: its CTAs spin on each other's progress counters
    int per_sm = 0;
    THB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_ll_kernel, 128, kLlSmem));
    if (per_sm < 1) THB_FAIL(THB_E_CUDA, "chol_ll_kernel does not fit an SM");
    ll_grid = std::min(per_sm * num_sms, h_cols[nblk]);
  }
  if (!legacy) return THB_OK;  // the second stream and the event rings belong to the r01 schedule only
  // the latency-bound diag/panel chain is the critical path: its CTAs must win free SM slots against the trailing update
  int prio_lo = 0, prio_hi = 0;
  THB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  THB_CUDA_CHECK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio_hi));
  THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
  for (int i = 0; i < 4; ++i) {
    THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_pp[i], cudaEventDisableTiming));
    THB_CUDA_CHECK(cudaEventCreateWithFlags(&ev_c2[i], cudaEventDisableTiming));
  }
  return THB_OK;
}

void DenseChol::Free(cudaStream_t st) {
  if (A) cudaFreeAsync(A, st);
  if (dinv) cudaFreeAsync(dinv, st);
  if (x) cudaFreeAsync(x, st);
  if (xtag) cudaFreeAsync(xtag, st);
  if (rdiag) cudaFreeAsync(rdiag, st);
  if (dscr) cudaFreeAsync(dscr, st);
  if (ll_sync) cudaFreeAsync(ll_sync, st);
  if (ll_cols) cudaFreeAsync(ll_cols, st);
  A = dinv = x = rdiag = dscr = nullptr; xtag = nullptr; ll_sync = ll_cols = nullptr;
  if (s2) { cudaStreamDestroy(s2); s2 = nullptr; }
  if (ev_start) { cudaEventDestroy(ev_start); ev_start = nullptr; }
  for (int i = 0; i < 4; ++i) {
    if (ev_pp[i]) { cudaEventDestroy(ev_pp[i]); ev_pp[i] = nullptr; }
    if (ev_c2[i]) { cudaEventDestroy(ev_c2[i]); ev_c2[i] = nullptr; }
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
  // diag + panel fused: 64-row tiles below the diagonal block; the last tile holds only the rhs row (guarded)
  if (ob == 0) {  // the first diagonal block has no producer kernel that could leave a copy: unfused pair
    chol_diag3_kernel<<<1, 64, 0, q>>>(A, ld, k0, rdiag + k0, fail_flag);
    chol_panel3_kernel<<<(n_pad - k0 - NB) / PR3 + 1, 64, kPanel3Smem, q>>>(A, ld, k0, rdiag + k0, rows_total);
    *launches += 1;
  } else {
    chol_dp_kernel<<<(n_pad - k0 - NB) / PR3 + 1, 128, kDpSmem, q>>>(A, ld, k0, rdiag + k0, fail_flag, rows_total, dscr);
  }
  {  // strip: columns k0+64 .. k0+127, rows k0+64 .. end, K = 64
    const int rows = rows_total - (k0 + NB);
    const int tr = (rows + NB - 1) / NB;
    if (tr * 2 <= num_sms) chol_update64_kernel<32><<<(rows + 31) / 32, 256, kUpd64Smem, q>>>(A, ld, k0, NB, k0 + NB, k0 + NB, rows_total, dscr + NB * NB);
    else chol_update64_kernel<64><<<tr, 256, kUpd64Smem, q>>>(A, ld, k0, NB, k0 + NB, k0 + NB, rows_total, dscr + NB * NB);
  }
  chol_dp_kernel<<<(n_pad - k0 - OB) / PR3 + 1, 128, kDpSmem, q>>>(A, ld, k0 + NB, rdiag + k0 + NB, fail_flag, rows_total, dscr + NB * NB);
  *launches += 3;
}

// Factor the lower triangle of A (n_pad x n_pad, row n_pad = rhs) and leave the solution in x.
//
// Schedule (outer blocks of 128 columns, b = 0 .. nob-1), two streams:
//   critical stream s2 (high priority):  L(b)  block column b  -=  panels b-2, b-1   (left-looking, K = 256)
//                                        PP(b) diag / panel / strip / diag / panel of block b
//   bulk stream st:                      C2(b) columns >= b+3  -=  panel b            (right-looking, K = 128)
// L(b) needs C2(b-3) and every earlier C2 (stream order), C2(b) needs PP(b). The serial chain L, PP, L, PP ... therefore
// runs two blocks ahead of the bulk updates instead of waiting for each of them (the r01 schedule had the next
// block-column update queued behind the whole trailing update). Event rings of four: a dependency spans <= 3 blocks.
int DenseChol::FactorAndSolve(cudaStream_t st, int* fail_flag, int* launches) {
  if (legacy) return FactorLegacy(st, fail_flag, launches);
  // one persistent kernel factors (left-looking tile tasks, see chol_ll_kernel), forward-substitutes the rhs row and leaves
  // the inverted diagonal factors in dinv; then the one-launch backward substitution
  THB_CUDA_CHECK(cudaMemsetAsync(ll_sync, 0, sizeof(int) * ll_sync_ints, st));
  long long* prof = nullptr;
  const bool want_prof = getenv("THB_K4_PROF") != nullptr;
  if (want_prof) { cudaMalloc(&prof, sizeof(long long) * nblk * 16); cudaMemset(prof, 0, sizeof(long long) * nblk * 16); }
  chol_ll_kernel<<<ll_grid, 128, kLlSmem, st>>>(A, ld, nblk, rows_total, ll_cols, ll_sync, dinv, fail_flag, prof);
  if (want_prof) {  // debugging aid: per-column phase stamps of the chain (ns, relative to the first stamp)
    std::vector<long long> h(nblk * 16);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), prof, sizeof(long long) * nblk * 16, cudaMemcpyDeviceToHost);
    cudaFree(prof);
    const long long t0 = h[0];
    for (int J = 0; J < nblk; ++J) {
      fprintf(stderr, "K4PROF J=%d", J);
      for (int m = 0; m < 2; ++m) { fprintf(stderr, " |"); for (int k = 0; k < 7; ++k) fprintf(stderr, " %lld", h[(J * 2 + m) * 8 + k] ? h[(J * 2 + m) * 8 + k] - t0 : -1); }
      fprintf(stderr, "\n");
    }
  }
  chol_ll_check_kernel<<<1, 1, 0, st>>>(ll_sync, fail_flag);
  chol_copy_row_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(A + (size_t)n_pad * ld, x, n_pad);
  if (++epoch == 0) epoch = 1;  // tag 0 is the cleared buffer
  chol_backsolve_kernel<<<std::min(nblk, num_sms), 256, kBackSmem, st>>>(A, ld, nblk, dinv, x, xtag, epoch, fail_flag);
  *launches += 4;
  THB_CUDA_CHECK(cudaGetLastError());
  return THB_OK;
}

int DenseChol::FactorLegacy(cudaStream_t st, int* fail_flag, int* launches) {
  const int nob = n_pad / OB;
  THB_CUDA_CHECK(cudaEventRecord(ev_start, st));
  THB_CUDA_CHECK(cudaStreamWaitEvent(s2, ev_start, 0));
  for (int b = 0; b < nob; ++b) {
    const int k0 = b * OB;
    if (b >= 1) {
      if (b >= 3) THB_CUDA_CHECK(cudaStreamWaitEvent(s2, ev_c2[(b - 3) & 3], 0));
      const int pb = b >= 2 ? b - 2 : 0;  // first pending panel
      const int rows = rows_total - k0;
      const dim3 grid((rows + NB - 1) / NB, OB / NB);
      if ((int)(grid.x * grid.y) * 2 <= num_sms)  // second half of the factorisation: halve the tiles, use more SMs
        chol_update64_kernel<32><<<dim3((rows + 31) / 32, OB / NB), 256, kUpd64Smem, s2>>>(A, ld, pb * OB, (b - pb) * OB, k0, k0, rows_total, dscr);
      else
        chol_update64_kernel<64><<<grid, 256, kUpd64Smem, s2>>>(A, ld, pb * OB, (b - pb) * OB, k0, k0, rows_total, dscr);
      *launches += 1;
    }
    PanelPair(s2, b, fail_flag, launches);
    THB_CUDA_CHECK(cudaEventRecord(ev_pp[b & 3], s2));
    const int ntr = nob - b - 3;  // 128-row tile rows of the region right of block column b+2
    if (ntr >= 1) {
      THB_CUDA_CHECK(cudaStreamWaitEvent(st, ev_pp[b & 3], 0));
      const int tiles = ntr * (ntr + 1) + 2 * ntr;  // 128 x 64 tiles of the lower triangle + the rhs tile-row
      chol_update2_kernel<true, true, 128><<<tiles, 128, kUpd2Smem, st>>>(A, ld, k0, OB, k0 + 3 * OB, rows_total);
      THB_CUDA_CHECK(cudaEventRecord(ev_c2[b & 3], st));
      *launches += 1;
    }
  }
  THB_CUDA_CHECK(cudaStreamWaitEvent(st, ev_pp[(nob - 1) & 3], 0));
  chol_inverse_kernel<<<nblk, 256, kInvSmem, st>>>(A, ld, dinv);
  chol_copy_row_kernel<<<(n_pad + 255) / 256, 256, 0, st>>>(A + (size_t)n_pad * ld, x, n_pad);
  if (++epoch == 0) epoch = 1;  // tag 0 is the cleared buffer
  chol_backsolve_kernel<<<std::min(nblk, num_sms), 256, kBackSmem, st>>>(A, ld, nblk, dinv, x, xtag, epoch, fail_flag);
  *launches += 3;
  THB_CUDA_CHECK(cudaGetLastError());
  return THB_OK;
}

}  // namespace thb
