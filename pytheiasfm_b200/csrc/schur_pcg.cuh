// Preconditioned conjugate gradients on the reduced camera system: ceres ITERATIVE_SCHUR with the SCHUR_JACOBI preconditioner
// (reference option BundleAdjustmentOptions::linear_solver_type, sfm/bundle_adjustment/bundle_adjustment.h:96-99; solver set up in
// bundle_adjuster.cc:63-88). Ceres 2.2 semantics restated (conjugate_gradients_solver.h, iterative_schur_complement_solver.cc):
// x0 = 0, z = M^-1 r with M = block diagonal of S (one block per camera / shared intrinsics block), termination by the Q-test
// i (Q_i - Q_{i-1}) / Q_i < eta (the residual test is off: the Levenberg-Marquardt strategy passes r_tolerance = -1), the residual
// recomputed as b - S x every 10th iteration, max_linear_solver_iterations. Ceres applies S implicitly through the Jacobian;
// here S is already in HBM (K2/K3 build it for the exact solver), so one CG iteration is one pass over its lower triangle.
//
// ONE persistent cooperative kernel runs the whole solve: every CTA multiplies its share of 64 x 128 tiles of the lower
// triangle (a lane owns 4 columns: 256-bit loads, column sums in registers, row sums by warp shuffle; both triangles of the
// product come from one read of S: 8 n^2 / 2 bytes per iteration), then a grid barrier; CTA 0 does the O(n) vector work and
// the termination test and publishes the next command. HBM-bound: n = 6016 -> 145 MB per iteration.
#ifndef THB_SCHUR_PCG_CUH_
#define THB_SCHUR_PCG_CUH_

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace thb {

enum { PCG_ARRIVE = 0, PCG_GEN = 1, PCG_CMD = 2, PCG_TIMEOUT = 3, PCG_ITERS = 4, PCG_TERM = 5, PCG_SYNC_INTS = 8 };
enum { PCG_CMD_EXIT = 0, PCG_CMD_SP = 1, PCG_CMD_SX = 2 };
enum { PCG_TERM_SUCCESS = 0, PCG_TERM_NO_CONVERGENCE = 1, PCG_TERM_FAILURE = 2 };
enum { PCG_RHO = 0, PCG_Q0 = 1, PCG_ALPHA = 2, PCG_ZETA = 3, PCG_VALS = 8 };
constexpr int kPcgThreads = 256, kPcgTileRows = 64, kPcgTileCols = 128, kPcgMaxBlockDim = 9, kPcgResetPeriod = 10;

__device__ __forceinline__ int pcg_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void pcg_st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void pcg_ld256(const double* p, double* o) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(o[0]), "=d"(o[1]), "=d"(o[2]), "=d"(o[3]) : "l"(p));
}

// inverse of every diagonal block of S (lower triangle stored) through its Cholesky factor; thread per block
__global__ void k_pcg_precond(int nblk, const int2* __restrict__ blk, const double* __restrict__ A, int ld, double* __restrict__ minv, int* fail) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblk) return;
  const int o = blk[b].x, d = blk[b].y;
  double L[kPcgMaxBlockDim * kPcgMaxBlockDim];
  bool ok = true;
  for (int i = 0; i < d; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[(size_t)(o + i) * ld + o + j];
      for (int k = 0; k < j; ++k) v -= L[i * d + k] * L[j * d + k];
      if (i == j) { if (!(v > 0.0) || !isfinite(v)) { ok = false; v = 1.0; } L[i * d + i] = sqrt(v); }
      else L[i * d + j] = v / L[j * d + j];
    }
  double* out = minv + (size_t)b * kPcgMaxBlockDim * kPcgMaxBlockDim;
  for (int c = 0; c < d; ++c) {
    double y[kPcgMaxBlockDim], x[kPcgMaxBlockDim];
    for (int i = 0; i < d; ++i) { double v = i == c ? 1.0 : 0.0; for (int k = 0; k < i; ++k) v -= L[i * d + k] * y[k]; y[i] = v / L[i * d + i]; }
    for (int i = d - 1; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < d; ++k) v -= L[k * d + i] * x[k]; x[i] = v / L[i * d + i]; }
    for (int i = 0; i < d; ++i) { out[i * d + c] = x[i]; if (!isfinite(x[i])) ok = false; }
  }
  if (!ok) atomicExch(fail, 1);
}

struct PcgArgs {
  const double* A; int ld; int n_pad;
  const double* b;       // rhs (row n_pad of A)
  double *x, *r, *p, *z, *q, *tmp;
  const double* minv; const int2* blk; int nblk;
  const int2* tasks; int ntasks;
  double eta; int max_it;
  int* sync; double* vals; int* fail; int* iters_total;
  long long* prof;  // THB_PCG_PROF: 3 accumulators (ns) or nullptr
};

// z = M^-1 r, returns this thread's share of r.z
__device__ __forceinline__ double pcg_apply_precond(const PcgArgs& a) {
  // restrict-qualified views: without them every store to z orders the loads of the next block behind it (one L2 round trip each)
  const double* __restrict__ R = a.r;
  double* __restrict__ Z = a.z;
  const double* __restrict__ Minv = a.minv;
  const int2* __restrict__ blk = a.blk;
  double part = 0.0;
#pragma unroll 2
  for (int b = threadIdx.x; b < a.nblk; b += blockDim.x) {
    const int o = blk[b].x, d = blk[b].y;
    const double* __restrict__ M = Minv + (size_t)b * kPcgMaxBlockDim * kPcgMaxBlockDim;
    double rr[kPcgMaxBlockDim], m[kPcgMaxBlockDim * kPcgMaxBlockDim];
#pragma unroll
    for (int j = 0; j < kPcgMaxBlockDim; ++j) rr[j] = j < d ? R[o + j] : 0.0;
#pragma unroll
    for (int e = 0; e < kPcgMaxBlockDim * kPcgMaxBlockDim; ++e) m[e] = e < d * d ? M[e] : 0.0;  // all loads of the block in flight at once
    if (d == 6) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) v += m[i * 6 + j] * rr[j];
        Z[o + i] = v;
        part += rr[i] * v;
      }
    } else {
      for (int i = 0; i < d; ++i) {
        double v = 0.0;
        for (int j = 0; j < d; ++j) v += M[i * d + j] * rr[j];
        Z[o + i] = v;
        part += rr[i] * v;
      }
    }
  }
  return part;
}

// x = 0, r = b, z = M^-1 r, p = z, rho = r.z; command for the persistent kernel
__global__ void k_pcg_init(PcgArgs a) {
  __shared__ double red[32];
  __shared__ double s_val;
  const int t = threadIdx.x;
  double nb = 0.0;
  for (int i = t; i < a.n_pad; i += blockDim.x) { const double v = a.b[i]; a.r[i] = v; a.x[i] = 0.0; a.z[i] = 0.0; a.q[i] = 0.0; a.tmp[i] = 0.0; nb += v * v; }
  nb = block_sum(nb, red);
  if (t == 0) s_val = nb;
  __syncthreads();
  const bool zero_rhs = s_val == 0.0;
  __syncthreads();
  double rho = zero_rhs ? 0.0 : pcg_apply_precond(a);
  rho = block_sum(rho, red);
  __syncthreads();
  for (int i = t; i < a.n_pad; i += blockDim.x) a.p[i] = a.z[i];
  if (t == 0) {
    a.vals[PCG_RHO] = rho; a.vals[PCG_Q0] = 0.0; a.vals[PCG_ZETA] = 0.0;
    int cmd = PCG_CMD_SP, term = PCG_TERM_NO_CONVERGENCE;
    if (zero_rhs) { cmd = PCG_CMD_EXIT; term = PCG_TERM_SUCCESS; }                                  // "Convergence. |b| = 0."
    else if (rho == 0.0 || isinf(rho) || isnan(rho)) { cmd = PCG_CMD_EXIT; term = PCG_TERM_FAILURE; atomicExch(a.fail, 1); }
    a.sync[PCG_ARRIVE] = 0; a.sync[PCG_GEN] = 0; a.sync[PCG_CMD] = cmd; a.sync[PCG_TIMEOUT] = 0; a.sync[PCG_ITERS] = 0; a.sync[PCG_TERM] = term;
  }
}

// out += S v over this CTA's tiles (out is zero when the pass starts)
__device__ __forceinline__ void pcg_symv_tiles(const PcgArgs& a, const double* __restrict__ v, double* __restrict__ out, double (*colsum)[kPcgTileCols]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  constexpr int RW = kPcgTileRows / (kPcgThreads / 32);  // rows per warp = 8
  for (int task = blockIdx.x; task < a.ntasks; task += gridDim.x) {
    const int r0 = a.tasks[task].x * kPcgTileRows + w * RW, c0 = a.tasks[task].y * kPcgTileCols + 4 * lane;
    double pc[4], acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) pc[k] = __ldcg(v + c0 + k);
    const double prow = lane < RW ? __ldcg(v + r0 + lane) : 0.0;
    double e[RW][4];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      if (c0 <= r0 + i) pcg_ld256(a.A + (size_t)(r0 + i) * a.ld + c0, e[i]);
      else { e[i][0] = e[i][1] = e[i][2] = e[i][3] = 0.0; }
    }
#pragma unroll
    for (int i = 0; i < RW; ++i) {
      const int row = r0 + i;
      const double pr = __shfl_sync(0xffffffffu, prow, i);
      double rowpart = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int col = c0 + k;
        const double s = col <= row ? e[i][k] : 0.0;  // the upper triangle is not part of the stored matrix
        rowpart = fma(s, pc[k], rowpart);
        if (col < row) acc[k] = fma(s, pr, acc[k]);
      }
      rowpart = warp_sum(rowpart);
      if (lane == 0) atomicAdd(out + row, rowpart);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) colsum[w][4 * lane + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < kPcgTileCols) {
      double s = 0.0;
#pragma unroll
      for (int u = 0; u < kPcgThreads / 32; ++u) s += colsum[u][threadIdx.x];
      if (s != 0.0) atomicAdd(out + a.tasks[task].y * kPcgTileCols + threadIdx.x, s);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPcgThreads, 2) k_pcg(PcgArgs a) {
  __shared__ double colsum[kPcgThreads / 32][kPcgTileCols];
  __shared__ double red[32];
  __shared__ double s_val[2];
  __shared__ int s_cmd;
  constexpr int kSpinLimit = 1 << 22;
  const int t = threadIdx.x;
  int gen = 0, it = 1;
  int cmd = a.sync[PCG_CMD];  // written by k_pcg_init
  double rho_cur = a.vals[PCG_RHO], q0_cur = a.vals[PCG_Q0];  // CTA 0 keeps the CG scalars in registers (every thread the same value)
  auto now = []() { long long g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return g; };
  long long t_mark = (a.prof && blockIdx.x == 0 && t == 0) ? now() : 0;
  auto lap = [&](int slot) {  // THB_PCG_PROF: nanoseconds CTA 0 spends in its own tiles / waiting for the grid / in the vector work
    if (a.prof && blockIdx.x == 0 && t == 0) { const long long n = now(); a.prof[slot] += n - t_mark; t_mark = n; }
  };
  while (cmd != PCG_CMD_EXIT) {
    pcg_symv_tiles(a, cmd == PCG_CMD_SP ? a.p : a.x, cmd == PCG_CMD_SP ? a.q : a.tmp, colsum);
    __threadfence();
    __syncthreads();
    lap(0);
    if (blockIdx.x != 0) {
      if (t == 0) {
        atomicAdd(&a.sync[PCG_ARRIVE], 1);
        int spins = 0, c = PCG_CMD_EXIT;
        bool released = true;
        while (pcg_ld_acquire(&a.sync[PCG_GEN]) == gen) {
          __nanosleep(200);
          if (++spins > kSpinLimit) { atomicExch(&a.sync[PCG_TIMEOUT], 1); released = false; break; }  // never hang the GPU
        }
        if (released) c = a.sync[PCG_CMD];
        s_cmd = c;
      }
      __syncthreads();
      cmd = s_cmd;
      ++gen;
      __syncthreads();
      continue;
    }
    // ---- CTA 0: wait for the other CTAs, then the vector work of this iteration ----
    if (t == 0) {
      int spins = 0, ok = 1;
      while (pcg_ld_acquire(&a.sync[PCG_ARRIVE]) != (int)gridDim.x - 1) {
        __nanosleep(100);
        if (++spins > kSpinLimit) { atomicExch(&a.sync[PCG_TIMEOUT], 1); ok = 0; break; }
      }
      a.sync[PCG_ARRIVE] = 0;
      s_cmd = ok;
    }
    __syncthreads();
    lap(1);
    int next = PCG_CMD_SP, term = -1;
    if (!s_cmd) { next = PCG_CMD_EXIT; term = PCG_TERM_FAILURE; }
    __syncthreads();
    bool have_r = false;
    // CTA 0 alone touches x, r, p, z between the barriers (the other CTAs read p / x through L2 after the release); q and tmp
    // are summed by every SM's atomics, hence __ldcg. restrict + unrolling keeps ~8 independent L2 round trips in flight per
    // thread: with plain struct members every store ordered the next load behind it (39 us of vector work per CG iteration).
    double* __restrict__ X = a.x; double* __restrict__ R = a.r; double* __restrict__ Pv = a.p;
    const double* __restrict__ Zv = a.z; double* __restrict__ Q = a.q; double* __restrict__ TMP = a.tmp;
    const double* __restrict__ Bv = a.b;
    const int n_pad = a.n_pad;
    if (next != PCG_CMD_EXIT && cmd == PCG_CMD_SP) {
      double pq = 0.0;
#pragma unroll 8
      for (int i = t; i < n_pad; i += kPcgThreads) pq += Pv[i] * __ldcg(Q + i);
      pq = block_sum(pq, red);
      if (t == 0) s_val[0] = pq;
      __syncthreads();
      pq = s_val[0];
      const double rho = rho_cur;
      const double alpha = rho / pq;
      if (pq <= 0.0 || isinf(pq) || isnan(pq)) { next = PCG_CMD_EXIT; term = PCG_TERM_NO_CONVERGENCE; }  // indefinite direction: keep x
      else if (isinf(alpha)) { next = PCG_CMD_EXIT; term = PCG_TERM_FAILURE; }
      else {
        if (it % kPcgResetPeriod == 0) {
#pragma unroll 8
          for (int i = t; i < n_pad; i += kPcgThreads) { X[i] += alpha * Pv[i]; Q[i] = 0.0; }
          next = PCG_CMD_SX;  // r = b - S x needs another pass over S
        } else {
#pragma unroll 8
          for (int i = t; i < n_pad; i += kPcgThreads) { const double qi = __ldcg(Q + i); X[i] += alpha * Pv[i]; R[i] -= alpha * qi; Q[i] = 0.0; }
          have_r = true;
        }
      }
      __syncthreads();
    } else if (next != PCG_CMD_EXIT) {  // cmd == PCG_CMD_SX
#pragma unroll 8
      for (int i = t; i < n_pad; i += kPcgThreads) { R[i] = Bv[i] - __ldcg(TMP + i); TMP[i] = 0.0; }
      have_r = true;
      __syncthreads();
    }
    if (have_r) {
      double q1 = 0.0;
#pragma unroll 8
      for (int i = t; i < n_pad; i += kPcgThreads) q1 -= X[i] * (Bv[i] + R[i]);
      q1 = block_sum(q1, red);
      if (t == 0) s_val[0] = q1;
      __syncthreads();
      q1 = s_val[0];
      const double zeta = it * (q1 - q0_cur) / q1;
      q0_cur = q1;
      if (t == 0) { a.vals[PCG_ZETA] = zeta; a.sync[PCG_ITERS] = it; }
      if (zeta < a.eta) { next = PCG_CMD_EXIT; term = PCG_TERM_SUCCESS; }
      else if (it >= a.max_it) { next = PCG_CMD_EXIT; term = PCG_TERM_NO_CONVERGENCE; }
      else {
        ++it;
        double rho = pcg_apply_precond(a);
        rho = block_sum(rho, red);
        if (t == 0) s_val[1] = rho;
        __syncthreads();
        rho = s_val[1];
        const double beta = rho / rho_cur;
        if (rho == 0.0 || isinf(rho) || isnan(rho) || beta == 0.0 || isinf(beta)) { next = PCG_CMD_EXIT; term = PCG_TERM_FAILURE; }
        else {
          rho_cur = rho;
#pragma unroll 8
          for (int i = t; i < n_pad; i += kPcgThreads) Pv[i] = Zv[i] + beta * Pv[i];
          next = PCG_CMD_SP;
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
      if (term >= 0) { a.sync[PCG_TERM] = term; if (term == PCG_TERM_FAILURE) atomicExch(a.fail, 1); }
      if (next == PCG_CMD_EXIT) atomicAdd(a.iters_total, a.sync[PCG_ITERS]);
      a.sync[PCG_CMD] = next;
      __threadfence();
      pcg_st_release(&a.sync[PCG_GEN], gen + 1);
    }
    lap(2);
    ++gen;
    cmd = next;
    __syncthreads();
  }
}

struct SchurPcg {
  int n_pad = 0, nblk = 0, ntasks = 0, grid = 0;
  double *r = nullptr, *p = nullptr, *z = nullptr, *q = nullptr, *tmp = nullptr, *minv = nullptr, *vals = nullptr;
  int2 *blk = nullptr, *tasks = nullptr;
  int* sync = nullptr;

  // blocks: (offset, dim <= 9) of every diagonal block of the preconditioner
  int Init(int n_pad_, const std::vector<int2>& blocks, cudaStream_t st) {
    n_pad = n_pad_; nblk = (int)blocks.size();
    std::vector<int2> t;
    for (int rt = 0; rt < n_pad / kPcgTileRows; ++rt)
      for (int ct = 0; ct * kPcgTileCols <= rt * kPcgTileRows + kPcgTileRows - 1; ++ct) t.push_back(make_int2(rt, ct));
    std::reverse(t.begin(), t.end());  // the long rows first
    ntasks = (int)t.size();
    THB_CUDA_CHECK(cudaMallocAsync(&r, sizeof(double) * 5 * n_pad, st));
    p = r + n_pad; z = p + n_pad; q = z + n_pad; tmp = q + n_pad;
    THB_CUDA_CHECK(cudaMallocAsync(&minv, sizeof(double) * std::max(1, nblk) * kPcgMaxBlockDim * kPcgMaxBlockDim, st));
    THB_CUDA_CHECK(cudaMallocAsync(&vals, sizeof(double) * PCG_VALS, st));
    THB_CUDA_CHECK(cudaMallocAsync(&blk, sizeof(int2) * std::max(1, nblk), st));
    THB_CUDA_CHECK(cudaMallocAsync(&tasks, sizeof(int2) * std::max(1, ntasks), st));
    THB_CUDA_CHECK(cudaMallocAsync(&sync, sizeof(int) * PCG_SYNC_INTS, st));
    if (nblk) THB_CUDA_CHECK(cudaMemcpyAsync(blk, blocks.data(), sizeof(int2) * nblk, cudaMemcpyHostToDevice, st));
    if (ntasks) THB_CUDA_CHECK(cudaMemcpyAsync(tasks, t.data(), sizeof(int2) * ntasks, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaStreamSynchronize(st));  // the host vectors go out of scope
    int dev = 0, sms = 0, per_sm = 0;
    THB_CUDA_CHECK(cudaGetDevice(&dev));
    THB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    THB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg, kPcgThreads, 0));
    grid = std::max(1, std::min(ntasks, sms * std::max(1, std::min(per_sm, 2))));
    return THB_OK;
  }
  void Free(cudaStream_t st) {
    if (r) cudaFreeAsync(r, st);
    if (minv) cudaFreeAsync(minv, st);
    if (vals) cudaFreeAsync(vals, st);
    if (blk) cudaFreeAsync(blk, st);
    if (tasks) cudaFreeAsync(tasks, st);
    if (sync) cudaFreeAsync(sync, st);
    r = minv = vals = nullptr; blk = tasks = nullptr; sync = nullptr;
  }
  // S y = rhs, y into x. `fail` is raised on LINEAR_SOLVER_FAILURE (an invalid step for the trust-region loop), the CG
  // iterations of this solve are added to *iters_total, sync[PCG_TIMEOUT] reports a broken grid barrier.
  int Solve(cudaStream_t st, const double* A, int ld, const double* rhs, double* x, double eta, int max_it, int* fail, int* iters_total, int* launches) {
    PcgArgs a;
    a.A = A; a.ld = ld; a.n_pad = n_pad; a.b = rhs; a.x = x; a.r = r; a.p = p; a.z = z; a.q = q; a.tmp = tmp;
    a.minv = minv; a.blk = blk; a.nblk = nblk; a.tasks = tasks; a.ntasks = ntasks; a.eta = eta; a.max_it = std::max(1, max_it);
    a.sync = sync; a.vals = vals; a.fail = fail; a.iters_total = iters_total;
    a.prof = nullptr;
    long long* d_prof = nullptr;
    if (getenv("THB_PCG_PROF")) { cudaMalloc(&d_prof, sizeof(long long) * 4); cudaMemsetAsync(d_prof, 0, sizeof(long long) * 4, st); a.prof = d_prof; }
    if (nblk) k_pcg_precond<<<(nblk + 127) / 128, 128, 0, st>>>(nblk, blk, A, ld, minv, fail);
    k_pcg_init<<<1, kPcgThreads, 0, st>>>(a);
    void* args[] = {&a};
    THB_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)k_pcg, dim3(grid), dim3(kPcgThreads), args, 0, st));
    *launches += 3;
    if (d_prof) {
      long long h[4] = {0, 0, 0, 0};
      cudaStreamSynchronize(st);
      cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
      cudaFree(d_prof);
      fprintf(stderr, "PCGPROF CTA 0: own tiles %.1f us, waiting for the grid %.1f us, vector work + publish %.1f us\n", h[0] * 1e-3, h[1] * 1e-3, h[2] * 1e-3);
    }
    return THB_OK;
  }
};

}  // namespace thb
#endif  // THB_SCHUR_PCG_CUH_
