// Micro-benchmarks behind the K4 design decisions (FP64 SIMT / DMMA latency and issue rate, per-kernel timings of the
// Cholesky building blocks at warm clocks). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -I pytheiasfm_b200/csrc tools/microbench/k4_micro.cu build/ba_solver.o ... (see tools/microbench/build_micro.sh). GPU box only.
#define THB_K4_PROBE
#include "../pytheiasfm_b200/csrc/dense_chol.cu"

#include <cstdio>
#include <vector>

using namespace thb;

__global__ void k_dfma_chain(double* out, int iters, int nchains) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double m = 0.999999, c = 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < nchains) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0); out[blockIdx.x * 2 + 1] = s; }
}
template <int NCH>
__global__ void k_dfma_chain_t(double* out, int iters) {
  double a[NCH];
  for (int i = 0; i < NCH; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  const double m = 0.999999, c = 1e-7;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NCH; ++i) s += a[i];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0); out[blockIdx.x * 2 + 1] = s; }
}
template <int NCH>
__global__ void k_dmma_chain_t(double* out, int iters) {
  double d[NCH][2];
  for (int i = 0; i < NCH; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; }
  const double a = 1e-3 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) dmma_m8n8k4(d[i][0], d[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < NCH; ++i) s += d[i][0] + d[i][1];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0); out[blockIdx.x * 2 + 1] = s; }
}
// the update kernel's register pattern: 8 A fragments x 4 B fragments -> 32 accumulators, no memory traffic
__global__ void __launch_bounds__(128) k_dmma_tile(double* out, int iters) {
  double acc[8][4][2], a[8], b[4];
  for (int u = 0; u < 8; ++u) { a[u] = 1e-3 * (threadIdx.x + u); for (int v = 0; v < 4; ++v) { acc[u][v][0] = u; acc[u][v][1] = v; } }
  for (int v = 0; v < 4; ++v) b[v] = 1e-3 * v;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) dmma_m8n8k4(acc[u][v][0], acc[u][v][1], a[u], b[v]);
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] += 1e-9;
  }
  long long t1 = clock64();
  double s = 0;
  for (int u = 0; u < 8; ++u) for (int v = 0; v < 4; ++v) s += acc[u][v][0] + acc[u][v][1];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0); out[blockIdx.x * 2 + 1] = s; }
}
__global__ void k_shfl_chain(double* out, int iters) {
  double a = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = (double)(t1 - t0); out[1] = a; }
}

__global__ void __launch_bounds__(256) k_dmma_peak(double* out, int iters) {
  double d[16][2];
  for (int i = 0; i < 16; ++i) { d[i][0] = threadIdx.x; d[i][1] = i; }
  const double a = 1e-3 * threadIdx.x, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma_m8n8k4(d[i][0], d[i][1], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += d[i][0] + d[i][1];
  if (s == 1.2345) out[0] = s;
}
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters) {
  double d[16];
  for (int i = 0; i < 16; ++i) d[i] = threadIdx.x + i;
  const double a = 0.999999, b = 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = fma(d[i], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += d[i];
  if (s == 1.2345) out[0] = s;
}

__global__ void k_clock_probe(double* out) {
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  const long long c0 = clock64();
  while (clock64() - c0 < 200000) {}
  const long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  out[0] = (double)(c1 - c0) / (double)(g1 - g0) * 1e3;  // MHz
}
static double probe_mhz(double* d_out) {
  k_clock_probe<<<1, 1>>>(d_out);
  double h; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  return h;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f();  // warm
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const bool blocks_only = argc > 1;  // one launch of every building block (for ncu --set full)
  double* d_out; cudaMalloc(&d_out, 4096);
  double h[4];
  const int iters = 4096;
  // warm the clocks
  for (int i = 0; i < (blocks_only ? 2 : 200); ++i) k_dfma_chain_t<8><<<148 * 4, 256>>>(d_out, 20000);
  cudaDeviceSynchronize();
#define RUN(label, kern, blocks, threads, nops)                                        \
  kern<<<blocks, threads>>>(d_out, iters); kern<<<blocks, threads>>>(d_out, iters);    \
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);                                    \
  printf("%-44s %8.2f cycles per op-set (%d ops) -> %.2f cycles/op\n", label, h[0] / iters, nops, h[0] / iters / nops);
  RUN("DFMA 1 warp, 1 chain (latency)", k_dfma_chain_t<1>, 1, 32, 1);
  RUN("DFMA 1 warp, 2 chains", k_dfma_chain_t<2>, 1, 32, 2);
  RUN("DFMA 1 warp, 4 chains", k_dfma_chain_t<4>, 1, 32, 4);
  RUN("DFMA 1 warp, 8 chains", k_dfma_chain_t<8>, 1, 32, 8);
  RUN("DFMA 4 warps (1/SMSP), 8 chains", k_dfma_chain_t<8>, 1, 128, 8);
  RUN("DFMA 8 warps (2/SMSP), 8 chains", k_dfma_chain_t<8>, 1, 256, 8);
  RUN("DFMA 16 warps (4/SMSP), 8 chains", k_dfma_chain_t<8>, 1, 512, 8);
  RUN("DMMA 1 warp, 1 chain (latency)", k_dmma_chain_t<1>, 1, 32, 1);
  RUN("DMMA 1 warp, 4 chains", k_dmma_chain_t<4>, 1, 32, 4);
  RUN("DMMA 1 warp, 16 chains", k_dmma_chain_t<16>, 1, 32, 16);
  RUN("DMMA 4 warps, 16 chains", k_dmma_chain_t<16>, 1, 128, 16);
  RUN("DMMA 8 warps, 16 chains", k_dmma_chain_t<16>, 1, 256, 16);
  RUN("DMMA 16 warps, 8 chains", k_dmma_chain_t<8>, 1, 512, 8);
  RUN("SHFL f64 chain 1 warp", k_shfl_chain, 1, 32, 1);
  RUN("DMMA 8x4 register tile, 4 warps (1/SMSP)", k_dmma_tile, 1, 128, 32);
  RUN("DMMA 8x4 register tile, 8 warps (2/SMSP, 2 CTAs)", k_dmma_tile, 2 * 148, 128, 32);

  {
    const int it2 = 20000;
    float ms = time_ms([&] { k_dmma_peak<<<148 * 2, 256>>>(d_out, it2); }, 5);
    printf("DMMA register-resident peak: %.2f TFLOP/s (%.2f ms)\n", 148.0 * 2 * 8 * it2 * 16 * 256 * 2 / ms / 1e9, ms);
    ms = time_ms([&] { k_dfma_peak<<<148 * 2, 256>>>(d_out, it2); }, 5);
    printf("DFMA register-resident peak: %.2f TFLOP/s (%.2f ms)\n", 148.0 * 2 * 256 * it2 * 16 * 2.0 / ms / 1e9, ms);
  }
  // ---- building blocks at warm clocks on an n = 6016 matrix ----
  DenseChol ch;
  ch.Init(6000, 0);
  int* d_fail; cudaMalloc(&d_fail, 4); cudaMemset(d_fail, 0, 4);
  ch.Clear(0);
  // SPD fill: diagonal n, small off-diagonal
  std::vector<double> hA((size_t)ch.rows_total * ch.ld, 0.0);
  for (int i = 0; i < ch.n_pad; ++i) { for (int j = 0; j < i; ++j) hA[(size_t)i * ch.ld + j] = 1e-3 * ((i * 31 + j * 17) % 13 - 6); hA[(size_t)i * ch.ld + i] = 10.0; }
  cudaMemcpy(ch.A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
  const int reps = blocks_only ? 1 : 50;
  double* A = ch.A; int ld = ch.ld, rows_total = ch.rows_total, n_pad = ch.n_pad;
  double* rd = ch.rdiag;
  {  // compact variants against the unrolled ones on the same block column
    double *A2, *rd2; cudaMalloc(&A2, hA.size() * 8); cudaMalloc(&rd2, 8 * n_pad);
    cudaMemcpy(A2, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
    chol_diag_kernel<<<1, 256>>>(A, ld, 0, rd, d_fail);
    chol_panel_kernel<<<(n_pad - NB) / PR + 1, 128, kPanelSmem>>>(A, ld, 0, rd, rows_total);
    chol_diag3_kernel<<<1, 64>>>(A2, ld, 0, rd2, d_fail);
    chol_panel3_kernel<<<(n_pad - NB) / PR3 + 1, 64, kPanel3Smem>>>(A2, ld, 0, rd2, rows_total);
    std::vector<double> h1(hA.size()), h2(hA.size());
    cudaMemcpy(h1.data(), A, hA.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(h2.data(), A2, hA.size() * 8, cudaMemcpyDeviceToHost);
    double md = 0, mp = 0;
    for (int i = 0; i < rows_total; ++i) for (int j = 0; j < 64; ++j) {
      if (i < 64 && j > i) continue;
      const double df = fabs(h1[(size_t)i * ld + j] - h2[(size_t)i * ld + j]);
      if (i < 64) md = fmax(md, df); else mp = fmax(mp, df);
    }
    int hf = 0; cudaMemcpy(&hf, d_fail, 4, cudaMemcpyDeviceToHost);
    printf("diag3 vs diag max diff %.3e, panel3 vs panel max diff %.3e (fail flag %d) [%s]\n", md, mp, hf, cudaGetErrorString(cudaGetLastError()));
    cudaMemcpy(A, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice);
    cudaFree(A2); cudaFree(rd2);
  }
  printf("diag64            %.2f us\n", 1e3 * time_ms([&] { chol_diag_kernel<<<1, 256>>>(A, ld, 4096, rd, d_fail); }, reps));
  printf("SM clock right after the diag loop: %.0f MHz\n", probe_mhz(d_out));
  printf("diag64 row/thread %.2f us\n", 1e3 * time_ms([&] { chol_diag3_kernel<<<1, 64>>>(A, ld, 4096, rd, d_fail); }, reps));
  for (int k0 : {0, 3008, 5760})
    printf("panel3 k0=%4d    %.2f us (%d CTAs)\n", k0, 1e3 * time_ms([&] { chol_panel3_kernel<<<(n_pad - k0 - NB) / PR3 + 1, 64, kPanel3Smem>>>(A, ld, k0, rd, rows_total); }, reps), (n_pad - k0 - NB) / PR3 + 1);
  for (int k0 : {0, 3008, 5760})
    printf("panel k0=%4d     %.2f us (%d CTAs)\n", k0, 1e3 * time_ms([&] { chol_panel_kernel<<<(n_pad - k0 - NB) / PR + 1, 128, kPanelSmem>>>(A, ld, k0, rd, rows_total); }, reps), (n_pad - k0 - NB) / PR + 1);
  {
    long long hp[16];
    chol_panel3_kernel<<<94, 64, kPanel3Smem>>>(A, ld, 0, rd, rows_total);
    cudaMemcpyFromSymbol(hp, g_probe, sizeof(hp));
    printf("panel3 CTA 0 (cycles): load %lld, blocks", hp[1] - hp[0]);
    for (int q = 0; q < 8; ++q) printf(" %lld", hp[2 + q] - hp[1 + q]);
    chol_diag3_kernel<<<1, 64>>>(A, ld, 4096, rd, d_fail);
    cudaMemcpyFromSymbol(hp, g_probe, sizeof(hp));
    printf("\ndiag3 (cycles): load %lld, blocks", hp[1] - hp[0]);
    for (int q = 0; q < 8; ++q) printf(" %lld", hp[2 + q] - hp[1 + q]);
    printf("\n");
  }
  printf("SM clock right after the panel loops: %.0f MHz\n", probe_mhz(d_out));
  for (int k0 : {0, 3008, 5760}) {
    const int rows = rows_total - (k0 + NB), tr = (rows + NB - 1) / NB;
    printf("strip k0=%4d     %.2f us (%d CTAs)\n", k0, 1e3 * time_ms([&] { chol_update_kernel<NB, NB><<<tr, 256, STAGES * (NB + NB) * LDK * sizeof(double)>>>(A, ld, k0, NB, k0 + NB, k0 + NB, rows_total, 1); }, reps), tr);
  }
  for (int k0 : {0, 3008, 5760}) {
    const int rows = rows_total - (k0 + NB), tr = (rows + NB - 1) / NB;
    printf("strip64 k0=%4d   %.2f us (%d CTAs)\n", k0, 1e3 * time_ms([&] { chol_update64_kernel<64><<<tr, 256, kUpd64Smem>>>(A, ld, k0, NB, k0 + NB, k0 + NB, rows_total, nullptr); }, reps), tr);
  }
  for (int b : {2, 24, 45}) {
    const int k0 = b * OB, rows = rows_total - k0;
    const dim3 grid((rows + NB - 1) / NB, OB / NB);
    printf("L(b) old b=%2d     %.2f us (%d CTAs)\n", b, 1e3 * time_ms([&] { chol_update_kernel<NB, NB><<<grid, 256, STAGES * (NB + NB) * LDK * sizeof(double)>>>(A, ld, (b - 2) * OB, 2 * OB, k0, k0, rows_total, 1); }, reps), grid.x * grid.y);
    printf("L(b) new b=%2d     %.2f us (%d CTAs)\n", b, 1e3 * time_ms([&] { chol_update64_kernel<64><<<grid, 256, kUpd64Smem>>>(A, ld, (b - 2) * OB, 2 * OB, k0, k0, rows_total, nullptr); }, reps), grid.x * grid.y);
    const dim3 g32((rows + 31) / 32, OB / NB);
    printf("L(b) 32r b=%2d     %.2f us (%d CTAs)\n", b, 1e3 * time_ms([&] { chol_update64_kernel<32><<<g32, 256, kUpd64Smem>>>(A, ld, (b - 2) * OB, 2 * OB, k0, k0, rows_total, nullptr); }, reps), g32.x * g32.y);
  }
  for (int ob : {0, 23, 44}) {
    const int k0 = ob * OB, nt = n_pad / OB - ob - 1;
    printf("colupd ob=%2d      %.2f us (%d CTAs)\n", ob, 1e3 * time_ms([&] { chol_update_kernel<OB, NB><<<2 * nt + 1, 256, STAGES * (NB + OB) * LDK * sizeof(double)>>>(A, ld, k0, OB, k0 + OB, k0 + OB, rows_total, 1); }, reps), 2 * nt + 1);
  }
  printf("SM clock right after the colupd loops: %.0f MHz\n", probe_mhz(d_out));
  for (int ob : {0, 10, 23, 36, 43}) {
    const int k0 = ob * OB, ntr = n_pad / OB - ob - 2;
    const int tiles = ntr * (ntr + 1) + 2 * ntr;
    cudaFuncSetAttribute(chol_update2_kernel<true, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpd2Smem);
    const float ms = time_ms([&] { chol_update2_kernel<true, true, 128><<<tiles, 128, kUpd2Smem>>>(A, ld, k0, OB, k0 + 2 * OB, rows_total); }, blocks_only ? 1 : 20);
    printf("update2 ob=%2d     %.2f us (%d tiles) %.2f TFLOP/s\n", ob, 1e3 * ms, tiles, tiles * 128.0 * 64 * 128 * 2 / ms / 1e9);
    cudaFuncSetAttribute(chol_update2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpd2Smem);
    cudaFuncSetAttribute(chol_update2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpd2Smem);
    const float msn = time_ms([&] { chol_update2_kernel<false, false><<<tiles, 128, kUpd2Smem>>>(A, ld, k0, OB, k0 + 2 * OB, rows_total); }, blocks_only ? 1 : 20);
    printf("update2 noC ob=%2d %.2f us (%d tiles) %.2f TFLOP/s\n", ob, 1e3 * msn, tiles, tiles * 128.0 * 64 * 128 * 2 / msn / 1e9);
    cudaFuncSetAttribute(chol_update2_kernel<true, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpd2Smem);
    const float ms8 = time_ms([&] { chol_update2_kernel<true, true, 256><<<tiles, 256, kUpd2Smem>>>(A, ld, k0, OB, k0 + 2 * OB, rows_total); }, blocks_only ? 1 : 20);
    printf("update2 8w  ob=%2d %.2f us (%d tiles) %.2f TFLOP/s\n", ob, 1e3 * ms8, tiles, tiles * 128.0 * 64 * 128 * 2 / ms8 / 1e9);
    const float msl = time_ms([&] { chol_update2_kernel<true, false><<<tiles, 128, kUpd2Smem>>>(A, ld, k0, OB, k0 + 2 * OB, rows_total); }, blocks_only ? 1 : 20);
    printf("update2 ldC ob=%2d %.2f us (%d tiles) %.2f TFLOP/s\n", ob, 1e3 * msl, tiles, tiles * 128.0 * 64 * 128 * 2 / msl / 1e9);
    const int tiles1 = ntr * (ntr + 1) / 2 + ntr;
    const float ms1 = time_ms([&] { chol_update_kernel<OB, OB><<<tiles1, 256, STAGES * (OB + OB) * LDK * sizeof(double)>>>(A, ld, k0, OB, k0 + 2 * OB, k0 + 2 * OB, rows_total, 0); }, blocks_only ? 1 : 20);
    printf("update1 ob=%2d     %.2f us (%d tiles) %.2f TFLOP/s\n", ob, 1e3 * ms1, tiles1, tiles1 * 128.0 * 128 * 128 * 2 / ms1 / 1e9);
  }
  for (int KT : {128, 256, 512, 1024}) {
    const int ntr = 30, tiles = ntr * (ntr + 1) + 2 * ntr, base = 2048;
    cudaFuncSetAttribute(chol_update2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpd2Smem);
    const float msn = time_ms([&] { chol_update2_kernel<false, false><<<tiles, 128, kUpd2Smem>>>(A, ld, 0, KT, base, rows_total); }, blocks_only ? 1 : 10);
    const float msc = time_ms([&] { chol_update2_kernel<true, true, 128><<<tiles, 128, kUpd2Smem>>>(A, ld, 0, KT, base, rows_total); }, blocks_only ? 1 : 10);
    printf("update2 K=%4d (%d tiles): no C %.2f TFLOP/s, with C %.2f TFLOP/s\n", KT, tiles, tiles * 128.0 * 64 * KT * 2 / msn / 1e9, tiles * 128.0 * 64 * KT * 2 / msc / 1e9);
  }
  printf("SM clock right after the update loops: %.0f MHz\n", probe_mhz(d_out));
  return 0;
}
