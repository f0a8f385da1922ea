// Host side of the BA hot path: the C-ABI of include/theia_b200.h over the kernels of ba_kernels.cuh.
//
// Stands behind theia::BundleAdjuster::Optimize -> ceres::Solve
// (/root/reference/src/theia/sfm/bundle_adjustment/bundle_adjuster.cc:315-355). The trust-region policy
// (Levenberg-Marquardt, Jacobi scaling, step acceptance, termination tests) restates what ceres::Solve
// does with the options Theia passes (bundle_adjuster.cc:63-89).
// The control loop runs on the host; every O(observations) operation is a kernel on the caller's stream.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <set>
#include <tuple>
#include <string>
#include <vector>

#include "ba_kernels.cuh"
#include "track_ba.cuh"
#include "inner_iter.cuh"
#include "ba_setup.cuh"
#include "dense_chol.cuh"
#include "schur_pcg.cuh"

namespace thb {
namespace {

thread_local std::string g_last_error;

// Stream-ordered allocation from the device's default memory pool. The pool keeps freed blocks (release threshold =
// max), so a solve after the first one pays no cudaMalloc/cudaFree: BundleAdjustTrack-style callers create thousands of
// short sessions.
void ConfigurePoolOnce() {
  static std::once_flag once[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  std::call_once(once[dev], [dev] {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  });
}

struct Arena {  // every device buffer of a session; released in one go on the session's stream
  std::vector<void*> blocks;
  cudaStream_t st = nullptr;
  template <typename T>
  int Get(T** p, size_t n) {
    *p = nullptr;
    void* q = nullptr;
    THB_CUDA_CHECK(cudaMallocAsync(&q, std::max<size_t>(1, n) * sizeof(T), st));
    blocks.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return THB_OK;
  }
  void Release() {
    for (void* q : blocks) cudaFreeAsync(q, st);
    blocks.clear();
  }
};

// Pinned host blocks for the per-iteration scalar read-back, recycled across sessions (cudaMallocHost is slow).
struct PinnedCache {
  std::mutex mu;
  std::vector<void*> free_blocks;
  static constexpr size_t kBytes = 1024;
  void* Get() {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (!free_blocks.empty()) { void* q = free_blocks.back(); free_blocks.pop_back(); return q; }
    }
    void* q = nullptr;
    if (cudaMallocHost(&q, kBytes) != cudaSuccess) return nullptr;
    return q;
  }
  void Put(void* q) {
    if (!q) return;
    std::lock_guard<std::mutex> lk(mu);
    free_blocks.push_back(q);
  }
};
PinnedCache g_pinned;

inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

struct PhaseTimer {  // CUDA-event timing of one phase on the solve stream
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  cudaStream_t st = nullptr;
  size_t used = 0;
  void Begin() {
    if (used == ev.size()) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ev.emplace_back(a, b);
    }
    cudaEventRecord(ev[used].first, st);
  }
  void End() { cudaEventRecord(ev[used].second, st); ++used; }
  double CollectMs() {  // call after a stream sync
    double ms = 0.0;
    for (size_t i = 0; i < used; ++i) { float t = 0.f; cudaEventElapsedTime(&t, ev[i].first, ev[i].second); ms += t; }
    used = 0;
    return ms;
  }
  void Free() { for (auto& e : ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); } ev.clear(); }
};

}  // namespace

void SetLastError(const std::string& s) { g_last_error = s; }
const char* GetLastError() { return g_last_error.c_str(); }

}  // namespace thb

using namespace thb;

struct ThbBaSession {
  ThbBaProblem prob;
  ThbBaOptions opt;
  cudaStream_t st = nullptr;
  int nc = 0, ng = 0, np = 0, no = 0;
  int PD = 3, model = -1, n_red = 0;
  int nvg = 0;                // intrinsics groups with free coordinates (slots of the reduced system)
  bool constrained = false;   // bounds on a non-constant block (ceres Program::IsBoundsConstrained)
  bool red_variable = false;  // any non-constant camera coordinate or intrinsics block
  bool any_variable = false;
  bool pt_variable = false;
  BaConst K{};
  BaState X{}, Xc{};
  ObsSoA Op{}, Oc{};
  // owned device memory
  int *d_cam_group = nullptr, *d_intr_model = nullptr, *d_intr_slot = nullptr;
  uint8_t *d_cam_const = nullptr, *d_pt_const = nullptr;
  uint8_t* d_has_prior = nullptr;  // camera priors: bit 0 position, bit 1 gravity, bit 2 orientation (ThbBaProblem::cam_has_*_prior); nullptr when the problem has none
  double* d_prior = nullptr;       // [nc][PRIOR_STRIDE]: position [sqrt information (9, row-major) | prior (3)], gravity, orientation [likewise]
  uint16_t* d_intr_const = nullptr;
  int *d_op_cam = nullptr, *d_op_pt = nullptr, *d_oc_cam = nullptr, *d_oc_pt = nullptr;
  double2 *d_op_xy = nullptr, *d_op_si = nullptr, *d_oc_xy = nullptr, *d_oc_si = nullptr;
  int *d_pt_start = nullptr, *d_cam_start = nullptr, *d_chunk_pt = nullptr;
  int nchunks = 0;
  double *d_r = nullptr, *d_jc = nullptr, *d_jp = nullptr, *d_ji = nullptr;
  int8_t* d_op_slot = nullptr;
  int* d_slot_group = nullptr;
  double *d_zt = nullptr, *d_ilo = nullptr, *d_ihi = nullptr;
  double *d_cs = nullptr, *d_ps = nullptr;
  double *d_vinv = nullptr, *d_gp = nullptr, *d_pdiag = nullptr, *d_braw = nullptr, *d_cdiag = nullptr, *d_yp = nullptr;
  double *d_jy = nullptr, *d_bsum = nullptr;  // back-substitution: Jc y_c per observation, Jp^T (Jc y_c) per point
  double* d_scal = nullptr;
  int* d_flag = nullptr;
  double* h_scal = nullptr;  // pinned
  int* h_flag = nullptr;     // pinned
  void* d_flush = nullptr;
  DenseChol chol;
  SchurPcg pcg;  // THB_SOLVER_SCHUR_PCG: conjugate gradients on the S that chol.A holds
  bool use_pcg = false;
  // trust-region state (ceres TrustRegionMinimizer / LevenbergMarquardtStrategy)
  double radius = 1e4, decrease_factor = 2.0;
  double x_cost = 0.0, x_norm = 0.0, min_cost = 0.0, fixed_cost = 0.0, gradient_max_norm = 0.0;
  int iteration = 0, num_consecutive_invalid = 0;
  bool step_is_successful = true, finished = false;
  bool grad_checked = false;
  ThbBaSummary sum{};
  PhaseTimer t_jac, t_normal, t_solve, t_update;
  std::chrono::steady_clock::time_point t_create, t_solve_start;
  Arena arena;
  void* h_block = nullptr;  // pinned block behind h_scal / h_flag
  // inner iterations (use_inner_iterations): ceres CoordinateDescentMinimizer state
  bool inner_enabled = false;
  double *d_bk_cam = nullptr, *d_bk_pts = nullptr, *d_bk_intr = nullptr;  // the candidate before the inner iterations
  double *d_inner_ps = nullptr, *d_pts_scratch = nullptr, *d_inner_out = nullptr, *d_inner_scale = nullptr;
  ThbTrackBaResult* d_inner_res = nullptr;
  std::vector<int> h_slot_group, h_model;
  std::vector<uint16_t> h_const;
  std::vector<double> h_ilo, h_ihi;
  int num_inner_steps = 0;
};

namespace {

void FreeSession(ThbBaSession* s) {
  if (!s) return;
  s->chol.Free(s->st);
  if (s->use_pcg) s->pcg.Free(s->st);
  s->arena.Release();
  g_pinned.Put(s->h_block);
  s->t_jac.Free(); s->t_normal.Free(); s->t_solve.Free(); s->t_update.Free();
  delete s;
}

// Copy `bytes` from the caller's array (host or device per memory_space) into a host vector.
template <typename T>
int FetchToHost(const void* src, size_t count, int space, std::vector<T>* out) {
  out->resize(count);
  if (count == 0) return THB_OK;
  if (space == THB_MEM_HOST) std::memcpy(out->data(), src, count * sizeof(T));
  else THB_CUDA_CHECK(cudaMemcpy(out->data(), src, count * sizeof(T), cudaMemcpyDeviceToHost));
  return THB_OK;
}

struct DevBufs {  // scoped device allocations
  std::vector<void*> p;
  ~DevBufs() { for (void* q : p) cudaFree(q); }
  template <typename T> T* get(size_t n) {
    void* q = nullptr;
    if (cudaMalloc(&q, std::max<size_t>(1, n) * sizeof(T)) != cudaSuccess) return nullptr;
    p.push_back(q);
    return (T*)q;
  }
};

int CheckDevice() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) THB_FAIL(THB_E_NO_DEVICE, "no CUDA device visible; libtheia_b200 has no CPU path");
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  THB_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) THB_FAIL(THB_E_NO_DEVICE, "device is not sm_100 (B200); kernels are built for sm_100a only");
  return THB_OK;
}

// K1 uses no shared memory: ask for the largest L1 so the 128-byte camera records stay resident.
// Function attributes are per (device, kernel): applied once each, safe under concurrent callers. Keyed by the kernel's
// ADDRESS - every instantiation of a kernel template has the same function-pointer type, so a per-type flag would set the
// attribute for the first instantiation only.
bool FirstUse(const void* kern, int what) {
  static std::mutex mu;
  static std::set<std::tuple<int, const void*, int>> seen;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  std::lock_guard<std::mutex> lk(mu);
  return seen.insert(std::make_tuple(dev, kern, what)).second;
}
template <typename Kern>
void PreferL1Once(Kern kern) {
  if (FirstUse(reinterpret_cast<const void*>(kern), 0)) cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
}
template <typename Kern>
void MaxDynamicSmemOnce(Kern kern, int bytes) {
  if (FirstUse(reinterpret_cast<const void*>(kern), 1)) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
int SmCount() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  static int cache[64] = {0};
  if (dev >= 0 && dev < 64 && cache[dev]) return cache[dev];
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  if (dev >= 0 && dev < 64) cache[dev] = sms;
  return sms;
}
// The shared-camera K1 (k_jacobian_sc) pays one copy of the camera table per SM; measured faster than the gather kernel from
// 50k observations (12.1 vs 15.9 us) to 1M (44.6 vs 58.9 us), profiles/r01_k1_ab_scale.txt: used whenever the table fits in
// shared memory and there are at least 32k observations. THB_K1_MODE=gather forces the L1-gather kernel (A/B timing),
// THB_K1_MODE=shared the shared-camera kernel whenever the table fits (parity tests at small sizes).
bool UseSharedCameraK1(const ThbBaSession* s) {
  if (s->nc > K1S_MAX_CAMS) return false;
  const char* e = getenv("THB_K1_MODE");
  if (e && !strcmp(e, "gather")) return false;
  if (e && !strcmp(e, "shared")) return true;
  return s->no >= 32768;
}
template <int MODEL, int PD, int NK, bool ROBUST>
void LaunchJacobianKernel(ThbBaSession* s, const double* cs, const double* ps, const double* is, double* ji) {
  // refined intrinsics (NK > 0) stay on the gather kernel: the shared-camera instantiation spills ~400 bytes there
  if constexpr (NK == 0) {
    if (UseSharedCameraK1(s)) {
      MaxDynamicSmemOnce(k_jacobian_sc<MODEL, PD, NK, ROBUST>, 227 * 1024);
      k_jacobian_sc<MODEL, PD, NK, ROBUST><<<std::min(SmCount(), cdiv(s->no, K1S_THREADS)), K1S_THREADS, K1S_PT_BYTES + K1S_INTR_BYTES + (size_t)s->nc * CAMD * 8, s->st>>>(
          s->K, s->X, s->Op, cs, ps, is, s->d_r, s->d_jc, s->d_jp, ji, s->d_scal, s->d_flag);
      return;
    }
  }
  PreferL1Once(k_jacobian<MODEL, PD, NK, ROBUST>);
  k_jacobian<MODEL, PD, NK, ROBUST><<<cdiv(s->no, 128 * K1_OBS_PER_THREAD), 128, 0, s->st>>>(s->K, s->X, s->Op, cs, ps, is, s->d_r, s->d_jc, s->d_jp, ji, s->d_scal, s->d_flag);
}
template <int MODEL, int PD>
void LaunchJacobian(ThbBaSession* s, const double* cs, const double* ps) {
  if (s->opt.loss_function_type == THB_LOSS_TRIVIAL) LaunchJacobianKernel<MODEL, PD, 0, false>(s, cs, ps, nullptr, nullptr);
  else LaunchJacobianKernel<MODEL, PD, 0, true>(s, cs, ps, nullptr, nullptr);
}
template <int PD>
void DispatchJacobian(ThbBaSession* s, const double* cs, const double* ps) {
  switch (s->model) {
    case THB_MODEL_PINHOLE: LaunchJacobian<THB_MODEL_PINHOLE, PD>(s, cs, ps); break;
    case THB_MODEL_FISHEYE: LaunchJacobian<THB_MODEL_FISHEYE, PD>(s, cs, ps); break;
    case THB_MODEL_FOV: LaunchJacobian<THB_MODEL_FOV, PD>(s, cs, ps); break;
    case THB_MODEL_DIVISION_UNDISTORTION: LaunchJacobian<THB_MODEL_DIVISION_UNDISTORTION, PD>(s, cs, ps); break;
    case THB_MODEL_DOUBLE_SPHERE: LaunchJacobian<THB_MODEL_DOUBLE_SPHERE, PD>(s, cs, ps); break;
    case THB_MODEL_EXTENDED_UNIFIED: LaunchJacobian<THB_MODEL_EXTENDED_UNIFIED, PD>(s, cs, ps); break;
    default: LaunchJacobian<-1, PD>(s, cs, ps); break;
  }
}
template <int PD>
void LaunchJacobianIntr(ThbBaSession* s, const double* cs, const double* ps) {
  LaunchJacobianKernel<-1, PD, NI, true>(s, cs, ps, cs + 6 * s->nc, s->d_ji);
}
void RunJacobian(ThbBaSession* s, const double* cs, const double* ps) {
  if (s->nvg > 0) { if (s->PD == 3) LaunchJacobianIntr<3>(s, cs, ps); else LaunchJacobianIntr<4>(s, cs, ps); }
  else if (s->PD == 3) DispatchJacobian<3>(s, cs, ps); else DispatchJacobian<4>(s, cs, ps);
  ++s->sum.gpu_launches;
}

template <int MODEL, int PD>
void LaunchCamPass(ThbBaSession* s, double inv_radius) {
  k_cam_pass<MODEL, PD><<<s->nc, 128, 0, s->st>>>(s->K, s->X, s->Oc, s->d_cam_start, s->d_cs, s->d_ps, s->d_vinv, s->d_gp, inv_radius,
                                                 s->opt.min_lm_diagonal, s->opt.max_lm_diagonal, s->chol.A, s->chol.ld,
                                                 s->chol.RhsRow(), s->d_braw, s->d_cdiag, s->d_flag, s->d_has_prior, s->d_prior);
}
template <int PD>
void DispatchCamPass(ThbBaSession* s, double inv_radius) {
  switch (s->model) {
    case THB_MODEL_PINHOLE: LaunchCamPass<THB_MODEL_PINHOLE, PD>(s, inv_radius); break;
    case THB_MODEL_FISHEYE: LaunchCamPass<THB_MODEL_FISHEYE, PD>(s, inv_radius); break;
    case THB_MODEL_FOV: LaunchCamPass<THB_MODEL_FOV, PD>(s, inv_radius); break;
    case THB_MODEL_DIVISION_UNDISTORTION: LaunchCamPass<THB_MODEL_DIVISION_UNDISTORTION, PD>(s, inv_radius); break;
    case THB_MODEL_DOUBLE_SPHERE: LaunchCamPass<THB_MODEL_DOUBLE_SPHERE, PD>(s, inv_radius); break;
    case THB_MODEL_EXTENDED_UNIFIED: LaunchCamPass<THB_MODEL_EXTENDED_UNIFIED, PD>(s, inv_radius); break;
    default: LaunchCamPass<-1, PD>(s, inv_radius); break;
  }
}

// position priors are residual blocks of their own: their cost joins whatever slot the reprojection blocks were summed into
void AddPriorCost(ThbBaSession* s, const BaState& st, int slot) {
  if (!s->d_has_prior) return;
  k_prior_cost<<<cdiv(s->nc, 128), 128, 0, s->st>>>(s->nc, s->d_has_prior, s->d_prior, s->d_cam_const, st.camd, s->d_scal + slot);
  ++s->sum.gpu_launches;
}

void RunCost(ThbBaSession* s, const BaState& st, int slot, int flag_slot) {
  const int g = cdiv(s->no, 256);
  switch (s->model) {
    case THB_MODEL_PINHOLE: k_cost<THB_MODEL_PINHOLE><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    case THB_MODEL_FISHEYE: k_cost<THB_MODEL_FISHEYE><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    case THB_MODEL_FOV: k_cost<THB_MODEL_FOV><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    case THB_MODEL_DIVISION_UNDISTORTION: k_cost<THB_MODEL_DIVISION_UNDISTORTION><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    case THB_MODEL_DOUBLE_SPHERE: k_cost<THB_MODEL_DOUBLE_SPHERE><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    case THB_MODEL_EXTENDED_UNIFIED: k_cost<THB_MODEL_EXTENDED_UNIFIED><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
    default: k_cost<-1><<<g, 256, 0, s->st>>>(s->K, st, s->Op, s->d_scal, slot, s->d_flag, flag_slot); break;
  }
  ++s->sum.gpu_launches;
  AddPriorCost(s, st, slot);
}

int ReadScalars(ThbBaSession* s) {
  THB_CUDA_CHECK(cudaMemcpyAsync(s->h_scal, s->d_scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s->st));
  THB_CUDA_CHECK(cudaMemcpyAsync(s->h_flag, s->d_flag, sizeof(int) * FL_COUNT, cudaMemcpyDeviceToHost, s->st));
  THB_CUDA_CHECK(cudaStreamSynchronize(s->st));
  s->sum.num_linear_solver_iterations = s->h_flag[FL_PCG_ITERS];
  s->sum.ms_jacobian += s->t_jac.CollectMs();
  s->sum.ms_normal += s->t_normal.CollectMs();
  s->sum.ms_solve += s->t_solve.CollectMs();
  s->sum.ms_update += s->t_update.CollectMs();
  return THB_OK;
}

void LogIter(ThbBaSession* s) {
  if (s->sum.iter_log_count < THB_MAX_ITER_LOG) {
    s->sum.iter_cost[s->sum.iter_log_count] = s->x_cost + s->fixed_cost;
    s->sum.iter_radius[s->sum.iter_log_count] = s->radius;
    ++s->sum.iter_log_count;
  }
}

// K1 at the current point: cost, residual/Jacobian planes (column-scaled).
int EvaluateJacobian(ThbBaSession* s) {
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_COST_X, 0, sizeof(double), s->st));
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_flag + FL_EVAL_X, 0, sizeof(int), s->st));
  s->t_jac.Begin();
  RunJacobian(s, s->d_cs, s->d_ps);
  s->t_jac.End();
  AddPriorCost(s, s->X, SC_COST_X);
  ++s->sum.num_jacobian_evaluations;
  return THB_OK;
}

// K2a + K2b (+ gradient max-norm) at the current Jacobian and radius.
int BuildBlocks(ThbBaSession* s, bool want_grad) {
  const double inv_radius = 1.0 / s->radius;
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_flag + FL_CHOL, 0, sizeof(int) * 2, s->st));
  s->t_normal.Begin();
  if (s->red_variable) { if (s->chol.Clear(s->st) != THB_OK) return THB_E_CUDA; ++s->sum.gpu_launches; }
  if (s->PD == 3)
    k_point_pass<3><<<cdiv(s->np, 128), 128, 0, s->st>>>(s->np, s->no, s->d_pt_start, s->d_r, s->d_jp, inv_radius, s->opt.min_lm_diagonal,
                                                        s->opt.max_lm_diagonal, s->d_vinv, s->d_gp, s->d_pdiag, s->d_flag);
  else
    k_point_pass<4><<<cdiv(s->np, 128), 128, 0, s->st>>>(s->np, s->no, s->d_pt_start, s->d_r, s->d_jp, inv_radius, s->opt.min_lm_diagonal,
                                                        s->opt.max_lm_diagonal, s->d_vinv, s->d_gp, s->d_pdiag, s->d_flag);
  if (s->PD == 3) DispatchCamPass<3>(s, inv_radius); else DispatchCamPass<4>(s, inv_radius);
  s->sum.gpu_launches += 2;
  if (s->nvg > 0) {
    const int base = 6 * s->nc, nvg = s->nvg;
    THB_CUDA_CHECK(cudaMemsetAsync(s->d_braw + base, 0, sizeof(double) * NI * nvg, s->st));
    THB_CUDA_CHECK(cudaMemsetAsync(s->d_cdiag + base, 0, sizeof(double) * NI * nvg, s->st));
    const dim3 gp_(cdiv(s->np, 128), nvg), gd_(std::min(cdiv(s->no, 256), 296), nvg), gi_(std::min(cdiv(s->np, 256), 148), nvg * (nvg + 1) / 2, NI),
        gc_(s->nc, nvg);
    if (s->PD == 3) {
      k_intr_point<3><<<gp_, 128, 0, s->st>>>(s->np, s->no, nvg, base, s->d_pt_start, s->d_op_slot, s->d_jp, s->d_ji, s->d_vinv, s->d_gp, s->d_zt, s->chol.RhsRow());
      k_intr_intr<3><<<gi_, 256, 0, s->st>>>(s->np, nvg, base, s->d_zt, s->chol.A, s->chol.ld);
      k_cam_intr_pass<3><<<gc_, 128, 0, s->st>>>(s->K, s->X, s->Oc, s->d_cam_start, s->d_cs, s->d_ps, nvg, s->d_zt, s->chol.A, s->chol.ld);
    } else {
      k_intr_point<4><<<gp_, 128, 0, s->st>>>(s->np, s->no, nvg, base, s->d_pt_start, s->d_op_slot, s->d_jp, s->d_ji, s->d_vinv, s->d_gp, s->d_zt, s->chol.RhsRow());
      k_intr_intr<4><<<gi_, 256, 0, s->st>>>(s->np, nvg, base, s->d_zt, s->chol.A, s->chol.ld);
      k_cam_intr_pass<4><<<gc_, 128, 0, s->st>>>(s->K, s->X, s->Oc, s->d_cam_start, s->d_cs, s->d_ps, nvg, s->d_zt, s->chol.A, s->chol.ld);
    }
    k_intr_direct<<<gd_, 256, 0, s->st>>>(s->no, base, s->d_op_slot, s->d_r, s->d_ji, s->chol.A, s->chol.ld, s->chol.RhsRow(), s->d_braw, s->d_cdiag);
    k_intr_finalize<<<cdiv(nvg * NI, 128), 128, 0, s->st>>>(nvg, base, s->d_slot_group, s->d_intr_model, s->d_intr_const, s->d_cdiag, inv_radius,
                                                          s->opt.min_lm_diagonal, s->opt.max_lm_diagonal, s->chol.A, s->chol.ld);
    s->sum.gpu_launches += 5;
  }
  if (want_grad) {
    THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_GRADMAX, 0, sizeof(double), s->st));
    if (s->constrained) {
      const int tot = s->nc + s->np + s->ng;
      if (s->PD == 3) k_grad_proj<3><<<cdiv(tot, 256), 256, 0, s->st>>>(s->K, s->X, s->d_braw, s->d_gp, s->d_cs, s->d_ps, s->d_ilo, s->d_ihi, s->d_scal);
      else k_grad_proj<4><<<cdiv(tot, 256), 256, 0, s->st>>>(s->K, s->X, s->d_braw, s->d_gp, s->d_cs, s->d_ps, s->d_ilo, s->d_ihi, s->d_scal);
      ++s->sum.gpu_launches;
    } else {
      k_grad_max<<<cdiv(s->n_red, 256), 256, 0, s->st>>>(s->n_red, s->d_braw, s->d_cs, s->d_scal);
      k_grad_max<<<cdiv((long long)s->np * s->PD, 256), 256, 0, s->st>>>(s->np * s->PD, s->d_gp, s->d_ps, s->d_scal);
      s->sum.gpu_launches += 2;
    }
  }
  s->t_normal.End();
  return THB_OK;
}

// candidate = Plus(x, alpha * delta) and its cost (ComputeCandidatePointAndEvaluateCost).
int ComputeCandidate(ThbBaSession* s, double alpha) {
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_COST_CAND, 0, sizeof(double) * 4, s->st));  // cand, step2, xnew2, gtd
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_flag + FL_EVAL_CAND, 0, sizeof(int), s->st));
  if (s->PD == 3) k_update_pts<3><<<cdiv(s->np, 128), 128, 0, s->st>>>(s->np, s->d_pt_const, s->X.pts, s->d_yp, s->d_gp, s->d_ps, alpha, s->Xc.pts, s->d_scal);
  else k_update_pts<4><<<cdiv(s->np, 128), 128, 0, s->st>>>(s->np, s->d_pt_const, s->X.pts, s->d_yp, s->d_gp, s->d_ps, alpha, s->Xc.pts, s->d_scal);
  k_update_cams<<<cdiv(s->nc, 128), 128, 0, s->st>>>(s->nc, s->d_cam_const, s->X.cam, s->chol.x, s->d_braw, s->d_cs, alpha, s->Xc.cam, s->d_scal);
  if (s->nvg > 0) {
    k_update_intr<<<cdiv(s->ng, 128), 128, 0, s->st>>>(s->ng, s->nc, s->d_intr_slot, s->d_intr_model, s->X.intr, s->chol.x, s->d_braw, s->d_cs, s->d_ilo,
                                                      s->d_ihi, alpha, s->Xc.intr, s->d_scal);
    ++s->sum.gpu_launches;
  }
  k_cam_derive<<<cdiv(s->nc, 128), 128, 0, s->st>>>(s->Xc.cam, s->Xc.camd, s->nc, s->d_cs, s->d_cam_const, s->d_cam_group);
  s->sum.gpu_launches += 3;
  RunCost(s, s->Xc, SC_COST_CAND, FL_EVAL_CAND);
  ++s->sum.num_cost_evaluations;
  return THB_OK;
}

// Schur complement off-diagonal blocks, factor + solve, back-substitution, candidate, candidate cost.
int SolveAndStep(ThbBaSession* s) {
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_MCC, 0, sizeof(double), s->st));
  if (s->red_variable) {
    s->t_normal.Begin();
    static const bool k3_red = []() { const char* m = getenv("THB_K3_MODE"); return m && std::string(m) == "red"; }();  // A/B: the r01 scalar-RED form
    if (s->PD == 3) {
      if (k3_red) k_schur_offdiag<3, false><<<s->nchunks, 128, 0, s->st>>>(s->no, s->d_chunk_pt, s->d_pt_start, s->d_op_cam, s->d_jc, s->d_jp, s->d_vinv, s->chol.A, s->chol.ld);
      else k_schur_offdiag<3, true><<<s->nchunks, 128, 0, s->st>>>(s->no, s->d_chunk_pt, s->d_pt_start, s->d_op_cam, s->d_jc, s->d_jp, s->d_vinv, s->chol.A, s->chol.ld);
    } else {
      if (k3_red) k_schur_offdiag<4, false><<<s->nchunks, 128, 0, s->st>>>(s->no, s->d_chunk_pt, s->d_pt_start, s->d_op_cam, s->d_jc, s->d_jp, s->d_vinv, s->chol.A, s->chol.ld);
      else k_schur_offdiag<4, true><<<s->nchunks, 128, 0, s->st>>>(s->no, s->d_chunk_pt, s->d_pt_start, s->d_op_cam, s->d_jc, s->d_jp, s->d_vinv, s->chol.A, s->chol.ld);
    }
    ++s->sum.gpu_launches;
    s->t_normal.End();
    s->t_solve.Begin();
    if (s->use_pcg) {
      if (s->pcg.Solve(s->st, s->chol.A, s->chol.ld, s->chol.RhsRow(), s->chol.x, s->opt.pcg_eta, s->opt.pcg_max_iterations, s->d_flag + FL_CHOL,
                       s->d_flag + FL_PCG_ITERS, &s->sum.gpu_launches) != THB_OK) return THB_E_CUDA;
    } else if (s->chol.FactorAndSolve(s->st, s->d_flag + FL_CHOL, &s->sum.gpu_launches) != THB_OK) return THB_E_CUDA;
    s->t_solve.End();
  } else {
    THB_CUDA_CHECK(cudaMemsetAsync(s->chol.x, 0, sizeof(double) * s->chol.n_pad, s->st));
  }
  ++s->sum.num_linear_solves;
  s->t_update.Begin();
  {
    const int gq = cdiv(s->no, 256), gp_ = cdiv(s->np, 128);
    THB_CUDA_CHECK(cudaMemsetAsync(s->d_bsum, 0, sizeof(double) * (size_t)s->np * s->PD, s->st));
    if (s->nvg > 0) {
      if (s->PD == 3) k_backsub_obs1<3, NI><<<gq, 256, 0, s->st>>>(s->no, s->nc, s->d_op_cam, s->d_op_pt, s->d_op_slot, s->d_jc, s->d_jp, s->d_ji, s->chol.x, s->d_jy, s->d_bsum);
      else k_backsub_obs1<4, NI><<<gq, 256, 0, s->st>>>(s->no, s->nc, s->d_op_cam, s->d_op_pt, s->d_op_slot, s->d_jc, s->d_jp, s->d_ji, s->chol.x, s->d_jy, s->d_bsum);
    } else {
      if (s->PD == 3) k_backsub_obs1<3, 0><<<gq, 256, 0, s->st>>>(s->no, s->nc, s->d_op_cam, s->d_op_pt, nullptr, s->d_jc, s->d_jp, nullptr, s->chol.x, s->d_jy, s->d_bsum);
      else k_backsub_obs1<4, 0><<<gq, 256, 0, s->st>>>(s->no, s->nc, s->d_op_cam, s->d_op_pt, nullptr, s->d_jc, s->d_jp, nullptr, s->chol.x, s->d_jy, s->d_bsum);
    }
    if (s->PD == 3) {
      k_backsub_pt<3><<<gp_, 128, 0, s->st>>>(s->np, s->d_vinv, s->d_gp, s->d_bsum, s->d_yp);
      k_backsub_obs2<3><<<gq, 256, 0, s->st>>>(s->no, s->d_op_pt, s->d_r, s->d_jp, s->d_jy, s->d_yp, s->d_scal);
    } else {
      k_backsub_pt<4><<<gp_, 128, 0, s->st>>>(s->np, s->d_vinv, s->d_gp, s->d_bsum, s->d_yp);
      k_backsub_obs2<4><<<gq, 256, 0, s->st>>>(s->no, s->d_op_pt, s->d_r, s->d_jp, s->d_jy, s->d_yp, s->d_scal);
    }
    s->sum.gpu_launches += 3;
    if (s->d_has_prior) {
      k_prior_mcc<<<cdiv(s->nc, 128), 128, 0, s->st>>>(s->nc, s->d_has_prior, s->d_prior, s->d_cam_const, s->X.camd, s->d_cs, s->chol.x, s->d_scal);
      ++s->sum.gpu_launches;
    }
  }
  const int rc = ComputeCandidate(s, 1.0);
  s->t_update.End();
  return rc;
}


// ---- inner iterations (ceres CoordinateDescentMinimizer, inner_iter.cuh) on the candidate Xc ------------------------------
template <int PD>
int InnerIntrEvaluate(ThbBaSession* s, int sl, bool want_j, double* h_out) {
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_inner_out, 0, sizeof(double) * II_VALS, s->st));
  const int grid = std::min(cdiv(s->no, 256), 4 * SmCount());
  if (want_j) k_inner_intr<PD, true><<<grid, 256, 0, s->st>>>(s->K, s->Xc, s->Op, s->d_op_slot, sl, s->d_inner_scale, s->d_inner_out);
  else k_inner_intr<PD, false><<<grid, 256, 0, s->st>>>(s->K, s->Xc, s->Op, s->d_op_slot, sl, s->d_inner_scale, s->d_inner_out);
  ++s->sum.gpu_launches;
  THB_CUDA_CHECK(cudaMemcpyAsync(h_out, s->d_inner_out, sizeof(double) * II_VALS, cudaMemcpyDeviceToHost, s->st));
  THB_CUDA_CHECK(cudaStreamSynchronize(s->st));
  return THB_OK;
}

// TrustRegionMinimizer on one shared intrinsics block, driven from the host (a block owns up to every observation: each
// evaluation is a full-grid launch). Mirrors k_inner_cam / the oracle's InnerSolve.
template <int PD>
int InnerIntrSolve(ThbBaSession* s, int sl) {
  const InnerLmParams P;
  const int g = s->h_slot_group[sl];
  const int Kg = num_intrinsics(s->h_model[g]);
  double x[KS], cand[KS], scale[NI], out[II_VALS];
  double* d_x = s->Xc.intr + (size_t)g * KS;
  THB_CUDA_CHECK(cudaMemcpyAsync(x, d_x, sizeof(double) * KS, cudaMemcpyDeviceToHost, s->st));
  THB_CUDA_CHECK(cudaStreamSynchronize(s->st));
  bool free_k[NI];
  for (int k = 0; k < NI; ++k) { free_k[k] = k < Kg && !((s->h_const[g] >> k) & 1); scale[k] = 1.0; }
  int rc;
  THB_CUDA_CHECK(cudaMemcpyAsync(s->d_inner_scale, scale, sizeof(double) * NI, cudaMemcpyHostToDevice, s->st));
  if ((rc = InnerIntrEvaluate<PD>(s, sl, true, out)) != THB_OK) return rc;
  if (out[55] > 0.0) return THB_OK;  // IterationZero failed: block untouched
  { int e = 0; for (int a = 0; a < NI; ++a) { e += a; scale[a] = free_k[a] ? 1.0 / (1.0 + std::sqrt(out[e])) : 0.0; ++e; } }
  THB_CUDA_CHECK(cudaMemcpyAsync(s->d_inner_scale, scale, sizeof(double) * NI, cudaMemcpyHostToDevice, s->st));
  if ((rc = InnerIntrEvaluate<PD>(s, sl, true, out)) != THB_OK) return rc;
  double H[45], gr[NI], diag[NI], x_cost = out[54], x_norm = 0.0, radius = P.radius0, decrease_factor = 2.0;
  for (int k = 0; k < 45; ++k) H[k] = out[k];
  for (int k = 0; k < NI; ++k) gr[k] = out[45 + k];
  for (int k = 0; k < Kg; ++k) x_norm += x[k] * x[k];
  x_norm = std::sqrt(x_norm);
  bool step_ok = true, reuse_diag = false, dirty = false;  // dirty: the device holds a rejected candidate
  int iteration = 0, invalid = 0;
  for (;;) {
    if (iteration >= P.max_num_iterations) break;
    if (step_ok) {
      double gmax = 0.0;
      for (int k = 0; k < NI; ++k) if (free_k[k]) gmax = std::max(gmax, std::fabs(gr[k] / scale[k]));
      if (gmax <= P.gtol) break;
    }
    if (radius <= P.min_radius) break;
    ++iteration;
    step_ok = false;
    if (!reuse_diag) { int e = 0; for (int a = 0; a < NI; ++a) { e += a; diag[a] = std::min(std::max(H[e], P.min_diag), P.max_diag); ++e; } }
    reuse_diag = true;
    double Mx[NI * NI], y[NI], mcc = 0.0;
    { int e = 0; for (int a = 0; a < NI; ++a) for (int b = 0; b <= a; ++b) { Mx[a * NI + b] = H[e]; Mx[b * NI + a] = H[e]; ++e; } }
    for (int a = 0; a < NI; ++a) Mx[a * NI + a] += diag[a] / radius;
    bool valid = spd_solve<NI>(Mx, gr, y);
    if (valid) {
      double yg = 0.0, yHy = 0.0;
      int e = 0;
      for (int a = 0; a < NI; ++a) { yg += y[a] * gr[a]; for (int b = 0; b <= a; ++b) { yHy += (a == b ? 1.0 : 2.0) * y[a] * H[e] * y[b]; ++e; } }
      mcc = yg - 0.5 * yHy;
      valid = std::isfinite(mcc) && mcc > 0.0;
    }
    if (!valid) {
      if (++invalid >= P.max_invalid) break;
      radius /= decrease_factor; decrease_factor *= 2.0;
      continue;
    }
    invalid = 0;
    for (int k = 0; k < KS; ++k) cand[k] = x[k];
    for (int k = 0; k < NI; ++k)  // Plus, then the box projection of ParameterBlock::Plus (bounds of bundle_adjuster.cc:396-427)
      if (k < KS) { if (free_k[k]) cand[k] = x[k] + (-y[k] * scale[k]); cand[k] = std::min(std::max(cand[k], s->h_ilo[NI * sl + k]), s->h_ihi[NI * sl + k]); }
    THB_CUDA_CHECK(cudaMemcpyAsync(d_x, cand, sizeof(double) * KS, cudaMemcpyHostToDevice, s->st));
    dirty = true;
    if ((rc = InnerIntrEvaluate<PD>(s, sl, false, out)) != THB_OK) return rc;
    const double cand_cost = out[55] > 0.0 ? std::numeric_limits<double>::max() : out[54];
    double sn = 0.0, cn = 0.0;
    for (int k = 0; k < Kg; ++k) { sn += (cand[k] - x[k]) * (cand[k] - x[k]); cn += cand[k] * cand[k]; }
    if (std::sqrt(sn) <= P.ptol * (x_norm + P.ptol)) break;
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= P.ftol * x_cost) break;
    const double rel = cand_cost >= std::numeric_limits<double>::max() ? std::numeric_limits<double>::lowest() : cost_change / mcc;
    if (rel > P.min_relative_decrease) {
      for (int k = 0; k < KS; ++k) x[k] = cand[k];
      dirty = false;
      x_norm = std::sqrt(cn);
      if ((rc = InnerIntrEvaluate<PD>(s, sl, true, out)) != THB_OK) return rc;
      if (out[55] > 0.0) break;
      for (int k = 0; k < 45; ++k) H[k] = out[k];
      for (int k = 0; k < NI; ++k) gr[k] = out[45 + k];
      x_cost = out[54];
      step_ok = true;
      const double u = 2.0 * rel - 1.0;
      radius = std::min(P.max_radius, radius / std::max(1.0 / 3.0, 1.0 - u * u * u));
      decrease_factor = 2.0; reuse_diag = false;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0;
    }
  }
  if (dirty) {
    THB_CUDA_CHECK(cudaMemcpyAsync(d_x, x, sizeof(double) * KS, cudaMemcpyHostToDevice, s->st));
    THB_CUDA_CHECK(cudaStreamSynchronize(s->st));  // x lives on this stack frame
  }
  return THB_OK;
}

// DoInnerIterationsIfNeeded (ceres trust_region_minimizer.cc; external): refines the candidate Xc in place; the scalars
// SC_COST_INNER / SC_STEP2_INNER / SC_XNEW2_INNER and FL_EVAL_INNER are on the host when it returns.
template <int PD>
int DoInnerIterations(ThbBaSession* s) {
  cudaStream_t st = s->st;
  const int nc = s->nc, np = s->np, ng = s->ng;
  THB_CUDA_CHECK(cudaMemcpyAsync(s->d_bk_cam, s->Xc.cam, sizeof(double) * 6 * nc, cudaMemcpyDeviceToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(s->d_bk_pts, s->Xc.pts, sizeof(double) * 4 * np, cudaMemcpyDeviceToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(s->d_bk_intr, s->Xc.intr, sizeof(double) * KS * ng, cudaMemcpyDeviceToDevice, st));
  const InnerLmParams P;
  // group 0: camera extrinsics
  if (nc > 0) {
    k_inner_cam<PD><<<nc, IC_THREADS, 0, st>>>(s->K, s->Xc, s->Oc, s->d_cam_start, P, s->d_has_prior, s->d_prior);
    k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(s->Xc.cam, s->Xc.camd, nc, s->d_cs, s->d_cam_const, s->d_cam_group);
    s->sum.gpu_launches += 2;
  }
  // group 1: shared intrinsics blocks
  for (int sl = 0; sl < s->nvg; ++sl) { const int rc = InnerIntrSolve<PD>(s, sl); if (rc != THB_OK) return rc; }
  // group 2: points
  if (np > 0) {
    TrackBaParams tp;
    tp.max_num_iterations = P.max_num_iterations; tp.max_invalid = P.max_invalid; tp.jacobi_scaling = 1;
    tp.ftol = P.ftol; tp.gtol = P.gtol; tp.ptol = P.ptol; tp.radius0 = P.radius0; tp.min_radius = P.min_radius; tp.max_radius = P.max_radius;
    tp.min_relative_decrease = P.min_relative_decrease; tp.min_diag = P.min_diag; tp.max_diag = P.max_diag;
    tp.rays = nullptr; tp.status = nullptr; tp.bundle_adjustment = 1; tp.cos_min_angle = -2.0; tp.sq_max_reprojection_error = 0.0;
    BaState scratch = s->Xc;
    scratch.pts = s->d_pts_scratch;
    THB_CUDA_CHECK(cudaMemcpyAsync(s->d_pts_scratch, s->Xc.pts, sizeof(double) * 4 * np, cudaMemcpyDeviceToDevice, st));
    k_track_ba<PD><<<cdiv(np, 64), 64, 0, st>>>(s->K, s->Xc, scratch, s->Op, s->d_pt_start, s->d_inner_ps, tp, s->d_inner_res);
    ++s->sum.gpu_launches;
  }
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_COST_INNER, 0, sizeof(double) * 3, st));
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_flag + FL_EVAL_INNER, 0, sizeof(int), st));
  RunCost(s, s->Xc, SC_COST_INNER, FL_EVAL_INNER);
  ++s->sum.num_cost_evaluations;
  k_step_norms<<<cdiv((long long)nc + np + ng, 256), 256, 0, st>>>(nc, np, ng, s->d_cam_const, s->d_pt_const, s->d_intr_slot, s->d_intr_model, s->X, s->Xc,
                                                                  s->d_scal + SC_STEP2_INNER, s->d_scal + SC_XNEW2_INNER);
  ++s->sum.gpu_launches;
  ++s->num_inner_steps;
  return ReadScalars(s);
}

int RestoreCandidate(ThbBaSession* s) {  // "Inner iteration failed": the trust-region candidate stands
  THB_CUDA_CHECK(cudaMemcpyAsync(s->Xc.cam, s->d_bk_cam, sizeof(double) * 6 * s->nc, cudaMemcpyDeviceToDevice, s->st));
  THB_CUDA_CHECK(cudaMemcpyAsync(s->Xc.pts, s->d_bk_pts, sizeof(double) * 4 * s->np, cudaMemcpyDeviceToDevice, s->st));
  THB_CUDA_CHECK(cudaMemcpyAsync(s->Xc.intr, s->d_bk_intr, sizeof(double) * KS * s->ng, cudaMemcpyDeviceToDevice, s->st));
  k_cam_derive<<<cdiv(s->nc, 128), 128, 0, s->st>>>(s->Xc.cam, s->Xc.camd, s->nc, s->d_cs, s->d_cam_const, s->d_cam_group);
  ++s->sum.gpu_launches;
  return THB_OK;
}

void Terminate(ThbBaSession* s, int type) { s->sum.termination_type = type; s->finished = true; }

// One pass of TrustRegionMinimizer's main loop. Returns THB_OK; sets s->finished on termination.
int OneIteration(ThbBaSession* s) {
  const ThbBaOptions& O = s->opt;
  // FinalizeIterationAndCheckIfMinimizerCanContinue (tests that need no device data)
  const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - s->t_solve_start).count();
  if (elapsed >= O.max_solver_time_in_seconds) { Terminate(s, THB_TERM_NO_CONVERGENCE); return THB_OK; }
  if (s->iteration >= O.max_num_iterations) { Terminate(s, THB_TERM_NO_CONVERGENCE); return THB_OK; }
  if (s->radius <= O.min_trust_region_radius) { Terminate(s, THB_TERM_CONVERGENCE); return THB_OK; }
  // The gradient test of the previous (successful) iteration needs b = J^T r at the current point, which
  // K2 produces anyway: the whole iteration is enqueued and the test is applied when the scalars arrive.
  const bool want_grad = s->step_is_successful;
  int rc = BuildBlocks(s, want_grad);
  if (rc != THB_OK) return rc;
  rc = SolveAndStep(s);
  if (rc != THB_OK) return rc;
  rc = ReadScalars(s);
  if (rc != THB_OK) return rc;
  if (want_grad) {
    s->gradient_max_norm = s->h_scal[SC_GRADMAX];
    if (O.gradient_tolerance >= 0.0 && s->gradient_max_norm <= O.gradient_tolerance) { Terminate(s, THB_TERM_CONVERGENCE); return THB_OK; }
  }
  ++s->iteration;
  s->step_is_successful = false;
  double model_cost_change = s->h_scal[SC_MCC];
  bool step_valid = !(s->h_flag[FL_CHOL] || s->h_flag[FL_POINT]) && std::isfinite(model_cost_change) && model_cost_change > 0.0;
  if (!step_valid) {
    // HandleInvalidStep / LevenbergMarquardtStrategy::StepIsInvalid
    if (++s->num_consecutive_invalid >= O.max_num_consecutive_invalid_steps) { Terminate(s, THB_TERM_FAILURE); return THB_OK; }
    s->radius /= s->decrease_factor; s->decrease_factor *= 2.0;
    LogIter(s);
    return THB_OK;
  }
  s->num_consecutive_invalid = 0;
  if (s->constrained) {
    // TrustRegionMinimizer::DoLineSearch (projected Armijo search along delta, first probe at step size 1 = the candidate
    // just evaluated). A failed probe halves the step until the sufficient-decrease test holds (at most 20 probes), as
    // the oracle does; Ceres contracts by polynomial interpolation instead (deviations).
    const double gtd = s->h_scal[SC_GTD];
    ++s->sum.num_cost_evaluations;
    auto probe_ok = [&](double alpha) {
      const double c = s->h_scal[SC_COST_CAND];
      return !s->h_flag[FL_EVAL_CAND] && std::isfinite(c) && c <= s->x_cost + 1e-4 * gtd * alpha;
    };
    if (!probe_ok(1.0)) {
      double alpha = 1.0;
      bool found = false;
      for (int it = 1; it < 20 && !found; ++it) {
        alpha *= 0.5;
        if ((rc = ComputeCandidate(s, alpha)) != THB_OK) return rc;
        if ((rc = ReadScalars(s)) != THB_OK) return rc;
        found = probe_ok(alpha);
      }
      if (!found) {
        if ((rc = ComputeCandidate(s, 1.0)) != THB_OK) return rc;
        if ((rc = ReadScalars(s)) != THB_OK) return rc;
      }
    }
  }
  double cand_cost = s->h_flag[FL_EVAL_CAND] ? std::numeric_limits<double>::max() : s->h_scal[SC_COST_CAND];
  double step2 = s->h_scal[SC_STEP2], xnew2 = s->h_scal[SC_XNEW2];
  // DoInnerIterationsIfNeeded
  bool inner_useful = false;
  if (s->inner_enabled && cand_cost < std::numeric_limits<double>::max()) {
    rc = s->PD == 3 ? DoInnerIterations<3>(s) : DoInnerIterations<4>(s);
    if (rc != THB_OK) return rc;
    if (s->h_flag[FL_EVAL_INNER]) {
      if ((rc = RestoreCandidate(s)) != THB_OK) return rc;
    } else {
      const double inner_cost = s->h_scal[SC_COST_INNER];
      model_cost_change += cand_cost - inner_cost;
      inner_useful = inner_cost < s->x_cost;
      s->inner_enabled = 1.0 - inner_cost / cand_cost > 1e-3;  // inner_iteration_tolerance
      cand_cost = inner_cost;
      step2 = s->h_scal[SC_STEP2_INNER]; xnew2 = s->h_scal[SC_XNEW2_INNER];
    }
  }
  // ParameterToleranceReached
  const double step_norm = std::sqrt(step2);
  if (O.parameter_tolerance >= 0.0 && step_norm <= O.parameter_tolerance * (s->x_norm + O.parameter_tolerance)) { Terminate(s, THB_TERM_CONVERGENCE); return THB_OK; }
  // FunctionToleranceReached
  const double cost_change = s->x_cost - cand_cost;
  if (O.function_tolerance >= 0.0 && std::fabs(cost_change) <= O.function_tolerance * s->x_cost) { Terminate(s, THB_TERM_CONVERGENCE); return THB_OK; }
  // IsStepSuccessful (monotonic)
  const double relative_decrease = cand_cost >= std::numeric_limits<double>::max() ? std::numeric_limits<double>::lowest()
                                                                                   : cost_change / model_cost_change;
  if (inner_useful || relative_decrease > O.min_relative_decrease) {
    std::swap(s->X, s->Xc);
    s->x_norm = std::sqrt(xnew2);
    rc = EvaluateJacobian(s);  // EvaluateGradientAndJacobian(new_evaluation_point = false)
    if (rc != THB_OK) return rc;
    THB_CUDA_CHECK(cudaMemcpyAsync(s->h_scal, s->d_scal, sizeof(double), cudaMemcpyDeviceToHost, s->st));
    THB_CUDA_CHECK(cudaMemcpyAsync(s->h_flag, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->st));
    THB_CUDA_CHECK(cudaStreamSynchronize(s->st));
    if (s->h_flag[FL_EVAL_X]) { Terminate(s, THB_TERM_FAILURE); return THB_OK; }
    s->x_cost = s->h_scal[SC_COST_X];
    s->step_is_successful = true;
    ++s->sum.num_successful_steps;
    s->radius = s->radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
    s->radius = std::min(O.max_trust_region_radius, s->radius);
    s->decrease_factor = 2.0;
    s->min_cost = std::min(s->min_cost, s->x_cost);
  } else {
    s->radius /= s->decrease_factor; s->decrease_factor *= 2.0;
  }
  LogIter(s);
  return THB_OK;
}

int ValidateAndCreate(const ThbBaProblem* P, const ThbBaOptions* O, void* stream, ThbBaSession** out) {
  if (!P || !O || !out) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  *out = nullptr;
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  if (P->num_cameras < 0 || P->num_groups < 0 || P->num_points < 0 || P->num_observations < 0)
    THB_FAIL(THB_E_INVALID_ARGUMENT, "negative size");
  if (P->memory_space != THB_MEM_HOST && P->memory_space != THB_MEM_DEVICE) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad memory_space");
  if (P->num_observations > 0 && (!P->cam_ext || !P->cam_group || !P->intr || !P->intr_model || !P->pts || !P->obs_cam || !P->obs_pt || !P->obs_xy))
    THB_FAIL(THB_E_INVALID_ARGUMENT, "null array");
  if (O->linear_solver != THB_SOLVER_SCHUR_CHOLESKY && O->linear_solver != THB_SOLVER_SCHUR_PCG) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad linear_solver");
  if (O->linear_solver == THB_SOLVER_SCHUR_PCG && (!(O->pcg_eta > 0.0) || O->pcg_max_iterations < 1))
    THB_FAIL(THB_E_INVALID_ARGUMENT, "pcg_eta must be positive and pcg_max_iterations at least 1");
  if (O->loss_function_type < THB_LOSS_TRIVIAL || O->loss_function_type > THB_LOSS_TRUNCATED) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad loss type");

  ThbBaSession* s = new ThbBaSession();
  s->t_create = std::chrono::steady_clock::now();
  s->prob = *P; s->opt = *O; s->st = (cudaStream_t)stream;
  s->nc = P->num_cameras; s->ng = P->num_groups; s->np = P->num_points; s->no = P->num_observations;
  s->PD = O->use_homogeneous_point_parametrization ? 3 : 4;
  const int nc = s->nc, ng = s->ng, np = s->np, no = s->no, sp = P->memory_space;
  s->t_jac.st = s->t_normal.st = s->t_solve.st = s->t_update.st = s->st;

#define THB_TRY(expr) do { rc = (expr); if (rc != THB_OK) { FreeSession(s); return rc; } } while (0)
#define THB_TRY_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { SetLastError(std::string(#expr) + ": " + cudaGetErrorString(_e)); FreeSession(s); return THB_E_CUDA; } } while (0)

  // THB_SETUP_PROF=1: host-side time stamps of the setup phases (debugging aid, stderr)
  const bool setup_prof = getenv("THB_SETUP_PROF") != nullptr;
  std::vector<std::pair<const char*, double>> marks;
  auto mark = [&](const char* name) {
    if (setup_prof) marks.emplace_back(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - s->t_create).count());
  };
  // ---- device memory (stream-ordered, pooled) ----
  ConfigurePoolOnce();
  s->arena.st = s->st;
  Arena& M = s->arena;
  cudaStream_t st = s->st;
  s->h_block = g_pinned.Get();
  if (!s->h_block) { FreeSession(s); THB_FAIL(THB_E_CUDA, "cudaMallocHost failed"); }
  s->h_scal = reinterpret_cast<double*>(s->h_block);
  s->h_flag = reinterpret_cast<int*>(reinterpret_cast<char*>(s->h_block) + 128);
  int* h_setup = reinterpret_cast<int*>(reinterpret_cast<char*>(s->h_block) + 256);  // SF_COUNT ints
  THB_TRY(M.Get(&s->X.cam, (size_t)nc * 6)); THB_TRY(M.Get(&s->X.camd, (size_t)nc * CAMD));
  THB_TRY(M.Get(&s->X.intr, (size_t)ng * KS)); THB_TRY(M.Get(&s->X.pts, (size_t)np * 4));
  THB_TRY(M.Get(&s->Xc.cam, (size_t)nc * 6)); THB_TRY(M.Get(&s->Xc.camd, (size_t)nc * CAMD));
  THB_TRY(M.Get(&s->Xc.intr, (size_t)ng * KS)); THB_TRY(M.Get(&s->Xc.pts, (size_t)np * 4));
  THB_TRY(M.Get(&s->d_cam_group, nc)); THB_TRY(M.Get(&s->d_intr_model, ng)); THB_TRY(M.Get(&s->d_intr_slot, ng));
  THB_TRY(M.Get(&s->d_cam_const, nc)); THB_TRY(M.Get(&s->d_pt_const, np)); THB_TRY(M.Get(&s->d_intr_const, ng));
  THB_TRY(M.Get(&s->d_op_cam, no)); THB_TRY(M.Get(&s->d_op_pt, no)); THB_TRY(M.Get(&s->d_oc_cam, no)); THB_TRY(M.Get(&s->d_oc_pt, no));
  THB_TRY(M.Get(&s->d_op_xy, no)); THB_TRY(M.Get(&s->d_op_si, no)); THB_TRY(M.Get(&s->d_oc_xy, no)); THB_TRY(M.Get(&s->d_oc_si, no));
  THB_TRY(M.Get(&s->d_pt_start, np + 1)); THB_TRY(M.Get(&s->d_cam_start, nc + 1));
  THB_TRY(M.Get(&s->d_r, (size_t)no * 2)); THB_TRY(M.Get(&s->d_jc, (size_t)no * 12)); THB_TRY(M.Get(&s->d_jp, (size_t)no * 2 * s->PD));
  THB_TRY(M.Get(&s->d_ps, (size_t)np * s->PD));
  THB_TRY(M.Get(&s->d_vinv, (size_t)np * s->PD * s->PD)); THB_TRY(M.Get(&s->d_gp, (size_t)np * s->PD)); THB_TRY(M.Get(&s->d_pdiag, (size_t)np * s->PD));
  THB_TRY(M.Get(&s->d_yp, (size_t)np * s->PD));
  THB_TRY(M.Get(&s->d_jy, (size_t)no * 2)); THB_TRY(M.Get(&s->d_bsum, (size_t)np * s->PD));
  THB_TRY(M.Get(&s->d_scal, SC_COUNT)); THB_TRY(M.Get(&s->d_flag, FL_COUNT));
  THB_TRY(M.Get(&s->d_op_slot, no));
  int *d_used = nullptr, *d_setup = nullptr, *d_perm = nullptr;
  THB_TRY(M.Get(&d_used, ng)); THB_TRY(M.Get(&d_setup, SF_COUNT)); THB_TRY(M.Get(&d_perm, no));

  // ---- the caller's arrays on the device (memory_space == DEVICE: used in place) ----
  const cudaMemcpyKind kin = sp == THB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const int *raw_cam = P->obs_cam, *raw_pt = P->obs_pt;
  const double2* raw_xy = reinterpret_cast<const double2*>(P->obs_xy);
  const double2* raw_si = reinterpret_cast<const double2*>(P->obs_sqrt_info);
  if (sp == THB_MEM_HOST && no > 0) {
    int *dc = nullptr, *dp = nullptr; double2 *dxy = nullptr, *dsi = nullptr;
    THB_TRY(M.Get(&dc, no)); THB_TRY(M.Get(&dp, no)); THB_TRY(M.Get(&dxy, no));
    THB_TRY_CUDA(cudaMemcpyAsync(dc, P->obs_cam, sizeof(int) * no, cudaMemcpyHostToDevice, st));
    THB_TRY_CUDA(cudaMemcpyAsync(dp, P->obs_pt, sizeof(int) * no, cudaMemcpyHostToDevice, st));
    THB_TRY_CUDA(cudaMemcpyAsync(dxy, P->obs_xy, sizeof(double2) * no, cudaMemcpyHostToDevice, st));
    if (P->obs_sqrt_info) {
      THB_TRY(M.Get(&dsi, no));
      THB_TRY_CUDA(cudaMemcpyAsync(dsi, P->obs_sqrt_info, sizeof(double2) * no, cudaMemcpyHostToDevice, st));
    }
    raw_cam = dc; raw_pt = dp; raw_xy = dxy; raw_si = dsi;
  }
  THB_TRY_CUDA(cudaMemcpyAsync(s->X.cam, P->cam_ext, sizeof(double) * nc * 6, kin, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->X.intr, P->intr, sizeof(double) * ng * KS, kin, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->X.pts, P->pts, sizeof(double) * np * 4, kin, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->Xc.intr, s->X.intr, sizeof(double) * ng * KS, cudaMemcpyDeviceToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_cam_group, P->cam_group, sizeof(int) * nc, kin, st));
  if (P->cam_const) THB_TRY_CUDA(cudaMemcpyAsync(s->d_cam_const, P->cam_const, nc, kin, st));
  else THB_TRY_CUDA(cudaMemsetAsync(s->d_cam_const, 0, std::max(nc, 1), st));
  if ((P->cam_has_position_prior || P->cam_has_gravity_prior || P->cam_has_orientation_prior) && nc > 0) {  // camera priors, packed per camera (see d_prior)
    if ((P->cam_has_position_prior && (!P->cam_position_prior || !P->cam_position_prior_sqrt_info)) ||
        (P->cam_has_gravity_prior && (!P->cam_gravity_prior || !P->cam_gravity_prior_sqrt_info)) ||
        (P->cam_has_orientation_prior && (!P->cam_orientation_prior || !P->cam_orientation_prior_sqrt_info))) {
      FreeSession(s); THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_has_*_prior without prior / sqrt information arrays");
    }
    uint8_t* d_h[3] = {nullptr, nullptr, nullptr};
    THB_TRY(M.Get(&s->d_has_prior, nc)); THB_TRY(M.Get(&s->d_prior, (size_t)nc * PRIOR_STRIDE));
    THB_TRY_CUDA(cudaMemsetAsync(s->d_prior, 0, sizeof(double) * PRIOR_STRIDE * nc, st));
    const uint8_t* has[3] = {P->cam_has_position_prior, P->cam_has_gravity_prior, P->cam_has_orientation_prior};
    const double* info[3] = {P->cam_position_prior_sqrt_info, P->cam_gravity_prior_sqrt_info, P->cam_orientation_prior_sqrt_info};
    const double* prior[3] = {P->cam_position_prior, P->cam_gravity_prior, P->cam_orientation_prior};
    for (int kind = 0; kind < PRIOR_KINDS; ++kind) {
      THB_TRY(M.Get(&d_h[kind], nc));
      THB_TRY_CUDA(cudaMemsetAsync(d_h[kind], 0, nc, st));
      if (!has[kind]) continue;
      THB_TRY_CUDA(cudaMemcpyAsync(d_h[kind], has[kind], nc, kin, st));
      THB_TRY_CUDA(cudaMemcpy2DAsync(s->d_prior + 12 * kind, PRIOR_STRIDE * sizeof(double), info[kind], 9 * sizeof(double), 9 * sizeof(double), nc, kin, st));
      THB_TRY_CUDA(cudaMemcpy2DAsync(s->d_prior + 12 * kind + 9, PRIOR_STRIDE * sizeof(double), prior[kind], 3 * sizeof(double), 3 * sizeof(double), nc, kin, st));
    }
    k_prior_flags<<<cdiv(nc, 256), 256, 0, st>>>(nc, d_h[0], d_h[1], d_h[2], s->d_has_prior);
  }
  if (P->pt_const) THB_TRY_CUDA(cudaMemcpyAsync(s->d_pt_const, P->pt_const, np, kin, st));
  else THB_TRY_CUDA(cudaMemsetAsync(s->d_pt_const, 0, std::max(np, 1), st));

  mark("alloc + copies enqueued");
  // ---- structure: validation, constness, the two observation orders (histogram + scan + stable radix sort) ----
  THB_TRY_CUDA(cudaMemsetAsync(s->d_pt_start, 0, sizeof(int) * (np + 1), st));
  THB_TRY_CUDA(cudaMemsetAsync(s->d_cam_start, 0, sizeof(int) * (nc + 1), st));
  THB_TRY_CUDA(cudaMemsetAsync(d_used, 0, sizeof(int) * std::max(ng, 1), st));
  THB_TRY_CUDA(cudaMemsetAsync(d_setup, 0, sizeof(int) * SF_COUNT, st));
  if (nc > 0) k_setup_check_groups<<<cdiv(nc, 256), 256, 0, st>>>(nc, ng, s->d_cam_group, d_setup);
  if (no > 0) k_setup_count<<<cdiv(no, 256), 256, 0, st>>>(no, nc, np, ng, raw_cam, raw_pt, s->d_cam_group, s->d_pt_start, s->d_cam_start, d_used, d_setup);
  // indices must be known good before they key a sort
  THB_TRY_CUDA(cudaMemcpyAsync(h_setup, d_setup, sizeof(int) * SF_COUNT, cudaMemcpyDeviceToHost, st));
  std::vector<int> h_intr_model;
  std::vector<uint16_t> h_intr_const(ng, 0xffff);
  THB_TRY(FetchToHost(P->intr_model, ng, sp, &h_intr_model));
  if (P->intr_const) THB_TRY(FetchToHost(P->intr_const, ng, sp, &h_intr_const));
  for (int g = 0; g < ng; ++g)
    if (num_intrinsics(h_intr_model[g]) < 0) { FreeSession(s); THB_FAIL(THB_E_UNSUPPORTED, "camera model not on the hot path"); }
  THB_TRY_CUDA(cudaStreamSynchronize(st));
  mark("sync 1 (counts, validation)");
  if (h_setup[SF_BAD_GROUP]) { FreeSession(s); THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_group out of range"); }
  if (h_setup[SF_BAD_INDEX]) { FreeSession(s); THB_FAIL(THB_E_INVALID_ARGUMENT, "observation index out of range"); }
  THB_TRY(GroupByKey(raw_pt, no, np, s->d_pt_start, d_perm, st));
  if (no > 0) k_setup_gather<<<cdiv(no, 256), 256, 0, st>>>(no, d_perm, raw_cam, raw_pt, raw_xy, raw_si, s->d_op_cam, s->d_op_pt, s->d_op_xy, s->d_op_si);
  THB_TRY(GroupByKey(raw_cam, no, nc, s->d_cam_start, d_perm, st));
  if (no > 0) k_setup_gather<<<cdiv(no, 256), 256, 0, st>>>(no, d_perm, raw_cam, raw_pt, raw_xy, raw_si, s->d_oc_cam, s->d_oc_pt, s->d_oc_xy, s->d_oc_si);
  if (nc + np > 0) k_setup_const<<<cdiv((long long)nc + np, 256), 256, 0, st>>>(nc, np, s->d_cam_start, s->d_pt_start, s->d_cam_const, s->d_pt_const, d_setup);
  std::vector<int> pt_start(np + 1, 0), used(ng, 0);
  THB_TRY_CUDA(cudaMemcpyAsync(pt_start.data(), s->d_pt_start, sizeof(int) * (np + 1), cudaMemcpyDeviceToHost, st));
  THB_TRY_CUDA(cudaMemcpyAsync(used.data(), d_used, sizeof(int) * ng, cudaMemcpyDeviceToHost, st));
  THB_TRY_CUDA(cudaMemcpyAsync(h_setup, d_setup, sizeof(int) * SF_COUNT, cudaMemcpyDeviceToHost, st));
  THB_TRY_CUDA(cudaStreamSynchronize(st));
  mark("sync 2 (sorts, gathers, pt_start D2H)");
  s->model = ng > 0 ? h_intr_model[0] : THB_MODEL_PINHOLE;
  for (int g = 1; g < ng; ++g) if (h_intr_model[g] != s->model) s->model = -1;
  // variable intrinsics groups: observed by at least one residual and with at least one free coordinate
  std::vector<int> slot(ng, -1), slot_group;
  for (int g = 0; g < ng; ++g) {
    const unsigned all = (1u << num_intrinsics(h_intr_model[g])) - 1;
    if (used[g] && (h_intr_const[g] & all) != all) { slot[g] = (int)slot_group.size(); slot_group.push_back(g); }
  }
  s->nvg = (int)slot_group.size();
  if (s->nvg > MAX_VG) {
    FreeSession(s);
    THB_FAIL(THB_E_UNSUPPORTED, "more than 8 intrinsics groups with free parameters are not supported (shared-intrinsics design)");
  }
  s->constrained = s->nvg > 0;  // focal length >= 1 is always set on a non-constant block (bundle_adjuster.cc:396-405)
  s->red_variable = s->nvg > 0 || h_setup[SF_RED_VARIABLE] != 0;
  s->pt_variable = h_setup[SF_PT_VARIABLE] != 0;
  s->any_variable = s->red_variable || s->pt_variable;
  // Schur chunks: consecutive points with <= 64 observations in total
  std::vector<int> chunk_pt;
  chunk_pt.push_back(0);
  for (int p = 0, acc = 0; p < np; ++p) {
    const int n = pt_start[p + 1] - pt_start[p];
    if (acc > 0 && acc + n > 64) { chunk_pt.push_back(p); acc = 0; }
    acc += n;
  }
  chunk_pt.push_back(np);
  s->nchunks = (int)chunk_pt.size() - 1;
  s->n_red = 6 * nc + NI * s->nvg;
  THB_TRY(M.Get(&s->d_chunk_pt, chunk_pt.size()));
  THB_TRY(M.Get(&s->d_cs, (size_t)s->n_red)); THB_TRY(M.Get(&s->d_braw, (size_t)s->n_red)); THB_TRY(M.Get(&s->d_cdiag, (size_t)s->n_red));
  THB_TRY(M.Get(&s->d_slot_group, s->nvg));
  THB_TRY(M.Get(&s->d_ilo, (size_t)NI * s->nvg)); THB_TRY(M.Get(&s->d_ihi, (size_t)NI * s->nvg));
  if (s->nvg > 0) {
    THB_TRY(M.Get(&s->d_ji, (size_t)no * 2 * NI));
    THB_TRY(M.Get(&s->d_zt, (size_t)np * s->nvg * 2 * NI * s->PD));
  }
  mark("chunks on host");
  THB_TRY(s->chol.Init(std::max(1, s->n_red), st));
  mark("chol.Init");
  s->use_pcg = O->linear_solver == THB_SOLVER_SCHUR_PCG;
  if (s->use_pcg) {  // SCHUR_JACOBI: one diagonal block per camera and per shared intrinsics block
    std::vector<int2> blocks;
    for (int c = 0; c < nc; ++c) blocks.push_back(make_int2(6 * c, 6));
    for (int sl = 0; sl < s->nvg; ++sl) blocks.push_back(make_int2(6 * nc + NI * sl, NI));
    static_assert(NI <= kPcgMaxBlockDim, "preconditioner block size");
    THB_TRY(s->pcg.Init(s->chol.n_pad, blocks, st));
  }
  // the preprocessor disables inner iterations on programs with fewer than two parameter blocks (ceres)
  s->inner_enabled = O->use_inner_iterations != 0 && h_setup[SF_NUM_CAM_VAR] + h_setup[SF_NUM_PT_VAR] + s->nvg >= 2;
  if (s->inner_enabled) {
    THB_TRY(M.Get(&s->d_bk_cam, (size_t)nc * 6)); THB_TRY(M.Get(&s->d_bk_pts, (size_t)np * 4)); THB_TRY(M.Get(&s->d_bk_intr, (size_t)ng * KS));
    THB_TRY(M.Get(&s->d_inner_ps, (size_t)np * s->PD)); THB_TRY(M.Get(&s->d_pts_scratch, (size_t)np * 4));
    THB_TRY(M.Get(&s->d_inner_out, II_VALS)); THB_TRY(M.Get(&s->d_inner_scale, NI)); THB_TRY(M.Get(&s->d_inner_res, np));
    s->h_slot_group = slot_group; s->h_model = h_intr_model; s->h_const = h_intr_const;
  }
  // box constraints of bundle_adjuster.cc:396-427
  std::vector<double> ilo((size_t)NI * s->nvg, -std::numeric_limits<double>::max()), ihi((size_t)NI * s->nvg, std::numeric_limits<double>::max());
  for (int sl = 0; sl < s->nvg; ++sl) {
    const int m = h_intr_model[slot_group[sl]];
    ilo[NI * sl + 0] = 1.0;  // focal length is parameter 0 of every model
    if (m == THB_MODEL_DOUBLE_SPHERE) { ilo[NI * sl + 5] = -1.0; ihi[NI * sl + 5] = 1.0; ilo[NI * sl + 6] = 0.0; ihi[NI * sl + 6] = 1.0; }
    else if (m == THB_MODEL_EXTENDED_UNIFIED) { ilo[NI * sl + 5] = 0.0; ihi[NI * sl + 5] = 1.0; ilo[NI * sl + 6] = 0.1; }
  }
  // pageable sources: these copies complete before the call returns
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_intr_model, h_intr_model.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_intr_const, h_intr_const.data(), sizeof(uint16_t) * ng, cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_intr_slot, slot.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_chunk_pt, chunk_pt.data(), sizeof(int) * chunk_pt.size(), cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_slot_group, slot_group.data(), sizeof(int) * s->nvg, cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_ilo, ilo.data(), sizeof(double) * ilo.size(), cudaMemcpyHostToDevice, st));
  THB_TRY_CUDA(cudaMemcpyAsync(s->d_ihi, ihi.data(), sizeof(double) * ihi.size(), cudaMemcpyHostToDevice, st));
  s->h_ilo = ilo; s->h_ihi = ihi;
  if (no > 0) k_setup_slots<<<cdiv(no, 256), 256, 0, st>>>(no, s->d_op_cam, s->d_op_pt, s->d_cam_group, s->d_intr_slot, s->d_cam_const, s->d_pt_const, s->d_op_slot, d_setup);
  // Ceres drops residual blocks whose parameter blocks are all constant (their cost is Summary::fixed_cost);
  // they would still contribute zero Jacobian columns here, so only the cost bookkeeping differs.
  THB_TRY_CUDA(cudaMemcpyAsync(h_setup, d_setup, sizeof(int) * SF_COUNT, cudaMemcpyDeviceToHost, st));
  THB_TRY_CUDA(cudaStreamSynchronize(st));  // also: the host vectors above may now go out of scope
  mark("sync 3 (slots)");
  if (h_setup[SF_HAS_FIXED] && s->any_variable) { FreeSession(s); THB_FAIL(THB_E_UNSUPPORTED, "observations whose camera, intrinsics and point are all constant are not supported in a problem with free blocks"); }
  s->Op = ObsSoA{s->d_op_cam, s->d_op_pt, s->d_op_xy, s->d_op_si};
  s->Oc = ObsSoA{s->d_oc_cam, s->d_oc_pt, s->d_oc_xy, s->d_oc_si};
  s->K = BaConst{nc, ng, np, no, s->d_cam_group, s->d_intr_model, s->d_cam_const, s->d_intr_const, s->d_pt_const, s->d_intr_slot,
                 O->loss_function_type, O->robust_loss_width};

  // ---- IterationZero ----
  std::memset(&s->sum, 0, sizeof(s->sum));
  s->sum.termination_type = THB_TERM_NO_CONVERGENCE;
  s->radius = O->initial_trust_region_radius; s->decrease_factor = 2.0;
  if (!s->any_variable || no == 0) {
    // nothing to optimise: report the cost of the (fixed) residual blocks
    THB_TRY_CUDA(cudaMemsetAsync(s->d_scal, 0, sizeof(double) * SC_COUNT, st));
    THB_TRY_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int) * FL_COUNT, st));
    if (no > 0) {
      k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(s->X.cam, s->X.camd, nc, s->d_cs, s->d_cam_const, s->d_cam_group);
      RunCost(s, s->X, SC_COST_X, FL_EVAL_X);
    }
    THB_TRY(ReadScalars(s));
    s->fixed_cost = s->h_scal[SC_COST_X]; s->x_cost = 0.0; s->min_cost = 0.0;
    s->sum.initial_cost = s->fixed_cost;
    Terminate(s, THB_TERM_CONVERGENCE);
    s->t_solve_start = std::chrono::steady_clock::now();
    *out = s;
    return THB_OK;
  }
  THB_TRY_CUDA(cudaMemsetAsync(s->d_scal, 0, sizeof(double) * SC_COUNT, st));
  THB_TRY_CUDA(cudaMemsetAsync(s->d_flag, 0, sizeof(int) * FL_COUNT, st));
  k_fill<<<cdiv(s->n_red, 256), 256, 0, st>>>(s->n_red, s->d_cs, 1.0);
  k_fill<<<cdiv((long long)np * s->PD, 256), 256, 0, st>>>(np * s->PD, s->d_ps, 1.0);
  k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(s->X.cam, s->X.camd, nc, s->d_cs, s->d_cam_const, s->d_cam_group);
  if (s->constrained) {  // IterationZero: x = Plus(x, 0) makes the start feasible
    k_update_intr<<<cdiv(ng, 128), 128, 0, st>>>(ng, nc, s->d_intr_slot, s->d_intr_model, s->X.intr, nullptr, nullptr, nullptr, s->d_ilo, s->d_ihi, 1.0,
                                                s->X.intr, nullptr);
    ++s->sum.gpu_launches;
  }
  k_xnorm<<<cdiv((long long)nc + np + ng, 256), 256, 0, st>>>(nc, np, ng, s->d_cam_const, s->d_pt_const, s->d_intr_slot, s->d_intr_model, s->X.cam,
                                                             s->X.pts, s->X.intr, s->d_scal + SC_XNEW2);
  s->sum.gpu_launches += 4;
  if (O->jacobi_scaling) {
    // column norms of the unscaled Jacobian -> scale = 1/(1+sqrt(norm^2)), fixed for the whole solve
    RunJacobian(s, s->d_cs, s->d_ps);
    THB_TRY(BuildBlocks(s, false));
    k_make_scale<<<cdiv(s->n_red, 256), 256, 0, st>>>(s->n_red, s->d_cdiag, s->d_cs);
    k_make_scale<<<cdiv((long long)np * s->PD, 256), 256, 0, st>>>(np * s->PD, s->d_pdiag, s->d_ps);
    k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(s->X.cam, s->X.camd, nc, s->d_cs, s->d_cam_const, s->d_cam_group);  // scale is part of the record
    s->sum.gpu_launches += 3;
    THB_TRY_CUDA(cudaMemsetAsync(s->d_scal + SC_COST_X, 0, sizeof(double), st));
  }
  THB_TRY(EvaluateJacobian(s));
  THB_TRY(ReadScalars(s));
  if (s->h_flag[FL_EVAL_X]) { FreeSession(s); THB_FAIL(THB_E_NUMERICAL, "initial residual and Jacobian evaluation failed"); }
  s->x_cost = s->h_scal[SC_COST_X];
  s->x_norm = std::sqrt(s->h_scal[SC_XNEW2]);
  s->min_cost = s->x_cost;
  s->sum.initial_cost = s->x_cost + s->fixed_cost;
  s->step_is_successful = true; s->iteration = 0;
  LogIter(s);
  mark("iteration zero");
  if (setup_prof) {
    fprintf(stderr, "THB_SETUP_PROF");
    for (auto& m : marks) fprintf(stderr, " | %s %.3f", m.first, m.second);
    fprintf(stderr, "\n");
  }
  s->t_solve_start = std::chrono::steady_clock::now();
  s->sum.setup_time_in_seconds = std::chrono::duration<double>(s->t_solve_start - s->t_create).count();
  *out = s;
  return THB_OK;
#undef THB_TRY
#undef THB_TRY_CUDA
}

}  // namespace

extern "C" {

int thb_ba_covariance(const ThbBaProblem* problem, const ThbBaOptions* options, double* cam_cov, uint8_t* cam_ok, double* pt_cov, uint8_t* pt_ok,
                      void* cuda_stream) {
  if (!problem || !options) THB_FAIL(THB_E_INVALID_ARGUMENT, "null argument");
  ThbBaOptions o = *options;
  o.jacobi_scaling = 0;          // the blocks are read in the problem's own units
  o.use_inner_iterations = 0;
  o.linear_solver = THB_SOLVER_SCHUR_CHOLESKY;
  ThbBaSession* s = nullptr;
  int rc = ValidateAndCreate(problem, &o, cuda_stream, &s);
  if (rc != THB_OK) return rc;
  const int nc = s->nc, np = s->np;
  const bool host = s->prob.memory_space == THB_MEM_HOST;
  const cudaMemcpyKind kout = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  auto fail = [&](int code, const char* msg) { FreeSession(s); SetLastError(msg); return code; };
  if (s->nvg > 0 || (s->red_variable && s->pt_variable))
    return fail(THB_E_UNSUPPORTED, "covariance needs a problem with only cameras or only points free (BundleAdjustView(s) / BundleAdjustTrack(s))");
  if (s->pt_variable && s->PD != 3) return fail(THB_E_INVALID_ARGUMENT, "point covariance needs use_homogeneous_point_parametrization (bundle_adjustment.cc:296)");
  if (s->red_variable && (!cam_cov || !cam_ok)) return fail(THB_E_INVALID_ARGUMENT, "cam_cov / cam_ok are NULL");
  if (s->pt_variable && (!pt_cov || !pt_ok)) return fail(THB_E_INVALID_ARGUMENT, "pt_cov / pt_ok are NULL");
  double* d_cov = nullptr; uint8_t* d_ok = nullptr;
  cudaError_t e = cudaSuccess;
  if (s->any_variable && s->no > 0) {
    s->radius = 1e30;  // no Levenberg-Marquardt damping: D^2 = diag / radius is far below the rounding of the blocks
    rc = BuildBlocks(s, false);
    if (rc != THB_OK) { FreeSession(s); return rc; }
  }
  const size_t n_items = s->red_variable ? (size_t)nc : (size_t)np, width = s->red_variable ? 36 : 9;
  if (s->any_variable && s->no > 0 && n_items > 0) {
    e = cudaMallocAsync(&d_cov, sizeof(double) * n_items * width, s->st);
    if (e == cudaSuccess) e = cudaMallocAsync(&d_ok, n_items, s->st);
    if (e == cudaSuccess) {
      if (s->red_variable) k_cov_cam<<<cdiv(nc, 64), 64, 0, s->st>>>(nc, s->d_cam_const, s->d_cam_start, s->chol.A, s->chol.ld, d_cov, d_ok);
      else k_cov_pt<<<cdiv(np, 128), 128, 0, s->st>>>(np, s->d_pt_const, s->d_pt_start, s->d_vinv, d_cov, d_ok);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->red_variable ? cam_cov : pt_cov, d_cov, sizeof(double) * n_items * width, kout, s->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->red_variable ? cam_ok : pt_ok, d_ok, n_items, kout, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (d_cov) cudaFreeAsync(d_cov, s->st);
    if (d_ok) cudaFreeAsync(d_ok, s->st);
  } else {  // nothing free: no block has a covariance
    if (cam_ok && nc > 0) e = host ? (std::memset(cam_ok, 0, nc), cudaSuccess) : cudaMemsetAsync(cam_ok, 0, nc, s->st);
    if (e == cudaSuccess && pt_ok && np > 0) e = host ? (std::memset(pt_ok, 0, np), cudaSuccess) : cudaMemsetAsync(pt_ok, 0, np, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
  }
  FreeSession(s);
  if (e != cudaSuccess) THB_FAIL(THB_E_CUDA, cudaGetErrorString(e));
  return THB_OK;
}

int thb_version(void) { return 100; }
const char* thb_last_error(void) { return thb::GetLastError(); }
int thb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

void thb_ba_default_options(ThbBaOptions* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  // BundleAdjustmentOptions defaults (bundle_adjustment.h:87-167), except use_inner_iterations: off here (the setting of
  // every benchmark and of BundleAdjustView / Track); the pt.sfm adapter passes the reference default (on) through
  o->loss_function_type = THB_LOSS_TRIVIAL; o->robust_loss_width = 2.0;
  o->linear_solver = THB_SOLVER_SCHUR_CHOLESKY;
  o->use_homogeneous_point_parametrization = 1; o->use_inner_iterations = 0;
  o->max_num_iterations = 100; o->jacobi_scaling = 1; o->verbose = 0;
  o->max_num_consecutive_invalid_steps = 5;
  o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->max_trust_region_radius = 1e12; o->initial_trust_region_radius = 1e4;
  o->min_trust_region_radius = 1e-32; o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->max_solver_time_in_seconds = 3600.0;
  o->pcg_eta = 0.1; o->pcg_max_iterations = 500;
}

int thb_ba_create(const ThbBaProblem* problem, const ThbBaOptions* options, void* cuda_stream, ThbBaSession** session) {
  return ValidateAndCreate(problem, options, cuda_stream, session);
}

int thb_ba_iterate(ThbBaSession* s, int32_t n, int32_t* ran) {
  if (!s) THB_FAIL(THB_E_INVALID_ARGUMENT, "null session");
  int done = 0;
  while (done < n && !s->finished) {
    const int before = s->iteration;
    const int rc = OneIteration(s);
    if (rc != THB_OK) return rc;
    if (s->iteration > before) ++done;
  }
  if (ran) *ran = done;
  return THB_OK;
}

int thb_ba_finish(ThbBaSession* s, ThbBaSummary* summary) {
  if (!s) THB_FAIL(THB_E_INVALID_ARGUMENT, "null session");
  int rc = THB_OK;
  s->sum.num_iterations = s->iteration;
  s->sum.final_cost = s->min_cost + s->fixed_cost;
  s->sum.success = s->sum.termination_type != THB_TERM_FAILURE;
  if (s->sum.success && s->any_variable) {
    const cudaMemcpyKind kout = s->prob.memory_space == THB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    cudaError_t e = cudaMemcpyAsync(s->prob.cam_ext, s->X.cam, sizeof(double) * s->nc * 6, kout, s->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(s->prob.pts, s->X.pts, sizeof(double) * s->np * 4, kout, s->st);
    if (e == cudaSuccess && s->nvg > 0) e = cudaMemcpyAsync(s->prob.intr, s->X.intr, sizeof(double) * s->ng * KS, kout, s->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);
    if (e != cudaSuccess) { SetLastError(cudaGetErrorString(e)); rc = THB_E_CUDA; }
  }
  cudaStreamSynchronize(s->st);
  s->sum.ms_jacobian += s->t_jac.CollectMs(); s->sum.ms_normal += s->t_normal.CollectMs();
  s->sum.ms_solve += s->t_solve.CollectMs(); s->sum.ms_update += s->t_update.CollectMs();
  s->sum.solve_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - s->t_solve_start).count();
  if (summary) *summary = s->sum;
  FreeSession(s);
  return rc;
}

int thb_ba_solve(const ThbBaProblem* problem, const ThbBaOptions* options, ThbBaSummary* summary, void* cuda_stream) {
  ThbBaSession* s = nullptr;
  int rc = thb_ba_create(problem, options, cuda_stream, &s);
  if (rc != THB_OK) return rc;
  rc = thb_ba_iterate(s, std::numeric_limits<int32_t>::max(), nullptr);
  if (rc != THB_OK) { FreeSession(s); return rc; }
  return thb_ba_finish(s, summary);
}

int thb_ba_time_jacobian(ThbBaSession* s, int32_t repeats, int32_t flush_l2, double* avg_ms) {
  if (!s || !avg_ms || repeats <= 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  // L2 flush = READ a buffer twice the size of the 126 MB L2, so the cache is full of CLEAN lines that belong to
  // nobody: a write-flush (memset) would leave dirty lines whose write-back gets billed to the timed kernel.
  const size_t flush_bytes = (size_t)256 << 20;
  if (flush_l2 && !s->d_flush) {
    THB_CUDA_CHECK(cudaMallocAsync(&s->d_flush, flush_bytes, s->st));
    s->arena.blocks.push_back(s->d_flush);
    THB_CUDA_CHECK(cudaMemsetAsync(s->d_flush, 0, flush_bytes, s->st));
    THB_CUDA_CHECK(cudaStreamSynchronize(s->st));
  }
  cudaEvent_t a, b;
  THB_CUDA_CHECK(cudaEventCreate(&a)); THB_CUDA_CHECK(cudaEventCreate(&b));
  double total = 0.0;
  for (int i = 0; i < repeats; ++i) {
    if (flush_l2) k_flush_read<<<148 * 8, 256, 0, s->st>>>(reinterpret_cast<const double2*>(s->d_flush), flush_bytes / sizeof(double2), s->d_scal + SC_SINK);
    THB_CUDA_CHECK(cudaEventRecord(a, s->st));
    RunJacobian(s, s->d_cs, s->d_ps);
    THB_CUDA_CHECK(cudaEventRecord(b, s->st));
    THB_CUDA_CHECK(cudaEventSynchronize(b));
    float ms = 0.f;
    THB_CUDA_CHECK(cudaEventElapsedTime(&ms, a, b));
    total += ms;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  // the cost accumulator was bumped by the extra launches: restore it
  THB_CUDA_CHECK(cudaMemsetAsync(s->d_scal + SC_COST_X, 0, sizeof(double), s->st));
  *avg_ms = total / repeats;
  return THB_OK;
}

int thb_dense_spd_solve(const double* A, const double* b, int32_t n, double* x, void* cuda_stream) {
  if (!A || !b || !x || n <= 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  DenseChol ch;
  ConfigurePoolOnce();
  if ((rc = ch.Init(n, st)) != THB_OK) { ch.Free(st); return rc; }
  DevBufs B;
  int* d_fail = B.get<int>(1);
  int launches = 0, h_fail = 0;
  cudaError_t e = cudaMemsetAsync(d_fail, 0, sizeof(int), st);
  if (e == cudaSuccess) rc = ch.Clear(st);
  if (e == cudaSuccess && rc == THB_OK) e = cudaMemcpy2DAsync(ch.A, sizeof(double) * ch.ld, A, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && rc == THB_OK) e = cudaMemcpyAsync(ch.RhsRow(), b, sizeof(double) * n, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && rc == THB_OK) rc = ch.FactorAndSolve(st, d_fail, &launches);
  if (e == cudaSuccess && rc == THB_OK) e = cudaMemcpyAsync(x, ch.x, sizeof(double) * n, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == THB_OK) e = cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == THB_OK) e = cudaStreamSynchronize(st);
  ch.Free(st);
  if (e != cudaSuccess) THB_FAIL(THB_E_CUDA, cudaGetErrorString(e));
  if (rc != THB_OK) return rc;
  if (h_fail) THB_FAIL(THB_E_NUMERICAL, "matrix is not positive definite");
  return THB_OK;
}

int thb_dense_spd_time(int32_t n, int32_t repeats, double* avg_ms, double* rel_residual, void* cuda_stream) {
  if (n <= 0 || repeats <= 0 || !avg_ms) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad argument");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  ConfigurePoolOnce();
  DenseChol ch;
  if ((rc = ch.Init(n, st)) != THB_OK) { ch.Free(st); return rc; }
  DevBufs B;
  int* d_fail = B.get<int>(1);
  double* d_res = B.get<double>(2);
  cudaEvent_t a = nullptr, b = nullptr;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double total = 0.0, h_res[2] = {0.0, 1.0};
  int launches = 0, h_fail = 0;
  cudaError_t e = cudaMemsetAsync(d_fail, 0, sizeof(int), st);
  for (int it = 0; it < repeats && e == cudaSuccess && rc == THB_OK; ++it) {
    rc = ch.Clear(st);
    k_synth_spd<<<dim3((ch.n_pad + 255) / 256, n + 1), 256, 0, st>>>(ch.A, ch.ld, n, ch.n_pad);
    cudaEventRecord(a, st);
    if (rc == THB_OK) rc = ch.FactorAndSolve(st, d_fail, &launches);
    cudaEventRecord(b, st);
    e = cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    total += ms;
  }
  if (e == cudaSuccess && rc == THB_OK) {
    cudaMemsetAsync(d_res, 0, 2 * sizeof(double), st);
    k_synth_spd_residual<<<(n + 127) / 128, 128, 0, st>>>(ch.x, n, d_res);
    cudaMemcpyAsync(h_res, d_res, 2 * sizeof(double), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  ch.Free(st);
  if (e != cudaSuccess) THB_FAIL(THB_E_CUDA, cudaGetErrorString(e));
  if (rc != THB_OK) return rc;
  if (h_fail) THB_FAIL(THB_E_NUMERICAL, "synthetic matrix not positive definite");
  *avg_ms = total / repeats;
  if (rel_residual) *rel_residual = h_res[1] > 0.0 ? h_res[0] / h_res[1] : 0.0;
  return THB_OK;
}

}  // extern "C"
namespace {
// shared by thb_ba_tracks_batch (rays == nullptr) and thb_estimate_tracks_batch. Host or device buffers: the observations are
// grouped by track ON THE DEVICE (histogram + stable radix sort, as the BA setup), so a host caller pays one H2D of its arrays
// and a device caller pays nothing but the kernels.
int RunTrackBatch(const ThbBaProblem* P, const ThbBaOptions* O, const double* rays, const ThbTrackEstimatorOptions* E, int32_t* status,
                  ThbTrackBaResult* results, void* cuda_stream) {
  if (!P || !O) THB_FAIL(THB_E_INVALID_ARGUMENT, "null problem or options");
  if (O->use_inner_iterations) THB_FAIL(THB_E_UNSUPPORTED, "BundleAdjustTrack runs without inner iterations (bundle_adjustment.cc:267)");
  const int sp = P->memory_space;
  if (sp != THB_MEM_HOST && sp != THB_MEM_DEVICE) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad memory_space");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  const int nc = P->num_cameras, ng = P->num_groups, np = P->num_points, no = P->num_observations;
  if (nc < 0 || ng < 0 || np < 0 || no < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "negative size");
  if (np == 0) return THB_OK;
  if (!P->pts || (no > 0 && (!P->cam_ext || !P->cam_group || !P->intr || !P->intr_model || !P->obs_cam || !P->obs_pt || !P->obs_xy)))
    THB_FAIL(THB_E_INVALID_ARGUMENT, "null array");
  std::vector<int> h_model;
  if ((rc = FetchToHost(P->intr_model, ng, sp, &h_model)) != THB_OK) return rc;
  for (int g = 0; g < ng; ++g) if (num_intrinsics(h_model[g]) < 0) THB_FAIL(THB_E_UNSUPPORTED, "camera model not on the hot path");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool host = sp == THB_MEM_HOST;
  const cudaMemcpyKind kin = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, kout = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const int PD = O->use_homogeneous_point_parametrization ? 3 : 4;
  // stream-ordered blocks from the device pool (kept across calls: TrackEstimator calls this once per batch of new tracks)
  ConfigurePoolOnce();
  struct ScopedArena : Arena { ~ScopedArena() { Release(); } } M;
  M.st = st;
  BaState X{}, Xc{};
  int *d_group = nullptr, *d_model = nullptr, *d_start = nullptr, *d_cstart = nullptr, *d_used = nullptr, *d_flags = nullptr, *d_perm = nullptr, *d_oc = nullptr,
      *d_op = nullptr, *d_raw_cam = nullptr, *d_raw_pt = nullptr;
  uint8_t *d_cc = nullptr, *d_pc = nullptr;
  double2 *d_xy = nullptr, *d_si = nullptr, *d_raw_xy = nullptr, *d_raw_si = nullptr;
  double *d_ps = nullptr, *d_orig = nullptr, *d_ray = nullptr, *d_raw_ray = nullptr;
  int* d_status = nullptr;
  ThbTrackBaResult* d_res = nullptr;
  const size_t no1 = std::max(no, 1);
  if ((rc = M.Get(&X.cam, (size_t)nc * 6)) != THB_OK || (rc = M.Get(&X.camd, (size_t)nc * CAMD)) != THB_OK || (rc = M.Get(&X.intr, (size_t)ng * KS)) != THB_OK ||
      (rc = M.Get(&X.pts, (size_t)np * 4)) != THB_OK || (rc = M.Get(&d_orig, (size_t)np * 4)) != THB_OK || (rc = M.Get(&d_group, nc)) != THB_OK ||
      (rc = M.Get(&d_model, ng)) != THB_OK || (rc = M.Get(&d_start, np + 1)) != THB_OK || (rc = M.Get(&d_cstart, nc + 1)) != THB_OK ||
      (rc = M.Get(&d_used, std::max(ng, 1))) != THB_OK || (rc = M.Get(&d_flags, SF_COUNT)) != THB_OK || (rc = M.Get(&d_perm, no1)) != THB_OK ||
      (rc = M.Get(&d_oc, no1)) != THB_OK || (rc = M.Get(&d_op, no1)) != THB_OK || (rc = M.Get(&d_cc, std::max(nc, 1))) != THB_OK || (rc = M.Get(&d_pc, np)) != THB_OK ||
      (rc = M.Get(&d_xy, no1)) != THB_OK || (rc = M.Get(&d_si, no1)) != THB_OK || (rc = M.Get(&d_ps, (size_t)np * PD)) != THB_OK || (rc = M.Get(&d_res, np)) != THB_OK)
    return rc;
  Xc = X;
  if ((rc = M.Get(&Xc.pts, (size_t)np * 4)) != THB_OK) return rc;
  // the caller's observation arrays: used in place when they are device memory
  const int *raw_cam = P->obs_cam, *raw_pt = P->obs_pt;
  const double2 *raw_xy = reinterpret_cast<const double2*>(P->obs_xy), *raw_si = reinterpret_cast<const double2*>(P->obs_sqrt_info);
  const double* raw_ray = rays;
  if (host && no > 0) {
    if ((rc = M.Get(&d_raw_cam, no1)) != THB_OK || (rc = M.Get(&d_raw_pt, no1)) != THB_OK || (rc = M.Get(&d_raw_xy, no1)) != THB_OK) return rc;
    THB_CUDA_CHECK(cudaMemcpyAsync(d_raw_cam, P->obs_cam, sizeof(int) * no, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(d_raw_pt, P->obs_pt, sizeof(int) * no, cudaMemcpyHostToDevice, st));
    THB_CUDA_CHECK(cudaMemcpyAsync(d_raw_xy, P->obs_xy, sizeof(double2) * no, cudaMemcpyHostToDevice, st));
    raw_cam = d_raw_cam; raw_pt = d_raw_pt; raw_xy = d_raw_xy;
    if (P->obs_sqrt_info) {
      if ((rc = M.Get(&d_raw_si, no1)) != THB_OK) return rc;
      THB_CUDA_CHECK(cudaMemcpyAsync(d_raw_si, P->obs_sqrt_info, sizeof(double2) * no, cudaMemcpyHostToDevice, st));
      raw_si = d_raw_si;
    }
    if (rays) {
      if ((rc = M.Get(&d_raw_ray, (size_t)no1 * 3)) != THB_OK) return rc;
      THB_CUDA_CHECK(cudaMemcpyAsync(d_raw_ray, rays, sizeof(double) * 3 * no, cudaMemcpyHostToDevice, st));
      raw_ray = d_raw_ray;
    }
  }
  if (rays && ((rc = M.Get(&d_ray, (size_t)no1 * 3)) != THB_OK || (rc = M.Get(&d_status, np)) != THB_OK)) return rc;
  THB_CUDA_CHECK(cudaMemcpyAsync(X.cam, P->cam_ext, sizeof(double) * nc * 6, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.intr, P->intr, sizeof(double) * ng * KS, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_orig, P->pts, sizeof(double) * np * 4, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.pts, d_orig, sizeof(double) * np * 4, cudaMemcpyDeviceToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_group, P->cam_group, sizeof(int) * nc, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_model, h_model.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cc, THB_CAM_CONST_ALL, std::max(nc, 1), st));
  if (P->pt_const) THB_CUDA_CHECK(cudaMemcpyAsync(d_pc, P->pt_const, np, kin, st));
  else THB_CUDA_CHECK(cudaMemsetAsync(d_pc, 0, np, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_start, 0, sizeof(int) * (np + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cstart, 0, sizeof(int) * (nc + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_used, 0, sizeof(int) * std::max(ng, 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_flags, 0, sizeof(int) * SF_COUNT, st));
  if (nc > 0) k_setup_check_groups<<<cdiv(nc, 256), 256, 0, st>>>(nc, ng, d_group, d_flags);
  if (no > 0) k_setup_count<<<cdiv(no, 256), 256, 0, st>>>(no, nc, np, ng, raw_cam, raw_pt, d_group, d_start, d_cstart, d_used, d_flags);
  int h_flags[SF_COUNT];
  THB_CUDA_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flags[SF_BAD_GROUP]) THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_group out of range");
  if (h_flags[SF_BAD_INDEX]) THB_FAIL(THB_E_INVALID_ARGUMENT, "observation index out of range");
  // observations grouped by track, in the caller's order inside a track (stable), as AddTrack walks track->ViewIds()
  if ((rc = GroupByKey(raw_pt, no, np, d_start, d_perm, st)) != THB_OK) return rc;
  if (no > 0) {
    k_setup_gather<<<cdiv(no, 256), 256, 0, st>>>(no, d_perm, raw_cam, raw_pt, raw_xy, raw_si, d_oc, d_op, d_xy, d_si);
    if (rays) k_gather_rays<<<cdiv(no, 256), 256, 0, st>>>(no, d_perm, raw_ray, d_ray);
  }
  if (nc > 0) k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(X.cam, X.camd, nc, nullptr, d_cc, d_group);
  BaConst K{};
  K.nc = nc; K.ng = ng; K.np = np; K.no = no;
  K.cam_group = d_group; K.intr_model = d_model; K.cam_const = d_cc; K.intr_const = nullptr; K.pt_const = d_pc; K.intr_slot = nullptr;
  K.loss_type = O->loss_function_type; K.loss_width = O->robust_loss_width;
  ObsSoA Ob{d_oc, nullptr, d_xy, d_si};
  TrackBaParams tp;
  tp.max_num_iterations = O->max_num_iterations; tp.max_invalid = O->max_num_consecutive_invalid_steps; tp.jacobi_scaling = O->jacobi_scaling;
  tp.ftol = O->function_tolerance; tp.gtol = O->gradient_tolerance; tp.ptol = O->parameter_tolerance;
  tp.radius0 = O->initial_trust_region_radius; tp.min_radius = O->min_trust_region_radius; tp.max_radius = O->max_trust_region_radius;
  tp.min_relative_decrease = O->min_relative_decrease; tp.min_diag = O->min_lm_diagonal; tp.max_diag = O->max_lm_diagonal;
  tp.rays = d_ray; tp.status = d_status; tp.bundle_adjustment = 1; tp.cos_min_angle = -2.0; tp.sq_max_reprojection_error = 0.0;
  if (rays) {
    tp.bundle_adjustment = E->bundle_adjustment;
    tp.cos_min_angle = std::cos(E->min_triangulation_angle_degrees * 3.14159265358979323846 / 180.0);  // DegToRad, triangulation.cc:240-241
    tp.sq_max_reprojection_error = E->max_acceptable_reprojection_error_pixels * E->max_acceptable_reprojection_error_pixels;
  }
  // 64-thread CTAs: a C5-sized batch (60k tracks) is only 0.4 threads per resident-thread slot of the GPU, small CTAs spread it
  // over all SMs (168 registers per thread). THB_TRACK_TIMING=1 prints the kernel's duration (tools/microbench/track_time.py).
  const bool timing = getenv("THB_TRACK_TIMING") != nullptr;
  const int cta = 64;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timing) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
  if (PD == 3) k_track_ba<3><<<cdiv(np, cta), cta, 0, st>>>(K, X, Xc, Ob, d_start, d_ps, tp, d_res);
  else k_track_ba<4><<<cdiv(np, cta), cta, 0, st>>>(K, X, Xc, Ob, d_start, d_ps, tp, d_res);
  THB_CUDA_CHECK(cudaGetLastError());
  if (timing) {
    cudaEventRecord(e1, st); cudaEventSynchronize(e1);
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "k_track_ba<%d>: %d tracks, %d-thread CTAs, %.3f ms\n", PD, np, cta, ms);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  // IsSolutionUsable() == false leaves the track as it was; EstimateTrack keeps only estimated tracks
  k_track_commit<<<cdiv(np, 256), 256, 0, st>>>(np, d_status, d_res, X.pts, d_orig);
  THB_CUDA_CHECK(cudaMemcpyAsync(P->pts, d_orig, sizeof(double) * np * 4, kout, st));
  if (results) THB_CUDA_CHECK(cudaMemcpyAsync(results, d_res, sizeof(ThbTrackBaResult) * np, kout, st));
  if (rays) THB_CUDA_CHECK(cudaMemcpyAsync(status, d_status, sizeof(int32_t) * np, kout, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}
}  // namespace
extern "C" {

int thb_ba_tracks_batch(const ThbBaProblem* P, const ThbBaOptions* O, ThbTrackBaResult* results, void* cuda_stream) {
  return RunTrackBatch(P, O, nullptr, nullptr, nullptr, results, cuda_stream);
}

int thb_estimate_tracks_batch(const ThbBaProblem* P, const double* ray_directions, const ThbTrackEstimatorOptions* E, const ThbBaOptions* O,
                              int32_t* status, ThbTrackBaResult* results, void* cuda_stream) {
  if (!ray_directions || !E || !status) THB_FAIL(THB_E_INVALID_ARGUMENT, "null ray_directions, options or status");
  return RunTrackBatch(P, O, ray_directions, E, status, results, cuda_stream);
}

int thb_set_outlier_tracks_batch(const ThbBaProblem* P, double max_err, double min_angle_deg, int32_t* status, int32_t* num_removed, void* cuda_stream) {
  if (!P || !status) THB_FAIL(THB_E_INVALID_ARGUMENT, "null problem or status");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  const int nc = P->num_cameras, ng = P->num_groups, np = P->num_points, no = P->num_observations, sp = P->memory_space;
  if (nc < 0 || ng < 0 || np < 0 || no < 0 || (sp != THB_MEM_HOST && sp != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad size or memory space");
  if (num_removed && sp == THB_MEM_HOST) *num_removed = 0;
  if (np == 0) return THB_OK;
  if (!P->pts || (no > 0 && (!P->cam_ext || !P->cam_group || !P->intr || !P->intr_model || !P->obs_cam || !P->obs_pt || !P->obs_xy)))
    THB_FAIL(THB_E_INVALID_ARGUMENT, "null array");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool host = sp == THB_MEM_HOST;
  const cudaMemcpyKind kin = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  ConfigurePoolOnce();
  struct ScopedArena : Arena { ~ScopedArena() { Release(); } } M;
  M.st = st;
  std::vector<int> h_model;
  if ((rc = FetchToHost(P->intr_model, ng, sp, &h_model)) != THB_OK) return rc;
  for (int g = 0; g < ng; ++g) if (num_intrinsics(h_model[g]) < 0) THB_FAIL(THB_E_UNSUPPORTED, "camera model not on the hot path");
  BaState X{};
  int *d_group = nullptr, *d_model = nullptr, *d_start = nullptr, *d_perm = nullptr, *d_oc = nullptr, *d_op = nullptr, *d_status = nullptr, *d_flags = nullptr, *d_cstart = nullptr, *d_used = nullptr;
  uint8_t *d_cc = nullptr, *d_pc = nullptr;
  double2* d_xy = nullptr;
  if ((rc = M.Get(&X.cam, (size_t)nc * 6)) != THB_OK || (rc = M.Get(&X.camd, (size_t)nc * CAMD)) != THB_OK || (rc = M.Get(&X.intr, (size_t)ng * KS)) != THB_OK ||
      (rc = M.Get(&X.pts, (size_t)np * 4)) != THB_OK || (rc = M.Get(&d_group, nc)) != THB_OK || (rc = M.Get(&d_model, ng)) != THB_OK ||
      (rc = M.Get(&d_start, np + 1)) != THB_OK || (rc = M.Get(&d_perm, no)) != THB_OK || (rc = M.Get(&d_oc, no)) != THB_OK || (rc = M.Get(&d_op, no)) != THB_OK ||
      (rc = M.Get(&d_xy, no)) != THB_OK || (rc = M.Get(&d_cc, nc)) != THB_OK || (rc = M.Get(&d_pc, np)) != THB_OK || (rc = M.Get(&d_status, np)) != THB_OK ||
      (rc = M.Get(&d_flags, SF_COUNT + 1)) != THB_OK || (rc = M.Get(&d_cstart, nc + 1)) != THB_OK || (rc = M.Get(&d_used, ng)) != THB_OK)
    return rc;
  THB_CUDA_CHECK(cudaMemcpyAsync(X.cam, P->cam_ext, sizeof(double) * 6 * nc, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.intr, P->intr, sizeof(double) * KS * ng, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.pts, P->pts, sizeof(double) * 4 * np, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_group, P->cam_group, sizeof(int) * nc, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_model, h_model.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_oc, P->obs_cam, sizeof(int) * no, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_op, P->obs_pt, sizeof(int) * no, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_xy, P->obs_xy, sizeof(double2) * no, kin, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cc, 0, std::max(nc, 1), st));
  if (P->pt_const) THB_CUDA_CHECK(cudaMemcpyAsync(d_pc, P->pt_const, np, kin, st));
  else THB_CUDA_CHECK(cudaMemsetAsync(d_pc, 0, np, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_start, 0, sizeof(int) * (np + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cstart, 0, sizeof(int) * (nc + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_used, 0, sizeof(int) * std::max(ng, 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_flags, 0, sizeof(int) * (SF_COUNT + 1), st));
  if (nc > 0) k_setup_check_groups<<<cdiv(nc, 256), 256, 0, st>>>(nc, ng, d_group, d_flags);
  if (no > 0) k_setup_count<<<cdiv(no, 256), 256, 0, st>>>(no, nc, np, ng, d_oc, d_op, d_group, d_start, d_cstart, d_used, d_flags);
  int h_flags[SF_COUNT + 1];
  THB_CUDA_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flags[SF_BAD_GROUP]) THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_group out of range");
  if (h_flags[SF_BAD_INDEX]) THB_FAIL(THB_E_INVALID_ARGUMENT, "observation index out of range");
  if ((rc = GroupByKey(d_op, no, np, d_start, d_perm, st)) != THB_OK) return rc;
  if (nc > 0) k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(X.cam, X.camd, nc, nullptr, d_cc, d_group);
  BaConst K{};
  K.nc = nc; K.ng = ng; K.np = np; K.no = no; K.cam_group = d_group; K.intr_model = d_model; K.cam_const = d_cc; K.pt_const = d_pc;
  K.loss_type = THB_LOSS_TRIVIAL; K.loss_width = 1.0;
  int* d_removed = d_flags + SF_COUNT;
  k_outlier_tracks<<<cdiv(np, 128), 128, 0, st>>>(K, X, d_start, d_perm, d_oc, d_xy, max_err * max_err,
                                                 std::cos(min_angle_deg * 3.14159265358979323846 / 180.0), d_status, d_removed);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(status, d_status, sizeof(int32_t) * np, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  int h_removed = 0;
  THB_CUDA_CHECK(cudaMemcpyAsync(&h_removed, d_removed, sizeof(int), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (num_removed) *num_removed = h_removed;
  return THB_OK;
}

int thb_select_good_tracks_batch(const ThbBaProblem* P, const uint8_t* cam_selected, int32_t long_thr, int32_t cell_size, int32_t min_per_view,
                                 uint8_t* selected, int32_t* num_selected, void* cuda_stream) {
  if (!P || !selected) THB_FAIL(THB_E_INVALID_ARGUMENT, "null problem or selected");
  if (cell_size <= 0 || long_thr < 0 || min_per_view < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad grid cell size / thresholds");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  const int nc = P->num_cameras, ng = P->num_groups, np = P->num_points, no = P->num_observations, sp = P->memory_space;
  if (nc < 0 || ng < 0 || np < 0 || no < 0 || (sp != THB_MEM_HOST && sp != THB_MEM_DEVICE)) THB_FAIL(THB_E_INVALID_ARGUMENT, "bad size or memory space");
  if (nc >= (1 << 22)) THB_FAIL(THB_E_UNSUPPORTED, "more than 4M views");
  if (num_selected) *num_selected = 0;
  if (np == 0) return THB_OK;
  if (!P->pts || (no > 0 && (!P->cam_ext || !P->cam_group || !P->intr || !P->intr_model || !P->obs_cam || !P->obs_pt || !P->obs_xy)))
    THB_FAIL(THB_E_INVALID_ARGUMENT, "null array");
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool host = sp == THB_MEM_HOST;
  const cudaMemcpyKind kin = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  ConfigurePoolOnce();
  struct ScopedArena : Arena { ~ScopedArena() { Release(); } } M;
  M.st = st;
  std::vector<int> h_model;
  if ((rc = FetchToHost(P->intr_model, ng, sp, &h_model)) != THB_OK) return rc;
  for (int g = 0; g < ng; ++g) if (num_intrinsics(h_model[g]) < 0) THB_FAIL(THB_E_UNSUPPORTED, "camera model not on the hot path");
  BaState X{};
  int *d_group = nullptr, *d_model = nullptr, *d_start = nullptr, *d_perm = nullptr, *d_oc = nullptr, *d_op = nullptr, *d_flags = nullptr, *d_cstart = nullptr,
      *d_used = nullptr, *d_len = nullptr, *d_iota = nullptr, *d_sorted_obs = nullptr, *d_cperm = nullptr;
  uint8_t *d_cc = nullptr, *d_sel = nullptr, *d_csel = nullptr;
  double2* d_xy = nullptr;
  double* d_mean = nullptr;
  unsigned long long *d_ckey = nullptr, *d_tkey = nullptr, *d_ckey_s = nullptr, *d_tkey_s = nullptr;
  const size_t no1 = std::max(no, 1);
  if ((rc = M.Get(&X.cam, (size_t)nc * 6)) != THB_OK || (rc = M.Get(&X.camd, (size_t)nc * CAMD)) != THB_OK || (rc = M.Get(&X.intr, (size_t)ng * KS)) != THB_OK ||
      (rc = M.Get(&X.pts, (size_t)np * 4)) != THB_OK || (rc = M.Get(&d_group, nc)) != THB_OK || (rc = M.Get(&d_model, ng)) != THB_OK ||
      (rc = M.Get(&d_start, np + 1)) != THB_OK || (rc = M.Get(&d_perm, no1)) != THB_OK || (rc = M.Get(&d_oc, no1)) != THB_OK || (rc = M.Get(&d_op, no1)) != THB_OK ||
      (rc = M.Get(&d_xy, no1)) != THB_OK || (rc = M.Get(&d_cc, std::max(nc, 1))) != THB_OK || (rc = M.Get(&d_sel, np)) != THB_OK || (rc = M.Get(&d_csel, std::max(nc, 1))) != THB_OK ||
      (rc = M.Get(&d_flags, SF_COUNT + 1)) != THB_OK || (rc = M.Get(&d_cstart, nc + 1)) != THB_OK || (rc = M.Get(&d_used, std::max(ng, 1))) != THB_OK ||
      (rc = M.Get(&d_len, np)) != THB_OK || (rc = M.Get(&d_mean, np)) != THB_OK || (rc = M.Get(&d_iota, no1)) != THB_OK || (rc = M.Get(&d_sorted_obs, no1)) != THB_OK ||
      (rc = M.Get(&d_cperm, no1)) != THB_OK || (rc = M.Get(&d_ckey, no1)) != THB_OK || (rc = M.Get(&d_tkey, no1)) != THB_OK || (rc = M.Get(&d_ckey_s, no1)) != THB_OK ||
      (rc = M.Get(&d_tkey_s, no1)) != THB_OK)
    return rc;
  THB_CUDA_CHECK(cudaMemcpyAsync(X.cam, P->cam_ext, sizeof(double) * 6 * nc, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.intr, P->intr, sizeof(double) * KS * ng, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.pts, P->pts, sizeof(double) * 4 * np, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_group, P->cam_group, sizeof(int) * nc, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_model, h_model.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_oc, P->obs_cam, sizeof(int) * no, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_op, P->obs_pt, sizeof(int) * no, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_xy, P->obs_xy, sizeof(double2) * no, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_sel, selected, np, kin, st));
  if (cam_selected) THB_CUDA_CHECK(cudaMemcpyAsync(d_csel, cam_selected, nc, kin, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cc, 0, std::max(nc, 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_start, 0, sizeof(int) * (np + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cstart, 0, sizeof(int) * (nc + 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_used, 0, sizeof(int) * std::max(ng, 1), st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_flags, 0, sizeof(int) * (SF_COUNT + 1), st));
  if (nc > 0) k_setup_check_groups<<<cdiv(nc, 256), 256, 0, st>>>(nc, ng, d_group, d_flags);
  if (no > 0) k_setup_count<<<cdiv(no, 256), 256, 0, st>>>(no, nc, np, ng, d_oc, d_op, d_group, d_start, d_cstart, d_used, d_flags);
  int h_flags[SF_COUNT + 1];
  THB_CUDA_CHECK(cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (h_flags[SF_BAD_GROUP]) THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_group out of range");
  if (h_flags[SF_BAD_INDEX]) THB_FAIL(THB_E_INVALID_ARGUMENT, "observation index out of range");
  if (no > 0) {
    if ((rc = GroupByKey(d_op, no, np, d_start, d_perm, st)) != THB_OK) return rc;
    if ((rc = GroupByKey(d_oc, no, nc, d_cstart, d_cperm, st)) != THB_OK) return rc;  // only the per-view offsets are used
    k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(X.cam, X.camd, nc, nullptr, d_cc, d_group);
    BaConst K{};
    K.nc = nc; K.ng = ng; K.np = np; K.no = no; K.cam_group = d_group; K.intr_model = d_model; K.cam_const = d_cc; K.pt_const = nullptr;
    K.loss_type = THB_LOSS_TRIVIAL; K.loss_width = 1.0;
    k_track_stats<<<cdiv(np, 128), 128, 0, st>>>(K, X, d_start, d_perm, d_oc, d_xy, long_thr, d_len, d_mean);
    k_select_keys<<<cdiv(no, 256), 256, 0, st>>>(no, d_oc, d_op, d_xy, cam_selected ? d_csel : nullptr, 1.0 / (double)cell_size, d_ckey, d_tkey, d_iota);
    if ((rc = SortPairsU64(d_ckey, d_ckey_s, d_iota, d_sorted_obs, no, st)) != THB_OK) return rc;
    k_select_cells<<<cdiv(no, 256), 256, 0, st>>>(no, d_ckey_s, d_sorted_obs, d_op, d_len, d_mean, d_sel);
    if ((rc = SortPairsU64(d_tkey, d_tkey_s, d_iota, d_cperm, no, st)) != THB_OK) return rc;
    k_select_top_ranked<<<1, 1024, 0, st>>>(nc, cam_selected ? d_csel : nullptr, d_cstart, d_tkey_s, min_per_view, d_sel);
  }
  int* d_count = d_flags + SF_COUNT;
  THB_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(int), st));
  k_count_selected<<<cdiv(np, 256), 256, 0, st>>>(np, d_sel, d_count);
  THB_CUDA_CHECK(cudaGetLastError());
  THB_CUDA_CHECK(cudaMemcpyAsync(selected, d_sel, np, host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  int h_count = 0;
  THB_CUDA_CHECK(cudaMemcpyAsync(&h_count, d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (num_selected) *num_selected = h_count;
  return THB_OK;
}

int thb_ba_evaluate(const ThbBaProblem* P, double* residuals, double* jac_cam, double* jac_intr, double* jac_pt,
                    uint8_t* ok, void* cuda_stream) {
  if (!P) THB_FAIL(THB_E_INVALID_ARGUMENT, "null problem");
  int rc = CheckDevice();
  if (rc != THB_OK) return rc;
  const int nc = P->num_cameras, ng = P->num_groups, np = P->num_points, no = P->num_observations;
  if (nc < 0 || ng < 0 || np < 0 || no < 0) THB_FAIL(THB_E_INVALID_ARGUMENT, "negative size");
  if (no == 0) return THB_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool host = P->memory_space == THB_MEM_HOST;
  const cudaMemcpyKind kin = host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  const cudaMemcpyKind kout = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  std::vector<int> h_group, h_model, h_cam, h_pt;
  if ((rc = FetchToHost(P->cam_group, nc, P->memory_space, &h_group)) != THB_OK) return rc;
  if ((rc = FetchToHost(P->intr_model, ng, P->memory_space, &h_model)) != THB_OK) return rc;
  if ((rc = FetchToHost(P->obs_cam, no, P->memory_space, &h_cam)) != THB_OK) return rc;
  if ((rc = FetchToHost(P->obs_pt, no, P->memory_space, &h_pt)) != THB_OK) return rc;
  for (int g = 0; g < ng; ++g) if (num_intrinsics(h_model[g]) < 0) THB_FAIL(THB_E_UNSUPPORTED, "camera model not on the hot path");
  for (int c = 0; c < nc; ++c) if (h_group[c] < 0 || h_group[c] >= ng) THB_FAIL(THB_E_INVALID_ARGUMENT, "cam_group out of range");
  for (int i = 0; i < no; ++i) if (h_cam[i] < 0 || h_cam[i] >= nc || h_pt[i] < 0 || h_pt[i] >= np) THB_FAIL(THB_E_INVALID_ARGUMENT, "observation index out of range");
  DevBufs B;
  BaState X{B.get<double>((size_t)nc * 6), B.get<double>((size_t)nc * CAMD), B.get<double>((size_t)ng * KS), B.get<double>((size_t)np * 4)};
  int* d_group = B.get<int>(nc); int* d_model = B.get<int>(ng); int* d_slot = B.get<int>(ng);
  uint8_t* d_cc = B.get<uint8_t>(nc); uint8_t* d_pc = B.get<uint8_t>(np); uint16_t* d_ic = B.get<uint16_t>(ng);
  int* d_oc = B.get<int>(no); int* d_op = B.get<int>(no);
  double2* d_xy = B.get<double2>(no); double2* d_si = B.get<double2>(no);
  double* d_res = B.get<double>((size_t)no * 2); double* d_jc = B.get<double>((size_t)no * 12);
  double* d_ji = B.get<double>((size_t)no * 2 * KS); double* d_jp = B.get<double>((size_t)no * 8);
  uint8_t* d_ok = B.get<uint8_t>(no);
  if (!X.cam || !X.camd || !X.intr || !X.pts || !d_group || !d_model || !d_slot || !d_cc || !d_pc || !d_ic || !d_oc || !d_op || !d_xy ||
      !d_si || !d_res || !d_jc || !d_ji || !d_jp || !d_ok) THB_FAIL(THB_E_CUDA, "cudaMalloc failed");
  THB_CUDA_CHECK(cudaMemcpyAsync(X.cam, P->cam_ext, sizeof(double) * nc * 6, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.intr, P->intr, sizeof(double) * ng * KS, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(X.pts, P->pts, sizeof(double) * np * 4, kin, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_group, h_group.data(), sizeof(int) * nc, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_model, h_model.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  std::vector<int> slot(ng);
  for (int g = 0; g < ng; ++g) slot[g] = g;
  THB_CUDA_CHECK(cudaMemcpyAsync(d_slot, slot.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_cc, 0, nc, st)); THB_CUDA_CHECK(cudaMemsetAsync(d_pc, 0, np, st));
  THB_CUDA_CHECK(cudaMemsetAsync(d_ic, 0, sizeof(uint16_t) * ng, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_oc, h_cam.data(), sizeof(int) * no, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_op, h_pt.data(), sizeof(int) * no, cudaMemcpyHostToDevice, st));
  THB_CUDA_CHECK(cudaMemcpyAsync(d_xy, P->obs_xy, sizeof(double) * 2 * no, kin, st));
  if (P->obs_sqrt_info) THB_CUDA_CHECK(cudaMemcpyAsync(d_si, P->obs_sqrt_info, sizeof(double) * 2 * no, kin, st));
  else k_fill<<<cdiv(2LL * no, 256), 256, 0, st>>>(2 * no, reinterpret_cast<double*>(d_si), 1.0);
  BaConst K{nc, ng, np, no, d_group, d_model, d_cc, d_ic, d_pc, d_slot, THB_LOSS_TRIVIAL, 1.0};
  ObsSoA O{d_oc, d_op, d_xy, d_si};
  k_cam_derive<<<cdiv(nc, 128), 128, 0, st>>>(X.cam, X.camd, nc, nullptr, d_cc, d_group);
  k_eval_ambient<<<cdiv(no, 128), 128, 0, st>>>(K, X, O, d_res, d_jc, d_ji, d_jp, d_ok);
  THB_CUDA_CHECK(cudaGetLastError());
  if (residuals) THB_CUDA_CHECK(cudaMemcpyAsync(residuals, d_res, sizeof(double) * 2 * no, kout, st));
  if (jac_cam) THB_CUDA_CHECK(cudaMemcpyAsync(jac_cam, d_jc, sizeof(double) * 12 * no, kout, st));
  if (jac_intr) THB_CUDA_CHECK(cudaMemcpyAsync(jac_intr, d_ji, sizeof(double) * 2 * KS * no, kout, st));
  if (jac_pt) THB_CUDA_CHECK(cudaMemcpyAsync(jac_pt, d_jp, sizeof(double) * 8 * no, kout, st));
  if (ok) THB_CUDA_CHECK(cudaMemcpyAsync(ok, d_ok, no, kout, st));
  THB_CUDA_CHECK(cudaStreamSynchronize(st));
  return THB_OK;
}

}  // extern "C"
