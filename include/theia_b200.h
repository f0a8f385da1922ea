/*
 * theia_b200.h — C-ABI of the B200-native bundle-adjustment / RANSAC hot paths.
 *
 * The reference (urbste/pyTheiaSfM) has no FFI seam for these paths: its seam is the
 * C++ free-function layer theia::BundleAdjust* / theia::Estimate*.  The entry points
 * declared here are what an adapter for that layer binds (gather -> C-ABI -> scatter);
 * each one names the reference function it stands behind (paths relative to
 * /root/reference/src/theia).  Plain pointers and sizes only; no C++ or torch types.
 *
 * All arithmetic is FP64 (the reference is FP64 throughout); ids are int32.
 * Every function returns THB_OK (0) or a negative THB_E_* code.  Nothing aborts or
 * throws across this boundary (the reference CHECK-aborts on contract violations,
 * e.g. sfm/bundle_adjustment/bundle_adjuster.cc:95,117).
 */
#ifndef THEIA_B200_H_
#define THEIA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define THB_OK 0
#define THB_E_INVALID_ARGUMENT (-1) /* null pointer, bad size, index out of range      */
#define THB_E_UNSUPPORTED (-2)      /* option combination the path does not implement  */
#define THB_E_CUDA (-3)             /* CUDA runtime error; see thb_last_error()        */
#define THB_E_NO_DEVICE (-4)        /* no sm_100 device visible: there is NO CPU path  */
#define THB_E_NUMERICAL (-5)        /* evaluation failed at the initial point          */

#define THB_MEM_HOST 0
#define THB_MEM_DEVICE 1

/* sfm/camera/camera_intrinsics_model_type.h:46-56 */
#define THB_MODEL_PINHOLE 0
#define THB_MODEL_PINHOLE_RADIAL_TANGENTIAL 1 /* not on the hot path: rejected */
#define THB_MODEL_FISHEYE 2
#define THB_MODEL_FOV 3
#define THB_MODEL_DIVISION_UNDISTORTION 4
#define THB_MODEL_DOUBLE_SPHERE 5
#define THB_MODEL_EXTENDED_UNIFIED 6

/* Intrinsics blocks are padded to this many doubles per group
 * (largest named model: FISHEYE, 9; sfm/camera/fisheye_camera_model.h:67-77). */
#define THB_INTR_STRIDE 10

/* sfm/bundle_adjustment/create_loss_function.h:51-59 */
#define THB_LOSS_TRIVIAL 0
#define THB_LOSS_HUBER 1
#define THB_LOSS_SOFTLONE 2
#define THB_LOSS_CAUCHY 3
#define THB_LOSS_ARCTAN 4
#define THB_LOSS_TUKEY 5
#define THB_LOSS_TRUNCATED 6

/* How the reduced camera system is solved (ceres::LinearSolverType as exposed by
 * BundleAdjustmentOptions::linear_solver_type, bundle_adjustment.h:98-100). */
#define THB_SOLVER_SCHUR_CHOLESKY 0 /* exact: DENSE_SCHUR / SPARSE_SCHUR / DENSE_QR   */
#define THB_SOLVER_SCHUR_PCG 1      /* ITERATIVE_SCHUR with SCHUR_JACOBI              */

/* Camera extrinsics constness bits (bundle_adjuster.cc:357-380, 478-507). */
#define THB_CAM_CONST_POSITION 1
#define THB_CAM_CONST_ORIENTATION 2
#define THB_CAM_CONST_ALL 3

/* ceres::TerminationType values reported in ThbBaSummary.termination_type */
#define THB_TERM_CONVERGENCE 0
#define THB_TERM_NO_CONVERGENCE 1
#define THB_TERM_FAILURE 2

/*
 * A bundle-adjustment problem in structure-of-arrays form: what
 * BundleAdjuster::AddView/AddTrack (bundle_adjuster.cc:116-221) would have registered
 * with ceres::Problem, flattened.  One observation = one ReprojectionError residual
 * block (sfm/camera/reprojection_error.h:49-114) on (cam_ext[obs_cam], intr[group],
 * pts[obs_pt]).  Parameter arrays are updated IN PLACE, like the reference mutates
 * Camera::camera_parameters_ / Track::point_ through raw double* (bundle_adjuster.cc:585-591).
 */
typedef struct ThbBaProblem {
  int32_t num_cameras;
  int32_t num_groups; /* shared intrinsics blocks (reconstruction.cc:129-140)        */
  int32_t num_points;
  int32_t num_observations;
  int32_t memory_space; /* THB_MEM_HOST or THB_MEM_DEVICE for EVERY pointer below    */
  int32_t reserved0;

  double* cam_ext;           /* [num_cameras*6]  [C(3), angle-axis(3)]  camera.h:202-204   */
  const uint8_t* cam_const;  /* [num_cameras]    THB_CAM_CONST_* bits; may be NULL (=0)    */
  const int32_t* cam_group;  /* [num_cameras]    index into intr / intr_model              */
  double* intr;              /* [num_groups*THB_INTR_STRIDE]  per-model layout, SURVEY B   */
  const int32_t* intr_model; /* [num_groups]     THB_MODEL_*                               */
  const uint16_t* intr_const;/* [num_groups]     bit k set => parameter k held constant;   */
                             /*                  all K bits set => block constant; NULL => all constant */
  double* pts;               /* [num_points*4]   homogeneous [X,Y,Z,W]   track.h           */
  const uint8_t* pt_const;   /* [num_points]     1 => constant point; may be NULL (=0)     */

  const int32_t* obs_cam;    /* [num_observations] */
  const int32_t* obs_pt;     /* [num_observations] */
  const double* obs_xy;      /* [num_observations*2]  Feature::point_                      */
  const double* obs_sqrt_info;/*[num_observations*2]  1/sqrt(cov(0,0)), 1/sqrt(cov(1,1));  */
                             /*                  NULL => 1 (reprojection_error.h:96-103)   */
  /* BundleAdjustmentOptions::use_position_priors (bundle_adjuster.cc:160-163, position_error.h:44-80): a camera with
   * cam_has_position_prior[c] != 0 gets the 3 residuals sqrt_info * (prior - position), no loss function. All three
   * pointers NULL (a zero-initialised struct) = no priors. A prior on a fully constant camera is a constant of the cost and
   * is left out. thb_ba_solve / create / iterate / covariance honour them; the track entries keep every camera constant. */
  const uint8_t* cam_has_position_prior;      /* [num_cameras]   View::HasPositionPrior()                          */
  const double* cam_position_prior;           /* [num_cameras*3] View::GetPositionPrior()                          */
  const double* cam_position_prior_sqrt_info; /* [num_cameras*9] View::GetPositionPriorSqrtInformation(), row-major */
  /* BundleAdjustmentOptions::use_gravity_priors (bundle_adjuster.cc:165-168, gravity_error.h:44-86): the 3 residuals
   * sqrt_info * (R(angle_axis) * (0, 0, -1) - gravity_prior) on the camera orientation, no loss function. Same conventions. */
  const uint8_t* cam_has_gravity_prior;       /* [num_cameras]   View::HasGravityPrior()                           */
  const double* cam_gravity_prior;            /* [num_cameras*3] View::GetGravityPrior()                           */
  const double* cam_gravity_prior_sqrt_info;  /* [num_cameras*9] View::GetGravityPriorSqrtInformation(), row-major  */
  /* BundleAdjustmentOptions::use_orientation_priors (bundle_adjuster.cc:170-172, orientation_error.h:44-80): the 3 residuals
   * sqrt_info * log(exp(angle_axis) * exp(orientation_prior)^-1) (SO(3) logarithm, Sophus) on the camera orientation, no loss. */
  const uint8_t* cam_has_orientation_prior;       /* [num_cameras]   View::HasOrientationPrior()                          */
  const double* cam_orientation_prior;            /* [num_cameras*3] View::GetOrientationPrior() (angle-axis)             */
  const double* cam_orientation_prior_sqrt_info;  /* [num_cameras*9] View::GetOrientationPriorSqrtInformation(), row-major */
} ThbBaProblem;

/* BundleAdjustmentOptions (bundle_adjustment.h:87-167) restricted to what reaches
 * ceres::Solver::Options (bundle_adjuster.cc:63-89), plus the Ceres trust-region
 * defaults Theia leaves untouched. Fill with thb_ba_default_options. */
typedef struct ThbBaOptions {
  int32_t loss_function_type; /* THB_LOSS_*                                           */
  int32_t linear_solver;      /* THB_SOLVER_*                                         */
  int32_t use_homogeneous_point_parametrization; /* SphereManifold<4> on points       */
  int32_t use_inner_iterations; /* ceres inner iterations with Theia's reversed ordering (bundle_adjuster.cc:329-334) */
  int32_t max_num_iterations;
  int32_t jacobi_scaling;     /* Ceres default true                                   */
  int32_t verbose;
  int32_t max_num_consecutive_invalid_steps; /* Ceres default 5                       */
  double robust_loss_width;
  double function_tolerance;  /* the three tolerances: Ceres semantics; a NEGATIVE value disables */
  double gradient_tolerance;  /* the test (extension used by benchmark drivers to force exactly   */
  double parameter_tolerance; /* max_num_iterations iterations; Ceres itself rejects negatives)   */
  double max_trust_region_radius;
  double initial_trust_region_radius; /* 1e4   */
  double min_trust_region_radius;     /* 1e-32 */
  double min_relative_decrease;       /* 1e-3  */
  double min_lm_diagonal;             /* 1e-6  */
  double max_lm_diagonal;             /* 1e32  */
  double max_solver_time_in_seconds;
  /* THB_SOLVER_SCHUR_PCG only (ceres ITERATIVE_SCHUR + SCHUR_JACOBI) */
  double pcg_eta;                     /* Solver::Options::eta, the forcing term: CG stops when the relative decrease of
                                         the quadratic model, i * (Q_i - Q_{i-1}) / Q_i, drops below it; default 0.1   */
  int32_t pcg_max_iterations;         /* Solver::Options::max_linear_solver_iterations; default 500                    */
  int32_t reserved0;
} ThbBaOptions;

#define THB_MAX_ITER_LOG 256

/* BundleAdjustmentSummary (bundle_adjustment.h:170-178) + the iteration record the
 * reference does not expose, + device timings for the roofline report. */
typedef struct ThbBaSummary {
  int32_t success;          /* ceres IsSolutionUsable()  bundle_adjuster.cc:352        */
  int32_t termination_type; /* THB_TERM_*                                              */
  int32_t num_iterations;   /* LM iterations attempted (excl. iteration 0)             */
  int32_t num_successful_steps;
  int32_t num_jacobian_evaluations;
  int32_t num_cost_evaluations;
  int32_t num_linear_solves;
  int32_t gpu_launches;     /* kernels launched by this call                           */
  double initial_cost;      /* 0.5 * sum rho(|r|^2)                                     */
  double final_cost;
  double setup_time_in_seconds;
  double solve_time_in_seconds;
  /* device time (CUDA events on the caller's stream), milliseconds, summed over calls */
  double ms_jacobian;       /* K1 residual+Jacobian kernel                             */
  double ms_normal;         /* K2/K3 block assembly + Schur complement build           */
  double ms_solve;          /* K4 reduced camera system factor/solve                   */
  double ms_update;         /* K5 back-substitution, Plus, cost evaluation             */
  int32_t iter_log_count;
  int32_t num_linear_solver_iterations; /* THB_SOLVER_SCHUR_PCG: conjugate-gradient iterations over all linear solves */
  double iter_cost[THB_MAX_ITER_LOG];   /* cost after each iteration (index 0 = initial) */
  double iter_radius[THB_MAX_ITER_LOG]; /* trust-region radius after each iteration      */
} ThbBaSummary;

/* Version / diagnostics. */
int thb_version(void);
const char* thb_last_error(void); /* thread-local text of the last THB_E_* */
int thb_device_count(void);

void thb_ba_default_options(ThbBaOptions* options);

/*
 * One-shot solve: theia::BundleAdjust{Reconstruction,PartialReconstruction,View(s),
 * Track(s)} after the adapter has flattened the Reconstruction
 * (bundle_adjustment.cc:111-143,188-217,220-258,261-285); the work done is that of
 * BundleAdjuster::Optimize -> ceres::Solve (bundle_adjuster.cc:315-355).
 * With THB_MEM_HOST pointers the H2D/D2H copies are part of the call.
 * cuda_stream: a cudaStream_t (may be NULL for the default stream).
 */
int thb_ba_solve(const ThbBaProblem* problem, const ThbBaOptions* options,
                 ThbBaSummary* summary, void* cuda_stream);

/*
 * Session form of the same solve, for callers that keep the problem resident in HBM
 * (benchmark drivers; incremental pipelines that re-run BA on a growing reconstruction).
 *   create  : validate, upload/alias inputs, build the point-major observation order,
 *             evaluate iteration 0 (cost, Jacobian, gradient).
 *   iterate : run up to n further LM iterations (stops early on convergence/failure);
 *             returns the number actually run in *ran.
 *   finish  : write parameters back to the problem's arrays, fill the summary, free.
 */
typedef struct ThbBaSession ThbBaSession;
int thb_ba_create(const ThbBaProblem* problem, const ThbBaOptions* options,
                  void* cuda_stream, ThbBaSession** session);
int thb_ba_iterate(ThbBaSession* session, int32_t n, int32_t* ran);
int thb_ba_finish(ThbBaSession* session, ThbBaSummary* summary);

/*
 * Kernel-level entry used by the parity tests and the roofline measurement: evaluate
 * every ReprojectionError block at the given parameters (reprojection_error.h:49-114
 * through create_reprojection_error_cost_function.h:54-136).  Outputs (host or device
 * as problem->memory_space; any may be NULL), in the caller's observation order:
 *   residuals [num_obs*2]
 *   jac_cam   [num_obs*2*6]   d r / d cam_ext           (row-major 2x6, ambient)
 *   jac_intr  [num_obs*2*THB_INTR_STRIDE] d r / d intr  (row-major 2xSTRIDE, zero padded)
 *   jac_pt    [num_obs*2*4]   d r / d point             (row-major 2x4, ambient)
 *   ok        [num_obs]       1 if the functor returned true
 * No loss function and no manifold are applied here (those are ceres-side).
 */
int thb_ba_evaluate(const ThbBaProblem* problem, double* residuals, double* jac_cam,
                    double* jac_intr, double* jac_pt, uint8_t* ok, void* cuda_stream);

/*
 * Time the production K1 kernel (tangent-space residual+Jacobian into the solver's
 * plane layout) `repeats` times on the session's stream with CUDA events; returns the
 * average milliseconds per launch.  flush_l2 != 0 rewrites a >L2-sized buffer between
 * launches.  Used only for roofline reporting.
 */
int thb_ba_time_jacobian(ThbBaSession* session, int32_t repeats, int32_t flush_l2,
                         double* avg_ms);

/*
 * Kernel-level entry for the parity tests of K4: solve the SPD system A x = b with the same dense
 * FP64 Cholesky the BA path uses for the reduced camera system (the reference delegates this to
 * Ceres' Schur solvers, bundle_adjuster.cc:65-88). A: n x n row-major (lower triangle read), host
 * pointers. Returns THB_E_NUMERICAL if A is not positive definite.
 */
int thb_dense_spd_solve(const double* A, const double* b, int32_t n, double* x, void* cuda_stream);

/*
 * Measurement entry for K4: factor + solve a synthetic diagonally dominant SPD system of order n that is built
 * on the device, `repeats` times; returns the average milliseconds of FactorAndSolve (CUDA events on the
 * stream; the matrix is rebuilt outside the timed region) and the max-norm residual |A x - b| of the last
 * solve relative to |b|. Used only for roofline reporting and the ncu launch lists.
 */
int thb_dense_spd_time(int32_t n, int32_t repeats, double* avg_ms, double* rel_residual, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------
 * RANSAC two-view geometric verification (hot path 2).
 * ---------------------------------------------------------------------------------------------- */

/* RansacParameters (solvers/sample_consensus_estimator.h:58-126). The reference's `rng` member is a
 * process-global generator that Python cannot set; here every pair carries its own seed and its
 * RANSAC run is what the reference does right after `RandomNumberGenerator rng(seed)` (util/random.cc:54-62). */
typedef struct ThbRansacParams {
  double error_thresh;        /* on SQUARED errors; must be > 0 (sample_consensus_estimator.h:217)   */
  double failure_probability; /* in (0, 1)                                                          */
  double min_inlier_ratio;    /* in [0, 1]                                                          */
  int32_t min_iterations;
  int32_t max_iterations;
  int32_t use_mle;            /* MLEQualityMeasurement instead of InlierSupport                     */
  int32_t use_lo;             /* LO-RANSAC: Estimator::RefineModel on every improved model from lo_start_iterations on and */
                              /* once on the final inliers (sample_consensus_estimator.h:372-380, 400-405). Relative pose  */
                              /* only (BundleAdjustTwoViewsAngular); THB_E_UNSUPPORTED for absolute pose, a no-op upstream  */
                              /* for homographies (Estimator::RefineModel default)                                          */
  int32_t lo_start_iterations;
  int32_t ransac_type;        /* RansacType: RANSAC (0), PROSAC (1, data sorted by quality; solvers/prosac_sampler.cc) or LMED (2, solvers/lmed.h + lmed_quality_measurement.h: cost = median of the squared residuals, not with use_lo); EXHAUSTIVE (3) is rejected: its sampler CHECK-aborts for sample sizes other than 2 in the reference too */
  int32_t use_tdd_test;       /* RansacParameters::use_Tdd_test: ComputeMaxIterations counts SampleSize + 1 draws */
                              /* (sample_consensus_estimator.h:272-279); the test itself is unimplemented upstream */
  int32_t reserved0;
} ThbRansacParams;

/* A batch of image pairs: pair p owns correspondences [pair_offset[p], pair_offset[p+1]). */
typedef struct ThbPairBatch {
  int32_t num_pairs;
  int32_t memory_space;       /* THB_MEM_HOST or THB_MEM_DEVICE for every pointer                   */
  const int64_t* pair_offset; /* [num_pairs + 1]                                                    */
  const double* corr;         /* [total * 4] normalised x1, y1, x2, y2 (FeatureCorrespondence::feature{1,2}.point_, */
                              /*             matching/feature_correspondence.h:49-72)               */
  const uint32_t* seed;       /* [num_pairs]                                                        */
} ThbPairBatch;

/* RelativePose (sfm/estimators/estimate_relative_pose.h:49-53) + RansacSummary
 * (sample_consensus_estimator.h:129-144); matrices row-major. */
typedef struct ThbRelPoseResult {
  int32_t success;            /* return value of EstimateRelativePose                               */
  int32_t num_inliers;
  int32_t num_iterations;
  int32_t num_input_data_points;
  double confidence;
  double best_cost;
  double essential_matrix[9];
  double rotation[9];
  double position[3];
  int32_t num_lo_iterations;  /* RansacSummary::num_lo_iterations */
  int32_t reserved0;
} ThbRelPoseResult;

void thb_ransac_default_params(ThbRansacParams* params);

/*
 * theia::EstimateRelativePose (sfm/estimators/estimate_relative_pose.cc:159-172) for every pair of the batch:
 * SampleConsensusEstimator::Estimate (sample_consensus_estimator.h:299-415) with RandomSampler, the five-point
 * solver, cheirality-gated Sampson scoring and the adaptive iteration bound, replayed exactly in iteration
 * order. results: [num_pairs]; inlier_mask: [total] (1 = RansacSummary::inliers contains the index), may be
 * NULL. Both in the batch's memory space.
 */
int thb_ransac_relpose_batch(const ThbPairBatch* batch, const ThbRansacParams* params,
                             ThbRelPoseResult* results, uint8_t* inlier_mask, void* cuda_stream);

/*
 * theia::EstimateCalibratedAbsolutePose (sfm/estimators/estimate_calibrated_absolute_pose.cc:176-190) with
 * PnPType::KNEIP (the estimator's default, :66): sample size 3, PoseFromThreePoints, squared normalised reprojection
 * error (:158-167). batch->corr is [total * 5] = FeatureCorrespondence2D3D {feature (x, y), world_point (X, Y, Z)}
 * (feature_correspondence_2d_3d.h:42-49). Results use ThbRelPoseResult with rotation / position =
 * CalibratedAbsolutePose::{rotation, position}; essential_matrix is zero. DLS / SQPnP are not replayable.
 */
int thb_ransac_abspose_batch(const ThbPairBatch* batch, const ThbRansacParams* params,
                             ThbRelPoseResult* results, uint8_t* inlier_mask, void* cuda_stream);

/*
 * theia::EstimateHomography (sfm/estimators/estimate_homography.cc:122-136): sample size 4, FourPointHomography,
 * one-way transfer error (:104-111). Results use ThbRelPoseResult with essential_matrix = the homography (row-major);
 * rotation / position are zero. This is also TwoViewMatchGeometricVerification::CountHomographyInliers
 * (two_view_match_geometric_verification.cc:331-366): num_inliers.
 */
int thb_ransac_homography_batch(const ThbPairBatch* batch, const ThbRansacParams* params,
                                ThbRelPoseResult* results, uint8_t* inlier_mask, void* cuda_stream);

/*
 * Work counters of the last thb_ransac_*_batch call made by the calling host thread, accumulated on the device
 * (one atomic add per pair): what the FP64 roofline of the RANSAC kernel is computed from.
 */
typedef struct ThbRansacStats {
  uint64_t pairs;
  uint64_t iterations;      /* RANSAC iterations consumed (sum of RansacSummary::num_iterations)                  */
  uint64_t samples_solved;  /* minimal problems solved (a batch solves up to 128 per pair before the replay stops) */
  uint64_t models_scored;   /* candidate models that entered scoring                                              */
  uint64_t data_scored;     /* per-datum error evaluations actually executed (early abandonment included)         */
  uint64_t reserved0;
  uint64_t cycles_draw;     /* where the time goes, per phase. Round-synchronous path: device time of the phase kernels in  */
  uint64_t cycles_solve;    /* nanoseconds; fused (LO) kernel: SM clock cycles the CTAs spent in the phase, summed over CTAs */
  uint64_t cycles_score;
  uint64_t cycles_scan;
} ThbRansacStats;
int thb_ransac_last_stats(ThbRansacStats* stats);

/*
 * Bit-packs per-correspondence inlier masks for transport (the pair queue's all-gather): pair p's n_p flags
 * mask[pair_offset[p] ...] go to words[word_offset[p] ...], bit i of word i/32 = flag i, ceil(n_p / 32) words per pair
 * (word_offset[p+1] - word_offset[p] must be at least that). All pointers in `memory_space`.
 */
int thb_pack_inlier_masks(const uint8_t* mask, const int64_t* pair_offset, const int64_t* word_offset, int32_t num_pairs,
                          int32_t memory_space, uint32_t* words, void* cuda_stream);

/*
 * Measurement entry: the FP64 FMA issue rate of the device (independent DFMA chains on every SM, best of `repeats`
 * launches, CUDA events), in TFLOP/s. The denominator of the FP64 rooflines bench.py reports (MEASURED_PEAKS.json
 * carries no FP64 figure).
 */
int thb_fp64_peak_tflops(int32_t repeats, double* tflops, void* cuda_stream);


/* ------------------------------------------------------------------------------------------------
 * Two-view geometry of image pairs from PIXEL correspondences (what the pipelines call once per pair).
 * ---------------------------------------------------------------------------------------------- */

/* A view's intrinsics as Camera::SetFromCameraIntrinsicsPriors leaves them (camera.cc), per-model parameter layout. */
typedef struct ThbViewIntrinsics {
  int32_t model;               /* THB_MODEL_*                                                           */
  int32_t image_width;         /* CameraIntrinsicsPrior::image_width / image_height (0 = unknown)        */
  int32_t image_height;
  int32_t focal_length_is_set; /* CameraIntrinsicsPrior::focal_length.is_set: both views set = calibrated */
  double params[THB_INTR_STRIDE];
} ThbViewIntrinsics;

/* EstimateTwoViewInfoOptions (sfm/estimate_twoview_info.h:51-80) + TwoViewMatchGeometricVerification::Options
 * (sfm/two_view_match_geometric_verification.h:53-92). Fill with thb_two_view_default_options. */
typedef struct ThbTwoViewOptions {
  double max_sampson_error_pixels;   /* 6.0, relative to a 1024-pixel image (reconstruction_estimator_utils.cc:97-110) */
  double expected_ransac_confidence; /* 0.9999 */
  int32_t min_ransac_iterations;     /* 10   */
  int32_t max_ransac_iterations;     /* 1000 */
  int32_t use_mle;                   /* 1    */
  int32_t use_lo;                    /* 0    */
  int32_t lo_start_iterations;       /* 10   */
  int32_t ransac_type;               /* RANSAC (0) only */
  /* verification only */
  int32_t min_num_inlier_matches;    /* 30 */
  int32_t bundle_adjustment;         /* 1  */
  double triangulation_max_reprojection_error; /* 15 px */
  double min_triangulation_angle_degrees;      /* 4     */
  double final_max_reprojection_error;         /* 5 px  */
} ThbTwoViewOptions;

/* TwoViewInfo (sfm/twoview_info.h:52-80) + the return value of the call. */
typedef struct ThbTwoViewInfo {
  int32_t success;                /* return value of EstimateTwoViewInfo / VerifyMatches               */
  int32_t num_verified_matches;
  int32_t num_homography_inliers; /* VerifyMatches only */
  int32_t visibility_score;       /* the reference computes it from the inlier list BEFORE filling it: always 0 (estimate_twoview_info.cc:186-189) */
  int32_t num_ransac_iterations;  /* RansacSummary::num_iterations of the relative-pose RANSAC          */
  int32_t num_triangulated;       /* VerifyMatches: matches that survived TriangulatePoints             */
  int32_t ba_iterations;          /* VerifyMatches: trust-region iterations of BundleAdjustTwoViews     */
  int32_t reserved0;
  double focal_length_1, focal_length_2;
  double position_2[3];
  double rotation_2[3];           /* angle-axis */
  double ba_initial_cost, ba_final_cost;
} ThbTwoViewInfo;

void thb_two_view_default_options(ThbTwoViewOptions* options);

/*
 * theia::EstimateTwoViewInfo (sfm/estimate_twoview_info.cc:259-305), calibrated branch (:133-192), for every pair of the
 * batch in one call: batch->corr holds PIXEL correspondences (x1, y1, x2, y2); they are normalised with each view's inverse
 * camera model (NormalizeFeatures :67-103 -> Camera::PixelToNormalizedCoordinates, all six models), the per-pair Sampson
 * threshold is scaled1 * scaled2 / (f1 * f2) (:155-167), then EstimateRelativePose runs with the pair's seed.
 * intrinsics1 / intrinsics2: [num_pairs]; info: [num_pairs]; inlier_mask: [total] (TwoViewInfo's inlier_indices), may be
 * NULL. Pairs whose views are not both calibrated return THB_E_UNSUPPORTED (the uncalibrated branch, :194-255, needs the
 * eight-point estimator). All pointers in batch->memory_space.
 */
int thb_estimate_two_view_info_batch(const ThbPairBatch* batch, const ThbViewIntrinsics* intrinsics1,
                                     const ThbViewIntrinsics* intrinsics2, const ThbTwoViewOptions* options,
                                     ThbTwoViewInfo* info, uint8_t* inlier_mask, void* cuda_stream);

/*
 * theia::TwoViewMatchGeometricVerification::VerifyMatches (sfm/two_view_match_geometric_verification.cc:114-183) for every
 * pair of the batch: CountHomographyInliers (:331-366, EstimateHomography on the pixel matches; its threshold uses the
 * not-yet-set-up cameras, i.e. max_sampson_error_pixels^2), EstimateTwoViewInfo on the SAME random generator, the
 * min_num_inlier_matches gates, and - options->bundle_adjustment - BundleAdjustRelativePose (:259-327): TriangulatePoints
 * (angle test, TriangulateMidpoint, 15 px reprojection gate), BundleAdjustTwoViews (camera 1 fixed, camera 2 extrinsics and
 * the 4-vector points free, no loss, DENSE_SCHUR; sfm/bundle_adjustment/bundle_adjust_two_views.cc:112-193) and the 5 px
 * reprojection filter. verified_mask [total]: 1 = the match is in verified_matches. Guided matching needs descriptors and
 * is not part of this entry. Calibrated pairs only.
 */
int thb_verify_two_view_matches_batch(const ThbPairBatch* batch, const ThbViewIntrinsics* intrinsics1,
                                      const ThbViewIntrinsics* intrinsics2, const ThbTwoViewOptions* options,
                                      ThbTwoViewInfo* info, uint8_t* verified_mask, void* cuda_stream);

/* theia::PoseFromThreePoints (sfm/pose/perspective_three_point.cc:182-291) for `count` independent samples (host
 * pointers): features [count*3*2], world_points [count*3*3]; R_out [count*4*9] row-major, t_out [count*4*3],
 * num_solutions [count] (0 => collinear world points, the reference returns false). */
int thb_p3p(const double* features, const double* world_points, int32_t count, double* R_out, double* t_out,
            int32_t* num_solutions, void* cuda_stream);

/* theia::FourPointHomography (sfm/pose/four_point_homography.cc:72-102), minimal case: corr [count*4*4] (x1,y1,x2,y2),
 * H_out [count*9] row-major, ok [count]. */
int thb_four_point_homography(const double* corr, int32_t count, double* H_out, int32_t* ok, void* cuda_stream);

/* theia::SevenPointFundamentalMatrix (sfm/pose/seven_point_fundamental_matrix.cc:72-152): corr [count*7*4],
 * F_out [count*3*9] row-major, num_solutions [count]. Reproduces the reference as written, including its
 * cubic-coefficient order. */
int thb_seven_point_fundamental_matrix(const double* corr, int32_t count, double* F_out, int32_t* num_solutions,
                                       void* cuda_stream);

/* Per-track record of thb_ba_tracks_batch: what BundleAdjustTrack's BundleAdjustmentSummary carries for one track. */
typedef struct ThbTrackBaResult {
  double initial_cost;
  double final_cost;
  int32_t num_iterations;   /* -1: track not optimised (constant or without observations) */
  int32_t termination_type; /* THB_TERM_*; THB_TERM_FAILURE leaves the point untouched (summary.success == false) */
} ThbTrackBaResult;

/*
 * theia::BundleAdjustTrack (sfm/bundle_adjustment/bundle_adjustment.cc:261-285) for EVERY non-constant point of the problem
 * in one launch: each track is its own trust-region problem over its 3 (homogeneous parametrisation) or 4 coordinates with
 * all cameras and intrinsics constant (bundle_adjuster.cc:176-221) - what TrackEstimator::EstimateTrack does once per track
 * from a thread pool (sfm/estimate_track.cc:286-290). problem->cam_const / intr_const are ignored (everything but the
 * points is constant), problem->pt_const selects the tracks, options as for thb_ba_solve (use_inner_iterations must be 0,
 * as BundleAdjustTrack forces, bundle_adjustment.cc:267). Refines problem->pts in place; results [num_points] may be NULL.
 * Any memory space (results in problem->memory_space); the observations are grouped by track on the device.
 */
int thb_ba_tracks_batch(const ThbBaProblem* problem, const ThbBaOptions* options, ThbTrackBaResult* results, void* cuda_stream);

/* TrackEstimator::Options (sfm/estimate_track.h:59-84), MIDPOINT triangulation (the default). */
typedef struct ThbTrackEstimatorOptions {
  double max_acceptable_reprojection_error_pixels; /* 5.0 */
  double min_triangulation_angle_degrees;          /* 3.0 */
  int32_t bundle_adjustment;                       /* 1   */
  int32_t reserved0;
} ThbTrackEstimatorOptions;

#define THB_TRACK_SKIPPED (-1)             /* constant (already estimated) track: TrackEstimator::EstimateTrack returns early */
#define THB_TRACK_ESTIMATED 0
#define THB_TRACK_BAD_ANGLE 1              /* < 2 observations or no pair of rays wider than the minimum angle (:230-236) */
#define THB_TRACK_FAILED_TRIANGULATION 2   /* TriangulateMidpoint returned false (:256-261)                               */
#define THB_TRACK_BA_FAILED 3              /* BundleAdjustTrack summary.success == false (:295-297)                       */
#define THB_TRACK_BAD_REPROJECTION 4       /* a view sees the point behind it, or mean squared error too large (:300-311) */

/*
 * TrackEstimator::EstimateTrack (sfm/estimate_track.cc:209-321) for every non-constant point of the problem in one launch:
 * angle test (SufficientTriangulationAngle, triangulation.cc:236-250), TriangulateMidpoint from the camera positions and
 * the caller's unit ray directions (ray_directions [num_observations*3], one per observation, what
 * Camera::PixelToUnitDepthRay(feature).normalized() gives, estimate_track.cc:80-87), BundleAdjustTrack, and the
 * reprojection test (AcceptableReprojectionError, :93-119). status [num_points] receives THB_TRACK_*; problem->pts is
 * written for THB_TRACK_ESTIMATED tracks only (the caller sets those estimated); ba_results may be NULL. The starting
 * values in problem->pts are ignored. Any memory space (ray_directions, status and ba_results in problem->memory_space).
 */
int thb_estimate_tracks_batch(const ThbBaProblem* problem, const double* ray_directions, const ThbTrackEstimatorOptions* options,
                              const ThbBaOptions* ba_options, int32_t* status, ThbTrackBaResult* ba_results, void* cuda_stream);

/*
 * The covariance blocks behind BundleAdjustView(s) / BundleAdjustTrack(s) WITH covariance (bundle_adjustment.cc:287-380,
 * 419-500 -> BundleAdjuster::GetCovarianceFor{View,Views,Track,Tracks}, bundle_adjuster.cc:660-773 = ceres::Covariance with
 * default options: the loss function applied, tangent space): evaluated at the parameters in `problem`, no solve.
 *   - every point constant (the AddView problems): cam_cov [num_cameras*36] receives (J_c^T J_c)^-1 of each camera's extrinsics
 *     block, row-major 6 x 6 over [position(3), angle-axis(3)]; rows / columns of constant coordinates are zero; cam_ok[c] = 0
 *     for constant or unobserved cameras and rank-deficient blocks (ceres::Covariance::Compute returns false for those);
 *   - every camera and intrinsics block constant (the AddTrack problems, which the reference forces onto the homogeneous
 *     parametrisation: options->use_homogeneous_point_parametrization must be set): pt_cov [num_points*9] receives the 3 x 3
 *     tangent-space covariance of each point on SphereManifold<4>, pt_ok likewise.
 * The caller scales by the empirical variance factor 2 * final_cost / redundancy (bundle_adjustment.cc:311-316). Problems with
 * both free cameras and free points (or refined intrinsics) couple every block: THB_E_UNSUPPORTED. Output pointers of the
 * unused kind may be NULL. Any memory space.
 */
int thb_ba_covariance(const ThbBaProblem* problem, const ThbBaOptions* options, double* cam_cov, uint8_t* cam_ok, double* pt_cov,
                      uint8_t* pt_ok, void* cuda_stream);

#define THB_OUTLIER_KEPT 0
#define THB_OUTLIER_BAD_REPROJECTION 1   /* a view sees the point behind it, or the mean squared reprojection error is too large */
#define THB_OUTLIER_BAD_ANGLE 2          /* no two viewing rays are at least min_triangulation_angle_degrees apart               */
#define THB_OUTLIER_SKIPPED (-1)         /* pt_const[p] != 0: the caller's "not estimated" tracks are not looked at              */

/*
 * theia::SetOutlierTracksToUnestimated (sfm/set_outlier_tracks_to_unestimated.cc:62-137), the filter every pipeline runs
 * after bundle adjustment (global_reconstruction_estimator.cc:266-270), for all tracks in one launch. The problem holds
 * the observations of the ESTIMATED views only (the reference skips the others, :86-88); points with pt_const != 0 are
 * skipped (tracks that are not estimated, :76-78). status [num_points] receives THB_OUTLIER_*; the caller sets every track
 * with status > 0 to unestimated. *num_removed = the function's return value. Any memory space.
 */
int thb_set_outlier_tracks_batch(const ThbBaProblem* problem, double max_inlier_reprojection_error,
                                 double min_triangulation_angle_degrees, int32_t* status, int32_t* num_removed, void* cuda_stream);

/*
 * theia::SelectGoodTracksForBundleAdjustment (sfm/select_good_tracks_for_bundle_adjustment.cc:263-325): the subset of tracks a
 * large bundle adjustment optimises. The problem holds the estimated views and estimated tracks with their observations
 * (the reference skips the others, :94-96,132-136,174-176). Per track: truncated length min(#views, long_track_length_threshold)
 * and mean squared reprojection error over all its views (:80-107). Stage 1 (:152-195): in every selected view the image is cut
 * into grid_cell_size_pixels cells ((int)(x / size), (int)(y / size), truncation as the reference's cast) and the track with the
 * smallest (truncated length, mean error) pair of each cell is chosen - std::pair's operator<, as the reference's
 * CompareGridCellElements. Stage 2 (:199-254), view after view: a view with fewer than min_num_optimized_tracks_per_view
 * chosen tracks gets its not-yet-chosen tracks in ascending track order until it has enough (std::partial_sort on
 * pair<TrackId, statistics> orders by TrackId first). Where the reference's order is unspecified (unordered containers: the
 * order of the views in stage 2, ties between equal statistics in a cell) this entry uses ascending index order: views and
 * tracks of the problem must be in ascending ViewId / TrackId order to reproduce a run of the reference that iterates
 * that way.
 * cam_selected [num_cameras]: the view_ids subset of the second overload (NULL = every view); selected [num_points]: in / out,
 * the caller's tracks_to_optimize (non-zero = chosen; normally zero on entry). *num_selected may be NULL. Any memory space.
 */
int thb_select_good_tracks_batch(const ThbBaProblem* problem, const uint8_t* cam_selected, int32_t long_track_length_threshold,
                                 int32_t grid_cell_size_pixels, int32_t min_num_optimized_tracks_per_view, uint8_t* selected,
                                 int32_t* num_selected, void* cuda_stream);

/*
 * theia::TriangulateMidpoint (sfm/triangulation/triangulation.cc:130-157) for a batch of tracks: track t owns the rays
 * ray_offset[t] .. ray_offset[t+1]-1 (origin and direction, 3 doubles each; directions as the caller passes them, the
 * reference does not normalise). A = sum (I4 - d d^T) with d = (direction, 0), b = sum (I4 - d d^T) (origin, 1), 4x4
 * Cholesky solve; ok[t] = 0 when the factorisation fails (Eigen::LLT info != Success) or the track has fewer than two
 * rays (the reference CHECK-aborts there). points_out [num_tracks*4]. This is the per-track kernel of
 * TrackEstimator::EstimateAllTracks (estimate_track.cc:124-321), one launch for all tracks. memory_space as in ThbBaProblem;
 * host offsets are validated (THB_E_INVALID_ARGUMENT), device offsets are trusted.
 */
int thb_triangulate_midpoint_batch(const double* ray_origins, const double* ray_directions, const int64_t* ray_offset,
                                   int32_t num_tracks, int32_t memory_space, double* points_out, uint8_t* ok,
                                   void* cuda_stream);

/*
 * theia::FivePointRelativePose (sfm/pose/five_point_relative_pose.cc:212-293), minimal case, for `count`
 * independent 5-point samples (host pointers): x1, x2 [count*5*2]; E_out [count*10*9] row-major, in the
 * reference's solution order; num_solutions [count] (0 => the reference returns false).
 */
int thb_five_point_relative_pose(const double* x1, const double* x2, int32_t count, double* E_out,
                                 int32_t* num_solutions, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* THEIA_B200_H_ */
