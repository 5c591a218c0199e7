/*
 * clic_cuda.h — C ABI of the B200-native continuous-time fixed-lag smoothing core.
 *
 * This is the drop-in boundary for ONE path of APRIL-ZJU/clic: B-spline factor residual +
 * analytic-Jacobian evaluation, normal-equation assembly, the LM/GN step and the
 * marginalization Schur complement.  The reference has no FFI layer: its
 * boundary is the C++ class clic::TrajectoryEstimator (src/estimator/trajectory_estimator.h:100-278)
 * on top of ceres::Problem.  Every entry point below cites the reference method it replaces.
 * include/clic/trajectory_estimator.hpp is a header-only C++ shim that reproduces the
 * reference method names / argument order on top of this ABI.
 *
 * Two libraries implement this header with identical symbols:
 *   clic_b200/lib/libclic_cuda.so   the product: hand-written sm_100a CUDA (no CPU path at all)
 *   oracle/_build/libclic_oracle.so test infrastructure only: a scalar CPU restatement of the
 *                                   reference arithmetic (see oracle/README.md)
 *
 * Conventions (SURVEY.md §8 "Conventions shared by all rows"):
 *   - rotations are unit quaternions stored (x,y,z,w)                 (sophus_lib/so3.hpp:153-160)
 *   - tangent update R <- R*exp(dtheta), p <- p + dp                    (factor/ceres_local_param.h:132-139)
 *   - all host pointers are borrowed for the duration of the call only
 *   - every function returns a clic_status; nothing throws or asserts across the ABI
 *   - time stamps are double seconds exactly as the reference structs hold them
 *     (lidar_feature.h:110-123 t_point, parameter_struct.h:52-58 timestamp); the library converts
 *     with the reference's truncation time_ns = (int64)(t * 1e9) (trajectory_estimator.cpp:275)
 *   - a measurement outside the window is silently dropped, exactly like
 *     TrajectoryEstimator::MeasuredTimeToNs returning false (trajectory_estimator.cpp:272-292,717);
 *     the number dropped is reported through an out-parameter
 *
 * Canonical tangent ordering of H and b (the reference never materialises one):
 *   [ knot k_f : dtheta(3) dp(3) | knot k_f+1 ... | knot K-1 ]   k_f = fixed_index + 1
 *   [ gyro bias slot 0..nb-1 (3 each) ]   if !lock_wb
 *   [ accel bias slot 0..nb-1 (3 each) ]  if !lock_ab
 *   [ t_offset IMU (1) ] if unlocked, [ t_offset camera (1) ] if unlocked
 *   [ inverse depth of landmark 0..L-1 ]  only landmarks added with fixed_depth = 0
 * H = sum J~^T J~, b = sum J~^T r~ with loss-corrected J~, r~ (marginalization_factor.cpp:38-61,138,144).
 */
#ifndef CLIC_CUDA_H_
#define CLIC_CUDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLIC_API __attribute__((visibility("default")))
#else
#define CLIC_API
#endif

typedef enum {
  CLIC_OK = 0,
  CLIC_ERR_BAD_HANDLE = 1,
  CLIC_ERR_BAD_ARGUMENT = 2,
  CLIC_ERR_CUDA = 3,          /* a CUDA runtime call or kernel failed; see clic_last_error() */
  CLIC_ERR_NO_DEVICE = 4,     /* no sm_100 device: the product path refuses to run (no CPU fallback) */
  CLIC_ERR_NCCL = 5,
  CLIC_ERR_NOT_POSITIVE_DEFINITE = 6,
  CLIC_ERR_UNSUPPORTED = 7,
  CLIC_ERR_NOTHING_TO_KEEP = 8 /* marginalization leaves no parameters (SaveMarginalizationInfo -> nullptr, trajectory_estimator.cpp:241-248) */
} clic_status;

/* SensorType, src/spline/trajectory.h:29-33 */
enum { CLIC_SENSOR_IMU = 0, CLIC_SENSOR_LIDAR = 1, CLIC_SENSOR_CAMERA = 2 };
/* GeometryType, src/lidar_odometry/lidar_feature.h:108 */
enum { CLIC_GEO_LINE = 0, CLIC_GEO_PLANE = 1 };

/* Parameter-block kinds a prior can refer to (the reference keys them by raw address,
 * marginalization_factor.h:125-138; we key them by kind + stable id). */
enum {
  CLIC_BLOCK_KNOT_ROT = 0,  /* id = global knot index, size 4, local 3 */
  CLIC_BLOCK_KNOT_POS = 1,  /* id = global knot index, size 3 */
  CLIC_BLOCK_BIAS_GYRO = 2, /* id = global bias slot, size 3 */
  CLIC_BLOCK_BIAS_ACCEL = 3,
  CLIC_BLOCK_GRAVITY = 4,   /* id = 0, size 3 (ambient; only ever appears inside priors) */
  CLIC_BLOCK_T_OFFSET = 5,  /* id = sensor, size 1 */
  CLIC_BLOCK_INV_DEPTH = 6  /* id = landmark index, size 1 */
};

typedef struct clic_estimator clic_estimator;
typedef struct clic_prior clic_prior;

/* The trajectory window + the non-spline states one solve touches.
 * Replaces: Trajectory (knot deques, src/spline/se3_spline.h; extrinsics EP_StoI_, trajectory.h:59-67),
 * TrajectoryManager::all_imu_bias_ / gravity_ (trajectory_manager.cpp:599-605) and
 * fea_id_inv_depths_ (:707), which the reference passes as raw double* parameter blocks. */
typedef struct {
  int64_t t0_ns;          /* time of the first uploaded knot's interval start = minTimeNs + knot_id0*dt_ns */
  int64_t dt_ns;          /* knot spacing, (int64)(knot_distance*1e9) (trajectory.h:46-48) */
  int32_t n_knots;        /* K >= 4 */
  int32_t knot_id0;       /* global index of uploaded knot 0 (priors store global ids) */
  const double* knot_q;   /* K*4, (x,y,z,w) */
  const double* knot_p;   /* K*3 */
  double S_LtoI[4], p_LinI[3]; /* unused by the adders (they take extrinsics explicitly like the reference) but kept for pose queries */
  double S_CtoI[4], p_CinI[3]; /* ImageFeatureFactor::SetParam (image_feature_factor.h:270-284) */
  double t_offset_ns[3];  /* per CLIC_SENSOR_*; ExtrinsicParam::t_offset_ns is a double (parameter_struct.h:108-150) */
  double gravity[3];
  int32_t n_bias;         /* number of bias slots (the reference uses 2 per solve: last scan, current scan) */
  int32_t bias_id0;       /* global id of slot 0 */
  const double* bias;     /* n_bias*6: gyro(3) accel(3) */
  int32_t n_landmarks;
  int32_t reserved0;
  const double* inv_depth; /* n_landmarks */
} clic_window_desc;

/* Mirrors TrajectoryEstimatorOptions (trajectory_estimator_options.h:32-72) + execution knobs. */
typedef struct {
  int32_t lock_traj;
  int32_t lock_ab, lock_wb, lock_g;     /* lock_g must be 1 for clic_solve (gravity uses a Ceres-internal
                                           HomogeneousVectorParameterization; locked in every shipped call,
                                           trajectory_manager.cpp:244,615) */
  int32_t lock_t_offset[3];             /* per sensor */
  int32_t is_marg_state;
  int64_t t_offset_padding_ns;
  int32_t ctrl_to_be_opt_now, ctrl_to_be_opt_later; /* global knot ids */
  int32_t marg_bias_param, marg_gravity_param, marg_t_offset_param;
  /* execution */
  int32_t device;        /* CUDA device ordinal */
  void* stream;          /* cudaStream_t to run on; NULL = library-owned stream */
  int32_t deterministic; /* 1 (default): fixed reduction order */
  int32_t reserved0;
  /* ImageFeatureFactor::sqrt_info (scalar x I2).  The class default is 450/1.5 (image_feature_factor.h:277-278), but
   * every run overwrites it with the configured `image_weight` through TrajectoryManager::InitFactorInfo
   * (trajectory_manager.cpp:53-61; config/ct_odometry_ntu.yaml:35 = 200, ct_odometry_lvi.yaml:35 = 450).
   * clic_default_options sets the class default; <= 0 also means the class default. */
  double image_sqrt_info;
  int32_t reserved[4];
} clic_options;

/* POD replacement of ceres::Solver::Summary (trajectory_estimator.h:225-226). */
enum { CLIC_TERM_CONVERGENCE = 0, CLIC_TERM_NO_CONVERGENCE = 1, CLIC_TERM_FAILURE = 2 };
enum {
  CLIC_REASON_MAX_ITERATIONS = 0, CLIC_REASON_GRADIENT_TOL = 1, CLIC_REASON_PARAMETER_TOL = 2,
  CLIC_REASON_FUNCTION_TOL = 3, CLIC_REASON_MIN_RADIUS = 4, CLIC_REASON_INVALID_STEPS = 5
};
typedef struct {
  int32_t iterations;            /* LM iterations after iteration 0 (successful + unsuccessful) */
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination;
  int32_t termination_reason;
  int32_t num_jacobian_evals;
  int32_t num_residual_evals;
  int32_t reserved;
  double initial_cost, final_cost, fixed_cost;
  double final_radius, final_gradient_max_norm;
} clic_summary;

/* Where each parameter block sits in H/b (-1 = constant / absent). */
typedef struct {
  int32_t D;                 /* number of columns */
  int32_t first_free_knot;   /* local knot index k_f; columns 6*(k-k_f) .. +5 = dtheta, dp */
  int32_t n_free_knots;
  int32_t col_bias_gyro;     /* + 3*slot */
  int32_t col_bias_accel;    /* + 3*slot */
  int32_t col_t_offset[3];
  int32_t col_inv_depth;     /* + free-landmark rank */
  int32_t n_free_landmarks;
} clic_layout;

CLIC_API const char* clic_last_error(void);
CLIC_API const char* clic_backend(void); /* "cuda-sm100a" or "oracle-cpu" */
CLIC_API void clic_default_options(clic_options* opt);

/* TrajectoryEstimator::TrajectoryEstimator (trajectory_estimator.cpp:142-170). */
CLIC_API int clic_create(const clic_window_desc* window, const clic_options* opt, clic_estimator** out);
CLIC_API int clic_destroy(clic_estimator* h);

/* TrajectoryEstimator::SetFixedIndex (trajectory_estimator.h:140); idx is a GLOBAL knot id;
 * knots with id <= idx are constant (trajectory_estimator.cpp:191-193). */
CLIC_API int clic_set_fixed_index(clic_estimator* h, int32_t idx);

/* AddLoamMeasurementAnalytic (trajectory_estimator.cpp:712-750), batched over n PointCorrespondence
 * records given as one array per struct field (lidar_feature.h:110-123).  geo_plane / geo_normal /
 * geo_point may each be NULL when no record of that type is present.  No loss (":743").
 * marg_this_factor routes to the marginalization set with the |r|/w < 0.05 gate (":733-741"). */
CLIC_API int clic_add_loam(clic_estimator* h, int64_t n, const double* t_point, const double* point /*3n*/,
                           const int32_t* geo_type, const double* geo_plane /*4n*/,
                           const double* geo_normal /*3n*/, const double* geo_point /*3n*/,
                           const double S_GtoM[4], const double p_GinM[3], const double S_LtoI[4],
                           const double p_LinI[3], double weight, int32_t marg_this_factor,
                           int64_t* n_dropped);

/* Same factors, packed input (81 instead of 116 bytes per record over PCIe): geo_type as one byte, and one
 * 6-double geometry record per feature — plane: (n.x, n.y, n.z, d, 0, 0) = geo_plane; line: (geo_point, geo_normal).
 * This is what include/clic/trajectory_estimator.hpp stages per AddLoamMeasurementAnalytic call. */
CLIC_API int clic_add_loam_packed(clic_estimator* h, int64_t n, const double* t_point, const double* point /*3n*/,
                                  const uint8_t* geo_type, const double* geo /*6n*/, const double S_GtoM[4],
                                  const double p_GinM[3], const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped);

/* Type-uniform batches of the same factors — what include/clic/trajectory_estimator.hpp stages: the shim routes each
 * AddLoamMeasurementAnalytic call to a plane batch or a line batch by pc.geo_type.  64 B per plane feature
 * (t 8 + point 24 + geo_plane 32) and 80 B per line feature (t 8 + point 24 + geo_point 24 + geo_normal 24): exactly
 * the information the factor reads, over PCIe and in HBM; and a type-uniform kernel (no intra-warp
 * divergence between the plane and the line residual).  Each batch must be time-sorted for full speed. */
CLIC_API int clic_add_loam_planes(clic_estimator* h, int64_t n, const double* t_point, const double* point /*3n*/,
                                  const double* plane /*4n: n.x n.y n.z d*/, const double S_GtoM[4],
                                  const double p_GinM[3], const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped);
CLIC_API int clic_add_loam_lines(clic_estimator* h, int64_t n, const double* t_point, const double* point /*3n*/,
                                 const double* line /*6n: geo_point xyz, geo_normal xyz*/, const double S_GtoM[4],
                                 const double p_GinM[3], const double S_LtoI[4], const double p_LinI[3], double weight,
                                 int32_t marg_this_factor, int64_t* n_dropped);

/* AddIMUMeasurementAnalytic (trajectory_estimator.cpp:369-455), batched over IMUData
 * (parameter_struct.h:52-58).  All samples of one call share a bias slot, like the reference's
 * call sites (trajectory_manager.cpp:731-735).  Loss CauchyLoss(marg ? 3 : 10) (":426-430"). */
CLIC_API int clic_add_imu(clic_estimator* h, int64_t n, const double* timestamp, const double* gyro /*3n*/,
                          const double* accel /*3n*/, int32_t bias_slot, const double info_vec[6],
                          int32_t marg_this_factor, int64_t* n_dropped);

/* AddBiasFactor (trajectory_estimator.cpp:457-498). */
CLIC_API int clic_add_bias(clic_estimator* h, int32_t slot_i, int32_t slot_j, double dt,
                           const double info_vec[6], int32_t marg_this_factor, int32_t marg_all_bias);

/* AddGravityFactor (trajectory_estimator.cpp:500-523) with GravityFactor (trajectory_value_factor.h:33-62):
 * r = diag(info_vec) (g - gravity), gravity = the value the caller passes (the reference passes the block's value at
 * add time, so r = 0 until the block moves).  In a solve the gravity block is constant (lock_g: every shipped call), so
 * the factor only adds its constant cost to fixed_cost; in a marginalization estimator with marg_this_factor it enters the
 * prior (J = diag(info_vec) on the 3 ambient gravity columns), the block dropped iff marg_gravity_param. */
CLIC_API int clic_add_gravity(clic_estimator* h, const double gravity[3], const double info_vec[3],
                              int32_t marg_this_factor);

/* AddGlobalVelocityMeasurement (trajectory_estimator.cpp:609-664) with auto_diff::IMUGlobalVelocityFactor
 * (factor/auto_diff/trajectory_value_factor.h:436-487): r = vel_weight * (p'(t + t_offset_IMU) - global_v) on the
 * position knots (+ the IMU time offset when it is free), CauchyLoss(60) (":642").  It is the one factor besides the
 * prior and the IMU factors that TrajectoryManager::InitTrajWithPropagation adds (trajectory_manager.cpp:233-300),
 * so with it the third solve of every update runs here too.  *added = 0 when the timestamp falls outside the window
 * (the reference returns false).  Not supported in a marginalization estimator (the reference never adds it there). */
CLIC_API int clic_add_global_velocity(clic_estimator* h, double timestamp, const double global_v[3], double vel_weight,
                                      int32_t* added);

/* ---- the remaining adders of TrajectoryEstimator: factors the shipped pipeline adds a handful of
 * per solve at most; they run on one CTA after the assembly (small_factors_kernel), not through the streaming kernels.
 * None of them is routed to the marginalization set by the reference's adders (no marg_this_factor argument), so a
 * marginalization estimator refuses them. ---- */
/* AddPoseMeasurementAnalytic (trajectory_estimator.cpp:326-368) with IMUPoseFactor (trajectory_value_factor.h:303-433):
 * r = diag(info_vec) [log(R(t') q_meas^-1); p(t') - p_meas], t' = t + (int64) t_offset_IMU; Jacobians on the 4+4 knots
 * of the segment and on the IMU time offset; no loss.  orientation: 4n (x,y,z,w). */
CLIC_API int clic_add_pose(clic_estimator* h, int64_t n, const double* timestamp, const double* position /*3n*/,
                           const double* orientation /*4n*/, const double info_vec[6], int64_t* n_dropped);
/* AddLocalVelocityMeasurementAnalytic (trajectory_estimator.cpp:666-708) with LocalVelocityFactor
 * (trajectory_value_factor.h:435-571): r = weight (R(t') v_local - p'(t')); no loss. */
CLIC_API int clic_add_local_velocity(clic_estimator* h, int64_t n, const double* timestamp, const double* local_v /*3n*/,
                                     double weight, int64_t* n_dropped);
/* AddRelativeRotationAnalytic (trajectory_estimator.cpp:586-607) with RelativeOrientationFactor
 * (trajectory_value_factor.h:966-1075): r = diag(info_vec) log(R(t_b)^T R(t_a) S_BtoA); rotation knots of both
 * segments; no loss.  *added = 0 when either time is outside the window. */
CLIC_API int clic_add_relative_rotation(clic_estimator* h, double ta, double tb, const double S_BtoA[4],
                                        const double info_vec[3], int32_t* added);
/* AddImageDepthAnalytic (trajectory_estimator.cpp:830-852) with ImageDepthFactor (image_feature_factor.h:658-707):
 * reprojection of a feature between two FIXED camera poses, the landmark's inverse depth the only parameter (it
 * becomes a free parameter of the solve), CauchyLoss(1), sqrt_info = clic_options::image_sqrt_info. */
CLIC_API int clic_add_image_depth(clic_estimator* h, const double p_i[3], const double p_j[3], const double S_CitoCj[4],
                                  const double p_CiinCj[3], int32_t landmark);

/* IntegrationBase (src/visual_odometry/integration_base.h:36-288) reduced to what PreIntegrationFactor reads.  The
 * integration itself (propagate / midPointIntegration over the IMU samples) stays with the caller like in the reference;
 * sqrt_info = LLT(covariance^-1).matrixL()^T, which the reference recomputes with Eigen inside every Evaluate
 * (trajectory_value_factor.h:744-748), is handed in (the C++ shim computes it from IntegrationBase::covariance). */
typedef struct {
  double sum_dt;
  double delta_p[3], delta_q[4] /* x,y,z,w */, delta_v[3];
  double linearized_ba[3], linearized_bg[3];
  double jacobian[225];  /* 15x15 row-major, state order P R V BA BG (integration_base.h:20-26) */
  double sqrt_info[225]; /* 15x15 row-major (upper triangular) */
  double gravity_norm;   /* the global GRAVITY_NORM (parameter_struct.cpp:23: -9.8); gravity = (0, 0, -GRAVITY_NORM) */
} clic_preintegration;

/* AddPreIntegrationAnalytic (trajectory_estimator.cpp:529-586) with PreIntegrationFactor
 * (trajectory_value_factor.h:618-964): 15 residuals [dp, dR, dv, dba, dbg] between the spline poses / velocities at
 * t_a < t_b and the integrated IMU deltas, on the 4+4 knots of both intervals and the gyro / accel bias of slots
 * slot_i (t_a) and slot_j (t_b); no loss.  ta_ns / tb_ns are the factor's constructor arguments (int64 ns on the
 * trajectory clock).  The reference's adder passes its double SECONDS there (":545", an implicit double -> int64
 * conversion; nothing in the shipped pipeline calls it) — the C++ shim passes the nanosecond values MeasuredTimeToNs
 * produced, see INTEGRATION.md "quirks".  *added = 0 when either time is outside the window.  Not routed to the
 * marginalization set (a marginalization estimator refuses it). */
CLIC_API int clic_add_preintegration(clic_estimator* h, int64_t ta_ns, int64_t tb_ns, const clic_preintegration* pre,
                                     int32_t slot_i, int32_t slot_j, int32_t* added);

/* ---- one analytic factor, residual and Jacobian: the surface src/app/test_analytic_factor.cpp drives
 * (cost_function->Evaluate + GetJacobian, ":846-900").  Covers every small analytic factor of SURVEY.md §8f-4,
 * including the five the reference never adds to a problem outside that test.  The knots are the estimator's current
 * window; the factor's other parameter blocks are given in x (the reference's block order after the knots). ---- */
enum {
  CLIC_FACTOR_POSE = 0,              /* IMUPoseFactor          t_a; d = position[3] orientation[4] info[6]; x = t_offset_ns */
  CLIC_FACTOR_LOCAL_VELOCITY = 1,    /* LocalVelocityFactor    t_a; d = local_v[3] weight; x = t_offset_ns */
  CLIC_FACTOR_RELATIVE_ROTATION = 2, /* RelativeOrientationFactor t_a, t_b; d = S_BtoA[4] info[3] */
  CLIC_FACTOR_IMAGE_DEPTH = 3,       /* ImageDepthFactor       d = p_i[3] p_j[3] S_CitoCj[4] p_CiinCj[3] sqrt_info; x = inv_depth */
  CLIC_FACTOR_PREINTEGRATION = 4,    /* PreIntegrationFactor   t_a, t_b; pre; x = bg_a[3] bg_b[3] ba_a[3] ba_b[3] */
  CLIC_FACTOR_IMAGE_3D2D = 5,        /* Image3D2DFactor (image_feature_factor.h:294-475)  t_a = t_j; d = p_j[3] sqrt_info;
                                        x = p_G[3] t_offset_ns */
  CLIC_FACTOR_IMAGE_ONE_POSE = 6,    /* ImageFeatureOnePoseFactor (":477-656")  t_a = t_j; d = p_i[3] S_IitoG[4] p_IiinG[3]
                                        p_j[3] sqrt_info; x = inv_depth */
  CLIC_FACTOR_EPIPOLAR = 7,          /* EpipolarFactor (":709-842")  t_a = t_i; d = x_i[3] x_k[3] S_GtoCk[4] p_CkinG[3] weight */
  CLIC_FACTOR_LOAM_OPT_MAP_POSE = 8, /* LoamFeatureOptMapPoseFactor (lidar_feature_factor.h:174-348)  t_a = t_point; geo_type;
                                        d = point[3] geo_plane[4] geo_normal[3] geo_point[3] weight S_LtoI[4] p_LinI[3];
                                        x = S_ImtoG[4] p_IminG[3] */
  CLIC_FACTOR_RELATIVE_LOAM = 9      /* RalativeLoamFeatureFactor (":350-553")  t_a = t_point, t_b = t_map; d as above */
};
typedef struct {
  int32_t kind;      /* CLIC_FACTOR_* */
  int32_t geo_type;  /* CLIC_GEO_* (the two LiDAR kinds) */
  int64_t t_a_ns, t_b_ns; /* measurement times on the trajectory clock, before the time offset */
  double d[32];      /* the factor's constants, per kind as listed above; camera factors use the window's S_CtoI / p_CinI */
  double x[16];      /* values of the parameter blocks after the knots */
  const clic_preintegration* pre; /* CLIC_FACTOR_PREINTEGRATION only */
} clic_factor_desc;
/* residual: *n_res values (<= 15).  jac: row-major *n_res x *n_cols, *n_cols = 6 K + the tangent sizes of the blocks in
 * x; column 6 k + c = d/d(theta_k)[c] (c < 3, right perturbation R <- R exp(dtheta), ceres_local_param.h:141-152) or
 * d/d(p_k)[c - 3] of window knot k; the columns after 6 K follow x (a rotation block has 3).  Returns
 * CLIC_ERR_BAD_ARGUMENT when a time is outside the window or jac_capacity (doubles) is too small. */
CLIC_API int clic_evaluate_factor(clic_estimator* h, const clic_factor_desc* f, int32_t* n_res, double* residual /*[15]*/,
                                  int32_t* n_cols, double* jac, int64_t jac_capacity);

/* AddImageFeatureAnalytic (trajectory_estimator.cpp:752-829), batched; landmark[i] indexes
 * clic_window_desc::inv_depth.  Loss CauchyLoss(marg ? 1 : 2) (":806-810"). */
CLIC_API int clic_add_image(clic_estimator* h, int64_t n, const double* t_i, const double* p_i /*3n*/,
                            const double* t_j, const double* p_j /*3n*/, const int32_t* landmark,
                            int32_t fixed_depth, int32_t marg_this_feature, int64_t* n_dropped);

/* AddMarginalizationFactor (trajectory_estimator.cpp:854-866).  The prior stays owned by the caller.
 * In marg mode (is_marg_state) drop_blocks lists the prior blocks (indices into the prior's block list)
 * to marginalize, mirroring PrepareMarginalizationInfo(RType_Prior, ...) (trajectory_manager.cpp:352-375). */
CLIC_API int clic_add_prior(clic_estimator* h, const clic_prior* prior, int32_t n_drop, const int32_t* drop_blocks);

/* Column layout of H/b for the factors added so far. */
CLIC_API int clic_get_layout(clic_estimator* h, clic_layout* out);

/* One fused residual + Jacobian + normal-equation pass at the current state:
 * H (D*D row-major, symmetric, both triangles), b (D), cost = 1/2 sum rho(|r|^2) (excluding factors whose
 * parameters are all constant, reported separately as fixed_cost like Ceres).  Any pointer may be NULL. */
CLIC_API int clic_evaluate(clic_estimator* h, double* H, double* b, double* cost, double* fixed_cost);

/* ResidualSummary-like read-back (trajectory_estimator.h:42-98, what `verbose` prints per update): the cost
 * 1/2 sum rho(|r|^2) of the solve problem at the current state split by ResidualType (marginalization_factor.h:31-46:
 * 0 Pose, 1 IMU, 2 Bias, 3 Gravity, 4 LiDAR, ... 8 GlobalVelocity, 11 Image, 13 Prior; slot 14 = the pose / local-velocity / pre-integration /
 * relative-rotation / image-depth adders together), constant-parameter factors included, and the number of factors added
 * per type.  A debugging aid: it runs one residual-only launch per stream. */
CLIC_API int clic_residual_summary(clic_estimator* h, double cost[16], int64_t count[16]);

/* Bit-exact indexing probe (SplineSegmentMeta::computeTIndexNs, spline_segment.h:67-84):
 * for the factors of one stream, in add order: local knot index s (-1 if the factor was dropped) and u.
 * stream: 0 LiDAR, 1 IMU, 2 image time i, 3 image time j.  IMU/image use the current time offset. */
CLIC_API int clic_get_indices(clic_estimator* h, int32_t stream, int64_t capacity, int32_t* s, double* u, int64_t* n_out);

/* TrajectoryEstimator::Solve (trajectory_estimator.cpp:987-1031): Ceres-1.14 LM trust region,
 * (H + D^2/mu) step by dense Cholesky.  State is updated in place on the device. */
CLIC_API int clic_solve(clic_estimator* h, int32_t max_iterations, clic_summary* summary);

/* Read the (updated) state back; any pointer may be NULL. */
CLIC_API int clic_read_state(clic_estimator* h, double* knot_q, double* knot_p, double* bias,
                             double* t_offset_ns /*3*/, double* inv_depth);
/* Overwrite the state (e.g. to re-run from the same start). */
CLIC_API int clic_write_state(clic_estimator* h, const double* knot_q, const double* knot_p, const double* bias,
                              const double* t_offset_ns /*3*/, const double* inv_depth);

/* ---- the step before the path: forward-only spline evaluation at the current state ----
 * Trajectory::GetSensorPose (src/spline/trajectory.cpp:100-115) batched: pose_S_to_G = poseNs(t) * EP_StoI for every
 * timestamp, with the reference's time arithmetic (double time_ns = t * 1e9 + t_offset_ns, range test in double, then
 * the implicit truncation to int64 of poseNs(int64_t)).  Se3Spline::poseNs = So3Spline::evaluate (so3_spline.h:220-262)
 * and RdSpline::evaluate (rd_spline.h:218-244).  q_out: 4n (x,y,z,w), p_out: 3n.  A time outside
 * [minTimeNs, maxTimeNs) — where the reference asserts — yields the identity pose and counts in n_invalid. */
CLIC_API int clic_query_poses(clic_estimator* h, int64_t n, const double* t_s, const double S_StoI[4],
                              const double p_SinI[3], double t_offset_ns, double* q_out, double* p_out,
                              int64_t* n_invalid);
/* Trajectory::UndistortScanInG (trajectory.cpp:69-97; in_global != 0): p_G = pose_Lk_to_G(t_k) * p_Lk, and
 * Trajectory::UndistortScan (trajectory.cpp:35-67; in_global == 0): p = pose_G_to_target * pose_Lk_to_G(t_k) * p_Lk
 * with pose_G_to_target = GetLidarPose(target_t_s).inverse().  Points are PosPoint x,y,z floats + double timestamp
 * (the cloud type's fields); arithmetic in double, stored back as float.  NaN input points (pcl_isnan(x), skipped
 * by the reference) and out-of-range times give NaN output and count in n_invalid. */
CLIC_API int clic_undistort_scan(clic_estimator* h, int64_t n, const double* t_s, const float* xyz_in /*3n*/,
                                 const double S_LtoI[4], const double p_LinI[3], double t_offset_ns, int32_t in_global,
                                 double target_t_s, float* xyz_out /*3n*/, int64_t* n_invalid);

/* ---- the producer of the path's input: LiDAR data association ----
 * LidarOdometry::SetTargetMap (src/lidar_odometry/lidar_odometry.cpp:472-481): the down-sampled feature map the scan is
 * matched against.  Points are PosPoint x,y,z floats.  corner_xyz may be NULL / n_corner 0 (use_corner_feature false). */
typedef struct clic_feature_map clic_feature_map;
CLIC_API int clic_feature_map_create(const float* surf_xyz /*3n*/, int64_t n_surf, const float* corner_xyz /*3n*/,
                                     int64_t n_corner, clic_feature_map** out);
CLIC_API void clic_feature_map_destroy(clic_feature_map* m);

/* LidarOdometry::FindCorrespondence (lidar_odometry.cpp:483-657).  For every step-th feature of the current scan
 * (step = n / 1500 + 1, ":501, :575"): exact 5 nearest map points of the feature expressed in the map frame (xyz_M, the
 * output of clic_undistort_scan) by squared float distance (pcl::KdTreeFLANN::nearestKSearch, exact search); the
 * feature is skipped when the 5th is farther than 1 m^2 (":508, :584").
 *  corner: centroid (double) and float covariance of the 5 points, float symmetric eigen-decomposition
 *          (cv::eigen, ":511-563"); kept when lambda_0 > 3 lambda_1 as a Line record: geo_point = centroid,
 *          geo_normal = principal direction.  Only when the corner map holds more than 10 points (":502").
 *  surface: skipped when its timestamp < 0 (":577"); float least squares A x = -1 over the 5 points by Householder QR
 *          with column pivoting (":586-603"), plane (x, 1) / |x|; kept when all 5 points lie within 0.2 m of it
 *          (":616-625") as a Plane record geo_plane = (n, d).
 * Outputs are the type-uniform batches clic_add_loam_lines / clic_add_loam_planes take, in the reference's order
 * (corner features first, each in scan order): t_point, point (the feature in the LiDAR frame, xyz_L, as double),
 * geometry, and the index of the source feature.  Each output array must hold ceil(n / step) records
 * (clic_correspondence_capacity).  Arithmetic: float / double exactly where the reference uses each. */
CLIC_API int64_t clic_correspondence_capacity(int64_t n_features);
CLIC_API int clic_find_correspondence(clic_feature_map* m,
                                      int64_t n_corner, const float* corner_xyz_L, const double* corner_t,
                                      const float* corner_xyz_M,
                                      int64_t n_surf, const float* surf_xyz_L, const double* surf_t,
                                      const float* surf_xyz_M,
                                      int64_t* n_lines, double* line_t, double* line_point /*3n*/,
                                      double* line_geo /*6n: geo_point, geo_normal*/, int64_t* line_src,
                                      int64_t* n_planes, double* plane_t, double* plane_point /*3n*/,
                                      double* plane_geo /*4n: n.x n.y n.z d*/, int64_t* plane_src);

/* The producer of one inner iteration, device-resident (odometry_manager.cpp:321-339 calls it twice per scan):
 * LidarOdometry::GetLoamFeatureAssociation (lidar_odometry.cpp:148-180) = Trajectory::UndistortScanInG of the scan's
 * corner / surface features with the current state of `h`, FindCorrespondence against `m`, std::sort by t_point,
 * DownSampleCorrespondence (every cor_downsample-th record, ":660-671") — followed by what TrajectoryManager does with
 * the result, one AddLoamMeasurementAnalytic per kept correspondence (trajectory_manager.cpp:676-688).  The scan goes
 * up once (16 B per feature); the undistorted scan, the neighbours, the fitted geometry and the factor records never
 * leave the device; two integers come back.  Records with equal timestamps keep their FindCorrespondence order (the
 * reference's std::sort leaves that open).  The factors land in a line batch and a plane batch, each time-sorted. */
CLIC_API int clic_add_loam_from_scan(clic_estimator* h, clic_feature_map* m,
                                     int64_t n_corner, const float* corner_xyz /*3n, LiDAR frame*/, const double* corner_t,
                                     int64_t n_surf, const float* surf_xyz /*3n*/, const double* surf_t,
                                     const double S_GtoM[4], const double p_GinM[3], const double S_LtoI[4],
                                     const double p_LinI[3], double t_offset_ns, int32_t cor_downsample, double weight,
                                     int32_t marg_this_factor, int64_t* n_lines, int64_t* n_planes, int64_t* n_dropped);

/* Parity probe: the 5 neighbour indices (nearest first; ties by smaller index) of every selected feature of the last
 * clic_find_correspondence call on this map; type 0 corner, 1 surface; -1 rows for features gated out before the
 * search; INT32_MAX for a neighbour outside the 1 m^2 gate (it cannot change the outcome, and is not searched for).
 * idx holds 5 * ceil(n / step) entries. */
CLIC_API int clic_correspondence_neighbours(clic_feature_map* m, int32_t type, int64_t capacity, int32_t* idx,
                                            int64_t* n_out);
/* Kernels launched on behalf of this map so far, and the device time (CUDA events on the map's stream) of the kernels
 * of the last clic_find_correspondence call; either pointer may be NULL. */
CLIC_API int clic_feature_map_launches(clic_feature_map* m, int64_t* launches, double* last_kernel_ms);

/* SaveMarginalizationInfo (trajectory_estimator.cpp:234-249) = preMarginalize + marginalize
 * (marginalization_factor.cpp:94-116,150-276) over the factors added with marg_* = 1. */
CLIC_API int clic_marginalize(clic_estimator* h, clic_prior** out);

/* Prior objects (MarginalizationInfo after marginalize(): linearized_jacobians J0 (n x n), linearized_residuals r0,
 * keep_block_size/idx/data, marginalization_factor.h:125-142). */
CLIC_API int clic_prior_create(int32_t n, const double* J0 /*n*n row-major*/, const double* r0, int32_t n_blocks,
                               const int32_t* kind, const int32_t* id, const double* x0 /*4 per block*/,
                               clic_prior** out);
CLIC_API int clic_prior_dims(const clic_prior* p, int32_t* n, int32_t* n_blocks, int32_t* m_dropped);
CLIC_API int clic_prior_read(const clic_prior* p, double* J0, double* r0, int32_t* kind, int32_t* id,
                             int32_t* tangent_offset, double* x0);
/* A = J0^T J0 and g = J0^T r0 of the prior (what parity compares: J0 itself has sign/rotation freedom). */
CLIC_API int clic_prior_normal(const clic_prior* p, double* A, double* g);
CLIC_API int clic_prior_release(clic_prior* p);

/* Multi-GPU: factors are sharded across ranks by the caller (each rank adds its shard);
 * H|b|cost are summed with one all-reduce per pass.  id is an ncclUniqueId (128 bytes). */
CLIC_API int clic_comm_unique_id(uint8_t id[128]);
CLIC_API int clic_comm_init(clic_estimator* h, int32_t n_ranks, int32_t rank, const uint8_t id[128]);

/* Measurement hooks used by bench.py: repeat the fused pass `reps` times on the handle's stream without
 * host synchronisation in between (CUDA events are recorded by the caller on that stream). */
/* Optional, after clic_comm_init on every rank of ONE node: map every rank's exchange area into every peer (CUDA IPC)
 * so that the per-pass all-reduce of H | b | cost becomes a single one-shot kernel over NVLink peer memory (pushed
 * slots + flags, summed in rank order: same bits on every rank) instead of ncclAllReduce.  export -> all-gather the
 * 64-byte handles by any means (bench.py: torch.distributed) -> import on every rank.  Process-wide like the
 * communicator.  If the import fails (no IPC between the processes) the NCCL path stays in use. */
CLIC_API int clic_comm_p2p_export(clic_estimator* h, uint8_t handle[64]);
CLIC_API int clic_comm_p2p_import(clic_estimator* h, int32_t n_ranks, int32_t rank, const uint8_t* handles /*64*n*/);
CLIC_API int clic_comm_p2p_ready(void);
CLIC_API int clic_comm_p2p_disable(void); /* back to ncclAllReduce (must be called on every rank) */

/* How the pass of the factors added so far executes.  *fused = 1: ONE persistent launch (clic_b200/csrc/pass_fused.cuh:
 * every factor stream on its own partition of the grid, then the record reduction, the assembly and — on several ranks
 * with mapped peer memory — the exchange of H | b | cost inside the same kernel); *grid = its CTAs, part_ctas[k] = CTAs
 * of stream kind k (0 LiDAR mixed, 1 LiDAR planes, 2 LiDAR lines, 3 IMU, 4 visual with pose packs, 5 visual per-factor
 * chain).  *fused = 0: one kernel per stream + reduction + assembly (+ all-reduce), e.g. two batches of one kind. */
CLIC_API int clic_pass_info(clic_estimator* h, int32_t* fused, int32_t* grid, int32_t part_ctas[6]);
/* Per-CTA %globaltimer stamps (ns) of the LAST fused pass: stamps[8*cta + {0..5}] = start, end of the factor phase,
 * after the grid barrier, after staging the slot keys, end of the assembly, end (after the exchange).  Reads what the
 * previous passes recorded (if enabled then), then sets the switch.
 * Measurement aid for the partition's cost model and bench.py's per-stream times. */
CLIC_API int clic_pass_stamps(clic_estimator* h, int32_t enable, int64_t capacity, uint64_t* stamps, int32_t* n_ctas);

CLIC_API int clic_linearize_async(clic_estimator* h, int32_t reps);
CLIC_API int clic_synchronize(clic_estimator* h);
CLIC_API int clic_num_kernel_launches(clic_estimator* h, int64_t* launches);
/* Per-kernel-class device time, measured with CUDA events on the handle's stream while enabled (bench.py's
 * roofline line).  Classes: 0 LiDAR factor kernel, 1 IMU factor kernel, 2 visual factor kernel,
 * 3 record reduction + assembly, 4 prior / bias terms, 5 LM step (scaling, Cholesky, solves, update), 6 all-reduce,
 * 7 marginalization.  clic_profile_read synchronizes and returns the sums since the last enable. */
CLIC_API int clic_profile_enable(clic_estimator* h, int32_t enable);
CLIC_API int clic_profile_read(clic_estimator* h, double ms[8], int64_t launches[8]);

#ifdef __cplusplus
}
#endif
#endif /* CLIC_CUDA_H_ */
