// Header-only C++ shim: the reference's clic::TrajectoryEstimator surface
// (src/estimator/trajectory_estimator.h:100-278) on top of the C ABI of include/clic_cuda.h.
//
// Method names, argument order and parameter-block semantics follow the reference: parameter blocks are
// caller-owned raw double* (knots inside the Trajectory, biases, inverse depths) that are updated in place when
// Solve() returns.  Per-measurement Add* calls only append to host staging arrays; the batched ABI calls (and the
// host->device upload) happen inside Solve() / SaveMarginalizationInfo().  No Ceres, no Eigen: small POD vector
// types replace Eigen::Vector3d / Sophus::SO3d (same memory layout: 3 doubles / 4 doubles x,y,z,w).
// Link against libclic_cuda.so (product) — or, in tests only, the oracle library, which exports the same symbols.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../clic_cuda.h"

namespace clic {

using Vec3d = std::array<double, 3>;
using Vec6d = std::array<double, 6>;
struct SO3d {  // unit quaternion (x,y,z,w): Sophus::SO3d::data() layout (src/sophus_lib/so3.hpp:153-160)
  std::array<double, 4> q{{0, 0, 0, 1}};
  const double* data() const { return q.data(); }
};

enum SensorType { IMUSensor = 0, LiDARSensor, CameraSensor };  // src/spline/trajectory.h:29-33
enum GeometryType { Line = 0, Plane };                         // src/lidar_odometry/lidar_feature.h:108

struct PointCorrespondence {  // src/lidar_odometry/lidar_feature.h:110-123
  double t_point = 0, t_map = 0;
  Vec3d point{};
  GeometryType geo_type = Plane;
  std::array<double, 4> geo_plane{};
  Vec3d geo_normal{};
  Vec3d geo_point{};
};
struct IMUData {  // src/utils/parameter_struct.h:52-58
  double timestamp = 0;
  Vec3d gyro{}, accel{};
};
struct PoseData {  // src/utils/parameter_struct.h:60-69
  double timestamp = 0;
  Vec3d position{};
  SO3d orientation;
};
// The members of IntegrationBase (src/visual_odometry/integration_base.h:254-288) PreIntegrationFactor reads.  The
// integration itself (push_back / midPointIntegration) stays with the caller's own class, as in the reference; copy its
// members here.  Matrices are row-major.
struct IntegrationBase {
  double sum_dt = 0;
  Vec3d delta_p{}, delta_v{};
  std::array<double, 4> delta_q{{0, 0, 0, 1}};  // x, y, z, w
  Vec3d linearized_ba{}, linearized_bg{};
  std::array<double, 225> jacobian{};    // 15 x 15, state order P R V BA BG
  std::array<double, 225> covariance{};  // 15 x 15, symmetric positive definite
  double gravity_norm = -9.8;            // the global GRAVITY_NORM (parameter_struct.cpp:23)

  // sqrt_info = LLT(covariance^-1).matrixL()^T (trajectory_value_factor.h:744-748), without Eigen: covariance = L L^T,
  // covariance^-1 = L^-T L^-1 =: M, then the Cholesky factor of M.  Throws if the covariance is not positive definite.
  clic_preintegration to_abi() const {
    clic_preintegration p;
    p.sum_dt = sum_dt;
    p.gravity_norm = gravity_norm;
    for (int i = 0; i < 3; i++) {
      p.delta_p[i] = delta_p[i]; p.delta_v[i] = delta_v[i];
      p.linearized_ba[i] = linearized_ba[i]; p.linearized_bg[i] = linearized_bg[i];
    }
    for (int i = 0; i < 4; i++) p.delta_q[i] = delta_q[i];
    for (int i = 0; i < 225; i++) p.jacobian[i] = jacobian[i];
    constexpr int n = 15;
    auto cholesky = [](const double* A, double* L) {  // lower, row-major
      for (int i = 0; i < n * n; i++) L[i] = 0.0;
      for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= L[j * n + k] * L[j * n + k];
        if (!(d > 0.0)) throw std::runtime_error("IntegrationBase: covariance is not positive definite");
        L[j * n + j] = std::sqrt(d);
        for (int i = j + 1; i < n; i++) {
          double v = A[i * n + j];
          for (int k = 0; k < j; k++) v -= L[i * n + k] * L[j * n + k];
          L[i * n + j] = v / L[j * n + j];
        }
      }
    };
    double L[225], Li[225], M[225], C[225];
    cholesky(covariance.data(), L);
    for (int c = 0; c < n; c++)  // Li = L^-1 (forward substitution, column by column)
      for (int r = 0; r < n; r++) {
        double v = r == c ? 1.0 : 0.0;
        for (int k = 0; k < r; k++) v -= L[r * n + k] * Li[k * n + c];
        Li[r * n + c] = v / L[r * n + r];
      }
    for (int r = 0; r < n; r++)
      for (int c = 0; c < n; c++) {
        double v = 0.0;
        for (int k = 0; k < n; k++) v += Li[k * n + r] * Li[k * n + c];
        M[r * n + c] = v;
      }
    cholesky(M, C);
    for (int r = 0; r < n; r++)
      for (int c = 0; c < n; c++) p.sqrt_info[r * n + c] = C[c * n + r];  // matrixL().transpose()
    return p;
  }
};
struct ExtrinsicParam {  // src/utils/parameter_struct.h:108-150 (the fields the hot path reads)
  SO3d so3;
  Vec3d p{};
  double t_offset_ns = 0;
};

// Knot storage + extrinsics: the part of clic::Trajectory (src/spline/trajectory.h, se3_spline.h) the estimator needs.
struct Trajectory {
  typedef std::shared_ptr<Trajectory> Ptr;
  int64_t t0_ns = 0, dt_ns = 0;
  int knot_id0 = 0;  // global index of knots_*[0]
  std::vector<std::array<double, 4>> knots_so3;
  std::vector<Vec3d> knots_pos;
  std::map<SensorType, ExtrinsicParam> EP_StoI;
  Trajectory(double time_interval, double start_time = 0)
      : t0_ns((int64_t)(start_time * 1e9)), dt_ns((int64_t)(time_interval * 1e9)) {
    EP_StoI[IMUSensor] = ExtrinsicParam();
    EP_StoI[LiDARSensor] = ExtrinsicParam();
    EP_StoI[CameraSensor] = ExtrinsicParam();
  }
  size_t numKnots() const { return knots_so3.size(); }
  double* getKnotSO3(size_t i) { return knots_so3[i].data(); }
  double* getKnotPos(size_t i) { return knots_pos[i].data(); }
};

struct LockExtrinsic {  // trajectory_estimator_options.h:26-30
  bool lock_P = true, lock_R = true, lock_t_offset = true;
};
struct TrajectoryEstimatorOptions {  // trajectory_estimator_options.h:32-72
  TrajectoryEstimatorOptions() {
    lock_EPs[IMUSensor] = LockExtrinsic();
    lock_EPs[LiDARSensor] = LockExtrinsic();
    lock_EPs[CameraSensor] = LockExtrinsic();
  }
  std::map<SensorType, LockExtrinsic> lock_EPs;
  bool use_auto_diff = false;
  int64_t t_offset_padding_ns = (int64_t)1e7;
  bool lock_traj = false;
  bool lock_ab = true, lock_wb = true, lock_g = true;
  bool is_marg_state = false;
  int ctrl_to_be_opt_now = 0, ctrl_to_be_opt_later = 0;
  bool marg_bias_param = true, marg_gravity_param = true, marg_t_offset_param = true;
  bool show_residual_summary = false;
  // execution (not in the reference)
  int device = 0;
  void* stream = nullptr;
};

// POD stand-in for ceres::Solver::Summary (the fields the reference reads: trajectory_manager.cpp:753-756)
struct Summary {
  clic_summary raw{};
  int num_successful_steps = 0, num_unsuccessful_steps = 0;
  double initial_cost = 0, final_cost = 0;
  std::string BriefReport() const {
    static const char* term[] = {"CONVERGENCE", "NO_CONVERGENCE", "FAILURE"};
    return "clic_b200 report: iterations: " + std::to_string(raw.iterations + 1) + ", initial cost: " +
           std::to_string(initial_cost) + ", final cost: " + std::to_string(final_cost) +
           ", termination: " + term[raw.termination < 0 || raw.termination > 2 ? 2 : raw.termination];
  }
};

struct PosPoint {  // my_pcl::PointXYZT, src/utils/mypcl_cloud_type.h:46-51 (x, y, z floats + laser timestamp)
  float x = 0, y = 0, z = 0;
  double timestamp = 0;
};
typedef std::vector<PosPoint> PosCloud;

// The target of LidarOdometry::SetTargetMap (lidar_odometry.cpp:472-481): the down-sampled surface / corner feature
// map, resident on the device (clic_feature_map).  An empty corner cloud = use_corner_feature false.
class LidarFeatureMap {
 public:
  typedef std::shared_ptr<LidarFeatureMap> Ptr;
  LidarFeatureMap(const PosCloud& surface_features, const PosCloud& corner_features) {
    std::vector<float> s = xyz(surface_features), c = xyz(corner_features);
    int rc = clic_feature_map_create(s.data(), (int64_t)surface_features.size(), c.empty() ? nullptr : c.data(),
                                     (int64_t)corner_features.size(), &h_);
    if (rc != CLIC_OK) throw std::runtime_error(std::string("clic_feature_map_create: ") + clic_last_error());
  }
  ~LidarFeatureMap() { clic_feature_map_destroy(h_); }
  LidarFeatureMap(const LidarFeatureMap&) = delete;
  LidarFeatureMap& operator=(const LidarFeatureMap&) = delete;
  clic_feature_map* handle() const { return h_; }
  static std::vector<float> xyz(const PosCloud& c) {
    std::vector<float> v(3 * c.size());
    for (size_t i = 0; i < c.size(); i++) { v[3 * i] = c[i].x; v[3 * i + 1] = c[i].y; v[3 * i + 2] = c[i].z; }
    return v;
  }

 private:
  clic_feature_map* h_ = nullptr;
};

// MarginalizationInfo (marginalization_factor.h:100-142): owns a clic_prior; the kept parameter blocks are
// returned to the caller as raw pointers exactly like getParameterBlocks() (marginalization_factor.cpp:301-317).
struct MarginalizationInfo {
  typedef std::shared_ptr<MarginalizationInfo> Ptr;
  clic_prior* prior = nullptr;
  int m = 0, n = 0;
  ~MarginalizationInfo() {
    if (prior) clic_prior_release(prior);
  }
};

// marginalization_factor.h:31-46
enum ResidualType {
  RType_Pose = 0, RType_IMU, RType_Bias, RType_Gravity, RType_LiDAR, RType_LiDAR_Relative, RType_LiDAROpt,
  RType_PreIntegration, RType_GlobalVelocity, RType_LocalVelocity, RType_Local6DoFVel, RType_Image, RType_Epipolar,
  RType_Prior
};

// MarginalizationFactor (marginalization_factor.h:145-154): the cost function the caller wraps a prior in before
// handing it to PrepareMarginalizationInfo(RType_Prior, ...) (trajectory_manager.cpp:370-376, 488-494).  Here it only
// carries the prior; its Evaluate (marginalization_factor.cpp:331-372) runs on the device (prior_kernel).
struct MarginalizationFactor {
  explicit MarginalizationFactor(MarginalizationInfo::Ptr _marginalization_info)
      : marginalization_info(std::move(_marginalization_info)) {}
  MarginalizationInfo::Ptr marginalization_info;
};

// 2 x 2 matrix with just enough arithmetic for `image_feature_weight * Matrix2d::Identity()` (trajectory_manager.cpp:57-58)
struct Matrix2d {
  double m[4] = {1, 0, 0, 1};
  static Matrix2d Identity() { return Matrix2d(); }
  friend Matrix2d operator*(double s, Matrix2d a) {
    for (double& x : a.m) x *= s;
    return a;
  }
  double operator()(int r, int c) const { return m[2 * r + c]; }
};

namespace analytic_derivative {
// The static parameters TrajectoryManager::InitFactorInfo sets on the visual factor class once per run
// (trajectory_manager.cpp:53-61; image_feature_factor.h:270-284): camera extrinsics and sqrt_info = image_weight * I2.
// The estimator reads them when it builds the problem (sqrt_info -> clic_options::image_sqrt_info; the camera extrinsics
// also travel in Trajectory::EP_StoI[CameraSensor], which is what the reference's SetParam call is fed from).
struct ImageFeatureFactor {
  static void SetParam(const SO3d& _S_CtoI, const Vec3d& _p_CinI) {
    S_CtoI() = _S_CtoI;
    p_CinI() = _p_CinI;
    param_set() = true;
  }
  static SO3d& S_CtoI() { static SO3d v; return v; }
  static Vec3d& p_CinI() { static Vec3d v{}; return v; }
  static bool& param_set() { static bool v = false; return v; }
  static inline Matrix2d sqrt_info = (450. / 1.5) * Matrix2d::Identity();  // class default, image_feature_factor.h:277-278
};
}  // namespace analytic_derivative

class TrajectoryEstimator {
 public:
  typedef std::shared_ptr<TrajectoryEstimator> Ptr;

  TrajectoryEstimator(Trajectory::Ptr trajectory, TrajectoryEstimatorOptions& option, std::string descri = "")
      : trajectory_(trajectory), options(option), descri_(descri), fixed_control_point_index_(-1) {}
  ~TrajectoryEstimator() {
    if (h_) clic_destroy(h_);
  }

  void SetFixedIndex(int idx) { fixed_control_point_index_ = idx; }  // trajectory_estimator.h:140

  // trajectory_estimator.h:153-156
  void AddIMUMeasurementAnalytic(const IMUData& imu_data, double* gyro_bias, double* accel_bias, double* gravity,
                                 const Vec6d& info_vec, bool marg_this_factor = false) {
    gravity_ = gravity;
    const int slot = bias_slot(gyro_bias, accel_bias);
    if (imu_.empty() || imu_.back().slot != slot || imu_.back().info != info_vec || imu_.back().marg != marg_this_factor) {
      ImuBatch b;
      b.slot = slot;
      b.info = info_vec;
      b.marg = marg_this_factor;
      imu_.push_back(b);
    }
    ImuBatch& b = imu_.back();
    b.t.push_back(imu_data.timestamp);
    b.gyro.insert(b.gyro.end(), imu_data.gyro.begin(), imu_data.gyro.end());
    b.accel.insert(b.accel.end(), imu_data.accel.begin(), imu_data.accel.end());
  }

  // trajectory_estimator.h:159-162
  void AddBiasFactor(double* bias_gyr_i, double* bias_gyr_j, double* bias_acc_i, double* bias_acc_j, double dt,
                     const Vec6d& info_vec, bool marg_this_factor = false, bool marg_all_bias = false) {
    BiasFac f;
    f.i = bias_slot(bias_gyr_i, bias_acc_i);
    f.j = bias_slot(bias_gyr_j, bias_acc_j);
    f.dt = dt;
    f.info = info_vec;
    f.marg = marg_this_factor;
    f.marg_all = marg_all_bias;
    bias_f_.push_back(f);
  }

  // trajectory_estimator.cpp:609-664 (IMUGlobalVelocityFactor, CauchyLoss(60)); false when the timestamp is outside
  // the window (MeasuredTimeToNs(IMUSensor), ":272-292" — repeated by the library when the problem is built)
  bool AddGlobalVelocityMeasurement(const double timestamp, const Vec3d& global_v, double vel_weight) {
    const double toff_s = 1e-9 * trajectory_->EP_StoI[IMUSensor].t_offset_ns;
    const int64_t K = (int64_t)trajectory_->numKnots();
    const int64_t t_min = (int64_t)((1e-9 * (double)trajectory_->t0_ns - toff_s) * 1e9);
    const int64_t t_max =  // Se3Spline::maxTimeNs = t0 + (n - N + 1) dt - 1 (rd_spline.h:142-144)
        (int64_t)((1e-9 * (double)(trajectory_->t0_ns + (K - 3) * trajectory_->dt_ns - 1) - toff_s) * 1e9);
    const int64_t pad = options.lock_EPs.at(IMUSensor).lock_t_offset ? 0 : options.t_offset_padding_ns;
    const int64_t t = (int64_t)(timestamp * 1e9);
    if (t - pad < t_min || t + pad >= t_max) return false;
    VelFac f;
    f.t = timestamp;
    f.v = global_v;
    f.w = vel_weight;
    vel_f_.push_back(f);
    return true;
  }

  // trajectory_estimator.h:188-193
  void AddLoamMeasurementAnalytic(const PointCorrespondence& pc, const SO3d& S_GtoM, const Vec3d& p_GinM,
                                  const SO3d& S_LtoI, const Vec3d& p_LinI, double weight,
                                  bool marg_this_factor = false) {
    if (loam_.empty() || loam_.back().S_GtoM.q != S_GtoM.q || loam_.back().p_GinM != p_GinM ||
        loam_.back().S_LtoI.q != S_LtoI.q || loam_.back().p_LinI != p_LinI || loam_.back().weight != weight ||
        loam_.back().marg != marg_this_factor) {
      LoamBatch b;
      b.S_GtoM = S_GtoM; b.p_GinM = p_GinM; b.S_LtoI = S_LtoI; b.p_LinI = p_LinI; b.weight = weight;
      b.marg = marg_this_factor;
      loam_.push_back(b);
    }
    // type-uniform staging (clic_add_loam_planes / clic_add_loam_lines): 64 B per plane, 80 B per line feature
    LoamBatch& b = loam_.back();
    if (pc.geo_type == Plane) {
      b.pl_t.push_back(pc.t_point);
      b.pl_point.insert(b.pl_point.end(), pc.point.begin(), pc.point.end());
      b.pl_geo.insert(b.pl_geo.end(), pc.geo_plane.begin(), pc.geo_plane.end());
    } else {
      b.ln_t.push_back(pc.t_point);
      b.ln_point.insert(b.ln_point.end(), pc.point.begin(), pc.point.end());
      b.ln_geo.insert(b.ln_geo.end(), pc.geo_point.begin(), pc.geo_point.end());
      b.ln_geo.insert(b.ln_geo.end(), pc.geo_normal.begin(), pc.geo_normal.end());
    }
  }

  // One inner iteration's LiDAR factors straight from the scan (no reference counterpart as one call):
  // LidarOdometry::GetLoamFeatureAssociation (lidar_odometry.cpp:148-180: UndistortScanInG with the trajectory as it
  // stands when the problem is built, FindCorrespondence against `map`, sort by t_point, DownSampleCorrespondence)
  // + the AddLoamMeasurementAnalytic loop of trajectory_manager.cpp:682-686, all on the device.
  void AddLoamMeasurementsFromScan(const LidarFeatureMap::Ptr& map, const PosCloud& corner_features,
                                   const PosCloud& surface_features, const SO3d& S_GtoM, const Vec3d& p_GinM,
                                   const SO3d& S_LtoI, const Vec3d& p_LinI, double weight, int cor_downsample = 1,
                                   bool marg_this_factor = false) {
    ScanBatch b;
    b.map = map;
    b.corner_xyz = LidarFeatureMap::xyz(corner_features);
    b.surf_xyz = LidarFeatureMap::xyz(surface_features);
    for (const auto& p : corner_features) b.corner_t.push_back(p.timestamp);
    for (const auto& p : surface_features) b.surf_t.push_back(p.timestamp);
    b.S_GtoM = S_GtoM; b.p_GinM = p_GinM; b.S_LtoI = S_LtoI; b.p_LinI = p_LinI; b.weight = weight;
    b.cor_downsample = cor_downsample;
    b.marg = marg_this_factor;
    scans_.push_back(std::move(b));
  }
  // (lines, planes) each scan batch contributed, after the problem has been built (Solve / SaveMarginalizationInfo)
  const std::vector<std::pair<int64_t, int64_t>>& ScanCorrespondenceCounts() const { return scan_counts_; }

  // Trajectory::GetLidarPose (trajectory.h:134-136) for many timestamps at once, at the current state.  Builds the
  // problem: call it after the Add* calls, or on an estimator made for the query.
  void GetLidarPoses(const std::vector<double>& t_s, std::vector<SO3d>* q_LtoG, std::vector<Vec3d>* p_LinG) {
    build();
    const ExtrinsicParam& ep = trajectory_->EP_StoI[LiDARSensor];
    std::vector<double> q(4 * t_s.size()), p(3 * t_s.size());
    int64_t bad = 0;
    check(clic_query_poses(h_, (int64_t)t_s.size(), t_s.data(), ep.so3.data(), ep.p.data(), ep.t_offset_ns, q.data(),
                           p.data(), &bad));
    q_LtoG->resize(t_s.size());
    p_LinG->resize(t_s.size());
    for (size_t i = 0; i < t_s.size(); i++) {
      std::memcpy((*q_LtoG)[i].q.data(), &q[4 * i], 4 * sizeof(double));
      std::memcpy((*p_LinG)[i].data(), &p[3 * i], 3 * sizeof(double));
    }
  }

  // trajectory_estimator.h:196-199
  void AddImageFeatureAnalytic(const double ti, const Vec3d& pi, const double tj, const Vec3d& pj, double* inv_depth,
                               bool fixed_depth = false, bool marg_this_fearure = false) {
    if (image_.empty() || image_.back().fixed != fixed_depth || image_.back().marg != marg_this_fearure) {
      ImageBatch b;
      b.fixed = fixed_depth;
      b.marg = marg_this_fearure;
      image_.push_back(b);
    }
    ImageBatch& b = image_.back();
    auto it = landmark_.find(inv_depth);
    if (it == landmark_.end()) {
      it = landmark_.emplace(inv_depth, (int)landmark_ptr_.size()).first;
      landmark_ptr_.push_back(inv_depth);
    }
    b.ti.push_back(ti);
    b.tj.push_back(tj);
    b.pi.insert(b.pi.end(), pi.begin(), pi.end());
    b.pj.insert(b.pj.end(), pj.begin(), pj.end());
    b.lm.push_back(it->second);
  }

  // trajectory_estimator.h:207-209.  In marg mode `drop_blocks` mirrors the drop_set the reference computes at
  // trajectory_manager.cpp:352-375 (indices into the prior's block list).
  void AddMarginalizationFactor(MarginalizationInfo::Ptr last_marginalization_info,
                                std::vector<double*>& /*last_marginalization_parameter_blocks*/,
                                const std::vector<int>& drop_blocks = {}) {
    priors_.push_back({last_marginalization_info, drop_blocks});
  }

  // trajectory_estimator.h:228-232 with a MarginalizationFactor (RType_Prior): what UpdateLIOPrior /
  // UpdateVisualOffsetPrior call to route the previous prior into the marginalization set, `drop_set` = indices into
  // `parameter_blocks` (trajectory_manager.cpp:352-376, 488-494).  Takes ownership of the factor like the reference's
  // MarginalizationInfo does (marginalization_factor.cpp:64-74).  Other residual types reach the marginalization set
  // through their Add* call with marg_this_factor = true, exactly as in the reference.
  void PrepareMarginalizationInfo(ResidualType r_type, MarginalizationFactor* cost_function, void* /*loss_function*/,
                                  std::vector<double*>& parameter_blocks, std::vector<int>& drop_set) {
    if (r_type != RType_Prior || !cost_function)
      throw std::runtime_error("PrepareMarginalizationInfo: only RType_Prior factors are passed in by the caller");
    std::unique_ptr<MarginalizationFactor> own(cost_function);
    AddMarginalizationFactor(own->marginalization_info, parameter_blocks, drop_set);
  }

  // trajectory_estimator.h:144 / .cpp:303-310: makes the time-offset blocks constant when their lock flag is set.
  // Here the lock flags travel with the options when the problem is built, so there is nothing left to do.
  void SetTimeoffsetState() {}

  // trajectory_estimator.cpp:251-270: every knot whose segment starts before max_time becomes constant.  (The reference
  // keeps the method but its only call site is commented out, trajectory_manager.cpp:263.)
  void SetKeyScanConstant(double max_time) {
    const double toff_s = 1e-9 * trajectory_->EP_StoI[LiDARSensor].t_offset_ns;
    const int64_t t_ns = (int64_t)((max_time + toff_s) * 1e9);
    if (t_ns < trajectory_->t0_ns) return;
    const int64_t s = (t_ns - trajectory_->t0_ns) / trajectory_->dt_ns;  // start index of the segment holding max_time
    const int last = trajectory_->knot_id0 + (int)std::min<int64_t>(s + 3, (int64_t)trajectory_->numKnots() - 1);
    if (last > fixed_control_point_index_) fixed_control_point_index_ = last;
  }

  // trajectory_estimator.h:150-151 (IMUPoseFactor); measurements outside the window are dropped like the reference's
  // early return (trajectory_estimator.cpp:329)
  void AddPoseMeasurementAnalytic(const PoseData& pose_data, const Vec6d& info_vec) {
    if (pose_f_.empty() || pose_f_.back().info != info_vec) {
      PoseBatch b;
      b.info = info_vec;
      pose_f_.push_back(b);
    }
    PoseBatch& b = pose_f_.back();
    b.t.push_back(pose_data.timestamp);
    b.p.insert(b.p.end(), pose_data.position.begin(), pose_data.position.end());
    b.q.insert(b.q.end(), pose_data.orientation.q.begin(), pose_data.orientation.q.end());
  }
  // trajectory_estimator.h:146-147 / .cpp:312-323: the pose at the trajectory's start time, weights 100
  void AddStartTimePose(const PoseData& pose) {
    PoseData pose_temp = pose;
    pose_temp.timestamp = 1e-9 * (double)trajectory_->t0_ns - 1e-9 * trajectory_->EP_StoI[LiDARSensor].t_offset_ns;
    AddPoseMeasurementAnalytic(pose_temp, Vec6d{{100, 100, 100, 100, 100, 100}});
  }
  // trajectory_estimator.h:183-185 (LocalVelocityFactor)
  void AddLocalVelocityMeasurementAnalytic(const double timestamp, const Vec3d& local_v, double weight) {
    LocalVel f;
    f.t = timestamp;
    f.v = local_v;
    f.w = weight;
    lvel_f_.push_back(f);
  }
  // trajectory_estimator.h:168-172 (PreIntegrationFactor).  The reference hands its double SECONDS ti / tj to the
  // factor's int64 nanosecond arguments (trajectory_estimator.cpp:545); here the factor gets the nanosecond values
  // MeasuredTimeToNs produced (INTEGRATION.md "quirks").  marg_this_factor: the reference routes the factor to the
  // marginalization set (":564-578"); not supported here (throws), nothing in the pipeline calls this adder.
  void AddPreIntegrationAnalytic(double ti, double tj, const IntegrationBase* pre_integration, double* gyro_bias_i,
                                 double* gyro_bias_j, double* accel_bias_i, double* accel_bias_j,
                                 bool marg_this_factor = false) {
    if (marg_this_factor && options.is_marg_state)
      throw std::runtime_error("AddPreIntegrationAnalytic: marg_this_factor is not supported");
    PreInt f;
    f.ta_ns = (int64_t)(ti * 1e9);
    f.tb_ns = (int64_t)(tj * 1e9);
    f.pre = pre_integration->to_abi();
    f.i = bias_slot(gyro_bias_i, accel_bias_i);
    f.j = bias_slot(gyro_bias_j, accel_bias_j);
    preint_f_.push_back(f);
  }
  // One analytic factor's residual and Jacobian at the current window: the surface the reference's
  // test_analytic_factor.cpp drives (cost_function->Evaluate + GetJacobian, ":846-900"); see clic_evaluate_factor.
  // jac: row-major n_res x n_cols (resized).  False when a measurement time is outside the window.
  bool EvaluateFactor(const clic_factor_desc& f, std::vector<double>* residual, std::vector<double>* jac, int* n_cols) {
    build();
    int32_t nr = 0, nc = 0;
    double r[15];
    std::vector<double> J((size_t)15 * (6 * trajectory_->numKnots() + 16));
    if (clic_evaluate_factor(h_, &f, &nr, r, &nc, J.data(), (int64_t)J.size()) != CLIC_OK) return false;
    if (residual) residual->assign(r, r + nr);
    if (jac) jac->assign(J.begin(), J.begin() + (size_t)nr * nc);
    if (n_cols) *n_cols = nc;
    return true;
  }
  // trajectory_estimator.h:174-175 (RelativeOrientationFactor)
  void AddRelativeRotationAnalytic(double ta, double tb, const SO3d& S_BtoA, const Vec3d& info_vec) {
    RelRot f;
    f.ta = ta;
    f.tb = tb;
    f.S = S_BtoA;
    f.info = info_vec;
    relrot_f_.push_back(f);
  }
  // trajectory_estimator.h:201-204 (ImageDepthFactor, CauchyLoss(1))
  void AddImageDepthAnalytic(const Vec3d& p_i, const Vec3d& p_j, const SO3d& S_CitoCj, const Vec3d& p_CiinCj,
                             double* inv_depth) {
    auto it = landmark_.find(inv_depth);
    if (it == landmark_.end()) {
      it = landmark_.emplace(inv_depth, (int)landmark_ptr_.size()).first;
      landmark_ptr_.push_back(inv_depth);
    }
    ImgDepth f;
    f.pi = p_i;
    f.pj = p_j;
    f.S = S_CitoCj;
    f.p = p_CiinCj;
    f.lm = it->second;
    depth_f_.push_back(f);
  }

  // trajectory_estimator.h:164-165
  void AddGravityFactor(double* gravity, const Vec3d& info_vec, bool marg_this_factor = false) {
    gravity_ = gravity;
    GravFac f;
    f.g0 = {{gravity[0], gravity[1], gravity[2]}};
    f.info = info_vec;
    f.marg = marg_this_factor;
    grav_f_.push_back(f);
  }

  // trajectory_estimator.h:225-226
  Summary Solve(int max_iterations = 50, bool /*progress*/ = false, int /*num_threads*/ = -1) {
    build();
    clic_summary s{};
    check(clic_solve(h_, max_iterations, &s));
    write_back();
    Summary out;
    out.raw = s;
    out.num_successful_steps = s.num_successful_steps;
    out.num_unsuccessful_steps = s.num_unsuccessful_steps;
    out.initial_cost = s.initial_cost;
    out.final_cost = s.final_cost;
    return out;
  }

  // trajectory_estimator.h:234-235
  void SaveMarginalizationInfo(MarginalizationInfo::Ptr& marg_info_out, std::vector<double*>& marg_param_blocks_out) {
    build();
    clic_prior* p = nullptr;
    int rc = clic_marginalize(h_, &p);
    marg_param_blocks_out.clear();
    if (rc == CLIC_ERR_NOTHING_TO_KEEP) {  // trajectory_estimator.cpp:241-248
      marg_info_out = nullptr;
      return;
    }
    check(rc);
    marg_info_out = std::make_shared<MarginalizationInfo>();
    marg_info_out->prior = p;
    int32_t n = 0, nb = 0, m = 0;
    check(clic_prior_dims(p, &n, &nb, &m));
    marg_info_out->n = n;
    marg_info_out->m = m;
    std::vector<int32_t> kind(nb), id(nb);
    check(clic_prior_read(p, nullptr, nullptr, kind.data(), id.data(), nullptr, nullptr));
    for (int i = 0; i < nb; i++) marg_param_blocks_out.push_back(block_address(kind[i], id[i]));
  }

  // one fused pass, for tests / parity probes (not in the reference API)
  void Evaluate(std::vector<double>* H, std::vector<double>* b, double* cost) {
    build();
    clic_layout L;
    check(clic_get_layout(h_, &L));
    if (H) H->assign((size_t)L.D * L.D, 0.0);
    if (b) b->assign(L.D, 0.0);
    check(clic_evaluate(h_, H ? H->data() : nullptr, b ? b->data() : nullptr, cost, nullptr));
  }

  TrajectoryEstimatorOptions options;

 private:
  struct LoamBatch {
    std::vector<double> pl_t, pl_point, pl_geo, ln_t, ln_point, ln_geo;
    SO3d S_GtoM, S_LtoI;
    Vec3d p_GinM{}, p_LinI{};
    double weight = 1;
    bool marg = false;
  };
  struct ScanBatch {
    LidarFeatureMap::Ptr map;
    std::vector<float> corner_xyz, surf_xyz;
    std::vector<double> corner_t, surf_t;
    SO3d S_GtoM, S_LtoI;
    Vec3d p_GinM{}, p_LinI{};
    double weight = 1;
    int cor_downsample = 1;
    bool marg = false;
  };
  struct ImuBatch {
    std::vector<double> t, gyro, accel;
    int slot = 0;
    Vec6d info{};
    bool marg = false;
  };
  struct ImageBatch {
    std::vector<double> ti, tj, pi, pj;
    std::vector<int32_t> lm;
    bool fixed = true, marg = false;
  };
  struct VelFac {
    double t, w;
    Vec3d v{};
  };
  struct BiasFac {
    int i, j;
    double dt;
    Vec6d info;
    bool marg, marg_all;
  };
  struct PriorRef {
    MarginalizationInfo::Ptr info;
    std::vector<int> drop;
  };
  struct GravFac {
    Vec3d g0{}, info{};
    bool marg = false;
  };
  struct PoseBatch {
    std::vector<double> t, p, q;
    Vec6d info{};
  };
  struct LocalVel {
    double t = 0, w = 0;
    Vec3d v{};
  };
  struct PreInt {
    int64_t ta_ns = 0, tb_ns = 0;
    clic_preintegration pre;
    int i = 0, j = 0;
  };
  struct RelRot {
    double ta = 0, tb = 0;
    SO3d S;
    Vec3d info{};
  };
  struct ImgDepth {
    Vec3d pi{}, pj{}, p{};
    SO3d S;
    int lm = 0;
  };

  static void check(int rc) {
    if (rc != CLIC_OK) throw std::runtime_error(std::string("clic: ") + clic_last_error());
  }
  int bias_slot(double* gyro, double* accel) {
    for (size_t i = 0; i < bias_ptr_.size(); i++)
      if (bias_ptr_[i].first == gyro) return (int)i;
    bias_ptr_.push_back({gyro, accel});
    return (int)bias_ptr_.size() - 1;
  }
  double* block_address(int kind, int id) {
    switch (kind) {
      case CLIC_BLOCK_KNOT_ROT: return trajectory_->getKnotSO3(id - trajectory_->knot_id0);
      case CLIC_BLOCK_KNOT_POS: return trajectory_->getKnotPos(id - trajectory_->knot_id0);
      case CLIC_BLOCK_BIAS_GYRO: return bias_ptr_.at(id).first;
      case CLIC_BLOCK_BIAS_ACCEL: return bias_ptr_.at(id).second;
      case CLIC_BLOCK_GRAVITY: return gravity_;
      case CLIC_BLOCK_T_OFFSET: return &trajectory_->EP_StoI[(SensorType)id].t_offset_ns;
      default: return landmark_ptr_.at(id);
    }
  }

  // create the handle and push every staged batch through the batched ABI
  void build() {
    if (h_) return;
    const size_t K = trajectory_->numKnots();
    std::vector<double> kq(4 * K), kp(3 * K), bias(6 * bias_ptr_.size()), invd(landmark_ptr_.size());
    for (size_t k = 0; k < K; k++) {
      std::memcpy(&kq[4 * k], trajectory_->knots_so3[k].data(), 4 * sizeof(double));
      std::memcpy(&kp[3 * k], trajectory_->knots_pos[k].data(), 3 * sizeof(double));
    }
    for (size_t i = 0; i < bias_ptr_.size(); i++) {
      std::memcpy(&bias[6 * i], bias_ptr_[i].first, 3 * sizeof(double));
      std::memcpy(&bias[6 * i + 3], bias_ptr_[i].second, 3 * sizeof(double));
    }
    for (size_t i = 0; i < landmark_ptr_.size(); i++) invd[i] = *landmark_ptr_[i];
    clic_window_desc w{};
    w.t0_ns = trajectory_->t0_ns;
    w.dt_ns = trajectory_->dt_ns;
    w.n_knots = (int32_t)K;
    w.knot_id0 = trajectory_->knot_id0;
    w.knot_q = kq.data();
    w.knot_p = kp.data();
    const ExtrinsicParam& epl = trajectory_->EP_StoI[LiDARSensor];
    const ExtrinsicParam& epc = trajectory_->EP_StoI[CameraSensor];
    std::memcpy(w.S_LtoI, epl.so3.data(), sizeof(w.S_LtoI));
    std::memcpy(w.p_LinI, epl.p.data(), sizeof(w.p_LinI));
    std::memcpy(w.S_CtoI, epc.so3.data(), sizeof(w.S_CtoI));
    std::memcpy(w.p_CinI, epc.p.data(), sizeof(w.p_CinI));
    for (int s = 0; s < 3; s++) w.t_offset_ns[s] = trajectory_->EP_StoI[(SensorType)s].t_offset_ns;
    if (gravity_) std::memcpy(w.gravity, gravity_, sizeof(w.gravity));
    w.n_bias = (int32_t)bias_ptr_.size();
    w.bias = bias.data();
    w.n_landmarks = (int32_t)invd.size();
    w.inv_depth = invd.data();
    clic_options o;
    clic_default_options(&o);
    o.lock_traj = options.lock_traj;
    o.lock_ab = options.lock_ab;
    o.lock_wb = options.lock_wb;
    o.lock_g = options.lock_g;
    for (int s = 0; s < 3; s++) o.lock_t_offset[s] = options.lock_EPs.at((SensorType)s).lock_t_offset;
    o.is_marg_state = options.is_marg_state;
    o.t_offset_padding_ns = options.t_offset_padding_ns;
    o.ctrl_to_be_opt_now = options.ctrl_to_be_opt_now;
    o.ctrl_to_be_opt_later = options.ctrl_to_be_opt_later;
    o.marg_bias_param = options.marg_bias_param;
    o.marg_gravity_param = options.marg_gravity_param;
    o.marg_t_offset_param = options.marg_t_offset_param;
    o.device = options.device;
    o.stream = options.stream;
    o.image_sqrt_info = analytic_derivative::ImageFeatureFactor::sqrt_info(0, 0);  // InitFactorInfo, trajectory_manager.cpp:57-61
    if (analytic_derivative::ImageFeatureFactor::param_set()) {  // ImageFeatureFactor::SetParam wins, as in the reference
      std::memcpy(w.S_CtoI, analytic_derivative::ImageFeatureFactor::S_CtoI().data(), sizeof(w.S_CtoI));
      std::memcpy(w.p_CinI, analytic_derivative::ImageFeatureFactor::p_CinI().data(), sizeof(w.p_CinI));
    }
    check(clic_create(&w, &o, &h_));
    check(clic_set_fixed_index(h_, fixed_control_point_index_));
    // the reference's call order is prior, LiDAR, image, IMU, bias (trajectory_manager.cpp:659-750).  Upload order here:
    // plane batch first (big asynchronous copy; the visual batch's host-side staging runs under it), image, IMU, line
    // batch last (the kernel left exposed after the last byte arrives is the smallest).  H, b do not depend on it.
    for (auto& p : priors_)
      check(clic_add_prior(h_, p.info->prior, (int32_t)p.drop.size(), p.drop.empty() ? nullptr : p.drop.data()));
    for (auto& b : loam_)
      if (!b.pl_t.empty())
        check(clic_add_loam_planes(h_, (int64_t)b.pl_t.size(), b.pl_t.data(), b.pl_point.data(), b.pl_geo.data(),
                                   b.S_GtoM.data(), b.p_GinM.data(), b.S_LtoI.data(), b.p_LinI.data(), b.weight, b.marg,
                                   nullptr));
    for (auto& b : image_)
      check(clic_add_image(h_, (int64_t)b.ti.size(), b.ti.data(), b.pi.data(), b.tj.data(), b.pj.data(), b.lm.data(),
                           b.fixed, b.marg, nullptr));
    for (auto& b : imu_)
      check(clic_add_imu(h_, (int64_t)b.t.size(), b.t.data(), b.gyro.data(), b.accel.data(), b.slot, b.info.data(),
                         b.marg, nullptr));
    for (auto& b : loam_)
      if (!b.ln_t.empty())
        check(clic_add_loam_lines(h_, (int64_t)b.ln_t.size(), b.ln_t.data(), b.ln_point.data(), b.ln_geo.data(),
                                  b.S_GtoM.data(), b.p_GinM.data(), b.S_LtoI.data(), b.p_LinI.data(), b.weight, b.marg,
                                  nullptr));
    for (auto& f : bias_f_) check(clic_add_bias(h_, f.i, f.j, f.dt, f.info.data(), f.marg, f.marg_all));
    for (auto& f : grav_f_) check(clic_add_gravity(h_, f.g0.data(), f.info.data(), f.marg));
    for (auto& b : pose_f_)
      check(clic_add_pose(h_, (int64_t)b.t.size(), b.t.data(), b.p.data(), b.q.data(), b.info.data(), nullptr));
    for (auto& f : lvel_f_) check(clic_add_local_velocity(h_, 1, &f.t, f.v.data(), f.w, nullptr));
    for (auto& f : preint_f_) check(clic_add_preintegration(h_, f.ta_ns, f.tb_ns, &f.pre, f.i, f.j, nullptr));
    for (auto& f : relrot_f_) check(clic_add_relative_rotation(h_, f.ta, f.tb, f.S.data(), f.info.data(), nullptr));
    for (auto& f : depth_f_) check(clic_add_image_depth(h_, f.pi.data(), f.pj.data(), f.S.data(), f.p.data(), f.lm));
    for (auto& f : vel_f_) check(clic_add_global_velocity(h_, f.t, f.v.data(), f.w, nullptr));
    for (auto& b : scans_) {
      int64_t nl = 0, np = 0;
      check(clic_add_loam_from_scan(h_, b.map->handle(), (int64_t)b.corner_t.size(), b.corner_xyz.data(),
                                    b.corner_t.data(), (int64_t)b.surf_t.size(), b.surf_xyz.data(), b.surf_t.data(),
                                    b.S_GtoM.data(), b.p_GinM.data(), b.S_LtoI.data(), b.p_LinI.data(),
                                    trajectory_->EP_StoI[LiDARSensor].t_offset_ns, b.cor_downsample, b.weight, b.marg, &nl,
                                    &np, nullptr));
      scan_counts_.push_back({nl, np});
    }
  }

  // Ceres updates the caller's parameter blocks in place; so do we
  void write_back() {
    const size_t K = trajectory_->numKnots();
    std::vector<double> kq(4 * K), kp(3 * K), bias(6 * bias_ptr_.size() + 1), invd(landmark_ptr_.size() + 1);
    double toff[3];
    check(clic_read_state(h_, kq.data(), kp.data(), bias.data(), toff, invd.data()));
    for (size_t k = 0; k < K; k++) {
      std::memcpy(trajectory_->knots_so3[k].data(), &kq[4 * k], 4 * sizeof(double));
      std::memcpy(trajectory_->knots_pos[k].data(), &kp[3 * k], 3 * sizeof(double));
    }
    for (size_t i = 0; i < bias_ptr_.size(); i++) {
      std::memcpy(bias_ptr_[i].first, &bias[6 * i], 3 * sizeof(double));
      std::memcpy(bias_ptr_[i].second, &bias[6 * i + 3], 3 * sizeof(double));
    }
    for (int s = 0; s < 3; s++) trajectory_->EP_StoI[(SensorType)s].t_offset_ns = toff[s];
    for (size_t i = 0; i < landmark_ptr_.size(); i++) *landmark_ptr_[i] = invd[i];  // free inverse depths (fixed_depth = false)
  }

  Trajectory::Ptr trajectory_;
  std::string descri_;
  int fixed_control_point_index_;
  clic_estimator* h_ = nullptr;
  double* gravity_ = nullptr;
  std::vector<LoamBatch> loam_;
  std::vector<ScanBatch> scans_;
  std::vector<std::pair<int64_t, int64_t>> scan_counts_;
  std::vector<ImuBatch> imu_;
  std::vector<ImageBatch> image_;
  std::vector<BiasFac> bias_f_;
  std::vector<GravFac> grav_f_;
  std::vector<PoseBatch> pose_f_;
  std::vector<LocalVel> lvel_f_;
  std::vector<RelRot> relrot_f_;
  std::vector<PreInt> preint_f_;
  std::vector<ImgDepth> depth_f_;
  std::vector<VelFac> vel_f_;
  std::vector<PriorRef> priors_;
  std::vector<std::pair<double*, double*>> bias_ptr_;
  std::map<double*, int> landmark_;
  std::vector<double*> landmark_ptr_;
};

}  // namespace clic
