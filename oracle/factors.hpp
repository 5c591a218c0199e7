// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference arithmetic (oracle).
//
// The analytic cost functions on the hot path, restated with the same
// ceres::CostFunction::Evaluate(parameters, residuals, jacobians) contract the reference uses:
// parameters = array of parameter-block pointers, jacobians[i] = row-major (num_residuals x block_size)
// or nullptr to skip.  Rotation blocks are 4 wide (xyzw); the first three Jacobian columns are the
// tangent Jacobian and the fourth is zero (src/estimator/factor/ceres_local_param.h:141-152).
#pragma once
#include <algorithm>
#include <limits>
#include <memory>
#include <vector>

#include "spline_views.hpp"

namespace oracle {

struct CostFunction {
  int num_residuals = 0;
  std::vector<int> block_sizes;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
};

// PointCorrespondence — src/lidar_odometry/lidar_feature.h:110-123
struct PointCorrespondence {
  double t_point;
  double t_map;
  Vec3 point;
  int geo_type;  // 0 Line, 1 Plane (":108")
  double geo_plane[4];
  Vec3 geo_normal;
  Vec3 geo_point;
};

// LoamFeatureFactor — src/estimator/factor/analytic_diff/lidar_feature_factor.h:31-172
struct LoamFeatureFactor : CostFunction {
  int64_t t_point_ns_;
  PointCorrespondence pc_;
  SegmentMeta meta_;
  Quat S_GtoM_;
  Vec3 p_GinM_;
  Quat S_LtoI_;
  Vec3 p_LinI_;
  double weight_;
  LoamFeatureFactor(int64_t t_point_ns, const PointCorrespondence& pc, const SegmentMeta& meta, const Quat& S_GtoM,
                    const Vec3& p_GinM, const Quat& S_LtoI, const Vec3& p_LinI, double weight)
      : t_point_ns_(t_point_ns), pc_(pc), meta_(meta), S_GtoM_(S_GtoM), p_GinM_(p_GinM), S_LtoI_(S_LtoI),
        p_LinI_(p_LinI), weight_(weight) {
    num_residuals = 1;
    size_t kont_num = meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    Vec3 p_Lk = pc_.point;
    Vec3 p_IK = so3_rotate(S_LtoI_, p_Lk) + p_LinI_;
    size_t kont_num = meta_.NumParameters();
    Quat S_ItoG = so3_evaluate_Rp(t_point_ns_, meta_, parameters, jacobians ? &J_R : nullptr);
    Vec3 p_IinG = rd_evaluate(0, t_point_ns_, meta_, parameters + kont_num, jacobians ? &J_p : nullptr);
    Vec3 p_M = so3_rotate(S_GtoM_, so3_rotate(S_ItoG, p_IK) + p_IinG) + p_GinM_;

    Vec3 J_pi = v3(0, 0, 0);
    if (pc_.geo_type == 1) {  // Plane
      Vec3 n = v3(pc_.geo_plane[0], pc_.geo_plane[1], pc_.geo_plane[2]);
      residuals[0] = dot(p_M, n) + pc_.geo_plane[3];
      J_pi = n;
    } else {
      Vec3 dist_vec = cross(p_M - pc_.geo_point, pc_.geo_normal);
      residuals[0] = norm(dist_vec);
      // J_pi = -dist_vec^T / r * hat(geo_normal)   (":101")
      J_pi = rowmul(-dist_vec / residuals[0], hat(pc_.geo_normal));
    }
    residuals[0] *= weight_;
    if (!jacobians) return true;

    for (size_t i = 0; i < kont_num; ++i) {
      if (jacobians[i])
        for (int c = 0; c < 4; c++) jacobians[i][c] = 0;
      if (jacobians[i + kont_num])
        for (int c = 0; c < 3; c++) jacobians[i + kont_num][c] = 0;
    }
    Mat3 J_Xm_R = -(so3_matrix(S_GtoM_) * so3_matrix(S_ItoG) * hat(p_IK));
    Vec3 jac_lhs_R = rowmul(J_pi, J_Xm_R);
    Vec3 jac_lhs_P = rowmul(J_pi, so3_matrix(S_GtoM_));
    for (size_t i = 0; i < kont_num; i++) {
      if (jacobians[i]) {
        // the factor is built on a 4-knot segment (trajectory_estimator.cpp:718), so i indexes d_val_d_knot directly
        Vec3 row = rowmul(jac_lhs_R, J_R.d_val_d_knot[i]);
        jacobians[i][0] = weight_ * row[0];
        jacobians[i][1] = weight_ * row[1];
        jacobians[i][2] = weight_ * row[2];
        jacobians[i][3] = weight_ * 0.0;
      }
    }
    for (size_t i = 0; i < kont_num; i++) {
      size_t idx = kont_num + i;
      if (jacobians[idx]) {
        for (int c = 0; c < 3; c++) jacobians[idx][c] = weight_ * (J_p.d_val_d_knot[i] * jac_lhs_P[c]);
      }
    }
    return true;
  }
};

// IMUData — src/utils/parameter_struct.h:52-58
struct IMUData {
  double timestamp;
  Vec3 gyro;
  Vec3 accel;
};

// IMUFactor — src/estimator/factor/analytic_diff/trajectory_value_factor.h:128-301
struct IMUFactor : CostFunction {
  int64_t time_ns_;
  IMUData imu_data_;
  SegmentMeta meta_;
  double info_vec_[6];
  IMUFactor(int64_t time_ns, const IMUData& imu, const SegmentMeta& meta, const double info_vec[6])
      : time_ns_(time_ns), imu_data_(imu), meta_(meta) {
    for (int i = 0; i < 6; i++) info_vec_[i] = info_vec[i];
    num_residuals = 6;
    size_t knot_num = meta_.NumParameters();
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(3);  // gyro bias
    block_sizes.push_back(3);  // accel bias
    block_sizes.push_back(3);  // gravity
    block_sizes.push_back(1);  // time_offset
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_rot_w, J_rot_a;
    RdJacobian J_pos;
    size_t knot_num = meta_.NumParameters();
    Vec3 gyro_bias = load_v(parameters[2 * knot_num]);
    Vec3 accel_bias = load_v(parameters[2 * knot_num + 1]);
    Vec3 gravity = load_v(parameters[2 * knot_num + 2]);
    double time_offset_in_ns = parameters[2 * knot_num + 3][0];
    int64_t t_corrected = time_ns_ + (int64_t)time_offset_in_ns;  // ":175" truncation
    SplineIMUData sd;
    if (jacobians)
      sd = split_evaluate(t_corrected, meta_, parameters, gravity, &J_rot_w, &J_rot_a, &J_pos);
    else
      sd = split_evaluate(t_corrected, meta_, parameters, gravity, nullptr, nullptr, nullptr);
    Vec3 rg = sd.gyro - (imu_data_.gyro - gyro_bias);
    Vec3 ra = sd.accel - (imu_data_.accel - accel_bias);
    for (int i = 0; i < 3; i++) {
      residuals[i] = info_vec_[i] * rg[i];
      residuals[3 + i] = info_vec_[3 + i] * ra[i];
    }
    if (!jacobians) return true;

    for (size_t i = 0; i < knot_num; ++i) {
      if (jacobians[i])
        for (int c = 0; c < 24; c++) jacobians[i][c] = 0;
      if (jacobians[i + knot_num])
        for (int c = 0; c < 18; c++) jacobians[i + knot_num][c] = 0;
    }
    for (size_t i = 0; i < (size_t)N; i++) {
      size_t idx = i + sd.start_idx;
      if (jacobians[idx]) {
        double* J = jacobians[idx];  // 6x4 row-major
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            J[r * 4 + c] = info_vec_[r] * J_rot_w.d_val_d_knot[i](r, c);
            J[(3 + r) * 4 + c] = info_vec_[3 + r] * J_rot_a.d_val_d_knot[i](r, c);
          }
        for (int r = 0; r < 6; r++) J[r * 4 + 3] = info_vec_[r] * 0.0;
      }
    }
    Mat3 Rinv = so3_matrix(sd.R_inv);
    for (size_t i = 0; i < (size_t)N; i++) {
      size_t idx = knot_num + i + sd.start_idx;
      if (jacobians[idx]) {
        double* J = jacobians[idx];  // 6x3 row-major
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            J[r * 3 + c] = info_vec_[r] * 0.0;
            J[(3 + r) * 3 + c] = info_vec_[3 + r] * (J_pos.d_val_d_knot[i] * Rinv(r, c));
          }
      }
    }
    if (jacobians[2 * knot_num]) {
      double* J = jacobians[2 * knot_num];
      for (int c = 0; c < 18; c++) J[c] = 0;
      for (int r = 0; r < 3; r++) J[r * 3 + r] = info_vec_[r];
    }
    if (jacobians[2 * knot_num + 1]) {
      double* J = jacobians[2 * knot_num + 1];
      for (int c = 0; c < 18; c++) J[c] = 0;
      for (int r = 0; r < 3; r++) J[(3 + r) * 3 + r] = info_vec_[3 + r];
    }
    if (jacobians[2 * knot_num + 2]) {
      double* J = jacobians[2 * knot_num + 2];
      for (int c = 0; c < 18; c++) J[c] = 0;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) J[(3 + r) * 3 + c] = info_vec_[3 + r] * Rinv(r, c);
    }
    if (jacobians[2 * knot_num + 3]) {
      double* J = jacobians[2 * knot_num + 3];  // 6x1
      Vec3 rot_accel = so3_acceleration_body(t_corrected, meta_, parameters);
      Vec3 jerk_R3 = rd_evaluate(3, t_corrected, meta_, parameters + knot_num, nullptr);
      Mat3 rot = so3_matrix(qconj(sd.R_inv));
      Mat3 gyro_hat = hat(sd.gyro);
      Mat3 rot_dot = rot * gyro_hat;
      Vec3 top = vee(transpose(rot_dot) * rot_dot) + vee(transpose(rot) * (rot_dot * gyro_hat + rot * hat(rot_accel)));
      Vec3 bot = transpose(rot_dot) * rot * sd.accel + transpose(rot) * jerk_R3;
      for (int r = 0; r < 3; r++) {
        J[r] = (1e-9 * info_vec_[r]) * top[r];
        J[3 + r] = (1e-9 * info_vec_[3 + r]) * bot[r];
      }
    }
    return true;
  }
};

// BiasFactor — trajectory_value_factor.h:65-126
// PoseData — src/utils/parameter_struct.h (timestamp, position, orientation)
struct PoseData {
  double timestamp;
  Vec3 position;
  Quat orientation;
};

// IMUPoseFactor — src/estimator/factor/analytic_diff/trajectory_value_factor.h:303-433
struct IMUPoseFactor : CostFunction {
  int64_t time_ns_;
  PoseData pose_data_;
  SegmentMeta meta_;
  double info_vec_[6];
  IMUPoseFactor(int64_t time_ns, const PoseData& pose, const SegmentMeta& meta, const double info_vec[6])
      : time_ns_(time_ns), pose_data_(pose), meta_(meta) {
    for (int i = 0; i < 6; i++) info_vec_[i] = info_vec[i];
    num_residuals = 6;
    size_t knot_num = meta_.NumParameters();
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(1);  // time_offset
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t knot_num = meta_.NumParameters();
    size_t P_offset = knot_num;
    double time_offset_in_ns = parameters[2 * knot_num][0];
    int64_t t_corrected = time_ns_ + (int64_t)time_offset_in_ns;
    Quat S_ItoG = so3_evaluate_rotation(t_corrected, meta_, parameters, jacobians ? &J_R : nullptr);
    Vec3 p_IinG = rd_evaluate(0, t_corrected, meta_, parameters + P_offset, jacobians ? &J_p : nullptr);
    Vec3 res_R = so3_log(so3_mul(S_ItoG, qconj(pose_data_.orientation)));
    Vec3 res_p = p_IinG - pose_data_.position;
    if (jacobians) {
      for (size_t i = 0; i < knot_num; ++i) {
        if (jacobians[i])
          for (int c = 0; c < 24; c++) jacobians[i][c] = 0;
        if (jacobians[i + knot_num])
          for (int c = 0; c < 18; c++) jacobians[i + knot_num][c] = 0;
      }
      Mat3 Jrot = left_jacobian_inv_so3(res_R);
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = J_R.start_idx + i;
        if (jacobians[idx]) {
          Mat3 JJ = Jrot * J_R.d_val_d_knot[i];
          double* J = jacobians[idx];  // 6x4 row-major
          for (int c = 0; c < 24; c++) J[c] = 0;
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) J[r * 4 + c] = info_vec_[r] * JJ(r, c);
        }
      }
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = J_R.start_idx + i;
        if (jacobians[idx + knot_num]) {
          double* J = jacobians[idx + knot_num];  // 6x3 row-major
          for (int c = 0; c < 18; c++) J[c] = 0;
          for (int r = 0; r < 3; r++) J[(3 + r) * 3 + r] = info_vec_[3 + r] * J_p.d_val_d_knot[i];
        }
      }
      if (jacobians[2 * knot_num]) {  // ":410-426"
        double* J = jacobians[2 * knot_num];
        Vec3 gyro = so3_velocity_body(t_corrected, meta_, parameters, nullptr);
        Vec3 vel = rd_evaluate(1, t_corrected, meta_, parameters + knot_num, nullptr);
        // SO3d(Eigen::Quaterniond(gyro_hat).normalized()): a quaternion made from the SKEW matrix, as written
        Quat qg = qnormalized(quat_from_matrix_eigen(hat(gyro)));
        Quat rot_dot = so3_mul(S_ItoG, qg);
        Vec3 jr = so3_log(so3_mul(rot_dot, qconj(pose_data_.orientation)));
        for (int r = 0; r < 3; r++) {
          J[r] = 1e-9 * info_vec_[r] * jr[r];
          J[3 + r] = 1e-9 * info_vec_[3 + r] * vel[r];
        }
      }
    }
    for (int r = 0; r < 3; r++) {
      residuals[r] = info_vec_[r] * res_R[r];
      residuals[3 + r] = info_vec_[3 + r] * res_p[r];
    }
    return true;
  }
};

// LocalVelocityFactor — trajectory_value_factor.h:435-571
struct LocalVelocityFactor : CostFunction {
  int64_t time_ns_;
  Vec3 local_velocity_;
  SegmentMeta meta_;
  double weight_;
  LocalVelocityFactor(int64_t time_ns, const Vec3& local_velocity, const SegmentMeta& meta, double weight)
      : time_ns_(time_ns), local_velocity_(local_velocity), meta_(meta), weight_(weight) {
    num_residuals = 3;
    size_t knot_num = meta_.NumParameters();
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(1);  // time_offset
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t knot_num = meta_.NumParameters();
    double time_offset_in_ns = parameters[2 * knot_num][0];
    int64_t t_corrected = time_ns_ + (int64_t)time_offset_in_ns;
    Quat S_ItoG = so3_evaluate_Rp(t_corrected, meta_, parameters, jacobians ? &J_R : nullptr);
    Vec3 v_inG = rd_evaluate(1, t_corrected, meta_, parameters + knot_num, jacobians ? &J_p : nullptr);
    Vec3 res = so3_rotate(S_ItoG, local_velocity_) - v_inG;
    for (int r = 0; r < 3; r++) residuals[r] = weight_ * res[r];
    if (!jacobians) return true;
    for (size_t i = 0; i < knot_num; ++i) {
      if (jacobians[i])
        for (int c = 0; c < 12; c++) jacobians[i][c] = 0;
      if (jacobians[i + knot_num])
        for (int c = 0; c < 9; c++) jacobians[i + knot_num][c] = 0;
    }
    Mat3 jac_lhs_R = -1.0 * (so3_matrix(S_ItoG) * hat(local_velocity_));
    for (size_t i = 0; i < (size_t)N; i++) {
      size_t idx = i + J_R.start_idx;
      if (jacobians[idx]) {
        Mat3 JJ = jac_lhs_R * J_R.d_val_d_knot[i];
        double* J = jacobians[idx];  // 3x4
        for (int c = 0; c < 12; c++) J[c] = 0;
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) J[r * 4 + c] = weight_ * JJ(r, c);
      }
    }
    for (size_t i = 0; i < (size_t)N; i++) {
      size_t idx = knot_num + i + J_p.start_idx;
      if (jacobians[idx]) {
        double* J = jacobians[idx];  // 3x3
        for (int c = 0; c < 9; c++) J[c] = 0;
        for (int r = 0; r < 3; r++) J[r * 3 + r] = weight_ * (J_p.d_val_d_knot[i] * -1.0);
      }
    }
    if (jacobians[2 * knot_num]) {  // ":548-563"
      double* J = jacobians[2 * knot_num];
      Vec3 gyro = so3_velocity_body(t_corrected, meta_, parameters, nullptr);
      Vec3 accel = rd_evaluate(2, t_corrected, meta_, parameters + knot_num, nullptr);
      Mat3 rot_dot = so3_matrix(S_ItoG) * hat(gyro);
      Vec3 jt = rot_dot * local_velocity_ - accel;
      for (int r = 0; r < 3; r++) J[r] = 1e-9 * weight_ * jt[r];
    }
    return true;
  }
};

// RelativeOrientationFactor — trajectory_value_factor.h:966-1075.  ta_ns_ / tb_ns_ are doubles in the reference
// (":1073") and go through the templated computeTIndexNs (floor of a double quotient, spline_segment.h:108-118); the
// times an estimator hands in are whole nanoseconds (MeasuredTimeToNs), for which both index paths agree except exactly
// on a knot boundary with u = 0 — restated here with the double path.
struct RelativeOrientationFactor : CostFunction {
  Quat S_BtoA_;
  double ta_ns_, tb_ns_;
  SplineMeta spline_meta_;
  double sqrt_info_[3];
  RelativeOrientationFactor(const Quat& S_BtoA, int64_t ta_ns, int64_t tb_ns, const SplineMeta& meta, const double info_vec[3])
      : S_BtoA_(S_BtoA), ta_ns_((double)ta_ns), tb_ns_((double)tb_ns), spline_meta_(meta) {
    for (int i = 0; i < 3; i++) sqrt_info_[i] = info_vec[i];
    num_residuals = 3;
    size_t knot_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(4);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R[2];
    size_t R_offset[2] = {0, 0};
    size_t seg_idx[2] = {0, 0};
    {
      double u;
      spline_meta_.ComputeSplineIndex((int64_t)ta_ns_, R_offset[0], u);
      spline_meta_.ComputeSplineIndex((int64_t)tb_ns_, R_offset[1], u);
      size_t segment0_knot_num = spline_meta_.segments.at(0).NumParameters();
      for (int i = 0; i < 2; ++i) {
        if (R_offset[i] >= segment0_knot_num) {
          seg_idx[i] = 1;
          R_offset[i] = segment0_knot_num;
        } else {
          R_offset[i] = 0;
        }
      }
    }
    Quat S_IatoG = so3_evaluate_rotation((int64_t)ta_ns_, spline_meta_.segments.at(seg_idx[0]), parameters + R_offset[0],
                                         jacobians ? &J_R[0] : nullptr);
    Quat S_GtoIb = so3_evaluate_logRT((int64_t)tb_ns_, spline_meta_.segments.at(seg_idx[1]), parameters + R_offset[1],
                                      jacobians ? &J_R[1] : nullptr);
    Vec3 res_R = so3_log(so3_mul(so3_mul(S_GtoIb, S_IatoG), S_BtoA_));
    for (int r = 0; r < 3; r++) residuals[r] = sqrt_info_[r] * res_R[r];
    size_t kont_num = spline_meta_.NumParameters();
    if (jacobians) {
      for (size_t i = 0; i < kont_num; ++i)
        if (jacobians[i])
          for (int c = 0; c < 12; c++) jacobians[i][c] = 0;
      Mat3 Jrot = left_jacobian_inv_so3(res_R);
      Mat3 jac_lhs[2] = {Jrot * so3_matrix(S_GtoIb), -1.0 * Jrot};
      for (int seg = 0; seg < 2; ++seg)
        for (size_t i = 0; i < (size_t)N; i++) {
          size_t idx = R_offset[seg] + J_R[seg].start_idx + i;
          if (jacobians[idx]) {
            Mat3 JJ = jac_lhs[seg] * J_R[seg].d_val_d_knot[i];
            double* J = jacobians[idx];  // 3x4, accumulated (":1056": +=)
            for (int r = 0; r < 3; r++)
              for (int c = 0; c < 3; c++) J[r * 4 + c] += sqrt_info_[r] * JJ(r, c);
          }
        }
    }
    return true;
  }
};

// ImageDepthFactor — src/estimator/factor/analytic_diff/image_feature_factor.h:658-707
struct ImageDepthFactor : CostFunction {
  Vec3 p_i_, p_j_;
  Quat S_CitoCj_;
  Vec3 p_CiinCj_;
  double sqrt_info;  // static 450/1.5 * I2 (":696-697"), overwritten by InitFactorInfo (trajectory_manager.cpp:70)
  ImageDepthFactor(const Vec3& p_i, const Vec3& p_j, const Quat& S_CitoCj, const Vec3& p_CiinCj, double _sqrt_info)
      : p_i_(p_i), p_j_(p_j), S_CitoCj_(S_CitoCj), p_CiinCj_(p_CiinCj), sqrt_info(_sqrt_info) {
    num_residuals = 2;
    block_sizes = {1};
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    double d_inv = parameters[0][0];
    Vec3 x_ci = p_i_ / d_inv;
    Vec3 x_j = so3_rotate(S_CitoCj_, x_ci) + p_CiinCj_;
    double depth_j_inv = 1.0 / x_j[2];
    double r0 = x_j[0] * depth_j_inv - p_j_[0], r1 = x_j[1] * depth_j_inv - p_j_[1];
    if (jacobians && jacobians[0]) {
      Vec3 J_Xm_d = -1.0 * (so3_matrix(S_CitoCj_) * x_ci) / d_inv;
      double j0 = depth_j_inv * J_Xm_d[0] - depth_j_inv * depth_j_inv * x_j[0] * J_Xm_d[2];
      double j1 = depth_j_inv * J_Xm_d[1] - depth_j_inv * depth_j_inv * x_j[1] * J_Xm_d[2];
      jacobians[0][0] = sqrt_info * j0;
      jacobians[0][1] = sqrt_info * j1;
    }
    residuals[0] = sqrt_info * r0;
    residuals[1] = sqrt_info * r1;
    return true;
  }
};

// GravityFactor — src/estimator/factor/analytic_diff/trajectory_value_factor.h:33-62
struct GravityFactor : CostFunction {
  Vec3 gravity_;
  double sqrt_info_[3];
  GravityFactor(const double gravity[3], const double sqrt_info[3]) : gravity_(load_v(gravity)) {
    for (int i = 0; i < 3; i++) sqrt_info_[i] = sqrt_info[i];
    num_residuals = 3;
    block_sizes = {3};
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    Vec3 g = load_v(parameters[0]);
    for (int i = 0; i < 3; i++) residuals[i] = sqrt_info_[i] * (g[i] - gravity_[i]);
    if (jacobians && jacobians[0]) {
      for (int c = 0; c < 9; c++) jacobians[0][c] = 0;
      for (int r = 0; r < 3; r++) jacobians[0][r * 3 + r] = sqrt_info_[r];
    }
    return true;
  }
};

struct BiasFactor : CostFunction {
  double sqrt_info_[6];
  BiasFactor(double dt, const double sqrt_info[6]) {
    double sqrt_dt = std::sqrt(dt);
    for (int i = 0; i < 6; i++) sqrt_info_[i] = sqrt_info[i] / sqrt_dt;
    num_residuals = 6;
    block_sizes = {3, 3, 3, 3};
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    Vec3 bg_i = load_v(parameters[0]), bg_j = load_v(parameters[1]);
    Vec3 ba_i = load_v(parameters[2]), ba_j = load_v(parameters[3]);
    Vec3 rg = bg_j - bg_i, ra = ba_j - ba_i;
    for (int i = 0; i < 3; i++) {
      residuals[i] = sqrt_info_[i] * rg[i];
      residuals[3 + i] = sqrt_info_[3 + i] * ra[i];
    }
    if (jacobians) {
      const double sign[4] = {-1, 1, -1, 1};
      const int row0[4] = {0, 0, 3, 3};
      for (int b = 0; b < 4; b++) {
        if (!jacobians[b]) continue;
        for (int c = 0; c < 18; c++) jacobians[b][c] = 0;
        for (int r = 0; r < 3; r++) jacobians[b][(row0[b] + r) * 3 + r] = sqrt_info_[row0[b] + r] * sign[b];
      }
    }
    return true;
  }
};

// IMUGlobalVelocityFactor — src/estimator/factor/auto_diff/trajectory_value_factor.h:436-487 (an auto-diff functor in
// the reference; the Jacobians here are its derivatives written out).  Parameters: the position knots of the spline
// meta, then the IMU time offset (1).  t_corrected = double(time_ns) + time_offset (NOT truncated: the offset is a Jet
// there), indexed through the templated SplineSegmentMeta::computeTIndexNs (spline_segment.h:108-118: floor of a double
// quotient, u = st - floor(st)); velocity = inv_dt * sum_i (M p'(u))_i p_i (ceres_spline_helper_jet.h:201-221).
struct GlobalVelocityFactor : CostFunction {
  int64_t time_ns_;
  Vec3 velocity_;
  SplineMeta spline_meta_;
  double vel_weight_, inv_dt_;
  GlobalVelocityFactor(int64_t time_ns, const Vec3& v, const SplineMeta& meta, double w)
      : time_ns_(time_ns), velocity_(v), spline_meta_(meta), vel_weight_(w) {
    inv_dt_ = 1e9 * 1.0 / (double)spline_meta_.segments.begin()->dt_ns;
    num_residuals = 3;
    block_sizes.assign(spline_meta_.NumParameters(), 3);
    block_sizes.push_back(1);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    const size_t Knot_offset = spline_meta_.NumParameters();
    const double t_corrected = (double)time_ns_ + parameters[Knot_offset][0];
    // SplineMeta::ComputeSplineIndex<T> over the segments
    size_t P_offset = 0;
    double u = 0;
    bool found = false;
    for (auto const& seg : spline_meta_.segments) {
      if (!(t_corrected < (double)seg.MinTimeNs() || t_corrected >= (double)seg.MaxTimeNs())) {
        const double st = (t_corrected - (double)seg.t0_ns) / (double)seg.dt_ns;
        const size_t s = (size_t)std::floor(st);
        u = st - (double)s;
        P_offset += s;
        found = true;
        break;
      }
      P_offset += seg.NumParameters();
    }
    if (!found) return false;
    // baseCoeffsWithTime<1>: p' = [0, 1, 2u, 3u^2]; coeff = inv_dt * M p'   (plain blending matrix M)
    const double u2 = u * u;
    double c[4], ca[4];
    c[0] = inv_dt_ * ((-3.0 + 6.0 * u - 3.0 * u2) / 6.0);
    c[1] = inv_dt_ * ((-12.0 * u + 9.0 * u2) / 6.0);
    c[2] = inv_dt_ * ((3.0 + 6.0 * u - 9.0 * u2) / 6.0);
    c[3] = inv_dt_ * ((3.0 * u2) / 6.0);
    // d coeff / du = inv_dt * M p'' with p'' = [0, 0, 2, 6u]
    ca[0] = inv_dt_ * ((6.0 - 6.0 * u) / 6.0);
    ca[1] = inv_dt_ * ((-12.0 + 18.0 * u) / 6.0);
    ca[2] = inv_dt_ * ((6.0 - 18.0 * u) / 6.0);
    ca[3] = inv_dt_ * ((6.0 * u) / 6.0);
    Vec3 v = v3(0, 0, 0), dv_du = v3(0, 0, 0);
    for (int i = 0; i < 4; i++) {
      const Vec3 p = load_v(parameters[P_offset + i]);
      v = v + c[i] * p;
      dv_du = dv_du + ca[i] * p;
    }
    for (int r = 0; r < 3; r++) residuals[r] = vel_weight_ * (v[r] - velocity_[r]);
    if (jacobians) {
      for (size_t k = 0; k < Knot_offset; k++) {
        if (!jacobians[k]) continue;
        for (int e = 0; e < 9; e++) jacobians[k][e] = 0;
        if (k >= P_offset && k < P_offset + 4)
          for (int r = 0; r < 3; r++) jacobians[k][r * 3 + r] = vel_weight_ * c[k - P_offset];
      }
      if (jacobians[Knot_offset]) {  // du / d t_offset = 1 / dt_ns
        const double du = 1.0 / (double)spline_meta_.segments.begin()->dt_ns;
        for (int r = 0; r < 3; r++) jacobians[Knot_offset][r] = vel_weight_ * dv_du[r] * du;
      }
    }
    return true;
  }
};

// ImageFeatureFactor — src/estimator/factor/analytic_diff/image_feature_factor.h:31-292
struct ImageFeatureFactor : CostFunction {
  int64_t t_i_ns_, t_j_ns_;
  Vec3 p_i_, p_j_;
  SplineMeta spline_meta_;
  Quat S_CtoI;
  Vec3 p_CinI;
  // static Eigen::Matrix2d sqrt_info, class default 450/1.5 * I2 (":277-278"), overwritten per run with
  // image_weight * I2 by TrajectoryManager::InitFactorInfo (trajectory_manager.cpp:53-61): a member here, set from
  // clic_options::image_sqrt_info
  double sqrt_info;
  ImageFeatureFactor(int64_t t_i_ns, const Vec3& p_i, int64_t t_j_ns, const Vec3& p_j, const SplineMeta& meta,
                     const Quat& _S_CtoI, const Vec3& _p_CinI, double _sqrt_info = 450. / 1.5)
      : t_i_ns_(t_i_ns), t_j_ns_(t_j_ns), p_i_(p_i), p_j_(p_j), spline_meta_(meta), S_CtoI(_S_CtoI),
        p_CinI(_p_CinI), sqrt_info(_sqrt_info) {
    num_residuals = 2;
    size_t kont_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(1);  // inverse depth
    block_sizes.push_back(1);  // time offset
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R[2];
    RdJacobian J_p[2];
    size_t Knot_offset = 2 * spline_meta_.NumParameters();
    double d_inv = parameters[Knot_offset][0];
    double time_offset_in_ns = parameters[Knot_offset + 1][0];
    int64_t ti_corrected_ns = t_i_ns_ + (int64_t)time_offset_in_ns;
    int64_t tj_corrected_ns = t_j_ns_ + (int64_t)time_offset_in_ns;
    size_t kont_num = spline_meta_.NumParameters();
    size_t R_offset[2] = {0, 0}, P_offset[2] = {0, 0}, seg_idx[2] = {0, 0};
    {
      double u;
      spline_meta_.ComputeSplineIndex(ti_corrected_ns, R_offset[0], u);
      spline_meta_.ComputeSplineIndex(tj_corrected_ns, R_offset[1], u);
      size_t segment0_knot_num = spline_meta_.segments.at(0).NumParameters();
      for (int i = 0; i < 2; ++i) {
        if (R_offset[i] >= segment0_knot_num) {
          seg_idx[i] = 1;
          R_offset[i] = segment0_knot_num;
        } else {
          R_offset[i] = 0;
        }
        P_offset[i] = R_offset[i] + kont_num;
      }
    }
    Vec3 x_ci = p_i_ / d_inv;
    Vec3 p_Ii = so3_rotate(S_CtoI, x_ci) + p_CinI;
    const SegmentMeta& seg_i = spline_meta_.segments.at(seg_idx[0]);
    const SegmentMeta& seg_j = spline_meta_.segments.at(seg_idx[1]);
    Quat S_IitoG = so3_evaluate_Rp(ti_corrected_ns, seg_i, parameters + R_offset[0], jacobians ? &J_R[0] : nullptr);
    Vec3 p_IiinG = rd_evaluate(0, ti_corrected_ns, seg_i, parameters + P_offset[0], jacobians ? &J_p[0] : nullptr);
    Vec3 p_G = so3_rotate(S_IitoG, p_Ii) + p_IiinG;
    Quat S_GtoIj = so3_evaluate_RTp(tj_corrected_ns, seg_j, parameters + R_offset[1], jacobians ? &J_R[1] : nullptr);
    Vec3 p_IjinG = rd_evaluate(0, tj_corrected_ns, seg_j, parameters + P_offset[1], jacobians ? &J_p[1] : nullptr);

    Vec3 gyro_i, gyro_j, vel_i, vel_j;
    if (jacobians && jacobians[Knot_offset + 1]) {
      gyro_i = so3_velocity_body(ti_corrected_ns, seg_i, parameters + R_offset[0], nullptr);
      vel_i = rd_evaluate(1, ti_corrected_ns, seg_i, parameters + P_offset[0], nullptr);
      gyro_j = so3_velocity_body(tj_corrected_ns, seg_j, parameters + R_offset[1], nullptr);
      vel_j = rd_evaluate(1, tj_corrected_ns, seg_j, parameters + P_offset[1], nullptr);
    }
    Quat S_ItoC = qconj(S_CtoI);
    Quat S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
    Vec3 x_j = so3_rotate(S_GtoCj, p_G - p_IjinG) - so3_rotate(S_ItoC, p_CinI);
    double depth_j_inv = 1.0 / x_j[2];
    residuals[0] = x_j[0] * depth_j_inv - p_j_[0];
    residuals[1] = x_j[1] * depth_j_inv - p_j_[1];

    if (jacobians) {
      for (size_t i = 0; i < kont_num; ++i) {
        if (jacobians[i])
          for (int c = 0; c < 8; c++) jacobians[i][c] = 0;
        if (jacobians[i + kont_num])
          for (int c = 0; c < 6; c++) jacobians[i + kont_num][c] = 0;
      }
      // J_v (2x3)
      double J_v[2][3] = {{depth_j_inv, 0, -depth_j_inv * depth_j_inv * x_j[0]},
                          {0, depth_j_inv, -depth_j_inv * depth_j_inv * x_j[1]}};
      auto mul23 = [](const double A[2][3], const Mat3& B, double out[2][3]) {
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 3; c++) out[r][c] = A[r][0] * B(0, c) + A[r][1] * B(1, c) + A[r][2] * B(2, c);
      };
      double jac_lhs_R[2][2][3], jac_lhs_P[2][2][3];
      Mat3 M_GtoCj = so3_matrix(S_GtoCj);
      {
        double negJv[2][3];
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 3; c++) negJv[r][c] = -J_v[r][c];
        Mat3 A = so3_matrix(so3_mul(S_GtoCj, S_IitoG)) * hat(p_Ii);
        mul23(negJv, A, jac_lhs_R[0]);
        mul23(J_v, M_GtoCj, jac_lhs_P[0]);
        Mat3 B = M_GtoCj * hat(p_G - p_IjinG);
        mul23(J_v, B, jac_lhs_R[1]);
        mul23(negJv, M_GtoCj, jac_lhs_P[1]);
      }
      for (int seg = 0; seg < 2; ++seg) {
        size_t pre_idx_R = R_offset[seg] + J_R[seg].start_idx;
        for (size_t i = 0; i < (size_t)N; i++) {
          size_t idx = pre_idx_R + i;
          if (jacobians[idx]) {
            double J_temp[2][3];
            mul23(jac_lhs_R[seg], J_R[seg].d_val_d_knot[i], J_temp);
            for (int r = 0; r < 2; r++)
              for (int c = 0; c < 3; c++) jacobians[idx][r * 4 + c] += sqrt_info * J_temp[r][c];
          }
        }
        size_t pre_idx_P = P_offset[seg] + J_p[seg].start_idx;
        for (size_t i = 0; i < (size_t)N; i++) {
          size_t idx = pre_idx_P + i;
          if (jacobians[idx]) {
            for (int r = 0; r < 2; r++)
              for (int c = 0; c < 3; c++)
                jacobians[idx][r * 3 + c] += sqrt_info * (J_p[seg].d_val_d_knot[i] * jac_lhs_P[seg][r][c]);
          }
        }
      }
      if (jacobians[Knot_offset]) {
        Mat3 M = so3_matrix(so3_mul(so3_mul(S_GtoCj, S_IitoG), S_CtoI));
        Vec3 J_Xm_d = -(M * x_ci) / d_inv;
        for (int r = 0; r < 2; r++)
          jacobians[Knot_offset][r] = sqrt_info * (J_v[r][0] * J_Xm_d[0] + J_v[r][1] * J_Xm_d[1] + J_v[r][2] * J_Xm_d[2]);
      }
      if (jacobians[Knot_offset + 1]) {
        Mat3 Ri_dot = so3_matrix(S_IitoG) * hat(gyro_i);
        Mat3 Rj_dot = so3_matrix(qconj(S_GtoIj)) * hat(gyro_j);
        Vec3 J_tj = so3_matrix(S_ItoC) * transpose(Rj_dot) * (p_G - p_IjinG) - so3_rotate(S_GtoCj, vel_j);
        Vec3 J_ti = M_GtoCj * (Ri_dot * p_Ii + vel_i);
        Vec3 sum = J_ti + J_tj;
        for (int r = 0; r < 2; r++)
          jacobians[Knot_offset + 1][r] = (1e-9 * sqrt_info) * (J_v[r][0] * sum[0] + J_v[r][1] * sum[1] + J_v[r][2] * sum[2]);
      }
    }
    residuals[0] = sqrt_info * residuals[0];
    residuals[1] = sqrt_info * residuals[1];
    return true;
  }
};

// ceres::CauchyLoss(a): rho(s) = b log(1 + s/b), b = a^2 (un-vendored Ceres; loss_function.cc CauchyLoss::Evaluate)
struct CauchyLoss {
  double b_, c_;
  explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
  void Evaluate(double s, double rho[3]) const {
    const double sum = 1.0 + s * c_;
    const double inv = 1.0 / sum;
    rho[0] = b_ * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c_ * (inv * inv);
  }
};

}  // namespace oracle
