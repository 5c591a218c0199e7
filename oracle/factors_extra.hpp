// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference arithmetic (oracle).
//
// The analytic cost functions of SURVEY.md §8f-4 that the reference builds but never adds from its pipeline
// (PreIntegrationFactor has an adder nobody calls; the other five are only constructed by
// src/app/test_analytic_factor.cpp), restated with the same Evaluate(parameters, residuals, jacobians) contract as
// factors.hpp.  Reached through clic_evaluate_factor (every kind) and clic_add_preintegration.
#pragma once
#include "factors.hpp"

namespace oracle {

// leftJacobianSO3 — src/utils/sophus_utils.hpp:238-268
inline Mat3 left_jacobian_so3(const Vec3& phi) {
  double phi_norm2 = sqnorm(phi);
  Mat3 phi_hat = hat(phi);
  Mat3 phi_hat2 = phi_hat * phi_hat;
  Mat3 J = eye3();
  if (phi_norm2 > kEps) {
    double phi_norm = std::sqrt(phi_norm2);
    double phi_norm3 = phi_norm2 * phi_norm;
    J = J + phi_hat * (1 - std::cos(phi_norm)) / phi_norm2;
    J = J + phi_hat2 * (phi_norm - std::sin(phi_norm)) / phi_norm3;
  } else {
    J = J + phi_hat / 2;
    J = J + phi_hat2 / 6;
  }
  return J;
}

// Two measurement times over a one- or two-segment SplineMeta: offsets of each time's segment inside the knot list
// (the block every two-time factor of the reference repeats, e.g. trajectory_value_factor.h:663-681).
struct TwoTimeOffsets {
  size_t R_offset[2] = {0, 0}, P_offset[2] = {0, 0}, seg_idx[2] = {0, 0};
  TwoTimeOffsets(const SplineMeta& meta, int64_t t0, int64_t t1) {
    double u;
    meta.ComputeSplineIndex(t0, R_offset[0], u);
    meta.ComputeSplineIndex(t1, R_offset[1], u);
    size_t kont_num = meta.NumParameters();
    size_t segment0_knot_num = meta.segments.at(0).NumParameters();
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
};

inline void zero_knot_jacobians(double** jacobians, size_t kont_num, int nres) {
  for (size_t i = 0; i < kont_num; ++i) {
    if (jacobians[i])
      for (int c = 0; c < 4 * nres; c++) jacobians[i][c] = 0;
    if (jacobians[i + kont_num])
      for (int c = 0; c < 3 * nres; c++) jacobians[i + kont_num][c] = 0;
  }
}

// IntegrationBase (src/visual_odometry/integration_base.h:36-288) reduced to what the factor reads: the integrated
// deltas, their bias Jacobian and the weight.  sqrt_info = LLT(covariance^-1).matrixL()^T is computed by the CALLER
// (trajectory_value_factor.h:744-748 does it with Eigen inside Evaluate; Eigen is not vendored, so the factorisation is
// handed in and both libraries use the same matrix).
struct PreIntegrationData {
  double sum_dt;
  Vec3 delta_p, delta_v;
  Quat delta_q;
  Vec3 linearized_ba, linearized_bg;
  double jacobian[15][15];
  double sqrt_info[15][15];
  double gravity_norm;  // the global GRAVITY_NORM (parameter_struct.cpp:23: -9.8)
  Mat3 jblock(int r0, int c0) const {
    Mat3 m;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) m(r, c) = jacobian[r0 + r][c0 + c];
    return m;
  }
};
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };  // StateOrder, integration_base.h:20-26

// IntegrationBase::evaluate333 — integration_base.h:207-252
inline void preintegration_residual(const PreIntegrationData& pre, const Vec3& Pi, const Quat& Si, const Vec3& Vi,
                                    const Vec3& Bai, const Vec3& Bgi, const Vec3& Pj, const Quat& Sj, const Vec3& Vj,
                                    const Vec3& Baj, const Vec3& Bgj, double residuals[15]) {
  Mat3 dp_dba = pre.jblock(O_P, O_BA), dp_dbg = pre.jblock(O_P, O_BG);
  Mat3 dq_dbg = pre.jblock(O_R, O_BG);
  Mat3 dv_dba = pre.jblock(O_V, O_BA), dv_dbg = pre.jblock(O_V, O_BG);
  Vec3 dba = Bai - pre.linearized_ba;
  Vec3 dbg = Bgi - pre.linearized_bg;
  Vec3 corrected_delta_p = pre.delta_p + dp_dba * dba + dp_dbg * dbg;
  Quat corrected_delta_q = qmul(pre.delta_q, so3_exp(dq_dbg * dbg));
  Vec3 corrected_delta_v = pre.delta_v + dv_dba * dba + dv_dbg * dbg;
  Quat S_j_to_i = qnormalized(corrected_delta_q);  // SO3T(Eigen::Quaternion): the constructor normalises
  Vec3 res_log = so3_log(so3_mul(so3_mul(qconj(S_j_to_i), qconj(Si)), Sj));
  Vec3 gravity = v3(0, 0, -pre.gravity_norm);
  Quat Si_inv = qconj(Si);
  Vec3 rp = so3_rotate(Si_inv, (0.5 * pre.sum_dt * pre.sum_dt) * gravity + Pj - Pi - Vi * pre.sum_dt) - corrected_delta_p;
  Vec3 rv = so3_rotate(Si_inv, gravity * pre.sum_dt + Vj - Vi) - corrected_delta_v;
  Vec3 rba = Baj - Bai, rbg = Bgj - Bgi;
  for (int r = 0; r < 3; r++) {
    residuals[O_P + r] = rp[r];
    residuals[O_R + r] = res_log[r];
    residuals[O_V + r] = rv[r];
    residuals[O_BA + r] = rba[r];
    residuals[O_BG + r] = rbg[r];
  }
}

// PreIntegrationFactor — trajectory_value_factor.h:618-964
struct PreIntegrationFactor : CostFunction {
  PreIntegrationData pre_;
  int64_t ta_ns_, tb_ns_;
  SplineMeta spline_meta_;
  PreIntegrationFactor(const PreIntegrationData& pre, int64_t ta_ns, int64_t tb_ns, const SplineMeta& meta)
      : pre_(pre), ta_ns_(ta_ns), tb_ns_(tb_ns), spline_meta_(meta) {
    num_residuals = 15;
    size_t knot_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < knot_num; ++i) block_sizes.push_back(3);
    for (int i = 0; i < 4; i++) block_sizes.push_back(3);  // gyro bias(ta), gyro bias(tb), accel bias(ta), accel bias(tb)
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R[2];
    RdJacobian J_p[2], J_v[2];
    size_t Knot_offset = 2 * spline_meta_.NumParameters();
    TwoTimeOffsets o(spline_meta_, ta_ns_, tb_ns_);
    Vec3 gyro_bias_a = load_v(parameters[Knot_offset]), gyro_bias_b = load_v(parameters[Knot_offset + 1]);
    Vec3 accel_bias_a = load_v(parameters[Knot_offset + 2]), accel_bias_b = load_v(parameters[Knot_offset + 3]);
    const SegmentMeta& seg_a = spline_meta_.segments.at(o.seg_idx[0]);
    const SegmentMeta& seg_b = spline_meta_.segments.at(o.seg_idx[1]);
    Quat S_GtoIa = so3_evaluate_RTp(ta_ns_, seg_a, parameters + o.R_offset[0], jacobians ? &J_R[0] : nullptr);
    Vec3 p_IainG = rd_evaluate(0, ta_ns_, seg_a, parameters + o.P_offset[0], jacobians ? &J_p[0] : nullptr);
    Vec3 v_IainG = rd_evaluate(1, ta_ns_, seg_a, parameters + o.P_offset[0], jacobians ? &J_v[0] : nullptr);
    Quat S_IbtoG = so3_evaluate_Rp(tb_ns_, seg_b, parameters + o.R_offset[1], jacobians ? &J_R[1] : nullptr);
    Vec3 p_IbinG = rd_evaluate(0, tb_ns_, seg_b, parameters + o.P_offset[1], jacobians ? &J_p[1] : nullptr);
    Vec3 v_IbinG = rd_evaluate(1, tb_ns_, seg_b, parameters + o.P_offset[1], jacobians ? &J_v[1] : nullptr);
    Quat S_IatoG = qconj(S_GtoIa), S_GtoIb = qconj(S_IbtoG);
    double res[15];
    preintegration_residual(pre_, p_IainG, S_IatoG, v_IainG, accel_bias_a, gyro_bias_a, p_IbinG, S_IbtoG, v_IbinG,
                            accel_bias_b, gyro_bias_b, res);
    Vec3 res_R = v3(res[O_R], res[O_R + 1], res[O_R + 2]);
    for (int r = 0; r < 15; r++) {
      double a = 0;
      for (int c = 0; c < 15; c++) a += pre_.sqrt_info[r][c] * res[c];
      residuals[r] = a;
    }
    if (!jacobians) return true;
    // unweighted Jacobian rows first (15 x width per block), the weight is applied at the end (":940-960")
    std::vector<std::vector<double>> raw(Knot_offset + 4);
    for (size_t i = 0; i < Knot_offset + 4; i++)
      if (jacobians[i]) raw[i].assign((size_t)15 * block_sizes[i], 0.0);
    auto add33 = [&](size_t idx, int row0, const Mat3& m) {
      const int w = block_sizes[idx];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) raw[idx][(size_t)(row0 + r) * w + c] += m(r, c);
    };
    Mat3 M_GtoIa = so3_matrix(S_GtoIa);
    Vec3 gravity = v3(0, 0, -pre_.gravity_norm);
    double dt = pre_.sum_dt;
    Mat3 jac_lhs_dp_dP[2] = {-1.0 * M_GtoIa, M_GtoIa};
    Vec3 r_p_part = p_IbinG - p_IainG - v_IainG * dt + (0.5 * dt * dt) * gravity;
    Mat3 jac_lhs_dp_dR[2] = {M_GtoIa * hat(r_p_part), zero3()};
    Mat3 Jrot = right_jacobian_inv_so3(res_R);
    Mat3 jac_lhs_dR_dR[2] = {-1.0 * (Jrot * so3_matrix(S_GtoIb)), Jrot};
    Mat3 jac_lhs_dv_dP[2] = {-1.0 * M_GtoIa, M_GtoIa};
    Mat3 jac_lhs_dv_dR[2] = {M_GtoIa * hat(v_IbinG - v_IainG + dt * gravity), zero3()};
    Mat3 dp_dba = pre_.jblock(O_P, O_BA), dv_dba = pre_.jblock(O_V, O_BA);
    Mat3 dp_dbg = pre_.jblock(O_P, O_BG), dq_dbg = pre_.jblock(O_R, O_BG), dv_dbg = pre_.jblock(O_V, O_BG);
    Mat3 jac_lhs_dR_dbg;
    {
      Vec3 delta_bg = dq_dbg * (gyro_bias_a - pre_.linearized_bg);
      Mat3 Jbias = left_jacobian_so3(-delta_bg);
      Mat3 Jrot_r = left_jacobian_inv_so3(res_R);
      jac_lhs_dR_dbg = -1.0 * (Jrot_r * Jbias * dq_dbg);
    }
    for (int seg = 0; seg < 2; ++seg) {
      size_t pre_idx_R = o.R_offset[seg] + J_R[seg].start_idx;
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = pre_idx_R + i;
        if (!jacobians[idx]) continue;
        if (seg == 0) add33(idx, O_P, jac_lhs_dp_dR[seg] * J_R[seg].d_val_d_knot[i]);
        add33(idx, O_R, jac_lhs_dR_dR[seg] * J_R[seg].d_val_d_knot[i]);
        if (seg == 0) add33(idx, O_V, jac_lhs_dv_dR[seg] * J_R[seg].d_val_d_knot[i]);
      }
      size_t pre_idx_P = o.P_offset[seg] + J_p[seg].start_idx;
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = pre_idx_P + i;
        if (!jacobians[idx]) continue;
        Mat3 J_temp1 = J_p[seg].d_val_d_knot[i] * jac_lhs_dp_dP[seg];
        if (seg == 0) J_temp1 = J_temp1 + (J_v[seg].d_val_d_knot[i] * jac_lhs_dp_dP[seg]) * pre_.sum_dt;
        add33(idx, O_P, J_temp1);
        add33(idx, O_V, J_v[seg].d_val_d_knot[i] * jac_lhs_dv_dP[seg]);
      }
    }
    if (jacobians[Knot_offset]) {  // gyro bias ta
      add33(Knot_offset, O_P, -1.0 * dp_dbg);
      add33(Knot_offset, O_R, jac_lhs_dR_dbg);
      add33(Knot_offset, O_V, -1.0 * dv_dbg);
      add33(Knot_offset, O_BG, -1.0 * eye3());
    }
    if (jacobians[Knot_offset + 1]) add33(Knot_offset + 1, O_BG, eye3());  // gyro bias tb
    if (jacobians[Knot_offset + 2]) {                                      // accel bias ta
      add33(Knot_offset + 2, O_P, -1.0 * dp_dba);
      add33(Knot_offset + 2, O_V, -1.0 * dv_dba);
      add33(Knot_offset + 2, O_BA, -1.0 * eye3());
    }
    if (jacobians[Knot_offset + 3]) add33(Knot_offset + 3, O_BA, eye3());  // accel bias tb
    for (size_t i = 0; i < Knot_offset + 4; i++) {
      if (!jacobians[i]) continue;
      const int w = block_sizes[i];
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < w; c++) {
          double a = 0;
          for (int k = 0; k < 15; k++) a += pre_.sqrt_info[r][k] * raw[i][(size_t)k * w + c];
          jacobians[i][(size_t)r * w + c] = a;
        }
    }
    return true;
  }
};

// J_v of a normalised-plane projection and 2x3 * 3x3 products, shared by the camera factors below
struct Proj {
  double J_v[2][3];
  Proj(const Vec3& x_j, double depth_j_inv) {
    J_v[0][0] = depth_j_inv; J_v[0][1] = 0; J_v[0][2] = -depth_j_inv * depth_j_inv * x_j[0];
    J_v[1][0] = 0; J_v[1][1] = depth_j_inv; J_v[1][2] = -depth_j_inv * depth_j_inv * x_j[1];
  }
  void mul(const Mat3& B, double out[2][3], double sign = 1.0) const {
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 3; c++) out[r][c] = sign * (J_v[r][0] * B(0, c) + J_v[r][1] * B(1, c) + J_v[r][2] * B(2, c));
  }
  double dot_row(int r, const Vec3& v) const { return J_v[r][0] * v[0] + J_v[r][1] * v[1] + J_v[r][2] * v[2]; }
};
inline void mul23_33(const double A[2][3], const Mat3& B, double out[2][3]) {
  for (int r = 0; r < 2; r++)
    for (int c = 0; c < 3; c++) out[r][c] = A[r][0] * B(0, c) + A[r][1] * B(1, c) + A[r][2] * B(2, c);
}

// Image3D2DFactor — image_feature_factor.h:294-475
struct Image3D2DFactor : CostFunction {
  int64_t t_j_ns_;
  Vec3 p_j_;
  SplineMeta spline_meta_;
  Quat S_CtoI;
  Vec3 p_CinI;
  double sqrt_info;
  Image3D2DFactor(int64_t t_j_ns, const Vec3& p_j, const SplineMeta& meta, const Quat& _S_CtoI, const Vec3& _p_CinI,
                  double _sqrt_info)
      : t_j_ns_(t_j_ns), p_j_(p_j), spline_meta_(meta), S_CtoI(_S_CtoI), p_CinI(_p_CinI), sqrt_info(_sqrt_info) {
    num_residuals = 2;
    size_t kont_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(3);  // p_inG
    block_sizes.push_back(1);  // time offset
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t kont_num = spline_meta_.NumParameters();
    size_t Knot_offset = 2 * kont_num;
    Vec3 p_G = load_v(parameters[Knot_offset]);
    double time_offset_in_ns = parameters[Knot_offset + 1][0];
    int64_t tj_corrected_ns = t_j_ns_ + (int64_t)time_offset_in_ns;
    const SegmentMeta& seg = spline_meta_.segments.at(0);
    Quat S_GtoIj = so3_evaluate_RTp(tj_corrected_ns, seg, parameters, jacobians ? &J_R : nullptr);
    Vec3 p_IjinG = rd_evaluate(0, tj_corrected_ns, seg, parameters + kont_num, jacobians ? &J_p : nullptr);
    Vec3 gyro_j, vel_j;
    if (jacobians && jacobians[Knot_offset + 1]) {
      gyro_j = so3_velocity_body(tj_corrected_ns, seg, parameters, nullptr);
      vel_j = rd_evaluate(1, tj_corrected_ns, seg, parameters + kont_num, nullptr);
    }
    Quat S_ItoC = qconj(S_CtoI);
    Quat S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
    Vec3 x_j = so3_rotate(S_GtoCj, p_G - p_IjinG) - so3_rotate(S_ItoC, p_CinI);
    double depth_j_inv = 1.0 / x_j[2];
    double r0 = x_j[0] * depth_j_inv - p_j_[0], r1 = x_j[1] * depth_j_inv - p_j_[1];
    if (jacobians) {
      zero_knot_jacobians(jacobians, kont_num, 2);
      Proj P(x_j, depth_j_inv);
      Mat3 M_GtoCj = so3_matrix(S_GtoCj);
      double jac_lhs_R[2][3], jac_lhs_P[2][3];
      double JvM[2][3];
      P.mul(M_GtoCj, JvM);
      mul23_33(JvM, hat(p_G - p_IjinG), jac_lhs_R);
      P.mul(M_GtoCj, jac_lhs_P, -1.0);
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = J_R.start_idx + i;
        if (jacobians[idx]) {
          double J_temp[2][3];
          mul23_33(jac_lhs_R, J_R.d_val_d_knot[i], J_temp);
          for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) jacobians[idx][r * 4 + c] += sqrt_info * J_temp[r][c];
        }
      }
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = J_p.start_idx + kont_num + i;
        if (jacobians[idx])
          for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) jacobians[idx][r * 3 + c] += sqrt_info * (J_p.d_val_d_knot[i] * jac_lhs_P[r][c]);
      }
      if (jacobians[Knot_offset]) {
        double J[2][3];
        P.mul(M_GtoCj, J);
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 3; c++) jacobians[Knot_offset][r * 3 + c] = sqrt_info * J[r][c];
      }
      if (jacobians[Knot_offset + 1]) {
        Mat3 Rj_dot = so3_matrix(qconj(S_GtoIj)) * hat(gyro_j);
        Vec3 J_tj = so3_matrix(S_ItoC) * transpose(Rj_dot) * (p_G - p_IjinG) - so3_rotate(S_GtoCj, vel_j);
        for (int r = 0; r < 2; r++) jacobians[Knot_offset + 1][r] = (1e-9 * sqrt_info) * P.dot_row(r, J_tj);
      }
    }
    residuals[0] = sqrt_info * r0;
    residuals[1] = sqrt_info * r1;
    return true;
  }
};

// ImageFeatureOnePoseFactor — image_feature_factor.h:477-656
struct ImageFeatureOnePoseFactor : CostFunction {
  Vec3 p_i_;
  Quat S_IitoG_;
  Vec3 p_IiinG_;
  int64_t t_j_ns_;
  Vec3 p_j_;
  SplineMeta spline_meta_;
  Quat S_CtoI;
  Vec3 p_CinI;
  double sqrt_info;
  ImageFeatureOnePoseFactor(const Vec3& p_i, const Quat& S_IitoG, const Vec3& p_IiinG, int64_t t_j_ns, const Vec3& p_j,
                            const SplineMeta& meta, const Quat& _S_CtoI, const Vec3& _p_CinI, double _sqrt_info)
      : p_i_(p_i), S_IitoG_(S_IitoG), p_IiinG_(p_IiinG), t_j_ns_(t_j_ns), p_j_(p_j), spline_meta_(meta),
        S_CtoI(_S_CtoI), p_CinI(_p_CinI), sqrt_info(_sqrt_info) {
    num_residuals = 2;
    size_t kont_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(1);  // inverse depth
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t kont_num = spline_meta_.NumParameters();
    size_t Knot_offset = 2 * kont_num;
    double d_inv = parameters[Knot_offset][0];
    size_t R_offset, P_offset;
    double u;
    spline_meta_.ComputeSplineIndex(t_j_ns_, R_offset, u);
    P_offset = R_offset + kont_num;
    Vec3 x_ci = p_i_ / d_inv;
    Vec3 p_Ii = so3_rotate(S_CtoI, x_ci) + p_CinI;
    Vec3 p_G = so3_rotate(S_IitoG_, p_Ii) + p_IiinG_;
    // as written (":533-547"): the knot pointers are offset by the INTERVAL index of t_j and the segment view then
    // indexes from there again — only consistent when t_j lies in the first interval of the meta (R_offset = 0), which
    // is how the reference's test builds it (CaculateSplineMeta over [t_j, t_j])
    const SegmentMeta& seg = spline_meta_.segments.at(0);
    Quat S_GtoIj = so3_evaluate_RTp(t_j_ns_, seg, parameters + R_offset, jacobians ? &J_R : nullptr);
    Vec3 p_IjinG = rd_evaluate(0, t_j_ns_, seg, parameters + P_offset, jacobians ? &J_p : nullptr);
    Quat S_ItoC = qconj(S_CtoI);
    Quat S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
    Vec3 x_j = so3_rotate(S_GtoCj, p_G - p_IjinG) - so3_rotate(S_ItoC, p_CinI);
    double depth_j_inv = 1.0 / x_j[2];
    double r0 = x_j[0] * depth_j_inv - p_j_[0], r1 = x_j[1] * depth_j_inv - p_j_[1];
    if (jacobians) {
      zero_knot_jacobians(jacobians, kont_num, 2);
      Proj P(x_j, depth_j_inv);
      Mat3 M_GtoCj = so3_matrix(S_GtoCj);
      double jac_lhs_R[2][3], jac_lhs_P[2][3];
      double JvM[2][3];
      P.mul(M_GtoCj, JvM);
      mul23_33(JvM, hat(p_G - p_IjinG), jac_lhs_R);
      P.mul(M_GtoCj, jac_lhs_P, -1.0);
      size_t pre_idx_R = R_offset + J_R.start_idx;
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = pre_idx_R + i;
        if (jacobians[idx]) {
          double J_temp[2][3];
          mul23_33(jac_lhs_R, J_R.d_val_d_knot[i], J_temp);
          for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) jacobians[idx][r * 4 + c] += sqrt_info * J_temp[r][c];
        }
      }
      size_t pre_idx_P = P_offset + J_p.start_idx;
      for (size_t i = 0; i < (size_t)N; i++) {
        size_t idx = pre_idx_P + i;
        if (jacobians[idx])
          for (int r = 0; r < 2; r++)
            for (int c = 0; c < 3; c++) jacobians[idx][r * 3 + c] += sqrt_info * (J_p.d_val_d_knot[i] * jac_lhs_P[r][c]);
      }
      if (jacobians[Knot_offset]) {
        Mat3 M = so3_matrix(so3_mul(so3_mul(S_GtoCj, S_IitoG_), S_CtoI));
        Vec3 J_Xm_d = -(M * x_ci) / d_inv;
        for (int r = 0; r < 2; r++) jacobians[Knot_offset][r] = sqrt_info * P.dot_row(r, J_Xm_d);
      }
    }
    residuals[0] = sqrt_info * r0;
    residuals[1] = sqrt_info * r1;
    return true;
  }
};

// EpipolarFactor — image_feature_factor.h:709-842
struct EpipolarFactor : CostFunction {
  int64_t t_i_ns_;
  Vec3 x_i_, x_k_;
  Quat S_GtoCk_;
  Vec3 p_CkinG_;
  SplineMeta spline_meta_;
  double weight_;
  Quat S_CtoI;
  Vec3 p_CinI;
  EpipolarFactor(int64_t t_i_ns, const Vec3& x_i, const Vec3& x_k, const Quat& S_GtoCk, const Vec3& p_CkinG,
                 const SplineMeta& meta, double weight, const Quat& _S_CtoI, const Vec3& _p_CinI)
      : t_i_ns_(t_i_ns), x_i_(x_i), x_k_(x_k), S_GtoCk_(S_GtoCk), p_CkinG_(p_CkinG), spline_meta_(meta),
        weight_(weight), S_CtoI(_S_CtoI), p_CinI(_p_CinI) {
    num_residuals = 1;
    size_t kont_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t kont_num = spline_meta_.NumParameters();
    const SegmentMeta& seg = spline_meta_.segments[0];
    Quat S_IitoG = so3_evaluate_Rp(t_i_ns_, seg, parameters, jacobians ? &J_R : nullptr);
    Vec3 p_IiinG = rd_evaluate(0, t_i_ns_, seg, parameters + kont_num, jacobians ? &J_p : nullptr);
    Vec3 Rxi = so3_rotate(so3_mul(so3_mul(S_GtoCk_, S_IitoG), S_CtoI), x_i_);
    Mat3 t_hat = hat(so3_rotate(S_GtoCk_, so3_rotate(S_IitoG, p_CinI) + p_IiinG - p_CkinG_));
    residuals[0] = dot(rowmul(x_k_, t_hat), Rxi);
    residuals[0] *= weight_;
    if (jacobians) {
      zero_knot_jacobians(jacobians, kont_num, 1);
      Vec3 jac_lhs = rowmul(rowmul(x_k_, hat(Rxi)), so3_matrix(S_GtoCk_));
      Vec3 jac_lhs_P = -jac_lhs;
      Vec3 jac_lhs_R = rowmul(rowmul(rowmul(-x_k_, t_hat), so3_matrix(so3_mul(S_GtoCk_, S_IitoG))), hat(so3_rotate(S_CtoI, x_i_)));
      jac_lhs_R = jac_lhs_R + rowmul(rowmul(jac_lhs, so3_matrix(S_IitoG)), hat(p_CinI));
      // as written (":793-822"): d_val_d_knot[i] is indexed by the BLOCK index, i.e. the meta must hold exactly the four
      // knots of the interval (the reference's test builds it over [t_i, t_i]); other knots are skipped here
      for (size_t i = 0; i < kont_num && i < (size_t)N; i++) {
        if (jacobians[i]) {
          Vec3 j = rowmul(jac_lhs_R, J_R.d_val_d_knot[i]);
          for (int c = 0; c < 3; c++) jacobians[i][c] = weight_ * j[c];
        }
        if (jacobians[kont_num + i])
          for (int c = 0; c < 3; c++) jacobians[kont_num + i][c] = weight_ * (J_p.d_val_d_knot[i] * jac_lhs_P[c]);
      }
    }
    return true;
  }
};

// point-to-plane / point-to-line distance of p and d(distance)/dp — lidar_feature_factor.h:237-250, 442-455
inline double loam_distance(const PointCorrespondence& pc, const Vec3& p, Vec3* J_pi) {
  double r;
  if (pc.geo_type == 1) {
    r = dot(p, v3(pc.geo_plane[0], pc.geo_plane[1], pc.geo_plane[2])) + pc.geo_plane[3];
    *J_pi = v3(pc.geo_plane[0], pc.geo_plane[1], pc.geo_plane[2]);
  } else {
    Vec3 dist_vec = cross(p - pc.geo_point, pc.geo_normal);
    r = norm(dist_vec);
    *J_pi = rowmul(-1.0 * dist_vec / r, hat(pc.geo_normal));
  }
  return r;
}

// LoamFeatureOptMapPoseFactor — lidar_feature_factor.h:174-348
struct LoamFeatureOptMapPoseFactor : CostFunction {
  int64_t t_point_ns_;
  PointCorrespondence pc_;
  SegmentMeta meta_;
  double weight_;
  Quat S_LtoI;
  Vec3 p_LinI;
  LoamFeatureOptMapPoseFactor(int64_t t_point_ns, const PointCorrespondence& pc, const SegmentMeta& meta, double weight,
                              const Quat& _S_LtoI, const Vec3& _p_LinI)
      : t_point_ns_(t_point_ns), pc_(pc), meta_(meta), weight_(weight), S_LtoI(_S_LtoI), p_LinI(_p_LinI) {
    num_residuals = 1;
    size_t kont_num = meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
    block_sizes.push_back(4);  // map rotation
    block_sizes.push_back(3);  // map position
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R;
    RdJacobian J_p;
    size_t kont_num = meta_.NumParameters();
    Quat S_IktoG = so3_evaluate_Rp(t_point_ns_, meta_, parameters, jacobians ? &J_R : nullptr);
    Vec3 p_IkinG = rd_evaluate(0, t_point_ns_, meta_, parameters + kont_num, jacobians ? &J_p : nullptr);
    Quat S_ImtoG = load_q(parameters[kont_num * 2]);
    Vec3 p_IminG = load_v(parameters[kont_num * 2 + 1]);
    Quat S_ItoL = qconj(S_LtoI);
    Quat S_GtoIm = qconj(S_ImtoG);
    Quat S_GtoM = so3_mul(S_ItoL, S_GtoIm);
    Vec3 p_GinM = so3_rotate(S_ItoL, so3_rotate(S_GtoIm, -p_IminG) - p_LinI);
    Vec3 p_IK = so3_rotate(S_LtoI, pc_.point) + p_LinI;
    Vec3 p_G = so3_rotate(S_IktoG, p_IK) + p_IkinG;
    Vec3 p_M = so3_rotate(S_GtoM, p_G) + p_GinM;
    Vec3 J_pi;
    residuals[0] = loam_distance(pc_, p_M, &J_pi);
    residuals[0] *= weight_;
    if (!jacobians) return true;
    zero_knot_jacobians(jacobians, kont_num, 1);
    Mat3 M_GtoM = so3_matrix(S_GtoM);
    Mat3 J_Xm_R = -1.0 * (M_GtoM * so3_matrix(S_IktoG) * hat(p_IK));
    Vec3 jac_lhs_R = rowmul(J_pi, J_Xm_R);
    Vec3 jac_lhs_P = rowmul(J_pi, M_GtoM);
    // (":284-311": d_val_d_knot[i] indexed by the block index — the segment holds the four knots of the interval)
    for (size_t i = 0; i < kont_num && i < (size_t)N; i++) {
      if (jacobians[i]) {
        Vec3 j = rowmul(jac_lhs_R, J_R.d_val_d_knot[i]);
        for (int c = 0; c < 3; c++) jacobians[i][c] = weight_ * j[c];
      }
      if (jacobians[kont_num + i])
        for (int c = 0; c < 3; c++) jacobians[kont_num + i][c] = weight_ * (J_p.d_val_d_knot[i] * jac_lhs_P[c]);
    }
    if (jacobians[kont_num * 2]) {
      Vec3 p_Im = so3_rotate(S_GtoIm, p_G - p_IminG);
      Vec3 j = rowmul(J_pi, so3_matrix(S_ItoL) * hat(p_Im));
      for (int c = 0; c < 3; c++) jacobians[kont_num * 2][c] = weight_ * j[c];
      jacobians[kont_num * 2][3] = 0;
    }
    if (jacobians[kont_num * 2 + 1]) {
      Vec3 j = -rowmul(J_pi, so3_matrix(so3_mul(S_ItoL, S_GtoIm)));
      for (int c = 0; c < 3; c++) jacobians[kont_num * 2 + 1][c] = weight_ * j[c];
    }
    return true;
  }
};

// RalativeLoamFeatureFactor — lidar_feature_factor.h:350-553
struct RalativeLoamFeatureFactor : CostFunction {
  int64_t t_point_ns_, t_map_ns_;
  PointCorrespondence pc_;
  SplineMeta spline_meta_;
  double weight_;
  Quat S_LtoI;
  Vec3 p_LinI;
  RalativeLoamFeatureFactor(int64_t t_point_ns, int64_t t_map_ns, const PointCorrespondence& pc, const SplineMeta& meta,
                            double weight, const Quat& _S_LtoI, const Vec3& _p_LinI)
      : t_point_ns_(t_point_ns), t_map_ns_(t_map_ns), pc_(pc), spline_meta_(meta), weight_(weight), S_LtoI(_S_LtoI),
        p_LinI(_p_LinI) {
    num_residuals = 1;
    size_t kont_num = spline_meta_.NumParameters();
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(4);
    for (size_t i = 0; i < kont_num; ++i) block_sizes.push_back(3);
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    So3Jacobian J_R[2];
    RdJacobian J_p[2];
    size_t kont_num = spline_meta_.NumParameters();
    // ":386-387": the two times pass through doubles (exact below 2^53 ns)
    int64_t t_i = (int64_t)(double)t_point_ns_, t_j = (int64_t)(double)t_map_ns_;
    TwoTimeOffsets o(spline_meta_, t_i, t_j);
    Vec3 p_Ii = so3_rotate(S_LtoI, pc_.point) + p_LinI;
    const SegmentMeta& seg_i = spline_meta_.segments.at(o.seg_idx[0]);
    const SegmentMeta& seg_j = spline_meta_.segments.at(o.seg_idx[1]);
    Quat S_IitoG = so3_evaluate_Rp(t_i, seg_i, parameters + o.R_offset[0], jacobians ? &J_R[0] : nullptr);
    Vec3 p_IiinG = rd_evaluate(0, t_i, seg_i, parameters + o.P_offset[0], jacobians ? &J_p[0] : nullptr);
    Vec3 p_G = so3_rotate(S_IitoG, p_Ii) + p_IiinG;
    Quat S_GtoIj = so3_evaluate_RTp(t_j, seg_j, parameters + o.R_offset[1], jacobians ? &J_R[1] : nullptr);
    Vec3 p_IjinG = rd_evaluate(0, t_j, seg_j, parameters + o.P_offset[1], jacobians ? &J_p[1] : nullptr);
    Quat S_ItoC = qconj(S_LtoI);
    Quat S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
    Vec3 x_j = so3_rotate(S_GtoCj, p_G - p_IjinG) - so3_rotate(S_ItoC, p_LinI);
    Vec3 J_pi;
    residuals[0] = loam_distance(pc_, x_j, &J_pi);
    residuals[0] *= weight_;
    if (jacobians) {
      zero_knot_jacobians(jacobians, kont_num, 1);
      Mat3 M_GtoCj = so3_matrix(S_GtoCj);
      Vec3 jac_lhs_R[2], jac_lhs_P[2];
      jac_lhs_R[0] = rowmul(rowmul(-J_pi, so3_matrix(so3_mul(S_GtoCj, S_IitoG))), hat(p_Ii));
      jac_lhs_P[0] = rowmul(J_pi, M_GtoCj);
      jac_lhs_R[1] = rowmul(rowmul(J_pi, M_GtoCj), hat(p_G - p_IjinG));
      jac_lhs_P[1] = -rowmul(J_pi, M_GtoCj);
      for (int seg = 0; seg < 2; ++seg) {
        size_t pre_idx_R = o.R_offset[seg] + J_R[seg].start_idx;
        for (size_t i = 0; i < (size_t)N; i++) {
          size_t idx = pre_idx_R + i;
          if (jacobians[idx]) {
            Vec3 j = rowmul(jac_lhs_R[seg], J_R[seg].d_val_d_knot[i]);
            for (int c = 0; c < 3; c++) jacobians[idx][c] += weight_ * j[c];
          }
        }
        size_t pre_idx_P = o.P_offset[seg] + J_p[seg].start_idx;
        for (size_t i = 0; i < (size_t)N; i++) {
          size_t idx = pre_idx_P + i;
          if (jacobians[idx])
            for (int c = 0; c < 3; c++) jacobians[idx][c] += weight_ * (J_p[seg].d_val_d_knot[i] * jac_lhs_P[seg][c]);
        }
      }
    }
    return true;
  }
};

}  // namespace oracle
