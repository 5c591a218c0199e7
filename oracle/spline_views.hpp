// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference arithmetic (oracle).
//
// Uniform cubic B-spline indexing + the analytic "spline views" (value and d/dknot) of
//   src/spline/spline_common.h, src/spline/spline_segment.h,
//   src/estimator/factor/analytic_diff/{so3,rd,split}_spline_view.h
// restated without Eigen, as written in the reference (no hoisting, same loop structure), so that
// the restructured CUDA kernels are checked against an independent formulation.
#pragma once
#include <array>
#include <cstddef>
#include <utility>
#include <vector>

#include "so3_math.hpp"

namespace oracle {

constexpr int N = 4;        // SplineOrder, src/spline/spline_common.h:46
constexpr int DEG = N - 1;
constexpr double S_TO_NS = 1e9;   // src/spline/spline_segment.h:34
constexpr double NS_TO_S = 1e-9;  // src/spline/spline_segment.h:33

// C_n_k — src/spline/spline_common.h:55-65
constexpr uint64_t C_n_k(uint64_t n, uint64_t k) {
  if (k > n) return 0;
  uint64_t r = 1;
  for (uint64_t d = 1; d <= k; ++d) {
    r *= n--;
    r /= d;
  }
  return r;
}

struct Mat4 {
  double m[N][N];
};

// computeBlendingMatrix<N,double,Cumulative> — src/spline/spline_common.h:72-103
inline Mat4 compute_blending_matrix(bool cumulative) {
  Mat4 M;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) M.m[i][j] = 0;
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < N; ++j) {
      double sum = 0;
      for (int s = j; s < N; ++s) {
        sum += std::pow(-1.0, s - j) * C_n_k(N, s - j) * std::pow(N - s - 1.0, N - 1.0 - i);
      }
      M.m[j][i] = C_n_k(N - 1, N - 1 - i) * sum;
    }
  }
  if (cumulative) {
    for (int i = 0; i < N; i++)
      for (int j = i + 1; j < N; j++)
        for (int c = 0; c < N; c++) M.m[i][c] += M.m[j][c];
  }
  uint64_t factorial = 1;
  for (int i = 2; i < N; ++i) factorial *= i;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) M.m[i][j] /= factorial;
  return M;
}

// computeBaseCoefficients — src/spline/spline_common.h:122-138
inline Mat4 compute_base_coefficients() {
  Mat4 B;
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) B.m[i][j] = 0;
  for (int j = 0; j < N; j++) B.m[0][j] = 1;
  int order = DEG;
  for (int n = 1; n < N; n++) {
    for (int i = DEG - order; i < N; i++) B.m[n][i] = (order - DEG + i) * B.m[n - 1][i];
    order--;
  }
  return B;
}

inline const Mat4& blend_cumulative() {
  static const Mat4 M = compute_blending_matrix(true);
  return M;
}
inline const Mat4& blend_plain() {
  static const Mat4 M = compute_blending_matrix(false);
  return M;
}
inline const Mat4& base_coeffs() {
  static const Mat4 M = compute_base_coefficients();
  return M;
}

// baseCoeffsWithTime<Derivative> — so3_spline_view.h:450-468 / rd_spline_view.h:113-131
inline void base_coeffs_with_time(int derivative, double p[N], double t) {
  for (int i = 0; i < N; i++) p[i] = 0;
  if (derivative < N) {
    const Mat4& B = base_coeffs();
    p[derivative] = B.m[derivative][derivative];
    double _t = t;
    for (int j = derivative + 1; j < N; j++) {
      p[j] = B.m[derivative][j] * _t;
      _t = _t * t;
    }
  }
}
inline void matvec4(const Mat4& M, const double p[N], double out[N]) {
  for (int i = 0; i < N; i++) out[i] = M.m[i][0] * p[0] + M.m[i][1] * p[1] + M.m[i][2] * p[2] + M.m[i][3] * p[3];
}

// scalar * M * p with Eigen's left-to-right evaluation: (scalar*M)*p
inline void scaled_matvec4(double sc, const Mat4& M, const double p[N], double out[N]) {
  for (int i = 0; i < N; i++)
    out[i] = (sc * M.m[i][0]) * p[0] + (sc * M.m[i][1]) * p[1] + (sc * M.m[i][2]) * p[2] + (sc * M.m[i][3]) * p[3];
}

// SplineSegmentMeta<4> — src/spline/spline_segment.h:40-118
struct SegmentMeta {
  int64_t t0_ns;
  int64_t dt_ns;
  size_t n;
  double pow_inv_dt[N];
  SegmentMeta() : t0_ns(0), dt_ns(1), n(0) {}
  SegmentMeta(int64_t _t0, int64_t _dt, size_t _n = 0) : t0_ns(_t0), dt_ns(_dt), n(_n) {
    pow_inv_dt[0] = 1.0;
    pow_inv_dt[1] = S_TO_NS / dt_ns;
    for (size_t i = 2; i < (size_t)N; i++) pow_inv_dt[i] = pow_inv_dt[i - 1] * pow_inv_dt[1];
  }
  size_t NumParameters() const { return n; }
  int64_t MinTimeNs() const { return t0_ns; }
  int64_t MaxTimeNs() const { return t0_ns + (int64_t)(n - DEG) * dt_ns; }
  // computeTIndexNs(int64) -> (u, s) — spline_segment.h:67-84 (the assert becomes a bool here)
  bool computeTIndexNs(int64_t time_ns, size_t& s, double& u) const {
    if (time_ns < MinTimeNs() || time_ns >= MaxTimeNs()) return false;
    int64_t st_ns = (time_ns - t0_ns);
    s = (size_t)(st_ns / dt_ns);
    u = double(st_ns % dt_ns) / double(dt_ns);
    return true;
  }
};

// SplineMeta<4> — src/spline/spline_segment.h:120-168
struct SplineMeta {
  std::vector<SegmentMeta> segments;
  size_t NumParameters() const {
    size_t n = 0;
    for (auto& s : segments) n += s.NumParameters();
    return n;
  }
  bool ComputeSplineIndex(int64_t time_ns, size_t& idx, double& u) const {
    idx = 0;
    for (auto const& seg : segments) {
      size_t s = 0;
      if (seg.computeTIndexNs(time_ns, s, u)) {
        idx += s;
        return true;
      } else {
        idx += seg.NumParameters();
      }
    }
    return false;
  }
};

inline Quat load_q(const double* p) { return Quat{p[0], p[1], p[2], p[3]}; }
inline Vec3 load_v(const double* p) { return v3(p[0], p[1], p[2]); }

struct So3Jacobian {  // So3SplineView::JacobianStruct — so3_spline_view.h:49-52
  size_t start_idx;
  std::array<Mat3, N> d_val_d_knot;
};
struct RdJacobian {  // RdSplineView::JacobianStruct — rd_spline_view.h:43-47
  size_t start_idx;
  std::array<double, N> d_val_d_knot;
};

// So3SplineView::EvaluateRp — so3_spline_view.h:127-183
inline Quat so3_evaluate_Rp(int64_t time_ns, const SegmentMeta& meta, double const* const* knots, So3Jacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  if (J) J->start_idx = s;

  Quat A_accum_inv{0, 0, 0, 1};
  Mat3 A_post_inv[DEG + 1];
  Mat3 Jr_inv_delta[DEG], Jr_kdelta[DEG];
  A_post_inv[DEG] = so3_matrix(A_accum_inv);
  for (int i = DEG - 1; i >= 0; i--) {
    Quat R0 = load_q(knots[s + i]);
    Quat R1 = load_q(knots[s + i + 1]);
    Vec3 delta = so3_log(so3_mul(qconj(R0), R1));
    Vec3 kdelta = delta * coeff[i + 1];
    A_accum_inv = so3_mul(A_accum_inv, so3_exp(-kdelta));
    if (J) {
      Jr_inv_delta[i] = right_jacobian_inv_so3(delta);
      Jr_kdelta[i] = right_jacobian_so3(kdelta);
      A_post_inv[i] = so3_matrix(A_accum_inv);
    }
  }
  Quat Ri = load_q(knots[s]);
  Quat res = so3_mul(Ri, qconj(A_accum_inv));
  if (J) {
    Mat3 J_helper = A_post_inv[0];
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < DEG; i++) {
      J_helper = coeff[i + 1] * A_post_inv[i + 1] * Jr_kdelta[i];
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - (J_helper * transpose(Jr_inv_delta[i]));
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// So3SplineView::EvaluateRTp — so3_spline_view.h:193-256.  Returns R(t)^T.
inline Quat so3_evaluate_RTp(int64_t time_ns, const SegmentMeta& meta, double const* const* knots, So3Jacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  if (J) J->start_idx = s;
  Quat Ri = load_q(knots[s]);
  Quat Si_A_pre[DEG + 1];
  Mat3 Ri_A_pre[DEG + 1];
  Mat3 Jr_inv_delta[DEG], Jr_kdelta[DEG];
  Si_A_pre[0] = Ri;
  for (int i = 0; i < DEG; i++) {
    Quat R0 = load_q(knots[s + i]);
    Quat R1 = load_q(knots[s + i + 1]);
    Vec3 delta = so3_log(so3_mul(qconj(R0), R1));
    Vec3 kdelta = delta * coeff[i + 1];
    Si_A_pre[i + 1] = so3_mul(Si_A_pre[i], so3_exp(kdelta));
    if (J) {
      Jr_inv_delta[i] = right_jacobian_inv_so3(delta);
      Jr_kdelta[i] = right_jacobian_so3(-kdelta);
    }
  }
  for (int i = 0; i < DEG + 1; i++) Ri_A_pre[i] = so3_matrix(Si_A_pre[i]);
  Quat res = qconj(Si_A_pre[DEG]);
  if (J) {
    Mat3 J_helper = Ri_A_pre[0];
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < DEG; i++) {
      J_helper = coeff[i + 1] * Ri_A_pre[i] * Jr_kdelta[i];
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - (J_helper * transpose(Jr_inv_delta[i]));
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// So3SplineView::EvaluateRotation — so3_spline_view.h:65-117.  d_val_d_knot[k] = d log-right-perturbation of R(t) /
// d(theta_k), built from the forward products R_tmp[i] = R_s prod_{j<=i} exp(k delta_j).
inline Quat so3_evaluate_rotation(int64_t time_ns, const SegmentMeta& meta, double const* const* knots, So3Jacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  Quat res = load_q(knots[s]);
  Mat3 J_helper = eye3();
  if (J) {
    J->start_idx = s;
    J_helper = so3_matrix(res);
  }
  Mat3 R_tmp[DEG];
  Mat3 Jr_inv_delta[DEG], Jr_kdelta[DEG];
  for (int i = 0; i < DEG; i++) {
    Quat p0 = load_q(knots[s + i]);
    Quat p1 = load_q(knots[s + i + 1]);
    Vec3 delta = so3_log(so3_mul(qconj(p0), p1));
    Vec3 kdelta = delta * coeff[i + 1];
    res = so3_mul(res, so3_exp(kdelta));
    Jr_inv_delta[i] = right_jacobian_inv_so3(delta);
    Jr_kdelta[i] = right_jacobian_so3(kdelta);
    R_tmp[i] = so3_matrix(res);
  }
  if (J) {
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < DEG; i++) {
      J_helper = coeff[i + 1] * R_tmp[i] * Jr_kdelta[i];
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - (J_helper * transpose(Jr_inv_delta[i]));
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// So3SplineView::EvaluateLogRT — so3_spline_view.h:263-321.  Returns R(t)^T, re-normalised through a rotation matrix
// and Eigen's matrix -> quaternion conversion exactly as written (":303-306").  One normalisation of the reference's
// behaviour: it fills A_post_inv[] only when J != nullptr but reads A_post_inv[0] for the value regardless (":290-303"),
// i.e. a cost-only evaluation reads an uninitialised matrix; here it is always filled.
inline Quat so3_evaluate_logRT(int64_t time_ns, const SegmentMeta& meta, double const* const* knots, So3Jacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  if (J) J->start_idx = s;
  Quat Ri = load_q(knots[s]);
  Quat A_accum_inv{0, 0, 0, 1};
  Mat3 A_post_inv[DEG + 1];
  Mat3 Jr_inv_delta[DEG], Jr_kdelta[DEG];
  A_post_inv[DEG] = eye3();
  for (int i = DEG - 1; i >= 0; i--) {
    Quat R0 = load_q(knots[s + i]);
    Quat R1 = load_q(knots[s + i + 1]);
    Vec3 delta = so3_log(so3_mul(qconj(R0), R1));
    Vec3 kdelta = delta * coeff[i + 1];
    A_accum_inv = so3_mul(A_accum_inv, so3_exp(-kdelta));
    Jr_inv_delta[i] = right_jacobian_inv_so3(delta);
    Jr_kdelta[i] = right_jacobian_so3(-kdelta);
    A_post_inv[i] = so3_matrix(A_accum_inv);
  }
  Mat3 R_T = A_post_inv[0] * so3_matrix(qconj(Ri));
  Quat q_t = qnormalized(quat_from_matrix_eigen(R_T));
  // SO3(q_t.toRotationMatrix()): Sophus converts the matrix back to a unit quaternion (so3.hpp:165-176)
  Quat res = qnormalized(quat_from_matrix_eigen(so3_matrix(q_t)));
  if (J) {
    Mat3 J_helper = A_post_inv[0];
    J->d_val_d_knot[0] = J_helper;
    for (int i = 0; i < DEG; i++) {
      J_helper = coeff[i + 1] * A_post_inv[i] * Jr_kdelta[i];
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - (J_helper * transpose(Jr_inv_delta[i]));
      J->d_val_d_knot[i + 1] = J_helper * Jr_inv_delta[i];
    }
  }
  return res;
}

// So3SplineView::VelocityBody — so3_spline_view.h:330-394 (J == nullptr is the only use on the hot path,
// image_feature_factor.h:144-158; the Jacobian part is restated too for completeness).
inline Vec3 so3_velocity_body(int64_t time_ns, const SegmentMeta& meta, double const* const* knots, So3Jacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N], dcoeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  base_coeffs_with_time(1, p, u);
  scaled_matvec4(meta.pow_inv_dt[1], blend_cumulative(), p, dcoeff);
  Vec3 delta_vec[DEG];
  Mat3 R_tmp[DEG];
  Quat accum{0, 0, 0, 1};
  Quat exp_k_delta[DEG];
  Mat3 Jr_delta_inv[DEG], Jr_kdelta[DEG];
  for (int i = DEG - 1; i >= 0; i--) {
    Quat p0 = load_q(knots[s + i]);
    Quat p1 = load_q(knots[s + i + 1]);
    Quat r01 = so3_mul(qconj(p0), p1);
    delta_vec[i] = so3_log(r01);
    Jr_delta_inv[i] = right_jacobian_inv_so3(delta_vec[i]);
    Jr_delta_inv[i] = Jr_delta_inv[i] * so3_matrix(qconj(p1));
    Vec3 k_delta = coeff[i + 1] * delta_vec[i];
    Jr_kdelta[i] = right_jacobian_so3(-k_delta);
    R_tmp[i] = so3_matrix(accum);
    exp_k_delta[i] = so3_exp(-k_delta);
    accum = so3_mul(accum, exp_k_delta[i]);
  }
  Mat3 d_vel_d_delta[DEG];
  d_vel_d_delta[0] = dcoeff[1] * R_tmp[0] * Jr_delta_inv[0];
  Vec3 rot_vel = delta_vec[0] * dcoeff[1];
  for (int i = 1; i < DEG; i++) {
    d_vel_d_delta[i] = R_tmp[i - 1] * hat(rot_vel) * Jr_kdelta[i] * coeff[i + 1] + R_tmp[i] * dcoeff[i + 1];
    d_vel_d_delta[i] = d_vel_d_delta[i] * Jr_delta_inv[i];
    rot_vel = so3_rotate(exp_k_delta[i], rot_vel) + delta_vec[i] * dcoeff[i + 1];
  }
  if (J) {
    J->start_idx = s;
    for (int i = 0; i < N; i++) J->d_val_d_knot[i] = zero3();
    for (int i = 0; i < DEG; i++) {
      J->d_val_d_knot[i] = J->d_val_d_knot[i] - d_vel_d_delta[i];
      J->d_val_d_knot[i + 1] = J->d_val_d_knot[i + 1] + d_vel_d_delta[i];
    }
  }
  return rot_vel;
}

// So3SplineView::accelerationBody — so3_spline_view.h:396-439
inline Vec3 so3_acceleration_body(int64_t time_ns, const SegmentMeta& meta, double const* const* knots) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], coeff[N], dcoeff[N], ddcoeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  base_coeffs_with_time(1, p, u);
  scaled_matvec4(meta.pow_inv_dt[1], blend_cumulative(), p, dcoeff);
  base_coeffs_with_time(2, p, u);
  scaled_matvec4(meta.pow_inv_dt[2], blend_cumulative(), p, ddcoeff);
  Vec3 rot_vel = v3(0, 0, 0), rot_accel = v3(0, 0, 0);
  for (int i = 0; i < DEG; i++) {
    Quat p0 = load_q(knots[s + i]);
    Quat p1 = load_q(knots[s + i + 1]);
    Quat r01 = so3_mul(qconj(p0), p1);
    Vec3 delta = so3_log(r01);
    Quat rot = so3_exp(-delta * coeff[i + 1]);
    rot_vel = so3_rotate(rot, rot_vel);
    Vec3 vel_current = dcoeff[i + 1] * delta;
    rot_vel = rot_vel + vel_current;
    rot_accel = so3_rotate(rot, rot_accel);
    rot_accel = rot_accel + (ddcoeff[i + 1] * delta + cross(rot_vel, vel_current));
  }
  return rot_accel;
}

// RdSplineView::evaluate<Derivative> — rd_spline_view.h:59-86
inline Vec3 rd_evaluate(int derivative, int64_t time_ns, const SegmentMeta& meta, double const* const* knots,
                        RdJacobian* J) {
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  double p[N], bp[N], coeff[N];
  base_coeffs_with_time(derivative, p, u);
  matvec4(blend_plain(), p, bp);
  for (int i = 0; i < N; i++) coeff[i] = meta.pow_inv_dt[derivative] * bp[i];
  Vec3 res = v3(0, 0, 0);
  for (int i = 0; i < N; i++) {
    Vec3 kp = load_v(knots[s + i]);
    res = res + coeff[i] * kp;
    if (J) J->d_val_d_knot[i] = coeff[i];
  }
  if (J) J->start_idx = s;
  return res;
}

// SplitSpineView::Evaluate — split_spline_view.h:56-188
struct SplineIMUData {
  int64_t time_ns;
  Vec3 gyro;
  Vec3 accel;
  Quat R_inv;
  size_t start_idx;
};
inline SplineIMUData split_evaluate(int64_t time_ns, const SegmentMeta& meta, double const* const* knots,
                                    const Vec3& gravity, So3Jacobian* J_rot_w, So3Jacobian* J_rot_a,
                                    RdJacobian* J_pos) {
  SplineIMUData d;
  d.time_ns = time_ns;
  size_t s = 0;
  double u = 0;
  meta.computeTIndexNs(time_ns, s, u);
  size_t R_knot_offset = s;
  size_t P_knot_offset = s + meta.NumParameters();
  d.start_idx = s;

  double Up[N], lambda_a[N];
  base_coeffs_with_time(2, Up, u);
  scaled_matvec4(meta.pow_inv_dt[2], blend_plain(), Up, lambda_a);
  double Ur[N], lambda_R[N];
  base_coeffs_with_time(0, Ur, u);
  matvec4(blend_cumulative(), Ur, lambda_R);
  double Uw[N], lambda_w[N];
  base_coeffs_with_time(1, Uw, u);
  scaled_matvec4(meta.pow_inv_dt[1], blend_cumulative(), Uw, lambda_w);

  Vec3 accelerate = v3(0, 0, 0);
  if (J_pos) J_pos->start_idx = s;
  for (int i = 0; i < N; i++) {
    Vec3 kp = load_v(knots[P_knot_offset + i]);
    accelerate = accelerate + lambda_a[i] * kp;
    if (J_pos) J_pos->d_val_d_knot[i] = lambda_a[i];
  }

  Vec3 d_vec[DEG];
  Quat A_rot_inv[DEG];
  Quat A_accum_inv{0, 0, 0, 1};
  Mat3 A_post_inv[N];
  Mat3 Jr_dvec_inv[DEG];
  Mat3 Jr_kdelta[DEG];
  A_post_inv[N - 1] = so3_matrix(A_accum_inv);
  for (int i = DEG - 1; i >= 0; i--) {
    Quat R0 = load_q(knots[R_knot_offset + i]);
    Quat R1 = load_q(knots[R_knot_offset + i + 1]);
    d_vec[i] = so3_log(so3_mul(qconj(R0), R1));
    Vec3 k_delta = lambda_R[i + 1] * d_vec[i];
    A_rot_inv[i] = so3_exp(-k_delta);
    A_accum_inv = so3_mul(A_accum_inv, A_rot_inv[i]);
    if (J_rot_w || J_rot_a) {
      A_post_inv[i] = so3_matrix(A_accum_inv);
      Jr_dvec_inv[i] = right_jacobian_inv_so3(d_vec[i]);
      Jr_kdelta[i] = right_jacobian_so3(-k_delta);
    }
  }

  Vec3 omega[N];
  {
    omega[0] = v3(0, 0, 0);
    for (int i = 0; i < DEG; i++) omega[i + 1] = so3_rotate(A_rot_inv[i], omega[i]) + lambda_w[i + 1] * d_vec[i];
    d.gyro = omega[3];
    Quat Ri = load_q(knots[R_knot_offset]);
    Quat R_inv = so3_mul(A_accum_inv, qconj(Ri));
    d.accel = so3_rotate(R_inv, accelerate + gravity);
    d.R_inv = R_inv;
  }

  if (J_rot_w) {
    J_rot_w->start_idx = s;
    for (int i = 0; i < N; i++) J_rot_w->d_val_d_knot[i] = zero3();
    Mat3 d_omega_d_delta[DEG];
    d_omega_d_delta[0] = lambda_w[1] * A_post_inv[1];
    for (int i = 1; i < DEG; i++) {
      d_omega_d_delta[i] =
          lambda_R[i + 1] * A_post_inv[i] * hat(omega[i]) * Jr_kdelta[i] + lambda_w[i + 1] * A_post_inv[i + 1];
    }
    for (int i = 0; i < DEG; i++) {
      J_rot_w->d_val_d_knot[i] = J_rot_w->d_val_d_knot[i] - d_omega_d_delta[i] * transpose(Jr_dvec_inv[i]);
      J_rot_w->d_val_d_knot[i + 1] = J_rot_w->d_val_d_knot[i + 1] + d_omega_d_delta[i] * Jr_dvec_inv[i];
    }
  }

  if (J_rot_a) {
    Mat3 R_accum[DEG];
    Quat R0 = load_q(knots[R_knot_offset]);
    R_accum[0] = so3_matrix(R0);
    for (int i = 1; i < DEG; i++) R_accum[i] = R_accum[i - 1] * transpose(so3_matrix(A_rot_inv[i - 1]));
    J_rot_a->start_idx = s;
    for (int i = 0; i < N; i++) J_rot_a->d_val_d_knot[i] = zero3();
    Mat3 lhs = so3_matrix(d.R_inv) * hat(accelerate + gravity);
    J_rot_a->d_val_d_knot[0] = J_rot_a->d_val_d_knot[0] + lhs * R_accum[0];
    for (int i = 0; i < DEG; i++) {
      Mat3 d_a_d_delta = lambda_R[i + 1] * lhs * R_accum[i] * Jr_kdelta[i];
      J_rot_a->d_val_d_knot[i] = J_rot_a->d_val_d_knot[i] - d_a_d_delta * transpose(Jr_dvec_inv[i]);
      J_rot_a->d_val_d_knot[i + 1] = J_rot_a->d_val_d_knot[i + 1] + d_a_d_delta * Jr_dvec_inv[i];
    }
  }
  return d;
}

}  // namespace oracle
