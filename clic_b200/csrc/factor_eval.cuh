// Per-factor residual + analytic Jacobian rows, restructured for one-factor-per-thread evaluation on sm_100a.
//
// Reference arithmetic being replaced:
//   So3SplineView::EvaluateRp / EvaluateRTp / VelocityBody / accelerationBody   analytic_diff/so3_spline_view.h:127-439
//   RdSplineView::evaluate<D>                                                    analytic_diff/rd_spline_view.h:59-86
//   SplitSpineView::Evaluate                                                     analytic_diff/split_spline_view.h:56-188
//   LoamFeatureFactor / IMUFactor / ImageFeatureFactor ::Evaluate                lidar_feature_factor.h:65-159,
//                                                                                trajectory_value_factor.h:162-294,
//                                                                                image_feature_factor.h:67-268
// What is different from a transliteration:
//   * everything that depends on the knots only — delta_k = log(R_k^-1 R_k+1), |delta_k|, its unit axis and
//     Jr^-1(delta_k) — is computed once per knot pair (PairData, staged in shared memory) instead of per factor;
//   * exp(-lambda*delta) and lambda*Jr(+-lambda*delta) are written on the unit axis of delta_k, so a factor needs
//     three sincos and no sqrt / division;
//   * the LiDAR Jacobian chain is evaluated vector-first (a 1x3 row pushed through the 3x3 chain), never
//     forming the 3x3 d_val_d_knot matrices.
// Output of every evaluator is a set of "rows" of the local Jacobian J~ (loss-corrected) in the local column
// order [knot s: dtheta(3) dp(3) | knot s+1 | knot s+2 | knot s+3 | extras... | r~], written through a Sink so the
// same code feeds shared memory (GPU) or a plain array (host emulation test).
#pragma once
#include "clic_math.cuh"

namespace clic {

// Per knot pair k -> k+1.  17 doubles.
struct PairData {
  double delta[3];
  double theta;
  double inv_theta;  // 0 when theta == 0 (identical knots): every use degrades to the exact theta = 0 limit
  double nhat[3];
  double Jrinv[9];   // rightJacobianInvSO3(delta), row-major
};
constexpr int kPairDoubles = 17;

CLIC_HD void compute_pair(const double* q0, const double* q1, PairData* P) {
  Q4 a; a.x = q0[0]; a.y = q0[1]; a.z = q0[2]; a.w = q0[3];
  Q4 b; b.x = q1[0]; b.y = q1[1]; b.z = q1[2]; b.w = q1[3];
  V3 d = so3_log(so3_mul(qconj(a), b));  // (R0.inverse() * R1).log(), so3_spline_view.h:151
  double th = sqrt(dot(d, d));
  double inv = th > 0.0 ? 1.0 / th : 0.0;
  P->delta[0] = d.x; P->delta[1] = d.y; P->delta[2] = d.z;
  P->theta = th;
  P->inv_theta = inv;
  P->nhat[0] = d.x * inv; P->nhat[1] = d.y * inv; P->nhat[2] = d.z * inv;
  M3 J = right_jacobian_inv(d);
#pragma unroll
  for (int i = 0; i < 9; i++) P->Jrinv[i] = J.m[i];
}

// Read-only view of the window state staged on chip.
struct WindowView {
  const double* kq;      // K*4
  const double* kp;      // K*3
  const PairData* pair;  // K-1
};

CLIC_HD Q4 ldq(const double* p) { Q4 q; q.x = p[0]; q.y = p[1]; q.z = p[2]; q.w = p[3]; return q; }
CLIC_HD V3 ldv(const double* p) { return mk(p[0], p[1], p[2]); }

// One cumulative-spline step on the unit axis: returns exp(-lambda*delta_k) and the two scalars of
// lambda*Jr(+-lambda*delta_k) = lambda I -+ ca hat(n) + cb hat(n)^2   (ca = (1-cos th)/theta_k, cb = lambda - sin th/theta_k)
CLIC_HD Q4 step_exp_neg(const PairData& P, double lam, double* ca, double* cb) {
  double th = lam * P.theta;
  double s, c;
  sincos_half_pi(0.5 * th, &s, &c);  // |th / 2| <= pi / 2: lam in [0, 1], theta <= pi
  *ca = (2.0 * s * s) * P.inv_theta;
  *cb = lam - (2.0 * s * c) * P.inv_theta;
  Q4 e; e.x = -s * P.nhat[0]; e.y = -s * P.nhat[1]; e.z = -s * P.nhat[2]; e.w = c;
  return e;
}
// row-vector w times lambda*Jr(sign*lambda*delta):  w*hat(n) = w x n
CLIC_HD V3 row_times_lamJr(V3 w, const PairData& P, double lam, double ca, double cb, double sign) {
  V3 n = ldv(P.nhat);
  V3 x1 = cross(w, n);
  V3 x2 = cross(x1, n);
  return fma3(cb, x2, fma3(-sign * ca, x1, lam * w));
}
// matrix lambda*Jr(sign*lambda*delta)
CLIC_HD M3 lamJr(const PairData& P, double lam, double ca, double cb, double sign) {
  M3 K = hat(ldv(P.nhat));
  M3 K2 = mmul(K, K);
  M3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = -sign * ca * K.m[i] + cb * K2.m[i];
  r.m[0] += lam; r.m[4] += lam; r.m[8] += lam;
  return r;
}

struct LoamShared {  // arguments of AddLoamMeasurementAnalytic shared by a batch (trajectory_estimator.cpp:712-716)
  Q4 S_GtoM; V3 p_GinM; Q4 S_LtoI; V3 p_LinI; double weight;
  bool gm_identity;  // S_GtoM = I, p_GinM = 0: every production call (trajectory_manager.cpp:676-680) -> skip two rotations
  bool point_in_imu; // the stream already holds p_I = S_LtoI p_L + p_LinI (rotated once at add time by the prep kernel)
};

// LoamFeatureFactor::Evaluate.  geo = plane (n.x n.y n.z d) or line (point xyz, direction xyz).
// J[24] gets the local row [knot: dtheta(3) dp(3)] x 4; *r the weighted residual.  No loss (":743").
// `jscale` multiplies the emitted Jacobian row (sh.weight for a contributing factor, 0 to mask a lane out).
// Out: out.put(col, value) for the 24 local columns (an array writer on the host, a shared-memory writer in the kernel).
struct ArrayOut24 {
  double* J;
  CLIC_HD void put(int c, double v) { J[c] = v; }
};
template <class Out>
CLIC_HD void loam_eval_to(const WindowView& W, int s, double u, V3 pL, bool is_plane, const double geo[6],
                          const LoamShared& sh, bool want_jac, double* r_out, Out& out, double jscale) {
  double lam[4], c[4];
  blend_cum(u, lam);
  blend_plain(u, c);
  Q4 acc; acc.x = acc.y = acc.z = 0.0; acc.w = 1.0;
  Q4 A[3];
  double ca[3], cb[3];
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    Q4 e = step_exp_neg(W.pair[s + i], lam[i + 1], &ca[i], &cb[i]);
    acc = qmul(acc, e);  // A_accum_inv *= exp(-kdelta)
    A[i] = acc;
  }
  Q4 R = qmul(ldq(W.kq + 4 * s), qconj(acc));  // Ri * A_accum_inv.inverse()
  V3 p_I = sh.point_in_imu ? pL : qrot(sh.S_LtoI, pL) + sh.p_LinI;
  V3 pos = mk(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 4; i++) pos = fma3(c[i], ldv(W.kp + 3 * (s + i)), pos);
  V3 p_G = qrot(R, p_I) + pos;
  V3 p_M = sh.gm_identity ? p_G : qrot(sh.S_GtoM, p_G) + sh.p_GinM;
  V3 Jpi;
  double r;
  if (is_plane) {
    Jpi = mk(geo[0], geo[1], geo[2]);
    r = dot(p_M, Jpi) + geo[3];
  } else {
    V3 gl = mk(geo[3], geo[4], geo[5]);
    V3 dv = cross(p_M - mk(geo[0], geo[1], geo[2]), gl);
    r = sqrt(dot(dv, dv));
    Jpi = cross((-1.0 / r) * dv, gl);  // -dist_vec^T / r * hat(geo_normal), lidar_feature_factor.h:101
  }
  *r_out = r * sh.weight;
  if (!want_jac) return;
  V3 g = sh.gm_identity ? Jpi : qrot_inv(sh.S_GtoM, Jpi);  // (J_pi^T S_GtoM)^T
  V3 m = qrot_inv(R, g);            // (J_pi^T S_GtoM R)^T
  V3 v = cross(p_I, m);             // J_pi^T (-S_GtoM R hat(p_I)) = p_I x m
  // rows of d_val_d_knot chain, vector-first (so3_spline_view.h:170-180)
  V3 w0 = qrot_inv(A[0], v), w1 = qrot_inv(A[1], v), w2 = qrot_inv(A[2], v);
  V3 t0 = row_times_lamJr(w1, W.pair[s + 0], lam[1], ca[0], cb[0], +1.0);
  V3 t1 = row_times_lamJr(w2, W.pair[s + 1], lam[2], ca[1], cb[1], +1.0);
  V3 t2 = row_times_lamJr(v, W.pair[s + 2], lam[3], ca[2], cb[2], +1.0);
  V3 j0 = w0 - vmat_t(t0, W.pair[s + 0].Jrinv);
  V3 j1 = vmat(t0, W.pair[s + 0].Jrinv) - vmat_t(t1, W.pair[s + 1].Jrinv);
  V3 j2 = vmat(t1, W.pair[s + 1].Jrinv) - vmat_t(t2, W.pair[s + 2].Jrinv);
  V3 j3 = vmat(t2, W.pair[s + 2].Jrinv);
  const double w = jscale;
  out.put(0, w * j0.x); out.put(1, w * j0.y); out.put(2, w * j0.z);
  out.put(6, w * j1.x); out.put(7, w * j1.y); out.put(8, w * j1.z);
  out.put(12, w * j2.x); out.put(13, w * j2.y); out.put(14, w * j2.z);
  out.put(18, w * j3.x); out.put(19, w * j3.y); out.put(20, w * j3.z);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double wc = w * c[k];
    out.put(6 * k + 3, wc * g.x); out.put(6 * k + 4, wc * g.y); out.put(6 * k + 5, wc * g.z);
  }
}
CLIC_HD void loam_eval(const WindowView& W, int s, double u, V3 pL, bool is_plane, const double geo[6],
                       const LoamShared& sh, bool want_jac, double* r_out, double J[24], double jscale) {
  ArrayOut24 o;
  o.J = J;
  loam_eval_to(W, s, u, pL, is_plane, geo, sh, want_jac, r_out, o, jscale);
}

// ceres::CauchyLoss(a) + Corrector (marginalization_factor.cpp:38-61): rho'' < 0 always, so the correction is the
// pure sqrt(rho') scaling of r and J.  Returns the scale; *rho = rho(s).
CLIC_HD double cauchy_scale(double sq_norm, double a, double* rho) {
  double b = a * a, cinv = 1.0 / b;
  double sum = 1.0 + sq_norm * cinv;
  double inv = 1.0 / sum;
  *rho = b * log(sum);
  return sqrt(inv);
}

struct ImuShared {  // per AddIMUMeasurementAnalytic batch
  double info[6];
  V3 bg, ba, gravity;
  double loss_a;     // Cauchy parameter (10, or 3 when marginalised); <= 0: no loss
  double inv_dt;     // 1e9 / dt_ns
  bool want_toff = true;  // false: the IMU time offset is a constant of this solve — its Jacobian column (body
                          // acceleration chain + jerk, ~200 FP64 instructions per factor) is never assembled: emit zeros
};

// IMUFactor::Evaluate on SplitSpineView::Evaluate.  Emits 6 rows of NC = 32 (solve) or 40 (marg: + gravity) columns:
//   [knot cols 0..23 | bg 24..26 | ba 27..29 | (solve) toff 30, r 31 | (marg) gravity 30..32, toff 33, r 34, pad]
// through sink.put(row, col, value) (all columns of a row are written, zeros included), then sink.row_done(row).
// Returns rho(|r|^2) (or |r|^2 without loss).
template <bool MARG, class Sink>
CLIC_HD double imu_eval(const WindowView& W, int s, double u, V3 gyro_m, V3 accel_m, const ImuShared& sh,
                        bool want_jac, Sink& sink) {
  constexpr int C_TOFF = MARG ? 33 : 30, C_R = MARG ? 34 : 31, NC = MARG ? 40 : 32;
  double lamR[4], lamW[4], lamA[4];
  blend_cum(u, lamR);
  blend_cum_d1(u, sh.inv_dt, lamW);
  blend_plain_d2(u, sh.inv_dt * sh.inv_dt, lamA);
  V3 acc_w = mk(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 4; i++) acc_w = fma3(lamA[i], ldv(W.kp + 3 * (s + i)), acc_w);
  Q4 acc; acc.x = acc.y = acc.z = 0.0; acc.w = 1.0;
  Q4 Ainv[3];
  Q4 Aq[3];
  double ca[3], cb[3];
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    Ainv[i] = step_exp_neg(W.pair[s + i], lamR[i + 1], &ca[i], &cb[i]);
    acc = qmul(acc, Ainv[i]);
    Aq[i] = acc;
  }
  V3 omega[4];
  omega[0] = mk(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++) omega[i + 1] = fma3(lamW[i + 1], ldv(W.pair[s + i].delta), qrot(Ainv[i], omega[i]));
  Q4 q_s = ldq(W.kq + 4 * s);
  Q4 R_inv = qmul(acc, qconj(q_s));
  V3 ag = acc_w + sh.gravity;
  V3 accel_b = qrot(R_inv, ag);
  double res[6];
  {
    V3 rg = omega[3] - (gyro_m - sh.bg);
    V3 ra = accel_b - (accel_m - sh.ba);
    res[0] = sh.info[0] * rg.x; res[1] = sh.info[1] * rg.y; res[2] = sh.info[2] * rg.z;
    res[3] = sh.info[3] * ra.x; res[4] = sh.info[4] * ra.y; res[5] = sh.info[5] * ra.z;
  }
  double sq = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) sq += res[i] * res[i];
  double rho = sq, sc = 1.0;
  if (sh.loss_a > 0) sc = cauchy_scale(sq, sh.loss_a, &rho);
  if (!want_jac) return rho;

  // Row-wise (vector-first) evaluation: a residual row only needs ROWS of the rotation matrices involved, pushed as
  // 1x3 vectors through lambda*Jr(-k_delta) and Jr^-1 — no 3x3 matrix products, short register live ranges.
  //   A_post_inv[i] = mat(Aq[i]) (A_post_inv[3] = I),  R_accum[i] = mat(P_i), P_0 = R_s, P_i+1 = P_i * exp(+k_delta_i)
  const Q4 P1 = qmul(q_s, qconj(Ainv[0]));
  const Q4 P2 = qmul(P1, qconj(Ainv[1]));
  M3 Rinv_m = qmat(R_inv);
  // time-offset column (trajectory_value_factor.h:268-291)
  V3 jt_top = mk(0, 0, 0), jt_bot = mk(0, 0, 0);
  if (sh.want_toff) {
    double lamWW[4];
    blend_cum_d2(u, sh.inv_dt * sh.inv_dt, lamWW);
    V3 rv = mk(0, 0, 0), ra = mk(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 3; i++) {  // So3SplineView::accelerationBody, so3_spline_view.h:396-439
      V3 d = ldv(W.pair[s + i].delta);
      rv = qrot(Ainv[i], rv);
      V3 vc = lamW[i + 1] * d;
      rv = rv + vc;
      ra = qrot(Ainv[i], ra);
      ra = ra + (lamWW[i + 1] * d + cross(rv, vc));
    }
    double cj[4];
    blend_plain_d3(sh.inv_dt * sh.inv_dt * sh.inv_dt, cj);
    V3 jerk = mk(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 4; i++) jerk = fma3(cj[i], ldv(W.kp + 3 * (s + i)), jerk);
    // With rot = R(t), rot_dot = R hat(omega):  vee(rot_dot^T rot_dot) + vee(rot^T (rot_dot hat(omega) + rot hat(ra)))
    // = vee(-hat(omega)^2) + vee(hat(omega)^2 + hat(ra)) = ra   (the symmetric parts cancel; the reference forms the
    // 3x3 products, equal to 1e-16 relative), and rot_dot^T rot a_b + rot^T jerk = a_b x omega + R^-1 jerk.
    jt_top = ra;
    jt_bot = cross(accel_b, omega[3]) + qrot(R_inv, jerk);
  }
  // knot rotation blocks, one residual row at a time:
  //   gyro  rows (split_spline_view.h:137-157): dW_i = lamR_i+1 A_post_inv[i] hat(omega_i) Jr(-kd_i) + lamW_i+1 A_post_inv[i+1]
  //   accel rows (":160-186"):                  dA_i = lamR_i+1 lhs R_accum[i] Jr(-kd_i),  lhs = R_inv hat(a_w + g)
  //   J[0] = X - d_0 Jri_0^T ; J[k] = d_k-1 Jri_k-1 - d_k Jri_k^T ; J[3] = d_2 Jri_2      (X = 0 | lhs R_s)
  // The two halves are distinct code (unrolled); the three rows of a half run as a LOOP (rr is a run-time value, every
  // per-row quantity is picked with selects): a third of the instructions of the fully unrolled form — the evaluator is
  // the bulk of the IMU stream's code, and a pass that starts with a cold instruction cache pays for every line.
#pragma unroll
  for (int half = 0; half < 2; half++) {
#pragma unroll 1
    for (int rr = 0; rr < 3; rr++) {
      const int row = half * 3 + rr;
      const double info_r = rr == 0 ? sh.info[half * 3] : (rr == 1 ? sh.info[half * 3 + 1] : sh.info[half * 3 + 2]);
      const double res_r = rr == 0 ? res[half * 3] : (rr == 1 ? res[half * 3 + 1] : res[half * 3 + 2]);
      const double sI = sc * info_r;
      const V3 Rrow = mk(rr == 0 ? Rinv_m.m[0] : (rr == 1 ? Rinv_m.m[3] : Rinv_m.m[6]),
                         rr == 0 ? Rinv_m.m[1] : (rr == 1 ? Rinv_m.m[4] : Rinv_m.m[7]),
                         rr == 0 ? Rinv_m.m[2] : (rr == 1 ? Rinv_m.m[5] : Rinv_m.m[8]));  // row rr of R_inv
      V3 d0, d1, d2, x;
      if (half == 0) {
        const V3 e = mk(rr == 0 ? 1.0 : 0.0, rr == 1 ? 1.0 : 0.0, rr == 2 ? 1.0 : 0.0);
        const V3 a1 = qrot_inv(Aq[1], e), a2 = qrot_inv(Aq[2], e);  // rows rr of A_post_inv[1], [2]; [3] = I -> e
        x = mk(0, 0, 0);
        d0 = lamW[1] * a1;
        d1 = fma3(lamW[2], a2, row_times_lamJr(cross(a1, omega[1]), W.pair[s + 1], lamR[2], ca[1], cb[1], -1.0));
        d2 = fma3(lamW[3], e, row_times_lamJr(cross(a2, omega[2]), W.pair[s + 2], lamR[3], ca[2], cb[2], -1.0));
      } else {
        const V3 v = cross(Rrow, ag);  // row rr of lhs
        x = qrot_inv(q_s, v);          // row rr of lhs * R_s
        d0 = row_times_lamJr(x, W.pair[s + 0], lamR[1], ca[0], cb[0], -1.0);
        d1 = row_times_lamJr(qrot_inv(P1, v), W.pair[s + 1], lamR[2], ca[1], cb[1], -1.0);
        d2 = row_times_lamJr(qrot_inv(P2, v), W.pair[s + 2], lamR[3], ca[2], cb[2], -1.0);
      }
      V3 j[4];
      j[0] = x - vmat_t(d0, W.pair[s + 0].Jrinv);
      j[1] = vmat(d0, W.pair[s + 0].Jrinv) - vmat_t(d1, W.pair[s + 1].Jrinv);
      j[2] = vmat(d1, W.pair[s + 1].Jrinv) - vmat_t(d2, W.pair[s + 2].Jrinv);
      j[3] = vmat(d2, W.pair[s + 2].Jrinv);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        sink.put(row, 6 * k + 0, sI * j[k].x);
        sink.put(row, 6 * k + 1, sI * j[k].y);
        sink.put(row, 6 * k + 2, sI * j[k].z);
        // position blocks: accel rows only, lambda_a[k] * R_inv (trajectory_value_factor.h:230-242)
        double pc = half == 0 ? 0.0 : sI * lamA[k];
        sink.put(row, 6 * k + 3, pc * Rrow.x);
        sink.put(row, 6 * k + 4, pc * Rrow.y);
        sink.put(row, 6 * k + 5, pc * Rrow.z);
      }
      // bias blocks (":245-257")
#pragma unroll
      for (int cc = 0; cc < 3; cc++) {
        sink.put(row, 24 + cc, (half == 0 && cc == rr) ? sI : 0.0);
        sink.put(row, 27 + cc, (half == 1 && cc == rr) ? sI : 0.0);
      }
      if (MARG) {  // gravity block (":260-266"), ambient 3 columns
        sink.put(row, 30, half == 1 ? sI * Rrow.x : 0.0);
        sink.put(row, 31, half == 1 ? sI * Rrow.y : 0.0);
        sink.put(row, 32, half == 1 ? sI * Rrow.z : 0.0);
      }
      double jt = half == 0 ? (rr == 0 ? jt_top.x : rr == 1 ? jt_top.y : jt_top.z)
                            : (rr == 0 ? jt_bot.x : rr == 1 ? jt_bot.y : jt_bot.z);
      sink.put(row, C_TOFF, (1e-9 * sI) * jt);
      sink.put(row, C_R, sc * res_r);
#pragma unroll
      for (int cc = C_R + 1; cc < NC; cc++) sink.put(row, cc, 0.0);
      sink.row_done(half * 3);  // only gyro (< 3) / accel matters to the sinks that act on it
    }
  }
  return rho;
}

// Vector-first Jacobian rows of a cumulative SO(3) spline value through a 1x3 row `a`.
//   RP  (EvaluateRp,  so3_spline_view.h:127-183):  rows of  a * d_val_d_knot[k],  d from A_post_inv (inverse accumulation)
//   RTP (EvaluateRTp, so3_spline_view.h:193-256):  rows of  a * d_val_d_knot[k],  d from Ri_A_pre (forward accumulation)
struct So3Chain {
  Q4 R;          // R(t)
  Q4 acc[3];     // RP: A_accum_inv after steps i = 2,1,0 stored at [i];  RTP: prefix products P_1..P_3 at [0..2]
  double lam[4], ca[3], cb[3];
};
CLIC_HD void so3_chain_rp(const WindowView& W, int s, double u, So3Chain* C) {
  blend_cum(u, C->lam);
  Q4 acc; acc.x = acc.y = acc.z = 0.0; acc.w = 1.0;
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    Q4 e = step_exp_neg(W.pair[s + i], C->lam[i + 1], &C->ca[i], &C->cb[i]);
    acc = qmul(acc, e);
    C->acc[i] = acc;
  }
  C->R = qmul(ldq(W.kq + 4 * s), qconj(acc));
}
CLIC_HD void so3_rows_rp(const WindowView& W, int s, const So3Chain& C, V3 a, V3 j[4]) {
  V3 w0 = qrot_inv(C.acc[0], a), w1 = qrot_inv(C.acc[1], a), w2 = qrot_inv(C.acc[2], a);
  V3 t0 = row_times_lamJr(w1, W.pair[s + 0], C.lam[1], C.ca[0], C.cb[0], +1.0);
  V3 t1 = row_times_lamJr(w2, W.pair[s + 1], C.lam[2], C.ca[1], C.cb[1], +1.0);
  V3 t2 = row_times_lamJr(a, W.pair[s + 2], C.lam[3], C.ca[2], C.cb[2], +1.0);
  j[0] = w0 - vmat_t(t0, W.pair[s + 0].Jrinv);
  j[1] = vmat(t0, W.pair[s + 0].Jrinv) - vmat_t(t1, W.pair[s + 1].Jrinv);
  j[2] = vmat(t1, W.pair[s + 1].Jrinv) - vmat_t(t2, W.pair[s + 2].Jrinv);
  j[3] = vmat(t2, W.pair[s + 2].Jrinv);
}
// forward accumulation: P_0 = R_s, P_{i+1} = P_i * exp(+lambda_{i+1} delta_i);  R(t) = P_3
CLIC_HD void so3_chain_rtp(const WindowView& W, int s, double u, So3Chain* C, Q4 P[4]) {
  blend_cum(u, C->lam);
  P[0] = ldq(W.kq + 4 * s);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Q4 e = qconj(step_exp_neg(W.pair[s + i], C->lam[i + 1], &C->ca[i], &C->cb[i]));
    P[i + 1] = qmul(P[i], e);
  }
  C->R = P[3];
}
CLIC_HD void so3_rows_rtp(const WindowView& W, int s, const So3Chain& C, const Q4 P[4], V3 a, V3 j[4]) {
  V3 w0 = qrot_inv(P[0], a), w1 = qrot_inv(P[1], a), w2 = qrot_inv(P[2], a);
  V3 t0 = row_times_lamJr(w0, W.pair[s + 0], C.lam[1], C.ca[0], C.cb[0], -1.0);
  V3 t1 = row_times_lamJr(w1, W.pair[s + 1], C.lam[2], C.ca[1], C.cb[1], -1.0);
  V3 t2 = row_times_lamJr(w2, W.pair[s + 2], C.lam[3], C.ca[2], C.cb[2], -1.0);
  j[0] = w0 - vmat_t(t0, W.pair[s + 0].Jrinv);
  j[1] = vmat(t0, W.pair[s + 0].Jrinv) - vmat_t(t1, W.pair[s + 1].Jrinv);
  j[2] = vmat(t1, W.pair[s + 1].Jrinv) - vmat_t(t2, W.pair[s + 2].Jrinv);
  j[3] = vmat(t2, W.pair[s + 2].Jrinv);
}
// So3SplineView::VelocityBody without Jacobian (so3_spline_view.h:330-394): omega_body(t)
CLIC_HD V3 so3_velocity_body(const WindowView& W, int s, double u, double inv_dt) {
  double lam[4], dl[4];
  blend_cum(u, lam);
  blend_cum_d1(u, inv_dt, dl);
  V3 w = mk(0, 0, 0);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double ca, cb;
    Q4 e = step_exp_neg(W.pair[s + i], lam[i + 1], &ca, &cb);
    w = fma3(dl[i + 1], ldv(W.pair[s + i].delta), qrot(e, w));
  }
  return w;
}

// ---- rows of the two remaining rotation views (used by the factors of SURVEY.md §8f-4) ----
// EvaluateRotation (so3_spline_view.h:65-117): forward products P_0 = R_s, P_i+1 = P_i exp(lambda_i+1 delta_i);
// J_helper_i = lambda_i+1 mat(P_i+1) Jr(+lambda delta_i).  Rows a * d_val_d_knot[k].
CLIC_HD void so3_rows_rot(const WindowView& W, int s, const So3Chain& C, const Q4 P[4], V3 a, V3 j[4]) {
  V3 w0 = qrot_inv(P[0], a);
  V3 t0 = row_times_lamJr(qrot_inv(P[1], a), W.pair[s + 0], C.lam[1], C.ca[0], C.cb[0], +1.0);
  V3 t1 = row_times_lamJr(qrot_inv(P[2], a), W.pair[s + 1], C.lam[2], C.ca[1], C.cb[1], +1.0);
  V3 t2 = row_times_lamJr(qrot_inv(P[3], a), W.pair[s + 2], C.lam[3], C.ca[2], C.cb[2], +1.0);
  j[0] = w0 - vmat_t(t0, W.pair[s + 0].Jrinv);
  j[1] = vmat(t0, W.pair[s + 0].Jrinv) - vmat_t(t1, W.pair[s + 1].Jrinv);
  j[2] = vmat(t1, W.pair[s + 1].Jrinv) - vmat_t(t2, W.pair[s + 2].Jrinv);
  j[3] = vmat(t2, W.pair[s + 2].Jrinv);
}
// EvaluateLogRT (so3_spline_view.h:263-321): inverse accumulation as in EvaluateRp (C from so3_chain_rp), but
// J_helper_i = lambda_i+1 A_post_inv[i] Jr(-lambda delta_i) and d_val_d_knot[0] starts from A_post_inv[0].
// The value R(t)^T = A_accum_inv R_s^-1 = conj(C.R).
CLIC_HD void so3_rows_logrt(const WindowView& W, int s, const So3Chain& C, V3 a, V3 j[4]) {
  V3 w0 = qrot_inv(C.acc[0], a), w1 = qrot_inv(C.acc[1], a), w2 = qrot_inv(C.acc[2], a);
  V3 t0 = row_times_lamJr(w0, W.pair[s + 0], C.lam[1], C.ca[0], C.cb[0], -1.0);
  V3 t1 = row_times_lamJr(w1, W.pair[s + 1], C.lam[2], C.ca[1], C.cb[1], -1.0);
  V3 t2 = row_times_lamJr(w2, W.pair[s + 2], C.lam[3], C.ca[2], C.cb[2], -1.0);
  j[0] = w0 - vmat_t(t0, W.pair[s + 0].Jrinv);
  j[1] = vmat(t0, W.pair[s + 0].Jrinv) - vmat_t(t1, W.pair[s + 1].Jrinv);
  j[2] = vmat(t1, W.pair[s + 1].Jrinv) - vmat_t(t2, W.pair[s + 2].Jrinv);
  j[3] = vmat(t2, W.pair[s + 2].Jrinv);
}

// ---- the remaining adders of TrajectoryEstimator: a handful of factors per solve at most, evaluated
// by ONE thread each into a small dense row set over the global columns the factor touches (small_factors_kernel).
// The kinds are the CLIC_FACTOR_* values of include/clic_cuda.h (clic_evaluate_factor takes them as they are).
enum {
  kSmallPose = 0, kSmallLocalVelocity = 1, kSmallRelativeRotation = 2, kSmallImageDepth = 3, kSmallPreIntegration = 4,
  kSmallImage3D2D = 5, kSmallImageOnePose = 6, kSmallEpipolar = 7, kSmallLoamOptMapPose = 8, kSmallRelativeLoam = 9
};
struct SmallFactor {
  int kind;
  int landmark;        // image depth
  long long t_a, t_b;  // measurement time(s) in ns (without offset)
  double d[24];        // the factor's constants, laid out as clic_factor_desc::d documents per kind
  double loss_a;       // Cauchy parameter, <= 0: none
  int geo_type;        // the two LiDAR kinds
  int ext;             // pre-integration: index into the PreintData array
  int slot_i, slot_j;  // pre-integration: bias slots of t_a, t_b
};
// IntegrationBase as the factor reads it (clic_preintegration), the five 3x3 blocks of its Jacobian pulled out
struct PreintData {
  double sum_dt, gravity_norm;
  double delta_p[3], delta_q[4], delta_v[3], linearized_ba[3], linearized_bg[3];
  double dp_dba[9], dp_dbg[9], dq_dbg[9], dv_dba[9], dv_dbg[9];
  double sqrt_info[225];
};
constexpr int kSmallMaxRows = 15, kSmallMaxCols = 64;
struct SmallOut {
  int nres, ncols;
  double r[kSmallMaxRows];
  int gcol[kSmallMaxCols];                  // distinct global columns (constant blocks are never entered)
  double J[kSmallMaxRows][kSmallMaxCols];
  CLIC_HD void reset(int n) {
    nres = n;
    ncols = 0;
    for (int i = 0; i < kSmallMaxRows; i++) r[i] = 0.0;
  }
  CLIC_HD void add(int row, int g, double v) {  // J[row][column of g] += v; a knot shared by two segments accumulates
    if (g < 0) return;
    int c = 0;
    while (c < ncols && gcol[c] != g) c++;
    if (c == ncols) {
      if (ncols == kSmallMaxCols) return;
      gcol[ncols++] = g;
      for (int i = 0; i < kSmallMaxRows; i++) J[i][c] = 0.0;
    }
    J[row][c] += v;
  }
};
struct SmallCols {      // where a factor's parameter blocks sit in H (device copies of the layout)
  const int* knot_col;  // [K] column of the knot's dtheta (dp = +3), -1 constant
  int col_toff_imu;     // -1 constant
  const int* lm_col;    // per landmark, -1 constant (nullptr: none free)
  int col_bias_gyro, col_bias_accel;  // + 3 * slot, -1 constant
};

// IMUPoseFactor::Evaluate (trajectory_value_factor.h:323-427).  False: the corrected time left the window.
CLIC_HD bool small_pose_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K, double toff,
                             const SmallCols& sc, SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a + (long long)toff, t0_ns, dt_ns, K, &s, &u)) return false;
  const double inv_dt = 1e9 / double(dt_ns);
  const V3 p_meas = ldv(f.d);
  const Q4 q_meas = ldq(f.d + 3);
  const double* info = f.d + 7;
  So3Chain C;
  Q4 P[4];
  so3_chain_rtp(W, s, u, &C, P);  // forward accumulation: R(t) = P[3]
  double c[4];
  blend_plain(u, c);
  V3 pos = mk(0, 0, 0);
  for (int k = 0; k < 4; k++) pos = fma3(c[k], ldv(W.kp + 3 * (s + k)), pos);
  const V3 res_R = so3_log(so3_mul(C.R, qconj(q_meas)));
  const V3 res_p = pos - p_meas;
  o->reset(6);
  const M3 Jrot = left_jacobian_inv(res_R);
  for (int r = 0; r < 3; r++) {
    V3 j[4];
    so3_rows_rot(W, s, C, P, mk(Jrot.m[3 * r], Jrot.m[3 * r + 1], Jrot.m[3 * r + 2]), j);
    for (int k = 0; k < 4; k++) {
      const int g = sc.knot_col[s + k];
      if (g < 0) continue;
      o->add(r, g + 0, info[r] * j[k].x);
      o->add(r, g + 1, info[r] * j[k].y);
      o->add(r, g + 2, info[r] * j[k].z);
      o->add(3 + r, g + 3 + r, info[3 + r] * c[k]);
    }
  }
  if (sc.col_toff_imu >= 0) {  // ":410-426": as written — a quaternion made from the SKEW matrix of the body rate
    const V3 gyro = so3_velocity_body(W, s, u, inv_dt);
    double d1[4];
    blend_plain_d1(u, inv_dt, d1);
    V3 vel = mk(0, 0, 0);
    for (int k = 0; k < 4; k++) vel = fma3(d1[k], ldv(W.kp + 3 * (s + k)), vel);
    // Eigen::Quaterniond(hat(gyro)): trace 0, zero diagonal -> i = 0: (x, y, z, w) = (1/2, 0, 0, gyro.x), normalised
    const double nq = sqrt(0.25 + gyro.x * gyro.x);
    Q4 qg; qg.x = 0.5 / nq; qg.y = 0.0; qg.z = 0.0; qg.w = gyro.x / nq;
    const V3 jr = so3_log(so3_mul(so3_mul(C.R, qg), qconj(q_meas)));
    o->add(0, sc.col_toff_imu, 1e-9 * info[0] * jr.x);
    o->add(1, sc.col_toff_imu, 1e-9 * info[1] * jr.y);
    o->add(2, sc.col_toff_imu, 1e-9 * info[2] * jr.z);
    o->add(3, sc.col_toff_imu, 1e-9 * info[3] * vel.x);
    o->add(4, sc.col_toff_imu, 1e-9 * info[4] * vel.y);
    o->add(5, sc.col_toff_imu, 1e-9 * info[5] * vel.z);
  }
  o->r[0] = info[0] * res_R.x; o->r[1] = info[1] * res_R.y; o->r[2] = info[2] * res_R.z;
  o->r[3] = info[3] * res_p.x; o->r[4] = info[4] * res_p.y; o->r[5] = info[5] * res_p.z;
  return true;
}

// LocalVelocityFactor::Evaluate (trajectory_value_factor.h:464-566)
CLIC_HD bool small_local_velocity_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                       double toff, const SmallCols& sc, SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a + (long long)toff, t0_ns, dt_ns, K, &s, &u)) return false;
  const double inv_dt = 1e9 / double(dt_ns);
  const V3 v = ldv(f.d);
  const double w = f.d[3];
  So3Chain C;
  so3_chain_rp(W, s, u, &C);
  double d1[4];
  blend_plain_d1(u, inv_dt, d1);
  V3 vel = mk(0, 0, 0);
  for (int k = 0; k < 4; k++) vel = fma3(d1[k], ldv(W.kp + 3 * (s + k)), vel);
  const V3 res = qrot(C.R, v) - vel;
  o->reset(3);
  for (int r = 0; r < 3; r++) {
    // row r of jac_lhs_R = -R hat(v):  (row r of R) x v negated = v x R_r
    const V3 Rr = qrot_inv(C.R, mk(r == 0 ? 1.0 : 0.0, r == 1 ? 1.0 : 0.0, r == 2 ? 1.0 : 0.0));
    V3 j[4];
    so3_rows_rp(W, s, C, cross(v, Rr), j);
    for (int k = 0; k < 4; k++) {
      const int g = sc.knot_col[s + k];
      if (g < 0) continue;
      o->add(r, g + 0, w * j[k].x);
      o->add(r, g + 1, w * j[k].y);
      o->add(r, g + 2, w * j[k].z);
      o->add(r, g + 3 + r, -w * d1[k]);
    }
  }
  if (sc.col_toff_imu >= 0) {  // ":548-563": rot_dot v - accel, rot_dot = R hat(omega)
    const V3 gyro = so3_velocity_body(W, s, u, inv_dt);
    double d2[4];
    blend_plain_d2(u, inv_dt * inv_dt, d2);
    V3 acc = mk(0, 0, 0);
    for (int k = 0; k < 4; k++) acc = fma3(d2[k], ldv(W.kp + 3 * (s + k)), acc);
    const V3 jt = qrot(C.R, cross(gyro, v)) - acc;
    o->add(0, sc.col_toff_imu, 1e-9 * w * jt.x);
    o->add(1, sc.col_toff_imu, 1e-9 * w * jt.y);
    o->add(2, sc.col_toff_imu, 1e-9 * w * jt.z);
  }
  o->r[0] = w * res.x; o->r[1] = w * res.y; o->r[2] = w * res.z;
  return true;
}

// RelativeOrientationFactor::Evaluate (trajectory_value_factor.h:990-1064).  R(t_b)^T is taken as the quaternion product
// itself; the reference re-normalises it through a rotation matrix (so3_spline_view.h:303-306): 1e-16 apart.
CLIC_HD bool small_relative_rotation_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                          const SmallCols& sc, SmallOut* o) {
  int sa = 0, sb = 0;
  double ua = 0.0, ub = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &sa, &ua) || !index_ns(f.t_b, t0_ns, dt_ns, K, &sb, &ub)) return false;
  const Q4 S_BtoA = ldq(f.d);
  const double* info = f.d + 4;
  So3Chain Ca, Cb;
  Q4 Pa[4];
  so3_chain_rtp(W, sa, ua, &Ca, Pa);  // R(t_a), forward
  so3_chain_rp(W, sb, ub, &Cb);       // inverse accumulation at t_b
  const Q4 RbT = qconj(Cb.R);
  const V3 res = so3_log(so3_mul(so3_mul(RbT, Ca.R), S_BtoA));
  o->reset(3);
  const M3 Jrot = left_jacobian_inv(res);
  for (int r = 0; r < 3; r++) {
    const V3 Jr = mk(Jrot.m[3 * r], Jrot.m[3 * r + 1], Jrot.m[3 * r + 2]);
    V3 ja[4], jb[4];
    so3_rows_rot(W, sa, Ca, Pa, qrot_inv(RbT, Jr), ja);  // row r of Jrot mat(R_b^T)
    so3_rows_logrt(W, sb, Cb, -Jr, jb);                  // row r of -Jrot
    for (int k = 0; k < 4; k++) {
      const int ga = sc.knot_col[sa + k], gb = sc.knot_col[sb + k];
      if (ga >= 0) {
        o->add(r, ga + 0, info[r] * ja[k].x);
        o->add(r, ga + 1, info[r] * ja[k].y);
        o->add(r, ga + 2, info[r] * ja[k].z);
      }
      if (gb >= 0) {
        o->add(r, gb + 0, info[r] * jb[k].x);
        o->add(r, gb + 1, info[r] * jb[k].y);
        o->add(r, gb + 2, info[r] * jb[k].z);
      }
    }
  }
  o->r[0] = info[0] * res.x; o->r[1] = info[1] * res.y; o->r[2] = info[2] * res.z;
  return true;
}

// ImageDepthFactor::Evaluate (image_feature_factor.h:666-693)
CLIC_HD bool small_image_depth_eval(const SmallFactor& f, double d_inv, int g /* column of the inverse depth, -1 constant */,
                                    SmallOut* o) {
  const V3 p_i = ldv(f.d), p_j = ldv(f.d + 3);
  const Q4 S = ldq(f.d + 6);
  const V3 p_CiinCj = ldv(f.d + 10);
  const double sqrt_info = f.d[13];
  const V3 x_ci = mk(p_i.x / d_inv, p_i.y / d_inv, p_i.z / d_inv);
  const V3 x_j = qrot(S, x_ci) + p_CiinCj;
  const double dj = 1.0 / x_j.z;
  o->reset(2);
  if (g >= 0) {
    const V3 Jd = (-1.0 / d_inv) * qrot(S, x_ci);
    o->add(0, g, sqrt_info * (dj * Jd.x - dj * dj * x_j.x * Jd.z));
    o->add(1, g, sqrt_info * (dj * Jd.y - dj * dj * x_j.y * Jd.z));
  }
  o->r[0] = sqrt_info * (x_j.x * dj - p_j.x);
  o->r[1] = sqrt_info * (x_j.y * dj - p_j.y);
  return true;
}

// ---- the factors below take the values of their non-knot parameter blocks in x (the reference's block order) and the
// first column of each block in xcol (-1: constant); clic_evaluate_factor and the solve both go through them ----
CLIC_HD V3 m3row(const M3& m, int r) { return mk(m.m[3 * r], m.m[3 * r + 1], m.m[3 * r + 2]); }
CLIC_HD V3 unit3(int r) { return mk(r == 0 ? 1.0 : 0.0, r == 1 ? 1.0 : 0.0, r == 2 ? 1.0 : 0.0); }
CLIC_HD double comp(V3 v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }
CLIC_HD void small_add3(SmallOut* o, int row, int g, V3 v, double w) {
  if (g < 0) return;
  o->add(row, g + 0, w * v.x);
  o->add(row, g + 1, w * v.y);
  o->add(row, g + 2, w * v.z);
}
CLIC_HD V3 spline_pos(const WindowView& W, int s, const double c[4]) {
  V3 p = mk(0, 0, 0);
  for (int k = 0; k < 4; k++) p = fma3(c[k], ldv(W.kp + 3 * (s + k)), p);
  return p;
}
// the two rows of J_v (projection onto the normalised plane) of a camera-frame point
CLIC_HD V3 proj_row(V3 x_j, double dj, int r) {
  return r == 0 ? mk(dj, 0.0, -dj * dj * x_j.x) : mk(0.0, dj, -dj * dj * x_j.y);
}

// Image3D2DFactor::Evaluate (image_feature_factor.h:322-452).  x = p_G[3], t_offset_ns; xcol = {p_G, t_offset}.
CLIC_HD bool small_image_3d2d_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                   Q4 S_CtoI, V3 p_CinI, const double* x, const int* xcol, const SmallCols& sc, SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a + (long long)x[3], t0_ns, dt_ns, K, &s, &u)) return false;
  const double inv_dt = 1e9 / double(dt_ns);
  const V3 p_j = ldv(f.d), p_G = ldv(x);
  const double sqrt_info = f.d[3];
  So3Chain C;
  Q4 P[4];
  so3_chain_rtp(W, s, u, &C, P);
  const Q4 S_GtoIj = qconj(C.R), S_ItoC = qconj(S_CtoI);
  const Q4 S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
  double c[4];
  blend_plain(u, c);
  const V3 rel = p_G - spline_pos(W, s, c);
  const V3 x_j = qrot(S_GtoCj, rel) - qrot(S_ItoC, p_CinI);
  const double dj = 1.0 / x_j.z;
  o->reset(2);
  for (int r = 0; r < 2; r++) {
    const V3 b = qrot_inv(S_GtoCj, proj_row(x_j, dj, r));  // row r of J_v mat(S_GtoCj)
    V3 j[4];
    so3_rows_rtp(W, s, C, P, cross(b, rel), j);
    for (int k = 0; k < 4; k++) {
      const int g = sc.knot_col[s + k];
      if (g < 0) continue;
      small_add3(o, r, g, j[k], sqrt_info);
      small_add3(o, r, g + 3, b, -sqrt_info * c[k]);
    }
    small_add3(o, r, xcol[0], b, sqrt_info);
  }
  if (xcol[1] >= 0) {  // ":440-451"
    const V3 gyro = so3_velocity_body(W, s, u, inv_dt);
    double d1[4];
    blend_plain_d1(u, inv_dt, d1);
    const V3 vel = spline_pos(W, s, d1);
    // Rj_dot^T v = -(gyro x R^T v)
    const V3 J_tj = qrot(S_ItoC, -cross(gyro, qrot(S_GtoIj, rel))) - qrot(S_GtoCj, vel);
    for (int r = 0; r < 2; r++) o->add(r, xcol[1], 1e-9 * sqrt_info * dot(proj_row(x_j, dj, r), J_tj));
  }
  o->r[0] = sqrt_info * (x_j.x * dj - p_j.x);
  o->r[1] = sqrt_info * (x_j.y * dj - p_j.y);
  return true;
}

// ImageFeatureOnePoseFactor::Evaluate (image_feature_factor.h:511-633).  x = inverse depth; xcol = {inverse depth}.
CLIC_HD bool small_image_one_pose_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                       Q4 S_CtoI, V3 p_CinI, const double* x, const int* xcol, const SmallCols& sc,
                                       SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &s, &u)) return false;
  const V3 p_i = ldv(f.d), p_IiinG = ldv(f.d + 7), p_j = ldv(f.d + 10);
  const Q4 S_IitoG = ldq(f.d + 3);
  const double sqrt_info = f.d[13], d_inv = x[0];
  const V3 x_ci = mk(p_i.x / d_inv, p_i.y / d_inv, p_i.z / d_inv);
  const V3 p_Ii = qrot(S_CtoI, x_ci) + p_CinI;
  const V3 p_G = qrot(S_IitoG, p_Ii) + p_IiinG;
  So3Chain C;
  Q4 P[4];
  so3_chain_rtp(W, s, u, &C, P);
  const Q4 S_GtoIj = qconj(C.R), S_ItoC = qconj(S_CtoI);
  const Q4 S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
  double c[4];
  blend_plain(u, c);
  const V3 rel = p_G - spline_pos(W, s, c);
  const V3 x_j = qrot(S_GtoCj, rel) - qrot(S_ItoC, p_CinI);
  const double dj = 1.0 / x_j.z;
  o->reset(2);
  const V3 J_Xm_d = (-1.0 / d_inv) * qrot(so3_mul(so3_mul(S_GtoCj, S_IitoG), S_CtoI), x_ci);
  for (int r = 0; r < 2; r++) {
    const V3 Jv = proj_row(x_j, dj, r);
    const V3 b = qrot_inv(S_GtoCj, Jv);
    V3 j[4];
    so3_rows_rtp(W, s, C, P, cross(b, rel), j);
    for (int k = 0; k < 4; k++) {
      const int g = sc.knot_col[s + k];
      if (g < 0) continue;
      small_add3(o, r, g, j[k], sqrt_info);
      small_add3(o, r, g + 3, b, -sqrt_info * c[k]);
    }
    if (xcol[0] >= 0) o->add(r, xcol[0], sqrt_info * dot(Jv, J_Xm_d));
  }
  o->r[0] = sqrt_info * (x_j.x * dj - p_j.x);
  o->r[1] = sqrt_info * (x_j.y * dj - p_j.y);
  return true;
}

// EpipolarFactor::Evaluate (image_feature_factor.h:741-822)
CLIC_HD bool small_epipolar_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K, Q4 S_CtoI,
                                 V3 p_CinI, const SmallCols& sc, SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &s, &u)) return false;
  const V3 x_i = ldv(f.d), x_k = ldv(f.d + 3), p_CkinG = ldv(f.d + 10);
  const Q4 S_GtoCk = ldq(f.d + 6);
  const double weight = f.d[13];
  So3Chain C;
  so3_chain_rp(W, s, u, &C);
  const Q4 S_IitoG = C.R;
  double c[4];
  blend_plain(u, c);
  const V3 p_IiinG = spline_pos(W, s, c);
  const V3 xi_I = qrot(S_CtoI, x_i);
  const V3 Rxi = qrot(so3_mul(so3_mul(S_GtoCk, S_IitoG), S_CtoI), x_i);
  const V3 tvec = qrot(S_GtoCk, qrot(S_IitoG, p_CinI) + p_IiinG - p_CkinG);
  o->reset(1);
  o->r[0] = weight * dot(cross(x_k, tvec), Rxi);  // (x_k^T hat(t)) Rxi
  const V3 jac_lhs = qrot_inv(S_GtoCk, cross(x_k, Rxi));  // x_k^T hat(Rxi) mat(S_GtoCk)
  V3 a = cross(qrot_inv(so3_mul(S_GtoCk, S_IitoG), -cross(x_k, tvec)), xi_I);
  a = a + cross(qrot_inv(S_IitoG, jac_lhs), p_CinI);
  V3 j[4];
  so3_rows_rp(W, s, C, a, j);
  for (int k = 0; k < 4; k++) {
    const int g = sc.knot_col[s + k];
    if (g < 0) continue;
    small_add3(o, 0, g, j[k], weight);
    small_add3(o, 0, g + 3, jac_lhs, -weight * c[k]);
  }
  return true;
}

// distance of p to the plane / line of a correspondence and its gradient row (lidar_feature_factor.h:237-250);
// d = point[3] geo_plane[4] geo_normal[3] geo_point[3]
CLIC_HD double small_loam_distance(const SmallFactor& f, V3 p, V3* J_pi) {
  if (f.geo_type == 1) {
    *J_pi = ldv(f.d + 3);
    return dot(p, *J_pi) + f.d[6];
  }
  const V3 normal = ldv(f.d + 7);
  const V3 dist_vec = cross(p - ldv(f.d + 10), normal);
  const double r = sqrt(dot(dist_vec, dist_vec));
  *J_pi = cross((-1.0 / r) * dist_vec, normal);  // (-dist_vec^T / r) hat(normal)
  return r;
}

// LoamFeatureOptMapPoseFactor::Evaluate (lidar_feature_factor.h:206-329).  x = S_ImtoG[4], p_IminG[3];
// xcol = {map rotation, map position}.
CLIC_HD bool small_loam_opt_map_pose_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                          const double* x, const int* xcol, const SmallCols& sc, SmallOut* o) {
  int s = 0;
  double u = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &s, &u)) return false;
  const double weight = f.d[13];
  const Q4 S_LtoI = ldq(f.d + 14), S_ImtoG = ldq(x);
  const V3 p_LinI = ldv(f.d + 18), p_IminG = ldv(x + 4);
  So3Chain C;
  so3_chain_rp(W, s, u, &C);
  const Q4 S_IktoG = C.R;
  double c[4];
  blend_plain(u, c);
  const V3 p_IkinG = spline_pos(W, s, c);
  const Q4 S_ItoL = qconj(S_LtoI), S_GtoIm = qconj(S_ImtoG);
  const Q4 S_GtoM = so3_mul(S_ItoL, S_GtoIm);
  const V3 p_GinM = qrot(S_ItoL, qrot(S_GtoIm, -p_IminG) - p_LinI);
  const V3 p_IK = qrot(S_LtoI, ldv(f.d)) + p_LinI;
  const V3 p_G = qrot(S_IktoG, p_IK) + p_IkinG;
  const V3 p_M = qrot(S_GtoM, p_G) + p_GinM;
  V3 J_pi;
  o->reset(1);
  o->r[0] = weight * small_loam_distance(f, p_M, &J_pi);
  const V3 lhs_P = qrot_inv(S_GtoM, J_pi);                        // J_pi^T mat(S_GtoM)
  const V3 a = cross(qrot_inv(S_IktoG, -lhs_P), p_IK);            // J_pi^T (-mat(S_GtoM) mat(S_IktoG) hat(p_IK))
  V3 j[4];
  so3_rows_rp(W, s, C, a, j);
  for (int k = 0; k < 4; k++) {
    const int g = sc.knot_col[s + k];
    if (g < 0) continue;
    small_add3(o, 0, g, j[k], weight);
    small_add3(o, 0, g + 3, lhs_P, weight * c[k]);
  }
  const V3 Jl = qrot_inv(S_ItoL, J_pi);  // J_pi^T mat(S_ItoL)
  small_add3(o, 0, xcol[0], cross(Jl, qrot(S_GtoIm, p_G - p_IminG)), weight);
  small_add3(o, 0, xcol[1], qrot_inv(S_GtoIm, Jl), -weight);
  return true;
}

// RalativeLoamFeatureFactor::Evaluate (lidar_feature_factor.h:376-526)
CLIC_HD bool small_relative_loam_eval(const WindowView& W, const SmallFactor& f, long long t0_ns, long long dt_ns, int K,
                                      const SmallCols& sc, SmallOut* o) {
  int si = 0, sj = 0;
  double ui = 0.0, uj = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &si, &ui) || !index_ns(f.t_b, t0_ns, dt_ns, K, &sj, &uj)) return false;
  const double weight = f.d[13];
  const Q4 S_LtoI = ldq(f.d + 14);
  const V3 p_LinI = ldv(f.d + 18);
  const V3 p_Ii = qrot(S_LtoI, ldv(f.d)) + p_LinI;
  So3Chain Ci, Cj;
  Q4 Pj[4];
  so3_chain_rp(W, si, ui, &Ci);
  so3_chain_rtp(W, sj, uj, &Cj, Pj);
  double ci[4], cj[4];
  blend_plain(ui, ci);
  blend_plain(uj, cj);
  const Q4 S_IitoG = Ci.R, S_GtoIj = qconj(Cj.R), S_ItoC = qconj(S_LtoI);
  const V3 p_G = qrot(S_IitoG, p_Ii) + spline_pos(W, si, ci);
  const V3 rel = p_G - spline_pos(W, sj, cj);
  const Q4 S_GtoCj = so3_mul(S_ItoC, S_GtoIj);
  const V3 x_j = qrot(S_GtoCj, rel) - qrot(S_ItoC, p_LinI);
  V3 J_pi;
  o->reset(1);
  o->r[0] = weight * small_loam_distance(f, x_j, &J_pi);
  const V3 b = qrot_inv(S_GtoCj, J_pi);  // J_pi^T mat(S_GtoCj)
  V3 ji[4], jj[4];
  so3_rows_rp(W, si, Ci, cross(qrot_inv(S_IitoG, -b), p_Ii), ji);
  so3_rows_rtp(W, sj, Cj, Pj, cross(b, rel), jj);
  for (int k = 0; k < 4; k++) {
    const int gi = sc.knot_col[si + k], gj = sc.knot_col[sj + k];
    if (gi >= 0) {
      small_add3(o, 0, gi, ji[k], weight);
      small_add3(o, 0, gi + 3, b, weight * ci[k]);
    }
    if (gj >= 0) {
      small_add3(o, 0, gj, jj[k], weight);
      small_add3(o, 0, gj + 3, b, -weight * cj[k]);
    }
  }
  return true;
}

// PreIntegrationFactor::Evaluate (trajectory_value_factor.h:651-964) with IntegrationBase::evaluate333
// (integration_base.h:207-252).  x = bg_a[3] bg_b[3] ba_a[3] ba_b[3]; xcol = their first columns.
CLIC_HD bool small_preintegration_eval(const WindowView& W, const SmallFactor& f, const PreintData& pre, long long t0_ns,
                                       long long dt_ns, int K, const double* x, const int* xcol, const SmallCols& sc,
                                       SmallOut* o) {
  int sa = 0, sb = 0;
  double ua = 0.0, ub = 0.0;
  if (!index_ns(f.t_a, t0_ns, dt_ns, K, &sa, &ua) || !index_ns(f.t_b, t0_ns, dt_ns, K, &sb, &ub)) return false;
  const double inv_dt = 1e9 / double(dt_ns);
  const V3 bg_a = ldv(x), bg_b = ldv(x + 3), ba_a = ldv(x + 6), ba_b = ldv(x + 9);
  So3Chain Ca, Cb;
  Q4 Pa[4];
  so3_chain_rtp(W, sa, ua, &Ca, Pa);  // EvaluateRTp at t_a
  so3_chain_rp(W, sb, ub, &Cb);       // EvaluateRp at t_b
  double ca[4], cb[4], da[4], db[4];
  blend_plain(ua, ca);
  blend_plain(ub, cb);
  blend_plain_d1(ua, inv_dt, da);
  blend_plain_d1(ub, inv_dt, db);
  const V3 p_a = spline_pos(W, sa, ca), p_b = spline_pos(W, sb, cb);
  const V3 v_a = spline_pos(W, sa, da), v_b = spline_pos(W, sb, db);
  const Q4 S_IatoG = Ca.R, S_GtoIa = qconj(Ca.R), S_IbtoG = Cb.R;
  const double dt = pre.sum_dt;
  const V3 gravity = mk(0.0, 0.0, -pre.gravity_norm);
  // evaluate333
  const V3 dba = ba_a - ldv(pre.linearized_ba), dbg = bg_a - ldv(pre.linearized_bg);
  M3 dp_dba, dp_dbg, dq_dbg, dv_dba, dv_dbg;
  for (int i = 0; i < 9; i++) {
    dp_dba.m[i] = pre.dp_dba[i]; dp_dbg.m[i] = pre.dp_dbg[i]; dq_dbg.m[i] = pre.dq_dbg[i];
    dv_dba.m[i] = pre.dv_dba[i]; dv_dbg.m[i] = pre.dv_dbg[i];
  }
  const V3 cdp = ldv(pre.delta_p) + mvec(dp_dba, dba) + mvec(dp_dbg, dbg);
  Q4 cdq = qmul(ldq(pre.delta_q), so3_exp(mvec(dq_dbg, dbg)));
  {
    const double n = sqrt(cdq.x * cdq.x + cdq.y * cdq.y + cdq.z * cdq.z + cdq.w * cdq.w);
    cdq.x /= n; cdq.y /= n; cdq.z /= n; cdq.w /= n;
  }
  const V3 cdv = ldv(pre.delta_v) + mvec(dv_dba, dba) + mvec(dv_dbg, dbg);
  const V3 r_p_part = p_b - p_a - dt * v_a + (0.5 * dt * dt) * gravity;
  const V3 r_v_part = v_b - v_a + dt * gravity;
  const V3 res_p = qrot(S_GtoIa, (0.5 * dt * dt) * gravity + p_b - p_a - dt * v_a) - cdp;
  const V3 res_R = so3_log(so3_mul(so3_mul(qconj(cdq), S_GtoIa), S_IbtoG));
  const V3 res_v = qrot(S_GtoIa, dt * gravity + v_b - v_a) - cdv;
  double raw[15];
  raw[0] = res_p.x; raw[1] = res_p.y; raw[2] = res_p.z;
  raw[3] = res_R.x; raw[4] = res_R.y; raw[5] = res_R.z;
  raw[6] = res_v.x; raw[7] = res_v.y; raw[8] = res_v.z;
  raw[9] = ba_b.x - ba_a.x; raw[10] = ba_b.y - ba_a.y; raw[11] = ba_b.z - ba_a.z;
  raw[12] = bg_b.x - bg_a.x; raw[13] = bg_b.y - bg_a.y; raw[14] = bg_b.z - bg_a.z;
  o->reset(15);
  // unweighted rows
  const M3 Jrot = right_jacobian_inv(res_R);
  M3 dR_dbg;
  {
    const V3 delta_bg = mvec(dq_dbg, dbg);
    dR_dbg = mscale(-1.0, mmul(mmul(left_jacobian_inv(res_R), left_jacobian(-delta_bg)), dq_dbg));
  }
  for (int r = 0; r < 3; r++) {
    const V3 b = qrot(S_IatoG, unit3(r));  // row r of mat(S_GtoIa)
    const V3 Jr = m3row(Jrot, r);
    V3 jp[4], jra[4], jrb[4], jv[4];
    so3_rows_rtp(W, sa, Ca, Pa, cross(b, r_p_part), jp);             // [1] position rows, rotation knots of t_a
    so3_rows_rtp(W, sa, Ca, Pa, -qrot(S_IbtoG, Jr), jra);            // [2] rotation rows: -Jrot mat(S_GtoIb)
    so3_rows_rp(W, sb, Cb, Jr, jrb);                                 //     ... and +Jrot at t_b
    so3_rows_rtp(W, sa, Ca, Pa, cross(b, r_v_part), jv);             // [3] velocity rows
    for (int k = 0; k < 4; k++) {
      const int ga = sc.knot_col[sa + k], gb = sc.knot_col[sb + k];
      if (ga >= 0) {
        small_add3(o, 0 + r, ga, jp[k], 1.0);
        small_add3(o, 3 + r, ga, jra[k], 1.0);
        small_add3(o, 6 + r, ga, jv[k], 1.0);
        small_add3(o, 0 + r, ga + 3, b, -(ca[k] + da[k] * dt));
        small_add3(o, 6 + r, ga + 3, b, -da[k]);
      }
      if (gb >= 0) {
        small_add3(o, 3 + r, gb, jrb[k], 1.0);
        small_add3(o, 0 + r, gb + 3, b, cb[k]);
        small_add3(o, 6 + r, gb + 3, b, db[k]);
      }
    }
    small_add3(o, 0 + r, xcol[0], m3row(dp_dbg, r), -1.0);  // gyro bias t_a
    small_add3(o, 3 + r, xcol[0], m3row(dR_dbg, r), 1.0);
    small_add3(o, 6 + r, xcol[0], m3row(dv_dbg, r), -1.0);
    if (xcol[0] >= 0) o->add(12 + r, xcol[0] + r, -1.0);
    if (xcol[1] >= 0) o->add(12 + r, xcol[1] + r, 1.0);     // gyro bias t_b
    small_add3(o, 0 + r, xcol[2], m3row(dp_dba, r), -1.0);  // accel bias t_a
    small_add3(o, 6 + r, xcol[2], m3row(dv_dba, r), -1.0);
    if (xcol[2] >= 0) o->add(9 + r, xcol[2] + r, -1.0);
    if (xcol[3] >= 0) o->add(9 + r, xcol[3] + r, 1.0);      // accel bias t_b
  }
  // weight: r <- sqrt_info r, J <- sqrt_info J (":746-748, 940-960")
  for (int r = 0; r < 15; r++) {
    double a = 0.0;
    for (int k = 0; k < 15; k++) a += pre.sqrt_info[15 * r + k] * raw[k];
    o->r[r] = a;
  }
  for (int c = 0; c < o->ncols; c++) {
    double col[15];
    for (int k = 0; k < 15; k++) col[k] = o->J[k][c];
    for (int r = 0; r < 15; r++) {
      double a = 0.0;
      for (int k = 0; k < 15; k++) a += pre.sqrt_info[15 * r + k] * col[k];
      o->J[r][c] = a;
    }
  }
  return true;
}

struct ImageShared {
  Q4 S_CtoI; V3 p_CinI;
  double sqrt_info;  // 450 / 1.5, image_feature_factor.h:277-278
  double loss_a;     // CauchyLoss(2) (1 when marginalised), trajectory_estimator.cpp:806-810
  double inv_dt;
  bool want_toff;    // camera time offset unlocked
};

// ---- the two spline poses of a visual factor: where R, p, the blend weights, the vector-first Jacobian rows and the
// body velocities come from.  ImageChainPoses evaluates the cumulative chain per factor (what the reference does);
// ImagePackPoses reads per-timestamp "pose packs" shared by every feature of a camera frame. ----
struct ImageChainPoses {
  const WindowView& W;
  int si, sj;
  double ui, uj, inv_dt;
  So3Chain Ci, Cj;
  Q4 Pj[4];
  CLIC_HD ImageChainPoses(const WindowView& W_, int si_, double ui_, int sj_, double uj_, double inv_dt_)
      : W(W_), si(si_), sj(sj_), ui(ui_), uj(uj_), inv_dt(inv_dt_) {
    so3_chain_rp(W, si, ui, &Ci);
    so3_chain_rtp(W, sj, uj, &Cj, Pj);
  }
  CLIC_HD Q4 Ri() const { return Ci.R; }
  CLIC_HD Q4 Rj() const { return Cj.R; }
  CLIC_HD void coeff_i(double c[4]) const { blend_plain(ui, c); }
  CLIC_HD void coeff_j(double c[4]) const { blend_plain(uj, c); }
  CLIC_HD V3 pos_i() const {
    double c[4];
    blend_plain(ui, c);
    V3 p = mk(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; k++) p = fma3(c[k], ldv(W.kp + 3 * (si + k)), p);
    return p;
  }
  CLIC_HD V3 pos_j() const {
    double c[4];
    blend_plain(uj, c);
    V3 p = mk(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; k++) p = fma3(c[k], ldv(W.kp + 3 * (sj + k)), p);
    return p;
  }
  CLIC_HD void rows_i(V3 a, V3 j[4]) const { so3_rows_rp(W, si, Ci, a, j); }
  CLIC_HD void rows_j(V3 a, V3 j[4]) const { so3_rows_rtp(W, sj, Cj, Pj, a, j); }
  CLIC_HD static V3 lin_vel(const WindowView& W, int s, double u, double inv_dt) {
    double d[4];
    blend_plain_d1(u, inv_dt, d);
    V3 v = mk(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; k++) v = fma3(d[k], ldv(W.kp + 3 * (s + k)), v);
    return v;
  }
  CLIC_HD V3 gyro_i() const { return so3_velocity_body(W, si, ui, inv_dt); }
  CLIC_HD V3 gyro_j() const { return so3_velocity_body(W, sj, uj, inv_dt); }
  CLIC_HD V3 vel_i() const { return lin_vel(W, si, ui, inv_dt); }
  CLIC_HD V3 vel_j() const { return lin_vel(W, sj, uj, inv_dt); }
};

// Pose pack of one timestamp in one role (i: anchor frame, RP chain; j: observing frame, RTP chain):
//   q[4] | pos[3] | c[4] | G[4][3][3] | gyro[3] | vel[3]      with  rows(a)[k] = a^T G[k]  (the rows are linear in a)
constexpr int kPackQ = 0, kPackPos = 4, kPackC = 7, kPackG = 11, kPackGyro = 47, kPackVel = 50, kPackDoubles = 54;
// One call fills the row `c` (0..2) of every G[k]; the call with c == 0 also fills the other fields.
// Rows c0 .. c1-1 of every G[k] from ONE evaluation of the chain; the call with c0 == 0 also fills the other fields.
CLIC_HD void compute_pose_pack_rows(const WindowView& W, int s, double u, int role_j, double inv_dt, int c0, int c1,
                                    double* pack) {
  So3Chain C;
  Q4 P[4];
  if (role_j) so3_chain_rtp(W, s, u, &C, P); else so3_chain_rp(W, s, u, &C);
  for (int c = c0; c < c1; c++) {
    const V3 e = mk(c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0);
    V3 j[4];
    if (role_j) so3_rows_rtp(W, s, C, P, e, j); else so3_rows_rp(W, s, C, e, j);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      pack[kPackG + 9 * k + 3 * c + 0] = j[k].x;
      pack[kPackG + 9 * k + 3 * c + 1] = j[k].y;
      pack[kPackG + 9 * k + 3 * c + 2] = j[k].z;
    }
  }
  const int c = c0;
  if (c != 0) return;
  pack[kPackQ + 0] = C.R.x; pack[kPackQ + 1] = C.R.y; pack[kPackQ + 2] = C.R.z; pack[kPackQ + 3] = C.R.w;
  double cc[4];
  blend_plain(u, cc);
  V3 pos = mk(0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    pos = fma3(cc[k], ldv(W.kp + 3 * (s + k)), pos);
    pack[kPackC + k] = cc[k];
  }
  pack[kPackPos + 0] = pos.x; pack[kPackPos + 1] = pos.y; pack[kPackPos + 2] = pos.z;
  const V3 g = so3_velocity_body(W, s, u, inv_dt), v = ImageChainPoses::lin_vel(W, s, u, inv_dt);
  pack[kPackGyro + 0] = g.x; pack[kPackGyro + 1] = g.y; pack[kPackGyro + 2] = g.z;
  pack[kPackVel + 0] = v.x; pack[kPackVel + 1] = v.y; pack[kPackVel + 2] = v.z;
}
CLIC_HD void compute_pose_pack(const WindowView& W, int s, double u, int role_j, double inv_dt, int c, double* pack) {
  compute_pose_pack_rows(W, s, u, role_j, inv_dt, c, c + 1, pack);
}
struct ImagePackPoses {
  const double* A;  // pack of t_i (role i)
  const double* B;  // pack of t_j (role j)
  CLIC_HD Q4 Ri() const { return ldq(A + kPackQ); }
  CLIC_HD Q4 Rj() const { return ldq(B + kPackQ); }
  CLIC_HD void coeff_i(double c[4]) const { for (int k = 0; k < 4; k++) c[k] = A[kPackC + k]; }
  CLIC_HD void coeff_j(double c[4]) const { for (int k = 0; k < 4; k++) c[k] = B[kPackC + k]; }
  CLIC_HD V3 pos_i() const { return ldv(A + kPackPos); }
  CLIC_HD V3 pos_j() const { return ldv(B + kPackPos); }
  CLIC_HD static void rows(const double* G, V3 a, V3 j[4]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double* g = G + 9 * k;
      j[k].x = a.x * g[0] + a.y * g[3] + a.z * g[6];
      j[k].y = a.x * g[1] + a.y * g[4] + a.z * g[7];
      j[k].z = a.x * g[2] + a.y * g[5] + a.z * g[8];
    }
  }
  CLIC_HD void rows_i(V3 a, V3 j[4]) const { rows(A + kPackG, a, j); }
  CLIC_HD void rows_j(V3 a, V3 j[4]) const { rows(B + kPackG, a, j); }
  CLIC_HD V3 gyro_i() const { return ldv(A + kPackGyro); }
  CLIC_HD V3 gyro_j() const { return ldv(B + kPackGyro); }
  CLIC_HD V3 vel_i() const { return ldv(A + kPackVel); }
  CLIC_HD V3 vel_j() const { return ldv(B + kPackVel); }
};

// ImageFeatureFactor::Evaluate (image_feature_factor.h:67-268), inverse depth held constant (fixed_depth = true, the
// only mode the shipped pipeline uses, trajectory_manager.cpp:709).  Emits 2 rows of 56 columns:
//   [window of t_i: 0..23 | window of t_j: 24..47 | camera t_offset 48 | r 49 | pad]
// `row_sel` (0 | 1) selects which of the two residual rows this call emits: the kernel gives row 0 to lanes 0-15 and
// row 1 to lanes 16-31 of the same 16 factors (data-dependent only, no divergent code), the host test calls it twice.
// free_depth: the landmark's inverse depth is a parameter; its Jacobian (image_feature_factor.h:238-247) goes to
// local column 50 (0 otherwise).
template <class Poses, class Sink>
CLIC_HD double image_eval_with(const Poses& PS, V3 p_i, V3 p_j, double d_inv, const ImageShared& sh, bool want_jac,
                               Sink& sink, int row_sel, bool free_depth = false) {
  V3 x_ci = mk(p_i.x / d_inv, p_i.y / d_inv, p_i.z / d_inv);
  V3 p_Ii = qrot(sh.S_CtoI, x_ci) + sh.p_CinI;
  const Q4 Ri = PS.Ri(), Rj = PS.Rj();
  double ci[4], cj[4];
  PS.coeff_i(ci);
  PS.coeff_j(cj);
  const V3 pos_i = PS.pos_i(), pos_j = PS.pos_j();
  V3 p_G = qrot(Ri, p_Ii) + pos_i;
  Q4 S_ItoC = qconj(sh.S_CtoI);
  Q4 S_GtoCj = qmul(S_ItoC, qconj(Rj));  // S_ItoC * S_GtoIj
  V3 dG = p_G - pos_j;
  V3 x_j = qrot(S_GtoCj, dG) - qrot(S_ItoC, sh.p_CinI);
  const double dinv = 1.0 / x_j.z;
  double res[2] = {sh.sqrt_info * (x_j.x * dinv - p_j.x), sh.sqrt_info * (x_j.y * dinv - p_j.y)};
  const double sq = res[0] * res[0] + res[1] * res[1];
  double rho = sq, sc = 1.0;
  if (sh.loss_a > 0) sc = cauchy_scale(sq, sh.loss_a, &rho);
  if (!want_jac) return rho;
  // J_v row (1x3 of the 2x3): row 0 = (1/z, 0, -x/z^2), row 1 = (0, 1/z, -y/z^2)
  const bool r0 = row_sel == 0;
  const V3 Jvr = mk(r0 ? dinv : 0.0, r0 ? 0.0 : dinv, -dinv * dinv * (r0 ? x_j.x : x_j.y));
  V3 jt = mk(0, 0, 0);
  if (sh.want_toff) {  // image_feature_factor.h:250-263
    const V3 gyro_i = PS.gyro_i(), gyro_j = PS.gyro_j(), vel_i = PS.vel_i(), vel_j = PS.vel_j();
    // J_tj = S_ItoC * Rj_dot^T * (p_G - p_Ij) - S_GtoCj * vel_j,  Rj_dot^T v = -(omega_j x (Rj^T v))
    V3 J_tj = qrot(S_ItoC, -cross(gyro_j, qrot_inv(Rj, dG))) - qrot(S_GtoCj, vel_j);
    // J_ti = S_GtoCj * (Ri_dot * p_Ii + vel_i),  Ri_dot p = Ri (omega_i x p)
    V3 J_ti = qrot(S_GtoCj, qrot(Ri, cross(gyro_i, p_Ii)) + vel_i);
    jt = J_ti + J_tj;
  }
  {
    const int row = row_sel;
    const double sI = sc * sh.sqrt_info;
    // g = (J_v S_GtoCj)^T  -> jac_lhs_P[0] = g, jac_lhs_P[1] = -g
    V3 g = qrot_inv(S_GtoCj, Jvr);
    // jac_lhs_R[0] = -J_v (S_GtoCj S_IitoG) hat(p_Ii):  with m = (J_v S_GtoCj R_i)^T, row = -(m x p_Ii) = p_Ii x m
    V3 m = qrot_inv(Ri, g);
    V3 a_i = cross(p_Ii, m);
    // jac_lhs_R[1] = J_v S_GtoCj hat(p_G - p_Ij):  row = g x dG
    V3 a_j = cross(g, dG);
    V3 ji[4], jj[4];
    PS.rows_i(a_i, ji);
    PS.rows_j(a_j, jj);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      sink.put(row, 6 * k + 0, sI * ji[k].x);
      sink.put(row, 6 * k + 1, sI * ji[k].y);
      sink.put(row, 6 * k + 2, sI * ji[k].z);
      sink.put(row, 6 * k + 3, sI * ci[k] * g.x);
      sink.put(row, 6 * k + 4, sI * ci[k] * g.y);
      sink.put(row, 6 * k + 5, sI * ci[k] * g.z);
      sink.put(row, 24 + 6 * k + 0, sI * jj[k].x);
      sink.put(row, 24 + 6 * k + 1, sI * jj[k].y);
      sink.put(row, 24 + 6 * k + 2, sI * jj[k].z);
      sink.put(row, 24 + 6 * k + 3, -sI * cj[k] * g.x);
      sink.put(row, 24 + 6 * k + 4, -sI * cj[k] * g.y);
      sink.put(row, 24 + 6 * k + 5, -sI * cj[k] * g.z);
    }
    sink.put(row, 48, sh.want_toff ? (1e-9 * sI) * dot(Jvr, jt) : 0.0);
    sink.put(row, 49, sc * (r0 ? res[0] : res[1]));
    double j_rho = 0.0;
    if (free_depth) {  // J_Xm_d = -(S_GtoCj R_i S_CtoI x_ci) / rho;  S_CtoI x_ci = p_Ii - p_CinI
      const V3 jx = (-1.0 / d_inv) * qrot(S_GtoCj, qrot(Ri, p_Ii - sh.p_CinI));
      j_rho = sI * dot(Jvr, jx);
    }
    sink.put(row, 50, j_rho);
#pragma unroll
    for (int cc = 51; cc < 56; cc++) sink.put(row, cc, 0.0);
    sink.row_done(row);
  }
  return rho;
}
template <class Sink>
CLIC_HD double image_eval(const WindowView& W, int si, double ui, int sj, double uj, V3 p_i, V3 p_j, double d_inv,
                          const ImageShared& sh, bool want_jac, Sink& sink, int row_sel, bool free_depth = false) {
  const ImageChainPoses PS(W, si, ui, sj, uj, sh.inv_dt);
  return image_eval_with(PS, p_i, p_j, d_inv, sh, want_jac, sink, row_sel, free_depth);
}
template <class Sink>
CLIC_HD double image_eval_packed(const double* pack_i, const double* pack_j, V3 p_i, V3 p_j, double d_inv,
                                 const ImageShared& sh, bool want_jac, Sink& sink, int row_sel, bool free_depth = false) {
  ImagePackPoses PS;
  PS.A = pack_i;
  PS.B = pack_j;
  return image_eval_with(PS, p_i, p_j, d_inv, sh, want_jac, sink, row_sel, free_depth);
}

}  // namespace clic
