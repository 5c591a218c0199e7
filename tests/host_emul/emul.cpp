// TEST INFRASTRUCTURE ONLY: compiles the product's __host__ __device__ factor math (clic_b200/csrc/*.cuh) with g++
// so the CPU-only build box can check it against the oracle before GPU time is spent.  Never shipped or loaded
// by the product.
#include <cstring>
#include <vector>

#include "../../clic_b200/csrc/factor_eval.cuh"

using namespace clic;

struct HostWindow {
  std::vector<PairData> pair;
  WindowView view;
  HostWindow(int K, const double* kq, const double* kp) : pair(K - 1) {
    for (int k = 0; k + 1 < K; k++) compute_pair(kq + 4 * k, kq + 4 * (k + 1), &pair[k]);
    view.kq = kq;
    view.kp = kp;
    view.pair = pair.data();
  }
};

struct ArraySink {
  double* out;
  int nc;
  void put(int row, int col, double v) { out[row * nc + col] = v; }
  void row_done(int) {}
};

extern "C" {

__attribute__((visibility("default"))) int emul_loam(int K, const double* kq, const double* kp, int64_t t0_ns,
                                                      int64_t dt_ns, int64_t t_ns, const double* pL, int is_plane,
                                                      const double* geo, const double* S_GtoM, const double* p_GinM,
                                                      const double* S_LtoI, const double* p_LinI, double weight,
                                                      double* r, double* J, int* s_out, double* u_out) {
  HostWindow W(K, kq, kp);
  int s;
  double u;
  if (!index_ns(t_ns, t0_ns, dt_ns, K, &s, &u)) return 1;
  LoamShared sh;
  sh.S_GtoM = ldq(S_GtoM); sh.p_GinM = ldv(p_GinM); sh.S_LtoI = ldq(S_LtoI); sh.p_LinI = ldv(p_LinI); sh.weight = weight;
  sh.gm_identity = false;
  sh.point_in_imu = false;
  loam_eval(W.view, s, u, ldv(pL), is_plane != 0, geo, sh, true, r, J, weight);
  *s_out = s;
  *u_out = u;
  return 0;
}

__attribute__((visibility("default"))) int emul_imu(int K, const double* kq, const double* kp, int64_t t0_ns,
                                                     int64_t dt_ns, int64_t t_corrected_ns, const double* gyro,
                                                     const double* accel, const double* info, const double* bg,
                                                     const double* ba, const double* gravity, double loss_a, int marg,
                                                     double* rows /*6*NC*/, double* rho, int* s_out) {
  HostWindow W(K, kq, kp);
  int s;
  double u;
  if (!index_ns(t_corrected_ns, t0_ns, dt_ns, K, &s, &u)) return 1;
  ImuShared sh;
  for (int i = 0; i < 6; i++) sh.info[i] = info[i];
  sh.bg = ldv(bg); sh.ba = ldv(ba); sh.gravity = ldv(gravity); sh.loss_a = loss_a; sh.inv_dt = 1e9 / double(dt_ns);
  if (marg) {
    ArraySink sink{rows, 40};
    *rho = imu_eval<true>(W.view, s, u, ldv(gyro), ldv(accel), sh, true, sink);
  } else {
    ArraySink sink{rows, 32};
    *rho = imu_eval<false>(W.view, s, u, ldv(gyro), ldv(accel), sh, true, sink);
  }
  *s_out = s;
  return 0;
}

__attribute__((visibility("default"))) int emul_image(int K, const double* kq, const double* kp, int64_t t0_ns,
                                                       int64_t dt_ns, int64_t ti_ns, int64_t tj_ns, const double* p_i,
                                                       const double* p_j, double d_inv, const double* S_CtoI,
                                                       const double* p_CinI, double loss_a, int want_toff,
                                                       double* rows /*2*56*/, double* rho, int* si_out, int* sj_out) {
  HostWindow W(K, kq, kp);
  int si, sj;
  double ui, uj;
  if (!index_ns(ti_ns, t0_ns, dt_ns, K, &si, &ui) || !index_ns(tj_ns, t0_ns, dt_ns, K, &sj, &uj)) return 1;
  ImageShared sh;
  sh.S_CtoI = ldq(S_CtoI); sh.p_CinI = ldv(p_CinI); sh.sqrt_info = 450.0 / 1.5; sh.loss_a = loss_a;
  sh.inv_dt = 1e9 / double(dt_ns); sh.want_toff = want_toff != 0;
  ArraySink sink{rows, 56};
  // local column 50 = d r / d(inverse depth) (free_depth = true); callers with a fixed depth ignore it
  *rho = image_eval(W.view, si, ui, sj, uj, ldv(p_i), ldv(p_j), d_inv, sh, true, sink, 0, true);
  *rho = image_eval(W.view, si, ui, sj, uj, ldv(p_i), ldv(p_j), d_inv, sh, true, sink, 1, true);
  *si_out = si;
  *sj_out = sj;
  return 0;
}
// same factor through per-timestamp pose packs (the packed visual kernel's path)
__attribute__((visibility("default"))) int emul_image_packed(int K, const double* kq, const double* kp, int64_t t0_ns,
                                                              int64_t dt_ns, int64_t ti_ns, int64_t tj_ns,
                                                              const double* p_i, const double* p_j, double d_inv,
                                                              const double* S_CtoI, const double* p_CinI,
                                                              double loss_a, int want_toff, double* rows /*2*56*/,
                                                              double* rho) {
  HostWindow W(K, kq, kp);
  int si, sj;
  double ui, uj;
  if (!index_ns(ti_ns, t0_ns, dt_ns, K, &si, &ui) || !index_ns(tj_ns, t0_ns, dt_ns, K, &sj, &uj)) return 1;
  ImageShared sh;
  sh.S_CtoI = ldq(S_CtoI); sh.p_CinI = ldv(p_CinI); sh.sqrt_info = 450.0 / 1.5; sh.loss_a = loss_a;
  sh.inv_dt = 1e9 / double(dt_ns); sh.want_toff = want_toff != 0;
  double pa[kPackDoubles], pb[kPackDoubles];
  for (int c = 0; c < 3; c++) {
    compute_pose_pack(W.view, si, ui, 0, sh.inv_dt, c, pa);
    compute_pose_pack(W.view, sj, uj, 1, sh.inv_dt, c, pb);
  }
  ArraySink sink{rows, 56};
  *rho = image_eval_packed(pa, pb, ldv(p_i), ldv(p_j), d_inv, sh, true, sink, 0, true);
  *rho = image_eval_packed(pa, pb, ldv(p_i), ldv(p_j), d_inv, sh, true, sink, 1, true);
  return 0;
}
}

extern "C" {
// The one-thread evaluators of the small factors: kind, the factor's numbers, the layout's column
// maps; returns r and the dense Jacobian over D global columns.
__attribute__((visibility("default"))) int emul_small(int K, const double* kq, const double* kp, int64_t t0_ns, int64_t dt_ns,
                                                       int kind, int64_t t_a, int64_t t_b, const double* d16, double loss_a,
                                                       int landmark, double toff_imu, double inv_depth, const int* knot_col,
                                                       int col_toff_imu, const int* lm_col, int D, double* r /*6*/,
                                                       double* J /*6*D*/, int* nres) {
  HostWindow W(K, kq, kp);
  SmallFactor f{};
  f.kind = kind;
  f.t_a = t_a;
  f.t_b = t_b;
  f.landmark = landmark;
  f.loss_a = loss_a;
  for (int i = 0; i < 16; i++) f.d[i] = d16[i];
  SmallCols sc;
  sc.knot_col = knot_col;
  sc.col_toff_imu = col_toff_imu;
  sc.lm_col = lm_col;
  sc.col_bias_gyro = sc.col_bias_accel = -1;
  SmallOut o;
  bool ok = false;
  switch (kind) {
    case kSmallPose: ok = small_pose_eval(W.view, f, t0_ns, dt_ns, K, toff_imu, sc, &o); break;
    case kSmallLocalVelocity: ok = small_local_velocity_eval(W.view, f, t0_ns, dt_ns, K, toff_imu, sc, &o); break;
    case kSmallRelativeRotation: ok = small_relative_rotation_eval(W.view, f, t0_ns, dt_ns, K, sc, &o); break;
    default: ok = small_image_depth_eval(f, inv_depth, lm_col ? lm_col[landmark] : -1, &o); break;
  }
  if (!ok) return 1;
  *nres = o.nres;
  for (int i = 0; i < o.nres; i++) {
    r[i] = o.r[i];
    for (int c = 0; c < D; c++) J[i * D + c] = 0.0;
    for (int c = 0; c < o.ncols; c++) J[i * D + o.gcol[c]] = o.J[i][c];
  }
  return 0;
}
}

#include "../../include/clic_cuda.h"

extern "C" {
// One small analytic factor (any CLIC_FACTOR_* kind) against the window with every knot free: the host-compiled twin of
// factor_jacobian_kernel (kernels.cuh).  J: row-major nres x n_cols, n_cols = 6 K + the widths of the blocks in x.
__attribute__((visibility("default"))) int emul_factor(int K, const double* kq, const double* kp, int64_t t0_ns, int64_t dt_ns,
                                                        const clic_factor_desc* fd, const double* S_CtoI, const double* p_CinI,
                                                        int n_cols, double* r /*15*/, double* J, int* nres) {
  static const int kWidths[10][4] = {{1, 0, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {1, 0, 0, 0}, {3, 3, 3, 3},
                                     {3, 1, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {3, 3, 0, 0}, {0, 0, 0, 0}};
  if (fd->kind < 0 || fd->kind > 9) return 2;
  HostWindow W(K, kq, kp);
  SmallFactor f{};
  f.kind = fd->kind;
  f.geo_type = fd->geo_type;
  f.t_a = fd->t_a_ns;
  f.t_b = fd->t_b_ns;
  for (int i = 0; i < 24; i++) f.d[i] = fd->d[i];
  int xcol[4], cols = 6 * K;
  for (int k = 0; k < 4; k++) {
    xcol[k] = kWidths[fd->kind][k] > 0 ? cols : -1;
    cols += kWidths[fd->kind][k];
  }
  if (cols != n_cols) return 3;
  std::vector<int> knot_col(K);
  for (int k = 0; k < K; k++) knot_col[k] = 6 * k;
  SmallCols sc;
  sc.knot_col = knot_col.data();
  sc.col_toff_imu = -1;
  sc.lm_col = nullptr;
  sc.col_bias_gyro = sc.col_bias_accel = -1;
  PreintData pre{};
  if (fd->pre) {
    const clic_preintegration* p = fd->pre;
    pre.sum_dt = p->sum_dt;
    pre.gravity_norm = p->gravity_norm;
    for (int i = 0; i < 3; i++) {
      pre.delta_p[i] = p->delta_p[i]; pre.delta_v[i] = p->delta_v[i];
      pre.linearized_ba[i] = p->linearized_ba[i]; pre.linearized_bg[i] = p->linearized_bg[i];
    }
    for (int i = 0; i < 4; i++) pre.delta_q[i] = p->delta_q[i];
    for (int a = 0; a < 3; a++)
      for (int c = 0; c < 3; c++) {
        pre.dp_dba[3 * a + c] = p->jacobian[(0 + a) * 15 + 9 + c];
        pre.dp_dbg[3 * a + c] = p->jacobian[(0 + a) * 15 + 12 + c];
        pre.dq_dbg[3 * a + c] = p->jacobian[(3 + a) * 15 + 12 + c];
        pre.dv_dba[3 * a + c] = p->jacobian[(6 + a) * 15 + 9 + c];
        pre.dv_dbg[3 * a + c] = p->jacobian[(6 + a) * 15 + 12 + c];
      }
    for (int i = 0; i < 225; i++) pre.sqrt_info[i] = p->sqrt_info[i];
  }
  const Q4 qc = ldq(S_CtoI);
  const V3 pc = ldv(p_CinI);
  const double* x = fd->x;
  std::vector<SmallOut> ob(1);
  SmallOut& o = ob[0];
  bool ok = false;
  switch (f.kind) {
    case kSmallPose: sc.col_toff_imu = xcol[0]; ok = small_pose_eval(W.view, f, t0_ns, dt_ns, K, x[0], sc, &o); break;
    case kSmallLocalVelocity: sc.col_toff_imu = xcol[0]; ok = small_local_velocity_eval(W.view, f, t0_ns, dt_ns, K, x[0], sc, &o); break;
    case kSmallRelativeRotation: ok = small_relative_rotation_eval(W.view, f, t0_ns, dt_ns, K, sc, &o); break;
    case kSmallImageDepth: ok = small_image_depth_eval(f, x[0], xcol[0], &o); break;
    case kSmallPreIntegration: ok = fd->pre && small_preintegration_eval(W.view, f, pre, t0_ns, dt_ns, K, x, xcol, sc, &o); break;
    case kSmallImage3D2D: ok = small_image_3d2d_eval(W.view, f, t0_ns, dt_ns, K, qc, pc, x, xcol, sc, &o); break;
    case kSmallImageOnePose: ok = small_image_one_pose_eval(W.view, f, t0_ns, dt_ns, K, qc, pc, x, xcol, sc, &o); break;
    case kSmallEpipolar: ok = small_epipolar_eval(W.view, f, t0_ns, dt_ns, K, qc, pc, sc, &o); break;
    case kSmallLoamOptMapPose: ok = small_loam_opt_map_pose_eval(W.view, f, t0_ns, dt_ns, K, x, xcol, sc, &o); break;
    default: ok = small_relative_loam_eval(W.view, f, t0_ns, dt_ns, K, sc, &o); break;
  }
  if (!ok) return 1;
  *nres = o.nres;
  for (int i = 0; i < o.nres; i++) {
    r[i] = o.r[i];
    for (int c = 0; c < n_cols; c++) J[(size_t)i * n_cols + c] = 0.0;
    for (int c = 0; c < o.ncols; c++) J[(size_t)i * n_cols + o.gcol[c]] = o.J[i][c];
  }
  return 0;
}
}
