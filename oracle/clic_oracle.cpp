// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference arithmetic (oracle).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
//
// clic::TrajectoryEstimator restated without Ceres/Eigen behind the C ABI of include/clic_cuda.h:
//   - problem construction rules           src/estimator/trajectory_estimator.cpp:172-197, 272-292, 369-498, 712-866
//   - loss correction (Ceres Corrector)    src/estimator/factor/analytic_diff/marginalization_factor.cpp:38-61
//   - dense normal equations               marginalization_factor.cpp:122-148
//   - Schur complement / prior             marginalization_factor.cpp:150-276, 320-372
//   - LM trust region                      Ceres 1.14 (un-vendored; README.md:19) TrustRegionMinimizer +
//                                          LevenbergMarquardtStrategy, call site trajectory_estimator.cpp:987-1031.
// PARITY UNPINNED for the LM loop and the eigen-based marginalization: the reference's only test
// (src/app/test_analytic_factor.cpp) never calls Solve() or marginalize(); every Ceres constant below is the
// documented Ceres 1.14 default and is marked "Ceres default".
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <thread>

#include "../include/clic_cuda.h"
#include "factors.hpp"
#include "factors_extra.hpp"
#include "association.hpp"

using namespace oracle;

namespace {

thread_local std::string g_last_error;
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int g_threads = 1;

struct BlockRef {
  int kind;
  int id;  // LOCAL id (local knot index, local bias slot, sensor, landmark)
  bool operator<(const BlockRef& o) const { return kind != o.kind ? kind < o.kind : id < o.id; }
  bool operator==(const BlockRef& o) const { return kind == o.kind && id == o.id; }
};
int block_size(int kind) {
  switch (kind) {
    case CLIC_BLOCK_KNOT_ROT: return 4;
    case CLIC_BLOCK_KNOT_POS: return 3;
    case CLIC_BLOCK_BIAS_GYRO: return 3;
    case CLIC_BLOCK_BIAS_ACCEL: return 3;
    case CLIC_BLOCK_GRAVITY: return 3;
    default: return 1;
  }
}
int local_size(int size) { return size == 4 ? 3 : size; }  // MarginalizationInfo::localSize, marginalization_factor.cpp:118-120

// Canonical ordering key for marginalization (the reference orders by unordered_map iteration, i.e. arbitrarily;
// marginalization_factor.cpp:154-166): knots by id with rot before pos, then bias, gravity, offsets, depths.
struct CanonKey {
  int cat, a, b;
  bool operator<(const CanonKey& o) const {
    if (cat != o.cat) return cat < o.cat;
    if (a != o.a) return a < o.a;
    return b < o.b;
  }
};
CanonKey canon_key(const BlockRef& r) {
  switch (r.kind) {
    case CLIC_BLOCK_KNOT_ROT: return {0, r.id, 0};
    case CLIC_BLOCK_KNOT_POS: return {0, r.id, 1};
    case CLIC_BLOCK_BIAS_GYRO: return {1, 0, r.id};
    case CLIC_BLOCK_BIAS_ACCEL: return {1, 1, r.id};
    case CLIC_BLOCK_GRAVITY: return {2, 0, 0};
    case CLIC_BLOCK_T_OFFSET: return {3, r.id, 0};
    default: return {4, r.id, 0};
  }
}

}  // namespace

// Prior = MarginalizationInfo after marginalize() (marginalization_factor.h:125-142)
struct clic_prior {
  int n = 0, m = 0;
  std::vector<double> J0;  // n*n row-major (linearized_jacobians)
  std::vector<double> r0;  // n
  struct Block {
    int kind, id /*GLOBAL*/, size, idx /*tangent offset in [0,n)*/;
    double x0[4];
  };
  std::vector<Block> blocks;  // keep_block_*
};

namespace {

// MarginalizationFactor — marginalization_factor.cpp:320-372
struct MarginalizationFactor : CostFunction {
  const clic_prior* info;
  explicit MarginalizationFactor(const clic_prior* p) : info(p) {
    for (auto& b : p->blocks) block_sizes.push_back(b.size);
    num_residuals = p->n;
  }
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    int n = info->n;
    std::vector<double> dx(n, 0.0);
    for (size_t i = 0; i < info->blocks.size(); i++) {
      const auto& b = info->blocks[i];
      const double* x = parameters[i];
      const double* x0 = b.x0;
      if (b.size != 4) {
        for (int k = 0; k < b.size; k++) dx[b.idx + k] = x[k] - x0[k];
      } else {
        Quat q0_inv = qconj(Quat{x0[0], x0[1], x0[2], x0[3]});
        // Eigen Quaternion::inverse() = conjugate / squaredNorm
        double n2 = qsqnorm(q0_inv);
        q0_inv = Quat{q0_inv.x / n2, q0_inv.y / n2, q0_inv.z / n2, q0_inv.w / n2};
        Quat qx = Quat{x[0], x[1], x[2], x[3]};
        Quat d = qmul(q0_inv, qx);
        double sgn = (d.w >= 0) ? 1.0 : -1.0;  // ":349-352" (positify() is the identity, eigen_utils.hpp:146-155)
        dx[b.idx + 0] = 2.0 * (sgn * d.x);
        dx[b.idx + 1] = 2.0 * (sgn * d.y);
        dx[b.idx + 2] = 2.0 * (sgn * d.z);
      }
    }
    for (int r = 0; r < n; r++) {
      double acc = 0;
      for (int c = 0; c < n; c++) acc += info->J0[(size_t)r * n + c] * dx[c];
      residuals[r] = info->r0[r] + acc;
    }
    if (jacobians) {
      for (size_t i = 0; i < info->blocks.size(); i++) {
        if (!jacobians[i]) continue;
        const auto& b = info->blocks[i];
        int ls = local_size(b.size);
        for (int r = 0; r < n; r++) {
          for (int c = 0; c < b.size; c++) jacobians[i][(size_t)r * b.size + c] = 0;
          for (int c = 0; c < ls; c++) jacobians[i][(size_t)r * b.size + c] = info->J0[(size_t)r * n + b.idx + c];
        }
      }
    }
    return true;
  }
};

struct ResidualBlock {
  std::unique_ptr<CostFunction> cost;
  std::vector<BlockRef> refs;
  bool has_loss = false;
  double loss_a = 0;
  std::vector<int> drop_set;  // marg only (indices into refs)
  int rtype = 0;
};

struct State {
  std::vector<double> knot_q, knot_p, bias, inv_depth;
  double t_offset_ns[3];
  double gravity[3];
};

}  // namespace

struct clic_estimator {
  // window
  int64_t t0_ns, dt_ns;
  int K, knot_id0, n_bias, bias_id0, n_lm;
  Quat S_CtoI;
  Vec3 p_CinI;
  State x;
  clic_options opt;
  int fixed_index_global = -1;  // fixed_control_point_index_, trajectory_estimator.cpp:145
  std::vector<ResidualBlock> blocks;       // problem_
  std::vector<ResidualBlock> marg_blocks;  // marginalization_info_->factors
  std::vector<char> lm_free;               // landmark added with fixed_depth = 0
  bool bounds_toff[3] = {false, false, false};
  double toff_lo[3], toff_hi[3];
  // index probes: per stream (time_ns, dropped)
  std::vector<std::pair<int64_t, char>> probe[4];

  // ---- trajectory helpers (src/spline/rd_spline.h:119-144, trajectory.h:100-108) ----
  int64_t minTimeNs() const { return t0_ns; }
  int64_t maxTimeNs() const { return t0_ns + (int64_t)(K - N + 1) * dt_ns - 1; }
  double minTime(int sensor) const { return NS_TO_S * minTimeNs() - NS_TO_S * x.t_offset_ns[sensor]; }
  double maxTime(int sensor) const { return NS_TO_S * maxTimeNs() - NS_TO_S * x.t_offset_ns[sensor]; }
  size_t GetCtrlIndex(int64_t t_ns) const { return (size_t)((t_ns - t0_ns) / dt_ns); }

  // TrajectoryEstimator::MeasuredTimeToNs — trajectory_estimator.cpp:272-292
  bool MeasuredTimeToNs(int sensor, double timestamp, int64_t& time_ns) const {
    time_ns = (int64_t)(timestamp * S_TO_NS);
    int64_t t_min_traj_ns = (int64_t)(minTime(sensor) * S_TO_NS);
    int64_t t_max_traj_ns = (int64_t)(maxTime(sensor) * S_TO_NS);
    if (!opt.lock_t_offset[sensor]) {
      if (time_ns - opt.t_offset_padding_ns < t_min_traj_ns || time_ns + opt.t_offset_padding_ns >= t_max_traj_ns)
        return false;
    } else {
      if (time_ns < t_min_traj_ns || time_ns >= t_max_traj_ns) return false;
    }
    return true;
  }

  // Se3Spline::CaculateSplineMeta — src/spline/se3_spline.h:433-467
  void CaculateSplineMeta(const std::vector<std::pair<int64_t, int64_t>>& times, SplineMeta& meta) const {
    int64_t master_dt_ns = dt_ns, master_t0_ns = t0_ns;
    size_t current_segment_start = 0, current_segment_end = 0;
    for (auto tt : times) {
      size_t i1 = GetCtrlIndex(tt.first);
      size_t i2 = GetCtrlIndex(tt.second);
      if (meta.segments.empty() || i1 > current_segment_end) {
        int64_t segment_t0_ns = master_t0_ns + master_dt_ns * (int64_t)i1;
        meta.segments.push_back(SegmentMeta(segment_t0_ns, master_dt_ns));
        current_segment_start = i1;
      } else {
        i1 = current_segment_end + 1;
      }
      auto& seg = meta.segments.back();
      for (size_t i = i1; i < (i2 + N); ++i) seg.n += 1;
      current_segment_end = current_segment_start + seg.n - 1;
    }
  }

  // TrajectoryEstimator::AddControlPoints — trajectory_estimator.cpp:172-197
  void AddControlPoints(const SplineMeta& meta, std::vector<BlockRef>& vec, bool addPosKnot) const {
    for (auto const& seg : meta.segments) {
      size_t start_idx = GetCtrlIndex(seg.t0_ns);
      for (size_t i = start_idx; i < (start_idx + seg.NumParameters()); ++i)
        vec.push_back(BlockRef{addPosKnot ? CLIC_BLOCK_KNOT_POS : CLIC_BLOCK_KNOT_ROT, (int)i});
    }
  }
  // Se3Spline::GetCtrlIdxs(spline_meta) — se3_spline.h:469-481 (returns local ids here)
  void GetCtrlIdxs(const SplineMeta& meta, std::vector<int>& ctrl_ids) const {
    size_t last_ctrl_id = 0;
    for (auto const& seg : meta.segments) {
      size_t start_idx = GetCtrlIndex(seg.t0_ns);
      for (size_t i = 0; i < seg.NumParameters(); i++) {
        if (i + start_idx > last_ctrl_id || (i + start_idx == 0)) ctrl_ids.push_back((int)(start_idx + i));
      }
      last_ctrl_id = (size_t)ctrl_ids.back();
    }
  }

  double* block_ptr(State& s, const BlockRef& r) const {
    switch (r.kind) {
      case CLIC_BLOCK_KNOT_ROT: return &s.knot_q[(size_t)r.id * 4];
      case CLIC_BLOCK_KNOT_POS: return &s.knot_p[(size_t)r.id * 3];
      case CLIC_BLOCK_BIAS_GYRO: return &s.bias[(size_t)r.id * 6];
      case CLIC_BLOCK_BIAS_ACCEL: return &s.bias[(size_t)r.id * 6 + 3];
      case CLIC_BLOCK_GRAVITY: return s.gravity;
      case CLIC_BLOCK_T_OFFSET: return &s.t_offset_ns[r.id];
      default: return &s.inv_depth[(size_t)r.id];
    }
  }

  // ---- column layout ----
  clic_layout layout() const {
    clic_layout L;
    int kf = fixed_index_global - knot_id0 + 1;  // ids <= fixed index are constant, trajectory_estimator.cpp:191
    if (kf < 0) kf = 0;
    if (kf > K) kf = K;
    if (opt.lock_traj) kf = K;
    L.first_free_knot = kf;
    L.n_free_knots = K - kf;
    int c = 6 * L.n_free_knots;
    L.col_bias_gyro = -1;
    L.col_bias_accel = -1;
    if (!opt.lock_wb && n_bias > 0) { L.col_bias_gyro = c; c += 3 * n_bias; }
    if (!opt.lock_ab && n_bias > 0) { L.col_bias_accel = c; c += 3 * n_bias; }
    for (int s = 0; s < 3; s++) L.col_t_offset[s] = -1;
    if (!opt.lock_t_offset[CLIC_SENSOR_IMU]) L.col_t_offset[CLIC_SENSOR_IMU] = c++;
    if (!opt.lock_t_offset[CLIC_SENSOR_CAMERA]) L.col_t_offset[CLIC_SENSOR_CAMERA] = c++;
    L.col_inv_depth = -1;
    L.n_free_landmarks = 0;
    for (int i = 0; i < n_lm; i++) L.n_free_landmarks += lm_free[i] ? 1 : 0;
    if (L.n_free_landmarks > 0) { L.col_inv_depth = c; c += L.n_free_landmarks; }
    L.D = c;
    return L;
  }
  int block_col(const clic_layout& L, const BlockRef& r) const {
    switch (r.kind) {
      case CLIC_BLOCK_KNOT_ROT: return r.id >= L.first_free_knot ? 6 * (r.id - L.first_free_knot) : -1;
      case CLIC_BLOCK_KNOT_POS: return r.id >= L.first_free_knot ? 6 * (r.id - L.first_free_knot) + 3 : -1;
      case CLIC_BLOCK_BIAS_GYRO: return L.col_bias_gyro >= 0 ? L.col_bias_gyro + 3 * r.id : -1;
      case CLIC_BLOCK_BIAS_ACCEL: return L.col_bias_accel >= 0 ? L.col_bias_accel + 3 * r.id : -1;
      case CLIC_BLOCK_GRAVITY: return -1;
      case CLIC_BLOCK_T_OFFSET: return L.col_t_offset[r.id];
      default: {
        if (L.col_inv_depth < 0 || !lm_free[r.id]) return -1;
        int rank = 0;
        for (int i = 0; i < r.id; i++) rank += lm_free[i] ? 1 : 0;
        return L.col_inv_depth + rank;
      }
    }
  }
};

namespace {

// ResidualBlockInfo::Evaluate loss correction — marginalization_factor.cpp:38-61 (== Ceres Corrector).
// Returns rho(s) through *rho0 (needed for the cost 1/2 rho(s), Ceres residual_block.cc).
void correct(const ResidualBlock& rb, int nres, std::vector<double>& res, std::vector<std::vector<double>>& jac,
             const std::vector<int>& sizes, double* rho0) {
  double sq_norm = 0;
  for (int i = 0; i < nres; i++) sq_norm += res[i] * res[i];
  if (!rb.has_loss) {
    *rho0 = sq_norm;
    return;
  }
  double rho[3];
  CauchyLoss(rb.loss_a).Evaluate(sq_norm, rho);
  *rho0 = rho[0];
  double sqrt_rho1_ = std::sqrt(rho[1]);
  double residual_scaling_, alpha_sq_norm_;
  if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
    residual_scaling_ = sqrt_rho1_;
    alpha_sq_norm_ = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling_ = sqrt_rho1_ / (1 - alpha);
    alpha_sq_norm_ = alpha / sq_norm;
  }
  for (size_t b = 0; b < jac.size(); b++) {
    if (jac[b].empty()) continue;
    int sz = sizes[b];
    // J = sqrt_rho1 * (J - alpha_sq_norm * r * (r^T J))
    std::vector<double> rtJ(sz, 0.0);
    for (int r = 0; r < nres; r++)
      for (int c = 0; c < sz; c++) rtJ[c] += res[r] * jac[b][(size_t)r * sz + c];
    for (int r = 0; r < nres; r++)
      for (int c = 0; c < sz; c++)
        jac[b][(size_t)r * sz + c] = sqrt_rho1_ * (jac[b][(size_t)r * sz + c] - alpha_sq_norm_ * res[r] * rtJ[c]);
  }
  for (int i = 0; i < nres; i++) res[i] *= residual_scaling_;
}

struct EvalOut {
  std::vector<double> H, b;
  double cost = 0, fixed_cost = 0;
};

// One residual + Jacobian + normal-equation pass (what Ceres' evaluator + the normal-equation solver do per
// iteration; dense accumulation as in ThreadsConstructA, marginalization_factor.cpp:122-148).
void evaluate(clic_estimator* E, State& s, bool want_jac, EvalOut& out) {
  clic_layout L = E->layout();
  const int D = L.D;
  const size_t nb = E->blocks.size();
  int nthreads = std::max(1, g_threads);
  std::vector<std::vector<double>> Hs(nthreads), bs(nthreads);
  std::vector<double> costs(nthreads, 0.0), fcosts(nthreads, 0.0);
  // factors are sharded contiguously over threads, each with a private H/b summed afterwards in thread
  // order (the reference's own scheme for marginalization, marginalization_factor.cpp:179-204)
  auto worker = [&](int tid) {
    std::vector<double>& H = Hs[tid];
    std::vector<double>& b = bs[tid];
    if (want_jac) {
      H.assign((size_t)D * D, 0.0);
      b.assign(D, 0.0);
    }
    std::vector<const double*> params;
    std::vector<double*> jptr;
    std::vector<std::vector<double>> jac;
    std::vector<double> res;
    std::vector<int> cols;
    double cost = 0, fcost = 0;
    const size_t lo = nb * (size_t)tid / (size_t)nthreads, hi = nb * (size_t)(tid + 1) / (size_t)nthreads;
    for (size_t bi = lo; bi < hi; bi++) {
      const ResidualBlock& rb = E->blocks[bi];
      const int nres = rb.cost->num_residuals;
      const size_t np = rb.refs.size();
      params.resize(np);
      cols.resize(np);
      bool any_free = false;
      for (size_t i = 0; i < np; i++) {
        params[i] = E->block_ptr(s, rb.refs[i]);
        cols[i] = E->block_col(L, rb.refs[i]);
        if (cols[i] >= 0) any_free = true;
      }
      res.assign(nres, 0.0);
      const bool do_jac = want_jac && any_free;
      if (do_jac) {
        jac.resize(np);
        jptr.resize(np);
        for (size_t i = 0; i < np; i++) {
          if (cols[i] >= 0) {
            jac[i].assign((size_t)nres * rb.cost->block_sizes[i], 0.0);
            jptr[i] = jac[i].data();
          } else {
            jac[i].clear();
            jptr[i] = nullptr;  // constant block: Ceres passes NULL
          }
        }
        rb.cost->Evaluate(params.data(), res.data(), jptr.data());
      } else {
        jac.clear();
        rb.cost->Evaluate(params.data(), res.data(), nullptr);
      }
      double rho0;
      correct(rb, nres, res, jac, rb.cost->block_sizes, &rho0);
      if (any_free)
        cost += 0.5 * rho0;
      else
        fcost += 0.5 * rho0;  // Ceres removes residual blocks with only constant parameters: "fixed_cost"
      if (!do_jac) continue;
      for (size_t i = 0; i < np; i++) {
        if (cols[i] < 0) continue;
        const int si = rb.cost->block_sizes[i], li = local_size(si);
        for (size_t j = i; j < np; j++) {
          if (cols[j] < 0) continue;
          const int sj = rb.cost->block_sizes[j], lj = local_size(sj);
          for (int a = 0; a < li; a++)
            for (int c = 0; c < lj; c++) {
              double acc = 0;
              for (int r = 0; r < nres; r++) acc += jac[i][(size_t)r * si + a] * jac[j][(size_t)r * sj + c];
              H[(size_t)(cols[i] + a) * D + cols[j] + c] += acc;
              if (i != j) H[(size_t)(cols[j] + c) * D + cols[i] + a] += acc;
            }
        }
        for (int a = 0; a < li; a++) {
          double acc = 0;
          for (int r = 0; r < nres; r++) acc += jac[i][(size_t)r * si + a] * res[r];
          b[cols[i] + a] += acc;
        }
      }
    }
    costs[tid] = cost;
    fcosts[tid] = fcost;
  };
  if (nthreads == 1) {
    worker(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(worker, t);
    for (auto& t : th) t.join();
  }
  out.cost = 0;
  out.fixed_cost = 0;
  if (want_jac) {
    out.H.assign((size_t)D * D, 0.0);
    out.b.assign(D, 0.0);
  }
  for (int t = 0; t < nthreads; t++) {
    out.cost += costs[t];
    out.fixed_cost += fcosts[t];
    if (want_jac && !Hs[t].empty()) {
      for (size_t i = 0; i < (size_t)D * D; i++) out.H[i] += Hs[t][i];
      for (int i = 0; i < D; i++) out.b[i] += bs[t][i];
    }
  }
}

// ---- Plus (LieAnalyticLocalParameterization::Plus, ceres_local_param.h:132-139; ParameterBlock::Plus box projection,
// Ceres internal/ceres/parameter_block.h) over the canonical tangent vector ----
void plus(const clic_estimator* E, const clic_layout& L, const State& x, const double* delta, State& out) {
  out = x;
  for (int k = L.first_free_knot; k < E->K; k++) {
    int c = 6 * (k - L.first_free_knot);
    Quat q = load_q(&x.knot_q[(size_t)k * 4]);
    Quat r = so3_mul(q, so3_exp(v3(delta[c], delta[c + 1], delta[c + 2])));
    out.knot_q[(size_t)k * 4 + 0] = r.x;
    out.knot_q[(size_t)k * 4 + 1] = r.y;
    out.knot_q[(size_t)k * 4 + 2] = r.z;
    out.knot_q[(size_t)k * 4 + 3] = r.w;
    for (int a = 0; a < 3; a++) out.knot_p[(size_t)k * 3 + a] = x.knot_p[(size_t)k * 3 + a] + delta[c + 3 + a];
  }
  for (int j = 0; j < E->n_bias; j++) {
    if (L.col_bias_gyro >= 0)
      for (int a = 0; a < 3; a++) out.bias[(size_t)j * 6 + a] = x.bias[(size_t)j * 6 + a] + delta[L.col_bias_gyro + 3 * j + a];
    if (L.col_bias_accel >= 0)
      for (int a = 0; a < 3; a++)
        out.bias[(size_t)j * 6 + 3 + a] = x.bias[(size_t)j * 6 + 3 + a] + delta[L.col_bias_accel + 3 * j + a];
  }
  for (int s = 0; s < 3; s++) {
    if (L.col_t_offset[s] < 0) continue;
    double v = x.t_offset_ns[s] + delta[L.col_t_offset[s]];
    if (E->bounds_toff[s]) {
      v = std::max(v, E->toff_lo[s]);
      v = std::min(v, E->toff_hi[s]);
    }
    out.t_offset_ns[s] = v;
  }
  if (L.col_inv_depth >= 0) {
    int rank = 0;
    for (int i = 0; i < E->n_lm; i++)
      if (E->lm_free[i]) {
        out.inv_depth[i] = x.inv_depth[i] + delta[L.col_inv_depth + rank];
        rank++;
      }
  }
}

// Ambient-space quantities of the reduced program: |x|, |x - y|, |x - y|_inf over the columns marked active.
void ambient_diff(const clic_estimator* E, const clic_layout& L, const State& x, const State& y,
                  const std::vector<char>& active, double* xnorm, double* dnorm, double* dinf) {
  double xs = 0, ds = 0, dm = 0;
  auto acc = [&](const double* a, const double* b, int n) {
    for (int i = 0; i < n; i++) {
      xs += a[i] * a[i];
      double d = a[i] - b[i];
      ds += d * d;
      dm = std::max(dm, std::fabs(d));
    }
  };
  for (int k = L.first_free_knot; k < E->K; k++) {
    int c = 6 * (k - L.first_free_knot);
    if (active[c]) acc(&x.knot_q[(size_t)k * 4], &y.knot_q[(size_t)k * 4], 4);
    if (active[c + 3]) acc(&x.knot_p[(size_t)k * 3], &y.knot_p[(size_t)k * 3], 3);
  }
  for (int j = 0; j < E->n_bias; j++) {
    if (L.col_bias_gyro >= 0 && active[L.col_bias_gyro + 3 * j]) acc(&x.bias[(size_t)j * 6], &y.bias[(size_t)j * 6], 3);
    if (L.col_bias_accel >= 0 && active[L.col_bias_accel + 3 * j])
      acc(&x.bias[(size_t)j * 6 + 3], &y.bias[(size_t)j * 6 + 3], 3);
  }
  for (int s = 0; s < 3; s++)
    if (L.col_t_offset[s] >= 0 && active[L.col_t_offset[s]]) acc(&x.t_offset_ns[s], &y.t_offset_ns[s], 1);
  if (L.col_inv_depth >= 0) {
    int rank = 0;
    for (int i = 0; i < E->n_lm; i++)
      if (E->lm_free[i]) {
        if (active[L.col_inv_depth + rank]) acc(&x.inv_depth[i], &y.inv_depth[i], 1);
        rank++;
      }
  }
  if (xnorm) *xnorm = std::sqrt(xs);
  if (dnorm) *dnorm = std::sqrt(ds);
  if (dinf) *dinf = dm;
}

// Dense Cholesky solve A x = rhs (A symmetric positive definite, n x n row-major).  Stands in for Ceres'
// SPARSE_NORMAL_CHOLESKY (trajectory_estimator.cpp:994) — same solution up to rounding.
bool cholesky_solve(std::vector<double> A, int n, const double* rhs, double* x) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = v / d;
    }
  }
  std::vector<double> y(n);
  for (int i = 0; i < n; i++) {
    double v = rhs[i];
    for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * y[k];
    y[i] = v / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < n; k++) v -= A[(size_t)k * n + i] * x[k];
    x[i] = v / A[(size_t)i * n + i];
  }
  return true;
}

// Cyclic Jacobi eigen-decomposition of a symmetric matrix (stands in for Eigen::SelfAdjointEigenSolver,
// marginalization_factor.cpp:210,253).  A = V diag(w) V^T, V column-major eigenvectors in V[i*n+k] = V(i,k).
void jacobi_eigh(std::vector<double> A, int n, std::vector<double>& w, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; j++) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-40 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        double tau = (aqq - app) / (2.0 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; i++) w[i] = A[(size_t)i * n + i];
}

// ---------------- Ceres 1.14 trust-region LM (TrustRegionMinimizer::Minimize) ----------------
int solve(clic_estimator* E, int max_iterations, clic_summary* sum) {
  // Ceres defaults (solver.h, Ceres 1.14) — unverifiable in this container:
  const double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
  const double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const int max_num_consecutive_invalid_steps = 5;
  // projected line search for bound-constrained problems (TrustRegionMinimizer::DoLineSearch)
  const double ls_sufficient_decrease = 1e-4, ls_max_step_contraction = 1e-3, ls_min_step_contraction = 0.6,
               ls_min_step_size = 1e-9;
  const int ls_max_iterations = 20;

  clic_layout L = E->layout();
  const int D = L.D;
  *sum = clic_summary{};
  bool is_constrained = false;
  for (int s = 0; s < 3; s++)
    if (L.col_t_offset[s] >= 0 && E->bounds_toff[s]) is_constrained = true;

  State x = E->x;
  std::vector<double> zero(D, 0.0);
  if (is_constrained) {  // IterationZero: project the start point onto the feasible set
    State xp;
    plus(E, L, x, zero.data(), xp);
    x = xp;
  }
  EvalOut ev;
  evaluate(E, x, true, ev);
  sum->num_jacobian_evals++;
  double x_cost = ev.cost;
  sum->initial_cost = ev.cost + ev.fixed_cost;
  sum->fixed_cost = ev.fixed_cost;
  std::vector<char> active(D, 0);
  for (int i = 0; i < D; i++) active[i] = ev.H[(size_t)i * D + i] > 0.0;
  // Jacobi scaling, estimated once at iteration 0: 1 / (1 + sqrt(|J_col|^2))
  std::vector<double> scale(D);
  for (int i = 0; i < D; i++) scale[i] = 1.0 / (1.0 + std::sqrt(ev.H[(size_t)i * D + i]));
  std::vector<double> Hs((size_t)D * D), gs(D), g(D);
  auto rescale = [&]() {
    for (int i = 0; i < D; i++) {
      g[i] = ev.b[i];
      gs[i] = ev.b[i] * scale[i];
      for (int j = 0; j < D; j++) Hs[(size_t)i * D + j] = ev.H[(size_t)i * D + j] * scale[i] * scale[j];
    }
  };
  rescale();
  double x_norm, gmax;
  auto gradient_norm = [&]() {
    std::vector<double> ng(D);
    for (int i = 0; i < D; i++) ng[i] = -g[i];
    State xg;
    plus(E, L, x, ng.data(), xg);
    ambient_diff(E, L, x, xg, active, &x_norm, nullptr, &gmax);
  };
  gradient_norm();

  double radius = initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  std::vector<double> diagonal(D), step(D), delta(D), rhs(D), A((size_t)D * D);
  int iteration = 0, invalid = 0;
  int termination = CLIC_TERM_NO_CONVERGENCE, reason = CLIC_REASON_MAX_ITERATIONS;

  while (true) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (iteration >= max_iterations) { termination = CLIC_TERM_NO_CONVERGENCE; reason = CLIC_REASON_MAX_ITERATIONS; break; }
    if (gmax <= gradient_tolerance) { termination = CLIC_TERM_CONVERGENCE; reason = CLIC_REASON_GRADIENT_TOL; break; }
    if (radius <= min_trust_region_radius) { termination = CLIC_TERM_CONVERGENCE; reason = CLIC_REASON_MIN_RADIUS; break; }
    iteration++;

    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int i = 0; i < D; i++) diagonal[i] = std::min(std::max(Hs[(size_t)i * D + i], min_lm_diagonal), max_lm_diagonal);
    }
    for (int i = 0; i < D; i++) {
      for (int j = 0; j < D; j++) A[(size_t)i * D + j] = Hs[(size_t)i * D + j];
      double lm = std::sqrt(diagonal[i] / radius);
      A[(size_t)i * D + i] += lm * lm;
      rhs[i] = gs[i];
    }
    bool ok = cholesky_solve(A, D, rhs.data(), step.data());
    if (ok)
      for (int i = 0; i < D; i++)
        if (!std::isfinite(step[i])) ok = false;
    for (int i = 0; i < D; i++) step[i] = -step[i];
    reuse_diagonal = true;
    double model_cost_change = 0;
    if (ok) {
      // model_cost_change = -(J s)^T (f + J s / 2) = -(g~^T s + 1/2 s^T H~ s)
      double gts = 0, shs = 0;
      for (int i = 0; i < D; i++) {
        gts += gs[i] * step[i];
        double row = 0;
        for (int j = 0; j < D; j++) row += Hs[(size_t)i * D + j] * step[j];
        shs += step[i] * row;
      }
      model_cost_change = -(gts + 0.5 * shs);
    }
    bool step_is_valid = ok && (model_cost_change > 0.0);
    if (!step_is_valid) {
      // HandleInvalidStep
      if (++invalid >= max_num_consecutive_invalid_steps) { termination = CLIC_TERM_FAILURE; reason = CLIC_REASON_INVALID_STEPS; break; }
      radius = radius / decrease_factor;  // StepIsInvalid -> StepRejected
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      sum->num_unsuccessful_steps++;
      continue;
    }
    invalid = 0;
    for (int i = 0; i < D; i++) delta[i] = step[i] * scale[i];

    State cand;
    double cand_cost;
    EvalOut ec;
    if (is_constrained) {
      // TrustRegionMinimizer::DoLineSearch (Armijo, projected).  DEVIATION (documented in DESIGN.md): when the
      // first Armijo test fails Ceres interpolates with a CUBIC through value+gradient at the samples; we
      // use the QUADRATIC through f(0), f'(0), f(current), same contraction clamps.
      double g0 = 0, dmax = 0;
      for (int i = 0; i < D; i++) {
        g0 += g[i] * delta[i];
        dmax = std::max(dmax, std::fabs(delta[i]));
      }
      double cur = 1.0, fcur;
      std::vector<double> sd(D);
      auto f_at = [&](double a) {
        for (int i = 0; i < D; i++) sd[i] = a * delta[i];
        State xs;
        plus(E, L, x, sd.data(), xs);
        EvalOut e;
        evaluate(E, xs, false, e);
        sum->num_residual_evals++;
        return e.cost;
      };
      fcur = f_at(cur);
      int it = 0;
      bool success = true;
      while (!std::isfinite(fcur) || fcur > x_cost + ls_sufficient_decrease * g0 * cur) {
        if (++it >= ls_max_iterations) { success = false; break; }
        double denom = 2.0 * (fcur - x_cost - g0 * cur);
        double q = (denom > 0 && std::isfinite(denom)) ? (-g0 * cur * cur / denom) : ls_min_step_contraction * cur;
        q = std::min(std::max(q, ls_max_step_contraction * cur), ls_min_step_contraction * cur);
        if (q * dmax < ls_min_step_size) { success = false; break; }
        cur = q;
        fcur = f_at(cur);
      }
      if (success)
        for (int i = 0; i < D; i++) delta[i] *= cur;
    }
    plus(E, L, x, delta.data(), cand);
    evaluate(E, cand, false, ec);
    sum->num_residual_evals++;
    cand_cost = ec.cost;
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();

    // ParameterToleranceReached
    double step_norm;
    ambient_diff(E, L, x, cand, active, nullptr, &step_norm, nullptr);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
      termination = CLIC_TERM_CONVERGENCE; reason = CLIC_REASON_PARAMETER_TOL; break;
    }
    // FunctionToleranceReached
    double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= function_tolerance * x_cost) {
      termination = CLIC_TERM_CONVERGENCE; reason = CLIC_REASON_FUNCTION_TOL; break;
    }
    // IsStepSuccessful
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      // HandleSuccessfulStep
      x = cand;
      evaluate(E, x, true, ev);
      sum->num_jacobian_evals++;
      x_cost = ev.cost;
      rescale();
      gradient_norm();
      sum->num_successful_steps++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_trust_region_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      // HandleUnsuccessfulStep
      sum->num_unsuccessful_steps++;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
    }
  }
  E->x = x;
  sum->iterations = iteration;
  sum->termination = termination;
  sum->termination_reason = reason;
  sum->final_cost = x_cost + ev.fixed_cost;
  sum->final_radius = radius;
  sum->final_gradient_max_norm = gmax;
  return CLIC_OK;
}

// ---------------- marginalization (MarginalizationInfo) ----------------
// The assembled system [dropped | kept] of the last marginalize() call, before the Schur complement (test probe:
// tests compare the Schur routes — the reference's eigen pseudo-inverse, the product's Cholesky — with an extended-
// precision elimination of this system).
static std::vector<double> g_last_marg_A, g_last_marg_b;
static int g_last_marg_pos = 0, g_last_marg_m = 0;

int marginalize(clic_estimator* E, clic_prior** out) {
  // addResidualBlockInfo: sizes of all blocks, idx=0 marks dropped — marginalization_factor.cpp:76-92
  std::map<CanonKey, BlockRef> all;
  std::map<CanonKey, char> dropped;
  for (auto& rb : E->marg_blocks) {
    for (auto& r : rb.refs) all[canon_key(r)] = r;
    for (int d : rb.drop_set) dropped[canon_key(rb.refs[d])] = 1;
  }
  // marginalize(): dropped first (m), then the rest (n) — ":154-166"
  std::map<CanonKey, int> idx;
  int pos = 0;
  for (auto& kv : all)
    if (dropped.count(kv.first)) {
      idx[kv.first] = pos;
      pos += local_size(block_size(kv.second.kind));
    }
  int m = pos;
  for (auto& kv : all)
    if (!dropped.count(kv.first)) {
      idx[kv.first] = pos;
      pos += local_size(block_size(kv.second.kind));
    }
  int n = pos - m;
  if (n <= 0) return fail(CLIC_ERR_NOTHING_TO_KEEP, "marginalization leaves no parameters");

  // preMarginalize (evaluate every factor with all Jacobians, ":94-116") + ThreadsConstructA (":122-148")
  std::vector<double> A((size_t)pos * pos, 0.0), b(pos, 0.0);
  for (auto& rb : E->marg_blocks) {
    const int nres = rb.cost->num_residuals;
    const size_t np = rb.refs.size();
    std::vector<const double*> params(np);
    std::vector<std::vector<double>> jac(np);
    std::vector<double*> jptr(np);
    for (size_t i = 0; i < np; i++) {
      params[i] = E->block_ptr(E->x, rb.refs[i]);
      jac[i].assign((size_t)nres * rb.cost->block_sizes[i], 0.0);
      jptr[i] = jac[i].data();
    }
    std::vector<double> res(nres, 0.0);
    rb.cost->Evaluate(params.data(), res.data(), jptr.data());
    double rho0;
    correct(rb, nres, res, jac, rb.cost->block_sizes, &rho0);
    for (size_t i = 0; i < np; i++) {
      int ci = idx[canon_key(rb.refs[i])];
      const int si = rb.cost->block_sizes[i], li = local_size(si);
      for (size_t j = i; j < np; j++) {
        int cj = idx[canon_key(rb.refs[j])];
        const int sj = rb.cost->block_sizes[j], lj = local_size(sj);
        for (int a = 0; a < li; a++)
          for (int c = 0; c < lj; c++) {
            double acc = 0;
            for (int r = 0; r < nres; r++) acc += jac[i][(size_t)r * si + a] * jac[j][(size_t)r * sj + c];
            A[(size_t)(ci + a) * pos + cj + c] += acc;
            if (i != j) A[(size_t)(cj + c) * pos + ci + a] += acc;
          }
      }
      for (int a = 0; a < li; a++) {
        double acc = 0;
        for (int r = 0; r < nres; r++) acc += jac[i][(size_t)r * si + a] * res[r];
        b[ci + a] += acc;
      }
    }
  }
  g_last_marg_A = A;
  g_last_marg_b = b;
  g_last_marg_pos = pos;
  g_last_marg_m = m;
  const double eps = 1e-15;  // marginalization_factor.h:142
  // Amm^+ via eigen-decomposition with eigenvalue cut — ":208-213"
  std::vector<double> Amm((size_t)m * m), Amm_inv((size_t)m * m, 0.0);
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) Amm[(size_t)i * m + j] = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
  if (m > 0) {
    std::vector<double> w, V;
    jacobi_eigh(Amm, m, w, V);
    for (int k = 0; k < m; k++) {
      double inv = (w[k] > eps) ? 1.0 / w[k] : 0.0;
      if (inv == 0.0) continue;
      for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) Amm_inv[(size_t)i * m + j] += V[(size_t)i * m + k] * inv * V[(size_t)j * m + k];
    }
  }
  // Schur complement — ":231-237"
  std::vector<double> T((size_t)n * m, 0.0);  // Arm * Amm_inv
  for (int i = 0; i < n; i++)
    for (int k = 0; k < m; k++) {
      double a = A[(size_t)(m + i) * pos + k];
      if (a == 0.0) continue;
      for (int j = 0; j < m; j++) T[(size_t)i * m + j] += a * Amm_inv[(size_t)k * m + j];
    }
  std::vector<double> An((size_t)n * n), bn(n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int k = 0; k < m; k++) acc += T[(size_t)i * m + k] * A[(size_t)k * pos + m + j];
      An[(size_t)i * n + j] = A[(size_t)(m + i) * pos + m + j] - acc;
    }
    double acc = 0;
    for (int k = 0; k < m; k++) acc += T[(size_t)i * m + k] * b[k];
    bn[i] = b[m + i] - acc;
  }
  // A = V S V^T, J0 = S^{1/2} V^T, r0 = S^{-1/2} V^T b — ":246-264"
  std::vector<double> w, V;
  jacobi_eigh(An, n, w, V);
  std::unique_ptr<clic_prior> P(new clic_prior);
  P->n = n;
  P->m = m;
  P->J0.assign((size_t)n * n, 0.0);
  P->r0.assign(n, 0.0);
  for (int k = 0; k < n; k++) {
    double S = (w[k] > eps) ? w[k] : 0.0;
    double S_inv = (w[k] > eps) ? 1.0 / w[k] : 0.0;
    double S_sqrt = std::sqrt(S), S_inv_sqrt = std::sqrt(S_inv);
    double vtb = 0;
    for (int i = 0; i < n; i++) {
      P->J0[(size_t)k * n + i] = S_sqrt * V[(size_t)i * n + k];
      vtb += V[(size_t)i * n + k] * bn[i];
    }
    P->r0[k] = S_inv_sqrt * vtb;
  }
  // getParameterBlocks — ":301-317"
  for (auto& kv : all) {
    if (dropped.count(kv.first)) continue;
    const BlockRef& r = kv.second;
    clic_prior::Block bl;
    bl.kind = r.kind;
    bl.size = block_size(r.kind);
    bl.idx = idx[kv.first] - m;
    bl.id = r.id;
    if (r.kind == CLIC_BLOCK_KNOT_ROT || r.kind == CLIC_BLOCK_KNOT_POS) bl.id += E->knot_id0;
    if (r.kind == CLIC_BLOCK_BIAS_GYRO || r.kind == CLIC_BLOCK_BIAS_ACCEL) bl.id += E->bias_id0;
    const double* d = E->block_ptr(E->x, r);
    for (int k = 0; k < 4; k++) bl.x0[k] = k < bl.size ? d[k] : 0.0;
    P->blocks.push_back(bl);
  }
  *out = P.release();
  return CLIC_OK;
}

// TrajectoryEstimator::PrepareMarginalizationInfo (spline overload) — trajectory_estimator.cpp:208-232
void prepare_marg_spline(clic_estimator* E, ResidualBlock&& rb, const SplineMeta& meta,
                         const std::vector<int>& drop_wo_ctrl) {
  std::vector<int> drop_set = drop_wo_ctrl;
  if (E->opt.ctrl_to_be_opt_later > E->opt.ctrl_to_be_opt_now) {
    std::vector<int> ctrl_id;
    E->GetCtrlIdxs(meta, ctrl_id);
    for (int i = 0; i < (int)ctrl_id.size(); ++i) {
      if (ctrl_id[i] + E->knot_id0 < E->opt.ctrl_to_be_opt_later) {
        drop_set.emplace_back(i);
        drop_set.emplace_back(i + (int)meta.NumParameters());
      }
    }
  }
  if (drop_set.size() > 0) {
    rb.drop_set = drop_set;
    E->marg_blocks.push_back(std::move(rb));
  }
}

bool valid(const clic_estimator* h) { return h != nullptr; }

}  // namespace

// =============================== C ABI ===============================
extern "C" {

CLIC_API const char* clic_last_error(void) { return g_last_error.c_str(); }
CLIC_API const char* clic_backend(void) { return "oracle-cpu"; }

CLIC_API void clic_default_options(clic_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->lock_traj = 0;
  o->lock_ab = o->lock_wb = o->lock_g = 1;
  for (int s = 0; s < 3; s++) o->lock_t_offset[s] = 1;
  o->t_offset_padding_ns = (int64_t)1e7;
  o->is_marg_state = 0;
  o->ctrl_to_be_opt_now = o->ctrl_to_be_opt_later = 0;
  o->marg_bias_param = o->marg_gravity_param = o->marg_t_offset_param = 1;
  o->device = 0;
  o->stream = nullptr;
  o->deterministic = 1;
  o->image_sqrt_info = 450.0 / 1.5;  // ImageFeatureFactor's class default (image_feature_factor.h:277-278)
}

// Test-only knob: number of OpenMP threads of the evaluation loop (the reference runs Ceres with 1 thread,
// trajectory_estimator.cpp:1004-1007, and marginalization with 4 pthreads, marginalization_factor.h:29).
// ResidualSummary-like read-back (trajectory_estimator.h:42-98): 1/2 rho(|r|^2) summed per ResidualType
// (marginalization_factor.h:31-46; slot 14 = the pose / local-velocity / relative-rotation / image-depth adders) over
// the solve problem at the current state, and the number of residual blocks of each type.
CLIC_API int clic_residual_summary(clic_estimator* E, double cost[16], int64_t count[16]) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  for (int i = 0; i < 16; i++) { if (cost) cost[i] = 0.0; if (count) count[i] = 0; }
  for (auto& rb : E->blocks) {
    const int nres = rb.cost->num_residuals;
    const size_t np = rb.refs.size();
    std::vector<const double*> params(np);
    for (size_t i = 0; i < np; i++) params[i] = E->block_ptr(E->x, rb.refs[i]);
    std::vector<double> res(nres, 0.0);
    rb.cost->Evaluate(params.data(), res.data(), nullptr);
    std::vector<std::vector<double>> nojac;
    double rho0;
    correct(rb, nres, res, nojac, rb.cost->block_sizes, &rho0);
    const int t = rb.rtype >= 0 && rb.rtype < 16 ? rb.rtype : 15;
    if (cost) cost[t] += 0.5 * rho0;
    if (count) count[t] += 1;
  }
  return CLIC_OK;
}

CLIC_API int clic_oracle_last_marg_system(double* A, double* b, int32_t* pos, int32_t* m) {
  if (pos) *pos = g_last_marg_pos;
  if (m) *m = g_last_marg_m;
  if (A) std::memcpy(A, g_last_marg_A.data(), g_last_marg_A.size() * sizeof(double));
  if (b) std::memcpy(b, g_last_marg_b.data(), g_last_marg_b.size() * sizeof(double));
  return CLIC_OK;
}

CLIC_API int clic_oracle_set_threads(int n) {
  g_threads = n < 1 ? 1 : n;
  return CLIC_OK;
}

// ---- test-only probes (not part of include/clic_cuda.h) ----
CLIC_API int64_t clic_oracle_num_blocks(clic_estimator* E, int32_t marg) {
  return valid(E) ? (int64_t)(marg ? E->marg_blocks.size() : E->blocks.size()) : -1;
}
// Dense local Jacobian (nres x D, canonical columns) and residual of one residual block of the solve problem,
// before (apply_loss = 0) or after (1) the loss correction.  What Problem::Evaluate -> CRS Jacobian gives the
// reference's test (src/app/test_analytic_factor.cpp:1051-1082).
CLIC_API int clic_oracle_block_jacobian(clic_estimator* E, int64_t block, int32_t apply_loss, int32_t* nres_out,
                                        double* res_out, double* J_out) {
  if (!valid(E) || block < 0 || block >= (int64_t)E->blocks.size()) return fail(CLIC_ERR_BAD_ARGUMENT, "block");
  clic_layout L = E->layout();
  const ResidualBlock& rb = E->blocks[block];
  const int nres = rb.cost->num_residuals;
  const size_t np = rb.refs.size();
  std::vector<const double*> params(np);
  std::vector<std::vector<double>> jac(np);
  std::vector<double*> jptr(np);
  for (size_t i = 0; i < np; i++) {
    params[i] = E->block_ptr(E->x, rb.refs[i]);
    jac[i].assign((size_t)nres * rb.cost->block_sizes[i], 0.0);
    jptr[i] = jac[i].data();
  }
  std::vector<double> res(nres, 0.0);
  rb.cost->Evaluate(params.data(), res.data(), jptr.data());
  if (apply_loss) {
    double rho0;
    correct(rb, nres, res, jac, rb.cost->block_sizes, &rho0);
  }
  if (nres_out) *nres_out = nres;
  if (res_out) std::copy(res.begin(), res.end(), res_out);
  if (J_out) {
    std::fill(J_out, J_out + (size_t)nres * L.D, 0.0);
    for (size_t i = 0; i < np; i++) {
      int c = E->block_col(L, rb.refs[i]);
      if (c < 0) continue;
      int sz = rb.cost->block_sizes[i], ls = local_size(sz);
      for (int r = 0; r < nres; r++)
        for (int a = 0; a < ls; a++) J_out[(size_t)r * L.D + c + a] += jac[i][(size_t)r * sz + a];
    }
  }
  return CLIC_OK;
}

CLIC_API int clic_create(const clic_window_desc* w, const clic_options* opt, clic_estimator** out) {
  if (!w || !opt || !out) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (w->n_knots < N || w->dt_ns <= 0) return fail(CLIC_ERR_BAD_ARGUMENT, "need >= 4 knots and dt > 0");
  auto* E = new clic_estimator;
  E->t0_ns = w->t0_ns;
  E->dt_ns = w->dt_ns;
  E->K = w->n_knots;
  E->knot_id0 = w->knot_id0;
  E->n_bias = w->n_bias;
  E->bias_id0 = w->bias_id0;
  E->n_lm = w->n_landmarks;
  E->S_CtoI = Quat{w->S_CtoI[0], w->S_CtoI[1], w->S_CtoI[2], w->S_CtoI[3]};
  E->p_CinI = v3(w->p_CinI[0], w->p_CinI[1], w->p_CinI[2]);
  E->x.knot_q.assign(w->knot_q, w->knot_q + (size_t)4 * E->K);
  E->x.knot_p.assign(w->knot_p, w->knot_p + (size_t)3 * E->K);
  if (E->n_bias > 0) E->x.bias.assign(w->bias, w->bias + (size_t)6 * E->n_bias);
  if (E->n_lm > 0) E->x.inv_depth.assign(w->inv_depth, w->inv_depth + E->n_lm);
  for (int s = 0; s < 3; s++) E->x.t_offset_ns[s] = w->t_offset_ns[s];
  for (int a = 0; a < 3; a++) E->x.gravity[a] = w->gravity[a];
  E->opt = *opt;
  E->lm_free.assign(std::max(0, E->n_lm), 0);
  *out = E;
  return CLIC_OK;
}

CLIC_API int clic_destroy(clic_estimator* h) {
  delete h;
  return CLIC_OK;
}

CLIC_API int clic_set_fixed_index(clic_estimator* h, int32_t idx) {
  if (!valid(h)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  h->fixed_index_global = idx;
  return CLIC_OK;
}

CLIC_API int clic_add_loam(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                           const int32_t* geo_type, const double* geo_plane, const double* geo_normal,
                           const double* geo_point, const double S_GtoM[4], const double p_GinM[3],
                           const double S_LtoI[4], const double p_LinI[3], double weight, int32_t marg_this_factor,
                           int64_t* n_dropped) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  int64_t dropped = 0;
  Quat SG = load_q(S_GtoM), SL = load_q(S_LtoI);
  Vec3 pG = load_v(p_GinM), pL = load_v(p_LinI);
  for (int64_t i = 0; i < n; i++) {
    PointCorrespondence pc{};
    pc.t_point = t_point[i];
    pc.t_map = 0;
    pc.point = load_v(point + 3 * i);
    pc.geo_type = geo_type[i];
    if (pc.geo_type == CLIC_GEO_PLANE) {
      if (!geo_plane) return fail(CLIC_ERR_BAD_ARGUMENT, "plane record without geo_plane");
      for (int k = 0; k < 4; k++) pc.geo_plane[k] = geo_plane[4 * i + k];
    } else {
      if (!geo_normal || !geo_point) return fail(CLIC_ERR_BAD_ARGUMENT, "line record without geo_normal/geo_point");
      pc.geo_normal = load_v(geo_normal + 3 * i);
      pc.geo_point = load_v(geo_point + 3 * i);
    }
    int64_t time_ns;
    if (!E->MeasuredTimeToNs(CLIC_SENSOR_LIDAR, pc.t_point, time_ns)) {  // trajectory_estimator.cpp:717
      dropped++;
      E->probe[0].push_back({time_ns, 1});
      continue;
    }
    E->probe[0].push_back({time_ns, 0});
    SplineMeta meta;
    E->CaculateSplineMeta({{time_ns, time_ns}}, meta);
    ResidualBlock rb;
    rb.cost.reset(new LoamFeatureFactor(time_ns, pc, meta.segments.at(0), SG, pG, SL, pL, weight));
    E->AddControlPoints(meta, rb.refs, false);
    E->AddControlPoints(meta, rb.refs, true);
    rb.rtype = 4;
    if (E->opt.is_marg_state && marg_this_factor) {
      // trajectory_estimator.cpp:729-741: gate on |r| / w < 0.05
      std::vector<const double*> params(rb.refs.size());
      for (size_t k = 0; k < rb.refs.size(); k++) params[k] = E->block_ptr(E->x, rb.refs[k]);
      double r;
      rb.cost->Evaluate(params.data(), &r, nullptr);
      double dist = std::sqrt((r / weight) * (r / weight));
      if (dist < 0.05) prepare_marg_spline(E, std::move(rb), meta, {});
    } else {
      E->blocks.push_back(std::move(rb));
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return CLIC_OK;
}

CLIC_API int clic_add_loam_packed(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                  const uint8_t* geo_type, const double* geo, const double S_GtoM[4],
                                  const double p_GinM[3], const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped) {
  std::vector<int32_t> type(n);
  std::vector<double> plane(4 * (size_t)n), normal(3 * (size_t)n), gpoint(3 * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    type[i] = geo_type[i];
    for (int k = 0; k < 4; k++) plane[4 * i + k] = geo[6 * i + k];
    for (int k = 0; k < 3; k++) {
      gpoint[3 * i + k] = geo[6 * i + k];
      normal[3 * i + k] = geo[6 * i + 3 + k];
    }
  }
  return clic_add_loam(E, n, t_point, point, type.data(), plane.data(), normal.data(), gpoint.data(), S_GtoM, p_GinM,
                       S_LtoI, p_LinI, weight, marg_this_factor, n_dropped);
}

CLIC_API int clic_add_loam_planes(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                  const double* plane, const double S_GtoM[4], const double p_GinM[3],
                                  const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped) {
  std::vector<int32_t> type((size_t)std::max<int64_t>(n, 0), 1);  // GeometryType::Plane
  return clic_add_loam(E, n, t_point, point, type.data(), plane, nullptr, nullptr, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                       marg_this_factor, n_dropped);
}
CLIC_API int clic_add_loam_lines(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                 const double* line, const double S_GtoM[4], const double p_GinM[3],
                                 const double S_LtoI[4], const double p_LinI[3], double weight,
                                 int32_t marg_this_factor, int64_t* n_dropped) {
  const size_t m = (size_t)std::max<int64_t>(n, 0);
  std::vector<int32_t> type(m, 0);  // GeometryType::Line
  std::vector<double> gpoint(3 * m), normal(3 * m);
  for (size_t i = 0; i < m; i++)
    for (int k = 0; k < 3; k++) {
      gpoint[3 * i + k] = line[6 * i + k];
      normal[3 * i + k] = line[6 * i + 3 + k];
    }
  return clic_add_loam(E, n, t_point, point, type.data(), nullptr, normal.data(), gpoint.data(), S_GtoM, p_GinM, S_LtoI,
                       p_LinI, weight, marg_this_factor, n_dropped);
}

CLIC_API int clic_add_imu(clic_estimator* E, int64_t n, const double* timestamp, const double* gyro,
                          const double* accel, int32_t bias_slot, const double info_vec[6],
                          int32_t marg_this_factor, int64_t* n_dropped) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (bias_slot < 0 || bias_slot >= E->n_bias) return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  int64_t dropped = 0;
  const int S = CLIC_SENSOR_IMU;
  for (int64_t i = 0; i < n; i++) {
    IMUData d{timestamp[i], load_v(gyro + 3 * i), load_v(accel + 3 * i)};
    int64_t time_ns;
    if (!E->MeasuredTimeToNs(S, d.timestamp, time_ns)) {
      dropped++;
      E->probe[1].push_back({time_ns, 1});
      continue;
    }
    E->probe[1].push_back({time_ns, 0});
    int64_t t_corrected_ns = (int64_t)(time_ns + E->x.t_offset_ns[S]);  // "time_ns + (*t_offset_ns)", ":379"
    SplineMeta meta;
    if (E->opt.lock_t_offset[S])
      E->CaculateSplineMeta({{t_corrected_ns, t_corrected_ns}}, meta);
    else
      E->CaculateSplineMeta(
          {{t_corrected_ns - E->opt.t_offset_padding_ns, t_corrected_ns + E->opt.t_offset_padding_ns}}, meta);
    ResidualBlock rb;
    rb.cost.reset(new IMUFactor(time_ns, d, meta.segments.at(0), info_vec));
    E->AddControlPoints(meta, rb.refs, false);
    E->AddControlPoints(meta, rb.refs, true);
    rb.refs.push_back({CLIC_BLOCK_BIAS_GYRO, bias_slot});
    rb.refs.push_back({CLIC_BLOCK_BIAS_ACCEL, bias_slot});
    rb.refs.push_back({CLIC_BLOCK_GRAVITY, 0});
    rb.refs.push_back({CLIC_BLOCK_T_OFFSET, S});
    if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {  // ":419-424" bounds around the value at add time
      double t_ns = (double)E->opt.t_offset_padding_ns;
      E->bounds_toff[S] = true;
      E->toff_lo[S] = E->x.t_offset_ns[S] - t_ns;
      E->toff_hi[S] = E->x.t_offset_ns[S] + t_ns;
    }
    rb.has_loss = true;
    rb.loss_a = marg_this_factor ? 3 : 10;  // ":426-430"
    rb.rtype = 1;
    if (E->opt.is_marg_state && marg_this_factor) {
      std::vector<int> drop;
      int Knot_size = 2 * (int)meta.NumParameters();
      if (E->opt.marg_bias_param) {
        drop.push_back(Knot_size);
        drop.push_back(Knot_size + 1);
      }
      if (E->opt.marg_gravity_param) drop.push_back(Knot_size + 2);
      if (E->opt.marg_t_offset_param) drop.push_back(Knot_size + 3);
      prepare_marg_spline(E, std::move(rb), meta, drop);
    } else {
      E->blocks.push_back(std::move(rb));
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return CLIC_OK;
}

// AddGlobalVelocityMeasurement — trajectory_estimator.cpp:609-664
CLIC_API int clic_add_global_velocity(clic_estimator* E, double timestamp, const double global_v[3], double vel_weight,
                                      int32_t* added) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!global_v) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (added) *added = 0;
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "global velocity factor in a marginalization estimator");
  const int S = CLIC_SENSOR_IMU;
  int64_t time_ns;
  if (!E->MeasuredTimeToNs(S, timestamp, time_ns)) return CLIC_OK;  // "return false"
  const int64_t t_corrected_ns = (int64_t)(time_ns + E->x.t_offset_ns[S]);  // ":615"
  SplineMeta meta;
  if (E->opt.lock_t_offset[S])
    E->CaculateSplineMeta({{t_corrected_ns, t_corrected_ns}}, meta);
  else
    E->CaculateSplineMeta(
        {{t_corrected_ns - E->opt.t_offset_padding_ns, t_corrected_ns + E->opt.t_offset_padding_ns}}, meta);
  ResidualBlock rb;
  rb.cost.reset(new GlobalVelocityFactor(time_ns, load_v(global_v), meta, vel_weight));
  E->AddControlPoints(meta, rb.refs, true);  // position knots only (":645")
  rb.refs.push_back({CLIC_BLOCK_T_OFFSET, S});
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {  // ":653-656"
    const double t_ns = (double)E->opt.t_offset_padding_ns;
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->x.t_offset_ns[S] - t_ns;
    E->toff_hi[S] = E->x.t_offset_ns[S] + t_ns;
  }
  rb.has_loss = true;
  rb.loss_a = 60;  // CauchyLoss(60), ":642"
  rb.rtype = 8;  // RType_GlobalVelocity
  E->blocks.push_back(std::move(rb));
  if (added) *added = 1;
  return CLIC_OK;
}

CLIC_API int clic_add_bias(clic_estimator* E, int32_t slot_i, int32_t slot_j, double dt, const double info_vec[6],
                           int32_t marg_this_factor, int32_t marg_all_bias) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (slot_i < 0 || slot_i >= E->n_bias || slot_j < 0 || slot_j >= E->n_bias)
    return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  ResidualBlock rb;
  rb.cost.reset(new BiasFactor(dt, info_vec));
  rb.refs = {{CLIC_BLOCK_BIAS_GYRO, slot_i}, {CLIC_BLOCK_BIAS_GYRO, slot_j}, {CLIC_BLOCK_BIAS_ACCEL, slot_i},
             {CLIC_BLOCK_BIAS_ACCEL, slot_j}};
  rb.rtype = 2;
  if (E->opt.is_marg_state && marg_this_factor) {
    std::vector<int> drop = {0, 2};  // ":483-489"
    if (!E->opt.marg_bias_param) drop.clear();
    if (E->opt.marg_bias_param && marg_all_bias) drop = {0, 1, 2, 3};
    // non-spline PrepareMarginalizationInfo: always added, loss discarded (":199-206")
    rb.drop_set = drop;
    E->marg_blocks.push_back(std::move(rb));
  } else {
    E->blocks.push_back(std::move(rb));
  }
  return CLIC_OK;
}

// TrajectoryEstimator::AddGravityFactor — trajectory_estimator.cpp:500-523
CLIC_API int clic_add_gravity(clic_estimator* E, const double gravity[3], const double info_vec[3],
                              int32_t marg_this_factor) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  ResidualBlock rb;
  rb.cost.reset(new GravityFactor(gravity, info_vec));
  rb.refs = {{CLIC_BLOCK_GRAVITY, 0}};
  rb.rtype = 3;  // RType_Gravity
  if (E->opt.is_marg_state && marg_this_factor) {
    std::vector<int> drop = {0};  // ":515-519"
    if (!E->opt.marg_gravity_param) drop.clear();
    rb.drop_set = drop;  // non-spline PrepareMarginalizationInfo: always added (":199-206")
    E->marg_blocks.push_back(std::move(rb));
  } else {
    E->blocks.push_back(std::move(rb));
  }
  return CLIC_OK;
}

// One measurement time at the IMU offset -> the (padded) spline meta of the pose / local-velocity adders
// (trajectory_estimator.cpp:330-344, 671-684) and the time-offset bounds (":355-361, :696-702")
static void imu_time_meta(clic_estimator* E, int64_t time_ns, SplineMeta& meta) {
  const int S = CLIC_SENSOR_IMU;
  int64_t t_corrected_ns = (int64_t)(time_ns + E->x.t_offset_ns[S]);
  if (E->opt.lock_t_offset[S])
    E->CaculateSplineMeta({{t_corrected_ns, t_corrected_ns}}, meta);
  else
    E->CaculateSplineMeta(
        {{t_corrected_ns - E->opt.t_offset_padding_ns, t_corrected_ns + E->opt.t_offset_padding_ns}}, meta);
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {
    double t_ns = (double)E->opt.t_offset_padding_ns;
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->x.t_offset_ns[S] - t_ns;
    E->toff_hi[S] = E->x.t_offset_ns[S] + t_ns;
  }
}

// AddPoseMeasurementAnalytic — trajectory_estimator.cpp:326-368 (IMUPoseFactor, no loss)
CLIC_API int clic_add_pose(clic_estimator* E, int64_t n, const double* timestamp, const double* position,
                           const double* orientation, const double info_vec[6], int64_t* n_dropped) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "pose factor in a marginalization estimator");
  int64_t dropped = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t time_ns;
    if (!E->MeasuredTimeToNs(CLIC_SENSOR_IMU, timestamp[i], time_ns)) { dropped++; continue; }
    SplineMeta meta;
    imu_time_meta(E, time_ns, meta);
    PoseData pd{timestamp[i], load_v(position + 3 * i), load_q(orientation + 4 * i)};
    ResidualBlock rb;
    rb.cost.reset(new IMUPoseFactor(time_ns, pd, meta.segments.at(0), info_vec));
    E->AddControlPoints(meta, rb.refs, false);
    E->AddControlPoints(meta, rb.refs, true);
    rb.refs.push_back({CLIC_BLOCK_T_OFFSET, CLIC_SENSOR_IMU});
    rb.rtype = 14;
    E->blocks.push_back(std::move(rb));
  }
  if (n_dropped) *n_dropped = dropped;
  return CLIC_OK;
}

// AddLocalVelocityMeasurementAnalytic — trajectory_estimator.cpp:666-708 (LocalVelocityFactor, no loss)
CLIC_API int clic_add_local_velocity(clic_estimator* E, int64_t n, const double* timestamp, const double* local_v,
                                     double weight, int64_t* n_dropped) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "local velocity factor in a marginalization estimator");
  int64_t dropped = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t time_ns;
    if (!E->MeasuredTimeToNs(CLIC_SENSOR_IMU, timestamp[i], time_ns)) { dropped++; continue; }
    SplineMeta meta;
    imu_time_meta(E, time_ns, meta);
    ResidualBlock rb;
    rb.cost.reset(new LocalVelocityFactor(time_ns, load_v(local_v + 3 * i), meta.segments.at(0), weight));
    E->AddControlPoints(meta, rb.refs, false);
    E->AddControlPoints(meta, rb.refs, true);
    rb.refs.push_back({CLIC_BLOCK_T_OFFSET, CLIC_SENSOR_IMU});
    rb.rtype = 14;
    E->blocks.push_back(std::move(rb));
  }
  if (n_dropped) *n_dropped = dropped;
  return CLIC_OK;
}

// AddRelativeRotationAnalytic — trajectory_estimator.cpp:586-607 (RelativeOrientationFactor, no loss)
CLIC_API int clic_add_relative_rotation(clic_estimator* E, double ta, double tb, const double S_BtoA[4],
                                        const double info_vec[3], int32_t* added) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "relative rotation factor in a marginalization estimator");
  if (added) *added = 0;
  int64_t time_a_ns, time_b_ns;
  if (!E->MeasuredTimeToNs(CLIC_SENSOR_IMU, ta, time_a_ns)) return CLIC_OK;
  if (!E->MeasuredTimeToNs(CLIC_SENSOR_IMU, tb, time_b_ns)) return CLIC_OK;
  int64_t t_min = std::min(time_a_ns, time_b_ns), t_max = std::max(time_a_ns, time_b_ns);
  SplineMeta meta;
  E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
  ResidualBlock rb;
  rb.cost.reset(new RelativeOrientationFactor(load_q(S_BtoA), time_a_ns, time_b_ns, meta, info_vec));
  E->AddControlPoints(meta, rb.refs, false);
  rb.rtype = 14;
  E->blocks.push_back(std::move(rb));
  if (added) *added = 1;
  return CLIC_OK;
}

// AddImageDepthAnalytic — trajectory_estimator.cpp:830-852 (ImageDepthFactor, CauchyLoss(1); the inverse depth is a
// free parameter of the solve)
CLIC_API int clic_add_image_depth(clic_estimator* E, const double p_i[3], const double p_j[3], const double S_CitoCj[4],
                                  const double p_CiinCj[3], int32_t landmark) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "image depth factor in a marginalization estimator");
  if (landmark < 0 || landmark >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "landmark index out of range");
  ResidualBlock rb;
  rb.cost.reset(new ImageDepthFactor(load_v(p_i), load_v(p_j), load_q(S_CitoCj), load_v(p_CiinCj),
                                     E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450. / 1.5));
  rb.refs.push_back({CLIC_BLOCK_INV_DEPTH, landmark});
  E->lm_free[landmark] = 1;
  rb.has_loss = true;
  rb.loss_a = 1.0;
  rb.rtype = 14;
  E->blocks.push_back(std::move(rb));
  return CLIC_OK;
}

namespace {
PreIntegrationData load_preintegration(const clic_preintegration* p) {
  PreIntegrationData d;
  d.sum_dt = p->sum_dt;
  d.delta_p = load_v(p->delta_p);
  d.delta_q = load_q(p->delta_q);
  d.delta_v = load_v(p->delta_v);
  d.linearized_ba = load_v(p->linearized_ba);
  d.linearized_bg = load_v(p->linearized_bg);
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) {
      d.jacobian[r][c] = p->jacobian[r * 15 + c];
      d.sqrt_info[r][c] = p->sqrt_info[r * 15 + c];
    }
  d.gravity_norm = p->gravity_norm;
  return d;
}
PointCorrespondence load_correspondence(const clic_factor_desc* f) {
  PointCorrespondence pc;
  pc.t_point = pc.t_map = 0;
  pc.point = load_v(f->d);
  pc.geo_type = f->geo_type;
  for (int i = 0; i < 4; i++) pc.geo_plane[i] = f->d[3 + i];
  pc.geo_normal = load_v(f->d + 7);
  pc.geo_point = load_v(f->d + 10);
  return pc;
}
}  // namespace

// AddPreIntegrationAnalytic — trajectory_estimator.cpp:529-586 (PreIntegrationFactor, no loss); see clic_cuda.h for the
// time arguments
CLIC_API int clic_add_preintegration(clic_estimator* E, int64_t ta_ns, int64_t tb_ns, const clic_preintegration* pre,
                                     int32_t slot_i, int32_t slot_j, int32_t* added) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!pre) return fail(CLIC_ERR_BAD_ARGUMENT, "null pre-integration");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "pre-integration factor in a marginalization estimator");
  if (slot_i < 0 || slot_j < 0 || slot_i >= E->n_bias || slot_j >= E->n_bias)
    return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  if (added) *added = 0;
  if (ta_ns < E->minTimeNs() || ta_ns > E->maxTimeNs() || tb_ns < E->minTimeNs() || tb_ns > E->maxTimeNs()) return CLIC_OK;
  int64_t t_min = std::min(ta_ns, tb_ns), t_max = std::max(ta_ns, tb_ns);
  SplineMeta meta;
  E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
  ResidualBlock rb;
  rb.cost.reset(new PreIntegrationFactor(load_preintegration(pre), ta_ns, tb_ns, meta));
  E->AddControlPoints(meta, rb.refs, false);
  E->AddControlPoints(meta, rb.refs, true);
  rb.refs.push_back({CLIC_BLOCK_BIAS_GYRO, slot_i});
  rb.refs.push_back({CLIC_BLOCK_BIAS_GYRO, slot_j});
  rb.refs.push_back({CLIC_BLOCK_BIAS_ACCEL, slot_i});
  rb.refs.push_back({CLIC_BLOCK_BIAS_ACCEL, slot_j});
  rb.rtype = 14;  // (RType_PreIntegration = 7 in the reference; both libraries report the small adders together in slot 14)
  E->blocks.push_back(std::move(rb));
  if (added) *added = 1;
  return CLIC_OK;
}

// One analytic factor against the current window: cost_function->Evaluate with the parameter blocks laid out like the
// reference's test does (src/app/test_analytic_factor.cpp: CaculateSplineMeta over the measurement time(s),
// AddControlPoints for rotation then position knots, then the factor's own blocks).
CLIC_API int clic_evaluate_factor(clic_estimator* E, const clic_factor_desc* f, int32_t* n_res, double* residual,
                                  int32_t* n_cols, double* jac, int64_t jac_capacity) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!f || !n_res || !residual || !n_cols) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  const double* d = f->d;
  double x[16];
  for (int i = 0; i < 16; i++) x[i] = f->x[i];
  std::unique_ptr<CostFunction> cost;
  std::vector<BlockRef> refs;
  std::vector<std::pair<double*, int>> extra;  // (pointer, size) of the blocks after the knots
  auto in_window = [&](int64_t t) { return t >= E->minTimeNs() && t <= E->maxTimeNs(); };
  auto knots_of = [&](const SplineMeta& meta) {
    E->AddControlPoints(meta, refs, false);
    E->AddControlPoints(meta, refs, true);
  };
  SplineMeta meta;
  const double sqrt_info_default = E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450. / 1.5;
  (void)sqrt_info_default;
  switch (f->kind) {
    case CLIC_FACTOR_POSE: {
      const int64_t tc = f->t_a_ns + (int64_t)x[0];
      if (!in_window(tc)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{tc, tc}}, meta);
      PoseData pose;
      pose.timestamp = 0;
      pose.position = load_v(d);
      pose.orientation = load_q(d + 3);
      cost.reset(new IMUPoseFactor(f->t_a_ns, pose, meta.segments.at(0), d + 7));
      knots_of(meta);
      extra.push_back({x, 1});
      break;
    }
    case CLIC_FACTOR_LOCAL_VELOCITY: {
      const int64_t tc = f->t_a_ns + (int64_t)x[0];
      if (!in_window(tc)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{tc, tc}}, meta);
      cost.reset(new LocalVelocityFactor(f->t_a_ns, load_v(d), meta.segments.at(0), d[3]));
      knots_of(meta);
      extra.push_back({x, 1});
      break;
    }
    case CLIC_FACTOR_RELATIVE_ROTATION: {
      if (!in_window(f->t_a_ns) || !in_window(f->t_b_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      const int64_t t_min = std::min(f->t_a_ns, f->t_b_ns), t_max = std::max(f->t_a_ns, f->t_b_ns);
      E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
      cost.reset(new RelativeOrientationFactor(load_q(d), f->t_a_ns, f->t_b_ns, meta, d + 4));
      E->AddControlPoints(meta, refs, false);
      break;
    }
    case CLIC_FACTOR_IMAGE_DEPTH:
      cost.reset(new ImageDepthFactor(load_v(d), load_v(d + 3), load_q(d + 6), load_v(d + 10), d[13]));
      extra.push_back({x, 1});
      break;
    case CLIC_FACTOR_PREINTEGRATION: {
      if (!f->pre) return fail(CLIC_ERR_BAD_ARGUMENT, "null pre-integration");
      if (!in_window(f->t_a_ns) || !in_window(f->t_b_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      const int64_t t_min = std::min(f->t_a_ns, f->t_b_ns), t_max = std::max(f->t_a_ns, f->t_b_ns);
      E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
      cost.reset(new PreIntegrationFactor(load_preintegration(f->pre), f->t_a_ns, f->t_b_ns, meta));
      knots_of(meta);
      for (int i = 0; i < 4; i++) extra.push_back({x + 3 * i, 3});
      break;
    }
    case CLIC_FACTOR_IMAGE_3D2D: {
      const int64_t tc = f->t_a_ns + (int64_t)x[3];
      if (!in_window(tc)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{tc, tc}}, meta);
      cost.reset(new Image3D2DFactor(f->t_a_ns, load_v(d), meta, E->S_CtoI, E->p_CinI, d[3]));
      knots_of(meta);
      extra.push_back({x, 3});
      extra.push_back({x + 3, 1});
      break;
    }
    case CLIC_FACTOR_IMAGE_ONE_POSE: {
      if (!in_window(f->t_a_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{f->t_a_ns, f->t_a_ns}}, meta);
      cost.reset(new ImageFeatureOnePoseFactor(load_v(d), load_q(d + 3), load_v(d + 7), f->t_a_ns, load_v(d + 10), meta,
                                               E->S_CtoI, E->p_CinI, d[13]));
      knots_of(meta);
      extra.push_back({x, 1});
      break;
    }
    case CLIC_FACTOR_EPIPOLAR: {
      if (!in_window(f->t_a_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{f->t_a_ns, f->t_a_ns}}, meta);
      cost.reset(new EpipolarFactor(f->t_a_ns, load_v(d), load_v(d + 3), load_q(d + 6), load_v(d + 10), meta, d[13],
                                    E->S_CtoI, E->p_CinI));
      knots_of(meta);
      break;
    }
    case CLIC_FACTOR_LOAM_OPT_MAP_POSE: {
      if (!in_window(f->t_a_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      E->CaculateSplineMeta({{f->t_a_ns, f->t_a_ns}}, meta);
      cost.reset(new LoamFeatureOptMapPoseFactor(f->t_a_ns, load_correspondence(f), meta.segments.at(0), d[13],
                                                 load_q(d + 14), load_v(d + 18)));
      knots_of(meta);
      extra.push_back({x, 4});
      extra.push_back({x + 4, 3});
      break;
    }
    case CLIC_FACTOR_RELATIVE_LOAM: {
      if (!in_window(f->t_a_ns) || !in_window(f->t_b_ns)) return fail(CLIC_ERR_BAD_ARGUMENT, "time outside the window");
      const int64_t t_min = std::min(f->t_a_ns, f->t_b_ns), t_max = std::max(f->t_a_ns, f->t_b_ns);
      E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
      cost.reset(new RalativeLoamFeatureFactor(f->t_a_ns, f->t_b_ns, load_correspondence(f), meta, d[13], load_q(d + 14),
                                               load_v(d + 18)));
      knots_of(meta);
      break;
    }
    default: return fail(CLIC_ERR_BAD_ARGUMENT, "unknown factor kind");
  }
  const int nres = cost->num_residuals;
  int ncols = 6 * E->K;
  for (auto& e : extra) ncols += local_size(e.second);
  *n_res = nres;
  *n_cols = ncols;
  const size_t np = refs.size() + extra.size();
  if (np != cost->block_sizes.size()) return fail(CLIC_ERR_BAD_ARGUMENT, "parameter block count mismatch");
  std::vector<const double*> params(np);
  std::vector<std::vector<double>> jb(np);
  std::vector<double*> jptr(np);
  for (size_t i = 0; i < refs.size(); i++) params[i] = E->block_ptr(E->x, refs[i]);
  for (size_t i = 0; i < extra.size(); i++) params[refs.size() + i] = extra[i].first;
  for (size_t i = 0; i < np; i++) {
    jb[i].assign((size_t)nres * cost->block_sizes[i], 0.0);
    jptr[i] = jb[i].data();
  }
  std::vector<double> res(nres, 0.0);
  cost->Evaluate(params.data(), res.data(), jac ? jptr.data() : nullptr);
  for (int r = 0; r < nres; r++) residual[r] = res[r];
  if (!jac) return CLIC_OK;
  if ((int64_t)nres * ncols > jac_capacity) return fail(CLIC_ERR_BAD_ARGUMENT, "Jacobian buffer too small");
  for (int64_t i = 0; i < (int64_t)nres * ncols; i++) jac[i] = 0.0;
  int col = 6 * E->K;
  for (size_t i = 0; i < np; i++) {
    const int sz = cost->block_sizes[i], ls = local_size(sz);
    int c0;
    if (i < refs.size()) {
      c0 = 6 * refs[i].id + (refs[i].kind == CLIC_BLOCK_KNOT_POS ? 3 : 0);
    } else {
      c0 = col;
      col += ls;
    }
    for (int r = 0; r < nres; r++)
      for (int c = 0; c < ls; c++) jac[(size_t)r * ncols + c0 + c] += jb[i][(size_t)r * sz + c];
  }
  return CLIC_OK;
}

CLIC_API int clic_add_image(clic_estimator* E, int64_t n, const double* t_i, const double* p_i, const double* t_j,
                            const double* p_j, const int32_t* landmark, int32_t fixed_depth,
                            int32_t marg_this_feature, int64_t* n_dropped) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  int64_t dropped = 0;
  const int S = CLIC_SENSOR_CAMERA;
  for (int64_t i = 0; i < n; i++) {
    int lm = landmark[i];
    if (lm < 0 || lm >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "landmark index out of range");
    int64_t time_i_ns, time_j_ns;
    bool oki = E->MeasuredTimeToNs(S, t_i[i], time_i_ns);
    bool okj = E->MeasuredTimeToNs(S, t_j[i], time_j_ns);
    if (!oki || !okj) {  // ":757-758"
      dropped++;
      E->probe[2].push_back({time_i_ns, 1});
      E->probe[3].push_back({time_j_ns, 1});
      continue;
    }
    E->probe[2].push_back({time_i_ns, 0});
    E->probe[3].push_back({time_j_ns, 0});
    int64_t ti_c = (int64_t)(time_i_ns + E->x.t_offset_ns[S]);
    int64_t tj_c = (int64_t)(time_j_ns + E->x.t_offset_ns[S]);
    int64_t t_min = std::min(ti_c, tj_c), t_max = std::max(ti_c, tj_c);
    SplineMeta meta;
    if (E->opt.lock_t_offset[S]) {
      E->CaculateSplineMeta({{t_min, t_min}, {t_max, t_max}}, meta);
    } else {
      int64_t pad = E->opt.t_offset_padding_ns;
      E->CaculateSplineMeta({{t_min - pad, t_min + pad}, {t_max - pad, t_max + pad}}, meta);
    }
    ResidualBlock rb;
    rb.cost.reset(new ImageFeatureFactor(time_i_ns, load_v(p_i + 3 * i), time_j_ns, load_v(p_j + 3 * i), meta,
                                         E->S_CtoI, E->p_CinI,
                                         E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450. / 1.5));
    E->AddControlPoints(meta, rb.refs, false);
    E->AddControlPoints(meta, rb.refs, true);
    rb.refs.push_back({CLIC_BLOCK_INV_DEPTH, lm});
    rb.refs.push_back({CLIC_BLOCK_T_OFFSET, S});
    if (!fixed_depth) E->lm_free[lm] = 1;
    if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {
      double t_ns = (double)E->opt.t_offset_padding_ns;
      E->bounds_toff[S] = true;
      E->toff_lo[S] = E->x.t_offset_ns[S] - t_ns;
      E->toff_hi[S] = E->x.t_offset_ns[S] + t_ns;
    }
    rb.has_loss = true;
    rb.loss_a = marg_this_feature ? 1 : 2;  // ":806-810"
    rb.rtype = 11;
    if (E->opt.is_marg_state && marg_this_feature) {
      std::vector<int> drop;
      int Knot_size = 2 * (int)meta.NumParameters();
      drop.push_back(Knot_size);
      if (E->opt.marg_t_offset_param) drop.push_back(Knot_size + 1);
      prepare_marg_spline(E, std::move(rb), meta, drop);
    } else {
      E->blocks.push_back(std::move(rb));
    }
  }
  if (n_dropped) *n_dropped = dropped;
  return CLIC_OK;
}

CLIC_API int clic_add_prior(clic_estimator* E, const clic_prior* prior, int32_t n_drop, const int32_t* drop_blocks) {
  if (!valid(E) || !prior) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  ResidualBlock rb;
  rb.cost.reset(new MarginalizationFactor(prior));
  for (auto& b : prior->blocks) {
    BlockRef r{b.kind, b.id};
    if (b.kind == CLIC_BLOCK_KNOT_ROT || b.kind == CLIC_BLOCK_KNOT_POS) {
      r.id = b.id - E->knot_id0;
      if (r.id < 0 || r.id >= E->K) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to a knot outside the window");
    } else if (b.kind == CLIC_BLOCK_BIAS_GYRO || b.kind == CLIC_BLOCK_BIAS_ACCEL) {
      r.id = b.id - E->bias_id0;
      if (r.id < 0 || r.id >= E->n_bias) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to a bias slot outside the window");
    } else if (b.kind == CLIC_BLOCK_INV_DEPTH) {
      if (r.id < 0 || r.id >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to an unknown landmark");
    }
    rb.refs.push_back(r);
  }
  rb.rtype = 13;
  if (E->opt.is_marg_state) {
    if (n_drop <= 0) return CLIC_OK;  // trajectory_manager.cpp:367: only added when the drop set is non-empty
    rb.drop_set.assign(drop_blocks, drop_blocks + n_drop);
    E->marg_blocks.push_back(std::move(rb));
  } else {
    E->blocks.push_back(std::move(rb));
  }
  return CLIC_OK;
}

CLIC_API int clic_get_layout(clic_estimator* E, clic_layout* out) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  *out = E->layout();
  return CLIC_OK;
}

CLIC_API int clic_evaluate(clic_estimator* E, double* H, double* b, double* cost, double* fixed_cost) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  EvalOut ev;
  evaluate(E, E->x, H != nullptr || b != nullptr, ev);
  if (H) std::copy(ev.H.begin(), ev.H.end(), H);
  if (b) std::copy(ev.b.begin(), ev.b.end(), b);
  if (cost) *cost = ev.cost;
  if (fixed_cost) *fixed_cost = ev.fixed_cost;
  return CLIC_OK;
}

CLIC_API int clic_get_indices(clic_estimator* E, int32_t stream, int64_t capacity, int32_t* s_out, double* u_out,
                              int64_t* n_out) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (stream < 0 || stream > 3) return fail(CLIC_ERR_BAD_ARGUMENT, "stream");
  const auto& pr = E->probe[stream];
  int sensor = stream == 0 ? CLIC_SENSOR_LIDAR : (stream == 1 ? CLIC_SENSOR_IMU : CLIC_SENSOR_CAMERA);
  SegmentMeta whole(E->t0_ns, E->dt_ns, (size_t)E->K);
  int64_t cnt = 0;
  for (size_t i = 0; i < pr.size() && cnt < capacity; i++, cnt++) {
    if (pr[i].second) {
      s_out[cnt] = -1;
      u_out[cnt] = 0;
      continue;
    }
    int64_t t = pr[i].first;
    if (stream != 0) t = t + (int64_t)E->x.t_offset_ns[sensor];  // trajectory_value_factor.h:175
    size_t s;
    double u;
    if (!whole.computeTIndexNs(t, s, u)) {
      s_out[cnt] = -1;
      u_out[cnt] = 0;
    } else {
      s_out[cnt] = (int32_t)s;
      u_out[cnt] = u;
    }
  }
  if (n_out) *n_out = (int64_t)pr.size();
  return CLIC_OK;
}

CLIC_API int clic_solve(clic_estimator* E, int32_t max_iterations, clic_summary* summary) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!E->opt.lock_g) return fail(CLIC_ERR_UNSUPPORTED, "gravity must be locked (lock_g = 1)");
  clic_summary s;
  int rc = solve(E, max_iterations, &s);
  if (summary) *summary = s;
  return rc;
}

// ---- forward-only spline evaluation ----
namespace {
// Se3Spline::poseNs(int64_t) = So3Spline::evaluate (so3_spline.h:220-262) + RdSpline::evaluate<0> (rd_spline.h:218-244)
void pose_ns(const clic_estimator* E, int64_t time_ns, Quat* R_out, Vec3* p_out) {
  const int64_t st_ns = time_ns - E->t0_ns;  // computeTIndexNs, rd_spline.h:119-121
  const size_t s = (size_t)(st_ns / E->dt_ns);
  const double u = double(st_ns % E->dt_ns) / double(E->dt_ns);
  double p[N], coeff[N];
  base_coeffs_with_time(0, p, u);
  matvec4(blend_cumulative(), p, coeff);
  Quat res = load_q(&E->x.knot_q[4 * s]);
  for (int i = 0; i < N - 1; i++) {
    const Quat p0 = load_q(&E->x.knot_q[4 * (s + i)]), p1 = load_q(&E->x.knot_q[4 * (s + i + 1)]);
    const Quat r01 = so3_mul(qconj(p0), p1);
    const Vec3 delta = so3_log(r01);
    const Vec3 kdelta = delta * coeff[i + 1];
    res = so3_mul(res, so3_exp(kdelta));
  }
  double c[N];
  matvec4(blend_plain(), p, c);  // pow_inv_dt_[0] = 1
  Vec3 pos = v3(0, 0, 0);
  for (int i = 0; i < N; i++) pos = pos + c[i] * load_v(&E->x.knot_p[3 * (s + i)]);
  *R_out = res;
  *p_out = pos;
}
// Trajectory::GetSensorPose, trajectory.cpp:100-115 (the assert becomes a bool)
bool sensor_pose(const clic_estimator* E, double timestamp, const Quat& S_StoI, const Vec3& p_SinI, double t_offset_ns,
                 Quat* R_out, Vec3* p_out) {
  const double time_ns = timestamp * S_TO_NS + t_offset_ns;
  if (!(time_ns >= (double)E->minTimeNs() && time_ns < (double)E->maxTimeNs())) return false;
  Quat R;
  Vec3 p;
  pose_ns(E, (int64_t)time_ns, &R, &p);
  *R_out = so3_mul(R, S_StoI);  // pose_I_to_G * EP_StoI.se3
  *p_out = so3_rotate(R, p_SinI) + p;
  return true;
}
}  // namespace

CLIC_API int clic_query_poses(clic_estimator* E, int64_t n, const double* t_s, const double S_StoI[4],
                              const double p_SinI[3], double t_offset_ns, double* q_out, double* p_out,
                              int64_t* n_invalid) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n > 0 && (!t_s || !S_StoI || !p_SinI || !q_out || !p_out)) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  int64_t bad = 0;
  const Quat S = load_q(S_StoI);
  const Vec3 ps = load_v(p_SinI);
  for (int64_t i = 0; i < n; i++) {
    Quat R{0, 0, 0, 1};
    Vec3 p = v3(0, 0, 0);
    if (!sensor_pose(E, t_s[i], S, ps, t_offset_ns, &R, &p)) { bad++; R = Quat{0, 0, 0, 1}; p = v3(0, 0, 0); }
    q_out[4 * i] = R.x; q_out[4 * i + 1] = R.y; q_out[4 * i + 2] = R.z; q_out[4 * i + 3] = R.w;
    p_out[3 * i] = p[0]; p_out[3 * i + 1] = p[1]; p_out[3 * i + 2] = p[2];
  }
  if (n_invalid) *n_invalid = bad;
  return CLIC_OK;
}

// ---- LiDAR data association, see association.hpp ----
struct clic_feature_map {
  std::vector<float> surf, corner;
  std::vector<int32_t> nn[2];
  assoc::Grid grid[2];  // corner, surface: built only in grid-search mode
  bool use_grid = false;
};
static bool g_assoc_grid = false;
// Oracle-only switch for the timed CPU baseline: 0 exhaustive search (the definition), 1 the same search through a
// uniform grid.  Applies to feature maps created afterwards.
CLIC_API int clic_oracle_set_association_search(int mode) {
  g_assoc_grid = mode != 0;
  return CLIC_OK;
}

CLIC_API int clic_feature_map_create(const float* surf_xyz, int64_t n_surf, const float* corner_xyz, int64_t n_corner,
                                     clic_feature_map** out) {
  if (!out || n_surf < 0 || n_corner < 0 || (n_surf > 0 && !surf_xyz) || (n_corner > 0 && !corner_xyz))
    return fail(CLIC_ERR_BAD_ARGUMENT, "bad feature map");
  auto* m = new clic_feature_map;
  m->surf.assign(surf_xyz, surf_xyz + 3 * n_surf);
  if (n_corner > 0) m->corner.assign(corner_xyz, corner_xyz + 3 * n_corner);
  m->use_grid = g_assoc_grid;
  if (m->use_grid) {
    m->grid[0].build(m->corner.data(), n_corner);
    m->grid[1].build(m->surf.data(), n_surf);
  }
  *out = m;
  return CLIC_OK;
}
CLIC_API void clic_feature_map_destroy(clic_feature_map* m) { delete m; }

CLIC_API int64_t clic_correspondence_capacity(int64_t n) {
  if (n <= 0) return 0;
  const int64_t step = n / 1500 + 1;  // lidar_odometry.cpp:501, 575
  return (n + step - 1) / step;
}

CLIC_API int clic_find_correspondence(clic_feature_map* m, int64_t n_corner, const float* corner_xyz_L,
                                      const double* corner_t, const float* corner_xyz_M, int64_t n_surf,
                                      const float* surf_xyz_L, const double* surf_t, const float* surf_xyz_M,
                                      int64_t* n_lines, double* line_t, double* line_point, double* line_geo,
                                      int64_t* line_src, int64_t* n_planes, double* plane_t, double* plane_point,
                                      double* plane_geo, int64_t* plane_src) {
  if (!m || !n_lines || !n_planes) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  *n_lines = *n_planes = 0;
  m->nn[0].clear();
  m->nn[1].clear();
  if ((n_corner > 0 && (!corner_xyz_L || !corner_t || !corner_xyz_M || !line_t || !line_point || !line_geo)) ||
      (n_surf > 0 && (!surf_xyz_L || !surf_t || !surf_xyz_M || !plane_t || !plane_point || !plane_geo)))
    return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  const int64_t n_map_c = (int64_t)m->corner.size() / 3, n_map_s = (int64_t)m->surf.size() / 3;
  float pts[5][3];
  auto gather = [&](const std::vector<float>& map, const assoc::Neighbours& nb) {
    for (int j = 0; j < 5; j++)
      for (int c = 0; c < 3; c++) pts[j][c] = map[3 * (size_t)nb.idx[j] + c];
  };
  if (n_corner > 0 && n_map_c > 10) {  // lidar_odometry.cpp:502
    const int64_t step = n_corner / 1500 + 1;
    for (int64_t i = 0; i < n_corner; i += step) {
      const assoc::Neighbours nb = m->use_grid ? m->grid[0].knn5(m->corner.data(), corner_xyz_M + 3 * i)
                                               : assoc::knn5(m->corner.data(), n_map_c, corner_xyz_M + 3 * i);
      for (int k = 0; k < 5; k++) m->nn[0].push_back(nb.d2[k] <= 1.0f ? nb.idx[k] : INT32_MAX);
      if (!(nb.d2[4] <= 1.0f)) continue;  // ":508"
      gather(m->corner, nb);
      double cp[3], dir[3];
      if (!assoc::fit_line(pts, cp, dir)) continue;
      const int64_t o = (*n_lines)++;
      line_t[o] = corner_t[i];
      for (int c = 0; c < 3; c++) {
        line_point[3 * o + c] = (double)corner_xyz_L[3 * i + c];
        line_geo[6 * o + c] = cp[c];
        line_geo[6 * o + 3 + c] = dir[c];
      }
      if (line_src) line_src[o] = i;
    }
  }
  if (n_surf > 0) {
    const int64_t step = n_surf / 1500 + 1;
    for (int64_t i = 0; i < n_surf; i += step) {
      if (surf_t[i] < 0) {  // ":577"
        m->nn[1].insert(m->nn[1].end(), 5, -1);
        continue;
      }
      const assoc::Neighbours nb = m->use_grid ? m->grid[1].knn5(m->surf.data(), surf_xyz_M + 3 * i)
                                               : assoc::knn5(m->surf.data(), n_map_s, surf_xyz_M + 3 * i);
      for (int k = 0; k < 5; k++) m->nn[1].push_back(nb.d2[k] <= 1.0f ? nb.idx[k] : INT32_MAX);
      if (!(nb.d2[4] <= 1.0f)) continue;  // ":584"; fewer than 5 map points leaves +inf here (":593")
      gather(m->surf, nb);
      double pl[4];
      if (!assoc::fit_plane(pts, pl)) continue;
      const int64_t o = (*n_planes)++;
      plane_t[o] = surf_t[i];
      for (int c = 0; c < 3; c++) plane_point[3 * o + c] = (double)surf_xyz_L[3 * i + c];
      for (int c = 0; c < 4; c++) plane_geo[4 * o + c] = pl[c];
      if (plane_src) plane_src[o] = i;
    }
  }
  return CLIC_OK;
}

CLIC_API int clic_undistort_scan(clic_estimator* E, int64_t n, const double* t_s, const float* xyz_in,
                                 const double S_LtoI[4], const double p_LinI[3], double t_offset_ns, int32_t in_global,
                                 double target_t_s, float* xyz_out, int64_t* n_invalid);

// lidar_odometry.cpp:148-180 + trajectory_manager.cpp:676-688, composed from the restatements above.
CLIC_API int clic_add_loam_from_scan(clic_estimator* E, clic_feature_map* m, int64_t n_corner, const float* corner_xyz,
                                     const double* corner_t, int64_t n_surf, const float* surf_xyz, const double* surf_t,
                                     const double S_GtoM[4], const double p_GinM[3], const double S_LtoI[4],
                                     const double p_LinI[3], double t_offset_ns, int32_t cor_downsample, double weight,
                                     int32_t marg_this_factor, int64_t* n_lines, int64_t* n_planes, int64_t* n_dropped) {
  if (!valid(E) || !m) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (cor_downsample < 1) return fail(CLIC_ERR_BAD_ARGUMENT, "bad argument");
  if (n_lines) *n_lines = 0;
  if (n_planes) *n_planes = 0;
  if (n_dropped) *n_dropped = 0;
  std::vector<float> cM(3 * (size_t)std::max<int64_t>(n_corner, 0)), sM(3 * (size_t)std::max<int64_t>(n_surf, 0));
  int rc;
  if ((rc = clic_undistort_scan(E, n_corner, corner_t, corner_xyz, S_LtoI, p_LinI, t_offset_ns, 1, 0.0, cM.data(), nullptr)) ||
      (rc = clic_undistort_scan(E, n_surf, surf_t, surf_xyz, S_LtoI, p_LinI, t_offset_ns, 1, 0.0, sM.data(), nullptr)))
    return rc;
  const int64_t cap_c = clic_correspondence_capacity(n_corner), cap_s = clic_correspondence_capacity(n_surf);
  std::vector<double> lt(cap_c), lp(3 * cap_c), lg(6 * cap_c), pt(cap_s), pp(3 * cap_s), pg(4 * cap_s);
  int64_t nl = 0, np = 0;
  if ((rc = clic_find_correspondence(m, n_corner, corner_xyz, corner_t, cM.data(), n_surf, surf_xyz, surf_t, sM.data(), &nl,
                                     lt.data(), lp.data(), lg.data(), nullptr, &np, pt.data(), pp.data(), pg.data(),
                                     nullptr)))
    return rc;
  // std::sort(by_timestamp) over [lines..., planes...] — stable here — then every cor_downsample-th record
  std::vector<int64_t> order(nl + np);
  for (int64_t i = 0; i < nl + np; i++) order[i] = i;
  auto tp = [&](int64_t i) { return i < nl ? lt[i] : pt[i - nl]; };
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return tp(a) < tp(b); });
  std::vector<double> klt, klp, klg, kpt, kpp, kpg;
  for (size_t k = 0; k < order.size(); k += cor_downsample) {
    const int64_t i = order[k];
    if (i < nl) {
      klt.push_back(lt[i]);
      klp.insert(klp.end(), lp.begin() + 3 * i, lp.begin() + 3 * i + 3);
      klg.insert(klg.end(), lg.begin() + 6 * i, lg.begin() + 6 * i + 6);
    } else {
      const int64_t j = i - nl;
      kpt.push_back(pt[j]);
      kpp.insert(kpp.end(), pp.begin() + 3 * j, pp.begin() + 3 * j + 3);
      kpg.insert(kpg.end(), pg.begin() + 4 * j, pg.begin() + 4 * j + 4);
    }
  }
  int64_t d0 = 0, d1 = 0;
  if ((rc = clic_add_loam_lines(E, (int64_t)klt.size(), klt.data(), klp.data(), klg.data(), S_GtoM, p_GinM, S_LtoI, p_LinI,
                                weight, marg_this_factor, &d0)) ||
      (rc = clic_add_loam_planes(E, (int64_t)kpt.size(), kpt.data(), kpp.data(), kpg.data(), S_GtoM, p_GinM, S_LtoI, p_LinI,
                                 weight, marg_this_factor, &d1)))
    return rc;
  if (n_lines) *n_lines = (int64_t)klt.size();
  if (n_planes) *n_planes = (int64_t)kpt.size();
  if (n_dropped) *n_dropped = d0 + d1;
  return CLIC_OK;
}

CLIC_API int clic_correspondence_neighbours(clic_feature_map* m, int32_t type, int64_t capacity, int32_t* idx,
                                            int64_t* n_out) {
  if (!m || type < 0 || type > 1 || !n_out) return fail(CLIC_ERR_BAD_ARGUMENT, "bad argument");
  const int64_t n = (int64_t)m->nn[type].size() / 5;
  *n_out = n;
  if (n == 0 || !idx) return CLIC_OK;
  if (capacity < n) return fail(CLIC_ERR_BAD_ARGUMENT, "capacity too small");
  std::copy(m->nn[type].begin(), m->nn[type].end(), idx);
  return CLIC_OK;
}
CLIC_API int clic_feature_map_launches(clic_feature_map* m, int64_t* launches, double* last_kernel_ms) {
  if (launches) *launches = 0;
  if (last_kernel_ms) *last_kernel_ms = 0;
  return CLIC_OK;
}

CLIC_API int clic_undistort_scan(clic_estimator* E, int64_t n, const double* t_s, const float* xyz_in,
                                 const double S_LtoI[4], const double p_LinI[3], double t_offset_ns, int32_t in_global,
                                 double target_t_s, float* xyz_out, int64_t* n_invalid) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n > 0 && (!t_s || !xyz_in || !S_LtoI || !p_LinI || !xyz_out)) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  const Quat S = load_q(S_LtoI);
  const Vec3 ps = load_v(p_LinI);
  Quat Rt{0, 0, 0, 1};
  Vec3 pt = v3(0, 0, 0);
  if (!in_global) {  // pose_G_to_target = GetLidarPose(target_timestamp).inverse(), trajectory.cpp:42
    Quat R;
    Vec3 p;
    if (!sensor_pose(E, target_t_s, S, ps, t_offset_ns, &R, &p)) return fail(CLIC_ERR_BAD_ARGUMENT, "target time outside the trajectory");
    Rt = qconj(R);
    pt = -so3_rotate(Rt, p);
  }
  const float nanf_ = std::numeric_limits<float>::quiet_NaN();
  int64_t bad = 0;
  for (int64_t i = 0; i < n; i++) {
    Quat R;
    Vec3 p;
    if (std::isnan(xyz_in[3 * i]) || !sensor_pose(E, t_s[i], S, ps, t_offset_ns, &R, &p)) {
      bad++;
      xyz_out[3 * i] = xyz_out[3 * i + 1] = xyz_out[3 * i + 2] = nanf_;
      continue;
    }
    const Vec3 p_Lk = v3(xyz_in[3 * i], xyz_in[3 * i + 1], xyz_in[3 * i + 2]);
    Vec3 out;
    if (in_global) {
      out = so3_rotate(R, p_Lk) + p;  // pose_Lk_to_G * p_Lk
    } else {                          // (pose_G_to_target * pose_Lk_to_G) * p_Lk: the poses are composed first
      const Quat Rc = so3_mul(Rt, R);
      const Vec3 pc = so3_rotate(Rt, p) + pt;
      out = so3_rotate(Rc, p_Lk) + pc;
    }
    xyz_out[3 * i] = (float)out[0]; xyz_out[3 * i + 1] = (float)out[1]; xyz_out[3 * i + 2] = (float)out[2];
  }
  if (n_invalid) *n_invalid = bad;
  return CLIC_OK;
}

CLIC_API int clic_read_state(clic_estimator* E, double* knot_q, double* knot_p, double* bias, double* t_offset_ns,
                             double* inv_depth) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (knot_q) std::copy(E->x.knot_q.begin(), E->x.knot_q.end(), knot_q);
  if (knot_p) std::copy(E->x.knot_p.begin(), E->x.knot_p.end(), knot_p);
  if (bias) std::copy(E->x.bias.begin(), E->x.bias.end(), bias);
  if (t_offset_ns)
    for (int s = 0; s < 3; s++) t_offset_ns[s] = E->x.t_offset_ns[s];
  if (inv_depth) std::copy(E->x.inv_depth.begin(), E->x.inv_depth.end(), inv_depth);
  return CLIC_OK;
}

CLIC_API int clic_write_state(clic_estimator* E, const double* knot_q, const double* knot_p, const double* bias,
                              const double* t_offset_ns, const double* inv_depth) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (knot_q) E->x.knot_q.assign(knot_q, knot_q + (size_t)4 * E->K);
  if (knot_p) E->x.knot_p.assign(knot_p, knot_p + (size_t)3 * E->K);
  if (bias && E->n_bias > 0) E->x.bias.assign(bias, bias + (size_t)6 * E->n_bias);
  if (t_offset_ns)
    for (int s = 0; s < 3; s++) E->x.t_offset_ns[s] = t_offset_ns[s];
  if (inv_depth && E->n_lm > 0) E->x.inv_depth.assign(inv_depth, inv_depth + E->n_lm);
  return CLIC_OK;
}

CLIC_API int clic_marginalize(clic_estimator* E, clic_prior** out) {
  if (!valid(E) || !out) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  return marginalize(E, out);
}

CLIC_API int clic_prior_create(int32_t n, const double* J0, const double* r0, int32_t n_blocks, const int32_t* kind,
                               const int32_t* id, const double* x0, clic_prior** out) {
  if (n <= 0 || !J0 || !r0 || !out) return fail(CLIC_ERR_BAD_ARGUMENT, "bad prior");
  auto* P = new clic_prior;
  P->n = n;
  P->m = 0;
  P->J0.assign(J0, J0 + (size_t)n * n);
  P->r0.assign(r0, r0 + n);
  int off = 0;
  for (int i = 0; i < n_blocks; i++) {
    clic_prior::Block b;
    b.kind = kind[i];
    b.id = id[i];
    b.size = block_size(kind[i]);
    b.idx = off;
    off += local_size(b.size);
    for (int k = 0; k < 4; k++) b.x0[k] = x0[4 * i + k];
    P->blocks.push_back(b);
  }
  if (off != n) {
    delete P;
    return fail(CLIC_ERR_BAD_ARGUMENT, "prior blocks do not add up to n");
  }
  *out = P;
  return CLIC_OK;
}

CLIC_API int clic_prior_dims(const clic_prior* p, int32_t* n, int32_t* n_blocks, int32_t* m_dropped) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  if (n) *n = p->n;
  if (n_blocks) *n_blocks = (int32_t)p->blocks.size();
  if (m_dropped) *m_dropped = p->m;
  return CLIC_OK;
}

CLIC_API int clic_prior_read(const clic_prior* p, double* J0, double* r0, int32_t* kind, int32_t* id,
                             int32_t* tangent_offset, double* x0) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  if (J0) std::copy(p->J0.begin(), p->J0.end(), J0);
  if (r0) std::copy(p->r0.begin(), p->r0.end(), r0);
  for (size_t i = 0; i < p->blocks.size(); i++) {
    if (kind) kind[i] = p->blocks[i].kind;
    if (id) id[i] = p->blocks[i].id;
    if (tangent_offset) tangent_offset[i] = p->blocks[i].idx;
    if (x0)
      for (int k = 0; k < 4; k++) x0[4 * i + k] = p->blocks[i].x0[k];
  }
  return CLIC_OK;
}

CLIC_API int clic_prior_normal(const clic_prior* p, double* A, double* g) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  int n = p->n;
  if (A)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double acc = 0;
        for (int k = 0; k < n; k++) acc += p->J0[(size_t)k * n + i] * p->J0[(size_t)k * n + j];
        A[(size_t)i * n + j] = acc;
      }
  if (g)
    for (int i = 0; i < n; i++) {
      double acc = 0;
      for (int k = 0; k < n; k++) acc += p->J0[(size_t)k * n + i] * p->r0[k];
      g[i] = acc;
    }
  return CLIC_OK;
}

CLIC_API int clic_prior_release(clic_prior* p) {
  delete p;
  return CLIC_OK;
}

CLIC_API int clic_comm_unique_id(uint8_t*) { return fail(CLIC_ERR_UNSUPPORTED, "oracle is single-process"); }
CLIC_API int clic_comm_init(clic_estimator*, int32_t, int32_t, const uint8_t*) {
  return fail(CLIC_ERR_UNSUPPORTED, "oracle is single-process");
}
CLIC_API int clic_comm_p2p_export(clic_estimator*, uint8_t handle[64]) { std::memset(handle, 0, 64); return CLIC_OK; }
CLIC_API int clic_comm_p2p_import(clic_estimator*, int32_t, int32_t, const uint8_t*) { return CLIC_OK; }
CLIC_API int clic_comm_p2p_ready(void) { return 0; }
CLIC_API int clic_comm_p2p_disable(void) { return CLIC_OK; }

CLIC_API int clic_linearize_async(clic_estimator* E, int32_t reps) {
  if (!valid(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  EvalOut ev;
  for (int i = 0; i < reps; i++) evaluate(E, E->x, true, ev);
  return CLIC_OK;
}
CLIC_API int clic_synchronize(clic_estimator*) { return CLIC_OK; }
CLIC_API int clic_num_kernel_launches(clic_estimator*, int64_t* l) {
  if (l) *l = 0;
  return CLIC_OK;
}
CLIC_API int clic_pass_info(clic_estimator*, int32_t* fused, int32_t* grid, int32_t part_ctas[6]) {
  if (fused) *fused = 0;
  if (grid) *grid = 0;
  if (part_ctas) for (int k = 0; k < 6; k++) part_ctas[k] = 0;
  return CLIC_OK;
}
CLIC_API int clic_pass_stamps(clic_estimator*, int32_t, int64_t, uint64_t*, int32_t* n_ctas) {
  if (n_ctas) *n_ctas = 0;
  return CLIC_OK;
}
CLIC_API int clic_profile_enable(clic_estimator*, int32_t) { return CLIC_OK; }
CLIC_API int clic_profile_read(clic_estimator*, double* ms, int64_t* n) {
  for (int i = 0; i < 8; i++) {
    if (ms) ms[i] = 0;
    if (n) n[i] = 0;
  }
  return CLIC_OK;
}

}  // extern "C"
