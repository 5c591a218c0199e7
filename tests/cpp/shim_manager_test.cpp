// The boundary as TrajectoryManager uses it (src/estimator/trajectory_manager.cpp): InitFactorInfo's static visual-factor
// parameters (:53-61), InitTrajWithPropagation's SetTimeoffsetState (:295), the LIC solve with image factors (:637-756),
// UpdateLIOPrior's marginalization estimator fed with PrepareMarginalizationInfo(RType_Prior, new
// MarginalizationFactor(info), NULL, blocks, drop_set) (:346-376) + AddGravityFactor + SaveMarginalizationInfo, and the
// next solve on the new prior.  Runs on the oracle (CPU) and on libclic_cuda.so; prints "SHIM2_OK <costs...>".
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>

#include "../../include/clic/trajectory_estimator.hpp"

using namespace clic;

static int fail(const char* what) {
  std::printf("SHIM2_FAIL %s\n", what);
  return 1;
}

int main() {
  auto traj = std::make_shared<Trajectory>(0.1);
  std::mt19937 rng(5);
  std::normal_distribution<double> N(0.0, 1.0);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  const int K = 14;
  for (int k = 0; k < K; k++) {
    double t = 0.1 * (k - 1);
    double w[3] = {0.3 * std::sin(0.9 * t) + 0.005 * N(rng), 0.2 * std::cos(0.6 * t) + 0.005 * N(rng), 0.5 * t};
    double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), s = th > 0 ? std::sin(0.5 * th) / th : 0.5;
    traj->knots_so3.push_back({s * w[0], s * w[1], s * w[2], std::cos(0.5 * th)});
    traj->knots_pos.push_back({2 * std::sin(0.7 * t) + 0.01 * N(rng), 1.5 * std::cos(0.5 * t) + 0.01 * N(rng), 0.3 * std::sin(1.1 * t)});
  }
  ExtrinsicParam& Ep_CtoI = traj->EP_StoI[CameraSensor];
  Ep_CtoI.so3.q = {{0.5, -0.5, 0.5, -0.5}};
  Ep_CtoI.p = {0.05, 0.02, -0.03};
  // TrajectoryManager::InitFactorInfo (trajectory_manager.cpp:53-61), image_weight = 200 (config/ct_odometry_ntu.yaml:35)
  const double image_feature_weight = 200.0;
  Matrix2d sqrt_info = image_feature_weight * Matrix2d::Identity();
  analytic_derivative::ImageFeatureFactor::SetParam(Ep_CtoI.so3, Ep_CtoI.p);
  analytic_derivative::ImageFeatureFactor::sqrt_info = sqrt_info;

  SO3d I;
  Vec3d zero{};
  double bg0[3] = {0, 0, 0}, ba0[3] = {0, 0, 0}, bg1[3] = {0.01, 0, 0}, ba1[3] = {0, 0.02, 0}, g[3] = {0, 0, 9.8};
  Vec6d info = {800, 800, 800, 36, 36, 36};
  auto add_lidar = [&](TrajectoryEstimator::Ptr& est, double t_lo, double t_hi, int n, bool marg) {
    for (int i = 0; i < n; i++) {
      PointCorrespondence pc;
      pc.t_point = t_lo + (t_hi - t_lo) * U(rng);
      pc.point = {10 * N(rng), 10 * N(rng), 3 * N(rng)};
      pc.geo_type = (i % 3 == 0) ? Line : Plane;
      double n3[3] = {N(rng), N(rng), N(rng)}, nn = std::sqrt(n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
      pc.geo_plane = {n3[0] / nn, n3[1] / nn, n3[2] / nn, 0.003 * N(rng)};
      pc.geo_normal = {n3[1] / nn, n3[2] / nn, n3[0] / nn};
      pc.geo_point = {N(rng), N(rng), N(rng)};
      est->AddLoamMeasurementAnalytic(pc, I, zero, I, zero, 5.0, marg);
    }
  };
  auto add_imu = [&](TrajectoryEstimator::Ptr& est, double t_lo, int n, bool marg) {
    for (int i = 0; i < n; i++) {
      IMUData d;
      d.timestamp = t_lo + i * 0.0101;
      d.gyro = {0.1, 0.2, 0.5};
      d.accel = {0.1, 0.2, 9.8};
      est->AddIMUMeasurementAnalytic(d, bg1, ba1, g, info, marg);
    }
  };

  // ---- solve 1 (UpdateTrajectoryWithLoamFeature, :637-756): LiDAR + image + IMU + bias ----
  double cost1_init, cost1_final;
  {
    TrajectoryEstimatorOptions option;
    option.lock_ab = false;
    option.lock_wb = false;
    TrajectoryEstimator::Ptr estimator(new TrajectoryEstimator(traj, option, "Update Traj"));
    estimator->SetFixedIndex(0);
    add_lidar(estimator, 0.001, 1.09, 600, false);
    double inv_depth[6] = {0.2, 0.25, 0.15, 0.3, 0.22, 0.18};
    for (int l = 0; l < 6; l++)
      for (int o = 1; o <= 4; o++) {
        Vec3d pi = {0.1 * (l - 3), 0.05 * l, 1.0}, pj = {0.1 * (l - 3) + 0.02 * o, 0.05 * l - 0.01 * o, 1.0};
        estimator->AddImageFeatureAnalytic(0.05, pi, 0.05 + 0.2 * o, pj, &inv_depth[l], true);
      }
    add_imu(estimator, 0.01, 100, false);
    Vec6d binfo = {1e4, 1e4, 1e4, 1e3, 1e3, 1e3};
    estimator->AddBiasFactor(bg0, bg1, ba0, ba1, 1, binfo);
    estimator->SetTimeoffsetState();  // as InitTrajWithPropagation does (:295)
    // the remaining adders of the class (trajectory_estimator.h:146-204): pose, local velocity, relative rotation,
    // image depth
    PoseData pose;
    pose.timestamp = 0.4;
    pose.position = {2 * std::sin(0.28), 1.5 * std::cos(0.2), 0.3 * std::sin(0.44)};
    pose.orientation.q = {{0.05, 0.04, 0.1, 0.9929}};
    estimator->AddPoseMeasurementAnalytic(pose, Vec6d{{20, 20, 20, 8, 8, 8}});
    estimator->AddStartTimePose(pose);
    estimator->AddLocalVelocityMeasurementAnalytic(0.52, Vec3d{{1.1, 0.2, 0.05}}, 3.0);
    SO3d S_BtoA;
    S_BtoA.q = {{0.02, -0.03, 0.15, 0.9881}};
    estimator->AddRelativeRotationAnalytic(0.31, 0.62, S_BtoA, Vec3d{{12, 14, 16}});
    SO3d S_CitoCj;
    S_CitoCj.q = {{0.01, 0.02, -0.01, 0.9997}};
    estimator->AddImageDepthAnalytic(Vec3d{{0.1, 0.05, 1.0}}, Vec3d{{0.13, 0.03, 1.0}}, S_CitoCj, Vec3d{{0.1, -0.05, 0.02}},
                                     &inv_depth[0]);
    // PreIntegrationFactor between the two bias slots (trajectory_estimator.h:168-172); weak, so the solve keeps its shape
    IntegrationBase pre;
    pre.sum_dt = 0.3;
    pre.delta_p = {0.31, -0.02, 0.01};
    pre.delta_v = {0.1, 0.05, -0.02};
    pre.delta_q = {{0.01, -0.02, 0.03, 0.9993}};
    for (int i = 0; i < 15; i++) {
      pre.jacobian[i * 15 + i] = 1.0;
      pre.covariance[i * 15 + i] = 4.0 + 0.1 * i;
    }
    pre.jacobian[0 * 15 + 9] = -0.04;   // dp/dba
    pre.jacobian[3 * 15 + 12] = -0.3;   // dq/dbg
    pre.jacobian[6 * 15 + 9] = -0.3;    // dv/dba
    pre.covariance[0 * 15 + 1] = pre.covariance[1 * 15 + 0] = 0.5;
    {  // IntegrationBase::to_abi: sqrt_info^T sqrt_info = covariance^-1, sqrt_info upper triangular
      const clic_preintegration abi = pre.to_abi();
      double worst = 0.0;
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < 15; c++) {
          double m = 0.0;  // (S^T S P)(r, c)
          for (int k = 0; k < 15; k++) {
            double sts = 0.0;
            for (int q = 0; q < 15; q++) sts += abi.sqrt_info[q * 15 + r] * abi.sqrt_info[q * 15 + k];
            m += sts * pre.covariance[k * 15 + c];
          }
          worst = std::max(worst, std::fabs(m - (r == c ? 1.0 : 0.0)));
          if (r > c && abi.sqrt_info[r * 15 + c] != 0.0) return fail("sqrt_info is not upper triangular");
        }
      if (worst > 1e-12) return fail("IntegrationBase::to_abi: sqrt_info^T sqrt_info covariance != I");
    }
    estimator->AddPreIntegrationAnalytic(0.3, 0.6, &pre, bg0, bg1, ba0, ba1);
    {  // one factor's residual / Jacobian (the reference's test_analytic_factor.cpp surface): an epipolar constraint
      clic_factor_desc fd{};
      fd.kind = CLIC_FACTOR_EPIPOLAR;
      fd.t_a_ns = 530000000;
      const double dd[14] = {0.1, 0.2, 1, 1.3, 1.2, 1, 0, 0, 0, 1, 1.3, 1.2, 1.0, 0.11};
      for (int i = 0; i < 14; i++) fd.d[i] = dd[i];
      std::vector<double> r, J;
      int nc = 0;
      if (!estimator->EvaluateFactor(fd, &r, &J, &nc) || r.size() != 1 || nc != 6 * (int)traj->numKnots() ||
          J.size() != (size_t)nc)
        return fail("EvaluateFactor(epipolar)");
      int nz = 0;
      for (double v : J) nz += v != 0.0;
      if (nz != 24) return fail("EvaluateFactor: an epipolar factor touches 4 knots x 6 columns");
    }
    Summary summary = estimator->Solve(8, false);
    cost1_init = summary.initial_cost;
    cost1_final = summary.final_cost;
  }

  // ---- a first prior: marginalise knots < 4 out of LiDAR + IMU factors (no previous prior yet) ----
  MarginalizationInfo::Ptr lidar_marg_info;
  std::vector<double*> lidar_marg_parameter_blocks;
  {
    TrajectoryEstimatorOptions option;
    option.is_marg_state = true;
    option.marg_bias_param = false;
    option.ctrl_to_be_opt_now = 0;
    option.ctrl_to_be_opt_later = 4;
    TrajectoryEstimator::Ptr estimator(new TrajectoryEstimator(traj, option, "Update Prior 0"));
    add_lidar(estimator, 0.001, 0.69, 300, true);
    add_imu(estimator, 0.01, 60, true);
    estimator->SaveMarginalizationInfo(lidar_marg_info, lidar_marg_parameter_blocks);
  }
  if (!lidar_marg_info) { std::printf("SHIM2_FAIL no first prior\n"); return 1; }
  const int n0 = lidar_marg_info->n, m0 = lidar_marg_info->m;

  // ---- UpdateLIOPrior (:346-465): previous prior through PrepareMarginalizationInfo(RType_Prior, ...) ----
  MarginalizationInfo::Ptr next_info;
  std::vector<double*> next_blocks;
  {
    TrajectoryEstimatorOptions option;
    option.is_marg_state = true;
    option.marg_bias_param = false;
    option.marg_gravity_param = true;
    option.ctrl_to_be_opt_now = 4;
    option.ctrl_to_be_opt_later = 8;
    TrajectoryEstimator::Ptr estimator(new TrajectoryEstimator(traj, option, "Update Prior"));
    std::vector<double*> drop_param_set;
    for (int i = option.ctrl_to_be_opt_now; i < option.ctrl_to_be_opt_later; ++i) {
      drop_param_set.emplace_back(traj->getKnotSO3(i));
      drop_param_set.emplace_back(traj->getKnotPos(i));
    }
    std::vector<int> drop_set;
    for (int j = 0; j < (int)lidar_marg_parameter_blocks.size(); j++)
      for (auto const& drop_param : drop_param_set)
        if (lidar_marg_parameter_blocks[j] == drop_param) { drop_set.emplace_back(j); break; }
    if (!drop_set.empty()) {
      MarginalizationFactor* marginalization_factor = new MarginalizationFactor(lidar_marg_info);
      estimator->PrepareMarginalizationInfo(RType_Prior, marginalization_factor, NULL, lidar_marg_parameter_blocks, drop_set);
    }
    add_lidar(estimator, 0.3, 1.09, 400, true);
    add_imu(estimator, 0.31, 70, true);
    Vec3d ginfo = {50, 50, 50};
    estimator->AddGravityFactor(g, ginfo, true);
    estimator->SaveMarginalizationInfo(next_info, next_blocks);
  }
  if (!next_info) { std::printf("SHIM2_FAIL no second prior\n"); return 1; }

  // ---- next solve on the new prior (:649-666) ----
  double cost2_init, cost2_final;
  {
    TrajectoryEstimatorOptions option;
    option.lock_ab = false;
    option.lock_wb = false;
    TrajectoryEstimator::Ptr estimator(new TrajectoryEstimator(traj, option, "Update Traj 2"));
    estimator->SetFixedIndex(7);
    estimator->AddMarginalizationFactor(next_info, next_blocks);
    add_lidar(estimator, 0.5, 1.09, 500, false);
    add_imu(estimator, 0.5, 55, false);
    Vec3d ginfo = {50, 50, 50};
    estimator->AddGravityFactor(g, ginfo);  // constant block in a solve: adds nothing but must be accepted
    Summary summary = estimator->Solve(8, false);
    cost2_init = summary.initial_cost;
    cost2_final = summary.final_cost;
  }
  bool ok = cost1_final < cost1_init && cost2_final <= cost2_init && std::isfinite(cost2_final) && m0 >= 24 &&
            next_info->m >= 24 && next_info->n > 0 && !next_blocks.empty();
  std::printf("%s %.9e %.9e %.9e %.9e %d %d %d %d\n", ok ? "SHIM2_OK" : "SHIM2_FAIL", cost1_init, cost1_final, cost2_init,
              cost2_final, n0, m0, next_info->n, next_info->m);
  return ok ? 0 : 1;
}
