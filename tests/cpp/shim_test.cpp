// Compiles reference-style host code against include/clic/trajectory_estimator.hpp and runs it on the oracle
// library (CPU) — the same source links against libclic_cuda.so on a GPU box.  Prints "SHIM_OK <initial> <final>".
#include <cmath>
#include <cstdio>
#include <random>

#include "../../include/clic/trajectory_estimator.hpp"

using namespace clic;

int main() {
  auto traj = std::make_shared<Trajectory>(0.1);
  std::mt19937 rng(3);
  std::normal_distribution<double> N(0.0, 1.0);
  const int K = 10;
  for (int k = 0; k < K; k++) {
    double t = 0.1 * (k - 1);
    double w[3] = {0.3 * std::sin(0.9 * t) + 0.01 * N(rng), 0.2 * std::cos(0.6 * t) + 0.01 * N(rng), 0.5 * t};
    double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), s = th > 0 ? std::sin(0.5 * th) / th : 0.5;
    traj->knots_so3.push_back({s * w[0], s * w[1], s * w[2], std::cos(0.5 * th)});
    traj->knots_pos.push_back({2 * std::sin(0.7 * t) + 0.03 * N(rng), 1.5 * std::cos(0.5 * t) + 0.03 * N(rng), 0.3 * std::sin(1.1 * t)});
  }
  TrajectoryEstimatorOptions option;
  option.lock_ab = false;
  option.lock_wb = false;
  TrajectoryEstimator::Ptr estimator(new TrajectoryEstimator(traj, option, "shim test"));
  estimator->SetFixedIndex(1);
  SO3d I;
  Vec3d zero{};
  std::uniform_real_distribution<double> U(0.0, 1.0);
  for (int i = 0; i < 400; i++) {  // loop of AddLoamMeasurementAnalytic per correspondence, trajectory_manager.cpp:682-686
    PointCorrespondence pc;
    pc.t_point = 0.001 + 0.69 * U(rng);
    pc.point = {10 * N(rng), 10 * N(rng), 3 * N(rng)};
    pc.geo_type = (i % 3 == 0) ? Line : Plane;
    double n[3] = {N(rng), N(rng), N(rng)}, nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    pc.geo_plane = {n[0] / nn, n[1] / nn, n[2] / nn, 0.3 * N(rng)};
    pc.geo_normal = {n[1] / nn, n[2] / nn, n[0] / nn};
    pc.geo_point = {N(rng), N(rng), N(rng)};
    estimator->AddLoamMeasurementAnalytic(pc, I, zero, I, zero, 5.0);
  }
  double bg0[3] = {0, 0, 0}, ba0[3] = {0, 0, 0}, bg1[3] = {0.01, 0, 0}, ba1[3] = {0, 0.02, 0}, g[3] = {0, 0, 9.8};
  Vec6d info = {800, 800, 800, 36, 36, 36};
  for (int i = 0; i < 60; i++) {
    IMUData d;
    d.timestamp = 0.01 + i * 0.011;
    d.gyro = {0.1, 0.2, 0.5};
    d.accel = {0.1, 0.2, 9.8};
    estimator->AddIMUMeasurementAnalytic(d, bg1, ba1, g, info);
  }
  Vec6d binfo = {1e4, 1e4, 1e4, 1e3, 1e3, 1e3};
  estimator->AddBiasFactor(bg0, bg1, ba0, ba1, 1, binfo);
  Summary summary = estimator->Solve(8, false);
  std::printf("%s\n", summary.BriefReport().c_str());
  bool ok = summary.final_cost < summary.initial_cost && std::isfinite(summary.final_cost);
  std::printf("%s %.6e %.6e %d\n", ok ? "SHIM_OK" : "SHIM_FAIL", summary.initial_cost, summary.final_cost,
              summary.num_successful_steps);
  return ok ? 0 : 1;
}
