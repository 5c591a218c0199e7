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
  // The same inner iteration's LiDAR factors straight from a scan: a floor, a wall and their edge as the feature map;
  // the scan is the map's surfaces seen from the trajectory (raw points p_L = pose(t)^-1 p_G, pose from the library).
  {
    PosCloud map_surf, map_corner, scan_surf, scan_corner;
    std::vector<Vec3d> surf_G, corner_G;
    std::vector<double> ts, tc;
    for (int i = 0; i < 6000; i++) {
      PosPoint p;
      if (i % 2) { p.x = 8 * U(rng) - 2; p.y = 8 * U(rng) - 2; p.z = -1.5f + 0.01 * N(rng); }   // floor z = -1.5
      else { p.x = 5.0f + 0.01 * N(rng); p.y = 8 * U(rng) - 2; p.z = 4 * U(rng) - 1.5; }        // wall x = 5
      map_surf.push_back(p);
    }
    for (int i = 0; i < 600; i++) {
      PosPoint p;
      p.x = 5.0f + 0.01 * N(rng); p.y = 8 * U(rng) - 2; p.z = -1.5f + 0.01 * N(rng);             // the edge
      map_corner.push_back(p);
    }
    for (int i = 0; i < 300; i++) {
      surf_G.push_back(i % 2 ? Vec3d{8 * U(rng) - 2, 8 * U(rng) - 2, -1.5} : Vec3d{5.0, 8 * U(rng) - 2, 4 * U(rng) - 1.5});
      ts.push_back(0.001 + 0.69 * U(rng));
    }
    for (int i = 0; i < 80; i++) {
      corner_G.push_back({5.0, 7 * U(rng) - 1.5, -1.5});
      tc.push_back(0.001 + 0.69 * U(rng));
    }
    TrajectoryEstimator::Ptr probe(new TrajectoryEstimator(traj, option, "pose probe"));
    auto to_scan = [&](const std::vector<Vec3d>& pg, const std::vector<double>& t, PosCloud* out) {
      std::vector<SO3d> q;
      std::vector<Vec3d> p;
      probe->GetLidarPoses(t, &q, &p);
      for (size_t i = 0; i < pg.size(); i++) {  // p_L = R^T (p_G - p), R from the unit quaternion (x, y, z, w)
        const double* c = q[i].data();
        const double x = c[0], y = c[1], z = c[2], w = c[3];
        const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                                {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                                {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
        const double d[3] = {pg[i][0] - p[i][0], pg[i][1] - p[i][1], pg[i][2] - p[i][2]};
        PosPoint o;
        o.x = (float)(R[0][0] * d[0] + R[1][0] * d[1] + R[2][0] * d[2]);
        o.y = (float)(R[0][1] * d[0] + R[1][1] * d[1] + R[2][1] * d[2]);
        o.z = (float)(R[0][2] * d[0] + R[1][2] * d[1] + R[2][2] * d[2]);
        o.timestamp = t[i];
        out->push_back(o);
      }
    };
    to_scan(surf_G, ts, &scan_surf);
    to_scan(corner_G, tc, &scan_corner);
    LidarFeatureMap::Ptr map(new LidarFeatureMap(map_surf, map_corner));
    estimator->AddLoamMeasurementsFromScan(map, scan_corner, scan_surf, I, zero, I, zero, 5.0, 2);
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
  // InitTrajWithPropagation's extra factor (trajectory_manager.cpp:279-288)
  bool vel_in = estimator->AddGlobalVelocityMeasurement(0.3, {1.2, 0.4, 0.1}, 10.0);
  bool vel_out = estimator->AddGlobalVelocityMeasurement(7.0, {1.2, 0.4, 0.1}, 10.0);
  Vec6d binfo = {1e4, 1e4, 1e4, 1e3, 1e3, 1e3};
  estimator->AddBiasFactor(bg0, bg1, ba0, ba1, 1, binfo);
  Summary summary = estimator->Solve(8, false);
  std::printf("%s\n", summary.BriefReport().c_str());
  const auto& sc = estimator->ScanCorrespondenceCounts();
  std::printf("scan correspondences: %lld lines, %lld planes\n", (long long)sc.at(0).first, (long long)sc.at(0).second);
  bool ok = summary.final_cost < summary.initial_cost && std::isfinite(summary.final_cost) && sc.at(0).first > 10 &&
            sc.at(0).second > 60 && vel_in && !vel_out;
  std::printf("%s %.6e %.6e %d\n", ok ? "SHIM_OK" : "SHIM_FAIL", summary.initial_cost, summary.final_cost,
              summary.num_successful_steps);
  return ok ? 0 : 1;
}
