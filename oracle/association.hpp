// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's LiDAR data association
// (LidarOdometry::FindCorrespondence, src/lidar_odometry/lidar_odometry.cpp:483-657).  Never linked into the product.
//
// Third-party arithmetic on this path that is absent from /root/reference, restated from the published algorithms
// (parity with their float rounding is UNPINNED: no test of the reference exercises this function):
//  * pcl::KdTreeFLANN<PosPoint>::nearestKSearch (PCL 1.8 / FLANN 1.9, un-pinned): exact k-nearest search under
//    flann::L2_Simple<float> — result += diff * diff over x, y, z in float, in that order.  An exact search returns
//    the 5 smallest distances whatever the tree looks like, so a scan over all map points is the same function; ties
//    (equal float distances) are broken by the smaller map index here.
//  * Eigen::ColPivHouseholderQR<Matrix<float,5,3>>::solve (Eigen 3.3, un-pinned): Householder QR with column pivoting on
//    the largest remaining column norm, then Q^T b and back substitution.  The remaining norms are recomputed here
//    instead of LAPACK-style down-dated, which can only change the pivot order when two norms agree to rounding.
//  * cv::eigen on a 3x3 CV_32F symmetric matrix (OpenCV 3/4 JacobiImpl_, un-pinned): Jacobi rotations on the largest
//    off-diagonal element until it is below FLT_EPSILON, eigenvalues sorted descending, eigenvectors as rows.
// Compiled with -ffp-contract=off: every float / double operation below rounds once, like the device code that is
// checked against it (clic_b200/csrc/assoc.cuh uses the explicit round-to-nearest intrinsics).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <algorithm>
#include <unordered_map>
#include <vector>

namespace assoc {

struct Neighbours {
  int idx[5];
  float d2[5];
};

// flann::L2_Simple<float> (dist.h): sequential float accumulate of squared differences.
inline float sq_dist(const float* a, const float* b) {
  float r = 0.f;
  for (int c = 0; c < 3; c++) {
    const float diff = a[c] - b[c];
    r += diff * diff;
  }
  return r;
}

inline Neighbours knn5(const float* map, int64_t n_map, const float* q) {
  Neighbours nb;
  for (int k = 0; k < 5; k++) {
    nb.idx[k] = INT32_MAX;
    nb.d2[k] = INFINITY;
  }
  for (int64_t j = 0; j < n_map; j++) {
    const float d = sq_dist(q, map + 3 * j);
    if (!(d < nb.d2[4])) continue;
    int k = 4;
    while (k > 0 && d < nb.d2[k - 1]) {  // strict: an equal distance stays behind the smaller index
      nb.d2[k] = nb.d2[k - 1];
      nb.idx[k] = nb.idx[k - 1];
      k--;
    }
    nb.d2[k] = d;
    nb.idx[k] = (int)j;
  }
  return nb;
}

// The same exact search through a uniform 1 m grid, for the timed CPU baseline only (the reference walks a kd-tree;
// an exhaustive scan would flatter the GPU).  Only neighbours inside the 1 m^2 gate can change the outcome, and those
// lie in the cells within reach of the query, so the kept records are identical; neighbours outside the gate come back
// as (+inf, INT32_MAX).  tests/test_association.py checks it against knn5.
struct Grid {
  std::unordered_map<uint64_t, std::vector<int>> cells;
  static uint64_t key(int64_t cx, int64_t cy, int64_t cz) {
    return ((uint64_t)(cx & 0x1fffff) << 42) | ((uint64_t)(cy & 0x1fffff) << 21) | (uint64_t)(cz & 0x1fffff);
  }
  void build(const float* map, int64_t n) {
    cells.clear();
    for (int64_t j = 0; j < n; j++) {
      const float* p = map + 3 * j;
      if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
      cells[key((int64_t)std::floor(p[0]), (int64_t)std::floor(p[1]), (int64_t)std::floor(p[2]))].push_back((int)j);
    }
  }
  Neighbours knn5(const float* map, const float* q) const {
    Neighbours nb;
    for (int k = 0; k < 5; k++) {
      nb.idx[k] = INT32_MAX;
      nb.d2[k] = INFINITY;
    }
    if (!(std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))) return nb;
    const float amax = std::max(std::fabs(q[0]), std::max(std::fabs(q[1]), std::fabs(q[2])));
    const float r = 1.0001f + 4e-7f * amax;  // reach of the gate plus the rounding of a float difference
    int64_t lo[3], hi[3];
    for (int c = 0; c < 3; c++) {
      lo[c] = (int64_t)std::floor(q[c] - r);
      hi[c] = (int64_t)std::floor(q[c] + r);
    }
    for (int64_t cx = lo[0]; cx <= hi[0]; cx++)
      for (int64_t cy = lo[1]; cy <= hi[1]; cy++)
        for (int64_t cz = lo[2]; cz <= hi[2]; cz++) {
          auto it = cells.find(key(cx, cy, cz));
          if (it == cells.end()) continue;
          for (int j : it->second) {
            const float d = sq_dist(q, map + 3 * (size_t)j);
            if (!(d <= 1.0f)) continue;
            if (!(d < nb.d2[4] || (d == nb.d2[4] && j < nb.idx[4]))) continue;
            int k = 4;
            while (k > 0 && (d < nb.d2[k - 1] || (d == nb.d2[k - 1] && j < nb.idx[k - 1]))) {
              nb.d2[k] = nb.d2[k - 1];
              nb.idx[k] = nb.idx[k - 1];
              k--;
            }
            nb.d2[k] = d;
            nb.idx[k] = j;
          }
        }
    return nb;
  }
};

inline float hyp(float a, float b) { return (float)std::sqrt((double)a * (double)a + (double)b * (double)b); }

// lidar_odometry.cpp:511-563.  Returns true for a valid line; cp = centroid, dir = principal direction.
inline bool fit_line(const float pts[5][3], double cp[3], double dir[3]) {
  cp[0] = cp[1] = cp[2] = 0;
  for (int j = 0; j < 5; j++)
    for (int c = 0; c < 3; c++) cp[c] += (double)pts[j][c];
  for (int c = 0; c < 3; c++) cp[c] /= 5.0;
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
  for (int j = 0; j < 5; j++) {
    const float ax = (float)((double)pts[j][0] - cp[0]);
    const float ay = (float)((double)pts[j][1] - cp[1]);
    const float az = (float)((double)pts[j][2] - cp[2]);
    a11 += ax * ax; a12 += ax * ay; a13 += ax * az;
    a22 += ay * ay; a23 += ay * az; a33 += az * az;
  }
  a11 /= 5; a12 /= 5; a13 /= 5; a22 /= 5; a23 /= 5; a33 /= 5;
  // cv::eigen: Jacobi on the upper triangle, W = diagonal, V rows = eigenvectors.
  float A[3][3] = {{a11, a12, a13}, {a12, a22, a23}, {a13, a23, a33}};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  float W[3] = {a11, a22, a33};
  for (int it = 0; it < 270; it++) {
    int k = 0, l = 1;
    float mv = std::fabs(A[0][1]);
    if (mv < std::fabs(A[0][2])) { mv = std::fabs(A[0][2]); k = 0; l = 2; }
    if (mv < std::fabs(A[1][2])) { mv = std::fabs(A[1][2]); k = 1; l = 2; }
    const float p = A[k][l];
    if (std::fabs(p) <= FLT_EPSILON) break;
    const float y = (W[l] - W[k]) * 0.5f;
    float t = std::fabs(y) + hyp(p, y);
    float s = hyp(p, t);
    const float c = t / s;
    s = p / s;
    t = (p / t) * p;
    if (y < 0) { s = -s; t = -t; }
    A[k][l] = 0;
    W[k] -= t;
    W[l] += t;
    auto rot = [&](float& v0, float& v1) {
      const float a0 = v0, b0 = v1;
      v0 = a0 * c - b0 * s;
      v1 = a0 * s + b0 * c;
    };
    for (int i = 0; i < k; i++) rot(A[i][k], A[i][l]);
    for (int i = k + 1; i < l; i++) rot(A[k][i], A[i][l]);
    for (int i = l + 1; i < 3; i++) rot(A[k][i], A[l][i]);
    for (int i = 0; i < 3; i++) rot(V[k][i], V[l][i]);
  }
  // sort descending (selection, as OpenCV does after the sweeps)
  for (int k = 0; k < 2; k++) {
    int m = k;
    for (int i = k + 1; i < 3; i++)
      if (W[m] < W[i]) m = i;
    if (m != k) {
      std::swap(W[m], W[k]);
      for (int i = 0; i < 3; i++) std::swap(V[m][i], V[k][i]);
    }
  }
  if (!(W[0] > 3 * W[1])) return false;
  for (int c = 0; c < 3; c++) dir[c] = (double)V[0][c];
  return true;
}

// lidar_odometry.cpp:586-625.  Returns true for a valid plane; plane = (pa, pb, pc, pd) as floats widened.
inline bool fit_plane(const float pts[5][3], double plane[4]) {
  float Q[5][3], cvec[5];
  int perm[3] = {0, 1, 2};
  for (int j = 0; j < 5; j++) {
    for (int c = 0; c < 3; c++) Q[j][c] = pts[j][c];
    cvec[j] = -1.f;
  }
  float tau[3] = {0, 0, 0};
  float maxn2 = 0;
  for (int c = 0; c < 3; c++) {
    float n2 = 0;
    for (int j = 0; j < 5; j++) n2 += Q[j][c] * Q[j][c];
    if (n2 > maxn2) maxn2 = n2;
  }
  const float maxn = std::sqrt(maxn2);
  const float thr = (maxn * FLT_EPSILON) * (maxn * FLT_EPSILON) / 5.f;
  int nz = 3;
  for (int k = 0; k < 3; k++) {
    int big = k;
    float bign2 = -1.f;
    for (int c = k; c < 3; c++) {
      float n2 = 0;
      for (int j = k; j < 5; j++) n2 += Q[j][c] * Q[j][c];
      if (n2 > bign2) { bign2 = n2; big = c; }
    }
    if (nz == 3 && bign2 < thr * (float)(5 - k)) nz = k;
    if (big != k) {
      for (int j = 0; j < 5; j++) std::swap(Q[j][k], Q[j][big]);
      std::swap(perm[k], perm[big]);
    }
    float tail2 = 0;
    for (int j = k + 1; j < 5; j++) tail2 += Q[j][k] * Q[j][k];
    const float c0 = Q[k][k];
    float beta;
    if (tail2 <= FLT_MIN) {
      tau[k] = 0;
      beta = c0;
      for (int j = k + 1; j < 5; j++) Q[j][k] = 0;
    } else {
      beta = std::sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      const float den = c0 - beta;
      for (int j = k + 1; j < 5; j++) Q[j][k] /= den;
      tau[k] = (beta - c0) / beta;
    }
    Q[k][k] = beta;
    for (int c = k + 1; c < 3; c++) {
      float tmp = 0;
      for (int j = k + 1; j < 5; j++) tmp += Q[j][k] * Q[j][c];
      tmp += Q[k][c];
      Q[k][c] -= tau[k] * tmp;
      for (int j = k + 1; j < 5; j++) Q[j][c] -= (tau[k] * Q[j][k]) * tmp;
    }
  }
  for (int k = 0; k < nz; k++) {  // c = Q^T b
    float tmp = 0;
    for (int j = k + 1; j < 5; j++) tmp += Q[j][k] * cvec[j];
    tmp += cvec[k];
    cvec[k] -= tau[k] * tmp;
    for (int j = k + 1; j < 5; j++) cvec[j] -= (tau[k] * Q[j][k]) * tmp;
  }
  float xs[3] = {0, 0, 0}, x[3] = {0, 0, 0};
  for (int i = nz - 1; i >= 0; i--) {
    float acc = cvec[i];
    for (int c = i + 1; c < nz; c++) acc -= Q[i][c] * xs[c];
    xs[i] = acc / Q[i][i];
  }
  for (int i = 0; i < nz; i++) x[perm[i]] = xs[i];
  float pa = x[0], pb = x[1], pc = x[2], pd = 1;
  const float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
  pa /= ps; pb /= ps; pc /= ps; pd /= ps;
  for (int j = 0; j < 5; j++) {
    const float v = pa * pts[j][0] + pb * pts[j][1] + pc * pts[j][2] + pd;
    if (!((double)std::fabs(v) <= 0.2)) return false;  // NaN (degenerate fit) is rejected too
  }
  plane[0] = pa; plane[1] = pb; plane[2] = pc; plane[3] = pd;
  return true;
}

}  // namespace assoc
