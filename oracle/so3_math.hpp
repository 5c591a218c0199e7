// TEST INFRASTRUCTURE ONLY — CPU restatement of the reference arithmetic (oracle).
// Nothing under oracle/ is linked into, imported by or called from the product (clic_b200/).
//
// Small fixed-size FP64 helpers + the SO(3) numerics the reference takes from Eigen / vendored Sophus.
// Each function cites the reference code whose formula (and operation order) it follows.
// Parity status: the reference cannot be compiled in this container (needs Eigen/Ceres), so these
// restatements are pinned by (a) the reference test's own property, analytic == auto-diff Jacobian
// to 1e-6 on its hard-coded inputs (src/app/test_analytic_factor.cpp:846-900), checked in
// tests/test_oracle_autodiff.py against an independent torch-f64 lineage, and (b) committed golden vectors.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

constexpr double kEps = 1e-10;  // Sophus::Constants<double>::epsilon(), src/sophus_lib/common.hpp:143-144
constexpr double kPi = 3.141592653589793238462643383279502884;

struct Vec3 {
  double v[3];
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
};
struct Mat3 {
  double m[3][3];
  double& operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
};
// Quaternion in Eigen coefficient order (x,y,z,w) — src/sophus_lib/so3.hpp:153-160.
struct Quat {
  double x, y, z, w;
};

inline Vec3 v3(double a, double b, double c) { return Vec3{{a, b, c}}; }
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return v3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return v3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 operator-(const Vec3& a) { return v3(-a[0], -a[1], -a[2]); }
inline Vec3 operator*(double s, const Vec3& a) { return v3(s * a[0], s * a[1], s * a[2]); }
inline Vec3 operator*(const Vec3& a, double s) { return v3(a[0] * s, a[1] * s, a[2] * s); }
inline Vec3 operator/(const Vec3& a, double s) { return v3(a[0] / s, a[1] / s, a[2] / s); }
inline double dot(const Vec3& a, const Vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline double sqnorm(const Vec3& a) { return dot(a, a); }
inline double norm(const Vec3& a) { return std::sqrt(sqnorm(a)); }

inline Mat3 zero3() {
  Mat3 r;
  std::memset(&r, 0, sizeof(r));
  return r;
}
inline Mat3 eye3() {
  Mat3 r = zero3();
  r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
  return r;
}
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, 0) * b(0, j) + a(i, 1) * b(1, j) + a(i, 2) * b(2, j);
  return r;
}
inline Vec3 operator*(const Mat3& a, const Vec3& b) {
  return v3(a(0, 0) * b[0] + a(0, 1) * b[1] + a(0, 2) * b[2], a(1, 0) * b[0] + a(1, 1) * b[1] + a(1, 2) * b[2],
            a(2, 0) * b[0] + a(2, 1) * b[1] + a(2, 2) * b[2]);
}
inline Mat3 operator*(double s, const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = s * a(i, j);
  return r;
}
inline Mat3 operator*(const Mat3& a, double s) { return s * a; }
inline Mat3 operator/(const Mat3& a, double s) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, j) / s;
  return r;
}
inline Mat3 operator+(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, j) + b(i, j);
  return r;
}
inline Mat3 operator-(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(i, j) - b(i, j);
  return r;
}
inline Mat3 operator-(const Mat3& a) { return -1.0 * a; }
inline Mat3 transpose(const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r(i, j) = a(j, i);
  return r;
}
// row-vector * matrix
inline Vec3 rowmul(const Vec3& a, const Mat3& m) {
  return v3(a[0] * m(0, 0) + a[1] * m(1, 0) + a[2] * m(2, 0), a[0] * m(0, 1) + a[1] * m(1, 1) + a[2] * m(2, 1),
            a[0] * m(0, 2) + a[1] * m(1, 2) + a[2] * m(2, 2));
}

// SO3::hat — src/sophus_lib/so3.hpp:618-627
inline Mat3 hat(const Vec3& w) {
  Mat3 r;
  r(0, 0) = 0;     r(0, 1) = -w[2]; r(0, 2) = w[1];
  r(1, 0) = w[2];  r(1, 1) = 0;     r(1, 2) = -w[0];
  r(2, 0) = -w[1]; r(2, 1) = w[0];  r(2, 2) = 0;
  return r;
}
// SO3::vee — (Omega(2,1), Omega(0,2), Omega(1,0))
inline Vec3 vee(const Mat3& m) { return v3(m(2, 1), m(0, 2), m(1, 0)); }

// Eigen quaternion product (Eigen/src/Geometry/Quaternion.h, un-vendored Eigen 3).
inline Quat qmul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// SO3::inverse = conjugate — src/sophus_lib/so3.hpp:200-202
inline Quat qconj(const Quat& a) { return Quat{-a.x, -a.y, -a.z, a.w}; }
inline double qsqnorm(const Quat& a) { return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w; }

// SO3::operator*= with first-order renormalisation — src/sophus_lib/so3.hpp:338-355.
// SO3::operator* is a copy + operator*= (":300-304"), so every group product renormalises.
inline Quat so3_mul(const Quat& a, const Quat& b) {
  Quat r = qmul(a, b);
  double sn = qsqnorm(r);
  if (sn != 1.0) {
    double s = 2.0 / (1.0 + sn);
    r.x *= s; r.y *= s; r.z *= s; r.w *= s;
  }
  return r;
}

// Quaternion -> rotation matrix: Eigen QuaternionBase::toRotationMatrix (used by SO3::matrix(), so3.hpp:282-284).
inline Mat3 so3_matrix(const Quat& q) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r(0, 0) = 1.0 - (tyy + tzz); r(0, 1) = txy - twz;         r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;         r(1, 1) = 1.0 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;         r(2, 1) = tyz + twx;         r(2, 2) = 1.0 - (txx + tyy);
  return r;
}

// SO3 * point: Eigen QuaternionBase::_transformVector (so3.hpp:320-322).
inline Vec3 so3_rotate(const Quat& q, const Vec3& v) {
  Vec3 qv = v3(q.x, q.y, q.z);
  Vec3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}

// SO3::exp — src/sophus_lib/so3.hpp:534-568
inline Quat so3_exp(const Vec3& omega) {
  double theta_sq = sqnorm(omega);
  double theta = std::sqrt(theta_sq);
  double half_theta = 0.5 * theta;
  double imag_factor, real_factor;
  if (theta < kEps) {
    double theta_po4 = theta_sq * theta_sq;
    imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
    real_factor = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
  } else {
    double sin_half_theta = std::sin(half_theta);
    imag_factor = sin_half_theta / theta;
    real_factor = std::cos(half_theta);
  }
  return Quat{imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2], real_factor};
}

// SO3::log (atan based) — src/sophus_lib/so3.hpp:220-263
inline Vec3 so3_log(const Quat& q) {
  double squared_n = q.x * q.x + q.y * q.y + q.z * q.z;
  double n = std::sqrt(squared_n);
  double w = q.w;
  double two_atan_nbyw_by_n;
  if (n < kEps) {
    double squared_w = w * w;
    two_atan_nbyw_by_n = 2.0 / w - 2.0 * (squared_n) / (w * squared_w);
  } else {
    if (std::fabs(w) < kEps) {
      if (w > 0.0) {
        two_atan_nbyw_by_n = kPi / n;
      } else {
        two_atan_nbyw_by_n = -kPi / n;
      }
    } else {
      two_atan_nbyw_by_n = 2.0 * std::atan(n / w) / n;
    }
  }
  return v3(two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z);
}

// rightJacobianSO3 — src/utils/sophus_utils.hpp:160-188
inline Mat3 right_jacobian_so3(const Vec3& phi) {
  double phi_norm2 = sqnorm(phi);
  Mat3 phi_hat = hat(phi);
  Mat3 phi_hat2 = phi_hat * phi_hat;
  Mat3 J = eye3();
  if (phi_norm2 > kEps) {
    double phi_norm = std::sqrt(phi_norm2);
    double phi_norm3 = phi_norm2 * phi_norm;
    J = J - phi_hat * (1 - std::cos(phi_norm)) / phi_norm2;
    J = J + phi_hat2 * (phi_norm - std::sin(phi_norm)) / phi_norm3;
  } else {
    J = J - phi_hat / 2;
    J = J + phi_hat2 / 6;
  }
  return J;
}

// rightJacobianInvSO3 — src/utils/sophus_utils.hpp:199-228
inline Mat3 right_jacobian_inv_so3(const Vec3& phi) {
  double phi_norm2 = sqnorm(phi);
  Mat3 phi_hat = hat(phi);
  Mat3 phi_hat2 = phi_hat * phi_hat;
  Mat3 J = eye3();
  J = J + phi_hat / 2;
  if (phi_norm2 > kEps) {
    double phi_norm = std::sqrt(phi_norm2);
    J = J + phi_hat2 * (1 / phi_norm2 - (1 + std::cos(phi_norm)) / (2 * phi_norm * std::sin(phi_norm)));
  } else {
    J = J + phi_hat2 / 12;
  }
  return J;
}

// leftJacobianInvSO3 — src/utils/sophus_utils.hpp:279-307
inline Mat3 left_jacobian_inv_so3(const Vec3& phi) {
  double phi_norm2 = sqnorm(phi);
  Mat3 phi_hat = hat(phi);
  Mat3 phi_hat2 = phi_hat * phi_hat;
  Mat3 J = eye3();
  J = J - phi_hat / 2;
  if (phi_norm2 > kEps) {
    double phi_norm = std::sqrt(phi_norm2);
    J = J + phi_hat2 * (1 / phi_norm2 - (1 + std::cos(phi_norm)) / (2 * phi_norm * std::sin(phi_norm)));
  } else {
    J = J + phi_hat2 / 12;
  }
  return J;
}

// Eigen::Quaterniond(Matrix3d) — Eigen is not vendored in the reference; this is its published conversion
// (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl<Other, 3, 3>, Shoemake's algorithm), applied as written
// to WHATEVER 3x3 matrix it is given (the reference feeds it a skew matrix in IMUPoseFactor's time-offset column).
inline Quat quat_from_matrix_eigen(const Mat3& m) {
  double q[4];  // x y z w
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m(2, 1) - m(1, 2)) * t;
    q[1] = (m(0, 2) - m(2, 0)) * t;
    q[2] = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m(k, j) - m(j, k)) * t;
    q[j] = (m(j, i) + m(i, j)) * t;
    q[k] = (m(k, i) + m(i, k)) * t;
  }
  return Quat{q[0], q[1], q[2], q[3]};
}
inline Quat qnormalized(const Quat& a) {
  double n = std::sqrt(qsqnorm(a));
  return Quat{a.x / n, a.y / n, a.z / n, a.w / n};
}

}  // namespace oracle
