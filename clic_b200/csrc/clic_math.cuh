// FP64 SO(3)/spline device math for the sm_100a kernels.
//
// Everything here is __host__ __device__ so that tests/host_emul can compile the SAME functions with g++ and
// compare them with the oracle on the CPU-only build box (test infrastructure; the product never runs them on
// the host).  Numerics follow the reference formulas (Sophus SO3 exp/log, sophus_utils Jacobians) with the same
// epsilon = 1e-10 branch thresholds, restructured for the GPU: per-knot-pair quantities are hoisted (they do
// not depend on u, so3_spline_view.h:151-165) and Jacobian chains are evaluated vector-first.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CLIC_HD __host__ __device__ __forceinline__
#else
#define CLIC_HD inline
#endif

namespace clic {

constexpr double kEps = 1e-10;  // Sophus::Constants<double>::epsilon(), src/sophus_lib/common.hpp:143-144
constexpr double kPi = 3.141592653589793238462643383279502884;

struct V3 {
  double x, y, z;
};
struct Q4 {  // (x,y,z,w), src/sophus_lib/so3.hpp:153-160
  double x, y, z, w;
};

CLIC_HD V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
CLIC_HD V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
CLIC_HD V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
CLIC_HD V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
CLIC_HD V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
CLIC_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
CLIC_HD V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
CLIC_HD V3 fma3(double s, V3 a, V3 b) { return mk(fma(s, a.x, b.x), fma(s, a.y, b.y), fma(s, a.z, b.z)); }

// Eigen quaternion product
CLIC_HD Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
CLIC_HD Q4 qconj(Q4 a) { Q4 r; r.x = -a.x; r.y = -a.y; r.z = -a.z; r.w = a.w; return r; }
// SO3::operator*= first-order renormalisation (src/sophus_lib/so3.hpp:338-355)
CLIC_HD Q4 so3_mul(Q4 a, Q4 b) {
  Q4 r = qmul(a, b);
  double sn = r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
  double s = 2.0 / (1.0 + sn);
  r.x *= s; r.y *= s; r.z *= s; r.w *= s;
  return r;
}
// q * v * q^-1 (Eigen _transformVector)
CLIC_HD V3 qrot(Q4 q, V3 v) {
  V3 qv = mk(q.x, q.y, q.z);
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
// q^-1 * v * q
CLIC_HD V3 qrot_inv(Q4 q, V3 v) { return qrot(qconj(q), v); }

// Row-major 3x3
struct M3 {
  double m[9];
};
CLIC_HD M3 qmat(Q4 q) {  // Eigen toRotationMatrix
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  M3 r;
  r.m[0] = 1.0 - (tyy + tzz); r.m[1] = txy - twz;         r.m[2] = txz + twy;
  r.m[3] = txy + twz;         r.m[4] = 1.0 - (txx + tzz); r.m[5] = tyz - twx;
  r.m[6] = txz - twy;         r.m[7] = tyz + twx;         r.m[8] = 1.0 - (txx + tyy);
  return r;
}
CLIC_HD M3 mmul(const M3& a, const M3& b) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
  return r;
}
CLIC_HD M3 mmul_t(const M3& a, const M3& b) {  // a * b^T
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++)
      r.m[i * 3 + j] = a.m[i * 3] * b.m[j * 3] + a.m[i * 3 + 1] * b.m[j * 3 + 1] + a.m[i * 3 + 2] * b.m[j * 3 + 2];
  return r;
}
CLIC_HD M3 mtrans(const M3& a) {
  M3 r;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r.m[i * 3 + j] = a.m[j * 3 + i];
  return r;
}
CLIC_HD M3 madd(const M3& a, const M3& b) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + b.m[i];
  return r; }
CLIC_HD M3 msub(const M3& a, const M3& b) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = a.m[i] - b.m[i];
  return r; }
CLIC_HD M3 mscale(double s, const M3& a) { M3 r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.m[i] = s * a.m[i];
  return r; }
CLIC_HD V3 mvec(const M3& a, V3 v) {
  return mk(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
            a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
CLIC_HD V3 vmat(V3 v, const double* m) {  // row-vector * matrix (row-major 3x3 at m)
  return mk(v.x * m[0] + v.y * m[3] + v.z * m[6], v.x * m[1] + v.y * m[4] + v.z * m[7],
            v.x * m[2] + v.y * m[5] + v.z * m[8]);
}
CLIC_HD V3 vmat_t(V3 v, const double* m) {  // row-vector * matrix^T
  return mk(v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[3] + v.y * m[4] + v.z * m[5],
            v.x * m[6] + v.y * m[7] + v.z * m[8]);
}
CLIC_HD M3 hat(V3 w) {  // src/sophus_lib/so3.hpp:618-627
  M3 r;
  r.m[0] = 0;    r.m[1] = -w.z; r.m[2] = w.y;
  r.m[3] = w.z;  r.m[4] = 0;    r.m[5] = -w.x;
  r.m[6] = -w.y; r.m[7] = w.x;  r.m[8] = 0;
  return r;
}

// sin and cos for |x| <= pi/2 (+2 %), which is all the spline chain ever asks for: x = lambda * theta / 2 with
// lambda in [0, 1] (cumulative blend weight) and theta = |log(R_k^-1 R_k+1)| <= pi.  Minimax polynomials in x^2 of degree
// 8 (sin: x * P(x^2)), fitted on [0, 1.02 (pi/2)^2] (mpmath chebyfit, fit error 2.5e-19 / 4.7e-18); measured error in
// double arithmetic <= 2.2e-16 absolute.  18 FP64 instructions, no range reduction, no slow path: the library sincos is
// ~45 plus a Payne-Hanek call site, three times per factor — 13 % of the LiDAR factor's FP64 instructions.
CLIC_HD void sincos_half_pi(double x, double* s, double* c) {
  const double z = x * x;
  double p = 2.71980544902946122e-15;
  p = fma(p, z, -7.64286117761000412e-13);
  p = fma(p, z, 1.60589346267857712e-10);
  p = fma(p, z, -2.50521067681336371e-08);
  p = fma(p, z, 2.75573192099090477e-06);
  p = fma(p, z, -1.98412698412009944e-04);
  p = fma(p, z, 8.33333333333316495e-03);
  p = fma(p, z, -1.66666666666666657e-01);
  p = fma(p, z, 1.0);
  double q = 4.60562600645516101e-14;
  q = fma(q, z, -1.14625887076970448e-11);
  q = fma(q, z, 2.08765500004449101e-09);
  q = fma(q, z, -2.75573161591531550e-07);
  q = fma(q, z, 2.48015872749139732e-05);
  q = fma(q, z, -1.38888888887584426e-03);
  q = fma(q, z, 4.16666666666634725e-02);
  q = fma(q, z, -4.99999999999999722e-01);
  q = fma(q, z, 1.0);
  *s = x * p;
  *c = q;
}
CLIC_HD void sincos_hd(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x);
  *c = cos(x);
#endif
}

// SO3::exp (src/sophus_lib/so3.hpp:534-568)
CLIC_HD Q4 so3_exp(V3 w) {
  double th2 = dot(w, w);
  double th = sqrt(th2);
  double imag, real;
  if (th < kEps) {
    double th4 = th2 * th2;
    imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
    real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
  } else {
    double s, c;
    sincos_hd(0.5 * th, &s, &c);
    imag = s / th;
    real = c;
  }
  Q4 q; q.x = imag * w.x; q.y = imag * w.y; q.z = imag * w.z; q.w = real;
  return q;
}

// SO3::log, atan based (src/sophus_lib/so3.hpp:220-263)
CLIC_HD V3 so3_log(Q4 q) {
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
  double n = sqrt(n2);
  double w = q.w;
  double f;
  if (n < kEps) {
    f = 2.0 / w - 2.0 * n2 / (w * w * w);
  } else if (fabs(w) < kEps) {
    f = (w > 0.0 ? kPi : -kPi) / n;
  } else {
    f = 2.0 * atan(n / w) / n;
  }
  return mk(f * q.x, f * q.y, f * q.z);
}

// rightJacobianInvSO3 (src/utils/sophus_utils.hpp:199-228)
CLIC_HD M3 right_jacobian_inv(V3 phi) {
  double n2 = dot(phi, phi);
  M3 K = hat(phi);
  M3 K2 = mmul(K, K);
  double c2;
  if (n2 > kEps) {
    double n = sqrt(n2);
    double s, c;
    sincos_hd(n, &s, &c);
    c2 = 1.0 / n2 - (1.0 + c) / (2.0 * n * s);
  } else {
    c2 = 1.0 / 12.0;
  }
  M3 J;
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] = 0.5 * K.m[i] + c2 * K2.m[i];
  J.m[0] += 1.0; J.m[4] += 1.0; J.m[8] += 1.0;
  return J;
}

// leftJacobianInvSO3 (src/utils/sophus_utils.hpp:279-307): the same series with -hat(phi)/2
CLIC_HD M3 left_jacobian_inv(V3 phi) {
  M3 J = right_jacobian_inv(phi);
  M3 K = hat(phi);
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] -= K.m[i];
  return J;
}

// leftJacobianSO3 (src/utils/sophus_utils.hpp:238-268)
CLIC_HD M3 left_jacobian(V3 phi) {
  double n2 = dot(phi, phi);
  M3 K = hat(phi);
  M3 K2 = mmul(K, K);
  double c1, c2;
  if (n2 > kEps) {
    double n = sqrt(n2);
    double s, c;
    sincos_hd(n, &s, &c);
    c1 = (1.0 - c) / n2;
    c2 = (n - s) / (n2 * n);
  } else {
    c1 = 0.5;
    c2 = 1.0 / 6.0;
  }
  M3 J;
#pragma unroll
  for (int i = 0; i < 9; i++) J.m[i] = c1 * K.m[i] + c2 * K2.m[i];
  J.m[0] += 1.0; J.m[4] += 1.0; J.m[8] += 1.0;
  return J;
}

// ---- uniform cubic B-spline blending (src/spline/spline_common.h:72-138), coefficients in closed form ----
// cumulative M~ = 1/6 [[6,0,0,0],[5,3,-3,1],[1,3,3,-2],[0,0,0,1]], plain M = 1/6 [[1,-3,3,-1],[4,0,-6,3],[1,3,3,-3],[0,0,0,1]]
CLIC_HD void blend_cum(double u, double lam[4]) {  // value coefficients lambda~_0..3 (lambda~_0 = 1)
  double u2 = u * u, u3 = u2 * u;
  lam[0] = 1.0;
  lam[1] = (5.0 + 3.0 * u - 3.0 * u2 + u3) * (1.0 / 6.0);
  lam[2] = (1.0 + 3.0 * u + 3.0 * u2 - 2.0 * u3) * (1.0 / 6.0);
  lam[3] = u3 * (1.0 / 6.0);
}
CLIC_HD void blend_cum_d1(double u, double inv_dt, double d[4]) {
  double u2 = u * u;
  d[0] = 0.0;
  d[1] = inv_dt * ((3.0 - 6.0 * u + 3.0 * u2) * (1.0 / 6.0));
  d[2] = inv_dt * ((3.0 + 6.0 * u - 6.0 * u2) * (1.0 / 6.0));
  d[3] = inv_dt * ((3.0 * u2) * (1.0 / 6.0));
}
CLIC_HD void blend_cum_d2(double u, double inv_dt2, double d[4]) {
  d[0] = 0.0;
  d[1] = inv_dt2 * ((-6.0 + 6.0 * u) * (1.0 / 6.0));
  d[2] = inv_dt2 * ((6.0 - 12.0 * u) * (1.0 / 6.0));
  d[3] = inv_dt2 * ((6.0 * u) * (1.0 / 6.0));
}
CLIC_HD void blend_plain(double u, double c[4]) {
  double u2 = u * u, u3 = u2 * u;
  c[0] = (1.0 - 3.0 * u + 3.0 * u2 - u3) * (1.0 / 6.0);
  c[1] = (4.0 - 6.0 * u2 + 3.0 * u3) * (1.0 / 6.0);
  c[2] = (1.0 + 3.0 * u + 3.0 * u2 - 3.0 * u3) * (1.0 / 6.0);
  c[3] = u3 * (1.0 / 6.0);
}
CLIC_HD void blend_plain_d1(double u, double s, double c[4]) {
  double u2 = u * u;
  c[0] = s * ((-3.0 + 6.0 * u - 3.0 * u2) * (1.0 / 6.0));
  c[1] = s * ((-12.0 * u + 9.0 * u2) * (1.0 / 6.0));
  c[2] = s * ((3.0 + 6.0 * u - 9.0 * u2) * (1.0 / 6.0));
  c[3] = s * ((3.0 * u2) * (1.0 / 6.0));
}
CLIC_HD void blend_plain_d2(double u, double s, double c[4]) {
  c[0] = s * ((6.0 - 6.0 * u) * (1.0 / 6.0));
  c[1] = s * ((-12.0 + 18.0 * u) * (1.0 / 6.0));
  c[2] = s * ((6.0 - 18.0 * u) * (1.0 / 6.0));
  c[3] = s * ((6.0 * u) * (1.0 / 6.0));
}
CLIC_HD void blend_plain_d3(double s, double c[4]) {
  c[0] = s * (-1.0);
  c[1] = s * (3.0);
  c[2] = s * (-3.0);
  c[3] = s * (1.0);
}

// int64-ns knot indexing, bit-exact with SplineSegmentMeta::computeTIndexNs (src/spline/spline_segment.h:67-84)
CLIC_HD bool index_ns(int64_t t_ns, int64_t t0_ns, int64_t dt_ns, int n_knots, int* s, double* u) {
  if (t_ns < t0_ns || t_ns >= t0_ns + (int64_t)(n_knots - 3) * dt_ns) return false;
  int64_t st = t_ns - t0_ns;
#if defined(__CUDA_ARCH__)
  // 64-bit integer division is emulated (~100 instructions): estimate the quotient with one FP64 multiply by the
  // (loop-invariant) reciprocal and repair it with an integer multiply; exact for st, dt < 2^52, so still bit-exact.
  const double inv_dt = 1.0 / (double)dt_ns;
  int64_t q = (int64_t)((double)st * inv_dt);
  int64_t r = st - q * dt_ns;
  if (r < 0) { q -= 1; r += dt_ns; }
  else if (r >= dt_ns) { q += 1; r -= dt_ns; }
  *s = (int)q;
  *u = double(r) / double(dt_ns);
#else
  *s = (int)(st / dt_ns);
  *u = double(st % dt_ns) / double(dt_ns);
#endif
  return true;
}

}  // namespace clic
