// Small dense kernels around the factor pass: the marginalization prior (K4), and — further down — the
// LM step (K6: Jacobi scaling, damping, dense FP64 Cholesky, triangular solves, model decrease), the
// manifold update (K7) and the accept/reject bookkeeping.  All are single-CTA kernels: D <= ~300 here, the work
// is O(D^3) <= 3e7 flop and latency-bound.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace clic {

// A = J0^T J0 (n x n), once per prior at add time.
__global__ void prior_gram_kernel(const double* J0, int n, double* A) {
  int i = blockIdx.y * blockDim.y + threadIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j >= n) return;
  double acc = 0.0;
  for (int k = 0; k < n; k++) acc += J0[(size_t)k * n + i] * J0[(size_t)k * n + j];
  A[(size_t)i * n + j] = acc;
}
__global__ void prior_grad_kernel(const double* J0, const double* r0, int n, double* g) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int k = 0; k < n; k++) acc += J0[(size_t)k * n + i] * r0[k];
  g[i] = acc;
}

// MarginalizationFactor::Evaluate (marginalization_factor.cpp:331-372) fused with its contribution to H, b, cost:
//   dx = x - x0 (vectors) | 2 vec(q0^-1 q) sign-fixed to w >= 0 (quaternions);  r = r0 + J0 dx;
//   H += J0^T J0 (precomputed A), b += J0^T r, cost += 1/2 |r|^2.
// meta per block: [state offset, size, tangent idx, column (-1 constant)].
__global__ void prior_kernel(const double* J0, const double* A, const double* r0, const double* x0, const int* meta,
                             int n, int n_blocks, const double* state, double* H, double* b, int D, double* cost2,
                             int jac, const int* ctl) {
  if (ctl && ctl[0]) return;
  extern __shared__ double sm[];
  double* dx = sm;
  double* r = dx + n;
  double* g = r + n;
  int* tcol = reinterpret_cast<int*>(g + n);
  __shared__ int any_free;
  if (threadIdx.x == 0) any_free = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n_blocks; i += blockDim.x) {
    const int off = meta[4 * i], size = meta[4 * i + 1], idx = meta[4 * i + 2], col = meta[4 * i + 3];
    const double* x = state + off;
    const double* xr = x0 + 4 * i;
    if (size != 4) {
      for (int k = 0; k < size; k++) {
        dx[idx + k] = x[k] - xr[k];
        tcol[idx + k] = col >= 0 ? col + k : -1;
      }
    } else {
      Q4 q0 = qconj(ldq(xr));
      double n2 = q0.x * q0.x + q0.y * q0.y + q0.z * q0.z + q0.w * q0.w;  // Eigen inverse() = conjugate / squaredNorm
      q0.x /= n2; q0.y /= n2; q0.z /= n2; q0.w /= n2;
      Q4 d = qmul(q0, ldq(x));
      double sgn = d.w >= 0 ? 1.0 : -1.0;
      dx[idx] = 2.0 * (sgn * d.x);
      dx[idx + 1] = 2.0 * (sgn * d.y);
      dx[idx + 2] = 2.0 * (sgn * d.z);
      for (int k = 0; k < 3; k++) tcol[idx + k] = col >= 0 ? col + k : -1;
    }
    if (col >= 0) any_free = 1;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double acc = 0.0;
    for (int c = 0; c < n; c++) acc += J0[(size_t)t * n + c] * dx[c];
    r[t] = r0[t] + acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double c = 0.0;
    for (int t = 0; t < n; t++) c += r[t] * r[t];
    cost2[any_free ? 0 : 1] += 0.5 * c;
  }
  if (!jac) return;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < n; k++) acc += J0[(size_t)k * n + t] * r[k];
    g[t] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int a = e / n, c = e % n;
    const int ca = tcol[a], cc = tcol[c];
    if (ca >= 0 && cc >= 0) H[(size_t)ca * D + cc] += A[(size_t)a * n + c];
  }
  for (int a = threadIdx.x; a < n; a += blockDim.x)
    if (tcol[a] >= 0) b[tcol[a]] += g[a];
}

}  // namespace clic
