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
  pdl_wait();
  pdl_trigger();
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
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  for (int t = warp; t < n; t += n_warps) {  // one warp per row of J0: coalesced loads, fixed-order butterfly sum
    double acc = 0.0;
    for (int c = lane; c < n; c += 32) acc += J0[(size_t)t * n + c] * dx[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) r[t] = r0[t] + acc;
  }
  __syncthreads();
  if (warp == 0) {
    double c = 0.0;
    for (int t = lane; t < n; t += 32) c += r[t] * r[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cost2[any_free ? 0 : 1] += 0.5 * c;
  }
  if (!jac) return;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {  // column t of J0: coalesced across threads
    double acc = 0.0;
#pragma unroll 8
    for (int k = 0; k < n; k++) acc += J0[(size_t)k * n + t] * r[k];
    g[t] = acc;
  }
  __syncthreads();
  for (int a = warp; a < n; a += n_warps) {
    const int ca = tcol[a];
    if (ca < 0) continue;
    for (int c = lane; c < n; c += 32) {
      const int cc = tcol[c];
      if (cc >= 0) H[(size_t)ca * D + cc] += A[(size_t)a * n + c];
    }
  }
  for (int a = threadIdx.x; a < n; a += blockDim.x)
    if (tcol[a] >= 0) b[tcol[a]] += g[a];
}

}  // namespace clic

namespace clic {

// ---------------- dense FP64 Cholesky / triangular solves, one CTA, blocked by 8 ----------------
// Right-looking, panel width kNB: (1) the kNB x kNB diagonal block is factored by warp 0 (warp-synchronous),
// (2) the panel below it is solved row-parallel, (3) the trailing lower triangle is updated entry-parallel.
// Three block-wide barriers per 8 columns instead of three per column.  A is n x n row-major (leading dimension ld),
// in shared memory when it fits, else in L2-resident global scratch; only the lower triangle is referenced.
constexpr int kNB = 8;

#ifdef CLIC_LM_TIMING  // developer build: phase time stamps of the last lm_step_kernel launch (bench_micro/lm_timeline.py)
__device__ long long g_lm_stamps[16];
#define LM_STAMP(i) do { if (threadIdx.x == 0) g_lm_stamps[i] = clock64(); } while (0)
#define LM_ACC(i, t0) do { if (threadIdx.x == 0) g_lm_stamps[i] += clock64() - (t0); } while (0)
#define LM_NOW() clock64()
#else
#define LM_STAMP(i) do { } while (0)
#define LM_ACC(i, t0) do { } while (0)
#define LM_NOW() 0
#endif

// Warp 0, lane l < nb holds row l of the kNB x kNB diagonal block in registers; columns are eliminated with warp
// shuffles (no shared-memory round trips on the serial critical path).  Writes L back and, if Li != nullptr, the
// inverse of the block (row-major kNB x kNB, zero padded) used by the panel and triangular solves.
__device__ __forceinline__ void warp_diag_block(double* A, int ld, int kb, int nb, double* dinv, int* ok, int lane,
                                                const double* dtol = nullptr) {
  // This is the serial critical path of the whole factorization.  Measured on B200 (bench_micro/lm_timeline.py): a
  // warp-parallel elimination (one row per lane, columns exchanged with shuffles) costs ~270 cycles per column however
  // its dependent chain is arranged — one warp alone issues an instruction every 3-4 cycles and the ~36 double
  // shuffles + selects per block are what it spends them on.  ONE thread holding the whole 8 x 8 block in registers
  // needs no exchange at all: ~330 instructions per block, independent updates fill the latency of the pivot chain.
  // Chain per column: 1/d (hardware reciprocal approximation, ~2^-20, + two Newton steps of two FMAs) -> t_i = a_ij / d
  // -> a_ic -= t_i a_cj.  Off the chain: 1/sqrt(d) for the stored column (MUFU.RSQ64H + two Newton steps; the pivot is
  // d * (1/sqrt(d))), and the pivot test, whose failure is only reported — the block then holds garbage and every
  // caller stops on !*ok.  The library rsqrt() / sqrt() + division put several times more dependent instructions here.
  long long tq0 = LM_NOW();
  if (lane != 0) return;
  double a[kNB][kNB];
#pragma unroll
  for (int i = 0; i < kNB; i++)
#pragma unroll
    for (int c = 0; c <= i; c++)  // rows / columns past a partial block are padded with the identity
      a[i][c] = i < nb ? A[(size_t)(kb + i) * ld + kb + c] : (c == i ? 1.0 : 0.0);
  LM_ACC(13, tq0);
  tq0 = LM_NOW();
  bool any_bad = false;
#pragma unroll
  for (int j = 0; j < kNB; j++) {
    const double d = a[j][j];
    double rd;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rd) : "d"(d));
    rd = fma(rd, fma(-d, rd, 1.0), rd);
    rd = fma(rd, fma(-d, rd, 1.0), rd);
#pragma unroll
    for (int i = j + 1; i < kNB; i++) {
      const double t = a[i][j] * rd;
#pragma unroll
      for (int c = j + 1; c <= i; c++) a[i][c] = fma(-t, a[c][j], a[i][c]);
    }
    any_bad |= (j < nb) && (!(d > (dtol ? dtol[kb + j] : 0.0)) || !isfinite(d));
    double inv;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(d));
    const double hd = 0.5 * d;
    inv = inv * fma(-hd, inv * inv, 1.5);
    inv = inv * fma(-hd, inv * inv, 1.5);
    if (j < nb) dinv[kb + j] = inv;
#pragma unroll
    for (int i = j; i < kNB; i++) a[i][j] *= inv;
  }
  LM_ACC(14, tq0);
  tq0 = LM_NOW();
  if (any_bad) *ok = 0;
#pragma unroll
  for (int i = 0; i < kNB; i++)
#pragma unroll
    for (int c = 0; c <= i; c++)
      if (i < nb) A[(size_t)(kb + i) * ld + kb + c] = a[i][c];
  LM_ACC(15, tq0);
}

// Inverse of the kNB x kNB diagonal block kb of L (row-major kNB x kNB, zero padded), by one warp.
__device__ __forceinline__ void warp_block_inverse(const double* A, int ld, int kb, int nb, const double* dinv, double* Li,
                                                   int lane) {
  // lane l computes column l of L^-1 by forward substitution on e_l
  double x[kNB];
#pragma unroll
  for (int i = 0; i < kNB; i++) {
    double acc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < i; k++)
      if (k >= lane && i < nb) acc -= A[(size_t)(kb + i) * ld + kb + k] * x[k];
    x[i] = (i >= lane && i < nb && lane < nb) ? acc * dinv[kb + i] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < kNB; i++)
    if (lane < kNB) Li[i * kNB + lane] = x[i];
}

// dinv: n doubles receiving 1 / L[j][j].  Linv: nullptr or ceil(n/kNB) * 64 doubles receiving the inverses of the
// diagonal blocks (computed after the factorization, one warp per block, off the serial path).
// dtol[j]: pivot j <= dtol[j] counts as "not positive definite" (nullptr = 0 for the damped LM system; in K8
// n * eps * the column's original diagonal, i.e. the column is numerically dependent on the previous ones).
// pan: nullptr, or (n + kNB) * kPanStride doubles of SHARED scratch used when A itself lives in global memory: the
// factored diagonal block and the panel are staged there, so the trailing update reads global memory once per entry.
constexpr int kPanStride = kNB + 1;  // odd: conflict-free when lanes stride the rows
// n_rows >= n (0 = n): rows n .. n_rows-1 are carried along as extra panel rows (L[i][0..n) = A[i][0..n) L^-T) and
// their trailing entries A[i][j], j < n, are updated — a right-hand side appended as row n becomes L^-1 rhs.
__device__ __forceinline__ void cta_cholesky(double* A, int ld, int n, int* ok, double* dinv, double* Linv = nullptr,
                                             const double* dtol = nullptr, double* pan = nullptr, int n_rows = 0) {
  if (n_rows < n) n_rows = n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, n_warps = nt >> 5;
  for (int kb = 0; kb < n; kb += kNB) {
    const int nb = min(kNB, n - kb);
    if (tid < 32) warp_diag_block(A, ld, kb, nb, dinv, ok, tid, dtol);
    __syncthreads();
    if (!*ok) return;
    const int r0 = kb + nb;
    double* pdiag = pan;                                   // [kNB][kPanStride]: the factored diagonal block
    double* prow = pan ? pan + kNB * kPanStride : nullptr;  // [n - r0][kPanStride]: the panel rows
    if (pan) {
      if (tid < kNB * kNB) {
        const int c = tid / kNB, k = tid % kNB;
        pdiag[c * kPanStride + k] = (c < nb && k <= c) ? A[(size_t)(kb + c) * ld + kb + k] : 0.0;
      }
      __syncthreads();
    }
    // panel: L[i, panel] = A[i, panel] * Lkk^-T by forward substitution along the row (multiplies by 1 / Lcc)
    for (int i = r0 + tid; i < n_rows; i += nt) {
      double x[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++) {
        if (c < nb) {
          double v = A[(size_t)i * ld + kb + c];
#pragma unroll
          for (int k = 0; k < kNB; k++)
            if (k < c) v -= x[k] * (pan ? pdiag[c * kPanStride + k] : A[(size_t)(kb + c) * ld + kb + k]);
          x[c] = v * dinv[kb + c];
        } else {
          x[c] = 0.0;
        }
      }
#pragma unroll
      for (int c = 0; c < kNB; c++) {
        if (c < nb) A[(size_t)i * ld + kb + c] = x[c];
        if (pan) prow[(size_t)(i - r0) * kPanStride + c] = x[c];
      }
    }
    __syncthreads();
    // trailing lower triangle: warps stride the rows, lanes the columns (no integer division)
    for (int i = r0 + warp; i < n_rows; i += n_warps) {
      double li[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++)
        li[c] = c < nb ? (pan ? prow[(size_t)(i - r0) * kPanStride + c] : A[(size_t)i * ld + kb + c]) : 0.0;
      for (int j = r0 + lane; j <= i && j < n; j += 32) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < kNB; c++)
          if (c < nb) acc += li[c] * (pan ? prow[(size_t)(j - r0) * kPanStride + c] : A[(size_t)j * ld + kb + c]);
        A[(size_t)i * ld + j] -= acc;
      }
    }
    __syncthreads();
  }
  if (Linv) {
    for (int blk = warp; blk * kNB < n; blk += n_warps)
      warp_block_inverse(A, ld, blk * kNB, min(kNB, n - blk * kNB), dinv, Linv + (size_t)blk * kNB * kNB, lane);
    __syncthreads();
  }
}

// The same factorization with one panel of look-ahead, for A in shared memory: after the panel solve the next block
// column is updated first, then warp 0 factors the next diagonal block while the other warps update the rest of the
// trailing matrix — the serial 8-column elimination (about half of a panel step at n ~ 100) leaves the critical path.
// Same operations on every entry in the same order as cta_cholesky, so the factor is identical bit for bit.
__device__ __forceinline__ void cta_cholesky_lookahead(double* A, int ld, int n, int* ok, double* dinv, int n_rows = 0) {
  if (n_rows < n) n_rows = n;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, n_warps = nt >> 5;
  if (tid < 32) warp_diag_block(A, ld, 0, min(kNB, n), dinv, ok, tid, nullptr);
  __syncthreads();
  if (!*ok) return;
#ifdef CLIC_LM_TIMING
  if (tid == 0) for (int i = 9; i < 16; i++) g_lm_stamps[i] = 0;
#endif
  for (int kb = 0; kb < n; kb += kNB) {
    const int nb = min(kNB, n - kb);
    const int r0 = kb + nb;
    long long tp0 = LM_NOW();
    // panel: L[i, panel] = A[i, panel] * Lkk^-T by forward substitution along the row
    for (int i = r0 + tid; i < n_rows; i += nt) {
      double x[kNB];
#pragma unroll
      for (int c = 0; c < kNB; c++) {
        if (c < nb) {
          double v = A[(size_t)i * ld + kb + c];
#pragma unroll
          for (int k = 0; k < kNB; k++)
            if (k < c) v -= x[k] * A[(size_t)(kb + c) * ld + kb + k];
          x[c] = v * dinv[kb + c];
        } else {
          x[c] = 0.0;
        }
      }
#pragma unroll
      for (int c = 0; c < kNB; c++)
        if (c < nb) A[(size_t)i * ld + kb + c] = x[c];
    }
    __syncthreads();
    LM_ACC(9, tp0);
    if (r0 >= n) break;
    const int nb1 = min(kNB, n - r0);
    tp0 = LM_NOW();
    // next block column first: entries (i, j), j in [r0, r0 + nb1), j <= i
    for (int idx = tid; idx < (n_rows - r0) * kNB; idx += nt) {
      const int i = r0 + (idx >> 3), j = r0 + (idx & 7);
      if (j <= i && j < r0 + nb1) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < kNB; c++)
          if (c < nb) acc += A[(size_t)i * ld + kb + c] * A[(size_t)j * ld + kb + c];
        A[(size_t)i * ld + j] -= acc;
      }
    }
    __syncthreads();
    LM_ACC(10, tp0);
    tp0 = LM_NOW();
    if (warp == 0) {
      warp_diag_block(A, ld, r0, nb1, dinv, ok, lane, nullptr);
      LM_ACC(11, tp0);
    } else {  // the rest of the trailing lower triangle: columns j >= r0 + nb1
      const int c0 = r0 + nb1;
      for (int i = c0 + warp - 1; i < n_rows; i += n_warps - 1) {
        double li[kNB];
#pragma unroll
        for (int c = 0; c < kNB; c++) li[c] = c < nb ? A[(size_t)i * ld + kb + c] : 0.0;
        for (int j = c0 + lane; j <= i && j < n; j += 32) {
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < kNB; c++)
            if (c < nb) acc += li[c] * A[(size_t)j * ld + kb + c];
          A[(size_t)i * ld + j] -= acc;
        }
      }
    }
    __syncthreads();
    LM_ACC(12, tp0);
    if (!*ok) return;
  }
}

// Both triangular solves by ONE warp (n <= 96: three unknowns per lane), no block barriers: y <- L^-1 y, y <- L^-T y.
// Lane l owns y[l], y[l + 32], y[l + 64].  With an odd leading dimension the column reads of the forward sweep are
// conflict-free.
// forward = false: y already holds L^-1 y (the right-hand side rode along the factorization as an extra row).
__device__ __forceinline__ void warp_cholesky_solve(const double* A, int ld, int n, double* y, const double* dinv, int lane,
                                                    bool forward = true) {
  const unsigned full = 0xffffffffu;
  double v[3];
#pragma unroll
  for (int q = 0; q < 3; q++) v[q] = lane + 32 * q < n ? y[lane + 32 * q] : 0.0;
  for (int i = 0; forward && i < n; i++) {  // forward: y_i /= L_ii; y_r -= L_ri y_i for r > i
    const int q = i >> 5;
    double yi = (q == 0 ? v[0] : q == 1 ? v[1] : v[2]);
    yi = __shfl_sync(full, yi, i & 31) * dinv[i];
    if (lane == (i & 31)) { if (q == 0) v[0] = yi; else if (q == 1) v[1] = yi; else v[2] = yi; }
#pragma unroll
    for (int p = 0; p < 3; p++) {
      const int r = lane + 32 * p;
      if (r > i && r < n) v[p] -= A[(size_t)r * ld + i] * yi;
    }
  }
  // backward: y_i /= L_ii; y_r -= L_ir y_i for r < i.  One warp, n dependent steps: the row of L and 1 / L_ii of the
  // NEXT step are loaded before the shuffle of this one, and the three 32-row segments use their registers statically,
  // so a step is shuffle -> multiply -> FMA on the dependent chain and nothing else.
  {
    int i = n - 1;
    double di = i >= 0 ? dinv[i] : 0.0;
    double r0v = (i >= 0 && lane < i) ? A[(size_t)i * ld + lane] : 0.0;
    double r1v = (i >= 0 && lane + 32 < i) ? A[(size_t)i * ld + lane + 32] : 0.0;
    double r2v = (i >= 0 && lane + 64 < i) ? A[(size_t)i * ld + lane + 64] : 0.0;
#define CLIC_BACK_STEP(VQ)                                                                   \
    {                                                                                        \
      const int in = i - 1;                                                                  \
      const double dn = in >= 0 ? dinv[in] : 0.0;                                            \
      const double n0 = (in >= 0 && lane < in) ? A[(size_t)in * ld + lane] : 0.0;            \
      const double n1 = (in >= 0 && lane + 32 < in) ? A[(size_t)in * ld + lane + 32] : 0.0;  \
      const double n2 = (in >= 0 && lane + 64 < in) ? A[(size_t)in * ld + lane + 64] : 0.0;  \
      const double yi = __shfl_sync(full, VQ, i & 31) * di;                                  \
      if (lane == (i & 31)) VQ = yi;                                                         \
      v[0] -= r0v * yi;                                                                      \
      v[1] -= r1v * yi;                                                                      \
      v[2] -= r2v * yi;                                                                      \
      di = dn; r0v = n0; r1v = n1; r2v = n2;                                                 \
    }
    for (; i >= 64; i--) CLIC_BACK_STEP(v[2])
    for (; i >= 32; i--) CLIC_BACK_STEP(v[1])
    for (; i >= 0; i--) CLIC_BACK_STEP(v[0])
#undef CLIC_BACK_STEP
  }
#pragma unroll
  for (int q = 0; q < 3; q++)
    if (lane + 32 * q < n) y[lane + 32 * q] = v[q];
}

// y <- L^-1 y then y <- L^-T y (L from cta_cholesky with Linv), blocked by kNB: the diagonal step is an 8-lane
// mat-vec with the block inverse, the off-diagonal update is row-parallel.
__device__ __forceinline__ void cta_cholesky_solve(const double* A, int ld, int n, double* y, const double* Linv) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double t8[kNB];
  for (int kb = 0; kb < n; kb += kNB) {
    const int nb = min(kNB, n - kb);
    const double* Li = Linv + (size_t)(kb / kNB) * kNB * kNB;
    if (tid < kNB) {
      double v = 0.0;
      for (int c = 0; c < nb; c++) v += Li[tid * kNB + c] * y[kb + c];
      t8[tid] = v;
    }
    __syncthreads();
    if (tid < nb) y[kb + tid] = t8[tid];
    for (int i = kb + nb + tid; i < n; i += nt) {
      double acc = 0.0;
      for (int c = 0; c < nb; c++) acc += A[(size_t)i * ld + kb + c] * t8[c];
      y[i] -= acc;
    }
    __syncthreads();
  }
  for (int kb = ((n - 1) / kNB) * kNB; kb >= 0; kb -= kNB) {
    const int nb = min(kNB, n - kb);
    const double* Li = Linv + (size_t)(kb / kNB) * kNB * kNB;
    if (tid < kNB) {
      double v = 0.0;
      for (int c = 0; c < nb; c++) v += Li[c * kNB + tid] * y[kb + c];  // Lkk^-T
      t8[tid] = v;
    }
    __syncthreads();
    if (tid < nb) y[kb + tid] = t8[tid];
    for (int i = tid; i < kb; i += nt) {
      double acc = 0.0;
      for (int c = 0; c < nb; c++) acc += A[(size_t)(kb + c) * ld + i] * t8[c];
      y[i] -= acc;
    }
    __syncthreads();
  }
}

}  // namespace clic

namespace clic {

// ---------------- Levenberg-Marquardt on the device (Ceres 1.14 TrustRegionMinimizer semantics) ----------------
// The constants are Ceres defaults (solver.h); the loop mirrors oracle/clic_oracle.cpp solve() line by line.
struct LmConst {
  double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  double max_radius = 1e16, min_radius = 1e-32, initial_radius = 1e4;
  int max_consecutive_invalid = 5;
};

struct LmState {
  double radius, decrease_factor;
  double x_cost, cand_cost, model_cost_change;
  double x_norm, gmax, step_norm;
  double initial_cost, fixed_cost;
  double ls_g0, ls_dmax, ls_cur, ls_fcur;  // projected line search (bounded problems)
  int reuse_diagonal, iteration, invalid, done, termination, reason;
  int n_succ, n_unsucc, n_jac, n_res;
  int skip_jac;  // control word of the Jacobian pass (1 = skip)
  int skip_res;  // control word of the residual-only pass
  int step_valid, max_iterations;
  int ls_active, ls_iter, ls_success;
  // developer timeline (build with -DCLIC_LM_TIMELINE, run with CLIC_LM_TIMELINE=1): entry / exit %globaltimer of the LM kernels
  int nts;
  int tag[96];
  unsigned long long ts[96];
};
// (developer build -DCLIC_LM_TIMELINE: every stamp is a global read-modify-write by thread 0, ~1 us)
__device__ __forceinline__ void lm_ts(LmState* s, int tag) {
#ifdef CLIC_LM_TIMELINE
  if (threadIdx.x == 0 && s->nts < 96) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    s->ts[s->nts] = t;
    s->tag[s->nts++] = tag;
  }
#else
  (void)s;
  (void)tag;
#endif
}

struct LayoutDev {
  const int* lm_col;  // per landmark: column of its inverse depth, -1 if constant (nullptr: none is free)
  int D, K, kf, nb, L;
  int col_bias_gyro, col_bias_accel, col_toff[3];
  int has_bounds[3];
  double lo[3], hi[3];
  StateLayout sl;
};

// x (+) delta for the whole state, one thread per block (ceres_local_param.h:132-139 + ParameterBlock::Plus box
// projection).  alpha scales delta (line search).
__device__ __forceinline__ void plus_state(const double* x, const double* delta, double alpha, double* out,
                                           const LayoutDev& L) {
  const int n_state = L.sl.size();
  for (int i = threadIdx.x; i < n_state; i += blockDim.x) out[i] = x[i];
  __syncthreads();
  for (int k = L.kf + threadIdx.x; k < L.K; k += blockDim.x) {
    const int c = 6 * (k - L.kf);
    Q4 q = ldq(x + L.sl.off_q() + 4 * k);
    Q4 r = so3_mul(q, so3_exp(mk(alpha * delta[c], alpha * delta[c + 1], alpha * delta[c + 2])));
    double* oq = out + L.sl.off_q() + 4 * k;
    oq[0] = r.x; oq[1] = r.y; oq[2] = r.z; oq[3] = r.w;
    double* op = out + L.sl.off_p() + 3 * k;
    const double* xp = x + L.sl.off_p() + 3 * k;
    for (int a = 0; a < 3; a++) op[a] = xp[a] + alpha * delta[c + 3 + a];
  }
  for (int j = threadIdx.x; j < L.nb; j += blockDim.x) {
    const double* xb = x + L.sl.off_bias() + 6 * j;
    double* ob = out + L.sl.off_bias() + 6 * j;
    if (L.col_bias_gyro >= 0)
      for (int a = 0; a < 3; a++) ob[a] = xb[a] + alpha * delta[L.col_bias_gyro + 3 * j + a];
    if (L.col_bias_accel >= 0)
      for (int a = 0; a < 3; a++) ob[3 + a] = xb[3 + a] + alpha * delta[L.col_bias_accel + 3 * j + a];
  }
  if (L.lm_col)
    for (int l = threadIdx.x; l < L.L; l += blockDim.x)
      if (L.lm_col[l] >= 0) out[L.sl.off_invd() + l] = x[L.sl.off_invd() + l] + alpha * delta[L.lm_col[l]];
  if (threadIdx.x < 3) {
    const int s = threadIdx.x;
    if (L.col_toff[s] >= 0) {
      double v = x[L.sl.off_toff() + s] + alpha * delta[L.col_toff[s]];
      if (L.has_bounds[s]) {
        v = fmax(v, L.lo[s]);
        v = fmin(v, L.hi[s]);
      }
      out[L.sl.off_toff() + s] = v;
    }
  }
  __syncthreads();
}

// Block-wide sum / max in a fixed order: butterfly inside every warp, then the warp results in warp order.
__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); w++) r += sh[w];
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_max(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = sh[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); w++) r = fmax(r, sh[w]);
  __syncthreads();
  return r;
}

// |x|^2, |x-y|^2, |x-y|_inf in the ambient space of the active (non-constant, used) blocks.
__device__ __forceinline__ void ambient_diff(const double* x, const double* y, const int* active, const LayoutDev& L,
                                             double* sh, double* xs_out, double* ds_out, double* dm_out) {
  double xs = 0, ds = 0, dm = 0;
  auto acc = [&](const double* a, const double* b, int n) {
    for (int i = 0; i < n; i++) {
      xs += a[i] * a[i];
      double d = a[i] - b[i];
      ds += d * d;
      dm = fmax(dm, fabs(d));
    }
  };
  // (the values are loaded whether or not the block is active: one memory round trip instead of two dependent ones)
  for (int k = L.kf + threadIdx.x; k < L.K; k += blockDim.x) {
    const int c = 6 * (k - L.kf);
    const int a0 = active[c], a1 = active[c + 3];
    double xv[7], yv[7];
    for (int i = 0; i < 4; i++) { xv[i] = x[L.sl.off_q() + 4 * k + i]; yv[i] = y[L.sl.off_q() + 4 * k + i]; }
    for (int i = 0; i < 3; i++) { xv[4 + i] = x[L.sl.off_p() + 3 * k + i]; yv[4 + i] = y[L.sl.off_p() + 3 * k + i]; }
    if (a0) acc(xv, yv, 4);
    if (a1) acc(xv + 4, yv + 4, 3);
  }
  for (int j = threadIdx.x; j < L.nb; j += blockDim.x) {
    const int a0 = L.col_bias_gyro >= 0 ? active[L.col_bias_gyro + 3 * j] : 0;
    const int a1 = L.col_bias_accel >= 0 ? active[L.col_bias_accel + 3 * j] : 0;
    double xv[6], yv[6];
    for (int i = 0; i < 6; i++) { xv[i] = x[L.sl.off_bias() + 6 * j + i]; yv[i] = y[L.sl.off_bias() + 6 * j + i]; }
    if (a0) acc(xv, yv, 3);
    if (a1) acc(xv + 3, yv + 3, 3);
  }
  if (threadIdx.x < 3 && L.col_toff[threadIdx.x] >= 0 && active[L.col_toff[threadIdx.x]])
    acc(x + L.sl.off_toff() + threadIdx.x, y + L.sl.off_toff() + threadIdx.x, 1);
  if (L.lm_col)
    for (int l = threadIdx.x; l < L.L; l += blockDim.x)
      if (L.lm_col[l] >= 0 && active[L.lm_col[l]]) acc(x + L.sl.off_invd() + l, y + L.sl.off_invd() + l, 1);
  // the three reductions share their barriers (same orders as block_sum / block_max: butterfly, then warps in order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xs += __shfl_xor_sync(0xffffffffu, xs, o);
    ds += __shfl_xor_sync(0xffffffffu, ds, o);
    dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
  }
  __syncthreads();
  const int nw = (int)(blockDim.x >> 5);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = xs;
    sh[nw + (threadIdx.x >> 5)] = ds;
    sh[2 * nw + (threadIdx.x >> 5)] = dm;
  }
  __syncthreads();
  double a = 0.0, b = 0.0, c = sh[2 * nw];
  for (int w = 0; w < nw; w++) {
    a += sh[w];
    b += sh[nw + w];
    c = fmax(c, sh[2 * nw + w]);
  }
  __syncthreads();
  if (xs_out) *xs_out = a;
  if (ds_out) *ds_out = b;
  if (dm_out) *dm_out = c;
}

struct LmBufs {
  double* H;       // D*D unscaled normal matrix (from the pass)
  double* b;       // D   unscaled gradient
  double* cost_x;  // [2] cost, fixed_cost of the pass at x
  double* cost_c;  // [2] of the pass at the candidate
  double* Hc;      // speculative evaluation: the system at the candidate (nullptr: residual-only candidate passes)
  double* bc;
  double* scale;   // D
  double* diagonal;  // D
  double* step;    // D
  double* delta;   // D
  double* A;       // D*D scratch (used when it does not fit shared memory)
  double* tmp;     // state-sized scratch (gradient-norm Plus)
  int* active;     // D
  double* x;       // state
  double* cand;    // candidate state
  LmState* st;
};

// gradient_max_norm = |x - Plus(x, -g)|_inf and |x| (TrustRegionMinimizer::EvaluateGradientAndJacobian)
// One pass: every thread forms Plus(x, -g) of ITS blocks in registers (the arithmetic of plus_state, block by block) and
// accumulates |x|^2 and |x - Plus|_inf in ambient_diff's order — the same numbers as plus_state into a scratch state +
// ambient_diff, in one memory round trip instead of four dependent ones.
__device__ __forceinline__ void gradient_norms(const LmBufs& B, const LayoutDev& L, double* sh) {
  const double* x = B.x;
  const double* g = B.b;
  const int* active = B.active;
  double xs = 0, ds = 0, dm = 0;
  auto acc = [&](const double* a, const double* b, int n) {
    for (int i = 0; i < n; i++) {
      xs += a[i] * a[i];
      double d = a[i] - b[i];
      ds += d * d;
      dm = fmax(dm, fabs(d));
    }
  };
  for (int k = L.kf + threadIdx.x; k < L.K; k += blockDim.x) {
    const int c = 6 * (k - L.kf);
    const int a0 = active[c], a1 = active[c + 3];
    double xv[7], yv[7], d[6];
    for (int i = 0; i < 4; i++) xv[i] = x[L.sl.off_q() + 4 * k + i];
    for (int i = 0; i < 3; i++) xv[4 + i] = x[L.sl.off_p() + 3 * k + i];
    for (int i = 0; i < 6; i++) d[i] = -g[c + i];
    const Q4 r = so3_mul(ldq(xv), so3_exp(mk(1.0 * d[0], 1.0 * d[1], 1.0 * d[2])));
    yv[0] = r.x; yv[1] = r.y; yv[2] = r.z; yv[3] = r.w;
    for (int a = 0; a < 3; a++) yv[4 + a] = xv[4 + a] + 1.0 * d[3 + a];
    if (a0) acc(xv, yv, 4);
    if (a1) acc(xv + 4, yv + 4, 3);
  }
  for (int j = threadIdx.x; j < L.nb; j += blockDim.x) {
    const int a0 = L.col_bias_gyro >= 0 ? active[L.col_bias_gyro + 3 * j] : 0;
    const int a1 = L.col_bias_accel >= 0 ? active[L.col_bias_accel + 3 * j] : 0;
    double xv[6], yv[6];
    for (int i = 0; i < 6; i++) xv[i] = yv[i] = x[L.sl.off_bias() + 6 * j + i];
    if (L.col_bias_gyro >= 0)
      for (int a = 0; a < 3; a++) yv[a] = xv[a] + 1.0 * (-g[L.col_bias_gyro + 3 * j + a]);
    if (L.col_bias_accel >= 0)
      for (int a = 0; a < 3; a++) yv[3 + a] = xv[3 + a] + 1.0 * (-g[L.col_bias_accel + 3 * j + a]);
    if (a0) acc(xv, yv, 3);
    if (a1) acc(xv + 3, yv + 3, 3);
  }
  if (threadIdx.x < 3 && L.col_toff[threadIdx.x] >= 0 && active[L.col_toff[threadIdx.x]]) {
    const int sidx = threadIdx.x;
    const double xv = x[L.sl.off_toff() + sidx];
    double v = xv + 1.0 * (-g[L.col_toff[sidx]]);
    if (L.has_bounds[sidx]) {
      v = fmax(v, L.lo[sidx]);
      v = fmin(v, L.hi[sidx]);
    }
    acc(&xv, &v, 1);
  }
  if (L.lm_col)
    for (int l = threadIdx.x; l < L.L; l += blockDim.x) {
      const int col = L.lm_col[l];
      if (col >= 0 && active[col]) {
        const double xv = x[L.sl.off_invd() + l];
        const double v = xv + 1.0 * (-g[col]);
        acc(&xv, &v, 1);
      }
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xs += __shfl_xor_sync(0xffffffffu, xs, o);
    dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
  }
  __syncthreads();
  const int nw = (int)(blockDim.x >> 5);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = xs;
    sh[nw + (threadIdx.x >> 5)] = dm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = sh[nw];
    for (int w = 0; w < nw; w++) {
      a += sh[w];
      c = fmax(c, sh[nw + w]);
    }
    B.st->x_norm = sqrt(a);
    B.st->gmax = c;
  }
  __syncthreads();
  (void)ds;
}

// IterationZero of a bounded problem: x <- Plus(x, 0) projects the start point onto the feasible set.
__global__ void lm_project_kernel(LmBufs B, LayoutDev L) {
  pdl_wait();
  pdl_trigger();
  for (int i = threadIdx.x; i < L.D; i += blockDim.x) B.step[i] = 0.0;
  __syncthreads();
  plus_state(B.x, B.step, 1.0, B.tmp, L);
  const int n_state = L.sl.size();
  for (int i = threadIdx.x; i < n_state; i += blockDim.x) B.x[i] = B.tmp[i];
}

__global__ void zero_cost_kernel(double* cost2, const int* ctl) {
  pdl_wait();
  pdl_trigger();
  if (ctl && ctl[0]) return;
  if (threadIdx.x < 2) cost2[threadIdx.x] = 0.0;
}

// After the iteration-0 Jacobian pass: Jacobi scaling, active mask, norms, LM state.
__global__ void lm_init_kernel(LmBufs B, LayoutDev L, LmConst C, int max_iterations, int is_constrained) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sh[];
  const int D = L.D;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const double d = B.H[(size_t)i * D + i];
    B.active[i] = d > 0.0 ? 1 : 0;
    B.scale[i] = 1.0 / (1.0 + sqrt(d));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    LmState* s = B.st;
  if (threadIdx.x == 0) s->nts = 0;
  lm_ts(s, 1);
    s->radius = C.initial_radius;
    s->decrease_factor = 2.0;
    s->x_cost = B.cost_x[0];
    s->fixed_cost = B.cost_x[1];
    s->initial_cost = B.cost_x[0] + B.cost_x[1];
    s->reuse_diagonal = 0;
    s->iteration = 0;
    s->invalid = 0;
    s->done = 0;
    s->termination = 1;
    s->reason = 0;
    s->n_succ = s->n_unsucc = 0;
    s->n_jac = 1;
    s->n_res = 0;
    s->skip_jac = 1;
    s->skip_res = 1;
    s->step_valid = 0;
    s->max_iterations = max_iterations;
    s->ls_active = 0;
    s->ls_iter = 0;
    s->ls_success = 1;
  }
  __syncthreads();
  gradient_norms(B, L, sh);
}

// One LM iteration, first half: termination tests, damping, Cholesky, step, model decrease, candidate = x (+) delta.
// A lives in shared memory when it fits (a_in_smem), else in global scratch.
// lm_step_kernel's middle: A = H~ + diag(diagonal / radius), Cholesky, (L L^T) y = g~.  Returns 0 if a pivot failed.
__device__ __forceinline__ int lm_factor_solve(double* A, int ld, int D, const LmBufs& B, const double* gs, double* y,
                                               double* dinv, double* Linv, double* pan, double radius, bool warp_solve,
                                               bool a_in_smem, int* ok) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid >> 5; i < D; i += nt >> 5) {  // warps stride the rows, lanes the columns
    const double si = B.scale[i];
    for (int j = tid & 31; j < D; j += 32) {
      double v = B.H[(size_t)i * D + j] * si * B.scale[j];
      if (i == j) {
        double lm = sqrt(B.diagonal[i] / radius);
        v += lm * lm;
      }
      A[(size_t)i * ld + j] = v;
    }
  }
  // D <= 96: the scaled gradient rides along the factorization as row D of A and comes out as L^-1 g~ (the forward
  // substitution costs no extra serial sweep); the host sized A for D + 1 rows
  if (warp_solve)
    for (int j = tid; j < D; j += nt) A[(size_t)D * ld + j] = gs[j];
  __syncthreads();
  LM_STAMP(2);
  // dense Cholesky A = L L^T (lower, in place), blocked right-looking
  if (warp_solve && a_in_smem) cta_cholesky_lookahead(A, ld, D, ok, dinv, D + 1);
  else cta_cholesky(A, ld, D, ok, dinv, warp_solve ? nullptr : Linv, nullptr, pan, warp_solve ? D + 1 : 0);
  __syncthreads();
  LM_STAMP(3);
  if (!*ok) return 0;
  // forward / backward substitution: (L L^T) z = g~
  for (int i = tid; i < D; i += nt) y[i] = warp_solve ? A[(size_t)D * ld + i] : gs[i];
  __syncthreads();
  if (warp_solve) {
    if (tid < 32) warp_cholesky_solve(A, ld, D, y, dinv, tid, false);
    __syncthreads();
  } else {
    cta_cholesky_solve(A, ld, D, y, Linv);
  }
  return 1;
}

__device__ __forceinline__ void lm_accept_phase(const LmBufs& B, const LayoutDev& L, const LmConst& C, double* sh);
__device__ __forceinline__ void lm_post_phase(const LmBufs& B, const LayoutDev& L, double* sh);

// with_accept: the launch first closes the PREVIOUS iteration (lm_accept_phase + lm_post_phase on the pass that ran in
// between) — the unconstrained solve with speculative Jacobians then is two launches per iteration (this kernel, the
// pass) instead of four.
__global__ void __launch_bounds__(512) lm_step_kernel(LmBufs B, LayoutDev L, LmConst C, int a_in_smem, int is_constrained,
                                                      int with_accept) {
  pdl_wait();
  if (with_accept != 2) pdl_trigger();
  extern __shared__ double sh[];
  lm_ts(B.st, 10);
  if (with_accept) {
    lm_accept_phase(B, L, C, sh);
    __syncthreads();
    lm_ts(B.st, 11);
    lm_post_phase(B, L, sh);
    __syncthreads();
    lm_ts(B.st, 12);
  }
  if (with_accept == 2) pdl_trigger();
  LmState* s = B.st;
  const int D = L.D;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int go;
  if (tid == 0) {
    go = 1;
    s->skip_res = 1;
    s->step_valid = 0;
    if (s->done) {
      go = 0;
    } else if (s->iteration >= s->max_iterations) {
      s->done = 1; s->termination = 1; s->reason = 0; go = 0;   // NO_CONVERGENCE / max iterations
    } else if (s->gmax <= C.gradient_tolerance) {
      s->done = 1; s->termination = 0; s->reason = 1; go = 0;
    } else if (s->radius <= C.min_radius) {
      s->done = 1; s->termination = 0; s->reason = 4; go = 0;
    } else {
      s->iteration++;
    }
  }
  __syncthreads();
  if (!go) return;
  LM_STAMP(0);
  double* red = sh;               // 512 doubles of reduction scratch
  double* gs = sh + 512;          // D scaled gradient
  double* y = gs + D;             // D
  double* dinv = y + D;           // D: 1 / L_jj
  double* Linv = dinv + D;        // ceil(D / 8) * 64 doubles: inverses of the diagonal blocks (block solves, D > 96)
  double* after_linv = Linv + ((D + kNB - 1) / kNB) * kNB * kNB;
  double* pan = a_in_smem ? nullptr : after_linv;  // (D + kNB) * kPanStride doubles: staged panel of the global path
  const int ld = a_in_smem ? (D | 1) : D;  // odd leading dimension in shared memory: conflict-free column reads
  const bool warp_solve = D <= 96;
  const double radius = s->radius;
  const int reuse = s->reuse_diagonal;
  // LevenbergMarquardtStrategy::ComputeStep: diagonal = clamp(diag(J~^T J~)), A = H~ + diag(diagonal / radius)
  for (int i = tid; i < D; i += nt) {
    const double si = B.scale[i];
    gs[i] = B.b[i] * si;
    if (!reuse) {
      double d = B.H[(size_t)i * D + i] * si * si;
      B.diagonal[i] = fmin(fmax(d, C.min_lm_diagonal), C.max_lm_diagonal);
    }
  }
  __syncthreads();
  LM_STAMP(1);
  // Build A = H~ + diag(diagonal / radius), factor it, solve.  Two inlined copies of the same code, so that the copy
  // working in shared memory addresses it with shared-memory instructions (the run-time choice of A would make every
  // access in the factorization a slower generic one).
  __shared__ int ok;
  if (tid == 0) ok = 1;
  int valid;
  if (a_in_smem) valid = lm_factor_solve(after_linv, ld, D, B, gs, y, dinv, Linv, nullptr, radius, warp_solve, true, &ok);
  else valid = lm_factor_solve(B.A, ld, D, B, gs, y, dinv, Linv, pan, radius, warp_solve, false, &ok);
  if (valid) {
    LM_STAMP(4);
    double bad = 0.0;
    for (int i = tid; i < D; i += nt) {
      double v = -y[i];
      B.step[i] = v;
      if (!isfinite(v)) bad = 1.0;
    }
    if (block_max(bad, red) > 0.0) valid = 0;
  }
  __syncthreads();
  LM_STAMP(5);
  double mcc = 0.0;
  if (valid) {
    // model_cost_change = -(g~^T s + 1/2 s^T H~ s)
    double part = 0.0;
    for (int i = tid >> 5; i < D; i += nt >> 5) {  // one warp per row of H~ (coalesced), butterfly over the lanes
      double row = 0.0;
      const double si = B.scale[i];
      for (int j = tid & 31; j < D; j += 32) row += (B.H[(size_t)i * D + j] * si * B.scale[j]) * B.step[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) row += __shfl_xor_sync(0xffffffffu, row, o);
      if ((tid & 31) == 0) part += gs[i] * B.step[i] + 0.5 * B.step[i] * row;
    }
    mcc = -block_sum(part, red);
    if (!(mcc > 0.0)) valid = 0;
  }
  if (tid == 0) s->reuse_diagonal = 1;
  if (!valid) {
    if (tid == 0) {  // HandleInvalidStep
      s->invalid++;
      if (s->invalid >= C.max_consecutive_invalid) {
        s->done = 1; s->termination = 2; s->reason = 5;
      } else {
        s->radius = s->radius / s->decrease_factor;
        s->decrease_factor *= 2.0;
        s->n_unsucc++;
      }
    }
    return;
  }
  LM_STAMP(6);
  for (int i = tid; i < D; i += nt) B.delta[i] = B.step[i] * B.scale[i];
  __syncthreads();
  plus_state(B.x, B.delta, 1.0, B.cand, L);
  LM_STAMP(7);
  double g0p = 0.0, dmx = 0.0;
  for (int i = tid; i < D; i += nt) {
    g0p += B.b[i] * B.delta[i];
    dmx = fmax(dmx, fabs(B.delta[i]));
  }
  const double g0 = block_sum(g0p, red), dmax = block_max(dmx, red);
  if (tid == 0) {
    s->invalid = 0;
    s->model_cost_change = mcc;
    s->step_valid = 1;
    s->skip_res = 0;
    s->n_res++;
    s->ls_g0 = g0;
    s->ls_dmax = dmax;
    s->ls_cur = 1.0;
    s->ls_iter = 0;
    s->ls_success = 1;
    s->ls_active = is_constrained;
  }
  LM_STAMP(8);
  lm_ts(s, 19);
}

// Projected Armijo line search of bounded problems (TrustRegionMinimizer::DoLineSearch).  Called by the host loop
// after each residual pass at cand = x (+) cur*delta while s->ls_active: decides accept / contract (oracle: quadratic
// interpolation with Ceres' contraction clamps), and writes the next candidate.
__global__ void lm_linesearch_kernel(LmBufs B, LayoutDev L, int max_ls_iterations) {
  pdl_wait();
  pdl_trigger();
  LmState* s = B.st;
  if (s->done || !s->step_valid || !s->ls_active) return;
  __shared__ int next;
  __shared__ double alpha;
  if (threadIdx.x == 0) {
    const double fcur = B.cost_c[0], cur = s->ls_cur, g0 = s->ls_g0, f0 = s->x_cost;
    next = 0;
    alpha = cur;
    if (!isfinite(fcur) || fcur > f0 + 1e-4 * g0 * cur) {
      s->ls_iter++;
      if (s->ls_iter >= max_ls_iterations) {
        s->ls_success = 0;
      } else {
        double denom = 2.0 * (fcur - f0 - g0 * cur);
        double q = (denom > 0 && isfinite(denom)) ? (-g0 * cur * cur / denom) : 0.6 * cur;
        q = fmin(fmax(q, 1e-3 * cur), 0.6 * cur);
        if (q * s->ls_dmax < 1e-9) {
          s->ls_success = 0;
        } else {
          s->ls_cur = q;
          alpha = q;
          next = 1;
          s->n_res++;
        }
      }
      if (!next) alpha = 1.0;  // failure: delta unchanged (Ceres scales delta only on success)
    } else {
      alpha = cur;             // success: delta *= optimal step
    }
    if (!next) s->ls_active = 0;
  }
  __syncthreads();
  const double a = alpha;
  if (!next) {
    for (int i = threadIdx.x; i < L.D; i += blockDim.x) B.delta[i] *= a;
    __syncthreads();
    plus_state(B.x, B.delta, 1.0, B.cand, L);
    // the final candidate needs its cost: identical to the last sample when a == ls_cur, otherwise re-evaluated
    if (threadIdx.x == 0) s->skip_res = (a == s->ls_cur && s->ls_success) ? 1 : 0;
  } else {
    plus_state(B.x, B.delta, a, B.cand, L);
    if (threadIdx.x == 0) s->skip_res = 0;
  }
}

// Second half of the iteration, after the pass at the candidate (all threads of one CTA; sh: >= 1024 doubles).
__device__ __forceinline__ void lm_accept_phase(const LmBufs& B, const LayoutDev& L, const LmConst& C, double* sh) {
  LmState* s = B.st;
  __shared__ int go;
  // thread 0: everything the decision reads, requested together with the go test (not one round trip later)
  double p_cand = 0, p_xcost = 0, p_mcc = 0, p_radius = 0, p_dec = 0, p_xnorm = 0;
  if (threadIdx.x == 0) {
    const int done = s->done, valid = s->step_valid;
    p_cand = B.cost_c[0]; p_xcost = s->x_cost; p_mcc = s->model_cost_change; p_radius = s->radius;
    p_dec = s->decrease_factor; p_xnorm = s->x_norm;
    s->skip_jac = 1;
    go = (!done && valid) ? 1 : 0;
  }
  __syncthreads();
  if (!go) return;
  // the candidate's system is on its way into registers while the decision is made (one memory round trip less on the
  // accept path; D <= 96: everything, else the first part)
  constexpr int kPre = 5;
  double2 pre[kPre];
  const int n2_pre = B.Hc ? (L.D * L.D) >> 1 : 0;
#pragma unroll
  for (int q = 0; q < kPre; q++) {
    const int i = threadIdx.x + q * blockDim.x;
    pre[q] = i < n2_pre ? reinterpret_cast<const double2*>(B.Hc)[i] : make_double2(0.0, 0.0);
  }
  double ds;
  ambient_diff(B.x, B.cand, B.active, L, sh, nullptr, &ds, nullptr);
  __shared__ int accept;
  if (threadIdx.x == 0) {
    accept = 0;
    double cand_cost = p_cand;
    if (!isfinite(cand_cost)) cand_cost = 1.7976931348623157e308;
    s->cand_cost = cand_cost;
    const double step_norm = sqrt(ds);
    s->step_norm = step_norm;
    const double cost_change = p_xcost - cand_cost;
    if (step_norm <= C.parameter_tolerance * (p_xnorm + C.parameter_tolerance)) {
      s->done = 1; s->termination = 0; s->reason = 2;
    } else if (fabs(cost_change) <= C.function_tolerance * p_xcost) {
      s->done = 1; s->termination = 0; s->reason = 3;
    } else {
      const double rd = cost_change / p_mcc;
      if (rd > C.min_relative_decrease) {
        accept = 1;
        s->n_succ++;
        const double t = 2.0 * rd - 1.0;  // (pow(t, 3) as Ceres writes it; t * t * t is what pow returns for an integer 3)
        s->radius = fmin(C.max_radius, p_radius / fmax(1.0 / 3.0, 1.0 - pow(t, 3)));
        s->decrease_factor = 2.0;
        s->reuse_diagonal = 0;
        s->skip_jac = 0;
        s->n_jac++;
      } else {
        s->n_unsucc++;
        s->radius = p_radius / p_dec;
        s->decrease_factor = p_dec * 2.0;
        s->reuse_diagonal = 1;
      }
    }
  }
  __syncthreads();
  if (accept) {
    const int n_state = L.sl.size();
    for (int i = threadIdx.x; i < n_state; i += blockDim.x) B.x[i] = B.cand[i];
    if (B.Hc) {  // the candidate was evaluated with its Jacobians: adopt its system, no further pass needed
      const int D = L.D;
      const int n2 = (D * D) >> 1;  // both buffers are 256-byte aligned: copy 16 bytes per thread and step
      const double2* src = reinterpret_cast<const double2*>(B.Hc);
      double2* dst = reinterpret_cast<double2*>(B.H);
#pragma unroll
      for (int q = 0; q < kPre; q++) {
        const int i = threadIdx.x + q * blockDim.x;
        if (i < n2) dst[i] = pre[q];
      }
#pragma unroll 4
      for (int i = threadIdx.x + kPre * blockDim.x; i < n2; i += blockDim.x) dst[i] = src[i];
      if ((D & 1) && threadIdx.x == 0) B.H[D * D - 1] = B.Hc[D * D - 1];
      for (int i = threadIdx.x; i < D; i += blockDim.x) B.b[i] = B.bc[i];
      if (threadIdx.x < 2) B.cost_x[threadIdx.x] = B.cost_c[threadIdx.x];
    }
  }
}
__global__ void __launch_bounds__(1024) lm_accept_kernel(LmBufs B, LayoutDev L, LmConst C) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sh[];
  lm_ts(B.st, 20);
  lm_accept_phase(B, L, C, sh);
  __syncthreads();
  lm_ts(B.st, 21);
}

// After the (conditional) Jacobian pass at the accepted point.
__device__ __forceinline__ void lm_post_phase(const LmBufs& B, const LayoutDev& L, double* sh) {
  LmState* s = B.st;
  if (s->done || s->skip_jac) return;
  if (threadIdx.x == 0) s->x_cost = B.cost_x[0];
  __syncthreads();
  gradient_norms(B, L, sh);
}
__global__ void lm_post_kernel(LmBufs B, LayoutDev L) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ double sh[];
  lm_ts(B.st, 30);
  lm_post_phase(B, L, sh);
  __syncthreads();
  lm_ts(B.st, 31);
}

}  // namespace clic

namespace clic {

// ---------------- K8: marginalization (MarginalizationInfo::marginalize, marginalization_factor.cpp:150-276) ----------------
// Rank-revealing Cholesky with diagonal pivoting (the dpstrf scheme), one CTA, unblocked, on a FULL symmetric n x n
// block (both triangles valid, leading dimension ld):  P^T A P = L L^T with L n x r, r = numerical rank (stop when the
// largest remaining diagonal <= n * eps * largest initial diagonal).  perm[i] = original index at position i.  On
// return the lower triangle holds L in pivoted position order (columns >= r are to be ignored).
__device__ __forceinline__ void cta_pivoted_cholesky(double* A, int ld, int n, int* perm, int* rank_out) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ double s_val[512];
  __shared__ int s_idx[512];
  __shared__ double s_tol;
  __shared__ int s_stop;
  for (int i = tid; i < n; i += nt) perm[i] = i;
  if (tid == 0) s_stop = 0;
  __syncthreads();
  int rank = n;
  for (int k = 0; k < n; k++) {
    // pivot = largest remaining diagonal (ties -> lowest index, fixed order)
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += nt) {
      const double d = A[(size_t)i * ld + i];
      if (d > best) { best = d; bi = i; }
    }
    s_val[tid] = best;
    s_idx[tid] = bi;
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) {
      if (tid < o) {
        const double v = s_val[tid + o];
        const int j = s_idx[tid + o];
        if (v > s_val[tid] || (v == s_val[tid] && j < s_idx[tid])) { s_val[tid] = v; s_idx[tid] = j; }
      }
      __syncthreads();
    }
    const int piv = s_idx[0];
    const double d = s_val[0];
    if (tid == 0) {
      if (k == 0) s_tol = (double)n * 2.220446049250313e-16 * d;
      if (!(d > s_tol) || !isfinite(d)) s_stop = 1;
    }
    __syncthreads();
    if (s_stop) { rank = k; break; }
    if (piv != k) {  // symmetric swap of rows / columns k <-> piv over the full block
      for (int j = tid; j < n; j += nt) {
        const double t = A[(size_t)k * ld + j];
        A[(size_t)k * ld + j] = A[(size_t)piv * ld + j];
        A[(size_t)piv * ld + j] = t;
      }
      __syncthreads();
      for (int i = tid; i < n; i += nt) {
        const double t = A[(size_t)i * ld + k];
        A[(size_t)i * ld + k] = A[(size_t)i * ld + piv];
        A[(size_t)i * ld + piv] = t;
      }
      if (tid == 0) { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
      __syncthreads();
    }
    const double lkk = sqrt(d), inv = 1.0 / lkk;
    for (int i = k + 1 + tid; i < n; i += nt) A[(size_t)i * ld + k] *= inv;  // column k of L (row k keeps A's values)
    __syncthreads();
    if (tid == 0) A[(size_t)k * ld + k] = lkk;
    // trailing block, both triangles: A[i][j] -= L[i][k] L[j][k]
    const int t = n - k - 1;
    for (int e = tid; e < t * t; e += nt) {
      const int i = k + 1 + e / t, j = k + 1 + e % t;
      A[(size_t)i * ld + j] -= A[(size_t)i * ld + k] * A[(size_t)j * ld + k];
    }
    __syncthreads();
  }
  if (tid == 0) *rank_out = rank;
  __syncthreads();
}

// Factor a full symmetric n x n block: plain blocked Cholesky first; if a pivot is <= n*eps*(its original diagonal)
// — a column numerically dependent on the previous ones — the block is restored, Jacobi-scaled (S A S, unit diagonal:
// the test is then independent of the units of the parameters, e.g. time offsets in ns next to metres) and factored
// with pivoting, and L is unscaled in place.  Returns the rank (every thread); perm = identity on the plain path.
// d0: n doubles of scratch (original diagonal), dtol: 2 n doubles of scratch (tolerances, inverse diagonal).
// pan: (320 + kNB) * kPanStride doubles of shared scratch (staged panel; A is in global memory).
__device__ __forceinline__ int cta_factor_block(double* A, int ld, int n, int* perm, double* d0, double* dtol, int* ok,
                                                int* rank_sh, double* pan) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n; i += nt) {
    d0[i] = A[(size_t)i * ld + i];
    dtol[i] = (double)n * 2.220446049250313e-16 * fabs(d0[i]);
    perm[i] = i;
  }
  if (tid == 0) *ok = 1;
  __syncthreads();
  cta_cholesky(A, ld, n, ok, dtol + n, nullptr, dtol, n <= 320 ? pan : nullptr);  // dinv scratch right behind dtol
  __syncthreads();
  if (*ok) return n;
  // restore (the failed attempt touched the lower triangle and the diagonal only) and scale
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    const double si = d0[i] > 0.0 ? rsqrt(d0[i]) : 0.0, sj = d0[j] > 0.0 ? rsqrt(d0[j]) : 0.0;
    const double v = i == j ? d0[i] : (j < i ? A[(size_t)j * ld + i] : A[(size_t)i * ld + j]);
    if (j <= i) {
      A[(size_t)i * ld + j] = v * si * sj;
      A[(size_t)j * ld + i] = v * si * sj;
    }
  }
  __syncthreads();
  cta_pivoted_cholesky(A, ld, n, perm, rank_sh);
  const int r = *rank_sh;
  for (int e = tid; e < n * r; e += nt) {  // L = S^-1 L_s: row i (position order) belongs to original column perm[i]
    const int i = e / r, k = e % r;
    if (k <= i) A[(size_t)i * ld + k] *= sqrt(fmax(d0[perm[i]], 0.0));
  }
  __syncthreads();
  if (tid == 0) *ok = 1;
  __syncthreads();
  return r;
}

// Schur complement of the first m (dropped) tangent dimensions of the pos x pos system (A, b), then the prior's
// square-root form.  The reference inverts A_mm and factors the result by symmetric eigendecompositions with an
// eigenvalue cut of 1e-15 (":208-264").  Here: plain Cholesky first — for the positive-definite systems of a
// well-constrained window it gives the same A' = A_rr - A_rm A_mm^-1 A_mr, b' and the same prior energy with
// J0 = L'^T, r0 = L'^-1 b' (J0^T J0 = A', J0^T r0 = b').  If a block is not positive definite, the rank-revealing
// pivoted factorization above takes over: for a positive semi-definite system the generalized Schur complement does not
// depend on which generalized inverse of A_mm is used, so G = P [L_r^-T L_r^-1, 0; 0, 0] P^T gives what the truncated
// eigen pseudo-inverse gives, and J0 = [L_r^T P^T; 0], r0 = [L_r^-1 (P^T b')_r; 0] reproduces J0^T J0 = A' and
// J0^T r0 = proj_range(A') b' = V V^T b'.   (Differs from the reference only where it inverts noise-level eigenvalues
// between 1e-15 and n*eps*max|diag|.)
//   A: pos x pos row-major, symmetric (both triangles valid); overwritten.   J0: n x n, r0: n.   perm: pos ints.
//   status[0] = 0 ok;  status[1] = rank deficiency of A_mm;  status[2] = rank deficiency of the Schur complement.
__global__ void __launch_bounds__(512) schur_kernel(double* A, double* b, int m, int n, double* J0, double* r0,
                                                    int* perm, double* dscr, double* Aug, int* status) {
  const int pos = m + n;
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int ok, rank;
  if (tid == 0) { ok = 1; status[1] = 0; status[2] = 0; }
  __syncthreads();
  // ---- fast path: ONE blocked Cholesky of the augmented matrix [A b; b^T .] over its first pos columns.  After the
  // first m columns the trailing block is the Schur complement A' and the appended row is b'; after all pos columns
  // the kept block holds L' (J0 = L'^T) and the appended row holds r0 = L'^-1 b'.  Works on a copy: if a pivot
  // fails the rank test the two-stage route below starts from the untouched A, b.
  __shared__ double s_pan[(320 + kNB) * kPanStride];
  if (Aug && pos + kNB <= 320) {
    const int ld = pos + 1;
    for (int e = tid; e < (pos + 1) * (pos + 1); e += nt) {
      const int i = e / ld, j = e % ld;
      double v = 0.0;
      if (i < pos && j < pos) v = (i < m && j < m) ? 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]) : A[(size_t)i * pos + j];
      else if (i == pos && j < pos) v = b[j];
      else if (j == pos && i < pos) v = b[i];
      Aug[e] = v;
    }
    for (int j = tid; j < pos; j += nt) dscr[pos + j] = (double)(j < m ? m : n) * 2.220446049250313e-16 * fabs(A[(size_t)j * pos + j]);
    __syncthreads();
    cta_cholesky(Aug, ld, pos, &ok, dscr + 2 * pos, nullptr, dscr + pos, s_pan, pos + 1);
    __syncthreads();
    if (ok) {
      for (int e = tid; e < n * n; e += nt) {
        const int k = e / n, a = e % n;  // J0[k][a] = L'[a][k]
        J0[e] = a >= k ? Aug[(size_t)(m + a) * ld + m + k] : 0.0;
      }
      for (int k = tid; k < n; k += nt) r0[k] = Aug[(size_t)pos * ld + m + k];
      if (tid == 0) status[0] = 0;
      return;
    }
    __syncthreads();
    if (tid == 0) ok = 1;
    __syncthreads();
  }
  if (m > 0) {
    // Amm = 0.5 (A_mm + A_mm^T)  (":208")
    for (int e = tid; e < m * m; e += nt) {
      const int i = e / m, j = e % m;
      if (j < i) {
        const double v = 0.5 * (A[(size_t)i * pos + j] + A[(size_t)j * pos + i]);
        A[(size_t)i * pos + j] = v;
        A[(size_t)j * pos + i] = v;
      }
    }
    __syncthreads();
    const int rm = cta_factor_block(A, pos, m, perm, dscr, dscr + pos, &ok, &rank, s_pan);
    if (tid == 0) status[1] = m - rm;
    // X = L_r^-1 (P^T A_mr) (rm x n) in place at rows perm[i], one column per thread; y = L_r^-1 (P^T b_m)
    for (int c = tid; c < n + 1; c += nt) {
      for (int i = 0; i < rm; i++) {
        const int ri = perm[i];
        double v = c < n ? A[(size_t)ri * pos + m + c] : b[ri];
        for (int k = 0; k < i; k++) v -= A[(size_t)i * pos + k] * (c < n ? A[(size_t)perm[k] * pos + m + c] : b[perm[k]]);
        v /= A[(size_t)i * pos + i];
        if (c < n) A[(size_t)ri * pos + m + c] = v; else b[ri] = v;
      }
    }
    __syncthreads();
    // A_rr -= X^T X (lower triangle), b_r -= X^T y
    for (int e = tid; e < n * n; e += nt) {
      const int i = e / n, j = e % n;
      if (j <= i) {
        double acc = 0.0;
        for (int k = 0; k < rm; k++) acc += A[(size_t)perm[k] * pos + m + i] * A[(size_t)perm[k] * pos + m + j];
        A[(size_t)(m + i) * pos + m + j] -= acc;
      }
    }
    for (int i = tid; i < n; i += nt) {
      double acc = 0.0;
      for (int k = 0; k < rm; k++) acc += A[(size_t)perm[k] * pos + m + i] * b[perm[k]];
      b[m + i] -= acc;
    }
    __syncthreads();
  }
  // A' = L' L'^T on the kept block; J0 = L'^T, r0 = L'^-1 b'.  The updated values live in the lower triangle: mirror
  // them first (cta_factor_block wants both triangles).
  double* Ar = A + (size_t)m * pos + m;
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    if (j < i) Ar[(size_t)j * pos + i] = Ar[(size_t)i * pos + j];
  }
  __syncthreads();
  int* permr = perm + m;
  const int rn = cta_factor_block(Ar, pos, n, permr, dscr, dscr + pos, &ok, &rank, s_pan);
  if (tid == 0) status[2] = n - rn;
  // J0[k][perm[a]] = L'[a][k] for k < rn (a >= k), zero rows below
  for (int e = tid; e < n * n; e += nt) J0[e] = 0.0;
  __syncthreads();
  for (int e = tid; e < n * n; e += nt) {
    const int k = e / n, a = e % n;
    if (k < rn && a >= k) J0[(size_t)k * n + permr[a]] = Ar[(size_t)a * pos + k];
  }
  if (tid == 0) {
    for (int i = 0; i < n; i++) {
      if (i < rn) {
        double v = b[m + permr[i]];
        for (int k = 0; k < i; k++) v -= Ar[(size_t)i * pos + k] * r0[k];
        r0[i] = v / Ar[(size_t)i * pos + i];
      } else {
        r0[i] = 0.0;
      }
    }
    status[0] = 0;
  }
}

// which interval keys received contributions: present[s] = (block diagonal of key s is non-zero)
__global__ void key_presence_kernel(const double* fin, int n_keys, int rd, int NT, int r_col, int* present) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_keys) return;
  // the diagonal of the knot columns and the r~ column are enough: a contributing factor has r != 0 or J != 0
  const double* blk = fin + (size_t)s * rd;
  bool any = false;
  for (int l = 0; l < 24 && !any; l++) any = block_entry(blk, NT, l, l) != 0.0;
  if (r_col >= 0) any = any || block_entry(blk, NT, r_col, r_col) != 0.0;  // physical column of r~
  present[s] = any ? 1 : 0;
}

}  // namespace clic
