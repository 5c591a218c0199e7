// LiDAR data association on the device: LidarOdometry::FindCorrespondence,
// src/lidar_odometry/lidar_odometry.cpp:483-657 — exact 5-nearest-neighbour search of every selected scan feature in
// the feature map, then a line fit (corner features) or a plane fit (surface features) over the 5 neighbours.
//
// The reference searches a kd-tree per feature on the host.  The map here is 10^4-10^5 float points (1.2 MB, L2
// resident) and at most ~1500 features per type are matched per call (the reference's own stride, ":501, :575"), so the
// device version is an exact brute-force scan, cut both ways to fill the GPU: a CTA takes 32 features (4 per warp) and
// one slice of the map, streamed through shared memory; every lane keeps a sorted best-5 per feature in registers, a
// 5-round warp arg-min merges the 32 lists, and a second kernel merges the slices and fits.
// Ties are broken by the smaller map index, so the result is a pure function of the inputs.
// The float / double arithmetic of the fits follows the reference operation by operation with explicit
// round-to-nearest intrinsics (no FMA contraction), so a host build of the same operations without contraction produces
// the same bits; the tests check exactly that.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace assoc_dev {

constexpr int kTile = 2048;       // map points per shared-memory tile (24 KB)
constexpr int kWarps = 8;         // features per CTA
// The reference drops a feature whose 5th neighbour is farther than 1 m^2 (":508, :584"), so only map points with
// d^2 <= 1 can matter: the lists start from the first float above 1 instead of +inf, and the insertion branch — which is
// what a brute-force scan spends its time in while the lists are still cold — fires only for the few points in reach.
constexpr float kInf = 1.00000011920928955078125f;  // nextafterf(1, 2): "no neighbour within the gate"

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float hyp(float a, float b) {
  return (float)__dsqrt_rn(__dadd_rn(__dmul_rn((double)a, (double)a), __dmul_rn((double)b, (double)b)));
}

struct Best5 {
  float d[5];
  int id[5];
};

__device__ __forceinline__ void best5_init(Best5& b) {
#pragma unroll
  for (int k = 0; k < 5; k++) {
    b.d[k] = kInf;
    b.id[k] = INT32_MAX;
  }
}
__device__ __forceinline__ void best5_insert(Best5& b, float d, int j) {  // caller checked d < b.d[4]
  b.d[4] = d;
  b.id[4] = j;
#pragma unroll
  for (int k = 4; k > 0; k--) {
    if (b.d[k] < b.d[k - 1]) {  // strict: an equal distance stays behind the smaller index (lane scans ascending j)
      const float td = b.d[k];
      b.d[k] = b.d[k - 1];
      b.d[k - 1] = td;
      const int ti = b.id[k];
      b.id[k] = b.id[k - 1];
      b.id[k - 1] = ti;
    }
  }
}

// lidar_odometry.cpp:511-563.  cv::eigen (OpenCV, absent) is restated as Jacobi rotations on the largest off-diagonal
// element; Eigen's ColPivHouseholderQR below as Householder QR with pivoting on the largest remaining column norm.
__device__ inline bool fit_line(const float (&pts)[5][3], double* cp, double* dir) {
  cp[0] = cp[1] = cp[2] = 0;
#pragma unroll
  for (int j = 0; j < 5; j++)
#pragma unroll
    for (int c = 0; c < 3; c++) cp[c] = __dadd_rn(cp[c], (double)pts[j][c]);
#pragma unroll
  for (int c = 0; c < 3; c++) cp[c] = __ddiv_rn(cp[c], 5.0);
  float a11 = 0, a12 = 0, a13 = 0, a22 = 0, a23 = 0, a33 = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float ax = (float)__dsub_rn((double)pts[j][0], cp[0]);
    const float ay = (float)__dsub_rn((double)pts[j][1], cp[1]);
    const float az = (float)__dsub_rn((double)pts[j][2], cp[2]);
    a11 = fadd(a11, fmul(ax, ax)); a12 = fadd(a12, fmul(ax, ay)); a13 = fadd(a13, fmul(ax, az));
    a22 = fadd(a22, fmul(ay, ay)); a23 = fadd(a23, fmul(ay, az)); a33 = fadd(a33, fmul(az, az));
  }
  a11 = fdiv(a11, 5.f); a12 = fdiv(a12, 5.f); a13 = fdiv(a13, 5.f);
  a22 = fdiv(a22, 5.f); a23 = fdiv(a23, 5.f); a33 = fdiv(a33, 5.f);
  // Jacobi on the upper triangle; off-diagonals o01, o02, o12; V rows are the eigenvectors.
  float o01 = a12, o02 = a13, o12 = a23;
  float W[3] = {a11, a22, a33};
  float V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int it = 0; it < 270; it++) {
    int sel = 0;  // 0: (0,1)  1: (0,2)  2: (1,2)
    float mv = fabsf(o01);
    if (mv < fabsf(o02)) { mv = fabsf(o02); sel = 1; }
    if (mv < fabsf(o12)) { mv = fabsf(o12); sel = 2; }
    const float p = sel == 0 ? o01 : (sel == 1 ? o02 : o12);
    if (fabsf(p) <= FLT_EPSILON) break;
    const int k = sel == 2 ? 1 : 0, l = sel == 0 ? 1 : 2;
    const float y = fmul(fsub(W[l], W[k]), 0.5f);
    float t = fadd(fabsf(y), hyp(p, y));
    float s = hyp(p, t);
    const float c = fdiv(t, s);
    s = fdiv(p, s);
    t = fmul(fdiv(p, t), p);
    if (y < 0) { s = -s; t = -t; }
    W[k] = fsub(W[k], t);
    W[l] = fadd(W[l], t);
    auto rot = [&](float& v0, float& v1) {
      const float a0 = v0, b0 = v1;
      v0 = fsub(fmul(a0, c), fmul(b0, s));
      v1 = fadd(fmul(a0, s), fmul(b0, c));
    };
    // the third index i is the one that is neither k nor l: i<k never (k,l>0 only for sel 2, i=0), k<i<l for sel 1
    if (sel == 0) { o01 = 0; rot(o02, o12); }            // i = 2 > l: rot(A[k][i], A[l][i])
    else if (sel == 1) { o02 = 0; rot(o01, o12); }       // i = 1, k < i < l: rot(A[k][i], A[i][l])
    else { o12 = 0; rot(o01, o02); }                     // i = 0 < k: rot(A[i][k], A[i][l])
#pragma unroll
    for (int i = 0; i < 3; i++) {
      // rows k and l of V; k, l are run-time values, so select rather than index dynamically
      float vk = k == 0 ? V[0][i] : V[1][i];
      float vl = l == 1 ? V[1][i] : V[2][i];
      rot(vk, vl);
      if (k == 0) V[0][i] = vk; else V[1][i] = vk;
      if (l == 1) V[1][i] = vl; else V[2][i] = vl;
    }
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    int m = k;
#pragma unroll
    for (int i = k + 1; i < 3; i++)
      if (W[m] < W[i]) m = i;
    if (m != k) {
      const float tw = W[m]; W[m] = W[k]; W[k] = tw;
#pragma unroll
      for (int i = 0; i < 3; i++) { const float tv = V[m][i]; V[m][i] = V[k][i]; V[k][i] = tv; }
    }
  }
  if (!(W[0] > fmul(3.f, W[1]))) return false;
#pragma unroll
  for (int c = 0; c < 3; c++) dir[c] = (double)V[0][c];
  return true;
}

// lidar_odometry.cpp:586-625.
__device__ inline bool fit_plane(const float (&pts)[5][3], double* plane) {
  float Q[5][3], cv[5];
  int perm[3] = {0, 1, 2};
#pragma unroll
  for (int j = 0; j < 5; j++) {
#pragma unroll
    for (int c = 0; c < 3; c++) Q[j][c] = pts[j][c];
    cv[j] = -1.f;
  }
  float tau[3] = {0, 0, 0};
  float maxn2 = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float n2 = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) n2 = fadd(n2, fmul(Q[j][c], Q[j][c]));
    if (n2 > maxn2) maxn2 = n2;
  }
  const float maxn = __fsqrt_rn(maxn2);
  const float me = fmul(maxn, FLT_EPSILON);
  const float thr = fdiv(fmul(me, me), 5.f);
  int nz = 3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    int big = k;
    float bign2 = -1.f;
#pragma unroll
    for (int c = k; c < 3; c++) {
      float n2 = 0;
#pragma unroll
      for (int j = k; j < 5; j++) n2 = fadd(n2, fmul(Q[j][c], Q[j][c]));
      if (n2 > bign2) { bign2 = n2; big = c; }
    }
    if (nz == 3 && bign2 < fmul(thr, (float)(5 - k))) nz = k;
#pragma unroll
    for (int c = k + 1; c < 3; c++) {  // swap columns k and big (static indices)
      if (big == c) {
#pragma unroll
        for (int j = 0; j < 5; j++) { const float tq = Q[j][k]; Q[j][k] = Q[j][c]; Q[j][c] = tq; }
        const int tp = perm[k]; perm[k] = perm[c]; perm[c] = tp;
      }
    }
    float tail2 = 0;
#pragma unroll
    for (int j = k + 1; j < 5; j++) tail2 = fadd(tail2, fmul(Q[j][k], Q[j][k]));
    const float c0 = Q[k][k];
    float beta;
    if (tail2 <= FLT_MIN) {
      tau[k] = 0;
      beta = c0;
#pragma unroll
      for (int j = k + 1; j < 5; j++) Q[j][k] = 0;
    } else {
      beta = __fsqrt_rn(fadd(fmul(c0, c0), tail2));
      if (c0 >= 0) beta = -beta;
      const float den = fsub(c0, beta);
#pragma unroll
      for (int j = k + 1; j < 5; j++) Q[j][k] = fdiv(Q[j][k], den);
      tau[k] = fdiv(fsub(beta, c0), beta);
    }
    Q[k][k] = beta;
#pragma unroll
    for (int c = k + 1; c < 3; c++) {
      float tmp = 0;
#pragma unroll
      for (int j = k + 1; j < 5; j++) tmp = fadd(tmp, fmul(Q[j][k], Q[j][c]));
      tmp = fadd(tmp, Q[k][c]);
      Q[k][c] = fsub(Q[k][c], fmul(tau[k], tmp));
#pragma unroll
      for (int j = k + 1; j < 5; j++) Q[j][c] = fsub(Q[j][c], fmul(fmul(tau[k], Q[j][k]), tmp));
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (k < nz) {
      float tmp = 0;
#pragma unroll
      for (int j = k + 1; j < 5; j++) tmp = fadd(tmp, fmul(Q[j][k], cv[j]));
      tmp = fadd(tmp, cv[k]);
      cv[k] = fsub(cv[k], fmul(tau[k], tmp));
#pragma unroll
      for (int j = k + 1; j < 5; j++) cv[j] = fsub(cv[j], fmul(fmul(tau[k], Q[j][k]), tmp));
    }
  }
  float xs[3] = {0, 0, 0}, x[3] = {0, 0, 0};
#pragma unroll
  for (int i = 2; i >= 0; i--) {
    if (i < nz) {
      float acc = cv[i];
#pragma unroll
      for (int c = i + 1; c < 3; c++)
        if (c < nz) acc = fsub(acc, fmul(Q[i][c], xs[c]));
      xs[i] = fdiv(acc, Q[i][i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    if (i < nz) {
#pragma unroll
      for (int c = 0; c < 3; c++)
        if (perm[i] == c) x[c] = xs[i];
    }
  }
  float pa = x[0], pb = x[1], pc = x[2], pd = 1.f;
  const float ps = __fsqrt_rn(fadd(fadd(fmul(pa, pa), fmul(pb, pb)), fmul(pc, pc)));
  pa = fdiv(pa, ps); pb = fdiv(pb, ps); pc = fdiv(pc, ps); pd = fdiv(pd, ps);
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const float v = fadd(fadd(fadd(fmul(pa, pts[j][0]), fmul(pb, pts[j][1])), fmul(pc, pts[j][2])), pd);
    if (!((double)fabsf(v) <= 0.2)) ok = false;
  }
  plane[0] = pa; plane[1] = pb; plane[2] = pc; plane[3] = pd;
  return ok;
}

// One feature type (corner or surface) of one clic_find_correspondence call.
struct AssocJob {
  const float* mx;        // feature map, structure of arrays (built once per map by soa_kernel)
  const float* my;
  const float* mz;
  int n_map;
  const float* xyz_M;     // features in the map frame, 3 n_feat
  const float* xyz_L;     // features in the LiDAR frame
  const double* t;        // feature timestamps
  long long step;         // every step-th feature is matched (":501, :575")
  int n_q;                // ceil(n_feat / step)
  int is_surface;         // 1: plane fit + timestamp gate; 0: line fit
  int n_slices;           // the map is cut into n_slices runs of slice_len points, one CTA row each
  int slice_len;
  float* part_d;          // per (feature, slice): the slice's best 5, sorted — 5 n_slices n_q
  int* part_i;
  unsigned char* valid;   // n_q
  double* geo;            // 6 n_q (plane uses 4)
  int* nn_idx;            // 5 n_q (kept for the parity tests)
  int geo_w;              // 4 (plane) or 6 (line)
  double* out_t;          // compacted records, the batches clic_add_loam_lines / _planes take
  double* out_point;
  double* out_geo;
  long long* out_src;
  long long* count;
};
struct AssocJobs {
  AssocJob j[2];
};

constexpr int kQPW = 4;                    // features per warp: every map point read from shared memory meets 4 queries
constexpr int kQPC = kQPW * kWarps;        // features per CTA

__global__ void soa_kernel(const float* __restrict__ xyz, int n, float* x, float* y, float* z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] = xyz[3 * (size_t)i];
    y[i] = xyz[3 * (size_t)i + 1];
    z[i] = xyz[3 * (size_t)i + 2];
  }
}

// 5 rounds of warp arg-min over the heads of the 32 sorted lane lists, (distance, index) lexicographic.
__device__ __forceinline__ void warp_merge5(Best5& b, float (&nd)[5], int (&nn)[5]) {
#pragma unroll
  for (int r = 0; r < 5; r++) {
    float bd = b.d[0];
    int bi = b.id[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    nn[r] = bi;
    nd[r] = bd;
    if (b.id[0] == bi && b.d[0] == bd) {  // the winner pops its head (indices are unique across lanes)
#pragma unroll
      for (int k = 0; k < 4; k++) { b.d[k] = b.d[k + 1]; b.id[k] = b.id[k + 1]; }
      b.d[4] = kInf;
      b.id[4] = INT32_MAX;
    }
  }
}

// Grid (ceil(max n_q / kQPC), max n_slices, 2 types).  CTA (g, s, type): 32 features against one slice of the map.
__global__ void __launch_bounds__(kWarps * 32) knn5_partial_kernel(AssocJobs jobs) {
  const AssocJob& J = jobs.j[blockIdx.z];
  if ((int)blockIdx.y >= J.n_slices || (int)blockIdx.x * kQPC >= J.n_q) return;
  __shared__ float sx[kTile], sy[kTile], sz[kTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQPC + warp * kQPW;
  float qx[kQPW], qy[kQPW], qz[kQPW];
  Best5 b[kQPW];
#pragma unroll
  for (int k = 0; k < kQPW; k++) {
    const int q = q0 + k;
    const long long feat = (long long)(q < J.n_q ? q : 0) * J.step;
    // a feature past the end compares as NaN: nothing is ever nearer than +inf
    qx[k] = q < J.n_q ? J.xyz_M[3 * feat] : __int_as_float(0x7fc00000);
    qy[k] = J.xyz_M[3 * feat + 1];
    qz[k] = J.xyz_M[3 * feat + 2];
    best5_init(b[k]);
  }
  const int s0 = blockIdx.y * J.slice_len, s1 = min(J.n_map, s0 + J.slice_len);
  for (int base = s0; base < s1; base += kTile) {
    const int tn = min(kTile, s1 - base);
    __syncthreads();
    for (int i = threadIdx.x; i < tn; i += blockDim.x) {
      sx[i] = J.mx[base + i];
      sy[i] = J.my[base + i];
      sz[i] = J.mz[base + i];
    }
    __syncthreads();
#pragma unroll 2
    for (int j = lane; j < tn; j += 32) {
      const float px = sx[j], py = sy[j], pz = sz[j];
      float d[kQPW];
      bool any = false;
#pragma unroll
      for (int k = 0; k < kQPW; k++) {
        const float dx = fsub(qx[k], px), dy = fsub(qy[k], py), dz = fsub(qz[k], pz);
        d[k] = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));  // flann::L2_Simple<float>
        any |= d[k] < b[k].d[4];
      }
      if (any) {  // one rarely taken branch per point instead of one per (point, feature)
#pragma unroll
        for (int k = 0; k < kQPW; k++)
          if (d[k] < b[k].d[4]) best5_insert(b[k], d[k], base + j);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kQPW; k++) {
    float nd[5];
    int nn[5];
    warp_merge5(b[k], nd, nn);
    const int q = q0 + k;
    if (q < J.n_q && lane < 5) {
      float vd = nd[0];
      int vi = nn[0];
#pragma unroll
      for (int r = 1; r < 5; r++)
        if (lane == r) { vd = nd[r]; vi = nn[r]; }
      const size_t o = ((size_t)q * J.n_slices + blockIdx.y) * 5 + lane;
      J.part_d[o] = vd;
      J.part_i[o] = vi;
    }
  }
}

// One warp per feature: merge the per-slice lists, apply the gates, fit.  Grid (ceil(max n_q / kWarps), 1, 2 types).
__global__ void __launch_bounds__(kWarps * 32) merge_fit_kernel(AssocJobs jobs) {
  const AssocJob& J = jobs.j[blockIdx.z];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * kWarps + warp;
  if (q >= J.n_q) return;
  const long long feat = (long long)q * J.step;
  const bool want = !(J.is_surface && J.t[feat] < 0);  // ":577"
  Best5 b;
  best5_init(b);
  const int n_cand = 5 * J.n_slices;
  for (int c = lane; c < n_cand; c += 32) {
    const float d = J.part_d[(size_t)q * n_cand + c];
    const int id = J.part_i[(size_t)q * n_cand + c];
    // a lane meets at most one candidate per slice, in slice order, so its indices ascend: strict insertion keeps
    // equal distances in index order
    if (d < b.d[4]) best5_insert(b, d, id);
  }
  float nd[5];
  int nn[5];
  warp_merge5(b, nd, nn);
  if (lane != 0) return;
  bool ok = want && nd[4] <= 1.0f;  // ":508, :584"; +inf when the map holds fewer than 5 points
  double geo[6] = {0, 0, 0, 0, 0, 0};
  if (ok) {
    float pts[5][3];
#pragma unroll
    for (int j = 0; j < 5; j++) {
      pts[j][0] = J.mx[nn[j]];
      pts[j][1] = J.my[nn[j]];
      pts[j][2] = J.mz[nn[j]];
    }
    ok = J.is_surface ? fit_plane(pts, geo) : fit_line(pts, geo, geo + 3);
  }
  J.valid[q] = ok ? 1 : 0;
#pragma unroll
  for (int c = 0; c < 6; c++) J.geo[6 * (size_t)q + c] = geo[c];
#pragma unroll
  for (int c = 0; c < 5; c++) J.nn_idx[5 * (size_t)q + c] = want ? nn[c] : -1;
}

// Order-preserving compaction of the kept records.  Grid (1, 1, 2 types), one CTA of 1024 threads per type; chunks
// of 1024 features with a running offset (n_q is <= 1500 in production).
__global__ void __launch_bounds__(1024) compact_kernel(AssocJobs jobs) {
  const AssocJob& C = jobs.j[blockIdx.z];
  __shared__ int warp_tot[32];
  __shared__ long long running;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int base = 0; base < C.n_q; base += 1024) {
    const int q = base + threadIdx.x;
    const bool v = q < C.n_q && C.valid[q];
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      const int c = warp_tot[w];
      if (w < warp) before += c;
      total += c;
    }
    const long long off = running + before + __popc(m & ((1u << lane) - 1u));
    if (v) {
      const long long feat = (long long)q * C.step;
      C.out_t[off] = C.t[feat];
#pragma unroll
      for (int c = 0; c < 3; c++) C.out_point[3 * off + c] = (double)C.xyz_L[3 * feat + c];
      for (int c = 0; c < C.geo_w; c++) C.out_geo[C.geo_w * off + c] = C.geo[6 * (size_t)q + c];
      C.out_src[off] = feat;
    }
    __syncthreads();
    if (threadIdx.x == 0) running += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *C.count = running;
}

// LidarOdometry::GetLoamFeatureAssociation after FindCorrespondence (lidar_odometry.cpp:170-175, 660-671): the
// correspondences — corner records first, then surface records — are sorted by t_point, every ds-th one is kept, and the
// kept ones are added as factors one by one.  Here: one CTA compacts the valid records of both types into shared memory,
// sorts them by (t_point, position in the unsorted list) with a bitonic network (the reference's std::sort leaves the
// order of equal timestamps open; this is the stable choice), keeps every ds-th and writes them out split by type, each
// batch still time-sorted — the layout the factor kernels want.  At most 1500 records per type exist by construction.
constexpr int kSortMax = 4096;
constexpr int kAlignKeys = 128;  // key-aligned output for windows of up to this many knot intervals
struct SortedOut {
  double* t[2];        // [0] lines, [1] planes
  double* point[2];    // 3 per record
  double* geo[2];      // 6 / 4 per record
  long long* src[2];
  long long* count;    // [0..1] stored records per type (padding included), [2..3] kept correspondences per type
  // Key-aligned storage (n_keys > 0): every knot interval's records start on a multiple of 32, the holes hold
  // records with a timestamp far outside any window, which the factor adder stores as dropped — each 32-factor slab
  // of the factor kernels then belongs to one interval (see build_key_gather in clic_cuda.cu).
  long long t0_ns, dt_ns;
  int n_keys;
};
constexpr double kHoleTime = -1e9;  // seconds

__global__ void __launch_bounds__(1024) sort_downsample_kernel(AssocJobs jobs, int ds, SortedOut out) {
  extern __shared__ double s_key[];               // kSortMax doubles
  int* s_ord = (int*)(s_key + kSortMax);          // kSortMax ints: (type << 16) | q, i.e. the unsorted-list order
  __shared__ int warp_tot[32];
  __shared__ int running, kept_run[2];
  __shared__ int cnt[2][kAlignKeys], start[2][kAlignKeys + 1], cum[2][kAlignKeys + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool align = out.n_keys > 0 && out.n_keys <= kAlignKeys;
  if (threadIdx.x == 0) { running = 0; kept_run[0] = kept_run[1] = 0; }
  for (int i = threadIdx.x; i < 2 * kAlignKeys; i += blockDim.x) cnt[i / kAlignKeys][i % kAlignKeys] = 0;
  __syncthreads();
  for (int type = 0; type < 2; type++) {
    const AssocJob& J = jobs.j[type];
    for (int base = 0; base < J.n_q; base += 1024) {
      const int q = base + threadIdx.x;
      const bool v = q < J.n_q && J.valid[q];
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (lane == 0) warp_tot[warp] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < 32; w++) {
        const int c = warp_tot[w];
        if (w < warp) before += c;
        total += c;
      }
      const int off = running + before + __popc(m & ((1u << lane) - 1u));
      if (v && off < kSortMax) {
        s_key[off] = J.t[(long long)q * J.step];
        s_ord[off] = (type << 16) | q;
      }
      __syncthreads();
      if (threadIdx.x == 0) running += total;
      __syncthreads();
    }
  }
  const int n = min(running, kSortMax);
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) {
    s_key[i] = __longlong_as_double(0x7ff0000000000000LL);
    s_ord[i] = INT32_MAX;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const double a = s_key[i], b = s_key[p];
          const int oa = s_ord[i], ob = s_ord[p];
          const bool up = (i & k) == 0;
          const bool gt = a > b || (a == b && oa > ob);
          if (gt == up) {
            s_key[i] = b; s_key[p] = a;
            s_ord[i] = ob; s_ord[p] = oa;
          }
        }
      }
      __syncthreads();
    }
  }
  // the knot interval of a record, as the factor kernels will index it (clamped: alignment is only a layout matter)
  auto key_of = [&](double t_s) {
    const long long rel = (long long)(t_s * 1e9) - out.t0_ns;
    long long k = rel < 0 ? 0 : rel / out.dt_ns;
    return (int)(k >= out.n_keys ? out.n_keys - 1 : k);
  };
  if (align) {
    for (int i = threadIdx.x * ds; i < n; i += blockDim.x * ds) atomicAdd(&cnt[s_ord[i] >> 16][key_of(s_key[i])], 1);
    __syncthreads();
    if (threadIdx.x < 2) {
      int a = 0, c = 0;
      for (int k = 0; k < out.n_keys; k++) {
        start[threadIdx.x][k] = a;
        cum[threadIdx.x][k] = c;
        a += (cnt[threadIdx.x][k] + 31) / 32 * 32;
        c += cnt[threadIdx.x][k];
      }
      start[threadIdx.x][out.n_keys] = a;
      cum[threadIdx.x][out.n_keys] = c;
    }
    __syncthreads();
    for (int type = 0; type < 2; type++) {  // holes first; the kept records overwrite their places below
      const int gw = jobs.j[type].geo_w, tot = start[type][out.n_keys];
      for (int o = threadIdx.x; o < tot; o += blockDim.x) {
        out.t[type][o] = kHoleTime;
        out.point[type][3 * o] = out.point[type][3 * o + 1] = out.point[type][3 * o + 2] = 0.0;
        for (int c = 0; c < gw; c++) out.geo[type][gw * o + c] = 0.0;
        out.src[type][o] = -1;
      }
    }
    __syncthreads();
  }
  // keep every ds-th of the sorted list; split by type with an order-preserving scan
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const bool kept = i < n && (i % ds) == 0;
    const int ord = kept ? s_ord[i] : 0;
    const int type = ord >> 16;
    const unsigned m0 = __ballot_sync(0xffffffffu, kept && type == 0);
    const unsigned m1 = __ballot_sync(0xffffffffu, kept && type == 1);
    if (lane == 0) warp_tot[warp] = __popc(m0) | (__popc(m1) << 8);
    __syncthreads();
    int before0 = 0, before1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      const int c0 = warp_tot[w] & 0xff, c1 = warp_tot[w] >> 8;
      if (w < warp) { before0 += c0; before1 += c1; }
      tot0 += c0;
      tot1 += c1;
    }
    if (kept) {
      const AssocJob& J = jobs.j[type];
      const int q = ord & 0xffff;
      const unsigned below = (1u << lane) - 1u;
      long long o = type == 0 ? kept_run[0] + before0 + __popc(m0 & below) : kept_run[1] + before1 + __popc(m1 & below);
      if (align) {  // o-th kept record of its type -> its interval's slab-aligned run
        const int k = key_of(s_key[i]);
        o = start[type][k] + (o - cum[type][k]);
      }
      const long long feat = (long long)q * J.step;
      out.t[type][o] = s_key[i];
#pragma unroll
      for (int c = 0; c < 3; c++) out.point[type][3 * o + c] = (double)J.xyz_L[3 * feat + c];
      for (int c = 0; c < J.geo_w; c++) out.geo[type][J.geo_w * o + c] = J.geo[6 * (size_t)q + c];
      out.src[type][o] = feat;
    }
    __syncthreads();
    if (threadIdx.x == 0) { kept_run[0] += tot0; kept_run[1] += tot1; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out.count[2] = kept_run[0];
    out.count[3] = kept_run[1];
    out.count[0] = align ? start[0][out.n_keys] : kept_run[0];
    out.count[1] = align ? start[1][out.n_keys] : kept_run[1];
  }
}

}  // namespace assoc_dev
