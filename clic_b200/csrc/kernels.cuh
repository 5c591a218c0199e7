// sm_100a kernels of the fused residual + Jacobian + normal-equation pass.
//
// Design:
//   * one factor per thread; the knot window and the per-knot-pair PairData are staged in shared memory once
//     per CTA (the whole window is <= 8 KB);
//   * FP64 on B200 is one pipe (DFMA and DMMA measured at 36-37 TFLOP/s and NOT overlapping,
//     profiles/r01_fp64_peaks.txt), so the J~^T J~ contraction goes through DMMA.8x8x4 only for its operand
//     economy: a warp transposes its 32 Jacobian rows through a bank-conflict-free shared tile and issues the
//     upper-triangular 8x8 tile products; the accumulators stay in registers across all slabs of the warp;
//   * a warp flushes its accumulators into a private record slot when the knot interval (the "key") of its
//     factors changes or at the end: no atomics, fixed order => bitwise run-to-run determinism;
//   * reduce_all_kernel + assemble_kernel turn records into the dense H, b (owner-computes, no atomics).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "factor_eval.cuh"

namespace clic {

#ifndef CLIC_WARPS_PER_CTA
#define CLIC_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = CLIC_WARPS_PER_CTA;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kJtStride = 36;  // doubles per column of the transposed tile: = 4 (mod 16) => conflict-free fragment loads
constexpr int kSlotsPerWarp = 4;
constexpr int64_t kDropped = INT64_MIN;  // t_ns sentinel of a measurement dropped by MeasuredTimeToNs

__host__ __device__ constexpr int n_tiles(int NT) { return NT * (NT + 1) / 2; }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// State vector layout on the device: [knot_q 4K | knot_p 3K | bias 6nb | t_offset 3 | gravity 3 | inv_depth L]
// Programmatic dependent launch (PDL): a kernel launched with the attribute may start while its predecessor in the
// stream is still draining.  pdl_trigger() lets the successor's CTAs be scheduled as SMs free up; pdl_wait() blocks
// until the predecessor grid has completed and its writes are visible.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct StateLayout {
  int K, nb, L;
  __host__ __device__ int off_q() const { return 0; }
  __host__ __device__ int off_p() const { return 4 * K; }
  __host__ __device__ int off_bias() const { return 7 * K; }
  __host__ __device__ int off_toff() const { return 7 * K + 6 * nb; }
  __host__ __device__ int off_grav() const { return 7 * K + 6 * nb + 3; }
  __host__ __device__ int off_invd() const { return 7 * K + 6 * nb + 6; }
  __host__ __device__ int size() const { return 7 * K + 6 * nb + 6 + L; }
};

struct PassParams {
  const double* state;  // state vector the pass is evaluated at
  StateLayout sl;
  int64_t t0_ns, dt_ns;
  int kf;               // first free local knot (factors whose 4 knots are all < kf only add to fixed_cost)
  int marg_mode;        // 1: marginalization set semantics (gates below), every block has a column
  int marg_later_local; // ctrl_to_be_opt_later - knot_id0
  const int* ctl;       // device control word (LM loop): ctl[0] != 0 => skip this pass
  // PDL: 1 = first factor kernel of a pass (its inputs come from the predecessor: wait before reading anything);
  // 0 = independent of its predecessor (another factor kernel): it overlaps that kernel's tail and only waits
  // before exiting, so that "this grid completed" still implies "every earlier grid completed".
  int pdl_first;
};
__device__ __forceinline__ bool pass_prologue(const PassParams& pp) {  // true: skip this pass
  if (pp.pdl_first) pdl_wait();
  pdl_trigger();
  if (pp.ctl && pp.ctl[0]) {
    if (!pp.pdl_first) pdl_wait();
    return true;
  }
  return false;
}
__device__ __forceinline__ void pass_epilogue(const PassParams& pp) {
  if (!pp.pdl_first) pdl_wait();
}

struct LoamStream {
  int64_t n;
  const int64_t* t_ns;
  const double* p[3];
  const double* g[6];
  const uint8_t* type;  // 1 plane, 0 line (mixed streams only)
  int geo_kind;         // 0 mixed, 1 planes only, 2 lines only (loam_kernel's GEO)
  LoamShared sh;
};

// Per-warp output of a factor kernel.
struct RecordBuf {
  double* payload;  // [n_warps * kSlotsPerWarp][n_tiles * 64]
  int* key;         // [n_warps * kSlotsPerWarp], -1 = unused
  double* cost;     // [n_warps][2]: cost, fixed_cost
  double* overflow; // [n_keys][n_tiles * 64] atomics fallback when a warp runs out of slots (unsorted input)
  int* overflow_flag;
};

// Record size in doubles: the upper-triangular 8x8 tiles + (EXTRA) one double per lane (LiDAR: b = J~^T r~, 24 used)
__host__ __device__ constexpr int rec_doubles(int NT, int EXTRA) { return n_tiles(NT) * 64 + 32 * EXTRA; }

template <int NT, int EXTRA = 0>
struct WarpAcc {
  double c[n_tiles(NT)][2];
  double extra;
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int t = 0; t < n_tiles(NT); t++) c[t][0] = c[t][1] = 0.0;
    extra = 0.0;
  }
  // Jt: this warp's transposed tile, Jt[col * kJtStride + row], rows 0..31 = the slab's Jacobian rows.
  // NA: only the first NA column tiles of these rows are non-zero (the others are skipped, not even read).
  template <int NA = NT>
  __device__ __forceinline__ void accumulate(const double* Jt, int lane) {
    const int r4 = lane & 3, c8 = lane >> 2;
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      double a[NA];
#pragma unroll
      for (int t = 0; t < NA; t++) a[t] = Jt[(8 * t + c8) * kJtStride + 4 * ks + r4];
#pragma unroll
      for (int ta = 0; ta < NA; ta++)
#pragma unroll
        for (int tb = ta; tb < NA; tb++) {
          constexpr int dummy = 0; (void)dummy;
          const int idx = ta * NT - ta * (ta - 1) / 2 + (tb - ta);
          dmma884(c[idx][0], c[idx][1], a[ta], a[tb]);
        }
    }
  }
  // the same over an arbitrary subset of the column tiles (bit t of MASK: tile t takes part)
  template <unsigned MASK>
  __device__ __forceinline__ void accumulate_mask(const double* Jt, int lane) {
    const int r4 = lane & 3, c8 = lane >> 2;
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      double a[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) a[t] = (MASK & (1u << t)) ? Jt[(8 * t + c8) * kJtStride + 4 * ks + r4] : 0.0;
#pragma unroll
      for (int ta = 0; ta < NT; ta++)
#pragma unroll
        for (int tb = ta; tb < NT; tb++)
          if ((MASK & (1u << ta)) && (MASK & (1u << tb))) {
            const int idx = ta * NT - ta * (ta - 1) / 2 + (tb - ta);
            dmma884(c[idx][0], c[idx][1], a[ta], a[tb]);
          }
    }
  }
  __device__ __forceinline__ void store(double* rec, int lane) const {
#pragma unroll
    for (int t = 0; t < n_tiles(NT); t++) {
      double2 v = make_double2(c[t][0], c[t][1]);
      *reinterpret_cast<double2*>(rec + t * 64 + 2 * lane) = v;  // row-major 8x8: (lane/4)*8 + (lane%4)*2
    }
    if (EXTRA) rec[n_tiles(NT) * 64 + lane] = extra;
  }
  __device__ __forceinline__ void atomic_add(double* dst, int lane) const {
#pragma unroll
    for (int t = 0; t < n_tiles(NT); t++) {
      atomicAdd(dst + t * 64 + 2 * lane, c[t][0]);
      atomicAdd(dst + t * 64 + 2 * lane + 1, c[t][1]);
    }
    if (EXTRA) atomicAdd(dst + n_tiles(NT) * 64 + lane, extra);
  }
};

__host__ __device__ __forceinline__ int tile_index(int NT, int ta, int tb) {  // ta <= tb
  return ta * NT - ta * (ta - 1) / 2 + (tb - ta);
}

// Shared-memory staging of the window: knots + PairData.  Returns the view; `rest` points past it.
__device__ __forceinline__ WindowView stage_window(double* smem, const PassParams& pp, double** rest) {
  const int K = pp.sl.K;
  double* kq = smem;
  double* kp = kq + 4 * K;
  PairData* pair = reinterpret_cast<PairData*>(kp + 3 * K);
  for (int i = threadIdx.x; i < 7 * K; i += blockDim.x) smem[i] = pp.state[i];  // off_q = 0, off_p = 4K contiguous
  __syncthreads();
  for (int k = threadIdx.x; k < K - 1; k += blockDim.x) compute_pair(kq + 4 * k, kq + 4 * (k + 1), &pair[k]);
  __syncthreads();
  WindowView W;
  W.kq = kq;
  W.kp = kp;
  W.pair = pair;
  size_t used = (size_t)7 * K + (size_t)kPairDoubles * (K - 1);
  used = (used + 1) & ~(size_t)1;  // keep what follows 16-byte aligned (double2 accesses)
  *rest = smem + used;
  return W;
}
__host__ __device__ inline size_t window_smem_doubles(int K) {
  return (((size_t)7 * K + (size_t)kPairDoubles * (K - 1)) + 1) & ~(size_t)1;
}

// A stream may accumulate in a COMPACT column basis and expand to the record layout everything downstream addresses
// only when an accumulator leaves the registers (once per interval and warp, not per factor): Xform::entry(rec, e) is
// element e of the standard record computed from the compact one.
struct NoXform {
  static constexpr bool kActive = false;
  __device__ __forceinline__ double entry(const double* rec, int e) const { return rec[e]; }
};

// DEFER (dense finish): a record that gets a slot stays compact — only this CTA reads it back, and its finish expands it
// with all threads instead of one warp inside the slab loop; the atomics fallback always takes expanded values.
template <int NT, int EXTRA = 0, class Xform = NoXform, bool DEFER = false>
__device__ __forceinline__ void flush_record(const WarpAcc<NT, EXTRA>& acc, int key, int& slot, int warp_global,
                                             const RecordBuf& rb, int lane, double* tmp = nullptr,
                                             const Xform& xf = Xform()) {
  if (key < 0) return;
  if (Xform::kActive && !(DEFER && slot < kSlotsPerWarp)) {  // tmp: rec_doubles(NT, EXTRA) doubles of the warp's own smem
    constexpr int RD = rec_doubles(NT, EXTRA);
    acc.store(tmp, lane);
    __syncwarp();
    const bool has_slot = slot < kSlotsPerWarp;
    const int r = warp_global * kSlotsPerWarp + slot;
    double* dst = has_slot ? rb.payload + (size_t)r * RD : rb.overflow + (size_t)key * RD;
    for (int e = lane; e < RD; e += 32) {
      const double v = xf.entry(tmp, e);
      if (has_slot) dst[e] = v; else atomicAdd(dst + e, v);
    }
    if (lane == 0) {
      if (has_slot) rb.key[r] = key; else *rb.overflow_flag = 1;
    }
    if (has_slot) slot++;
    __syncwarp();
    return;
  }
  if (slot < kSlotsPerWarp) {
    int r = warp_global * kSlotsPerWarp + slot;
    acc.store(rb.payload + (size_t)r * rec_doubles(NT, EXTRA), lane);
    if (lane == 0) rb.key[r] = key;
    slot++;
  } else {  // unsorted input: correct but unordered
    acc.atomic_add(rb.overflow + (size_t)key * rec_doubles(NT, EXTRA), lane);
    if (lane == 0) *rb.overflow_flag = 1;
  }
}

// One descriptor per factor stream whose records were reduced to per-key blocks.
struct StreamDesc {
  const double* fin;     // [n_keys][rd] per-key blocks of reduce_all_kernel (atomics-fallback block folded in)
  int n_keys, NT, rd;
  int r_col;             // local column holding r~ (when b lives in the tiles)
  int b_off;             // >= 0: b = J~^T r~ is stored as a plain vector at this offset of the record (LiDAR)
  int n_extra;           // local columns extra_base .. extra_base+n_extra-1 map to extra_global[] (-1: constant)
  int extra_base;        // 24 (one knot window) or 48 (two windows)
  int kind;              // 0: key = interval s (LiDAR / IMU);  1: key = s_i * kdim + s_j (visual, two knot windows)
  int kdim;
  int extra_global[12];
  unsigned char perm[64]; // logical local column -> physical column inside the record tiles
};

struct ColMaps {
  int D, K;
  const int* knot_col;   // [K] column of knot's dtheta (dp = +3), -1 if constant
  const int* col_knot;   // [D] knot of a column (-1: not a knot column)
  const int* col_sub;    // [D] 0..5 within the knot
};


__host__ __device__ __forceinline__ int block_offset(int NT, int la, int lb) {  // la, lb: PHYSICAL columns of the record tiles
  int ta = la >> 3, tb = lb >> 3;
  if (ta > tb) { int t = ta; ta = tb; tb = t; t = la; la = lb; lb = t; }
  return tile_index(NT, ta, tb) * 64 + (la & 7) * 8 + (lb & 7);
}

// ---- dense finish (pass_kernel, pass_fused.cuh): instead of emitting records for a later reduction, a CTA adds what it
// accumulated straight into ITS OWN copy of the packed system [upper triangle of H row-major | b | cost, fixed cost]
// in shared memory and writes that out contiguously; after one grid barrier the assembly is a sum over CTAs of
// identically laid out vectors (coalesced, no keys to look up).  Everything stays deterministic: groups of warps with
// the same interval are summed element-wise in warp order, then every target entry of the packed system is owned by
// one thread, which pulls its <= 4 terms from the group's record in a fixed order.
struct DenseOut {
  double* Hs;              // shared memory: the CTA's packed system, n_ent + 2 doubles, zeroed at kernel start
  int n_ent;               // D (D + 1) / 2 + D
  int D;
  const StreamDesc* sd;    // this stream's column descriptor (a shared-memory copy: it is read per target entry)
  const int* knot_col;     // [K] column of the knot's dtheta, -1 if constant (shared-memory copy)
  int* key_cols;           // shared memory scratch: [1 + 64 * 3] ints (see build_key_cols)
  unsigned long long* stamp;  // null, or this CTA's phase stamps: [6] = warp 0 starts its slabs, [7] = warp 0 is through them
  int slab_lo, slab_hi;       // this CTA's slabs of its stream (the fused pass balances them per CTA, build_fused)
};
// The slabs of one warp: an even split of the CTA's range when the caller fixed one, else of the whole stream over the
// warps of the grid.
__device__ __forceinline__ void warp_slab_range(const DenseOut* dn, int total_slabs, int warp_global, int warp, int warps_per_cta,
                                                int64_t n_warps_grid, int64_t* slab0, int* n_slabs) {
  if (dn) {
    const int64_t len = dn->slab_hi - dn->slab_lo;
    *slab0 = dn->slab_lo + len * warp / warps_per_cta;
    *n_slabs = (int)(dn->slab_lo + len * (warp + 1) / warps_per_cta - *slab0);
  } else {
    *slab0 = (int64_t)total_slabs * warp_global / n_warps_grid;
    *n_slabs = (int)((int64_t)total_slabs * (warp_global + 1) / n_warps_grid - *slab0);
  }
}
__device__ __forceinline__ void dense_stamp(const DenseOut* dn, int i) {
  if (dn && dn->stamp && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    dn->stamp[i] = t;
  }
}
__device__ __forceinline__ int packed_row_start(int D, int gi) { return gi * D - gi * (gi - 1) / 2; }

// The distinct global columns a record of interval key `key` touches, ascending, each with its (<= 2) local logical
// columns: kc[0] = U, then U triples (global column, local candidate 0, local candidate 1 or -1).  Built by warp 0.
__device__ __forceinline__ void build_key_cols(const StreamDesc& sd, const int* knot_col, int key, int* kc, int lane) {
  int base = 0;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    const int i = half * 32 + lane;
    int g = -1, c0 = -1, c1 = -1;
    if (sd.kind == 0) {
      if (i < 24) {
        const int kcol = knot_col[key + i / 6];
        if (kcol >= 0) { g = kcol + i % 6; c0 = i; }
      } else if (i < 24 + sd.n_extra) {
        g = sd.extra_global[i - 24];
        c0 = sd.extra_base + (i - 24);
      }
    } else {
      const int si = key / sd.kdim, sj = key - si * sd.kdim;
      const int lo = min(si, sj), hi = max(si, sj);
      if (i < 48) {  // candidate knots lo..lo+3, then hi..hi+3 (those not among the first four): ascending, unique
        const int m = i / 6, c = i % 6;
        const int k = m < 4 ? lo + m : hi + (m - 4);
        if (m < 4 || k > lo + 3) {
          const int kcol = knot_col[k];
          if (kcol >= 0) {
            g = kcol + c;
            const int di = k - si, dj = k - sj;
            const int a = (di >= 0 && di <= 3) ? 6 * di + c : -1, b = (dj >= 0 && dj <= 3) ? 24 + 6 * dj + c : -1;
            c0 = a >= 0 ? a : b;
            c1 = a >= 0 ? b : -1;
          }
        }
      } else if (i < 48 + sd.n_extra) {
        g = sd.extra_global[i - 48];
        c0 = sd.extra_base + (i - 48);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, g >= 0);
    if (g >= 0) {
      const int pos = base + __popc(m & ((1u << lane) - 1u));
      kc[1 + 3 * pos] = g;
      kc[2 + 3 * pos] = c0;
      kc[3 + 3 * pos] = c1;
    }
    base += __popc(m);
  }
  if (lane == 0) kc[0] = base;
}

// rec (shared or global memory): one record of this stream (tiles [+ b vector]); adds it into dn.Hs.  All threads of the
// CTA call; two block barriers inside.
template <int NTHREADS>
__device__ __noinline__ void dense_scatter_record(const DenseOut& dn, const double* rec, int key) {
  const StreamDesc& sd = *dn.sd;
  __syncthreads();  // key_cols of the previous record are no longer read; rec is complete
  if (threadIdx.x < 32) build_key_cols(sd, dn.knot_col, key, dn.key_cols, threadIdx.x);
  __syncthreads();
  const int* kc = dn.key_cols;
  const int U = kc[0];
  const int n_pairs = U * (U + 1) / 2;
  const int D = dn.D, n_upper = D * (D + 1) / 2;
  const int NT = sd.NT;
  for (int t = threadIdx.x; t < n_pairs + U; t += NTHREADS) {
    double v = 0.0;
    int dst;
    if (t < n_pairs) {  // unrank u <= v over U (float estimate, exact integer correction)
      const float q = 2.0f * U + 1.0f;
      int u = (int)((q - sqrtf(q * q - 8.0f * (float)t)) * 0.5f);
      u = max(0, min(U - 1, u));
      while (u > 0 && u * U - u * (u - 1) / 2 > t) u--;
      while ((u + 1) * U - (u + 1) * u / 2 <= t) u++;
      const int w = u + (t - (u * U - u * (u - 1) / 2));
      const int gi = kc[1 + 3 * u], gj = kc[1 + 3 * w];
#pragma unroll
      for (int x = 0; x < 2; x++) {
        const int la = kc[2 + x + 3 * u];
        if (la < 0) continue;
#pragma unroll
        for (int y = 0; y < 2; y++) {
          const int lb = kc[2 + y + 3 * w];
          if (lb < 0) continue;
          v += rec[block_offset(NT, sd.perm[la], sd.perm[lb])];
        }
      }
      dst = packed_row_start(D, gi) + (gj - gi);
    } else {
      const int u = t - n_pairs;
#pragma unroll
      for (int x = 0; x < 2; x++) {
        const int la = kc[2 + x + 3 * u];
        if (la < 0) continue;
        v += sd.b_off >= 0 ? rec[sd.b_off + la] : rec[block_offset(NT, sd.perm[la], sd.perm[sd.r_col])];
      }
      dst = n_upper + kc[1 + 3 * u];
    }
    dn.Hs[dst] += v;
  }
}

// End of a factor body in dense mode (replaces cta_merge_and_flush + the per-warp cost words).
template <int NT, int EXTRA, int WARPS, class Xform = NoXform>
__device__ __forceinline__ void cta_finish_dense(const WarpAcc<NT, EXTRA>& acc, int acc_key, int slot, int warp, int lane,
                                                 int warp_global, const RecordBuf& rb, double* scratch, const DenseOut& dn,
                                                 double cost, double fcost, bool jac, const Xform& xf = Xform()) {
  constexpr int RD = rec_doubles(NT, EXTRA);
  __shared__ int s_keys[WARPS], s_slot[WARPS];
  __shared__ double s_cost[WARPS][2];
  __syncthreads();  // the Jt tiles under `scratch` are dead for every warp
  if (jac && acc_key >= 0) acc.store(scratch + (size_t)warp * RD, lane);
  if (lane == 0) {
    s_keys[warp] = jac ? acc_key : -1;
    s_slot[warp] = slot;
    s_cost[warp][0] = cost;
    s_cost[warp][1] = fcost;
  }
  __syncthreads();
  if (threadIdx.x < 2) {  // the CTA's cost words: warps in order
    double c = 0.0;
#pragma unroll
    for (int w = 0; w < WARPS; w++) c += s_cost[w][threadIdx.x];
    dn.Hs[dn.n_ent + threadIdx.x] = c;  // a CTA runs one stream: plain store
  }
  if (!jac) return;
  // groups of warps with the same key, in order of their first warp: element-wise sum in warp order into the first
  // warp's slice, then one scatter per group
  for (int w0 = 0; w0 < WARPS; w0++) {
    const int kc = s_keys[w0];
    if (kc < 0) continue;
    bool first = true;
    for (int w = 0; w < w0; w++) first = first && s_keys[w] != kc;
    if (!first) continue;
    bool alone = true;
    for (int w = w0 + 1; w < WARPS; w++) alone = alone && s_keys[w] != kc;
    if (!alone) {
      for (int e = threadIdx.x; e < RD; e += WARPS * 32) {
        double a = scratch[(size_t)w0 * RD + e];
        for (int w = w0 + 1; w < WARPS; w++)
          if (s_keys[w] == kc) a += scratch[(size_t)w * RD + e];
        scratch[(size_t)w0 * RD + e] = a;
      }
    }
    if (Xform::kActive) {  // expand the (merged) compact record behind the parked ones, scatter that
      double* xbuf = scratch + (size_t)WARPS * RD;
      __syncthreads();  // the merge is complete; the previous scatter no longer reads xbuf
      for (int e = threadIdx.x; e < RD; e += WARPS * 32) xbuf[e] = xf.entry(scratch + (size_t)w0 * RD, e);
      dense_scatter_record<WARPS * 32>(dn, xbuf, kc);
    } else {
      dense_scatter_record<WARPS * 32>(dn, scratch + (size_t)w0 * RD, kc);
    }
  }
  // records flushed in mid-run (a warp whose interval changed): read back from this CTA's own slots, in (warp, slot)
  // order, through the (now free) parking area — every load of a record in flight at once; scattering straight from
  // global memory would pay an L2 round trip per term (in-order issue)
  for (int w = 0; w < WARPS; w++)
    for (int sl = 0; sl < min(s_slot[w], kSlotsPerWarp); sl++) {
      const int r = (warp_global - warp + w) * kSlotsPerWarp + sl;
      const double* src = rb.payload + (size_t)r * RD;
      constexpr int PER = (RD + WARPS * 32 - 1) / (WARPS * 32);
      double v[PER];
      __syncthreads();  // the previous scatter no longer reads the staging area
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int e = q * WARPS * 32 + threadIdx.x;
        v[q] = e < RD ? src[e] : 0.0;
      }
      const int key = rb.key[r];
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int e = q * WARPS * 32 + threadIdx.x;
        if (e < RD) scratch[e] = v[q];
      }
      if (Xform::kActive) {  // stored compact (flush_record<..., DEFER>): expand next to it
        double* xbuf = scratch + (size_t)WARPS * RD;
        __syncthreads();
        for (int e = threadIdx.x; e < RD; e += WARPS * 32) xbuf[e] = xf.entry(scratch, e);
        dense_scatter_record<WARPS * 32>(dn, xbuf, key);
      } else {
        dense_scatter_record<WARPS * 32>(dn, scratch, key);
      }
    }
  __syncthreads();
}

// End of a factor kernel: the live accumulators of the CTA's warps that share the key of the first live warp (all
// of them, for time-sorted input) become ONE record; the others are flushed individually.  8x fewer records for the
// reduction stage, no atomics, fixed order.  Every warp first parks its accumulator in its own slice of `scratch`
// (WARPS * rec_doubles doubles; may alias the Jt tiles, dead by now) in parallel, then all threads of the CTA sum
// their elements over the matching warps in warp order and write the record straight to global memory — two block
// barriers instead of one per warp (the visual record is 14 KB per warp: the serial variant cost ~8 us there).
template <int NT, int EXTRA = 0, int WARPS = kWarpsPerCta, class Xform = NoXform>
__device__ __forceinline__ void cta_merge_and_flush(const WarpAcc<NT, EXTRA>& acc, int acc_key, int& slot, int warp, int lane,
                                                    int warp_global, const RecordBuf& rb, double* scratch,
                                                    const Xform& xf = Xform()) {
  constexpr int RD = rec_doubles(NT, EXTRA);
  __shared__ int s_keys[WARPS], s_slot[WARPS];
  __syncthreads();  // the Jt tiles under `scratch` are dead for every warp
  if (acc_key >= 0) acc.store(scratch + (size_t)warp * RD, lane);
  if (lane == 0) { s_keys[warp] = acc_key; s_slot[warp] = slot; }
  __syncthreads();
  int kc = -1, w_first = -1;
#pragma unroll
  for (int w = WARPS - 1; w >= 0; w--)
    if (s_keys[w] >= 0) { kc = s_keys[w]; w_first = w; }
  if (kc < 0) return;
  if (acc_key >= 0 && acc_key != kc)  // (its own parked slice doubles as the staging area: nobody else reads it)
    flush_record<NT, EXTRA, Xform>(acc, acc_key, slot, warp_global, rb, lane, scratch + (size_t)warp * RD, xf);
  const int slot_f = s_slot[w_first];
  const bool has_slot = slot_f < kSlotsPerWarp;
  const int r = (warp_global - warp + w_first) * kSlotsPerWarp + slot_f;
  double* dst = has_slot ? rb.payload + (size_t)r * RD : rb.overflow + (size_t)kc * RD;
  if (Xform::kActive) {  // merge in the compact basis behind the parked records, expand on the way out
    double* merged = scratch + (size_t)WARPS * RD;
    for (int e = threadIdx.x; e < RD; e += WARPS * 32) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < WARPS; w++)
        if (s_keys[w] == kc) a += scratch[(size_t)w * RD + e];
      merged[e] = a;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < RD; e += WARPS * 32) {
      const double v = xf.entry(merged, e);
      if (has_slot) dst[e] = v; else atomicAdd(dst + e, v);
    }
  } else {
    for (int e = threadIdx.x; e < RD; e += WARPS * 32) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < WARPS; w++)
        if (s_keys[w] == kc) a += scratch[(size_t)w * RD + e];
      if (has_slot) dst[e] = a; else atomicAdd(dst + e, a);
    }
  }
  if (warp == w_first) {
    if (has_slot) {
      if (lane == 0) rb.key[r] = kc;
      slot++;
    } else if (lane == 0) {
      *rb.overflow_flag = 1;
    }
  }
}

// Lane L returns the sum over the warp of v[L] (v[24..31] are implicit zeros): the classic transpose-reduce, 31
// add + shuffle steps instead of 24 x 5.
__device__ __forceinline__ double warp_transpose_reduce24(const double (&v)[24], int lane) {
  double a[16];
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const double lo_v = v[i], hi_v = (i + 16 < 24) ? v[i + 16] : 0.0;
      const double send = hi ? lo_v : hi_v, keep = hi ? hi_v : lo_v;
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  double b8[8];
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double send = hi ? a[i] : a[i + 8], keep = hi ? a[i + 8] : a[i];
      b8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  double c4[4];
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const double send = hi ? b8[i] : b8[i + 4], keep = hi ? b8[i + 4] : b8[i];
      c4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  double d2[2];
  {
    const bool hi = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      const double send = hi ? c4[i] : c4[i + 2], keep = hi ? c4[i + 2] : c4[i];
      d2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool hi = lane & 1;
  const double send = hi ? d2[0] : d2[1], keep = hi ? d2[1] : d2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// K1: LoamFeatureFactor over a stream (lidar_feature_factor.h:65-159), fused with the normal-equation accumulation.
// Local columns: [knot s..s+3: dtheta(3) dp(3)] = 0..23.  J~^T J~ goes through 6 DMMA tiles (NT = 3); J~^T r~ is a
// warp transpose-reduce (FP64 DMMA and DFMA share one pipe on B200, so the 4th tile column would cost 3x more pipe
// time than the 24 multiplies + 31 adds); sum r~^2 is the cost.  Evaluation is divergence-free: every lane evaluates,
// masked lanes with a zero scale.  The loads of the next slab are issued before the DMMA phase of the current one.
// GEO: 0 = planes and lines mixed in one stream (type byte per feature), 1 = planes only (4 geometry doubles, the
// line code is compiled out), 2 = lines only (no type byte; no intra-warp divergence between the two residuals).
// The body is shared by loam_kernel (one stream per launch) and pass_kernel (pass_fused.cuh: every stream of a pass in
// one launch, this stream on CTAs cta = 0 .. n_ctas-1 of its partition).
template <bool JAC, int WARPS, int GEO, bool DENSE = false>
__device__ __forceinline__ void loam_body(const LoamStream& st, const PassParams& pp, int total_slabs, const RecordBuf& rb,
                                          const WindowView& W, double* rest, int cta, int n_ctas,
                                          const DenseOut* dn = nullptr) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_global = cta * WARPS + warp;
  double* Jt = rest + (size_t)warp * 24 * kJtStride;
  WarpAcc<3, 1> acc;
  acc.zero();
  int acc_key = -1, slot = 0;
  double cost = 0.0, fcost = 0.0;
  // balanced contiguous split of the slabs over all warps of the grid (sizes differ by at most one)
  const int64_t n_warps_grid = (int64_t)n_ctas * WARPS;
  int64_t slab0;
  int slabs_per_warp;
  warp_slab_range(DENSE ? dn : nullptr, total_slabs, warp_global, warp, WARPS, n_warps_grid, &slab0, &slabs_per_warp);
  // software pipeline registers (next slab)
  int64_t n_t = kDropped;
  double n_p0 = 0, n_p1 = 0, n_p2 = 0, n_g0 = 0, n_g1 = 0, n_g2 = 0, n_g3 = 0, n_g4 = 0, n_g5 = 0;
  unsigned char n_ty = 1;
  auto prefetch = [&](int64_t idx) {
    n_t = kDropped;
    if (idx < st.n) {
      n_t = st.t_ns[idx];
      n_p0 = st.p[0][idx]; n_p1 = st.p[1][idx]; n_p2 = st.p[2][idx];
      n_g0 = st.g[0][idx]; n_g1 = st.g[1][idx]; n_g2 = st.g[2][idx];
      n_g3 = st.g[3][idx];
      if (GEO != 1) { n_g4 = st.g[4][idx]; n_g5 = st.g[5][idx]; }
      if (GEO == 0) n_ty = st.type[idx];
    }
  };
#ifndef CLIC_LOAM_PREFETCH
#define CLIC_LOAM_PREFETCH 1
#endif
  if (CLIC_LOAM_PREFETCH) prefetch(slab0 * 32 + lane);
  if (DENSE) dense_stamp(dn, 6);
  for (int sl = 0; sl < slabs_per_warp; sl++) {
    if ((slab0 + sl) * 32 >= st.n) break;
    if (!CLIC_LOAM_PREFETCH) prefetch((slab0 + sl) * 32 + lane);  // plain load at use (more resident warps hide it)
    const int64_t next_idx = (CLIC_LOAM_PREFETCH && sl + 1 < slabs_per_warp) ? (slab0 + sl + 1) * 32 + lane : st.n;
    const int64_t t = n_t;
    int s = 0;
    double u = 0.0;
    bool valid = (t != kDropped) && index_ns(t, pp.t0_ns, pp.dt_ns, pp.sl.K, &s, &u);
    // masked lanes evaluate a harmless dummy (plane with zero normal at s = 0)
    const V3 pL = valid ? mk(n_p0, n_p1, n_p2) : mk(0, 0, 0);
    double geo[6];
    geo[0] = valid ? n_g0 : 0.0; geo[1] = valid ? n_g1 : 0.0; geo[2] = valid ? n_g2 : 0.0;
    geo[3] = valid ? n_g3 : 0.0; geo[4] = valid ? n_g4 : 0.0; geo[5] = valid ? n_g5 : 0.0;
    const bool is_plane = GEO == 1 ? true : (valid ? (GEO == 2 ? false : (n_ty != 0)) : true);
    if (!valid) { s = 0; u = 0.0; }
    if (pp.marg_mode && valid && s >= pp.marg_later_local) valid = false;  // empty drop set: not part of the prior
    const int key = valid ? s : -1;
    bool prefetched = false;
    if (!JAC) { prefetch(next_idx); prefetched = true; }
    // distinct keys of the slab in ascending order (one, for time-sorted input)
    unsigned pending = __ballot_sync(0xffffffffu, key >= 0);
    while (pending) {
      const int kmin = __reduce_min_sync(0xffffffffu, (key >= 0 && ((pending >> lane) & 1)) ? key : 0x7fffffff);
      const bool mine = (key == kmin) && ((pending >> lane) & 1);
      double r;
      // every lane evaluates; the scale masks it out.  The 24 Jacobian columns are streamed to the warp's transposed
      // shared tile as they are produced (short register live ranges).
      struct SmemOut24 {
        double* Jt;
        __device__ __forceinline__ void put(int c, double v) { Jt[c * kJtStride] = v; }
      } jout;
      jout.Jt = Jt + lane;
      loam_eval_to(W, s, u, pL, is_plane, geo, st.sh, JAC, &r, jout, mine ? st.sh.weight : 0.0);
      // marg gate (trajectory_estimator.cpp:729-741): only |r| / w < 0.05 enters the prior
      bool contributes = mine;
      if (pp.marg_mode && mine && !(fabs(r / st.sh.weight) < 0.05)) contributes = false;
      const double r_eff = contributes ? r : 0.0;
      if (s + 3 >= pp.kf || pp.marg_mode) cost += 0.5 * r_eff * r_eff; else fcost += 0.5 * r_eff * r_eff;
      if (JAC) {
        if (kmin != acc_key) {
          flush_record<3, 1>(acc, acc_key, slot, warp_global, rb, lane);
          acc.zero();
          acc_key = kmin;
        }
        double pb[24];
        if (pp.marg_mode && contributes != mine) {  // gated out after the fact: zero the row
#pragma unroll
          for (int c = 0; c < 24; c++) Jt[c * kJtStride + lane] = 0.0;
        }
#pragma unroll
        for (int c = 0; c < 24; c++) pb[c] = Jt[c * kJtStride + lane] * r_eff;
        acc.extra += warp_transpose_reduce24(pb, lane);
        __syncwarp();
        if (!prefetched) { prefetch(next_idx); prefetched = true; }
        acc.accumulate(Jt, lane);
        __syncwarp();
      }
      pending &= ~__ballot_sync(0xffffffffu, mine);
    }
    if (!prefetched) prefetch(next_idx);
  }
  // deterministic butterfly
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(0xffffffffu, cost, o);
    fcost += __shfl_xor_sync(0xffffffffu, fcost, o);
  }
  if (DENSE) {
    dense_stamp(dn, 7);
    cta_finish_dense<3, 1, WARPS>(acc, acc_key, slot, warp, lane, warp_global, rb, rest, *dn, cost, fcost, JAC);
    return;
  }
  if (JAC) {
    cta_merge_and_flush<3, 1, WARPS>(acc, acc_key, slot, warp, lane, warp_global, rb, rest);
    for (int sidx = slot + lane; sidx < kSlotsPerWarp; sidx += 32) rb.key[warp_global * kSlotsPerWarp + sidx] = -1;
  }
  if (lane == 0) {
    rb.cost[2 * warp_global] = cost;
    rb.cost[2 * warp_global + 1] = fcost;
  }
}
template <bool JAC, int WARPS, int MINB, int GEO = 0>
__global__ void __launch_bounds__(WARPS * 32, MINB)
loam_kernel(LoamStream st, PassParams pp, int total_slabs, RecordBuf rb) {
  if (pass_prologue(pp)) return;
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  loam_body<JAC, WARPS, GEO>(st, pp, total_slabs, rb, W, rest, blockIdx.x, gridDim.x);
  pass_epilogue(pp);
}

// Fixed-tree sum of the per-warp (cost, fixed_cost) pairs by one 256-thread CTA; valid in thread 0.
__device__ __forceinline__ void cta_cost_sum(const double* warp_cost, int n_warps, double* c_out, double* f_out) {
  __shared__ double sh[2][256];
  double c = 0, f = 0;
  for (int i = threadIdx.x; i < n_warps; i += blockDim.x) {
    c += warp_cost[2 * i];
    f += warp_cost[2 * i + 1];
  }
  __syncthreads();  // sh may still be read by a previous call
  sh[0][threadIdx.x] = c;
  sh[1][threadIdx.x] = f;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  *c_out = sh[0][0];
  *f_out = sh[1][0];
}

// All factor streams of a pass in ONE launch: CTA (job, key) fetches every slot key of the stream with all loads
// in flight, compacts the hits in slot order (256 threads, two-level scan) and sums the matching payloads in that
// order, NB records in flight per thread; then folds the atomics-fallback block in and re-arms it.  For time-sorted streams a key has ~n_CTAs / n_keys records, so this is a handful of load rounds.  The extra
// last CTA sums every job's per-warp costs.  jac == 0: only that CTA is launched.
struct ReduceJob {
  const double* payload;
  const int* key;
  double* fin;       // [n_keys][rd]
  double* overflow;  // [n_keys][rd], folded into fin and zeroed again
  const double* warp_cost;
  int n_slots, n_keys, rd, n_warps;
  int cta0;          // first CTA of the job in the flattened grid
};
template <int RD>
__device__ __forceinline__ void reduce_key_body(const ReduceJob& J, int k, int* skey, int* list) {
  constexpr int EPT = (RD + 255) / 256;
  constexpr int NB = RD <= 640 ? 8 : 4;  // records in flight per thread
  __shared__ int warp_cnt[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_slots = J.n_slots;
  // every slot key of the stream, all loads in flight
  for (int r = threadIdx.x; r < n_slots; r += 256) skey[r] = J.key[r];
  __syncthreads();
  // ordered compaction: thread t owns the contiguous slots [t * per, (t + 1) * per); counts are scanned within the
  // warp (shuffles) and across the 8 warps (shared), then every thread writes its hits at its offset
  const int per = (n_slots + 255) / 256;
  const int s0 = min(n_slots, (int)threadIdx.x * per), s1 = min(n_slots, s0 + per);
  int cnt = 0;
  for (int r = s0; r < s1; r++) cnt += skey[r] == k ? 1 : 0;
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_cnt[warp] = incl;
  __syncthreads();
  int base = incl - cnt, n = 0;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    if (w < warp) base += warp_cnt[w];
    n += warp_cnt[w];
  }
  for (int r = s0; r < s1; r++)
    if (skey[r] == k) list[base++] = r;
  __syncthreads();
  double acc[EPT];
#pragma unroll
  for (int i = 0; i < EPT; i++) acc[i] = 0.0;
  for (int j = 0; j < n; j += NB) {
    double v[NB][EPT];
#pragma unroll
    for (int q = 0; q < NB; q++) {
      const bool on = j + q < n;
      const double* src = J.payload + (size_t)(on ? list[j + q] : 0) * RD;
#pragma unroll
      for (int i = 0; i < EPT; i++) {
        const int e = threadIdx.x + i * 256;
        v[q][i] = (on && e < RD) ? src[e] : 0.0;
      }
    }
#pragma unroll
    for (int q = 0; q < NB; q++)
#pragma unroll
      for (int i = 0; i < EPT; i++) acc[i] += v[q][i];
  }
  double* ov = J.overflow + (size_t)k * RD;
  double* fin = J.fin + (size_t)k * RD;
#pragma unroll
  for (int i = 0; i < EPT; i++) {
    const int e = threadIdx.x + i * 256;
    if (e < RD) {
      const double o = ov[e];
      if (o != 0.0) ov[e] = 0.0;
      fin[e] = acc[i] + o;
    }
  }
}
__global__ void __launch_bounds__(256, 3)
reduce_all_kernel(const ReduceJob* jobs, int n_jobs, int n_key_ctas, int max_slots, double* cost2, const int* ctl) {
  pdl_wait();
  pdl_trigger();
  if (ctl && ctl[0]) return;
  if ((int)blockIdx.x == n_key_ctas) {
    double c_tot = 0.0, f_tot = 0.0;
    for (int j = 0; j < n_jobs; j++) {
      double c, f;
      cta_cost_sum(jobs[j].warp_cost, jobs[j].n_warps, &c, &f);
      c_tot += c;
      f_tot += f;
    }
    if (threadIdx.x == 0) {
      cost2[0] = c_tot;
      cost2[1] = f_tot;
    }
    return;
  }
  int j = 0;
  while (j + 1 < n_jobs && (int)blockIdx.x >= jobs[j + 1].cta0) j++;
  const ReduceJob& J = jobs[j];
  const int k = blockIdx.x - J.cta0;
  extern __shared__ int red_smem[];  // [max n_slots] keys + [max n_slots] ordered hit list
  int* skey = red_smem;
  int* list = red_smem + max_slots;
  switch (J.rd) {
    case 416: reduce_key_body<416>(J, k, skey, list); break;
    case 640: reduce_key_body<640>(J, k, skey, list); break;
    case 960: reduce_key_body<960>(J, k, skey, list); break;
    default: reduce_key_body<1792>(J, k, skey, list); break;
  }
}

}  // namespace clic

namespace clic {

// ---------------- K2: IMUFactor over a stream ----------------
struct ImuStream {
  int64_t n;
  const int64_t* t_ns;  // measurement time (without offset); kDropped if dropped at add time
  const double* gyro[3];
  const double* accel[3];
  double info[6];
  double loss_a;
  int bias_slot;
  int extras_free;  // a bias / time-offset block of these factors is a free parameter (options, known at add time)
  int marg_always;  // marg mode: a non-knot block of these factors is dropped, so they all enter the prior
  int toff_free;    // the IMU time offset has a column in H (set per pass from the layout)
  const int4* xform_tab;  // ImuCompactXform's expansion table (solve mode)
};

template <int NT>
struct SmemRowSink {
  double* Jt;
  int lane;
  bool on;
  WarpAcc<NT>* acc;
  __device__ __forceinline__ void put(int, int col, double v) { Jt[col * kJtStride + lane] = on ? v : 0.0; }
  __device__ __forceinline__ void row_done(int) {
    __syncwarp();
    acc->accumulate(Jt, lane);
    __syncwarp();
  }
};

// IMU rows are stored in a PHYSICAL column order that packs the columns the gyro rows touch into the first three
// tiles:  [dtheta x12 | bg | r | t_off | ba | (gravity) | dp x12 | pad].  Gyro rows (0..2) have zero dp / ba /
// gravity blocks, so their last tile(s) are skipped: 6 instead of 10 (15 in marg mode) tile products — and only 3
// (two column tiles) when the IMU time offset is a constant (every shipped LIO call): its column is never assembled.
// Logical order (what factor_eval emits, what the assembly addresses): [knot k: dtheta dp] x4 | bg | ba | (g) | t_off | r.
template <bool MARG>
__host__ __device__ constexpr int imu_phys_col(int l) {
  if (l < 24) return (l % 6 < 3) ? (l / 6) * 3 + l % 6 : (MARG ? 23 : 20) + (l / 6) * 3 + (l % 6 - 3);
  if (l < 27) return 12 + (l - 24);                       // bg
  if (l < 30) return 17 + (l - 27);                       // ba
  if (MARG) {
    if (l < 33) return 20 + (l - 30);                     // gravity
    if (l == 33) return 16;                               // t_off
    if (l == 34) return 15;                               // r
    return l;                                             // pad 35..39
  }
  return l == 30 ? 16 : 15;                               // t_off, r
}
template <bool MARG>
struct ImuRowSink {
  static constexpr int NT = MARG ? 5 : 4;
  double* Jt;
  int lane;
  bool on;
  bool gyro2;  // t_off is constant: gyro rows only need the first two column tiles
  WarpAcc<NT>* acc;
  __device__ __forceinline__ void put(int row, int col, double v) {
    const int pc = imu_phys_col<MARG>(col);
    if (pc >= 24) {  // never read for gyro rows (row is a run-time value inside the evaluator's row loop)
      if (row >= 3) Jt[pc * kJtStride + lane] = on ? v : 0.0;
      return;
    }
    Jt[pc * kJtStride + lane] = on ? v : 0.0;
  }
  __device__ __forceinline__ void row_done(int row) {
    __syncwarp();
    if (row < 3) {
      if (gyro2) acc->template accumulate<2>(Jt, lane); else acc->template accumulate<3>(Jt, lane);
    } else {
      acc->template accumulate<NT>(Jt, lane);
    }
    __syncwarp();
  }
};
// ---- solve-mode IMU rows in a COMPACT column basis: 9 instead of 13 tile products per factor ----
// The accel rows' 12 position columns are lambda_k(u) * A with ONE 1x3 row A = s_I * row(R^-1) per residual row and
// lambda_k = inv_dt^2 * (1 - u, -2 + 3u, 1 - 3u, u) (second derivative of the cubic blend): rank 2 over the four knots.
// They are accumulated as 6 basis columns  B0 = lambda_0 A + lambda_3 A (= inv_dt^2 A),  B1 = lambda_3 A (= inv_dt^2 u A)
// and expanded (dp_k = M0[k] B0 + M1[k] B1, M0 = (1, -2, 1, 0), M1 = (-1, 3, -3, 1)) when an accumulator leaves the
// registers.  Compact physical layout, chosen so that a row type touches as few column tiles as possible:
//   tile 0: dtheta 0..7        tile 1: dtheta 8..11 | r | bg        (gyro rows)
//   tile 2: dtheta 8..11 | r | ba   (accel rows)                    tile 3: B0 | B1 | t_off | 0
// gyro rows multiply tiles {0, 1} (+ 3 when the IMU time offset is free), accel rows tiles {0, 2, 3}; the two copies of
// dtheta 8..11 and of r are summed by the expansion, tile (1, 2) is never touched.  The expansion is exact algebra; the
// result differs from the direct accumulation by rounding only (lambda_0 + lambda_3 vs inv_dt^2, order of the sums).
struct ImuCompactSink {
  double* Jt;
  int lane;
  bool on;
  bool toff_free;
  WarpAcc<4>* acc;
  __device__ __forceinline__ void put(int row, int col, double v) {
    v = on ? v : 0.0;
    double* J = Jt + lane;
    if (col < 24) {
      const int k = col / 6, c = col % 6;
      if (c < 3) {
        const int i = 3 * k + c;
        J[(i < 8 ? i : (row < 3 ? i : i + 8)) * kJtStride] = v;
      } else if (row >= 3 || toff_free) {  // (gyro rows: zeros, only read when tile 3 takes part)
        if (k == 0) J[(24 + c - 3) * kJtStride] = v;
        if (k == 3) { J[(27 + c - 3) * kJtStride] = v; J[(24 + c - 3) * kJtStride] += v; }
      }
    } else if (col < 27) {
      if (row < 3) J[(13 + col - 24) * kJtStride] = v;
    } else if (col < 30) {
      if (row >= 3) J[(21 + col - 27) * kJtStride] = v;
    } else if (col == 30) {
      J[30 * kJtStride] = v;
    } else {
      J[(row < 3 ? 12 : 20) * kJtStride] = v;
    }
  }
  __device__ __forceinline__ void row_done(int row) {
    __syncwarp();
    if (row < 3) {
      if (toff_free) acc->template accumulate_mask<0xBu>(Jt, lane); else acc->template accumulate_mask<0x3u>(Jt, lane);
    } else {
      acc->template accumulate_mask<0xDu>(Jt, lane);
    }
    __syncwarp();
  }
};
// standard physical column (imu_phys_col<false>) -> logical column
__host__ __device__ __forceinline__ int imu_std_logical(int p) {
  if (p < 12) return (p / 3) * 6 + p % 3;
  if (p < 15) return 24 + (p - 12);
  if (p == 15) return 31;
  if (p == 16) return 30;
  if (p < 20) return 27 + (p - 17);
  const int j = p - 20;
  return (j / 3) * 6 + 3 + j % 3;
}
// a logical column as a combination of compact columns: returns the number of terms
__host__ __device__ __forceinline__ int imu_compact_terms(int l, int col[2], double w[2]) {
  w[0] = w[1] = 1.0;
  if (l < 24) {
    const int k = l / 6, c = l % 6;
    if (c < 3) {
      const int i = 3 * k + c;
      col[0] = i;
      col[1] = i + 8;
      return i < 8 ? 1 : 2;
    }
    col[0] = 24 + c - 3;
    col[1] = 27 + c - 3;
    w[0] = k == 0 ? 1.0 : (k == 1 ? -2.0 : (k == 2 ? 1.0 : 0.0));
    w[1] = k == 0 ? -1.0 : (k == 1 ? 3.0 : (k == 2 ? -3.0 : 1.0));
    return 2;
  }
  if (l < 27) { col[0] = 13 + l - 24; return 1; }
  if (l < 30) { col[0] = 21 + l - 27; return 1; }
  if (l == 30) { col[0] = 30; return 1; }
  col[0] = 12;
  col[1] = 20;
  return 2;
}
// Every element of the standard record is a combination of <= 4 elements of the compact one with small integer
// weights: a table (offset << 8 | weight + 128, four terms per element; weight 0 = unused) built once on the host
// (ImuStream::xform_tab, 10 KB of global memory), so that an expansion is one 16-byte load + 4 loads + 4 FMAs per element.
constexpr int kImuRecDoubles = 640;  // rec_doubles(4, 0)
struct ImuCompactXform {
  static constexpr bool kActive = true;
  const int4* tab;
  __device__ __forceinline__ double entry(const double* cmp, int e) const {
    const int4 q = __ldg(tab + e);
    double v = (double)((q.x & 255) - 128) * cmp[q.x >> 8];
    v += (double)((q.y & 255) - 128) * cmp[q.y >> 8];
    v += (double)((q.z & 255) - 128) * cmp[q.z >> 8];
    v += (double)((q.w & 255) - 128) * cmp[q.w >> 8];
    return v;
  }
  static void build_host(int* tab /*[kImuRecDoubles * 4]*/) {
    for (int e = 0; e < kImuRecDoubles; e++) {
      const int t = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
      const int ta = t < 4 ? 0 : (t < 7 ? 1 : (t < 9 ? 2 : 3));
      const int tb = t < 4 ? t : (t < 7 ? t - 3 : (t < 9 ? t - 5 : 3));
      int ca[2], cb[2];
      double wa[2], wb[2];
      const int na = imu_compact_terms(imu_std_logical(8 * ta + rr), ca, wa);
      const int nb = imu_compact_terms(imu_std_logical(8 * tb + cc), cb, wb);
      int term[4] = {128, 128, 128, 128};  // offset 0, weight 0
      int n = 0;
      for (int x = 0; x < na; x++)
        for (int y = 0; y < nb; y++) {
          const int w = (int)(wa[x] * wb[y]);
          if (w != 0) term[n++] = (block_offset(4, ca[x], cb[y]) << 8) | (w + 128);
        }
      for (int i = 0; i < 4; i++) tab[4 * e + i] = term[i];
    }
  }
};

struct NullSink {
  __device__ __forceinline__ void put(int, int, double) {}
  __device__ __forceinline__ void row_done(int) {}
};

__device__ __forceinline__ void imu_xform_setup(NoXform&, const int4*) {}
__device__ __forceinline__ void imu_xform_setup(ImuCompactXform& xf, const int4* tab) { xf.tab = tab; }
__device__ __forceinline__ void imu_sink_mode(ImuRowSink<true>& s, bool) { s.gyro2 = false; }
__device__ __forceinline__ void imu_sink_mode(ImuRowSink<false>& s, bool toff_free) { s.gyro2 = !toff_free; }
__device__ __forceinline__ void imu_sink_mode(ImuCompactSink& s, bool toff_free) { s.toff_free = toff_free; }

// Local columns (solve, NT=4): [knots 0..23 | bg 24..26 | ba 27..29 | toff 30 | r 31]
//               (marg,  NT=5): [knots 0..23 | bg | ba | gravity 30..32 | toff 33 | r 34 | pad]
template <bool JAC, bool MARG, bool DENSE = false>
__device__ __forceinline__ void imu_body(const ImuStream& st, const PassParams& pp, int total_slabs, const RecordBuf& rb,
                                         const WindowView& W, double* rest, int cta, int n_ctas,
                                         const DenseOut* dn = nullptr) {
  constexpr int NT = MARG ? 5 : 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_global = cta * kWarpsPerCta + warp;
  double* Jt = rest + (size_t)warp * (8 * NT) * kJtStride;
  ImuShared sh;
#pragma unroll
  for (int i = 0; i < 6; i++) sh.info[i] = st.info[i];
  const double* bias = pp.state + pp.sl.off_bias() + 6 * st.bias_slot;
  sh.bg = mk(bias[0], bias[1], bias[2]);
  sh.ba = mk(bias[3], bias[4], bias[5]);
  const double* gr = pp.state + pp.sl.off_grav();
  sh.gravity = mk(gr[0], gr[1], gr[2]);
  sh.loss_a = st.loss_a;
  sh.inv_dt = 1e9 / double(pp.dt_ns);
  sh.want_toff = MARG || st.toff_free;
  const int64_t toff = (int64_t)pp.state[pp.sl.off_toff() + 0];  // (int64_t)time_offset_in_ns, trajectory_value_factor.h:175
  // The compact accumulation basis pays where the expansion is done by the whole CTA (dense finish of the fused pass);
  // the per-stream kernels, which serve the small windows, keep the direct layout (an expansion per warp and record cost
  // them 2-3 us per pass).
  constexpr bool CMP = !MARG && DENSE;
  using Xf = typename std::conditional<CMP, ImuCompactXform, NoXform>::type;
  Xf xf;
  imu_xform_setup(xf, st.xform_tab);
  if (CMP && JAC) {
    if (threadIdx.x < (kImuRecDoubles * 16 + 127) / 128)  // the table is read at the very end: have it in L2 by then
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(st.xform_tab) + 128 * threadIdx.x));
    Jt[31 * kJtStride + lane] = 0.0;  // the one column of the compact layout nobody writes
    __syncwarp();
  }

  WarpAcc<NT> acc;
  acc.zero();
  int acc_key = -1, slot = 0;
  double cost = 0.0, fcost = 0.0;
  const int64_t n_warps_grid = (int64_t)n_ctas * kWarpsPerCta;
  int64_t slab0;
  int slabs_per_warp;
  warp_slab_range(DENSE ? dn : nullptr, total_slabs, warp_global, warp, kWarpsPerCta, n_warps_grid, &slab0, &slabs_per_warp);
  if (DENSE) dense_stamp(dn, 6);
  for (int sl = 0; sl < slabs_per_warp; sl++) {
    if ((slab0 + sl) * 32 >= st.n) break;
    const int64_t idx = (slab0 + sl) * 32 + lane;
    int64_t t = kDropped;
    if (idx < st.n) t = st.t_ns[idx];
    int s = -1;
    double u = 0.0;
    bool valid = (t != kDropped) && index_ns(t + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &s, &u);
    V3 gy = mk(0, 0, 0), ac = mk(0, 0, 0);
    if (valid) {
      gy = mk(st.gyro[0][idx], st.gyro[1][idx], st.gyro[2][idx]);
      ac = mk(st.accel[0][idx], st.accel[1][idx], st.accel[2][idx]);
    }
    // marg set membership: the drop set is non-empty iff an extra block is dropped (marg_always) or one of the
    // knots is older than ctrl_to_be_opt_later (trajectory_estimator.cpp:208-232)
    if (MARG && valid && !st.marg_always && s >= pp.marg_later_local) valid = false;
    int key = valid ? s : -1;
    unsigned pending = __ballot_sync(0xffffffffu, key >= 0);
    while (pending) {
      int kmin = __reduce_min_sync(0xffffffffu, (key >= 0 && ((pending >> lane) & 1)) ? key : 0x7fffffff);
      bool mine = (key == kmin) && ((pending >> lane) & 1);
      if (JAC) {
        if (kmin != acc_key) {
          if (CMP) {  // the compact record is staged in the warp's own tile (free between two row sets) and expanded
            flush_record<NT, 0, Xf, DENSE>(acc, acc_key, slot, warp_global, rb, lane, Jt, xf);
            Jt[31 * kJtStride + lane] = 0.0;
            __syncwarp();
          } else {
            flush_record<NT>(acc, acc_key, slot, warp_global, rb, lane);
          }
          acc.zero();
          acc_key = kmin;
        }
        typename std::conditional<CMP, ImuCompactSink, ImuRowSink<MARG>>::type sink;
        sink.Jt = Jt; sink.lane = lane; sink.on = mine; sink.acc = &acc;
        imu_sink_mode(sink, st.toff_free != 0);
        double rho = imu_eval<MARG>(W, mine ? s : kmin, mine ? u : 0.0, gy, ac, sh, true, sink);
        if (mine) { if (s + 3 >= pp.kf || st.extras_free || MARG) cost += 0.5 * rho; else fcost += 0.5 * rho; }
      } else {
        if (mine) {
          NullSink sink;
          const double rho = imu_eval<MARG>(W, s, u, gy, ac, sh, false, sink);
          if (s + 3 >= pp.kf || st.extras_free || MARG) cost += 0.5 * rho; else fcost += 0.5 * rho;
        }
      }
      pending &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(0xffffffffu, cost, o);
    fcost += __shfl_xor_sync(0xffffffffu, fcost, o);
  }
  if (DENSE) {
    dense_stamp(dn, 7);
    cta_finish_dense<NT, 0, kWarpsPerCta, Xf>(acc, acc_key, slot, warp, lane, warp_global, rb, rest, *dn, cost, fcost, JAC, xf);
    return;
  }
  if (JAC) {
    cta_merge_and_flush<NT, 0, kWarpsPerCta, Xf>(acc, acc_key, slot, warp, lane, warp_global, rb, rest, xf);
    for (int sidx = slot + lane; sidx < kSlotsPerWarp; sidx += 32) rb.key[warp_global * kSlotsPerWarp + sidx] = -1;
  }
  if (lane == 0) {
    rb.cost[2 * warp_global] = cost;
    rb.cost[2 * warp_global + 1] = fcost;
  }
}
template <bool JAC, bool MARG>
__global__ void __launch_bounds__(kThreads, 1)
imu_kernel(ImuStream st, PassParams pp, int total_slabs, RecordBuf rb) {
  if (pass_prologue(pp)) return;
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  imu_body<JAC, MARG>(st, pp, total_slabs, rb, W, rest, blockIdx.x, gridDim.x);
  pass_epilogue(pp);
}

// ---------------- K3: ImageFeatureFactor over a stream ----------------
struct ImageStream {
  int64_t n;
  const int64_t* ti_ns;  // kDropped if either time was outside the window at add time
  const int64_t* tj_ns;
  const double* pi[3];
  const double* pj[2];
  const int32_t* landmark;
  double loss_a;
  // Pose packs (n_packs > 0): every feature of a camera frame shares its timestamp, so the spline poses and their
  // Jacobian maps are evaluated once per distinct (timestamp, role) by each CTA and the factors only compose them.
  // Free inverse depths (fixed_depth = false): the terms that involve a landmark's column are gathered per LANDMARK by
  // rho_side_kernel (one warp per landmark over its observations, fixed order: deterministic; round 1 scattered them per
  // factor with atomics) into rho_side[rank][col] = H[rho, col] (col < Dtot, any non-rho column),
  // [Dtot] = H[rho, rho], [Dtot + 1] = b[rho]  (two factors of one interval key see different landmarks, so these
  // cannot ride in the per-key tiles).  The rho columns are rho0 .. rho0 + n_rho - 1 (last in a solve, among the
  // dropped ones in a marginalization).
  const int32_t* lm_col;     // per landmark: its column of H, or -1 (constant)
  const int* knot_col;       // per knot: column of its d(theta), -1 if constant
  double* rho_side;
  int Dtot, rho0, col_toff_cam;  // Dtot = number of columns of the system, rho0 = first inverse-depth column
  const int32_t* id_i;       // per factor: pack of t_i (0 .. n_pack_i-1)
  const int32_t* id_j;       // per factor: pack of t_j (n_pack_i .. n_packs-1)
  const int64_t* pack_t_ns;  // [n_packs] distinct times (ns, without offset)
  int n_pack_i, n_packs;
};
constexpr int kMaxPosePacks = 96;

// Local columns (NT = 7): [window of t_i 0..23 | window of t_j 24..47 | camera t_offset 48 | r 49 | pad]; key = s_i*(K-3)+s_j.
// A slab is 16 factors x 2 residual rows: lanes 0-15 evaluate row 0, lanes 16-31 row 1 of the same factors, so one
// 32-row tile accumulation covers a slab and each lane builds one Jacobian row only (this stream is small and
// latency-bound: shorter per-lane work + twice as many slabs to spread over the SMs).
constexpr int kImageFactorsPerSlab = 16;
template <bool JAC, bool PACKED, bool DENSE = false>
__device__ __forceinline__ void image_body(const ImageStream& st, const PassParams& pp, int total_slabs, const RecordBuf& rb,
                                           const ImageShared& sh, const WindowView& W, double* rest, int cta, int n_ctas,
                                           const DenseOut* dn = nullptr) {
  constexpr int NT = 7;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_global = cta * kWarpsPerCta + warp;
  const int row_sel = lane >> 4;
  const int64_t toff = (int64_t)pp.state[pp.sl.off_toff() + 2];  // camera offset, image_feature_factor.h:75-76
  const double* packs = rest;
  if (PACKED) {  // thread (pack, c): row c of the pack's Jacobian maps; c == 0 also the pose itself
    double* pk = rest;
    // every task is one chain evaluation (~10 us of dependent FP64 latency): as many tasks per pack as fit ONE round
    // of the CTA's threads — 3 (one row of the maps each), 2 (row 0 + pose | rows 1, 2) or 1
    const int tpp = 3 * st.n_packs <= (int)blockDim.x ? 3 : (2 * st.n_packs <= (int)blockDim.x ? 2 : 1);
    for (int t = threadIdx.x; t < tpp * st.n_packs; t += blockDim.x) {
      const int id = t / tpp, q = t % tpp;
      const int c0 = tpp == 3 ? q : (tpp == 2 ? (q == 0 ? 0 : 1) : 0);
      const int c1 = tpp == 3 ? q + 1 : (tpp == 2 ? (q == 0 ? 1 : 3) : 3);
      int s = 0;
      double u = 0.0;
      if (!index_ns(st.pack_t_ns[id] + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &s, &u)) { s = 0; u = 0.0; }  // unused pack
      compute_pose_pack_rows(W, s, u, id >= st.n_pack_i ? 1 : 0, sh.inv_dt, c0, c1, pk + (size_t)id * kPackDoubles);
    }
    rest += (size_t)st.n_packs * kPackDoubles;  // even: what follows stays 16-byte aligned
    __syncthreads();
  }
  double* Jt = rest + (size_t)warp * (8 * NT) * kJtStride;
  const double* invd = pp.state + pp.sl.off_invd();
  const int nk = pp.sl.K - 3;
  WarpAcc<NT> acc;
  acc.zero();
  int acc_key = -1, slot = 0;
  double cost = 0.0, fcost = 0.0;
  const int64_t n_warps_grid = (int64_t)n_ctas * kWarpsPerCta;
  int64_t slab0;
  int slabs_per_warp;
  warp_slab_range(DENSE ? dn : nullptr, total_slabs, warp_global, warp, kWarpsPerCta, n_warps_grid, &slab0, &slabs_per_warp);
  if (DENSE) dense_stamp(dn, 6);
  for (int sl = 0; sl < slabs_per_warp; sl++) {
    if ((slab0 + sl) * kImageFactorsPerSlab >= st.n) break;
    const int64_t idx = (slab0 + sl) * kImageFactorsPerSlab + (lane & 15);
    int64_t ti = kDropped, tj = kDropped;
    if (idx < st.n) { ti = st.ti_ns[idx]; tj = st.tj_ns[idx]; }
    int si = 0, sj = 0;
    double ui = 0.0, uj = 0.0;
    bool valid = (ti != kDropped) && (tj != kDropped) && index_ns(ti + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &si, &ui) &&
                 index_ns(tj + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &sj, &uj);
    V3 p_i = mk(0, 0, 1), p_j = mk(0, 0, 1);
    double rho_inv = 1.0;
    int pack_i = 0, pack_j = 0, rho_col = -1;
    if (valid) {
      p_i = mk(st.pi[0][idx], st.pi[1][idx], st.pi[2][idx]);
      p_j = mk(st.pj[0][idx], st.pj[1][idx], 1.0);
      rho_inv = invd[st.landmark[idx]];
      if (st.lm_col) rho_col = st.lm_col[st.landmark[idx]];
      if (PACKED) { pack_i = st.id_i[idx]; pack_j = st.id_j[idx]; }
    } else {
      si = sj = 0; ui = uj = 0.0;
    }
    int key = valid ? si * nk + sj : -1;
    unsigned pending = __ballot_sync(0xffffffffu, key >= 0);
    while (pending) {
      int kmin = __reduce_min_sync(0xffffffffu, (key >= 0 && ((pending >> lane) & 1)) ? key : 0x7fffffff);
      bool mine = (key == kmin) && ((pending >> lane) & 1);
      // a visual factor touches knots of both windows: it has a free parameter unless both windows are all constant
      const bool any_free = (max(si, sj) + 3 >= pp.kf) || sh.want_toff || rho_col >= 0 || pp.marg_mode;
      const bool counts = mine && row_sel == 0;  // the factor's cost is counted once
      if (JAC) {
        if (kmin != acc_key) {
          flush_record<NT>(acc, acc_key, slot, warp_global, rb, lane);
          acc.zero();
          acc_key = kmin;
        }
        SmemRowSink<NT> sink;
        sink.Jt = Jt; sink.lane = lane; sink.on = mine; sink.acc = &acc;
        double rho = PACKED ? image_eval_packed(packs + (size_t)pack_i * kPackDoubles, packs + (size_t)pack_j * kPackDoubles,
                                                p_i, p_j, rho_inv, sh, true, sink, row_sel, rho_col >= 0)
                            : image_eval(W, si, ui, sj, uj, p_i, p_j, rho_inv, sh, true, sink, row_sel, rho_col >= 0);
        if (counts) { if (any_free) cost += 0.5 * rho; else fcost += 0.5 * rho; }
        __syncwarp();
      } else {
        NullSink sink;
        double rho = PACKED ? image_eval_packed(packs + (size_t)pack_i * kPackDoubles, packs + (size_t)pack_j * kPackDoubles,
                                                p_i, p_j, rho_inv, sh, false, sink, row_sel)
                            : image_eval(W, si, ui, sj, uj, p_i, p_j, rho_inv, sh, false, sink, row_sel);
        if (counts) { if (any_free) cost += 0.5 * rho; else fcost += 0.5 * rho; }
      }
      pending &= ~__ballot_sync(0xffffffffu, mine);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_xor_sync(0xffffffffu, cost, o);
    fcost += __shfl_xor_sync(0xffffffffu, fcost, o);
  }
  if (DENSE) {
    dense_stamp(dn, 7);
    cta_finish_dense<NT, 0, kWarpsPerCta>(acc, acc_key, slot, warp, lane, warp_global, rb, rest, *dn, cost, fcost, JAC);
    return;
  }
  if (JAC) {
    cta_merge_and_flush<NT>(acc, acc_key, slot, warp, lane, warp_global, rb, rest);
    for (int sidx = slot + lane; sidx < kSlotsPerWarp; sidx += 32) rb.key[warp_global * kSlotsPerWarp + sidx] = -1;
  }
  if (lane == 0) {
    rb.cost[2 * warp_global] = cost;
    rb.cost[2 * warp_global + 1] = fcost;
  }
}
template <bool JAC, bool PACKED>
__global__ void __launch_bounds__(kThreads, 1)
image_kernel(ImageStream st, PassParams pp, int total_slabs, RecordBuf rb, ImageShared sh) {
  if (pass_prologue(pp)) return;
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  image_body<JAC, PACKED>(st, pp, total_slabs, rb, sh, W, rest, blockIdx.x, gridDim.x);
  pass_epilogue(pp);
}

// The inverse-depth columns of the system (free or marginalised landmarks): side[rank][g] = H[rho, g] for every other
// column g, [Dtot] = H[rho, rho], [Dtot + 1] = b[rho].  One warp per landmark: its observations (lm_obs[lm_start[l] ..
// lm_start[l+1]) = stored positions, ascending) are evaluated 16 at a time — lanes 2o, 2o + 1 build the two residual rows
// of observation o in shared memory — and then added observation by observation, lanes over the local columns (window i,
// then window j, then the time offset: two windows may share a knot), into the warp's shared side row, which goes out
// with plain stores.  Same order every run; no atomics.
constexpr int kRhoWarps = 4;
__host__ __device__ inline size_t rho_side_smem_doubles(int K, int Dtot) {
  return window_smem_doubles(K) + (size_t)kRhoWarps * (32 * 56 + ((Dtot + 2 + 1) & ~1));
}
struct SmemPlainRow {
  double* row;
  __device__ __forceinline__ void put(int, int col, double v) { row[col] = v; }
  __device__ __forceinline__ void row_done(int) {}
};
__global__ void __launch_bounds__(kRhoWarps * 32) rho_side_kernel(ImageStream st, PassParams pp, ImageShared sh,
                                                                   const int* lm_start, const int* lm_obs, int n_lm) {
  pdl_wait();
  pdl_trigger();
  if (pp.ctl && pp.ctl[0]) return;
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int side_n = st.Dtot + 2, side_pad = (side_n + 1) & ~1;
  double* rows = rest + (size_t)warp * (32 * 56 + side_pad);
  double* side_s = rows + 32 * 56;
  const int64_t toff = (int64_t)pp.state[pp.sl.off_toff() + 2];
  const double* invd = pp.state + pp.sl.off_invd();
  for (int l = blockIdx.x * kRhoWarps + warp; l < n_lm; l += gridDim.x * kRhoWarps) {
    const int col = st.lm_col[l];
    if (col < 0) continue;
    for (int c = lane; c < side_n; c += 32) side_s[c] = 0.0;
    const int o0 = lm_start[l], o1 = lm_start[l + 1];
    for (int base = o0; base < o1; base += 16) {
      __syncwarp();
      const int o = base + (lane >> 1);
      int si = 0, sj = 0;
      bool valid = false;
      if (o < o1) {
        const int64_t idx = lm_obs[o];
        const int64_t ti = st.ti_ns[idx], tj = st.tj_ns[idx];
        double ui = 0.0, uj = 0.0;
        valid = ti != kDropped && tj != kDropped && index_ns(ti + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &si, &ui) &&
                index_ns(tj + toff, pp.t0_ns, pp.dt_ns, pp.sl.K, &sj, &uj);
        if (valid) {
          SmemPlainRow sink;
          sink.row = rows + lane * 56;
          image_eval(W, si, ui, sj, uj, mk(st.pi[0][idx], st.pi[1][idx], st.pi[2][idx]), mk(st.pj[0][idx], st.pj[1][idx], 1.0),
                     invd[l], sh, true, sink, lane & 1, true);
        }
      }
      __syncwarp();
      for (int q = 0; q < 16 && base + q < o1; q++) {
        const int v_q = __shfl_sync(0xffffffffu, valid ? 1 : 0, 2 * q);
        const int si_q = __shfl_sync(0xffffffffu, si, 2 * q), sj_q = __shfl_sync(0xffffffffu, sj, 2 * q);
        if (!v_q) continue;
        for (int rr = 0; rr < 2; rr++) {
          const double* row = rows + (2 * q + rr) * 56;
          const double jr = row[50];
          if (lane < 24) {  // window of t_i
            const int kc = st.knot_col[si_q + lane / 6];
            if (kc >= 0) side_s[kc + lane % 6] += jr * row[lane];
          }
          __syncwarp();
          if (lane < 24) {  // window of t_j
            const int kc = st.knot_col[sj_q + lane / 6];
            if (kc >= 0) side_s[kc + lane % 6] += jr * row[24 + lane];
          }
          __syncwarp();
          if (lane == 0) {
            if (st.col_toff_cam >= 0) side_s[st.col_toff_cam] += jr * row[48];
            side_s[st.Dtot] += jr * jr;
            side_s[st.Dtot + 1] += jr * row[49];
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
    double* side = st.rho_side + (size_t)(col - st.rho0) * side_n;
    for (int c = lane; c < side_n; c += 32) side[c] += side_s[c];  // (zeroed per pass; batches run one after the other)
    __syncwarp();
  }
}

// ---------------- one-shot all-reduce over NVLink peer memory ----------------
// The exchange step of the path is tiny (H | b | cost: D^2 + D + 2 doubles, 42 KB at C5) and latency-bound, so instead
// of a ring / tree collective every rank PUSHES its vector into a per-source slot of every peer's exchange area
// (posted NVLink stores), publishes the sequence number with a system-scope release store, waits for the n_ranks
// flags of its own area with acquire loads, and sums the n_ranks slots in rank order from LOCAL memory: one launch,
// one NVLink hop, and the same summation order on every rank (bit-identical results).  Slots are double-buffered by
// the parity of the sequence number: a rank can only be one all-reduce ahead of its slowest peer, so a slot is never
// overwritten while it is still being read.  The wait is bounded (~1 s): a missing peer sets *err instead of hanging.
constexpr int kP2PMaxRanks = 16;
constexpr size_t kP2PMaxDoubles = 128 * 1024;  // per slot: D <= ~360
// flags: [kP2PMaxRanks] per-rank words of p2p_allreduce_kernel, then (offset 256) [kP2PMaxRanks][512] per-(rank, CTA) words
// of the fused pass_kernel exchange (pass_fused.cuh)
constexpr size_t kP2PFlagBytes = 128 * 1024;
struct P2PView {
  char* peer[kP2PMaxRanks];  // exchange areas: [flags: kP2PMaxRanks x u64][slots: 2 x kP2PMaxRanks x kP2PMaxDoubles doubles]
  int n, rank;
};
__host__ __device__ inline size_t p2p_area_bytes() {
  return kP2PFlagBytes + (size_t)2 * kP2PMaxRanks * kP2PMaxDoubles * sizeof(double);
}
__device__ __forceinline__ double* p2p_slot(char* area, int parity, int src_rank) {
  return reinterpret_cast<double*>(area + kP2PFlagBytes) + ((size_t)parity * kP2PMaxRanks + src_rank) * kP2PMaxDoubles;
}
// The wait is bounded in wall-clock time (timeout_ns of %globaltimer, tens of seconds by default: a peer that is merely
// slow — first-call module load, a large copy ahead of it — is not a missing peer).  On a timeout *err is set and dst
// is ZEROED, never a partial sum.
__global__ void __launch_bounds__(1024) p2p_allreduce_kernel(const double* src, double* dst, int count, P2PView v,
                                                             unsigned long long seq, unsigned long long timeout_ns,
                                                             int* err) {
  pdl_wait();
  pdl_trigger();
  const int tid = threadIdx.x, parity = (int)(seq & 1ull);
  for (int i = tid; i < count; i += blockDim.x) {
    const double x = src[i];
    for (int r = 0; r < v.n; r++) p2p_slot(v.peer[r], parity, v.rank)[i] = x;
  }
  __syncthreads();
  bool timed_out = false;
  if (tid < v.n) {
    __threadfence_system();  // cumulative over the CTA's stores ordered before the barrier
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(v.peer[tid]) + v.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(seq) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(v.peer[v.rank]) + tid;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned long long got = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(mine) : "memory");
      if (got >= seq) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) { timed_out = true; break; }
      __nanosleep(64);
    }
  }
  const int any_timeout = __syncthreads_or(timed_out ? 1 : 0);
  if (any_timeout && tid == 0) *err = 1;
  char* local = v.peer[v.rank];
  for (int i = tid; i < count; i += blockDim.x) {
    double acc = 0.0;
    if (!any_timeout)
      for (int r = 0; r < v.n; r++) acc += __ldcv(p2p_slot(local, parity, r) + i);
    dst[i] = acc;
  }
}

// ---------------- forward-only spline evaluation (the step before the path) ----------------
// Trajectory::GetSensorPose (trajectory.cpp:100-115) for a batch of timestamps: one query per thread, grid-stride;
// the knot-only part of So3Spline::evaluate (delta_i = log(R_i^-1 R_i+1), so3_spline.h:240-244) comes from the
// per-CTA PairData like everywhere else.  HBM traffic: 8 B in, 56 B out per pose (32 B in+out per undistorted point).
struct PoseQuery {
  int64_t n;
  const double* t_s;
  Q4 S_StoI;
  V3 p_SinI;
  double t_offset_ns, t_min_ns, t_max_ns;  // minTimeNs / maxTimeNs as doubles (the reference compares in double)
};
__device__ __forceinline__ bool sensor_pose(const WindowView& W, const PassParams& pp, const PoseQuery& q, double t_s,
                                            Q4* R_out, V3* p_out) {
  const double tn = __dadd_rn(__dmul_rn(t_s, 1e9), q.t_offset_ns);  // product rounded first (trajectory.cpp:101), no FMA contraction
  if (!(tn >= q.t_min_ns && tn < q.t_max_ns)) return false;
  int s = 0;
  double u = 0.0;
  if (!index_ns((int64_t)tn, pp.t0_ns, pp.dt_ns, pp.sl.K, &s, &u)) return false;
  So3Chain C;
  Q4 P[4];
  so3_chain_rtp(W, s, u, &C, P);  // forward accumulation R_s * prod exp(lambda_i+1 delta_i)
  double c[4];
  blend_plain(u, c);
  V3 pos = mk(0, 0, 0);
#pragma unroll
  for (int k = 0; k < 4; k++) pos = fma3(c[k], ldv(W.kp + 3 * (s + k)), pos);
  *R_out = qmul(C.R, q.S_StoI);              // pose_I_to_G * EP_StoI.se3
  *p_out = qrot(C.R, q.p_SinI) + pos;
  return true;
}
__global__ void __launch_bounds__(256) pose_query_kernel(PassParams pp, PoseQuery q, double* q_out, double* p_out,
                                                         unsigned long long* n_invalid) {
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  unsigned long long bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * blockDim.x) {
    Q4 R; R.x = R.y = R.z = 0.0; R.w = 1.0;
    V3 p = mk(0, 0, 0);
    if (!sensor_pose(W, pp, q, q.t_s[i], &R, &p)) { bad++; R.x = R.y = R.z = 0.0; R.w = 1.0; p = mk(0, 0, 0); }
    q_out[4 * i] = R.x; q_out[4 * i + 1] = R.y; q_out[4 * i + 2] = R.z; q_out[4 * i + 3] = R.w;
    p_out[3 * i] = p.x; p_out[3 * i + 1] = p.y; p_out[3 * i + 2] = p.z;
  }
  if (bad && n_invalid) atomicAdd(n_invalid, bad);
}
// target: 7 doubles (q xyzw, p) of GetLidarPose(target) written by pose_query_kernel, or nullptr for the G frame.
__global__ void __launch_bounds__(256) undistort_kernel(PassParams pp, PoseQuery q, const float* xyz_in,
                                                        const double* target, float* xyz_out,
                                                        unsigned long long* n_invalid) {
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  Q4 Rt; Rt.x = Rt.y = Rt.z = 0.0; Rt.w = 1.0;
  V3 pt = mk(0, 0, 0);
  if (target) {  // pose_G_to_target = pose.inverse()
    Rt = qconj(ldq(target));
    pt = -1.0 * qrot(Rt, ldv(target + 4));
  }
  unsigned long long bad = 0;
  const float nan_f = __int_as_float(0x7fc00000);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = xyz_in[3 * i], y = xyz_in[3 * i + 1], z = xyz_in[3 * i + 2];
    Q4 R;
    V3 p;
    if (isnan(x) || !sensor_pose(W, pp, q, q.t_s[i], &R, &p)) {
      bad++;
      xyz_out[3 * i] = xyz_out[3 * i + 1] = xyz_out[3 * i + 2] = nan_f;
      continue;
    }
    const V3 pl = mk((double)x, (double)y, (double)z);
    V3 o;
    if (!target) {
      o = qrot(R, pl) + p;
    } else {  // (pose_G_to_target * pose_Lk_to_G) * p_Lk: poses composed first, as the reference does
      const Q4 Rc = qmul(Rt, R);
      const V3 pc = qrot(Rt, p) + pt;
      o = qrot(Rc, pl) + pc;
    }
    xyz_out[3 * i] = (float)o.x; xyz_out[3 * i + 1] = (float)o.y; xyz_out[3 * i + 2] = (float)o.z;
  }
  if (bad && n_invalid) atomicAdd(n_invalid, bad);
}

__global__ void prep_image_kernel(int64_t n, const double* ti_s, const double* tj_s, const double* p_i, const double* p_j,
                                  int64_t t_lo, int64_t t_hi, int64_t pad, int64_t* ti_ns, int64_t* tj_ns, double* pi0,
                                  double* pi1, double* pi2, double* pj0, double* pj1, unsigned long long* n_dropped) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t a = (int64_t)(ti_s[i] * 1e9), b = (int64_t)(tj_s[i] * 1e9);
  bool ok = !(a - pad < t_lo || a + pad >= t_hi) && !(b - pad < t_lo || b + pad >= t_hi);
  ti_ns[i] = ok ? a : kDropped;
  tj_ns[i] = ok ? b : kDropped;
  if (!ok) atomicAdd(n_dropped, 1ULL);
  pi0[i] = p_i[3 * i]; pi1[i] = p_i[3 * i + 1]; pi2[i] = p_i[3 * i + 2];
  pj0[i] = p_j[3 * i]; pj1[i] = p_j[3 * i + 1];
}

// ---------------- assembly of the dense normal equations ----------------
// (StreamDesc / ColMaps: defined near the top of this file)
__device__ __forceinline__ double block_entry(const double* blk, int NT, int la, int lb) {
  int ta = la >> 3, tb = lb >> 3;  // la, lb: PHYSICAL columns
  if (ta > tb) { int t = ta; ta = tb; tb = t; t = la; la = lb; lb = t; }
  return blk[tile_index(NT, ta, tb) * 64 + (la & 7) * 8 + (lb & 7)];
}

// BiasFactor (trajectory_value_factor.h:65-126) added to H, b, cost by one thread (6 residuals, +-I Jacobians).
struct BiasFactorDesc {
  int slot_i, slot_j;
  double sqrt_info[6];   // info / sqrt(dt)
  int col_gi, col_gj, col_ai, col_aj;  // columns (-1 constant)
};
// H[gi][gj] (upper, mirrored) and b[gi]: owner-computes gather over the per-key blocks.  One WARP per entry: the
// lanes stride the block list in a fixed assignment and are combined by a fixed butterfly,
// so the sum order never changes between runs.
// bias (n_bias > 0, solve path): the BiasFactor terms (+-I Jacobians) are added by the owner of each entry, and the
// first warp adds their cost — one launch less per pass than bias_factor_kernel.
// Entry `wid` of the packed system [upper triangle of H, row-major | b]: its value on lane 0 of the calling warp (all 32
// lanes must call); *gi_out, *gj_out = its row / column (gj = -1 for an entry of b); *bias_cost_out = cost of the
// BiasFactors (same on every entry).
__device__ __forceinline__ double assemble_entry(int wid, int lane, const ColMaps& cm, const StreamDesc* descs, int n_desc,
                                                 const BiasFactorDesc* bias, int n_bias, const double* state, int off_bias,
                                                 const double* rho_side, int rho0, int n_rho, int* gi_out, int* gj_out,
                                                 double* bias_cost_out, bool* bias_free_out) {
  const int D = cm.D;
  const int n_upper = D * (D + 1) / 2;
  int gi, gj;
  const bool is_b = wid >= n_upper;
  if (is_b) {
    gi = wid - n_upper;
    gj = -1;
  } else {  // unrank (gi <= gj) from the row-major upper-triangular index: rows before gi hold gi*D - gi(gi-1)/2 entries
    const double q = 2.0 * D + 1.0;
    gi = (int)((q - sqrt(q * q - 8.0 * (double)wid)) * 0.5);
    gi = max(0, min(D - 1, gi));
    while (gi > 0 && gi * D - gi * (gi - 1) / 2 > wid) gi--;
    while ((gi + 1) * D - (gi + 1) * gi / 2 <= wid) gi++;
    gj = gi + (wid - (gi * D - gi * (gi - 1) / 2));
  }
  *gi_out = gi;
  *gj_out = gj;
  const int ki = cm.col_knot[gi], ci = cm.col_sub[gi];
  const int kj = is_b ? -1 : cm.col_knot[gj], cj = is_b ? 0 : cm.col_sub[gj];
  double acc = 0.0;
  for (int d = 0; d < n_desc; d++) {
    const StreamDesc& sd = descs[d];
    int ei = -1, ej = -1;  // extra local columns
    for (int e = 0; e < sd.n_extra; e++) {
      if (sd.extra_global[e] == gi) ei = sd.extra_base + e;
      if (!is_b && sd.extra_global[e] == gj) ej = sd.extra_base + e;
    }
    if (ki < 0 && ei < 0) continue;
    if (!is_b && kj < 0 && ej < 0) continue;
    if (sd.kind == 0) {
      int s_lo = 0, s_hi = sd.n_keys - 1;
      if (ki >= 0) { s_lo = max(s_lo, ki - 3); s_hi = min(s_hi, ki); }
      if (!is_b && kj >= 0) { s_lo = max(s_lo, kj - 3); s_hi = min(s_hi, kj); }
      const int n_items = max(0, s_hi - s_lo + 1);
      for (int it = lane; it < n_items; it += 32) {
        const int s = s_lo + it;
        const int la = ki >= 0 ? 6 * (ki - s) + ci : ei;
        const int lb = is_b ? sd.r_col : (kj >= 0 ? 6 * (kj - s) + cj : ej);
        const double* blk = sd.fin + (size_t)s * sd.rd;
        acc += (is_b && sd.b_off >= 0) ? blk[sd.b_off + la] : block_entry(blk, sd.NT, sd.perm[la], sd.perm[lb]);
      }
    } else {
      // two knot windows: a global knot column may appear in window i, window j or both (overlap): J_g = sum of locals
      const int nk = sd.kdim;
      const int n_items = nk * nk;
      for (int it = lane; it < n_items; it += 32) {
        const int key = it;
        const int si = key / nk, sj = key % nk;
        int la[2], lb[2], na = 0, nb = 0;
        if (ki >= 0) {
          if (ki - si >= 0 && ki - si <= 3) la[na++] = 6 * (ki - si) + ci;
          if (ki - sj >= 0 && ki - sj <= 3) la[na++] = 24 + 6 * (ki - sj) + ci;
        } else {
          la[na++] = ei;
        }
        if (is_b) {
          lb[nb++] = sd.r_col;
        } else if (kj >= 0) {
          if (kj - si >= 0 && kj - si <= 3) lb[nb++] = 6 * (kj - si) + cj;
          if (kj - sj >= 0 && kj - sj <= 3) lb[nb++] = 24 + 6 * (kj - sj) + cj;
        } else {
          lb[nb++] = ej;
        }
        if (na == 0 || nb == 0) continue;
        const double* blk = sd.fin + (size_t)key * sd.rd;
        for (int x = 0; x < na; x++)
          for (int y = 0; y < nb; y++) acc += block_entry(blk, sd.NT, sd.perm[la[x]], sd.perm[lb[y]]);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (rho_side) {  // entries of the inverse-depth columns (visual factors with a free / marginalised landmark)
      const bool ri = gi >= rho0 && gi < rho0 + n_rho, rj = !is_b && gj >= rho0 && gj < rho0 + n_rho;
      const size_t ld = (size_t)D + 2;
      if (is_b) {
        if (ri) acc += rho_side[(size_t)(gi - rho0) * ld + D + 1];
      } else if (ri && rj) {
        if (gi == gj) acc += rho_side[(size_t)(gi - rho0) * ld + D];
      } else if (rj) {
        acc += rho_side[(size_t)(gj - rho0) * ld + gi];
      } else if (ri) {
        acc += rho_side[(size_t)(gi - rho0) * ld + gj];
      }
    }
    double bias_cost = 0.0;
    bool bias_free = false;
    for (int f = 0; f < n_bias; f++) {
      const BiasFactorDesc& bf = bias[f];
      const double* bi = state + off_bias + 6 * bf.slot_i;
      const double* bj = state + off_bias + 6 * bf.slot_j;
      bias_free = bias_free || bf.col_gi >= 0 || bf.col_gj >= 0 || bf.col_ai >= 0 || bf.col_aj >= 0;
      double tot = 0.0;
      for (int r = 0; r < 6; r++) {
        const double w = bf.sqrt_info[r], res = w * (bj[r] - bi[r]);
        tot += res * res;
        const int base_i = r < 3 ? bf.col_gi : bf.col_ai, base_j = r < 3 ? bf.col_gj : bf.col_aj;
        const int ci = base_i >= 0 ? base_i + r % 3 : -1, cj = base_j >= 0 ? base_j + r % 3 : -1;
        if (is_b) {
          if (ci >= 0 && gi == ci) acc += -w * res;
          if (cj >= 0 && gi == cj) acc += w * res;
        } else {
          if (ci >= 0 && gi == ci && gj == ci) acc += w * w;
          if (cj >= 0 && gi == cj && gj == cj) acc += w * w;
          if (ci >= 0 && cj >= 0 && gi == min(ci, cj) && gj == max(ci, cj)) acc += -w * w;
        }
      }
      bias_cost += 0.5 * tot;
    }
    *bias_cost_out = bias_cost;
    *bias_free_out = bias_free;
  }
  return acc;
}
__global__ void __launch_bounds__(256) assemble_kernel(double* H, double* b, ColMaps cm, const StreamDesc* descs,
                                                       int n_desc, const int* ctl, const BiasFactorDesc* bias = nullptr,
                                                       int n_bias = 0, const double* state = nullptr,
                                                       int off_bias = 0, double* cost2 = nullptr,
                                                       const double* rho_side = nullptr, int rho0 = 0, int n_rho = 0) {
  pdl_wait();
  pdl_trigger();
  if (ctl && ctl[0]) return;
  const int D = cm.D;
  const int n_upper = D * (D + 1) / 2;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wid >= n_upper + D) return;
  int gi, gj;
  double bias_cost = 0.0;
  bool bias_free = false;
  const double acc = assemble_entry(wid, lane, cm, descs, n_desc, bias, n_bias, state, off_bias, rho_side, rho0, n_rho,
                                    &gi, &gj, &bias_cost, &bias_free);
  if (lane == 0) {
    if (wid == 0 && n_bias > 0) cost2[bias_free ? 0 : 1] += bias_cost;
    if (gj < 0) {
      b[gi] = acc;
    } else {
      H[(size_t)gi * D + gj] = acc;
      H[(size_t)gj * D + gi] = acc;
    }
  }
}

__global__ void bias_factor_kernel(BiasFactorDesc bf, const double* state, StateLayout sl, double* H, double* b, int D,
                                   double* cost2, int jac, const int* ctl) {
  pdl_wait();
  pdl_trigger();
  if (ctl && ctl[0]) return;
  // one lane per residual row: rows touch disjoint H / b entries (row r -> offset r % 3 of the gyro or accel blocks)
  const int r = threadIdx.x;
  double c = 0.0;
  if (r < 6) {
    const double* bi = state + sl.off_bias() + 6 * bf.slot_i;
    const double* bj = state + sl.off_bias() + 6 * bf.slot_j;
    const double w = bf.sqrt_info[r];
    const double res = w * (bj[r] - bi[r]);
    c = res * res;
    if (jac) {
      const int ci = (r < 3 ? bf.col_gi : bf.col_ai), cj = (r < 3 ? bf.col_gj : bf.col_aj);
      const int o = r % 3;
      if (ci >= 0) {
        H[(size_t)(ci + o) * D + ci + o] += w * w;
        b[ci + o] += -w * res;
      }
      if (cj >= 0) {
        H[(size_t)(cj + o) * D + cj + o] += w * w;
        b[cj + o] += w * res;
      }
      if (ci >= 0 && cj >= 0) {
        H[(size_t)(ci + o) * D + cj + o] += -w * w;
        H[(size_t)(cj + o) * D + ci + o] += -w * w;
      }
    }
  }
  // fixed-order sum of the 6 terms
  double tot = 0.0;
  for (int k = 0; k < 6; k++) tot += __shfl_sync(0xffffffffu, c, k);
  if (r == 0) {
    const bool any_free = bf.col_gi >= 0 || bf.col_gj >= 0 || bf.col_ai >= 0 || bf.col_aj >= 0;
    cost2[any_free ? 0 : 1] += 0.5 * tot;
  }
}

// GravityFactor (trajectory_value_factor.h:33-62) added to the marginalization system: r = w (g - g0), J = diag(w) on the
// three ambient gravity columns starting at `col` (-1: no column).  One lane per row.
__global__ void gravity_factor_kernel(double g0x, double g0y, double g0z, double wx, double wy, double wz, const double* g,
                                      int col, double* A, double* b, int D, double* cost2) {
  const int r = threadIdx.x;
  double c = 0.0;
  if (r < 3) {
    const double w = r == 0 ? wx : (r == 1 ? wy : wz), g0 = r == 0 ? g0x : (r == 1 ? g0y : g0z);
    const double res = w * (g[r] - g0);
    c = res * res;
    if (col >= 0) {
      A[(size_t)(col + r) * D + col + r] += w * w;
      b[col + r] += w * res;
    }
  }
  double tot = 0.0;
  for (int k = 0; k < 3; k++) tot += __shfl_sync(0xffffffffu, c, k);
  if (r == 0) cost2[0] += 0.5 * tot;
}

// auto_diff::IMUGlobalVelocityFactor (factor/auto_diff/trajectory_value_factor.h:436-487) with CauchyLoss(60)
// (trajectory_estimator.cpp:642): r = w (p'(t) - v), t = double(time_ns) + t_offset_IMU (not truncated: a Jet in the
// reference), indexed by the templated SplineSegmentMeta::computeTIndexNs (spline_segment.h:108-118: floor of a double
// quotient).  One factor per solve (InitTrajWithPropagation): a single small CTA adds it to H, b, cost after the
// assembly, every thread owning distinct entries.
struct VelFactorDesc {
  long long time_ns;
  double v[3];
  double w;
};
__global__ void velocity_factor_kernel(VelFactorDesc vf, const double* state, StateLayout sl, long long t0_ns,
                                       long long dt_ns, int K, const int* knot_col, int col_toff, double* H, double* b,
                                       int D, double* cost2, int jac, const int* ctl) {
  pdl_wait();
  pdl_trigger();
  if (ctl && ctl[0]) return;
  const double t_corr = (double)vf.time_ns + state[sl.off_toff() + 0];  // IMUSensor = 0
  const double t_lo = (double)t0_ns, t_hi = (double)(t0_ns + (long long)(K - 3) * dt_ns);
  if (t_corr < t_lo || t_corr >= t_hi) return;
  const double st = (t_corr - (double)t0_ns) / (double)dt_ns;
  const int s = (int)floor(st);
  const double u = st - (double)s;
  const double inv_dt = 1e9 * 1.0 / (double)dt_ns;
  double c[4], ca[4];
  blend_plain_d1(u, inv_dt, c);       // inv_dt * M p'(u)
  blend_plain_d2(u, inv_dt, ca);      // d/du of the above
  const double* kp = state + sl.off_p();
  double v[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
  for (int k = 0; k < 4; k++)
    for (int r = 0; r < 3; r++) {
      v[r] += c[k] * kp[3 * (s + k) + r];
      dv[r] += ca[k] * kp[3 * (s + k) + r];
    }
  double res[3], jt[3], sq = 0.0;
  for (int r = 0; r < 3; r++) {
    res[r] = vf.w * (v[r] - vf.v[r]);
    jt[r] = vf.w * dv[r] * (1.0 / (double)dt_ns);  // d r / d t_offset: du / dt_offset = 1 / dt_ns
    sq += res[r] * res[r];
  }
  double rho;
  const double sc = cauchy_scale(sq, 60.0, &rho);
  const double rp = sc * sc;  // rho'
  int col[4];
  bool any_free = col_toff >= 0;
  for (int k = 0; k < 4; k++) {
    col[k] = knot_col[s + k] >= 0 ? knot_col[s + k] + 3 : -1;
    any_free |= col[k] >= 0;
  }
  const int t = threadIdx.x;
  if (t == 63) cost2[any_free ? 0 : 1] += 0.5 * rho;
  if (!jac) return;
  if (t < 48) {  // H[p_k + r][p_l + r] += rho' w^2 c_k c_l
    const int k = t / 12, l = (t / 3) % 4, r = t % 3;
    if (col[k] >= 0 && col[l] >= 0) H[(size_t)(col[k] + r) * D + col[l] + r] += rp * (vf.w * c[k]) * (vf.w * c[l]);
  } else if (t < 60) {  // cross terms with the time offset, and b on the position knots
    const int k = (t - 48) / 3, r = (t - 48) % 3;
    if (col[k] >= 0) {
      b[col[k] + r] += rp * (vf.w * c[k]) * res[r];
      if (col_toff >= 0) {
        const double h = rp * (vf.w * c[k]) * jt[r];
        H[(size_t)(col[k] + r) * D + col_toff] += h;
        H[(size_t)col_toff * D + col[k] + r] += h;
      }
    }
  } else if (t == 60 && col_toff >= 0) {
    H[(size_t)col_toff * D + col_toff] += rp * (jt[0] * jt[0] + jt[1] * jt[1] + jt[2] * jt[2]);
    b[col_toff] += rp * (jt[0] * res[0] + jt[1] * res[1] + jt[2] * res[2]);
  }
}

// ---- the small analytic factors ----
// Values (x) and first columns (xcol, -1 = constant) of a small factor's parameter blocks after the knots.
struct SmallX {
  double x[16];
  int xcol[4];
};
// One small factor into o (thread 0 of the CTA).  False: a measurement time left the window.
__device__ __noinline__ bool small_dispatch(const WindowView& W, const SmallFactor& f, const PreintData* pre, const PassParams& pp,
                                            Q4 S_CtoI, V3 p_CinI, const SmallX& X, SmallCols sc, SmallOut* o) {
  const long long t0 = pp.t0_ns, dt = pp.dt_ns;
  const int K = pp.sl.K;
  switch (f.kind) {
    case kSmallPose: sc.col_toff_imu = X.xcol[0]; return small_pose_eval(W, f, t0, dt, K, X.x[0], sc, o);
    case kSmallLocalVelocity: sc.col_toff_imu = X.xcol[0]; return small_local_velocity_eval(W, f, t0, dt, K, X.x[0], sc, o);
    case kSmallRelativeRotation: return small_relative_rotation_eval(W, f, t0, dt, K, sc, o);
    case kSmallImageDepth: return small_image_depth_eval(f, X.x[0], X.xcol[0], o);
    case kSmallPreIntegration: return small_preintegration_eval(W, f, pre[f.ext], t0, dt, K, X.x, X.xcol, sc, o);
    case kSmallImage3D2D: return small_image_3d2d_eval(W, f, t0, dt, K, S_CtoI, p_CinI, X.x, X.xcol, sc, o);
    case kSmallImageOnePose: return small_image_one_pose_eval(W, f, t0, dt, K, S_CtoI, p_CinI, X.x, X.xcol, sc, o);
    case kSmallEpipolar: return small_epipolar_eval(W, f, t0, dt, K, S_CtoI, p_CinI, sc, o);
    case kSmallLoamOptMapPose: return small_loam_opt_map_pose_eval(W, f, t0, dt, K, X.x, X.xcol, sc, o);
    case kSmallRelativeLoam: return small_relative_loam_eval(W, f, t0, dt, K, sc, o);
    default: return false;
  }
}

// The remaining adders of TrajectoryEstimator: pose, local-velocity, relative-rotation, image-depth and pre-integration
// factors — a handful per solve at most.  One CTA walks them in add order: thread 0 evaluates a factor into a small
// dense row set over the distinct global columns it touches (factor_eval.cuh), then every thread owns some of the
// (column, column) pairs and adds J~^T J~, J~^T r~ into H, b: no atomics, fixed order.  Runs after the assembly (or
// into the addend system of the fused pass), like velocity_factor_kernel.
__global__ void __launch_bounds__(128) small_factors_kernel(const SmallFactor* fs, int n, const PreintData* pre, PassParams pp,
                                                            SmallCols sc, double* H, double* b, int D, double* cost2,
                                                            int jac) {
  pdl_wait();
  pdl_trigger();
  if (pp.ctl && pp.ctl[0]) return;
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  SmallOut* o = reinterpret_cast<SmallOut*>(rest);
  __shared__ int s_ok;
  __shared__ double s_cost[2];
  if (threadIdx.x == 0) s_cost[0] = s_cost[1] = 0.0;
  const double toff_imu = pp.state[pp.sl.off_toff() + 0];
  const double* invd = pp.state + pp.sl.off_invd();
  Q4 qid; qid.x = qid.y = qid.z = 0.0; qid.w = 1.0;
  for (int i = 0; i < n; i++) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const SmallFactor f = fs[i];
      SmallX X;
      for (int k = 0; k < 4; k++) X.xcol[k] = -1;
      if (f.kind == kSmallPose || f.kind == kSmallLocalVelocity) {
        X.x[0] = toff_imu;
        X.xcol[0] = sc.col_toff_imu;
      } else if (f.kind == kSmallImageDepth) {
        X.x[0] = invd[f.landmark];
        X.xcol[0] = sc.lm_col ? sc.lm_col[f.landmark] : -1;
      } else if (f.kind == kSmallPreIntegration) {
        const double* bi = pp.state + pp.sl.off_bias() + 6 * f.slot_i;
        const double* bj = pp.state + pp.sl.off_bias() + 6 * f.slot_j;
        for (int c = 0; c < 3; c++) {
          X.x[c] = bi[c]; X.x[3 + c] = bj[c]; X.x[6 + c] = bi[3 + c]; X.x[9 + c] = bj[3 + c];
        }
        X.xcol[0] = sc.col_bias_gyro >= 0 ? sc.col_bias_gyro + 3 * f.slot_i : -1;
        X.xcol[1] = sc.col_bias_gyro >= 0 ? sc.col_bias_gyro + 3 * f.slot_j : -1;
        X.xcol[2] = sc.col_bias_accel >= 0 ? sc.col_bias_accel + 3 * f.slot_i : -1;
        X.xcol[3] = sc.col_bias_accel >= 0 ? sc.col_bias_accel + 3 * f.slot_j : -1;
      }
      const bool ok = small_dispatch(W, f, pre, pp, qid, mk(0, 0, 0), X, sc, o);
      if (ok) {
        double sq = 0.0;
        for (int r = 0; r < o->nres; r++) sq += o->r[r] * o->r[r];
        double rho = sq, scale = 1.0;
        if (f.loss_a > 0.0) scale = cauchy_scale(sq, f.loss_a, &rho);  // Corrector: rho'' < 0 -> pure sqrt(rho') scaling
        if (scale != 1.0) {
          for (int r = 0; r < o->nres; r++) {
            o->r[r] *= scale;
            for (int c = 0; c < o->ncols; c++) o->J[r][c] *= scale;
          }
        }
        s_cost[o->ncols > 0 ? 0 : 1] += 0.5 * rho;
      }
      s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok || !jac) continue;
    const int nc = o->ncols, nr = o->nres;
    for (int t = threadIdx.x; t < nc * nc + nc; t += blockDim.x) {
      if (t < nc * nc) {
        const int a = t / nc, c = t - a * nc;
        double v = 0.0;
        for (int r = 0; r < nr; r++) v += o->J[r][a] * o->J[r][c];
        H[(size_t)o->gcol[a] * D + o->gcol[c]] += v;
      } else {
        const int a = t - nc * nc;
        double v = 0.0;
        for (int r = 0; r < nr; r++) v += o->J[r][a] * o->r[r];
        b[o->gcol[a]] += v;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    cost2[0] += s_cost[0];
    cost2[1] += s_cost[1];
  }
}

// clic_evaluate_factor: one factor against the window with EVERY knot free (column 6 k of knot k) and the factor's own
// blocks after them; out = [ok, nres, -, -, r[16], J (nres x n_cols, row-major)].
__global__ void __launch_bounds__(128) factor_jacobian_kernel(SmallFactor f, const PreintData* pre, PassParams pp, Q4 S_CtoI,
                                                              V3 p_CinI, SmallX X, int n_cols, double* out) {
  extern __shared__ double smem[];
  double* rest;
  WindowView W = stage_window(smem, pp, &rest);
  SmallOut* o = reinterpret_cast<SmallOut*>(rest);
  int* kc = reinterpret_cast<int*>(o + 1);
  for (int k = threadIdx.x; k < pp.sl.K; k += blockDim.x) kc[k] = 6 * k;
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    SmallCols sc;
    sc.knot_col = kc;
    sc.col_toff_imu = -1;
    sc.lm_col = nullptr;
    sc.col_bias_gyro = sc.col_bias_accel = -1;
    s_ok = small_dispatch(W, f, pre, pp, S_CtoI, p_CinI, X, sc, o) ? 1 : 0;
    out[0] = (double)s_ok;
    out[1] = s_ok ? (double)o->nres : 0.0;
  }
  __syncthreads();
  if (!s_ok) return;
  const int nr = o->nres;
  for (int t = threadIdx.x; t < nr; t += blockDim.x) out[4 + t] = o->r[t];
  for (int t = threadIdx.x; t < nr * n_cols; t += blockDim.x) out[20 + t] = 0.0;
  __syncthreads();
  for (int t = threadIdx.x; t < nr * o->ncols; t += blockDim.x) {
    const int r = t / o->ncols, c = t - r * o->ncols;
    out[20 + (size_t)r * n_cols + o->gcol[c]] = o->J[r][c];
  }
}

// K0 at add time: double seconds -> int64 ns with the reference's truncation and window test
// (trajectory_estimator.cpp:272-292); AoS(3n) -> planar.
__global__ void prep_loam_kernel(int64_t n, const double* t_s, const double* point, const int32_t* geo_type,
                                 const double* geo_plane, const double* geo_normal, const double* geo_point,
                                 Q4 S_LtoI, V3 p_LinI, int64_t t_lo, int64_t t_hi, int64_t* t_ns, double* p0, double* p1, double* p2,
                                 double* g0, double* g1, double* g2, double* g3, double* g4, double* g5, uint8_t* type,
                                 unsigned long long* n_dropped, int* bad_input, const int32_t* gather) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // stored position
  if (o >= n) return;
  const int64_t i = gather ? gather[o] : o;                          // input record (-1: padding of a key-aligned batch)
  if (i < 0) {
    t_ns[o] = kDropped;
    p0[o] = p1[o] = p2[o] = 0.0;
    g0[o] = g1[o] = g2[o] = g3[o] = g4[o] = g5[o] = 0.0;
    type[o] = 1;
    return;
  }
  int64_t t = (int64_t)(t_s[i] * 1e9);
  bool ok = !(t < t_lo || t >= t_hi);
  t_ns[o] = ok ? t : kDropped;
  if (!ok) atomicAdd(n_dropped, 1ULL);
  const V3 pI = qrot(S_LtoI, mk(point[3 * i], point[3 * i + 1], point[3 * i + 2])) + p_LinI;  // LoamShared::point_in_imu
  p0[o] = pI.x; p1[o] = pI.y; p2[o] = pI.z;
  const bool plane = geo_type[i] == 1;
  type[o] = plane ? 1 : 0;
  if ((plane && !geo_plane) || (!plane && (!geo_normal || !geo_point))) {  // record type present, its array missing
    *bad_input = 1;
    t_ns[o] = kDropped;
    g0[o] = g1[o] = g2[o] = g3[o] = g4[o] = g5[o] = 0.0;
    return;
  }
  if (plane) {
    g0[o] = geo_plane[4 * i]; g1[o] = geo_plane[4 * i + 1]; g2[o] = geo_plane[4 * i + 2]; g3[o] = geo_plane[4 * i + 3];
    g4[o] = 0.0; g5[o] = 0.0;
  } else {
    g0[o] = geo_point[3 * i]; g1[o] = geo_point[3 * i + 1]; g2[o] = geo_point[3 * i + 2];
    g3[o] = geo_normal[3 * i]; g4[o] = geo_normal[3 * i + 1]; g5[o] = geo_normal[3 * i + 2];
  }
}

// geo_type == nullptr: a type-uniform batch (fixed_plane) whose geometry records are gstride (4 | 6) doubles long.
__global__ void prep_loam_packed_kernel(int64_t n, const double* t_s, const double* point, const uint8_t* geo_type,
                                        const double* geo, int gstride, int fixed_plane, Q4 S_LtoI, V3 p_LinI,
                                        int64_t t_lo, int64_t t_hi, int64_t* t_ns, double* p0,
                                        double* p1, double* p2, double* g0, double* g1, double* g2, double* g3,
                                        double* g4, double* g5, uint8_t* type, unsigned long long* n_dropped,
                                        const int32_t* gather) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // stored position
  if (o >= n) return;
  const int64_t i = gather ? gather[o] : o;                          // input record (-1: padding of a key-aligned batch)
  if (i < 0) {
    t_ns[o] = kDropped;
    p0[o] = p1[o] = p2[o] = 0.0;
    if (type) type[o] = (unsigned char)fixed_plane;
    g0[o] = g1[o] = g2[o] = g3[o] = 0.0;
    if (gstride > 4) { g4[o] = 0.0; g5[o] = 0.0; }
    return;
  }
  int64_t t = (int64_t)(t_s[i] * 1e9);
  bool ok = !(t < t_lo || t >= t_hi);
  t_ns[o] = ok ? t : kDropped;
  if (!ok) atomicAdd(n_dropped, 1ULL);
  const V3 pI = qrot(S_LtoI, mk(point[3 * i], point[3 * i + 1], point[3 * i + 2])) + p_LinI;  // LoamShared::point_in_imu
  p0[o] = pI.x; p1[o] = pI.y; p2[o] = pI.z;
  if (type) type[o] = geo_type ? (geo_type[i] == 1 ? 1 : 0) : (unsigned char)fixed_plane;
  const double* gi = geo + (size_t)gstride * i;
  g0[o] = gi[0]; g1[o] = gi[1]; g2[o] = gi[2]; g3[o] = gi[3];
  if (gstride > 4) { g4[o] = gi[4]; g5[o] = gi[5]; }
}

__global__ void prep_imu_kernel(int64_t n, const double* t_s, const double* gyro, const double* accel, int64_t t_lo,
                                int64_t t_hi, int64_t pad, int64_t* t_ns, double* g0, double* g1, double* g2,
                                double* a0, double* a1, double* a2, unsigned long long* n_dropped,
                                const int32_t* gather) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // stored position
  if (o >= n) return;
  const int64_t i = gather ? gather[o] : o;                          // input record (-1: padding of a key-aligned batch)
  if (i < 0) {
    t_ns[o] = kDropped;
    g0[o] = g1[o] = g2[o] = a0[o] = a1[o] = a2[o] = 0.0;
    return;
  }
  int64_t t = (int64_t)(t_s[i] * 1e9);
  bool ok = !(t - pad < t_lo || t + pad >= t_hi);
  t_ns[o] = ok ? t : kDropped;
  if (!ok) atomicAdd(n_dropped, 1ULL);
  g0[o] = gyro[3 * i]; g1[o] = gyro[3 * i + 1]; g2[o] = gyro[3 * i + 2];
  a0[o] = accel[3 * i]; a1[o] = accel[3 * i + 1]; a2[o] = accel[3 * i + 2];
}

// Index probe (clic_get_indices): s, u per factor of a stream at the current offset.
__global__ void index_probe_kernel(int64_t n, const int64_t* t_ns, int64_t toff, int64_t t0_ns, int64_t dt_ns, int K,
                                   int32_t* s_out, double* u_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = -1;
  double u = 0.0;
  int64_t t = t_ns[i];
  if (t == kDropped || !index_ns(t + toff, t0_ns, dt_ns, K, &s, &u)) { s = -1; u = 0.0; }
  s_out[i] = s;
  u_out[i] = u;
}

}  // namespace clic
