// Host-side cost model of the fused pass (pass_fused.cuh): how the grid is divided among the factor streams of a pass and
// which slabs every CTA gets.  Plain C++ (no CUDA): build_fused (clic_cuda.cu) calls it, tests/cpp/partition_model_test.cpp
// exercises it on the CPU.  Everything here is a pure function of its arguments — the partition decides the grouping of
// the sums, so it must not depend on anything but the problem (bit-reproducible H, b).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace clic {
namespace partition {

constexpr int kKinds = 6;   // 0 LiDAR mixed, 1 LiDAR planes, 2 LiDAR lines, 3 IMU, 4 visual (pose packs), 5 visual (chain)
constexpr int kWarps = 8;   // warps per CTA (one CTA per SM)

// A stream on n CTAs takes  f + w * max(p, lat * ceil(p)) + ramp * min(p, ramp_to)  microseconds, p = slabs / (8 n) the
// slabs per warp: w per 32-factor slab (16 for the visual stream) of one warp with two warps per SM sub-partition — a
// CTA's 8 warps share one FP64 pipe, so the throughput term turns into a latency term (lat * w per whole round: the time
// a warp needs for a slab when it has the SM to itself) only at 1-2 slabs per warp —, f the stream's fixed part (window
// staging, pose packs, first-slab latency, the dense finish), ramp: the first ramp_to slabs of a warp cost more (the
// visual stream: nearly every early slab of a warp opens a new interval pair, i.e. one more record for the CTA's dense
// finish; later slabs repeat pairs).  Fitted to %globaltimer stamps of the C5 window and of its 1/2, 1/4, 1/8 shards on
// B200 (bench_micro/fused_probe.py, cta_ends.py; profiles/).  Only the ratios matter, and a wrong model costs balance,
// never correctness.
struct Model {
  double w[kKinds], f[kKinds], ramp[kKinds], ramp_to[kKinds];
  double lat;
};
inline Model default_model() {
  return Model{{3.6, 3.04, 3.32, 11.15, 7.3, 12.0},
               {14.0, 12.3, 14.8, 27.5, 31.2, 50.0},
               {0.0, 0.0, 0.0, 0.0, 15.0, 0.0},
               {0.0, 0.0, 0.0, 0.0, 1.8, 0.0},
               0.4};
}

struct Stream {
  int kind;
  int64_t slabs;    // total slabs of the stream (>= 1)
  int cap;          // most CTAs the stream may take (its record buffers are sized for that many)
  double w_scale;   // multiplies w (IMU with a free time offset: 1.25)
};

inline double time_of(const Model& M, const Stream& s, int n) {
  const double per_warp = (double)s.slabs / (double)(kWarps * n);
  const double rounds = (double)((s.slabs + (int64_t)kWarps * n - 1) / ((int64_t)kWarps * n));
  return M.f[s.kind] + M.w[s.kind] * s.w_scale * std::max(per_warp, M.lat * rounds) +
         M.ramp[s.kind] * std::min(per_warp, M.ramp_to[s.kind]);
}

// Modelled warp-time of the pass (the automatic choice between the fused pass and the per-stream kernels).
inline double work(const Model& M, const std::vector<Stream>& js) {
  double v = 0.0;
  for (const Stream& s : js) v += M.w[s.kind] * s.w_scale * (double)s.slabs;
  return v;
}

// CTAs per stream: min-max under the model — the smallest finishing time T for which every stream fits,
// n_j(T) = fewest CTAs with time_of <= T — then the leftover CTAs (and the whole grid when the model found no fit) one
// by one to the stream that would finish last.  Every stream gets at least one CTA; the sum never exceeds n_sm when
// js.size() <= n_sm.
inline std::vector<int> split_grid(const Model& M, const std::vector<Stream>& js, int n_sm) {
  std::vector<int> out(js.size(), 1);
  auto need = [&](double T, std::vector<int>& n) -> bool {
    int tot = 0;
    for (size_t i = 0; i < js.size(); i++) {
      if (time_of(M, js[i], js[i].cap) > T) return false;
      int lo = 1, hi = js[i].cap;
      while (lo < hi) {  // time_of is non-increasing in n
        const int mid = (lo + hi) / 2;
        if (time_of(M, js[i], mid) <= T) hi = mid; else lo = mid + 1;
      }
      n[i] = lo;
      tot += lo;
    }
    return tot <= n_sm;
  };
  double lo = 0.0, hi = 0.0;
  for (const Stream& s : js) hi = std::max(hi, time_of(M, s, 1));
  std::vector<int> n(js.size(), 1), best;
  if (need(hi, n)) best = n;
  for (int it = 0; it < 48 && !best.empty(); it++) {  // feasibility is monotone in T
    const double mid = 0.5 * (lo + hi);
    if (need(mid, n)) { best = n; hi = mid; } else { lo = mid; }
  }
  if (!best.empty()) out = best;
  int used = 0;
  for (int v : out) used += v;
  while (used < n_sm) {
    int pick = -1;
    double worst = 0.0;
    for (size_t i = 0; i < js.size(); i++) {
      if (out[i] >= js[i].cap) continue;
      const double t = time_of(M, js[i], out[i]);
      if (pick < 0 || t > worst) { pick = (int)i; worst = t; }
    }
    if (pick < 0) break;
    out[pick]++;
    used++;
  }
  return out;
}

// Last slab (exclusive) of every CTA of a stream's partition.  Even split, except that a CTA which contains an interval
// boundary of a time-sorted LiDAR / IMU stream gets fewer slabs: it has a record flushed in mid-run plus a second live
// group to finish (measured + 4 us LiDAR, + 7 us IMU: these CTAs were the stragglers of their partitions) — worth
// 10 / 5 slabs.  key_change: slabs at which the stream enters the next knot interval (host estimate from the caller's
// timestamps; empty = unknown).  Three fixed-point rounds (which CTA holds a boundary depends on the split).
inline std::vector<int64_t> cta_slab_ends(int kind, int64_t slabs, int n, const std::vector<int32_t>& key_change) {
  const double pen = kind <= 2 ? 10.0 : (kind == 3 ? 5.0 : 0.0);
  std::vector<int64_t> end((size_t)n);
  for (int c = 0; c < n; c++) end[c] = slabs * (c + 1) / n;
  if (pen > 0.0 && !key_change.empty() && n > 1 && slabs > 4 * (int64_t)n) {
    for (int round = 0; round < 3; round++) {
      std::vector<int> nb((size_t)n, 0);
      int total = 0;
      for (int32_t b : key_change) {
        const int c = (int)(std::upper_bound(end.begin(), end.end(), (int64_t)b) - end.begin());
        if (c < n) { nb[c]++; total++; }
      }
      const double T = ((double)slabs + pen * total) / n;
      double acc = 0.0;
      for (int c = 0; c < n; c++) {
        acc += std::max(1.0, T - pen * nb[c]);
        end[c] = std::min<int64_t>(slabs, (int64_t)(acc + 0.5));
      }
      end[n - 1] = slabs;
    }
  }
  return end;
}

}  // namespace partition
}  // namespace clic
