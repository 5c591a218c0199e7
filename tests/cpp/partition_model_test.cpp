// The host-side partition model of the fused pass (clic_b200/csrc/partition_model.hpp) on the CPU: grid split among the
// factor streams and the boundary-aware slab ranges of the CTAs.  Prints "PARTITION_OK" or "PARTITION_FAIL <what>".
#include <cstdio>
#include <numeric>

#include "../../clic_b200/csrc/partition_model.hpp"

using namespace clic::partition;

static int fail(const char* what) {
  std::printf("PARTITION_FAIL %s\n", what);
  return 1;
}

static std::vector<Stream> window(int64_t planes, int64_t lines, int64_t imu, int64_t vis_slabs, int n_sm) {
  auto mk = [&](int kind, int64_t slabs, int min_per_cta) {
    Stream s;
    s.kind = kind;
    s.slabs = std::max<int64_t>(1, slabs);
    s.cap = (int)std::max<int64_t>(1, std::min<int64_t>(n_sm, (s.slabs + min_per_cta - 1) / min_per_cta));
    s.w_scale = 1.0;
    return s;
  };
  return {mk(1, (planes + 31) / 32, 8), mk(2, (lines + 31) / 32, 8), mk(3, (imu + 31) / 32, 8), mk(4, vis_slabs, 4)};
}

int main() {
  const Model M = default_model();
  const int n_sm = 148;
  // time_of never grows with more CTAs
  for (const Stream& s : window(700403, 299597, 200000, 1300, n_sm))
    for (int n = 1; n < s.cap; n++)
      if (time_of(M, s, n + 1) > time_of(M, s, n) + 1e-12) return fail("time_of grows with n");
  // BASELINE C5 and its 1/2, 1/4, 1/8 shards: the whole grid is used, every stream within its cap, the modelled finishing
  // times within 10 % of each other (what the partition is for)
  for (int shard : {1, 2, 4, 8}) {
    const std::vector<Stream> js = window(700403 / shard, 299597 / shard, 200000 / shard, 1300 / shard + 8, n_sm);
    const std::vector<int> n = split_grid(M, js, n_sm);
    if ((int)n.size() != 4) return fail("size");
    if (std::accumulate(n.begin(), n.end(), 0) != n_sm) return fail("the grid is not used up");
    double tmin = 1e30, tmax = 0.0;
    for (size_t i = 0; i < js.size(); i++) {
      if (n[i] < 1 || n[i] > js[i].cap) return fail("cap");
      const double t = time_of(M, js[i], n[i]);
      tmin = std::min(tmin, t);
      tmax = std::max(tmax, t);
    }
    if (tmax > 1.12 * tmin) return fail("unbalanced partition");
    if (split_grid(M, js, n_sm) != n) return fail("not a pure function");
    if (shard == 1 && !(n[0] >= 48 && n[0] <= 56 && n[1] >= 22 && n[1] <= 28 && n[2] >= 56 && n[2] <= 64 && n[3] >= 9 && n[3] <= 13))
      return fail("C5 partition far from the measured optimum (52 / 25 / 60 / 11)");
  }
  // a tiny window: every stream gets at least one CTA, nobody more than its cap, leftovers stay unused
  {
    const std::vector<Stream> js = window(40, 40, 40, 1, n_sm);
    const std::vector<int> n = split_grid(M, js, n_sm);
    for (size_t i = 0; i < js.size(); i++)
      if (n[i] != 1) return fail("tiny window");
  }
  if (work(M, window(2000, 0, 200, 0, n_sm)) >= 6000.0) return fail("C1 must stay on the per-stream kernels");
  if (work(M, window(700403, 299597, 200000, 1300, n_sm)) < 6000.0) return fail("C5 must take the fused pass");
  // slab ranges: even without boundaries; monotone, complete; a CTA with a boundary gets fewer slabs
  {
    const int64_t S = 6250;
    const int n = 60;
    const std::vector<int64_t> even = cta_slab_ends(3, S, n, {});
    for (int c = 0; c < n; c++)
      if (even[c] != S * (c + 1) / n) return fail("even split");
    std::vector<int32_t> kb;
    for (int k = 1; k <= 6; k++) kb.push_back((int32_t)(S * k / 7));
    const std::vector<int64_t> end = cta_slab_ends(3, S, n, kb);
    if (end.back() != S) return fail("last end");
    int with = 0;
    double len_with = 0.0, len_without = 0.0;
    for (int c = 0; c < n; c++) {
      const int64_t lo = c ? end[c - 1] : 0;
      if (end[c] <= lo) return fail("empty or non-monotone range");
      bool has = false;
      for (int32_t b : kb) has = has || (b >= lo && b < end[c]);
      (has ? len_with : len_without) += (double)(end[c] - lo);
      with += has ? 1 : 0;
    }
    if (with != 6) return fail("every boundary in its own CTA");
    len_with /= with;
    len_without /= (n - with);
    if (!(len_without - len_with > 3.5 && len_without - len_with < 6.5)) return fail("IMU boundary CTAs should get ~5 slabs less");
    const std::vector<int64_t> lidar = cta_slab_ends(1, 21888, 52, kb);
    if (lidar.back() != 21888) return fail("lidar last end");
    if (cta_slab_ends(4, 1300, 11, kb) != cta_slab_ends(4, 1300, 11, {})) return fail("visual ranges ignore the hint");
    if (cta_slab_ends(3, 5, 1, kb)[0] != 5) return fail("single CTA");
    // too few slabs for the correction: even split
    if (cta_slab_ends(3, 100, 60, kb) != cta_slab_ends(3, 100, 60, {})) return fail("small stream");
  }
  std::printf("PARTITION_OK\n");
  return 0;
}
