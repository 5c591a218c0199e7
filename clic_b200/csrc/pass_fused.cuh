// pass_kernel: the whole fused residual + Jacobian + normal-equation pass of a solve — and, on several GPUs, its one
// exchange step — in ONE persistent launch.
//
//   phase F  every factor stream of the pass runs on its own partition of the grid (one CTA per SM; the partition
//            sizes and every CTA's slab range come from a static cost model on the host, so the grouping of the sums —
//            and with it every bit of H, b — is a function of the problem alone).  The stream bodies are the ones of loam_kernel /
//            imu_kernel / image_kernel (kernels.cuh).
//   barrier  (the only grid-wide one)
//            ... at the end of which every CTA has added what it accumulated into ITS OWN copy of the packed system
//            [upper triangle of H | b | cost, fixed cost] in shared memory (cta_finish_dense, kernels.cuh) and writes
//            it out contiguously (21.6 KB at D = 72).
//   phase A  owner-computes assembly: CTA c owns a contiguous chunk of the packed system and sums that chunk over the
//            producing CTAs — identically laid out vectors, coalesced loads, nothing to look up — in a fixed order.
//            (Round 1 reduced per-warp records per interval and assembled the blocks afterwards: two dependent kernels,
//            ~35 us; summing the record entries per target entry directly cost as much in scattered 8-byte sector
//            traffic.)
//   phase X  (several ranks, peer memory mapped) the chunk is pushed straight into slot `rank` of every peer's exchange
//            area over NVLink, every double as two 8-byte words (32 data bits | 32-bit sequence number): the data is its
//            own arrival flag.  CTA c polls the words of CTA c of every peer — all ranks partition the system
//            identically, so no grid-wide synchronisation is involved — and sums the slots in rank order: every rank
//            ends up with bit-identical H, b, cost.  The wait is bounded in wall-clock time (zeroed result + error).
//
// Replaces  loam/imu/image kernels -> reduce_all_kernel -> assemble_kernel -> p2p_allreduce_kernel  (round 1: ~60 us of
// serial tail that did not shrink with the number of GPUs).  Grid barriers need every CTA resident: the grid never
// exceeds one CTA per SM and the kernel is launched cooperatively.
#pragma once
#include "kernels.cuh"

namespace clic {

constexpr int kFusedKinds = 6;  // 0 LiDAR mixed, 1 LiDAR planes, 2 LiDAR lines, 3 IMU, 4 visual (pose packs), 5 visual (chain)
constexpr int kP2PMaxCtas = 512;  // per-CTA flags of the fused exchange live after the per-rank flags of the exchange area
constexpr size_t kP2PCtaFlagOffset = 256;  // bytes
static_assert(kP2PCtaFlagOffset + (size_t)kP2PMaxRanks * kP2PMaxCtas * sizeof(unsigned long long) <= kP2PFlagBytes,
              "per-CTA flags must fit the flag area");

struct FusedPart {
  int cta0, n_ctas, total_slabs;  // n_ctas == 0: no stream of this kind in the pass
  int desc;                       // index of the stream's column descriptor
  RecordBuf rb;                   // slots of records flushed in mid-run (read back by the same CTA) + atomics fallback
};
constexpr int kMaxFusedJobs = 5;
constexpr int kMaxFusedCtas = 160;  // one CTA per SM (148 on B200)
constexpr int kKeyColsInts = 1 + 3 * 64;
constexpr int kMaxFusedBias = 4;
__host__ __device__ inline size_t fused_persist_bytes(int K, int nb) {
  return ((sizeof(StreamDesc) + 15) & ~(size_t)15) + (((size_t)K * sizeof(int) + 15) & ~(size_t)15) +
         kMaxFusedBias * ((sizeof(BiasFactorDesc) + 15) & ~(size_t)15) + (size_t)6 * nb * sizeof(double) + 16;
}

// Everything a pass needs that does not change between passes of one layout; passed by value (__grid_constant__) so
// that the stream descriptors stay immediate constant-bank operands inside the factor bodies, as in the per-stream
// kernels.
struct FusedArgs {
  LoamStream loam[3];  // by geo_kind
  ImuStream imu;
  ImageStream image;
  ImageShared ish;
  FusedPart part[kFusedKinds];
  int n_factor_ctas;
  int hs_off;               // offset (doubles) of the CTA's packed system inside the dynamic shared memory
  int persist_off;          // offset (doubles) of the region that lives through all phases: [StreamDesc | knot_col K |
                            // BiasFactorDesc x n_bias | bias state 6 nb]
  double* hloc;             // [n_factor_ctas][n_ent + 2] the CTAs' packed systems
  ColMaps cm;
  const StreamDesc* descs;      // one column descriptor per stream
  const StreamDesc* descs_ovf;  // the same with fin = the stream's atomics-fallback blocks
  int n_jobs;
  const BiasFactorDesc* bias;
  int n_bias, off_bias;
  const double* rho_side;
  int rho0, n_rho;
  // grid barrier: one generation word per CTA; ovf_flag: "a record went to the atomics fallback in this pass"
  unsigned* bar;
  unsigned* ovf_flag;
  double* ovf_blocks[kMaxFusedJobs];
  int ovf_doubles[kMaxFusedJobs];
  // per-CTA phase stamps (globaltimer ns) when not null: [cta][8] = start, end of F, after the barrier, tile loaded,
  // end of A, end
  unsigned long long* stamps;
  // last slab (exclusive) of every factor CTA within its stream; a partition's first CTA starts at slab 0 (build_fused)
  int cta_slab_end[kMaxFusedCtas];
};

// per-pass arguments
struct FusedOut {
  double* H;      // dense D x D (this rank's sum; the global sum when there is no exchange in the kernel)
  double* b;
  double* cost2;  // cost, fixed cost
  const double* addend;  // nullptr, or [H D*D | b D | cost 2] terms added by other kernels (prior, velocity, bias factors)
  // exchange (n_ranks > 1 and peer memory): destination of the all-reduced system
  double* Hg;
  double* bg;
  double* cost2g;
  unsigned long long seq;
  unsigned long long timeout_ns;
  int* err;
  int exchange;   // 1: phase X runs
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Grid barrier on one generation word per CTA: arriving is a plain release store to the CTA's own word (no contended
// atomic: 148 same-address atomics cost ~3 us), waiting is warp 0 polling all words, every lane its share with all loads
// in flight.  Invariant between passes: every word holds the same generation (a pass skipped by the control word touches
// none), so a CTA learns the generation from its own word — `gen`, read at kernel start.  Release / acquire through the
// words makes the global writes of every CTA before the barrier visible to every CTA after it.  Every CTA of the grid
// calls it the same number of times, with gen advanced by one each time.
constexpr int kBarMaxPerLane = 8;  // grids of up to 256 CTAs
__device__ __forceinline__ void grid_barrier(unsigned* flags, unsigned gen) {
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const unsigned target = gen + 1u;
    if (lane == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(target) : "memory");
    }
    const int n = gridDim.x;
    while (true) {
      unsigned g[kBarMaxPerLane];
#pragma unroll
      for (int q = 0; q < kBarMaxPerLane; q++) {
        const int c = q * 32 + lane;
        g[q] = target;
        if (c < n) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g[q]) : "l"(flags + c) : "memory");
      }
      bool ok = true;
#pragma unroll
      for (int q = 0; q < kBarMaxPerLane; q++) ok = ok && (int)(g[q] - target) >= 0;
      if (__all_sync(0xffffffffu, ok)) break;
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ unsigned long long* p2p_cta_flag(char* area, int src_rank, int cta) {
  return reinterpret_cast<unsigned long long*>(area + kP2PCtaFlagOffset) + (size_t)src_rank * kP2PMaxCtas + cta;
}

// Lane-0 terms of one entry that do not come from factor streams: inverse-depth side rows and the BiasFactors.
__device__ __forceinline__ double entry_side_terms(int gi, int gj, const FusedArgs& fa, const BiasFactorDesc* bias,
                                                   const double* bias_state) {
  const int D = fa.cm.D;
  const bool is_b = gj < 0;
  double acc = 0.0;
  if (fa.rho_side) {  // entries of the inverse-depth columns (visual factors with a free / marginalised landmark)
    const int rho0 = fa.rho0, n_rho = fa.n_rho;
    const bool ri = gi >= rho0 && gi < rho0 + n_rho, rj = !is_b && gj >= rho0 && gj < rho0 + n_rho;
    const size_t ld = (size_t)D + 2;
    if (is_b) {
      if (ri) acc += fa.rho_side[(size_t)(gi - rho0) * ld + D + 1];
    } else if (ri && rj) {
      if (gi == gj) acc += fa.rho_side[(size_t)(gi - rho0) * ld + D];
    } else if (rj) {
      acc += fa.rho_side[(size_t)(gj - rho0) * ld + gi];
    } else if (ri) {
      acc += fa.rho_side[(size_t)(gi - rho0) * ld + gj];
    }
  }
  for (int f = 0; f < fa.n_bias; f++) {  // BiasFactor (trajectory_value_factor.h:65-126): +-I Jacobians
    const BiasFactorDesc& bf = bias[f];
    const double* bi = bias_state + 6 * bf.slot_i;
    const double* bj = bias_state + 6 * bf.slot_j;
    for (int r = 0; r < 6; r++) {
      const double w = bf.sqrt_info[r], res = w * (bj[r] - bi[r]);
      const int base_i = r < 3 ? bf.col_gi : bf.col_ai, base_j = r < 3 ? bf.col_gj : bf.col_aj;
      const int c_i = base_i >= 0 ? base_i + r % 3 : -1, c_j = base_j >= 0 ? base_j + r % 3 : -1;
      if (is_b) {
        if (c_i >= 0 && gi == c_i) acc += -w * res;
        if (c_j >= 0 && gi == c_j) acc += w * res;
      } else {
        if (c_i >= 0 && gi == c_i && gj == c_i) acc += w * w;
        if (c_j >= 0 && gi == c_j && gj == c_j) acc += w * w;
        if (c_i >= 0 && c_j >= 0 && gi == min(c_i, c_j) && gj == max(c_i, c_j)) acc += -w * w;
      }
    }
  }
  return acc;
}
__device__ __forceinline__ double bias_factor_cost(const FusedArgs& fa, const BiasFactorDesc* bias, const double* bias_state,
                                                   int which) {
  double bias_cost = 0.0;
  bool bias_free = false;
  for (int f = 0; f < fa.n_bias; f++) {
    const BiasFactorDesc& bf = bias[f];
    const double* bi = bias_state + 6 * bf.slot_i;
    const double* bj = bias_state + 6 * bf.slot_j;
    bias_free = bias_free || bf.col_gi >= 0 || bf.col_gj >= 0 || bf.col_ai >= 0 || bf.col_aj >= 0;
    double tot = 0.0;
    for (int r = 0; r < 6; r++) {
      const double res = bf.sqrt_info[r] * (bj[r] - bi[r]);
      tot += res * res;
    }
    bias_cost += 0.5 * tot;
  }
  return (fa.n_bias > 0 && (bias_free ? 0 : 1) == which) ? bias_cost : 0.0;
}

__device__ __forceinline__ size_t n_ent_full_stride(const FusedArgs& fa) {
  const int D = fa.cm.D;
  return (size_t)D * (D + 1) / 2 + D + 2;
}

template <bool JAC>
__global__ void __launch_bounds__(kThreads, 1)
pass_kernel(const __grid_constant__ FusedArgs fa, const __grid_constant__ PassParams pp, const __grid_constant__ FusedOut out,
            const __grid_constant__ P2PView pv) {
  if (pass_prologue(pp)) return;  // uniform over the grid: nobody reaches the barrier
  extern __shared__ double smem[];
  const int cta = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long* stamp = fa.stamps ? fa.stamps + 8 * cta : nullptr;
  if (stamp && threadIdx.x == 0) stamp[0] = globaltimer_ns();
  unsigned bar_gen = 0;
  if (threadIdx.x == 0)  // the barrier generation, read early: the load is long done when the barrier needs it
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(bar_gen) : "l"(fa.bar + cta) : "memory");
  const int D = fa.cm.D;
  const int n_ent = JAC ? D * (D + 1) / 2 + D : 0;
  const int n_tot = n_ent + 2;

  // the region of shared memory that lives through all phases: copies of what the dense finish and the assembly read per
  // entry — from global memory each of those reads is a dependent cache miss right after the L2 was flushed
  char* persist = reinterpret_cast<char*>(smem + fa.persist_off);
  StreamDesc* p_sd = reinterpret_cast<StreamDesc*>(persist);
  int* p_knot_col = reinterpret_cast<int*>(persist + ((sizeof(StreamDesc) + 15) & ~(size_t)15));
  BiasFactorDesc* p_bias = reinterpret_cast<BiasFactorDesc*>(reinterpret_cast<char*>(p_knot_col) +
                                                             (((size_t)pp.sl.K * sizeof(int) + 15) & ~(size_t)15));
  double* p_bstate = reinterpret_cast<double*>(reinterpret_cast<char*>(p_bias) +
                                               kMaxFusedBias * ((sizeof(BiasFactorDesc) + 15) & ~(size_t)15));
  {
    int my_desc = -1;
#pragma unroll
    for (int k = 0; k < kFusedKinds; k++)
      if (fa.part[k].n_ctas && cta >= fa.part[k].cta0 && cta < fa.part[k].cta0 + fa.part[k].n_ctas) my_desc = fa.part[k].desc;
    if (my_desc >= 0)
      for (int i = threadIdx.x; i < (int)(sizeof(StreamDesc) / sizeof(int)); i += kThreads)
        reinterpret_cast<int*>(p_sd)[i] = reinterpret_cast<const int*>(fa.descs + my_desc)[i];
    for (int i = threadIdx.x; i < pp.sl.K; i += kThreads) p_knot_col[i] = fa.cm.knot_col[i];
    for (int i = threadIdx.x; i < fa.n_bias * (int)(sizeof(BiasFactorDesc) / sizeof(int)); i += kThreads)
      reinterpret_cast<int*>(p_bias)[i] = reinterpret_cast<const int*>(fa.bias)[i];
    if (fa.n_bias > 0)
      for (int i = threadIdx.x; i < 6 * pp.sl.nb; i += kThreads) p_bstate[i] = pp.state[fa.off_bias + i];
  }

  // ---- phase F: the factor streams, each on its partition; every CTA fills its own packed system in shared memory ----
  if (cta < fa.n_factor_ctas) {
    double* Hs = smem + fa.hs_off;
    for (int i = threadIdx.x; i < n_tot; i += kThreads) Hs[i] = 0.0;
    double* rest;
    WindowView W = stage_window(smem, pp, &rest);  // (block barriers inside: Hs is zero for everybody afterwards)
    DenseOut dn;
    dn.Hs = Hs;
    dn.n_ent = n_ent;
    dn.D = D;
    dn.knot_col = p_knot_col;
    dn.key_cols = reinterpret_cast<int*>(Hs + ((n_tot + 1) & ~1));
    dn.stamp = stamp;
    {
      bool first = false;
#pragma unroll
      for (int k = 0; k < kFusedKinds; k++) first = first || (fa.part[k].n_ctas && cta == fa.part[k].cta0);
      dn.slab_lo = first ? 0 : fa.cta_slab_end[cta - 1];
      dn.slab_hi = fa.cta_slab_end[cta];
    }
    if (fa.part[1].n_ctas && cta >= fa.part[1].cta0 && cta < fa.part[1].cta0 + fa.part[1].n_ctas) {
      dn.sd = p_sd;
      loam_body<JAC, kWarpsPerCta, 1, true>(fa.loam[1], pp, fa.part[1].total_slabs, fa.part[1].rb, W, rest,
                                            cta - fa.part[1].cta0, fa.part[1].n_ctas, &dn);
    } else if (fa.part[2].n_ctas && cta >= fa.part[2].cta0 && cta < fa.part[2].cta0 + fa.part[2].n_ctas) {
      dn.sd = p_sd;
      loam_body<JAC, kWarpsPerCta, 2, true>(fa.loam[2], pp, fa.part[2].total_slabs, fa.part[2].rb, W, rest,
                                            cta - fa.part[2].cta0, fa.part[2].n_ctas, &dn);
    } else if (fa.part[0].n_ctas && cta >= fa.part[0].cta0 && cta < fa.part[0].cta0 + fa.part[0].n_ctas) {
      dn.sd = p_sd;
      loam_body<JAC, kWarpsPerCta, 0, true>(fa.loam[0], pp, fa.part[0].total_slabs, fa.part[0].rb, W, rest,
                                            cta - fa.part[0].cta0, fa.part[0].n_ctas, &dn);
    } else if (fa.part[3].n_ctas && cta >= fa.part[3].cta0 && cta < fa.part[3].cta0 + fa.part[3].n_ctas) {
      dn.sd = p_sd;
      imu_body<JAC, false, true>(fa.imu, pp, fa.part[3].total_slabs, fa.part[3].rb, W, rest, cta - fa.part[3].cta0,
                                 fa.part[3].n_ctas, &dn);
    } else if (fa.part[4].n_ctas && cta >= fa.part[4].cta0 && cta < fa.part[4].cta0 + fa.part[4].n_ctas) {
      dn.sd = p_sd;
      image_body<JAC, true, true>(fa.image, pp, fa.part[4].total_slabs, fa.part[4].rb, fa.ish, W, rest,
                                  cta - fa.part[4].cta0, fa.part[4].n_ctas, &dn);
    }  // kind 5 (visual, per-factor chain) is never fused: build_fused keeps such a pass on the per-stream kernels
    __syncthreads();
    double* dst = fa.hloc + (size_t)cta * (n_ent_full_stride(fa));
    for (int i = threadIdx.x; i < n_tot; i += kThreads) dst[i] = Hs[i];
  }
  if (stamp && threadIdx.x == 0) stamp[1] = globaltimer_ns();
  bar_gen = __shfl_sync(0xffffffffu, bar_gen, 0);  // (only warp 0 uses it)
  grid_barrier(fa.bar, bar_gen);
  if (stamp && threadIdx.x == 0) stamp[2] = globaltimer_ns();

  // ---- phase A (+ X): every CTA owns a contiguous chunk of the packed system [upper(H) | b | cost, fixed cost] ----
  const int epc = (n_tot + gridDim.x - 1) / gridDim.x;
  const int e0 = min(n_tot, cta * epc), e1 = min(n_tot, e0 + epc);
  const int my_n = e1 - e0;
  const int P = fa.n_factor_ctas;
  // shared memory: [stage: epc doubles][st_g: 2 ints per staged value][tile: P x my_n doubles]
  double* stage = smem;
  int* st_g = reinterpret_cast<int*>(stage + epc);   // row, column (-1: b; -2: cost; -3: fixed cost)
  double* tile = stage + 2 * epc;
  const unsigned ovf_word = *reinterpret_cast<volatile unsigned*>(fa.ovf_flag);
  const bool ovf = JAC && ovf_word != 0u;
  if (my_n > 0) {
    const size_t stride = n_ent_full_stride(fa);
    const int total = P * my_n;
    constexpr int PER = 12;  // loads in flight per thread
    for (int base = 0; base < total; base += PER * kThreads) {
      double v[PER];
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int i = base + q * kThreads + threadIdx.x;
        const int p = i / my_n, k = i - p * my_n;
        v[q] = i < total ? __ldcg(fa.hloc + (size_t)p * stride + e0 + k) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < PER; q++) {
        const int i = base + q * kThreads + threadIdx.x;
        if (i < total) tile[i] = v[q];
      }
    }
    __syncthreads();
    if (stamp && threadIdx.x == 0) stamp[3] = globaltimer_ns();
    // 8 threads per entry: thread j sums the producers p = j, j + 8, ... in ascending order, then a fixed 3-level tree
    for (int t = threadIdx.x; t < ((my_n * 8 + 31) & ~31); t += kThreads) {
      const int k = t >> 3, j = t & 7;
      double a = 0.0;
      if (k < my_n)
        for (int p = j; p < P; p += 8) a += tile[p * my_n + k];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if (j == 0 && k < my_n) {
        const int e = e0 + k;
        int gi = 0, gj = -2 - (e - n_ent);  // cost words: -2, -3
        if (e < n_ent) {
          const int n_upper = D * (D + 1) / 2;
          if (e >= n_upper) {
            gi = e - n_upper;
            gj = -1;
          } else {  // unrank (gi <= gj): rows before gi hold gi*D - gi(gi-1)/2 entries
            const double q = 2.0 * D + 1.0;
            gi = (int)((q - sqrt(q * q - 8.0 * (double)e)) * 0.5);
            gi = max(0, min(D - 1, gi));
            while (gi > 0 && gi * D - gi * (gi - 1) / 2 > e) gi--;
            while ((gi + 1) * D - (gi + 1) * gi / 2 <= e) gi++;
            gj = gi + (e - (gi * D - gi * (gi - 1) / 2));
          }
          a += entry_side_terms(gi, gj, fa, p_bias, p_bstate);
          if (out.addend) a += gj >= 0 ? out.addend[(size_t)gi * D + gj] : out.addend[(size_t)D * D + gi];
        } else {
          const int which = e - n_ent;
          a += bias_factor_cost(fa, p_bias, p_bstate, which);
          if (out.addend) a += out.addend[(size_t)D * D + D + which];
        }
        stage[k] = a;
        st_g[2 * k] = gi;
        st_g[2 * k + 1] = gj;
      }
    }
    __syncthreads();
    if (ovf) {  // unsorted input beyond the slots went to per-interval blocks with atomics: add them (rare path)
      for (int k = warp; k < my_n; k += kWarpsPerCta) {
        const int e = e0 + k;
        if (e >= n_ent) continue;
        int gi, gj;
        double bc;
        bool bf;
        const double v = assemble_entry(e, lane, fa.cm, fa.descs_ovf, fa.n_jobs, nullptr, 0, pp.state, fa.off_bias, nullptr,
                                        0, 0, &gi, &gj, &bc, &bf);
        if (lane == 0) stage[k] += v;
      }
      __syncthreads();
    }
  }
  if (stamp && threadIdx.x == 0) stamp[4] = globaltimer_ns();
  // this rank's own system is always written out (the NCCL fallback reduces it; tests read it)
  for (int i = threadIdx.x; i < my_n; i += blockDim.x) {
    const double v = stage[i];
    const int gi = st_g[2 * i], gj = st_g[2 * i + 1];
    if (gj >= 0) {
      out.H[(size_t)gi * D + gj] = v;
      out.H[(size_t)gj * D + gi] = v;
    } else if (gj == -1) {
      out.b[gi] = v;
    } else {
      out.cost2[-2 - gj] = v;
    }
  }
  if (out.exchange && my_n > 0) {
    // Low-latency exchange: every double travels as two 8-byte words (32 data bits | 32-bit sequence number), so the
    // data is its own arrival flag — no system-scope fence and no separate flag round trip (an 8-byte store is
    // single-copy atomic over NVLink).  CTA c pushes its chunk into slot `rank` of every peer, then polls slot r of its
    // own area for every rank r's chunk c and sums in rank order: same bits on every rank.  Slots are double-buffered by
    // the parity of the sequence number (a rank can be at most one exchange ahead of its slowest peer).
    const int parity = (int)(out.seq & 1ull);
    const unsigned long long tag = (out.seq & 0xffffffffull) << 32;
    for (int i = threadIdx.x; i < my_n * pv.n; i += blockDim.x) {
      const int r = i / my_n, k = i - r * my_n;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(stage[k]);
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(p2p_slot(pv.peer[r], parity, pv.rank)) + 2 * (e0 + k);
      const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
      asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
    }
    bool timed_out = false;
    double acc = 0.0;
    if ((int)threadIdx.x < my_n) {
      const unsigned long long* base = reinterpret_cast<const unsigned long long*>(p2p_slot(pv.peer[pv.rank], parity, 0)) +
                                       2 * (e0 + threadIdx.x);
      const unsigned long long t0 = globaltimer_ns();
      for (int r = 0; r < pv.n; r++) {
        const unsigned long long* src = base + (size_t)r * kP2PMaxDoubles;  // slot r of my own area
        unsigned long long w0, w1;
        while (true) {
          asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
          if ((w0 >> 32) == (tag >> 32) && (w1 >> 32) == (tag >> 32)) break;
          if (globaltimer_ns() - t0 > out.timeout_ns) { timed_out = true; break; }
        }
        if (timed_out) break;
        acc += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
      }
    }
    const int any_timeout = __syncthreads_or(timed_out ? 1 : 0);
    if (any_timeout && threadIdx.x == 0) *out.err = 1;
    if ((int)threadIdx.x < my_n) {
      if (any_timeout) acc = 0.0;  // a missing peer: the result is zeroed (never a partial sum) and the error word is set
      const int i = threadIdx.x;
      const int gi = st_g[2 * i], gj = st_g[2 * i + 1];
      if (gj >= 0) {
        out.Hg[(size_t)gi * D + gj] = acc;
        out.Hg[(size_t)gj * D + gi] = acc;
      } else if (gj == -1) {
        out.bg[gi] = acc;
      } else {
        out.cost2g[-2 - gj] = acc;
      }
    }
  }
  if (ovf) {  // re-arm the atomics-fallback blocks once everybody has read them
    grid_barrier(fa.bar, bar_gen + 1u);
    for (int j = 0; j < fa.n_jobs; j++) {
      const size_t n = (size_t)fa.ovf_doubles[j];
      for (size_t i = (size_t)cta * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        fa.ovf_blocks[j][i] = 0.0;
    }
    if (cta == 0 && threadIdx.x == 0) *fa.ovf_flag = 0u;
  }
  if (stamp && threadIdx.x == 0) stamp[5] = globaltimer_ns();
  pass_epilogue(pp);
}

}  // namespace clic
