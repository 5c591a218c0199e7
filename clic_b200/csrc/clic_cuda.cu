// libclic_cuda.so — the C ABI of include/clic_cuda.h on top of the sm_100a kernels.  Host code here only
// stages inputs, launches kernels and copies results; there is no CPU implementation of any arithmetic of
// the path (no fallback: without a CUDA device every compute entry point returns CLIC_ERR_NO_DEVICE).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/clic_cuda.h"
#include "kernels.cuh"
#include "pass_fused.cuh"
#include "partition_model.hpp"
#include "solver.cuh"
#include "assoc.cuh"

using namespace clic;

#ifndef CLIC_LOAM_WARPS
#define CLIC_LOAM_WARPS 8
#endif
#ifndef CLIC_LOAM_MINB
#define CLIC_LOAM_MINB 1
#endif

namespace {

constexpr int kLoamWarps = CLIC_LOAM_WARPS, kLoamMinB = CLIC_LOAM_MINB;

thread_local std::string g_err;
int fail(int code, const std::string& m) {
  g_err = m;
  return code;
}
#define CU(x)                                                                                          \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess)                                                                             \
      return fail(CLIC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " @" + std::to_string(__LINE__)); \
  } while (0)

// ---- NCCL, resolved at run time (the library must load on boxes without NCCL; torch's bundled libnccl.so.2 is
// already mapped in a torchrun process and is what dlopen returns) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (int (*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (int (*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
      api.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
      api.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce;
    }
  }
  return api;
}
// One communicator per process, shared by every estimator (an estimator is built per solve; ncclCommInitRank is not).
struct ProcessComm {
  ncclComm_t comm = nullptr;
  int n_ranks = 1, rank = 0;
} g_comm;
constexpr int kNcclDouble = 8, kNcclSum = 0;

// Peer-memory exchange areas for the one-shot all-reduce (kernels.cuh), one per process, mapped into every peer with
// CUDA IPC.  Optional: without it (single rank, IPC unavailable, D too large) the all-reduce is ncclAllReduce.
struct ProcessP2P {
  bool ready = false;
  void* local = nullptr;
  clic::P2PView view{};
  unsigned long long seq = 0;  // all-reduce sequence number; every rank issues the same sequence of all-reduces
  int* d_err = nullptr;
} g_p2p;

// Device buffer from the stream-ordered memory pool (cudaMallocAsync): an estimator is built per solve like the
// reference's (trajectory_manager.cpp:637), so allocations must be cheap; the pool keeps freed blocks cached.
thread_local cudaStream_t g_alloc_stream = nullptr;
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  bool owned = true;  // false: a view into another DevBuf (carve), never freed here
  ~DevBuf() { if (p && owned) cudaFreeAsync(p, st); }
  void view(void* ptr, size_t n) {
    if (p && owned) cudaFreeAsync(p, st);
    p = ptr;
    bytes = n;
    owned = false;
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    if (p && owned) cudaFreeAsync(p, st);
    owned = true;
    p = nullptr;
    bytes = 0;
    st = g_alloc_stream;
    cudaError_t e = cudaMallocAsync(&p, n, st);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
// One stream-ordered allocation for a group of buffers (a factor batch has ~10 arrays, its execution plan 5 more: at a
// few microseconds per cudaMallocAsync the allocations dominated the host time of adding a production-size batch).
struct CarveItem {
  DevBuf* buf;
  size_t bytes;
};
cudaError_t carve(DevBuf& arena, std::initializer_list<CarveItem> items) {
  size_t total = 0;
  for (const auto& it : items) total += (std::max<size_t>(it.bytes, 1) + 255) & ~(size_t)255;
  cudaError_t e = arena.ensure(total);
  if (e != cudaSuccess) return e;
  size_t off = 0;
  for (const auto& it : items) {
    it.buf->view(static_cast<char*>(arena.p) + off, it.bytes);
    off += (std::max<size_t>(it.bytes, 1) + 255) & ~(size_t)255;
  }
  return cudaSuccess;
}
struct AllocScope {  // every ABI entry point that may allocate names the estimator's stream
  cudaStream_t prev;
  explicit AllocScope(cudaStream_t s) : prev(g_alloc_stream) { g_alloc_stream = s; }
  ~AllocScope() { g_alloc_stream = prev; }
};

// CLIC_TRACE=1: host time of every traced ABI call (stderr), for chasing host-side stalls.
struct ScopedTrace {
  const char* name;
  std::chrono::steady_clock::time_point t0;
  static bool on() { static const bool v = getenv("CLIC_TRACE") != nullptr; return v; }
  explicit ScopedTrace(const char* n) : name(n) { if (on()) t0 = std::chrono::steady_clock::now(); }
  void lap(const char* what) {
    if (!on()) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[clic trace] %s/%s %.3f ms\n", name, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
  ~ScopedTrace() { lap("end"); }
};

// Kernel launch with the programmatic-dependent-launch attribute (kernels.cuh: pdl_wait / pdl_trigger).  CLIC_NO_PDL=1
// launches plainly (A/B switch); the device-side hooks are no-ops then.
bool pdl_enabled() { static const bool v = getenv("CLIC_NO_PDL") == nullptr; return v; }
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Cooperative launch (every CTA resident: pass_kernel synchronises the grid) with the PDL attribute on top when the
// driver accepts the pair; probed on the first launch.  CLIC_FUSED_COOP=0 drops the cooperative attribute (experiments
// only: residency is then merely likely), CLIC_FUSED_NO_PDL=1 the other one.
template <class... KArgs, class... Args>
cudaError_t launch_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static int pdl_ok = -1;
  static const bool coop = !(getenv("CLIC_FUSED_COOP") && getenv("CLIC_FUSED_COOP")[0] == '0');
  static const bool no_pdl = getenv("CLIC_FUSED_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (coop) {
    at[n].id = cudaLaunchAttributeCooperative;
    at[n].val.cooperative = 1;
    n++;
  }
  const bool try_pdl = pdl_enabled() && !no_pdl && pdl_ok != 0;
  if (try_pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
  if (try_pdl && pdl_ok < 0) {
    if (e == cudaSuccess) {
      pdl_ok = 1;
    } else {  // the pair is not accepted: cooperative only from now on
      cudaGetLastError();
      pdl_ok = 0;
      cfg.numAttrs = n - 1;
      e = cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
    }
  }
  return e;
}

// Launch geometry + record buffers of one factor stream.
struct StreamExec {
  int64_t n = 0;
  int grid = 0, slabs_per_warp = 0, n_warps = 0, warps_per_cta = kWarpsPerCta;
  int NT = 4, rd = 640, n_keys = 0;
  DevBuf payload, key, cost, overflow;
  DevBuf fin;  // solve path: per-key final blocks of reduce_all_kernel
  DevBuf arena;  // one allocation behind the five buffers above (plan_stream)
  // The adders upload on the estimator's copy stream; `ready` marks the batch complete there and the compute stream
  // waits for it once, right before the first kernel that reads the batch (H2D of later batches overlaps compute).
  cudaEvent_t ready = nullptr;
  bool ready_waited = true;
  long seq = 0;  // upload order (mark_ready)
  // slabs at which the knot interval of a time-sorted batch changes (host estimate at add time; a load-balance hint for
  // the fused pass: a CTA that contains such a slab has more records to finish, see build_fused)
  std::vector<int32_t> key_change_slab;
  StreamExec() = default;
  StreamExec(const StreamExec&) = delete;
  StreamExec& operator=(const StreamExec&) = delete;
  ~StreamExec() { if (ready) cudaEventDestroy(ready); }
  int* overflow_flag = nullptr;
  RecordBuf rb() const {
    RecordBuf r;
    r.payload = payload.as<double>();
    r.key = key.as<int>();
    r.cost = cost.as<double>();
    r.overflow = overflow.as<double>();
    r.overflow_flag = overflow_flag;
    return r;
  }
};

struct LoamBatch {
  LoamStream st;
  DevBuf t_ns, p[3], g[6], type;
  DevBuf raw[6];  // the arrays as uploaded (kept until destroy, see the adders)
  DevBuf arena;   // one allocation behind t_ns / p / g / type
  DevBuf d_gather;
  std::vector<int32_t> gather;  // key-aligned storage (build_key_gather): stored position -> input record, -1 = padding
  int64_t n_input = 0;          // records the caller added (st.n counts the padding too)
  int marg = 0;
  StreamExec ex;
};
struct ImuBatch {
  ImuStream st;
  DevBuf t_ns, gy[3], ac[3];
  DevBuf raw[3];
  DevBuf arena;
  DevBuf d_gather;
  std::vector<int32_t> gather;
  int64_t n_input = 0;
  int marg = 0;
  StreamExec ex;
};
struct ImageBatch {
  ImageStream st;
  DevBuf ti, tj, pi[3], pj[2], lm;
  DevBuf raw[4];
  DevBuf id_i, id_j, pack_t;  // pose packs (st.n_packs > 0)
  DevBuf lm_start, lm_obs;      // landmark -> its observations' stored positions (CSR), for rho_side_kernel
  std::vector<int32_t> h_lm_start, h_lm_obs;  // (host copies: the source of the asynchronous upload lives with the batch)
  std::vector<int32_t> gather;  // stored position -> add order, -1 = padding (buckets start on slab boundaries)
  int64_t n_input = 0;          // factors the caller added (st.n counts the padding too)
  std::vector<char> lm_used;  // landmarks referenced by a kept factor (marg batches: their inverse depths are dropped blocks)
  int marg = 0;
  StreamExec ex;
};
struct BiasFac {
  int slot_i, slot_j;
  double sqrt_info[6];
  int marg = 0, marg_all = 0;
};

}  // namespace

struct clic_prior {
  int n = 0, m = 0;
  std::vector<double> J0, r0;
  struct Block {
    int kind, id, size, idx;
    double x0[4];
  };
  std::vector<Block> blocks;
};

struct clic_estimator {
  // window
  int64_t t0_ns, dt_ns;
  int K, knot_id0, n_bias, bias_id0, n_lm;
  double S_CtoI[4], p_CinI[3];
  clic_options opt;
  int fixed_index_global = -1;
  StateLayout sl;
  std::vector<double> h_state;  // host mirror used only for the add-time window test (needs t_offset)
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of measurement batches + their prep kernels (process-wide, per device)
  bool own_stream = false;
  int device = 0, n_sm = 148;
  DevBuf d_state, d_cand, d_flags, d_counter;
  // normal equations: [H D*D | b D | cost_x 2 | cost_cand 2]; `loc` is what this rank's kernels write, `glb` the
  // all-reduced copy every consumer reads (the same buffer when there is one rank)
  DevBuf d_nrm_loc, d_nrm_glb;
  DevBuf d_nrm_loc2, d_nrm_glb2;  // same layout: the speculative Jacobian pass at the LM candidate (set 1)
  bool use_comm = false;
  double* H_loc() const { return d_nrm_loc.as<double>(); }
  double* b_loc() const { return d_nrm_loc.as<double>() + (size_t)std::max(1, L.D) * std::max(1, L.D); }
  double* cost_loc() const { return b_loc() + std::max(1, L.D); }
  double* nrm_glb() const { return use_comm ? d_nrm_glb.as<double>() : d_nrm_loc.as<double>(); }
  double* H_glb() const { return nrm_glb(); }
  double* b_glb() const { return nrm_glb() + (size_t)std::max(1, L.D) * std::max(1, L.D); }
  double* cost_glb() const { return b_glb() + std::max(1, L.D); }
  // set 0: the system at x;  set 1: the system at the LM candidate (lm_solve, speculative evaluation)
  double* nrm_loc_set(int set) const { return set ? d_nrm_loc2.as<double>() : d_nrm_loc.as<double>(); }
  double* nrm_glb_set(int set) const {
    return use_comm ? (set ? d_nrm_glb2.as<double>() : d_nrm_glb.as<double>()) : nrm_loc_set(set);
  }
  size_t off_b() const { return (size_t)std::max(1, L.D) * std::max(1, L.D); }
  size_t off_cost() const { return off_b() + std::max(1, L.D); }
  DevBuf d_knot_col, d_col_knot, d_col_sub, d_descs;
  std::vector<StreamDesc> h_descs;  // host copy of d_descs
  DevBuf d_lm;  // LM workspace (solver.cuh)
  std::vector<std::unique_ptr<LoamBatch>> loam;
  std::vector<std::unique_ptr<ImuBatch>> imu;
  std::vector<VelFactorDesc> vel_factors;  // AddGlobalVelocityMeasurement (one per init solve in the reference)
  std::vector<SmallFactor> small_f;        // pose / local velocity / relative rotation / image depth factors (§8f-4)
  DevBuf d_small;
  bool small_dirty = true;
  std::vector<PreintData> small_pre;       // pre-integration data of the kSmallPreIntegration factors (SmallFactor::ext)
  DevBuf d_small_pre;
  std::vector<std::unique_ptr<ImageBatch>> image;
  std::vector<BiasFac> bias_f;
  struct GravFac { double g0[3], sqrt_info[3]; int marg; };
  std::vector<GravFac> grav_f;  // AddGravityFactor
  // GravityFactors of the solve problem: gravity is a constant block there, so their cost is a constant of the problem
  double grav_fixed_cost() const {
    double c = 0.0;
    for (auto& f : grav_f) {
      if (f.marg) continue;
      for (int i = 0; i < 3; i++) {
        const double r = f.sqrt_info[i] * (h_state[sl.off_grav() + i] - f.g0[i]);
        c += 0.5 * r * r;
      }
    }
    return c;
  }
  struct PriorRef {
    const clic_prior* p;
    std::vector<int> drop;
    DevBuf J0, r0, x0, meta, A;  // device copies (+ A = J0^T J0)
    std::vector<int> h_meta;
  };
  std::vector<std::unique_ptr<PriorRef>> priors;
  std::vector<char> lm_free;
  bool layout_dirty = true;
  clic_layout L;
  int n_desc = 0;
  DevBuf d_jobs;
  long add_seq = 0;
  int64_t factors_since_pass = 0;  // uploaded by the adders since the last pass (run_pass: schedule of the first pass)
  bool packs_fit = true;  // the packed visual kernel's shared memory fits this window
  int n_jobs = 0, n_part_ctas = 0, max_slots = 0;
  DevBuf d_bias_descs;
  DevBuf d_lm_col, d_rho_side;  // free inverse depths: landmark -> column, per-landmark side rows of H | b
  bool bias_desc_dirty = true;
  int64_t launches = 0;
  // fused pass (pass_fused.cuh): one persistent launch per pass; built by ensure_layout when the problem qualifies
  struct Fused {
    bool ok = false;
    FusedArgs fa{};
    int grid = 0;
    size_t smem = 0;
    int part_ctas[kFusedKinds] = {0, 0, 0, 0, 0, 0};
    LoamBatch* lb[3] = {nullptr, nullptr, nullptr};
    ImuBatch* ib = nullptr;
    ImageBatch* gb = nullptr;
    DevBuf d_bar, d_stamps, d_addend, d_hloc, d_descs_ovf, arena;
    std::vector<StreamDesc> h_descs_ovf;
    bool stamps_on = false;
  } fused;
  bool bounds_toff[3] = {false, false, false};
  double toff_lo[3], toff_hi[3];
  // profiling (clic_profile_*)
  bool prof_on = false;
  struct ProfSpan { int kind; cudaEvent_t a, b; };
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t new_event() {
    cudaEvent_t e;
    if (!ev_pool.empty()) { e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
  }
  void prof_begin(int kind) {
    if (!prof_on) return;
    ProfSpan s{kind, new_event(), new_event()};
    cudaEventRecord(s.a, stream);
    prof_spans.push_back(s);
  }
  void prof_end() {
    if (!prof_on) return;
    cudaEventRecord(prof_spans.back().b, stream);
  }

  int64_t minTimeNs() const { return t0_ns; }
  int64_t maxTimeNs() const { return t0_ns + (int64_t)(K - 4 + 1) * dt_ns - 1; }
  // trajectory.h:100-108 — double arithmetic exactly as the reference
  double minTime(int s) const { return 1e-9 * minTimeNs() - 1e-9 * h_state[sl.off_toff() + s]; }
  double maxTime(int s) const { return 1e-9 * maxTimeNs() - 1e-9 * h_state[sl.off_toff() + s]; }
};

namespace {

bool have_device() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

clic_layout compute_layout(const clic_estimator* E) {
  clic_layout L;
  int kf = E->fixed_index_global - E->knot_id0 + 1;
  kf = std::max(0, std::min(E->K, kf));
  if (E->opt.lock_traj) kf = E->K;
  L.first_free_knot = kf;
  L.n_free_knots = E->K - kf;
  int c = 6 * L.n_free_knots;
  L.col_bias_gyro = L.col_bias_accel = -1;
  if (!E->opt.lock_wb && E->n_bias > 0) { L.col_bias_gyro = c; c += 3 * E->n_bias; }
  if (!E->opt.lock_ab && E->n_bias > 0) { L.col_bias_accel = c; c += 3 * E->n_bias; }
  for (int s = 0; s < 3; s++) L.col_t_offset[s] = -1;
  if (!E->opt.lock_t_offset[CLIC_SENSOR_IMU]) L.col_t_offset[CLIC_SENSOR_IMU] = c++;
  if (!E->opt.lock_t_offset[CLIC_SENSOR_CAMERA]) L.col_t_offset[CLIC_SENSOR_CAMERA] = c++;
  L.col_inv_depth = -1;
  L.n_free_landmarks = 0;
  for (int i = 0; i < E->n_lm; i++) L.n_free_landmarks += E->lm_free[i] ? 1 : 0;
  if (L.n_free_landmarks > 0) { L.col_inv_depth = c; c += L.n_free_landmarks; }
  L.D = c;
  return L;
}

int plan_stream(clic_estimator* E, StreamExec& ex, int64_t n, int NT, int ctas_per_sm, bool pair_keys = false,
                int extra = 0, int warps_per_cta = kWarpsPerCta, int factors_per_slab = 32) {
  const int kWarpsPerCta = warps_per_cta;  // shadows the global default for this stream
  ex.warps_per_cta = warps_per_cta;
  ex.n = n;
  E->factors_since_pass += n;
  ex.NT = NT;
  ex.rd = rec_doubles(NT, extra);
  ex.n_keys = pair_keys ? (E->K - 3) * (E->K - 3) : E->K - 3;
  const int64_t slabs = (n + factors_per_slab - 1) / factors_per_slab;
  const int max_ctas = E->n_sm * ctas_per_sm;
  // a full grid whenever there is at least one slab per warp; the kernels split the slabs evenly over its warps
  // (the visual stream down to 4 slabs per CTA: see the partition model of the fused pass)
  const int64_t min_per_cta = pair_keys ? std::max(1, kWarpsPerCta / 2) : kWarpsPerCta;
  int64_t ctas = std::min<int64_t>(max_ctas, (slabs + min_per_cta - 1) / min_per_cta);
  ex.grid = (int)std::max<int64_t>(ctas, 1);
  ex.slabs_per_warp = (int)slabs;  // = total slabs (kernel argument)
  ex.n_warps = ex.grid * kWarpsPerCta;
  const size_t slots = (size_t)ex.n_warps * kSlotsPerWarp;
  CU(carve(ex.arena, {{&ex.payload, slots * ex.rd * sizeof(double)},
                      {&ex.key, slots * sizeof(int)},
                      {&ex.cost, (size_t)ex.n_warps * 2 * sizeof(double)},
                      {&ex.overflow, (size_t)ex.n_keys * ex.rd * sizeof(double)},
                      {&ex.fin, (size_t)ex.n_keys * ex.rd * sizeof(double)}}));
  // reduce_all_kernel folds the atomics-fallback block into fin and re-zeroes it
  CU(cudaMemsetAsync(ex.overflow.p, 0, (size_t)ex.n_keys * ex.rd * sizeof(double), g_alloc_stream));
  ex.overflow_flag = E->d_flags.as<int>();
  return CLIC_OK;
}

size_t factor_smem_bytes(int K, int NT) {
  // NT column tiles of the transposed Jacobian tile per warp; the end-of-kernel merge scratch aliases it
  size_t jt = (size_t)kWarpsPerCta * 8 * NT * kJtStride;
  jt = std::max<size_t>(jt, (size_t)kWarpsPerCta * rec_doubles(NT, 1));  // every warp parks its accumulator there
  return (window_smem_doubles(K) + jt) * sizeof(double);
}

size_t imu_smem_bytes(int K) { return factor_smem_bytes(K, 4); }

// ImuCompactXform's expansion table: once per process and device
int imu_xform_table(int device, const int4** out) {
  static std::mutex mu;
  static std::map<int, int4*> tabs;
  std::lock_guard<std::mutex> lock(mu);
  auto it = tabs.find(device);
  if (it == tabs.end()) {
    std::vector<int> h((size_t)kImuRecDoubles * 4);
    ImuCompactXform::build_host(h.data());
    int4* d = nullptr;
    CU(cudaMalloc(&d, h.size() * sizeof(int)));
    CU(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    it = tabs.emplace(device, d).first;
  }
  *out = it->second;
  return CLIC_OK;
}

size_t image_smem_bytes(int K, int n_packs) {
  return factor_smem_bytes(K, 7) + (size_t)n_packs * kPackDoubles * sizeof(double);
}

size_t loam_smem_bytes(int K) {
  size_t jt = (size_t)kLoamWarps * 24 * kJtStride;
  jt = std::max<size_t>(jt, (size_t)kLoamWarps * rec_doubles(3, 1));
  return (window_smem_doubles(K) + jt) * sizeof(double);
}

// End of an adder: the batch is complete on the copy stream when `ready` fires.
int mark_ready(clic_estimator* E, StreamExec& ex, cudaStream_t producer = nullptr) {
  if (!ex.ready) CU(cudaEventCreateWithFlags(&ex.ready, cudaEventDisableTiming));
  CU(cudaEventRecord(ex.ready, producer ? producer : E->copy_stream));
  ex.ready_waited = false;
  ex.seq = ++E->add_seq;
  return CLIC_OK;
}
int wait_ready(clic_estimator* E, StreamExec& ex) {
  if (ex.ready_waited) return CLIC_OK;
  CU(cudaStreamWaitEvent(E->stream, ex.ready, 0));
  ex.ready_waited = true;
  return CLIC_OK;
}
int join_adds(clic_estimator* E) {
  int rc;
  for (auto& b : E->loam) if ((rc = wait_ready(E, b->ex))) return rc;
  for (auto& b : E->imu) if ((rc = wait_ready(E, b->ex))) return rc;
  for (auto& b : E->image) if ((rc = wait_ready(E, b->ex))) return rc;
  return CLIC_OK;
}

// After a synchronisation: did a one-shot all-reduce give up waiting for a peer?
int check_p2p(clic_estimator* E) {
  if (!E->use_comm || !g_p2p.ready) return CLIC_OK;
  int err = 0;
  CU(cudaMemcpy(&err, g_p2p.d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) {
    CU(cudaMemset(g_p2p.d_err, 0, sizeof(int)));  // reported once; later all-reduces start clean
    return fail(CLIC_ERR_NCCL, "peer-memory all-reduce timed out waiting for a rank (result zeroed)");
  }
  return CLIC_OK;
}

// ---- fused pass (pass_fused.cuh) ----
// CLIC_FUSED=0: one kernel per stream + reduction + assembly (+ all-reduce) as in round 1; =1: the fused pass whenever
// it is applicable; unset: automatic (build_fused: the fused pass for windows with enough work to amortise its fixed
// cost — per-CTA dense system, grid barrier — and always on several ranks, where it carries the exchange).  Read at
// every layout so that a test can flip it between two estimators.
int fused_mode() { const char* e = getenv("CLIC_FUSED"); return !e ? -1 : (e[0] == '0' ? 0 : 1); }
// Cost model of the grid partition: partition_model.hpp.  CLIC_FUSED_W / CLIC_FUSED_F = "mixed,planes,lines,imu,visual
// packed,visual chain" and CLIC_FUSED_LAT override its constants (tuning runs).
const partition::Model& fused_model() {
  static partition::Model m = partition::default_model();
  static bool init = false;
  if (!init) {
    init = true;
    double v[kFusedKinds];
    if (const char* e = getenv("CLIC_FUSED_W"))
      if (sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == kFusedKinds)
        for (int i = 0; i < kFusedKinds; i++) if (v[i] > 0) m.w[i] = v[i];
    if (const char* e = getenv("CLIC_FUSED_F"))
      if (sscanf(e, "%lf,%lf,%lf,%lf,%lf,%lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == kFusedKinds)
        for (int i = 0; i < kFusedKinds; i++) if (v[i] >= 0) m.f[i] = v[i];
    if (const char* e = getenv("CLIC_FUSED_LAT")) m.lat = atof(e);
  }
  return m;
}
unsigned long long p2p_timeout_ns() {
  static unsigned long long v = 0;
  if (!v) {
    double ms = 30000.0;
    if (const char* e = getenv("CLIC_P2P_TIMEOUT_MS")) ms = std::max(1.0, atof(e));
    v = (unsigned long long)(ms * 1e6);
  }
  return v;
}

// Decides whether the pass of this layout runs as ONE pass_kernel launch and prepares its constant arguments: at most one
// non-marginalization batch per stream kind (every shipped call pattern: a plane batch, a line batch, one IMU batch, one
// visual batch), a device that launches cooperatively, and the grid partition from the cost model above.
int build_fused(clic_estimator* E) {
  auto& F = E->fused;
  F.ok = false;
  for (int k = 0; k < 3; k++) F.lb[k] = nullptr;
  F.ib = nullptr;
  F.gb = nullptr;
  if (fused_mode() == 0 || E->L.D <= 0) return CLIC_OK;
  for (auto& b : E->loam) {
    if (b->marg) continue;
    const int k = b->st.geo_kind;
    if (k < 0 || k > 2 || F.lb[k]) return CLIC_OK;
    F.lb[k] = b.get();
  }
  for (auto& b : E->imu) {
    if (b->marg) continue;
    if (F.ib) return CLIC_OK;
    F.ib = b.get();
  }
  for (auto& b : E->image) {
    if (b->marg) continue;
    if (F.gb || b->st.n_packs <= 0) return CLIC_OK;  // the per-factor chain variant (> 96 distinct frame times) stays unfused
    F.gb = b.get();
  }
  int coop = 0;
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, E->device));
  if (!coop) return CLIC_OK;
  struct Job { int kind; StreamExec* ex; int n; };
  std::vector<Job> js;
  std::vector<partition::Stream> streams;
  const partition::Model& M = fused_model();
  auto push = [&](int kind, StreamExec& ex, double scale = 1.0) {
    partition::Stream s;
    s.kind = kind;
    s.slabs = std::max<int64_t>(1, ex.slabs_per_warp);  // (the field holds the stream's total slab count)
    // at most one slab per warp; the visual stream may spread thinner (down to 4 slabs per CTA): its dense finish costs
    // one scatter per record, i.e. per slab when every slab is its own interval pair
    const int64_t min_per_cta = kind >= 4 ? kWarpsPerCta / 2 : kWarpsPerCta;
    s.cap = (int)std::max<int64_t>(1, std::min<int64_t>(ex.grid, (s.slabs + min_per_cta - 1) / min_per_cta));
    s.w_scale = scale;
    streams.push_back(s);
    js.push_back(Job{kind, &ex, 1});
  };
  // same order as the descriptors / reduce jobs of ensure_layout: LiDAR batches in add order, IMU, visual
  for (auto& b : E->loam) if (!b->marg) push(b->st.geo_kind, b->ex);
  if (F.ib) push(3, F.ib->ex, E->L.col_t_offset[CLIC_SENSOR_IMU] >= 0 ? 1.25 : 1.0);
  if (F.gb) push(F.gb->st.n_packs > 0 ? 4 : 5, F.gb->ex);
  if (js.empty() || (int)js.size() > E->n_sm || E->n_sm > kMaxFusedCtas) return CLIC_OK;
  if (fused_mode() < 0 && !E->use_comm) {
    // Small windows (the reference's production size: a few hundred factors, ~30 slabs) are latency-bound and run
    // faster as per-stream launches (measured, bench_micro/small_window.py: C1 85 vs 100 us per LM iteration); from
    // ~C2's size on the fused pass wins (C3 45.6 vs 54.4 us per pass, C5 204 vs 244 us).  Work = modelled warp-time.
    if (partition::work(M, streams) < 6000.0) return CLIC_OK;
  }
  {
    const std::vector<int> n = partition::split_grid(M, streams, E->n_sm);
    for (size_t i = 0; i < js.size(); i++) js[i].n = n[i];
  }
  if ((int)js.size() > kMaxFusedJobs || (int)js.size() != E->n_desc) return CLIC_OK;
  FusedArgs& fa = F.fa;
  std::memset(&fa, 0, sizeof(fa));
  if (!F.d_bar.p) {  // one barrier word per CTA + the overflow flag (word 1024)
    CU(F.d_bar.ensure(1025 * sizeof(unsigned)));
    CU(cudaMemsetAsync(F.d_bar.p, 0, 1025 * sizeof(unsigned), E->stream));
  }
  const int D = E->L.D;
  const int n_tot = D * (D + 1) / 2 + D + 2;
  int cta0 = 0;
  size_t body_smem = 0;
  for (int k = 0; k < kFusedKinds; k++) F.part_ctas[k] = 0;
  int jn = 0;
  for (auto& j : js) {
    FusedPart& P = fa.part[j.kind];
    P.cta0 = cta0;
    P.n_ctas = j.n;
    P.total_slabs = j.ex->slabs_per_warp;
    P.desc = jn;
    P.rb = j.ex->rb();
    P.rb.overflow_flag = reinterpret_cast<int*>(F.d_bar.as<unsigned>() + 1024);  // read back by phase A of the same pass
    F.part_ctas[j.kind] = j.n;
    {  // slab range of every CTA of the partition (partition_model.hpp: boundary-aware)
      const std::vector<int64_t> end =
          partition::cta_slab_ends(j.kind, j.ex->slabs_per_warp, j.n, j.ex->key_change_slab);
      for (int c = 0; c < j.n; c++) fa.cta_slab_end[cta0 + c] = (int)end[c];
    }
    cta0 += j.n;
    fa.ovf_blocks[jn] = j.ex->overflow.as<double>();
    fa.ovf_doubles[jn] = j.ex->n_keys * j.ex->rd;
    jn++;
    if (j.kind <= 2) body_smem = std::max(body_smem, loam_smem_bytes(E->K));
    else if (j.kind == 3) body_smem = std::max(body_smem, imu_smem_bytes(E->K));
    else body_smem = std::max(body_smem, image_smem_bytes(E->K, F.gb->st.n_packs));
  }
  fa.n_jobs = jn;
  fa.n_factor_ctas = cta0;
  for (int k = 0; k < 3; k++) if (F.lb[k]) fa.loam[k] = F.lb[k]->st;
  if (F.ib) fa.imu = F.ib->st;
  if (F.gb) fa.image = F.gb->st;
  F.grid = std::min(E->n_sm, std::max(cta0, (n_tot + 15) / 16));
  const int epc = (n_tot + F.grid - 1) / F.grid;
  // phase F: the stream bodies' shared memory, then the CTA's packed system and the key-column scratch;
  // phase A: staged chunk + coordinates + the (producers x chunk) tile
  body_smem = (body_smem + 15) & ~(size_t)15;
  fa.hs_off = (int)(body_smem / sizeof(double));
  size_t smem = body_smem + (size_t)((n_tot + 1) & ~1) * sizeof(double) + kKeyColsInts * sizeof(int);
  smem = std::max(smem, (size_t)2 * epc * sizeof(double) + (size_t)cta0 * epc * sizeof(double));
  smem = (smem + 15) & ~(size_t)15;
  fa.persist_off = (int)(smem / sizeof(double));
  smem += fused_persist_bytes(E->K, E->n_bias);
  int max_optin = 0;
  CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, E->device));
  if (smem > (size_t)max_optin) return CLIC_OK;  // a system too wide for a per-CTA copy (C4, D = 240): per-stream kernels
  F.smem = smem;
  {  // an estimator is built per solve: the attribute / occupancy queries are cached per device and size
    static std::mutex mu;
    static size_t set_smem[16] = {};
    static int occ_of[16] = {};
    std::lock_guard<std::mutex> lock(mu);
    const int dv = E->device & 15;
    if (smem > set_smem[dv]) {
      CU(cudaFuncSetAttribute(pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU(cudaFuncSetAttribute(pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_smem[dv] = smem;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_of[dv], pass_kernel<true>, kThreads, smem));
    }
    if (occ_of[dv] < 1 || F.grid > occ_of[dv] * E->n_sm) return CLIC_OK;  // the grid barrier needs every CTA resident
  }
  // one allocation behind the kernel's scratch buffers
  CU(carve(F.arena, {{&F.d_stamps, (size_t)E->n_sm * 8 * sizeof(unsigned long long)},
                     {&F.d_addend, ((size_t)D * D + D + 2) * sizeof(double)},
                     {&F.d_hloc, (size_t)cta0 * n_tot * sizeof(double)},
                     {&F.d_descs_ovf, (size_t)kMaxFusedJobs * sizeof(StreamDesc)}}));
  {  // descriptors of the atomics-fallback blocks: the per-interval block layout of the per-stream path
    F.h_descs_ovf = E->h_descs;  // (member: the asynchronous copy below reads it after this function returns)
    for (int j = 0; j < jn; j++) F.h_descs_ovf[j].fin = fa.ovf_blocks[j];
    CU(cudaMemcpyAsync(F.d_descs_ovf.p, F.h_descs_ovf.data(), F.h_descs_ovf.size() * sizeof(StreamDesc),
                       cudaMemcpyHostToDevice, E->stream));
  }
  fa.hloc = F.d_hloc.as<double>();
  fa.cm.D = D;
  fa.cm.K = E->K;
  fa.cm.knot_col = E->d_knot_col.as<int>();
  fa.cm.col_knot = E->d_col_knot.as<int>();
  fa.cm.col_sub = E->d_col_sub.as<int>();
  fa.descs = E->d_descs.as<StreamDesc>();
  fa.descs_ovf = F.d_descs_ovf.as<StreamDesc>();
  fa.off_bias = E->sl.off_bias();
  fa.bar = F.d_bar.as<unsigned>();
  fa.ovf_flag = F.d_bar.as<unsigned>() + 1024;
  F.ok = true;
  return CLIC_OK;
}

// Column maps + stream descriptors for assemble_kernel.
int ensure_layout(clic_estimator* E) {
  if (!E->layout_dirty) return CLIC_OK;
  E->L = compute_layout(E);
  const clic_layout& L = E->L;
  const int D = std::max(1, L.D), K = E->K;
  std::vector<int> knot_col(K), col_knot(D, -1), col_sub(D, 0);
  for (int k = 0; k < K; k++) {
    knot_col[k] = k >= L.first_free_knot ? 6 * (k - L.first_free_knot) : -1;
    if (knot_col[k] >= 0)
      for (int c = 0; c < 6; c++) {
        col_knot[knot_col[k] + c] = k;
        col_sub[knot_col[k] + c] = c;
      }
  }
  CU(E->d_knot_col.ensure(K * sizeof(int)));
  CU(E->d_col_knot.ensure(D * sizeof(int)));
  CU(E->d_col_sub.ensure(D * sizeof(int)));
  CU(cudaMemcpyAsync(E->d_knot_col.p, knot_col.data(), K * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  CU(cudaMemcpyAsync(E->d_col_knot.p, col_knot.data(), D * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  CU(cudaMemcpyAsync(E->d_col_sub.p, col_sub.data(), D * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  std::vector<StreamDesc> descs;
  std::vector<ReduceJob> jobs;
  int n_part_ctas = 0;
  auto add_job = [&](StreamExec& ex) {
    ReduceJob j{};
    j.payload = ex.payload.as<double>();
    j.key = ex.key.as<int>();
    j.fin = ex.fin.as<double>();
    j.overflow = ex.overflow.as<double>();
    j.warp_cost = ex.cost.as<double>();
    j.n_slots = ex.n_warps * kSlotsPerWarp;
    j.n_keys = ex.n_keys; j.rd = ex.rd; j.n_warps = ex.n_warps;
    j.cta0 = n_part_ctas;
    n_part_ctas += ex.n_keys;
    jobs.push_back(j);
  };
  for (auto& b : E->loam) {
    if (b->marg) continue;
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    add_job(b->ex);
    d.r_col = -1;
    d.b_off = 384;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)l;
    d.n_extra = 0;
    d.extra_base = 24;
    d.kind = 0;
    d.kdim = E->K - 3;
    descs.push_back(d);
  }
  for (auto& b : E->imu) {
    if (b->marg) continue;
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    add_job(b->ex);
    d.r_col = 31;
    d.b_off = -1;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)(l < 32 ? imu_phys_col<false>(l) : l);
    d.n_extra = 7;
    d.extra_base = 24;
    d.kind = 0;
    d.kdim = E->K - 3;
    for (int c = 0; c < 3; c++) {
      d.extra_global[c] = L.col_bias_gyro >= 0 ? L.col_bias_gyro + 3 * b->st.bias_slot + c : -1;
      d.extra_global[3 + c] = L.col_bias_accel >= 0 ? L.col_bias_accel + 3 * b->st.bias_slot + c : -1;
    }
    d.extra_global[6] = L.col_t_offset[CLIC_SENSOR_IMU];
    descs.push_back(d);
  }
  for (auto& b : E->image) {
    if (b->marg) continue;
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    add_job(b->ex);
    d.r_col = 49;
    d.b_off = -1;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)l;
    d.n_extra = 1;
    d.extra_base = 48;
    d.kind = 1;
    d.kdim = E->K - 3;
    d.extra_global[0] = L.col_t_offset[CLIC_SENSOR_CAMERA];
    descs.push_back(d);
  }
  E->n_desc = (int)descs.size();
  E->h_descs = descs;
  CU(E->d_descs.ensure(std::max<size_t>(1, descs.size()) * sizeof(StreamDesc)));
  if (!descs.empty())
    CU(cudaMemcpyAsync(E->d_descs.p, descs.data(), descs.size() * sizeof(StreamDesc), cudaMemcpyHostToDevice, E->stream));
  if (L.n_free_landmarks > 0) {
    std::vector<int32_t> lm_col(E->n_lm, -1);
    int rank = 0;
    for (int i = 0; i < E->n_lm; i++)
      if (E->lm_free[i]) lm_col[i] = L.col_inv_depth + rank++;
    CU(E->d_lm_col.ensure((size_t)E->n_lm * sizeof(int32_t)));
    CU(cudaMemcpyAsync(E->d_lm_col.p, lm_col.data(), lm_col.size() * sizeof(int32_t), cudaMemcpyHostToDevice, E->stream));
    CU(E->d_rho_side.ensure((size_t)L.n_free_landmarks * (L.D + 2) * sizeof(double)));
    CU(cudaStreamSynchronize(E->stream));  // lm_col is a local
  }
  E->n_jobs = (int)jobs.size();
  E->max_slots = 1;
  for (auto& j : jobs) E->max_slots = std::max(E->max_slots, j.n_slots);
  if ((size_t)2 * E->max_slots * sizeof(int) > 32 * 1024)  // static shared memory comes on top of it
    CU(cudaFuncSetAttribute(reduce_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * E->max_slots * sizeof(int))));
  E->n_part_ctas = n_part_ctas;
  CU(E->d_jobs.ensure(std::max<size_t>(1, jobs.size()) * sizeof(ReduceJob)));
  if (!jobs.empty())
    CU(cudaMemcpyAsync(E->d_jobs.p, jobs.data(), jobs.size() * sizeof(ReduceJob), cudaMemcpyHostToDevice, E->stream));
  CU(E->d_nrm_loc.ensure(((size_t)D * D + D + 8) * sizeof(double)));
  CU(E->d_nrm_loc2.ensure(((size_t)D * D + D + 8) * sizeof(double)));
  if (E->use_comm) {
    CU(E->d_nrm_glb.ensure(((size_t)D * D + D + 8) * sizeof(double)));
    CU(E->d_nrm_glb2.ensure(((size_t)D * D + D + 8) * sizeof(double)));
  }
  CU(cudaStreamSynchronize(E->stream));  // host vectors above go out of scope
  E->layout_dirty = false;
  E->bias_desc_dirty = true;
  return build_fused(E);
}


int launch_prior(clic_estimator* E, clic_estimator::PriorRef* pr, const double* state, double* cost2, bool jac,
                 const int* ctl, double* Hp, double* bp) {
  const int n = pr->p->n, nb = (int)pr->p->blocks.size();
  const int D = std::max(1, E->L.D);
  size_t smem = (size_t)3 * n * sizeof(double) + (size_t)n * sizeof(int);
  CU(launch_pdl(prior_kernel, 1, 1024, smem, E->stream, pr->J0.as<double>(), pr->A.as<double>(), pr->r0.as<double>(),
                pr->x0.as<double>(), pr->meta.as<int>(), n, nb, state, Hp, bp, D, cost2, jac ? 1 : 0,
                ctl));
  E->launches++;
  return CLIC_OK;
}

// column of a prior block under the current layout (same rule as every other block)
int fill_prior_columns(clic_estimator* E) {
  const clic_layout& L = E->L;
  for (auto& pr : E->priors) {
    const int nb = (int)pr->p->blocks.size();
    for (int i = 0; i < nb; i++) {
      const auto& b = pr->p->blocks[i];
      int col = -1;
      switch (b.kind) {
        case CLIC_BLOCK_KNOT_ROT:
        case CLIC_BLOCK_KNOT_POS: {
          int k = b.id - E->knot_id0;
          if (k >= L.first_free_knot) col = 6 * (k - L.first_free_knot) + (b.kind == CLIC_BLOCK_KNOT_POS ? 3 : 0);
        } break;
        case CLIC_BLOCK_BIAS_GYRO:
          if (L.col_bias_gyro >= 0) col = L.col_bias_gyro + 3 * (b.id - E->bias_id0);
          break;
        case CLIC_BLOCK_BIAS_ACCEL:
          if (L.col_bias_accel >= 0) col = L.col_bias_accel + 3 * (b.id - E->bias_id0);
          break;
        case CLIC_BLOCK_T_OFFSET: col = L.col_t_offset[b.id]; break;
        default: col = -1;  // gravity: locked; inverse depth: fixed_depth only in this revision
      }
      pr->h_meta[4 * i + 3] = col;
    }
    CU(cudaMemcpyAsync(pr->meta.p, pr->h_meta.data(), pr->h_meta.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
  }
  if (!E->priors.empty()) CU(cudaStreamSynchronize(E->stream));
  return CLIC_OK;
}

PassParams pass_params(const clic_estimator* E, const double* state, const int* ctl) {
  PassParams pp;
  pp.state = state;
  pp.sl = E->sl;
  pp.t0_ns = E->t0_ns;
  pp.dt_ns = E->dt_ns;
  pp.kf = E->L.first_free_knot;
  pp.marg_mode = 0;
  pp.marg_later_local = E->opt.ctrl_to_be_opt_later - E->knot_id0;
  pp.ctl = ctl;
  pp.pdl_first = 0;
  return pp;
}

std::vector<BiasFactorDesc> solve_bias_descs(const clic_estimator* E) {
  std::vector<BiasFactorDesc> bias_descs;
  for (auto& bf : E->bias_f) {
    if (bf.marg) continue;
    BiasFactorDesc d;
    d.slot_i = bf.slot_i;
    d.slot_j = bf.slot_j;
    for (int i = 0; i < 6; i++) d.sqrt_info[i] = bf.sqrt_info[i];
    d.col_gi = E->L.col_bias_gyro >= 0 ? E->L.col_bias_gyro + 3 * bf.slot_i : -1;
    d.col_gj = E->L.col_bias_gyro >= 0 ? E->L.col_bias_gyro + 3 * bf.slot_j : -1;
    d.col_ai = E->L.col_bias_accel >= 0 ? E->L.col_bias_accel + 3 * bf.slot_i : -1;
    d.col_aj = E->L.col_bias_accel >= 0 ? E->L.col_bias_accel + 3 * bf.slot_j : -1;
    bias_descs.push_back(d);
  }
  return bias_descs;
}

ImageShared image_shared(const clic_estimator* E, const ImageBatch* b) {
  ImageShared ish;
  ish.S_CtoI = ldq(E->S_CtoI);
  ish.p_CinI = ldv(E->p_CinI);
  ish.sqrt_info = E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450.0 / 1.5;
  ish.loss_a = b->st.loss_a;
  ish.inv_dt = 1e9 / double(E->dt_ns);
  ish.want_toff = E->L.col_t_offset[CLIC_SENSOR_CAMERA] >= 0;
  return ish;
}
void image_pass_fields(const clic_estimator* E, ImageStream& st, bool jac) {
  const bool free_rho = E->L.n_free_landmarks > 0;
  (void)jac;
  st.rho_side = nullptr;  // the landmark columns are gathered by rho_side_kernel, not scattered by the factor kernel
  st.lm_col = free_rho ? E->d_lm_col.as<int32_t>() : nullptr;
  st.knot_col = E->d_knot_col.as<int>();
  st.Dtot = E->L.D;
  st.rho0 = E->L.col_inv_depth;
  st.col_toff_cam = E->L.col_t_offset[CLIC_SENSOR_CAMERA];
}

// The inverse-depth columns of the solve system (free landmarks), one deterministic gather per visual batch into the
// zeroed side rows; runs before the assembly reads them.
int launch_rho_side(clic_estimator* E, const double* state, const int* ctl) {
  if (E->L.n_free_landmarks <= 0) return CLIC_OK;
  CU(cudaMemsetAsync(E->d_rho_side.p, 0, (size_t)E->L.n_free_landmarks * (E->L.D + 2) * sizeof(double), E->stream));
  PassParams pp = pass_params(E, state, ctl);
  int rc;
  for (auto& b : E->image) {
    if (b->marg) continue;
    if ((rc = wait_ready(E, b->ex))) return rc;
    ImageStream st = b->st;
    image_pass_fields(E, st, true);
    st.rho_side = E->d_rho_side.as<double>();
    const ImageShared ish = image_shared(E, b.get());
    const size_t smem = rho_side_smem_doubles(E->K, E->L.D) * sizeof(double);
    static size_t set_smem = 0;
    if (smem > set_smem) {
      CU(cudaFuncSetAttribute(rho_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_smem = smem;
    }
    const int grid = std::max(1, std::min(2 * E->n_sm, (E->n_lm + kRhoWarps - 1) / kRhoWarps));
    CU(launch_pdl(rho_side_kernel, grid, kRhoWarps * 32, smem, E->stream, st, pp, ish, (const int*)b->lm_start.as<int>(),
                  (const int*)b->lm_obs.as<int>(), E->n_lm));
    E->launches++;
  }
  return CLIC_OK;
}

// small_factors_kernel over E->small_f into (Hp, bp, cost2): after the assembly, or into the fused pass's addend system
int launch_small_factors(clic_estimator* E, const double* state, double* Hp, double* bp, double* cost2, bool jac,
                         const int* ctl) {
  if (E->small_f.empty()) return CLIC_OK;
  if (E->small_dirty) {
    CU(E->d_small.ensure(E->small_f.size() * sizeof(SmallFactor)));
    CU(cudaMemcpyAsync(E->d_small.p, E->small_f.data(), E->small_f.size() * sizeof(SmallFactor), cudaMemcpyHostToDevice,
                       E->stream));
    if (!E->small_pre.empty()) {
      CU(E->d_small_pre.ensure(E->small_pre.size() * sizeof(PreintData)));
      CU(cudaMemcpyAsync(E->d_small_pre.p, E->small_pre.data(), E->small_pre.size() * sizeof(PreintData),
                         cudaMemcpyHostToDevice, E->stream));
    }
    CU(cudaStreamSynchronize(E->stream));
    E->small_dirty = false;
  }
  PassParams pp = pass_params(E, state, ctl);
  SmallCols sc;
  sc.knot_col = E->d_knot_col.as<int>();
  sc.col_toff_imu = E->L.col_t_offset[CLIC_SENSOR_IMU];
  sc.lm_col = E->L.n_free_landmarks > 0 ? E->d_lm_col.as<int32_t>() : nullptr;
  sc.col_bias_gyro = E->L.col_bias_gyro;
  sc.col_bias_accel = E->L.col_bias_accel;
  const size_t smem = window_smem_doubles(E->K) * sizeof(double) + sizeof(SmallOut) + 16;
  CU(launch_pdl(small_factors_kernel, 1, 128, smem, E->stream, E->d_small.as<SmallFactor>(), (int)E->small_f.size(),
                (const PreintData*)E->d_small_pre.as<PreintData>(), pp, sc, Hp, bp, std::max(1, E->L.D), cost2, jac ? 1 : 0));
  E->launches++;
  return CLIC_OK;
}

// The pass as ONE persistent launch (pass_fused.cuh).  Terms that other kernels own (priors, the global-velocity
// factor, more BiasFactors than the kernel takes) are evaluated first into a zeroed addend system that phase A adds
// entry by entry, so they ride through the in-kernel exchange like everything else.
int run_pass_fused(clic_estimator* E, const double* state, bool jac, double* Hp, double* bp, double* cost2,
                   const int* ctl, int set) {
  auto& F = E->fused;
  int rc;
  const int D = E->L.D;
  for (int k = 0; k < 3; k++) if (F.lb[k] && (rc = wait_ready(E, F.lb[k]->ex))) return rc;
  if (F.ib && (rc = wait_ready(E, F.ib->ex))) return rc;
  if (F.gb && (rc = wait_ready(E, F.gb->ex))) return rc;
  PassParams pp = pass_params(E, state, ctl);
  pp.pdl_first = 1;
  FusedArgs fa = F.fa;
  if (F.ib) fa.imu.toff_free = E->L.col_t_offset[CLIC_SENSOR_IMU] >= 0 ? 1 : 0;
  if (F.gb) {
    image_pass_fields(E, fa.image, jac);
    fa.ish = image_shared(E, F.gb);
  }
  if (jac && E->L.n_free_landmarks > 0) {
    if ((rc = launch_rho_side(E, state, ctl))) return rc;
    fa.rho_side = E->d_rho_side.as<double>();
    fa.rho0 = E->L.col_inv_depth;
    fa.n_rho = E->L.n_free_landmarks;
  }
  std::vector<BiasFactorDesc> bias_descs = solve_bias_descs(E);
  const bool bias_in_kernel = !bias_descs.empty() && (int)bias_descs.size() <= kMaxFusedBias;
  if (bias_in_kernel) {  // the descriptors only depend on the layout: uploaded once per layout
    if (E->bias_desc_dirty) {
      CU(E->d_bias_descs.ensure(bias_descs.size() * sizeof(BiasFactorDesc)));
      CU(cudaMemcpyAsync(E->d_bias_descs.p, bias_descs.data(), bias_descs.size() * sizeof(BiasFactorDesc),
                         cudaMemcpyHostToDevice, E->stream));
      CU(cudaStreamSynchronize(E->stream));
      E->bias_desc_dirty = false;
    }
    fa.bias = E->d_bias_descs.as<BiasFactorDesc>();
    fa.n_bias = (int)bias_descs.size();
  }
  const bool have_priors = !E->priors.empty() && !E->opt.is_marg_state;
  const bool need_addend =
      have_priors || !E->vel_factors.empty() || !E->small_f.empty() || (!bias_descs.empty() && !bias_in_kernel);
  double* addend = nullptr;
  if (need_addend) {
    addend = F.d_addend.as<double>();
    double* Ha = addend;
    double* ba = addend + (size_t)D * D;
    double* ca = ba + D;
    CU(cudaMemsetAsync(addend, 0, ((size_t)D * D + D + 2) * sizeof(double), E->stream));
    E->prof_begin(4);
    if (!bias_in_kernel)
      for (auto& d : bias_descs) {
        CU(launch_pdl(bias_factor_kernel, 1, 32, 0, E->stream, d, state, E->sl, Ha, ba, D, ca, jac ? 1 : 0, ctl));
        E->launches++;
      }
    if (have_priors)
      for (auto& pr : E->priors)
        if ((rc = launch_prior(E, pr.get(), state, ca, jac, ctl, Ha, ba))) return rc;
    for (auto& vf : E->vel_factors) {
      CU(launch_pdl(velocity_factor_kernel, 1, 64, 0, E->stream, vf, state, E->sl, (long long)E->t0_ns, (long long)E->dt_ns,
                    E->K, (const int*)E->d_knot_col.as<int>(), E->L.col_t_offset[CLIC_SENSOR_IMU], Ha, ba, D, ca,
                    jac ? 1 : 0, ctl));
      E->launches++;
    }
    if ((rc = launch_small_factors(E, state, Ha, ba, ca, jac, ctl))) return rc;
    E->prof_end();
  }
  if (F.stamps_on) fa.stamps = F.d_stamps.as<unsigned long long>();
  FusedOut out{};
  out.H = Hp;
  out.b = bp;
  out.cost2 = cost2;
  out.addend = addend;
  const size_t n_tot = jac ? (size_t)D * (D + 1) / 2 + D + 2 : 2;
  // in-kernel exchange: two 8-byte words per double in the peer slots, one thread per value of a CTA's chunk
  const size_t chunk = (n_tot + F.grid - 1) / F.grid;
  const bool exchange = E->use_comm && g_p2p.ready && 2 * n_tot <= kP2PMaxDoubles && chunk <= (size_t)kThreads;
  if (exchange) {
    double* G = E->nrm_glb_set(set);
    out.Hg = G;
    out.bg = G + E->off_b();
    out.cost2g = G + (cost2 - Hp);
    out.seq = ++g_p2p.seq;
    out.timeout_ns = p2p_timeout_ns();
    out.err = g_p2p.d_err;
    out.exchange = 1;
  }
  E->prof_begin(0);
  if (jac) CU(launch_coop(pass_kernel<true>, F.grid, kThreads, F.smem, E->stream, fa, pp, out, g_p2p.view));
  else CU(launch_coop(pass_kernel<false>, F.grid, kThreads, F.smem, E->stream, fa, pp, out, g_p2p.view));
  E->prof_end();
  E->launches++;
  CU(cudaGetLastError());
  if (E->use_comm && !exchange) {  // no peer mapping (or a system too large for the slots): ncclAllReduce of the dense system
    NcclApi& nc = nccl_api();
    const size_t DD = (size_t)D * D + D;
    const double* ar_src = jac ? Hp : cost2;
    double* ar_dst = jac ? E->nrm_glb_set(set) : E->nrm_glb_set(set) + (cost2 - Hp);
    E->prof_begin(6);
    const int e = nc.AllReduce(ar_src, ar_dst, jac ? DD + 2 : 2, kNcclDouble, kNcclSum, g_comm.comm, E->stream);
    E->prof_end();
    if (e != 0) return fail(CLIC_ERR_NCCL, std::string("ncclAllReduce: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"));
  }
  return CLIC_OK;
}

template <class KernelT>
int set_smem(KernelT k, size_t bytes) {
  CU(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CLIC_OK;
}

// One fused residual (+ Jacobian + normal-equation) pass over the solve problem at `state`.
// jac = true: H, b, cost2 are produced; false: cost2 only.  ctl: optional device skip word.
// set = 1 (with jac): the whole system goes to the second buffer set (H_c | b_c | cost_c), cost in its slot 0.
constexpr int64_t kOverlapFactors = 250000;  // ~16 MB of factor records: 0.3 ms of PCIe to hide a kernel under
int run_pass(clic_estimator* E, const double* state, bool jac, int which, const int* ctl, int set = 0) {
  int rc = ensure_layout(E);
  if (rc) return rc;
  double* Hp = E->nrm_loc_set(set);
  double* bp = Hp + E->off_b();
  // slot 0: pass at x, slot 1: pass at the LM candidate (set 0);  set 1 keeps its cost in its own slot 0
  double* cost2 = Hp + E->off_cost() + (set ? 0 : 2 * which);
  // The first pass after a large upload (automatic mode, one GPU): the per-stream kernels below are launched in upload
  // order, run under the copies still in flight and leave only the last batch's kernel exposed, where the fused pass
  // cannot start before the last byte has arrived (C5 from pinned buffers: 2.03 instead of 2.15 ms per create + add +
  // evaluate, bench_micro/e2e_modes.py).  Decided from the call sequence, not from event state: the same calls take the
  // same path, so results stay bit-reproducible.  Ranks of a communicator always take the fused pass (one protocol).
  static const bool overlap_on = !(getenv("CLIC_OVERLAP_UPLOADS") && getenv("CLIC_OVERLAP_UPLOADS")[0] == '0');
  const bool overlap_uploads = overlap_on && fused_mode() < 0 && !E->use_comm && E->factors_since_pass >= kOverlapFactors;
  E->factors_since_pass = 0;
  if (E->fused.ok && !overlap_uploads) return run_pass_fused(E, state, jac, Hp, bp, cost2, ctl, set);
  PassParams pp = pass_params(E, state, ctl);
  const int D = E->L.D;
  if (jac && E->L.n_free_landmarks > 0)  // side rows of the inverse-depth columns: gathered per landmark
    if ((rc = launch_rho_side(E, state, ctl))) return rc;
  pp.pdl_first = 1;  // the first factor kernel of the pass; cleared after each launch
  auto launch_loam = [&](LoamBatch* b) -> int {
    StreamExec& ex = b->ex;
    if ((rc = wait_ready(E, ex))) return rc;
    const size_t smem = loam_smem_bytes(E->K);
    if (jac) {
      E->prof_begin(0);
      if (b->st.geo_kind == 1) CU(launch_pdl(loam_kernel<true, kLoamWarps, kLoamMinB, 1>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      else if (b->st.geo_kind == 2) CU(launch_pdl(loam_kernel<true, kLoamWarps, kLoamMinB, 2>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      else CU(launch_pdl(loam_kernel<true, kLoamWarps, kLoamMinB, 0>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      pp.pdl_first = 0;
      E->prof_end();
      E->launches += 1;
    } else {
      E->prof_begin(0);
      if (b->st.geo_kind == 1) CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 1>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      else if (b->st.geo_kind == 2) CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 2>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      else CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 0>, ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      pp.pdl_first = 0;
      E->prof_end();
      E->launches += 1;
    }
    return CLIC_OK;
  };
  auto launch_imu = [&](ImuBatch* b) -> int {
    StreamExec& ex = b->ex;
    if ((rc = wait_ready(E, ex))) return rc;
    const size_t smem = imu_smem_bytes(E->K);
    b->st.toff_free = E->L.col_t_offset[CLIC_SENSOR_IMU] >= 0 ? 1 : 0;
    if (jac) {
      E->prof_begin(1);
      CU(launch_pdl(imu_kernel<true, false>, ex.grid, kThreads, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      pp.pdl_first = 0;
      E->prof_end();
      E->launches += 1;
    } else {
      E->prof_begin(1);
      CU(launch_pdl(imu_kernel<false, false>, ex.grid, kThreads, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb()));
      pp.pdl_first = 0;
      E->prof_end();
      E->launches += 1;
    }
    return CLIC_OK;
  };
  auto launch_image = [&](ImageBatch* b) -> int {
    StreamExec& ex = b->ex;
    if ((rc = wait_ready(E, ex))) return rc;
    const size_t smem = factor_smem_bytes(E->K, 7);
    const ImageShared ish = image_shared(E, b);
    const bool packed = b->st.n_packs > 0;
    const size_t smem_p = image_smem_bytes(E->K, b->st.n_packs);
    image_pass_fields(E, b->st, jac);
    E->prof_begin(2);
    if (jac && packed) CU(launch_pdl(image_kernel<true, true>, ex.grid, kThreads, smem_p, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb(), ish));
    else if (jac) CU(launch_pdl(image_kernel<true, false>, ex.grid, kThreads, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb(), ish));
    else if (packed) CU(launch_pdl(image_kernel<false, true>, ex.grid, kThreads, smem_p, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb(), ish));
    else CU(launch_pdl(image_kernel<false, false>, ex.grid, kThreads, smem, E->stream, b->st, pp, ex.slabs_per_warp, ex.rb(), ish));
    E->prof_end();
    E->launches += 1;
    pp.pdl_first = 0;
    return CLIC_OK;
  };
  // Factor kernels are independent of each other: launch them in the order their batches were uploaded, so that a
  // kernel whose data has arrived runs under the uploads still in flight (per-batch ready events on the copy stream).
  struct Item { long seq; int kind; void* p; };
  std::vector<Item> order;
  for (auto& b : E->loam) if (!b->marg) order.push_back({b->ex.seq, 0, b.get()});
  for (auto& b : E->imu) if (!b->marg) order.push_back({b->ex.seq, 1, b.get()});
  for (auto& b : E->image) if (!b->marg) order.push_back({b->ex.seq, 2, b.get()});
  std::sort(order.begin(), order.end(), [](const Item& x, const Item& y) { return x.seq < y.seq; });
  for (auto& it : order) {
    if (it.kind == 0) rc = launch_loam(static_cast<LoamBatch*>(it.p));
    else if (it.kind == 1) rc = launch_imu(static_cast<ImuBatch*>(it.p));
    else rc = launch_image(static_cast<ImageBatch*>(it.p));
    if (rc) return rc;
  }
  if (E->n_jobs == 0) {  // no factor stream at all (prior-only problem)
    CU(launch_pdl(zero_cost_kernel, 1, 32, 0, E->stream, cost2, ctl));
  } else {  // every stream's records -> per-key blocks, and the pass cost, in one launch (cost only without jac)
    const int part_ctas = jac ? E->n_part_ctas : 0;
    E->prof_begin(3);
    CU(launch_pdl(reduce_all_kernel, part_ctas + 1, 256, (size_t)2 * E->max_slots * sizeof(int), E->stream,
                  E->d_jobs.as<ReduceJob>(), E->n_jobs, part_ctas, E->max_slots, cost2, ctl));
    E->prof_end();
  }
  E->launches += 1;
  std::vector<BiasFactorDesc> bias_descs = solve_bias_descs(E);
  const bool bias_in_assemble = jac && D > 0 && !bias_descs.empty() && bias_descs.size() <= 4;
  if (jac && D > 0) {
    ColMaps cm;
    cm.D = D;
    cm.K = E->K;
    cm.knot_col = E->d_knot_col.as<int>();
    cm.col_knot = E->d_col_knot.as<int>();
    cm.col_sub = E->d_col_sub.as<int>();
    const int n_entries = D * (D + 1) / 2 + D;
    const BiasFactorDesc* d_bias = nullptr;
    if (bias_in_assemble) {  // the descriptors only depend on the layout: uploaded once per layout
      if (E->bias_desc_dirty) {
        CU(E->d_bias_descs.ensure(bias_descs.size() * sizeof(BiasFactorDesc)));
        CU(cudaMemcpyAsync(E->d_bias_descs.p, bias_descs.data(), bias_descs.size() * sizeof(BiasFactorDesc),
                           cudaMemcpyHostToDevice, E->stream));
        CU(cudaStreamSynchronize(E->stream));
        E->bias_desc_dirty = false;
      }
      d_bias = E->d_bias_descs.as<BiasFactorDesc>();
    }
    E->prof_begin(3);
    CU(launch_pdl(assemble_kernel, (n_entries + 7) / 8, 256, 0, E->stream, Hp, bp, cm, E->d_descs.as<StreamDesc>(),
                  E->n_desc, ctl, d_bias, bias_in_assemble ? (int)bias_descs.size() : 0, state, E->sl.off_bias(),
                  cost2, E->L.n_free_landmarks > 0 ? (const double*)E->d_rho_side.as<double>() : (const double*)nullptr,
                  E->L.col_inv_depth, E->L.n_free_landmarks));
    E->prof_end();
    E->launches += 1;
  }
  if (!bias_in_assemble)
    for (auto& d : bias_descs) {
      CU(launch_pdl(bias_factor_kernel, 1, 32, 0, E->stream, d, state, E->sl, Hp, bp, std::max(1, D), cost2,
                    jac ? 1 : 0, ctl));
      E->launches += 1;
    }
  for (auto& pr : E->priors) {
    if (E->opt.is_marg_state) continue;
    if ((rc = launch_prior(E, pr.get(), state, cost2, jac, ctl, Hp, bp))) return rc;
  }
  for (auto& vf : E->vel_factors) {
    CU(launch_pdl(velocity_factor_kernel, 1, 64, 0, E->stream, vf, state, E->sl, (long long)E->t0_ns, (long long)E->dt_ns,
                  E->K, (const int*)E->d_knot_col.as<int>(), E->L.col_t_offset[CLIC_SENSOR_IMU], Hp, bp, std::max(1, D),
                  cost2, jac ? 1 : 0, ctl));
    E->launches++;
  }
  if ((rc = launch_small_factors(E, state, Hp, bp, cost2, jac, ctl))) return rc;
  CU(cudaGetLastError());
  if (E->use_comm) {
    // the single exchange step of the path: sum of H | b | cost over the ranks' factor shards.
    // Out of place (loc -> glb): a pass skipped by `ctl` leaves loc untouched and the sum is simply recomputed.
    E->prof_begin(6);
    NcclApi& nc = nccl_api();
    const size_t DD = (size_t)std::max(1, D) * std::max(1, D) + std::max(1, D);
    int e = 0;
    const double* ar_src = jac ? Hp : cost2;
    double* ar_dst = jac ? E->nrm_glb_set(set) : E->nrm_glb_set(set) + (cost2 - Hp);
    const size_t ar_count = jac ? DD + 2 : 2;
    if (g_p2p.ready && ar_count <= kP2PMaxDoubles) {
      CU(launch_pdl(p2p_allreduce_kernel, 1, 1024, 0, E->stream, ar_src, ar_dst, (int)ar_count, g_p2p.view, ++g_p2p.seq,
                    p2p_timeout_ns(), g_p2p.d_err));
      E->launches++;
    } else {
      e = nc.AllReduce(ar_src, ar_dst, ar_count, kNcclDouble, kNcclSum, g_comm.comm, E->stream);
    }
    E->prof_end();
    if (e != 0) return fail(CLIC_ERR_NCCL, std::string("ncclAllReduce: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"));
  }
  return CLIC_OK;
}

// copies a host array to a temporary device buffer on the estimator's stream
template <class T>
int to_device(clic_estimator* E, DevBuf& buf, const T* src, size_t count) {
  CU(buf.ensure(std::max<size_t>(1, count) * sizeof(T)));  // from g_alloc_stream's pool: the copy goes there too
  if (count) CU(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, g_alloc_stream));
  (void)E;
  return CLIC_OK;
}




// Key-aligned storage for small batches.  A warp evaluates 32 consecutive factors at a time and accumulates them into
// the 24-column block of ONE knot interval, so a 32-factor slab that straddles k intervals costs k evaluation rounds.
// With 10^5-10^6 factors per window that never matters; at the reference's production size (a few hundred LiDAR and
// ~100 IMU factors over ~10 intervals) nearly every slab straddles 2-4 intervals.  For a small batch the
// host therefore starts every interval's factors on a slab boundary (padding entries are stored as dropped factors):
// one round per slab, and the slabs of different intervals run on different warps.  Only the grouping of the sums
// changes (H, b move by rounding, ~1e-16).  Returns an empty map when the batch is large or already aligned.
// t_s: the caller's timestamps; shift_ns: what the factor adds before indexing ((int64) time offset for the IMU).
constexpr int64_t kKeyAlignMax = 1 << 15;
bool key_align_enabled() { static const bool v = getenv("CLIC_NO_KEY_ALIGN") == nullptr; return v; }
// Slab indices (factors_per_slab factors each) at which a time-SORTED stream enters the next knot interval: binary search
// per interval boundary on the caller's timestamps.  Only a hint (an unsorted stream gives meaningless, harmless values).
void key_change_slabs(const clic_estimator* E, const double* t_s, int64_t n, int64_t toff_ns, int factors_per_slab,
                      std::vector<int32_t>& out) {
  out.clear();
  if (!t_s || n <= 0) return;
  for (int k = 1; k <= E->K - 4; k++) {
    const int64_t tb = E->t0_ns + (int64_t)k * E->dt_ns;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
      const int64_t mid = (lo + hi) / 2;
      if ((int64_t)(t_s[mid] * 1e9) + toff_ns < tb) lo = mid + 1; else hi = mid;
    }
    if (lo > 0 && lo < n) out.push_back((int32_t)(lo / factors_per_slab));
  }
}

std::vector<int32_t> build_key_gather(const clic_estimator* E, const double* t_s, int64_t n, int64_t t_lo, int64_t t_hi,
                                      int64_t pad, int64_t shift_ns) {
  std::vector<int32_t> g;
  if (!key_align_enabled() || n > kKeyAlignMax || n <= 1) return g;
  std::vector<int32_t> key((size_t)n);
  const int n_keys = E->K - 3;
  std::vector<int32_t> count((size_t)n_keys, 0);
  int prev = -1;
  int64_t n_bad = 0;
  bool straddle = false;
  int64_t run = 0;  // valid records since the last slab boundary in the unpadded layout
  for (int64_t i = 0; i < n; i++) {
    const int64_t t = (int64_t)(t_s[i] * 1e9);  // prep_*_kernel's conversion and window test
    const bool ok = !(t - pad < t_lo || t + pad >= t_hi);
    int k = -1;
    if (ok) {
      const int64_t rel = t + shift_ns - E->t0_ns;
      k = rel < 0 ? -1 : (int)std::min<int64_t>(rel / E->dt_ns, n_keys - 1);
    }
    key[i] = k;
    if (k < 0) { n_bad++; continue; }
    // an unsorted batch is bucketed by interval all the same (a stable counting sort: arrival order within an
    // interval), which also takes it off the kernels' atomics fallback for records that revisit an interval
    if (k != prev && ((run % 32) != 0 || k < prev)) straddle = true;
    prev = k;
    count[k]++;
    run++;
  }
  if (!straddle && n_bad == 0) return g;
  std::vector<int64_t> start((size_t)n_keys + 1, 0);
  for (int k = 0; k < n_keys; k++) start[k + 1] = start[k] + (count[k] + 31) / 32 * 32;
  const int64_t total = start[n_keys] + n_bad;  // dropped records last: still counted by the prep kernel
  if (total > INT32_MAX) return g;
  g.assign((size_t)total, -1);
  std::vector<int64_t> fill(start.begin(), start.end() - 1);
  int64_t bad_at = start[n_keys];
  for (int64_t i = 0; i < n; i++) {
    if (key[i] >= 0) g[(size_t)fill[key[i]]++] = (int32_t)i;
    else g[(size_t)bad_at++] = (int32_t)i;
  }
  return g;
}

// clic_solve: the LM loop runs on the device; the host only enqueues kernels (and, for bounded problems, polls the
// line-search flag once per sample).
int lm_solve(clic_estimator* E, int max_iterations, clic_summary* summary) {
  const clic_layout& Lh = E->L;
  const int D = Lh.D;
  if (D <= 0) return fail(CLIC_ERR_BAD_ARGUMENT, "nothing to optimise (every parameter block is constant)");
  const size_t nst = E->sl.size();
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_scale = take(D * 8), o_diag = take(D * 8), o_step = take(D * 8), o_delta = take(D * 8),
               o_A = take((size_t)D * D * 8), o_tmp = take(nst * 8), o_act = take(D * 4), o_st = take(sizeof(LmState));
  CU(E->d_lm.ensure(off));
  char* base = E->d_lm.as<char>();
  LmBufs B;
  B.H = E->H_glb();
  B.b = E->b_glb();
  B.cost_x = E->cost_glb();
  B.cost_c = E->cost_glb() + 2;
  // Unconstrained problems evaluate the candidate WITH its Jacobians (second buffer set): an accepted step — the
  // common case — then needs no further pass; lm_accept_kernel adopts H_c, b_c, cost_c.  CLIC_LM_NO_SPECULATE=1 or a
  // bounded problem (line search on residual-only passes) keeps the residual pass + conditional Jacobian pass.
  B.Hc = nullptr;
  B.bc = nullptr;
  B.scale = (double*)(base + o_scale);
  B.diagonal = (double*)(base + o_diag);
  B.step = (double*)(base + o_step);
  B.delta = (double*)(base + o_delta);
  B.A = (double*)(base + o_A);
  B.tmp = (double*)(base + o_tmp);
  B.active = (int*)(base + o_act);
  B.x = E->d_state.as<double>();
  B.cand = E->d_cand.as<double>();
  B.st = (LmState*)(base + o_st);
  LayoutDev L;
  L.D = D; L.K = E->K; L.kf = Lh.first_free_knot; L.nb = E->n_bias; L.L = E->n_lm;
  L.lm_col = Lh.n_free_landmarks > 0 ? E->d_lm_col.as<int32_t>() : nullptr;
  L.col_bias_gyro = Lh.col_bias_gyro; L.col_bias_accel = Lh.col_bias_accel;
  bool constrained = false;
  for (int s = 0; s < 3; s++) {
    L.col_toff[s] = Lh.col_t_offset[s];
    L.has_bounds[s] = E->bounds_toff[s] ? 1 : 0;
    L.lo[s] = E->toff_lo[s];
    L.hi[s] = E->toff_hi[s];
    if (L.col_toff[s] >= 0 && E->bounds_toff[s]) constrained = true;
  }
  L.sl = E->sl;
  const bool speculate = !constrained && !getenv("CLIC_LM_NO_SPECULATE");
  if (speculate) {
    double* g2 = E->nrm_glb_set(1);
    B.Hc = g2;
    B.bc = g2 + E->off_b();
    B.cost_c = g2 + E->off_cost();
  }
  LmConst C;
  const int* ctl_res = &B.st->skip_res;
  const int* ctl_jac = &B.st->skip_jac;
  const size_t sm_small = 1024 * sizeof(double);
  size_t sm_step = (size_t)(std::max(512 + 3 * D + ((D + 7) / 8) * 64, 1024)) * sizeof(double);  // >= the accept / post scratch
  int a_in_smem = 0;
  if (sm_step + (size_t)(D + 1) * (D | 1) * sizeof(double) <= 200 * 1024) {
    a_in_smem = 1;
    sm_step += (size_t)(D + 1) * (D | 1) * sizeof(double);  // + one row: the right-hand side of the D <= 96 path
  } else {
    sm_step += (size_t)(D + kNB) * kPanStride * sizeof(double);  // staged panel (A stays in L2-resident global memory)
  }
  CU(cudaFuncSetAttribute(lm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_step));
  int rc;
  if (constrained) {
    CU(launch_pdl(lm_project_kernel, 1, 256, 0, E->stream, B, L));
    E->launches++;
  }
  if ((rc = run_pass(E, B.x, true, 0, nullptr))) return rc;
  E->prof_begin(5);
  CU(launch_pdl(lm_init_kernel, 1, 256, sm_small, E->stream, B, L, C, max_iterations, constrained ? 1 : 0));
  E->prof_end();
  E->launches++;
  const int max_ls = 20;  // Ceres default max_num_line_search_step_size_iterations
  // speculative Jacobians (the common case): accept + post of iteration it - 1 can run at the head of lm_step_kernel of
  // iteration it instead of as two launches of their own.  Measured (bench_micro/solve_time.py, CLIC_LM_TIMELINE): a
  // production-size window (per-stream kernels; C1) gains 8 % per LM iteration (86 -> 79 us); with the fused pass it
  // is neutral (C3) to 2 % slower (C5: the step part of the merged kernel takes 52 instead of 47 us), so those keep
  // the separate launches.  CLIC_LM_FUSE_ACCEPT=0/1 forces.
  static const int fuse_env = [] { const char* e = getenv("CLIC_LM_FUSE_ACCEPT"); return e ? atoi(e) : -1; }();
  const bool fuse_accept = speculate && !constrained && (fuse_env < 0 ? !E->fused.ok : fuse_env != 0);
  for (int it = 0; it <= max_iterations; it++) {
    E->prof_begin(5);
    CU(launch_pdl(lm_step_kernel, 1, 512, sm_step, E->stream, B, L, C, a_in_smem, constrained ? 1 : 0,
                  (fuse_accept && it > 0) ? (fuse_env == 2 ? 2 : 1) : 0));
    E->prof_end();
    E->launches++;
    if (it == max_iterations) break;  // the last launch only records the max-iterations termination
    if ((rc = run_pass(E, B.cand, speculate, 1, ctl_res, speculate ? 1 : 0))) return rc;
    if (constrained) {
      for (int ls = 0; ls < max_ls + 1; ls++) {
        CU(launch_pdl(lm_linesearch_kernel, 1, 256, 0, E->stream, B, L, max_ls));
        E->launches++;
        if ((rc = run_pass(E, B.cand, false, 1, ctl_res))) return rc;
        int active = 0;
        CU(cudaMemcpyAsync(&active, &B.st->ls_active, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
        CU(cudaStreamSynchronize(E->stream));
        if (!active) break;
      }
    }
    if (!fuse_accept) {
      E->prof_begin(5);
      CU(launch_pdl(lm_accept_kernel, 1, 1024, sm_small, E->stream, B, L, C));
      E->prof_end();
      E->launches++;
      if (!speculate && (rc = run_pass(E, B.x, true, 0, ctl_jac))) return rc;
      E->prof_begin(5);
      CU(launch_pdl(lm_post_kernel, 1, 256, sm_small, E->stream, B, L));
      E->prof_end();
      E->launches++;
    }
    // Long solves (Solve(50) of InitTrajWithPropagation) usually converge within a handful of iterations; every
    // iteration enqueued after that is ~9 launches that exit at once.  One 4-byte poll every 8 iterations stops the
    // enqueueing instead (Solve(8), the inner-iteration solve, never polls).
    if (max_iterations > 12 && (it % 8) == 7 && it + 1 < max_iterations) {
      int done = 0;
      CU(cudaMemcpyAsync(&done, &B.st->done, sizeof(int), cudaMemcpyDeviceToHost, E->stream));
      CU(cudaStreamSynchronize(E->stream));
      if (done) break;
    }
  }
  LmState st;
  CU(cudaMemcpyAsync(&st, B.st, sizeof(st), cudaMemcpyDeviceToHost, E->stream));
  int in_flags[2] = {0, 0};
  CU(cudaMemcpyAsync(in_flags, E->d_flags.p, sizeof(in_flags), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  CU(cudaGetLastError());
  if (in_flags[1]) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_loam: geometry array missing for a record type that is present");
  if (int prc = check_p2p(E)) return prc;
#ifdef CLIC_LM_TIMELINE
  if (getenv("CLIC_LM_TIMELINE")) {  // developer aid: entry / exit of the LM kernels on the device clock
    fprintf(stderr, "LM timeline (us since the first stamp; 1 init, 10 step entry, 11 after accept, 12 after post, 19 step exit, "
                    "20/21 accept kernel, 30/31 post kernel):\n");
    for (int i = 0; i < st.nts && i < 96; i++)
      fprintf(stderr, "  %2d %9.2f%s", st.tag[i], (double)(st.ts[i] - st.ts[0]) * 1e-3, (i % 6 == 5) ? "\n" : "");
    fprintf(stderr, "\n");
  }
#endif
  if (summary) {
    std::memset(summary, 0, sizeof(*summary));
    summary->iterations = st.iteration;
    summary->num_successful_steps = st.n_succ;
    summary->num_unsuccessful_steps = st.n_unsucc;
    summary->termination = st.termination;
    summary->termination_reason = st.reason;
    summary->num_jacobian_evals = st.n_jac;
    summary->num_residual_evals = st.n_res;
    const double gfc = E->grav_fixed_cost();  // GravityFactors on the constant gravity block
    summary->initial_cost = st.initial_cost + gfc;
    summary->final_cost = st.x_cost + st.fixed_cost + gfc;
    summary->fixed_cost = st.fixed_cost + gfc;
    summary->final_radius = st.radius;
    summary->final_gradient_max_norm = st.gmax;
  }
  return CLIC_OK;
}

// clic_marginalize: preMarginalize + marginalize (marginalization_factor.cpp:94-116, 150-276) on the device.
int marginalize_impl(clic_estimator* E, clic_prior** out) {
  ScopedTrace trace("marginalize");
  const int K = E->K;
  const int later_local = E->opt.ctrl_to_be_opt_later - E->knot_id0;
  const bool drop_knots = E->opt.ctrl_to_be_opt_later > E->opt.ctrl_to_be_opt_now;  // trajectory_estimator.cpp:214
  if (int jrc = join_adds(E)) return jrc;
  PassParams pp = pass_params(E, E->d_state.as<double>(), nullptr);
  pp.marg_mode = 1;
  pp.kf = 0;
  pp.marg_later_local = drop_knots ? later_local : 0;  // no knot is dropped => only factors with other dropped blocks stay
  int rc;
  DevBuf d_present;
  CU(d_present.ensure((size_t)(K - 3) * sizeof(int)));
  std::vector<char> knot_used(K, 0);
  std::vector<int> present(K - 3);
  DevBuf d_job;
  CU(d_job.ensure(sizeof(ReduceJob)));
  // records of one marg stream -> per-key blocks ex.fin (the same reduction kernel as the solve path, one job)
  auto reduce_stream = [&](StreamExec& ex) -> int {
    ReduceJob j{};
    j.payload = ex.payload.as<double>();
    j.key = ex.key.as<int>();
    j.fin = ex.fin.as<double>();
    j.overflow = ex.overflow.as<double>();
    j.warp_cost = ex.cost.as<double>();
    j.n_slots = ex.n_warps * kSlotsPerWarp;
    j.n_keys = ex.n_keys; j.rd = ex.rd; j.n_warps = ex.n_warps;
    j.cta0 = 0;
    CU(cudaMemcpyAsync(d_job.p, &j, sizeof(j), cudaMemcpyHostToDevice, E->stream));
    const size_t sm = (size_t)2 * j.n_slots * sizeof(int);
    if (sm > 32 * 1024) CU(cudaFuncSetAttribute(reduce_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    reduce_all_kernel<<<ex.n_keys + 1, 256, sm, E->stream>>>(d_job.as<ReduceJob>(), 1, ex.n_keys, j.n_slots,
                                                             E->cost_loc() + 4, nullptr);
    CU(cudaStreamSynchronize(E->stream));  // `j` is a local
    E->launches++;
    return CLIC_OK;
  };
  auto mark = [&](const StreamExec& ex, int r_col) -> int {
    key_presence_kernel<<<(ex.n_keys + 63) / 64, 64, 0, E->stream>>>(ex.fin.as<double>(), ex.n_keys, ex.rd, ex.NT, r_col,
                                                                    d_present.as<int>());
    E->launches++;
    CU(cudaMemcpyAsync(present.data(), d_present.p, ex.n_keys * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    for (int s = 0; s < ex.n_keys; s++)
      if (present[s])
        for (int k = s; k < s + 4; k++) knot_used[k] = 1;
    return CLIC_OK;
  };
  bool any_imu = false;
  std::vector<char> bias_used(2 * std::max(1, E->n_bias), 0), bias_drop(2 * std::max(1, E->n_bias), 0);
  E->prof_begin(7);
  for (auto& b : E->loam) {
    if (!b->marg) continue;
    StreamExec& ex = b->ex;
    if (b->st.geo_kind == 1) loam_kernel<true, kLoamWarps, kLoamMinB, 1><<<ex.grid, kLoamWarps * 32, loam_smem_bytes(K), E->stream>>>(b->st, pp, ex.slabs_per_warp, ex.rb());
    else if (b->st.geo_kind == 2) loam_kernel<true, kLoamWarps, kLoamMinB, 2><<<ex.grid, kLoamWarps * 32, loam_smem_bytes(K), E->stream>>>(b->st, pp, ex.slabs_per_warp, ex.rb());
    else loam_kernel<true, kLoamWarps, kLoamMinB, 0><<<ex.grid, kLoamWarps * 32, loam_smem_bytes(K), E->stream>>>(b->st, pp, ex.slabs_per_warp, ex.rb());
    E->launches += 1;
    if ((rc = reduce_stream(ex)) || (rc = mark(ex, -1))) return rc;
  }
  for (auto& b : E->imu) {
    if (!b->marg) continue;
    StreamExec& ex = b->ex;
    imu_kernel<true, true><<<ex.grid, kThreads, factor_smem_bytes(K, 5), E->stream>>>(b->st, pp, ex.slabs_per_warp, ex.rb());
    E->launches += 1;
    if ((rc = reduce_stream(ex)) || (rc = mark(ex, imu_phys_col<true>(34)))) return rc;
    bool used = false;
    for (int s = 0; s < ex.n_keys; s++) used = used || present[s];
    if (used) {
      any_imu = true;
      bias_used[b->st.bias_slot] = bias_used[E->n_bias + b->st.bias_slot] = 1;
      if (E->opt.marg_bias_param) bias_drop[b->st.bias_slot] = bias_drop[E->n_bias + b->st.bias_slot] = 1;
    }
  }
  // visual factors (UpdateVisualOffsetPrior, trajectory_manager.cpp:466-553): every marginalised feature drops its
  // inverse depth, so every such factor is in the set; first run = records only (which knots are touched decides the
  // column layout), the landmark side rows are scattered in a second run once the columns are known.
  std::vector<char> lm_used(std::max(1, E->n_lm), 0);
  bool any_image = false;
  auto image_shared = [&](const ImageBatch& b) {
    ImageShared ish;
    ish.S_CtoI = ldq(E->S_CtoI);
    ish.p_CinI = ldv(E->p_CinI);
    ish.sqrt_info = E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450.0 / 1.5;
    ish.loss_a = b.st.loss_a;
    ish.inv_dt = 1e9 / double(E->dt_ns);
    ish.want_toff = true;  // the marginalization evaluates every Jacobian block
    return ish;
  };
  auto launch_image_marg = [&](ImageBatch& b) -> int {
    StreamExec& ex = b.ex;
    const ImageShared ish = image_shared(b);
    if (b.st.n_packs > 0)
      image_kernel<true, true><<<ex.grid, kThreads, image_smem_bytes(K, b.st.n_packs), E->stream>>>(b.st, pp, ex.slabs_per_warp, ex.rb(), ish);
    else
      image_kernel<true, false><<<ex.grid, kThreads, factor_smem_bytes(K, 7), E->stream>>>(b.st, pp, ex.slabs_per_warp, ex.rb(), ish);
    E->launches += 1;
    return reduce_stream(ex);
  };
  std::vector<int> present2((size_t)(K - 3) * (K - 3));
  DevBuf d_present2;
  for (auto& b : E->image) {
    if (!b->marg) continue;
    StreamExec& ex = b->ex;
    b->st.rho_side = nullptr;
    b->st.lm_col = nullptr;
    if ((rc = launch_image_marg(*b))) return rc;
    CU(d_present2.ensure(present2.size() * sizeof(int)));
    key_presence_kernel<<<(ex.n_keys + 63) / 64, 64, 0, E->stream>>>(ex.fin.as<double>(), ex.n_keys, ex.rd, ex.NT, 49,
                                                                    d_present2.as<int>());
    E->launches++;
    CU(cudaMemcpyAsync(present2.data(), d_present2.p, ex.n_keys * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    const int nk = K - 3;
    for (int key = 0; key < ex.n_keys; key++)
      if (present2[key]) {
        any_image = true;
        for (int k = 0; k < 4; k++) knot_used[key / nk + k] = knot_used[key % nk + k] = 1;
      }
    for (int l = 0; l < E->n_lm; l++) lm_used[l] = lm_used[l] || b->lm_used[l];
  }
  for (auto& bf : E->bias_f) {
    if (!bf.marg) continue;
    for (int sl : {bf.slot_i, bf.slot_j}) bias_used[sl] = bias_used[E->n_bias + sl] = 1;
    if (E->opt.marg_bias_param) {  // trajectory_estimator.cpp:483-489
      bias_drop[bf.slot_i] = bias_drop[E->n_bias + bf.slot_i] = 1;
      if (bf.marg_all) bias_drop[bf.slot_j] = bias_drop[E->n_bias + bf.slot_j] = 1;
    }
  }
  trace.lap("factor kernels + presence");
  // ---- blocks in canonical order, dropped first (marginalization_factor.cpp:154-166) ----
  struct Blk { int kind, id, size, used, drop, idx; };
  std::vector<Blk> blocks;
  std::vector<char> knot_drop(K, 0);
  bool any_grav_f = false;
  for (auto& gf : E->grav_f) any_grav_f = any_grav_f || gf.marg;
  // an IMU factor or a GravityFactor of the marginalization set brings the gravity block in; dropped iff marg_gravity_param
  // (trajectory_estimator.cpp:434-440, 515-519)
  bool grav_used = any_imu || any_grav_f, grav_drop = (any_imu || any_grav_f) && E->opt.marg_gravity_param;
  bool toff_used[3] = {any_imu, false, any_image};
  bool toff_drop[3] = {any_imu && E->opt.marg_t_offset_param != 0, false, any_image && E->opt.marg_t_offset_param != 0};
  for (auto& pr : E->priors) {
    std::vector<char> dropped(pr->p->blocks.size(), 0);
    for (int d : pr->drop) if (d >= 0 && d < (int)dropped.size()) dropped[d] = 1;
    for (size_t i = 0; i < pr->p->blocks.size(); i++) {
      const auto& b = pr->p->blocks[i];
      switch (b.kind) {
        case CLIC_BLOCK_KNOT_ROT:
        case CLIC_BLOCK_KNOT_POS: knot_used[b.id - E->knot_id0] = 1; if (dropped[i]) knot_drop[b.id - E->knot_id0] = 1; break;
        case CLIC_BLOCK_BIAS_GYRO: bias_used[b.id - E->bias_id0] = 1; if (dropped[i]) bias_drop[b.id - E->bias_id0] = 1; break;
        case CLIC_BLOCK_BIAS_ACCEL: bias_used[E->n_bias + b.id - E->bias_id0] = 1; if (dropped[i]) bias_drop[E->n_bias + b.id - E->bias_id0] = 1; break;
        case CLIC_BLOCK_GRAVITY: grav_used = true; if (dropped[i]) grav_drop = true; break;
        case CLIC_BLOCK_T_OFFSET: toff_used[b.id] = true; if (dropped[i]) toff_drop[b.id] = true; break;
        default: return fail(CLIC_ERR_UNSUPPORTED, "inverse-depth blocks in a prior are not built on the GPU yet");
      }
    }
  }
  for (int k = 0; k < K; k++) {
    if (!knot_used[k]) continue;
    const int drop = (drop_knots && k < later_local) || knot_drop[k] ? 1 : 0;
    blocks.push_back({CLIC_BLOCK_KNOT_ROT, k, 4, 1, drop, -1});
    blocks.push_back({CLIC_BLOCK_KNOT_POS, k, 3, 1, drop, -1});
  }
  for (int j = 0; j < E->n_bias; j++) if (bias_used[j]) blocks.push_back({CLIC_BLOCK_BIAS_GYRO, j, 3, 1, bias_drop[j], -1});
  for (int j = 0; j < E->n_bias; j++)
    if (bias_used[E->n_bias + j]) blocks.push_back({CLIC_BLOCK_BIAS_ACCEL, j, 3, 1, bias_drop[E->n_bias + j], -1});
  if (grav_used) blocks.push_back({CLIC_BLOCK_GRAVITY, 0, 3, 1, grav_drop ? 1 : 0, -1});
  for (int sn = 0; sn < 3; sn++) if (toff_used[sn]) blocks.push_back({CLIC_BLOCK_T_OFFSET, sn, 1, 1, toff_drop[sn] ? 1 : 0, -1});
  int n_rho = 0;
  if (any_image)  // the inverse depth of a marginalised feature is always dropped (trajectory_estimator.cpp:813-818)
    for (int l = 0; l < E->n_lm; l++)
      if (lm_used[l]) { blocks.push_back({CLIC_BLOCK_INV_DEPTH, l, 1, 1, 1, -1}); n_rho++; }
  int pos = 0;
  for (auto& b : blocks) if (b.drop) { b.idx = pos; pos += b.size == 4 ? 3 : b.size; }
  const int m = pos;
  for (auto& b : blocks) if (!b.drop) { b.idx = pos; pos += b.size == 4 ? 3 : b.size; }
  const int n = pos - m;
  if (n <= 0 || blocks.empty()) return fail(CLIC_ERR_NOTHING_TO_KEEP, "marginalization leaves no parameters");
  auto find = [&](int kind, int id) -> int {
    for (auto& b : blocks) if (b.kind == kind && b.id == id) return b.idx;
    return -1;
  };
  // ---- column maps of the marg system ----
  std::vector<int> knot_col(K, -1), col_knot(pos, -1), col_sub(pos, 0);
  for (int k = 0; k < K; k++) {
    int c = find(CLIC_BLOCK_KNOT_ROT, k);
    knot_col[k] = c;
    if (c >= 0) for (int a = 0; a < 6; a++) { col_knot[c + a] = k; col_sub[c + a] = a; }
  }
  DevBuf d_kc, d_ck, d_cs, d_desc, d_A, d_b, d_J0, d_r0, d_status, d_cost, d_perm, d_dscr, d_aug;
  if ((rc = to_device(E, d_kc, knot_col.data(), K)) || (rc = to_device(E, d_ck, col_knot.data(), pos)) ||
      (rc = to_device(E, d_cs, col_sub.data(), pos)))
    return rc;
  std::vector<StreamDesc> descs;
  for (auto& b : E->loam) {
    if (!b->marg) continue;
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    d.r_col = -1; d.b_off = 384; d.n_extra = 0; d.extra_base = 24; d.kind = 0; d.kdim = K - 3;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)l;
    descs.push_back(d);
  }
  for (auto& b : E->imu) {
    if (!b->marg) continue;
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    d.r_col = 34; d.b_off = -1; d.n_extra = 10; d.extra_base = 24; d.kind = 0; d.kdim = K - 3;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)(l < 40 ? imu_phys_col<true>(l) : l);
    const int cg = find(CLIC_BLOCK_BIAS_GYRO, b->st.bias_slot), ca = find(CLIC_BLOCK_BIAS_ACCEL, b->st.bias_slot),
              cgr = find(CLIC_BLOCK_GRAVITY, 0), ct = find(CLIC_BLOCK_T_OFFSET, CLIC_SENSOR_IMU);
    for (int c = 0; c < 3; c++) {
      d.extra_global[c] = cg >= 0 ? cg + c : -1;
      d.extra_global[3 + c] = ca >= 0 ? ca + c : -1;
      d.extra_global[6 + c] = cgr >= 0 ? cgr + c : -1;
    }
    d.extra_global[9] = ct;
    descs.push_back(d);
  }
  DevBuf d_lm_col, d_side;
  int rho0 = 0;
  if (n_rho > 0) {
    std::vector<int32_t> lm_col(E->n_lm, -1);
    rho0 = pos;
    for (int l = 0; l < E->n_lm; l++)
      if (lm_used[l]) { lm_col[l] = find(CLIC_BLOCK_INV_DEPTH, l); rho0 = std::min(rho0, (int)lm_col[l]); }
    if ((rc = to_device(E, d_lm_col, lm_col.data(), lm_col.size()))) return rc;
    CU(d_side.ensure((size_t)n_rho * (pos + 2) * sizeof(double)));
    CU(cudaMemsetAsync(d_side.p, 0, (size_t)n_rho * (pos + 2) * sizeof(double), E->stream));
    CU(cudaStreamSynchronize(E->stream));  // lm_col is a local
  }
  for (auto& b : E->image) {
    if (!b->marg) continue;
    if (n_rho > 0) {  // the landmark side rows under the final column layout: per-landmark gather (deterministic)
      ImageStream st = b->st;
      st.rho_side = d_side.as<double>();
      st.lm_col = d_lm_col.as<int32_t>();
      st.knot_col = d_kc.as<int>();
      st.Dtot = pos;
      st.rho0 = rho0;
      st.col_toff_cam = find(CLIC_BLOCK_T_OFFSET, CLIC_SENSOR_CAMERA);
      const ImageShared ish = image_shared(*b);
      const size_t smem = rho_side_smem_doubles(K, pos) * sizeof(double);
      CU(cudaFuncSetAttribute(rho_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      PassParams ppm = pass_params(E, E->d_state.as<double>(), nullptr);
      ppm.marg_mode = 1;
      const int grid = std::max(1, std::min(2 * E->n_sm, (E->n_lm + kRhoWarps - 1) / kRhoWarps));
      rho_side_kernel<<<grid, kRhoWarps * 32, smem, E->stream>>>(st, ppm, ish, (const int*)b->lm_start.as<int>(),
                                                                 (const int*)b->lm_obs.as<int>(), E->n_lm);
      E->launches++;
    }
    StreamDesc d{};
    d.fin = b->ex.fin.as<double>();
    d.n_keys = b->ex.n_keys; d.NT = b->ex.NT; d.rd = b->ex.rd;
    d.r_col = 49; d.b_off = -1; d.n_extra = 1; d.extra_base = 48; d.kind = 1; d.kdim = K - 3;
    for (int l = 0; l < 64; l++) d.perm[l] = (unsigned char)l;
    d.extra_global[0] = find(CLIC_BLOCK_T_OFFSET, CLIC_SENSOR_CAMERA);
    descs.push_back(d);
  }
  if ((rc = to_device(E, d_desc, descs.data(), descs.size()))) return rc;
  CU(d_A.ensure((size_t)pos * pos * sizeof(double)));
  CU(d_b.ensure((size_t)pos * sizeof(double)));
  CU(d_J0.ensure((size_t)n * n * sizeof(double)));
  CU(d_r0.ensure((size_t)n * sizeof(double)));
  CU(d_status.ensure(4 * sizeof(int)));
  CU(d_perm.ensure((size_t)pos * sizeof(int)));
  CU(d_dscr.ensure((size_t)3 * pos * sizeof(double)));
  CU(d_aug.ensure((size_t)(pos + 1) * (pos + 1) * sizeof(double)));
  CU(d_cost.ensure(2 * sizeof(double)));
  CU(cudaMemsetAsync(d_cost.p, 0, 2 * sizeof(double), E->stream));
  ColMaps cm;
  cm.D = pos; cm.K = K;
  cm.knot_col = d_kc.as<int>(); cm.col_knot = d_ck.as<int>(); cm.col_sub = d_cs.as<int>();
  const int n_entries = pos * (pos + 1) / 2 + pos;
  assemble_kernel<<<(n_entries + 7) / 8, 256, 0, E->stream>>>(
      d_A.as<double>(), d_b.as<double>(), cm, d_desc.as<StreamDesc>(), (int)descs.size(), nullptr, nullptr, 0, nullptr, 0,
      nullptr, n_rho > 0 ? (const double*)d_side.as<double>() : (const double*)nullptr, rho0, n_rho);
  E->launches++;
  for (auto& bf : E->bias_f) {
    if (!bf.marg) continue;
    BiasFactorDesc d;
    d.slot_i = bf.slot_i; d.slot_j = bf.slot_j;
    for (int i = 0; i < 6; i++) d.sqrt_info[i] = bf.sqrt_info[i];
    d.col_gi = find(CLIC_BLOCK_BIAS_GYRO, bf.slot_i); d.col_gj = find(CLIC_BLOCK_BIAS_GYRO, bf.slot_j);
    d.col_ai = find(CLIC_BLOCK_BIAS_ACCEL, bf.slot_i); d.col_aj = find(CLIC_BLOCK_BIAS_ACCEL, bf.slot_j);
    bias_factor_kernel<<<1, 32, 0, E->stream>>>(d, E->d_state.as<double>(), E->sl, d_A.as<double>(), d_b.as<double>(), pos,
                                                d_cost.as<double>(), 1, nullptr);
    E->launches++;
  }
  for (auto& gf : E->grav_f) {
    if (!gf.marg) continue;
    const int col = find(CLIC_BLOCK_GRAVITY, 0);
    gravity_factor_kernel<<<1, 32, 0, E->stream>>>(gf.g0[0], gf.g0[1], gf.g0[2], gf.sqrt_info[0], gf.sqrt_info[1],
                                                   gf.sqrt_info[2], E->d_state.as<double>() + E->sl.off_grav(), col,
                                                   d_A.as<double>(), d_b.as<double>(), pos, d_cost.as<double>());
    E->launches++;
  }
  for (auto& pr : E->priors) {
    const int nb = (int)pr->p->blocks.size();
    for (int i = 0; i < nb; i++) {
      const auto& b = pr->p->blocks[i];
      int id = b.id;
      if (b.kind == CLIC_BLOCK_KNOT_ROT || b.kind == CLIC_BLOCK_KNOT_POS) id -= E->knot_id0;
      if (b.kind == CLIC_BLOCK_BIAS_GYRO || b.kind == CLIC_BLOCK_BIAS_ACCEL) id -= E->bias_id0;
      pr->h_meta[4 * i + 3] = find(b.kind, id);
    }
    CU(cudaMemcpyAsync(pr->meta.p, pr->h_meta.data(), pr->h_meta.size() * sizeof(int), cudaMemcpyHostToDevice, E->stream));
    const int pn = pr->p->n;
    size_t smem = (size_t)3 * pn * sizeof(double) + (size_t)pn * sizeof(int);
    prior_kernel<<<1, 1024, smem, E->stream>>>(pr->J0.as<double>(), pr->A.as<double>(), pr->r0.as<double>(),
                                              pr->x0.as<double>(), pr->meta.as<int>(), pn, nb, E->d_state.as<double>(),
                                              d_A.as<double>(), d_b.as<double>(), pos, d_cost.as<double>(), 1, nullptr);
    E->launches++;
    CU(cudaStreamSynchronize(E->stream));  // h_meta is reused by the solve path
  }
  if (ScopedTrace::on()) cudaStreamSynchronize(E->stream);
  trace.lap("assemble + bias + prior");
  int status4[4] = {-1, -1, -1, -1};
  int& status = status4[0];
  CU(cudaMemsetAsync(d_status.p, 0xff, 4 * sizeof(int), E->stream));
  schur_kernel<<<1, 512, 0, E->stream>>>(d_A.as<double>(), d_b.as<double>(), m, n, d_J0.as<double>(), d_r0.as<double>(),
                                         d_perm.as<int>(), d_dscr.as<double>(),
                                         getenv("CLIC_SCHUR_TWO_STAGE") ? nullptr : d_aug.as<double>(), d_status.as<int>());
  E->launches++;
  E->prof_end();
  std::unique_ptr<clic_prior> P(new clic_prior);
  P->n = n;
  P->m = m;
  P->J0.resize((size_t)n * n);
  P->r0.resize(n);
  std::vector<double> h(E->sl.size());
  CU(cudaMemcpyAsync(status4, d_status.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(P->J0.data(), d_J0.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(P->r0.data(), d_r0.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(h.data(), E->d_state.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  CU(cudaGetLastError());
  trace.lap("schur + read back");
  if (ScopedTrace::on())
    fprintf(stderr, "[clic trace] schur: m %d n %d status %d rank deficiency A_mm %d, Schur complement %d\n", m, n, status4[0],
            status4[1], status4[2]);
  if (status != 0) return fail(CLIC_ERR_CUDA, "schur kernel did not complete");
  for (auto& b : blocks) {
    if (b.drop) continue;
    clic_prior::Block pb;
    pb.kind = b.kind;
    pb.size = b.size;
    pb.idx = b.idx - m;
    pb.id = b.id;
    const double* src = nullptr;
    switch (b.kind) {
      case CLIC_BLOCK_KNOT_ROT: pb.id += E->knot_id0; src = &h[E->sl.off_q() + 4 * b.id]; break;
      case CLIC_BLOCK_KNOT_POS: pb.id += E->knot_id0; src = &h[E->sl.off_p() + 3 * b.id]; break;
      case CLIC_BLOCK_BIAS_GYRO: pb.id += E->bias_id0; src = &h[E->sl.off_bias() + 6 * b.id]; break;
      case CLIC_BLOCK_BIAS_ACCEL: pb.id += E->bias_id0; src = &h[E->sl.off_bias() + 6 * b.id + 3]; break;
      case CLIC_BLOCK_GRAVITY: src = &h[E->sl.off_grav()]; break;
      default: src = &h[E->sl.off_toff() + b.id]; break;
    }
    for (int k = 0; k < 4; k++) pb.x0[k] = k < b.size ? src[k] : 0.0;
    P->blocks.push_back(pb);
  }
  *out = P.release();
  return CLIC_OK;
}

int upload_state(clic_estimator* E) {
  CU(cudaMemcpyAsync(E->d_state.p, E->h_state.data(), E->h_state.size() * sizeof(double), cudaMemcpyHostToDevice,
                     E->stream));
  return CLIC_OK;
}

bool ok_handle(const clic_estimator* h) { return h != nullptr; }

}  // namespace

extern "C" {

CLIC_API const char* clic_last_error(void) { return g_err.c_str(); }
CLIC_API const char* clic_backend(void) { return "cuda-sm100a"; }

CLIC_API void clic_default_options(clic_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->lock_ab = o->lock_wb = o->lock_g = 1;
  for (int s = 0; s < 3; s++) o->lock_t_offset[s] = 1;
  o->t_offset_padding_ns = (int64_t)1e7;
  o->marg_bias_param = o->marg_gravity_param = o->marg_t_offset_param = 1;
  o->deterministic = 1;
  o->image_sqrt_info = 450.0 / 1.5;  // ImageFeatureFactor's class default (image_feature_factor.h:277-278)
}

CLIC_API int clic_create(const clic_window_desc* w, const clic_options* opt, clic_estimator** out) {
  ScopedTrace trace("clic_create");
  if (!w || !opt || !out) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (w->n_knots < 4 || w->dt_ns <= 0) return fail(CLIC_ERR_BAD_ARGUMENT, "need >= 4 knots and dt > 0");
  if (!have_device()) return fail(CLIC_ERR_NO_DEVICE, "no CUDA device: libclic_cuda has no CPU fallback");
  std::unique_ptr<clic_estimator> E(new clic_estimator);
  E->device = opt->device;
  CU(cudaSetDevice(E->device));
  // device queries are slow (~ms): once per process and device
  static int cached_dev = -1, cached_major = 0, cached_sms = 0;
  if (cached_dev != E->device) {
    int major = 0, sms = 0;
    CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, E->device));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, E->device));
    cached_dev = E->device; cached_major = major; cached_sms = sms;
  }
  if (cached_major < 10) return fail(CLIC_ERR_NO_DEVICE, "device is not sm_100 class (kernels are built for sm_100a only)");
  E->n_sm = cached_sms;
  E->t0_ns = w->t0_ns;
  E->dt_ns = w->dt_ns;
  E->K = w->n_knots;
  E->knot_id0 = w->knot_id0;
  E->n_bias = w->n_bias;
  E->bias_id0 = w->bias_id0;
  E->n_lm = w->n_landmarks;
  std::memcpy(E->S_CtoI, w->S_CtoI, sizeof(E->S_CtoI));
  std::memcpy(E->p_CinI, w->p_CinI, sizeof(E->p_CinI));
  E->opt = *opt;
  E->sl.K = E->K;
  E->sl.nb = E->n_bias;
  E->sl.L = E->n_lm;
  E->h_state.assign(E->sl.size(), 0.0);
  std::copy(w->knot_q, w->knot_q + 4 * E->K, E->h_state.begin() + E->sl.off_q());
  std::copy(w->knot_p, w->knot_p + 3 * E->K, E->h_state.begin() + E->sl.off_p());
  if (E->n_bias > 0) std::copy(w->bias, w->bias + 6 * E->n_bias, E->h_state.begin() + E->sl.off_bias());
  for (int s = 0; s < 3; s++) E->h_state[E->sl.off_toff() + s] = w->t_offset_ns[s];
  for (int a = 0; a < 3; a++) E->h_state[E->sl.off_grav() + a] = w->gravity[a];
  if (E->n_lm > 0) std::copy(w->inv_depth, w->inv_depth + E->n_lm, E->h_state.begin() + E->sl.off_invd());
  E->lm_free.assign(std::max(0, E->n_lm), 0);
  if (opt->stream) {
    E->stream = (cudaStream_t)opt->stream;
  } else {
    CU(cudaStreamCreateWithFlags(&E->stream, cudaStreamNonBlocking));
    E->own_stream = true;
  }
  {  // one copy stream per process and device, shared by every estimator (ordering is by per-batch events)
    static std::vector<cudaStream_t> copy_streams;
    if ((int)copy_streams.size() <= E->device) copy_streams.resize(E->device + 1, nullptr);
    if (!copy_streams[E->device]) CU(cudaStreamCreateWithFlags(&copy_streams[E->device], cudaStreamNonBlocking));
    E->copy_stream = copy_streams[E->device];
  }
  static int pool_dev = -1;
  if (pool_dev != E->device) {
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, E->device));
    uint64_t thr = UINT64_MAX;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    pool_dev = E->device;
  }
  AllocScope alloc_scope(E->stream);
  CU(E->d_state.ensure(E->sl.size() * sizeof(double)));
  CU(E->d_cand.ensure(E->sl.size() * sizeof(double)));
  CU(E->d_counter.ensure(2 * sizeof(unsigned long long)));  // [0]: adders (copy stream); [1]: queries / scan association (compute stream)
  CU(E->d_flags.ensure(16 * sizeof(int)));
  CU(cudaMemsetAsync(E->d_flags.p, 0, 16 * sizeof(int), E->stream));
  int rc = upload_state(E.get());
  if (rc) return rc;
  // opt in to > 48 KB dynamic shared memory once
  size_t smem4 = imu_smem_bytes(E->K), smem5 = factor_smem_bytes(E->K, 5);
  static int attr_K = -1, attr_dev = -1;  // the opt-in is per function and device; only grows with K
  if (attr_dev != E->device || E->K > attr_K) {
  if ((rc = set_smem(loam_kernel<true, kLoamWarps, kLoamMinB, 0>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(loam_kernel<false, kLoamWarps, kLoamMinB, 0>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(loam_kernel<true, kLoamWarps, kLoamMinB, 1>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(loam_kernel<false, kLoamWarps, kLoamMinB, 1>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(loam_kernel<true, kLoamWarps, kLoamMinB, 2>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(loam_kernel<false, kLoamWarps, kLoamMinB, 2>, loam_smem_bytes(E->K))) ||
      (rc = set_smem(imu_kernel<true, false>, smem4)) || (rc = set_smem(imu_kernel<false, false>, smem4)) ||
      (rc = set_smem(imu_kernel<true, true>, smem5)) ||
      (rc = set_smem(image_kernel<true, false>, factor_smem_bytes(E->K, 7))) ||
      (rc = set_smem(image_kernel<false, false>, factor_smem_bytes(E->K, 7))))
    return rc;
  E->packs_fit = image_smem_bytes(E->K, kMaxPosePacks) <= (size_t)227 * 1024;
  if (E->packs_fit && ((rc = set_smem(image_kernel<true, true>, image_smem_bytes(E->K, kMaxPosePacks))) ||
                       (rc = set_smem(image_kernel<false, true>, image_smem_bytes(E->K, kMaxPosePacks)))))
    return rc;
  attr_K = E->K;
  attr_dev = E->device;
  }
  CU(cudaStreamSynchronize(E->stream));
  *out = E.release();
  return CLIC_OK;
}

CLIC_API int clic_destroy(clic_estimator* h) {
  ScopedTrace trace("clic_destroy");
  if (!h) return CLIC_OK;
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  const bool own = h->own_stream;
  cudaStreamSynchronize(h->copy_stream);
  cudaStreamSynchronize(st);
  for (auto& sp : h->prof_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  delete h;  // buffers go back to the stream-ordered pool
  if (own) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  return CLIC_OK;
}

CLIC_API int clic_set_fixed_index(clic_estimator* h, int32_t idx) {
  if (!ok_handle(h)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  h->fixed_index_global = idx;
  h->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_add_loam(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                           const int32_t* geo_type, const double* geo_plane, const double* geo_normal,
                           const double* geo_point, const double S_GtoM[4], const double p_GinM[3],
                           const double S_LtoI[4], const double p_LinI[3], double weight, int32_t marg_this_factor,
                           int64_t* n_dropped) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_dropped) *n_dropped = 0;
  if (n <= 0) return CLIC_OK;
  CU(cudaSetDevice(E->device));
  cudaStream_t cs = E->copy_stream;
  AllocScope alloc_scope(cs);
  std::unique_ptr<LoamBatch> B(new LoamBatch);
  // MeasuredTimeToNs(LiDARSensor): host-side scalar bounds in the reference's double arithmetic
  const int S = CLIC_SENSOR_LIDAR;
  int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  if (!E->opt.lock_t_offset[S]) {
    t_min += E->opt.t_offset_padding_ns;
    t_max -= E->opt.t_offset_padding_ns;
  }
  // raw (as-uploaded) arrays stay with the batch until the estimator is destroyed: freeing them while their copy is
  // still in flight makes the stream-ordered pool's next large allocation unpredictably slow (measured: 1-30 ms)
  DevBuf &r_t = B->raw[0], &r_pt = B->raw[1], &r_type = B->raw[2], &r_plane = B->raw[3], &r_normal = B->raw[4], &r_gpoint = B->raw[5];
  int rc;
  if ((rc = to_device(E, r_t, t_point, n)) || (rc = to_device(E, r_pt, point, 3 * n)) ||
      (rc = to_device(E, r_type, geo_type, n)))
    return rc;
  // a missing geometry array is only legal when no record of that type exists: checked on the device by the prep
  // kernel (flag word 1), reported by the next evaluate / solve
  if ((rc = to_device(E, r_plane, geo_plane, geo_plane ? 4 * n : 0)) ||
      (rc = to_device(E, r_normal, geo_normal, geo_normal ? 3 * n : 0)) ||
      (rc = to_device(E, r_gpoint, geo_point, geo_point ? 3 * n : 0)))
    return rc;
  B->n_input = n;
  B->gather = build_key_gather(E, t_point, n, t_min, t_max, 0, 0);
  if (B->gather.empty()) key_change_slabs(E, t_point, n, 0, 32, B->ex.key_change_slab);
  const int32_t* d_gather = nullptr;
  if (!B->gather.empty()) {
    if ((rc = to_device(E, B->d_gather, B->gather.data(), B->gather.size()))) return rc;
    d_gather = B->d_gather.as<int32_t>();
    n = (int64_t)B->gather.size();  // stored length: every interval's factors start on a slab boundary
  }
  {
    const size_t nd = (size_t)n * sizeof(double);
    CU(carve(B->arena, {{&B->t_ns, (size_t)n * sizeof(int64_t)}, {&B->p[0], nd}, {&B->p[1], nd}, {&B->p[2], nd},
                        {&B->g[0], nd}, {&B->g[1], nd}, {&B->g[2], nd}, {&B->g[3], nd}, {&B->g[4], nd}, {&B->g[5], nd},
                        {&B->type, (size_t)n}}));
  }
  CU(cudaMemsetAsync(E->d_counter.p, 0, sizeof(unsigned long long), cs));
  prep_loam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(
      n, r_t.as<double>(), r_pt.as<double>(), r_type.as<int32_t>(), r_plane.as<double>(), r_normal.as<double>(),
      r_gpoint.as<double>(), ldq(S_LtoI), ldv(p_LinI), t_min, t_max, B->t_ns.as<int64_t>(), B->p[0].as<double>(), B->p[1].as<double>(),
      B->p[2].as<double>(), B->g[0].as<double>(), B->g[1].as<double>(), B->g[2].as<double>(), B->g[3].as<double>(),
      B->g[4].as<double>(), B->g[5].as<double>(), B->type.as<uint8_t>(), E->d_counter.as<unsigned long long>(),
      E->d_flags.as<int>() + 1, d_gather);
  E->launches++;
  if (n_dropped) {  // the count is only read back (one synchronisation) when the caller asks for it
    unsigned long long dropped = 0;
    CU(cudaMemcpyAsync(&dropped, E->d_counter.p, sizeof(dropped), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    *n_dropped = (int64_t)dropped;
  }  // the temporary raw buffers are freed in stream order (cudaFreeAsync), no host wait needed
  B->st.n = n;
  B->st.t_ns = B->t_ns.as<int64_t>();
  for (int k = 0; k < 3; k++) B->st.p[k] = B->p[k].as<double>();
  for (int k = 0; k < 6; k++) B->st.g[k] = B->g[k].as<double>();
  B->st.type = B->type.as<uint8_t>();
  B->st.geo_kind = 0;
  B->st.sh.S_GtoM = ldq(S_GtoM);
  B->st.sh.p_GinM = ldv(p_GinM);
  B->st.sh.S_LtoI = ldq(S_LtoI);
  B->st.sh.p_LinI = ldv(p_LinI);
  B->st.sh.weight = weight;
  B->st.sh.point_in_imu = true;  // the prep kernel stored S_LtoI p_L + p_LinI
  B->st.sh.gm_identity = S_GtoM[0] == 0.0 && S_GtoM[1] == 0.0 && S_GtoM[2] == 0.0 && S_GtoM[3] == 1.0 &&
                         p_GinM[0] == 0.0 && p_GinM[1] == 0.0 && p_GinM[2] == 0.0;
  B->marg = (E->opt.is_marg_state && marg_this_factor) ? 1 : 0;
  if ((rc = plan_stream(E, B->ex, n, 3, kLoamMinB, false, 1, kLoamWarps))) return rc;
  if ((rc = mark_ready(E, B->ex, cs))) return rc;
  E->loam.push_back(std::move(B));
  E->layout_dirty = true;
  return CLIC_OK;
}

namespace {
// geo_kind 0: mixed (geo_type per feature, 6 geometry doubles each); 1: planes (4 doubles); 2: lines (6 doubles)
// device_inputs: t_point / point / geo_type / geo already live on the device and were produced on stream `cs`
// (clic_add_loam_from_scan); otherwise they are host arrays and go through the copy stream.
int add_loam_records(clic_estimator* E, int geo_kind, int64_t n, const double* t_point, const double* point,
                     const uint8_t* geo_type, const double* geo, const double S_GtoM[4], const double p_GinM[3],
                     const double S_LtoI[4], const double p_LinI[3], double weight, int32_t marg_this_factor,
                     int64_t* n_dropped, bool device_inputs = false, cudaStream_t dev_stream = nullptr) {
  ScopedTrace trace("clic_add_loam_records");
  const int gstride = geo_kind == 1 ? 4 : 6;
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_dropped) *n_dropped = 0;
  if (n <= 0) return CLIC_OK;
  if (!t_point || !point || (geo_kind == 0 && !geo_type) || !geo)
    return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_loam_*: null array");
  CU(cudaSetDevice(E->device));
  cudaStream_t cs = device_inputs ? dev_stream : E->copy_stream;
  AllocScope alloc_scope(cs);
  std::unique_ptr<LoamBatch> B(new LoamBatch);
  // MeasuredTimeToNs(LiDARSensor): host-side scalar bounds in the reference's double arithmetic
  const int S = CLIC_SENSOR_LIDAR;
  int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  if (!E->opt.lock_t_offset[S]) {
    t_min += E->opt.t_offset_padding_ns;
    t_max -= E->opt.t_offset_padding_ns;
  }
  DevBuf &r_t = B->raw[0], &r_pt = B->raw[1], &r_type = B->raw[2], &r_geo = B->raw[3];
  int rc;
  const double *in_t = t_point, *in_pt = point, *in_geo = geo;
  const uint8_t* in_type = geo_type;
  if (!device_inputs) {
    if ((rc = to_device(E, r_t, t_point, n)) || (rc = to_device(E, r_pt, point, 3 * n)) ||
        (geo_kind == 0 && (rc = to_device(E, r_type, geo_type, n))) || (rc = to_device(E, r_geo, geo, gstride * n)))
      return rc;
    in_t = r_t.as<double>();
    in_pt = r_pt.as<double>();
    in_type = r_type.as<uint8_t>();
    in_geo = r_geo.as<double>();
  }
  B->n_input = n;
  const int32_t* d_gather = nullptr;
  if (!device_inputs) {
    B->gather = build_key_gather(E, t_point, n, t_min, t_max, 0, 0);
    if (B->gather.empty()) key_change_slabs(E, t_point, n, 0, 32, B->ex.key_change_slab);
    if (!B->gather.empty()) {
      if ((rc = to_device(E, B->d_gather, B->gather.data(), B->gather.size()))) return rc;
      d_gather = B->d_gather.as<int32_t>();
      n = (int64_t)B->gather.size();  // stored length: every interval's factors start on a slab boundary
    }
  }
  {
    const size_t nd = (size_t)n * sizeof(double), g45 = gstride > 4 ? nd : 0;
    CU(carve(B->arena, {{&B->t_ns, (size_t)n * sizeof(int64_t)}, {&B->p[0], nd}, {&B->p[1], nd}, {&B->p[2], nd},
                        {&B->g[0], nd}, {&B->g[1], nd}, {&B->g[2], nd}, {&B->g[3], nd}, {&B->g[4], g45}, {&B->g[5], g45},
                        {&B->type, geo_kind == 0 ? (size_t)n : 0}}));
    if (gstride <= 4) { B->g[4].view(nullptr, 0); B->g[5].view(nullptr, 0); }
    if (geo_kind != 0) B->type.view(nullptr, 0);
  }
  CU(cudaMemsetAsync(E->d_counter.p, 0, sizeof(unsigned long long), cs));
  prep_loam_packed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(
      n, in_t, in_pt, geo_kind == 0 ? in_type : nullptr, in_geo, gstride,
      geo_kind == 1 ? 1 : 0, ldq(S_LtoI), ldv(p_LinI), t_min, t_max,
      B->t_ns.as<int64_t>(), B->p[0].as<double>(), B->p[1].as<double>(), B->p[2].as<double>(), B->g[0].as<double>(),
      B->g[1].as<double>(), B->g[2].as<double>(), B->g[3].as<double>(), B->g[4].as<double>(), B->g[5].as<double>(),
      geo_kind == 0 ? B->type.as<uint8_t>() : nullptr, E->d_counter.as<unsigned long long>(), d_gather);
  E->launches++;
  if (n_dropped) {  // the count is only read back (one synchronisation) when the caller asks for it
    unsigned long long dropped = 0;
    CU(cudaMemcpyAsync(&dropped, E->d_counter.p, sizeof(dropped), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    *n_dropped = (int64_t)dropped;
  }  // the temporary raw buffers are freed in stream order (cudaFreeAsync), no host wait needed
  B->st.n = n;
  B->st.t_ns = B->t_ns.as<int64_t>();
  for (int k = 0; k < 3; k++) B->st.p[k] = B->p[k].as<double>();
  for (int k = 0; k < 6; k++) B->st.g[k] = B->g[k].as<double>();
  B->st.type = B->type.as<uint8_t>();
  B->st.geo_kind = geo_kind;
  B->st.sh.S_GtoM = ldq(S_GtoM);
  B->st.sh.p_GinM = ldv(p_GinM);
  B->st.sh.S_LtoI = ldq(S_LtoI);
  B->st.sh.p_LinI = ldv(p_LinI);
  B->st.sh.weight = weight;
  B->st.sh.point_in_imu = true;  // the prep kernel stored S_LtoI p_L + p_LinI
  B->st.sh.gm_identity = S_GtoM[0] == 0.0 && S_GtoM[1] == 0.0 && S_GtoM[2] == 0.0 && S_GtoM[3] == 1.0 &&
                         p_GinM[0] == 0.0 && p_GinM[1] == 0.0 && p_GinM[2] == 0.0;
  B->marg = (E->opt.is_marg_state && marg_this_factor) ? 1 : 0;
  if ((rc = plan_stream(E, B->ex, n, 3, kLoamMinB, false, 1, kLoamWarps))) return rc;
  if ((rc = mark_ready(E, B->ex, cs))) return rc;
  E->loam.push_back(std::move(B));
  E->layout_dirty = true;
  return CLIC_OK;
}
}  // namespace

CLIC_API int clic_add_loam_packed(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                  const uint8_t* geo_type, const double* geo, const double S_GtoM[4],
                                  const double p_GinM[3], const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped) {
  return add_loam_records(E, 0, n, t_point, point, geo_type, geo, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                          marg_this_factor, n_dropped);
}
CLIC_API int clic_add_loam_planes(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                  const double* plane, const double S_GtoM[4], const double p_GinM[3],
                                  const double S_LtoI[4], const double p_LinI[3], double weight,
                                  int32_t marg_this_factor, int64_t* n_dropped) {
  return add_loam_records(E, 1, n, t_point, point, nullptr, plane, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                          marg_this_factor, n_dropped);
}
CLIC_API int clic_add_loam_lines(clic_estimator* E, int64_t n, const double* t_point, const double* point,
                                 const double* line, const double S_GtoM[4], const double p_GinM[3],
                                 const double S_LtoI[4], const double p_LinI[3], double weight,
                                 int32_t marg_this_factor, int64_t* n_dropped) {
  return add_loam_records(E, 2, n, t_point, point, nullptr, line, S_GtoM, p_GinM, S_LtoI, p_LinI, weight,
                          marg_this_factor, n_dropped);
}

CLIC_API int clic_add_imu(clic_estimator* E, int64_t n, const double* timestamp, const double* gyro,
                          const double* accel, int32_t bias_slot, const double info_vec[6],
                          int32_t marg_this_factor, int64_t* n_dropped) {
  ScopedTrace trace("clic_add_imu");
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (bias_slot < 0 || bias_slot >= E->n_bias) return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  if (n_dropped) *n_dropped = 0;
  if (n <= 0) return CLIC_OK;
  CU(cudaSetDevice(E->device));
  cudaStream_t cs = E->copy_stream;
  AllocScope alloc_scope(cs);
  std::unique_ptr<ImuBatch> B(new ImuBatch);
  const int S = CLIC_SENSOR_IMU;
  const int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  const int64_t pad = E->opt.lock_t_offset[S] ? 0 : E->opt.t_offset_padding_ns;
  DevBuf &r_t = B->raw[0], &r_g = B->raw[1], &r_a = B->raw[2];
  int rc;
  if ((rc = to_device(E, r_t, timestamp, n)) || (rc = to_device(E, r_g, gyro, 3 * n)) ||
      (rc = to_device(E, r_a, accel, 3 * n)))
    return rc;
  B->n_input = n;
  B->gather = build_key_gather(E, timestamp, n, t_min, t_max, pad, (int64_t)E->h_state[E->sl.off_toff() + S]);
  if (B->gather.empty())
    key_change_slabs(E, timestamp, n, (int64_t)E->h_state[E->sl.off_toff() + S], 32, B->ex.key_change_slab);
  const int32_t* d_gather = nullptr;
  if (!B->gather.empty()) {
    if ((rc = to_device(E, B->d_gather, B->gather.data(), B->gather.size()))) return rc;
    d_gather = B->d_gather.as<int32_t>();
    n = (int64_t)B->gather.size();  // stored length: every interval's factors start on a slab boundary
  }
  {
    const size_t nd = (size_t)n * sizeof(double);
    CU(carve(B->arena, {{&B->t_ns, (size_t)n * sizeof(int64_t)}, {&B->gy[0], nd}, {&B->gy[1], nd}, {&B->gy[2], nd},
                        {&B->ac[0], nd}, {&B->ac[1], nd}, {&B->ac[2], nd}}));
  }
  CU(cudaMemsetAsync(E->d_counter.p, 0, sizeof(unsigned long long), cs));
  prep_imu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(
      n, r_t.as<double>(), r_g.as<double>(), r_a.as<double>(), t_min, t_max, pad, B->t_ns.as<int64_t>(),
      B->gy[0].as<double>(), B->gy[1].as<double>(), B->gy[2].as<double>(), B->ac[0].as<double>(),
      B->ac[1].as<double>(), B->ac[2].as<double>(), E->d_counter.as<unsigned long long>(), d_gather);
  E->launches++;
  if (n_dropped) {
    unsigned long long dropped = 0;
    CU(cudaMemcpyAsync(&dropped, E->d_counter.p, sizeof(dropped), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    *n_dropped = (int64_t)dropped;
  }
  B->st.n = n;
  B->st.t_ns = B->t_ns.as<int64_t>();
  for (int k = 0; k < 3; k++) {
    B->st.gyro[k] = B->gy[k].as<double>();
    B->st.accel[k] = B->ac[k].as<double>();
  }
  for (int i = 0; i < 6; i++) B->st.info[i] = info_vec[i];
  B->st.loss_a = marg_this_factor ? 3.0 : 10.0;  // trajectory_estimator.cpp:426-430
  B->st.bias_slot = bias_slot;
  {
    int rc_tab = imu_xform_table(E->device, &B->st.xform_tab);
    if (rc_tab) return rc_tab;
  }
  B->marg = (E->opt.is_marg_state && marg_this_factor) ? 1 : 0;
  B->st.extras_free = (!E->opt.lock_wb || !E->opt.lock_ab || !E->opt.lock_t_offset[S]) ? 1 : 0;
  B->st.marg_always = (E->opt.marg_bias_param || E->opt.marg_gravity_param || E->opt.marg_t_offset_param) ? 1 : 0;
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {  // trajectory_estimator.cpp:419-424
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->h_state[E->sl.off_toff() + S] - (double)E->opt.t_offset_padding_ns;
    E->toff_hi[S] = E->h_state[E->sl.off_toff() + S] + (double)E->opt.t_offset_padding_ns;
  }
  if ((rc = plan_stream(E, B->ex, n, B->marg ? 5 : 4, 1))) return rc;
  if ((rc = mark_ready(E, B->ex))) return rc;
  E->imu.push_back(std::move(B));
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_add_global_velocity(clic_estimator* E, double timestamp, const double global_v[3], double vel_weight,
                                      int32_t* added) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!global_v) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (added) *added = 0;
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "global velocity factor in a marginalization estimator");
  // MeasuredTimeToNs(IMUSensor), trajectory_estimator.cpp:272-292: same host arithmetic as the IMU adder's bounds
  const int S = CLIC_SENSOR_IMU;
  const int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  const int64_t pad = E->opt.lock_t_offset[S] ? 0 : E->opt.t_offset_padding_ns;
  const int64_t t = (int64_t)(timestamp * 1e9);
  if (t - pad < t_min || t + pad >= t_max) return CLIC_OK;  // the reference returns false
  VelFactorDesc vf;
  vf.time_ns = t;
  for (int k = 0; k < 3; k++) vf.v[k] = global_v[k];
  vf.w = vel_weight;
  E->vel_factors.push_back(vf);
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {  // trajectory_estimator.cpp:653-656
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->h_state[E->sl.off_toff() + S] - (double)E->opt.t_offset_padding_ns;
    E->toff_hi[S] = E->h_state[E->sl.off_toff() + S] + (double)E->opt.t_offset_padding_ns;
  }
  E->layout_dirty = true;
  if (added) *added = 1;
  return CLIC_OK;
}

CLIC_API int clic_add_bias(clic_estimator* E, int32_t slot_i, int32_t slot_j, double dt, const double info_vec[6],
                           int32_t marg_this_factor, int32_t marg_all_bias) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (slot_i < 0 || slot_i >= E->n_bias || slot_j < 0 || slot_j >= E->n_bias)
    return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  BiasFac f;
  f.slot_i = slot_i;
  f.slot_j = slot_j;
  const double sqrt_dt = sqrt(dt);  // BiasFactor ctor, trajectory_value_factor.h:67-71
  for (int i = 0; i < 6; i++) f.sqrt_info[i] = info_vec[i] / sqrt_dt;
  f.marg = (E->opt.is_marg_state && marg_this_factor) ? 1 : 0;
  f.marg_all = marg_all_bias;
  E->bias_f.push_back(f);
  E->bias_desc_dirty = true;
  return CLIC_OK;
}

namespace {
// MeasuredTimeToNs(IMUSensor) (trajectory_estimator.cpp:272-292) with the host arithmetic of the IMU adder; false: dropped
bool imu_time_in_window(clic_estimator* E, double timestamp, int64_t* t_ns) {
  const int S = CLIC_SENSOR_IMU;
  const int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  const int64_t pad = E->opt.lock_t_offset[S] ? 0 : E->opt.t_offset_padding_ns;
  const int64_t t = (int64_t)(timestamp * 1e9);
  *t_ns = t;
  return !(t - pad < t_min || t + pad >= t_max);
}
void imu_offset_bounds(clic_estimator* E) {  // trajectory_estimator.cpp:355-361, 696-702
  const int S = CLIC_SENSOR_IMU;
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->h_state[E->sl.off_toff() + S] - (double)E->opt.t_offset_padding_ns;
    E->toff_hi[S] = E->h_state[E->sl.off_toff() + S] + (double)E->opt.t_offset_padding_ns;
  }
}
}  // namespace

CLIC_API int clic_add_pose(clic_estimator* E, int64_t n, const double* timestamp, const double* position,
                           const double* orientation, const double info_vec[6], int64_t* n_dropped) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "pose factor in a marginalization estimator");
  if (n > 0 && (!timestamp || !position || !orientation || !info_vec)) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  int64_t dropped = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t t;
    if (!imu_time_in_window(E, timestamp[i], &t)) { dropped++; continue; }
    SmallFactor f{};
    f.kind = kSmallPose;
    f.t_a = t;
    for (int k = 0; k < 3; k++) f.d[k] = position[3 * i + k];
    for (int k = 0; k < 4; k++) f.d[3 + k] = orientation[4 * i + k];
    for (int k = 0; k < 6; k++) f.d[7 + k] = info_vec[k];
    E->small_f.push_back(f);
    imu_offset_bounds(E);
  }
  if (n_dropped) *n_dropped = dropped;
  E->small_dirty = true;
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_add_local_velocity(clic_estimator* E, int64_t n, const double* timestamp, const double* local_v,
                                     double weight, int64_t* n_dropped) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "local velocity factor in a marginalization estimator");
  if (n > 0 && (!timestamp || !local_v)) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  int64_t dropped = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t t;
    if (!imu_time_in_window(E, timestamp[i], &t)) { dropped++; continue; }
    SmallFactor f{};
    f.kind = kSmallLocalVelocity;
    f.t_a = t;
    for (int k = 0; k < 3; k++) f.d[k] = local_v[3 * i + k];
    f.d[3] = weight;
    E->small_f.push_back(f);
    imu_offset_bounds(E);
  }
  if (n_dropped) *n_dropped = dropped;
  E->small_dirty = true;
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_add_relative_rotation(clic_estimator* E, double ta, double tb, const double S_BtoA[4],
                                        const double info_vec[3], int32_t* added) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "relative rotation factor in a marginalization estimator");
  if (!S_BtoA || !info_vec) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (added) *added = 0;
  int64_t t_a, t_b;
  if (!imu_time_in_window(E, ta, &t_a) || !imu_time_in_window(E, tb, &t_b)) return CLIC_OK;
  SmallFactor f{};
  f.kind = kSmallRelativeRotation;
  f.t_a = t_a;  // used as they are: the factor has no time-offset parameter (trajectory_estimator.cpp:586-607)
  f.t_b = t_b;
  for (int k = 0; k < 4; k++) f.d[k] = S_BtoA[k];
  for (int k = 0; k < 3; k++) f.d[4 + k] = info_vec[k];
  E->small_f.push_back(f);
  E->small_dirty = true;
  E->layout_dirty = true;
  if (added) *added = 1;
  return CLIC_OK;
}

CLIC_API int clic_add_image_depth(clic_estimator* E, const double p_i[3], const double p_j[3], const double S_CitoCj[4],
                                  const double p_CiinCj[3], int32_t landmark) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "image depth factor in a marginalization estimator");
  if (!p_i || !p_j || !S_CitoCj || !p_CiinCj) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  if (landmark < 0 || landmark >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "landmark index out of range");
  SmallFactor f{};
  f.kind = kSmallImageDepth;
  f.landmark = landmark;
  for (int k = 0; k < 3; k++) { f.d[k] = p_i[k]; f.d[3 + k] = p_j[k]; f.d[10 + k] = p_CiinCj[k]; }
  for (int k = 0; k < 4; k++) f.d[6 + k] = S_CitoCj[k];
  f.d[13] = E->opt.image_sqrt_info > 0.0 ? E->opt.image_sqrt_info : 450.0 / 1.5;  // ImageDepthFactor::sqrt_info
  f.loss_a = 1.0;                                                                  // CauchyLoss(1.0), ":842"
  E->small_f.push_back(f);
  E->lm_free[landmark] = 1;  // the inverse depth is a free parameter of the solve
  E->small_dirty = true;
  E->layout_dirty = true;
  return CLIC_OK;
}

namespace {
PreintData to_preint(const clic_preintegration* p) {
  PreintData d;
  d.sum_dt = p->sum_dt;
  d.gravity_norm = p->gravity_norm;
  for (int i = 0; i < 3; i++) {
    d.delta_p[i] = p->delta_p[i]; d.delta_v[i] = p->delta_v[i];
    d.linearized_ba[i] = p->linearized_ba[i]; d.linearized_bg[i] = p->linearized_bg[i];
  }
  for (int i = 0; i < 4; i++) d.delta_q[i] = p->delta_q[i];
  // StateOrder (integration_base.h:20-26): O_P 0, O_R 3, O_V 6, O_BA 9, O_BG 12
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      d.dp_dba[3 * r + c] = p->jacobian[(0 + r) * 15 + 9 + c];
      d.dp_dbg[3 * r + c] = p->jacobian[(0 + r) * 15 + 12 + c];
      d.dq_dbg[3 * r + c] = p->jacobian[(3 + r) * 15 + 12 + c];
      d.dv_dba[3 * r + c] = p->jacobian[(6 + r) * 15 + 9 + c];
      d.dv_dbg[3 * r + c] = p->jacobian[(6 + r) * 15 + 12 + c];
    }
  for (int i = 0; i < 225; i++) d.sqrt_info[i] = p->sqrt_info[i];
  return d;
}
}  // namespace

CLIC_API int clic_add_preintegration(clic_estimator* E, int64_t ta_ns, int64_t tb_ns, const clic_preintegration* pre,
                                     int32_t slot_i, int32_t slot_j, int32_t* added) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!pre) return fail(CLIC_ERR_BAD_ARGUMENT, "null pre-integration");
  if (E->opt.is_marg_state) return fail(CLIC_ERR_UNSUPPORTED, "pre-integration factor in a marginalization estimator");
  if (slot_i < 0 || slot_j < 0 || slot_i >= E->n_bias || slot_j >= E->n_bias)
    return fail(CLIC_ERR_BAD_ARGUMENT, "bias slot out of range");
  if (added) *added = 0;
  if (ta_ns < E->minTimeNs() || ta_ns > E->maxTimeNs() || tb_ns < E->minTimeNs() || tb_ns > E->maxTimeNs()) return CLIC_OK;
  SmallFactor f{};
  f.kind = kSmallPreIntegration;
  f.t_a = ta_ns;
  f.t_b = tb_ns;
  f.slot_i = slot_i;
  f.slot_j = slot_j;
  f.ext = (int)E->small_pre.size();
  E->small_pre.push_back(to_preint(pre));
  E->small_f.push_back(f);
  E->small_dirty = true;
  E->layout_dirty = true;
  if (added) *added = 1;
  return CLIC_OK;
}

CLIC_API int clic_evaluate_factor(clic_estimator* E, const clic_factor_desc* fd, int32_t* n_res, double* residual,
                                  int32_t* n_cols, double* jac, int64_t jac_capacity) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!fd || !n_res || !residual || !n_cols) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  // tangent widths of the parameter blocks after the knots, per kind (the reference's block order)
  static const int kWidths[10][4] = {{1, 0, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {1, 0, 0, 0}, {3, 3, 3, 3},
                                     {3, 1, 0, 0}, {1, 0, 0, 0}, {0, 0, 0, 0}, {3, 3, 0, 0}, {0, 0, 0, 0}};
  if (fd->kind < 0 || fd->kind > 9) return fail(CLIC_ERR_BAD_ARGUMENT, "unknown factor kind");
  if (fd->kind == CLIC_FACTOR_PREINTEGRATION && !fd->pre) return fail(CLIC_ERR_BAD_ARGUMENT, "null pre-integration");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  SmallFactor f{};
  f.kind = fd->kind;
  f.geo_type = fd->geo_type;
  f.t_a = fd->t_a_ns;
  f.t_b = fd->t_b_ns;
  for (int i = 0; i < 24; i++) f.d[i] = fd->d[i];
  SmallX X;
  for (int i = 0; i < 16; i++) X.x[i] = fd->x[i];
  int cols = 6 * E->K;
  for (int k = 0; k < 4; k++) {
    X.xcol[k] = kWidths[fd->kind][k] > 0 ? cols : -1;
    cols += kWidths[fd->kind][k];
  }
  DevBuf d_pre, d_out;
  if (fd->pre) {
    const PreintData pd = to_preint(fd->pre);
    CU(d_pre.ensure(sizeof(PreintData)));
    CU(cudaMemcpyAsync(d_pre.p, &pd, sizeof(PreintData), cudaMemcpyHostToDevice, E->stream));
    CU(cudaStreamSynchronize(E->stream));  // pd is a local
  }
  const size_t n_out = 20 + (size_t)kSmallMaxRows * cols;
  CU(d_out.ensure(n_out * sizeof(double)));
  CU(cudaMemsetAsync(d_out.p, 0, 4 * sizeof(double), E->stream));
  PassParams pp = pass_params(E, E->d_state.as<double>(), nullptr);
  const size_t smem = window_smem_doubles(E->K) * sizeof(double) + sizeof(SmallOut) + (size_t)E->K * sizeof(int) + 16;
  factor_jacobian_kernel<<<1, 128, smem, E->stream>>>(f, (const PreintData*)d_pre.as<PreintData>(), pp, ldq(E->S_CtoI),
                                                      ldv(E->p_CinI), X, cols, d_out.as<double>());
  CU(cudaGetLastError());
  E->launches++;
  std::vector<double> h(n_out);
  CU(cudaMemcpyAsync(h.data(), d_out.p, n_out * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  if (h[0] == 0.0) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_evaluate_factor: a measurement time is outside the window");
  const int nr = (int)h[1];
  *n_res = nr;
  *n_cols = cols;
  for (int r = 0; r < nr; r++) residual[r] = h[4 + r];
  if (jac) {
    if ((int64_t)nr * cols > jac_capacity) return fail(CLIC_ERR_BAD_ARGUMENT, "Jacobian buffer too small");
    std::memcpy(jac, h.data() + 20, (size_t)nr * cols * sizeof(double));
  }
  return CLIC_OK;
}

CLIC_API int clic_add_gravity(clic_estimator* E, const double gravity[3], const double info_vec[3],
                              int32_t marg_this_factor) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!gravity || !info_vec) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_gravity: null argument");
  clic_estimator::GravFac f;
  for (int i = 0; i < 3; i++) { f.g0[i] = gravity[i]; f.sqrt_info[i] = info_vec[i]; }
  f.marg = (E->opt.is_marg_state && marg_this_factor) ? 1 : 0;
  E->grav_f.push_back(f);
  return CLIC_OK;
}

CLIC_API int clic_add_image(clic_estimator* E, int64_t n, const double* t_i, const double* p_i, const double* t_j,
                            const double* p_j, const int32_t* landmark, int32_t fixed_depth,
                            int32_t marg_this_feature, int64_t* n_dropped) {
  ScopedTrace trace("clic_add_image");
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_dropped) *n_dropped = 0;
  if (n <= 0) return CLIC_OK;
  if (!fixed_depth && E->opt.is_marg_state)
    return fail(CLIC_ERR_UNSUPPORTED, "free inverse depths in a marginalization estimator are not built on the GPU");
  for (int64_t i = 0; i < n; i++)
    if (landmark[i] < 0 || landmark[i] >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "landmark index out of range");

  CU(cudaSetDevice(E->device));
  cudaStream_t cs = E->copy_stream;
  AllocScope alloc_scope(cs);
  std::unique_ptr<ImageBatch> B(new ImageBatch);
  const int S = CLIC_SENSOR_CAMERA;
  const int64_t t_min = (int64_t)(E->minTime(S) * 1e9), t_max = (int64_t)(E->maxTime(S) * 1e9);
  const int64_t pad = E->opt.lock_t_offset[S] ? 0 : E->opt.t_offset_padding_ns;
  if (!fixed_depth)  // trajectory_estimator.cpp:789-793: the block of a KEPT factor stays variable unless fixed_depth
    for (int64_t i = 0; i < n; i++) {  // same window test as prep_image_kernel (a dropped factor adds no block)
      const int64_t a = (int64_t)(t_i[i] * 1e9), b = (int64_t)(t_j[i] * 1e9);
      if (!(a - pad < t_min || a + pad >= t_max) && !(b - pad < t_min || b + pad >= t_max)) E->lm_free[landmark[i]] = 1;
    }
  // Order by the (s_i, s_j) knot-interval key (stable counting sort, O(n)), every key's factors starting on a slab
  // boundary (padding = dropped factors): a 16-factor slab then holds ONE key — a slab that straddles k keys costs k
  // evaluation rounds — and a warp's accumulators are flushed rarely.  The key here only orders the inputs (host
  // arithmetic at the current offset); the kernels recompute the exact indices.
  const int64_t n_in = n;
  if (n_in > INT32_MAX / 2) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_image: batch too large");
  {
    const int nk = std::max(1, E->K - 3);
    const int64_t toff = (int64_t)E->h_state[E->sl.off_toff() + S];
    auto interval = [&](double t) {
      const int64_t q = ((int64_t)(t * 1e9) + toff - E->t0_ns) / E->dt_ns;
      return (int)std::min<int64_t>(nk - 1, std::max<int64_t>(0, q));
    };
    std::vector<int> key(n_in);
    std::vector<int64_t> start((size_t)nk * nk + 1, 0);
    for (int64_t i = 0; i < n_in; i++) {
      key[i] = interval(t_i[i]) * nk + interval(t_j[i]);
      start[key[i] + 1]++;
    }
    for (size_t k = 0; k < (size_t)nk * nk; k++)
      start[k + 1] = start[k] + (start[k + 1] + kImageFactorsPerSlab - 1) / kImageFactorsPerSlab * kImageFactorsPerSlab;
    B->gather.assign((size_t)start[(size_t)nk * nk], -1);
    for (int64_t i = 0; i < n_in; i++) B->gather[start[key[i]]++] = (int32_t)i;
  }
  n = (int64_t)B->gather.size();  // stored factors, padding included
  B->n_input = n_in;
  // Permuted copies go through a page-locked staging arena (process-wide, grown on demand, guarded by the event of
  // its last upload): a cudaMemcpyAsync from pageable vectors would block this thread until every earlier copy on
  // the stream (the LiDAR batch) has drained.
  static std::mutex arena_mu;
  static void* arena = nullptr;
  static size_t arena_bytes = 0;
  static cudaEvent_t arena_done = nullptr;
  std::lock_guard<std::mutex> arena_lock(arena_mu);
  const size_t need = (size_t)n * (8 * sizeof(double) + 3 * sizeof(int32_t)) + kMaxPosePacks * sizeof(int64_t);
  if (arena_done) CU(cudaEventSynchronize(arena_done));
  trace.lap("arena wait");
  if (need > arena_bytes) {
    if (arena) cudaFreeHost(arena);
    arena = nullptr;
    arena_bytes = 0;
    CU(cudaHostAlloc(&arena, need + need / 2, cudaHostAllocDefault));
    arena_bytes = need + need / 2;
  }
  if (!arena_done) CU(cudaEventCreateWithFlags(&arena_done, cudaEventDisableTiming));
  double* s_ti = reinterpret_cast<double*>(arena);
  double* s_tj = s_ti + n;
  double* s_pi = s_tj + n;
  double* s_pj = s_pi + 3 * n;
  int64_t* s_pack_t = reinterpret_cast<int64_t*>(s_pj + 3 * n);
  int32_t* s_lm = reinterpret_cast<int32_t*>(s_pack_t + kMaxPosePacks);
  int32_t* s_idi = s_lm + n;
  int32_t* s_idj = s_idi + n;
  constexpr double kPadTime = -1.0e9;  // seconds: far before any window, exactly representable in ns
  for (int64_t k = 0; k < n; k++) {
    const int64_t i = B->gather[k];
    if (i < 0) {
      s_ti[k] = s_tj[k] = kPadTime;
      for (int c = 0; c < 3; c++) { s_pi[3 * k + c] = 0.0; s_pj[3 * k + c] = 0.0; }
      s_lm[k] = 0;
      s_idi[k] = s_idj[k] = 0;
      continue;
    }
    s_ti[k] = t_i[i];
    s_tj[k] = t_j[i];
    for (int c = 0; c < 3; c++) { s_pi[3 * k + c] = p_i[3 * i + c]; s_pj[3 * k + c] = p_j[3 * i + c]; }
    s_lm[k] = landmark[i];
  }
  {  // CSR landmark -> stored positions (ascending), for the per-landmark gather of the inverse-depth columns
    std::vector<int32_t>& start = B->h_lm_start;
    std::vector<int32_t>& obs = B->h_lm_obs;
    start.assign((size_t)E->n_lm + 1, 0);
    obs.assign((size_t)n_in, 0);
    for (int64_t k = 0; k < n; k++) if (B->gather[k] >= 0) start[s_lm[k] + 1]++;
    for (int l = 0; l < E->n_lm; l++) start[l + 1] += start[l];
    std::vector<int32_t> fill(start.begin(), start.end() - 1);
    for (int64_t k = 0; k < n; k++) if (B->gather[k] >= 0) obs[fill[s_lm[k]]++] = (int32_t)k;
    int rc2;
    if ((rc2 = to_device(E, B->lm_start, start.data(), start.size())) || (rc2 = to_device(E, B->lm_obs, obs.data(), obs.size())))
      return rc2;
  }
  trace.lap("sort+stage");
  // Distinct (timestamp, role) pairs -> pose packs (tiny open-addressing table; abandoned beyond kMaxPosePacks,
  // e.g. rolling-shutter data with per-feature times, which then takes the per-factor chain kernel).
  int n_pack_i = 0, n_pack_j = 0;
  bool use_packs = E->packs_fit && !getenv("CLIC_NO_POSE_PACKS");
  {
    constexpr int kTab = 512;
    int64_t tab_t[2][kTab] = {};
    int tab_id[2][kTab];
    for (int r = 0; r < 2; r++) for (int q = 0; q < kTab; q++) tab_id[r][q] = -1;
    int64_t pack_t_i[kMaxPosePacks], pack_t_j[kMaxPosePacks];
    auto lookup = [&](int r, int64_t t, int& count, int64_t* list) -> int {
      unsigned h = (unsigned)(((uint64_t)t * 0x9E3779B97F4A7C15ull) >> 55) & (kTab - 1);
      while (tab_id[r][h] >= 0 && tab_t[r][h] != t) h = (h + 1) & (kTab - 1);
      if (tab_id[r][h] < 0) {
        if (n_pack_i + n_pack_j >= kMaxPosePacks) return -1;
        tab_t[r][h] = t;
        tab_id[r][h] = count;
        list[count++] = t;
      }
      return tab_id[r][h];
    };
    for (int64_t k = 0; k < n && use_packs; k++) {
      if (B->gather[k] < 0) continue;
      const int a = lookup(0, (int64_t)(s_ti[k] * 1e9), n_pack_i, pack_t_i);
      const int b = a < 0 ? -1 : lookup(1, (int64_t)(s_tj[k] * 1e9), n_pack_j, pack_t_j);
      if (a < 0 || b < 0) { use_packs = false; break; }
      s_idi[k] = a;
      s_idj[k] = b;
    }
    if (use_packs) {
      for (int64_t k = 0; k < n; k++) if (B->gather[k] >= 0) s_idj[k] += n_pack_i;
      for (int q = 0; q < n_pack_i; q++) s_pack_t[q] = pack_t_i[q];
      for (int q = 0; q < n_pack_j; q++) s_pack_t[n_pack_i + q] = pack_t_j[q];
    }
  }
  trace.lap("packs");
  DevBuf &r_ti = B->raw[0], &r_tj = B->raw[1], &r_pi = B->raw[2], &r_pj = B->raw[3];
  int rc;
  if (use_packs && ((rc = to_device(E, B->id_i, s_idi, n)) || (rc = to_device(E, B->id_j, s_idj, n)) ||
                    (rc = to_device(E, B->pack_t, s_pack_t, n_pack_i + n_pack_j))))
    return rc;
  if ((rc = to_device(E, r_ti, s_ti, n)) || (rc = to_device(E, r_tj, s_tj, n)) ||
      (rc = to_device(E, r_pi, s_pi, 3 * n)) || (rc = to_device(E, r_pj, s_pj, 3 * n)) ||
      (rc = to_device(E, B->lm, s_lm, n)))
    return rc;
  CU(cudaEventRecord(arena_done, cs));
  trace.lap("uploads");
  CU(B->ti.ensure(n * sizeof(int64_t)));
  CU(B->tj.ensure(n * sizeof(int64_t)));
  for (int k = 0; k < 3; k++) CU(B->pi[k].ensure(n * sizeof(double)));
  for (int k = 0; k < 2; k++) CU(B->pj[k].ensure(n * sizeof(double)));
  trace.lap("ensures");
  CU(cudaMemsetAsync(E->d_counter.p, 0, sizeof(unsigned long long), cs));
  prep_image_kernel<<<(unsigned)((n + 255) / 256), 256, 0, cs>>>(
      n, r_ti.as<double>(), r_tj.as<double>(), r_pi.as<double>(), r_pj.as<double>(), t_min, t_max, pad,
      B->ti.as<int64_t>(), B->tj.as<int64_t>(), B->pi[0].as<double>(), B->pi[1].as<double>(), B->pi[2].as<double>(),
      B->pj[0].as<double>(), B->pj[1].as<double>(), E->d_counter.as<unsigned long long>());
  E->launches++;
  if (n_dropped) {
    unsigned long long dropped = 0;
    CU(cudaMemcpyAsync(&dropped, E->d_counter.p, sizeof(dropped), cudaMemcpyDeviceToHost, cs));
    CU(cudaStreamSynchronize(cs));
    *n_dropped = (int64_t)dropped - (n - n_in);  // the padding counted as dropped
  }
  B->st.n = n;
  B->st.ti_ns = B->ti.as<int64_t>();
  B->st.tj_ns = B->tj.as<int64_t>();
  for (int k = 0; k < 3; k++) B->st.pi[k] = B->pi[k].as<double>();
  for (int k = 0; k < 2; k++) B->st.pj[k] = B->pj[k].as<double>();
  B->st.landmark = B->lm.as<int32_t>();
  B->st.id_i = B->id_i.as<int32_t>();
  B->st.id_j = B->id_j.as<int32_t>();
  B->st.pack_t_ns = B->pack_t.as<int64_t>();
  B->st.n_pack_i = use_packs ? n_pack_i : 0;
  B->st.n_packs = use_packs ? n_pack_i + n_pack_j : 0;
  B->st.loss_a = marg_this_feature ? 1.0 : 2.0;  // trajectory_estimator.cpp:806-810
  B->marg = (E->opt.is_marg_state && marg_this_feature) ? 1 : 0;
  B->lm_used.assign(std::max(0, E->n_lm), 0);
  for (int64_t i = 0; i < n_in; i++) {  // same window test as prep_image_kernel
    const int64_t a = (int64_t)(t_i[i] * 1e9), b = (int64_t)(t_j[i] * 1e9);
    if (!(a - pad < t_min || a + pad >= t_max) && !(b - pad < t_min || b + pad >= t_max)) B->lm_used[landmark[i]] = 1;
  }
  if (!E->opt.lock_t_offset[S] && !E->bounds_toff[S]) {  // trajectory_estimator.cpp:795-800
    E->bounds_toff[S] = true;
    E->toff_lo[S] = E->h_state[E->sl.off_toff() + S] - (double)E->opt.t_offset_padding_ns;
    E->toff_hi[S] = E->h_state[E->sl.off_toff() + S] + (double)E->opt.t_offset_padding_ns;
  }
  trace.lap("prep");
  if ((rc = plan_stream(E, B->ex, n, 7, 1, true, 0, kWarpsPerCta, kImageFactorsPerSlab))) return rc;
  trace.lap("plan");
  if ((rc = mark_ready(E, B->ex))) return rc;
  E->image.push_back(std::move(B));
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_add_prior(clic_estimator* E, const clic_prior* prior, int32_t n_drop, const int32_t* drop_blocks) {
  if (!ok_handle(E) || !prior) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  std::unique_ptr<clic_estimator::PriorRef> R(new clic_estimator::PriorRef);
  R->p = prior;
  if (n_drop > 0) R->drop.assign(drop_blocks, drop_blocks + n_drop);
  if (E->opt.is_marg_state && n_drop <= 0) return CLIC_OK;  // trajectory_manager.cpp:367
  const int nb = (int)prior->blocks.size();
  std::vector<int> meta(4 * nb);  // per block: state offset, size, tangent idx, column (-1 const) [column filled per pass]
  std::vector<double> x0(4 * nb);
  for (int i = 0; i < nb; i++) {
    const auto& b = prior->blocks[i];
    int off = -1;
    switch (b.kind) {
      case CLIC_BLOCK_KNOT_ROT: {
        int k = b.id - E->knot_id0;
        if (k < 0 || k >= E->K) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to a knot outside the window");
        off = E->sl.off_q() + 4 * k;
      } break;
      case CLIC_BLOCK_KNOT_POS: {
        int k = b.id - E->knot_id0;
        if (k < 0 || k >= E->K) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to a knot outside the window");
        off = E->sl.off_p() + 3 * k;
      } break;
      case CLIC_BLOCK_BIAS_GYRO:
      case CLIC_BLOCK_BIAS_ACCEL: {
        int s = b.id - E->bias_id0;
        if (s < 0 || s >= E->n_bias) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to a bias slot outside the window");
        off = E->sl.off_bias() + 6 * s + (b.kind == CLIC_BLOCK_BIAS_ACCEL ? 3 : 0);
      } break;
      case CLIC_BLOCK_GRAVITY: off = E->sl.off_grav(); break;
      case CLIC_BLOCK_T_OFFSET: off = E->sl.off_toff() + b.id; break;
      case CLIC_BLOCK_INV_DEPTH:
        if (b.id < 0 || b.id >= E->n_lm) return fail(CLIC_ERR_BAD_ARGUMENT, "prior refers to an unknown landmark");
        off = E->sl.off_invd() + b.id;
        break;
      default: return fail(CLIC_ERR_BAD_ARGUMENT, "unknown prior block kind");
    }
    meta[4 * i + 0] = off;
    meta[4 * i + 1] = b.size;
    meta[4 * i + 2] = b.idx;
    meta[4 * i + 3] = -1;
    for (int k = 0; k < 4; k++) x0[4 * i + k] = b.x0[k];
  }
  int rc;
  if ((rc = to_device(E, R->J0, prior->J0.data(), prior->J0.size())) ||
      (rc = to_device(E, R->r0, prior->r0.data(), prior->r0.size())) || (rc = to_device(E, R->x0, x0.data(), x0.size())) ||
      (rc = to_device(E, R->meta, meta.data(), meta.size())))
    return rc;
  R->h_meta = meta;
  CU(R->A.ensure((size_t)prior->n * prior->n * sizeof(double)));
  {
    dim3 blk(16, 16), grd((prior->n + 15) / 16, (prior->n + 15) / 16);
    prior_gram_kernel<<<grd, blk, 0, E->stream>>>(R->J0.as<double>(), prior->n, R->A.as<double>());
    E->launches++;
  }
  CU(cudaStreamSynchronize(E->stream));
  E->priors.push_back(std::move(R));
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_get_layout(clic_estimator* E, clic_layout* out) {
  if (!ok_handle(E) || !out) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  *out = compute_layout(E);
  return CLIC_OK;
}

CLIC_API int clic_evaluate(clic_estimator* E, double* H, double* b, double* cost, double* fixed_cost) {
  ScopedTrace trace("clic_evaluate");
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  rc = fill_prior_columns(E);
  if (rc) return rc;
  const bool jac = H || b;
  rc = run_pass(E, E->d_state.as<double>(), jac, 0, nullptr);
  if (rc) return rc;
  const int D = E->L.D;
  double c2[2];
  if (H && D > 0) CU(cudaMemcpyAsync(H, E->H_glb(), (size_t)D * D * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  if (b && D > 0) CU(cudaMemcpyAsync(b, E->b_glb(), (size_t)D * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(c2, E->cost_glb(), sizeof(c2), cudaMemcpyDeviceToHost, E->stream));
  int flags[2] = {0, 0};
  CU(cudaMemcpyAsync(flags, E->d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  if (flags[1]) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_loam: geometry array missing for a record type that is present");
  if (int prc = check_p2p(E)) return prc;
  if (cost) *cost = c2[0];
  if (fixed_cost) *fixed_cost = c2[1] + E->grav_fixed_cost();
  return CLIC_OK;
}

CLIC_API int clic_residual_summary(clic_estimator* E, double cost[16], int64_t count[16]) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  if ((rc = fill_prior_columns(E))) return rc;
  for (int i = 0; i < 16; i++) { if (cost) cost[i] = 0.0; if (count) count[i] = 0; }
  DevBuf tmp;
  CU(tmp.ensure(64 * sizeof(double)));  // [0, 32): per type (cost, fixed cost); [32, 64): per stream
  CU(cudaMemsetAsync(tmp.p, 0, 64 * sizeof(double), E->stream));
  double* d_c = tmp.as<double>();
  const double* state = E->d_state.as<double>();
  PassParams pp = pass_params(E, state, nullptr);
  // one residual-only launch per stream + its cost sum (the cost CTA of reduce_all_kernel on that stream's job alone);
  // jobs are in ensure_layout's order: LiDAR batches, IMU batches, visual batches (marginalization batches skipped)
  int j = 0;
  auto sum_job = [&](int type) -> int {
    CU(launch_pdl(reduce_all_kernel, 1, 256, (size_t)2 * E->max_slots * sizeof(int), E->stream,
                  E->d_jobs.as<ReduceJob>() + j, 1, 0, E->max_slots, d_c + 32 + 2 * j, (const int*)nullptr));
    (void)type;
    if (++j > 16) return fail(CLIC_ERR_UNSUPPORTED, "clic_residual_summary: more than 16 factor streams");
    return CLIC_OK;
  };
  std::vector<int> job_type;
  for (auto& b : E->loam) {
    if (b->marg) continue;
    if ((rc = wait_ready(E, b->ex))) return rc;
    const size_t smem = loam_smem_bytes(E->K);
    if (b->st.geo_kind == 1) CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 1>, b->ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, b->ex.slabs_per_warp, b->ex.rb()));
    else if (b->st.geo_kind == 2) CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 2>, b->ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, b->ex.slabs_per_warp, b->ex.rb()));
    else CU(launch_pdl(loam_kernel<false, kLoamWarps, kLoamMinB, 0>, b->ex.grid, kLoamWarps * 32, smem, E->stream, b->st, pp, b->ex.slabs_per_warp, b->ex.rb()));
    job_type.push_back(4);  // RType_LiDAR
    if (count) count[4] += b->n_input;
    if ((rc = sum_job(4))) return rc;
  }
  for (auto& b : E->imu) {
    if (b->marg) continue;
    if ((rc = wait_ready(E, b->ex))) return rc;
    b->st.toff_free = E->L.col_t_offset[CLIC_SENSOR_IMU] >= 0 ? 1 : 0;
    CU(launch_pdl(imu_kernel<false, false>, b->ex.grid, kThreads, imu_smem_bytes(E->K), E->stream, b->st, pp,
                  b->ex.slabs_per_warp, b->ex.rb()));
    job_type.push_back(1);  // RType_IMU
    if (count) count[1] += b->n_input;
    if ((rc = sum_job(1))) return rc;
  }
  for (auto& b : E->image) {
    if (b->marg) continue;
    if ((rc = wait_ready(E, b->ex))) return rc;
    const ImageShared ish = image_shared(E, b.get());
    image_pass_fields(E, b->st, false);
    if (b->st.n_packs > 0)
      CU(launch_pdl(image_kernel<false, true>, b->ex.grid, kThreads, image_smem_bytes(E->K, b->st.n_packs), E->stream, b->st,
                    pp, b->ex.slabs_per_warp, b->ex.rb(), ish));
    else
      CU(launch_pdl(image_kernel<false, false>, b->ex.grid, kThreads, factor_smem_bytes(E->K, 7), E->stream, b->st, pp,
                    b->ex.slabs_per_warp, b->ex.rb(), ish));
    job_type.push_back(11);  // RType_Image
    if (count) count[11] += b->n_input;
    if ((rc = sum_job(11))) return rc;
  }
  
  double* H_scratch = E->nrm_loc_set(1);  // the kernels below take H, b pointers but do not touch them without jac
  double* b_scratch = H_scratch + E->off_b();
  const int D = std::max(1, E->L.D);
  for (auto& d : solve_bias_descs(E))
    CU(launch_pdl(bias_factor_kernel, 1, 32, 0, E->stream, d, state, E->sl, H_scratch, b_scratch, D, d_c + 2 * 2, 0,
                  (const int*)nullptr));
  if (count) for (auto& bf : E->bias_f) if (!bf.marg) count[2]++;
  if (!E->opt.is_marg_state)
    for (auto& pr : E->priors) {
      if ((rc = launch_prior(E, pr.get(), state, d_c + 2 * 13, false, nullptr, H_scratch, b_scratch))) return rc;
      if (count) count[13]++;
    }
  for (auto& vf : E->vel_factors) {
    CU(launch_pdl(velocity_factor_kernel, 1, 64, 0, E->stream, vf, state, E->sl, (long long)E->t0_ns, (long long)E->dt_ns,
                  E->K, (const int*)E->d_knot_col.as<int>(), E->L.col_t_offset[CLIC_SENSOR_IMU], H_scratch, b_scratch, D,
                  d_c + 2 * 8, 0, (const int*)nullptr));
    if (count) count[8]++;
  }
  if ((rc = launch_small_factors(E, state, H_scratch, b_scratch, d_c + 2 * 14, false, nullptr))) return rc;
  if (count) count[14] += (int64_t)E->small_f.size();
  double h[64];
  CU(cudaMemcpyAsync(h, tmp.p, sizeof(h), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  CU(cudaGetLastError());
  if (cost) {
    for (int t : {2, 8, 13, 14}) cost[t] = h[2 * t] + h[2 * t + 1];
    for (int q = 0; q < j; q++) cost[job_type[q]] += h[32 + 2 * q] + h[32 + 2 * q + 1];
    cost[3] = E->grav_fixed_cost();
  }
  if (count) for (auto& gf : E->grav_f) if (!gf.marg) count[3]++;
  return CLIC_OK;
}

CLIC_API int clic_get_indices(clic_estimator* E, int32_t stream, int64_t capacity, int32_t* s_out, double* u_out,
                              int64_t* n_out) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (stream < 0 || stream > 3) return fail(CLIC_ERR_BAD_ARGUMENT, "stream");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int64_t total = 0;
  if (stream == 0) for (auto& b : E->loam) total += b->n_input;
  if (stream == 1) for (auto& b : E->imu) total += b->n_input;
  if (stream >= 2) for (auto& b : E->image) total += b->n_input;
  if (n_out) *n_out = total;
  if (capacity <= 0 || total == 0) return CLIC_OK;
  if (int jrc = join_adds(E)) return jrc;
  // current offset is read from the device state
  double toff_d[3];
  CU(cudaMemcpyAsync(toff_d, E->d_state.as<double>() + E->sl.off_toff(), sizeof(toff_d), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  DevBuf ds, du;
  int64_t done = 0;
  // gather: the batch is stored key-aligned (stored position -> input record, -1 padding): report in add order
  auto run = [&](int64_t n, const int64_t* t_ns, int64_t toff, const std::vector<int32_t>* gather = nullptr,
                 int64_t n_input = 0) -> int {
    if (gather && !gather->empty()) {
      if (n_input > capacity - done) return fail(CLIC_ERR_BAD_ARGUMENT, "clic_get_indices: capacity too small");
      CU(ds.ensure(n * sizeof(int32_t)));
      CU(du.ensure(n * sizeof(double)));
      index_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, E->stream>>>(n, t_ns, toff, E->t0_ns, E->dt_ns, E->K,
                                                                           ds.as<int32_t>(), du.as<double>());
      E->launches++;
      std::vector<int32_t> ts((size_t)n);
      std::vector<double> tu((size_t)n);
      CU(cudaMemcpyAsync(ts.data(), ds.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
      CU(cudaMemcpyAsync(tu.data(), du.p, n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
      CU(cudaStreamSynchronize(E->stream));
      for (int64_t o = 0; o < n; o++) {
        const int32_t i = (*gather)[(size_t)o];
        if (i >= 0) { s_out[done + i] = ts[(size_t)o]; u_out[done + i] = tu[(size_t)o]; }
      }
      done += n_input;
      return CLIC_OK;
    }
    int64_t m = std::min(n, capacity - done);
    if (m <= 0) return CLIC_OK;
    CU(ds.ensure(m * sizeof(int32_t)));
    CU(du.ensure(m * sizeof(double)));
    index_probe_kernel<<<(unsigned)((m + 255) / 256), 256, 0, E->stream>>>(m, t_ns, toff, E->t0_ns, E->dt_ns, E->K,
                                                                         ds.as<int32_t>(), du.as<double>());
    E->launches++;
    CU(cudaMemcpyAsync(s_out + done, ds.p, m * sizeof(int32_t), cudaMemcpyDeviceToHost, E->stream));
    CU(cudaMemcpyAsync(u_out + done, du.p, m * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    done += m;
    return CLIC_OK;
  };
  int rc = CLIC_OK;
  if (stream == 0)
    for (auto& b : E->loam) if ((rc = run(b->st.n, b->st.t_ns, 0, &b->gather, b->n_input))) return rc;
  if (stream == 1)
    for (auto& b : E->imu)
      if ((rc = run(b->st.n, b->st.t_ns, (int64_t)toff_d[CLIC_SENSOR_IMU], &b->gather, b->n_input))) return rc;
  if (stream >= 2)  // the stream is stored sorted by (s_i, s_j) with padding: reported in add order through the gather map
    for (auto& b : E->image)
      if ((rc = run(b->st.n, stream == 2 ? b->st.ti_ns : b->st.tj_ns, (int64_t)toff_d[CLIC_SENSOR_CAMERA], &b->gather,
                    b->n_input)))
        return rc;
  return CLIC_OK;
}

CLIC_API int clic_solve(clic_estimator* E, int32_t max_iterations, clic_summary* summary) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!E->opt.lock_g) return fail(CLIC_ERR_UNSUPPORTED, "gravity must be locked (lock_g = 1)");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  if ((rc = fill_prior_columns(E))) return rc;
  return lm_solve(E, max_iterations, summary);
}

CLIC_API int clic_read_state(clic_estimator* E, double* knot_q, double* knot_p, double* bias, double* t_offset_ns,
                             double* inv_depth) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  std::vector<double> h(E->sl.size());
  CU(cudaMemcpyAsync(h.data(), E->d_state.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  E->h_state = h;
  if (knot_q) std::copy(h.begin() + E->sl.off_q(), h.begin() + E->sl.off_q() + 4 * E->K, knot_q);
  if (knot_p) std::copy(h.begin() + E->sl.off_p(), h.begin() + E->sl.off_p() + 3 * E->K, knot_p);
  if (bias) std::copy(h.begin() + E->sl.off_bias(), h.begin() + E->sl.off_bias() + 6 * E->n_bias, bias);
  if (t_offset_ns) std::copy(h.begin() + E->sl.off_toff(), h.begin() + E->sl.off_toff() + 3, t_offset_ns);
  if (inv_depth) std::copy(h.begin() + E->sl.off_invd(), h.begin() + E->sl.off_invd() + E->n_lm, inv_depth);
  return CLIC_OK;
}

CLIC_API int clic_write_state(clic_estimator* E, const double* knot_q, const double* knot_p, const double* bias,
                              const double* t_offset_ns, const double* inv_depth) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  auto& h = E->h_state;
  if (knot_q) std::copy(knot_q, knot_q + 4 * E->K, h.begin() + E->sl.off_q());
  if (knot_p) std::copy(knot_p, knot_p + 3 * E->K, h.begin() + E->sl.off_p());
  if (bias) std::copy(bias, bias + 6 * E->n_bias, h.begin() + E->sl.off_bias());
  if (t_offset_ns) std::copy(t_offset_ns, t_offset_ns + 3, h.begin() + E->sl.off_toff());
  if (inv_depth) std::copy(inv_depth, inv_depth + E->n_lm, h.begin() + E->sl.off_invd());
  int rc = upload_state(E);
  if (rc) return rc;
  CU(cudaStreamSynchronize(E->stream));
  return CLIC_OK;
}

namespace {
PoseQuery make_pose_query(const clic_estimator* E, int64_t n, const double* d_t, const double S[4], const double p[3],
                          double t_offset_ns) {
  PoseQuery q;
  q.n = n;
  q.t_s = d_t;
  q.S_StoI = ldq(S);
  q.p_SinI = ldv(p);
  q.t_offset_ns = t_offset_ns;
  q.t_min_ns = (double)E->minTimeNs();
  q.t_max_ns = (double)E->maxTimeNs();
  return q;
}
int query_grid(const clic_estimator* E, int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)E->n_sm * 8, (n + 255) / 256));
}
}  // namespace

CLIC_API int clic_query_poses(clic_estimator* E, int64_t n, const double* t_s, const double S_StoI[4],
                              const double p_SinI[3], double t_offset_ns, double* q_out, double* p_out,
                              int64_t* n_invalid) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_invalid) *n_invalid = 0;
  if (n <= 0) return CLIC_OK;
  if (!t_s || !S_StoI || !p_SinI || !q_out || !p_out) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  DevBuf d_t, d_q, d_p;
  int rc;
  if ((rc = to_device(E, d_t, t_s, n))) return rc;
  CU(d_q.ensure((size_t)4 * n * sizeof(double)));
  CU(d_p.ensure((size_t)3 * n * sizeof(double)));
  CU(cudaMemsetAsync(E->d_counter.as<unsigned long long>() + 1, 0, sizeof(unsigned long long), E->stream));
  PassParams pp = pass_params(E, E->d_state.as<double>(), nullptr);
  const PoseQuery q = make_pose_query(E, n, d_t.as<double>(), S_StoI, p_SinI, t_offset_ns);
  pose_query_kernel<<<query_grid(E, n), 256, window_smem_doubles(E->K) * sizeof(double), E->stream>>>(
      pp, q, d_q.as<double>(), d_p.as<double>(), E->d_counter.as<unsigned long long>() + 1);
  E->launches++;
  unsigned long long bad = 0;
  CU(cudaMemcpyAsync(q_out, d_q.p, (size_t)4 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(p_out, d_p.p, (size_t)3 * n * sizeof(double), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(&bad, E->d_counter.as<unsigned long long>() + 1, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  CU(cudaGetLastError());
  if (n_invalid) *n_invalid = (int64_t)bad;
  return CLIC_OK;
}

CLIC_API int clic_undistort_scan(clic_estimator* E, int64_t n, const double* t_s, const float* xyz_in,
                                 const double S_LtoI[4], const double p_LinI[3], double t_offset_ns, int32_t in_global,
                                 double target_t_s, float* xyz_out, int64_t* n_invalid) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_invalid) *n_invalid = 0;
  if (n <= 0) return CLIC_OK;
  if (!t_s || !xyz_in || !S_LtoI || !p_LinI || !xyz_out) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  DevBuf d_t, d_in, d_out, d_target, d_tt;
  int rc;
  if ((rc = to_device(E, d_t, t_s, n)) || (rc = to_device(E, d_in, xyz_in, (size_t)3 * n))) return rc;
  CU(d_out.ensure((size_t)3 * n * sizeof(float)));
  PassParams pp = pass_params(E, E->d_state.as<double>(), nullptr);
  const size_t smem = window_smem_doubles(E->K) * sizeof(double);
  const double* target = nullptr;
  unsigned long long bad = 0;
  if (!in_global) {  // GetLidarPose(target_timestamp) on the device, then its inverse inside the kernel
    if ((rc = to_device(E, d_tt, &target_t_s, 1))) return rc;
    CU(d_target.ensure(7 * sizeof(double)));
    CU(cudaMemsetAsync(E->d_counter.as<unsigned long long>() + 1, 0, sizeof(unsigned long long), E->stream));
    const PoseQuery qt = make_pose_query(E, 1, d_tt.as<double>(), S_LtoI, p_LinI, t_offset_ns);
    pose_query_kernel<<<1, 256, smem, E->stream>>>(pp, qt, d_target.as<double>(), d_target.as<double>() + 4,
                                                   E->d_counter.as<unsigned long long>() + 1);
    E->launches++;
    CU(cudaMemcpyAsync(&bad, E->d_counter.as<unsigned long long>() + 1, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    if (bad) return fail(CLIC_ERR_BAD_ARGUMENT, "target time outside the trajectory");
    target = d_target.as<double>();
  }
  CU(cudaMemsetAsync(E->d_counter.as<unsigned long long>() + 1, 0, sizeof(unsigned long long), E->stream));
  const PoseQuery q = make_pose_query(E, n, d_t.as<double>(), S_LtoI, p_LinI, t_offset_ns);
  undistort_kernel<<<query_grid(E, n), 256, smem, E->stream>>>(pp, q, d_in.as<float>(), target, d_out.as<float>(),
                                                              E->d_counter.as<unsigned long long>() + 1);
  E->launches++;
  CU(cudaMemcpyAsync(xyz_out, d_out.p, (size_t)3 * n * sizeof(float), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaMemcpyAsync(&bad, E->d_counter.as<unsigned long long>() + 1, sizeof(bad), cudaMemcpyDeviceToHost, E->stream));
  CU(cudaStreamSynchronize(E->stream));
  CU(cudaGetLastError());
  if (n_invalid) *n_invalid = (int64_t)bad;
  return CLIC_OK;
}

CLIC_API int clic_marginalize(clic_estimator* E, clic_prior** out) {
  if (!ok_handle(E) || !out) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  return marginalize_impl(E, out);
}

CLIC_API int clic_prior_create(int32_t n, const double* J0, const double* r0, int32_t n_blocks, const int32_t* kind,
                               const int32_t* id, const double* x0, clic_prior** out) {
  if (n <= 0 || !J0 || !r0 || !out) return fail(CLIC_ERR_BAD_ARGUMENT, "bad prior");
  std::unique_ptr<clic_prior> P(new clic_prior);
  P->n = n;
  P->m = 0;
  P->J0.assign(J0, J0 + (size_t)n * n);
  P->r0.assign(r0, r0 + n);
  int off = 0;
  for (int i = 0; i < n_blocks; i++) {
    clic_prior::Block b;
    b.kind = kind[i];
    b.id = id[i];
    b.size = (b.kind == CLIC_BLOCK_KNOT_ROT) ? 4 : (b.kind <= CLIC_BLOCK_GRAVITY ? 3 : 1);
    b.idx = off;
    off += b.size == 4 ? 3 : b.size;
    for (int k = 0; k < 4; k++) b.x0[k] = x0[4 * i + k];
    P->blocks.push_back(b);
  }
  if (off != n) return fail(CLIC_ERR_BAD_ARGUMENT, "prior blocks do not add up to n");
  *out = P.release();
  return CLIC_OK;
}

CLIC_API int clic_prior_dims(const clic_prior* p, int32_t* n, int32_t* n_blocks, int32_t* m_dropped) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  if (n) *n = p->n;
  if (n_blocks) *n_blocks = (int32_t)p->blocks.size();
  if (m_dropped) *m_dropped = p->m;
  return CLIC_OK;
}

CLIC_API int clic_prior_read(const clic_prior* p, double* J0, double* r0, int32_t* kind, int32_t* id,
                             int32_t* tangent_offset, double* x0) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  if (J0) std::copy(p->J0.begin(), p->J0.end(), J0);
  if (r0) std::copy(p->r0.begin(), p->r0.end(), r0);
  for (size_t i = 0; i < p->blocks.size(); i++) {
    if (kind) kind[i] = p->blocks[i].kind;
    if (id) id[i] = p->blocks[i].id;
    if (tangent_offset) tangent_offset[i] = p->blocks[i].idx;
    if (x0)
      for (int k = 0; k < 4; k++) x0[4 * i + k] = p->blocks[i].x0[k];
  }
  return CLIC_OK;
}

// A = J0^T J0, g = J0^T r0: a read-back convenience on a small host-resident object (not on the solve path,
// where the prior is applied by prior_kernel on the device).
CLIC_API int clic_prior_normal(const clic_prior* p, double* A, double* g) {
  if (!p) return fail(CLIC_ERR_BAD_HANDLE, "bad prior");
  if (!have_device()) return fail(CLIC_ERR_NO_DEVICE, "no CUDA device");
  const int n = p->n;
  DevBuf dJ, dr, dA, dg;
  CU(dJ.ensure((size_t)n * n * sizeof(double)));
  CU(dr.ensure(n * sizeof(double)));
  CU(dA.ensure((size_t)n * n * sizeof(double)));
  CU(dg.ensure(n * sizeof(double)));
  CU(cudaMemcpy(dJ.p, p->J0.data(), (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(dr.p, p->r0.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  dim3 blk(16, 16), grd((n + 15) / 16, (n + 15) / 16);
  prior_gram_kernel<<<grd, blk>>>(dJ.as<double>(), n, dA.as<double>());
  prior_grad_kernel<<<(n + 127) / 128, 128>>>(dJ.as<double>(), dr.as<double>(), n, dg.as<double>());
  if (A) CU(cudaMemcpy(A, dA.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost));
  if (g) CU(cudaMemcpy(g, dg.p, n * sizeof(double), cudaMemcpyDeviceToHost));
  return CLIC_OK;
}

CLIC_API int clic_prior_release(clic_prior* p) {
  delete p;
  return CLIC_OK;
}

CLIC_API int clic_comm_unique_id(uint8_t id[128]) {
  NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(CLIC_ERR_NCCL, "libnccl.so.2 could not be loaded");
  ncclUniqueId u;
  int e = nc.GetUniqueId(&u);
  if (e != 0) return fail(CLIC_ERR_NCCL, "ncclGetUniqueId failed");
  std::memcpy(id, u.internal, 128);
  return CLIC_OK;
}

// Attaches the estimator to the process communicator, creating it on the first call (collective over all ranks);
// later estimators of the same process reuse it and ignore `id`.
CLIC_API int clic_comm_init(clic_estimator* E, int32_t n_ranks, int32_t rank, const uint8_t id[128]) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (n_ranks <= 1) return CLIC_OK;
  NcclApi& nc = nccl_api();
  if (!nc.ok) return fail(CLIC_ERR_NCCL, "libnccl.so.2 could not be loaded");
  CU(cudaSetDevice(E->device));
  if (!g_comm.comm) {
    if (!id) return fail(CLIC_ERR_BAD_ARGUMENT, "unique id required for the first clic_comm_init of the process");
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    int e = nc.CommInitRank(&g_comm.comm, n_ranks, u, rank);
    if (e != 0) return fail(CLIC_ERR_NCCL, std::string("ncclCommInitRank: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "?"));
    g_comm.n_ranks = n_ranks;
    g_comm.rank = rank;
  } else if (g_comm.n_ranks != n_ranks || g_comm.rank != rank) {
    return fail(CLIC_ERR_BAD_ARGUMENT, "the process communicator was created with another size / rank");
  }
  E->use_comm = true;
  E->layout_dirty = true;
  return CLIC_OK;
}

CLIC_API int clic_comm_p2p_export(clic_estimator* E, uint8_t handle[64]) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaSetDevice(E->device));
  if (!g_p2p.local) {
    CU(cudaMalloc(&g_p2p.local, p2p_area_bytes()));
    CU(cudaMemset(g_p2p.local, 0, p2p_area_bytes()));
    CU(cudaMalloc(&g_p2p.d_err, sizeof(int)));
    CU(cudaMemset(g_p2p.d_err, 0, sizeof(int)));
    CU(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, g_p2p.local));
  std::memcpy(handle, &h, 64);
  return CLIC_OK;
}

CLIC_API int clic_comm_p2p_import(clic_estimator* E, int32_t n_ranks, int32_t rank, const uint8_t* handles) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (g_p2p.ready) return CLIC_OK;
  if (n_ranks < 2 || n_ranks > kP2PMaxRanks || rank < 0 || rank >= n_ranks || !handles || !g_p2p.local)
    return fail(CLIC_ERR_BAD_ARGUMENT, "clic_comm_p2p_import: export first; 2..16 ranks");
  CU(cudaSetDevice(E->device));
  g_p2p.view.n = n_ranks;
  g_p2p.view.rank = rank;
  for (int r = 0; r < n_ranks; r++) {
    if (r == rank) { g_p2p.view.peer[r] = static_cast<char*>(g_p2p.local); continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(CLIC_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    g_p2p.view.peer[r] = static_cast<char*>(p);
  }
  g_p2p.ready = true;
  return CLIC_OK;
}

CLIC_API int clic_comm_p2p_ready(void) { return g_p2p.ready ? 1 : 0; }
CLIC_API int clic_comm_p2p_disable(void) { g_p2p.ready = false; return CLIC_OK; }

CLIC_API int clic_linearize_async(clic_estimator* E, int32_t reps) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  if ((rc = fill_prior_columns(E))) return rc;
  for (int i = 0; i < reps; i++)
    if ((rc = run_pass(E, E->d_state.as<double>(), true, 0, nullptr))) return rc;
  return CLIC_OK;
}

CLIC_API int clic_synchronize(clic_estimator* E) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaStreamSynchronize(E->copy_stream));
  CU(cudaStreamSynchronize(E->stream));
  return check_p2p(E);
}

CLIC_API int clic_profile_enable(clic_estimator* E, int32_t enable) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  CU(cudaStreamSynchronize(E->stream));
  for (auto& sp : E->prof_spans) { E->ev_pool.push_back(sp.a); E->ev_pool.push_back(sp.b); }
  E->prof_spans.clear();
  E->prof_on = enable != 0;
  return CLIC_OK;
}

CLIC_API int clic_profile_read(clic_estimator* E, double ms[8], int64_t launches[8]) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  CU(cudaStreamSynchronize(E->stream));
  for (int i = 0; i < 8; i++) { if (ms) ms[i] = 0; if (launches) launches[i] = 0; }
  for (auto& sp : E->prof_spans) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, sp.a, sp.b));
    if (sp.kind >= 0 && sp.kind < 8) { if (ms) ms[sp.kind] += t; if (launches) launches[sp.kind] += 1; }
  }
  return CLIC_OK;
}

CLIC_API int clic_pass_info(clic_estimator* E, int32_t* fused, int32_t* grid, int32_t part_ctas[6]) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  AllocScope alloc_scope(E->stream);
  int rc = ensure_layout(E);
  if (rc) return rc;
  if (fused) *fused = E->fused.ok ? 1 : 0;
  if (grid) *grid = E->fused.ok ? E->fused.grid : 0;
  if (part_ctas)
    for (int k = 0; k < kFusedKinds; k++) part_ctas[k] = E->fused.ok ? E->fused.part_ctas[k] : 0;
  return CLIC_OK;
}

CLIC_API int clic_pass_stamps(clic_estimator* E, int32_t enable, int64_t capacity, uint64_t* stamps, int32_t* n_ctas) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  CU(cudaSetDevice(E->device));
  CU(cudaStreamSynchronize(E->stream));
  if (n_ctas) *n_ctas = E->fused.ok ? E->fused.grid : 0;
  if (stamps && E->fused.ok && E->fused.stamps_on && capacity >= (int64_t)E->fused.grid * 8)
    CU(cudaMemcpy(stamps, E->fused.d_stamps.p, (size_t)E->fused.grid * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  E->fused.stamps_on = enable != 0;
  return CLIC_OK;
}

CLIC_API int clic_num_kernel_launches(clic_estimator* E, int64_t* l) {
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (l) *l = E->launches;
  return CLIC_OK;
}

// ---- LiDAR data association (kernels in assoc.cuh) ----
namespace {
struct AssocSide {  // device layout of one feature type inside the scratch buffer
  int64_t n_feat = 0, step = 1, n_q = 0;
  int geo_w = 4, n_slices = 1, slice_len = 1;
  size_t o_xyzM = 0, o_xyzL = 0, o_t = 0, o_part_d = 0, o_part_i = 0, o_valid = 0, o_geo = 0, o_nn = 0, o_out_t = 0,
         o_out_pt = 0, o_out_geo = 0, o_out_src = 0;
};
size_t assoc_plan(AssocSide& s, size_t off) {
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~(size_t)255;
    return o;
  };
  s.o_xyzM = take((size_t)3 * s.n_feat * sizeof(float));
  s.o_xyzL = take((size_t)3 * s.n_feat * sizeof(float));
  s.o_t = take((size_t)s.n_feat * sizeof(double));
  s.o_part_d = take((size_t)5 * s.n_slices * s.n_q * sizeof(float));
  s.o_part_i = take((size_t)5 * s.n_slices * s.n_q * sizeof(int));
  s.o_valid = take((size_t)s.n_q);
  s.o_geo = take((size_t)6 * s.n_q * sizeof(double));
  s.o_nn = take((size_t)5 * s.n_q * sizeof(int));
  const size_t cap = (size_t)s.n_q + 32 * assoc_dev::kAlignKeys;  // key-aligned output: up to 31 holes per knot interval
  s.o_out_t = take(cap * sizeof(double));
  s.o_out_pt = take(3 * cap * sizeof(double));
  s.o_out_geo = take(s.geo_w * cap * sizeof(double));
  s.o_out_src = take(cap * sizeof(long long));
  return off;
}
}  // namespace

struct clic_feature_map {
  uint32_t magic = 0x4d41504cu;
  int device = 0, n_sm = 148;
  cudaStream_t stream = nullptr;
  float* d_map[2] = {nullptr, nullptr};  // corner, surface: structure of arrays x[n] y[n] z[n]
  int64_t n_map[2] = {0, 0};
  void* d_scratch = nullptr;  // per-call inputs, per-feature results and the compacted outputs
  size_t scratch_bytes = 0;
  int64_t launches = 0;
  AssocSide last[2];  // layout of the last call (clic_correspondence_neighbours)
  cudaEvent_t ev[2] = {nullptr, nullptr};  // around the kernels of the last call
};

CLIC_API int clic_feature_map_create(const float* surf_xyz, int64_t n_surf, const float* corner_xyz, int64_t n_corner,
                                     clic_feature_map** out) {
  if (!out || n_surf < 0 || n_corner < 0 || (n_surf > 0 && !surf_xyz) || (n_corner > 0 && !corner_xyz))
    return fail(CLIC_ERR_BAD_ARGUMENT, "bad feature map");
  if (n_surf > INT32_MAX / 4 || n_corner > INT32_MAX / 4) return fail(CLIC_ERR_BAD_ARGUMENT, "feature map too large");
  if (!have_device()) return fail(CLIC_ERR_NO_DEVICE, "no CUDA device: libclic_cuda has no CPU fallback");
  std::unique_ptr<clic_feature_map> m(new clic_feature_map);
  CU(cudaGetDevice(&m->device));
  CU(cudaDeviceGetAttribute(&m->n_sm, cudaDevAttrMultiProcessorCount, m->device));
  CU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&m->ev[0]));
  CU(cudaEventCreate(&m->ev[1]));
  const float* src[2] = {corner_xyz, surf_xyz};
  m->n_map[0] = n_corner;
  m->n_map[1] = n_surf;
  float* d_tmp = nullptr;
  const size_t nmax = (size_t)std::max(n_corner, n_surf);
  if (nmax > 0) CU(cudaMalloc(&d_tmp, 3 * nmax * sizeof(float)));
  for (int s = 0; s < 2; s++) {
    const size_t n = (size_t)m->n_map[s];
    if (n == 0) continue;
    CU(cudaMalloc(&m->d_map[s], 3 * n * sizeof(float)));
    CU(cudaMemcpyAsync(d_tmp, src[s], 3 * n * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    assoc_dev::soa_kernel<<<(int)std::min<size_t>((n + 255) / 256, 1184), 256, 0, m->stream>>>(
        d_tmp, (int)n, m->d_map[s], m->d_map[s] + n, m->d_map[s] + 2 * n);
    m->launches++;
  }
  CU(cudaStreamSynchronize(m->stream));
  cudaFree(d_tmp);
  CU(cudaGetLastError());
  *out = m.release();
  return CLIC_OK;
}

CLIC_API void clic_feature_map_destroy(clic_feature_map* m) {
  if (!m || m->magic != 0x4d41504cu) return;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->d_map[0]);
  cudaFree(m->d_map[1]);
  cudaFree(m->d_scratch);
  for (auto e : m->ev) if (e) cudaEventDestroy(e);
  if (m->stream) cudaStreamDestroy(m->stream);
  m->magic = 0;
  delete m;
}

CLIC_API int64_t clic_correspondence_capacity(int64_t n) {
  if (n <= 0) return 0;
  const int64_t step = n / 1500 + 1;  // lidar_odometry.cpp:501, 575
  return (n + step - 1) / step;
}

CLIC_API int clic_feature_map_launches(clic_feature_map* m, int64_t* launches, double* last_kernel_ms) {
  if (!m || m->magic != 0x4d41504cu) return fail(CLIC_ERR_BAD_HANDLE, "bad feature map");
  if (launches) *launches = m->launches;
  if (last_kernel_ms) {
    *last_kernel_ms = 0;
    if (m->last[0].n_q + m->last[1].n_q > 0) {
      float ms = 0;
      CU(cudaSetDevice(m->device));
      CU(cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]));
      *last_kernel_ms = ms;
    }
  }
  return CLIC_OK;
}

namespace {
struct AssocPlan {
  AssocSide side[2];
  size_t total = 256;  // two counters in front
  int max_groups = 0, max_slices = 0, max_q = 0;
};
// Stride, grid cut and scratch layout of one call: n_feat[0] corner features, n_feat[1] surface features.
int assoc_plan_call(clic_feature_map* m, int64_t n_corner, int64_t n_surf, AssocPlan& P) {
  P.side[0].n_feat = (n_corner > 0 && m->n_map[0] > 10 /* ":502" */) ? n_corner : 0;
  P.side[0].geo_w = 6;
  P.side[1].n_feat = (n_surf > 0 && m->n_map[1] > 0) ? n_surf : 0;
  P.side[1].geo_w = 4;
  for (int s = 0; s < 2; s++) {
    AssocSide& S = P.side[s];
    S.step = S.n_feat / 1500 + 1;
    S.n_q = (S.n_feat + S.step - 1) / S.step;
    if (S.n_q > 0) {
      // cut the map so that the grid holds about 4 CTAs per SM; a slice is a whole number of shared-memory tiles
      const int groups = (int)((S.n_q + assoc_dev::kQPC - 1) / assoc_dev::kQPC);
      const int tiles = (int)((m->n_map[s] + assoc_dev::kTile - 1) / assoc_dev::kTile);
      S.n_slices = std::max(1, std::min(tiles, (4 * m->n_sm + groups - 1) / groups));
      const int tiles_per_slice = (tiles + S.n_slices - 1) / S.n_slices;
      S.slice_len = tiles_per_slice * assoc_dev::kTile;
      S.n_slices = (int)((m->n_map[s] + S.slice_len - 1) / S.slice_len);
      P.max_groups = std::max(P.max_groups, groups);
      P.max_slices = std::max(P.max_slices, S.n_slices);
      P.max_q = std::max(P.max_q, (int)S.n_q);
    }
    P.total = assoc_plan(S, P.total);
  }
  if (P.total > m->scratch_bytes) {
    CU(cudaDeviceSynchronize());
    cudaFree(m->d_scratch);
    m->d_scratch = nullptr;
    m->scratch_bytes = 0;
    CU(cudaMalloc(&m->d_scratch, P.total + P.total / 4));
    m->scratch_bytes = P.total + P.total / 4;
  }
  m->last[0] = P.side[0];
  m->last[1] = P.side[1];
  return CLIC_OK;
}
assoc_dev::AssocJobs assoc_jobs(clic_feature_map* m, const AssocPlan& P) {
  char* base = (char*)m->d_scratch;
  assoc_dev::AssocJobs jobs;
  for (int s = 0; s < 2; s++) {
    const AssocSide& S = P.side[s];
    assoc_dev::AssocJob& J = jobs.j[s];
    const size_t nm = (size_t)m->n_map[s];
    J.mx = m->d_map[s];
    J.my = m->d_map[s] + nm;
    J.mz = m->d_map[s] + 2 * nm;
    J.n_map = (int)nm;
    J.xyz_M = (const float*)(base + S.o_xyzM);
    J.xyz_L = (const float*)(base + S.o_xyzL);
    J.t = (const double*)(base + S.o_t);
    J.step = S.step;
    J.n_q = (int)S.n_q;
    J.is_surface = s;
    J.n_slices = S.n_slices;
    J.slice_len = S.slice_len;
    J.part_d = (float*)(base + S.o_part_d);
    J.part_i = (int*)(base + S.o_part_i);
    J.valid = (unsigned char*)(base + S.o_valid);
    J.geo = (double*)(base + S.o_geo);
    J.nn_idx = (int*)(base + S.o_nn);
    J.geo_w = S.geo_w;
    J.out_t = (double*)(base + S.o_out_t);
    J.out_point = (double*)(base + S.o_out_pt);
    J.out_geo = (double*)(base + S.o_out_geo);
    J.out_src = (long long*)(base + S.o_out_src);
    J.count = (long long*)base + s;
  }
  return jobs;
}
// The search and the fits: valid / geo / nn_idx of every selected feature.
void assoc_search(clic_feature_map* m, const AssocPlan& P, const assoc_dev::AssocJobs& jobs, cudaStream_t st) {
  assoc_dev::knn5_partial_kernel<<<dim3(P.max_groups, P.max_slices, 2), assoc_dev::kWarps * 32, 0, st>>>(jobs);
  assoc_dev::merge_fit_kernel<<<dim3((P.max_q + assoc_dev::kWarps - 1) / assoc_dev::kWarps, 1, 2), assoc_dev::kWarps * 32, 0,
                                st>>>(jobs);
  m->launches += 2;
}
}  // namespace

CLIC_API int clic_find_correspondence(clic_feature_map* m, int64_t n_corner, const float* corner_xyz_L,
                                      const double* corner_t, const float* corner_xyz_M, int64_t n_surf,
                                      const float* surf_xyz_L, const double* surf_t, const float* surf_xyz_M,
                                      int64_t* n_lines, double* line_t, double* line_point, double* line_geo,
                                      int64_t* line_src, int64_t* n_planes, double* plane_t, double* plane_point,
                                      double* plane_geo, int64_t* plane_src) {
  if (!m || m->magic != 0x4d41504cu) return fail(CLIC_ERR_BAD_HANDLE, "bad feature map");
  if (!n_lines || !n_planes) return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  *n_lines = *n_planes = 0;
  if ((n_corner > 0 && (!corner_xyz_L || !corner_t || !corner_xyz_M || !line_t || !line_point || !line_geo)) ||
      (n_surf > 0 && (!surf_xyz_L || !surf_t || !surf_xyz_M || !plane_t || !plane_point || !plane_geo)))
    return fail(CLIC_ERR_BAD_ARGUMENT, "null argument");
  CU(cudaSetDevice(m->device));
  AssocPlan P;
  int rc = assoc_plan_call(m, n_corner, n_surf, P);
  if (rc) return rc;
  if (P.max_q == 0) return CLIC_OK;
  char* base = (char*)m->d_scratch;
  const float* in_M[2] = {corner_xyz_M, surf_xyz_M};
  const float* in_L[2] = {corner_xyz_L, surf_xyz_L};
  const double* in_t[2] = {corner_t, surf_t};
  for (int s = 0; s < 2; s++) {
    const AssocSide& S = P.side[s];
    if (S.n_q == 0) continue;
    CU(cudaMemcpyAsync(base + S.o_xyzM, in_M[s], (size_t)3 * S.n_feat * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(base + S.o_xyzL, in_L[s], (size_t)3 * S.n_feat * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(base + S.o_t, in_t[s], (size_t)S.n_feat * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  }
  const assoc_dev::AssocJobs jobs = assoc_jobs(m, P);
  CU(cudaEventRecord(m->ev[0], m->stream));
  assoc_search(m, P, jobs, m->stream);
  assoc_dev::compact_kernel<<<dim3(1, 1, 2), 1024, 0, m->stream>>>(jobs);
  m->launches += 1;
  CU(cudaEventRecord(m->ev[1], m->stream));
  long long h_count[2] = {0, 0};
  CU(cudaMemcpyAsync(h_count, base, sizeof(h_count), cudaMemcpyDeviceToHost, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  CU(cudaGetLastError());
  double* out_t[2] = {line_t, plane_t};
  double* out_pt[2] = {line_point, plane_point};
  double* out_geo[2] = {line_geo, plane_geo};
  int64_t* out_src[2] = {line_src, plane_src};
  for (int s = 0; s < 2; s++) {
    const AssocSide& S = P.side[s];
    const size_t c = (size_t)h_count[s];
    if (c == 0) continue;
    CU(cudaMemcpyAsync(out_t[s], base + S.o_out_t, c * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaMemcpyAsync(out_pt[s], base + S.o_out_pt, 3 * c * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaMemcpyAsync(out_geo[s], base + S.o_out_geo, S.geo_w * c * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    if (out_src[s])
      CU(cudaMemcpyAsync(out_src[s], base + S.o_out_src, c * sizeof(long long), cudaMemcpyDeviceToHost, m->stream));
  }
  CU(cudaStreamSynchronize(m->stream));
  *n_lines = h_count[0];
  *n_planes = h_count[1];
  return CLIC_OK;
}

CLIC_API int clic_add_loam_from_scan(clic_estimator* E, clic_feature_map* m, int64_t n_corner, const float* corner_xyz,
                                     const double* corner_t, int64_t n_surf, const float* surf_xyz, const double* surf_t,
                                     const double S_GtoM[4], const double p_GinM[3], const double S_LtoI[4],
                                     const double p_LinI[3], double t_offset_ns, int32_t cor_downsample, double weight,
                                     int32_t marg_this_factor, int64_t* n_lines, int64_t* n_planes, int64_t* n_dropped) {
  ScopedTrace trace("clic_add_loam_from_scan");
  if (!ok_handle(E)) return fail(CLIC_ERR_BAD_HANDLE, "bad handle");
  if (!m || m->magic != 0x4d41504cu) return fail(CLIC_ERR_BAD_HANDLE, "bad feature map");
  if (n_lines) *n_lines = 0;
  if (n_planes) *n_planes = 0;
  if (n_dropped) *n_dropped = 0;
  if ((n_corner > 0 && (!corner_xyz || !corner_t)) || (n_surf > 0 && (!surf_xyz || !surf_t)) || !S_GtoM || !p_GinM ||
      !S_LtoI || !p_LinI || cor_downsample < 1)
    return fail(CLIC_ERR_BAD_ARGUMENT, "bad argument");
  if (m->device != E->device) return fail(CLIC_ERR_BAD_ARGUMENT, "feature map lives on another device");
  CU(cudaSetDevice(E->device));
  cudaStream_t st = E->stream;
  AllocScope alloc_scope(st);
  AssocPlan P;
  int rc = assoc_plan_call(m, n_corner, n_surf, P);
  if (rc) return rc;
  if (P.max_q == 0) return CLIC_OK;
  // earlier asynchronous adders may still be running their prep kernels on the copy stream; they share the dropped-
  // record counter with the adders below
  if ((rc = join_adds(E))) return rc;
  char* base = (char*)m->d_scratch;
  const float* in_L[2] = {corner_xyz, surf_xyz};
  const double* in_t[2] = {corner_t, surf_t};
  const PassParams pp = pass_params(E, E->d_state.as<double>(), nullptr);
  const size_t smem = window_smem_doubles(E->K) * sizeof(double);
  CU(cudaMemsetAsync(E->d_counter.as<unsigned long long>() + 1, 0, sizeof(unsigned long long), st));
  for (int s = 0; s < 2; s++) {
    const AssocSide& S = P.side[s];
    if (S.n_q == 0) continue;
    CU(cudaMemcpyAsync(base + S.o_xyzL, in_L[s], (size_t)3 * S.n_feat * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + S.o_t, in_t[s], (size_t)S.n_feat * sizeof(double), cudaMemcpyHostToDevice, st));
    // Trajectory::UndistortScanInG with the current state: the scan in the map frame stays on the device
    const PoseQuery q = make_pose_query(E, S.n_feat, (const double*)(base + S.o_t), S_LtoI, p_LinI, t_offset_ns);
    undistort_kernel<<<query_grid(E, S.n_feat), 256, smem, st>>>(pp, q, (const float*)(base + S.o_xyzL), nullptr,
                                                                 (float*)(base + S.o_xyzM),
                                                                 E->d_counter.as<unsigned long long>() + 1);
    E->launches++;
  }
  const assoc_dev::AssocJobs jobs = assoc_jobs(m, P);
  // sort_downsample_kernel packs (type << 16) | q and sorts at most kSortMax records in one CTA; the plan's stride
  // (n / 1500 + 1) keeps n_q <= 1500 per type — refuse instead of truncating if that ever stops being true
  if (jobs.j[0].n_q + jobs.j[1].n_q > assoc_dev::kSortMax || jobs.j[0].n_q > 65535 || jobs.j[1].n_q > 65535)
    return fail(CLIC_ERR_BAD_ARGUMENT, "clic_add_loam_from_scan: more selected features than the correspondence sort holds");
  assoc_search(m, P, jobs, st);
  assoc_dev::SortedOut out;
  for (int s = 0; s < 2; s++) {
    out.t[s] = jobs.j[s].out_t;
    out.point[s] = jobs.j[s].out_point;
    out.geo[s] = jobs.j[s].out_geo;
    out.src[s] = jobs.j[s].out_src;
  }
  out.count = (long long*)base;
  out.t0_ns = E->t0_ns;
  out.dt_ns = E->dt_ns;
  out.n_keys = key_align_enabled() ? E->K - 3 : 0;
  static const bool sort_attr = [] {
    return cudaFuncSetAttribute(assoc_dev::sort_downsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                assoc_dev::kSortMax * 12) == cudaSuccess;
  }();
  if (!sort_attr) return fail(CLIC_ERR_CUDA, "cannot reserve shared memory for the correspondence sort");
  assoc_dev::sort_downsample_kernel<<<1, 1024, assoc_dev::kSortMax * 12, st>>>(jobs, cor_downsample, out);
  m->launches += 1;
  E->launches += 3;
  long long h_count[4] = {0, 0, 0, 0};  // stored (padded) lines, planes; kept lines, planes
  CU(cudaMemcpyAsync(h_count, base, sizeof(h_count), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  int64_t dropped = 0, d1 = 0;
  // line batch first, then the plane batch — each time-sorted, read by the prep kernel straight from the scratch
  if ((rc = add_loam_records(E, 2, h_count[0], out.t[0], out.point[0], nullptr, out.geo[0], S_GtoM, p_GinM, S_LtoI, p_LinI,
                             weight, marg_this_factor, n_dropped ? &d1 : nullptr, true, st)))
    return rc;
  dropped += d1;
  d1 = 0;
  if ((rc = add_loam_records(E, 1, h_count[1], out.t[1], out.point[1], nullptr, out.geo[1], S_GtoM, p_GinM, S_LtoI, p_LinI,
                             weight, marg_this_factor, n_dropped ? &d1 : nullptr, true, st)))
    return rc;
  dropped += d1;
  CU(cudaStreamSynchronize(st));  // the scratch may be reused by the next call on any stream
  if (n_lines) *n_lines = h_count[2];
  if (n_planes) *n_planes = h_count[3];
  // the holes of the key-aligned layout were stored as dropped records: they are not measurements
  if (n_dropped) *n_dropped = dropped - (h_count[0] - h_count[2]) - (h_count[1] - h_count[3]);
  return CLIC_OK;
}

CLIC_API int clic_correspondence_neighbours(clic_feature_map* m, int32_t type, int64_t capacity, int32_t* idx /*5n*/,
                                            int64_t* n_out) {
  if (!m || m->magic != 0x4d41504cu) return fail(CLIC_ERR_BAD_HANDLE, "bad feature map");
  if (type < 0 || type > 1 || !n_out) return fail(CLIC_ERR_BAD_ARGUMENT, "bad argument");
  const AssocSide& S = m->last[type];
  *n_out = S.n_q;
  if (S.n_q == 0 || !idx) return CLIC_OK;
  if (capacity < S.n_q) return fail(CLIC_ERR_BAD_ARGUMENT, "capacity too small");
  CU(cudaSetDevice(m->device));
  CU(cudaMemcpyAsync(idx, (char*)m->d_scratch + S.o_nn, (size_t)5 * S.n_q * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  CU(cudaStreamSynchronize(m->stream));
  return CLIC_OK;
}

#ifdef CLIC_LM_TIMING
CLIC_API int clic_debug_lm_stamps(long long* out /*16*/) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(out, clic::g_lm_stamps, 16 * sizeof(long long)));
  return CLIC_OK;
}
#endif

}  // extern "C"
