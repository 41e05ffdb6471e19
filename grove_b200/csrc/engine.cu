// engine.cu -- libgrove_place.so: host side of the placement engine and the C ABI of
// include/grove_place.h.  Everything that computes a placement runs in the kernels of kernels.cuh;
// the host sorts the topology once per label change, validates and uploads tables, and drives the
// optimistic rounds.  There is no CPU fallback: without a CUDA device the engine cannot be created.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <array>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>
#include <numeric>
#include <string>
#include <vector>

#include "kernels.cuh"

using namespace grove;

namespace {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  // grows geometrically: a pending set that creeps up tick after tick (churn) must not pay a cudaFree + cudaMalloc
  // of every table per cycle; falls back to the exact size when the padded one does not fit
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    const size_t want = std::max<size_t>(std::max<size_t>(n, (64u << 10) / sizeof(T)), cap + cap / 2);
    release();
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) { cap = want; return e; }
    (void)cudaGetLastError();
    e = cudaMalloc(reinterpret_cast<void**>(&p), std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = std::max<size_t>(n, 1); else p = nullptr;
    return e;
  }
};

template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    // pinned allocations cost milliseconds: start at 256 KB and double, so a growing pending set re-allocates rarely
    const size_t want = std::max<size_t>(std::max<size_t>(n, (256u << 10) / sizeof(T)), cap * 2);
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) cap = want; else p = nullptr;
    return e;
  }
};

// Size of the engine's host thread team (table validation, signature interning, staging copies).  One size for every
// parallel region: libgomp re-creates its team whenever the size changes.  Not taken from OMP_NUM_THREADS: launchers
// such as torchrun export OMP_NUM_THREADS=1 for every rank, which would serialise the per-cycle host work; the knob is
// GROVE_HOST_THREADS (default 8), capped by the processors this process may run on.
int host_threads() {
  static const int t = [] {
    int want = 8;
    if (const char* v = std::getenv("GROVE_HOST_THREADS")) want = std::atoi(v);
    return std::max(1, std::min(want, omp_get_num_procs()));
  }();
  return t;
}

// a host table the engine owns, in pinned memory: the caller's array is copied in once (by a few threads when it is
// large) and the upload from it is a plain asynchronous DMA
template <typename T>
struct PinVec {
  PinBuf<T> b;
  size_t n = 0;
  T& operator[](size_t i) { return b.p[i]; }
  const T& operator[](size_t i) const { return b.p[i]; }
  T* data() { return b.p; }
  cudaError_t assign(const T* src, size_t cnt) {
    cudaError_t e = b.ensure(cnt);
    if (e != cudaSuccess) return e;
    n = cnt;
    const size_t bytes = cnt * sizeof(T);
    // one team size for every parallel region of the engine: libgomp re-creates its thread team whenever the size changes
    const int T_ = bytes >= (256u << 10) ? host_threads() : 1;
#pragma omp parallel for num_threads(T_) schedule(static)
    for (int t = 0; t < T_; ++t) {
      const size_t a = bytes * t / T_, z = bytes * (t + 1) / T_;
      std::memcpy(reinterpret_cast<char*>(b.p) + a, reinterpret_cast<const char*>(src) + a, z - a);
    }
    return cudaSuccess;
  }
};

uint32_t fmix32(uint32_t x) {
  x = x * 0x9E3779B1u + 0x7F4A7C15u;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}

}  // namespace

struct grove_engine {
  grove_config_t cfg{};
  std::string err;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_score = nullptr;  // K2 runs beside K3 (they only share the fit data)
  cudaEvent_t ev_fit = nullptr, ev_score = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
  cudaEvent_t ev_nodes_up = nullptr, ev_tables_up = nullptr;  // the last upload out of the pinned node staging buffer / gang tables
  cudaEvent_t ev[10]{};

  // ---- topology (static until labels change) ----
  uint32_t N = 0, Npad = 0, L = 0, words = 0;
  std::vector<uint32_t> raw_dom;      // caller order, N * MAX_LEVELS: cache key
  std::vector<uint32_t> perm, inv;    // sorted <-> caller
  std::vector<uint32_t> tdom;         // sorted order, N * MAX_LEVELS tree-ified ids
  std::vector<uint8_t> vdepth;
  std::vector<uint32_t> dom_lo[GROVE_MAX_LEVELS], dom_hi[GROVE_MAX_LEVELS];
  uint32_t n_dom[GROVE_MAX_LEVELS]{}, unit[GROVE_MAX_LEVELS]{};
  DevBuf<grove_node_t> d_nodes_in;    // caller order, as loaded
  DevBuf<uint4> d_nres, d_ndom;
  DevBuf<uint32_t> d_perm, d_inv;
  DevBuf<uint8_t> d_vdepth;
  DevBuf<uint32_t> d_dom_lo[GROVE_MAX_LEVELS], d_dom_hi[GROVE_MAX_LEVELS], d_next_dom[GROVE_MAX_LEVELS];
  bool nodes_loaded = false;

  // ---- gang tables ----
  uint32_t G = 0, Q = 0, S = 0, P = 0;
  PinVec<grove_gang_t> gangs;
  PinVec<grove_clique_t> cliques;
  PinVec<grove_scope_t> scopes;
  PinBuf<uint8_t> h_state0;        // initial gang states of a cycle
  PinBuf<GangInfo> ginfo_pin;      // derived tables are built straight into pinned memory: their upload is a plain DMA
  PinBuf<CliqueInfo> cinfo_pin;
  GangInfo* ginfo = nullptr;
  CliqueInfo* cinfo = nullptr;
  std::vector<uint4> sigs;
  uint32_t n_sigs = 0;
  bool gangs_loaded = false, ginfo_dirty = true;
  DevBuf<grove_gang_t> d_gangs;
  DevBuf<grove_clique_t> d_cliques;
  DevBuf<grove_scope_t> d_scopes;
  DevBuf<GangInfo> d_ginfo;
  DevBuf<CliqueInfo> d_cinfo;
  DevBuf<uint4> d_sigs;
  DevBuf<uint32_t> d_sig_stamp, d_sig_list;

  // ---- round state ----
  DevBuf<uint8_t> d_state, d_round, d_spec_score, d_T;
  DevBuf<uint16_t> d_spec_n, d_ent_meta;
  DevBuf<uint32_t> d_active, d_rows, d_counters, d_spec_top, d_ent_node, d_claim, d_F, d_totals;
  DevBuf<grove_gang_status_t> d_status;
  DevBuf<grove_placement_t> d_out;
  DevBuf<uint32_t> d_xbuf, d_active_all, d_flags, d_capsum, d_capmax;
  DevBuf<uint8_t> d_taken, d_cur, d_prop;
  uint32_t K = GROVE_MAX_ALTERNATIVES;
  bool any_preferred = false;  // some gang / scope / clique carries a Preferred level
  uint32_t n_constrained = 0, n_unconstrained = 0;  // gangs with / without a gang-level Required or Preferred level
  uint32_t max_gang_pods = 0;
  int resolve_blocks_per_sm = 0;
  uint32_t n_sm = 148;
  DevBuf<uint8_t> d_cap8;
  uint32_t cap_off[GROVE_MAX_LEVELS]{}, cap_stride = 0;
  bool prefilter = false;
  bool dbg_on = false;
  DevBuf<uint32_t> d_dbg;
  int tune_prefilter = 2;   // 0 off, 1 tables for the pre-filter, 2 tables also feed the packing
  uint32_t tune_width0 = 24;  // packing attempts per window in the warp-per-gang kernel
  uint32_t tune_width1 = 32;  // packing attempts per warp in the first window of the CTA-per-gang kernels (doubles per window)
  // admission kernel form by number of active gangs: >= 10 per SM a warp per gang (gangs in flight matter),
  // >= 4 per SM a 4-warp CTA per gang, below that an 8-warp CTA per gang (latency of one gang matters)
  uint32_t tune_warp_min = 1480, tune_wide_max = 592;
  bool tune_overlap = true;      // K2 on a second stream beside K3 (GROVE_TUNE_OVERLAP=0 serialises them, e.g. to time K2 alone)
  uint32_t tune_resolve_bps = 8;  // k_resolve CTAs per SM at most (fewer CTAs = cheaper grid barriers)
  DevBuf<uint32_t> d_upd_idx;
  DevBuf<grove_node_t> d_upd_recs;
  DevBuf<grove_node_t> d_nodes_out;  // grove_get_nodes scratch
  PinBuf<uint32_t> h_counters;
  PinBuf<grove_gang_status_t> h_status;
  PinBuf<grove_placement_t> h_out;
  PinBuf<grove_node_t> h_stage_nodes;
  uint32_t n_out = 0;
  bool have_results = false;
  grove_cycle_stats_t last{};

  // stepping state (multi-GPU)
  uint32_t round_no = 0;
  bool in_cycle = false;
  bool stream_ordered = false;  // stepping calls return without the trailing host sync
  uint64_t pairs = 0, launches = 0;
};

#define CU_TRY(e, expr)                                                                       \
  do {                                                                                        \
    cudaError_t _c = (expr);                                                                  \
    if (_c != cudaSuccess) {                                                                  \
      (e)->err = std::string(#expr) + ": " + cudaGetErrorString(_c);                          \
      return _c == cudaErrorMemoryAllocation ? GROVE_ERR_OOM : GROVE_ERR_CUDA;                \
    }                                                                                         \
  } while (0)

static int32_t fail(grove_engine* e, int32_t code, const char* msg) { e->err = msg; return code; }

// ---------------------------------------------------------------------------------------------
// topology: sort by label path, tree-ify ids, domain ranges.  Host side; runs only when labels change.
// Restates what the reference hands a scheduler as the ordered level list
// (operator/internal/scheduler/kai/topology.go:103-135) plus the per-node label values.
// ---------------------------------------------------------------------------------------------
static int32_t build_topology(grove_engine* e, const grove_node_t* nodes, uint32_t n) {
  const uint32_t L = e->L;
  e->N = n;
  e->Npad = (n + 1023u) & ~1023u;
  e->words = e->Npad / 32;
  e->raw_dom.resize(size_t(n) * GROVE_MAX_LEVELS);
  for (uint32_t i = 0; i < n; ++i) std::memcpy(&e->raw_dom[size_t(i) * GROVE_MAX_LEVELS], nodes[i].dom, sizeof(uint32_t) * GROVE_MAX_LEVELS);
  e->perm.resize(n); e->inv.resize(n);
  std::iota(e->perm.begin(), e->perm.end(), 0u);
  const uint32_t* rd = e->raw_dom.data();
  std::sort(e->perm.begin(), e->perm.end(), [rd, L](uint32_t a, uint32_t b) {
    const uint32_t* x = rd + size_t(a) * GROVE_MAX_LEVELS; const uint32_t* y = rd + size_t(b) * GROVE_MAX_LEVELS;
    for (uint32_t l = 0; l < L; ++l) if (x[l] != y[l]) return x[l] < y[l];
    return a < b;
  });
  for (uint32_t i = 0; i < n; ++i) e->inv[e->perm[i]] = i;
  e->tdom.assign(size_t(e->Npad) * GROVE_MAX_LEVELS, GROVE_DOM_ABSENT);
  e->vdepth.assign(e->Npad, 0);
  for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) { e->dom_lo[l].clear(); e->dom_hi[l].clear(); e->n_dom[l] = 0; e->unit[l] = 0; }
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t* cur = rd + size_t(e->perm[i]) * GROVE_MAX_LEVELS;
    const uint32_t* prv = i ? rd + size_t(e->perm[i - 1]) * GROVE_MAX_LEVELS : nullptr;
    bool same_path = prv != nullptr;  // do cur and prv agree (and exist) on every level above l?
    uint32_t depth = 0;
    for (uint32_t l = 0; l < L; ++l) {
      if (cur[l] == GROVE_DOM_ABSENT) break;  // deeper labels are ignored once one is missing
      bool same = same_path && e->vdepth[i - 1] > l && prv[l] == cur[l];
      if (same) {
        uint32_t id = e->tdom[size_t(i - 1) * GROVE_MAX_LEVELS + l];
        e->tdom[size_t(i) * GROVE_MAX_LEVELS + l] = id;
        e->dom_hi[l][id] = i + 1;
      } else {
        e->tdom[size_t(i) * GROVE_MAX_LEVELS + l] = uint32_t(e->dom_lo[l].size());
        e->dom_lo[l].push_back(i); e->dom_hi[l].push_back(i + 1);
      }
      same_path = same;
      depth = l + 1;
    }
    e->vdepth[i] = uint8_t(depth);
  }
  for (uint32_t l = 0; l < L; ++l) {
    e->n_dom[l] = uint32_t(e->dom_lo[l].size());
    bool u = e->n_dom[l] > 0;
    for (uint32_t d = 0; d < e->n_dom[l] && u; ++d) u = (e->dom_hi[l][d] - e->dom_lo[l][d]) == 1;
    e->unit[l] = u ? 1u : 0u;
  }
  e->cap_stride = 0;
  for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) {
    e->cap_off[l] = e->cap_stride;
    if (l < L && !e->unit[l]) e->cap_stride += e->n_dom[l];
  }
  // upload static tables
  CU_TRY(e, e->d_perm.ensure(n)); CU_TRY(e, e->d_inv.ensure(n));
  CU_TRY(e, e->d_ndom.ensure(e->Npad)); CU_TRY(e, e->d_nres.ensure(e->Npad)); CU_TRY(e, e->d_vdepth.ensure(e->Npad));
  CU_TRY(e, cudaMemcpyAsync(e->d_perm.p, e->perm.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_inv.p, e->inv.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_ndom.p, e->tdom.data(), sizeof(uint4) * e->Npad, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_vdepth.p, e->vdepth.data(), e->Npad, cudaMemcpyHostToDevice, e->stream));
  std::vector<uint32_t> nxt(size_t(n) + 1);
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t nd = e->n_dom[l];
    CU_TRY(e, e->d_dom_lo[l].ensure(nd)); CU_TRY(e, e->d_dom_hi[l].ensure(nd)); CU_TRY(e, e->d_next_dom[l].ensure(size_t(n) + 1));
    if (nd) {
      CU_TRY(e, cudaMemcpyAsync(e->d_dom_lo[l].p, e->dom_lo[l].data(), sizeof(uint32_t) * nd, cudaMemcpyHostToDevice, e->stream));
      CU_TRY(e, cudaMemcpyAsync(e->d_dom_hi[l].p, e->dom_hi[l].data(), sizeof(uint32_t) * nd, cudaMemcpyHostToDevice, e->stream));
    }
    uint32_t d = 0;
    for (uint32_t i = 0; i <= n; ++i) {  // first domain whose lo >= i
      while (d < nd && e->dom_lo[l][d] < i) ++d;
      nxt[i] = d;
    }
    CU_TRY(e, cudaMemcpyAsync(e->d_next_dom[l].p, nxt.data(), sizeof(uint32_t) * (size_t(n) + 1), cudaMemcpyHostToDevice, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));  // nxt is reused
  }
  e->ginfo_dirty = true;
  return GROVE_OK;
}

static Topo make_topo(grove_engine* e) {
  Topo t{};
  t.nres = e->d_nres.p; t.ndom = e->d_ndom.p;
  for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) {
    t.dom_lo[l] = e->d_dom_lo[l].p; t.dom_hi[l] = e->d_dom_hi[l].p; t.next_dom[l] = e->d_next_dom[l].p;
    t.n_dom[l] = e->n_dom[l]; t.unit[l] = e->unit[l];
  }
  t.n = e->N; t.npad = e->Npad; t.L = e->L; t.words = e->words;
  for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) t.cap_off[l] = e->cap_off[l];
  t.cap_stride = e->cap_stride;
  return t;
}

static Tables make_tables(grove_engine* e) {
  Tables t{};
  t.gangs = e->d_gangs.p; t.cliques = e->d_cliques.p; t.scopes = e->d_scopes.p;
  t.ginfo = e->d_ginfo.p; t.cinfo = e->d_cinfo.p; t.sigs = e->d_sigs.p; t.G = e->G; t.Q = e->Q; t.S = e->n_sigs;
  return t;
}

static RoundBufs make_bufs(grove_engine* e) {
  RoundBufs r{};
  r.state = e->d_state.p; r.round = e->d_round.p; r.active = e->d_active.p; r.rows = e->d_rows.p;
  r.counters = e->d_counters.p; r.spec_score = e->d_spec_score.p;
  r.spec_n = e->d_spec_n.p; r.spec_top = e->d_spec_top.p; r.ent_node = e->d_ent_node.p; r.ent_meta = e->d_ent_meta.p;
  r.sig_stamp = e->d_sig_stamp.p; r.sig_list = e->d_sig_list.p;
  r.active_all = e->d_active_all.p; r.taken = e->d_taken.p; r.cur = e->d_cur.p; r.prop = e->d_prop.p; r.flags = e->d_flags.p;
  {
    const size_t KP = size_t(e->K) * e->P, GK = size_t(e->G) * e->K;
    uint32_t* x = e->d_xbuf.p;
    r.alt_node = x; r.alt_meta = x + KP; r.alt_n = x + 2 * KP; r.alt_score = x + 2 * KP + GK; r.alt_top = x + 2 * KP + 2 * GK;
    r.alt_nmin = x + 2 * KP + 3 * GK; r.nalt = x + 2 * KP + 4 * GK; r.K = e->K; r.P = e->P;
  }
  r.claim = e->d_claim.p; r.F = e->d_F.p; r.T = e->d_T.p;
  r.cap8 = e->prefilter ? e->d_cap8.p : nullptr; r.capsum = e->d_capsum.p; r.capmax = e->d_capmax.p;
  r.caps_in_attempts = e->tune_prefilter >= 2; r.width0 = e->tune_width0; r.width1 = e->tune_width1; r.dbg = e->dbg_on ? e->d_dbg.p : nullptr;
  return r;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t grove_abi_version(void) { return GROVE_ABI_VERSION; }

const char* grove_last_error(grove_engine_t* e) { return e ? e->err.c_str() : "null engine"; }

int32_t grove_engine_create(const grove_config_t* cfg, grove_engine_t** out) {
  if (!cfg || !out) return GROVE_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->abi_version != GROVE_ABI_VERSION) return GROVE_ERR_INVALID_ARG;
  if (cfg->n_levels < 1 || cfg->n_levels > GROVE_MAX_LEVELS) return GROVE_ERR_INVALID_ARG;
  if (cfg->world > 1 && cfg->rank >= cfg->world) return GROVE_ERR_INVALID_ARG;
  if (cfg->alternatives > GROVE_MAX_ALTERNATIVES) return GROVE_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GROVE_ERR_NO_DEVICE;  // no CPU fallback
  if (cfg->device < 0 || cfg->device >= ndev) return GROVE_ERR_NO_DEVICE;
  if (cudaSetDevice(cfg->device) != cudaSuccess) return GROVE_ERR_NO_DEVICE;
  grove_engine* e = new (std::nothrow) grove_engine();
  if (!e) return GROVE_ERR_OOM;
  e->cfg = *cfg; e->L = cfg->n_levels;
  e->K = cfg->alternatives ? cfg->alternatives : GROVE_MAX_ALTERNATIVES;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->resolve_blocks_per_sm, k_resolve, kResolveThreads, 0);
  if (e->resolve_blocks_per_sm < 1) e->resolve_blocks_per_sm = 1;
  { int sm = 0; if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && sm > 0) e->n_sm = uint32_t(sm); }
  e->tune_warp_min = 10 * e->n_sm; e->tune_wide_max = 4 * e->n_sm;
  if (const char* v = std::getenv("GROVE_TUNE_WARP_MIN")) e->tune_warp_min = uint32_t(std::max(1, std::atoi(v)));
  if (const char* v = std::getenv("GROVE_TUNE_WIDE_MAX")) e->tune_wide_max = uint32_t(std::max(1, std::atoi(v)));
  if (const char* v = std::getenv("GROVE_TUNE_PREFILTER")) e->tune_prefilter = std::atoi(v);
  if (std::getenv("GROVE_DEBUG_ADMIT")) e->dbg_on = true;

  if (const char* v = std::getenv("GROVE_TUNE_OVERLAP")) e->tune_overlap = std::atoi(v) != 0;
  if (const char* v = std::getenv("GROVE_TUNE_RESOLVE_BPS")) e->tune_resolve_bps = uint32_t(std::max(1, std::atoi(v)));
  if (const char* v = std::getenv("GROVE_TUNE_WIDTH0")) e->tune_width0 = uint32_t(std::min(32, std::max(1, std::atoi(v))));
  if (const char* v = std::getenv("GROVE_TUNE_ALTERNATIVES")) if (!cfg->alternatives) e->K = uint32_t(std::min<int>(GROVE_MAX_ALTERNATIVES, std::max(1, std::atoi(v))));
  if (const char* v = std::getenv("GROVE_TUNE_WIDTH1")) e->tune_width1 = uint32_t(std::min(32, std::max(1, std::atoi(v))));
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // the latency-bound admission gets the SMs first
  if (cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete e; return GROVE_ERR_CUDA; }
  if (cudaStreamCreateWithPriority(&e->stream_score, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete e; return GROVE_ERR_CUDA; }
  if (cudaEventCreateWithFlags(&e->ev_fit, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_score, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_nodes_up, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&e->ev_tables_up, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreate(&e->ev_s0) != cudaSuccess || cudaEventCreate(&e->ev_s1) != cudaSuccess) { delete e; return GROVE_ERR_CUDA; }
  for (auto& ev : e->ev) if (cudaEventCreate(&ev) != cudaSuccess) { delete e; return GROVE_ERR_CUDA; }
  if (e->h_counters.ensure(8) != cudaSuccess) { delete e; return GROVE_ERR_OOM; }
  *out = e;
  return GROVE_OK;
}

void grove_engine_destroy(grove_engine_t* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
  if (e->ev_s0) cudaEventDestroy(e->ev_s0);
  if (e->ev_s1) cudaEventDestroy(e->ev_s1);
  if (e->ev_fit) cudaEventDestroy(e->ev_fit);
  if (e->ev_nodes_up) cudaEventDestroy(e->ev_nodes_up);
  if (e->ev_tables_up) cudaEventDestroy(e->ev_tables_up);
  if (e->ev_score) cudaEventDestroy(e->ev_score);
  if (e->stream_score) cudaStreamDestroy(e->stream_score);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int32_t load_nodes_common(grove_engine* e, const grove_node_t* host_nodes, const void* dev_nodes, uint32_t n) {
  if (n == 0 || n > GROVE_MAX_NODES) return fail(e, GROVE_ERR_INVALID_ARG, "node count out of range");
  if (e->in_cycle) return fail(e, GROVE_ERR_STATE, "cycle in flight");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  if (host_nodes) {
    CU_TRY(e, cudaEventSynchronize(e->ev_nodes_up));  // an earlier upload may still be reading the staging buffer
    CU_TRY(e, e->d_nodes_in.ensure(n));
    CU_TRY(e, e->h_stage_nodes.ensure(n));
    // one pass over the caller's snapshot by a few threads: copy to the pinned staging buffer (the caller's array is
    // not retained) and compare the labels with the cached topology's
    const bool cached = e->nodes_loaded && n == e->N;
    const int T = n >= 8192 ? host_threads() : 1;
    int differs = 0;
#pragma omp parallel for num_threads(T) schedule(static) reduction(| : differs)
    for (int t = 0; t < T; ++t) {
      const uint32_t a = uint32_t(uint64_t(n) * t / T), z = uint32_t(uint64_t(n) * (t + 1) / T);
      std::memcpy(e->h_stage_nodes.p + a, host_nodes + a, sizeof(grove_node_t) * (z - a));
      if (cached) {
        const uint32_t* rd = e->raw_dom.data();
        int d = 0;
        for (uint32_t i = a; i < z; ++i)
          d |= std::memcmp(rd + size_t(i) * GROVE_MAX_LEVELS, host_nodes[i].dom, sizeof(uint32_t) * GROVE_MAX_LEVELS) != 0;
        differs |= d;
      }
    }
    if (!cached || differs) { int32_t rc = build_topology(e, host_nodes, n); if (rc) return rc; }
    CU_TRY(e, cudaMemcpyAsync(e->d_nodes_in.p, e->h_stage_nodes.p, sizeof(grove_node_t) * n, cudaMemcpyHostToDevice, e->stream));
    CU_TRY(e, cudaEventRecord(e->ev_nodes_up, e->stream));
  } else {
    if (!e->nodes_loaded || n != e->N) return fail(e, GROVE_ERR_STATE, "device load needs a prior host load with the same labels");
    CU_TRY(e, cudaMemcpyAsync(e->d_nodes_in.p, dev_nodes, sizeof(grove_node_t) * n, cudaMemcpyDeviceToDevice, e->stream));
  }
  k_gather<<<(e->Npad + 255) / 256, 256, 0, e->stream>>>(e->d_nodes_in.p, e->d_perm.p, e->d_vdepth.p, e->d_nres.p, e->N, e->Npad);
  CU_TRY(e, cudaGetLastError());
  if (dev_nodes) CU_TRY(e, cudaStreamSynchronize(e->stream));  // the caller's device buffer is free again on return
  e->nodes_loaded = true;
  return GROVE_OK;
}

int32_t grove_load_nodes(grove_engine_t* e, const grove_node_t* nodes, uint32_t n) {
  if (!e || !nodes) return GROVE_ERR_INVALID_ARG;
  return load_nodes_common(e, nodes, nullptr, n);
}

int32_t grove_load_nodes_device(grove_engine_t* e, const void* d_nodes, uint32_t n) {
  if (!e || !d_nodes) return GROVE_ERR_INVALID_ARG;
  return load_nodes_common(e, nullptr, d_nodes, n);
}

int32_t grove_update_nodes(grove_engine_t* e, const uint32_t* idx, const grove_node_t* recs, uint32_t n) {
  if (!e || (n && (!idx || !recs))) return GROVE_ERR_INVALID_ARG;
  if (!e->nodes_loaded) return fail(e, GROVE_ERR_STATE, "no node snapshot loaded");
  if (e->in_cycle) return fail(e, GROVE_ERR_STATE, "cycle in flight");
  if (n == 0) return GROVE_OK;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  std::vector<uint32_t> sidx(n);
  for (uint32_t i = 0; i < n; ++i) {
    if (idx[i] >= e->N) return fail(e, GROVE_ERR_INVALID_ARG, "node index out of range");
    if (std::memcmp(&e->raw_dom[size_t(idx[i]) * GROVE_MAX_LEVELS], recs[i].dom, sizeof(uint32_t) * GROVE_MAX_LEVELS) != 0)
      return fail(e, GROVE_ERR_INVALID_ARG, "grove_update_nodes cannot change labels; reload the snapshot");
    sidx[i] = e->inv[idx[i]];
  }
  CU_TRY(e, e->d_upd_idx.ensure(n)); CU_TRY(e, e->d_upd_recs.ensure(n));
  CU_TRY(e, cudaMemcpyAsync(e->d_upd_idx.p, sidx.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaMemcpyAsync(e->d_upd_recs.p, recs, sizeof(grove_node_t) * n, cudaMemcpyHostToDevice, e->stream));
  k_update<<<(n + 255) / 256, 256, 0, e->stream>>>(e->d_upd_idx.p, e->d_upd_recs.p, e->d_vdepth.p, e->d_perm.p, e->d_nres.p, e->d_nodes_in.p, n);
  CU_TRY(e, cudaGetLastError());
  CU_TRY(e, cudaStreamSynchronize(e->stream));  // sidx is a local, recs the caller's
  return GROVE_OK;
}

int32_t grove_get_nodes(grove_engine_t* e, grove_node_t* out, uint32_t cap) {
  if (!e || !out) return GROVE_ERR_INVALID_ARG;
  if (!e->nodes_loaded) return fail(e, GROVE_ERR_STATE, "no node snapshot loaded");
  if (cap < e->N) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, e->d_nodes_out.ensure(e->N));
  k_scatter<<<(e->N + 255) / 256, 256, 0, e->stream>>>(e->d_nodes_out.p, e->d_nodes_in.p, e->d_perm.p, e->d_nres.p, e->N);
  CU_TRY(e, cudaGetLastError());
  CU_TRY(e, cudaMemcpyAsync(out, e->d_nodes_out.p, sizeof(grove_node_t) * e->N, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  return GROVE_OK;
}

// one gang; returns nullptr when it is well formed, else the reason (code through *code)
static const char* validate_gang(const grove_engine* e, const grove_gang_t* gangs, uint32_t G, const grove_clique_t* cliques, uint32_t Q,
                                 const grove_scope_t* scopes, uint32_t S, uint32_t gi, int32_t* code) {
  const uint32_t L = e->L;
  const grove_gang_t& g = gangs[gi];
  *code = GROVE_ERR_LIMIT;
  if (g.n_cliques == 0 || g.n_cliques > GROVE_MAX_GANG_CLIQUES) return "gang clique count out of range";
  if (g.n_scopes == 0 || g.n_scopes > GROVE_MAX_GANG_SCOPES) return "gang scope count out of range";
  *code = GROVE_ERR_INVALID_ARG;
  if (uint64_t(g.clique_off) + g.n_cliques > Q || uint64_t(g.scope_off) + g.n_scopes > S) return "gang table offsets out of range";
  if (g.level != GROVE_LEVEL_NONE && g.level >= L) return "gang level out of range";
  if (g.preferred != GROVE_LEVEL_NONE && (g.preferred >= L || (g.level != GROVE_LEVEL_NONE && g.preferred <= g.level)))
    return "gang preferred level must be a level deeper than the required one";
  if (g.anchor_node != GROVE_NONE_U32 && e->nodes_loaded && g.anchor_node >= e->N) return "anchor node out of range";
  if (g.base_gang != GROVE_NONE_U32 && (g.base_gang >= G || g.base_gang == gi)) return "base gang out of range";
  uint32_t pods = 0, next = 0;
  for (uint32_t si = 0; si < g.n_scopes; ++si) {
    const grove_scope_t& s = scopes[g.scope_off + si];
    if (s.first_clique != next || s.n_cliques == 0) return "scopes must tile the gang's cliques in order";
    if (s.level != GROVE_LEVEL_NONE && s.level >= L) return "scope level out of range";
    if (s.preferred1 && (s.preferred1 > L || (s.level != GROVE_LEVEL_NONE && uint32_t(s.preferred1) - 1u <= s.level)))
      return "scope preferred level must be a level deeper than the required one";
    if (next + s.n_cliques > g.n_cliques) return "scope exceeds gang";
    for (uint32_t i = 0; i < s.n_cliques; ++i) {
      const grove_clique_t& q = cliques[g.clique_off + next + i];
      if (GROVE_CLIQUE_SCOPE(q.scope) != si) return "clique.scope does not match its scope";
      if (q.level != GROVE_LEVEL_NONE && q.level >= L) return "clique level out of range";
      const uint32_t qp = GROVE_CLIQUE_PREFERRED(q.scope);
      if (qp != GROVE_LEVEL_NONE && (qp >= L || (q.level != GROVE_LEVEL_NONE && qp <= q.level)))
        return "clique preferred level must be a level deeper than the required one";
      if (q.replicas < q.min_replicas) return "replicas < min_replicas";
      pods += q.replicas;
    }
    next += s.n_cliques;
  }
  if (next != g.n_cliques) return "scopes do not cover the gang";
  if (pods > GROVE_MAX_GANG_PODS) { *code = GROVE_ERR_LIMIT; return "gang exceeds GROVE_MAX_GANG_PODS"; }
  return nullptr;
}

// same structural rules as the PodGang admission webhook guarantees
// (operator/internal/webhook/admission/pcs/validation/topologyconstraints.go:195-280); the first offending gang
// (lowest index) is reported whatever the thread count
static int32_t validate(grove_engine* e, const grove_gang_t* gangs, uint32_t G, const grove_clique_t* cliques, uint32_t Q,
                        const grove_scope_t* scopes, uint32_t S) {
  uint32_t first_bad = GROVE_NONE_U32;
  const int T = G >= 2048 ? host_threads() : 1;
#pragma omp parallel for num_threads(T) schedule(static) reduction(min : first_bad)
  for (uint32_t gi = 0; gi < G; ++gi) {
    int32_t code;
    if (gi < first_bad && validate_gang(e, gangs, G, cliques, Q, scopes, S, gi, &code)) first_bad = gi;
  }
  if (first_bad == GROVE_NONE_U32) return GROVE_OK;
  int32_t code = GROVE_ERR_INVALID_ARG;
  const char* why = validate_gang(e, gangs, G, cliques, Q, scopes, S, first_bad, &code);
  return fail(e, code, why ? why : "malformed gang tables");
}

int32_t grove_submit_gangs(grove_engine_t* e, const grove_gang_t* gangs, uint32_t n_gangs, const grove_clique_t* cliques,
                           uint32_t n_cliques, const grove_scope_t* scopes, uint32_t n_scopes) {
  if (!e || (n_gangs && (!gangs || !cliques || !scopes))) return GROVE_ERR_INVALID_ARG;
  if (e->in_cycle) return fail(e, GROVE_ERR_STATE, "cycle in flight");
  if (n_gangs >= (1u << 24)) return fail(e, GROVE_ERR_LIMIT, "too many gangs");  // order rank + sub-round tag share a claim word
  int32_t rc = validate(e, gangs, n_gangs, cliques, n_cliques, scopes, n_scopes);
  if (rc) return rc;
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, cudaEventSynchronize(e->ev_tables_up));  // an earlier upload may still be reading the pinned tables
  e->G = n_gangs; e->Q = n_cliques; e->S = n_scopes;
  CU_TRY(e, e->gangs.assign(gangs, n_gangs));
  CU_TRY(e, e->cliques.assign(cliques, n_cliques));
  CU_TRY(e, e->scopes.assign(scopes, n_scopes));
  e->gangs_loaded = true; e->ginfo_dirty = true; e->have_results = false;
  e->n_constrained = e->n_unconstrained = 0;
  bool pref = false;
  for (uint32_t g = 0; g < n_gangs; ++g) {
    (gangs[g].level == GROVE_LEVEL_NONE && gangs[g].preferred == GROVE_LEVEL_NONE ? e->n_unconstrained : e->n_constrained)++;
    pref |= gangs[g].preferred != GROVE_LEVEL_NONE;
  }
  for (uint32_t i = 0; i < n_scopes && !pref; ++i) pref |= scopes[i].preferred1 != 0;
  for (uint32_t i = 0; i < n_cliques && !pref; ++i) pref |= (cliques[i].scope >> 5) != 0;
  e->any_preferred = pref;
  CU_TRY(e, e->d_gangs.ensure(n_gangs)); CU_TRY(e, e->d_cliques.ensure(n_cliques)); CU_TRY(e, e->d_scopes.ensure(n_scopes));
  if (n_gangs) CU_TRY(e, cudaMemcpyAsync(e->d_gangs.p, e->gangs.data(), sizeof(grove_gang_t) * n_gangs, cudaMemcpyHostToDevice, e->stream));
  if (n_cliques) CU_TRY(e, cudaMemcpyAsync(e->d_cliques.p, e->cliques.data(), sizeof(grove_clique_t) * n_cliques, cudaMemcpyHostToDevice, e->stream));
  if (n_scopes) CU_TRY(e, cudaMemcpyAsync(e->d_scopes.p, e->scopes.data(), sizeof(grove_scope_t) * n_scopes, cudaMemcpyHostToDevice, e->stream));
  CU_TRY(e, cudaEventRecord(e->ev_tables_up, e->stream));
  return GROVE_OK;  // the uploads read engine-owned pinned memory: nothing of the caller's is referenced any more
}

// derived per-gang / per-clique tables: order rank, anchor (sorted index + ancestor ranges), entry slots
static int32_t build_ginfo(grove_engine* e) {
  const uint32_t G = e->G, Q = e->Q;
  const auto t_b0 = std::chrono::steady_clock::now();
  CU_TRY(e, e->ginfo_pin.ensure(G)); CU_TRY(e, e->cinfo_pin.ensure(Q));
  e->ginfo = e->ginfo_pin.p; e->cinfo = e->cinfo_pin.p;
  {  // ginfo zeroed, cinfo all-ones (gang == NONE marks a row no gang owns yet), by the engine's thread team
    const int T0 = G >= 2048 ? host_threads() : 1;
#pragma omp parallel for num_threads(T0) schedule(static)
    for (int t = 0; t < T0; ++t) {
      const size_t g0 = size_t(G) * t / T0, g1 = size_t(G) * (t + 1) / T0, q0 = size_t(Q) * t / T0, q1 = size_t(Q) * (t + 1) / T0;
      std::memset(e->ginfo + g0, 0, sizeof(GangInfo) * (g1 - g0));
      std::memset(e->cinfo + q0, 0xFF, sizeof(CliqueInfo) * (q1 - q0));
    }
  }
  e->sigs.clear();
  const auto t_m0 = std::chrono::steady_clock::now();
  // order rank = position by (priority desc, index asc), and the anchors.  PriorityClasses are few: every thread counts
  // the distinct priorities of its gang range, one thread turns the counts into start offsets per (priority, thread),
  // and every thread numbers its own range -- a stable counting sort without the sorted array.
  bool bad_anchor = false;
  {
    constexpr size_t kMaxPrio = 64;
    const int T1 = G >= 2048 ? host_threads() : 1;
    std::vector<std::vector<std::pair<int32_t, uint32_t>>> seen(T1);   // per thread: (priority, count) in first-seen order
    std::vector<int32_t> prios;                                         // distinct, descending
    std::vector<uint32_t> start;                                        // [thread][priority index] first rank
    bool many = false;
#pragma omp parallel num_threads(T1)
    {
      const int t = omp_get_thread_num();
      const uint32_t g0 = uint32_t(uint64_t(G) * t / T1), g1 = uint32_t(uint64_t(G) * (t + 1) / T1);
      auto& mine = seen[t];
      bool over = false;
      for (uint32_t g = g0; g < g1 && !over; ++g) {
        const int32_t p = e->gangs[g].priority;
        size_t i = 0;
        while (i < mine.size() && mine[i].first != p) ++i;
        if (i == mine.size()) { if (mine.size() == kMaxPrio) { over = true; break; } mine.push_back({p, 0u}); }
        mine[i].second++;
      }
      if (over) {
#pragma omp atomic write
        many = true;
      }
#pragma omp barrier
#pragma omp single
      {
        for (const auto& v : seen) for (const auto& pc : v) prios.push_back(pc.first);
        std::sort(prios.begin(), prios.end(), std::greater<int32_t>());
        prios.erase(std::unique(prios.begin(), prios.end()), prios.end());
        if (prios.size() > kMaxPrio) many = true;
        if (!many) {
          start.assign(size_t(T1) * prios.size(), 0u);
          uint32_t run = 0;
          for (size_t k = 0; k < prios.size(); ++k)
            for (int tt = 0; tt < T1; ++tt) {
              start[size_t(tt) * prios.size() + k] = run;
              for (const auto& pc : seen[tt]) if (pc.first == prios[k]) run += pc.second;
            }
        }
      }  // implicit barrier
      bool bad = false;
      uint32_t next[kMaxPrio];
      if (!many) for (size_t k = 0; k < prios.size(); ++k) next[k] = start[size_t(t) * prios.size() + k];
      for (uint32_t gi = g0; gi < g1; ++gi) {
        const grove_gang_t& g = e->gangs[gi];
        if (!many) {
          size_t k = 0;
          while (prios[k] != g.priority) ++k;
          e->ginfo[gi].order = next[k]++;
        }
        if (g.anchor_node != GROVE_NONE_U32 && g.anchor_node >= e->N) { bad = true; continue; }
        e->ginfo[gi].anchor = g.anchor_node != GROVE_NONE_U32 ? e->inv[g.anchor_node] : fmix32(gi) % e->N;
      }
      if (bad) {
#pragma omp atomic write
        bad_anchor = true;
      }
    }
    if (many) {  // more distinct priorities than a scheduler has PriorityClasses: plain stable sort
      std::vector<uint32_t> ord(G);
      std::iota(ord.begin(), ord.end(), 0u);
      std::stable_sort(ord.begin(), ord.end(), [e](uint32_t a, uint32_t b) { return e->gangs[a].priority > e->gangs[b].priority; });
      for (uint32_t r = 0; r < G; ++r) e->ginfo[ord[r]].order = r;
    }
  }
  if (bad_anchor) return fail(e, GROVE_ERR_INVALID_ARG, "anchor node out of range");
  const auto t_m1 = std::chrono::steady_clock::now();
  // per-clique derived data + signature interning.  PodCliques stamped from one template (PCS / PCSG
  // replicas) share requests, selector class and binding depth: they share one fit-bitmap row.  Gang
  // ranges are processed by a few host threads with thread-local signature tables, merged afterwards.
  const auto t_m2 = std::chrono::steady_clock::now();
  struct SigTab {
    std::vector<std::array<uint32_t, 6>> tab = std::vector<std::array<uint32_t, 6>>(256, std::array<uint32_t, 6>{0, 0, 0, 0, 0, GROVE_NONE_U32});
    std::vector<std::array<uint32_t, 5>> keys;
    static uint32_t hash5(const std::array<uint32_t, 5>& k) { uint32_t h = k[0] * 0x9E3779B1u ^ k[1] * 0x85EBCA6Bu ^ (k[2] << 20) ^ (k[3] << 4) ^ k[4]; return h ^ (h >> 15); }
    uint32_t intern(const std::array<uint32_t, 5>& key) {
      if (keys.size() * 2 >= tab.size()) {
        std::vector<std::array<uint32_t, 6>> nt(tab.size() * 4, std::array<uint32_t, 6>{0, 0, 0, 0, 0, GROVE_NONE_U32});
        for (const auto& r : tab) if (r[5] != GROVE_NONE_U32) {
          size_t h = hash5({r[0], r[1], r[2], r[3], r[4]}) & (nt.size() - 1);
          while (nt[h][5] != GROVE_NONE_U32) h = (h + 1) & (nt.size() - 1);
          nt[h] = r;
        }
        tab.swap(nt);
      }
      size_t h = hash5(key) & (tab.size() - 1);
      while (tab[h][5] != GROVE_NONE_U32 && !(tab[h][0] == key[0] && tab[h][1] == key[1] && tab[h][2] == key[2] && tab[h][3] == key[3] && tab[h][4] == key[4]))
        h = (h + 1) & (tab.size() - 1);
      if (tab[h][5] == GROVE_NONE_U32) { tab[h] = {key[0], key[1], key[2], key[3], key[4], uint32_t(keys.size())}; keys.push_back(key); }
      return tab[h][5];
    }
  };
  const int T = G >= 2048 ? host_threads() : 1;
  std::vector<SigTab> local(T);
  std::vector<uint32_t> gang_pods(G, 0);
  int shared_rows = 0;
#pragma omp parallel num_threads(T) reduction(+ : shared_rows)
  {
    const int t = omp_get_thread_num();
    const uint32_t g0 = uint32_t(uint64_t(G) * t / T), g1 = uint32_t(uint64_t(G) * (t + 1) / T);
    SigTab& st = local[t];
    for (uint32_t gi = g0; gi < g1; ++gi) {
      const grove_gang_t& g = e->gangs[gi];
      uint32_t pods = 0;
      for (uint32_t si = 0; si < g.n_scopes; ++si) {
        const grove_scope_t& s = e->scopes[g.scope_off + si];
        for (uint32_t i = 0; i < s.n_cliques; ++i) {
          const uint32_t qi = g.clique_off + s.first_clique + i;
          const grove_clique_t& q = e->cliques[qi];
          uint32_t nd = 0;
          if (g.level != GROVE_LEVEL_NONE) nd = std::max(nd, uint32_t(g.level) + 1);
          if (s.level != GROVE_LEVEL_NONE) nd = std::max(nd, uint32_t(s.level) + 1);
          if (q.level != GROVE_LEVEL_NONE) nd = std::max(nd, uint32_t(q.level) + 1);
          if (e->cinfo[qi].gang != GROVE_NONE_U32) shared_rows++;  // rows of different threads never overlap unless the tables are malformed
          e->cinfo[qi] = CliqueInfo{gi, nd, st.intern({q.req_cpu_milli, q.req_mem_mib, q.req_gpu, q.class_mask, nd}), uint32_t(t)};
          pods += q.replicas;
        }
      }
      gang_pods[gi] = pods;
    }
  }
  const auto t_m3 = std::chrono::steady_clock::now();
  if (shared_rows) return fail(e, GROVE_ERR_INVALID_ARG, "clique rows shared between gangs");
  SigTab global;
  std::vector<std::vector<uint32_t>> remap(T);
  for (int t = 0; t < T; ++t) { remap[t].resize(local[t].keys.size()); for (size_t k = 0; k < local[t].keys.size(); ++k) remap[t][k] = global.intern(local[t].keys[k]); }
  e->sigs.resize(global.keys.size());
  for (size_t k = 0; k < global.keys.size(); ++k) { const auto& key = global.keys[k]; e->sigs[k] = make_uint4(key[0], key[1], key[2], key[3] | (key[4] << 16)); }
  // thread-local signature ids -> global ids (pad carried the id of the thread that interned the row)
#pragma omp parallel for num_threads(T) schedule(static)
  for (uint32_t qi = 0; qi < Q; ++qi)
    if (e->cinfo[qi].gang != GROVE_NONE_U32) { e->cinfo[qi].sig = remap[e->cinfo[qi].pad][e->cinfo[qi].sig]; e->cinfo[qi].pad = 0; }
  uint32_t pod_off = 0;
  e->max_gang_pods = 0;
  for (uint32_t gi = 0; gi < G; ++gi) { e->ginfo[gi].pod_off = pod_off; pod_off += gang_pods[gi]; e->max_gang_pods = std::max(e->max_gang_pods, gang_pods[gi]); }
  e->P = pod_off;
  const auto t_b1 = std::chrono::steady_clock::now();
  for (uint32_t qi = 0; qi < Q; ++qi)
    if (e->cinfo[qi].gang == GROVE_NONE_U32) return fail(e, GROVE_ERR_INVALID_ARG, "clique row owned by no gang");
  e->n_sigs = uint32_t(e->sigs.size());
  CU_TRY(e, e->d_ginfo.ensure(G)); CU_TRY(e, e->d_cinfo.ensure(Q)); CU_TRY(e, e->d_sigs.ensure(e->n_sigs));
  CU_TRY(e, e->d_sig_stamp.ensure(e->n_sigs)); CU_TRY(e, e->d_sig_list.ensure(e->n_sigs));
  if (e->n_sigs) CU_TRY(e, cudaMemcpyAsync(e->d_sigs.p, e->sigs.data(), sizeof(uint4) * e->n_sigs, cudaMemcpyHostToDevice, e->stream));
  if (G) {
    CU_TRY(e, cudaMemcpyAsync(e->d_ginfo.p, e->ginfo, sizeof(GangInfo) * G, cudaMemcpyHostToDevice, e->stream));
    k_anchor<<<(G + 255) / 256, 256, 0, e->stream>>>(make_topo(e), e->d_ginfo.p, G);  // ancestor ranges from the device-resident tree
    CU_TRY(e, cudaGetLastError());
  }
  if (Q) CU_TRY(e, cudaMemcpyAsync(e->d_cinfo.p, e->cinfo, sizeof(CliqueInfo) * Q, cudaMemcpyHostToDevice, e->stream));
  e->ginfo_dirty = false;
  if (std::getenv("GROVE_DEBUG_HOST")) {
    const auto t_b2 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    std::fprintf(stderr, "build_ginfo: tables %ld us (init %ld, order %ld, anchors %ld, cliques %ld, merge %ld), upload %ld us\n", us(t_b0, t_b1), us(t_b0, t_m0), us(t_m0, t_m1), us(t_m1, t_m2), us(t_m2, t_m3), us(t_m3, t_b1), us(t_b1, t_b2));
  }
  return GROVE_OK;
}

int32_t grove_cycle_begin(grove_engine_t* e) {
  if (!e) return GROVE_ERR_INVALID_ARG;
  if (!e->nodes_loaded) return fail(e, GROVE_ERR_STATE, "no node snapshot loaded");
  if (!e->gangs_loaded) return fail(e, GROVE_ERR_STATE, "no gangs submitted");
  if (e->in_cycle) return fail(e, GROVE_ERR_STATE, "cycle in flight");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  if (e->ginfo_dirty) { int32_t rc = build_ginfo(e); if (rc) return rc; }
  const uint32_t G = e->G, Q = e->Q;
  CU_TRY(e, e->d_state.ensure(G)); CU_TRY(e, e->d_round.ensure(G)); CU_TRY(e, e->d_active.ensure(G)); CU_TRY(e, e->d_rows.ensure(Q));
  CU_TRY(e, e->d_counters.ensure(8)); CU_TRY(e, e->d_spec_score.ensure(G));
  CU_TRY(e, e->d_spec_n.ensure(G)); CU_TRY(e, e->d_spec_top.ensure(G)); CU_TRY(e, e->d_ent_node.ensure(e->P)); CU_TRY(e, e->d_ent_meta.ensure(e->P));
  CU_TRY(e, e->d_claim.ensure(e->N)); CU_TRY(e, e->d_totals.ensure(4));
  CU_TRY(e, e->d_taken.ensure(e->N)); CU_TRY(e, e->d_cur.ensure(G)); CU_TRY(e, e->d_prop.ensure(G)); CU_TRY(e, e->d_flags.ensure(GROVE_SUBROUNDS));
  CU_TRY(e, e->d_active_all.ensure(G));
  CU_TRY(e, e->d_xbuf.ensure(2 * size_t(e->K) * e->P + 4 * size_t(G) * e->K + G));
  CU_TRY(e, cudaMemsetAsync(e->d_flags.p, 0, sizeof(uint32_t) * GROVE_SUBROUNDS, e->stream));  // round-stamped; rounds count from 1
  CU_TRY(e, e->d_status.ensure(G)); CU_TRY(e, e->d_out.ensure(e->P));
  CU_TRY(e, e->h_status.ensure(G)); CU_TRY(e, e->h_out.ensure(e->P));
  // the Q x N matrices
  {
    const size_t fw = size_t(e->n_sigs) * e->words, tb = size_t(Q) * e->Npad;
    if (e->d_F.ensure(fw) != cudaSuccess || e->d_T.ensure(tb) != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(e, GROVE_ERR_OOM, "fit/score matrices do not fit in device memory");
    }
  }
  // capacity tables for the candidate pre-filter: worth building only while signatures are few
  e->prefilter = false;
  if (e->tune_prefilter && e->n_sigs && uint64_t(e->n_sigs) * e->Npad <= (64ull << 20) && e->cap_stride > 0) {
    const size_t tw = size_t(e->n_sigs) * e->cap_stride;
    if (e->d_cap8.ensure(size_t(e->n_sigs) * e->Npad) == cudaSuccess && e->d_capsum.ensure(tw) == cudaSuccess &&
        e->d_capmax.ensure(tw) == cudaSuccess) e->prefilter = true;
    else (void)cudaGetLastError();
  }
  if (e->dbg_on) CU_TRY(e, e->d_dbg.ensure(size_t(G) * 8));
  // initial states: gated gangs are skipped (pods still hold the scheduling gate, pod.go:70,164)
  CU_TRY(e, e->h_state0.ensure(G));
  uint8_t* st = e->h_state0.p;   // rewritten only by the next cycle_begin, long after this upload has been consumed
  for (uint32_t g = 0; g < G; ++g) st[g] = (e->gangs[g].flags & GROVE_GANG_GATED) ? GROVE_GANG_GATED_SKIP : GROVE_GANG_PENDING;
  if (G) CU_TRY(e, cudaMemcpyAsync(e->d_state.p, st, G, cudaMemcpyHostToDevice, e->stream));
  if (G) CU_TRY(e, cudaMemsetAsync(e->d_round.p, 0, G, e->stream));
  if (G) CU_TRY(e, cudaMemsetAsync(e->d_spec_n.p, 0, sizeof(uint16_t) * G, e->stream));
  if (e->n_sigs) CU_TRY(e, cudaMemsetAsync(e->d_sig_stamp.p, 0, sizeof(uint32_t) * e->n_sigs, e->stream));
  e->round_no = 0; e->pairs = 0; e->launches = 0; e->in_cycle = true; e->have_results = false;
  std::memset(&e->last, 0, sizeof(e->last));
  return GROVE_OK;
}

static size_t xbuf_words(const grove_engine* e) { return 2 * size_t(e->K) * e->P + 4 * size_t(e->G) * e->K + e->G; }

// K3 launches of one round.  kPref: some gang / scope / clique of the submission carries a Preferred level
// (the level walks are compiled in); otherwise the loop-free Required-only instantiations run.
extern "C++" {
template <bool kPref>
static void launch_admit(grove_engine* e, const Topo& tp, const Tables& tb, const RoundBufs& rb, uint32_t na, bool caps, bool small) {
  if (e->n_constrained) {
    if (na >= e->tune_warp_min) {  // throughput round: a warp per gang
      const uint32_t nb = (na + kAdmitWarpGangs - 1) / kAdmitWarpGangs;
      const int th = kAdmitWarpGangs * 32;
      if (caps && small) k_admit_warp<true, kEntSmem, kPref><<<nb, th, 0, e->stream>>>(tp, tb, rb);
      else if (caps) k_admit_warp<true, 0, kPref><<<nb, th, 0, e->stream>>>(tp, tb, rb);
      else if (small) k_admit_warp<false, kEntSmem, kPref><<<nb, th, 0, e->stream>>>(tp, tb, rb);
      else k_admit_warp<false, 0, kPref><<<nb, th, 0, e->stream>>>(tp, tb, rb);
    } else if (na >= e->tune_wide_max) {  // middle: a 4-warp CTA per gang
      if (caps && small) k_admit<kAdmitThreads, 0, kEntSmem, kPref><<<na, kAdmitThreads, 0, e->stream>>>(tp, tb, rb);
      else if (caps) k_admit<kAdmitThreads, 0, 0, kPref><<<na, kAdmitThreads, 0, e->stream>>>(tp, tb, rb);
      else if (small) k_admit<kAdmitThreads, 1, kEntSmem, kPref><<<na, kAdmitThreads, 0, e->stream>>>(tp, tb, rb);
      else k_admit<kAdmitThreads, 1, 0, kPref><<<na, kAdmitThreads, 0, e->stream>>>(tp, tb, rb);
    } else {                // latency round: an 8-warp CTA per gang
      if (caps && small) k_admit<kAdmitThreadsWide, 0, kEntSmem, kPref><<<na, kAdmitThreadsWide, 0, e->stream>>>(tp, tb, rb);
      else if (caps) k_admit<kAdmitThreadsWide, 0, 0, kPref><<<na, kAdmitThreadsWide, 0, e->stream>>>(tp, tb, rb);
      else if (small) k_admit<kAdmitThreadsWide, 1, kEntSmem, kPref><<<na, kAdmitThreadsWide, 0, e->stream>>>(tp, tb, rb);
      else k_admit<kAdmitThreadsWide, 1, 0, kPref><<<na, kAdmitThreadsWide, 0, e->stream>>>(tp, tb, rb);
    }
    e->launches += 1;
  }
  if (e->n_unconstrained) { k_admit<kAdmitThreads, 2, 0, kPref><<<na, kAdmitThreads, 0, e->stream>>>(tp, tb, rb); e->launches += 1; }
}
}  // extern "C++"

// evaluation half of a round on this handle's share of the gangs: prepare -> fit -> (capacity tables)
// -> score -> admit.  Counters land in h_counters.
static int32_t round_eval(grove_engine* e, bool timed) {
  const Topo tp = make_topo(e); const Tables tb = make_tables(e); const RoundBufs rb = make_bufs(e);
  e->round_no++;
  if (e->G == 0) { std::memset(e->h_counters.p, 0, sizeof(uint32_t) * 8); return GROVE_OK; }  // empty submission: nothing to launch
  CU_TRY(e, cudaMemsetAsync(e->d_counters.p, 0, sizeof(uint32_t) * 8, e->stream));
  k_prepare<<<(e->G + 1023) / 1024, 1024, 0, e->stream>>>(tb, rb, e->round_no, e->cfg.rank, e->cfg.world);
  CU_TRY(e, cudaGetLastError());
  CU_TRY(e, cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  e->launches += 1;
  const uint32_t na = e->h_counters.p[0], nr = e->h_counters.p[1], ns = e->h_counters.p[4];
  if (e->cfg.world > 1 && e->h_counters.p[5])  // all-reduce SUM payload: everything this rank does not own stays zero
    CU_TRY(e, cudaMemsetAsync(e->d_xbuf.p, 0, sizeof(uint32_t) * xbuf_words(e), e->stream));
  if (na == 0) return GROVE_OK;
  if (timed) CU_TRY(e, cudaEventRecord(e->ev[0], e->stream));
  dim3 gfit(e->Npad / 1024, std::min<uint32_t>((ns + kFitTile - 1) / kFitTile, 65535u));
  k_fit<<<gfit, 1024, 0, e->stream>>>(tp, tb, rb);
  // fork: the score matrix goes to the second stream and overlaps the admission
  cudaStream_t ss = e->tune_overlap ? e->stream_score : e->stream;
  if (e->tune_overlap) {
    CU_TRY(e, cudaEventRecord(e->ev_fit, e->stream));
    CU_TRY(e, cudaStreamWaitEvent(e->stream_score, e->ev_fit, 0));
  }
  if (timed) CU_TRY(e, cudaEventRecord(e->ev_s0, ss));
  k_score<<<dim3(1, std::min<uint32_t>(nr, 65535u)), 256, 0, ss>>>(tp, tb, rb, nr);  // one CTA per row
  if (timed) CU_TRY(e, cudaEventRecord(e->ev_s1, ss));
  if (e->tune_overlap) CU_TRY(e, cudaEventRecord(e->ev_score, e->stream_score));
  if (e->prefilter && ns) {
    k_cap8<<<dim3(e->Npad / 256, ns), 256, 0, e->stream>>>(tp, tb, rb, e->d_cap8.p);
    k_capsum<<<dim3((e->cap_stride * 32 + 255) / 256, ns), 256, 0, e->stream>>>(tp, rb, e->d_cap8.p, e->d_capsum.p, e->d_capmax.p);
    e->launches += 2;
  }
  if (timed) CU_TRY(e, cudaEventRecord(e->ev[1], e->stream));
  if (timed) CU_TRY(e, cudaEventRecord(e->ev[2], e->stream));
  if (e->dbg_on) { cudaMemsetAsync(e->d_dbg.p, 0, size_t(e->G) * 32, e->stream); k_dbg_init<<<(e->G + 255) / 256, 256, 0, e->stream>>>(e->d_dbg.p, e->G); }
  const bool caps = e->prefilter && e->tune_prefilter >= 2;
  const bool small = e->max_gang_pods <= kEntSmem;  // per-lane entry stacks fit the shared-memory form
  if (e->any_preferred) launch_admit<true>(e, tp, tb, rb, na, caps, small); else launch_admit<false>(e, tp, tb, rb, na, caps, small);
  // join: scores of the alternatives need both the score matrix and the alternatives
  if (e->tune_overlap) CU_TRY(e, cudaStreamWaitEvent(e->stream, e->ev_score, 0));
  k_alt_scores<<<(na * e->K * 32 + 255) / 256, 256, 0, e->stream>>>(tp, tb, rb);
  CU_TRY(e, cudaGetLastError());
  if (timed) CU_TRY(e, cudaEventRecord(e->ev[3], e->stream));
  e->launches += 3;
  e->pairs += uint64_t(nr) * e->N;
  return GROVE_OK;
}

// resolution half: sub-rounds over the alternatives of every rank's active gangs (replicated), commits
static int32_t round_resolve(grove_engine* e, bool timed) {
  const uint32_t na_all = e->h_counters.p[5];
  if (na_all == 0) return GROVE_OK;
  Topo tp = make_topo(e); Tables tb = make_tables(e); RoundBufs rb = make_bufs(e);
  // The per-round scratch is stamped instead of cleared.  claim[n] carries a tag in its top byte that DEcreases with
  // every (round, sub-round), so a newer claim always beats a stale one through atomicMin; taken[n] holds the stamp of
  // the round that committed on n; flags[sub] the number of the round that proposed.  The arrays are reset only when a
  // stamp would wrap: every kClaimRounds rounds (claim: the top byte stays below 0x7F = "no claim") / 254 rounds (taken).
  const uint32_t r0 = e->round_no - 1;   // rounds count from 1
  if (r0 % kClaimRounds == 0) CU_TRY(e, cudaMemsetAsync(e->d_claim.p, 0x7F, sizeof(uint32_t) * e->N, e->stream));
  if (r0 % 254u == 0) CU_TRY(e, cudaMemsetAsync(e->d_taken.p, 0, e->N, e->stream));
  uint32_t tag_hi = 0x7Eu - (r0 % kClaimRounds) * GROVE_SUBROUNDS;   // tag of sub-round 0 of this round
  uint32_t tk = r0 % 254u + 1u;                                       // 1..254
  uint4* nres = e->d_nres.p; uint32_t rn = e->round_no;
  const uint32_t want = (na_all * 32 + kResolveThreads - 1) / kResolveThreads;  // a warp per gang
  const uint32_t blocks = std::max(1u, std::min<uint32_t>(want, std::min<uint32_t>(uint32_t(e->resolve_blocks_per_sm), e->tune_resolve_bps) * e->n_sm));
  void* args[] = {&tp, &tb, &rb, &nres, &rn, &tag_hi, &tk};
  CU_TRY(e, cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_resolve), dim3(blocks), dim3(kResolveThreads), args, 0, e->stream));
  if (timed) CU_TRY(e, cudaEventRecord(e->ev[4], e->stream));
  e->launches += 1;
  return GROVE_OK;
}

static int32_t finish_cycle(grove_engine* e, grove_cycle_stats_t* stats) {
  const Topo tp = make_topo(e); const Tables tb = make_tables(e); const RoundBufs rb = make_bufs(e);
  k_finalize<<<1, 1024, 0, e->stream>>>(tp, tb, rb, e->d_status.p, e->d_totals.p);
  if (e->G) k_emit<<<(e->G * 32 + 255) / 256, 256, 0, e->stream>>>(tb, rb, e->d_perm.p, e->d_status.p, e->d_out.p);
  CU_TRY(e, cudaGetLastError());
  e->launches += 2;
  CU_TRY(e, cudaMemcpyAsync(e->h_counters.p, e->d_totals.p, sizeof(uint32_t) * 4, cudaMemcpyDeviceToHost, e->stream));
  if (e->G) CU_TRY(e, cudaMemcpyAsync(e->h_status.p, e->d_status.p, sizeof(grove_gang_status_t) * e->G, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  e->n_out = e->h_counters.p[0];
  if (e->n_out) CU_TRY(e, cudaMemcpyAsync(e->h_out.p, e->d_out.p, sizeof(grove_placement_t) * e->n_out, cudaMemcpyDeviceToHost, e->stream));
  CU_TRY(e, cudaStreamSynchronize(e->stream));
  e->last.rounds = e->round_no; e->last.gangs_admitted = e->h_counters.p[1]; e->last.gangs_rejected = e->h_counters.p[2];
  e->last.pods_bound = e->n_out; e->last.pairs_evaluated = e->pairs; e->last.kernel_launches = e->launches;
  e->have_results = true; e->in_cycle = false;
  if (stats) *stats = e->last;
  return GROVE_OK;
}

// what the prepare pass decided (identical on every rank: it sees the replicated gang state):
//  1 = a round is on, 0 = the cycle is over
static int32_t after_prepare(grove_engine* e, uint32_t* go) {
  *go = 0;
  const uint32_t unres = e->h_counters.p[2], glob = e->h_counters.p[5];
  if (unres == 0) { if (!e->h_counters.p[3]) e->round_no--; return GROVE_OK; }  // nothing left (and nothing propagated): not a round
  if (glob == 0) {  // dependency cycle / unreachable base: nothing can ever become active
    k_reject_rest<<<(e->G + 255) / 256, 256, 0, e->stream>>>(make_tables(e), make_bufs(e), e->round_no);
    CU_TRY(e, cudaGetLastError());
    e->launches += 1;
    return GROVE_OK;
  }
  *go = 1;
  return GROVE_OK;
}

int32_t grove_run_cycle(grove_engine_t* e, grove_cycle_stats_t* stats) {
  if (!e) return GROVE_ERR_INVALID_ARG;
  if (e->cfg.world > 1) return fail(e, GROVE_ERR_STATE, "sharded handle: drive the cycle with grove_round_* and reduce between the steps");
  const auto t_h0 = std::chrono::steady_clock::now();
  int32_t rc = grove_cycle_begin(e);
  const auto t_h1 = std::chrono::steady_clock::now();
  if (rc) return rc;
  float ms_fit = 0, ms_score = 0, ms_admit = 0, ms_commit = 0;
  CU_TRY(e, cudaEventRecord(e->ev[8], e->stream));
  for (;;) {
    if (e->cfg.max_rounds && e->round_no >= e->cfg.max_rounds) break;
    rc = round_eval(e, true);
    if (rc) { e->in_cycle = false; return rc; }
    uint32_t go = 0;
    rc = after_prepare(e, &go);
    if (rc) { e->in_cycle = false; return rc; }
    if (!go) break;
    const uint32_t na = e->h_counters.p[0];
    rc = round_resolve(e, true);
    if (rc) { e->in_cycle = false; return rc; }
    CU_TRY(e, cudaEventSynchronize(e->ev[4]));
    float t;
    cudaEventElapsedTime(&t, e->ev[0], e->ev[1]); ms_fit += t;
    cudaEventElapsedTime(&t, e->ev_s0, e->ev_s1); ms_score += t;  // on its own stream, overlapping the admission
    cudaEventElapsedTime(&t, e->ev[2], e->ev[3]); ms_admit += t;
    cudaEventElapsedTime(&t, e->ev[3], e->ev[4]); ms_commit += t;
    if (e->dbg_on) {  // GROVE_DEBUG_ADMIT: per-round admission statistics on stderr
      std::vector<uint32_t> h(size_t(e->G) * 8); std::vector<uint32_t> act(na);
      cudaMemcpy(h.data(), e->d_dbg.p, h.size() * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(act.data(), e->d_active.p, na * 4, cudaMemcpyDeviceToHost);
      uint64_t sp = 0, sa = 0, sk = 0, won = 0, maxa = 0, ss = 0, full = 0;
      for (uint32_t i = 0; i < na; ++i) {
        const uint32_t* d = &h[size_t(act[i]) * 8]; sp += d[1]; sa += d[2]; ss += d[0]; maxa = std::max<uint64_t>(maxa, d[2]);
        if (d[0] >= e->K) full++;
        if (d[3] != 0xFFFFFFFFu) { won++; sk += d[3]; }
      }
      std::fprintf(stderr, "round %u: active %u plausible/gang %.1f attempts/gang %.1f (max %llu) successes/gang %.1f gangs with K %llu feasible %llu mean first feasible k %.1f\n",
                   e->round_no, na, double(sp) / na, double(sa) / na, (unsigned long long)maxa, double(ss) / na, (unsigned long long)full, (unsigned long long)won,
                   won ? double(sk) / won : 0.0);
      // the slowest gangs of the round (SM cycles of their admission) and what they did
      std::vector<uint32_t> idx(na); std::iota(idx.begin(), idx.end(), 0u);
      std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return h[size_t(act[a]) * 8 + 4] > h[size_t(act[b]) * 8 + 4]; });
      uint64_t cyc = 0; for (uint32_t i = 0; i < na; ++i) cyc += h[size_t(act[i]) * 8 + 4];
      std::fprintf(stderr, "  mean cycles/gang %.0f, median %u; slowest:", double(cyc) / na, h[size_t(act[idx[na / 2]]) * 8 + 4]);
      for (uint32_t i = 0; i < std::min(na, 6u); ++i) {
        const uint32_t g = act[idx[i]]; const uint32_t* d = &h[size_t(g) * 8];
        std::fprintf(stderr, " [g%u lvl%u cliques%u: %u cyc, %u chunks, %u plaus, %u att, %u succ]", g, unsigned(e->gangs[g].level), unsigned(e->gangs[g].n_cliques), d[4], d[5], d[1], d[2], d[0]);
      }
      std::fprintf(stderr, "\n");
    }
  }
  CU_TRY(e, cudaEventRecord(e->ev[9], e->stream));
  const auto t_h2 = std::chrono::steady_clock::now();
  rc = finish_cycle(e, nullptr);
  if (std::getenv("GROVE_DEBUG_HOST")) {
    const auto t_h3 = std::chrono::steady_clock::now();
    auto us = [](auto a, auto b) { return (long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    std::fprintf(stderr, "host: begin %ld us, rounds %ld us, finish %ld us\n", us(t_h0, t_h1), us(t_h1, t_h2), us(t_h2, t_h3));
  }
  if (rc) { e->in_cycle = false; return rc; }
  float tot = 0; cudaEventElapsedTime(&tot, e->ev[8], e->ev[9]);
  e->last.ms_fit = ms_fit; e->last.ms_score = ms_score; e->last.ms_admit = ms_admit; e->last.ms_commit = ms_commit; e->last.ms_total = tot;
  if (stats) *stats = e->last;
  return GROVE_OK;
}

int32_t grove_get_placements(grove_engine_t* e, grove_placement_t* out, uint32_t cap, uint32_t* n_out) {
  if (!e || !n_out) return GROVE_ERR_INVALID_ARG;
  if (!e->have_results) return fail(e, GROVE_ERR_STATE, "no completed cycle");
  *n_out = e->n_out;
  if (!out) return GROVE_OK;
  if (cap < e->n_out) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  std::memcpy(out, e->h_out.p, sizeof(grove_placement_t) * e->n_out);
  return GROVE_OK;
}

int32_t grove_get_gang_status(grove_engine_t* e, grove_gang_status_t* out, uint32_t cap) {
  if (!e || !out) return GROVE_ERR_INVALID_ARG;
  if (!e->have_results) return fail(e, GROVE_ERR_STATE, "no completed cycle");
  if (cap < e->G) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  std::memcpy(out, e->h_status.p, sizeof(grove_gang_status_t) * e->G);
  return GROVE_OK;
}

// ---- stepping API (multi-GPU hosts reduce the returned buffer between the two calls of a round) ----
int32_t grove_round_eval(grove_engine_t* e, void** d_words, uint32_t* n_words, uint32_t* go) {
  if (!e || !d_words || !n_words || !go) return GROVE_ERR_INVALID_ARG;
  if (!e->in_cycle) return fail(e, GROVE_ERR_STATE, "grove_cycle_begin first");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  *go = 0; *d_words = e->d_xbuf.p; *n_words = uint32_t(xbuf_words(e));
  if (e->cfg.max_rounds && e->round_no >= e->cfg.max_rounds) return GROVE_OK;
  int32_t rc = round_eval(e, false);
  if (rc) return rc;
  rc = after_prepare(e, go);
  if (rc) return rc;
  if (!e->stream_ordered) CU_TRY(e, cudaStreamSynchronize(e->stream));
  return GROVE_OK;
}

int32_t grove_round_resolve(grove_engine_t* e, uint32_t* remaining) {
  if (!e) return GROVE_ERR_INVALID_ARG;
  if (!e->in_cycle) return fail(e, GROVE_ERR_STATE, "grove_cycle_begin first");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  int32_t rc = round_resolve(e, false);
  if (rc) return rc;
  if (!e->stream_ordered) CU_TRY(e, cudaStreamSynchronize(e->stream));
  if (remaining) *remaining = e->h_counters.p[2];  // unresolved before this round's commits (upper bound)
  return GROVE_OK;
}

int32_t grove_cycle_end(grove_engine_t* e, grove_cycle_stats_t* stats) {
  if (!e) return GROVE_ERR_INVALID_ARG;
  if (!e->in_cycle) return fail(e, GROVE_ERR_STATE, "grove_cycle_begin first");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  return finish_cycle(e, stats);
}

int32_t grove_engine_stream(grove_engine_t* e, void** stream) {
  if (!e || !stream) return GROVE_ERR_INVALID_ARG;
  *stream = e->stream;
  return GROVE_OK;
}

int32_t grove_set_stream_ordered(grove_engine_t* e, int32_t on) {
  if (!e) return GROVE_ERR_INVALID_ARG;
  e->stream_ordered = on != 0;
  return GROVE_OK;
}

// ---- introspection for parity tests ----
int32_t grove_debug_get_perm(grove_engine_t* e, uint32_t* sorted_to_caller, uint32_t cap) {
  if (!e || !sorted_to_caller) return GROVE_ERR_INVALID_ARG;
  if (!e->nodes_loaded) return fail(e, GROVE_ERR_STATE, "no node snapshot loaded");
  if (cap < e->N) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  std::memcpy(sorted_to_caller, e->perm.data(), sizeof(uint32_t) * e->N);
  return GROVE_OK;
}

int32_t grove_debug_get_fit_row(grove_engine_t* e, uint32_t clique, uint32_t* words, uint32_t cap_words) {
  if (!e || !words) return GROVE_ERR_INVALID_ARG;
  if (!e->have_results || clique >= e->Q) return fail(e, GROVE_ERR_STATE, "no completed cycle / bad clique");
  const uint32_t w = (e->N + 31) / 32;
  if (cap_words < w) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, cudaMemcpy(words, e->d_F.p + size_t(e->cinfo[clique].sig) * e->words, sizeof(uint32_t) * w, cudaMemcpyDeviceToHost));
  return GROVE_OK;
}

int32_t grove_debug_get_score_row(grove_engine_t* e, uint32_t clique, uint8_t* bytes, uint32_t cap_bytes) {
  if (!e || !bytes) return GROVE_ERR_INVALID_ARG;
  if (!e->have_results || clique >= e->Q) return fail(e, GROVE_ERR_STATE, "no completed cycle / bad clique");
  if (cap_bytes < e->N) return fail(e, GROVE_ERR_LIMIT, "output buffer too small");
  CU_TRY(e, cudaSetDevice(e->cfg.device));
  CU_TRY(e, cudaMemcpy(bytes, e->d_T.p + size_t(clique) * e->Npad, e->N, cudaMemcpyDeviceToHost));
  return GROVE_OK;
}

}  // extern "C"
