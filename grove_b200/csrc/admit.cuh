// admit.cuh -- K3: gang admission (candidate pre-filter, greedy packing, K alternatives per gang).
#pragma once
#include "common.cuh"

namespace grove {
// ------------------------------------------------------------------------------------------------
// K3: gang admission.
//
// A gang with a Required level tries the domains of that level in score order; every domain in which
// every PodClique's MinReplicas can be packed (all-or-nothing) is feasible, and the first K feasible
// ones become the gang's alternatives.  Candidate domains are independent of each other, so they are
// evaluated in parallel: a cheap pre-filter on per-signature capacity tables discards hopeless domains,
// then one LANE per plausible candidate runs the greedy packing as scalar code (ScalarEv) and ballots
// rank the successes.  Three launch forms share that code: a warp per gang (k_admit_warp) while a
// round has many gangs, a 4-warp or an 8-warp CTA per gang (k_admit) when it has few.  A gang without
// a gang-level constraint has a single candidate (the whole cluster): warp 0 packs it cooperatively
// (CoopEv), lanes = nodes of a fit word, prefix sums over the per-node capacities.
// Both evaluators implement the same DESIGN.md semantics and are checked against the oracle.
// ------------------------------------------------------------------------------------------------
struct GangRegs {   // per-gang constants
  uint32_t a, L, n;
  uint32_t anc_lo[GROVE_MAX_LEVELS], anc_hi[GROVE_MAX_LEVELS];
  uint32_t clique_off;
};

struct GangShared {
  uint4 clq[GROVE_MAX_GANG_CLIQUES];       // req_cpu, req_mem, req_gpu, min | replicas << 8 | level << 16 | preferred << 24
  grove_scope_t scopes[GROVE_MAX_GANG_SCOPES];
  uint32_t sig[GROVE_MAX_GANG_CLIQUES];    // fit-bitmap row of each clique
  // cooperative evaluator state (warp 0)
  uint32_t ent_node[GROVE_MAX_GANG_PODS];
  uint16_t ent_meta[GROVE_MAX_GANG_PODS];  // clique_rel | score << 8
  uint32_t Hlo[GROVE_MAX_GANG_CLIQUES], Hhi[GROVE_MAX_GANG_CLIQUES];
};

// Ordered pieces of [lo,hi): descending score, ties by ascending rotated index (n - anchor) mod N.
// from == L: node granularity (ring l = anc_l \ anc_{l+1}, upper part then lower part);
// from <  L: level-`from` domain granularity (the anchor's own domain whole, then the rings).
__device__ __forceinline__ int make_pieces(const GangRegs& g, uint32_t lo, uint32_t hi, uint32_t from,
                                           uint32_t* plo, uint32_t* phi) {
  int k = 0;
  if (g.a < lo || g.a >= hi) { plo[0] = lo; phi[0] = hi; return hi > lo ? 1 : 0; }
  uint32_t prev_lo = g.a, prev_hi = g.a;
  if (from < g.L) {
    prev_lo = max(g.anc_lo[from], lo); prev_hi = min(g.anc_hi[from], hi);
    if (prev_hi > prev_lo) { plo[k] = prev_lo; phi[k] = prev_hi; ++k; }
  }
  for (int l = int(min(from, g.L)) - 1; l >= 0; --l) {
    uint32_t cl = max(g.anc_lo[l], lo), ch = min(g.anc_hi[l], hi);
    if (ch > prev_hi) { plo[k] = prev_hi; phi[k] = ch; ++k; }
    if (prev_lo > cl) { plo[k] = cl; phi[k] = prev_lo; ++k; }
    prev_lo = min(cl, prev_lo); prev_hi = max(ch, prev_hi);
  }
  if (hi > prev_hi) { plo[k] = prev_hi; phi[k] = hi; ++k; }
  if (prev_lo > lo) { plo[k] = lo; phi[k] = prev_lo; ++k; }
  return k;
}

// The same ordered pieces as make_pieces, produced one at a time from a few registers (no per-thread
// arrays: the scalar evaluator runs one attempt per LANE and every local-memory word costs a cache line
// per warp).
struct PieceIt {
  uint32_t lo, hi, prev_lo, prev_hi, first_lo, first_hi;
  int l, stage;  // stage: 0 first piece pending, 1 rings (upper), 2 rings (lower), 3 tail upper, 4 tail lower, 5 done
  bool outside;
  __device__ __forceinline__ void init(const GangRegs& g, uint32_t lo_, uint32_t hi_, uint32_t from) {
    lo = lo_; hi = hi_;
    outside = g.a < lo || g.a >= hi;
    stage = 0; prev_lo = g.a; prev_hi = g.a; first_lo = first_hi = 0;
    l = int(min(from, g.L)) - 1;
    if (!outside && from < g.L) {
      first_lo = max(g.anc_lo[from], lo); first_hi = min(g.anc_hi[from], hi);
      prev_lo = first_lo; prev_hi = first_hi;
    }
  }
  __device__ __forceinline__ bool next(const GangRegs& g, uint32_t& a, uint32_t& b) {
    if (outside) { if (stage == 0 && hi > lo) { stage = 5; a = lo; b = hi; return true; } return false; }
    if (stage == 0) { stage = 1; if (first_hi > first_lo) { a = first_lo; b = first_hi; return true; } }
    while (stage == 1 || stage == 2) {
      if (l < 0) { stage = 3; break; }
      const uint32_t cl = max(g.anc_lo[l], lo), ch = min(g.anc_hi[l], hi);
      if (stage == 1) { stage = 2; if (ch > prev_hi) { a = prev_hi; b = ch; return true; } }
      // stage 2: lower part of ring l, then move one level out
      const uint32_t pl = prev_lo;
      prev_lo = min(cl, prev_lo); prev_hi = max(ch, prev_hi);
      --l; stage = 1;
      if (pl > cl) { a = cl; b = pl; return true; }
    }
    if (stage == 3) { stage = 4; if (hi > prev_hi) { a = prev_hi; b = hi; return true; } }
    if (stage == 4) { stage = 5; if (prev_lo > lo) { a = lo; b = prev_lo; return true; } }
    return false;
  }
};

__device__ __forceinline__ uint32_t cap_from(uint32_t cpu, uint32_t mem, uint32_t gpu, uint32_t pods, const uint4& q) {
  uint32_t c = pods;
  if (q.x) c = min(c, cpu / q.x);
  if (q.y) c = min(c, mem / q.y);
  if (q.z) c = min(c, gpu / q.z);
  return c;
}

// ---- cooperative evaluator: the whole warp packs ONE candidate range -----------------------------
template <bool kPref_>
struct CoopEv {
  static constexpr bool kPref = kPref_;
  const Topo& tp; const RoundBufs& rb; GangShared& sh; const GangRegs& g; uint32_t lane;
  uint32_t np;
  uint32_t tmask = 0;
  __device__ CoopEv(const Topo& t, const RoundBufs& r, GangShared& s, const GangRegs& gr, uint32_t ln)
      : tp(t), rb(r), sh(s), g(gr), lane(ln), np(0) {}

  __device__ __forceinline__ uint32_t cap_now(uint32_t cr, uint32_t n) const {
    const uint4 r = __ldg(tp.nres + n);
    uint32_t cpu = r.x, mem = r.y, gpu = r.z & 0xFFFFu, pods = r.z >> 16;
    for (uint32_t i = 0; i < np; ++i) {
      if (sh.ent_node[i] == n) {
        const uint4 o = sh.clq[sh.ent_meta[i] & 0xFFu];
        cpu -= o.x; mem -= o.y; gpu -= o.z; pods -= 1;
      }
    }
    return cap_from(cpu, mem, gpu, pods, sh.clq[cr]);
  }

  // up to `want` pods of clique cr on fit nodes of [lo,hi) in score order; returns pods placed
  __device__ uint32_t take(uint32_t cr, uint32_t lo, uint32_t hi, uint32_t want) {
    if (want == 0 || hi <= lo) return 0;
    const uint32_t* Frow = rb.F + size_t(sh.sig[cr]) * tp.words;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    uint32_t placed = 0;
    for (uint32_t a, b; placed < want && pit.next(g, a, b);) {
      const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
      for (uint32_t wb = w0; wb <= w1 && placed < want; wb += 32) {
        uint32_t myw = 0;  // 32 fit words at a time, one per lane
        if (wb + lane <= w1) {
          myw = __ldg(Frow + wb + lane);
          if (wb + lane == w0) myw &= kFull << (a & 31);
          if (wb + lane == w1 && (b & 31)) myw &= (1u << (b & 31)) - 1u;
        }
        uint32_t nz = __ballot_sync(kFull, myw != 0);
        while (nz && placed < want) {
          const uint32_t src = __ffs(nz) - 1; nz &= nz - 1;
          const uint32_t bits = __shfl_sync(kFull, myw, src);
          const uint32_t n = ((wb + src) << 5) + lane;
          const bool mine = (bits >> lane) & 1u;
          const uint32_t c = mine ? cap_now(cr, n) : 0u;
          const uint32_t incl = warp_incl_scan(c, lane), excl = incl - c, rem = want - placed;
          const uint32_t t = excl >= rem ? 0u : min(c, rem - excl);
          const uint32_t tincl = warp_incl_scan(t, lane);
          __syncwarp();  // every lane's cap_now scan of the entry stack (which may read one slot ahead) is over before it grows
          if (t) {
            const uint16_t meta = uint16_t(cr);
            const uint32_t pos = np + tincl - t;
            for (uint32_t j = 0; j < t; ++j) { sh.ent_node[pos + j] = n; sh.ent_meta[pos + j] = meta; }
          }
          const uint32_t tot = __shfl_sync(kFull, tincl, 31);
          np += tot; placed += tot;
          __syncwarp();
        }
      }
    }
    return placed;
  }

  __device__ bool fill_min(uint32_t cr, uint32_t lo, uint32_t hi) {
    const uint32_t m = sh.clq[cr].w & 0xFFu;
    const uint32_t mark = np;
    if (take(cr, lo, hi, m) < m) { np = mark; return false; }
    if (lane == 0) { sh.Hlo[cr] = lo; sh.Hhi[cr] = hi; }
    __syncwarp();
    return true;
  }

  // clique whose own Required level is a unit level (one node per domain, e.g. hostname):
  // first node of [lo,hi) in score order that takes all m pods
  __device__ bool find_unit(uint32_t cr, uint32_t lo, uint32_t hi) {
    const uint32_t m = sh.clq[cr].w & 0xFFu;
    const uint32_t* Frow = rb.F + size_t(sh.sig[cr]) * tp.words;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    for (uint32_t a, b; pit.next(g, a, b);) {
      const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
      for (uint32_t wb = w0; wb <= w1; wb += 32) {
        uint32_t myw = 0;
        if (wb + lane <= w1) {
          myw = __ldg(Frow + wb + lane);
          if (wb + lane == w0) myw &= kFull << (a & 31);
          if (wb + lane == w1 && (b & 31)) myw &= (1u << (b & 31)) - 1u;
        }
        uint32_t nz = __ballot_sync(kFull, myw != 0);
        while (nz) {
          const uint32_t src = __ffs(nz) - 1; nz &= nz - 1;
          const uint32_t bits = __shfl_sync(kFull, myw, src);
          const uint32_t n = ((wb + src) << 5) + lane;
          const bool mine = (bits >> lane) & 1u;
          const uint32_t c = mine ? cap_now(cr, n) : 0u;
          const uint32_t okb = __ballot_sync(kFull, mine && c >= m);
          __syncwarp();
          if (okb) {
            const uint32_t nn = ((wb + src) << 5) + (__ffs(okb) - 1);
            const uint16_t meta = uint16_t(cr);
            for (uint32_t j = lane; j < m; j += 32) { sh.ent_node[np + j] = nn; sh.ent_meta[np + j] = meta; }
            if (lane == 0) { sh.Hlo[cr] = nn; sh.Hhi[cr] = nn + 1; }
            np += m;
            __syncwarp();
            return true;
          }
        }
      }
    }
    return false;
  }
};

// ---- scalar evaluator: ONE lane packs one candidate range (lanes of a warp hold different candidates)
// kEnt > 0: the per-lane entry stack (pods placed so far) lives in shared memory, kEnt entries per lane,
// laid out [entry][thread] -- per-thread local arrays are what made this kernel thrash L1 (every local
// word is a 128 B line per warp).  kEnt == 0: local arrays sized for the largest legal gang.
template <bool kCaps, int kEnt, bool kPref_>
struct ScalarEv {
  static constexpr bool kPref = kPref_;
  const Topo& tp; const RoundBufs& rb; const GangShared& sh; const GangRegs& g;
  uint32_t np;
  uint32_t tmask;  // bit (n & 31) set for every node this attempt has put a pod on: quick 'untouched' test
  uint32_t k;   // candidate index of this lane
  uint32_t* sen; uint16_t* sem; uint32_t stride;
  uint32_t ent_node_l[kEnt ? 1 : GROVE_MAX_GANG_PODS];
  uint16_t ent_meta_l[kEnt ? 1 : GROVE_MAX_GANG_PODS];
  uint32_t Hlo[GROVE_MAX_GANG_CLIQUES], Hhi[GROVE_MAX_GANG_CLIQUES];  // written only for cliques with surplus replicas
  __device__ ScalarEv(const Topo& t, const RoundBufs& r, const GangShared& s, const GangRegs& gr, uint32_t* sen_, uint16_t* sem_, uint32_t stride_)
      : tp(t), rb(r), sh(s), g(gr), np(0), tmask(0), k(0), sen(sen_), sem(sem_), stride(stride_) {}
  __device__ __forceinline__ uint32_t& en(uint32_t i) { if constexpr (kEnt > 0) return sen[i * stride]; else return ent_node_l[i]; }
  __device__ __forceinline__ uint16_t& em(uint32_t i) { if constexpr (kEnt > 0) return sem[i * stride]; else return ent_meta_l[i]; }
  __device__ __forceinline__ uint32_t en(uint32_t i) const { if constexpr (kEnt > 0) return sen[i * stride]; else return ent_node_l[i]; }
  __device__ __forceinline__ uint16_t em(uint32_t i) const { if constexpr (kEnt > 0) return sem[i * stride]; else return ent_meta_l[i]; }
  __device__ __forceinline__ void note_domain(uint32_t cr, uint32_t lo, uint32_t hi) {
    const uint32_t w = sh.clq[cr].w;
    if (((w >> 8) & 0xFFu) > (w & 0xFFu)) { Hlo[cr] = lo; Hhi[cr] = hi; }
  }

  __device__ __forceinline__ uint32_t cap_now(uint32_t cr, uint32_t n) const {
    const uint4 r = __ldg(tp.nres + n);
    uint32_t cpu = r.x, mem = r.y, gpu = r.z & 0xFFFFu, pods = r.z >> 16;
    for (uint32_t i = 0; i < np; ++i) {
      if (en(i) == n) {
        const uint4 o = sh.clq[em(i) & 0xFFu];
        cpu -= o.x; mem -= o.y; gpu -= o.z; pods -= 1;
      }
    }
    return cap_from(cpu, mem, gpu, pods, sh.clq[cr]);
  }

  // has this gang already put pods on node n?
  __device__ __forceinline__ bool touched(uint32_t n) const {
    if (!((tmask >> (n & 31)) & 1u)) return false;
    for (uint32_t i = 0; i < np; ++i) if (en(i) == n) return true;
    return false;
  }

  // 32 capacity bytes [base, base+32) of one signature row as 8 independent word loads; returns the
  // mask of nodes in [a,b) whose capacity byte is non-zero
  __device__ __forceinline__ uint32_t load_caps(const uint8_t* row, uint32_t base, uint32_t a, uint32_t b) const {
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t v = (base + 4u * i < b) ? __ldg(reinterpret_cast<const uint32_t*>(row + base) + i) : 0u;
      const uint32_t nz = ((((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u) >> 7;   // bit 0 of each byte = byte != 0
      mask |= (((nz * 0x01020408u) >> 24) & 0xFu) << (4 * i);                             // gather the 4 flags, byte 0 first
    }
    if (a > base) mask &= kFull << (a - base);
    if (b < base + 32u) mask &= (1u << (b - base)) - 1u;
    return mask;
  }

  // capacity-table path: per-node capacities come as bytes (computed once per round for the signature);
  // only nodes this gang already touched, or saturated bytes, are recomputed from the node record
  __device__ __forceinline__ uint32_t take_caps(uint32_t cr, uint32_t lo, uint32_t hi, uint32_t want) {
    const uint8_t* row = rb.cap8 + size_t(sh.sig[cr]) * tp.npad;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    uint32_t placed = 0;
    for (uint32_t a, b; placed < want && pit.next(g, a, b);) {
      for (uint32_t base = a & ~3u; base < b && placed < want; base += 32) {
        uint32_t mask = load_caps(row, base, a, b);
        while (mask && placed < want) {
          const uint32_t j = __ffs(mask) - 1; mask &= mask - 1;
          const uint32_t n = base + j;
          uint32_t c = __ldg(row + n);  // the line was just fetched by load_caps
          if (c == 255u || touched(n)) c = cap_now(cr, n);
          const uint32_t t = min(c, want - placed);
          if (t) {
            const uint16_t meta = uint16_t(cr);
            for (uint32_t x = 0; x < t; ++x) { en(np + x) = n; em(np + x) = meta; }
            tmask |= 1u << (n & 31);
            np += t; placed += t;
          }
        }
      }
    }
    return placed;
  }

  __device__ __forceinline__ bool find_unit_caps(uint32_t cr, uint32_t lo, uint32_t hi) {
    const uint32_t m = sh.clq[cr].w & 0xFFu;
    const uint8_t* row = rb.cap8 + size_t(sh.sig[cr]) * tp.npad;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    for (uint32_t a, b; pit.next(g, a, b);) {
      for (uint32_t base = a & ~3u; base < b; base += 32) {
        uint32_t mask = load_caps(row, base, a, b);
        while (mask) {
          const uint32_t j = __ffs(mask) - 1; mask &= mask - 1;
          const uint32_t n = base + j;
          uint32_t c = __ldg(row + n);  // the line was just fetched by load_caps
          if (c < m && c != 255u) continue;   // capacities only shrink inside an attempt
          if (c == 255u || touched(n)) c = cap_now(cr, n);
          if (c >= m) {
            const uint16_t meta = uint16_t(cr);
            for (uint32_t x = 0; x < m; ++x) { en(np + x) = n; em(np + x) = meta; }
            tmask |= 1u << (n & 31);
            np += m; note_domain(cr, n, n + 1);
            return true;
          }
        }
      }
    }
    return false;
  }

  __device__ __forceinline__ uint32_t take(uint32_t cr, uint32_t lo, uint32_t hi, uint32_t want) {
    if (want == 0 || hi <= lo) return 0;
    if constexpr (kCaps) return take_caps(cr, lo, hi, want);
    const uint32_t* Frow = rb.F + size_t(sh.sig[cr]) * tp.words;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    uint32_t placed = 0;
    for (uint32_t a, b; placed < want && pit.next(g, a, b);) {
      const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
      for (uint32_t w = w0; w <= w1 && placed < want; ++w) {
        uint32_t bits = __ldg(Frow + w);
        if (w == w0) bits &= kFull << (a & 31);
        if (w == w1 && (b & 31)) bits &= (1u << (b & 31)) - 1u;
        while (bits && placed < want) {
          const uint32_t n = (w << 5) + (__ffs(bits) - 1); bits &= bits - 1;
          const uint32_t c = cap_now(cr, n);
          const uint32_t t = min(c, want - placed);
          if (t) {
            const uint16_t meta = uint16_t(cr);
            for (uint32_t j = 0; j < t; ++j) { en(np + j) = n; em(np + j) = meta; }
            tmask |= 1u << (n & 31);
            np += t; placed += t;
          }
        }
      }
    }
    return placed;
  }

  __device__ __forceinline__ bool fill_min(uint32_t cr, uint32_t lo, uint32_t hi) {
    const uint32_t m = sh.clq[cr].w & 0xFFu;
    const uint32_t mark = np;
    if (take(cr, lo, hi, m) < m) { np = mark; return false; }
    note_domain(cr, lo, hi);
    return true;
  }

  __device__ __forceinline__ bool find_unit(uint32_t cr, uint32_t lo, uint32_t hi) {
    if constexpr (kCaps) return find_unit_caps(cr, lo, hi);
    const uint32_t m = sh.clq[cr].w & 0xFFu;
    const uint32_t* Frow = rb.F + size_t(sh.sig[cr]) * tp.words;
    PieceIt pit; pit.init(g, lo, hi, g.L);
    for (uint32_t a, b; pit.next(g, a, b);) {
      const uint32_t w0 = a >> 5, w1 = (b - 1) >> 5;
      for (uint32_t w = w0; w <= w1; ++w) {
        uint32_t bits = __ldg(Frow + w);
        if (w == w0) bits &= kFull << (a & 31);
        if (w == w1 && (b & 31)) bits &= (1u << (b & 31)) - 1u;
        while (bits) {
          const uint32_t n = (w << 5) + (__ffs(bits) - 1); bits &= bits - 1;
          if (cap_now(cr, n) >= m) {
            const uint16_t meta = uint16_t(cr);
            for (uint32_t j = 0; j < m; ++j) { en(np + j) = n; em(np + j) = meta; }
            tmask |= 1u << (n & 31);
            np += m; note_domain(cr, n, n + 1);
            return true;
          }
        }
      }
    }
    return false;
  }
};

// ---- candidate pre-filter: a NECESSARY condition for place_in(lo, hi) to succeed ------------------
// (each clique alone must find MinReplicas worth of capacity in a domain it could be packed into, and
// the cliques of a scope must find it inside one common scope domain).  Reads only the small
// per-signature capacity tables.
__device__ __forceinline__ bool clique_plausible(const Topo& tp, const RoundBufs& rb, const GangShared& sh, uint32_t cr,
                                 uint32_t lo, uint32_t hi, int lvl, uint32_t dE) {
  const uint32_t w = sh.clq[cr].w;
  const uint32_t m = w & 0xFFu, ql = (w >> 16) & 0xFFu;
  if (m == 0) return true;
  const size_t row = size_t(sh.sig[cr]);
  const bool tabled = lvl >= 0 && !tp.unit[lvl];
  if (ql != GROVE_LEVEL_NONE && int(ql) > lvl) {
    if (tp.unit[ql]) {  // all m pods on one node
      if (tabled) return __ldg(rb.capmax + row * tp.cap_stride + tp.cap_off[lvl] + dE) >= m;
      for (uint32_t n = lo; n < hi; ++n) if (__ldg(rb.cap8 + row * tp.npad + n) >= m) return true;
      return false;
    }
    const uint32_t d0 = __ldg(tp.next_dom[ql] + lo), d1 = __ldg(tp.next_dom[ql] + hi);
    uint32_t any = 0;  // no early exit: the look-ups are independent and overlap
    for (uint32_t d = d0; d < d1; ++d) any |= __ldg(rb.capsum + row * tp.cap_stride + tp.cap_off[ql] + d) >= m;
    return any != 0;
  }
  if (tabled) return __ldg(rb.capsum + row * tp.cap_stride + tp.cap_off[lvl] + dE) >= m;
  uint32_t sum = 0;
  for (uint32_t n = lo; n < hi && sum < m; ++n) sum += __ldg(rb.cap8 + row * tp.npad + n);
  return sum >= m;
}

__device__ __forceinline__ bool scope_plausible(const Topo& tp, const RoundBufs& rb, const GangShared& sh, const grove_scope_t& s,
                                uint32_t lo, uint32_t hi, int lvl, uint32_t dE) {
  uint32_t all = 1;
  for (uint32_t i = 0; i < s.n_cliques; ++i) all &= clique_plausible(tp, rb, sh, s.first_clique + i, lo, hi, lvl, dE);
  return all != 0;
}

__device__ bool gang_plausible(const Topo& tp, const RoundBufs& rb, const GangShared& sh, uint32_t n_scopes,
                               uint32_t lo, uint32_t hi, int lvl, uint32_t dD) {
  for (uint32_t si = 0; si < n_scopes; ++si) {
    const grove_scope_t s = sh.scopes[si];
    bool ok = false;
    if (s.level != GROVE_LEVEL_NONE && int(s.level) > lvl) {
      const uint32_t d0 = __ldg(tp.next_dom[s.level] + lo), d1 = __ldg(tp.next_dom[s.level] + hi);
      uint32_t any = 0;  // no early exit: children are independent table look-ups
      for (uint32_t d = d0; d < d1; ++d)
        any |= scope_plausible(tp, rb, sh, s, __ldg(tp.dom_lo[s.level] + d), __ldg(tp.dom_hi[s.level] + d), int(s.level), d);
      ok = any != 0;
    } else {
      ok = scope_plausible(tp, rb, sh, s, lo, hi, lvl, dD);
    }
    if (!ok) return false;
  }
  return true;
}

// Levels a unit (gang / scope / clique) with constraint (req, pref) is tried at inside a parent range of
// level lvl (-1 = whole cluster): `first` down to the returned base.  base = the hard level (Required when
// deeper than the parent's, else the parent range itself); a deeper Preferred level is tried first and
// widened level by level up to base (podgang.go:110-117).  Oracle: level_span().  kPref = false: the submission
// carries no Preferred level anywhere, the walk is the single step `base` (the kernels are instantiated both ways so
// that Required-only workloads run the loop-free code).
template <bool kPref>
__device__ __forceinline__ int level_span(uint32_t req, uint32_t pref, int lvl, int& first) {
  const int base = (req != GROVE_LEVEL_NONE && int(req) > lvl) ? int(req) : lvl;
  first = (kPref && pref != GROVE_LEVEL_NONE && int(pref) > base) ? int(pref) : base;
  return base;
}

template <class Ev>
__device__ __forceinline__ bool place_scope(Ev& ev, const grove_scope_t& s, uint32_t lo, uint32_t hi, int lvl) {
  const Topo& tp = ev.tp;
  const uint32_t mark = ev.np;
  for (uint32_t i = 0; i < s.n_cliques; ++i) {
    const uint32_t cr = s.first_clique + i;
    const uint32_t w = ev.sh.clq[cr].w;
    const uint32_t m = w & 0xFFu;
    bool ok = false;
    int first;
    const int base = level_span<Ev::kPref>((w >> 16) & 0xFFu, w >> 24, lvl, first);
    int ql = first;
    do {
      if (ql > lvl) {
        // find_unit scans fit NODES: every fit node is a domain of a Required unit level (the fit row demands the
        // labels down to it), but not of a Preferred one (ragged labels) -- those take the domain walk below
        if (tp.unit[ql] && m >= 1 && ql == base) {
          ok = ev.find_unit(cr, lo, hi);
        } else {
          PieceIt pit; pit.init(ev.g, lo, hi, uint32_t(ql));
          for (uint32_t pa, pb; !ok && pit.next(ev.g, pa, pb);) {
            const uint32_t d0 = __ldg(tp.next_dom[ql] + pa), d1 = __ldg(tp.next_dom[ql] + pb);
            for (uint32_t d = d0; d < d1 && !ok; ++d)
              ok = ev.fill_min(cr, __ldg(tp.dom_lo[ql] + d), __ldg(tp.dom_hi[ql] + d));
          }
        }
      } else {
        ok = ev.fill_min(cr, lo, hi);
      }
    } while (Ev::kPref && !ok && --ql >= base);
    if (!ok) { ev.np = mark; return false; }
  }
  return true;
}

template <class Ev>
__device__ __forceinline__ bool place_in(Ev& ev, uint32_t n_scopes, uint32_t lo, uint32_t hi, int lvl) {
  const Topo& tp = ev.tp;
  ev.np = 0; ev.tmask = 0;
  for (uint32_t si = 0; si < n_scopes; ++si) {
    const grove_scope_t s = ev.sh.scopes[si];
    bool ok = false;
    int first;
    const int base = level_span<Ev::kPref>(s.level, s.preferred1 ? uint32_t(s.preferred1) - 1u : uint32_t(GROVE_LEVEL_NONE), lvl, first);
    int sl = first;
    do {
      if (sl > lvl) {
        PieceIt pit; pit.init(ev.g, lo, hi, uint32_t(sl));
        for (uint32_t pa, pb; !ok && pit.next(ev.g, pa, pb);) {
          const uint32_t d0 = __ldg(tp.next_dom[sl] + pa), d1 = __ldg(tp.next_dom[sl] + pb);
          for (uint32_t d = d0; d < d1 && !ok; ++d) {
            const uint32_t el = __ldg(tp.dom_lo[sl] + d), eh = __ldg(tp.dom_hi[sl] + d);
            // round-start capacities are an upper bound: a scope domain that lacks them cannot be packed
            if (ev.rb.cap8 && !scope_plausible(tp, ev.rb, ev.sh, s, el, eh, sl, d)) continue;
            ok = place_scope(ev, s, el, eh, sl);
          }
        }
      } else {
        ok = place_scope(ev, s, lo, hi, lvl);
      }
    } while (Ev::kPref && !ok && --sl >= base);
    if (!ok) { ev.np = 0; return false; }
  }
  return true;
}

// surplus beyond MinReplicas (best effort) of a successful scalar attempt
template <class Ev>
__device__ void finish_gang(Ev& ev, uint32_t n_cliques, uint32_t& n_min) {
  n_min = ev.np;
  for (uint32_t cr = 0; cr < n_cliques; ++cr) {
    const uint32_t w = ev.sh.clq[cr].w;
    const uint32_t mn = w & 0xFFu, rp = (w >> 8) & 0xFFu;
    if (rp > mn) ev.take(cr, ev.Hlo[cr], ev.Hhi[cr], rp - mn);
  }
}

__global__ void k_dbg_init(uint32_t* dbg, uint32_t G) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) dbg[g * 8 + 3] = GROVE_NONE_U32;
}

#ifndef GROVE_ADMIT_MINBLOCKS
#define GROVE_ADMIT_MINBLOCKS 6
#endif
#ifndef GROVE_ADMIT_MINBLOCKS_WIDE
#define GROVE_ADMIT_MINBLOCKS_WIDE 1   // 8-warp CTAs (2 per SM measured no faster: these rounds wait on their slowest gang)
#endif
constexpr int kAdmitThreads = 128;      // throughput rounds (many gangs): 4 warps per gang
constexpr int kAdmitThreadsWide = 256;  // latency rounds (few gangs): 8 warps per gang

// kMode 0: gangs with a gang-level constraint, packing from capacity bytes; 1: same, packing from fit
// words + node records (no capacity tables this cycle); 2: gangs without a gang-level constraint
// (cooperative).  Each instantiation skips the gangs of the other kind.
template <int kThreads, int kMode, int kEnt, bool kPref>
__global__ void __launch_bounds__(kThreads, kThreads == 128 ? GROVE_ADMIT_MINBLOCKS : GROVE_ADMIT_MINBLOCKS_WIDE) k_admit(Topo tp, Tables tb, RoundBufs rb) {
  __shared__ GangShared sh;
  __shared__ uint32_t s_en[(kEnt && kMode != 2 ? kEnt : 1) * kThreads];
  __shared__ uint16_t s_em[(kEnt && kMode != 2 ? kEnt : 1) * kThreads];
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ai = blockIdx.x;
  if (ai >= rb.counters[0]) return;
  const uint32_t gi = rb.active[ai];
  const grove_gang_t gg = tb.gangs[gi];
  if ((gg.level == GROVE_LEVEL_NONE && gg.preferred == GROVE_LEVEL_NONE) != (kMode == 2)) return;  // handled by the other instantiation
  const GangInfo info = tb.ginfo[gi];
  GangRegs g;
  g.a = info.anchor; g.L = tp.L; g.n = tp.n; g.clique_off = gg.clique_off;
#pragma unroll
  for (int l = 0; l < GROVE_MAX_LEVELS; ++l) { g.anc_lo[l] = info.anc_lo[l]; g.anc_hi[l] = info.anc_hi[l]; }
  for (uint32_t c = tid; c < gg.n_cliques; c += blockDim.x) {
    const grove_clique_t q = tb.cliques[gg.clique_off + c];
    sh.clq[c] = make_uint4(q.req_cpu_milli, q.req_mem_mib, q.req_gpu,
                           uint32_t(q.min_replicas) | (uint32_t(q.replicas) << 8) | (uint32_t(q.level) << 16) |
                               (GROVE_CLIQUE_PREFERRED(q.scope) << 24));
    sh.Hlo[c] = 0; sh.Hhi[c] = 0; sh.sig[c] = tb.cinfo[gg.clique_off + c].sig;
  }
  for (uint32_t si = tid; si < gg.n_scopes; si += blockDim.x) sh.scopes[si] = tb.scopes[gg.scope_off + si];
  __syncthreads();

  const uint32_t K = rb.K, P = rb.P;
  const long long dbg_t0 = rb.dbg ? clock64() : 0;
  if constexpr (kMode == 2) {
    // single candidate: the whole cluster, packed cooperatively by warp 0 (one alternative at most)
    if (warp != 0) return;
    CoopEv<kPref> ev(tp, rb, sh, g, lane);
    const bool ok = place_in(ev, gg.n_scopes, 0, tp.n, -1);
    uint32_t n_min = 0;
    if (ok) {
      n_min = ev.np;
      for (uint32_t cr = 0; cr < gg.n_cliques; ++cr) {
        const uint32_t w = sh.clq[cr].w;
        const uint32_t mn = w & 0xFFu, rp = (w >> 8) & 0xFFu;
        if (rp > mn) ev.take(cr, sh.Hlo[cr], sh.Hhi[cr], rp - mn);
      }
      for (uint32_t i = lane; i < ev.np; i += 32) {
        rb.alt_node[info.pod_off + i] = sh.ent_node[i];
        rb.alt_meta[info.pod_off + i] = sh.ent_meta[i];
      }
    }
    if (lane == 0) {
      rb.nalt[gi] = ok ? 1u : 0u;
      rb.alt_n[size_t(gi) * K] = ok ? ev.np : 0u;
      rb.alt_nmin[size_t(gi) * K] = n_min;
      rb.alt_top[size_t(gi) * K] = 0u;
    }
    return;
  } else {

  ScalarEv<kMode == 0, kEnt, kPref> ev(tp, rb, sh, g, s_en + tid, s_em + tid, kThreads);
  __shared__ uint32_t s_wcnt[kAdmitThreadsWide / 32];
  __shared__ uint32_t s_okmask[kAdmitThreadsWide / 32];
  __shared__ uint32_t s_ck[kAdmitThreadsWide], s_cl[kAdmitThreadsWide], s_ch[kAdmitThreadsWide];  // plausible candidates of the chunk, in order
  const uint32_t nwarp = blockDim.x >> 5;
  uint32_t nsucc = 0;  // feasible candidates found so far (block-uniform)
  // candidate levels: the Preferred level first (if any), widened level by level up to the Required one
  // (gl == -1: the whole cluster as a single candidate)
  int gfirst;
  const int gbase = level_span<kPref>(gg.level, gg.preferred, -1, gfirst);
  int gl = gfirst;
  do {
  // candidate domains of level gl in score order: up to kMaxPieces ranges of domain indices
  uint32_t r0[kMaxPieces], rcnt[kMaxPieces], D = 0;
  int npc = 1;
  if (gl >= 0) {
    uint32_t plo[kMaxPieces], phi[kMaxPieces];
    npc = make_pieces(g, 0, tp.n, uint32_t(gl), plo, phi);
    for (int p = 0; p < npc; ++p) {
      r0[p] = __ldg(tp.next_dom[gl] + plo[p]);
      rcnt[p] = __ldg(tp.next_dom[gl] + phi[p]) - r0[p];
      D += rcnt[p];
    }
  } else { r0[0] = 0; rcnt[0] = 1; D = 1; }
  // chunks of blockDim.x candidates in order: pre-filter all of them in parallel (cheap table look-ups),
  // compact the plausible ones, then run the packing on them one lane per candidate.  The first K
  // feasible candidates in order become the gang's alternatives.
  for (uint32_t base = 0; base < D && nsucc < K; base += blockDim.x) {
    {
      const uint32_t k = base + tid;
      uint32_t d = 0, dl = 0, dh = 0;
      bool plaus = false;
      if (k < D) {
        uint32_t rem = k;
        for (int p = 0; p < npc; ++p) { if (rem < rcnt[p]) { d = r0[p] + rem; break; } rem -= rcnt[p]; }
        if (gl >= 0) {
          dl = __ldg(tp.dom_lo[gl] + d); dh = __ldg(tp.dom_hi[gl] + d);
          plaus = rb.cap8 == nullptr || gang_plausible(tp, rb, sh, gg.n_scopes, dl, dh, gl, d);
        } else { dl = 0; dh = tp.n; plaus = true; }
      }
      const uint32_t pb = __ballot_sync(kFull, plaus);
      if (lane == 0) s_wcnt[warp] = __popc(pb);
      __syncthreads();
      uint32_t rank = __popc(pb & ((1u << lane) - 1u));
      for (uint32_t w = 0; w < warp; ++w) rank += s_wcnt[w];
      if (plaus) { s_ck[rank] = k; s_cl[rank] = dl; s_ch[rank] = dh; }
    }
    uint32_t total = 0;
    for (uint32_t w = 0; w < nwarp; ++w) total += s_wcnt[w];
    if (rb.dbg && tid == 0) { atomicAdd(rb.dbg + gi * 8 + 1, total); atomicAdd(rb.dbg + gi * 8 + 5, 1u); }
    __syncthreads();
    // attempt windows: `per_warp` candidates per warp, doubling (few attempts when the first candidates succeed,
    // a logarithmic number of windows when they do not)
    for (uint32_t abase = 0, per_warp = rb.width1; abase < total && nsucc < K; per_warp = min(32u, per_warp * 2u)) {
      const uint32_t width = per_warp * nwarp;
      // compacted candidate `slot` of the window goes to warp slot % nwarp: lanes of a warp run DIFFERENT packings
      // (divergent, serialised), so a window of w candidates costs ~w / nwarp attempts per warp, not min(w, 32)
      const uint32_t slot = lane * nwarp + warp;
      if (tid < (kAdmitThreadsWide / 32)) s_okmask[tid] = 0;
      __syncthreads();
      bool ok = false; uint32_t k = 0, dl = 0;
      if (slot < width && abase + slot < total) {
        k = s_ck[abase + slot]; dl = s_cl[abase + slot];
        ev.k = k;
        ok = place_in(ev, gg.n_scopes, dl, s_ch[abase + slot], gl);
        if (rb.dbg) { atomicAdd(rb.dbg + gi * 8 + 2, 1u); if (ok) atomicAdd(rb.dbg + gi * 8 + 0, 1u); }
        if (ok) atomicOr(&s_okmask[slot >> 5], 1u << (slot & 31));
      }
      __syncthreads();
      uint32_t stot = 0, srank = nsucc;
      for (uint32_t w = 0; w < (kAdmitThreadsWide / 32); ++w) {
        const uint32_t m = s_okmask[w];
        stot += __popc(m);
        if (ok) { if (w < (slot >> 5)) srank += __popc(m); else if (w == (slot >> 5)) srank += __popc(m & ((1u << (slot & 31)) - 1u)); }
      }
      if (ok && srank < K) {  // this lane holds the srank-th feasible candidate: publish it as an alternative
        uint32_t n_min;
        finish_gang(ev, gg.n_cliques, n_min);
        const size_t o = size_t(srank) * P + info.pod_off;
        for (uint32_t i = 0; i < ev.np; ++i) { rb.alt_node[o + i] = ev.en(i); rb.alt_meta[o + i] = ev.em(i); }
        rb.alt_n[size_t(gi) * K + srank] = ev.np;
        rb.alt_nmin[size_t(gi) * K + srank] = n_min;
        rb.alt_top[size_t(gi) * K + srank] = dl;
        if (rb.dbg && srank == 0) rb.dbg[gi * 8 + 3] = k;
      }
      nsucc += stot;
      abase += width;
      __syncthreads();  // s_okmask is rewritten by the next window
    }
    __syncthreads();  // the candidate list is rewritten by the next chunk
  }
  } while (kPref && nsucc < K && --gl >= gbase);  // candidate levels
  if (tid == 0) rb.nalt[gi] = min(nsucc, K);
  if (rb.dbg && tid == 0) rb.dbg[gi * 8 + 4] = uint32_t(clock64() - dbg_t0);
  }
}

// ------------------------------------------------------------------------------------------------
// K3, throughput form: ONE WARP per gang (4 gangs per CTA), used while a round has many gangs.  The
// packing of one gang is a chain of dependent L2 look-ups (latency-bound), so what matters is how many
// gangs are in flight per SM: a warp per gang keeps 24 of them resident instead of 6 with a CTA per
// gang, and every intra-gang barrier is a __syncwarp.  Same semantics as k_admit: candidates in chunks
// of 32 (one lane each), pre-filter, packing attempts on the plausible lanes in windows of `width0`,
// ballots rank the successes, the first K in order are published as alternatives.
// ------------------------------------------------------------------------------------------------
constexpr int kAdmitWarpGangs = 4;
constexpr int kEntSmem = 16;  // per-lane entry stack depth of the shared-memory form (gangs of <= 16 pods)

template <bool kCaps, int kEnt, bool kPref>
__global__ void __launch_bounds__(kAdmitWarpGangs * 32, GROVE_ADMIT_MINBLOCKS) k_admit_warp(Topo tp, Tables tb, RoundBufs rb) {
  __shared__ GangShared shs[kAdmitWarpGangs];
  __shared__ uint32_t s_en[(kEnt ? kEnt : 1) * kAdmitWarpGangs * 32];
  __shared__ uint16_t s_em[(kEnt ? kEnt : 1) * kAdmitWarpGangs * 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ai = blockIdx.x * kAdmitWarpGangs + warp;
  if (ai >= rb.counters[0]) return;
  const uint32_t gi = rb.active[ai];
  const grove_gang_t gg = tb.gangs[gi];
  if (gg.level == GROVE_LEVEL_NONE && gg.preferred == GROVE_LEVEL_NONE) return;  // unconstrained gangs: k_admit<.,2>
  GangShared& sh = shs[warp];
  const GangInfo info = tb.ginfo[gi];
  GangRegs g;
  g.a = info.anchor; g.L = tp.L; g.n = tp.n; g.clique_off = gg.clique_off;
#pragma unroll
  for (int l = 0; l < GROVE_MAX_LEVELS; ++l) { g.anc_lo[l] = info.anc_lo[l]; g.anc_hi[l] = info.anc_hi[l]; }
  for (uint32_t c = lane; c < gg.n_cliques; c += 32) {
    const grove_clique_t q = tb.cliques[gg.clique_off + c];
    sh.clq[c] = make_uint4(q.req_cpu_milli, q.req_mem_mib, q.req_gpu,
                           uint32_t(q.min_replicas) | (uint32_t(q.replicas) << 8) | (uint32_t(q.level) << 16) |
                               (GROVE_CLIQUE_PREFERRED(q.scope) << 24));
    sh.sig[c] = tb.cinfo[gg.clique_off + c].sig;
  }
  for (uint32_t si = lane; si < gg.n_scopes; si += 32) sh.scopes[si] = tb.scopes[gg.scope_off + si];
  __syncwarp();
  const uint32_t K = rb.K, P = rb.P;
  const long long dbg_t0 = rb.dbg ? clock64() : 0;
  ScalarEv<kCaps, kEnt, kPref> ev(tp, rb, sh, g, s_en + threadIdx.x, s_em + threadIdx.x, kAdmitWarpGangs * 32);
  uint32_t nsucc = 0;
  // candidate levels: the Preferred level first (if any), widened level by level up to the Required one
  // (gl == -1: the whole cluster as a single candidate)
  int gfirst;
  const int gbase = level_span<kPref>(gg.level, gg.preferred, -1, gfirst);
  int gl = gfirst;
  do {
  uint32_t r0[kMaxPieces], rcnt[kMaxPieces], D = 0;
  int npc = 1;
  if (gl >= 0) {
    uint32_t plo[kMaxPieces], phi[kMaxPieces];
    npc = make_pieces(g, 0, tp.n, uint32_t(gl), plo, phi);
    for (int p = 0; p < npc; ++p) {
      r0[p] = __ldg(tp.next_dom[gl] + plo[p]);
      rcnt[p] = __ldg(tp.next_dom[gl] + phi[p]) - r0[p];
      D += rcnt[p];
    }
  } else { r0[0] = 0; rcnt[0] = 1; D = 1; }
  for (uint32_t base = 0; base < D && nsucc < K; base += 32) {
    const uint32_t k = base + lane;
    uint32_t d = 0, dl = 0, dh = 0;
    bool plaus = false;
    if (k < D) {
      uint32_t rem = k;
      for (int p = 0; p < npc; ++p) { if (rem < rcnt[p]) { d = r0[p] + rem; break; } rem -= rcnt[p]; }
      if (gl >= 0) {
        dl = __ldg(tp.dom_lo[gl] + d); dh = __ldg(tp.dom_hi[gl] + d);
        plaus = rb.cap8 == nullptr || gang_plausible(tp, rb, sh, gg.n_scopes, dl, dh, gl, d);
      } else { dl = 0; dh = tp.n; plaus = true; }
    }
    uint32_t todo = __ballot_sync(kFull, plaus);
    if (rb.dbg && lane == 0) { atomicAdd(rb.dbg + gi * 8 + 1, __popc(todo)); atomicAdd(rb.dbg + gi * 8 + 5, 1u); }
    bool first_window = base == 0;
    while (todo && nsucc < K) {
      // first window: a few more candidates than alternatives wanted (in an uncongested cluster nearly all
      // fit); if that was not enough the cluster is congested: take every plausible candidate of the chunk
      uint32_t sel = 0, t = todo;
      const uint32_t wmax = first_window ? rb.width0 : 32u;
      first_window = false;
      for (uint32_t i = 0; i < wmax && t; ++i) { const uint32_t b = t & (0u - t); sel |= b; t ^= b; }
      todo &= ~sel;
      bool ok = false;
      if ((sel >> lane) & 1u) {
        ev.k = k;
        ok = place_in(ev, gg.n_scopes, dl, dh, gl);
        if (rb.dbg) { atomicAdd(rb.dbg + gi * 8 + 2, 1u); if (ok) atomicAdd(rb.dbg + gi * 8 + 0, 1u); }
      }
      const uint32_t sb = __ballot_sync(kFull, ok);
      const uint32_t srank = nsucc + __popc(sb & ((1u << lane) - 1u));
      if (ok && srank < K) {  // this lane holds the srank-th feasible candidate: publish it as an alternative
        uint32_t n_min;
        finish_gang(ev, gg.n_cliques, n_min);
        const size_t o = size_t(srank) * P + info.pod_off;
        for (uint32_t i = 0; i < ev.np; ++i) { rb.alt_node[o + i] = ev.en(i); rb.alt_meta[o + i] = ev.em(i); }
        rb.alt_n[size_t(gi) * K + srank] = ev.np;
        rb.alt_nmin[size_t(gi) * K + srank] = n_min;
        rb.alt_top[size_t(gi) * K + srank] = dl;
        if (rb.dbg && srank == 0) rb.dbg[gi * 8 + 3] = k;
      }
      nsucc += __popc(sb);
      __syncwarp();
    }
  }
  } while (kPref && nsucc < K && --gl >= gbase);  // candidate levels
  if (lane == 0) rb.nalt[gi] = min(nsucc, K);
  if (rb.dbg && lane == 0) rb.dbg[gi * 8 + 4] = uint32_t(clock64() - dbg_t0);
}

}  // namespace grove
