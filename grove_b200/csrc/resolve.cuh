// resolve.cuh -- conflict resolution + commit of a round (cooperative), and the output kernels.
#pragma once
#include "common.cuh"

namespace grove {
// ------------------------------------------------------------------------------------------------
// Conflict resolution of one round (cooperative launch: grid-wide barriers between the phases).
// Up to GROVE_SUBROUNDS passes over the alternatives computed by k_admit: every undecided gang
// proposes its first alternative that touches no node committed earlier in this round; proposals
// claim their nodes with the gang's order rank (atomicMin); a gang that holds every node it claimed
// (warp ballot) commits: node table decremented, nodes marked taken, placement copied to the final
// arrays.  Gangs without any alternative are rejected.  One warp per gang, strided over the grid; a
// gang is always handled by the same warp, so its cur/prop bytes need no cross-CTA visibility; taken,
// claim and flags do and are read with ld.cg / volatile.
// ------------------------------------------------------------------------------------------------
constexpr int kResolveThreads = 1024;  // few, fat CTAs: the grid barrier is what this kernel waits on
constexpr uint32_t kClaimRounds = 0x7Eu / GROVE_SUBROUNDS;  // rounds whose (round, sub-round) tags fit below the "no claim" byte

// tag_hi: claim tag of this round's sub-round 0 (tags decrease: newer claims win atomicMin against stale ones);
// tk: this round's stamp in taken[]; flags[] are stamped with the round number.  See round_resolve().
__global__ void __launch_bounds__(kResolveThreads) k_resolve(Topo tp, Tables tb, RoundBufs rb, uint4* nres, uint32_t round_no,
                                                            uint32_t tag_hi, uint32_t tk) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const uint32_t na = rb.counters[5];
  const uint8_t r8 = uint8_t(round_no > 255 ? 255 : round_no);
  const uint32_t K = rb.K, P = rb.P;
  const volatile uint32_t* vflags = rb.flags;
  if (na <= nw) {
    // Fast path (the usual one: the grid is sized for it): a warp owns at most ONE gang for the whole round,
    // so its constants and the nodes of its current alternative stay in registers and every phase is one
    // level of look-ups (taken / claim) instead of a chain of six.
    const bool have = gw < na;
    uint32_t g = 0, nalt = 0, po = 0, order0 = 0, c = 0, cnt = 0;
    uint32_t nd[4] = {GROVE_NONE_U32, GROVE_NONE_U32, GROVE_NONE_U32, GROVE_NONE_U32};
    bool pending = false;
    auto load_alt = [&]() {
      cnt = rb.alt_n[size_t(g) * K + c];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const uint32_t i = lane + 32u * j; nd[j] = i < cnt ? rb.alt_node[size_t(c) * P + po + i] : GROVE_NONE_U32; }
    };
    if (have) {
      g = rb.active_all[gw];
      nalt = rb.nalt[g]; po = tb.ginfo[g].pod_off; order0 = tb.ginfo[g].order;
      pending = nalt > 0;
      if (!pending && lane == 0) { rb.state[g] = GROVE_GANG_REJECTED; rb.round[g] = r8; }
      if (pending) load_alt();
    }
    for (uint32_t sub = 0; sub < GROVE_SUBROUNDS; ++sub) {
      const uint32_t order = order0 | ((tag_hi - sub) << 24);
      bool proposed = false;
      if (pending) {
        while (c < nalt) {  // first alternative that touches no node committed earlier in this round
          bool hit = false;
#pragma unroll
          for (int j = 0; j < 4; ++j) if (nd[j] != GROVE_NONE_U32) hit |= __ldcg(rb.taken + nd[j]) == tk;
          if (!__any_sync(kFull, hit)) break;
          if (++c < nalt) load_alt();
        }
        if (c < nalt) {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (nd[j] != GROVE_NONE_U32) atomicMin(rb.claim + nd[j], order);
          if (lane == 0) rb.flags[sub] = round_no;
          proposed = true;
        } else {
          pending = false;  // nothing left to propose: re-evaluated next round
        }
      }
      grid.sync();
      if (vflags[sub] != round_no) break;  // no proposal anywhere: the round is settled
      if (proposed) {
        bool win = true;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (nd[j] != GROVE_NONE_U32) win &= __ldcg(rb.claim + nd[j]) == order;
        if (__all_sync(kFull, win)) {
          const uint32_t coff = tb.gangs[g].clique_off;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (nd[j] == GROVE_NONE_U32) continue;
            const uint32_t i = lane + 32u * j;
            const uint32_t meta = rb.alt_meta[size_t(c) * P + po + i];
            const grove_clique_t q = tb.cliques[coff + (meta & 0xFFu)];
            uint32_t* r = reinterpret_cast<uint32_t*>(nres + nd[j]);
            // winners own their nodes exclusively in a sub-round; atomics only order this gang's own pods
            if (q.req_cpu_milli) atomicSub(r + 0, q.req_cpu_milli);
            if (q.req_mem_mib) atomicSub(r + 1, q.req_mem_mib);
            atomicSub(r + 2, uint32_t(q.req_gpu) | (1u << 16));
            rb.taken[nd[j]] = uint8_t(tk);
            rb.ent_node[po + i] = nd[j]; rb.ent_meta[po + i] = uint16_t(meta);
          }
          if (lane == 0) {
            rb.spec_n[g] = uint16_t(cnt); rb.spec_score[g] = uint8_t(rb.alt_score[size_t(g) * K + c]);
            rb.spec_top[g] = rb.alt_top[size_t(g) * K + c];
            rb.state[g] = GROVE_GANG_ADMITTED; rb.round[g] = r8;
          }
          pending = false;
        }
      }
      grid.sync();
    }
    return;
  }
  // generic path: more gangs than warps
  for (uint32_t ai = gw; ai < na; ai += nw) {
    const uint32_t g = rb.active_all[ai];
    if (lane == 0) {
      rb.cur[g] = 0; rb.prop[g] = 0;
      if (rb.nalt[g] == 0) { rb.state[g] = GROVE_GANG_REJECTED; rb.round[g] = r8; }
    }
  }
  __syncwarp();
  for (uint32_t sub = 0; sub < GROVE_SUBROUNDS; ++sub) {
    // claims carry the (round, sub-round) in their top bits so that a later sub-round always beats stale claims of
    // an earlier one (atomicMin): nothing has to be withdrawn between sub-rounds or rounds
    const uint32_t tag = (tag_hi - sub) << 24;
    // ---- propose ----
    for (uint32_t ai = gw; ai < na; ai += nw) {
      const uint32_t g = rb.active_all[ai];
      if (rb.state[g] != GROVE_GANG_PENDING) continue;
      const uint32_t nalt = rb.nalt[g], po = tb.ginfo[g].pod_off, order = tb.ginfo[g].order | tag;
      uint32_t c = rb.cur[g], cnt = 0;
      while (c < nalt) {  // first alternative that touches no node committed earlier in this round
        cnt = rb.alt_n[size_t(g) * K + c];
        bool hit = false;
        for (uint32_t i = lane; i < cnt; i += 32) hit |= __ldcg(rb.taken + rb.alt_node[size_t(c) * P + po + i]) == tk;
        if (!__any_sync(kFull, hit)) break;
        ++c;
      }
      if (lane == 0) rb.cur[g] = uint8_t(c);
      if (c >= nalt) continue;  // nothing left to propose: re-evaluated next round
      for (uint32_t i = lane; i < cnt; i += 32) atomicMin(rb.claim + rb.alt_node[size_t(c) * P + po + i], order);
      if (lane == 0) { rb.prop[g] = uint8_t(sub + 1); rb.flags[sub] = round_no; }
    }
    grid.sync();
    if (vflags[sub] != round_no) break;  // no proposal anywhere: the round is settled
    // ---- decide ----
    for (uint32_t ai = gw; ai < na; ai += nw) {
      const uint32_t g = rb.active_all[ai];
      if (rb.prop[g] != sub + 1 || rb.state[g] != GROVE_GANG_PENDING) continue;
      const uint32_t po = tb.ginfo[g].pod_off, order = tb.ginfo[g].order | tag, c = rb.cur[g];
      const uint32_t cnt = rb.alt_n[size_t(g) * K + c];
      bool win = true;
      for (uint32_t i = lane; i < cnt; i += 32) win &= __ldcg(rb.claim + rb.alt_node[size_t(c) * P + po + i]) == order;
      if (!__all_sync(kFull, win)) continue;
      const uint32_t coff = tb.gangs[g].clique_off;
      for (uint32_t i = lane; i < cnt; i += 32) {
        const uint32_t nd = rb.alt_node[size_t(c) * P + po + i], meta = rb.alt_meta[size_t(c) * P + po + i];
        const grove_clique_t q = tb.cliques[coff + (meta & 0xFFu)];
        uint32_t* r = reinterpret_cast<uint32_t*>(nres + nd);
        // winners own their nodes exclusively in a sub-round; atomics only order this gang's own pods
        if (q.req_cpu_milli) atomicSub(r + 0, q.req_cpu_milli);
        if (q.req_mem_mib) atomicSub(r + 1, q.req_mem_mib);
        atomicSub(r + 2, uint32_t(q.req_gpu) | (1u << 16));
        rb.taken[nd] = uint8_t(tk);
        rb.ent_node[po + i] = nd; rb.ent_meta[po + i] = uint16_t(meta);
      }
      if (lane == 0) {
        rb.spec_n[g] = uint16_t(cnt); rb.spec_score[g] = uint8_t(rb.alt_score[size_t(g) * K + c]);
        rb.spec_top[g] = rb.alt_top[size_t(g) * K + c];
        rb.state[g] = GROVE_GANG_ADMITTED; rb.round[g] = r8;
      }
    }
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// outputs: compact the admitted gangs' entries into caller order / caller node indices
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_finalize(Topo tp, Tables tb, RoundBufs rb, grove_gang_status_t* status, uint32_t* totals) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_run, s_tot, s_adm, s_rej;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { s_run = 0; s_adm = 0; s_rej = 0; }
  __syncthreads();
  for (uint32_t base = 0; base < tb.G; base += 1024) {
    const uint32_t g = base + tid;
    uint32_t cnt = 0; uint8_t st = 0;
    if (g < tb.G) { st = rb.state[g]; if (st == GROVE_GANG_ADMITTED) cnt = rb.spec_n[g]; }
    const uint32_t incl = warp_incl_scan(cnt, lane);
    if (lane == 31) s_warp[warp] = incl;
    const uint32_t na = __popc(__ballot_sync(kFull, st == GROVE_GANG_ADMITTED));
    const uint32_t nr = __popc(__ballot_sync(kFull, st == GROVE_GANG_REJECTED || st == GROVE_GANG_BASE_REJECTED));
    if (lane == 0) { if (na) atomicAdd(&s_adm, na); if (nr) atomicAdd(&s_rej, nr); }
    __syncthreads();
    if (warp == 0) {
      const uint32_t v = s_warp[lane];
      const uint32_t s = warp_incl_scan(v, lane);
      s_warp[lane] = s - v;
      if (lane == 31) s_tot = s;
    }
    __syncthreads();
    if (g < tb.G) {
      grove_gang_status_t o;
      o.state = st; o.round = rb.round[g]; o.n_pods = cnt; o.placement_off = s_run + s_warp[warp] + incl - cnt;
      o.score_num = 0; o.score_den = 0; o.top_domain_lo = GROVE_NONE_U32;
      if (st == GROVE_GANG_ADMITTED) { o.score_num = rb.spec_score[g]; o.score_den = uint8_t(tp.L + 1); o.top_domain_lo = rb.spec_top[g]; }
      status[g] = o;
    }
    __syncthreads();
    if (tid == 0) s_run += s_tot;
    __syncthreads();
  }
  if (tid == 0) { totals[0] = s_run; totals[1] = s_adm; totals[2] = s_rej; }
}

// one warp per gang: its entries -> caller node indices, at the offset k_finalize assigned
__global__ void k_emit(Tables tb, RoundBufs rb, const uint32_t* __restrict__ perm, const grove_gang_status_t* __restrict__ status,
                       grove_placement_t* out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= tb.G) return;
  const grove_gang_status_t st = status[g];
  if (st.state != GROVE_GANG_ADMITTED) return;
  const uint32_t po = tb.ginfo[g].pod_off, coff = tb.gangs[g].clique_off;
  for (uint32_t i = lane; i < st.n_pods; i += 32) {
    grove_placement_t p;
    p.clique = coff + (rb.ent_meta[po + i] & 0xFFu);
    p.node = perm[rb.ent_node[po + i]];
    out[st.placement_off + i] = p;
  }
}

}  // namespace grove
