// tables.cuh -- table packing (caller order <-> topology order), anchor ancestors, round bookkeeping.
#pragma once
#include "common.cuh"

namespace grove {
// ------------------------------------------------------------------------------------------------
// table packing
// ------------------------------------------------------------------------------------------------
__global__ void k_gather(const grove_node_t* __restrict__ in, const uint32_t* __restrict__ perm,
                         const uint8_t* __restrict__ vdepth, uint4* __restrict__ nres, uint32_t n, uint32_t npad) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npad) return;
  uint4 r = make_uint4(0, 0, 0, 0);
  if (i < n) {
    const uint4* p = reinterpret_cast<const uint4*>(in + perm[i]);  // first 16 B of the 32 B record
    uint4 a = __ldg(p);
    r.x = a.x; r.y = a.y; r.z = a.z;                                // gpu | pods << 16 is already packed
    r.w = (a.w & 0xFFFFu) | (uint32_t(vdepth[i]) << 16);
  }
  nres[i] = r;
}

__global__ void k_scatter(grove_node_t* __restrict__ out, const grove_node_t* __restrict__ orig,
                          const uint32_t* __restrict__ perm, const uint4* __restrict__ nres, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = perm[i];
  grove_node_t nd = orig[c];
  uint4 r = nres[i];
  nd.free_cpu_milli = r.x; nd.free_mem_mib = r.y; nd.free_gpu = uint16_t(r.z & 0xFFFFu); nd.free_pods = uint16_t(r.z >> 16);
  out[c] = nd;
}

// churn deltas: new records for a few nodes (labels unchanged); also kept in the caller-order mirror that
// grove_get_nodes scatters back
__global__ void k_update(const uint32_t* __restrict__ idx_sorted, const grove_node_t* __restrict__ recs,
                         const uint8_t* __restrict__ vdepth, const uint32_t* __restrict__ perm, uint4* __restrict__ nres,
                         grove_node_t* __restrict__ nodes_in, uint32_t m) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  uint32_t s = idx_sorted[i];
  grove_node_t nd = recs[i];
  nres[s] = make_uint4(nd.free_cpu_milli, nd.free_mem_mib, uint32_t(nd.free_gpu) | (uint32_t(nd.free_pods) << 16),
                       (nd.flags & 0xFFFFu) | (uint32_t(vdepth[s]) << 16));
  nodes_in[perm[s]] = nd;
}

// anchor ancestors: node range of the anchor's domain at every level ([a,a) where its label is absent)
__global__ void k_anchor(Topo tp, GangInfo* ginfo, uint32_t G) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const uint32_t a = ginfo[g].anchor;
  const uint4 dm = tp.ndom[a];
  const uint32_t d[4] = {dm.x, dm.y, dm.z, dm.w};
#pragma unroll
  for (int l = 0; l < GROVE_MAX_LEVELS; ++l) {
    uint32_t lo = a, hi = a;
    if (l < (int)tp.L && d[l] != GROVE_DOM_ABSENT) { lo = tp.dom_lo[l][d[l]]; hi = tp.dom_hi[l][d[l]]; }
    ginfo[g].anc_lo[l] = lo; ginfo[g].anc_hi[l] = hi;
  }
}

// ------------------------------------------------------------------------------------------------
// round bookkeeping: which gangs are evaluated this round, and their clique rows
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, uint32_t lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(kFull, v, d);
    if (lane >= (uint32_t)d) v += t;
  }
  return v;
}

// grid of 1024-thread CTAs over the gangs; counters must be zeroed before the launch.
// Order inside active[] / rows[] depends on CTA arrival order; no result depends on it.
__global__ void __launch_bounds__(1024) k_prepare(Tables tb, RoundBufs rb, uint32_t round_no, uint32_t rank, uint32_t world) {
  __shared__ uint32_t s_warp_a[32], s_warp_r[32];
  __shared__ uint32_t s_base_a, s_base_r;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t r8 = uint8_t(round_no > 255 ? 255 : round_no);
  const uint32_t g = blockIdx.x * 1024 + tid;
  uint32_t act = 0, ncl = 0, unres = 0, coff = 0, prop = 0;
  bool ready = false;
  if (g < tb.G && rb.state[g] == GROVE_GANG_PENDING) {
    const grove_gang_t gg = tb.gangs[g];
    // walk the base chain: a scaled gang is rejected with any rejected / skipped ancestor (transitively),
    // and is ready once its direct base gang is admitted (pod/syncflow.go:319-358)
    bool dead = false;
    uint32_t b = gg.base_gang;
    for (int hop = 0; hop < 64 && b != GROVE_NONE_U32; ++hop) {
      const uint8_t bs = rb.state[b];
      if (bs == GROVE_GANG_REJECTED || bs == GROVE_GANG_BASE_REJECTED || bs == GROVE_GANG_GATED_SKIP) { dead = true; break; }
      if (bs == GROVE_GANG_ADMITTED) break;
      b = tb.gangs[b].base_gang;
    }
    if (dead) {
      rb.state[g] = GROVE_GANG_BASE_REJECTED; rb.round[g] = r8; prop = 1;
    } else {
      unres = 1;
      ready = gg.base_gang == GROVE_NONE_U32 || rb.state[gg.base_gang] == GROVE_GANG_ADMITTED;
      if (ready) rb.active_all[atomicAdd(rb.counters + 5, 1u)] = g;
      const bool mine = world <= 1 || (g % world) == rank;
      if (ready && mine) { act = 1; ncl = gg.n_cliques; coff = gg.clique_off; }
    }
  }
  const uint32_t ia = warp_incl_scan(act, lane), ir = warp_incl_scan(ncl, lane);
  const uint32_t un = __popc(__ballot_sync(kFull, unres));
  const uint32_t pr = __ballot_sync(kFull, prop);
  if (lane == 31) { s_warp_a[warp] = ia; s_warp_r[warp] = ir; }
  if (lane == 0) {
    if (un) atomicAdd(rb.counters + 2, un);
    if (pr) atomicOr(rb.counters + 3, 1u);
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t va = s_warp_a[lane], vr = s_warp_r[lane];
    const uint32_t sa = warp_incl_scan(va, lane), sr = warp_incl_scan(vr, lane);
    s_warp_a[lane] = sa - va; s_warp_r[lane] = sr - vr;  // exclusive per-warp offsets
    if (lane == 31) {
      s_base_a = sa ? atomicAdd(rb.counters + 0, sa) : 0u;
      s_base_r = sr ? atomicAdd(rb.counters + 1, sr) : 0u;
    }
  }
  __syncthreads();
  if (act) {
    const uint32_t oa = s_base_a + s_warp_a[warp] + ia - act;
    const uint32_t orr = s_base_r + s_warp_r[warp] + ir - ncl;
    rb.active[oa] = g;
    for (uint32_t i = 0; i < ncl; ++i) {
      rb.rows[orr + i] = coff + i;
      const uint32_t sg = tb.cinfo[coff + i].sig;
      // thousands of cliques share a few signatures: look before exchanging, so that only the first few arrivals per
      // signature pay a same-address atomic
      if (__ldcg(rb.sig_stamp + sg) != round_no && atomicExch(rb.sig_stamp + sg, round_no) != round_no)
        rb.sig_list[atomicAdd(rb.counters + 4, 1u)] = sg;
    }
  }
}

// dependency cycle or unreachable base: nothing can become active any more
__global__ void k_reject_rest(Tables tb, RoundBufs rb, uint32_t round_no) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < tb.G && rb.state[g] == GROVE_GANG_PENDING) {
    rb.state[g] = GROVE_GANG_BASE_REJECTED;
    rb.round[g] = uint8_t(round_no > 255 ? 255 : round_no);
  }
}

}  // namespace grove
