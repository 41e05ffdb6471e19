// tables.cuh -- table packing (caller order <-> topology order), anchor ancestors.
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

}  // namespace grove
