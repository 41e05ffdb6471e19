// fit.cuh -- K1: node x clique resource-fit bitmap, and the per-signature capacity tables built from it.
#pragma once
#include "common.cuh"

namespace grove {
// ------------------------------------------------------------------------------------------------
// K1: node x clique resource-fit bitmap.
// CTA = 1024 nodes (lane = node, record in registers) x a tile of kFitTile clique rows whose
// requirements are staged in shared memory.  One __ballot_sync per (warp, clique) yields the 32-bit
// fit word; words are staged in shared memory and written back as full 128-byte lines per row.
// ------------------------------------------------------------------------------------------------
constexpr int kFitTile = 128;

__global__ void __launch_bounds__(1024) k_fit(Topo tp, Tables tb, uint32_t* __restrict__ F) {
  __shared__ uint4 s_prm[kFitTile];
  __shared__ uint32_t s_row[kFitTile];
  __shared__ uint32_t s_out[kFitTile][32];
  const uint32_t n_rows = tb.S;                            // every signature of the submission
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t node = blockIdx.x * 1024 + tid;           // npad is a multiple of 1024
  const uint4 r = __ldg(tp.nres + node);
  const uint32_t gpu = r.z & 0xFFFFu, pods = r.z >> 16;
  const uint32_t depth = (r.w >> 16) & 0xFu;
  const uint32_t onehot = ((r.w & GROVE_NODE_SCHEDULABLE) && pods >= 1) ? (1u << ((r.w >> GROVE_NODE_CLASS_SHIFT) & 0xFu)) : 0u;
  for (uint32_t r0 = blockIdx.y * kFitTile; r0 < n_rows; r0 += gridDim.y * kFitTile) {
    __syncthreads();
    if (tid < kFitTile) {
      uint4 p = make_uint4(kFull, kFull, kFull, 0);  // never fits
      uint32_t sg = 0;
      if (r0 + tid < n_rows) { sg = r0 + tid; p = tb.sigs[sg]; }
      s_prm[tid] = p; s_row[tid] = sg;
    }
    __syncthreads();
    const int cnt = int(min(uint32_t(kFitTile), n_rows - r0));
#pragma unroll 4
    for (int c = 0; c < cnt; ++c) {
      const uint4 p = s_prm[c];
      bool ok = (r.x >= p.x) & (r.y >= p.y) & (gpu >= p.z) & ((p.w & onehot) != 0) & (depth >= (p.w >> 16));
      uint32_t b = __ballot_sync(kFull, ok);
      if (lane == 0) s_out[c][warp] = b;
    }
    __syncthreads();
    for (int c = warp; c < cnt; c += 32) F[size_t(s_row[c]) * tp.words + blockIdx.x * 32 + lane] = s_out[c][lane];
  }
}

// ------------------------------------------------------------------------------------------------
// Capacity tables for K3 (candidate pre-filter and packing):
// cap8[sig][n] = whole pods of the signature that fit on node n (0 if unfit, saturating at 255) and
// its per-domain sum / max.  "sum over the fill domain >= MinReplicas" is a necessary condition for a
// clique to be packable there, so domains failing it can be skipped without changing any result.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_cap8(Topo tp, Tables tb, const uint32_t* __restrict__ F, uint8_t* cap8) {
  const uint32_t sg = blockIdx.y;
  const uint32_t n = blockIdx.x * 256 + threadIdx.x;
  if (n >= tp.npad) return;
  uint32_t c = 0;
  if ((__ldg(F + size_t(sg) * tp.words + (n >> 5)) >> (n & 31)) & 1u) {
    const uint4 r = __ldg(tp.nres + n);
    const uint4 q = tb.sigs[sg];
    c = r.z >> 16;
    if (q.x) c = min(c, r.x / q.x);
    if (q.y) c = min(c, r.y / q.y);
    if (q.z) c = min(c, (r.z & 0xFFFFu) / q.z);
    c = min(c, 255u);
  }
  cap8[size_t(sg) * tp.npad + n] = uint8_t(c);
}

// one warp per (active signature, non-unit domain)
__global__ void __launch_bounds__(256) k_capsum(Topo tp, const uint8_t* __restrict__ cap8,
                                                uint32_t* capsum, uint32_t* capmax) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t j = (blockIdx.x * 256 + threadIdx.x) >> 5;   // column in the table row
  if (j >= tp.cap_stride) return;
  const uint32_t sg = blockIdx.y;
  uint32_t l = 0;
#pragma unroll
  for (uint32_t k = 0; k < GROVE_MAX_LEVELS; ++k)
    if (k < tp.L && !tp.unit[k] && j >= tp.cap_off[k]) l = k;
  const uint32_t d = j - tp.cap_off[l];
  const uint32_t lo = __ldg(tp.dom_lo[l] + d), hi = __ldg(tp.dom_hi[l] + d);
  const uint8_t* row = cap8 + size_t(sg) * tp.npad;
  uint32_t sum = 0, mx = 0;
  for (uint32_t n = lo + lane; n < hi; n += 32) { const uint32_t c = row[n]; sum += c; mx = max(mx, c); }
#pragma unroll
  for (int o = 16; o; o >>= 1) { sum += __shfl_xor_sync(kFull, sum, o); mx = max(mx, __shfl_xor_sync(kFull, mx, o)); }
  if (lane == 0) { capsum[size_t(sg) * tp.cap_stride + j] = sum; capmax[size_t(sg) * tp.cap_stride + j] = mx; }
}

}  // namespace grove
