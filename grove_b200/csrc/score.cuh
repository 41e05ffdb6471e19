// score.cuh -- K2: topology-distance score matrix.
#pragma once
#include "common.cuh"

namespace grove {
// ------------------------------------------------------------------------------------------------
// K2: topology-distance score matrix.  T[q][n] = fit ? 1 + #levels at which n shares the anchor's
// domain : 0.  Nodes are stored in topology order, so the anchor's domains are index ranges and the
// score is piecewise constant: a thread owns 16 consecutive nodes of one row, expands its 16 fit bits
// to bytes and writes one 128-bit store; only chunks straddling an ancestor boundary go per byte.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t spread4(uint32_t nib) {  // 4 bits -> 4 bytes of 0/1
  return (nib * 0x00204081u) & 0x01010101u;
}

__global__ void __launch_bounds__(256) k_score(Topo tp, Tables tb, const uint32_t* __restrict__ F, uint8_t* __restrict__ T, uint32_t n_rows) {
  const uint32_t cpr = tp.npad >> 4;  // 16-node chunks per row
  for (uint32_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
    const uint32_t q = r;   // one row per clique of the submission
    const CliqueInfo ci = tb.cinfo[q];
    const GangInfo* gi = tb.ginfo + ci.gang;
    const uint4 alo = __ldg(reinterpret_cast<const uint4*>(gi->anc_lo));
    const uint4 ahi = __ldg(reinterpret_cast<const uint4*>(gi->anc_hi));
    const uint32_t lo[4] = {alo.x, alo.y, alo.z, alo.w}, hi[4] = {ahi.x, ahi.y, ahi.z, ahi.w};
    const uint32_t* Frow = F + size_t(ci.sig) * tp.words;
    uint8_t* Trow = T + size_t(q) * tp.npad;
    for (uint32_t ch = blockIdx.x * blockDim.x + threadIdx.x; ch < cpr; ch += gridDim.x * blockDim.x) {
      const uint32_t n0 = ch << 4;
      const uint32_t bits = (__ldg(Frow + (n0 >> 5)) >> (n0 & 16)) & 0xFFFFu;
      uint4 out = make_uint4(0, 0, 0, 0);
      if (bits) {
        uint32_t inside = 0; bool uniform = true;
#pragma unroll
        for (int l = 0; l < GROVE_MAX_LEVELS; ++l) {
          if (l < (int)tp.L) {
            bool in = n0 >= lo[l] && n0 + 16 <= hi[l];
            bool outl = n0 + 16 <= lo[l] || n0 >= hi[l];
            inside += in; uniform &= (in | outl);
          }
        }
        if (uniform) {
          const uint32_t v = inside + 1;
          out.x = spread4(bits & 0xF) * v; out.y = spread4((bits >> 4) & 0xF) * v;
          out.z = spread4((bits >> 8) & 0xF) * v; out.w = spread4(bits >> 12) * v;
        } else {
          uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if ((bits >> j) & 1u) {
              uint32_t n = n0 + j, c = 1;
#pragma unroll
              for (int l = 0; l < GROVE_MAX_LEVELS; ++l) c += (l < (int)tp.L && n >= lo[l] && n < hi[l]);
              w[j >> 2] |= c << ((j & 3) * 8);
            }
          }
          out = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      *reinterpret_cast<uint4*>(Trow + n0) = out;
    }
  }
}

}  // namespace grove
