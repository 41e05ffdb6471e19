// score.cuh -- K2: topology-distance score matrix, and the join kernel that scores the alternatives.
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

__global__ void __launch_bounds__(256) k_score(Topo tp, Tables tb, RoundBufs rb, uint32_t n_rows) {
  const uint32_t cpr = tp.npad >> 4;  // 16-node chunks per row
  for (uint32_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
    const uint32_t q = __ldg(rb.rows + r);
    const CliqueInfo ci = tb.cinfo[q];
    const GangInfo* gi = tb.ginfo + ci.gang;
    const uint4 alo = __ldg(reinterpret_cast<const uint4*>(gi->anc_lo));
    const uint4 ahi = __ldg(reinterpret_cast<const uint4*>(gi->anc_hi));
    const uint32_t lo[4] = {alo.x, alo.y, alo.z, alo.w}, hi[4] = {ahi.x, ahi.y, ahi.z, ahi.w};
    const uint32_t* Frow = rb.F + size_t(ci.sig) * tp.words;
    uint8_t* Trow = rb.T + size_t(q) * tp.npad;
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

// ------------------------------------------------------------------------------------------------
// Scores of the alternatives.  The score matrix (K2) and the admission (K3) only share the fit data, so
// they run concurrently on two streams (K2 is HBM-write-bound, K3 is latency-bound: they overlap almost
// perfectly); this kernel joins them: one warp per (active gang, alternative) looks up T[clique row][node]
// for every entry, stores it next to the entry and reduces the minimum over the MinReplicas entries --
// the PlacementScore numerator (podgang.go:187-189).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alt_scores(Topo tp, Tables tb, RoundBufs rb) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t K = rb.K, P = rb.P;
  const uint32_t ai = w / K, a = w - ai * K;
  if (ai >= rb.counters[0]) return;
  const uint32_t g = rb.active[ai];
  if (a >= rb.nalt[g]) return;
  const uint32_t po = tb.ginfo[g].pod_off, coff = tb.gangs[g].clique_off;
  const uint32_t cnt = rb.alt_n[size_t(g) * K + a], nmin = rb.alt_nmin[size_t(g) * K + a];
  uint32_t mn = tp.L + 1;
  for (uint32_t i = lane; i < cnt; i += 32) {
    const size_t o = size_t(a) * P + po + i;
    const uint32_t cr = rb.alt_meta[o] & 0xFFu;
    const uint32_t sc = rb.T[size_t(coff + cr) * tp.npad + rb.alt_node[o]];
    rb.alt_meta[o] = cr | (sc << 8);
    if (i < nmin) mn = min(mn, sc);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) mn = min(mn, __shfl_xor_sync(kFull, mn, d));
  if (lane == 0) rb.alt_score[size_t(g) * K + a] = mn;
}

}  // namespace grove
