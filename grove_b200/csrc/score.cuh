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

// c0 / cpr / tstride: the 16-node chunks [c0, c0 + cpr) of every row are built, into rows of tstride bytes (the whole row:
// 0, npad / 16, npad; a node-range shard of a multi-GPU score pass: its chunks only -- engine.cu grove_run_score_pass)
__global__ void __launch_bounds__(256) k_score(Topo tp, Tables tb, const uint32_t* __restrict__ F, uint8_t* __restrict__ T, uint32_t n_rows,
                                               uint32_t c0, uint32_t cpr, size_t tstride) {
  for (uint32_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
    const uint32_t q = r;   // one row per clique of the submission
    const CliqueInfo ci = tb.cinfo[q];
    const GangInfo* gi = tb.ginfo + ci.gang;
    const uint4 alo = __ldg(reinterpret_cast<const uint4*>(gi->anc_lo));
    const uint4 ahi = __ldg(reinterpret_cast<const uint4*>(gi->anc_hi));
    const uint32_t lo[4] = {alo.x, alo.y, alo.z, alo.w}, hi[4] = {ahi.x, ahi.y, ahi.z, ahi.w};
    const uint32_t* Frow = F + size_t(ci.sig) * tp.words;
    uint8_t* Trow = T + size_t(q) * tstride - (size_t(c0) << 4);
    for (uint32_t ch = c0 + blockIdx.x * blockDim.x + threadIdx.x; ch < c0 + cpr; ch += gridDim.x * blockDim.x) {
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
// Node-range shard summary (multi-GPU score pass): what ONE all-reduce(SUM) over the ranks turns into cluster-wide feasibility.
//   out[g]      gangs with a Required level: domains of that level STARTING in [lo, hi) in which every clique of the gang finds
//               MinReplicas worth of capacity on its own (a necessary condition for the gang to fit there); 0 for the others
//   out[G + q]  pods of clique q that fit on the nodes of [lo, hi) (capacity bytes, saturating at 255 a node)
// One warp per gang, then one thread per clique.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_shard_summary(Topo tp, Tables tb, const uint8_t* __restrict__ cap8, const uint32_t* __restrict__ capsum,
                                                       const uint32_t* __restrict__ sig_sum, uint32_t lo, uint32_t hi, int32_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t g = w; g < tb.G; g += nw) {
    const grove_gang_t gg = tb.gangs[g];
    uint32_t cnt = 0;
    if (gg.level != GROVE_LEVEL_NONE && gg.level < tp.L) {
      const uint32_t l = gg.level;
      const uint32_t d0 = __ldg(tp.next_dom[l] + lo), d1 = __ldg(tp.next_dom[l] + hi);
      for (uint32_t d = d0 + lane; d < d1; d += 32) {
        bool ok = true;
        for (uint32_t c = 0; c < gg.n_cliques && ok; ++c) {
          const grove_clique_t q = tb.cliques[gg.clique_off + c];
          if (q.min_replicas == 0) continue;
          const uint32_t sg = tb.cinfo[gg.clique_off + c].sig;
          uint32_t sum;
          if (tp.unit[l]) sum = cap8[size_t(sg) * tp.npad + __ldg(tp.dom_lo[l] + d)];
          else sum = capsum[size_t(sg) * tp.cap_stride + tp.cap_off[l] + d];
          ok = sum >= q.min_replicas;
        }
        cnt += ok;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    }
    if (lane == 0) out[g] = int32_t(cnt);
  }
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < tb.Q; q += gridDim.x * blockDim.x)
    out[tb.G + q] = int32_t(sig_sum[tb.cinfo[q].sig]);
}

// sig_sum[s] = sum of the capacity bytes of signature s over [lo, hi): one CTA per signature
__global__ void __launch_bounds__(256) k_sig_range_sum(Topo tp, const uint8_t* __restrict__ cap8, uint32_t lo, uint32_t hi, uint32_t* __restrict__ sig_sum) {
  __shared__ uint32_t s_w[8];
  const uint8_t* row = cap8 + size_t(blockIdx.x) * tp.npad;
  uint32_t sum = 0;
  for (uint32_t n = lo + threadIdx.x; n < hi; n += 256) sum += row[n];
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < 8; ++i) t += s_w[i]; sig_sum[blockIdx.x] = t; }
}

}  // namespace grove
