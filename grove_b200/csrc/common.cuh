// common.cuh -- shared structs of the placement kernels (device views of the tables, per-round buffers).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/grove_place.h"


namespace grove {

constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr int kMaxPieces = 2 * GROVE_MAX_LEVELS + 3;

struct GangInfo {     // 48 B, built on the host at submit time
  uint32_t anchor;    // sorted node index
  uint32_t order;     // rank by (priority desc, index asc)
  uint32_t pod_off;   // first slot in the entry arrays
  uint32_t pad;
  uint32_t anc_lo[GROVE_MAX_LEVELS];  // node range of the anchor's domain per level; [a,a) if label absent
  uint32_t anc_hi[GROVE_MAX_LEVELS];
};

struct CliqueInfo {   // 16 B
  uint32_t gang;
  uint32_t need_depth;  // labels a candidate node must carry (deepest binding Required level + 1)
  uint32_t sig;         // fit signature: cliques with identical (requests, class mask, need_depth) share a fit row
  uint32_t pad;
};

struct Topo {
  const uint4* nres;       // [npad] dynamic: free_cpu, free_mem, free_gpu | free_pods << 16, flags | vdepth << 16
  const uint4* ndom;       // [npad] static: tree-ified domain index per level
  const uint32_t* dom_lo[GROVE_MAX_LEVELS];
  const uint32_t* dom_hi[GROVE_MAX_LEVELS];
  const uint32_t* next_dom[GROVE_MAX_LEVELS];  // [n+1] first level-l domain starting at or after node i
  uint32_t n_dom[GROVE_MAX_LEVELS];
  uint32_t unit[GROVE_MAX_LEVELS];             // every domain of the level is a single node
  uint32_t n, npad, L, words;                  // words = npad / 32 (row stride of the fit bitmap)
  uint32_t cap_off[GROVE_MAX_LEVELS];          // column offset of level l in a capacity-table row (non-unit levels)
  uint32_t cap_stride;                         // columns per signature row = sum of n_dom over non-unit levels
};

struct Tables {
  const grove_gang_t* gangs;
  const grove_clique_t* cliques;
  const grove_scope_t* scopes;
  const GangInfo* ginfo;
  const CliqueInfo* cinfo;
  const uint4* sigs;     // [S] req_cpu, req_mem, req_gpu, class_mask | need_depth << 16
  uint32_t G, Q, S;
};

struct RoundBufs {
  uint8_t* state;        // [G] GROVE_GANG_*
  uint8_t* round;        // [G]
  uint32_t* active;      // [G] gangs evaluated this round
  uint32_t* rows;        // [Q] clique rows evaluated this round
  uint32_t* counters;    // [0] n_active [1] n_rows [2] unresolved [3] base rejections propagated [4] n_sigs
                         // [5] active gangs over all ranks [6] gangs resolved by this round's apply
  uint32_t* sig_stamp;   // [S] last round in which the signature was active
  uint32_t* sig_list;    // [S] signatures needed this round
  uint8_t* spec_score;   // [G]
  uint16_t* spec_n;      // [G] entries incl. surplus
  uint32_t* spec_top;    // [G]
  uint32_t* ent_node;    // [P]
  uint16_t* ent_meta;    // [P] clique_rel | score << 8
  uint32_t* active_all;  // [G] active gangs of every rank (replicated decision)
  uint32_t* claim;       // [n] (tag << 24 | order rank) of the best proposal for the node; tags decrease per (round, sub-round)
  uint8_t* taken;        // [n] stamp of the last round that committed pods on the node
  uint8_t* cur;          // [G] next alternative a gang will propose
  uint8_t* prop;         // [G] sub-round (1-based) of the gang's last proposal
  uint32_t* flags;       // [GROVE_SUBROUNDS] number of the last round with a proposal in sub-round s
  // exchange buffer of the round (also the all-reduce payload of the sharded cycle), u32 words:
  uint32_t* alt_node;    // [K][P] entry i of alternative a of gang g at a*P + pod_off[g] + i
  uint32_t* alt_meta;    // [K][P] clique_rel | score << 8
  uint32_t* alt_n;       // [G][K] entries incl. surplus
  uint32_t* alt_score;   // [G][K] min score over the MinReplicas entries (written by k_alt_scores)
  uint32_t* alt_nmin;    // [G][K] entries of the MinReplicas phase (the rest is best-effort surplus)
  uint32_t* alt_top;     // [G][K]
  uint32_t* nalt;        // [G]
  uint32_t K, P;
  uint32_t* F;           // [S][words] fit bitmap, one row per signature
  uint8_t* T;            // [Q][npad]
  const uint8_t* cap8;   // [S][npad] pods of the signature that fit on the node now (saturating), or null
  const uint32_t* capsum; // [S][cap_stride] per-domain sum of cap8 (non-unit levels)
  const uint32_t* capmax; // [S][cap_stride] per-domain max of cap8
  uint32_t caps_in_attempts;  // 1: the scalar evaluator packs from cap8 bytes, 0: from fit words + node records
  uint32_t width0;            // candidates attempted in the very first step of a gang (1..32), warp-per-gang form
  uint32_t width1;            // candidates per warp in the first attempt window of the CTA-per-gang forms (1..32)
  uint32_t* dbg;              // [G][8] optional: successes, plausible, attempts, winning candidate, cycles, chunks
};

}  // namespace grove
