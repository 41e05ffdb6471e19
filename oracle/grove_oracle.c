/*
 * grove_oracle.c -- CPU restatement of the gang-placement cycle.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libgrove_place.so) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" at node level.  The reference tree (ai-dynamo/grove @ 08ad3b37)
 * contains no scheduler: placement is done by KAI-Scheduler v0.14.0 (operator/go.mod:11), which is
 * not vendored, and there is no Go toolchain in this image.  What the reference does pin -- and
 * what tests/test_oracle_e2e_properties.py checks this file against -- is
 *   * the input schema             scheduler/api/core/v1alpha1/podgang.go:51-131
 *   * MinReplicas = gang guarantee, surplus best effort                      podgang.go:80-83
 *   * Required pack constraint = all pods of the scope share one label value podgang.go:101-109
 *   * nodes lacking the label are not candidates      docs/proposals/244-topology-aware-scheduling/README.md:65
 *   * level index 0 is the broadest                                          ibid. :143
 *   * scope nesting gang >= group-config >= pod-group   operator/internal/webhook/admission/pcs/validation/topologyconstraints.go:195-202
 *   * all-or-nothing admission               operator/internal/controller/podclique/components/pod/syncflow.go:319-358
 *   * scaled gangs gated behind their base gang                              ibid. :255-314
 *   * the outcome properties of the live-cluster e2e suites GS1-GS12 / TAS2-TAS17
 *     (operator/e2e/tests/gang_scheduling_test.go, topology_test.go), the step-by-step pod counts of the
 *     multi-step suites GS2-GS12 (tests/test_oracle_e2e_sequences.py), and the reference's own workload
 *     files run end to end (tests/test_workload_fixtures.py)
 *   * Preferred = best effort, widening level by level up to Required        podgang.go:110-117
 *     (no reference test exercises it: tests/test_oracle_preferred.py restates the API comment).
 * Everything below those (which node, which domain among feasible ones, the score value) is
 * defined by DESIGN.md "Placement semantics"; this file is its executable form, written as plain
 * scalar loops that derive every ordering from the score matrix itself (no piece iterator, no
 * warp tricks) so that it shares no code or shortcut with the CUDA path.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/grove_place.h"

#define VDEPTH_SHIFT 16u /* internal: valid label depth in flags bits 16..19 */

typedef struct oracle_stats {
  uint32_t rounds;
  uint32_t gangs_admitted;
  uint32_t gangs_rejected;
  uint32_t pods_bound;
  uint64_t pairs_evaluated;
  double seconds_eval;   /* fit+score+admit per-gang evaluation, all rounds */
  double seconds_total;
  uint32_t non_tree_labels; /* raw label ids that appeared under more than one parent */
  uint32_t threads;
} oracle_stats_t;

typedef struct topo {
  uint32_t n, L;
  uint32_t* perm; /* sorted -> caller */
  uint32_t* inv;  /* caller -> sorted */
  grove_node_t* nodes; /* sorted; dom[] tree-ified; flags |= vdepth << 16 */
  uint32_t n_dom[GROVE_MAX_LEVELS];
  uint32_t* dom_lo[GROVE_MAX_LEVELS];
  uint32_t* dom_hi[GROVE_MAX_LEVELS];
  uint32_t non_tree;
} topo_t;

static const grove_node_t* g_sort_nodes;
static uint32_t g_sort_L;
static int cmp_nodes(const void* pa, const void* pb) {
  uint32_t a = *(const uint32_t*)pa, b = *(const uint32_t*)pb;
  for (uint32_t l = 0; l < g_sort_L; ++l) {
    uint32_t x = g_sort_nodes[a].dom[l], y = g_sort_nodes[b].dom[l];
    if (x != y) return x < y ? -1 : 1;
  }
  return a < b ? -1 : (a > b ? 1 : 0);
}

static void topo_free(topo_t* t) {
  free(t->perm); free(t->inv); free(t->nodes);
  for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) { free(t->dom_lo[l]); free(t->dom_hi[l]); }
  memset(t, 0, sizeof(*t));
}

/* Sort nodes by label path (absent labels last), then make every level's domain id the index of the
 * distinct PATH prefix ("tree-ify"): two nodes are in the same level-l domain iff their labels agree
 * on levels 0..l and are all present.  KWOK's 28/20/7 arithmetic (kwok.py:64-68) is not nested; path
 * semantics splits such a block per zone and counts it in non_tree. */
static int topo_build(topo_t* t, const grove_node_t* in, uint32_t n, uint32_t L) {
  memset(t, 0, sizeof(*t));
  t->n = n; t->L = L;
  t->perm = malloc(sizeof(uint32_t) * (n ? n : 1));
  t->inv = malloc(sizeof(uint32_t) * (n ? n : 1));
  t->nodes = malloc(sizeof(grove_node_t) * (n ? n : 1));
  if (!t->perm || !t->inv || !t->nodes) return -1;
  for (uint32_t i = 0; i < n; ++i) t->perm[i] = i;
  g_sort_nodes = in; g_sort_L = L;
  qsort(t->perm, n, sizeof(uint32_t), cmp_nodes);
  for (uint32_t i = 0; i < n; ++i) { t->inv[t->perm[i]] = i; t->nodes[i] = in[t->perm[i]]; }
  for (uint32_t l = 0; l < L; ++l) {
    t->dom_lo[l] = malloc(sizeof(uint32_t) * (n ? n : 1));
    t->dom_hi[l] = malloc(sizeof(uint32_t) * (n ? n : 1));
    if (!t->dom_lo[l] || !t->dom_hi[l]) return -1;
  }
  /* raw label -> first parent seen, to count non-tree label sets (diagnostic only) */
  uint32_t cnt[GROVE_MAX_LEVELS] = {0, 0, 0, 0};
  uint32_t prev_raw[GROVE_MAX_LEVELS], prev_id[GROVE_MAX_LEVELS];
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t depth = 0;
    int parent_same = 1; /* is the tree-ified parent the same as the previous node's? */
    uint32_t raw[GROVE_MAX_LEVELS];
    for (uint32_t l = 0; l < L; ++l) raw[l] = in[t->perm[i]].dom[l];
    int alive = 1;
    for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) {
      if (l >= L || !alive || raw[l] == GROVE_DOM_ABSENT) {
        alive = 0;
        t->nodes[i].dom[l] = GROVE_DOM_ABSENT;
        if (l < L) { prev_raw[l] = GROVE_DOM_ABSENT; prev_id[l] = GROVE_DOM_ABSENT; }
        parent_same = 0;
        continue;
      }
      int same = (i > 0) && parent_same && prev_id[l] != GROVE_DOM_ABSENT && prev_raw[l] == raw[l];
      uint32_t id;
      if (same) {
        id = prev_id[l];
        t->dom_hi[l][id] = i + 1;
      } else {
        id = cnt[l]++;
        t->dom_lo[l][id] = i;
        t->dom_hi[l][id] = i + 1;
      }
      t->nodes[i].dom[l] = id;
      prev_raw[l] = raw[l]; prev_id[l] = id;
      parent_same = same;
      depth = l + 1;
    }
    t->nodes[i].flags = (t->nodes[i].flags & 0xFFFFu) | (depth << VDEPTH_SHIFT);
  }
  for (uint32_t l = 0; l < L; ++l) t->n_dom[l] = cnt[l];
  /* non-tree diagnostic: a raw id at level l>0 owned by >1 tree domain */
  t->non_tree = 0;
  for (uint32_t l = 1; l < L; ++l) {
    /* collect (raw) of every domain's first node, sort, count duplicates */
    uint32_t m = cnt[l];
    if (m < 2) continue;
    uint32_t* raws = malloc(sizeof(uint32_t) * m);
    if (!raws) return -1;
    for (uint32_t d = 0; d < m; ++d) raws[d] = in[t->perm[t->dom_lo[l][d]]].dom[l];
    /* shell sort (keeps this file free of a second comparator) */
    for (uint32_t gap = m / 2; gap > 0; gap /= 2)
      for (uint32_t x = gap; x < m; ++x) {
        uint32_t v = raws[x]; uint32_t y = x;
        while (y >= gap && raws[y - gap] > v) { raws[y] = raws[y - gap]; y -= gap; }
        raws[y] = v;
      }
    for (uint32_t x = 1; x < m; ++x) if (raws[x] == raws[x - 1]) t->non_tree++;
    free(raws);
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------- */
typedef struct ctx {
  topo_t T;
  uint32_t G, Q, S;
  const grove_gang_t* gangs;
  const grove_clique_t* cliques;
  const grove_scope_t* scopes;
  uint32_t* order;   /* gang -> rank by (priority desc, index asc) */
  uint32_t* anchor;  /* gang -> sorted node index */
  uint32_t* pod_off; /* gang -> first slot */
} ctx_t;

typedef struct entry { uint32_t node; uint8_t clique_rel; uint8_t score; } entry_t;

/* Exchange buffer of one round (int32 words; also the all-reduce payload of the multi-rank protocol):
 *   [0, K*P)          alt_node   entry i of alternative a of gang g at a*P + pod_off[g] + i (sorted node index)
 *   [K*P, 2*K*P)      alt_meta   clique_rel | score << 8
 *   then G*K words each: alt_n (entries incl. surplus), alt_score (min score over MinReplicas pods),
 *   alt_top (first sorted node of the gang domain); then G words nalt. */
typedef struct xlay { size_t node, meta, n, score, top, nalt, words; uint32_t K; } xlay_t;
static xlay_t xlayout(uint32_t K, uint32_t P, uint32_t G) {
  xlay_t x; x.K = K;
  x.node = 0; x.meta = (size_t)K * P; x.n = 2 * (size_t)K * P; x.score = x.n + (size_t)G * K; x.top = x.score + (size_t)G * K;
  x.nalt = x.top + (size_t)G * K; x.words = x.nalt + G;
  return x;
}

static uint32_t fmix32(uint32_t x) {
  x = x * 0x9E3779B1u + 0x7F4A7C15u;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}

static inline uint32_t node_vdepth(const grove_node_t* nd) { return (nd->flags >> VDEPTH_SHIFT) & 0xFu; }
static inline uint32_t node_class(const grove_node_t* nd) { return (nd->flags & GROVE_NODE_CLASS_MASK) >> GROVE_NODE_CLASS_SHIFT; }

/* deepest Required level that binds clique q (its own, its scope's, its gang's) + 1; 0 if none.
 * A node must carry labels down to that depth to be a candidate (GREP-244 README.md:65). */
static uint32_t need_depth(const grove_gang_t* g, const grove_scope_t* s, const grove_clique_t* q) {
  uint32_t d = 0;
  if (g->level != GROVE_LEVEL_NONE && g->level + 1u > d) d = g->level + 1u;
  if (s->level != GROVE_LEVEL_NONE && s->level + 1u > d) d = s->level + 1u;
  if (q->level != GROVE_LEVEL_NONE && q->level + 1u > d) d = q->level + 1u;
  return d;
}

static int static_ok(const grove_node_t* nd, const grove_clique_t* q, uint32_t nd_need) {
  if (!(nd->flags & GROVE_NODE_SCHEDULABLE)) return 0;
  if (!((q->class_mask >> node_class(nd)) & 1u)) return 0;
  if (node_vdepth(nd) < nd_need) return 0;
  return 1;
}

/* K1 semantics: can node nd host at least one pod of q right now? */
static int fit(const grove_node_t* nd, const grove_clique_t* q, uint32_t nd_need) {
  return static_ok(nd, q, nd_need) && nd->free_cpu_milli >= q->req_cpu_milli &&
         nd->free_mem_mib >= q->req_mem_mib && nd->free_gpu >= q->req_gpu && nd->free_pods >= 1;
}

/* K2 semantics: number of levels at which n and the anchor share a (tree-ified) domain */
static uint32_t closeness(const topo_t* T, uint32_t n, uint32_t a) {
  uint32_t c = 0;
  for (uint32_t l = 0; l < T->L; ++l) {
    uint32_t x = T->nodes[n].dom[l];
    if (x != GROVE_DOM_ABSENT && x == T->nodes[a].dom[l]) c++;
  }
  return c;
}

/* Levels a unit (gang / scope / clique) with constraint (req, pref) is tried at inside a parent range of
 * level `lvl` (-1 = the whole cluster): from *first down to the returned base.  base is the hard level
 * (Required if it is deeper than the parent's, else the parent range itself); Preferred, when deeper than
 * base, is tried first and widened level by level up to base (podgang.go:110-117: "Scheduler can fall back
 * to higher topology levels (upto Required constraint) if preferred cannot be satisfied"). */
static int level_span(uint32_t req, uint32_t pref, int lvl, int* first) {
  int base = (req != GROVE_LEVEL_NONE && (int)req > lvl) ? (int)req : lvl;
  *first = (pref != GROVE_LEVEL_NONE && (int)pref > base) ? (int)pref : base;
  return base;
}
static inline uint32_t scope_pref(const grove_scope_t* s) { return s->preferred1 ? (uint32_t)s->preferred1 - 1u : GROVE_LEVEL_NONE; }

typedef struct geval {
  const ctx_t* C;
  uint32_t g;
  uint32_t a;
  uint8_t* Trow[GROVE_MAX_GANG_CLIQUES]; /* score rows of this gang's cliques, round-start state */
  uint32_t ndepth[GROVE_MAX_GANG_CLIQUES];
  entry_t st[GROVE_MAX_GANG_PODS];
  uint32_t np;
  uint32_t Hlo[GROVE_MAX_GANG_CLIQUES], Hhi[GROVE_MAX_GANG_CLIQUES];
} geval_t;

/* how many more pods of clique cr fit on node n given the pods this gang already put there */
static uint32_t cap_now(const geval_t* E, uint32_t cr, uint32_t n) {
  const ctx_t* C = E->C;
  const grove_gang_t* g = &C->gangs[E->g];
  const grove_clique_t* q = &C->cliques[g->clique_off + cr];
  const grove_node_t* nd = &C->T.nodes[n];
  if (!static_ok(nd, q, E->ndepth[cr])) return 0;
  uint64_t ucpu = 0, umem = 0, ugpu = 0, upods = 0;
  for (uint32_t i = 0; i < E->np; ++i)
    if (E->st[i].node == n) {
      const grove_clique_t* o = &C->cliques[g->clique_off + E->st[i].clique_rel];
      ucpu += o->req_cpu_milli; umem += o->req_mem_mib; ugpu += o->req_gpu; upods += 1;
    }
  if (nd->free_cpu_milli < ucpu || nd->free_mem_mib < umem || nd->free_gpu < ugpu || nd->free_pods < upods) return 0;
  uint64_t c = nd->free_pods - upods;
  if (q->req_cpu_milli) { uint64_t k = (nd->free_cpu_milli - ucpu) / q->req_cpu_milli; if (k < c) c = k; }
  if (q->req_mem_mib) { uint64_t k = (nd->free_mem_mib - umem) / q->req_mem_mib; if (k < c) c = k; }
  if (q->req_gpu) { uint64_t k = (nd->free_gpu - ugpu) / q->req_gpu; if (k < c) c = k; }
  return (uint32_t)c;
}

/* Put up to `want` pods of clique cr on the fit nodes of [lo,hi), visiting them in descending score,
 * ties by ascending rotated index (n - anchor) mod N.  Returns pods placed. */
static uint32_t take(geval_t* E, uint32_t cr, uint32_t lo, uint32_t hi, uint32_t want) {
  const topo_t* T = &E->C->T;
  uint32_t placed = 0;
  if (want == 0 || hi <= lo) return 0;
  uint32_t len = hi - lo;
  uint32_t start = (E->a >= lo && E->a < hi) ? E->a - lo : 0; /* ascending rot == ascending from the anchor, wrapping */
  for (uint32_t s = T->L + 1; s >= 1 && placed < want; --s) {
    for (uint32_t k = 0; k < len && placed < want; ++k) {
      uint32_t n = lo + (start + k) % len;
      if (E->Trow[cr][n] != s) continue;
      uint32_t c = cap_now(E, cr, n);
      uint32_t t = c < (want - placed) ? c : (want - placed);
      for (uint32_t j = 0; j < t; ++j) {
        E->st[E->np].node = n; E->st[E->np].clique_rel = (uint8_t)cr; E->st[E->np].score = (uint8_t)s;
        E->np++;
      }
      placed += t;
    }
  }
  return placed;
}

static int fill_min(geval_t* E, uint32_t cr, uint32_t lo, uint32_t hi) {
  const grove_gang_t* g = &E->C->gangs[E->g];
  uint32_t m = E->C->cliques[g->clique_off + cr].min_replicas;
  uint32_t mark = E->np;
  if (take(E, cr, lo, hi, m) < m) { E->np = mark; return 0; }
  E->Hlo[cr] = lo; E->Hhi[cr] = hi;
  return 1;
}

typedef struct cand { uint32_t sc; uint32_t rot; uint32_t lo, hi; } cand_t;
static int cmp_cand(const void* pa, const void* pb) {
  const cand_t* a = pa; const cand_t* b = pb;
  if (a->sc != b->sc) return a->sc > b->sc ? -1 : 1;
  if (a->rot != b->rot) return a->rot < b->rot ? -1 : 1;
  return 0;
}

/* Level-l domains inside [lo,hi), ordered by descending score of the domain (closeness of its nodes
 * to the anchor, which is uniform outside the anchor's own level-l domain and capped at l+1 inside
 * it), ties by ascending rotated index of the domain's first node. */
static cand_t* subdomains(const geval_t* E, uint32_t l, uint32_t lo, uint32_t hi, uint32_t* n_out) {
  const topo_t* T = &E->C->T;
  uint32_t cnt = 0;
  for (uint32_t n = lo; n < hi; ++n) {
    uint32_t d = T->nodes[n].dom[l];
    if (d != GROVE_DOM_ABSENT && T->dom_lo[l][d] == n) cnt++;
  }
  cand_t* v = malloc(sizeof(cand_t) * (cnt ? cnt : 1));
  uint32_t k = 0;
  for (uint32_t n = lo; n < hi; ++n) {
    uint32_t d = T->nodes[n].dom[l];
    if (d == GROVE_DOM_ABSENT || T->dom_lo[l][d] != n) continue;
    uint32_t c = closeness(T, n, E->a);
    v[k].sc = c < l + 1 ? c : l + 1;
    v[k].rot = (n + T->n - E->a) % T->n;
    v[k].lo = n; v[k].hi = T->dom_hi[l][d];
    k++;
  }
  qsort(v, cnt, sizeof(cand_t), cmp_cand);
  *n_out = cnt;
  return v;
}

/* cliques of one scope inside range E_=[lo,hi) whose level is `lvl` (-1 = ROOT) */
static int place_scope(geval_t* E, const grove_scope_t* s, uint32_t lo, uint32_t hi, int lvl) {
  const grove_gang_t* g = &E->C->gangs[E->g];
  uint32_t mark = E->np;
  for (uint32_t i = 0; i < s->n_cliques; ++i) {
    uint32_t cr = s->first_clique + i;
    const grove_clique_t* q = &E->C->cliques[g->clique_off + cr];
    int ok = 0, first, base = level_span(q->level, GROVE_CLIQUE_PREFERRED(q->scope), lvl, &first);
    for (int l = first; l >= base && !ok; --l) {
      if (l > lvl) {
        uint32_t nc; cand_t* v = subdomains(E, (uint32_t)l, lo, hi, &nc);
        for (uint32_t k = 0; k < nc && !ok; ++k) ok = fill_min(E, cr, v[k].lo, v[k].hi);
        free(v);
      } else {
        ok = fill_min(E, cr, lo, hi);
      }
    }
    if (!ok) { E->np = mark; return 0; }
  }
  return 1;
}

static int place_in(geval_t* E, uint32_t lo, uint32_t hi, int lvl) {
  const ctx_t* C = E->C;
  const grove_gang_t* g = &C->gangs[E->g];
  E->np = 0;
  for (uint32_t si = 0; si < g->n_scopes; ++si) {
    const grove_scope_t* s = &C->scopes[g->scope_off + si];
    int ok = 0, first, base = level_span(s->level, scope_pref(s), lvl, &first);
    for (int l = first; l >= base && !ok; --l) {
      if (l > lvl) {
        uint32_t nc; cand_t* v = subdomains(E, (uint32_t)l, lo, hi, &nc);
        for (uint32_t k = 0; k < nc && !ok; ++k) ok = place_scope(E, s, v[k].lo, v[k].hi, l);
        free(v);
      } else {
        ok = place_scope(E, s, lo, hi, lvl);
      }
    }
    if (!ok) { E->np = 0; return 0; }
  }
  return 1;
}

/* surplus beyond MinReplicas (best effort, podgang.go:80-83) and publication of one alternative */
static void emit_alt(geval_t* E, const xlay_t* X, int32_t* xb, uint32_t P, uint32_t pod_off, uint32_t a, uint32_t top_lo) {
  const ctx_t* C = E->C;
  const grove_gang_t* g = &C->gangs[E->g];
  uint32_t min_score = C->T.L + 1;
  for (uint32_t i = 0; i < E->np; ++i) if (E->st[i].score < min_score) min_score = E->st[i].score;
  for (uint32_t cr = 0; cr < g->n_cliques; ++cr) {
    const grove_clique_t* q = &C->cliques[g->clique_off + cr];
    uint32_t extra = q->replicas > q->min_replicas ? (uint32_t)(q->replicas - q->min_replicas) : 0;
    if (extra) take(E, cr, E->Hlo[cr], E->Hhi[cr], extra);
  }
  for (uint32_t i = 0; i < E->np; ++i) {
    xb[X->node + (size_t)a * P + pod_off + i] = (int32_t)E->st[i].node;
    xb[X->meta + (size_t)a * P + pod_off + i] = (int32_t)((uint32_t)E->st[i].clique_rel | ((uint32_t)E->st[i].score << 8));
  }
  xb[X->n + (size_t)E->g * X->K + a] = (int32_t)E->np;
  xb[X->score + (size_t)E->g * X->K + a] = (int32_t)min_score;
  xb[X->top + (size_t)E->g * X->K + a] = (int32_t)top_lo;
}

/* one gang against the round-start state: its first K feasible gang-level domains in score order, each
 * packed independently ("alternatives"); a gang without a gang-level constraint has one candidate (the
 * whole cluster).  With a Preferred level the candidate list is the Preferred level's domains in score
 * order, then each wider level's, down to the Required level (or the whole cluster). */
static void eval_gang(const ctx_t* C, uint32_t gi, uint8_t* const* Trow, const xlay_t* X, int32_t* xb, uint32_t P) {
  geval_t* E = malloc(sizeof(geval_t));
  memset(E, 0, sizeof(*E));
  const grove_gang_t* g = &C->gangs[gi];
  E->C = C; E->g = gi; E->a = C->anchor[gi];
  for (uint32_t si = 0; si < g->n_scopes; ++si) {
    const grove_scope_t* s = &C->scopes[g->scope_off + si];
    for (uint32_t i = 0; i < s->n_cliques; ++i) {
      uint32_t cr = s->first_clique + i;
      E->ndepth[cr] = need_depth(g, s, &C->cliques[g->clique_off + cr]);
    }
  }
  for (uint32_t cr = 0; cr < g->n_cliques; ++cr) E->Trow[cr] = Trow[cr];
  uint32_t na = 0;
  int first, base = level_span(g->level, g->preferred, -1, &first);
  for (int l = first; l >= base && na < X->K; --l) {
    if (l < 0) {
      if (place_in(E, 0, C->T.n, -1)) emit_alt(E, X, xb, P, C->pod_off[gi], na++, 0);
    } else {
      uint32_t nc; cand_t* v = subdomains(E, (uint32_t)l, 0, C->T.n, &nc);
      for (uint32_t k = 0; k < nc && na < X->K; ++k)
        if (place_in(E, v[k].lo, v[k].hi, l)) emit_alt(E, X, xb, P, C->pod_off[gi], na++, v[k].lo);
      free(v);
    }
  }
  xb[X->nalt + gi] = (int32_t)na;
  free(E);
}

static double now_s(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

typedef struct ord { int32_t pr; uint32_t g; } ord_t;
static int cmp_ord(const void* pa, const void* pb) {
  const ord_t* a = pa; const ord_t* b = pb;
  if (a->pr != b->pr) return a->pr > b->pr ? -1 : 1;
  return a->g < b->g ? -1 : (a->g > b->g ? 1 : 0);
}

int32_t oracle_validate(const grove_gang_t* gangs, uint32_t G, const grove_clique_t* cliques, uint32_t Q,
                        const grove_scope_t* scopes, uint32_t S, uint32_t L, uint32_t n_nodes) {
  if (L < 1 || L > GROVE_MAX_LEVELS) return GROVE_ERR_INVALID_ARG;
  for (uint32_t gi = 0; gi < G; ++gi) {
    const grove_gang_t* g = &gangs[gi];
    if (g->n_cliques == 0 || g->n_cliques > GROVE_MAX_GANG_CLIQUES) return GROVE_ERR_LIMIT;
    if (g->n_scopes == 0 || g->n_scopes > GROVE_MAX_GANG_SCOPES) return GROVE_ERR_LIMIT;
    if ((uint64_t)g->clique_off + g->n_cliques > Q || (uint64_t)g->scope_off + g->n_scopes > S) return GROVE_ERR_INVALID_ARG;
    if (g->level != GROVE_LEVEL_NONE && g->level >= L) return GROVE_ERR_INVALID_ARG;
    if (g->preferred != GROVE_LEVEL_NONE && (g->preferred >= L || (g->level != GROVE_LEVEL_NONE && g->preferred <= g->level))) return GROVE_ERR_INVALID_ARG;
    if (g->anchor_node != GROVE_NONE_U32 && g->anchor_node >= n_nodes) return GROVE_ERR_INVALID_ARG;
    if (g->base_gang != GROVE_NONE_U32 && (g->base_gang >= G || g->base_gang == gi)) return GROVE_ERR_INVALID_ARG;
    uint32_t pods = 0, next = 0;
    for (uint32_t si = 0; si < g->n_scopes; ++si) {
      const grove_scope_t* s = &scopes[g->scope_off + si];
      if (s->first_clique != next || s->n_cliques == 0) return GROVE_ERR_INVALID_ARG; /* scopes tile the gang's cliques in order */
      if (s->level != GROVE_LEVEL_NONE && s->level >= L) return GROVE_ERR_INVALID_ARG;
      if (s->preferred1 && (s->preferred1 > L || (s->level != GROVE_LEVEL_NONE && s->preferred1 - 1u <= s->level))) return GROVE_ERR_INVALID_ARG;
      for (uint32_t i = 0; i < s->n_cliques; ++i) {
        if (next + i >= g->n_cliques) return GROVE_ERR_INVALID_ARG;
        const grove_clique_t* q = &cliques[g->clique_off + next + i];
        if (GROVE_CLIQUE_SCOPE(q->scope) != si) return GROVE_ERR_INVALID_ARG;
        if (q->level != GROVE_LEVEL_NONE && q->level >= L) return GROVE_ERR_INVALID_ARG;
        { uint32_t qp = GROVE_CLIQUE_PREFERRED(q->scope);
          if (qp != GROVE_LEVEL_NONE && (qp >= L || (q->level != GROVE_LEVEL_NONE && qp <= q->level))) return GROVE_ERR_INVALID_ARG; }
        if (q->replicas < q->min_replicas) return GROVE_ERR_INVALID_ARG;
        pods += q->replicas;
      }
      next += s->n_cliques;
    }
    if (next != g->n_cliques) return GROVE_ERR_INVALID_ARG;
    if (pods > GROVE_MAX_GANG_PODS) return GROVE_ERR_LIMIT;
  }
  return GROVE_OK;
}

/*
 * One scheduling cycle.  Optimistic rounds (DESIGN.md "Cycle"):
 *   round: every active gang is evaluated against the round-start node state (fit -> score -> its
 *   first K feasible domains, each packed = K alternatives); then up to GROVE_SUBROUNDS sub-rounds
 *   resolve conflicts without re-evaluating: every undecided gang proposes its first alternative
 *   that touches no node committed earlier in this round, proposals claim their nodes with the gang's
 *   order rank (min wins), a gang that holds every node it claimed commits.  Gangs left undecided are
 *   re-evaluated next round; a gang with no feasible domain is rejected for the cycle.
 *
 * Written as steps over a shard context so that the multi-rank protocol (gang rows dealt g % world
 * to ranks, node table and gang state replicated -- DESIGN.md section 7) can be exercised on CPU:
 *   begin; repeat { eval (own gangs -> exchange buffer) -> [all-reduce SUM] -> resolve (replicated) }; end.
 * oracle_run_cycle is the same steps with world = 1 and no reduction.
 */
#define CLAIM_NONE 0x7F7F7F7F /* > any order rank (< 2^24) */

typedef struct oshard {
  ctx_t C;
  uint32_t n, L, G, Q, S, P, K;
  xlay_t X;
  const grove_node_t* nodes_in;
  uint32_t rank, world, max_rounds;
  uint8_t* state; uint8_t* rnd;
  uint32_t* active; uint32_t na_local;        /* this rank's share of the round */
  uint32_t* active_all; uint32_t na_all;      /* every rank's (replicated decision) */
  int32_t* fin_node; int32_t* fin_meta; uint32_t* fin_n; uint8_t* fin_score; uint32_t* fin_top;
  uint32_t round, unresolved;
  uint64_t pairs;
  double t0, t_eval;
  uint32_t* out_fit; uint8_t* out_score;
} oshard_t;

static void shard_free(oshard_t* h) {
  if (!h) return;
  free(h->active); free(h->active_all); free(h->rnd); free(h->state);
  free(h->fin_node); free(h->fin_meta); free(h->fin_n); free(h->fin_score); free(h->fin_top);
  free(h->C.order); free(h->C.anchor); free(h->C.pod_off);
  topo_free(&h->C.T);
  free(h);
}

void oracle_shard_abort(oshard_t* h) { shard_free(h); }

int32_t oracle_shard_begin(const grove_node_t* nodes_in, uint32_t n, uint32_t L, const grove_gang_t* gangs, uint32_t G,
                           const grove_clique_t* cliques, uint32_t Q, const grove_scope_t* scopes, uint32_t S,
                           uint32_t max_rounds, uint32_t alternatives, int32_t threads, uint32_t rank, uint32_t world,
                           uint32_t* out_fit, uint8_t* out_score, oshard_t** out) {
  int32_t rc = oracle_validate(gangs, G, cliques, Q, scopes, S, L, n);
  if (rc != GROVE_OK) return rc;
  if (n == 0 || n > GROVE_MAX_NODES || !out) return GROVE_ERR_INVALID_ARG;
  if (alternatives == 0) alternatives = GROVE_MAX_ALTERNATIVES;
  if (alternatives > GROVE_MAX_ALTERNATIVES) return GROVE_ERR_INVALID_ARG;
  if (world == 0) world = 1;
  if (rank >= world) return GROVE_ERR_INVALID_ARG;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  oshard_t* h = calloc(1, sizeof(oshard_t));
  if (!h) return GROVE_ERR_OOM;
  h->t0 = now_s();
  if (topo_build(&h->C.T, nodes_in, n, L)) { shard_free(h); return GROVE_ERR_OOM; }
  h->n = n; h->L = L; h->G = G; h->Q = Q; h->S = S; h->nodes_in = nodes_in; h->K = alternatives;
  h->rank = rank; h->world = world; h->max_rounds = max_rounds; h->out_fit = out_fit; h->out_score = out_score;
  ctx_t* C = &h->C;
  C->G = G; C->Q = Q; C->S = S; C->gangs = gangs; C->cliques = cliques; C->scopes = scopes;
  C->order = malloc(sizeof(uint32_t) * (G ? G : 1));
  C->anchor = malloc(sizeof(uint32_t) * (G ? G : 1));
  C->pod_off = malloc(sizeof(uint32_t) * (G + 1));
  ord_t* ov = malloc(sizeof(ord_t) * (G ? G : 1));
  for (uint32_t g = 0; g < G; ++g) { ov[g].pr = gangs[g].priority; ov[g].g = g; }
  qsort(ov, G, sizeof(ord_t), cmp_ord);
  for (uint32_t r = 0; r < G; ++r) C->order[ov[r].g] = r;
  free(ov);
  uint32_t po = 0;
  for (uint32_t g = 0; g < G; ++g) {
    C->anchor[g] = gangs[g].anchor_node != GROVE_NONE_U32 ? C->T.inv[gangs[g].anchor_node] : fmix32(g) % n;
    C->pod_off[g] = po;
    for (uint32_t c = 0; c < gangs[g].n_cliques; ++c) po += cliques[gangs[g].clique_off + c].replicas;
  }
  C->pod_off[G] = po; h->P = po;
  h->X = xlayout(h->K, h->P, G);
  h->state = calloc(G ? G : 1, 1);
  h->rnd = calloc(G ? G : 1, 1);
  h->active = malloc(sizeof(uint32_t) * (G ? G : 1));
  h->active_all = malloc(sizeof(uint32_t) * (G ? G : 1));
  h->fin_node = calloc(po ? po : 1, sizeof(int32_t)); h->fin_meta = calloc(po ? po : 1, sizeof(int32_t));
  h->fin_n = calloc(G ? G : 1, sizeof(uint32_t)); h->fin_score = calloc(G ? G : 1, 1); h->fin_top = calloc(G ? G : 1, sizeof(uint32_t));
  for (uint32_t g = 0; g < G; ++g) {
    if (gangs[g].flags & GROVE_GANG_GATED) h->state[g] = GROVE_GANG_GATED_SKIP; else h->unresolved++;
  }
  *out = h;
  return GROVE_OK;
}

uint32_t oracle_shard_xbuf_words(const oshard_t* h) { return (uint32_t)h->X.words; }

/* Step 1: decide who is active (replicated state => identical on every rank), evaluate this rank's
 * share into the exchange buffer (zero elsewhere).  *go = 0 when the cycle is over. */
int32_t oracle_shard_eval(oshard_t* h, int32_t* xb, uint32_t* go) {
  ctx_t* C = &h->C;
  const grove_gang_t* gangs = C->gangs; const grove_clique_t* cliques = C->cliques; const grove_scope_t* scopes = C->scopes;
  const uint32_t G = h->G, n = h->n;
  *go = 0; h->na_local = 0; h->na_all = 0;
  if (h->unresolved == 0 || (h->max_rounds && h->round >= h->max_rounds)) return GROVE_OK;
  h->round++;
  const uint8_t r8 = (uint8_t)(h->round > 255 ? 255 : h->round);
  /* scaled gangs become active once their base gang is admitted (pod/syncflow.go:319-358) */
  int changed = 1;
  while (changed) { /* propagate base rejections transitively */
    changed = 0;
    for (uint32_t g = 0; g < G; ++g) {
      if (h->state[g] != GROVE_GANG_PENDING || gangs[g].base_gang == GROVE_NONE_U32) continue;
      uint8_t bs = h->state[gangs[g].base_gang];
      if (bs == GROVE_GANG_REJECTED || bs == GROVE_GANG_BASE_REJECTED || bs == GROVE_GANG_GATED_SKIP) {
        h->state[g] = GROVE_GANG_BASE_REJECTED; h->rnd[g] = r8; h->unresolved--; changed = 1;
      }
    }
  }
  for (uint32_t g = 0; g < G; ++g) {
    if (h->state[g] != GROVE_GANG_PENDING) continue;
    if (gangs[g].base_gang != GROVE_NONE_U32 && h->state[gangs[g].base_gang] != GROVE_GANG_ADMITTED) continue;
    h->active_all[h->na_all++] = g;
    if (g % h->world == h->rank) h->active[h->na_local++] = g;
  }
  if (h->na_all == 0) { /* dependency cycle: nothing can ever become active */
    for (uint32_t g = 0; g < G; ++g)
      if (h->state[g] == GROVE_GANG_PENDING) { h->state[g] = GROVE_GANG_BASE_REJECTED; h->rnd[g] = r8; h->unresolved--; }
    return GROVE_OK;
  }
  *go = 1;
  memset(xb, 0, sizeof(int32_t) * h->X.words);
  const uint32_t words = (n + 31) / 32;
  uint64_t pairs = 0;
  double te0 = now_s();
  const uint32_t na = h->na_local;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : pairs)
  for (uint32_t ai = 0; ai < na; ++ai) {
    uint32_t gi = h->active[ai];
    const grove_gang_t* g = &gangs[gi];
    uint8_t* Trow[GROVE_MAX_GANG_CLIQUES];
    uint32_t* Frow = malloc(sizeof(uint32_t) * words);
    for (uint32_t si = 0; si < g->n_scopes; ++si) {
      const grove_scope_t* s = &scopes[g->scope_off + si];
      for (uint32_t i = 0; i < s->n_cliques; ++i) {
        uint32_t cr = s->first_clique + i;
        const grove_clique_t* q = &cliques[g->clique_off + cr];
        uint32_t ndp = need_depth(g, s, q);
        Trow[cr] = malloc(n);
        memset(Frow, 0, sizeof(uint32_t) * words);
        /* K1: fit bitmap row; K2: score row = fit ? closeness + 1 : 0 (one pass over the node table) */
        for (uint32_t nn = 0; nn < n; ++nn) {
          const int f = fit(&C->T.nodes[nn], q, ndp);
          if (f) Frow[nn >> 5] |= 1u << (nn & 31);
          Trow[cr][nn] = f ? (uint8_t)(closeness(&C->T, nn, C->anchor[gi]) + 1) : 0;
        }
        pairs += n;
        if (h->round == 1 && h->out_fit) memcpy(h->out_fit + (size_t)(g->clique_off + cr) * words, Frow, sizeof(uint32_t) * words);
        if (h->round == 1 && h->out_score) memcpy(h->out_score + (size_t)(g->clique_off + cr) * n, Trow[cr], n);
      }
    }
    eval_gang(C, gi, Trow, &h->X, xb, h->P);
    for (uint32_t cr = 0; cr < g->n_cliques; ++cr) free(Trow[cr]);
    free(Frow);
  }
  h->pairs += pairs;
  h->t_eval += now_s() - te0;
  return GROVE_OK;
}

/* Step 2 (exchange buffer summed over ranks): conflict resolution and commits, identical on every rank */
int32_t oracle_shard_resolve(oshard_t* h, const int32_t* xb, uint32_t* remaining) {
  ctx_t* C = &h->C;
  const xlay_t* X = &h->X;
  const uint32_t n = h->n, P = h->P, K = h->K;
  const uint8_t r8 = (uint8_t)(h->round > 255 ? 255 : h->round);
  uint8_t* taken = calloc(n, 1);
  int32_t* claim = malloc(sizeof(int32_t) * n);
  uint32_t* cur = calloc(h->na_all ? h->na_all : 1, sizeof(uint32_t));
  uint8_t* prop = calloc(h->na_all ? h->na_all : 1, 1);
  for (uint32_t ai = 0; ai < h->na_all; ++ai) {
    uint32_t g = h->active_all[ai];
    if (xb[X->nalt + g] == 0) { h->state[g] = GROVE_GANG_REJECTED; h->rnd[g] = r8; h->unresolved--; }
  }
  for (uint32_t sub = 0; sub < GROVE_SUBROUNDS; ++sub) {
    for (uint32_t i = 0; i < n; ++i) claim[i] = CLAIM_NONE;
    uint32_t nprop = 0;
    for (uint32_t ai = 0; ai < h->na_all; ++ai) {
      uint32_t g = h->active_all[ai];
      prop[ai] = 0;
      if (h->state[g] != GROVE_GANG_PENDING) continue;
      const uint32_t nalt = (uint32_t)xb[X->nalt + g], po = C->pod_off[g];
      while (cur[ai] < nalt) { /* first alternative that touches no node committed earlier in this round */
        const uint32_t a = cur[ai], cnt = (uint32_t)xb[X->n + (size_t)g * K + a];
        int hit = 0;
        for (uint32_t i = 0; i < cnt && !hit; ++i) hit = taken[(uint32_t)xb[X->node + (size_t)a * P + po + i]];
        if (!hit) break;
        cur[ai]++;
      }
      if (cur[ai] >= nalt) continue; /* nothing left to propose: re-evaluated next round */
      prop[ai] = 1; nprop++;
      const uint32_t a = cur[ai], cnt = (uint32_t)xb[X->n + (size_t)g * K + a];
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t nd = (uint32_t)xb[X->node + (size_t)a * P + po + i];
        if ((int32_t)C->order[g] < claim[nd]) claim[nd] = (int32_t)C->order[g];
      }
    }
    if (nprop == 0) break;
    for (uint32_t ai = 0; ai < h->na_all; ++ai) {
      if (!prop[ai]) continue;
      uint32_t g = h->active_all[ai];
      const uint32_t a = cur[ai], cnt = (uint32_t)xb[X->n + (size_t)g * K + a], po = C->pod_off[g];
      int win = 1;
      for (uint32_t i = 0; i < cnt && win; ++i) win = claim[(uint32_t)xb[X->node + (size_t)a * P + po + i]] == (int32_t)C->order[g];
      if (!win) continue;
      for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t nd = (uint32_t)xb[X->node + (size_t)a * P + po + i];
        int32_t meta = xb[X->meta + (size_t)a * P + po + i];
        const grove_clique_t* q = &C->cliques[C->gangs[g].clique_off + ((uint32_t)meta & 0xFFu)];
        grove_node_t* node = &C->T.nodes[nd];
        node->free_cpu_milli -= q->req_cpu_milli; node->free_mem_mib -= q->req_mem_mib;
        node->free_gpu -= q->req_gpu; node->free_pods -= 1;
        taken[nd] = 1;
        h->fin_node[po + i] = (int32_t)nd; h->fin_meta[po + i] = meta;
      }
      h->fin_n[g] = cnt; h->fin_score[g] = (uint8_t)xb[X->score + (size_t)g * K + a]; h->fin_top[g] = (uint32_t)xb[X->top + (size_t)g * K + a];
      h->state[g] = GROVE_GANG_ADMITTED; h->rnd[g] = r8; h->unresolved--;
    }
  }
  free(taken); free(claim); free(cur); free(prop);
  if (remaining) *remaining = h->unresolved;
  return GROVE_OK;
}

/* Step 3: outputs in caller node indices; frees the context */
int32_t oracle_shard_end(oshard_t* h, grove_placement_t* out_pl, uint32_t cap_pl, uint32_t* n_pl,
                         grove_gang_status_t* out_status, grove_node_t* out_nodes, uint32_t* out_perm, oracle_stats_t* stats) {
  ctx_t* C = &h->C;
  const uint32_t G = h->G, n = h->n;
  uint32_t np = 0, adm = 0, rej = 0;
  for (uint32_t g = 0; g < G; ++g) {
    grove_gang_status_t st; memset(&st, 0, sizeof(st));
    st.state = h->state[g]; st.round = h->rnd[g]; st.top_domain_lo = GROVE_NONE_U32; st.placement_off = np;
    if (h->state[g] == GROVE_GANG_ADMITTED) {
      adm++;
      uint32_t cnt = h->fin_n[g];
      st.score_num = h->fin_score[g]; st.score_den = (uint8_t)(h->L + 1);
      st.n_pods = cnt; st.top_domain_lo = h->fin_top[g];
      for (uint32_t i = 0; i < cnt; ++i) {
        if (out_pl && np < cap_pl) {
          out_pl[np].clique = C->gangs[g].clique_off + ((uint32_t)h->fin_meta[C->pod_off[g] + i] & 0xFFu);
          out_pl[np].node = C->T.perm[(uint32_t)h->fin_node[C->pod_off[g] + i]];
        }
        np++;
      }
    } else if (h->state[g] == GROVE_GANG_REJECTED || h->state[g] == GROVE_GANG_BASE_REJECTED) rej++;
    if (out_status) out_status[g] = st;
  }
  if (n_pl) *n_pl = np;
  if (out_nodes)
    for (uint32_t i = 0; i < n; ++i) {
      grove_node_t nd = h->nodes_in[C->T.perm[i]];
      nd.free_cpu_milli = C->T.nodes[i].free_cpu_milli; nd.free_mem_mib = C->T.nodes[i].free_mem_mib;
      nd.free_gpu = C->T.nodes[i].free_gpu; nd.free_pods = C->T.nodes[i].free_pods;
      out_nodes[C->T.perm[i]] = nd;
    }
  if (out_perm) memcpy(out_perm, C->T.perm, sizeof(uint32_t) * n);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->rounds = h->round; stats->gangs_admitted = adm; stats->gangs_rejected = rej; stats->pods_bound = np;
    stats->pairs_evaluated = h->pairs; stats->seconds_eval = h->t_eval; stats->seconds_total = now_s() - h->t0;
    stats->non_tree_labels = C->T.non_tree;
#ifdef _OPENMP
    stats->threads = (uint32_t)omp_get_max_threads();
#else
    stats->threads = 1;
#endif
  }
  int32_t ret = (out_pl && np > cap_pl) ? GROVE_ERR_LIMIT : GROVE_OK;
  shard_free(h);
  return ret;
}

/* Outputs are in caller node indices.  out_fit/out_score (nullable) receive the round-1 rows of all
 * cliques in SORTED node order (row stride words = ceil(n/32) and n bytes). */
int32_t oracle_run_cycle(const grove_node_t* nodes_in, uint32_t n, uint32_t L,
                         const grove_gang_t* gangs, uint32_t G, const grove_clique_t* cliques, uint32_t Q,
                         const grove_scope_t* scopes, uint32_t S, uint32_t max_rounds, uint32_t alternatives, int32_t threads,
                         grove_placement_t* out_pl, uint32_t cap_pl, uint32_t* n_pl,
                         grove_gang_status_t* out_status, grove_node_t* out_nodes, uint32_t* out_perm,
                         uint32_t* out_fit, uint8_t* out_score, oracle_stats_t* stats) {
  oshard_t* h = NULL;
  int32_t rc = oracle_shard_begin(nodes_in, n, L, gangs, G, cliques, Q, scopes, S, max_rounds, alternatives, threads, 0, 1,
                                  out_fit, out_score, &h);
  if (rc != GROVE_OK) return rc;
  int32_t* xb = malloc(sizeof(int32_t) * (h->X.words ? h->X.words : 1));
  if (!xb) { shard_free(h); return GROVE_ERR_OOM; }
  for (;;) {
    uint32_t go = 0;
    oracle_shard_eval(h, xb, &go);
    if (!go) break;
    oracle_shard_resolve(h, xb, NULL);
  }
  rc = oracle_shard_end(h, out_pl, cap_pl, n_pl, out_status, out_nodes, out_perm, stats);
  free(xb);
  return rc;
}

/* topology preprocessing alone, for tests of the engine's host-side sort: perm + tree-ified dom ids */
int32_t oracle_topology(const grove_node_t* nodes_in, uint32_t n, uint32_t L, uint32_t* out_perm,
                        uint32_t* out_dom /* n * GROVE_MAX_LEVELS, sorted order */, uint32_t* out_n_dom /* L */,
                        uint32_t* out_non_tree) {
  topo_t T;
  if (L < 1 || L > GROVE_MAX_LEVELS || n == 0) return GROVE_ERR_INVALID_ARG;
  if (topo_build(&T, nodes_in, n, L)) return GROVE_ERR_OOM;
  for (uint32_t i = 0; i < n; ++i) {
    if (out_perm) out_perm[i] = T.perm[i];
    if (out_dom) for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) out_dom[i * GROVE_MAX_LEVELS + l] = T.nodes[i].dom[l];
  }
  if (out_n_dom) for (uint32_t l = 0; l < L; ++l) out_n_dom[l] = T.n_dom[l];
  if (out_non_tree) *out_non_tree = T.non_tree;
  topo_free(&T);
  return GROVE_OK;
}

uint32_t oracle_abi_version(void) { return GROVE_ABI_VERSION; }
