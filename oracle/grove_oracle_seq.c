/*
 * grove_oracle_seq.c -- CPU restatement of the gang-placement cycle: the SEQUENTIAL, priority-ordered pass.
 * TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libgrove_place.so) never links, loads or calls it.
 *
 * What it is: gangs are taken ONE AT A TIME in (priority desc, submission index asc) order
 * (PodGangSpec.PriorityClassName, scheduler/api/core/v1alpha1/podgang.go:62-64);
 * each is evaluated against the node state the earlier ones left, packed all-or-nothing, and its
 * resources are subtracted before the next gang is looked at.  There are no rounds, no alternatives and
 * no concurrency in here: this is the definition the CUDA engine (which reaches the same answer by
 * parallel relaxation) is compared with bit for bit.
 *
 * PARITY STATUS: "parity unpinned" at node level.  The reference tree (ai-dynamo/grove @ 08ad3b37)
 * contains no scheduler: placement is done by KAI-Scheduler v0.14.0 (operator/go.mod:11), which is
 * not vendored, and there is no Go toolchain in this image.  What the reference does pin -- and
 * what tests/test_oracle_e2e_properties.py checks this file against -- is
 *   * the input schema             scheduler/api/core/v1alpha1/podgang.go:51-131
 *   * priority order of gangs                                                 podgang.go:62-64
 *   * MinReplicas = gang guarantee, surplus best effort                      podgang.go:80-83
 *   * Required pack constraint = all pods of the scope share one label value podgang.go:101-109
 *   * Preferred = best effort, widening level by level up to Required        podgang.go:110-117
 *   * PlacementScore 1.0 = best possible placement                           podgang.go:187-189
 *   * nodes lacking the label are not candidates      docs/proposals/244-topology-aware-scheduling/README.md:65
 *   * level index 0 is the broadest                                          ibid. :143
 *   * scope nesting gang >= group-config >= pod-group   operator/internal/webhook/admission/pcs/validation/topologyconstraints.go:195-202
 *   * all-or-nothing admission               operator/internal/controller/podclique/components/pod/syncflow.go:319-358
 *   * scaled gangs gated behind their base gang                              ibid. :255-314
 *   * the outcome properties of the live-cluster e2e suites GS1-GS12 / TAS2-TAS17
 *     (operator/e2e/tests/gang_scheduling_test.go, topology_test.go), the step-by-step pod counts of the
 *     multi-step suites GS2-GS12 (tests/test_oracle_e2e_sequences.py), and the reference's own workload
 *     files run end to end (tests/test_workload_fixtures.py).
 * Everything below those (which node, which domain among feasible ones) is defined by DESIGN.md
 * "Placement semantics"; this file is its executable form, written as plain scalar loops that derive
 * every ordering from the score row itself (no piece iterator, no warp tricks, no capacity tables) so
 * that it shares no code or shortcut with the CUDA path.
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
  uint32_t gangs_admitted;
  uint32_t gangs_rejected;   /* REJECTED + BASE_REJECTED */
  uint32_t pods_bound;
  uint32_t threads;          /* OpenMP threads used for the per-gang fit/score rows (the gang loop is sequential) */
  uint64_t pairs_evaluated;  /* (clique,node) pairs through fit+score */
  double seconds_total;
  uint32_t non_tree_labels;  /* raw label ids that appeared under more than one parent */
  uint32_t reserved;
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
  uint32_t* anchor;  /* gang -> sorted node index */
} ctx_t;

typedef struct entry { uint32_t node; uint8_t clique_rel; } entry_t;

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

/* PlacementScore bookkeeping (podgang.go:187-189): a unit that carries a pack
 * constraint asked for `want` = Preferred if set else Required; it was packed at level `got`
 * (-1 = no domain of its own).  Levels count from 1 (level index + 1): credit min(got, want) + 1 of want + 1. */
typedef struct score { uint32_t num, den; } score_t;
static void score_unit(score_t* sc, uint32_t req, uint32_t pref, int got) {
  if (req == GROVE_LEVEL_NONE && pref == GROVE_LEVEL_NONE) return;
  const int want = pref != GROVE_LEVEL_NONE ? (int)pref : (int)req;
  sc->den += (uint32_t)want + 1u;
  sc->num += (uint32_t)((got < want ? got : want) + 1);
}

typedef struct geval {
  const ctx_t* C;
  uint32_t g;
  uint32_t a;
  uint8_t* Trow[GROVE_MAX_GANG_CLIQUES]; /* score rows of this gang's cliques against the current state */
  uint32_t ndepth[GROVE_MAX_GANG_CLIQUES];
  entry_t st[GROVE_MAX_GANG_PODS];
  uint32_t np;
  uint32_t Hlo[GROVE_MAX_GANG_CLIQUES], Hhi[GROVE_MAX_GANG_CLIQUES];
  int c_got[GROVE_MAX_GANG_CLIQUES];   /* level each clique / scope was packed at (-1: the parent range) */
  int s_got[GROVE_MAX_GANG_SCOPES];
  uint32_t s_lo[GROVE_MAX_GANG_SCOPES];
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
        E->st[E->np].node = n; E->st[E->np].clique_rel = (uint8_t)cr;
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

/* cliques of one scope inside range [lo,hi) whose level is `lvl` (-1 = ROOT) */
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
      if (ok) E->c_got[cr] = l;
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
        for (uint32_t k = 0; k < nc && !ok; ++k) { ok = place_scope(E, s, v[k].lo, v[k].hi, l); if (ok) E->s_lo[si] = v[k].lo; }
        free(v);
      } else {
        ok = place_scope(E, s, lo, hi, lvl);
        if (ok) E->s_lo[si] = lo;
      }
      if (ok) E->s_got[si] = l;
    }
    if (!ok) { E->np = 0; return 0; }
  }
  return 1;
}

/* One gang against the current node state: its first feasible gang-level domain in score order (the whole
 * cluster when it has no gang-level constraint).  With a Preferred level the candidate list is the Preferred
 * level's domains in score order, then each wider level's, down to the Required level (or the whole
 * cluster).  On success E->st[0..np) holds the pods (MinReplicas of every clique in scope/clique order, then
 * the best-effort surplus, podgang.go:80-83) and *n_min the MinReplicas part.  Returns 1 if feasible. */
static int eval_gang(geval_t* E, const ctx_t* C, uint32_t gi, uint8_t* const* Trow, int* g_got, uint32_t* g_lo, uint32_t* n_min) {
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
  int ok = 0, first, base = level_span(g->level, g->preferred, -1, &first);
  for (int l = first; l >= base && !ok; --l) {
    if (l < 0) {
      ok = place_in(E, 0, C->T.n, -1);
      if (ok) { *g_got = -1; *g_lo = 0; }
    } else {
      uint32_t nc; cand_t* v = subdomains(E, (uint32_t)l, 0, C->T.n, &nc);
      for (uint32_t k = 0; k < nc && !ok; ++k) { ok = place_in(E, v[k].lo, v[k].hi, l); if (ok) { *g_got = l; *g_lo = v[k].lo; } }
      free(v);
    }
  }
  if (!ok) return 0;
  *n_min = E->np;
  for (uint32_t cr = 0; cr < g->n_cliques; ++cr) {
    const grove_clique_t* q = &C->cliques[g->clique_off + cr];
    uint32_t extra = q->replicas > q->min_replicas ? (uint32_t)(q->replicas - q->min_replicas) : 0;
    if (extra) take(E, cr, E->Hlo[cr], E->Hhi[cr], extra);
  }
  return 1;
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
 * One scheduling cycle, sequentially (DESIGN.md section 1 "Cycle").
 *   for every gang in (priority desc, index asc) order:
 *     gated                                  -> GATED_SKIP
 *     base gang not ADMITTED so far          -> BASE_REJECTED   (pod/syncflow.go:319-358: a scaled gang's pods stay
 *                                                                gated until the base gang is scheduled; it waits
 *                                                                for a later cycle)
 *     fit row + score row of each clique over all nodes (K1, K2 semantics) against the CURRENT state,
 *     first feasible domain in score order   -> ADMITTED, resources subtracted
 *     none                                   -> REJECTED, nothing bound
 * Outputs are in caller node indices.  out_fit/out_score (nullable) receive the rows of all cliques against
 * the cycle-START state in SORTED node order (row stride words = ceil(n/32) and n bytes): what K1 and K2 of
 * the engine produce.  out_scopes (nullable): S records.  `threads` parallelises the per-gang row computation
 * over nodes; the gang loop is sequential by definition.
 */
int32_t oracle_run_cycle(const grove_node_t* nodes_in, uint32_t n, uint32_t L,
                         const grove_gang_t* gangs, uint32_t G, const grove_clique_t* cliques, uint32_t Q,
                         const grove_scope_t* scopes, uint32_t S, int32_t threads,
                         grove_placement_t* out_pl, uint32_t cap_pl, uint32_t* n_pl,
                         grove_gang_status_t* out_status, grove_scope_status_t* out_scopes, grove_node_t* out_nodes, uint32_t* out_perm,
                         uint32_t* out_fit, uint8_t* out_score, oracle_stats_t* stats) {
  int32_t rc = oracle_validate(gangs, G, cliques, Q, scopes, S, L, n);
  if (rc != GROVE_OK) return rc;
  if (n == 0 || n > GROVE_MAX_NODES) return GROVE_ERR_INVALID_ARG;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  const double t0 = now_s();
  ctx_t Cx; memset(&Cx, 0, sizeof(Cx));
  ctx_t* C = &Cx;
  if (topo_build(&C->T, nodes_in, n, L)) { topo_free(&C->T); return GROVE_ERR_OOM; }
  C->G = G; C->Q = Q; C->S = S; C->gangs = gangs; C->cliques = cliques; C->scopes = scopes;
  C->anchor = malloc(sizeof(uint32_t) * (G ? G : 1));
  ord_t* ov = malloc(sizeof(ord_t) * (G ? G : 1));
  for (uint32_t g = 0; g < G; ++g) {
    ov[g].pr = gangs[g].priority; ov[g].g = g;
    C->anchor[g] = gangs[g].anchor_node != GROVE_NONE_U32 ? C->T.inv[gangs[g].anchor_node] : fmix32(g) % n;
  }
  qsort(ov, G, sizeof(ord_t), cmp_ord);
  const uint32_t words = (n + 31) / 32;
  uint64_t pairs = 0;
  /* K1 / K2 over the cycle-start snapshot, every clique (debug / parity rows) */
  if (out_fit || out_score) {
    for (uint32_t gi = 0; gi < G; ++gi) {
      const grove_gang_t* g = &gangs[gi];
      for (uint32_t si = 0; si < g->n_scopes; ++si) {
        const grove_scope_t* s = &scopes[g->scope_off + si];
        for (uint32_t i = 0; i < s->n_cliques; ++i) {
          const uint32_t qi = g->clique_off + s->first_clique + i;
          const grove_clique_t* q = &cliques[qi];
          const uint32_t ndp = need_depth(g, s, q);
          if (out_fit) memset(out_fit + (size_t)qi * words, 0, sizeof(uint32_t) * words);
          for (uint32_t nn = 0; nn < n; ++nn) {
            const int f = fit(&C->T.nodes[nn], q, ndp);
            if (f && out_fit) out_fit[(size_t)qi * words + (nn >> 5)] |= 1u << (nn & 31);
            if (out_score) out_score[(size_t)qi * n + nn] = f ? (uint8_t)(closeness(&C->T, nn, C->anchor[gi]) + 1) : 0;
          }
        }
      }
    }
  }
  uint8_t* state = calloc(G ? G : 1, 1);
  /* placements are kept per gang and emitted in submission order at the end */
  uint32_t* pod_off = malloc(sizeof(uint32_t) * (G + 1));
  uint32_t po = 0;
  for (uint32_t g = 0; g < G; ++g) { pod_off[g] = po; for (uint32_t c = 0; c < gangs[g].n_cliques; ++c) po += cliques[gangs[g].clique_off + c].replicas; }
  pod_off[G] = po;
  entry_t* fin = malloc(sizeof(entry_t) * (po ? po : 1));
  uint32_t* fin_n = calloc(G ? G : 1, sizeof(uint32_t));
  grove_gang_status_t* gst = calloc(G ? G : 1, sizeof(grove_gang_status_t));
  grove_scope_status_t* sst = malloc(sizeof(grove_scope_status_t) * (S ? S : 1));
  for (uint32_t i = 0; i < S; ++i) { memset(&sst[i], 0, sizeof(sst[i])); sst[i].level = GROVE_LEVEL_NONE; sst[i].domain_node = GROVE_NONE_U32; }
  geval_t* E = malloc(sizeof(geval_t));
  uint8_t* Trow[GROVE_MAX_GANG_CLIQUES];
  for (uint32_t c = 0; c < GROVE_MAX_GANG_CLIQUES; ++c) Trow[c] = malloc(n);

  for (uint32_t r = 0; r < G; ++r) {
    const uint32_t gi = ov[r].g;
    const grove_gang_t* g = &gangs[gi];
    gst[gi].level = GROVE_LEVEL_NONE; gst[gi].domain_node = GROVE_NONE_U32;
    if (g->flags & GROVE_GANG_GATED) { state[gi] = GROVE_GANG_GATED_SKIP; continue; }
    if (g->base_gang != GROVE_NONE_U32 && state[g->base_gang] != GROVE_GANG_ADMITTED) { state[gi] = GROVE_GANG_BASE_REJECTED; continue; }
    for (uint32_t si = 0; si < g->n_scopes; ++si) {
      const grove_scope_t* s = &scopes[g->scope_off + si];
      for (uint32_t i = 0; i < s->n_cliques; ++i) {
        const uint32_t cr = s->first_clique + i;
        const grove_clique_t* q = &cliques[g->clique_off + cr];
        const uint32_t ndp = need_depth(g, s, q);
        uint8_t* row = Trow[cr];
        const uint32_t a = C->anchor[gi];
        /* K1: fit; K2: score = fit ? closeness + 1 : 0 (one pass over the node table) */
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (uint32_t nn = 0; nn < n; ++nn)
          row[nn] = fit(&C->T.nodes[nn], q, ndp) ? (uint8_t)(closeness(&C->T, nn, a) + 1) : 0;
        pairs += n;
      }
    }
    int g_got = -1; uint32_t g_lo = 0, n_min = 0;
    if (!eval_gang(E, C, gi, Trow, &g_got, &g_lo, &n_min)) { state[gi] = GROVE_GANG_REJECTED; continue; }
    state[gi] = GROVE_GANG_ADMITTED;
    /* commit: the next gang sees what is left */
    for (uint32_t i = 0; i < E->np; ++i) {
      const grove_clique_t* q = &cliques[g->clique_off + E->st[i].clique_rel];
      grove_node_t* node = &C->T.nodes[E->st[i].node];
      node->free_cpu_milli -= q->req_cpu_milli; node->free_mem_mib -= q->req_mem_mib;
      node->free_gpu -= q->req_gpu; node->free_pods -= 1;
      fin[pod_off[gi] + i] = E->st[i];
    }
    fin_n[gi] = E->np;
    /* PlacementScore and the chosen domains */
    score_t sc = {0, 0};
    score_unit(&sc, g->level, g->preferred, g_got);
    for (uint32_t si = 0; si < g->n_scopes; ++si) {
      const grove_scope_t* s = &scopes[g->scope_off + si];
      score_unit(&sc, s->level, scope_pref(s), E->s_got[si]);
      /* a scope packed at the gang's own level has no domain of its own */
      if (E->s_got[si] > g_got) { sst[g->scope_off + si].level = (uint8_t)E->s_got[si]; sst[g->scope_off + si].domain_node = C->T.perm[E->s_lo[si]]; }
      for (uint32_t i = 0; i < s->n_cliques; ++i) {
        const uint32_t cr = s->first_clique + i;
        const grove_clique_t* q = &cliques[g->clique_off + cr];
        score_unit(&sc, q->level, GROVE_CLIQUE_PREFERRED(q->scope), E->c_got[cr]);
      }
    }
    if (sc.den == 0) { sc.num = 1; sc.den = 1; }
    gst[gi].score_num = (uint16_t)sc.num; gst[gi].score_den = (uint16_t)sc.den;
    if (g_got >= 0) { gst[gi].level = (uint8_t)g_got; gst[gi].domain_node = C->T.perm[g_lo]; }
  }

  uint32_t np = 0, adm = 0, rej = 0;
  for (uint32_t g = 0; g < G; ++g) {
    gst[g].state = state[g]; gst[g].placement_off = np; gst[g].n_pods = 0;
    if (state[g] == GROVE_GANG_ADMITTED) {
      adm++;
      gst[g].n_pods = fin_n[g];
      for (uint32_t i = 0; i < fin_n[g]; ++i) {
        if (out_pl && np < cap_pl) {
          out_pl[np].clique = gangs[g].clique_off + fin[pod_off[g] + i].clique_rel;
          out_pl[np].node = C->T.perm[fin[pod_off[g] + i].node];
        }
        np++;
      }
    } else if (state[g] == GROVE_GANG_REJECTED || state[g] == GROVE_GANG_BASE_REJECTED) rej++;
    if (out_status) out_status[g] = gst[g];
  }
  if (n_pl) *n_pl = np;
  if (out_scopes) memcpy(out_scopes, sst, sizeof(grove_scope_status_t) * S);
  if (out_nodes)
    for (uint32_t i = 0; i < n; ++i) {
      grove_node_t nd = nodes_in[C->T.perm[i]];
      nd.free_cpu_milli = C->T.nodes[i].free_cpu_milli; nd.free_mem_mib = C->T.nodes[i].free_mem_mib;
      nd.free_gpu = C->T.nodes[i].free_gpu; nd.free_pods = C->T.nodes[i].free_pods;
      out_nodes[C->T.perm[i]] = nd;
    }
  if (out_perm) memcpy(out_perm, C->T.perm, sizeof(uint32_t) * n);
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->gangs_admitted = adm; stats->gangs_rejected = rej; stats->pods_bound = np;
    stats->pairs_evaluated = pairs; stats->seconds_total = now_s() - t0;
    stats->non_tree_labels = C->T.non_tree;
#ifdef _OPENMP
    stats->threads = (uint32_t)omp_get_max_threads();
#else
    stats->threads = 1;
#endif
  }
  for (uint32_t c = 0; c < GROVE_MAX_GANG_CLIQUES; ++c) free(Trow[c]);
  free(E); free(state); free(pod_off); free(fin); free(fin_n); free(gst); free(sst); free(ov); free(C->anchor);
  topo_free(&C->T);
  return (out_pl && np > cap_pl) ? GROVE_ERR_LIMIT : GROVE_OK;
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
