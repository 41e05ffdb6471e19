/*
 * grove_place.h -- C ABI of libgrove_place.so, the B200 gang-placement engine.
 *
 * This is the drop-in boundary for the scheduling-cycle hot path named in
 * BASELINE.json (resource fit -> topology score -> all-or-nothing gang
 * admission).  The reference (ai-dynamo/grove @ 08ad3b37) ships no scheduler
 * and no FFI; what it pins is the INPUT contract (the PodGang CRD) and the
 * plug-in seam a scheduler hangs off.  Every entry point below names the
 * reference interface it stands in for:
 *
 *   grove_engine_create / _destroy   scheduler.Backend.Init()
 *                                    operator/internal/scheduler/types.go:41-43
 *   grove_load_nodes / _update_nodes node snapshot a scheduler session takes of
 *                                    Node.status.allocatable + topology labels;
 *                                    generator: operator/hack/infra_manager/kwok.py:55-117
 *                                    level schema: operator/internal/scheduler/kai/topology.go:103-135
 *   grove_submit_gangs               scheduler.Backend.SyncPodGang(ctx, *PodGang)
 *                                    operator/internal/scheduler/types.go:45-47
 *                                    payload = PodGangSpec, scheduler/api/core/v1alpha1/podgang.go:51-131
 *   grove_run_cycle                  the scheduling cycle itself (absent from the
 *                                    reference tree; KAI-Scheduler v0.14.0, operator/go.mod:11)
 *   grove_get_placements             Pod.spec.nodeName bindings, counted back by
 *                                    operator/internal/controller/podclique/reconcilestatus.go:134-141
 *   grove_get_gang_status            PodGangStatus{Phase,PlacementScore},
 *                                    scheduler/api/core/v1alpha1/podgang.go:141-150,182-190
 *
 * cgo rules honoured: plain pointers + sizes only, every input array is copied
 * before the call returns (no Go pointer is retained), outputs go to
 * caller-owned buffers, there are no callbacks.  A handle is NOT thread-safe:
 * one cycle in flight per handle; distinct handles are independent.
 *
 * There is no CPU fallback behind this ABI.  If no CUDA device is usable
 * grove_engine_create fails with GROVE_ERR_NO_DEVICE.
 */
#ifndef GROVE_PLACE_H
#define GROVE_PLACE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GROVE_ABI_VERSION 1u

#define GROVE_MAX_LEVELS 4u          /* topology levels, index 0 = broadest (GREP-244 README.md:143) */
#define GROVE_LEVEL_NONE 0xFFu       /* "no pack constraint at this scope" */
#define GROVE_DOM_ABSENT 0xFFFFFFFFu /* node lacks the label of that level (GREP-244 README.md:65) */
#define GROVE_NONE_U32 0xFFFFFFFFu
#define GROVE_MAX_GANG_PODS 128u     /* sum of replicas over a gang's cliques */
#define GROVE_MAX_GANG_CLIQUES 32u
#define GROVE_MAX_GANG_SCOPES 32u
#define GROVE_MAX_NODES (1u << 24)
#ifndef GROVE_MAX_ALTERNATIVES
#define GROVE_MAX_ALTERNATIVES 8u    /* feasible placements kept per gang per round */
#endif
#ifndef GROVE_SUBROUNDS
#define GROVE_SUBROUNDS 8u           /* conflict-resolution passes over the alternatives per round */
#endif

/* error codes (all entry points return 0 on success, <0 on error) */
#define GROVE_OK 0
#define GROVE_ERR_INVALID_ARG (-1)
#define GROVE_ERR_NO_DEVICE (-2)
#define GROVE_ERR_CUDA (-3)
#define GROVE_ERR_LIMIT (-4)
#define GROVE_ERR_STATE (-5)
#define GROVE_ERR_OOM (-6)

/* ---- node-state table: 32 B per node -------------------------------------------------------- */
#define GROVE_NODE_SCHEDULABLE 0x1u               /* not cordoned / Ready */
#define GROVE_NODE_CLASS_SHIFT 8u                 /* bits 8..11: selector class id 0..15 */
#define GROVE_NODE_CLASS_MASK 0xF00u
typedef struct grove_node {
  uint32_t free_cpu_milli; /* allocatable - requested, millicores */
  uint32_t free_mem_mib;   /* MiB */
  uint16_t free_gpu;       /* nvidia.com/gpu */
  uint16_t free_pods;      /* pod slots left */
  uint32_t flags;          /* GROVE_NODE_SCHEDULABLE | class << 8 */
  uint32_t dom[GROVE_MAX_LEVELS]; /* interned label value per level, or GROVE_DOM_ABSENT */
} grove_node_t;

/* ---- gang-request tables -------------------------------------------------------------------- */
/* one PodGroup (PodClique) of a PodGang: podgang.go:75-91 */
typedef struct grove_clique {
  uint32_t req_cpu_milli; /* per-pod request */
  uint32_t req_mem_mib;
  uint16_t req_gpu;
  uint8_t min_replicas;   /* PodGroup.MinReplicas: gang-guaranteed */
  uint8_t replicas;       /* len(PodReferences) >= min_replicas: surplus is best effort (podgang.go:80-83) */
  uint16_t class_mask;    /* bit c set <=> node selector class c is acceptable (nodeSelector+tolerations) */
  uint8_t level;          /* PodGroup.TopologyConstraint Required level index, or GROVE_LEVEL_NONE */
  uint8_t scope;          /* bits 0..4: index of the owning scope inside the gang (non-decreasing over the gang's
                             cliques); bits 5..7: Preferred level index + 1, 0 = none (GROVE_CLIQUE_SCOPE_PREF) */
} grove_clique_t;
#define GROVE_CLIQUE_SCOPE(x) ((uint32_t)(x) & 0x1Fu)
#define GROVE_CLIQUE_PREFERRED(x) ((((uint32_t)(x)) >> 5) ? (((uint32_t)(x)) >> 5) - 1u : (uint32_t)GROVE_LEVEL_NONE)
#define GROVE_CLIQUE_SCOPE_PREF(scope, pref) \
  ((uint8_t)(((scope) & 0x1Fu) | (((pref) == GROVE_LEVEL_NONE ? 0u : (uint32_t)(pref) + 1u) << 5)))

/* one TopologyConstraintGroupConfig (podgang.go:120-131), or the implicit scope of loose PodGroups */
typedef struct grove_scope {
  uint16_t first_clique;  /* relative to the gang's clique_off */
  uint16_t n_cliques;
  uint8_t level;          /* Required level index, or GROVE_LEVEL_NONE */
  uint8_t preferred1;     /* Preferred level index + 1 (deeper than `level`), 0 = none: zero-filled records
                             written against the earlier layout keep their meaning */
  uint8_t reserved[2];
} grove_scope_t;

#define GROVE_GANG_GATED 0x1u /* pods still carry the grove.io/podgang-pending-creation gate: skip */
typedef struct grove_gang {
  uint32_t clique_off;    /* into the cliques array of the same submission */
  uint32_t scope_off;     /* into the scopes array */
  uint16_t n_cliques;
  uint16_t n_scopes;
  int32_t priority;       /* PriorityClass value; higher goes first */
  uint32_t anchor_node;   /* caller's node index to score distance against (ReuseReservationRef hint,
                             podgang.go:66-71), or GROVE_NONE_U32 = engine derives hash(gang index) % N */
  uint32_t base_gang;     /* scaled gang: index of its base gang in this submission (gated behind it,
                             pod/syncflow.go:319-358), or GROVE_NONE_U32 */
  uint8_t level;          /* PodGangSpec.TopologyConstraint Required level index, or GROVE_LEVEL_NONE */
  uint8_t preferred;      /* PackConstraint.Preferred level index (podgang.go:110-117): best effort, tried before
                             falling back level by level up to `level`; deeper than `level`; or GROVE_LEVEL_NONE */
  uint16_t flags;
  uint32_t reserved;
} grove_gang_t;

/* ---- outputs -------------------------------------------------------------------------------- */
/* one bound pod: the r-th entry of a clique binds that clique's r-th PodReference */
typedef struct grove_placement {
  uint32_t clique; /* global clique index in the submission */
  uint32_t node;   /* caller's node index (position in the grove_load_nodes array) */
} grove_placement_t;

#define GROVE_GANG_PENDING 0u       /* never seen by a cycle */
#define GROVE_GANG_ADMITTED 1u      /* all MinReplicas bound (Scheduled=True) */
#define GROVE_GANG_REJECTED 2u      /* no feasible domain: Unschedulable, nothing bound */
#define GROVE_GANG_GATED_SKIP 3u    /* GROVE_GANG_GATED set */
#define GROVE_GANG_BASE_REJECTED 4u /* base gang was not admitted */
typedef struct grove_gang_status {
  uint8_t state;
  uint8_t score_num;       /* PlacementScore = score_num / score_den, (0,1]; 0/0 when not admitted */
  uint8_t score_den;
  uint8_t round;           /* optimistic round in which the gang was resolved */
  uint32_t n_pods;         /* pods bound (>= sum MinReplicas when admitted) */
  uint32_t placement_off;  /* first entry in grove_get_placements output */
  uint32_t top_domain_lo;  /* caller-order is not contiguous: first SORTED node index of chosen gang domain, or NONE */
} grove_gang_status_t;

typedef struct grove_config {
  uint32_t abi_version;  /* GROVE_ABI_VERSION */
  int32_t device;        /* CUDA ordinal */
  uint32_t n_levels;     /* 1..GROVE_MAX_LEVELS */
  uint32_t max_rounds;   /* 0 = default (unbounded until every gang is resolved) */
  uint32_t rank;         /* gang-row sharding: this handle evaluates gangs g with g % world == rank */
  uint32_t world;        /* 0 or 1 = unsharded */
  uint32_t alternatives; /* 1..GROVE_MAX_ALTERNATIVES, 0 = default (GROVE_MAX_ALTERNATIVES) */
  uint32_t reserved;
} grove_config_t;

typedef struct grove_cycle_stats {
  uint32_t rounds;
  uint32_t gangs_admitted;
  uint32_t gangs_rejected;
  uint32_t pods_bound;
  uint64_t pairs_evaluated;  /* (clique,node) pairs through fit+score, all rounds */
  uint64_t kernel_launches;
  float ms_fit;              /* CUDA-event device time per kernel family, summed over rounds */
  float ms_score;
  float ms_admit;
  float ms_commit;
  float ms_total;            /* first launch -> last kernel done */
  float reserved;
} grove_cycle_stats_t;

typedef struct grove_engine grove_engine_t;

int32_t grove_engine_create(const grove_config_t* cfg, grove_engine_t** out);
void grove_engine_destroy(grove_engine_t* e);
const char* grove_last_error(grove_engine_t* e);
uint32_t grove_abi_version(void);

/* full snapshot; copied (host -> device) before return */
int32_t grove_load_nodes(grove_engine_t* e, const grove_node_t* nodes, uint32_t n);
/* churn deltas: replace free_* and flags of nodes idx[i] (labels must not change) */
int32_t grove_update_nodes(grove_engine_t* e, const uint32_t* idx, const grove_node_t* recs, uint32_t n);
/* read back the engine's current node table in caller order (after commits) */
int32_t grove_get_nodes(grove_engine_t* e, grove_node_t* out, uint32_t cap);

int32_t grove_submit_gangs(grove_engine_t* e, const grove_gang_t* gangs, uint32_t n_gangs,
                           const grove_clique_t* cliques, uint32_t n_cliques,
                           const grove_scope_t* scopes, uint32_t n_scopes);

/* blocking: fit -> score -> admit -> commit, optimistic rounds until every gang is resolved */
int32_t grove_run_cycle(grove_engine_t* e, grove_cycle_stats_t* stats);

int32_t grove_get_placements(grove_engine_t* e, grove_placement_t* out, uint32_t cap, uint32_t* n_out);
int32_t grove_get_gang_status(grove_engine_t* e, grove_gang_status_t* out, uint32_t cap);

/* ---- device-resident variants (inputs already in HBM; used by bench.py `value`) -------------- */
/* d_nodes: device pointer to n grove_node_t in caller order, labels identical to the last load */
int32_t grove_load_nodes_device(grove_engine_t* e, const void* d_nodes, uint32_t n);

/* ---- multi-GPU stepping: one handle per rank (cfg.rank / cfg.world) -------------------------------
 * Gang rows are dealt g % world to ranks; the node table and the gang states are replicated.  A round
 * has two halves: each rank EVALUATES its own gangs (fit -> score -> up to `alternatives` feasible
 * placements per gang) into an exchange buffer that is zero for gangs it does not own; after one
 * all-reduce SUM over that buffer every rank RESOLVES the conflicts and commits identically.  Results
 * are bit-identical for every world size.  The buffer is a DEVICE pointer owned by the engine, int32
 * words; the host (torch.distributed / NCCL) reduces it in place:
 *
 *   grove_cycle_begin
 *   loop: grove_round_eval(&buf,&n,&go); if (!go) break;     all-reduce SUM over buf[n]
 *         grove_round_resolve(&remaining);
 *   grove_cycle_end(&stats)
 *
 * Every call is synchronous (its kernels are complete on return).  With world <= 1 the same sequence
 * is valid without the reduction, and grove_run_cycle is the fused form of it. */
int32_t grove_cycle_begin(grove_engine_t* e);
int32_t grove_round_eval(grove_engine_t* e, void** d_words, uint32_t* n_words, uint32_t* go);
int32_t grove_round_resolve(grove_engine_t* e, uint32_t* remaining);
int32_t grove_cycle_end(grove_engine_t* e, grove_cycle_stats_t* stats);
/* The CUDA stream (cudaStream_t) every engine kernel runs on.  A host that enqueues its reduction ON THIS
 * STREAM (ncclAllReduce(..., stream), or torch.cuda.ExternalStream) may switch the stepping calls to
 * stream-ordered completion with grove_set_stream_ordered(e, 1): they then return without waiting for their
 * kernels, and eval -> all-reduce -> resolve are ordered by the stream alone (no host synchronisation). */
int32_t grove_engine_stream(grove_engine_t* e, void** stream);
int32_t grove_set_stream_ordered(grove_engine_t* e, int32_t on);

/* ---- introspection for the parity tests (sorted node order; see DESIGN.md "Data layout") ------ */
int32_t grove_debug_get_perm(grove_engine_t* e, uint32_t* sorted_to_caller, uint32_t cap);
/* matrices as the last round left them (create the engine with max_rounds = 1 to read round 1):
 * fit bitmap row (ceil(n/32) words) and score row (n bytes) of one clique */
int32_t grove_debug_get_fit_row(grove_engine_t* e, uint32_t clique, uint32_t* words, uint32_t cap_words);
int32_t grove_debug_get_score_row(grove_engine_t* e, uint32_t clique, uint8_t* bytes, uint32_t cap_bytes);
#ifdef __cplusplus
}
#endif
#endif /* GROVE_PLACE_H */
