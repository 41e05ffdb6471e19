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
 *                                    reference tree; KAI-Scheduler v0.14.0, operator/go.mod:11).
 *                                    Result = the SEQUENTIAL pass: gangs one at a time in (priority desc,
 *                                    submission index asc) order, each against the state the earlier ones left
 *                                    (PriorityClassName, podgang.go:62-64); the engine reaches it by parallel
 *                                    relaxation, bit-identical to oracle/grove_oracle_seq.c
 *   grove_get_placements             Pod.spec.nodeName bindings, counted back by
 *                                    operator/internal/controller/podclique/reconcilestatus.go:134-141
 *   grove_get_gang_status            PodGangStatus{Phase,PlacementScore},
 *                                    scheduler/api/core/v1alpha1/podgang.go:141-150,182-190
 *   grove_get_scope_domains          the topology domain chosen for every TopologyConstraintGroupConfig
 *                                    (podgang.go:120-131) -- what a binder needs to explain a placement
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

#define GROVE_ABI_VERSION 2u

#define GROVE_MAX_LEVELS 4u          /* topology levels, index 0 = broadest (GREP-244 README.md:143) */
#define GROVE_LEVEL_NONE 0xFFu       /* "no pack constraint at this scope" */
#define GROVE_DOM_ABSENT 0xFFFFFFFFu /* node lacks the label of that level (GREP-244 README.md:65) */
#define GROVE_NONE_U32 0xFFFFFFFFu
#define GROVE_MAX_GANG_PODS 128u     /* sum of replicas over a gang's cliques */
#define GROVE_MAX_GANG_CLIQUES 32u
#define GROVE_MAX_GANG_SCOPES 32u
#define GROVE_MAX_NODES (1u << 24)

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
  int32_t priority;       /* PriorityClass value (PodGangSpec.PriorityClassName resolved, podgang.go:62-64):
                             gangs are placed strictly in (priority desc, submission index asc) order */
  uint32_t anchor_node;   /* caller's node index to score distance against (ReuseReservationRef hint,
                             podgang.go:66-71), or GROVE_NONE_U32 = engine derives hash(gang index) % N */
  uint32_t base_gang;     /* scaled gang: index of its base gang in this submission (gated behind it,
                             pod/syncflow.go:319-358), or GROVE_NONE_U32.  Considered at its own turn: admitted
                             only if the base gang was admitted EARLIER in this cycle's order, else
                             GROVE_GANG_BASE_REJECTED (it waits for the next cycle, like its gated pods do) */
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
#define GROVE_GANG_REJECTED 2u      /* no feasible domain at its turn: Unschedulable, nothing bound */
#define GROVE_GANG_GATED_SKIP 3u    /* GROVE_GANG_GATED set */
#define GROVE_GANG_BASE_REJECTED 4u /* base gang not admitted before this gang's turn */
typedef struct grove_gang_status {
  uint8_t state;
  uint8_t level;           /* level of the domain the whole gang was packed into (its Required level, or the deeper
                              Preferred level that held), GROVE_LEVEL_NONE = the whole cluster / not admitted */
  uint16_t reserved0;      /* flags: GROVE_STATUS_PREEMPTOR = admitted by the reclaim pass of grove_run_cycle_preempt; else 0 */
  uint16_t score_num;      /* PlacementScore = score_num / score_den (podgang.go:187-189): over the gang, its scopes and
                              its cliques that carry a pack constraint, (levels honoured) / (levels asked for), where
                              asked = Preferred if set else Required; 1/1 when nothing was asked or everything held;
                              0/0 when not admitted */
  uint16_t score_den;
  uint32_t n_pods;         /* pods bound (>= sum MinReplicas when admitted) */
  uint32_t placement_off;  /* first entry in grove_get_placements output */
  uint32_t domain_node;    /* caller index of the first node (topology order) of that gang domain, or GROVE_NONE_U32 */
  uint32_t reserved1[3];
} grove_gang_status_t;

/* chosen domain of one scope (same indexing as the scopes array of the submission) */
typedef struct grove_scope_status {
  uint8_t level;           /* level of the domain the scope's cliques share; GROVE_LEVEL_NONE = no own domain (packed in
                              the gang's domain) or gang not admitted */
  uint8_t reserved[3];
  uint32_t domain_node;    /* caller index of the first node (topology order) of that domain, or GROVE_NONE_U32 */
} grove_scope_status_t;

typedef struct grove_config {
  uint32_t abi_version;  /* GROVE_ABI_VERSION */
  int32_t device;        /* CUDA ordinal */
  uint32_t n_levels;     /* 1..GROVE_MAX_LEVELS */
  uint32_t window;       /* gangs beyond the settled prefix that relax concurrently; 0 = default.  A tuning knob:
                            results do not depend on it */
  uint32_t rank;         /* multi-GPU: this handle's rank ... */
  uint32_t world;        /* ... of `world` handles driving the same cycle; 0 or 1 = single GPU */
  uint32_t reserved[2];
} grove_config_t;

typedef struct grove_cycle_stats {
  uint32_t rounds;           /* relaxation rounds the engine took (an implementation figure, not a result) */
  uint32_t gangs_admitted;
  uint32_t gangs_rejected;   /* REJECTED + BASE_REJECTED */
  uint32_t pods_bound;
  uint64_t pairs_evaluated;  /* (clique,node) pairs through fit+score (K1/K2 over the cycle-start snapshot) */
  uint64_t kernel_launches;
  uint64_t evaluations;      /* gang evaluations (K3) over all rounds; >= gangs considered */
  float ms_fit;              /* CUDA-event device time: K1 + capacity tables */
  float ms_score;            /* K2 (beside the relaxation, on its own stream) */
  float ms_admit;            /* relaxation rounds: evaluate / publish / detect / settle */
  float ms_commit;           /* output compaction */
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

/* blocking: fit -> score -> admit -> commit; on return every gang is decided */
int32_t grove_run_cycle(grove_engine_t* e, grove_cycle_stats_t* stats);

int32_t grove_get_placements(grove_engine_t* e, grove_placement_t* out, uint32_t cap, uint32_t* n_out);
int32_t grove_get_gang_status(grove_engine_t* e, grove_gang_status_t* out, uint32_t cap);
int32_t grove_get_scope_domains(grove_engine_t* e, grove_scope_status_t* out, uint32_t cap);

/* ---- preemption / reclaim ------------------------------------------------------------------------
 * The API reserves it (PodGangConditionTypeDisruptionTarget, podgang.go:166-170: "PodGang is preempted by a higher priority
 * PodGang"; PriorityClassName podgang.go:62-64); the arithmetic is ours ("Reclaim pass").  RUNNING gangs
 * (admitted by earlier cycles) are handed in with what they hold per node.  After the ordinary pass every REJECTED gang, in
 * order rank, is evaluated once more against the RECLAIM VIEW of its priority p -- free resources plus everything held by
 * running gangs of priority < p that are still standing.  If it fits, it is admitted (GROVE_STATUS_PREEMPTOR) and, node by
 * node in the order of its placement, running gangs are evicted WHOLE (lowest priority first, then highest running index)
 * until the node's free resources cover what it puts there; everything an evicted gang held anywhere returns to the free
 * pool.  grove_get_victims lists (running gang, preemptor) pairs: the caller sets DisruptionTarget on them. */
typedef struct grove_holding {
  uint32_t node;          /* caller's node index */
  uint32_t cpu_milli, mem_mib;
  uint16_t gpu, pods;
} grove_holding_t;
typedef struct grove_running_gang {
  int32_t priority;
  uint32_t holding_off;   /* into the holdings array */
  uint32_t n_holdings;
  uint32_t reserved;
} grove_running_gang_t;
typedef struct grove_victim {
  uint32_t running;       /* index into the running array */
  uint32_t preemptor;     /* gang index of the submission */
} grove_victim_t;
#define GROVE_STATUS_PREEMPTOR 0x1u /* grove_gang_status_t.reserved0: admitted by the reclaim pass */
/* ordinary pass + reclaim pass; outputs through the usual getters (statuses, placements, scope domains, node table = free
 * resources after evictions and admissions) */
int32_t grove_run_cycle_preempt(grove_engine_t* e, const grove_running_gang_t* running, uint32_t n_running,
                                const grove_holding_t* holdings, uint32_t n_holdings, grove_cycle_stats_t* stats);
int32_t grove_get_victims(grove_engine_t* e, grove_victim_t* out, uint32_t cap, uint32_t* n_out);

/* K2: the topology-distance score matrix T[clique][node] = fit ? 1 + levels shared with the gang's anchor : 0 (u8) over the
 * snapshot the last cycle started from -- what a scheduler's Score extension point would hand out.  The admission does
 * not read it (its visiting order comes from the anchor's domain ranges), so it is built on request; GROVE_TUNE_SCORE=1
 * builds it every cycle beside the admission.  *ms (nullable): device time of the kernel. */
int32_t grove_build_score_matrix(grove_engine_t* e, float* ms);

/* ---- multi-GPU score pass (north_star: node-range shards + one all-reduce of per-shard feasibility) -------------
 * A handle created with grove_config_t.world = W > 1 and rank = r builds K1 + K2 (fit / score matrix) only for ITS node range:
 * the topology-sorted node table cut at top-level domain boundaries into W ranges of about equal size (every handle loads the
 * whole -- small -- node table; the Q x N score matrix is what outgrows a GPU).  grove_run_score_pass: K1 + K2 over the loaded
 * snapshot for the submitted gangs, no admission; *ms (nullable) = device time.  grove_shard_summary_device writes int32[G + Q]
 * into the caller's DEVICE buffer: [g] = domains of gang g's Required level starting in the shard in which each of its cliques
 * finds MinReplicas worth of capacity (0 for gangs without a Required level), [G + q] = pods of clique q that fit on the
 * shard's nodes.  The caller sums the vectors of all ranks with ONE all-reduce (NCCL over NVLink; gloo in the CPU tests) and
 * reads cluster-wide feasibility off the sum: a gang with a Required level and sum 0, or a clique whose summed capacity is
 * below MinReplicas, cannot be admitted from this snapshot.  With world <= 1 the shard is the whole table. */
int32_t grove_shard_range(grove_engine_t* e, uint32_t* lo, uint32_t* hi); /* topology-sorted node indices [lo, hi) */
int32_t grove_run_score_pass(grove_engine_t* e, float* ms);
int32_t grove_shard_summary_device(grove_engine_t* e, void* d_out, uint32_t cap_words);

/* ---- device-resident variants (inputs already in HBM; used by bench.py `value`) -------------- */
/* d_nodes: device pointer to n grove_node_t in caller order, labels identical to the last load.  Read ASYNCHRONOUSLY on the
 * engine's stream: the buffer must stay valid and unchanged until the next blocking call on the handle (grove_run_cycle,
 * grove_run_score_pass, grove_get_nodes) has returned. */
int32_t grove_load_nodes_device(grove_engine_t* e, const void* d_nodes, uint32_t n);

/* ---- introspection for the parity tests (sorted node order; see DESIGN.md "Data layout") ------ */
int32_t grove_debug_get_perm(grove_engine_t* e, uint32_t* sorted_to_caller, uint32_t cap);
/* K1 / K2 output over the cycle-start snapshot: fit bitmap row (ceil(n/32) words) and score row (n bytes) of one clique */
int32_t grove_debug_get_fit_row(grove_engine_t* e, uint32_t clique, uint32_t* words, uint32_t cap_words);
int32_t grove_debug_get_score_row(grove_engine_t* e, uint32_t clique, uint8_t* bytes, uint32_t cap_bytes);
#ifdef __cplusplus
}
#endif
#endif /* GROVE_PLACE_H */
