// grove_host.hpp -- host side above the C ABI, in C++ because the reference is compiled code (Go) and
// no Go toolchain exists in this image.  It mirrors, name for name, the reference interfaces that sit
// either side of the placement hot path:
//
//   PodGang / PodGroup / TopologyConstraint ...   scheduler/api/core/v1alpha1/podgang.go:51-190
//   Backend / TopologyAwareSchedBackend           operator/internal/scheduler/types.go:37-96
//   ComputeExpectedPodGangs (the producer)        operator/internal/controller/podcliqueset/components/podgang/syncflow.go:145-371
//   name generation                               operator/api/common/namegen.go:70-117
//   CreatePodGroupsForPodGang                     .../components/podgang/podgang.go:165-186
//   isBasePodGangScheduled / scheduling gates     operator/internal/controller/podclique/components/pod/syncflow.go:255-358
//
// (all paths relative to /root/reference).  Error convention: Go returns `error`; here every fallible
// call returns std::optional<Error> (nullopt == nil), and Error carries the GroveError fields
// (operator/internal/errors/errors.go:41-86).
#pragma once
#include <cstdint>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <optional>
#include <string>
#include <vector>

#include "../../../include/grove_place.h"

namespace grove::host {

// ---- errors ------------------------------------------------------------------------------------
struct Error {
  std::string code;       // e.g. "ERR_SYNC_PODGANG"
  std::string operation;  // e.g. "SyncPodGang"
  std::string message;
};
using Err = std::optional<Error>;

// ---- scheduler/api/core/v1alpha1/podgang.go ------------------------------------------------------
struct NamespacedName { std::string Namespace, Name; };                              // :133-139
struct TopologyPackConstraint { std::optional<std::string> Required, Preferred; };   // :101-118 (label KEYS)
struct TopologyConstraint { std::optional<TopologyPackConstraint> PackConstraint; }; // :93-99
struct PodGroup {                                                                    // :75-91
  std::string Name;
  std::vector<NamespacedName> PodReferences;
  int32_t MinReplicas = 0;
  std::optional<TopologyConstraint> Topology;
};
struct TopologyConstraintGroupConfig {                                               // :120-131
  std::string Name;
  std::vector<std::string> PodGroupNames;
  std::optional<TopologyConstraint> Topology;
};
struct PodGangSpec {                                                                 // :51-72
  std::vector<PodGroup> PodGroups;
  std::optional<TopologyConstraint> Topology;
  std::vector<TopologyConstraintGroupConfig> TopologyConstraintGroupConfigs;
  std::string PriorityClassName;
  std::optional<NamespacedName> ReuseReservationRef;
};
enum class PodGangPhase { Pending, Starting, Running };                              // :141-150
struct PodGangStatus {                                                               // :182-190
  PodGangPhase Phase = PodGangPhase::Pending;
  bool Scheduled = false;                 // condition PodGangConditionTypeScheduled :155
  std::string ScheduledReason;            // "" | "Unschedulable" | "BaseNotScheduled" | "Gated"
  std::string ScheduledMessage;           // condition message (why a PodGang cannot be handed to the engine)
  std::optional<double> PlacementScore;   // :187-189
  uint32_t UnboundPods = 0;               // best-effort pods (beyond MinReplicas) of a scheduled PodGang that found no node
  bool DisruptionTarget = false;          // condition PodGangConditionTypeDisruptionTarget :166-170, reason 1: "PodGang is preempted
  std::string DisruptionMessage;          // by a higher priority PodGang" -- names the preemptor; the caller terminates the pods
};
struct PodGang {
  std::string Namespace, Name;
  std::map<std::string, std::string> Labels;  // grove.io/scheduler-name etc.
  PodGangSpec Spec;
  PodGangStatus Status;
  // not part of the CRD: what the pod template of each PodGroup requests (a scheduler reads it from the Pods)
  struct Requests { uint32_t cpu_milli = 0, mem_mib = 0; uint16_t gpu = 0; std::map<std::string, std::string> nodeSelector; std::vector<std::string> tolerationKeys; };
  std::map<std::string, Requests> PodGroupRequests;  // by PodGroup name
  std::string BasePodGangName;                        // scaled PodGang: gated behind this one (syncflow.go:319-358)
  bool Gated = false;                                 // pods still carry grove.io/podgang-pending-creation
};

// ---- operator/api/core/v1alpha1: the slice of PodCliqueSet the producer reads ----------------------
struct TopologyLevel { std::string Domain, Key; };                                   // clustertopology.go:111-124
struct PackDomain { std::string packDomain; };                                       // TopologyConstraint{PackDomain}
struct PodCliqueTemplateSpec {
  std::string Name;
  int32_t Replicas = 1;
  std::optional<int32_t> MinAvailable;          // defaulted to Replicas (defaulting/podcliqueset.go:74-83)
  std::optional<PackDomain> Topology;
  PodGang::Requests Requests;
};
struct PodCliqueScalingGroupConfig {
  std::string Name;
  std::vector<std::string> CliqueNames;
  std::optional<int32_t> Replicas;               // kubebuilder default 1 (podcliqueset.go:308-309)
  std::optional<int32_t> MinAvailable;           // kubebuilder default 1 (podcliqueset.go:321-322)
  std::optional<PackDomain> Topology;
};
struct PodCliqueSet {
  std::string Namespace = "default", Name;
  int32_t Replicas = 1;
  std::vector<PodCliqueTemplateSpec> Cliques;
  std::vector<PodCliqueScalingGroupConfig> PodCliqueScalingGroupConfigs;
  std::optional<PackDomain> Topology;
  std::string PriorityClassName;
};

// ---- namegen.go ------------------------------------------------------------------------------------
std::string GeneratePodCliqueName(const std::string& owner, int replica, const std::string& pclqTemplateName);  // :70-72
std::string GeneratePodCliqueScalingGroupName(const std::string& pcs, int replica, const std::string& name);    // :76-78
std::string GenerateBasePodGangName(const std::string& pcs, int replica);                                       // :82-84
std::string CreatePodGangNameFromPCSGFQN(const std::string& pcsgFQN, int scaledPodGangIndex);                   // :88-90

// ---- the producer (syncflow.go) ----------------------------------------------------------------------
struct PclqInfo {            // syncflow.go:776-783
  std::string fqn;
  int32_t replicas = 0, minAvailable = 0;
  std::optional<TopologyConstraint> topologyConstraint;
  std::string templateName;
};
struct PodGangInfo {         // syncflow.go:763-774
  std::string fqn;
  std::optional<TopologyConstraint> topologyConstraint;
  std::vector<PclqInfo> pclqs;
  std::vector<TopologyConstraintGroupConfig> pcsgTopologyConstraints;
  std::string baseFqn;       // "" for base PodGangs
};
// applies the defaulting webhook's MinAvailable := Replicas, then computeExpectedPodGangs
Err ComputeExpectedPodGangs(const PodCliqueSet& pcs, const std::vector<TopologyLevel>& topologyLevels, bool tasEnabled,
                            std::vector<PodGangInfo>* out);
// buildResource + createPodGroupsForPodGang (podgang.go:128-186); pod names are <pclq>-<ordinal>, sorted
PodGang BuildPodGang(const PodCliqueSet& pcs, const PodGangInfo& info);

// ---- webhook rule the packing relies on (webhook/admission/pcs/validation/topologyconstraints.go:195-268) ----
// A parent's pack domain must not be narrower than a child's (PodCliqueSet vs PodClique, PodCliqueSet vs
// PodCliqueScalingGroup, PodCliqueScalingGroup vs its PodCliques); domains unknown to the ClusterTopology are skipped.
// One FieldError per violating pair, in the reference's order, with the reference's field paths.
struct FieldError { std::string field, message; };
std::vector<FieldError> ValidateHierarchicalTopologyConstraints(const PodCliqueSet& pcs, const std::vector<std::string>& clusterTopologyDomains);

// ---- the gang predicate (podclique/components/pod/syncflow.go:316-358) ---------------------------------------
// A base PodGang is "scheduled" when every PodGroup has ScheduledReplicas >= MinReplicas; scaled PodGangs stay gated
// until then.  base == nullptr (PodGang not found) and a PodGroup without a PodClique status are errors: the caller
// requeues, as the reference does for every Get failure.  scheduledReplicas: PodClique name -> Status.ScheduledReplicas
// (podclique/reconcilestatus.go:134-141).
Err IsBasePodGangScheduled(const PodGang* base, const std::map<std::string, int32_t>& scheduledReplicas, bool* scheduled);

// What the bindings turn into on the way back (podclique/reconcilestatus.go:134-141, 255-274): ScheduledReplicas = pods of
// the PodClique with PodScheduled=True, and the PodCliqueScheduled condition on it.
struct Condition { std::string Type, Status, Reason, Message; };
Condition ComputePodCliqueScheduledCondition(int32_t scheduledReplicas, int32_t minAvailable);
// bindings of one cycle -> ScheduledReplicas per PodClique (pod names are <pclq fqn>-<ordinal>)
std::map<std::string, int32_t> CountScheduledReplicas(const std::vector<struct Binding>& bindings);

// Which pods reach a scheduler at all (checkAndRemovePodSchedulingGates, pod/syncflow.go:255-312): a pod loses its
// grove.io/podgang-pending-creation gate when it carries the gate, is already listed in its PodGang's PodReferences,
// and its PodGang is a base PodGang or its base PodGang is scheduled.  basePodGangName empty = the pod's PodGang is a
// base PodGang.  *removed: the gate is removed now; *skipped: the pod keeps its gate and is counted as skipped.
Err CheckPodSchedulingGate(bool podHasGate, bool podListedInPodGang, const std::string& basePodGangName, const PodGang* base,
                           const std::map<std::string, int32_t>& scheduledReplicas, bool* removed, bool* skipped);

// ---- nodes ------------------------------------------------------------------------------------------
struct Node {   // the slice of corev1.Node + bound Pods a scheduler snapshots; shape of kwok.py:74-117
  std::string Name;
  std::map<std::string, std::string> Labels;
  uint32_t alloc_cpu_milli = 0, alloc_mem_mib = 0; uint16_t alloc_gpu = 0, alloc_pods = 110;
  uint32_t used_cpu_milli = 0, used_mem_mib = 0; uint16_t used_gpu = 0, used_pods = 0;
  bool Unschedulable = false;                 // cordon
  std::vector<std::string> TaintKeys;         // NoSchedule taints
};

struct Binding { std::string PodNamespace, PodName, NodeName; };   // Pod.spec.nodeName

// ---- operator/internal/scheduler/types.go ---------------------------------------------------------------
class Backend {                                                                       // :37-58
 public:
  virtual ~Backend() = default;
  virtual std::string Name() const = 0;
  virtual Err Init() = 0;
  virtual Err SyncPodGang(const PodGang& podGang) = 0;          // read-only on the cache-owned object
  virtual Err OnPodGangDelete(const PodGang& podGang) = 0;
  virtual void PreparePod(std::string* schedulerName) const = 0; // sets pod.spec.schedulerName
  virtual Err ValidatePodCliqueSet(const PodCliqueSet& pcs) const = 0;
};
struct GroupVersionResource { std::string Group, Version, Resource; };                // schema.GroupVersionResource
class TopologyAwareSchedBackend {                                                     // :64-96
 public:
  virtual ~TopologyAwareSchedBackend() = default;
  virtual GroupVersionResource TopologyGVR() const = 0;                               // :65-68
  virtual std::string TopologyResourceName(const std::string& clusterTopologyName) const = 0;
  virtual Err SyncTopology(const std::vector<TopologyLevel>& levels) = 0;
  virtual Err OnTopologyDelete() = 0;
  // (inSync, message)
  virtual std::pair<bool, std::string> CheckTopologyDrift(const std::vector<TopologyLevel>& levels) const = 0;
};

// packed tables of one cycle, as handed to libgrove_place.so
struct Tables {
  std::vector<grove_node_t> nodes;
  std::vector<grove_gang_t> gangs;
  std::vector<grove_clique_t> cliques;
  std::vector<grove_scope_t> scopes;
  std::vector<std::string> gangNames;                 // row -> PodGang "<ns>/<name>"
  std::vector<std::pair<uint32_t, uint32_t>> cliqueOf; // clique row -> (gang row, PodGroup index in Spec.PodGroups)
  std::vector<uint8_t> remainder;                       // gang row -> 1: unbound pods of an already scheduled PodGang (MinReplicas 0)
  std::vector<uint32_t> refBase;                        // clique row -> first PodReference its placement entries bind (remainders)
  std::map<std::string, std::string> skipped;           // PodGang "<ns>/<name>" -> why it was left out of this pass
};

// The `gpu` scheduler backend: the third case of newBackendForProfile (manager/manager.go:35-52).
class GpuBackend : public Backend, public TopologyAwareSchedBackend {
 public:
  static constexpr const char* kName = "gpu-scheduler";
  explicit GpuBackend(int device = 0, std::string classLabelKey = "node_role.e2e.grove.nvidia.com");
  ~GpuBackend() override;
  std::string Name() const override { return kName; }
  Err Init() override;                       // creates the engine handle; fails (no fallback) without a CUDA device
  Err SyncPodGang(const PodGang& podGang) override;
  Err OnPodGangDelete(const PodGang& podGang) override;
  void PreparePod(std::string* schedulerName) const override { *schedulerName = kName; }
  Err ValidatePodCliqueSet(const PodCliqueSet& pcs) const override;
  // this backend keeps no topology CR of its own: the ordered level keys of the ClusterTopology ARE its topology
  // (kai/topology.go:40-46 returns KAI's CRD here); the CT controller's dynamic watch lands on the ClusterTopology itself
  GroupVersionResource TopologyGVR() const override { return {"grove.io", "v1alpha1", "clustertopologies"}; }
  std::string TopologyResourceName(const std::string& ct) const override { return ct; }
  Err SyncTopology(const std::vector<TopologyLevel>& levels) override;
  Err OnTopologyDelete() override;
  std::pair<bool, std::string> CheckTopologyDrift(const std::vector<TopologyLevel>& levels) const override;

  // encode the pending PodGangs + a node snapshot into packed tables (no GPU needed; unit-testable)
  Err Encode(const std::vector<Node>& nodes, Tables* out) const;
  // one scheduling cycle: Encode -> grove_load_nodes / submit / run_cycle -> bindings + PodGang statuses
  Err RunCycle(const std::vector<Node>& nodes, std::vector<Binding>* bindings, std::map<std::string, PodGangStatus>* statuses,
               grove_cycle_stats_t* stats = nullptr);
  size_t Pending() const { std::lock_guard<std::mutex> l(mu_); return pending_.size(); }   // PodGangs with unbound pods (unscheduled, or scheduled with a remainder)
  size_t Unscheduled() const { std::lock_guard<std::mutex> l(mu_); return pending_.size() - bound_.size(); }
  void SetPriorityClass(const std::string& name, int32_t value) { std::lock_guard<std::mutex> l(mu_); priorityClasses_[name] = value; }
  // Preemption (podgang.go:166-170): with it on, a cycle is grove_run_cycle_preempt -- PodGangs this backend scheduled earlier
  // are the RUNNING gangs (priority + what they hold per node, remembered from their bindings); a PodGang the ordinary pass
  // rejects may evict running PodGangs of a lower priority.  Victims come back in `statuses` with DisruptionTarget set and
  // are forgotten here: the caller deletes their pods (the node snapshot of the next cycle shows the resources as free).
  void SetPreemption(bool on) { std::lock_guard<std::mutex> l(mu_); preemption_ = on; }
  size_t Running() const { std::lock_guard<std::mutex> l(mu_); return running_.size(); }
  // why a PodGang cannot be handed to the engine (limits of the packed tables), or empty.  SyncPodGang accepts every
  // PodGang -- the reference's reconciler would requeue forever on an error -- and RunCycle reports these as
  // Unschedulable with this message instead of failing the pass for everybody.
  static std::string WhyNotEncodable(const PodGang& pg);

  // The cycle loop of the backend: SyncPodGang / OnPodGangDelete are called concurrently by
  // the PodGang reconcilers (controller/podgang/register.go:34-36, MaxConcurrentReconciles) and only touch the pending
  // set under the mutex; ONE thread takes node snapshots, runs cycles on the engine handle (not thread-safe) and hands
  // bindings + statuses to the binder callback.
  using SnapshotFn = std::function<std::vector<Node>()>;
  using BindFn = std::function<void(const std::vector<Binding>&, const std::map<std::string, PodGangStatus>&, const grove_cycle_stats_t&)>;
  Err Start(SnapshotFn snapshot, BindFn bind, std::chrono::milliseconds period);
  void Stop();
  uint64_t Cycles() const { return cycles_.load(); }

 private:
  int device_;
  std::string classKey_;
  std::vector<TopologyLevel> levels_;
  std::map<std::string, PodGang> pending_;   // "<ns>/<name>" -> copy (the cache-owned object is never retained)
  std::map<std::string, int32_t> priorityClasses_;
  std::map<std::string, std::string> lastNode_;  // scheduled PodGang -> a node it landed on (ReuseReservationRef hint)
  std::map<std::string, std::vector<uint32_t>> bound_;  // scheduled PodGang still pending with unbound pods -> pods bound per PodGroup
  struct Held { std::string node; uint32_t cpu_milli = 0, mem_mib = 0; uint16_t gpu = 0, pods = 0; };
  struct RunningGang { int32_t priority = 0; std::vector<Held> held; };
  std::map<std::string, RunningGang> running_;   // PodGangs scheduled by this backend and not deleted since
  bool preemption_ = false;
  grove_engine_t* engine_ = nullptr;
  uint32_t engineLevels_ = 0;
  mutable std::mutex mu_;        // pending set, bindings, levels: everything SyncPodGang / OnPodGangDelete / SyncTopology touch
  std::mutex cycleMu_;           // one cycle at a time on the engine handle
  std::thread loop_;
  std::condition_variable cv_;
  bool stop_ = false;
  std::atomic<uint64_t> cycles_{0};
  Err EncodeLocked(const std::vector<Node>& nodes, Tables* out) const;
};

}  // namespace grove::host
