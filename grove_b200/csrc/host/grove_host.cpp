// grove_host.cpp -- see grove_host.hpp.  Host-side mirror of the reference interfaces around the
// placement hot path; the placement itself happens behind the C ABI (libgrove_place.so).
#include "grove_host.hpp"

#include <algorithm>
#include <set>

namespace grove::host {

static Error mkerr(const char* code, const char* op, std::string msg) { return Error{code, op, std::move(msg)}; }

// ---- namegen.go ------------------------------------------------------------------------------------
std::string GeneratePodCliqueName(const std::string& owner, int replica, const std::string& t) {
  return owner + "-" + std::to_string(replica) + "-" + t;
}
std::string GeneratePodCliqueScalingGroupName(const std::string& pcs, int replica, const std::string& name) {
  return pcs + "-" + std::to_string(replica) + "-" + name;
}
std::string GenerateBasePodGangName(const std::string& pcs, int replica) { return pcs + "-" + std::to_string(replica); }
std::string CreatePodGangNameFromPCSGFQN(const std::string& pcsgFQN, int idx) { return pcsgFQN + "-" + std::to_string(idx); }

// ---- producer: syncflow.go:145-371 -------------------------------------------------------------------
// createTopologyPackConstraint (syncflow.go:349-371): packDomain -> node-label key through the ordered
// ClusterTopology levels; TAS disabled, no constraint, or an unknown domain => nil.  Only Required is set.
static std::optional<TopologyConstraint> createTopologyPackConstraint(const std::vector<TopologyLevel>& levels, bool tas,
                                                                      const std::optional<PackDomain>& req) {
  if (!tas || !req) return std::nullopt;
  for (const auto& l : levels)
    if (l.Domain == req->packDomain) {
      TopologyConstraint tc;
      tc.PackConstraint = TopologyPackConstraint{l.Key, std::nullopt};
      return tc;
    }
  return std::nullopt;  // stale domain: the reference logs and nullifies the constraint
}

static const PodCliqueTemplateSpec* findClique(const PodCliqueSet& pcs, const std::string& name) {
  for (const auto& c : pcs.Cliques) if (c.Name == name) return &c;
  return nullptr;
}
static const PodCliqueScalingGroupConfig* findScalingGroupForClique(const PodCliqueSet& pcs, const std::string& clique) {
  for (const auto& g : pcs.PodCliqueScalingGroupConfigs)
    if (std::find(g.CliqueNames.begin(), g.CliqueNames.end(), clique) != g.CliqueNames.end()) return &g;
  return nullptr;
}

// buildPodCliqueInfo (syncflow.go:335-345); MinAvailable defaulted as the webhook does (defaulting/podcliqueset.go:74-83)
static PclqInfo buildPodCliqueInfo(const std::vector<TopologyLevel>& levels, bool tas, const PodCliqueTemplateSpec& t, const std::string& fqn) {
  PclqInfo p;
  p.fqn = fqn; p.replicas = t.Replicas; p.minAvailable = t.MinAvailable.value_or(t.Replicas); p.templateName = t.Name;
  p.topologyConstraint = createTopologyPackConstraint(levels, tas, t.Topology);
  return p;
}

Err ComputeExpectedPodGangs(const PodCliqueSet& pcs, const std::vector<TopologyLevel>& levels, bool tas, std::vector<PodGangInfo>* out) {
  out->clear();
  // base PodGang per PCS replica: standalone cliques + PCSG replicas [0, minAvailable) (syncflow.go:172-275)
  for (int r = 0; r < pcs.Replicas; ++r) {
    PodGangInfo pg;
    pg.fqn = GenerateBasePodGangName(pcs.Name, r);
    pg.topologyConstraint = createTopologyPackConstraint(levels, tas, pcs.Topology);
    for (const auto& t : pcs.Cliques)
      if (!findScalingGroupForClique(pcs, t.Name))
        pg.pclqs.push_back(buildPodCliqueInfo(levels, tas, t, GeneratePodCliqueName(pcs.Name, r, t.Name)));
    for (const auto& cfg : pcs.PodCliqueScalingGroupConfigs) {
      const int minAvailable = cfg.MinAvailable.value_or(1);
      const std::string pcsgFQN = GeneratePodCliqueScalingGroupName(pcs.Name, r, cfg.Name);
      for (int ri = 0; ri < minAvailable; ++ri) {
        std::vector<std::string> fqns;
        for (const auto& cn : cfg.CliqueNames) {
          const PodCliqueTemplateSpec* t = findClique(pcs, cn);
          if (!t) return mkerr("ERR_SYNC_PODGANG", "ComputeExpectedPodGangs",
                               "PodCliqueScalingGroup \"" + cfg.Name + "\" references a PodClique \"" + cn + "\" that does not exist in the PodCliqueSet");
          const std::string fqn = GeneratePodCliqueName(pcsgFQN, ri, cn);
          pg.pclqs.push_back(buildPodCliqueInfo(levels, tas, *t, fqn));
          fqns.push_back(fqn);
        }
        if (tas && cfg.Topology) {  // one TopologyConstraintGroupConfig per PCSG replica (syncflow.go:262-271)
          TopologyConstraintGroupConfig gc;
          gc.Name = pcsgFQN + "-" + std::to_string(ri);
          gc.PodGroupNames = fqns;
          gc.Topology = createTopologyPackConstraint(levels, tas, cfg.Topology);
          pg.pcsgTopologyConstraints.push_back(std::move(gc));
        }
      }
    }
    out->push_back(std::move(pg));
  }
  // scaled PodGang per PCSG replica >= minAvailable (syncflow.go:277-333)
  for (int r = 0; r < pcs.Replicas; ++r)
    for (const auto& cfg : pcs.PodCliqueScalingGroupConfigs) {
      const std::string pcsgFQN = GeneratePodCliqueScalingGroupName(pcs.Name, r, cfg.Name);
      const int replicas = cfg.Replicas.value_or(1), minAvailable = cfg.MinAvailable.value_or(1);
      for (int idx = 0, pr = minAvailable; idx < replicas - minAvailable; ++idx, ++pr) {
        PodGangInfo pg;
        pg.fqn = CreatePodGangNameFromPCSGFQN(pcsgFQN, idx);
        pg.baseFqn = GenerateBasePodGangName(pcs.Name, r);
        for (const auto& cn : cfg.CliqueNames) {
          const PodCliqueTemplateSpec* t = findClique(pcs, cn);
          if (!t) return mkerr("ERR_SYNC_PODGANG", "ComputeExpectedPodGangs", "PodCliqueScalingGroup references unknown PodClique \"" + cn + "\"");
          pg.pclqs.push_back(buildPodCliqueInfo(levels, tas, *t, GeneratePodCliqueName(pcsgFQN, pr, cn)));
        }
        // PCSG constraint if set, else fall back to the PCS constraint (syncflow.go:311-324)
        if (tas) pg.topologyConstraint = cfg.Topology ? createTopologyPackConstraint(levels, tas, cfg.Topology)
                                                      : createTopologyPackConstraint(levels, tas, pcs.Topology);
        out->push_back(std::move(pg));
      }
    }
  return std::nullopt;
}

PodGang BuildPodGang(const PodCliqueSet& pcs, const PodGangInfo& info) {
  PodGang pg;
  pg.Namespace = pcs.Namespace; pg.Name = info.fqn;
  pg.Labels["grove.io/scheduler-name"] = GpuBackend::kName;
  pg.Spec.Topology = info.topologyConstraint;
  pg.Spec.TopologyConstraintGroupConfigs = info.pcsgTopologyConstraints;
  pg.Spec.PriorityClassName = pcs.PriorityClassName;
  pg.BasePodGangName = info.baseFqn;
  for (const auto& p : info.pclqs) {  // createPodGroupsForPodGang (podgang.go:165-186)
    PodGroup g;
    g.Name = p.fqn; g.MinReplicas = p.minAvailable; g.Topology = p.topologyConstraint;
    for (int i = 0; i < p.replicas; ++i) g.PodReferences.push_back({pcs.Namespace, p.fqn + "-" + std::to_string(i)});
    std::sort(g.PodReferences.begin(), g.PodReferences.end(), [](const NamespacedName& a, const NamespacedName& b) { return a.Name < b.Name; });
    pg.Spec.PodGroups.push_back(std::move(g));
    if (const PodCliqueTemplateSpec* t = findClique(pcs, p.templateName)) pg.PodGroupRequests[p.fqn] = t->Requests;
  }
  return pg;
}

// ---- GpuBackend ----------------------------------------------------------------------------------------
GpuBackend::GpuBackend(int device, std::string classLabelKey) : device_(device), classKey_(std::move(classLabelKey)) {}
GpuBackend::~GpuBackend() { Stop(); if (engine_) grove_engine_destroy(engine_); }

Err GpuBackend::Init() {
  if (levels_.empty()) return std::nullopt;  // the engine is created once the topology is known (SyncTopology)
  if (engine_ && engineLevels_ == levels_.size()) return std::nullopt;
  if (engine_) { grove_engine_destroy(engine_); engine_ = nullptr; }
  grove_config_t cfg{};
  cfg.abi_version = GROVE_ABI_VERSION; cfg.device = device_; cfg.n_levels = uint32_t(levels_.size());
  const int32_t rc = grove_engine_create(&cfg, &engine_);
  if (rc != GROVE_OK) { engine_ = nullptr; return mkerr("ERR_INIT_BACKEND", "Init", "grove_engine_create failed with " + std::to_string(rc) + " (there is no CPU fallback)"); }
  engineLevels_ = uint32_t(levels_.size());
  return std::nullopt;
}

Err GpuBackend::SyncTopology(const std::vector<TopologyLevel>& levels) {
  if (levels.empty() || levels.size() > GROVE_MAX_LEVELS)
    return mkerr("ERR_SYNC_TOPOLOGY", "SyncTopology", "the engine supports 1.." + std::to_string(GROVE_MAX_LEVELS) + " topology levels");
  std::lock_guard<std::mutex> l(mu_);
  levels_ = levels;  // ordered broadest -> narrowest, as desiredKAITopologyLevels hands them over (kai/topology.go:103-135)
  return std::nullopt;
}
Err GpuBackend::OnTopologyDelete() { std::lock_guard<std::mutex> l(mu_); levels_.clear(); return std::nullopt; }
std::pair<bool, std::string> GpuBackend::CheckTopologyDrift(const std::vector<TopologyLevel>& levels) const {
  std::lock_guard<std::mutex> l(mu_);
  if (levels.size() != levels_.size()) return {false, "level count differs"};
  for (size_t i = 0; i < levels.size(); ++i)
    if (levels[i].Key != levels_[i].Key) return {false, "level " + std::to_string(i) + " key " + levels_[i].Key + " != " + levels[i].Key};
  return {true, ""};
}

std::vector<FieldError> ValidateHierarchicalTopologyConstraints(const PodCliqueSet& pcs, const std::vector<std::string>& domains) {
  std::vector<FieldError> errs;
  auto index = [&domains](const std::string& d) { auto it = std::find(domains.begin(), domains.end(), d); return it == domains.end() ? -1 : int(it - domains.begin()); };
  auto violation = [&index](const std::string& parent, const std::string& child) {   // hasHierarchyViolation :195-202
    const int p = index(parent), c = index(child);
    return p != -1 && c != -1 && p > c;
  };
  auto msg = [](const std::string& pd, const std::string& pkind, const std::string& pname, const std::string& cd, const std::string& ckind, const std::string& cname) {
    const std::string who = pname.empty() ? pkind : pkind + " '" + pname + "'";
    return who + " topology constraint domain '" + pd + "' is narrower than " + ckind + " '" + cname + "' topology constraint domain '" + cd + "'";
  };
  const std::string root = "spec.template";
  if (pcs.Topology) {
    const std::string& pd = pcs.Topology->packDomain;
    for (const auto& c : pcs.Cliques)
      if (c.Topology && violation(pd, c.Topology->packDomain))
        errs.push_back({root + ".topologyConstraint", msg(pd, "PodCliqueSet", "", c.Topology->packDomain, "PodClique", c.Name)});
    for (const auto& g : pcs.PodCliqueScalingGroupConfigs)
      if (g.Topology && violation(pd, g.Topology->packDomain))
        errs.push_back({root + ".topologyConstraint", msg(pd, "PodCliqueSet", "", g.Topology->packDomain, "PodCliqueScalingGroup", g.Name)});
  }
  for (size_t i = 0; i < pcs.PodCliqueScalingGroupConfigs.size(); ++i) {
    const auto& g = pcs.PodCliqueScalingGroupConfigs[i];
    if (!g.Topology) continue;
    for (const auto& cn : g.CliqueNames) {
      auto it = std::find_if(pcs.Cliques.begin(), pcs.Cliques.end(), [&cn](const PodCliqueTemplateSpec& t) { return t.Name == cn; });
      if (it != pcs.Cliques.end() && it->Topology && violation(g.Topology->packDomain, it->Topology->packDomain))
        errs.push_back({root + ".podCliqueScalingGroups[" + std::to_string(i) + "].topologyConstraint",
                        msg(g.Topology->packDomain, "PodCliqueScalingGroup", g.Name, it->Topology->packDomain, "PodClique", it->Name)});
    }
  }
  return errs;
}

Err GpuBackend::ValidatePodCliqueSet(const PodCliqueSet& pcs) const {
  std::vector<std::string> domains;
  for (const auto& l : levels_) domains.push_back(l.Domain);
  if (auto errs = ValidateHierarchicalTopologyConstraints(pcs, domains); !errs.empty())
    return mkerr("ERR_VALIDATE_PCS", "ValidatePodCliqueSet", errs[0].field + ": " + errs[0].message);
  std::vector<PodGangInfo> infos;
  if (auto e = ComputeExpectedPodGangs(pcs, levels_, true, &infos)) return e;
  for (const auto& pg : infos) {
    int pods = 0;
    for (const auto& p : pg.pclqs) pods += p.replicas;
    if (pg.pclqs.size() > GROVE_MAX_GANG_CLIQUES || pods > int(GROVE_MAX_GANG_PODS) || pg.pcsgTopologyConstraints.size() + 1 > GROVE_MAX_GANG_SCOPES)
      return mkerr("ERR_VALIDATE_PCS", "ValidatePodCliqueSet", "PodGang " + pg.fqn + " exceeds the gpu-scheduler limits (32 PodCliques / 32 constraint groups / 128 pods per PodGang)");
  }
  return std::nullopt;
}

Err GpuBackend::SyncPodGang(const PodGang& podGang) {
  std::lock_guard<std::mutex> l(mu_);   // called concurrently by the PodGang reconcilers (podgang/register.go:34-36)
  pending_[podGang.Namespace + "/" + podGang.Name] = podGang;  // a copy: the cache-owned object is neither mutated nor retained
  return std::nullopt;
}
Err IsBasePodGangScheduled(const PodGang* base, const std::map<std::string, int32_t>& scheduledReplicas, bool* scheduled) {
  *scheduled = false;
  if (!base) return mkerr("ERR_GET_PODGANG", "Sync", "failed to get base PodGang");
  for (const auto& g : base->Spec.PodGroups) {
    auto it = scheduledReplicas.find(g.Name);
    if (it == scheduledReplicas.end())
      return mkerr("ERR_GET_PODCLIQUE", "Sync", "failed to get PodClique " + g.Name + " in namespace " + base->Namespace + " for base PodGang readiness check");
    if (it->second < g.MinReplicas) return std::nullopt;  // not scheduled yet: a legitimate state, no error
  }
  *scheduled = true;
  return std::nullopt;
}

Condition ComputePodCliqueScheduledCondition(int32_t scheduledReplicas, int32_t minAvailable) {
  const std::string counts = "expected at least: " + std::to_string(minAvailable) + ", found: " + std::to_string(scheduledReplicas);
  if (scheduledReplicas < minAvailable) return {"PodCliqueScheduled", "False", "InsufficientScheduledPods", "Insufficient scheduled pods. " + counts};
  return {"PodCliqueScheduled", "True", "SufficientScheduledPods", "Sufficient scheduled pods found. " + counts};
}

std::map<std::string, int32_t> CountScheduledReplicas(const std::vector<Binding>& bindings) {
  std::map<std::string, int32_t> n;
  for (const auto& b : bindings) n[b.PodName.substr(0, b.PodName.rfind('-'))]++;
  return n;
}

Err CheckPodSchedulingGate(bool podHasGate, bool podListedInPodGang, const std::string& basePodGangName, const PodGang* base,
                           const std::map<std::string, int32_t>& scheduledReplicas, bool* removed, bool* skipped) {
  *removed = false; *skipped = false;
  bool baseScheduled = true;   // computed once per PodClique, before any pod is looked at (syncflow.go:259-269)
  if (!basePodGangName.empty()) {
    if (auto e = IsBasePodGangScheduled(base, scheduledReplicas, &baseScheduled))
      return mkerr("ERR_REMOVE_POD_SCHEDULING_GATE", "Sync", "failed to check if base PodGang is scheduled for PodClique: " + e->message);
  }
  if (!podHasGate) return std::nullopt;                                 // nothing to do, not a skip
  if (!podListedInPodGang) { *skipped = true; return std::nullopt; }    // not yet in PodGang.Spec.PodGroups[].PodReferences
  if (!baseScheduled) { *skipped = true; return std::nullopt; }         // scaled PodGang behind an unscheduled base
  *removed = true;
  return std::nullopt;
}

Err GpuBackend::OnPodGangDelete(const PodGang& podGang) {
  std::lock_guard<std::mutex> l(mu_);
  const std::string key = podGang.Namespace + "/" + podGang.Name;
  pending_.erase(key); bound_.erase(key); running_.erase(key);
  return std::nullopt;
}

std::string GpuBackend::WhyNotEncodable(const PodGang& pg) {
  if (pg.Spec.PodGroups.empty()) return "spec.podgroups: a PodGang needs at least one PodGroup";
  if (pg.Spec.PodGroups.size() > GROVE_MAX_GANG_CLIQUES) return "spec.podgroups: more than " + std::to_string(GROVE_MAX_GANG_CLIQUES) + " PodGroups";
  if (pg.Spec.TopologyConstraintGroupConfigs.size() + 1 > GROVE_MAX_GANG_SCOPES) return "spec.topologyConstraintGroupConfigs: more than " + std::to_string(GROVE_MAX_GANG_SCOPES - 1) + " group configs";
  size_t pods = 0;
  for (const auto& p : pg.Spec.PodGroups) {
    if (p.MinReplicas < 0 || p.MinReplicas > 255) return "spec.podgroups[" + p.Name + "].minReplicas: must be within 0..255";
    if (p.PodReferences.size() > 255) return "spec.podgroups[" + p.Name + "].podReferences: more than 255 pods";
    if (size_t(p.MinReplicas) > p.PodReferences.size()) return "spec.podgroups[" + p.Name + "].minReplicas: exceeds the pods referenced";
    pods += p.PodReferences.size();
  }
  if (pods > GROVE_MAX_GANG_PODS) return "spec.podgroups: more than " + std::to_string(GROVE_MAX_GANG_PODS) + " pods in one PodGang";
  for (const auto& gc : pg.Spec.TopologyConstraintGroupConfigs)
    for (const auto& n : gc.PodGroupNames)
      if (std::none_of(pg.Spec.PodGroups.begin(), pg.Spec.PodGroups.end(), [&n](const PodGroup& p) { return p.Name == n; }))
        return "spec.topologyConstraintGroupConfigs[" + gc.Name + "].podGroupNames: unknown PodGroup " + n;
  return "";
}

Err GpuBackend::Encode(const std::vector<Node>& nodes, Tables* out) const {
  std::lock_guard<std::mutex> l(mu_);
  return EncodeLocked(nodes, out);
}

Err GpuBackend::EncodeLocked(const std::vector<Node>& nodes, Tables* out) const {
  *out = Tables{};
  if (levels_.empty()) return mkerr("ERR_SYNC_PODGANG", "Encode", "no ClusterTopology levels synced");
  const uint32_t L = uint32_t(levels_.size());
  // nodes: interned label value per level; selector class from one label key; taints per class
  std::vector<std::map<std::string, uint32_t>> intern(L);
  std::map<std::string, uint32_t> classOf;  // label value -> class id (0 = label absent)
  std::vector<std::set<std::string>> classTaints(1);
  std::vector<std::string> classValue(1, "");
  out->nodes.resize(nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) {
    const Node& n = nodes[i];
    grove_node_t& r = out->nodes[i];
    r.free_cpu_milli = n.alloc_cpu_milli > n.used_cpu_milli ? n.alloc_cpu_milli - n.used_cpu_milli : 0;
    r.free_mem_mib = n.alloc_mem_mib > n.used_mem_mib ? n.alloc_mem_mib - n.used_mem_mib : 0;
    r.free_gpu = uint16_t(n.alloc_gpu > n.used_gpu ? n.alloc_gpu - n.used_gpu : 0);
    r.free_pods = uint16_t(n.alloc_pods > n.used_pods ? n.alloc_pods - n.used_pods : 0);
    uint32_t cls = 0;
    if (auto it = n.Labels.find(classKey_); it != n.Labels.end()) {
      auto [ci, fresh] = classOf.emplace(it->second, uint32_t(classOf.size() + 1));
      if (fresh) { classTaints.emplace_back(); classValue.push_back(it->second); }
      cls = ci->second;
      if (cls > 15) return mkerr("ERR_SYNC_PODGANG", "Encode", "more than 15 values of the selector-class label " + classKey_);
    }
    for (const auto& t : n.TaintKeys) classTaints[cls].insert(t);
    r.flags = (n.Unschedulable ? 0u : GROVE_NODE_SCHEDULABLE) | (cls << GROVE_NODE_CLASS_SHIFT);
    for (uint32_t l = 0; l < GROVE_MAX_LEVELS; ++l) {
      r.dom[l] = GROVE_DOM_ABSENT;
      if (l >= L) continue;
      if (auto it = n.Labels.find(levels_[l].Key); it != n.Labels.end())
        r.dom[l] = intern[l].emplace(it->second, uint32_t(intern[l].size())).first->second;
    }
  }
  // (Required, Preferred) label keys -> level indices; a Preferred level that is not deeper than the Required
  // one adds nothing (best effort, podgang.go:110-117) and is dropped
  auto levelOf = [this](const std::optional<TopologyConstraint>& tc, uint8_t* lvl, uint8_t* pref) -> Err {
    *lvl = GROVE_LEVEL_NONE; *pref = GROVE_LEVEL_NONE;
    if (!tc || !tc->PackConstraint) return std::nullopt;
    // a key that is no (longer a) level of the synced ClusterTopology is dropped, not an error: the operator does the same
    // for a pack domain it cannot find (createTopologyPackConstraint, syncflow.go:349-371, logs and returns nil)
    auto find = [this](const std::string& key, const char*, uint8_t* o) -> Err {
      for (size_t l = 0; l < levels_.size(); ++l) if (levels_[l].Key == key) { *o = uint8_t(l); return std::nullopt; }
      return std::nullopt;
    };
    if (tc->PackConstraint->Required) if (auto e = find(*tc->PackConstraint->Required, "Required", lvl)) return e;
    if (tc->PackConstraint->Preferred) if (auto e = find(*tc->PackConstraint->Preferred, "Preferred", pref)) return e;
    if (*pref != GROVE_LEVEL_NONE && *lvl != GROVE_LEVEL_NONE && *pref <= *lvl) *pref = GROVE_LEVEL_NONE;
    return std::nullopt;
  };
  // Row order: PodGangs that still have to get their minimum (key order), then the REMAINDERS of scheduled PodGangs:
  // pods that found no node when the gang was admitted stay Pending and are retried with MinReplicas 0 (the gang
  // guarantee is met), after everybody's minimum -- what the e2e suites wait for step by step
  // (gang_scheduling_test.go GS5-GS12).  A remainder carries no placement of its own, so it is only resubmitted for
  // PodGangs without pack constraints (with constraints its pods would have to rejoin the domains chosen earlier).
  auto constrained = [](const PodGang& pg) {
    auto has = [](const std::optional<TopologyConstraint>& tc) { return tc && tc->PackConstraint && (tc->PackConstraint->Required || tc->PackConstraint->Preferred); };
    if (has(pg.Spec.Topology)) return true;
    for (const auto& gc : pg.Spec.TopologyConstraintGroupConfigs) if (has(gc.Topology)) return true;
    for (const auto& p : pg.Spec.PodGroups) if (has(p.Topology)) return true;
    return false;
  };
  // The best-effort pods of a PodGang that is not scheduled yet follow the same route: it goes in as its minimum
  // only, and its surplus as a remainder row gated behind it (base_gang) -- minimums of every gang before anybody's
  // surplus, which is what the suites' step descriptions say happens (GS8 :583, GS10 :779, GS12 :1023).
  auto hasSurplus = [](const PodGang& pg) { for (const auto& p : pg.Spec.PodGroups) if (int64_t(p.PodReferences.size()) > p.MinReplicas) return true; return false; };
  // PodGangs the packed tables cannot hold are left out of the pass (and so is a scaled PodGang behind one); everybody else
  // is scheduled as usual
  for (const auto& kv : pending_) { const std::string why = WhyNotEncodable(kv.second); if (!why.empty()) out->skipped[kv.first] = why; }
  for (const auto& kv : pending_)
    if (!out->skipped.count(kv.first) && !kv.second.BasePodGangName.empty() && out->skipped.count(kv.second.Namespace + "/" + kv.second.BasePodGangName))
      out->skipped[kv.first] = "base PodGang " + kv.second.BasePodGangName + " cannot be scheduled: " + out->skipped[kv.second.Namespace + "/" + kv.second.BasePodGangName];
  std::vector<const std::string*> order;
  for (const auto& kv : pending_) if (!bound_.count(kv.first) && !out->skipped.count(kv.first)) order.push_back(&kv.first);
  const size_t nFull = order.size();
  for (const auto& kv : pending_) {
    if (out->skipped.count(kv.first)) continue;
    const bool scheduled = bound_.count(kv.first) != 0;
    if (!constrained(kv.second) && (scheduled || hasSurplus(kv.second))) order.push_back(&kv.first);
  }
  std::map<std::string, uint32_t> row;  // PodGang key -> gang row (PodGangs still to be scheduled only: what a base_gang can name)
  for (size_t i = 0; i < nFull; ++i) row.emplace(*order[i], uint32_t(i));
  std::map<std::string, uint32_t> nodeIndex;
  for (size_t i = 0; i < nodes.size(); ++i) nodeIndex[nodes[i].Name] = uint32_t(i);
  for (size_t oi = 0; oi < order.size(); ++oi) {
    const std::string& key = *order[oi];
    const PodGang& pg = pending_.at(key);
    const bool remainder = oi >= nFull;
    const bool split = !remainder && !constrained(pg) && hasSurplus(pg);          // this row carries the minimum only
    std::vector<uint32_t> minOnly;                                                 // surplus row of an unscheduled PodGang: its minimum is row `row[key]`
    const std::vector<uint32_t>* done = nullptr;                                   // pods covered elsewhere, per PodGroup
    if (remainder) {
      if (auto it = bound_.find(key); it != bound_.end()) done = &it->second;
      else { for (const auto& p : pg.Spec.PodGroups) minOnly.push_back(uint32_t(std::max<int32_t>(0, p.MinReplicas))); done = &minOnly; }
    }
    grove_gang_t g{};
    g.clique_off = uint32_t(out->cliques.size()); g.scope_off = uint32_t(out->scopes.size());
    g.anchor_node = GROVE_NONE_U32; g.base_gang = GROVE_NONE_U32; g.preferred = GROVE_LEVEL_NONE;
    g.flags = pg.Gated ? GROVE_GANG_GATED : 0;
    if (auto e = levelOf(pg.Spec.Topology, &g.level, &g.preferred)) return e;
    if (auto it = priorityClasses_.find(pg.Spec.PriorityClassName); it != priorityClasses_.end()) g.priority = it->second;
    if (!remainder && !pg.BasePodGangName.empty())   // a base that is already scheduled gates nothing
      if (auto it = row.find(pg.Namespace + "/" + pg.BasePodGangName); it != row.end()) g.base_gang = it->second;
    if (remainder && !minOnly.empty()) g.base_gang = row.at(key);   // surplus of a PodGang whose minimum is in this very pass
    if (remainder) {  // stay close to where the gang landed
      if (auto it = lastNode_.find(key); it != lastNode_.end()) if (auto nt = nodeIndex.find(it->second); nt != nodeIndex.end()) g.anchor_node = nt->second;
    } else if (pg.Spec.ReuseReservationRef) {  // locality hint: score distance against where that PodGang was placed
      auto it = lastNode_.find(pg.Spec.ReuseReservationRef->Namespace + "/" + pg.Spec.ReuseReservationRef->Name);
      if (it != lastNode_.end()) if (auto nt = nodeIndex.find(it->second); nt != nodeIndex.end()) g.anchor_node = nt->second;
    }
    // scopes: the loose PodGroups first (one implicit scope), then one scope per TopologyConstraintGroupConfig
    std::set<std::string> grouped;
    for (const auto& gc : pg.Spec.TopologyConstraintGroupConfigs) for (const auto& n : gc.PodGroupNames) grouped.insert(n);
    struct ScopeRows { uint8_t first, pref; std::vector<uint32_t> second; };  // (Required level, Preferred level, PodGroup indices)
    std::vector<ScopeRows> scopes;
    std::vector<uint32_t> loose;
    auto left = [&](uint32_t i) { return pg.Spec.PodGroups[i].PodReferences.size() - (done ? (*done)[i] : 0u); };
    for (uint32_t i = 0; i < pg.Spec.PodGroups.size(); ++i) if (!grouped.count(pg.Spec.PodGroups[i].Name) && (!remainder || left(i))) loose.push_back(i);
    if (!loose.empty()) scopes.push_back({uint8_t(GROVE_LEVEL_NONE), uint8_t(GROVE_LEVEL_NONE), loose});
    for (const auto& gc : pg.Spec.TopologyConstraintGroupConfigs) {
      uint8_t lvl, pref; if (auto e = levelOf(gc.Topology, &lvl, &pref)) return e;
      std::vector<uint32_t> members;
      for (const auto& n : gc.PodGroupNames) {
        auto it = std::find_if(pg.Spec.PodGroups.begin(), pg.Spec.PodGroups.end(), [&n](const PodGroup& p) { return p.Name == n; });
        if (it == pg.Spec.PodGroups.end()) return mkerr("ERR_SYNC_PODGANG", "Encode", "group config " + gc.Name + " names unknown PodGroup " + n);
        const uint32_t gi = uint32_t(it - pg.Spec.PodGroups.begin());
        if (!remainder || left(gi)) members.push_back(gi);
      }
      if (!members.empty()) scopes.push_back({lvl, pref, members});
    }
    uint32_t rel = 0, pods = 0;
    for (size_t si = 0; si < scopes.size(); ++si) {
      grove_scope_t s{}; s.first_clique = uint16_t(rel); s.n_cliques = uint16_t(scopes[si].second.size()); s.level = scopes[si].first;
      s.preferred1 = scopes[si].pref == GROVE_LEVEL_NONE ? uint8_t(0) : uint8_t(scopes[si].pref + 1);
      out->scopes.push_back(s);
      for (uint32_t gi : scopes[si].second) {
        const PodGroup& p = pg.Spec.PodGroups[gi];
        grove_clique_t c{};
        PodGang::Requests rq;
        if (auto it = pg.PodGroupRequests.find(p.Name); it != pg.PodGroupRequests.end()) rq = it->second;
        c.req_cpu_milli = rq.cpu_milli; c.req_mem_mib = rq.mem_mib; c.req_gpu = rq.gpu;
        if (p.MinReplicas < 0 || p.MinReplicas > 255 || p.PodReferences.size() > 255 || size_t(p.MinReplicas) > p.PodReferences.size())
          return mkerr("ERR_SYNC_PODGANG", "Encode", "PodGroup " + p.Name + ": MinReplicas / PodReferences out of range");
        c.min_replicas = uint8_t(p.MinReplicas); c.replicas = uint8_t(p.PodReferences.size());
        if (remainder) { c.min_replicas = 0; c.replicas = uint8_t(left(gi)); }
        if (split) c.replicas = c.min_replicas;
        uint8_t cpref;
        if (auto e = levelOf(p.Topology, &c.level, &cpref)) return e;
        c.scope = GROVE_CLIQUE_SCOPE_PREF(uint32_t(si), cpref);
        // class mask: classes whose label value satisfies the nodeSelector and whose taints are all tolerated
        uint16_t mask = 0;
        auto sel = rq.nodeSelector.find(classKey_);
        for (uint32_t cl = 0; cl < classValue.size(); ++cl) {
          if (sel != rq.nodeSelector.end() && (cl == 0 || classValue[cl] != sel->second)) continue;
          bool tolerated = true;
          for (const auto& t : classTaints[cl]) tolerated &= std::find(rq.tolerationKeys.begin(), rq.tolerationKeys.end(), t) != rq.tolerationKeys.end();
          if (tolerated) mask |= uint16_t(1u << cl);
        }
        c.class_mask = mask;
        out->cliques.push_back(c);
        out->cliqueOf.push_back({uint32_t(out->gangs.size()), gi});
        out->refBase.push_back(done ? (*done)[gi] : 0u);
        pods += c.replicas; ++rel;
      }
    }
    if (rel == 0 || rel > GROVE_MAX_GANG_CLIQUES || scopes.size() > GROVE_MAX_GANG_SCOPES || pods > GROVE_MAX_GANG_PODS)
      return mkerr("ERR_SYNC_PODGANG", "Encode", "PodGang " + key + " exceeds the gpu-scheduler limits");
    g.n_cliques = uint16_t(rel); g.n_scopes = uint16_t(scopes.size());
    out->gangs.push_back(g);
    out->gangNames.push_back(key);
    out->remainder.push_back(remainder ? 1 : 0);
  }
  return std::nullopt;
}

Err GpuBackend::RunCycle(const std::vector<Node>& nodes, std::vector<Binding>* bindings, std::map<std::string, PodGangStatus>* statuses,
                         grove_cycle_stats_t* stats) {
  std::lock_guard<std::mutex> cycle(cycleMu_);   // the engine handle is not thread-safe: one cycle at a time
  bindings->clear(); statuses->clear();
  if (stats) *stats = grove_cycle_stats_t{};
  Tables t;
  std::map<std::string, PodGang> snap;            // the PodGangs of this pass: reconcilers may add / delete others meanwhile
  {
    std::lock_guard<std::mutex> l(mu_);
    if (pending_.empty()) return std::nullopt;
    if (auto e = Init()) return e;
    if (!engine_) return mkerr("ERR_SYNC_PODGANG", "RunCycle", "no ClusterTopology synced");
    if (auto e = EncodeLocked(nodes, &t)) return e;
    snap = pending_;
  }
  for (const auto& kv : t.skipped) { PodGangStatus s; s.ScheduledReason = "Unschedulable"; s.ScheduledMessage = kv.second; (*statuses)[kv.first] = s; }
  if (t.gangs.empty()) return std::nullopt;
  auto chk = [this](int32_t rc, const char* what) -> Err {
    if (rc == GROVE_OK) return std::nullopt;
    return mkerr("ERR_SYNC_PODGANG", what, std::string(grove_last_error(engine_)) + " (" + std::to_string(rc) + ")");
  };
  if (auto e = chk(grove_load_nodes(engine_, t.nodes.data(), uint32_t(t.nodes.size())), "grove_load_nodes")) return e;
  if (auto e = chk(grove_submit_gangs(engine_, t.gangs.data(), uint32_t(t.gangs.size()), t.cliques.data(), uint32_t(t.cliques.size()),
                                      t.scopes.data(), uint32_t(t.scopes.size())), "grove_submit_gangs")) return e;
  grove_cycle_stats_t st{};
  // running PodGangs -> the reclaim pass's tables (holdings on nodes that left the snapshot are dropped)
  std::vector<grove_running_gang_t> run; std::vector<grove_holding_t> held; std::vector<std::string> runNames;
  bool preempt = false;
  {
    std::lock_guard<std::mutex> l(mu_);
    preempt = preemption_ && !running_.empty();
    if (preempt) {
      std::map<std::string, uint32_t> nodeIdx;
      for (uint32_t i = 0; i < nodes.size(); ++i) nodeIdx[nodes[i].Name] = i;
      for (const auto& kv : running_) {
        if (snap.count(kv.first)) continue;   // still pending with a remainder: not evictable while it is being placed
        grove_running_gang_t r{kv.second.priority, uint32_t(held.size()), 0u, 0u};
        for (const auto& h : kv.second.held) {
          auto it = nodeIdx.find(h.node);
          if (it == nodeIdx.end()) continue;
          held.push_back(grove_holding_t{it->second, h.cpu_milli, h.mem_mib, h.gpu, h.pods}); ++r.n_holdings;
        }
        run.push_back(r); runNames.push_back(kv.first);
      }
    }
  }
  if (preempt) { if (auto e = chk(grove_run_cycle_preempt(engine_, run.data(), uint32_t(run.size()), held.data(), uint32_t(held.size()), &st), "grove_run_cycle_preempt")) return e; }
  else if (auto e = chk(grove_run_cycle(engine_, &st), "grove_run_cycle")) return e;
  if (stats) *stats = st;
  std::vector<grove_placement_t> pl(st.pods_bound + 1);
  uint32_t n = 0;
  if (auto e = chk(grove_get_placements(engine_, pl.data(), uint32_t(pl.size()), &n), "grove_get_placements")) return e;
  std::vector<grove_gang_status_t> gs(t.gangs.size());
  if (auto e = chk(grove_get_gang_status(engine_, gs.data(), uint32_t(gs.size())), "grove_get_gang_status")) return e;
  std::vector<uint32_t> seen(t.cliques.size(), 0);  // the r-th entry of a clique binds its r-th PodReference
  for (uint32_t i = 0; i < n; ++i) {
    const auto [grow, pgi] = t.cliqueOf[pl[i].clique];
    const PodGang& pg = snap.at(t.gangNames[grow]);
    const PodGroup& grp = pg.Spec.PodGroups[pgi];
    const NamespacedName& pod = grp.PodReferences[t.refBase[pl[i].clique] + seen[pl[i].clique]++];
    bindings->push_back({pod.Namespace, pod.Name, nodes[pl[i].node].Name});
  }
  for (size_t g = 0; g < gs.size(); ++g) {
    PodGangStatus s;
    switch (gs[g].state) {
      case GROVE_GANG_ADMITTED:
        s.Phase = PodGangPhase::Starting; s.Scheduled = true;
        if (!t.remainder[g]) s.PlacementScore = double(gs[g].score_num) / double(gs[g].score_den);  // a remainder re-reports no score
        break;
      case GROVE_GANG_REJECTED: s.ScheduledReason = "Unschedulable"; break;
      case GROVE_GANG_BASE_REJECTED: s.ScheduledReason = "BaseNotScheduled"; break;
      case GROVE_GANG_GATED_SKIP: s.ScheduledReason = "Gated"; break;
      default: s.ScheduledReason = "Pending"; break;
    }
    if (t.remainder[g] && statuses->count(t.gangNames[g])) continue;   // the minimum row of the same PodGang already spoke
    (*statuses)[t.gangNames[g]] = s;
  }
  // a scheduled PodGang leaves the pending set once every PodReference is bound; until then its unbound pods are retried
  // as a remainder (Encode).  Remember where gangs landed for ReuseReservationRef hints.
  auto constrained = [](const PodGang& pg) {
    auto has = [](const std::optional<TopologyConstraint>& tc) { return tc && tc->PackConstraint && (tc->PackConstraint->Required || tc->PackConstraint->Preferred); };
    if (has(pg.Spec.Topology)) return true;
    for (const auto& gc : pg.Spec.TopologyConstraintGroupConfigs) if (has(gc.Topology)) return true;
    for (const auto& p : pg.Spec.PodGroups) if (has(p.Topology)) return true;
    return false;
  };
  std::lock_guard<std::mutex> l(mu_);
  if (preempt) {   // victims: DisruptionTarget on the PodGang (podgang.go:166-170), forgotten as running
    uint32_t nv = 0;
    if (auto e = chk(grove_get_victims(engine_, nullptr, 0, &nv), "grove_get_victims")) return e;
    std::vector<grove_victim_t> vs(nv + 1);
    if (auto e = chk(grove_get_victims(engine_, vs.data(), uint32_t(vs.size()), &nv), "grove_get_victims")) return e;
    for (uint32_t i = 0; i < nv; ++i) {
      PodGangStatus s; s.Phase = PodGangPhase::Running; s.Scheduled = true;
      s.DisruptionTarget = true; s.DisruptionMessage = "preempted by higher priority PodGang " + t.gangNames[vs[i].preemptor];
      (*statuses)[runNames[vs[i].running]] = s;
      running_.erase(runNames[vs[i].running]);
    }
  }
  for (uint32_t i = 0; i < gs.size(); ++i) {
    if (gs[i].state != GROVE_GANG_ADMITTED) continue;
    const std::string& key = t.gangNames[i];
    auto pit = pending_.find(key);
    if (pit == pending_.end()) continue;   // deleted while the cycle ran
    const PodGang& pg = pit->second;
    {   // what the PodGang holds from now on (a remainder adds to what its minimum took)
      RunningGang& rg = running_[key];
      rg.priority = t.gangs[i].priority;
      for (uint32_t k = 0; k < gs[i].n_pods; ++k) {
        const grove_placement_t& e = pl[gs[i].placement_off + k];
        const grove_clique_t& q = t.cliques[e.clique];
        const std::string& nn = nodes[e.node].Name;
        auto it = std::find_if(rg.held.begin(), rg.held.end(), [&](const Held& h) { return h.node == nn; });
        if (it == rg.held.end()) { rg.held.push_back(Held{nn}); it = rg.held.end() - 1; }
        it->cpu_milli += q.req_cpu_milli; it->mem_mib += q.req_mem_mib; it->gpu = uint16_t(it->gpu + q.req_gpu); it->pods = uint16_t(it->pods + 1);
      }
    }
    if (gs[i].n_pods && !lastNode_.count(key)) lastNode_[key] = nodes[pl[gs[i].placement_off].node].Name;
    std::vector<uint32_t>& done = bound_[key];
    done.resize(pg.Spec.PodGroups.size(), 0u);
    for (uint32_t c = t.gangs[i].clique_off; c < t.gangs[i].clique_off + t.gangs[i].n_cliques; ++c) done[t.cliqueOf[c].second] += seen[c];
    bool all = true; uint32_t unbound = 0;
    for (size_t k = 0; k < done.size(); ++k) {
      all &= done[k] >= pg.Spec.PodGroups[k].PodReferences.size();
      if (done[k] < pg.Spec.PodGroups[k].PodReferences.size()) unbound += uint32_t(pg.Spec.PodGroups[k].PodReferences.size()) - done[k];
    }
    // A PodGang WITH pack constraints is not resubmitted as a remainder (its best-effort pods would have to rejoin the
    // domains chosen for the gang): it is scheduled, leaves the pending set, and the status says how many of its
    // best-effort pods stayed Pending (such gangs used to sit in the pending set for ever)
    if (!all && constrained(pg)) { (*statuses)[key].UnboundPods = unbound; all = true; }
    if (all) { pending_.erase(key); bound_.erase(key); }
  }
  return std::nullopt;
}

Err GpuBackend::Start(SnapshotFn snapshot, BindFn bind, std::chrono::milliseconds period) {
  std::lock_guard<std::mutex> l(mu_);
  if (loop_.joinable()) return mkerr("ERR_SYNC_PODGANG", "Start", "cycle loop already running");
  stop_ = false;
  loop_ = std::thread([this, snapshot, bind, period] {
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (cv_.wait_for(lk, period, [this] { return stop_; })) return;
        if (pending_.empty()) continue;
      }
      std::vector<Binding> b; std::map<std::string, PodGangStatus> st; grove_cycle_stats_t cs{};
      const std::vector<Node> nodes = snapshot();
      if (auto e = RunCycle(nodes, &b, &st, &cs)) continue;   // the next tick retries (the reconciler's requeue)
      cycles_.fetch_add(1);
      bind(b, st, cs);
    }
  });
  return std::nullopt;
}

void GpuBackend::Stop() {
  { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
  cv_.notify_all();
  if (loop_.joinable()) loop_.join();
}

}  // namespace grove::host
