// kernels.cuh -- sm_100a kernels of the gang-placement cycle (umbrella include).
//
//   tables.cuh   k_gather / k_scatter / k_update / k_anchor               table (un)packing
//   fit.cuh      k_fit  K1  node x clique-signature resource-fit bitmap   (Filter; oracle: fit())
//                k_cap8 / k_capsum  per-signature pod capacities and their per-domain sum / max
//   score.cuh    k_score K2 topology-distance score matrix, u8            (Score;  oracle: closeness())
//   admit.cuh    k_eval  K3  one gang against its view: all-or-nothing admission, warp-cooperative packing
//   relax.cuh    k_select / k_apply / k_detect / k_fold    the relaxation that makes the evaluations of all gangs
//                agree with the sequential, priority-ordered pass; k_fin_* / k_emit  outputs
//
// Semantics are DESIGN.md "Placement semantics"; the reference contract they restate is cited in
// include/grove_place.h.  Integer / compare work only: no tensor cores, no floating point.
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "fit.cuh"
#include "score.cuh"
#include "admit.cuh"
#include "relax.cuh"
