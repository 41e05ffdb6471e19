// kernels.cuh -- sm_100a kernels of the gang-placement cycle (umbrella include).
//
//   tables.cuh   k_gather / k_scatter / k_update / k_anchor / k_prepare   table (un)packing, round bookkeeping
//   fit.cuh      k_fit  K1  node x clique-signature resource-fit bitmap   (Filter; oracle: fit())
//                k_cap8 / k_capsum  per-signature pod capacities and their per-domain sum / max
//   score.cuh    k_score K2 topology-distance score matrix, u8            (Score;  oracle: closeness())
//                k_alt_scores       joins the K2 stream with K3: score of every alternative
//   admit.cuh    k_admit_warp / k_admit  K3  per-gang all-or-nothing admission, K alternatives per gang
//   resolve.cuh  k_resolve  conflict resolution + commit of one round (cooperative launch);
//                k_finalize / k_emit  outputs
//
// Semantics are DESIGN.md "Placement semantics"; the reference contract they restate is cited in
// include/grove_place.h.  Integer / compare work only: no tensor cores, no floating point.
#pragma once
#include "common.cuh"
#include "tables.cuh"
#include "fit.cuh"
#include "score.cuh"
#include "admit.cuh"
#include "resolve.cuh"
