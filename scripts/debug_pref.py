"""debug helper: first gang whose status differs from the oracle on a random Preferred case"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_random_parity_gpu import random_case
from grove_b200.engine import PlacementEngine
from grove_b200 import tables as T
from oracle import oracle_py as O

seeds = [int(x) for x in sys.argv[1:]] or [4000]
for seed in seeds:
    nodes, L, (g, c, s) = random_case(seed, big=seed >= 4000, pref=True)
    for mr in (1, 0):
        ref = O.run_cycle(nodes, L, g, c, s, threads=8, max_rounds=mr)
        with PlacementEngine(L, max_rounds=mr) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s); st = e.run_cycle()
            gs = e.gang_status()
            bad = np.nonzero(gs != ref["status"])[0]
            print("seed", seed, "max_rounds", mr, "L", L, "n", len(nodes), "G", len(g), "rounds", st["rounds"], ref["stats"]["rounds"], "mismatches", len(bad))
            for gi in bad[:3]:
                print("  gang", gi, "level", g["level"][gi], "pref", g["preferred"][gi], "gpu", gs[gi], "ref", ref["status"][gi])
                for si in range(g["n_scopes"][gi]):
                    sc = s[g["scope_off"][gi] + si]
                    print("    scope", sc["level"], sc["preferred1"], [(int(q["min_replicas"]), int(q["replicas"]), int(q["level"]), int(q["scope"]) >> 5)
                                                                         for q in c[g["clique_off"][gi] + sc["first_clique"]: g["clique_off"][gi] + sc["first_clique"] + sc["n_cliques"]]])
        if len(bad):
            break
