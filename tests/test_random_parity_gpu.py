"""-m gpu: differential test on seeded random snapshots that mix every feature of the table format:
ragged topologies (absent labels at any level, non-tree label sets), cordoned nodes, selector classes,
zero / partial requests, MinReplicas = 0, surplus replicas, nested Required levels at gang / scope /
clique, Preferred levels at gang / scope / clique (with and without a Required level above them),
priorities, explicit anchors, base-gang chains (incl. rejected and gated bases).  Every case must match
the oracle bit for bit."""
import numpy as np
import pytest

from grove_b200 import tables as T

pytestmark = pytest.mark.gpu


def random_case(seed, big=False, pref=False):
    rng = np.random.default_rng(seed)
    prng = np.random.default_rng(seed + 1_000_003)  # Preferred levels draw from their own stream: pref=False cases never change

    def preferred(req):
        """a level deeper than the unit's own Required one (it may be shallower than the parent's: then it is moot)"""
        lo = 0 if req is None else req + 1
        return int(prng.integers(lo, L)) if pref and lo < L and prng.random() < 0.45 else None

    L = int(rng.integers(1, 5))
    n = int(rng.integers(1, 400)) if not big else int(rng.integers(1500, 4000))
    fan = [int(rng.integers(2, 9)) for _ in range(L)]
    nodes = T.make_nodes(n)
    idx = np.arange(n)
    span = n
    for l in range(L):
        span = max(1, span // fan[l])
        nodes["dom"][:, l] = idx // span if l < L - 1 or rng.random() < 0.7 else idx
    if rng.random() < 0.4:  # non-tree: shuffle one level's ids
        l = int(rng.integers(0, L)); nodes["dom"][:, l] = rng.permutation(nodes["dom"][:, l])
    if rng.random() < 0.5:  # ragged: labels missing from some level down, for some nodes
        for i in rng.choice(n, size=max(1, n // 6), replace=False):
            nodes["dom"][i, int(rng.integers(0, L)):] = T.DOM_ABSENT
    nodes["free_cpu_milli"] = rng.integers(0, 64001, n)
    nodes["free_mem_mib"] = rng.integers(0, 524289, n)
    nodes["free_gpu"] = rng.integers(0, 9, n)
    nodes["free_pods"] = rng.integers(0, 6, n) if rng.random() < 0.3 else 110
    cls = rng.integers(0, 4, n)
    sched = rng.random(n) > 0.1
    nodes["flags"] = (sched * T.NODE_SCHEDULABLE) | (cls.astype(np.uint32) << T.NODE_CLASS_SHIFT)
    b = T.GangTableBuilder()
    G = int(rng.integers(1, 40)) if not big else int(rng.integers(700, 1500))  # big: many gangs relaxing at once
    for gi in range(G):
        glevel = None if rng.random() < 0.35 else int(rng.integers(0, L))
        scopes, pods = [], 0
        for _ in range(int(rng.integers(1, 4))):
            lo = -1 if glevel is None else glevel
            slevel = None if rng.random() < 0.5 or lo + 1 >= L else int(rng.integers(lo + 1, L))
            cliques = []
            for _ in range(int(rng.integers(1, 4))):
                lo2 = lo if slevel is None else slevel
                clevel = None if rng.random() < 0.5 or lo2 + 1 >= L else int(rng.integers(lo2 + 1, L))
                mn = int(rng.integers(0, 7)); rep = mn + (int(rng.integers(0, 4)) if rng.random() < 0.3 else 0)
                if pods + rep > 60:
                    mn = rep = 1
                pods += rep
                gpu = int(rng.choice([0, 1, 2, 4, 8]))
                cliques.append(dict(cpu=int(rng.choice([0, 500, 2000, 16000])), mem=int(rng.choice([0, 1024, 65536])), gpu=gpu,
                                    min=mn, replicas=rep, level=clevel, preferred=preferred(clevel),
                                    class_mask=int(rng.choice([0xFFFF, 0x1, 0x6, 0x8]))))
            scopes.append((slevel, cliques, preferred(slevel)))
        base = None
        if gi > 0 and rng.random() < (0.3 if not big else 0.05):
            base = int(rng.integers(0, gi)) if not big else int(rng.integers(max(0, gi - 50), gi))
        b.add_gang(scopes, level=glevel, priority=int(rng.integers(0, 3)), anchor=None if rng.random() < 0.5 else int(rng.integers(0, n)),
                   base=base, gated=bool(rng.random() < 0.05), preferred=preferred(glevel))
    return nodes, L, b.build()


@pytest.mark.parametrize("block", range(8))
def test_random_snapshots_match_the_oracle(built_lib, oracle, block):
    from grove_b200.engine import PlacementEngine
    for seed in range(block * 40, block * 40 + 40):
        nodes, L, (g, c, s) = random_case(seed)
        ref = oracle.run_cycle(nodes, L, g, c, s, threads=1)
        with PlacementEngine(L) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s)
            e.run_cycle()
            assert np.array_equal(e.debug_perm(), ref["perm"]), seed
            assert np.array_equal(e.gang_status(), ref["status"]), seed
            assert np.array_equal(e.placements(), ref["placements"]), seed
            assert np.array_equal(e.nodes(), ref["nodes_after"]), seed


def test_random_snapshots_one_gang_at_a_time(built_lib, oracle):
    """window = 1: the engine itself runs the sequential pass (one evaluation per round)"""
    from grove_b200.engine import PlacementEngine
    for seed in range(1000, 1030):
        nodes, L, (g, c, s) = random_case(seed, pref=seed % 2 == 1)
        ref = oracle.run_cycle(nodes, L, g, c, s)
        with PlacementEngine(L, window=1) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s); e.run_cycle()
            assert np.array_equal(e.gang_status(), ref["status"]), seed
            assert np.array_equal(e.scope_domains(), ref["scope_status"]), seed
            assert np.array_equal(e.placements(), ref["placements"]), seed


def test_random_big_snapshots(built_lib, oracle):
    from grove_b200.engine import PlacementEngine
    for seed in range(2000, 2006):
        nodes, L, (g, c, s) = random_case(seed, big=True)
        ref = oracle.run_cycle(nodes, L, g, c, s, threads=8)
        with PlacementEngine(L) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s)
            e.run_cycle()
            assert np.array_equal(e.gang_status(), ref["status"]), seed
            assert np.array_equal(e.placements(), ref["placements"]), seed
            assert np.array_equal(e.nodes(), ref["nodes_after"]), seed


@pytest.mark.parametrize("block", range(10))
def test_random_snapshots_with_preferred_levels(built_lib, oracle, block):
    """Preferred levels (podgang.go:110-117) on all three kinds of unit; includes gangs with a Preferred but no
    Required level, whose last candidate is the whole cluster packed by the scalar evaluator."""
    from grove_b200.engine import PlacementEngine
    seen = 0
    for seed in range(3000 + block * 40, 3000 + block * 40 + 40):
        nodes, L, (g, c, s) = random_case(seed, pref=True)
        seen += int((g["preferred"] != T.LEVEL_NONE).sum() + (s["preferred1"] != 0).sum() + ((c["scope"] >> 5) != 0).sum())
        ref = oracle.run_cycle(nodes, L, g, c, s, threads=1)
        with PlacementEngine(L) as e:
            e.load_nodes(nodes); e.submit_gangs(g, c, s)
            e.run_cycle()
            assert np.array_equal(e.gang_status(), ref["status"]), seed
            assert np.array_equal(e.scope_domains(), ref["scope_status"]), seed
            assert np.array_equal(e.placements(), ref["placements"]), seed
            assert np.array_equal(e.nodes(), ref["nodes_after"]), seed
    assert seen > 100


def test_random_big_snapshots_with_preferred_levels(built_lib, oracle):
    from grove_b200.engine import PlacementEngine
    for seed in range(4000, 4008):
        nodes, L, (g, c, s) = random_case(seed, big=True, pref=True)
        ref = oracle.run_cycle(nodes, L, g, c, s, threads=8)
        for W in (0, 300):
            with PlacementEngine(L, window=W) as e:
                e.load_nodes(nodes); e.submit_gangs(g, c, s)
                e.run_cycle()
                assert np.array_equal(e.gang_status(), ref["status"]), seed
                assert np.array_equal(e.scope_domains(), ref["scope_status"]), seed
                assert np.array_equal(e.placements(), ref["placements"]), seed
                assert np.array_equal(e.nodes(), ref["nodes_after"]), seed
