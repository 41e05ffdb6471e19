"""The reference's multi-step gang-scheduling scenarios (gang_scheduling_test.go GS2-GS12), replayed step by step
against the oracle: the pod counts the suites wait for (WaitForPodPhases / WaitForRunningPods) after every
cordon / uncordon / scale step.  One pod per node (80 MiB pods on 150 MiB nodes), as in the suites."""
import pytest

from e2e_sim import E2ESim
from grove_b200 import synth


def oracle_place(oracle):
    return lambda nodes, g, c, s: oracle.run_cycle(nodes, synth.E2E_LEVELS, g, c, s)


def replay(place, script, split_surplus=False, after_step=None, same_pass_unlock=True):
    """script: (n_nodes, cordoned, workload, [(action, arg, expected running pods, expected total pods), ...])"""
    n, cordoned, wl, steps = script
    sim = E2ESim(n, cordoned, split_surplus=split_surplus, same_pass_unlock=same_pass_unlock)
    sim.deploy(wl)
    trace = []
    for action, arg, running, total in steps:
        if action == "uncordon":
            sim.uncordon(arg)
        elif action == "pcsg":
            sim.scale_pcsg(*arg)
        elif action == "pcs":
            sim.scale_pcs(arg)
        sim.step(place)
        trace.append((sim.running(), sim.pods()))
        assert (sim.running(), sim.pods()) == (running, total), (action, arg, trace)
        if after_step:
            after_step(len(trace), sim)
    return sim


# scenario tables: the numbered steps of the test headers (file:line), as (action, argument, running, total pods)
GS = {
    "GS2": (14, 5, 1, [("none", 0, 0, 10), ("uncordon", 1, 10, 10), ("pcsg", (0, 3), 10, 14), ("uncordon", 4, 14, 14)]),                     # :76-85
    "GS3": (20, 11, 1, [("none", 0, 0, 10), ("uncordon", 1, 10, 10), ("pcs", 2, 10, 20), ("uncordon", 10, 20, 20)]),                         # :146-154
    "GS4": (28, 19, 1, [("none", 0, 0, 10), ("uncordon", 1, 10, 10), ("pcsg", (0, 3), 10, 14), ("uncordon", 4, 14, 14), ("pcs", 2, 14, 24),
                        ("pcsg", (1, 3), 14, 28), ("uncordon", 14, 28, 28)]),                                                              # :207-219
    "GS5": (10, 8, 2, [("none", 0, 0, 10), ("uncordon", 1, 3, 10), ("uncordon", 7, 10, 10)]),                                               # :277-285
    "GS6": (14, 12, 2, [("none", 0, 0, 10), ("uncordon", 1, 3, 10), ("uncordon", 7, 10, 10), ("pcsg", (0, 3), 10, 14), ("uncordon", 2, 12, 14),
                        ("uncordon", 2, 14, 14)]),                                                                                         # :339-353
    "GS7": (14, 12, 2, [("none", 0, 0, 10), ("uncordon", 1, 3, 10), ("uncordon", 2, 5, 10), ("uncordon", 5, 10, 10), ("pcsg", (0, 3), 10, 14),
                        ("uncordon", 2, 12, 14), ("uncordon", 2, 14, 14)]),                                                                # :449-465
    "GS8": (14, 12, 2, [("none", 0, 0, 10), ("pcsg", (0, 3), 0, 14), ("uncordon", 1, 3, 14), ("uncordon", 4, 7, 14), ("uncordon", 7, 14, 14)]),   # :575-586
    "GS9": (20, 18, 2, [("none", 0, 0, 10), ("uncordon", 1, 3, 10), ("uncordon", 7, 10, 10), ("pcs", 2, 10, 20), ("uncordon", 3, 13, 20),
                        ("uncordon", 7, 20, 20)]),                                                                                         # :669-681
    "GS10": (20, 18, 2, [("none", 0, 0, 10), ("pcs", 2, 0, 20), ("uncordon", 4, 6, 20), ("uncordon", 4, 10, 20), ("uncordon", 10, 20, 20)]),      # :772-784
    "GS11": (28, 26, 2, [("none", 0, 0, 10), ("uncordon", 1, 3, 10), ("uncordon", 7, 10, 10), ("pcsg", (0, 3), 10, 14), ("uncordon", 2, 12, 14),
                         ("uncordon", 2, 14, 14), ("pcs", 2, 14, 24), ("uncordon", 3, 17, 24), ("uncordon", 7, 24, 24), ("pcsg", (1, 3), 24, 28),
                         ("uncordon", 2, 26, 28), ("uncordon", 2, 28, 28)]),                                                               # :866-887
    "GS12": (28, 26, 2, [("none", 0, 0, 10), ("pcs", 2, 0, 20), ("pcsg", (0, 3), 0, 24), ("pcsg", (1, 3), 0, 28), ("uncordon", 4, 6, 28),
                         ("uncordon", 8, 14, 28), ("uncordon", 14, 28, 28)]),                                                              # :1016-1028
}


@pytest.mark.parametrize("name", sorted(GS, key=lambda k: int(k[2:])))
def test_gang_scheduling_sequences(oracle, name):
    sim = replay(oracle_place(oracle), GS[name])
    # final state of every suite: all pods Running on distinct nodes (ListPodsAndAssertDistinctNodes)
    assert sim.running() == sim.pods()
    assert all(g.scheduled for g in sim.gangs)


@pytest.mark.parametrize("name", sorted(GS, key=lambda k: int(k[2:])))
def test_gang_scheduling_sequences_minimums_first(oracle, name):
    """the same scenarios with the surplus of unscheduled PodGangs submitted as gated remainder rows, and with scaled
    PodGangs joining only the pass after their base was scheduled (operator latency): same counts"""
    for unlock in (True, False):
        sim = replay(oracle_place(oracle), GS[name], split_surplus=True, same_pass_unlock=unlock)
        assert sim.running() == sim.pods() and all(g.scheduled for g in sim.gangs)


def test_minimums_of_every_gang_before_anybodys_surplus(oracle):
    """What the step descriptions say beyond the counts.  GS8 step 8 (:583): four nodes appear while sg-x-1 and sg-x-2
    wait -> "pcs-0-{sg-x-1-pc-b=1, sg-x-1-pc-c=1}, pcs-0-{sg-x-2-pc-b=1, sg-x-2-pc-c=1}": each scaled gang gets its
    two minimum pods, neither takes four.  GS10 step 6 (:779): six nodes, two PCS replicas -> "pcs-0-{pc-a=1,
    sg-x-0-pc-b=1, sg-x-0-pc-c=1}, pcs-1-{...}": three each.  GS12 step 8 (:1023) likewise.  Scaled PodGangs join the
    pass after their base was scheduled, as they do behind the operator's gate removal."""
    by = lambda sim: {g.name: g.n_bound() for g in sim.gangs}

    def gs8(step, sim):
        if step == 4:
            assert by(sim) == {"pcs-0": 3, "pcs-0-sg-x-1": 2, "pcs-0-sg-x-2": 2}
    replay(oracle_place(oracle), GS["GS8"], split_surplus=True, same_pass_unlock=False, after_step=gs8)

    def gs10(step, sim):
        if step == 3:
            assert by(sim) == {"pcs-0": 3, "pcs-0-sg-x-1": 0, "pcs-1": 3, "pcs-1-sg-x-1": 0}
        if step == 4:
            assert by(sim) == {"pcs-0": 3, "pcs-0-sg-x-1": 2, "pcs-1": 3, "pcs-1-sg-x-1": 2}
    replay(oracle_place(oracle), GS["GS10"], split_surplus=True, same_pass_unlock=False, after_step=gs10)

    def gs12(step, sim):
        if step == 5:
            b = by(sim)
            assert b["pcs-0"] == 3 and b["pcs-1"] == 3 and sum(b.values()) == 6
        if step == 6:
            assert all(v >= 2 for k, v in by(sim).items() if "sg-x" in k)   # "remaining PCSG pods": every scaled gang has its minimum
    replay(oracle_place(oracle), GS["GS12"], split_surplus=True, same_pass_unlock=False, after_step=gs12)
