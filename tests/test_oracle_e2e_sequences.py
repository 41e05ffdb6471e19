"""The reference's multi-step gang-scheduling scenarios (gang_scheduling_test.go GS2-GS12), replayed step by step
against the oracle: the pod counts the suites wait for (WaitForPodPhases / WaitForRunningPods) after every
cordon / uncordon / scale step.  One pod per node (80 MiB pods on 150 MiB nodes), as in the suites."""
import pytest

from e2e_sim import E2ESim
from grove_b200 import synth


def oracle_place(oracle):
    return lambda nodes, g, c, s: oracle.run_cycle(nodes, synth.E2E_LEVELS, g, c, s)


def replay(place, script):
    """script: (n_nodes, cordoned, workload, [(action, arg, expected running pods, expected total pods), ...])"""
    n, cordoned, wl, steps = script
    sim = E2ESim(n, cordoned)
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
