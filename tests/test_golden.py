"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must
reproduce them on CPU; the CUDA path must reproduce them on the GPU box."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    return z, (z["gangs"], z["cliques"], z["scopes"])


def test_fixtures_exist():
    assert len(FIXTURES) >= 8


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_golden(oracle, path):
    z, (g, c, s) = _load(path)
    r = oracle.run_cycle(z["nodes"], int(z["n_levels"]), g, c, s, threads=2)
    assert np.array_equal(r["placements"], z["placements"])
    assert np.array_equal(r["status"], z["status"])
    assert np.array_equal(r["scope_status"], z["scope_status"])
    assert np.array_equal(r["nodes_after"], z["nodes_after"])


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_oracle_is_thread_count_invariant(oracle, threads):
    z, (g, c, s) = _load(os.path.join(HERE, "golden", "c4_small.npz"))
    r = oracle.run_cycle(z["nodes"], int(z["n_levels"]), g, c, s, threads=threads)
    assert np.array_equal(r["placements"], z["placements"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_cuda_reproduces_golden(built_lib, path):
    from grove_b200.engine import PlacementEngine
    z, (g, c, s) = _load(path)
    with PlacementEngine(int(z["n_levels"])) as e:
        e.load_nodes(z["nodes"]); e.submit_gangs(g, c, s)
        e.run_cycle()
        assert np.array_equal(e.placements(), z["placements"])
        assert np.array_equal(e.gang_status(), z["status"])
        assert np.array_equal(e.scope_domains(), z["scope_status"])
        assert np.array_equal(e.nodes(), z["nodes_after"])
