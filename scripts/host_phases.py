"""Wall-clock of the host-facing calls of one e2e cycle (GROVE_DEBUG_HOST=1 adds the engine's own split)."""
import sys, time
sys.path.insert(0, ".")
from grove_b200 import synth
from grove_b200.engine import PlacementEngine

c = synth.config_c4(); g, cl, sc = c["tables"]; nodes = c["nodes"]
with PlacementEngine(4) as e:
    for it in range(5):
        t0 = time.perf_counter(); e.load_nodes(nodes); t1 = time.perf_counter(); e.submit_gangs(g, cl, sc); t2 = time.perf_counter()
        st = e.run_cycle(); t3 = time.perf_counter(); pl = e.placements(copy=False); gs = e.gang_status(copy=False); t4 = time.perf_counter()
        print("load %.2f submit %.2f cycle %.2f (dev %.2f) get %.2f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), st["ms_total"], 1e3 * (t4 - t3)))
