python - <<'PY'
import sys, time; sys.path.insert(0,'.')
import numpy as np
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
c=synth.config_c4(); g,cl,sc=c['tables']; nodes=c['nodes']
with PlacementEngine(4) as e:
    for it in range(4):
        t0=time.perf_counter(); e.load_nodes(nodes); t1=time.perf_counter(); e.submit_gangs(g,cl,sc); t2=time.perf_counter()
        st=e.run_cycle(); t3=time.perf_counter(); pl=e.placements(); gs=e.gang_status(); t4=time.perf_counter()
        print(f"load {1e3*(t1-t0):.2f} submit {1e3*(t2-t1):.2f} cycle {1e3*(t3-t2):.2f} (dev {st['ms_total']:.2f}) get {1e3*(t4-t3):.2f} ms  rounds {st['rounds']} launches {st['kernel_launches']}")
PY
