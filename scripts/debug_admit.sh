GROVE_DEBUG_ADMIT=1 python - <<'PY'
import sys; sys.path.insert(0,'.')
from grove_b200 import synth
from grove_b200.engine import PlacementEngine
c=synth.config_c4(); g,cl,sc=c['tables']
with PlacementEngine(4) as e:
    e.load_nodes(c['nodes']); e.submit_gangs(g,cl,sc); e.run_cycle(); print(e.run_cycle() if False else '')
PY
