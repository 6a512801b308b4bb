#!/bin/bash
# round 2, session d: the whole GPU suite + the default bench line
set -u
out=gpurun_out/${1:-r2d}
mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $out/pytest.log 2>&1
tail -15 $out/pytest.log
python bench.py > $out/bench.json 2> $out/bench.err
tail -3 $out/bench.err
python - $out/bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d.get("e2e",{}).get("value"))
print("parity",d.get("parity")); print("mixture",d.get("mixture")); p=d.get("pipeline")
if p: print("pipeline",{k:p[k] for k in p if k not in ("workload",)})
PY
