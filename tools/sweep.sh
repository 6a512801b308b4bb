#!/bin/bash
# Read-sharded throughput sweep through the host program (BASELINE.json configs[4]): the same synthetic read stream on
# 1..N GPUs of one box, dynamic dispatch from one shared queue, ordered output.
# usage: bash tools/sweep.sh <out.json> <synth spec> "<gpu counts>" [extra nanocall-b200 options]
# e.g.   bash tools/sweep.sh gpurun_out/sweep_1M.json 1000000:1:4096:mix "1 2 4 8" --pool-gb 100
set -u
out=$1; spec=$2; counts=$3; shift 3
tmp=$(mktemp -d)
echo "[" > $out
first=1
for n in $counts; do
  s=$(date +%s.%N)
  nanocall_b200/bin/nanocall-b200 --pore r73 --synth $spec --gpus $n -o /dev/null --log warning --summary-json $tmp/s$n.json "$@" 2> $tmp/err$n.txt
  rc=$?
  e=$(date +%s.%N)
  if [ $rc -ne 0 ]; then echo "run with $n GPUs failed:" >&2; tail -5 $tmp/err$n.txt >&2; continue; fi
  [ $first -eq 1 ] || echo "," >> $out
  first=0
  python - $tmp/s$n.json $s $e "$spec" "$*" >> $out <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
d["process_wall_s"] = float(sys.argv[3]) - float(sys.argv[2])
d["synth"] = sys.argv[4]
d["options"] = sys.argv[5]
print(json.dumps(d))
PY
  python - $tmp/s$n.json <<'PY' >&2
import json, sys
d = json.load(open(sys.argv[1]))
print(f"gpus={d['n_gpus']} reads={d['reads']} read_events/s={d['read_events_per_s']:.4g} steady_s={d['steady_wall_s']:.2f} tail_s={d['tail_s']:.2f}",
      "busy:", [round(x['train_s'] + x['basecall_s'], 1) for x in d['devices']], "wait:", [round(x['wait_s'], 1) for x in d['devices']])
PY
done
echo "]" >> $out
rm -rf $tmp
