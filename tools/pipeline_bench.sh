#!/bin/bash
# BASELINE.json configs[2]/[3] on one GPU through the CLI: synthetic 2D reads (template + complement, random per-read
# scaling), Forward/Backward training rounds, candidate-model selection (r73.t x {r73.c.p1, r73.c.p2}), Viterbi with
# the trained parameters.  Prints the CLI's per-GPU summary and derived events/s.
# usage: bash tools/pipeline_bench.sh <out_dir> [n_reads=1000] [n_template=5000] [n_complement=5000]
set -e
out=${1:-gpurun_out/pipeline}; n=${2:-1000}; nt=${3:-5000}; nc=${4:-5000}
mkdir -p $out
python tools/make_synth_ncev.py /tmp/pipe.ncev $n $nt $nc 5 > /dev/null
/usr/bin/env time -v true 2>/dev/null || true
s=$(date +%s.%N)
nanocall_b200/bin/nanocall-b200 --pore r73 -o $out/pipe.fa --stats $out/pipe_stats.tsv --log info /tmp/pipe.ncev 2> $out/pipe.err
e=$(date +%s.%N)
grep "^gpu " $out/pipe.err | tail -1 > $out/pipe_summary.txt
python - $out/pipe_summary.txt $s $e $n $nt $nc <<'PY'
import re, sys, json
line = open(sys.argv[1]).read()
kv = dict(re.findall(r"(\w+)=([\d.]+)", line))
wall = float(sys.argv[3]) - float(sys.argv[2])
n, nt, nc = map(int, sys.argv[4:7])
fb_ev, fb_ms = float(kv["fwbw_events"]), float(kv["train_kernel_ms"])
v_ev, v_ms = float(kv["viterbi_events"]), float(kv["viterbi_kernel_ms"])
print(json.dumps({"workload": f"{n} reads x ({nt} template + {nc} complement events), training on, 2 candidate model pairs",
                  "train_rounds": int(float(kv["train_rounds"])), "fwbw_events": fb_ev, "train_kernel_ms": fb_ms,
                  "fwbw_events_per_s_kernel": fb_ev / fb_ms * 1e3 if fb_ms else None,
                  "viterbi_events": v_ev, "viterbi_kernel_ms": v_ms,
                  "viterbi_events_per_s_kernel": v_ev / v_ms * 1e3 if v_ms else None,
                  "cli_wall_s": wall, "basecalled_events_per_s_wall": n * (nt + nc) / wall}))
PY
