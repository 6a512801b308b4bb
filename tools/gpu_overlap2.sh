#!/bin/bash
# pipeline throughput after a host-side change: 2D reads and the length mixture, one GPU
set -u
out=gpurun_out/${1:-overlap2}
mkdir -p $out
run() {
  local name=$1 spec=$2; shift 2
  NC_TRAIN_TIMING=1 nanocall_b200/bin/nanocall-b200 --pore r73 --synth $spec -o /tmp/$name.fa --log warning --summary-json $out/$name.json "$@" 2> $out/$name.err
  echo "$name rc=$? md5=$(md5sum < /tmp/$name.fa | cut -c1-12)" | tee -a $out/summary.txt
  grep "host time" $out/$name.err | tee -a $out/summary.txt
  python - $out/$name.json <<'PY' | tee -a $out/summary.txt
import json, sys
d = json.load(open(sys.argv[1])); x = d['devices'][0]
print(f"  reads={d['reads']} read_events/s={d['read_events_per_s']:.4g} steady_s={d['steady_wall_s']:.3f} kernels_s={(x['train_kernel_ms']+x['viterbi_kernel_ms'])/1e3:.3f} "
      f"train_s={x['train_s']:.3f} (call {x['train_call_s']:.3f}, kernels {x['train_kernel_ms']/1e3:.3f}) basecall_s={x['basecall_s']:.3f} (call {x['viterbi_call_s']:.3f}) wait_s={x['wait_s']:.3f} hand_wait_s={x.get('hand_wait_s',0):.3f} batches={x['batches']}")
PY
}
run 2d_overlap 6000:1:2048:2d:5000:5000 --batch-reads 2000 --batch-mevents 64
run mix_overlap 40000:1:4096:mix --batch-reads 4096 --batch-mevents 48 --pool-gb 60
