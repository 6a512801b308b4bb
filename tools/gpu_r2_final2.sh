#!/bin/bash
# end-of-round evidence (final state of round 2): the whole GPU suite, the default bench line, the reference arm, the launch
# list of the bench command, the length mixture on 20000 reads
set -u
out=gpurun_out/${1:-r2final3}
mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
python bench.py > $out/bench.json 2> $out/bench.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2>> $out/bench.err
python -c "
import json
d=json.load(open('$out/bench.json')); p=d['pipeline']
print('value',d['value'],'e2e',d['e2e']['value'],'mixture',d['mixture']['value'],'parity',d['parity']['identical'],'/',d['parity']['reads_checked'], d['clocks'])
print('pipeline read_events_per_s',p['read_events_per_s'],'fwbw_events_per_s',p['fwbw_events_per_s'],p['kernel_ms'],p['roofline']['frac'])
r=json.load(open('$out/bench_reference.json')); print('reference arm', r.get('value'), r.get('unit'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pipeline-reads 500 --mix-reads 1000 > $out/launches_bench.log 2>&1
python tools/launch_summary.py $out/launches_bench.csv > $out/launches_bench_summary.txt; tail -12 $out/launches_bench_summary.txt
python bench.py --mix --reads 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --pipeline-reads 0 --mix-reads 0 > $out/bench_mix20k.json 2> $out/bench_mix20k.err
python -c "
import json
d=json.load(open('$out/bench_mix20k.json')); print('mixture 20000 reads', d['value'], d['ms_per_step'], d['config'].get('workload'))
"
