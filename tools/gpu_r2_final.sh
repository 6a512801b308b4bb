#!/bin/bash
# end-of-round evidence: the whole GPU suite, the default bench line, the reference arm, the launch list of the bench
# command, ncu --set full captures of the three rewritten training kernels
set -u
out=gpurun_out/${1:-r2final}
mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
python bench.py > $out/bench.json 2> $out/bench.err
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_reference.json 2>> $out/bench.err
python -c "
import json
d=json.load(open('$out/bench.json')); p=d['pipeline']
print('value',d['value'],'e2e',d['e2e']['value'],'mixture',d['mixture']['value'],'parity',d['parity']['identical'],'/',d['parity']['reads_checked'])
print('pipeline read_events_per_s',p['read_events_per_s'],'fwbw_events_per_s',p['fwbw_events_per_s'],p['kernel_ms'],p['roofline']['frac'])
r=json.load(open('$out/bench_reference.json')); print('reference arm', r.get('value'), r.get('unit'))
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pipeline-reads 500 --mix-reads 1000 > $out/launches_bench.log 2>&1
python tools/launch_summary.py $out/launches_bench.csv > $out/launches_bench_summary.txt; tail -12 $out/launches_bench_summary.txt
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 256 5000 5000 7 > /dev/null
for k in st_stats_kernel fwbw_kernel pm_stats_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/$k \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/$k.log 2>&1
tail -1 $out/$k.log
done
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipe_summary.txt
rm -f $out/pipe.fa $out/pipe.err $out/pipe_stats.tsv
