#!/bin/bash
# round 2, session a: parity tests + pipeline timing with the new Forward/Backward + trainer kernels
set -u
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
lscpu | head -20 > $out/cpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
tail -5 $out/pytest.log
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipeline.json
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 64 5000 5000 7 > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches_pipeline.csv \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/launches_pipeline.log 2>&1
python tools/launch_summary.py $out/launches_pipeline.csv | tee $out/launches_pipeline_summary.txt
rm -f $out/pipe.fa
ls -la $out
