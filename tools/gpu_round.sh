#!/bin/bash
# One GPU session: parity tests, the bench line, the ncu launch list and one full capture of the dominant kernel,
# the pipeline (configs 3/4) through the CLI and one full capture of the Forward/Backward kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
lscpu | head -20 > $out/cpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
tail -3 $out/pytest.log
python bench.py > $out/bench.json 2> $out/bench.err
cat $out/bench.json
python bench.py --impl reference --steps 1 --warmup 0 > $out/bench_ref.json 2>> $out/bench.err
cat $out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --reads 1500 --no-cpu-baseline --no-e2e > $out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:viterbi_alpha -c 1 -f -o $out/vit_alpha \
    python bench.py --reads 1480 --events 3000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
python bench.py --mix --reads 4000 --no-cpu-baseline > $out/bench_mix.json 2>> $out/bench.err
cut -c1-160 $out/bench_mix.json
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipeline.json
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 64 5000 5000 7 > /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/launches_pipeline.csv \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/launches_pipeline.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwbw_kernel -s 2 -c 1 -f -o $out/fwbw \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/ncu_fwbw.log 2>&1
tail -2 $out/ncu_fwbw.log
ls -la $out
