#!/bin/bash
# bench-configuration evidence: launch list of the default bench command, one ncu --set full capture of viterbi_alpha_kernel on
# the 10000 x 10000 launch (dram bytes = the `traffic` of the roofline), pipeline per-kernel times of the current build
set -u
out=gpurun_out/${1:-r2g}
mkdir -p $out
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipe_summary.txt | tr ' ' '\n' | grep -E "_ms" | tr '\n' ' '; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pipeline-reads 500 --mix-reads 1000 > $out/launches_bench.log 2>&1
python tools/launch_summary.py $out/launches_bench.csv | tee $out/launches_bench_summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:viterbi_alpha -c 1 -f -o $out/vit_alpha_10kx10k \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --pipeline-reads 0 --mix-reads 0 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
rm -f $out/pipe.fa $out/pipe.err $out/pipe_stats.tsv
ls -la $out
