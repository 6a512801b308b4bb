#!/bin/bash
# F/B parity tests, pipeline timing with per-kernel device times, and one ncu --set full capture of fwbw_kernel
set -u
out=gpurun_out/${1:-r2e}
mkdir -p $out
( time timeout 900 python -m pytest tests/test_fwbw_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q ) > $out/pytest.log 2>&1
grep -E "passed|failed" $out/pytest.log
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipe_summary.txt
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 256 5000 5000 7 > /dev/null
for k in ${2:-fwbw_kernel}; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/$k \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/$k.log 2>&1
done
rm -f $out/pipe.fa $out/pipe.err $out/pipe_stats.tsv
