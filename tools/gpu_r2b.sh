#!/bin/bash
# round 2, session b: F/B parity tests + pipeline timing with per-kernel device times
set -u
tag=${1:-r2b}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 900 python -m pytest tests/test_fwbw_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
bash tools/pipeline_bench.sh $out 1000 5000 5000 > $out/pipeline.json 2> $out/pipeline.err
cat $out/pipeline.json; cat $out/pipe_summary.txt
rm -f $out/pipe.fa $out/pipe.err $out/pipe_stats.tsv
