#!/bin/bash
# 2-GPU session: whole GPU suite, then the small sweeps (dispatcher, host overhead)
set -u
out=gpurun_out/${1:-r2f}
mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -q -x ) > $out/pytest.log 2>&1
tail -4 $out/pytest.log
bash tools/sweep.sh $out/sweep_2d_20k.json 20000:1:2048:2d:5000:5000 "1 2" --batch-reads 2000 2> $out/sweep_2d_20k.log
cat $out/sweep_2d_20k.log
bash tools/sweep.sh $out/sweep_mix_100k.json 100000:1:4096:mix "1 2" --batch-reads 4096 --batch-mevents 48 --pool-gb 60 2> $out/sweep_mix_100k.log
cat $out/sweep_mix_100k.log
