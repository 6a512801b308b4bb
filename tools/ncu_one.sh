#!/bin/bash
# usage: bash tools/ncu_one.sh <kernel-regex> <out-name> [n_reads=256]  -- one ncu --set full capture from the pipeline CLI
k=$1; name=$2; n=${3:-256}
mkdir -p gpurun_out/ncu
python tools/make_synth_ncev.py /tmp/pipe_small.ncev $n 5000 5000 7 > /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/ncu/$name \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > gpurun_out/ncu/$name.log 2>&1
tail -1 gpurun_out/ncu/$name.log
