#!/bin/bash
# one ncu --set full capture (with source) of kernel $2 (launch index $3) during the CLI on 256 synthetic 2D reads
set -u
out=gpurun_out/${1:-ncu1}
k=${2:-st_stats_kernel}
skip=${3:-2}
mkdir -p $out
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 256 5000 5000 7 > /dev/null
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $out/$k \
    nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/$k.log 2>&1
tail -3 $out/$k.log
