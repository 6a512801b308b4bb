#!/bin/bash
# round 2, session c: ncu --set full of the three training kernels (one launch each, 256-read pipeline)
set -u
out=gpurun_out/${1:-r2c}
mkdir -p $out
python tools/make_synth_ncev.py /tmp/pipe_small.ncev 256 5000 5000 7 > /dev/null
for k in fwbw_kernel pm_stats_kernel st_stats_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/$k \
      nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p.fa --log warning /tmp/pipe_small.ncev > $out/$k.log 2>&1
  tail -1 $out/$k.log
done
ls -la $out
