#!/bin/bash
# one ncu --set full capture of viterbi_alpha_kernel on the bench's 10000 x 10000 launch (dram bytes = the roofline's `traffic`,
# executed warp instructions = roofline_issue's numerator)
set -u
out=gpurun_out/${1:-ncu_vit}
mkdir -p $out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:viterbi_alpha -c 1 -f -o $out/vit_alpha_10kx10k \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --pipeline-reads 0 --mix-reads 0 > $out/ncu_full.log 2>&1
tail -2 $out/ncu_full.log
