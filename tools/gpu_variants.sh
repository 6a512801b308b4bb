#!/bin/bash
# A/B timing of library variants (variants/<name>/libnanocall_b200.so) on the 1000-read pipeline
set -u
out=gpurun_out/${1:-variants}
mkdir -p $out
python tools/make_synth_ncev.py /tmp/pipe.ncev 1000 5000 5000 5 > /dev/null
for v in base $(ls variants); do
  if [ $v = base ]; then lp=""; else lp=$PWD/variants/$v; fi
  LD_LIBRARY_PATH=$lp nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/p_$v.fa --log info /tmp/pipe.ncev 2> /tmp/err_$v.txt
  echo "$v: $(grep '^gpu ' /tmp/err_$v.txt | tr ' ' '\n' | grep -E 'fwbw_ms|pm_stats_ms|st_stats_ms|train_kernel_ms' | tr '\n' ' ') md5=$(md5sum < /tmp/p_$v.fa | cut -c1-8)" | tee -a $out/variants.txt
done
