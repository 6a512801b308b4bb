#!/bin/bash
# usage (under gpurun): bash tools/run_variants.sh <n> <n> ...   -- vit_diag.py for each experiment build
for n in "$@"; do
  echo "== variant $n"
  if [ "$n" = "0" ]; then python tools/vit_diag.py --reads 1480 --reps 1 2>&1 | grep mode
  else NC_LIB_PATH=$PWD/nanocall_b200/variants/libnc_exp$n.so python tools/vit_diag.py --reads 1480 --reps 1 2>&1 | grep mode; fi
done
