#!/bin/bash
# per-wave kernel times (tools/time_train_round.py) for the product library and every variants/<name>
out=gpurun_out/${1:-ttr}; shift
mkdir -p $out
for v in base $(ls variants 2>/dev/null); do
  if [ $v = base ]; then lib=$PWD/nanocall_b200/libnanocall_b200.so; else lib=$PWD/variants/$v/libnanocall_b200.so; fi
  echo "== $v" | tee -a $out/times.txt
  NC_LIB_PATH=$lib python tools/time_train_round.py "$@" 2>&1 | tee -a $out/times.txt
done
