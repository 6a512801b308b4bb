#!/bin/bash
# usage: bash tools/gpu_sweep.sh <tag> "<gpu counts>" [mix_reads] [2d_reads]
set -u
tag=$1; counts=$2; nmix=${3:-1000000}; n2d=${4:-100000}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $out/gpus.txt
lscpu | grep -E "Model name|^CPU\(s\)" > $out/cpu.txt
bash tools/sweep.sh $out/sweep_mix.json $nmix:1:4096:mix "$counts" --batch-reads 4096 --batch-mevents 48 --pool-gb 110 2> $out/sweep_mix.log
cat $out/sweep_mix.log
bash tools/sweep.sh $out/sweep_2d.json $n2d:1:2048:2d:5000:5000 "$counts" --batch-reads 2000 2> $out/sweep_2d.log
cat $out/sweep_2d.log
