#!/bin/bash
# The CLI's read sharding over N GPUs (one host thread + context per GPU, FASTA in input order): same bytes as one GPU.
# usage (under gpurun --gpus N): bash tools/cli_multi_gpu_check.sh N
n=${1:-2}
mkdir -p gpurun_out
python tools/make_synth_ncev.py /tmp/mg.ncev 600 3000 3000 9 > /dev/null
s=$(date +%s.%N); nanocall_b200/bin/nanocall-b200 --pore r73 -o /tmp/mg1.fa --log info /tmp/mg.ncev 2> /tmp/mg1.err; e=$(date +%s.%N)
s2=$(date +%s.%N); nanocall_b200/bin/nanocall-b200 --pore r73 --gpus $n -o /tmp/mgN.fa --log info /tmp/mg.ncev 2> /tmp/mgN.err; e2=$(date +%s.%N)
{
  echo "reads=600 x (3000+3000) events, full pipeline"
  echo "1 gpu: $(echo "$e - $s" | bc -l 2>/dev/null || python -c "print($e-$s)") s; $(grep '^gpu ' /tmp/mg1.err | tr '\n' ' ')"
  echo "$n gpus: $(python -c "print($e2-$s2)") s; $(grep '^gpu ' /tmp/mgN.err | tr '\n' ' ')"
  if cmp -s /tmp/mg1.fa /tmp/mgN.fa; then echo "FASTA identical ($(wc -c < /tmp/mg1.fa) bytes, $(grep -c '>' /tmp/mg1.fa) sequences)"; else echo "FASTA DIFFERS"; fi
} | tee gpurun_out/cli_multi_gpu_$n.txt
