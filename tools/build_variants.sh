#!/bin/bash
# Timing-experiment builds of the library: nanocall_b200/variants/libnc_exp<N>.so (loaded through NC_LIB_PATH).
set -e
cd "$(dirname "$0")/.."
mkdir -p nanocall_b200/variants
for n in "$@"; do
  make -s -C nanocall_b200/csrc BUILD=$PWD/nanocall_b200/csrc/build_exp$n OUT=$PWD/nanocall_b200/variants/libnc_exp$n.so NC_EXTRA_NVFLAGS=-DNC_EXP=$n 2>&1 | grep -E "error|alpha.*Used|spill" | tail -3
done
ls -la nanocall_b200/variants
