#!/bin/bash
# Development aid: A/B the library variants in $VAR (gpurun_variants/) on one box (3 rounds each).
VAR=${VAR:-$PWD/gpurun_variants}
for r in 1 2 3; do
for v in $VAR/*.so; do
  echo -n "$(basename $v)  "
  NSS_B200_LIB=$v python tools/quick_bench.py ${1:-2e7} 2>&1 | grep "events/s"
done; done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
