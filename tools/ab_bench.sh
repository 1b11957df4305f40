#!/bin/bash
# Development aid: A/B the library variants in gpurun_variants/ on one box (3 rounds each).
for r in 1 2 3; do
for v in gpurun_variants/*.so; do
  echo -n "$v  "
  NSS_B200_LIB=$PWD/$v python tools/quick_bench.py ${1:-2e7} 2>&1 | grep "events/s"
done; done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
