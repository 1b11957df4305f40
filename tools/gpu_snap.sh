#!/bin/bash
# Development aid: freeze the working tree into gpurun_variants/snap_<tag>/ (git-ignored, travels to
# the GPU box) and run a command from that frozen copy under gpurun, retrying while the pod is busy.
# The working tree can be edited while the call waits in the queue.
#   [GPUS=2|4|8] tools/gpu_snap.sh <tag> <timeout_s> '<command run inside the frozen copy>'
# Inside the command, $OUT is the box's gpurun_out/ (merged back) and $VAR is gpurun_variants/.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TAG=$1; TO=$2; CMD=$3
SNAP=$ROOT/gpurun_variants/snap_$TAG
mkdir -p "$SNAP"
rm -rf "$SNAP"; mkdir -p "$SNAP"
(cd "$ROOT" && tar cf - --exclude=./.git --exclude=./gpurun_out --exclude=./gpurun_variants \
      --exclude=__pycache__ --exclude=.pytest_cache . ) | (cd "$SNAP" && tar xf -)
LOG=/tmp/gpu_$TAG.log
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$TO" -- \
    "export OUT=\$PWD/gpurun_out VAR=\$PWD/gpurun_variants; mkdir -p \$OUT; cd gpurun_variants/snap_$TAG && ( $CMD )" > "$LOG" 2>&1 || true
  if grep -q "status=transient" "$LOG"; then sleep 60; continue; fi
  break
done
echo "done after $i tries" >> "$LOG"
rm -rf "$SNAP"
