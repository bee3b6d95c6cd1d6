#!/bin/bash
# Round-2 ncu captures (run under gpurun).  Each ncu pass follows a plain run of the same command that exited 0.
# usage: tools/capture_r2.sh "<dtype> <n> <T> [rows]:<tag>" ...
set -u
mkdir -p gpurun_out
for cfg in "$@"; do
  args="${cfg%%:*}"; tag="${cfg##*:}"
  python tools/prof_case.py $args > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed: $tag"; cat gpurun_out/plain_$tag.log; continue; }
  ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -s 1 -f -o gpurun_out/prof_r2_$tag \
      python tools/prof_case.py $args > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
done
