#!/bin/bash
# Runs on the GPU box (gpurun): row N3 evidence - plain timings of k_ik_track, then one ncu --set full capture of the
# same command (after it exited 0 without ncu).
set -u
mkdir -p gpurun_out
for n in 1024 16384 131072; do python tools/ik_case.py $n || exit 1; done > gpurun_out/ik_plain.log 2>&1
cat gpurun_out/ik_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_ik_track -c 1 -s 1 -f -o gpurun_out/prof_r1_ik_16384 \
    python tools/ik_case.py 16384 > gpurun_out/ncu_ik.log 2>&1
tail -2 gpurun_out/ncu_ik.log
