#!/bin/bash
# Runs on the GPU box (gpurun): both bench arms, then the ncu evidence profiles/README.md describes.
# Every ncu pass follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err || exit 1
python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/ncu_bench.log 2>&1
for cfg in "float64 4096 100 rows:f64_4096" "float64 131072 2:f64_131072" "float32 131072 2:f32_131072"; do
  args="${cfg%%:*}"; tag="${cfg##*:}"
  python tools/prof_case.py $args > gpurun_out/plain_$tag.log 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 -s 1 -f -o gpurun_out/prof_r1_final_$tag \
      python tools/prof_case.py $args > gpurun_out/ncu_$tag.log 2>&1
done
tail -c 600 gpurun_out/bench_default.json; echo; cat gpurun_out/bench_reference.json | cut -c1-400
