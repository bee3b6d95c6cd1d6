#!/bin/bash
# Round-2 evidence (run under gpurun): GPU tests, both bench arms, the ncu launch list of the bench command and full captures
# of the dominant kernels.  Every ncu pass follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_final.log
tail -3 gpurun_out/r2_gputests_final.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err || { echo bench failed; tail -5 gpurun_out/r2_bench_default.err; exit 1; }
python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err || exit 1
python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r2_ncu_bench.log 2>&1
tools/capture_r2.sh "float64 4096 100 rows:f64_4096" "float64 131072 2:f64_131072" "float32 131072 2:f32_131072"
python tools/contact_perf.py > gpurun_out/r2_contact_perf.log 2>&1
python tools/mpc_frame_probe.py 131072 50 > gpurun_out/r2_mpc_frame.log 2>&1
cut -c1-1500 gpurun_out/r2_bench_default.json; echo; cut -c1-600 gpurun_out/r2_bench_reference.json; cat gpurun_out/r2_contact_perf.log gpurun_out/r2_mpc_frame.log
