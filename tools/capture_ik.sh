set -u
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; tail -3 gpurun_out/gputests.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 900 gpurun_out/bench_default.json; echo
for n in 1024 16384 131072; do python tools/ik_case.py $n; done > gpurun_out/ik_plain.log 2>&1; cat gpurun_out/ik_plain.log
ncu --set full --clock-control none --import-source on -k regex:k_ik_track -c 1 -s 1 -f -o gpurun_out/prof_r1_ik_16384 python tools/ik_case.py 16384 > gpurun_out/ncu_ik.log 2>&1; tail -2 gpurun_out/ncu_ik.log
