#!/bin/bash
# One GPU-box call: GPU test suite, bench line, ncu capture of the whole-step kernel, launch list of the bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 200 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench exit $?"
cut -c1-400 gpurun_out/bench_1gpu.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:step_warp_fused -c 2 -o gpurun_out/step_fused -f python tools/profile_target.py 16384 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list exit $?"
ls -la gpurun_out | head -30
