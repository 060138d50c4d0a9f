#!/bin/bash
# One GPU-box call: A/B of the stage-kernel variants, then the GPU tests and the bench with the winner.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_gpu.txt 2>&1
timeout 240 python tools/ab_stage.py > gpurun_out/ab.log 2>&1
echo "ab exit $?" >> gpurun_out/ab.log
tail -3 gpurun_out/ab.log
V=$(cat gpurun_out/ab_stage_best.txt 2>/dev/null || echo 5000)
echo "best variant $V"
PSK_STAGE_VARIANT=$V timeout 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py tests/test_gpu_edge_cases.py tests/test_gpu_api.py -x -q > gpurun_out/tests_best.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_best.log
tail -3 gpurun_out/tests_best.log
PSK_STAGE_VARIANT=$V timeout 150 python bench.py > gpurun_out/bench_best.json 2> gpurun_out/bench_best.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench_best.json
