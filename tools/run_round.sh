#!/bin/bash
# One GPU-box call: GPU test suite, then the bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 200 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
echo "bench exit $?"
cut -c1-200 gpurun_out/bench_1gpu.json
timeout 60 python bench.py --workload slab --transport p2p-step --cells 134217728 --steps 10 > gpurun_out/bench_slab_step_1gpu.json 2> gpurun_out/bench_slab_step_1gpu.err
echo "slab exit $?"; cut -c1-200 gpurun_out/bench_slab_step_1gpu.json
