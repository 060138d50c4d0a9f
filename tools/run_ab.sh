#!/bin/bash
# One GPU-box call: the GPU test suite on the current default, then the A/B of the stage-kernel variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 240 python tools/ab_stage.py > gpurun_out/ab.log 2>&1
echo "ab exit $?" >> gpurun_out/ab.log
tail -2 gpurun_out/ab.log
