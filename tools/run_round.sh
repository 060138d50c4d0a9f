#!/bin/bash
# One GPU-box call: whole-step shapes A/B, then compute-sanitizer over the new kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 python tools/ab_stage.py > gpurun_out/ab.log 2>&1
echo "ab exit $?" >> gpurun_out/ab.log
tail -10 gpurun_out/ab.log
timeout 150 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_step.py > gpurun_out/sanitizer_step_memcheck.txt 2>&1
echo "memcheck exit $?"; tail -3 gpurun_out/sanitizer_step_memcheck.txt
