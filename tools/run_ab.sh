#!/bin/bash
# One GPU-box call: interleaved A/B of the stage-kernel variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/ab_stage.py > gpurun_out/ab.log 2>&1
echo "ab exit $?" >> gpurun_out/ab.log
tail -14 gpurun_out/ab.log
