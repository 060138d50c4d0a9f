#!/bin/bash
# One GPU-box call: GPU test suite, then the adjoint bench lines (configs[4]).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1
echo "tests exit $?" >> gpurun_out/tests_gpu.log
tail -4 gpurun_out/tests_gpu.log
timeout 200 python bench.py --workload adjoint --steps 1000 > gpurun_out/bench_adjoint_1000.json 2> gpurun_out/bench_adjoint_1000.err
echo "adjoint 1000 exit $?"; cut -c1-300 gpurun_out/bench_adjoint_1000.json
timeout 200 python bench.py --workload adjoint --steps 200 > gpurun_out/bench_adjoint_200.json 2> gpurun_out/bench_adjoint_200.err
echo "adjoint 200 exit $?"; cut -c1-300 gpurun_out/bench_adjoint_200.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
