#!/bin/bash
# Multi-GPU lines for profiles/ (run with: gpurun --gpus N -- 'bash tools/run_scaling.sh N'):
# the ensemble bench (weak scaling, whole-step kernel) and the single-grid slab workload with the
# per-stage fused exchange (p2p) and the whole-step exchange (p2p-step), N = the GPUs of the box.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() {  # run <tag> <bench args...>
  local tag=$1; shift
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29517 bench.py --gpus "$N" "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  fi
  echo "$tag exit $?"; tail -n 1 gpurun_out/$tag.json | cut -c1-220
}
run bench_ensemble_${N}gpu --steps 20 --warmup 3
run bench_slab_p2p_${N}gpu --workload slab --transport p2p --steps 10 --warmup 3
run bench_slab_step_${N}gpu --workload slab --transport p2p-step --steps 10 --warmup 3
run bench_slab_step_small_${N}gpu --workload slab --transport p2p-step --cells 16777216 --steps 50 --warmup 5
run bench_slab_p2p_small_${N}gpu --workload slab --transport p2p --cells 16777216 --steps 50 --warmup 5
