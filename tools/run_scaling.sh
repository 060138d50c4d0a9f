#!/bin/bash
# Multi-GPU lines for profiles/ (run with: gpurun --gpus N -- 'bash tools/run_scaling.sh N [tag]'):
# the single-grid slab workload (BASELINE configs[3]) at N = the GPUs of the box, for the transports
#   p2p-step-fused  whole SSPRK33 step + 9-cell exchange in ONE launch (psk_ssprk33_step_p2p)
#   p2p-step        wait / whole-step kernel / push (3 launches per step)
#   p2p             per-stage kernel with the 3-cell exchange fused (3 launches per step)
#   nccl            NCCL send/recv per stage (the baseline)
# at 2^30 cells (the named config) and at 2^27 .. 2^20 (where the exchange starts to cost).
N=${1:-2}
TAG=${2:-r2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out/${TAG}_slab_scaling_${N}gpu.jsonl
: > $OUT
run() {  # run <bench args...>
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --workload slab "$@" 2>> gpurun_out/${TAG}_slab_${N}gpu.err | tail -n 1 >> $OUT
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 \
      --master-port 29517 bench.py --gpus "$N" --workload slab "$@" 2>> gpurun_out/${TAG}_slab_${N}gpu.err | tail -n 1 >> $OUT
  fi
  tail -n 1 $OUT | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['cells'], d['transport'], 'ms/step %.4f' % d['ms_per_step'], '%.3e' % d['value'], 'parity', (d['parity'] or {}).get('max_rel'), 'launches', d['gpu_launches'])"
}
run --transport p2p-step-fused --steps 20 --warmup 5
run --transport p2p-step --steps 20 --warmup 5
run --transport p2p --steps 20 --warmup 5
for cells in 134217728 16777216 1048576; do
  for tr in p2p-step-fused p2p nccl; do
    run --transport $tr --cells $cells --steps 100 --warmup 10
  done
done
run --transport p2p-step --cells 16777216 --steps 100 --warmup 10
run --transport p2p-step --cells 1048576 --steps 100 --warmup 10
