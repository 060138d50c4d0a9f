set -x
N=${1:-8}
timeout 400 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | grep -E "passed|failed|Error|error" | tail -5
for cells in 1073741824 16777216; do
for t in p2p nccl; do
  steps=300; if [ $cells -gt 100000000 ]; then steps=30; fi
  timeout 300 python bench.py --workload slab --gpus $N --transport $t --steps $steps --cells $cells > gpurun_out/slab${N}_${t}_${cells}.json 2> gpurun_out/slab${N}_${t}_${cells}.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/slab${N}_${t}_${cells}.json').read().strip().splitlines()[-1])
print('RESULT $N $t', $cells, d['value'], d['ms_per_step'])" || tail -5 gpurun_out/slab${N}_${t}_${cells}.err
done
done
