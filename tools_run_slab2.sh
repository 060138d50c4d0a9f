set -x
timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | grep -E "rank|passed|failed|Error|error" | tail -30
for cells in 1048576 16777216 1073741824; do
for t in p2p p2p-serial nccl; do
  steps=300; if [ $cells -gt 100000000 ]; then steps=20; fi
  timeout 300 python bench.py --workload slab --gpus 2 --transport $t --steps $steps --cells $cells > gpurun_out/slab2_${t}_${cells}.json 2> gpurun_out/slab2_${t}_${cells}.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/slab2_${t}_${cells}.json').read().strip().splitlines()[-1])
print('RESULT $t', $cells, d['value'], d['ms_per_step'])" || tail -5 gpurun_out/slab2_${t}_${cells}.err
done
done
