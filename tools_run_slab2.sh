set -x
timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q 2>&1 | grep -E "rank|passed|failed|Error" | tail -25
for cells in 1048576 16777216; do
for t in p2p p2p-serial nccl; do
  timeout 300 python bench.py --workload slab --gpus 2 --transport $t --steps 300 --cells $cells > gpurun_out/slab2_${t}_${cells}.json 2> gpurun_out/slab2_${t}_${cells}.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/slab2_${t}_${cells}.json').read().strip().splitlines()[-1])
print('$t', $cells, d['value'], d['ms_per_step'])"
done
done
