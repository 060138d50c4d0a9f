set -x
timeout 200 python -m pytest tests/test_gpu_distributed.py -q -k multi_gpu 2>&1 | grep -E "passed|failed|Error|error" | tail -5
for cells in 1048576 16777216; do
  timeout 100 python bench.py --workload slab --gpus 2 --transport p2p --graph --steps 300 --cells $cells > gpurun_out/slab2_p2p-graph_${cells}.json 2> gpurun_out/slab2_p2p-graph_${cells}.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/slab2_p2p-graph_${cells}.json').read().strip().splitlines()[-1])
print('RESULT p2p graph', $cells, d['value'], d['ms_per_step'], d['config']['finite'])" || grep -v "^\*\|OMP" gpurun_out/slab2_p2p-graph_${cells}.err | head -12
done
