#!/bin/bash
# bench.py under torchrun exactly as the driver launches it (N ranks, one per GPU), then the reference arm
N=${1:-2}
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
tail -5 gpurun_out/scale_n$N.err
python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/scale_n$N.json').read().strip().splitlines() if x.startswith('{')][-1])
print('N', l['n_gpus'], 'value', round(l['value']), 'ms/step', round(l['ms_per_step'],4), 'frac', round(l['roofline']['frac'],3), 'e2e', round(l['e2e']['value']))
for k,v in l.get('configs',{}).items():
    print(k, v.get('error') or (v['n_gpus'], round(v['value']), round(v['ms_per_step'],3), round(v['roofline']['frac'],3), v['config'].get('sequences_per_gpu')))
PY
