#!/bin/bash
mkdir -p gpurun_out/r02
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02/bench_${N}gpu.json 2> gpurun_out/r02/bench_${N}gpu.err
tail -c 300 gpurun_out/r02/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_${N}gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],3), d.get('dp_parity'))
PY
