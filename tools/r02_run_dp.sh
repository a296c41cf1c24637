#!/bin/bash
# round-2 GPU check: data-parallel numerics on 2 GPUs (pytest) + bench.py --gpus 2 with its dp_parity self-check
mkdir -p gpurun_out/r02
( time python -m pytest tests/test_gpu_dp.py -m gpu -x -q ) > gpurun_out/r02/pytest_dp.log 2>&1
tail -5 gpurun_out/r02/pytest_dp.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02/bench_dp2.json 2> gpurun_out/r02/bench_dp2.err
tail -c 400 gpurun_out/r02/bench_dp2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_dp2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('dp_parity'))
PY
