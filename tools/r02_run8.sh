#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_shapes.py -m gpu -q -x 2>&1 | tail -3
for wl in C4 C3 C2 C1; do
  for mode in chain steps; do
    if [ $mode = steps ]; then export MOGPE_CHAIN_STEPS=1; else unset MOGPE_CHAIN_STEPS; fi
    python bench.py --workload $wl --steps 30 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_8_${wl}_$mode.json 2> gpurun_out/r02/bench_8_${wl}_$mode.err
    python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_8_${wl}_$mode.json') if l.startswith('{')][-1])
print('$wl','$mode', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))
PY
  done
done
unset MOGPE_CHAIN_STEPS
bash tools/r02_ncu_launches.sh i8step_v7 | head -8
