#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_shapes.py tests/test_gpu_int8.py -m gpu -q -x 2>&1 | tail -3
for mode in chain steps; do
  if [ $mode = steps ]; then export MOGPE_CHAIN_STEPS=1; else unset MOGPE_CHAIN_STEPS; fi
  python bench.py --workload C4 --steps 30 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_9_$mode.json 2> gpurun_out/r02/bench_9_$mode.err
  python tools/show_bench.py gpurun_out/r02/bench_9_$mode.json | cut -c1-140
done
unset MOGPE_CHAIN_STEPS
bash tools/r02_ncu_launches.sh i8step_v8 | head -12
