#!/bin/bash
mkdir -p gpurun_out/r02
( time python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -m gpu -q -k "c4 or int8" ) > gpurun_out/r02/pytest_i8step.log 2>&1
tail -8 gpurun_out/r02/pytest_i8step.log
bash tools/r02_ncu_launches.sh i8step_v2 | head -16
MOGPE_KGR_ROWS=2 bash tools/r02_ncu_launches.sh i8step_v2_kgr2 | grep "kuf_grad"
python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_3.json 2> gpurun_out/r02/bench_3.err
python tools/show_bench.py gpurun_out/r02/bench_3.json
