#!/bin/bash
mkdir -p gpurun_out/r02
( time python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -m gpu -q -x -k "c4 or int8" ) > gpurun_out/r02/pytest_i8step.log 2>&1
tail -8 gpurun_out/r02/pytest_i8step.log
for d in 64 32; do MOGPE_I8_DBG=$d python tools/i8_epilogue_probe.py 2>&1 | grep -A7 "timed step" ; done | tee gpurun_out/r02/i8_epilogue_probe2.log
python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_4.json 2> gpurun_out/r02/bench_4.err
python tools/show_bench.py gpurun_out/r02/bench_4.json
