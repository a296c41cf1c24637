#!/bin/bash
mkdir -p gpurun_out/r02
( time python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py -m gpu -q -k "c4 or int8" ) > gpurun_out/r02/pytest_i8step.log 2>&1
tail -30 gpurun_out/r02/pytest_i8step.log
python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_2.json 2> gpurun_out/r02/bench_2.err
tail -c 600 gpurun_out/r02/bench_2.err
python tools/show_bench.py gpurun_out/r02/bench_2.json
