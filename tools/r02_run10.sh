#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_shapes.py tests/test_gpu_int8.py -m gpu -q -x -k "c4 or int8" 2>&1 | tail -3
MOGPE_I8_DBG=64 python tools/i8_epilogue_probe.py 2>&1 | grep -A7 "timed step"
python bench.py --workload C4 --steps 30 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_10.json 2> gpurun_out/r02/bench_10.err
python tools/show_bench.py gpurun_out/r02/bench_10.json | cut -c1-200
