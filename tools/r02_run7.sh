#!/bin/bash
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02/pytest_gpu_7.log 2>&1
tail -6 gpurun_out/r02/pytest_gpu_7.log
bash tools/r02_ncu_launches.sh i8step_v5 | head -16
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_7.json 2> gpurun_out/r02/bench_7.err
python tools/show_bench.py gpurun_out/r02/bench_7.json
