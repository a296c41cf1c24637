#!/bin/bash
# round-2 GPU check 1: full GPU suite + default bench line
mkdir -p gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/smi.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02/pytest_gpu_1.log 2>&1
tail -5 gpurun_out/r02/pytest_gpu_1.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02/bench_1.json 2> gpurun_out/r02/bench_1.err
tail -c 600 gpurun_out/r02/bench_1.err
python tools/show_bench.py gpurun_out/r02/bench_1.json 2>/dev/null | head -40
