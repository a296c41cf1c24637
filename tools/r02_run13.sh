#!/bin/bash
mkdir -p gpurun_out/r02
python tools/i8_grad_slices_accuracy.py 2>&1 | tee gpurun_out/r02/i8_grad_slices_accuracy.log | tail -10
python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_shapes.py tests/test_gpu_int8.py -m gpu -q -k "c4 or int8" 2>&1 | tail -4
MOGPE_I8_DBG=64 python tools/i8_epilogue_probe.py 2>&1 | grep -A7 "timed step"
python bench.py --steps 30 --warmup 5 --no-predict --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['roofline']['achieved'], d['roofline']['frac'])"
