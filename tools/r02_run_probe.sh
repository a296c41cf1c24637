#!/bin/bash
# parity tests + per-product times of the INT8 step under the epilogue switches (128: timing print only; 64: plain Horner
# recurrence; 32: Horner only, nothing stored; 96: both)
O=gpurun_out/r02p
mkdir -p $O
python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for d in 128 64 32 96; do MOGPE_I8_DBG=$d python tools/i8_epilogue_probe.py 2>&1 | grep -A8 "timed step" ; done > $O/probe.log 2>&1
cat $O/probe.log
python bench.py --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-200
