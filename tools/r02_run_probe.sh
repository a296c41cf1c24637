#!/bin/bash
# parity tests + per-product times of the INT8 step under the epilogue switches (128: timing print only; 32: Horner only)
O=gpurun_out/r02p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for ew in 16 8; do for d in 128 32; do echo "== EW $ew"; MOGPE_I8_EW=$ew MOGPE_I8_DBG=$d timeout 300 python tools/i8_epilogue_probe.py 2>&1 | grep -A8 "timed step" ; done; done > $O/probe.log 2>&1
cat $O/probe.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-200
