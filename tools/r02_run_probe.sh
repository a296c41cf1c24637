#!/bin/bash
# parity tests + per-product times of the INT8 step (MOGPE_I8_DBG=128: timing print only) + launch list
O=gpurun_out/r02p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
MOGPE_I8_DBG=128 timeout 300 python tools/i8_epilogue_probe.py 2>&1 | grep -A8 "timed step" > $O/probe.log 2>&1
cat $O/probe.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-predict --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt; head -12 $O/launches_summary.txt
