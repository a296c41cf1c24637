#!/bin/bash
# quick loop: bench-shape + INT8 engine tests, C4 bench line, serialised launch list
O=gpurun_out/r02h
mkdir -p $O
python -m pytest tests/test_gpu_bench_shapes.py tests/test_gpu_engine.py tests/test_gpu_int8.py -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python bench.py --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-predict --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt; head -14 $O/launches_summary.txt
