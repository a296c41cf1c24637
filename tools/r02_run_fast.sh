#!/bin/bash
# slice-count modes of the INT8 step: full GPU suite, accuracy log, bench lines (default and fast mode)
O=gpurun_out/r02g
mkdir -p $O
( time python -m pytest tests -m gpu -q -x ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
python tools/i8_grad_slices_accuracy.py > $O/accuracy.log 2>&1; cat $O/accuracy.log
python bench.py --fast --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4_fast.json 2> $O/bench_c4_fast.err; python tools/show_bench.py $O/bench_c4_fast.json | cut -c1-260
python bench.py --steps 50 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-260
