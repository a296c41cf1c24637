#!/bin/bash
# round-2 final check on one B200: full GPU suite, smoke, default bench line (+ reference arm), the other workloads, strong-scaling
# N=1 point, fast training mode, slice-count accuracy log, launch list and one `ncu --set full` capture of the INT8 products
O=gpurun_out/r02f
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
( time python -m pytest tests -m gpu -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --steps 20 --warmup 5 > $O/bench_c4.json 2> $O/bench_c4.err; python tools/show_bench.py $O/bench_c4.json | cut -c1-260
python bench.py --steps 100 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4_100.json 2> $O/bench_c4_100.err; python tools/show_bench.py $O/bench_c4_100.json | cut -c1-160
python bench.py --fast --steps 100 --warmup 5 --no-predict --no-cpu-baseline > $O/bench_c4_fast.json 2> $O/bench_c4_fast.err; python tools/show_bench.py $O/bench_c4_fast.json | cut -c1-160
for wl in C1 C2 C3; do python bench.py --workload $wl --steps 50 --warmup 5 --no-predict > $O/bench_$wl.json 2> $O/bench_$wl.err; python tools/show_bench.py $O/bench_$wl.json | cut -c1-200; done
python bench.py --scaling strong --steps 10 --warmup 3 --no-predict --no-cpu-baseline > $O/bench_c4_strong1.json 2> $O/bench_c4_strong1.err; python tools/show_bench.py $O/bench_c4_strong1.json | cut -c1-160
python tools/i8_grad_slices_accuracy.py > $O/accuracy.log 2>&1; tail -5 $O/accuracy.log
MOGPE_I8_DBG=128 python tools/i8_epilogue_probe.py 2>&1 | grep -A8 "timed step" > $O/probe.log; cat $O/probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-predict --no-cpu-baseline > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt; head -12 $O/launches_summary.txt
ncu --set full --import-source on --clock-control none -k regex:gemm_i8_sliced --launch-skip 6 --launch-count 6 -o $O/full_i8 -f python tools/one_step.py C4 3 > $O/ncu_full.log 2>&1
ls -la $O/full_i8.ncu-rep
