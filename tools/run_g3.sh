timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in C1 C2 C4; do timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r01k_$w.json 2> gpurun_out/bench_r01k_$w.err; tail -1 gpurun_out/bench_r01k_$w.err; python tools/show_bench.py gpurun_out/bench_r01k_$w.json; done
