# refresh the single-GPU bench lines of every workload (profiles/bench_r01_<workload>_1gpu.json)
for w in ${WORKLOADS:-C1 C2 C3 C4}; do
  timeout 400 python bench.py --workload $w --steps 100 --warmup 5 $( [ $w != C4 ] && echo --no-predict ) > gpurun_out/bench_final_$w.json 2> gpurun_out/bench_final_$w.err
  python tools/show_bench.py gpurun_out/bench_final_$w.json
done
