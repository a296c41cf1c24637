#!/bin/bash
mkdir -p gpurun_out/r02
bash tools/r02_ncu_launches.sh i8step_v4 | head -14
MOGPE_KGS=1 bash tools/r02_ncu_launches.sh i8step_v4_kgs | head -8
python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_6.json 2> gpurun_out/r02/bench_6.err
python tools/show_bench.py gpurun_out/r02/bench_6.json
MOGPE_KGS=1 python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_6k.json 2> gpurun_out/r02/bench_6k.err
python tools/show_bench.py gpurun_out/r02/bench_6k.json
