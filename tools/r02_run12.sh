#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_engine.py tests/test_gpu_bench_shapes.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 30 --warmup 5 --no-predict --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"; done
bash tools/r02_ncu_launches.sh i8step_v10 | grep -E "chol|kuf8|one step"
