#!/bin/bash
mkdir -p gpurun_out/r02
( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r02/pytest_gpu_11.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r02/pytest_gpu_11.log | tail -5
python bench.py --steps 20 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_11.json 2> gpurun_out/r02/bench_11.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02/bench_11.json') if l.startswith('{')][-1])
print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), {k:(round(v['us_per_step'],1), round(v['frac'],3)) for k,v in d['roofline']['streaming'].items()})
PY
for wl in C2 C1; do python bench.py --workload $wl --steps 50 --warmup 5 --no-predict --no-cpu-baseline > gpurun_out/r02/bench_11_$wl.json 2>/dev/null; python tools/show_bench.py gpurun_out/r02/bench_11_$wl.json | cut -c1-120; done
