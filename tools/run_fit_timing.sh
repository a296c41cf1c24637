MOGPE_FIT_TIMING=1 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -5
