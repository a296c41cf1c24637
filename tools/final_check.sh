# final verification of the round: full GPU suite, smoke, default bench line (C5 at full size: tools/predict_bench.py)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err; python tools/show_bench.py gpurun_out/bench_default.json
