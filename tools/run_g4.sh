timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/predict_bench.py 1000000 1024 tf32 2>&1 | grep "predict_y +"
timeout 400 python tools/predict_bench.py 100000000 1024 tf32 2>&1 | grep "predict_y +"
timeout 600 python tools/predict_bench.py 100000000 1024 float64 2>&1 | grep "predict_y +"
