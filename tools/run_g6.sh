timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/predict_bench.py 100000000 1024 int8 2>&1 | grep "predict_y +"
timeout 300 python tools/predict_bench.py 4000000 512 int8 2>&1 | grep "predict_y +"
timeout 300 python tools/predict_bench.py 4000000 512 float64 2>&1 | grep "predict_y +"
