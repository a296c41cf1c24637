for p in tf32 float64; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/predict_bench.py 100000000 1024 $p 2>&1 | grep "predict_y +"
done
