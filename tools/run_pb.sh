for M in 1024 512; do for p in float64 tf32; do timeout 300 python tools/predict_bench.py 1000000 $M $p 2>&1 | grep "predict_y +"; done; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_predict_tf32_r01.csv python tools/predict_bench.py 262144 1024 tf32 > gpurun_out/ncu_pb.log 2>&1
tail -2 gpurun_out/ncu_pb.log
