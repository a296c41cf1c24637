#!/bin/bash
mkdir -p gpurun_out/r02
for d in 64 1 2 3 4 8 12 16 32; do MOGPE_I8_DBG=$d python tools/i8_epilogue_probe.py 2>&1 | grep -A7 "timed step" ; done | tee gpurun_out/r02/i8_epilogue_probe3.log | grep -E "timed|batch=16 kmode=1|batch=8 kmode=2"
