# same-box A/B: the products of one C4 step with the default epilogue warp counts and with eight everywhere
for rep in 1 2; do for ew in 0 8; do echo "MOGPE_I8_EW=$ew"; MOGPE_I8_EW=$ew MOGPE_I8_DBG=128 timeout 300 python tools/i8_epilogue_probe.py 2>&1 | grep -E "K=512"; done; done
