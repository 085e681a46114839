timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -3
echo nores; DH_NORES=1 timeout 120 python tools/prof_conv.py conv 128 32 32 48 576 1 3 10
echo res1; timeout 120 python tools/prof_conv.py conv 128 32 32 48 576 1 3 10
echo res2; DH_RES2=1 timeout 120 python tools/prof_conv.py conv 128 32 32 48 576 1 3 10
for d in 0 27; do echo "dbg=$d"; DH_DBG=$d timeout 120 python tools/prof_conv.py sep 128 32 32 576 576 5 3 10; done
timeout 120 python tools/prof_conv.py conv 128 32 32 576 576 1 3 10
