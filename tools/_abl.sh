timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu 2>&1 | tail -3
for d in 0 64 27; do echo "dbg=$d"; DH_DBG=$d timeout 120 python tools/prof_conv.py sep 128 32 32 576 576 5 3 10; done
timeout 120 python tools/prof_conv.py sep 128 32 32 288 288 5 3 10
timeout 120 python tools/prof_conv.py sep 128 16 16 288 288 5 3 10
timeout 120 python tools/prof_conv.py sep 128 8 8 288 288 5 3 10
