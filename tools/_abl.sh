for d in 0 2048; do echo "dbg=$d"; DH_DBG=$d timeout 120 python tools/prof_conv.py sep 128 32 32 576 576 5 3 10; done
for d in 0 2048; do echo "dbg=$d"; DH_DBG=$d timeout 120 python tools/prof_conv.py sep 128 16 16 288 288 5 3 10; done
for d in 0 2048; do echo "dbg=$d"; DH_DBG=$d timeout 120 python tools/prof_conv.py conv 128 32 32 576 576 1 3 10; done
