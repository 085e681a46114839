for v in 0 1 2 3; do echo "variant $v"; DH_DBG=$v DH_RES2=1 timeout 120 python tools/prof_conv.py conv 128 32 32 48 576 1 3 10; done
