import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import test_gpu_tc as T
from gpu_util import Dev
dev = Dev(torch)
case = eval(sys.argv[1])
try:
    T._run_sepconv(dev, case, 3, 2)
    print('OK', case)
except Exception as e:
    print('FAIL', case, str(e)[:200])
