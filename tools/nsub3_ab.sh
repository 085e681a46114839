mkdir -p gpurun_out/r2v
for v in 0 1; do
  DH_NSUB3=$v DH_SAVE=gpurun_out/r2v/out$v.npy timeout 60 python tools/prof_conv.py sep 256 32 32 576 576 5 3 10
  DH_NSUB3=$v timeout 60 python tools/prof_conv.py sep 256 32 32 576 576 5 3 10
  DH_NSUB3=$v DH_RES2=1 timeout 60 python tools/prof_conv.py sep 256 32 32 576 576 5 3 10
  DH_NSUB3=$v timeout 60 python tools/prof_conv.py sep 256 32 32 384 576 3 3 10
done
python -c "
import numpy as np
a=np.load('gpurun_out/r2v/out0.npy'); b=np.load('gpurun_out/r2v/out1.npy'); print('max diff', np.abs(a-b).max(), 'equal', np.array_equal(a,b))"
rm -f gpurun_out/r2v/*.npy
