#!/usr/bin/env python
"""Driver for an ncu launch list of the byte kernels (see gpu_byte_kernels.py): a few calls of each.
N positions (argv[1], default 4 Mi: 1.3 GB of encode traffic, far beyond the 126 MB L2, so that the DRAM
counters see the writes too); the profiler range opens after the context is calibrated and warm:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/<name>.csv python tools/gpu_byte_kernels_ncu.py"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import _lib as L, weights as W, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
sq, bl = synthetic.random_positions(4096, seed=9)
dev = torch.device("cuda", 0)
sq_t = torch.from_numpy(np.tile(sq, (N // 4096, 1))).to(dev)
bl_t = torch.from_numpy(np.tile(bl, N // 4096)).to(dev)
planes = torch.empty((N, 8, 8, 4), dtype=torch.int8, device=dev)
mask = torch.empty((N, 155), dtype=torch.uint8, device=dev)
s1 = ctypes.c_void_p(1)
L.check(ev._lib.wann_encode(ev._ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, planes.data_ptr(), L.MEM_DEVICE, s1))
L.check(ev._lib.wann_legal_mask(ev._ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, mask.data_ptr(), L.MEM_DEVICE, s1))
ev.eval_expand_seeded(sq_t[:65536], bl_t[:65536])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(3):
    L.check(ev._lib.wann_encode(ev._ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, planes.data_ptr(), L.MEM_DEVICE, s1))
    L.check(ev._lib.wann_legal_mask(ev._ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, mask.data_ptr(), L.MEM_DEVICE, s1))
    ev.eval_expand_seeded(sq_t[:65536], bl_t[:65536])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done', N, flush=True)
