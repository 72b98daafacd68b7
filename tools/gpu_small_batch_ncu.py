#!/usr/bin/env python
"""The small-batch path under ncu (development tool): evaluations of 1, 256 and 1,024 positions in the
headline arithmetic (fp16c as created), device pointers, ranked child priors out.  The profiled range is
opened with cudaProfilerStart only after the context is calibrated and warm, so that

  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/<name> \\
      python tools/gpu_small_batch_ncu.py

captures exactly one conv_stack_tc launch per batch size (in that order).  Without ncu it prints the
CUDA-event time of each size (that run, not the ncu one, is the timing)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

SIZES = (1, 256, 1024)
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), mode="fp16c")
sq, bl = synthetic.random_positions(max(SIZES), seed=1)
sqt, blt = torch.from_numpy(sq).cuda(), torch.from_numpy(bl).cuda()
outs = {}
for n in SIZES:                                   # warm: lanes, workspaces, instruction cache
    outs[n] = ev.eval_topk(sqt[:n], blt[:n])
    for _ in range(5):
        ev.eval_topk(sqt[:n], blt[:n], out=outs[n])
torch.cuda.synchronize()
for n in SIZES:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        ev.eval_topk(sqt[:n], blt[:n], out=outs[n])
    b.record()
    torch.cuda.synchronize()
    print("N=%5d  %.1f us per call back to back" % (n, a.elapsed_time(b) / 50 * 1e3), flush=True)
torch.cuda.cudart().cudaProfilerStart()
for n in SIZES:
    ev.eval_topk(sqt[:n], blt[:n], out=outs[n])
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
idx = outs[1][0] if isinstance(outs[1], (tuple, list)) else outs[1]
print("done", np.asarray(idx.cpu() if hasattr(idx, "cpu") else idx).ravel()[:4], flush=True)
ev.close()
