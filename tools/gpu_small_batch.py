#!/usr/bin/env python
"""Where does the time go for small batches? (development tool)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
sq, bl = synthetic.random_positions(4096, seed=1)
sq, bl = np.tile(sq, (16, 1)), np.tile(bl, 16)
sqt, blt = torch.from_numpy(sq).cuda(), torch.from_numpy(bl).cuda()
for pdl in (0,):
  ev.set_option("pdl", pdl); print("pdl", pdl)
  for n in (1, 8, 64, 256, 592, 1024, 4096, 65536):
      out = ev.eval_topk(sqt[:n], blt[:n])
      ev.set_option("time_kernels", 1); ev.kernel_times()
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      reps = 50
      t0 = time.perf_counter(); a.record()
      for _ in range(reps):
          ev.eval_topk(sqt[:n], blt[:n], out=out)
      b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
      ms, cnt = ev.kernel_times(); ev.set_option("time_kernels", 0)
      print("N=%5d  gpu %.1f us/call  conv kernel %.1f us  host enqueue %.1f us/call" % (
          n, a.elapsed_time(b) / reps * 1e3, ms / cnt * 1e3, (t1 - t0) / reps * 1e6), flush=True)
