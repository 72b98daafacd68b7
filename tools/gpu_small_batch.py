#!/usr/bin/env python
"""Where does the time go for small batches? (development tool)
Alternates the `small_batch_halves` option per batch size to cancel clock drift."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
sq, bl = synthetic.random_positions(4096, seed=1)
sqt, blt = torch.from_numpy(sq).cuda(), torch.from_numpy(bl).cuda()

def measure(n, reps=50):
    out = ev.eval_topk(sqt[:n], blt[:n])
    ev.set_option("time_kernels", 1); ev.kernel_times()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ev.eval_topk(sqt[:n], blt[:n], out=out)
    b.record(); torch.cuda.synchronize()
    ms, cnt = ev.kernel_times(); ev.set_option("time_kernels", 0)
    return a.elapsed_time(b) / reps * 1e3, ms / cnt * 1e3

for n in (1, 8, 64, 128, 256, 296, 400, 592, 1024):
    res = {0: [], 1: []}
    for rnd in range(3):
        for opt in (0, 1):
            ev.set_option("small_batch_halves", opt)
            res[opt].append(measure(n))
    f = lambda xs: "gpu %.1f us conv %.1f us" % (min(x[0] for x in xs), min(x[1] for x in xs))
    print("N=%5d  full-tiles: %s | one-half: %s" % (n, f(res[0]), f(res[1])), flush=True)
