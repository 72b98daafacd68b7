#!/usr/bin/env python
"""Latency of small HOST calls -- the reference's real call pattern (evaluate([node]) from the tree,
tree_builder.pyx:189,830) -- with and without the zero-copy path, plus the 11-thread coalesced
throughput (expansion_MCTS_functions.pyx:105-119).  Development tool, run under gpurun."""
import os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

mode = os.environ.get("MODE", "bf16")
refine = float(os.environ.get("REFINE", "0"))
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), mode=mode, refine_margin=refine)
sq, bl = synthetic.random_positions(4096, seed=1)


def lat(fn, n, reps=400):
    for _ in range(20):
        fn(sq[:n], bl[:n])
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn(sq[:n], bl[:n])
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    return np.median(ts), np.percentile(ts, 10), np.percentile(ts, 99)


for n in (1, 8, 64, 256, 512):
    row = "N=%4d" % n
    for zc in (0, 1):
        ev.set_option("zero_copy", zc)
        a = lat(ev.eval_topk, n)
        b = lat(ev.eval_probs, n)
        row += " | zero_copy=%d topk %6.1f us (p10 %6.1f p99 %6.1f)  probs %6.1f us" % (zc, a[0], a[1], a[2], b[0])
    print(row, flush=True)

for zc in (0, 1):
    for co in (0, 1):
        ev.set_option("zero_copy", zc)
        ev.set_option("coalesce", co)
        stop, counts = [False], [0] * 11

        def worker(i):
            s1, b1 = sq[i:i + 1].copy(), bl[i:i + 1].copy()
            while not stop[0]:
                ev.eval_probs(s1, b1)
                counts[i] += 1
        th = [threading.Thread(target=worker, args=(i,)) for i in range(11)]
        [t.start() for t in th]
        time.sleep(0.3)
        c0, t0 = sum(counts), time.perf_counter()
        time.sleep(1.5)
        c1, t1 = sum(counts), time.perf_counter()
        stop[0] = True
        [t.join() for t in th]
        print("11 threads x batch 1, zero_copy=%d coalesce=%d: %.0f evals/s" % (zc, co, (c1 - c0) / (t1 - t0)), flush=True)
ev.close()
