#!/usr/bin/env python
"""Experiment: offsets of the centred mode restricted to few mantissa bits (WANN_B200_OFFSET_BITS) -- accuracy on
20,000 positions and sustained throughput / clock at 65,536 positions per launch.  Development tool."""
import json, os, subprocess, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic
w = W.synthetic("trained_like", 1)
n = 20000
sq, bl = synthetic.random_positions(n, seed=20260601)
with wann_b200.PolicyEvaluator(w, mode="fp32") as ev:
    p32, l32 = ev.eval_probs(sq, bl, return_logits=True)
dev = torch.device("cuda", 0)
B = 65536
bsq = torch.from_numpy(np.tile(sq, (4, 1))[:B]).to(dev); bbl = torch.from_numpy(np.tile(bl, 4)[:B]).to(dev)
for bits in [int(x) for x in os.environ.get("ZERO_BUCKETS", "0,2,3,4,5,6").split(",")]:
    os.environ["WANN_B200_ZERO_BUCKETS"] = str(bits)
    with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        ev.set_option("split_layers", int(os.environ.get("SPLIT", "0")))
        p, l = ev.eval_probs(sq, bl, return_logits=True)
        d = np.abs(l - l32)
        acc = (d.max() / np.abs(l32).max(), np.sqrt((d.astype(np.float64) ** 2).mean()) / l32.std(), np.median(np.abs(p - p32) / p32))
        for _ in range(5):
            ev.eval_topk(bsq, bbl)
        torch.cuda.synchronize()
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                               stdout=subprocess.PIPE, text=True)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        steps = 350
        for _ in range(steps):
            ev.eval_topk(bsq, bbl)
        b.record(); torch.cuda.synchronize()
        smi.terminate()
        rows = [r.split(",") for r in smi.stdout.read().strip().splitlines()]
        clk = np.median([float(r[0]) for r in rows[3:]]) if len(rows) > 4 else -1
        pw = np.median([float(r[1]) for r in rows[3:]]) if len(rows) > 4 else -1
        print("zero buckets %s: logits maxnorm %.2e relrms %.2e pmed %.2e | sustained %.2f M pos/s over %.1f s, %.0f MHz, %.0f W | offsets %s" % (
            bits, acc[0], acc[1], acc[2], B * steps / (a.elapsed_time(b) * 1e-3) / 1e6, a.elapsed_time(b) * 1e-3, clk, pw,
            np.round(ev.centering()[0][1], 4).tolist()), flush=True)
