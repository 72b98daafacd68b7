#!/usr/bin/env python
"""Accuracy vs throughput of the centred fp16 mode's PARTIAL hi/lo weight split (run under gpurun):
for each (split_layers, lo_n, lo_kgroups): error against the fp32 mode on N positions and device-resident
throughput at 65,536 positions per call, a short burst and a sustained (power-capped) run.
Writes gpurun_out/partial_split[_seedS].json.   CONFIGS="14:64:1,14:96:1" overrides the list."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic

n = int(os.environ.get("N", "100000"))
seed = int(os.environ.get("WSEED", "1"))
sustain_s = float(os.environ.get("SUSTAIN", "1.5"))
w = W.synthetic("trained_like", seed)
sq, bl = synthetic.random_positions(n, seed=20260601)
rep = {"positions": n, "weights_seed": seed}

with wann_b200.PolicyEvaluator(w, mode="fp32") as ev:
    p32, l32 = ev.eval_probs(sq, bl, return_logits=True)
    i32, _, c32 = ev.eval_topk(sq, bl)

def accuracy(ev):
    p, l = ev.eval_probs(sq, bl, return_logits=True)
    idx, pri, cnt = ev.eval_topk(sq, bl)
    d = np.abs(l - l32)
    rel = np.abs(p - p32) / p32
    l2 = np.linalg.norm(p - p32, axis=1) / np.linalg.norm(p32, axis=1)
    return {"logits_maxnorm": float(d.max() / np.abs(l32).max()), "logits_rel_rms": float(np.sqrt((d.astype(np.float64) ** 2).mean()) / l32.std()),
            "probs_rel_median": float(np.median(rel)), "probs_rowL2_median": float(np.median(l2)),
            "top1_agreement": float((idx[:, 0] == i32[:, 0]).mean()), "legal_count_equal": bool((cnt == c32).all())}

m = 65536
dev = torch.device("cuda", 0)
reps_n = (m + n - 1) // n
sq_m, bl_m = np.tile(sq, (reps_n, 1))[:m], np.tile(bl, reps_n)[:m]
pool = [(torch.from_numpy(np.roll(sq_m, 1000 * i, axis=0).copy()).to(dev),
         torch.from_numpy(np.roll(bl_m, 1000 * i).copy()).to(dev)) for i in range(4)]

def throughput(ev, reps):
    outs = [ev.eval_topk(*pool[i]) for i in range(4)]
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for r in range(reps):
        ev.eval_topk(*pool[r % 4], out=outs[r % 4])
    t1.record(); torch.cuda.synchronize()
    return m * reps / (t0.elapsed_time(t1) * 1e-3)

default = "0:64:1,14:32:1,14:48:1,14:64:1,14:96:1,14:128:1,12:64:1,10:64:1,14:64:2,8:192:3"
for cfg in os.environ.get("CONFIGS", default).split(","):
    split, lo_n, lo_kg = (int(x) for x in cfg.split(":"))
    with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        ev.set_option("lo_n", lo_n)
        ev.set_option("lo_kgroups", lo_kg)
        if split:
            ev.set_option("split_layers", split)
        r = accuracy(ev) if n > 0 else {}
        r["burst_positions_per_s"] = throughput(ev, 20)
        r["sustained_positions_per_s"] = throughput(ev, int(sustain_s * 9e6 / m))
        nl = bin(split).count("1")
        r["mma_overhead"] = nl * lo_kg / 3.0 * lo_n / 192.0 / 3.0
        rep[cfg] = r
        print("split_layers=%d lo_n=%d lo_kgroups=%d" % (split, lo_n, lo_kg), json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "partial_split%s.json" % ("" if seed == 1 else "_seed%d" % seed)), "w"), indent=1)
