#!/usr/bin/env python
"""Bring-up / accuracy-vs-throughput report of the centred fp16 mode (run under gpurun):
per-layer agreement with the emulation of its roundings, error against the fp32 mode for each
split_layers mask, and device-resident throughput.  Writes gpurun_out/fp16c_report.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic
from oracle import oracle as O

n = int(os.environ.get("N", "100000"))
seed = int(os.environ.get("WSEED", "1"))
w = W.synthetic("trained_like", seed)
sq, bl = synthetic.random_positions(n, seed=20260601)
rep = {"positions": n, "weights_seed": seed}

def nerr(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())

with wann_b200.PolicyEvaluator(w, mode="fp32") as ev:
    p32, l32 = ev.eval_probs(sq, bl, return_logits=True)
    i32, _, c32 = ev.eval_topk(sq, bl)
rep["logit_sigma"] = float(l32.std())

def accuracy(ev):
    p, l = ev.eval_probs(sq, bl, return_logits=True)
    idx, pri, cnt = ev.eval_topk(sq, bl)
    d = np.abs(l - l32)
    rel = np.abs(p - p32) / p32
    l2 = np.linalg.norm(p - p32, axis=1) / np.linalg.norm(p32, axis=1)
    return {"logits_maxnorm": float(d.max() / np.abs(l32).max()), "logits_rel_rms": float(np.sqrt((d ** 2).mean()) / l32.std()),
            "probs_rel_median": float(np.median(rel)), "probs_rel_p99": float(np.quantile(rel, 0.99)),
            "probs_rowL2_median": float(np.median(l2)), "probs_rowL2_max": float(l2.max()),
            "probs_max_abs": float(np.abs(p - p32).max()),
            "top1_agreement": float((idx[:, 0] == i32[:, 0]).mean()), "legal_count_equal": bool((cnt == c32).all())}

def throughput(ev, m=65536, reps=20):
    dev = torch.device("cuda", 0)
    reps_sq = np.tile(sq, ((m + n - 1) // n, 1))[:m]; reps_bl = np.tile(bl, (m + n - 1) // n)[:m]
    dsq, dbl = torch.from_numpy(reps_sq).to(dev), torch.from_numpy(reps_bl).to(dev)
    for _ in range(3):
        ev.eval_topk(dsq, dbl)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        ev.eval_topk(dsq, dbl)
    t1.record(); torch.cuda.synchronize()
    return m * reps / (t0.elapsed_time(t1) * 1e-3)

# 1. emulation check (256 positions), plain centred mode and with split layers
m = 256
planes = O.encode_planes(sq[:m], bl[:m])
for split in (0, 14):
    with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        if split:
            ev.set_option("split_layers", split)
        off, pos, kap = ev.centering()
        lg_e, pr_e, acts_e = O.forward_fp16c_emulated(planes, w, off, pos, split_mask=split, return_all=True)
        r = {"offsets": off.tolist(), "kappa": kap.tolist()}
        for layer in (1, 2, 3, 4, 5):
            a = ev.debug_activations(sq[:m], bl[:m], layer)
            r["layer%d_vs_emulation" % layer] = nerr(a, acts_e[layer - 1])
        pr, lg = ev.eval_probs(sq[:m], bl[:m], return_logits=True)
        r["logits_vs_emulation"] = nerr(lg, lg_e)
        r["emulation_vs_fp32_logits"] = nerr(lg_e, l32[:m])
        rep["emulation_split%d" % split] = r
        print("emulation split=%d" % split, json.dumps({k: v for k, v in r.items() if k not in ("offsets", "kappa")}), flush=True)

# 2. accuracy + throughput per configuration
for name, mode, split, margin in (("fp16", "fp16", 0, 0.0), ("fp16c", "fp16c", 0, 0.0), ("fp16c+refine0.02", "fp16c", 0, 0.02),
                                  ("fp16c+refine0.01", "fp16c", 0, 0.01),
                                  ("fp16c+split_layers8", "fp16c", 8, 0.0), ("fp16c+split_layers12", "fp16c", 12, 0.0),
                                  ("fp16c+split_layers14", "fp16c", 14, 0.0), ("fp16c+split_layers14+refine0.01", "fp16c", 14, 0.01), ("bf16", "bf16", 0, 0.0)):
    with wann_b200.PolicyEvaluator(w, mode=mode, refine_margin=margin) as ev:
        if split:
            ev.set_option("split_layers", split)
        r = accuracy(ev)
        r["refined_fraction"] = ev.stats()["refined_positions"] / (2.0 * n)
        r["positions_per_s_65536"] = throughput(ev)
        r["positions_per_s_1024"] = throughput(ev, 1024, 200)
        rep[name] = r
        print(name, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "fp16c_report%s.json" % ("" if seed == 1 else "_seed%d" % seed)), "w"), indent=1)
