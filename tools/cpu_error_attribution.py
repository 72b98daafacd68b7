#!/usr/bin/env python
"""Per-layer / per-operand attribution of the fp16 tensor-core mode's error (CPU, torch float64
arithmetic with selective fp16 roundings).  Sources: w1..w5 = conv weights of layer l rounded to
fp16, a1..a4 = stored activations of layer l rounded to fp16 (the input of layer l+1).
Usage: python tools/cpu_error_attribution.py [n_positions]"""
import itertools, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wann_b200 import weights as W, synthetic
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
fmt = os.environ.get("FMT", "fp16")
sq, bl = synthetic.random_positions(n, seed=20260601)
w = W.synthetic("trained_like", int(os.environ.get("WSEED", "1")))
planes = torch.from_numpy(O.encode_planes(sq, bl).astype(np.float64)).permute(0, 3, 1, 2)

def rnd(t):
    if fmt == "fp16":
        return t.to(torch.float16).to(torch.float64)
    return t.to(torch.bfloat16).to(torch.float64)

def forward(sources):
    x = planes
    for i in range(1, 6):
        k = torch.from_numpy(w["hidden_layer/%d/weights" % i].astype(np.float64)).permute(3, 2, 0, 1)
        if "w%d" % i in sources:
            k = rnd(k.float())
        b = torch.from_numpy(w["hidden_layer/%d/bias" % i].astype(np.float64))
        x = torch.relu(torch.nn.functional.conv2d(x, k, b, padding=k.shape[-1] // 2))
        if i < 5 and "a%d" % i in sources:
            x = rnd(x.float())
    h = x.permute(0, 2, 3, 1).reshape(n, 64)
    lg = h @ torch.from_numpy(w["output_layer/weights"].astype(np.float64)) + torch.from_numpy(w["output_layer/bias"].astype(np.float64))
    return lg.numpy(), torch.softmax(lg, 1).numpy()

ref_l, ref_p = forward(())
ALL = ["w1", "a1", "w2", "a2", "w3", "a3", "w4", "a4", "w5"]
def report(name, sources):
    l, p = forward(sources)
    d = np.abs(l - ref_l)
    r = {"logits_maxnorm": d.max() / np.abs(ref_l).max(), "logits_rel_rms": np.sqrt((d ** 2).mean()) / ref_l.std(),
         "probs_rel_median": float(np.median(np.abs(p - ref_p) / ref_p)),
         "probs_rel_p99": float(np.quantile(np.abs(p - ref_p) / ref_p, 0.99)),
         "probs_maxnorm": float((np.abs(p - ref_p).max(1) / ref_p.max(1)).max())}
    print("%-40s maxnorm %.2e relrms %.2e pmed %.2e p99 %.2e pmaxnorm %.2e" % (name, r["logits_maxnorm"], r["logits_rel_rms"], r["probs_rel_median"], r["probs_rel_p99"], r["probs_maxnorm"]))
    return r
out = {}
out["all"] = report("all (the fp16 mode)", ALL)
for s in ALL:
    out[s] = report("only " + s, [s])
combos = {
    "free fixes (w1,w5,a4 exact)": [s for s in ALL if s not in ("w1", "w5", "a4")],
    "free + all mid weights exact": ["a1", "a2", "a3"],
    "free + all mid acts exact": ["w2", "w3", "w4"],
    "free + L4 exact": ["a1", "w2", "a2", "w3"],
    "free + L2 exact": ["w3", "a2", "a3", "w4"],
    "free + w2,w3 exact": ["a1", "a2", "a3", "w4"],
    "free + w4,a3 + w3 exact": ["a1", "w2", "a2"],
}
for k, v in combos.items():
    out[k] = report(k, v)
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_error_attribution_%s.json" % fmt), "w"), indent=1)
