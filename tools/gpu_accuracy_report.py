#!/usr/bin/env python
"""Accuracy report on a large suite (development tool, run under gpurun): the tensor-core modes
against the fp32 CUDA-core mode (itself within 3e-6 of the CPU oracle, tests/test_gpu_parity.py)
and, on a sub-sample, against the CPU oracle directly."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic
from oracle import oracle as O, c_oracle as C

n = int(os.environ.get("N", "100000"))
sq, bl = synthetic.random_positions(n, seed=20260601)
seed = int(os.environ.get("WSEED", "1"))
w = W.synthetic("trained_like", seed)
with wann_b200.PolicyEvaluator(w, mode="fp32") as ev:
    p32, l32 = ev.eval_probs(sq, bl, return_logits=True)
    i32, _, c32 = ev.eval_topk(sq, bl)
mask = O.legal_mask(sq, bl)
lg = np.where(mask, l32, -np.inf); srt = np.sort(lg, axis=1); margin = srt[:, -1] - srt[:, -2]
m = 4096
C.set_threads(os.cpu_count())
lo, po = C.forward(O.encode_planes(sq[:m], bl[:m]), w)
rep = {"positions": n, "weights": "synthetic trained_like seed %d" % seed, "logit_sigma": float(l32.std()),
       "fp32_gpu_vs_cpu_oracle_logits_maxnorm": float(np.abs(l32[:m] - lo).max() / np.abs(lo).max()),
       "fp32_gpu_vs_cpu_oracle_top1": float((i32[:m, 0] == O.top1_legal(po, mask[:m])).mean())}
for mode, refine in (("bf16", 0.0), ("fp16", 0.0), ("bf16", 0.15), ("fp16", 0.02), ("fp16", 0.01)):
    with wann_b200.PolicyEvaluator(w, mode=mode, refine_margin=refine) as ev:
        p, l = ev.eval_probs(sq, bl, return_logits=True)
        r0 = ev.stats()["refined_positions"]
        idx, pri, cnt = ev.eval_topk(sq, bl)
        refined = ev.stats()["refined_positions"] - r0
    agree = idx[:, 0] == i32[:, 0]
    d = np.abs(l - l32)
    rep[mode if not refine else "%s+refine%g" % (mode, refine)] = {
        "refine_margin_logits": refine, "refined_fraction": refined / n,
        "logits_max_abs": float(d.max()), "logits_maxnorm": float(d.max() / np.abs(l32).max()),
        "logits_mean_abs": float(d.mean()),
        "logits_rel_rms": float(np.sqrt((d ** 2).mean()) / l32.std()),
        "probs_max_abs": float(np.abs(p - p32).max()),
        "probs_rel_median": float(np.median(np.abs(p - p32) / p32)),
        "top1_agreement": float(agree.mean()), "top1_disagreements": int((~agree).sum()),
        "max_oracle_margin_among_disagreements": float(margin[~agree].max()) if (~agree).any() else 0.0,
        "top1_agreement_where_margin_ge_0.05": float(agree[margin >= 0.05].mean()),
        "legal_count_equal": bool((cnt == c32).all()),
    }
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "accuracy_report%s.json" % ("" if seed == 1 else "_seed%d" % seed)), "w").write(json.dumps(rep, indent=1))
