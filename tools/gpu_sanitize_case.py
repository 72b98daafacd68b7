#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (development tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic
sq, bl = synthetic.random_positions(77, seed=2)
for mode in ("bf16", "fp32"):
    with wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), mode=mode, chunk=32) as ev:
        p, l = ev.eval_probs(sq, bl, return_logits=True)
        out = ev.eval_expand(sq, bl)
        ev.encode(sq, bl); ev.legal_mask(sq, bl)
        ev.eval_planes(ev.encode(sq, bl))
        if mode == "bf16":
            ev.debug_activations(sq[:9], bl[:9], 2)
        print(mode, "ok", float(p.sum()), int(out[2].sum()))
