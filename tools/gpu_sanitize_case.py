#!/usr/bin/env python
"""Small end-to-end case for compute-sanitizer (development tool): every kernel of the library on small,
ragged inputs -- small chunks (the zero-copy capacity fix), the centred mode with split layers and margin
refinement, top-k truncation, the byte kernels, child construction / seeding, the device-resident forest."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic
sq, bl = synthetic.random_positions(77, seed=2)
w = W.synthetic("trained_like", 1)
for mode, margin in (("bf16", 0.0), ("fp32", 0.0), ("fp16c", 0.05), ("fp32_tc", 0.0)):
    with wann_b200.PolicyEvaluator(w, mode=mode, chunk=32, refine_margin=margin) as ev:
        if mode == "fp16c":
            ev.set_option("split_layers", 14)
        p, l = ev.eval_probs(sq, bl, return_logits=True)
        out = ev.eval_expand(sq, bl)
        ev.eval_expand_seeded(sq[:33], bl[:33], k=5)
        ev.encode(sq, bl); ev.legal_mask(sq, bl)
        ev.eval_planes(ev.encode(sq, bl))
        if mode in ("bf16", "fp16c"):
            ev.debug_activations(sq[:9], bl[:9], 2)
            d = torch.device("cuda", 0)
            ev.eval_topk(torch.from_numpy(sq).to(d), torch.from_numpy(bl).to(d), k=3)
            torch.cuda.synchronize()
        print(mode, "ok", float(p.sum()), int(out[2].sum()), flush=True)
with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
    f = wann_b200.SearchForest(sq[:9], bl[:9], evaluator=ev, device_resident=True, node_capacity=2048)
    print("forest", f.run(ev, 3), f.run_steps(ev, 5), f.run(ev, 0), flush=True)
    print("play", f.play()[0][:4], f.info(0)["nodes"], flush=True)
    f.close()
