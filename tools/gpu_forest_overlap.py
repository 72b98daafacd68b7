#!/usr/bin/env python
"""1,024 (and 8,192) trees as K device-resident forests driven by K host threads on one evaluator: the
lock steps of different forests overlap on the GPU (tree kernels of one under the conv stack of another).
Development tool, run under gpurun; writes gpurun_out/forest_overlap.json."""
import json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic

sims = int(os.environ.get("SIMS", "20"))
w = W.synthetic("trained_like", 1)
rep = {}
for mode, split, margin in (("fp16c", 8, 0.0), ("fp16c", 0, 0.02)):
    with wann_b200.PolicyEvaluator(w, mode=mode, refine_margin=margin) as ev:
        if split:
            ev.set_option("split_layers", split)
        for T in (1024, 8192):
            sq, bl = synthetic.random_positions(T, seed=3)
            for K in (1, 2, 4):
                forests = [wann_b200.SearchForest(sq[k::K], bl[k::K], evaluator=ev, device_resident=True) for k in range(K)]
                for f in forests:
                    f.run(ev, 1)
                res = [None] * K

                def work(k):
                    res[k] = forests[k].run(ev, sims)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                th = [threading.Thread(target=work, args=(k,)) for k in range(K)]
                [t.start() for t in th]
                [t.join() for t in th]
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                evals = sum(r[1] for r in res)
                key = "%s split%d refine%g trees%d forests%d" % (mode, split, margin, T, K)
                rep[key] = {"evaluations": evals, "wall_s": round(dt, 4), "evaluations_per_s": round(evals / dt, 1),
                            "steps_per_forest": res[0][0]}
                print(key, rep[key], flush=True)
                for f in forests:
                    f.close()
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "forest_overlap.json"), "w"), indent=1)
