#!/usr/bin/env python
"""Phase timing of the device-resident forest's lock steps at steady state (WANN_FOREST_TIMING=1 makes
wann_forest_run_steps bracket the phases of its first 48 steps with CUDA events).  Development tool."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, wann_b200
from wann_b200 import weights as W, synthetic
w = W.synthetic("trained_like", 1)
for T in [int(x) for x in os.environ.get("TREES", "1024,8192").split(",")]:
    sq, bl = synthetic.random_positions(T, seed=3)
    with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        ev.set_option("split_layers", int(os.environ.get("SPLIT", "0")))
        f = wann_b200.SearchForest(sq, bl, evaluator=ev, device_resident=True, node_capacity=32768)
        f.run_steps(ev, 60)
        os.environ["WANN_FOREST_TIMING"] = "1"
        f.run_steps(ev, 48)
        del os.environ["WANN_FOREST_TIMING"]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st, evs = f.run_steps(ev, 200)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(T, "trees: steady state", st, "steps", evs, "evals", "%.0f evals/s" % (evs / dt), "%.1f us/step" % (dt * 1e6 / st),
              "%.0f rows/step" % (evs / st), flush=True)
        f.close()
