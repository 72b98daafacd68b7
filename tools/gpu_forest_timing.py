import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch, wann_b200
from wann_b200 import weights as W, synthetic
w = W.synthetic("trained_like", 1)
for T in (1024, 8192):
    sq, bl = synthetic.random_positions(T, seed=3)
    with wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        ev.set_option("split_layers", 8)
        f = wann_b200.SearchForest(sq, bl, evaluator=ev, device_resident=True)
        f.run(ev, 1)
        os.environ["WANN_FOREST_TIMING"] = "1"
        f.run(ev, 3)
        del os.environ["WANN_FOREST_TIMING"]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st, evs = f.run(ev, 20)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(T, "steps", st, "evals", evs, "evals/s", evs / dt, "us/step", dt * 1e6 / st, flush=True)
        f.close()
