#!/usr/bin/env python
"""Device-resident search forest (wann_forest_create_device / wann_forest_run) against the host forest:
evaluations per second for T trees, same evaluator (fp16c + 0.02 margin refinement), and a CUDA-event
split of one lock step.  Development tool, run under gpurun; writes gpurun_out/forest_device.json."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import wann_b200
from wann_b200 import weights as W, synthetic

sims = int(os.environ.get("SIMS", "20"))
w = W.synthetic("trained_like", 1)
rep = {"simulations_per_tree": sims, "host_cores": os.cpu_count(), "mode": "fp16c + 0.02 margin refinement"}
with wann_b200.PolicyEvaluator(w, mode="fp16c", refine_margin=0.02) as ev:
    for T in [int(x) for x in os.environ.get("TREES", "1,64,1024,8192,32768").split(",")]:
        sq = np.tile(synthetic.initial_squares()[None], (T, 1))
        bl = np.zeros(T, np.uint8)
        if T > 1:
            rs, rb = synthetic.random_positions(T - 1, seed=3)
            sq[1:], bl[1:] = rs, rb
        r = {"trees": T}
        for kind in ("device", "host"):
            if kind == "host" and T > 8192:
                continue
            f = wann_b200.SearchForest(sq, bl, evaluator=ev, device_resident=(kind == "device"),
                                       node_capacity=int(os.environ.get("CAP", "16384")))
            f.run(ev, 1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            steps, evals = f.run(ev, sims)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if kind == "device":                         # the tree kernels, step by step (last 64 lock steps)
                f2 = wann_b200.SearchForest(sq, bl, evaluator=ev, device_resident=True,
                                            node_capacity=int(os.environ.get("CAP", "16384")))
                f2.run(ev, 1)                            # ~40 lock steps: all inside the 64-step history
                c = f2.debug_cycles()
                f2.close()
                c = c[c[:, 0] > 0]
                r["tree_kernel_cycles_by_step"] = [[int(x) for x in row] for row in c[::max(1, len(c) // 12)]]
                r["tree_kernel_cycles_per_step"] = {
                    "steps": int(len(c)), "rows": float(c[:, 0].mean()),
                    "next_mean_per_tree": float(c[:, 1].sum() / T / max(len(c), 1)), "next_max": float(c[:, 2].mean()),
                    "apply_mean_per_row": float((c[:, 3] / np.maximum(c[:, 0], 1)).mean()), "apply_max": float(c[:, 4].mean())}
            sample = range(0, T, max(1, T // 128))
            infos = [f.info(t) for t in sample]
            r[kind] = {"steps": steps, "evaluations": evals, "wall_s": round(dt, 4), "evaluations_per_s": round(evals / dt, 1),
                       "us_per_step": round(dt * 1e6 / max(steps, 1), 1), "mean_batch": round(evals / max(steps, 1), 1),
                       "mean_nodes_per_tree": float(np.mean([i["nodes"] for i in infos])),
                       "overflowed": int(sum(i["finished"] == 2 for i in infos))}
            f.close()
        rep[str(T)] = r
        print(T, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "forest_device.json"), "w"), indent=1)
