#!/usr/bin/env python
"""Single-process multi-GPU throughput (MultiGpuEvaluator: one host thread per GPU, host buffers
sharded by position) next to the one-GPU number in the same process.  Development tool (gpurun --gpus N)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

G = torch.cuda.device_count()
per = int(os.environ.get("B", "65536"))
w = W.synthetic("trained_like", 1)
base_sq, base_bl = synthetic.random_positions(4096, seed=5)


def pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype).pin_memory().numpy()


def run(devices):
    n = per * len(devices)
    sq, bl = pinned((n, 64), torch.int8), pinned((n,), torch.uint8)
    sq[:] = np.tile(base_sq, (n // 4096, 1))
    bl[:] = np.tile(base_bl, n // 4096)
    out = (pinned((n, 48), torch.uint8), pinned((n, 48), torch.float32), pinned((n,), torch.uint8))
    with wann_b200.MultiGpuEvaluator(w, devices=devices, mode="bf16") as m:
        for _ in range(3):
            m.eval_topk(sq, bl, out=out)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            m.eval_topk(sq, bl, out=out)
        dt = time.perf_counter() - t0
    return n * reps / dt


rep = {"gpus": G, "positions_per_gpu_per_call": per, "api": "MultiGpuEvaluator.eval_topk, pinned host buffers, end to end"}
rep["one_gpu_positions_per_s"] = run([0])
rep["all_gpus_positions_per_s"] = run(list(range(G)))
rep["scaling"] = rep["all_gpus_positions_per_s"] / rep["one_gpu_positions_per_s"]
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "multi_single_process_%dgpu.json" % G), "w").write(json.dumps(rep, indent=1))
