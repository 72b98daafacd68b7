#!/usr/bin/env python
"""The honest GPU library comparator: the same policy net as plain
torch/cuDNN bf16 conv2d (channels_last) + matmul + softmax, timed like bench.py
(65,536 positions per step, CUDA events, device-resident inputs), next to this repo's kernels in
the same process.  Library code only -- NOT a product path; development tool, run under gpurun."""
import json, os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

B = int(os.environ.get("B", "65536"))
steps = int(os.environ.get("STEPS", "20"))
dev = torch.device("cuda", 0)
w = W.synthetic("trained_like", 1)
sq, bl = synthetic.random_positions(4096, seed=3)
reps = B // 4096
sq_t = torch.from_numpy(np.tile(sq, (reps, 1))).to(dev)
bl_t = torch.from_numpy(np.tile(bl, reps)).to(dev)
ev = wann_b200.PolicyEvaluator(w, mode="bf16")
planes = torch.from_numpy(ev.encode(np.tile(sq, (reps, 1)), np.tile(bl, reps))).to(dev)   # [B,8,8,4] int8

convs = []
for i in range(1, 6):
    k = torch.from_numpy(w["hidden_layer/%d/weights" % i]).permute(3, 2, 0, 1).contiguous()   # HWIO -> OIHW
    convs.append((k.to(dev, torch.bfloat16).contiguous(memory_format=torch.channels_last),
                  torch.from_numpy(w["hidden_layer/%d/bias" % i]).to(dev, torch.bfloat16)))
fc_w = torch.from_numpy(w["output_layer/weights"]).to(dev)
fc_b = torch.from_numpy(w["output_layer/bias"]).to(dev)
torch.backends.cudnn.benchmark = True


def torch_forward(pl):
    x = pl.permute(0, 3, 1, 2).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for k, b in convs:
        x = F.relu(F.conv2d(x, k, b, padding=1))
    h = x.float().reshape(-1, 64)
    return torch.softmax(h @ fc_w + fc_b, dim=1)


def timeit(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


chunk = 16384                     # cuDNN activations: 16,384 x 64 x 192 bf16 = 403 MB per layer
def lib_step():
    for o in range(0, B, chunk):
        torch_forward(planes[o:o + chunk])

out = (torch.empty((B, 48), dtype=torch.uint8, device=dev), torch.empty((B, 48), dtype=torch.float32, device=dev),
       torch.empty((B,), dtype=torch.uint8, device=dev))
probs = torch.empty((B, 155), dtype=torch.float32, device=dev)
ms_lib = timeit(lib_step, steps)
ms_ours_probs = timeit(lambda: ev.eval_probs(sq_t, bl_t, out=probs), steps)
ms_ours = timeit(lambda: ev.eval_topk(sq_t, bl_t, out=out), steps)
p_lib = torch_forward(planes[:1024]).float().cpu().numpy()
p_ours = ev.eval_probs(sq_t[:1024], bl_t[:1024]).cpu().numpy()
flop = 128527744.0 * B
rep = {"positions_per_step": B, "steps": steps,
       "torch_cudnn_bf16": {"ms_per_step": ms_lib, "positions_per_s": B / ms_lib * 1e3, "tflops": flop / ms_lib / 1e9,
                            "what": "pre-encoded planes -> 5 x F.conv2d(channels_last, bf16)+relu -> fp32 FC -> softmax; no legality / ranking"},
       "wann_b200_bf16_probs": {"ms_per_step": ms_ours_probs, "positions_per_s": B / ms_ours_probs * 1e3,
                                "what": "boards -> encode -> conv stack -> FC -> softmax ([N,155] out)"},
       "wann_b200_bf16_topk": {"ms_per_step": ms_ours, "positions_per_s": B / ms_ours * 1e3, "tflops": flop / ms_ours / 1e9,
                               "what": "boards -> ... -> legal gather + sort (ranked priors out)"},
       "speedup_vs_cudnn": ms_lib / ms_ours,
       "max_abs_prob_diff_vs_cudnn": float(np.abs(p_lib - p_ours).max()),
       "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "cudnn_comparator.json"), "w").write(json.dumps(rep, indent=1))
