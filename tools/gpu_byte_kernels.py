#!/usr/bin/env python
"""Achieved HBM bandwidth of the byte/integer kernels beside the tensor-core path (encode, legal mask,
child construction, child seeding), device resident, CUDA events, against the measured copy
bandwidth in MEASURED_PEAKS.json.  Development tool (run under gpurun)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import _lib as L, weights as W, synthetic

N = int(os.environ.get("N", str(1 << 20)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6533.5) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6533.5
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
ev.set_option("chunk", 16384)
lib, ctx = ev._lib, ev._ctx
sq, bl = synthetic.random_positions(4096, seed=9)
dev = torch.device("cuda", 0)
sq_t = torch.from_numpy(np.tile(sq, (N // 4096, 1))).to(dev)
bl_t = torch.from_numpy(np.tile(bl, N // 4096)).to(dev)
stream = ctypes.c_void_p(1)
planes = torch.empty((N, 8, 8, 4), dtype=torch.int8, device=dev)
mask = torch.empty((N, 155), dtype=torch.uint8, device=dev)


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


rep = {"positions": N, "hbm_peak_gbs": peak}
t = timed(lambda: L.check(lib.wann_encode(ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, planes.data_ptr(), L.MEM_DEVICE, stream)))
rep["encode_kernel"] = {"bytes_per_position": 65 + 256, "seconds": t, "gbs": N * 321 / t / 1e9}
t = timed(lambda: L.check(lib.wann_legal_mask(ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, mask.data_ptr(), L.MEM_DEVICE, stream)))
rep["legal_mask_kernel"] = {"bytes_per_position": 65 + 155, "seconds": t, "gbs": N * 220 / t / 1e9}
for k in ("encode_kernel", "legal_mask_kernel"):
    rep[k]["frac_of_hbm_peak"] = rep[k]["gbs"] / peak
    rep[k]["positions_per_s"] = N / rep[k]["seconds"]
rep["note"] = ("expand_children_kernel / seed_children_kernel run behind an evaluation: their durations come from the ncu "
               "launch list (tools/gpu_byte_kernels_ncu.py, profiles/r01_ncu_byte_kernels.csv)")
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "byte_kernels.json"), "w").write(json.dumps(rep, indent=1))
