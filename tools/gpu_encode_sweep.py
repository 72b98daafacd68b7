import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import wann_b200
from wann_b200 import _lib as L, weights as W, synthetic
N = 1 << 20
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
lib, ctx = ev._lib, ev._ctx
sq, bl = synthetic.random_positions(4096, seed=9)
dev = torch.device("cuda", 0)
sq_t = torch.from_numpy(np.tile(sq, (N // 4096, 1))).to(dev); bl_t = torch.from_numpy(np.tile(bl, N // 4096)).to(dev)
planes = torch.empty((N, 8, 8, 4), dtype=torch.int8, device=dev)
stream = ctypes.c_void_p(1)
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for unroll_sel, name in ((0, "u4 stcs"), (1, "u2 stcs"), (2, "u8 stcs"), (3, "u4 plain"), (4, "u8 plain"), (5, "u1 plain")):
    for per_sm in (4, 8, 16, 0):
        ev.set_option("variant", (unroll_sel << 8) | (per_sm << 12))
        t = timed(lambda: L.check(lib.wann_encode(ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, planes.data_ptr(), L.MEM_DEVICE, stream)))
        print(name, "blocks/SM", per_sm or 16, "%.1f us  %.0f GB/s" % (t * 1e6, N * 321 / t / 1e9), flush=True)
