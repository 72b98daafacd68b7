import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import _lib as L, weights as W, synthetic
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
lib, ctx = ev._lib, ev._ctx
sq, bl = synthetic.random_positions(4096, seed=9)
dev = torch.device("cuda", 0)
stream = ctypes.c_void_p(1)
def timed(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
for N in (1 << 18, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
    sq_t = torch.from_numpy(np.tile(sq, (N // 4096, 1))).to(dev); bl_t = torch.from_numpy(np.tile(bl, N // 4096)).to(dev)
    planes = torch.empty((N, 8, 8, 4), dtype=torch.int8, device=dev); mask = torch.empty((N, 155), dtype=torch.uint8, device=dev)
    t1 = timed(lambda: L.check(lib.wann_encode(ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, planes.data_ptr(), L.MEM_DEVICE, stream)))
    t2 = timed(lambda: L.check(lib.wann_legal_mask(ctx, sq_t.data_ptr(), bl_t.data_ptr(), N, mask.data_ptr(), L.MEM_DEVICE, stream)))
    ref = torch.empty(N * 321 // 8, dtype=torch.int32, device=dev); ref2 = torch.empty_like(ref)
    t3 = timed(lambda: ref2.copy_(ref))
    print("N=%d encode %.1f us %.0f GB/s | legal mask %.1f us %.0f GB/s | torch copy of the same bytes %.1f us %.0f GB/s" % (
        N, t1 * 1e6, N * 321 / t1 / 1e9, t2 * 1e6, N * 220 / t2 / 1e9, t3 * 1e6, ref.numel() * 8 / t3 / 1e9), flush=True)
    del sq_t, bl_t, planes, mask, ref, ref2
