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
M = min(N, 1 << 17)
idx = torch.empty((M, 48), dtype=torch.uint8, device=dev); pri = torch.empty((M, 48), dtype=torch.float32, device=dev)
cnt = torch.empty((M,), dtype=torch.uint8, device=dev)
ev.eval_topk(sq_t[:M], bl_t[:M], out=(idx, pri, cnt))
kids = torch.empty((M, 48, 64), dtype=torch.int8, device=dev); flg = torch.empty((M, 48), dtype=torch.uint8, device=dev)
st = torch.empty((M, 48, 4), dtype=torch.int32, device=dev); ps = torch.empty((M, 8), dtype=torch.int32, device=dev)
t = timed(lambda: L.check(lib.wann_expand_ranked(ctx, sq_t.data_ptr(), bl_t.data_ptr(), M, idx.data_ptr(), pri.data_ptr(), cnt.data_ptr(),
                                                 kids.data_ptr(), flg.data_ptr(), None, None, stream)))
rep["expand_children_kernel"] = {"positions": M, "bytes_per_position": 65 + 49 + 3072 + 48, "seconds": t, "gbs": M * 3234 / t / 1e9}
t = timed(lambda: L.check(lib.wann_expand_ranked(ctx, sq_t.data_ptr(), bl_t.data_ptr(), M, idx.data_ptr(), pri.data_ptr(), cnt.data_ptr(),
                                                 None, flg.data_ptr(), st.data_ptr(), ps.data_ptr(), stream)))
rep["seed_children_kernel"] = {"positions": M, "bytes_per_position": 65 + 48 + 192 + 1 + 768 + 48 + 32, "seconds": t, "gbs": M * 1154 / t / 1e9}
# context for the write-heavy kernels: what this GPU does for a pure write (fill) and a pure read (sum) of 1 GiB
big = torch.empty(1 << 28, dtype=torch.int32, device=dev)
t = timed(lambda: big.fill_(7)); rep["reference_fill_gbs"] = big.numel() * 4 / t / 1e9
t = timed(lambda: big.sum()); rep["reference_read_sum_gbs"] = big.numel() * 4 / t / 1e9
half = big[:1 << 27]; other = big[1 << 27:]
t = timed(lambda: other.copy_(half)); rep["reference_copy_gbs"] = half.numel() * 8 / t / 1e9
del big
# bit-exactness of what was just timed, on a sample, against the oracle
from oracle import oracle as O
k = 3000
assert (planes[:k].cpu().numpy() == O.encode_planes(sq_t[:k].cpu().numpy(), bl_t[:k].cpu().numpy())).all()
assert (mask[:k].cpu().numpy().astype(bool) == O.legal_mask(sq_t[:k].cpu().numpy(), bl_t[:k].cpu().numpy())).all()
assert (mask[N - k:].cpu().numpy().astype(bool) == O.legal_mask(sq_t[N - k:].cpu().numpy(), bl_t[N - k:].cpu().numpy())).all()
rep["bit_exact_vs_oracle"] = True
for k in ("encode_kernel", "legal_mask_kernel", "expand_children_kernel", "seed_children_kernel"):
    rep[k]["frac_of_hbm_peak"] = rep[k]["gbs"] / peak
    rep[k]["positions_per_s"] = rep[k].get("positions", N) / rep[k]["seconds"]
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "byte_kernels.json"), "w").write(json.dumps(rep, indent=1))
