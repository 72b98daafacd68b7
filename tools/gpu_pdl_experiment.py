import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import wann_b200
from wann_b200 import weights as W, synthetic
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1))
sq, bl = synthetic.random_positions(8192, seed=1)
sqt, blt = torch.from_numpy(np.tile(sq,(8,1))).cuda(), torch.from_numpy(np.tile(bl,8)).cuda()
def measure(n, reps=100):
    out = ev.eval_topk(sqt[:n], blt[:n])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ev.eval_topk(sqt[:n], blt[:n], out=out)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
ref = None
for n in (1, 256, 592, 1024, 4096, 65536):
    res = {0: [], 1: []}
    for rnd in range(3):
        for opt in (0, 1):
            ev.set_option("pdl", opt)
            res[opt].append(measure(n, 100 if n < 65536 else 20))
    print("N=%6d  pdl=0: %.1f us   pdl=1: %.1f us" % (n, min(res[0]), min(res[1])), flush=True)
# correctness under pdl
ev.set_option("pdl", 0); a = [x.cpu().numpy() for x in ev.eval_topk(sqt[:5000], blt[:5000])]
ev.set_option("pdl", 1)
for _ in range(5): b = ev.eval_topk(sqt[:5000], blt[:5000])
b = [x.cpu().numpy() for x in b]
print("pdl results identical:", all(np.array_equal(x, y) for x, y in zip(a, b)))
