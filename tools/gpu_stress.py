#!/usr/bin/env python
"""Soak test: random batch sizes, call paths and option combinations back to back (asynchronous device
calls with programmatic dependent launch, host calls, 4 concurrent host threads), every result compared
with a reference evaluation of the same rows (the path is batch-invariant: row i depends on input i
only).  Development tool, run under gpurun:  SECONDS=60 python tools/gpu_stress.py"""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

secs = float(os.environ.get("SECONDS", "40"))
rng = np.random.RandomState(int(os.environ.get("SEED", "1")))
w = W.synthetic("trained_like", 1)
P = 20000
sq, bl = synthetic.random_positions(P, seed=11)
dev = torch.device("cuda", 0)
sq_t, bl_t = torch.from_numpy(sq).to(dev), torch.from_numpy(bl).to(dev)
report = {}
for mode, margin in (("bf16", 0.0), ("fp16", 0.02), ("fp32_tc", 0.0)):
    with wann_b200.PolicyEvaluator(w, mode=mode, refine_margin=margin) as ev:
        ref_idx, ref_pri, ref_cnt = ev.eval_topk(sq, bl)
        ref_pr = ev.eval_probs(sq, bl)
        ref_seed = ev.eval_expand_seeded(sq, bl)           # idx, prior, n_legal, children, flags, stats, parent
        n_checks, errs = [0], []

        def check(lo, n, idx, pri, cnt, pr):
            ok = (np.array_equal(idx, ref_idx[lo:lo + n]) and np.array_equal(pri, ref_pri[lo:lo + n]) and
                  np.array_equal(cnt, ref_cnt[lo:lo + n]) and (pr is None or np.array_equal(pr, ref_pr[lo:lo + n])))
            n_checks[0] += 1
            if not ok:
                errs.append((lo, n))

        def host_thread(seed):
            r = np.random.RandomState(seed)
            while time.time() < t_end and not errs:
                n = int(r.choice([1, 2, 5, 17, 64, 300, 513, 2000]))
                lo = int(r.randint(0, P - n))
                i, p, c = ev.eval_topk(sq[lo:lo + n], bl[lo:lo + n])
                check(lo, n, i, p, c, ev.eval_probs(sq[lo:lo + n], bl[lo:lo + n]) if r.rand() < 0.3 else None)
                if r.rand() < 0.2:                             # the widest call, host path
                    got = ev.eval_expand_seeded(sq[lo:lo + n], bl[lo:lo + n], with_children=bool(r.randint(2)))
                    n_checks[0] += 1
                    if not all(g is None or np.array_equal(g, rf[lo:lo + n]) for g, rf in zip(got, ref_seed)):
                        errs.append(("seeded-host", lo, n))

        t_end = time.time() + secs / 3
        threads = [threading.Thread(target=host_thread, args=(s,)) for s in range(4)]
        [t.start() for t in threads]
        while time.time() < t_end and not errs:
            for key in ("fuse_tail", "zero_copy", "small_batch_halves", "coalesce"):
                ev.set_option(key, int(rng.randint(2)))
            ev.set_option("pdl", int(rng.randint(-1, 2)))
            pend = []
            for _ in range(int(rng.randint(1, 12))):           # a burst of asynchronous device calls
                n = int(rng.choice([1, 3, 8, 100, 296, 297, 592, 1024, 1500, 4096, 9000]))
                lo = int(rng.randint(0, P - n))
                out = ev.eval_topk(sq_t[lo:lo + n], bl_t[lo:lo + n])
                pend.append((lo, n, out))
            n = int(rng.choice([1, 7, 600, 5000]))
            lo = int(rng.randint(0, P - n))
            seeded = ev.eval_expand_seeded(sq_t[lo:lo + n], bl_t[lo:lo + n], with_children=bool(rng.randint(2)))
            torch.cuda.synchronize()
            n_checks[0] += 1
            if not all(g is None or np.array_equal(g.cpu().numpy(), rf[lo:lo + n]) for g, rf in zip(seeded, ref_seed)):
                errs.append(("seeded-device", lo, n))
            for lo, n, out in pend:
                check(lo, n, out[0].cpu().numpy(), out[1].cpu().numpy(), out[2].cpu().numpy(), None)
        [t.join() for t in threads]
        report[mode] = (n_checks[0], errs[:3], ev.stats())
        print(mode, "margin", margin, ":", n_checks[0], "checks,", len(errs), "mismatches", errs[:3], flush=True)
assert all(not r[1] for r in report.values()), "MISMATCH"
print("stress ok")
