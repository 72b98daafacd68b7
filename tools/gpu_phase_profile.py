#!/usr/bin/env python
"""Print per-layer MMA / epilogue cycle counts of CTA pair 0 (development tool).
VARIANTS env: comma list of descriptor/timing variants to try (timing variants give garbage results)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O
import wann_b200
from wann_b200 import weights as W
n = int(os.environ.get("N", "8192"))
sq, bl = O.random_positions_fast(min(n, 2048), seed=5)
reps = (n + len(sq) - 1) // len(sq)
sq = np.tile(sq, (reps, 1))[:n]; bl = np.tile(bl, reps)[:n]
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), mode="bf16")
for variant in [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]:
    ev.set_option("variant", variant)
    for _ in range(2):
        st = ev.debug_profile(sq, bl)
    mma0, mma1, ep0, ep1 = st
    k = int((mma0 > 0).sum())
    print("== variant %d: layer-steps recorded %d" % (variant, k))
    print("step layer  mma_issue  wait_acc(ep0-mma1)  epilogue  a_ready->mma(next)  total_step")
    for i in range(5, min(k, 15)):
        nxt = mma0[i + 1] - ep1[i] if i + 1 < k else 0
        tot = mma0[i + 1] - mma0[i] if i + 1 < k else 0
        print("%3d   %d   %8d   %8d   %8d   %8d   %8d" % (i, i % 5, mma1[i] - mma0[i], ep0[i] - mma1[i], ep1[i] - ep0[i], nxt, tot))
    per_st = (mma0[5:k:5][1:] - mma0[5:k:5][:-1])
    print("cycles per supertile (pair 0): mean", per_st.mean() if len(per_st) else None, flush=True)
