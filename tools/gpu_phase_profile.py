#!/usr/bin/env python
"""Print per-layer MMA / epilogue cycle counts of CTA pair 0, both halves (development tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic
n = int(os.environ.get("N", "8192"))
sq, bl = synthetic.random_positions(min(n, 2048), seed=5)
reps = (n + len(sq) - 1) // len(sq)
sq = np.tile(sq, (reps, 1))[:n]; bl = np.tile(bl, reps)[:n]
ev = wann_b200.PolicyEvaluator(W.synthetic("trained_like", 1), mode=os.environ.get("MODE", "bf16"))
if os.environ.get("SPLIT"):
    ev.set_option("split_layers", int(os.environ["SPLIT"]))
print("mode", os.environ.get("MODE", "bf16"), "split_layers", os.environ.get("SPLIT", "0"))
for variant in [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]:
  ev.set_option("variant", variant & 0xFFFF)
  ev.set_option("skew", variant >> 16)
  print("==== variant", variant & 0xFFFF, "skew", variant >> 16)
  for _ in range(2):
    st = ev.debug_profile(sq, bl)
  t0 = st[st > 0].min()
  k = int((st[0] > 0).sum())
  print("layer-steps recorded", k, "(times relative to first stamp, cycles)")
  print("step L |  H0: mma_start  mma_issue  epi_wait  epi   |  H1: mma_start  mma_issue  epi_wait  epi  | lag(H1-H0 mma_start)")
  for i in range(5 if k > 10 else 0, min(k, 20)):
      a = st[0:4, i]; b = st[4:8, i]
      print("%3d  %d | %9d %9d %8d %6d | %9d %9d %8d %6d | %6d" % (
          i, i % 5, a[0] - t0, a[1] - a[0], a[2] - a[1], a[3] - a[2],
          b[0] - t0, b[1] - b[0], b[2] - b[1], b[3] - b[2], b[0] - a[0]))
  per_st = (st[0][5:k:5][1:] - st[0][5:k:5][:-1])
  print("cycles per supertile (pair 0): mean", per_st.mean() if len(per_st) else None, flush=True)
