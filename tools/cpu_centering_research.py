#!/usr/bin/env python
"""Offline study of the centred fp16 mode's offsets (CPU, torch float64 with the mode's fp16 roundings).
Which per-bucket offsets c minimise the end-to-end error?  A clamped activation (a = 0) is stored as
exactly -c, so only the POSITIVE activations carry a storage rounding error, while the next layer's weight
rounding error multiplies a' = a - c of every activation.
Usage: python tools/cpu_centering_research.py [n_positions]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wann_b200 import weights as W, synthetic
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.set_num_threads(os.cpu_count())
sq, bl = synthetic.random_positions(n, seed=20260601)
w = W.synthetic("trained_like", int(os.environ.get("WSEED", "1")))
planes = torch.from_numpy(O.encode_planes(sq, bl).astype(np.float64)).permute(0, 3, 1, 2)
K = [torch.from_numpy(w["hidden_layer/%d/weights" % i].astype(np.float64)).permute(3, 2, 0, 1).contiguous() for i in range(1, 6)]
B = [torch.from_numpy(w["hidden_layer/%d/bias" % i].astype(np.float64)) for i in range(1, 6)]
FW = torch.from_numpy(w["output_layer/weights"].astype(np.float64))
FB = torch.from_numpy(w["output_layer/bias"].astype(np.float64))

def r16(t):
    return t.to(torch.float32).to(torch.float16).to(torch.float64)

def head(x):
    h = x.permute(0, 2, 3, 1).reshape(n, 64)
    lg = h @ FW + FB
    return lg.numpy(), torch.softmax(lg, 1).numpy()

def exact():
    x = planes
    acts = []
    for i in range(5):
        x = torch.relu(torch.nn.functional.conv2d(x, K[i], B[i], padding=1))
        acts.append(x)
    return head(x), acts

(ref_l, ref_p), ref_acts = exact()

def centred(offsets, split=()):
    """offsets: list of 4 tensors [192] (per TF channel).  split: layers (2..4) whose weights are exact."""
    x = planes
    c_prev = None
    for i in range(5):
        layer = i + 1
        k16 = K[i] if (layer in (1, 5) or layer in split) else r16(K[i])
        if c_prev is None:
            z = torch.nn.functional.conv2d(x, k16, B[i], padding=1)
        else:
            xp = torch.empty(n, 192, 10, 10, dtype=torch.float64)
            xp[:] = -c_prev[None, :, None, None]
            xp[:, :, 1:-1, 1:-1] = x
            const = (K[i] * c_prev[None, :, None, None]).sum(dim=(1, 2, 3))
            z = torch.nn.functional.conv2d(xp, k16, B[i] + const)
        if layer < 5:
            c = offsets[i]
            x = r16(torch.maximum(z - c[None, :, None, None], -c[None, :, None, None]))
            c_prev = c
        else:
            x = torch.relu(z)
    return head(x)

def report(name, offsets, split=()):
    l, p = centred(offsets, split)
    d = np.abs(l - ref_l)
    print("%-54s maxnorm %.2e relrms %.3e pmed %.3e" % (name, d.max() / np.abs(ref_l).max(),
          np.sqrt((d ** 2).mean()) / ref_l.std(), float(np.median(np.abs(p - ref_p) / ref_p))), flush=True)

def bucketise(vals, mean, zero_buckets=0):
    """8 buckets by sorted MEAN (as calibrate_centering), offset of a bucket = fp16(mean of vals over it)."""
    order = torch.argsort(mean, stable=True)
    c = torch.zeros(192, dtype=torch.float64)
    for j in range(8):
        idx = order[j * 24:(j + 1) * 24]
        c[idx] = 0.0 if j < zero_buckets else float(r16(vals[idx].mean()))
    return c

means = [a.mean(dim=(0, 2, 3)) for a in ref_acts[:4]]
pos_means = [(a.sum(dim=(0, 2, 3)) / (a > 0).sum(dim=(0, 2, 3)).clamp(min=1)) for a in ref_acts[:4]]
frac_pos = [(a > 0).double().mean(dim=(0, 2, 3)) for a in ref_acts[:4]]
print("fraction positive per layer:", [float(f.mean()) for f in frac_pos])
print("mean / positive-mean per layer:", [(float(m.mean()), float(pm.mean())) for m, pm in zip(means, pos_means)])

zero = [torch.zeros(192, dtype=torch.float64)] * 4
FULL = os.environ.get("FULL", "0") == "1"      # FULL=1: also the exploratory sections
def report0(*a, **k):
    if FULL: report(*a, **k)
report("plain fp16 (offsets 0), w1/w5 exact", zero)
report0("bucket mean (current, no zero buckets)", [bucketise(m, m) for m in means])
report0("bucket mean, 2 zero buckets (shipped)", [bucketise(m, m, 2) for m in means])
report0("bucket mean, 2 zero buckets, split L4", [bucketise(m, m, 2) for m in means], split=(4,))
for alpha in (1.25, 1.5, 2.0):
    report0("bucket mean x %.2f" % alpha, [bucketise(m * alpha, m) for m in means])
# J-optimal per channel: minimise 2 * P+ E[(a-c)^2 | a>0] + P0 c^2  ->  c = 2 S1 / (2 P+ + P0), S1 = E[a] (a>0 part)
jopt = [2 * m / (2 * f + (1 - f)) for m, f in zip(means, frac_pos)]
report0("J-optimal offsets, bucketised", [bucketise(j, m) for j, m in zip(jopt, means)])
report0("J-optimal offsets, bucketised, 2 zero buckets", [bucketise(j, m, 2) for j, m in zip(jopt, means)])
report0("per-channel mean (not implementable: bound)", [r16(m) for m in means])
report0("per-channel J-optimal (bound)", [r16(j) for j in jopt])
report0("J-optimal, split L4 (c3 = positive mean)", [bucketise(j, m) for j, m in zip(jopt[:3], means[:3])] + [bucketise(jopt[3], means[3])], split=(4,))

# ---- finer bucketing: G groups of consecutive chunks x 8 positions = 8 G buckets -------------------
def bucketise_n(vals, mean, nb, zero_buckets=0):
    order = torch.argsort(mean, stable=True)
    c = torch.zeros(192, dtype=torch.float64)
    per = 192 // nb
    for j in range(nb):
        idx = order[j * per:(j + 1) * per]
        c[idx] = 0.0 if j < zero_buckets else float(r16(vals[idx].mean()))
    return c
for nb in (16, 24, 48):
    report0("%d buckets, mean" % nb, [bucketise_n(m, m, nb) for m in means])
    report0("%d buckets, J-optimal" % nb, [bucketise_n(j, m, nb) for j, m in zip(jopt, means)])
report0("24 buckets, mean, 6 zero buckets", [bucketise_n(m, m, 24, 6) for m in means])
report0("24 buckets, mean, 3 zero buckets", [bucketise_n(m, m, 24, 3) for m in means])

# ---- partial hi/lo split: exact weights only for the (output, input) channel block that matters most -------
def centred_partial(offsets, keep_in, keep_out):
    """keep_in[l] / keep_out[l] (l = 2..4): index tensors of input / output channels whose weights are exact
    (hi + lo) in layer l; everything else fp16-rounded."""
    x = planes
    c_prev = None
    for i in range(5):
        layer = i + 1
        if layer in (1, 5):
            k16 = K[i]
        else:
            k16 = r16(K[i])
            if layer in keep_in:
                ki, ko = keep_in[layer], keep_out[layer]
                blk = K[i][ko][:, ki]
                tmp = k16[ko]
                tmp[:, ki] = blk
                k16[ko] = tmp
        if c_prev is None:
            z = torch.nn.functional.conv2d(x, k16, B[i], padding=1)
        else:
            xp = torch.empty(n, 192, 10, 10, dtype=torch.float64)
            xp[:] = -c_prev[None, :, None, None]
            xp[:, :, 1:-1, 1:-1] = x
            const = (K[i] * c_prev[None, :, None, None]).sum(dim=(1, 2, 3))
            z = torch.nn.functional.conv2d(xp, k16, B[i] + const)
        if layer < 5:
            c = offsets[i]
            x = r16(torch.maximum(z - c[None, :, None, None], -c[None, :, None, None]))
            c_prev = c
        else:
            x = torch.relu(z)
    return head(x)

def report_partial(name, offsets, keep_in, keep_out):
    l, p = centred_partial(offsets, keep_in, keep_out)
    d = np.abs(l - ref_l)
    print("%-64s maxnorm %.2e relrms %.3e pmed %.3e" % (name, d.max() / np.abs(ref_l).max(),
          np.sqrt((d ** 2).mean()) / ref_l.std(), float(np.median(np.abs(p - ref_p) / ref_p))), flush=True)

off24 = [bucketise_n(m, m, 24) for m in means]
off8 = [bucketise(m, m, 2) for m in means]
var = [a.var(dim=(0, 2, 3)) for a in ref_acts[:4]]
top_in = {l: torch.argsort(var[l - 2], descending=True) for l in (2, 3, 4)}       # inputs of layer l = outputs of l-1
top_out = {l: torch.argsort(frac_pos[l - 1], descending=True) for l in (2, 3, 4)}
for l in (2, 3, 4):
    v = var[l - 2][top_in[l]]
    print("layer %d input variance share of the top 64 / 128 channels: %.2f %.2f" % (l, float(v[:64].sum() / v.sum()), float(v[:128].sum() / v.sum())))
allc = torch.arange(192)
for offs, oname in ((off8, "8b"), (off24, "24b")):
    for nk, no, layers in ((64, 192, (2, 3, 4)), (64, 96, (2, 3, 4)), (128, 96, (2, 3, 4)), (64, 192, (3, 4)), (128, 192, (4,)),
                           (64, 128, (2, 3, 4)), (128, 128, (2, 3, 4))):
        ki = {l: top_in[l][:nk] for l in layers}
        ko = {l: top_out[l][:no] for l in layers}
        cost = sum(nk / 192 * no / 192 for l in layers) / 3
        if FULL: report_partial("%s: exact top %d in x top %d out, layers %s (+%.1f %% MMAs)" % (oname, nk, no, layers, 100 * cost), offs, ki, ko)

# ---- the implementable scheme: positions by descending variance in 3 groups of 64 (own halo rows per group),
# ---- 8 mean-buckets inside a group, lo pass = K group 0 (top 64 inputs) x output positions 0..95 -------------
def scheme(zero_frac=0.25):
    offs, pos = [], []
    for l in range(4):
        order = torch.argsort(var[l], descending=True)
        c = torch.zeros(192, dtype=torch.float64)
        p = torch.zeros(192, dtype=torch.long)
        bucket_means = []
        for g in range(3):
            grp = order[64 * g:64 * (g + 1)]
            grp = grp[torch.argsort(means[l][grp], stable=True)]          # by mean inside the group
            for j in range(8):
                idx = grp[8 * j:8 * (j + 1)]
                idx = idx[torch.argsort(var[l][idx], descending=True)]      # most important first inside a bucket
                cj = float(r16(means[l][idx].mean()))
                bucket_means.append((cj, idx))
                p[idx] = 64 * g + 8 * torch.arange(8) + j
        bucket_means.sort(key=lambda t: t[0])
        nz = int(round(zero_frac * 24))
        for r, (cj, idx) in enumerate(bucket_means):
            c[idx] = 0.0 if r < nz else cj
        offs.append(c); pos.append(p)
    return offs, pos
for zf in (0.0, 0.25):
    offs, pos = scheme(zf)
    for nk, no, layers in ((64, 96, (2, 3, 4)), (64, 96, (3, 4)), (64, 96, (2, 4)), (64, 96, (2, 3)), (64, 64, (2, 3, 4)), (64, 128, (2, 3, 4)), (64, 48, (2, 3, 4)),
                           (64, 192, (2, 3, 4)), (128, 96, (2, 3, 4)), (0, 0, ())):
        ki = {l: torch.nonzero(pos[l - 2] < nk).flatten() for l in layers}
        ko = {l: torch.nonzero(pos[l - 1] < no).flatten() for l in layers}
        cost = sum(nk / 192 * no / 192 for l in layers) / 3
        report_partial("scheme zf=%.2f: lo K<%d x N<%d, layers %s (+%.1f %% MMAs)" % (zf, nk, no, layers, 100 * cost), offs, ki, ko)
