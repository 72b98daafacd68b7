#!/usr/bin/env python
"""Offline study of data-aware (GPTQ-style) rounding of the 192->192 layers' weights to fp16 (CPU, torch float64 with
the centred mode's roundings): second moments of the calibration inputs, sequential rounding with error compensation,
end-to-end error on HELD-OUT positions; variants (order, damping, calibration size, fp32 H, zero buckets).
The numbers it printed are what the kernel later measured."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wann_b200 import weights as W, synthetic
from oracle import oracle as O
torch.set_num_threads(os.cpu_count())
n_cal, n_test = 512, 1024
n = n_cal + n_test
sq, bl = synthetic.random_positions(n, seed=20260601)
w = W.synthetic("trained_like", int(os.environ.get("WSEED", "1")))
planes = torch.from_numpy(O.encode_planes(sq, bl).astype(np.float64)).permute(0, 3, 1, 2)
K = [torch.from_numpy(w["hidden_layer/%d/weights" % i].astype(np.float64)).permute(3, 2, 0, 1).contiguous() for i in range(1, 6)]
B = [torch.from_numpy(w["hidden_layer/%d/bias" % i].astype(np.float64)) for i in range(1, 6)]
FW = torch.from_numpy(w["output_layer/weights"].astype(np.float64)); FB = torch.from_numpy(w["output_layer/bias"].astype(np.float64))
def r16(t): return t.to(torch.float32).to(torch.float16).to(torch.float64)
def head(x):
    h = x.permute(0, 2, 3, 1).reshape(x.shape[0], 64); lg = h @ FW + FB
    return lg.numpy(), torch.softmax(lg, 1).numpy()
x = planes; ref_acts = []
for i in range(5):
    x = torch.relu(torch.nn.functional.conv2d(x, K[i], B[i], padding=1)); ref_acts.append(x)
ref_l, ref_p = head(x)
means = [a[:n_cal].mean(dim=(0, 2, 3)) for a in ref_acts[:4]]
var = [a[:n_cal].var(dim=(0, 2, 3)) for a in ref_acts[:4]]
# the shipped centring: 3 groups by variance, 8 buckets by mean, 12 lowest-mean buckets zero
def scheme(zero=12):
    offs, pos = [], []
    for l in range(4):
        order = torch.argsort(var[l], descending=True)
        c = torch.zeros(192, dtype=torch.float64); p = torch.zeros(192, dtype=torch.long); bm = []
        for g in range(3):
            grp = order[64 * g:64 * (g + 1)]; grp = grp[torch.argsort(means[l][grp], stable=True)]
            for j in range(8):
                idx = grp[8 * j:8 * (j + 1)]; idx = idx[torch.argsort(var[l][idx], descending=True)]
                bm.append((float(r16(means[l][idx].mean())), idx)); p[idx] = 64 * g + 8 * torch.arange(8) + j
        bm.sort(key=lambda t: t[0])
        for r, (cj, idx) in enumerate(bm): c[idx] = 0.0 if r < zero else cj
        offs.append(c); pos.append(p)
    return offs, pos
offs, pos = scheme()

def run(k16s, offs, lo=None, bias_fix=None):
    """k16s: the (possibly optimised) fp16 weights of layers 2..4 (dict layer->tensor [O,I,3,3])."""
    x = planes; c_prev = None
    for i in range(5):
        layer = i + 1
        k16 = K[i] if layer in (1, 5) else k16s[layer]
        if c_prev is None:
            z = torch.nn.functional.conv2d(x, k16, B[i], padding=1)
        else:
            xp = torch.empty(x.shape[0], 192, 10, 10, dtype=torch.float64); xp[:] = -c_prev[None, :, None, None]; xp[:, :, 1:-1, 1:-1] = x
            const = (K[i] * c_prev[None, :, None, None]).sum(dim=(1, 2, 3))
            b = B[i] + const
            if bias_fix is not None and layer in bias_fix: b = b + bias_fix[layer]
            z = torch.nn.functional.conv2d(xp, k16, b)
        if layer < 5:
            c = offs[i]; x = r16(torch.maximum(z - c[None, :, None, None], -c[None, :, None, None])); c_prev = c
        else:
            x = torch.relu(z)
    return head(x)
def report(name, k16s, **kw):
    l, p = run(k16s, offs, **kw)
    sl = slice(n_cal, n)
    d = np.abs(l[sl] - ref_l[sl])
    print("%-50s maxnorm %.2e relrms %.3e pmed %.3e" % (name, d.max() / np.abs(ref_l[sl]).max(), np.sqrt((d ** 2).mean()) / ref_l[sl].std(),
          float(np.median(np.abs(p[sl] - ref_p[sl]) / ref_p[sl]))), flush=True)

rtn = {l: r16(K[l - 1]) for l in (2, 3, 4)}
report("RTN (shipped, no lo)", rtn)
def with_lo(base, layers=(2, 3, 4), nk=64, no=64):
    out = {}
    for l in (2, 3, 4):
        k = base[l].clone()
        if l in layers:
            ki = torch.nonzero(pos[l - 2] < nk).flatten(); ko = torch.nonzero(pos[l - 1] < no).flatten()
            tmp = k[ko]; tmp[:, ki] = K[l - 1][ko][:, ki]; k[ko] = tmp
        out[l] = k
    return out
report("RTN + lo 14:64:1", with_lo(rtn))

# bias fold of the systematic weight error: sum e * (mean - c) over all taps (interior pixels)
bf = {}
for l in (2, 3, 4):
    e = K[l - 1] - rtn[l]
    bf[l] = (e * (means[l - 2] - offs[l - 2])[None, :, None, None]).sum(dim=(1, 2, 3))
report("RTN + bias fold of e*(mean-c)", rtn, bias_fix=bf)

# ---- GPTQ: Hessian of the im2col'd centred activations of the calibration positions
def im2col(l):   # input of layer l (2..4) = stored activation of layer l-1, centred, with -c halo; exact activations used
    a = ref_acts[l - 2][:n_cal]; c = offs[l - 2]
    xp = torch.empty(n_cal, 192, 10, 10, dtype=torch.float64); xp[:] = -c[None, :, None, None]; xp[:, :, 1:-1, 1:-1] = a - c[None, :, None, None]
    cols = torch.nn.functional.unfold(xp, 3)            # [n, 192*9, 64]  ordering (channel, ky, kx)
    return cols.permute(0, 2, 1).reshape(-1, 1728)
def gptq(l, damp=0.01, order="act"):
    X = im2col(l)
    H = (X.T @ X) / X.shape[0]
    ev = torch.linalg.eigvalsh(H)
    print("layer %d: H trace %.3g, effective rank %.1f, top eig share %.3f, eig quantiles %s" % (l, float(ev.sum()), float(ev.sum() ** 2 / (ev ** 2).sum()), float(ev[-1] / ev.sum()),
          [float("%.3g" % ev[int(q * 1727)]) for q in (0.0, 0.1, 0.5, 0.9, 0.99)]), flush=True)
    Wm = K[l - 1].reshape(192, 1728).clone()            # (channel, ky, kx) ordering matches unfold
    perm = torch.argsort(torch.diag(H), descending=True) if order == "act" else torch.arange(1728)
    Hp = H[perm][:, perm] + damp * torch.diag(H).mean() * torch.eye(1728, dtype=torch.float64)
    Wp = Wm[:, perm]
    Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hp)), upper=True)
    Q = torch.zeros_like(Wp)
    for k in range(1728):
        q = r16(Wp[:, k]); Q[:, k] = q
        err = (Wp[:, k] - q) / Hinv[k, k]
        if k + 1 < 1728: Wp[:, k + 1:] -= err[:, None] * Hinv[k, k + 1:][None, :]
    inv = torch.argsort(perm)
    Qm = Q[:, inv]
    # layer-local check: output error energy RTN vs GPTQ on the calibration set
    E_rtn = (Wm - r16(Wm)); E_g = (Wm - Qm)
    print("   layer-local error energy  RTN %.4g  GPTQ %.4g  ratio %.3f" % (float(((E_rtn @ H) * E_rtn).sum()), float(((E_g @ H) * E_g).sum()),
          float(((E_g @ H) * E_g).sum() / ((E_rtn @ H) * E_rtn).sum())), flush=True)
    return Qm.reshape(192, 192, 3, 3)
t0 = time.time()
g = {l: gptq(l) for l in (2, 3, 4)}
print("gptq time", time.time() - t0)
report("GPTQ (no lo)", g)
report("GPTQ + lo 14:64:1", with_lo(g))
report("GPTQ + lo 14:32:1", with_lo(g, no=32))

print("---- variants")
def gptq2(l, damp=0.01, order="act", ncal=n_cal, h32=False, verbose=False):
    a = ref_acts[l - 2][:ncal]; c = offs[l - 2]
    xp = torch.empty(ncal, 192, 10, 10, dtype=torch.float64); xp[:] = -c[None, :, None, None]; xp[:, :, 1:-1, 1:-1] = a - c[None, :, None, None]
    X = torch.nn.functional.unfold(xp, 3).permute(0, 2, 1).reshape(-1, 1728)
    if h32:
        X = X.float(); H = ((X.T @ X) / X.shape[0]).double()
    else:
        H = (X.T @ X) / X.shape[0]
    Wm = K[l - 1].reshape(192, 1728).clone()
    perm = torch.argsort(torch.diag(H), descending=True) if order == "act" else torch.arange(1728)
    Hp = H[perm][:, perm] + damp * torch.diag(H).mean() * torch.eye(1728, dtype=torch.float64)
    Wp = Wm[:, perm]
    Hinv = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hp)), upper=True)
    Q = torch.zeros_like(Wp)
    for k in range(1728):
        q = r16(Wp[:, k]); Q[:, k] = q
        err = (Wp[:, k] - q) / Hinv[k, k]
        if k + 1 < 1728: Wp[:, k + 1:] -= err[:, None] * Hinv[k, k + 1:][None, :]
    return Q[:, torch.argsort(perm)].reshape(192, 192, 3, 3)
report("GPTQ natural order", {l: gptq2(l, order="nat") for l in (2, 3, 4)})
report("GPTQ damp 0.001", {l: gptq2(l, damp=0.001) for l in (2, 3, 4)})
report("GPTQ damp 0.1", {l: gptq2(l, damp=0.1) for l in (2, 3, 4)})
report("GPTQ 128 calibration positions", {l: gptq2(l, ncal=128) for l in (2, 3, 4)})
report("GPTQ 256 calibration positions", {l: gptq2(l, ncal=256) for l in (2, 3, 4)})
report("GPTQ fp32 Hessian", {l: gptq2(l, h32=True) for l in (2, 3, 4)})
g = {l: gptq2(l) for l in (2, 3, 4)}
for z in (6, 12, 18, 24):
    offs, pos = scheme(z)
    report("GPTQ (offsets' H reused), zero buckets %d" % z, g) if z == 12 else None
    g2 = {l: gptq2(l) for l in (2, 3, 4)}
    report("GPTQ, zero buckets %d" % z, g2)
