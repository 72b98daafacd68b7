#!/usr/bin/env python
"""GPU bring-up / diagnosis script (development tool, run under gpurun).

Checks the fp32 CUDA-core path and the bf16 tcgen05 path layer by layer against the
oracle, tries the descriptor variants if layer 1 mismatches, and prints quick timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import wann_b200  # noqa: E402
from wann_b200 import weights as W  # noqa: E402


def stats(name, got, ref):
    d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    scale = np.abs(ref).max() + 1e-30
    print("  %-28s max|d|=%.3e  max|d|/max|ref|=%.3e  mean|d|=%.3e  ref max=%.3e  nan=%d" % (
        name, d.max(), d.max() / scale, d.mean(), scale, int(np.isnan(got).sum())), flush=True)
    return d.max() / scale


def main():
    n = int(os.environ.get("BRINGUP_N", "64"))
    sq, bl = O.random_positions_fast(n, seed=11)
    w = W.synthetic("trained_like", 1)
    planes = O.encode_planes(sq, bl)
    lg_ref, pr_ref, acts_ref = O.forward_fp32(planes, w, return_all=True)
    lg_bf, pr_bf, acts_bf = O.forward_bf16_emulated(planes, w, return_all=True)
    mask_ref = O.legal_mask(sq, bl)

    print("== fp32 mode ==", flush=True)
    ev = wann_b200.PolicyEvaluator(w, mode="fp32")
    assert (ev.encode(sq, bl) == planes).all(), "encode mismatch"
    assert (ev.legal_mask(sq, bl) == mask_ref).all(), "legal mask mismatch"
    print("  encode + legal mask bit-exact", flush=True)
    for layer in range(1, 6):
        a = ev.debug_activations(sq, bl, layer)
        stats("layer %d act" % layer, a, acts_ref[layer - 1])
    pr, lg = ev.eval_probs(sq, bl, return_logits=True)
    stats("logits", lg, lg_ref)
    stats("probs", pr, pr_ref)
    idx, pri, cnt = ev.eval_topk(sq, bl)
    ridx, rpri, rcnt = O.ranked_children(pr, mask_ref)
    print("  topk idx equal:", (idx == ridx).all(), "count equal:", (cnt == rcnt).all(),
          "prior max diff", np.abs(pri - rpri).max(), flush=True)
    pr_pl = ev.eval_planes(planes)
    print("  planes(int8) path vs squares path max diff", np.abs(pr_pl - pr).max(), flush=True)
    ev.close()

    print("== bf16 tcgen05 mode ==", flush=True)
    ev = wann_b200.PolicyEvaluator(w, mode="bf16")
    good_variant = None
    for variant in ([0, 1, 2, 3, 4, 5, 6, 7]):
        ev.set_option("variant", variant)
        try:
            t0 = time.time()
            a1 = ev.debug_activations(sq, bl, 1)
            dt = time.time() - t0
        except wann_b200.WannError as e:
            print("  variant %d: ERROR %s" % (variant, e), flush=True)
            continue
        r = stats("variant %d layer 1 (%.2fs)" % (variant, dt), a1, acts_bf[0])
        if r < 2e-2:
            good_variant = variant
            break
    if good_variant is None:
        print("  NO VARIANT MATCHES LAYER 1; dumping a few values for diagnosis", flush=True)
        ev.set_option("variant", 0)
        a1 = ev.debug_activations(sq, bl, 1)
        np.set_printoptions(precision=3, linewidth=200, suppress=True)
        print("got  board0 px0 ch0..15", a1[0, 0, 0, :16])
        print("ref  board0 px0 ch0..15", acts_bf[0][0, 0, 0, :16])
        print("got  board0 px9 ch0..15", a1[0, 1, 1, :16])
        print("ref  board0 px9 ch0..15", acts_bf[0][0, 1, 1, :16])
        print("got  board0 ch0 8x8\n", a1[0, :, :, 0])
        print("ref  board0 ch0 8x8\n", acts_bf[0][0, :, :, 0])
        print("got  board0 ch100 8x8\n", a1[0, :, :, 100])
        print("ref  board0 ch100 8x8\n", acts_bf[0][0, :, :, 100])
        print("got  board1 ch0 8x8\n", a1[1, :, :, 0])
        print("ref  board1 ch0 8x8\n", acts_bf[0][1, :, :, 0])
        print("got  board5 ch0 8x8\n", a1[5, :, :, 0])
        print("ref  board5 ch0 8x8\n", acts_bf[0][5, :, :, 0])
        return 1
    print("  using variant", good_variant, flush=True)
    for layer in range(2, 6):
        a = ev.debug_activations(sq, bl, layer)
        stats("layer %d act vs bf16-emu" % layer, a, acts_bf[layer - 1])
        stats("layer %d act vs fp32" % layer, a, acts_ref[layer - 1])
    pr, lg = ev.eval_probs(sq, bl, return_logits=True)
    stats("logits vs bf16-emu", lg, lg_bf)
    stats("logits vs fp32 oracle", lg, lg_ref)
    stats("probs vs fp32 oracle", pr, pr_ref)
    rel = np.abs(pr - pr_ref) / pr_ref
    print("  probs elementwise rel err: max %.3e median %.3e" % (rel.max(), np.median(rel)), flush=True)
    top_ref = O.top1_legal(pr_ref, mask_ref)
    idx, pri, cnt = ev.eval_topk(sq, bl)
    print("  top-1 agreement vs fp32 oracle: %.4f" % (idx[:, 0] == top_ref).mean(), flush=True)
    ridx, rpri, rcnt = O.ranked_children(pr, mask_ref)
    print("  topk (own probs) idx equal:", (idx == ridx).all(), "count equal:", (cnt == rcnt).all(), flush=True)

    # ---- quick timing, device-resident ----
    import torch
    for nb in (256, 8192, 65536):
        sqb, blb = O.random_positions_fast(min(nb, 4096), seed=5)
        reps = (nb + len(sqb) - 1) // len(sqb)
        sqt = torch.from_numpy(np.tile(sqb, (reps, 1))[:nb]).cuda()
        blt = torch.from_numpy(np.tile(blb, reps)[:nb]).cuda()
        out = None
        for _ in range(3):
            out = ev.eval_topk(sqt, blt, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            out = ev.eval_topk(sqt, blt, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        pos_s = nb / ms * 1e3
        print("  N=%6d  %.3f ms/batch  %.3f M pos/s  (%.1f TFLOP/s conv stack)" % (
            nb, ms, pos_s / 1e6, pos_s * 127401984 / 1e12), flush=True)
    ev.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
