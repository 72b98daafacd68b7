"""MultiGpuEvaluator: ONE process driving several B200s (the reference is a single Python process,
so this is what a maintainer gets without torchrun).  The path shards by position -- NN_output[i]
depends only on input i (tree_builder.pyx:318) -- so a host batch is cut into contiguous slices
(sharding.shard_range), one per GPU, each evaluated by that GPU's own context on its own host
thread (the C ABI releases the GIL and is re-entrant), results written straight into the slices of
one output array.  No collective, no peer copies; weights are replicated (4 MB per GPU).

Small batches (< min_per_gpu positions per GPU) are not worth splitting and go to one GPU, chosen
round-robin so that concurrent small callers (the reference's 11 search threads) spread over the
devices."""
import itertools
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib as L
from .evaluator import PolicyEvaluator
from .sharding import shard_range


def plan_shards(n, n_devices, min_per_gpu=512):
    """-> list of (device slot, lo, hi) covering [0, n): contiguous, sizes differ by <= 1, and no
    more devices than keep every slice >= min_per_gpu (at least one)."""
    n = int(n)
    if n <= 0:
        return []
    use = max(1, min(int(n_devices), n // max(1, int(min_per_gpu))))
    return [(r,) + shard_range(n, r, use) for r in range(use)]


class MultiGpuEvaluator(object):
    def __init__(self, weights=None, devices=None, min_per_gpu=512, evaluator_factory=None, **kw):
        """devices: list of CUDA ordinals (default: all visible).  **kw goes to every PolicyEvaluator
        (mode, refine_margin, coalesce, ...).  evaluator_factory(device) is a test seam."""
        if devices is None:
            import torch
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise L.WannError("no CUDA device: wann_b200 has no CPU fallback")
        self.devices = list(devices)
        self.min_per_gpu = int(min_per_gpu)
        make = evaluator_factory or (lambda d: PolicyEvaluator(weights, device=d, **kw))
        self.evs = [make(d) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.evs), thread_name_prefix="wann-gpu")
        self._rr = itertools.count()
        self._rr_lock = threading.Lock()

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_pool", None) is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
            for ev in self.evs:
                ev.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_weights(self, weights):
        for ev in self.evs:
            ev.set_weights(weights)

    def set_refine_margin(self, margin_logits):
        for ev in self.evs:
            ev.set_refine_margin(margin_logits)

    def stats(self):
        per = [ev.stats() for ev in self.evs]
        return {k: sum(p[k] for p in per) for k in per[0]}

    def _plan(self, n):
        shards = plan_shards(n, len(self.evs), self.min_per_gpu)
        if len(shards) == 1:                       # one GPU: spread successive small calls round-robin
            with self._rr_lock:
                slot = next(self._rr) % len(self.evs)
            shards = [(slot, 0, n)]
        return shards

    def _run(self, n, fn):
        shards = self._plan(n)
        if len(shards) == 1:
            slot, lo, hi = shards[0]
            fn(self.evs[slot], lo, hi)
            return
        futs = [self._pool.submit(fn, self.evs[slot], lo, hi) for slot, lo, hi in shards]
        for f in futs:
            f.result()                              # re-raises a worker's WannError here

    # ------------------------------------------------------------------ the evaluator surface (host arrays)
    def eval_probs(self, squares, black, return_logits=False, out=None):
        sq, bl = PolicyEvaluator._np_in(squares, black)
        n = sq.shape[0]
        probs = out if out is not None else np.empty((n, L.N_MOVES), np.float32)
        logits = np.empty((n, L.N_MOVES), np.float32) if return_logits else None

        def work(ev, lo, hi):
            if return_logits:
                p, lg = ev.eval_probs(sq[lo:hi], bl[lo:hi], return_logits=True, out=probs[lo:hi])
                logits[lo:hi] = lg
            else:
                ev.eval_probs(sq[lo:hi], bl[lo:hi], out=probs[lo:hi])
        self._run(n, work)
        return (probs, logits) if return_logits else probs

    def eval_planes(self, planes):
        p = np.ascontiguousarray(planes).reshape(-1, 8, 8, 4)
        n = p.shape[0]
        probs = np.empty((n, L.N_MOVES), np.float32)

        def work(ev, lo, hi):
            probs[lo:hi] = ev.eval_planes(p[lo:hi])
        self._run(n, work)
        return probs

    def eval_topk(self, squares, black, out=None, k=0):
        sq, bl = PolicyEvaluator._np_in(squares, black)
        n = sq.shape[0]
        if out is None:
            out = (np.empty((n, L.MAX_LEGAL), np.uint8), np.empty((n, L.MAX_LEGAL), np.float32),
                   np.empty(n, np.uint8))
        idx, pri, cnt = out

        def work(ev, lo, hi):
            ev.eval_topk(sq[lo:hi], bl[lo:hi], out=(idx[lo:hi], pri[lo:hi], cnt[lo:hi]), k=k)
        self._run(n, work)
        return idx, pri, cnt

    def eval_expand_seeded(self, squares, black, with_children=True, k=0):
        sq, bl = PolicyEvaluator._np_in(squares, black)
        n = sq.shape[0]
        parts, lock = {}, threading.Lock()

        def work(ev, lo, hi):
            res = ev.eval_expand_seeded(sq[lo:hi], bl[lo:hi], with_children=with_children, k=k)
            with lock:
                parts[lo] = res
        self._run(n, work)
        cols = list(zip(*[parts[lo] for lo in sorted(parts)]))
        return tuple(None if c[0] is None else np.concatenate(c) for c in cols)
