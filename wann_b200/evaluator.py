"""PolicyEvaluator: Python handle on a wann_ctx (one GPU, one set of weights).

numpy arrays go through the HOST path of the C ABI (copies inside the call);
torch CUDA tensors go through the DEVICE path on torch's current stream (no copies,
asynchronous; outputs are torch tensors).  torch is used for buffers/streams only."""
import ctypes

import numpy as np

from . import _lib as L
from . import weights as W


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class PolicyEvaluator(object):
    def __init__(self, weights=None, device=0, mode="bf16", final_kernel=3, chunk=None, coalesce=False,
                 refine_margin=0.0):
        lib = L.load()
        self.final_kernel = final_kernel
        w, cw = self._c_weights(weights, final_kernel)
        self._keep = w
        self.mode = L.MODES[mode] if isinstance(mode, str) else int(mode)
        self.device = int(device)
        ctx = ctypes.c_void_p()
        L.check(lib.wann_create(ctypes.byref(cw), self.device, self.mode, ctypes.byref(ctx)))
        self._ctx = ctx
        self._lib = lib
        if chunk is not None:
            self.set_option("chunk", chunk)
        if coalesce:
            self.set_option("coalesce", 1)
        if refine_margin:
            self.set_refine_margin(refine_margin)

    @staticmethod
    def _c_weights(weights, final_kernel):
        w = W.validate(weights if weights is not None else W.default_weights(final_kernel=final_kernel), final_kernel)
        cw = L.WannWeights()
        for i in range(5):
            cw.conv_w[i] = w["hidden_layer/%d/weights" % (i + 1)].ctypes.data
            cw.conv_b[i] = w["hidden_layer/%d/bias" % (i + 1)].ctypes.data
        cw.fc_w = w["output_layer/weights"].ctypes.data
        cw.fc_b = w["output_layer/bias"].ctypes.data
        cw.final_kernel = final_kernel
        return w, cw

    def set_weights(self, weights):
        """Swap in a newer checkpoint of the same architecture (wann_set_weights): waits for the
        evaluations in flight, later ones use the new weights."""
        w, cw = self._c_weights(weights, self.final_kernel)
        L.check(self._lib.wann_set_weights(self._ctx, ctypes.byref(cw)))
        self._keep = w

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.wann_destroy(self._ctx)
            self._ctx = None
            for p in getattr(self, "_pinned", []):
                self._lib.wann_host_free(p)
            self._pinned = []

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key, value):
        L.check(self._lib.wann_set_option(self._ctx, key.encode(), int(value)))

    def set_refine_margin(self, margin_logits):
        """bf16/fp16 modes: re-evaluate, inside the same call, every position whose two best legal
        moves are closer than `margin_logits` with fp32-class arithmetic (0 switches it off)."""
        self.set_option("refine_margin_micro", int(round(float(margin_logits) * 1e6)))

    def centering(self):
        """fp16c mode: (offsets [4,24] by (position // 64) * 8 + position % 8, position_of [4,192], kappa [9])
        of the calibrated centring."""
        off = np.zeros((4, 24), np.float32)
        pos = np.zeros((4, 192), np.uint8)
        kap = np.zeros(9, np.float32)
        L.check(self._lib.wann_debug_centering(self._ctx, off.ctypes.data, pos.ctypes.data, kap.ctypes.data))
        return off, pos, kap

    def effective_weights(self):
        """bf16 / fp16 / fp16c modes: {layer: float32 [3,3,192,192] HWIO} for hidden_layer/2..4 -- the weights exactly as
        the tensor-core kernel multiplies them (16-bit values; hi + lo where a lo stage covers the entry; with
        the fp16c mode's data-aware rounding applied)."""
        out = {}
        for layer in (2, 3, 4):
            w = np.zeros((3, 3, 192, 192), np.float32)
            L.check(self._lib.wann_debug_effective_weights(self._ctx, layer, w.ctypes.data))
            out[layer] = w
        return out

    def stats(self):
        arr = (ctypes.c_uint64 * 8)()
        L.check(self._lib.wann_get_stats(self._ctx, ctypes.byref(arr)))
        keys = ["positions", "kernel_launches", "calls", "h2d_bytes", "d2h_bytes", "merged_launches",
                "merged_calls", "refined_positions"]
        return dict(zip(keys, list(arr)))

    @staticmethod
    def _np_in(squares, black):
        sq = np.ascontiguousarray(squares, dtype=np.int8).reshape(-1, 64)
        bl = np.ascontiguousarray(black, dtype=np.uint8).reshape(-1)
        if bl.shape[0] != sq.shape[0]:
            raise ValueError("squares/black_to_move length mismatch")
        return sq, bl

    @staticmethod
    def _torch_stream(t):
        import torch
        # torch's default stream has handle 0, which the C ABI reads as "no stream: use the
        # context's own stream and synchronise"; name the legacy default stream explicitly
        # (cudaStreamLegacy == 0x1) so that the call stays asynchronous.
        h = torch.cuda.current_stream(t.device).cuda_stream
        return ctypes.c_void_p(h if h else 1)

    # ------------------------------------------------------------------ byte kernels
    def encode(self, squares, black):
        sq, bl = self._np_in(squares, black)
        out = np.empty((sq.shape[0], 8, 8, 4), np.int8)
        L.check(self._lib.wann_encode(self._ctx, sq.ctypes.data, bl.ctypes.data, sq.shape[0],
                                      out.ctypes.data, L.MEM_HOST, None))
        return out

    def legal_mask(self, squares, black):
        sq, bl = self._np_in(squares, black)
        out = np.empty((sq.shape[0], L.N_MOVES), np.uint8)
        L.check(self._lib.wann_legal_mask(self._ctx, sq.ctypes.data, bl.ctypes.data, sq.shape[0],
                                          out.ctypes.data, L.MEM_HOST, None))
        return out.astype(bool)

    # ------------------------------------------------------------------ evaluation
    def eval_probs(self, squares, black, return_logits=False, out=None):
        if _is_torch(squares):
            import torch
            n = squares.shape[0]
            probs = out if out is not None else torch.empty((n, L.N_MOVES), dtype=torch.float32, device=squares.device)
            logits = torch.empty_like(probs) if return_logits else None
            L.check(self._lib.wann_eval_probs(self._ctx, squares.data_ptr(), black.data_ptr(), n,
                                              probs.data_ptr(), logits.data_ptr() if return_logits else None,
                                              L.MEM_DEVICE, self._torch_stream(squares)))
            return (probs, logits) if return_logits else probs
        sq, bl = self._np_in(squares, black)
        n = sq.shape[0]
        probs = out if out is not None else np.empty((n, L.N_MOVES), np.float32)
        logits = np.empty((n, L.N_MOVES), np.float32) if return_logits else None
        L.check(self._lib.wann_eval_probs(self._ctx, sq.ctypes.data, bl.ctypes.data, n, probs.ctypes.data,
                                          logits.ctypes.data if return_logits else None, L.MEM_HOST, None))
        return (probs, logits) if return_logits else probs

    def eval_planes(self, planes, return_logits=False):
        p = np.asarray(planes)
        if p.dtype == np.int8:
            dt = L.PLANES_I8
        else:
            p = p.astype(np.float32, copy=False)
            dt = L.PLANES_F32
        p = np.ascontiguousarray(p).reshape(-1, 8, 8, 4)
        n = p.shape[0]
        probs = np.empty((n, L.N_MOVES), np.float32)
        logits = np.empty((n, L.N_MOVES), np.float32) if return_logits else None
        L.check(self._lib.wann_eval_planes(self._ctx, p.ctypes.data, dt, n, probs.ctypes.data,
                                           logits.ctypes.data if return_logits else None, L.MEM_HOST, None))
        return (probs, logits) if return_logits else probs

    def eval_topk(self, squares, black, out=None, k=0):
        """-> idx uint8 [n,48], prior float32 [n,48], n_legal uint8 [n] (sorted, descending).
        k: keep the k best legal children only (num_top of get_top_children, board_utils.pyx:279-291);
        0 = all."""
        if _is_torch(squares):
            import torch
            n = squares.shape[0]
            dev = squares.device
            if out is None:
                out = (torch.empty((n, L.MAX_LEGAL), dtype=torch.uint8, device=dev),
                       torch.empty((n, L.MAX_LEGAL), dtype=torch.float32, device=dev),
                       torch.empty((n,), dtype=torch.uint8, device=dev))
            idx, pri, cnt = out
            L.check(self._lib.wann_eval_topk(self._ctx, squares.data_ptr(), black.data_ptr(), n, int(k), idx.data_ptr(),
                                             pri.data_ptr(), cnt.data_ptr(), L.MEM_DEVICE,
                                             self._torch_stream(squares)))
            return idx, pri, cnt
        sq, bl = self._np_in(squares, black)
        n = sq.shape[0]
        if out is None:
            out = (np.empty((n, L.MAX_LEGAL), np.uint8), np.empty((n, L.MAX_LEGAL), np.float32),
                   np.empty(n, np.uint8))
        idx, pri, cnt = out
        L.check(self._lib.wann_eval_topk(self._ctx, sq.ctypes.data, bl.ctypes.data, n, int(k), idx.ctypes.data,
                                         pri.ctypes.data, cnt.ctypes.data, L.MEM_HOST, None))
        return idx, pri, cnt

    def eval_expand(self, squares, black, k=0):
        """eval_topk + device-side child construction.  -> idx, prior, n_legal, child_squares int8
        [n,48,64], child_flags uint8 [n,48] (bit0 win, bit1 capture).  numpy in -> numpy out,
        torch CUDA tensors in -> torch CUDA tensors out (children stay on the device, ready to be
        the next leaf batch)."""
        if _is_torch(squares):
            import torch
            n, dev = squares.shape[0], squares.device
            idx = torch.empty((n, L.MAX_LEGAL), dtype=torch.uint8, device=dev)
            pri = torch.empty((n, L.MAX_LEGAL), dtype=torch.float32, device=dev)
            cnt = torch.empty((n,), dtype=torch.uint8, device=dev)
            kids = torch.empty((n, L.MAX_LEGAL, 64), dtype=torch.int8, device=dev)
            flg = torch.empty((n, L.MAX_LEGAL), dtype=torch.uint8, device=dev)
            L.check(self._lib.wann_eval_expand(self._ctx, squares.data_ptr(), black.data_ptr(), n, int(k), idx.data_ptr(),
                                               pri.data_ptr(), cnt.data_ptr(), kids.data_ptr(), flg.data_ptr(),
                                               L.MEM_DEVICE, self._torch_stream(squares)))
            return idx, pri, cnt, kids, flg
        sq, bl = self._np_in(squares, black)
        n = sq.shape[0]
        idx = np.empty((n, L.MAX_LEGAL), np.uint8)
        pri = np.empty((n, L.MAX_LEGAL), np.float32)
        cnt = np.empty(n, np.uint8)
        kids = np.empty((n, L.MAX_LEGAL, 64), np.int8)
        flg = np.empty((n, L.MAX_LEGAL), np.uint8)
        L.check(self._lib.wann_eval_expand(self._ctx, sq.ctypes.data, bl.ctypes.data, n, int(k), idx.ctypes.data,
                                           pri.ctypes.data, cnt.ctypes.data, kids.ctypes.data, flg.ctypes.data,
                                           L.MEM_HOST, None))
        return idx, pri, cnt, kids, flg

    def eval_expand_seeded(self, squares, black, with_children=True, out=None, k=0):
        """eval_expand + child seeding (wann_eval_expand_seeded).  -> idx, prior, n_legal,
        child_squares (or None), child_flags uint8 [n,48] (6 bits), child_stats int32 [n,48,4]
        (visits, wins, gameover_visits, gameover_wins), parent_summary int32 [n,8].
        `out` (host path): the seven arrays to fill (e.g. pinned ones from pinned_rows), rows [:n] are used."""
        if _is_torch(squares):
            import torch
            n, dev = squares.shape[0], squares.device
            mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)   # noqa: E731
            idx, pri, cnt = mk((n, L.MAX_LEGAL), torch.uint8), mk((n, L.MAX_LEGAL), torch.float32), mk((n,), torch.uint8)
            kids = mk((n, L.MAX_LEGAL, 64), torch.int8) if with_children else None
            flg, st, ps = mk((n, L.MAX_LEGAL), torch.uint8), mk((n, L.MAX_LEGAL, 4), torch.int32), mk((n, 8), torch.int32)
            L.check(self._lib.wann_eval_expand_seeded(
                self._ctx, squares.data_ptr(), black.data_ptr(), n, int(k), idx.data_ptr(), pri.data_ptr(), cnt.data_ptr(),
                kids.data_ptr() if with_children else None, flg.data_ptr(), st.data_ptr(), ps.data_ptr(),
                L.MEM_DEVICE, self._torch_stream(squares)))
            return idx, pri, cnt, kids, flg, st, ps
        sq, bl = self._np_in(squares, black)
        n = sq.shape[0]
        if out is not None:
            idx, pri, cnt, kids, flg, st, ps = [None if a is None else a[:n] for a in out]
            with_children = with_children and kids is not None
        else:
            idx, pri, cnt = np.empty((n, L.MAX_LEGAL), np.uint8), np.empty((n, L.MAX_LEGAL), np.float32), np.empty(n, np.uint8)
            kids = np.empty((n, L.MAX_LEGAL, 64), np.int8) if with_children else None
            flg, st, ps = np.empty((n, L.MAX_LEGAL), np.uint8), np.empty((n, L.MAX_LEGAL, 4), np.int32), np.empty((n, 8), np.int32)
        L.check(self._lib.wann_eval_expand_seeded(
            self._ctx, sq.ctypes.data, bl.ctypes.data, n, int(k), idx.ctypes.data, pri.ctypes.data, cnt.ctypes.data,
            kids.ctypes.data if with_children else None, flg.ctypes.data, st.ctypes.data, ps.ctypes.data,
            L.MEM_HOST, None))
        return idx, pri, cnt, kids, flg, st, ps

    def expand_ranked(self, squares, black, idx, prior, n_legal, with_children=True, with_seeds=True):
        """Child boards / child seeds for positions whose ranked children are already known (the outputs of
        eval_topk): wann_expand_ranked, torch CUDA tensors only.  -> child_squares (or None), child_flags,
        child_stats (or None), parent_summary (or None)."""
        import torch
        n, dev = squares.shape[0], squares.device
        kids = torch.empty((n, L.MAX_LEGAL, 64), dtype=torch.int8, device=dev) if with_children else None
        flg = torch.empty((n, L.MAX_LEGAL), dtype=torch.uint8, device=dev)
        st = torch.empty((n, L.MAX_LEGAL, 4), dtype=torch.int32, device=dev) if with_seeds else None
        ps = torch.empty((n, 8), dtype=torch.int32, device=dev) if with_seeds else None
        L.check(self._lib.wann_expand_ranked(
            self._ctx, squares.data_ptr(), black.data_ptr(), n, idx.data_ptr(), prior.data_ptr(), n_legal.data_ptr(),
            kids.data_ptr() if with_children else None, flg.data_ptr(), st.data_ptr() if with_seeds else None,
            ps.data_ptr() if with_seeds else None, self._torch_stream(squares)))
        return kids, flg, st, ps

    def pinned_rows(self, n, with_children=False):
        """Pinned host arrays for eval_expand_seeded(out=...) over up to n positions (wann_host_alloc):
        the copies of a large host call then run at PCIe speed and nothing is allocated per call.
        The arrays live as long as this evaluator."""
        shapes = [((n, L.MAX_LEGAL), np.uint8), ((n, L.MAX_LEGAL), np.float32), ((n,), np.uint8),
                  ((n, L.MAX_LEGAL, 64), np.int8) if with_children else None,
                  ((n, L.MAX_LEGAL), np.uint8), ((n, L.MAX_LEGAL, 4), np.int32), ((n, 8), np.int32)]
        return tuple(None if sh is None else self.pinned_array(*sh) for sh in shapes)

    def pinned_array(self, shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = ctypes.c_void_p()
        L.check(self._lib.wann_host_alloc(ctypes.byref(ptr), max(nbytes, 1)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def tree_step(self, squares, black, greedy=False, seed=0, step=0, want_outputs=True):
        """In-place self-play step on torch CUDA tensors (see wann_tree_step).  Returns
        (chosen_idx uint8 [n], finished uint8 [n]) device tensors, or None."""
        import torch
        n = squares.shape[0]
        chosen = torch.empty((n,), dtype=torch.uint8, device=squares.device) if want_outputs else None
        fin = torch.empty((n,), dtype=torch.uint8, device=squares.device) if want_outputs else None
        L.check(self._lib.wann_tree_step(self._ctx, squares.data_ptr(), black.data_ptr(), n, int(bool(greedy)),
                                         int(seed), int(step), chosen.data_ptr() if want_outputs else None,
                                         fin.data_ptr() if want_outputs else None, self._torch_stream(squares)))
        return (chosen, fin) if want_outputs else None

    def debug_activations(self, squares, black, layer):
        sq, bl = self._np_in(squares, black)
        n = sq.shape[0]
        c = 1 if layer == 5 else 192
        out = np.empty((n, 8, 8, c), np.float32)
        L.check(self._lib.wann_debug_activations(self._ctx, sq.ctypes.data, bl.ctypes.data, n, int(layer),
                                                 out.ctypes.data, L.MEM_HOST))
        return out

    def debug_profile(self, squares, black):
        """clock64 stamps [4,64] of CTA pair 0 (MMA start/end, epilogue start/end per layer-step)."""
        sq, bl = self._np_in(squares, black)
        out = np.zeros((8, 64), np.int64)
        L.check(self._lib.wann_debug_profile(self._ctx, sq.ctypes.data, bl.ctypes.data, sq.shape[0],
                                             out.ctypes.data))
        return out

    def kernel_times(self):
        """(summed conv-stack kernel ms, launches) since the last call; needs option time_kernels=1."""
        out = (ctypes.c_double * 2)()
        L.check(self._lib.wann_kernel_times(self._ctx, ctypes.byref(out)))
        return float(out[0]), int(out[1])
