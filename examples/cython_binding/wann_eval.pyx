# cython: language_level=3, boundscheck=False
"""The binding a WaNN maintainer would add next to monte_carlo_tree_search/MCTS.pyx: a Cython class with
the surface of NeuralNetsCombined_128 (MCTS.pyx:76-133) calling the C ABI directly -- no ctypes, no
Python objects on the call path, the GIL released for the duration of the evaluation (the reference's
10 search threads keep running, tree_builder.pyx:186-189).

Build:  python examples/cython_binding/setup.py build_ext --inplace
"""
cimport wann_b200 as W
import numpy as np
cimport numpy as cnp
from libc.stdint cimport int8_t, uint8_t, int64_t

cnp.import_array()

cdef dict _VALUE = {'e': 0, 'w': 1, 'b': 2}
cdef str _FILES = 'abcdefgh'


def abi_version():
    return W.wann_version()


cdef class PolicyNet:
    """evaluate(game_nodes) -> ndarray [N,155] float32, rows in input order (tree_builder.pyx:318)."""
    cdef W.wann_ctx* ctx
    cdef object keep            # the weight arrays must outlive wann_create's copy only; kept for set_weights

    def __cinit__(self):
        self.ctx = NULL

    def __init__(self, dict weights, int device=0, int mode=W.WANN_MODE_FP16C, int final_kernel=1,
                 double refine_margin=0.02):
        cdef W.wann_weights w
        cdef cnp.ndarray a
        self.keep = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
        for i in range(5):
            a = self.keep["hidden_layer/%d/weights" % (i + 1)]
            w.conv_w[i] = <const float*> a.data
            a = self.keep["hidden_layer/%d/bias" % (i + 1)]
            w.conv_b[i] = <const float*> a.data
        a = self.keep["output_layer/weights"]
        w.fc_w = <const float*> a.data
        a = self.keep["output_layer/bias"]
        w.fc_b = <const float*> a.data
        w.final_kernel = final_kernel
        if W.wann_create(&w, device, mode, &self.ctx) != W.WANN_OK:
            raise RuntimeError(W.wann_last_error().decode())
        W.wann_set_option(self.ctx, b"coalesce", 1)
        if refine_margin > 0:
            W.wann_set_option(self.ctx, b"refine_margin_micro", <int64_t> (refine_margin * 1e6))

    def evaluate(self, list game_nodes, player_color=None, is_player=True, already_converted=False):
        cdef int64_t n = len(game_nodes), i
        cdef cnp.ndarray[int8_t, ndim=2] sq = np.empty((n, 64), np.int8)
        cdef cnp.ndarray[uint8_t, ndim=1] bl = np.empty(n, np.uint8)
        cdef cnp.ndarray[float, ndim=2] probs = np.empty((n, W.WANN_N_MOVES), np.float32)
        cdef dict board
        cdef int r, c, rc
        for i in range(n):                       # dict-of-dicts board (tools/utils.pyx:518-541) -> 64 bytes
            board = game_nodes[i]['game_board']
            for r in range(8):
                row = board[r + 1]
                for c in range(8):
                    sq[i, r * 8 + c] = _VALUE[row[_FILES[c]]]
            bl[i] = game_nodes[i]['color'] == 'Black'
        with nogil:
            rc = W.wann_eval_probs(self.ctx, <const int8_t*> sq.data, <const uint8_t*> bl.data, n,
                                   <float*> probs.data, NULL, W.WANN_MEM_HOST, NULL)
        if rc != W.WANN_OK:
            raise RuntimeError(W.wann_last_error().decode())
        return probs

    def close(self):
        if self.ctx != NULL:
            W.wann_destroy(self.ctx)
            self.ctx = NULL

    def __dealloc__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):   # MCTS.pyx:132-133
        self.close()
