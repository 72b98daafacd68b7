"""SearchForest: many independent search trees with the reference's own search semantics, advancing
in lock step on one evaluator (BASELINE config 5).  Python handle on a wann_forest (include/wann_b200.h):
the trees live in the C++ runtime as flat arrays; each step

    positions = forest.next(max_simulations)        # every tree runs to its next policy_net.evaluate
    rows = evaluator.eval_expand_seeded(positions)   # ONE batched GPU call for all of them
    forest.apply(rows)                               # children + seeds attached, statistics propagated

What a tree does between two evaluations is exactly what one thread of the reference's search does
(expansion_MCTS_functions.pyx:300-350, tree_search_utils.pyx, tree_builder.pyx), see array_tree.hpp.

SearchForest(..., evaluator=ev, device_resident=True) keeps the trees in the evaluator's GPU memory:
`run` is then ONE C call (wann_forest_run) whose lock steps -- tree kernels, conv stack, child seeding --
are enqueued on one stream with no host round trip; the trees are the same, node for node."""
import ctypes

import numpy as np

from . import _lib as L

DUMP_FIELDS = ("parent", "index", "black", "visits", "wins", "gameover_visits", "gameover_wins", "win_status",
               "gameover", "subtree_checked", "subtree_being_checked", "threads_checking_node", "n_children", "height")


def host_threads(n=0):
    """Host threads the forest's tree code uses (OpenMP over trees); n > 0 sets it.  Launchers such as
    torchrun export OMP_NUM_THREADS=1: call host_threads(os.cpu_count() // world_size) there."""
    return int(L.load().wann_host_threads(int(n)))


class SearchForest(object):
    def __init__(self, squares, black_to_move, heights=None, evaluator=None, device_resident=False,
                 node_capacity=16384):
        self._lib = L.load()
        sq = np.ascontiguousarray(squares, np.int8).reshape(-1, 64)
        bl = np.ascontiguousarray(black_to_move, np.uint8).reshape(-1)
        self.n = sq.shape[0]
        h = None if heights is None else np.ascontiguousarray(heights, np.int32).reshape(-1)
        ptr = ctypes.c_void_p()
        self.device_resident = bool(device_resident)
        self._owner = evaluator                      # a device forest lives in (and dies before) its evaluator's context
        if self.device_resident:
            if evaluator is None or not hasattr(evaluator, "_ctx"):
                raise ValueError("a device-resident forest needs the PolicyEvaluator whose GPU holds it")
            L.check(self._lib.wann_forest_create_device(evaluator._ctx, self.n, sq.ctypes.data, bl.ctypes.data,
                                                        None if h is None else h.ctypes.data, int(node_capacity),
                                                        ctypes.byref(ptr)))
        else:
            L.check(self._lib.wann_forest_create(self.n, sq.ctypes.data, bl.ctypes.data,
                                                 None if h is None else h.ctypes.data, ctypes.byref(ptr)))
        self._f = ptr
        self._sq = np.empty((self.n, 64), np.int8)
        self._bl = np.empty(self.n, np.uint8)
        self._ids = np.empty(self.n, np.int64)

    def close(self):
        if getattr(self, "_f", None):
            self._lib.wann_forest_destroy(self._f)
            self._f = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def next(self, max_simulations):
        """-> (squares int8 [m,64], black uint8 [m], tree_ids int64 [m]) of the trees that need an
        evaluation; m == 0 when every tree has done max_simulations simulations (or is solved)."""
        m = ctypes.c_int64()
        L.check(self._lib.wann_forest_next(self._f, int(max_simulations), self._sq.ctypes.data, self._bl.ctypes.data,
                                           self._ids.ctypes.data, ctypes.byref(m)))
        m = m.value
        return self._sq[:m], self._bl[:m], self._ids[:m]

    def apply(self, idx, prior, n_legal, child_flags, child_stats, parent_summary):
        a = [np.ascontiguousarray(idx, np.uint8), np.ascontiguousarray(prior, np.float32),
             np.ascontiguousarray(n_legal, np.uint8), np.ascontiguousarray(child_flags, np.uint8),
             np.ascontiguousarray(child_stats, np.int32), np.ascontiguousarray(parent_summary, np.int32)]
        L.check(self._lib.wann_forest_apply(self._f, *[x.ctypes.data for x in a]))

    def run(self, evaluator, simulations):
        """Grow every tree by `simulations` simulations on `evaluator` (anything with
        eval_expand_seeded(squares, black, with_children=False)).  Returns (steps, evaluations)."""
        if self.device_resident or (hasattr(evaluator, "_ctx") and getattr(self, "native_run", True)):
            # the whole search in the library: wann_forest_run (device forests: no host round trip per step)
            out2 = (ctypes.c_int64 * 2)()
            L.check(self._lib.wann_forest_run(evaluator._ctx, self._f, int(simulations), out2))
            return int(out2[0]), int(out2[1])
        steps = evals = 0
        out = None
        if hasattr(evaluator, "pinned_rows"):            # pinned result rows, allocated once per evaluator
            cache = evaluator.__dict__.setdefault("_forest_rows", {})
            if cache.get("n", 0) < self.n:
                cache.update(n=self.n, rows=evaluator.pinned_rows(self.n))
            out = cache["rows"]
        while True:
            sq, bl, _ = self.next(simulations)
            if len(sq) == 0:
                return steps, evals
            if out is not None:
                idx, pri, cnt, _, flg, st, ps = evaluator.eval_expand_seeded(sq, bl, with_children=False, out=out)
            else:
                idx, pri, cnt, _, flg, st, ps = evaluator.eval_expand_seeded(sq, bl, with_children=False)
            self.apply(idx, pri, cnt, flg, st, ps)
            steps += 1
            evals += len(sq)

    def run_steps(self, evaluator, lock_steps, simulations=1 << 60):
        """Exactly `lock_steps` lock steps (fewer if every tree finishes `simulations` first), trees may be
        left mid-simulation (wann_forest_run_steps): the steady state of BASELINE config 5, every tree
        contributing a leaf per step.  -> (steps, evaluations)."""
        out2 = (ctypes.c_int64 * 2)()
        L.check(self._lib.wann_forest_run_steps(evaluator._ctx, self._f, int(simulations), int(lock_steps), out2))
        return int(out2[0]), int(out2[1])

    def run_for(self, evaluator, seconds):
        """Time-budgeted search, the reference's `time_to_think` (MCTS_with_expansions,
        expansion_MCTS_functions.pyx:113-119): lock steps until `seconds` have passed, then the
        simulations in flight are finished (no new ones start), so that every tree is between two
        simulations when this returns.  -> (steps, evaluations)."""
        import time
        t_end = time.perf_counter() + float(seconds)
        steps = evals = 0
        if self.device_resident or (hasattr(evaluator, "_ctx") and getattr(self, "native_run", True)):
            while time.perf_counter() < t_end:
                s, e = self.run_steps(evaluator, 32)
                steps, evals = steps + s, evals + e
                if s == 0:
                    break
            s, e = self.run(evaluator, 0)                        # drain: finish what is in flight, start nothing new
            return steps + s, evals + e
        budget = 1 << 60
        while True:
            if budget and time.perf_counter() >= t_end:
                budget = 0                                   # drain: continue, but start nothing new
            sq, bl, _ = self.next(budget)
            if len(sq) == 0:
                if budget == 0:
                    return steps, evals
                budget = 0
                continue
            idx, pri, cnt, _, flg, st, ps = evaluator.eval_expand_seeded(sq, bl, with_children=False)
            self.apply(idx, pri, cnt, flg, st, ps)
            steps += 1
            evals += len(sq)

    def advance(self, tree, move_index):
        """Root reuse: the child reached by `move_index` becomes the root of tree's next search."""
        L.check(self._lib.wann_forest_advance(self._f, int(tree), int(move_index)))

    def play(self, restart_squares=None, restart_black=0):
        """Every tree plays the move its search chose and continues from that child (subtree reused); a
        winning move restarts the tree from `restart_squares` (default: the opening position).
        -> (moves uint8 [n], finished uint8 [n])."""
        from . import synthetic
        rs = np.ascontiguousarray(synthetic.initial_squares() if restart_squares is None else restart_squares, np.int8)
        moves, fin = np.empty(self.n, np.uint8), np.empty(self.n, np.uint8)
        L.check(self._lib.wann_forest_play(self._f, rs.ctypes.data, int(restart_black), moves.ctypes.data, fin.ctypes.data))
        return moves, fin

    def self_play(self, evaluator, simulations_per_move, plies):
        """`plies` rounds of: search `simulations_per_move` simulations in every tree, play the chosen
        moves.  Returns (evaluations, games finished)."""
        evals = games = 0
        for _ in range(plies):
            evals += self.run(evaluator, simulations_per_move)[1]
            games += int(self.play()[1].sum())
        return evals, games

    def debug_cycles(self):
        """device-resident forests: the last 64 lock steps, int64 [64, 5] = rows requested, cycles in the
        next kernel (sum over trees, max), cycles in the apply kernel (sum, max)."""
        out = np.zeros((64, 5), np.int64)
        L.check(self._lib.wann_forest_debug_cycles(self._f, out.ctypes.data))
        return out

    def info(self, tree):
        out = (ctypes.c_int64 * 8)()
        L.check(self._lib.wann_forest_info(self._f, int(tree), out))
        keys = ("nodes", "simulations", "evaluations", "root_visits", "root_wins", "root_win_status", "finished",
                "move_index")
        return dict(zip(keys, list(out)))

    def dump(self, tree):
        n = self._lib.wann_forest_dump(self._f, int(tree), None, 0)
        if n < 0:
            raise L.WannError("bad tree index")
        rows = np.empty((n, 14), np.int64)
        self._lib.wann_forest_dump(self._f, int(tree), rows.ctypes.data, n)
        return rows
