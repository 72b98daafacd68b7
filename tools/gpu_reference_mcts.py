#!/usr/bin/env python
"""End to end through the REFERENCE'S OWN search: expansion_MCTS_functions.MCTS_with_expansions
(ThreadPool(10) + main thread, compiled from the reference's .pyx into oracle/_ref) thinking about
the opening position for a fixed time on top of
  (a) wann_b200.NeuralNetsCombined_128 (this repo, shim defaults, coalescer on), and
  (b) a CPU stand-in session: the same graph as torch-CPU ops behind the same evaluate() surface
      (oracle/torch_oracle.py; TensorFlow 1.0 is not installable here).
Reports policy-net evaluations per second of search and the size of the tree that was built.
Development tool (needs oracle/_ref and a GPU): run under gpurun."""
import io, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import boards
from oracle.ref_loader import load_ref
from oracle import oracle as O, torch_oracle as T

ref = load_ref()
assert ref is not None, "oracle/_ref not built"
mc = ref.module("monte_carlo_tree_search.expansion_MCTS_functions")
think = float(os.environ.get("THINK", "4"))
w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)


class Counting(object):
    def __init__(self, fn):
        self.fn, self.calls, self.positions, self.lock, self.secs = fn, 0, 0, threading.Lock(), 0.0

    def evaluate(self, nodes, **kw):
        t0 = time.perf_counter()
        out = self.fn(nodes, **kw)
        dt = time.perf_counter() - t0
        with self.lock:
            self.calls += 1
            self.positions += len(nodes)
            self.secs += dt
        return out


def cpu_evaluate(nodes, **kw):
    # the reference's own encode (tools/utils.pyx:546-602) in front of the CPU graph, like MCTS.pyx:99-126
    planes = np.array([ref.utils.convert_board_to_2d_matrix_POEB(n["game_board"], n["color"]) for n in nodes],
                      dtype=np.float32)
    return T.forward(planes, w1, threads=os.cpu_count())[1]


def count_nodes(root):
    n, stack = 0, [root]
    while stack:
        node = stack.pop()
        n += 1
        if node["children"]:
            stack.extend(node["children"])
    return n


def search(net):
    board = ref.utils.initial_game_board()
    t0 = time.perf_counter()
    best_child, move = mc.MCTS_with_expansions(board, "White", think, 5, None, None, 0, io.StringIO(), "WaNN", net, 0)
    wall = time.perf_counter() - t0
    root = best_child["parent"]
    return {"move": move, "wall_s": round(wall, 2), "evaluate_calls": net.calls, "positions": net.positions,
            "evals_per_s": round(net.positions / wall, 1), "mean_ms_per_call": round(1e3 * net.secs / max(net.calls, 1), 3),
            "tree_nodes": count_nodes(root)}


rep = {"think_s": think, "host_cores": os.cpu_count(), "search": "reference MCTS_with_expansions, opening position, White"}
rep["cpu_stand_in"] = search(Counting(cpu_evaluate))
with wann_b200.NeuralNetsCombined_128(w1) as net:
    net.evaluate([{"game_board": ref.utils.initial_game_board(), "color": "White"}])     # warm-up
    rep["wann_b200"] = search(Counting(net.evaluate))
    rep["wann_b200"]["stats"] = net.sess.evaluator("white_net").stats()
rep["speedup_evals_per_s"] = round(rep["wann_b200"]["evals_per_s"] / rep["cpu_stand_in"]["evals_per_s"], 2)
print(json.dumps(rep, indent=1))
open(os.path.join(ROOT, "gpurun_out", "reference_mcts.json"), "w").write(json.dumps(rep, indent=1))
