#!/usr/bin/env python
"""Throughput of the native search forest (BASELINE config 5 with the reference's search semantics per
tree): T trees advancing in lock step, one wann_eval_expand_seeded call per step.  Reports evaluations
per second, tree nodes per second and where the wall time goes (tree code / evaluator), next to the
reference's own search on the same evaluator (profiles/r01_reference_mcts_end_to_end.json: 1.46 k
evaluations/s, GIL-bound).  Development tool, run under gpurun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wann_b200
from wann_b200 import weights as W, synthetic

sims = int(os.environ.get("SIMS", "20"))
w = W.synthetic("trained_like", 1)
rep = {"simulations_per_tree": sims, "host_cores": os.cpu_count(), "mode": "fp16 + 0.02 margin refinement"}
with wann_b200.PolicyEvaluator(w, mode="fp16", refine_margin=0.02) as ev:
    for T in (1, 64, 1024, 8192):
        sq = np.tile(synthetic.initial_squares()[None], (T, 1))
        bl = np.zeros(T, np.uint8)
        if T > 1:                                   # different games: random positions from play-outs
            rs, rb = synthetic.random_positions(T - 1, seed=3)
            sq[1:], bl[1:] = rs, rb
        forest = wann_b200.SearchForest(sq, bl)
        rows = ev.pinned_rows(T)
        t_next = t_eval = t_apply = 0.0
        steps = evals = 0
        t0 = time.perf_counter()
        while True:
            a = time.perf_counter()
            s, b, _ = forest.next(sims)
            c = time.perf_counter()
            t_next += c - a
            if len(s) == 0:
                break
            out = ev.eval_expand_seeded(s, b, with_children=False, out=rows)
            d = time.perf_counter()
            t_eval += d - c
            forest.apply(out[0], out[1], out[2], out[4], out[5], out[6])
            t_apply += time.perf_counter() - d
            steps += 1
            evals += len(s)
        wall = time.perf_counter() - t0
        nodes = sum(forest.info(t)["nodes"] for t in range(T))
        rep[str(T)] = {"trees": T, "steps": steps, "evaluations": evals, "nodes": nodes, "wall_s": round(wall, 3),
                       "evaluations_per_s": round(evals / wall, 1), "nodes_per_s": round(nodes / wall, 1),
                       "mean_batch": round(evals / max(steps, 1), 1),
                       "time_share": {"tree_next": round(t_next / wall, 3), "evaluator": round(t_eval / wall, 3),
                                      "tree_apply": round(t_apply / wall, 3)}}
        forest.close()
        print(T, rep[str(T)], flush=True)
    # two half-forests on two host threads: one's tree code overlaps the other's evaluation (the C ABI
    # releases the GIL and is re-entrant)
    import threading
    for T in (1024, 8192):
        rs, rb = synthetic.random_positions(T, seed=3)
        halves = [wann_b200.SearchForest(rs[h::2], rb[h::2]) for h in range(2)]
        res = [None, None]

        def work(h):
            res[h] = halves[h].run(ev, sims)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(h,)) for h in range(2)]
        [t.start() for t in th]
        [t.join() for t in th]
        wall = time.perf_counter() - t0
        evals = res[0][1] + res[1][1]
        nodes = sum(f.info(t)["nodes"] for f in halves for t in range(f.n))
        rep["%d_two_half_forests" % T] = {"trees": T, "evaluations": evals, "nodes": nodes, "wall_s": round(wall, 3),
                                          "evaluations_per_s": round(evals / wall, 1), "nodes_per_s": round(nodes / wall, 1)}
        print(T, "two half forests", rep["%d_two_half_forests" % T], flush=True)
        [f.close() for f in halves]
    # self-play: search, play the chosen move (subtree reused), restart finished games
    for T, spm, plies in ((1024, 8, 40), (8192, 8, 20)):
        rs, rb = synthetic.random_positions(T, seed=5)          # games in progress: they end (and restart) at different times
        forest = wann_b200.SearchForest(rs, rb)
        t0 = time.perf_counter()
        evals, games = forest.self_play(ev, spm, plies)
        wall = time.perf_counter() - t0
        nodes = sum(forest.info(t)["nodes"] for t in range(T))
        rep["self_play_%d" % T] = {"trees": T, "simulations_per_move": spm, "plies": plies, "moves_played": T * plies,
                                   "games_finished": games, "evaluations": evals, "wall_s": round(wall, 3),
                                   "evaluations_per_s": round(evals / wall, 1), "moves_per_s": round(T * plies / wall, 1),
                                   "nodes_alive_at_end": nodes}
        print("self-play", rep["self_play_%d" % T], flush=True)
        forest.close()
open(os.path.join(ROOT, "gpurun_out", "forest.json"), "w").write(json.dumps(rep, indent=1))
