#!/usr/bin/env python
"""Quick start: evaluate Breakthrough positions on a B200 three ways.

    python examples/quickstart.py

1. the reference's call surface (drop-in for MCTS.pyx's NeuralNetsCombined_128),
2. the array API (numpy in / numpy out through the C ABI's host path),
3. the device-resident API (torch CUDA tensors in / out, children ready for the next leaf batch).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wann_b200  # noqa: E402
from wann_b200 import boards, synthetic  # noqa: E402

sq, black = synthetic.random_positions(4096, seed=0)

# 1. reference surface: list of node dicts -> ndarray [N,155] float32 (MCTS.pyx:85-127)
nodes = [{"game_board": boards.squares_to_board(s, 0 if b else 1), "color": "Black" if b else "White"}
         for s, b in zip(sq[:8], black[:8])]
# (no weights given -> $WANN_B200_CHECKPOINT / $WANN_B200_FROZEN_GRAPH; here: the synthetic preset, explicitly)
weights = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
with wann_b200.NeuralNetsCombined_128(weights) as policy_net:
    probs = policy_net.evaluate(nodes)
    best = [boards.move_lookup_by_index(int(i), n["color"]) for i, n in zip(probs.argmax(1), nodes)]
    print("evaluate():", probs.shape, probs.dtype, "argmax moves", best[:4])

# 2. array API: ranked legal children (indices in the reference's 155-move alphabet, priors)
with wann_b200.PolicyEvaluator(wann_b200.weights.synthetic("trained_like", 1), mode="fp16c") as ev:
    idx, prior, n_legal = ev.eval_topk(sq, black)
    print("eval_topk(): policy moves", idx[:4, 0], "priors", np.round(prior[:4, 0], 3), "legal", n_legal[:4])

    # 3. device-resident: children boards stay on the GPU
    import torch
    out = ev.eval_expand(torch.from_numpy(sq).cuda(), torch.from_numpy(black).cuda())
    torch.cuda.synchronize()
    print("eval_expand(): children", tuple(out[3].shape), "on", out[3].device, "| stats", ev.stats())

    # 4. everything the reference derives from the priors when it expands a node (child seeds, flags,
    #    the parent's gained statistics and win status) -- tree_search_utils.update_child & co.
    idx, prior, n_legal, _, flags, stats, parent = ev.eval_expand_seeded(sq[:4], black[:4], with_children=False)
    print("eval_expand_seeded(): first child visits/wins/gameover_visits/gameover_wins", stats[0, 0].tolist(),
          "| parent gained", parent[0, :4].tolist(), "win_status", int(parent[0, 4]))

# 5. many searches at once: array-based trees with the reference's search semantics, one GPU call per lock step
with wann_b200.PolicyEvaluator(wann_b200.weights.synthetic("trained_like", 1), mode="fp16c", refine_margin=0.02) as ev, \
        wann_b200.SearchForest(sq[:256], black[:256]) as forest:
    steps, evals = forest.run(ev, simulations=5)
    info = forest.info(0)
    print("SearchForest: 256 trees, 5 simulations each:", steps, "lock steps,", evals, "evaluations; tree 0 has",
          info["nodes"], "nodes and would play move index", info["move_index"])
