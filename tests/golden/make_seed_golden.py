#!/usr/bin/env python
"""Golden vectors for CHILD SEEDING, produced by the compiled reference tree code (oracle/_ref).

For each position the reference's own expansion step runs once --
  tree_builder.expand_descendants_to_depth_wrt_NN([root], ...)      tree_builder.pyx:122-215
    -> enumerate_update_and_prune / get_child / assign_children      :428-445, :530-620
    -> check_for_winning_move / set_game_over_values(_1_step_lookahead)   :902-921
    -> tree_search_utils.update_child                                tree_search_utils.pyx:907-1017
    -> eval_children / evaluation_function                           :812-858, :1056-1071
    -> update_win_status_from_children                               :86-132
-- with a stub policy net that returns a seeded synthetic NN output (committed below) and stops
the search after that one expansion.  Recorded: every child's seeded statistics and flags, the
reference's child ORDER, and the statistics / win status the (parentless) root ends up with.

Run in the build container only:  python tests/golden/make_seed_golden.py
Output: tests/golden/ref_seed_golden.npz (committed)."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_ref  # noqa: E402

FILES = "abcdefgh"
V = {"e": 0, "w": 1, "b": 2}
WS = {None: -1, False: 0, True: 1}


def to_squares(board):
    return np.array([V[board[r][c]] for r in range(1, 9) for c in FILES], np.int8)


def apply_move(B, board, move, color):
    nb = B.copy_game_board(board)
    f, t = move.split("-")
    nb[int(t[1])][t[0]] = board[int(f[1])][f[0]]
    nb[int(f[1])][f[0]] = "e"
    nb[9] = 0 if color == "White" else 1
    return nb


class OneShotNet(object):
    """policy_net stand-in: the first evaluate() returns the prepared output, the second one ends
    the search (do_updates_in_the_same_process re-reads sim_info['start_time'], tree_builder.pyx:308)."""

    def __init__(self, out, sim_info):
        self.out, self.sim_info, self.calls = out, sim_info, 0

    def evaluate(self, nodes, **kw):
        self.calls += 1
        if self.calls > 1:
            self.sim_info["start_time"] = 0.0
        return np.repeat(self.out[None], len(nodes), axis=0)


def main():
    ref = load_ref()
    assert ref is not None, "build oracle/_ref first: python oracle/build_ref.py"
    U, B = ref.utils, ref.board_utils
    tb = ref.module("monte_carlo_tree_search.tree_builder")
    TreeNode = ref.module("monte_carlo_tree_search.TreeNode").TreeNode
    rng = np.random.RandomState(20261019)
    cases = [(U.initial_game_board(), "White"), (U.initial_game_board(), "Black"),
             (B.debug_game_board(), "White"), (B.debug_game_board(), "Black")]
    while len(cases) < 640:
        board, color = U.initial_game_board(), "White"
        traj = []
        while not B.game_over(board)[0]:
            legal = B.enumerate_legal_moves(board, color)
            if not legal:
                break
            traj.append((board, color))
            board = apply_move(B, board, legal[rng.randint(len(legal))], color)
            color = "Black" if color == "White" else "White"
        # whole-game spread plus the last plies (threats, winning moves, game-saving captures)
        picks = set(rng.choice(len(traj), size=min(6, len(traj)), replace=False).tolist())
        picks.update(range(max(0, len(traj) - 4), len(traj)))
        for k in sorted(picks):
            cases.append(traj[k])
    n = len(cases)
    squares = np.zeros((n, 64), np.int8)
    black = np.zeros(n, np.uint8)
    nn = np.zeros((n, 155), np.float32)
    c_idx = np.full((n, 48), 255, np.uint8)          # children in ascending move-index order
    c_stats = np.zeros((n, 48, 4), np.int64)         # visits, wins, gameover_visits, gameover_wins
    c_gameover = np.zeros((n, 48), np.uint8)
    c_win = np.full((n, 48), -1, np.int8)
    c_uct = np.zeros((n, 48), np.float64)
    order = np.full((n, 48), 255, np.uint8)          # the reference's child order (sorted by UCT_multiplier)
    n_children = np.zeros(n, np.uint8)
    root_stats = np.zeros((n, 4), np.int64)
    root_win = np.full(n, -1, np.int8)
    for k, (board, color) in enumerate(cases):
        squares[k] = to_squares(board)
        black[k] = color == "Black"
        lg = rng.standard_normal(155) * 2.5
        if k % 5 == 0:                                # some peaked outputs: priors > 0.25 / > 0.30 branches
            lg[rng.randint(154)] += 9.0
        e = np.exp(lg - lg.max())
        nn[k] = (e / e.sum()).astype(np.float32)
        wp = [c + str(r) for r in range(1, 9) for c in FILES if board[r][c] == "w"]
        bp = [c + str(r) for r in range(1, 9) for c in FILES if board[r][c] == "b"]
        root = TreeNode(B.copy_game_board(board), wp, bp, color, 0, None, 0)
        sim_info = {"root": root, "time_to_think": 1000.0, "start_time": time.time(), "game_tree": [],
                    "main_pid": "not-this-thread", "do_eval": False, "counter": 0, "prev_game_tree_size": 0,
                    "game_num": 0, "eval_times": []}
        net = OneShotNet(nn[k], sim_info)
        tb.expand_descendants_to_depth_wrt_NN([root], 0, 1, sim_info, threading.Lock(), net)
        kids = root["children"]
        assert kids is not None and net.calls >= 1
        assert all(kid["children"] is None for kid in kids), "search went deeper than one expansion"
        n_children[k] = len(kids)
        order[k, :len(kids)] = [kid["index"] for kid in kids]
        for slot, kid in enumerate(sorted(kids, key=lambda c: c["index"])):
            c_idx[k, slot] = kid["index"]
            c_stats[k, slot] = [kid["visits"], kid["wins"], kid["gameover_visits"], kid["gameover_wins"]]
            c_gameover[k, slot] = bool(kid["gameover"])
            c_win[k, slot] = WS[kid["win_status"]]
            c_uct[k, slot] = kid["UCT_multiplier"]
        root_stats[k] = [root["visits"], root["wins"], root["gameover_visits"], root["gameover_wins"]]
        root_win[k] = WS[root["win_status"]]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_seed_golden.npz")
    np.savez_compressed(out, squares=squares, black=black, nn=nn, c_idx=c_idx, c_stats=c_stats,
                        c_gameover=c_gameover, c_win=c_win, c_uct=c_uct, order=order, n_children=n_children,
                        root_stats=root_stats, root_win=root_win)
    threat = (c_stats[:, :, 0] == 65536).any(1) | (c_stats[:, :, 0] == 65537).any(1)
    print("wrote", out, "cases", n, "with 65536-class children", int(threat.sum()),
          "root win statuses", np.bincount(root_win + 1, minlength=3).tolist(),
          "winning children", int(c_gameover.sum()), os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
