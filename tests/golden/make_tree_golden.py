#!/usr/bin/env python
"""Golden TREES, grown by the compiled reference search (oracle/_ref), single-threaded and therefore
deterministic: for a few start positions, K calls of
  expansion_MCTS_functions.run_MCTS_with_expansions_simulation            (:300-316)
    -> select_UCT_child / choose_UCT_or_best_child / find_best_UCT_child   (:339-350, tree_search_utils.pyx:251-403)
    -> expand_leaf_node -> tree_builder.expand_descendants_to_depth_wrt_NN (:333-337, tree_builder.pyx:122-215)
with a deterministic synthetic policy net (oracle.synthetic_policy: a seeded function of the position,
probabilities quantised to 2^-20 so that 1 + p is exact in float32).  The whole tree is dumped after
every checkpoint: per node parent, move index, colour, visits, wins, gameover_visits, gameover_wins,
win_status, gameover, subtree_checked, subtree_being_checked, threads_checking_node and the children
in the reference's order.

Run in the build container only:  python tests/golden/make_tree_golden.py
Output: tests/golden/ref_tree_golden.npz (committed)."""
import io
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_ref  # noqa: E402
from oracle import oracle as O  # noqa: E402

FILES = "abcdefgh"
V = {"e": 0, "w": 1, "b": 2}
WS = {None: -1, False: 0, True: 1}


def to_squares(board):
    return np.array([V[board[r][c]] for r in range(1, 9) for c in FILES], np.int8)


class Net(object):
    def __init__(self):
        self.calls = 0

    def evaluate(self, nodes, **kw):
        self.calls += len(nodes)
        return np.stack([O.synthetic_policy(to_squares(n["game_board"]), n["color"] == "Black") for n in nodes])


def dump(root):
    """BFS over the reference's dict tree -> arrays (children of a node are contiguous, in its order)."""
    rows, queue, parent_of = [], [(root, -1)], []
    i = 0
    while i < len(queue):
        node, par = queue[i]
        kids = node["children"]
        rows.append((par, node["index"], node["color"] == "Black", node["visits"], int(node["wins"]),
                     node["gameover_visits"], int(node["gameover_wins"]), WS[node["win_status"]],
                     bool(node["gameover"]), bool(node["subtree_checked"]), bool(node["subtree_being_checked"]),
                     node["threads_checking_node"], -1 if kids is None else len(kids), node["height"]))
        if kids is not None:
            for k in kids:
                queue.append((k, i))
        i += 1
    return np.array(rows, np.int64)


def main():
    ref = load_ref()
    assert ref is not None
    U, B = ref.utils, ref.board_utils
    mc = ref.module("monte_carlo_tree_search.expansion_MCTS_functions")
    tsu = ref.module("monte_carlo_tree_search.tree_search_utils")
    TreeNode = ref.module("monte_carlo_tree_search.TreeNode").TreeNode
    rng = np.random.RandomState(77)
    # start positions: the opening and two positions from a seeded random play-out (plies 14 and 31)
    starts = [(U.initial_game_board(), "White", 0)]
    board, color = U.initial_game_board(), "White"
    for ply in range(59):
        if B.game_over(board)[0]:
            break
        legal = B.enumerate_legal_moves(board, color)
        if ply in (14, 31, 44, 52, 58):
            starts.append((B.copy_game_board(board), color, ply))
        mv = legal[rng.randint(len(legal))]
        f, t = mv.split("-")
        nb = B.copy_game_board(board)
        nb[int(t[1])][t[0]] = board[int(f[1])][f[0]]
        nb[int(f[1])][f[0]] = "e"
        board, color = nb, ("Black" if color == "White" else "White")
    out = {}
    checkpoints = (1, 2, 5, 12, 25, 40)
    for s, (board, color, ply) in enumerate(starts):
        wp = [c + str(r) for r in range(1, 9) for c in FILES if board[r][c] == "w"]
        bp = [c + str(r) for r in range(1, 9) for c in FILES if board[r][c] == "b"]
        root = TreeNode(B.copy_game_board(board), wp, bp, color, 0, None, ply)
        sim_info = tsu.SimulationInfo(io.StringIO())
        sim_info.update(time_to_think=10 ** 6, start_time=time.time(), root=root, main_pid="not-this-thread")
        net, lock = Net(), threading.Lock()
        out["start%d_squares" % s] = to_squares(board)
        out["start%d_black" % s] = np.uint8(color == "Black")
        done_sims = 0
        for cp in checkpoints:
            while done_sims < cp:
                mc.run_MCTS_with_expansions_simulation(root, 5, sim_info, net, sim_info["start_time"], 10 ** 6, lock)
                done_sims += 1
            out["start%d_after%d" % (s, cp)] = dump(root)
            out["start%d_evals%d" % (s, cp)] = np.int64(net.calls)
            # the move the search would play now (expansion_MCTS_functions.pyx:129: randomly_choose_a_winning_move,
            # deterministic despite its name) as the move index of the chosen child
            out["start%d_move%d" % (s, cp)] = np.int64(tsu.randomly_choose_a_winning_move(root)["index"])
        # root reuse (assign_root, expansion_MCTS_functions.pyx:209-266, for the position after the move the
        # search chose): the chosen child becomes the root of the next search, reset_thread_flag (:197-207)
        # clears the thread bookkeeping of its subtree, 15 more simulations follow
        best = tsu.randomly_choose_a_winning_move(root)
        if not best["gameover"]:
            sim2 = tsu.SimulationInfo(io.StringIO())
            sim2.update(time_to_think=10 ** 6, start_time=time.time(), main_pid="not-this-thread")
            new_root = mc.assign_root(best["game_board"], best["color"], root, None, ply + 2, net, sim2, lock)
            assert new_root is best
            sim2["root"] = new_root
            mc.reset_thread_flag(new_root)
            for _ in range(15):
                mc.run_MCTS_with_expansions_simulation(new_root, 5, sim2, net, sim2["start_time"], 10 ** 6, lock)
            out["start%d_reuse_move" % s] = np.int64(best["index"])
            out["start%d_reuse_after15" % s] = dump(new_root)
            out["start%d_reuse_evals" % s] = np.int64(net.calls)
            out["start%d_reuse_next_move" % s] = np.int64(tsu.randomly_choose_a_winning_move(new_root)["index"])
        # init_new_root (tree_builder.pyx:809-897): the position after a move that is NOT in the tree yet --
        # the previous node was never expanded.  A fresh tree is searched 12 simulations, an UNEXPANDED child
        # u of its root stands in for "our move", the opponent answers with u's lowest-index legal move:
        # assign_root then expands u (with sim_info['root'] still None), finds the answer among u's new
        # children and makes it the root; 10 simulations follow.
        root3 = TreeNode(B.copy_game_board(board), list(wp), list(bp), color, 0, None, ply)
        sim3 = tsu.SimulationInfo(io.StringIO())
        sim3.update(time_to_think=10 ** 6, start_time=time.time(), root=root3, main_pid="not-this-thread")
        net3 = Net()
        for _ in range(12):
            mc.run_MCTS_with_expansions_simulation(root3, 5, sim3, net3, sim3["start_time"], 10 ** 6, lock)
        cands = [c for c in root3["children"] if c["children"] is None and not c["gameover"] and c["win_status"] is None]
        if cands:
            u = cands[-1]
            legal = B.enumerate_legal_moves(u["game_board"], u["color"])
            reply = min(legal, key=U.index_lookup_by_move)
            nb = B.move_piece_update_piece_arrays(u["game_board"], reply, u["color"])[0]   # incl. the side-to-move keys
            reply_color = "Black" if u["color"] == "White" else "White"
            sim4 = tsu.SimulationInfo(io.StringIO())
            sim4.update(time_to_think=10 ** 6, start_time=time.time(), main_pid="not-this-thread")
            new_root = mc.assign_root(nb, reply_color, u, reply, ply + 2, net3, sim4, lock)
            assert new_root["parent"] is u and new_root["index"] == U.index_lookup_by_move(reply)
            sim4["root"] = new_root
            mc.reset_thread_flag(new_root)
            for _ in range(10):
                mc.run_MCTS_with_expansions_simulation(new_root, 5, sim4, net3, sim4["start_time"], 10 ** 6, lock)
            out["start%d_newroot_u" % s] = np.int64(u["index"])
            out["start%d_newroot_reply" % s] = np.int64(new_root["index"])
            out["start%d_newroot_after10" % s] = dump(new_root)
            out["start%d_newroot_evals" % s] = np.int64(net3.calls)
        print("start", s, color, "ply", ply, "sims", done_sims, "evaluations", net.calls, "nodes",
              len(out["start%d_after%d" % (s, checkpoints[-1])]), "root win_status", root["win_status"])
    out["checkpoints"] = np.array(checkpoints)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tree_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
