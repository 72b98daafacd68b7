#!/usr/bin/env python
"""Generate golden vectors FROM THE COMPILED REFERENCE (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run in the build container only; the
output tests/golden/ref_golden.npz is committed so that the GPU box (which has no
/root/reference) can check against the reference's own answers.

Contents (all produced by reference code, not by the oracle):
  squares[N,64] int8, black[N] uint8      positions (opening, debug board both colours,
                                          README game replay, seeded random play-outs)
  planes[N,8,8,4] int8                    tools.utils.convert_board_to_2d_matrix_POEB
  legal[N,155] uint8                      board_utils.enumerate_legal_moves ->
                                          tools.utils.index_lookup_by_move
  legal_pa[N,155] uint8                   same via enumerate_legal_moves_using_piece_arrays
  child_squares[64,48,64] int8,           for the first 64 cases, every legal move in ascending
  child_flags[64,48] uint8                index order: board_utils.move_piece_update_piece_arrays
                                          (bit1 = remove_opponent_piece) and the win rule
                                          index > 131 of tree_builder.check_for_winning_move (bit0)
  white_moves[155], black_moves[155]      tools.utils.move_lookup_by_index
  gen_white[155], gen_black[155]          tools.utils.generate_move_lookup
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_ref  # noqa: E402

FILES = "abcdefgh"
V = {"e": 0, "w": 1, "b": 2}


def to_squares(board):
    return np.array([V[board[r][c]] for r in range(1, 9) for c in FILES], np.int8)


def apply_move(B, board, move, color):
    nb = B.copy_game_board(board)
    f, t = move.split("-")
    nb[int(t[1])][t[0]] = board[int(f[1])][f[0]]
    nb[int(f[1])][f[0]] = "e"
    nb[9] = 0 if color == "White" else 1
    return nb


def main():
    ref = load_ref()
    assert ref is not None, "build oracle/_ref first: python oracle/build_ref.py"
    U, B = ref.utils, ref.board_utils
    rng = np.random.RandomState(20260601)
    cases = []  # (board, color)
    cases.append((U.initial_game_board(), "White"))
    cases.append((U.initial_game_board(), "Black"))
    cases.append((B.debug_game_board(), "White"))
    cases.append((B.debug_game_board(), "Black"))
    # README.md:5 game record (trmph): replay every position
    rec = "g2h3b7b6h1g2c7c6a2b3b6b5f2f3a7a6b3a4b5a4b2b3a4b3c2b3c6c5g2f2h7g6f3f4d7c6f2f3e7e6d2d3a8b7b1c2b8a7h3g4a6a5h2h3g6f5g4f5e6f5b3a4a5a4c2b3a4b3a1a2b3a2c1c2a2a1"
    # trmph moves have no separator: from,to squares in pairs; keep only well-formed ones
    board, color = U.initial_game_board(), "White"
    i = 0
    while i + 4 <= len(rec):
        mv = rec[i:i + 2] + "-" + rec[i + 2:i + 4]
        i += 4
        legal = B.enumerate_legal_moves(board, color)
        if mv not in legal or B.game_over(board)[0]:
            break
        cases.append((board, color))
        board = apply_move(B, board, mv, color)
        color = "Black" if color == "White" else "White"
    n_readme = len(cases) - 4
    # seeded random play-outs
    while len(cases) < 4 + n_readme + 400:
        board, color = U.initial_game_board(), "White"
        traj = []
        while not B.game_over(board)[0]:
            legal = B.enumerate_legal_moves(board, color)
            if not legal:
                break
            traj.append((board, color))
            board = apply_move(B, board, legal[rng.randint(len(legal))], color)
            color = "Black" if color == "White" else "White"
        for k in rng.choice(len(traj), size=min(8, len(traj)), replace=False):
            cases.append(traj[k])
    n = len(cases)
    squares = np.zeros((n, 64), np.int8)
    black = np.zeros(n, np.uint8)
    planes = np.zeros((n, 8, 8, 4), np.int8)
    legal = np.zeros((n, 155), np.uint8)
    legal_pa = np.zeros((n, 155), np.uint8)
    for k, (board, color) in enumerate(cases):
        squares[k] = to_squares(board)
        black[k] = color == "Black"
        planes[k] = U.convert_board_to_2d_matrix_POEB(board, color)
        for m in B.enumerate_legal_moves(board, color):
            legal[k, U.index_lookup_by_move(m)] = 1
        me = "w" if color == "White" else "b"
        pieces = [c + str(r) for r in range(1, 9) for c in FILES if board[r][c] == me]
        for m in B.enumerate_legal_moves_using_piece_arrays(color, board, pieces):
            legal_pa[k, U.index_lookup_by_move(m)] = 1
    child_squares = np.zeros((64, 48, 64), np.int8)
    child_flags = np.zeros((64, 48), np.uint8)
    for k, (board, color) in enumerate(cases[:64]):
        for slot, j in enumerate(np.nonzero(legal[k])[0]):
            mv = U.move_lookup_by_index(int(j), color)
            nb, add, rem, cap = B.move_piece_update_piece_arrays(board, mv, color)
            child_squares[k, slot] = to_squares(nb)
            child_flags[k, slot] = (1 if (j > 131 and j != 154) else 0) | (2 if cap else 0)
            assert nb[9] == (0 if color == "White" else 1)      # the other colour moves next
    wm = np.array([U.move_lookup_by_index(i, "White") for i in range(155)])
    bm = np.array([U.move_lookup_by_index(i, "Black") for i in range(155)])
    gw, gb = U.generate_move_lookup()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(out, squares=squares, black=black, planes=planes, legal=legal,
                        legal_pa=legal_pa, child_squares=child_squares, child_flags=child_flags, white_moves=wm, black_moves=bm,
                        gen_white=np.array(gw), gen_black=np.array(gb),
                        n_readme=np.int64(n_readme))
    print("wrote", out, "cases", n, "readme plies", n_readme, "black share", black.mean(),
          "max legal", legal.sum(1).max())


if __name__ == "__main__":
    main()
