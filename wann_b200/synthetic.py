"""Synthetic leaf positions for benchmarks and demos: seeded random legal play-outs from the
Breakthrough opening (host-side numpy; not on the hot path).  Rules as in the reference's
move generator (Breakthrough_Player/board_utils.pyx:331-400): forward needs an empty target,
diagonals need a target that is not an own piece; a side reaching the far rank ends the game."""
import numpy as np


def initial_squares():
    sq = np.zeros(64, np.int8)
    sq[:16] = 1
    sq[48:] = 2
    return sq


def _pov(sq, black):
    """[n,64] absolute squares -> side-to-move frame (0 empty, 1 own, 2 opponent), [n,8,8]."""
    black = black.astype(bool)[:, None]
    rot = np.where(black, sq[:, ::-1], sq)
    own = np.where(black, 2, 1)
    return np.where(rot == 0, 0, np.where(rot == own, 1, 2)).astype(np.int8).reshape(-1, 8, 8)


def _legal(pov):
    """[n,8,8] -> bool [n,7,8,3]: piece on (r,c) may move to (r+1, c+d), d = -1,0,+1."""
    own = pov[:, :7] == 1
    nxt = pov[:, 1:]
    m = np.zeros(own.shape + (3,), bool)
    m[..., 1] = own & (nxt == 0)
    m[:, :, 1:, 0] = own[:, :, 1:] & (nxt[:, :, :-1] != 1)
    m[:, :, :-1, 2] = own[:, :, :-1] & (nxt[:, :, 1:] != 1)
    return m


def random_positions(n, seed=20260601, max_plies=60):
    """n non-terminal positions: n games played in lock-step for a random number of plies each.
    Returns squares int8 [n,64] (0 'e', 1 'w', 2 'b'; index (rank-1)*8+file), black_to_move uint8 [n]."""
    rng = np.random.RandomState(seed)
    sq = np.tile(initial_squares(), (n, 1))
    black = np.zeros(n, np.uint8)
    target = rng.randint(0, max_plies, size=n)
    ply = np.zeros(n, np.int64)
    active = ply < target
    for _ in range(max_plies):
        ids = np.nonzero(active)[0]
        if ids.size == 0:
            break
        pov = _pov(sq[ids], black[ids])
        m = _legal(pov).reshape(ids.size, -1)
        pick = (rng.random_sample(m.shape) * m).argmax(1)
        has = m.any(1)
        r, rem = np.divmod(pick, 24)
        c, d = np.divmod(rem, 3)
        frm, to = r * 8 + c, (r + 1) * 8 + c + d - 1
        wins = (r + 1) == 7                                   # reaching the far rank ends the game
        go = has & ~wins
        blk = black[ids].astype(bool)
        frm = np.where(blk, 63 - frm, frm)
        to = np.where(blk, 63 - to, to)
        g = ids[go]
        sq[g, to[go]] = sq[g, frm[go]]
        sq[g, frm[go]] = 0
        black[g] ^= 1
        ply[g] += 1
        active[ids[~go]] = False                              # keep the pre-terminal position
        active[g] = ply[g] < target[g]
    return sq.astype(np.int8), black
