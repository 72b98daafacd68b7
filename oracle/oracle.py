"""CPU oracle for the WaNN batched policy-net leaf-evaluation path (numpy).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  The product
(wann_b200/) never does: it fails loudly when its CUDA library is missing.

Every function restates a piece of the reference (paths relative to
/root/reference) and cites the lines it follows.

PARITY STATUS
  * encode / move index / legal moves: PINNED.  Checked against (a) the literal
    155-entry move tables tools/utils.pyx:71-113,126-141,300-315, (b) the compiled
    reference itself (oracle/_ref, built by oracle/build_ref.py) on thousands of
    play-out positions, (c) committed fixtures tests/golden/*.npz generated from the
    compiled reference by tests/golden/make_golden.py.
  * logits / probabilities: **PARITY UNPINNED**.  The arithmetic lives in TensorFlow
    1.0.0 (model.meta meta_info_def.tensorflow_version), which is not vendored in
    the reference tree, not installed and not installable here; the weight shard
    policy_net_model/model.data-00000-of-00001 is absent (.MISSING_LARGE_BLOBS:1);
    no reference test pins any logit.  forward_fp32() restates the published
    semantics of tf.nn.conv2d(NHWC, SAME, stride 1) / bias_add / relu / reshape /
    matmul / softmax at the reference's call sites
    Breakthrough_Player/policy_net_utils.pyx:106-124,149-178,215-239.
"""
import numpy as np

N_MOVES = 155          # tools/utils.pyx:605-641 (154 from-to moves + 'no-move')
NO_MOVE = 154          # tools/utils.pyx:113
MAX_LEGAL = 48         # 16 pieces x 3 directions
EMPTY, WHITE, BLACK = 0, 1, 2
FILES = "abcdefgh"

# ----------------------------------------------------------------------------
# Board type  (tools/utils.pyx:518-541, Breakthrough_Player/board_utils.pyx:18-30)
# compact form: squares[64] int8, index = (rank-1)*8 + file, 0='e' 1='w' 2='b'
# ----------------------------------------------------------------------------
_CH2V = {"e": EMPTY, "w": WHITE, "b": BLACK}
_V2CH = "ewb"


def squares_from_board(board):
    """Reference board dict {1..8:{'a'..'h':'w'|'b'|'e'}, 9:.., 10:..} -> int8[64]."""
    sq = np.zeros(64, dtype=np.int8)
    for rank in range(1, 9):
        row = board[rank]
        for f, ch in enumerate(FILES):
            sq[(rank - 1) * 8 + f] = _CH2V[row[ch]]
    return sq


def board_from_squares(sq, white_to_move=1, last_mover=-1):
    """Inverse of squares_from_board (keys 9/10 as in tools/utils.pyx:533-536)."""
    board = {9: int(white_to_move), 10: int(last_mover)}
    for rank in range(1, 9):
        board[rank] = {ch: _V2CH[int(sq[(rank - 1) * 8 + f])] for f, ch in enumerate(FILES)}
    return board


def initial_squares():
    """tools/utils.pyx:518-541 initial_game_board(): white ranks 1-2, black 7-8."""
    sq = np.zeros(64, dtype=np.int8)
    sq[0:16] = WHITE
    sq[48:64] = BLACK
    return sq


def pov_squares(squares, black_to_move):
    """Side-to-move point of view: values 0=empty 1=player 2=opponent.

    tools/utils.pyx:563-600: for Black the board is rotated 180 degrees (column
    mirror :566-584 + row swap :586-599), i.e. pov index = 63 - index; player/opponent
    letters swap (:842-848).
    squares [N,64] int8, black_to_move [N] -> [N,64] int8.
    """
    squares = np.asarray(squares, dtype=np.int8).reshape(-1, 64)
    black = np.asarray(black_to_move).astype(bool).reshape(-1)
    pov = np.where(black[:, None], squares[:, ::-1], squares)
    # recolour: for white-to-move 1=w stays player, 2=b opponent; for black swap
    swapped = np.where(pov == WHITE, 2, np.where(pov == BLACK, 1, 0)).astype(np.int8)
    return np.where(black[:, None], swapped, pov).astype(np.int8)


def encode_planes(squares, black_to_move):
    """convert_board_to_2d_matrix_POEB (tools/utils.pyx:546-602) +
    generate_binary_plane_POEB (:819-866).  -> int8 [N,8,8,4] planes
    (Player, Opponent, Empty, Bias=1); row index 0 = side-to-move's home rank."""
    pov = pov_squares(squares, black_to_move)
    n = pov.shape[0]
    planes = np.zeros((n, 64, 4), dtype=np.int8)
    planes[:, :, 0] = pov == 1
    planes[:, :, 1] = pov == 2
    planes[:, :, 2] = pov == 0
    planes[:, :, 3] = 1                       # utils.pyx:828 [[0,0,0,bias]]
    return planes.reshape(n, 8, 8, 4)


# ----------------------------------------------------------------------------
# Move-index codec (tools/utils.pyx:69-114, 118-477, 480-516, 605-641)
# ----------------------------------------------------------------------------
def move_index(move, color=None):
    """generate_transition_vector (tools/utils.pyx:605-641) closed form.
    move like 'a2-a3'; colour inferred from direction when not given."""
    move = move.lower()
    if move == "no-move":
        return NO_MOVE
    fc, fr, tc, tr = ord(move[0]) - 97, int(move[1]), ord(move[3]) - 97, int(move[4])
    if color is None:
        color = "White" if tr > fr else "Black"
    col_off = fc * 3
    if color == "Black":
        row_off = (tr - 1) * 22
        assert row_off == (fr - 2) * 22       # utils.pyx:632
        return 153 - ((tc - fc) + col_off + row_off)
    row_off = (fr - 1) * 22
    assert row_off == (tr - 2) * 22           # utils.pyx:637
    return (tc - fc) + col_off + row_off


def move_tables():
    """generate_move_lookup (tools/utils.pyx:480-516): index -> move string for
    White and Black (each 155 long, last entry 'no-move')."""
    chars = FILES
    white = []
    for k in range(1, 8):
        for i in range(8):
            if i > 0:
                white.append("%s%d-%s%d" % (chars[i], k, chars[i - 1], k + 1))
            white.append("%s%d-%s%d" % (chars[i], k, chars[i], k + 1))
            if i < 7:
                white.append("%s%d-%s%d" % (chars[i], k, chars[i + 1], k + 1))
    black = []
    for k in range(8, 1, -1):
        for i in range(7, -1, -1):
            if i < 7:
                black.append("%s%d-%s%d" % (chars[i], k, chars[i + 1], k - 1))
            black.append("%s%d-%s%d" % (chars[i], k, chars[i], k - 1))
            if i > 0:
                black.append("%s%d-%s%d" % (chars[i], k, chars[i - 1], k - 1))
    white.append("no-move")
    black.append("no-move")
    return white, black


def pov_move_index(r, c, d):
    """POV-frame closed form: piece on pov (r,c) moving to (r+1,c+d)."""
    return 22 * r + 3 * c + d


# ----------------------------------------------------------------------------
# Legal moves (Breakthrough_Player/board_utils.pyx:293-400)
# ----------------------------------------------------------------------------
def legal_mask(squares, black_to_move):
    """-> bool [N,155]: policy-net indices of the side-to-move's legal moves.

    check_forward_move (:355-378): forward legal iff target == 'e'.
    check_left/right_diagonal_move (:331-353, :380-400): diagonal legal iff
    in-board and target != own piece.  Index by generate_transition_vector.
    Pieces on the far rank generate nothing (board_utils.pyx:364,372: the game is
    already over there)."""
    pov = pov_squares(squares, black_to_move).reshape(-1, 8, 8)
    n = pov.shape[0]
    mask = np.zeros((n, N_MOVES), dtype=bool)
    for r in range(7):
        for c in range(8):
            own = pov[:, r, c] == 1
            for d in (-1, 0, 1):
                cc = c + d
                if cc < 0 or cc > 7:
                    continue
                tgt = pov[:, r + 1, cc]
                ok = (tgt == 0) if d == 0 else (tgt != 1)
                mask[:, 22 * r + 3 * c + d] = own & ok
    return mask


def game_over(squares):
    """board_utils.pyx:408-422: white on rank 8 or black on rank 1."""
    squares = np.asarray(squares).reshape(-1, 64)
    return (squares[:, 56:64] == WHITE).any(1) | (squares[:, 0:8] == BLACK).any(1)


# ----------------------------------------------------------------------------
# Policy net forward (Breakthrough_Player/policy_net_utils.pyx:106-124,149-178,215-239
#                     == policy_net_model/PolicyNet.py:38-114)
# ----------------------------------------------------------------------------
WEIGHT_NAMES = ["hidden_layer/%d/%s" % (i, k) for i in range(1, 6) for k in ("weights", "bias")] + \
               ["output_layer/weights", "output_layer/bias"]


def weight_shapes(final_kernel=3):
    """model.index shapes: HWIO conv kernels, [64,155] FC."""
    s = {}
    cin = 4
    for i in range(1, 6):
        cout = 192 if i < 5 else 1
        k = 3 if i < 5 else final_kernel
        s["hidden_layer/%d/weights" % i] = (k, k, cin, cout)
        s["hidden_layer/%d/bias" % i] = (cout,)
        cin = cout
    s["output_layer/weights"] = (64, 155)
    s["output_layer/bias"] = (155,)
    return s


def synthetic_weights(preset="trained_like", seed=1, final_kernel=3):
    """Deterministic stand-ins for the missing weight shard.

    'ref_init'    : the reference's initialisers -- conv N(0, sqrt(2/prod(prev shape[1:])))
                    (policy_net_utils.pyx:152,161), conv bias 0.01 (:166-167), FC Xavier
                    uniform (:227), FC bias 0 (:233).
    'trained_like': fan-in-correct He conv N(0, sqrt(2/(k*k*Cin))), small random conv
                    biases, FC scaled so that logits have sigma ~ 2.5 (a peaked policy, so
                    that top-1 agreement is a meaningful test)."""
    rng = np.random.RandomState(seed)
    w = {}
    shapes = weight_shapes(final_kernel)
    for i in range(1, 6):
        shp = shapes["hidden_layer/%d/weights" % i]
        k, _, cin, cout = shp
        if preset == "ref_init":
            std = np.sqrt(2.0 / (64 * cin))
            b = np.full((cout,), 0.01, np.float32)
        else:
            std = np.sqrt(2.0 / (k * k * cin))
            b = (0.05 * rng.standard_normal(cout)).astype(np.float32)
        w["hidden_layer/%d/weights" % i] = (std * rng.standard_normal(shp)).astype(np.float32)
        w["hidden_layer/%d/bias" % i] = b
    if preset == "ref_init":
        lim = np.sqrt(6.0 / (64 + 155))
        w["output_layer/weights"] = rng.uniform(-lim, lim, (64, 155)).astype(np.float32)
        w["output_layer/bias"] = np.zeros(155, np.float32)
    else:
        w["output_layer/weights"] = (0.9 * rng.standard_normal((64, 155))).astype(np.float32)
        w["output_layer/bias"] = (0.1 * rng.standard_normal(155)).astype(np.float32)
    return w


def conv2d_same_nhwc(x, w_hwio, dtype=np.float32):
    """tf.nn.conv2d(input NHWC, filter HWIO, strides 1, padding 'SAME')
    (policy_net_utils.pyx:171-175): cross-correlation, symmetric zero pad (k-1)/2."""
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w_hwio.shape
    ph, pw = (kh - 1) // 2, (kw - 1) // 2
    xp = np.zeros((n, h + 2 * ph, wd + 2 * pw, cin), dtype=dtype)
    xp[:, ph:ph + h, pw:pw + wd, :] = x
    cols = np.empty((n, h, wd, kh * kw * cin), dtype=dtype)
    for i in range(kh):
        for j in range(kw):
            cols[..., (i * kw + j) * cin:(i * kw + j + 1) * cin] = xp[:, i:i + h, j:j + wd, :]
    out = cols.reshape(n * h * wd, kh * kw * cin) @ w_hwio.reshape(kh * kw * cin, cout).astype(dtype)
    return out.reshape(n, h, wd, cout)


def forward_fp32(planes, weights, dtype=np.float32, return_all=False):
    """build_policy_net (policy_net_utils.pyx:106-124): 5 x [conv3x3 SAME + bias_add + relu]
    (hidden_layer_init :149-178; ReLU also on the 1-channel 5th map, default activation
    :149), reshape [-1,64] (:216), matmul [64,155] + bias (:235-238), softmax (:217).
    planes [N,8,8,4] any numeric dtype.  Returns (logits, probs) [N,155]."""
    x = np.asarray(planes).astype(dtype)
    acts = []
    for i in range(1, 6):
        x = conv2d_same_nhwc(x, weights["hidden_layer/%d/weights" % i], dtype)
        x = np.maximum(x + weights["hidden_layer/%d/bias" % i].astype(dtype), 0)
        acts.append(x)
    h = x.reshape(-1, 64)
    logits = h @ weights["output_layer/weights"].astype(dtype) + weights["output_layer/bias"].astype(dtype)
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    probs = e / e.sum(axis=1, keepdims=True)
    if return_all:
        return logits, probs, acts
    return logits, probs


# ----------------------------------------------------------------------------
# Prior gather + ranking (tree_search_utils.pyx:979-986, tree_builder.pyx:557,606-609;
# board_utils.pyx:262-291)
# ----------------------------------------------------------------------------
def ranked_children(probs, mask):
    """For each position: legal indices sorted by prior descending (prior =
    softmax155[idx], NO renormalisation: tree_search_utils.pyx:983-986; sort
    tree_builder.pyx:557).  Ties broken by ascending move index (the reference's tie
    order depends on history-dependent piece-list order and is not a function of the
    board; DESIGN.md 'tie-break').  Returns idx uint8 [N,48] (pad 255),
    prior float32 [N,48] (pad 0), n_legal uint8 [N]."""
    probs = np.asarray(probs, dtype=np.float32)
    n = probs.shape[0]
    idx = np.full((n, MAX_LEGAL), 255, np.uint8)
    pri = np.zeros((n, MAX_LEGAL), np.float32)
    cnt = np.zeros(n, np.uint8)
    for i in range(n):
        legal = np.nonzero(mask[i])[0]
        order = np.lexsort((legal, -probs[i, legal]))     # primary -prob, secondary idx
        legal = legal[order]
        k = len(legal)
        idx[i, :k] = legal
        pri[i, :k] = probs[i, legal]
        cnt[i] = k
    return idx, pri, cnt


def top1_legal(probs, mask):
    """get_best_move (board_utils.pyx:262-277): first legal index in the descending
    ranking == argmax of prior over legal indices."""
    p = np.where(mask, probs, -1.0)
    return p.argmax(axis=1)


# ----------------------------------------------------------------------------
# Position generator: seeded random legal play-outs
# ----------------------------------------------------------------------------
def apply_pov_move(squares, black, idx):
    """Apply policy index idx for the side to move; returns new squares (absolute)."""
    r, rem = divmod(int(idx), 22)
    # invert 22r + 3c + d
    c = (rem + 1) // 3
    d = rem - 3 * c
    frm, to = r * 8 + c, (r + 1) * 8 + c + d
    if black:
        frm, to = 63 - frm, 63 - to
    sq = squares.copy()
    sq[to] = sq[frm]
    sq[frm] = EMPTY
    return sq


def random_positions(n, seed=20260601, balance=True):
    """n non-terminal positions from seeded random play-outs from the opening; one
    position sampled uniformly per play-out ply range.  -> squares int8 [n,64],
    black_to_move uint8 [n]."""
    rng = np.random.RandomState(seed)
    out_sq, out_b = [], []
    want_black = 0
    while len(out_sq) < n:
        sq = initial_squares()
        black = 0
        traj = []
        for _ in range(200):
            if game_over(sq)[0]:
                break
            m = np.nonzero(legal_mask(sq[None], [black])[0])[0]
            if len(m) == 0:
                break
            traj.append((sq, black))
            sq = apply_pov_move(sq, black, m[rng.randint(len(m))])
            black ^= 1
        if balance:
            cand = [t for t in traj if t[1] == want_black]
        else:
            cand = traj
        if not cand:
            continue
        s, b = cand[rng.randint(len(cand))]
        out_sq.append(s)
        out_b.append(b)
        want_black ^= 1
    return np.stack(out_sq).astype(np.int8), np.array(out_b, dtype=np.uint8)


def random_positions_fast(n, seed=20260601):
    """Vectorised variant for large suites: plays n games in lock-step for a random
    number of plies each (uniform 0..59), keeping a game frozen once its target ply is
    reached or the next move would end it.  Same distributional intent as
    random_positions, much faster for n >= 1e4."""
    rng = np.random.RandomState(seed)
    sq = np.tile(initial_squares(), (n, 1))
    black = np.zeros(n, dtype=np.uint8)
    target = rng.randint(0, 60, size=n)
    ply = np.zeros(n, dtype=np.int64)
    active = ply < target
    for _ in range(60):
        if not active.any():
            break
        ids = np.nonzero(active)[0]
        m = legal_mask(sq[ids], black[ids])
        # choose a random legal move per active board
        r = rng.random_sample(m.shape) * m
        choice = r.argmax(axis=1)
        has = m.any(axis=1)
        for j, i in enumerate(ids):
            if not has[j]:
                active[i] = False
                continue
            nsq = apply_pov_move(sq[i], black[i], choice[j])
            if game_over(nsq)[0]:
                active[i] = False          # keep the pre-terminal position
                continue
            sq[i] = nsq
            black[i] ^= 1
            ply[i] += 1
            if ply[i] >= target[i]:
                active[i] = False
    return sq.astype(np.int8), black


# ----------------------------------------------------------------------------
# bf16 emulation of the tensor-core path (for tight per-layer checks of the CUDA kernel)
# ----------------------------------------------------------------------------
def bf16_round(a):
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (what cvt.rn.bf16.f32 does)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(np.shape(a))


def forward_bf16_emulated(planes, weights, return_all=False):
    """Same graph as forward_fp32 but with the roundings the tcgen05 path applies: conv
    weights and layer inputs rounded to bf16, fp32 accumulation, bias+ReLU in fp32, the
    activations of layers 1-4 rounded to bf16 when stored; layer 5 and the FC/softmax in fp32."""
    x = bf16_round(np.asarray(planes).astype(np.float32))
    acts = []
    for i in range(1, 6):
        w = bf16_round(weights["hidden_layer/%d/weights" % i])
        x = conv2d_same_nhwc(x, w, np.float32)
        x = np.maximum(x + weights["hidden_layer/%d/bias" % i], 0).astype(np.float32)
        if i < 5:
            x = bf16_round(x)
        acts.append(x)
    h = x.reshape(-1, 64)
    logits = h @ weights["output_layer/weights"] + weights["output_layer/bias"]
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    probs = e / e.sum(axis=1, keepdims=True)
    if return_all:
        return logits, probs, acts
    return logits, probs


def f16_round(a):
    """Round-to-nearest-even float32 -> float16 -> float32, saturating at +-65504."""
    return np.clip(np.asarray(a, np.float32), -65504.0, 65504.0).astype(np.float16).astype(np.float32)


def forward_16bit_emulated(planes, weights, fmt="bf16", return_all=False):
    """forward_bf16_emulated generalised to the operand format of the tensor-core mode
    ('bf16' or 'fp16')."""
    rnd = bf16_round if fmt == "bf16" else f16_round
    x = rnd(np.asarray(planes).astype(np.float32))
    acts = []
    for i in range(1, 6):
        x = conv2d_same_nhwc(x, rnd(weights["hidden_layer/%d/weights" % i]), np.float32)
        x = np.maximum(x + weights["hidden_layer/%d/bias" % i], 0).astype(np.float32)
        if i < 5:
            x = rnd(x)
        acts.append(x)
    h = x.reshape(-1, 64)
    logits = h @ weights["output_layer/weights"] + weights["output_layer/bias"]
    z = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(z)
    probs = e / e.sum(axis=1, keepdims=True)
    if return_all:
        return logits, probs, acts
    return logits, probs


def gptq_round(H, W, damp=0.01, exact=None):
    """numpy restatement of wann_b200/csrc/gptq.hpp (the GPTQ / optimal-brain-quantisation sequential rounding):
    H [n,n] second moments of a layer's inputs, W [rows,n] its weights.  Column k of every row is rounded to
    fp16 (kept as is where exact[r,k]); the error is pushed onto the columns j > k along
    U[k,j] / U[k,k], (H + damp * mean(diag H) * I)^-1 = U^T U, U upper triangular.  -> rounded copy of W (float64)."""
    H = np.asarray(H, np.float64)
    n = H.shape[0]
    Hd = 0.5 * (H + H.T) + damp * np.trace(H) / n * np.eye(n)
    U = np.linalg.cholesky(np.linalg.inv(Hd)).T.copy()        # inv = L L^T = U^T U with U = L^T
    W = np.array(W, np.float64)
    for k in range(n):
        q = f16_round(W[:, k]).astype(np.float64)
        if exact is not None:
            q = np.where(exact[:, k] != 0, W[:, k].astype(np.float32).astype(np.float64), q)
        err = (W[:, k] - q) / U[k, k]
        W[:, k] = q
        if k + 1 < n:
            W[:, k + 1:] -= err[:, None] * U[k, k + 1:][None, :]
    return W


def forward_fp16c_emulated(planes, weights, offsets, position_of, split_mask=0, lo_n=192, lo_kgroups=3,
                           return_all=False, w_eff=None):
    """Emulation of the roundings of the CENTRED fp16 tensor-core mode (WANN_MODE_FP16C,
    wann_b200/csrc/conv_stack_tc.cuh) for the same graph as forward_fp32
    (policy_net_utils.pyx:106-124,149-178,215-239).  offsets [4,24] / position_of [4,192] are the
    calibrated centring of the context (wann_debug_centering): channel c of hidden layer l sits at
    position q = position_of[l-1][c] and is stored as a' = fp16(relu(z) - offsets[l-1][(q // 64) * 8 + q % 8]);
    the zero padding of the next conv therefore reads -offset, and that layer's bias absorbs
    sum(w * offset) with the TRUE fp32 weights.  hidden_layer/1 and /5 see their weights as fp16
    hi + lo pairs (22 bits); in the layers of split_mask (bits 1..3 = hidden_layer/2..4) the block
    [input positions < 64 * lo_kgroups] x [output positions < lo_n] too; all other weights are rounded
    to fp16.  w_eff (optional): {layer: [3,3,192,192]} the weights hidden_layer/2..4 are multiplied with, as the
    context reports them (wann_debug_effective_weights) -- the fp16c mode rounds them with error compensation
    (gptq_round), which only the context's calibration data determines.  Accumulation, bias and ReLU in fp32;
    FC/softmax in fp32."""
    def two16(w):                     # hi + lo/4096 pair, as the kernel's conv1 / conv5 operands
        hi = f16_round(w)
        return hi + f16_round((w - hi) * 4096.0) / np.float32(4096.0)
    def hilo(w):                      # hi + unscaled lo stage of a split 192->192 layer
        hi = f16_round(w)
        return hi + f16_round(w - hi)
    offsets = np.asarray(offsets, np.float32)
    position_of = np.asarray(position_of)
    x = np.asarray(planes).astype(np.float32)        # exact in fp16
    acts = []
    c_prev = None
    for i in range(1, 6):
        w = weights["hidden_layer/%d/weights" % i].astype(np.float32)
        b = weights["hidden_layer/%d/bias" % i].astype(np.float64)
        if i == 1 or i == 5:
            w16 = two16(w)
        elif w_eff is not None and i in w_eff:
            w16 = np.asarray(w_eff[i], np.float32)
        else:
            w16 = f16_round(w)
            if (split_mask >> (i - 1)) & 1:
                k_in = position_of[i - 2] < 64 * lo_kgroups          # inputs of layer i = outputs of layer i-1
                k_out = position_of[i - 1] < lo_n
                blk = np.ix_(range(w.shape[0]), range(w.shape[1]), np.nonzero(k_in)[0], np.nonzero(k_out)[0])
                w16[blk] = hilo(w)[blk]
        if c_prev is None:
            z = conv2d_same_nhwc(x, w16, np.float32)
        else:
            n, hh, ww, cin = x.shape
            xp = np.empty((n, hh + 2, ww + 2, cin), np.float32)
            xp[:] = -c_prev                              # the halo holds -offset
            xp[:, 1:-1, 1:-1, :] = x
            kh = w16.shape[0]
            if kh == 1:
                z = x.reshape(-1, cin) @ w16.reshape(cin, -1)
                z = z.reshape(n, hh, ww, -1)
                b = b + (w.reshape(cin, -1).astype(np.float64) * c_prev[:, None].astype(np.float64)).sum(0)
            else:
                cols = np.concatenate([xp[:, dy:dy + hh, dx:dx + ww, :] for dy in range(3) for dx in range(3)], axis=-1)
                z = (cols.reshape(-1, 9 * cin) @ w16.reshape(9 * cin, -1)).reshape(n, hh, ww, -1)
                b = b + (w.astype(np.float64) * c_prev[None, None, :, None].astype(np.float64)).sum(axis=(0, 1, 2))
        if i < 5:
            q = position_of[i - 1].astype(np.int64)
            c = offsets[i - 1][(q // 64) * 8 + q % 8].astype(np.float32)       # per TF channel
            t = z + (b - c.astype(np.float64)).astype(np.float32)
            x = f16_round(np.maximum(t, -c))                                   # a' = relu(z + b') - c
            acts.append(x + c)
            c_prev = c
        else:
            x = np.maximum(z + b.astype(np.float32), 0).astype(np.float32)
            acts.append(x)
    h = x.reshape(-1, 64)
    logits = h @ weights["output_layer/weights"] + weights["output_layer/bias"]
    zz = logits - logits.max(axis=1, keepdims=True)
    e = np.exp(zz)
    probs = e / e.sum(axis=1, keepdims=True)
    if return_all:
        return logits, probs, acts
    return logits, probs


# ----------------------------------------------------------------------------
# Child construction: get_child / get_child_attributes
# (monte_carlo_tree_search/tree_builder.pyx:530-537, 623-658),
# move_piece_update_piece_arrays (Breakthrough_Player/board_utils.pyx:90-131),
# check_for_winning_move (tree_builder.pyx:902-905)
# ----------------------------------------------------------------------------
FLAG_WIN, FLAG_CAPTURE = 1, 2


def expand_children(squares, black_to_move, idx, n_legal):
    """For each position and each listed move index (policy alphabet, side-to-move frame) build
    the child board: the piece moves from->to, the origin becomes empty (board_utils.pyx:119-120);
    a diagonal move onto an opponent piece is a capture (:124-130); a move index > 131 reaches
    the far rank = game over, won by the mover (tree_builder.pyx:902-905).
    idx uint8 [N,48] (255 = pad), n_legal [N] -> child_squares int8 [N,48,64] (absolute frame,
    zeros for pads), flags uint8 [N,48] (bit0 win, bit1 capture).  The child is to be moved by
    the other colour (get_opponent_color, tree_builder.pyx:646)."""
    squares = np.asarray(squares, np.int8).reshape(-1, 64)
    n = squares.shape[0]
    out = np.zeros((n, MAX_LEGAL, 64), np.int8)
    flags = np.zeros((n, MAX_LEGAL), np.uint8)
    for i in range(n):
        blk = int(black_to_move[i])
        for k in range(int(n_legal[i])):
            j = int(idx[i, k])
            r, rem = divmod(j, 22)
            c = (rem + 1) // 3
            d = rem - 3 * c
            frm, to = r * 8 + c, (r + 1) * 8 + c + d
            if blk:
                frm, to = 63 - frm, 63 - to
            child = squares[i].copy()
            f = FLAG_WIN if j > 131 else 0
            if child[to] != 0:
                f |= FLAG_CAPTURE
            child[to] = child[frm]
            child[frm] = 0
            out[i, k] = child
            flags[i, k] = f
    return out, flags


# ----------------------------------------------------------------------------
# Self-play style step (BASELINE config 5; wann_tree_step): pick one child per game from the
# ranked legal priors, apply it, restart finished games.  The selection rule is ours (the
# reference's tree policy is out of scope); the pieces it is made of are the reference's:
# policy move = first ranked legal index (board_utils.pyx:262-277), move application
# (board_utils.pyx:119-120), game over when the move index > 131 (tree_builder.pyx:902-905),
# restart from initial_game_board (tools/utils.pyx:518-541).
# ----------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def tree_step(squares, black_to_move, idx, prior, n_legal, greedy=False, seed=0, step=0):
    """-> (new squares int8 [N,64], new black_to_move uint8 [N], chosen idx uint8 [N] (255 = none),
    finished uint8 [N]).  Sampling: u = (splitmix64(splitmix64(seed ^ step*K) ^ game) >> 40) / 2^24
    * sum(priors) in float32, first slot whose running float32 sum exceeds u."""
    squares = np.asarray(squares, np.int8).reshape(-1, 64)
    n = squares.shape[0]
    out_sq = squares.copy()
    out_bl = np.asarray(black_to_move, np.uint8).copy()
    chosen = np.full(n, 255, np.uint8)
    finished = np.ones(n, np.uint8)
    base = splitmix64((seed ^ ((step * 0xD1342543DE82EF95) & _M64)) & _M64)
    for g in range(n):
        cnt = int(n_legal[g])
        k = 0
        if not greedy and cnt > 1:
            pr = np.asarray(prior[g, :cnt], np.float32)
            total = np.float32(0)
            for p in pr:
                total = np.float32(total + p)
            u = np.float32(np.float32(splitmix64(base ^ g) >> 40) * np.float32(1.0 / 16777216.0)) * total
            cum, k = np.float32(0), cnt - 1
            for i, p in enumerate(pr):
                cum = np.float32(cum + p)
                if u < cum:
                    k = i
                    break
        if cnt > 0:
            chosen[g] = idx[g, k]
            finished[g] = 1 if chosen[g] > 131 else 0
        if finished[g]:
            out_sq[g] = initial_squares()
            out_bl[g] = 0
        else:
            out_sq[g] = apply_pov_move(squares[g], int(out_bl[g]), int(chosen[g]))
            out_bl[g] ^= 1
    return out_sq, out_bl, chosen, finished


# ----------------------------------------------------------------------------
# Child seeding: what the reference does with the priors right after the evaluation
# (SURVEY 8(a) a10 + 8(f)-1).  Restates, per parent position,
#   get_child                      tree_builder.pyx:530-536  (TreeNode zeros, TreeNode.pyx:7-20)
#   check_for_winning_move         tree_builder.pyx:902-905 -> set_game_over_values :907-914
#   update_child                   tree_search_utils.pyx:907-1017 (capture close to home :963-980,
#                                  prior -> visits/wins seeds :982-1014)
#   assign_children                tree_builder.pyx:538-620 ("maybe gameover next move" :563-581:
#                                  game-saving captures, set_game_over_values_1_step_lookahead :916-921;
#                                  otherwise eval_children, tree_search_utils.pyx:1056-1071 with
#                                  evaluation_function :812-858)
#   update_tree_wins / _losses     tree_search_utils.pyx:43-66 (one level: child -> parent)
#   update_win_status_from_children  tree_search_utils.pyx:86-132
# for a parent WITHOUT a parent of its own (the propagation above the parent is tree business).
# Pinned by tests/golden/ref_seed_golden.npz (made by the compiled reference).
# ----------------------------------------------------------------------------
SEED_WIN, SEED_CAPTURE, SEED_CLOSE_TO_HOME, SEED_GAME_SAVING, SEED_DOOMED, SEED_EVAL_ONE = 1, 2, 4, 8, 16, 32

# evaluation_function's weighted_board (tree_search_utils.pyx:814-823), rank 1..8 x file a..h,
# used in ABSOLUTE coordinates for both colours
EVAL_TABLE = np.array([
    [5, 28, 28, 12, 12, 28, 28, 5],
    [2, 3, 3, 3, 3, 3, 3, 2],
    [3, 5, 10, 10, 10, 10, 5, 3],
    [6, 9, 16, 16, 16, 16, 9, 6],
    [9, 15, 21, 21, 21, 21, 15, 9],
    [14, 22, 22, 22, 22, 22, 22, 14],
    [21, 23, 23, 23, 23, 23, 23, 21],
    [24, 24, 24, 24, 24, 24, 24, 24]], np.int32).reshape(64)


def seed_children(squares, black_to_move, idx, prior, n_legal):
    """idx uint8 [N,48] / prior float32 [N,48] / n_legal [N]: the children in ANY order (e.g. the
    ranked order of ranked_children).  Returns
      stats  int32 [N,48,4]  visits, wins, gameover_visits, gameover_wins of every child,
      flags  uint8 [N,48]    SEED_* bits,
      parent int32 [N,6]     visits, wins, gameover_visits, gameover_wins the parent has gained,
                             its win_status (-1 None / 0 False / 1 True), threat (the opponent can
                             reach the far rank next move)."""
    squares = np.asarray(squares, np.int8).reshape(-1, 64)
    n = squares.shape[0]
    stats = np.zeros((n, MAX_LEGAL, 4), np.int32)
    flags = np.zeros((n, MAX_LEGAL), np.uint8)
    parent = np.zeros((n, 6), np.int32)
    threshold = float(np.float32(1.001))                      # `float threshold = 1.001`, tree_search_utils.pyx:95
    for i in range(n):
        blk = int(black_to_move[i])
        pov = pov_squares(squares[i:i + 1], np.array([blk], np.uint8))[0]
        threat = bool((pov[8:16] == 2).any())                 # tree_builder.pyx:556-563 in the mover's frame
        pv = pw = pgv = pgw = 0
        statuses, considering = [], []
        ev_l = ev_w = 0
        for k in range(int(n_legal[i])):
            j = int(idx[i, k])
            p32 = np.float32(prior[i, k])
            r, rem = divmod(j, 22)
            c = (rem + 1) // 3
            d = rem - 3 * c
            to = (r + 1) * 8 + c + d
            capture = pov[to] == 2
            win = j > 131                                     # tree_builder.pyx:904 (154 is never a child)
            v = w = gv = gw = 0
            ws = None
            f = (SEED_WIN if win else 0) | (SEED_CAPTURE if capture else 0)
            uct = 1.0 + float(p32)                            # `1 + child_val` with `double child_val`, :984
            if win:                                           # set_game_over_values + update_tree_losses(child)
                w, gv, v, ws = 0, 65536 + 1, 65536 + 1, False
                pw += 1; pv += 1; pgw += 1; pgv += 1          # -> update_tree_wins(parent, 1, gameover=True)
            if 23 <= j <= 42 and capture:                     # tree_search_utils.pyx:963-980
                gv, gw, v, w = 100, 0, 100, 0
                uct = float(np.float32(1) + p32)              # `1+NN_output[child_index]`: a float32 scalar
                f |= SEED_CLOSE_TO_HOME
            elif not win:                                     # :982-1014
                nv = int(p32 * np.float32(100))               # float32 product (numpy scalar * int)
                v = 100
                ww = v - nv
                w = min(ww, 75)
                if ww < 70:
                    gv, gw = v, ww
            if threat:                                        # tree_builder.pyx:563-576
                if capture:
                    gv, gw, v, w = 1000, 0, 65536, 0
                    f |= SEED_GAME_SAVING
                elif not win:                                 # set_game_over_values_1_step_lookahead + update_tree_wins(child)
                    v = w = gw = gv = 65536 + 1
                    ws = True
                    f |= SEED_DOOMED
                    pv += 1; pgv += 1                         # -> update_tree_losses(parent, 1, gameover=True)
            else:                                             # eval_children, tree_search_utils.pyx:1056-1071
                child = apply_pov_move(squares[i], blk, j)
                result = int(EVAL_TABLE[child == WHITE].sum()) - int(EVAL_TABLE[child == BLACK].sum())
                child_is_white = bool(blk)                    # the child is to be moved by the other colour
                outcome = 1 if (result > 0) == child_is_white else 0
                if outcome == 1:
                    ev_l += 1
                    f |= SEED_EVAL_ONE
                else:
                    ev_w += 1
            stats[i, k] = (v, w, gv, gw)
            flags[i, k] = f
            statuses.append(ws)
            if uct > threshold:
                considering.append(ws)
        pv += ev_l + ev_w                                     # update_tree_losses(parent, L); update_tree_wins(parent, W)
        pw += ev_w
        if False not in statuses and considering:             # tree_search_utils.pyx:113-114
            statuses = considering
        pws = -1
        if int(n_legal[i]) > 0:
            if False in statuses:                             # :122-123
                pws = 1
            elif True in statuses and None not in statuses:   # :124-125 -> update_tree_losses(parent, 1, gameover=True)
                pws = 0
                pv += 1; pgv += 1
        parent[i] = (pv, pw, pgv, pgw, pws, int(threat))
    return stats, flags, parent


# ----------------------------------------------------------------------------
# A deterministic stand-in policy for TREE tests (tests/golden/make_tree_golden.py): a seeded function
# of the position, so that the reference's search and ours can be fed the same "network".
# ----------------------------------------------------------------------------
def synthetic_policy(squares, black_to_move):
    """-> float32 [155], sums to ~1.  Every value is a multiple of 2^-23 (so 1 + p is exact in float32,
    which keeps the reference's mixed float32 / double UCT_multiplier sort keys out of the picture) and
    carries its move index in its low 8 bits (so no two moves tie: the reference breaks ties by its
    history-dependent piece-list order, which a board alone does not determine)."""
    import zlib
    sq = np.asarray(squares, np.int8).reshape(64)
    seed = zlib.crc32(sq.tobytes() + (b"B" if black_to_move else b"W"))
    lg = np.random.RandomState(seed).standard_normal(N_MOVES) * 2.5
    e = np.exp(lg - lg.max())
    p = e / e.sum()
    q = np.round(p * (1 << 15)).astype(np.int64) * 256 + np.arange(N_MOVES)
    return (q.astype(np.float64) / (1 << 23)).astype(np.float32)
