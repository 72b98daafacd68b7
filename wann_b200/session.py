"""Drop-in call surface of the reference's TensorFlow session for the policy net.

Mirrors (names, argument meaning, return types) of
  monte_carlo_tree_search/MCTS.pyx:76-133          class NeuralNetsCombined_128
  Breakthrough_Player/policy_net_utils.pyx:9-16    call_policy_net
  Breakthrough_Player/policy_net_utils.pyx:19-104  instantiate_session*()  -> (sess, y_pred, X)
  policy_net_model/PolicyNet.py:16-36              PolicyNet.call_policy_net / instantiate_session
so that MCTS.pyx, expansion_MCTS_functions.pyx / tree_builder.pyx (policy_net.evaluate(...))
and policy_net_utils.pyx (sess.run(y_pred, feed_dict={X: batch})) work unchanged on top of the
B200 kernels.  Documented divergence: inputs longer than 16,384 return the correct [N,155]
array (the reference's multi-chunk concatenation, MCTS.pyx:124, is broken)."""
import os

import numpy as np

from . import boards
from . import weights as W
from .evaluator import PolicyEvaluator


class _Endpoint(object):
    """Stands in for a TF tensor handle (y_pred / X): identifies a net inside a Session."""

    def __init__(self, net, kind):
        self.net, self.kind = net, kind

    def __repr__(self):
        return "<wann_b200 %s of net %r>" % (self.kind, self.net)


class Session(object):
    """tf.Session-shaped object: run(y_pred, feed_dict={X: batch}) and close()."""

    def __init__(self, evaluators):
        self._nets = evaluators          # name -> PolicyEvaluator
        self._closed = False

    def run(self, fetches, feed_dict=None):
        if self._closed:
            raise RuntimeError("Attempted to use a closed Session.")
        if not isinstance(fetches, _Endpoint) or fetches.kind != "y_pred":
            raise TypeError("fetches must be a y_pred endpoint returned by instantiate_session*()")
        feed_dict = feed_dict or {}
        X = None
        for k, v in feed_dict.items():
            if isinstance(k, _Endpoint) and k.kind == "X" and k.net == fetches.net:
                X = v
        if X is None:
            raise ValueError("feed_dict has no value for the input placeholder of net %r" % fetches.net)
        planes = np.asarray(X)
        if planes.ndim == 3:
            planes = planes[None]
        if planes.shape[1:] != (8, 8, 4):
            raise ValueError("Cannot feed value of shape %s for placeholder of shape (?, 8, 8, 4)"
                             % (planes.shape,))
        return eval_planes_as_boards(self._nets[fetches.net], planes)

    def evaluator(self, net):
        return self._nets[net]

    def close(self):
        if not self._closed:
            for ev in set(self._nets.values()):
                ev.close()
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def eval_planes_as_boards(ev, planes):
    """sess.run(y_pred, {X: planes}) / evaluate(already_converted=True) (MCTS.pyx:103-104).  Planes that
    are a valid P/O/E/1 encoding (tools/utils.pyx:842-865) are turned back into side-to-move boards --
    the encoding is invertible, and a side-to-move board IS a White-to-move board -- and go through
    eval_probs, so that they get the same treatment as every other input form (margin refinement,
    coalescing of concurrent small calls) and bit-identical outputs; anything else (soft or malformed
    planes) is evaluated as given."""
    p = np.asarray(planes)
    if p.size and p.ndim == 4:
        own, opp, emp, one = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
        if ((own == 0) | (own == 1)).all() and ((opp == 0) | (opp == 1)).all() and \
                ((emp == 0) | (emp == 1)).all() and (one == 1).all() and (own + opp + emp == 1).all():
            sq = (own + 2 * opp).astype(np.int8).reshape(-1, 64)
            return ev.eval_probs(sq, np.zeros(sq.shape[0], np.uint8))
    return ev.eval_planes(p)


# Default arithmetic of the shim: the CENTRED fp16 mode on the tensor cores (logits within 1e-3 of the
# fp32 graph, profiles/r02_accuracy_fp16c.json) plus the margin-refinement pass, which re-evaluates
# the ~2 % of positions whose two best legal moves are within 0.02 logits with fp32-class arithmetic:
# the policy move (get_best_move, board_utils.pyx:262-277) then equals the fp32 graph's on every
# one of 100,000 test positions.
# Override per call (mode=..., refine_margin=...) or with WANN_B200_MODE / WANN_B200_REFINE_MARGIN.
DEFAULT_MODE = os.environ.get("WANN_B200_MODE", "fp16c")
DEFAULT_REFINE_MARGIN = float(os.environ.get("WANN_B200_REFINE_MARGIN", "0.02"))


def _make(weights, device, mode, final_kernel, refine_margin=None, scope=""):
    mode = DEFAULT_MODE if mode is None else mode
    if refine_margin is None:
        refine_margin = DEFAULT_REFINE_MARGIN if mode in ("fp16", "bf16", "fp16c") else 0.0
    # coalesce=True: the reference calls evaluate concurrently from its ThreadPool(10) + main
    # thread, each with a single node (expansion_MCTS_functions.pyx:105-119); concurrent small
    # calls are merged into one launch inside the C++ runtime.
    w = weights if weights is not None else W.default_weights(scope=scope, final_kernel=final_kernel)
    if isinstance(device, (list, tuple)):          # one process, several GPUs: host batches are sharded by position
        from .multi import MultiGpuEvaluator
        return MultiGpuEvaluator(w, devices=list(device), mode=mode, final_kernel=final_kernel, coalesce=True,
                                 refine_margin=refine_margin)
    return PolicyEvaluator(w, device=device, mode=mode, final_kernel=final_kernel, coalesce=True,
                           refine_margin=refine_margin)


def instantiate_session(weights=None, device=0, mode=None, refine_margin=None):
    """policy_net_utils.pyx:64-76 / PolicyNet.py:25-36 -> (sess, y_pred, X)."""
    sess = Session({"policy": _make(weights, device, mode, 3, refine_margin)})
    return sess, _Endpoint("policy", "y_pred"), _Endpoint("policy", "X")


def _both(weights_white, weights_black, device, mode, final_kernel, refine_margin=None):
    # policy_net_utils.pyx:22-25,43-46: the dual-net checkpoints keep their variables under the
    # 'white_net/' and 'black_net/' scopes; with explicit weights, one dict may serve both colours
    white = _make(weights_white, device, mode, final_kernel, refine_margin, scope="white_net/")
    if weights_black is None and weights_white is not None:
        black = white
    else:
        black = _make(weights_black, device, mode, final_kernel, refine_margin, scope="black_net/")
    sess = Session({"white_net": white, "black_net": black})
    return (sess, _Endpoint("white_net", "y_pred"), _Endpoint("white_net", "X"),
            _Endpoint("black_net", "y_pred"), _Endpoint("black_net", "X"))


def instantiate_session_both(weights_white=None, weights_black=None, device=0, mode=None, refine_margin=None):
    """policy_net_utils.pyx:19-39 -> (sess, y_pred_white, X_white, y_pred_black, X_black)."""
    return _both(weights_white, weights_black, device, mode, 3, refine_margin)


def instantiate_session_both_128(weights_white=None, weights_black=None, device=0, mode=None,
                                 final_kernel=1, refine_margin=None):
    """policy_net_utils.pyx:40-60: the `_128` nets (last conv 1x1, :135)."""
    return _both(weights_white, weights_black, device, mode, final_kernel, refine_margin)


def call_policy_net(board_representation, weights=None):
    """policy_net_utils.pyx:9-16 (deprecated one-shot call)."""
    sess, y_pred, X = instantiate_session(weights)
    try:
        if isinstance(board_representation, list):
            return sess.run(y_pred, feed_dict={X: board_representation})
        return sess.run(y_pred, feed_dict={X: [board_representation]})
    finally:
        sess.close()


class NeuralNetsCombined_128(object):
    """MCTS.pyx:76-133.  evaluate() keeps the reference signature and its three input forms."""

    BATCH_SIZE = 16384   # MCTS.pyx:96 (kept as the internal chunk size)

    def __init__(self, weights_white=None, weights_black=None, device=0, mode=None, final_kernel=1,
                 refine_margin=None):
        (self.sess, self.output_white, self.input_white,
         self.output_black, self.input_black) = instantiate_session_both_128(
            weights_white, weights_black, device, mode, final_kernel, refine_margin)

    def _net(self, is_player):
        # MCTS.pyx:111-116: is_player -> the 'white' (winner) net, else the 'black' net
        return self.sess.evaluator("white_net" if is_player else "black_net")

    def evaluate(self, game_nodes, player_color=None, is_player=True, already_converted=False):
        ev = self._net(is_player)
        if already_converted:                                   # MCTS.pyx:103-104
            return eval_planes_as_boards(ev, np.asarray(game_nodes))
        if player_color is not None:                            # MCTS.pyx:98-99: one board + colour
            sq = boards.board_to_squares(game_nodes)[None]
            bl = np.array([boards.color_is_black(player_color)], np.uint8)
        else:                                                   # MCTS.pyx:100-102: list of nodes
            sq, bl = self._nodes_to_arrays(game_nodes)
        return ev.eval_probs(sq, bl)

    @staticmethod
    def _nodes_to_arrays(game_nodes):
        sq = boards.boards_to_squares([node['game_board'] for node in game_nodes])
        bl = np.fromiter((boards.color_is_black(node['color']) for node in game_nodes), np.uint8,
                         len(game_nodes))
        return sq, bl

    def evaluate_children(self, game_nodes, is_player=True, num_top=0):
        """Fused variant for callers that only need the ranked legal priors
        (what enumerate_update_and_prune consumes, tree_builder.pyx:428-445).  num_top keeps the best
        num_top legal children (get_top_children's argument, board_utils.pyx:279-291; 0 = all)."""
        sq, bl = self._nodes_to_arrays(game_nodes)
        return self._net(is_player).eval_topk(sq, bl, k=num_top)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        self.sess.close()                                       # MCTS.pyx:132-133
