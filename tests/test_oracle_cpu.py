"""CPU suite (-m "not gpu"): pins the ORACLE against the reference's golden vectors and
checks host logic + that the C-ABI library loads and exports every declared symbol.
No compute call on the CUDA library happens here (there is no GPU)."""
import os
import re

import numpy as np
import pytest

from oracle import c_oracle as C
from oracle import oracle as O
from oracle.ref_loader import load_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- golden vectors (reference-made)
def test_move_tables_match_reference_literals(golden):
    white, black = O.move_tables()
    assert list(golden["white_moves"]) == white       # tools/utils.pyx:126-141
    assert list(golden["black_moves"]) == black       # tools/utils.pyx:300-315
    assert list(golden["gen_white"]) == white         # tools/utils.pyx:480-516
    assert list(golden["gen_black"]) == black
    for i in range(155):
        assert O.move_index(white[i], "White") == i   # tools/utils.pyx:605-641
        assert O.move_index(black[i], "Black") == i


def test_encode_matches_reference_planes(golden):
    planes = O.encode_planes(golden["squares"], golden["black"])
    assert planes.dtype == np.int8 and (planes == golden["planes"]).all()
    assert (C.encode(golden["squares"], golden["black"]) == golden["planes"]).all()
    # one-hot invariant (self_play_logs_to_data_structures.py:292-293) + bias plane
    assert (planes[..., :3].sum(-1) == 1).all() and (planes[..., 3] == 1).all()


def test_legal_mask_matches_reference(golden):
    m = O.legal_mask(golden["squares"], golden["black"])
    assert (m == golden["legal"].astype(bool)).all()
    assert (m == golden["legal_pa"].astype(bool)).all()
    assert (C.legal_mask(golden["squares"], golden["black"]) == m).all()
    assert not m[:, 154].any()                         # 'no-move' is never enumerated


def test_known_answers_opening_and_debug_board(golden):
    # SURVEY 8c KATs: opening has 22 legal moves = indices 22..43 for either colour;
    # planes identical for both colours; debug board has 28 legal moves each side.
    sq, bl = golden["squares"], golden["black"]
    for k in (0, 1):
        assert list(np.nonzero(golden["legal"][k])[0]) == list(range(22, 44))
    assert (golden["planes"][0] == golden["planes"][1]).all()
    white_dbg = {110, 111, 93, 95, 68, 69, 80, 82, 58, 83, 84, 85, 30, 31, 32, 8, 10, 5, 6, 2, 3, 4,
                 14, 15, 16, 17, 18, 19}
    black_dbg = {19, 18, 16, 14, 7, 6, 5, 4, 3, 2, 38, 36, 60, 58, 54, 53, 51, 49, 47, 46, 73, 71,
                 104, 103, 102, 120, 119, 118}
    assert set(np.nonzero(O.legal_mask(sq[2:3], bl[2:3])[0])[0]) == white_dbg
    assert set(np.nonzero(O.legal_mask(sq[3:4], bl[3:4])[0])[0]) == black_dbg
    # Black planes == 180-degree rotation of White planes with P/O swapped
    pw, pb = golden["planes"][2], golden["planes"][3]
    rot = pw[::-1, ::-1]
    assert (pb[..., 0] == rot[..., 1]).all() and (pb[..., 1] == rot[..., 0]).all()


@pytest.mark.skipif(load_ref() is None, reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_against_live_compiled_reference():
    ref = load_ref()
    sq, bl = O.random_positions(60, seed=123)
    planes = O.encode_planes(sq, bl)
    mask = O.legal_mask(sq, bl)
    for i in range(len(sq)):
        color = "Black" if bl[i] else "White"
        board = O.board_from_squares(sq[i], white_to_move=0 if bl[i] else 1)
        assert (ref.utils.convert_board_to_2d_matrix_POEB(board, color) == planes[i]).all()
        idx = sorted(ref.utils.index_lookup_by_move(m)
                     for m in ref.board_utils.enumerate_legal_moves(board, color))
        assert idx == list(np.nonzero(mask[i])[0])
        # get_best_move (board_utils.pyx:262-277) == argmax of prior over legal indices
        probs = np.random.RandomState(i).dirichlet(np.ones(155)).astype(np.float32)[None]
        mv = ref.board_utils.get_best_move(board, probs)
        assert ref.utils.index_lookup_by_move(mv) == O.top1_legal(probs, mask[i:i + 1])[0]


# ---------------------------------------------------------------- forward restatement
def test_forward_numpy_vs_c_oracle_and_explicit_conv():
    sq, bl = O.random_positions_fast(24, seed=3)
    w = O.synthetic_weights("trained_like", 2)
    planes = O.encode_planes(sq, bl)
    lg, pr = O.forward_fp32(planes, w)
    lg_c, pr_c = C.forward(planes, w)
    assert np.abs(lg - lg_c).max() < 2e-4 and np.abs(pr - pr_c).max() < 1e-5
    assert np.allclose(pr.sum(1), 1, atol=1e-5)
    # explicit NHWC/HWIO cross-correlation for one output element (SAME = zero pad 1)
    x = planes[:1].astype(np.float64)
    k = w["hidden_layer/1/weights"].astype(np.float64)
    y, xx, co = 0, 7, 5
    acc = 0.0
    for i in range(3):
        for j in range(3):
            yy, xj = y + i - 1, xx + j - 1
            if 0 <= yy < 8 and 0 <= xj < 8:
                acc += (x[0, yy, xj] * k[i, j, :, co]).sum()
    ref = max(acc + w["hidden_layer/1/bias"][co], 0)
    got = O.forward_fp32(planes[:1], w, return_all=True)[2][0][0, y, xx, co]
    assert abs(got - ref) < 1e-5


def test_forward_final_kernel_1_variant():
    sq, bl = O.random_positions_fast(8, seed=4)
    w = O.synthetic_weights("trained_like", 3, final_kernel=1)
    planes = O.encode_planes(sq, bl)
    lg, _ = O.forward_fp32(planes, w)
    lg_c, _ = C.forward(planes, w, final_k=1)
    assert np.abs(lg - lg_c).max() < 2e-4


def test_ranked_children_semantics():
    sq, bl = O.random_positions_fast(16, seed=9)
    mask = O.legal_mask(sq, bl)
    probs = np.random.RandomState(0).dirichlet(np.ones(155), size=16).astype(np.float32)
    probs[:, 30] = probs[:, 31]                      # force exact ties -> ascending index
    idx, pri, cnt = O.ranked_children(probs, mask)
    idx_c, pri_c, cnt_c = C.rank(probs, mask)
    assert (idx == idx_c).all() and (pri == pri_c).all() and (cnt == cnt_c).all()
    for i in range(16):
        k = cnt[i]
        assert k == mask[i].sum() and (idx[i, k:] == 255).all()
        assert set(idx[i, :k]) == set(np.nonzero(mask[i])[0])
        assert (np.diff(pri[i, :k]) <= 0).all()                       # descending, no renorm
        assert (pri[i, :k] == probs[i, idx[i, :k]]).all()
        assert idx[i, 0] == O.top1_legal(probs[i:i + 1], mask[i:i + 1])[0]


def test_tree_step_oracle_semantics():
    """The self-play step restatement: greedy takes ranked slot 0; sampling is a pure function of
    (seed, step, game); winning moves (index > 131) restart the game; everything else applies the move."""
    sq, bl = O.random_positions_fast(64, seed=12)
    mask = O.legal_mask(sq, bl)
    probs = np.random.RandomState(1).dirichlet(np.ones(155), size=64).astype(np.float32)
    idx, pri, cnt = O.ranked_children(probs, mask)
    s1, b1, ch, fin = O.tree_step(sq, bl, idx, pri, cnt, greedy=True)
    assert (ch == idx[:, 0]).all() and np.array_equal(fin, (ch > 131).astype(np.uint8))
    kids, flg = O.expand_children(sq, bl, ch[:, None], np.ones(64, np.uint8))
    live = fin == 0
    assert np.array_equal(s1[live], kids[live, 0]) and (b1[live] == (bl[live] ^ 1)).all()
    assert (s1[~live] == O.initial_squares()).all() and (b1[~live] == 0).all()
    a = O.tree_step(sq, bl, idx, pri, cnt, greedy=False, seed=7, step=3)
    b = O.tree_step(sq, bl, idx, pri, cnt, greedy=False, seed=7, step=3)
    c = O.tree_step(sq, bl, idx, pri, cnt, greedy=False, seed=7, step=4)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and not np.array_equal(a[2], c[2])
    assert mask[np.arange(64), a[2]].all()
    # empirical frequencies follow the priors: one position, many (seed, step) draws
    i = int(np.argmax(cnt))
    rep = lambda x: np.repeat(x[i:i + 1], 1, axis=0)
    hits = np.zeros(int(cnt[i]))
    for s in range(2000):
        hits[list(idx[i]).index(O.tree_step(rep(sq), rep(bl), rep(idx), rep(pri), rep(cnt), seed=1, step=s)[2][0])] += 1
    p = pri[i, :cnt[i]] / pri[i, :cnt[i]].sum()
    assert np.abs(hits / 2000 - p).max() < 0.04
    assert O.splitmix64(0) == 0xE220A8397B1DCDAF      # published SplitMix64 first output for seed 0


def test_oracle_matches_the_references_own_serialized_graph():
    """The reference ships its inference graph twice (cppModels/policyNet.pb and the checkpoint's
    model.meta, TensorFlow 1.0.0).  Executing THAT node list (ops, edges, strides/padding/data_format,
    the Reshape constant, variable shapes) with textbook op kernels gives the oracle's logits and
    probabilities: the oracle's reading of policy_net_utils.pyx is the graph TensorFlow built."""
    from oracle import graph_interp as G
    pb, meta = G.load("policyNet_pb"), G.load("model_meta")
    assert [dict(n, name="") for n in pb[:-1]] == [dict(n, name="") for n in meta[:-1]]
    assert [n["name"] for n in pb[:-1]] == [n["name"] for n in meta[:-1]]
    assert (pb[-1]["op"], pb[-1]["inputs"]) == (meta[-1]["op"], meta[-1]["inputs"]) == ("Softmax", ["output_layer/output"])
    ops = [n["op"] for n in pb]
    assert ops.count("Conv2D") == 5 and ops.count("Relu") == 5 and ops.count("MatMul") == 1
    var_shapes = {n["name"]: tuple(n["attr"]["shape"]) for n in pb if n["op"] == "VariableV2"}
    assert var_shapes == {k: tuple(v) for k, v in O.weight_shapes().items()}
    sq, bl = O.random_positions_fast(12, seed=21)
    sq[0], bl[0] = O.initial_squares(), 0
    planes = O.encode_planes(sq, bl)
    for preset, seed in (("trained_like", 1), ("ref_init", 2)):
        w = O.synthetic_weights(preset, seed)
        vals = G.run(pb, {"Placeholder": planes.astype(np.float32)}, w, return_all=True)
        lg, pr, acts = O.forward_fp32(planes, w, return_all=True)
        assert vals["output_layer/Softmax"].shape == (12, 155)
        assert np.abs(vals["output_layer/output"] - lg).max() <= 2e-5 * np.abs(lg).max()
        assert np.abs(vals["output_layer/Softmax"] - pr).max() <= 2e-6
        for i in range(1, 6):
            a = vals["hidden_layer/%d/Relu" % i]
            assert np.abs(a - acts[i - 1]).max() <= 2e-5 * max(np.abs(a).max(), 1e-6)


def _golden_children(g):
    """children of the seed golden in ascending index order with the NN output gathered as priors"""
    idx, cnt, nn = g["c_idx"], g["n_children"], g["nn"]
    pri = np.zeros(idx.shape, np.float32)
    for i in range(len(idx)):
        pri[i, :cnt[i]] = nn[i, idx[i, :cnt[i]]]
    return idx, pri, cnt


def test_frozen_graphdef_writer_matches_the_references_graph(tmp_path):
    """tf_graphdef.save_frozen_graph: the GraphDef it writes parses with the TensorFlow protobuf schema
    (tensorboard's copy), is node for node the reference's own inference graph (tests/golden/ref_graph.json)
    with every VariableV2 frozen into a Const, executes in the graph interpreter to the oracle's
    probabilities, and reads back to the same weights."""
    from wann_b200 import tf_graphdef as GD
    from oracle import graph_interp as G
    w = O.synthetic_weights("trained_like", 5)
    path = str(tmp_path / "policyNet_frozen.pb")
    GD.save_frozen_graph(path, w)
    back = GD.load_frozen_weights(path)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
    ours = GD.parse_graph(open(path, "rb").read())
    ref = G.load("policyNet_pb")
    assert [n["name"] for n in ours] == [n["name"] for n in ref]
    for a, b in zip(ours, ref):
        assert a["inputs"] == b["inputs"]
        if b["op"] == "VariableV2":
            assert a["op"] == "Const" and list(a["attr"]["value"].shape) == b["attr"]["shape"]
        elif b["op"] == "Const":
            assert a["op"] == "Const" and a["attr"]["value"].tolist() == b["attr"]["value"]["values"]
        else:
            assert a["op"] == b["op"]
            for k, v in b["attr"].items():
                if b["op"] == "Placeholder" and k == "shape":
                    continue                                     # the reference's placeholder leaves it unknown
                assert a["attr"][k] == v, (a["name"], k)
    # the real TensorFlow schema reads it too
    gp = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    g = gp.GraphDef()
    g.ParseFromString(open(path, "rb").read())
    assert len(g.node) == len(ref) and g.versions.producer == 21
    conv = [n for n in g.node if n.op == "Conv2D"][0]
    assert conv.attr["padding"].s == b"SAME" and list(conv.attr["strides"].list.i) == [1, 1, 1, 1]
    k1 = [n for n in g.node if n.name == "hidden_layer/1/weights"][0].attr["value"].tensor
    assert np.array_equal(np.frombuffer(k1.tensor_content, np.float32).reshape(3, 3, 4, 192), w["hidden_layer/1/weights"])
    # and it EXECUTES: frozen graph -> interpreter -> the oracle's probabilities
    nodes = [{"name": n["name"], "op": n["op"], "inputs": n["inputs"],
              "attr": ({"value": {"dtype": str(n["attr"]["value"].dtype), "shape": list(n["attr"]["value"].shape),
                                  "values": n["attr"]["value"].reshape(-1).tolist()}} if n["op"] == "Const"
                       else dict(n["attr"], dtype=n["attr"].get("dtype", "float32")))} for n in ours]
    sq, bl = O.random_positions_fast(6, seed=2)
    planes = O.encode_planes(sq, bl)
    pr = G.run(nodes, {"Placeholder": planes.astype(np.float32)}, {})
    assert np.abs(pr - O.forward_fp32(planes, w)[1]).max() <= 2e-6
    with pytest.raises(KeyError):
        open(path, "wb").write(open("/root/reference/cppModels/policyNet.pb", "rb").read()
                               if os.path.exists("/root/reference/cppModels/policyNet.pb") else b"")
        GD.load_frozen_weights(path)                             # the reference's own .pb is NOT frozen


def _build_cython_binding(tmp_path):
    import subprocess
    import sys
    ex = os.path.join(ROOT, "examples", "cython_binding")
    env = dict(os.environ, WANN_CYTHON_BUILD_DIR=str(tmp_path / "gen"))
    subprocess.check_call([sys.executable, os.path.join(ex, "setup.py"), "-q", "build_ext", "--build-lib", str(tmp_path),
                           "--build-temp", str(tmp_path / "tmp")], cwd=str(tmp_path), env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    sys.path.insert(0, str(tmp_path))
    try:
        import importlib
        return importlib.import_module("wann_eval")
    finally:
        sys.path.pop(0)


def test_cython_binding_compiles_against_the_header(lib_built, tmp_path):
    """include/wann_b200.pxd (the reference is Cython) declares the same ABI as wann_b200.h: the example
    binding a maintainer would add next to MCTS.pyx compiles against both and links to libwann_b200.so;
    without a GPU the class fails loudly (no CPU fallback)."""
    pytest.importorskip("Cython")
    mod = _build_cython_binding(tmp_path)
    assert mod.abi_version() == 200
    import re
    hdr = open(os.path.join(ROOT, "include", "wann_b200.h")).read()
    pxd = open(os.path.join(ROOT, "include", "wann_b200.pxd")).read()
    declared = set(re.findall(r"\b(wann_[a-z0-9_]+)\s*\(", hdr))
    assert declared and declared <= set(re.findall(r"\b(wann_[a-z0-9_]+)\s*\(", pxd))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
            mod.PolicyNet(O.synthetic_weights("trained_like", 1, final_kernel=1))


def test_array_tree_restatement_reproduces_the_references_trees():
    """tests/golden/ref_tree_golden.npz holds whole search trees grown by the compiled reference
    (run_MCTS_with_expansions_simulation, single-threaded, a deterministic synthetic policy) from six
    start positions (plies 0..58; three end up solved), checkpointed after 1..40 simulations (up to 29,955
    nodes), with the move randomly_choose_a_winning_move would play at each checkpoint.  oracle/tree_oracle.py --
    the reference's selection / expansion / statistics / solver restated over flat arrays -- grows the
    SAME trees, node for node, field for field (visits, wins, gameover statistics, win status, the
    checked / being-checked flags, thread counters, child order)."""
    from oracle.tree_oracle import ArrayTree
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_tree_golden.npz"))
    for s, height in enumerate((0, 14, 31, 44, 52, 58)):
        t = ArrayTree(g["start%d_squares" % s], int(g["start%d_black" % s]), O.synthetic_policy, height=height)
        sims = 0
        for cp in g["checkpoints"]:
            while sims < cp:
                t.run_simulation()
                sims += 1
            assert t.evaluations == int(g["start%d_evals%d" % (s, cp)])
            assert np.array_equal(t.dump(), g["start%d_after%d" % (s, cp)]), (s, cp)
            assert t.index[t.best_move()] == int(g["start%d_move%d" % (s, cp)])      # the move the search plays
        if s < 3:                                    # root reuse (assign_root + reset_thread_flag), 15 more simulations
            t.advance_root(int(g["start%d_reuse_move" % s]))
            for _ in range(15):
                t.run_simulation()
            assert np.array_equal(t.dump(), g["start%d_reuse_after15" % s]) and t.evaluations == int(g["start%d_reuse_evals" % s])
            assert t.index[t.best_move()] == int(g["start%d_reuse_next_move" % s])
            t2 = ArrayTree(g["start%d_squares" % s], int(g["start%d_black" % s]), O.synthetic_policy, height=height)
            for _ in range(12):                      # a move that is not in the tree yet (init_new_root)
                t2.run_simulation()
            t2.advance_root(int(g["start%d_newroot_u" % s]))
            t2.advance_root(int(g["start%d_newroot_reply" % s]))
            for _ in range(10):
                t2.run_simulation()
            assert np.array_equal(t2.dump(), g["start%d_newroot_after10" % s]) and t2.evaluations == int(g["start%d_newroot_evals" % s])
    p = O.synthetic_policy(O.initial_squares(), 0)
    assert p.dtype == np.float32 and len(set(p.tolist())) == 155 and abs(float(p.sum()) - 1) < 1e-2
    assert np.array_equal((np.float32(1) + p).astype(np.float64), 1.0 + p.astype(np.float64))   # 1 + p exact in float32


class _CpuSeeds(object):
    """Stands in for PolicyEvaluator.eval_expand_seeded in CPU tests: the same rows, computed by the
    oracle from the deterministic synthetic policy."""

    def eval_expand_seeded(self, sq, bl, with_children=False):
        probs = np.stack([O.synthetic_policy(s, b) for s, b in zip(sq, bl)])
        idx, pri, cnt = O.ranked_children(probs, O.legal_mask(sq, bl))
        st, fl, ps = O.seed_children(sq, bl, idx, pri, cnt)
        ps8 = np.zeros((len(sq), 8), np.int32)
        ps8[:, :6] = ps
        return idx, pri, cnt, None, fl, st, ps8


def test_native_search_forest_reproduces_the_references_trees(lib_built):
    """The PRODUCT tree (wann_forest_*, array_tree.hpp: C++ struct-of-arrays, resumable at the evaluation,
    six trees advancing in lock step in one forest) grows node for node the trees of the compiled
    reference search (tests/golden/ref_tree_golden.npz) -- three open positions, three that the solver
    decides -- checkpoint by checkpoint, and plays the move the reference's search would play."""
    import wann_b200
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_tree_golden.npz"))
    sq = np.stack([g["start%d_squares" % s] for s in range(6)])
    bl = np.array([int(g["start%d_black" % s]) for s in range(6)], np.uint8)
    with wann_b200.SearchForest(sq, bl, heights=[0, 14, 31, 44, 52, 58]) as forest:
        for cp in g["checkpoints"]:
            steps, evals = forest.run(_CpuSeeds(), int(cp))
            assert steps > 0 and evals >= steps
            for s in range(6):
                info = forest.info(s)
                assert info["simulations"] == cp and info["evaluations"] == int(g["start%d_evals%d" % (s, cp)])
                want = g["start%d_after%d" % (s, cp)]
                assert np.array_equal(forest.dump(s), want), (s, cp)
                assert info["nodes"] == len(want) and info["root_visits"] == want[0, 3] and info["root_wins"] == want[0, 4]
                assert info["move_index"] == int(g["start%d_move%d" % (s, cp)])
        sq2, bl2, ids = forest.next(40)
        assert len(sq2) == 0 and len(ids) == 0                     # budget reached: nothing to evaluate
        # root reuse: play the chosen move in the three open positions, search 15 simulations more
        for s in range(3):
            assert forest.info(s)["move_index"] == int(g["start%d_reuse_move" % s])
            forest.advance(s, int(g["start%d_reuse_move" % s]))
        with pytest.raises(wann_b200.WannError):
            forest.advance(0, 154)                                 # no such child
        forest.run(_CpuSeeds(), 15)
        for s in range(3):
            assert np.array_equal(forest.dump(s), g["start%d_reuse_after15" % s]), s
            assert forest.info(s)["move_index"] == int(g["start%d_reuse_next_move" % s])
            assert forest.info(s)["evaluations"] == int(g["start%d_reuse_evals" % s])
    # a move that is NOT in the tree (init_new_root): 12 simulations, "our move" to an unexpanded child, the
    # opponent's answer, 10 more simulations -- the evaluations of the late expansion come through next/apply
    with wann_b200.SearchForest(sq[:3], bl[:3], heights=[0, 14, 31]) as forest:
        forest.run(_CpuSeeds(), 12)
        for s in range(3):
            forest.advance(s, int(g["start%d_newroot_u" % s]))
            assert forest.info(s)["nodes"] == 1                     # the unexpanded child is the whole tree now
            forest.advance(s, int(g["start%d_newroot_reply" % s]))
        forest.run(_CpuSeeds(), 10)
        for s in range(3):
            assert np.array_equal(forest.dump(s), g["start%d_newroot_after10" % s]), s
            assert forest.info(s)["evaluations"] == int(g["start%d_newroot_evals" % s]) and forest.info(s)["simulations"] == 10
    with pytest.raises(wann_b200.WannError):
        f = wann_b200.SearchForest(sq, bl)
        f.next(1)
        f.next(1)                                                  # apply must follow next


def test_search_forest_self_play_follows_the_restatement(lib_built):
    """wann_forest_play: every tree plays its search's move and reuses the subtree; a winning move restarts
    the game.  Three games (an open one, two near their end) stay node for node equal to the array-tree
    restatement advanced the same way, move after move, until they end and restart."""
    import wann_b200
    from oracle.tree_oracle import ArrayTree
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_tree_golden.npz"))
    pick, heights = (1, 3, 5), (14, 44, 58)
    sq = np.stack([g["start%d_squares" % s] for s in pick])
    bl = np.array([int(g["start%d_black" % s]) for s in pick], np.uint8)
    models = [ArrayTree(sq[i], int(bl[i]), O.synthetic_policy, height=heights[i]) for i in range(3)]
    done = [False] * 3
    with wann_b200.SearchForest(sq, bl, heights=heights) as forest:
        finished_games = 0
        for ply in range(12):
            forest.run(_CpuSeeds(), 6)
            want_moves = []
            for i, m in enumerate(models):
                if done[i]:
                    want_moves.append(None)
                    continue
                for _ in range(6):
                    m.run_simulation()
                assert np.array_equal(forest.dump(i), m.dump()), (ply, i)
                want_moves.append(m.best_move())
            moves, fin = forest.play()
            for i, m in enumerate(models):
                if done[i]:
                    continue
                c = want_moves[i]
                assert moves[i] == m.index[c] and bool(fin[i]) == bool(m.gameover[c])
                if m.gameover[c]:
                    done[i] = True                                   # the forest restarted this tree from the opening
                    assert forest.info(i)["nodes"] == 1 and np.array_equal(forest.next(0)[0], np.zeros((0, 64), np.int8))
                else:
                    m.advance_root(int(moves[i]))
            finished_games += int(fin.sum())
        assert finished_games >= 2 and done[1] and done[2]
        # a time budget instead of a simulation count (the reference's time_to_think): when it returns every
        # tree is between two simulations, so moves can be played
        steps, evals = forest.run_for(_CpuSeeds(), 0.3)
        assert steps > 0 and evals > 0 and len(forest.next(0)[0]) == 0
        moves, fin = forest.play()
        assert (moves != 255).all()
        root0 = forest.dump(0)
        assert root0[0, 0] == -1 and forest.info(0)["nodes"] == len(root0)   # compacted: only the reused subtree is kept


def test_child_seeding_oracle_vs_reference_golden(seed_golden):
    """641 positions expanded ONCE by the compiled reference tree code (expand_descendants_to_depth_wrt_NN
    -> get_child / update_child / assign_children / eval_children / update_win_status_from_children) with
    a synthetic NN output: the oracle's seed_children reproduces every child's visits / wins /
    gameover_visits / gameover_wins / gameover / win_status and the root's resulting statistics."""
    g = seed_golden
    sq, bl = g["squares"], g["black"]
    idx, pri, cnt = _golden_children(g)
    assert np.array_equal(cnt, O.legal_mask(sq, bl).sum(1))
    stats, flags, parent = O.seed_children(sq, bl, idx, pri, cnt)
    live = idx != 255
    assert np.array_equal(stats[live], g["c_stats"][live].astype(np.int32))
    assert np.array_equal((flags & O.SEED_WIN)[live] != 0, g["c_gameover"][live] != 0)
    ws = np.where(flags & O.SEED_WIN, 0, np.where(flags & O.SEED_DOOMED, 1, -1))
    assert np.array_equal(ws[live], g["c_win"][live])
    assert np.array_equal(parent[:, :4], g["root_stats"].astype(np.int32))
    assert np.array_equal(parent[:, 4], g["root_win"])
    assert np.allclose(1.0 + pri[live].astype(np.float64), g["c_uct"][live], rtol=0, atol=1.2e-7)
    # every branch is exercised by the fixture
    for bit in (O.SEED_WIN, O.SEED_CAPTURE, O.SEED_CLOSE_TO_HOME, O.SEED_GAME_SAVING, O.SEED_DOOMED, O.SEED_EVAL_ONE):
        assert ((flags & bit) != 0).sum() > 10
    assert set(np.unique(parent[:, 4])) == {-1, 0, 1} and 0 < parent[:, 5].sum() < len(sq)
    # the reference's child order (sort by UCT_multiplier descending, tree_builder.pyx:557) is the
    # ranked order of ranked_children: identical where 1 + prior separates neighbours; a capture close
    # to home carries a float32-rounded multiplier (tree_search_utils.pyx:978), so it may swap with a
    # neighbour whose prior is within one float32 ulp of 1.0 (1.2e-7) -- documented tie-class divergence
    ridx, rpri, rcnt = O.ranked_children(g["nn"], O.legal_mask(sq, bl))
    same = 0
    for i in range(len(sq)):
        k = int(cnt[i])
        uct_of = dict(zip(g["c_idx"][i, :k].tolist(), g["c_uct"][i, :k].tolist()))
        ref_u = np.array([uct_of[j] for j in g["order"][i, :k].tolist()])
        our_u = np.array([uct_of[j] for j in ridx[i, :k].tolist()])
        assert (np.diff(ref_u) <= 1.2e-7).all()      # mixed float32 / double keys compare in float32 (numpy scalars)
        assert (np.diff(our_u) <= 2.4e-7).all()
        assert sorted(ridx[i, :k].tolist()) == sorted(g["order"][i, :k].tolist())
        assert ridx[i, 0] == g["order"][i, 0] or abs(ref_u[0] - our_u[0]) <= 2.4e-7      # best_child
        same += np.array_equal(ridx[i, :k], g["order"][i, :k])
    assert same >= 0.9 * len(sq)


def test_bf16_round_matches_torch():
    import torch
    x = np.random.RandomState(0).standard_normal(4096).astype(np.float32) * 3
    assert (O.bf16_round(x) == torch.tensor(x).bfloat16().float().numpy()).all()


# ---------------------------------------------------------------- host logic / boundary
def test_library_loads_and_exports_every_declared_symbol(lib_built):
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "wann_b200.h")).read()
    declared = set(re.findall(r"\b(wann_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"wann_ctx", "wann_weights"}
    assert len(declared) >= 14
    lib = ctypes.CDLL(lib_built)
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export %s" % name
    from wann_b200 import _lib
    assert set(_lib.EXPORTS) == declared
    assert _lib.load().wann_version() >= 200


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wann_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # no import / path / library reference to oracle/ (prose mentions are fine)
                assert not re.search(r"(from|import)\s+oracle|oracle[./\\]|liboracle|oracle_", src), f


def test_no_cpu_fallback_without_gpu(lib_built, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import wann_b200
    monkeypatch.setenv("WANN_B200_SYNTHETIC", "1")
    with pytest.raises(wann_b200.WannError):
        wann_b200.PolicyEvaluator()
    with pytest.raises(wann_b200.WannError):
        wann_b200.NeuralNetsCombined_128()


def test_board_helpers_and_move_tables(golden):
    from wann_b200 import boards
    assert boards.WHITE_MOVES == list(golden["white_moves"])
    assert boards.BLACK_MOVES == list(golden["black_moves"])
    for i in (0, 21, 22, 77, 153, 154):
        for color in ("White", "Black"):
            mv = boards.move_lookup_by_index(i, color)
            assert boards.index_lookup_by_move(mv) == i
    sq = golden["squares"][2]
    board = boards.squares_to_board(sq)
    assert (boards.board_to_squares(board) == sq).all()
    assert boards.color_is_black("Black") == 1 and boards.color_is_black("White") == 0


def test_weight_presets_and_validation():
    from wann_b200 import weights as W
    w = W.synthetic("trained_like", 1)
    o = O.synthetic_weights("trained_like", 1)
    for k in W.NAMES:
        assert (w[k] == o[k]).all()                  # product preset == oracle preset (same seed)
    assert sum(v.size for v in w.values()) == 1014812      # SURVEY 8a.a7 parameter count
    W.validate(w)
    bad = dict(w)
    bad["hidden_layer/2/weights"] = bad["hidden_layer/2/weights"][:, :, :10]
    with pytest.raises(ValueError):
        W.validate(bad)
    with pytest.raises(KeyError):
        W.validate({})


def test_shard_range_partition():
    from wann_b200.sharding import shard_range, tree_owner
    for n in (0, 1, 7, 8, 1024, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [tree_owner(t, 8) for t in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wann_b200.sharding import gather_counters, shard_range
    sq, bl = O.random_positions_fast(101, seed=1)
    lo, hi = shard_range(len(sq), rank, world)
    # each rank works on its own slice only (no data-path collective); stand-in work = the
    # oracle's legal-move count, then the end-of-run counter gather
    local = int(O.legal_mask(sq[lo:hi], bl[lo:hi]).sum())
    g = gather_counters([hi - lo, local])
    q.put((rank, g.tolist(), int(O.legal_mask(sq, bl).sum())))
    dist.destroy_process_group()


def test_multi_gpu_evaluator_shards_by_position_with_fake_devices():
    """MultiGpuEvaluator (one process, several GPUs): contiguous slices per device, outputs assembled in
    input order, small calls go to ONE device round-robin, a device's error surfaces in the caller.
    The evaluators here are stand-ins (evaluator_factory seam): this is host logic, no kernels."""
    from wann_b200.multi import MultiGpuEvaluator, plan_shards
    assert plan_shards(0, 8) == [] and plan_shards(100, 8) == [(0, 0, 100)]
    assert plan_shards(1024, 8, 512) == [(0, 0, 512), (1, 512, 1024)]
    sh = plan_shards(65536 + 5, 8)
    assert [s[0] for s in sh] == list(range(8)) and sh[0][1] == 0 and sh[-1][2] == 65541
    assert all(a[2] == b[1] for a, b in zip(sh, sh[1:])) and max(h - l for _, l, h in sh) - min(h - l for _, l, h in sh) <= 1

    class Fake(object):
        def __init__(self, dev):
            self.dev, self.calls, self.closed = dev, [], False

        def eval_probs(self, sq, bl, return_logits=False, out=None):
            self.calls.append(len(sq))
            if self.dev == 3 and len(sq) == 700:
                raise RuntimeError("device 3 failed")
            out[:] = sq[:, :1].astype(np.float32) + 1000 * self.dev          # row id + which device served it
            return (out, out * 2) if return_logits else out

        def eval_topk(self, sq, bl, out=None, k=0):
            out[0][:] = self.dev
            out[1][:] = sq[:, :1]
            out[2][:] = 7
            return out

        def stats(self):
            return {"positions": sum(self.calls)}

        def close(self):
            self.closed = True

    m = MultiGpuEvaluator(devices=[0, 1, 2, 3], min_per_gpu=100, evaluator_factory=Fake)
    n = 1003
    sq = np.zeros((n, 64), np.int8)
    sq[:, 0] = np.arange(n) % 100
    bl = np.zeros(n, np.uint8)
    pr, lg = m.eval_probs(sq, bl, return_logits=True)
    assert pr.shape == (n, 155) and np.array_equal(pr[:, 0] % 1000, np.arange(n) % 100)
    assert np.array_equal(pr[:, 0] // 1000, np.repeat([0, 1, 2, 3], [251, 251, 251, 250])) and np.array_equal(lg, pr * 2)
    idx, pri, cnt = m.eval_topk(sq, bl)
    assert np.array_equal(idx[:, 0], np.repeat([0, 1, 2, 3], [251, 251, 251, 250])) and (cnt == 7).all()
    served = [int(m.eval_probs(sq[:5], bl[:5])[0, 0] // 1000) for _ in range(8)]
    assert sorted(set(served)) == [0, 1, 2, 3]                                # small calls rotate over the devices
    assert m.stats()["positions"] == n + 8 * 5
    with pytest.raises(RuntimeError, match="device 3 failed"):
        m.eval_probs(np.zeros((2800, 64), np.int8), np.zeros(2800, np.uint8))
    m.close()
    assert all(ev.closed for ev in m.evs)


def test_two_rank_gloo_sharding_and_counter_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, g, total in res:
        assert len(g) == 2 and g[0][0] + g[1][0] == 101       # every unit owned exactly once
        assert g[0][1] + g[1][1] == total                     # sharded work == whole-job work


# ---------------------------------------------------------------- TF-V2 bundle I/O (host only)
SURVEY_INDEX = {   # (offset, size, masked CRC32C) decoded from policy_net_model/model.index
    "hidden_layer/1/bias": (8, 768, 0x5104bcb9), "hidden_layer/1/weights": (2312, 27648, 0x6f2c5271),
    "hidden_layer/2/bias": (85256, 768, 0xa8156cc6), "hidden_layer/2/weights": (87560, 1327104, 0xc0bd8ad7),
    "hidden_layer/3/bias": (4068872, 768, 0xbeb65717), "hidden_layer/3/weights": (4071176, 1327104, 0x24038a77),
    "hidden_layer/4/bias": (8052488, 768, 0xd96d363a), "hidden_layer/4/weights": (8054792, 1327104, 0x2b93de69),
    "hidden_layer/5/bias": (12036104, 4, 0xc1d762cd), "hidden_layer/5/weights": (12036116, 6912, 0xdc238e50),
    "output_layer/bias": (12056852, 620, 0xf13a2743), "output_layer/weights": (12058712, 39680, 0xa72e0fa6),
}


def test_tf_bundle_roundtrip_and_crc(lib_built, tmp_path):
    from wann_b200 import tf_bundle as T
    from wann_b200 import weights as W
    assert T.crc32c(b"123456789") == 0xE3069283               # CRC-32C check value
    assert T.crc32c(b"") == 0
    w = W.synthetic("trained_like", 4)
    prefix = str(tmp_path / "model")
    T.save_policy_weights(prefix, w)
    r = T.load_policy_weights(prefix)
    assert all((r[k] == w[k]).all() for k in W.NAMES)
    idx = T.read_index(prefix + ".index")
    assert idx["hidden_layer/2/weights"]["shape"] == (3, 3, 192, 192)
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[123456] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        T.load_policy_weights(prefix)                          # corruption is detected, not loaded
    T.save_policy_weights(prefix, w, scope="white_net/")
    assert (T.load_policy_weights(prefix, scope="white_net/")["output_layer/bias"] == w["output_layer/bias"]).all()
    with pytest.raises(KeyError):
        T.load_policy_weights(prefix)                          # un-scoped names are absent now


@pytest.mark.skipif(not os.path.exists("/root/reference/policy_net_model/model.index"),
                    reason="reference tree not present")
def test_tf_bundle_reads_the_reference_index(lib_built):
    from wann_b200 import tf_bundle as T
    e = T.read_index("/root/reference/policy_net_model/model.index")
    assert len(e) == 38                                        # 12 params + 24 Adam slots + 2 beta powers
    for name, (off, size, crc) in SURVEY_INDEX.items():
        assert (e[name]["offset"], e[name]["size"], e[name]["crc32c"]) == (off, size, crc)
    assert max(v["offset"] + v["size"] for v in e.values()) == 12177752
    with pytest.raises(FileNotFoundError):                     # the shard itself is not shipped
        T.load_policy_weights("/root/reference/policy_net_model/model")


def test_child_construction_oracle_vs_reference_golden(golden):
    """the oracle's child boards / capture / win flags equal the reference's
    move_piece_update_piece_arrays + check_for_winning_move on every legal move of 64 cases."""
    sq, bl = golden["squares"][:64], golden["black"][:64]
    m = golden["legal"][:64].astype(bool)
    idx = np.full((64, 48), 255, np.uint8)
    cnt = m.sum(1).astype(np.uint8)
    for i in range(64):
        idx[i, :cnt[i]] = np.nonzero(m[i])[0]
    kids, flg = O.expand_children(sq, bl, idx, cnt)
    assert np.array_equal(kids, golden["child_squares"]) and np.array_equal(flg, golden["child_flags"])
    assert (flg & O.FLAG_CAPTURE).any() and (flg & O.FLAG_WIN).any()


def test_forward_against_an_independent_torch_restatement():
    """Third, independent restatement of the graph with torch CPU ops (conv2d on NCHW with
    OIHW kernels = w.permute(3,2,0,1), padding 1) -- the oracle must agree with it."""
    from oracle import torch_oracle as T
    sq, bl = O.random_positions_fast(32, seed=21)
    for fk in (3, 1):
        w = O.synthetic_weights("trained_like", 7, final_kernel=fk)
        planes = O.encode_planes(sq, bl)
        logits, probs = T.forward(planes, w)
        lg, pr = O.forward_fp32(planes, w)
        assert np.abs(lg - logits).max() < 2e-4 and np.abs(pr - probs).max() < 1e-5


def test_bulk_board_conversion_equals_per_square_conversion(golden):
    from wann_b200 import boards
    sq = golden["squares"][:100]
    dicts = [boards.squares_to_board(s) for s in sq]
    assert np.array_equal(boards.boards_to_squares(dicts), sq)
    assert boards.boards_to_squares([]).shape == (0, 64)
    bad = boards.squares_to_board(sq[0])
    bad[3]["c"] = "x"
    with pytest.raises(ValueError):
        boards.boards_to_squares([bad])


def test_missing_checkpoint_fails_loudly_and_dual_net_bundles_load_by_scope(tmp_path, monkeypatch):
    """(1) the reference-shaped constructors must not fall back to random weights silently
    (the reference fails when its checkpoint is missing, policy_net_utils.pyx:68-73); (2) the dual-net
    bundles keep their variables under 'white_net/' and 'black_net/' (policy_net_utils.pyx:22-25,43-46)."""
    from wann_b200 import weights as W, tf_bundle
    for k in ("WANN_B200_CHECKPOINT", "WANN_B200_FROZEN_GRAPH", "WANN_B200_SYNTHETIC"):
        monkeypatch.delenv(k, raising=False)
    with pytest.raises(W.MissingCheckpoint):
        W.default_weights()
    monkeypatch.setenv("WANN_B200_SYNTHETIC", "1")
    assert np.array_equal(W.default_weights()["output_layer/bias"], W.synthetic("trained_like", 1)["output_layer/bias"])
    monkeypatch.delenv("WANN_B200_SYNTHETIC")
    white, black = W.synthetic("trained_like", 3, final_kernel=1), W.synthetic("trained_like", 4, final_kernel=1)
    both = {"white_net/" + k: v for k, v in white.items()}
    both.update({"black_net/" + k: v for k, v in black.items()})
    prefix = str(tmp_path / "dual")
    tf_bundle.save_policy_weights(prefix, both)
    monkeypatch.setenv("WANN_B200_CHECKPOINT", prefix)
    got_w = W.default_weights(scope="white_net/", final_kernel=1)
    got_b = W.default_weights(scope="black_net/", final_kernel=1)
    for k in white:
        assert np.array_equal(got_w[k], white[k]) and np.array_equal(got_b[k], black[k])
    assert not np.array_equal(got_w["hidden_layer/2/weights"], got_b["hidden_layer/2/weights"])
    with pytest.raises(KeyError):
        W.default_weights(scope="", final_kernel=1)             # no unscoped net in a dual bundle
    single = str(tmp_path / "single")                           # a single-net bundle serves both colours
    tf_bundle.save_policy_weights(single, white)
    monkeypatch.setenv("WANN_B200_CHECKPOINT", single)
    assert np.array_equal(W.default_weights(scope="black_net/", final_kernel=1)["output_layer/weights"],
                          white["output_layer/weights"])


def test_centred_fp16_emulation_halves_the_fp16_error():
    """The arithmetic idea behind WANN_MODE_FP16C, on the CPU: storing ReLU outputs as a - c (c = bucket
    mean, halo = -c, bias absorbs sum(w*c)) and giving hidden_layer/1 and /5 their weights as hi + lo
    pairs roughly halves the error of plain fp16 operands; split layers reduce it further."""
    from wann_b200 import weights as W
    w = W.synthetic("trained_like", 1)
    sq, bl = O.random_positions_fast(192, seed=11)
    planes = O.encode_planes(sq, bl)
    lg, pr, acts = O.forward_fp32(planes, w, return_all=True)
    off = np.zeros((4, 24), np.float32)
    pos = np.zeros((4, 192), np.uint8)
    for l in range(4):                                          # the calibration of wann_api.cu, restated
        a = acts[l].reshape(-1, 192).astype(np.float64)
        m, v = a.mean(0), a.var(0)
        order = np.argsort(-v, kind="stable")                   # 3 groups of 64 by descending variance
        for g in range(3):
            grp = order[64 * g:64 * (g + 1)]
            grp = grp[np.argsort(m[grp], kind="stable")]        # 8 buckets of 8 by ascending mean
            for j in range(8):
                bk = grp[8 * j:8 * (j + 1)]
                bk = bk[np.argsort(-v[bk], kind="stable")]
                pos[l][bk] = 64 * g + 8 * np.arange(8) + j
                off[l][8 * g + j] = np.float32(np.float16(m[bk].mean()))
    rms = lambda d: float(np.sqrt((d.astype(np.float64) ** 2).mean()) / lg.std())   # noqa: E731
    ident = np.tile(np.arange(192, dtype=np.uint8), (4, 1))
    e_plain = rms(O.forward_16bit_emulated(planes, w, "fp16")[0] - lg)
    e_zero = rms(O.forward_fp16c_emulated(planes, w, np.zeros((4, 24), np.float32), ident)[0] - lg)
    e_cen = rms(O.forward_fp16c_emulated(planes, w, off, pos)[0] - lg)
    e_split = rms(O.forward_fp16c_emulated(planes, w, off, pos, split_mask=14)[0] - lg)
    # the partial split: lo weights for the top-variance third of the inputs x third of the outputs of every
    # 192->192 layer (+11 % MMAs) removes a good part of what the full split (+100 %) removes
    e_part = rms(O.forward_fp16c_emulated(planes, w, off, pos, split_mask=14, lo_n=64, lo_kgroups=1)[0] - lg)
    assert e_zero < e_plain and e_cen < 0.65 * e_plain and e_split < 0.8 * e_cen, (e_plain, e_zero, e_cen, e_split)
    assert e_split < e_part < 0.93 * e_cen, (e_cen, e_part, e_split)
    assert e_cen < 1e-3
    # the data-aware rounding of the 192->192 layers' weights (wann_b200/csrc/gptq.hpp, restated in
    # oracle.gptq_round): second moments of each layer's centred 3x3 input windows on HALF of the positions,
    # error measured on the OTHER half -- it removes more than the partial split does, without any lo stage
    cal = slice(0, 96)
    w_eff = {}
    for layer in (2, 3, 4):
        q = pos[layer - 2].astype(np.int64)
        c = off[layer - 2][(q // 64) * 8 + q % 8]                       # per TF channel
        a = acts[layer - 2][cal].astype(np.float64) - c
        xp = np.empty((a.shape[0], 10, 10, 192))
        xp[:] = -c
        xp[:, 1:-1, 1:-1, :] = a
        X = np.concatenate([xp[:, dy:dy + 8, dx:dx + 8, :] for dy in range(3) for dx in range(3)], axis=-1).reshape(-1, 1728)
        H = X.T @ X / X.shape[0]
        wt = w["hidden_layer/%d/weights" % layer].reshape(1728, 192).T    # [out, (tap, in)]
        w_eff[layer] = O.gptq_round(H, wt, 0.01).T.reshape(3, 3, 192, 192).astype(np.float32)
        assert np.array_equal(w_eff[layer], w_eff[layer].astype(np.float16).astype(np.float32))
    test_half = slice(96, 192)
    rms2 = lambda d: float(np.sqrt((d.astype(np.float64) ** 2).mean()) / lg[test_half].std())   # noqa: E731
    e_rtn = rms2(O.forward_fp16c_emulated(planes[test_half], w, off, pos)[0] - lg[test_half])
    e_gptq = rms2(O.forward_fp16c_emulated(planes[test_half], w, off, pos, w_eff=w_eff)[0] - lg[test_half])
    assert e_gptq < 0.9 * e_rtn, (e_rtn, e_gptq)
    print("fp16c emulation, held-out half: rel-RMS round-to-nearest %.3e, data-aware %.3e" % (e_rtn, e_gptq))


def test_data_aware_weight_rounding_equals_its_restatement(lib_built):
    """wann_gptq_round (wann_b200/csrc/gptq.hpp, host code) against oracle.gptq_round on a small layer with
    strongly correlated inputs: same rounded weights, every unprotected entry an fp16 value within 2 ulp of the
    largest weight, and an output error energy far below round-to-nearest's -- also on held-out inputs."""
    from wann_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(5)
    n, rows, rank = 96, 24, 6
    F = rng.normal(size=(n, rank)) * rng.uniform(0.3, 2.0, size=rank)
    draw = lambda m: rng.normal(size=(m, rank)) @ F.T + 0.05 * rng.normal(size=(m, n)) + 0.2   # noqa: E731
    X, X2 = draw(4 * n), draw(4 * n)
    H = (X.T @ X / X.shape[0]).astype(np.float32)
    W = (rng.normal(size=(rows, n)) * 0.034).astype(np.float32)
    exact = (rng.random((rows, n)) < 0.1).astype(np.uint8)
    ref = O.gptq_round(H, W, 0.01, exact)
    Wc = W.copy()
    L.check(lib.wann_gptq_round(H.ctypes.data, n, Wc.ctypes.data, rows, exact.ctypes.data, 0.01))
    assert (ref.astype(np.float32) == Wc).mean() > 0.99
    free = exact == 0
    assert np.array_equal(Wc[free], Wc[free].astype(np.float16).astype(np.float32))
    assert np.array_equal(Wc[~free], W[~free]) or np.abs(Wc[~free] - W[~free]).max() < 1e-4   # compensation only
    rtn = np.where(free, O.f16_round(W), W).astype(np.float64)
    # the weights move about as far as round-to-nearest moves them (just not independently of each other)
    assert np.abs(Wc - W).max() <= 2 * float(np.spacing(np.float16(np.abs(W).max())))
    assert np.sqrt(((Wc - W)[free] ** 2).mean()) < 2 * np.sqrt(((rtn - W)[free] ** 2).mean())
    energy = lambda Q, Z: float((((W - Q) @ Z.T) ** 2).mean())       # noqa: E731
    assert energy(Wc, X) < 0.1 * energy(rtn, X) and energy(Wc, X2) < 0.2 * energy(rtn, X2)
    # a matrix that is not positive definite is refused
    bad = -np.eye(4, dtype=np.float32)
    w4 = np.ones((1, 4), np.float32)
    assert lib.wann_gptq_round(bad.ctypes.data, 4, w4.ctypes.data, 1, None, 0.0) != 0


def test_concurrent_processes_do_not_live_lock_the_calibration(lib_built):
    """The ranks of a multi-GPU job calibrate at the same moment on shared host cores.  Four processes running the
    rounding (wann_gptq_round, n = 768) at once must take about as long as four runs one after the other -- an
    earlier OpenMP version (a spinning barrier per Cholesky column) needed two orders of magnitude longer."""
    import subprocess
    import sys
    import time
    code = ("import sys, time, numpy as np\n"
            "sys.path.insert(0, %r)\n"
            "from wann_b200 import _lib as L\n"
            "lib = L.load()\n"
            "rng = np.random.default_rng(3)\n"
            "n, rows = 768, 32\n"
            "X = rng.normal(size=(2 * n, 12)) @ rng.normal(size=(12, n)) + 0.05 * rng.normal(size=(2 * n, n))\n"
            "H = (X.T @ X / X.shape[0]).astype(np.float32)\n"
            "W = (rng.normal(size=(rows, n)) * 0.03).astype(np.float32)\n"
            "t0 = time.time()\n"
            "L.check(lib.wann_gptq_round(H.ctypes.data, n, W.ctypes.data, rows, None, 0.01))\n"
            "print(time.time() - t0)\n") % ROOT
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    alone = float(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout)
    t0 = time.time()
    procs = [subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, text=True) for _ in range(4)]
    times = [float(p.communicate()[0]) for p in procs]
    assert all(p.returncode == 0 for p in procs)
    assert max(times) < 20 * max(alone, 0.05) + 2.0, (alone, times)


def test_shim_turns_valid_planes_back_into_boards():
    """sess.run(y_pred, {X: planes}) / evaluate(already_converted=True): valid P/O/E/1 planes take the
    boards path (refinement + coalescing apply), anything else is evaluated as planes."""
    from wann_b200 import session

    class Fake(object):
        def eval_probs(self, sq, bl):
            return ("boards", sq.copy(), bl.copy())

        def eval_planes(self, p):
            return ("planes", p)
    sq, bl = O.random_positions_fast(40, seed=3)
    planes = O.encode_planes(sq, bl)
    kind, got_sq, got_bl = session.eval_planes_as_boards(Fake(), planes)
    assert kind == "boards" and not got_bl.any()
    assert np.array_equal(O.encode_planes(got_sq, got_bl), planes)        # the same network input
    assert session.eval_planes_as_boards(Fake(), planes.astype(np.float32))[0] == "boards"
    soft = planes.astype(np.float32)
    soft[0, 0, 0, 0] = 0.5
    assert session.eval_planes_as_boards(Fake(), soft)[0] == "planes"
    bad = planes.copy()
    bad[1, 2, 3, 3] = 0
    assert session.eval_planes_as_boards(Fake(), bad)[0] == "planes"


def test_forest_play_refuses_mid_expansion_and_restarts_unsearched_trees(lib_built):
    """wann_forest_play (1) must not treat a tree that is waiting for an evaluation as a won game
    (it returns WANN_E_INVALID and leaves every tree alone), and (2) restarts a tree whose root has no
    children (never searched, or a finished game) instead of leaving it stuck with move 255."""
    import wann_b200
    from wann_b200 import _lib as L
    sq, bl = O.random_positions_fast(3, seed=21)
    with wann_b200.SearchForest(sq, bl) as forest:
        moves, fin = forest.play()                                  # nothing searched yet: all three restart
        assert (moves == 255).all() and (fin == 1).all()
        assert all(forest.info(t)["nodes"] == 1 for t in range(3))
        s, b, ids = forest.next(2)                                  # every tree now waits for its first evaluation
        assert len(s) == 3
        before = [forest.dump(t).copy() for t in range(3)]
        with pytest.raises(L.WannError):
            forest.play()
        # (the pending rows are still owed: apply them, then the trees can move)
        idx, pri, cnt = O.ranked_children(np.stack([O.synthetic_policy(x, int(c)) for x, c in zip(s, b)]), O.legal_mask(s, b))
        seeds, flags, parent = O.seed_children(s, b, idx, pri, cnt)
        parent8 = np.zeros((3, 8), np.int32)
        parent8[:, :6] = parent
        for t in range(3):
            assert np.array_equal(forest.dump(t), before[t])
        forest.apply(idx, pri, cnt, flags, seeds, parent8)
        forest.run(_CpuSeeds(), 2)
        moves, fin = forest.play()
        assert (moves != 255).all()


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference runs WITHOUT a GPU (the oracle port on the host cores) and prints exactly one
    JSON line on stdout with the keys the driver reads; the metric, unit and direction are the ones of the GPU arm."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, WANN_B200_BENCH_CPU_SECONDS="0.2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], env=env, capture_output=True, timeout=300, check=True).stdout.decode()
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "policy_net_leaf_evals_per_sec"
    assert line["unit"] == "positions/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "positions/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_any_fp32_implementation_of_the_graph_is_the_same_reference_for_the_gates():
    """The reference's logits are unpinned (TensorFlow 1.0 and the weight shard are absent), so the question is how
    much the choice of fp32 IMPLEMENTATION can matter for north_star's 1e-3 gates.  Three fp32 implementations with
    three different summation orders (numpy/OpenBLAS, plain C, torch/oneDNN) are compared with the graph evaluated
    in float64: all sit within a few 1e-6 of it -- more than two orders of magnitude below the gate, so a result
    within 1e-3 of one fp32 implementation is within 1.01e-3 of any other, Eigen's included."""
    from oracle import torch_oracle as T
    w = O.synthetic_weights("trained_like", 1)
    sq, bl = O.random_positions_fast(512, seed=5)
    planes = O.encode_planes(sq, bl)
    l64, p64 = O.forward_fp32(planes, w, dtype=np.float64)
    for name, (lg, pr) in (("numpy", O.forward_fp32(planes, w)), ("C", C.forward(planes, w)),
                           ("torch", T.forward(planes, w, threads=4))):
        d = np.abs(lg - l64)
        rel = np.abs(pr - p64) / p64
        assert d.max() / np.abs(l64).max() < 1e-5, name
        assert np.median(rel) < 1e-5 and rel.max() < 1e-4, name
