"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the
oracle on the same seeded inputs, against the committed reference-made golden fixtures, and --
at BASELINE.json's full sizes -- through size-independent properties.

Tolerances (written here, derived from north_star and the physics of the operand formats):
  planes / legal masks / move indices / ranked indices ...... bit-exact
  fp32 mode   logits, probs vs fp32 oracle ................. 1e-5 (max |d| / max |ref|)
  fp32_tc     logits 1e-5, probs 2e-5 (fp16 hi/lo split on the tensor cores, fp32-class)
  fp16 mode   logits vs fp32 oracle ........................ 2e-3   (measured 1.4e-3: fp16 operands alone miss 1e-3)
  fp16c mode  logits vs fp32 oracle ........................ 1e-3   north_star's tolerance ("1e-3 relative"):
              max |d| / max |ref| <= 1e-3 and RMS(d) / sigma(ref) <= 1e-3 on 100,000 positions; probabilities:
              median elementwise relative error and median per-row relative L2 error <= 1e-3 (the mode as
              created = the benchmark's configuration); top-1 >= 99.9 % (100 % with refinement)
  bf16 mode   logits vs fp32 oracle ........................ 2e-2   (bf16 has 8 mantissa bits;
              a 5-layer K=1728 stack cannot reach 1e-3 with bf16 operands -- DESIGN.md)
  fp16/bf16   vs the oracle's emulation of the SAME roundings  5e-3 / 5e-3 (tight check of the kernel)
Logits/probabilities have no reference-made golden values (PARITY UNPINNED, see oracle/oracle.py)."""
import os
import threading

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nerr(got, ref):
    return float(np.abs(got.astype(np.float64) - ref).max() / (np.abs(ref).max() + 1e-30))


@pytest.fixture(scope="module")
def weights():
    from wann_b200 import weights as W
    return W.synthetic("trained_like", 1)


@pytest.fixture(scope="module")
def suite():
    sq, bl = O.random_positions_fast(1024, seed=20260601)     # BASELINE config 2 size
    return sq, bl


@pytest.fixture(scope="module")
def ref1k(suite, weights):
    sq, bl = suite
    planes = O.encode_planes(sq, bl)
    lg, pr, acts = O.forward_fp32(planes, weights, return_all=True)
    return dict(planes=planes, logits=lg, probs=pr, acts=acts, mask=O.legal_mask(sq, bl))


@pytest.fixture(scope="module")
def ev_bf16(weights):
    import wann_b200
    ev = wann_b200.PolicyEvaluator(weights, mode="bf16")
    yield ev
    ev.close()


@pytest.fixture(scope="module")
def ev_fp16(weights):
    import wann_b200
    ev = wann_b200.PolicyEvaluator(weights, mode="fp16")
    yield ev
    ev.close()


@pytest.fixture(scope="module")
def ev_fp32(weights):
    import wann_b200
    ev = wann_b200.PolicyEvaluator(weights, mode="fp32")
    yield ev
    ev.close()


# ------------------------------------------------------------------ bit-exact byte/index work
def test_encode_and_legal_mask_bit_exact_vs_reference_golden(ev_bf16, golden):
    assert (ev_bf16.encode(golden["squares"], golden["black"]) == golden["planes"]).all()
    assert (ev_bf16.legal_mask(golden["squares"], golden["black"]) == golden["legal"].astype(bool)).all()


def test_encode_and_legal_mask_bit_exact_vs_oracle(ev_bf16, suite, ref1k):
    sq, bl = suite
    assert (ev_bf16.encode(sq, bl) == ref1k["planes"]).all()
    assert (ev_bf16.legal_mask(sq, bl) == ref1k["mask"]).all()


def test_opening_position_config1(ev_fp32, ev_bf16, weights):
    # BASELINE config 1: batch = 1 on the opening position, both colours
    sq = np.stack([O.initial_squares()] * 2)
    bl = np.array([0, 1], np.uint8)
    lg_ref, pr_ref = O.forward_fp32(O.encode_planes(sq, bl), weights)
    for i in range(2):
        pr = ev_fp32.eval_probs(sq[i:i + 1], bl[i:i + 1])
        assert pr.shape == (1, 155) and nerr(pr, pr_ref[i:i + 1]) < 1e-5
        idx, pri, cnt = ev_bf16.eval_topk(sq[i:i + 1], bl[i:i + 1])
        assert cnt[0] == 22 and sorted(idx[0, :22]) == list(range(22, 44))
    # the opening is colour-symmetric in the side-to-move frame
    assert np.array_equal(ev_fp32.eval_probs(sq[:1], bl[:1]), ev_fp32.eval_probs(sq[1:], bl[1:]))


# ------------------------------------------------------------------ floating point
def test_fp32_mode_matches_oracle_1e5(ev_fp32, suite, ref1k):
    sq, bl = suite
    for layer in (1, 3, 5):
        a = ev_fp32.debug_activations(sq[:64], bl[:64], layer)
        assert nerr(a, ref1k["acts"][layer - 1][:64]) < 1e-5
    pr, lg = ev_fp32.eval_probs(sq, bl, return_logits=True)
    assert nerr(lg, ref1k["logits"]) < 1e-5
    assert nerr(pr, ref1k["probs"]) < 1e-5
    assert np.abs(pr.sum(1) - 1).max() < 1e-5


@pytest.mark.parametrize("mode,tol_emu,tol_fp32", [("bf16", 5e-3, 2e-2), ("fp16", 5e-3, 2e-3)])
def test_tensor_core_modes(mode, tol_emu, tol_fp32, ev_bf16, ev_fp16, suite, ref1k, weights):
    ev = ev_bf16 if mode == "bf16" else ev_fp16
    sq, bl = suite
    n = 256                                                    # BASELINE config 3 size
    lg_e, pr_e, acts_e = O.forward_16bit_emulated(ref1k["planes"][:n], weights, mode, return_all=True)
    a1 = ev.debug_activations(sq[:n], bl[:n], 1)
    # conv1 of 0/1 planes: the products are exact; the bias enters the fp32 accumulator FIRST (the layer's opening
    # MMA), the emulation adds it last, so a stored value may differ by one ulp of the operand format -- rarely
    ulp = np.spacing(np.abs(acts_e[0]).astype(np.float16)).astype(np.float32) * (8 if mode == "bf16" else 1)
    assert (np.abs(a1 - acts_e[0]) <= ulp).all() and (a1 != acts_e[0]).mean() < 1e-3
    for layer in (2, 4, 5):
        a = ev.debug_activations(sq[:n], bl[:n], layer)
        assert nerr(a, acts_e[layer - 1]) < tol_emu            # 1 ulp rounding flips at most
    pr, lg = ev.eval_probs(sq, bl, return_logits=True)
    assert nerr(lg[:n], lg_e) < tol_emu
    assert nerr(lg, ref1k["logits"]) < tol_fp32
    assert nerr(pr, ref1k["probs"]) < 2 * tol_fp32
    assert np.abs(pr.sum(1) - 1).max() < 1e-5


@pytest.mark.parametrize("split,lo_n,lo_kg,gptq", [(0, 64, 1, 1), (0, 64, 1, 0), (14, 64, 1, 0), (14, 96, 1, 1), (6, 48, 2, 0),
                                                   (8, 192, 3, 0), (14, 192, 3, 1)])
def test_centred_fp16_mode_matches_the_emulation_of_its_roundings(split, lo_n, lo_kg, gptq, weights, suite, ref1k):
    """WANN_MODE_FP16C: the kernel against oracle.forward_fp16c_emulated fed with the context's own
    calibrated centring (offsets, channel buckets) -- layer by layer, TF channel order -- for the plain
    centred mode, partial hi/lo weight splits (lo stages for output positions < lo_n x the first lo_kg
    64-channel input groups) and full splits; with round-to-nearest weights (gptq = 0: the emulation rounds
    them itself) and with the data-aware rounding (gptq = 1: the emulation multiplies the weights the context
    reports, which are checked to be fp16 values no further from the true weights than 2 ulp of the largest one).
    Differences are accumulation order and one-ulp rounding flips only."""
    import wann_b200
    sq, bl = suite
    n = 256

    def configure(e):
        e.set_option("gptq", gptq)
        e.set_option("lo_n", lo_n)
        e.set_option("lo_kgroups", lo_kg)
        if split:
            e.set_option("split_layers", split)
    with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev:
        configure(ev)
        off, pos, kap = ev.centering()
        # the centring itself: per layer a permutation; 3 groups x 8 buckets of non-negative fp16 offsets,
        # ascending inside a group, the lowest-mean half of the buckets zero
        for l in range(4):
            assert sorted(pos[l].tolist()) == list(range(192))
            assert np.array_equal(off[l], off[l].astype(np.float16).astype(np.float32))
            assert (off[l] >= 0).all() and off[l].max() > 0 and (off[l] == 0).sum() == 12
            for g in range(3):
                nz = off[l][8 * g:8 * g + 8]
                assert (np.diff(nz[nz > 0]) >= 0).all()
        w_eff = None
        if gptq:
            w_eff = ev.effective_weights()
            for layer in (2, 3, 4):
                w_true = weights["hidden_layer/%d/weights" % layer]
                k_in, k_out = pos[layer - 2] < 64 * lo_kg, pos[layer - 1] < lo_n
                covered = np.zeros((3, 3, 192, 192), bool)
                if (split >> (layer - 1)) & 1:
                    covered[:, :, np.ix_(k_in, k_out)[0], np.ix_(k_in, k_out)[1]] = True
                we = w_eff[layer]
                assert np.array_equal(we[~covered], we[~covered].astype(np.float16).astype(np.float32))
                assert np.abs(we - w_true).max() <= 2 * float(np.spacing(np.float16(np.abs(w_true).max())))
                if (~covered).any():
                    assert (we != O.f16_round(w_true))[~covered].mean() > 0.05     # it is not round-to-nearest
        lg_e, pr_e, acts_e = O.forward_fp16c_emulated(ref1k["planes"][:n], weights, off, pos, split_mask=split,
                                                      lo_n=lo_n, lo_kgroups=lo_kg, return_all=True, w_eff=w_eff)
        for layer in (1, 2, 3, 4, 5):
            a = ev.debug_activations(sq[:n], bl[:n], layer)
            assert nerr(a, acts_e[layer - 1]) < 1.5e-3
        pr, lg = ev.eval_probs(sq, bl, return_logits=True)
        assert nerr(lg[:n], lg_e) < 1e-3
        assert nerr(lg, ref1k["logits"]) < 1e-3                    # north_star's tolerance, 1,024 positions
        assert np.abs(pr.sum(1) - 1).max() < 1e-5
        # the legal-move ranking is the oracle's ranking of these probabilities, planes input agrees
        idx, pri, cnt = ev.eval_topk(sq, bl)
        ridx, rpri, rcnt = O.ranked_children(pr, ref1k["mask"])
        assert np.array_equal(idx, ridx) and np.array_equal(cnt, rcnt)
        assert np.array_equal(ev.eval_planes(ref1k["planes"][:300]), pr[:300])
        # calibration is deterministic: a second context has the same centring and the same outputs
        with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev2:
            configure(ev2)
            off2, pos2, _ = ev2.centering()
            assert np.array_equal(off, off2) and np.array_equal(pos, pos2)
            assert np.array_equal(ev2.eval_probs(sq[:300], bl[:300]), pr[:300])
        if split:
            ev.set_option("split_layers", 0)                       # back to one MMA per K-step
            with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev0:
                ev0.set_option("gptq", gptq)
                assert np.array_equal(ev.eval_probs(sq[:300], bl[:300]), ev0.eval_probs(sq[:300], bl[:300]))


def test_centred_mode_calibrates_on_degenerate_weights(weights, suite):
    """A net whose hidden layers are dead (all-zero conv weights: every activation 0, the calibration's second-moment
    matrices singular) still creates, calibrates and evaluates in the fp16c mode -- the affected layers fall back
    to round-to-nearest -- and gives the fp32 mode's answer (softmax of the FC bias row)."""
    import wann_b200
    sq, bl = suite
    w = {k: (np.zeros_like(v) if k.startswith("hidden_layer") else v.copy()) for k, v in weights.items()}
    with wann_b200.PolicyEvaluator(w, mode="fp32") as ev32, wann_b200.PolicyEvaluator(w, mode="fp16c") as ev:
        p32, p = ev32.eval_probs(sq[:64], bl[:64]), ev.eval_probs(sq[:64], bl[:64])
        assert np.isfinite(p).all() and np.abs(p - p32).max() < 1e-6
        off, pos, _ = ev.centering()
        assert (off == 0).all()
    w2 = {k: v.copy() for k, v in weights.items()}
    w2["hidden_layer/3/weights"][:] = 0                    # dead from hidden_layer/3 on: its inputs are alive
    w2["hidden_layer/3/bias"][:] = 0
    with wann_b200.PolicyEvaluator(w2, mode="fp32") as ev32, wann_b200.PolicyEvaluator(w2, mode="fp16c") as ev:
        p32, p = ev32.eval_probs(sq[:64], bl[:64]), ev.eval_probs(sq[:64], bl[:64])
        assert np.isfinite(p).all() and nerr(p, p32) < 1e-3


def test_centred_fp16_mode_meets_the_north_star_gates_on_100k_positions(weights):
    """north_star: logits and probabilities within 1e-3 relative of the fp32 graph, top-1 >= 99.9 % on a
    fixed suite of random legal positions -- measured on 100,000 positions against the fp32 mode (itself
    within 1e-5 of the CPU oracle, test_fp32_mode_matches_oracle_1e5), for the benchmark's configuration
    (fp16c as created: centred activations, weights rounded with error compensation, one MMA per K-step, no
    refinement pass), the same with margin refinement, and -- with round-to-nearest weights -- the partial
    hi/lo split that meets the gates (+11 % MMAs) and the plain centred mode that just misses one."""
    import wann_b200
    from wann_b200 import synthetic
    sq, bl = synthetic.random_positions(100000, seed=20260601)
    with wann_b200.PolicyEvaluator(weights, mode="fp32") as ev:
        p32, l32 = ev.eval_probs(sq, bl, return_logits=True)
        i32, _, c32 = ev.eval_topk(sq, bl)

    def measure(ev):
        p, l = ev.eval_probs(sq, bl, return_logits=True)
        idx, _, cnt = ev.eval_topk(sq, bl)
        d = np.abs(l - l32)
        return dict(maxnorm=d.max() / np.abs(l32).max(), relrms=np.sqrt((d.astype(np.float64) ** 2).mean()) / l32.std(),
                    pmed=np.median(np.abs(p - p32) / p32),
                    prow=np.median(np.linalg.norm(p - p32, axis=1) / np.linalg.norm(p32, axis=1)),
                    top1=(idx[:, 0] == i32[:, 0]).mean(), counts=np.array_equal(cnt, c32))
    with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev:      # the headline configuration
        m = measure(ev)
        assert m["maxnorm"] <= 1e-3 and m["relrms"] <= 1e-3, m       # logits (measured 5.7e-4 / 4.2e-4)
        assert m["pmed"] <= 1e-3 and m["prow"] <= 1e-3, m            # probabilities (measured 7.4e-4 / 7.1e-4)
        assert m["pmed"] <= 8.5e-4, m                                 # ... with room: the data-aware rounding works
        assert m["top1"] >= 0.999 and m["counts"], m                 # policy move (measured 99.96 %)
        ev.set_refine_margin(0.01)                                    # + margin refinement: every policy move agrees
        m = measure(ev)
        assert m["maxnorm"] <= 1e-3 and m["pmed"] <= 1e-3 and m["top1"] == 1.0, m
    with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev:      # round-to-nearest weights
        ev.set_option("gptq", 0)
        m = measure(ev)                                               # plain centred mode: full speed, one gate missed
        assert m["maxnorm"] <= 1e-3 and m["relrms"] <= 1e-3 and m["top1"] >= 0.999, m
        assert 8.5e-4 < m["pmed"] <= 1.25e-3, m                       # (measured 1.01e-3)
        ev.set_option("split_layers", 14)                             # partial hi/lo split (lo_n = 64, lo_kgroups = 1)
        m = measure(ev)
        assert m["maxnorm"] <= 1e-3 and m["relrms"] <= 1e-3, m       # logits (measured 6.2e-4 / 4.9e-4)
        assert m["pmed"] <= 1e-3 and m["prow"] <= 1e-3, m            # probabilities (measured 9.0e-4 / 8.5e-4)
        assert m["top1"] >= 0.999 and m["counts"], m


def test_top1_agreement_vs_fp32_oracle(ev_bf16, ev_fp16, suite, ref1k):
    sq, bl = suite
    top_ref = O.top1_legal(ref1k["probs"], ref1k["mask"])
    # margin (in logit units) between the oracle's best and second-best legal move
    lg = np.where(ref1k["mask"], ref1k["logits"], -np.inf)
    srt = np.sort(lg, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    for ev, raw_min, tol in ((ev_bf16, 0.97, 0.20), (ev_fp16, 0.995, 0.03)):
        idx, pri, cnt = ev.eval_topk(sq, bl)
        agree = idx[:, 0] == top_ref
        assert agree.mean() >= raw_min
        # every disagreement is a near-tie of the oracle itself (margin below the mode's
        # logit tolerance): agreement modulo ties is 100 %
        assert (margin[~agree] < tol).all()


def test_topk_ranking_is_exactly_the_oracle_ranking_of_gpu_probs(ev_bf16, suite, ref1k):
    sq, bl = suite
    pr = ev_bf16.eval_probs(sq, bl)
    idx, pri, cnt = ev_bf16.eval_topk(sq, bl)
    ridx, rpri, rcnt = O.ranked_children(pr, ref1k["mask"])
    assert np.array_equal(idx, ridx) and np.array_equal(cnt, rcnt) and np.array_equal(pri, rpri)


def test_planes_input_paths(ev_bf16, ev_fp32, suite, ref1k):
    sq, bl = suite
    for ev in (ev_bf16, ev_fp32):
        base = ev.eval_probs(sq[:200], bl[:200])
        assert np.array_equal(ev.eval_planes(ref1k["planes"][:200]), base)                   # int8
        assert np.array_equal(ev.eval_planes(ref1k["planes"][:200].astype(np.float32)), base)  # f32


def test_final_kernel_1_variant(weights):
    import wann_b200
    from wann_b200 import weights as W
    w = W.synthetic("trained_like", 5, final_kernel=1)
    sq, bl = O.random_positions_fast(64, seed=8)
    lg_ref, pr_ref = O.forward_fp32(O.encode_planes(sq, bl), w)
    for mode, tol in (("fp32", 1e-5), ("fp16", 2e-3), ("fp16c", 1e-3), ("bf16", 2e-2)):
        with wann_b200.PolicyEvaluator(w, mode=mode, final_kernel=1) as ev:
            pr, lg = ev.eval_probs(sq, bl, return_logits=True)
            assert nerr(lg, lg_ref) < tol


# ------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("n", [0, 1, 2, 7, 8, 9, 15, 17, 63, 65, 255, 300])
def test_ragged_batch_sizes(ev_bf16, suite, n):
    sq, bl = suite
    full = ev_bf16.eval_probs(sq[:304], bl[:304])
    pr = ev_bf16.eval_probs(sq[:n], bl[:n])
    assert pr.shape == (n, 155) and np.array_equal(pr, full[:n])   # row i <-> input i, batch-invariant
    idx, pri, cnt = ev_bf16.eval_topk(sq[:n], bl[:n])
    assert idx.shape == (n, 48) and cnt.shape == (n,)


def test_batch_larger_than_chunk_and_many_pairs(weights):
    import wann_b200
    sq, bl = O.random_positions_fast(512, seed=77)
    sq = np.tile(sq, (9, 1))[:4500]
    bl = np.tile(bl, 9)[:4500]
    with wann_b200.PolicyEvaluator(weights, mode="bf16", chunk=1024) as ev:   # 5 chunks, last ragged
        pr = ev.eval_probs(sq, bl)
    with wann_b200.PolicyEvaluator(weights, mode="bf16") as ev:
        ref = ev.eval_probs(sq, bl)
    assert np.array_equal(pr, ref)
    assert np.array_equal(pr[:512], pr[512:1024])            # same inputs -> same outputs (any tile slot)


def test_extreme_boards(ev_bf16, ev_fp32):
    # maximum legal-move fan-out and minimal material
    sq = np.zeros((3, 64), np.int8)
    sq[0, 8:24] = 1; sq[0, 56:64] = 2           # 16 white pieces with free squares ahead
    sq[1, 27] = 1; sq[1, 36] = 2                # lone pieces
    sq[2, 0:8] = 1; sq[2, 8:16] = 1; sq[2, 16:24] = 2   # fully blocked front: forward illegal, captures legal
    bl = np.zeros(3, np.uint8)
    mask = O.legal_mask(sq, bl)
    assert (ev_bf16.legal_mask(sq, bl) == mask).all()
    idx, pri, cnt = ev_fp32.eval_topk(sq, bl)
    assert list(cnt) == list(mask.sum(1))
    for i in range(3):
        assert set(idx[i, :cnt[i]]) == set(np.nonzero(mask[i])[0])


def test_errors_are_reported_not_swallowed(ev_bf16):
    import wann_b200
    with pytest.raises(wann_b200.WannError):
        ev_bf16.debug_activations(np.zeros((1, 64), np.int8), np.zeros(1, np.uint8), 9)
    with pytest.raises(ValueError):
        ev_bf16.eval_probs(np.zeros((2, 64), np.int8), np.zeros(3, np.uint8))


# ------------------------------------------------------------------ size-independent properties at full size
def test_full_size_properties_65536(ev_bf16):
    import torch
    sq0, bl0 = O.random_positions_fast(4096, seed=31)
    n = 65536                                                   # BASELINE config 4 top of sweep
    sq = torch.from_numpy(np.tile(sq0, (16, 1))).cuda()
    bl = torch.from_numpy(np.tile(bl0, 16)).cuda()
    probs = ev_bf16.eval_probs(sq, bl)
    idx, pri, cnt = ev_bf16.eval_topk(sq, bl)
    torch.cuda.synchronize()
    probs, idx, pri, cnt = probs.cpu().numpy(), idx.cpu().numpy(), pri.cpu().numpy(), cnt.cpu().numpy()
    assert np.isfinite(probs).all() and np.abs(probs.sum(1) - 1).max() < 1e-5      # softmax
    assert np.array_equal(probs[:4096], probs[n - 4096:])                          # determinism / tile-slot invariance
    mask = O.legal_mask(sq0, bl0)
    assert np.array_equal(cnt[:4096], mask.sum(1).astype(np.uint8))                # legality
    k = np.arange(48)[None, :] < cnt[:, None]
    assert (np.diff(pri, axis=1)[k[:, 1:]] <= 0).all()                             # sortedness
    rows = np.arange(n)[:, None].repeat(48, 1)
    assert np.array_equal(pri[k], probs[rows[k], idx[k]])                          # gather, no renormalisation
    assert (idx[~k] == 255).all() and (pri[~k] == 0).all()
    # colour symmetry: a position and its 180-degree rotated, colour-swapped twin are the
    # same input in the side-to-move frame -> identical outputs
    sq_t = np.where(sq0[:, ::-1] == 1, 2, np.where(sq0[:, ::-1] == 2, 1, 0)).astype(np.int8)
    p_a = ev_bf16.eval_probs(sq0, bl0)
    p_b = ev_bf16.eval_probs(sq_t, (1 - bl0).astype(np.uint8))
    assert np.array_equal(p_a, p_b)


# ------------------------------------------------------------------ the reference-facing surface
def test_session_shim_is_a_drop_in(weights, suite, ref1k):
    import wann_b200
    from wann_b200 import boards
    sq, bl = suite
    w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
    nodes = [{"game_board": boards.squares_to_board(sq[i], 0 if bl[i] else 1),
              "color": "Black" if bl[i] else "White"} for i in range(40)]
    with wann_b200.NeuralNetsCombined_128(w1, mode="fp32") as net:
        out = net.evaluate(nodes, is_player=True)                      # tree_builder.pyx:189
        assert isinstance(out, np.ndarray) and out.shape == (40, 155) and out.dtype == np.float32
        one = net.evaluate(nodes[3]["game_board"], nodes[3]["color"])  # MCTS.pyx:57 (Policy mode)
        assert np.array_equal(one[0], out[3])
        conv = net.evaluate(ref1k["planes"][:40], already_converted=True)   # policy_net_metrics.py:101
        assert np.array_equal(conv, out)
        assert np.array_equal(net.evaluate([nodes[5]])[0], out[5])     # tree_builder.pyx:830
        idx, pri, cnt = net.evaluate_children(nodes)
        assert np.array_equal(idx[:, 0], O.top1_legal(out, ref1k["mask"][:40]))
    # the shim's DEFAULT arithmetic (fp16 operands + margin refinement): every policy move equals
    # the fp32 graph's (get_best_move, board_utils.pyx:262-277), probabilities in the 1e-3 class
    all_nodes = [{"game_board": boards.squares_to_board(sq[i], 0 if bl[i] else 1),
                  "color": "Black" if bl[i] else "White"} for i in range(len(sq))]
    _, pr_ref = O.forward_fp32(ref1k["planes"], w1)
    with wann_b200.NeuralNetsCombined_128(w1) as net:
        out = net.evaluate(all_nodes)
        assert nerr(out, pr_ref) < 5e-3          # probabilities: d p = p * d logit, logits are in the 1.5e-3 class
        assert np.array_equal(O.top1_legal(out, ref1k["mask"]), O.top1_legal(pr_ref, ref1k["mask"]))
        assert net.sess.evaluator("white_net").stats()["refined_positions"] > 0
    sess, y_pred, X = wann_b200.instantiate_session(weights, mode="fp32")    # policy_net_utils.pyx:64
    got = sess.run(y_pred, feed_dict={X: [p for p in ref1k["planes"][:16]]})
    assert nerr(got, ref1k["probs"][:16]) < 1e-5
    sess.close()
    with pytest.raises(RuntimeError):
        sess.run(y_pred, feed_dict={X: ref1k["planes"][:1]})


def test_concurrent_callers_are_reentrant(ev_bf16, suite):
    # the reference calls evaluate from 11 threads without a lock (tree_builder.pyx:186-189)
    sq, bl = suite
    ref = ev_bf16.eval_probs(sq, bl)
    errs = []

    def worker(t):
        try:
            for k in range(20):
                i = (37 * t + 11 * k) % 1000
                n = 1 + (t + k) % 9
                out = ev_bf16.eval_probs(sq[i:i + n], bl[i:i + n])
                if not np.array_equal(out, ref[i:i + n]):
                    errs.append((t, k))
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(11)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", ".built")),
                    reason="compiled reference (oracle/_ref) not shipped")
def test_unmodified_reference_tree_code_runs_on_our_evaluator(weights):
    """Drive the reference's own compiled expand_descendants_to_depth_wrt_NN through the real
    boundary (policy_net.evaluate) with our evaluator standing in for the TF session."""
    import time
    import wann_b200
    from oracle.ref_loader import load_ref
    ref = load_ref()
    tb = ref.module("monte_carlo_tree_search.tree_builder")
    TreeNode = ref.module("monte_carlo_tree_search.TreeNode").TreeNode
    w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
    calls = []
    with wann_b200.NeuralNetsCombined_128(w1, mode="fp32") as net:
        class Spy(object):
            def evaluate(self, nodes, **kw):
                out = net.evaluate(nodes, **kw)
                calls.append((len(nodes), out))
                return out
        board = ref.utils.initial_game_board()
        wp, bp = ref.board_utils.initial_piece_arrays()
        root = TreeNode(board, wp, bp, "White", 0, None, 0)
        net.evaluate([root])      # first call allocates the lane workspaces: keep it off the search clock
        sim_info = {"root": root, "time_to_think": 5.0, "start_time": time.time(), "game_tree": [],
                    "main_pid": "not-this-thread", "do_eval": False, "counter": 0, "prev_game_tree_size": 0,
                    "game_num": 0, "eval_times": []}
        tb.expand_descendants_to_depth_wrt_NN([root], 0, 1, sim_info, threading.Lock(), Spy())
    assert calls and calls[0][0] == 1
    kids = root["children"]
    assert kids is not None and len(kids) == 22                  # all legal moves kept
    pri = [k["UCT_multiplier"] for k in kids]
    assert pri == sorted(pri, reverse=True)                      # tree_builder.pyx:557
    sq = O.initial_squares()[None]
    want = O.ranked_children(calls[0][1], O.legal_mask(sq, [0]))
    got_idx = [k["index"] for k in kids]
    assert got_idx == want[0][0, :22].tolist()                   # same children, same order
    assert abs((pri[0] - 1) - want[1][0, 0]) < 1e-6              # 1 + NN_output[idx], tree_search_utils.pyx:986


def test_coalescer_merges_concurrent_small_calls(weights, suite):
    """11 threads x batch-1 calls (the reference's pattern): results identical to direct calls,
    and fewer launches than calls (flat combining in the C++ runtime)."""
    import wann_b200
    sq, bl = suite
    with wann_b200.PolicyEvaluator(weights, mode="bf16") as ev:
        ref_p = ev.eval_probs(sq[:512], bl[:512])
        ref_i, ref_r, ref_c = ev.eval_topk(sq[:512], bl[:512])
    with wann_b200.PolicyEvaluator(weights, mode="bf16", coalesce=True) as ev:
        assert np.array_equal(ev.eval_probs(sq[:3], bl[:3]), ref_p[:3])      # lone caller
        errs = []

        def worker(t):
            try:
                for k in range(40):
                    i = (41 * t + 13 * k) % 500
                    n = 1 + (k % 3 == 0) * (t % 5)
                    if k % 2:
                        out = ev.eval_probs(sq[i:i + n], bl[i:i + n])
                        ok = np.array_equal(out, ref_p[i:i + n])
                    else:
                        a, b, c = ev.eval_topk(sq[i:i + n], bl[i:i + n])
                        ok = np.array_equal(a, ref_i[i:i + n]) and np.array_equal(b, ref_r[i:i + n]) and \
                            np.array_equal(c, ref_c[i:i + n])
                    if not ok:
                        errs.append((t, k))
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))

        ts = [threading.Thread(target=worker, args=(t,)) for t in range(11)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        st = ev.stats()
        assert not errs
        assert st["merged_calls"] == 1 + 11 * 40
        assert st["merged_launches"] < st["merged_calls"]     # concurrency was actually merged
        big = ev.eval_probs(sq[:400], bl[:400])                # > 256 positions bypasses the combiner
        assert np.array_equal(big, ref_p[:400]) and ev.stats()["merged_calls"] == st["merged_calls"]


def test_child_construction_bit_exact(ev_fp32, ev_bf16, golden, suite):
    """children of every ranked legal move, against reference-made golden children
    (board_utils.move_piece_update_piece_arrays) and the oracle."""
    sq, bl = golden["squares"][:64], golden["black"][:64]
    idx, pri, cnt, kids, flg = ev_fp32.eval_expand(sq, bl)
    # golden children are stored in ascending move-index order; ours come in ranked order
    for i in range(64):
        order = np.argsort(idx[i, :cnt[i]], kind="stable")
        assert np.array_equal(kids[i, :cnt[i]][order], golden["child_squares"][i, :cnt[i]])
        assert np.array_equal(flg[i, :cnt[i]][order], golden["child_flags"][i, :cnt[i]])
        assert not kids[i, cnt[i]:].any() and not flg[i, cnt[i]:].any()
    sq, bl = suite
    idx, pri, cnt, kids, flg = ev_bf16.eval_expand(sq, bl)
    okids, oflg = O.expand_children(sq, bl, idx, cnt)
    assert np.array_equal(kids, okids) and np.array_equal(flg, oflg)
    i2, p2, c2 = ev_bf16.eval_topk(sq, bl)
    assert np.array_equal(idx, i2) and np.array_equal(pri, p2) and np.array_equal(cnt, c2)
    # children are valid positions for the other colour: material changes only by captures
    mat = (sq != 0).sum(1)[:, None] - ((flg & 2) != 0)
    k = np.arange(48)[None, :] < cnt[:, None]
    assert np.array_equal((kids != 0).sum(2)[k], mat[k])
    # device-resident path gives the same children (ready to be the next leaf batch)
    import torch
    out = ev_bf16.eval_expand(torch.from_numpy(sq).cuda(), torch.from_numpy(bl).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(out[3].cpu().numpy(), kids) and np.array_equal(out[4].cpu().numpy(), flg)


def test_malformed_boards_do_not_corrupt_memory(ev_bf16, suite):
    """Boards that cannot occur in Breakthrough (random squares, > 16 pieces a side, > 48 'legal'
    moves) are outside the reference's domain; the kernels must stay memory-safe: the mask is still
    the rule-based mask, the ranking is truncated to 48 entries, neighbours are unaffected."""
    rng = np.random.RandomState(5)
    bad = rng.randint(0, 3, size=(64, 64)).astype(np.int8)
    bl = rng.randint(0, 2, size=64).astype(np.uint8)
    assert (ev_bf16.legal_mask(bad, bl) == O.legal_mask(bad, bl)).all()
    assert (ev_bf16.encode(bad, bl) == O.encode_planes(bad, bl)).all()
    sq, b2 = suite
    mix_sq = np.concatenate([sq[:64], bad, sq[64:128]])
    mix_bl = np.concatenate([b2[:64], bl, b2[64:128]])
    idx, pri, cnt, kids, flg = ev_bf16.eval_expand(mix_sq, mix_bl)
    assert cnt.max() <= 48 and np.isfinite(pri).all()
    ref = ev_bf16.eval_expand(np.concatenate([sq[:64], sq[64:128]]), np.concatenate([b2[:64], b2[64:128]]))
    for got, want in zip((idx, pri, cnt, kids, flg), ref):
        assert np.array_equal(np.concatenate([got[:64], got[128:]]), want)


def test_treefarm_steady_state_plays_legal_games(ev_bf16):
    """BASELINE config 5 in miniature: device-resident games, one leaf per game per step; every move
    taken is a legal move of the position it was taken in (oracle), finished games restart."""
    import torch
    from wann_b200.treefarm import TreeFarm
    farm = TreeFarm(ev_bf16, 256, seed=3)
    finished = 0
    for step in range(70):
        sq_before = farm.sq.cpu().numpy()
        bl_before = farm.black.cpu().numpy()
        idx, pri, cnt, k = farm.step()
        idx, cnt, k = idx.cpu().numpy(), cnt.cpu().numpy(), k.cpu().numpy()
        mask = O.legal_mask(sq_before, bl_before)
        chosen = idx[np.arange(256), k]
        assert (k < cnt).all() and mask[np.arange(256), chosen].all()
        kids, flg = O.expand_children(sq_before, bl_before, chosen[:, None].astype(np.uint8), np.ones(256, np.uint8))
        reset = farm.last_reset.cpu().numpy()
        assert np.array_equal(reset, (flg[:, 0] & 1).astype(bool))
        after = farm.sq.cpu().numpy()
        assert np.array_equal(after[~reset], kids[~reset, 0])
        assert (after[reset] == O.initial_squares()).all()
        assert not O.game_over(after).any()
        finished += int(reset.sum())
    assert finished > 0                          # random games of Breakthrough end within ~60 plies


def test_tree_step_fused_matches_oracle(ev_bf16, suite):
    """wann_tree_step (config 5 as ONE call): the child chosen per game (greedy and sampled, counter-based
    RNG), the in-place move and the restart equal the oracle's restatement applied to the ranked
    priors the same context returns for the same boards."""
    import torch
    sq, bl = suite
    dev = torch.device("cuda", ev_bf16.device)
    big_sq, big_bl = np.tile(sq, (20, 1)), np.tile(bl, 20)
    for greedy, n, seed, step in ((True, 300, 0, 0), (False, 1024, 11, 5), (False, 1, 3, 9), (False, 1024, 11, 6),
                                  (False, 20000, 4, 2)):            # 20,000 games: two internal chunks
        s0, b0 = big_sq[:n].copy(), big_bl[:n].copy()
        idx, pri, cnt = ev_bf16.eval_topk(s0, b0)
        want = O.tree_step(s0, b0, idx, pri, cnt, greedy=greedy, seed=seed, step=step)
        ts, tb = torch.from_numpy(s0).to(dev), torch.from_numpy(b0).to(dev)
        chosen, fin = ev_bf16.tree_step(ts, tb, greedy=greedy, seed=seed, step=step)
        torch.cuda.synchronize()
        got = (ts.cpu().numpy(), tb.cpu().numpy(), chosen.cpu().numpy(), fin.cpu().numpy())
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        if not greedy and n > 1:
            assert (want[2] != idx[:n, 0]).any()          # sampling really leaves the policy move sometimes
    # a farm playing whole games through the fused call: always legal, games finish and restart
    from wann_b200.treefarm import TreeFarm
    farm = TreeFarm(ev_bf16, 512, seed=5)
    finished = 0
    for _ in range(80):
        before_sq, before_bl = farm.sq.cpu().numpy(), farm.black.cpu().numpy()
        chosen, fin = farm.step_fused(want_outputs=True)
        chosen, fin = chosen.cpu().numpy(), fin.cpu().numpy().astype(bool)
        mask = O.legal_mask(before_sq, before_bl)
        assert mask[np.arange(512), chosen].all()
        after = farm.sq.cpu().numpy()
        assert (after[fin] == O.initial_squares()).all() and not O.game_over(after).any()
        assert np.array_equal(fin, chosen > 131)
        finished += int(fin.sum())
    assert finished > 0


@pytest.mark.parametrize("mode,margin", [("fp16", 0.02), ("bf16", 0.15), ("fp16c", 0.02)])
def test_margin_refinement_pass(mode, margin, weights, suite, ref1k):
    """refine_margin: positions whose two best LEGAL priors are closer than the margin (the only ones
    where 16-bit operand rounding can flip get_best_move's choice) are re-evaluated in the same call
    with the fp32-class split kernel.  Flagged rows == the fp32_tc mode's rows, every other row is
    untouched, the counter matches, and top-1 then agrees with the fp32 oracle (north_star: >= 99.9 %)."""
    import torch
    import wann_b200
    sq, bl = suite
    n = sq.shape[0]
    with wann_b200.PolicyEvaluator(weights, mode=mode) as plain, \
            wann_b200.PolicyEvaluator(weights, mode="fp32_tc") as exact, \
            wann_b200.PolicyEvaluator(weights, mode=mode, refine_margin=margin) as ev:
        i0, p0, c0 = plain.eval_topk(sq, bl)
        pr0, lg0 = plain.eval_probs(sq, bl, return_logits=True)
        ix, px, cx = exact.eval_topk(sq, bl)
        prx, lgx = exact.eval_probs(sq, bl, return_logits=True)
        ratio = np.float32(np.exp(-np.float32(margin)))
        flagged = (c0 >= 2) & (p0[:, 1] >= p0[:, 0] * ratio)
        edge = np.abs(p0[:, 1] - p0[:, 0] * ratio) <= 1e-6 * p0[:, 0]      # float-rounding of the threshold itself
        assert 0 < flagged.sum() < n // 2
        before = ev.stats()["refined_positions"]
        i1, p1, c1 = ev.eval_topk(sq, bl)
        assert ev.stats()["refined_positions"] - before == flagged.sum() or edge.any()
        pr1, lg1 = ev.eval_probs(sq, bl, return_logits=True)
        f, u = flagged & ~edge, ~flagged & ~edge
        for got, fast, slow in ((i1, i0, ix), (p1, p0, px), (c1, c0, cx), (pr1, pr0, prx), (lg1, lg0, lgx)):
            assert np.array_equal(got[u], fast[u])                        # untouched rows: bit-identical
            assert np.array_equal(got[f], slow[f])                        # refined rows: the fp32_tc result
        # the policy move now agrees with the fp32 oracle
        want = O.top1_legal(ref1k["probs"], ref1k["mask"])
        assert (i1[:, 0] == want).mean() >= 0.999
        # device path, several chunks (65,536 positions per device chunk) and a ragged tail
        big = 70001
        reps = (big + n - 1) // n
        bsq, bbl = np.tile(sq, (reps, 1))[:big], np.tile(bl, reps)[:big]
        dev = torch.device("cuda", ev.device)
        di, dp, dc = ev.eval_topk(torch.from_numpy(bsq).to(dev), torch.from_numpy(bbl).to(dev))
        torch.cuda.synchronize()
        assert np.array_equal(di.cpu().numpy(), np.tile(i1, (reps, 1))[:big])
        assert np.array_equal(dp.cpu().numpy(), np.tile(p1, (reps, 1))[:big])
        assert np.array_equal(dc.cpu().numpy(), np.tile(c1, reps)[:big])
        # children are built from the refined ranking
        ei, ep, ec, kids, flg = ev.eval_expand(sq[:300], bl[:300])
        assert np.array_equal(ei, i1[:300]) and np.array_equal(ep, p1[:300])
        wk, wf = O.expand_children(sq[:300], bl[:300], ei, ec)
        assert np.array_equal(kids, wk) and np.array_equal(flg, wf)
        # small host calls (the host reads the flagged count and launches pass 2 itself) with the fused and
        # with the stand-alone tail (regression: the latter mirrored the count before the flagging ran)
        for fuse in (1, 0):
            ev.set_option("fuse_tail", fuse)
            for lo, m in ((0, 64), (100, 1), (7, 300)):
                si, sp, sc = ev.eval_topk(sq[lo:lo + m], bl[lo:lo + m])
                assert np.array_equal(si, i1[lo:lo + m]) and np.array_equal(sp, p1[lo:lo + m]) and np.array_equal(sc, c1[lo:lo + m])
                assert np.array_equal(ev.eval_probs(sq[lo:lo + m], bl[lo:lo + m]), pr1[lo:lo + m])
        ev.set_option("fuse_tail", 1)
        # switching it off restores the plain path
        ev.set_refine_margin(0)
        i2, p2, c2 = ev.eval_topk(sq, bl)
        assert np.array_equal(i2, i0) and np.array_equal(p2, p0)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_zero_copy_small_host_calls_equal_the_dma_path(mode, weights, suite, ref1k):
    """Host calls of <= 512 positions read their inputs from / write their outputs to pinned mapped
    host memory directly (no cudaMemcpy): bit-identical to the staged DMA path, for every input form."""
    import wann_b200
    sq, bl = suite
    with wann_b200.PolicyEvaluator(weights, mode=mode) as ev:
        for n in (1, 7, 300, 512):
            res = {}
            for zc in (0, 1):
                ev.set_option("zero_copy", zc)
                pr, lg = ev.eval_probs(sq[:n], bl[:n], return_logits=True)
                res[zc] = (pr, lg) + tuple(ev.eval_topk(sq[:n], bl[:n])) + \
                    (ev.eval_planes(ref1k["planes"][:n]), ev.eval_planes(ref1k["planes"][:n].astype(np.float32)))
            for a, b in zip(res[0], res[1]):
                assert np.array_equal(a, b)
        ev.set_option("zero_copy", 1)
        h0 = ev.stats()["h2d_bytes"]
        ev.eval_topk(sq[:5], bl[:5])
        assert ev.stats()["h2d_bytes"] - h0 == 5 * 65


def test_child_seeding_bit_exact(ev_fp32, ev_bf16, seed_golden, suite):
    """wann_eval_expand_seeded: child statistics, flags and parent summaries are integer work -> bit-exact
    against the oracle applied to the ranked priors the same call returns (host and device paths), and
    against the compiled reference's own expansion (golden) when the kernel is fed the golden NN output
    through an identity-like check of the oracle on the same children."""
    import torch
    g = seed_golden
    for ev, (sq, bl) in ((ev_bf16, (g["squares"], g["black"])), (ev_fp32, (g["squares"][:200], g["black"][:200])),
                         (ev_bf16, suite)):
        idx, pri, cnt, kids, flg, st, ps = ev.eval_expand_seeded(sq, bl)
        i0, p0, c0, k0, f0 = ev.eval_expand(sq, bl)
        assert np.array_equal(idx, i0) and np.array_equal(pri, p0) and np.array_equal(kids, k0)
        assert np.array_equal(flg & 3, f0)
        want_st, want_fl, want_ps = O.seed_children(sq, bl, idx, pri, cnt)
        assert np.array_equal(st, want_st) and np.array_equal(flg, want_fl)
        assert np.array_equal(ps[:, :6], want_ps)
        assert np.array_equal(ps[:, 6], ((flg & 1) != 0).sum(1)) and np.array_equal(ps[:, 7], ((flg & 16) != 0).sum(1))
        # seeds only (no child boards), device path
        dev = torch.device("cuda", ev.device)
        out = ev.eval_expand_seeded(torch.from_numpy(sq).to(dev), torch.from_numpy(bl).to(dev), with_children=False)
        torch.cuda.synchronize()
        assert out[3] is None
        for got, want in zip((out[0], out[1], out[2], out[4], out[5], out[6]), (idx, pri, cnt, flg, st, ps)):
            assert np.array_equal(got.cpu().numpy(), want)
    # the integer part that does not depend on the net at all must equal the reference-made golden:
    # which children win, which are doomed / game saving, and the threat flag
    sq, bl = g["squares"], g["black"]
    idx, pri, cnt, _, flg, st, ps = ev_bf16.eval_expand_seeded(sq, bl, with_children=False)
    for i in range(len(sq)):
        k = int(cnt[i])
        order = np.argsort(idx[i, :k])
        assert np.array_equal(idx[i, :k][order], g["c_idx"][i, :k])
        big = g["c_stats"][i, :k, 0] >= 65536                       # win / doomed / game-saving seeds
        assert np.array_equal(st[i, :k][order][big], g["c_stats"][i, :k][big].astype(np.int32))
        ws = np.where(flg[i, :k] & 1, 0, np.where(flg[i, :k] & 16, 1, -1))[order]
        assert np.array_equal(ws, g["c_win"][i, :k])


@pytest.mark.parametrize("mode", ["bf16", "fp16", "fp16c"])
def test_fused_tail_equals_the_standalone_tail_kernel(mode, weights, suite, ref1k):
    """bf16/fp16 modes run the tail (FC-155 -> softmax -> legal gather -> rank) on two extra warps INSIDE
    the conv-stack kernel; the stand-alone tail_kernel (option fuse_tail = 0) must give bit-identical
    probabilities, logits and ranked priors for every batch shape, input form and call path."""
    import torch
    import wann_b200
    sq, bl = suite
    big_sq, big_bl = np.tile(sq, (5, 1))[:4999], np.tile(bl, 5)[:4999]
    with wann_b200.PolicyEvaluator(weights, mode=mode) as ev:
        dev = torch.device("cuda", ev.device)
        for n in (1, 3, 4, 5, 8, 9, 295, 296, 297, 600, 1024, 4999):
            s_, b_ = big_sq[:n], big_bl[:n]
            res = {}
            for fuse in (0, 1):
                ev.set_option("fuse_tail", fuse)
                k0 = ev.stats()["kernel_launches"]
                pr, lg = ev.eval_probs(s_, b_, return_logits=True)
                launches = ev.stats()["kernel_launches"] - k0
                assert launches == (1 if fuse else 2)
                tk = ev.eval_topk(s_, b_)
                dv = ev.eval_topk(torch.from_numpy(s_).to(dev), torch.from_numpy(b_).to(dev))
                torch.cuda.synchronize()
                res[fuse] = (pr, lg) + tuple(tk) + tuple(x.cpu().numpy() for x in dv)
                if n <= 1024:
                    res[fuse] += (ev.eval_planes(ref1k["planes"][:n]),)
            for a, b in zip(res[0], res[1]):
                assert np.array_equal(a, b)


def test_full_size_seeded_expansion_properties(weights):
    """BASELINE's top batch (65,536 positions, plus a ragged tail crossing the device chunk) through the
    widest call -- fused evaluation + margin refinement + child seeding, device resident: size-independent
    invariants on every row and the oracle on a sample."""
    import torch
    import wann_b200
    from wann_b200 import synthetic
    n = 65536 + 17
    base_sq, base_bl = synthetic.random_positions(4096, seed=77)
    reps = (n + 4095) // 4096
    sq, bl = np.tile(base_sq, (reps, 1))[:n], np.tile(base_bl, reps)[:n]
    with wann_b200.PolicyEvaluator(weights, mode="fp16", refine_margin=0.02) as ev:
        dev = torch.device("cuda", ev.device)
        out = ev.eval_expand_seeded(torch.from_numpy(sq).to(dev), torch.from_numpy(bl).to(dev), with_children=False)
        torch.cuda.synchronize()
        idx, pri, cnt, _, flg, st, ps = [x.cpu().numpy() if x is not None else None for x in out]
    mask = O.legal_mask(sq[:4096], bl[:4096])
    assert np.array_equal(cnt[:4096], mask.sum(1))
    live = np.arange(48)[None] < cnt[:, None]
    assert (np.diff(pri, axis=1)[live[:, 1:]] <= 0).all()                       # ranked, no renormalisation
    assert (idx[~live] == 255).all() and (pri[~live] == 0).all() and (st[~live] == 0).all() and (flg[~live] == 0).all()
    assert ((pri * live).sum(1) <= 1.0 + 1e-5).all()
    # identical positions (the pool is tiled) give identical rows, wherever they sit in a chunk or tile
    for a in (idx, pri, cnt, flg, st, ps):
        assert np.array_equal(a[:4096], a[4096:8192]) and np.array_equal(a[:17], a[65536:])
    assert np.array_equal(ps[:, 6], ((flg & 1) != 0).sum(1)) and np.array_equal(ps[:, 7], ((flg & 16) != 0).sum(1))
    assert ((ps[:, 5] == 1) == (((flg & 24) != 0).any(1) | ((ps[:, 5] == 1) & ((flg & 32) == 0).all(1)))).all()
    assert (ps[:, 2] <= ps[:, 0]).all() and (ps[:, 3] <= ps[:, 1]).all()        # gameover stats are a subset
    pick = np.random.RandomState(3).choice(n, 300, replace=False)
    want_st, want_fl, want_ps = O.seed_children(sq[pick], bl[pick], idx[pick], pri[pick], cnt[pick])
    assert np.array_equal(st[pick], want_st) and np.array_equal(flg[pick], want_fl) and np.array_equal(ps[pick, :6], want_ps)


def test_multi_gpu_evaluator_single_process(weights, suite, ev_bf16):
    """One process, every visible GPU (sharded by position, one host thread per GPU): bit-identical to
    one GPU evaluating the whole batch; also the session shim with device=[...]."""
    import torch
    import wann_b200
    sq, bl = suite
    big_sq, big_bl = np.tile(sq, (4, 1))[:3001], np.tile(bl, 4)[:3001]
    devices = list(range(torch.cuda.device_count()))
    want_pr, want_lg = ev_bf16.eval_probs(big_sq, big_bl, return_logits=True)
    want_tk = ev_bf16.eval_topk(big_sq, big_bl)
    with wann_b200.MultiGpuEvaluator(weights, devices=devices, mode="bf16", min_per_gpu=256) as m:
        pr, lg = m.eval_probs(big_sq, big_bl, return_logits=True)
        assert np.array_equal(pr, want_pr) and np.array_equal(lg, want_lg)
        for got, want in zip(m.eval_topk(big_sq, big_bl), want_tk):
            assert np.array_equal(got, want)
        out = m.eval_expand_seeded(big_sq[:700], big_bl[:700], with_children=False)
        ref = ev_bf16.eval_expand_seeded(big_sq[:700], big_bl[:700], with_children=False)
        for got, want in zip(out, ref):
            assert (got is None and want is None) or np.array_equal(got, want)
        assert np.array_equal(m.eval_probs(sq[:3], bl[:3]), want_pr[:3])
        assert m.stats()["positions"] >= 2 * 3001 + 700 + 3
        if len(devices) > 1:
            per = [ev.stats()["positions"] for ev in m.evs]
            assert min(per) > 0                                   # every GPU took a slice
    w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
    with wann_b200.NeuralNetsCombined_128(w1, device=devices, mode="fp32") as multi, \
            wann_b200.NeuralNetsCombined_128(w1, device=0, mode="fp32") as single:
        from wann_b200 import boards
        nodes = [{"game_board": boards.squares_to_board(sq[i], 0 if bl[i] else 1),
                  "color": "Black" if bl[i] else "White"} for i in range(64)]
        assert np.array_equal(multi.evaluate(nodes), single.evaluate(nodes))


@pytest.mark.parametrize("mode", ["bf16", "fp32", "fp32_tc", "fp16c"])
def test_set_weights_swaps_a_live_context(mode, suite):
    """wann_set_weights: a newer checkpoint of the same architecture replaces the weights of a live
    context (every device-side form: fp32 tensors, tensor-core images, split image, bias) -- the results
    afterwards are those of a context created with the new weights, also while other threads evaluate."""
    import wann_b200
    sq, bl = suite
    w_a = wann_b200.weights.synthetic("trained_like", 1)
    w_b = wann_b200.weights.synthetic("trained_like", 2)
    with wann_b200.PolicyEvaluator(w_a, mode=mode, refine_margin=0.05 if mode == "bf16" else 0.0) as ev, \
            wann_b200.PolicyEvaluator(w_b, mode=mode, refine_margin=0.05 if mode == "bf16" else 0.0) as fresh:
        n = 600 if mode != "fp32" else 200
        before = ev.eval_topk(sq[:n], bl[:n])
        want = fresh.eval_topk(sq[:n], bl[:n]) + (fresh.eval_probs(sq[:n], bl[:n]),)
        assert not np.array_equal(before[1], want[1])
        stop, errs = [False], []

        def hammer():
            try:
                while not stop[0]:
                    ev.eval_probs(sq[:64], bl[:64])
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=hammer) for _ in range(3)]
        [t.start() for t in th]
        ev.set_weights(w_b)
        stop[0] = True
        [t.join() for t in th]
        assert not errs
        got = ev.eval_topk(sq[:n], bl[:n]) + (ev.eval_probs(sq[:n], bl[:n]),)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
        with pytest.raises(ValueError):                                   # another architecture: refused by the shim ...
            ev.set_weights(wann_b200.weights.synthetic("trained_like", 2, final_kernel=1))
        _, cw = ev._c_weights(w_b, 3)
        cw.final_kernel = 1                                               # ... and by the C ABI itself
        import ctypes
        assert ev._lib.wann_set_weights(ev._ctx, ctypes.byref(cw)) != 0


def test_cython_binding_evaluates_like_the_shim(tmp_path, suite):
    """The Cython class of examples/cython_binding (cimport wann_b200, nogil call into the C ABI) returns
    what the Python shim returns for the same nodes."""
    pytest.importorskip("Cython")
    import wann_b200
    from wann_b200 import boards
    from test_oracle_cpu import _build_cython_binding
    mod = _build_cython_binding(tmp_path)
    sq, bl = suite
    w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
    nodes = [{"game_board": boards.squares_to_board(sq[i], 0 if bl[i] else 1),
              "color": "Black" if bl[i] else "White"} for i in range(300)]
    with mod.PolicyNet(w1) as net, wann_b200.NeuralNetsCombined_128(w1) as shim:
        got = net.evaluate(nodes)
        assert got.shape == (300, 155) and got.dtype == np.float32
        assert np.array_equal(got, shim.evaluate(nodes))
        assert np.array_equal(net.evaluate(nodes[:1]), got[:1])


def test_soak_random_batches_paths_and_options():
    """tools/gpu_stress.py for a few seconds: random batch sizes, asynchronous device bursts (PDL on / off /
    auto), 4 concurrent host threads, options toggled between bursts, bf16 and fp16 + refinement -- every
    result equals the reference evaluation of the same rows."""
    import subprocess
    import sys
    env = dict(os.environ, SECONDS="10")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_stress.py")], env=env, capture_output=True,
                         text=True, timeout=200)
    assert out.returncode == 0 and "stress ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_search_forest_on_the_gpu_evaluator(weights, suite):
    """BASELINE config 5 with the reference's own search per tree: a forest of native array trees
    (wann_forest_*) advancing in lock step on ONE wann_eval_expand_seeded call per step.  Every tree equals
    the array-tree restatement (oracle/tree_oracle.py -- itself node-for-node the compiled reference's
    tree, tests/golden/ref_tree_golden.npz) grown on the same network one evaluation at a time."""
    import wann_b200
    from oracle.tree_oracle import ArrayTree
    sq, bl = suite
    pick = [0, 5, 17, 33, 101, 250, 400, 777, 901, 1000]
    starts_sq = np.concatenate([O.initial_squares()[None], sq[pick]])
    starts_bl = np.concatenate([[0], bl[pick]]).astype(np.uint8)
    heights = np.arange(len(starts_sq), dtype=np.int32) * 3
    with wann_b200.PolicyEvaluator(weights, mode="fp16", refine_margin=0.02) as ev:
        with wann_b200.SearchForest(starts_sq, starts_bl, heights) as forest:
            steps, evals = forest.run(ev, 6)
            assert steps > 10 and evals > steps            # several trees per step
            policy = lambda s, b: ev.eval_probs(np.asarray(s, np.int8)[None], np.array([b], np.uint8))[0]   # noqa: E731
            for t in range(len(starts_sq)):
                ref = ArrayTree(starts_sq[t], int(starts_bl[t]), policy, height=int(heights[t]))
                for _ in range(6):
                    ref.run_simulation()
                got, want = forest.dump(t), ref.dump()
                assert got.shape == want.shape and np.array_equal(got, want), t
                info = forest.info(t)
                assert info["evaluations"] == ref.evaluations and info["simulations"] == 6
                assert info["move_index"] == ref.index[ref.best_move()]            # the move the search plays
        # self-play on the real network: whole games, subtree reuse, restarts
        with wann_b200.SearchForest(np.tile(O.initial_squares()[None], (64, 1)), np.zeros(64, np.uint8)) as sp:
            evals, games = sp.self_play(ev, 4, 70)
            assert evals > 64 * 70 and games >= 32             # Breakthrough games end within ~60 plies
        # a larger forest just runs: 512 trees, every evaluation a legal expansion
        big = wann_b200.SearchForest(np.tile(sq[:512], (1, 1)), bl[:512])
        steps, evals = big.run(ev, 3)
        nodes = sum(big.info(t)["nodes"] for t in range(512))
        assert evals > 512 * 3 and nodes > evals * 10
        big.close()


def test_no_out_of_bounds_writes_canary(ev_bf16, suite):
    """compute-sanitizer is not available on the GPU pool: carve every device output out of one
    sentinel-filled arena (guard bands before/after each buffer) and check that the kernels wrote
    exactly their own bytes, for ragged sizes that do not fill a tile / warp group / chunk."""
    import ctypes
    import torch
    from wann_b200 import _lib as L
    lib = L.load()
    sq, bl = suite
    for n in (1, 5, 13, 127, 1000):
        sizes = [("idx", n * 48), ("prior", n * 48 * 4), ("cnt", n), ("kids", n * 48 * 64), ("flg", n * 48),
                 ("probs", n * 155 * 4), ("logits", n * 155 * 4)]
        guard = 4096
        total = guard + sum((s + guard + 255) // 256 * 256 for _, s in sizes)
        arena = torch.full((total,), 0xA5, dtype=torch.uint8, device="cuda")
        offs, pos = {}, guard
        for name, s in sizes:
            offs[name] = (pos, s)
            pos += (s + guard + 255) // 256 * 256
        base = arena.data_ptr()
        dsq = torch.from_numpy(sq[:n]).cuda()
        dbl = torch.from_numpy(bl[:n]).cuda()
        ptr = lambda k: ctypes.c_void_p(base + offs[k][0])   # noqa: E731
        L.check(lib.wann_eval_expand(ev_bf16._ctx, dsq.data_ptr(), dbl.data_ptr(), n, 0, ptr("idx"), ptr("prior"),
                                     ptr("cnt"), ptr("kids"), ptr("flg"), L.MEM_DEVICE, ctypes.c_void_p(1)))
        L.check(lib.wann_eval_probs(ev_bf16._ctx, dsq.data_ptr(), dbl.data_ptr(), n, ptr("probs"), ptr("logits"),
                                    L.MEM_DEVICE, ctypes.c_void_p(1)))
        torch.cuda.synchronize()
        host = arena.cpu().numpy()
        inside = np.zeros(total, bool)
        for name, (o, s) in offs.items():
            inside[o:o + s] = True
        assert (host[~inside] == 0xA5).all(), "a kernel wrote outside its output buffer (n=%d)" % n
        # and the in-bounds results are the usual ones
        o, s = offs["cnt"]
        assert np.array_equal(host[o:o + s], ev_bf16.eval_topk(sq[:n], bl[:n])[2])


def test_fp32_tc_mode_fp32_class_accuracy_on_tensor_cores(weights, suite, ref1k):
    """WANN_MODE_FP32_TC: fp16 hi/lo operand splitting, 3 MMAs per K-step.  Must meet the same
    1e-5 bound as the CUDA-core fp32 mode (north_star's fp32-mode tolerance)."""
    import wann_b200
    sq, bl = suite
    with wann_b200.PolicyEvaluator(weights, mode="fp32_tc") as ev:
        for layer in (1, 2, 4, 5):
            a = ev.debug_activations(sq[:64], bl[:64], layer)
            assert nerr(a, ref1k["acts"][layer - 1][:64]) < 1e-5
        pr, lg = ev.eval_probs(sq, bl, return_logits=True)
        assert nerr(lg, ref1k["logits"]) < 1e-5
        # probabilities: measured 1.05e-5 of the largest probability (the tensor core's fp32
        # accumulator truncates when aligning addends); bound 2e-5, logits stay inside 1e-5
        assert nerr(pr, ref1k["probs"]) < 2e-5
        idx, pri, cnt = ev.eval_topk(sq, bl)
        assert np.array_equal(idx[:, 0], O.top1_legal(ref1k["probs"], ref1k["mask"]))
        for n in (1, 3, 4, 5, 9, 300):                      # ragged sizes vs the 4-board supertile
            assert np.array_equal(ev.eval_probs(sq[:n], bl[:n]), pr[:n])
    w1 = wann_b200.weights.synthetic("trained_like", 5, final_kernel=1)
    lg_ref, _ = O.forward_fp32(O.encode_planes(sq[:64], bl[:64]), w1)
    with wann_b200.PolicyEvaluator(w1, mode="fp32_tc", final_kernel=1) as ev:
        assert nerr(ev.eval_probs(sq[:64], bl[:64], return_logits=True)[1], lg_ref) < 1e-5


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", ".built")),
                    reason="compiled reference (oracle/_ref) not shipped")
def test_reference_multithreaded_mcts_search_runs_on_our_evaluator():
    """The reference's own MCTS_with_expansions (ThreadPool(10) + main thread, every evaluate call a
    single node, expansion_MCTS_functions.pyx:77-167) searching on top of NeuralNetsCombined_128:
    returns a legal move; the coalescer merged concurrent calls."""
    import io
    import wann_b200
    from oracle.ref_loader import load_ref
    ref = load_ref()
    mc = ref.module("monte_carlo_tree_search.expansion_MCTS_functions")
    w1 = wann_b200.weights.synthetic("trained_like", 1, final_kernel=1)
    names, sizes = set(), []
    with wann_b200.NeuralNetsCombined_128(w1, mode="bf16") as net:
        class Spy(object):
            def evaluate(self, nodes, **kw):
                names.add(threading.current_thread().name)
                sizes.append(len(nodes))
                return net.evaluate(nodes, **kw)
        board = ref.utils.initial_game_board()
        best_child, move = mc.MCTS_with_expansions(board, "White", 1, 5, None, None, 0, io.StringIO(), "WaNN",
                                                   Spy(), 0)
        st = net.sess.evaluator("white_net").stats()
    legal = ref.board_utils.enumerate_legal_moves(board, "White")
    assert move in legal
    assert len(sizes) > 100 and len(names) >= 2
    assert st["merged_calls"] >= len(sizes) - 16 and st["merged_launches"] <= st["merged_calls"]


@pytest.mark.parametrize("k", [1, 2, 5, 47, 48, 0, 200])
def test_top_k_keeps_the_best_k_children(k, ev_bf16, suite):
    """`k` of wann_eval_topk / _expand / _expand_seeded = num_top of get_top_children
    (board_utils.pyx:279-291) / num_to_keep of the pruning branch (tree_builder.pyx:585-609): the ranked
    list is cut after the k best legal children, children beyond it are neither built nor seeded."""
    import torch
    sq, bl = suite
    n = 300
    full_i, full_p, full_c = ev_bf16.eval_topk(sq[:n], bl[:n])
    kk = k if 1 <= k < 48 else 48
    want_c = np.minimum(full_c, kk)
    want_i, want_p = full_i.copy(), full_p.copy()
    want_i[:, kk:], want_p[:, kk:] = 255, 0
    idx, pri, cnt = ev_bf16.eval_topk(sq[:n], bl[:n], k=k)
    assert np.array_equal(idx, want_i) and np.array_equal(pri, want_p) and np.array_equal(cnt, want_c)
    dev = torch.device("cuda", ev_bf16.device)
    di, dp, dc = ev_bf16.eval_topk(torch.from_numpy(sq[:n]).to(dev), torch.from_numpy(bl[:n]).to(dev), k=k)
    torch.cuda.synchronize()
    assert np.array_equal(di.cpu().numpy(), want_i) and np.array_equal(dc.cpu().numpy(), want_c)
    ei, ep, ec, kids, flg, seeds, parent = ev_bf16.eval_expand_seeded(sq[:n], bl[:n], k=k)
    assert np.array_equal(ei, want_i) and np.array_equal(ec, want_c)
    wk, _ = O.expand_children(sq[:n], bl[:n], ei, ec)
    ws, wf, wp = O.seed_children(sq[:n], bl[:n], ei, ep, ec)
    assert np.array_equal(kids, wk) and np.array_equal(seeds, ws) and np.array_equal(flg, wf)
    assert np.array_equal(parent[:, :6], wp)
    xi, xp, xc, xk, xf = ev_bf16.eval_expand(sq[:n], bl[:n], k=k)
    assert np.array_equal(xi, want_i) and np.array_equal(xk, wk)


@pytest.mark.parametrize("mode,margin", [("fp32", 0.0), ("fp16", 0.05), ("fp32_tc", 0.0), ("fp16c", 0.05)])
def test_small_chunk_host_calls_stay_inside_the_workspaces(mode, margin, weights, suite):
    """with the `chunk` option below 512 a small host call must not take the one-chunk
    zero-copy path past the slot capacity (chunk = 32, n = 77 wrote past d_act / d_h5r); results equal
    those of a default-chunk context, with the fused and the stand-alone tail."""
    import wann_b200
    sq, bl = suite
    with wann_b200.PolicyEvaluator(weights, mode=mode, refine_margin=margin) as ref:
        want = ref.eval_topk(sq[:300], bl[:300]) + (ref.eval_probs(sq[:300], bl[:300]),)
    for chunk in (8, 32):
        with wann_b200.PolicyEvaluator(weights, mode=mode, chunk=chunk, refine_margin=margin) as ev:
            for fuse in ((1, 0) if mode in ("fp16", "fp16c") else (1,)):
                ev.set_option("fuse_tail", fuse)
                for n in (1, 33, 77, 129, 300):
                    got = ev.eval_topk(sq[:n], bl[:n]) + (ev.eval_probs(sq[:n], bl[:n]),)
                    for g, w in zip(got, want):
                        assert np.array_equal(g, w[:n])


def test_checkpoint_bundle_and_frozen_graph_load_into_the_session(tmp_path, monkeypatch, suite):
    """SURVEY 8f-3 on hardware: weights written as a TF-V2 bundle (dual-net scopes, as
    instantiate_session_both_128 restores them, policy_net_utils.pyx:40-60) or as a frozen GraphDef,
    loaded through WANN_B200_CHECKPOINT / WANN_B200_FROZEN_GRAPH by the reference-shaped no-argument
    constructors, give bit-identical outputs to passing the same dicts directly."""
    import wann_b200
    from wann_b200 import tf_bundle, tf_graphdef, weights as W
    sq, bl = suite
    nodes_sq, nodes_bl = sq[:200], bl[:200]
    white, black = W.synthetic("trained_like", 6, final_kernel=1), W.synthetic("trained_like", 7, final_kernel=1)
    for k in ("WANN_B200_CHECKPOINT", "WANN_B200_FROZEN_GRAPH", "WANN_B200_SYNTHETIC"):
        monkeypatch.delenv(k, raising=False)
    with pytest.raises(W.MissingCheckpoint):
        wann_b200.NeuralNetsCombined_128()
    both = {"white_net/" + k: v for k, v in white.items()}
    both.update({"black_net/" + k: v for k, v in black.items()})
    prefix = str(tmp_path / "model")
    tf_bundle.save_policy_weights(prefix, both)
    with wann_b200.NeuralNetsCombined_128(white, black) as direct:
        want_w = direct._net(True).eval_probs(nodes_sq, nodes_bl)
        want_b = direct._net(False).eval_probs(nodes_sq, nodes_bl)
    assert not np.array_equal(want_w, want_b)
    monkeypatch.setenv("WANN_B200_CHECKPOINT", prefix)
    with wann_b200.NeuralNetsCombined_128() as loaded:
        assert np.array_equal(loaded._net(True).eval_probs(nodes_sq, nodes_bl), want_w)
        assert np.array_equal(loaded._net(False).eval_probs(nodes_sq, nodes_bl), want_b)
        planes = O.encode_planes(nodes_sq[:50], nodes_bl[:50])
        assert np.array_equal(loaded.evaluate(planes, already_converted=True), want_w[:50])
    # single 3x3 net: bundle and frozen .pb through instantiate_session()
    single = W.synthetic("trained_like", 8)
    with wann_b200.PolicyEvaluator(single, mode=wann_b200.session.DEFAULT_MODE,
                                   refine_margin=wann_b200.session.DEFAULT_REFINE_MARGIN) as ev:
        want = ev.eval_probs(nodes_sq, nodes_bl)
    tf_bundle.save_policy_weights(str(tmp_path / "one"), single)
    monkeypatch.setenv("WANN_B200_CHECKPOINT", str(tmp_path / "one"))
    sess, y_pred, X = wann_b200.instantiate_session()
    assert np.array_equal(sess.run(y_pred, feed_dict={X: O.encode_planes(nodes_sq, nodes_bl)}), want)
    sess.close()
    monkeypatch.delenv("WANN_B200_CHECKPOINT")
    pb = str(tmp_path / "policyNet_frozen.pb")
    tf_graphdef.save_frozen_graph(pb, single)
    monkeypatch.setenv("WANN_B200_FROZEN_GRAPH", pb)
    sess, y_pred, X = wann_b200.instantiate_session()
    assert np.array_equal(sess.run(y_pred, feed_dict={X: O.encode_planes(nodes_sq, nodes_bl)}), want)
    sess.close()


@pytest.mark.parametrize("mode,margin", [("fp16c", 0.02), ("bf16", 0.0)])
def test_device_resident_forest_equals_the_host_forest(mode, margin, weights, suite):
    """SURVEY 8f-2 moved onto the GPU: the device-resident forest (array_tree.hpp compiled for the device,
    one warp per tree, lock steps enqueued on one stream by wann_forest_run) grows node for node the trees
    of the host forest -- themselves pinned to the compiled reference's trees
    (test_native_search_forest_reproduces_the_references_trees) -- on the same evaluator, through
    searches, self-play moves with subtree reuse, and restarts of finished games."""
    import wann_b200
    sq, bl = suite
    n = 96
    with wann_b200.PolicyEvaluator(weights, mode=mode, refine_margin=margin) as ev:
        host = wann_b200.SearchForest(sq[:n], bl[:n])
        dev = wann_b200.SearchForest(sq[:n], bl[:n], evaluator=ev, device_resident=True, node_capacity=65536)
        host.native_run = False                                  # the Python next/eval/apply loop
        steps_h, evals_h = host.run(ev, 6)
        steps_d, evals_d = dev.run(ev, 6)
        assert evals_d == evals_h and evals_h > 6 * n and steps_d >= steps_h   # (a device tree yields its step after an evaluation-free simulation)
        for t in range(n):
            assert np.array_equal(dev.dump(t), host.dump(t)), t
            assert dev.info(t) == host.info(t)
        host.native_run = True                                   # wann_forest_run on a host forest
        for ply in range(4):                                     # self-play: move, reuse the subtree, search on
            mh, fh = host.play()
            md, fd = dev.play()
            assert np.array_equal(mh, md) and np.array_equal(fh, fd)
            assert host.run(ev, 3)[1] == dev.run(ev, 3)[1]
        for t in range(0, n, 7):
            assert np.array_equal(dev.dump(t), host.dump(t)), t
        # near-finished games: winning moves restart the tree from the opening
        late_sq, late_bl = O.random_positions_fast(4000, seed=5)
        pieces = (late_sq != 0).sum(1)
        pick = np.argsort(pieces)[:32]
        host2 = wann_b200.SearchForest(late_sq[pick], late_bl[pick])
        dev2 = wann_b200.SearchForest(late_sq[pick], late_bl[pick], evaluator=ev, device_resident=True, node_capacity=65536)
        finished = 0
        for ply in range(6):
            assert host2.run(ev, 4)[1] == dev2.run(ev, 4)[1]
            mh, fh = host2.play()
            md, fd = dev2.play()
            assert np.array_equal(mh, md) and np.array_equal(fh, fd)
            finished += int(fd.sum())
        for t in range(32):
            assert np.array_equal(dev2.dump(t), host2.dump(t)), t
        # a tree that fills its arena stops cleanly (finished == 2), the others carry on
        tiny = wann_b200.SearchForest(sq[:8], bl[:8], evaluator=ev, device_resident=True, node_capacity=256)
        tiny.run(ev, 50)
        infos = [tiny.info(t) for t in range(8)]
        assert all(i["nodes"] <= 256 for i in infos) and any(i["finished"] == 2 for i in infos)
        for f in (host, dev, host2, dev2, tiny):
            f.close()


def test_forest_step_budget_and_drain(weights, suite):
    """wann_forest_run_steps: a budget of lock steps (the reference's search has a time budget,
    expansion_MCTS_functions.pyx:113-119) may leave trees mid-simulation; draining with a simulation budget
    of 0 finishes them, after which moves can be played.  Host and device forests stay identical."""
    import wann_b200
    sq, bl = suite
    n = 48
    with wann_b200.PolicyEvaluator(weights, mode="fp16c") as ev:
        host = wann_b200.SearchForest(sq[:n], bl[:n])
        dev = wann_b200.SearchForest(sq[:n], bl[:n], evaluator=ev, device_resident=True, node_capacity=32768)
        for forest in (host, dev):
            steps, evals = forest.run_steps(ev, 25)
            assert steps == 25 and evals > 20 * n                 # (nearly) every tree asks every step
        with pytest.raises(wann_b200.WannError):                  # trees are mid-simulation
            dev.play()
        # the two drivers order evaluation-free simulations differently (a device tree yields its step), so the
        # forests are compared after both have completed the same number of simulations
        sims = max(host.info(t)["simulations"] for t in range(n)) + 2
        host.run(ev, sims)
        dev.run(ev, sims)
        for t in range(n):
            assert np.array_equal(dev.dump(t), host.dump(t)), t
        dev.run(ev, 0)
        mh, fh = host.play()
        md, fd = dev.play()
        assert np.array_equal(mh, md) and (md != 255).all()
        s, e = dev.run_for(ev, 0.05)                               # time budget + drain
        assert e > 0
        dev.play()
        host.close()
        dev.close()


def test_expand_ranked_equals_the_fused_call(ev_bf16, suite):
    """wann_expand_ranked: child construction and seeding on their own, from ranked children that are already
    on the device, give what wann_eval_expand_seeded gives in one call (and hence the oracle's, bit for bit)."""
    import torch
    sq, bl = suite
    n = 500
    dev = torch.device("cuda", ev_bf16.device)
    dsq, dbl = torch.from_numpy(sq[:n]).to(dev), torch.from_numpy(bl[:n]).to(dev)
    idx, pri, cnt, kids, flg, st, ps = ev_bf16.eval_expand_seeded(dsq, dbl)
    k2, f2, s2, p2 = ev_bf16.expand_ranked(dsq, dbl, idx, pri, cnt)
    torch.cuda.synchronize()
    for a, b in ((kids, k2), (flg, f2), (st, s2), (ps, p2)):
        assert torch.equal(a, b)
    k3, f3, _, _ = ev_bf16.expand_ranked(dsq, dbl, idx, pri, cnt, with_seeds=False)      # children only: 2 flag bits
    torch.cuda.synchronize()
    assert torch.equal(k3, kids) and torch.equal(f3 & 3, flg & 3)
    wk, wf = O.expand_children(sq[:n], bl[:n], idx.cpu().numpy(), cnt.cpu().numpy())
    assert np.array_equal(k3.cpu().numpy(), wk) and np.array_equal(f3.cpu().numpy(), wf)
