"""Weights for the policy net: names and shapes of policy_net_model/model.index
(hidden_layer/{1..5}/{weights,bias}, output_layer/{weights,bias}; HWIO conv kernels).

The reference's weight shard (model.data-00000-of-00001) is absent from the reference tree
(.MISSING_LARGE_BLOBS:1), so besides the TF-V2 bundle reader (tf_bundle.py) this module
provides deterministic synthetic presets used by the tests and the benchmark."""
import os

import numpy as np

NAMES = ["hidden_layer/%d/%s" % (i, k) for i in range(1, 6) for k in ("weights", "bias")] + \
        ["output_layer/weights", "output_layer/bias"]


def shapes(final_kernel=3):
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


def synthetic(preset="trained_like", seed=1, final_kernel=3):
    """'ref_init': the reference initialisers (policy_net_utils.pyx:152,161,166-167,227,233).
    'trained_like': fan-in He conv weights, small biases, FC scaled to logit sigma ~ 2.5."""
    rng = np.random.RandomState(seed)
    w = {}
    shp = shapes(final_kernel)
    for i in range(1, 6):
        s = shp["hidden_layer/%d/weights" % i]
        k, _, cin, cout = s
        if preset == "ref_init":
            std = np.sqrt(2.0 / (64 * cin))
            b = np.full((cout,), 0.01, np.float32)
        else:
            std = np.sqrt(2.0 / (k * k * cin))
            b = (0.05 * rng.standard_normal(cout)).astype(np.float32)
        w["hidden_layer/%d/weights" % i] = (std * rng.standard_normal(s)).astype(np.float32)
        w["hidden_layer/%d/bias" % i] = b
    if preset == "ref_init":
        lim = np.sqrt(6.0 / (64 + 155))
        w["output_layer/weights"] = rng.uniform(-lim, lim, (64, 155)).astype(np.float32)
        w["output_layer/bias"] = np.zeros(155, np.float32)
    else:
        w["output_layer/weights"] = (0.9 * rng.standard_normal((64, 155))).astype(np.float32)
        w["output_layer/bias"] = (0.1 * rng.standard_normal(155)).astype(np.float32)
    return w


def validate(w, final_kernel=3):
    out = {}
    for name, shape in shapes(final_kernel).items():
        if name not in w:
            raise KeyError("missing weight %r" % name)
        a = np.ascontiguousarray(w[name], dtype=np.float32)
        if a.shape != shape:
            raise ValueError("weight %r has shape %s, expected %s" % (name, a.shape, shape))
        out[name] = a
    return out


class MissingCheckpoint(FileNotFoundError):
    """No weights were given and no checkpoint is configured."""


def default_weights(scope="", final_kernel=3):
    """Weights used when the reference-shaped constructors are called with no arguments
    (the reference hard-codes its checkpoint path, policy_net_utils.pyx:69, and fails when the
    checkpoint is missing).  Order: $WANN_B200_CHECKPOINT (TF-V2 bundle prefix, CRC-verified; `scope`
    = 'white_net/' / 'black_net/' for the dual-net bundles of policy_net_utils.pyx:22-25, falling
    back to the unscoped names when the bundle holds a single net), else $WANN_B200_FROZEN_GRAPH
    (frozen GraphDef .pb, tf_graphdef.py).  Synthetic weights are used ONLY on explicit request
    ($WANN_B200_SYNTHETIC=1: the 'trained_like' preset, seed from $WANN_B200_SEED, default 1);
    otherwise a missing checkpoint raises MissingCheckpoint -- a drop-in that silently played on
    random weights would be a trap."""
    prefix = os.environ.get("WANN_B200_CHECKPOINT")
    if prefix:
        from . import tf_bundle
        if scope:
            names = tf_bundle.read_index(prefix + ".index")
            if not any(k.startswith(scope) for k in names):
                scope = ""                        # a single-net bundle serves both colours
        return tf_bundle.load_policy_weights(prefix, scope=scope, final_kernel=final_kernel)
    frozen = os.environ.get("WANN_B200_FROZEN_GRAPH")
    if frozen:                                    # a frozen GraphDef (.pb) of the same architecture
        from . import tf_graphdef
        return tf_graphdef.load_frozen_weights(frozen, final_kernel=final_kernel)
    if os.environ.get("WANN_B200_SYNTHETIC") == "1":
        return synthetic("trained_like", int(os.environ.get("WANN_B200_SEED", "1")), final_kernel)
    raise MissingCheckpoint(
        "no policy-net weights: pass weights=..., or set WANN_B200_CHECKPOINT (TF-V2 bundle prefix, e.g. "
        ".../policy_net_model/model) or WANN_B200_FROZEN_GRAPH (.pb).  The reference repository does not ship "
        "its weight shard (policy_net_model/model.data-00000-of-00001); WANN_B200_SYNTHETIC=1 asks for the "
        "synthetic 'trained_like' preset explicitly (tests, benchmarks).")
