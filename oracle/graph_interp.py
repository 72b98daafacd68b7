"""TEST INFRASTRUCTURE (oracle): a small interpreter for the reference's own serialized inference
graph (tests/golden/ref_graph.json, extracted from /root/reference/cppModels/policyNet.pb and
policy_net_model/model.meta by tests/golden/make_graph_fixture.py).

oracle.forward_fp32 restates policy_net_utils.pyx:106-124,149-178,215-239 by reading the source;
this module instead EXECUTES the node list TensorFlow serialized for that code -- op types, edges,
`strides`, `padding`, `data_format`, `transpose_a/b`, the Reshape constant, the variable shapes --
so the topology and attributes are the reference's artefact, not our reading of it.  What remains
restated are the op kernels themselves, following TensorFlow 1.0's published definitions:
  Conv2D (NHWC, filter HWIO): out[b,i,j,k] = sum_{di,dj,q} in[b, s1*i+di-pt, s2*j+dj-pl, q] * f[di,dj,q,k]
          SAME: out = ceil(in/stride), pad_total = max((out-1)*stride + k - in, 0), pad_before = pad_total//2
  BiasAdd: broadcast over the last (NHWC) dimension;  Relu: max(x, 0);  Reshape: row-major;
  MatMul: a @ b with optional transposes;  Softmax: exp(x - max) / sum over the last dimension.
Logits/probabilities remain PARITY UNPINNED (no TensorFlow, no trained weights; see oracle.py)."""
import json
import os

import numpy as np

FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                       "ref_graph.json")


def load(which="policyNet_pb", path=FIXTURE):
    with open(path) as f:
        return json.load(f)[which]


def _conv2d(x, f, attr):
    assert attr["data_format"] == "NHWC" and attr["padding"] in ("SAME", "VALID")
    s = attr["strides"]
    assert s[0] == 1 and s[3] == 1
    n, h, w, cin = x.shape
    kh, kw, fin, cout = f.shape
    assert fin == cin
    if attr["padding"] == "SAME":
        oh, ow = -(-h // s[1]), -(-w // s[2])
        ph, pw = max((oh - 1) * s[1] + kh - h, 0), max((ow - 1) * s[2] + kw - w, 0)
    else:
        oh, ow, ph, pw = (h - kh) // s[1] + 1, (w - kw) // s[2] + 1, 0, 0
    xp = np.zeros((n, h + ph, w + pw, cin), x.dtype)
    xp[:, ph // 2:ph // 2 + h, pw // 2:pw // 2 + w] = x
    out = np.zeros((n, oh, ow, cout), np.float64)
    for di in range(kh):
        for dj in range(kw):
            patch = xp[:, di:di + (oh - 1) * s[1] + 1:s[1], dj:dj + (ow - 1) * s[2] + 1:s[2]]
            out += patch.reshape(-1, cin).astype(np.float64).dot(f[di, dj].astype(np.float64)).reshape(n, oh, ow, cout)
    return out.astype(x.dtype)


def run(nodes, feed, variables, fetch=None, return_all=False):
    """Execute `nodes` (already in dependency order).  feed: {placeholder name: array};
    variables: {VariableV2 name: array} (shapes are checked against the graph's).  Returns the value
    of the last node (the Softmax), or of `fetch`, or every value."""
    val = {}
    for nd in nodes:
        op, a = nd["op"], nd["attr"]
        ins = [val[i] for i in nd["inputs"]]
        if op == "Placeholder":
            v = np.asarray(feed[nd["name"]], dtype=a["dtype"])
        elif op == "VariableV2":
            v = np.asarray(variables[nd["name"]], dtype=a["dtype"])
            assert list(v.shape) == a["shape"], (nd["name"], v.shape, a["shape"])
        elif op == "Const":
            t = a["value"]
            v = np.array(t["values"], dtype=t["dtype"]).reshape(t["shape"])
        elif op == "Identity":
            v = ins[0]
        elif op == "Conv2D":
            v = _conv2d(ins[0], ins[1], a)
        elif op == "BiasAdd":
            assert a["data_format"] == "NHWC"
            v = ins[0] + ins[1]
        elif op == "Relu":
            v = np.maximum(ins[0], 0)
        elif op == "Reshape":
            v = ins[0].reshape([int(d) for d in ins[1]])
        elif op == "MatMul":
            x, y = (ins[0].T if a["transpose_a"] else ins[0]), (ins[1].T if a["transpose_b"] else ins[1])
            v = x.astype(np.float64).dot(y.astype(np.float64)).astype(x.dtype)
        elif op == "Softmax":
            e = np.exp(ins[0] - ins[0].max(axis=-1, keepdims=True))
            v = e / e.sum(axis=-1, keepdims=True)
        else:
            raise NotImplementedError("op %s (%s) is not on the reference's inference path" % (op, nd["name"]))
        val[nd["name"]] = v
    if return_all:
        return val
    return val[fetch or nodes[-1]["name"]]
