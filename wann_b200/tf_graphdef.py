"""Frozen TensorFlow GraphDef I/O for the policy net -- the other on-disk format beside the path
(SURVEY 8f-3).  The reference ships `cppModels/policyNet.pb`, an UNFROZEN GraphDef of the inference
graph (12 VariableV2 nodes, no weights), and shows the freeze call commented out
(policy_net_utils.pyx:74-75).  This module

  * writes the same inference graph FROZEN: identical node names, ops, edges and attributes
    (Placeholder -> 5 x [Conv2D NHWC/SAME/strides 1 -> BiasAdd -> Relu] -> Reshape [-1,64] -> MatMul
    -> BiasAdd -> Softmax), with every VariableV2 replaced by a Const holding the weights -- what
    TensorFlow's convert_variables_to_constants produces -- so that TF tooling (or the reference's
    planned C++ loader) can consume the weights this library evaluates;
  * reads such a file back into the weights dict (`hidden_layer/{1..5}/{weights,bias}`,
    `output_layer/{weights,bias}`), e.g. to run a frozen model somebody else exported.

No TensorFlow dependency: the handful of protobuf messages involved (GraphDef, NodeDef, AttrValue,
TensorProto, TensorShapeProto) are encoded / decoded by hand with tf_bundle's wire helpers."""
import struct

import numpy as np

from .tf_bundle import _enc_varint, _parse_fields, _pb
from . import weights as W

DT_FLOAT, DT_INT32 = 1, 3
GRAPH_PRODUCER = 21            # TensorFlow 1.0.0 (model.meta meta_info_def / versions.producer)


# ------------------------------------------------------------------ encoding
def _shape(dims):
    return b"".join(_pb(2, 2, _pb(1, 0, d & 0xFFFFFFFFFFFFFFFF)) for d in dims)      # int64 size, -1 as two's complement


def _tensor(arr):
    a = np.ascontiguousarray(arr)
    dt = DT_FLOAT if a.dtype == np.float32 else DT_INT32
    return _pb(1, 0, dt) + _pb(2, 2, _shape(a.shape)) + _pb(4, 2, a.astype(a.dtype.newbyteorder("<")).tobytes())


def _attr(kind, v):
    if kind == "type":
        return _pb(6, 0, v)
    if kind == "s":
        return _pb(2, 2, v.encode())
    if kind == "b":
        return _pb(5, 0, 1 if v else 0)
    if kind == "shape":
        return _pb(7, 2, _shape(v))
    if kind == "ints":          # AttrValue.list { repeated int64 i = 3 [packed] }
        return _pb(1, 2, _pb(3, 2, b"".join(_enc_varint(i) for i in v)))
    if kind == "tensor":
        return _pb(8, 2, _tensor(v))
    raise ValueError(kind)


def _node(name, op, inputs=(), **attrs):
    body = _pb(1, 2, name.encode()) + _pb(2, 2, op.encode())
    for i in inputs:
        body += _pb(3, 2, i.encode())
    for key in sorted(attrs):                      # map<string, AttrValue>: entry {key = 1, value = 2}
        kind, v = attrs[key]
        body += _pb(5, 2, _pb(1, 2, key.encode()) + _pb(2, 2, _attr(kind, v)))
    return _pb(1, 2, body)                         # GraphDef.node = 1


def frozen_graph_bytes(weights, final_kernel=3):
    w = W.validate(weights, final_kernel)
    F = ("type", DT_FLOAT)
    out = _node("Placeholder", "Placeholder", dtype=F, shape=("shape", [-1, 8, 8, 4]))
    prev = "Placeholder"

    def variable(name):                            # VariableV2 -> Const (frozen), the /read Identity is kept
        return _node(name, "Const", dtype=F, value=("tensor", w[name])) + \
            _node(name + "/read", "Identity", [name], T=F)
    for i in range(1, 6):
        s = "hidden_layer/%d/" % i
        out += variable(s + "weights")
        out += _node(s + "Conv2D", "Conv2D", [prev, s + "weights/read"], T=F, data_format=("s", "NHWC"),
                     padding=("s", "SAME"), strides=("ints", [1, 1, 1, 1]), use_cudnn_on_gpu=("b", True))
        out += variable(s + "bias")
        out += _node(s + "BiasAdd", "BiasAdd", [s + "Conv2D", s + "bias/read"], T=F, data_format=("s", "NHWC"))
        out += _node(s + "Relu", "Relu", [s + "BiasAdd"], T=F)
        prev = s + "Relu"
    out += _node("Reshape/shape", "Const", dtype=("type", DT_INT32), value=("tensor", np.array([-1, 64], np.int32)))
    out += _node("Reshape", "Reshape", [prev, "Reshape/shape"], T=F, Tshape=("type", DT_INT32))
    out += variable("output_layer/weights")
    out += _node("output_layer/MatMul", "MatMul", ["Reshape", "output_layer/weights/read"], T=F,
                 transpose_a=("b", False), transpose_b=("b", False))
    out += variable("output_layer/bias")
    out += _node("output_layer/output", "BiasAdd", ["output_layer/MatMul", "output_layer/bias/read"], T=F,
                 data_format=("s", "NHWC"))
    out += _node("output_layer/Softmax", "Softmax", ["output_layer/output"], T=F)
    return out + _pb(4, 2, _pb(1, 0, GRAPH_PRODUCER))          # GraphDef.versions { producer }


def save_frozen_graph(path, weights, final_kernel=3):
    with open(path, "wb") as f:
        f.write(frozen_graph_bytes(weights, final_kernel))


# ------------------------------------------------------------------ decoding
def _dec_shape(buf):
    dims = []
    for f, _, v in _parse_fields(buf):
        if f == 2:
            size = 0
            for f2, _, v2 in _parse_fields(v):
                if f2 == 1:
                    size = v2 - (1 << 64) if v2 >= (1 << 63) else v2
            dims.append(size)
    return dims


def _dec_tensor(buf):
    dt, shape, content, fvals, ivals = DT_FLOAT, [], b"", [], []
    for f, wt, v in _parse_fields(buf):
        if f == 1:
            dt = v
        elif f == 2:
            shape = _dec_shape(v)
        elif f == 4:
            content = v
        elif f == 5:                               # float_val: packed or single fixed32
            fvals += list(np.frombuffer(v, "<f4")) if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
        elif f == 7:                               # int_val: packed or single varint
            if wt == 2:
                pos = 0
                while pos < len(v):
                    x, pos = _parse_varint(v, pos)
                    ivals.append(x)
            else:
                ivals.append(v)
    np_dt = "<f4" if dt == DT_FLOAT else "<i4"
    n = int(np.prod(shape)) if shape else 1
    if content:
        arr = np.frombuffer(content, np_dt)
    else:
        vals = fvals if dt == DT_FLOAT else [x - (1 << 64) if x >= (1 << 63) else x for x in ivals]
        arr = np.array(vals, np_dt) if len(vals) == n else np.full(n, vals[0] if vals else 0, np_dt)
    return arr.reshape(shape).astype(np.float32 if dt == DT_FLOAT else np.int32)


def _parse_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def parse_graph(data):
    """GraphDef bytes -> list of {name, op, inputs, attr} (attr values: str / bool / int / list / ndarray)."""
    nodes = []
    for f, _, v in _parse_fields(data):
        if f != 1:
            continue
        nd = {"name": "", "op": "", "inputs": [], "attr": {}}
        for f2, _, v2 in _parse_fields(v):
            if f2 == 1:
                nd["name"] = v2.decode()
            elif f2 == 2:
                nd["op"] = v2.decode()
            elif f2 == 3:
                nd["inputs"].append(v2.decode())
            elif f2 == 5:
                key, val = "", None
                for f3, _, v3 in _parse_fields(v2):
                    if f3 == 1:
                        key = v3.decode()
                    elif f3 == 2:
                        for f4, wt4, v4 in _parse_fields(v3):
                            if f4 == 2:
                                val = v4.decode()
                            elif f4 == 5:
                                val = bool(v4)
                            elif f4 == 3:
                                val = int(v4)
                            elif f4 == 6:
                                val = {DT_FLOAT: "float32", DT_INT32: "int32"}.get(v4, v4)
                            elif f4 == 7:
                                val = _dec_shape(v4)
                            elif f4 == 8:
                                val = _dec_tensor(v4)
                            elif f4 == 1:          # list: only packed / repeated int64 are used by this graph
                                ints = []
                                for f5, wt5, v5 in _parse_fields(v4):
                                    if f5 == 3 and wt5 == 2:
                                        pos = 0
                                        while pos < len(v5):
                                            x, pos = _parse_varint(v5, pos)
                                            ints.append(x)
                                    elif f5 == 3:
                                        ints.append(v5)
                                val = ints
                nd["attr"][key] = val
        nodes.append(nd)
    return nodes


def load_frozen_weights(path, final_kernel=None):
    """Weights dict from a frozen GraphDef of this architecture (Const nodes named like the
    reference's variables).  final_kernel is inferred from hidden_layer/5/weights when None."""
    with open(path, "rb") as f:
        nodes = parse_graph(f.read())
    consts = {n["name"]: n["attr"].get("value") for n in nodes if n["op"] == "Const"}
    out = {}
    for name in W.shapes(3):
        if name not in consts or consts[name] is None:
            raise KeyError("frozen graph has no constant %r (is it frozen?)" % name)
        out[name] = consts[name]
    fk = int(out["hidden_layer/5/weights"].shape[0]) if final_kernel is None else final_kernel
    return W.validate(out, fk)
