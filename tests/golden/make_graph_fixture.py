"""Extract the reference's OWN inference graph -- the nodes `output_layer/Softmax` depends on -- from
its two serialized TensorFlow graphs (/root/reference/cppModels/policyNet.pb, a GraphDef, and
/root/reference/policy_net_model/model.meta, the checkpoint's MetaGraphDef) into
tests/golden/ref_graph.json, so that oracle/graph_interp.py can execute the reference's graph
topology and attributes (not our reading of policy_net_utils.pyx) on the GPU box, where
/root/reference does not exist.  Run here:  python tests/golden/make_graph_fixture.py

Only node names, op types, input edges, attributes and small Const tensors are extracted; the
protobuf schema comes from the `tensorboard` package in this image (TensorFlow itself is absent)."""
import json
import os

import numpy as np
from tensorboard.compat.proto import graph_pb2, meta_graph_pb2, types_pb2

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_graph.json")
DT = {types_pb2.DT_FLOAT: "float32", types_pb2.DT_INT32: "int32"}


def attr_value(v):
    kind = v.WhichOneof("value")
    if kind == "s":
        return v.s.decode()
    if kind == "b":
        return bool(v.b)
    if kind == "i":
        return int(v.i)
    if kind == "type":
        return DT.get(v.type, int(v.type))
    if kind == "shape":
        return [int(d.size) for d in v.shape.dim]
    if kind == "list":
        return [int(x) for x in v.list.i]
    if kind == "tensor":
        t = v.tensor
        dt = DT[t.dtype]
        shape = [int(d.size) for d in t.tensor_shape.dim]
        if t.tensor_content:
            arr = np.frombuffer(t.tensor_content, dtype=dt).reshape(shape)
        else:
            vals = list(t.int_val) if dt == "int32" else list(t.float_val)
            arr = np.array(vals, dtype=dt).reshape(shape) if len(vals) == int(np.prod(shape)) else \
                np.full(shape, vals[0], dtype=dt)
        return {"dtype": dt, "shape": shape, "values": arr.reshape(-1).tolist()}
    return None


def inference_subgraph(graph_def):
    """Nodes the graph's one Softmax op depends on, in execution order; the fetch is last.
    (policyNet.pb names it output_layer/Softmax -- policy_net_utils.pyx:217,235; the training
    graph in model.meta names it Softmax.)"""
    by_name = {n.name: n for n in graph_def.node}
    fetches = [n.name for n in graph_def.node if n.op == "Softmax"]
    assert len(fetches) == 1, fetches
    fetch = fetches[0]
    order, seen = [], set()

    def visit(name):
        name = name.split(":")[0]
        if name in seen or name.startswith("^"):
            return
        seen.add(name)
        node = by_name[name]
        if node.op != "VariableV2":                 # a variable's initializer is not on the inference path
            for i in node.input:
                visit(i)
        order.append(node)

    visit(fetch)
    out = []
    for n in order:
        out.append({"name": n.name, "op": n.op,
                    "inputs": [] if n.op == "VariableV2" else [i.split(":")[0] for i in n.input],
                    "attr": {k: attr_value(v) for k, v in sorted(n.attr.items()) if not k.startswith("_")}})
    return out


def main():
    g = graph_pb2.GraphDef()
    g.ParseFromString(open(os.path.join(REF, "cppModels", "policyNet.pb"), "rb").read())
    m = meta_graph_pb2.MetaGraphDef()
    m.ParseFromString(open(os.path.join(REF, "policy_net_model", "model.meta"), "rb").read())
    doc = {"made_by": "tests/golden/make_graph_fixture.py",
           "tensorflow_version": m.meta_info_def.tensorflow_version,
           "policyNet_pb": inference_subgraph(g),
           "model_meta": inference_subgraph(m.graph_def)}
    with open(OUT, "w") as f:
        json.dump(doc, f, indent=0, sort_keys=True)
    print(OUT, len(doc["policyNet_pb"]), len(doc["model_meta"]), os.path.getsize(OUT))


if __name__ == "__main__":
    main()
