"""TEST INFRASTRUCTURE — minimal ONNX protobuf reader + NumPy evaluator.

``onnx`` / ``onnxruntime`` are not installed in this image.  The reference's plant
model ``sustaingym/data/cogen/onnx_model/model.onnx`` (ir 8, opset 12, producer sclblonnx) only uses
11 op types (Constant, Gather, Sign, Abs, Add, Sub, Mul, Concat, MatMul, Sigmoid, Relu), so a
hand-rolled wire-format walker using the public ONNX field numbers is enough to decode it, and a
dict-of-arrays interpreter evaluates it in file order (the graph is topologically sorted).

The reference evaluates this graph with ONNX Runtime at batch 1 (``cogen/env.py:305``); ORT's CPU
kernels (MLAS sgemm) sum in a different order, so agreement is ~1e-6 relative, not bit-exact.
"""
from __future__ import annotations

import struct

import numpy as np


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yields (field_number, wire_type, value) for one message; length-delimited values are bytes."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fn, wt, v


_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 11: np.float64}


def _tensor(buf):
    dims, dtype, name, raw, floats, int64s = [], 1, "", None, [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            if wt == 2:      # packed
                p = 0
                while p < len(v):
                    d, p = _varint(v, p); dims.append(d)
            else:
                dims.append(v)
        elif fn == 2:
            dtype = v
        elif fn == 4:
            floats += list(struct.unpack(f"<{len(v) // 4}f", v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fn == 7:
            if wt == 2:
                p = 0
                while p < len(v):
                    d, p = _varint(v, p); int64s.append(d if d < 1 << 63 else d - (1 << 64))
            else:
                int64s.append(v)
        elif fn == 8:
            name = v.decode()
        elif fn == 9:
            raw = v
    dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=dt).copy()
    elif floats:
        arr = np.array(floats, dtype=dt)
    else:
        arr = np.array(int64s, dtype=dt)
    return name, arr.reshape(dims) if dims else arr.reshape(())


def _attribute(buf):
    name, val = "", None
    for fn, wt, v in _fields(buf):
        if fn == 1:
            name = v.decode()
        elif fn == 3:
            val = v if v < 1 << 63 else v - (1 << 64)
        elif fn == 5:
            val = _tensor(v)[1]
        elif fn == 2:
            val = struct.unpack("<f", v)[0]
    return name, val


def _node(buf):
    node = {"input": [], "output": [], "name": "", "op": "", "attr": {}}
    for fn, wt, v in _fields(buf):
        if fn == 1:
            node["input"].append(v.decode())
        elif fn == 2:
            node["output"].append(v.decode())
        elif fn == 3:
            node["name"] = v.decode()
        elif fn == 4:
            node["op"] = v.decode()
        elif fn == 5:
            k, a = _attribute(v)
            node["attr"][k] = a
    return node


def _value_info_name(buf):
    for fn, wt, v in _fields(buf):
        if fn == 1:
            return v.decode()
    return ""


def load(path: str) -> dict:
    """Returns {'nodes': [...], 'init': {name: array}, 'inputs': [...], 'outputs': [...]}."""
    buf = open(path, "rb").read()
    graph = None
    for fn, wt, v in _fields(buf):
        if fn == 7:
            graph = v
    g = {"nodes": [], "init": {}, "inputs": [], "outputs": []}
    for fn, wt, v in _fields(graph):
        if fn == 1:
            g["nodes"].append(_node(v))
        elif fn == 5:
            name, arr = _tensor(v)
            g["init"][name] = arr
        elif fn == 11:
            g["inputs"].append(_value_info_name(v))
        elif fn == 12:
            g["outputs"].append(_value_info_name(v))
    g["inputs"] = [i for i in g["inputs"] if i not in g["init"]]
    return g


def run(g: dict, X: np.ndarray, dtype=np.float32) -> np.ndarray:
    """Evaluates the graph on X [B, 18] in the given float dtype."""
    env = {k: (v.astype(dtype) if v.dtype.kind == "f" else v) for k, v in g["init"].items()}
    env[g["inputs"][0]] = np.asarray(X, dtype=dtype)
    for nd in g["nodes"]:
        i = [env[k] for k in nd["input"]]
        op = nd["op"]
        if op == "Constant":
            v = nd["attr"]["value"]
            o = v.astype(dtype) if v.dtype.kind == "f" else v
        elif op == "Gather":
            o = np.take(i[0], i[1], axis=nd["attr"].get("axis", 0))
        elif op == "Sign":
            o = np.sign(i[0])
        elif op == "Abs":
            o = np.abs(i[0])
        elif op == "Add":
            o = i[0] + i[1]
        elif op == "Sub":
            o = i[0] - i[1]
        elif op == "Mul":
            o = i[0] * i[1]
        elif op == "Concat":
            o = np.concatenate(i, axis=nd["attr"]["axis"])
        elif op == "MatMul":
            o = i[0] @ i[1]
        elif op == "Sigmoid":
            o = (1 / (1 + np.exp(-i[0]))).astype(dtype)
        elif op == "Relu":
            o = np.maximum(i[0], 0)
        else:
            raise NotImplementedError(op)
        env[nd["output"][0]] = o
    return env[g["outputs"][0]]
