"""Packs the reference's Cogen plant model into the engine's layout and writes golden vectors.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/gen_cogen_pack.py

Reads ``sustaingym/data/cogen/onnx_model/model.onnx`` + ``model.json`` with the raw protobuf reader
in ``oracle/onnx_reader.py`` (onnx / onnxruntime are not installed), recognises the 23 sub-models
``(x - mu) * sinv -> Dense -> act -> Dense -> act -> Dense(1)`` and writes

  sustaingym_b200/data/cogen_plant.npz   weights in the kernel's layout (reference DATA, needed at run
                                         time on the GPU box where /root/reference does not exist)
  tests/golden/cogen_model_io.npz        4096 random inputs (the sampler of SURVEY.md Appendix E) with
                                         the graph's outputs evaluated by the NumPy interpreter in f32
                                         and f64, plus the SURVEY D.2 mid-point vector

Extended feature vector E[25] the sub-models index: E[0:18] = model inputs (model.json order),
E[18:21] = GTi_PAC_FFU * g(GTi_PWR), E[21:24] = GTi_EVC_FFU * g(GTi_PWR), E[24] = IPPROC_M * g(ST_PWR),
with g(v) = |sign(v)|.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import onnx_reader as R  # noqa: E402

REF = "/root/reference/sustaingym/data/cogen/onnx_model"


def main():
    g = R.load(os.path.join(REF, "model.onnx"))
    meta = json.load(open(os.path.join(REF, "model.json")))
    in_ids = [e["id"] for e in meta["inputs"]]
    out_ids = [e["id"] for e in meta["outputs"]]
    fidx = {f"X:{n}": i for i, n in enumerate(in_ids)}
    for i in (1, 2, 3):
        fidx[f"X:GT{i}_PAC_FFU:raw"] = fidx[f"X:GT{i}_PAC_FFU"]
        fidx[f"X:GT{i}_EVC_FFU:raw"] = fidx[f"X:GT{i}_EVC_FFU"]
        fidx[f"X:GT{i}_PAC_FFU"] = 18 + (i - 1)
        fidx[f"X:GT{i}_EVC_FFU"] = 21 + (i - 1)
    fidx["IPPROC_M:corr"] = 24
    init = g["init"]
    prod = {n["output"][0]: n for n in g["nodes"]}
    cons = {}
    for n in g["nodes"]:
        for i in n["input"]:
            cons.setdefault(i, []).append(n)

    subs = []
    for n in g["nodes"]:
        if n["op"] == "Concat" and n["output"][0].endswith(":X"):
            name = n["output"][0][:-2]
            feats = [fidx[i] for i in n["input"]]
            cur = n["output"][0]
            chain = []
            while True:
                nxt = [c for c in cons.get(cur, []) if c["output"][0].startswith(name + ":")
                       or (c["op"] == "Add" and c["output"][0].startswith("Y:") and len(chain) == 9)]
                if not nxt:
                    break
                nd = nxt[0]
                chain.append(nd)
                cur = nd["output"][0]
                if cur == name + ":Y" or cur.startswith("Y:"):
                    break
            ops = [c["op"] for c in chain]
            assert ops[:2] == ["Sub", "Mul"] and ops[2::3][:3] == ["MatMul"] * 3, (name, ops)
            mu = init[chain[0]["input"][1]].reshape(-1)
            sinv = init[chain[1]["input"][1]].reshape(-1)
            W1, b1 = init[chain[2]["input"][1]], init[chain[3]["input"][1]]
            act1 = chain[4]["op"]
            W2, b2 = init[chain[5]["input"][1]], init[chain[6]["input"][1]]
            act2 = chain[7]["op"]
            W3, b3 = init[chain[8]["input"][1]], init[chain[9]["input"][1]]
            assert act1 == act2 and act1 in ("Sigmoid", "Relu") and W3.shape[1] == 1
            subs.append(dict(name=name, feats=feats, mu=mu, sinv=sinv, W1=W1, b1=b1, W2=W2, b2=b2,
                             W3=W3[:, 0], b3=float(b3.reshape(-1)[0]), act=0 if act1 == "Sigmoid" else 1,
                             out=cur))
    assert len(subs) == 23
    # where each raw sub-model output lands in Y (after the optional gate) -------------------------
    slot_of = {}
    for s, sub in enumerate(subs):
        o = sub["out"]
        users = cons.get(o, [])
        gate = -1
        if users and users[0]["op"] == "Mul":                     # Y:NAME = raw * gate
            gname = users[0]["input"][1]
            gate = {"X:GT1_PWR:gt0": 0, "X:GT2_PWR:gt0": 1, "X:GT3_PWR:gt0": 2,
                    "X:GT1_PWR+HR1_HPIP_M_PROC:gt0": 3, "X:GT2_PWR+HR2_HPIP_M_PROC:gt0": 4,
                    "X:GT3_PWR+HR3_HPIP_M_PROC:gt0": 5}[gname]
            o = users[0]["output"][0]
        final = prod["Y"]["input"]
        sub["y_index"] = final.index(o)
        sub["gate"] = gate
        slot_of[sub["name"]] = s

    order = sorted(range(23), key=lambda s: (subs[s]["W2"].shape, subs[s]["act"], s))   # group equal shapes
    pack = {"n_sub": 23, "in_ids": np.array(in_ids), "out_ids": np.array(out_ids)}
    dmax = max(len(s["feats"]) for s in subs)
    feats = np.full((23, dmax), -1, np.int32)
    mu = np.zeros((23, dmax), np.float32)
    sinv = np.zeros((23, dmax), np.float32)
    meta_i = np.zeros((23, 6), np.int32)      # d, H1, H2, act, y_index, gate
    b3 = np.zeros(23, np.float32)
    blobs, offs = [], np.zeros((23, 5), np.int64)   # offsets of W1, b1, W2, b2, W3 in the f32 blob
    pos = 0
    names = []
    for j, s in enumerate(order):
        sub = subs[s]
        d, H1 = sub["W1"].shape
        H2 = sub["W2"].shape[1]
        feats[j, :d] = sub["feats"]; mu[j, :d] = sub["mu"]; sinv[j, :d] = sub["sinv"]
        meta_i[j] = [d, H1, H2, sub["act"], sub["y_index"], sub["gate"]]
        b3[j] = sub["b3"]
        names.append(sub["name"])
        for c, arr in enumerate([sub["W1"], sub["b1"], sub["W2"], sub["b2"], sub["W3"]]):
            a = np.ascontiguousarray(arr, np.float32).reshape(-1)
            offs[j, c] = pos
            blobs.append(a)
            pos += a.size
            pad = (-pos) % 4                      # keep every array 16-byte aligned
            if pad:
                blobs.append(np.zeros(pad, np.float32)); pos += pad
    pack.update(names=np.array(names), feats=feats, mu=mu, sinv=sinv, meta=meta_i, b3=b3, offsets=offs,
                blob=np.concatenate(blobs),
                in_min=np.array([e["min"] for e in meta["inputs"]]), in_max=np.array([e["max"] for e in meta["inputs"]]),
                out_min=np.array([e["min"] for e in meta["outputs"]], float),
                out_max=np.array([e["max"] for e in meta["outputs"]], float))
    dst = os.path.join(ROOT, "sustaingym_b200", "data", "cogen_plant.npz")
    np.savez_compressed(dst, **pack)
    nparam = sum(b.size for b in blobs)
    print(f"{dst}: 23 sub-models, {nparam} packed f32, {os.path.getsize(dst)} B")

    # golden I/O ------------------------------------------------------------------------------------
    rng = np.random.default_rng(0)
    lo, hi = pack["in_min"], pack["in_max"]
    X = rng.uniform(lo, hi, (4096, 18))
    for c in (3, 4, 6, 7, 9, 10):
        X[:, c] = rng.integers(0, 2, 4096)
    X[:, 17] = rng.integers(1, 13, 4096)
    # a few rows with units switched off (gates = 0), as the sign/abs gating allows
    X[:64, 5] = 0; X[32:96, 8] = 0; X[64:128, 11] = 0; X[100:140, 15] = 0; X[:16, 12] = 0; X[8:24, 13] = 0
    X = X.astype(np.float32)
    mid = np.array([[73.50001, 14.5, 0.4999989, 1, 0, 104.95397, 0, 1, 104.95189, 1, 1, 109.45038, 611.3647,
                     607.01306, 654.63226, 54.59582, -768.14154, 6]], np.float32)
    X = np.concatenate([mid, X])
    with np.errstate(over="ignore"):
        Y32 = R.run(g, X, np.float32)
        Y64 = R.run(g, X.astype(np.float64), np.float64)
    gpath = os.path.join(HERE, "cogen_model_io.npz")
    np.savez_compressed(gpath, X=X, Y32=Y32, Y64=Y64)
    print(f"{gpath}: X{X.shape} Y{Y32.shape}; f32-vs-f64 max rel {np.max(np.abs(Y32 - Y64) / np.maximum(np.abs(Y64), 1e-3)):.2e}")
    print("mid-point Y:", Y32[0][:6])


if __name__ == "__main__":
    main()
