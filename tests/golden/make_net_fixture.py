#!/usr/bin/env python
"""Golden OUTPUTS of the reference's own network, evaluated by hand, operator by operator.

The reference evaluates positions with Keras / TFLite (python/main.pyx:70-83,
rating/tourney.pyx:139-155); neither runtime exists in this image. Its shipped TFLite graphs do:
this script executes one of them (corintho_ai/docker/tflite_model.tflite -- the network behind
the public web app) exactly as stored, one operator at a time in the stored order, with the
float32 semantics of TFLite's reference kernels (tensorflow/lite/kernels/internal/reference):

  FULLY_CONNECTED  out[b,o] = act(bias[o] + sum_d in[b,d] * w[o,d]), float accumulator, d ascending
  RELU / MUL / ADD elementwise float32 (broadcast of a constant vector)
  TANH             std::tanh on float32
  SOFTMAX          exp(x - max) / sum, float32

Nothing here shares code with the weight importer's folding algebra
(corintho_ai_b200/tflite_import.py:load_tflite_weights) or with the engine's kernels: the graph is
walked generically from its operator list. Output: tests/golden/net_fixture.npz with
  positions [N,70] (uint8, quarter units: x = q / 4), value [N] f32, policy [N,96] f32 (float32
  execution as above), value64 / policy64 (the same graph in float64, for error budgets), and the
  graph's output order (tourney.pyx:153-154: TFLite returns policy first, value second).
Run in the build container:  python tests/golden/make_net_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from corintho_ai_b200.tflite_import import (ACT_NONE, ACT_RELU, ACT_TANH, OP_ADD, OP_FULLY_CONNECTED,  # noqa: E402
                                            OP_MUL, OP_RELU, OP_SOFTMAX, OP_TANH, parse_tflite)

SRC = "/root/reference/corintho_ai/docker/tflite_model.tflite"
N = 512


def positions(n, seed=2024):
    """Reachable positions from seeded random legal play (oracle rules), as 70-float rows."""
    from oracle.pyoracle import OracleLib
    O = OracleLib()
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        st = O.start()
        for _ in range(40):
            rows.append(O.encode(st))
            mask, _ = O.legal(st)
            ids = [m for m in range(96) if (mask[m >> 5] >> (m & 31)) & 1]
            if not ids:
                break
            st = O.do_move(st, int(rng.choice(ids)))
    return np.stack(rows[:n]).astype(np.float32)


def activation(x, act, dtype):
    if act == ACT_NONE:
        return x
    if act == ACT_RELU:
        return np.maximum(x, dtype(0))
    if act == ACT_TANH:
        return np.tanh(x).astype(dtype)
    raise ValueError(f"fused activation {act}")


def run_graph(path, x, dtype):
    """Execute subgraph 0 on a batch x [n,70]; returns {output tensor index: array}."""
    tensors, ops, inputs, outputs = parse_tflite(path)
    val = {inputs[0]: np.asarray(x, dtype)}

    def get(i):
        if i in val:
            return val[i]
        d = tensors[i]["data"]
        if d is None:
            raise ValueError(f"tensor {i} has no value")
        return d.astype(dtype)

    for op in ops:
        code, ins, outs = op["op"], op["inputs"], op["outputs"]
        if code == OP_FULLY_CONNECTED:
            a, w = get(ins[0]), get(ins[1])                      # [n,in], [out,in]
            acc = np.zeros((a.shape[0], w.shape[0]), dtype)
            for d in range(w.shape[1]):                           # depth ascending, one rounding per step
                acc = (acc + a[:, d:d + 1] * w[None, :, d]).astype(dtype)
            if len(ins) > 2 and ins[2] >= 0:
                acc = (acc + get(ins[2])[None, :]).astype(dtype)
            r = activation(acc, op["act"], dtype)
        elif code in (OP_MUL, OP_ADD):
            a, b = get(ins[0]), get(ins[1])
            r = (a * b if code == OP_MUL else a + b).astype(dtype)
            r = activation(r, op["act"], dtype)
        elif code == OP_RELU:
            r = np.maximum(get(ins[0]), dtype(0))
        elif code == OP_TANH:
            r = np.tanh(get(ins[0])).astype(dtype)
        elif code == OP_SOFTMAX:
            a = get(ins[0])
            e = np.exp((a - a.max(-1, keepdims=True)).astype(dtype)).astype(dtype)
            r = (e / e.sum(-1, keepdims=True).astype(dtype)).astype(dtype)
        else:
            raise ValueError(f"operator {code} is not part of the reference's exported graphs")
        val[outs[0]] = r
    return [val[o] for o in outputs], [tensors[o]["shape"] for o in outputs]


if __name__ == "__main__":
    x = positions(N)
    o32, shapes = run_graph(SRC, x, np.float32)
    o64, _ = run_graph(SRC, x, np.float64)
    # which output is which: the policy has 96 columns, the value 1 (tourney.pyx:153-154 reads
    # output 0 as probabilities and output 1 as evaluations)
    order = ["policy" if o.shape[-1] == 96 else "value" for o in o32]
    named32 = dict(zip(order, o32))
    named64 = dict(zip(order, o64))
    q = np.rint(x * 4).astype(np.uint8)
    assert np.array_equal(q.astype(np.float32) / 4, x)
    np.savez_compressed(os.path.join(HERE, "net_fixture.npz"), positions=q,
                        value=named32["value"].reshape(-1).astype(np.float32),
                        policy=named32["policy"].astype(np.float32),
                        value64=named64["value"].reshape(-1), policy64=named64["policy"],
                        output_order=np.array(order), source=np.bytes_(SRC))
    print("outputs", order, "shapes", shapes, "| f32 vs f64: value %.2e policy %.2e"
          % (np.abs(named32["value"].reshape(-1) - named64["value"].reshape(-1)).max(),
             np.abs(named32["policy"] - named64["policy"]).max()))
