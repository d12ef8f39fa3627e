"""Weight import from the reference's shipped TFLite checkpoints.

The reference exports every generation's network with the TFLite converter
(corintho_ai/rating/tflite_models/*.tflite, corintho_ai/docker/tflite_model.tflite) and runs
them with tflite_runtime (rating/tourney.pyx:139-155). Neither TensorFlow nor the flatbuffers
package is available here, so this module reads the FlatBuffer container directly (format:
little-endian, root uoffset at byte 0, tables with vtables; schema ids below are those of
tensorflow/lite/schema/schema.fbs) and rebuilds the network of wrapper.py:256-271 as the flat
fp32 weight vector the engine consumes (see corintho_ai_b200.fold_batchnorm for the layout).

Only the operators such an export contains are understood: FULLY_CONNECTED (with optional fused
RELU), MUL / ADD by a constant vector (the inference-time BatchNormalization that FOLLOWS the
ReLU), TANH / LOGISTIC-free heads and SOFTMAX. Anything else raises ValueError.
"""
import struct

import numpy as np

# BuiltinOperator codes (schema.fbs)
OP_ADD, OP_FULLY_CONNECTED, OP_MUL, OP_RELU, OP_SOFTMAX, OP_TANH = 0, 9, 18, 19, 25, 28
# ActivationFunctionType
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 4


class _FB:
    """Minimal FlatBuffer reader."""

    def __init__(self, buf):
        self.b = memoryview(buf)

    def u8(self, o):
        return self.b[o]

    def i8(self, o):
        return struct.unpack_from("<b", self.b, o)[0]

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def i32(self, o):
        return struct.unpack_from("<i", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def root(self):
        return self.u32(0)

    def field(self, table, idx):
        """Absolute offset of field `idx` of the table at `table`, or None if absent."""
        vt = table - self.i32(table)
        vt_size = self.u16(vt)
        slot = 4 + 2 * idx
        if slot >= vt_size:
            return None
        off = self.u16(vt + slot)
        return table + off if off else None

    def indirect(self, o):
        return o + self.u32(o)

    def vector(self, table, idx):
        """(start offset of elements, length) of a vector field, or (None, 0)."""
        f = self.field(table, idx)
        if f is None:
            return None, 0
        v = self.indirect(f)
        return v + 4, self.u32(v)

    def table_vector(self, table, idx):
        start, n = self.vector(table, idx)
        return [self.indirect(start + 4 * i) for i in range(n)]

    def i32_vector(self, table, idx):
        start, n = self.vector(table, idx)
        return [self.i32(start + 4 * i) for i in range(n)]

    def string(self, table, idx):
        f = self.field(table, idx)
        if f is None:
            return ""
        s = self.indirect(f)
        n = self.u32(s)
        return bytes(self.b[s + 4:s + 4 + n]).decode("utf-8", "replace")

    def scalar(self, table, idx, kind, default=0):
        f = self.field(table, idx)
        if f is None:
            return default
        return {"i8": self.i8, "u8": self.u8, "i32": self.i32, "u32": self.u32}[kind](f)


def parse_tflite(path):
    """Returns (tensors, operators, inputs, outputs) of subgraph 0:
    tensors[i] = {"name", "shape", "type", "data" (np.float32 array or None)};
    operators = [{"op": builtin code, "inputs": [...], "outputs": [...], "act": fused activation}]."""
    buf = open(path, "rb").read()
    if buf[4:8] != b"TFL3":
        raise ValueError(f"{path}: not a TFLite flatbuffer")
    fb = _FB(buf)
    model = fb.root()
    opcodes = []
    for oc in fb.table_vector(model, 1):  # Model.operator_codes
        dep = fb.scalar(oc, 0, "i8")       # deprecated_builtin_code (int8)
        new = fb.scalar(oc, 3, "i32")      # builtin_code
        opcodes.append(max(dep, new))
    buffers = []
    for bt in fb.table_vector(model, 4):   # Model.buffers
        start, n = fb.vector(bt, 0)        # Buffer.data
        buffers.append(bytes(fb.b[start:start + n]) if n else b"")
    sg = fb.table_vector(model, 2)[0]      # Model.subgraphs[0]
    tensors = []
    for tt in fb.table_vector(sg, 0):      # SubGraph.tensors
        shape = fb.i32_vector(tt, 0)
        ttype = fb.scalar(tt, 1, "i8")     # TensorType: 0 = FLOAT32
        bidx = fb.scalar(tt, 2, "u32")
        data = None
        if bidx < len(buffers) and buffers[bidx] and ttype == 0:
            data = np.frombuffer(buffers[bidx], np.float32).copy()
            if shape:
                data = data.reshape(shape)
        tensors.append({"name": fb.string(tt, 3), "shape": shape, "type": ttype, "data": data})
    ops = []
    for ot in fb.table_vector(sg, 3):      # SubGraph.operators
        code = opcodes[fb.scalar(ot, 0, "u32")]
        act = ACT_NONE
        f = fb.field(ot, 4)                # builtin_options (union table)
        if f is not None and code in (OP_FULLY_CONNECTED, OP_ADD, OP_MUL):
            act = fb.scalar(fb.indirect(f), 0, "i8")  # fused_activation_function is field 0
        ops.append({"op": code, "inputs": fb.i32_vector(ot, 1), "outputs": fb.i32_vector(ot, 2), "act": act})
    return tensors, ops, fb.i32_vector(sg, 1), fb.i32_vector(sg, 2)


def load_tflite_weights(path):
    """Fold a reference TFLite checkpoint into the engine's flat weight vector (127 997 floats).

    Walks the operator chain from the input: every FULLY_CONNECTED becomes one Dense layer, the
    constant MUL/ADD that follow its ReLU (the exported BatchNormalization) are folded into the
    NEXT Dense exactly like corintho_ai_b200.fold_batchnorm does. The two heads are the
    FULLY_CONNECTED ops with 1 output (tanh value) and 96 outputs (softmax policy)."""
    tensors, ops, inputs, outputs = parse_tflite(path)
    if len(inputs) != 1:
        raise ValueError("expected a single input tensor")
    producers = {}
    for op in ops:
        for o in op["outputs"]:
            producers[o] = op
    consumers = {}
    for op in ops:
        for i in op["inputs"]:
            consumers.setdefault(i, []).append(op)

    def const(idx):
        return tensors[idx]["data"] if idx >= 0 else None

    cur = inputs[0]
    s = np.ones(70, np.float64)   # pending affine of the activation feeding the next Dense
    t = np.zeros(70, np.float64)
    hidden = []
    heads = {}
    guard = 0
    while True:
        guard += 1
        if guard > 200:
            raise ValueError("operator chain too long")
        nxt = consumers.get(cur, [])
        if not nxt:
            break
        fcs = [op for op in nxt if op["op"] == OP_FULLY_CONNECTED]
        if len(fcs) == 2 and len(nxt) == 2:  # the two heads branch off the last hidden activation
            for op in fcs:
                W = const(op["inputs"][1]).astype(np.float64)     # [out, in]
                b = const(op["inputs"][2]) if len(op["inputs"]) > 2 and op["inputs"][2] >= 0 else None
                b = np.zeros(W.shape[0]) if b is None else b.astype(np.float64)
                Wf = (W * s[None, :]).T                            # [in, out]
                bf = W @ t + b
                heads[W.shape[0]] = (Wf, bf)
            break
        if len(nxt) != 1:
            raise ValueError("unexpected branching in the exported graph")
        op = nxt[0]
        if op["op"] == OP_FULLY_CONNECTED:
            W = const(op["inputs"][1]).astype(np.float64)          # [out, in]
            b = const(op["inputs"][2]) if len(op["inputs"]) > 2 and op["inputs"][2] >= 0 else None
            b = np.zeros(W.shape[0]) if b is None else b.astype(np.float64)
            hidden.append(((W * s[None, :]).T, W @ t + b))
            s, t = np.ones(W.shape[0]), np.zeros(W.shape[0])
            if op["act"] not in (ACT_NONE, ACT_RELU):
                raise ValueError("unexpected fused activation on a hidden Dense")
        elif op["op"] == OP_RELU:
            pass
        elif op["op"] == OP_MUL:
            c = const(op["inputs"][1]) if const(op["inputs"][1]) is not None else const(op["inputs"][0])
            s, t = s * c.astype(np.float64).ravel(), t * c.astype(np.float64).ravel()
        elif op["op"] == OP_ADD:
            c = const(op["inputs"][1]) if const(op["inputs"][1]) is not None else const(op["inputs"][0])
            t = t + c.astype(np.float64).ravel()
        else:
            raise ValueError(f"unsupported operator {op['op']} in the hidden chain")
        cur = op["outputs"][0]
    if len(hidden) != 12 or 1 not in heads or 96 not in heads:
        raise ValueError(f"unexpected architecture: {len(hidden)} hidden layers, heads {sorted(heads)}")
    out = []
    for Wf, bf in hidden:
        out.append(Wf.astype(np.float32).ravel())
        out.append(bf.astype(np.float32))
    Wh = np.concatenate([heads[1][0], heads[96][0]], 1)
    bh = np.concatenate([heads[1][1], heads[96][1]])
    out.append(Wh.astype(np.float32).ravel())
    out.append(bh.astype(np.float32))
    flat = np.concatenate(out)
    if flat.size != 127997:
        raise ValueError(f"unexpected weight count {flat.size}")
    return flat
