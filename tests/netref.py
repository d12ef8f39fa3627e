"""CPU reference of the policy/value network (test infrastructure): a numpy restatement of
corintho_ai/python/wrapper.py:256-271 with standard Keras inference semantics
(Dense y = xW + b, W [in,out]; ReLU; BatchNormalization gamma (x-mean)/sqrt(var+1e-3) + beta;
heads Dense(1,tanh), Dense(96,softmax)). PARITY UNPINNED by the reference: Keras itself is an
un-vendored third-party dependency (keras 2.9-2.12 per scripts/bash/startup_script.sh:7-8 and
model/keras_metadata.pb) and no reference test touches the network."""
import numpy as np


def forward_unfolded(params, x, dtype=np.float64):
    h = np.asarray(x, dtype)
    for L in params["layers"]:
        h = h @ L["W"].astype(dtype) + L["b"].astype(dtype)
        h = np.maximum(h, 0)
        h = L["gamma"].astype(dtype) * (h - L["mean"].astype(dtype)) / np.sqrt(L["var"].astype(dtype) + dtype(1e-3)) \
            + L["beta"].astype(dtype)
    H = params["head"]
    v = np.tanh(h @ H["Wv"].astype(dtype) + H["bv"].astype(dtype))[:, 0]
    z = h @ H["Wp"].astype(dtype) + H["bp"].astype(dtype)
    z = z - z.max(1, keepdims=True)
    e = np.exp(z)
    return v, e / e.sum(1, keepdims=True)


def forward_folded(flat, x, dtype=np.float32, round_bf16=False):
    """Same network from the folded C-ABI weight vector (the layout the engine consumes)."""
    def bf16(a):
        if not round_bf16:
            return a
        b = np.ascontiguousarray(a, np.float32).view(np.uint32)
        b = ((b + np.uint32(0x7FFF) + ((b >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000))
        return b.view(np.float32)

    h = bf16(np.asarray(x, np.float32)).astype(dtype)
    off = 0
    dims = [70] + [100] * 12 + [97]
    for l in range(13):
        K, N = dims[l], dims[l + 1]
        W = bf16(flat[off:off + K * N].reshape(K, N)).astype(dtype)
        off += K * N
        b = flat[off:off + N].astype(dtype)
        off += N
        h = h @ W + b
        if l < 12:
            h = bf16(np.maximum(h, 0).astype(np.float32)).astype(dtype)
    v = np.tanh(h[:, 0])
    z = h[:, 1:] - h[:, 1:].max(1, keepdims=True)
    e = np.exp(z)
    return v.astype(np.float32), (e / e.sum(1, keepdims=True)).astype(np.float32)
