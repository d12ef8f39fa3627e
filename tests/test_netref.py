"""BatchNorm folding is exact algebra: folded fp64 forward == unfolded fp64 forward."""
import numpy as np

import corintho_ai_b200 as cb
from netref import forward_folded, forward_unfolded
from oracle.pyoracle import OracleLib, RefLib, have_ref


def test_folding_matches_unfolded_network():
    rng = np.random.default_rng(0)
    p = cb.random_weights(5)
    for L in p["layers"]:  # non-trivial BN statistics
        n = L["gamma"].size
        L["gamma"] = rng.uniform(0.5, 1.5, n).astype(np.float32)
        L["beta"] = rng.uniform(-0.2, 0.2, n).astype(np.float32)
        L["mean"] = rng.uniform(-0.1, 0.3, n).astype(np.float32)
        L["var"] = rng.uniform(0.5, 2.0, n).astype(np.float32)
        L["b"] = rng.uniform(-0.1, 0.1, n).astype(np.float32)
    flat = cb.fold_batchnorm(p)
    O = OracleLib()
    st = O.start()
    xs = [O.encode(st)]
    for m in (48, 70, 90, 55):
        st = O.do_move(st, m)
        xs.append(O.encode(st))
    x = np.stack(xs)
    v0, p0 = forward_unfolded(p, x, np.float64)
    v1, p1 = forward_folded(flat, x, np.float64)
    assert np.allclose(v0, v1, rtol=1e-5, atol=1e-6) and np.allclose(p0, p1, rtol=1e-5, atol=1e-7)
    assert np.allclose(p1.sum(1), 1, atol=1e-5)
