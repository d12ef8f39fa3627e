"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares.
No compute entry point is called here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "corintho_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_trainer_surface():
    syms = declared_symbols()
    for name in ["create", "destroy", "do_iteration", "num_requests", "write_requests", "num_samples",
                 "write_samples", "score", "avg_mate_length", "write_scores"]:
        assert "cb200_trainer_" + name in syms


def test_library_exports_every_declared_symbol():
    path = os.path.join(ROOT, "corintho_ai_b200", "libcorintho_b200.so")
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    lib.cb200_last_error.restype = ctypes.c_char_p
    assert lib.cb200_last_error() is not None


def test_python_mirror_has_reference_method_names():
    import corintho_ai_b200 as cb
    for m in ["doIteration", "num_requests", "writeRequests", "num_samples", "writeSamples", "score",
              "avg_mate_length", "writeScores"]:
        assert hasattr(cb.Trainer, m)


def test_no_cpu_fallback_without_gpu():
    """Without a usable GPU the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import corintho_ai_b200 as cb
    with pytest.raises(cb.Corintho200Error):
        cb.Trainer(2, "", 1, 16, 4)
    import numpy as np
    with pytest.raises(cb.Corintho200Error):
        cb.game_step(np.zeros((4, 2), np.uint64))


def test_product_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under corintho_ai_b200/ may import/link it."""
    pkg = os.path.join(ROOT, "corintho_ai_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "corintho_oracle" not in txt and "libcorintho_ref" not in txt, f


def test_weight_folding_layout():
    import numpy as np
    import corintho_ai_b200 as cb
    p = cb.random_weights(3)
    flat = cb.fold_batchnorm(p)
    assert flat.size == cb.WEIGHT_FLOATS == 70 * 100 + 100 + 11 * (100 * 100 + 100) + 100 * 97 + 97
    # first layer is unscaled; second layer rows are scaled by 1/sqrt(1+1e-3)
    assert np.array_equal(flat[:7000].reshape(70, 100), p["layers"][0]["W"])
    s = np.float64(1.0) / np.sqrt(1.0 + 1e-3)
    w2 = flat[7100:7100 + 10000].reshape(100, 100)
    assert np.allclose(w2, p["layers"][1]["W"] * s, rtol=1e-6)


def test_python_mirror_validates_buffers_like_the_cython_binding():
    """The reference's Cython layer takes typed float32 memoryviews (main.pyx:30-38) and rejects
    anything else; the ctypes mirror passes raw pointers, so it checks dtype / layout / size itself
    (advisor finding, round 1). No GPU needed: the checks run before the library is called."""
    import numpy as np
    import pytest
    import corintho_ai_b200 as cb
    ok = np.zeros((4, 70), np.float32)
    assert cb._out_f32(ok, 280, "x") is ok
    for bad in (np.zeros((4, 70), np.float64), np.zeros((8, 70), np.float32)[::2], [0.0] * 280):
        with pytest.raises(cb.Corintho200Error):
            cb._out_f32(bad, 280, "x")
    with pytest.raises(cb.Corintho200Error):
        cb._out_f32(ok, 281, "x")                       # too small
    ro = np.zeros(280, np.float32)
    ro.setflags(write=False)
    with pytest.raises(cb.Corintho200Error):
        cb._out_f32(ro, 280, "x")                       # not writable
    # inputs are converted (float64 / strided views become contiguous float32) but never short
    conv = cb._in_f32(np.arange(8, dtype=np.float64)[::2], 4, "y")
    assert conv.dtype == np.float32 and conv.flags.c_contiguous and conv.tolist() == [0.0, 2.0, 4.0, 6.0]
    with pytest.raises(cb.Corintho200Error):
        cb._in_f32(np.zeros(3, np.float32), 4, "y")
    with pytest.raises(cb.Corintho200Error):
        cb._in_f32(None, 1, "y")
    assert cb._in_f32(None, 0, "y") is None
