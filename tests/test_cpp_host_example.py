"""examples/selfplay_main.cpp: the self-play path driven from a C++ host through the C ABI only
(the reference's host language; no Python between the caller and the library)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "selfplay_main")


def build_example():
    lib_dir = os.path.join(ROOT, "corintho_ai_b200")
    if not os.path.exists(os.path.join(lib_dir, "libcorintho_b200.so")):
        pytest.skip("libcorintho_b200.so not built")
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "selfplay_main.cpp"), "-L" + lib_dir, "-lcorintho_b200",
           "-Wl,-rpath," + lib_dir, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_cpp_host_example_builds_against_the_c_abi_and_fails_loudly_without_a_gpu():
    exe = build_example()
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([exe, "4", "16"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_example_produces_the_python_mirrors_samples(tmp_path):
    """Same weights file, same configuration: the C++ host's dumped samples (completion order)
    equal the Python mirror's writeSamples rows after a stable sort by game."""
    import corintho_ai_b200 as cb
    exe = build_example()
    flat = cb.fold_batchnorm(cb.random_weights(17))
    wfile = tmp_path / "weights.f32"
    flat.astype("<f4").tofile(wfile)
    prefix = str(tmp_path / "out")
    r = subprocess.run([exe, "48", "40", "7", str(wfile), prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    game_of = np.fromfile(prefix + "_game_of.i32", np.int32)
    gs = np.fromfile(prefix + "_game_states.f32", np.float32).reshape(-1, 70)
    ev = np.fromfile(prefix + "_evaluation_labels.f32", np.float32)
    pr = np.fromfile(prefix + "_probability_labels.f32", np.float32).reshape(-1, 96)
    t = cb.Trainer(48, "", 7, 40, 16, 1.0, 0.25)
    t.set_weights(flat, 0, "bf16x3")
    assert t.run_selfplay(0, stagger=False)
    g2, e2, p2 = t.write_samples()
    order = np.argsort(game_of, kind="stable")
    rows = (order[:, None] * 8 + np.arange(8)[None, :]).ravel()
    assert len(game_of) == t.num_samples()
    assert gs[rows].tobytes() == g2.tobytes() and ev[rows].tobytes() == e2.tobytes() and pr[rows].tobytes() == p2.tobytes()
