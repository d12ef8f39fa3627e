"""The reference-side binding, for real: bindings/cython/corintho_b200_cy.pyx keeps the reference's
`cdef cppclass Trainer / Tourney` blocks (python/main.pyx:17-38, rating/tourney.pyx:15-32) and reads
them from include/corintho_b200.hpp instead of the reference's .cpp files. It is compiled here with
Cython + g++ (no CUDA toolchain involved) and the engine is driven through it with the reference's
own loops; results must equal the oracle's bit for bit."""
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cy(tmp_path_factory):
    pytest.importorskip("Cython")
    if not os.path.exists(os.path.join(ROOT, "corintho_ai_b200", "libcorintho_b200.so")):
        import __graft_entry__
        __graft_entry__.build()
    d = tmp_path_factory.mktemp("cy")
    env = dict(os.environ, CB200_CY_BUILD=str(d / "build"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bindings", "cython", "setup.py"), "-q", "build_ext",
                        "--build-lib", str(d / "out"), "--build-temp", str(d / "tmp")],
                       cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    sys.path.insert(0, str(d / "out"))
    try:
        yield importlib.import_module("corintho_b200_cy")
    finally:
        sys.path.remove(str(d / "out"))


def test_binding_compiles_imports_and_raises_through_except_plus(cy):
    """No GPU needed: argument validation precedes every CUDA call, and the C++ exception of the
    shim class arrives as a Python exception like the reference's `except +` members."""
    assert callable(cy.play_games) and callable(cy.run_tourney)
    with pytest.raises(RuntimeError, match="invalid argument"):
        cy.trainer_rejects(0, 4)


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(24, 9, 64, 8, 1.0, 0.25), (7, 3, 200, 16, 1.5, 0.1)])
def test_play_games_through_cython_equals_oracle(cy, oracle, cfg):
    from oracle.pyoracle import play_out, synth_eval
    g, seed, ms, spe, cp, eps = cfg
    st, ev, pr, score, mate = cy.play_games(g, "", seed, ms, spe, cp, eps, 0, 1, False, synth_eval)
    o = oracle.trainer(num_games=g, seed=seed, max_searches=ms, searches_per_eval=spe, c_puct=cp, epsilon=eps)
    play_out(o, synth_eval)
    ost, oev, opr = o.write_samples()
    assert st.shape == ost.shape and st.shape[0] == o.num_samples() * 8
    assert _digest(st, ev, pr) == _digest(ost, oev, opr)
    assert np.float32(score).tobytes() == np.float32(o.score()).tobytes()
    assert np.float32(mate).tobytes() == np.float32(o.avg_mate_length()).tobytes()


@pytest.mark.gpu
def test_testing_mode_through_cython_equals_oracle(cy, oracle, tmp_path):
    """main.pyx:151-182: two evaluators, to_play flips when one side has nothing pending."""
    from oracle.pyoracle import play_out, synth_eval

    def two(rows, to_play):
        e, p = synth_eval(rows)
        return (e, p) if to_play == 0 else (-e, p[:, ::-1].copy())

    seen = []
    out = cy.play_games(6, "", 11, 48, 6, 1.0, 0.0, 0, 1, True, lambda r, tp: (seen.append(tp), two(r, tp))[1],
                        scores_file=tmp_path / "scores.txt")
    o = oracle.trainer(num_games=6, seed=11, max_searches=48, searches_per_eval=6, c_puct=1.0, epsilon=0.0,
                       testing=True)
    rec = []
    # the oracle driver calls evaluator(rows); replay the same side sequence through `two`
    sides = iter(seen)
    play_out(o, lambda r: two(r, next(sides)), to_play=0, record=rec)
    assert [tp for tp, _ in rec] == seen
    assert np.float32(out[3]).tobytes() == np.float32(o.score()).tobytes()
    assert np.float32(out[4]).tobytes() == np.float32(o.avg_mate_length()).tobytes()
    assert (tmp_path / "scores.txt").read_text().strip() != ""


@pytest.mark.gpu
def test_run_tourney_through_cython_equals_oracle(cy, oracle, tmp_path):
    """rating/tourney.pyx's file formats and loop; scores and every request row vs the oracle."""
    from oracle.pyoracle import play_tourney, synth_eval
    from util import TOURNEY_CASES
    players, matches = TOURNEY_CASES["with_random"]
    assert [p[0] for p in players] == list(range(len(players)))     # the file format numbers players by line
    (tmp_path / "players.txt").write_text(
        f"{len(players)}\n" + "".join(f"{m} {ms} {spe} {cp} {eps} {int(rnd)}\n" for _, m, ms, spe, cp, eps, rnd in players))
    (tmp_path / "matches.txt").write_text(f"{len(matches)}\n" + "".join(f"{a} {b} 0\n" for a, b in matches))
    model_ids = []
    for p in players:
        if p[1] not in model_ids:
            model_ids.append(p[1])
    rec = []
    rounds = cy.run_tourney({m: synth_eval for m in model_ids}, str(tmp_path / "players.txt"),
                            str(tmp_path / "matches.txt"), str(tmp_path), 2, record=rec)

    t = oracle.tourney(2, "")
    for pl in players:
        t.add_player(*pl)
    for a, b in matches:
        t.add_match(a, b)
    orec = []
    assert play_tourney(t, None, record=orec) == rounds
    assert [m for m, _ in rec] == [m for m, _ in orec]
    assert _digest(*[r for _, r in rec]) == _digest(*[r for _, r in orec])
    got = np.array([[float(x) for x in line.split()] for line in (tmp_path / "scores.txt").read_text().splitlines()
                    if line.strip()])
    want = np.array(t.scores(), np.float64).reshape(-1, 3)
    assert got.shape == want.shape and (got == want).all()


@pytest.mark.gpu
def test_fused_selfplay_through_cython_equals_the_ctypes_mirror(cy):
    """setWeights + runSelfplay of the C++ header (network on the device) vs corintho_ai_b200.Trainer."""
    import corintho_ai_b200 as cb
    w = cb.fold_batchnorm(cb.random_weights(3))
    st, ev, pr, score, mate = cy.play_games_fused(96, "", 5, 64, 8, 1.0, 0.25, 0, w, 1)
    t = cb.Trainer(96, "", 5, 64, 8, 1.0, 0.25, 0, 1, False)
    t.set_weights(w, 0, "bf16")
    while not t.run_selfplay(0, True):
        pass
    n = t.num_samples()
    a, b, c = (np.zeros((n * 8, 70), np.float32), np.zeros(n * 8, np.float32), np.zeros((n * 8, 96), np.float32))
    t.writeSamples(a, b, c)
    assert st.shape[0] == n * 8 and _digest(st, ev, pr) == _digest(a, b, c)
    assert np.float32(score).tobytes() == np.float32(t.score()).tobytes()
