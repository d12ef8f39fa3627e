"""Edge cases of the fused paths: tiny budgets, a single game, random-only tourneys, maximal
searches_per_eval -- each compared with the oracle (or the lock-step path) on the same inputs."""
import numpy as np
import pytest

import corintho_ai_b200 as cb
from util import run_tourney, run_trainer

pytestmark = pytest.mark.gpu


def _fused_vs_oracle(oracle, cfg, prec):
    flat = cb.fold_batchnorm(cb.random_weights(cfg["seed"]))
    t = cb.Trainer(cfg["num_games"], "", cfg["seed"], cfg["max_searches"], cfg["searches_per_eval"],
                   cfg["c_puct"], cfg["epsilon"])
    t.set_weights(flat, 0, prec)
    assert t.run_selfplay(0, stagger=True)
    helper = cb.Trainer(max(cfg["num_games"], 4), "", 1, 64, cfg["searches_per_eval"])
    helper.set_weights(flat, 0, prec)
    r = run_trainer(oracle.trainer(**cfg), lambda req: helper.evaluate(req))
    gs, ev, pr = t.write_samples()
    assert t.num_samples() == r["num_samples"]
    assert gs.tobytes() == r["samples"][0].tobytes() and pr.tobytes() == r["samples"][2].tobytes()
    assert ev.tobytes() == r["samples"][1].tobytes()
    assert t.score().tobytes() == r["score"].tobytes()


@pytest.mark.parametrize("cfg", [
    dict(num_games=1, seed=3, max_searches=24, searches_per_eval=8, c_puct=1.0, epsilon=0.25),
    dict(num_games=9, seed=4, max_searches=5, searches_per_eval=1, c_puct=2.0, epsilon=0.0),
    dict(num_games=33, seed=6, max_searches=16, searches_per_eval=16, c_puct=1.0, epsilon=1.0),
    dict(num_games=3, seed=8, max_searches=1, searches_per_eval=1, c_puct=1.0, epsilon=0.25),
], ids=lambda c: "g%d_m%d_e%d" % (c["num_games"], c["max_searches"], c["searches_per_eval"]))
def test_small_fused_runs_equal_oracle(oracle, cfg):
    _fused_vs_oracle(oracle, cfg, "bf16")


def test_random_only_and_wide_budget_tourneys(oracle):
    class E:
        def tourney(self, n, f):
            return cb.Tourney(n, f)

    def build(L):
        t = L.tourney(1, "")
        t.add_player(0, -1, 1, 1, 1.0, 0.25, True)
        t.add_player(1, 2, 64, 64, 1.0, 0.25)   # searches_per_eval = max_searches = 64
        t.add_player(2, 2, 3, 1, 4.0, 0.0)
        for a, b in [(0, 0), (0, 0), (1, 2), (2, 1), (1, 0), (0, 2), (1, 1)]:
            t.add_match(a, b)
        return t
    a, b = run_tourney(build(oracle)), run_tourney(build(E()))
    assert a["rounds"] == b["rounds"] and a["req_hash"] == b["req_hash"]
    assert (a["scores"] == b["scores"]).all() and (a["counts"] == b["counts"]).all()


def test_arena_overflow_in_fused_and_persistent_runs_fails_loudly(monkeypatch):
    """A node arena that is too small must end the fused run with CB200_ERR_OVERFLOW -- from the
    lock-step loop and from inside the persistent kernel alike -- never with a hang or bad data."""
    monkeypatch.setenv("CB200_ARENA_NODES", "24")
    flat = cb.fold_batchnorm(cb.random_weights(2))
    for no_persistent in ("1", None):
        if no_persistent:
            monkeypatch.setenv("CB200_NO_PERSISTENT", no_persistent)
        else:
            monkeypatch.delenv("CB200_NO_PERSISTENT")
        t = cb.Trainer(16, "", 3, 64, 16, 1.0, 0.25)
        t.set_weights(flat, 0, "bf16")
        with pytest.raises(cb.Corintho200Error) as e:
            t.run_selfplay(0, stagger=False)
        assert "arena" in str(e.value)
