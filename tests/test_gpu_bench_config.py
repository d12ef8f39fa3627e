"""Parity at the BENCHMARKED configuration (BASELINE.json configs[2] and configs[4] shapes).

bench.py times fused bf16 self-play of 4096 games x 800 sims with the default scheduling: six
stream groups, parking, the 16-games-per-CTA persistent kernel that takes over at <= 2368 live
games, and the arena sizing of a full-size trainer. None of the small parity tests reaches those
code paths, so they are pinned here:

  * default scheduling == plain lock-step scheduling (one group, no persistent kernel), byte for
    byte on samples, scores and exact counters;
  * 64-game shards of that very run == the CPU oracle (reference restatement, trainer.cpp:164-236
    / trainmc.cpp:602-696) driven by the same network, byte for byte;
  * the wide (16 games per CTA) persistent kernel is asserted to have run;
  * the two-model persistent path at the configs[4] per-GPU shape (1250 games x 1600 sims, no
    noise) == the lock-step two-model path, and a 64-game subset == the oracle.
"""
import numpy as np
import pytest

import corintho_ai_b200 as cb
from util import run_trainer, sha

pytestmark = pytest.mark.gpu

G, SIMS, SPE, SEED = 4096, 800, 16, 2000


def _digest(t):
    gs, ev, pr = t.write_samples()
    c = t.counters()
    return {"samples": sha(gs, ev, pr), "score": t.score().tobytes(), "mate": t.avg_mate_length().tobytes(),
            "results": t.game_results().tobytes(),
            "counters": (c["simulations"], c["moves"], c["leaf_evals"])}, (gs, ev, pr)


@pytest.fixture(scope="module")
def bench_run():
    """The run bench.py times: fused bf16, default scheduling, 4096 x 800, no staggering."""
    flat = cb.fold_batchnorm(cb.random_weights(0))
    t = cb.Trainer(G, "", 12345, SIMS, SPE, 1.0, 0.25)
    t.set_weights(flat, 0, "bf16")
    t.set_profiling(True)
    t.reset(SEED)
    assert t.run_selfplay(0, stagger=False)
    kt = t.kernel_times()
    split = t.phase_split()
    dig, samples = _digest(t)
    st, pr, lb, game_of = t.raw_samples()
    out = {"flat": flat, "digest": dig, "samples": samples, "game_of": game_of.copy(), "kt": kt,
           "split": split, "results": t.game_results().copy()}
    t.close()
    return out


def test_benchmark_run_uses_the_scheduling_it_claims(bench_run):
    kt, split = bench_run["kt"], bench_run["split"]
    assert kt["game_step"]["launches"] > 1000 and kt["network"]["launches"] > 1000
    assert kt["fused_tail_wide"]["launches"] >= 1, "k_selfplay_persistent<*,16> never ran"
    assert kt["fused_tail"]["launches"] >= 1, "k_selfplay_persistent<*,8> never ran"
    # the running search counter ends equal to the simulation count of the finished games
    assert split["simulations"] == bench_run["digest"]["counters"][0]
    assert split["leaf_evals"] == bench_run["digest"]["counters"][2]
    assert 0 < split["lockstep_simulations"] < split["simulations"]


def test_default_scheduling_equals_plain_lock_step(bench_run, monkeypatch):
    """Six stream groups + parking + both persistent kernels vs ONE group, lock-step all the
    way: identical samples (sha256 of all 8-fold augmented rows), scores, results, counters."""
    monkeypatch.setenv("CB200_GROUPS", "1")
    monkeypatch.setenv("CB200_NO_PERSISTENT", "1")
    monkeypatch.setenv("CB200_YIELD", "0")
    t = cb.Trainer(G, "", 12345, SIMS, SPE, 1.0, 0.25)
    t.set_weights(bench_run["flat"], 0, "bf16")
    t.set_profiling(True)
    t.reset(SEED)
    assert t.run_selfplay(0, stagger=False)
    kt = t.kernel_times()
    assert kt["fused_tail"]["launches"] == 0 and kt["fused_tail_wide"]["launches"] == 0
    dig, _ = _digest(t)
    t.close()
    assert dig == bench_run["digest"]


@pytest.mark.parametrize("first", [0, 2048, 4032])
def test_shard_of_the_benchmark_run_equals_the_oracle(bench_run, oracle, first):
    """Games are independent units (trainer.cpp:175-196): games [first, first+64) of the timed
    4096-game run must equal the oracle playing those 64 games with the same network."""
    n = 64
    helper = cb.Trainer(n, "", 1, 64, SPE)
    helper.set_weights(bench_run["flat"], 0, "bf16")
    o = oracle.trainer(num_games=n, seed=SEED, max_searches=SIMS, searches_per_eval=SPE, c_puct=1.0,
                       epsilon=0.25, num_threads=8, first_game=first)
    r = run_trainer(o, lambda req: helper.evaluate(req))
    game_of = bench_run["game_of"]
    lo, hi = np.searchsorted(game_of, first), np.searchsorted(game_of, first + n)
    gs, ev, pr = bench_run["samples"]
    assert hi - lo == r["num_samples"]
    assert gs[8 * lo:8 * hi].tobytes() == r["samples"][0].tobytes()
    assert ev[8 * lo:8 * hi].tobytes() == r["samples"][1].tobytes()
    assert pr[8 * lo:8 * hi].tobytes() == r["samples"][2].tobytes()


def test_wide_persistent_kernel_short_run(monkeypatch):
    """2400 games x 64 sims: every game fits the 16-games-per-CTA persistent kernel from the
    first iteration after the roots (2400 > 148 * 8). Must equal the lock-step run."""
    flat = cb.fold_batchnorm(cb.random_weights(3))

    def run():
        t = cb.Trainer(2400, "", 5, 64, 16, 1.0, 0.25)
        t.set_weights(flat, 0, "bf16")
        t.set_profiling(True)
        assert t.run_selfplay(0, stagger=False)
        kt = t.kernel_times()
        d, _ = _digest(t)
        t.close()
        return d, kt

    a, kt = run()
    assert kt["fused_tail_wide"]["launches"] >= 1
    monkeypatch.setenv("CB200_NO_PERSISTENT", "1")
    monkeypatch.setenv("CB200_GROUPS", "1")
    b, kt2 = run()
    assert kt2["fused_tail_wide"]["launches"] == 0 and kt2["fused_tail"]["launches"] == 0
    assert a == b


def test_two_model_match_at_config4_shape(oracle, monkeypatch):
    """configs[4] per-GPU shape: 1250 games, 1600 sims, epsilon 0, two networks, testing mode
    (main.pyx:329-350). Persistent two-model kernel == lock-step two-model path; games [0, 64)
    == the oracle driven by the same two networks (clean alternation, see test_gpu_net.py)."""
    fa, fb = cb.fold_batchnorm(cb.random_weights(1)), cb.fold_batchnorm(cb.random_weights(2))
    Gm, MS = 1250, 1600

    def run():
        t = cb.Trainer(Gm, "", 12345, MS, SPE, 1.0, 0.0, 0, 1, True)
        t.set_weights(fa, 0, "bf16")
        t.set_weights(fb, 1, "bf16")
        t.set_profiling(True)
        assert t.run_selfplay(0)
        c = t.counters()
        out = (t.game_results().copy(), t.score().tobytes(), (c["simulations"], c["moves"], c["leaf_evals"]))
        kt = t.kernel_times()
        t.close()
        return out, kt

    a, kt = run()
    assert kt["fused_tail"]["launches"] + kt["fused_tail_wide"]["launches"] >= 1
    monkeypatch.setenv("CB200_NO_PERSISTENT", "1")
    b, kt2 = run()
    assert kt2["fused_tail"]["launches"] + kt2["fused_tail_wide"]["launches"] == 0
    assert (a[0] == b[0]).all() and a[1:] == b[1:]
    # parity subset: the first 64 games against the oracle
    n = 64
    h = [cb.Trainer(n, "", 1, 16, SPE) for _ in range(2)]
    h[0].set_weights(fa, 0, "bf16")
    h[1].set_weights(fb, 0, "bf16")
    o = oracle.trainer(num_games=n, seed=12345, max_searches=MS, searches_per_eval=SPE, c_puct=1.0,
                       epsilon=0.0, testing=True, num_threads=8)
    ev = np.zeros(n * SPE, np.float32)
    pr = np.zeros((n * SPE, 96), np.float32)
    done, tp = False, 0
    while not done:
        k = o.num_requests(tp)
        if k:
            e, p = h[tp].evaluate(o.write_requests(tp))
            ev[:k], pr[:k] = e, p
        done = o.do_iteration(ev, pr, tp)
        tp = 1 - tp
    assert (a[0][:n] == oracle.game_results(o)).all()
