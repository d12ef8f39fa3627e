"""Match / Tourney: the oracle restatement vs golden transcripts generated from
the compiled reference and vs the compiled reference itself."""
import os

import numpy as np
import pytest

from util import TOURNEY_CASES, make_tourney, run_tourney

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "tourney.npz"))


def check_against_golden(r, name):
    assert r["rounds"] == int(GOLD[name + "/rounds"])
    assert (r["models"] == GOLD[name + "/models"]).all()
    assert (r["counts"] == GOLD[name + "/counts"]).all()
    assert r["req_hash"].encode() == bytes(GOLD[name + "/req_hash"])
    assert (r["scores"] == GOLD[name + "/scores"]).all()


@pytest.mark.parametrize("name", list(TOURNEY_CASES))
def test_oracle_tourney_matches_golden(oracle, name):
    check_against_golden(run_tourney(make_tourney(oracle, name)), name)


def test_oracle_tourney_matches_compiled_reference(oracle, ref):
    a = run_tourney(make_tourney(ref, "with_random", 1))
    b = run_tourney(make_tourney(oracle, "with_random", 3))
    assert a["rounds"] == b["rounds"] and a["req_hash"] == b["req_hash"]
    assert (a["scores"] == b["scores"]).all()


def test_random_player_only_tourney_finishes_without_evaluations(oracle):
    """Two random players (match.cpp:27-33, 193-206): the whole game is played inside one
    doIteration and no request is ever made; scores are 0, 0.5 or 1."""
    t = oracle.tourney(1, "")
    t.add_player(7, -1, 1, 1, 1.0, 0.25, True)
    for _ in range(6):
        t.add_match(7, 7)
    r = run_tourney(t)
    assert r["rounds"] == 1 and len(r["counts"]) == 0
    assert set(r["scores"][:, 2]) <= {0.0, 0.5, 1.0} and len(r["scores"]) == 6
