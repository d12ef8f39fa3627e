"""Reference on-disk sample format (main.pyx:189-204): three .npz files keyed arr_0."""
import numpy as np

from corintho_ai_b200.samples import load_samples, save_samples


class _OracleAdapter:
    """Gives an oracle trainer the reference-named writeSamples signature."""

    def __init__(self, t):
        self.t = t

    def num_samples(self):
        return self.t.num_samples()

    def writeSamples(self, gs, ev, pr):
        a, b, c = self.t.write_samples()
        gs[:], ev[:], pr[:] = a, b, c


def test_sample_folder_round_trip(oracle, tmp_path):
    from oracle.pyoracle import play_out
    t = oracle.trainer(num_games=3, seed=5, max_searches=16, searches_per_eval=8)
    play_out(t)
    rows = save_samples(_OracleAdapter(t), str(tmp_path / "samples" / "gen_0"))
    assert rows == t.num_samples() * 8
    for name in ("game_states", "evaluation_labels", "probability_labels"):
        z = np.load(tmp_path / "samples" / "gen_0" / (name + ".npz"))
        assert z.files == ["arr_0"] and z["arr_0"].dtype == np.float32
    gs, ev, pr = load_samples(str(tmp_path / "samples" / "gen_0"))
    a, b, c = t.write_samples()
    assert gs.tobytes() == a.tobytes() and ev.tobytes() == b.tobytes() and pr.tobytes() == c.tobytes()
