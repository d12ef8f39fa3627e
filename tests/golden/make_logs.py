#!/usr/bin/env python
"""Generate tests/golden/logs/*.txt: the per-game text logs (selfplayer.cpp:124-204) the COMPILED
REFERENCE (oracle/_ref) writes for the first num_logged games of two small runs under the
synthetic evaluator. Run in the build container only:

    make -C oracle ref && python tests/golden/make_logs.py
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle.pyoracle import RefLib, play_out, play_tourney  # noqa: E402
from util import LOG_CASES, LOGGED_MATCHES, TOURNEY_CASES  # noqa: E402


def main():
    R = RefLib()
    out = os.path.join(HERE, "logs")
    os.makedirs(out, exist_ok=True)
    for name, (kw, to_play) in LOG_CASES.items():
        d = tempfile.mkdtemp()
        t = R.trainer(log_folder=d, **kw)
        play_out(t, to_play=to_play)
        t.close()
        for f in sorted(os.listdir(d)):
            shutil.copy(os.path.join(d, f), os.path.join(out, name + "_" + f))
            print(name, f, os.path.getsize(os.path.join(d, f)))
        shutil.rmtree(d)
    # Tourney match logs (match.cpp:79-180): the "with_random" field with some matches logged
    d = tempfile.mkdtemp()
    players, matches = TOURNEY_CASES["with_random"]
    t = R.tourney(1, d)
    for pl in players:
        t.add_player(*pl)
    for i, (a, b) in enumerate(matches):
        t.add_match(a, b, i in LOGGED_MATCHES)
    play_tourney(t)
    t.close()
    for f in sorted(os.listdir(d)):
        shutil.copy(os.path.join(d, f), os.path.join(out, f))
        print("tourney", f, os.path.getsize(os.path.join(d, f)))
    shutil.rmtree(d)


if __name__ == "__main__":
    main()
