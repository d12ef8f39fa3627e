"""On-disk training-sample format of the reference.

The reference saves three compressed numpy archives per generation under
``<run>/samples/gen_<k>/`` (corintho_ai/python/main.pyx:189-204, wrapper.py:191-199):
``game_states.npz`` [N*8, 70], ``evaluation_labels.npz`` [N*8], ``probability_labels.npz``
[N*8, 96], each holding one float32 array under the default key ``arr_0``
(np.savez_compressed with a positional argument)."""
import os

import numpy as np

FILES = ("game_states", "evaluation_labels", "probability_labels")


def save_samples(trainer, sample_folder):
    """Write the trainer's samples (8 symmetries, Trainer::writeSamples order) exactly where and
    how main.pyx:189-204 does. Returns the number of rows written."""
    n = trainer.num_samples()
    gs = np.zeros((n * 8, 70), np.float32)
    ev = np.zeros(n * 8, np.float32)
    pr = np.zeros((n * 8, 96), np.float32)
    if n:
        trainer.writeSamples(gs, ev, pr)
    os.makedirs(sample_folder, exist_ok=True)
    np.savez_compressed(os.path.join(sample_folder, "game_states"), gs)
    np.savez_compressed(os.path.join(sample_folder, "evaluation_labels"), ev)
    np.savez_compressed(os.path.join(sample_folder, "probability_labels"), pr)
    return n * 8


def load_samples(sample_folder):
    """Read a reference-format sample folder (main.pyx:207-216 reads key ``arr_0``)."""
    out = []
    for name in FILES:
        with np.load(os.path.join(sample_folder, name + ".npz")) as z:
            out.append(z["arr_0"])
    gs, ev, pr = out
    return gs.reshape(-1, 70), ev, pr.reshape(-1, 96)
