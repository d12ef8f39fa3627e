#!/usr/bin/env python
"""Fold one of the reference's shipped checkpoints (corintho_ai/docker/tflite_model.tflite, the
network behind the public web app) into the engine's flat weight vector and store it as a test
fixture: tests/golden/trained_net.npz (key `flat`, 127 997 float32). Run in the build container:

    python tests/golden/make_weights.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from corintho_ai_b200.tflite_import import load_tflite_weights  # noqa: E402

SRC = "/root/reference/corintho_ai/docker/tflite_model.tflite"

if __name__ == "__main__":
    flat = load_tflite_weights(SRC)
    np.savez_compressed(os.path.join(HERE, "trained_net.npz"), flat=flat, source=np.bytes_(SRC))
    print("wrote", flat.size, "floats")
