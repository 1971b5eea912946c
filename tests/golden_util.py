"""Helpers for the committed golden vectors (tests/golden/*.npz, made by oracle/make_golden.py)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sc_cases():
    return sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "sc_*.npz")))


def beergame_cases():
    return sorted(os.path.basename(p)[9:-4] for p in glob.glob(os.path.join(GOLDEN, "beergame_*.npz")))


def load(kind, name):
    z = np.load(os.path.join(GOLDEN, "%s_%s.npz" % (kind, name)))
    meta = json.loads(str(z["meta"]))
    return meta, z

