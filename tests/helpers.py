"""Shared helpers for the test-suite (test infrastructure)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RES_FIELDS = ("origin_position", "origin_normal", "hit_position", "hit_normal", "radiance", "visibility", "w_sum",
              "ucw", "M")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def small_scene(name):
    return np.ascontiguousarray(golden("scenes_small.npz")[name])


def same(a, b):
    """bit-level equality that treats NaN == NaN (compares raw bytes of float arrays)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def reservoir_mismatch(a, b):
    """number of pixels whose reservoirs differ in any field (padding bytes ignored)."""
    bad = np.zeros(len(a), bool)
    for f in RES_FIELDS:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            x, y = x.view(np.uint32), y.view(np.uint32)
        d = x != y
        bad |= d.reshape(len(a), -1).any(1)
    return int(bad.sum())
