"""TEST/BENCH INFRASTRUCTURE — stages the reference's scenes as binary Triangle[] caches.

The GPU box has no /root/reference, so the scenes travel with the repo snapshot as git-ignored files under
assets/<name>.tri.xz: the exact bytes of the reference loader's std::vector<Triangle>
(common/loader.hpp:11-66 + tinyobjloader 1.0.6, run through oracle/_ref/libref_loader.so), LZMA-compressed.
Nothing is committed; rerun here whenever oracle/_ref is rebuilt:   python oracle/stage_assets.py
"""
import lzma
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import orc  # noqa: E402

ASSETS = "/root/reference/assets/"
OUT = os.path.join(HERE, "..", "assets")
SCENES = ("cornellbox1", "blocks_ao", "blocks_pt", "blocks_restir")


def scene_path(name):
    return os.path.join(OUT, name + ".tri.xz")


def load_scene(name):
    """Triangle[] of a staged scene (numpy structured array, orc.TRIANGLE)."""
    with lzma.open(scene_path(name), "rb") as f:
        raw = f.read()
    return np.frombuffer(raw, dtype=orc.TRIANGLE).copy()


def have_scene(name):
    return os.path.exists(scene_path(name))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in SCENES:
        t0 = time.time()
        tris = orc.load_obj_reference(ASSETS + name + ".obj", ASSETS)
        with lzma.open(scene_path(name), "wb", preset=1) as f:
            f.write(tris.tobytes())
        print("%-14s %9d tris  %6.1f MB -> %6.1f MB  fnv1a64 %s  (%.1f s)" % (
            name, len(tris), tris.nbytes / 1e6, os.path.getsize(scene_path(name)) / 1e6, orc.fnv1a64(tris),
            time.time() - t0))


if __name__ == "__main__":
    main()
