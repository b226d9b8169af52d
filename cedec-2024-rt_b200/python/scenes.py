"""Scene helpers of the headless harness: binary Triangle[] caches, tiling (BASELINE config 5) and a procedural
stand-in scene for machines where the reference's assets were not staged."""
import lzma
import os

import numpy as np

from cedecrt import TRIANGLE

_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
ASSET_DIRS = (os.path.join(_ROOT, "assets"),)


def find_scene(name):
    for d in ASSET_DIRS:
        p = os.path.join(d, name + ".tri.xz")
        if os.path.exists(p):
            return p
    return None


def load_scene(name):
    """Triangle[] exactly as the reference loader produced it (staged as assets/<name>.tri.xz)."""
    p = find_scene(name)
    if p is None:
        raise FileNotFoundError("scene cache %s.tri.xz not found under %s" % (name, ASSET_DIRS))
    with lzma.open(p, "rb") as f:
        return np.frombuffer(f.read(), dtype=TRIANGLE).copy()


def tile_scene(tris, nx, nz, pitch_x, pitch_z):
    """BASELINE config 5 / SURVEY.md section 8d: nx x nz copies on a grid in x,z, tile-major triangle order so
    that primID = tile * len(tris) + i; tile 0 is the original (the camera stays inside it)."""
    out = np.empty(len(tris) * nx * nz, TRIANGLE)
    k = 0
    for iz in range(nz):
        for ix in range(nx):
            t = tris.copy()
            t["vertices"][:, :, 0] += np.float32(ix * pitch_x)
            t["vertices"][:, :, 2] += np.float32(iz * pitch_z)
            out[k * len(tris):(k + 1) * len(tris)] = t
            k += 1
    return out


def procedural_blocks(n_blocks=20000, seed=1, extent=60.0, emissive_fraction=0.08):
    """A city of axis-aligned boxes on a ground plane with some emissive boxes: same character as
    blocks_restir (12 triangles per block). Used only when the real asset cache is absent."""
    rng = np.random.default_rng(seed)
    tris = np.zeros(n_blocks * 12 + 2, TRIANGLE)
    cx = rng.uniform(-extent, extent, n_blocks).astype(np.float32)
    cz = rng.uniform(-extent, extent, n_blocks).astype(np.float32)
    sx = rng.uniform(0.3, 1.5, n_blocks).astype(np.float32)
    sz = rng.uniform(0.3, 1.5, n_blocks).astype(np.float32)
    h = rng.uniform(0.5, 25.0, n_blocks).astype(np.float32) * (rng.random(n_blocks) < 0.3) + rng.uniform(
        0.3, 4.0, n_blocks).astype(np.float32)
    col = rng.uniform(0.2, 0.9, (n_blocks, 3)).astype(np.float32)
    emi = (rng.random(n_blocks) < emissive_fraction)[:, None] * rng.uniform(2, 40, (n_blocks, 3)).astype(np.float32)
    corners = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1], [-1, 1, -1], [1, 1, -1], [1, 1, 1],
                        [-1, 1, 1]], np.float32)
    faces = [(0, 2, 1), (0, 3, 2), (4, 5, 6), (4, 6, 7), (0, 1, 5), (0, 5, 4), (1, 2, 6), (1, 6, 5), (2, 3, 7),
             (2, 7, 6), (3, 0, 4), (3, 4, 7)]
    scale = np.stack([sx, h.astype(np.float32), sz], 1)
    base = np.stack([cx, np.zeros_like(cx), cz], 1)
    v = corners[None] * scale[:, None] + base[:, None]  # (n,8,3)
    for f, (a, b, c) in enumerate(faces):
        tris["vertices"][f:n_blocks * 12:12, 0] = v[:, a]
        tris["vertices"][f:n_blocks * 12:12, 1] = v[:, b]
        tris["vertices"][f:n_blocks * 12:12, 2] = v[:, c]
        tris["color"][f:n_blocks * 12:12] = col
        tris["emissive"][f:n_blocks * 12:12] = emi
    e = np.float32(extent * 1.2)
    g = n_blocks * 12
    tris["vertices"][g] = [[-e, 0, -e], [e, 0, e], [e, 0, -e]]
    tris["vertices"][g + 1] = [[-e, 0, -e], [-e, 0, e], [e, 0, e]]
    tris["color"][g:] = 0.6
    return tris
