#!/usr/bin/env python
"""MEASUREMENT INFRASTRUCTURE — runs the reference GPU comparator (baseline/_ref/bin/ref_gpu: the reference's own
Orochi/HIPRT CUDA build of examples/10_restir_di behind a headless host, built by `make -C oracle refgpu`) on the
GPU box and prints its JSON line; optionally compares its primitive-id image with this repo's raycast.

  python oracle/ref_gpu/run_ref_gpu.py [--width 3840 --height 2160 --tiles 3 2 --frames 8 --warmup 3] [--compare]

The scene is the staged Triangle[] cache (assets/blocks_restir.tri.xz — bytes of the reference loader's output),
decompressed to a scratch file; the binary tiles it like BASELINE config 5.  If HIPRT cannot run on this GPU the
binary (or this script) prints {"impl": "reference-hiprt", "unavailable": "..."} instead."""
import argparse
import json
import lzma
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BIN = os.path.join(ROOT, "baseline", "_ref", "bin")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--tiles", type=int, nargs=2, default=[3, 2])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scene", default="blocks_restir")
    ap.add_argument("--compare", action="store_true", help="compare the primitive-id image with this repo's raycast")
    ap.add_argument("--compare-radiance", action="store_true",
                    help="also render the same number of frames with this repo's fused frame and compare the mean radiance")
    ap.add_argument("--timeout", type=int, default=900)
    args = ap.parse_args()
    exe = os.path.join(BIN, "ref_gpu")
    if not os.path.exists(exe):
        print(json.dumps({"impl": "reference-hiprt", "unavailable": "baseline/_ref/bin/ref_gpu not built (make -C oracle refgpu)"}))
        return
    src = os.path.join(ROOT, "assets", args.scene + ".tri.xz")
    tmp = tempfile.mkdtemp(prefix="refgpu_")
    tri = os.path.join(tmp, args.scene + ".tri")
    with lzma.open(src, "rb") as f, open(tri, "wb") as g:
        g.write(f.read())
    vis_path = os.path.join(tmp, "vis.bin")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = BIN + ":/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    cmd = [exe, "--base", "../", "--tri", tri, "--width", str(args.width), "--height", str(args.height), "--tiles-x",
           str(args.tiles[0]), "--tiles-z", str(args.tiles[1]), "--frames", str(args.frames), "--warmup", str(args.warmup)]
    if args.compare:
        cmd += ["--dump-vis", vis_path]
    try:
        p = subprocess.run(cmd, cwd=BIN, env=env, capture_output=True, text=True, timeout=args.timeout)
    except subprocess.TimeoutExpired:
        print(json.dumps({"impl": "reference-hiprt", "unavailable": "timed out after %d s" % args.timeout}))
        return
    sys.stderr.write(p.stderr[-4000:])
    line = [l for l in p.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(json.dumps({"impl": "reference-hiprt", "unavailable": "exit code %d, no result line; stderr tail: %s" % (
            p.returncode, p.stderr[-300:].replace("\n", " | "))}))
        return
    out = json.loads(line[-1])
    if args.compare and os.path.exists(vis_path):
        out["primid_vs_cedecrt"] = compare(args, vis_path)
    if args.compare_radiance:
        out["radiance_vs_cedecrt"] = compare_radiance(args, out)
    print(json.dumps(out))


def compare(args, vis_path):
    """per-pixel closest-hit primitive ids: HIPRT's intersector vs this repo's raycast on the same scene and camera"""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "cedec-2024-rt_b200", "python"))
    import cedecrt
    import scenes

    tris = scenes.tile_scene(scenes.load_scene(args.scene), args.tiles[0], args.tiles[1], 130.0, 82.0)
    cam = ((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192))
    rt = cedecrt.Runtime(0)
    app = cedecrt.RestirDI(rt, args.width, args.height, tris, *cam, cedecrt.Options())
    rt.raycast(args.width, args.height, app.geom, app.triangles, app.raygen, app.visibility)
    mine = app.visibility.to_host()
    theirs = np.fromfile(vis_path, dtype=mine.dtype)
    diff = mine["index"] != theirs["index"]
    both = (~diff) & (mine["index"] >= 0)
    duv = np.abs(mine["uv"][both] - theirs["uv"][both]).max() if both.any() else 0.0
    res = {"pixels": int(diff.size), "primid_mismatches": int(diff.sum()),
           "mismatch_rate": float(diff.mean()), "max_abs_uv_diff_where_equal": float(duv),
           "sky_mine": int((mine["index"] < 0).sum()), "sky_hiprt": int((theirs["index"] < 0).sum())}
    # where do the uv differences on equal primitives come from?  Per pixel: |duv|, the hit triangle's shape (area over
    # squared longest edge: 0.43 for an equilateral triangle, -> 0 for a needle) and the position error the uv difference
    # amounts to in world units
    if both.any():
        idx = mine["index"][both]
        d = np.abs(mine["uv"][both] - theirs["uv"][both]).max(1)
        v = tris["vertices"][idx].astype(np.float64)
        e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
        area = 0.5 * np.linalg.norm(np.cross(e1, e2), axis=1)
        longest = np.maximum(np.maximum(np.linalg.norm(e1, axis=1), np.linalg.norm(e2, axis=1)), np.linalg.norm(v[:, 2] - v[:, 1], axis=1))
        shape = area / np.maximum(longest ** 2, 1e-30)
        duv2 = (mine["uv"][both] - theirs["uv"][both]).astype(np.float64)
        dpos = np.linalg.norm(duv2[:, :1] * e1 + duv2[:, 1:] * e2, axis=1)  # P = (1-u-v) v0 + u v1 + v v2
        q = [0.5, 0.9, 0.99, 0.999, 0.9999, 1.0]
        big = d > 1e-3
        res["uv_diff"] = {
            "quantiles": {str(x): float(np.quantile(d, x)) for x in q},
            "pixels_above_1e-3": int(big.sum()),
            "median_shape_all": float(np.median(shape)), "median_shape_above_1e-3": float(np.median(shape[big])) if big.any() else None,
            "median_longest_edge_above_1e-3": float(np.median(longest[big])) if big.any() else None,
            "position_error_quantiles": {str(x): float(np.quantile(dpos, x)) for x in q},
            "position_error_over_distance_max": float((dpos / np.maximum(np.linalg.norm(
                (1 - mine["uv"][both].sum(1))[:, None] * v[:, 0] + mine["uv"][both][:, :1] * v[:, 1] + mine["uv"][both][:, 1:] * v[:, 2]
                - np.asarray(cam[0]), axis=1), 1e-9)).max()),
            "note": "uv differences on equal primitive ids are where a small position difference (HIPRT computes the hit "
                    "from its own ray/triangle arithmetic) is divided by a small triangle: compare the position error"}
    rt.close()
    return res


def compare_radiance(args, theirs):
    """mean accumulated radiance of the reference GPU build against this repo's fused frame after the same number of frames
    (warm-up frames included: both accumulate from frame 1).  Per pixel the two differ by Monte-Carlo noise — HIPRT's
    intersector and its compiler's FMA contraction change individual random decisions — so the comparison is of the
    image means, whose standard error is far below 1e-3 at 8.3 M pixels."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "cedec-2024-rt_b200", "python"))
    import cedecrt
    import scenes

    tris = scenes.tile_scene(scenes.load_scene(args.scene), args.tiles[0], args.tiles[1], 130.0, 82.0)
    cam = ((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192))
    rt = cedecrt.Runtime(0)
    opt = cedecrt.Options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    app = cedecrt.RestirDI(rt, args.width, args.height, tris, *cam, opt, fused=True)
    n = args.frames + args.warmup
    for _ in range(n):
        app.frame()
    acc = app.accumulation.to_host().view(np.float32).reshape(-1, 4).astype(np.float64)
    mine = (acc[:, :3].sum(0) / acc[:, 3].sum()).tolist()
    ref = theirs["mean_radiance"]
    rel = [abs(a - b) / abs(b) for a, b in zip(mine, ref)]
    rt.close()
    return {"frames": n, "mean_radiance_mine": mine, "mean_radiance_hiprt": ref, "relative_difference": rel,
            "max_relative_difference": max(rel), "within_1e-3": max(rel) <= 1e-3}


if __name__ == "__main__":
    main()
