#!/usr/bin/env python
"""SURVEY.md section 8(f) row 1 — BVH build time, builder quality trade-off and refit on the config-5 scene
(blocks_restir x6, 9 590 208 triangles): for each builder (PLOC, the default; LBVH) the build time, the tree size,
the 4K primary-ray time it gives (k_raycast, 3840x2160, config-5 camera) and the refit time after every vertex
moved.  Runs on the GPU box:  python profiles/bvh_build_bench.py > gpurun_out/bvh_build.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cedec-2024-rt_b200", "python"))
import cedecrt  # noqa: E402
import scenes  # noqa: E402

CAM = ((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192))
W, H = 3840, 2160


def main():
    rt = cedecrt.Runtime(0)
    tris = scenes.tile_scene(scenes.load_scene("blocks_restir"), 3, 2, 130.0, 82.0)
    d_tris = rt.to_device(tris)
    vis = rt.buffer(cedecrt.VISIBILITY, W * H)
    rg = cedecrt.lookat(CAM[0], CAM[1], W, H)
    out = {"triangles": int(len(tris)), "device": rt.device_name(), "builders": {}}
    for builder in ("ploc", "lbvh"):
        os.environ["CRT_BVH_BUILDER"] = builder
        builds = []
        for _ in range(3):
            g = rt.build_geometry(d_tris)
            builds.append(g.stats()["build_ms"])
            if _ < 2:
                g.destroy()
        st = g.stats()
        for _ in range(2):
            rt.raycast(W, H, g, d_tris, rg, vis)
        rt.profile_begin()
        for _ in range(5):
            rt.raycast(W, H, g, d_tris, rg, vis)
        ray_ms = [ms for name, ms in rt.profile_end() if name == "raycast"]
        ids_before = vis.to_host()["index"].copy()
        # every vertex moves (a rigid shift keeps the image), then refit; the image must be the same
        moved = tris.copy()
        moved["vertices"] += np.float32(0.25)
        d_tris.upload(moved)
        refits = []
        for _ in range(3):
            g.refit()
            refits.append(g.stats()["refit_ms"])
        eye = tuple(np.float32(c) + np.float32(0.25) for c in CAM[0])
        ctr = tuple(np.float32(c) + np.float32(0.25) for c in CAM[1])
        rt.raycast(W, H, g, d_tris, cedecrt.lookat(eye, ctr, W, H), vis)
        same_ids = float((vis.to_host()["index"] == ids_before).mean())
        d_tris.upload(tris)
        out["builders"][builder] = {
            "build_ms": [round(x, 2) for x in builds], "nodes": int(st["n_nodes"]), "depth": int(st["max_depth"]),
            "node_mb": round(st["node_bytes"] / 1e6, 1), "raycast_4k_ms": round(float(np.median(ray_ms)), 3),
            "primary_grays_per_s": round(W * H / float(np.median(ray_ms)) / 1e6, 3),
            "refit_ms": [round(x, 2) for x in refits], "ids_equal_after_shift_and_refit": round(same_ids, 6)}
        g.destroy()
    os.environ.pop("CRT_BVH_BUILDER", None)
    print(json.dumps(out))
    rt.close()


if __name__ == "__main__":
    main()
