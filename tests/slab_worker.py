"""TEST INFRASTRUCTURE — one rank of the multi-process row-slab test (launched by tests/test_slabs.py through
torch.distributed.run, gloo backend, CPU).  Each rank runs the ReSTIR DI frame loop of the CPU oracle restricted to
its slab (orc.set_range), exchanging halo rows with cedec-2024-rt_b200/python/slabs.py exactly as bench.py does on
GPUs, and checks its rows against the full single-process frame.  SLAB_REPROJECT=1: the camera moves twice and temporal
resampling looks its history up at the reprojected pixel (oracle: reproject_pixel) — before such a frame every rank
gathers every other rank's rows of the history (slabs.full_plan), because the look-up is bounded by the camera motion,
not by the spatial halo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", os.path.join("cedec-2024-rt_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import orc  # noqa: E402
import slabs  # noqa: E402
from helpers import reservoir_mismatch, same, small_scene  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, frames = 48, int(os.environ.get("SLAB_H", "200")), 3
    reproject = os.environ.get("SLAB_REPROJECT", "0") == "1"
    moves = {2: ((8.4, 7.8, 8.1), (0.1, 0.0, -0.1)), 3: ((8.9, 7.0, 8.3), (0.2, 0.6, -0.2))} if reproject else {}  # before frame f
    tris = small_scene("blocks_ao").copy()
    tris["emissive"][100:140] = (5.0, 4.0, 3.0)
    cam = ((8.0, 8.0, 8.0), (0.0, 0.0, 0.0))
    opt = orc.make_options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    o = orc.load("port")
    o.set_threads(2)
    g = o.geom_build(tris)

    full = orc.RestirChain(o, W, H, tris, g, *cam, opt, reproject=reproject)
    for f in range(1, frames + 1):
        if f in moves:
            full.set_camera(*moves[f])
        full.step()

    edges = [slabs.slab_rows(H, world, r)[0] for r in range(world)] + [H]
    y0, y1 = edges[rank], edges[rank + 1]
    plan = slabs.halo_plan(H, edges, rank)
    everything = slabs.full_plan(H, edges, rank)
    mine = orc.RestirChain(o, W, H, tris, g, *cam, opt, reproject=reproject)
    prev_rg = None
    # poison everything outside the slab so that a missing halo row cannot go unnoticed
    t_of = {}

    def tensor(arr):
        return torch.from_numpy(arr.view(np.uint8).reshape(-1))

    o.set_range(y0 * W, y1 * W)  # tid = yi * W + xi
    for f in range(1, frames + 1):
        if f in moves:
            o.set_range(0, -1)
            mine.set_camera(*moves[f])  # clears the accumulation like the reference's host (10_restir_di.cpp:257-267)
            o.set_range(y0 * W, y1 * W)
        eye = mine.eye
        for name in ("vis", "buf0", "buf1") + (("temporal",) if reproject else ()):
            a = getattr(mine, name).view(np.uint8).reshape(H, -1)  # bottom-up rows
            keep = a[H - y1:H - y0].copy()
            a[:] = 0xEE
            a[H - y1:H - y0] = keep
        o.raycast(W, H, g, tris, mine.rg, mine.vis)
        slabs.exchange(dist, tensor(mine.vis), W, H, slabs.AOS_VISIBILITY, plan)
        o.generate_candidate(W, H, f, g, tris, mine.vis, eye, mine.lights, opt, mine.buf0)
        if reproject and prev_rg is not None:
            # the history of every slab: poisoned above outside my rows, so a row that did not travel cannot go unnoticed
            slabs.exchange(dist, tensor(mine.temporal), W, H, slabs.AOS_RESERVOIR, everything)
            o.temporal_resampling_reprojected(W, H, f, g, tris, mine.vis, eye, opt, prev_rg, mine.temporal, mine.buf0)
        else:
            o.temporal_resampling(W, H, f, g, tris, mine.vis, eye, opt, mine.temporal, mine.buf0)
        prev_rg = mine.rg
        o.save_temporal_reservoir(W, H, mine.buf0, mine.temporal)
        bi, bo = mine.buf0, mine.buf1
        for k in range(int(opt["spatial_resampling_passes"])):
            if k:
                bi, bo = bo, bi
            slabs.exchange(dist, tensor(bi), W, H, slabs.AOS_RESERVOIR, plan)
            o.spatial_resampling(W, H, f, k, g, tris, mine.vis, eye, opt, bi, bo)
        o.resolve(mine.accum, W, H, g, tris, mine.vis, eye, opt, bo)
    o.set_range(0, -1)

    rows = slice((H - y1) * W, (H - y0) * W)
    ok = same(mine.accum[rows], full.accum[rows]) and same(mine.vis[rows], full.vis[rows])
    ok = ok and reservoir_mismatch(bo[rows], full.out[rows]) == 0
    ok = ok and reservoir_mismatch(mine.temporal[rows], full.temporal[rows]) == 0
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print("SLABS_OK" if all(flags) else "SLABS_MISMATCH %s" % flags, "world", world, "edges", edges, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
