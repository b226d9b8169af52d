"""TEST INFRASTRUCTURE — one rank of the multi-GPU row-slab test (tests/test_gpu_parity.py::test_slabs_on_gpus,
torch.distributed.run, NCCL).  Every rank renders its slab of the frame with halo exchange (slabs.SlabRenderer, the
object bench.py times) and, on the same GPU, the whole frame as a single slab; its rows must be identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("oracle", "tests", os.path.join("cedec-2024-rt_b200", "python")):
    sys.path.insert(0, os.path.join(ROOT, p))
import slabs  # noqa: E402
import stage_assets  # noqa: E402  (scene cache reader; test infrastructure)
from helpers import small_scene  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    # fused (direct peer stores) | fused_nccl | dropin (NCCL) | group | fused_reproject | dropin_reproject (a camera that moves
    # twice, temporal resampling with the history looked up at the reprojected pixel: every rank gathers every slab's history)
    mode = os.environ.get("SLAB_MODE", "fused")
    reproject = mode.endswith("_reproject")
    fused = not mode.startswith("dropin")
    if stage_assets.have_scene("blocks_restir"):
        tris = stage_assets.load_scene("blocks_restir")
        cam, W, H = ((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192)), 960, 540
        if mode == "group":
            H = 1080  # two slabs per GPU, each at least a halo tall
    else:
        tris = small_scene("blocks_ao").copy()
        tris["emissive"][100:140] = (5.0, 4.0, 3.0)
        cam, W, H = ((8.0, 8.0, 8.0), (0.0, 0.0, 0.0)), 320, 400
    if mode == "group":  # two slabs per GPU on two streams (slabs.SlabGroup), linked by pointers within a process
        part = slabs.SlabGroup(torch, dist, rank, world, tris, cam, W, H, sub=2)
        part.calibrate(rounds=1, frames=1)  # uneven slabs, and a history reset in between
    else:
        part = slabs.SlabRenderer(torch, dist, rank, world, tris, cam, W, H, fused=fused, p2p=(mode == "fused"), reproject=reproject)
    full = slabs.SlabRenderer(torch, None, 0, 1, tris, cam, W, H, fused=fused, reproject=reproject)
    e, c = np.asarray(cam[0], np.float64), np.asarray(cam[1], np.float64)
    moves = {1: (tuple(e + (0.4, -0.2, 0.1)), tuple(c + (0.1, 0.0, -0.1))),  # after frame 1: most surface points stay on screen
             2: (tuple(e + (0.9, -1.0, 0.3)), tuple(c + (0.2, 0.6, -0.2)))} if reproject else {}  # a larger, mostly vertical move
    for k in range(3):
        part.frame()
        full.frame()
        if k + 1 in moves:
            part.set_camera(*moves[k + 1])
            full.set_camera(*moves[k + 1])
    if reproject:
        part.frame()
        full.frame()
    torch.cuda.synchronize()
    ok = True
    for sl in part.slabs:
        for name, elem in (("t_acc", 16), ("t_pix", 4), ("t_vis", 16)):
            a = sl._rows(getattr(sl, name), elem, sl.y0, sl.y1)
            b = full._rows(getattr(full, name), elem, sl.y0, sl.y1)
            ok = ok and bool(torch.equal(a, b))
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("GPU_SLABS_OK" if all(flags) else "GPU_SLABS_MISMATCH %s" % flags, "world", world, "mode", mode,
              "p2p", part.p2p, "edges", part.edges, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
