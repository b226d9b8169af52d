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
    mode = os.environ.get("SLAB_MODE", "fused")  # fused (direct peer stores) | fused_nccl | dropin (NCCL) | group
    fused = mode != "dropin"
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
        part = slabs.SlabRenderer(torch, dist, rank, world, tris, cam, W, H, fused=fused, p2p=(mode == "fused"))
    full = slabs.SlabRenderer(torch, None, 0, 1, tris, cam, W, H, fused=fused)
    for _ in range(3):
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
