#!/usr/bin/env python
"""How much would overlapping independent kernel sequences on one GPU recover at slab size?  Two renderers (two
contexts, two streams) each render a 3840 x H frame of config 5's scene; K frames each, issued alternately, are timed
against the same 2K frames issued on one stream.  The gain is the upper bound of what pipelining a frame's tail under the
next frame's head could give a slab at 8 GPUs.  usage: overlap_probe.py [H]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cedec-2024-rt_b200", "python"))
import bench  # noqa: E402
from slabs import SlabRenderer  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 272
K = 64
tris, cam, _ = bench.load_workload()
rs = []
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for s in streams:
    with torch.cuda.stream(s):
        rs.append(SlabRenderer(torch, None, 0, 1, tris, cam, 3840, H, fused=True))


def run(two_streams):
    for r, s in zip(rs, streams):
        r.rt.set_stream((s if two_streams else streams[0]).cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for _ in range(K):
        for r in rs:
            r.frame()
    if two_streams:
        done = torch.cuda.Event()
        done.record(streams[1])
        streams[0].wait_event(done)
    e1.record(streams[0])
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (2 * K)


for r in rs:
    for _ in range(4):
        r.frame()
out = {"height": H, "frames_each": K}
out["ms_per_frame_one_stream"] = round(run(False), 4)
out["ms_per_frame_two_streams"] = round(run(True), 4)
out["gain"] = round(out["ms_per_frame_one_stream"] / out["ms_per_frame_two_streams"], 3)
print(json.dumps(out))
