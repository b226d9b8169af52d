#!/usr/bin/env python
"""MEASUREMENT TOOL (one GPU): where do small slabs lose?  Renders the 4K config-5 frame as N row slabs one after the
other on the same GPU (no neighbours, no halo traffic, no waiting) and compares the per-kernel device times, summed over
the slabs, with the whole frame's — the part of the multi-GPU loss that is launch size alone.
usage: python profiles/slab_probe.py [N ...]   -> one JSON line per N"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cedec-2024-rt_b200", "python"))
import torch  # noqa: E402

import bench  # noqa: E402
import slabs  # noqa: E402


def kernel_ms(r, frames=3):
    for _ in range(3):
        r.frame()
    r.rt.profile_begin()
    for _ in range(frames):
        r.frame()
    out = {}
    for name, ms in r.rt.profile_end():
        out[name] = out.get(name, 0.0) + ms / frames
    return out


def main():
    ns = [int(a) for a in sys.argv[1:]] or [8]
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    tris, cam, _ = bench.load_workload()
    W, H = 3840, 2160
    full = slabs.SlabRenderer(torch, None, 0, 1, tris, cam, W, H, fused=True)
    base = kernel_ms(full)
    print(json.dumps({"slabs": 1, "kernels_ms": {k: round(v, 4) for k, v in base.items()}, "sum_ms": round(sum(base.values()), 4)}))
    for n in ns:
        total = {}
        per_slab = []
        for r in range(n):
            y0, y1 = slabs.slab_rows(H, n, r)
            s = slabs.SlabRenderer(torch, None, 0, 1, tris, cam, W, H, fused=True, edges=[y0, y1], share=full)
            k = kernel_ms(s)
            per_slab.append(round(sum(k.values()), 4))
            for name, ms in k.items():
                total[name] = total.get(name, 0.0) + ms
            del s
        print(json.dumps({"slabs": n, "kernels_ms_summed_over_slabs": {k: round(v, 4) for k, v in total.items()},
                          "sum_ms": round(sum(total.values()), 4), "per_slab_ms": per_slab,
                          "ratio_to_whole_frame": {k: round(total[k] / base[k], 3) for k in total if k in base},
                          "ratio_sum": round(sum(total.values()) / sum(base.values()), 3)}))


if __name__ == "__main__":
    main()
