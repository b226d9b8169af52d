"""slabs — row-slab partition of a frame over the GPUs of one node and the halo exchange it needs.

The reference is single-GPU; this is new host logic.  Every kernel is per-pixel independent
except spatial_resampling, which reads the reservoirs and the sky/emissive class of up to 5 neighbours within
+-86.4 px (sample_2d_gaussian with rv0 >= 2^-23 and radius/1.96, examples/10_restir_di/10_restir_di.cu:309-313), so

  * rank r computes image rows yi in [y0, y1) (crt_set_row_range); buffers are full-size and indices global;
  * before every spatial pass each rank sends the HALO = 87 boundary rows of the pass's input reservoirs to its
    slab neighbours and receives theirs into the same global positions; once per frame the same happens for the
    neighbour-rejection data (Visibility rows in drop-in mode, the 1-byte pixel-class plane in fused mode).

Nothing here touches pixel data itself: the functions compute byte ranges of bottom-up buffers
(pixel_idx = xi + (H - yi - 1) * W, 10_restir_di.cu:18-20) and issue torch.distributed P2P ops on them, so the same
code runs over NCCL on GPU tensors (bench.py) and over gloo on CPU tensors (tests/test_slabs.py).
"""

HALO = 87  # rows; |neighbour offset| <= 86.4 px


def slab_rows(H, world, rank, align=8):
    """rows [y0, y1) of rank `rank`: equal shares rounded to multiples of `align` rows (the kernels' tile height)"""
    edges = [((H * r // world) + align - 1) // align * align for r in range(world)] + [H]
    edges[0] = 0
    edges = [min(e, H) for e in edges]
    return edges[rank], edges[rank + 1]


def weighted_slab_rows(row_cost, world, align=8):
    """slab edges that balance a per-row cost estimate (e.g. last frame's per-row ray counts) instead of row counts;
    returns the list of world + 1 edges"""
    H = len(row_cost)
    total = float(sum(row_cost))
    edges, acc, r = [0], 0.0, 1
    for y, c in enumerate(row_cost):
        acc += c
        while r < world and acc >= total * r / world:
            e = min(H, max(edges[-1], (y + 1 + align - 1) // align * align))
            edges.append(e)
            r += 1
    while len(edges) < world:
        edges.append(H)
    edges.append(H)
    return edges


def halo_plan(H, edges, rank, halo=HALO):
    """[(peer, send_rows, recv_rows)]: the rows of mine every other rank's spatial pass can read, and the rows of
    theirs mine can.  With slabs taller than `halo` this is the two adjacent slabs; thin slabs reach further."""
    y0, y1 = edges[rank], edges[rank + 1]
    plan = []
    for peer in range(len(edges) - 1):
        if peer == rank:
            continue
        p0, p1 = edges[peer], edges[peer + 1]
        if p1 <= p0 or y1 <= y0:
            continue
        # rows of mine within `halo` of the peer's slab, and vice versa
        send = (max(y0, p0 - halo), min(y1, p1 + halo))
        recv = (max(p0, y0 - halo), min(p1, y1 + halo))
        if send[0] < send[1] and recv[0] < recv[1]:
            plan.append((peer, send, recv))
    return plan


def row_byte_ranges(W, H, rows, layout):
    """byte ranges [(begin, end)] that image rows [a, b) occupy in a bottom-up buffer.
    layout: ("aos", elem_bytes) — one contiguous range; ("planes", [elem_bytes...]) — planar storage over W*H
    pixels, planes back to back in the order given, each plane padded to 16 * W * H bytes except the last ones
    as in csrc/restir_fast.cuh (plane p starts at sum of 16*W*H for the planes before it)."""
    a, b = rows
    first, last = (H - b) * W, (H - a) * W
    if layout[0] == "aos":
        e = layout[1]
        return [(first * e, last * e)]
    out, off = [], 0
    for e in layout[1]:
        out.append((off + first * e, off + last * e))
        off += 16 * W * H if e == 16 else e * W * H
    return out


SOA_RESERVOIR = ("planes", [16, 16, 16, 16, 8])  # csrc/restir_fast.cuh: SoaStore
AOS_RESERVOIR = ("aos", 76)
AOS_VISIBILITY = ("aos", 16)
CLASS_PLANE = ("aos", 1)


def exchange(dist, tensor, W, H, layout, plan):
    """send/receive the halo rows of one flat uint8 tensor (a full-size bottom-up buffer) according to `plan`;
    returns the bytes this rank sent"""
    if not plan:
        return 0
    ops, sent = [], 0
    for peer, send_rows, recv_rows in plan:
        for (s0, s1) in row_byte_ranges(W, H, send_rows, layout):
            ops.append(dist.P2POp(dist.isend, tensor[s0:s1], peer))
            sent += s1 - s0
        for (r0, r1) in row_byte_ranges(W, H, recv_rows, layout):
            ops.append(dist.P2POp(dist.irecv, tensor[r0:r1], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return sent
