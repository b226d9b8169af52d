"""Multi-GPU host logic on CPU: slab partition, halo plans, byte ranges of AoS / planar buffers, and a
world_size-2 and -3 gloo run of the slab-restricted frame loop against the single-process frame."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import orc
import slabs
from helpers import reservoir_mismatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_rows_cover_the_frame():
    for H in (1, 7, 8, 54, 1080, 2160):
        for world in (1, 2, 3, 4, 8):
            edges = [slabs.slab_rows(H, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == H
            for a, b in zip(edges, edges[1:]):
                assert a[1] == b[0] and a[0] <= a[1]
            assert all(e[0] % 8 == 0 or e[0] == H for e in edges)


def test_weighted_slabs_balance_cost():
    cost = np.ones(2160)
    cost[:540] = 0.1  # cheap sky rows on top
    edges = slabs.weighted_slab_rows(cost, 8)
    assert edges[0] == 0 and edges[-1] == 2160 and len(edges) == 9 and sorted(edges) == edges
    per = [cost[a:b].sum() for a, b in zip(edges, edges[1:])]
    assert max(per) / (sum(per) / 8) < 1.05
    assert edges[1] > 540  # the first slab swallows the cheap rows


def test_halo_plan_is_symmetric_and_sufficient():
    H = 2160
    for world in (2, 4, 8, 32):
        edges = [slabs.slab_rows(H, world, r)[0] for r in range(world)] + [H]
        plans = [slabs.halo_plan(H, edges, r) for r in range(world)]
        for r, plan in enumerate(plans):
            y0, y1 = edges[r], edges[r + 1]
            got = np.zeros(H, bool)
            got[y0:y1] = True
            for peer, send, recv in plan:
                back = [p for p in plans[peer] if p[0] == r]
                assert len(back) == 1 and back[0][1] == recv and back[0][2] == send  # my recv is their send
                got[recv[0]:recv[1]] = True
            need = np.zeros(H, bool)
            need[max(0, y0 - slabs.HALO):min(H, y1 + slabs.HALO)] = True
            assert (got | ~need).all()  # every row a neighbour lookup can touch is present


def test_halo_rows_bound_the_neighbour_reach():
    """slabs.halo_rows(radius) (and csrc/restir_fast.cuh: halo_rows_for, the same formula) bounds |neighbour - pixel|
    for every random pair the reference can draw: rv0 in {2^-23 ... 1 - 2^-23}, any angle (reservoir.hpp:89-95,
    10_restir_di.cu:309-313: float arithmetic, truncation towards zero)."""
    assert slabs.halo_rows(30.0) == slabs.HALO == 87
    rv0 = np.float32(2.0 ** -23)  # the smallest non-zero uniformf() value gives the longest offset
    phi = np.linspace(0, 2 * np.pi, 100001).astype(np.float32)
    for radius in (1.0, 7.5, 30.0, 64.0, 200.0):
        r = np.sqrt(np.maximum(np.float32(-2.0) * np.log(rv0), np.float32(0)), dtype=np.float32)
        gx = r * np.cos(phi, dtype=np.float32)
        for yi in (0, 1, 500, 2159):
            y = (np.float32(yi) + np.float32(radius) / np.float32(1.96) * gx).astype(np.float32)
            reach = np.abs(np.trunc(y).astype(np.int64) - yi).max()
            assert reach <= slabs.halo_rows(radius), (radius, yi, reach)  # towards smaller y the truncation rounds the reach up
        assert slabs.halo_rows(radius) <= int(radius / 1.96 * 5.65) + 2  # and it is tight


def test_halo_plan_with_a_wider_reach():
    H, world = 2160, 8
    edges = [slabs.slab_rows(H, world, r)[0] for r in range(world)] + [H]
    wide = slabs.halo_rows(120.0)  # 346 rows: beyond the adjacent slab (270 rows)
    plan = slabs.halo_plan(H, edges, 3, wide)
    assert sorted(p[0] for p in plan) == [1, 2, 4, 5]
    covered = sorted(r for _, _, recv in plan for r in range(*recv))
    assert covered[0] == edges[3] - wide and covered[-1] == edges[4] + wide - 1


def test_planar_row_ranges_select_exactly_the_rows(emu_lib):
    W, H = 16, 40
    n = W * H
    a = np.zeros(n, orc.RESERVOIR)
    rows_of_pixel = H - 1 - np.arange(n) // W  # bottom-up storage: yi of every pixel_idx
    for f in ("origin_position", "origin_normal", "hit_position", "hit_normal", "radiance"):
        a[f] = (rows_of_pixel[:, None] + 1).astype(np.float32)
    a["w_sum"] = a["ucw"] = rows_of_pixel + 1
    a["M"] = rows_of_pixel + 1
    soa = np.zeros(n * 76, np.uint8)
    emu_lib.emu_aos_to_soa(a.ctypes.data_as(C.c_void_p), soa.ctypes.data_as(C.c_void_p), C.c_long(n))
    rows = (13, 29)
    part = np.zeros_like(soa)
    for b, e in slabs.row_byte_ranges(W, H, rows, slabs.SOA_RESERVOIR):
        part[b:e] = soa[b:e]
    back = np.zeros(n, orc.RESERVOIR)
    emu_lib.emu_soa_to_aos(part.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p), C.c_long(n))
    inside = (rows_of_pixel >= rows[0]) & (rows_of_pixel < rows[1])
    assert reservoir_mismatch(back[inside], a[inside]) == 0
    assert (back["M"][~inside] == 0).all() and (back["w_sum"][~inside] == 0).all()
    # AoS: one contiguous range
    (b, e), = slabs.row_byte_ranges(W, H, rows, slabs.AOS_RESERVOIR)
    assert (e - b) == (rows[1] - rows[0]) * W * 76 and b == (H - rows[1]) * W * 76


@pytest.fixture(scope="module")
def emu_lib():
    so = os.path.join(ROOT, "tests", "emu", "libemu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    inc = os.path.join(ROOT, "cedec-2024-rt_b200", "csrc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run([gxx, "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                        "-I" + inc, "-o", so, src], check=True)
    return C.CDLL(so)


def test_full_plan_moves_every_row_to_every_rank():
    H = 1080
    for world in (2, 3, 8):
        edges = [slabs.slab_rows(H, world, r)[0] for r in range(world)] + [H]
        for r in range(world):
            plan = slabs.full_plan(H, edges, r)
            assert sorted(p[0] for p in plan) == [q for q in range(world) if q != r]
            for peer, send, recv in plan:
                assert send == (edges[r], edges[r + 1]) and recv == (edges[peer], edges[peer + 1])


@pytest.mark.parametrize("world,height,reproject", [(2, 200, 0), (3, 200, 0), (2, 200, 1), (3, 120, 1)])
def test_slab_frames_equal_single_process_frames_gloo(world, height, reproject, port):
    """world 2: slabs taller than the halo (the 8-GPU 4K case); world 3 at H = 200: slabs thinner than the halo, so
    rows travel between non-adjacent ranks too; reproject: a camera that moves twice, temporal resampling with the
    history looked up at the reprojected pixel, every rank gathering every slab's history rows first"""
    env = dict(os.environ, SLAB_H=str(height), OMP_NUM_THREADS="2", SLAB_REPROJECT=str(reproject))
    port_no = 29500 + (os.getpid() % 1000) + world + 7 * reproject
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port_no),
                        os.path.join(ROOT, "tests", "slab_worker.py")], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "SLABS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
