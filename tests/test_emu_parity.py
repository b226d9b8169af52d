"""Logic of the CUDA path, checked without a GPU: tests/emu compiles the device functions of
cedec-2024-rt_b200/csrc/*.cuh (BVH build steps, wide-BVH walk, every per-pixel body) for the host and these
tests compare them bit for bit with the oracle (canonical left-to-right argument order, both math modes).
The real parity tests — the CUDA kernels through the C ABI — are in test_gpu_parity.py (-m gpu)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import orc
from helpers import reservoir_mismatch, same, small_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CAM_CB = ((0.0, 2.7, 9.0), (0.0, 2.7, 0.0))
CAM_AO = ((8.0, 8.0, 8.0), (0.0, 0.0, 0.0))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, "tests", "emu", "libemu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    inc = os.path.join(ROOT, "cedec-2024-rt_b200", "csrc")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                    "-I" + inc, "-o", so, src], check=True)
    return orc.Oracle(so)


def lit_blocks_ao():
    t = small_scene("blocks_ao").copy()
    t["emissive"][100:140] = (5.0, 4.0, 3.0)
    return t


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("scene", ["cornellbox1", "blocks_ao_lit"])
def test_restir_chain(emu, port, scene, mode):
    tris = small_scene("cornellbox1") if scene == "cornellbox1" else lit_blocks_ao()
    cam, (W, H) = (CAM_CB, (96, 54)) if scene == "cornellbox1" else (CAM_AO, (128, 72))
    opt = orc.make_options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    outs = []
    for o in (port, emu):
        o.set_math_mode(mode)
        g = o.geom_build(tris)
        ch = orc.RestirChain(o, W, H, tris, g, *cam, opt)
        for _ in range(3):
            ch.step()
        outs.append((ch.vis.copy(), ch.buf0.copy(), ch.buf1.copy(), ch.temporal.copy(), ch.accum.copy(),
                     o.tone_mapping(ch.accum, W, H).copy()))
        o.geom_free(g)
        o.set_math_mode(0)
    a, b = outs
    assert same(a[0]["index"], b[0]["index"]) and same(a[0]["uv"], b[0]["uv"])
    for k in (1, 2, 3):
        assert reservoir_mismatch(a[k], b[k]) == 0
    assert same(a[4], b[4]) and same(a[5], b[5].view(np.uint8))
    assert float(a[4][:, :3].sum()) > 0


def test_option_variants(emu, port):
    tris = small_scene("cornellbox1")
    W, H = 64, 36
    for kw in (dict(use_shadowed_target_function=1, use_visibility_reuse=0, use_temporal_resampling=1,
                    use_spatial_resampling=1, ris_sample_count=8, spatial_resampling_passes=2),
               dict(use_spatial_resampling=0, use_temporal_resampling=1),
               dict(use_spatial_resampling=1, use_temporal_resampling=0, spatial_resampling_radius=5.0,
                    spatial_resampling_sample_count=2, accumulate=1)):
        opt = orc.make_options(**kw)
        outs = []
        for o in (port, emu):
            g = o.geom_build(tris)
            ch = orc.RestirChain(o, W, H, tris, g, *CAM_CB, opt)
            ch.step()
            ch.step()
            outs.append((ch.buf0.copy(), ch.buf1.copy(), ch.accum.copy()))
            o.geom_free(g)
        assert reservoir_mismatch(outs[0][0], outs[1][0]) == 0 and reservoir_mismatch(outs[0][1], outs[1][1]) == 0
        assert same(outs[0][2], outs[1][2]), kw


@pytest.mark.parametrize("ex", [7, 8, 9, 109])
def test_path_tracers(emu, port, ex):
    """109: example 09 with use_shadowed_target_function (BASELINE config 4): a shadow ray inside every candidate's target
    function — the CUDA path walks those in a per-lane loop of walk steps (restir_core.cuh: ris_candidates_shadowed)"""
    tris = small_scene("cornellbox1")
    W, H = 64, 36
    opt = orc.make_options(accumulate=1, max_depth=4, ris_sample_count=8, sky_color=(0.1, 0.2, 0.3),
                           use_shadowed_target_function=1 if ex == 109 else 0)
    ex = 9 if ex == 109 else ex
    lights = orc.light_indices(tris)
    outs = []
    for o in (port, emu):
        o.set_example(ex)
        g = o.geom_build(tris)
        acc = np.zeros((W * H, 4), np.float32)
        rg = o.lookat(*CAM_CB, W, H)
        for frame in (1, 2):
            o.path_trace(W, H, frame, g, tris, lights, rg, opt, acc)
        outs.append(acc)
        o.geom_free(g)
    port.set_example(9)
    assert same(outs[0], outs[1]) and float(outs[0][:, :3].sum()) > 0


def test_ao(emu, port):
    tris = small_scene("blocks_ao")
    W, H = 96, 54
    outs = []
    for o in (port, emu):
        o.set_example(6)
        g = o.geom_build(tris)
        outs.append(o.ao(W, H, g, tris, o.lookat(*CAM_AO, W, H), 16).view(np.uint8).copy())
        o.geom_free(g)
    port.set_example(9)
    assert same(outs[0], outs[1]) and len(np.unique(outs[0])) > 4


def test_wide_bvh_equals_brute_force(emu, port):
    """closest hit (incl. the tie rule) and any-hit of the compressed wide BVH vs the exhaustive loop"""
    tris = small_scene("blocks_ao")
    g = emu.geom_build(tris)
    lib = port.lib
    lib.orc_closest_hit_brute.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    emu.lib.emu_any_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
    emu.lib.emu_any_hit_far_first.argtypes = emu.lib.emu_any_hit.argtypes
    rng = np.random.default_rng(7)
    hits = 0
    for i in range(600):
        o = rng.uniform(-6, 6, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if i % 5 == 0:
            d[rng.integers(3)] = 0.0  # axis-parallel component
        if i % 7 == 0:
            o = np.round(o)  # origins on grid planes where many block faces lie
        tmax = 3.402823466e38 if i % 3 else float(rng.uniform(0.5, 8))
        idx, tuv = emu.closest_hit(g, o, d, 0.0, tmax)
        tuv2 = np.zeros(3, np.float32)
        idx2 = lib.orc_closest_hit_brute(tris.ctypes.data, len(tris), o.ctypes.data, d.ctypes.data, 0.0, tmax,
                                         tuv2.ctypes.data)
        assert idx == idx2 and (idx < 0 or same(tuv, tuv2)), (i, o, d)
        assert emu.lib.emu_any_hit(g, o.ctypes.data, d.ctypes.data, 0.0, tmax) == (1 if idx2 >= 0 else 0)
        # far end first (shadow rays towards lights): another visiting order, the same answer
        assert emu.lib.emu_any_hit_far_first(g, o.ctypes.data, d.ctypes.data, 0.0, tmax) == (1 if idx2 >= 0 else 0)
        # the same walk with triangle groups postponed whenever the walk allows it (the divergence control of
        # the CUDA path; any visiting order must give the same closest hit)
        emu.lib.emu_set_postpone(1)
        idx3, tuv3 = emu.closest_hit(g, o, d, 0.0, tmax)
        any3 = emu.lib.emu_any_hit(g, o.ctypes.data, d.ctypes.data, 0.0, tmax)
        emu.lib.emu_set_postpone(0)
        assert idx3 == idx2 and (idx3 < 0 or same(tuv3, tuv2)) and any3 == (1 if idx2 >= 0 else 0)
        hits += idx >= 0
    assert hits > 100
    emu.geom_free(g)


def test_degenerate_inputs(emu, port):
    """empty scene, single triangle, 3 triangles (one leaf), duplicated triangles (identical Morton keys)"""
    W, H = 32, 18
    cb = small_scene("cornellbox1")
    for tris in (cb[:0], cb[:1], cb[:3], np.concatenate([cb[:5]] * 4)):
        tris = np.ascontiguousarray(tris)
        outs = []
        for o in (port, emu):
            g = o.geom_build(tris)
            outs.append(o.raycast(W, H, g, tris, o.lookat(*CAM_CB, W, H)))
            o.geom_free(g)
        assert same(outs[0]["index"], outs[1]["index"]) and same(outs[0]["uv"], outs[1]["uv"])


def test_refit_after_moving_vertices(emu, port):
    """crt_refit_geometry's device functions (bvh_build.cuh: refit_tri / refit_node): after vertices move, the
    refitted tree gives the closest hits of a tree built from scratch — the oracle's — bit for bit"""
    tris = small_scene("blocks_ao").copy()
    W, H = 128, 72
    ge = emu.geom_build(tris)
    rng = np.random.default_rng(3)
    for step in range(2):
        # whole blocks slide, single vertices jitter, and the scene box grows
        tris["vertices"][300:900] += rng.uniform(-1.5, 1.5, 3).astype(np.float32)
        tris["vertices"][1200:1300] += rng.uniform(-0.05, 0.05, (100, 3, 3)).astype(np.float32)
        tris["vertices"][2000:2012] += np.float32(25.0 * (step + 1))
        emu.lib.emu_geom_refit(ge)
        g = port.geom_build(tris)
        rg = port.lookat(*CAM_AO, W, H)
        a, b = port.raycast(W, H, g, tris, rg), emu.raycast(W, H, ge, tris, rg)
        assert same(a["index"], b["index"]) and same(a["uv"], b["uv"]), step
        assert (a["index"] >= 0).sum() > 1000
        port.geom_free(g)
    emu.geom_free(ge)


def test_hinted_raycast_equals_the_plain_walk_for_any_hint(emu, port):
    """px_raycast_hinted (the walk seeded with the triangle the pixel's record names before the call): the stored record is
    the oracle's whatever the hint — the right triangle, a neighbour's, a random one, -1, garbage — also when the camera
    moves between the frame that wrote the hints and the frame that reads them, and with duplicated triangles (ties on t
    go to the larger id: a hint with the smaller id of a duplicate pair must not survive)"""
    tris = small_scene("blocks_ao").copy()
    tris = np.concatenate([tris, tris[500:700]])  # duplicates: equal t, larger id wins
    W, H = 128, 72
    ge = emu.geom_build(tris)
    g = port.geom_build(tris)
    L = emu.lib
    L.emu_raycast_hinted.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11)
    rg0 = port.lookat(*CAM_AO, W, H)
    want0 = port.raycast(W, H, g, tris, rg0)
    rg1 = port.lookat((7.0, 9.0, 8.5), (0.5, 0.0, -0.5), W, H)
    want1 = port.raycast(W, H, g, tris, rg1)
    assert (want0["index"] >= 0).sum() > 1000 and not same(want0["index"], want1["index"])
    assert (want0["index"] >= len(tris) - 200).sum() > 0  # some pixels see a duplicated triangle (and report the larger id)

    def run(rg, hints):
        vis = np.zeros(W * H, dtype=want0.dtype)
        vis["index"] = hints
        vis["uv"] = rng.uniform(-5, 5, (W * H, 2)).astype(np.float32)
        L.emu_raycast_hinted(W, H, ge, tris.ctypes.data, len(tris), rg.ctypes.data, vis.ctypes.data)
        return vis

    n = len(tris)
    hint_sets = {
        "own answer": want0["index"].copy(),
        "other camera's answer": want1["index"].copy(),
        "shifted by one pixel": np.roll(want0["index"], 1),
        "random triangles": rng.integers(0, n, W * H).astype(np.int32),
        "none": np.full(W * H, -1, np.int32),
        "garbage": rng.integers(-2**31, 2**31 - 1, W * H).astype(np.int32),
        "the duplicate with the smaller id": np.where(want0["index"] >= n - 200, want0["index"] - (n - 200) + 500, want0["index"]).astype(np.int32),
    }
    for name, hints in hint_sets.items():
        for rg, want in ((rg0, want0), (rg1, want1)):
            got = run(rg, hints)
            assert same(got["index"], want["index"]) and same(got["uv"], want["uv"]), name
    port.geom_free(g)
    emu.geom_free(ge)


def test_seeded_walk_on_triangle_soups(emu, port):
    """bvh.cuh: trace_seeded against the oracle's closest hit on random triangle soups with duplicated, degenerate (zero
    area, repeated vertex) and coplanar overlapping triangles, random rays (a third of them aimed at a triangle's
    centroid so that ties on t between duplicates occur), every kind of seed: the true answer, its duplicate with the
    smaller id, a triangle behind the hit, a random one, out of range — both postponing settings of the host walk"""
    rng = np.random.default_rng(23)
    L = emu.lib
    L.emu_closest_hit_seeded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    base = small_scene("cornellbox1")
    for n_tri in (1, 7, 200, 1500):
        tris = np.zeros(n_tri, base.dtype)
        c = rng.uniform(-4, 4, (n_tri, 1, 3)).astype(np.float32)
        tris["vertices"] = c + rng.normal(0, 0.6, (n_tri, 3, 3)).astype(np.float32)
        k = max(1, n_tri // 10)
        tris["vertices"][-k:] = tris["vertices"][:k]  # duplicates: the larger id must win
        if n_tri > 20:
            tris["vertices"][10] = tris["vertices"][10][[0, 0, 1]]  # repeated vertex
            tris["vertices"][11] = tris["vertices"][11][0] + np.outer([0.0, 1.0, 2.0], [1.0, 0.5, 0.25]).astype(np.float32)  # collinear
            tris["vertices"][12] = tris["vertices"][13] * np.float32(0.5) + tris["vertices"][13].mean(0) * np.float32(0.5)  # coplanar, inside 13
        tris = np.ascontiguousarray(tris)
        g, ge = port.geom_build(tris), emu.geom_build(tris)
        n_rays = 600
        o = rng.uniform(-8, 8, (n_rays, 3)).astype(np.float32)
        d = rng.normal(0, 1, (n_rays, 3)).astype(np.float32)
        aim = rng.integers(0, n_tri, n_rays)
        cen = tris["vertices"][aim].mean(1)
        d[::3] = (cen - o)[::3]
        for post in (0, 1):
            L.emu_set_postpone(post)
            for i in range(n_rays):
                want, tuv = port.closest_hit(g, o[i], d[i])
                seeds = [want, -1, int(rng.integers(0, n_tri)), int(aim[i]), n_tri + 5, -7]
                if want >= n_tri - k:
                    seeds.append(want - (n_tri - k))  # the duplicate with the smaller id
                for hint in seeds:
                    got_tuv = np.zeros(3, np.float32)
                    got = L.emu_closest_hit_seeded(ge, tris.ctypes.data, n_tri, o[i].ctypes.data, d[i].ctypes.data, int(hint),
                                                   got_tuv.ctypes.data)
                    assert got == want and (want < 0 or same(got_tuv, tuv)), (n_tri, i, hint, got, want)
        L.emu_set_postpone(0)
        port.geom_free(g)
        emu.geom_free(ge)


# ---------------------------------------------------------------------------------------------- fused frame
class EmuFusedFrame:
    """Drives emu_restir_frame_fast (the kernel sequence of crt_restir_di_frame, csrc/kernels_fast.cu) on planar
    SoA reservoir storage and converts back to the reference's AoS for comparison."""

    def __init__(self, emu, W, H, tris, g, eye, center, opt):
        self.L, self.W, self.H, self.tris, self.g, self.opt = emu.lib, W, H, tris, g, opt
        self.eye = np.asarray(eye, np.float32)
        self.rg = emu.lookat(eye, center, W, H)
        self.lights = orc.light_indices(tris)
        n = W * H
        self.vis = np.zeros(n, orc.VISIBILITY)
        self.T, self.A, self.B = (np.zeros(n * 76, np.uint8) for _ in range(3))
        self.accum = np.zeros((n, 4), np.float32)
        self.pixels = np.zeros(n, np.uint32)
        self.frame = 0
        self.emu, self.reproject, self.prev_rg = emu, False, None

    def set_camera(self, eye, center):
        """the reference's loop on a camera move: new ray generator, accumulation cleared (10_restir_di.cpp:257-267)"""
        self.eye = np.asarray(eye, np.float32)
        self.rg = self.emu.lookat(eye, center, self.W, self.H)
        self.accum[:] = 0

    def step(self):
        self.frame += 1
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lights = self.lights if len(self.lights) else np.zeros(1, np.uint32)
        prev = p(self.prev_rg) if self.reproject and self.prev_rg is not None else None
        self.L.emu_restir_frame_fast_reprojected(self.W, self.H, self.frame, self.g, p(self.tris), p(self.rg), p(self.eye),
                                                 p(lights), len(self.lights), p(self.opt), p(self.vis), p(self.T), p(self.A),
                                                 p(self.B), p(self.accum), p(self.pixels), prev)
        self.prev_rg = self.rg.copy()

    def aos(self, soa):
        out = np.zeros(self.W * self.H, orc.RESERVOIR)
        self.L.emu_soa_to_aos(soa.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), C.c_long(len(out)))
        return out

    def output(self):
        passes = int(self.opt["spatial_resampling_passes"])
        if not self.opt["use_spatial_resampling"] or passes == 0:
            return self.aos(self.T)
        return self.aos(self.A if passes % 2 else self.B)


def diffuse_mask(vis, tris):
    em = (tris["emissive"] > 0).any(1)
    m = vis["index"] >= 0
    m[m] = ~em[vis["index"][m]]
    return m


FUSED_VARIANTS = [
    dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1),
    dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1, spatial_resampling_passes=2),
    dict(accumulate=0, use_temporal_resampling=1, use_spatial_resampling=1, spatial_resampling_passes=1),
    dict(accumulate=1, use_temporal_resampling=0, use_spatial_resampling=1),
    dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=0),
    dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1, use_visibility_reuse=0),
    dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1, ris_sample_count=7,
         spatial_resampling_sample_count=3, spatial_resampling_radius=10.0),
]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("variant", range(len(FUSED_VARIANTS)))
def test_fused_frame_equals_kernel_chain(emu, port, variant, mode):
    """the fused data flow (SoA planes, candidate+temporal in one body, in-place temporal buffer, conditional
    visibility-reuse ray, G-buffer, tone mapping in the resolve epilogue) gives the oracle's buffers bit for bit"""
    tris = lit_blocks_ao()
    W, H = 96, 54
    opt = orc.make_options(**FUSED_VARIANTS[variant])
    port.set_math_mode(mode)
    emu.set_math_mode(mode)
    try:
        g = port.geom_build(tris)
        ch = orc.RestirChain(port, W, H, tris, g, *CAM_AO, opt)
        ge = emu.geom_build(tris)
        fu = EmuFusedFrame(emu, W, H, tris, ge, *CAM_AO, opt)
        for _ in range(3):
            ch.step()
            fu.step()
            if not opt["use_spatial_resampling"]:
                # the reference still copies buf0 -> buf1 in its (disabled) spatial passes; same reservoirs
                pass
        d = diffuse_mask(ch.vis, tris)
        assert d.sum() > 1000
        assert same(ch.vis["index"], fu.vis["index"]) and same(ch.vis["uv"], fu.vis["uv"])
        assert reservoir_mismatch(ch.temporal, fu.aos(fu.T)) == 0
        assert reservoir_mismatch(ch.out[d], fu.output()[d]) == 0
        assert same(ch.accum, fu.accum)
        assert same(port.tone_mapping(ch.accum, W, H), fu.pixels.view(np.uint8))
        port.geom_free(g)
        emu.geom_free(ge)
    finally:
        port.set_math_mode(0)
        emu.set_math_mode(0)


def test_resolve_reuses_traced_visibility(emu, port):
    """resolve's shadow ray towards a sample whose origin is this pixel's own surface is the visibility-reuse ray
    already traced for that sample: using the stored answer (restir_fast.cuh: kTracedBit) changes no bit of the
    image and saves rays"""
    tris = lit_blocks_ao()
    W, H = 96, 54
    opt = orc.make_options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    g = port.geom_build(tris)
    ch = orc.RestirChain(port, W, H, tris, g, *CAM_AO, opt)
    ge = emu.geom_build(tris)
    emu.lib.emu_resolve_rays.restype = C.c_long
    rays = {}
    try:
        for reuse in (1, 0):
            emu.lib.emu_set_resolve_reuse(reuse)
            fu = EmuFusedFrame(emu, W, H, tris, ge, *CAM_AO, opt)
            emu.lib.emu_resolve_rays()
            for f in range(4):
                fu.step()
                if reuse:
                    ch.step()
                    assert same(ch.accum, fu.accum), f
            rays[reuse] = emu.lib.emu_resolve_rays()
            assert same(ch.accum, fu.accum)
    finally:
        emu.lib.emu_set_resolve_reuse(1)
    d = int(diffuse_mask(ch.vis, tris).sum())
    assert rays[0] == 4 * d and 0 < rays[1] < 0.95 * rays[0], (rays, d)
    port.geom_free(g)
    emu.geom_free(ge)


def test_fused_frame_follows_a_moving_camera(emu, port):
    """the reference keeps its reservoirs when the camera moves (only the accumulation is cleared,
    10_restir_di.cpp:257-267): history then holds samples whose origin is another surface — the fused frame must
    neither reuse their traced visibility in resolve nor lose the merge order"""
    tris = lit_blocks_ao()
    W, H = 96, 54
    opt = orc.make_options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    g = port.geom_build(tris)
    ge = emu.geom_build(tris)
    ch = orc.RestirChain(port, W, H, tris, g, *CAM_AO, opt)
    fu = EmuFusedFrame(emu, W, H, tris, ge, *CAM_AO, opt)
    cams = [CAM_AO, CAM_AO, ((8.5, 7.5, 8.0), (0.0, 0.5, 0.0)), ((8.5, 7.5, 8.0), (0.0, 0.5, 0.0)), CAM_AO, CAM_AO]
    for i, (eye, ctr) in enumerate(cams):
        for o, drv in ((port, ch), (emu, fu)):
            drv.eye = np.asarray(eye, np.float32)
            drv.rg = o.lookat(eye, ctr, W, H)
        ch.step()
        fu.step()
        d = diffuse_mask(ch.vis, tris)
        assert same(ch.vis["index"], fu.vis["index"]), i
        assert reservoir_mismatch(ch.temporal, fu.aos(fu.T)) == 0, i
        assert reservoir_mismatch(ch.out[d], fu.output()[d]) == 0, i
        assert same(ch.accum, fu.accum), i
    port.geom_free(g)
    emu.geom_free(ge)


def test_soa_aos_round_trip(emu):
    rng = np.random.default_rng(5)
    n = 1000
    a = np.zeros(n, orc.RESERVOIR)
    for f in ("origin_position", "origin_normal", "hit_position", "hit_normal", "radiance"):
        a[f] = rng.standard_normal((n, 3)).astype(np.float32)
    a["visibility"] = rng.integers(0, 2, n)
    a["w_sum"], a["ucw"] = rng.random(n, np.float32), rng.random(n, np.float32)
    a["M"] = rng.integers(0, 2**29 - 1, n)  # 29 bits of M in the record, 3 flag bits (restir_fast.cuh)
    soa = np.zeros(n * 76, np.uint8)
    emu.lib.emu_aos_to_soa(a.ctypes.data_as(C.c_void_p), soa.ctypes.data_as(C.c_void_p), C.c_long(n))
    b = np.zeros(n, orc.RESERVOIR)
    emu.lib.emu_soa_to_aos(soa.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_long(n))
    assert reservoir_mismatch(a, b) == 0
    assert not soa[72 * n:].any()  # 72 of the 76 bytes per pixel are used
