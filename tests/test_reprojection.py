"""Temporal reprojection — an extension: the reference reads the previous reservoir at the same
pixel and assumes a static camera.  oracle/port (reproject_pixel, k_temporal with a previous camera) is the specification;
here it is checked for the properties that tie it to the reference, and the CUDA path's device code — compiled for the
host by tests/emu — is checked against it bit for bit.  The GPU test proper is tests/test_gpu_configs.py."""
import ctypes as C

import numpy as np
import pytest

import orc
from helpers import reservoir_mismatch, same, small_scene
from test_emu_parity import EmuFusedFrame, emu, lit_blocks_ao  # noqa: F401  (fixture)

CAM_CB = ((0.0, 2.7, 9.0), (0.0, 2.7, 0.0))
CAM_AO = ((8.0, 8.0, 8.0), (0.0, 0.0, 0.0))
KW = dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)


def test_reprojection_inverts_the_camera(port):
    """a point on pixel (xi, yi)'s own ray projects back to (xi, yi) in the same camera; points behind the camera or
    outside the image have no history"""
    W, H = 96, 54
    rg = port.lookat(*CAM_CB, W, H)
    lib = port.lib
    lib.orc_reproject_pixel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    o, right, up = rg["m_origin"], rg["m_right"], rg["m_up"]
    forward = np.cross(up, right)
    forward /= np.linalg.norm(forward)
    xy = np.zeros(2, np.int32)
    rng = np.random.default_rng(3)
    for _ in range(2000):
        xi, yi = int(rng.integers(W)), int(rng.integers(H))
        u, v = np.float32(xi) / np.float32(W), np.float32(yi) / np.float32(H)
        to = o + forward + right * (2 * u - 1) + up * (1 - 2 * v)  # RayGenerator::shoot, camera.hpp:27-35
        p = (o + (to - o) * np.float32(rng.uniform(0.5, 40.0))).astype(np.float32)
        assert lib.orc_reproject_pixel(rg.ctypes.data, W, H, p.ctypes.data, xy.ctypes.data) == 1
        assert (int(xy[0]), int(xy[1])) == (xi, yi)
    behind = (o - forward * 3).astype(np.float32)
    assert lib.orc_reproject_pixel(rg.ctypes.data, W, H, behind.ctypes.data, xy.ctypes.data) == 0
    aside = (o + forward + right * 1.5).astype(np.float32)
    assert lib.orc_reproject_pixel(rg.ctypes.data, W, H, aside.ctypes.data, xy.ctypes.data) == 0


@pytest.mark.parametrize("scene", ["cornellbox1", "blocks_ao_lit"])
def test_static_camera_equals_the_reference_kernel(port, scene):
    """with an unmoved camera the reprojected lookup is the pixel itself: four frames of the chain with reprojection equal
    the reference's chain bit for bit (reservoirs and accumulation)"""
    tris, cam = (small_scene("cornellbox1"), CAM_CB) if scene == "cornellbox1" else (lit_blocks_ao(), CAM_AO)
    W, H = 96, 54
    port.set_math_mode(0)
    g = port.geom_build(tris)
    a = orc.RestirChain(port, W, H, tris, g, *cam, orc.make_options(**KW))
    b = orc.RestirChain(port, W, H, tris, g, *cam, orc.make_options(**KW), reproject=True)
    for _ in range(4):
        a.step()
        b.step()
    assert same(a.accum, b.accum) and reservoir_mismatch(a.temporal, b.temporal) == 0 and reservoir_mismatch(a.out, b.out) == 0
    assert float(a.accum[:, :3].sum()) > 0
    port.geom_free(g)


def moved_camera_chain(o, tris, W, H, mode, reproject):
    o.set_math_mode(mode)
    g = o.geom_build(tris)
    ch = orc.RestirChain(o, W, H, tris, g, *CAM_AO, orc.make_options(**KW), reproject=reproject)
    for _ in range(2):
        ch.step()
    ch.set_camera((8.4, 7.8, 8.1), (0.1, 0.0, -0.1))  # a small move: most surface points stay on screen
    ch.step()
    ch.set_camera((8.9, 7.5, 8.3), (0.2, 0.1, -0.2))
    ch.step()
    o.set_math_mode(0)
    return ch


@pytest.mark.parametrize("mode", [0, 1])
def test_device_code_equals_the_specification_with_a_moving_camera(emu, port, mode):
    tris = lit_blocks_ao()
    W, H = 128, 72
    a = moved_camera_chain(port, tris, W, H, mode, True)
    b = moved_camera_chain(emu, tris, W, H, mode, True)
    assert same(a.vis["index"], b.vis["index"])
    assert reservoir_mismatch(a.temporal, b.temporal) == 0 and reservoir_mismatch(a.out, b.out) == 0
    assert same(a.accum, b.accum)
    # and reprojection does find history: after the move most diffuse pixels carry more than this frame's 32 candidates,
    # which the same-pixel lookup (the reference's behaviour) also does — but with other, wrong-surface reservoirs
    c = moved_camera_chain(port, tris, W, H, mode, False)
    d = a.vis["index"] >= 0
    em = (tris["emissive"] > 0).any(1)
    d[d] = ~em[a.vis["index"][d]]
    assert (a.temporal["M"][d] > 32).mean() > 0.8
    assert reservoir_mismatch(a.temporal[d], c.temporal[d]) > 0.3 * d.sum()


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_frame_with_reprojection_equals_the_specification(emu, port, mode):
    """crt_restir_set_previous_camera: the fused frame's candidate + temporal body with the history looked up at the
    reprojected pixel in a snapshot of last frame's records (the merge itself stays in place) — device code on the host —
    against the specification's kernel chain, camera moving twice; and with an unmoved camera against the plain fused frame"""
    tris = lit_blocks_ao()
    W, H = 128, 72
    spec = moved_camera_chain(port, tris, W, H, mode, True)
    emu.set_math_mode(mode)
    g = emu.geom_build(tris)
    f = EmuFusedFrame(emu, W, H, tris, g, *CAM_AO, orc.make_options(**KW))
    f.reproject = True
    for _ in range(2):
        f.step()
    f.set_camera((8.4, 7.8, 8.1), (0.1, 0.0, -0.1))
    f.step()
    f.set_camera((8.9, 7.5, 8.3), (0.2, 0.1, -0.2))
    f.step()
    assert same(f.vis["index"], spec.vis["index"])
    d = spec.vis["index"] >= 0
    em = (tris["emissive"] > 0).any(1)
    d[d] = ~em[spec.vis["index"][d]]
    assert reservoir_mismatch(spec.temporal[d], f.aos(f.T)[d]) == 0 and reservoir_mismatch(spec.out[d], f.output()[d]) == 0
    assert same(f.accum, spec.accum)
    # unmoved camera: reprojection on and off give the same frames
    a = EmuFusedFrame(emu, W, H, tris, g, *CAM_AO, orc.make_options(**KW))
    b = EmuFusedFrame(emu, W, H, tris, g, *CAM_AO, orc.make_options(**KW))
    b.reproject = True
    for _ in range(3):
        a.step()
        b.step()
    assert same(a.T, b.T) and same(a.accum, b.accum) and same(a.pixels, b.pixels)
    emu.set_math_mode(0)
    emu.geom_free(g)
