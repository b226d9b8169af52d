"""Edge cases of the hot path on the GPU, through the C ABI, against the CPU oracle (oracle/port): image sizes that do
not fill the kernels' 32x8 tiles (down to one pixel), degenerate geometry (zero-area and duplicated triangles, a
triangle the ray grazes edge-on), a scene with a single light, rays that start on the triangle that stops them (the
own-triangle pre-test's case).  Bars as in test_gpu_parity.py: bit-exact in CRT_MATH_EXACT."""
import numpy as np
import pytest

import cedecrt
import orc
from helpers import DeviceAsOracle, reservoir_mismatch, same, small_scene

pytestmark = pytest.mark.gpu

CAM_CB = ((0.0, 2.7, 9.0), (0.0, 2.7, 0.0))
CAM_AO = ((8.0, 8.0, 8.0), (0.0, 0.0, 0.0))


@pytest.fixture(scope="module")
def rt():
    r = cedecrt.Runtime(0)
    yield r
    r.close()


@pytest.fixture()
def exact(rt, port):
    rt.set_math_mode(cedecrt.MATH_EXACT)
    port.set_math_mode(1)
    yield DeviceAsOracle(rt)
    rt.set_math_mode(cedecrt.MATH_LIBDEVICE)
    port.set_math_mode(0)
    port.set_example(9)


def lit_blocks_ao():
    t = small_scene("blocks_ao").copy()
    t["emissive"][100:140] = (5.0, 4.0, 3.0)
    return t


def diffuse_mask(vis, tris):
    idx = vis["index"]
    hit = idx >= 0
    em = np.zeros(len(idx), bool)
    em[hit] = (tris["emissive"][idx[hit]] > 0).any(axis=1)
    return hit & ~em


@pytest.mark.parametrize("size", [(1, 1), (33, 17), (100, 37), (257, 9), (31, 65)])
def test_ragged_image_sizes_fused_and_per_kernel(rt, port, exact, size):
    """sizes that leave partial 32x8 tiles and partial warps everywhere: the per-kernel chain and the fused frame against
    the oracle's chain, three frames, every buffer the reference's loop leaves behind"""
    W, H = size
    tris = lit_blocks_ao()
    kw = dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    opt = orc.make_options(**kw)
    gp = port.geom_build(tris)
    ref = orc.RestirChain(port, W, H, tris, gp, *CAM_AO, opt)
    g = exact.geom_build(tris)
    per_kernel = orc.RestirChain(exact, W, H, tris, g, *CAM_AO, opt)
    app = cedecrt.RestirDI(rt, W, H, tris, *CAM_AO, cedecrt.Options(**kw), fused=True)
    for _ in range(3):
        ref.step()
        per_kernel.step()
        app.frame()
    assert same(per_kernel.vis, ref.vis)
    assert reservoir_mismatch(ref.buf0, per_kernel.buf0) == 0 and reservoir_mismatch(ref.buf1, per_kernel.buf1) == 0
    assert same(per_kernel.accum, ref.accum)
    vis = app.visibility.to_host()
    assert same(vis["index"], ref.vis["index"]) and same(vis["uv"], ref.vis["uv"])
    d = diffuse_mask(ref.vis, tris)
    assert reservoir_mismatch(ref.temporal, app.export_aos(app.temporal)) == 0
    assert reservoir_mismatch(ref.out[d], app.output_reservoirs()[d]) == 0
    acc = app.accumulation.to_host().view(np.float32).reshape(-1, 4)
    assert same(acc, ref.accum)
    assert same(app.pixels.to_host(), port.tone_mapping(ref.accum, W, H))
    port.geom_free(gp)


@pytest.mark.parametrize("ex", [6, 8, 9])
def test_ragged_image_sizes_examples(port, exact, ex):
    """06_ao / 08_nee / 09_ris (shadowed target: the wavefront form with its hole records) at a size with partial tiles"""
    W, H = 71, 29
    if ex == 6:
        tris = small_scene("blocks_ao")
        port.set_example(6)
        gp, g = port.geom_build(tris), exact.geom_build(tris)
        a = port.ao(W, H, gp, tris, port.lookat(*CAM_AO, W, H), 16)
        b = exact.ao(W, H, g, tris, exact.lookat(*CAM_AO, W, H), 16)
        assert same(a, b)
        return
    tris = small_scene("cornellbox1")
    opt = orc.make_options(accumulate=1, max_depth=5, ris_sample_count=8, use_shadowed_target_function=1 if ex == 9 else 0,
                           sky_color=(0.1, 0.2, 0.3))
    lights = orc.light_indices(tris)
    port.set_example(ex)
    gp, g = port.geom_build(tris), exact.geom_build(tris)
    a, b = np.zeros((W * H, 4), np.float32), np.zeros((W * H, 4), np.float32)
    for frame in (1, 2, 3):
        port.path_trace(W, H, frame, gp, tris, lights, port.lookat(*CAM_CB, W, H), opt, a)
        exact.path_trace(ex, W, H, frame, g, tris, lights, exact.lookat(*CAM_CB, W, H), opt, b)
    assert same(a, b) and float(a[:, :3].sum()) > 0


def test_degenerate_triangles_do_not_change_a_hit(rt):
    """zero-area triangles (two or three equal vertices, collinear vertices), exact duplicates of real triangles (the tie
    rule: the larger primitive id wins) and far-away slivers mixed into a scene: the tree's closest hits and any-hits
    still equal the exhaustive loop's on random and on axis-parallel rays"""
    base = small_scene("cornellbox1")
    rng = np.random.default_rng(5)
    extra = np.zeros(64, dtype=base.dtype)
    extra["color"] = 0.5
    v = extra["vertices"]
    p = rng.uniform(-3, 6, (64, 3)).astype(np.float32)
    v[:, 0] = v[:, 1] = v[:, 2] = p                       # points
    v[16:32, 1] = p[16:32] + np.float32(0.5)              # segments (two equal vertices)
    v[32:48, 1] = p[32:48] + np.float32([1, 0, 0])
    v[32:48, 2] = p[32:48] + np.float32([2, 0, 0])        # collinear
    v[48:, 1] = p[48:] + np.float32([1e3, 1e-4, 0])
    v[48:, 2] = p[48:] + np.float32([1e3, 0, 1e-4])       # slivers a thousand units long
    tris = np.concatenate([base, extra, base[:12]])       # ... and duplicates of the first twelve triangles at the end
    d_tris = rt.to_device(tris)
    g = rt.build_geometry(d_tris)
    n = 8192
    o = rng.uniform(-7, 7, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[::4, 0] = 0.0
    d[1::4, 1] = 0.0
    d[2::8, 2] = 0.0
    o[::3] = np.round(o[::3])
    p1, t1 = rt.trace_closest(g, o, d)
    p2, t2 = rt.trace_closest(g, o, d, brute=True)
    assert same(p1, p2) and same(t1, t2)
    dup = np.isin(p1, np.arange(len(base) + 64, len(tris)))
    assert dup.sum() > 0 and not np.isin(p1, np.arange(12)).any()  # the duplicate with the larger id wins every tie
    a1 = rt.trace_any(g, o, d, 0.0, 4.0)
    p3, _ = rt.trace_closest(g, o, d, 0.0, 4.0, brute=True)
    assert same(a1, (p3 >= 0).astype(np.int32))
    g.destroy()


def test_single_light_and_own_triangle_rays(rt, port, exact):
    """one emissive triangle *below* most of the scene's surfaces: nearly every candidate's visibility ray starts on the
    triangle that stops it — the case the emitting kernels settle with the own-triangle pre-test.  Fused frame against
    the oracle bit for bit, and the pre-test's counter against the number of such rays counted on the host."""
    tris = small_scene("blocks_ao").copy()
    tris["emissive"][:] = 0
    ymin = tris["vertices"][:, :, 1].mean(axis=1)
    low = int(np.argmin(ymin))
    tris["emissive"][low] = (30.0, 30.0, 30.0)
    W, H = 96, 54
    kw = dict(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1, ris_sample_count=4)
    opt = orc.make_options(**kw)
    gp = port.geom_build(tris)
    ref = orc.RestirChain(port, W, H, tris, gp, *CAM_AO, opt)
    app = cedecrt.RestirDI(rt, W, H, tris, *CAM_AO, cedecrt.Options(**kw), fused=True)
    before = rt.rays_decided_at_emission()[0], rt.shadow_rays_traced()[0]
    for _ in range(2):
        ref.step()
        app.frame()
    after = rt.rays_decided_at_emission()[0], rt.shadow_rays_traced()[0]
    d = diffuse_mask(ref.vis, tris)
    assert d.sum() > 1000
    assert reservoir_mismatch(ref.temporal, app.export_aos(app.temporal)) == 0
    assert reservoir_mismatch(ref.out[d], app.output_reservoirs()[d]) == 0
    acc = app.accumulation.to_host().view(np.float32).reshape(-1, 4)
    assert same(acc, ref.accum)
    decided, walked = after[0] - before[0], after[1] - before[1]
    print("visibility-reuse rays: %d settled by the own-triangle pre-test, %d walked" % (decided, walked))
    assert decided > 0 and decided + walked <= 2 * int(d.sum())
    port.geom_free(gp)
