"""TEST INFRASTRUCTURE — ctypes front end of the two CPU oracles.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product path (cedec-2024-rt_b200/) never does.

  kind="port"       oracle/liboracle_port.so   restatement of the reference algorithm (oracle/port/)
  kind="reference"  oracle/_ref/libref_NN.so   the reference's own unmodified .cu compiled as host C++

Both export the same `orc_*` C API (oracle/ref_shim/ref_driver.cpp, oracle/port/oracle_port.cpp).
Struct layouts follow the reference: Triangle common/core.hpp:38-43 (60 B), Visibility core.hpp:167-172
(16 B), ReservoirSample/Reservoir common/reservoir.hpp:5-38 (64/76 B), Options common/options.hpp:4-23
(48 B), RayGenerator common/camera.hpp:5-9 (36 B).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

f3 = (np.float32, (3,))
TRIANGLE = np.dtype([("vertices", np.float32, (3, 3)), ("color", *f3), ("emissive", *f3)])
VISIBILITY = np.dtype([("uv", np.float32, (2,)), ("index", np.int32), ("_pad", np.int32)])
RESERVOIR = np.dtype(
    [
        ("origin_position", *f3),
        ("origin_normal", *f3),
        ("hit_position", *f3),
        ("hit_normal", *f3),
        ("radiance", *f3),
        ("visibility", np.uint8),
        ("_pad", np.uint8, (3,)),
        ("w_sum", np.float32),
        ("ucw", np.float32),
        ("M", np.int32),
    ]
)
OPTIONS = np.dtype(
    [
        ("accumulate", np.uint8),
        ("_p0", np.uint8, (3,)),
        ("max_depth", np.int32),
        ("sky_color", *f3),
        ("ris_sample_count", np.int32),
        ("rejection_heuristics_threshold", np.float32),
        ("use_temporal_resampling", np.uint8),
        ("use_spatial_resampling", np.uint8),
        ("_p1", np.uint8, (2,)),
        ("spatial_resampling_sample_count", np.int32),
        ("spatial_resampling_radius", np.float32),
        ("spatial_resampling_passes", np.int32),
        ("use_shadowed_target_function", np.uint8),
        ("use_visibility_reuse", np.uint8),
        ("_p2", np.uint8, (2,)),
    ]
)
RAYGEN = np.dtype([("m_origin", *f3), ("m_right", *f3), ("m_up", *f3)])
assert TRIANGLE.itemsize == 60 and VISIBILITY.itemsize == 16 and RESERVOIR.itemsize == 76
assert OPTIONS.itemsize == 48 and RAYGEN.itemsize == 36


def make_options(**kw):
    """Options with the reference defaults (common/options.hpp:4-23)."""
    o = np.zeros((), OPTIONS)
    o["accumulate"] = 0
    o["max_depth"] = 6
    o["ris_sample_count"] = 32
    o["rejection_heuristics_threshold"] = 0.2
    o["spatial_resampling_sample_count"] = 5
    o["spatial_resampling_radius"] = 30.0
    o["spatial_resampling_passes"] = 3
    o["use_visibility_reuse"] = 1
    for k, v in kw.items():
        o[k] = v
    return o


SURVEY_FNV_BASIS = 1469598103934665603  # the offset basis of the survey's probe: the standard one minus its last digit


def fnv1a64(a, basis=None):
    """FNV-1a-64 over the bytes of an array.  basis=SURVEY_FNV_BASIS reproduces the goldens SURVEY.md sections 2.3 / 4
    quote (their probe started from a truncated offset basis); the default is the standard 0xcbf29ce484222325."""
    data = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    lib = _fnv_lib()
    if basis is None:
        return "%016x" % lib.orc_fnv1a64(data.ctypes.data_as(C.c_void_p), C.c_size_t(data.size))
    lib.orc_fnv1a64_from.restype = C.c_uint64
    lib.orc_fnv1a64_from.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
    return "%016x" % lib.orc_fnv1a64_from(data.ctypes.data_as(C.c_void_p), C.c_size_t(data.size), C.c_uint64(basis))


_FNV = None


def _fnv_lib():
    global _FNV
    if _FNV is None:
        _FNV = load("port").lib
    return _FNV


def light_indices(tris):
    """Indices of emissive triangles in ascending order (10_restir_di.cpp:196-206)."""
    e = tris["emissive"]
    return np.nonzero((e[:, 0] > 0) | (e[:, 1] > 0) | (e[:, 2] > 0))[0].astype(np.uint32)


def build(targets=("port",)):
    subprocess.run(["make", "-s", "-C", HERE, *targets], check=True)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One loaded oracle library.  example: 10 (ReSTIR DI), 9, 8, 7 (path tracers), 6, 4 (AO)."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_kind.restype = C.c_char_p
        L.orc_geom_build.restype = C.c_void_p
        L.orc_geom_build.argtypes = [C.c_void_p, C.c_int]
        L.orc_geom_free.argtypes = [C.c_void_p]
        L.orc_set_range.argtypes = [C.c_long, C.c_long]
        L.orc_closest_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.orc_lookat.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p]
        if hasattr(L, "orc_fnv1a64"):
            L.orc_fnv1a64.restype = C.c_uint64
            L.orc_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
        self.kind = L.orc_kind().decode()
        self.path = path

    # -- control
    def threads(self):
        return self.lib.orc_threads()

    def set_threads(self, n):
        self.lib.orc_set_threads(int(n))

    def set_math_mode(self, m):
        """0 = libm float functions (the reference as host C++), 1 = correctly rounded via double."""
        self.lib.orc_set_math_mode(int(m))

    def set_arg_order(self, o):
        """port only: 0 = left-to-right argument evaluation (GPU compilers; canonical), 1 = g++'s right-to-left."""
        self.lib.orc_set_arg_order(int(o))

    def set_example(self, e):
        """port only: which example orc_path_trace / orc_ao restate (7, 8, 9 / 4, 6)."""
        self.lib.orc_set_example(int(e))

    def set_range(self, begin=0, end=-1):
        self.lib.orc_set_range(int(begin), int(end))

    # -- geometry
    def geom_build(self, tris):
        assert tris.dtype == TRIANGLE and tris.flags.c_contiguous
        g = self.lib.orc_geom_build(_p(tris), len(tris))
        return C.c_void_p(g)

    def geom_free(self, g):
        self.lib.orc_geom_free(g)

    def lookat(self, eye, center, W, H, up=(0, 1, 0), fovy=None):
        # fovy = PI / 4.0f with float PI (10_restir_di.cpp:243)
        if fovy is None:
            fovy = np.float32(np.float32(3.14159265358979323846) / np.float32(4.0))
        rg = np.zeros((), RAYGEN)
        e, c, u = (np.asarray(v, np.float32) for v in (eye, center, up))
        self.lib.orc_lookat(_p(e), _p(c), _p(u), C.c_float(float(fovy)), W, H, _p(rg))
        return rg

    def closest_hit(self, g, o, d, tmin=0.0, tmax=3.402823466e38):
        o, d = np.asarray(o, np.float32), np.asarray(d, np.float32)
        tuv = np.zeros(3, np.float32)
        idx = self.lib.orc_closest_hit(g, _p(o), _p(d), C.c_float(tmin), C.c_float(tmax), _p(tuv))
        return idx, tuv

    # -- example 10 kernels (10_restir_di.cu)
    def raycast(self, W, H, g, tris, rg, vis=None):
        vis = np.zeros(W * H, VISIBILITY) if vis is None else vis
        self.lib.orc_raycast(W, H, g, _p(tris), len(tris), _p(rg), _p(vis))
        return vis

    def generate_candidate(self, W, H, frame, g, tris, vis, eye, lights, opt, res=None):
        res = np.zeros(W * H, RESERVOIR) if res is None else res
        eye = np.asarray(eye, np.float32)
        self.lib.orc_generate_candidate(
            W, H, frame, g, _p(tris), len(tris), _p(vis), _p(eye), _p(lights), len(lights), _p(opt), _p(res)
        )
        return res

    def temporal_resampling(self, W, H, frame, g, tris, vis, eye, opt, prev, res):
        eye = np.asarray(eye, np.float32)
        self.lib.orc_temporal_resampling(W, H, frame, g, _p(tris), len(tris), _p(vis), _p(eye), _p(opt), _p(prev), _p(res))
        return res

    def temporal_resampling_reprojected(self, W, H, frame, g, tris, vis, eye, opt, prev_rg, prev, res):
        """extension (oracle/port: reproject_pixel is its specification): the previous reservoir is
        read at the pixel the current surface point had in the previous frame's camera `prev_rg`"""
        eye = np.asarray(eye, np.float32)
        self.lib.orc_temporal_resampling_reprojected(W, H, frame, g, _p(tris), len(tris), _p(vis), _p(eye), _p(opt),
                                                     _p(prev_rg), _p(prev), _p(res))
        return res

    def save_temporal_reservoir(self, W, H, src, dst):
        self.lib.orc_save_temporal_reservoir(W, H, _p(src), _p(dst))
        return dst

    def spatial_resampling(self, W, H, frame, pas, g, tris, vis, eye, opt, rin, rout):
        eye = np.asarray(eye, np.float32)
        self.lib.orc_spatial_resampling(
            W, H, frame, pas, g, _p(tris), len(tris), _p(vis), _p(eye), _p(opt), _p(rin), _p(rout)
        )
        return rout

    def resolve(self, accum, W, H, g, tris, vis, eye, opt, res):
        eye = np.asarray(eye, np.float32)
        self.lib.orc_resolve(_p(accum), W, H, g, _p(tris), len(tris), _p(vis), _p(eye), _p(opt), _p(res))
        return accum

    def clear(self, buf, W, H):
        self.lib.orc_clear(_p(buf), W, H)
        return buf

    def tone_mapping(self, accum, W, H, pixels=None):
        pixels = np.zeros(W * H * 4, np.uint8) if pixels is None else pixels
        self.lib.orc_tone_mapping(_p(pixels), _p(accum), W, H)
        return pixels

    # -- examples 07/08/09 (path_trace) and 04/06 (AO)
    def path_trace(self, W, H, frame, g, tris, lights, rg, opt, accum):
        self.lib.orc_path_trace(W, H, frame, g, _p(tris), len(tris), _p(lights), len(lights), _p(rg), _p(opt), _p(accum))
        return accum

    def ao(self, W, H, g, tris, rg, n_rays=64, pixels=None):
        pixels = np.zeros(W * H * 4, np.uint8) if pixels is None else pixels
        rc = self.lib.orc_ao(_p(pixels), _p(rg), W, H, g, _p(tris), len(tris), n_rays)
        if rc != 0:
            raise ValueError("this oracle supports n_rays=64 only (hard-coded in the reference)")
        return pixels


class RestirChain:
    """The frame loop of 10_restir_di.cpp:229-380 driven over any object with the Oracle kernel methods."""

    def __init__(self, orc, W, H, tris, g, eye, center, opt, lights=None, reproject=False):
        """reproject=True (extension): temporal resampling reads the previous reservoir at the pixel
        the surface point had in the previous frame's camera; set_camera() moves the camera between frames and clears the
        accumulation like the reference does on a move (10_restir_di.cpp:257-267)"""
        self.o, self.W, self.H, self.tris, self.g, self.opt = orc, W, H, tris, g, opt
        self.reproject, self.prev_rg = reproject, None
        self.eye = np.asarray(eye, np.float32)
        self.rg = orc.lookat(eye, center, W, H)
        self.lights = light_indices(tris) if lights is None else lights
        n = W * H
        self.vis = np.zeros(n, VISIBILITY)
        self.buf0 = np.zeros(n, RESERVOIR)
        self.buf1 = np.zeros(n, RESERVOIR)
        self.temporal = np.zeros(n, RESERVOIR)  # zero-initialised (frame-1 rule)
        self.accum = np.zeros((n, 4), np.float32)
        self.frame = 0
        self.out = self.buf1

    def set_camera(self, eye, center):
        self.eye = np.asarray(eye, np.float32)
        self.rg = self.o.lookat(eye, center, self.W, self.H)
        self.accum[:] = 0  # `clear` on a camera move (10_restir_di.cpp:257-267)

    def step(self):
        o, W, H, g, t, opt, eye = self.o, self.W, self.H, self.g, self.tris, self.opt, self.eye
        self.frame += 1  # first frame is 1 (10_restir_di.cpp:233)
        o.raycast(W, H, g, t, self.rg, self.vis)
        o.generate_candidate(W, H, self.frame, g, t, self.vis, eye, self.lights, opt, self.buf0)
        if self.reproject and self.prev_rg is not None:
            o.temporal_resampling_reprojected(W, H, self.frame, g, t, self.vis, eye, opt, self.prev_rg, self.temporal, self.buf0)
        else:
            o.temporal_resampling(W, H, self.frame, g, t, self.vis, eye, opt, self.temporal, self.buf0)
        self.prev_rg = self.rg  # set_camera replaces the object, it never mutates it
        o.save_temporal_reservoir(W, H, self.buf0, self.temporal)
        bi, bo = self.buf0, self.buf1
        for k in range(int(opt["spatial_resampling_passes"])):
            if k != 0:
                bi, bo = bo, bi
            o.spatial_resampling(W, H, self.frame, k, g, t, self.vis, eye, opt, bi, bo)
        self.out = bo
        o.resolve(self.accum, W, H, g, t, self.vis, eye, opt, bo)
        return self.accum


def load(kind="port", example=10):
    if kind == "port":
        path = os.path.join(HERE, "liboracle_port.so")
        if not os.path.exists(path):
            build(("port",))
    elif kind == "reference":
        path = os.path.join(HERE, "_ref", "libref_%02d.so" % example)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle ref` where /root/reference exists)")
    else:
        raise ValueError(kind)
    return Oracle(path)


def have_reference():
    return os.path.exists(os.path.join(HERE, "_ref", "libref_10.so"))


def load_obj_reference(obj_path, mtl_dir):
    """Triangle[] through the reference's own loader (common/loader.hpp:11-66); needs oracle/_ref."""
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_loader.so"))
    lib.orc_load_obj.restype = C.c_long
    lib.orc_load_obj.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.orc_free.argtypes = [C.c_void_p]
    out = C.c_void_p()
    n = lib.orc_load_obj(obj_path.encode(), mtl_dir.encode(), C.byref(out))
    buf = (C.c_char * (n * 60)).from_address(out.value)
    tris = np.frombuffer(buf, dtype=TRIANGLE, count=n).copy()
    lib.orc_free(out)
    return tris
