"""CPU-side checks of the drop-in boundary: libcedecrt.so loads, exports every symbol include/cedecrt.h
declares, refuses to run without a GPU (no CPU fallback), and its host arithmetic equals the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import orc
from helpers import same

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cedecrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import cedecrt

    lib = cedecrt.lib()
    names = _declared()
    assert len(names) >= 35, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_reference():
    import cedecrt

    assert C.sizeof(cedecrt.Options) == 48 and C.sizeof(cedecrt.RayGenerator) == 36 and C.sizeof(cedecrt._Buffer) == 16
    assert cedecrt.TRIANGLE.itemsize == 60 and cedecrt.VISIBILITY.itemsize == 16 and cedecrt.RESERVOIR.itemsize == 76
    # Options defaults = common/options.hpp:4-23, byte for byte the oracle's record
    assert bytes(cedecrt.Options()) == orc.make_options().tobytes()
    o = cedecrt.Options(accumulate=1, use_temporal_resampling=1, sky_color=(0.1, 0.2, 0.3))
    assert bytes(o) == orc.make_options(accumulate=1, use_temporal_resampling=1, sky_color=(0.1, 0.2, 0.3)).tobytes()


def test_lookat_matches_oracle(port):
    import cedecrt

    for eye, ctr, W, H in (((8, 8, 8), (0, 0, 0), 1920, 1080), ((0, 2.7, 9), (0, 2.7, 0), 96, 54),
                           ((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192), 3840, 2160)):
        rg = cedecrt.lookat(eye, ctr, W, H)
        assert bytes(rg) == port.lookat(eye, ctr, W, H).tobytes()


def test_no_cpu_fallback():
    """without a CUDA device the product path must fail loudly (CRT_ENODEVICE), never compute on the CPU"""
    import cedecrt

    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cedecrt.CrtError, match="no CUDA device"):
        cedecrt.Runtime(0)


def test_product_sources_never_touch_the_oracle():
    pkg = os.path.join(ROOT, "cedec-2024-rt_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".cpp", ".hpp")):
                text = open(os.path.join(d, f), errors="ignore").read()
                text = text.replace("oracle/orc.py OPTIONS", "")  # a docstring naming the record layout
                assert "liboracle" not in text and "import orc" not in text and "oracle/" not in text, os.path.join(d, f)
