"""Generates tests/golden/*.npz from oracle/_ref — the reference's own unmodified kernels compiled as
host C++ (oracle/Makefile) — and the reference's own OBJ loader.  Needs /root/reference; run once here:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures are what pins the port (oracle/port/) on machines without /root/reference.
All runs use math mode 0 (glibc float functions) and g++'s right-to-left argument evaluation, i.e. exactly
what the reference does when built as host C++ (see oracle/port/oracle_port.cpp, g_arg_order).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import orc  # noqa: E402

A = "/root/reference/assets/"
CORNELL_CAM = ((0.0, 2.7, 9.0), (0.0, 2.7, 0.0))


def main():
    cb = orc.load_obj_reference(A + "cornellbox1.obj", A)
    ao = orc.load_obj_reference(A + "blocks_ao.obj", A)
    np.savez_compressed(os.path.join(HERE, "scenes_small.npz"), cornellbox1=cb, blocks_ao=ao)

    # --- example 10: 4-frame chain, every option on (SURVEY.md section 4 golden: sky 3554, mean M 115.66)
    W, H = 96, 54
    R = orc.load("reference", 10)
    g = R.geom_build(cb)
    opt = orc.make_options(accumulate=1, use_temporal_resampling=1, use_spatial_resampling=1)
    ch = orc.RestirChain(R, W, H, cb, g, *CORNELL_CAM, opt)
    for _ in range(4):
        ch.step()
    pix = R.tone_mapping(ch.accum, W, H)
    out = dict(W=W, H=H, eye=CORNELL_CAM[0], center=CORNELL_CAM[1], opt=opt, rg=ch.rg, vis=ch.vis, buf0=ch.buf0,
               buf1=ch.buf1, temporal=ch.temporal, accum=ch.accum, pixels=pix)
    # shadowed target function, no visibility reuse, 1 frame
    opt2 = orc.make_options(use_temporal_resampling=1, use_spatial_resampling=1, use_shadowed_target_function=1,
                            use_visibility_reuse=0, spatial_resampling_passes=2, ris_sample_count=8)
    ch2 = orc.RestirChain(R, W, H, cb, g, *CORNELL_CAM, opt2)
    ch2.step()
    ch2.step()
    out.update(opt2=opt2, b_buf0=ch2.buf0, b_buf1=ch2.buf1, b_accum=ch2.accum)
    np.savez_compressed(os.path.join(HERE, "restir_cornell_96x54.npz"), **out)

    # --- primary visibility on blocks_ao at the 06 camera (eye (8,8,8) -> (0,0,0))
    W2, H2 = 320, 180
    g2 = R.geom_build(ao)
    rg2 = R.lookat((8, 8, 8), (0, 0, 0), W2, H2)
    vis2 = R.raycast(W2, H2, g2, ao, rg2)
    np.savez_compressed(os.path.join(HERE, "raycast_blocks_ao_320x180.npz"), W=W2, H=H2, rg=rg2, vis=vis2)

    # --- path tracers 07/08/09 on cornellbox1, 2 accumulated frames, max_depth 4
    res = {}
    for ex in (7, 8, 9):
        E = orc.load("reference", ex)
        ge = E.geom_build(cb)
        o = orc.make_options(accumulate=1, max_depth=4, ris_sample_count=8, sky_color=(0.1, 0.2, 0.3))
        acc = np.zeros((W * H, 4), np.float32)
        rg = E.lookat(*CORNELL_CAM, W, H)
        for frame in (1, 2):
            E.path_trace(W, H, frame, ge, cb, orc.light_indices(cb), rg, o, acc)
        res["pt%02d" % ex] = acc
        res["opt%02d" % ex] = o
    o9 = orc.make_options(max_depth=2, ris_sample_count=4, use_shadowed_target_function=1)
    E = orc.load("reference", 9)
    acc = np.zeros((W * H, 4), np.float32)
    E.path_trace(W, H, 1, E.geom_build(cb), cb, orc.light_indices(cb), E.lookat(*CORNELL_CAM, W, H), o9, acc)
    res["pt09_shadowed"] = acc
    res["opt09_shadowed"] = o9
    np.savez_compressed(os.path.join(HERE, "pt_cornell_96x54.npz"), **res)

    # --- AO: 06 (BVH) on blocks_ao and 04 (brute force) on cornellbox1, 64 rays (reference's hard-coded count)
    E6 = orc.load("reference", 6)
    W3, H3 = 160, 90
    p6 = E6.ao(W3, H3, E6.geom_build(ao), ao, E6.lookat((8, 8, 8), (0, 0, 0), W3, H3))
    E4 = orc.load("reference", 4)
    p4 = E4.ao(64, 64, None, cb, E4.lookat((8, 8, 8), (0, 0, 0), 64, 64))
    np.savez_compressed(os.path.join(HERE, "ao_goldens.npz"), ao06=p6, W06=W3, H06=H3, ao04=p4, W04=64, H04=64)
    print("wrote goldens:", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
