"""Long-horizon parity of the fused frame against the CPU oracle: 64 accumulated frames of config 4/5's scene and
camera (blocks_restir, 1.6 M triangles) at 480x270, in the three math modes.  Runs on the GPU box (the oracle port
travels as a built .so).  Output committed as profiles/r1/long_horizon_parity.txt."""
import os, sys, time
import numpy as np
ROOT=os.environ.get("GRAFT_REPO_ROOT","/root/repo")
sys.path.insert(0, ROOT+"/oracle"); sys.path.insert(0, ROOT+"/cedec-2024-rt_b200/python"); sys.path.insert(0, ROOT+"/tests")
import orc, cedecrt, stage_assets
tris=stage_assets.load_scene("blocks_restir")
CAM=((-0.579885, 22.194597, -6.567105), (5.224952, 20.847435, 1.431192))
W,H,N=480,270,64
port=orc.load("port")
kw=dict(accumulate=1,use_temporal_resampling=1,use_spatial_resampling=1)
rt=cedecrt.Runtime(0)
def rel_l1(a,b):
    ra, rb = a[:, :3] / a[:, 3:4], b[:, :3] / b[:, 3:4]
    return float(np.abs(ra - rb).sum() / np.abs(rb).sum())
g=port.geom_build(tris)
for mode in (1,0,2):
    port.set_math_mode(1 if mode == 1 else 0)
    rt.set_math_mode({1: cedecrt.MATH_EXACT, 0: cedecrt.MATH_LIBDEVICE, 2: cedecrt.MATH_FAST}[mode])
    ch=orc.RestirChain(port,W,H,tris,g,*CAM,orc.make_options(**kw))
    app=cedecrt.RestirDI(rt,W,H,tris,*CAM,cedecrt.Options(**kw),fused=True)
    t=time.time()
    for f in range(N):
        ch.step(); app.frame()
        if f in (0,1,3,7,15,31,63):
            acc=app.accumulation.to_host().view(np.float32).reshape(-1,4)
            ids = bool((app.visibility.to_host()["index"] == ch.vis["index"]).all())
            print("mode %s frame %d: bit-equal %s  primitive ids equal %s  rel L1 %.3e  (%.0fs)" % (
                {1: "exact", 0: "libdevice", 2: "fast"}[mode], f + 1, acc.tobytes() == ch.accum.tobytes(), ids,
                rel_l1(acc, ch.accum), time.time() - t), flush=True)
