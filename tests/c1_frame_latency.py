import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers as H
from prototype2_b200 import capi
g = np.load('/root/repo/tests/golden/c1_level.npz')
for mode,name in ((capi.ORDER_LITERAL,'literal'),(capi.ORDER_CANONICAL,'canonical')):
    ctx, tf, ph = H.build_gpu(H.c1_scene(g), order_mode=mode)
    ts=[]
    for f in range(int(g["frames"])):
        if g["jump_script"][f]: ph["velocity"][0,1]=np.float32(0.35)
        ph["movement"][0]=g["move_script"][f]
        t=time.perf_counter(); ctx.update_characters(float(g["dt"]), tf, ph); ts.append(time.perf_counter()-t)
    ts=np.array(ts[5:])*1e3
    print(name, "C1 frame ms: median %.3f p95 %.3f max %.3f"%(np.median(ts), np.percentile(ts,95), ts.max()))
# L0 single-thread timing for comparison
from oraclelib import L0World
w = H.build_oracle(H.c1_scene(g), "l0", variant="plain")
ts=[]
for f in range(int(g["frames"])):
    s=w.get_state(); vel=s["vel"].copy()
    if g["jump_script"][f]: vel[0,1]=np.float32(0.35)
    w.set_state(vel=vel, move=g["move_script"][f:f+1])
    t=time.perf_counter(); w.step(float(g["dt"])); ts.append(time.perf_counter()-t)
ts=np.array(ts[5:])*1e3
print("reference (L0, 1 thread) C1 frame ms: median %.4f p95 %.4f"%(np.median(ts), np.percentile(ts,95)))
