"""Experiment: do two independent frame pipelines on one GPU overlap well enough (shade under trace, tails) to raise throughput?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nexus_b200 as nx
from nexus_b200 import scenes
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "instanced10m_4k"
res = bench.WORKLOADS[wl]["res"]
desc = bench.make_desc(wl)
K = 8
ctxs, scs, pts = [], [], []
for i in range(2):
    c = nx.Context(0); s = scenes.build(c, desc, res); p = nx.PathTracer(c, res)
    p.Render(s, frames=2); c.synchronize()
    ctxs.append(c); scs.append(s); pts.append(p)

def run(i, first):
    pts[i].ResetFrameNumber()
    pts[i].Render(scs[i], frames=K, firstFrame=first)
    ctxs[i].synchronize()

for rep in range(2):
    t0 = time.time(); run(0, 1); t1 = time.time() - t0
    st = pts[0].Stats(); rays1 = st["extension_rays"] + st["shadow_rays"]
    print(f"one tracer : {K} frames {t1*1e3/K:7.2f} ms/frame wall, device {st['device_ms']/K:7.2f} ms/frame, {rays1/t1/1e6:8.1f} Mrays/s", flush=True)
    th = [threading.Thread(target=run, args=(i, 1 + i * K)) for i in range(2)]
    t0 = time.time(); [t.start() for t in th]; [t.join() for t in th]; t2 = time.time() - t0
    rays2 = sum(p.Stats()["extension_rays"] + p.Stats()["shadow_rays"] for p in pts)
    print(f"two tracers: {2*K} frames {t2*1e3/(2*K):7.2f} ms/frame wall, {rays2/t2/1e6:8.1f} Mrays/s", flush=True)
