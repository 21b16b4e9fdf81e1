"""Diagnostic: where does the merged-BLAS traversal spend its time?  Times closest-hit traces of ray subsets and bisects the slow ones."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nexus_b200 as nx
from nexus_b200 import scenes
import bench

res = (3840, 2160)
ctx = nx.Context(0)
desc = bench.make_desc("instanced10m_4k")
scene = scenes.build(ctx, desc, res)
m = scene.ExportMerged(bounds=False)
print("merged prims", m["bvh"].primCount if m else 0, "nodes", m["bvh"].nodeCount if m else 0, "entries", len(scene.ExportTlasEntries()), flush=True)
o, d = scenes.camera_rays(desc["camera"], res)
rays = nx.make_rays(o, d)
h, ms = scene.TraceClosest(rays, timed=True)
h, ms = scene.TraceClosest(rays, timed=True)
print("primary 4K:", round(ms, 2), "ms; hit fraction", float((h["prim"] != 0xffffffff).mean()), flush=True)
# secondary rays as the path tracer makes them: from hit points, random directions, offset 1e-3
rng = np.random.default_rng(3)
ok = np.nonzero(h["prim"] != 0xffffffff)[0]
dd = rng.normal(size=(len(ok), 3)).astype(np.float32); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
oo = (rays["origin"][ok] + rays["direction"][ok] * h["t"][ok, None] + 1e-3 * dd).astype(np.float32)
sec = nx.make_rays(oo, dd)
h2, ms2 = scene.TraceClosest(sec, timed=True)
print("secondary:", len(sec), round(ms2, 2), "ms", flush=True)
import torch
dev_r = torch.from_numpy(rays.view(np.uint8).reshape(-1, 32)).cuda(); dev_h = torch.empty((len(rays), 20), dtype=torch.uint8, device="cuda")
st = scene.TraceStats(dev_r.data_ptr(), len(rays), dev_h.data_ptr())
print("primary per ray:", {k: round(v / max(st["rays"], 1), 2) for k, v in st.items()})
dev_s = torch.from_numpy(sec.view(np.uint8).reshape(-1, 32)).cuda(); dev_h2 = torch.empty((len(sec), 20), dtype=torch.uint8, device="cuda")
st = scene.TraceStats(dev_s.data_ptr(), len(sec), dev_h2.data_ptr())
print("secondary per ray:", {k: round(v / max(st["rays"], 1), 2) for k, v in st.items()})
pt = nx.PathTracer(ctx, res)
import copy
for L in (4, 8):
    st_ = copy.copy(desc["settings"]); st_.pathLength = L
    scene.SetRenderSettings(st_)
    pt.ResetFrameNumber(); pt.SetProfiling(events=True, work=False)
    pt.Render(scene, frames=1, firstFrame=1)
    pr = pt.Profile(); s2 = pt.Stats()
    print("pathLength", L, "ms/frame", round(s2["device_ms"], 2), {k: round(pr[k]["ms"], 2) for k in pt.KERNELS}, "rays", s2["extension_rays"], s2["shadow_rays"], flush=True)
scene.SetRenderSettings(desc["settings"])
pt.SetProfiling(events=False, work=False)
pt.ResetFrameNumber()
pt.Render(scene, frames=4, firstFrame=1); ctx.synchronize()
s2 = pt.Stats()
print("render 4 frames, overlapped: ms/frame", round(s2["device_ms"] / 4, 2), "Mrays/s", round((s2["extension_rays"] + s2["shadow_rays"]) / s2["device_ms"] / 1e3, 1), "mean", float(pt.ReadAccumulation().mean()), flush=True)
sys.exit(0)
pt.ResetFrameNumber(); pt.SetProfiling(events=True, work=False)
pt.Render(scene, frames=2, firstFrame=1)
pr = pt.Profile(); s2 = pt.Stats()
print("render: ms/frame", round(s2["device_ms"] / 2, 2), {k: round(pr[k]["ms"] / 2, 2) for k in pt.KERNELS}, flush=True)
