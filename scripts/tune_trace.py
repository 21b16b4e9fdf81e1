"""Sweeps the traversal batching thresholds on a bench workload (GPU): frame time and per-kernel times for each setting."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nexus_b200 as nx
from nexus_b200 import scenes
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "instanced10m_4k"
res = bench.WORKLOADS[wl]["res"]
ctx = nx.Context(0)
desc = bench.make_desc(wl)
if os.environ.get("NX_PATHLEN"):      # e.g. 1: primary rays only, to tune the first (coherent) launch by itself
    desc["settings"].pathLength = int(os.environ["NX_PATHLEN"])
scene = scenes.build(ctx, desc, res)
pt = nx.PathTracer(ctx, res)
pt.Render(scene, frames=2); ctx.synchronize()
if os.environ.get("NX_NO_SPHERE"):
    ctx.SetSphereCull(False)
grid = [tuple(int(x) for x in a.split(",")) for a in sys.argv[2:]] or [(1, 1), (4, 1), (8, 1), (8, 4), (12, 4), (16, 4), (16, 8), (24, 8), (32, 16)]
for tri, inst in grid:
    ctx.SetTraceTuning(tri, inst)
    pt.ResetFrameNumber(); pt.SetProfiling(events=True, work=False)
    pt.Render(scene, frames=4, firstFrame=1)
    st, pr = pt.Stats(), pt.Profile()
    rays = st["extension_rays"] + st["shadow_rays"]
    pt.ResetFrameNumber(); pt.SetProfiling(events=False, work=True)
    pt.Render(scene, frames=1, firstFrame=1); ctx.synchronize()
    w = pt.Profile()
    cw, aw = w["closest_work"], w["any_work"]
    sc = w["closest_sched"]
    it = max(sc["iters"], 1)
    print(f"   sched: iters/ray {32*it/cw['rays']:.1f} node lanes/iter {sc['lanes_node']/it:.1f} | tri rounds/iter {sc['tri_rounds']/it:.2f} lanes/round {sc['tri_lanes']/max(sc['tri_rounds'],1):.1f}"
          f" | setup rounds/iter {sc['setup_rounds']/it:.2f} lanes/round {sc['setup_lanes']/max(sc['setup_rounds'],1):.1f} | sphere-culled/ray {sc['sphere_culled']/cw['rays']:.2f}")
    print(f"tri={tri:2d} inst={inst:2d}: {st['device_ms']/4:7.2f} ms/frame {rays/st['device_ms']/1e3:7.1f} Mrays/s | closest {pr['trace_closest']['ms']/4:6.2f} any {pr['trace_any']['ms']/4:6.2f} shade {pr['shade']['ms']/4:5.2f}"
          f" | per ray nodes {cw['nodes']/cw['rays']:.2f} tris {cw['tris']/cw['rays']:.2f} insts {cw['insts']/cw['rays']:.2f} | shadow nodes {aw['nodes']/max(aw['rays'],1):.2f} tris {aw['tris']/max(aw['rays'],1):.2f} | mean {pt.ReadAccumulation().mean():.5f}", flush=True)
