"""Ray-pool traversal against the lane-bound loop on a bench workload (GPU): frame time, per-kernel times and the round statistics
for a list of pool thresholds "node,tri,inst,fetch".  Usage: tune_pool.py [workload] [n,t,i,f ...]"""
import os, sys
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
if os.environ.get("NX_PATHLEN"):
    desc["settings"].pathLength = int(os.environ["NX_PATHLEN"])
scene = scenes.build(ctx, desc, res)
pt = nx.PathTracer(ctx, res)
pt.Render(scene, frames=2); ctx.synchronize()
FR = int(os.environ.get("NX_FRAMES", "4"))


def run(label, stats=True):
    pt.ResetFrameNumber(); pt.SetProfiling(events=False, work=False)
    pt.Render(scene, frames=FR, firstFrame=1)
    st = pt.Stats()
    frame_ms = st["device_ms"] / FR
    rays = st["extension_rays"] + st["shadow_rays"]
    mean = pt.ReadAccumulation().mean()
    pt.ResetFrameNumber(); pt.SetProfiling(events=True, work=False)
    pt.Render(scene, frames=2, firstFrame=1)
    pr = pt.Profile()
    line = (f"{label:28s} {frame_ms:7.2f} ms/frame {rays/st['device_ms']/1e3:7.1f} Mrays/s | serial: closest {pr['trace_closest']['ms']/2:6.2f} any {pr['trace_any']['ms']/2:6.2f}"
            f" shade {pr['shade']['ms']/2:5.2f} | mean {mean:.5f}")
    if stats:
        pt.ResetFrameNumber(); pt.SetProfiling(events=False, work=True)
        pt.Render(scene, frames=1, firstFrame=1); ctx.synchronize()
        w = pt.Profile()
        for k in ("closest", "any"):
            cw, sc = w[k + "_work"], w[k + "_sched"]
            it = max(sc["iters"], 1); nr = max(sc["node_rounds"], 1) if sc["node_rounds"] else it
            line += (f"\n      {k}: per ray nodes {cw['nodes']/max(cw['rays'],1):.2f} tris {cw['tris']/max(cw['rays'],1):.2f} insts {cw['insts']/max(cw['rays'],1):.2f} culled {sc['sphere_culled']/max(cw['rays'],1):.2f}"
                     f" | rounds {it} node {nr} ({sc['lanes_node']/nr:.1f} lanes) tri {sc['tri_rounds']} ({sc['tri_lanes']/max(sc['tri_rounds'],1):.1f}) inst/setup {sc['setup_rounds']} ({sc['setup_lanes']/max(sc['setup_rounds'],1):.1f})"
                     f" fetch {sc['fetch_rounds']} ({sc['fetch_lanes']/max(sc['fetch_rounds'],1):.1f})")
    pt.SetProfiling(events=False, work=False)
    print(line, flush=True)


# arguments: "lane:tri,inst"  "duo:tri,inst"  "pool:n,t,i,f"
specs = sys.argv[2:] or ["lane:6,8", "duo:6,8", "duo:10,10", "duo:14,12", "duo:20,16", "duo:4,6", "lane:6,8"]
for spec in specs:
    mode, _, vals = spec.partition(":")
    v = [int(x) for x in vals.split(",")] if vals else []
    ctx.SetTraceMode(mode)
    if mode == "pool":
        ctx.SetPoolTuning(*v); ctx.SetPoolTuning(*v, any_hit=True)
    elif v:
        ctx.SetTraceTuning(*v)
    run(spec, stats=True)
