"""Experiment (GPU): how much would re-ordering the secondary rays of configs[2] buy the closest-hit kernel?
Secondary rays from the primary hit points of the full 4K frame, in the renderer's queue order (8x4 pixel tiles), into random directions
of the hemisphere away from the incoming ray; timed in that order, stably binned by direction octant, sorted by a Morton code of the
origin (with and without the octant on top), and shuffled.  Prints ms per launch (device timed) and node / triangle visits per ray."""
import sys
import numpy as np
import bench
import nexus_b200 as nx
from nexus_b200 import scenes


def morton(o, bits=10):
    lo, hi = o.min(0), o.max(0)
    q = ((o - lo) / (hi - lo + 1e-9) * ((1 << bits) - 1)).astype(np.uint64)
    def spread(v):
        v = (v | (v << 32)) & 0x1f00000000ffff
        v = (v | (v << 16)) & 0x1f0000ff0000ff
        v = (v | (v << 8)) & 0x100f00f00f00f00f
        v = (v | (v << 4)) & 0x10c30c30c30c30c3
        v = (v | (v << 2)) & 0x1249249249249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "instanced10m_4k"
    ctx = nx.Context(0)
    desc = bench.make_desc(wl); res = bench.WORKLOADS[wl]["res"]
    scene = scenes.build(ctx, desc, res)
    o, d = scenes.camera_rays(desc["camera"], res)
    w, h = res
    slot = np.arange(w * h, dtype=np.uint32)
    tile = slot >> 5; ty, tx = tile // (w >> 3), tile % (w >> 3)
    pix = (ty * 4 + ((slot >> 3) & 3)) * w + tx * 8 + (slot & 7)
    primary = nx.make_rays(o[pix], d[pix])
    ph = scene.TraceClosest(primary)
    ok = np.nonzero(ph["prim"] != 0xffffffff)[0]
    rng = np.random.default_rng(5)
    p = primary["origin"][ok] + primary["direction"][ok] * ph["t"][ok, None]
    nd = rng.normal(size=(len(ok), 3)).astype(np.float32); nd /= np.linalg.norm(nd, axis=1, keepdims=True)
    flip = (nd * primary["direction"][ok]).sum(1) > 0                       # away from the incoming ray: a stand-in for the hemisphere
    nd[flip] = -nd[flip]
    sec = nx.make_rays((p + 1e-3 * nd).astype(np.float32), nd)
    octant = ((nd[:, 0] < 0).astype(np.uint64) << 2) | ((nd[:, 1] < 0).astype(np.uint64) << 1) | (nd[:, 2] < 0).astype(np.uint64)
    m = morton(sec["origin"].astype(np.float64))
    orders = {
        "queue order (8x4 tiles)": np.arange(len(sec)),
        "stable bins by octant": np.argsort(octant, kind="stable"),
        "origin Morton": np.argsort(m, kind="stable"),
        "octant, then origin Morton": np.argsort((octant << 40) | (m >> 0 & ((1 << 30) - 1)), kind="stable"),
        "origin Morton (12 bits), then octant": np.argsort(((m >> 18) << 3) | octant, kind="stable"),
        "shuffled": rng.permutation(len(sec)),
    }
    hits_dev = ctx.malloc(20 * len(sec))
    print(f"{wl}: {len(sec)} secondary rays")
    for name, idx in orders.items():
        dev = ctx.upload(sec[idx])
        ts = [scene.TraceClosestDevice(dev, len(sec), hits_dev) for _ in range(5)]
        st = scene.TraceStats(dev, len(sec), hits_dev)
        print(f"  {name:40s} {np.median(ts[1:]):7.3f} ms   nodes/ray {st['nodes'] / st['rays']:.2f}  tris/ray {st['tris'] / st['rays']:.2f}")
        ctx.free(dev)
    ctx.free(hits_dev); scene.close(); ctx.close()


if __name__ == "__main__":
    main()
