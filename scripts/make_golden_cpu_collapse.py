"""Golden vectors of the reference's own CPU collapse (Nexus/src/Geometry/BVH/BVH8Builder.cpp, compiled unmodified by
`make -C oracle refcpu`; plain g++, so this runs in the build container, no GPU).  For every case: the host BVH2 handed to the
reference (root at node 0; built by the oracle's binned-SAH builder, stored so the golden does not depend on that builder), the CWBVH8
nodes and primitive order the reference produced and its C(root, 1).  For BASELINE configs[0]'s 100,352-triangle sphere only digests
are kept (the arrays would add 4 MB).   python scripts/make_golden_cpu_collapse.py  ->  tests/golden/cpu_collapse_ref.npz"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import oracle_lib as O
from golden_cases import cpu_collapse_cases

assert O.have_refcpu(), "build oracle/_ref/libnexus_refcpu.so first: make -C oracle refcpu (needs /root/reference)"
out = {}
for name, tris, keep in cpu_collapse_cases():
    n = len(tris)
    pb, _ = O.prim_bounds(tris, 1)
    bvh2 = O.sah_build_bvh2(pb, threads=4)
    nodes, prim, cost = O.ref_cpu_collapse(bvh2, n)
    out[name + "/cost"] = np.float32(cost)
    out[name + "/node_count"] = np.uint32(len(nodes))
    out[name + "/sha256"] = np.frombuffer(hashlib.sha256(nodes.tobytes() + prim.tobytes()).digest(), np.uint8)
    out[name + "/bvh2_sha256"] = np.frombuffer(hashlib.sha256(bvh2.tobytes()).digest(), np.uint8)
    if keep:
        out[name + "/bvh2"], out[name + "/nodes"], out[name + "/prim_idx"] = bvh2, nodes, prim
    print(f"{name}: {n} triangles -> {len(nodes)} nodes, C(root, 1) = {cost:.6f}")
path = os.path.join(ROOT, "tests", "golden", "cpu_collapse_ref.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
