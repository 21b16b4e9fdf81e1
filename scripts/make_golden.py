"""Generates tests/golden/*.npz on a GPU box from the UNMODIFIED reference kernels (oracle/_ref/libnexus_ref.so).

    gpurun -- 'python scripts/make_golden.py gpurun_out/golden'      then copy gpurun_out/golden/*.npz to tests/golden/

No product code runs here: inputs are procedural (numpy, fixed seeds), the host-side scene assembly is oracle_lib's numpy
restatement, outputs are what the reference's own CUDA code returns.  The reference ships no tests or golden vectors
(SURVEY.md §4), so these files are the pin for the CPU oracle and, through it and directly, for the product kernels.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from golden_cases import builder_cases, trace_scenes, trace_rays, render_cases  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)

# ---- builder: canonical BVH2 / BVH8 of the reference for every case
blob = {}
for name, prims, speed in builder_cases():
    n = prims.shape[0]
    r2, b2 = O.ref_build_bvh2(prims, speed)
    r8, rp8, b8 = O.ref_build_bvh8(prims, speed)
    c8, cp8 = O.canon_bvh8(r8, rp8)
    blob[name + "/bvh2"] = O.canon_bvh2(r2, n) if n > 1 else r2
    blob[name + "/bvh8"] = c8
    blob[name + "/prim_idx"] = cp8
    blob[name + "/bounds"] = b8
    blob[name + "/morton"] = O.ref_morton(prims, not speed)[0]     # the reference's fast-math keys (div.approx)
    print(name, n, "bvh8 nodes", len(c8), flush=True)
np.savez_compressed(os.path.join(out, "builder_ref.npz"), **blob)

# ---- traversal: closest hits of the reference TraceKernel on fixed ray batches
blob = {}
for name, desc, res in trace_scenes():
    O.ref_load_scene_standalone(desc, res)
    rays = trace_rays(name, desc, res)
    hits, _ = O.ref_trace(rays)
    blob[name + "/hits"] = hits
    print(name, len(rays), "rays,", int((hits["t"] < 1e29).sum()), "hit", flush=True)
np.savez_compressed(os.path.join(out, "trace_ref.npz"), **blob)

# ---- converged images of the reference renderer (linear accumulation, block-averaged) + its own noise floor
blob = {}
for name, desc, res, spp, block in render_cases():
    O.ref_load_scene_standalone(desc, res)
    O.ref_render(1, spp)
    a = O.ref_read_accum(res)
    O.ref_render(spp + 1, spp)          # the running mean continues: now the mean of frames 1 .. 2spp
    b = 2.0 * O.ref_read_accum(res).astype(np.float64) - a   # => mean of frames spp+1 .. 2spp, independent of `a`
    h, w = res[1] // block, res[0] // block
    da = a.reshape(h, block, w, block, 3).mean((1, 3))
    db = b.reshape(h, block, w, block, 3).mean((1, 3))
    blob[name + "/mean_a"] = da.astype(np.float32)
    blob[name + "/mean_b"] = db.astype(np.float32)
    print(name, "mean", a.mean((0, 1)), b.mean((0, 1)), "rel rmse a-b (blocks)", float(np.sqrt(((da - db) ** 2).mean()) / da.mean()), flush=True)
np.savez_compressed(os.path.join(out, "render_ref.npz"), **blob)
print("golden written to", out)
