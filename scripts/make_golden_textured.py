"""Golden converged images of the reference renderer for the textured scene, same protocol and
format as scripts/make_golden.py's render section:  gpurun -- 'python scripts/make_golden_textured.py gpurun_out/golden'"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from golden_cases import textured_render_cases

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
blob = {}
for name, desc, res, spp, block in textured_render_cases():
    O.ref_load_scene_standalone(desc, res)
    O.ref_render(1, spp)
    a = O.ref_read_accum(res)
    O.ref_render(spp + 1, spp)
    b = 2.0 * O.ref_read_accum(res).astype(np.float64) - a
    h, w = res[1] // block, res[0] // block
    da = a.reshape(h, block, w, block, 3).mean((1, 3)); db = b.reshape(h, block, w, block, 3).mean((1, 3))
    blob[name + "/mean_a"] = da.astype(np.float32); blob[name + "/mean_b"] = db.astype(np.float32)
    print(name, "mean", a.mean((0, 1)), b.mean((0, 1)), "rel rmse a-b (blocks)", float(np.sqrt(((da - db) ** 2).mean()) / da.mean()), flush=True)
np.savez_compressed(os.path.join(out, "render_textured_ref.npz"), **blob)
