"""Golden vectors of the reference's display transform (run on a GPU box where oracle/_ref travelled):

    gpurun -- 'python scripts/make_golden_display.py gpurun_out/golden'   # then: cp gpurun_out/golden/display_ref.npz tests/golden/

A fixed 64x64 linear HDR test image goes through the UNMODIFIED AccumulateKernel (oracle/ref/ref_harness.cu: nxref_display)
for every tone-mapping mode and three exposures; the RGBA8 render buffers are the golden outputs."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from golden_cases import display_image, DISPLAY_EXPOSURES
from nexus_b200 import scenes   # scene generators only

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
os.makedirs(out, exist_ok=True)
img = display_image()
h, w = img.shape[:2]
O.ref_load_scene_standalone(scenes.with_triangle_data(scenes.cornell_box()), (w, h))
res = {"image": img}
for mode in range(6):
    for e in DISPLAY_EXPOSURES:
        rgba = np.zeros((h, w), np.uint32)
        rc = O.ref().nxref_display(C.c_int(mode), C.c_float(e), img.ctypes.data_as(C.c_void_p), C.c_uint32(w * h), rgba.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        res[f"m{mode}_e{e:+.1f}"] = rgba
np.savez_compressed(os.path.join(out, "display_ref.npz"), **res)
print("wrote", os.path.join(out, "display_ref.npz"), {k: v.shape for k, v in res.items() if k != "image"})
