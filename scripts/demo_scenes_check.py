"""The reference's own demo scenes (assets/demo_scenes/*.glb) through the product and through the UNMODIFIED reference kernels
(oracle/_ref) on the same GPU: converged-image agreement and Mrays/s, scene by scene.  GPU only; the assets are not part of this
repository - stage them in an untracked directory that travels with gpurun (e.g. `cp -r /root/reference/Nexus/assets/demo_scenes
/root/repo/_demo_scenes`, listed in .gitignore) and run

    gpurun -- 'python scripts/demo_scenes_check.py _demo_scenes 64 1280x720 > gpurun_out/demo_scenes.json'

Per scene: the .glb is loaded by nexus_b200.gltf (materials, textures, node transforms), the camera is placed on a diagonal of the
scene's bounds (the files carry none), a dim sky is switched on so that scenes without emitters are not black, and `spp` frames are
rendered by both renderers; the images are compared on 8x8-pixel block means against the reference's own noise floor (two independent
halves of the reference render), exactly like tests/test_gpu_render.py::test_converged_image_matches_reference."""
import glob
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import nexus_b200 as nx
from nexus_b200 import gltf, scenes
import oracle_lib as O

src = sys.argv[1] if len(sys.argv) > 1 else "_demo_scenes"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
res = tuple(int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "1280x720").split("x"))
block = 8
ctx = nx.Context(0)
have_ref = O.have_ref() if hasattr(O, "have_ref") else os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libnexus_ref.so"))
out = []


def blocks(img):
    h, w = img.shape[0] // block, img.shape[1] // block
    return img[:h * block, :w * block].reshape(h, block, w, block, 3).mean((1, 3)).astype(np.float64)


for path in sorted(glob.glob(os.path.join(src, "**", "*.glb"), recursive=True)):
    name = os.path.basename(path)
    t0 = time.time()
    desc = gltf.load_glb(path, path_length=6)
    # world bounds from the instance matrices and the mesh boxes
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for inst in desc["instances"]:
        v = desc["meshes"][inst["mesh"]]["triangles"].reshape(-1, 3)
        b = np.array([v.min(0), v.max(0)], np.float64)
        corners = np.array([[b[(c >> a) & 1][a] for a in range(3)] for c in range(8)])
        m = inst["matrix"].astype(np.float64)
        wc = corners @ m[:3, :3].T + m[:3, 3]
        lo, hi = np.minimum(lo, wc.min(0)), np.maximum(hi, wc.max(0))
    centre, radius = 0.5 * (lo + hi), 0.5 * float(np.linalg.norm(hi - lo))
    eye = centre + radius * 1.25 * np.array([0.55, 0.35, 0.76]) / np.linalg.norm([0.55, 0.35, 0.76])
    fwd = (centre - eye) / np.linalg.norm(centre - eye)
    desc["camera"] = nx.Camera(position=tuple(eye), forward=tuple(fwd), horizontalFOV=50.0, focusDistance=float(np.linalg.norm(centre - eye)), defocusAngle=0.0)
    st = desc["settings"]; st.backgroundColor = (0.55, 0.65, 0.8); st.backgroundIntensity = 0.6
    t_load = time.time() - t0
    t0 = time.time()
    scene = scenes.build(ctx, desc, res); ctx.synchronize()
    t_build = time.time() - t0
    pt = nx.PathTracer(ctx, res)
    pt.Render(scene, frames=2); ctx.synchronize(); pt.ResetFrameNumber()
    pt.Render(scene, frames=spp, firstFrame=1)
    s = pt.Stats()
    ours = pt.ReadAccumulation()
    row = {"scene": name, "triangles": int(sum(len(m["triangles"]) for m in desc["meshes"])), "instances": len(desc["instances"]), "textures": len(desc["textures"]),
           "resolution": list(res), "spp": spp, "load_s": round(t_load, 2), "scene_setup_s": round(t_build, 2),
           "ours_ms_per_frame": round(s["device_ms"] / spp, 3), "ours_Mrays_per_s": round((s["extension_rays"] + s["shadow_rays"]) / s["device_ms"] / 1e3, 1),
           "ours_mean": [round(float(v), 5) for v in ours.mean((0, 1))], "finite": bool(np.isfinite(ours).all())}
    if have_ref:
        O.ref_load_scene(desc, scene, res)
        half = spp // 2
        O.ref_render(1, 2)                                   # warm-up of the reference's kernels on this scene
        O.ref_load_scene(desc, scene, res)                   # fresh accumulation
        ms_a, ext_a, sh_a = O.ref_render(1, half)
        a = O.ref_read_accum(res).astype(np.float64)
        ms_b, ext_b, sh_b = O.ref_render(half + 1, half)
        both = O.ref_read_accum(res).astype(np.float64)      # running mean of frames 1 .. 2 * half
        b = 2.0 * both - a
        ba, bb, bo = blocks(a), blocks(b), blocks(ours.astype(np.float64))
        ref = 0.5 * (ba + bb)
        floor = float(np.sqrt(((ba - bb) ** 2).mean()) / max(ref.mean(), 1e-30)) / np.sqrt(2.0)     # noise of the full reference render against the truth
        rmse = float(np.sqrt(((bo - ref) ** 2).mean()) / max(ref.mean(), 1e-30))
        row.update({"ref_ms_per_frame": round((ms_a + ms_b) / (2 * half), 3), "ref_Mrays_per_s": round((ext_a + sh_a + ext_b + sh_b) / (ms_a + ms_b) / 1e3, 1),
                    "ref_mean": [round(float(v), 5) for v in both.mean((0, 1))], "rel_rmse_blocks": round(rmse, 5), "ref_noise_floor": round(floor, 5),
                    "mean_ratio": round(float(ours.mean() / max(both.mean(), 1e-30)), 4)})
        row["speedup"] = round(row["ref_ms_per_frame"] / row["ours_ms_per_frame"], 3)
    out.append(row)
    print(json.dumps(row), flush=True)
    pt.close(); scene.close()
