"""Deterministic inputs shared by scripts/make_golden.py (which runs the reference on them) and the parity tests."""
import numpy as np


def _rand_tris(rng, n, spread=5.0, size=0.1):
    c = rng.uniform(-spread, spread, (n, 1, 3)).astype(np.float32)
    return (c + rng.uniform(-size, size, (n, 3, 3)).astype(np.float32)).reshape(n, 9)


def builder_cases():
    """(name, prims (n,9) triangles or (n,6) AABBs, prioritizeSpeed).  Covers single/tiny inputs, the PLOC merge threshold
    (16/17/33), duplicates (equal Morton keys and equal areas), a degenerate flat cloud, AABB primitives, a closed mesh."""
    from nexus_b200 import scenes
    rng = np.random.default_rng(20261017)
    cases = []
    for n in (1, 2, 3, 16, 17, 33, 100, 1000, 5000):
        t = _rand_tris(rng, n)
        cases.append((f"tri{n}_m32", t, True))
        cases.append((f"tri{n}_m64", t, False))
    lo = rng.uniform(-3, 3, (50, 3)).astype(np.float32)
    cases.append(("aabb50_m64", np.concatenate([lo, lo + rng.uniform(0.01, 0.5, (50, 3)).astype(np.float32)], 1), False))
    cases.append(("aabb50_m32", cases[-1][1], True))
    cases.append(("dup40_m32", np.tile(_rand_tris(rng, 1), (40, 1)), True))
    flat = _rand_tris(rng, 500); flat[:, 2::3] = 0.0
    cases.append(("flat500_m32", flat, True))
    cases.append(("sphere2048_m32", scenes.uv_sphere(32, 32), True))
    cases.append(("sphere8192_m32", scenes.uv_sphere(64, 64), True))
    # lattice with a power-of-two extent: Morton normalisation is exact in both IEEE and approximate division
    g = rng.integers(0, 256, (3000, 1, 3)).astype(np.float32) / 8.0
    lat = (g + rng.integers(0, 5, (3000, 3, 3)).astype(np.float32) / 16.0).reshape(3000, 9)
    lat[0, :3] = 0.0; lat[1, :3] = 32.0
    cases.append(("lattice3000_m32", lat, True))
    cases.append(("lattice3000_m64", lat, False))
    return cases


def trace_scenes():
    from nexus_b200 import scenes
    return [
        ("cornell", scenes.with_triangle_data(scenes.cornell_box()), (128, 128)),
        ("instanced", scenes.with_triangle_data(scenes.instanced_scene(n_blas=6, n_instances=20, nu=16, nv=14)), (128, 128)),
    ]


def trace_rays(name, desc, res):
    """Camera rays + rays from random points in random directions (inside the scene's bounding region)."""
    import nexus_b200 as nx
    from nexus_b200 import scenes
    rng = np.random.default_rng(7 + len(name))
    o, d = scenes.camera_rays(desc["camera"], res, n=6000, seed=3)
    lo, hi = np.array([1e30] * 3), np.array([-1e30] * 3)
    for m in desc["meshes"][:2] if name == "instanced" else desc["meshes"]:
        v = m["triangles"].reshape(-1, 3)
        lo, hi = np.minimum(lo, v.min(0)), np.maximum(hi, v.max(0))
    if name == "instanced":
        lo, hi = np.array([-8.0, 0.1, -8.0]), np.array([8.0, 4.0, 8.0])
    ro = rng.uniform(lo, hi, (6000, 3)).astype(np.float32)
    rd = rng.normal(size=(6000, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    return nx.make_rays(np.concatenate([o, ro]), np.concatenate([d, rd]))


def render_cases():
    """(name, desc, resolution, spp, block): converged block means are compared, not pixels (the reference's RNG is keyed
    on racing queue slots, so its frames are not reproducible run to run — SURVEY.md §3.1)."""
    from nexus_b200 import scenes
    return [
        ("cornell", scenes.with_triangle_data(scenes.cornell_box(path_length=6)), (128, 128), 2048, 8),
        ("instanced", scenes.with_triangle_data(scenes.instanced_scene(n_blas=6, n_instances=20, nu=16, nv=14, path_length=6)), (128, 72), 2048, 8),
    ]


def textured_render_cases():
    """SURVEY.md §8 row f-2 (material maps): same protocol as render_cases(), golden file render_textured_ref.npz."""
    from nexus_b200 import scenes
    return [("cornell_textured", scenes.textured_cornell(path_length=6), (128, 128), 2048, 8)]


# ------------------------------------------------------------------ display transform (tests/golden/display_ref.npz) ----
DISPLAY_EXPOSURES = (0.0, -1.5, 2.0)


def display_image():
    """64x64 linear HDR test image: log-uniform radiance over 1e-4..60 per channel, a grey ramp row, exact zeros and ones."""
    rng = np.random.default_rng(3)
    img = np.exp(rng.uniform(np.log(1e-4), np.log(60.0), (64, 64, 3))).astype(np.float32)
    img[0] = (np.linspace(0.0, 4.0, 64, dtype=np.float32) ** 2)[:, None]
    img[1, :8] = 0.0
    img[1, 8:16] = 1.0
    return img


def cpu_collapse_cases():
    """(name, triangles (n, 9), keep arrays in the golden file) for the reference CPU collapse golden (scripts/make_golden_cpu_collapse.py).
    No input with a zero-extent axis: there the reference evaluates log2f(0) and casts NaN to uint8_t (BVH8Builder.cpp:305-307,
    346-353), which is undefined behaviour; the restatement gives such an axis the smallest normal cell instead (oracle_sah.cpp)."""
    from nexus_b200 import scenes
    rng = np.random.default_rng(2024)
    cases = [("sphere32", np.asarray(scenes.uv_sphere(32, 32), np.float32).reshape(-1, 9), True),
             ("rock912", np.asarray(scenes.rock(3, 24, 20), np.float32).reshape(-1, 9), True),
             ("soup1", _rand_tris(rng, 1), True), ("soup2", _rand_tris(rng, 2), True), ("soup3", _rand_tris(rng, 3), True),
             ("soup9", _rand_tris(rng, 9), True), ("soup500", _rand_tris(rng, 500), True), ("soup1900", _rand_tris(rng, 1900, spread=2.0, size=0.3), True),
             ("rock9798", np.asarray(scenes.rock(11), np.float32).reshape(-1, 9), False),
             ("sphere224", np.asarray(scenes.uv_sphere(224, 224), np.float32).reshape(-1, 9), False)]
    return cases


def host_cases():
    """Inputs of the host-record golden (scripts/make_golden_host.py): instance transforms (position, Euler degrees, scale, mesh box,
    ids) and cameras (position, forward, horizontal FOV, focus distance, defocus angle, resolution)."""
    rng = np.random.default_rng(77)
    n = 48
    inst = {"position": rng.uniform(-20, 20, (n, 3)).astype(np.float32), "rotation": rng.uniform(-360, 720, (n, 3)).astype(np.float32),
            "scale": rng.uniform(0.05, 4.0, (n, 3)).astype(np.float32),
            "mesh_bounds": np.concatenate([rng.uniform(-3, 0, (n, 3)), rng.uniform(0, 3, (n, 3))], 1).astype(np.float32),
            "mesh_idx": rng.integers(0, 1000, n).astype(np.uint32), "material_idx": rng.integers(0, 50, n).astype(np.uint32)}
    inst["position"][0] = 0; inst["rotation"][0] = 0; inst["scale"][0] = 1                      # identity
    inst["rotation"][1] = (90, 0, 0); inst["rotation"][2] = (0, 90, 0); inst["rotation"][3] = (0, 0, 90)   # one axis at a time
    inst["scale"][4] = (-1, 2, 0.5)                                                             # mirrored
    inst["scale"][5] = (1.5, 1.5, 1.5); inst["rotation"][5] = (10, 200, 35)
    m = 12
    fwd = rng.normal(size=(m, 3)); fwd[:, 1] *= 0.3
    fwd /= np.linalg.norm(fwd, axis=1, keepdims=True)
    cams = {"position": rng.uniform(-10, 10, (m, 3)).astype(np.float32), "forward": fwd.astype(np.float32),
            "hfov": rng.uniform(20, 100, m).astype(np.float32), "focus": rng.uniform(0.5, 20, m).astype(np.float32),
            "defocus": rng.uniform(0, 5, m).astype(np.float32),
            "res": np.array([(1920, 1080), (3840, 2160), (640, 480), (96, 64)] * 3, np.uint32)}
    cams["position"][0] = (0, 4, 14); cams["forward"][0] = (0, 0, -1); cams["hfov"][0] = 45; cams["focus"][0] = 5; cams["defocus"][0] = 0   # Scene.cpp:9-10
    return inst, cams


def bsdf_cases(n=8192, seed=5):
    """Random (material, wi, wo) triples for the BSDF golden (scripts/make_golden_bsdf.py): nx_material / D_Material records (92 B as 23
    words) covering metal / dielectric / plastic mixes, anisotropy, specular weight and colour, ior, transmission; directions in the
    local shading frame, wi mostly in the upper hemisphere, wo in both (reflection and refraction)."""
    rng = np.random.default_rng(seed)
    mat = np.zeros((n, 23), np.float32)
    mat[:, 0:3] = rng.uniform(0.02, 1, (n, 3))
    mat[:, 3] = rng.choice([0.0, 1.0, 0.3, 0.7], n)             # metalness
    mat[:, 4] = rng.uniform(0.03, 1.0, n)                       # roughness
    mat[:, 5] = rng.choice([0.0, 0.0, 0.5, 0.9], n)             # anisotropy
    mat[:, 6] = rng.choice([1.0, 0.0, 0.5], n)                  # specular weight
    mat[:, 7:10] = rng.uniform(0.2, 1, (n, 3))                  # specular colour
    mat[:, 10] = rng.uniform(1.05, 2.2, n)                      # ior
    mat[:, 11] = rng.choice([0.0, 1.0, 0.4], n)                 # transmission
    mat[:, 12:15] = 1; mat[:, 15] = 0; mat[:, 16] = 1           # emission colour, intensity, opacity
    mat.view(np.int32)[:, 17:23] = -1                           # no maps

    def dirs(lower_frac):
        v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
        v[:, 2] = np.abs(v[:, 2]); v[rng.uniform(size=n) < lower_frac, 2] *= -1
        return v.astype(np.float32)
    return mat, dirs(0.15), dirs(0.4)
