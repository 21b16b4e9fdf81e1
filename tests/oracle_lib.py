"""ctypes binding of the CPU oracle (oracle/liboracle.so) and of the compiled reference (oracle/_ref/libnexus_ref.so).

TEST INFRASTRUCTURE: importable only from tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke().  The product package nexus_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libnexus_ref.so")

_oracle = None
_ref = None


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_bvh2_cost.restype = C.c_double
        _oracle.orc_bvh8_cost.restype = C.c_double
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


# The UNMODIFIED reference CPU collapse (Nexus/src/Geometry/BVH/BVH8Builder.cpp), oracle/Makefile target `refcpu`: plain g++, runs anywhere.
REFCPU_SO = os.path.join(os.path.dirname(REF_SO), "libnexus_refcpu.so")
_refcpu = None


def have_refcpu():
    return os.path.exists(REFCPU_SO)


def ref_cpu_collapse(bvh2_nodes, n):
    """BVH8Builder::Init + CollapseNode of the reference itself on a host BVH2 with its root at node 0.  Returns nodes8, primIdx,
    C(root, 1) — the same triple as sah_collapse, the restatement it pins."""
    global _refcpu
    if _refcpu is None:
        _refcpu = C.CDLL(REFCPU_SO)
    bvh2 = np.ascontiguousarray(bvh2_nodes)
    cap = (4 * n - 1) // 7 + 1
    nodes = np.zeros((cap, 20), np.uint32)
    prim_idx = np.zeros(n, np.uint32)
    cnt, cost = C.c_uint32(0), C.c_float(0)
    rc = _refcpu.ref_cpu_bvh8_collapse(_p(bvh2), C.c_uint32(bvh2.shape[0]), C.c_uint32(n), _p(nodes), _p(prim_idx), C.byref(cnt), C.byref(cost))
    assert rc == 0, rc
    return nodes[:cnt.value].copy(), prim_idx, cost.value


def ref():
    """The unmodified reference kernels + our headless driver (needs a GPU to *run*; loading works anywhere)."""
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
    return _ref


# ------------------------------------------------------------------ oracle: builder ----
def prim_bounds(prims, prim_type):
    prims = np.ascontiguousarray(prims, dtype=np.float32)
    n = prims.shape[0]
    bounds = np.empty((n, 6), np.float32)
    scene = np.empty(6, np.float32)
    oracle().orc_prim_bounds(_p(prims), C.c_uint32(n), C.c_int(prim_type), _p(bounds), _p(scene))
    return bounds, scene


def morton(bounds, scene, bits64):
    n = bounds.shape[0]
    out = np.empty(n, np.uint64)
    oracle().orc_morton(_p(bounds), C.c_uint32(n), _p(scene), C.c_int(int(bits64)), _p(out))
    return out


def build_bvh2(bounds, codes, bits64):
    n = bounds.shape[0]
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    rc = oracle().orc_build_bvh2(_p(np.ascontiguousarray(bounds, np.float32)), _p(np.ascontiguousarray(codes, np.uint64)),
                                 C.c_uint32(n), C.c_int(int(bits64)), _p(nodes))
    assert rc == 0, rc
    return nodes


def build_bvh8(bvh2_nodes, n):
    cap = (4 * n - 1 + 6) // 7
    nodes = np.zeros((cap, 20), np.uint32)
    prim_idx = np.zeros(n, np.uint32)
    cnt = C.c_uint32(0)
    rc = oracle().orc_build_bvh8(_p(np.ascontiguousarray(bvh2_nodes)), C.c_uint32(n), _p(nodes), _p(prim_idx), C.byref(cnt))
    assert rc == 0, rc
    return nodes[:cnt.value].copy(), prim_idx


def build_bvh8_optimal(bvh2_nodes, n, max_leaf_prims=3):
    """SAH-optimal collapse (reference CPU BVH8Builder's C(n, i) table) in the GPU builder's node conventions."""
    cap = (4 * n - 1 + 6) // 7
    nodes = np.zeros((cap, 20), np.uint32)
    prim_idx = np.zeros(n, np.uint32)
    cnt = C.c_uint32(0)
    rc = oracle().orc_build_bvh8_optimal(_p(np.ascontiguousarray(bvh2_nodes)), C.c_uint32(n), C.c_uint32(max_leaf_prims), _p(nodes), _p(prim_idx), C.byref(cnt))
    assert rc == 0, rc
    return nodes[:cnt.value].copy(), prim_idx


def canon_bvh8(nodes, prim_idx):
    nodes = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 20)
    prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
    out = np.zeros_like(nodes)
    pout = np.zeros_like(prim_idx)
    rc = oracle().orc_canon_bvh8(_p(nodes), _p(prim_idx), C.c_uint32(nodes.shape[0]), C.c_uint32(prim_idx.shape[0]), _p(out), _p(pout))
    assert rc == 0, f"canon_bvh8 failed: {rc}"
    return out, pout


def canon_bvh2(nodes, n):
    nodes = np.ascontiguousarray(nodes).view(np.uint32).reshape(-1, 8)
    out = np.zeros_like(nodes)
    rc = oracle().orc_canon_bvh2(_p(nodes), C.c_uint32(n), _p(out))
    assert rc == 0, f"canon_bvh2 failed: {rc}"
    return out


def bvh2_cost(nodes, scene):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_bvh2_cost(_p(nodes), C.c_uint32(nodes.shape[0]), _p(np.ascontiguousarray(scene, np.float32)))


def bvh8_cost(nodes, scene):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_bvh8_cost(_p(nodes), C.c_uint32(nodes.shape[0]), _p(np.ascontiguousarray(scene, np.float32)))


def check_bvh8(nodes, prim_idx, bounds):
    nodes = np.ascontiguousarray(nodes)
    return oracle().orc_check_bvh8(_p(nodes), _p(np.ascontiguousarray(prim_idx, np.uint32)), C.c_uint32(nodes.shape[0]),
                                   C.c_uint32(prim_idx.shape[0]), _p(np.ascontiguousarray(bounds, np.float32)))


# ------------------------------------------------------------------ oracle: config 1 (CPU binned SAH + SAH-optimal collapse) ----
def sah_build_bvh2(bounds, threads=1):
    """Binned-SAH BVH2, root at node 0 (oracle/oracle_sah.cpp; parity unpinned: the reference snapshot has no such source)."""
    bounds = np.ascontiguousarray(bounds, np.float32)
    n = bounds.shape[0]
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    assert oracle().orc_sah_build_bvh2(_p(bounds), C.c_uint32(n), _p(nodes), C.c_int(threads)) == 0
    return nodes


def sah_collapse(bvh2_nodes, n, root=0):
    """CPU BVH8Builder restatement (Ylitie et al. dynamic-programming collapse).  Returns nodes8, primIdx, C(root, 1)."""
    cap = (4 * n - 1) // 7 + 1
    nodes = np.zeros((cap, 20), np.uint32)
    prim_idx = np.zeros(n, np.uint32)
    cnt, cost = C.c_uint32(0), C.c_float(0)
    rc = oracle().orc_sah_collapse(_p(np.ascontiguousarray(bvh2_nodes)), C.c_uint32(n), C.c_uint32(root), _p(nodes), _p(prim_idx), C.byref(cnt), C.byref(cost))
    assert rc == 0, rc
    return nodes[:cnt.value].copy(), prim_idx, cost.value


def sah_bvh2_cost(bvh2_nodes, n):
    oracle().orc_sah_bvh2_cost.restype = C.c_double
    return oracle().orc_sah_bvh2_cost(_p(np.ascontiguousarray(bvh2_nodes)), C.c_uint32(n))


# ------------------------------------------------------------------ oracle: traversal ----
RAY_DTYPE = np.dtype([("origin", np.float32, 3), ("tmax", np.float32), ("direction", np.float32, 3), ("pad", np.uint32)])
HIT_DTYPE = np.dtype([("t", np.float32), ("u", np.float32), ("v", np.float32), ("prim", np.uint32), ("instance", np.uint32)])


class OracleScene:
    """Two-level scene for the CPU traversal oracle (oracle/oracle_trace.cpp)."""

    def __init__(self):
        L = oracle()
        L.orc_scene_create.restype = C.c_void_p
        L.orc_triangle_t.restype = C.c_float
        self._h = C.c_void_p(L.orc_scene_create())
        self.n_instances = 0

    def __del__(self):
        try:
            oracle().orc_scene_destroy(self._h)
        except Exception:
            pass

    def add_mesh(self, tris, nodes8, prim_idx):
        tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        nodes8 = np.ascontiguousarray(nodes8).view(np.uint32).reshape(-1, 20)
        prim_idx = np.ascontiguousarray(prim_idx, np.uint32)
        return oracle().orc_scene_add_mesh(self._h, _p(tris), C.c_uint32(tris.shape[0]), _p(nodes8), C.c_uint32(nodes8.shape[0]), _p(prim_idx))

    def set_instances(self, mesh_idx, inv12, tlas_nodes, tlas_prim_idx):
        mesh_idx = np.ascontiguousarray(mesh_idx, np.uint32)
        inv12 = np.ascontiguousarray(inv12, np.float32).reshape(-1, 12)
        tlas_nodes = np.ascontiguousarray(tlas_nodes).view(np.uint32).reshape(-1, 20)
        tlas_prim_idx = np.ascontiguousarray(tlas_prim_idx, np.uint32)
        self.n_instances = mesh_idx.shape[0]
        oracle().orc_scene_set_instances(self._h, _p(mesh_idx), _p(inv12), C.c_uint32(mesh_idx.shape[0]), _p(tlas_nodes),
                                         C.c_uint32(tlas_nodes.shape[0]), _p(tlas_prim_idx))

    def set_instance_ids(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32)
        assert oracle().orc_scene_set_instance_ids(self._h, _p(ids), C.c_uint32(ids.shape[0])) == 0

    def set_merged(self, mesh, inst_of, prim_of, inv12):
        inst_of, prim_of = np.ascontiguousarray(inst_of, np.uint32), np.ascontiguousarray(prim_of, np.uint32)
        inv12 = np.ascontiguousarray(inv12, np.float32).reshape(-1, 12)
        assert oracle().orc_scene_set_merged(self._h, C.c_int(mesh), _p(inst_of), _p(prim_of), C.c_uint32(inst_of.shape[0]), _p(inv12), C.c_uint32(inv12.shape[0])) == 0

    def trace_closest(self, rays, threads=8, stats=False):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        st = np.zeros(3, np.uint64)
        oracle().orc_trace_closest(self._h, _p(rays), C.c_uint32(rays.shape[0]), _p(hits), C.c_int(threads), _p(st))
        return (hits, {"nodes": int(st[0]), "tris": int(st[1]), "insts": int(st[2])}) if stats else hits

    def trace_any(self, rays, threads=8):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        occ = np.empty(rays.shape[0], np.uint8)
        oracle().orc_trace_any(self._h, _p(rays), C.c_uint32(rays.shape[0]), _p(occ), C.c_int(threads))
        return occ

    def trace_brute(self, rays, threads=8):
        rays = np.ascontiguousarray(rays, RAY_DTYPE)
        hits = np.empty(rays.shape[0], HIT_DTYPE)
        oracle().orc_trace_brute(self._h, _p(rays), C.c_uint32(rays.shape[0]), _p(hits), C.c_int(threads))
        return hits

    def triangle_t(self, ray, inst, prim):
        ray = np.ascontiguousarray(ray, RAY_DTYPE)
        return float(oracle().orc_triangle_t(self._h, _p(ray), C.c_uint32(int(inst)), C.c_uint32(int(prim))))


def cpu_build_bvh8(prims, prim_type, bits64):
    """Whole CPU pipeline: bounds -> Morton (IEEE division) -> H-PLOC -> collapse.  Returns nodes8, primIdx, scene bounds."""
    bounds, scene = prim_bounds(prims, prim_type)
    codes = morton(bounds, scene, bits64)
    n2 = build_bvh2(bounds, codes, bits64)
    n8, pidx = build_bvh8(n2, bounds.shape[0])
    return n8, pidx, scene


def compare_hits(ora, rays, got, want, rel=1e-5):
    """Classifies differences between two closest-hit arrays.  Returns dict(exact, id_mismatch_tie, id_mismatch_hard, t_bad).
    A 'tie' is an id mismatch where the other side's triangle, evaluated by the oracle for the same ray, lies within `rel`
    of the reported distance (coincident or abutting geometry; the reference itself is order-dependent there)."""
    same_id = (got["prim"] == want["prim"]) & (got["instance"] == want["instance"])
    miss_both = (got["t"] == np.float32(1e30)) & (want["t"] == np.float32(1e30))
    denom = np.maximum(np.abs(want["t"].astype(np.float64)), 1e-30)
    t_ok = (np.abs(got["t"].astype(np.float64) - want["t"].astype(np.float64)) / denom <= rel) | miss_both
    res = {"n": int(len(got)), "exact_id": int(same_id.sum()), "t_bad": int((same_id & ~t_ok).sum()), "tie": 0, "hard": 0, "hard_idx": []}
    for i in np.nonzero(~same_id)[0]:
        tg, tw = float(got["t"][i]), float(want["t"][i])
        alt = ora.triangle_t(rays[i:i + 1], want["instance"][i], want["prim"][i]) if want["prim"][i] != 0xffffffff else 1e30
        alt2 = ora.triangle_t(rays[i:i + 1], got["instance"][i], got["prim"][i]) if got["prim"][i] != 0xffffffff else 1e30
        close = lambda a, b: abs(a - b) <= rel * max(abs(a), abs(b), 1e-30)  # noqa: E731
        if close(tg, tw) or close(alt, tg) or close(alt2, tw):
            res["tie"] += 1
        else:
            res["hard"] += 1
            res["hard_idx"].append(int(i))
    return res


def grazing_margin(desc, inst_records, ray, instance, prim):
    """Float64 Moeller-Trumbore of one ray against one triangle of one instance (world space): returns (t, min(u, v, 1 - u - v), scale).
    A margin close to zero means the ray passes through an edge or vertex of the triangle.  Whether that counts as a hit is decided by
    the roundings of the fp32 evaluation - above all of the world -> object transform of the ray, whose absolute error is a few ulps
    of the COORDINATES (tens of units) while the triangle is centimetres across.  scale = (|origin| + t) / shortest edge is the factor
    that turns an ulp of a coordinate into barycentric units; two correct fp32 implementations (IEEE here, fast-math contraction in
    the reference) may disagree on a hit whose margin is below a few tens of 2^-23 * scale."""
    m = inst_records[instance, 8:72].copy().view(np.float32).reshape(4, 4).astype(np.float64)
    mesh = int(inst_records[instance, 0:4].copy().view(np.uint32)[0])
    tri = np.asarray(desc["meshes"][mesh]["triangles"], np.float64).reshape(-1, 3, 3)[prim]
    w = tri @ m[:3, :3].T + m[:3, 3]
    o = np.asarray(ray["origin"], np.float64).reshape(3); d = np.asarray(ray["direction"], np.float64).reshape(3)
    e0, e1 = w[1] - w[0], w[2] - w[0]
    pv = np.cross(d, e1); det = float(e0 @ pv)
    if det == 0.0:
        return 1e30, -1.0, 1.0
    s = o - w[0]; u = float(s @ pv) / det
    qv = np.cross(s, e0); v = float(d @ qv) / det
    t = float(e1 @ qv) / det
    edge = min(np.linalg.norm(e0), np.linalg.norm(e1), np.linalg.norm(w[2] - w[1]))
    return t, min(u, v, 1.0 - u - v), (float(np.abs(o).max()) + abs(t)) / max(edge, 1e-30)


def incidence_cos(desc, inst_records, ray, instance, prim):
    """|cos| of the angle between the ray and the triangle's plane normal (float64, world space)."""
    m = inst_records[instance, 8:72].copy().view(np.float32).reshape(4, 4).astype(np.float64)
    mesh = int(inst_records[instance, 0:4].copy().view(np.uint32)[0])
    tri = np.asarray(desc["meshes"][mesh]["triangles"], np.float64).reshape(-1, 3, 3)[prim]
    w = tri @ m[:3, :3].T + m[:3, 3]
    nrm = np.cross(w[1] - w[0], w[2] - w[0])
    d = np.asarray(ray["direction"], np.float64).reshape(3)
    return abs(float(nrm @ d)) / max(float(np.linalg.norm(nrm) * np.linalg.norm(d)), 1e-300)


def classify_hard(desc, scene, rays, got, want, hard_idx, ulps=32.0, cap=2e-2):
    """Splits 'hard' id mismatches (compare_hits) into ill-conditioned ones and real ones.  The NEARER of the two reported hits - the
    one the other side did not see - is examined in float64 (grazing_margin).  Two ways in which two correct fp32 evaluations (IEEE
    here, fast-math contraction in the reference) legitimately disagree on whether that triangle is hit:
      * edge grazing: |barycentric margin| below ulps * 2^-23 * scale, never above `cap`;
      * origin on the surface: the ray STARTS within the fp32 rounding of the triangle's plane, so the sign of t (t > 0 is a hit, t <= 0
        is behind the origin) is decided by rounding: |t| * |d| below ulps * 2^-23 * max|origin| / cos(incidence) - the transform of the
        origin into object space rounds at the magnitude of the world coordinates, and the distance along the ray amplifies that
        by 1 / cos.  (Secondary rays start 1e-3 above the surface they left and meet its neighbours at such distances.)
    Returns (edge grazing, origin on surface, real)."""
    inst = scene.ExportInstances()
    graze, origin, real = [], [], []
    for i in hard_idx:
        near = got if float(got["t"][i]) < float(want["t"][i]) else want
        t64, m, scale = grazing_margin(desc, inst, rays[i], int(near["instance"][i]), int(near["prim"][i]))
        if abs(m) <= min(cap, ulps * 2.0 ** -23 * scale):
            graze.append(int(i)); continue
        c = incidence_cos(desc, inst, rays[i], int(near["instance"][i]), int(near["prim"][i]))
        mag = float(np.abs(np.asarray(rays["origin"][i], np.float64)).max())
        dlen = float(np.linalg.norm(np.asarray(rays["direction"][i], np.float64)))
        (origin if abs(t64) * dlen <= ulps * 2.0 ** -23 * mag / max(c, 1e-3) else real).append(int(i))
    return graze, origin, real


def t_outliers(rays, got, want, rel=1e-5):
    """north_star's bar on hit distance is 1e-5 relative.  Hits whose distance is tiny compared with the coordinates involved
    (a ray starting almost on a surface) are ill-conditioned in fp32: the reference's own result is then 1e-5..1e-4 away
    from the float64 distance.  Returns (indices failing the relative bar, subset of those that also fail the conditioning
    bound |dt| <= rel * max(|t|, |origin|_inf)).  Tests require the first set to be a counted handful and the second empty."""
    same = (got["prim"] == want["prim"]) & (got["instance"] == want["instance"]) & (got["prim"] != 0xffffffff)
    dt = np.abs(got["t"].astype(np.float64) - want["t"].astype(np.float64))
    tw = np.abs(want["t"].astype(np.float64))
    bad = np.nonzero(same & (dt > rel * tw))[0]
    scale = np.maximum(tw[bad], np.abs(rays["origin"][bad].astype(np.float64)).max(axis=1))
    return bad, bad[dt[bad] > rel * scale]


# ------------------------------------------------------------------ reference arm helpers (GPU only) ----
def ref_build_bvh8(prims, prioritize_speed, metrics=False):
    prims = np.ascontiguousarray(prims, np.float32)
    n, tri = prims.shape[0], 1 if prims.shape[1] == 9 else 0
    cap = (4 * n - 1 + 6) // 7
    nodes = np.zeros((cap, 20), np.uint32)
    pidx = np.zeros(n, np.uint32)
    cnt = C.c_uint32(0)
    bounds = np.zeros(6, np.float32)
    m = np.zeros(9, np.float32)
    rc = ref().nxref_build_bvh8(_p(prims), C.c_uint32(n), C.c_int(tri), C.c_int(int(prioritize_speed)), _p(nodes), _p(pidx), C.byref(cnt),
                                _p(bounds), _p(m) if metrics else None)
    assert rc == 0
    out = (nodes[:cnt.value].copy(), pidx, bounds)
    return out + (m,) if metrics else out


def ref_morton(prims, bits64):
    """The reference's own (fast-math) Morton keys, primitive order."""
    prims = np.ascontiguousarray(prims, np.float32)
    n, tri = prims.shape[0], 1 if prims.shape[1] == 9 else 0
    out = np.zeros(n, np.uint64)
    bounds = np.zeros(6, np.float32)
    assert ref().nxref_morton(_p(prims), C.c_uint32(n), C.c_int(tri), C.c_int(int(bits64)), _p(out), _p(bounds)) == 0
    return out, bounds


def ref_build_bvh2(prims, prioritize_speed):
    prims = np.ascontiguousarray(prims, np.float32)
    n, tri = prims.shape[0], 1 if prims.shape[1] == 9 else 0
    nodes = np.zeros((2 * n - 1, 8), np.uint32)
    bounds = np.zeros(6, np.float32)
    rc = ref().nxref_build_bvh2(_p(prims), C.c_uint32(n), C.c_int(tri), C.c_int(int(prioritize_speed)), _p(nodes), _p(bounds), None)
    assert rc == 0
    return nodes, bounds


def _ref_textures(R, desc):
    for k, (pixels, srgb) in enumerate(desc.get("textures", [])):
        pixels = np.ascontiguousarray(pixels)
        idx = R.nxref_add_texture(_p(pixels), C.c_uint32(pixels.shape[1]), C.c_uint32(pixels.shape[0]), C.c_int(int(pixels.dtype == np.float32)), C.c_int(int(srgb)))
        assert idx == k, (idx, k)


def ref_load_scene(desc, scene, resolution):
    """Feeds the reference harness the same scene the product got: identical triangles, shading data, materials, and the
    instance matrices / camera / light list exported by the product's host layer in the reference's device layouts."""
    R = ref()
    R.nxref_scene_reset()
    for m in desc["meshes"]:
        tris = np.ascontiguousarray(m["triangles"], np.float32)
        td = np.ascontiguousarray(m["triangle_data"], np.float32)
        assert R.nxref_add_mesh(_p(tris), _p(td), C.c_uint32(tris.shape[0])) >= 0
    inst = scene.ExportInstances()
    assert R.nxref_set_instances(_p(inst), C.c_uint32(inst.shape[0])) == 0
    _ref_textures(R, desc)
    mats = np.frombuffer(b"".join(bytes(m.pod()) for m in desc["materials"]), np.uint8).copy()
    assert R.nxref_set_materials(_p(mats), C.c_uint32(len(desc["materials"]))) == 0
    lights = scene.ExportLights()
    assert R.nxref_set_lights(_p(lights) if len(lights) else None, C.c_uint32(len(lights))) == 0
    cam = scene.ExportCamera()
    assert R.nxref_set_camera(_p(cam)) == 0
    st = desc["settings"]
    bg = np.asarray(st.backgroundColor, np.float32)
    assert R.nxref_set_settings(C.c_int(int(st.useMIS)), C.c_int(int(st.pathLength)), _p(bg), C.c_float(st.backgroundIntensity)) == 0
    if desc.get("hdr") is not None:
        hdr = np.ascontiguousarray(desc["hdr"], np.float32)
        assert R.nxref_set_hdr(_p(hdr), C.c_uint32(hdr.shape[1]), C.c_uint32(hdr.shape[0])) == 0
    assert R.nxref_render_init(C.c_uint32(resolution[0]), C.c_uint32(resolution[1])) == 0


def ref_render(first_frame, n_frames):
    ms = C.c_float(0)
    rays = (C.c_ulonglong * 2)()
    assert ref().nxref_render(C.c_uint32(first_frame), C.c_uint32(n_frames), C.byref(ms), rays) == 0
    return ms.value, int(rays[0]), int(rays[1])


def ref_read_accum(resolution):
    out = np.empty((resolution[1], resolution[0], 3), np.float32)
    assert ref().nxref_read_accum(_p(out)) == 0
    return out


def ref_trace(rays):
    rays = np.ascontiguousarray(rays, RAY_DTYPE)
    n = rays.shape[0]
    o = np.ascontiguousarray(rays["origin"])
    d = np.ascontiguousarray(rays["direction"])
    t, u, v = (np.empty(n, np.float32) for _ in range(3))
    tri, inst = np.empty(n, np.uint32), np.empty(n, np.uint32)
    ms = C.c_float(0)
    rc = ref().nxref_trace(_p(o), _p(d), C.c_uint32(n), _p(t), _p(u), _p(v), _p(tri), _p(inst), C.byref(ms))
    assert rc == 0, rc
    hits = np.empty(n, HIT_DTYPE)
    hits["t"], hits["u"], hits["v"], hits["prim"], hits["instance"] = t, u, v, tri, inst
    miss = hits["t"] == np.float32(1e30)
    hits["prim"][miss] = 0xffffffff
    hits["instance"][miss] = 0xffffffff
    return hits, ms.value


class ProductOracle:
    """CPU oracle over the acceleration structures the PRODUCT built (so traversal, not building, is under test), answering in the
    product's terms: instance ids and primitive ids inside the instance's mesh.

    The product's TLAS is over entries - instances with a BLAS of their own, then the merged BLAS (nx_scene_export_merged: nodes in
    world space, entered with the identity transform, every triangle tested in the object space of its instance).  `trav` mirrors
    exactly that (same trees, same arithmetic: bit-identical hits); `plain` is the object-space scene (every instance with its own
    mesh, no trees needed) for brute force and single-triangle evaluation."""

    def __init__(self, desc, scene):
        inst = scene.ExportInstances()
        self.inst_records = inst
        mesh_idx = inst[:, 0:4].copy().view(np.uint32).ravel()
        inv = inst[:, 72:136].copy().view(np.float32).reshape(-1, 16)[:, :12]
        entries = scene.ExportTlasEntries()
        merged = scene.ExportMerged(bounds=False) if (entries == 0xffffffff).any() else None
        tn, tp = scene.TLAS().ToHost()
        T = OracleScene()
        slot_of_mesh, e_mesh, e_inv, added, merged_slot = {}, [], [], 0, None
        for e in entries:
            if e == 0xffffffff:
                nodes, pidx = merged["bvh"].ToHost()
                # object-space triangles in merged order: instance by instance, each with its mesh's triangles in order
                tris = np.empty((len(merged["prim"]), 9), np.float32)
                first = np.nonzero(merged["prim"] == 0)[0]
                for f, l in zip(first, list(first[1:]) + [len(tris)]):
                    tris[f:l] = desc["meshes"][int(mesh_idx[merged["instance"][f]])]["triangles"][:l - f]
                T.add_mesh(tris, nodes, pidx)
                merged_slot = added
                e_mesh.append(added); added += 1
                e_inv.append(np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32))
                continue
            k = int(mesh_idx[e])
            if k not in slot_of_mesh:
                nodes, pidx = scene.MeshBVH(k).ToHost()
                T.add_mesh(desc["meshes"][k]["triangles"], nodes, pidx)
                slot_of_mesh[k] = added; added += 1
            e_mesh.append(slot_of_mesh[k]); e_inv.append(inv[e])
        T.set_instances(np.asarray(e_mesh, np.uint32), np.ascontiguousarray(np.stack(e_inv), np.float32), tn, tp)
        T.set_instance_ids(entries)                       # hits report scene instance ids (the merged BLAS reports per triangle)
        if merged_slot is not None:
            T.set_merged(merged_slot, merged["instance"], merged["prim"], inv)
        self.trav, self.entries, self.merged = T, entries, merged
        # object-space scene for brute force: one dummy node per mesh is enough, the BVHs are never walked
        P = OracleScene()
        dummy = np.zeros((1, 20), np.uint32)
        for m in desc["meshes"]:
            P.add_mesh(m["triangles"], dummy, np.arange(len(m["triangles"]), dtype=np.uint32))
        P.set_instances(mesh_idx, inv, dummy, np.arange(len(mesh_idx), dtype=np.uint32))
        self.plain = P

    def trace_closest(self, rays, threads=8, stats=False):
        return self.trav.trace_closest(rays, threads=threads, stats=stats)

    def trace_any(self, rays, threads=8):
        return self.trav.trace_any(rays, threads=threads)

    def trace_brute(self, rays, threads=8):
        return self.plain.trace_brute(rays, threads=threads)

    def triangle_t(self, ray, inst, prim):
        return self.plain.triangle_t(ray, inst, prim)


def oracle_scene_from_product(desc, scene):
    """CPU oracle scene that uses the BVHs the product built on the GPU (so traversal, not building, is under test)."""
    return ProductOracle(desc, scene)


def _trs(position, rotation_deg, scale):
    rx, ry, rz = (np.radians(np.float64(a)) for a in rotation_deg)
    T = np.eye(4); T[:3, 3] = position
    S = np.diag([scale[0], scale[1], scale[2], 1.0])
    Rx = np.eye(4); Rx[1, 1], Rx[1, 2], Rx[2, 1], Rx[2, 2] = np.cos(rx), -np.sin(rx), np.sin(rx), np.cos(rx)
    Ry = np.eye(4); Ry[0, 0], Ry[0, 2], Ry[2, 0], Ry[2, 2] = np.cos(ry), np.sin(ry), -np.sin(ry), np.cos(ry)
    Rz = np.eye(4); Rz[0, 0], Rz[0, 1], Rz[1, 0], Rz[1, 1] = np.cos(rz), -np.sin(rz), np.sin(rz), np.cos(rz)
    return T @ Rz @ Ry @ Rx @ S


def camera_record(cam, resolution):
    """D_Camera (88 B) from Camera ctor arguments."""
    w, h = resolution
    f = np.asarray(cam.forward, np.float64)
    r = np.asarray(cam.right, np.float64)
    if not r.any():
        r = np.cross(f, [0.0, 1.0, 0.0])
    up = np.cross(r, f)
    half_w = cam.focusDistance * np.tan(np.radians(cam.horizontalFOV / 2.0))
    half_h = half_w / (w / h)
    lens = cam.focusDistance * np.tan(np.radians(cam.defocusAngle / 2.0))
    pos = np.asarray(cam.position, np.float64)
    vx, vy = 2 * half_w * r, 2 * half_h * up
    ll = pos - vx / 2 - vy / 2 + f * cam.focusDistance
    rec = np.zeros(22, np.float32)
    rec[0:3], rec[3:6], rec[6:9], rec[9] = pos, r, up, lens
    rec[10:13], rec[13:16], rec[16:19] = ll, vx, vy
    out = rec.view(np.uint8).copy()
    out[80:88] = np.array([w, h], np.uint32).view(np.uint8)
    return out


def host_instances(desc, mesh_bounds):
    """D_MeshInstance records (n, 160) uint8 + the material index of every instance."""
    inst = np.zeros((len(desc["instances"]), 160), np.uint8)
    mats_of = []
    for k, i in enumerate(desc["instances"]):
        M = _trs(i["position"], i["rotation"], i["scale"])
        Mi = np.linalg.inv(M)
        b = mesh_bounds[i["mesh"]]
        corners = np.array([[b[0 + 3 * (c & 1)], b[1 + 3 * ((c >> 1) & 1)], b[2 + 3 * ((c >> 2) & 1)], 1.0] for c in range(8)])
        wc = (M.astype(np.float32).astype(np.float64) @ corners.T).T[:, :3]
        mat = i.get("material", -1)
        mat = desc["meshes"][i["mesh"]]["material"] if mat < 0 else mat
        mats_of.append(mat)
        inst[k, 0:8] = np.array([i["mesh"], mat], np.uint32).view(np.uint8)
        inst[k, 8:72] = M.astype(np.float32).ravel().view(np.uint8)
        inst[k, 72:136] = Mi.astype(np.float32).ravel().view(np.uint8)
        inst[k, 136:160] = np.concatenate([wc.min(0), wc.max(0)]).astype(np.float32).view(np.uint8)
    return inst, mats_of


def oracle_scene_from_desc(desc):
    """CPU-only two-level scene: every BLAS and the TLAS built by the CPU oracle pipeline (IEEE Morton keys), instance
    records from the numpy host restatement.  No GPU, no product code."""
    S = OracleScene()
    mesh_bounds = []
    for m in desc["meshes"]:
        n8, pidx, sb = cpu_build_bvh8(m["triangles"], 1, 0)       # Mesh::Mesh: prioritizeSpeed = true -> 32-bit keys
        S.add_mesh(m["triangles"], n8, pidx)
        mesh_bounds.append(sb)
    inst, _ = host_instances(desc, mesh_bounds)
    bounds = inst[:, 136:160].copy().view(np.float32).reshape(-1, 6)
    tn, tp, _ = cpu_build_bvh8(bounds, 0, 1)                      # Scene::BuildTLAS: default config -> 64-bit keys
    S.set_instances(inst[:, 0:4].copy().view(np.uint32).ravel(), inst[:, 72:136].copy().view(np.float32).reshape(-1, 16)[:, :12], tn, tp)
    S.instances = inst
    return S


def ref_load_scene_standalone(desc, resolution):
    """Loads a scene description into the reference harness without touching the product library."""
    R = ref()
    R.nxref_scene_reset()
    mesh_bounds = []
    for m in desc["meshes"]:
        tris = np.ascontiguousarray(m["triangles"], np.float32)
        td = np.ascontiguousarray(m["triangle_data"], np.float32)
        idx = R.nxref_add_mesh(_p(tris), _p(td), C.c_uint32(tris.shape[0]))
        assert idx >= 0
        b = np.zeros(6, np.float32)
        R.nxref_mesh_bounds(C.c_int(idx), _p(b))
        mesh_bounds.append(b)
    inst, mats_of = host_instances(desc, mesh_bounds)
    assert R.nxref_set_instances(_p(inst), C.c_uint32(inst.shape[0])) == 0
    _ref_textures(R, desc)
    mats = np.frombuffer(b"".join(bytes(m.pod()) for m in desc["materials"]), np.uint8).copy()
    assert R.nxref_set_materials(_p(mats), C.c_uint32(len(desc["materials"]))) == 0
    lights = []
    for mi, m in enumerate(desc["materials"]):
        if max(m.emissionColor) > 0.0 and m.intensity > 0.0:
            for k, mo in enumerate(mats_of):
                if mo == mi:
                    rec = np.zeros(52, np.uint8)
                    rec[0:4] = np.array([k], np.uint32).view(np.uint8)
                    rec[48] = 3
                    lights.append(rec)
    lights = np.stack(lights) if lights else np.zeros((0, 52), np.uint8)
    assert R.nxref_set_lights(_p(lights) if len(lights) else None, C.c_uint32(len(lights))) == 0
    assert R.nxref_set_camera(_p(camera_record(desc["camera"], resolution))) == 0
    st = desc["settings"]
    bg = np.asarray(st.backgroundColor, np.float32)
    assert R.nxref_set_settings(C.c_int(int(st.useMIS)), C.c_int(int(st.pathLength)), _p(bg), C.c_float(st.backgroundIntensity)) == 0
    if desc.get("hdr") is not None:
        hdr = np.ascontiguousarray(desc["hdr"], np.float32)
        assert R.nxref_set_hdr(_p(hdr), C.c_uint32(hdr.shape[1]), C.c_uint32(hdr.shape[0])) == 0
    assert R.nxref_render_init(C.c_uint32(resolution[0]), C.c_uint32(resolution[1])) == 0
