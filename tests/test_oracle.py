"""CPU tests of the oracle (oracle/*.cpp, test infrastructure) against the committed outputs of the unmodified reference
kernels (tests/golden/, see its README), plus structural invariants the reference implies.  No GPU."""
import os

import numpy as np
import pytest

import oracle_lib as O
from golden_cases import builder_cases, trace_rays, trace_scenes

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = builder_cases()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "builder_ref.npz"))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_builder_oracle_equals_reference_trees(gold, case):
    """Given the reference's own (fast-math) Morton keys, the CPU H-PLOC + collapse restatement reproduces the reference's
    BVH2 and CWBVH8 bit for bit (canonical numbering)."""
    name, prims, speed = case
    n, tri, b64 = prims.shape[0], 1 if prims.shape[1] == 9 else 0, 0 if speed else 1
    pb, sb = O.prim_bounds(prims, tri)
    assert (sb == gold[name + "/bounds"]).all()
    o2 = O.build_bvh2(pb, gold[name + "/morton"], b64)
    if n > 1:
        assert (O.canon_bvh2(o2, n) == gold[name + "/bvh2"]).all()
    o8, op8 = O.build_bvh8(o2, n)
    c8, cp8 = O.canon_bvh8(o8, op8)
    assert c8.shape == gold[name + "/bvh8"].shape and (c8 == gold[name + "/bvh8"]).all()
    assert (cp8 == gold[name + "/prim_idx"]).all()
    assert O.check_bvh8(o8, op8, pb) == 0


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_builder_oracle_pure_cpu_pipeline(gold, case):
    """Same with the oracle's own IEEE-division Morton keys: 32-bit keys agree with the GPU's exactly on every golden case;
    64-bit keys may differ in the last quantisation cell for a few primitives (div.approx), counted here, and on these
    cases never change the tree."""
    name, prims, speed = case
    n, tri, b64 = prims.shape[0], 1 if prims.shape[1] == 9 else 0, 0 if speed else 1
    pb, sb = O.prim_bounds(prims, tri)
    codes = O.morton(pb, sb, b64)
    mism = int((codes != gold[name + "/morton"]).sum())
    assert mism == 0 if not b64 else mism <= 0.3 * n + 4, mism   # 21-bit cells vs a 2-ulp approximate division
    n8, pidx, _ = O.cpu_build_bvh8(prims, tri, b64)
    c8, cp8 = O.canon_bvh8(n8, pidx)
    assert c8.shape == gold[name + "/bvh8"].shape and (c8 == gold[name + "/bvh8"]).all() and (cp8 == gold[name + "/prim_idx"]).all()


def test_sort_contract_and_bvh2_shape():
    """Keys are sorted on bits [2,32) / [1,64) only, stably (Setup.cu:74-78); BVH2 = 2n-1 nodes, leaves first, root last."""
    rng = np.random.default_rng(3)
    import ctypes as C
    codes = rng.integers(0, 1 << 30, 1000, dtype=np.uint64)
    codes[100:200] = codes[100] ^ np.arange(100, dtype=np.uint64) % 4          # equal after dropping the two low bits
    out_c, out_i = np.zeros(1000, np.uint64), np.zeros(1000, np.uint32)
    O.oracle().orc_sort(O._p(codes), C.c_uint32(1000), C.c_int(0), O._p(out_c), O._p(out_i))
    key = out_c >> np.uint64(2)
    assert (np.diff(key.astype(np.int64)) >= 0).all()
    same = np.nonzero(np.diff(key.astype(np.int64)) == 0)[0]
    assert (out_i[same + 1] > out_i[same]).all()                                # stable
    n = 777
    c = rng.uniform(-3, 3, (n, 1, 3)).astype(np.float32)
    prims = (c + rng.uniform(-0.1, 0.1, (n, 3, 3)).astype(np.float32)).reshape(n, 9)
    pb, sb = O.prim_bounds(prims, 1)
    n2 = O.build_bvh2(pb, O.morton(pb, sb, 0), 0)
    leaf = n2[:, 6] == 0xffffffff
    assert leaf[:n].all() and not leaf[n:].any() and (n2[:n, 7] == np.arange(n)).all()
    kids = np.concatenate([n2[n:, 6], n2[n:, 7]])
    assert len(np.unique(kids)) == 2 * n - 2 and 2 * n - 2 not in kids
    f = n2.view(np.float32)
    for i in range(n, 2 * n - 1):                                               # parent box = union of child boxes
        l, r = n2[i, 6], n2[i, 7]
        assert (f[i, :3] == np.minimum(f[l, :3], f[r, :3])).all() and (f[i, 3:6] == np.maximum(f[l, 3:6], f[r, 3:6])).all()
    c2 = O.canon_bvh2(n2, n)
    assert (O.canon_bvh2(c2, n) == c2).all()                                    # canonical form is a fixed point


def test_bvh8_invariants_on_config1_mesh():
    """BASELINE.json configs[0]: the ~100K-triangle tessellated mesh, host only."""
    from nexus_b200 import scenes
    prims = scenes.uv_sphere(224, 224)
    n = prims.shape[0]
    assert n == 100352
    n8, pidx, sb = O.cpu_build_bvh8(prims, 1, 0)
    assert len(n8) <= (4 * n - 1 + 6) // 7
    assert (np.sort(pidx) == np.arange(n)).all()
    pb, _ = O.prim_bounds(prims, 1)
    assert O.check_bvh8(n8, pidx, pb) == 0
    c8, cp = O.canon_bvh8(n8, pidx)
    assert (O.canon_bvh8(c8, cp)[0] == c8).all()
    cost = O.bvh8_cost(n8, sb)
    assert 10.0 < cost < 1000.0


@pytest.mark.parametrize("case", trace_scenes(), ids=lambda c: c[0])
def test_trace_oracle_equals_reference_hits(case):
    """CPU two-level traversal (own CPU-built BVHs) vs the reference TraceKernel's hits on the same rays: primitive and
    instance ids equal except exact-distance ties, t within 1e-5 relative except a counted handful of ill-conditioned hits
    (see oracle_lib.t_outliers) that must still meet the conditioning bound; brute force agrees with the BVH traversal."""
    name, desc, res = case
    gold = np.load(os.path.join(GOLD, "trace_ref.npz"))[name + "/hits"]
    rays = trace_rays(name, desc, res)
    ora = O.oracle_scene_from_desc(desc)
    got = ora.trace_closest(rays)
    cmp = O.compare_hits(ora, rays, got, gold, rel=1e-5)
    assert cmp["hard"] == 0 and cmp["tie"] <= 0.001 * cmp["n"], cmp
    bad, worse = O.t_outliers(rays, got, gold, rel=1e-5)
    assert len(bad) <= 1e-3 * len(rays) and len(worse) == 0, (len(bad), len(worse))
    brute = ora.trace_brute(rays[:1500])
    cmp2 = O.compare_hits(ora, rays[:1500], got[:1500], brute, rel=1e-6)
    assert cmp2["hard"] == 0 and cmp2["t_bad"] == 0, cmp2
    occ = ora.trace_any(np.array(rays))
    assert (occ.astype(bool) == (got["t"] < np.float32(1e30))).all()


# ------------------------------------------------------------------ display transform ----
@pytest.mark.parametrize("mode", range(6), ids=["none", "aces", "uncharted2", "agx", "agx_golden", "agx_punchy"])
def test_display_oracle_equals_reference_render_buffer(mode):
    """The numpy restatement of AccumulateKernel's display transform (oracle/oracle_display.py) against the RGBA8 buffers the
    unmodified reference kernel produced for the same image (tests/golden/display_ref.npz).  Integer output: exact for the
    rational curves; the AgX path goes through log2 / pow (fast-math in the reference), where a value that lands on a
    quantisation boundary may differ by one code: at most 1 LSB on at most 0.2 % of the channels."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "oracle"))
    import oracle_display as D
    from golden_cases import DISPLAY_EXPOSURES, display_image
    g = np.load(os.path.join(GOLD, "display_ref.npz"))
    assert (g["image"] == display_image()).all()
    for e in DISPLAY_EXPOSURES:
        ref, got = D.unpack(g[f"m{mode}_e{e:+.1f}"]), D.unpack(D.display(g["image"], mode, e))
        diff = np.abs(ref - got)
        if mode <= D.UNCHARTED2:
            assert diff.max() == 0, (mode, e)
        else:
            assert diff.max() <= 1 and (diff > 0).mean() <= 0.002, (mode, e, diff.max(), (diff > 0).mean())
        assert (g[f"m{mode}_e{e:+.1f}"] >> 24 == 0xff).all()


# ------------------------------------------------------------------ BASELINE.json configs[0]: CPU binned SAH + SAH-optimal collapse ----
def test_config1_cpu_sah_build_and_optimal_collapse():
    """configs[0]: the procedurally tessellated 100,352-triangle sphere through the CPU path (oracle/oracle_sah.cpp): binned-SAH
    BVH2 (written from the README's description, PARITY UNPINNED: the reference snapshot has no such source) and the
    restatement of the reference's CPU BVH8Builder (Ylitie et al. collapse).  No golden vectors exist for either, so the
    checks are the invariants the algorithms imply: a complete binary tree with every primitive in exactly one leaf and
    parents enclosing children; a CWBVH8 whose decoded child boxes enclose their subtrees, leaves of at most P_MAX = 3
    triangles, at most 24 triangles and 8 children per node; C(root, 1) no worse than the greedy collapse of the same tree
    evaluated with the same cost model; and traversal of the result equal to brute force."""
    from nexus_b200 import scenes
    tris = scenes.uv_sphere(224, 224).reshape(-1, 9)
    n = tris.shape[0]
    assert n == 100352
    pb, sb = O.prim_bounds(tris, 1)
    n2 = O.sah_build_bvh2(pb, threads=4)
    assert (n2 == O.sah_build_bvh2(pb, threads=1)).all()                     # numbering does not depend on the thread count
    leaf = n2[:, 6] == 0xffffffff
    assert leaf.sum() == n and (np.sort(n2[leaf, 7]) == np.arange(n)).all()
    inner = np.nonzero(~leaf)[0]
    kids = np.concatenate([n2[inner, 6], n2[inner, 7]])
    assert len(np.unique(kids)) == 2 * n - 2 and 0 not in kids             # root = node 0, every other node has one parent
    f = n2.view(np.float32)
    for side in (6, 7):
        c = n2[inner, side]
        assert (f[c, 0:3] >= f[inner, 0:3]).all() and (f[c, 3:6] <= f[inner, 3:6]).all()
    n8, pidx, root_cost = O.sah_collapse(n2, n)
    assert O.check_bvh8(n8, pidx, pb) == 0
    meta = n8.view(np.uint8).reshape(-1, 80)[:, 24:32]
    is_inner = ((meta & 0x1f) >= 24) & (meta != 0)
    tri_cnt = np.where((meta != 0) & ~is_inner, np.log2((meta >> 5).astype(np.float64) + 1), 0).astype(int)
    assert tri_cnt.max() <= 3 and tri_cnt.sum(1).max() <= 24 and tri_cnt.sum() == n
    assert len(n8) <= (4 * n - 1) // 7 + 1 and (meta != 0).sum(1).mean() > 6.0    # the optimal collapse fills its nodes
    # the same BVH2 through the reference GPU builder's greedy collapse (WideConverter restatement) needs more nodes
    g8, gp = O.build_bvh8(_root_last(n2, n), n)
    assert O.check_bvh8(g8, gp, pb) == 0 and len(n8) < len(g8)
    assert root_cost > 0 and np.isfinite(root_cost)
    # usable for traversal: closest hits equal brute force
    S = O.OracleScene()
    S.add_mesh(tris, n8, pidx)
    tn, tp, _ = O.cpu_build_bvh8(np.concatenate([sb[:3], sb[3:]])[None].astype(np.float32), 0, 1)
    S.set_instances(np.zeros(1, np.uint32), np.eye(4, dtype=np.float32)[:3].reshape(1, 12), tn, tp)
    rng = np.random.default_rng(0)
    o = (rng.normal(size=(4000, 3)) * 3).astype(np.float32)
    d = (-o / np.linalg.norm(o, axis=1)[:, None] + rng.normal(size=o.shape) * 0.2).astype(np.float32)
    rays = np.zeros(len(o), O.RAY_DTYPE); rays["origin"], rays["direction"], rays["tmax"] = o, d, 1e30
    cmp = O.compare_hits(S, rays, S.trace_closest(rays), S.trace_brute(rays))
    assert cmp["hard"] == 0 and cmp["t_bad"] == 0, cmp


def _root_last(n2, n):
    """Renumbers a root-at-0 BVH2 into NexusBVH's convention (leaves [0, n) in primitive order, root at 2n-2)."""
    leaf = n2[:, 6] == 0xffffffff
    new = np.empty(2 * n - 1, np.uint32)
    new[np.nonzero(leaf)[0]] = n2[leaf, 7]
    inner = np.nonzero(~leaf)[0]
    new[inner] = (2 * n - 2 - np.arange(len(inner))).astype(np.uint32)      # pre-order index 0 (root) -> 2n-2: parents after children
    out = np.zeros_like(n2)
    out[new] = n2
    out[new[inner], 6] = new[n2[inner, 6]]
    out[new[inner], 7] = new[n2[inner, 7]]
    return out


# ------------------------------------------------------------------ CPU collapse pinned against the reference's own code ----
CPU_COLLAPSE_GOLD = os.path.join(os.path.dirname(__file__), "golden", "cpu_collapse_ref.npz")


def _sha(*arrays):
    import hashlib
    return np.frombuffer(hashlib.sha256(b"".join(np.ascontiguousarray(a).tobytes() for a in arrays)).digest(), np.uint8)


@pytest.mark.parametrize("case", [c[0] for c in __import__("golden_cases").cpu_collapse_cases()])
def test_cpu_collapse_restatement_equals_the_reference_bvh8builder(case):
    """The oracle's restatement of the reference's CPU collapse (oracle_sah.cpp, orc_sah_collapse) against the golden output of the
    UNMODIFIED Nexus/src/Geometry/BVH/BVH8Builder.cpp (scripts/make_golden_cpu_collapse.py): same CWBVH8 nodes byte for byte, same
    primitive order, same C(root, 1) bit for bit - decisions, child order (greedy slot assignment), quantisation, numbering.  When
    the compiled reference travelled with the repository it is also run live on the same input."""
    from golden_cases import cpu_collapse_cases
    g = np.load(CPU_COLLAPSE_GOLD)
    tris = next(c[1] for c in cpu_collapse_cases() if c[0] == case)
    n = len(tris)
    if case + "/bvh2" in g.files:
        bvh2 = g[case + "/bvh2"]
    else:                                        # big cases: the input is rebuilt and checked against its digest
        pb, _ = O.prim_bounds(tris, 1)
        bvh2 = O.sah_build_bvh2(pb, threads=4)
        assert (_sha(bvh2) == g[case + "/bvh2_sha256"]).all(), "the oracle's BVH2 builder changed: regenerate the golden file"
    nodes, prim, cost = O.sah_collapse(bvh2, n)
    assert len(nodes) == int(g[case + "/node_count"]) and np.float32(cost) == g[case + "/cost"]
    assert (_sha(nodes, prim) == g[case + "/sha256"]).all()
    if case + "/nodes" in g.files:
        assert (nodes == g[case + "/nodes"]).all() and (prim == g[case + "/prim_idx"]).all()
    if O.have_refcpu():
        rn, rp, rc = O.ref_cpu_collapse(bvh2, n)
        assert (rn == nodes).all() and (rp == prim).all() and rc == cost


def _subtree_prim_sets(nodes8, prim_idx):
    """Topology of a CWBVH8 independent of node numbering, child order and quantisation: the primitive set under every node and
    the primitive group of every leaf child, as sorted tuples."""
    nodes8 = np.ascontiguousarray(nodes8).view(np.uint32).reshape(-1, 20)
    meta = nodes8.view(np.uint8).reshape(-1, 80)[:, 24:32]
    inner_sets, leaf_groups = [], []

    def walk(i):
        prims = []
        imask, child_base, prim_base = int(nodes8[i, 3] >> 24), int(nodes8[i, 4]), int(nodes8[i, 5])
        for s in range(8):
            m = int(meta[i, s])
            if not m:
                continue
            if (imask >> s) & 1:
                prims += walk(child_base + bin(imask & ((1 << s) - 1)).count("1"))
            else:
                cnt, first = bin(m >> 5).count("1"), m & 0x1f
                group = sorted(int(p) for p in prim_idx[prim_base + first: prim_base + first + cnt])
                leaf_groups.append(tuple(group)); prims += group
        inner_sets.append(tuple(sorted(prims)))
        return prims

    import sys
    sys.setrecursionlimit(10000)
    walk(0)
    return sorted(inner_sets), sorted(leaf_groups)


@pytest.mark.skipif(not O.have_refcpu(), reason="oracle/_ref/libnexus_refcpu.so (the compiled reference BVH8Builder) is not present")
@pytest.mark.parametrize("case", ["sphere32", "rock912", "soup500", "soup1900"])
def test_gpu_mode_optimal_collapse_takes_the_reference_builders_decisions(case):
    """NX_COLLAPSE_SAH_OPTIMAL (the GPU kernels dp_eval_kernel + collapse_kernel<true>) is checked bit for bit against its CPU
    restatement orc_build_bvh8_optimal (tests/test_gpu_builder.py).  Here that restatement - GPU-form areas, the GPU converter's slot
    assignment and quantisation, root-last BVH2 numbering - is checked against the UNMODIFIED reference CPU BVH8Builder on the same
    tree with P_MAX = 3: the same BVH8 topology (primitive set under every node, primitive group of every leaf), i.e. the same
    LEAF / INTERNAL / DISTRIBUTE decisions, although node numbering, child order and quantisation are the GPU converter's."""
    from golden_cases import cpu_collapse_cases
    tris = next(c[1] for c in cpu_collapse_cases() if c[0] == case)
    n = len(tris)
    pb, _ = O.prim_bounds(tris, 1)
    bvh2 = O.sah_build_bvh2(pb, threads=4)
    ref_nodes, ref_prim, _ = O.ref_cpu_collapse(bvh2, n)
    got_nodes, got_prim = O.build_bvh8_optimal(_root_last(bvh2, n), n, max_leaf_prims=3)[:2]
    assert len(got_nodes) == len(ref_nodes)
    assert _subtree_prim_sets(got_nodes, got_prim) == _subtree_prim_sets(ref_nodes, ref_prim)
