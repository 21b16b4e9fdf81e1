"""Parity at BENCH scale (BASELINE.json configs[2] and configs[3]), against the live reference kernels when oracle/_ref travelled
to the box, and against the CPU oracle otherwise.

 * configs[2]'s actual scene (1,024 BLASes, 10 M triangles, TLAS over 1,024 instances): one million primary rays plus one million
   incoherent secondary rays through the product's trace kernels and through the UNMODIFIED reference TraceKernel
   (PathTracer.cu:98-113 -> BVH8Trace, BVH8Traversal.cuh:149-324): primitive and instance ids exact except counted exact-distance
   ties, hit distance within 1e-5 relative; a 200k subset against the CPU oracle bit for bit; both traversal loops (ray pool and
   lane-bound) byte-identical.
 * configs[3]: the 10 M and the 50 M triangle builds of the NexusBVH benchmark mesh canonical-tree-equal to the live reference's
   BuildBVH8 (BVHBuilder.cpp:173-267), compared through a hash of the canonical node array and leaf order when the arrays are large.
 * the traversal-stack overflow report (nx_ctx_set_stack_limit).
"""
import hashlib

import numpy as np
import pytest

import bench
import nexus_b200 as nx
import oracle_lib as O
from nexus_b200 import scenes

pytestmark = pytest.mark.gpu
REL = 1e-5


def secondary_rays(rays, hits, n, seed):
    """Incoherent rays: from the hit points of the first n primary rays that hit, into random directions of the sphere."""
    rng = np.random.default_rng(seed)
    ok = np.nonzero(hits["prim"] != 0xffffffff)[0][:n]
    o = rays["origin"][ok] + rays["direction"][ok] * hits["t"][ok, None]
    d = rng.normal(size=(len(ok), 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = (o + 1e-3 * d).astype(np.float32)
    return nx.make_rays(o, d.astype(np.float32))


@pytest.fixture(scope="module")
def config2(ctx):
    desc = bench.make_desc("instanced10m_4k")
    res = bench.WORKLOADS["instanced10m_4k"]["res"]
    scene = scenes.build(ctx, desc, res)
    yield desc, res, scene
    scene.close()


def test_config2_scene_hits_equal_reference_and_oracle(ctx, have_ref, config2):
    desc, res, scene = config2
    assert len(desc["meshes"]) == 1024 and sum(len(m["triangles"]) for m in desc["meshes"]) >= 10_000_000
    o, d = scenes.camera_rays(desc["camera"], res)
    sel = np.random.default_rng(2).choice(len(o), 1_000_000, replace=False)
    primary = nx.make_rays(o[sel], d[sel])
    ph = scene.TraceClosest(primary)
    secondary = secondary_rays(primary, ph, 1_000_000, seed=3)
    assert len(secondary) >= 800_000
    ora = O.oracle_scene_from_product(desc, scene)
    for name, rays in (("primary", primary), ("secondary", secondary)):
        ctx.SetTraceMode("pool")
        got = scene.TraceClosest(rays)
        ctx.SetTraceMode("lane")
        lane = scene.TraceClosest(rays)
        ctx.SetTraceMode("pool")
        assert (got.view(np.uint8) == lane.view(np.uint8)).all(), name          # the two loops agree byte for byte
        sub = np.random.default_rng(4).choice(len(rays), 200_000, replace=False)
        want = ora.trace_closest(rays[sub])
        for f in ("prim", "instance"):
            assert (got[f][sub] == want[f]).all(), (name, f)
        for f in ("t", "u", "v"):
            assert (got[f][sub].view(np.uint32) == want[f].view(np.uint32)).all(), (name, f)
        occ = scene.TraceAny(rays)
        assert (occ.astype(bool) == (got["t"] < nx.MISS_T)).all(), name
    if have_ref:
        O.ref_load_scene(desc, scene, res)
        for name, rays in (("primary", primary), ("secondary", secondary)):
            got = scene.TraceClosest(rays)
            live, _ = O.ref_trace(rays)
            cmp = O.compare_hits(ora, rays, got, live, rel=REL)
            assert cmp["hard"] == 0, (name, cmp)
            assert cmp["tie"] <= 0.001 * cmp["n"], (name, cmp)
            bad, worse = O.t_outliers(rays, got, live, rel=REL)
            assert len(bad) <= 5e-3 * len(rays), (name, len(bad))
            assert len(worse) == 0, name


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        flat = np.ascontiguousarray(a).view(np.uint8).ravel()
        for k in range(0, flat.size, 1 << 28):
            h.update(flat[k:k + (1 << 28)].data)
    return h.hexdigest()


@pytest.mark.parametrize("n", [10_000_000, 50_000_000])
def test_bench_size_trees_equal_live_reference(ctx, have_ref, n):
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    prims = scenes.test_triangles(n)
    b8 = nx.BuildBVH8(ctx, prims, prioritizeSpeed=True)
    n8, p8 = b8.ToHost(); b8.Free()
    ours = _digest(*O.canon_bvh8(n8, p8))
    count = len(n8)
    del n8, p8
    r8, rp8, _ = O.ref_build_bvh8(prims, True)
    assert len(r8) == count
    assert _digest(*O.canon_bvh8(r8, rp8)) == ours


def test_stack_overflow_is_reported(ctx):
    """A push beyond the traversal-stack limit is refused, counted and reported (the reference's 32-entry stack overflows silently).
    No built tree gets near the real limit of 40 entries, so the test lowers it to the minimum."""
    desc = scenes.with_triangle_data(scenes.instanced_scene(n_blas=8, n_instances=64, nu=24, nv=24))
    res = (256, 256)
    scene = scenes.build(ctx, desc, res)
    o, d = scenes.camera_rays(desc["camera"], res)
    rays = nx.make_rays(o, d)
    want = scene.TraceClosest(rays)
    try:
        for mode in ("pool", "lane"):
            ctx.SetTraceMode(mode)
            ctx.SetStackLimit(40)
            assert (scene.TraceClosest(rays).view(np.uint8) == want.view(np.uint8)).all()
            ctx.SetStackLimit(8)
            # 8 entries are enough for most rays of this scene but not for all of them: either the call succeeds with identical hits or
            # it fails loudly; it never returns different hits silently
            try:
                got = scene.TraceClosest(rays)
                assert (got.view(np.uint8) == want.view(np.uint8)).all()
            except nx.NexusError as e:
                assert "stack overflow" in str(e)
    finally:
        ctx.SetStackLimit(40); ctx.SetTraceMode("pool")
    # a limit the scene certainly exceeds: forced through a deep chain of nested hits
    deep = scenes.with_triangle_data(scenes.instanced_scene(n_blas=4, n_instances=512, nu=32, nv=32))
    s2 = scenes.build(ctx, deep, res)
    o, d = scenes.camera_rays(deep["camera"], res)
    rays = nx.make_rays(o, d)
    ok = s2.TraceClosest(rays)
    depth_needed = None
    try:
        for limit in range(39, 1, -1):        # the ray-pool loop accepts limits down to 2 entries (the lane-bound loop clamps at its 8 shared entries)
            ctx.SetStackLimit(limit)
            try:
                got = s2.TraceClosest(rays)
                assert (got.view(np.uint8) == ok.view(np.uint8)).all()
            except nx.NexusError as e:
                assert "stack overflow" in str(e)
                depth_needed = limit + 1
                break
    finally:
        ctx.SetStackLimit(40)
    assert depth_needed is not None, "no stack limit down to 2 entries made this scene overflow"
    assert (s2.TraceClosest(rays).view(np.uint8) == ok.view(np.uint8)).all()    # the error state does not stick
    scene.close(); s2.close()
