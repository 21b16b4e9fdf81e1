"""SURVEY.md §8(e), "BVH build, many meshes": BLAS builds are independent units.  A scene assembled from BLASes that were built
standalone (nx_scene_build_blas), packed, exchanged (nexus_b200.multigpu.exchange_blas; world size 1 here, world size 2 over gloo in
tests/test_multigpu_host.py and over NCCL in scripts/check_sharded_build.py) and imported (nx_scene_add_mesh_prebuilt) must be the scene a
plain build produces: byte-identical BLASes, identical closest hits, identical frame."""
import numpy as np
import pytest
import torch

import nexus_b200 as nx
import oracle_lib as O
from nexus_b200 import multigpu, scenes

pytestmark = pytest.mark.gpu


def test_sharded_scene_build_equals_the_plain_build(ctx):
    res = (160, 120)
    desc = scenes.with_triangle_data(scenes.instanced_scene(n_blas=5, n_instances=14, nu=20, nv=18, path_length=4))
    plain = scenes.build(ctx, desc, res)
    shard = multigpu.build_scene_sharded(ctx, desc, res)
    for k in range(len(desc["meshes"])):
        a, b = plain.MeshBVH(k), shard.MeshBVH(k)
        # two builds of one mesh number the nodes of a level in schedule order: compare the canonical renumbering
        (na, pa), (nb, pb) = O.canon_bvh8(*a.ToHost()), O.canon_bvh8(*b.ToHost())
        assert a.nodeCount == b.nodeCount and (na == nb).all() and (pa == pb).all() and (a.bounds == b.bounds).all(), k
    o, d = scenes.camera_rays(desc["camera"], res)
    rays = nx.make_rays(o, d)
    ha, hb = plain.TraceClosest(rays), shard.TraceClosest(rays)
    assert ha.tobytes() == hb.tobytes()
    pa, pb = nx.PathTracer(ctx, res), nx.PathTracer(ctx, res)
    pa.Render(plain, frames=2, firstFrame=1); pb.Render(shard, frames=2, firstFrame=1)
    assert np.allclose(pa.ReadAccumulation(), pb.ReadAccumulation(), rtol=1e-4, atol=1e-5)
    pa.close(); pb.close(); plain.close(); shard.close()


def test_standalone_blas_is_the_mesh_blas(ctx):
    tris = scenes.rock(5, nu=24, nv=20)
    desc = {"meshes": [{"name": "rock", "material": 0, "triangles": tris}]}
    bvh = nx.BuildBLAS(ctx, tris)
    scene = nx.Scene(ctx, (32, 32))
    am = scene.GetAssetManager()
    am.AddMaterial(nx.Material())
    am.AddMesh("rock", 0, tris)
    nodes, prim = O.canon_bvh8(*bvh.ToHost())
    n2, p2 = O.canon_bvh8(*scene.MeshBVH(0).ToHost())
    assert (nodes == n2).all() and (prim == p2).all()
    # imported as is: the scene holds the very bytes it was given
    raw_nodes, raw_prim = bvh.ToHost()
    k = am.AddMeshPrebuilt("copy", 0, tris, None, bvh.h.nodes, bvh.nodeCount, bvh.h.prim_idx, bvh.bounds)
    n3, p3 = scene.MeshBVH(k).ToHost()
    assert k == 1 and (n3 == raw_nodes).all() and (p3 == raw_prim).all() and (scene.MeshBVH(k).bounds == bvh.bounds).all()
    # a prebuilt import is validated: too many nodes for the triangle count is refused
    dev = torch.zeros(80 * 4, dtype=torch.uint8, device="cuda")
    with pytest.raises(nx.NexusError):
        am.AddMeshPrebuilt("bad", 0, tris[:2], None, dev.data_ptr(), 4, dev.data_ptr(), bvh.bounds)
    bvh.Free(); scene.close()
