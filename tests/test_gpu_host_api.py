"""The kept host API (Scene / asset surface) through the Python mirror: host objects edited in place and pushed by
Scene.Update after Invalidate*, exactly the pattern of the reference's UI code (Scene.h:19-49, MeshInstance.h:22-53, Camera.h:19-35,
AssetManager.h:18-44).  An edited scene must equal a scene created in the final state."""
import numpy as np
import pytest

import nexus_b200 as nx
from nexus_b200 import scenes

pytestmark = pytest.mark.gpu
RES = (96, 72)


def _render(ctx, scene, frames=6):
    pt = nx.PathTracer(ctx, RES)
    pt.Render(scene, frames=frames, firstFrame=1)
    img = pt.ReadAccumulation().copy()
    pt.close()
    return img


def _desc():
    return scenes.with_triangle_data(scenes.instanced_scene(n_blas=3, n_instances=6, nu=14, nv=12, path_length=3))


def _same_hits(a, b):
    """Same scene, possibly different acceleration structure.  An instance that was moved keeps a BLAS of its own, one that was created
    in place may live in the scene's merged world-space BLAS: the hit distance is then evaluated in object space in one scene and in
    world space in the other, so ids must agree (all but edge-grazing rays) and distances agree to fp32 rounding."""
    same = (a["prim"] == b["prim"]) & (a["instance"] == b["instance"])
    assert same.mean() >= 0.999, same.mean()
    hit = same & (a["prim"] != 0xffffffff)
    assert np.allclose(a["t"][hit], b["t"][hit], rtol=2e-5, atol=0)


def test_instance_edits_through_host_objects(ctx):
    want_desc = _desc()
    want_desc["instances"][2].update(position=(0.5, 1.5, -1.0), rotation=(15.0, 40.0, -20.0), scale=(0.7, 0.7, 0.7), material=1)
    want = scenes.build(ctx, want_desc, RES)
    scene = scenes.build(ctx, _desc(), RES)
    inst = scene.GetMeshInstances()[2]
    assert len(scene.GetMeshInstances()) == len(want_desc["instances"]) and not scene.IsInvalid() and not scene.IsEmpty()
    inst.SetPosition((0.5, 1.5, -1.0)); inst.SetRotationX(15.0); inst.SetRotationY(40.0); inst.SetRotationZ(-20.0); inst.SetScale(0.7)
    inst.AssignMaterial(1)
    scene.InvalidateMeshInstance(inst.index)
    assert scene.IsInvalid()
    scene.Update()
    assert not scene.IsInvalid()
    # T * Rz * Ry * Rx * S (MeshInstance.h:36-40) and the world box of the eight transformed corners (MeshInstance.h:42-53)
    ax, ay, az = np.radians([15.0, 40.0, -20.0])
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    m = inst.GetTransfromationMatrix().astype(np.float64)
    assert np.allclose(m[:3, :3], rz @ ry @ rx * 0.7, atol=1e-6) and np.allclose(m[:3, 3], (0.5, 1.5, -1.0)) and np.allclose(m[3], (0, 0, 0, 1))
    mb = scene.MeshBounds(inst.meshIdx).astype(np.float64)
    corners = np.array([[mb[3 * (i & 1)], mb[1 + 3 * ((i >> 1) & 1)], mb[2 + 3 * ((i >> 2) & 1)]] for i in range(8)]) @ m[:3, :3].T + m[:3, 3]
    assert np.allclose(inst.GetBounds(), np.concatenate([corners.min(0), corners.max(0)]), atol=1e-5)
    o, d = scenes.camera_rays(want_desc["camera"], RES)
    rays = nx.make_rays(o, d)
    _same_hits(scene.TraceClosest(rays), want.TraceClosest(rays))
    assert np.allclose(_render(ctx, scene), _render(ctx, want), rtol=2e-3, atol=1e-3)
    scene.close(); want.close()
    # without instance merging both scenes hold the same trees: bit for bit
    ctx.SetInstanceMerging(False)
    try:
        want = scenes.build(ctx, want_desc, RES)
        scene = scenes.build(ctx, _desc(), RES)
        inst = scene.GetMeshInstances()[2]
        inst.SetPosition((0.5, 1.5, -1.0)); inst.SetRotationX(15.0); inst.SetRotationY(40.0); inst.SetRotationZ(-20.0); inst.SetScale(0.7)
        inst.AssignMaterial(1)
        scene.InvalidateMeshInstance(inst.index); scene.Update()
        assert scene.TraceClosest(rays).tobytes() == want.TraceClosest(rays).tobytes()
        assert np.allclose(_render(ctx, scene), _render(ctx, want), rtol=1e-4, atol=1e-5)
        scene.close(); want.close()
    finally:
        ctx.SetInstanceMerging(True)


def test_camera_settings_material_and_light_edits(ctx):
    desc = _desc()
    desc["lights"] = [nx.Light(nx.Light.POINT, position=(0.0, 6.0, 0.0), color=(1.0, 0.9, 0.8), intensity=30.0)]
    scene = scenes.build(ctx, desc, RES)
    base = _render(ctx, scene)
    # camera: edit the host object, Invalidate, Update == a scene given that camera
    cam = scene.GetCamera()
    cam.SetPosition((2.0, 5.0, 12.0)); cam.SetHorizontalFOV(35.0); cam.SetFocusDist(7.0); cam.Invalidate()
    assert scene.IsInvalid()
    scene.Update()
    d2 = _desc(); d2["lights"] = desc["lights"]
    d2["camera"] = nx.Camera(position=(2.0, 5.0, 12.0), forward=desc["camera"].forward, horizontalFOV=35.0, focusDistance=7.0,
                             defocusAngle=desc["camera"].defocusAngle)
    other = scenes.build(ctx, d2, RES)
    moved = _render(ctx, scene)
    assert np.allclose(moved, _render(ctx, other), rtol=1e-4, atol=1e-5) and not np.allclose(moved, base, rtol=1e-2, atol=1e-3)
    # render settings: edit the host object; Update uploads it
    rs = scene.GetRenderSettings()
    rs.backgroundColor, rs.backgroundIntensity = (0.2, 0.4, 0.9), 2.0
    assert scene.IsInvalid()
    scene.Update()
    sky = _render(ctx, scene)
    assert sky[..., 2].mean() > moved[..., 2].mean() * 1.05
    # materials: GetMaterials()[i] edited + InvalidateMaterial(i) + Update
    mats = scene.GetMaterials()
    assert len(mats) == len(desc["materials"]) and scene.GetAssetManager().GetMaterials() is mats
    for i in range(len(mats)):
        mats[i].baseColor, mats[i].metalness, mats[i].transmission = (0.9, 0.05, 0.05), 0.0, 0.0
        scene.GetAssetManager().InvalidateMaterial(i)
    assert scene.IsInvalid()
    scene.Update()
    red = _render(ctx, scene)
    assert red[..., 0].mean() / red[..., 1].mean() > 1.1 * sky[..., 0].mean() / sky[..., 1].mean()
    # lights: brighten in place, then remove
    assert len(scene.GetLights()) == 1
    scene.GetLights()[0].intensity = 3000.0
    scene.InvalidateLight(0); scene.Update()
    bright = _render(ctx, scene)
    scene.RemoveLight(0); scene.Update()
    assert scene.GetLights() == []
    dark = _render(ctx, scene)
    assert bright.mean() > red.mean() * 1.05 and dark.mean() < bright.mean() / 1.05
    with pytest.raises(nx.NexusError):
        scene.InvalidateLight(0)
    with pytest.raises(nx.NexusError):
        scene.GetAssetManager().InvalidateMaterial(99)
    scene.close(); other.close()


def test_assets_and_environment_from_files(ctx, tmp_path):
    """Scene::CreateMeshInstanceFromFile and Scene::AddHDRMap(filePath, fileName) (Scene.cpp:97-107): importing a .glb and an .obj
    into an existing scene appends their materials, meshes and instances (material indices shifted past the scene's own), and an
    environment map read from a Radiance .hdr file equals the same pixels handed over directly."""
    import test_gltf
    import test_hdr
    import test_obj
    from nexus_b200 import gltf, hdr, obj
    test_gltf._two_quads_glb(tmp_path / "q.glb")
    cube = test_obj._write(tmp_path)
    sky = np.random.RandomState(2).uniform(0.2, 1.5, (16, 32, 3)).astype(np.float32)
    test_hdr._write(tmp_path / "sky.hdr", test_hdr._rgbe(sky), rle=True)
    sky_px = hdr.load_hdr(tmp_path / "sky.hdr")

    def base_scene():
        scene = nx.Scene(ctx, RES)
        am = scene.GetAssetManager()
        am.AddMaterial(nx.Material(baseColor=(0.3, 0.7, 0.3), roughness=0.8))
        am.AddMesh("floor", 0, np.array([[-6, -0.5, 6, 6, -0.5, 6, 6, -0.5, -6], [-6, -0.5, 6, 6, -0.5, -6, -6, -0.5, -6]], np.float32))
        scene.CreateMeshInstance(0)
        scene.SetCamera(nx.Camera(position=(2.0, 4.0, 12.0), forward=(0.0, -0.19611614, -0.98058068), horizontalFOV=45.0))
        scene.SetRenderSettings(nx.RenderSettings(pathLength=3))
        return scene

    a = base_scene()
    got_glb = a.CreateMeshInstanceFromFile(str(tmp_path) + "/", "q.glb")
    got_obj = a.CreateMeshInstanceFromFile(cube)
    a.AddHDRMap(str(tmp_path) + "/", "sky.hdr")
    a.Update()
    assert len(got_glb) == 2 and len(got_obj) == 2 and len(a.GetMeshInstances()) == 5 and len(a.GetMaterials()) == 1 + 2 + 2
    assert got_glb[0].materialIdx == -1 and got_obj[1].meshIdx == 4 and got_glb[1].name.startswith("quads")
    with pytest.raises(nx.NexusError):
        a.CreateMeshInstanceFromFile(str(tmp_path / "scene.fbx"))

    # the same content assembled by hand
    b = base_scene()
    am = b.GetAssetManager()
    dg, do = gltf.load_glb(tmp_path / "q.glb"), obj.load_obj(cube)
    for d in (dg, do):
        m0 = len(am.GetMaterials())
        for m in d["materials"]:
            am.AddMaterial(m)
        ids = [am.AddMesh(m["name"], m0 + m["material"], m["triangles"], m["triangle_data"]) for m in d["meshes"]]
        for i in d["instances"]:
            b.CreateMeshInstanceMatrix(ids[i["mesh"]], i["matrix"], -1)
    b.AddHDRMap(sky_px)
    b.Update()
    o, d = scenes.camera_rays(a.GetCamera(), RES)
    rays = nx.make_rays(o, d)
    ha, hb = a.TraceClosest(rays), b.TraceClosest(rays)
    assert ha.tobytes() == hb.tobytes() and len(np.unique(ha["instance"][ha["prim"] != 0xffffffff])) >= 2
    ia, ib = _render(ctx, a), _render(ctx, b)
    assert ia.mean() > 0.05 and np.allclose(ia, ib, rtol=1e-4, atol=1e-5)
    a.close(); b.close()
