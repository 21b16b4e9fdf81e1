"""Host logic of the Python mirror without a GPU: the reference's edit pattern (host objects edited in place, Invalidate*, Scene::Update,
Scene.h:19-49 / Scene.cpp:34-63) is bookkeeping above the C ABI, so it is tested here against a recording stand-in for libnexus_b200.so:
which ABI calls Update() makes, for which objects, in which order, and that nothing is pushed twice or forgotten."""
import ctypes as C

import numpy as np
import pytest

import nexus_b200 as nx


class FakeLib:
    """Records every ABI call as (name, plain-python arguments); add_* calls return consecutive indices like the library."""

    def __init__(self):
        self.calls, self.counts = [], {}

    def __getattr__(self, name):
        def fn(*args):
            plain = []
            for a in args:
                v = getattr(a, "value", a)
                plain.append(v if isinstance(v, (int, float, str, bytes, type(None))) else type(a).__name__)
            self.calls.append((name, plain))
            if name.startswith("nx_scene_add_"):
                key = "nx_scene_add_instance" if name == "nx_scene_add_instance_matrix" else name     # instances share one index space
                k = self.counts.get(key, 0); self.counts[key] = k + 1
                return k
            if name == "nx_scene_instance_matrix":
                return 0
            return 0
        return fn

    def names(self):
        return [c[0] for c in self.calls]

    def clear(self):
        self.calls = []


@pytest.fixture
def fake(monkeypatch):
    f = FakeLib()
    monkeypatch.setattr(nx, "lib", lambda: f)
    return f


class _Ctx:
    _h = C.c_void_p(1)
    device = 0


def _scene(fake):
    s = nx.Scene(_Ctx(), (64, 48))
    am = s.GetAssetManager()
    for k in range(3):
        assert am.AddMaterial(nx.Material(roughness=0.1 * (k + 1))) == k
    tri = np.zeros((2, 9), np.float32)
    assert am.AddMesh("a", 0, tri) == 0 and am.AddMesh("b", 1, tri) == 1
    for k in range(4):
        s.CreateMeshInstance(k % 2, position=(k, 0, 0))
    for k in range(3):
        assert s.AddLight(nx.Light(nx.Light.POINT, position=(0, k, 0))) == k
    s.Update()
    fake.clear()
    return s


def test_update_pushes_exactly_what_was_invalidated(fake):
    s = _scene(fake)
    assert not s.IsInvalid()
    s.Update()
    assert fake.names() == ["nx_scene_update"]                       # nothing dirty: just the library's own update
    fake.clear()
    # instance 2: moved and re-assigned; instance 1: only invalidated by the caller (nothing changed: nothing to push)
    inst = s.GetMeshInstances()[2]
    inst.SetPosition((1, 2, 3)); inst.SetRotationY(45.0); inst.AssignMaterial(2)
    s.InvalidateMeshInstance(1)
    # material 1 edited in place, light 0 edited in place, camera moved, path length changed
    s.GetMaterials()[1].roughness = 0.9; s.GetAssetManager().InvalidateMaterial(1)
    s.GetLights()[0].intensity = 5.0; s.InvalidateLight(0)
    s.GetCamera().SetPosition((0, 1, 2)); s.GetCamera().Invalidate()
    s.GetRenderSettings().pathLength = 4
    assert s.IsInvalid() and fake.names() == []                      # edits alone make no ABI call
    s.Update()
    assert fake.names() == ["nx_scene_set_camera", "nx_scene_set_render_settings", "nx_scene_set_material", "nx_scene_set_instance_transform",
                            "nx_scene_set_instance_material", "nx_scene_set_light", "nx_scene_update"]
    by = dict((n, a) for n, a in fake.calls)
    assert by["nx_scene_set_material"][1] == 1 and by["nx_scene_set_instance_transform"][1] == 2
    assert by["nx_scene_set_instance_material"][1:] == [2, 2] and by["nx_scene_set_light"][1] == 0
    assert not s.IsInvalid()
    fake.clear()
    s.Update()
    assert fake.names() == ["nx_scene_update"]                       # and nothing is pushed twice


def test_remove_light_renumbers_pending_invalidations(fake):
    s = _scene(fake)
    s.GetLights()[2].intensity = 9.0; s.InvalidateLight(2)
    s.GetLights()[0].intensity = 3.0; s.InvalidateLight(0)
    s.RemoveLight(1)                                                 # light 2 becomes light 1
    assert len(s.GetLights()) == 2 and s.GetLights()[1].intensity == 9.0
    s.Update()
    pushed = [a[1] for n, a in fake.calls if n == "nx_scene_set_light"]
    assert fake.names()[0] == "nx_scene_remove_light" and pushed == [0, 1]
    with pytest.raises(nx.NexusError):
        s.RemoveLight(5)
    with pytest.raises(nx.NexusError):
        s.InvalidateLight(2)


def test_set_transform_applies_at_once_and_matrix_instances_keep_their_matrix(fake):
    s = _scene(fake)
    inst = s.GetMeshInstances()[0]
    inst.SetTransform((1, 1, 1), (0, 90, 0), (2, 2, 2))
    assert fake.names() == ["nx_scene_set_instance_transform"] and not s.IsInvalid()
    fake.clear()
    m = s.CreateMeshInstanceMatrix(1, np.eye(4, dtype=np.float32))
    m.AssignMaterial(0)
    s.Update()
    assert "nx_scene_set_instance_transform" not in fake.names() and "nx_scene_set_instance_material" in fake.names()
    fake.clear()
    m.SetScale(3.0)                                                  # from now on position / rotation / scale define it
    s.Update()
    assert "nx_scene_set_instance_transform" in fake.names()


def test_materials_replaced_through_the_legacy_call_upload_at_once(fake):
    s = _scene(fake)
    s.GetAssetManager().InvalidateMaterial(2, nx.Material(roughness=0.77))
    assert fake.names() == ["nx_scene_set_material"] and s.GetMaterials()[2].roughness == 0.77 and not s.IsInvalid()
    with pytest.raises(nx.NexusError):
        s.GetAssetManager().InvalidateMaterial(7)


def test_importing_into_a_populated_scene_shifts_material_and_texture_indices(fake, tmp_path):
    """Scene.CreateMeshInstanceFromFile appends: an asset's material indices are offset by the materials the scene already has and its
    texture ids by the textures already registered (checked on the MaterialPod records handed to the library)."""
    PIL = pytest.importorskip("PIL.Image")
    import io
    import test_gltf
    px = np.random.RandomState(1).randint(0, 256, (2, 2, 4)).astype(np.uint8)
    buf = io.BytesIO(); PIL.fromarray(px, "RGBA").save(buf, format="PNG")
    png = buf.getvalue()
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    binary = pos.tobytes() + png
    js = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": [{"mesh": 0}],
          "meshes": [{"primitives": [{"attributes": {"POSITION": 0}, "material": 1}]}],
          "materials": [{}, {"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}, "normalTexture": {"index": 0}}],
          "textures": [{"source": 0}], "images": [{"bufferView": 1, "mimeType": "image/png"}],
          "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3"}],
          "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": pos.nbytes}, {"buffer": 0, "byteOffset": pos.nbytes, "byteLength": len(png)}],
          "buffers": [{"byteLength": len(binary)}]}
    test_gltf._write_glb(tmp_path / "t.glb", js, binary)

    pods = []
    real_getattr = FakeLib.__getattr__

    def spy(self, name):
        fn = real_getattr(self, name)
        if name != "nx_scene_add_material":
            return fn

        def wrapped(*args):
            pods.append(args[1]._obj)            # byref(MaterialPod)
            return fn(*args)
        return wrapped
    FakeLib.__getattr__ = spy
    try:
        s = _scene(fake)                          # 3 materials, 2 meshes, 4 instances
        s.GetAssetManager().AddTexture(np.zeros((2, 2, 4), np.uint8))
        s.GetAssetManager().AddTexture(np.zeros((2, 2, 4), np.float32))
        pods.clear(); fake.clear()
        created = s.CreateMeshInstanceFromFile(str(tmp_path) + "/", "t.glb")
    finally:
        FakeLib.__getattr__ = real_getattr
    assert len(created) == 1 and created[0].index == 4 and created[0].meshIdx == 2
    assert fake.names() == ["nx_scene_add_texture", "nx_scene_add_texture", "nx_scene_add_material", "nx_scene_add_material", "nx_scene_add_mesh", "nx_scene_add_instance_matrix"]
    assert len(pods) == 2 and pods[0].base_color_map == -1
    assert pods[1].base_color_map == 2 and pods[1].normal_map == 3            # the asset's textures 0 (sRGB) and 1 (linear) behind the scene's two
    add_mesh = [a for n, a in fake.calls if n == "nx_scene_add_mesh"][0]
    assert add_mesh[-1] == 3 + 1                                              # material 1 of the asset = material 4 of the scene
    assert len(s.GetMaterials()) == 5
