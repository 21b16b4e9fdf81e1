"""Minimal binary glTF 2.0 (.glb) import into the scene description the host API consumes.

The reference imports assets through Assimp (src/Assets/OBJLoader.cpp:8-446): one Mesh per primitive, one MeshInstance per
node that references it (node transform = instance transform), materials from the glTF PBR block and the KHR extensions its
demo scenes use.  Assimp is not available here, so this reader covers exactly what those assets need:

  geometry   triangle primitives (mode 4), indexed or not; POSITION, NORMAL, TANGENT, TEXCOORD_0 -> NXB::Triangle rows and
             TriangleData rows (normals, tangents, texture coordinates; geometric normals when the asset has none)
  nodes      matrix or translation / rotation (quaternion) / scale, accumulated down the hierarchy -> instance matrices
  materials  pbrMetallicRoughness (baseColorFactor -> baseColor + opacity, metallicFactor, roughnessFactor), emissiveFactor,
             KHR_materials_emissive_strength -> intensity (1.0 when only emissiveFactor is set, OBJLoader.cpp:113-118),
             KHR_materials_specular -> specularWeight / specularColor, KHR_materials_ior -> ior,
             KHR_materials_transmission -> transmission                                        (OBJLoader.cpp:96-119)
  camera     the first perspective camera (yfov + aspect -> horizontal FOV); otherwise the reference's default camera
             (position (0, 4, 14), forward (0, 0, -1), 45 degrees, focus 5: src/Scene/Scene.cpp:9-10)

Texture images (embedded PNG / JPEG) are decoded with Pillow when it is installed (the reference uses stb_image,
src/Assets/IMGLoader.cpp:13-43: 8-bit RGBA, rows top to bottom), or with a caller-supplied
`decode_image(bytes) -> (h, w, 4) uint8`; without either, assets that use textures are rejected with a clear error.
Pure host code: numpy (+ optional Pillow), no GPU.
"""
import json
import struct

import numpy as np

from . import Camera, Material, RenderSettings

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}


class GltfError(ValueError):
    pass


def _pillow_decoder():
    try:
        from PIL import Image
    except ImportError:
        return None
    import io

    def decode(data):
        return np.asarray(Image.open(io.BytesIO(bytes(data))).convert("RGBA"), np.uint8)
    return decode


def _chunks(blob):
    if len(blob) < 20 or blob[:4] != b"glTF":
        raise GltfError("not a binary glTF file")
    version, length = struct.unpack_from("<II", blob, 4)
    if version != 2:
        raise GltfError(f"glTF version {version} is not supported")
    off, js, binary = 12, None, b""
    while off + 8 <= min(length, len(blob)):
        clen, ctype = struct.unpack_from("<I4s", blob, off)
        body = blob[off + 8: off + 8 + clen]
        if ctype == b"JSON":
            js = json.loads(body.decode("utf-8"))
        elif ctype == b"BIN\x00":
            binary = body
        off += 8 + clen + (-clen % 4)
    if js is None:
        raise GltfError("the file has no JSON chunk")
    return js, binary


def _accessor(js, binary, index):
    acc = js["accessors"][index]
    if "sparse" in acc:
        raise GltfError("sparse accessors are not supported")
    dtype, width = np.dtype(_COMPONENT[acc["componentType"]]), _WIDTH[acc["type"]]
    count = acc["count"]
    if "bufferView" not in acc:
        return np.zeros((count, width), dtype)
    view = js["bufferViews"][acc["bufferView"]]
    if view.get("buffer", 0) != 0:
        raise GltfError("only the embedded binary buffer is supported")
    start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = view.get("byteStride", 0) or dtype.itemsize * width
    raw = np.frombuffer(binary, np.uint8, count=stride * (count - 1) + dtype.itemsize * width, offset=start)
    rows = np.lib.stride_tricks.as_strided(raw, shape=(count, dtype.itemsize * width), strides=(stride, 1))
    out = np.ascontiguousarray(rows).view(dtype).reshape(count, width)
    if acc.get("normalized") and dtype.kind in "ui":
        out = out.astype(np.float32) / np.float32(np.iinfo(dtype).max)
    return out


def _node_matrix(node):
    if "matrix" in node:
        return np.asarray(node["matrix"], np.float64).reshape(4, 4).T          # glTF stores column-major
    t, r, s = node.get("translation", (0, 0, 0)), node.get("rotation", (0, 0, 0, 1)), node.get("scale", (1, 1, 1))
    x, y, z, w = (float(v) for v in r)
    rot = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                    [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                    [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    m = np.eye(4)
    m[:3, :3] = rot * np.asarray(s, np.float64)[None, :]
    m[:3, 3] = t
    return m


def _material(js, m, texture_id):
    pbr = m.get("pbrMetallicRoughness", {})
    ext = m.get("extensions", {})
    base = pbr.get("baseColorFactor", (1.0, 1.0, 1.0, 1.0))
    emission = tuple(float(c) for c in m.get("emissiveFactor", (0.0, 0.0, 0.0)))
    out = Material(baseColor=tuple(float(c) for c in base[:3]), opacity=float(base[3]),
                   metalness=float(pbr.get("metallicFactor", 1.0)), roughness=float(pbr.get("roughnessFactor", 1.0)),
                   emissionColor=emission)
    out.intensity = 1.0 if max(emission) > 0.0 else 0.0
    if "KHR_materials_emissive_strength" in ext:
        out.intensity = float(ext["KHR_materials_emissive_strength"].get("emissiveStrength", 1.0))
    spec = ext.get("KHR_materials_specular", {})
    out.specularWeight = float(spec.get("specularFactor", 1.0))
    out.specularColor = tuple(float(c) for c in spec.get("specularColorFactor", (1.0, 1.0, 1.0)))
    out.ior = float(ext.get("KHR_materials_ior", {}).get("ior", 1.5))
    out.transmission = float(ext.get("KHR_materials_transmission", {}).get("transmissionFactor", 0.0))
    for key, attr, srgb in (("baseColorTexture", "baseColorMap", True), ("metallicRoughnessTexture", "metallicRoughnessMap", False)):
        if key in pbr:
            setattr(out, attr, texture_id(pbr[key]["index"], srgb))
    for key, attr, srgb in (("normalTexture", "normalMap", False), ("emissiveTexture", "emissiveMap", True)):
        if key in m:
            setattr(out, attr, texture_id(m[key]["index"], srgb))
    return out


def _triangle_rows(js, binary, prim):
    if prim.get("mode", 4) != 4:
        raise GltfError("only triangle primitives (mode 4) are supported")
    att = prim["attributes"]
    pos = _accessor(js, binary, att["POSITION"]).astype(np.float32)
    idx = _accessor(js, binary, prim["indices"]).astype(np.int64).ravel() if "indices" in prim else np.arange(len(pos))
    if len(idx) % 3:
        raise GltfError("index count is not a multiple of three")
    idx = idx.reshape(-1, 3)
    tris = pos[idx].reshape(-1, 9)
    n = len(tris)
    data = np.zeros((n, 24), np.float32)
    if "NORMAL" in att:
        data[:, 0:9] = _accessor(js, binary, att["NORMAL"]).astype(np.float32)[idx].reshape(n, 9)
    else:
        g = np.cross(tris[:, 3:6] - tris[:, 0:3], tris[:, 6:9] - tris[:, 0:3])
        g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-30)
        data[:, 0:3] = data[:, 3:6] = data[:, 6:9] = g
    if "TANGENT" in att:
        data[:, 9:18] = _accessor(js, binary, att["TANGENT"]).astype(np.float32)[:, :3][idx].reshape(n, 9)
    if "TEXCOORD_0" in att:
        data[:, 18:24] = _accessor(js, binary, att["TEXCOORD_0"]).astype(np.float32)[idx].reshape(n, 6)
    return np.ascontiguousarray(tris, np.float32), data


def load_glb(path, path_length=10, decode_image=None):
    """Reads a .glb file into the scene description `scenes.build` instantiates: meshes (one per primitive), instances (one per
    node and primitive, `matrix` = accumulated node transform), materials, textures, camera, settings."""
    with open(path, "rb") as f:
        js, binary = _chunks(f.read())
    textures, texture_of = [], {}
    if decode_image is None:
        decode_image = _pillow_decoder()

    def texture_id(tex_index, srgb):
        key = (tex_index, srgb)
        if key not in texture_of:
            if decode_image is None:
                raise GltfError("the asset uses texture images; install Pillow or pass decode_image(bytes) -> (h, w, 4) uint8 to load them")
            image = js["images"][js["textures"][tex_index]["source"]]
            view = js["bufferViews"][image["bufferView"]]
            data = binary[view.get("byteOffset", 0): view.get("byteOffset", 0) + view["byteLength"]]
            texture_of[key] = len(textures)
            textures.append((np.ascontiguousarray(decode_image(data), np.uint8), srgb))
        return texture_of[key]

    materials = [_material(js, m, texture_id) for m in js.get("materials", [])] or [Material()]
    meshes, mesh_of_prim, instances = [], {}, []
    camera = None

    def visit(node_index, parent):
        nonlocal camera
        node = js["nodes"][node_index]
        world = parent @ _node_matrix(node)
        if "mesh" in node:
            mesh = js["meshes"][node["mesh"]]
            for k, prim in enumerate(mesh["primitives"]):
                key = (node["mesh"], k)
                if key not in mesh_of_prim:
                    tris, data = _triangle_rows(js, binary, prim)
                    mesh_of_prim[key] = len(meshes)
                    meshes.append({"name": f"{mesh.get('name', 'mesh')}.{k}", "triangles": tris, "triangle_data": data, "material": int(prim.get("material", 0))})
                instances.append({"mesh": mesh_of_prim[key], "material": -1, "matrix": world.astype(np.float32)})
        if "camera" in node and camera is None:
            cam = js["cameras"][node["camera"]]
            if cam.get("type") == "perspective":
                p = cam["perspective"]
                aspect = float(p.get("aspectRatio", 16.0 / 9.0))
                hfov = np.degrees(2.0 * np.arctan(np.tan(0.5 * float(p["yfov"])) * aspect))
                fwd = world[:3, :3] @ np.array([0.0, 0.0, -1.0])
                camera = Camera(position=tuple(world[:3, 3]), forward=tuple(fwd / np.linalg.norm(fwd)), horizontalFOV=float(hfov), focusDistance=5.0, defocusAngle=0.0)
        for child in node.get("children", []):
            visit(child, world)

    scene = js["scenes"][js.get("scene", 0)] if js.get("scenes") else {"nodes": list(range(len(js.get("nodes", []))))}
    for root in scene.get("nodes", []):
        visit(root, np.eye(4))
    if not meshes:
        raise GltfError("the asset contains no triangle geometry")
    return {"name": str(path), "meshes": meshes, "instances": instances, "materials": materials, "textures": textures, "lights": [],
            "camera": camera or Camera(), "settings": RenderSettings(useMIS=True, pathLength=path_length)}
