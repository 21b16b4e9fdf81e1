"""Procedural scenes for the configurations BASELINE.json names.  Pure numpy, deterministic.

A scene description is a dict:
  meshes:    list of {"name", "triangles" (n,9) float32, "material" int}
  instances: list of {"mesh", "material" (or -1), "position", "rotation" (Euler degrees), "scale"}
  materials: list of nexus_b200.Material
  lights:    list of nexus_b200.Light (punctual only; emissive instances become lights automatically)
  camera:    nexus_b200.Camera
  settings:  nexus_b200.RenderSettings
  hdr:       optional (h, w, 4) float32 equirect map
"""
import numpy as np

from . import Camera, Material, RenderSettings, Scene


def _quad(a, b, c, d):
    """Two triangles (a,b,c), (a,c,d) of a quad given in order."""
    return [np.concatenate([a, b, c]), np.concatenate([a, c, d])]


def _box(corners_bottom, height_vec):
    """Five-sided-plus-bottom box from four bottom corners (ccw seen from above) and an extrusion vector."""
    b = [np.asarray(c, np.float32) for c in corners_bottom]
    t = [c + np.asarray(height_vec, np.float32) for c in b]
    tris = _quad(t[0], t[1], t[2], t[3])
    for i in range(4):
        j = (i + 1) % 4
        tris += _quad(b[i], b[j], t[j], t[i])
    return tris


def cornell_box(path_length=10):
    """Config 2.  Dimensions, colours and the 35x emissive 0.47 x 0.38 ceiling quad follow the reference's demo asset
    (Nexus/assets/demo_scenes/cornell_box/cornell_box.glb: 8 primitives, 32 triangles, node rotated +90 deg about X,
    all materials specularFactor 0 / ior 1 / roughness 0.9); y is up, the open side faces +z."""
    white, green, red = (0.725, 0.71, 0.68), (0.14, 0.45, 0.091), (0.63, 0.065, 0.05)

    def diffuse(c):
        return Material(baseColor=c, roughness=0.9, ior=1.0, specularWeight=0.0)

    materials = [diffuse(white), diffuse(white), diffuse(white), diffuse(green), diffuse(red), diffuse(white), diffuse(white),
                 Material(baseColor=(0.78, 0.78, 0.78), roughness=0.9, ior=1.0, specularWeight=0.0, emissionColor=(1.0, 1.0, 1.0), intensity=35.0)]
    P = lambda x, y, z: np.array([x, y, z], np.float32)  # noqa: E731
    x0, x1, y0, y1, z0, z1 = -1.0, 1.0, 0.0, 1.99, -1.04, 0.99
    meshes = [
        ("floor", _quad(P(x0, y0, z1), P(x1, y0, z1), P(x1, y0, z0), P(x0, y0, z0)), 0),
        ("ceiling", _quad(P(x0, y1, z0), P(x1, y1, z0), P(x1, y1, z1), P(x0, y1, z1)), 1),
        ("backWall", _quad(P(x0, y0, z0), P(x1, y0, z0), P(x1, y1, z0), P(x0, y1, z0)), 2),
        ("rightWall", _quad(P(x1, y0, z0), P(x1, y0, z1), P(x1, y1, z1), P(x1, y1, z0)), 3),
        ("leftWall", _quad(P(x0, y0, z1), P(x0, y0, z0), P(x0, y1, z0), P(x0, y1, z1)), 4),
        ("shortBox", _box([P(0.53, 0.0, 0.75), P(0.70, 0.0, 0.17), P(0.13, 0.0, 0.0), P(-0.05, 0.0, 0.57)], (0.0, 0.6, 0.0)), 5),
        ("tallBox", _box([P(-0.53, 0.0, 0.09), P(0.04, 0.0, -0.09), P(-0.14, 0.0, -0.67), P(-0.71, 0.0, -0.49)], (0.0, 1.2, 0.0)), 6),
        ("light", _quad(P(-0.24, 1.98, -0.22), P(0.23, 1.98, -0.22), P(0.23, 1.98, 0.16), P(-0.24, 1.98, 0.16)), 7),
    ]
    return {
        "name": "cornell_box",
        "meshes": [{"name": n, "triangles": np.stack(t).astype(np.float32), "material": m} for n, t, m in meshes],
        "instances": [{"mesh": i, "material": -1, "position": (0, 0, 0), "rotation": (0, 0, 0), "scale": (1, 1, 1)} for i in range(len(meshes))],
        "materials": materials,
        "lights": [],
        "camera": Camera(position=(0.0, 0.995, 3.9), forward=(0.0, 0.0, -1.0), horizontalFOV=45.0, focusDistance=5.0, defocusAngle=0.0),
        "settings": RenderSettings(useMIS=True, pathLength=path_length, backgroundColor=(0.0, 0.0, 0.0)),
    }


def planar_triangle_data(tris, scale=1.0):
    """D_TriangleData rows (n, 24) for flat-shaded geometry with texture coordinates: the geometric normal at every vertex, a
    unit tangent along the triangle's first edge and (u, v) = the vertex position projected on the two axes most orthogonal
    to the normal, times `scale`."""
    tris = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    out = flat_triangle_data(tris.reshape(-1, 9))
    n = out[:, 0:3]
    e = tris[:, 1] - tris[:, 0]
    t = e - (e * n).sum(1, keepdims=True) * n
    t = t / np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-30)
    out[:, 9:12] = out[:, 12:15] = out[:, 15:18] = t.astype(np.float32)
    drop = np.abs(n).argmax(1)
    for k in range(tris.shape[0]):
        ax = [a for a in range(3) if a != drop[k]]
        for v in range(3):
            out[k, 18 + 2 * v: 20 + 2 * v] = tris[k, v, ax] * scale
    return out


def textured_cornell(path_length=6):
    """the Cornell box with every kind of material map the reference samples (PathTracer.cu:373-411,
    263-269): sRGB RGBA8 checker base colour (with alpha < 1 squares: texture transparency) on the floor, a linear RGBA8
    normal map on the back wall, an RGBA8 roughness ramp on the short box, a glTF metallic-roughness map on the tall box and
    an RGBA32F emissive map on the light."""
    d = cornell_box(path_length)
    for m in d["meshes"]:
        m["triangle_data"] = planar_triangle_data(m["triangles"], scale=1.5)
    yy, xx = np.mgrid[0:64, 0:64]
    checker = np.zeros((64, 64, 4), np.uint8)
    on = ((xx // 8 + yy // 8) % 2).astype(bool)
    checker[..., 0], checker[..., 1], checker[..., 2] = np.where(on, 230, 60), np.where(on, 200, 90), np.where(on, 120, 200)
    checker[..., 3] = np.where((xx // 16 + yy // 16) % 4 == 0, 160, 255)
    normal = np.zeros((32, 32, 4), np.uint8)
    nx_, ny_ = 0.35 * np.sin(xx[:32, :32] * (2 * np.pi / 16)), 0.35 * np.cos(yy[:32, :32] * (2 * np.pi / 16))
    nz_ = np.sqrt(np.maximum(1.0 - nx_ ** 2 - ny_ ** 2, 0.0))
    normal[..., 0], normal[..., 1], normal[..., 2], normal[..., 3] = (nx_ * 0.5 + 0.5) * 255, (ny_ * 0.5 + 0.5) * 255, (nz_ * 0.5 + 0.5) * 255, 255
    rough = np.zeros((16, 16, 4), np.uint8); rough[..., 0] = np.linspace(40, 255, 16)[None, :]; rough[..., 1:] = 255
    mr = np.zeros((16, 16, 4), np.uint8); mr[..., 1] = np.linspace(30, 200, 16)[:, None]; mr[..., 2] = np.where((xx[:16, :16] // 4) % 2, 255, 40); mr[..., 0] = mr[..., 3] = 255
    emis = np.ones((8, 8, 4), np.float32)
    emis[..., 0], emis[..., 1], emis[..., 2] = 1.0, 0.55 + 0.45 * ((xx[:8, :8] + yy[:8, :8]) % 2), 0.3 + 0.7 * (xx[:8, :8] / 7.0)
    d["textures"] = [(checker, True), (normal, False), (rough, False), (mr, False), (emis, False)]
    mats = d["materials"]
    mats[0].baseColorMap = 0
    mats[2].normalMap = 1; mats[2].roughness = 0.35; mats[2].specularWeight = 1.0; mats[2].ior = 1.5
    mats[5].roughnessMap = 2; mats[5].roughness = 1.0; mats[5].specularWeight = 1.0; mats[5].ior = 1.5
    mats[6].metallicRoughnessMap = 3; mats[6].metalness = 1.0; mats[6].roughness = 1.0; mats[6].baseColor = (0.95, 0.8, 0.5)
    mats[7].emissiveMap = 4
    d["name"] = "cornell_textured"
    return d


def uv_sphere(nu=224, nv=224, radius=1.0, displace=None):
    """Tessellated sphere, 2*nu*nv triangles (224 x 224 x 2 = 100,352: config 1's mesh).  Pole rows keep their (degenerate)
    second triangle so the count is exact."""
    u = np.linspace(0.0, 2.0 * np.pi, nu + 1, dtype=np.float64)
    v = np.linspace(0.0, np.pi, nv + 1, dtype=np.float64)
    uu, vv = np.meshgrid(u, v, indexing="xy")
    r = np.full_like(uu, radius)
    if displace is not None:
        r = r * displace(uu, vv)
    p = np.stack([r * np.sin(vv) * np.cos(uu), r * np.cos(vv), r * np.sin(vv) * np.sin(uu)], -1).astype(np.float32)
    a, b, c, d = p[:-1, :-1], p[:-1, 1:], p[1:, 1:], p[1:, :-1]
    t0 = np.concatenate([a, c, d], -1).reshape(-1, 9)   # outward-facing winding
    t1 = np.concatenate([a, b, c], -1).reshape(-1, 9)
    return np.concatenate([t0, t1], 0).astype(np.float32)


def rock(seed, nu=71, nv=70):
    """One BLAS of config 3: a sphere displaced by a few random low-frequency lobes; 2*nu*(nv-1) non-degenerate triangles."""
    rs = np.random.RandomState(seed)
    k = rs.randint(1, 6, size=(6, 2)).astype(np.float64)
    amp = rs.uniform(0.02, 0.12, size=6)
    ph = rs.uniform(0.0, 2.0 * np.pi, size=(6, 2))

    def displace(uu, vv):
        d = np.ones_like(uu)
        for i in range(6):
            d = d + amp[i] * np.sin(k[i, 0] * uu + ph[i, 0]) * np.sin(k[i, 1] * vv + ph[i, 1]) * np.sin(vv)
        return d

    tris = uv_sphere(nu, nv, 1.0, displace)
    e0, e1 = tris[:, 3:6] - tris[:, 0:3], tris[:, 6:9] - tris[:, 0:3]
    keep = np.linalg.norm(np.cross(e0, e1), axis=1) > 1e-12   # drops the collapsed pole triangles
    return np.ascontiguousarray(tris[keep])


def instanced_scene(n_blas=1024, n_instances=1024, nu=71, nv=70, path_length=8, shared_blas=False):
    """Config 3: ~10 M triangles as n_blas displaced-sphere BLASes (9,798 triangles each at the defaults), n_instances
    instances on a jittered sqrt(n) x sqrt(n) grid with random Euler rotation and scale 0.5-1.5, a ground plane, one
    emissive quad above, and dielectric-glossy / gold-metal / translucent materials assigned round-robin."""
    materials = [
        Material(baseColor=(0.55, 0.55, 0.6), roughness=0.8),                                                  # 0 ground
        Material(baseColor=(1.0, 1.0, 1.0), emissionColor=(1.0, 0.96, 0.9), intensity=18.0),                     # 1 light
        Material(baseColor=(0.7, 0.2, 0.15), roughness=0.3),                                                    # 2 glossy dielectric
        Material(baseColor=(1.0, 0.78, 0.34), metalness=1.0, roughness=0.2),                                    # 3 gold
        Material(baseColor=(0.9, 0.95, 1.0), transmission=1.0, ior=1.5, roughness=0.1),                          # 4 translucent
    ]
    side = int(np.ceil(np.sqrt(n_instances)))
    spacing = 3.0
    half = 0.5 * side * spacing
    P = lambda x, y, z: np.array([x, y, z], np.float32)  # noqa: E731
    ground = np.stack(_quad(P(-half - 4, 0, half + 4), P(half + 4, 0, half + 4), P(half + 4, 0, -half - 4), P(-half - 4, 0, -half - 4)))
    lh, ls = 14.0, 0.35 * half
    light = np.stack(_quad(P(-ls, lh, -ls), P(ls, lh, -ls), P(ls, lh, ls), P(-ls, lh, ls)))
    meshes = [{"name": "ground", "triangles": ground, "material": 0}, {"name": "light", "triangles": light, "material": 1}]
    n_unique = 1 if shared_blas else n_blas
    for k in range(n_unique):
        meshes.append({"name": f"rock{k}", "triangles": rock(1000 + k, nu, nv), "material": 2 + k % 3})
    rs = np.random.RandomState(7)
    instances = [{"mesh": 0, "material": -1, "position": (0, 0, 0), "rotation": (0, 0, 0), "scale": (1, 1, 1)},
                 {"mesh": 1, "material": -1, "position": (0, 0, 0), "rotation": (0, 0, 0), "scale": (1, 1, 1)}]
    for i in range(n_instances):
        gx, gz = i % side, i // side
        jitter = rs.uniform(-0.6, 0.6, size=2)
        s = float(rs.uniform(0.5, 1.5))
        rot = rs.uniform(0.0, 360.0, size=3)
        pos = (float((gx + 0.5) * spacing - half + jitter[0]), 1.25 * s, float((gz + 0.5) * spacing - half + jitter[1]))
        instances.append({"mesh": 2 + (i % n_unique), "material": 2 + i % 3, "position": pos, "rotation": tuple(float(r) for r in rot), "scale": (s, s, s)})
    cam_pos = np.array([0.0, 0.45 * half + 6.0, half + 10.0])
    fwd = np.array([0.0, 1.0, 0.0]) - cam_pos
    fwd = fwd / np.linalg.norm(fwd)
    return {
        "name": f"instanced_{n_unique}blas_{n_instances}inst",
        "meshes": meshes, "instances": instances, "materials": materials, "lights": [],
        "camera": Camera(position=tuple(cam_pos), forward=tuple(fwd), horizontalFOV=50.0, focusDistance=float(np.linalg.norm(cam_pos)), defocusAngle=0.0),
        "settings": RenderSettings(useMIS=True, pathLength=path_length, backgroundColor=(0.02, 0.03, 0.05)),
    }


def procedural_sky(w=4096, h=2048, sun_dir=(0.35, 0.6, 0.72), sun_radiance=5.0e4, sun_angle_deg=0.5):
    """Config 5's environment: analytic gradient sky + sun disc, RGBA32F equirect in the reference's (u, v) convention
    (SampleBackground, PathTracer.cu:40-58: u = (atan2(z, x) + pi) / 2pi, v = 1 - (asin(y) + pi/2) / pi)."""
    u = (np.arange(w, dtype=np.float64) + 0.5) / w
    v = (np.arange(h, dtype=np.float64) + 0.5) / h
    theta = u * 2.0 * np.pi - np.pi
    phi = (1.0 - v) * np.pi - 0.5 * np.pi
    tt, pp = np.meshgrid(theta, phi, indexing="xy")
    d = np.stack([np.cos(pp) * np.cos(tt), np.sin(pp), np.cos(pp) * np.sin(tt)], -1)
    up = np.clip(d[..., 1], 0.0, 1.0)
    horizon, zenith, ground = np.array([0.9, 0.95, 1.0]), np.array([0.15, 0.35, 0.9]), np.array([0.25, 0.22, 0.2])
    sky = horizon[None, None] * (1.0 - up[..., None]) ** 3 + zenith[None, None] * (1.0 - (1.0 - up[..., None]) ** 3)
    img = np.where(d[..., 1:2] >= 0.0, sky, ground[None, None] * (0.4 + 0.6 * np.exp(6.0 * d[..., 1:2])))
    s = np.asarray(sun_dir, np.float64)
    s = s / np.linalg.norm(s)
    in_sun = (d @ s) > np.cos(np.radians(sun_angle_deg))
    img = np.where(in_sun[..., None], sun_radiance, img)
    out = np.ones((h, w, 4), np.float32)
    out[..., :3] = img
    return out


def test_triangles(n, seed=12345, grid=1000):
    """Config 4: the NexusBVH benchmark's synthetic mesh (vendor/NexusBVH/Test/src/Main.cpp:31-64): n small triangles in
    random cells of a grid^3 lattice over [0, 10]^3 (same distribution; numpy's MT19937 stream rather than libstdc++'s)."""
    rs = np.random.RandomState(seed)
    cell = np.float32(10.0 / grid)
    out = np.empty((n, 9), np.float32)
    chunk = 4_000_000
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        base = rs.randint(0, grid, size=(m, 3)).astype(np.float32) * cell
        v0 = base + rs.uniform(0.1 * cell, 0.9 * cell, size=(m, 3)).astype(np.float32)
        v1 = v0 + rs.uniform(-0.4 * cell, 0.4 * cell, size=(m, 3)).astype(np.float32)
        v2 = v0 + rs.uniform(-0.4 * cell, 0.4 * cell, size=(m, 3)).astype(np.float32)
        near1 = np.linalg.norm(v1 - v0, axis=1) < 0.1 * cell
        near2 = np.linalg.norm(v2 - v0, axis=1) < 0.1 * cell
        v1[near1, 0] += 0.2 * cell
        v2[near2, 1] += 0.2 * cell
        out[s:s + m, 0:3], out[s:s + m, 3:6], out[s:s + m, 6:9] = v0, v1, v2
    return out


def build(ctx, desc, resolution, blas=None):
    """Instantiates a scene description through the reference-shaped host API.  blas: optional {mesh index: (nodes_dev, node_count,
    prim_idx_dev, bounds)} of BLASes built elsewhere (nexus_b200.multigpu.build_scene_sharded); those meshes are imported, the
    others built here."""
    scene = Scene(ctx, resolution)
    am = scene.GetAssetManager()
    for t in desc.get("textures", []):          # [(pixels, sRGB)]: indices in order of appearance
        am.AddTexture(t[0], t[1])
    for m in desc["materials"]:
        am.AddMaterial(m)
    for k, m in enumerate(desc["meshes"]):
        if blas is not None and k in blas:
            am.AddMeshPrebuilt(m["name"], m["material"], m["triangles"], m.get("triangle_data"), *blas[k])
        else:
            am.AddMesh(m["name"], m["material"], m["triangles"], m.get("triangle_data"))
    for i in desc["instances"]:
        if "matrix" in i:      # imported assets carry the accumulated node transform (nexus_b200.gltf)
            scene.CreateMeshInstanceMatrix(i["mesh"], i["matrix"], i.get("material", -1))
        else:
            scene.CreateMeshInstance(i["mesh"], i.get("material", -1), i["position"], i["rotation"], i["scale"])
    for l in desc.get("lights", []):
        scene.AddLight(l)
    scene.SetCamera(desc["camera"])
    scene.SetRenderSettings(desc["settings"])
    if desc.get("hdr") is not None:
        scene.AddHDRMap(desc["hdr"])
    scene.Update()
    return scene


def camera_rays(desc_camera, resolution, n=None, seed=0):
    """Deterministic pinhole rays through pixel centres (for traversal parity batches)."""
    w, h = resolution
    pos = np.asarray(desc_camera.position, np.float64)
    fwd = np.asarray(desc_camera.forward, np.float64)
    right = np.cross(fwd, [0.0, 1.0, 0.0])
    up = np.cross(right, fwd)
    half_w = desc_camera.focusDistance * np.tan(np.radians(desc_camera.horizontalFOV / 2.0))
    half_h = half_w / (w / h)
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    x = (xs.ravel() + 0.5) / w * 2.0 - 1.0
    y = (ys.ravel() + 0.5) / h * 2.0 - 1.0
    d = fwd[None] * desc_camera.focusDistance + x[:, None] * half_w * right[None] + y[:, None] * half_h * up[None]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    o = np.broadcast_to(pos, d.shape)
    if n is not None and n < len(d):
        idx = np.random.RandomState(seed).choice(len(d), n, replace=False)
        o, d = o[idx], d[idx]
    return o.astype(np.float32), d.astype(np.float32)


def flat_triangle_data(tris):
    """D_TriangleData rows (n, 24) with the geometric normal at all three vertices, zero tangents and texture coordinates."""
    tris = np.asarray(tris, np.float32).reshape(-1, 9)
    n = np.cross(tris[:, 3:6] - tris[:, 0:3], tris[:, 6:9] - tris[:, 0:3]).astype(np.float32)
    l = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(l > 0, n / np.maximum(l, 1e-30), 0.0).astype(np.float32)
    out = np.zeros((tris.shape[0], 24), np.float32)
    out[:, 0:3] = out[:, 3:6] = out[:, 6:9] = n
    return out


def with_triangle_data(desc):
    """Adds explicit flat shading data to every mesh so two renderers can be fed byte-identical inputs."""
    for m in desc["meshes"]:
        if m.get("triangle_data") is None:
            m["triangle_data"] = flat_triangle_data(m["triangles"])
    return desc
