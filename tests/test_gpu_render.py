"""GPU checks of the wavefront renderer through the C ABI.

Bar (north_star): converged images within a stated RMSE of the reference.  Per-frame images cannot be compared: the
reference seeds its RNG with racing queue slots and is not reproducible run to run.  The stated bound:
on 8x8-pixel block means at 2048 spp, relative RMSE(ours, reference) <= 3 x the reference's own relative RMSE between two
independent 2048-spp halves (its noise floor, stored in the golden file), and overall mean radiance within 1 %.
"""
import os

import numpy as np
import pytest

import nexus_b200 as nx
import oracle_lib as O
from golden_cases import render_cases
from nexus_b200 import scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "render_ref.npz")


def _blocks(img, block):
    h, w = img.shape[0] // block, img.shape[1] // block
    return img.reshape(h, block, w, block, 3).mean((1, 3))


@pytest.mark.parametrize("case", render_cases(), ids=lambda c: c[0])
def test_converged_image_matches_reference(ctx, case):
    name, desc, res, spp, block = case
    gold = np.load(GOLD)
    ref_a, ref_b = gold[name + "/mean_a"].astype(np.float64), gold[name + "/mean_b"].astype(np.float64)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    pt.Render(scene, frames=spp)
    ours = _blocks(pt.ReadAccumulation().astype(np.float64), block)
    ref = 0.5 * (ref_a + ref_b)                      # 2 x spp samples of the reference
    floor = np.sqrt(((ref_a - ref_b) ** 2).mean()) / ref.mean()
    rmse = np.sqrt(((ours - ref) ** 2).mean()) / ref.mean()
    assert rmse <= 3.0 * floor + 1e-3, (rmse, floor)
    assert abs(ours.mean() - ref.mean()) <= 0.01 * ref.mean(), (ours.mean(), ref.mean())
    for c in range(3):
        assert abs(ours[..., c].mean() - ref[..., c].mean()) <= 0.015 * ref[..., c].mean()
    pt.close(); scene.close()


def test_frames_are_reproducible_and_additive(ctx):
    """Ours keys the RNG on (pixel, frame, bounce): a frame is a pure function of its index, so rendering frames 1..8 in one
    call, in two calls, or on two renderers and summing (the multi-GPU sample partition) gives the same image up to float
    summation order."""
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=5))
    res = (96, 64)
    scene = scenes.build(ctx, desc, res)
    a = nx.PathTracer(ctx, res); a.Render(scene, frames=8, firstFrame=1)
    b = nx.PathTracer(ctx, res); b.Render(scene, frames=4, firstFrame=1); b.Render(scene, frames=4, firstFrame=5)
    c1 = nx.PathTracer(ctx, res); c1.Render(scene, frames=4, firstFrame=1)
    c2 = nx.PathTracer(ctx, res); c2.Render(scene, frames=4, firstFrame=5)
    ia, ib = a.ReadAccumulation(), b.ReadAccumulation()
    ic = 0.5 * (c1.ReadAccumulation() + c2.ReadAccumulation())
    assert a.GetFrameNumber() == 8 and b.GetFrameNumber() == 8
    assert np.allclose(ia, ib, rtol=1e-4, atol=1e-5) and np.allclose(ia, ic, rtol=1e-4, atol=1e-5)
    sa, sb = a.Stats(), b.Stats()
    assert sa["extension_rays"] >= 8 * res[0] * res[1]
    for p in (a, b, c1, c2):
        p.close()
    scene.close()


def test_queue_accounting_and_settings(ctx):
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=3))
    res = (64, 64)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    pt.SetProfiling(events=True, work=True)
    pt.Render(scene, frames=2)
    st, pr = pt.Stats(), pt.Profile()
    # kernels per frame: generate + trace + pathLength x (shade + shadow trace) + (pathLength - 1) x trace + totals
    assert st["kernel_launches"] == 2 * (2 + 3 * 2 + 2 + 1)
    assert pr["trace_closest"]["launches"] == 2 * 3 and pr["trace_any"]["launches"] == 2 * 3 and pr["shade"]["launches"] == 2 * 3
    assert pr["closest_work"]["rays"] == st["extension_rays"] and pr["any_work"]["rays"] == st["shadow_rays"]
    assert pr["closest_work"]["nodes"] >= st["extension_rays"]          # every ray visits at least the TLAS root
    assert st["shaded_hits"] <= st["extension_rays"] and st["shadow_rays"] <= st["shaded_hits"]
    # pathLength 1 and no MIS: only directly visible emitters contribute, no shadow rays at all
    desc["settings"] = nx.RenderSettings(useMIS=False, pathLength=1)
    scene.SetRenderSettings(desc["settings"])
    pt.SetProfiling(events=False, work=False)
    pt.ResetFrameNumber(); pt.Render(scene, frames=4)
    st = pt.Stats(); img = pt.ReadAccumulation()
    assert st["shadow_rays"] == 0 and st["extension_rays"] == 4 * res[0] * res[1]
    lit = img.max(axis=2) > 0
    assert 0 < lit.sum() < 0.1 * lit.size and np.allclose(img[lit].max(), 35.0, rtol=1e-3)   # the 35x emitter, nothing else
    # background: open the scene up and every miss returns the constant background
    pt.close(); scene.close()


def test_resize_and_outputs(ctx, tmp_path):
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=4))
    scene = scenes.build(ctx, desc, (80, 48))
    pt = nx.PathTracer(ctx, (40, 24))
    with pytest.raises(nx.NexusError):
        pt.Render(scene, frames=1)                  # resolution mismatch is an error, not a crash
    pt.OnResize((80, 48))
    pt.Render(scene, frames=16)
    # and the other way round: an existing scene follows a resize end to end (Camera::OnResize), nothing is rebuilt
    scene.OnResize((64, 40)); pt.OnResize((64, 40))
    pt.Render(scene, frames=2)
    assert pt.ReadAccumulation().shape == (40, 64, 3)
    scene.OnResize((80, 48)); pt.OnResize((80, 48))
    pt.Render(scene, frames=16)
    img = pt.ReadAccumulation()
    assert img.shape == (48, 80, 3) and np.isfinite(img).all() and img.mean() > 0.05
    rgba = pt.ReadRGBA8(scene)
    assert rgba.shape == (48, 80) and (rgba >> 24 == 0xff).all() and (rgba & 0xffffff).any()
    nx.write_pfm(tmp_path / "c.pfm", img); nx.write_exr(tmp_path / "c.exr", img)
    assert os.path.getsize(tmp_path / "c.pfm") > 48 * 80 * 12 and os.path.getsize(tmp_path / "c.exr") > 48 * 80 * 12
    # orientation: the Cornell box's light is on the ceiling.  The accumulation has row 0 at the bottom (the reference's pixel order),
    # PFM stores rows bottom to top, EXR scanline 0 is the top: the emitter must be in the LAST rows of the PFM body and in the FIRST
    # scanlines of the EXR file
    assert img[36:].max() > 4.0 * img[:12].max()
    body = np.frombuffer(open(tmp_path / "c.pfm", "rb").read().split(b"\n", 3)[3], "<f4").reshape(48, 80, 3)
    assert body[36:].max() > 4.0 * body[:12].max()
    raw = open(tmp_path / "c.exr", "rb").read()
    chunk = 8 + 80 * 12
    lines = np.frombuffer(raw[-48 * chunk:], np.uint8).reshape(48, chunk)[:, 8:].copy().view("<f4").reshape(48, 3, 80)   # scanline, (B, G, R), x
    assert lines[:12].max() > 4.0 * lines[36:].max()
    pt.close(); scene.close()


def test_hdr_environment_and_punctual_lights(ctx):
    """Config 5's ingredients: equirect HDR background through a CUDA texture, no emissive geometry; plus a point light."""
    sky = np.zeros((8, 16, 4), np.float32); sky[..., :3] = (0.5, 1.0, 2.0); sky[..., 3] = 1.0
    tri = np.array([[-50, 0, 50, 50, 0, 50, 50, 0, -50], [-50, 0, 50, 50, 0, -50, -50, 0, -50]], np.float32)   # ground quad, faces +y
    desc = scenes.with_triangle_data({
        "meshes": [{"name": "g", "triangles": tri, "material": 0}],
        "instances": [{"mesh": 0, "material": -1, "position": (0, 0, 0), "rotation": (0, 0, 0), "scale": (1, 1, 1)}],
        "materials": [nx.Material(baseColor=(0.5, 0.5, 0.5), roughness=1.0, specularWeight=0.0)], "lights": [],
        "camera": nx.Camera(position=(0, 2, 0), forward=(0, 0, -1), horizontalFOV=60.0),
        "settings": nx.RenderSettings(useMIS=True, pathLength=2), "hdr": sky})
    res = (64, 64)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res); pt.Render(scene, frames=256)
    img = pt.ReadAccumulation()
    top = img[40:, :, :].reshape(-1, 3).mean(0)         # rows above the horizon see the sky directly
    assert np.allclose(top, (0.5, 1.0, 2.0), rtol=1e-3)
    # white furnace-ish: a grey Lambertian ground under a constant sky reflects albedo x sky (one bounce, no occluders)
    bottom = img[:20, :, :].reshape(-1, 3).mean(0)
    assert np.allclose(bottom, 0.5 * np.array([0.5, 1.0, 2.0]), rtol=0.03), bottom
    pt.close(); scene.close()
    # a point light over the same ground, black background: radiance = albedo/pi * I * cos / d^2 straight below the light
    desc2 = dict(desc); desc2["hdr"] = None
    desc2["lights"] = [nx.Light(nx.Light.POINT, position=(0, 1, -3), color=(1, 1, 1), intensity=10.0)]
    desc2["camera"] = nx.Camera(position=(0, 3, -3), forward=(0, -1, 0), right=(1, 0, 0), horizontalFOV=20.0)
    scene = scenes.build(ctx, desc2, res)
    pt = nx.PathTracer(ctx, res); pt.Render(scene, frames=64)
    img = pt.ReadAccumulation()
    centre = img[30:34, 30:34].mean()
    assert abs(centre - 0.5 / np.pi * 10.0) <= 0.03 * (0.5 / np.pi * 10.0), centre
    pt.close(); scene.close()


@pytest.mark.parametrize("mode", range(6), ids=["none", "aces", "uncharted2", "agx", "agx_golden", "agx_punchy"])
def test_display_transform_matches_reference(ctx, mode):
    """exposure + tone curve + gamma + RGBA8 pack of the product against the reference kernel's own
    render buffers (golden) and against the numpy oracle.  Integer output; both sides evaluate log2 / pow with the hardware
    approximations, so a channel may differ by one code where the value sits on a quantisation boundary: at most 1 LSB on at
    most 1 % of the channels, exact alpha."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "..", "oracle"))
    import oracle_display as D
    from golden_cases import DISPLAY_EXPOSURES
    g = np.load(os.path.join(os.path.dirname(GOLD), "display_ref.npz"))
    for e in DISPLAY_EXPOSURES:
        got = nx.display_transform(ctx, g["image"], mode, e)
        for want in (g[f"m{mode}_e{e:+.1f}"], D.display(g["image"], mode, e)):
            diff = np.abs(D.unpack(got) - D.unpack(want))
            assert diff.max() <= 1 and (diff > 0).mean() <= 0.01, (mode, e, diff.max(), (diff > 0).mean())
        assert (got >> 24 == 0xff).all()


def test_read_rgba8_uses_the_scene_tone_mapping(ctx):
    """ReadRGBA8 = display transform of the running mean with the scene's RenderSettings (toneMapping, exposure)."""
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=4))
    res = (64, 48)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    pt.Render(scene, frames=8)
    lin = pt.ReadAccumulation()
    for mode, e in ((nx.TONE_NONE, 0.0), (nx.TONE_ACES, 1.0), (nx.TONE_AGX_PUNCHY, -0.5)):
        rs = desc["settings"]; rs.toneMapping, rs.exposure = mode, e
        scene.SetRenderSettings(rs)
        assert (pt.ReadRGBA8(scene) == nx.display_transform(ctx, lin, mode, e)).all()
    with pytest.raises(nx.NexusError):
        nx.display_transform(ctx, lin, 9, 0.0)
    pt.close(); scene.close()


def test_dynamic_scene_updates_rebuild_the_tlas(ctx):
    """MeshInstance::SetTransform / AssetManager::InvalidateMaterial mark the scene dirty and the next
    Update rebuilds the TLAS, the traversal records and the light list (Scene::Update, Scene.cpp:34-63; the reference rebuilds
    its TLAS from scratch on any instance change).  An edited scene must behave exactly like one created in the final state:
    identical closest hits (bit for bit) and an identical frame (the RNG is keyed on pixel, frame and bounce)."""
    res = (160, 120)
    ctx.SetInstanceMerging(False)     # bit-for-bit needs the same trees in both scenes; the merged-BLAS variant follows below

    def make(final):
        desc = scenes.with_triangle_data(scenes.instanced_scene(n_blas=4, n_instances=9, nu=16, nv=16, path_length=4))
        if final:
            desc["instances"][3]["position"] = (1.5, 2.0, -0.5); desc["instances"][3]["rotation"] = (10.0, 200.0, 35.0); desc["instances"][3]["scale"] = (1.3, 1.3, 1.3)
            desc["materials"][2].baseColor = (0.1, 0.8, 0.2); desc["materials"][2].roughness = 0.6
        return desc, scenes.build(ctx, desc, res)

    desc_a, a = make(False)
    desc_b, b = make(True)
    o, d = scenes.camera_rays(desc_a["camera"], res)
    rays = nx.make_rays(o, d)
    before = a.TraceClosest(rays)
    # edit scene a into scene b's state through the host API
    nx.MeshInstance(a, 3, desc_a["instances"][3]["mesh"]).SetTransform((1.5, 2.0, -0.5), (10.0, 200.0, 35.0), (1.3, 1.3, 1.3))
    m = desc_a["materials"][2]; m.baseColor = (0.1, 0.8, 0.2); m.roughness = 0.6
    a.GetAssetManager().InvalidateMaterial(2, m)
    a.Update()
    after, want = a.TraceClosest(rays), b.TraceClosest(rays)
    assert (after.view(np.uint8) == want.view(np.uint8)).all()
    assert not (after.view(np.uint8) == before.view(np.uint8)).all()          # the edit is visible
    pa, pb = nx.PathTracer(ctx, res), nx.PathTracer(ctx, res)
    pa.Render(a, frames=2, firstFrame=1); pb.Render(b, frames=2, firstFrame=1)
    ia, ib = pa.ReadAccumulation(), pb.ReadAccumulation()
    assert np.allclose(ia, ib, rtol=1e-4, atol=1e-5)                          # float atomics: summation order only
    pa.close(); pb.close(); a.close(); b.close()
    ctx.SetInstanceMerging(True)

    # With instance merging (the default) the rocks of this scene share BLASes (4 meshes, 9 instances) and keep them, but the ground and
    # the light - one instance each - start in the merged world-space BLAS.  Moving the ground takes it out: it gets a BLAS of its own
    # (nothing else is rebuilt but the merged BLAS, once), the hits are those of a scene created in the final state, and a second move
    # only rebuilds the TLAS.
    desc_c = scenes.with_triangle_data(scenes.instanced_scene(n_blas=4, n_instances=9, nu=16, nv=16, path_length=4))
    c = scenes.build(ctx, desc_c, res)
    assert (c.ExportTlasEntries() == 0xffffffff).sum() == 1 and c.ExportMerged(bounds=False)["bvh"].primCount == 4
    nx.MeshInstance(c, 0, desc_c["instances"][0]["mesh"]).SetTransform((0.0, -0.25, 0.0), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
    c.Update()
    assert c.ExportMerged(bounds=False)["bvh"].primCount == 2                 # the light's two triangles are what is left in it
    desc_d = scenes.with_triangle_data(scenes.instanced_scene(n_blas=4, n_instances=9, nu=16, nv=16, path_length=4))
    desc_d["instances"][0]["position"] = (0.0, -0.25, 0.0)
    d = scenes.build(ctx, desc_d, res)
    hc, hd = c.TraceClosest(rays), d.TraceClosest(rays)
    same = (hc["prim"] == hd["prim"]) & (hc["instance"] == hd["instance"])
    assert same.mean() >= 0.999 and np.allclose(hc["t"][same], hd["t"][same], rtol=2e-5)
    assert (hc["instance"][hc["prim"] != 0xffffffff] == 0).any()                 # the moved ground is being hit
    c.close(); d.close()


# ------------------------------------------------------------------ material maps ----
def test_textured_converged_image_matches_reference(ctx):
    """Base colour (sRGB RGBA8, with transparency), normal, roughness, metallic-roughness and emissive (RGBA32F) maps: converged
    block means against the reference renderer's (tests/golden/render_textured_ref.npz, scripts/make_golden_textured.py),
    same bound as the untextured cases: relative RMSE <= 3 x the reference's own noise floor, means within 1 - 1.5 %."""
    from golden_cases import textured_render_cases
    gold = np.load(os.path.join(os.path.dirname(GOLD), "render_textured_ref.npz"))
    for name, desc, res, spp, block in textured_render_cases():
        ref_a, ref_b = gold[name + "/mean_a"].astype(np.float64), gold[name + "/mean_b"].astype(np.float64)
        scene = scenes.build(ctx, desc, res)
        pt = nx.PathTracer(ctx, res)
        pt.Render(scene, frames=spp)
        ours = _blocks(pt.ReadAccumulation().astype(np.float64), block)
        ref = 0.5 * (ref_a + ref_b)
        floor = np.sqrt(((ref_a - ref_b) ** 2).mean()) / ref.mean()
        rmse = np.sqrt(((ours - ref) ** 2).mean()) / ref.mean()
        assert rmse <= 3.0 * floor + 1e-3, (rmse, floor)
        assert abs(ours.mean() - ref.mean()) <= 0.01 * ref.mean(), (ours.mean(), ref.mean())
        for c in range(3):
            assert abs(ours[..., c].mean() - ref[..., c].mean()) <= 0.015 * ref[..., c].mean()
        pt.close(); scene.close()


def test_constant_maps_equal_scaled_materials(ctx):
    """A constant map multiplies the material parameter (PathTracer.cu:394-411): a scene whose floor carries a uniform 0.5-grey
    base-colour map and whose box a uniform roughness map renders the same frames as the scene with the parameters pre-multiplied
    and no maps (same RNG keys, so the comparison is per pixel, not statistical)."""
    res = (96, 96)

    def make(maps):
        d = scenes.cornell_box(path_length=4)
        for m in d["meshes"]:
            m["triangle_data"] = scenes.planar_triangle_data(m["triangles"])
        half = np.full((4, 4, 4), 0.5, np.float32); half[..., 3] = 1.0
        quarter = np.full((4, 4, 4), 0.25, np.float32)
        d["materials"][5].specularWeight = 1.0; d["materials"][5].ior = 1.5; d["materials"][5].roughness = 0.8
        if maps:
            d["textures"] = [(half, False), (quarter, False)]
            d["materials"][0].baseColorMap = 0
            d["materials"][5].roughnessMap = 1
        else:
            d["materials"][0].baseColor = tuple(0.5 * c for c in d["materials"][0].baseColor)
            d["materials"][5].roughness = 0.8 * 0.25
        return d, scenes.build(ctx, d, res)

    (da, a), (db, b) = make(True), make(False)
    pa, pb = nx.PathTracer(ctx, res), nx.PathTracer(ctx, res)
    pa.Render(a, frames=4, firstFrame=1); pb.Render(b, frames=4, firstFrame=1)
    ia, ib = pa.ReadAccumulation(), pb.ReadAccumulation()
    assert ia.mean() > 0 and np.allclose(ia, ib, rtol=2e-3, atol=1e-5)
    pa.close(); pb.close(); a.close(); b.close()
    # a material that names a texture the scene does not have is rejected at Update
    d = scenes.with_triangle_data(scenes.cornell_box(path_length=2)); d["materials"][0].baseColorMap = 3
    with pytest.raises(nx.NexusError):
        scenes.build(ctx, d, (8, 8)).TraceClosest(nx.make_rays(np.zeros((1, 3), np.float32), np.array([[0, 0, -1]], np.float32)))


def test_tlas_refit_gives_the_hits_of_a_rebuilt_scene(ctx):
    """with nx_ctx_set_tlas_refit the Update after moving instances refits the TLAS in place (same topology and
    leaf order) whenever the set of TLAS entries is unchanged; the hits are those of a scene created in the final state, byte for byte
    (the TLAS only culls: ids, t, u, v come from the triangle tests).  The default stays the reference's rebuild."""
    res = (160, 120)
    rs = np.random.RandomState(5)

    def desc_with(moves):
        d = scenes.with_triangle_data(scenes.instanced_scene(n_blas=5, n_instances=24, nu=14, nv=12, path_length=3))
        for i, (pos, rot, sc) in moves.items():
            d["instances"][i]["position"], d["instances"][i]["rotation"], d["instances"][i]["scale"] = pos, rot, sc
        return d

    def random_moves(ids):
        return {i: (tuple(float(v) for v in rs.uniform(-6, 6, 3) + np.array([0, 3.0, 0])), tuple(float(v) for v in rs.uniform(0, 360, 3)), (float(rs.uniform(0.6, 1.4)),) * 3) for i in ids}

    base = desc_with({})
    o, d = scenes.camera_rays(base["camera"], res)
    rays = nx.make_rays(o, d)
    for refit in (True, False):
        ctx.SetTlasRefit(refit)
        try:
            s = scenes.build(ctx, base, res)
            assert s.TlasHistory() == (1, 0)
            moves = {}
            for step, ids in enumerate(([4, 9, 17], [4, 9, 17], [9, 21], [3, 4, 9, 17, 21])):
                moves.update(random_moves(ids))
                for i in ids:
                    nx.MeshInstance(s, i, base["instances"][i]["mesh"]).SetTransform(*moves[i])
                s.Update()
                got = s.TraceClosest(rays)
                ctx.SetTlasRefit(False)
                fresh = scenes.build(ctx, desc_with(moves), res)
                ctx.SetTlasRefit(refit)
                want = fresh.TraceClosest(rays)
                assert (got.view(np.uint8) == want.view(np.uint8)).all(), (refit, step)
                fresh.close()
            builds, refits = s.TlasHistory()
            # steps 0, 2 and 3 move instances for the first time: they leave the merged BLAS or - shared meshes - stay entries; the entry
            # set changes only when a single-use instance moves for the first time.  Step 1 moves the same instances again: a pure refit.
            assert builds + refits == 5 and (refits >= 1 if refit else refits == 0), (refit, builds, refits)
            s.close()
        finally:
            ctx.SetTlasRefit(False)


def test_present_is_a_pipelined_read_rgba8(ctx):
    """nx_renderer_present / present_wait (the reference's PBO path: Render returns without synchronising and the display reads the
    pixel buffer a frame later, PathTracer.cpp:170-199): the image and queue totals a ticket delivers are those of the frames
    submitted before Present, even though later frames were queued behind it before anybody waited."""
    import torch
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=4))
    res = (128, 96)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    bufs = [torch.zeros((res[1], res[0]), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32) for _ in range(2)]
    tickets, want_img, want_stats = [], [], []
    # reference answers with the synchronous calls
    for f in (1, 2, 3):
        pt.Render(scene, frames=1, firstFrame=f)
        want_stats.append(pt.Stats()); want_img.append(pt.ReadRGBA8(scene).copy())
    pt.ResetFrameNumber()
    got_img, got_stats = [], []
    for f in (1, 2, 3):
        pt.Render(scene, frames=1, firstFrame=f)
        tickets.append(pt.Present(scene, bufs[(f - 1) & 1]))
        if f > 1:       # wait for the previous frame only after this one has been queued
            got_stats.append(pt.PresentWait(tickets[-2])); got_img.append(bufs[(f - 2) & 1].copy())
    got_stats.append(pt.PresentWait(tickets[-1])); got_img.append(bufs[0].copy())
    assert tickets == [0, 1, 0]
    for k in range(3):
        assert (got_img[k] == want_img[k]).all(), k
        for key in ("extension_rays", "shadow_rays", "shaded_hits", "frames"):
            assert got_stats[k][key] == want_stats[k][key], (k, key)
    fresh = nx.PathTracer(ctx, res)
    with pytest.raises(nx.NexusError):
        fresh.PresentWait(1)
    fresh.close(); pt.close(); scene.close()


def test_present_device_writes_the_display_image_into_caller_memory(ctx):
    """nx_renderer_present_device: the headless form of the reference's OpenGL display path (the mapped pixel buffer's device pointer
    handed to AccumulateKernel, PixelBuffer.cpp:4-40 / Renderer.cpp:41-48): the RGBA8 image lands in caller-supplied device memory, queued
    behind the frame, and equals ReadRGBA8; host memory is refused."""
    import torch
    desc = scenes.with_triangle_data(scenes.cornell_box(path_length=3))
    res = (96, 64)
    scene = scenes.build(ctx, desc, res)
    pt = nx.PathTracer(ctx, res)
    pbo = torch.zeros(res[0] * res[1], dtype=torch.int32, device="cuda")       # stands in for the mapped pixel buffer object
    torch.cuda.synchronize()
    pt.Render(scene, frames=2, firstFrame=1)
    pt.PresentDevice(scene, pbo.data_ptr())                                      # no host synchronisation in between
    ctx.synchronize()
    want = pt.ReadRGBA8(scene)
    assert (pbo.cpu().numpy().view(np.uint32).reshape(res[1], res[0]) == want).all() and want.any()
    host = np.zeros(res[0] * res[1], np.uint32)
    with pytest.raises(nx.NexusError):
        pt.PresentDevice(scene, host.ctypes.data)
    pt.close(); scene.close()


def test_pixel_query_returns_the_primary_hit_instance(ctx):
    """PathTracer::SetPixelQuery / SynchronizePixelQuery (PathTracer.cpp:221-240, PathTracer.cu:150-151, 459-460): the instance the
    pixel's primary ray hits in the next frame, -1 for a miss.  Checked on pixels whose whole 3x3 neighbourhood of centre rays
    agrees (the primary ray is jittered inside the pixel), at a tiled (8x4) and a non-tiled resolution."""
    desc = scenes.with_triangle_data(scenes.instanced_scene(n_blas=3, n_instances=12, nu=12, nv=12, path_length=2))
    for res in ((160, 120), (150, 90)):
        scene = scenes.build(ctx, desc, res)
        pt = nx.PathTracer(ctx, res)
        assert not pt.PixelQueryPending() and pt.SynchronizePixelQuery() == -1
        o, d = scenes.camera_rays(desc["camera"], res)
        hits = scene.TraceClosest(nx.make_rays(o, d))
        inst = np.where(hits["prim"] == 0xffffffff, -1, hits["instance"].astype(np.int64)).reshape(res[1], res[0])
        stable = np.ones_like(inst, bool)
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                stable &= np.roll(np.roll(inst, dy, 0), dx, 1) == inst
        stable[0, :] = stable[-1, :] = False; stable[:, 0] = stable[:, -1] = False
        ys, xs = np.nonzero(stable)
        rs = np.random.RandomState(3)
        picks = list(rs.choice(len(ys), 24, replace=False))
        seen = set()
        for frame, i in enumerate(picks, start=1):
            x, y = int(xs[i]), int(ys[i])
            pt.SetPixelQuery(x, y)
            assert pt.PixelQueryPending()
            pt.Render(scene, frames=1, firstFrame=frame)
            got = pt.SynchronizePixelQuery()
            assert got == inst[y, x] and pt.GetSelectedInstance() == got and not pt.PixelQueryPending(), (res, x, y, got, inst[y, x])
            seen.add(got)
        assert len(seen) >= 3                        # several different instances (and usually the background) were picked
        with pytest.raises(nx.NexusError):
            pt.SetPixelQuery(res[0], 0)
        pt.close(); scene.close()


def test_scene_and_renderer_cycles_return_their_device_memory():
    """Create / render / close, six times on one context, then close the context: the device's free memory after the sixth cycle is
    what it was after the first (pools are warm from then on, nothing is lost per cycle), and closing the context returns the pools.
    The reference frees its buffers in destructors (Scene / PathTracer / DeviceBuffer); here ownership is explicit handles behind the C
    ABI, so the balance is checked on the device."""
    import torch
    def free_mb():
        torch.cuda.synchronize()
        return torch.cuda.mem_get_info()[0] / 2 ** 20
    before = free_mb()
    ctx = nx.Context(0)
    desc = scenes.instanced_scene(n_blas=6, n_instances=40, nu=24, nv=20, path_length=4)
    res = (160, 96)
    marks = []
    for _ in range(6):
        scene = scenes.build(ctx, desc, res)
        pt = nx.PathTracer(ctx, res)
        pt.Render(scene, frames=2)
        img = pt.ReadAccumulation()
        assert np.isfinite(img).all() and img.mean() > 0
        pt.close(); scene.close()
        marks.append(free_mb())
    assert abs(marks[-1] - marks[0]) <= 8.0, marks            # MiB: nothing accumulates from cycle to cycle
    ctx.close()
    after = free_mb()
    assert before - after <= 64.0, (before, marks, after)     # what stays is the driver's own (module, primary context), not ours
