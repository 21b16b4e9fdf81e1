// (types are spelled nexus::float3 / nexus::uint2: a translation unit that also includes the CUDA runtime headers has global ones too)
// The Cornell box of the examples, built through the reference-shaped host API (include/nexus_b200.hpp): shared by
// render_headless.cpp and render_multigpu.cpp.
#pragma once
#include <vector>
#include "nexus_b200.hpp"

namespace cornell {
using namespace nexus;

inline void quad(std::vector<NXB::Triangle>& t, nexus::float3 a, nexus::float3 b, nexus::float3 c, nexus::float3 d)
{
    t.push_back(NXB::Triangle{{a.x, a.y, a.z}, {b.x, b.y, b.z}, {c.x, c.y, c.z}});
    t.push_back(NXB::Triangle{{a.x, a.y, a.z}, {c.x, c.y, c.z}, {d.x, d.y, d.z}});
}
inline std::vector<NXB::Triangle> box(const nexus::float3 b[4], float height)
{
    std::vector<NXB::Triangle> t;
    nexus::float3 u[4]; for (int i = 0; i < 4; i++) u[i] = nexus::float3{b[i].x, b[i].y + height, b[i].z};
    quad(t, u[0], u[1], u[2], u[3]);
    for (int i = 0; i < 4; i++) { int j = (i + 1) % 4; quad(t, b[i], b[j], u[j], u[i]); }
    return t;
}


inline void Build(Scene& scene)
{
    AssetManager& am = scene.GetAssetManager();
    auto diffuse = [](nexus::float3 c) { Material m; m.baseColor = c; m.roughness = 0.9f; m.ior = 1.0f; m.specularWeight = 0.0f; return m; };
    const nexus::float3 white{0.725f, 0.71f, 0.68f}, green{0.14f, 0.45f, 0.091f}, red{0.63f, 0.065f, 0.05f};
    const uint32_t mWhite = am.AddMaterial(diffuse(white)), mGreen = am.AddMaterial(diffuse(green)), mRed = am.AddMaterial(diffuse(red));
    Material light = diffuse(nexus::float3{0.78f, 0.78f, 0.78f}); light.intensity = 35.0f;
    const uint32_t mLight = am.AddMaterial(light);
    const float x0 = -1.0f, x1 = 1.0f, y0 = 0.0f, y1 = 1.99f, z0 = -1.04f, z1 = 0.99f;
    std::vector<NXB::Triangle> t;
    auto add = [&](const char* name, uint32_t mat) { scene.CreateMeshInstance(am.AddMesh(name, mat, t)); t.clear(); };
    quad(t, {x0, y0, z1}, {x1, y0, z1}, {x1, y0, z0}, {x0, y0, z0}); add("floor", mWhite);
    quad(t, {x0, y1, z0}, {x1, y1, z0}, {x1, y1, z1}, {x0, y1, z1}); add("ceiling", mWhite);
    quad(t, {x0, y0, z0}, {x1, y0, z0}, {x1, y1, z0}, {x0, y1, z0}); add("back", mWhite);
    quad(t, {x1, y0, z0}, {x1, y0, z1}, {x1, y1, z1}, {x1, y1, z0}); add("right", mGreen);
    quad(t, {x0, y0, z1}, {x0, y0, z0}, {x0, y1, z0}, {x0, y1, z1}); add("left", mRed);
    const nexus::float3 sb[4] = {{0.53f, 0, 0.75f}, {0.70f, 0, 0.17f}, {0.13f, 0, 0.0f}, {-0.05f, 0, 0.57f}};
    const nexus::float3 tb[4] = {{-0.53f, 0, 0.09f}, {0.04f, 0, -0.09f}, {-0.14f, 0, -0.67f}, {-0.71f, 0, -0.49f}};
    t = box(sb, 0.6f); add("shortBox", mWhite);
    t = box(tb, 1.2f); add("tallBox", mWhite);
    quad(t, {-0.24f, 1.98f, -0.22f}, {0.23f, 1.98f, -0.22f}, {0.23f, 1.98f, 0.16f}, {-0.24f, 1.98f, 0.16f}); add("light", mLight);
    Camera cam; cam.position = {0.0f, 0.995f, 3.9f}; cam.forward = {0.0f, 0.0f, -1.0f};
    scene.SetCamera(cam);
    scene.SetRenderSettings(RenderSettings{});
    scene.Update();
}
}  // namespace cornell
