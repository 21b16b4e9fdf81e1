// nexus_b200.hpp — C++17 host classes over the C ABI (include/nexus_b200.h), keeping the names, argument meaning and
// call order of the reference's host API so a Nexus application switches by changing includes:
//
//   NXB::BuildBVH2 / BuildBVH8 / ToHost / FreeDeviceBVH / BenchmarkBuild   vendor/NexusBVH/NexusBVH/include/NXB/BVHBuilder.h:19-55
//   Material, Light, Camera, RenderSettings                                src/Assets/Material.h:6-26, src/Scene/Light.h:10-54,
//                                                                          src/Scene/Camera.h:9-52, src/Renderer/RenderSettings.h:5-17
//   AssetManager::AddMaterial / AddMesh                                    src/Assets/AssetManager.h:18-44
//   Scene::CreateMeshInstance / AddLight / AddHDRMap / Update              src/Scene/Scene.h:19-49
//   MeshInstance::SetTransform / AssignMaterial                            src/Scene/MeshInstance.h:22-53
//   PathTracer::Render / ResetFrameNumber / OnResize / GetFrameNumber      src/Renderer/PathTracer.h:12-29
//
// Differences from the reference, all deliberate: errors throw nexus::Error instead of exit(99) (src/Utils/Utils.cpp:3-12);
// every object belongs to a Context (one per GPU) instead of process-global __constant__ state, so one process can drive
// eight GPUs; the render target is an owned float accumulation buffer (headless) instead of an OpenGL PBO.
// Header-only; link with -lnexus_b200.  No CUDA headers needed by the host application.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "nexus_b200.h"

namespace nexus {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

struct float3 { float x, y, z; };
struct uint2 { uint32_t x, y; };

class Context {
public:
    explicit Context(int device = 0) { if (nx_ctx_create(device, &h_) != NX_OK) throw Error("nx_ctx_create failed: no usable CUDA device (there is no CPU fallback)"); }
    ~Context() { nx_ctx_destroy(h_); }
    Context(const Context&) = delete; Context& operator=(const Context&) = delete;
    nx_ctx* handle() const { return h_; }
    void Synchronize() { check(nx_ctx_synchronize(h_), "synchronize"); }
    int check(int rc, const char* what) const { if (rc < 0) throw Error(std::string(what) + ": " + nx_last_error(h_)); return rc; }
private:
    nx_ctx* h_ = nullptr;
};

}  // namespace nexus

// ------------------------------------------------------------------------------------------------ NXB builder API ----
namespace NXB {

using Triangle = nx_triangle;           // 3 x float3, 36 B (NXB/Triangle.h:10-47)
using AABB = nx_aabb;                   // 24 B (NXB/AABB.h:8-53)
struct BuildConfig { bool prioritizeSpeed = false; };            // NXB/BuildConfig.h:6-12
using BVHBuildMetrics = nx_build_metrics;                        // NXB/BVHBuildMetrics.h:7-61
struct BVH2 { nx_bvh2 h{}; };
struct BVH8 { nx_bvh8 h{}; uint32_t nodeCount() const { return h.node_count; } uint32_t primCount() const { return h.prim_count; } };

// primitives: DEVICE pointer, caller-owned, not modified (as the reference).  Blocking.
inline BVH2 BuildBVH2(nexus::Context& ctx, const Triangle* dPrims, uint32_t n, BuildConfig cfg = {}, BVHBuildMetrics* m = nullptr)
{ BVH2 b; nx_build_config c{cfg.prioritizeSpeed}; ctx.check(nx_bvh2_build_tri(ctx.handle(), dPrims, n, &c, m, &b.h), "BuildBVH2<Triangle>"); return b; }
inline BVH2 BuildBVH2(nexus::Context& ctx, const AABB* dPrims, uint32_t n, BuildConfig cfg = {}, BVHBuildMetrics* m = nullptr)
{ BVH2 b; nx_build_config c{cfg.prioritizeSpeed}; ctx.check(nx_bvh2_build_aabb(ctx.handle(), dPrims, n, &c, m, &b.h), "BuildBVH2<AABB>"); return b; }
inline BVH8 BuildBVH8(nexus::Context& ctx, const Triangle* dPrims, uint32_t n, BuildConfig cfg = {}, BVHBuildMetrics* m = nullptr)
{ BVH8 b; nx_build_config c{cfg.prioritizeSpeed}; ctx.check(nx_bvh8_build_tri(ctx.handle(), dPrims, n, &c, m, &b.h), "BuildBVH8<Triangle>"); return b; }
inline BVH8 BuildBVH8(nexus::Context& ctx, const AABB* dPrims, uint32_t n, BuildConfig cfg = {}, BVHBuildMetrics* m = nullptr)
{ BVH8 b; nx_build_config c{cfg.prioritizeSpeed}; ctx.check(nx_bvh8_build_aabb(ctx.handle(), dPrims, n, &c, m, &b.h), "BuildBVH8<AABB>"); return b; }
inline std::vector<nx_bvh2_node> ToHost(nexus::Context& ctx, const BVH2& b)
{ std::vector<nx_bvh2_node> v(b.h.node_count); ctx.check(nx_bvh2_to_host(ctx.handle(), &b.h, v.data()), "ToHost"); return v; }
inline void FreeDeviceBVH(nexus::Context& ctx, BVH2& b) { nx_bvh2_free(ctx.handle(), &b.h); }
inline void FreeDeviceBVH(nexus::Context& ctx, BVH8& b) { nx_bvh8_free(ctx.handle(), &b.h); }
// BenchmarkBuild(BuildBVH8<PrimT>, warmup, iterations, ...) (NXB/BVHBuildMetrics.h:63-108): averaged per-stage times
template <typename PrimT>
inline BVHBuildMetrics BenchmarkBuild(nexus::Context& ctx, uint32_t warmup, uint32_t iterations, const PrimT* dPrims, uint32_t n, BuildConfig cfg = {})
{
    BVHBuildMetrics m{}; nx_build_config c{cfg.prioritizeSpeed}; uint32_t nodes = 0;
    ctx.check(nx_bvh8_benchmark(ctx.handle(), dPrims, n, sizeof(PrimT) == sizeof(Triangle) ? 1 : 0, &c, (int)warmup, (int)iterations, &m, &nodes), "BenchmarkBuild");
    return m;
}

}  // namespace NXB

// ------------------------------------------------------------------------------------------ scene / renderer API ----
namespace nexus {

struct Material {                       // src/Assets/Material.h:6-26, same defaults
    float3 baseColor{0.8f, 0.8f, 0.8f}; float metalness = 0.0f, roughness = 0.3f, anisotropy = 0.0f, specularWeight = 1.0f;
    float3 specularColor{1.0f, 1.0f, 1.0f}; float ior = 1.5f, transmission = 0.0f;
    float3 emissionColor{1.0f, 1.0f, 1.0f}; float intensity = 0.0f, opacity = 1.0f;
    int32_t baseColorMapId = -1, emissiveMapId = -1, normalMapId = -1, roughnessMapId = -1, metalnessMapId = -1, metallicRoughnessMapId = -1;
    nx_material pod() const
    {
        nx_material m{};
        m.base_color[0] = baseColor.x; m.base_color[1] = baseColor.y; m.base_color[2] = baseColor.z;
        m.metalness = metalness; m.roughness = roughness; m.anisotropy = anisotropy; m.specular_weight = specularWeight;
        m.specular_color[0] = specularColor.x; m.specular_color[1] = specularColor.y; m.specular_color[2] = specularColor.z;
        m.ior = ior; m.transmission = transmission;
        m.emission_color[0] = emissionColor.x; m.emission_color[1] = emissionColor.y; m.emission_color[2] = emissionColor.z;
        m.intensity = intensity; m.opacity = opacity;
        m.base_color_map = baseColorMapId; m.emissive_map = emissiveMapId; m.normal_map = normalMapId;
        m.roughness_map = roughnessMapId; m.metalness_map = metalnessMapId; m.metallic_roughness_map = metallicRoughnessMapId;
        return m;
    }
};

struct Light {                          // src/Scene/Light.h:10-54
    enum class Type { POINT = 0, SPOT = 1, DIRECTIONAL = 2, MESH = 3 } type = Type::POINT;
    float3 position{0, 0, 0}, direction{0, -1, 0}, color{1, 1, 1}; float intensity = 1.0f; uint32_t meshId = 0;
};

struct Camera {                         // src/Scene/Camera.h:9-52 (ctor arguments)
    float3 position{0.0f, 4.0f, 14.0f}, forward{0.0f, 0.0f, -1.0f}; float horizontalFOV = 45.0f, focusDistance = 5.0f, defocusAngle = 0.0f;
};

struct RenderSettings {                 // src/Renderer/RenderSettings.h:5-17
    bool useMIS = true; int pathLength = 10; float3 backgroundColor{0, 0, 0}; float backgroundIntensity = 1.0f; int toneMapping = 3; float exposure = 0.0f;
};

class Scene;

class MeshInstance {                    // src/Scene/MeshInstance.h:9-77
public:
    MeshInstance(Scene* s, uint32_t idx) : scene_(s), idx_(idx) {}
    void SetTransform(float3 position, float3 rotationDeg, float3 scale);
    uint32_t index() const { return idx_; }
private:
    Scene* scene_; uint32_t idx_;
};

class AssetManager {                    // src/Assets/AssetManager.h:13-57
public:
    explicit AssetManager(Scene* s) : scene_(s) {}
    uint32_t AddMaterial(const Material& m);
    // AddMesh builds the BLAS immediately (Mesh::Mesh, src/Assets/Mesh.h:15-46: BuildBVH8<Triangle>, prioritizeSpeed = true)
    uint32_t AddMesh(const std::string& name, uint32_t materialIdx, const std::vector<NXB::Triangle>& triangles, const std::vector<nx_triangle_data>& triangleData = {});
    void InvalidateMaterial(uint32_t idx, const Material& m);
    // AddTexture + Texture::ToDevice (src/Assets/Texture.cpp:12-46): RGBA8 (isHDR false) or RGBA32F pixels; returns the id Material::*MapId uses
    uint32_t AddTexture(const void* rgbaPixels, uint32_t width, uint32_t height, bool isHDR = false, bool sRGB = false);
private:
    Scene* scene_;
};

class Scene {                           // src/Scene/Scene.h:16-77
public:
    Scene(Context& ctx, uint2 resolution) : ctx_(ctx), assets_(this), res_(resolution) { ctx.check(nx_scene_create(ctx.handle(), resolution.x, resolution.y, &h_), "Scene"); }
    ~Scene() { nx_scene_destroy(h_); }
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    AssetManager& GetAssetManager() { return assets_; }
    uint32_t AddMaterial(const Material& m) { return assets_.AddMaterial(m); }
    MeshInstance CreateMeshInstance(uint32_t meshId, float3 position = {0, 0, 0}, float3 rotationDeg = {0, 0, 0}, float3 scale = {1, 1, 1}, int materialIdx = -1)
    {
        const float p[3] = {position.x, position.y, position.z}, r[3] = {rotationDeg.x, rotationDeg.y, rotationDeg.z}, s[3] = {scale.x, scale.y, scale.z};
        return MeshInstance(this, (uint32_t)ctx_.check(nx_scene_add_instance(h_, meshId, materialIdx, p, r, s), "CreateMeshInstance"));
    }
    void AddLight(const Light& l)
    {
        nx_light p{}; p.type = (int32_t)l.type;
        p.position[0] = l.position.x; p.position[1] = l.position.y; p.position[2] = l.position.z;
        p.direction[0] = l.direction.x; p.direction[1] = l.direction.y; p.direction[2] = l.direction.z;
        p.color[0] = l.color.x; p.color[1] = l.color.y; p.color[2] = l.color.z; p.intensity = l.intensity; p.instance = l.meshId;
        ctx_.check(nx_scene_add_light(h_, &p), "AddLight");
    }
    void AddHDRMap(const float* rgba, uint32_t w, uint32_t h) { ctx_.check(nx_scene_set_hdr_map(h_, rgba, w, h), "AddHDRMap"); }
    void SetCamera(const Camera& c)
    {
        nx_camera p{}; p.position[0] = c.position.x; p.position[1] = c.position.y; p.position[2] = c.position.z;
        p.forward[0] = c.forward.x; p.forward[1] = c.forward.y; p.forward[2] = c.forward.z;
        p.horizontal_fov_deg = c.horizontalFOV; p.focus_distance = c.focusDistance; p.defocus_angle_deg = c.defocusAngle;
        ctx_.check(nx_scene_set_camera(h_, &p), "SetCamera");
    }
    void SetRenderSettings(const RenderSettings& r)
    {
        nx_render_settings p{}; p.use_mis = r.useMIS; p.path_length = r.pathLength;
        p.background_color[0] = r.backgroundColor.x; p.background_color[1] = r.backgroundColor.y; p.background_color[2] = r.backgroundColor.z;
        p.background_intensity = r.backgroundIntensity; p.tone_mapping = r.toneMapping; p.exposure = r.exposure;
        ctx_.check(nx_scene_set_render_settings(h_, &p), "SetRenderSettings");
    }
    void Update() { ctx_.check(nx_scene_update(h_), "Scene::Update"); }      // uploads dirty state, rebuilds the TLAS, refreshes the light list
    void BuildTLAS() { Update(); }
    NXB::BVH8 GetTLAS() { NXB::BVH8 b; ctx_.check(nx_scene_tlas(h_, &b.h), "TLAS"); return b; }
    nx_scene* handle() const { return h_; }
    Context& context() const { return ctx_; }
    uint2 resolution() const { return res_; }
private:
    Context& ctx_; nx_scene* h_ = nullptr; AssetManager assets_; uint2 res_;
};

inline void MeshInstance::SetTransform(float3 position, float3 rotationDeg, float3 scale)
{
    const float p[3] = {position.x, position.y, position.z}, r[3] = {rotationDeg.x, rotationDeg.y, rotationDeg.z}, s[3] = {scale.x, scale.y, scale.z};
    scene_->context().check(nx_scene_set_instance_transform(scene_->handle(), idx_, p, r, s), "SetTransform");
}
inline uint32_t AssetManager::AddMaterial(const Material& m) { const nx_material p = m.pod(); return (uint32_t)scene_->context().check(nx_scene_add_material(scene_->handle(), &p), "AddMaterial"); }
inline uint32_t AssetManager::AddTexture(const void* px, uint32_t w, uint32_t h, bool isHDR, bool sRGB) { return (uint32_t)scene_->context().check(nx_scene_add_texture(scene_->handle(), px, w, h, isHDR, sRGB), "AddTexture"); }
inline void AssetManager::InvalidateMaterial(uint32_t idx, const Material& m) { const nx_material p = m.pod(); scene_->context().check(nx_scene_set_material(scene_->handle(), idx, &p), "InvalidateMaterial"); }
inline uint32_t AssetManager::AddMesh(const std::string& name, uint32_t materialIdx, const std::vector<NXB::Triangle>& tris, const std::vector<nx_triangle_data>& data)
{
    if (!data.empty() && data.size() != tris.size()) throw Error("AddMesh(" + name + "): triangleData must have one entry per triangle");
    return (uint32_t)scene_->context().check(nx_scene_add_mesh(scene_->handle(), tris.data(), data.empty() ? nullptr : data.data(), (uint32_t)tris.size(), materialIdx), "AddMesh");
}

class PathTracer {                      // src/Renderer/PathTracer.h:9-66
public:
    PathTracer(Context& ctx, uint2 resolution) : ctx_(ctx), res_(resolution) { ctx.check(nx_renderer_create(ctx.handle(), resolution.x, resolution.y, &h_), "PathTracer"); }
    ~PathTracer() { nx_renderer_destroy(h_); }
    PathTracer(const PathTracer&) = delete; PathTracer& operator=(const PathTracer&) = delete;
    void ResetFrameNumber() { ctx_.check(nx_renderer_reset_accumulation(h_), "ResetFrameNumber"); frame_ = 0; }
    void OnResize(uint2 resolution) { ctx_.check(nx_renderer_resize(h_, resolution.x, resolution.y), "OnResize"); res_ = resolution; frame_ = 0; }
    // One call = one frame = one sample per pixel, like PathTracer::Render (src/Renderer/PathTracer.cpp:166-200); asynchronous.
    void Render(Scene& scene) { Render(scene, 1); }
    void Render(Scene& scene, uint32_t frames) { ctx_.check(nx_renderer_render(h_, scene.handle(), frame_ + 1, frames), "Render"); frame_ += frames; }
    uint32_t GetFrameNumber() const { return frame_; }
    uint2 GetResolution() const { return res_; }
    nx_frame_stats Stats() { nx_frame_stats s{}; ctx_.check(nx_renderer_stats(h_, &s), "Stats"); return s; }
    std::vector<float> ReadAccumulation() { std::vector<float> v(3ull * res_.x * res_.y); ctx_.check(nx_renderer_read_accum(h_, v.data()), "ReadAccumulation"); return v; }
    std::vector<uint32_t> ReadRGBA8(Scene& scene) { std::vector<uint32_t> v((size_t)res_.x * res_.y); ctx_.check(nx_renderer_read_rgba8(h_, scene.handle(), v.data()), "ReadRGBA8"); return v; }
    // Pipelined display read-back (the reference's PBO path): Present queues resolve + copy into a caller-owned (pinned) image
    // and returns a ticket; PresentWait blocks until that image is on the host and returns the render call's queue totals.
    int Present(Scene& scene, uint32_t* hostRgba) { int t = 0; ctx_.check(nx_renderer_present(h_, scene.handle(), hostRgba, &t), "Present"); return t; }
    nx_frame_stats PresentWait(int ticket) { nx_frame_stats s{}; ctx_.check(nx_renderer_present_wait(h_, ticket, &s), "PresentWait"); return s; }
    // src/Renderer/PathTracer.h:23-27
    void SetPixelQuery(uint32_t x, uint32_t y) { ctx_.check(nx_renderer_set_pixel_query(h_, x, y), "SetPixelQuery"); }
    bool PixelQueryPending() const { return nx_renderer_pixel_query_pending(h_) == 1; }
    int32_t SynchronizePixelQuery() { ctx_.check(nx_renderer_sync_pixel_query(h_, &selected_), "SynchronizePixelQuery"); return selected_; }
    int32_t GetSelectedInstance() const { return selected_; }
    nx_renderer* handle() const { return h_; }
private:
    Context& ctx_; nx_renderer* h_ = nullptr; uint2 res_; uint32_t frame_ = 0; int32_t selected_ = -1;
};

inline void WritePFM(const std::string& path, const std::vector<float>& rgb, uint2 res) { if (nx_write_pfm(path.c_str(), rgb.data(), res.x, res.y) != NX_OK) throw Error("cannot write " + path); }
inline void WriteEXR(const std::string& path, const std::vector<float>& rgb, uint2 res) { if (nx_write_exr(path.c_str(), rgb.data(), res.x, res.y) != NX_OK) throw Error("cannot write " + path); }

}  // namespace nexus
