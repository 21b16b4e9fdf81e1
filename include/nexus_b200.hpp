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
#include <deque>
#include <memory>
#include <set>
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
    void* stream() const { return nx_ctx_stream(h_); }                    // the device stream (the runtime's stream handle) the render kernels are queued on
    // on (default): instances whose mesh no other instance uses, and that were never moved, share one world-space BLAS (same hits)
    void SetInstanceMerging(bool enabled) { check(nx_ctx_set_instance_merging(h_, enabled ? 1 : 0), "SetInstanceMerging"); }
    // Scene::Update after instances moved: rebuild the TLAS (default, as the reference does) or refit it in place when the entry set is unchanged
    void SetTlasRefit(bool enabled) { check(nx_ctx_set_tlas_refit(h_, enabled ? 1 : 0), "SetTlasRefit"); }
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
inline void FreeHostBVH(std::vector<nx_bvh2_node>& hostNodes) { std::vector<nx_bvh2_node>().swap(hostNodes); }   // NXB::FreeHostBVH (BVHBuilder.h:42)
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
    nx_light pod() const
    {
        nx_light p{}; p.type = (int32_t)type;
        p.position[0] = position.x; p.position[1] = position.y; p.position[2] = position.z;
        p.direction[0] = direction.x; p.direction[1] = direction.y; p.direction[2] = direction.z;
        p.color[0] = color.x; p.color[1] = color.y; p.color[2] = color.z; p.intensity = intensity; p.instance = meshId;
        return p;
    }
};

struct Camera {                         // src/Scene/Camera.h:9-52 (ctor arguments, setters, invalid flag)
    float3 position{0.0f, 4.0f, 14.0f}, forward{0.0f, 0.0f, -1.0f}; float horizontalFOV = 45.0f, focusDistance = 5.0f, defocusAngle = 0.0f;
    float3 right{0.0f, 0.0f, 0.0f};     // zero: derived as cross(forward, +Y)
    float& GetHorizontalFOV() { return horizontalFOV; }
    void SetHorizontalFOV(float v) { horizontalFOV = v; }
    float& GetDefocusAngle() { return defocusAngle; }
    void SetDefocusAngle(float v) { defocusAngle = v; }
    float& GetFocusDist() { return focusDistance; }
    void SetFocusDist(float v) { focusDistance = v; }
    float3& GetPosition() { return position; }
    void SetPosition(const float3& v) { position = v; }
    float3& GetForwardDirection() { return forward; }
    void SetForwardDirection(const float3& v) { forward = v; }
    float3& GetRightDirection() { return right; }
    void SetRightDirection(const float3& v) { right = v; }
    bool IsInvalid() const { return invalid_; }
    void SetInvalid(bool v) { invalid_ = v; }
    void Invalidate() { invalid_ = true; }
private:
    bool invalid_ = true;
};

struct RenderSettings {                 // src/Renderer/RenderSettings.h:5-17
    bool useMIS = true; int pathLength = 10; float3 backgroundColor{0, 0, 0}; float backgroundIntensity = 1.0f; int toneMapping = 3; float exposure = 0.0f;
};

class Scene;

// src/Scene/MeshInstance.h:9-77: host object; edits reach the device at the next Scene::Update() after
// Scene::InvalidateMeshInstance(index) - the setters below also invalidate by themselves.  (The reference's SetRotationY / SetRotationZ
// write `position`, MeshInstance.h:24-25; that bug is not reproduced.)
class MeshInstance {
public:
    MeshInstance(Scene* s, uint32_t idx, uint32_t mesh, int material, float3 p, float3 r, float3 sc)
        : name("instance " + std::to_string(idx)), position(p), rotation(r), scale(sc), meshIdx(mesh), materialIdx(material), scene_(s), idx_(idx) {}
    void SetPosition(float3 p) { position = p; Edited(); }
    void SetRotationX(float r) { rotation.x = r; Edited(); }
    void SetRotationY(float r) { rotation.y = r; Edited(); }
    void SetRotationZ(float r) { rotation.z = r; Edited(); }
    void SetScale(float s) { scale = float3{s, s, s}; Edited(); }
    void SetScale(float3 s) { scale = s; Edited(); }
    void SetTransform(float3 p, float3 r, float3 s);          // applied at once
    void AssignMaterial(int mIdx);
    std::array<float, 16> GetTransfromationMatrix();           // T * Rz * Ry * Rx * S, row-major (MeshInstance.h:36-40; the reference's spelling)
    NXB::AABB GetBounds();                                     // world AABB of the mesh box's eight corners (MeshInstance.h:42-53)
    uint32_t index() const { return idx_; }
    std::string name;
    float3 position, rotation, scale;
    uint32_t meshIdx; int materialIdx;
private:
    friend class Scene;
    void Edited();
    Scene* scene_; uint32_t idx_; bool trsEdited_ = false, materialEdited_ = false;
};

class AssetManager {                    // src/Assets/AssetManager.h:13-57
public:
    explicit AssetManager(Scene* s) : scene_(s) {}
    uint32_t AddMaterial(const Material& m);
    std::vector<Material>& GetMaterials() { return materials_; }        // edit an entry, then InvalidateMaterial(index)
    // AddMesh builds the BLAS immediately (Mesh::Mesh, src/Assets/Mesh.h:15-46: BuildBVH8<Triangle>, prioritizeSpeed = true)
    uint32_t AddMesh(const std::string& name, uint32_t materialIdx, const std::vector<NXB::Triangle>& triangles, const std::vector<nx_triangle_data>& triangleData = {});
    void InvalidateMaterial(uint32_t idx) { if (idx >= materials_.size()) throw Error("InvalidateMaterial: no such material"); invalid_.insert(idx); }
    void InvalidateMaterial(uint32_t idx, const Material& m) { if (idx >= materials_.size()) throw Error("InvalidateMaterial: no such material"); materials_[idx] = m; invalid_.insert(idx); SendDataToDevice(); }
    bool SendDataToDevice();                                            // AssetManager.cpp:74-84: uploads the invalidated materials
    bool IsInvalid() const { return !invalid_.empty(); }
    // AddTexture + Texture::ToDevice (src/Assets/Texture.cpp:12-46): RGBA8 (isHDR false) or RGBA32F pixels; returns the id Material::*MapId uses
    uint32_t AddTexture(const void* rgbaPixels, uint32_t width, uint32_t height, bool isHDR = false, bool sRGB = false);
    size_t GetTextureCount() const { return textureCount_; }
private:
    Scene* scene_; std::vector<Material> materials_; std::set<uint32_t> invalid_; size_t textureCount_ = 0;
};

class Scene {                           // src/Scene/Scene.h:16-77
public:
    Scene(Context& ctx, uint2 resolution) : ctx_(ctx), assets_(this), res_(resolution), camera_(std::make_shared<Camera>()) { ctx.check(nx_scene_create(ctx.handle(), resolution.x, resolution.y, &h_), "Scene"); }
    ~Scene() { nx_scene_destroy(h_); }
    // Camera::OnResize (src/Scene/Camera.cpp:118-128): only the camera record depends on the resolution; pair with PathTracer::OnResize
    void OnResize(uint2 resolution) { ctx_.check(nx_scene_set_resolution(h_, resolution.x, resolution.y), "Scene::OnResize"); res_ = resolution; }
    Scene(const Scene&) = delete; Scene& operator=(const Scene&) = delete;
    std::shared_ptr<Camera> GetCamera() { return camera_; }               // edit, Invalidate(), Update()
    AssetManager& GetAssetManager() { return assets_; }
    uint32_t AddMaterial(const Material& m) { return assets_.AddMaterial(m); }
    std::vector<Material>& GetMaterials() { return assets_.GetMaterials(); }
    RenderSettings& GetRenderSettings() { return settings_; }             // edits are uploaded by the next Update()
    const RenderSettings& GetRenderSettings() const { return settings_; }
    bool IsEmpty() const { return instances_.empty(); }
    bool IsInvalid() const { return !invalidInstances_.empty() || !invalidLights_.empty() || camera_->IsInvalid() || assets_.IsInvalid(); }
    MeshInstance& CreateMeshInstance(uint32_t meshId, float3 position = {0, 0, 0}, float3 rotationDeg = {0, 0, 0}, float3 scale = {1, 1, 1}, int materialIdx = -1)
    {
        const float p[3] = {position.x, position.y, position.z}, r[3] = {rotationDeg.x, rotationDeg.y, rotationDeg.z}, s[3] = {scale.x, scale.y, scale.z};
        const uint32_t idx = (uint32_t)ctx_.check(nx_scene_add_instance(h_, meshId, materialIdx, p, r, s), "CreateMeshInstance");
        instances_.emplace_back(this, idx, meshId, materialIdx, position, rotationDeg, scale);
        return instances_.back();
    }
    // an instance placed by an explicit row-major 4x4 (imported node transforms); position / rotation / scale stay at their defaults until set
    MeshInstance& CreateMeshInstanceMatrix(uint32_t meshId, const float matrix[16], int materialIdx = -1)
    {
        const uint32_t idx = (uint32_t)ctx_.check(nx_scene_add_instance_matrix(h_, meshId, materialIdx, matrix), "CreateMeshInstance");
        instances_.emplace_back(this, idx, meshId, materialIdx, float3{0, 0, 0}, float3{0, 0, 0}, float3{1, 1, 1});
        return instances_.back();
    }
    std::deque<MeshInstance>& GetMeshInstances() { return instances_; }   // a deque: references stay valid as instances are added
    void InvalidateMeshInstance(uint32_t instanceId) { invalidInstances_.insert(instanceId); }
    size_t AddLight(const Light& l)
    {
        const nx_light p = l.pod();
        const size_t idx = (size_t)ctx_.check(nx_scene_add_light(h_, &p), "AddLight");
        lights_.push_back(l);
        return idx;
    }
    std::vector<Light>& GetLights() { return lights_; }                   // edit an entry, then InvalidateLight(index)
    void InvalidateLight(uint32_t lightIdx) { if (lightIdx >= lights_.size()) throw Error("InvalidateLight: no such light"); invalidLights_.insert(lightIdx); }
    void RemoveLight(size_t index)
    {
        if (index >= lights_.size()) throw Error("RemoveLight: no such light");
        ctx_.check(nx_scene_remove_light(h_, (uint32_t)index), "RemoveLight");
        lights_.erase(lights_.begin() + (std::ptrdiff_t)index);
        std::set<uint32_t> moved;
        for (uint32_t i : invalidLights_) if (i != index) moved.insert(i > index ? i - 1 : i);
        invalidLights_.swap(moved);
    }
    void AddHDRMap(const float* rgba, uint32_t w, uint32_t h) { ctx_.check(nx_scene_set_hdr_map(h_, rgba, w, h), "AddHDRMap"); }
    void SetCamera(const Camera& c)
    {
        *camera_ = c;
        nx_camera p{}; p.position[0] = c.position.x; p.position[1] = c.position.y; p.position[2] = c.position.z;
        p.forward[0] = c.forward.x; p.forward[1] = c.forward.y; p.forward[2] = c.forward.z;
        p.right[0] = c.right.x; p.right[1] = c.right.y; p.right[2] = c.right.z;
        p.horizontal_fov_deg = c.horizontalFOV; p.focus_distance = c.focusDistance; p.defocus_angle_deg = c.defocusAngle;
        ctx_.check(nx_scene_set_camera(h_, &p), "SetCamera");
        camera_->SetInvalid(false);
    }
    void SetRenderSettings(const RenderSettings& r)
    {
        settings_ = r;
        nx_render_settings p{}; p.use_mis = r.useMIS; p.path_length = r.pathLength;
        p.background_color[0] = r.backgroundColor.x; p.background_color[1] = r.backgroundColor.y; p.background_color[2] = r.backgroundColor.z;
        p.background_intensity = r.backgroundIntensity; p.tone_mapping = r.toneMapping; p.exposure = r.exposure;
        ctx_.check(nx_scene_set_render_settings(h_, &p), "SetRenderSettings");
    }
    // Scene::Update (Scene.cpp:34-63): uploads what was invalidated (camera, settings, materials, instances, lights), rebuilds the TLAS
    // when an instance changed, refreshes the light list
    void Update()
    {
        if (camera_->IsInvalid()) SetCamera(Camera(*camera_));
        SetRenderSettings(RenderSettings(settings_));                     // a host-side struct copy; no device work
        assets_.SendDataToDevice();
        for (uint32_t i : std::set<uint32_t>(invalidInstances_)) if (i < instances_.size()) Flush(instances_[i]);
        invalidInstances_.clear();
        for (uint32_t i : invalidLights_) { const nx_light p = lights_[i].pod(); ctx_.check(nx_scene_set_light(h_, i, &p), "InvalidateLight"); }
        invalidLights_.clear();
        ctx_.check(nx_scene_update(h_), "Scene::Update");
    }
    void BuildTLAS() { Update(); }                                        // the library rebuilds the TLAS inside Update when an instance changed
    NXB::BVH8 GetTLAS() { NXB::BVH8 b; ctx_.check(nx_scene_tlas(h_, &b.h), "TLAS"); return b; }
    nx_scene* handle() const { return h_; }
    Context& context() const { return ctx_; }
    uint2 resolution() const { return res_; }
private:
    friend class MeshInstance;
    void Flush(MeshInstance& m)
    {
        if (m.trsEdited_) {
            const float p[3] = {m.position.x, m.position.y, m.position.z}, r[3] = {m.rotation.x, m.rotation.y, m.rotation.z}, s[3] = {m.scale.x, m.scale.y, m.scale.z};
            ctx_.check(nx_scene_set_instance_transform(h_, m.idx_, p, r, s), "InvalidateMeshInstance");
        }
        if (m.materialEdited_) ctx_.check(nx_scene_set_instance_material(h_, m.idx_, m.materialIdx), "AssignMaterial");
        m.trsEdited_ = m.materialEdited_ = false;
        invalidInstances_.erase(m.idx_);
    }
    Context& ctx_; nx_scene* h_ = nullptr; AssetManager assets_; uint2 res_;
    std::shared_ptr<Camera> camera_; RenderSettings settings_;
    std::deque<MeshInstance> instances_; std::set<uint32_t> invalidInstances_;
    std::vector<Light> lights_; std::set<uint32_t> invalidLights_;
};

inline void MeshInstance::Edited() { trsEdited_ = true; scene_->InvalidateMeshInstance(idx_); }
inline void MeshInstance::SetTransform(float3 p, float3 r, float3 s) { position = p; rotation = r; scale = s; Edited(); scene_->Flush(*this); }
inline void MeshInstance::AssignMaterial(int mIdx) { materialIdx = mIdx; materialEdited_ = true; scene_->InvalidateMeshInstance(idx_); }
inline std::array<float, 16> MeshInstance::GetTransfromationMatrix()
{
    scene_->Flush(*this);
    std::array<float, 16> m{}; scene_->context().check(nx_scene_instance_matrix(scene_->handle(), idx_, m.data()), "GetTransfromationMatrix"); return m;
}
inline NXB::AABB MeshInstance::GetBounds()
{
    scene_->Flush(*this);
    NXB::AABB b{}; scene_->context().check(nx_scene_instance_bounds(scene_->handle(), idx_, &b), "GetBounds"); return b;
}
inline uint32_t AssetManager::AddMaterial(const Material& m)
{
    const nx_material p = m.pod();
    const uint32_t idx = (uint32_t)scene_->context().check(nx_scene_add_material(scene_->handle(), &p), "AddMaterial");
    materials_.push_back(m);
    return idx;
}
inline bool AssetManager::SendDataToDevice()
{
    const bool any = !invalid_.empty();
    for (uint32_t i : invalid_) { const nx_material p = materials_[i].pod(); scene_->context().check(nx_scene_set_material(scene_->handle(), i, &p), "InvalidateMaterial"); }
    invalid_.clear();
    return any;
}
inline uint32_t AssetManager::AddTexture(const void* px, uint32_t w, uint32_t h, bool isHDR, bool sRGB)
{
    const uint32_t id = (uint32_t)scene_->context().check(nx_scene_add_texture(scene_->handle(), px, w, h, isHDR, sRGB), "AddTexture");
    textureCount_++;
    return id;
}
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
    // multi-GPU sample partition (include/nexus_b200_nccl.hpp): frames [firstFrame, firstFrame + frames) - frames are pure functions of
    // (pixel, frame index, bounce), so disjoint blocks on different GPUs add up to the image one GPU would accumulate
    void RenderFrames(Scene& scene, uint32_t firstFrame, uint32_t frames) { ctx_.check(nx_renderer_render(h_, scene.handle(), firstFrame, frames), "Render"); frame_ += frames; }
    float* AccumulationDevice() { float* d = nullptr; uint32_t f = 0; ctx_.check(nx_renderer_accum_device(h_, &d, &f), "AccumulationDevice"); return d; }   // 3 * W * H float SUMS
    void SetAccumulatedFrames(uint32_t frames) { ctx_.check(nx_renderer_set_accum_frames(h_, frames), "SetAccumulatedFrames"); frame_ = frames; }       // after an external reduce
    Context& context() { return ctx_; }
    void Reset() { ResetFrameNumber(); }                                  // PathTracer::Reset (PathTracer.cpp:61-159): the queues here are sized once per resolution
    void UpdateDeviceScene(Scene& scene) { scene.Update(); }              // PathTracer.cpp:216-219: the scene view is a per-call parameter block; flush pending edits
    uint2 GetResolution() const { return res_; }
    nx_frame_stats Stats() { nx_frame_stats s{}; ctx_.check(nx_renderer_stats(h_, &s), "Stats"); return s; }
    std::vector<float> ReadAccumulation() { std::vector<float> v(3ull * res_.x * res_.y); ctx_.check(nx_renderer_read_accum(h_, v.data()), "ReadAccumulation"); return v; }
    std::vector<uint32_t> ReadRGBA8(Scene& scene) { std::vector<uint32_t> v((size_t)res_.x * res_.y); ctx_.check(nx_renderer_read_rgba8(h_, scene.handle(), v.data()), "ReadRGBA8"); return v; }
    // Pipelined display read-back (the reference's PBO path): Present queues resolve + copy into a caller-owned (pinned) image
    // and returns a ticket; PresentWait blocks until that image is on the host and returns the render call's queue totals.
    int Present(Scene& scene, uint32_t* hostRgba) { int t = 0; ctx_.check(nx_renderer_present(h_, scene.handle(), hostRgba, &t), "Present"); return t; }
    // ... or straight into device memory: the mapped pixel buffer of an OpenGL viewer (PixelBuffer::GetDevicePtr); queued on Context::stream()
    void PresentDevice(Scene& scene, uint32_t* devRgba) { ctx_.check(nx_renderer_present_device(h_, scene.handle(), devRgba), "PresentDevice"); }
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
