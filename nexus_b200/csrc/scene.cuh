// Device-side scene mirrors shared by scene.cu (upload) and render.cu (kernels).
#pragma once
#include "nx_common.cuh"
#include "traverse.cuh"

struct DMesh {                 // replaces D_Mesh (src/Cuda/Scene/Mesh.cuh:5-9)
    const float* tris;         // 36-byte NXB::Triangle AoS, original order (shading + light sampling address by primitive id)
    const float* tridata;      // 96-byte D_TriangleData AoS
    uint32_t primCount;
    uint32_t pad;
};

struct DShadeInst {            // replaces the shading half of D_MeshInstance (src/Cuda/Scene/MeshInstance.cuh:9-16), 112 B
    float4 m0, m1, m2;         // object -> world rows
    float4 i0, i1, i2;         // world -> object rows
    uint32_t meshIdx, materialIdx, pad0, pad1;
};

struct DLight {                // flattened D_Light (src/Cuda/Scene/Light.cuh:4-47)
    int32_t type;
    float px, py, pz, dx, dy, dz, cr, cg, cb, intensity;
    uint32_t instance;
};

struct DCamera {               // D_Camera (src/Cuda/Scene/Camera.cuh:5-14)
    float position[3], right[3], up[3]; float lensRadius;
    float lowerLeft[3], viewportX[3], viewportY[3];
    uint32_t pad_;            // uint2 resolution is 8-byte aligned in the reference
    uint32_t resX, resY;
};
static_assert(sizeof(DCamera) == 88, "D_Camera layout");

struct DSceneView {            // kernel parameter block (replaces the __constant__ D_Scene / tlas / meshes symbols)
    TraceScene trace;
    const DShadeInst* shadeInst;
    const DMesh* meshes;
    const nx_material* materials;
    const DLight* lights;
    uint32_t lightCount;
    uint32_t hasHdr;
    cudaTextureObject_t hdr;
    const cudaTextureObject_t* textures;   // material maps (D_Scene::textures, src/Cuda/Scene/Scene.cuh:30), indexed by nx_material::*_map
    DCamera camera;
    uint32_t useMIS, pathLength;
    float bg[3], bgIntensity;
};

struct HostMesh {
    float* dTris = nullptr; float* dTriData = nullptr;
    float4* dLeafTris = nullptr;
    nx_bvh8 bvh{};
    uint32_t materialIdx = 0;
    double sphere[4] = {0, 0, 0, 0};   // object-space bounding sphere of the vertices (centre, radius)
};

struct HostInstance {
    uint32_t meshIdx = 0, materialIdx = 0;
    float m[16]; float inv[16];
    nx_aabb bounds{};
};

struct nx_scene {
    nx_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0;
    std::vector<nx_material> materials;
    std::vector<HostMesh> meshes;
    std::vector<HostInstance> instances;
    std::vector<nx_light> userLights;      // punctual lights added through nx_scene_add_light
    std::vector<DLight> lights;            // device list: punctual + emissive-instance lights
    nx_camera camera{};
    nx_render_settings settings{};
    bool dirtyInstances = true, dirtyMaterials = true, dirtyLights = true;

    // device mirrors
    nx_bvh8 tlas{};
    DTravInst* dTravInst = nullptr;        // inside dTop
    // Everything a ray touches before it enters an instance, in ONE allocation so that one L2 access-policy window covers it:
    // [copy of the TLAS nodes | traversal records in TLAS leaf order].  Marked persisting in L2 on both trace streams.
    void* dTop = nullptr; size_t topBytes = 0; const float4* dTopNodes = nullptr;
    DShadeInst* dShadeInst = nullptr;
    DMesh* dMeshes = nullptr;
    nx_material* dMaterials = nullptr;
    DLight* dLights = nullptr;
    cudaTextureObject_t hdr = 0; cudaArray_t hdrArray = nullptr; bool hasHdr = false;
    std::vector<cudaTextureObject_t> textures; std::vector<cudaArray_t> textureArrays;   // AssetManager::AddTexture
    cudaTextureObject_t* dTextures = nullptr; bool dirtyTextures = false;
    uint32_t dMeshCount = 0;
};

// scene.cu
int nxi_scene_view(nx_scene* s, DSceneView* out);
DCamera nxi_camera_to_device(const nx_camera& c, uint32_t w, uint32_t h);
// bvh_builder.cu
int nxi_build_bvh8(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, int prioritizeSpeed, nx_bvh8* out);
// render.cu
int nxi_trace_closest(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, nx_hit* dHits, float* outMs);
int nxi_trace_any(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, uint8_t* dOcc, float* outMs);
