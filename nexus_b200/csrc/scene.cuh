// Device-side scene mirrors shared by scene.cu (upload) and render.cu (kernels).
#pragma once
#include "nx_common.cuh"
#include "traverse.cuh"

struct DMesh {                 // replaces D_Mesh (src/Cuda/Scene/Mesh.cuh:5-9)
    const float* tris;         // 36-byte NXB::Triangle AoS, original order (shading + light sampling address by primitive id)
    const float* tridata;      // 96-byte D_TriangleData AoS (tangents and texture coordinates are read from here)
    const float4* shade;       // 80-byte shading record per primitive, 16-byte aligned: positions + vertex normals in five float4
    uint32_t primCount;
    uint32_t pad;
};

// Shading record: what every shaded hit (and every light sample) reads.  The reference fetches nine position floats at a 36-byte
// stride and nine normal floats at a 96-byte stride with 4-byte loads (PathTracer.cu:355-371); here it is
// five LDG.128: {v0, n0.x} {v1, n0.y} {v2, n0.z} {n1, n2.x} {n2.y, n2.z, -, -}.
#define NX_SHADE_REC_F4 5

struct __align__(16) DMaterial { nx_material m; uint32_t pad; };   // 96 B: six LDG.128 instead of 23 scalar loads
static_assert(sizeof(DMaterial) == 96, "DMaterial layout");

struct DShadeInst {            // replaces the shading half of D_MeshInstance (src/Cuda/Scene/MeshInstance.cuh:9-16), 112 B
    float4 m0, m1, m2;         // object -> world rows
    float4 i0, i1, i2;         // world -> object rows
    uint32_t meshIdx, materialIdx;
    const float4* shade;       // = meshes[meshIdx].shade: a shaded hit goes instance -> shading record without the mesh table in between,
                               // and shade_kernel's logic phase can prefetch the record of a surviving hit (shade.cu)
};
static_assert(sizeof(DShadeInst) == 112, "DShadeInst layout");

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
    const DMaterial* materials;
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
    float4* dShadeRec = nullptr;
    nx_bvh8 bvh{};
    uint32_t materialIdx = 0;
    double sphere[4] = {0, 0, 0, 0};   // object-space bounding sphere of the vertices (centre, radius)
    bool blasBuilt = false;            // the mesh has a BLAS of its own (built on demand: meshes that only live in the merged BLAS never get one)
    uint32_t useCount = 0;             // instances referring to the mesh
    bool arenaOwned = false;           // all device arrays live in the scene's arena: nothing is freed per mesh
    bool prebuilt = false;             // BLAS supplied by the caller (nx_scene_add_mesh_prebuilt): indices validated on the device
    bool pending = false;              // BLAS build in flight on a build stream: node_count not known yet (flush_builds)
};

struct HostInstance {
    uint32_t meshIdx = 0, materialIdx = 0;
    float m[16]; float inv[16];
    nx_aabb bounds{};
    bool dynamic = false;              // moved after creation (nx_scene_set_instance_transform): never merged
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
    DMaterial* dMaterials = nullptr;
    DLight* dLights = nullptr;
    cudaTextureObject_t hdr = 0; cudaArray_t hdrArray = nullptr; bool hasHdr = false;
    std::vector<cudaTextureObject_t> textures; std::vector<cudaArray_t> textureArrays;   // AssetManager::AddTexture
    cudaTextureObject_t* dTextures = nullptr; bool dirtyTextures = false;
    uint32_t dMeshCount = 0;
    // scene set-up pipeline: device counters of the BLAS builds in flight (8 words per mesh, chunks of 1024 meshes)
    std::vector<uint32_t*> buildCounterChunks;
    std::vector<nx_bump> arena;            // slabs holding every mesh's geometry, shading records and BLAS (freed with the scene)
    uint32_t pendingBuilds = 0;
    // merged BLAS: one world-space BLAS over all instances whose mesh is used exactly once and that have not been moved
    std::vector<uint32_t> mergedInstances;   // instance ids in it, ascending
    std::vector<uint32_t> mergedFirst;       // first merged primitive of each of them
    nx_bvh8 merged{};
    float4* dMergedLeaf = nullptr;
    float4* dInstInv = nullptr;              // instance id -> rows of its world -> object 3x4
    uint32_t mergedSlot = 0xffffffffu;       // its TLAS leaf slot
    std::vector<uint32_t> tlasEntryInst;     // TLAS primitive (entry) -> instance id, 0xffffffff for the merged BLAS
    uint32_t tlasRefits = 0, tlasBuilds = 0; // how the TLAS came to its current state (parity hook: nx_scene_tlas_history)
    uint32_t* dSlotInst = nullptr;           // TLAS leaf slot -> instance id
};

// scene.cu
int nxi_scene_view(nx_scene* s, DSceneView* out);
DCamera nxi_camera_to_device(const nx_camera& c, uint32_t w, uint32_t h);
// bvh_builder.cu
int nxi_build_bvh8(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, int prioritizeSpeed, nx_bvh8* out, uint32_t maxLeafPrims = 0);
int nxi_refit_bvh8(nx_ctx* ctx, nx_bvh8* bvh, const void* dBounds);
int nxi_build_bvh8_async(nx_ctx* ctx, cudaStream_t stream, const void* dPrims, uint32_t n, int primType, int prioritizeSpeed, uint32_t* dCounters, nx_bvh8* out,
                         nx_bump* ws, nx_bump* outputs);
size_t nxi_build_workspace_bytes(uint32_t n);
// render.cu
int nxi_trace_closest(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, nx_hit* dHits, float* outMs);
int nxi_trace_any(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, uint8_t* dOcc, float* outMs);
