// TEST INFRASTRUCTURE — headless driver for the UNMODIFIED reference kernels.
//
// This file is ours; everything it drives is compiled straight from
// /root/reference (see oracle/Makefile, target `_ref`) and is never copied into
// this repository.  It gives the tests and bench.py's `--impl reference` arm a
// C ABI onto:
//   * NXB::BuildBVH2 / NXB::BuildBVH8            (B/src/BVHBuilder.cpp:115,174)
//   * TraceKernel on a caller-supplied ray batch (N/Cuda/PathTracer/PathTracer.cu:98)
//   * the full frame loop of PathTracer::Render  (N/Renderer/PathTracer.cpp:166-200)
// The launch sequence below replays PathTracer::Reset/Render because the
// reference's own host class needs an OpenGL PBO (N/Renderer/PathTracer.cpp:7)
// and cannot be constructed headless.
//
// Only tests/, __graft_entry__.smoke() and bench.py may load the resulting
// oracle/_ref/libnexus_ref.so; the product never does.
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#include "Cuda/PathTracer/PathTracer.cuh"
#include "Cuda/Scene/Scene.cuh"
#include "Cuda/Scene/Mesh.cuh"
#include "Cuda/Scene/MeshInstance.cuh"
#include "NXB/BVHBuilder.h"
#include "NXB/BVHBuildMetrics.h"
#include "Cuda/Setup.h"          // B/src/Cuda/Setup.h: the reference's bounds / Morton kernels (explicitly instantiated in Setup.cu:114-118)

#define REF_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    std::fprintf(stderr, "[nxref] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)

namespace {

struct RefMesh {
    NXB::BVH8 bvh{};
    NXB::Triangle* tris = nullptr;
    D_TriangleData* tridata = nullptr;
};

struct RefState {
    std::vector<RefMesh> meshes;
    D_Mesh* dMeshes = nullptr;
    D_MeshInstance* dInstances = nullptr;
    uint32_t instanceCount = 0;
    D_Material* dMaterials = nullptr;
    D_Light* dLights = nullptr;
    uint32_t lightCount = 0;
    NXB::BVH8 tlas{};
    bool hasTlas = false;
    cudaTextureObject_t hdr = 0;
    cudaArray_t hdrArray = nullptr;
    bool hasHdr = false;
    std::vector<cudaTextureObject_t> textures; std::vector<cudaArray_t> textureArrays; cudaTextureObject_t* dTextures = nullptr;
    D_Camera camera{};
    D_RenderSettings settings{};

    // render buffers (PathTracer::Reset, N/Renderer/PathTracer.cpp:61-159)
    uint32_t w = 0, h = 0;
    std::vector<void*> allocs;
    float3* accum = nullptr;
    uint32_t* renderBuffer = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graphExec = nullptr;
    cudaStream_t graphStream = nullptr;
    dim3 pixelGrid, traceGrid, shadowGrid;
    unsigned long long* rayTotals = nullptr;  // [0] extension rays, [1] shadow rays
} g;

__global__ void AddQueueTotals(const D_QueueSize* q, unsigned long long* totals, int pathLength)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long ext = 0, sh = 0;
    for (int b = 0; b <= pathLength && b < PATH_MAX_LENGTH; b++) { ext += q->traceSize[b]; sh += q->traceShadowSize[b]; }
    totals[0] += ext;
    totals[1] += sh;
}

template <typename T> T* devAlloc(size_t n)
{
    T* p = nullptr;
    if (cudaMalloc((void**)&p, sizeof(T) * (n ? n : 1)) != cudaSuccess) return nullptr;
    g.allocs.push_back(p);
    return p;
}

void copyMetrics(const NXB::BVHBuildMetrics& m, float* out)
{
    if (!out) return;
    out[0] = m.computeSceneBoundsTime; out[1] = m.computeMortonCodesTime; out[2] = m.radixSortTime;
    out[3] = m.bvhBuildTime; out[4] = m.bvh8ConversionTime; out[5] = m.totalTime;
    out[6] = m.bvh2Cost; out[7] = m.bvh8Cost; out[8] = m.averageChildPerNode;
}

int occupancyGrid(const void* fn, int blockSize)
{
    // CUDAKernel::SetMinimalLaunchConfigurationWithBlockSize, N/Device/Kernels/CUDAKernel.h:31-45
    int dev = 0, perSm = 0;
    cudaDeviceProp prop;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, blockSize, 0);
    return perSm * prop.multiProcessorCount;
}

} // namespace

extern "C" {

// ---------------------------------------------------------------- builder ----
// primType: 0 = NXB::AABB (24 B), 1 = NXB::Triangle (36 B). Host pointers in, host pointers out.
int nxref_build_bvh2(const void* hostPrims, uint32_t n, int primType, int prioritizeSpeed,
                     void* outNodes /* (2n-1)*32 B */, float* outBounds6, float* outMetrics9)
{
    size_t stride = primType ? sizeof(NXB::Triangle) : sizeof(NXB::AABB);
    void* dPrims = nullptr;
    REF_CHECK(cudaMalloc(&dPrims, stride * n));
    REF_CHECK(cudaMemcpy(dPrims, hostPrims, stride * n, cudaMemcpyHostToDevice));
    NXB::BuildConfig cfg; cfg.prioritizeSpeed = prioritizeSpeed != 0;
    NXB::BVHBuildMetrics metrics{};
    NXB::BVH2 bvh = primType ? NXB::BuildBVH2<NXB::Triangle>((NXB::Triangle*)dPrims, n, cfg, outMetrics9 ? &metrics : nullptr)
                             : NXB::BuildBVH2<NXB::AABB>((NXB::AABB*)dPrims, n, cfg, outMetrics9 ? &metrics : nullptr);
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaMemcpy(outNodes, bvh.nodes, sizeof(NXB::BVH2::Node) * bvh.nodeCount, cudaMemcpyDeviceToHost));
    if (outBounds6) std::memcpy(outBounds6, &bvh.bounds, 24);
    copyMetrics(metrics, outMetrics9);
    NXB::FreeDeviceBVH(bvh);
    cudaFree(dPrims);
    return 0;
}

int nxref_build_bvh8(const void* hostPrims, uint32_t n, int primType, int prioritizeSpeed,
                     void* outNodes /* cap ceil((4n-1)/7)*80 B */, uint32_t* outPrimIdx, uint32_t* outNodeCount,
                     float* outBounds6, float* outMetrics9)
{
    size_t stride = primType ? sizeof(NXB::Triangle) : sizeof(NXB::AABB);
    void* dPrims = nullptr;
    REF_CHECK(cudaMalloc(&dPrims, stride * n));
    REF_CHECK(cudaMemcpy(dPrims, hostPrims, stride * n, cudaMemcpyHostToDevice));
    NXB::BuildConfig cfg; cfg.prioritizeSpeed = prioritizeSpeed != 0;
    NXB::BVHBuildMetrics metrics{};
    NXB::BVH8 bvh = primType ? NXB::BuildBVH8<NXB::Triangle>((NXB::Triangle*)dPrims, n, cfg, outMetrics9 ? &metrics : nullptr)
                             : NXB::BuildBVH8<NXB::AABB>((NXB::AABB*)dPrims, n, cfg, outMetrics9 ? &metrics : nullptr);
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaMemcpy(outNodes, bvh.nodes, sizeof(NXB::BVH8::Node) * bvh.nodeCount, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outPrimIdx, bvh.primIdx, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    *outNodeCount = bvh.nodeCount;
    if (outBounds6) std::memcpy(outBounds6, &bvh.bounds, 24);
    copyMetrics(metrics, outMetrics9);
    NXB::FreeDeviceBVH(bvh);
    cudaFree(dPrims);
    return 0;
}

// NXB::BenchmarkBuild's protocol (BVHBuildMetrics.h:63-108: warm-up builds, then the mean of the per-stage metrics over the
// measured builds) without its std::cout report, on primitives uploaded once.  outMs2: [0] device time of the measured
// builds back to back (CUDA events around the loop, FreeDeviceBVH included as in BenchmarkBuild), [1] host wall time.
int nxref_benchmark_bvh8(const void* hostPrims, uint32_t n, int primType, int prioritizeSpeed, int warmup, int iters,
                         float* outMetrics9, uint32_t* outNodeCount, float* outMs2)
{
    size_t stride = primType ? sizeof(NXB::Triangle) : sizeof(NXB::AABB);
    void* dPrims = nullptr;
    REF_CHECK(cudaMalloc(&dPrims, stride * n));
    REF_CHECK(cudaMemcpy(dPrims, hostPrims, stride * n, cudaMemcpyHostToDevice));
    NXB::BuildConfig cfg; cfg.prioritizeSpeed = prioritizeSpeed != 0;
    NXB::BVHBuildMetrics agg{};
    uint32_t nodes = 0;
    auto build = [&](NXB::BVHBuildMetrics* m) {
        NXB::BVH8 bvh = primType ? NXB::BuildBVH8<NXB::Triangle>((NXB::Triangle*)dPrims, n, cfg, m) : NXB::BuildBVH8<NXB::AABB>((NXB::AABB*)dPrims, n, cfg, m);
        nodes = bvh.nodeCount;
        NXB::FreeDeviceBVH(bvh);
    };
    for (int i = 0; i < warmup; i++) { NXB::BVHBuildMetrics dummy{}; build(&dummy); }
    for (int i = 0; i < iters; i++) { NXB::BVHBuildMetrics m{}; build(&m); agg += m; }
    agg = agg / (float)iters;
    copyMetrics(agg, outMetrics9);
    if (outNodeCount) *outNodeCount = nodes;
    if (outMs2) {
        // the same builds without per-stage metrics (no event synchronisation inside): what a caller of BuildBVH8 waits for
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        REF_CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters; i++) build(nullptr);
        cudaEventRecord(e1, 0);
        REF_CHECK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&outMs2[0], e0, e1);
        outMs2[1] = 0.f;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaFree(dPrims);
    return 0;
}

// The reference's own Morton keys: ComputeSceneBoundsKernel + ComputeMortonCodesKernel launched as BuildBVH2 /
// BuildBVH2Impl launch them (B/src/BVHBuilder.cpp:119-153, 17-49).  outCodes: n x uint64 in primitive order.
} // extern "C"
template <typename PrimT, typename McT>
static int refMorton(const void* hostPrims, uint32_t n, uint64_t* outCodes, float* outBounds6)
{
    PrimT* dPrims = nullptr; McT* dCodes = nullptr;
    NXB::BVH2BuildState st{};
    st.primCount = n;
    REF_CHECK(cudaMalloc((void**)&dPrims, sizeof(PrimT) * n));
    REF_CHECK(cudaMemcpy(dPrims, hostPrims, sizeof(PrimT) * n, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMalloc((void**)&st.sceneBounds, sizeof(NXB::AABB)));
    REF_CHECK(cudaMalloc((void**)&st.nodes, sizeof(NXB::BVH2::Node) * (2 * (size_t)n - 1)));
    REF_CHECK(cudaMalloc((void**)&st.clusterIdx, 4 * (size_t)n));
    REF_CHECK(cudaMalloc((void**)&dCodes, sizeof(McT) * n));
    NXB::AABB sb; sb.Clear();
    REF_CHECK(cudaMemcpy(st.sceneBounds, &sb, sizeof(sb), cudaMemcpyHostToDevice));
    void* a1[2] = {&st, &dPrims};
    REF_CHECK(cudaLaunchKernel((void*)NXB::ComputeSceneBoundsKernel<PrimT>, dim3(occupancyGrid((const void*)NXB::ComputeSceneBoundsKernel<PrimT>, 64)), dim3(64), a1, 0, 0));
    void* a2[2] = {&st, &dCodes};
    REF_CHECK(cudaLaunchKernel((void*)NXB::ComputeMortonCodesKernel<McT>, dim3(occupancyGrid((const void*)NXB::ComputeMortonCodesKernel<McT>, 64)), dim3(64), a2, 0, 0));
    REF_CHECK(cudaDeviceSynchronize());
    std::vector<McT> h(n);
    REF_CHECK(cudaMemcpy(h.data(), dCodes, sizeof(McT) * n, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n; i++) outCodes[i] = (uint64_t)h[i];
    if (outBounds6) REF_CHECK(cudaMemcpy(outBounds6, st.sceneBounds, 24, cudaMemcpyDeviceToHost));
    cudaFree(dPrims); cudaFree(dCodes); cudaFree(st.sceneBounds); cudaFree(st.nodes); cudaFree(st.clusterIdx);
    return 0;
}
extern "C" {
int nxref_morton(const void* hostPrims, uint32_t n, int primType, int bits64, uint64_t* outCodes, float* outBounds6)
{
    if (primType) return bits64 ? refMorton<NXB::Triangle, uint64_t>(hostPrims, n, outCodes, outBounds6) : refMorton<NXB::Triangle, uint32_t>(hostPrims, n, outCodes, outBounds6);
    return bits64 ? refMorton<NXB::AABB, uint64_t>(hostPrims, n, outCodes, outBounds6) : refMorton<NXB::AABB, uint32_t>(hostPrims, n, outCodes, outBounds6);
}

// ------------------------------------------------------------------ scene ----
int nxref_scene_reset()
{
    cudaDeviceSynchronize();
    for (auto& m : g.meshes) { NXB::FreeDeviceBVH(m.bvh); cudaFree(m.tris); cudaFree(m.tridata); }
    g.meshes.clear();
    if (g.dMeshes) cudaFree(g.dMeshes), g.dMeshes = nullptr;
    if (g.dInstances) cudaFree(g.dInstances), g.dInstances = nullptr;
    if (g.dMaterials) cudaFree(g.dMaterials), g.dMaterials = nullptr;
    if (g.dLights) cudaFree(g.dLights), g.dLights = nullptr;
    if (g.hasTlas) NXB::FreeDeviceBVH(g.tlas), g.hasTlas = false;
    if (g.hasHdr) { cudaDestroyTextureObject(g.hdr); cudaFreeArray(g.hdrArray); g.hasHdr = false; }
    for (size_t i = 0; i < g.textures.size(); i++) { cudaDestroyTextureObject(g.textures[i]); cudaFreeArray(g.textureArrays[i]); }
    g.textures.clear(); g.textureArrays.clear();
    if (g.dTextures) cudaFree(g.dTextures), g.dTextures = nullptr;
    g.instanceCount = g.lightCount = 0;
    return 0;
}

// Mesh::Mesh (N/Assets/Mesh.h:15-46): upload + BuildBVH8<Triangle> with prioritizeSpeed = true.
int nxref_add_mesh(const float* tris /* n*9 */, const float* tridata /* n*24 */, uint32_t n)
{
    RefMesh m;
    REF_CHECK(cudaMalloc((void**)&m.tris, sizeof(NXB::Triangle) * n));
    REF_CHECK(cudaMalloc((void**)&m.tridata, sizeof(D_TriangleData) * n));
    REF_CHECK(cudaMemcpy(m.tris, tris, sizeof(NXB::Triangle) * n, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(m.tridata, tridata, sizeof(D_TriangleData) * n, cudaMemcpyHostToDevice));
    NXB::BuildConfig cfg; cfg.prioritizeSpeed = true;
    m.bvh = NXB::BuildBVH8<NXB::Triangle>(m.tris, n, cfg);
    g.meshes.push_back(m);
    return (int)g.meshes.size() - 1;
}

int nxref_mesh_bounds(int meshIdx, float* outBounds6)
{
    std::memcpy(outBounds6, &g.meshes[meshIdx].bvh.bounds, 24);
    return 0;
}

// instances: n * sizeof(D_MeshInstance) (160 B) in the reference's device layout.
// Uploads them, then Scene::BuildTLAS (N/Scene/Scene.cpp:65-78): BuildBVH8<AABB> with the default config.
int nxref_set_instances(const void* instances, uint32_t n)
{
    static_assert(sizeof(D_MeshInstance) == 160, "D_MeshInstance layout");
    if (g.dInstances) cudaFree(g.dInstances);
    if (g.hasTlas) NXB::FreeDeviceBVH(g.tlas), g.hasTlas = false;
    REF_CHECK(cudaMalloc((void**)&g.dInstances, sizeof(D_MeshInstance) * n));
    REF_CHECK(cudaMemcpy(g.dInstances, instances, sizeof(D_MeshInstance) * n, cudaMemcpyHostToDevice));
    g.instanceCount = n;

    std::vector<NXB::AABB> bounds(n);
    const D_MeshInstance* inst = (const D_MeshInstance*)instances;
    for (uint32_t i = 0; i < n; i++) std::memcpy(&bounds[i], &inst[i].bounds, sizeof(NXB::AABB));
    NXB::AABB* dBounds = nullptr;
    REF_CHECK(cudaMalloc((void**)&dBounds, sizeof(NXB::AABB) * n));
    REF_CHECK(cudaMemcpy(dBounds, bounds.data(), sizeof(NXB::AABB) * n, cudaMemcpyHostToDevice));
    g.tlas = NXB::BuildBVH8<NXB::AABB>(dBounds, n);
    g.hasTlas = true;
    cudaFree(dBounds);

    // AssetManager::AddMesh (N/Assets/AssetManager.cpp:24-33): D_Mesh[] array
    std::vector<D_Mesh> hm(g.meshes.size());
    for (size_t i = 0; i < g.meshes.size(); i++) { hm[i].bvh = g.meshes[i].bvh; hm[i].triangles = g.meshes[i].tris; hm[i].triangleData = g.meshes[i].tridata; }
    if (g.dMeshes) cudaFree(g.dMeshes);
    REF_CHECK(cudaMalloc((void**)&g.dMeshes, sizeof(D_Mesh) * hm.size()));
    REF_CHECK(cudaMemcpy(g.dMeshes, hm.data(), sizeof(D_Mesh) * hm.size(), cudaMemcpyHostToDevice));
    return 0;
}

int nxref_set_materials(const void* mats, uint32_t n)
{
    static_assert(sizeof(D_Material) == 92, "D_Material layout");
    if (g.dMaterials) cudaFree(g.dMaterials);
    REF_CHECK(cudaMalloc((void**)&g.dMaterials, sizeof(D_Material) * n));
    REF_CHECK(cudaMemcpy(g.dMaterials, mats, sizeof(D_Material) * n, cudaMemcpyHostToDevice));
    return 0;
}

// Texture::ToDevice (N/Assets/Texture.cpp:12-46) for one material map; returns its index in D_Scene::textures.
int nxref_add_texture(const void* rgba, uint32_t w, uint32_t h, int isHdr, int srgb)
{
    cudaChannelFormatDesc desc = isHdr ? cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat) : cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    cudaArray_t arr = nullptr;
    REF_CHECK(cudaMallocArray(&arr, &desc, w, h));
    const size_t pitch = (size_t)w * (isHdr ? 16 : 4);
    REF_CHECK(cudaMemcpy2DToArray(arr, 0, 0, rgba, pitch, pitch, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res; std::memset(&res, 0, sizeof(res)); res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc tex; std::memset(&tex, 0, sizeof(tex));
    tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeWrap; tex.sRGB = (srgb && !isHdr) ? 1 : 0;
    tex.filterMode = cudaFilterModeLinear; tex.readMode = isHdr ? cudaReadModeElementType : cudaReadModeNormalizedFloat; tex.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    REF_CHECK(cudaCreateTextureObject(&obj, &res, &tex, nullptr));
    g.textures.push_back(obj); g.textureArrays.push_back(arr);
    if (g.dTextures) cudaFree(g.dTextures);
    REF_CHECK(cudaMalloc((void**)&g.dTextures, sizeof(cudaTextureObject_t) * g.textures.size()));
    REF_CHECK(cudaMemcpy(g.dTextures, g.textures.data(), sizeof(cudaTextureObject_t) * g.textures.size(), cudaMemcpyHostToDevice));
    return (int)g.textures.size() - 1;
}

int nxref_set_lights(const void* lights, uint32_t n)
{
    static_assert(sizeof(D_Light) == 52, "D_Light layout");
    if (g.dLights) cudaFree(g.dLights), g.dLights = nullptr;
    if (n) {
        REF_CHECK(cudaMalloc((void**)&g.dLights, sizeof(D_Light) * n));
        REF_CHECK(cudaMemcpy(g.dLights, lights, sizeof(D_Light) * n, cudaMemcpyHostToDevice));
    }
    g.lightCount = n;
    return 0;
}

int nxref_set_camera(const void* dCamera88)
{
    static_assert(sizeof(D_Camera) == 88, "D_Camera layout");
    std::memcpy(&g.camera, dCamera88, sizeof(D_Camera));
    return 0;
}

int nxref_set_settings(int useMIS, int pathLength, const float* bgColor3, float bgIntensity)
{
    std::memset(&g.settings, 0, sizeof(g.settings));
    g.settings.useMIS = useMIS != 0;
    g.settings.visualizeBvh = false;
    g.settings.wireFrameBvh = false;
    g.settings.pathLength = (unsigned char)pathLength;
    g.settings.backgroundColor = make_float3(bgColor3[0], bgColor3[1], bgColor3[2]);
    g.settings.backgroundIntensity = bgIntensity;
    g.settings.toneMapping = ColorUtils::ToneMapping::NONE;
    g.settings.exposure = 0.0f;
    return 0;
}

// Texture::ToDevice for an RGBA32F equirect map (N/Assets/Texture.cpp:12-46; sRGB off as in Scene::AddHDRMap, N/Scene/Scene.cpp:102-107).
int nxref_set_hdr(const float* rgba, uint32_t w, uint32_t h)
{
    if (g.hasHdr) { cudaDestroyTextureObject(g.hdr); cudaFreeArray(g.hdrArray); g.hasHdr = false; }
    cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    REF_CHECK(cudaMallocArray(&g.hdrArray, &desc, w, h));
    REF_CHECK(cudaMemcpy2DToArray(g.hdrArray, 0, 0, rgba, w * 16, w * 16, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res; std::memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = g.hdrArray;
    cudaTextureDesc tex; std::memset(&tex, 0, sizeof(tex));
    tex.addressMode[0] = cudaAddressModeWrap; tex.addressMode[1] = cudaAddressModeWrap;
    tex.sRGB = 0; tex.filterMode = cudaFilterModeLinear; tex.readMode = cudaReadModeElementType; tex.normalizedCoords = 1;
    REF_CHECK(cudaCreateTextureObject(&g.hdr, &res, &tex, nullptr));
    g.hasHdr = true;
    return 0;
}

// ----------------------------------------------------------------- render ----
static int uploadSymbols()
{
    D_Scene s; std::memset(&s, 0, sizeof(s));
    s.hasHdrMap = g.hasHdr; s.hdrMap = g.hdr; s.textures = g.dTextures;
    s.lights = g.dLights; s.lightCount = g.lightCount;
    s.materials = g.dMaterials; s.camera = g.camera;
    s.meshInstances = g.dInstances;
    s.renderSettings = g.settings;
    s.renderSettings.resolution = g.camera.resolution;
    REF_CHECK(cudaMemcpy(GetDeviceSceneAddress(), &s, sizeof(D_Scene), cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(GetDeviceTLASAddress(), &g.tlas, sizeof(NXB::BVH8), cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(GetDeviceMeshesAdress(), &g.dMeshes, sizeof(D_Mesh*), cudaMemcpyHostToDevice));
    return 0;
}

static void freeRenderBuffers()
{
    cudaDeviceSynchronize();
    for (void* p : g.allocs) cudaFree(p);
    g.allocs.clear();
    if (g.graphExec) cudaGraphExecDestroy(g.graphExec), g.graphExec = nullptr;
    if (g.graph) cudaGraphDestroy(g.graph), g.graph = nullptr;
    if (g.graphStream) cudaStreamDestroy(g.graphStream), g.graphStream = nullptr;
}

// PathTracer::Reset (N/Renderer/PathTracer.cpp:61-159): queue allocation, launch shapes, per-bounce graph.
int nxref_render_init(uint32_t w, uint32_t h)
{
    freeRenderBuffers();
    g.w = w; g.h = h;
    const size_t count = (size_t)w * h;

    g.accum = devAlloc<float3>(count);
    g.renderBuffer = devAlloc<uint32_t>(count);
    g.rayTotals = devAlloc<unsigned long long>(2);
    REF_CHECK(cudaMemset(g.rayTotals, 0, 16));
    REF_CHECK(cudaMemcpy(GetDeviceAccumulationBufferAddress(), &g.accum, sizeof(float3*), cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(GetDeviceRenderBufferAddress(), &g.renderBuffer, sizeof(uint32_t*), cudaMemcpyHostToDevice));

    D_PathStateSOA ps;
    ps.lastPdf = devAlloc<float>(count); ps.throughput = devAlloc<float3>(count);
    ps.radiance = devAlloc<float3>(count); ps.allowMIS = devAlloc<bool>(count);
    REF_CHECK(cudaMemcpy(GetDevicePathStateAddress(), &ps, sizeof(ps), cudaMemcpyHostToDevice));

    auto makeIsect = [&]() { D_IntersectionSOA i; i.hitDistance = devAlloc<float>(count); i.instanceIdx = devAlloc<uint32_t>(count);
                             i.triIdx = devAlloc<uint32_t>(count); i.u = devAlloc<float>(count); i.v = devAlloc<float>(count); return i; };
    D_TraceRequestSOA tr;
    tr.intersection = makeIsect();
    tr.ray.origin = devAlloc<float3>(count); tr.ray.direction = devAlloc<float3>(count);
    tr.pixelIdx = devAlloc<uint32_t>(count);
    REF_CHECK(cudaMemcpy(GetDeviceTraceRequestAddress(), &tr, sizeof(tr), cudaMemcpyHostToDevice));

    D_ShadowTraceRequestSOA sr;
    sr.hitDistance = devAlloc<float>(count); sr.pixelIdx = devAlloc<uint32_t>(count); sr.radiance = devAlloc<float3>(count);
    sr.ray.origin = devAlloc<float3>(count); sr.ray.direction = devAlloc<float3>(count);
    REF_CHECK(cudaMemcpy(GetDeviceShadowTraceRequestAddress(), &sr, sizeof(sr), cudaMemcpyHostToDevice));

    D_MaterialRequestSOA mr;
    mr.intersection = makeIsect();
    mr.rayDirection = devAlloc<float3>(count); mr.pixelIdx = devAlloc<uint32_t>(count);
    REF_CHECK(cudaMemcpy(GetDeviceMaterialRequestAddress(), &mr, sizeof(mr), cudaMemcpyHostToDevice));

    D_PixelQuery pq; pq.pixelIdx = -1; pq.instanceIdx = -1;
    REF_CHECK(cudaMemcpy(GetDevicePixelQueryAddress(), &pq, sizeof(pq), cudaMemcpyHostToDevice));

    g.pixelGrid = dim3((unsigned)(count / BLOCK_SIZE + 1), 1, 1);
    g.traceGrid = dim3(occupancyGrid((const void*)TraceKernel, BLOCK_SIZE), 1, 1);
    g.shadowGrid = dim3(occupancyGrid((const void*)TraceShadowKernel, BLOCK_SIZE), 1, 1);

    // CUDAGraph (N/Device/Kernels/CUDAGraph.cpp:5-48): Logic -> Material -> (Trace || TraceShadow)
    REF_CHECK(cudaGraphCreate(&g.graph, 0));
    REF_CHECK(cudaStreamCreate(&g.graphStream));
    auto addNode = [&](void* fn, dim3 grid, cudaGraphNode_t* deps, size_t nDeps, cudaGraphNode_t* out) {
        cudaKernelNodeParams p; std::memset(&p, 0, sizeof(p));
        p.func = fn; p.gridDim = grid; p.blockDim = dim3(BLOCK_SIZE, 1, 1); p.kernelParams = nullptr; p.extra = nullptr; p.sharedMemBytes = 0;
        return cudaGraphAddKernelNode(out, g.graph, deps, nDeps, &p);
    };
    cudaGraphNode_t logic, material, trace, shadow;
    REF_CHECK(addNode((void*)LogicKernel, g.pixelGrid, nullptr, 0, &logic));
    REF_CHECK(addNode((void*)MaterialKernel, g.pixelGrid, &logic, 1, &material));
    REF_CHECK(addNode((void*)TraceKernel, g.traceGrid, &material, 1, &trace));
    REF_CHECK(addNode((void*)TraceShadowKernel, g.shadowGrid, &material, 1, &shadow));
    REF_CHECK(cudaGraphInstantiate(&g.graphExec, g.graph, 0));
    return uploadSymbols();
}

int nxref_update_scene() { return uploadSymbols(); }

// One frame: PathTracer::Render (N/Renderer/PathTracer.cpp:166-200) minus the GL map/unmap.
static int renderOneFrame(uint32_t frameNumber)
{
    uint32_t bounce = 0;
    REF_CHECK(cudaMemcpyAsync(GetDeviceFrameNumberAddress(), &frameNumber, 4, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpyAsync(GetDeviceBounceAddress(), &bounce, 4, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemsetAsync(GetDeviceQueueSizeAddress(), 0, sizeof(D_QueueSize)));
    void** noArgs = nullptr;
    REF_CHECK(cudaLaunchKernel((void*)GenerateKernel, g.pixelGrid, dim3(BLOCK_SIZE), noArgs, 0, 0));
    REF_CHECK(cudaLaunchKernel((void*)TraceKernel, g.traceGrid, dim3(BLOCK_SIZE), noArgs, 0, 0));
    bounce = 1;
    REF_CHECK(cudaMemcpyAsync(GetDeviceBounceAddress(), &bounce, 4, cudaMemcpyHostToDevice));
    for (uint32_t i = 0; i < g.settings.pathLength; i++)
    {
        REF_CHECK(cudaGraphLaunch(g.graphExec, g.graphStream));
        bounce = i + 2;
        REF_CHECK(cudaMemcpyAsync(GetDeviceBounceAddress(), &bounce, 4, cudaMemcpyHostToDevice));
    }
    REF_CHECK(cudaLaunchKernel((void*)AccumulateKernel, g.pixelGrid, dim3(BLOCK_SIZE), noArgs, 0, 0));
    AddQueueTotals<<<1, 32>>>(GetDeviceQueueSizeAddress(), g.rayTotals, g.settings.pathLength);
    return 0;
}

// Renders frames firstFrame .. firstFrame+nFrames-1 (frame numbers start at 1). If outMs != null the whole batch is
// timed with CUDA events on the legacy stream (which the per-bounce graph stream synchronises with implicitly).
int nxref_render(uint32_t firstFrame, uint32_t nFrames, float* outMs, unsigned long long* outRays2)
{
    cudaEvent_t e0, e1;
    REF_CHECK(cudaMemset(g.rayTotals, 0, 16));
    REF_CHECK(cudaEventCreate(&e0)); REF_CHECK(cudaEventCreate(&e1));
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaEventRecord(e0, 0));
    for (uint32_t f = 0; f < nFrames; f++)
        if (renderOneFrame(firstFrame + f)) return -1;
    REF_CHECK(cudaEventRecord(e1, 0));
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaGetLastError());
    if (outMs) REF_CHECK(cudaEventElapsedTime(outMs, e0, e1));
    if (outRays2) REF_CHECK(cudaMemcpy(outRays2, g.rayTotals, 16, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

int nxref_read_accum(float* outRgb /* w*h*3 */)
{
    REF_CHECK(cudaMemcpy(outRgb, g.accum, sizeof(float3) * (size_t)g.w * g.h, cudaMemcpyDeviceToHost));
    return 0;
}

// The reference's display transform (AccumulateKernel, N/Cuda/PathTracer/PathTracer.cu:513-549 with N/Utils/ColorUtils.h:27-212)
// applied to a caller-supplied linear image: the image is written to pathState.radiance, frameNumber is set to 1 (so the
// running mean becomes the image itself) and AccumulateKernel fills the RGBA8 render buffer.  count must be w*h of the last
// nxref_render_init.  Overwrites the accumulation buffer.
int nxref_display(int toneMapping, float exposure, const float* rgb, uint32_t count, uint32_t* outRgba)
{
    if ((size_t)count != (size_t)g.w * g.h) return -2;
    D_PathStateSOA ps;
    REF_CHECK(cudaMemcpy(&ps, GetDevicePathStateAddress(), sizeof(ps), cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(ps.radiance, rgb, 12ull * count, cudaMemcpyHostToDevice));
    g.settings.toneMapping = (ColorUtils::ToneMapping)toneMapping;
    g.settings.exposure = exposure;
    if (uploadSymbols()) return -1;
    const uint32_t one = 1;
    REF_CHECK(cudaMemcpy(GetDeviceFrameNumberAddress(), &one, 4, cudaMemcpyHostToDevice));
    void** noArgs = nullptr;
    REF_CHECK(cudaLaunchKernel((void*)AccumulateKernel, g.pixelGrid, dim3(BLOCK_SIZE), noArgs, 0, 0));
    REF_CHECK(cudaDeviceSynchronize());
    REF_CHECK(cudaMemcpy(outRgba, g.renderBuffer, 4ull * count, cudaMemcpyDeviceToHost));
    g.settings.toneMapping = ColorUtils::ToneMapping::NONE; g.settings.exposure = 0.0f;
    return uploadSymbols();
}

// Closest hit for a caller-supplied ray batch through the reference TraceKernel (bounce 0 queue).
// n must be <= w*h of the last nxref_render_init.
int nxref_trace(const float* origins, const float* dirs, uint32_t n,
                float* outT, float* outU, float* outV, uint32_t* outTri, uint32_t* outInst, float* outMs)
{
    if ((size_t)n > (size_t)g.w * g.h) return -2;
    D_TraceRequestSOA tr;
    REF_CHECK(cudaMemcpy(&tr, GetDeviceTraceRequestAddress(), sizeof(tr), cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(tr.ray.origin, origins, 12ull * n, cudaMemcpyHostToDevice));
    REF_CHECK(cudaMemcpy(tr.ray.direction, dirs, 12ull * n, cudaMemcpyHostToDevice));
    D_QueueSize* q = GetDeviceQueueSizeAddress();
    REF_CHECK(cudaMemset(q, 0, sizeof(D_QueueSize)));
    int32_t size = (int32_t)n;
    REF_CHECK(cudaMemcpy(&q->traceSize[0], &size, 4, cudaMemcpyHostToDevice));
    uint32_t bounce = 0;
    REF_CHECK(cudaMemcpy(GetDeviceBounceAddress(), &bounce, 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    REF_CHECK(cudaEventCreate(&e0)); REF_CHECK(cudaEventCreate(&e1));
    REF_CHECK(cudaEventRecord(e0, 0));
    void** noArgs = nullptr;
    REF_CHECK(cudaLaunchKernel((void*)TraceKernel, g.traceGrid, dim3(BLOCK_SIZE), noArgs, 0, 0));
    REF_CHECK(cudaEventRecord(e1, 0));
    REF_CHECK(cudaDeviceSynchronize());
    if (outMs) REF_CHECK(cudaEventElapsedTime(outMs, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    REF_CHECK(cudaMemcpy(outT, tr.intersection.hitDistance, 4ull * n, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outU, tr.intersection.u, 4ull * n, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outV, tr.intersection.v, 4ull * n, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outTri, tr.intersection.triIdx, 4ull * n, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outInst, tr.intersection.instanceIdx, 4ull * n, cudaMemcpyDeviceToHost));
    return 0;
}

int nxref_tlas_to_host(void* outNodes, uint32_t* outPrimIdx, uint32_t* outNodeCount)
{
    REF_CHECK(cudaMemcpy(outNodes, g.tlas.nodes, sizeof(NXB::BVH8::Node) * g.tlas.nodeCount, cudaMemcpyDeviceToHost));
    REF_CHECK(cudaMemcpy(outPrimIdx, g.tlas.primIdx, 4ull * g.tlas.primCount, cudaMemcpyDeviceToHost));
    *outNodeCount = g.tlas.nodeCount;
    return 0;
}

int nxref_sizes(uint32_t* out /* [0]=D_MeshInstance [1]=D_Material [2]=D_Light [3]=D_Camera [4]=D_Scene [5]=D_Mesh */)
{
    out[0] = sizeof(D_MeshInstance); out[1] = sizeof(D_Material); out[2] = sizeof(D_Light);
    out[3] = sizeof(D_Camera); out[4] = sizeof(D_Scene); out[5] = sizeof(D_Mesh);
    return 0;
}

} // extern "C"
