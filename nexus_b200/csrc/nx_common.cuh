// Shared device/host helpers for the nexus_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/nexus_b200.h"

#define NX_WARP 32u
#define NX_FULL 0xffffffffu
#define NX_INVALID 0xffffffffu

// bump allocator over one device allocation
struct nx_bump { char* base = nullptr; size_t cap = 0, used = 0; };

// ---------------------------------------------------------------------------------------------- context ----
struct nx_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;      // main stream: build, generate, closest-hit trace, shade
    cudaStream_t stream_aux = nullptr;  // shadow rays overlap the next extension trace
    std::string error;
    // traversal batching thresholds in lanes (traverse.cuh TraceTuning); overridable with NX_TRACE_TUNE="tri,inst"
    uint32_t tune_tri = 6, tune_inst = 8, tune_sphere = 1;
    uint32_t tune_tri_any = 6, tune_inst_any = 8;   // any-hit kernel (NX_TRACE_TUNE_ANY)
    bool tune_user = false;              // thresholds set by the caller / environment: used for every scene kind.  Otherwise scenes that are one
                                         // merged BLAS (no instance entries: the set-up phase is only the ray fetch) run with 8 / 4, measured 1.8 %
                                         // faster than 6 / 8 on the 10M-triangle scene (profiles/r02_tune_direct.log)
    // BVH2 -> BVH8 collapse used for the BLASes and the TLAS the scene code builds (nx_build_config::collapse / max_leaf_prims).
    // Default: the SAH-optimal collapse with at most 2 primitives per leaf - same hits, 22 % fewer node visits per ray than the
    // reference GPU converter's trees on the 10M-triangle scene (25.5 -> 22.7 ms per 4K frame); nx_ctx_set_scene_collapse /
    // NX_SCENE_COLLAPSE="0,0" restore trees identical to NexusBVH's.
    int scene_collapse = NX_COLLAPSE_SAH_OPTIMAL, scene_max_leaf_prims = 2;
    // Instances whose mesh is used once (and that have not been moved) are transformed to world space and share one BLAS: no instance
    // entry, no per-object tree overlap.  Off in the NexusBVH-identical mode.  NX_MERGE_INSTANCES=0 / nx_ctx_set_instance_merging.
    int merge_instances = 1;
    uint32_t merge_min_prims = 0;        // meshes with fewer primitives stay instances of their own (0: every single-use mesh is merged; measured on BASELINE configs[2]:
                                         // keeping the two-triangle ground and light out costs 30 %: each is then a scene-sized TLAS entry every ray enters) NX_MERGE_MIN_PRIMS
    int scene_blas_speed = 1;   // Mesh::Mesh builds its BLAS with prioritizeSpeed = true (32-bit Morton keys), N/Assets/Mesh.h:37
    // L2 set-aside for persisting accesses (top-level nodes + instance records of the scene being rendered); 0 = hints off
    size_t l2_persist_bytes = 0, l2_window_max = 0;
    // Traversal loop: 1 = ray pool (traverse_pool.cuh: 64 rays per warp in shared memory, lanes take rays by kind of work), 0 = one
    // ray per lane (traverse.cuh trace_loop).  Same hits either way.  NX_TRACE_MODE / nx_ctx_set_trace_mode.
    int trace_mode = 0;
    uint32_t pool_node = 28, pool_tri = 24, pool_inst = 16, pool_fetch = 16;           // ray-pool round thresholds in rays (NX_POOL_TUNE="n,t,i,f")
    uint32_t pool_node_any = 28, pool_tri_any = 24, pool_inst_any = 16, pool_fetch_any = 16;
    uint32_t stack_limit = 40;           // NX_STACK_TOTAL; nx_ctx_set_stack_limit lowers it in the overflow test
    uint32_t* dOverflow = nullptr;       // device counter of refused traversal-stack pushes (TraceScene::overflow)
    uint32_t* hOverflow = nullptr;       // pinned mirror ([1]: rays with non-finite origin / direction, answered as misses)
    unsigned long long nonfinite_rays = 0;
    void* poolSpill[2] = {nullptr, nullptr}; size_t poolSpillWarps[2] = {0, 0};       // global spill stacks of the two trace streams
    int tlas_refit = 0;                  // Scene::Update refits the TLAS instead of rebuilding it when the entry set is unchanged (nx_ctx_set_tlas_refit)
    int trace_generic = 0;               // 1: always the general traversal loop, never the scene-kind specialisations (NX_TRACE_GENERIC; tests)
    int gridCache[24] = {0};             // persistent-grid sizes per kernel (occupancy x SM count of THIS context's device)
    int sort_mode = 1;                   // 1 = radix_sort.cuh (own onesweep sort), 0 = cub::DeviceRadixSort (measurement only); NX_SORT=0|1
    int dp_waves = 1;                    // 1 = C(n, i) tables level by level on builds of 200k+ primitives (NX_DP_WAVES=0: always the climb)
    int collapse_cta = 1;                // 1 = single-block collapse for builds of up to 40k primitives (NX_COLLAPSE_CTA=0: always the grid-wide kernel)
    int hploc_mode = 2;                  // 0 = one-phase kernel, 1 = two-phase (block-local phase in shared memory + global phase), 2 = one-phase with the shared-memory merge table; NX_HPLOC
    // Scene set-up pipeline (scene.cu add_mesh): BLAS builds of successive meshes are issued round-robin on these streams without any
    // host synchronisation; host data reaches the device through a pinned staging ring.  nx_scene_update waits for all of them once.
    cudaStream_t buildStreams[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int buildStreamCount = 0;
    char* stagePinned = nullptr; size_t stageBytes = 0, stageUsed = 0;
    nx_bump buildWsStore[8];             // one workspace per build stream (temporaries of the BLAS build running on it)
    nx_bump* buildWs = nullptr;          // set while a workspace build is being issued (bvh_builder.cu allocAsync)
    nx_bump* outArena = nullptr;         // where such a build puts its outputs
    // scratch reused by the builder's parity hook
    std::vector<uint64_t> dbg_codes;
};

#define NX_FAIL(ctx, code, ...) do { char b_[512]; std::snprintf(b_, sizeof(b_), __VA_ARGS__); (ctx)->error = b_; return (code); } while (0)
#define NX_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    char b_[512]; std::snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    (ctx)->error = b_; return NX_ERR_CUDA; } } while (0)

// context.cu: synchronises the context's streams and turns a non-zero traversal-stack overflow counter into NX_ERR_STATE
int nxi_check_overflow(nx_ctx* ctx);

// Runs builder code (which issues everything on ctx->stream) on another stream for the lifetime of the object.
struct StreamSwap {
    nx_ctx* c; cudaStream_t old;
    StreamSwap(nx_ctx* ctx, cudaStream_t s) : c(ctx), old(ctx->stream) { ctx->stream = s; }
    ~StreamSwap() { c->stream = old; }
};

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// ------------------------------------------------------------------------------------------- float math ----
// All geometry arithmetic that decides topology or hit ids is written with explicit-rounding intrinsics so the compiler
// can neither contract nor reorder it; the CPU oracle restates the same operation sequence with fmaf().
struct V3 { float x, y, z; };
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__host__ __device__ __forceinline__ V3 v3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 vmin3(V3 a, V3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
__device__ __forceinline__ V3 vmax3(V3 a, V3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
// dot(a,b) = fma(a.x,b.x, fma(a.y,b.y, a.z*b.z))
__device__ __forceinline__ float xdot(V3 a, V3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z))); }
// cross(a,b).x = fma(a.y,b.z, -(a.z*b.y)) ...
__device__ __forceinline__ V3 xcross(V3 a, V3 b)
{
    return {__fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x))};
}

struct Box { V3 lo, hi; };
// Half surface area in the operation order the reference compiles to for sm_100a (mul dy*dz; fma dx*dy; fma dx*dz) with
// flush-to-zero, because H-PLOC compares these values bit for bit (B/include/NXB/AABB.h:43-47, BinaryBuilder.cu:143-148).
__device__ __forceinline__ float half_area_ref(const Box& b)
{
    float dx, dy, dz, t;
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(dx) : "f"(b.hi.x), "f"(b.lo.x));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(dy) : "f"(b.hi.y), "f"(b.lo.y));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(dz) : "f"(b.hi.z), "f"(b.lo.z));
    asm("mul.ftz.f32 %0, %1, %2;" : "=f"(t) : "f"(dy), "f"(dz));
    asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(dx), "f"(dy), "f"(t));
    asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(t) : "f"(dx), "f"(dz), "f"(t));
    return t;
}
__device__ __forceinline__ void box_grow(Box& a, const Box& b) { a.lo = vmin3(a.lo, b.lo); a.hi = vmax3(a.hi, b.hi); }

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__host__ __device__ __forceinline__ uint32_t f2ord(float f)
{
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    std::memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t k)
{
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t bits_below(uint32_t x, uint32_t i) { return __popc(x & ((1u << i) - 1u)); }

// 16-byte cache-global (L2) loads/stores for data produced by other SMs within the same kernel
__device__ __forceinline__ float4 ld_cg4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg4(float4* p, float4 v) { __stcg(p, v); }

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
