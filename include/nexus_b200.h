/* nexus_b200 — C ABI of the B200-native (sm_100a) wavefront path-tracing hot path.
 *
 * Drop-in boundary for three C++ surfaces of StokastX/Nexus (reference paths relative to /root/reference/Nexus):
 *   builder  : NXB::BuildBVH2 / BuildBVH8 / ToHost / FreeDeviceBVH / BenchmarkBuild
 *              vendor/NexusBVH/NexusBVH/include/NXB/BVHBuilder.h:19-55, BVHBuildMetrics.h:7-108, BuildConfig.h:6-12
 *   scene    : Scene / AssetManager / Mesh / MeshInstance / Material / Light / Camera / RenderSettings
 *              src/Scene/Scene.h:19-49, src/Assets/AssetManager.h:18-44, src/Scene/MeshInstance.h:22-66,
 *              src/Assets/Material.h:6-26, src/Scene/Light.h:10-54, src/Scene/Camera.h:9-52, src/Renderer/RenderSettings.h:5-17
 *   renderer : PathTracer::{Reset, ResetFrameNumber, Render, OnResize, UpdateDeviceScene, GetFrameNumber}
 *              src/Renderer/PathTracer.h:12-29 (+ the six kernels of src/Cuda/PathTracer/PathTracer.cuh:69-74)
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative nx_status and records a
 * message retrievable with nx_last_error() (the reference prints and exit(99)s: src/Utils/Utils.cpp:3-12).  One host
 * thread per context.  All device work of a context is issued on the context's own streams; there is no process-global
 * device state, so one process can drive several GPUs (the reference keeps its state in __constant__ symbols,
 * src/Cuda/PathTracer/PathTracer.cu:21-37, which limits it to one renderer per process).
 */
#ifndef NEXUS_B200_H
#define NEXUS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NX_ABI_VERSION 1

typedef enum nx_status {
    NX_OK = 0,
    NX_ERR_CUDA = -1,
    NX_ERR_INVALID = -2,
    NX_ERR_OOM = -3,
    NX_ERR_STATE = -4
} nx_status;

typedef struct nx_ctx nx_ctx;
typedef struct nx_scene nx_scene;
typedef struct nx_renderer nx_renderer;

/* ---------------------------------------------------------------- PODs (layouts are part of the contract) ---- */
typedef struct nx_aabb { float bmin[3]; float bmax[3]; } nx_aabb;                      /* NXB::AABB, 24 B            */
typedef struct nx_triangle { float v0[3], v1[3], v2[3]; } nx_triangle;                 /* NXB::Triangle, 36 B        */
typedef struct nx_triangle_data {                                                       /* D_TriangleData, 96 B       */
    float normal0[3], normal1[3], normal2[3];
    float tangent0[3], tangent1[3], tangent2[3];
    float uv0[2], uv1[2], uv2[2];
} nx_triangle_data;

/* NXB::BVH2::Node (32 B): leaf <=> left == 0xffffffff, then right = primitive id.  Leaves are [0,n), root = 2n-2. */
typedef struct nx_bvh2_node { nx_aabb bounds; uint32_t left, right; } nx_bvh2_node;
/* NXB::BVH8::NodeExplicit (80 B, 16-byte aligned on the device). */
typedef struct nx_bvh8_node {
    float p[3]; uint8_t e[3]; uint8_t imask;
    uint32_t child_base, prim_base;
    uint8_t meta[8];
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
} nx_bvh8_node;

/* Handles returned by value like NXB::BVH2 / NXB::BVH8 (BVH.h:18-96); pointers are DEVICE pointers owned by the library. */
typedef struct nx_bvh2 { nx_bvh2_node* nodes; uint32_t node_count; uint32_t prim_count; nx_aabb bounds; } nx_bvh2;
typedef struct nx_bvh8 { nx_bvh8_node* nodes; uint32_t node_count; uint32_t* prim_idx; uint32_t prim_count; nx_aabb bounds; } nx_bvh8;

/* NXB::BuildConfig (BuildConfig.h:6-12) plus the choice of BVH2 -> BVH8 collapse:
 *   NX_COLLAPSE_REFERENCE_GPU  (0, default) the reference GPU converter's rule (WideConverter.cu:225-414): trees identical to NexusBVH's;
 *   NX_COLLAPSE_SAH_OPTIMAL    the SAH-optimal collapse of the reference's CPU BVH8Builder (src/Geometry/BVH/BVH8Builder.cpp:31-199:
 *                              C(n, i) table, C_PRIM 0.3, C_NODE 1.0), leaf children of up to max_leaf_prims (1..3, 0 = P_MAX = 3)
 *                              primitives, on the GPU; node layout, slot assignment and quantisation as in the GPU converter. */
enum { NX_COLLAPSE_REFERENCE_GPU = 0, NX_COLLAPSE_SAH_OPTIMAL = 1 };
typedef struct nx_build_config { int prioritize_speed; int collapse; int max_leaf_prims; } nx_build_config;
typedef struct nx_build_metrics {                                                       /* NXB::BVHBuildMetrics       */
    float scene_bounds_ms, morton_ms, sort_ms, bvh2_ms, bvh8_ms, total_ms;
    float bvh2_cost, bvh8_cost, avg_children_per_node;
} nx_build_metrics;

typedef struct nx_material {                                                            /* Material / D_Material, 92 B */
    float base_color[3]; float metalness; float roughness; float anisotropy; float specular_weight;
    float specular_color[3]; float ior; float transmission;
    float emission_color[3]; float intensity; float opacity;
    int32_t base_color_map, emissive_map, normal_map, roughness_map, metalness_map, metallic_roughness_map;
} nx_material;

typedef enum nx_light_type { NX_LIGHT_POINT = 0, NX_LIGHT_SPOT = 1, NX_LIGHT_DIRECTIONAL = 2, NX_LIGHT_MESH = 3 } nx_light_type;
typedef struct nx_light {                 /* flattened Light union (src/Scene/Light.h:10-54) */
    int32_t type;
    float position[3]; float direction[3]; float color[3]; float intensity;
    float falloff_start, falloff_end;
    uint32_t instance;                    /* NX_LIGHT_MESH: emissive instance index (Light::mesh.meshId) */
} nx_light;

typedef struct nx_camera {                /* Camera ctor arguments, src/Scene/Camera.cpp:24-30 */
    float position[3]; float forward[3]; float right[3];   /* right = {0,0,0} => cross(forward, +Y) as the ctor does */
    float horizontal_fov_deg; float focus_distance; float defocus_angle_deg;
} nx_camera;

/* ColorUtils::ToneMapping, src/Utils/ColorUtils.h:9-16 */
enum { NX_TONE_NONE = 0, NX_TONE_ACES = 1, NX_TONE_UNCHARTED2 = 2, NX_TONE_AGX_DEFAULT = 3, NX_TONE_AGX_GOLDEN = 4, NX_TONE_AGX_PUNCHY = 5 };

typedef struct nx_render_settings {       /* RenderSettings, src/Renderer/RenderSettings.h:5-17 */
    int32_t use_mis; int32_t path_length;
    float background_color[3]; float background_intensity;
    int32_t tone_mapping; float exposure; /* display transform only; the accumulation buffer stays linear */
} nx_render_settings;

typedef struct nx_ray { float origin[3]; float tmax; float direction[3]; uint32_t pad; } nx_ray;     /* 32 B */
typedef struct nx_hit { float t, u, v; uint32_t prim; uint32_t instance; } nx_hit;                    /* D_Intersection */

typedef struct nx_frame_stats {           /* queue totals of the frames rendered by the last nx_render_frames call */
    uint64_t extension_rays; uint64_t shadow_rays; uint64_t shaded_hits; uint64_t frames;
    float device_ms;                      /* CUDA-event time of that call on the context's stream */
    uint32_t kernel_launches;
} nx_frame_stats;

/* Per-kernel device times and traversal work of the last nx_renderer_render call (see nx_renderer_set_profiling).
 * Kernel classes: 0 generate, 1 closest-hit trace, 2 shade, 3 any-hit (shadow) trace. */
typedef struct nx_kernel_profile {
    float ms[4]; uint32_t launches[4];
    uint64_t closest_work[4];             /* nodes visited, triangles tested, instances entered, rays (flags & 2) */
    uint64_t any_work[4];
    /* warp scheduling of the traversal loop (flags & 2), summed over warps: loop iterations, lanes that tested a node,
     * triangle rounds, lanes in them, set-up (new ray / instance entry) rounds, lanes in them, instances culled by their sphere;
     * ray-pool loop only: node rounds, fetch rounds, rays fetched */
    uint64_t closest_sched[10];
    uint64_t any_sched[10];
} nx_kernel_profile;

/* ------------------------------------------------------------------------------------------------ context ---- */
int nx_abi_version(void);
int nx_ctx_create(int device, nx_ctx** out);
void nx_ctx_destroy(nx_ctx* ctx);
const char* nx_last_error(const nx_ctx* ctx);
int nx_ctx_synchronize(nx_ctx* ctx);
int nx_ctx_sm_count(const nx_ctx* ctx);
void* nx_ctx_stream(nx_ctx* ctx);                                   /* cudaStream_t all work is issued on */
/* Traversal batching thresholds, in lanes of a warp: the triangle / instance-entry phase of the traversal kernels runs when
 * at least this many lanes have such work (or one lane has nothing else to do).  Results do not depend on them (equal-distance
 * hits are resolved by id); they trade SIMT efficiency against front-to-back culling.  The reference's counterpart is the
 * compile-time 20 % postponing rule (BVH8Traversal.cuh:15-22,270-278).  Also settable with NX_TRACE_TUNE="tri,inst". */
int nx_ctx_set_trace_tuning(nx_ctx* ctx, uint32_t tri_lanes, uint32_t inst_lanes);
/* Collapse used for the BLASes / TLAS that nx_scene_* builds from now on (default NX_COLLAPSE_SAH_OPTIMAL, 2 primitives per leaf;
 * NX_COLLAPSE_REFERENCE_GPU reproduces NexusBVH's trees).  Hit ids and distances do not depend on the choice. */
int nx_ctx_set_scene_collapse(nx_ctx* ctx, int collapse, int max_leaf_prims);
/* Traversal loop of the trace kernels.  NX_TRACE_LANE (default, the fastest measured): one ray per lane, the loop specialised for what
 * the scene holds (two-level only / merged BLAS only / both); NX_TRACE_LANE_GENERAL: the same loop without the specialisation.
 * NX_TRACE_POOL: a warp owns 64 rays whose state lives in shared memory and hands its lanes the rays that want the kind of work of
 * the round (node test, triangle test, instance entry, fetch).  NX_TRACE_DUO: two rays per lane.  Hits are identical in all of them
 * (same arithmetic, deterministic tie-break).  Also NX_TRACE_MODE=lane|pool|duo, NX_TRACE_GENERIC=1. */
enum { NX_TRACE_LANE = 0, NX_TRACE_POOL = 1, NX_TRACE_DUO = 2, NX_TRACE_LANE_GENERAL = 3 };
int nx_ctx_set_trace_mode(nx_ctx* ctx, int mode);
/* Scene::Update after instances moved: 0 (default, the reference's behaviour) rebuilds the TLAS; 1 refits it (nx_bvh8_refit_aabb) whenever
 * the set of TLAS entries is unchanged - cheaper per edit, same hits, tree quality degrades as objects travel.  NX_TLAS_REFIT=0|1. */
int nx_ctx_set_tlas_refit(nx_ctx* ctx, int enabled);
/* Ray-pool round thresholds, in rays of the warp's pool: a round of that kind runs when at least this many rays want it (else the
 * fullest kind runs).  any_hit != 0 sets the shadow-ray kernel's.  Also NX_POOL_TUNE / NX_POOL_TUNE_ANY = "node,tri,inst,fetch". */
int nx_ctx_set_pool_tuning(nx_ctx* ctx, int any_hit, uint32_t node_rays, uint32_t tri_rays, uint32_t inst_rays, uint32_t fetch_rays);
/* Traversal-stack entries per ray (2..40, default 40; the lane-bound loop never goes below its 8 shared-memory entries).  A push beyond the limit is refused and COUNTED; the next nx_ctx_synchronize /
 * nx_renderer_stats / nx_trace_* returns NX_ERR_STATE with the count in nx_last_error (the reference's 32-entry stack overflows
 * silently, BVH8Traversal.cuh:164).  Lowering the limit exists for the test of that report. */
int nx_ctx_set_stack_limit(nx_ctx* ctx, uint32_t entries);
/* Instance merging (default on; off in the NexusBVH-identical collapse mode): at Scene::Update the instances whose mesh no other instance
 * uses, and that have not been moved since they were created, are transformed to world space and share ONE BLAS under the TLAS.  Hit
 * records are unchanged (instance id, primitive id inside the mesh, world-space distance); an instance that is moved later gets a
 * BLAS of its own, so moving things still only rebuilds the TLAS.  Applies to scenes updated after the call.  NX_MERGE_INSTANCES=0|1. */
int nx_ctx_set_instance_merging(nx_ctx* ctx, int enabled);
/* 0 switches the per-instance bounding-sphere test off (measurement only; results are identical either way). */
int nx_ctx_set_sphere_cull(nx_ctx* ctx, int enabled);

/* device memory helpers so a C / ctypes host needs no CUDA runtime of its own (replace N/Device/CudaMemory.h) */
int nx_malloc(nx_ctx* ctx, size_t bytes, void** out_dev);
int nx_free(nx_ctx* ctx, void* dev);
int nx_memcpy_h2d(nx_ctx* ctx, void* dev, const void* host, size_t bytes);
int nx_memcpy_d2h(nx_ctx* ctx, void* host, const void* dev, size_t bytes);

/* ------------------------------------------------------------------------------------------------ builder ---- */
/* d_prims: DEVICE pointer, caller-owned, not modified (as NXB::BuildBVH*).  metrics may be NULL; when non-NULL every
 * stage is bracketed by CUDA events like the reference (BVHBuilder.cpp:35-46).  Blocking like the reference. */
int nx_bvh2_build_tri(nx_ctx* ctx, const nx_triangle* d_prims, uint32_t n, const nx_build_config* cfg, nx_build_metrics* metrics, nx_bvh2* out);
int nx_bvh2_build_aabb(nx_ctx* ctx, const nx_aabb* d_prims, uint32_t n, const nx_build_config* cfg, nx_build_metrics* metrics, nx_bvh2* out);
int nx_bvh8_build_tri(nx_ctx* ctx, const nx_triangle* d_prims, uint32_t n, const nx_build_config* cfg, nx_build_metrics* metrics, nx_bvh8* out);
int nx_bvh8_build_aabb(nx_ctx* ctx, const nx_aabb* d_prims, uint32_t n, const nx_build_config* cfg, nx_build_metrics* metrics, nx_bvh8* out);
/* Refit of a BVH8 that nx_bvh8_build_aabb produced, in place: same topology and leaf order, node frames and child boxes recomputed
 * bottom-up from d_bounds (prim_count boxes in device memory, primitive order); bvh->bounds is updated.  No counterpart in the
 * reference, which rebuilds its TLAS on every instance change (Scene::BuildTLAS, src/Scene/Scene.cpp:65-78).  Blocking. */
int nx_bvh8_refit_aabb(nx_ctx* ctx, nx_bvh8* bvh, const nx_aabb* d_bounds);
int nx_bvh2_to_host(nx_ctx* ctx, const nx_bvh2* bvh, nx_bvh2_node* host_nodes);                       /* NXB::ToHost */
int nx_bvh8_to_host(nx_ctx* ctx, const nx_bvh8* bvh, nx_bvh8_node* host_nodes, uint32_t* host_prim_idx);
int nx_bvh2_free(nx_ctx* ctx, nx_bvh2* bvh);                                                           /* FreeDeviceBVH */
int nx_bvh8_free(nx_ctx* ctx, nx_bvh8* bvh);
/* NXB::BenchmarkBuild: averaged per-stage times over `iters` builds after `warmup` builds (prim_type 0 = AABB, 1 = triangle). */
int nx_bvh8_benchmark(nx_ctx* ctx, const void* d_prims, uint32_t n, int prim_type, const nx_build_config* cfg,
                      int warmup, int iters, nx_build_metrics* metrics, uint32_t* out_node_count);
/* Parity hook: the sorted Morton keys / primitive order the last *_build call of this context used (uint64 per key). */
int nx_bvh_debug_morton(nx_ctx* ctx, const void* d_prims, uint32_t n, int prim_type, int bits64, uint64_t* host_codes /* n, primitive order */);

/* -------------------------------------------------------------------------------------------------- scene ---- */
int nx_scene_create(nx_ctx* ctx, uint32_t width, uint32_t height, nx_scene** out);    /* Scene(uint2 resolution) */
void nx_scene_destroy(nx_scene* scene);
int nx_scene_add_material(nx_scene* scene, const nx_material* m);                       /* returns index >= 0  */
int nx_scene_set_material(nx_scene* scene, uint32_t idx, const nx_material* m);         /* InvalidateMaterial  */
/* AssetManager::AddMesh: host triangles + per-triangle shading data (may be NULL => geometric normals); builds the BLAS
 * with prioritizeSpeed = true exactly like Mesh::Mesh (src/Assets/Mesh.h:15-46).  Returns mesh index >= 0. */
int nx_scene_add_mesh(nx_scene* scene, const nx_triangle* tris, const nx_triangle_data* data, uint32_t n, uint32_t material_idx);
/* Sharded BLAS builds ("BVH build, many meshes"): BLAS builds are independent units, so ranks build disjoint
 * subsets and exchange the results.  nx_scene_build_blas builds one mesh's BLAS from HOST triangles exactly as nx_scene_add_mesh
 * would (the context's scene collapse mode, leaf size and Morton width) and returns an owned handle (free with nx_bvh8_free).
 * nx_scene_add_mesh_prebuilt is nx_scene_add_mesh with the BLAS supplied instead of built: `d_nodes` / `d_prim_idx` are DEVICE
 * pointers (e.g. into an NCCL all-gather buffer) and are copied; the result is indistinguishable from a locally built mesh. */
int nx_scene_build_blas(nx_ctx* ctx, const nx_triangle* host_tris, uint32_t n, nx_bvh8* out);
int nx_scene_add_mesh_prebuilt(nx_scene* scene, const nx_triangle* tris, const nx_triangle_data* data, uint32_t n, uint32_t material_idx,
                               const nx_bvh8_node* d_nodes, uint32_t node_count, const uint32_t* d_prim_idx, const nx_aabb* bounds);
int nx_scene_mesh_bounds(nx_scene* scene, uint32_t mesh_idx, nx_aabb* out);
int nx_scene_mesh_bvh(nx_scene* scene, uint32_t mesh_idx, nx_bvh8* out);                /* borrowed handle */
/* Scene::CreateMeshInstance + MeshInstance::SetTransform (T * Rz * Ry * Rx * S, Euler degrees). Returns instance index. */
int nx_scene_add_instance(nx_scene* scene, uint32_t mesh_idx, int32_t material_idx /* <0: mesh default */,
                          const float position[3], const float rotation_deg[3], const float scale[3]);
/* Same, with an explicit row-major 4x4 (used by parity tests so both arms see identical matrices). */
int nx_scene_add_instance_matrix(nx_scene* scene, uint32_t mesh_idx, int32_t material_idx, const float m[16]);
int nx_scene_set_instance_transform(nx_scene* scene, uint32_t inst, const float position[3], const float rotation_deg[3], const float scale[3]);
int nx_scene_set_instance_material(nx_scene* scene, uint32_t inst, int32_t material_idx);      /* MeshInstance::AssignMaterial + InvalidateMeshInstance */
int nx_scene_instance_matrix(nx_scene* scene, uint32_t inst, float out_row_major[16]);         /* MeshInstance::GetTransfromationMatrix */
int nx_scene_instance_bounds(nx_scene* scene, uint32_t inst, nx_aabb* out);                    /* MeshInstance::GetBounds (world AABB) */
int nx_scene_add_light(nx_scene* scene, const nx_light* light);                         /* Scene::AddLight */
int nx_scene_set_light(nx_scene* scene, uint32_t idx, const nx_light* light);           /* GetLights()[idx] = ...; InvalidateLight(idx) */
int nx_scene_remove_light(nx_scene* scene, uint32_t idx);                               /* Scene::RemoveLight */
int nx_scene_light_count(nx_scene* scene);                                              /* lights added through AddLight (mesh lights are automatic) */
/* Camera::OnResize: the output resolution of the scene's camera (must equal the renderer's at render time). */
int nx_scene_set_resolution(nx_scene* s, uint32_t width, uint32_t height);
int nx_scene_set_camera(nx_scene* scene, const nx_camera* cam);
int nx_scene_set_render_settings(nx_scene* scene, const nx_render_settings* rs);
/* AssetManager::AddTexture + Texture::ToDevice (src/Assets/Texture.cpp:12-46): HOST RGBA8 (is_hdr 0; srgb: decode in the sampler) or RGBA32F
 * (is_hdr 1) pixels, wrap addressing, linear filtering.  Returns the index nx_material::*_map refers to, or < 0. */
int nx_scene_add_texture(nx_scene* s, const void* host_rgba, uint32_t w, uint32_t h, int is_hdr, int srgb);
int nx_scene_set_hdr_map(nx_scene* scene, const float* rgba, uint32_t w, uint32_t h);   /* Scene::AddHDRMap (RGBA32F equirect) */
/* Scene::Update: uploads dirty instances/materials, rebuilds the TLAS (BuildBVH8<AABB>, default config) and maintains the
 * emissive-mesh light list (Scene::UpdateSceneLighting, src/Scene/Scene.cpp:157-219). */
int nx_scene_update(nx_scene* scene);
/* Device-layout mirrors for parity with the reference arm (160-byte D_MeshInstance, 88-byte D_Camera, 52-byte D_Light). */
int nx_scene_export_instances(nx_scene* scene, void* out160 /* n*160 */, uint32_t* out_count);
int nx_scene_export_camera(nx_scene* scene, void* out88);
int nx_scene_export_lights(nx_scene* scene, void* out52 /* n*52 */, uint32_t* out_count);
/* Parity hooks for the merged BLAS (nx_ctx_set_instance_merging).  The TLAS is built over ENTRIES - the instances that keep a BLAS of
 * their own, then the merged BLAS - and its prim_idx holds entry numbers: out_inst[e] = instance id of entry e, 0xffffffff = merged. */
int nx_scene_export_tlas_entries(nx_scene* scene, uint32_t* out_inst, uint32_t* out_count);
/* Handle of the merged BLAS and, per merged primitive, the padded world-space box it was built over (6 floats: min, max), its instance
 * and its primitive id inside that instance's mesh.  The tree's nodes are in world space; its triangles are tested in the object space
 * of their instance, so the hits are the two-level scene's.  out_count = 0 when the scene has no merged BLAS.  Any output pointer may be null. */
int nx_scene_export_merged(nx_scene* scene, nx_bvh8* out_bvh, float* host_bounds, uint32_t* host_inst, uint32_t* host_prim, uint32_t* out_count);
int nx_scene_tlas(nx_scene* scene, nx_bvh8* out);                                       /* borrowed handle */
int nx_scene_tlas_history(nx_scene* scene, uint32_t* out_builds, uint32_t* out_refits);  /* how often the TLAS was built / refitted */
/* The same records computed without a scene or a GPU (pure host arithmetic, the code paths nx_scene_add_instance and
 * nx_scene_export_camera use): MeshInstance::ToDevice (src/Scene/MeshInstance.h:36-66) and Camera::ToDevice (src/Scene/Camera.cpp:130-156). */
int nx_host_instance_record(const float position[3], const float rotation_deg[3], const float scale[3], const nx_aabb* mesh_bounds,
                            uint32_t mesh_idx, uint32_t material_idx, void* out160);
int nx_host_camera_record(const nx_camera* cam, uint32_t width, uint32_t height, void* out88);

/* Parity hook: closest hits of a ray batch through the product traversal kernel (HOST buffers). */
int nx_trace_closest(nx_scene* scene, const nx_ray* rays, uint32_t n, nx_hit* hits, float* out_device_ms);
/* Any-hit (shadow) variant: out_occluded[i] = 1 if anything lies in (0, tmax). */
int nx_trace_any(nx_scene* scene, const nx_ray* rays, uint32_t n, uint8_t* out_occluded, float* out_device_ms);
/* Same on DEVICE buffers (rays/hits already resident), used by bench.py for the kernel-only number. */
int nx_trace_closest_device(nx_scene* scene, const nx_ray* d_rays, uint32_t n, nx_hit* d_hits, float* out_device_ms);

/* Traversal work counters over a DEVICE ray batch: out4 = {nodes visited, triangles tested, instances entered, rays}.  These are
 * the per-scene averages SURVEY.md 8(d)'s algorithmic-byte formula needs. */
int nx_trace_stats(nx_scene* scene, const nx_ray* d_rays, uint32_t n, nx_hit* d_hits, uint64_t out4[4]);

/* ----------------------------------------------------------------------------------------------- renderer ---- */
int nx_renderer_create(nx_ctx* ctx, uint32_t width, uint32_t height, nx_renderer** out);   /* PathTracer(uint2) + Reset  */
void nx_renderer_destroy(nx_renderer* r);
int nx_renderer_resize(nx_renderer* r, uint32_t width, uint32_t height);                   /* PathTracer::OnResize       */
int nx_renderer_reset_accumulation(nx_renderer* r);                                        /* ResetFrameNumber           */
/* Renders frames first_frame .. first_frame + n - 1 (frame numbers start at 1 and seed the RNG, N/Cuda/Random.cuh:67-78)
 * and adds them to the accumulation.  Asynchronous like PathTracer::Render; nx_renderer_stats() synchronises. */
int nx_renderer_render(nx_renderer* r, nx_scene* scene, uint32_t first_frame, uint32_t n_frames);
int nx_renderer_frame_count(const nx_renderer* r);                                         /* GetFrameNumber             */
int nx_renderer_stats(nx_renderer* r, nx_frame_stats* out);
/* Measurement hooks (the reference's only instrumentation is BVHBuildMetrics; MetricsPanel.cpp:22-40 counts primary rays).
 * flags & 1: bracket every kernel launch of nx_renderer_render with CUDA events on its stream; flags & 2: run the counting
 * variants of the two traversal kernels (nodes / triangles / instances per ray for SURVEY.md 8(d)'s byte formula). */
int nx_renderer_set_profiling(nx_renderer* r, int flags);
int nx_renderer_profile(nx_renderer* r, nx_kernel_profile* out);
/* Linear radiance mean, float RGB, row-major, HOST buffer of w*h*3 floats. */
int nx_renderer_read_accum(nx_renderer* r, float* host_rgb);
/* DEVICE pointer to the running float3 SUM (w*h*3 floats) and the number of frames in it — what the NCCL reduce consumes. */
int nx_renderer_accum_device(nx_renderer* r, float** out_dev_sum, uint32_t* out_frames);
int nx_renderer_set_accum_frames(nx_renderer* r, uint32_t frames);                         /* after an external all-reduce */
/* Tone-mapped RGBA8 (AccumulateKernel's display transform, PathTracer.cu:527-548), HOST buffer of w*h uint32. */
int nx_renderer_read_rgba8(nx_renderer* r, nx_scene* scene, uint32_t* host_rgba);
/* Pipelined display read-back, the headless counterpart of the reference's pixel-buffer path (PathTracer::Render maps a GL PBO
 * and returns without synchronising, src/Renderer/PathTracer.cpp:170-199; Renderer::UnpackToTexture consumes it a frame later,
 * src/Renderer/Renderer.cpp:41-48).  nx_renderer_present resolves the current accumulation to RGBA8 on the render stream into
 * one of two device images and queues its copy to `host_rgba` (w*h uint32, PINNED host memory for a truly asynchronous copy) on
 * a copy stream, together with the queue totals of the last nx_renderer_render call; it returns at once with a ticket
 * (0 or 1, alternating).  nx_renderer_present_wait blocks until that ticket's image and totals are in host memory.  A slot is
 * reused two presents later: wait for a ticket before presenting twice more into the same host buffer. */
int nx_renderer_present(nx_renderer* r, nx_scene* scene, uint32_t* host_rgba, int* out_ticket);
/* The same resolve straight into caller-supplied DEVICE memory (W * H RGBA8 words): the mapped pixel-buffer object of an OpenGL viewer
 * (PixelBuffer::GetDevicePtr, src/OpenGL/PixelBuffer.cpp:4-40; Renderer::Render maps it, src/Renderer/Renderer.cpp:41-48).  Queued on
 * nx_ctx_stream() behind the frame, returns at once; synchronise that stream before unmapping the buffer. */
int nx_renderer_present_device(nx_renderer* r, nx_scene* scene, uint32_t* dev_rgba);
int nx_renderer_present_wait(nx_renderer* r, int ticket, nx_frame_stats* out_stats /* may be NULL; device_ms is 0 */);
/* PathTracer::SetPixelQuery / PixelQueryPending / SynchronizePixelQuery (src/Renderer/PathTracer.h:23-27, PathTracer.cpp:221-240;
 * device side PathTracer.cu:150-151, 459-460): the instance under pixel (x, y) — row 0 is the bottom row — as seen by the
 * primary ray of the next rendered frame; -1 when that ray leaves the scene.  Synchronise returns the instance id and clears the
 * pending flag; without a rendered frame in between it returns the previous answer (-1 initially), like the reference. */
int nx_renderer_set_pixel_query(nx_renderer* r, uint32_t x, uint32_t y);
int nx_renderer_pixel_query_pending(const nx_renderer* r);
int nx_renderer_sync_pixel_query(nx_renderer* r, int32_t* out_instance);
/* The same display transform (exposure, tone curve NX_TONE_*, gamma 2.2, RGBA8 pack: src/Utils/ColorUtils.h:27-212) applied on
 * the device to a HOST linear float RGB image of `count` pixels; for previews of EXR/PFM output and for the parity tests. */
int nx_display_transform(nx_ctx* ctx, const float* host_rgb, uint32_t count, int tone_mapping, float exposure, uint32_t* host_rgba);
/* Headless output (north_star): PFM (little-endian float RGB) and EXR (uncompressed scanline, float RGB).  `rgb` is in the renderer's
 * layout - row 0 = BOTTOM row, what nx_renderer_read_accum returns - and both files come out upright. */
int nx_write_pfm(const char* path, const float* rgb, uint32_t w, uint32_t h);
int nx_write_exr(const char* path, const float* rgb, uint32_t w, uint32_t h);

#ifdef __cplusplus
}
#endif
#endif /* NEXUS_B200_H */
