// Wavefront state shared by the two translation units of the renderer: render.cu (traversal kernels + host driver, IEEE
// arithmetic) and shade.cu (ray generation, shading, display transform; compiled with --use_fast_math like the reference).
#pragma once
#include "scene.cuh"
#include "bsdf.cuh"

constexpr int kShadeBlock = 128;
#ifndef NX_TILED_PIXELS
#define NX_TILED_PIXELS 1
#endif
#ifndef NX_SHADE_MIN_BLOCKS
#define NX_SHADE_MIN_BLOCKS 5
#endif
constexpr uint32_t kMaxBounce = 256;

struct WaveCounters {                 // zeroed at the start of every frame
    uint32_t extCount[kMaxBounce];    // extension rays queued for bounce b
    uint32_t extFetch[kMaxBounce];    // persistent-thread fetch cursor of the bounce-b trace
    uint32_t shCount[kMaxBounce];     // shadow rays queued while shading bounce b
    uint32_t shFetch[kMaxBounce];
    uint32_t shaded[kMaxBounce];      // surviving hits shaded at bounce b
};
struct WaveTotals { unsigned long long ext, shadow, shaded, frames; };

struct WaveBuffers {
    nx_ray* ext[2];        // extension-ray queues (ping-pong); nx_ray::pad carries the pixel index
    float4* state[2];      // (throughput.rgb, last bsdf pdf) of the path that owns the ray
    nx_hit* hits;          // closest hits, same index as the traced queue
    nx_ray* shadow[2];     // shadow rays of bounce b in shadow[b & 1]; tmax = distance to the light sample, pad = pixel index
    float4* shadowRad[2];  // radiance to add when the shadow ray is unoccluded
    float* accum;          // running SUM of radiance, 3 floats per pixel, row 0 = bottom row like the reference
    WaveCounters* counters;
    WaveTotals* totals;
};

// Primary-ray queue order (generate_kernel): 8x4 pixel tiles when the resolution allows, row-major otherwise.
__host__ __device__ __forceinline__ bool pixels_tiled(uint32_t resX, uint32_t resY) { return NX_TILED_PIXELS && (resX & 7u) == 0u && (resY & 3u) == 0u; }
__host__ __device__ __forceinline__ void slot_to_pixel(uint32_t i, uint32_t resX, uint32_t resY, uint32_t& px, uint32_t& py)
{
    if (pixels_tiled(resX, resY)) {
        const uint32_t tile = i >> 5, tilesX = resX >> 3, ty = tile / tilesX, tx = tile - ty * tilesX;
        px = tx * 8u + (i & 7u); py = ty * 4u + ((i >> 3) & 3u);
    } else { py = i / resX; px = i - py * resX; }
}
__host__ __device__ __forceinline__ uint32_t pixel_to_slot(uint32_t pixel, uint32_t resX, uint32_t resY)
{
    if (!pixels_tiled(resX, resY)) return pixel;
    const uint32_t py = pixel / resX, px = pixel - py * resX;
    return (((py >> 2) * (resX >> 3) + (px >> 3)) << 5) | ((py & 3u) << 3) | (px & 7u);
}

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// one atomic per warp; every lane of the warp must call this (convergent point)
__device__ __forceinline__ uint32_t warp_append(uint32_t* counter, bool want)
{
    const uint32_t mask = __ballot_sync(NX_FULL, want);
    if (!mask) return 0;
    const uint32_t leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane_id() == leader) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(NX_FULL, base, leader);
    return base + __popc(mask & lanemask_lt());
}

__device__ __forceinline__ void add_radiance(float* accum, uint32_t pixel, F3 L)
{
    if (L.x != 0.f) atomicAdd(accum + 3 * (size_t)pixel, L.x);
    if (L.y != 0.f) atomicAdd(accum + 3 * (size_t)pixel + 1, L.y);
    if (L.z != 0.f) atomicAdd(accum + 3 * (size_t)pixel + 2, L.z);
}

// launchers of the kernels in shade.cu
int nxi_shade_grid(nx_ctx* ctx);
void nxi_launch_generate(const DSceneView& sv, const WaveBuffers& wb, uint32_t frame, int grid, cudaStream_t s);
void nxi_launch_shade(const DSceneView& sv, const WaveBuffers& wb, uint32_t bounce, uint32_t frame, int grid, cudaStream_t s);
void nxi_launch_resolve(int grid, cudaStream_t s, const float* accum, uint32_t count, float invFrames, float exposure, int mode, uint32_t* out);
