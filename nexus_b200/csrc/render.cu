// Wavefront path tracer for sm_100a: generate -> trace -> shade (logic + OpenPBR material + NEE/MIS) -> {trace, trace shadow}
// -> accumulate.  Replaces the six kernels of src/Cuda/PathTracer/PathTracer.cu:60-549 and PathTracer::Render
// (src/Renderer/PathTracer.cpp:166-200) of the reference.
//
// Differences that matter for throughput:
//   * Logic and Material are one kernel: the miss / Russian-roulette decisions are taken in registers right before shading
//     instead of through a 36 B/ray material queue round trip;
//   * path throughput and last pdf travel with the ray (one 16-byte record in queue order) instead of being gathered and
//     scattered through per-pixel arrays;
//   * queue appends are one atomic per warp (ballot + popc prefix) at a convergent point;
//   * radiance goes straight into the float sum buffer with RED.ADD (no per-frame radiance buffer, no accumulate pass over
//     W*H pixels per frame); the mean is taken on read-out, which is also what the multi-GPU reduction wants;
//   * all kernels are persistent and sized from the SM count; queue lengths are read on the device, so no kernel is
//     launched over W*H threads for a nearly empty queue and the host never waits inside a frame;
//   * rays are 32-byte records, hits 20-byte records: 16-byte vector loads in queue order.
#include "scene.cuh"
#include "bsdf.cuh"

namespace {

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
    nx_ray* shadow;        // shadow rays; tmax = distance to the light sample, pad = pixel index
    float4* shadowRad;     // radiance to add when the shadow ray is unoccluded
    float* accum;          // running SUM of radiance, 3 floats per pixel, row 0 = bottom row like the reference
    WaveCounters* counters;
    WaveTotals* totals;
};

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

// ------------------------------------------------------------------------------------------ trace kernels ----
// Persistent warps over a device-side queue; the loop itself is trace_loop() in traverse.cuh.
struct ClosestSink {
    nx_hit* hits;
    __device__ __forceinline__ void finish(const TraceScene& sc, uint32_t rayIdx, uint32_t, float t, float u, float v, uint32_t prim, uint32_t slot, bool)
    {
        nx_hit h; h.t = t; h.u = u; h.v = v; h.prim = prim;
        h.instance = slot != NX_INVALID ? __ldg(sc.tlasPrimIdx + slot) : NX_INVALID;
        hits[rayIdx] = h;
    }
};
// Any-hit.  occluded != null: write occlusion flags (parity hook); else add the queued radiance to the pixel when unoccluded
// (TraceShadowKernel's fused accumulate, PathTracer.cu:115-122 / BVH8Traversal.cuh:517-519).
struct AnySink {
    uint8_t* occluded; const float4* radiance; float* accum;
    __device__ __forceinline__ void finish(const TraceScene&, uint32_t rayIdx, uint32_t pixel, float, float, float, uint32_t, uint32_t, bool occ)
    {
        if (occluded) occluded[rayIdx] = occ ? 1 : 0;
        else if (!occ) { const float4 L = __ldg(radiance + rayIdx); add_radiance(accum, pixel, f3(L.x, L.y, L.z)); }
    }
};

template <bool STATS>
__global__ void __launch_bounds__(NX_TRACE_BLOCK, NX_TRACE_MIN_BLOCKS) trace_closest_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                         uint32_t* cursor, nx_hit* __restrict__ hits, TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[NX_TRACE_SMEM_BYTES / 4];
    ClosestSink sink{hits};
    trace_loop<false, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}

template <bool STATS>
__global__ void __launch_bounds__(NX_TRACE_BLOCK, NX_TRACE_MIN_BLOCKS) trace_any_kernel(TraceScene sc, const nx_ray* __restrict__ rays, uint32_t nImm, const uint32_t* nPtr,
                                                                     uint32_t* cursor, uint8_t* occluded, const float4* __restrict__ radiance, float* accum,
                                                                     TraceStats* stats, TraceTuning tune)
{
    __shared__ __align__(16) uint32_t smem[NX_TRACE_SMEM_BYTES / 4];
    AnySink sink{occluded, radiance, accum};
    trace_loop<true, STATS>(sc, rays, nPtr ? __ldg(nPtr) : nImm, cursor, tune, smem, sink, stats);
}

// --------------------------------------------------------------------------------------------- generate ----
// Camera rays with pixel jitter and thin-lens sampling (GenerateKernel, PathTracer.cu:60-95).
__global__ void __launch_bounds__(256) generate_kernel(const __grid_constant__ DSceneView sv, WaveBuffers wb, uint32_t frame)
{
    const DCamera& cam = sv.camera;
    const uint32_t count = cam.resX * cam.resY;
    if (blockIdx.x == 0 && threadIdx.x == 0) wb.counters->extCount[0] = count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        // Queue slot i -> pixel: 8x4 pixel tiles, so the 32 rays a warp fetches together cover a compact footprint (and, because
        // every later queue inherits this order through compaction, so do their bounces).  Row-major order when the
        // resolution is not a multiple of the tile.  The pixel index, not the slot, keys the RNG and addresses the image.
        uint32_t px, py;
        if (NX_TILED_PIXELS && (cam.resX & 7u) == 0u && (cam.resY & 3u) == 0u) {
            const uint32_t tile = i >> 5, tilesX = cam.resX >> 3, ty = tile / tilesX, tx = tile - ty * tilesX;
            px = tx * 8u + (i & 7u); py = ty * 4u + ((i >> 3) & 3u);
        } else { py = i / cam.resX; px = i - py * cam.resX; }
        const uint32_t pixel = py * cam.resX + px;
        uint32_t rng = rng_seed(pixel, frame, 0u);
        const float x = ((float)px + rng_next(rng)) / (float)cam.resX;
        const float y = ((float)py + rng_next(rng)) / (float)cam.resY;
        const float u0 = rng_next(rng), u1 = rng_next(rng);   // concentric-free polar disk sample (Random.cuh:100-107)
        float sn, cs; __sincosf(NX_TWO_PI * u1, &sn, &cs);
        const float r = cam.lensRadius * sqrtf(u0);
        const F3 right = f3(cam.right[0], cam.right[1], cam.right[2]), up = f3(cam.up[0], cam.up[1], cam.up[2]);
        const F3 off = right * (r * cs) + up * (r * sn);
        const F3 pos = f3(cam.position[0], cam.position[1], cam.position[2]);
        const F3 org = pos + off;
        const F3 target = f3(cam.lowerLeft[0], cam.lowerLeft[1], cam.lowerLeft[2]) + x * f3(cam.viewportX[0], cam.viewportX[1], cam.viewportX[2]) +
                          y * f3(cam.viewportY[0], cam.viewportY[1], cam.viewportY[2]);
        const F3 dir = normalize(target - pos - off);
        float4* out = reinterpret_cast<float4*>(wb.ext[0] + i);
        out[0] = make_float4(org.x, org.y, org.z, NX_MISS_T);
        out[1] = make_float4(dir.x, dir.y, dir.z, __uint_as_float(pixel));
        wb.state[0][i] = make_float4(1.f, 1.f, 1.f, 0.f);
    }
}

// ------------------------------------------------------------------------------------------------ shade ----
__device__ __forceinline__ F3 load3(const float* p) { return f3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ F3 bary(F3 a, F3 b, F3 c, float u, float v) { return u * b + v * c + (1.0f - u - v) * a; }   // Utils.cuh:59-63
__device__ __forceinline__ F3 xf_point(float4 r0, float4 r1, float4 r2, F3 p)
{
    return f3(r0.x * p.x + r0.y * p.y + r0.z * p.z + r0.w, r1.x * p.x + r1.y * p.y + r1.z * p.z + r1.w, r2.x * p.x + r2.y * p.y + r2.z * p.z + r2.w);
}
// (M^-1)^T * n: normals transform with the transposed inverse (PathTracer.cu:385-389)
__device__ __forceinline__ F3 xf_normal(float4 i0, float4 i1, float4 i2, F3 n)
{
    return f3(i0.x * n.x + i1.x * n.y + i2.x * n.z, i0.y * n.x + i1.y * n.y + i2.y * n.z, i0.z * n.x + i1.z * n.y + i2.z * n.z);
}
// "A Fast and Robust Method for Avoiding Self-Intersection" (Ray Tracing Gems ch. 6; src/Cuda/Utils.cuh:65-86)
__device__ __forceinline__ float offset_axis(float p, float n)
{
    const int of = (int)(256.0f * n);
    const float pi = __int_as_float(__float_as_int(p) + (p < 0.0f ? -of : of));
    return fabsf(p) < (1.0f / 32.0f) ? p + (1.0f / 65536.0f) * n : pi;
}
__device__ __forceinline__ F3 offset_ray(F3 p, F3 n) { return f3(offset_axis(p.x, n.x), offset_axis(p.y, n.y), offset_axis(p.z, n.z)); }
__device__ __forceinline__ float power_heuristic(float a, float b) { return a * a / (a * a + b * b); }   // Sampler.cuh:22-25

__device__ __forceinline__ F3 background(const DSceneView& sv, F3 d)   // SampleBackground, PathTracer.cu:40-58
{
    if (sv.hasHdr) {
        const float theta = atan2f(d.z, d.x), phi = asinf(d.y);
        const float u = (theta + NX_PI) * NX_INV_PI * 0.5f, v = 1.0f - (phi + NX_PI * 0.5f) * NX_INV_PI;
        const float4 c = tex2D<float4>(sv.hdr, u, v);
        return f3(c.x, c.y, c.z) * sv.bgIntensity;
    }
    return f3(sv.bg[0], sv.bg[1], sv.bg[2]) * sv.bgIntensity;
}

struct Surface { F3 p, n, gn; };

struct ShadeOut {
    bool ext, shadow;
    F3 extO, extD, thr; float pdf;
    F3 shO, shD, shL; float shDist;
};

// Next-event estimation: one light picked uniformly, one point on it, MIS against the BSDF (PathTracer.cu:176-343).
// The light-specific part only produces (direction, distance, pdf, emission); the BSDF is evaluated once, at one call site,
// which keeps the kernel's code size (and with it the instruction-cache pressure ncu showed) down.
__device__ __forceinline__ void next_event(const DSceneView& sv, const nx_material& mat, const Surface& sf, const Frame& fr, F3 wi, F3 rayDir, F3 thr,
                                           uint32_t& rng, ShadeOut& out)
{
    const uint32_t li = (uint32_t)floorf(rng_next(rng) * (float)sv.lightCount);
    const DLight L = sv.lights[min(li, sv.lightCount - 1u)];
    F3 toLight, emissive, dir, origin; float lightPdf, dist; bool mis = false;
    if (L.type == NX_LIGHT_MESH)
    {
        const DShadeInst I = sv.shadeInst[L.instance];
        const DMesh mesh = sv.meshes[I.meshIdx];
        const uint32_t ti = min((uint32_t)floorf(rng_next(rng) * (float)mesh.primCount), mesh.primCount - 1u);
        const float a = rng_next(rng), b = rng_next(rng), su = sqrtf(a);
        const float u = 1.0f - su, v = b * su;                                   // uniform triangle sample (Sampler.cuh:41-48)
        const float* t = mesh.tris + 9 * (size_t)ti; const float* td = mesh.tridata + 24 * (size_t)ti;
        const F3 v0 = load3(t), v1 = load3(t + 3), v2 = load3(t + 6);
        F3 lp = xf_point(I.m0, I.m1, I.m2, bary(v0, v1, v2, u, v));
        const F3 lgn = normalize(xf_normal(I.i0, I.i1, I.i2, cross(v1 - v0, v2 - v0)));
        const F3 ln = normalize(xf_normal(I.i0, I.i1, I.i2, bary(load3(td), load3(td + 3), load3(td + 6), u, v)));
        toLight = lp - sf.p;
        const bool sameSide = dot(-rayDir, sf.gn) * dot(toLight, sf.gn) > 0.0f;
        if (!sameSide && mat.transmission == 0.0f) return;
        origin = offset_ray(sf.p, sf.gn * sign_or_one(dot(toLight, sf.n)));
        lp = offset_ray(lp, lgn * sign_or_one(dot(-toLight, ln)));
        const F3 seg = lp - origin;
        dist = length(seg); dir = seg / dist;
        const float cosL = fabsf(dot(ln, dir));
        const F3 w0 = xf_point(I.m0, I.m1, I.m2, v0), w1 = xf_point(I.m0, I.m1, I.m2, v1), w2 = xf_point(I.m0, I.m1, I.m2, v2);
        const float area = 0.5f * length(cross(w1 - w0, w2 - w0));
        lightPdf = 1.0f / ((float)sv.lightCount * (float)mesh.primCount * area);
        lightPdf *= dot(toLight, toLight) / cosL;                                // area measure -> solid angle
        if (!pdf_ok(lightPdf)) return;
        const nx_material& lm = sv.materials[I.materialIdx];
        emissive = f3(__ldg(&lm.emission_color[0]), __ldg(&lm.emission_color[1]), __ldg(&lm.emission_color[2])) * __ldg(&lm.intensity);
        mis = true;
    }
    else if (L.type == NX_LIGHT_POINT || L.type == NX_LIGHT_DIRECTIONAL)
    {
        const bool point = L.type == NX_LIGHT_POINT;
        toLight = point ? f3(L.px, L.py, L.pz) - sf.p : -f3(L.dx, L.dy, L.dz);
        lightPdf = 1.0f / (float)sv.lightCount;
        if (point) { lightPdf *= dot(toLight, toLight); if (!pdf_ok(lightPdf)) return; }
        emissive = f3(L.cr, L.cg, L.cb) * L.intensity;
        const bool sameSide = dot(-rayDir, sf.gn) * dot(toLight, sf.gn) > 0.0f;
        if (!sameSide && mat.transmission == 0.0f) return;
        origin = offset_ray(sf.p, sf.gn * sign_or_one(dot(toLight, sf.n)));
        dist = point ? length(toLight) : NX_MISS_T;
        dir = point ? toLight / dist : normalize(toLight);
    }
    else return;   // spot lights are declared but have no NEE branch in the reference either (PathTracer.cu:274-334)
    F3 f; float bsdfPdf;
    if (!principled_eval(mat, wi, fr.toLocal(dir), f, bsdfPdf)) return;
    const float weight = mis ? power_heuristic(lightPdf, bsdfPdf) : 1.0f;
    out.shL = weight * thr * f * emissive / lightPdf;
    out.shadow = true; out.shO = origin; out.shD = dir; out.shDist = dist;
}

// LogicKernel + MaterialKernel for one traced ray (PathTracer.cu:124-173, 346-511).
__device__ __forceinline__ void shade_one(const DSceneView& sv, const WaveBuffers& wb, uint32_t bounce, uint32_t frame, const nx_hit& hit, F3 rayDir,
                                          uint32_t pixel, F3 thr, float lastPdf, ShadeOut& out, bool& survived)
{
    if (hit.t == NX_MISS_T) { add_radiance(wb.accum, pixel, thr * background(sv, rayDir)); return; }

    uint32_t rng = rng_seed(pixel, frame, bounce);
    // Russian roulette on the largest throughput component, from the first bounce, no clamp (PathTracer.cu:158-166)
    const float survive = max3(thr);
    if (!(rng_next(rng) < survive)) return;
    thr = thr / survive;
    survived = true;

    const DShadeInst I = sv.shadeInst[hit.instance];
    const DMesh mesh = sv.meshes[I.meshIdx];
    const float* t = mesh.tris + 9 * (size_t)hit.prim; const float* td = mesh.tridata + 24 * (size_t)hit.prim;
    const F3 v0 = load3(t), v1 = load3(t + 3), v2 = load3(t + 6);
    const nx_material mat = sv.materials[I.materialIdx];

    Surface sf;
    sf.p = xf_point(I.m0, I.m1, I.m2, bary(v0, v1, v2, hit.u, hit.v));
    sf.n = normalize(xf_normal(I.i0, I.i1, I.i2, normalize(bary(load3(td), load3(td + 3), load3(td + 6), hit.u, hit.v))));
    sf.gn = normalize(xf_normal(I.i0, I.i1, I.i2, cross(v1 - v0, v2 - v0)));
    const Frame fr(sf.n);

    // emission seen by the BSDF-sampled ray, MIS-weighted against light sampling except on primary hits (PathTracer.cu:414-447)
    const F3 Le = f3(mat.emission_color[0], mat.emission_color[1], mat.emission_color[2]) * mat.intensity;
    if (max3(Le) > 0.0f)
    {
        float w = 1.0f;
        if (bounce > 1u && sv.useMIS) {
            const float cosL = fabsf(dot(sf.n, rayDir));
            const F3 w0 = xf_point(I.m0, I.m1, I.m2, v0), w1 = xf_point(I.m0, I.m1, I.m2, v1), w2 = xf_point(I.m0, I.m1, I.m2, v2);
            const float area = 0.5f * length(cross(w1 - w0, w2 - w0));
            float lightPdf = 1.0f / ((float)sv.lightCount * (float)mesh.primCount * area);
            lightPdf *= sqr(hit.t) / cosL;
            w = pdf_ok(lightPdf) ? power_heuristic(lastPdf, lightPdf) : 0.0f;
        }
        add_radiance(wb.accum, pixel, w * Le * thr);
    }
    if (bounce == sv.pathLength) return;

    const F3 wi = fr.toLocal(-rayDir);
    if (rng_next(rng) > mat.opacity)
    {
        // alpha pass-through: continue straight on, path state unchanged (PathTracer.cu:464-475)
        const F3 wo = fr.toWorld(-wi);
        out.ext = true; out.extO = offset_ray(sf.p, sf.gn * sign_or_one(dot(wo, sf.n))); out.extD = wo; out.thr = thr; out.pdf = lastPdf;
        return;
    }
    if (sv.useMIS && sv.lightCount > 0u) next_event(sv, mat, sf, fr, wi, rayDir, thr, rng, out);

    const LobeSample s = principled_sample(mat, wi, rng);
    if (!s.ok) return;
    const F3 wo = fr.toWorld(s.wo);
    const bool sameSide = dot(-rayDir, sf.gn) * dot(wo, sf.gn) > 0.0f;
    if (!sameSide && mat.transmission == 0.0f) return;
    out.ext = true; out.extO = offset_ray(sf.p, sf.gn * sign_or_one(dot(wo, sf.n))); out.extD = wo; out.thr = thr * s.weight; out.pdf = s.pdf;
}

__global__ void __launch_bounds__(kShadeBlock, NX_SHADE_MIN_BLOCKS) shade_kernel(const __grid_constant__ DSceneView sv, WaveBuffers wb, uint32_t bounce, uint32_t frame)
{
    const uint32_t n = wb.counters->extCount[bounce - 1];
    const uint32_t in = (bounce - 1) & 1u, outQ = bounce & 1u;
    uint32_t shadedHere = 0;
    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
    {
        const uint32_t i = base + threadIdx.x;
        ShadeOut o; o.ext = false; o.shadow = false;
        uint32_t pixel = 0;
        if (i < n)
        {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(wb.ext[in] + i) + 1);
            const float4 st = __ldg(wb.state[in] + i);
            const nx_hit h = wb.hits[i];
            pixel = __float_as_uint(d4.w);
            bool survived = false;
            shade_one(sv, wb, bounce, frame, h, f3(d4.x, d4.y, d4.z), pixel, f3(st.x, st.y, st.z), st.w, o, survived);
            shadedHere += survived ? 1u : 0u;
        }
        const uint32_t e = warp_append(&wb.counters->extCount[bounce], o.ext);
        if (o.ext) {
            float4* r = reinterpret_cast<float4*>(wb.ext[outQ] + e);
            r[0] = make_float4(o.extO.x, o.extO.y, o.extO.z, NX_MISS_T);
            r[1] = make_float4(o.extD.x, o.extD.y, o.extD.z, __uint_as_float(pixel));
            wb.state[outQ][e] = make_float4(o.thr.x, o.thr.y, o.thr.z, o.pdf);
        }
        const uint32_t s = warp_append(&wb.counters->shCount[bounce], o.shadow);
        if (o.shadow) {
            float4* r = reinterpret_cast<float4*>(wb.shadow + s);
            r[0] = make_float4(o.shO.x, o.shO.y, o.shO.z, o.shDist);
            r[1] = make_float4(o.shD.x, o.shD.y, o.shD.z, __uint_as_float(pixel));
            wb.shadowRad[s] = make_float4(o.shL.x, o.shL.y, o.shL.z, 0.f);
        }
    }
    for (int off = 16; off > 0; off >>= 1) shadedHere += __shfl_xor_sync(NX_FULL, shadedHere, off);
    if (lane_id() == 0 && shadedHere) atomicAdd(&wb.counters->shaded[bounce], shadedHere);
}

__global__ void frame_totals_kernel(WaveBuffers wb, uint32_t pathLength)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long e = 0, s = 0, h = 0;
    for (uint32_t b = 0; b <= pathLength && b < kMaxBounce; b++) { e += wb.counters->extCount[b]; s += wb.counters->shCount[b]; h += wb.counters->shaded[b]; }
    wb.totals->ext += e; wb.totals->shadow += s; wb.totals->shaded += h; wb.totals->frames += 1;
}

// Display transform of AccumulateKernel (PathTracer.cu:527-548) on the mean; tone curves other than NONE are added with
// SURVEY.md §8 row f-1, until then every mode maps to exposure + gamma 2.2.
__global__ void resolve_rgba8_kernel(const float* __restrict__ accum, uint32_t count, float invFrames, float exposure, uint32_t* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const float s = invFrames * exp2f(exposure);
        const float r = __powf(fmaxf(accum[3 * (size_t)i] * s, 0.f), 1.0f / 2.2f), g = __powf(fmaxf(accum[3 * (size_t)i + 1] * s, 0.f), 1.0f / 2.2f),
                    b = __powf(fmaxf(accum[3 * (size_t)i + 2] * s, 0.f), 1.0f / 2.2f);
        out[i] = (uint32_t)(__saturatef(r) * 255.0f) | ((uint32_t)(__saturatef(g) * 255.0f) << 8) | ((uint32_t)(__saturatef(b) * 255.0f) << 16) | 0xff000000u;
    }
}

TraceTuning trace_tuning(const nx_ctx* ctx) { TraceTuning t; t.triLanes = ctx->tune_tri; t.instLanes = ctx->tune_inst; t.sphereCull = ctx->tune_sphere; t.k47 = 0x47000000u; return t; }

int persistent_grid(nx_ctx* ctx, const void* fn, int block, int* cache)
{
    if (*cache) return *cache;
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, block, 0);
    if (perSm < 1) perSm = 1;
    *cache = perSm * ctx->sm_count;
    return *cache;
}
int g_gridClosest = 0, g_gridClosestStats = 0, g_gridAny = 0, g_gridAnyStats = 0, g_gridShade = 0;

} // namespace

struct nx_renderer {
    nx_ctx* ctx = nullptr;
    uint32_t width = 0, height = 0;
    WaveBuffers wb{};
    uint32_t frames = 0;            // frames in the accumulation
    cudaEvent_t evStart = nullptr, evStop = nullptr, evShade = nullptr, evShadow = nullptr;
    bool timed = false;
    uint32_t launches = 0;
    uint32_t pathLengthLast = 0;
    // optional per-kernel profiling (nx_renderer_set_profiling): event pairs per launch, grouped by kernel
    int profFlags = 0;
    std::vector<cudaEvent_t> evPool; size_t evUsed = 0;
    std::vector<std::pair<int, size_t>> evLaunches;   // (kernel class, index of the start event; stop = index + 1)
    TraceStats* dWork = nullptr;                      // [0] closest-hit traversal work, [1] any-hit traversal work

    cudaEvent_t next_event() { if (evUsed == evPool.size()) { cudaEvent_t e; cudaEventCreate(&e); evPool.push_back(e); } return evPool[evUsed++]; }
    void prof_begin(int cls, cudaStream_t st) { if (profFlags & 1) { evLaunches.push_back({cls, evUsed}); cudaEventRecord(next_event(), st); next_event(); } }
    void prof_end(cudaStream_t st) { if (profFlags & 1) cudaEventRecord(evPool[evLaunches.back().second + 1], st); }
};

int nxi_trace_closest(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, nx_hit* dHits, float* outMs)
{
    DeviceGuard guard(ctx->device);
    uint32_t* cursor = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
    const int grid = persistent_grid(ctx, (const void*)trace_closest_kernel<false>, NX_TRACE_BLOCK, &g_gridClosest);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (outMs) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
    trace_closest_kernel<false><<<grid, NX_TRACE_BLOCK, 0, ctx->stream>>>(sc, dRays, n, nullptr, cursor, dHits, nullptr, trace_tuning(ctx));
    if (outMs) { cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(outMs, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    cudaFreeAsync(cursor, ctx->stream);
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NX_CUDA(ctx, cudaGetLastError());
    return NX_OK;
}

int nxi_trace_any(nx_ctx* ctx, const TraceScene& sc, const nx_ray* dRays, uint32_t n, uint8_t* dOcc, float* outMs)
{
    DeviceGuard guard(ctx->device);
    uint32_t* cursor = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream));
    const int grid = persistent_grid(ctx, (const void*)trace_any_kernel<false>, NX_TRACE_BLOCK, &g_gridAny);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (outMs) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream); }
    trace_any_kernel<false><<<grid, NX_TRACE_BLOCK, 0, ctx->stream>>>(sc, dRays, n, nullptr, cursor, dOcc, nullptr, nullptr, nullptr, trace_tuning(ctx));
    if (outMs) { cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(outMs, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); }
    cudaFreeAsync(cursor, ctx->stream);
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    NX_CUDA(ctx, cudaGetLastError());
    return NX_OK;
}

// Traversal work counters for the roofline's algorithmic-byte figure: nodes, triangles, instances per ray.
extern "C" int nx_trace_stats(nx_scene* s, const nx_ray* dRays, uint32_t n, nx_hit* dHits, uint64_t out4[4])
{
    if (!s || !dRays || !dHits || !out4) return NX_ERR_INVALID;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    DSceneView v; int rc = nxi_scene_view(s, &v); if (rc) return rc;
    uint32_t* cursor = nullptr; TraceStats* st = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&cursor, 4, ctx->stream)); NX_CUDA(ctx, cudaMallocAsync((void**)&st, sizeof(TraceStats), ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(cursor, 0, 4, ctx->stream)); NX_CUDA(ctx, cudaMemsetAsync(st, 0, sizeof(TraceStats), ctx->stream));
    const int grid = persistent_grid(ctx, (const void*)trace_closest_kernel<true>, NX_TRACE_BLOCK, &g_gridClosestStats);
    trace_closest_kernel<true><<<grid, NX_TRACE_BLOCK, 0, ctx->stream>>>(v.trace, dRays, n, nullptr, cursor, dHits, st, trace_tuning(ctx));
    TraceStats h{};
    NX_CUDA(ctx, cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out4[0] = h.nodes; out4[1] = h.tris; out4[2] = h.insts; out4[3] = h.rays;
    cudaFreeAsync(cursor, ctx->stream); cudaFreeAsync(st, ctx->stream);
    return NX_OK;
}

namespace {

int free_buffers(nx_renderer* r)
{
    nx_ctx* ctx = r->ctx;
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream_aux);
    WaveBuffers& w = r->wb;
    cudaFree(w.ext[0]); cudaFree(w.ext[1]); cudaFree(w.state[0]); cudaFree(w.state[1]); cudaFree(w.hits); cudaFree(w.shadow);
    cudaFree(w.shadowRad); cudaFree(w.accum); cudaFree(w.counters); cudaFree(w.totals);
    w = WaveBuffers{};
    return NX_OK;
}

// PathTracer::Reset (src/Renderer/PathTracer.cpp:61-159): 32+32+16+16+20+32+16 = 164 B of queues per pixel + 12 B accumulation
int alloc_buffers(nx_renderer* r, uint32_t w, uint32_t h)
{
    nx_ctx* ctx = r->ctx;
    const size_t px = (size_t)w * h;
    WaveBuffers& b = r->wb;
    NX_CUDA(ctx, cudaMalloc((void**)&b.ext[0], sizeof(nx_ray) * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.ext[1], sizeof(nx_ray) * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.state[0], 16 * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.state[1], 16 * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.hits, sizeof(nx_hit) * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.shadow, sizeof(nx_ray) * px)); NX_CUDA(ctx, cudaMalloc((void**)&b.shadowRad, 16 * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.accum, 12 * px));
    NX_CUDA(ctx, cudaMalloc((void**)&b.counters, sizeof(WaveCounters))); NX_CUDA(ctx, cudaMalloc((void**)&b.totals, sizeof(WaveTotals)));
    NX_CUDA(ctx, cudaMemsetAsync(b.accum, 0, 12 * px, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(b.totals, 0, sizeof(WaveTotals), ctx->stream));
    r->width = w; r->height = h; r->frames = 0;
    return NX_OK;
}

} // namespace

extern "C" {

int nx_renderer_create(nx_ctx* ctx, uint32_t width, uint32_t height, nx_renderer** out)
{
    if (!ctx || !out || !width || !height) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    nx_renderer* r = new nx_renderer(); r->ctx = ctx;
    int rc = alloc_buffers(r, width, height);
    if (rc) { delete r; return rc; }
    cudaEventCreate(&r->evStart); cudaEventCreate(&r->evStop);
    cudaEventCreateWithFlags(&r->evShade, cudaEventDisableTiming); cudaEventCreateWithFlags(&r->evShadow, cudaEventDisableTiming);
    if (cudaMalloc((void**)&r->dWork, 2 * sizeof(TraceStats)) != cudaSuccess || cudaMemset(r->dWork, 0, 2 * sizeof(TraceStats)) != cudaSuccess) {
        ctx->error = "nx_renderer_create: cudaMalloc failed"; nx_renderer_destroy(r); return NX_ERR_CUDA;
    }
    *out = r;
    return NX_OK;
}

void nx_renderer_destroy(nx_renderer* r)
{
    if (!r) return;
    DeviceGuard guard(r->ctx->device);
    free_buffers(r);
    cudaEventDestroy(r->evStart); cudaEventDestroy(r->evStop); cudaEventDestroy(r->evShade); cudaEventDestroy(r->evShadow);
    for (cudaEvent_t e : r->evPool) cudaEventDestroy(e);
    cudaFree(r->dWork);
    delete r;
}

int nx_renderer_resize(nx_renderer* r, uint32_t width, uint32_t height)
{
    if (!r || !width || !height) return NX_ERR_INVALID;
    if (width == r->width && height == r->height) return NX_OK;
    DeviceGuard guard(r->ctx->device);
    free_buffers(r);
    return alloc_buffers(r, width, height);
}

int nx_renderer_reset_accumulation(nx_renderer* r)
{
    if (!r) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemsetAsync(r->wb.accum, 0, 12 * (size_t)r->width * r->height, ctx->stream));
    NX_CUDA(ctx, cudaMemsetAsync(r->wb.totals, 0, sizeof(WaveTotals), ctx->stream));
    r->frames = 0; r->timed = false;
    return NX_OK;
}

int nx_renderer_render(nx_renderer* r, nx_scene* scene, uint32_t firstFrame, uint32_t nFrames)
{
    if (!r || !scene || scene->ctx != r->ctx) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    if (scene->width != r->width || scene->height != r->height) NX_FAIL(ctx, NX_ERR_INVALID, "scene resolution %ux%u != renderer %ux%u", scene->width, scene->height, r->width, r->height);
    DeviceGuard guard(ctx->device);
    DSceneView sv; int rc = nxi_scene_view(scene, &sv); if (rc) return rc;
    cudaStream_t s = ctx->stream, sa = ctx->stream_aux;
    const uint32_t L = sv.pathLength;
    const bool work = (r->profFlags & 2) != 0;
    const TraceTuning tune = trace_tuning(ctx);
    const int gClosest = work ? persistent_grid(ctx, (const void*)trace_closest_kernel<true>, NX_TRACE_BLOCK, &g_gridClosestStats)
                              : persistent_grid(ctx, (const void*)trace_closest_kernel<false>, NX_TRACE_BLOCK, &g_gridClosest);
    const int gAny = work ? persistent_grid(ctx, (const void*)trace_any_kernel<true>, NX_TRACE_BLOCK, &g_gridAnyStats)
                          : persistent_grid(ctx, (const void*)trace_any_kernel<false>, NX_TRACE_BLOCK, &g_gridAny);
    const int gShade = persistent_grid(ctx, (const void*)shade_kernel, kShadeBlock, &g_gridShade);
    const int gGen = ctx->sm_count * 8;
    WaveBuffers& wb = r->wb;
    r->evUsed = 0; r->evLaunches.clear();
    if (work) NX_CUDA(ctx, cudaMemsetAsync(r->dWork, 0, 2 * sizeof(TraceStats), s));

    auto closest = [&](const nx_ray* q, uint32_t b) {
        r->prof_begin(1, s);
        if (work) trace_closest_kernel<true><<<gClosest, NX_TRACE_BLOCK, 0, s>>>(sv.trace, q, 0, &wb.counters->extCount[b], &wb.counters->extFetch[b], wb.hits, r->dWork, tune);
        else trace_closest_kernel<false><<<gClosest, NX_TRACE_BLOCK, 0, s>>>(sv.trace, q, 0, &wb.counters->extCount[b], &wb.counters->extFetch[b], wb.hits, nullptr, tune);
        r->prof_end(s);
        r->launches++;
    };

    NX_CUDA(ctx, cudaMemsetAsync(wb.totals, 0, sizeof(WaveTotals), s));
    NX_CUDA(ctx, cudaEventRecord(r->evStart, s));
    r->launches = 0;
    for (uint32_t f = 0; f < nFrames; f++)
    {
        const uint32_t frame = firstFrame + f;
        NX_CUDA(ctx, cudaMemsetAsync(wb.counters, 0, sizeof(WaveCounters), s));
        r->prof_begin(0, s);
        generate_kernel<<<gGen, 256, 0, s>>>(sv, wb, frame);
        r->prof_end(s);
        r->launches++;
        closest(wb.ext[0], 0);
        for (uint32_t b = 1; b <= L; b++)
        {
            // the shadow rays of bounce b-1 must have been consumed before shade(b) refills the shadow queue
            if (b > 1) NX_CUDA(ctx, cudaStreamWaitEvent(s, r->evShadow, 0));
            r->prof_begin(2, s);
            shade_kernel<<<gShade, kShadeBlock, 0, s>>>(sv, wb, b, frame);
            r->prof_end(s);
            r->launches++;
            NX_CUDA(ctx, cudaEventRecord(r->evShade, s));
            // shadow rays on the auxiliary stream overlap the extension trace (the reference's graph runs them as siblings)
            NX_CUDA(ctx, cudaStreamWaitEvent(sa, r->evShade, 0));
            r->prof_begin(3, sa);
            if (work) trace_any_kernel<true><<<gAny, NX_TRACE_BLOCK, 0, sa>>>(sv.trace, wb.shadow, 0, &wb.counters->shCount[b], &wb.counters->shFetch[b], nullptr, wb.shadowRad, wb.accum, r->dWork + 1, tune);
            else trace_any_kernel<false><<<gAny, NX_TRACE_BLOCK, 0, sa>>>(sv.trace, wb.shadow, 0, &wb.counters->shCount[b], &wb.counters->shFetch[b], nullptr, wb.shadowRad, wb.accum, nullptr, tune);
            r->prof_end(sa);
            NX_CUDA(ctx, cudaEventRecord(r->evShadow, sa));
            r->launches++;
            if (b < L) closest(wb.ext[b & 1u], b);
        }
        NX_CUDA(ctx, cudaStreamWaitEvent(s, r->evShadow, 0));
        frame_totals_kernel<<<1, 32, 0, s>>>(wb, L);
        r->launches++;
    }
    NX_CUDA(ctx, cudaEventRecord(r->evStop, s));
    NX_CUDA(ctx, cudaGetLastError());
    r->frames += nFrames; r->timed = true; r->pathLengthLast = L;
    return NX_OK;
}

int nx_renderer_frame_count(const nx_renderer* r) { return r ? (int)r->frames : NX_ERR_INVALID; }

int nx_renderer_stats(nx_renderer* r, nx_frame_stats* out)
{
    if (!r || !out) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    std::memset(out, 0, sizeof(*out));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaGetLastError());
    WaveTotals t{};
    NX_CUDA(ctx, cudaMemcpy(&t, r->wb.totals, sizeof(t), cudaMemcpyDeviceToHost));
    out->extension_rays = t.ext; out->shadow_rays = t.shadow; out->shaded_hits = t.shaded; out->frames = t.frames;
    if (r->timed) NX_CUDA(ctx, cudaEventElapsedTime(&out->device_ms, r->evStart, r->evStop));
    out->kernel_launches = r->launches;
    return NX_OK;
}

int nx_renderer_set_profiling(nx_renderer* r, int flags)
{
    if (!r) return NX_ERR_INVALID;
    r->profFlags = flags;
    return NX_OK;
}

int nx_renderer_profile(nx_renderer* r, nx_kernel_profile* out)
{
    if (!r || !out) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    std::memset(out, 0, sizeof(*out));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    for (const auto& l : r->evLaunches) {
        float ms = 0.f;
        NX_CUDA(ctx, cudaEventElapsedTime(&ms, r->evPool[l.second], r->evPool[l.second + 1]));
        out->ms[l.first] += ms; out->launches[l.first]++;
    }
    if (r->profFlags & 2) {
        TraceStats h[2];
        NX_CUDA(ctx, cudaMemcpy(h, r->dWork, sizeof(h), cudaMemcpyDeviceToHost));
        out->closest_work[0] = h[0].nodes; out->closest_work[1] = h[0].tris; out->closest_work[2] = h[0].insts; out->closest_work[3] = h[0].rays;
        out->any_work[0] = h[1].nodes; out->any_work[1] = h[1].tris; out->any_work[2] = h[1].insts; out->any_work[3] = h[1].rays;
        const unsigned long long* a = &h[0].iters; const unsigned long long* b = &h[1].iters;
        for (int k = 0; k < 7; k++) { out->closest_sched[k] = a[k]; out->any_sched[k] = b[k]; }
    }
    return NX_OK;
}

int nx_renderer_read_accum(nx_renderer* r, float* hostRgb)
{
    if (!r || !hostRgb) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    const size_t n = 3 * (size_t)r->width * r->height;
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgb, r->wb.accum, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float inv = r->frames ? 1.0f / (float)r->frames : 0.f;
    for (size_t i = 0; i < n; i++) hostRgb[i] *= inv;
    return NX_OK;
}

int nx_renderer_accum_device(nx_renderer* r, float** outSum, uint32_t* outFrames)
{
    if (!r || !outSum) return NX_ERR_INVALID;
    *outSum = r->wb.accum; if (outFrames) *outFrames = r->frames;
    return NX_OK;
}
int nx_renderer_set_accum_frames(nx_renderer* r, uint32_t frames) { if (!r) return NX_ERR_INVALID; r->frames = frames; return NX_OK; }

int nx_renderer_read_rgba8(nx_renderer* r, nx_scene* scene, uint32_t* hostRgba)
{
    if (!r || !scene || !hostRgba) return NX_ERR_INVALID;
    nx_ctx* ctx = r->ctx;
    DeviceGuard guard(ctx->device);
    const uint32_t count = r->width * r->height;
    uint32_t* d = nullptr;
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream_aux));
    NX_CUDA(ctx, cudaMallocAsync((void**)&d, 4 * (size_t)count, ctx->stream));
    resolve_rgba8_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(r->wb.accum, count, r->frames ? 1.0f / (float)r->frames : 0.f, scene->settings.exposure, d);
    NX_CUDA(ctx, cudaMemcpyAsync(hostRgba, d, 4 * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFreeAsync(d, ctx->stream);
    return NX_OK;
}

} // extern "C"
