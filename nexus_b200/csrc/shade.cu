// Ray generation, shading (logic + OpenPBR material + NEE/MIS) and the display transform of the wavefront path tracer.
// Replaces GenerateKernel, LogicKernel, MaterialKernel + NextEventEstimation and AccumulateKernel of
// src/Cuda/PathTracer/PathTracer.cu:60-95, 124-173, 176-511, 513-549.
//
// This translation unit is compiled with --use_fast_math, as the reference compiles all of its device code
// (Nexus/CMakeLists.txt:75-78): divisions are MUFU.RCP + FMUL, square roots MUFU.RSQ/SQRT, denormals flush to zero.
// Traversal (render.cu) is NOT: its arithmetic decides hit ids and is written with explicit IEEE roundings.
#include "wave.cuh"

namespace {

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
        slot_to_pixel(i, cam.resX, cam.resY, px, py);
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

// positions and vertex normals of one primitive: five 16-byte loads from the shading record (scene.cuh)
struct ShadeTri { F3 v0, v1, v2, n0, n1, n2; };
__device__ __forceinline__ ShadeTri load_shade_tri(const float4* __restrict__ rec, uint32_t prim)
{
    const float4* r = rec + NX_SHADE_REC_F4 * (size_t)prim;
    const float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2), d = __ldg(r + 3), e = __ldg(r + 4);
    ShadeTri t;
    t.v0 = f3(a.x, a.y, a.z); t.v1 = f3(b.x, b.y, b.z); t.v2 = f3(c.x, c.y, c.z);
    t.n0 = f3(a.w, b.w, c.w); t.n1 = f3(d.x, d.y, d.z); t.n2 = f3(d.w, e.x, e.y);
    return t;
}

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
        const float* td = mesh.tridata + 24 * (size_t)ti;
        const ShadeTri st = load_shade_tri(mesh.shade, ti);
        const F3 v0 = st.v0, v1 = st.v1, v2 = st.v2;
        F3 lp = xf_point(I.m0, I.m1, I.m2, bary(v0, v1, v2, u, v));
        const F3 lgn = normalize(xf_normal(I.i0, I.i1, I.i2, cross(v1 - v0, v2 - v0)));
        const F3 ln = normalize(xf_normal(I.i0, I.i1, I.i2, bary(st.n0, st.n1, st.n2, u, v)));
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
        const nx_material& lm = sv.materials[I.materialIdx].m;
        const int32_t emap = __ldg(&lm.emissive_map);
        if (emap != -1) {   // the map REPLACES the emission colour here (PathTracer.cu:263-269), unlike in the material kernel
            const float tu = u * __ldg(td + 20) + v * __ldg(td + 22) + (1.0f - u - v) * __ldg(td + 18);
            const float tv = u * __ldg(td + 21) + v * __ldg(td + 23) + (1.0f - u - v) * __ldg(td + 19);
            const float4 c = tex2D<float4>(sv.textures[emap], tu, tv);
            emissive = f3(c.x, c.y, c.z) * __ldg(&lm.intensity);
        }
        else emissive = f3(__ldg(&lm.emission_color[0]), __ldg(&lm.emission_color[1]), __ldg(&lm.emission_color[2])) * __ldg(&lm.intensity);
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

// LogicKernel for one traced ray (PathTracer.cu:124-173): environment on a miss, Russian roulette.  Returns true when the
// hit goes on to be shaded.
__device__ __forceinline__ bool logic_one(const DSceneView& sv, const WaveBuffers& wb, uint32_t bounce, uint32_t frame, float hitT, F3 rayDir, uint32_t pixel, F3 thr)
{
    if (hitT == NX_MISS_T) { add_radiance(wb.accum, pixel, thr * background(sv, rayDir)); return false; }
    uint32_t rng = rng_seed(pixel, frame, bounce);
    // Russian roulette on the largest throughput component, from the first bounce, no clamp (PathTracer.cu:158-166)
    return rng_next(rng) < max3(thr);
}

// MaterialKernel for one surviving hit (PathTracer.cu:346-511).
__device__ __forceinline__ void material_one(const DSceneView& sv, const WaveBuffers& wb, uint32_t bounce, uint32_t frame, const nx_hit& hit, F3 rayDir,
                                             uint32_t pixel, F3 thr, float lastPdf, ShadeOut& out)
{
    uint32_t rng = rng_seed(pixel, frame, bounce);
    rng_next(rng);                       // the roulette draw logic_one consumed
    thr = thr / max3(thr);

    const DShadeInst I = sv.shadeInst[hit.instance];
    const ShadeTri st = load_shade_tri(I.shade, hit.prim);
    const F3 v0 = st.v0, v1 = st.v1, v2 = st.v2;
    const DMaterial dmat = sv.materials[I.materialIdx];     // 96 B, 16-byte aligned: six LDG.128
    nx_material mat = dmat.m;

    Surface sf;
    sf.p = xf_point(I.m0, I.m1, I.m2, bary(v0, v1, v2, hit.u, hit.v));
    F3 nObj = normalize(bary(st.n0, st.n1, st.n2, hit.u, hit.v));
    // material maps (PathTracer.cu:373-411): all six indices are -1 <=> their AND is -1
    if ((mat.base_color_map & mat.emissive_map & mat.normal_map & mat.roughness_map & mat.metalness_map & mat.metallic_roughness_map) != -1)
    {
        const float* td = sv.meshes[I.meshIdx].tridata + 24 * (size_t)hit.prim;
        const float w0 = 1.0f - hit.u - hit.v;
        const float tu = hit.u * __ldg(td + 20) + hit.v * __ldg(td + 22) + w0 * __ldg(td + 18);
        const float tv = hit.u * __ldg(td + 21) + hit.v * __ldg(td + 23) + w0 * __ldg(td + 19);
        if (mat.normal_map != -1) {
            // tangent-space normal about (normal, Gram-Schmidt tangent) (TangentFrame(n, t), src/Math/TangentFrame.h:24-29)
            const float4 c = tex2D<float4>(sv.textures[mat.normal_map], tu, tv);
            const F3 tn = normalize(2.0f * f3(c.x, c.y, c.z) - f3(1.0f));
            const F3 tg = bary(load3(td + 9), load3(td + 12), load3(td + 15), hit.u, hit.v);
            const F3 t = normalize(tg - dot(tg, nObj) * nObj), b = cross(nObj, t);
            nObj = t * tn.x + b * tn.y + nObj * tn.z;
        }
        if (mat.base_color_map != -1) {
            const float4 c = tex2D<float4>(sv.textures[mat.base_color_map], tu, tv);
            mat.base_color[0] *= c.x; mat.base_color[1] *= c.y; mat.base_color[2] *= c.z; mat.opacity *= c.w;
        }
        if (mat.emissive_map != -1) {
            const float4 c = tex2D<float4>(sv.textures[mat.emissive_map], tu, tv);
            mat.emission_color[0] *= c.x; mat.emission_color[1] *= c.y; mat.emission_color[2] *= c.z;
        }
        if (mat.roughness_map != -1) mat.roughness *= tex2D<float4>(sv.textures[mat.roughness_map], tu, tv).x;
        if (mat.metalness_map != -1) mat.metalness *= tex2D<float4>(sv.textures[mat.metalness_map], tu, tv).x;
        if (mat.metallic_roughness_map != -1) {
            const float4 c = tex2D<float4>(sv.textures[mat.metallic_roughness_map], tu, tv);   // glTF packing: G roughness, B metalness
            mat.roughness *= c.y; mat.metalness *= c.z;
        }
    }
    sf.n = normalize(xf_normal(I.i0, I.i1, I.i2, nObj));
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
            float lightPdf = 1.0f / ((float)sv.lightCount * (float)__ldg(&sv.meshes[I.meshIdx].primCount) * area);
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

// Logic and material in ONE kernel, with the survivors compacted inside the block in between: every thread decides miss /
// roulette for its own ray, survivors' queue indices go into a shared-memory ring, and the material code runs only when a
// full block of survivors is available (plus one flush at the end), so its warps are full instead of carrying the 30-40 %
// of lanes whose path just ended.  The reference gets the same effect with a 36-byte-per-ray material queue in HBM between
// two kernels; here the intermediate is 4 bytes per survivor in shared memory.
__global__ void __launch_bounds__(kShadeBlock, NX_SHADE_MIN_BLOCKS) shade_kernel(const __grid_constant__ DSceneView sv, WaveBuffers wb, uint32_t bounce, uint32_t frame)
{
    __shared__ uint32_t ring[2 * kShadeBlock];
    __shared__ uint32_t ringCount;                 // survivors appended so far (monotone)
    const uint32_t n = wb.counters->extCount[bounce - 1];
    const uint32_t in = (bounce - 1) & 1u, outQ = bounce & 1u;
    uint32_t shadedHere = 0, consumed = 0;         // consumed is block-uniform
    if (threadIdx.x == 0) ringCount = 0;
    __syncthreads();

    auto material_round = [&](bool have, uint32_t i) {    // all threads of the block call this; `have`: the thread owns survivor i
        ShadeOut o; o.ext = false; o.shadow = false;
        uint32_t pixel = 0;
        if (have)
        {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(wb.ext[in] + i) + 1);
            const float4 st = __ldg(wb.state[in] + i);
            const nx_hit h = wb.hits[i];
            pixel = __float_as_uint(d4.w);
            material_one(sv, wb, bounce, frame, h, f3(d4.x, d4.y, d4.z), pixel, f3(st.x, st.y, st.z), st.w, o);
        }
        // both queue appends of the warp with their two atomics in flight together (lane 0: extension queue, lane 1: shadow queue)
        const uint32_t mE = __ballot_sync(NX_FULL, o.ext), mS = __ballot_sync(NX_FULL, o.shadow);
        uint32_t qbase = 0;
        if (lane_id() == 0 && mE) qbase = atomicAdd(&wb.counters->extCount[bounce], __popc(mE));
        if (lane_id() == 1 && mS) qbase = atomicAdd(&wb.counters->shCount[bounce], __popc(mS));
        const uint32_t e = __shfl_sync(NX_FULL, qbase, 0) + __popc(mE & lanemask_lt());
        const uint32_t s = __shfl_sync(NX_FULL, qbase, 1) + __popc(mS & lanemask_lt());
        if (o.ext) {
            float4* r = reinterpret_cast<float4*>(wb.ext[outQ] + e);
            r[0] = make_float4(o.extO.x, o.extO.y, o.extO.z, NX_MISS_T);
            r[1] = make_float4(o.extD.x, o.extD.y, o.extD.z, __uint_as_float(pixel));
            wb.state[outQ][e] = make_float4(o.thr.x, o.thr.y, o.thr.z, o.pdf);
        }
        if (o.shadow) {
            float4* r = reinterpret_cast<float4*>(wb.shadow[bounce & 1u] + s);
            r[0] = make_float4(o.shO.x, o.shO.y, o.shO.z, o.shDist);
            r[1] = make_float4(o.shD.x, o.shD.y, o.shD.z, __uint_as_float(pixel));
            wb.shadowRad[bounce & 1u][s] = make_float4(o.shL.x, o.shL.y, o.shL.z, 0.f);
        }
    };

    for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x)
    {
        const uint32_t i = base + threadIdx.x;
        bool surv = false;
        if (i < n)
        {
            // the next iteration's queue entries, requested first: the logic phase is a streaming read whose DRAM latency is otherwise
            // exposed once per iteration (ncu: long-scoreboard waits on these three loads, 5 CTAs per SM to hide them)
            const uint32_t nx_i = i + gridDim.x * blockDim.x;
            if (nx_i < n && (threadIdx.x & 3u) == 0u) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(wb.ext[in] + nx_i));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(wb.state[in] + nx_i));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(wb.hits + nx_i));
            }
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(wb.ext[in] + i) + 1);
            const float4 st = __ldg(wb.state[in] + i);
            surv = logic_one(sv, wb, bounce, frame, wb.hits[i].t, f3(d4.x, d4.y, d4.z), __float_as_uint(d4.w), f3(st.x, st.y, st.z));
        }
        const uint32_t mask = __ballot_sync(NX_FULL, surv);
        if (mask) {
            const uint32_t leader = __ffs(mask) - 1;
            uint32_t pos = 0;
            if (lane_id() == leader) pos = atomicAdd(&ringCount, __popc(mask));
            pos = __shfl_sync(NX_FULL, pos, leader);
            if (surv) ring[(pos + __popc(mask & lanemask_lt())) & (2 * kShadeBlock - 1)] = i;
        }
        shadedHere += surv ? 1u : 0u;
        __syncthreads();                                         // this iteration's appends are visible
        const uint32_t avail = ringCount - consumed;             // < 2 * kShadeBlock: at most kShadeBlock - 1 were left over
        const bool full = avail >= kShadeBlock;                  // block-uniform
        const uint32_t item = full ? ring[(consumed + threadIdx.x) & (2 * kShadeBlock - 1)] : 0u;
        __syncthreads();                                         // count and ring slots read before anybody appends again
        if (full) { material_round(true, item); consumed += kShadeBlock; }
    }
    const uint32_t rest = ringCount - consumed;                  // no appends after the loop's last barrier
    if (rest) material_round(threadIdx.x < rest, threadIdx.x < rest ? ring[(consumed + threadIdx.x) & (2 * kShadeBlock - 1)] : 0u);
    for (int off = 16; off > 0; off >>= 1) shadedHere += __shfl_xor_sync(NX_FULL, shadedHere, off);
    if (lane_id() == 0 && shadedHere) atomicAdd(&wb.counters->shaded[bounce], shadedHere);
}

// ------------------------------------------------------------------------------------------------ display ----
// Display transform of AccumulateKernel (PathTracer.cu:527-548): exposure, one of the reference's six tone curves
// (src/Utils/ColorUtils.h:9-16: NONE, ACES, UNCHARTED2, AGX_DEFAULT, AGX_GOLDEN, AGX_PUNCHY), gamma 2.2, RGBA8 pack.  The curves
// are the published fits the reference cites: S. Hill's ACES RRT+ODT fit, J. Hable's Uncharted 2 operator with white point
// 11.2, B. Wrensch's minimal AgX (7th-order sigmoid fit) with the golden / punchy ASC-CDL looks.
__device__ __forceinline__ F3 mul3x3(const float (&m)[9], F3 v)
{
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ F3 pow3(F3 v, F3 e) { return f3(powf(v.x, e.x), powf(v.y, e.y), powf(v.z, e.z)); }

__device__ __forceinline__ F3 tone_aces(F3 c)                                    // ColorUtils.h:72-97
{
    const float in[9] = {0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f};
    const float outm[9] = {1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f};
    const F3 v = mul3x3(in, c);
    auto fit = [](float x) { return (x * (x + 0.0245786f) - 0.000090537f) / (x * (0.983729f * x + 0.4329510f) + 0.238081f); };
    const F3 o = mul3x3(outm, f3(fit(v.x), fit(v.y), fit(v.z)));
    return f3(clampf(o.x, 0.f, 1.f), clampf(o.y, 0.f, 1.f), clampf(o.z, 0.f, 1.f));
}
__device__ __forceinline__ float hable(float x)                                 // ColorUtils.h:99-110
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return (x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F) - E / F;
}
__device__ __forceinline__ F3 tone_uncharted2(F3 c)                              // ColorUtils.h:112-117
{
    const float scale = 1.0f / hable(11.2f);
    return f3(hable(1.6f * c.x) * scale, hable(1.6f * c.y) * scale, hable(1.6f * c.z) * scale);
}
__device__ __forceinline__ F3 tone_agx(F3 c, int mode)                           // ColorUtils.h:119-212
{
    const float inset[9] = {0.842479062253094f, 0.0784335999999992f, 0.0792237451477643f, 0.0423282422610123f, 0.878468636469772f, 0.0791661274605434f,
                            0.0423756549057051f, 0.0784336f, 0.879142973793104f};
    const float outset[9] = {1.19687900512017f, -0.0980208811401368f, -0.0990297440797205f, -0.0528968517574562f, 1.15190312990417f, -0.0989611768448433f,
                             -0.0529716355144438f, -0.0980434501171241f, 1.15107367264116f};
    const float minEv = -12.47393f, maxEv = 4.026069f;
    F3 v = mul3x3(inset, c);
    auto enc = [&](float x) { return (clampf(log2f(x), minEv, maxEv) - minEv) / (maxEv - minEv); };
    auto sig = [](float x) {                                                     // 7th-order fit of the AgX default contrast curve
        const float x2 = x * x, x4 = x2 * x2, x6 = x4 * x2;
        return -17.86f * x6 * x + 78.01f * x6 - 126.7f * x4 * x + 92.06f * x4 - 28.72f * x2 * x + 4.361f * x2 - 0.1718f * x + 0.002857f;
    };
    v = f3(sig(enc(v.x)), sig(enc(v.y)), sig(enc(v.z)));
    // look: ASC CDL slope / power, then saturation about the Rec.709 luma
    F3 slope = f3(1.0f), power = f3(1.0f); float sat = 1.0f;
    if (mode == NX_TONE_AGX_GOLDEN) { slope = f3(1.0f, 0.9f, 0.5f); power = f3(0.8f); sat = 0.8f; }
    else if (mode == NX_TONE_AGX_PUNCHY) { power = f3(1.35f); sat = 1.4f; }
    v = pow3(v * slope, power);
    const float luma = 0.2126f * v.x + 0.7152f * v.y + 0.0722f * v.z;
    v = f3(luma + sat * (v.x - luma), luma + sat * (v.y - luma), luma + sat * (v.z - luma));
    return pow3(mul3x3(outset, v), f3(2.2f));                                    // outset, then linearise (gamma is re-applied below)
}
__device__ __forceinline__ uint32_t display_pixel(F3 c, int mode, float gain)
{
    c = c * gain;
    if (mode == NX_TONE_ACES) c = tone_aces(c);
    else if (mode == NX_TONE_UNCHARTED2) c = tone_uncharted2(c);
    else if (mode >= NX_TONE_AGX_DEFAULT && mode <= NX_TONE_AGX_PUNCHY) c = tone_agx(c, mode);
    c = pow3(c, f3(1.0f / 2.2f));
    // ToColorUInt (ColorUtils.h:46-56): clamp (NaN -> 0), scale, truncate
    return (uint32_t)(clampf(c.x, 0.f, 1.f) * 255.0f) | ((uint32_t)(clampf(c.y, 0.f, 1.f) * 255.0f) << 8) | ((uint32_t)(clampf(c.z, 0.f, 1.f) * 255.0f) << 16) | 0xff000000u;
}
__global__ void resolve_rgba8_kernel(const float* __restrict__ accum, uint32_t count, float invFrames, float exposure, int mode, uint32_t* __restrict__ out)
{
    const float gain = invFrames * exp2f(exposure);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        out[i] = display_pixel(f3(accum[3 * (size_t)i], accum[3 * (size_t)i + 1], accum[3 * (size_t)i + 2]), mode, gain);
}

} // namespace

int nxi_shade_grid(nx_ctx* ctx)
{
    int& cache = ctx->gridCache[12];     // per context: resident CTAs per SM x SM count of THIS context's device
    if (cache) return cache;
    int perSm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, (const void*)shade_kernel, kShadeBlock, 0);
    if (perSm < 1) perSm = 1;
    cache = perSm * ctx->sm_count;
    return cache;
}
void nxi_launch_generate(const DSceneView& sv, const WaveBuffers& wb, uint32_t frame, int grid, cudaStream_t s) { generate_kernel<<<grid, 256, 0, s>>>(sv, wb, frame); }
void nxi_launch_shade(const DSceneView& sv, const WaveBuffers& wb, uint32_t bounce, uint32_t frame, int grid, cudaStream_t s)
{
    shade_kernel<<<grid, kShadeBlock, 0, s>>>(sv, wb, bounce, frame);
}
void nxi_launch_resolve(int grid, cudaStream_t s, const float* accum, uint32_t count, float invFrames, float exposure, int mode, uint32_t* out)
{
    resolve_rgba8_kernel<<<grid, 256, 0, s>>>(accum, count, invFrames, exposure, mode, out);
}
