// OpenPBR-style principled BSDF for the shade kernel: metalness-mix( F82-tint conductor,
// transmission-mix( rough dielectric, glossy-diffuse "plastic" ) ), anisotropic GGX with Heitz-2018 VNDF sampling.
//
// Same estimator as the reference (src/Cuda/BSDF/PrincipledBSDF.cuh:12-83, ConductorBSDF.cuh:10-62,
// DielectricBSDF.cuh:17-126, PlasticBSDF.cuh:18-97, Microfacet.cuh:23-62,126-176, Fresnel.cuh:12-85) so converged images
// agree; restructured around one shared microfacet context so D/G1/alpha are evaluated once per lobe call, all in fp32
// (the reference evaluates 2*PI*u in fp64 because its PI is a double literal, src/Utils/Utils.h:7).
#pragma once
#include "nx_common.cuh"
#ifndef NX_BSDF_INLINE
#define NX_BSDF_INLINE __forceinline__
#endif

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ F3 f3(float s) { return {s, s, s}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator-(F3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ F3 operator/(F3 a, float s) { const float i = 1.0f / s; return {a.x * i, a.y * i, a.z * i}; }
__device__ __forceinline__ F3& operator+=(F3& a, F3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
__device__ __forceinline__ F3& operator*=(F3& a, F3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float length(F3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ F3 normalize(F3 a) { return a * rsqrtf(dot(a, a)); }
__device__ __forceinline__ float max3(F3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
__device__ __forceinline__ F3 clamp01(F3 a) { return {__saturatef(a.x), __saturatef(a.y), __saturatef(a.z)}; }
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float pow5(float x) { const float x2 = x * x; return x2 * x2 * x; }
__device__ __forceinline__ float sign_or_one(float x) { return x < 0.0f ? -1.0f : 1.0f; }   // Utils::SgnE

#define NX_PI 3.14159265358979f
#define NX_INV_PI 0.31830988618f
#define NX_TWO_PI 6.28318530718f

// ------------------------------------------------------------------------------------------------- RNG ----
// xorshift32 seeded through a Jenkins hash (src/Cuda/Random.cuh:21-85).  The reference keys the seed on the *queue slot*
// (racing atomics => run-to-run different images); here it is keyed on (pixel, frame, bounce), so a frame is reproducible.
__device__ __forceinline__ uint32_t hash_jenkins(uint32_t x) { x += x << 10; x ^= x >> 6; x += x << 3; x ^= x >> 11; x += x << 15; return x; }
__device__ __forceinline__ uint32_t rng_seed(uint32_t pixel, uint32_t frame, uint32_t bounce)
{
    uint32_t s = (pixel * 0x9E3779B1u + bounce * 0x85EBCA6Bu) ^ hash_jenkins(frame);
    if (s == 0u) s = 1u;
    s = hash_jenkins(s);
    return s ? s : 1u;
}
__device__ __forceinline__ float rng_next(uint32_t& s)
{
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return __uint_as_float(0x3f800000u | (s >> 9)) - 1.0f;
}

// ---------------------------------------------------------------------------------------- tangent frame ----
struct Frame {   // Duff et al. 2017 (src/Math/TangentFrame.h:11-22)
    F3 t, b, n;
    __device__ __forceinline__ explicit Frame(F3 nn) : n(nn)
    {
        const float sg = copysignf(1.0f, nn.z);
        const float a = -1.0f / (sg + nn.z);
        const float c = nn.x * nn.y * a;
        t = f3(1.0f + sg * nn.x * nn.x * a, sg * c, -sg * nn.x);
        b = f3(c, sg + nn.y * nn.y * a, -nn.y);
    }
    __device__ __forceinline__ F3 toLocal(F3 v) const { return f3(dot(v, t), dot(v, b), dot(v, n)); }
    __device__ __forceinline__ F3 toWorld(F3 v) const { return t * v.x + b * v.y + n * v.z; }
};

// ------------------------------------------------------------------------------------------- microfacet ----
struct Ggx {
    float ax, ay;
    // OpenPBR roughness/anisotropy -> alpha mapping, clamped to [1e-4, 1] (ConductorBSDF.cuh:14-19)
    __device__ __forceinline__ Ggx(float roughness, float anisotropy)
    {
        ax = sqr(roughness) * sqrtf(2.0f / (1.0f + sqr(1.0f - anisotropy)));
        ay = (1.0f - anisotropy) * ax;
        ax = fminf(fmaxf(ax, 1.0e-4f), 1.0f); ay = fminf(fmaxf(ay, 1.0e-4f), 1.0f);
    }
    __device__ __forceinline__ float D(F3 m) const { return 1.0f / (NX_PI * ax * ay * sqr(sqr(m.x / ax) + sqr(m.y / ay) + sqr(m.z))); }
    __device__ __forceinline__ float lambda(F3 w) const { return 0.5f * sqrtf(1.0f + (sqr(ax * w.x) + sqr(ay * w.y)) / sqr(w.z)) - 0.5f; }
    __device__ __forceinline__ float G1(F3 w) const { return 1.0f / (1.0f + lambda(w)); }
    __device__ __forceinline__ float G2(F3 wi, F3 wo) const { return 1.0f / (1.0f + lambda(wi) + lambda(wo)); }
    // Heitz 2018 "Sampling the GGX Distribution of Visible Normals" (Microfacet.cuh:126-147,166-176)
    __device__ __forceinline__ F3 sampleVndf(F3 wi, uint32_t& rng) const
    {
        const F3 v = normalize(f3(wi.x * ax, wi.y * ay, wi.z));
        const float u0 = rng_next(rng), u1 = rng_next(rng);
        const float l2 = v.x * v.x + v.y * v.y;
        const F3 w1 = l2 > 0.0f ? f3(-v.y, v.x, 0.0f) * rsqrtf(l2) : f3(1.0f, 0.0f, 0.0f);
        const F3 w2 = cross(v, w1);
        float sn, cs; __sincosf(NX_TWO_PI * u0, &sn, &cs);
        const float r = sqrtf(u1);
        const float t1 = r * cs;
        float t2 = r * sn;
        const float s = 0.5f * (1.0f + v.z);
        t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
        const float t3 = sqrtf(fmaxf(1.0f - t1 * t1 - t2 * t2, 0.0f));
        const F3 h = t1 * w1 + t2 * w2 + t3 * v;
        return normalize(f3(h.x * ax, h.y * ay, h.z));
    }
};
__device__ __forceinline__ float pdf_reflect(float D, float G1, float absCosI) { return G1 * D / (4.0f * absCosI); }
__device__ __forceinline__ float pdf_refract(float D, float G1, float eta, float cosI, float iDotM, float oDotM)
{
    return G1 * D * fabsf(iDotM * oDotM) / (fabsf(cosI) * sqr(eta * iDotM + oDotM));
}
__device__ __forceinline__ bool pdf_ok(float pdf) { return isfinite(pdf) && pdf > 1.0e-4f; }   // Sampler::IsPdfValid
__device__ __forceinline__ F3 reflect_about(F3 wi, F3 m) { return 2.0f * dot(wi, m) * m - wi; }  // reflect(-wi, m)

// ---------------------------------------------------------------------------------------------- Fresnel ----
__device__ __forceinline__ float fresnel_dielectric(float eta, float cosI)   // Fresnel.cuh:14-27
{
    const float sin2T = eta * eta * (1.0f - cosI * cosI);
    if (sin2T >= 1.0f) return 1.0f;
    const float cosT = sqrtf(1.0f - sin2T);
    const float rp = (eta * cosI - cosT) / (eta * cosI + cosT), rs = (eta * cosT - cosI) / (eta * cosT + cosI);
    return 0.5f * (rp * rp + rs * rs);
}
// OpenPBR specular_weight: scales F0 without disturbing TIR (Fresnel.cuh:31-49)
__device__ __forceinline__ float fresnel_dielectric(float eta, float cosI, float specWeight)
{
    if (specWeight == 1.0f) return fresnel_dielectric(eta, cosI);
    const float F0 = sqr((eta - 1.0f) / (eta + 1.0f));
    const float eps = copysignf(fminf(1.0f, sqrtf(specWeight * F0)), 1.0f - eta);
    const float etaP = (1.0f - eps) / fmaxf(1.1920929e-7f, 1.0f + eps);
    if (eta <= 1.0f) return fresnel_dielectric(etaP, cosI);
    const float cos2T = 1.0f - (1.0f - sqr(cosI)) * sqr(eta);
    if (cos2T <= 0.0f) return 1.0f;
    return fresnel_dielectric(1.0f / eta, sqrtf(cos2T));
}
// F82-tint conductor Fresnel (Kutz et al. 2021; Fresnel.cuh:62-85), F82 = specularColor * Schlick(F0, cos)
__device__ __forceinline__ F3 fresnel_f82(F3 F0, F3 specColor, float specWeight, float cosT)
{
    const float k = pow5(1.0f - cosT);
    const F3 schlick = F0 + (f3(1.0f) - F0) * k;
    const F3 F82 = specColor * schlick;
    constexpr float cm = 1.0f / 7.0f, om = 1.0f - cm;
    constexpr float om5 = om * om * om * om * om, om6 = om5 * om;
    const F3 oneMinusF0 = f3(1.0f) - F0;
    const F3 b = ((F0 + oneMinusF0 * om5) * (f3(1.0f) - F82)) * (1.0f / (cm * om6));
    const F3 F = F0 + (oneMinusF0 - b * (cosT * (1.0f - cosT))) * k;
    return specWeight * clamp01(F);
}

// ------------------------------------------------------------------------------------------------ lobes ----
// All lobes work in the local shading frame (z = normal).  eval returns f * |cos(wo)| and the solid-angle pdf.
struct LobeSample { F3 wo; F3 weight; float pdf; bool ok; };

__device__ __forceinline__ bool conductor_eval(const nx_material& M, const Ggx& g, F3 wi, F3 wo, F3& f, float& pdf)
{
    const F3 m = normalize(wo + wi);
    const F3 F = fresnel_f82(f3(M.base_color[0], M.base_color[1], M.base_color[2]), f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]),
                             M.specular_weight, fabsf(dot(wi, m)));
    const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, wo);
    f = F * (G2 * D / (4.0f * fabsf(wi.z)));
    pdf = pdf_reflect(D, G1, fabsf(wi.z));
    return pdf_ok(pdf);
}
__device__ __forceinline__ LobeSample conductor_sample(const nx_material& M, const Ggx& g, F3 wi, uint32_t& rng)
{
    LobeSample s; s.ok = false; s.pdf = 0.f; s.weight = f3(0.f);
    const F3 m = g.sampleVndf(wi, rng);
    const F3 F = fresnel_f82(f3(M.base_color[0], M.base_color[1], M.base_color[2]), f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]),
                             M.specular_weight, fabsf(dot(wi, m)));
    s.wo = reflect_about(wi, m);
    if (s.wo.z * wi.z < 0.0f) return s;
    const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, s.wo);
    s.weight = F * (G2 / G1);
    s.pdf = pdf_reflect(D, G1, fabsf(wi.z));
    s.ok = pdf_ok(s.pdf);
    return s;
}

__device__ __forceinline__ float dielectric_eta(const nx_material& M, F3 wi, bool nudge)
{
    float eta = wi.z < 0.0f ? M.ior : 1.0f / M.ior;
    if (nudge && eta == 1.0f) eta += 1.0e-4f;   // avoids null refracted directions (DielectricBSDF.cuh:25-27)
    return eta;
}
__device__ __forceinline__ bool dielectric_eval(const nx_material& M, const Ggx& g, F3 wi, F3 wo, F3& f, float& pdf)
{
    const float eta = dielectric_eta(M, wi, true);
    const bool refl = wi.z * wo.z > 0.0f;
    const F3 m = refl ? normalize(wo + wi) : normalize(wi * eta + wo);
    const float iM = dot(wi, m), oM = dot(wo, m);
    const float F = fresnel_dielectric(eta, fabsf(iM), M.specular_weight);
    const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, wo);
    if (refl) {
        f = f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]) * (F * G2 * D / (4.0f * fabsf(wi.z)));
        pdf = F * pdf_reflect(D, G1, fabsf(wi.z));
    } else {
        f = f3(M.base_color[0], M.base_color[1], M.base_color[2]) * ((1.0f - F) * G2 * D * fabsf(iM * oM) / (fabsf(wi.z) * sqr(eta * iM + oM)));
        pdf = (1.0f - F) * pdf_refract(D, G1, eta, wi.z, iM, oM);
    }
    return pdf_ok(pdf);
}
__device__ __forceinline__ LobeSample dielectric_sample(const nx_material& M, const Ggx& g, F3 wi, uint32_t& rng)
{
    LobeSample s; s.ok = false; s.pdf = 0.f; s.weight = f3(0.f); s.wo = f3(0.f, 0.f, 1.f);
    const float eta = dielectric_eta(M, wi, true);
    // VNDF sampling works in the upper hemisphere: mirror, sample, mirror back (DielectricBSDF.cuh:66-76)
    F3 up = wi; if (wi.z < 0.0f) up.z = -up.z;
    F3 m = g.sampleVndf(up, rng);
    if (wi.z < 0.0f) m.z = -m.z;
    const float iM = dot(wi, m);
    const float F = fresnel_dielectric(eta, iM, M.specular_weight);
    const bool refl = rng_next(rng) < F;
    if (refl) { s.wo = reflect_about(wi, m); if (s.wo.z * wi.z < 0.0f) return s; }
    else {
        const float cosT = sqrtf(1.0f - sqr(eta) * (1.0f - sqr(iM)));
        s.wo = (eta * iM - sign_or_one(iM) * cosT) * m - eta * wi;
        if (s.wo.z * wi.z > 0.0f) return s;
    }
    const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, s.wo);
    if (refl) {
        s.weight = f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]) * (F * G2 / (G1 * F));
        s.pdf = F * pdf_reflect(D, G1, fabsf(wi.z));
    } else {
        s.weight = f3(M.base_color[0], M.base_color[1], M.base_color[2]) * (G2 / G1);
        s.pdf = (1.0f - F) * pdf_refract(D, G1, eta, wi.z, iM, dot(s.wo, m));
    }
    s.ok = pdf_ok(s.pdf);
    return s;
}

__device__ __forceinline__ bool plastic_eval(const nx_material& M, const Ggx& g, F3 wi, F3 wo, F3& f, float& pdf)
{
    const float eta = dielectric_eta(M, wi, false);
    const F3 m = normalize(wo + wi);
    const float F = fresnel_dielectric(eta, dot(wi, m), M.specular_weight);
    const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, wo);
    const F3 spec = f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]) * (F * G2 * D / (4.0f * fabsf(wi.z)));
    const F3 diff = f3(M.base_color[0], M.base_color[1], M.base_color[2]) * ((1.0f - F) * NX_INV_PI * fabsf(wo.z));
    f = spec + diff;
    pdf = F * pdf_reflect(D, G1, fabsf(wi.z)) + (1.0f - F) * fabsf(wo.z) * NX_INV_PI;
    return pdf_ok(pdf);
}
__device__ __forceinline__ LobeSample plastic_sample(const nx_material& M, const Ggx& g, F3 wi, uint32_t& rng)
{
    LobeSample s; s.ok = false; s.pdf = 0.f; s.weight = f3(0.f);
    const float eta = dielectric_eta(M, wi, false);
    const F3 m = g.sampleVndf(wi, rng);
    const float F = fresnel_dielectric(eta, dot(wi, m), M.specular_weight);
    if (rng_next(rng) < F) {
        s.wo = reflect_about(wi, m);
        if (s.wo.z * wi.z < 0.0f) return s;
        const float D = g.D(m), G1 = g.G1(wi), G2 = g.G2(wi, s.wo);
        s.weight = f3(M.specular_color[0], M.specular_color[1], M.specular_color[2]) * (G2 * F / (G1 * F));
        s.pdf = F * pdf_reflect(D, G1, fabsf(wi.z));
    } else {
        // cosine-weighted hemisphere around +z (Random.cuh:85-98)
        const float r1 = rng_next(rng), r2 = rng_next(rng);
        float sn, cs; __sincosf(NX_TWO_PI * r1, &sn, &cs);
        const float b = sqrtf(r2);
        s.wo = f3(cs * b, sn * b, sqrtf(1.0f - r2));
        s.weight = f3(M.base_color[0], M.base_color[1], M.base_color[2]);
        s.pdf = (1.0f - F) * NX_INV_PI * s.wo.z;
    }
    s.ok = pdf_ok(s.pdf);
    return s;
}

// Principled mix (PrincipledBSDF.cuh:12-83): eval sums the lobes with their selection weights, sample picks one lobe.
__device__ NX_BSDF_INLINE bool principled_eval(const nx_material& M, F3 wi, F3 wo, F3& f, float& pdf)
{
    const Ggx g(M.roughness, M.anisotropy);
    f = f3(0.f); pdf = 0.f;
    F3 lf; float lp;
    if (M.metalness > 0.0f && conductor_eval(M, g, wi, wo, lf, lp)) { f += M.metalness * lf; pdf += M.metalness * lp; }
    if (M.transmission > 0.0f) {
        const float w = (1.0f - M.metalness) * M.transmission;
        if (dielectric_eval(M, g, wi, wo, lf, lp)) { f += w * lf; pdf += w * lp; }
    }
    {
        const float w = (1.0f - M.metalness) * (1.0f - M.transmission);
        if (plastic_eval(M, g, wi, wo, lf, lp)) { f += w * lf; pdf += w * lp; }
    }
    return pdf_ok(pdf);
}
__device__ NX_BSDF_INLINE LobeSample principled_sample(const nx_material& M, F3 wi, uint32_t& rng)
{
    const Ggx g(M.roughness, M.anisotropy);
    LobeSample s; float w;
    if (rng_next(rng) < M.metalness) { s = conductor_sample(M, g, wi, rng); w = M.metalness; }
    else if (rng_next(rng) < M.transmission) { s = dielectric_sample(M, g, wi, rng); w = (1.0f - M.metalness) * M.transmission; }
    else { s = plastic_sample(M, g, wi, rng); w = (1.0f - M.metalness) * (1.0f - M.transmission); }
    s.pdf *= w;
    s.ok = s.ok && pdf_ok(s.pdf);
    return s;
}
