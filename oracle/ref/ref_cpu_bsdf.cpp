// TEST INFRASTRUCTURE ONLY.  The reference's BSDF code (Nexus/src/Cuda/BSDF/*.cuh: D_PrincipledBSDF and the conductor / dielectric / plastic
// lobes, Microfacet, Fresnel) is device-only, but it is plain arithmetic: with host shims for the handful of device intrinsics it uses it
// compiles UNMODIFIED with g++ and runs on the CPU (oracle/Makefile target `refcpu`).  That turns row a8 of SURVEY.md section 8 from
// "pinned statistically through converged images" into a function-level pin: D_PrincipledBSDF::Eval for given (material, wi, wo).
// Host arithmetic is IEEE; on the GPU the reference is compiled with --use_fast_math, so values agree to approximation error only.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "Utils/Utils.h"
static inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
extern "C" void __sincosf(float x, float* s, float* c) noexcept { *s = std::sin(x); *c = std::cos(x); }   // declared (not defined) for the host by the CUDA headers
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float __fdividef(float a, float b) { return a / b; }
// CUDA provides max / min overloads for floats in device code; on the host the reference's cuda_math.h only declares the int ones,
// which would silently truncate (e.g. sqrt(max(1 - t1*t1 - t2*t2, 0.0f)) in Microfacet::SampleVndf_Heitz)
static inline float max(float a, float b) { return a > b ? a : b; }
static inline float min(float a, float b) { return a < b ? a : b; }
static inline double max(double a, double b) { return a > b ? a : b; }
static inline double min(double a, double b) { return a < b ? a : b; }
#include "Cuda/BSDF/PrincipledBSDF.cuh"

static_assert(sizeof(D_Material) == 92, "D_Material layout");

// mats: n x 92 B (D_Material), wi / wo: n x 3 floats in the local shading frame (z = normal).  out: bsdf n x 3, pdf n, ok n.
extern "C" int ref_bsdf_eval(const void* mats, const float* wi, const float* wo, uint32_t n, float* outBsdf, float* outPdf, uint8_t* outOk)
{
    for (uint32_t i = 0; i < n; i++) {
        D_Material m; std::memcpy(&m, (const uint8_t*)mats + 92 * (size_t)i, 92);
        float3 f; float pdf;
        const bool ok = D_PrincipledBSDF::Eval(m, make_float3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), make_float3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), f, pdf);
        outBsdf[3 * i] = f.x; outBsdf[3 * i + 1] = f.y; outBsdf[3 * i + 2] = f.z; outPdf[i] = pdf; outOk[i] = ok ? 1 : 0;
    }
    return 0;
}

// TangentFrame(n) (Nexus/src/Math/TangentFrame.h:11-22, Duff et al. 2017): tangent, bitangent, normal as 9 floats per input normal.
#include "Math/TangentFrame.h"
extern "C" int ref_tangent_frame(const float* normals, uint32_t n, float* out9)
{
    for (uint32_t i = 0; i < n; i++) {
        const TangentFrame f(make_float3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
        const float v[9] = {f.tangent.x, f.tangent.y, f.tangent.z, f.bitangent.x, f.bitangent.y, f.bitangent.z, f.normal.x, f.normal.y, f.normal.z};
        std::memcpy(out9 + 9 * (size_t)i, v, sizeof(v));
    }
    return 0;
}

// D_PrincipledBSDF::Sample with the reference's own RNG seeded by seeds[i]: outgoing direction, path weight ("throughput"), pdf.
extern "C" int ref_bsdf_sample(const void* mats, const float* wi, const uint32_t* seeds, uint32_t n, float* outWo, float* outWeight, float* outPdf, uint8_t* outOk)
{
    for (uint32_t i = 0; i < n; i++) {
        D_Material m; std::memcpy(&m, (const uint8_t*)mats + 92 * (size_t)i, 92);
        float3 wo = make_float3(0.0f), w = make_float3(0.0f); float pdf = 0.0f; unsigned int rng = seeds[i];
        const bool ok = D_PrincipledBSDF::Sample(m, make_float3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), wo, w, pdf, rng);
        outWo[3 * i] = wo.x; outWo[3 * i + 1] = wo.y; outWo[3 * i + 2] = wo.z;
        outWeight[3 * i] = w.x; outWeight[3 * i + 1] = w.y; outWeight[3 * i + 2] = w.z;
        outPdf[i] = pdf; outOk[i] = ok ? 1 : 0;
    }
    return 0;
}
