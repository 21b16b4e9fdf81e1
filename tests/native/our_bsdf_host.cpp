// TEST HELPER.  The product's BSDF header (nexus_b200/csrc/bsdf.cuh, device code of shade_kernel) compiled for the host with shims for the
// device intrinsics it uses, so that principled_eval / principled_sample can be called from the CPU tests and compared, value for value,
// with the reference's D_PrincipledBSDF compiled the same way (oracle/ref/ref_cpu_bsdf.cpp).  Built on the fly by tests/test_bsdf_host.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
using std::isfinite;
static inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }
extern "C" void __sincosf(float x, float* s, float* c) noexcept { *s = std::sin(x); *c = std::cos(x); }   // declared (not defined) for the host by the CUDA headers
static inline float __saturatef(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline int __popc(unsigned int v) { return __builtin_popcount(v); }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcg(T* p, T v) { *p = v; }
static struct { unsigned x, y, z; } threadIdx;
#define __forceinline__ inline
#include "bsdf.cuh"

static_assert(sizeof(nx_material) == 92, "nx_material layout");

extern "C" int our_bsdf_eval(const void* mats, const float* wi, const float* wo, uint32_t n, float* outBsdf, float* outPdf, uint8_t* outOk)
{
    for (uint32_t i = 0; i < n; i++) {
        nx_material m; std::memcpy(&m, (const uint8_t*)mats + 92 * (size_t)i, 92);
        F3 f; float pdf;
        const bool ok = principled_eval(m, f3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), f3(wo[3 * i], wo[3 * i + 1], wo[3 * i + 2]), f, pdf);
        outBsdf[3 * i] = f.x; outBsdf[3 * i + 1] = f.y; outBsdf[3 * i + 2] = f.z; outPdf[i] = pdf; outOk[i] = ok ? 1 : 0;
    }
    return 0;
}
// One sample per input with the product's RNG seeded by `seeds[i]`: outgoing direction, path weight (f * |cos| / pdf) and pdf.
extern "C" int our_bsdf_sample(const void* mats, const float* wi, const uint32_t* seeds, uint32_t n, float* outWo, float* outWeight, float* outPdf, uint8_t* outOk)
{
    for (uint32_t i = 0; i < n; i++) {
        nx_material m; std::memcpy(&m, (const uint8_t*)mats + 92 * (size_t)i, 92);
        uint32_t rng = seeds[i];
        const LobeSample s = principled_sample(m, f3(wi[3 * i], wi[3 * i + 1], wi[3 * i + 2]), rng);
        outWo[3 * i] = s.wo.x; outWo[3 * i + 1] = s.wo.y; outWo[3 * i + 2] = s.wo.z;
        outWeight[3 * i] = s.weight.x; outWeight[3 * i + 1] = s.weight.y; outWeight[3 * i + 2] = s.weight.z;
        outPdf[i] = s.pdf; outOk[i] = s.ok ? 1 : 0;
    }
    return 0;
}
// Frame(n) of bsdf.cuh: tangent, bitangent, normal as 9 floats per input normal.
extern "C" int our_tangent_frame(const float* normals, uint32_t n, float* out9)
{
    for (uint32_t i = 0; i < n; i++) {
        const Frame f(f3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
        const float v[9] = {f.t.x, f.t.y, f.t.z, f.b.x, f.b.y, f.b.z, f.n.x, f.n.y, f.n.z};
        std::memcpy(out9 + 9 * (size_t)i, v, sizeof(v));
    }
    return 0;
}
