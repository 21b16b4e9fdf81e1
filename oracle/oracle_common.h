// TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference algorithms.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use this.
// The product library (nexus_b200/csrc) never includes, links or calls anything under oracle/.
//
// Parity pin: the reference ships no tests or golden vectors.  This restatement is
// pinned against outputs of the reference's own CUDA code (oracle/_ref, built by `make ref`) run on a
// B200 and committed under tests/golden/ (see tests/golden/README.md for the generating command).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>

namespace orc {

struct f3 { float x, y, z; };
static inline f3 mk(float x, float y, float z) { return {x, y, z}; }
static inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(float s, f3 a) { return {a.x * s, a.y * s, a.z * s}; }
static inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
// PTX min.f32 / max.f32 order the zeros (-0 < +0); libm's fminf/fmaxf may return either.  Leaf bounds of triangles with a
// vertex at -0 depend on it (seen on the UV sphere's pole rows), so the restatement pins the GPU behaviour.
static inline float gmin(float a, float b) { if (a == 0.0f && b == 0.0f) return (std::signbit(a) || std::signbit(b)) ? -0.0f : 0.0f; return fminf(a, b); }
static inline float gmax(float a, float b) { if (a == 0.0f && b == 0.0f) return (std::signbit(a) && std::signbit(b)) ? -0.0f : 0.0f; return fmaxf(a, b); }
static inline f3 vmin(f3 a, f3 b) { return {gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)}; }
static inline f3 vmax(f3 a, f3 b) { return {gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)}; }
static inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static inline float length(f3 a) { return sqrtf(dot(a, a)); }
static inline f3 normalize(f3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// The reference is compiled with -ftz=true (--use_fast_math): subnormal results are flushed to +-0.
static inline float ftz(float f) { uint32_t u = f2u(f); return (u & 0x7f800000u) == 0 ? u2f(u & 0x80000000u) : f; }

// NXB::AABB, B/include/NXB/AABB.h:8-53
struct AABB {
    f3 bMin, bMax;
    void clear() { bMin = mk(FLT_MAX, FLT_MAX, FLT_MAX); bMax = mk(-FLT_MAX, -FLT_MAX, -FLT_MAX); }
    void grow(const AABB& o) { bMin = vmin(bMin, o.bMin); bMax = vmax(bMax, o.bMax); }
    void grow(f3 p) { bMin = vmin(bMin, p); bMax = vmax(bMax, p); }
    // AABB::Area() (AABB.h:43-47) as the reference compiles it for sm_100a (fmad contraction, checked in the
    // emitted PTX): mul dy*dz ; fma dx*dy + . ; fma dx*dz + .
    float area() const {
        float dx = bMax.x - bMin.x, dy = bMax.y - bMin.y, dz = bMax.z - bMin.z;
        return fmaf(dx, dz, fmaf(dx, dy, dy * dz));
    }
    // The same expression as a host compiler evaluates it (x86-64 without FMA: three products, two sums, each rounded): what the
    // reference's CPU BVH8Builder computes.  The translation unit is built with -ffp-contract=off, so this stays uncontracted.
    float areaHost() const {
        float dx = bMax.x - bMin.x, dy = bMax.y - bMin.y, dz = bMax.z - bMin.z;
        return dx * dy + dy * dz + dz * dx;
    }
};

// NXB::BVH2::Node, B/include/NXB/BVH.h:20-29 (32 B)
struct Node2 { AABB bounds; uint32_t left, right; };
static_assert(sizeof(Node2) == 32, "BVH2 node layout");

// NXB::BVH8::NodeExplicit, B/include/NXB/BVH.h:42-69 (80 B)
struct Node8 {
    f3 p; uint8_t e[3]; uint8_t imask;
    uint32_t childBaseIdx, primBaseIdx;
    uint8_t meta[8];
    uint8_t qlox[8], qloy[8], qloz[8], qhix[8], qhiy[8], qhiz[8];
};
static_assert(sizeof(Node8) == 80, "BVH8 node layout");

constexpr uint32_t INVALID = 0xffffffffu;

} // namespace orc
