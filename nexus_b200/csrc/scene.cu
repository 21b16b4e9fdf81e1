// Host scene logic behind the C ABI: materials, meshes (+BLAS), instances (+TLAS), lights, camera, device mirrors.
// Mirrors Scene / AssetManager / Mesh / MeshInstance / Camera of the reference (src/Scene/*.cpp, src/Assets/*.h); the
// reference's DeviceVector/DeviceInstance plumbing is replaced by plain uploads in nx_scene_update().
#include "scene.cuh"
#include <cmath>
#include <algorithm>
#include <chrono>
#include <cstdlib>

namespace {

// ---------------------------------------------------------------------------------- row-major 4x4 helpers ----
struct M4 { float c[16]; };
M4 m4_identity() { M4 r; for (int i = 0; i < 16; i++) r.c[i] = (i % 5 == 0) ? 1.f : 0.f; return r; }
M4 m4_mul(const M4& a, const M4& b)
{
    M4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
        r.c[4 * i + j] = a.c[4 * i] * b.c[j] + a.c[4 * i + 1] * b.c[4 + j] + a.c[4 * i + 2] * b.c[8 + j] + a.c[4 * i + 3] * b.c[12 + j];
    return r;
}
float to_radians(float deg) { return (float)((double)deg * 3.14159265358979323846 / 180.0); }   // Utils::ToRadians (double PI)
M4 m4_translate(const float p[3]) { M4 r = m4_identity(); r.c[3] = p[0]; r.c[7] = p[1]; r.c[11] = p[2]; return r; }
M4 m4_scale(const float s[3]) { M4 r = m4_identity(); r.c[0] = s[0]; r.c[5] = s[1]; r.c[10] = s[2]; return r; }
M4 m4_rotx(float a) { M4 r = m4_identity(); r.c[5] = cosf(a); r.c[6] = -sinf(a); r.c[9] = sinf(a); r.c[10] = cosf(a); return r; }
M4 m4_roty(float a) { M4 r = m4_identity(); r.c[0] = cosf(a); r.c[2] = sinf(a); r.c[8] = -sinf(a); r.c[10] = cosf(a); return r; }
M4 m4_rotz(float a) { M4 r = m4_identity(); r.c[0] = cosf(a); r.c[1] = -sinf(a); r.c[4] = sinf(a); r.c[5] = cosf(a); return r; }

// General inverse by cofactors (the instance matrices are affine, but user matrices need not be).
bool m4_inverse(const M4& m, M4& out)
{
    const float* a = m.c;
    float s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
    float s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
    float c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
    float c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
    float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    if (det == 0.f) { out = m4_identity(); return false; }
    float id = 1.0f / det;
    float* b = out.c;
    b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * id;   b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * id;
    b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * id; b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * id;
    b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * id;  b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * id;
    b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * id; b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * id;
    b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * id;   b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * id;
    b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * id; b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * id;
    b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * id; b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * id;
    b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * id; b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * id;
    return true;
}
void m4_point(const float* m, const float p[3], float out[3])
{
    for (int r = 0; r < 3; r++) out[r] = m[4 * r] * p[0] + m[4 * r + 1] * p[1] + m[4 * r + 2] * p[2] + m[4 * r + 3];
}

// MeshInstance::GetTransfromationMatrix: T * Rz * Ry * Rx * S (src/Scene/MeshInstance.h:36-40)
M4 compose_trs(const float pos[3], const float rotDeg[3], const float scale[3])
{
    return m4_mul(m4_mul(m4_mul(m4_mul(m4_translate(pos), m4_rotz(to_radians(rotDeg[2]))), m4_roty(to_radians(rotDeg[1]))), m4_rotx(to_radians(rotDeg[0]))),
                  m4_scale(scale));
}
// MeshInstance::GetBounds: AABB of the eight transformed corners (MeshInstance.h:42-53)
nx_aabb transformed_bounds(const float* m, const nx_aabb& b)
{
    nx_aabb r; for (int k = 0; k < 3; k++) { r.bmin[k] = 3.402823466e38f; r.bmax[k] = -3.402823466e38f; }
    for (int i = 0; i < 8; i++) {
        float p[3] = {(i & 1) ? b.bmax[0] : b.bmin[0], (i & 2) ? b.bmax[1] : b.bmin[1], (i & 4) ? b.bmax[2] : b.bmin[2]}, q[3];
        m4_point(m, p, q);
        for (int k = 0; k < 3; k++) { r.bmin[k] = fminf(r.bmin[k], q[k]); r.bmax[k] = fmaxf(r.bmax[k], q[k]); }
    }
    return r;
}

// Bounds of a mesh's vertices: the AABB (exact float min / max - the same values the builder's scene-bounds reduction produces) and the
// bounding sphere around the AABB centre, radius = farthest vertex, padded for the float evaluation.  Written over blocks of four
// vertices with twelve independent accumulators so that the host compiler vectorises it: the scalar double-precision version was the
// largest single item of the scene set-up (0.46 s of 0.8 s for the 1,026 meshes of BASELINE configs[2]).
void mesh_bounds_sphere(const nx_triangle* tris, uint32_t n, nx_aabb* box, double out[4])
{
    const float* v = (const float*)tris;
    const size_t nf = (size_t)n * 9, blocks = nf / 12;
    float lo[12], hi[12];
    for (int k = 0; k < 12; k++) { lo[k] = 3.402823466e38f; hi[k] = -3.402823466e38f; }
    for (size_t b = 0; b < blocks; b++) {
        const float* p = v + 12 * b;
        for (int k = 0; k < 12; k++) { lo[k] = p[k] < lo[k] ? p[k] : lo[k]; hi[k] = p[k] > hi[k] ? p[k] : hi[k]; }
    }
    float l3[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, h3[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int k = 0; k < 12; k++) { l3[k % 3] = fminf(l3[k % 3], lo[k]); h3[k % 3] = fmaxf(h3[k % 3], hi[k]); }
    for (size_t i = 12 * blocks; i < nf; i++) { l3[i % 3] = fminf(l3[i % 3], v[i]); h3[i % 3] = fmaxf(h3[i % 3], v[i]); }
    for (int k = 0; k < 3; k++) { box->bmin[k] = l3[k]; box->bmax[k] = h3[k]; out[k] = 0.5 * ((double)l3[k] + (double)h3[k]); }
    const float c[3] = {(float)out[0], (float)out[1], (float)out[2]};
    float c12[12], r4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 12; k++) c12[k] = c[k % 3];
    for (size_t b = 0; b < blocks; b++) {           // four vertices per block: twelve independent squares, four sums
        const float* p = v + 12 * b;
        float sq[12];
        for (int k = 0; k < 12; k++) { const float d = p[k] - c12[k]; sq[k] = d * d; }
        for (int j = 0; j < 4; j++) { const float d2 = sq[3 * j] + sq[3 * j + 1] + sq[3 * j + 2]; r4[j] = d2 > r4[j] ? d2 : r4[j]; }
    }
    float r2 = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
    for (size_t i = 12 * blocks; i + 2 < nf; i += 3) {
        const float dx = v[i] - c[0], dy = v[i + 1] - c[1], dz = v[i + 2] - c[2];
        r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
    }
    // float evaluation: centre rounded to float (<= 2^-24 relative of the coordinates), each product and sum rounded once more
    const double mag = std::fabs(out[0]) + std::fabs(out[1]) + std::fabs(out[2]);
    out[3] = std::sqrt((double)r2) * (1.0 + 1e-5) + 1e-6 * mag + 1e-30;
}

// Largest singular value of the upper-left 3x3 of a row-major 4x4 (power iteration on A^T A, double): the factor by which the
// instance transform can stretch a sphere radius.
double max_stretch(const float* m)
{
    double a[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = m[4 * i + j];
    double g[3][3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { g[i][j] = 0; for (int k = 0; k < 3; k++) g[i][j] += a[k][i] * a[k][j]; }
    double best = 0.0;
    const double starts[3][3] = {{1, 0.3, 0.2}, {0.2, 1, 0.3}, {0.3, 0.2, 1}};
    for (const auto& s0 : starts) {
        double v[3] = {s0[0], s0[1], s0[2]}, lambda = 0.0;
        for (int it = 0; it < 64; it++) {
            double w[3] = {0, 0, 0};
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) w[i] += g[i][j] * v[j];
            const double nrm = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            if (nrm == 0.0) break;
            lambda = nrm;
            for (int i = 0; i < 3; i++) v[i] = w[i] / nrm;
        }
        best = lambda > best ? lambda : best;
    }
    // power iteration approaches the top eigenvalue from below: bound it by the Frobenius norm and pad
    const double fro = g[0][0] + g[1][1] + g[2][2];
    double sigma = std::sqrt(best) * (1.0 + 1e-3);
    const double cap = std::sqrt(fro);
    return sigma < cap ? sigma : cap;
}

bool material_emits(const nx_material& m)   // Scene::UpdateSceneLighting's test (Scene.cpp:162-185)
{
    float mx = fmaxf(m.emission_color[0], fmaxf(m.emission_color[1], m.emission_color[2]));
    return (m.emissive_map != -1 || mx > 0.0f) && m.intensity > 0.0f;
}

// ------------------------------------------------------------------------------------------- device prep ----
// Leaf-ordered triangle stream: slot k of the BLAS gets {v0, primId}, {v1 - v0, 0}, {v2 - v0, 0}.
__global__ void leaf_triangles_kernel(const float* __restrict__ tris, const uint32_t* __restrict__ primIdx, uint32_t n, float4* __restrict__ out, uint32_t* bad)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(primIdx + k);
        if (p >= n) { if (bad) atomicAdd(bad, 1u); continue; }     // only a caller-supplied (prebuilt) BLAS can get here; reported by Scene::Update
        const float* t = tris + 9 * (size_t)p;
        const V3 a = v3(__ldg(t), __ldg(t + 1), __ldg(t + 2)), b = v3(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5)), c = v3(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8));
        const V3 e0 = b - a, e1 = c - a;
        out[3 * (size_t)k] = make_float4(a.x, a.y, a.z, __uint_as_float(p));
        out[3 * (size_t)k + 1] = make_float4(e0.x, e0.y, e0.z, 0.f);
        out[3 * (size_t)k + 2] = make_float4(e1.x, e1.y, e1.z, 0.f);
    }
}

// Shading records (scene.cuh): positions from the triangle array, vertex normals from the D_TriangleData array.
__global__ void shade_record_kernel(const float* __restrict__ tris, const float* __restrict__ tridata, uint32_t n, float4* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* t = tris + 9 * (size_t)i; const float* d = tridata + 24 * (size_t)i;
        float4* o = out + NX_SHADE_REC_F4 * (size_t)i;
        o[0] = make_float4(t[0], t[1], t[2], d[0]);
        o[1] = make_float4(t[3], t[4], t[5], d[1]);
        o[2] = make_float4(t[6], t[7], t[8], d[2]);
        o[3] = make_float4(d[3], d[4], d[5], d[6]);
        o[4] = make_float4(d[7], d[8], 0.f, 0.f);
    }
}

// ---- merged BLAS: every instance whose mesh nobody else uses, under ONE tree built over the world-space boxes of their triangles ----
// The leaf records stay in object space (traverse.cuh: a triangle is tested with the ray taken into its instance's object space, so the
// hits are the two-level scene's bit for bit); only the NODES are in world space.
struct BakeSrc { const float* tris; float m[12]; uint32_t first, count, inst; float mag; };   // source j covers merged primitives [first, first + count)

__device__ __forceinline__ uint32_t bake_source_of(const BakeSrc* __restrict__ src, uint32_t nSrc, uint32_t p)
{
    uint32_t lo = 0, hi = nSrc - 1;                                // the last source whose `first` is <= p
    while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (__ldg(&src[mid].first) <= p) lo = mid; else hi = mid - 1; }
    return lo;
}
// World-space AABB of every merged primitive (nx_aabb: min, max), padded.  The box is that of the fp32-transformed vertices; the
// triangle test, however, runs in object space on the fp32-transformed RAY, so the surface the ray can hit sits where the exact
// transform would put it, give or take the rounding of both transforms: a few ulps of the largest magnitude involved (world
// coordinates and translation).  pad = 2^-21 of that magnitude (4 ulps) per side.
__global__ void bake_bounds_kernel(const BakeSrc* __restrict__ src, uint32_t nSrc, uint32_t total, float* __restrict__ out, uint32_t* __restrict__ srcOf)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const uint32_t j = bake_source_of(src, nSrc, p);
        const BakeSrc S = src[j];
        const float* t = S.tris + 9 * (size_t)(p - S.first);
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        float mag = S.mag;
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const float x = __ldg(t + 3 * v), y = __ldg(t + 3 * v + 1), z = __ldg(t + 3 * v + 2);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float w = __fmaf_rn(S.m[4 * a], x, __fmaf_rn(S.m[4 * a + 1], y, __fmaf_rn(S.m[4 * a + 2], z, S.m[4 * a + 3])));
                lo[a] = fminf(lo[a], w); hi[a] = fmaxf(hi[a], w); mag = fmaxf(mag, fabsf(w));
            }
        }
        const float pad = __fmul_rn(mag, 0x1p-21f);
        float* o = out + 6 * (size_t)p;
#pragma unroll
        for (int a = 0; a < 3; a++) { o[a] = __fsub_rd(lo[a], pad); o[3 + a] = __fadd_ru(hi[a], pad); }
        srcOf[p] = j;
    }
}
// leaf-ordered stream of the merged BLAS, OBJECT space: {v0 | primitive id inside its mesh}, {v1 - v0 | instance id}, {v2 - v0 | 0}:
// the floats leaf_triangles_kernel writes for the mesh's own BLAS
__global__ void merged_leaf_kernel(const uint32_t* __restrict__ primIdx, uint32_t n, const uint32_t* __restrict__ srcOf,
                                   const BakeSrc* __restrict__ src, float4* __restrict__ out)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(primIdx + k);
        const uint32_t j = __ldg(srcOf + p);
        const uint32_t q = p - __ldg(&src[j].first);
        const float* t = src[j].tris + 9 * (size_t)q;
        const V3 a = v3(__ldg(t), __ldg(t + 1), __ldg(t + 2)), b = v3(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5)), c = v3(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8));
        const V3 e0 = b - a, e1 = c - a;
        out[3 * (size_t)k] = make_float4(a.x, a.y, a.z, __uint_as_float(q));
        out[3 * (size_t)k + 1] = make_float4(e0.x, e0.y, e0.z, __uint_as_float(__ldg(&src[j].inst)));
        out[3 * (size_t)k + 2] = make_float4(e1.x, e1.y, e1.z, 0.f);
    }
}

// Default shading data when the caller passes none: flat geometric normal, zero tangents and texture coordinates.
// World-space AABB of an instance's actual geometry: one block per instance transforms every vertex of its mesh with the
// instance matrix (row-major 3x4) and reduces min / max.  out: 6 floats per instance.
struct InstGeom { const float* tris; uint32_t triCount; float m[12]; };
__global__ void __launch_bounds__(256) instance_geometry_bounds_kernel(const InstGeom* __restrict__ inst, float* __restrict__ out)
{
    const InstGeom g = inst[blockIdx.x];
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (uint32_t v = threadIdx.x; v < 3u * g.triCount; v += blockDim.x) {
        const float x = __ldg(g.tris + 3 * (size_t)v), y = __ldg(g.tris + 3 * (size_t)v + 1), z = __ldg(g.tris + 3 * (size_t)v + 2);
        for (int a = 0; a < 3; a++) {
            const float w = g.m[4 * a] * x + g.m[4 * a + 1] * y + g.m[4 * a + 2] * z + g.m[4 * a + 3];
            lo[a] = fminf(lo[a], w); hi[a] = fmaxf(hi[a], w);
        }
    }
    __shared__ float red[6][8];
    for (int a = 0; a < 3; a++) {
        for (int o = 16; o > 0; o >>= 1) { lo[a] = fminf(lo[a], __shfl_xor_sync(NX_FULL, lo[a], o)); hi[a] = fmaxf(hi[a], __shfl_xor_sync(NX_FULL, hi[a], o)); }
        if ((threadIdx.x & 31) == 0) { red[a][threadIdx.x >> 5] = lo[a]; red[3 + a][threadIdx.x >> 5] = hi[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[threadIdx.x][0];
        for (int w = 1; w < 8; w++) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][w]) : fmaxf(v, red[threadIdx.x][w]);
        out[6 * (size_t)blockIdx.x + threadIdx.x] = v;
    }
}

__global__ void default_tridata_kernel(const float* __restrict__ tris, uint32_t n, float* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* t = tris + 9 * (size_t)i;
        const float ax = t[3] - t[0], ay = t[4] - t[1], az = t[5] - t[2], bx = t[6] - t[0], by = t[7] - t[1], bz = t[8] - t[2];
        float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
        const float l = sqrtf(nx * nx + ny * ny + nz * nz), il = l > 0.f ? 1.0f / l : 0.f;
        nx *= il; ny *= il; nz *= il;
        float* o = out + 24 * (size_t)i;
        for (int v = 0; v < 3; v++) { o[3 * v] = nx; o[3 * v + 1] = ny; o[3 * v + 2] = nz; }
        for (int k = 9; k < 24; k++) o[k] = 0.f;
    }
}

template <typename T> int upload_vec(nx_ctx* ctx, T** dst, const std::vector<T>& src)
{
    if (*dst) { cudaFreeAsync(*dst, ctx->stream); *dst = nullptr; }
    NX_CUDA(ctx, cudaMallocAsync((void**)dst, sizeof(T) * std::max<size_t>(src.size(), 1), ctx->stream));
    if (!src.empty()) NX_CUDA(ctx, cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
    return NX_OK;
}

} // namespace

// Camera::ToDevice (src/Scene/Camera.cpp:130-156)
DCamera nxi_camera_to_device(const nx_camera& c, uint32_t w, uint32_t h)
{
    DCamera d{};
    float f[3] = {c.forward[0], c.forward[1], c.forward[2]};
    float r[3] = {c.right[0], c.right[1], c.right[2]};
    if (r[0] == 0.f && r[1] == 0.f && r[2] == 0.f) { r[0] = f[1] * 0.f - f[2] * 1.f; r[1] = f[2] * 0.f - f[0] * 0.f; r[2] = f[0] * 1.f - f[1] * 0.f; }  // cross(forward, +Y)
    float up[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};                                                 // cross(right, forward)
    const float aspect = (float)w / (float)h;
    const float halfW = c.focus_distance * tanf((float)((double)(c.horizontal_fov_deg / 2.0f) * 3.14159265358979323846 / 180.0));
    const float halfH = halfW / aspect;
    const float lens = c.focus_distance * tanf((float)((double)(c.defocus_angle_deg / 2.0f) * 3.14159265358979323846 / 180.0));
    for (int k = 0; k < 3; k++) {
        d.position[k] = c.position[k]; d.right[k] = r[k]; d.up[k] = up[k];
        d.viewportX[k] = 2 * halfW * r[k]; d.viewportY[k] = 2 * halfH * up[k];
        d.lowerLeft[k] = c.position[k] - d.viewportX[k] / 2.0f - d.viewportY[k] / 2.0f + f[k] * c.focus_distance;
    }
    d.lensRadius = lens; d.resX = w; d.resY = h;
    return d;
}

int nxi_scene_view(nx_scene* s, DSceneView* v)
{
    if (s->dirtyInstances || s->dirtyMaterials || s->dirtyLights || s->dirtyTextures) { int rc = nx_scene_update(s); if (rc) return rc; }
    std::memset(v, 0, sizeof(*v));
    v->trace.tlasNodes = s->dTopNodes; v->trace.tlasPrimIdx = s->dSlotInst; v->trace.inst = s->dTravInst; v->trace.overflow = s->ctx->dOverflow;
    v->trace.mergedSlot = s->mergedSlot;
    v->trace.direct = (s->mergedSlot != NX_INVALID && s->tlasEntryInst.size() == 1) ? 1u : 0u;
    v->trace.mNodes = (const float4*)s->merged.nodes; v->trace.mLtris = s->dMergedLeaf; v->trace.instInv = s->dInstInv;
    v->shadeInst = s->dShadeInst; v->meshes = s->dMeshes; v->materials = s->dMaterials; v->lights = s->dLights;
    v->lightCount = (uint32_t)s->lights.size(); v->hasHdr = s->hasHdr ? 1u : 0u; v->hdr = s->hdr; v->textures = s->dTextures;
    v->camera = nxi_camera_to_device(s->camera, s->width, s->height);
    v->useMIS = s->settings.use_mis ? 1u : 0u; v->pathLength = (uint32_t)s->settings.path_length;
    for (int k = 0; k < 3; k++) v->bg[k] = s->settings.background_color[k];
    v->bgIntensity = s->settings.background_intensity;
    return NX_OK;
}

extern "C" {

int nx_scene_create(nx_ctx* ctx, uint32_t width, uint32_t height, nx_scene** out)
{
    if (!ctx || !out || !width || !height) return NX_ERR_INVALID;
    nx_scene* s = new nx_scene();
    s->ctx = ctx; s->width = width; s->height = height;
    // Scene::Scene default camera (src/Scene/Scene.cpp:8-12) and RenderSettings defaults (RenderSettings.h:5-17)
    s->camera = nx_camera{{0.f, 4.f, 14.f}, {0.f, 0.f, -1.f}, {0.f, 0.f, 0.f}, 45.f, 5.f, 0.f};
    s->settings = nx_render_settings{1, 10, {0.f, 0.f, 0.f}, 1.0f, 3, 0.0f};
    *out = s;
    return NX_OK;
}

void nx_scene_destroy(nx_scene* s)
{
    if (!s) return;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->stream_aux);
    for (int k = 0; k < ctx->buildStreamCount; k++) cudaStreamSynchronize(ctx->buildStreams[k]);
    for (uint32_t* c : s->buildCounterChunks) cudaFree(c);
    for (auto& m : s->meshes) if (!m.arenaOwned) { cudaFree(m.dTris); cudaFree(m.dTriData); cudaFree(m.dLeafTris); cudaFree(m.dShadeRec); nx_bvh8_free(ctx, &m.bvh); }
    for (nx_bump& slab : s->arena) cudaFree(slab.base);
    if (s->tlas.nodes) nx_bvh8_free(ctx, &s->tlas);
    if (s->merged.nodes) nx_bvh8_free(ctx, &s->merged);
    cudaFree(s->dMergedLeaf); cudaFree(s->dSlotInst); cudaFree(s->dInstInv);
    if (s->dTop && ctx->l2_persist_bytes) {   // drop the window that points at this scene's top-level block
        cudaStreamAttrValue attr; std::memset(&attr, 0, sizeof(attr));
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaStreamSetAttribute(ctx->stream_aux, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    for (size_t i = 0; i < s->textures.size(); i++) { cudaDestroyTextureObject(s->textures[i]); cudaFreeArray(s->textureArrays[i]); }
    cudaFree(s->dTextures);
    cudaFree(s->dTop); cudaFree(s->dShadeInst); cudaFree(s->dMeshes); cudaFree(s->dMaterials); cudaFree(s->dLights);
    if (s->hasHdr) { cudaDestroyTextureObject(s->hdr); cudaFreeArray(s->hdrArray); }
    cudaStreamSynchronize(ctx->stream);
    delete s;
}

int nx_scene_add_material(nx_scene* s, const nx_material* m)
{
    if (!s || !m) return NX_ERR_INVALID;
    s->materials.push_back(*m); s->dirtyMaterials = true; s->dirtyLights = true;
    return (int)s->materials.size() - 1;
}
int nx_scene_set_material(nx_scene* s, uint32_t idx, const nx_material* m)
{
    if (!s || !m || idx >= s->materials.size()) return NX_ERR_INVALID;
    s->materials[idx] = *m; s->dirtyMaterials = true; s->dirtyLights = true;
    return NX_OK;
}

struct PrebuiltBlas { const nx_bvh8_node* dNodes; uint32_t nodeCount; const uint32_t* dPrimIdx; nx_aabb bounds; };

// NX_PROFILE_SETUP=1: host-side time per section of add_mesh, summed over the scene, printed when the builds are collected
static double g_setupT[6] = {0, 0, 0, 0, 0, 0};
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Scene arena: every mesh's device arrays are carved out of large slabs (one cudaMalloc per 512 MB instead of six stream-ordered
// allocations per mesh, whose pool has to grow by gigabytes the first time a large scene is loaded: 2.3 s against 0.8 s warm for the
// 1,026-mesh scene).  Freed as a whole with the scene.
static int arena_alloc(nx_scene* s, size_t bytes, void** out)
{
    nx_ctx* ctx = s->ctx;
    bytes = (bytes + 255) & ~(size_t)255;
    if (s->arena.empty() || s->arena.back().used + bytes > s->arena.back().cap) {
        nx_bump slab; slab.cap = std::max<size_t>((size_t)512 << 20, bytes);
        NX_CUDA(ctx, cudaMalloc((void**)&slab.base, slab.cap));
        s->arena.push_back(slab);
    }
    nx_bump& b = s->arena.back();
    *out = b.base + b.used; b.used += bytes;
    return NX_OK;
}

// Host data -> device through the context's pinned staging ring (no stream synchronisation unless the ring wraps).
static int stage_upload(nx_ctx* ctx, cudaStream_t st, void* dst, const void* src, size_t bytes)
{
    constexpr size_t kRing = 64u << 20;
    if (!ctx->stagePinned) {
        if (cudaMallocHost((void**)&ctx->stagePinned, kRing) != cudaSuccess) { cudaGetLastError(); ctx->stagePinned = nullptr; ctx->stageBytes = 0; }
        else ctx->stageBytes = kRing;
    }
    if (bytes > ctx->stageBytes / 2) {                        // larger than the ring is worth: plain (staged by the driver) copy
        NX_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return NX_OK;
    }
    const size_t aligned = (bytes + 255) & ~(size_t)255;
    if (ctx->stageUsed + aligned > ctx->stageBytes) {         // wrap: everything staged so far must have left the ring
        for (int k = 0; k < ctx->buildStreamCount; k++) NX_CUDA(ctx, cudaStreamSynchronize(ctx->buildStreams[k]));
        NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stageUsed = 0;
    }
    char* p = ctx->stagePinned + ctx->stageUsed;
    ctx->stageUsed += aligned;
    std::memcpy(p, src, bytes);
    NX_CUDA(ctx, cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, st));
    return NX_OK;
}

static int build_stream(nx_ctx* ctx, size_t meshIdx, cudaStream_t* out)
{
    constexpr int kStreams = (int)(sizeof(ctx->buildStreams) / sizeof(ctx->buildStreams[0]));
    while (ctx->buildStreamCount < kStreams) {
        NX_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->buildStreams[ctx->buildStreamCount], cudaStreamNonBlocking));
        ctx->buildStreamCount++;
    }
    *out = ctx->buildStreams[meshIdx % kStreams];
    return NX_OK;
}

// Waits for the BLAS builds in flight and collects their node counts (one synchronisation and one read-back for all of them).
static int nxi_scene_flush_builds(nx_scene* s)
{
    if (!s->pendingBuilds) return NX_OK;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    const double tF = now_s();
    for (int k = 0; k < ctx->buildStreamCount; k++) NX_CUDA(ctx, cudaStreamSynchronize(ctx->buildStreams[k]));
    if (std::getenv("NX_PROFILE_SETUP")) {
        std::fprintf(stderr, "[nx setup] %.0f meshes: alloc %.3f s, staging %.3f s, bounds+sphere %.3f s, build issue %.3f s, final wait %.3f s\n",
                     g_setupT[5], g_setupT[0], g_setupT[1], g_setupT[2], g_setupT[3], now_s() - tF);
        for (double& v : g_setupT) v = 0;
    }
    NX_CUDA(ctx, cudaGetLastError());
    std::vector<uint32_t> host(8 * 1024);
    for (size_t c = 0; c < s->buildCounterChunks.size(); c++) {
        bool any = false;
        for (size_t k = c * 1024; k < std::min(s->meshes.size(), (c + 1) * 1024); k++) any = any || s->meshes[k].pending;
        if (!any) continue;
        NX_CUDA(ctx, cudaMemcpy(host.data(), s->buildCounterChunks[c], 4 * host.size(), cudaMemcpyDeviceToHost));
        for (size_t k = c * 1024; k < std::min(s->meshes.size(), (c + 1) * 1024); k++) {
            HostMesh& m = s->meshes[k];
            if (!m.pending) continue;
            const uint32_t* cnt = host.data() + 8 * (k - c * 1024);
            m.pending = false;
            if (m.prebuilt) {
                if (cnt[2]) NX_FAIL(ctx, NX_ERR_INVALID, "AddMesh(prebuilt) (mesh %zu): %u leaf slots refer to primitives outside [0, %u)", k, cnt[2], m.bvh.prim_count);
                continue;
            }
            if (cnt[1] != m.bvh.prim_count) NX_FAIL(ctx, NX_ERR_STATE, "BuildBVH8 (mesh %zu): collapse placed %u of %u primitives", k, cnt[1], m.bvh.prim_count);
            m.bvh.node_count = cnt[0];
        }
    }
    s->pendingBuilds = 0;
    return NX_OK;
}

// Issues the BLAS build of one mesh on its build stream (no host synchronisation; nxi_scene_flush_builds collects the result).
static int issue_blas_build(nx_scene* s, size_t meshIdx)
{
    nx_ctx* ctx = s->ctx;
    HostMesh& m = s->meshes[meshIdx];
    if (m.blasBuilt) return NX_OK;
    const uint32_t n = m.bvh.prim_count;
    cudaStream_t st = nullptr;
    int rc = build_stream(ctx, meshIdx, &st); if (rc) return rc;
    uint32_t* counters = s->buildCounterChunks[meshIdx / 1024] + 8 * (meshIdx % 1024);
    rc = arena_alloc(s, 48 * (size_t)n, (void**)&m.dLeafTris); if (rc) return rc;
    // temporaries: this build stream's workspace (grown when a larger mesh arrives; builds on one stream run one after the other);
    // outputs: a region of the arena sized for the worst case (ceil((4n - 1) / 7) nodes, BVHBuilder.cpp:184-186)
    nx_bump& ws = ctx->buildWsStore[meshIdx % (sizeof(ctx->buildWsStore) / sizeof(ctx->buildWsStore[0]))];
    const size_t need = nxi_build_workspace_bytes(n);
    if (ws.cap < need) {
        NX_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(ws.base); ws.base = nullptr; ws.cap = 0;
        NX_CUDA(ctx, cudaMalloc((void**)&ws.base, need + need / 2));
        ws.cap = need + need / 2;
    }
    nx_bump outputs; outputs.cap = (((size_t)4 * n - 1 + 6) / 7) * sizeof(nx_bvh8_node) + 4 * (size_t)n + 1024;
    rc = arena_alloc(s, outputs.cap, (void**)&outputs.base); if (rc) return rc;
    const nx_aabb box = m.bvh.bounds;                         // = the builder's scene bounds (exact min / max of the same floats)
    rc = nxi_build_bvh8_async(ctx, st, m.dTris, n, 1, ctx->scene_blas_speed /* Mesh::Mesh: prioritizeSpeed = true */, counters, &m.bvh, &ws, &outputs);
    if (rc) return rc;
    m.bvh.bounds = box;
    const int grid = (int)std::min<uint32_t>(div_up(n, 256), (uint32_t)ctx->sm_count * 8u);
    leaf_triangles_kernel<<<grid, 256, 0, st>>>(m.dTris, m.bvh.prim_idx, n, m.dLeafTris, nullptr);
    NX_CUDA(ctx, cudaGetLastError());
    m.blasBuilt = true; m.pending = true; s->pendingBuilds++;
    return NX_OK;
}

// Mesh::Mesh (N/Assets/Mesh.h:29-40).  With `pre` the BLAS is taken from the caller (device arrays, copied) instead of built.
// Nothing in here waits for the GPU: uploads go through the pinned ring on one of the context's build streams.  The BLAS itself is
// built when it is first needed - by Scene::Update for the meshes that keep a BLAS of their own (the others end up in the scene's
// merged world-space BLAS), or by a query (nx_scene_mesh_bvh) - again without waiting, all meshes of a scene pipelined over the build
// streams and collected once.  A scene of 1,026 meshes took 4.25 s to set up with one synchronising build per mesh (BENCH_r01
// scene_setup_s) for 0.36 s of GPU work.
static int add_mesh(nx_scene* s, const nx_triangle* tris, const nx_triangle_data* data, uint32_t n, uint32_t materialIdx, const PrebuiltBlas* pre)
{
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    const double tA = now_s();
    const size_t meshIdx = s->meshes.size();
    cudaStream_t st = nullptr;
    int rc = build_stream(ctx, meshIdx, &st); if (rc) return rc;
    if (meshIdx / 1024 >= s->buildCounterChunks.size()) {
        uint32_t* chunk = nullptr;
        NX_CUDA(ctx, cudaMalloc((void**)&chunk, 4 * 8 * 1024));
        s->buildCounterChunks.push_back(chunk);
    }
    HostMesh m; m.materialIdx = materialIdx; m.arenaOwned = true;
    uint32_t* counters = s->buildCounterChunks[meshIdx / 1024] + 8 * (meshIdx % 1024);
    rc = arena_alloc(s, 36 * (size_t)n, (void**)&m.dTris); if (rc) return rc;
    rc = arena_alloc(s, 96 * (size_t)n, (void**)&m.dTriData); if (rc) return rc;
    rc = arena_alloc(s, 16 * NX_SHADE_REC_F4 * (size_t)n, (void**)&m.dShadeRec); if (rc) return rc;
    const double tB = now_s();
    rc = stage_upload(ctx, st, m.dTris, tris, 36 * (size_t)n); if (rc) return rc;
    const int grid = (int)std::min<uint32_t>(div_up(n, 256), (uint32_t)ctx->sm_count * 8u);
    if (data) { rc = stage_upload(ctx, st, m.dTriData, data, 96 * (size_t)n); if (rc) return rc; }
    else default_tridata_kernel<<<grid, 256, 0, st>>>(m.dTris, n, m.dTriData);
    shade_record_kernel<<<grid, 256, 0, st>>>(m.dTris, m.dTriData, n, m.dShadeRec);
    const double tC = now_s();
    nx_aabb box;
    mesh_bounds_sphere(tris, n, &box, m.sphere);
    const double tD = now_s();
    m.bvh.prim_count = n; m.bvh.bounds = box;
    s->pendingBuilds++;                                        // the uploads: Update must wait for them even when it builds nothing
    if (pre) {
        rc = arena_alloc(s, 48 * (size_t)n, (void**)&m.dLeafTris); if (rc) return rc;
        rc = arena_alloc(s, sizeof(nx_bvh8_node) * (size_t)pre->nodeCount, (void**)&m.bvh.nodes); if (rc) return rc;
        rc = arena_alloc(s, 4 * (size_t)n, (void**)&m.bvh.prim_idx); if (rc) return rc;
        // the caller's arrays were produced on other streams: the caller synchronises before handing them over (nx_scene_build_blas does)
        NX_CUDA(ctx, cudaMemcpyAsync(m.bvh.nodes, pre->dNodes, sizeof(nx_bvh8_node) * (size_t)pre->nodeCount, cudaMemcpyDeviceToDevice, st));
        NX_CUDA(ctx, cudaMemcpyAsync(m.bvh.prim_idx, pre->dPrimIdx, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        m.bvh.node_count = pre->nodeCount; m.bvh.bounds = pre->bounds;
        NX_CUDA(ctx, cudaMemsetAsync(counters, 0, 32, st));
        leaf_triangles_kernel<<<grid, 256, 0, st>>>(m.dTris, m.bvh.prim_idx, n, m.dLeafTris, counters + 2);
        m.blasBuilt = true; m.pending = true; m.prebuilt = true;   // Update waits for the copies and reads the index check's verdict
    }
    NX_CUDA(ctx, cudaGetLastError());
    const double tE = now_s();
    g_setupT[0] += tB - tA; g_setupT[1] += tC - tB; g_setupT[2] += tD - tC; g_setupT[3] += tE - tD; g_setupT[5] += 1;
    s->meshes.push_back(m);
    return (int)s->meshes.size() - 1;
}

int nx_scene_add_mesh(nx_scene* s, const nx_triangle* tris, const nx_triangle_data* data, uint32_t n, uint32_t materialIdx)
{
    if (!s || !tris || !n) return NX_ERR_INVALID;
    return add_mesh(s, tris, data, n, materialIdx, nullptr);
}

int nx_scene_add_mesh_prebuilt(nx_scene* s, const nx_triangle* tris, const nx_triangle_data* data, uint32_t n, uint32_t materialIdx,
                               const nx_bvh8_node* dNodes, uint32_t nodeCount, const uint32_t* dPrimIdx, const nx_aabb* bounds)
{
    if (!s || !tris || !n || !dNodes || !dPrimIdx || !bounds) return NX_ERR_INVALID;
    if (!nodeCount || (size_t)nodeCount > ((size_t)4 * n - 1 + 6) / 7) NX_FAIL(s->ctx, NX_ERR_INVALID, "AddMesh(prebuilt): %u nodes for %u triangles", nodeCount, n);
    const PrebuiltBlas pre{dNodes, nodeCount, dPrimIdx, *bounds};
    return add_mesh(s, tris, data, n, materialIdx, &pre);
}

int nx_scene_build_blas(nx_ctx* ctx, const nx_triangle* hostTris, uint32_t n, nx_bvh8* out)
{
    if (!ctx || !hostTris || !n || !out) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    float* d = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&d, 36 * (size_t)n, ctx->stream));
    NX_CUDA(ctx, cudaMemcpyAsync(d, hostTris, 36 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    const int rc = nxi_build_bvh8(ctx, d, n, 1, ctx->scene_blas_speed, out);   // synchronises the stream
    cudaFreeAsync(d, ctx->stream);
    return rc;
}

int nx_scene_mesh_bounds(nx_scene* s, uint32_t meshIdx, nx_aabb* out)
{
    if (!s || !out || meshIdx >= s->meshes.size()) return NX_ERR_INVALID;
    *out = s->meshes[meshIdx].bvh.bounds; return NX_OK;
}
int nx_scene_mesh_bvh(nx_scene* s, uint32_t meshIdx, nx_bvh8* out)
{
    if (!s || !out || meshIdx >= s->meshes.size()) return NX_ERR_INVALID;
    DeviceGuard guard(s->ctx->device);
    { int rc = issue_blas_build(s, meshIdx); if (rc) return rc; rc = nxi_scene_flush_builds(s); if (rc) return rc; }
    *out = s->meshes[meshIdx].bvh; return NX_OK;
}

int nx_scene_add_instance_matrix(nx_scene* s, uint32_t meshIdx, int32_t materialIdx, const float mtx[16])
{
    if (!s || !mtx || meshIdx >= s->meshes.size()) return NX_ERR_INVALID;
    HostInstance inst; inst.meshIdx = meshIdx;
    inst.materialIdx = materialIdx >= 0 ? (uint32_t)materialIdx : s->meshes[meshIdx].materialIdx;
    M4 m, inv; std::memcpy(m.c, mtx, 64);
    m4_inverse(m, inv);
    std::memcpy(inst.m, m.c, 64); std::memcpy(inst.inv, inv.c, 64);
    inst.bounds = transformed_bounds(inst.m, s->meshes[meshIdx].bvh.bounds);
    s->meshes[meshIdx].useCount++;
    s->instances.push_back(inst); s->dirtyInstances = true; s->dirtyLights = true;
    return (int)s->instances.size() - 1;
}
int nx_scene_add_instance(nx_scene* s, uint32_t meshIdx, int32_t materialIdx, const float pos[3], const float rot[3], const float scale[3])
{
    if (!s || !pos || !rot || !scale) return NX_ERR_INVALID;
    M4 m = compose_trs(pos, rot, scale);
    return nx_scene_add_instance_matrix(s, meshIdx, materialIdx, m.c);
}
int nx_scene_set_instance_transform(nx_scene* s, uint32_t idx, const float pos[3], const float rot[3], const float scale[3])
{
    if (!s || idx >= s->instances.size()) return NX_ERR_INVALID;
    HostInstance& inst = s->instances[idx];
    M4 m = compose_trs(pos, rot, scale), inv; m4_inverse(m, inv);
    std::memcpy(inst.m, m.c, 64); std::memcpy(inst.inv, inv.c, 64);
    inst.bounds = transformed_bounds(inst.m, s->meshes[inst.meshIdx].bvh.bounds);
    inst.dynamic = true;        // an instance that has been moved keeps a BLAS of its own from now on: further moves only rebuild the TLAS
    s->dirtyInstances = true;
    return NX_OK;
}

// MeshInstance::AssignMaterial + Scene::InvalidateMeshInstance (MeshInstance.h:35, Scene.cpp:109-112); < 0 = the mesh's own material
int nx_scene_set_instance_material(nx_scene* s, uint32_t idx, int32_t materialIdx)
{
    if (!s || idx >= s->instances.size()) return NX_ERR_INVALID;
    if (materialIdx >= 0 && (size_t)materialIdx >= s->materials.size()) NX_FAIL(s->ctx, NX_ERR_INVALID, "AssignMaterial: material %d of %zu", materialIdx, s->materials.size());
    HostInstance& inst = s->instances[idx];
    inst.materialIdx = materialIdx >= 0 ? (uint32_t)materialIdx : s->meshes[inst.meshIdx].materialIdx;
    s->dirtyInstances = true; s->dirtyLights = true;          // an emissive material turns the instance into a light (Scene.cpp:157-219)
    return NX_OK;
}
// MeshInstance::GetTransfromationMatrix / GetBounds (MeshInstance.h:36-53): row-major 4x4; world AABB of the mesh box's eight corners
int nx_scene_instance_matrix(nx_scene* s, uint32_t idx, float out[16])
{
    if (!s || !out || idx >= s->instances.size()) return NX_ERR_INVALID;
    std::memcpy(out, s->instances[idx].m, 64); return NX_OK;
}
int nx_scene_instance_bounds(nx_scene* s, uint32_t idx, nx_aabb* out)
{
    if (!s || !out || idx >= s->instances.size()) return NX_ERR_INVALID;
    *out = s->instances[idx].bounds; return NX_OK;
}

int nx_scene_add_light(nx_scene* s, const nx_light* l)
{
    if (!s || !l) return NX_ERR_INVALID;
    s->userLights.push_back(*l); s->dirtyLights = true;
    return (int)s->userLights.size() - 1;
}
// Scene::InvalidateLight after editing GetLights()[idx] (Scene.cpp:114-117), Scene::RemoveLight (Scene.cpp:129-132; later lights move down)
int nx_scene_set_light(nx_scene* s, uint32_t idx, const nx_light* l)
{
    if (!s || !l || idx >= s->userLights.size()) return NX_ERR_INVALID;
    s->userLights[idx] = *l; s->dirtyLights = true;
    return NX_OK;
}
int nx_scene_remove_light(nx_scene* s, uint32_t idx)
{
    if (!s || idx >= s->userLights.size()) return NX_ERR_INVALID;
    s->userLights.erase(s->userLights.begin() + idx); s->dirtyLights = true;
    return NX_OK;
}
int nx_scene_light_count(nx_scene* s) { return s ? (int)s->userLights.size() : NX_ERR_INVALID; }
// AssetManager::AddTexture + Texture::ToDevice (src/Assets/AssetManager.h:31, src/Assets/Texture.cpp:12-46): RGBA8 (normalised float
// reads, optional sRGB decode in the sampler) or RGBA32F pixels into a CUDA array behind a texture object with wrap addressing,
// linear filtering and normalised coordinates.  Returns the texture index that nx_material::*_map refers to.
int nx_scene_add_texture(nx_scene* s, const void* rgba, uint32_t w, uint32_t h, int isHdr, int srgb)
{
    if (!s || !rgba || !w || !h) return NX_ERR_INVALID;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    const cudaChannelFormatDesc desc = isHdr ? cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat) : cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
    cudaArray_t arr = nullptr;
    NX_CUDA(ctx, cudaMallocArray(&arr, &desc, w, h));
    const size_t pitch = (size_t)w * (isHdr ? 16 : 4);
    NX_CUDA(ctx, cudaMemcpy2DToArray(arr, 0, 0, rgba, pitch, pitch, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res; std::memset(&res, 0, sizeof(res));
    res.resType = cudaResourceTypeArray; res.res.array.array = arr;
    cudaTextureDesc tex; std::memset(&tex, 0, sizeof(tex));
    tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeWrap;
    tex.sRGB = (srgb && !isHdr) ? 1 : 0;
    tex.filterMode = cudaFilterModeLinear;
    tex.readMode = isHdr ? cudaReadModeElementType : cudaReadModeNormalizedFloat;
    tex.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    NX_CUDA(ctx, cudaCreateTextureObject(&obj, &res, &tex, nullptr));
    s->textures.push_back(obj); s->textureArrays.push_back(arr); s->dirtyTextures = true;
    return (int)s->textures.size() - 1;
}

// Camera::OnResize (src/Scene/Camera.cpp:118-128): the resolution only feeds the camera record (aspect ratio, pixel grid); nothing is rebuilt.
int nx_scene_set_resolution(nx_scene* s, uint32_t width, uint32_t height)
{
    if (!s || !width || !height) return NX_ERR_INVALID;
    s->width = width; s->height = height;
    return NX_OK;
}
int nx_scene_set_camera(nx_scene* s, const nx_camera* c) { if (!s || !c) return NX_ERR_INVALID; s->camera = *c; return NX_OK; }
int nx_scene_set_render_settings(nx_scene* s, const nx_render_settings* r)
{
    if (!s || !r || r->path_length < 1 || r->path_length > 255) return NX_ERR_INVALID;
    s->settings = *r; return NX_OK;
}

// Texture::ToDevice for the HDR environment (src/Assets/Texture.cpp:12-46): RGBA32F array, wrap, linear, normalised coords.
int nx_scene_set_hdr_map(nx_scene* s, const float* rgba, uint32_t w, uint32_t h)
{
    if (!s || !rgba || !w || !h) return NX_ERR_INVALID;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    if (s->hasHdr) { cudaStreamSynchronize(ctx->stream); cudaDestroyTextureObject(s->hdr); cudaFreeArray(s->hdrArray); s->hasHdr = false; }
    cudaChannelFormatDesc desc = cudaCreateChannelDesc(32, 32, 32, 32, cudaChannelFormatKindFloat);
    NX_CUDA(ctx, cudaMallocArray(&s->hdrArray, &desc, w, h));
    NX_CUDA(ctx, cudaMemcpy2DToArray(s->hdrArray, 0, 0, rgba, 16 * (size_t)w, 16 * (size_t)w, h, cudaMemcpyHostToDevice));
    cudaResourceDesc res; std::memset(&res, 0, sizeof(res)); res.resType = cudaResourceTypeArray; res.res.array.array = s->hdrArray;
    cudaTextureDesc tex; std::memset(&tex, 0, sizeof(tex));
    tex.addressMode[0] = tex.addressMode[1] = cudaAddressModeWrap; tex.filterMode = cudaFilterModeLinear;
    tex.readMode = cudaReadModeElementType; tex.normalizedCoords = 1; tex.sRGB = 0;
    NX_CUDA(ctx, cudaCreateTextureObject(&s->hdr, &res, &tex, nullptr));
    s->hasHdr = true;
    return NX_OK;
}

int nx_scene_update(nx_scene* s)
{
    if (!s) return NX_ERR_INVALID;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    { const int rc = nxi_scene_flush_builds(s); if (rc) return rc; }
    if (s->instances.empty()) NX_FAIL(ctx, NX_ERR_STATE, "Scene::Update: the scene has no mesh instances");
    if (s->materials.empty()) NX_FAIL(ctx, NX_ERR_STATE, "Scene::Update: the scene has no materials");
    for (const auto& i : s->instances) if (i.materialIdx >= s->materials.size()) NX_FAIL(ctx, NX_ERR_INVALID, "instance refers to material %u of %zu", i.materialIdx, s->materials.size());

    for (const auto& m : s->materials)
        for (int32_t id : {m.base_color_map, m.emissive_map, m.normal_map, m.roughness_map, m.metalness_map, m.metallic_roughness_map})
            if (id < -1 || id >= (int32_t)s->textures.size()) NX_FAIL(ctx, NX_ERR_INVALID, "a material refers to texture %d of %zu", id, s->textures.size());
    if (s->dirtyMaterials) {
        std::vector<DMaterial> dmat(s->materials.size());
        for (size_t i = 0; i < dmat.size(); i++) { dmat[i].m = s->materials[i]; dmat[i].pad = 0; }
        int rc = upload_vec(ctx, &s->dMaterials, dmat); if (rc) return rc;
    }
    if (s->dirtyTextures) { int rc = upload_vec(ctx, &s->dTextures, s->textures); if (rc) return rc; s->dirtyTextures = false; }

    if (s->dirtyInstances)
    {
        int rc;
        // ---- which instances share the merged world-space BLAS: mesh used by this instance only, never moved, merging on, and not in
        // the NexusBVH-identical mode (whose point is reference-identical trees)
        std::vector<uint32_t> mergeSet;
        if (ctx->merge_instances && ctx->scene_collapse != NX_COLLAPSE_REFERENCE_GPU)
            for (uint32_t i = 0; i < s->instances.size(); i++)
                if (!s->instances[i].dynamic && s->meshes[s->instances[i].meshIdx].useCount == 1 &&
                    s->meshes[s->instances[i].meshIdx].bvh.prim_count >= ctx->merge_min_prims) mergeSet.push_back(i);
        std::vector<uint8_t> isMerged(s->instances.size(), 0);
        for (uint32_t i : mergeSet) isMerged[i] = 1;
        // every other instance needs its mesh's own BLAS: issue the missing builds now, they run while the merged BLAS is prepared
        for (uint32_t i = 0; i < s->instances.size(); i++)
            if (!isMerged[i]) { rc = issue_blas_build(s, s->instances[i].meshIdx); if (rc) return rc; }

        if (mergeSet != s->mergedInstances)
        {
            if (s->merged.nodes) { nx_bvh8_free(ctx, &s->merged); s->merged = nx_bvh8{}; }
            if (s->dMergedLeaf) { cudaFreeAsync(s->dMergedLeaf, ctx->stream); s->dMergedLeaf = nullptr; }
            s->mergedInstances = mergeSet; s->mergedFirst.clear();
            if (!mergeSet.empty())
            {
                std::vector<BakeSrc> src(mergeSet.size());
                uint64_t total = 0;
                for (size_t j = 0; j < src.size(); j++) {
                    const HostInstance& h = s->instances[mergeSet[j]];
                    const HostMesh& hm = s->meshes[h.meshIdx];
                    src[j].tris = hm.dTris; std::memcpy(src[j].m, h.m, 48);
                    src[j].first = (uint32_t)total; src[j].count = hm.bvh.prim_count; src[j].inst = mergeSet[j];
                    src[j].mag = std::max(std::fabs(h.m[3]), std::max(std::fabs(h.m[7]), std::fabs(h.m[11])));
                    s->mergedFirst.push_back((uint32_t)total);
                    total += hm.bvh.prim_count;
                }
                if (total > 0x07ffffffull) NX_FAIL(ctx, NX_ERR_INVALID, "Scene::Update: %llu primitives in the merged BLAS (limit 2^27)", (unsigned long long)total);
                const uint32_t n = (uint32_t)total;
                BakeSrc* dSrc = nullptr; float* dWorld = nullptr; uint32_t* dSrcOf = nullptr;
                rc = upload_vec(ctx, &dSrc, src); if (rc) return rc;
                NX_CUDA(ctx, cudaMallocAsync((void**)&dWorld, 24 * (size_t)n, ctx->stream));
                NX_CUDA(ctx, cudaMallocAsync((void**)&dSrcOf, 4 * (size_t)n, ctx->stream));
                NX_CUDA(ctx, cudaMallocAsync((void**)&s->dMergedLeaf, 48 * (size_t)n, ctx->stream));
                // the geometry uploads ran on the build streams
                for (int k = 0; k < ctx->buildStreamCount; k++) NX_CUDA(ctx, cudaStreamSynchronize(ctx->buildStreams[k]));
                const int grid = (int)std::min<uint32_t>(div_up(n, 256), (uint32_t)ctx->sm_count * 8u);
                bake_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(dSrc, (uint32_t)src.size(), n, dWorld, dSrcOf);
                // 64-bit Morton keys: the merged BLAS spans the whole scene, and 10 bits per axis would put many of its small triangles
                // into the same cell.  Built over boxes (primType 0) with the triangle BLASes' leaf size.
                rc = nxi_build_bvh8(ctx, dWorld, n, 0, /*prioritizeSpeed=*/0, &s->merged, ctx->scene_max_leaf_prims); if (rc) return rc;
                merged_leaf_kernel<<<grid, 256, 0, ctx->stream>>>(s->merged.prim_idx, n, dSrcOf, dSrc, s->dMergedLeaf);
                NX_CUDA(ctx, cudaGetLastError());
                cudaFreeAsync(dWorld, ctx->stream); cudaFreeAsync(dSrcOf, ctx->stream); cudaFreeAsync(dSrc, ctx->stream);
            }
        }
        rc = nxi_scene_flush_builds(s); if (rc) return rc;

        // device mesh table
        std::vector<DMesh> dm(s->meshes.size());
        for (size_t i = 0; i < dm.size(); i++) { dm[i].tris = s->meshes[i].dTris; dm[i].tridata = s->meshes[i].dTriData; dm[i].shade = s->meshes[i].dShadeRec; dm[i].primCount = s->meshes[i].bvh.prim_count; dm[i].pad = 0; }
        rc = upload_vec(ctx, &s->dMeshes, dm); if (rc) return rc;

        // shading records in instance order
        std::vector<DShadeInst> si(s->instances.size());
        std::vector<nx_aabb> bounds(s->instances.size());
        for (size_t i = 0; i < si.size(); i++) {
            const HostInstance& h = s->instances[i];
            std::memcpy(&si[i].m0, h.m, 48); std::memcpy(&si[i].i0, h.inv, 48);
            si[i].meshIdx = h.meshIdx; si[i].materialIdx = h.materialIdx; si[i].shade = s->meshes[h.meshIdx].dShadeRec;
            bounds[i] = h.bounds;
        }
        rc = upload_vec(ctx, &s->dShadeInst, si); if (rc) return rc;
        {   // instance id -> world -> object rows, for the object-space triangle tests inside the merged BLAS
            std::vector<float4> ii(3 * s->instances.size());
            for (size_t i = 0; i < s->instances.size(); i++) std::memcpy(&ii[3 * i], s->instances[i].inv, 48);
            rc = upload_vec(ctx, &s->dInstInv, ii); if (rc) return rc;
        }

        // TLAS entries: the instances that keep their own BLAS, then the merged BLAS as one more primitive
        std::vector<uint32_t> entryInst;
        for (uint32_t i = 0; i < s->instances.size(); i++) if (!isMerged[i]) entryInst.push_back(i);
        const size_t nOwn = entryInst.size();

        // world-space bounding sphere of every such instance's geometry (conservative: radius and centre rounding padded)
        std::vector<float4> spheres(nOwn);
        for (size_t e = 0; e < nOwn; e++) {
            const HostInstance& h = s->instances[entryInst[e]];
            const HostMesh& hm = s->meshes[h.meshIdx];
            double c[3];
            for (int r = 0; r < 3; r++) c[r] = (double)h.m[4 * r] * hm.sphere[0] + (double)h.m[4 * r + 1] * hm.sphere[1] + (double)h.m[4 * r + 2] * hm.sphere[2] + (double)h.m[4 * r + 3];
            const double rad = hm.sphere[3] * max_stretch(h.m) * (1.0 + 1e-5) + 1e-30;
            spheres[e] = make_float4((float)c[0], (float)c[1], (float)c[2], (float)(rad * (1.0 + 1e-6)) + 1e-6f * (float)(std::fabs(c[0]) + std::fabs(c[1]) + std::fabs(c[2])));
        }
        std::vector<nx_aabb> ebounds(nOwn);
        for (size_t e = 0; e < nOwn; e++) ebounds[e] = bounds[entryInst[e]];
        // The reference bounds an instance by the world AABB of the eight transformed corners of its mesh AABB
        // (MeshInstance::GetBounds, N/Scene/MeshInstance.h:42-66), which for a rotated object is up to sqrt(3) too wide.  Outside
        // the NexusBVH-identical mode the TLAS is built over that box clipped to the bounding sphere's box: still conservative,
        // and the node test then rejects what the per-instance sphere test would otherwise have to.
        if (ctx->scene_collapse != NX_COLLAPSE_REFERENCE_GPU && nOwn)
        {
            // ... and to the box of the transformed vertices themselves (padded for the rounding of the transform)
            std::vector<InstGeom> ig(nOwn);
            for (size_t e = 0; e < nOwn; e++) {
                const HostInstance& h = s->instances[entryInst[e]];
                ig[e].tris = s->meshes[h.meshIdx].dTris; ig[e].triCount = s->meshes[h.meshIdx].bvh.prim_count; std::memcpy(ig[e].m, h.m, 48);
            }
            InstGeom* dIg = nullptr; float* dGb = nullptr;
            rc = upload_vec(ctx, &dIg, ig); if (rc) return rc;
            NX_CUDA(ctx, cudaMallocAsync((void**)&dGb, 24 * ig.size(), ctx->stream));
            instance_geometry_bounds_kernel<<<(uint32_t)ig.size(), 256, 0, ctx->stream>>>(dIg, dGb);
            std::vector<float> gb(6 * ig.size());
            NX_CUDA(ctx, cudaMemcpyAsync(gb.data(), dGb, 24 * ig.size(), cudaMemcpyDeviceToHost, ctx->stream));
            NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFreeAsync(dIg, ctx->stream); cudaFreeAsync(dGb, ctx->stream);
            for (size_t i = 0; i < ebounds.size(); i++)
                for (int a = 0; a < 3; a++) {
                    const float ext = gb[6 * i + 3 + a] - gb[6 * i + a], mag = std::max(std::fabs(gb[6 * i + a]), std::fabs(gb[6 * i + 3 + a]));
                    const float pad = 1.0e-5f * ext + 4.0e-6f * mag + 1.0e-30f;
                    const float nlo = std::max(ebounds[i].bmin[a], gb[6 * i + a] - pad), nhi = std::min(ebounds[i].bmax[a], gb[6 * i + 3 + a] + pad);
                    if (nlo <= nhi) { ebounds[i].bmin[a] = nlo; ebounds[i].bmax[a] = nhi; }
                }
            for (size_t i = 0; i < ebounds.size(); i++) {
                const float4 sp = spheres[i];
                const float c3[3] = {sp.x, sp.y, sp.z};
                for (int a = 0; a < 3; a++) {
                    const float lo = std::nextafterf(c3[a] - sp.w, -INFINITY), hi = std::nextafterf(c3[a] + sp.w, INFINITY);
                    const float nlo = std::max(ebounds[i].bmin[a], lo), nhi = std::min(ebounds[i].bmax[a], hi);
                    if (nlo <= nhi) { ebounds[i].bmin[a] = nlo; ebounds[i].bmax[a] = nhi; }
                }
            }
        }
        const bool hasMerged = !s->mergedInstances.empty();
        if (hasMerged) { entryInst.push_back(NX_INVALID); ebounds.push_back(s->merged.bounds); }
        // same entries in the same order as the TLAS in hand: it can be refitted (nx_ctx_set_tlas_refit) instead of rebuilt
        const bool refit = ctx->tlas_refit && s->tlas.nodes && entryInst == s->tlasEntryInst;
        s->tlasEntryInst = entryInst;

        // Scene::BuildTLAS (src/Scene/Scene.cpp:65-78): BuildBVH8<AABB> over the entry bounds, default config (64-bit keys)
        nx_aabb* dBounds = nullptr;
        rc = upload_vec(ctx, &dBounds, ebounds); if (rc) return rc;
        if (refit) { rc = nxi_refit_bvh8(ctx, &s->tlas, dBounds); s->tlasRefits++; }
        else {
            if (s->tlas.nodes) nx_bvh8_free(ctx, &s->tlas);
            rc = nxi_build_bvh8(ctx, dBounds, (uint32_t)ebounds.size(), 0, 0, &s->tlas);
            s->tlasBuilds++;
        }
        cudaFreeAsync(dBounds, ctx->stream);
        if (rc) return rc;

        // traversal records and instance ids in TLAS leaf order
        std::vector<uint32_t> order(ebounds.size());
        NX_CUDA(ctx, cudaMemcpyAsync(order.data(), s->tlas.prim_idx, 4 * order.size(), cudaMemcpyDeviceToHost, ctx->stream));
        NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<DTravInst> ti(order.size());
        std::vector<uint32_t> slotInst(order.size());
        s->mergedSlot = NX_INVALID;
        for (size_t k = 0; k < order.size(); k++) {
            const uint32_t e = order[k];
            slotInst[k] = entryInst[e];
            if (entryInst[e] == NX_INVALID) {
                // the merged BLAS: identity transform, a sphere that is never culled
                s->mergedSlot = (uint32_t)k;
                ti[k].sphere = make_float4(0.f, 0.f, 0.f, INFINITY);
                ti[k].r0 = make_float4(1.f, 0.f, 0.f, 0.f); ti[k].r1 = make_float4(0.f, 1.f, 0.f, 0.f); ti[k].r2 = make_float4(0.f, 0.f, 1.f, 0.f);
                ti[k].nodes = (const float4*)s->merged.nodes; ti[k].ltris = s->dMergedLeaf;
                continue;
            }
            const HostInstance& h = s->instances[entryInst[e]];
            std::memcpy(&ti[k].r0, h.inv, 48);
            ti[k].sphere = spheres[e];
            ti[k].nodes = (const float4*)s->meshes[h.meshIdx].bvh.nodes; ti[k].ltris = s->meshes[h.meshIdx].dLeafTris;
        }
        rc = upload_vec(ctx, &s->dSlotInst, slotInst); if (rc) return rc;
        // top-level block: TLAS nodes + traversal records, one allocation, L2-persisting window on the two trace streams
        {
            const size_t nodeBytes = 80 * (size_t)s->tlas.node_count, recBytes = sizeof(DTravInst) * ti.size();
            if (s->dTop) { cudaFreeAsync(s->dTop, ctx->stream); s->dTop = nullptr; }
            s->topBytes = nodeBytes + recBytes;
            NX_CUDA(ctx, cudaMallocAsync(&s->dTop, s->topBytes, ctx->stream));
            s->dTopNodes = (const float4*)s->dTop;
            s->dTravInst = (DTravInst*)((char*)s->dTop + nodeBytes);
            NX_CUDA(ctx, cudaMemcpyAsync(s->dTop, s->tlas.nodes, nodeBytes, cudaMemcpyDeviceToDevice, ctx->stream));
            NX_CUDA(ctx, cudaMemcpyAsync(s->dTravInst, ti.data(), recBytes, cudaMemcpyHostToDevice, ctx->stream));
            NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // ti is a local
            if (ctx->l2_persist_bytes) {
                cudaStreamAttrValue attr; std::memset(&attr, 0, sizeof(attr));
                attr.accessPolicyWindow.base_ptr = s->dTop;
                attr.accessPolicyWindow.num_bytes = std::min(s->topBytes, ctx->l2_window_max);
                attr.accessPolicyWindow.hitRatio = s->topBytes <= ctx->l2_persist_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)s->topBytes;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
                cudaStreamSetAttribute(ctx->stream_aux, cudaStreamAttributeAccessPolicyWindow, &attr);
                cudaGetLastError();   // a hint: failure to set it is not an error
            }
        }
    }

    if (s->dirtyLights || s->dirtyInstances || s->dirtyMaterials)
    {
        // punctual lights first, then one MESH light per emissive instance (Scene::UpdateSceneLighting, Scene.cpp:157-219)
        s->lights.clear();
        for (const nx_light& l : s->userLights) {
            if (l.type == NX_LIGHT_MESH && l.instance >= s->instances.size())
                NX_FAIL(ctx, NX_ERR_INVALID, "a MESH light refers to instance %u of %zu", l.instance, s->instances.size());
            DLight d{}; d.type = l.type;
            d.px = l.position[0]; d.py = l.position[1]; d.pz = l.position[2];
            d.dx = l.direction[0]; d.dy = l.direction[1]; d.dz = l.direction[2];
            d.cr = l.color[0]; d.cg = l.color[1]; d.cb = l.color[2]; d.intensity = l.intensity; d.instance = l.instance;
            s->lights.push_back(d);
        }
        for (uint32_t mat = 0; mat < s->materials.size(); mat++) {
            if (!material_emits(s->materials[mat])) continue;
            for (uint32_t j = 0; j < s->instances.size(); j++)
                if (s->instances[j].materialIdx == mat) { DLight d{}; d.type = NX_LIGHT_MESH; d.instance = j; s->lights.push_back(d); }
        }
        int rc = upload_vec(ctx, &s->dLights, s->lights); if (rc) return rc;
    }
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    s->dirtyInstances = s->dirtyMaterials = s->dirtyLights = false;
    return NX_OK;
}

// ---- exports in the reference's device layouts (parity with the reference arm) ----
int nx_scene_export_instances(nx_scene* s, void* out160, uint32_t* outCount)
{
    if (!s || !outCount) return NX_ERR_INVALID;
    *outCount = (uint32_t)s->instances.size();
    if (!out160) return NX_OK;
    uint8_t* o = (uint8_t*)out160;
    for (size_t i = 0; i < s->instances.size(); i++, o += 160) {
        const HostInstance& h = s->instances[i];
        std::memcpy(o, &h.meshIdx, 4); std::memcpy(o + 4, &h.materialIdx, 4);
        std::memcpy(o + 8, h.m, 64); std::memcpy(o + 72, h.inv, 64); std::memcpy(o + 136, &h.bounds, 24);
    }
    return NX_OK;
}
int nx_scene_export_camera(nx_scene* s, void* out88)
{
    if (!s || !out88) return NX_ERR_INVALID;
    DCamera c = nxi_camera_to_device(s->camera, s->width, s->height);
    std::memcpy(out88, &c, 88);
    return NX_OK;
}
// The host arithmetic behind nx_scene_add_instance / nx_scene_set_instance_transform and nx_scene_export_camera, callable without a
// scene or a GPU: the D_MeshInstance (160 B) and D_Camera (88 B) records for the given inputs, through the very functions the scene
// code uses (compose_trs, m4_inverse, transformed_bounds, nxi_camera_to_device).  Lets the CPU tests pin them against the
// reference's own host code (MeshInstance::ToDevice, Camera::ToDevice).
int nx_host_instance_record(const float pos[3], const float rotDeg[3], const float scale[3], const nx_aabb* meshBounds, uint32_t meshIdx,
                            uint32_t materialIdx, void* out160)
{
    if (!pos || !rotDeg || !scale || !meshBounds || !out160) return NX_ERR_INVALID;
    M4 m = compose_trs(pos, rotDeg, scale), inv; m4_inverse(m, inv);
    const nx_aabb b = transformed_bounds(m.c, *meshBounds);
    uint8_t* o = (uint8_t*)out160;
    std::memcpy(o, &meshIdx, 4); std::memcpy(o + 4, &materialIdx, 4);
    std::memcpy(o + 8, m.c, 64); std::memcpy(o + 72, inv.c, 64); std::memcpy(o + 136, &b, 24);
    return NX_OK;
}
int nx_host_camera_record(const nx_camera* cam, uint32_t width, uint32_t height, void* out88)
{
    if (!cam || !out88 || !width || !height) return NX_ERR_INVALID;
    const DCamera c = nxi_camera_to_device(*cam, width, height);
    std::memcpy(out88, &c, 88);
    return NX_OK;
}

int nx_scene_export_lights(nx_scene* s, void* out52, uint32_t* outCount)
{
    if (!s || !outCount) return NX_ERR_INVALID;
    if (s->dirtyLights || s->dirtyInstances || s->dirtyMaterials) { int rc = nx_scene_update(s); if (rc) return rc; }
    *outCount = (uint32_t)s->lights.size();
    if (!out52) return NX_OK;
    uint8_t* o = (uint8_t*)out52;
    for (const DLight& l : s->lights) {   // D_Light: 48-byte union + type byte at offset 48
        std::memset(o, 0, 52);
        if (l.type == NX_LIGHT_MESH) std::memcpy(o, &l.instance, 4);
        else if (l.type == NX_LIGHT_POINT) { float v[7] = {l.px, l.py, l.pz, l.cr, l.cg, l.cb, l.intensity}; std::memcpy(o, v, 28); }
        else if (l.type == NX_LIGHT_DIRECTIONAL) { float v[7] = {l.cr, l.cg, l.cb, l.dx, l.dy, l.dz, l.intensity}; std::memcpy(o, v, 28); }
        o[48] = (uint8_t)l.type;
        o += 52;
    }
    return NX_OK;
}
// The TLAS is built over ENTRIES: the instances that have a BLAS of their own and, last, the merged BLAS.  outInst[e] = instance id of
// entry e, 0xffffffff for the merged BLAS (the TLAS's prim_idx holds entry numbers).
int nx_scene_export_tlas_entries(nx_scene* s, uint32_t* outInst, uint32_t* outCount)
{
    if (!s || !outCount) return NX_ERR_INVALID;
    if (s->dirtyInstances) { int rc = nx_scene_update(s); if (rc) return rc; }
    *outCount = (uint32_t)s->tlasEntryInst.size();
    if (outInst) std::memcpy(outInst, s->tlasEntryInst.data(), 4 * s->tlasEntryInst.size());
    return NX_OK;
}
// The merged BLAS for the parity tests: its handle, and per merged primitive the padded world-space box it was built over (exactly
// those floats: the same kernel again), the instance it belongs to and its primitive id inside that instance's mesh.  Any pointer may be null.
int nx_scene_export_merged(nx_scene* s, nx_bvh8* outBvh, float* hostBounds, uint32_t* hostInst, uint32_t* hostPrim, uint32_t* outCount)
{
    if (!s || !outCount) return NX_ERR_INVALID;
    if (s->dirtyInstances) { int rc = nx_scene_update(s); if (rc) return rc; }
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    *outCount = s->mergedInstances.empty() ? 0u : s->merged.prim_count;
    if (outBvh) *outBvh = s->merged;
    if (!*outCount) return NX_OK;
    const uint32_t n = s->merged.prim_count;
    if (hostInst || hostPrim)
        for (size_t j = 0; j < s->mergedInstances.size(); j++) {
            const uint32_t first = s->mergedFirst[j], count = s->meshes[s->instances[s->mergedInstances[j]].meshIdx].bvh.prim_count;
            for (uint32_t p = 0; p < count; p++) { if (hostInst) hostInst[first + p] = s->mergedInstances[j]; if (hostPrim) hostPrim[first + p] = p; }
        }
    if (hostBounds) {
        std::vector<BakeSrc> src(s->mergedInstances.size());
        for (size_t j = 0; j < src.size(); j++) {
            const HostInstance& h = s->instances[s->mergedInstances[j]];
            src[j].tris = s->meshes[h.meshIdx].dTris; std::memcpy(src[j].m, h.m, 48);
            src[j].first = s->mergedFirst[j]; src[j].count = s->meshes[h.meshIdx].bvh.prim_count; src[j].inst = s->mergedInstances[j];
            src[j].mag = std::max(std::fabs(h.m[3]), std::max(std::fabs(h.m[7]), std::fabs(h.m[11])));
        }
        BakeSrc* dSrc = nullptr; float* dWorld = nullptr; uint32_t* dSrcOf = nullptr;
        int rc = upload_vec(ctx, &dSrc, src); if (rc) return rc;
        NX_CUDA(ctx, cudaMallocAsync((void**)&dWorld, 24 * (size_t)n, ctx->stream));
        NX_CUDA(ctx, cudaMallocAsync((void**)&dSrcOf, 4 * (size_t)n, ctx->stream));
        const int grid = (int)std::min<uint32_t>(div_up(n, 256), (uint32_t)ctx->sm_count * 8u);
        bake_bounds_kernel<<<grid, 256, 0, ctx->stream>>>(dSrc, (uint32_t)src.size(), n, dWorld, dSrcOf);
        NX_CUDA(ctx, cudaMemcpyAsync(hostBounds, dWorld, 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFreeAsync(dSrc, ctx->stream); cudaFreeAsync(dWorld, ctx->stream); cudaFreeAsync(dSrcOf, ctx->stream);
    }
    return NX_OK;
}
int nx_scene_tlas_history(nx_scene* s, uint32_t* outBuilds, uint32_t* outRefits)
{
    if (!s) return NX_ERR_INVALID;
    if (s->dirtyInstances) { int rc = nx_scene_update(s); if (rc) return rc; }
    if (outBuilds) *outBuilds = s->tlasBuilds;
    if (outRefits) *outRefits = s->tlasRefits;
    return NX_OK;
}
int nx_scene_tlas(nx_scene* s, nx_bvh8* out)
{
    if (!s || !out) return NX_ERR_INVALID;
    if (s->dirtyInstances) { int rc = nx_scene_update(s); if (rc) return rc; }
    *out = s->tlas; return NX_OK;
}

// ---- parity hooks ----
int nx_trace_closest_device(nx_scene* s, const nx_ray* dRays, uint32_t n, nx_hit* dHits, float* outMs)
{
    if (!s || !dRays || !dHits) return NX_ERR_INVALID;
    DSceneView v; int rc = nxi_scene_view(s, &v); if (rc) return rc;
    return nxi_trace_closest(s->ctx, v.trace, dRays, n, dHits, outMs);
}
int nx_trace_closest(nx_scene* s, const nx_ray* rays, uint32_t n, nx_hit* hits, float* outMs)
{
    if (!s || !rays || !hits) return NX_ERR_INVALID;
    if (n == 0) return NX_OK;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    DSceneView v; int rc = nxi_scene_view(s, &v); if (rc) return rc;
    nx_ray* dRays = nullptr; nx_hit* dHits = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&dRays, sizeof(nx_ray) * (size_t)n, ctx->stream));
    NX_CUDA(ctx, cudaMallocAsync((void**)&dHits, sizeof(nx_hit) * (size_t)n, ctx->stream));
    NX_CUDA(ctx, cudaMemcpyAsync(dRays, rays, sizeof(nx_ray) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = nxi_trace_closest(ctx, v.trace, dRays, n, dHits, outMs);
    if (!rc) { NX_CUDA(ctx, cudaMemcpyAsync(hits, dHits, sizeof(nx_hit) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream)); NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    cudaFreeAsync(dRays, ctx->stream); cudaFreeAsync(dHits, ctx->stream);
    return rc;
}
int nx_trace_any(nx_scene* s, const nx_ray* rays, uint32_t n, uint8_t* occluded, float* outMs)
{
    if (!s || !rays || !occluded) return NX_ERR_INVALID;
    if (n == 0) return NX_OK;
    nx_ctx* ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    DSceneView v; int rc = nxi_scene_view(s, &v); if (rc) return rc;
    nx_ray* dRays = nullptr; uint8_t* dOcc = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&dRays, sizeof(nx_ray) * (size_t)n, ctx->stream));
    NX_CUDA(ctx, cudaMallocAsync((void**)&dOcc, n, ctx->stream));
    NX_CUDA(ctx, cudaMemcpyAsync(dRays, rays, sizeof(nx_ray) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = nxi_trace_any(ctx, v.trace, dRays, n, dOcc, outMs);
    if (!rc) { NX_CUDA(ctx, cudaMemcpyAsync(occluded, dOcc, n, cudaMemcpyDeviceToHost, ctx->stream)); NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); }
    cudaFreeAsync(dRays, ctx->stream); cudaFreeAsync(dOcc, ctx->stream);
    return rc;
}

} // extern "C"
