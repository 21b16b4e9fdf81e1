// TEST INFRASTRUCTURE — CPU restatement of the reference's two-level CWBVH8 traversal, plus a brute-force closest hit.
// See oracle_common.h for who may use this and how it is pinned.
//
// Follows BVH8Trace / ChildTrace (Nexus/src/Cuda/BVH/BVH8Traversal.cuh:56-324), TriangleTrace
// (Nexus/src/Cuda/Geometry/Triangle.cuh:29-62) and Mat4::TransformPoint/Vector (Nexus/src/Math/Mat4.h:217-230).
// Arithmetic note: the reference is built with --use_fast_math (approximate reciprocals, compiler-chosen FMA
// contraction), which no CPU can reproduce bit for bit.  This restatement therefore fixes one explicit IEEE operation
// order (fmaf where a multiply-add is fused, correctly rounded reciprocals); the product kernels use the same order, so
// product-vs-oracle comparisons are exact, while both are compared to the real reference kernels (oracle/_ref, on a
// GPU) with the tolerance north_star states: ids equal, |dt| <= 1e-5 relative.
#include "oracle_common.h"
#include <thread>
#include <atomic>

using namespace orc;

namespace {

struct Mesh {
    std::vector<float> tris;          // n * 9
    std::vector<Node8> nodes;
    std::vector<uint32_t> primIdx;
    // The product's merged BLAS (nexus_b200/csrc/scene.cu; no counterpart in the reference): nodes in world space over many instances'
    // triangles, each triangle kept in the OBJECT space of its instance and tested with the ray transformed the way an instance entry
    // transforms it, so that the arithmetic of every triangle test is the two-level scene's.  instOf / primOf: per primitive of this
    // mesh, the scene instance it belongs to and its index inside that instance's own mesh (what a hit reports).
    std::vector<uint32_t> instOf, primOf;
};
struct Instance { uint32_t mesh; float inv[12]; uint32_t id; };   // id: the instance id a hit reports (the entry's own index unless set)
struct SceneO {
    std::vector<Mesh> meshes;
    std::vector<Instance> inst;
    std::vector<Node8> tlas;
    std::vector<uint32_t> tlasPrim;
    std::vector<float> instInv;       // merged BLAS only: scene instance id -> world -> object 3x4 rows
};

struct Ray { f3 o; float tmax; f3 d; uint32_t pad; };
struct Hit { float t, u, v; uint32_t prim, inst; };
static_assert(sizeof(Ray) == 32 && sizeof(Hit) == 20, "ray / hit layout");

inline float xdot(f3 a, f3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline f3 xcross(f3 a, f3 b) { return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))}; }
inline f3 xpoint(const float* m, f3 p)
{
    return {fmaf(m[0], p.x, fmaf(m[1], p.y, fmaf(m[2], p.z, m[3]))), fmaf(m[4], p.x, fmaf(m[5], p.y, fmaf(m[6], p.z, m[7]))),
            fmaf(m[8], p.x, fmaf(m[9], p.y, fmaf(m[10], p.z, m[11])))};
}
inline f3 xvector(const float* m, f3 p)
{
    return {fmaf(m[0], p.x, fmaf(m[1], p.y, m[2] * p.z)), fmaf(m[4], p.x, fmaf(m[5], p.y, m[6] * p.z)), fmaf(m[8], p.x, fmaf(m[9], p.y, m[10] * p.z))};
}
inline uint32_t octantInv(f3 d) { return 7u - (((d.x < 0.f) ? 4u : 0u) | ((d.y < 0.f) ? 2u : 0u) | ((d.z < 0.f) ? 1u : 0u)); }

// Triangle.cuh:29-62 on (v0, e0 = v1 - v0, e1 = v2 - v0).  Returns the distance of a valid intersection (t > 0) or -1.
inline float triangleT(const float* t, f3 o, f3 d, float& u, float& v)
{
    f3 v0 = mk(t[0], t[1], t[2]);
    f3 e0 = mk(t[3], t[4], t[5]) - v0, e1 = mk(t[6], t[7], t[8]) - v0;
    f3 pv = xcross(d, e1);
    float det = xdot(e0, pv);
    float inv = 1.0f / det;
    f3 s = o - v0;
    u = inv * xdot(s, pv);
    if (!(u >= 0.0f && u <= 1.0f)) return -1.0f;
    f3 qv = xcross(s, e0);
    v = inv * xdot(d, qv);
    if (!(v >= 0.0f && u + v <= 1.0f)) return -1.0f;
    float tt = inv * xdot(e1, qv);
    return tt > 0.0f ? tt : -1.0f;
}
inline bool triangle(const float* t, f3 o, f3 d, float& best, float& bu, float& bv)
{
    f3 v0 = mk(t[0], t[1], t[2]);
    f3 e0 = mk(t[3], t[4], t[5]) - v0, e1 = mk(t[6], t[7], t[8]) - v0;
    f3 pv = xcross(d, e1);
    float det = xdot(e0, pv);
    float inv = 1.0f / det;
    f3 s = o - v0;
    float u = inv * xdot(s, pv);
    if (u < 0.0f || u > 1.0f) return false;
    f3 qv = xcross(s, e0);
    float v = inv * xdot(d, qv);
    if (v < 0.0f || u + v > 1.0f) return false;
    float tt = inv * xdot(e1, qv);
    if (tt > 0.0f && tt < best) { best = tt; bu = u; bv = v; return true; }
    return false;
}

// reciprocal of a direction component for the slab test, clamped like the product's rcp_dir (traverse.cuh): an infinite reciprocal
// turns the fused slab formula into inf - inf and the axis is ignored, so an axis-parallel ray would visit every node it passes in
// the other axes.  Hits do not depend on it (the box test is only conservative culling).
inline float rcpDir(float x) { float r = 1.0f / x; return r > 1.0e18f ? 1.0e18f : (r < -1.0e18f ? -1.0e18f : r); }

// ChildTrace, BVH8Traversal.cuh:56-147: 32-bit hit mask, inner children in bits 24..31 ordered by octant
inline uint32_t childHits(const Node8& N, f3 o, f3 d, f3 inv, uint32_t oinv, float tmax)
{
    float sx = u2f((uint32_t)N.e[0] << 23) * inv.x, sy = u2f((uint32_t)N.e[1] << 23) * inv.y, sz = u2f((uint32_t)N.e[2] << 23) * inv.z;
    float ox = (N.p.x - o.x) * inv.x, oy = (N.p.y - o.y) * inv.y, oz = (N.p.z - o.z) * inv.z;
    uint32_t hits = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t meta = N.meta[i];
        if (!meta) continue;
        bool inner = (meta & 0x1f) >= 24;
        uint32_t bitIdx = inner ? ((meta ^ oinv) & 0x1f) : (meta & 0x1f);
        uint32_t bits = meta >> 5;
        float lox = N.qlox[i], loy = N.qloy[i], loz = N.qloz[i], hix = N.qhix[i], hiy = N.qhiy[i], hiz = N.qhiz[i];
        float nx = d.x < 0.f ? hix : lox, fx = d.x < 0.f ? lox : hix;
        float ny = d.y < 0.f ? hiy : loy, fy = d.y < 0.f ? loy : hiy;
        float nz = d.z < 0.f ? hiz : loz, fz = d.z < 0.f ? loz : hiz;
        float t0x = fmaf(nx, sx, ox), t1x = fmaf(fx, sx, ox), t0y = fmaf(ny, sy, oy), t1y = fmaf(fy, sy, oy), t0z = fmaf(nz, sz, oz), t1z = fmaf(fz, sz, oz);
        float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
        float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
        if (tn <= tf) hits |= bits << bitIdx;
    }
    return hits;
}

struct Stats { uint64_t nodes = 0, tris = 0, insts = 0; };

// BVH8Trace without the SIMT bookkeeping: depth-first, children in octant order, leaves of a node before its stacked siblings.
template <bool ANY>
bool traverseBlas(const SceneO& S, const Mesh& M, f3 o, f3 d, float& best, Hit& hit, uint32_t instId, Stats& st)
{
    const bool merged = !M.instOf.empty();
    f3 inv = mk(rcpDir(d.x), rcpDir(d.y), rcpDir(d.z));
    uint32_t oinv = octantInv(d);
    struct Entry { uint32_t base, hits, imask; bool tri; };
    std::vector<Entry> stack;
    Entry ngroup{0, 0x80000000u, 0, false};
    while (true) {
        Entry tgroup{0, 0, 0, true};
        if (ngroup.hits & 0xff000000u) {
            uint32_t bit = 31 - __builtin_clz(ngroup.hits);
            ngroup.hits &= ~(1u << bit);
            if (ngroup.hits & 0xff000000u) stack.push_back(ngroup);
            uint32_t slot = (bit - 24) ^ oinv;
            uint32_t child = ngroup.base + __builtin_popcount(ngroup.imask & ((1u << slot) - 1));
            const Node8& N = M.nodes[child];
            st.nodes++;
            uint32_t h = childHits(N, o, d, inv, oinv, best);
            ngroup = Entry{N.childBaseIdx, h & 0xff000000u, N.imask, false};
            tgroup = Entry{N.primBaseIdx, h & 0x00ffffffu, 0, true};
        } else { tgroup = ngroup; tgroup.tri = true; ngroup = Entry{0, 0, 0, false}; }
        while (tgroup.hits) {
            uint32_t bit = 31 - __builtin_clz(tgroup.hits);
            tgroup.hits &= ~(1u << bit);
            uint32_t prim = M.primIdx[tgroup.base + bit];
            st.tris++;
            float u, v;
            f3 to = o, td = d;
            if (merged) { const float* inv = &S.instInv[12 * (size_t)M.instOf[prim]]; to = xpoint(inv, o); td = xvector(inv, d); }
            const float tt = triangleT(&M.tris[9 * (size_t)prim], to, td, u, v);
            if (merged) { instId = M.instOf[prim]; prim = M.primOf[prim]; }
            if (tt > 0.0f) {
                // closer hit, or an exact tie resolved by (instance id, primitive id) so that the visiting order cannot matter
                // (the reference keeps whichever it met first, which depends on its warp schedule: SURVEY.md §7)
                bool take = tt < best;
                if (!ANY && !take && tt == hit.t && hit.prim != INVALID) take = instId < hit.inst || (instId == hit.inst && prim < hit.prim);
                if (take) {
                    best = tt; hit.t = tt; hit.u = u; hit.v = v; hit.prim = prim; hit.inst = instId;
                    if (ANY) return true;
                }
            }
        }
        if ((ngroup.hits & 0xff000000u) == 0) {
            if (stack.empty()) return false;
            ngroup = stack.back(); stack.pop_back();
            if (ngroup.tri) { /* never pushed in this restatement */ }
        }
    }
}

template <bool ANY>
bool traverse(const SceneO& S, const Ray& r, Hit& hit, Stats& st)
{
    hit.t = 1.0e30f; hit.u = hit.v = 0.f; hit.prim = INVALID; hit.inst = INVALID;
    float best = ANY ? r.tmax : fminf(r.tmax, 1.0e30f);
    f3 o = r.o, d = r.d;
    f3 inv = mk(rcpDir(d.x), rcpDir(d.y), rcpDir(d.z));
    uint32_t oinv = octantInv(d);
    struct Entry { uint32_t base, hits, imask; };
    std::vector<Entry> stack;
    Entry ngroup{0, 0x80000000u, 0};
    while (true) {
        Entry tgroup{0, 0, 0};
        if (ngroup.hits & 0xff000000u) {
            uint32_t bit = 31 - __builtin_clz(ngroup.hits);
            ngroup.hits &= ~(1u << bit);
            if (ngroup.hits & 0xff000000u) stack.push_back(ngroup);
            uint32_t slot = (bit - 24) ^ oinv;
            uint32_t child = ngroup.base + __builtin_popcount(ngroup.imask & ((1u << slot) - 1));
            const Node8& N = S.tlas[child];
            st.nodes++;
            uint32_t h = childHits(N, o, d, inv, oinv, best);
            ngroup = Entry{N.childBaseIdx, h & 0xff000000u, N.imask};
            tgroup = Entry{N.primBaseIdx, h & 0x00ffffffu, 0};
        } else { tgroup = ngroup; ngroup = Entry{0, 0, 0}; }
        if (tgroup.hits) {   // TLAS leaf = instance (BVH8Traversal.cuh:235-268): one instance, the rest of the node waits on the stack
            uint32_t bit = 31 - __builtin_clz(tgroup.hits);
            tgroup.hits &= ~(1u << bit);
            if (tgroup.hits) stack.push_back(tgroup);
            if (ngroup.hits & 0xff000000u) stack.push_back(ngroup);
            ngroup = Entry{0, 0, 0};
            const Instance& I = S.inst[S.tlasPrim[tgroup.base + bit]];
            st.insts++;
            f3 lo = xpoint(I.inv, o), ld = xvector(I.inv, d);   // direction not renormalised: t stays in world units
            if (traverseBlas<ANY>(S, S.meshes[I.mesh], lo, ld, best, hit, I.id, st) && ANY) return true;
        }
        if ((ngroup.hits & 0xff000000u) == 0) {
            if (stack.empty()) return false;
            ngroup = stack.back(); stack.pop_back();
        }
    }
}

template <typename F> void parallelFor(uint32_t n, int threads, F&& f)
{
    if (threads <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> pool;
    std::atomic<uint32_t> next{0};
    const uint32_t chunk = 4096;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() { while (true) { uint32_t b = next.fetch_add(chunk); if (b >= n) break; f(b, std::min(n, b + chunk), t); } });
    for (auto& th : pool) th.join();
}

} // namespace

extern "C" {

void* orc_scene_create() { return new SceneO(); }
void orc_scene_destroy(void* s) { delete (SceneO*)s; }

int orc_scene_add_mesh(void* s, const float* tris, uint32_t n, const void* nodes8, uint32_t nodeCount, const uint32_t* primIdx)
{
    SceneO* S = (SceneO*)s;
    Mesh m;
    m.tris.assign(tris, tris + 9 * (size_t)n);
    m.nodes.resize(nodeCount); std::memcpy(m.nodes.data(), nodes8, 80 * (size_t)nodeCount);
    m.primIdx.assign(primIdx, primIdx + n);
    S->meshes.push_back(std::move(m));
    return (int)S->meshes.size() - 1;
}

// inv: n * 12 floats, rows of the world->object 3x4
int orc_scene_set_instances(void* s, const uint32_t* meshIdx, const float* inv, uint32_t n, const void* tlasNodes, uint32_t tlasNodeCount, const uint32_t* tlasPrimIdx)
{
    SceneO* S = (SceneO*)s;
    S->inst.resize(n);
    for (uint32_t i = 0; i < n; i++) { S->inst[i].mesh = meshIdx[i]; std::memcpy(S->inst[i].inv, inv + 12 * (size_t)i, 48); S->inst[i].id = i; }
    S->tlas.resize(tlasNodeCount); std::memcpy(S->tlas.data(), tlasNodes, 80 * (size_t)tlasNodeCount);
    S->tlasPrim.assign(tlasPrimIdx, tlasPrimIdx + n);
    return 0;
}

// The instance id each TLAS entry reports in a hit (default: its own index).
int orc_scene_set_instance_ids(void* s, const uint32_t* ids, uint32_t n)
{
    SceneO* S = (SceneO*)s;
    if (n != S->inst.size()) return -1;
    for (uint32_t i = 0; i < n; i++) S->inst[i].id = ids[i];
    return 0;
}

// Marks mesh `mesh` as the product's merged BLAS: instOf / primOf per primitive (n = its triangle count; its triangles are the
// object-space triangles in merged order), invTab = nInst * 12 floats, world -> object rows by scene instance id.
int orc_scene_set_merged(void* s, int mesh, const uint32_t* instOf, const uint32_t* primOf, uint32_t n, const float* invTab, uint32_t nInst)
{
    SceneO* S = (SceneO*)s;
    if (mesh < 0 || (size_t)mesh >= S->meshes.size() || S->meshes[mesh].tris.size() != 9 * (size_t)n) return -1;
    for (uint32_t i = 0; i < n; i++) if (instOf[i] >= nInst) return -2;
    S->meshes[mesh].instOf.assign(instOf, instOf + n); S->meshes[mesh].primOf.assign(primOf, primOf + n);
    S->instInv.assign(invTab, invTab + 12 * (size_t)nInst);
    return 0;
}

// rays: n * 32 B {origin, tmax, direction, pad}; hits: n * 20 B {t, u, v, prim, instance}; stats3: nodes, triangles, instances visited
void orc_trace_closest(void* s, const void* rays, uint32_t n, void* hits, int threads, uint64_t* stats3)
{
    const SceneO* S = (const SceneO*)s; const Ray* R = (const Ray*)rays; Hit* H = (Hit*)hits;
    std::vector<Stats> st(std::max(threads, 1));
    parallelFor(n, threads, [&](uint32_t b, uint32_t e, int t) { for (uint32_t i = b; i < e; i++) traverse<false>(*S, R[i], H[i], st[t]); });
    if (stats3) { stats3[0] = stats3[1] = stats3[2] = 0; for (auto& x : st) { stats3[0] += x.nodes; stats3[1] += x.tris; stats3[2] += x.insts; } }
}

void orc_trace_any(void* s, const void* rays, uint32_t n, uint8_t* occluded, int threads)
{
    const SceneO* S = (const SceneO*)s; const Ray* R = (const Ray*)rays;
    std::vector<Stats> st(std::max(threads, 1));
    parallelFor(n, threads, [&](uint32_t b, uint32_t e, int t) { for (uint32_t i = b; i < e; i++) { Hit h; occluded[i] = traverse<true>(*S, R[i], h, st[t]) ? 1 : 0; } });
}

// Every instance x triangle, same triangle arithmetic, instances in index order, triangles in index order.
void orc_trace_brute(void* s, const void* rays, uint32_t n, void* hits, int threads)
{
    const SceneO* S = (const SceneO*)s; const Ray* R = (const Ray*)rays; Hit* H = (Hit*)hits;
    parallelFor(n, threads, [&](uint32_t b, uint32_t e, int) {
        for (uint32_t i = b; i < e; i++) {
            Hit h; h.t = 1.0e30f; h.u = h.v = 0.f; h.prim = INVALID; h.inst = INVALID;
            float best = fminf(R[i].tmax, 1.0e30f);
            for (uint32_t k = 0; k < S->inst.size(); k++) {
                const Instance& I = S->inst[k]; const Mesh& M = S->meshes[I.mesh];
                f3 lo = xpoint(I.inv, R[i].o), ld = xvector(I.inv, R[i].d);
                uint32_t nt = (uint32_t)(M.tris.size() / 9);
                for (uint32_t p = 0; p < nt; p++)
                    if (triangle(&M.tris[9 * (size_t)p], lo, ld, best, h.u, h.v)) { h.t = best; h.prim = p; h.inst = k; }
            }
            H[i] = h;
        }
    });
}

// t of one specific (instance, triangle) pair for a ray, or 1e30: used to classify id mismatches as exact ties.
float orc_triangle_t(void* s, const void* ray, uint32_t inst, uint32_t prim)
{
    const SceneO* S = (const SceneO*)s; const Ray& r = *(const Ray*)ray;
    if (inst >= S->inst.size()) return 1.0e30f;
    const Instance& I = S->inst[inst]; const Mesh& M = S->meshes[I.mesh];
    if ((size_t)prim * 9 >= M.tris.size()) return 1.0e30f;
    float best = 1.0e30f, u, v;
    triangle(&M.tris[9 * (size_t)prim], xpoint(I.inv, r.o), xvector(I.inv, r.d), best, u, v);
    return best;
}

} // extern "C"
