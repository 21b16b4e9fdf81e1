// Two-level CWBVH8 traversal (closest hit and any hit) for sm_100a.
//
// Replaces BVH8Trace / BVH8TraceShadow / ChildTrace / TriangleTrace of the reference
// (src/Cuda/BVH/BVH8Traversal.cuh:56-521, src/Cuda/Geometry/Triangle.cuh:29-93).  Same 80-byte node, same octant-ordered
// hit mask, same Moeller-Trumbore formulation, so the closest primitive / instance ids are the reference's.
//
// What is different, and why:
//   * leaves index straight into a leaf-ordered, 48-byte, 16-byte-aligned triangle stream {v0 | primId, e0, e1}: three
//     LDG.128 per triangle instead of one dependent index load plus nine 4-byte loads at 36-byte stride;
//   * TLAS leaves index a 64-byte instance record {inverse 3x4, BLAS node pointer, BLAS triangle pointer} stored in TLAS
//     leaf order: one 64-byte line instead of primIdx -> 160-byte instance -> 72-byte mesh (three dependent loads);
//   * rays are fetched for the whole warp with one atomic (ballot + popc prefix) instead of one atomic per ray;
//   * the traversal stack lives in shared memory (conflict-free, lane-interleaved) with a short local spill tail;
//   * arithmetic uses explicit-rounding intrinsics, IEEE reciprocal included, so results do not depend on compiler
//     contraction and the CPU oracle can reproduce every hit bit for bit.
#pragma once
#include "nx_common.cuh"

#ifndef NX_TRACE_BLOCK
#define NX_TRACE_BLOCK 128
#endif
#ifndef NX_STACK_SHARED
#define NX_STACK_SHARED 10
#endif
#define NX_STACK_TOTAL 32
#define NX_MISS_T 1.0e30f   // miss sentinel of the reference (PathTracer.cu:140)

struct DTravInst {           // 64 B, one cache-line half; TLAS-leaf order
    float4 r0, r1, r2;       // rows of the inverse instance transform (world -> object), row-major 3x4
    const float4* nodes;     // BLAS nodes (5 x float4 each)
    const float4* ltris;     // BLAS leaf-ordered triangles (3 x float4 each)
};

struct TraceScene {
    const float4* tlasNodes;
    const uint32_t* tlasPrimIdx;   // TLAS leaf slot -> instance id (read once per ray, at the end)
    const DTravInst* inst;         // TLAS leaf slot -> traversal record
};

struct HitRec { float t, u, v; uint32_t prim, slot; };

struct TraceStats { unsigned long long nodes, tris, insts, rays; };

// stack: first NX_STACK_SHARED entries in shared memory (entry e of thread t at [e * blockDim + t]), the rest in local memory
struct TravStack {
    uint2* sh; uint2 loc[NX_STACK_TOTAL - NX_STACK_SHARED]; int sp;
    __device__ __forceinline__ void push(uint2 v) { if (sp < NX_STACK_SHARED) sh[sp * NX_TRACE_BLOCK] = v; else loc[sp - NX_STACK_SHARED] = v; sp++; }
    __device__ __forceinline__ uint2 pop() { sp--; return sp < NX_STACK_SHARED ? sh[sp * NX_TRACE_BLOCK] : loc[sp - NX_STACK_SHARED]; }
};

__device__ __forceinline__ uint32_t octant_inv(V3 d) { return 7u - (((d.x < 0.f) ? 4u : 0u) | ((d.y < 0.f) ? 2u : 0u) | ((d.z < 0.f) ? 1u : 0u)); }
__device__ __forceinline__ float rcp_ieee(float x) { return __frcp_rn(x); }

// Slab test of the eight quantised child boxes of one node.  Returns the inner-node group (childBase, hits<<24 | imask)
// and the leaf group (primBase, hit bits 0..23).
__device__ __forceinline__ void intersect_children(const float4* __restrict__ nodes, uint32_t idx, V3 o, V3 d, V3 inv, uint32_t octinv4, float tmax,
                                                   uint2& inner, uint2& leaves)
{
    const float4* nd = nodes + 5 * (size_t)idx;
    const float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3), n4 = __ldg(nd + 4);
    const uint32_t eim = __float_as_uint(n0.w);
    // per-axis cell size 2^(e-127) scaled by 1/d, and the node origin relative to the ray, scaled by 1/d
    const float sx = __fmul_rn(__uint_as_float((eim & 0xffu) << 23), inv.x);
    const float sy = __fmul_rn(__uint_as_float(((eim >> 8) & 0xffu) << 23), inv.y);
    const float sz = __fmul_rn(__uint_as_float(((eim >> 16) & 0xffu) << 23), inv.z);
    const float ox = __fmul_rn(__fsub_rn(n0.x, o.x), inv.x), oy = __fmul_rn(__fsub_rn(n0.y, o.y), inv.y), oz = __fmul_rn(__fsub_rn(n0.z, o.z), inv.z);
    uint32_t hits = 0;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const uint32_t meta4 = __float_as_uint(h ? n1.w : n1.z);
        const uint32_t inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;                 // bit 4 of a byte set <=> low5 >= 24 <=> inner child
        const uint32_t innerFF = (inner4 >> 4) * 0xffu;                               // 0xff in the bytes of inner children
        const uint32_t bitIdx4 = (meta4 ^ (octinv4 & innerFF)) & 0x1f1f1f1fu;         // inner: 24 + (slot ^ octinv); leaf: first triangle bit
        const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;                            // inner: 1; leaf: unary triangle count
        const uint32_t lox = __float_as_uint(h ? n2.y : n2.x), loy = __float_as_uint(h ? n2.w : n2.z), loz = __float_as_uint(h ? n3.y : n3.x);
        const uint32_t hix = __float_as_uint(h ? n3.w : n3.z), hiy = __float_as_uint(h ? n4.y : n4.x), hiz = __float_as_uint(h ? n4.w : n4.z);
        const uint32_t nearx = d.x < 0.f ? hix : lox, farx = d.x < 0.f ? lox : hix;
        const uint32_t neary = d.y < 0.f ? hiy : loy, fary = d.y < 0.f ? loy : hiy;
        const uint32_t nearz = d.z < 0.f ? hiz : loz, farz = d.z < 0.f ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const uint32_t sh = 8u * j;
            const float t0x = __fmaf_rn((float)((nearx >> sh) & 0xffu), sx, ox), t1x = __fmaf_rn((float)((farx >> sh) & 0xffu), sx, ox);
            const float t0y = __fmaf_rn((float)((neary >> sh) & 0xffu), sy, oy), t1y = __fmaf_rn((float)((fary >> sh) & 0xffu), sy, oy);
            const float t0z = __fmaf_rn((float)((nearz >> sh) & 0xffu), sz, oz), t1z = __fmaf_rn((float)((farz >> sh) & 0xffu), sz, oz);
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            if (tn <= tf) hits |= ((bits4 >> sh) & 0xffu) << ((bitIdx4 >> sh) & 0xffu);
        }
    }
    inner = make_uint2(__float_as_uint(n1.x), (hits & 0xff000000u) | (eim >> 24));
    leaves = make_uint2(__float_as_uint(n1.y), hits & 0x00ffffffu);
}

// Moeller-Trumbore on {v0, e0 = v1 - v0, e1 = v2 - v0}; no back-face culling; accepts 0 < t < best (strict), as
// Triangle.cuh:29-62.  Returns true when the hit was accepted.
__device__ __forceinline__ bool intersect_triangle(const float4* __restrict__ tri, V3 o, V3 d, float& best, float& bu, float& bv, uint32_t& prim)
{
    const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
    const V3 e0 = v3(b.x, b.y, b.z), e1 = v3(c.x, c.y, c.z);
    const V3 pv = xcross(d, e1);
    const float det = xdot(e0, pv);
    const float invDet = rcp_ieee(det);
    const V3 s = o - v3(a.x, a.y, a.z);
    const float u = __fmul_rn(invDet, xdot(s, pv));
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qv = xcross(s, e0);
    const float v = __fmul_rn(invDet, xdot(d, qv));
    if (v < 0.0f || __fadd_rn(u, v) > 1.0f) return false;
    const float t = __fmul_rn(invDet, xdot(e1, qv));
    if (t > 0.0f && t < best) { best = t; bu = u; bv = v; prim = __float_as_uint(a.w); return true; }
    return false;
}

__device__ __forceinline__ V3 xform_point(const DTravInst& I, V3 p)
{
    return v3(__fmaf_rn(I.r0.x, p.x, __fmaf_rn(I.r0.y, p.y, __fmaf_rn(I.r0.z, p.z, I.r0.w))),
              __fmaf_rn(I.r1.x, p.x, __fmaf_rn(I.r1.y, p.y, __fmaf_rn(I.r1.z, p.z, I.r1.w))),
              __fmaf_rn(I.r2.x, p.x, __fmaf_rn(I.r2.y, p.y, __fmaf_rn(I.r2.z, p.z, I.r2.w))));
}
__device__ __forceinline__ V3 xform_vector(const DTravInst& I, V3 p)
{
    return v3(__fmaf_rn(I.r0.x, p.x, __fmaf_rn(I.r0.y, p.y, __fmul_rn(I.r0.z, p.z))),
              __fmaf_rn(I.r1.x, p.x, __fmaf_rn(I.r1.y, p.y, __fmul_rn(I.r1.z, p.z))),
              __fmaf_rn(I.r2.x, p.x, __fmaf_rn(I.r2.y, p.y, __fmul_rn(I.r2.z, p.z))));
}

// Per-lane traversal state machine.  One call to step() intersects one node (or pops) and then works off the leaf group.
// Written as a resumable state so the persistent kernel can refill finished lanes between steps.
template <bool ANY_HIT, bool STATS>
struct Traverser {
    V3 o, d, inv;          // current-space ray
    V3 wo, wd;             // world-space ray (restored when leaving an instance)
    float tmax;            // any-hit: fixed limit; closest-hit: shrinks with every accepted hit
    HitRec hit;
    uint2 ngroup, tgroup;
    const float4* nodes; const float4* ltris;
    uint32_t octinv4, curSlot;
    int instDepth;         // stack depth at which the current instance was entered, -1 in the TLAS
    bool occluded;
    TravStack st;
    uint32_t cNodes, cTris, cInsts;

    __device__ __forceinline__ void begin(const TraceScene& sc, V3 ro, V3 rd, float limit)
    {
        o = wo = ro; d = wd = rd;
        inv = v3(rcp_ieee(rd.x), rcp_ieee(rd.y), rcp_ieee(rd.z));
        tmax = limit; hit.t = NX_MISS_T; hit.u = hit.v = 0.f; hit.prim = NX_INVALID; hit.slot = NX_INVALID;
        ngroup = make_uint2(0u, 0x80000000u); tgroup = make_uint2(0u, 0u);
        nodes = sc.tlasNodes; ltris = nullptr;
        octinv4 = octant_inv(rd) * 0x01010101u; curSlot = NX_INVALID; instDepth = -1; occluded = false; st.sp = 0;
        if (STATS) cNodes = cTris = cInsts = 0;
    }

    // returns true when the ray is finished
    __device__ __forceinline__ bool step(const TraceScene& sc)
    {
        if (ngroup.y & 0xff000000u)
        {
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) st.push(ngroup);
            const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
            const uint32_t child = ngroup.x + __popc(ngroup.y & ((1u << slot) - 1u) & 0xffu);
            intersect_children(nodes, child, o, d, inv, octinv4, ANY_HIT ? tmax : fminf(tmax, hit.t), ngroup, tgroup);
            if (STATS) cNodes++;
        }
        else { tgroup = ngroup; ngroup = make_uint2(0u, 0u); }

        while (tgroup.y)
        {
            const uint32_t bit = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << bit);
            if (instDepth < 0)
            {
                // TLAS leaf: enter the instance.  What is left of this node goes on the stack first.
                if (tgroup.y) st.push(tgroup);
                if (ngroup.y & 0xff000000u) st.push(ngroup);
                instDepth = st.sp;
                curSlot = tgroup.x + bit;
                const DTravInst* I = sc.inst + curSlot;
                DTravInst T; T.r0 = __ldg(&I->r0); T.r1 = __ldg(&I->r1); T.r2 = __ldg(&I->r2);
                const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&I->nodes));
                nodes = reinterpret_cast<const float4*>(((uint64_t)ptrs.y << 32) | ptrs.x);
                ltris = reinterpret_cast<const float4*>(((uint64_t)ptrs.w << 32) | ptrs.z);
                o = xform_point(T, wo); d = xform_vector(T, wd);     // direction is not renormalised: t stays in world units
                inv = v3(rcp_ieee(d.x), rcp_ieee(d.y), rcp_ieee(d.z));
                octinv4 = octant_inv(d) * 0x01010101u;
                ngroup = make_uint2(0u, 0x80000000u); tgroup = make_uint2(0u, 0u);
                if (STATS) cInsts++;
                return false;
            }
            if (STATS) cTris++;
            float best = ANY_HIT ? tmax : fminf(tmax, hit.t);
            if (intersect_triangle(ltris + 3 * (size_t)(tgroup.x + bit), o, d, best, hit.u, hit.v, hit.prim))
            {
                if (ANY_HIT) { occluded = true; return true; }
                hit.t = best; hit.slot = curSlot;
            }
        }

        if ((ngroup.y & 0xff000000u) == 0u)
        {
            if (st.sp == 0) return true;
            if (st.sp == instDepth)
            {
                o = wo; d = wd; inv = v3(rcp_ieee(wd.x), rcp_ieee(wd.y), rcp_ieee(wd.z));
                octinv4 = octant_inv(wd) * 0x01010101u;
                nodes = sc.tlasNodes; instDepth = -1;
            }
            ngroup = st.pop();
        }
        return false;
    }
};
