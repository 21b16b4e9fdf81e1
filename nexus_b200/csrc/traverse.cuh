// Two-level CWBVH8 traversal (closest hit and any hit) for sm_100a.
//
// Replaces BVH8Trace / BVH8TraceShadow / ChildTrace / TriangleTrace of the reference
// (src/Cuda/BVH/BVH8Traversal.cuh:56-521, src/Cuda/Geometry/Triangle.cuh:29-93).  Same 80-byte node, same octant-ordered
// hit mask, same Moeller-Trumbore formulation, so the closest primitive / instance ids are the reference's.
//
// What is different, and why (ncu evidence in profiles/):
//   * leaves index straight into a leaf-ordered, 48-byte, 16-byte-aligned triangle stream {v0 | primId, e0, e1}: three
//     LDG.128 per triangle instead of one dependent index load plus nine 4-byte loads at 36-byte stride;
//   * TLAS leaves index a 64-byte instance record {inverse 3x4, BLAS node pointer, BLAS triangle pointer} stored in TLAS
//     leaf order: one 64-byte line instead of primIdx -> 160-byte instance -> 72-byte mesh (three dependent loads);
//   * rays are fetched for the whole warp with one atomic (ballot + popc prefix) instead of one atomic per ray;
//   * the warp runs in PHASES with convergent votes in between (pop/refill -> enter instance -> node -> triangles), so all
//     lanes that have a node to test execute the 8-wide slab test together, and triangle / instance work is batched until
//     enough lanes have some (or a lane has nothing else to do).  The first version ran one resumable state machine per
//     lane and measured 12-13 active lanes per instruction (ncu smsp__thread_inst_executed_per_inst_executed);
//   * all per-ray state lives in registers; the traversal stack is in shared memory (lane-interleaved, conflict-free) with
//     a separate local spill array that only deep paths touch; the world-space ray is parked in shared memory while the
//     lane is inside an instance;
//   * equal-distance hits are resolved by (instance id, primitive id), not by visiting order, so the answer does not depend
//     on which rays share a warp or on how triangle work is batched (the reference's result on exact ties is schedule
//     dependent);
//   * arithmetic uses explicit-rounding intrinsics, IEEE reciprocal included, so results do not depend on compiler
//     contraction and the CPU oracle can reproduce every hit bit for bit.
#pragma once
#include "nx_common.cuh"

#ifndef NX_TRACE_BLOCK
#define NX_TRACE_BLOCK 128
#endif
#ifndef NX_STACK_SHARED
#define NX_STACK_SHARED 8
#endif
#ifndef NX_TRACE_MIN_BLOCKS
#define NX_TRACE_MIN_BLOCKS 8
#endif
#define NX_STACK_TOTAL 40
#define NX_MISS_T 1.0e30f   // miss sentinel of the reference (PathTracer.cu:140)

struct DTravInst {           // 80 B; TLAS-leaf order
    float4 sphere;           // world-space bounding sphere of the instance's geometry (centre, radius): culls the loose
                             // world AABB of a rotated object before the ray is transformed (a conservative test, results unchanged)
    float4 r0, r1, r2;       // rows of the inverse instance transform (world -> object), row-major 3x4
    const float4* nodes;     // BLAS nodes (5 x float4 each)
    const float4* ltris;     // BLAS leaf-ordered triangles (3 x float4 each)
};

struct TraceScene {
    const float4* tlasNodes;
    const uint32_t* tlasPrimIdx;   // TLAS leaf slot -> instance id (read once per ray, at the end); unused for the merged BLAS's slot
    const DTravInst* inst;         // TLAS leaf slot -> traversal record
    // Merged BLAS (scene.cu): the instances whose mesh nobody else uses share ONE BLAS whose NODES are in world space (built over the
    // world-space boxes of their triangles), entered with the identity transform.  Its leaf records stay in OBJECT space and carry the
    // instance id ({v0 | prim}, {e0 | instance}, {e1 | -}): a triangle is tested with the ray taken into its instance's object space by
    // the same arithmetic an instance entry uses, so t, u, v and the accept / reject decision are bit for bit those of the two-level
    // scene (a world-space triangle test rounds differently: at coordinates of ~1000 units it moved hits by 1e-4..1e-3 of a unit and
    // flipped near-origin hits of secondary rays, 2e-4 of the rays of the 10 M-triangle scene).
    uint32_t mergedSlot;           // TLAS leaf slot of that BLAS, NX_INVALID when the scene has none
    uint32_t direct;               // 1: it is the ONLY TLAS entry, rays start inside it and never see the TLAS
    const float4* mNodes; const float4* mLtris;
    const float4* instInv;         // instance id -> rows of its world -> object 3x4 (3 x float4), for the triangles of the merged BLAS
    uint32_t* overflow;            // device counter of traversal-stack pushes refused (a tree deeper than NX_STACK_TOTAL entries); the
                                   // host turns a non-zero value into an error (the reference's 32-entry stack has no check, BVH8Traversal.cuh:164)
};

struct TraceTuning {               // batching thresholds (lanes): run a phase when at least this many lanes want it
    uint32_t triLanes;             // triangle phase
    uint32_t instLanes;            // new-ray / instance-entry phase
    uint32_t sphereCull;           // 0 disables the per-instance bounding-sphere test (measurement only)
    uint32_t k47;                  // always 0x47000000 (see trace_loop: a constant the compiler must not see)
    uint32_t stackLimit;           // NX_STACK_TOTAL; a test lowers it (not below NX_STACK_SHARED) to exercise the overflow report
};

struct TraceStats {
    unsigned long long nodes, tris, insts, rays;
    // warp scheduling: loop iterations, lanes that tested a node, triangle rounds / lanes, set-up rounds / lanes (per warp, summed)
    unsigned long long iters, lanesN, roundsT, lanesT, roundsX, lanesX, sphereCulled;
    unsigned long long roundsN, roundsF, lanesF;   // ray-pool loop only: node rounds, fetch rounds and the rays they fetched
};

// Shared memory per block: stack entries [NX_STACK_SHARED][block] of uint2, then the parked world-space ray, 48 B per
// thread: {origin, octant word} {direction, -} {reciprocal direction, -}.  Three LDS.128 / STS.128 per instance exit /
// entry; a 48-byte thread stride keeps the 16-byte accesses of a quarter warp on distinct banks.
#define NX_TRACE_SMEM_BYTES ((NX_STACK_SHARED * 8 + 48) * NX_TRACE_BLOCK)
#define NX_TRACE_SMEM_BYTES_DIRECT (NX_STACK_SHARED * 8 * NX_TRACE_BLOCK)   // NX_SCENE_DIRECT: no instance is ever entered, nothing is parked

// (7 - octant) replicated into four bytes, octant = sign bits of the direction (x: 4, y: 2, z: 1).  Any value works as long
// as the same one decodes the hit mask it encoded (it only fixes the visiting order), so the sign BITS are used: shifts
// and one multiply instead of three float compares and selects.
__device__ __forceinline__ uint32_t octant_inv4(V3 d)
{
    const uint32_t oct = ((__float_as_uint(d.x) >> 31) << 2) | ((__float_as_uint(d.y) >> 31) << 1) | (__float_as_uint(d.z) >> 31);
    return (7u - oct) * 0x01010101u;
}
__device__ __forceinline__ float rcp_ieee(float x) { return __frcp_rn(x); }

// A hit remembers WHERE it was found as one word: a TLAS leaf slot, or - inside the merged BLAS - the instance id from the triangle
// record with the top bit set.  NX_INVALID (no hit) has that bit set too and is tested first.
__device__ __forceinline__ uint32_t hit_instance(const TraceScene& sc, uint32_t where)
{
    if (where == NX_INVALID) return NX_INVALID;
    return (where & 0x80000000u) ? (where & 0x7fffffffu) : __ldg(sc.tlasPrimIdx + where);
}

// Which of the six quantised planes per child are converted byte -> float on the ALU pipe (PRMT into the mantissa of
// 2^15, the bias folded into the FMA addend) instead of the XU pipe (I2F.U8): bit a = near plane of axis a, bit 3 + a =
// far plane.  ncu (profiles/r01_trace_closest.md): 48 I2F.U8 per node kept the XU pipe at 54-61 % with the I2Fs holding
// 15 % of all stall samples; splitting the conversions over both pipes removes that queue.
#ifndef NX_MAGIC_PLANES
#define NX_MAGIC_PLANES 0x2d
#endif

__device__ __forceinline__ float sqrt_fast(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // one MUFU; 2^-22 relative
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// Reciprocal of a direction component for the slab test, clamped to +-1e18.  With an infinite reciprocal (a ray exactly parallel to an
// axis) the fused form q * (cell / d) + (p - o) / d of the slab test is inf - inf = NaN on that axis, the NaN drops out of min / max and
// the axis is IGNORED: conservative, but a vertical ray then visits every node of the scene whose height range it crosses, whatever
// their x and z (found in round 2: one such ray per frame took 5 s in a flat 10 M-triangle BVH).  With a huge finite reciprocal the
// planes of that axis sit at -huge / +huge when the origin is inside the slab and at the same sign when it is outside: a miss.
__device__ __forceinline__ float rcp_dir(float x) { return fminf(fmaxf(rcp_fast(x), -1.0e18f), 1.0e18f); }
__device__ __forceinline__ uint32_t shl_wrap(uint32_t v, uint32_t n) { uint32_t r; asm("shf.l.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0u), "r"(v), "r"(n)); return r; }

// Byte j of `word` as a float.  MAGIC: the byte is spliced into 0x4700qq00 = 32768 + q (one PRMT on the ALU pipe); the
// caller folds -32768 * s into the addend.  Otherwise I2F.U8 on the XU pipe.
template <bool MAGIC>
__device__ __forceinline__ float plane_q(uint32_t word, int j, uint32_t k47)
{
    if (MAGIC) return __uint_as_float(__byte_perm(word, k47, 0x7504u | (uint32_t)(j << 4)));
    return (float)((word >> (8 * j)) & 0xffu);
}
// {t0, t1} = {q0, q1} * s + {c0, c1}: one FFMA2 (sm_100 packed fp32, scalar-broadcast multiplier) = one issue slot for the
// near and the far plane of an axis.  Each half is an IEEE fma, so the CPU oracle's fmaf() reproduces it.
#ifndef NX_FFMA2
#define NX_FFMA2 1
#endif
__device__ __forceinline__ void fma2_bcast(float& t0, float& t1, float q0, float q1, float s, float c0, float c1)
{
#if !NX_FFMA2
    t0 = __fmaf_rn(q0, s, c0); t1 = __fmaf_rn(q1, s, c1); return;
#endif
    asm("{\n\t.reg .b64 ra, rs, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rd, ra, rs, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(t0), "=f"(t1) : "f"(q0), "f"(q1), "f"(s), "f"(c0), "f"(c1));
}

// Slab test of the eight quantised child boxes of one node.  Returns the inner-node group (childBase, hits<<24 | imask)
// and the leaf group (primBase, hit bits 0..23).
//
// The test is CONSERVATIVE, not exact: `inv` is the hardware reciprocal approximation and every slab is widened by
// eps = 2^-6 cell + 2^-20 |(p - o) * inv| per axis, which covers the approximation (2^-22 relative), the 2^-9 cell lost to
// the folded bias and the roundings of the original formulation.  A widened box can only add node visits; which triangle
// is the closest hit is decided by the triangle test alone (exact arithmetic, deterministic tie-break), so hits are
// unchanged.  NaNs (0 * inf on axis-parallel rays) drop out of min/max, i.e. the axis is ignored: also conservative.
__device__ __forceinline__ void intersect_children(const float4* __restrict__ nodes, uint32_t idx, V3 o, V3 inv, uint32_t octinv4, float tmax,
                                                   uint32_t k47, uint2& inner, uint2& leaves)
{
    const float4* nd = nodes + 5 * (size_t)idx;
    const float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3), n4 = __ldg(nd + 4);
    const uint32_t eim = __float_as_uint(n0.w);
    // per-axis cell size 2^(e-127) scaled by 1/d, and the node origin relative to the ray, scaled by 1/d
    const float sx = __fmul_rn(__uint_as_float((eim & 0xffu) << 23), inv.x);
    const float sy = __fmul_rn(__uint_as_float(((eim >> 8) & 0xffu) << 23), inv.y);
    const float sz = __fmul_rn(__uint_as_float(((eim >> 16) & 0xffu) << 23), inv.z);
    const float ox = __fmul_rn(__fsub_rn(n0.x, o.x), inv.x), oy = __fmul_rn(__fsub_rn(n0.y, o.y), inv.y), oz = __fmul_rn(__fsub_rn(n0.z, o.z), inv.z);
    const float ex = __fmaf_rn(fabsf(ox), 0x1p-20f, __fmul_rn(fabsf(sx), 0x1p-6f));
    const float ey = __fmaf_rn(fabsf(oy), 0x1p-20f, __fmul_rn(fabsf(sy), 0x1p-6f));
    const float ez = __fmaf_rn(fabsf(oz), 0x1p-20f, __fmul_rn(fabsf(sz), 0x1p-6f));
    constexpr bool MNX = NX_MAGIC_PLANES & 1, MNY = NX_MAGIC_PLANES & 2, MNZ = NX_MAGIC_PLANES & 4;
    constexpr bool MFX = NX_MAGIC_PLANES & 8, MFY = NX_MAGIC_PLANES & 16, MFZ = NX_MAGIC_PLANES & 32;
    float nx_ = __fsub_rn(ox, ex), ny_ = __fsub_rn(oy, ey), nz_ = __fsub_rn(oz, ez);      // near-plane addends
    float fx_ = __fadd_rn(ox, ex), fy_ = __fadd_rn(oy, ey), fz_ = __fadd_rn(oz, ez);      // far-plane addends
    if (MNX) nx_ = __fmaf_rn(-32768.0f, sx, nx_);
    if (MNY) ny_ = __fmaf_rn(-32768.0f, sy, ny_);
    if (MNZ) nz_ = __fmaf_rn(-32768.0f, sz, nz_);
    if (MFX) fx_ = __fmaf_rn(-32768.0f, sx, fx_);
    if (MFY) fy_ = __fmaf_rn(-32768.0f, sy, fy_);
    if (MFZ) fz_ = __fmaf_rn(-32768.0f, sz, fz_);
    const bool negx = inv.x < 0.f, negy = inv.y < 0.f, negz = inv.z < 0.f;                // sign(inv) == sign(d), -0 included
    uint32_t hits = 0;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const uint32_t meta4 = __float_as_uint(h ? n1.w : n1.z);
        const uint32_t inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;                 // bit 4 of a byte set <=> low5 >= 24 <=> inner child
        const uint32_t innerFF = (inner4 >> 4) * 0xffu;                               // 0xff in the bytes of inner children
        const uint32_t bitIdx4 = meta4 ^ (octinv4 & innerFF);                         // low 5 bits per byte: inner 24 + (slot ^ octinv); leaf first triangle bit
        const uint32_t bits4 = (meta4 >> 5) & 0x07070707u;                            // inner: 1; leaf: unary triangle count
        const uint32_t lox = __float_as_uint(h ? n2.y : n2.x), loy = __float_as_uint(h ? n2.w : n2.z), loz = __float_as_uint(h ? n3.y : n3.x);
        const uint32_t hix = __float_as_uint(h ? n3.w : n3.z), hiy = __float_as_uint(h ? n4.y : n4.x), hiz = __float_as_uint(h ? n4.w : n4.z);
        const uint32_t nearx = negx ? hix : lox, farx = negx ? lox : hix;
        const uint32_t neary = negy ? hiy : loy, fary = negy ? loy : hiy;
        const uint32_t nearz = negz ? hiz : loz, farz = negz ? loz : hiz;
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            float t0x, t1x, t0y, t1y, t0z, t1z;
            fma2_bcast(t0x, t1x, plane_q<MNX>(nearx, j, k47), plane_q<MFX>(farx, j, k47), sx, nx_, fx_);
            fma2_bcast(t0y, t1y, plane_q<MNY>(neary, j, k47), plane_q<MFY>(fary, j, k47), sy, ny_, fy_);
            fma2_bcast(t0z, t1z, plane_q<MNZ>(nearz, j, k47), plane_q<MFZ>(farz, j, k47), sz, nz_, fz_);
            const float tn = fmaxf(fmaxf(t0x, t0y), fmaxf(t0z, 0.0f));
            const float tf = fminf(fminf(t1x, t1y), fminf(t1z, tmax));
            // child bits shifted to the child's position: the shift uses the low 5 bits of its count only
            const uint32_t c = shl_wrap(__byte_perm(bits4, 0u, 0x4440u | (uint32_t)j), bitIdx4 >> (8 * j));
            if (tn <= tf) hits |= c;
        }
    }
    inner = make_uint2(__float_as_uint(n1.x), (hits & 0xff000000u) | (eim >> 24));
    leaves = make_uint2(__float_as_uint(n1.y), hits & 0x00ffffffu);
}

__device__ __forceinline__ V3 xform_point(float4 r0, float4 r1, float4 r2, V3 p)
{
    return v3(__fmaf_rn(r0.x, p.x, __fmaf_rn(r0.y, p.y, __fmaf_rn(r0.z, p.z, r0.w))),
              __fmaf_rn(r1.x, p.x, __fmaf_rn(r1.y, p.y, __fmaf_rn(r1.z, p.z, r1.w))),
              __fmaf_rn(r2.x, p.x, __fmaf_rn(r2.y, p.y, __fmaf_rn(r2.z, p.z, r2.w))));
}
__device__ __forceinline__ V3 xform_vector(float4 r0, float4 r1, float4 r2, V3 p)
{
    return v3(__fmaf_rn(r0.x, p.x, __fmaf_rn(r0.y, p.y, __fmul_rn(r0.z, p.z))),
              __fmaf_rn(r1.x, p.x, __fmaf_rn(r1.y, p.y, __fmul_rn(r1.z, p.z))),
              __fmaf_rn(r2.x, p.x, __fmaf_rn(r2.y, p.y, __fmul_rn(r2.z, p.z))));
}

// The ray a triangle of the merged BLAS is tested with: the world-space ray taken into the object space of the triangle's instance,
// exactly as an instance entry does it (same operations, same order: the two-level scene's o', d').
__device__ __forceinline__ void merged_object_ray(const TraceScene& sc, uint32_t inst, V3& o, V3& d)
{
    const float4* m = sc.instInv + 3 * (size_t)inst;
    const float4 r0 = __ldg(m), r1 = __ldg(m + 1), r2 = __ldg(m + 2);
    const V3 wo = o, wd = d;
    o = xform_point(r0, r1, r2, wo); d = xform_vector(r0, r1, r2, wd);
}

// Warp-granular dynamic fetch: a warp reserves 32 queue slots with one atomic and hands them to lanes as they finish, so
// lanes never idle while the queue still has rays (the reference fetches one ray per atomic, BVH8Traversal.cuh:179).
struct WarpFetcher {
    uint32_t next = 0, end = 0;   // warp-uniform
    __device__ __forceinline__ uint32_t take(uint32_t* cursor, uint32_t mask, uint32_t lane_lt)
    {
        const uint32_t cnt = __popc(mask), rank = __popc(mask & lane_lt);
        const uint32_t avail = end - next;
        uint32_t fresh = 0;
        if (cnt > avail) {
            if (lane_id() == 0) fresh = atomicAdd(cursor, 32u);
            fresh = __shfl_sync(NX_FULL, fresh, 0);
        }
        const uint32_t idx = rank < avail ? next + rank : fresh + (rank - avail);
        if (cnt > avail) { next = fresh + (cnt - avail); end = fresh + 32u; } else next += cnt;
        return idx;
    }
};

// The traversal loop, shared by the closest-hit and the any-hit kernels.
//   Sink::finish(rayIdx, pad, t, u, v, prim, slot, occluded) is called once per ray.
// KIND specialises it for what the scene holds (the host picks the kernel; results are identical, the general loop handles everything):
//   NX_SCENE_MIXED      TLAS over instances with their own BLAS and the merged BLAS;
//   NX_SCENE_TWO_LEVEL  no merged BLAS: the merged-triangle path is compiled out of the triangle test;
//   NX_SCENE_DIRECT     the merged BLAS is the whole scene: rays start inside it, there is no TLAS, no instance entry or exit, no parked
//                       ray, and the node / triangle pointers are launch constants instead of per-lane registers.
#define NX_SCENE_MIXED 0
#define NX_SCENE_TWO_LEVEL 1
#define NX_SCENE_DIRECT 2
template <bool ANY_HIT, bool STATS, int KIND, typename Sink>
__device__ __forceinline__ void trace_loop(const TraceScene& sc, const nx_ray* __restrict__ rays, uint32_t n, uint32_t* cursor, TraceTuning tune,
                                           uint32_t* smem, Sink& sink, TraceStats* stats)
{
    uint2* const sstack = reinterpret_cast<uint2*>(smem) + threadIdx.x;                       // entry e at sstack[e * NX_TRACE_BLOCK]
    float4* const park4 = reinterpret_cast<float4*>(smem + 2 * NX_STACK_SHARED * NX_TRACE_BLOCK) + 3 * threadIdx.x;
    uint2 spill[NX_STACK_TOTAL - NX_STACK_SHARED];
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    // 0x47000000 arrives as a kernel parameter so that ptxas cannot fold it: PRMT then takes the constant from the
    // parameter bank / a register and its selector as the immediate, and the byte -> float splice is ONE instruction (with
    // the constant known, ptxas makes IT the immediate and reloads the selector into a register before every PRMT)
    const uint32_t k47 = tune.k47;

    // per-lane state (registers)
    V3 o = v3(0, 0, 0), d = v3(0, 0, 1), inv = v3(0, 0, 0);
    float tmax = 0.f, hitT = NX_MISS_T, hitU = 0.f, hitV = 0.f;
    uint32_t hitPrim = NX_INVALID, hitSlot = NX_INVALID;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    const float4* nodes = KIND == NX_SCENE_DIRECT ? sc.mNodes : sc.tlasNodes; const float4* ltris = KIND == NX_SCENE_DIRECT ? sc.mLtris : nullptr;
    uint32_t octinv4 = 0, curSlot = NX_INVALID, rayIdx = 0, rayPad = 0;
    int sp = 0, instDepth = KIND == NX_SCENE_DIRECT ? 0 : -1;
    bool live = false, dead = false, occluded = false;
#ifdef NX_TRACE_WATCHDOG
    uint32_t wdSteps = 0;
#endif
    WarpFetcher fetch;
    unsigned long long cN = 0, cT = 0, cI = 0, cR = 0, cS = 0;
    unsigned long long wIt = 0, wLN = 0, wRT = 0, wLT = 0, wRX = 0, wLX = 0;   // lane 0 only

    // deeper than NX_STACK_TOTAL entries (never seen on a built tree; the reference's 32-entry stack has no check at all): the entry
    // is refused and counted, and the host reports the count as an error (nx_last_error)
    auto push = [&](uint2 v) { if (sp < NX_STACK_SHARED) sstack[sp * NX_TRACE_BLOCK] = v; else if (sp < (int)tune.stackLimit) spill[sp - NX_STACK_SHARED] = v; else { atomicAdd(sc.overflow, 1u); return; } sp++; };
    auto pop = [&]() -> uint2 { sp--; return sp < NX_STACK_SHARED ? sstack[sp * NX_TRACE_BLOCK] : spill[sp - NX_STACK_SHARED]; };

    // One Moeller-Trumbore test for the highest set bit of the lane's triangle group; {v0, e0 = v1 - v0, e1 = v2 - v0},
    // no back-face culling (Triangle.cuh:29-62).
    auto test_triangle = [&]() {
        const uint32_t bit = 31u - __clz(tgroup.y);
        tgroup.y &= ~(1u << bit);
        if (STATS) cT++;
        const float4* tri = ltris + 3 * (size_t)(tgroup.x + bit);
        const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
        V3 to = o, td = d;
        const bool inMerged = KIND == NX_SCENE_DIRECT || (KIND == NX_SCENE_MIXED && curSlot == sc.mergedSlot);
        if (inMerged) merged_object_ray(sc, __float_as_uint(b.w), to, td);
        const V3 e0 = v3(b.x, b.y, b.z), e1 = v3(c.x, c.y, c.z);
        const V3 pv = xcross(td, e1);
        const float det = xdot(e0, pv);
        const float invDet = rcp_ieee(det);
        const V3 s = to - v3(a.x, a.y, a.z);
        const float u = __fmul_rn(invDet, xdot(s, pv));
        const V3 qv = xcross(s, e0);
        const float v = __fmul_rn(invDet, xdot(td, qv));
        const float t = __fmul_rn(invDet, xdot(e1, qv));
        if (u >= 0.0f && u <= 1.0f && v >= 0.0f && __fadd_rn(u, v) <= 1.0f && t > 0.0f)
        {
            const uint32_t prim = __float_as_uint(a.w);
            if (ANY_HIT) { if (t < tmax) occluded = true; }
            else {
                const uint32_t here = inMerged ? (0x80000000u | __float_as_uint(b.w)) : curSlot;
                bool take = t < fminf(tmax, hitT);
                if (!take && t == hitT && hitPrim != NX_INVALID) {
                    // exact tie: the smaller (instance id, primitive id) wins, whatever the visiting order
                    const uint32_t ia = hit_instance(sc, here), ib = hit_instance(sc, hitSlot);
                    take = ia < ib || (ia == ib && prim < hitPrim);
                }
                if (take) { hitT = t; hitU = u; hitV = v; hitPrim = prim; hitSlot = here; }
            }
        }
    };

    while (true)
    {
        // ---------------------------------------------------------------- phase P: retire / pop (cheap, every iteration) ----
        if (live && ((ANY_HIT && occluded) || (!(ngroup.y & 0xff000000u) && !tgroup.y)))
        {
            if (sp == 0 || (ANY_HIT && occluded)) {
                sink.finish(sc, rayIdx, rayPad, hitT, hitU, hitV, hitPrim, hitSlot, occluded);
                if (STATS) cR++;
                live = false; ngroup = make_uint2(0u, 0u); tgroup = make_uint2(0u, 0u);
            } else {
                if (KIND != NX_SCENE_DIRECT && sp == instDepth) {      // leaving an instance: restore the parked world-space ray
                    const float4 p0 = park4[0], p1 = park4[1], p2 = park4[2];
                    o = v3(p0.x, p0.y, p0.z); d = v3(p1.x, p1.y, p1.z); inv = v3(p2.x, p2.y, p2.z);
                    octinv4 = __float_as_uint(p0.w);
                    nodes = sc.tlasNodes; instDepth = -1;
                }
                const uint2 e = pop();
                if (e.y & 0xff000000u) ngroup = e; else tgroup = e;
            }
        }
        // what every lane could do next; phases other than N run when enough lanes want them or nobody has a node to test
        const bool hasN = live && (ngroup.y & 0xff000000u) != 0u;
        const bool needR = !live && !dead;
        const bool wantI = KIND != NX_SCENE_DIRECT && live && instDepth < 0 && tgroup.y != 0u;
        const uint32_t mN = __ballot_sync(NX_FULL, hasN);
        const uint32_t mX = __ballot_sync(NX_FULL, needR || wantI);

        // ---------------------------------------------------------------- phase X: new ray / enter an instance ----
        // Both end in the same reciprocal-direction + octant set-up, so they share it.
        if (STATS) { wIt++; wLN += __popc(mN); }
        if (mX && (__popc(mX) >= tune.instLanes || mN == 0u))
        {
            if (STATS) { wRX++; wLX += __popc(mX); }
            const uint32_t mR = __ballot_sync(NX_FULL, needR);
            uint32_t got = 0;
            if (mR) got = fetch.take(cursor, mR, lane_lt);
            bool setup = false, bad = false;
            if (needR) {
                if (got < n) {
                    const float4* r = reinterpret_cast<const float4*>(rays + got);
                    const float4 a = __ldg(r), b = __ldg(r + 1);
                    o = v3(a.x, a.y, a.z); d = v3(b.x, b.y, b.z); tmax = a.w;
                    rayIdx = got; rayPad = __float_as_uint(b.w);
#ifdef NX_TRACE_WATCHDOG
                    wdSteps = 0;
#endif
                    hitT = NX_MISS_T; hitU = hitV = 0.f; hitPrim = NX_INVALID; hitSlot = NX_INVALID; occluded = false;
                    sp = 0;
                    if (KIND != NX_SCENE_DIRECT) { nodes = sc.tlasNodes; curSlot = NX_INVALID; instDepth = -1; }
                    // the merged BLAS is the only TLAS entry: rays start inside it (its nodes are in world space)
                    if (KIND == NX_SCENE_MIXED && sc.direct) { nodes = sc.mNodes; ltris = sc.mLtris; curSlot = sc.mergedSlot; instDepth = 0; }
                    live = true; setup = true;
                    // a ray with a non-finite origin or direction passes every conservative box test (NaNs drop out of min / max) and would
                    // walk the whole scene: it is a miss, and it is counted
                    bad = !(fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z) < 3.0e38f);
                    if (bad) atomicAdd(sc.overflow + 1, 1u);
                } else dead = true;
            } else if (KIND != NX_SCENE_DIRECT && wantI) {
                // first instance of the group whose bounding sphere the ray can reach before its current limit
                const float dd = xdot(d, d), limit = ANY_HIT ? tmax : fminf(tmax, hitT);
                const float dlen = sqrt_fast(dd), far = limit * dd * 1.0001f;   // the 1e-4 slack also covers the approximate root
                uint32_t bit = 0; bool found = false;
                while (tgroup.y && !found) {
                    bit = 31u - __clz(tgroup.y);
                    tgroup.y &= ~(1u << bit);
                    const float4 sp4 = __ldg(&sc.inst[tgroup.x + bit].sphere);
                    const V3 oc = v3(sp4.x - o.x, sp4.y - o.y, sp4.z - o.z);
                    const float b = xdot(oc, d), c2 = xdot(oc, oc), r2 = sp4.w * sp4.w;
                    // miss if the closest approach is outside the sphere (slack covers rounding), if the sphere lies behind
                    // the origin, or if it starts beyond the current limit
                    const bool miss = (c2 * dd - b * b) > (r2 + 1.0e-4f * c2) * dd || (b < 0.0f && c2 > r2) || (b - sp4.w * dlen) > far;
                    found = !miss || !tune.sphereCull;
                    if (STATS && !found) cS++;
                }
                if (found) {
                if (tgroup.y) push(tgroup);
                if (ngroup.y & 0xff000000u) push(ngroup);
                instDepth = sp;
                curSlot = tgroup.x + bit;
                const DTravInst* I = sc.inst + curSlot;
                const float4 r0 = __ldg(&I->r0), r1 = __ldg(&I->r1), r2 = __ldg(&I->r2);
                const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&I->nodes));
                nodes = reinterpret_cast<const float4*>(((uint64_t)ptrs.y << 32) | ptrs.x);
                ltris = reinterpret_cast<const float4*>(((uint64_t)ptrs.w << 32) | ptrs.z);
                park4[0] = make_float4(o.x, o.y, o.z, __uint_as_float(octinv4));
                park4[1] = make_float4(d.x, d.y, d.z, 0.f);
                park4[2] = make_float4(inv.x, inv.y, inv.z, 0.f);
                const V3 wo = o, wd = d;
                o = xform_point(r0, r1, r2, wo); d = xform_vector(r0, r1, r2, wd);   // direction is not renormalised: t stays in world units
                if (STATS) cI++;
                setup = true;
                }
            }
            __syncwarp();   // new rays and instance entries reconverge here: one pass through the shared set-up, not one per branch
            if (setup) {
                inv = v3(rcp_dir(d.x), rcp_dir(d.y), rcp_dir(d.z));
                octinv4 = octant_inv4(inv);
                ngroup = make_uint2(0u, bad ? 0u : 0x80000000u); tgroup = make_uint2(0u, 0u);
            }
            if (__all_sync(NX_FULL, dead)) break;
        }

        // ---------------------------------------------------------------- phase N: one node per lane ----
        if (live && (ngroup.y & 0xff000000u))
        {
            if (tgroup.y) { push(tgroup); tgroup = make_uint2(0u, 0u); }     // postponed triangles / instances wait on the stack
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) push(ngroup);
            const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
            const uint32_t child = ngroup.x + __popc(ngroup.y & ((1u << slot) - 1u) & 0xffu);
            intersect_children(nodes, child, o, inv, octinv4, ANY_HIT ? tmax : fminf(tmax, hitT), k47, ngroup, tgroup);
            if (STATS) cN++;
#ifdef NX_TRACE_WATCHDOG
            if (++wdSteps == (1u << 18)) printf("WD ray %u o %.9g %.9g %.9g d %.9g %.9g %.9g tmax %g hitT %g inv %g %g %g sp %d inst %d child %u\n", rayIdx, o.x, o.y, o.z, d.x, d.y, d.z, tmax, hitT, inv.x, inv.y, inv.z, sp, instDepth, child);
#endif
        }

        // ---------------------------------------------------------------- phase T: triangles (rounds until too few lanes) ----
        while (true)
        {
            const bool wantT = live && (KIND == NX_SCENE_DIRECT || instDepth >= 0) && tgroup.y != 0u && !(ANY_HIT && occluded);
            const uint32_t mT = __ballot_sync(NX_FULL, wantT);
            if (!mT) break;
            // lanes that still have a node to test after this one keep the warp busy; otherwise triangles are all there is
            const uint32_t mN2 = __ballot_sync(NX_FULL, live && (ngroup.y & 0xff000000u) != 0u);
            if (__popc(mT) < tune.triLanes && mN2 != 0u) break;
            if (STATS) { wRT++; wLT += __popc(mT); }
            if (wantT) test_triangle();
        }
    }
    if (STATS) {
        atomicAdd(&stats->nodes, cN); atomicAdd(&stats->tris, cT); atomicAdd(&stats->insts, cI); atomicAdd(&stats->rays, cR); atomicAdd(&stats->sphereCulled, cS);
        if (lane_id() == 0) {
            atomicAdd(&stats->iters, wIt); atomicAdd(&stats->lanesN, wLN); atomicAdd(&stats->roundsT, wRT); atomicAdd(&stats->lanesT, wLT);
            atomicAdd(&stats->roundsX, wRX); atomicAdd(&stats->lanesX, wLX);
        }
    }
}
