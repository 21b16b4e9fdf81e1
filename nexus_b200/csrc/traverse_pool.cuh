// Ray-pool traversal: a warp owns MORE rays than it has lanes.
//
// Same algorithm and same arithmetic as trace_loop() in traverse.cuh (which follows BVH8Trace / BVH8TraceShadow of the
// reference, src/Cuda/BVH/BVH8Traversal.cuh:149-324, 343-521): identical slab test (intersect_children), identical
// Moeller-Trumbore test, identical (instance id, primitive id) tie-break, so the hits are the same bit for bit.  What changes
// is who executes what.  trace_loop() binds a ray to a lane for its whole life, and ncu showed what that costs
// (profiles/r01_ncu_trace_closest_after.md): 19-21 of 32 lanes per instruction, because the triangle and the set-up phases
// can only run on the lanes whose OWN ray wants them (8 of 32 on average) while everybody else idles.
//
// Here the state of R = 64 rays per warp lives in shared memory (structure of float4 arrays, one element per pool slot).
// Every round the warp counts how many of its rays want each of the four kinds of work - FETCH a new ray, test a NODE,
// test a TRIangle, enter an INSTance -, picks one kind, hands the first 32 takers to its 32 lanes (ballot + popc ranks
// through a 32-entry table), and the lanes load just the slice of state that kind of work needs, run it and store what changed.
// A round therefore runs at (nearly) 32 lanes whatever it is, and rays that wait for a triangle round cost a pool slot,
// not an idle lane.  Registers no longer limit occupancy (no per-ray state survives a round); shared memory does:
// 212 B per ray, 13.7 KB per warp, 16 warps per SM.
//
// Because a ray is parked between rounds anyway, the address of the next node is known long before it is tested: the tail of
// every round prefetches it (the lane-bound loop had nothing to overlap a prefetch with).
#pragma once
#include "traverse.cuh"

#ifndef NX_POOL_WARPS
#define NX_POOL_WARPS 4          // warps per CTA
#endif
#ifndef NX_POOL_MIN_BLOCKS
#define NX_POOL_MIN_BLOCKS 4     // CTAs per SM the register allocation must allow (shared memory allows 4 at R = 64)
#endif
#ifndef NX_POOL_STACK
#define NX_POOL_STACK 8          // traversal-stack entries per ray in shared memory; deeper entries go to a global spill area
#endif
#ifndef NX_POOL_PREFETCH
#define NX_POOL_PREFETCH 1
#endif
#define NX_POOL_R 64u
#define NX_POOL_BLOCK (32 * NX_POOL_WARPS)
#define NX_POOL_SPILL (NX_STACK_TOTAL - NX_POOL_STACK)

enum : uint32_t { PS_FREE = 0u, PS_NODE = 1u, PS_TRI = 2u, PS_INST = 3u };

struct PoolTuning {
    uint32_t nodeLanes, triLanes, instLanes, fetchLanes;   // run that kind of round when at least this many rays want it
    uint32_t sphereCull, k47;
    uint32_t stackLimit;                                   // entries per ray before a push is refused (NX_STACK_TOTAL; lower only in tests)
};

struct __align__(16) PoolWarp {            // shared memory of one warp; [slot]
    float4 Wo[NX_POOL_R];                  // world-space origin, tmax
    float4 Wd[NX_POOL_R];                  // world-space direction, nx_ray::pad (pixel index)
    float4 Co[NX_POOL_R];                  // origin in the current space (world or the entered instance's), current limit min(tmax, hit t)
    float4 Cd[NX_POOL_R];                  // direction in the current space
    float4 Ci[NX_POOL_R];                  // reciprocal direction in the current space, octant word
    uint4 G[NX_POOL_R];                    // inner-node group (childBase, hits << 24 | imask), leaf group (primBase, hit bits)
    uint4 P[NX_POOL_R];                    // node pointer of the current level, leaf-triangle pointer of the entered instance
    uint4 M[NX_POOL_R];                    // sp | instDepth << 8 (0xff: top level), current TLAS slot, ray index, TLAS slot of the hit
    float4 H[NX_POOL_R];                   // hit t, u, v, primitive id
    uint2 stk[NX_POOL_STACK][NX_POOL_R];
    uint32_t st[NX_POOL_R];                // PS_*
    uint32_t assign[32];                   // lane -> slot of this round
};
#define NX_POOL_SMEM_BYTES (sizeof(PoolWarp) * NX_POOL_WARPS)

__device__ __forceinline__ const float4* ptr_from(uint32_t lo, uint32_t hi) { return reinterpret_cast<const float4*>(((uint64_t)hi << 32) | lo); }

template <bool ANY_HIT, bool STATS, typename Sink>
__device__ __forceinline__ void trace_pool_loop(const TraceScene& sc, const nx_ray* __restrict__ rays, uint32_t n, uint32_t* cursor, PoolTuning tune,
                                                PoolWarp& pw, Sink& sink, TraceStats* stats, uint2* __restrict__ spill)
{
    const uint32_t lane = lane_id();
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    const uint32_t k47 = tune.k47;
    const uint32_t tlasLo = (uint32_t)reinterpret_cast<uint64_t>(sc.tlasNodes), tlasHi = (uint32_t)(reinterpret_cast<uint64_t>(sc.tlasNodes) >> 32);
    pw.st[lane] = PS_FREE; pw.st[lane + 32u] = PS_FREE;
    bool exhausted = false;                                                    // warp-uniform: the ray queue has nothing left
    unsigned long long cN = 0, cT = 0, cI = 0, cR = 0, cS = 0;
    unsigned long long wIt = 0, wRN = 0, wLN = 0, wRT = 0, wLT = 0, wRX = 0, wLX = 0, wRF = 0, wLF = 0;   // lane 0 only

    while (true)
    {
        __syncwarp();
        // ------------------------------------------------------------------ what do the 64 rays want? ----
        const uint32_t s0 = pw.st[lane], s1 = pw.st[lane + 32u];
        const uint32_t a0 = __ballot_sync(NX_FULL, s0 & 1u), b0 = __ballot_sync(NX_FULL, s0 & 2u);
        const uint32_t a1 = __ballot_sync(NX_FULL, s1 & 1u), b1 = __ballot_sync(NX_FULL, s1 & 2u);
        const uint32_t qN = __popc(a0 & ~b0) + __popc(a1 & ~b1), qT = __popc(b0 & ~a0) + __popc(b1 & ~a1), qI = __popc(a0 & b0) + __popc(a1 & b1);
        const uint32_t qF = exhausted ? 0u : 64u - qN - qT - qI;
        uint32_t ph;
        if (qF >= tune.fetchLanes) ph = PS_FREE;
        else if (qN >= tune.nodeLanes) ph = PS_NODE;
        else if (qT >= tune.triLanes) ph = PS_TRI;
        else if (qI >= tune.instLanes) ph = PS_INST;
        else {
            uint32_t best = qN; ph = PS_NODE;
            if (qT > best) { best = qT; ph = PS_TRI; }
            if (qI > best) { best = qI; ph = PS_INST; }
            if (qF > best) { best = qF; ph = PS_FREE; }
            if (best == 0u) break;                                             // every slot is free and the queue is empty
        }
        // ------------------------------------------------------------------ hand the first 32 takers to the lanes ----
        const uint32_t x0 = (ph & 1u) ? a0 : ~a0, y0 = (ph & 2u) ? b0 : ~b0, x1 = (ph & 1u) ? a1 : ~a1, y1 = (ph & 2u) ? b1 : ~b1;
        const uint32_t m0 = x0 & y0, m1 = x1 & y1, c0 = __popc(m0);
        if ((m0 >> lane) & 1u) pw.assign[__popc(m0 & lane_lt)] = lane;
        { const uint32_t r1 = c0 + __popc(m1 & lane_lt); if (((m1 >> lane) & 1u) && r1 < 32u) pw.assign[r1] = lane + 32u; }
        __syncwarp();
        const uint32_t cnt = min(32u, c0 + __popc(m1));
        const bool active = lane < cnt;
        const uint32_t s = active ? pw.assign[lane] : 0u;
        if (STATS) wIt++;

        // ------------------------------------------------------------------ FETCH: up to 32 new rays, one atomic ----
        if (ph == PS_FREE)
        {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, cnt);
            base = __shfl_sync(NX_FULL, base, 0);
            if (base + cnt >= n) exhausted = true;
            const uint32_t idx = base + lane;
            if (STATS) { wRF++; wLF += cnt; }
            if (active && idx < n)
            {
                const float4* r = reinterpret_cast<const float4*>(rays + idx);
                const float4 a = __ldg(r), b = __ldg(r + 1);
                const V3 inv = v3(rcp_dir(b.x), rcp_dir(b.y), rcp_dir(b.z));
                pw.Wo[s] = a; pw.Wd[s] = b;
                pw.Co[s] = make_float4(a.x, a.y, a.z, ANY_HIT ? a.w : fminf(a.w, NX_MISS_T)); pw.Cd[s] = b;
                pw.Ci[s] = make_float4(inv.x, inv.y, inv.z, __uint_as_float(octant_inv4(inv)));
                pw.G[s] = make_uint4(0u, 0x80000000u, 0u, 0u);
                if (sc.direct) {      // the merged BLAS is the whole scene: start inside it (instance depth 0, world space is its object space)
                    const uint64_t mn = reinterpret_cast<uint64_t>(sc.mNodes), ml = reinterpret_cast<uint64_t>(sc.mLtris);
                    pw.P[s] = make_uint4((uint32_t)mn, (uint32_t)(mn >> 32), (uint32_t)ml, (uint32_t)(ml >> 32));
                    pw.M[s] = make_uint4(0x0000u, sc.mergedSlot, idx, NX_INVALID);
                } else {
                    pw.P[s] = make_uint4(tlasLo, tlasHi, 0u, 0u);
                    pw.M[s] = make_uint4(0xff00u, NX_INVALID, idx, NX_INVALID);
                }
                pw.H[s] = make_float4(NX_MISS_T, 0.f, 0.f, __uint_as_float(NX_INVALID));
                pw.st[s] = PS_NODE;
                if (NX_POOL_PREFETCH) { prefetch_l1(sc.tlasNodes); }
            }
            continue;
        }

        // ------------------------------------------------------------------ NODE / TRI / INST on the lane's ray ----
        if (active)
        {
            const uint4 g = pw.G[s];
            uint4 m = pw.M[s];
            float4 co = pw.Co[s];
            uint2 ng = make_uint2(g.x, g.y), tg = make_uint2(g.z, g.w);
            uint32_t sp = m.x & 0xffu, idp = (m.x >> 8) & 0xffu;
            bool occluded = false;
            uint32_t pfLo = 0, pfHi = 0, pfOct = 0;                              // node pointer + octant word for the prefetch in the tail

            auto push = [&](uint2 v) {
                if (sp >= tune.stackLimit) { atomicAdd(sc.overflow, 1u); return; }   // never seen on a built tree; reported through nx_last_error
                if (sp < NX_POOL_STACK) pw.stk[sp][s] = v; else spill[(size_t)(sp - NX_POOL_STACK) * NX_POOL_R + s] = v;
                sp++;
            };
            auto pop = [&]() -> uint2 { sp--; return sp < NX_POOL_STACK ? pw.stk[sp][s] : spill[(size_t)(sp - NX_POOL_STACK) * NX_POOL_R + s]; };

            if (ph == PS_NODE)
            {
                const float4 ci = pw.Ci[s];
                const uint2 np = *reinterpret_cast<const uint2*>(&pw.P[s]);
                const uint32_t octinv4 = __float_as_uint(ci.w);
                const uint32_t bit = 31u - __clz(ng.y);
                ng.y &= ~(1u << bit);
                if (ng.y & 0xff000000u) push(ng);
                const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
                const uint32_t child = ng.x + __popc(ng.y & ((1u << slot) - 1u) & 0xffu);
                intersect_children(ptr_from(np.x, np.y), child, v3(co.x, co.y, co.z), v3(ci.x, ci.y, ci.z), octinv4, co.w, k47, ng, tg);
                pfLo = np.x; pfHi = np.y; pfOct = octinv4;
                if (STATS) cN++;
            }
            else if (ph == PS_TRI)
            {
                // one Moeller-Trumbore test (Triangle.cuh:29-62), {v0 | primId, e0, e1} records in leaf order
                const float4 cd = pw.Cd[s];
                const uint2 lp = reinterpret_cast<const uint2*>(&pw.P[s])[1];
                const uint32_t bit = 31u - __clz(tg.y);
                tg.y &= ~(1u << bit);
                if (STATS) cT++;
                const float4* tri = ptr_from(lp.x, lp.y) + 3 * (size_t)(tg.x + bit);
                const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
                V3 o = v3(co.x, co.y, co.z), d = v3(cd.x, cd.y, cd.z);
                if (m.y == sc.mergedSlot) merged_object_ray(sc, __float_as_uint(b.w), o, d);   // traverse.cuh: object-space test inside the merged BLAS
                const V3 e0 = v3(b.x, b.y, b.z), e1 = v3(c.x, c.y, c.z);
                const V3 pv = xcross(d, e1);
                const float det = xdot(e0, pv);
                const float invDet = rcp_ieee(det);
                const V3 sv = o - v3(a.x, a.y, a.z);
                const float u = __fmul_rn(invDet, xdot(sv, pv));
                const V3 qv = xcross(sv, e0);
                const float v = __fmul_rn(invDet, xdot(d, qv));
                const float t = __fmul_rn(invDet, xdot(e1, qv));
                if (u >= 0.0f && u <= 1.0f && v >= 0.0f && __fadd_rn(u, v) <= 1.0f && t > 0.0f)
                {
                    if (ANY_HIT) { if (t < co.w) occluded = true; }
                    else {
                        const uint32_t prim = __float_as_uint(a.w);
                        const uint32_t here = m.y == sc.mergedSlot ? (0x80000000u | __float_as_uint(b.w)) : m.y;
                        bool take = t < co.w;                                  // co.w = min(tmax, hit t)
                        if (!take && t == co.w) {
                            // exact tie with the hit so far: the smaller (instance id, primitive id) wins, whatever the visiting order
                            const float4 h = pw.H[s];
                            if (__float_as_uint(h.w) != NX_INVALID && t == h.x) {
                                const uint32_t ia = hit_instance(sc, here), ib = hit_instance(sc, m.w);
                                take = ia < ib || (ia == ib && prim < __float_as_uint(h.w));
                            }
                        }
                        if (take) { pw.H[s] = make_float4(t, u, v, a.w); m.w = here; co.w = t; pw.Co[s].w = t; }
                    }
                }
            }
            else
            {
                // first instance of the group whose bounding sphere the ray can reach before its current limit, then enter it
                const float4 cd = pw.Cd[s];
                const V3 o = v3(co.x, co.y, co.z), d = v3(cd.x, cd.y, cd.z);
                const float dd = xdot(d, d), limit = co.w;
                const float dlen = sqrt_fast(dd), far = limit * dd * 1.0001f;   // the 1e-4 slack also covers the approximate root
                uint32_t bit = 0; bool found = false;
                while (tg.y && !found) {
                    bit = 31u - __clz(tg.y);
                    tg.y &= ~(1u << bit);
                    const float4 sp4 = __ldg(&sc.inst[tg.x + bit].sphere);
                    const V3 oc = v3(sp4.x - o.x, sp4.y - o.y, sp4.z - o.z);
                    const float b = xdot(oc, d), c2 = xdot(oc, oc), r2 = sp4.w * sp4.w;
                    const bool miss = (c2 * dd - b * b) > (r2 + 1.0e-4f * c2) * dd || (b < 0.0f && c2 > r2) || (b - sp4.w * dlen) > far;
                    found = !miss || !tune.sphereCull;
                    if (STATS && !found) cS++;
                }
                if (found) {
                    if (tg.y) push(tg);
                    if (ng.y & 0xff000000u) push(ng);
                    idp = sp;
                    m.y = tg.x + bit;
                    const DTravInst* I = sc.inst + m.y;
                    const float4 r0 = __ldg(&I->r0), r1 = __ldg(&I->r1), r2 = __ldg(&I->r2);
                    const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&I->nodes));
                    const V3 oo = xform_point(r0, r1, r2, o), od = xform_vector(r0, r1, r2, d);   // direction is not renormalised: t stays in world units
                    const V3 inv = v3(rcp_dir(od.x), rcp_dir(od.y), rcp_dir(od.z));
                    pfOct = octant_inv4(inv); pfLo = ptrs.x; pfHi = ptrs.y;
                    pw.P[s] = ptrs;
                    pw.Co[s] = make_float4(oo.x, oo.y, oo.z, limit);
                    pw.Cd[s] = make_float4(od.x, od.y, od.z, 0.f);
                    pw.Ci[s] = make_float4(inv.x, inv.y, inv.z, __uint_as_float(pfOct));
                    ng = make_uint2(0u, 0x80000000u); tg = make_uint2(0u, 0u);
                    if (STATS) cI++;
                }
            }

            // ---- what next for this ray: pending leaves first (they shrink the limit), then inner nodes, then the stack ----
            uint32_t ns;
            bool done = false;
            if (ANY_HIT && occluded) done = true;
            else if (tg.y) ns = idp != 0xffu ? PS_TRI : PS_INST;
            else if (ng.y & 0xff000000u) ns = PS_NODE;
            else if (sp == 0u) done = true;
            else {
                if (sp == idp) {      // leaving the instance: back to the world-space ray
                    const float4 wo = pw.Wo[s], wd = pw.Wd[s];
                    const V3 inv = v3(rcp_dir(wd.x), rcp_dir(wd.y), rcp_dir(wd.z));
                    pw.Co[s] = make_float4(wo.x, wo.y, wo.z, co.w);
                    pw.Cd[s] = wd;
                    pw.Ci[s] = make_float4(inv.x, inv.y, inv.z, __uint_as_float(octant_inv4(inv)));
                    *reinterpret_cast<uint2*>(&pw.P[s]) = make_uint2(tlasLo, tlasHi);
                    idp = 0xffu; pfLo = 0u; pfHi = 0u;
                }
                const uint2 e = pop();
                if (e.y & 0xff000000u) { ng = e; ns = PS_NODE; } else { tg = e; ns = idp != 0xffu ? PS_TRI : PS_INST; }
            }
            if (done) {
                const float4 h = pw.H[s], wd = pw.Wd[s];
                sink.finish(sc, m.z, __float_as_uint(wd.w), h.x, h.y, h.z, __float_as_uint(h.w), m.w, occluded);
                if (STATS) cR++;
                ns = PS_FREE;
            } else {
                pw.G[s] = make_uint4(ng.x, ng.y, tg.x, tg.y);
                m.x = sp | (idp << 8);
                pw.M[s] = m;
                if (NX_POOL_PREFETCH && ns == PS_NODE && (pfLo | pfHi)) {
                    // the node this ray tests next (same selection as the NODE round will make)
                    const uint32_t bit = 31u - __clz(ng.y);
                    const uint32_t slot = (bit - 24u) ^ (pfOct & 0xffu);
                    const uint32_t child = ng.x + __popc(ng.y & ~(1u << bit) & ((1u << slot) - 1u) & 0xffu);
                    const float4* nd = ptr_from(pfLo, pfHi) + 5 * (size_t)child;
                    prefetch_l1(nd); prefetch_l1(nd + 4);
                }
            }
            pw.st[s] = ns;
        }
        if (STATS) {
            if (ph == PS_NODE) { wRN++; wLN += cnt; } else if (ph == PS_TRI) { wRT++; wLT += cnt; } else { wRX++; wLX += cnt; }
        }
    }
    if (STATS) {
        atomicAdd(&stats->nodes, cN); atomicAdd(&stats->tris, cT); atomicAdd(&stats->insts, cI); atomicAdd(&stats->rays, cR); atomicAdd(&stats->sphereCulled, cS);
        if (lane == 0) {
            atomicAdd(&stats->iters, wIt); atomicAdd(&stats->lanesN, wLN); atomicAdd(&stats->roundsT, wRT); atomicAdd(&stats->lanesT, wLT);
            atomicAdd(&stats->roundsX, wRX); atomicAdd(&stats->lanesX, wLX);
            atomicAdd(&stats->roundsN, wRN); atomicAdd(&stats->roundsF, wRF); atomicAdd(&stats->lanesF, wLF);
        }
    }
}
