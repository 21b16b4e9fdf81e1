// Two rays per lane ("duo") traversal loop.
//
// Same phased persistent-warp loop, same arithmetic and the same hits as trace_loop() in traverse.cuh (BVH8Trace /
// BVH8TraceShadow of the reference, src/Cuda/BVH/BVH8Traversal.cuh:149-324, 343-521).  The difference is what a lane does when
// its ray cannot take part in the phase the warp runs.  ncu on trace_loop (profiles/r01_ncu_trace_closest_after.md): 19-21 of 32
// lanes per instruction, because a lane whose ray waits for a triangle round or an instance entry idles through the node rounds
// of the others.  The ray-pool loop (traverse_pool.cuh) removes that by re-assigning rays to lanes every round, and pays for it
// with ~150 instructions of scheduling and state traffic per round and half the resident warps: measured 35 % SLOWER
// (profiles/r02_ncu_trace_closest_pool.md).
//
// Here every lane owns TWO rays: one active (state in registers, exactly as in trace_loop) and one parked in shared memory
// (80 bytes).  When the active ray has no node to test and the parked one does, the lane swaps them: 5 LDS.128 + 5 STS.128 and a
// reciprocal, only at that moment and without any cross-lane traffic or scheduling.  Each ray keeps its own traversal stack and
// instance park area (selected by one bit), so nothing but the register state moves.  A lane is idle in the node phase only when
// BOTH of its rays are blocked.
#pragma once
#include "traverse.cuh"


#ifndef NX_DUO_BLOCK
#define NX_DUO_BLOCK 128
#endif
#ifndef NX_DUO_STACK
#define NX_DUO_STACK 6           // shared-memory stack entries per ray
#endif
#ifndef NX_DUO_MIN_BLOCKS
#define NX_DUO_MIN_BLOCKS 6
#endif
// per thread: 2 stacks, 2 instance park areas (3 x float4), the parked ray (5 x float4)
#define NX_DUO_SMEM_BYTES ((2 * NX_DUO_STACK * 8 + 2 * 48 + 80) * NX_DUO_BLOCK)

enum : uint32_t { DW_NONE = 0u, DW_X = 1u, DW_T = 2u, DW_N = 3u };   // what a ray wants next, by priority

template <bool ANY_HIT, bool STATS, typename Sink>
__device__ __forceinline__ void trace_loop_duo(const TraceScene& sc, const nx_ray* __restrict__ rays, uint32_t n, uint32_t* cursor, TraceTuning tune,
                                               uint32_t* smem, Sink& sink, TraceStats* stats)
{
    constexpr int B = NX_DUO_BLOCK, S = NX_DUO_STACK;
    uint2* const stk = reinterpret_cast<uint2*>(smem) + threadIdx.x;                          // entry e of stack c at stk[(c * S + e) * B]
    float4* const park = reinterpret_cast<float4*>(smem + 2 * 2 * S * B) + threadIdx.x;       // row k of park area c at park[(c * 3 + k) * B]
    float4* const rec = reinterpret_cast<float4*>(smem + 2 * 2 * S * B + 2 * 12 * B) + threadIdx.x;   // row k of the parked ray at rec[k * B]
    uint2 spill[2][NX_STACK_TOTAL - S];
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    const uint32_t k47 = tune.k47;

    // active ray (registers)
    V3 o = v3(0, 0, 0), d = v3(0, 0, 1), inv = v3(0, 0, 0);
    float tmax = 0.f, hitT = NX_MISS_T, hitU = 0.f, hitV = 0.f;
    uint32_t hitPrim = NX_INVALID, hitSlot = NX_INVALID;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u);
    const float4* nodes = sc.tlasNodes; const float4* ltris = nullptr;
    uint32_t octinv4 = 0, curSlot = NX_INVALID, rayIdx = 0, rayPad = 0;
    int sp = 0, instDepth = -1;
    bool live = false, dead = false, occluded = false;
    // parked ray: what it wants (DW_*), whether the slot holds a ray; `cur` selects the active ray's stack / park area
    uint32_t bWant = DW_X, cur = 0;
    bool bLive = false;
    rec[4 * B] = make_float4(0.f, 0.f, 0.f, 0.f);                                             // parked slot: no ray
    WarpFetcher fetch;
    unsigned long long cN = 0, cT = 0, cI = 0, cR = 0, cS = 0;
    unsigned long long wIt = 0, wLN = 0, wRT = 0, wLT = 0, wRX = 0, wLX = 0, wSw = 0;

    auto push = [&](uint2 v) {
        if (sp < S) stk[(cur * S + sp) * B] = v;
        else if (sp < (int)tune.stackLimit) spill[cur][sp - S] = v;
        else { atomicAdd(sc.overflow, 1u); return; }
        sp++;
    };
    auto pop = [&]() -> uint2 { sp--; return sp < S ? stk[(cur * S + sp) * B] : spill[cur][sp - S]; };

    auto wantOf = [&]() -> uint32_t {            // the active ray's next step
        if (live) {
            if (ANY_HIT && occluded) return DW_N;   // only waits to be retired: it must never count as wanting triangles (two such rays would swap for ever)
            if (ngroup.y & 0xff000000u) return DW_N;
            if (tgroup.y) return instDepth >= 0 ? DW_T : DW_X;
            return DW_N;                          // nothing pending: the pop at the top of the next iteration decides (treated as runnable)
        }
        return dead ? DW_NONE : DW_X;
    };

    // exchange the active ray with the parked one
    auto swap = [&]() {
        const float4 a0 = make_float4(o.x, o.y, o.z, tmax), a1 = make_float4(d.x, d.y, d.z, __uint_as_float(rayPad));
        const float4 a2 = make_float4(hitT, hitU, hitV, __uint_as_float(hitPrim));
        const float4 a3 = make_float4(__uint_as_float(ngroup.x), __uint_as_float(ngroup.y), __uint_as_float(tgroup.x), __uint_as_float(tgroup.y));
        const float4 a4 = make_float4(__uint_as_float(hitSlot), __uint_as_float(curSlot), __uint_as_float(rayIdx),
                                      __uint_as_float((uint32_t)sp | ((uint32_t)(instDepth + 1) << 8) | (live ? 0x10000u : 0u) | (occluded ? 0x20000u : 0u)));
        const uint32_t aWant = wantOf();
        const bool aLive = live;
        const float4 b0 = rec[0], b1 = rec[B], b2 = rec[2 * B], b3 = rec[3 * B], b4 = rec[4 * B];
        rec[0] = a0; rec[B] = a1; rec[2 * B] = a2; rec[3 * B] = a3; rec[4 * B] = a4;
        o = v3(b0.x, b0.y, b0.z); tmax = b0.w; d = v3(b1.x, b1.y, b1.z); rayPad = __float_as_uint(b1.w);
        hitT = b2.x; hitU = b2.y; hitV = b2.z; hitPrim = __float_as_uint(b2.w);
        ngroup = make_uint2(__float_as_uint(b3.x), __float_as_uint(b3.y)); tgroup = make_uint2(__float_as_uint(b3.z), __float_as_uint(b3.w));
        hitSlot = __float_as_uint(b4.x); curSlot = __float_as_uint(b4.y); rayIdx = __float_as_uint(b4.z);
        const uint32_t f = __float_as_uint(b4.w);
        sp = (int)(f & 0xffu); instDepth = (int)((f >> 8) & 0xffu) - 1; live = (f & 0x10000u) != 0u; occluded = (f & 0x20000u) != 0u;
        cur ^= 1u;
        bWant = aWant; bLive = aLive;
        nodes = sc.tlasNodes;
        if (live) {
            if (instDepth >= 0) {
                const uint4 ptrs = __ldg(reinterpret_cast<const uint4*>(&sc.inst[curSlot].nodes));
                nodes = reinterpret_cast<const float4*>(((uint64_t)ptrs.y << 32) | ptrs.x);
                ltris = reinterpret_cast<const float4*>(((uint64_t)ptrs.w << 32) | ptrs.z);
            }
            inv = v3(rcp_dir(d.x), rcp_dir(d.y), rcp_dir(d.z));       // the same values the set-up phase computed from the same d
            octinv4 = octant_inv4(inv);
        } else { ngroup = make_uint2(0u, 0u); tgroup = make_uint2(0u, 0u); }
        if (STATS) wSw++;
    };

    auto test_triangle = [&]() {
        const uint32_t bit = 31u - __clz(tgroup.y);
        tgroup.y &= ~(1u << bit);
        if (STATS) cT++;
        const float4* tri = ltris + 3 * (size_t)(tgroup.x + bit);
        const float4 a = __ldg(tri), b = __ldg(tri + 1), c = __ldg(tri + 2);
        V3 to = o, td = d;
        if (curSlot == sc.mergedSlot) merged_object_ray(sc, __float_as_uint(b.w), to, td);   // traverse.cuh: object-space test inside the merged BLAS
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
                const uint32_t here = curSlot == sc.mergedSlot ? (0x80000000u | __float_as_uint(b.w)) : curSlot;
                bool take = t < fminf(tmax, hitT);
                if (!take && t == hitT && hitPrim != NX_INVALID) {
                    const uint32_t ia = hit_instance(sc, here), ib = hit_instance(sc, hitSlot);
                    take = ia < ib || (ia == ib && prim < hitPrim);
                }
                if (take) { hitT = t; hitU = u; hitV = v; hitPrim = prim; hitSlot = here; }
            }
        }
    };

#ifdef NX_DUO_WATCHDOG
    unsigned long long wd = 0;
#endif
    while (true)
    {
#ifdef NX_DUO_WATCHDOG
        if (++wd > (1ull << 21)) {
            if (atomicAdd(sc.overflow, 1u << 16) == 0u || (wd & 0xffff) == 1)
                printf("WD blk %d thr %d live %d dead %d bLive %d bWant %u cur %u sp %d inst %d ng %08x/%08x tg %08x/%08x ray %u hitT %g occ %d\n", blockIdx.x, threadIdx.x, (int)live, (int)dead, (int)bLive, bWant, cur, sp,
                       instDepth, ngroup.x, ngroup.y, tgroup.x, tgroup.y, rayIdx, hitT, (int)occluded);
            if (wd > (1ull << 21) + 2) break;
        }
#endif
        // ---------------------------------------------------------------- phase P: retire / pop (active ray) ----
        if (live && ((ANY_HIT && occluded) || (!(ngroup.y & 0xff000000u) && !tgroup.y)))
        {
            if (sp == 0 || (ANY_HIT && occluded)) {
                sink.finish(sc, rayIdx, rayPad, hitT, hitU, hitV, hitPrim, hitSlot, occluded);
                if (STATS) cR++;
                live = false; ngroup = make_uint2(0u, 0u); tgroup = make_uint2(0u, 0u);
            } else {
                if (sp == instDepth) {      // leaving an instance: restore the parked world-space ray
                    const float4 p0 = park[(cur * 3) * B], p1 = park[(cur * 3 + 1) * B], p2 = park[(cur * 3 + 2) * B];
                    o = v3(p0.x, p0.y, p0.z); d = v3(p1.x, p1.y, p1.z); inv = v3(p2.x, p2.y, p2.z);
                    octinv4 = __float_as_uint(p0.w);
                    nodes = sc.tlasNodes; instDepth = -1;
                }
                const uint2 e = pop();
                if (e.y & 0xff000000u) ngroup = e; else tgroup = e;
            }
        }
        // ---------------------------------------------------------------- which of the lane's two rays works this iteration ----
        // a node test always runs, so a parked ray that has one replaces an active ray that has none at once; for the batched phases
        // the parked ray only votes, and is swapped in when its phase actually runs
        bool hasN = live && (ngroup.y & 0xff000000u) != 0u;
        if (!hasN && (bWant == DW_N || (bLive && !live && dead))) { swap(); hasN = live && (ngroup.y & 0xff000000u) != 0u; }
        if (dead && !bLive && bWant == DW_X) bWant = DW_NONE;                     // the empty parked slot will never be refilled
        const bool needR = !live && !dead;
        const bool wantI = live && instDepth < 0 && tgroup.y != 0u;
        const bool parkedX = !hasN && bWant == DW_X && !(needR || wantI);         // the parked slot wants a new ray / an instance entry
        const uint32_t mN = __ballot_sync(NX_FULL, hasN);
        const uint32_t mX = __ballot_sync(NX_FULL, needR || wantI || parkedX);
        if (!__any_sync(NX_FULL, live || bLive || !dead)) break;

        // ---------------------------------------------------------------- phase X: new ray / enter an instance ----
        if (STATS) { wIt++; wLN += __popc(mN); }
        if (mX && (__popc(mX) >= tune.instLanes || mN == 0u))
        {
            if (STATS) { wRX++; wLX += __popc(mX); }
            if (parkedX) swap();                                                  // the blocked active ray is parked, its slot's work comes in
            const bool needR2 = !live && !dead;
            const bool wantI2 = live && instDepth < 0 && tgroup.y != 0u;
            const uint32_t mR = __ballot_sync(NX_FULL, needR2);
            uint32_t got = 0;
            if (mR) got = fetch.take(cursor, mR, lane_lt);
            bool setup = false;
            if (needR2) {
                if (got < n) {
                    const float4* r = reinterpret_cast<const float4*>(rays + got);
                    const float4 a = __ldg(r), b = __ldg(r + 1);
                    o = v3(a.x, a.y, a.z); d = v3(b.x, b.y, b.z); tmax = a.w;
                    rayIdx = got; rayPad = __float_as_uint(b.w);
                    hitT = NX_MISS_T; hitU = hitV = 0.f; hitPrim = NX_INVALID; hitSlot = NX_INVALID; occluded = false;
                    nodes = sc.tlasNodes; curSlot = NX_INVALID; instDepth = -1; sp = 0;
                    if (sc.direct) { nodes = sc.mNodes; ltris = sc.mLtris; curSlot = sc.mergedSlot; instDepth = 0; }
                    live = true; setup = true;
                } else dead = true;
            } else if (wantI2) {
                const float dd = xdot(d, d), limit = ANY_HIT ? tmax : fminf(tmax, hitT);
                const float dlen = sqrt_fast(dd), far = limit * dd * 1.0001f;
                uint32_t bit = 0; bool found = false;
                while (tgroup.y && !found) {
                    bit = 31u - __clz(tgroup.y);
                    tgroup.y &= ~(1u << bit);
                    const float4 sp4 = __ldg(&sc.inst[tgroup.x + bit].sphere);
                    const V3 oc = v3(sp4.x - o.x, sp4.y - o.y, sp4.z - o.z);
                    const float b = xdot(oc, d), c2 = xdot(oc, oc), r2 = sp4.w * sp4.w;
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
                    park[(cur * 3) * B] = make_float4(o.x, o.y, o.z, __uint_as_float(octinv4));
                    park[(cur * 3 + 1) * B] = make_float4(d.x, d.y, d.z, 0.f);
                    park[(cur * 3 + 2) * B] = make_float4(inv.x, inv.y, inv.z, 0.f);
                    const V3 wo = o, wd = d;
                    o = xform_point(r0, r1, r2, wo); d = xform_vector(r0, r1, r2, wd);
                    if (STATS) cI++;
                    setup = true;
                }
            }
            __syncwarp();
            if (setup) {
                inv = v3(rcp_dir(d.x), rcp_dir(d.y), rcp_dir(d.z));
                octinv4 = octant_inv4(inv);
                ngroup = make_uint2(0u, 0x80000000u); tgroup = make_uint2(0u, 0u);
            }
        }

        // ---------------------------------------------------------------- phase N: one node per lane ----
        if (live && (ngroup.y & 0xff000000u))
        {
            if (tgroup.y) { push(tgroup); tgroup = make_uint2(0u, 0u); }
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y & 0xff000000u) push(ngroup);
            const uint32_t slot = (bit - 24u) ^ (octinv4 & 0xffu);
            const uint32_t child = ngroup.x + __popc(ngroup.y & ((1u << slot) - 1u) & 0xffu);
            intersect_children(nodes, child, o, inv, octinv4, ANY_HIT ? tmax : fminf(tmax, hitT), k47, ngroup, tgroup);
            if (STATS) cN++;
        }

        // ---------------------------------------------------------------- phase T: triangles (rounds until too few lanes) ----
        bool swappedT = false;                                                    // at most one swap per lane and iteration in here
        while (true)
        {
            bool wantT = live && instDepth >= 0 && tgroup.y != 0u && !(ANY_HIT && occluded);
            const bool nodeLeft = live && (ngroup.y & 0xff000000u) != 0u;
            const bool parkedT = !swappedT && !wantT && !nodeLeft && bWant == DW_T;   // only the parked ray has triangles and the active one is blocked
            const uint32_t mT = __ballot_sync(NX_FULL, wantT || parkedT);
            if (!mT) break;
            const uint32_t mN2 = __ballot_sync(NX_FULL, nodeLeft);
            if (__popc(mT) < tune.triLanes && mN2 != 0u) break;
            if (STATS) { wRT++; wLT += __popc(mT); }
            if (parkedT) { swap(); swappedT = true; wantT = live && instDepth >= 0 && tgroup.y != 0u && !(ANY_HIT && occluded); }
            if (wantT) test_triangle();
        }
    }
    if (STATS) {
        atomicAdd(&stats->nodes, cN); atomicAdd(&stats->tris, cT); atomicAdd(&stats->insts, cI); atomicAdd(&stats->rays, cR); atomicAdd(&stats->sphereCulled, cS);
        if (lane_id() == 0) {
            atomicAdd(&stats->iters, wIt); atomicAdd(&stats->lanesN, wLN); atomicAdd(&stats->roundsT, wRT); atomicAdd(&stats->lanesT, wLT);
            atomicAdd(&stats->roundsX, wRX); atomicAdd(&stats->lanesX, wLX); atomicAdd(&stats->roundsN, wIt);
        }
        atomicAdd(&stats->lanesF, wSw);   // duo loop: number of swaps (all lanes)
    }
}
