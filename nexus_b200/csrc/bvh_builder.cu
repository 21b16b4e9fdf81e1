// H-PLOC BVH2 builder + BVH2 -> CWBVH8 collapse/compression for sm_100a.
//
// Replaces NXB::BuildBVH2 / BuildBVH8 (vendor/NexusBVH/NexusBVH/src/BVHBuilder.cpp:115-267) and its kernels
// (src/Cuda/Setup.cu, BinaryBuilder.cu, WideConverter.cu, Eval.cu).  The produced trees are the reference's trees bit for
// bit (same Morton keys, same PLOC merges, same collapse decisions, same quantisation); only the node numbering, which
// in the reference depends on the GPU schedule, is ours (BVH8 nodes come out level by level).
//
// Design deltas against the reference, all HBM/latency motivated:
//   * primitives are staged through shared memory with 16-byte loads instead of nine 4-byte loads at a 36-byte stride;
//   * scene bounds use order-preserving integer atomics (one RED per block and component) instead of CAS loops per warp;
//   * PLOC nearest-neighbour search sends each candidate once (shfl_up) instead of a three-shuffle round trip;
//   * the collapse is a persistent cooperative kernel that walks the BVH8 level by level: leaf-ness and the primitive id
//     of a BVH2 child are read off its index (< n), allocation is one atomic per warp and counter, and there is no
//     spin-waiting work queue.
#include "nx_common.cuh"
#include "radix_sort.cuh"
#ifdef NX_WITH_CUB   // measurement builds only (scripts/build_variant.sh cub -DNX_WITH_CUB): the product library contains no library kernel
#include <cub/device/device_radix_sort.cuh>
#endif
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr uint32_t kSearchRadius = 8;     // H-PLOC search radius (BinaryBuilder.cu:9)
constexpr uint32_t kMergeThreshold = 16;  // clusters kept per LBVH range (BinaryBuilder.cu:10)
constexpr int kSetupBlock = 256;
// One warp per CTA: most leaf threads stop after a step or two and a warp lives as long as its last climbing lane, so resources are
// best released warp by warp - a CTA holds its slot until its last warp is done.  Measured at 10 M triangles: 32 / 64 / 128 / 256 threads
// per CTA = 1.75 / 1.80 / 1.79 (round-2 start) / 2.21 ms.
#ifndef NX_PLOC_BLOCK
#define NX_PLOC_BLOCK 32
#endif
#ifndef NX_DP_BLOCK
#define NX_DP_BLOCK 128
#endif
constexpr int kPlocBlock = NX_PLOC_BLOCK;   // one-phase kernel and the global phase of the two-phase builder
constexpr int kDpBlock = NX_DP_BLOCK;
#ifndef NX_PLOC_CHUNK
#define NX_PLOC_CHUNK 512
#endif
#ifndef NX_PLOC_MINB
#define NX_PLOC_MINB 3
#endif
constexpr uint32_t kPlocChunk = NX_PLOC_CHUNK;   // sorted leaves per CTA in the block-local phase of hploc2_kernel
constexpr int kCollapseBlock = 256;

struct SceneKeys { uint32_t lo[3], hi[3]; };  // order-preserving uint encoding of the scene AABB

// ------------------------------------------------------------------------------------------ leaf bounds ----
// One leaf node per primitive: {bounds, left = INVALID, right = i}; scene bounds by min/max reduction.
template <int FLOATS>  // 9 = triangle, 6 = AABB
__global__ void __launch_bounds__(kSetupBlock) leaf_bounds_kernel(const float* __restrict__ prims, uint32_t n, float4* __restrict__ nodes, SceneKeys* scene)
{
    __shared__ float tile[kSetupBlock * FLOATS];
    __shared__ float red[6][kSetupBlock / 32];
    Box acc; acc.lo = v3(3.402823466e38f, 3.402823466e38f, 3.402823466e38f); acc.hi = v3(-3.402823466e38f, -3.402823466e38f, -3.402823466e38f);

    const uint32_t tiles = (n + kSetupBlock - 1) / kSetupBlock;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x)
    {
        const uint32_t first = t * kSetupBlock;
        const uint32_t count = min((uint32_t)kSetupBlock, n - first);
        const size_t base = (size_t)first * FLOATS;             // kSetupBlock*FLOATS*4 is a multiple of 16 bytes
        const uint32_t floats = count * FLOATS;
        const uint32_t vec = ((reinterpret_cast<uintptr_t>(prims) & 15u) == 0) ? floats / 4 : 0;   // 16-byte path needs an aligned base
        const float4* src4 = reinterpret_cast<const float4*>(prims + base);
        float4* dst4 = reinterpret_cast<float4*>(tile);
        for (uint32_t i = threadIdx.x; i < vec; i += kSetupBlock) dst4[i] = __ldg(src4 + i);
        for (uint32_t i = vec * 4 + threadIdx.x; i < floats; i += kSetupBlock) tile[i] = __ldg(prims + base + i);
        __syncthreads();
        if (threadIdx.x < count)
        {
            const float* p = tile + threadIdx.x * FLOATS;
            Box b;
            if (FLOATS == 9) {
                V3 a = v3(p[0], p[1], p[2]), c = v3(p[3], p[4], p[5]), d = v3(p[6], p[7], p[8]);
                b.lo = vmin3(a, vmin3(c, d)); b.hi = vmax3(a, vmax3(c, d));
            } else {
                b.lo = v3(p[0], p[1], p[2]); b.hi = v3(p[3], p[4], p[5]);
            }
            const uint32_t i = first + threadIdx.x;
            nodes[2 * (size_t)i] = make_float4(b.lo.x, b.lo.y, b.lo.z, b.hi.x);
            nodes[2 * (size_t)i + 1] = make_float4(b.hi.y, b.hi.z, __uint_as_float(NX_INVALID), __uint_as_float(i));
            box_grow(acc, b);
        }
        __syncthreads();
    }
    float v[6] = {acc.lo.x, acc.lo.y, acc.lo.z, acc.hi.x, acc.hi.y, acc.hi.z};
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { float w = __shfl_xor_sync(NX_FULL, v[k], o); v[k] = k < 3 ? fminf(v[k], w) : fmaxf(v[k], w); }
        if (lane_id() == 0) red[k][threadIdx.x / 32] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 6)
    {
        const int k = threadIdx.x;
        float r = red[k][0];
        for (int w = 1; w < kSetupBlock / 32; w++) r = k < 3 ? fminf(r, red[k][w]) : fmaxf(r, red[k][w]);
        if (k < 3) atomicMin(&scene->lo[k], f2ord(r)); else atomicMax(&scene->hi[k - 3], f2ord(r));
    }
}

// ------------------------------------------------------------------------------------------ Morton keys ----
// Bit-identical to the reference's fast-math Morton normalisation: add/mul/sub/div.approx with flush-to-zero, truncating
// conversion (checked against the PTX nvcc emits for Setup.cu:43-59 with --use_fast_math).
__device__ __forceinline__ uint32_t morton_axis(float lo, float hi, float smin, float smax, float scale)
{
    float c, num, den, q, m; uint32_t r;
    asm("add.ftz.f32 %0, %1, %2;" : "=f"(c) : "f"(lo), "f"(hi));
    asm("mul.ftz.f32 %0, %1, 0f3F000000;" : "=f"(c) : "f"(c));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(num) : "f"(c), "f"(smin));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(den) : "f"(smax), "f"(smin));
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(q) : "f"(num), "f"(den));
    asm("mul.ftz.f32 %0, %1, %2;" : "=f"(m) : "f"(q), "f"(scale));
    asm("cvt.rzi.ftz.u32.f32 %0, %1;" : "=r"(r) : "f"(m));
    return r;
}
__device__ __forceinline__ uint32_t spread10(uint32_t x)
{
    x = (x | (x << 16)) & 0x030000ffu; x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;  x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint64_t spread21(uint64_t x)
{
    x = (x | (x << 32)) & 0x001f00000000ffffull; x = (x | (x << 16)) & 0x001f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;  x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// Also accumulates the digit histograms of every pass of the radix sort that follows (radix_sort.cuh), so the sort never makes
// a histogram pass over the keys: shared-memory counters per CTA, one global atomic per non-empty bin at the end.
template <typename KeyT>
__global__ void __launch_bounds__(kSetupBlock) morton_kernel(const float4* __restrict__ nodes, uint32_t n, const SceneKeys* __restrict__ scene,
                                                             KeyT* __restrict__ keys, uint32_t* __restrict__ order, float* __restrict__ sceneOut, uint32_t* __restrict__ hist)
{
    constexpr int PASSES = SortShape<KeyT>::kPasses;
    __shared__ uint32_t sHist[PASSES * 256];
    if (hist) { for (uint32_t i = threadIdx.x; i < PASSES * 256; i += blockDim.x) sHist[i] = 0u; __syncthreads(); }
    const float sx0 = ord2f(scene->lo[0]), sy0 = ord2f(scene->lo[1]), sz0 = ord2f(scene->lo[2]);
    const float sx1 = ord2f(scene->hi[0]), sy1 = ord2f(scene->hi[1]), sz1 = ord2f(scene->hi[2]);
    if (blockIdx.x == 0 && threadIdx.x == 0) { sceneOut[0] = sx0; sceneOut[1] = sy0; sceneOut[2] = sz0; sceneOut[3] = sx1; sceneOut[4] = sy1; sceneOut[5] = sz1; }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 a = __ldg(nodes + 2 * (size_t)i), b = __ldg(nodes + 2 * (size_t)i + 1);
        if (sizeof(KeyT) == 4) {
            const float s = 1023.0f;
            uint32_t x = morton_axis(a.x, a.w, sx0, sx1, s), y = morton_axis(a.y, b.x, sy0, sy1, s), z = morton_axis(a.z, b.y, sz0, sz1, s);
            keys[i] = (KeyT)(spread10(x) | (spread10(y) << 1) | (spread10(z) << 2));
        } else {
            const float s = 2097151.0f;
            uint32_t x = morton_axis(a.x, a.w, sx0, sx1, s), y = morton_axis(a.y, b.x, sy0, sy1, s), z = morton_axis(a.z, b.y, sz0, sz1, s);
            keys[i] = (KeyT)(spread21(x) | (spread21(y) << 1) | (spread21(z) << 2));
        }
        order[i] = i;
        if (hist) {
            const KeyT k = keys[i];
#pragma unroll
            for (int p = 0; p < PASSES; p++) atomicAdd(&sHist[p * 256 + sort_digit<KeyT>(k, p)], 1u);
        }
    }
    if (hist) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < PASSES * 256; i += blockDim.x) { const uint32_t c = sHist[i]; if (c) atomicAdd(hist + i, c); }
    }
}

// ----------------------------------------------------------------------------------------------- H-PLOC ----
struct PlocArgs {
    float4* nodes;        // BVH2 nodes, 2 x float4 each: (lo.xyz, hi.x) (hi.y, hi.z, left, right)
    uint32_t* cluster;    // cluster ids, sorted-primitive order; rewritten in place as ranges merge
    uint32_t* parent;     // LBVH "other boundary" slots, 0xffffffff until the first child arrives
    uint32_t* allocated;  // number of BVH2 nodes handed out so far (starts at n)
    uint32_t n;
    uint32_t* seedLo; uint32_t* seedHi;   // two-phase builder: ranges left at the chunk borders (two per chunk, INVALID when unused)
    uint8_t* height;      // optional (global merge path): height of every node above the leaves, saturated at 255 - lets the SAH-optimal
                          // collapse evaluate its C(n, i) tables level by level instead of climbing with atomics (dp_wave_kernel)
};

// Highest differing bit between neighbouring keys, with the position as tie-break (Apetrei 2014; the 32-bit flavour
// concatenates key and position, the 64-bit one falls back to the position only on equal keys - BinaryBuilder.cu:16-28).
__device__ __forceinline__ uint64_t key_delta(const uint32_t* keys, uint32_t a, uint32_t b)
{
    return (((uint64_t)__ldg(keys + a) << 32) | a) ^ (((uint64_t)__ldg(keys + b) << 32) | b);
}
__device__ __forceinline__ uint64_t key_delta(const uint64_t* keys, uint32_t a, uint32_t b)
{
    const uint64_t d = __ldg(keys + a) ^ __ldg(keys + b);
    return d ? d : (uint64_t)(a ^ b);
}

__device__ __forceinline__ Box shfl_box(const Box& b, uint32_t src)
{
    Box r;
    r.lo.x = __shfl_sync(NX_FULL, b.lo.x, src); r.lo.y = __shfl_sync(NX_FULL, b.lo.y, src); r.lo.z = __shfl_sync(NX_FULL, b.lo.z, src);
    r.hi.x = __shfl_sync(NX_FULL, b.hi.x, src); r.hi.y = __shfl_sync(NX_FULL, b.hi.y, src); r.hi.z = __shfl_sync(NX_FULL, b.hi.z, src);
    return r;
}

// Warp-cooperative PLOC over the clusters of one LBVH range [lo, hiEnd) split at mid.  Every lane holds at most one cluster.
__device__ void ploc_merge_range(const PlocArgs& a, uint32_t lo, uint32_t mid, uint32_t hiEnd, bool isRoot)
{
    const uint32_t lane = lane_id();
    uint32_t id = NX_INVALID;

    // gather: up to kMergeThreshold ids from the left range into lanes [0..), then from the right range behind them
    const uint32_t takeL = min(mid - lo, kMergeThreshold);
    if (lane < takeL) id = __ldcg(a.cluster + lo + lane);
    const uint32_t numL = __popc(__ballot_sync(NX_FULL, lane < takeL && id != NX_INVALID));
    const uint32_t takeR = min(hiEnd - mid, kMergeThreshold);
    const uint32_t offR = lane - numL;  // wraps for lanes below numL
    if (offR < takeR) id = __ldcg(a.cluster + mid + offR);
    const uint32_t numR = __popc(__ballot_sync(NX_FULL, offR < takeR && id != NX_INVALID));
    const uint32_t loaded = numL + numR;
    uint32_t num = loaded;

    Box box; box.lo = v3(0.f, 0.f, 0.f); box.hi = v3(0.f, 0.f, 0.f);
    if (lane < num) {
        const float4 p = ld_cg4(a.nodes + 2 * (size_t)id), q = ld_cg4(a.nodes + 2 * (size_t)id + 1);
        box.lo = v3(p.x, p.y, p.z); box.hi = v3(p.w, q.x, q.y);
    }

    const uint32_t keep = isRoot ? 1u : kMergeThreshold;
    while (num > keep)
    {
        // nearest neighbour within +-kSearchRadius by merged half-area, compared on the float bits; ties keep the
        // candidate seen first (+1, -1, +2, -2, ...)
        uint32_t bestArea = NX_INVALID, bestLane = NX_INVALID;
#pragma unroll
        for (uint32_t r = 1; r <= kSearchRadius; r++)
        {
            Box other;
            other.lo.x = __shfl_down_sync(NX_FULL, box.lo.x, r); other.lo.y = __shfl_down_sync(NX_FULL, box.lo.y, r); other.lo.z = __shfl_down_sync(NX_FULL, box.lo.z, r);
            other.hi.x = __shfl_down_sync(NX_FULL, box.hi.x, r); other.hi.y = __shfl_down_sync(NX_FULL, box.hi.y, r); other.hi.z = __shfl_down_sync(NX_FULL, box.hi.z, r);
            uint32_t fwd = NX_INVALID;
            if (lane + r < num) {
                box_grow(other, box);
                fwd = __float_as_uint(half_area_ref(other));
                if (fwd < bestArea) { bestArea = fwd; bestLane = lane + r; }
            }
            const uint32_t bwd = __shfl_up_sync(NX_FULL, fwd, r);   // the same pair seen from the other side
            if (lane >= r && bwd < bestArea) { bestArea = bwd; bestLane = lane - r; }
        }

        const bool alive = lane < num;
        const uint32_t theirs = __shfl_sync(NX_FULL, bestLane, bestLane);
        const bool mutual = alive && theirs == lane;
        const bool owner = mutual && lane < bestLane;          // the lower lane of a mutual pair creates the node
        const uint32_t ownerMask = __ballot_sync(NX_FULL, owner);
        const uint32_t created = __popc(ownerMask);
        uint32_t base = 0;
        if (lane == 0 && created) base = atomicAdd(a.allocated, created);
        base = __shfl_sync(NX_FULL, base, 0);

        const uint32_t partnerId = __shfl_sync(NX_FULL, id, bestLane);
        const Box partnerBox = shfl_box(box, bestLane);
        if (owner) {
            box_grow(box, partnerBox);
            const uint32_t node = base + __popc(ownerMask & ((1u << lane) - 1u));
            st_cg4(a.nodes + 2 * (size_t)node, make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x));
            st_cg4(a.nodes + 2 * (size_t)node + 1, make_float4(box.hi.y, box.hi.z, __uint_as_float(id), __uint_as_float(partnerId)));
            id = node;
        }
        // compact: survivors are the pair owners and every cluster without a mutual partner, order preserved
        const uint32_t keepMask = __ballot_sync(NX_FULL, alive && (owner || !mutual));
        const uint32_t src = __fns(keepMask, 0, lane + 1);
        const uint32_t movedId = __shfl_sync(NX_FULL, id, src);
        box = shfl_box(box, src);
        num -= created;
        id = lane < num ? movedId : NX_INVALID;
    }

    if (lane < loaded) __stcg(a.cluster + lo + lane, id);
    __threadfence();
}

template <typename KeyT>
__global__ void __launch_bounds__(kPlocBlock) hploc_kernel(PlocArgs a, const KeyT* __restrict__ keys)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lo = i, hi = i, mid = 0;
    bool climbing = i < a.n;

    while (__ballot_sync(NX_FULL, climbing))
    {
        if (climbing)
        {
            // Apetrei's bottom-up step: the parent is on the side with the smaller key delta
            const bool parentAtHi = lo == 0 || (hi != a.n - 1 && key_delta(keys, hi, hi + 1) < key_delta(keys, lo - 1, lo));
            uint32_t other;
            if (parentAtHi) { other = atomicExch(a.parent + hi, lo); if (other != NX_INVALID) { mid = hi + 1; hi = other; } }
            else            { other = atomicExch(a.parent + lo - 1, hi); if (other != NX_INVALID) { mid = lo; lo = other; } }
            if (other == NX_INVALID) climbing = false;   // first to arrive: the sibling's thread continues
            else __threadfence();                        // second to arrive: acquire the sibling's cluster list
        }
        const uint32_t size = hi - lo + 1;
        const bool isRoot = climbing && size == a.n;
        uint32_t todo = __ballot_sync(NX_FULL, (climbing && size > kMergeThreshold) || isRoot);
        while (todo)
        {
            const uint32_t src = __ffs(todo) - 1;
            ploc_merge_range(a, __shfl_sync(NX_FULL, lo, src), __shfl_sync(NX_FULL, mid, src), __shfl_sync(NX_FULL, hi, src) + 1,
                             __shfl_sync(NX_FULL, (int)isRoot, src) != 0);
            todo &= todo - 1;
        }
    }
}

// ------------------------------------------------------------------------------- H-PLOC, two-phase version ----
// Same algorithm and the same merges as hploc_kernel above (so the same tree, bit for bit), organised for the memory system:
//
//   Phase A (block local).  A CTA owns a chunk of C consecutive sorted leaves.  Every LBVH range that lies inside the chunk
//   is climbed and PLOC-merged entirely in shared memory: leaf boxes, the nodes created, the cluster lists, the parent slots
//   and the Morton keys are staged there, the hand-off between sibling ranges is a shared-memory atomicExch + block-scope
//   fence, nodes get provisional chunk-local ids.  ncu on the one-phase kernel (profiles/r01_ncu_hploc_build10m.md): 38 % of
//   the stall samples wait for global round trips (atomicExch, cluster ids, boxes, the node-index atomic inside every merge
//   iteration) and 25 % for device-scope fences, at 41 % issue utilisation; more than nine merges in ten never leave a chunk.
//   Flush.  One atomicAdd per CTA reserves the global node indices of everything the chunk created; nodes, remapped cluster
//   ids and the half-arrived parent slots go to global memory with coalesced stores.
//   Phase B (global).  Only ranges that cross a chunk border continue with the global protocol of the one-phase kernel:
//   threads whose parent slot lies on the chunk border, and the holders of inside slots whose sibling (a range reaching in
//   from a neighbouring chunk) has already arrived in global memory.
//
// The merge itself keeps each cluster's (box, id) in a per-warp shared-memory table: the nearest-neighbour search reads
// neighbour boxes with two LDS instead of six SHFL, and compaction is a scatter by rank instead of __fns + seven SHFL.
constexpr uint32_t kSlotDone = 0xfffffffeu;

struct WarpTable {
    float4 a[40], b[40];   // per warp: a = {lo.xyz, hi.x}, b = {hi.y, hi.z, id, -}; 8 entries of slack for lane + radius
    float4 pa[32], pb[32]; // global path: the nodes a merge has created so far, with provisional ids (kTempId | k), written out at its end
    uint32_t ph[32];       // ... and their heights
};
constexpr uint32_t kTempId = 0x80000000u;   // node ids stay below 2^31 (n <= 2^30)

template <uint32_t C> struct ChunkMem {
    float4 node[2 * C][2];     // [0, C): leaves of the chunk in sorted order; [C, 2C): nodes created here (provisional ids)
    uint32_t cluster[C];       // cluster ids (chunk-local node ids) by sorted position
    uint32_t leafId[C];        // global node id (= primitive id) of the chunk's leaves
    uint32_t parent[C];        // LBVH parent slots inside the chunk
    unsigned long long key[C + 2];   // Morton keys of positions chunkStart - 1 .. chunkStart + C
    WarpTable table[C / 32];
    uint32_t created, base;
};

template <typename KeyT>
__device__ __forceinline__ uint64_t key_delta_vals(KeyT ka, KeyT kb, uint32_t a, uint32_t b)
{
    if (sizeof(KeyT) == 4) return (((uint64_t)ka << 32) | a) ^ (((uint64_t)kb << 32) | b);
    const uint64_t d = (uint64_t)ka ^ (uint64_t)kb;
    return d ? d : (uint64_t)(a ^ b);
}

// PLOC merge of the clusters of one LBVH range.  LOCAL: positions, ids and nodes are the chunk's (shared memory);
// otherwise global memory, exactly like ploc_merge_range.
template <bool LOCAL, uint32_t C>
__device__ __forceinline__ void ploc_merge2(const PlocArgs& a, ChunkMem<C>& M, WarpTable& tb, uint32_t lo, uint32_t mid, uint32_t hiEnd, bool isRoot)
{
    const uint32_t lane = lane_id();
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    uint32_t id = NX_INVALID;
    const uint32_t takeL = min(mid - lo, kMergeThreshold);
    if (lane < takeL) id = LOCAL ? M.cluster[lo + lane] : __ldcg(a.cluster + lo + lane);
    const uint32_t numL = __popc(__ballot_sync(NX_FULL, lane < takeL && id != NX_INVALID));
    const uint32_t takeR = min(hiEnd - mid, kMergeThreshold);
    const uint32_t offR = lane - numL;  // wraps for lanes below numL
    if (offR < takeR) id = LOCAL ? M.cluster[mid + offR] : __ldcg(a.cluster + mid + offR);
    const uint32_t numR = __popc(__ballot_sync(NX_FULL, offR < takeR && id != NX_INVALID));
    const uint32_t loaded = numL + numR;
    uint32_t num = loaded;

    Box box; box.lo = v3(0.f, 0.f, 0.f); box.hi = v3(0.f, 0.f, 0.f);
    uint32_t h = 0;                                       // height of the lane's cluster (global path with a.height only)
    if (lane < num) {
        float4 p, q;
        if (LOCAL) { p = M.node[id][0]; q = M.node[id][1]; } else { p = ld_cg4(a.nodes + 2 * (size_t)id); q = ld_cg4(a.nodes + 2 * (size_t)id + 1); }
        if (!LOCAL && a.height) h = __ldcg(a.height + id);
        box.lo = v3(p.x, p.y, p.z); box.hi = v3(p.w, q.x, q.y);
    }
    __syncwarp();
    tb.a[lane] = make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x);
    tb.b[lane] = make_float4(box.hi.y, box.hi.z, __uint_as_float(id), __uint_as_float(h));
    __syncwarp();

    const uint32_t keep = isRoot ? 1u : kMergeThreshold;
    uint32_t pending = 0;   // global path: nodes created by this merge so far.  Their indices are reserved with ONE atomic at the end
                            // (the one-phase kernel pays a global atomic round trip inside every iteration); the root, created by the
                            // last iteration of the last merge, still gets the last index 2n - 2.
    while (num > keep)
    {
        // nearest neighbour within +-kSearchRadius by merged half-area, compared on the float bits; ties keep the candidate seen
        // first (+1, -1, +2, -2, ...), BinaryBuilder.cu:127-167
        uint32_t bestArea = NX_INVALID, bestLane = NX_INVALID;
#pragma unroll
        for (uint32_t r = 1; r <= kSearchRadius; r++)
        {
            const float4 o0 = tb.a[lane + r];
            const float2 o1 = *reinterpret_cast<const float2*>(&tb.b[lane + r]);
            uint32_t fwd = NX_INVALID;
            if (lane + r < num) {
                Box other; other.lo = v3(o0.x, o0.y, o0.z); other.hi = v3(o0.w, o1.x, o1.y);
                box_grow(other, box);
                fwd = __float_as_uint(half_area_ref(other));
                if (fwd < bestArea) { bestArea = fwd; bestLane = lane + r; }
            }
            const uint32_t bwd = __shfl_up_sync(NX_FULL, fwd, r);   // the same pair seen from the other side
            if (lane >= r && bwd < bestArea) { bestArea = bwd; bestLane = lane - r; }
        }
        const bool alive = lane < num;
        const uint32_t theirs = __shfl_sync(NX_FULL, bestLane, bestLane);
        const bool mutual = alive && theirs == lane;
        const bool owner = mutual && lane < bestLane;          // the lower lane of a mutual pair creates the node
        const uint32_t ownerMask = __ballot_sync(NX_FULL, owner);
        const uint32_t created = __popc(ownerMask);
        uint32_t base = pending;
        if (LOCAL) {
            if (lane == 0 && created) base = atomicAdd(&M.created, created);
            base = __shfl_sync(NX_FULL, base, 0);
        }
        pending += created;
        if (owner) {
            const float4 p0 = tb.a[bestLane], p1 = tb.b[bestLane];
            Box pb; pb.lo = v3(p0.x, p0.y, p0.z); pb.hi = v3(p0.w, p1.x, p1.y);
            box_grow(box, pb);
            const uint32_t k = base + __popc(ownerMask & lane_lt);
            const float4 n0 = make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x), n1 = make_float4(box.hi.y, box.hi.z, __uint_as_float(id), p1.z);
            if (LOCAL) { M.node[C + k][0] = n0; M.node[C + k][1] = n1; id = C + k; }
            else { h = min(255u, 1u + max(h, __float_as_uint(p1.w))); tb.pa[k] = n0; tb.pb[k] = n1; tb.ph[k] = h; id = kTempId | k; }
        }
        // compact: survivors are the pair owners and every cluster without a mutual partner, order preserved
        const bool stays = alive && (owner || !mutual);
        const uint32_t keepMask = __ballot_sync(NX_FULL, stays);
        __syncwarp();                                         // every read of the table is done
        if (stays) {
            const uint32_t rank = __popc(keepMask & lane_lt);
            tb.a[rank] = make_float4(box.lo.x, box.lo.y, box.lo.z, box.hi.x);
            tb.b[rank] = make_float4(box.hi.y, box.hi.z, __uint_as_float(id), __uint_as_float(h));
        }
        __syncwarp();
        num -= created;
        id = NX_INVALID;
        if (lane < num) {
            const float4 p = tb.a[lane], q = tb.b[lane];
            box.lo = v3(p.x, p.y, p.z); box.hi = v3(p.w, q.x, q.y); id = __float_as_uint(q.z); h = __float_as_uint(q.w);
        }
    }
    if (!LOCAL) {
        // reserve the indices, write the nodes with their children's provisional ids resolved, resolve the surviving cluster ids
        uint32_t base = 0;
        if (lane == 0 && pending) base = atomicAdd(a.allocated, pending);
        base = __shfl_sync(NX_FULL, base, 0);
        __syncwarp();
        if (lane < pending) {
            const float4 n0 = tb.pa[lane]; float4 n1 = tb.pb[lane];
            uint32_t l = __float_as_uint(n1.z), r = __float_as_uint(n1.w);
            if (l & kTempId) l = base + (l & ~kTempId);
            if (r & kTempId) r = base + (r & ~kTempId);
            n1.z = __uint_as_float(l); n1.w = __uint_as_float(r);
            st_cg4(a.nodes + 2 * (size_t)(base + lane), n0); st_cg4(a.nodes + 2 * (size_t)(base + lane) + 1, n1);
            if (a.height) a.height[base + lane] = (uint8_t)tb.ph[lane];
        }
        if (id != NX_INVALID && (id & kTempId)) id = base + (id & ~kTempId);
        __syncwarp();
    }
    if (lane < loaded) { if (LOCAL) M.cluster[lo + lane] = id; else __stcg(a.cluster + lo + lane, id); }
    // Global path: no fence here.  The nodes, heights and cluster ids written above are published by the RELEASE exchange the climbing
    // lane performs next (exch_release in hploc_seed_kernel); the reader takes them with ld.cg after it has seen that exchange.
    // __threadfence() would be MEMBAR.SC.GPU + CCTL.IVALL: a sequentially consistent barrier plus an invalidation of the SM's whole L1,
    // once per merge, for every warp on the SM.
    if (LOCAL) __threadfence_block(); else __syncwarp();   // every lane's stores are ordered before the owning lane's release
}

template <typename KeyT, uint32_t C>
__global__ void __launch_bounds__(C, NX_PLOC_MINB) hploc2_kernel(PlocArgs a, const KeyT* __restrict__ keys)
{
    extern __shared__ __align__(16) unsigned char chunk_raw[];
    ChunkMem<C>& M = *reinterpret_cast<ChunkMem<C>*>(chunk_raw);
    WarpTable& tb = M.table[threadIdx.x >> 5];
    const uint32_t t = threadIdx.x, n = a.n;
    const uint32_t chunkStart = blockIdx.x * C, count = min(C, n - chunkStart);

    // ---- stage the chunk: leaf nodes in sorted order, identity cluster list, empty parent slots, keys with one of margin each side
    if (t < count) {
        const uint32_t prim = __ldg(a.cluster + chunkStart + t);
        M.leafId[t] = prim; M.cluster[t] = t; M.parent[t] = NX_INVALID;
        M.node[t][0] = __ldg(a.nodes + 2 * (size_t)prim); M.node[t][1] = __ldg(a.nodes + 2 * (size_t)prim + 1);
    }
    for (uint32_t i = t; i < count + 2; i += C) {
        const uint64_t g = (uint64_t)chunkStart + i;            // key[i] = keys[chunkStart + i - 1]
        M.key[i] = (g >= 1 && g - 1 < n) ? (unsigned long long)__ldg(keys + (g - 1)) : 0ull;
    }
    if (t == 0) M.created = 0;
    __syncthreads();

    // ---- phase A: climb and merge inside the chunk (positions are chunk-local)
    uint32_t lo = t, hi = t, mid = 0;
    bool climbing = t < count, blocked = false, blockedRight = false;
    auto delta = [&](uint32_t x, uint32_t y) {      // local positions x, y = x + 1 (may be -1 / count: the margins)
        return key_delta_vals<KeyT>((KeyT)M.key[x + 1], (KeyT)M.key[y + 1], chunkStart + x, chunkStart + y);
    };
    while (__ballot_sync(NX_FULL, climbing))
    {
        if (climbing)
        {
            const uint32_t gl = chunkStart + lo, gh = chunkStart + hi;
            const bool parentAtHi = gl == 0 || (gh != n - 1 && delta(hi, hi + 1) < delta(lo - 1, lo));
            const bool inside = parentAtHi ? (hi + 1 < count) : (lo > 0);      // the sibling range starts / ends inside this chunk
            if (!inside) { blocked = true; blockedRight = parentAtHi; climbing = false; }
            else {
                uint32_t other;
                if (parentAtHi) { other = atomicExch(&M.parent[hi], lo); if (other != NX_INVALID) { M.parent[hi] = kSlotDone; mid = hi + 1; hi = other; } }
                else            { other = atomicExch(&M.parent[lo - 1], hi); if (other != NX_INVALID) { M.parent[lo - 1] = kSlotDone; mid = lo; lo = other; } }
                if (other == NX_INVALID) climbing = false;   // first to arrive: the sibling's thread continues
                else __threadfence_block();                  // second to arrive: acquire the sibling's cluster list
            }
        }
        const uint32_t size = hi - lo + 1;
        const bool isRoot = climbing && size == n;
        uint32_t todo = __ballot_sync(NX_FULL, (climbing && size > kMergeThreshold) || isRoot);
        while (todo)
        {
            const uint32_t src = __ffs(todo) - 1;
            ploc_merge2<true, C>(a, M, tb, __shfl_sync(NX_FULL, lo, src), __shfl_sync(NX_FULL, mid, src), __shfl_sync(NX_FULL, hi, src) + 1,
                                 __shfl_sync(NX_FULL, (int)isRoot, src) != 0);
            todo &= todo - 1;
        }
    }
    __syncthreads();

    // ---- flush: reserve global node ids, write the nodes created here, the cluster lists (global ids) of the chunk
    const uint32_t made = M.created;
    if (t == 0) M.base = made ? atomicAdd(a.allocated, made) : 0u;
    __syncthreads();
    const uint32_t base = M.base;
    auto remap = [&](uint32_t id) { return id == NX_INVALID ? NX_INVALID : (id < C ? M.leafId[id] : base + (id - C)); };
    for (uint32_t k = t; k < made; k += C) {
        const float4 n0 = M.node[C + k][0]; float4 n1 = M.node[C + k][1];
        n1.z = __uint_as_float(remap(__float_as_uint(n1.z))); n1.w = __uint_as_float(remap(__float_as_uint(n1.w)));
        st_cg4(a.nodes + 2 * (size_t)(base + k), n0); st_cg4(a.nodes + 2 * (size_t)(base + k) + 1, n1);
    }
    if (t < count) __stcg(a.cluster + chunkStart + t, remap(M.cluster[t]));
    __threadfence();
    __syncthreads();

    // ---- hand-over to the global phase (hploc_seed_kernel, launched behind this kernel): the parent slots inside the chunk where only
    // one side arrived become ordinary first arrivals in global memory (nothing else touches them before the next kernel), and the
    // at most two ranges that stopped at the chunk's borders become the seeds the global phase climbs on from.
    if (t + 1 < count) {
        const uint32_t v = M.parent[t];
        a.parent[chunkStart + t] = (v != NX_INVALID && v != kSlotDone) ? chunkStart + v : NX_INVALID;
    }
    if (t < 2) { a.seedLo[2 * blockIdx.x + t] = NX_INVALID; a.seedHi[2 * blockIdx.x + t] = NX_INVALID; }
    __syncthreads();
    if (blocked) { const uint32_t k = blockedRight ? 1u : 0u;   // at most one range stops at each border
        a.seedLo[2 * blockIdx.x + k] = chunkStart + lo; a.seedHi[2 * blockIdx.x + k] = chunkStart + hi; }
}

// atomicExch with release semantics at GPU scope: everything this thread wrote before it is visible to whoever observes the exchanged
// value (MEMBAR.ALL.GPU + ATOMG; no L1 invalidation - nothing here is read through L1 after an acquire).
__device__ __forceinline__ uint32_t exch_release(uint32_t* p, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.release.gpu.global.exch.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

// Global phase of the two-phase builder: the one-phase protocol (hploc_kernel) started from the ranges the chunks could not finish.
#ifndef NX_SEED_MINB
#define NX_SEED_MINB 1
#endif
template <typename KeyT>
__global__ void __launch_bounds__(kPlocBlock, NX_SEED_MINB) hploc_seed_kernel(PlocArgs a, const KeyT* __restrict__ keys, uint32_t nSeeds)
{
    __shared__ WarpTable tables[kPlocBlock / 32];
    WarpTable& tb = tables[threadIdx.x >> 5];
    ChunkMem<32>* none = nullptr;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, n = a.n;
    // without a seed list every leaf is its own seed: the one-phase protocol with the shared-memory merge table
    uint32_t lo = i < nSeeds ? (a.seedLo ? __ldg(a.seedLo + i) : i) : NX_INVALID, hi = i < nSeeds ? (a.seedHi ? __ldg(a.seedHi + i) : i) : NX_INVALID, mid = 0;
    bool climbing = lo != NX_INVALID;
    bool wrote = a.seedLo != nullptr;      // seeds of the two-phase builder arrive with data written by the previous kernel (visible already; release is harmless)
    while (__ballot_sync(NX_FULL, climbing))
    {
        if (climbing)
        {
            const bool parentAtHi = lo == 0 || (hi != n - 1 && key_delta(keys, hi, hi + 1) < key_delta(keys, lo - 1, lo));
            // a lane whose range was merged in the previous iteration has nodes / cluster ids to publish: its exchange is a release;
            // a lane that only climbs (ranges of up to kMergeThreshold leaves: nothing written yet) uses the plain exchange
            uint32_t* slot = parentAtHi ? a.parent + hi : a.parent + lo - 1;
            const uint32_t mine = parentAtHi ? lo : hi;
            uint32_t other = wrote ? exch_release(slot, mine) : atomicExch(slot, mine);
            if (other != NX_INVALID) { if (parentAtHi) { mid = hi + 1; hi = other; } else { mid = lo; lo = other; } }
            if (other == NX_INVALID) climbing = false;   // first to arrive: the sibling's thread continues (cluster lists are read with ld.cg)
        }
        const uint32_t size = hi - lo + 1;
        const bool isRoot = climbing && size == n;
        const bool merges = (climbing && size > kMergeThreshold) || isRoot;
        wrote = merges;
        uint32_t todo = __ballot_sync(NX_FULL, merges);
        while (todo)
        {
            const uint32_t src = __ffs(todo) - 1;
            ploc_merge2<false, 32>(a, *none, tb, __shfl_sync(NX_FULL, lo, src), __shfl_sync(NX_FULL, mid, src), __shfl_sync(NX_FULL, hi, src) + 1,
                                   __shfl_sync(NX_FULL, (int)isRoot, src) != 0);
            todo &= todo - 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------- collapse ----
__device__ __forceinline__ Box load_box2(const float4* n2, uint32_t i)
{
    const float4 p = __ldg(n2 + 2 * (size_t)i), q = __ldg(n2 + 2 * (size_t)i + 1);
    Box b; b.lo = v3(p.x, p.y, p.z); b.hi = v3(p.w, q.x, q.y);
    return b;
}

struct CollapseArgs {
    const float4* n2;      // BVH2 nodes
    float4* n8;            // CWBVH8 nodes, 5 x float4 each
    uint32_t* primIdx;     // leaf slot -> primitive id
    uint32_t* bvh2Of;      // work map: BVH8 node index -> BVH2 node it collapses
    uint32_t* counters;    // [0] nodes allocated, [1] leaf slots allocated, [2..4] nodes created per level (rotating, see collapse_kernel)
    uint32_t n;
    // SAH-optimal collapse only (NX_COLLAPSE_SAH_OPTIMAL): C(n, i) decisions of every BVH2 node, dp_eval_kernel
    const unsigned long long* dpDec;
};

// ------------------------------------------------------------------------------- SAH-optimal collapse: the C(n, i) table ----
// The reference's CPU BVH8Builder (Nexus/src/Geometry/BVH/BVH8Builder.cpp:31-145; Ylitie, Karras, Laine 2017 §3): for every
// BVH2 node n and i = 1..7, C(n, i) is the cheapest way to turn n's subtree into at most i BVH8 children, with the
// decision that achieves it - LEAF (at most max_leaf_prims primitives), INTERNAL (n becomes a BVH8 node: 7 roots shared
// between its two children, plus C_NODE * area) or DISTRIBUTE (the roots are split k / i-1-k between the children).  The
// recursion with a memo table becomes one bottom-up pass here: a thread starts at every leaf and climbs, the second
// thread to arrive at a node (atomic counter) evaluates it from its finished children.
// dpDec[n]: byte i = decision (bits 0-1) | left count (bits 2-4) | right count (bits 5-7) for i = 0..6; byte 7 = primitives
// in the subtree, saturated at 255.  dpCost[n * 8 + i] = C(n, i + 1), i = 0..6 (8 floats per node: two 16-byte accesses).
constexpr uint32_t kDpLeaf = 0u, kDpInternal = 1u, kDpDistribute = 2u;
constexpr float kCPrim = 0.3f, kCNode = 1.0f;      // BVH8Builder.h:7-8

struct DpArgs {
    const float4* n2; uint32_t n;
    uint32_t* parent; uint32_t* arrived;           // arrived: one counter per inner node
    float* cost; unsigned long long* dec;
    uint32_t maxLeafPrims;
};

__global__ void dp_parent_kernel(DpArgs a)
{
    for (uint32_t i = a.n + blockIdx.x * blockDim.x + threadIdx.x; i < 2 * a.n - 1; i += gridDim.x * blockDim.x) {
        const float4 q = __ldg(a.n2 + 2 * (size_t)i + 1);
        a.parent[__float_as_uint(q.z)] = i; a.parent[__float_as_uint(q.w)] = i;
        a.arrived[i - a.n] = 0u;
    }
}

// Cost table: 8 floats per node (C(n, 1..7) and a pad), so a child's table is two LDG.128 and a node's two STG.128.
// C(p, 1..7) and the decisions of inner node p from its children's finished tables.
__device__ __forceinline__ void dp_eval_node(const DpArgs& a, uint32_t p)
{
    float4* const cost4 = reinterpret_cast<float4*>(a.cost);
    const float4 q = __ldg(a.n2 + 2 * (size_t)p + 1);
    const uint32_t L = __float_as_uint(q.z), R = __float_as_uint(q.w);
    float cl[7], cr[7], c[7];
    // A BVH2 leaf has no table in memory: C(leaf, i) = area * C_PRIM for every i (CLeaf(node, 1), BVH8Builder.cpp:31-37), one primitive,
    // decision LEAF - computed here from the leaf's box (one 32-byte sector) instead of gathering a 32-byte table and an 8-byte
    // decision word that would first have to be written for every primitive.  Half of all tables are leaves'.
    auto table = [&](uint32_t child, float (&t)[7]) -> uint32_t {
        if (child < a.n) {
            const float v = __fmul_rn(__fmul_rn(half_area_ref(load_box2(a.n2, child)), 1.0f), kCPrim);
#pragma unroll
            for (int k = 0; k < 7; k++) t[k] = v;
            return 1u;
        }
        const float4 t0 = __ldcg(cost4 + 2 * (size_t)child), t1 = __ldcg(cost4 + 2 * (size_t)child + 1);
        t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y; t[6] = t1.z;
        return (uint32_t)(__ldcg(a.dec + child) >> 56);
    };
    const uint32_t trisL = table(L, cl), trisR = table(R, cr);
    const uint32_t tris = min(255u, trisL + trisR);
    const float area = half_area_ref(load_box2(a.n2, p));
    unsigned long long dec = (unsigned long long)tris << 56;
    // CDistribute(node, j) (:39-57): best split of j - 1 roots: k to the left child, j - 1 - k to the right
    auto distribute = [&](int j, uint32_t& l, uint32_t& r) {
        float best = 1.0e30f;
#pragma unroll
        for (int k = 0; k < 7; k++) if (k < j) { const float v = __fadd_rn(cl[k], cr[j - 1 - k]); if (v < best) { best = v; l = (uint32_t)k; r = (uint32_t)(j - 1 - k); } }
        return best;
    };
    {   // i = 0: leaf or internal (:92-113)
        uint32_t l = 0, r = 0;
        const float internal = __fadd_rn(distribute(7, l, r), __fmul_rn(area, kCNode));
        const float leafCost = tris > a.maxLeafPrims ? 1.0e30f : __fmul_rn(__fmul_rn(area, (float)tris), kCPrim);
        if (leafCost < internal) { c[0] = leafCost; dec |= kDpLeaf; }
        else { c[0] = internal; dec |= kDpInternal | (l << 2) | (r << 5); }
    }
#pragma unroll
    for (int i = 1; i < 7; i++) {   // i roots + 1: distribute, or keep the solution with one root fewer (:115-131)
        uint32_t l = 0, r = 0;
        const float d = distribute(i, l, r);
        if (d < c[i - 1]) { c[i] = d; dec |= (unsigned long long)(kDpDistribute | (l << 2) | (r << 5)) << (8 * i); }
        else { c[i] = c[i - 1]; dec |= ((dec >> (8 * (i - 1))) & 0xffull) << (8 * i); }
    }
    __stcg(cost4 + 2 * (size_t)p, make_float4(c[0], c[1], c[2], c[3])); __stcg(cost4 + 2 * (size_t)p + 1, make_float4(c[4], c[5], c[6], 0.f));
    __stcg(a.dec + p, dec);
}
// decision word of any BVH2 node: leaves (ids below n) have none in memory - LEAF for every i, one primitive
__device__ __forceinline__ unsigned long long dp_dec_of(const unsigned long long* __restrict__ dec, uint32_t n, uint32_t node)
{
    return node < n ? (1ull << 56) : __ldg(dec + node);
}

// Bottom-up by climbing: a thread per leaf, the second thread to arrive at a node evaluates it (small inputs, and whenever the
// builder recorded no node heights).
__global__ void __launch_bounds__(kDpBlock) dp_eval_kernel(DpArgs a)
{
    const uint32_t leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= a.n) return;
    uint32_t p = __ldcg(a.parent + leaf);
    while (p != NX_INVALID)
    {
        __threadfence();                                            // release this subtree's tables before announcing it
        if (atomicAdd(a.arrived + (p - a.n), 1u) == 0u) return;     // the sibling subtree is not finished: its thread continues
        // second to arrive: the sibling's tables are read with ld.cg (L2), which is where its release made them visible
        dp_eval_node(a, p);
        p = __ldcg(a.parent + p);
    }
}

// Bottom-up by levels: H-PLOC records the height of every node it creates (PlocArgs::height); the inner nodes are counting-sorted by
// height and one launch per height evaluates all nodes of that height at once - no atomics, no fences, no thread that climbs to the
// root while its block idles.  The climb above takes 2.2 ms at 10 M triangles but 15 ms at 50 M.
struct DpWaveArgs { const uint8_t* height; uint32_t* hist; uint32_t* cursor; uint32_t* order; };

__global__ void dp_leaf_hist_kernel(DpArgs a, DpWaveArgs w)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t i = a.n + blockIdx.x * blockDim.x + threadIdx.x; i < 2 * a.n - 1; i += gridDim.x * blockDim.x) atomicAdd(&sh[w.height[i]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(w.hist + threadIdx.x, sh[threadIdx.x]);
}
__global__ void dp_height_scan_kernel(DpWaveArgs w)     // one block of 256: cursor[h] = number of inner nodes lower than h
{
    __shared__ uint32_t warpSum[8];
    const uint32_t v = w.hist[threadIdx.x];
    uint32_t s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(NX_FULL, s, o); if (lane_id() >= (uint32_t)o) s += t; }
    if (lane_id() == 31) warpSum[threadIdx.x >> 5] = s;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t k = 0; k < (threadIdx.x >> 5); k++) before += warpSum[k];
    w.cursor[threadIdx.x] = before + s - v;
}
__global__ void dp_height_scatter_kernel(DpArgs a, DpWaveArgs w)
{
    uint32_t lane_lt; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lane_lt));
    const uint32_t total = a.n - 1, stride = gridDim.x * blockDim.x;
    for (uint32_t k0 = blockIdx.x * blockDim.x; k0 < total; k0 += stride) {      // whole warps iterate together
        const uint32_t k = k0 + threadIdx.x;
        const bool ok = k < total;
        const uint32_t node = a.n + k, h = ok ? w.height[node] : 0xffffffffu;
        const uint32_t peers = __match_any_sync(NX_FULL, h);                     // neighbouring ids mostly share a height: one atomic per group
        uint32_t base = 0;
        const uint32_t leader = __ffs(peers) - 1;
        if (ok && lane_id() == leader) base = atomicAdd(w.cursor + h, __popc(peers));
        base = __shfl_sync(NX_FULL, base, leader);
        if (ok) w.order[base + __popc(peers & lane_lt)] = node;
    }
}
__global__ void __launch_bounds__(128) dp_wave_kernel(DpArgs a, const uint32_t* __restrict__ order, uint32_t first, uint32_t count)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dp_eval_node(a, __ldg(order + first + k));
}

__device__ __forceinline__ uint32_t ceil_log2_biased(float x)   // biased exponent of the smallest power of two >= x
{
    const uint32_t u = __float_as_uint(x);
    return ((u >> 23) & 0xffu) + ((u & 0x7fffffu) ? 1u : 0u);
}
__device__ __forceinline__ uint32_t quant(float c, float p, float inv, bool up)
{
    float d, m, r; uint32_t q;
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(c), "f"(p));
    asm("mul.ftz.f32 %0, %1, %2;" : "=f"(m) : "f"(d), "f"(inv));
    if (up) asm("cvt.rpi.ftz.f32.f32 %0, %1;" : "=f"(r) : "f"(m)); else asm("cvt.rmi.ftz.f32.f32 %0, %1;" : "=f"(r) : "f"(m));
    asm("cvt.rzi.ftz.u32.f32 %0, %1;" : "=r"(q) : "f"(r));
    return q & 0xffu;
}

// Collapses BVH2 node `root2` into BVH8 node `self`.  OPT = false: the reference GPU converter's rule (open the inner child in
// the highest array position until eight children, WideConverter.cu:291-324).  OPT = true: the children the C(n, i) table
// prescribes (GetChildrenIndices, BVH8Builder.cpp:165-199), leaf children holding up to max_leaf_prims primitives; slot
// assignment, quantisation and node layout are the GPU converter's in both modes.
template <bool OPT>
__device__ void collapse_one(const CollapseArgs& a, uint32_t self, uint32_t root2, uint32_t* levelCreated)
{
    const uint32_t lane = lane_id();
    const uint32_t n = a.n;
    uint32_t child[8];
    uint32_t innerMask = 0, count = 0;
    const bool live = root2 != NX_INVALID;
    Box parent; parent.lo = v3(0, 0, 0); parent.hi = v3(0, 0, 0);
    uint32_t slotOf = 0;   // 4 bits per slot: child index or 0xf

    if (live)
    {
        const float4 p = __ldg(a.n2 + 2 * (size_t)root2), q = __ldg(a.n2 + 2 * (size_t)root2 + 1);
        parent.lo = v3(p.x, p.y, p.z); parent.hi = v3(p.w, q.x, q.y);
        uint32_t l = __float_as_uint(q.z), r = __float_as_uint(q.w);
        int open = 0;
        if (OPT)
        {
            // entries: node | count << 28 | expand flag << 31; the left entry is pushed last so that it is processed first and
            // the children come out in the recursion's order
            uint32_t stack[9]; int sp = 0;
            auto decOf = [&](uint32_t node, uint32_t i) { return (uint32_t)(dp_dec_of(a.dpDec, n, node) >> (8 * i)) & 0xffu; };
            const uint32_t d0 = decOf(root2, 0);
            if ((d0 & 3u) == kDpLeaf) stack[sp++] = root2;                  // the whole tree is one leaf child
            else stack[sp++] = root2 | 0x80000000u;                          // count 0
            while (sp)
            {
                const uint32_t e = stack[--sp], node = e & 0x0fffffffu;
                if (!(e & 0x80000000u)) {
#pragma unroll
                    for (int k = 0; k < 8; k++) if (k == (int)count) child[k] = node;
                    if ((decOf(node, 0) & 3u) == kDpInternal) innerMask |= 1u << count;
                    count++;
                    continue;
                }
                const uint32_t d = decOf(node, (e >> 28) & 7u), lc = (d >> 2) & 7u, rc = (d >> 5) & 7u;
                const float4 nq = __ldg(a.n2 + 2 * (size_t)node + 1);
                const uint32_t L = __float_as_uint(nq.z), R = __float_as_uint(nq.w);
                stack[sp++] = R | (rc << 28) | ((decOf(R, rc) & 3u) == kDpDistribute ? 0x80000000u : 0u);
                stack[sp++] = L | (lc << 28) | ((decOf(L, lc) & 3u) == kDpDistribute ? 0x80000000u : 0u);
            }
        }
        else
        // Open inner children (highest array position first) until eight children or only leaves remain.  Of each opened
        // pair the smaller-area child takes the vacated position, the other is appended (WideConverter.cu:291-324).
        while (true)
        {
            const float al = half_area_ref(load_box2(a.n2, l)), ar = half_area_ref(load_box2(a.n2, r));
            const uint32_t first = al < ar ? l : r, second = al < ar ? r : l;
#pragma unroll
            for (int k = 0; k < 8; k++) {   // register-resident "child[open] = first; child[count] = second"
                if (k == open) child[k] = first;
            }
            if (first >= n) innerMask |= 1u << open;
            count++;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k == (int)count) child[k] = second;
            if (second >= n) innerMask |= 1u << count;
            count++;
            open = 31 - __clz(innerMask);
            if (open < 0 || count == 8) break;
            innerMask &= ~(1u << open);
            count--;
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k == open) o = child[k];
            const float4 oq = __ldg(a.n2 + 2 * (size_t)o + 1);
            l = __float_as_uint(oq.z); r = __float_as_uint(oq.w);
        }

        // Greedy octant-slot assignment (Ylitie et al. 2017 §4.2 as WideConverter.cu:106-153 implements it): children in
        // array order, each takes the free slot maximising (parentCentroid - childCentroid) . signs(slot); first max wins.
        const V3 pc = parent.lo + parent.hi;
        slotOf = 0xffffffffu;
        uint32_t freeSlots = 0xffu;
#pragma unroll
        for (int c = 0; c < 8; c++)
        {
            if (c < (int)count)
            {
                const Box cb = load_box2(a.n2, child[c]);
                const V3 off = pc - (cb.hi + cb.lo);
                float best = -3.402823466e38f; uint32_t bestSlot = 0xf;
#pragma unroll
                for (uint32_t s = 0; s < 8; s++)
                {
                    const float cost = __fadd_rn(__fadd_rn((s & 4) ? -off.x : off.x, (s & 2) ? -off.y : off.y), (s & 1) ? -off.z : off.z);
                    if (((freeSlots >> s) & 1u) && cost > best) { best = cost; bestSlot = s; }
                }
                freeSlots &= ~(1u << bestSlot);
                slotOf = (slotOf & ~(0xfu << (4 * bestSlot))) | ((uint32_t)c << (4 * bestSlot));
            }
        }
    }

    // slot-ordered masks
    uint32_t slotInner = 0, slotLeaf = 0;
#pragma unroll
    for (uint32_t s = 0; s < 8; s++) {
        const uint32_t c = (slotOf >> (4 * s)) & 0xfu;
        if (live && c != 0xfu) { if ((innerMask >> c) & 1u) slotInner |= 1u << s; else slotLeaf |= 1u << s; }
    }
    // primitives per leaf slot (4 bits each): one in the reference GPU mode, the subtree's count in the SAH-optimal mode
    uint32_t slotPrims = 0;
#pragma unroll
    for (uint32_t s = 0; s < 8; s++) {
        if (!((slotLeaf >> s) & 1u)) continue;
        uint32_t cnt = 1;
        if (OPT) {
            const uint32_t c = (slotOf >> (4 * s)) & 0xfu;
            uint32_t id = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) if (k == (int)c) id = child[k];
            cnt = (uint32_t)(dp_dec_of(a.dpDec, n, id) >> 56);
        }
        slotPrims |= cnt << (4 * s);
    }
    uint32_t nLeaf = 0;
#pragma unroll
    for (uint32_t s = 0; s < 8; s++) nLeaf += (slotPrims >> (4 * s)) & 0xfu;
    const uint32_t nInner = __popc(slotInner);

    // one atomic per warp and counter: inclusive scan of (inner | leaf << 16)
    uint32_t packed = nInner | (nLeaf << 16), scan = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(NX_FULL, scan, o); if (lane >= (uint32_t)o) scan += t; }
    const uint32_t total = __shfl_sync(NX_FULL, scan, 31);
    uint32_t baseNode = 0, baseLeaf = 0;
    if (lane == 31) {
        if (total & 0xffffu) { baseNode = atomicAdd(a.counters + 0, total & 0xffffu); atomicAdd(levelCreated, total & 0xffffu); }
        if (total >> 16) baseLeaf = atomicAdd(a.counters + 1, total >> 16);
    }
    baseNode = __shfl_sync(NX_FULL, baseNode, 31) + ((scan - packed) & 0xffffu);
    baseLeaf = __shfl_sync(NX_FULL, baseLeaf, 31) + ((scan - packed) >> 16);
    if (!live) return;

    const uint32_t childBase = nInner ? baseNode : 0u, primBase = nLeaf ? baseLeaf : 0u;

    // quantisation frame: origin = parent min, per-axis power-of-two cell >= extent / 255 (WideConverter.cu:160-171)
    float ext[3];
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[0]) : "f"(parent.hi.x), "f"(parent.lo.x));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[1]) : "f"(parent.hi.y), "f"(parent.lo.y));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[2]) : "f"(parent.hi.z), "f"(parent.lo.z));
    uint32_t e[3]; float inv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float s; asm("mul.ftz.f32 %0, %1, 0f3B808081;" : "=f"(s) : "f"(ext[k]));   // * (1.0f / 255.0f)
        e[k] = ceil_log2_biased(s) & 0xffu;
        inv[k] = __uint_as_float((254u - e[k]) << 23);
    }

    uint32_t meta[2] = {0, 0}, qlo[3][2] = {{0, 0}, {0, 0}, {0, 0}}, qhi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
    for (uint32_t s = 0; s < 8; s++)
    {
        const uint32_t c = (slotOf >> (4 * s)) & 0xfu;
        if (c == 0xfu) continue;
        uint32_t id = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) if (k == (int)c) id = child[k];
        uint32_t m;
        if ((slotInner >> s) & 1u) {
            m = 0x20u | (24u + s);
            a.bvh2Of[childBase + bits_below(slotInner, s)] = id;
        } else {
            uint32_t off = 0;
#pragma unroll
            for (uint32_t t = 0; t < 8; t++) if (t < s) off += (slotPrims >> (4 * t)) & 0xfu;
            const uint32_t cnt = (slotPrims >> (4 * s)) & 0xfu;
            m = (((1u << cnt) - 1u) << 5) | off;              // unary primitive count | first primitive (1 -> 0x20 | off)
            if (!OPT) a.primIdx[primBase + off] = id;         // a BVH2 leaf's index is its primitive id
            else {
                // the subtree's leaves, left to right (CountTriangles, BVH8Builder.cpp:282-293): at most three
                uint32_t st[4]; int sp = 0; uint32_t k = 0;
                st[sp++] = id;
                while (sp) {
                    const uint32_t u = st[--sp];
                    if (u < n) a.primIdx[primBase + off + k++] = u;
                    else { const float4 uq = __ldg(a.n2 + 2 * (size_t)u + 1); st[sp++] = __float_as_uint(uq.w); st[sp++] = __float_as_uint(uq.z); }
                }
            }
        }
        const Box cb = load_box2(a.n2, id);
        const uint32_t w = s >> 2, sh = (s & 3u) * 8u;
        meta[w] |= m << sh;
        qlo[0][w] |= quant(cb.lo.x, parent.lo.x, inv[0], false) << sh;
        qlo[1][w] |= quant(cb.lo.y, parent.lo.y, inv[1], false) << sh;
        qlo[2][w] |= quant(cb.lo.z, parent.lo.z, inv[2], false) << sh;
        qhi[0][w] |= quant(cb.hi.x, parent.lo.x, inv[0], true) << sh;
        qhi[1][w] |= quant(cb.hi.y, parent.lo.y, inv[1], true) << sh;
        qhi[2][w] |= quant(cb.hi.z, parent.lo.z, inv[2], true) << sh;
    }
    float4* out = a.n8 + 5 * (size_t)self;
    out[0] = make_float4(parent.lo.x, parent.lo.y, parent.lo.z, __uint_as_float(e[0] | (e[1] << 8) | (e[2] << 16) | (slotInner << 24)));
    out[1] = make_float4(__uint_as_float(childBase), __uint_as_float(primBase), __uint_as_float(meta[0]), __uint_as_float(meta[1]));
    out[2] = make_float4(__uint_as_float(qlo[0][0]), __uint_as_float(qlo[0][1]), __uint_as_float(qlo[1][0]), __uint_as_float(qlo[1][1]));
    out[3] = make_float4(__uint_as_float(qlo[2][0]), __uint_as_float(qlo[2][1]), __uint_as_float(qhi[0][0]), __uint_as_float(qhi[0][1]));
    out[4] = make_float4(__uint_as_float(qhi[1][0]), __uint_as_float(qhi[1][1]), __uint_as_float(qhi[2][0]), __uint_as_float(qhi[2][1]));
}

// Persistent cooperative kernel: BVH8 level L is exactly the node index range allocated while level L-1 was processed.
// The size of the next level cannot be read from the allocation cursor after the barrier: threads released early are already
// allocating for the level after it.  Each level therefore also counts what it creates in one of three rotating slots
// (counters[2 + (L + 1) % 3]); the slot a level adds to was cleared one level earlier, after the barrier that guarantees every
// thread has read its previous content, and is only read after the barrier that ends the level.
template <bool OPT>
__global__ void __launch_bounds__(kCollapseBlock, 4) collapse_kernel(CollapseArgs a)
{
    cg::grid_group grid = cg::this_grid();
    uint32_t begin = 0, end = 1, level = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    while (begin < end)
    {
        uint32_t* created = a.counters + 2 + (level + 1) % 3;
        if (blockIdx.x == 0 && threadIdx.x == 0) __stcg(a.counters + 2 + (level + 2) % 3, 0u);
        const uint32_t span = end - begin;
        const uint32_t rounds = (span + stride - 1) / stride;
        for (uint32_t r = 0; r < rounds; r++)
        {
            const uint32_t k = r * stride + blockIdx.x * blockDim.x + threadIdx.x;
            const uint32_t node = begin + k;
            collapse_one<OPT>(a, node, k < span ? __ldcg(a.bvh2Of + node) : NX_INVALID, created);   // whole warps call in, idle lanes pass INVALID
        }
        grid.sync();          // orders this level's stores before the next level's loads by itself: no __threadfence() per thread on top of it
                              // (that was a MEMBAR.SC.GPU + an L1 invalidation executed by every thread of the grid, every level)
        begin = end;
        end += __ldcg(created);
        level++;
    }
}

// The same level-by-level collapse inside ONE thread block, for inputs whose widest BVH8 level is a few rounds of a block: the barrier
// between levels is __syncthreads (tens of cycles) instead of a grid-wide barrier (13 us per level measured at 25 CTAs, 18 levels for
// the 100k-triangle sphere: profiles/r01_ncu_collapse_build100k.md shows 77 % of the stall samples at grid.sync), and the launch is an
// ordinary one, so the BLAS builds of a many-mesh scene overlap on the context's build streams instead of queueing as cooperative
// launches.  Two barriers per level: one to finish the level's allocations, one to read its size before the next level allocates.
constexpr uint32_t kDpWaveMinPrims = 200000u;   // below this the climb is as fast as ~40 launches
constexpr int kCollapseCtaThreads = 1024;
constexpr uint32_t kCollapseCtaMaxPrims = 40000u;   // widest level <= ~2 rounds of the block; beyond that the grid-wide kernel has more warps to hide the per-node latency
template <bool OPT>
__global__ void __launch_bounds__(kCollapseCtaThreads, 1) collapse_cta_kernel(CollapseArgs a)
{
    __shared__ uint32_t sEnd;
    uint32_t begin = 0, end = 1;
    while (begin < end)
    {
        const uint32_t span = end - begin;
        const uint32_t rounds = (span + kCollapseCtaThreads - 1) / kCollapseCtaThreads;
        for (uint32_t r = 0; r < rounds; r++)
        {
            const uint32_t k = r * kCollapseCtaThreads + threadIdx.x;
            collapse_one<OPT>(a, begin + k, k < span ? __ldcg(a.bvh2Of + begin + k) : NX_INVALID, a.counters + 2);
        }
        __threadfence_block();
        __syncthreads();
        if (threadIdx.x == 0) sEnd = __ldcg(a.counters);
        __syncthreads();
        begin = end;
        end = sEnd;
    }
}

// Single-primitive BVH: one node with one leaf child in slot 0 (WideConverter.cu:209-222).
__global__ void single_leaf_kernel(CollapseArgs a)
{
    const float4 p = a.n2[0], q = a.n2[1];
    float ext[3];
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[0]) : "f"(p.w), "f"(p.x));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[1]) : "f"(q.x), "f"(p.y));
    asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext[2]) : "f"(q.y), "f"(p.z));
    uint32_t e[3]; float inv[3];
    for (int k = 0; k < 3; k++) { float s; asm("mul.ftz.f32 %0, %1, 0f3B808081;" : "=f"(s) : "f"(ext[k])); e[k] = ceil_log2_biased(s) & 0xffu; inv[k] = __uint_as_float((254u - e[k]) << 23); }
    a.n8[0] = make_float4(p.x, p.y, p.z, __uint_as_float(e[0] | (e[1] << 8) | (e[2] << 16)));
    a.n8[1] = make_float4(__uint_as_float(0u), __uint_as_float(0u), __uint_as_float(0x20u), __uint_as_float(0u));
    a.n8[2] = make_float4(__uint_as_float(quant(p.x, p.x, inv[0], false)), 0.f, __uint_as_float(quant(p.y, p.y, inv[1], false)), 0.f);
    a.n8[3] = make_float4(__uint_as_float(quant(p.z, p.z, inv[2], false)), 0.f, __uint_as_float(quant(p.w, p.x, inv[0], true)), 0.f);
    a.n8[4] = make_float4(__uint_as_float(quant(q.x, p.y, inv[1], true)), 0.f, __uint_as_float(quant(q.y, p.z, inv[2], true)), 0.f);
    a.primIdx[0] = 0;
    a.counters[0] = 1; a.counters[1] = 1;
}

// ------------------------------------------------------------------------------------------------ refit ----
// Refit of a CWBVH8 built over boxes (a TLAS): same topology, same leaf order, every node's frame and child boxes recomputed bottom-up
// from new primitive boxes.  The reference rebuilds its TLAS from scratch on every instance change (Scene::BuildTLAS,
// src/Scene/Scene.cpp:65-78); a refit is what an interactive edit of a few instances wants - no sort, no hierarchy, no collapse, and the
// leaf order (hence every traversal record) stays where it is.  Hits cannot change (boxes stay conservative: each is the union of what
// lies below it, quantised outwards exactly like the collapse does it); the tree's quality degrades as objects move far, which is the
// caller's trade (nx_ctx_set_tlas_refit).
// One CTA: the trees this serves have a few thousand nodes at most.  The collapse numbers nodes level by level, so level L + 1 is the
// index range that follows level L and holds popc(imask) nodes per node of L; levels are found top-down, then processed bottom-up with a
// block barrier in between.  box[i]: the refitted bounds of node i (6 floats), read by its parent one level later.
constexpr int kRefitBlock = 512, kRefitMaxLevels = 64;
__global__ void __launch_bounds__(kRefitBlock) refit_kernel(float4* __restrict__ n8, uint32_t nodeCount, const uint32_t* __restrict__ primIdx,
                                                             const float* __restrict__ primBounds, float* __restrict__ box, uint32_t* status)
{
    __shared__ uint32_t levelEnd[kRefitMaxLevels + 1];
    __shared__ uint32_t red[kRefitBlock / 32];
    __shared__ uint32_t nLevels;
    // ---- level boundaries
    uint32_t begin = 0, end = 1, L = 0;
    if (threadIdx.x == 0) levelEnd[0] = 0;
    while (true) {
        uint32_t cnt = 0;
        for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) cnt += __popc(__float_as_uint(n8[5 * (size_t)i].w) >> 24);
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(NX_FULL, cnt, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
        __syncthreads();
        uint32_t total = 0;
        for (int w = 0; w < kRefitBlock / 32; w++) total += red[w];
        if (threadIdx.x == 0) levelEnd[L + 1] = end;
        __syncthreads();
        L++;
        if (total == 0 || L == kRefitMaxLevels) break;
        begin = end; end += total;
    }
    if (threadIdx.x == 0) { nLevels = L; if (end != nodeCount) atomicExch(status, 1u); }   // not a level-ordered tree of nodeCount nodes: refused
    __syncthreads();
    if (end != nodeCount) return;
    // ---- bottom-up
    for (int lev = (int)nLevels - 1; lev >= 0; lev--)
    {
        for (uint32_t i = levelEnd[lev] + threadIdx.x; i < levelEnd[lev + 1]; i += blockDim.x)
        {
            float4* nd = n8 + 5 * (size_t)i;
            const float4 n0 = nd[0], n1 = nd[1];
            const uint32_t imask = __float_as_uint(n0.w) >> 24, childBase = __float_as_uint(n1.x), primBase = __float_as_uint(n1.y);
            const uint32_t meta[2] = {__float_as_uint(n1.z), __float_as_uint(n1.w)};
            float clo[8][3], chi[8][3];
            float plo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, phi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
            for (uint32_t s = 0; s < 8; s++)
            {
                const uint32_t m = (meta[s >> 2] >> (8 * (s & 3u))) & 0xffu;
                for (int k = 0; k < 3; k++) { clo[s][k] = 3.0e38f; chi[s][k] = -3.0e38f; }
                if (!m) continue;
                if ((m & 0x1fu) >= 24u) {
                    const float* b = box + 6 * (size_t)(childBase + __popc(imask & ((1u << s) - 1u)));
                    for (int k = 0; k < 3; k++) { clo[s][k] = b[k]; chi[s][k] = b[3 + k]; }
                } else {
                    const uint32_t first = primBase + (m & 0x1fu), cnt = __popc(m >> 5);
                    for (uint32_t t = 0; t < cnt; t++) {
                        const float* b = primBounds + 6 * (size_t)primIdx[first + t];
                        for (int k = 0; k < 3; k++) { clo[s][k] = fminf(clo[s][k], b[k]); chi[s][k] = fmaxf(chi[s][k], b[3 + k]); }
                    }
                }
                for (int k = 0; k < 3; k++) { plo[k] = fminf(plo[k], clo[s][k]); phi[k] = fmaxf(phi[k], chi[s][k]); }
            }
            // quantisation frame and child planes exactly as the collapse writes them (collapse_one)
            uint32_t e[3]; float inv[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float ext, sc;
                asm("sub.ftz.f32 %0, %1, %2;" : "=f"(ext) : "f"(phi[k]), "f"(plo[k]));
                asm("mul.ftz.f32 %0, %1, 0f3B808081;" : "=f"(sc) : "f"(ext));
                e[k] = ceil_log2_biased(sc) & 0xffu;
                inv[k] = __uint_as_float((254u - e[k]) << 23);
            }
            uint32_t qlo[3][2] = {{0, 0}, {0, 0}, {0, 0}}, qhi[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
            for (uint32_t s = 0; s < 8; s++)
            {
                const uint32_t m = (meta[s >> 2] >> (8 * (s & 3u))) & 0xffu;
                if (!m) continue;
                const uint32_t w = s >> 2, sh = (s & 3u) * 8u;
#pragma unroll
                for (int k = 0; k < 3; k++) { qlo[k][w] |= quant(clo[s][k], plo[k], inv[k], false) << sh; qhi[k][w] |= quant(chi[s][k], plo[k], inv[k], true) << sh; }
            }
            nd[0] = make_float4(plo[0], plo[1], plo[2], __uint_as_float(e[0] | (e[1] << 8) | (e[2] << 16) | (imask << 24)));
            nd[2] = make_float4(__uint_as_float(qlo[0][0]), __uint_as_float(qlo[0][1]), __uint_as_float(qlo[1][0]), __uint_as_float(qlo[1][1]));
            nd[3] = make_float4(__uint_as_float(qlo[2][0]), __uint_as_float(qlo[2][1]), __uint_as_float(qhi[0][0]), __uint_as_float(qhi[0][1]));
            nd[4] = make_float4(__uint_as_float(qhi[1][0]), __uint_as_float(qhi[1][1]), __uint_as_float(qhi[2][0]), __uint_as_float(qhi[2][1]));
            float* b = box + 6 * (size_t)i;
            for (int k = 0; k < 3; k++) { b[k] = plo[k]; b[3 + k] = phi[k]; }
        }
        __syncthreads();
    }
}

// -------------------------------------------------------------------------------------------------- SAH ----
// Cost metrics as Eval.cu:12-80 defines them (BVH2: 3 per inner, 2 per leaf; BVH8: 2 per inner child, 3 per leaf child).
__global__ void bvh2_cost_kernel(const float4* __restrict__ n2, uint32_t nodeCount, Box scene, double* out)
{
    const float rootArea = half_area_ref(scene);
    double v = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nodeCount; i += gridDim.x * blockDim.x) {
        const float4 q = __ldg(n2 + 2 * (size_t)i + 1);
        v += (double)((__float_as_uint(q.z) != NX_INVALID ? 3.0f : 2.0f) * __fdividef(half_area_ref(load_box2(n2, i)), rootArea));
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NX_FULL, v, o);
    if (lane_id() == 0) atomicAdd(out, v);
}
__global__ void bvh8_cost_kernel(const float4* __restrict__ n8, uint32_t nodeCount, Box scene, double* out)   // out[0] cost, out[1] children
{
    const float rootArea = half_area_ref(scene);
    double v = 0.0, kids = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nodeCount; i += gridDim.x * blockDim.x) {
        const float4 h0 = __ldg(n8 + 5 * (size_t)i), h1 = __ldg(n8 + 5 * (size_t)i + 1), h2 = __ldg(n8 + 5 * (size_t)i + 2),
                     h3 = __ldg(n8 + 5 * (size_t)i + 3), h4 = __ldg(n8 + 5 * (size_t)i + 4);
        const uint32_t eim = __float_as_uint(h0.w);
        const float ex = __uint_as_float((eim & 0xffu) << 23), ey = __uint_as_float(((eim >> 8) & 0xffu) << 23), ez = __uint_as_float(((eim >> 16) & 0xffu) << 23);
        const uint32_t meta[2] = {__float_as_uint(h1.z), __float_as_uint(h1.w)};
        const uint32_t lox[2] = {__float_as_uint(h2.x), __float_as_uint(h2.y)}, loy[2] = {__float_as_uint(h2.z), __float_as_uint(h2.w)};
        const uint32_t loz[2] = {__float_as_uint(h3.x), __float_as_uint(h3.y)}, hix[2] = {__float_as_uint(h3.z), __float_as_uint(h3.w)};
        const uint32_t hiy[2] = {__float_as_uint(h4.x), __float_as_uint(h4.y)}, hiz[2] = {__float_as_uint(h4.z), __float_as_uint(h4.w)};
        for (uint32_t s = 0; s < 8; s++) {
            const uint32_t w = s >> 2, sh = (s & 3u) * 8u;
            const uint32_t m = (meta[w] >> sh) & 0xffu;
            if (!m) continue;
            kids += 1.0;
            Box b;
            b.lo = v3(__fmaf_rn(ex, (float)((lox[w] >> sh) & 0xffu), h0.x), __fmaf_rn(ey, (float)((loy[w] >> sh) & 0xffu), h0.y), __fmaf_rn(ez, (float)((loz[w] >> sh) & 0xffu), h0.z));
            b.hi = v3(__fmaf_rn(ex, (float)((hix[w] >> sh) & 0xffu), h0.x), __fmaf_rn(ey, (float)((hiy[w] >> sh) & 0xffu), h0.y), __fmaf_rn(ez, (float)((hiz[w] >> sh) & 0xffu), h0.z));
            v += (double)(((m & 0x1fu) >= 24u ? 2.0f : 3.0f) * __fdividef(half_area_ref(b), rootArea));
        }
    }
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(NX_FULL, v, o); kids += __shfl_xor_sync(NX_FULL, kids, o); }
    if (lane_id() == 0) { atomicAdd(out, v); atomicAdd(out + 1, kids); }
}

// ------------------------------------------------------------------------------------------ host driver ----
struct StageTimer {
    cudaStream_t s; bool on; cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(cudaStream_t s_, bool on_) : s(s_), on(on_) { if (on) { cudaEventCreate(&a); cudaEventCreate(&b); } }
    ~StageTimer() { if (on) { cudaEventDestroy(a); cudaEventDestroy(b); } }
    void begin() { if (on) cudaEventRecord(a, s); }
    float end() { if (!on) return 0.f; cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }
};

// Temporaries of a build.  Ordinary builds take them from the stream-ordered pool; a build issued by the scene set-up pipeline
// (nxi_build_bvh8_async) takes them from its build stream's workspace by bumping a pointer - no driver call, nothing to free - and its
// outputs from the scene's arena.
template <typename T> cudaError_t allocAsync(nx_ctx* ctx, T** p, size_t count, cudaStream_t s, bool output = false)
{
    const size_t bytes = (sizeof(T) * (count ? count : 1) + 255) & ~(size_t)255;
    nx_bump* b = output ? ctx->outArena : ctx->buildWs;
    if (b) {
        if (b->used + bytes > b->cap) return cudaErrorMemoryAllocation;
        *p = (T*)(b->base + b->used); b->used += bytes;
        return cudaSuccess;
    }
    return cudaMallocAsync((void**)p, bytes, s);
}
inline void freeAsync(nx_ctx* ctx, void* p, cudaStream_t s) { if (!ctx->buildWs) cudaFreeAsync(p, s); }

struct Bvh2Result { float4* nodes = nullptr; nx_aabb bounds{}; uint8_t* height = nullptr; /* optional: node heights for the level-wise C(n, i) pass */ };

template <typename KeyT>
int build_bvh2_keys(nx_ctx* ctx, uint32_t n, float4* nodes, SceneKeys* dScene, float* dSceneOut, nx_build_metrics* metrics, StageTimer& timer,
                    std::vector<uint64_t>* dbgCodes, uint8_t* heightOut)
{
    cudaStream_t s = ctx->stream;
    KeyT *keys = nullptr, *keysAlt = nullptr; uint32_t *order = nullptr, *orderAlt = nullptr, *parent = nullptr, *allocated = nullptr;
    NX_CUDA(ctx, allocAsync(ctx, &keys, n, s)); NX_CUDA(ctx, allocAsync(ctx, &keysAlt, n, s));
    NX_CUDA(ctx, allocAsync(ctx, &order, n, s)); NX_CUDA(ctx, allocAsync(ctx, &orderAlt, n, s));
    NX_CUDA(ctx, allocAsync(ctx, &parent, n, s)); NX_CUDA(ctx, allocAsync(ctx, &allocated, 1, s));
    NX_CUDA(ctx, cudaMemsetAsync(parent, 0xff, sizeof(uint32_t) * (size_t)n, s));
    NX_CUDA(ctx, cudaMemcpyAsync(allocated, &n, 4, cudaMemcpyHostToDevice, s));

    const bool ownSort = ctx->sort_mode != 0;
    uint32_t* sortHist = nullptr; uint32_t* sortStatus = nullptr;
    constexpr int kHistWords = SortShape<KeyT>::kPasses * 256 + SortShape<KeyT>::kPasses;     // histograms, then one tile counter per pass
    if (ownSort) {
        NX_CUDA(ctx, allocAsync(ctx, &sortHist, (size_t)kHistWords, s)); NX_CUDA(ctx, allocAsync(ctx, &sortStatus, radix_sort_status_words<KeyT>(n), s));
        NX_CUDA(ctx, cudaMemsetAsync(sortHist, 0, sizeof(uint32_t) * kHistWords, s));
    }
    const int grid = (int)std::min<uint32_t>(div_up(n, kSetupBlock), (uint32_t)ctx->sm_count * 8u);
    timer.begin();
    morton_kernel<KeyT><<<grid, kSetupBlock, 0, s>>>(nodes, n, dScene, keys, order, dSceneOut, sortHist);
    if (metrics) metrics->morton_ms = timer.end();
    if (dbgCodes) {
        dbgCodes->assign(n, 0);
        std::vector<KeyT> tmp(n);
        NX_CUDA(ctx, cudaMemcpyAsync(tmp.data(), keys, sizeof(KeyT) * (size_t)n, cudaMemcpyDeviceToHost, s));
        NX_CUDA(ctx, cudaStreamSynchronize(s));
        for (uint32_t i = 0; i < n; i++) (*dbgCodes)[i] = (uint64_t)tmp[i];
    }

    // stable LSD radix sort over the same bit window as the reference (Setup.cu:74-78): [2,32) or [1,64)
    KeyT* sortedKeys = keys; uint32_t* sortedOrder = order;
    void* temp = nullptr;
    if (ownSort) {
        timer.begin();
        NX_CUDA(ctx, radix_sort_pairs<KeyT>(keys, order, keysAlt, orderAlt, n, sortHist, sortStatus, sortHist + SortShape<KeyT>::kPasses * 256, s));
        if (metrics) metrics->sort_ms = timer.end();
        freeAsync(ctx, sortHist, s); freeAsync(ctx, sortStatus, s);
    } else {
#ifdef NX_WITH_CUB
        cub::DoubleBuffer<KeyT> kb(keys, keysAlt); cub::DoubleBuffer<uint32_t> vb(order, orderAlt);
        const int beginBit = sizeof(KeyT) == 4 ? 2 : 1, endBit = sizeof(KeyT) * 8;
        size_t tempBytes = 0;
        NX_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, kb, vb, (int)n, beginBit, endBit, s));
        NX_CUDA(ctx, cudaMallocAsync(&temp, tempBytes ? tempBytes : 1, s));
        timer.begin();
        NX_CUDA(ctx, cub::DeviceRadixSort::SortPairs(temp, tempBytes, kb, vb, (int)n, beginBit, endBit, s));
        if (metrics) metrics->sort_ms = timer.end();
        sortedKeys = kb.Current(); sortedOrder = vb.Current();
#else
        NX_FAIL(ctx, NX_ERR_INVALID, "NX_SORT=0 needs a library built with -DNX_WITH_CUB (measurement builds only)");
#endif
    }

    PlocArgs pa; pa.nodes = nodes; pa.cluster = sortedOrder; pa.parent = parent; pa.allocated = allocated; pa.n = n; pa.seedLo = pa.seedHi = nullptr; pa.height = heightOut;
    timer.begin();
    if (ctx->hploc_mode == 0) hploc_kernel<KeyT><<<div_up(n, kPlocBlock), kPlocBlock, 0, s>>>(pa, sortedKeys);
    else if (ctx->hploc_mode == 2) hploc_seed_kernel<KeyT><<<div_up(n, kPlocBlock), kPlocBlock, 0, s>>>(pa, sortedKeys, n);
    else {
        constexpr uint32_t C = kPlocChunk;
        static bool attr = false;   // per function, not per context: a property of the loaded module
        if (!attr) { cudaFuncSetAttribute((const void*)hploc2_kernel<KeyT, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChunkMem<C>)); attr = true; }
        const uint32_t chunks = div_up(n, C), nSeeds = 2 * chunks;
        NX_CUDA(ctx, allocAsync(ctx, &pa.seedLo, nSeeds, s)); NX_CUDA(ctx, allocAsync(ctx, &pa.seedHi, nSeeds, s));
        hploc2_kernel<KeyT, C><<<chunks, C, sizeof(ChunkMem<C>), s>>>(pa, sortedKeys);
        hploc_seed_kernel<KeyT><<<div_up(nSeeds, kPlocBlock), kPlocBlock, 0, s>>>(pa, sortedKeys, nSeeds);
        freeAsync(ctx, pa.seedLo, s); freeAsync(ctx, pa.seedHi, s);
    }
    if (metrics) metrics->bvh2_ms = timer.end();
    NX_CUDA(ctx, cudaGetLastError());

    if (temp) cudaFreeAsync(temp, s);
    freeAsync(ctx, keys, s); freeAsync(ctx, keysAlt, s); freeAsync(ctx, order, s); freeAsync(ctx, orderAlt, s);
    freeAsync(ctx, parent, s); freeAsync(ctx, allocated, s);
    return NX_OK;
}

int build_bvh2(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, const nx_build_config* cfg, nx_build_metrics* metrics,
               Bvh2Result* out, std::vector<uint64_t>* dbgCodes = nullptr, int forceBits64 = -1, bool finalSync = true, bool readBounds = true, bool wantHeights = false)
{
    if (!dPrims || n == 0) NX_FAIL(ctx, NX_ERR_INVALID, "BuildBVH2: empty primitive list");
    if (n > 0x7fffffffu / 2) NX_FAIL(ctx, NX_ERR_INVALID, "BuildBVH2: primitive count %u exceeds 2^30", n);
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    if (metrics) std::memset(metrics, 0, sizeof(*metrics));
    StageTimer timer(s, metrics != nullptr);

    float4* nodes = nullptr; SceneKeys* dScene = nullptr; float* dSceneOut = nullptr;
    NX_CUDA(ctx, allocAsync(ctx, &nodes, 2 * (size_t)(2 * (size_t)n - 1), s));
    NX_CUDA(ctx, allocAsync(ctx, &dScene, 1, s)); NX_CUDA(ctx, allocAsync(ctx, &dSceneOut, 6, s));
    SceneKeys init; for (int k = 0; k < 3; k++) { init.lo[k] = 0xffffffffu; init.hi[k] = 0u; }
    NX_CUDA(ctx, cudaMemcpyAsync(dScene, &init, sizeof(init), cudaMemcpyHostToDevice, s));

    const int grid = (int)std::min<uint32_t>(div_up(n, kSetupBlock), (uint32_t)ctx->sm_count * 8u);
    timer.begin();
    if (primType) leaf_bounds_kernel<9><<<grid, kSetupBlock, 0, s>>>((const float*)dPrims, n, nodes, dScene);
    else leaf_bounds_kernel<6><<<grid, kSetupBlock, 0, s>>>((const float*)dPrims, n, nodes, dScene);
    if (metrics) metrics->scene_bounds_ms = timer.end();

    uint8_t* height = nullptr;
    if (wantHeights && ctx->hploc_mode == 2) {       // only the global merge path records heights
        NX_CUDA(ctx, allocAsync(ctx, &height, 2 * (size_t)n - 1, s));
        NX_CUDA(ctx, cudaMemsetAsync(height, 0, 2 * (size_t)n - 1, s));
    }
    const bool bits64 = forceBits64 >= 0 ? forceBits64 != 0 : !(cfg && cfg->prioritize_speed);
    int rc = bits64 ? build_bvh2_keys<uint64_t>(ctx, n, nodes, dScene, dSceneOut, metrics, timer, dbgCodes, height)
                    : build_bvh2_keys<uint32_t>(ctx, n, nodes, dScene, dSceneOut, metrics, timer, dbgCodes, height);
    if (rc) return rc;
    out->height = height;
    if (readBounds) NX_CUDA(ctx, cudaMemcpyAsync(&out->bounds, dSceneOut, 24, cudaMemcpyDeviceToHost, s));
    if (metrics)
    {
        metrics->total_ms = metrics->scene_bounds_ms + metrics->morton_ms + metrics->sort_ms + metrics->bvh2_ms;
        NX_CUDA(ctx, cudaStreamSynchronize(s));
        double* dCost = nullptr; NX_CUDA(ctx, allocAsync(ctx, &dCost, 1, s));
        NX_CUDA(ctx, cudaMemsetAsync(dCost, 0, 8, s));
        Box sb; sb.lo = v3(out->bounds.bmin[0], out->bounds.bmin[1], out->bounds.bmin[2]); sb.hi = v3(out->bounds.bmax[0], out->bounds.bmax[1], out->bounds.bmax[2]);
        bvh2_cost_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(nodes, 2 * n - 1, sb, dCost);
        double cost = 0; NX_CUDA(ctx, cudaMemcpyAsync(&cost, dCost, 8, cudaMemcpyDeviceToHost, s));
        NX_CUDA(ctx, cudaStreamSynchronize(s));
        metrics->bvh2_cost = (float)cost;
        freeAsync(ctx, dCost, s);
    }
    freeAsync(ctx, dScene, s); freeAsync(ctx, dSceneOut, s);
    // BuildBVH8 continues on the same stream and synchronises once at its end: out->bounds is valid from then on
    if (finalSync) NX_CUDA(ctx, cudaStreamSynchronize(s));
    NX_CUDA(ctx, cudaGetLastError());
    out->nodes = nodes;
    return NX_OK;
}

// asyncCounters != nullptr: fire and forget.  Nothing is read back and the stream is never waited for; the node / leaf-slot counts stay
// in asyncCounters[0..1] (device memory of the caller, 5 words) and out->node_count / out->bounds are left to the caller.
int build_bvh8(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, const nx_build_config* cfg, nx_build_metrics* metrics, nx_bvh8* out,
               uint32_t* asyncCounters = nullptr)
{
    Bvh2Result b2;
    const bool wantHeights = cfg && cfg->collapse == NX_COLLAPSE_SAH_OPTIMAL && n >= kDpWaveMinPrims && !asyncCounters && ctx->dp_waves;
    int rc = build_bvh2(ctx, dPrims, n, primType, cfg, metrics, &b2, nullptr, -1, /*finalSync=*/false, /*readBounds=*/asyncCounters == nullptr, wantHeights);
    if (rc) return rc;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    StageTimer timer(s, metrics != nullptr);

    const size_t cap = ((size_t)4 * n - 1 + 6) / 7;   // worst case node count (BVHBuilder.cpp:184-186)
    CollapseArgs ca; ca.n2 = b2.nodes; ca.n = n; ca.dpDec = nullptr;
    NX_CUDA(ctx, allocAsync(ctx, &ca.n8, 5 * cap, s, /*output=*/true));
    NX_CUDA(ctx, allocAsync(ctx, &ca.primIdx, n, s, /*output=*/true));
    NX_CUDA(ctx, allocAsync(ctx, &ca.bvh2Of, cap, s));
    if (asyncCounters) ca.counters = asyncCounters; else NX_CUDA(ctx, allocAsync(ctx, &ca.counters, 5, s));
    const uint32_t initCounters[5] = {1u, 0u, 0u, 0u, 0u}, root2 = 2 * n - 2;
    NX_CUDA(ctx, cudaMemcpyAsync(ca.counters, initCounters, sizeof(initCounters), cudaMemcpyHostToDevice, s));
    NX_CUDA(ctx, cudaMemcpyAsync(ca.bvh2Of, &root2, 4, cudaMemcpyHostToDevice, s));

    const bool optimal = cfg && cfg->collapse == NX_COLLAPSE_SAH_OPTIMAL && n > 1;
    if (cfg && cfg->collapse != NX_COLLAPSE_REFERENCE_GPU && cfg->collapse != NX_COLLAPSE_SAH_OPTIMAL) NX_FAIL(ctx, NX_ERR_INVALID, "BuildBVH8: unknown collapse mode %d", cfg->collapse);
    DpArgs dp; std::memset(&dp, 0, sizeof(dp));
    timer.begin();
    if (optimal)
    {
        if (2ull * n - 1 > 0x0fffffffull) NX_FAIL(ctx, NX_ERR_INVALID, "BuildBVH8: the SAH-optimal collapse packs node ids into 28 bits (n = %u)", n);
        dp.n2 = b2.nodes; dp.n = n;
        dp.maxLeafPrims = cfg->max_leaf_prims >= 1 && cfg->max_leaf_prims <= 3 ? (uint32_t)cfg->max_leaf_prims : 3u;   // P_MAX, BVH8Builder.h:9
        NX_CUDA(ctx, allocAsync(ctx, &dp.cost, 8 * (2 * (size_t)n - 1), s));
        NX_CUDA(ctx, allocAsync(ctx, &dp.dec, 2 * (size_t)n - 1, s));
        bool waves = b2.height != nullptr;
        if (waves)
        {
            // level by level (dp_wave_kernel): histogram of the heights, counting sort of the inner nodes, one launch per height.  The
            // host needs the level sizes, so this path synchronises once; the set-up pipeline (asyncCounters) never records heights.
            DpWaveArgs w; w.height = b2.height;
            NX_CUDA(ctx, allocAsync(ctx, &w.hist, 512, s)); w.cursor = w.hist + 256;
            NX_CUDA(ctx, allocAsync(ctx, &w.order, n, s));
            NX_CUDA(ctx, cudaMemsetAsync(w.hist, 0, 4 * 512, s));
            dp_leaf_hist_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(dp, w);
            dp_height_scan_kernel<<<1, 256, 0, s>>>(w);
            uint32_t hist[256];
            NX_CUDA(ctx, cudaMemcpyAsync(hist, w.hist, sizeof(hist), cudaMemcpyDeviceToHost, s));
            dp_height_scatter_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(dp, w);
            NX_CUDA(ctx, cudaStreamSynchronize(s));
            if (hist[0] != 0u) NX_FAIL(ctx, NX_ERR_STATE, "BuildBVH8: %u inner nodes without a height", hist[0]);
            if (hist[255] != 0u) waves = false;          // a tree taller than 254 levels: the saturated heights no longer order the nodes; climb instead
            uint32_t first = 0;
            for (uint32_t h = 1; h < 255 && waves; h++) {
                if (!hist[h]) continue;
                dp_wave_kernel<<<div_up(hist[h], 128u), 128, 0, s>>>(dp, w.order, first, hist[h]);
                first += hist[h];
            }
            if (waves && first != n - 1) NX_FAIL(ctx, NX_ERR_STATE, "BuildBVH8: %u of %u inner nodes have a height", first, n - 1);
            freeAsync(ctx, w.hist, s); freeAsync(ctx, w.order, s);
        }
        if (!waves)
        {
            NX_CUDA(ctx, allocAsync(ctx, &dp.parent, 2 * (size_t)n - 1, s));
            NX_CUDA(ctx, allocAsync(ctx, &dp.arrived, n, s));
            NX_CUDA(ctx, cudaMemsetAsync(dp.parent, 0xff, 4 * (2 * (size_t)n - 1), s));
            dp_parent_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(dp);
            dp_eval_kernel<<<div_up(n, (uint32_t)kDpBlock), kDpBlock, 0, s>>>(dp);
        }
        ca.dpDec = dp.dec;
    }
    if (n == 1) single_leaf_kernel<<<1, 1, 0, s>>>(ca);
    else if (n <= kCollapseCtaMaxPrims && ctx->collapse_cta) {
        if (optimal) collapse_cta_kernel<true><<<1, kCollapseCtaThreads, 0, s>>>(ca);
        else collapse_cta_kernel<false><<<1, kCollapseCtaThreads, 0, s>>>(ca);
    }
    else {
        int perSm = 0;
        const void* fn = optimal ? (const void*)collapse_kernel<true> : (const void*)collapse_kernel<false>;
        NX_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, fn, kCollapseBlock, 0));
        // One thread per node of the widest level is all a level can use, and the grid-wide barrier between levels costs more the
        // more CTAs take part: n / 16 threads cover the widest level of every mesh measured (5,011 nodes for the 100,352-triangle
        // sphere) and cut the 100 k collapse from 231 to 190 us, the 1 M one from 317 to 236 us; from 2.4 M primitives on the grid is
        // the resident maximum as before.
        uint32_t grid = std::min<uint32_t>((uint32_t)(perSm * ctx->sm_count), std::max<uint32_t>(8u, div_up(n / 16u + 1u, kCollapseBlock)));
        void* args[] = {&ca};
        NX_CUDA(ctx, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kCollapseBlock), args, 0, s));
    }
    if (optimal) { if (dp.parent) { freeAsync(ctx, dp.parent, s); freeAsync(ctx, dp.arrived, s); } freeAsync(ctx, dp.cost, s); freeAsync(ctx, dp.dec, s); }
    if (b2.height) freeAsync(ctx, b2.height, s);
    if (metrics) { metrics->bvh8_ms = timer.end(); metrics->total_ms += metrics->bvh8_ms; }
    if (asyncCounters) {
        NX_CUDA(ctx, cudaGetLastError());
        out->nodes = (nx_bvh8_node*)ca.n8; out->node_count = 0; out->prim_idx = ca.primIdx; out->prim_count = n;
        freeAsync(ctx, b2.nodes, s); freeAsync(ctx, ca.bvh2Of, s);
        return NX_OK;
    }
    uint32_t counters[2] = {0, 0};
    NX_CUDA(ctx, cudaMemcpyAsync(counters, ca.counters, 8, cudaMemcpyDeviceToHost, s));
    NX_CUDA(ctx, cudaStreamSynchronize(s));
    NX_CUDA(ctx, cudaGetLastError());
    if (counters[1] != n) NX_FAIL(ctx, NX_ERR_STATE, "BuildBVH8: collapse placed %u of %u primitives", counters[1], n);

    out->nodes = (nx_bvh8_node*)ca.n8; out->node_count = counters[0]; out->prim_idx = ca.primIdx; out->prim_count = n; out->bounds = b2.bounds;
    if (metrics)
    {
        double* dCost = nullptr; NX_CUDA(ctx, allocAsync(ctx, &dCost, 2, s));
        NX_CUDA(ctx, cudaMemsetAsync(dCost, 0, 16, s));
        Box sb; sb.lo = v3(b2.bounds.bmin[0], b2.bounds.bmin[1], b2.bounds.bmin[2]); sb.hi = v3(b2.bounds.bmax[0], b2.bounds.bmax[1], b2.bounds.bmax[2]);
        bvh8_cost_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(ca.n8, counters[0], sb, dCost);
        double cost[2] = {0, 0}; NX_CUDA(ctx, cudaMemcpyAsync(cost, dCost, 16, cudaMemcpyDeviceToHost, s));
        NX_CUDA(ctx, cudaStreamSynchronize(s));
        metrics->bvh8_cost = (float)cost[0];
        metrics->avg_children_per_node = (float)(cost[1] / (double)counters[0]);   // = (n + nodes - 1) / nodes with one primitive per leaf
        freeAsync(ctx, dCost, s);
    }
    freeAsync(ctx, b2.nodes, s); freeAsync(ctx, ca.bvh2Of, s); freeAsync(ctx, ca.counters, s);
    NX_CUDA(ctx, cudaStreamSynchronize(s));
    return NX_OK;
}

} // namespace

// Internal entry used by the scene code (same translation-unit-free interface as the public C ABI).
int nxi_build_bvh8(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, int prioritizeSpeed, nx_bvh8* out, uint32_t maxLeafPrims)
{
    // TLAS (AABB primitives = instances): one instance per leaf child.  Entering an instance costs a ray transform and a BLAS
    // root visit, far more than the C_PRIM = 0.3 the cost model charges a primitive, so sharing a leaf box between two
    // instances (3.8 -> 5.3 instance candidates per ray on the CPU oracle) is a loss; BLAS leaves hold up to max_leaf_prims.
    nx_build_config cfg; cfg.prioritize_speed = prioritizeSpeed; cfg.collapse = ctx->scene_collapse;
    cfg.max_leaf_prims = maxLeafPrims ? maxLeafPrims : (primType ? ctx->scene_max_leaf_prims : 1);   // the merged BLAS is built over boxes of triangles
    return build_bvh8(ctx, dPrims, n, primType, &cfg, nullptr, out);
}

// Refit of a box-primitive BVH8 in place (refit_kernel).  dBounds: prim_count x nx_aabb in device memory, primitive order.
int nxi_refit_bvh8(nx_ctx* ctx, nx_bvh8* bvh, const void* dBounds)
{
    if (!bvh || !bvh->nodes || !dBounds || !bvh->node_count) return NX_ERR_INVALID;
    cudaStream_t s = ctx->stream;
    float* dBox = nullptr; uint32_t* dStatus = nullptr;
    NX_CUDA(ctx, cudaMallocAsync((void**)&dBox, 24 * (size_t)bvh->node_count + 4, s));
    dStatus = reinterpret_cast<uint32_t*>(dBox + 6 * (size_t)bvh->node_count);
    NX_CUDA(ctx, cudaMemsetAsync(dStatus, 0, 4, s));
    refit_kernel<<<1, kRefitBlock, 0, s>>>((float4*)bvh->nodes, bvh->node_count, bvh->prim_idx, (const float*)dBounds, dBox, dStatus);
    NX_CUDA(ctx, cudaGetLastError());
    float root[6]; uint32_t status = 0;
    NX_CUDA(ctx, cudaMemcpyAsync(root, dBox, 24, cudaMemcpyDeviceToHost, s));
    NX_CUDA(ctx, cudaMemcpyAsync(&status, dStatus, 4, cudaMemcpyDeviceToHost, s));
    NX_CUDA(ctx, cudaStreamSynchronize(s));
    cudaFreeAsync(dBox, s);
    if (status) NX_FAIL(ctx, NX_ERR_INVALID, "RefitBVH8: the node array is not a level-ordered tree of %u nodes (only trees built by this library can be refitted)", bvh->node_count);
    for (int k = 0; k < 3; k++) { bvh->bounds.bmin[k] = root[k]; bvh->bounds.bmax[k] = root[3 + k]; }
    return NX_OK;
}

// The same build issued on `stream` without any host synchronisation (scene set-up pipeline, scene.cu).
// ws: the build stream's workspace (temporaries; reset here - builds on one stream run one after the other), outputs: where the CWBVH8
// nodes and the leaf order go (the scene's arena).
int nxi_build_bvh8_async(nx_ctx* ctx, cudaStream_t stream, const void* dPrims, uint32_t n, int primType, int prioritizeSpeed, uint32_t* dCounters, nx_bvh8* out,
                         nx_bump* ws, nx_bump* outputs)
{
    nx_build_config cfg; cfg.prioritize_speed = prioritizeSpeed; cfg.collapse = ctx->scene_collapse;
    cfg.max_leaf_prims = primType ? ctx->scene_max_leaf_prims : 1;
    StreamSwap swap(ctx, stream);
    if (ws) ws->used = 0;
    ctx->buildWs = ws; ctx->outArena = outputs;
    const int rc = build_bvh8(ctx, dPrims, n, primType, &cfg, nullptr, out, dCounters);
    ctx->buildWs = nullptr; ctx->outArena = nullptr;
    return rc;
}
// upper bound of the temporaries of one build of n primitives (bytes): BVH2 nodes 64n, keys and order 2 x 12n, parent 4n, sort status,
// work map 2.3n, C(n, i) tables 2n x (32 + 8 + 4) + 4n, plus alignment slack per array
size_t nxi_build_workspace_bytes(uint32_t n) { return (size_t)n * 200u + (1u << 20); }

extern "C" {

int nx_bvh2_build_tri(nx_ctx* ctx, const nx_triangle* p, uint32_t n, const nx_build_config* cfg, nx_build_metrics* m, nx_bvh2* out)
{
    if (!ctx || !out) return NX_ERR_INVALID;
    Bvh2Result r; int rc = build_bvh2(ctx, p, n, 1, cfg, m, &r); if (rc) return rc;
    out->nodes = (nx_bvh2_node*)r.nodes; out->node_count = 2 * n - 1; out->prim_count = n; out->bounds = r.bounds; return NX_OK;
}
int nx_bvh2_build_aabb(nx_ctx* ctx, const nx_aabb* p, uint32_t n, const nx_build_config* cfg, nx_build_metrics* m, nx_bvh2* out)
{
    if (!ctx || !out) return NX_ERR_INVALID;
    Bvh2Result r; int rc = build_bvh2(ctx, p, n, 0, cfg, m, &r); if (rc) return rc;
    out->nodes = (nx_bvh2_node*)r.nodes; out->node_count = 2 * n - 1; out->prim_count = n; out->bounds = r.bounds; return NX_OK;
}
int nx_bvh8_build_tri(nx_ctx* ctx, const nx_triangle* p, uint32_t n, const nx_build_config* cfg, nx_build_metrics* m, nx_bvh8* out)
{
    if (!ctx || !out) return NX_ERR_INVALID;
    return build_bvh8(ctx, p, n, 1, cfg, m, out);
}
int nx_bvh8_build_aabb(nx_ctx* ctx, const nx_aabb* p, uint32_t n, const nx_build_config* cfg, nx_build_metrics* m, nx_bvh8* out)
{
    if (!ctx || !out) return NX_ERR_INVALID;
    return build_bvh8(ctx, p, n, 0, cfg, m, out);
}

int nx_bvh8_refit_aabb(nx_ctx* ctx, nx_bvh8* bvh, const nx_aabb* dBounds)
{
    if (!ctx || !bvh || !dBounds) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    return nxi_refit_bvh8(ctx, bvh, dBounds);
}

int nx_bvh2_to_host(nx_ctx* ctx, const nx_bvh2* b, nx_bvh2_node* host)
{
    if (!ctx || !b || !host) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    NX_CUDA(ctx, cudaMemcpyAsync(host, b->nodes, sizeof(nx_bvh2_node) * (size_t)b->node_count, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NX_OK;
}
int nx_bvh8_to_host(nx_ctx* ctx, const nx_bvh8* b, nx_bvh8_node* hostNodes, uint32_t* hostPrim)
{
    if (!ctx || !b) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    if (hostNodes) NX_CUDA(ctx, cudaMemcpyAsync(hostNodes, b->nodes, sizeof(nx_bvh8_node) * (size_t)b->node_count, cudaMemcpyDeviceToHost, ctx->stream));
    if (hostPrim) NX_CUDA(ctx, cudaMemcpyAsync(hostPrim, b->prim_idx, 4 * (size_t)b->prim_count, cudaMemcpyDeviceToHost, ctx->stream));
    NX_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return NX_OK;
}
int nx_bvh2_free(nx_ctx* ctx, nx_bvh2* b)
{
    if (!ctx || !b) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    if (b->nodes) cudaFreeAsync(b->nodes, ctx->stream);
    b->nodes = nullptr; b->node_count = 0;
    return NX_OK;
}
int nx_bvh8_free(nx_ctx* ctx, nx_bvh8* b)
{
    if (!ctx || !b) return NX_ERR_INVALID;
    DeviceGuard guard(ctx->device);
    if (b->nodes) cudaFreeAsync(b->nodes, ctx->stream);
    if (b->prim_idx) cudaFreeAsync(b->prim_idx, ctx->stream);
    b->nodes = nullptr; b->prim_idx = nullptr; b->node_count = 0;
    return NX_OK;
}

int nx_bvh8_benchmark(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, const nx_build_config* cfg, int warmup, int iters,
                      nx_build_metrics* metrics, uint32_t* outNodeCount)
{
    if (!ctx || !metrics || iters <= 0) return NX_ERR_INVALID;
    nx_build_metrics agg; std::memset(&agg, 0, sizeof(agg));
    uint32_t nodes = 0;
    for (int i = 0; i < warmup + iters; i++)
    {
        nx_build_metrics m; nx_bvh8 b;
        int rc = build_bvh8(ctx, dPrims, n, primType, cfg, &m, &b); if (rc) return rc;
        nodes = b.node_count;
        nx_bvh8_free(ctx, &b);
        if (i >= warmup) {
            float* d = (float*)&agg; const float* sre = (const float*)&m;
            for (size_t k = 0; k < sizeof(agg) / sizeof(float); k++) d[k] += sre[k];
        }
    }
    float* d = (float*)&agg;
    for (size_t k = 0; k < sizeof(agg) / sizeof(float); k++) d[k] /= (float)iters;
    *metrics = agg;
    if (outNodeCount) *outNodeCount = nodes;
    return NX_OK;
}

int nx_bvh_debug_morton(nx_ctx* ctx, const void* dPrims, uint32_t n, int primType, int bits64, uint64_t* hostCodes)
{
    if (!ctx || !hostCodes) return NX_ERR_INVALID;
    Bvh2Result r; std::vector<uint64_t> codes;
    int rc = build_bvh2(ctx, dPrims, n, primType, nullptr, nullptr, &r, &codes, bits64);
    if (rc) return rc;
    std::memcpy(hostCodes, codes.data(), 8 * (size_t)n);
    cudaFreeAsync(r.nodes, ctx->stream);
    return NX_OK;
}

} // extern "C"
