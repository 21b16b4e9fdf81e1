// TEST INFRASTRUCTURE — CPU restatement of the NexusBVH builder (H-PLOC BVH2 + BVH2->CWBVH8 collapse).
// See oracle_common.h for the rules on who may use this and how it is pinned.
//
// The reference builder is warp-cooperative CUDA whose *node numbering* depends on the GPU schedule
// (atomicAdd allocation, B/src/Cuda/BinaryBuilder.cu:88, WideConverter.cu:386-391) while the *tree* does not.
// This file simulates one valid schedule sequentially (32 "lanes" as arrays) and offers canonical
// renumberings so trees built by different schedules compare bit for bit.
#include "oracle_common.h"
#include <queue>

using namespace orc;

namespace {

constexpr uint32_t SEARCH_RADIUS = 8;       // BinaryBuilder.cu:9
constexpr uint32_t MERGING_THRESHOLD = 16;  // BinaryBuilder.cu:10

// ----------------------------------------------------------------------------------------------
// Setup.cu:13-39  leaf bounds + scene bounds
AABB primBounds(const float* p, int primType)
{
    AABB b;
    if (primType == 1) { // NXB::Triangle::Bounds -> AABB(v0, v1, v2), AABB.h:16-20
        f3 v0 = mk(p[0], p[1], p[2]), v1 = mk(p[3], p[4], p[5]), v2 = mk(p[6], p[7], p[8]);
        b.bMin = vmin(v0, vmin(v1, v2));
        b.bMax = vmax(v0, vmax(v1, v2));
    } else {
        b.bMin = mk(p[0], p[1], p[2]); b.bMax = mk(p[3], p[4], p[5]);
    }
    return b;
}

// BuilderUtils.h:108-196  bit interleave (the standard magic-number expansion)
uint32_t expand10(uint32_t x)
{
    x &= 0x3ff;
    x = (x | (x << 16)) & 0x30000ff;
    x = (x | (x << 8)) & 0x300f00f;
    x = (x | (x << 4)) & 0x30c30c3;
    x = (x | (x << 2)) & 0x9249249;
    return x;
}
uint64_t expand21(uint64_t x)
{
    x &= 0x1fffff;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}

// cvt.rzi.u32.f32 semantics (NaN -> 0, clamp to [0, 2^32-1])
uint32_t cvtRziU32(float f)
{
    if (!(f == f)) return 0;
    if (f <= 0.0f) return 0;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

// ----------------------------------------------------------------------------------------------
// BinaryBuilder.cu:16-28  Delta()
struct Keys {
    const uint64_t* code; bool bits64;
    uint64_t delta(uint32_t a, uint32_t b) const
    {
        if (bits64) { uint64_t d = code[a] ^ code[b]; return d == 0 ? (uint64_t)(a ^ b) : d; }
        return (((uint64_t)(uint32_t)code[a] << 32) | a) ^ (((uint64_t)(uint32_t)code[b] << 32) | b);
    }
};

struct Ploc {
    std::vector<Node2>& nodes;
    std::vector<uint32_t>& clusterIdx;
    uint32_t& clusterCount;

    // BinaryBuilder.cu:42-57
    uint32_t loadIndices(uint32_t start, uint32_t end, uint32_t lane_cluster[32], uint32_t offset)
    {
        uint32_t count = 0;
        uint32_t limit = std::min(end - start, MERGING_THRESHOLD);
        for (uint32_t lane = 0; lane < 32; lane++) {
            uint32_t index = lane - offset;             // unsigned wrap for lanes below offset, as on the GPU
            bool valid = index < limit;
            if (valid) lane_cluster[lane] = clusterIdx[start + index];
            if (valid && lane_cluster[lane] != INVALID) count++;
        }
        return count;
    }

    // BinaryBuilder.cu:127-163  radius-8 nearest neighbour, area compared on float bits
    void findNearestNeighbor(uint32_t numPrim, const AABB bounds[32], uint32_t nn[32])
    {
        uint32_t minArea[32], minIdx[32];
        for (int l = 0; l < 32; l++) { minArea[l] = INVALID; minIdx[l] = INVALID; }
        for (uint32_t r = 1; r <= SEARCH_RADIUS; r++) {
            uint32_t area[32];
            for (uint32_t l = 0; l < 32; l++) {
                uint32_t nb = l + r;
                area[l] = INVALID;
                if (nb < numPrim) {
                    AABB m = bounds[nb & 31];
                    m.grow(bounds[l]);
                    area[l] = f2u(ftz(m.area()));
                    if (area[l] < minArea[l]) { minArea[l] = area[l]; minIdx[l] = nb; }
                }
            }
            // neighbour (l + r) receives lane l's candidate; strict '<' keeps the earlier one on ties
            uint32_t nnA[32], nnI[32];
            for (uint32_t l = 0; l < 32; l++) {
                uint32_t src = (l + r) & 31;
                nnA[l] = minArea[src]; nnI[l] = minIdx[src];
                if (area[l] < nnA[l]) { nnA[l] = area[l]; nnI[l] = l; }
            }
            for (uint32_t l = 0; l < 32; l++) {
                uint32_t src = (l - r) & 31;
                minArea[l] = nnA[src]; minIdx[l] = nnI[src];
            }
        }
        for (int l = 0; l < 32; l++) nn[l] = minIdx[l];
    }

    // BinaryBuilder.cu:72-124  mutual-nearest-neighbour merge + compaction
    uint32_t mergeClusters(uint32_t numPrim, const uint32_t nn[32], uint32_t cluster[32], AABB bounds[32])
    {
        bool mutual[32], merge[32];
        uint32_t mergeMask = 0;
        for (uint32_t l = 0; l < 32; l++) {
            bool active = l < numPrim;
            uint32_t nnnn = nn[nn[l] & 31];
            mutual[l] = active && l == nnnn;
            merge[l] = mutual[l] && l < nn[l];
            if (merge[l]) mergeMask |= 1u << l;
        }
        uint32_t mergeCount = __builtin_popcount(mergeMask);
        uint32_t baseIdx = clusterCount;
        clusterCount += mergeCount;

        uint32_t oldCluster[32]; AABB oldBounds[32];
        std::memcpy(oldCluster, cluster, sizeof(oldCluster));
        std::memcpy(oldBounds, bounds, sizeof(oldBounds));
        for (uint32_t l = 0; l < 32; l++) {
            if (!merge[l]) continue;
            uint32_t rel = l == 0 ? 0 : __builtin_popcount(mergeMask << (32 - l));
            uint32_t nb = nn[l] & 31;
            bounds[l].grow(oldBounds[nb]);
            Node2 node; node.bounds = bounds[l]; node.left = oldCluster[l]; node.right = oldCluster[nb];
            cluster[l] = baseIdx + rel;
            nodes[cluster[l]] = node;
        }
        // compaction: keep merged lanes and non-mutual lanes (lanes >= numPrim count as "non mutual")
        uint32_t validMask = 0;
        for (uint32_t l = 0; l < 32; l++) if (merge[l] || !mutual[l]) validMask |= 1u << l;
        uint32_t c2[32]; AABB b2[32];
        std::memcpy(c2, cluster, sizeof(c2)); std::memcpy(b2, bounds, sizeof(b2));
        uint32_t m = validMask;
        for (uint32_t l = 0; l < 32; l++) {
            if (m) { uint32_t src = __builtin_ctz(m); m &= m - 1; cluster[l] = c2[src]; bounds[l] = b2[src]; }
            else { cluster[l] = INVALID; bounds[l] = b2[31]; }
        }
        return numPrim - mergeCount;
    }

    // BinaryBuilder.cu:165-197
    void merge(uint32_t left, uint32_t right, uint32_t split, bool final)
    {
        uint32_t lStart = left, rEnd = right + 1, lEnd = split, rStart = split;
        uint32_t cluster[32];
        for (int l = 0; l < 32; l++) cluster[l] = INVALID;
        uint32_t numLeft = loadIndices(lStart, lEnd, cluster, 0);
        uint32_t numRight = loadIndices(rStart, rEnd, cluster, numLeft);
        uint32_t numPrim = numLeft + numRight;
        AABB bounds[32];
        for (int l = 0; l < 32; l++) bounds[l].clear();
        for (uint32_t l = 0; l < numPrim; l++) bounds[l] = nodes[cluster[l]].bounds;
        uint32_t threshold = final ? 1 : MERGING_THRESHOLD;
        while (numPrim > threshold) {
            uint32_t nn[32];
            findNearestNeighbor(numPrim, bounds, nn);
            numPrim = mergeClusters(numPrim, nn, cluster, bounds);
        }
        uint32_t prev = numLeft + numRight;  // StoreIndices, BinaryBuilder.cu:59-69
        for (uint32_t l = 0; l < prev && l < 32; l++) clusterIdx[lStart + l] = cluster[l];
    }
};

// ----------------------------------------------------------------------------------------------
// WideConverter.cu helpers
uint32_t ceilLog2Bits(float x)  // BuilderUtils.h:335-342
{
    uint32_t ix = f2u(x);
    uint32_t exp = (ix >> 23) & 0xff;
    bool isPow2 = (ix & ((1u << 23) - 1)) == 0;
    return exp + !isPow2;
}
float invPow2(uint8_t eBiased) { return u2f((uint32_t)(254 - eBiased) << 23); }  // BuilderUtils.h:344-347
uint32_t nibble(uint32_t x, uint32_t i) { return (x >> (4 * i)) & 0xf; }
void setNibble(uint32_t& x, uint32_t i, uint32_t v) { x &= ~(0xfu << (4 * i)); x |= v << (4 * i); }
uint32_t bitsBelow(uint32_t x, uint32_t i) { return __builtin_popcount(x & ((1u << i) - 1)); }

uint8_t quantU8(float v, bool up)  // (uint8_t)floorf / ceilf then cvt.rzi.u32 and byte truncation, WideConverter.cu:195-201
{
    float r = up ? ceilf(ftz(v)) : floorf(ftz(v));
    return (uint8_t)cvtRziU32(r);
}

Node8 makeNode8(const std::vector<Node2>& n2, const AABB& bounds, const uint32_t child[8], uint32_t childBase, uint32_t primBase,
                uint32_t assignments, uint32_t innerMask, uint32_t leafMask)
{
    // WideConverter.cu:156-206
    Node8 nd; std::memset(&nd, 0, sizeof(nd));
    const float quantStep = 1.0f / 255.0f;  // constexpr quantStep, WideConverter.cu:19
    f3 diag = bounds.bMax - bounds.bMin;
    nd.p = bounds.bMin;
    nd.e[0] = (uint8_t)ceilLog2Bits(ftz(ftz(diag.x) * quantStep));
    nd.e[1] = (uint8_t)ceilLog2Bits(ftz(ftz(diag.y) * quantStep));
    nd.e[2] = (uint8_t)ceilLog2Bits(ftz(ftz(diag.z) * quantStep));
    nd.childBaseIdx = childBase; nd.primBaseIdx = primBase;
    f3 invE = mk(invPow2(nd.e[0]), invPow2(nd.e[1]), invPow2(nd.e[2]));
    for (uint32_t i = 0; i < 8; i++) {
        uint32_t a = nibble(assignments, i);
        if (innerMask & (1u << i)) { nd.imask |= 1u << i; nd.meta[i] = (1u << 5) | (24 + i); }
        else if (a != 0xf) nd.meta[i] = (1u << 5) | bitsBelow(leafMask, i);
        else continue;
        const AABB& cb = n2[child[a]].bounds;
        nd.qlox[i] = quantU8((cb.bMin.x - bounds.bMin.x) * invE.x, false);
        nd.qloy[i] = quantU8((cb.bMin.y - bounds.bMin.y) * invE.y, false);
        nd.qloz[i] = quantU8((cb.bMin.z - bounds.bMin.z) * invE.z, false);
        nd.qhix[i] = quantU8((cb.bMax.x - bounds.bMin.x) * invE.x, true);
        nd.qhiy[i] = quantU8((cb.bMax.y - bounds.bMin.y) * invE.y, true);
        nd.qhiz[i] = quantU8((cb.bMax.z - bounds.bMin.z) * invE.z, true);
    }
    return nd;
}

} // namespace

extern "C" {

// Leaf bounds and scene bounds.  Setup.cu:13-39.
void orc_prim_bounds(const float* prims, uint32_t n, int primType, float* outBounds, float* outScene6)
{
    AABB scene; scene.clear();
    size_t stride = primType ? 9 : 6;
    for (uint32_t i = 0; i < n; i++) {
        AABB b = primBounds(prims + stride * i, primType);
        std::memcpy(outBounds + 6 * (size_t)i, &b, 24);
        scene.grow(b);
    }
    std::memcpy(outScene6, &scene, 24);
}

// Morton codes.  Setup.cu:43-59 + BuilderUtils.h:225-253.  The reference normalises with div.approx.ftz.f32
// (fast-math); IEEE division is used here, so individual codes can differ from the GPU's by one quantisation
// cell.  Tests therefore feed the GPU-computed codes to orc_build_bvh2 and check these separately with a count.
void orc_morton(const float* bounds, uint32_t n, const float* scene6, int bits64, uint64_t* out)
{
    f3 smin = mk(scene6[0], scene6[1], scene6[2]), smax = mk(scene6[3], scene6[4], scene6[5]);
    f3 ext = mk(ftz(smax.x - smin.x), ftz(smax.y - smin.y), ftz(smax.z - smin.z));
    for (uint32_t i = 0; i < n; i++) {
        const float* b = bounds + 6 * (size_t)i;
        f3 c = mk(ftz(ftz(b[0] + b[3]) * 0.5f), ftz(ftz(b[1] + b[4]) * 0.5f), ftz(ftz(b[2] + b[5]) * 0.5f));
        f3 q = mk(ftz(ftz(c.x - smin.x) / ext.x), ftz(ftz(c.y - smin.y) / ext.y), ftz(ftz(c.z - smin.z) / ext.z));
        if (bits64) {
            uint64_t x = cvtRziU32(ftz(q.x * 2097151.0f)), y = cvtRziU32(ftz(q.y * 2097151.0f)), z = cvtRziU32(ftz(q.z * 2097151.0f));
            out[i] = expand21(x) | (expand21(y) << 1) | (expand21(z) << 2);
        } else {
            uint32_t x = cvtRziU32(ftz(q.x * 1023.0f)), y = cvtRziU32(ftz(q.y * 1023.0f)), z = cvtRziU32(ftz(q.z * 1023.0f));
            out[i] = expand10(x) | (expand10(y) << 1) | (expand10(z) << 2);
        }
    }
}

// Stable LSD radix sort contract of cub::DeviceRadixSort::SortPairs on bits [2,32) / [1,64): Setup.cu:63-112.
void orc_sort(const uint64_t* codes, uint32_t n, int bits64, uint64_t* outCodes, uint32_t* outIdx)
{
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; i++) idx[i] = i;
    int shift = bits64 ? 1 : 2;
    uint64_t mask = bits64 ? ~0ull : 0xffffffffull;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return ((codes[a] & mask) >> shift) < ((codes[b] & mask) >> shift); });
    for (uint32_t i = 0; i < n; i++) { outIdx[i] = idx[i]; outCodes[i] = codes[idx[i]]; }
}

// H-PLOC.  bounds: n*6 floats (leaf bounds in primitive order); codes: UNSORTED Morton codes (one uint64 per primitive,
// low 32 bits used when !bits64).  outNodes: (2n-1) * 32 B; leaves are [0,n) (leaf i = primitive i), inner nodes are
// numbered in this simulation's allocation order with the root last (2n-2), as in BuildBVH2Impl (BVHBuilder.cpp:14-111).
int orc_build_bvh2(const float* bounds, const uint64_t* codes, uint32_t n, int bits64, void* outNodes)
{
    if (n == 0) return -1;
    std::vector<Node2> nodes(2 * (size_t)n - 1);
    for (uint32_t i = 0; i < n; i++) {
        std::memcpy(&nodes[i].bounds, bounds + 6 * (size_t)i, 24);
        nodes[i].left = INVALID; nodes[i].right = i;
    }
    std::vector<uint64_t> sorted(n); std::vector<uint32_t> clusterIdx(n);
    orc_sort(codes, n, bits64, sorted.data(), clusterIdx.data());
    std::vector<uint32_t> parentIdx(n, INVALID);
    uint32_t clusterCount = n;
    Keys keys{sorted.data(), bits64 != 0};
    Ploc ploc{nodes, clusterIdx, clusterCount};

    // BuildBVH2Kernel (BinaryBuilder.cu:200-273), one "thread" at a time: climb until the sibling has not arrived yet.
    for (uint32_t idx = 0; idx < n; idx++) {
        uint32_t left = idx, right = idx, split = 0;
        while (true) {
            uint32_t prev;
            bool goRight = left == 0 || (right != n - 1 && keys.delta(right, right + 1) < keys.delta(left - 1, left));  // FindParentId :33-40
            if (goRight) {
                prev = parentIdx[right]; parentIdx[right] = left;
                if (prev != INVALID) { split = right + 1; right = prev; }
            } else {
                prev = parentIdx[left - 1]; parentIdx[left - 1] = right;
                if (prev != INVALID) { split = left; left = prev; }
            }
            if (prev == INVALID) break;
            uint32_t size = right - left + 1;
            bool final = size == n;
            if (size > MERGING_THRESHOLD || final) ploc.merge(left, right, split, final);
        }
    }
    std::memcpy(outNodes, nodes.data(), nodes.size() * sizeof(Node2));
    return clusterCount == 2 * n - 1 ? 0 : -2;
}

// BVH2 -> CWBVH8 collapse (BuildBVH8Kernel, WideConverter.cu:225-414) in FIFO (= canonical BFS) order.
// bvh2: (2n-1) nodes, root at index 2n-2 (BVHBuilder.cpp:195).  outNodes8 capacity ceil((4n-1)/7).
int orc_build_bvh8(const void* bvh2Nodes, uint32_t n, void* outNodes8, uint32_t* outPrimIdx, uint32_t* outNodeCount)
{
    std::vector<Node2> n2(2 * (size_t)n - 1);
    std::memcpy(n2.data(), bvh2Nodes, n2.size() * sizeof(Node2));
    Node8* out = (Node8*)outNodes8;
    if (n == 1) {  // CreateBVH8SingleLeaf, WideConverter.cu:209-222
        uint32_t child[8] = {0};
        out[0] = makeNode8(n2, n2[0].bounds, child, 0, 0, 0xfffffff0u, 0x0, 0x1);
        outPrimIdx[0] = 0;
        *outNodeCount = 1;
        return 0;
    }
    uint32_t nodeCounter = 1, leafCounter = 0;
    std::queue<std::pair<uint32_t, uint32_t>> work;  // (bvh2 node, bvh8 slot)
    work.push({2 * n - 2, 0});
    while (!work.empty()) {
        auto [i2, i8] = work.front(); work.pop();
        const Node2& node = n2[i2];
        if (node.left == INVALID) { outPrimIdx[i8] = node.right; continue; }

        uint32_t innerMask = 0, childCount = 0, child[8];
        uint32_t lr[2] = {node.left, node.right};
        int msb = 0;
        while (true) {  // :291-324
            float a0 = ftz(n2[lr[0]].bounds.area()), a1 = ftz(n2[lr[1]].bounds.area());
            uint32_t first = a0 < a1 ? 0 : 1;
            for (uint32_t i = 0; i < 2; i++) {
                uint32_t idx = i == 0 ? (uint32_t)msb : childCount;
                if (n2[lr[first]].left != INVALID) innerMask |= 1u << idx;
                child[idx] = lr[first];
                childCount++;
                first = !first;
            }
            msb = innerMask ? 31 - __builtin_clz(innerMask) : -1;
            if (msb < 0 || childCount == 8) break;
            innerMask &= ~(1u << msb);
            childCount--;
            uint32_t open = child[msb];
            lr[0] = n2[open].left; lr[1] = n2[open].right;
        }

        // GreedyAssignment, :106-153
        f3 parentCentroid = node.bounds.bMin + node.bounds.bMax;
        uint32_t assignments = INVALID;
        for (uint32_t c = 0; c < childCount; c++) {
            const AABB& cb = n2[child[c]].bounds;
            f3 off = parentCentroid - (cb.bMax + cb.bMin);
            float best = -FLT_MAX; uint32_t bestSlot = 0xf;
            for (uint32_t s = 0; s < 8; s++) {
                if (nibble(assignments, s) != 0xf) continue;
                float cost = ((s >> 2) & 1 ? -1.0f : 1.0f) * off.x + ((s >> 1) & 1 ? -1.0f : 1.0f) * off.y + (s & 1 ? -1.0f : 1.0f) * off.z;
                if (cost > best) { best = cost; bestSlot = s; }
            }
            setNibble(assignments, bestSlot, c);
        }
        uint32_t newInner = 0, leafMask = 0;
        for (uint32_t i = 0; i < 8; i++) {
            uint32_t a = nibble(assignments, i);
            if (a == 0xf) continue;
            bool bit = (innerMask >> a) & 1;
            newInner |= (uint32_t)bit << i; leafMask |= (uint32_t)(!bit) << i;
        }
        innerMask = newInner;
        uint32_t innerCount = __builtin_popcount(innerMask), leafCount = childCount - innerCount;
        uint32_t childBase = innerCount ? nodeCounter : 0; nodeCounter += innerCount;  // reference value is schedule-dependent when innerCount == 0
        uint32_t primBase = 0;
        if (leafCount > 0) { primBase = leafCounter; leafCounter += leafCount; }
        for (uint32_t i = 0; i < 8; i++) {
            uint32_t a = nibble(assignments, i);
            if (a == 0xf) continue;
            if (innerMask & (1u << i)) work.push({child[a], childBase + bitsBelow(innerMask, i)});
            else work.push({child[a], primBase + bitsBelow(leafMask, i)});
        }
        out[i8] = makeNode8(n2, node.bounds, child, childBase, primBase, assignments, innerMask, leafMask);
    }
    *outNodeCount = nodeCounter;
    return leafCounter == n ? 0 : -2;
}

// SAH-optimal collapse on the GPU builder's conventions: the C(n, i) table and child selection of the reference's CPU
// BVH8Builder (Nexus/src/Geometry/BVH/BVH8Builder.cpp:31-199, constants BVH8Builder.h:7-9; restated in full, with its own slot
// ordering and quantisation, in oracle_sah.cpp), emitted with the GPU converter's slot assignment (GreedyAssignment,
// WideConverter.cu:106-153), quantisation and node layout (CreateBVH8Node, :156-206), in FIFO (= canonical) order.  This is
// what nx_build_config::collapse = NX_COLLAPSE_SAH_OPTIMAL builds.  bvh2: root at 2n-2, leaves [0, n) in primitive order.
int orc_build_bvh8_optimal(const void* bvh2Nodes, uint32_t n, uint32_t maxLeafPrims, void* outNodes8, uint32_t* outPrimIdx, uint32_t* outNodeCount)
{
    if (n == 1) return orc_build_bvh8(bvh2Nodes, n, outNodes8, outPrimIdx, outNodeCount);
    const uint32_t total = 2 * n - 1, root = 2 * n - 2;
    std::vector<Node2> n2(total);
    std::memcpy(n2.data(), bvh2Nodes, total * sizeof(Node2));
    enum { LEAF = 0, INTERNAL = 1, DISTRIBUTE = 2, UNDEF = 3 };
    struct Ev { float cost; uint8_t dec = UNDEF, l = 0, r = 0; };
    std::vector<Ev> ev((size_t)total * 7);
    std::vector<uint32_t> tris(total, 0);
    // children before parents: post-order from the root, iterative
    std::vector<uint32_t> order; order.reserve(total);
    { std::vector<uint32_t> st{root}; while (!st.empty()) { uint32_t u = st.back(); st.pop_back(); order.push_back(u); if (n2[u].left != INVALID) { st.push_back(n2[u].left); st.push_back(n2[u].right); } } }
    for (size_t k = order.size(); k-- > 0;) {
        const uint32_t u = order[k];
        const Node2& nd = n2[u];
        const float area = ftz(nd.bounds.area());
        Ev* e = &ev[(size_t)u * 7];
        if (nd.left == INVALID) { tris[u] = 1; for (int i = 0; i < 7; i++) { e[i].cost = (area * 1.0f) * 0.3f; e[i].dec = LEAF; } continue; }
        tris[u] = std::min(255u, tris[nd.left] + tris[nd.right]);
        const Ev* cl = &ev[(size_t)nd.left * 7]; const Ev* cr = &ev[(size_t)nd.right * 7];
        auto distribute = [&](int j, uint8_t& l, uint8_t& r) {
            float best = 1.0e30f;
            for (int k2 = 0; k2 < j; k2++) { const float v = cl[k2].cost + cr[j - 1 - k2].cost; if (v < best) { best = v; l = (uint8_t)k2; r = (uint8_t)(j - 1 - k2); } }
            return best;
        };
        { uint8_t l = 0, r = 0;
          const float internal = distribute(7, l, r) + area * 1.0f;
          const float leaf = tris[u] > maxLeafPrims ? 1.0e30f : (area * (float)tris[u]) * 0.3f;
          if (leaf < internal) { e[0].cost = leaf; e[0].dec = LEAF; } else { e[0].cost = internal; e[0].dec = INTERNAL; e[0].l = l; e[0].r = r; } }
        for (int i = 1; i < 7; i++) {
            uint8_t l = 0, r = 0;
            const float d = distribute(i, l, r);
            if (d < e[i - 1].cost) { e[i].cost = d; e[i].dec = DISTRIBUTE; e[i].l = l; e[i].r = r; } else e[i] = e[i - 1];
        }
    }
    auto dec = [&](uint32_t u, int i) -> const Ev& { return ev[(size_t)u * 7 + i]; };

    Node8* out = (Node8*)outNodes8;
    uint32_t nodeCounter = 1, leafCounter = 0;
    std::queue<std::pair<uint32_t, uint32_t>> work;
    work.push({root, 0});
    while (!work.empty()) {
        auto [i2, i8] = work.front(); work.pop();
        const Node2& node = n2[i2];
        uint32_t child[8], childCount = 0, innerMask = 0;
        // GetChildrenIndices (BVH8Builder.cpp:165-199), iteratively: entries are (node, count, expand)
        struct Ent { uint32_t node; int cnt; bool expand; };
        std::vector<Ent> st;
        if (dec(i2, 0).dec == LEAF) st.push_back({i2, 0, false}); else st.push_back({i2, 0, true});
        while (!st.empty()) {
            Ent e = st.back(); st.pop_back();
            if (!e.expand) { if (dec(e.node, 0).dec == INTERNAL) innerMask |= 1u << childCount; child[childCount++] = e.node; continue; }
            const Ev& d = dec(e.node, e.cnt);
            const uint32_t L = n2[e.node].left, R = n2[e.node].right;
            st.push_back({R, d.r, dec(R, d.r).dec == DISTRIBUTE});
            st.push_back({L, d.l, dec(L, d.l).dec == DISTRIBUTE});
        }
        // GreedyAssignment, WideConverter.cu:106-153
        f3 parentCentroid = node.bounds.bMin + node.bounds.bMax;
        uint32_t assignments = INVALID;
        for (uint32_t c = 0; c < childCount; c++) {
            const AABB& cb = n2[child[c]].bounds;
            f3 off = parentCentroid - (cb.bMax + cb.bMin);
            float best = -FLT_MAX; uint32_t bestSlot = 0xf;
            for (uint32_t s = 0; s < 8; s++) {
                if (nibble(assignments, s) != 0xf) continue;
                float cost = ((s >> 2) & 1 ? -1.0f : 1.0f) * off.x + ((s >> 1) & 1 ? -1.0f : 1.0f) * off.y + (s & 1 ? -1.0f : 1.0f) * off.z;
                if (cost > best) { best = cost; bestSlot = s; }
            }
            setNibble(assignments, bestSlot, c);
        }
        uint32_t newInner = 0, leafMask = 0, leafPrims = 0, primsOf[8] = {0};
        for (uint32_t i = 0; i < 8; i++) {
            uint32_t a = nibble(assignments, i);
            if (a == 0xf) continue;
            bool bit = (innerMask >> a) & 1;
            newInner |= (uint32_t)bit << i; leafMask |= (uint32_t)(!bit) << i;
            if (!bit) { primsOf[i] = tris[child[a]]; leafPrims += primsOf[i]; }
        }
        innerMask = newInner;
        uint32_t innerCount = __builtin_popcount(innerMask);
        uint32_t childBase = innerCount ? nodeCounter : 0; nodeCounter += innerCount;
        uint32_t primBase = leafPrims ? leafCounter : 0; leafCounter += leafPrims;
        Node8 nd = makeNode8(n2, node.bounds, child, childBase, primBase, assignments, innerMask, leafMask);
        uint32_t off = 0;
        for (uint32_t i = 0; i < 8; i++) {
            uint32_t a = nibble(assignments, i);
            if (a == 0xf) continue;
            if (innerMask & (1u << i)) { work.push({child[a], childBase + bitsBelow(innerMask, i)}); continue; }
            nd.meta[i] = (uint8_t)((((1u << primsOf[i]) - 1u) << 5) | off);
            // the subtree's leaves left to right (CountTriangles, BVH8Builder.cpp:282-293)
            std::vector<uint32_t> ls{child[a]}; uint32_t k = 0;
            while (!ls.empty()) { uint32_t u = ls.back(); ls.pop_back(); if (n2[u].left == INVALID) outPrimIdx[primBase + off + k++] = n2[u].right; else { ls.push_back(n2[u].right); ls.push_back(n2[u].left); } }
            off += primsOf[i];
        }
        out[i8] = nd;
    }
    *outNodeCount = nodeCounter;
    return leafCounter == n ? 0 : -2;
}

// Canonical renumbering of a BVH8 (root = node 0): breadth-first, children in slot order.  Node payloads are copied
// byte for byte except childBaseIdx / primBaseIdx.  Returns 0, or <0 if the input is structurally broken.
int orc_canon_bvh8(const void* nodesIn, const uint32_t* primIdxIn, uint32_t nodeCount, uint32_t primCount,
                   void* nodesOut, uint32_t* primIdxOut)
{
    const Node8* in = (const Node8*)nodesIn;
    Node8* out = (Node8*)nodesOut;
    std::queue<std::pair<uint32_t, uint32_t>> work;  // (old id, new id)
    work.push({0, 0});
    uint32_t nodeCounter = 1, leafCounter = 0;
    std::vector<uint8_t> seen(nodeCount, 0);
    while (!work.empty()) {
        auto [o, nw] = work.front(); work.pop();
        if (o >= nodeCount || seen[o]) return -1;
        seen[o] = 1;
        Node8 nd = in[o];
        uint32_t innerCount = __builtin_popcount(nd.imask);
        uint32_t leafPrims = 0;
        for (int i = 0; i < 8; i++) if (nd.meta[i] && !(nd.imask & (1u << i))) leafPrims += __builtin_popcount(nd.meta[i] >> 5);
        uint32_t childBase = innerCount ? nodeCounter : 0; nodeCounter += innerCount;
        uint32_t primBase = leafPrims ? leafCounter : 0;
        for (int i = 0; i < 8; i++) {
            if (!nd.meta[i]) continue;
            if (nd.imask & (1u << i)) work.push({nd.childBaseIdx + bitsBelow(nd.imask, i), childBase + bitsBelow(nd.imask, i)});
            else {
                uint32_t cnt = __builtin_popcount(nd.meta[i] >> 5), off = nd.meta[i] & 0x1f;
                for (uint32_t k = 0; k < cnt; k++) {
                    if (nd.primBaseIdx + off + k >= primCount || primBase + off + k >= primCount) return -2;
                    primIdxOut[primBase + off + k] = primIdxIn[nd.primBaseIdx + off + k];
                }
            }
        }
        leafCounter += leafPrims;
        nd.childBaseIdx = childBase; nd.primBaseIdx = primBase;
        // the reference never initialises the quantised boxes of empty slots (NodeExplicit is a stack temporary,
        // WideConverter.cu:160); they carry no information, so the canonical form zeroes them
        for (int i = 0; i < 8; i++) if (!nd.meta[i]) nd.qlox[i] = nd.qloy[i] = nd.qloz[i] = nd.qhix[i] = nd.qhiy[i] = nd.qhiz[i] = 0;
        if (nw >= nodeCount) return -3;
        out[nw] = nd;
    }
    return (nodeCounter == nodeCount && leafCounter == primCount) ? 0 : -4;
}

// Canonical renumbering of a BVH2: leaves stay [0,n); inner nodes are renumbered in post-order (left subtree, right
// subtree, node), so the root lands on 2n-2.  Iterative to survive degenerate depth.
int orc_canon_bvh2(const void* nodesIn, uint32_t n, void* nodesOut)
{
    const Node2* in = (const Node2*)nodesIn;
    Node2* out = (Node2*)nodesOut;
    uint32_t total = 2 * n - 1;
    for (uint32_t i = 0; i < n; i++) out[i] = in[i];
    if (n == 1) return 0;
    std::vector<uint32_t> newId(total, INVALID);
    for (uint32_t i = 0; i < n; i++) newId[i] = i;
    struct Frame { uint32_t node; int state; };
    std::vector<Frame> stack; stack.push_back({total - 1, 0});
    uint32_t next = n;
    while (!stack.empty()) {
        Frame& f = stack.back();
        const Node2& nd = in[f.node];
        if (nd.left == INVALID) { stack.pop_back(); continue; }
        if (nd.left >= total || nd.right >= total) return -1;
        if (f.state == 0) { f.state = 1; stack.push_back({nd.left, 0}); }
        else if (f.state == 1) { f.state = 2; stack.push_back({nd.right, 0}); }
        else {
            if (next >= total) return -2;
            newId[f.node] = next;
            out[next].bounds = nd.bounds; out[next].left = newId[nd.left]; out[next].right = newId[nd.right];
            next++;
            stack.pop_back();
        }
    }
    return next == total ? 0 : -3;
}

// SAH costs exactly as Eval.cu:12-80 defines them (including the swapped C_I / C_T in the BVH8 variant), summed in
// double so the value is independent of the GPU's float-atomic summation order.
double orc_bvh2_cost(const void* nodes, uint32_t nodeCount, const float* scene6)
{
    const Node2* nd = (const Node2*)nodes;
    AABB s; std::memcpy(&s, scene6, 24);
    double root = s.area(), cost = 0.0;
    for (uint32_t i = 0; i < nodeCount; i++) cost += (nd[i].left != INVALID ? 3.0 : 2.0) * ((double)nd[i].bounds.area() / root);
    return cost;
}

double orc_bvh8_cost(const void* nodes, uint32_t nodeCount, const float* scene6)
{
    const Node8* nd = (const Node8*)nodes;
    AABB s; std::memcpy(&s, scene6, 24);
    double root = s.area(), cost = 0.0;
    for (uint32_t n = 0; n < nodeCount; n++) {
        const Node8& N = nd[n];
        float ex = u2f((uint32_t)N.e[0] << 23), ey = u2f((uint32_t)N.e[1] << 23), ez = u2f((uint32_t)N.e[2] << 23);
        for (int i = 0; i < 8; i++) {
            if (!N.meta[i]) continue;
            bool internal = (N.meta[i] & 0x1f) >= 24;
            AABB b;
            b.bMin = mk(N.p.x + ex * N.qlox[i], N.p.y + ey * N.qloy[i], N.p.z + ez * N.qloz[i]);
            b.bMax = mk(N.p.x + ex * N.qhix[i], N.p.y + ey * N.qhiy[i], N.p.z + ez * N.qhiz[i]);
            cost += (internal ? 2.0 : 3.0) * ((double)b.area() / root);
        }
    }
    return cost;
}

// Structural invariants implied by the reference.  Returns 0 when all hold.
int orc_check_bvh8(const void* nodes, const uint32_t* primIdx, uint32_t nodeCount, uint32_t primCount, const float* primBounds /* n*6 */)
{
    const Node8* nd = (const Node8*)nodes;
    if (nodeCount > (4ull * primCount - 1 + 6) / 7) return -1;
    std::vector<uint8_t> seenPrim(primCount, 0);
    // every decoded child box must contain the true bounds of everything below it: check leaves directly
    std::vector<std::pair<uint32_t, AABB>> stack;  // node, decoded box from parent
    AABB any; any.bMin = mk(-FLT_MAX, -FLT_MAX, -FLT_MAX); any.bMax = mk(FLT_MAX, FLT_MAX, FLT_MAX);
    stack.push_back({0, any});
    uint32_t visited = 0;
    while (!stack.empty()) {
        auto [ni, box] = stack.back(); stack.pop_back();
        if (ni >= nodeCount) return -2;
        visited++;
        const Node8& N = nd[ni];
        float ex = u2f((uint32_t)N.e[0] << 23), ey = u2f((uint32_t)N.e[1] << 23), ez = u2f((uint32_t)N.e[2] << 23);
        for (int i = 0; i < 8; i++) {
            if (!N.meta[i]) continue;
            AABB b;
            b.bMin = mk(N.p.x + ex * N.qlox[i], N.p.y + ey * N.qloy[i], N.p.z + ez * N.qloz[i]);
            b.bMax = mk(N.p.x + ex * N.qhix[i], N.p.y + ey * N.qhiy[i], N.p.z + ez * N.qhiz[i]);
            if (N.imask & (1u << i)) stack.push_back({N.childBaseIdx + bitsBelow(N.imask, i), b});
            else {
                uint32_t cnt = __builtin_popcount(N.meta[i] >> 5), off = N.meta[i] & 0x1f;
                for (uint32_t k = 0; k < cnt; k++) {
                    uint32_t slot = N.primBaseIdx + off + k;
                    if (slot >= primCount) return -3;
                    uint32_t prim = primIdx[slot];
                    if (prim >= primCount || seenPrim[prim]) return -4;
                    seenPrim[prim] = 1;
                    const float* pb = primBounds + 6 * (size_t)prim;
                    // fl(cmin - p) may round across a grid line, so allow a sliver of one quantisation cell
                    const float sx = ex * (1.0f / 1024.0f), sy = ey * (1.0f / 1024.0f), sz = ez * (1.0f / 1024.0f);
                    if (pb[0] < b.bMin.x - sx || pb[1] < b.bMin.y - sy || pb[2] < b.bMin.z - sz ||
                        pb[3] > b.bMax.x + sx || pb[4] > b.bMax.y + sy || pb[5] > b.bMax.z + sz) return -5;
                }
            }
        }
    }
    if (visited != nodeCount) return -6;
    for (uint32_t i = 0; i < primCount; i++) if (!seenPrim[i]) return -7;
    return 0;
}

} // extern "C"
