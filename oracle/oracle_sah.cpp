// TEST INFRASTRUCTURE — CPU path of BASELINE.json configs[0]: a binned-SAH BVH2 build followed by the SAH-optimal
// BVH2 -> CWBVH8 collapse (Ylitie, Karras, Laine 2017), host only.
//
// (1) Collapse: restates the reference's CPU BVH8Builder (Nexus/src/Geometry/BVH/BVH8Builder.cpp:11-390, constants
//     BVH8Builder.h:7-10).  That class is dead code in the reference snapshot (nothing instantiates it, SURVEY.md header
//     note 1) and cannot run as written: Build() never allocates bvh8.primIdx although CountTriangles() writes it (:21-29 vs
//     :261).  The restatement keeps its semantics - C(n, i) table with LEAF / INTERNAL / DISTRIBUTE decisions, C_PRIM 0.3,
//     C_NODE 1.0, P_MAX 3, greedy global-minimum slot assignment, e = ceil(log2(extent / 255)) quantisation - and only adds
//     what it needs to run: the primIdx allocation, a guard for zero-extent axes, and a root that may be a single leaf.
//     PINNED against the reference's own code: BVH8Builder.cpp compiles unmodified with g++ (oracle/Makefile target `refcpu`,
//     harness oracle/ref/ref_cpu_collapse.cpp allocates primIdx and calls Init() / CollapseNode()); on 10 inputs from 1 to
//     100,352 triangles this restatement reproduces its nodes, primitive order and C(root, 1) bit for bit
//     (tests/golden/cpu_collapse_ref.npz, tests/test_oracle.py).  Areas are evaluated as the host compiler does (areaHost(), no
//     FMA contraction), not as the GPU builder does.  Not pinned: inputs with a zero-extent axis, where the reference takes
//     log2f(0) and casts NaN to uint8_t (undefined behaviour).
// (2) Binned SAH BVH2: the reference README advertises "Standard SAH-based BVH (BVH2) using binned building"
//     (README.md:31) but the snapshot contains no such source; this is written from that description with the choices
//     SURVEY.md §8(c) records: 8 bins per axis over the centroid bounds, cost A(L)·N(L) + A(R)·N(R), one primitive per leaf
//     (BVH8Builder requires it, BVH8Builder.cpp:71-78), object-median fallback when binning cannot separate, root at node 0
//     (BVH8Builder.cpp:16-17, 26).  PARITY UNPINNED: there is no reference code, test or vector for this builder; the
//     tests check structural invariants and that traversal of the resulting tree equals brute force.
// Only tests/ and bench.py's cpu_baseline leg use this file.
#include "oracle_common.h"
#include <thread>
#include <numeric>

using namespace orc;

namespace {

constexpr float kCPrim = 0.3f, kCNode = 1.0f;   // BVH8Builder.h:7-8
constexpr int kPMax = 3;                          // BVH8Builder.h:9
constexpr int kBins = 8;

// ------------------------------------------------------------------------------------------- binned SAH BVH2 ----
struct Sah2 {
    const AABB* pb; Node2* nodes; uint32_t* idx; int maxParDepth;

    // nodes of the subtree over idx[lo, hi) live in [base, base + 2 * (hi - lo) - 1): numbering is schedule independent
    void build(uint32_t base, uint32_t lo, uint32_t hi, int depth)
    {
        AABB box; box.clear(); AABB cb; cb.clear();
        for (uint32_t i = lo; i < hi; i++) { const AABB& b = pb[idx[i]]; box.grow(b); cb.grow((b.bMin + b.bMax) * 0.5f); }
        Node2& nd = nodes[base];
        nd.bounds = box;
        const uint32_t n = hi - lo;
        if (n == 1) { nd.left = INVALID; nd.right = idx[lo]; return; }   // leaf: rightChild = primitive id (BVH.h:20-29)

        int bestAxis = -1, bestSplit = 0; float bestCost = FLT_MAX;
        const float ext[3] = {cb.bMax.x - cb.bMin.x, cb.bMax.y - cb.bMin.y, cb.bMax.z - cb.bMin.z};
        const float mn[3] = {cb.bMin.x, cb.bMin.y, cb.bMin.z};
        for (int a = 0; a < 3; a++) {
            if (!(ext[a] > 0.0f)) continue;
            AABB bb[kBins]; uint32_t cnt[kBins] = {};
            for (auto& b : bb) b.clear();
            const float scale = (float)kBins / ext[a];
            for (uint32_t i = lo; i < hi; i++) {
                const AABB& b = pb[idx[i]];
                const float c = 0.5f * ((&b.bMin.x)[a] + (&b.bMax.x)[a]);
                const int k = std::min(kBins - 1, (int)((c - mn[a]) * scale));
                bb[k].grow(b); cnt[k]++;
            }
            float rightArea[kBins]; uint32_t rightCnt[kBins];
            AABB acc; acc.clear(); uint32_t c = 0;
            for (int k = kBins - 1; k > 0; k--) { if (cnt[k]) acc.grow(bb[k]); c += cnt[k]; rightArea[k] = c ? acc.area() : 0.0f; rightCnt[k] = c; }
            acc.clear(); c = 0;
            for (int k = 1; k < kBins; k++) {                     // split between bin k-1 and bin k
                if (cnt[k - 1]) acc.grow(bb[k - 1]); c += cnt[k - 1];
                if (c == 0 || rightCnt[k] == 0) continue;
                const float cost = acc.area() * (float)c + rightArea[k] * (float)rightCnt[k];
                if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = k; }
            }
        }
        uint32_t mid;
        if (bestAxis >= 0) {
            const float scale = (float)kBins / ext[bestAxis];
            auto left = [&](uint32_t p) {
                const AABB& b = pb[p];
                const float c = 0.5f * ((&b.bMin.x)[bestAxis] + (&b.bMax.x)[bestAxis]);
                return std::min(kBins - 1, (int)((c - mn[bestAxis]) * scale)) < bestSplit;
            };
            mid = (uint32_t)(std::partition(idx + lo, idx + hi, left) - idx);
        } else {
            // binning cannot separate the centroids: object median along the longest box axis, ties by primitive id
            const float bx[3] = {box.bMax.x - box.bMin.x, box.bMax.y - box.bMin.y, box.bMax.z - box.bMin.z};
            const int a = bx[0] >= bx[1] && bx[0] >= bx[2] ? 0 : (bx[1] >= bx[2] ? 1 : 2);
            mid = lo + n / 2;
            std::nth_element(idx + lo, idx + mid, idx + hi, [&](uint32_t p, uint32_t q) {
                const float cp = (&pb[p].bMin.x)[a] + (&pb[p].bMax.x)[a], cq = (&pb[q].bMin.x)[a] + (&pb[q].bMax.x)[a];
                return cp < cq || (cp == cq && p < q);
            });
        }
        const uint32_t nL = mid - lo;
        nd.left = base + 1; nd.right = base + 2 * nL;
        if (depth < maxParDepth && n > 16384) {
            std::thread t([&] { build(base + 1, lo, mid, depth + 1); });
            build(base + 2 * nL, mid, hi, depth + 1);
            t.join();
        } else {
            build(base + 1, lo, mid, depth + 1);
            build(base + 2 * nL, mid, hi, depth + 1);
        }
    }
};

// ----------------------------------------------------------------------------- SAH-optimal collapse (BVH8Builder) ----
enum Decision : int8_t { UNDEFINED = -1, LEAF = 0, INTERNAL = 1, DISTRIBUTE = 2 };   // BVH8Builder.h:18-24
struct Eval { float cost; Decision decision = UNDEFINED; int8_t leftCount = 0, rightCount = 0; };

struct Collapse {
    const Node2* n2; uint32_t nodeCount2, primCount;
    std::vector<Eval> evals;       // [node][7]: C(n, i + 1) of the paper
    std::vector<uint32_t> triCount;
    Node8* out; uint32_t* primIdx; uint32_t usedNodes = 0, usedIndices = 0;

    Eval& ev(uint32_t n, int i) { return evals[(size_t)n * 7 + i]; }

    uint32_t countTris(uint32_t n)                               // ComputeNodeTriCount, BVH8Builder.cpp:147-163
    {
        const Node2& nd = n2[n];
        triCount[n] = nd.left == INVALID ? 1u : countTris(nd.left) + countTris(nd.right);
        return triCount[n];
    }
    float cLeaf(const Node2& nd, uint32_t tris) { return tris > (uint32_t)kPMax ? 1.0e30f : nd.bounds.areaHost() * (float)tris * kCPrim; }   // :31-37
    float cDistribute(const Node2& nd, int j, int8_t& l, int8_t& r)                                                            // :39-57
    {
        float best = 1.0e30f;
        for (int k = 0; k < j; k++) {
            const float c = cost(nd.left, k) + cost(nd.right, j - 1 - k);
            if (c < best) { best = c; l = (int8_t)k; r = (int8_t)(j - 1 - k); }
        }
        return best;
    }
    float cost(uint32_t n, int i)                                // ComputeNodeCost, BVH8Builder.cpp:64-145
    {
        Eval& e = ev(n, i);
        if (e.decision != UNDEFINED) return e.cost;
        const Node2& nd = n2[n];
        if (nd.left == INVALID) { e.decision = LEAF; e.cost = cLeaf(nd, 1); return e.cost; }
        if (i == 0) {
            int8_t l = 0, r = 0;
            const float leaf = cLeaf(nd, triCount[n]);
            const float internal = cDistribute(nd, 7, l, r) + nd.bounds.areaHost() * kCNode;      // CInternal, :59-62
            Eval& e0 = ev(n, 0);
            if (leaf < internal) { e0.decision = LEAF; e0.cost = leaf; }
            else { e0.decision = INTERNAL; e0.cost = internal; e0.leftCount = l; e0.rightCount = r; }
            return e0.cost;
        }
        int8_t l = 0, r = 0;
        const float dist = cDistribute(nd, i, l, r);
        const float fewer = cost(n, i - 1);
        Eval& ei = ev(n, i);
        if (dist < fewer) { ei.decision = DISTRIBUTE; ei.cost = dist; ei.leftCount = l; ei.rightCount = r; }
        else ei = ev(n, i - 1);
        return ei.cost;
    }
    void children(uint32_t n, int i, int* idx, int& cnt)         // GetChildrenIndices, BVH8Builder.cpp:165-199
    {
        const Eval& e = ev(n, i);
        if (e.decision == LEAF) { idx[cnt++] = (int)n; return; }
        const Node2& nd = n2[n];
        if (ev(nd.left, e.leftCount).decision == DISTRIBUTE) children(nd.left, e.leftCount, idx, cnt); else idx[cnt++] = (int)nd.left;
        if (ev(nd.right, e.rightCount).decision == DISTRIBUTE) children(nd.right, e.rightCount, idx, cnt); else idx[cnt++] = (int)nd.right;
    }
    void order(uint32_t n, int* idx)                             // OrderChildren, BVH8Builder.cpp:201-280: greedy global minimum
    {
        const f3 pc = (n2[n].bounds.bMax + n2[n].bounds.bMin) * 0.5f;
        float c[8][8]; int count = 0;
        for (int k = 0; k < 8 && idx[k] != -1; k++, count++) {
            const f3 cc = (n2[idx[k]].bounds.bMin + n2[idx[k]].bounds.bMax) * 0.5f;
            for (int s = 0; s < 8; s++) c[k][s] = dot(cc - pc, mk((s & 4) ? -1.0f : 1.0f, (s & 2) ? -1.0f : 1.0f, (s & 1) ? -1.0f : 1.0f));
        }
        bool used[8] = {}; int slotOf[8]; std::fill(slotOf, slotOf + 8, -1);
        while (true) {
            float best = FLT_MAX; int bn = -1, bs = -1;
            for (int k = 0; k < count; k++) { if (slotOf[k] != -1) continue; for (int s = 0; s < 8; s++) if (!used[s] && c[k][s] < best) { best = c[k][s]; bn = k; bs = s; } }
            if (bn < 0) break;
            slotOf[bn] = bs; used[bs] = true;
        }
        int copy[8]; std::memcpy(copy, idx, sizeof(copy));
        std::fill(idx, idx + 8, -1);
        for (int k = 0; k < count; k++) idx[slotOf[k]] = copy[k];
    }
    uint32_t emitTris(uint32_t n)                                // CountTriangles, BVH8Builder.cpp:282-293
    {
        const Node2& nd = n2[n];
        if (nd.left == INVALID) { primIdx[usedIndices++] = nd.right; return 1; }
        return emitTris(nd.left) + emitTris(nd.right);
    }
    void collapse(uint32_t n, uint32_t self)                     // CollapseNode, BVH8Builder.cpp:296-390
    {
        const Node2& nd = n2[n];
        Node8& o = out[self];
        std::memset(&o, 0, sizeof(o));
        const float denom = 1.0f / 255.0f;
        const float extent[3] = {nd.bounds.bMax.x - nd.bounds.bMin.x, nd.bounds.bMax.y - nd.bounds.bMin.y, nd.bounds.bMax.z - nd.bounds.bMin.z};
        float scale[3];
        for (int a = 0; a < 3; a++) {
            // e = ceil(log2(extent / 255)); a flat axis (log2(0) = -inf in the reference) gets the smallest normal cell instead
            const float e = extent[a] > 0.0f ? std::max(-126.0f, ceilf(log2f(extent[a] * denom))) : -126.0f;
            const float cell = exp2f(e);
            o.e[a] = (uint8_t)(f2u(cell) >> 23);
            scale[a] = 1.0f / cell;
        }
        o.childBaseIdx = usedNodes; o.primBaseIdx = usedIndices; o.p = nd.bounds.bMin; o.imask = 0;
        int idx[8]; std::fill(idx, idx + 8, -1); int cnt = 0;
        if (nd.left == INVALID) idx[cnt++] = (int)n;             // single-primitive tree: the root box is its own leaf child
        else children(n, 0, idx, cnt);
        order(n, idx);
        auto q = [](float v) { return (uint8_t)std::min(255.0f, std::max(0.0f, v)); };
        uint32_t trisTotal = 0;
        for (int i = 0; i < 8; i++) {
            if (idx[i] == -1) { o.meta[i] = 0; continue; }
            const Node2& c = n2[idx[i]];
            o.qlox[i] = q(floorf((c.bounds.bMin.x - o.p.x) * scale[0])); o.qloy[i] = q(floorf((c.bounds.bMin.y - o.p.y) * scale[1])); o.qloz[i] = q(floorf((c.bounds.bMin.z - o.p.z) * scale[2]));
            o.qhix[i] = q(ceilf((c.bounds.bMax.x - o.p.x) * scale[0])); o.qhiy[i] = q(ceilf((c.bounds.bMax.y - o.p.y) * scale[1])); o.qhiz[i] = q(ceilf((c.bounds.bMax.z - o.p.z) * scale[2]));
            if (ev((uint32_t)idx[i], 0).decision == INTERNAL && c.left != INVALID && (uint32_t)idx[i] != n) {
                usedNodes++;
                o.meta[i] = (uint8_t)(0x20 | (24 + i));
                o.imask |= (uint8_t)(1u << i);
            } else {
                const uint32_t t = emitTris((uint32_t)idx[i]);   // LEAF decision: at most P_MAX triangles
                o.meta[i] = (uint8_t)((((1u << t) - 1u) << 5) | trisTotal);
                trisTotal += t;
            }
        }
        const uint32_t childBase = o.childBaseIdx; uint32_t k = 0;
        for (int i = 0; i < 8; i++)
            if (idx[i] != -1 && (out[self].imask & (1u << i))) collapse((uint32_t)idx[i], childBase + k++);
    }
};

} // namespace

extern "C" {

// bounds: n x {min xyz, max xyz}.  nodes: (2n-1) x 32 B NXB::BVH2::Node, root at 0.  threads >= 1.
int orc_sah_build_bvh2(const float* bounds, uint32_t n, void* nodes, int threads)
{
    if (!n) return -1;
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    int depth = 0; while ((1 << depth) < threads) depth++;
    Sah2 b{reinterpret_cast<const AABB*>(bounds), reinterpret_cast<Node2*>(nodes), idx.data(), depth};
    b.build(0, 0, n, 0);
    return 0;
}

// Leaf-SAH of a BVH2 with its root at node 0, relative to the root area: sum of inner areas * C_NODE + leaf areas * C_PRIM.
double orc_sah_bvh2_cost(const void* nodes, uint32_t n)
{
    const Node2* nd = reinterpret_cast<const Node2*>(nodes);
    double c = 0;
    for (uint32_t i = 0; i < 2 * n - 1; i++) c += (double)nd[i].bounds.area() * (nd[i].left == INVALID ? kCPrim : kCNode);
    return c / (double)nd[0].bounds.area();
}

// bvh2: any numbering, root given (BVH8Builder assumes node 0, BVH8Builder.cpp:16-17; NexusBVH puts it at 2n-2).
// out: capacity (4n-1)/7 + 1 nodes; primIdx: n.  rootCost = C(root, 1).
int orc_sah_collapse(const void* bvh2, uint32_t n, uint32_t root, void* outNodes, uint32_t* primIdx, uint32_t* outNodeCount, float* rootCost)
{
    if (!n) return -1;
    Collapse c;
    c.n2 = reinterpret_cast<const Node2*>(bvh2); c.nodeCount2 = 2 * n - 1; c.primCount = n;
    c.evals.assign((size_t)c.nodeCount2 * 7, Eval{});
    c.triCount.assign(c.nodeCount2, 0);
    c.out = reinterpret_cast<Node8*>(outNodes); c.primIdx = primIdx;
    c.countTris(root);
    const float rc = c.cost(root, 0);
    c.usedNodes = 1;
    c.collapse(root, 0);
    if (outNodeCount) *outNodeCount = c.usedNodes;
    if (rootCost) *rootCost = rc;
    return c.usedIndices == n ? 0 : -2;
}

} // extern "C"
