// TEST INFRASTRUCTURE ONLY.  Drives the UNMODIFIED reference CPU collapse, Nexus/src/Geometry/BVH/BVH8Builder.cpp (compiled from
// /root/reference where it lies, oracle/Makefile target `refcpu`), so that the oracle's restatement (oracle_sah.cpp, orc_sah_collapse)
// can be pinned against the reference's own code.  That class is dead code in the reference snapshot, and BVH8Builder::Build() as
// written cannot run: it allocates bvh8.nodes but writes primitive ids through the never-allocated bvh8.primIdx
// (BVH8Builder.cpp:21-29, 262).  The harness therefore does what Build() does, with both arrays allocated, through the class's
// own public Init() / CollapseNode(); the one private member Build() sets first (m_UsedNodes = 1) is reached by compiling this
// translation unit with `private` made public, which changes no layout and touches no reference source.
#include <cstring>
#define private public
#include "Geometry/BVH/BVH8Builder.h"
#undef private

static_assert(sizeof(NXB::BVH2::Node) == 32 && sizeof(NXB::BVH8::Node) == 80, "reference node layouts");

// bvh2: nodeCount 32-byte nodes, root at index 0 (BVH8Builder.cpp:16-17, 26).  outNodes: room for (4 * primCount - 1) / 7 + 1 nodes.
extern "C" int ref_cpu_bvh8_collapse(const void* bvh2, uint32_t nodeCount, uint32_t primCount, void* outNodes, uint32_t* outPrimIdx,
                                      uint32_t* outNodeCount, float* outRootCost)
{
    if (!bvh2 || !nodeCount || !primCount || !outNodes || !outPrimIdx) return -1;
    NXB::BVH2 in;
    in.nodes = (NXB::BVH2::Node*)bvh2; in.nodeCount = nodeCount; in.primCount = primCount; in.bounds = in.nodes[0].bounds;
    BVH8Builder builder(in);
    builder.Init();
    NXB::BVH8 out;
    std::memset(&out, 0, sizeof(out));
    out.nodes = (NXB::BVH8::Node*)outNodes; out.primIdx = outPrimIdx;
    builder.m_UsedNodes = 1;                       // BVH8Builder::Build, BVH8Builder.cpp:23
    builder.CollapseNode(out, 0, 0);
    if (outNodeCount) *outNodeCount = builder.m_UsedNodes;
    if (outRootCost) *outRootCost = builder.m_Evals[0][0].cost;
    return builder.m_UsedIndices == primCount ? 0 : -2;
}
