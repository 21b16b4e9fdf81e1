// Multi-GPU sample partition for the C++ host layer: the NCCL reduce of the float accumulation buffers north_star describes, next to
// the Python form of the same thing (nexus_b200/multigpu.py).  Header-only; needs <nccl.h> and -lnccl in the APPLICATION - the library
// itself (libnexus_b200.so) does not link NCCL.
//
// Rendering shards by sample: every GPU holds the whole scene, GPU g of G renders frame indices [first + g * K, first + (g + 1) * K),
// and one all-reduce(sum) of 3 * W * H floats over NVLink leaves the sum of all G * K frames on every GPU (frames are pure functions
// of (pixel, frame index, bounce), so the result is the image one GPU would accumulate over the same indices, up to float summation
// order).  The reduce is queued on each context's render stream, behind that GPU's frames: no host synchronisation in between.
//
//   one process, G GPUs (examples/render_multigpu.cpp):   ncclCommInitAll + nexus::RenderPartitioned(group)
//   one process per GPU (MPI / torchrun style):           ncclCommInitRank  + nexus::RenderPartitioned(pt, scene, comm, rank, world, ...)
#ifndef NEXUS_B200_NCCL_HPP
#define NEXUS_B200_NCCL_HPP

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "nexus_b200.hpp"

namespace nexus {

inline void NcclCheck(ncclResult_t r, const char* what) { if (r != ncclSuccess) throw Error(std::string(what) + ": " + ncclGetErrorString(r)); }

// One rank's share: renders its block of frames into an EMPTY accumulation and all-reduces it in place.  Call between ncclGroupStart /
// ncclGroupEnd when one thread drives several ranks.  Afterwards the accumulation holds world * framesPerRank frames.
inline void RenderPartitioned(PathTracer& pt, Scene& scene, ncclComm_t comm, int rank, int world, uint32_t framesPerRank, uint32_t firstFrame = 1)
{
    if (pt.GetFrameNumber() != 0) throw Error("RenderPartitioned: the accumulation must be empty (ResetFrameNumber first): a buffer that already holds frames is the global sum on every rank and would be counted world times");
    pt.RenderFrames(scene, firstFrame + (uint32_t)rank * framesPerRank, framesPerRank);
    const uint2 res = pt.GetResolution();
    float* acc = pt.AccumulationDevice();
    NcclCheck(ncclAllReduce(acc, acc, 3ull * res.x * res.y, ncclFloat, ncclSum, comm, (cudaStream_t)pt.context().stream()), "ncclAllReduce");
    pt.SetAccumulatedFrames((uint32_t)world * framesPerRank);
}

// One process driving G GPUs: a context, a scene replica and a path tracer per GPU (built by the caller, who owns them).
struct GpuRank { PathTracer* pathTracer; Scene* scene; };
inline void RenderPartitioned(std::vector<GpuRank>& ranks, std::vector<ncclComm_t>& comms, uint32_t framesPerRank, uint32_t firstFrame = 1)
{
    const int world = (int)ranks.size();
    if ((int)comms.size() != world) throw Error("RenderPartitioned: one communicator per GPU");
    // the renders first (asynchronous), then all reduces inside one NCCL group so that a single thread cannot deadlock on them
    for (int g = 0; g < world; g++) {
        if (ranks[g].pathTracer->GetFrameNumber() != 0) throw Error("RenderPartitioned: the accumulations must be empty");
        ranks[g].pathTracer->RenderFrames(*ranks[g].scene, firstFrame + (uint32_t)g * framesPerRank, framesPerRank);
    }
    NcclCheck(ncclGroupStart(), "ncclGroupStart");
    for (int g = 0; g < world; g++) {
        PathTracer& pt = *ranks[g].pathTracer;
        const uint2 res = pt.GetResolution();
        float* acc = pt.AccumulationDevice();
        NcclCheck(ncclAllReduce(acc, acc, 3ull * res.x * res.y, ncclFloat, ncclSum, comms[g], (cudaStream_t)pt.context().stream()), "ncclAllReduce");
    }
    NcclCheck(ncclGroupEnd(), "ncclGroupEnd");
    for (int g = 0; g < world; g++) ranks[g].pathTracer->SetAccumulatedFrames((uint32_t)world * framesPerRank);
}

}  // namespace nexus
#endif /* NEXUS_B200_NCCL_HPP */
