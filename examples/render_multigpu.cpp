// Multi-GPU C++ host: one process, G B200s, sample partition + NCCL all-reduce of the float accumulation buffers over NVLink
// (include/nexus_b200_nccl.hpp).  Every GPU gets a context, a replica of the scene and a path tracer; GPU g renders frames
// [1 + g * K, 1 + (g + 1) * K); after the reduce every GPU holds the sum of all G * K frames.  The program then renders the same G * K
// frame indices on GPU 0 alone and compares: the two images may differ by float summation order only.
//
//   g++ -std=c++17 -Iinclude -I/usr/local/cuda/include examples/render_multigpu.cpp -Lnexus_b200 -lnexus_b200 -lnccl
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/nexus_b200 -o render_multigpu
//   ./render_multigpu [gpus = all] [width 640] [height 480] [frames per GPU 8] [out prefix]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include "nexus_b200_nccl.hpp"
#include "cornell_scene.hpp"

using namespace nexus;

int main(int argc, char** argv)
{
    int available = 0;
    if (cudaGetDeviceCount(&available) != cudaSuccess || available < 1) { std::fprintf(stderr, "error: no CUDA device\n"); return 1; }
    const int G = argc > 1 && std::atoi(argv[1]) > 0 ? std::min(std::atoi(argv[1]), available) : available;
    const uint32_t w = argc > 2 ? std::atoi(argv[2]) : 640, h = argc > 3 ? std::atoi(argv[3]) : 480, K = argc > 4 ? std::atoi(argv[4]) : 8;
    const std::string out = argc > 5 ? argv[5] : "";
    try {
        std::vector<std::unique_ptr<Context>> ctx; std::vector<std::unique_ptr<Scene>> scene; std::vector<std::unique_ptr<PathTracer>> pt;
        std::vector<int> devs;
        for (int g = 0; g < G; g++) {
            devs.push_back(g);
            ctx.emplace_back(new Context(g));
            scene.emplace_back(new Scene(*ctx[g], nexus::uint2{w, h}));
            cornell::Build(*scene[g]);
            pt.emplace_back(new PathTracer(*ctx[g], nexus::uint2{w, h}));
        }
        std::vector<ncclComm_t> comms((size_t)G);
        NcclCheck(ncclCommInitAll(comms.data(), G, devs.data()), "ncclCommInitAll");

        std::vector<GpuRank> ranks;
        for (int g = 0; g < G; g++) ranks.push_back(GpuRank{pt[g].get(), scene[g].get()});
        RenderPartitioned(ranks, comms, K);                     // warm-up (NCCL channels, kernels)
        for (int g = 0; g < G; g++) { ctx[g]->Synchronize(); pt[g]->ResetFrameNumber(); }
        cudaEvent_t e0, e1;
        cudaSetDevice(0); cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, (cudaStream_t)ctx[0]->stream());
        RenderPartitioned(ranks, comms, K);
        cudaSetDevice(0); cudaEventRecord(e1, (cudaStream_t)ctx[0]->stream());
        for (int g = 0; g < G; g++) ctx[g]->Synchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);

        // every GPU holds the same reduced image ...
        const std::vector<float> multi = pt[0]->ReadAccumulation();
        for (int g = 1; g < G; g++) {
            const std::vector<float> other = pt[g]->ReadAccumulation();
            for (size_t i = 0; i < multi.size(); i++) if (other[i] != multi[i]) { std::fprintf(stderr, "error: GPU %d holds a different reduced image\n", g); return 3; }
        }
        // ... and it is the image one GPU accumulates over the same frame indices
        PathTracer solo(*ctx[0], nexus::uint2{w, h});
        solo.Render(*scene[0], (uint32_t)G * K);
        const std::vector<float> single = solo.ReadAccumulation();
        double maxAbs = 0, mean = 0, peak = 0;
        for (size_t i = 0; i < multi.size(); i++) { maxAbs = std::fmax(maxAbs, std::fabs((double)multi[i] - single[i])); mean += single[i]; peak = std::fmax(peak, (double)single[i]); }
        mean /= (double)multi.size();
        std::printf("%d GPU(s), %ux%u, %u frames each: %.2f ms on GPU 0 incl. the all-reduce of %.1f MB; mean radiance %.5f; max |multi - single| = %.3g (peak %.3g)\n",
                    G, w, h, K, ms, 12e-6 * w * h, mean, maxAbs, peak);
        if (!out.empty()) WritePFM(out + ".pfm", multi, nexus::uint2{w, h});
        for (int g = 0; g < G; g++) ncclCommDestroy(comms[g]);
        if (!(mean > 0.0) || maxAbs > 1e-5 * peak * (double)(G * K) + 1e-4 * mean) { std::fprintf(stderr, "error: the reduced image differs from the single-GPU image\n"); return 2; }
        std::printf("multi gpu ok\n");
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
