// Headless C++ host: builds the Cornell box through the reference-shaped host API (include/nexus_b200.hpp), renders N
// samples per pixel on one B200 and writes the linear accumulation buffer as PFM and EXR.
//
//   g++ -std=c++17 -Iinclude examples/render_headless.cpp -Lnexus_b200 -lnexus_b200 -Wl,-rpath,$PWD/nexus_b200 -o render_headless
//   ./render_headless 1920 1080 256 out
#include <cstdio>
#include <cstdlib>
#include <string>
#include "nexus_b200.hpp"
#include "cornell_scene.hpp"

using namespace nexus;

int main(int argc, char** argv)
{
    const uint32_t w = argc > 1 ? std::atoi(argv[1]) : 512, h = argc > 2 ? std::atoi(argv[2]) : 512, spp = argc > 3 ? std::atoi(argv[3]) : 64;
    const std::string out = argc > 4 ? argv[4] : "cornell";
    try {
        Context ctx(0);
        Scene scene(ctx, uint2{w, h});
        cornell::Build(scene);

        PathTracer pt(ctx, uint2{w, h});
        pt.Render(scene, spp);
        const nx_frame_stats st = pt.Stats();
        const std::vector<float> img = pt.ReadAccumulation();
        double mean = 0; for (float v : img) mean += v; mean /= (double)img.size();
        WritePFM(out + ".pfm", img, uint2{w, h});
        WriteEXR(out + ".exr", img, uint2{w, h});
        std::printf("%ux%u, %u spp: %.2f ms, %.1f Mrays/s, mean radiance %.5f -> %s.pfm / %s.exr\n", w, h, spp, st.device_ms,
                    (double)(st.extension_rays + st.shadow_rays) / (st.device_ms * 1e3), mean, out.c_str(), out.c_str());
        if (!(mean > 0.0)) return 2;
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
