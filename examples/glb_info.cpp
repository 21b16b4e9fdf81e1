// Host-only: loads a .glb with the C++ reader (include/nexus_b200_import.hpp, textures decoded by include/nexus_b200_image.hpp) and prints a
// one-line JSON summary; with a second argument the decoded textures are written as raw RGBA8 files <prefix><index>.rgba.
// tests/test_cpp_import.py runs it over the reference's demo scenes where they are mounted.
#include <cstdio>
#include "nexus_b200_import.hpp"

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: glb_info file.glb [texture dump prefix]\n"); return 2; }
    try {
        const nexus::ImportedScene g = nexus::LoadGLB(argv[1]);
        size_t tris = 0; for (const auto& m : g.meshes) tris += m.triangles.size();
        std::printf("{\"meshes\": %zu, \"instances\": %zu, \"triangles\": %zu, \"materials\": %zu, \"camera\": %s, \"textures\": [", g.meshes.size(), g.instances.size(), tris,
                    g.materials.size(), g.hasCamera ? "true" : "false");
        for (size_t i = 0; i < g.textures.size(); i++) {
            const nexus::ImportedTexture& t = g.textures[i];
            std::printf("%s{\"width\": %u, \"height\": %u, \"srgb\": %s}", i ? ", " : "", t.image.width, t.image.height, t.sRGB ? "true" : "false");
            if (argc > 2) {
                const std::string path = std::string(argv[2]) + std::to_string(i) + ".rgba";
                FILE* f = std::fopen(path.c_str(), "wb");
                if (!f) { std::fprintf(stderr, "cannot write %s\n", path.c_str()); return 2; }
                std::fwrite(t.image.rgba.data(), 1, t.image.rgba.size(), f); std::fclose(f);
            }
        }
        std::printf("]}\n");
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
