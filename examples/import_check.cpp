// Host-only check of include/nexus_b200_import.hpp (no GPU, nothing of the library is called): parses an .obj (+ .mtl) and a .hdr and
// prints what it read as JSON, for tests/test_cpp_import.py to compare with the Python readers.
//   g++ -std=c++17 -Iinclude examples/import_check.cpp -Lnexus_b200 -lnexus_b200 -o import_check && ./import_check cube.obj sky.hdr
#include <cstdio>
#include "nexus_b200_import.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: import_check file.obj file.hdr\n"); return 2; }
    try {
        const nexus::ImportedAsset a = nexus::LoadOBJ(argv[1]);
        std::printf("{\"materials\": [");
        for (size_t i = 0; i < a.materials.size(); i++) {
            const nexus::Material& m = a.materials[i];
            std::printf("%s{\"baseColor\": [%.9g, %.9g, %.9g], \"emissionColor\": [%.9g, %.9g, %.9g], \"intensity\": %.9g, \"ior\": %.9g, \"opacity\": %.9g, \"roughness\": %.9g, \"metalness\": %.9g}",
                        i ? ", " : "", m.baseColor.x, m.baseColor.y, m.baseColor.z, m.emissionColor.x, m.emissionColor.y, m.emissionColor.z, m.intensity, m.ior, m.opacity, m.roughness, m.metalness);
        }
        std::printf("], \"meshes\": [");
        for (size_t i = 0; i < a.meshes.size(); i++) {
            const nexus::ImportedMesh& m = a.meshes[i];
            std::printf("%s{\"name\": \"%s\", \"material\": %u, \"triangles\": [", i ? ", " : "", m.name.c_str(), m.material);
            const float* t = reinterpret_cast<const float*>(m.triangles.data());
            for (size_t k = 0; k < 9 * m.triangles.size(); k++) std::printf("%s%.9g", k ? ", " : "", t[k]);
            std::printf("], \"triangle_data\": [");
            const float* d = reinterpret_cast<const float*>(m.triangleData.data());
            for (size_t k = 0; k < 24 * m.triangleData.size(); k++) std::printf("%s%.9g", k ? ", " : "", d[k]);
            std::printf("]}");
        }
        if (argc > 3) {      // optional .glb: appended as "glb": {...} below
            const nexus::ImportedScene g = nexus::LoadGLB(argv[3]);
            std::printf("], \"glb\": {\"camera\": %s, \"camera_position\": [%.9g, %.9g, %.9g], \"camera_forward\": [%.9g, %.9g, %.9g], \"camera_hfov\": %.9g, \"materials\": [", g.hasCamera ? "true" : "false",
                        g.camera.position.x, g.camera.position.y, g.camera.position.z, g.camera.forward.x, g.camera.forward.y, g.camera.forward.z, g.camera.horizontalFOV);
            for (size_t i = 0; i < g.materials.size(); i++) {
                const nexus::Material& m = g.materials[i];
                std::printf("%s[%.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g, %.9g]", i ? ", " : "", m.baseColor.x, m.baseColor.y, m.baseColor.z, m.opacity, m.metalness,
                            m.roughness, m.emissionColor.x, m.emissionColor.y, m.emissionColor.z, m.intensity, m.specularWeight, m.specularColor.x, m.ior, m.transmission, m.specularColor.z);
            }
            std::printf("], \"maps\": [");
            for (size_t i = 0; i < g.materials.size(); i++) {
                const nexus::Material& m = g.materials[i];
                std::printf("%s[%d, %d, %d, %d]", i ? ", " : "", m.baseColorMapId, m.emissiveMapId, m.normalMapId, m.metallicRoughnessMapId);
            }
            std::printf("], \"textures\": [");
            for (size_t i = 0; i < g.textures.size(); i++) {
                const nexus::ImportedTexture& t = g.textures[i];
                std::printf("%s{\"width\": %u, \"height\": %u, \"srgb\": %s, \"rgba\": [", i ? ", " : "", t.image.width, t.image.height, t.sRGB ? "true" : "false");
                for (size_t k = 0; k < t.image.rgba.size(); k++) std::printf("%s%u", k ? ", " : "", (unsigned)t.image.rgba[k]);
                std::printf("]}");
            }
            std::printf("], \"instances\": [");
            for (size_t i = 0; i < g.instances.size(); i++) {
                std::printf("%s{\"mesh\": %u, \"matrix\": [", i ? ", " : "", g.instances[i].mesh);
                for (int k = 0; k < 16; k++) std::printf("%s%.9g", k ? ", " : "", g.instances[i].matrix[k]);
                std::printf("]}");
            }
            std::printf("], \"meshes\": [");
            for (size_t i = 0; i < g.meshes.size(); i++) {
                const nexus::ImportedMesh& m = g.meshes[i];
                std::printf("%s{\"name\": \"%s\", \"material\": %u, \"triangles\": [", i ? ", " : "", m.name.c_str(), m.material);
                const float* t = reinterpret_cast<const float*>(m.triangles.data());
                for (size_t k = 0; k < 9 * m.triangles.size(); k++) std::printf("%s%.9g", k ? ", " : "", t[k]);
                std::printf("], \"triangle_data\": [");
                const float* d = reinterpret_cast<const float*>(m.triangleData.data());
                for (size_t k = 0; k < 24 * m.triangleData.size(); k++) std::printf("%s%.9g", k ? ", " : "", d[k]);
                std::printf("]}");
            }
            std::printf("]}, \"end_of_glb\": [");
        }
        const nexus::HdrImage img = nexus::LoadHDR(argv[2]);
        std::printf("], \"hdr\": {\"width\": %u, \"height\": %u, \"rgba\": [", img.width, img.height);
        for (size_t k = 0; k < img.rgba.size(); k++) std::printf("%s%.9g", k ? ", " : "", img.rgba[k]);
        std::printf("]}}\n");
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
