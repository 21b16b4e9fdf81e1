// Asset readers for the C++ host layer (header-only, standard library only): Wavefront OBJ + MTL and Radiance .hdr, the two formats the
// Python mirror also reads (nexus_b200/obj.py, nexus_b200/hdr.py - same rules, same results).  The reference imports through Assimp and
// stb_image (src/Assets/OBJLoader.cpp:420-446, src/Assets/IMGLoader.cpp:13-31); an application that keeps those libraries feeds
// AssetManager::AddMesh / AddTexture / Scene::AddHDRMap directly and does not need this file.
//
//   nexus::ImportedAsset a = nexus::LoadOBJ("cube.obj");           // one mesh per material group, fan triangulation, V flip
//   nexus::CreateMeshInstanceFromFile(scene, "assets/", "cube.obj"); // Scene::CreateMeshInstanceFromFile (Scene.cpp:97-100)
//   nexus::AddHDRMap(scene, "assets/", "sky.hdr");                  // Scene::AddHDRMap (Scene.cpp:102-107)
#ifndef NEXUS_B200_IMPORT_HPP
#define NEXUS_B200_IMPORT_HPP

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "nexus_b200.hpp"

namespace nexus {

struct ImportedMesh { std::string name; uint32_t material = 0; std::vector<NXB::Triangle> triangles; std::vector<nx_triangle_data> triangleData; };
struct ImportedAsset { std::vector<Material> materials; std::vector<ImportedMesh> meshes; };
struct HdrImage { uint32_t width = 0, height = 0; std::vector<float> rgba; };   // rows top to bottom, alpha 1

namespace detail {
inline std::vector<std::string> Tokens(const std::string& line)
{
    std::vector<std::string> t; std::istringstream in(line.substr(0, line.find('#'))); std::string w;
    while (in >> w) t.push_back(w);
    return t;
}
inline float ToFloat(const std::string& s, const std::string& where)
{
    char* end = nullptr; const float v = std::strtof(s.c_str(), &end);
    if (end == s.c_str() || *end) throw Error(where + ": malformed number '" + s + "'");
    return v;
}
inline std::string Dir(const std::string& path) { const size_t p = path.find_last_of("/\\"); return p == std::string::npos ? "" : path.substr(0, p + 1); }
}  // namespace detail

// MTL: Kd -> baseColor, Ke -> emissionColor (intensity 1), Ni -> ior, d / Tr -> opacity, Pr -> roughness, Pm -> metalness.
inline std::map<std::string, Material> LoadMTL(const std::string& path, bool ignoreMaps = false)
{
    std::ifstream f(path);
    if (!f) throw Error("cannot open " + path);
    std::map<std::string, Material> mats; Material* cur = nullptr; std::string line; int ln = 0;
    while (std::getline(f, line)) {
        ln++;
        const auto t = detail::Tokens(line);
        if (t.empty()) continue;
        const std::string where = path + ":" + std::to_string(ln);
        if (t[0] == "newmtl") { std::string name; for (size_t i = 1; i < t.size(); i++) name += (i > 1 ? " " : "") + t[i]; cur = &mats[name]; *cur = Material(); continue; }
        if (!cur) throw Error(where + ": statement before the first newmtl");
        auto f3 = [&](float3& out) { if (t.size() < 4) throw Error(where + ": malformed '" + t[0] + "' statement"); out = {detail::ToFloat(t[1], where), detail::ToFloat(t[2], where), detail::ToFloat(t[3], where)}; };
        auto f1 = [&]() { if (t.size() < 2) throw Error(where + ": malformed '" + t[0] + "' statement"); return detail::ToFloat(t[1], where); };
        if (t[0] == "Kd") f3(cur->baseColor);
        else if (t[0] == "Ke") { float3 e; f3(e); if (e.x != 0 || e.y != 0 || e.z != 0) { cur->emissionColor = e; cur->intensity = 1.0f; } }
        else if (t[0] == "Ni") cur->ior = f1();
        else if (t[0] == "d") cur->opacity = f1();
        else if (t[0] == "Tr") cur->opacity = 1.0f - f1();
        else if (t[0] == "Pr") cur->roughness = f1();
        else if (t[0] == "Pm") cur->metalness = f1();
        else if ((t[0].rfind("map_", 0) == 0 || t[0] == "bump" || t[0] == "disp" || t[0] == "norm") && !ignoreMaps)
            throw Error(where + ": texture maps need an image decoder (pass ignoreMaps = true to drop them)");
    }
    return mats;
}

inline ImportedAsset LoadOBJ(const std::string& path, bool ignoreMaps = false)
{
    std::ifstream f(path);
    if (!f) throw Error("cannot open " + path);
    struct Corner { int v, t, n; };
    std::vector<float3> V, VN; std::vector<std::array<float, 2>> VT;
    std::vector<std::string> order; std::map<std::string, std::vector<std::array<Corner, 3>>> groups;
    std::map<std::string, Material> mtl; std::string current = "\x01none"; std::string line; int ln = 0;
    while (std::getline(f, line)) {
        ln++;
        const auto t = detail::Tokens(line);
        if (t.empty()) continue;
        const std::string where = path + ":" + std::to_string(ln);
        auto index = [&](const std::string& tok, size_t count, const char* what) {
            if (tok.empty()) return -1;
            char* end = nullptr; const long i = std::strtol(tok.c_str(), &end, 10);
            if (end == tok.c_str() || *end) throw Error(where + ": malformed 'f' statement");
            const long j = i > 0 ? i - 1 : (long)count + i;
            if (i == 0 || j < 0 || j >= (long)count) throw Error(where + ": " + what + " index " + std::to_string(i) + " out of range");
            return (int)j;
        };
        if (t[0] == "v" || t[0] == "vn") {
            if (t.size() < 4) throw Error(where + ": malformed '" + t[0] + "' statement");
            (t[0] == "v" ? V : VN).push_back({detail::ToFloat(t[1], where), detail::ToFloat(t[2], where), detail::ToFloat(t[3], where)});
        } else if (t[0] == "vt") {
            if (t.size() < 2) throw Error(where + ": malformed 'vt' statement");
            VT.push_back({detail::ToFloat(t[1], where), t.size() > 2 ? detail::ToFloat(t[2], where) : 0.0f});
        } else if (t[0] == "f") {
            if (t.size() < 4) throw Error(where + ": a face needs at least three vertices");
            std::vector<Corner> c;
            for (size_t k = 1; k < t.size(); k++) {
                std::string p[3]; size_t a = 0; int part = 0;
                for (size_t i = 0; i <= t[k].size() && part < 3; i++) if (i == t[k].size() || t[k][i] == '/') { p[part++] = t[k].substr(a, i - a); a = i + 1; }
                c.push_back({index(p[0], V.size(), "vertex"), index(p[1], VT.size(), "texture"), index(p[2], VN.size(), "normal")});
                if (c.back().v < 0) throw Error(where + ": malformed 'f' statement");
            }
            if (!groups.count(current)) order.push_back(current);
            for (size_t k = 1; k + 1 < c.size(); k++) groups[current].push_back({c[0], c[k], c[k + 1]});     // fan, like aiProcess_Triangulate
        } else if (t[0] == "usemtl") { current.clear(); for (size_t i = 1; i < t.size(); i++) current += (i > 1 ? " " : "") + t[i]; }
        else if (t[0] == "mtllib") for (size_t i = 1; i < t.size(); i++) for (auto& kv : LoadMTL(detail::Dir(path) + t[i], ignoreMaps)) mtl[kv.first] = kv.second;
    }
    if (order.empty()) throw Error(path + ": the file contains no faces");
    ImportedAsset out;
    const std::string base = path.substr(detail::Dir(path).size());
    for (const std::string& name : order) {
        const bool none = name == "\x01none";
        if (!none && !mtl.count(name)) throw Error("material '" + name + "' is not defined by any mtllib");
        ImportedMesh m; m.name = base + "." + (none ? "default" : name); m.material = (uint32_t)out.materials.size();
        out.materials.push_back(none ? Material() : mtl[name]);
        for (const auto& tri : groups[name]) {
            const float3 p[3] = {V[tri[0].v], V[tri[1].v], V[tri[2].v]};
            m.triangles.push_back(NXB::Triangle{{p[0].x, p[0].y, p[0].z}, {p[1].x, p[1].y, p[1].z}, {p[2].x, p[2].y, p[2].z}});
            const float e0[3] = {p[1].x - p[0].x, p[1].y - p[0].y, p[1].z - p[0].z}, e1[3] = {p[2].x - p[0].x, p[2].y - p[0].y, p[2].z - p[0].z};
            float g[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
            const float len = std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), il = 1.0f / std::fmax(len, 1e-30f);
            for (float& x : g) x *= il;
            nx_triangle_data d; std::memset(&d, 0, sizeof(d));
            const bool hasN = tri[0].n >= 0 && tri[1].n >= 0 && tri[2].n >= 0, hasT = tri[0].t >= 0 && tri[1].t >= 0 && tri[2].t >= 0;
            float* nrm[3] = {d.normal0, d.normal1, d.normal2}; float* uv[3] = {d.uv0, d.uv1, d.uv2};
            for (int k = 0; k < 3; k++) {
                if (hasN) { nrm[k][0] = VN[tri[k].n].x; nrm[k][1] = VN[tri[k].n].y; nrm[k][2] = VN[tri[k].n].z; } else std::memcpy(nrm[k], g, 12);
                if (hasT) { uv[k][0] = VT[tri[k].t][0]; uv[k][1] = 1.0f - VT[tri[k].t][1]; }                      // aiProcess_FlipUVs
            }
            m.triangleData.push_back(d);
        }
        out.meshes.push_back(std::move(m));
    }
    return out;
}

// Radiance RGBE: #?RADIANCE / #?RGBE, FORMAT=32-bit_rle_rgbe, -Y h +X w, flat or new-style run-length encoded scanlines; stb_image's decoding.
inline HdrImage LoadHDR(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    const std::string blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = blob.find('\n');
    auto strip = [](std::string s) { while (!s.empty() && (s.back() == '\r' || s.back() == ' ')) s.pop_back(); return s; };
    if (pos == std::string::npos || (strip(blob.substr(0, pos)) != "#?RADIANCE" && strip(blob.substr(0, pos)) != "#?RGBE")) throw Error(path + ": not a Radiance HDR file");
    bool fmt = false;
    while (true) {
        const size_t end = blob.find('\n', pos + 1);
        if (end == std::string::npos) throw Error(path + ": truncated header");
        const std::string line = strip(blob.substr(pos + 1, end - pos - 1));
        pos = end;
        if (line.empty()) break;
        if (line == "FORMAT=32-bit_rle_rgbe") fmt = true;
    }
    if (!fmt) throw Error(path + ": unsupported format (only 32-bit_rle_rgbe)");
    const size_t end = blob.find('\n', pos + 1);
    int h = 0, w = 0; char sy[8] = {0}, sx[8] = {0};
    if (end == std::string::npos || std::sscanf(blob.substr(pos + 1, end - pos - 1).c_str(), "%7s %d %7s %d", sy, &h, sx, &w) != 4 || std::strcmp(sy, "-Y") || std::strcmp(sx, "+X") || h <= 0 || w <= 0)
        throw Error(path + ": unsupported orientation (only -Y h +X w)");
    const unsigned char* data = (const unsigned char*)blob.data() + end + 1;
    const size_t len = blob.size() - end - 1;
    std::vector<unsigned char> rgbe((size_t)4 * w * h);
    const bool flat = w < 8 || w >= 32768 || (len == (size_t)4 * w * h && !(data[0] == 2 && data[1] == 2 && !(data[2] & 0x80)));
    if (flat) {
        if (len < rgbe.size()) throw Error(path + ": truncated pixel data");
        std::memcpy(rgbe.data(), data, rgbe.size());
    } else {
        size_t p = 0;
        for (int y = 0; y < h; y++) {
            if (p + 4 > len || data[p] != 2 || data[p + 1] != 2 || ((data[p + 2] << 8) | data[p + 3]) != w) throw Error(path + ": bad run-length header in scanline " + std::to_string(y));
            p += 4;
            for (int c = 0; c < 4; c++) for (int x = 0; x < w;) {
                if (p >= len) throw Error(path + ": truncated pixel data");
                int count = data[p++];
                if (count > 128) {
                    count -= 128;
                    if (x + count > w || p >= len) throw Error(path + ": run overflows scanline " + std::to_string(y));
                    for (int i = 0; i < count; i++) rgbe[4 * ((size_t)y * w + x + i) + c] = data[p];
                    p++;
                } else {
                    if (count == 0 || x + count > w || p + count > len) throw Error(path + ": dump overflows scanline " + std::to_string(y));
                    for (int i = 0; i < count; i++) rgbe[4 * ((size_t)y * w + x + i) + c] = data[p + i];
                    p += count;
                }
                x += count;
            }
        }
    }
    HdrImage img; img.width = (uint32_t)w; img.height = (uint32_t)h; img.rgba.resize((size_t)4 * w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const int e = rgbe[4 * i + 3];
        const float s = e ? std::ldexp(1.0f, e - 136) : 0.0f;
        for (int c = 0; c < 3; c++) img.rgba[4 * i + c] = (float)rgbe[4 * i + c] * s;
        img.rgba[4 * i + 3] = 1.0f;
    }
    return img;
}

// Scene::CreateMeshInstanceFromFile (Scene.cpp:97-100): the asset's materials and meshes are appended to the scene's asset manager
// and one identity instance per mesh is created.  Returns the indices of the new instances.
inline std::vector<uint32_t> CreateMeshInstanceFromFile(Scene& scene, const std::string& filePath, const std::string& fileName, bool ignoreMaps = false)
{
    const std::string path = filePath + fileName;
    if (path.size() < 4 || path.substr(path.size() - 4) != ".obj") throw Error("CreateMeshInstanceFromFile: unsupported asset type (only .obj in the C++ layer): " + path);
    const ImportedAsset a = LoadOBJ(path, ignoreMaps);
    AssetManager& am = scene.GetAssetManager();
    const uint32_t mat0 = (uint32_t)am.GetMaterials().size();
    for (const Material& m : a.materials) am.AddMaterial(m);
    std::vector<uint32_t> created;
    for (const ImportedMesh& m : a.meshes) {
        MeshInstance& inst = scene.CreateMeshInstance(am.AddMesh(m.name, mat0 + m.material, m.triangles, m.triangleData));
        inst.name = m.name;
        created.push_back(inst.index());
    }
    return created;
}

// Scene::AddHDRMap(filePath, fileName) (Scene.cpp:102-107)
inline void AddHDRMap(Scene& scene, const std::string& filePath, const std::string& fileName)
{
    const HdrImage img = LoadHDR(filePath + fileName);
    scene.AddHDRMap(img.rgba.data(), img.width, img.height);
}

}  // namespace nexus
#endif /* NEXUS_B200_IMPORT_HPP */
