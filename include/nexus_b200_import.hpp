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
#include <functional>
#include <map>
#include <sstream>

#include "nexus_b200.hpp"
#include "nexus_b200_image.hpp"

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

// ----------------------------------------------------------------------------------------------------- binary glTF ----
// Minimal .glb reader with the same coverage and rules as nexus_b200/gltf.py: triangle primitives (indexed or not; POSITION, NORMAL,
// TANGENT, TEXCOORD_0, strided views, normalised integers), node hierarchy (matrix or translation / rotation / scale) accumulated into
// one instance matrix per node and primitive, pbrMetallicRoughness + emissive + KHR_materials_{emissive_strength, specular, ior,
// transmission} materials (OBJLoader.cpp:96-119), the first perspective camera.  Embedded PNG / JPEG texture images are decoded with
// nexus_b200_image.hpp (the reference: stb_image, IMGLoader.cpp:13-43): base-colour and emissive textures are flagged sRGB, normal and
// metallic-roughness textures linear, an image shared by several materials is decoded once per (image, colour space); the materials'
// *MapId fields index ImportedScene::textures.  ignoreTextures drops all maps instead.
struct ImportedInstance { uint32_t mesh = 0; float matrix[16]; };                       // row-major object -> world
struct ImportedTexture { DecodedImage image; bool sRGB = false; };
struct ImportedScene { std::vector<Material> materials; std::vector<ImportedMesh> meshes; std::vector<ImportedInstance> instances; std::vector<ImportedTexture> textures;
                       bool hasCamera = false; Camera camera; };

namespace detail {
struct Json {                               // just enough JSON for glTF
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false; double num = 0; std::string str; std::vector<Json> arr; std::vector<std::pair<std::string, Json>> obj;
    const Json* find(const std::string& k) const { for (const auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
    bool has(const std::string& k) const { return find(k) != nullptr; }
    const Json& at(const std::string& k) const { const Json* j = find(k); if (!j) throw Error("glTF: missing key '" + k + "'"); return *j; }
    const Json& at(size_t i) const { if (kind != Array || i >= arr.size()) throw Error("glTF: index out of range"); return arr[i]; }
    double number(const std::string& k, double def) const { const Json* j = find(k); return j && j->kind == Number ? j->num : def; }
    size_t size() const { return arr.size(); }
};
struct JsonParser {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    [[noreturn]] void fail(const char* what) { throw Error(std::string("glTF: malformed JSON (") + what + ")"); }
    Json value()
    {
        ws(); if (p >= end) fail("unexpected end");
        Json j;
        if (*p == '{') { p++; j.kind = Json::Object; ws(); if (p < end && *p == '}') { p++; return j; }
            while (true) { ws(); if (p >= end || *p != '"') fail("key"); std::string k = string(); ws(); if (p >= end || *p != ':') fail("colon"); p++; j.obj.emplace_back(std::move(k), value()); ws();
                if (p < end && *p == ',') { p++; continue; } if (p < end && *p == '}') { p++; return j; } fail("object"); } }
        if (*p == '[') { p++; j.kind = Json::Array; ws(); if (p < end && *p == ']') { p++; return j; }
            while (true) { j.arr.push_back(value()); ws(); if (p < end && *p == ',') { p++; continue; } if (p < end && *p == ']') { p++; return j; } fail("array"); } }
        if (*p == '"') { j.kind = Json::String; j.str = string(); return j; }
        if (end - p >= 4 && !std::strncmp(p, "true", 4)) { p += 4; j.kind = Json::Bool; j.b = true; return j; }
        if (end - p >= 5 && !std::strncmp(p, "false", 5)) { p += 5; j.kind = Json::Bool; return j; }
        if (end - p >= 4 && !std::strncmp(p, "null", 4)) { p += 4; return j; }
        char* e = nullptr; const std::string tmp(p, std::min<size_t>(64, (size_t)(end - p))); j.num = std::strtod(tmp.c_str(), &e);
        if (e == tmp.c_str()) fail("value");
        p += e - tmp.c_str(); j.kind = Json::Number; return j;
    }
    std::string string()
    {
        std::string s; p++;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) { p++; switch (*p) { case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'u': s += '?'; p += 4; break; default: s += *p; } p++; }
            else s += *p++;
        }
        if (p >= end) fail("string");
        p++; return s;
    }
};
struct M44 { double m[16]; };
inline M44 Identity() { M44 r{}; for (int i = 0; i < 4; i++) r.m[5 * i] = 1.0; return r; }
inline M44 Mul(const M44& a, const M44& b) { M44 r{}; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) r.m[4 * i + j] += a.m[4 * i + k] * b.m[4 * k + j]; return r; }
}  // namespace detail

inline ImportedScene LoadGLB(const std::string& path, bool ignoreTextures = false)
{
    using detail::Json;
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    const std::string blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto u32 = [&](size_t off) { uint32_t v; std::memcpy(&v, blob.data() + off, 4); return v; };
    if (blob.size() < 20 || blob.compare(0, 4, "glTF") != 0) throw Error(path + ": not a binary glTF file");
    if (u32(4) != 2) throw Error(path + ": glTF version " + std::to_string(u32(4)) + " is not supported");
    Json js; bool haveJson = false; const unsigned char* bin = nullptr; size_t binLen = 0;
    for (size_t off = 12; off + 8 <= std::min<size_t>(u32(8), blob.size());) {
        const uint32_t clen = u32(off), ctype = u32(off + 4);
        if (off + 8 + clen > blob.size()) throw Error(path + ": truncated chunk");
        if (ctype == 0x4E4F534Au) { detail::JsonParser jp{blob.data() + off + 8, blob.data() + off + 8 + clen}; js = jp.value(); haveJson = true; }
        else if (ctype == 0x004E4942u) { bin = (const unsigned char*)blob.data() + off + 8; binLen = clen; }
        off += 8 + clen + ((4 - clen % 4) % 4);
    }
    if (!haveJson) throw Error(path + ": the file has no JSON chunk");

    // accessor -> rows of `width` floats
    auto accessor = [&](size_t index, size_t& width) {
        const Json& acc = js.at("accessors").at(index);
        if (acc.has("sparse")) throw Error("glTF: sparse accessors are not supported");
        const int ct = (int)acc.at("componentType").num; const std::string& ty = acc.at("type").str;
        width = ty == "SCALAR" ? 1 : ty == "VEC2" ? 2 : ty == "VEC3" ? 3 : ty == "VEC4" ? 4 : ty == "MAT4" ? 16 : 0;
        const size_t csize = ct == 5120 || ct == 5121 ? 1 : ct == 5122 || ct == 5123 ? 2 : ct == 5125 || ct == 5126 ? 4 : 0;
        if (!width || !csize) throw Error("glTF: unsupported accessor type");
        const size_t count = (size_t)acc.at("count").num;
        std::vector<double> out(count * width, 0.0);
        if (!acc.has("bufferView")) return out;
        const Json& view = js.at("bufferViews").at((size_t)acc.at("bufferView").num);
        if (view.number("buffer", 0) != 0) throw Error("glTF: only the embedded binary buffer is supported");
        const size_t start = (size_t)view.number("byteOffset", 0) + (size_t)acc.number("byteOffset", 0);
        size_t stride = (size_t)view.number("byteStride", 0); if (!stride) stride = csize * width;
        if (count && start + stride * (count - 1) + csize * width > binLen) throw Error("glTF: accessor runs past the binary chunk");
        const Json* nrm = acc.find("normalized"); const bool normalized = nrm && nrm->kind == Json::Bool && nrm->b;
        for (size_t i = 0; i < count; i++) for (size_t c = 0; c < width; c++) {
            const unsigned char* q = bin + start + stride * i + csize * c; double v = 0, maxv = 1;
            switch (ct) {
                case 5120: { int8_t x; std::memcpy(&x, q, 1); v = x; maxv = 127; break; }
                case 5121: { uint8_t x; std::memcpy(&x, q, 1); v = x; maxv = 255; break; }
                case 5122: { int16_t x; std::memcpy(&x, q, 2); v = x; maxv = 32767; break; }
                case 5123: { uint16_t x; std::memcpy(&x, q, 2); v = x; maxv = 65535; break; }
                case 5125: { uint32_t x; std::memcpy(&x, q, 4); v = x; maxv = 4294967295.0; break; }
                default: { float x; std::memcpy(&x, q, 4); v = x; }
            }
            out[i * width + c] = normalized && ct != 5126 ? (double)((float)v / (float)maxv) : v;
        }
        return out;
    };

    ImportedScene out;
    std::map<std::pair<size_t, bool>, int32_t> textureOf;
    auto textureId = [&](const Json& ref, bool srgb) -> int32_t {
        const size_t ti = (size_t)ref.at("index").num;
        const auto key = std::make_pair(ti, srgb);
        auto it = textureOf.find(key); if (it != textureOf.end()) return it->second;
        const Json& image = js.at("images").at((size_t)js.at("textures").at(ti).at("source").num);
        if (!image.has("bufferView")) throw Error(path + ": only images embedded in the binary chunk are supported");
        const Json& view = js.at("bufferViews").at((size_t)image.at("bufferView").num);
        const size_t start = (size_t)view.number("byteOffset", 0), len = (size_t)view.at("byteLength").num;
        if (start + len > binLen) throw Error(path + ": image runs past the binary chunk");
        ImportedTexture t; t.sRGB = srgb;
        try { t.image = DecodeImage(bin + start, len); } catch (const ImageError& e) { throw Error(path + ": texture " + std::to_string(ti) + ": " + e.what()); }
        out.textures.push_back(std::move(t));
        return textureOf[key] = (int32_t)out.textures.size() - 1;
    };
    auto vec = [](const Json* j, size_t n, std::initializer_list<double> def) { std::vector<double> v(def); if (j && j->kind == Json::Array) for (size_t i = 0; i < n && i < j->size(); i++) v[i] = j->arr[i].num; return v; };
    if (const Json* mats = js.find("materials")) for (const Json& m : mats->arr) {
        static const Json empty; const Json* pj = m.find("pbrMetallicRoughness"); const Json& pbr = pj ? *pj : empty; const Json* ej = m.find("extensions"); const Json& ext = ej ? *ej : empty;
        const auto base = vec(pbr.find("baseColorFactor"), 4, {1, 1, 1, 1}); const auto em = vec(m.find("emissiveFactor"), 3, {0, 0, 0});
        Material o; o.baseColor = {(float)base[0], (float)base[1], (float)base[2]}; o.opacity = (float)base[3];
        o.metalness = (float)pbr.number("metallicFactor", 1.0); o.roughness = (float)pbr.number("roughnessFactor", 1.0);
        o.emissionColor = {(float)em[0], (float)em[1], (float)em[2]}; o.intensity = std::max(em[0], std::max(em[1], em[2])) > 0.0 ? 1.0f : 0.0f;
        if (const Json* e = ext.find("KHR_materials_emissive_strength")) o.intensity = (float)e->number("emissiveStrength", 1.0);
        const Json* sp = ext.find("KHR_materials_specular");
        o.specularWeight = sp ? (float)sp->number("specularFactor", 1.0) : 1.0f;
        const auto sc = vec(sp ? sp->find("specularColorFactor") : nullptr, 3, {1, 1, 1}); o.specularColor = {(float)sc[0], (float)sc[1], (float)sc[2]};
        const Json* ior = ext.find("KHR_materials_ior"); o.ior = ior ? (float)ior->number("ior", 1.5) : 1.5f;
        const Json* tr = ext.find("KHR_materials_transmission"); o.transmission = tr ? (float)tr->number("transmissionFactor", 0.0) : 0.0f;
        if (!ignoreTextures) {
            if (const Json* t = pbr.find("baseColorTexture")) o.baseColorMapId = textureId(*t, true);
            if (const Json* t = pbr.find("metallicRoughnessTexture")) o.metallicRoughnessMapId = textureId(*t, false);
            if (const Json* t = m.find("normalTexture")) o.normalMapId = textureId(*t, false);
            if (const Json* t = m.find("emissiveTexture")) o.emissiveMapId = textureId(*t, true);
        }
        out.materials.push_back(o);
    }
    if (out.materials.empty()) out.materials.push_back(Material());

    std::map<std::pair<size_t, size_t>, uint32_t> meshOfPrim;
    auto primitive = [&](const Json& mesh, size_t meshIdx, size_t k) -> uint32_t {
        const auto key = std::make_pair(meshIdx, k);
        auto it = meshOfPrim.find(key); if (it != meshOfPrim.end()) return it->second;
        const Json& prim = mesh.at("primitives").at(k);
        if (prim.number("mode", 4) != 4) throw Error("glTF: only triangle primitives (mode 4) are supported");
        const Json& att = prim.at("attributes");
        size_t w = 0; const std::vector<double> pos = accessor((size_t)att.at("POSITION").num, w);
        const size_t nv = pos.size() / 3;
        std::vector<size_t> idx;
        if (prim.has("indices")) { size_t wi; for (double v : accessor((size_t)prim.at("indices").num, wi)) idx.push_back((size_t)v); } else for (size_t i = 0; i < nv; i++) idx.push_back(i);
        if (idx.size() % 3) throw Error("glTF: index count is not a multiple of three");
        for (size_t i : idx) if (i >= nv) throw Error("glTF: vertex index out of range");
        std::vector<double> nrm, tan, uv; size_t wn = 0, wt = 0, wu = 0;
        if (att.has("NORMAL")) nrm = accessor((size_t)att.at("NORMAL").num, wn);
        if (att.has("TANGENT")) tan = accessor((size_t)att.at("TANGENT").num, wt);
        if (att.has("TEXCOORD_0")) uv = accessor((size_t)att.at("TEXCOORD_0").num, wu);
        ImportedMesh m; const Json* nm = mesh.find("name"); m.name = (nm ? nm->str : std::string("mesh")) + "." + std::to_string(k); m.material = (uint32_t)prim.number("material", 0);
        for (size_t t = 0; t < idx.size(); t += 3) {
            float p[3][3]; for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) p[c][a] = (float)pos[3 * idx[t + c] + a];
            m.triangles.push_back(NXB::Triangle{{p[0][0], p[0][1], p[0][2]}, {p[1][0], p[1][1], p[1][2]}, {p[2][0], p[2][1], p[2][2]}});
            nx_triangle_data d; std::memset(&d, 0, sizeof(d));
            float* N[3] = {d.normal0, d.normal1, d.normal2}; float* T[3] = {d.tangent0, d.tangent1, d.tangent2}; float* U[3] = {d.uv0, d.uv1, d.uv2};
            if (!nrm.empty()) { for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) N[c][a] = (float)nrm[wn * idx[t + c] + a]; }
            else {
                const float e0[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]}, e1[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
                float g[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
                const float il = 1.0f / std::fmax(std::sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-30f);
                for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) N[c][a] = g[a] * il;
            }
            if (!tan.empty()) for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) T[c][a] = (float)tan[wt * idx[t + c] + a];
            if (!uv.empty()) for (int c = 0; c < 3; c++) for (int a = 0; a < 2; a++) U[c][a] = (float)uv[wu * idx[t + c] + a];
            m.triangleData.push_back(d);
        }
        out.meshes.push_back(std::move(m));
        return meshOfPrim[key] = (uint32_t)out.meshes.size() - 1;
    };

    auto nodeMatrix = [&](const Json& node) {
        detail::M44 r = detail::Identity();
        if (const Json* mj = node.find("matrix")) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[4 * i + j] = mj->at((size_t)(4 * j + i)).num; return r; }   // column-major in the file
        const auto t = vec(node.find("translation"), 3, {0, 0, 0}), q = vec(node.find("rotation"), 4, {0, 0, 0, 1}), s = vec(node.find("scale"), 3, {1, 1, 1});
        const double x = q[0], y = q[1], z = q[2], w = q[3];
        const double rot[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                               2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.m[4 * i + j] = rot[3 * i + j] * s[j]; r.m[4 * i + 3] = t[i]; }
        return r;
    };
    std::function<void(size_t, const detail::M44&, int)> visit = [&](size_t ni, const detail::M44& parent, int depth) {
        if (depth > 256) throw Error("glTF: node hierarchy too deep (cycle?)");
        const Json& node = js.at("nodes").at(ni);
        const detail::M44 world = detail::Mul(parent, nodeMatrix(node));
        if (node.has("mesh")) {
            const size_t mi = (size_t)node.at("mesh").num; const Json& mesh = js.at("meshes").at(mi);
            for (size_t k = 0; k < mesh.at("primitives").size(); k++) {
                ImportedInstance inst; inst.mesh = primitive(mesh, mi, k);
                for (int i = 0; i < 16; i++) inst.matrix[i] = (float)world.m[i];
                out.instances.push_back(inst);
            }
        }
        if (node.has("camera") && !out.hasCamera) {
            const Json& cam = js.at("cameras").at((size_t)node.at("camera").num);
            const Json* ty = cam.find("type");
            if (ty && ty->str == "perspective") {
                const Json& p = cam.at("perspective");
                const double aspect = p.number("aspectRatio", 16.0 / 9.0), yfov = p.at("yfov").num;
                double fw[3] = {-world.m[2], -world.m[6], -world.m[10]}; const double l = std::sqrt(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
                out.camera.position = {(float)world.m[3], (float)world.m[7], (float)world.m[11]};
                out.camera.forward = {(float)(fw[0] / l), (float)(fw[1] / l), (float)(fw[2] / l)};
                out.camera.horizontalFOV = (float)(2.0 * std::atan(std::tan(0.5 * yfov) * aspect) * 180.0 / 3.14159265358979323846);
                out.camera.focusDistance = 5.0f; out.camera.defocusAngle = 0.0f; out.hasCamera = true;
            }
        }
        if (const Json* ch = node.find("children")) for (const Json& c : ch->arr) visit((size_t)c.num, world, depth + 1);
    };
    const Json* scenes = js.find("scenes");
    if (scenes && scenes->size()) { const Json& sc = scenes->at((size_t)js.number("scene", 0)); if (const Json* roots = sc.find("nodes")) for (const Json& r : roots->arr) visit((size_t)r.num, detail::Identity(), 0); }
    else if (const Json* nodes = js.find("nodes")) for (size_t i = 0; i < nodes->size(); i++) visit(i, detail::Identity(), 0);
    if (out.meshes.empty()) throw Error(path + ": the asset contains no triangle geometry");
    return out;
}

// Scene::CreateMeshInstanceFromFile (Scene.cpp:97-100): the asset's materials and meshes are appended to the scene's asset manager
// and one identity instance per mesh is created.  Returns the indices of the new instances.
inline std::vector<uint32_t> CreateMeshInstanceFromFile(Scene& scene, const std::string& filePath, const std::string& fileName, bool ignoreMaps = false)
{
    const std::string path = filePath + fileName;
    const std::string ext = path.size() >= 4 ? path.substr(path.size() - 4) : "";
    AssetManager& am = scene.GetAssetManager();
    const uint32_t mat0 = (uint32_t)am.GetMaterials().size();
    if (ext == ".glb") {      // one mesh per primitive, one instance per node and primitive with the accumulated node transform
        const ImportedScene g = LoadGLB(path, ignoreMaps);
        // textures first: their ids in this scene replace the asset-local ones in the materials
        std::vector<int32_t> texIds;
        for (const ImportedTexture& t : g.textures) texIds.push_back((int32_t)am.AddTexture(t.image.rgba.data(), t.image.width, t.image.height, false, t.sRGB));
        for (Material m : g.materials) {
            for (int32_t* id : {&m.baseColorMapId, &m.emissiveMapId, &m.normalMapId, &m.roughnessMapId, &m.metalnessMapId, &m.metallicRoughnessMapId}) if (*id >= 0) *id = texIds[(size_t)*id];
            am.AddMaterial(m);
        }
        std::vector<uint32_t> meshIds, created;
        for (const ImportedMesh& m : g.meshes) meshIds.push_back(am.AddMesh(m.name, mat0 + m.material, m.triangles, m.triangleData));
        for (const ImportedInstance& i : g.instances) {
            MeshInstance& inst = scene.CreateMeshInstanceMatrix(meshIds[i.mesh], i.matrix);
            inst.name = g.meshes[i.mesh].name;
            created.push_back(inst.index());
        }
        return created;
    }
    if (ext != ".obj") throw Error("CreateMeshInstanceFromFile: unsupported asset type (.obj and .glb are): " + path);
    const ImportedAsset a = LoadOBJ(path, ignoreMaps);
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
