// Host-only check of include/nexus_b200_image.hpp: decodes one PNG / JPEG file and writes "width height\n" + raw RGBA8 to a file.
// tests/test_cpp_image.py compares the pixels with Pillow's decoding of the same file.
#include <cstdio>
#include "nexus_b200_image.hpp"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: image_check in.{png,jpg} out.rgba\n"); return 2; }
    try {
        const nexus::DecodedImage img = nexus::LoadImageFile(argv[1]);
        FILE* f = std::fopen(argv[2], "wb");
        if (!f) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 2; }
        std::fprintf(f, "%u %u\n", img.width, img.height);
        std::fwrite(img.rgba.data(), 1, img.rgba.size(), f);
        std::fclose(f);
    } catch (const std::exception& e) { std::fprintf(stderr, "error: %s\n", e.what()); return 1; }
    return 0;
}
