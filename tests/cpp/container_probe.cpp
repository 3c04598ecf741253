// tests/cpp/container_probe.cpp — TEST INFRASTRUCTURE.  Exercises the ScratchImage constructors of the DirectXTex API (1D / 2D / cube /
// from images / OverrideFormat) and prints HRESULTs, metadata, the layout of every image and a checksum of the pixel memory.  Built twice
// by tests/test_cpu_abi.py (reference header + reference build, mirror header + libdxtex_b200.so); the outputs must be identical.
#ifdef PROBE_REFERENCE
#include "DirectXTex.h"
#else
#include "DirectXTexB200.h"
#endif
#include <cstdio>
#include <cstdint>
#include <vector>
using namespace DirectX;
static void dump(const char* what, HRESULT hr, const ScratchImage& s)
{
    printf("%s hr %08x", what, (unsigned)hr);
    if (hr >= 0)
    {
        const TexMetadata& m = s.GetMetadata();
        printf(" w %zu h %zu d %zu arr %zu mips %zu misc %x misc2 %x fmt %u dim %u n %zu size %zu", m.width, m.height, m.depth, m.arraySize, m.mipLevels,
               (unsigned)m.miscFlags, (unsigned)m.miscFlags2, (unsigned)m.format, (unsigned)m.dimension, s.GetImageCount(), s.GetPixelsSize());
        uint64_t sum = 1469598103934665603ull;
        for (size_t i = 0; i < s.GetPixelsSize(); ++i) sum = (sum ^ s.GetPixels()[i]) * 1099511628211ull;
        printf(" fnv %016llx\n", (unsigned long long)sum);
        for (size_t i = 0; i < s.GetImageCount(); ++i)
        {
            const Image& im = s.GetImages()[i];
            printf("   [%zu] %zux%zu fmt %u row %zu slice %zu off %zu\n", i, im.width, im.height, (unsigned)im.format, im.rowPitch, im.slicePitch, (size_t)(im.pixels - s.GetPixels()));
        }
    }
    else printf("\n");
}
int main()
{
    std::vector<uint8_t> px(6 * 40 * 24 * 4 + 64);
    for (size_t i = 0; i < px.size(); ++i) px[i] = (uint8_t)(i * 131u + (i >> 7));
    std::vector<Image> im(6);
    for (size_t i = 0; i < 6; ++i) { im[i].width = 20; im[i].height = 12; im[i].format = DXGI_FORMAT_R8G8B8A8_UNORM; im[i].rowPitch = 96; im[i].slicePitch = 96 * 12; im[i].pixels = px.data() + i * 96 * 12; }
    { ScratchImage s; dump("1d", s.Initialize1D(DXGI_FORMAT_R8_UNORM, 37, 3, 0), s); }
    { ScratchImage s; dump("1d-bad", s.Initialize1D(DXGI_FORMAT_R8_UNORM, 0, 3, 0), s); }
    { ScratchImage s; dump("2d", s.Initialize2D(DXGI_FORMAT_R16G16B16A16_FLOAT, 33, 17, 2, 0), s); }
    { ScratchImage s; dump("2d-bc", s.Initialize2D(DXGI_FORMAT_BC3_UNORM, 30, 18, 1, 3), s); }
    { ScratchImage s; dump("2d-toomany", s.Initialize2D(DXGI_FORMAT_R8G8B8A8_UNORM, 16, 16, 1, 9), s); }
    { ScratchImage s; dump("cube", s.InitializeCube(DXGI_FORMAT_R8G8B8A8_UNORM, 8, 8, 2, 2), s); }
    { ScratchImage s; dump("cube-bad", s.InitializeCube(DXGI_FORMAT_R8G8B8A8_UNORM, 8, 8, 0, 1), s); }
    { ScratchImage s; dump("from", s.InitializeFromImage(im[1]), s); }
    { Image one = im[2]; one.height = 1; one.slicePitch = 96; ScratchImage s; dump("from-1d", s.InitializeFromImage(one, true), s); ScratchImage t; dump("from-1d-as-2d", t.InitializeFromImage(one, false), t); }
    { ScratchImage s; dump("array", s.InitializeArrayFromImages(im.data(), 5), s); }
    { ScratchImage s; dump("array-null", s.InitializeArrayFromImages(nullptr, 5), s); }
    { std::vector<Image> bad(im); bad[3].width = 19; ScratchImage s; dump("array-mismatch", s.InitializeArrayFromImages(bad.data(), 5), s); }
    { std::vector<Image> bad(im); bad[4].pixels = nullptr; ScratchImage s; dump("array-nullpixels", s.InitializeArrayFromImages(bad.data(), 5), s); }
    { ScratchImage s; dump("cubeimg", s.InitializeCubeFromImages(im.data(), 6), s); }
    { ScratchImage s; dump("cubeimg-5", s.InitializeCubeFromImages(im.data(), 5), s); }
    { ScratchImage s; s.Initialize2D(DXGI_FORMAT_R8G8B8A8_UNORM, 8, 4, 1, 2); const bool a = s.OverrideFormat(DXGI_FORMAT_R8G8B8A8_UNORM_SRGB), b = s.OverrideFormat(DXGI_FORMAT_NV12), c = s.OverrideFormat(DXGI_FORMAT_UNKNOWN), d = s.OverrideFormat(DXGI_FORMAT_P8);
      printf("override %d %d %d %d\n", (int)a, (int)b, (int)c, (int)d); dump("override", 0, s); ScratchImage e; printf("override-empty %d\n", (int)e.OverrideFormat(DXGI_FORMAT_R8G8B8A8_UNORM)); }
    { TexMetadata m{}; m.width = 64; m.height = 32; m.depth = 1; m.arraySize = 6; m.mipLevels = 4; m.miscFlags = TEX_MISC_TEXTURECUBE; m.miscFlags2 = 2; m.format = DXGI_FORMAT_BC1_UNORM; m.dimension = TEX_DIMENSION_TEXTURE2D;
      printf("meta cube %d pm %d vol %d alpha %u sub %u %u %u %u idx %zu %zu\n", (int)m.IsCubemap(), (int)m.IsPMAlpha(), (int)m.IsVolumemap(), (unsigned)m.GetAlphaMode(),
             (unsigned)m.CalculateSubresource(3, 5), (unsigned)m.CalculateSubresource(4, 0), (unsigned)m.CalculateSubresource(1, 6), (unsigned)m.CalculateSubresource(2, 1, 1),
             m.ComputeIndex(3, 5, 0), m.ComputeIndex(0, 6, 0)); }
    return 0;
}
