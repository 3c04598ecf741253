// tests/cpp/pitch_probe.cpp — TEST INFRASTRUCTURE.  ComputePitch for every DXGI_FORMAT value 0..200 x 5 sizes x 11 CP_FLAGS settings, and a
// ScratchImage laid out with an alignment flag; built against the reference and against the mirror (tests/test_cpu_abi.py), outputs must match.
#ifdef PROBE_REFERENCE
#include "DirectXTex.h"
#else
#include "DirectXTexB200.h"
#endif
#include <cstdio>
using namespace DirectX;
int main()
{
    const size_t sz[5][2] = { { 1, 1 }, { 5, 7 }, { 64, 32 }, { 1023, 3 }, { 130, 258 } };
    const unsigned fl[11] = { 0, 0x1, 0x2, 0x4, 0x8, 0x200, 0x1000, 0x10000, 0x20000, 0x40000, 0x1 | 0x2 | 0x10000 };
    for (unsigned v = 0; v <= 200; ++v)
        for (auto& s : sz)
            for (unsigned f : fl)
            {
                size_t row = 123, slice = 456;
                const HRESULT hr = ComputePitch(static_cast<DXGI_FORMAT>(v), s[0], s[1], row, slice, static_cast<CP_FLAGS>(f));
                if (hr >= 0) printf("%u %zux%zu %x -> %zu %zu\n", v, s[0], s[1], f, row, slice); else printf("%u %zux%zu %x -> hr %08x\n", v, s[0], s[1], f, (unsigned)hr);
            }
    for (unsigned f : { 0x1u, 0x2u, 0x8u, 0x200u })
    {
        ScratchImage s;
        const HRESULT hr = s.Initialize2D(DXGI_FORMAT_R8G8_UNORM, 37, 11, 2, 0, static_cast<CP_FLAGS>(f));
        printf("init flags %x hr %08x n %zu size %zu\n", f, (unsigned)hr, s.GetImageCount(), s.GetPixelsSize());
        for (size_t i = 0; i < s.GetImageCount(); ++i) printf("  [%zu] %zux%zu row %zu slice %zu off %zu\n", i, s.GetImages()[i].width, s.GetImages()[i].height, s.GetImages()[i].rowPitch, s.GetImages()[i].slicePitch, (size_t)(s.GetImages()[i].pixels - s.GetPixels()));
    }
    return 0;
}
