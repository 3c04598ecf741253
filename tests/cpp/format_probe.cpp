// tests/cpp/format_probe.cpp — TEST INFRASTRUCTURE.  Prints what every format utility of the DirectXTex API answers for every
// DXGI_FORMAT value 0..200 (and a few heights for ComputeScanlines).  Built twice by tests/test_cpu_abi.py: against the reference's
// DirectXTex.h + oracle/_ref/libdxtex_ref.so (-DPROBE_REFERENCE) and against DirectXTexB200.h + libdxtex_b200.so; the outputs must be equal.
#ifdef PROBE_REFERENCE
#include "DirectXTex.h"
#else
#include "DirectXTexB200.h"
#endif
#include <cstdio>
using namespace DirectX;
int main()
{
    for (unsigned v = 0; v <= 200; ++v)
    {
        const DXGI_FORMAT f = static_cast<DXGI_FORMAT>(v);
        printf("%u valid %d bc %d packed %d video %d planar %d planar12 %d pal %d ds %d srgb %d bgr %d typeless %d typelessfull %d alpha %d bpp %zu bpc %zu",
               v, (int)IsValid(f), (int)IsCompressed(f), (int)IsPacked(f), (int)IsVideo(f), (int)IsPlanar(f), (int)IsPlanar(f, true), (int)IsPalettized(f),
               (int)IsDepthStencil(f), (int)IsSRGB(f), (int)IsBGR(f), (int)IsTypeless(f), (int)IsTypeless(f, false), (int)HasAlpha(f), BitsPerPixel(f), BitsPerColor(f));
        printf(" type %u", (unsigned)FormatDataType(f));
        printf(" mk %u %u %u %u %u", (unsigned)MakeSRGB(f), (unsigned)MakeLinear(f), (unsigned)MakeTypeless(f), (unsigned)MakeTypelessUNORM(f), (unsigned)MakeTypelessFLOAT(f));
        printf(" scan %zu %zu %zu %zu\n", ComputeScanlines(f, 1), ComputeScanlines(f, 5), ComputeScanlines(f, 64), ComputeScanlines(f, 1023));
    }
    return 0;
}
