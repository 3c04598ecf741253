// tests/cpp/abi_probe.cpp — TEST INFRASTRUCTURE.  Prints the memory layout (sizeof / offsetof) of the public types and the values of
// the public enumerators of the DirectXTex API for the hot path.  Built twice by tests/test_cpu_abi.py:
//   -DPROBE_REFERENCE : against the reference's own DirectXTex.h (through oracle/compat, where /root/reference is mounted)
//   (default)         : against directxtex_b200/host/DirectXTexB200.h
// and the two outputs must be identical line for line: a program compiled against the reference header and linked to
// libdxtex_b200.so sees the same structs and constants.
#ifdef PROBE_REFERENCE
#include "DirectXTex.h"
#else
#include "DirectXTexB200.h"
#endif
#include <cstddef>
#include <cstdio>
using namespace DirectX;

#define SZ(T) printf("sizeof(" #T ") %zu align %zu\n", sizeof(T), alignof(T))
#define OFF(T, m) printf("offsetof(" #T ", " #m ") %zu size %zu\n", offsetof(T, m), sizeof(((T*)nullptr)->m))
#define VAL(e) printf(#e " 0x%llx\n", (unsigned long long)(e))

int main()
{
    SZ(Image); OFF(Image, width); OFF(Image, height); OFF(Image, format); OFF(Image, rowPitch); OFF(Image, slicePitch); OFF(Image, pixels);
    SZ(TexMetadata); OFF(TexMetadata, width); OFF(TexMetadata, height); OFF(TexMetadata, depth); OFF(TexMetadata, arraySize); OFF(TexMetadata, mipLevels);
    OFF(TexMetadata, miscFlags); OFF(TexMetadata, miscFlags2); OFF(TexMetadata, format); OFF(TexMetadata, dimension);
    SZ(ScratchImage); SZ(Blob);
    SZ(CompressOptions); OFF(CompressOptions, flags); OFF(CompressOptions, threshold); OFF(CompressOptions, alphaWeight);
    SZ(ConvertOptions); OFF(ConvertOptions, filter); OFF(ConvertOptions, threshold);
    VAL(TEX_DIMENSION_TEXTURE1D); VAL(TEX_DIMENSION_TEXTURE2D); VAL(TEX_DIMENSION_TEXTURE3D);
    VAL(TEX_FILTER_DEFAULT); VAL(TEX_FILTER_WRAP_U); VAL(TEX_FILTER_WRAP_V); VAL(TEX_FILTER_WRAP); VAL(TEX_FILTER_MIRROR_U); VAL(TEX_FILTER_MIRROR_V); VAL(TEX_FILTER_MIRROR);
    VAL(TEX_FILTER_SEPARATE_ALPHA); VAL(TEX_FILTER_FLOAT_X2BIAS); VAL(TEX_FILTER_RGB_COPY_RED); VAL(TEX_FILTER_RGB_COPY_GREEN); VAL(TEX_FILTER_RGB_COPY_BLUE);
    VAL(TEX_FILTER_DITHER); VAL(TEX_FILTER_DITHER_DIFFUSION); VAL(TEX_FILTER_POINT); VAL(TEX_FILTER_LINEAR); VAL(TEX_FILTER_CUBIC); VAL(TEX_FILTER_BOX);
    VAL(TEX_FILTER_TRIANGLE); VAL(TEX_FILTER_SRGB_IN); VAL(TEX_FILTER_SRGB_OUT); VAL(TEX_FILTER_SRGB); VAL(TEX_FILTER_FORCE_NON_WIC);
    VAL(TEX_COMPRESS_DEFAULT); VAL(TEX_COMPRESS_RGB_DITHER); VAL(TEX_COMPRESS_A_DITHER); VAL(TEX_COMPRESS_DITHER); VAL(TEX_COMPRESS_UNIFORM);
    VAL(TEX_COMPRESS_BC7_USE_3SUBSETS); VAL(TEX_COMPRESS_BC7_QUICK); VAL(TEX_COMPRESS_SRGB_IN); VAL(TEX_COMPRESS_SRGB_OUT); VAL(TEX_COMPRESS_SRGB); VAL(TEX_COMPRESS_PARALLEL);
    VAL(TEX_PMALPHA_DEFAULT); VAL(TEX_PMALPHA_IGNORE_SRGB); VAL(TEX_PMALPHA_REVERSE); VAL(TEX_PMALPHA_SRGB_IN); VAL(TEX_PMALPHA_SRGB_OUT);
    VAL(DDS_FLAGS_NONE); VAL(DDS_FLAGS_LEGACY_DWORD); VAL(DDS_FLAGS_FORCE_DX10_EXT); VAL(DDS_FLAGS_FORCE_DX10_EXT_MISC2); VAL(DDS_FLAGS_FORCE_DX9_LEGACY);
    VAL(DDS_FLAGS_FORCE_DXT5_RXGB); VAL(DDS_FLAGS_IGNORE_MIPS); VAL(DDS_FLAGS_ALLOW_LARGE_FILES);
    VAL(TEX_ALPHA_MODE_UNKNOWN); VAL(TEX_ALPHA_MODE_STRAIGHT); VAL(TEX_ALPHA_MODE_PREMULTIPLIED); VAL(TEX_ALPHA_MODE_OPAQUE); VAL(TEX_ALPHA_MODE_CUSTOM);
    VAL(CP_FLAGS_NONE);
    VAL(DXGI_FORMAT_R32G32B32A32_FLOAT); VAL(DXGI_FORMAT_R16G16B16A16_FLOAT); VAL(DXGI_FORMAT_R8G8B8A8_UNORM); VAL(DXGI_FORMAT_R8_UNORM);
    VAL(DXGI_FORMAT_BC1_UNORM); VAL(DXGI_FORMAT_BC3_UNORM); VAL(DXGI_FORMAT_BC4_UNORM); VAL(DXGI_FORMAT_BC5_SNORM); VAL(DXGI_FORMAT_BC6H_UF16); VAL(DXGI_FORMAT_BC6H_SF16); VAL(DXGI_FORMAT_BC7_UNORM);
    printf("TEX_THRESHOLD_DEFAULT %g TEX_ALPHA_WEIGHT_DEFAULT %g\n", (double)TEX_THRESHOLD_DEFAULT, (double)TEX_ALPHA_WEIGHT_DEFAULT);
    return 0;
}
