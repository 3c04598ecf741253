// tests/cpp/dds_roundtrip.cpp — the C++ DDS mirror (SaveToDDSMemory / LoadFromDDSMemory / GetMetadataFromDDSMemory / Blob,
// directxtex_b200/host) used the way a DirectXTex caller uses it; host-side only, runs without a GPU.
//   dds_roundtrip <out.dds>   builds a 2-item, 3-level R8G8B8A8_UNORM array and a BC3 premultiplied image, saves, reloads, compares
#include "DirectXTexB200.h"
#include <cstdio>
#include <cstring>

using namespace DirectX;

static int fail(const char* what, HRESULT hr) { std::printf("FAIL %s hr=0x%08X\n", what, unsigned(hr)); return 1; }

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    ScratchImage tex;
    HRESULT hr = tex.Initialize2D(DXGI_FORMAT_R8G8B8A8_UNORM, 20, 12, 2, 3);
    if (FAILED(hr)) return fail("Initialize2D", hr);
    for (size_t i = 0; i < tex.GetPixelsSize(); ++i) tex.GetPixels()[i] = uint8_t(i * 37u + 11u);

    Blob blob;
    hr = SaveToDDSMemory(tex.GetImages(), tex.GetImageCount(), tex.GetMetadata(), DDS_FLAGS_NONE, blob);
    if (FAILED(hr)) return fail("SaveToDDSMemory", hr);
    // arrays need the DX10 extension: magic + 124-byte header + 20-byte extension + pixels
    if (blob.GetBufferSize() != 4 + 124 + 20 + tex.GetPixelsSize()) return fail("blob size", E_FAIL);

    TexMetadata md{};
    hr = GetMetadataFromDDSMemory(blob.GetConstBufferPointer(), blob.GetBufferSize(), DDS_FLAGS_NONE, md);
    if (FAILED(hr)) return fail("GetMetadataFromDDSMemory", hr);
    if (md.width != 20 || md.height != 12 || md.arraySize != 2 || md.mipLevels != 3 || md.format != DXGI_FORMAT_R8G8B8A8_UNORM) return fail("metadata", E_FAIL);

    ScratchImage back;
    hr = LoadFromDDSMemory(blob.GetConstBufferPointer(), blob.GetBufferSize(), DDS_FLAGS_NONE, nullptr, back);
    if (FAILED(hr)) return fail("LoadFromDDSMemory", hr);
    if (back.GetPixelsSize() != tex.GetPixelsSize() || std::memcmp(back.GetPixels(), tex.GetPixels(), tex.GetPixelsSize()) != 0) return fail("pixels", E_FAIL);

    // legacy header path + premultiplied alpha: BC3 with TEX_ALPHA_MODE_PREMULTIPLIED is written as 'DXT4'
    ScratchImage bc;
    hr = bc.Initialize2D(DXGI_FORMAT_BC3_UNORM, 16, 8, 1, 1);
    if (FAILED(hr)) return fail("Initialize2D bc3", hr);
    for (size_t i = 0; i < bc.GetPixelsSize(); ++i) bc.GetPixels()[i] = uint8_t(i * 13u);
    TexMetadata bm = bc.GetMetadata();
    bm.SetAlphaMode(TEX_ALPHA_MODE_PREMULTIPLIED);
    hr = SaveToDDSFile(bc.GetImages(), bc.GetImageCount(), bm, DDS_FLAGS_NONE, argv[1]);
    if (FAILED(hr)) return fail("SaveToDDSFile", hr);
    TexMetadata lm{}; ScratchImage lb;
    hr = LoadFromDDSFile(argv[1], DDS_FLAGS_NONE, &lm, lb);
    if (FAILED(hr)) return fail("LoadFromDDSFile", hr);
    if (!lm.IsPMAlpha() || lm.format != DXGI_FORMAT_BC3_UNORM || std::memcmp(lb.GetPixels(), bc.GetPixels(), bc.GetPixelsSize()) != 0) return fail("bc3 round trip", E_FAIL);

    // error paths keep the reference's codes
    hr = LoadFromDDSMemory(blob.GetConstBufferPointer(), 64, DDS_FLAGS_NONE, nullptr, back);
    if (hr != static_cast<HRESULT>(0x8007000D)) return fail("short buffer code", hr);
    std::printf("OK %zu bytes\n", blob.GetBufferSize());
    return 0;
}
