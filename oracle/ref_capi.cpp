// oracle/ref_capi.cpp — TEST INFRASTRUCTURE.  Not part of the product.
//
// extern "C" access (for ctypes: tests/, bench.py cpu_baseline / --impl reference,
// __graft_entry__.smoke()) to the UNMODIFIED reference implementation compiled from
// /root/reference/DirectXTex/*.cpp by oracle/Makefile into oracle/_ref/libdxtex_ref.so.
// Every function forwards to the reference's own public API (DirectXTex.h:818-968,
// 1041) or to its per-block codec entry points (BC.h:321-343).
#include "DirectXTexP.h"
#include "BC.h"
#include <omp.h>

using namespace DirectX;

extern "C" {

int ref_omp_max_threads() { return omp_get_max_threads(); }
void ref_omp_set_threads(int n) { omp_set_num_threads(n); }

// ComputePitch (DirectXTexUtil.cpp:961)
int32_t ref_compute_pitch(uint32_t fmt, size_t w, size_t h, size_t* rowPitch, size_t* slicePitch)
{
    return ComputePitch(static_cast<DXGI_FORMAT>(fmt), w, h, *rowPitch, *slicePitch, CP_FLAGS_NONE);
}

static Image make_image(const uint8_t* px, size_t w, size_t h, uint32_t fmt, size_t rowPitch)
{
    Image img;
    img.width = w; img.height = h; img.format = static_cast<DXGI_FORMAT>(fmt);
    size_t rp = 0, sp = 0;
    ComputePitch(img.format, w, h, rp, sp, CP_FLAGS_NONE);
    if (rowPitch == 0) rowPitch = rp;
    img.rowPitch = rowPitch;
    img.slicePitch = (rowPitch == rp) ? sp : rowPitch * ComputeScanlines(img.format, h);
    img.pixels = const_cast<uint8_t*>(px);
    return img;
}

// DirectX::Compress single image (DirectXTexCompress.cpp:632)
int32_t ref_compress(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, size_t srcRowPitch,
                     uint32_t dstFmt, uint32_t flags, float threshold, uint8_t* dst, size_t dstBytes)
{
    Image img = make_image(src, w, h, srcFmt, srcRowPitch);
    ScratchImage out;
    HRESULT hr = Compress(img, static_cast<DXGI_FORMAT>(dstFmt), static_cast<TEX_COMPRESS_FLAGS>(flags), threshold, out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

// Times only the Compress() call (allocation of the output included, as in the reference);
// returns seconds, negative on failure.
double ref_compress_timed(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, size_t srcRowPitch,
                          uint32_t dstFmt, uint32_t flags, float threshold)
{
    Image img = make_image(src, w, h, srcFmt, srcRowPitch);
    ScratchImage out;
    const double t0 = omp_get_wtime();
    HRESULT hr = Compress(img, static_cast<DXGI_FORMAT>(dstFmt), static_cast<TEX_COMPRESS_FLAGS>(flags), threshold, out);
    const double t1 = omp_get_wtime();
    return FAILED(hr) ? -1.0 : (t1 - t0);
}

// DirectX::Decompress single image (DirectXTexCompress.cpp:852)
int32_t ref_decompress(const uint8_t* blocks, size_t w, size_t h, uint32_t bcFmt, uint32_t dstFmt, uint8_t* dst, size_t dstBytes)
{
    Image img = make_image(blocks, w, h, bcFmt, 0);
    ScratchImage out;
    HRESULT hr = Decompress(img, static_cast<DXGI_FORMAT>(dstFmt), out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

// DirectX::Convert single image (DirectXTexConvert.cpp:5091)
int32_t ref_convert(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, size_t srcRowPitch,
                    uint32_t dstFmt, uint32_t filter, float threshold, uint8_t* dst, size_t dstBytes)
{
    Image img = make_image(src, w, h, srcFmt, srcRowPitch);
    ScratchImage out;
    HRESULT hr = Convert(img, static_cast<DXGI_FORMAT>(dstFmt), static_cast<TEX_FILTER_FLAGS>(filter), threshold, out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

double ref_convert_timed(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, uint32_t dstFmt, uint32_t filter, float threshold)
{
    Image img = make_image(src, w, h, srcFmt, 0);
    ScratchImage out;
    const double t0 = omp_get_wtime();
    HRESULT hr = Convert(img, static_cast<DXGI_FORMAT>(dstFmt), static_cast<TEX_FILTER_FLAGS>(filter), threshold, out);
    const double t1 = omp_get_wtime();
    return FAILED(hr) ? -1.0 : (t1 - t0);
}

// DirectX::GenerateMipMaps single image (DirectXTexMipmaps.cpp:2828). Output = the whole
// chain (level 0 included) exactly as laid out in the ScratchImage; *outLevels receives the level count.
int32_t ref_generate_mipmaps(const uint8_t* src, size_t w, size_t h, uint32_t fmt, size_t srcRowPitch,
                             uint32_t filter, size_t levels, uint8_t* dst, size_t dstBytes, size_t* outLevels, size_t* outBytes)
{
    Image img = make_image(src, w, h, fmt, srcRowPitch);
    ScratchImage out;
    HRESULT hr = GenerateMipMaps(img, static_cast<TEX_FILTER_FLAGS>(filter), levels, out, false);
    if (FAILED(hr)) return hr;
    if (outLevels) *outLevels = out.GetMetadata().mipLevels;
    if (outBytes) *outBytes = out.GetPixelsSize();
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

// DirectX::Resize single image (DirectXTexResize.cpp:854-935; custom filters :255-837).  dst = width*height pixels, tightly
// packed as the result ScratchImage lays them out.
int32_t ref_resize(const uint8_t* src, size_t w, size_t h, uint32_t fmt, size_t srcRowPitch,
                   size_t width, size_t height, uint32_t filter, uint8_t* dst, size_t dstBytes)
{
    Image img = make_image(src, w, h, fmt, srcRowPitch);
    ScratchImage out;
    HRESULT hr = Resize(img, width, height, static_cast<TEX_FILTER_FLAGS>(filter), out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

// DirectX::PremultiplyAlpha single image (DirectXTexPMAlpha.cpp:214-266)
int32_t ref_premultiply_alpha(const uint8_t* src, size_t w, size_t h, uint32_t fmt, size_t srcRowPitch, uint32_t flags, uint8_t* dst, size_t dstBytes)
{
    Image img = make_image(src, w, h, fmt, srcRowPitch);
    ScratchImage out;
    HRESULT hr = PremultiplyAlpha(img, static_cast<TEX_PMALPHA_FLAGS>(flags), out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

// DDS container (DirectXTexDDS.cpp): a 2D texture (array / cubemap / mip chain) whose images live in `pixels` in ScratchImage
// order.  meta = {width, height, arraySize, mipLevels, format, miscFlags, miscFlags2}.
static HRESULT build_scratch(const uint8_t* pixels, size_t pixelBytes, const uint64_t* meta, ScratchImage& img)
{
    TexMetadata md = {};
    md.width = meta[0]; md.height = meta[1]; md.depth = 1; md.arraySize = meta[2]; md.mipLevels = meta[3];
    md.format = static_cast<DXGI_FORMAT>(meta[4]); md.miscFlags = static_cast<uint32_t>(meta[5]); md.miscFlags2 = static_cast<uint32_t>(meta[6]);
    md.dimension = TEX_DIMENSION_TEXTURE2D;
    HRESULT hr = img.Initialize(md);
    if (FAILED(hr)) return hr;
    if (img.GetPixelsSize() != pixelBytes) return E_INVALIDARG;
    memcpy(img.GetPixels(), pixels, pixelBytes);
    return S_OK;
}
int32_t ref_dds_save(const uint8_t* pixels, size_t pixelBytes, const uint64_t* meta, uint32_t flags, uint8_t* dst, size_t cap, size_t* written)
{
    ScratchImage img;
    HRESULT hr = build_scratch(pixels, pixelBytes, meta, img);
    if (FAILED(hr)) return hr;
    Blob blob;
    hr = SaveToDDSMemory(img.GetImages(), img.GetImageCount(), img.GetMetadata(), static_cast<DDS_FLAGS>(flags), blob);
    if (FAILED(hr)) return hr;
    *written = blob.GetBufferSize();
    if (blob.GetBufferSize() > cap) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, blob.GetBufferPointer(), blob.GetBufferSize());
    return hr;
}
// meta receives the 7 fields above; pixels = the loaded ScratchImage memory
int32_t ref_dds_load(const uint8_t* src, size_t size, uint32_t flags, uint64_t* meta, uint8_t* pixels, size_t cap, size_t* pixelBytes)
{
    TexMetadata md; ScratchImage img;
    HRESULT hr = LoadFromDDSMemory(src, size, static_cast<DDS_FLAGS>(flags), &md, img);
    if (FAILED(hr)) return hr;
    meta[0] = md.width; meta[1] = md.height; meta[2] = md.arraySize; meta[3] = md.mipLevels; meta[4] = md.format;
    meta[5] = md.miscFlags; meta[6] = md.miscFlags2;
    *pixelBytes = img.GetPixelsSize();
    if (img.GetPixelsSize() > cap) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(pixels, img.GetPixels(), img.GetPixelsSize());
    return hr;
}

// GenerateMipMaps then ScaleMipMapsAlphaForCoverage (DirectXTexMipmaps.cpp:3483-3552) on one image, as texconv does
// (Texconv/texconv.cpp:3457-3480).  dst receives the whole chain in ScratchImage layout.
int32_t ref_mips_alpha_coverage(const uint8_t* src, size_t w, size_t h, uint32_t fmt, uint32_t filter, float alphaRef,
                                uint8_t* dst, size_t dstBytes, uint8_t* plainChain)
{
    Image img = make_image(src, w, h, fmt, 0);
    ScratchImage mips;
    HRESULT hr = GenerateMipMaps(img, static_cast<TEX_FILTER_FLAGS>(filter), 0, mips, false);
    if (FAILED(hr)) return hr;
    if (plainChain) memcpy(plainChain, mips.GetPixels(), mips.GetPixelsSize());
    ScratchImage out;
    hr = out.Initialize(mips.GetMetadata());
    if (FAILED(hr)) return hr;
    hr = ScaleMipMapsAlphaForCoverage(mips.GetImages(), mips.GetMetadata().mipLevels, mips.GetMetadata(), 0, alphaRef, out);
    if (FAILED(hr)) return hr;
    if (out.GetPixelsSize() > dstBytes) return E_NOT_SUFFICIENT_BUFFER;
    memcpy(dst, out.GetPixels(), out.GetPixelsSize());
    return hr;
}

double ref_generate_mipmaps_timed(const uint8_t* src, size_t w, size_t h, uint32_t fmt, uint32_t filter, size_t levels)
{
    Image img = make_image(src, w, h, fmt, 0);
    ScratchImage out;
    const double t0 = omp_get_wtime();
    HRESULT hr = GenerateMipMaps(img, static_cast<TEX_FILTER_FLAGS>(filter), levels, out, false);
    const double t1 = omp_get_wtime();
    return FAILED(hr) ? -1.0 : (t1 - t0);
}

// Layout of one mip chain as ScratchImage lays it out (DirectXTexImage.cpp:34-268):
// fills offsets[level], widths, heights, rowPitches; returns number of levels and total bytes.
int32_t ref_mipchain_layout(uint32_t fmt, size_t w, size_t h, size_t levels, size_t* outLevels, size_t* totalBytes,
                            size_t* offsets, size_t* widths, size_t* heights, size_t* rowPitches, size_t cap)
{
    ScratchImage s;
    HRESULT hr = s.Initialize2D(static_cast<DXGI_FORMAT>(fmt), w, h, 1, levels);
    if (FAILED(hr)) return hr;
    const size_t n = s.GetMetadata().mipLevels;
    if (outLevels) *outLevels = n;
    if (totalBytes) *totalBytes = s.GetPixelsSize();
    for (size_t i = 0; i < n && i < cap; ++i)
    {
        const Image* im = s.GetImage(i, 0, 0);
        offsets[i] = static_cast<size_t>(im->pixels - s.GetPixels());
        widths[i] = im->width; heights[i] = im->height; rowPitches[i] = im->rowPitch;
    }
    return hr;
}

// DirectX::ComputeMSE (DirectXTexMisc.cpp:388)
int32_t ref_compute_mse(const uint8_t* a, uint32_t fmtA, const uint8_t* b, uint32_t fmtB, size_t w, size_t h,
                        float* mse, float* mseV4, uint32_t flags)
{
    Image ia = make_image(a, w, h, fmtA, 0);
    Image ib = make_image(b, w, h, fmtB, 0);
    return ComputeMSE(ia, ib, *mse, mseV4, static_cast<CMSE_FLAGS>(flags));
}

// Per-block codec entry points (BC.h:321-343). `rgba` = 16 pixels x 4 floats, row-major.
int32_t ref_encode_block(uint32_t dstFmt, const float* rgba, uint32_t bcflags, float threshold, uint8_t* out)
{
    XM_ALIGNED_DATA(16) XMVECTOR temp[16];
    memcpy(temp, rgba, sizeof(temp));
    switch (static_cast<DXGI_FORMAT>(dstFmt))
    {
    case DXGI_FORMAT_BC1_UNORM: case DXGI_FORMAT_BC1_UNORM_SRGB: D3DXEncodeBC1(out, temp, threshold, bcflags); break;
    case DXGI_FORMAT_BC2_UNORM: case DXGI_FORMAT_BC2_UNORM_SRGB: D3DXEncodeBC2(out, temp, bcflags); break;
    case DXGI_FORMAT_BC3_UNORM: case DXGI_FORMAT_BC3_UNORM_SRGB: D3DXEncodeBC3(out, temp, bcflags); break;
    case DXGI_FORMAT_BC4_UNORM: D3DXEncodeBC4U(out, temp, bcflags); break;
    case DXGI_FORMAT_BC4_SNORM: D3DXEncodeBC4S(out, temp, bcflags); break;
    case DXGI_FORMAT_BC5_UNORM: D3DXEncodeBC5U(out, temp, bcflags); break;
    case DXGI_FORMAT_BC5_SNORM: D3DXEncodeBC5S(out, temp, bcflags); break;
    case DXGI_FORMAT_BC6H_UF16: D3DXEncodeBC6HU(out, temp, bcflags); break;
    case DXGI_FORMAT_BC6H_SF16: D3DXEncodeBC6HS(out, temp, bcflags); break;
    case DXGI_FORMAT_BC7_UNORM: case DXGI_FORMAT_BC7_UNORM_SRGB: D3DXEncodeBC7(out, temp, bcflags); break;
    default: return HRESULT_E_NOT_SUPPORTED;
    }
    return S_OK;
}

int32_t ref_decode_blocks(uint32_t bcFmt, const uint8_t* blocks, size_t nblocks, float* rgba)
{
    size_t bs = 16;
    BC_DECODE fn = nullptr;
    switch (static_cast<DXGI_FORMAT>(bcFmt))
    {
    case DXGI_FORMAT_BC1_UNORM: case DXGI_FORMAT_BC1_UNORM_SRGB: fn = D3DXDecodeBC1; bs = 8; break;
    case DXGI_FORMAT_BC2_UNORM: case DXGI_FORMAT_BC2_UNORM_SRGB: fn = D3DXDecodeBC2; break;
    case DXGI_FORMAT_BC3_UNORM: case DXGI_FORMAT_BC3_UNORM_SRGB: fn = D3DXDecodeBC3; break;
    case DXGI_FORMAT_BC4_UNORM: fn = D3DXDecodeBC4U; bs = 8; break;
    case DXGI_FORMAT_BC4_SNORM: fn = D3DXDecodeBC4S; bs = 8; break;
    case DXGI_FORMAT_BC5_UNORM: fn = D3DXDecodeBC5U; break;
    case DXGI_FORMAT_BC5_SNORM: fn = D3DXDecodeBC5S; break;
    case DXGI_FORMAT_BC6H_UF16: fn = D3DXDecodeBC6HU; break;
    case DXGI_FORMAT_BC6H_SF16: fn = D3DXDecodeBC6HS; break;
    case DXGI_FORMAT_BC7_UNORM: case DXGI_FORMAT_BC7_UNORM_SRGB: fn = D3DXDecodeBC7; break;
    default: return HRESULT_E_NOT_SUPPORTED;
    }
    #pragma omp parallel for
    for (long i = 0; i < static_cast<long>(nblocks); ++i)
    {
        XM_ALIGNED_DATA(16) XMVECTOR temp[16];
        fn(temp, blocks + static_cast<size_t>(i) * bs);
        memcpy(rgba + static_cast<size_t>(i) * 64, temp, sizeof(temp));
    }
    return S_OK;
}

} // extern "C"
