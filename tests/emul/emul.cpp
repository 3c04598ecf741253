// tests/emul/emul.cpp — TEST INFRASTRUCTURE, never shipped.
//
// Lock-step HOST build of the exact arithmetic sources the CUDA kernels are compiled from
// (directxtex_b200/csrc/*.cuh with DXB_DEV = plain inline).  It exists so that parity against the
// oracle can be debugged in this GPU-less container; the GPU tests then check the sm_100a build
// against the oracle AND against this emulator.  The product library never contains, loads or
// calls any of this: there is no CPU fallback.
//
// Build flags mirror the device build's numeric contract: -ffp-contract=off (== nvcc -fmad=false),
// IEEE division/sqrt, -mfma only so that explicit fmaf() is one instruction (fmaf is exact either way).
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>

#include "dxb_portable.h"
#include "dxb_formats.h"
#include "dxb_pixel.cuh"
#include "dxb_block.cuh"
#include "dxb_bc15.cuh"
#ifdef DXB_EMUL_BC7
#include "dxb_bc7.cuh"
#endif

extern "C" {

int32_t emul_compress(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, size_t rowPitch,
                      uint32_t dstFmt, uint32_t flags, float threshold, uint8_t* dst)
{
    const uint32_t inF = dxb_convert_flags(srcFmt), outF = dxb_convert_flags(dstFmt);
    const uint32_t bs = dxb_bc_block_bytes(dstFmt);
    if (!inF || !outF || !bs || (inF & DXB_CONVF_BC)) return DXB_E_NOT_SUPPORTED;
    if (rowPitch == 0) rowPitch = w * dxb_bytes_per_pixel(srcFmt);
    uint32_t cflags = 0;
    if (dstFmt == DXB_FMT_BC4_UNORM || dstFmt == DXB_FMT_BC4_SNORM) cflags = DXB_FILTER_RGB_COPY_RED;
    if (dstFmt == DXB_FMT_BC5_UNORM || dstFmt == DXB_FMT_BC5_SNORM) cflags = DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_GREEN;
    cflags |= (flags & DXB_FILTER_SRGB_MASK);
    cflags = dxb_resolve_srgb_convert(cflags, srcFmt, dstFmt);
    const uint32_t bcflags = flags & (DXB_BC_FLAGS_DITHER_RGB | DXB_BC_FLAGS_DITHER_A | DXB_BC_FLAGS_UNIFORM | DXB_BC_FLAGS_USE_3SUBSETS | DXB_BC_FLAGS_FORCE_BC7_MODE6);
    dxb_image_desc img; img.pixels = src; img.rowPitch = rowPitch; img.width = (uint32_t)w; img.height = (uint32_t)h; img.format = srcFmt;
    const uint32_t nbx = (uint32_t)((w + 3) / 4), nby = (uint32_t)((h + 3) / 4);
    #pragma omp parallel for schedule(dynamic, 8)
    for (long by = 0; by < (long)nby; ++by)
        for (uint32_t bx = 0; bx < nbx; ++bx)
        {
            dxb_px px[16];
            dxb_gather_block(img, bx, (uint32_t)by, inF, outF, cflags, px);
            alignas(16) uint8_t blk[16];
#ifdef DXB_EMUL_BC7
            if (dstFmt == DXB_FMT_BC7_UNORM || dstFmt == DXB_FMT_BC7_UNORM_SRGB)
                dxb_bc7_encode_block_emul(px, bcflags, blk);
            else
#endif
                dxb_encode_block_bc15(dstFmt, px, bcflags, threshold, blk);
            memcpy(dst + ((size_t)by * nbx + bx) * bs, blk, bs);
        }
    return DXB_S_OK;
}

// Row-wise format conversion (ConvertCustom no-dither path, DirectXTexConvert.cpp:4888-4908)
int32_t emul_convert(const uint8_t* src, size_t w, size_t h, uint32_t srcFmt, size_t srcPitch,
                     uint32_t dstFmt, size_t dstPitch, uint32_t filter, uint8_t* dst)
{
    const uint32_t inF = dxb_convert_flags(srcFmt), outF = dxb_convert_flags(dstFmt);
    if (!inF || !outF || ((inF | outF) & DXB_CONVF_BC)) return DXB_E_NOT_SUPPORTED;
    if (srcPitch == 0) srcPitch = w * dxb_bytes_per_pixel(srcFmt);
    if (dstPitch == 0) dstPitch = w * dxb_bytes_per_pixel(dstFmt);
    const uint32_t flags = dxb_resolve_srgb_convert(filter, srcFmt, dstFmt);
    for (size_t y = 0; y < h; ++y)
        for (size_t x = 0; x < w; ++x)
        {
            dxb_px v = dxb_load_pixel(srcFmt, src + y * srcPitch, x);
            v = dxb_convert_pixel(v, inF, outF, flags);
            dxb_store_pixel(dstFmt, dst + y * dstPitch, x, v);
        }
    return DXB_S_OK;
}

} // extern "C"
