// dxb_block.cuh — gather one 4x4 block the way CompressBC does (DirectXTexCompress.cpp:123-189):
// up to 4 rows x up to 4 pixels through LoadScanline, partial blocks completed by replicating
// with the {0,0,0,1} source map (:159-187), then ConvertScanline(16 px) into the BC format's class.
#pragma once
#include "dxb_pixel.cuh"

struct dxb_image_desc
{
    const uint8_t* pixels;   // device (or, in tests/emul, host) pointer to row 0
    size_t rowPitch;
    uint32_t width, height;
    uint32_t format;
};

DXB_DEV void dxb_gather_block(const dxb_image_desc& img, uint32_t bx, uint32_t by,
                              uint32_t inF, uint32_t outF, uint32_t cflags, dxb_px* px)
{
    const uint32_t x0 = bx * 4, y0 = by * 4;
    const uint32_t pw = (img.width - x0 < 4u) ? (img.width - x0) : 4u;
    const uint32_t ph = (img.height - y0 < 4u) ? (img.height - y0) : 4u;
    for (uint32_t t = 0; t < ph; ++t)
    {
        const uint8_t* row = img.pixels + (size_t)(y0 + t) * img.rowPitch;
        for (uint32_t s = 0; s < pw; ++s) px[(t << 2) | s] = dxb_load_pixel(img.format, row, x0 + s);
    }
    if (pw != 4 || ph != 4)
    {
        // uSrc = {0,0,0,1}
        if (pw < 4)
            for (uint32_t t = 0; t < ph; ++t)
                for (uint32_t s = pw; s < 4; ++s) px[(t << 2) | s] = px[(t << 2) | (s == 3 ? 1u : 0u)];
        if (ph < 4)
            for (uint32_t t = ph; t < 4; ++t)
                for (uint32_t s = 0; s < 4; ++s) px[(t << 2) | s] = px[((t == 3 ? 1u : 0u) << 2) | s];
    }
    for (int i = 0; i < 16; ++i) px[i] = dxb_convert_pixel(px[i], inF, outF, cflags);
}

#if DXB_ON_DEVICE
// Same result as dxb_gather_block for a compile-time source format: a full 4x4 block whose rows are suitably aligned
// is fetched with ONE vector load per row (4 pixels = 4..64 bytes) instead of 16 scalar pixel loads, then decoded
// from registers by the same dxb_load_pixel code.
template <uint32_t SF>
__device__ __forceinline__ void dxb_gather_block_t(const dxb_image_desc& img, uint32_t bx, uint32_t by,
                                                   uint32_t inF, uint32_t outF, uint32_t cflags, dxb_px* px)
{
    constexpr uint32_t B = dxb_bytes_per_pixel(SF);
    constexpr uint32_t ROWB = 4u * B;                                     // bytes of 4 pixels
    constexpr uint32_t V = (ROWB >= 16u) ? 16u : ROWB;                    // vector width used
    const uint32_t x0 = bx * 4u, y0 = by * 4u;
    const uint8_t* p0 = img.pixels + (size_t)y0 * img.rowPitch + (size_t)x0 * B;
    const bool full = (img.width - x0 >= 4u) && (img.height - y0 >= 4u);
    if (!(full && (((uintptr_t)p0 | img.rowPitch) & (V - 1u)) == 0u))
    {
        dxb_gather_block(img, bx, by, inF, outF, cflags, px);
        return;
    }
    #pragma unroll
    for (uint32_t t = 0; t < 4; ++t)
    {
        __align__(16) uint8_t row[ROWB];
        const uint8_t* p = p0 + (size_t)t * img.rowPitch;
        #pragma unroll
        for (uint32_t k = 0; k < ROWB; k += V)
        {
            if (V == 16u) *reinterpret_cast<uint4*>(row + k) = __ldg(reinterpret_cast<const uint4*>(p + k));
            else if (V == 8u) *reinterpret_cast<uint2*>(row + k) = __ldg(reinterpret_cast<const uint2*>(p + k));
            else *reinterpret_cast<uint32_t*>(row + k) = __ldg(reinterpret_cast<const uint32_t*>(p + k));
        }
        #pragma unroll
        for (uint32_t s2 = 0; s2 < 4; ++s2) px[(t << 2) | s2] = dxb_convert_pixel(dxb_load_pixel(SF, row, s2), inF, outF, cflags);
    }
}
#endif
