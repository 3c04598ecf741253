// dxb_pixel.cuh — per-pixel load / convert / store in the reference's canonical intermediate
// (RGBA fp32, one "XMVECTOR" per pixel).  Restates, for the implemented format subset:
//   LoadScanline      DirectXTexConvert.cpp:779-1628   (+ packed loads of DirectXMath, see oracle/compat)
//   ConvertScanline   DirectXTexConvert.cpp:3080-3854
//   StoreScanline     DirectXTexConvert.cpp:1643-2530  (8-bit bias constant :198-199)
//   sRGB helpers      XMColorSRGBToRGB / XMColorRGBToSRGB as used at :3169-3180, :3842-3853
// Every arithmetic step keeps the reference's operation order; the translation unit is compiled
// with -fmad=false so no multiply-add is contracted.
#pragma once
#include "dxb_portable.h"
#include "dxb_formats.h"

struct dxb_px { float x, y, z, w; };

DXB_DEV dxb_px dxb_make_px(float x, float y, float z, float w) { dxb_px p; p.x = x; p.y = y; p.z = z; p.w = w; return p; }

// a / b for an INTEGER-valued a and a literal b in {255, 127, 65535, 32767}: q = a*(1/b); q += fma(-q,b,a)*(1/b) equals the
// correctly rounded IEEE quotient for every value of the 8/16-bit input domains (checked exhaustively on the CPU and in
// tests/test_gpu_parity.py::test_convert_exhaustive_small_domains); 3 FP ops instead of a ~15-instruction division.
DXB_DEV float dxb_div_small(float a, float b)
{
    const float rcp = 1.0f / b;
    const float q = a * rcp;
    const float r = dxb_fma(-q, b, a);
    return dxb_fma(r, rcp, q);
}

DXB_DEV float dxb_snorm_load(int32_t v, float rcp) { return dxb_ssemax(dxb_i2f_small((int32_t)v) * rcp, -1.0f); }
DXB_DEV float dxb_clamp(float v, float lo, float hi) { return dxb_ssemin(dxb_ssemax(v, lo), hi); }

// ---------------------------------------------------------------------------------------------
// Load pixel `i` of a row starting at `row` (byte pointer).  Missing channels default to (0,0,0,1).

// ---- packed small-float formats, restated from the oracle's DirectXMath stand-in (oracle/compat/DirectXPackedVector.h,
// XMLoadFloat3PK / XMStoreFloat3PK / XMLoadFloat3SE / XMStoreFloat3SE; marked (M) there: DirectXMath itself is not in the tree)
DXB_DEV float dxb_smallfloat_decode(uint32_t m, uint32_t e, uint32_t mbits)
{
    uint32_t bits;
    if (e == 0x1fu) bits = 0x7f800000u | (m << (23u - mbits));
    else
    {
        int32_t ex;
        if (e != 0u) ex = (int32_t)e;
        else if (m != 0u)
        {
            ex = 1;
            do { ex--; m <<= 1; } while ((m & (1u << mbits)) == 0u);
            m &= (1u << mbits) - 1u;
        }
        else ex = -112;
        bits = ((uint32_t)(ex + 112) << 23) | (m << (23u - mbits));
    }
    return dxb_uint_as_float(bits);
}
DXB_DEV uint32_t dxb_smallfloat_encode(float f, uint32_t mbits)
{
    uint32_t I = dxb_float_as_uint(f);
    const uint32_t sign = I & 0x80000000u;
    I &= 0x7FFFFFFFu;
    const uint32_t maxv = (mbits == 6u) ? 0x7C0u : 0x3E0u;
    const uint32_t shift = 23u - mbits;
    if ((I & 0x7F800000u) == 0x7F800000u)
    {
        uint32_t r = maxv;
        if (I & 0x7FFFFFu) r = maxv | (((I >> shift) | (I >> (shift - 6u)) | (I >> (shift - 12u)) | I) & ((1u << mbits) - 1u));
        else if (sign) r = 0u;
        return r;
    }
    if (sign) return 0u;
    if (I > 0x477E0000u && mbits == 6u) return 0x7BFu;
    if (I > 0x477C0000u && mbits == 5u) return 0x3DFu;
    if (I < 0x38800000u)
    {
        const uint32_t sh = 113u - (I >> 23);
        I = (sh < 32u) ? ((0x800000u | (I & 0x7FFFFFu)) >> sh) : 0u;
    }
    else I += 0xC8000000u;
    const uint32_t half = (1u << (shift - 1u)) - 1u;
    return ((I + half + ((I >> shift) & 1u)) >> shift) & ((mbits == 6u) ? 0x7FFu : 0x3FFu);
}
// pk::st_int of the stand-in: clamp, round to nearest even
DXB_DEV uint32_t dxb_store_int_rne(float v, float hi) { return (uint32_t)dxb_f2i_rn(dxb_clamp(v, 0.0f, hi)); }

DXB_DEV dxb_px dxb_load_pixel(uint32_t fmt, const uint8_t* row, size_t i)
{
    switch (fmt)
    {
    case DXB_FMT_R11G11B10_FLOAT:        // XMLoadFloat3PK, alpha 1 (DirectXTexConvert.cpp:906-920)
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        return dxb_make_px(dxb_smallfloat_decode(v & 0x3Fu, (v >> 6) & 0x1Fu, 6u), dxb_smallfloat_decode((v >> 11) & 0x3Fu, (v >> 17) & 0x1Fu, 6u),
                           dxb_smallfloat_decode((v >> 22) & 0x1Fu, (v >> 27) & 0x1Fu, 5u), 1.0f);
    }
    case DXB_FMT_R9G9B9E5_SHAREDEXP:     // XMLoadFloat3SE, alpha 1 (:1189-1203)
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        const float scale = dxb_uint_as_float(0x33800000u + ((v >> 27) << 23));
        return dxb_make_px(scale * (float)(v & 0x1FFu), scale * (float)((v >> 9) & 0x1FFu), scale * (float)((v >> 18) & 0x1FFu), 1.0f);
    }
    case DXB_FMT_B5G6R5_UNORM:           // XMLoadU565 * {1/31, 1/63, 1/31}, swizzled to RGB, alpha 1 (:1227-1242)
    {
        const uint32_t v = ((const uint16_t*)row)[i];
        return dxb_make_px((float)((v >> 11) & 0x1Fu) * (1.0f / 31.0f), (float)((v >> 5) & 0x3Fu) * (1.0f / 63.0f), (float)(v & 0x1Fu) * (1.0f / 31.0f), 1.0f);
    }
    case DXB_FMT_B5G5R5A1_UNORM:         // XMLoadU555 * {1/31 x3, 1}, swizzled (:1244-1258)
    {
        const uint32_t v = ((const uint16_t*)row)[i];
        return dxb_make_px((float)((v >> 10) & 0x1Fu) * (1.0f / 31.0f), (float)((v >> 5) & 0x1Fu) * (1.0f / 31.0f), (float)(v & 0x1Fu) * (1.0f / 31.0f), (float)(v >> 15) * 1.0f);
    }
    case DXB_FMT_B4G4R4A4_UNORM:         // XMLoadUNibble4 * 1/15, swizzled (:1511-1526)
    {
        const uint32_t v = ((const uint16_t*)row)[i];
        return dxb_make_px((float)((v >> 8) & 0xFu) * (1.0f / 15.0f), (float)((v >> 4) & 0xFu) * (1.0f / 15.0f), (float)(v & 0xFu) * (1.0f / 15.0f), (float)(v >> 12) * (1.0f / 15.0f));
    }
    case DXB_FMT_R32G32B32A32_FLOAT:
    {
        const float* p = (const float*)row + i * 4;
        return dxb_make_px(p[0], p[1], p[2], p[3]);
    }
    case DXB_FMT_R32G32B32_FLOAT:
    {
        const float* p = (const float*)row + i * 3;
        return dxb_make_px(p[0], p[1], p[2], 1.0f);
    }
    case DXB_FMT_R16G16B16A16_FLOAT:
    {
        const uint16_t* p = (const uint16_t*)row + i * 4;
        return dxb_make_px(dxb_half_to_float(p[0]), dxb_half_to_float(p[1]), dxb_half_to_float(p[2]), dxb_half_to_float(p[3]));
    }
    case DXB_FMT_R16G16B16A16_UNORM:
    {
        const uint16_t* p = (const uint16_t*)row + i * 4;
        const float s = 1.0f / 65535.0f;
        return dxb_make_px(dxb_i2f_small((int32_t)p[0]) * s, dxb_i2f_small((int32_t)p[1]) * s, dxb_i2f_small((int32_t)p[2]) * s, dxb_i2f_small((int32_t)p[3]) * s);
    }
    case DXB_FMT_R16G16B16A16_SNORM:
    {
        const int16_t* p = (const int16_t*)row + i * 4;
        const float s = 1.0f / 32767.0f;
        return dxb_make_px(dxb_snorm_load(p[0], s), dxb_snorm_load(p[1], s), dxb_snorm_load(p[2], s), dxb_snorm_load(p[3], s));
    }
    case DXB_FMT_R32G32_FLOAT:
    {
        const float* p = (const float*)row + i * 2;
        return dxb_make_px(p[0], p[1], 0.0f, 1.0f);
    }
    case DXB_FMT_R10G10B10A2_UNORM:
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        const float s = 1.0f / 1023.0f;
        return dxb_make_px(dxb_i2f_small((int32_t)(v & 0x3FF)) * s, dxb_i2f_small((int32_t)((v >> 10) & 0x3FF)) * s, dxb_i2f_small((int32_t)((v >> 20) & 0x3FF)) * s, dxb_i2f_small((int32_t)(v >> 30)) * (1.0f / 3.0f));
    }
    case DXB_FMT_R8G8B8A8_UNORM:
    case DXB_FMT_R8G8B8A8_UNORM_SRGB:
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        const float s = 1.0f / 255.0f;
        return dxb_make_px(dxb_byte_to_float(v, 0) * s, dxb_byte_to_float(v, 1) * s, dxb_byte_to_float(v, 2) * s, dxb_byte_to_float(v, 3) * s);
    }
    case DXB_FMT_B8G8R8A8_UNORM:
    case DXB_FMT_B8G8R8A8_UNORM_SRGB:
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        const float s = 1.0f / 255.0f;
        return dxb_make_px(dxb_byte_to_float(v, 2) * s, dxb_byte_to_float(v, 1) * s, dxb_byte_to_float(v, 0) * s, dxb_byte_to_float(v, 3) * s);
    }
    case DXB_FMT_B8G8R8X8_UNORM:
    case DXB_FMT_B8G8R8X8_UNORM_SRGB:
    {
        const uint32_t v = ((const uint32_t*)row)[i];
        const float s = 1.0f / 255.0f;
        return dxb_make_px(dxb_byte_to_float(v, 2) * s, dxb_byte_to_float(v, 1) * s, dxb_byte_to_float(v, 0) * s, 1.0f);
    }
    case DXB_FMT_R8G8B8A8_SNORM:
    {
        const int8_t* p = (const int8_t*)row + i * 4;
        const float s = 1.0f / 127.0f;
        return dxb_make_px(dxb_snorm_load(p[0], s), dxb_snorm_load(p[1], s), dxb_snorm_load(p[2], s), dxb_snorm_load(p[3], s));
    }
    case DXB_FMT_R16G16_FLOAT:
    {
        const uint16_t* p = (const uint16_t*)row + i * 2;
        return dxb_make_px(dxb_half_to_float(p[0]), dxb_half_to_float(p[1]), 0.0f, 1.0f);
    }
    case DXB_FMT_R16G16_UNORM:
    {
        const uint16_t* p = (const uint16_t*)row + i * 2;
        const float s = 1.0f / 65535.0f;
        return dxb_make_px(dxb_i2f_small((int32_t)p[0]) * s, dxb_i2f_small((int32_t)p[1]) * s, 0.0f, 1.0f);
    }
    case DXB_FMT_R16G16_SNORM:
    {
        const int16_t* p = (const int16_t*)row + i * 2;
        const float s = 1.0f / 32767.0f;
        return dxb_make_px(dxb_snorm_load(p[0], s), dxb_snorm_load(p[1], s), 0.0f, 1.0f);
    }
    case DXB_FMT_R32_FLOAT:
        return dxb_make_px(((const float*)row)[i], 0.0f, 0.0f, 1.0f);
    case DXB_FMT_R8G8_UNORM:
    {
        const uint8_t* p = row + i * 2;
        const float s = 1.0f / 255.0f;
        return dxb_make_px(dxb_i2f_small((int32_t)p[0]) * s, dxb_i2f_small((int32_t)p[1]) * s, 0.0f, 1.0f);
    }
    case DXB_FMT_R8G8_SNORM:
    {
        const int8_t* p = (const int8_t*)row + i * 2;
        const float s = 1.0f / 127.0f;
        return dxb_make_px(dxb_snorm_load(p[0], s), dxb_snorm_load(p[1], s), 0.0f, 1.0f);
    }
    case DXB_FMT_R16_FLOAT:
        return dxb_make_px(dxb_half_to_float(((const uint16_t*)row)[i]), 0.0f, 0.0f, 1.0f);
    case DXB_FMT_R16_UNORM:      // true division: DirectXTexConvert.cpp:1062
        return dxb_make_px(dxb_div_small(dxb_i2f_small((int32_t)((const uint16_t*)row)[i]), 65535.0f), 0.0f, 0.0f, 1.0f);
    case DXB_FMT_R16_SNORM:      // :1088 (no clamp of -32768)
        return dxb_make_px(dxb_div_small(dxb_i2f_small((int32_t)((const int16_t*)row)[i]), 32767.0f), 0.0f, 0.0f, 1.0f);
    case DXB_FMT_R8_UNORM:       // :1113
        return dxb_make_px(dxb_div_small(dxb_i2f_small((int32_t)row[i]), 255.0f), 0.0f, 0.0f, 1.0f);
    case DXB_FMT_R8_SNORM:       // :1139
        return dxb_make_px(dxb_div_small(dxb_i2f_small((int32_t)((const int8_t*)row)[i]), 127.0f), 0.0f, 0.0f, 1.0f);
    case DXB_FMT_A8_UNORM:       // :1165
        return dxb_make_px(0.0f, 0.0f, 0.0f, dxb_div_small(dxb_i2f_small((int32_t)row[i]), 255.0f));
    default:
        return dxb_make_px(0.0f, 0.0f, 0.0f, 1.0f);
    }
}

// ---------------------------------------------------------------------------------------------
// sRGB <-> linear on xyz (w untouched).  powf is not bit-identical between glibc and CUDA libm:
// sRGB paths are tolerance-checked.
DXB_DEV float dxb_srgb_to_linear1(float c)
{
    const float v = dxb_ssemin(dxb_ssemax(c, 0.0f), 1.0f);
    const float v0 = v * (1.0f / 12.92f);
    const float t = (v + 0.055f) * (1.0f / 1.055f);
    const float v1 = powf(t, 2.4f);
    return (v > 0.04045f) ? v1 : v0;
}
DXB_DEV float dxb_linear_to_srgb1(float c)
{
    const float v = dxb_ssemin(dxb_ssemax(c, 0.0f), 1.0f);
    const float v0 = v * 12.92f;
    const float p = powf(v, 1.0f / 2.4f);
    const float v1 = 1.055f * p - 0.055f;
    return (v < 0.0031308f) ? v0 : v1;
}
DXB_DEV dxb_px dxb_srgb_to_linear(dxb_px v) { v.x = dxb_srgb_to_linear1(v.x); v.y = dxb_srgb_to_linear1(v.y); v.z = dxb_srgb_to_linear1(v.z); return v; }
DXB_DEV dxb_px dxb_linear_to_srgb(dxb_px v) { v.x = dxb_linear_to_srgb1(v.x); v.y = dxb_linear_to_srgb1(v.y); v.z = dxb_linear_to_srgb1(v.z); return v; }

DXB_DEV float dxb_grayscale(dxb_px v)
{
    // XMVector3Dot(v, {0.2125, 0.7154, 0.0721}) = (x*a + y*b) + z*c, unfused
    const float a = v.x * 0.2125f, b = v.y * 0.7154f, c = v.z * 0.0721f;
    const float ab = a + b;
    return ab + c;
}
DXB_DEV float dxb_madd(float a, float b, float c) { const float p = a * b; return p + c; }   // XMVectorMultiplyAdd, unfused

// ConvertScanline for one pixel.  `inF`/`outF` = dxb_convert_flags of the two formats, `flags` =
// TEX_FILTER flags with the sRGB bits already resolved by dxb_resolve_srgb_convert.  Depth, UINT,
// SINT, POS_ONLY, XR, YUV and PACKED formats are rejected on the host before launch.
DXB_DEV dxb_px dxb_convert_pixel(dxb_px v, uint32_t inF, uint32_t outF, uint32_t flags)
{
    if (flags & DXB_FILTER_SRGB_IN)
    {
        if ((inF & DXB_CONVF_FLOAT) || (inF & DXB_CONVF_UNORM)) v = dxb_srgb_to_linear(v);
    }

    const uint32_t diff = inF ^ outF;
    if (diff != 0)
    {
        if (outF & DXB_CONVF_UNORM)
        {
            if (inF & DXB_CONVF_SNORM)
            {
                v.x = dxb_madd(v.x, 0.5f, 0.5f); v.y = dxb_madd(v.y, 0.5f, 0.5f); v.z = dxb_madd(v.z, 0.5f, 0.5f); v.w = dxb_madd(v.w, 0.5f, 0.5f);
            }
            else if (inF & DXB_CONVF_FLOAT)
            {
                if (!(inF & DXB_CONVF_POS_ONLY) && (flags & DXB_FILTER_FLOAT_X2BIAS))
                {
                    v.x = dxb_madd(dxb_clamp(v.x, -1.0f, 1.0f), 0.5f, 0.5f); v.y = dxb_madd(dxb_clamp(v.y, -1.0f, 1.0f), 0.5f, 0.5f);
                    v.z = dxb_madd(dxb_clamp(v.z, -1.0f, 1.0f), 0.5f, 0.5f); v.w = dxb_madd(dxb_clamp(v.w, -1.0f, 1.0f), 0.5f, 0.5f);
                }
                else
                {
                    v.x = dxb_clamp(v.x, 0.0f, 1.0f); v.y = dxb_clamp(v.y, 0.0f, 1.0f); v.z = dxb_clamp(v.z, 0.0f, 1.0f); v.w = dxb_clamp(v.w, 0.0f, 1.0f);
                }
            }
        }
        else if (outF & DXB_CONVF_SNORM)
        {
            if (inF & DXB_CONVF_UNORM)
            {
                v.x = dxb_madd(v.x, 2.0f, -1.0f); v.y = dxb_madd(v.y, 2.0f, -1.0f); v.z = dxb_madd(v.z, 2.0f, -1.0f); v.w = dxb_madd(v.w, 2.0f, -1.0f);
            }
            else if (inF & DXB_CONVF_FLOAT)
            {
                if ((inF & DXB_CONVF_POS_ONLY) && (flags & DXB_FILTER_FLOAT_X2BIAS))
                {
                    // FLOAT (positive only, x2 bias) -> SNORM (:3506-3516)
                    v.x = dxb_madd(dxb_clamp(v.x, 0.0f, 1.0f), 2.0f, -1.0f); v.y = dxb_madd(dxb_clamp(v.y, 0.0f, 1.0f), 2.0f, -1.0f);
                    v.z = dxb_madd(dxb_clamp(v.z, 0.0f, 1.0f), 2.0f, -1.0f); v.w = dxb_madd(dxb_clamp(v.w, 0.0f, 1.0f), 2.0f, -1.0f);
                }
                else
                {
                    v.x = dxb_clamp(v.x, -1.0f, 1.0f); v.y = dxb_clamp(v.y, -1.0f, 1.0f); v.z = dxb_clamp(v.z, -1.0f, 1.0f); v.w = dxb_clamp(v.w, -1.0f, 1.0f);
                }
            }
        }
        else if (diff & DXB_CONVF_UNORM)
        {
            if ((outF & DXB_CONVF_FLOAT) && !(outF & DXB_CONVF_POS_ONLY) && (flags & DXB_FILTER_FLOAT_X2BIAS))
            {
                v.x = dxb_madd(v.x, 2.0f, -1.0f); v.y = dxb_madd(v.y, 2.0f, -1.0f); v.z = dxb_madd(v.z, 2.0f, -1.0f); v.w = dxb_madd(v.w, 2.0f, -1.0f);
            }
        }
        else if ((diff & DXB_CONVF_POS_ONLY) && (flags & DXB_FILTER_FLOAT_X2BIAS))
        {
            // positive-only float formats with the x2 bias (:3545-3583)
            if (inF & DXB_CONVF_POS_ONLY)
            {
                if (outF & DXB_CONVF_FLOAT)
                {
                    v.x = dxb_madd(dxb_clamp(v.x, 0.0f, 1.0f), 2.0f, -1.0f); v.y = dxb_madd(dxb_clamp(v.y, 0.0f, 1.0f), 2.0f, -1.0f);
                    v.z = dxb_madd(dxb_clamp(v.z, 0.0f, 1.0f), 2.0f, -1.0f); v.w = dxb_madd(dxb_clamp(v.w, 0.0f, 1.0f), 2.0f, -1.0f);
                }
            }
            else if (outF & DXB_CONVF_POS_ONLY)
            {
                if (inF & DXB_CONVF_FLOAT)
                {
                    v.x = dxb_madd(dxb_clamp(v.x, -1.0f, 1.0f), 0.5f, 0.5f); v.y = dxb_madd(dxb_clamp(v.y, -1.0f, 1.0f), 0.5f, 0.5f);
                    v.z = dxb_madd(dxb_clamp(v.z, -1.0f, 1.0f), 0.5f, 0.5f); v.w = dxb_madd(dxb_clamp(v.w, -1.0f, 1.0f), 0.5f, 0.5f);
                }
                else if (inF & DXB_CONVF_SNORM)
                {
                    v.x = dxb_madd(v.x, 0.5f, 0.5f); v.y = dxb_madd(v.y, 0.5f, 0.5f); v.z = dxb_madd(v.z, 0.5f, 0.5f); v.w = dxb_madd(v.w, 0.5f, 0.5f);
                }
            }
        }

        const uint32_t inRGB = inF & DXB_CONVF_RGB_MASK, outRGB = outF & DXB_CONVF_RGB_MASK;
        const uint32_t RGB = DXB_CONVF_R | DXB_CONVF_G | DXB_CONVF_B, RG = DXB_CONVF_R | DXB_CONVF_G;
        if (((outF & DXB_CONVF_RGBA_MASK) == DXB_CONVF_A) && !(inF & DXB_CONVF_A))
        {
            float s;
            switch (flags & (DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_BLUE))
            {
            case DXB_FILTER_RGB_COPY_GREEN: s = v.y; break;
            case DXB_FILTER_RGB_COPY_BLUE: s = v.z; break;
            case DXB_FILTER_RGB_COPY_RED: s = v.x; break;
            default: s = ((inF & DXB_CONVF_UNORM) && inRGB == RGB) ? dxb_grayscale(v) : v.x; break;
            }
            v = dxb_make_px(s, s, s, s);
        }
        else if (((inF & DXB_CONVF_RGBA_MASK) == DXB_CONVF_A) && !(outF & DXB_CONVF_A))
        {
            v = dxb_make_px(v.w, v.w, v.w, v.w);
        }
        else if (inRGB == DXB_CONVF_R)
        {
            if (outRGB == RGB) { v.y = v.x; v.z = v.x; }
            else if (outRGB == RG) { v.y = v.x; }
        }
        else if (inRGB == RGB)
        {
            if (outRGB == DXB_CONVF_R)
            {
                switch (flags & (DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_BLUE | DXB_FILTER_RGB_COPY_ALPHA))
                {
                case DXB_FILTER_RGB_COPY_GREEN: v.x = v.y; v.z = v.y; break;
                case DXB_FILTER_RGB_COPY_BLUE: v.x = v.z; v.y = v.z; break;
                case DXB_FILTER_RGB_COPY_ALPHA: v.x = v.w; v.y = v.w; v.z = v.w; break;
                case DXB_FILTER_RGB_COPY_RED: break;
                default:
                    if (inF & DXB_CONVF_UNORM) { const float g = dxb_grayscale(v); v.x = g; v.y = g; v.z = g; }
                    break;
                }
            }
            else if (outRGB == RG)
            {
                if ((flags & DXB_FILTER_RGB_COPY_ALPHA) && (inF & DXB_CONVF_A))
                {
                    switch (flags & (DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_BLUE | DXB_FILTER_RGB_COPY_ALPHA))
                    {
                    case (DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_ALPHA): v.x = v.y; v.y = v.w; break;
                    case (DXB_FILTER_RGB_COPY_BLUE | DXB_FILTER_RGB_COPY_ALPHA): v.x = v.z; v.y = v.w; break;
                    default: v.y = v.w; break;
                    }
                }
                else
                {
                    switch (flags & (DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_BLUE))
                    {
                    case (DXB_FILTER_RGB_COPY_RED | DXB_FILTER_RGB_COPY_BLUE): v.y = v.z; break;
                    case (DXB_FILTER_RGB_COPY_GREEN | DXB_FILTER_RGB_COPY_BLUE): v.x = v.y; v.y = v.z; break;
                    default: break;
                    }
                }
            }
        }
    }

    if (flags & DXB_FILTER_SRGB_OUT)
    {
        if ((outF & DXB_CONVF_FLOAT) || (outF & DXB_CONVF_UNORM)) v = dxb_linear_to_srgb(v);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
DXB_DEV uint32_t dxb_unorm8_trunc(float v)       // +0.5/255 bias, saturate, *255, truncate
{
    const float b = v + (0.5f / 255.0f);
    return dxb_f2u_trunc_small(dxb_clamp(b, 0.0f, 1.0f) * 255.0f);
}
// four channels stored with dxb_unorm8_trunc semantics, packed b0 | b1 << 8 | b2 << 16 | b3 << 24
DXB_DEV uint32_t dxb_pack_unorm8x4(float b0, float b1, float b2, float b3)
{
#if DXB_ON_DEVICE
    // 2^23 + x rounded toward zero keeps floor(x) (<= 255) in the low mantissa byte: three PRMTs gather the four bytes
    const uint32_t u0 = __float_as_uint(__fadd_rz(dxb_clamp(b0 + (0.5f / 255.0f), 0.0f, 1.0f) * 255.0f, 8388608.0f));
    const uint32_t u1 = __float_as_uint(__fadd_rz(dxb_clamp(b1 + (0.5f / 255.0f), 0.0f, 1.0f) * 255.0f, 8388608.0f));
    const uint32_t u2 = __float_as_uint(__fadd_rz(dxb_clamp(b2 + (0.5f / 255.0f), 0.0f, 1.0f) * 255.0f, 8388608.0f));
    const uint32_t u3 = __float_as_uint(__fadd_rz(dxb_clamp(b3 + (0.5f / 255.0f), 0.0f, 1.0f) * 255.0f, 8388608.0f));
    return __byte_perm(__byte_perm(u0, u1, 0x0040), __byte_perm(u2, u3, 0x0040), 0x5410);
#else
    return dxb_unorm8_trunc(b0) | (dxb_unorm8_trunc(b1) << 8) | (dxb_unorm8_trunc(b2) << 16) | (dxb_unorm8_trunc(b3) << 24);
#endif
}
DXB_DEV uint32_t dxb_unorm8_scalar(float v)      // scalar R8/A8 path: std::max(std::min(v,1),0)
{
    float b = v + (0.5f / 255.0f);
    b = (1.0f < b) ? 1.0f : b;
    b = (b < 0.0f) ? 0.0f : b;
    return dxb_f2u_trunc_small(fmaxf(b, 0.0f) * 255.0f);      // fmaxf also maps NaN to 0 like the cast did
}
DXB_DEV float dxb_stdclamp(float v, float lo, float hi)   // std::max(std::min(v, hi), lo)
{
    v = (hi < v) ? hi : v;
    v = (v < lo) ? lo : v;
    return v;
}

// Store pixel `i` of a row starting at `row`.  `threshold`: alpha threshold of the 1-bit alpha format (StoreScanline's
// last parameter, default 0; Convert passes the caller's, DirectXTexConvert.cpp:2116-2140).
DXB_DEV void dxb_store_pixel(uint32_t fmt, uint8_t* row, size_t i, dxb_px v, float threshold = 0.0f)
{
    switch (fmt)
    {
    case DXB_FMT_R11G11B10_FLOAT:        // XMStoreFloat3PK (:1756-1767)
        ((uint32_t*)row)[i] = (dxb_smallfloat_encode(v.x, 6u) & 0x7FFu) | ((dxb_smallfloat_encode(v.y, 6u) & 0x7FFu) << 11) | ((dxb_smallfloat_encode(v.z, 5u) & 0x3FFu) << 22);
        return;
    case DXB_FMT_R9G9B9E5_SHAREDEXP:     // XMStoreFloat3SE (:2057-2068)
    {
        const float maxf9 = (float)(0x1FF << 7), minf9 = 1.0f / (float)(1 << 16);
        const float x = (v.x >= 0.0f) ? ((v.x > maxf9) ? maxf9 : v.x) : 0.0f;
        const float y = (v.y >= 0.0f) ? ((v.y > maxf9) ? maxf9 : v.y) : 0.0f;
        const float z = (v.z >= 0.0f) ? ((v.z > maxf9) ? maxf9 : v.z) : 0.0f;
        const float mxy = (x > y) ? x : y, mxyz = (mxy > z) ? mxy : z;
        const float maxColor = (mxyz > minf9) ? mxyz : minf9;
        const uint32_t mi = dxb_float_as_uint(maxColor) + 0x00004000u;
        const uint32_t ex = mi >> 23;
        const float scaleR = dxb_uint_as_float(0x83000000u - (ex << 23));
        ((uint32_t*)row)[i] = ((uint32_t)dxb_lround(x * scaleR) & 0x1FFu) | (((uint32_t)dxb_lround(y * scaleR) & 0x1FFu) << 9)
                            | (((uint32_t)dxb_lround(z * scaleR) & 0x1FFu) << 18) | (((ex - 0x6fu) & 0x1Fu) << 27);
        return;
    }
    case DXB_FMT_B5G6R5_UNORM:           // swizzle to BGR, * {31, 63, 31}, XMStoreU565 (:2096-2114)
        ((uint16_t*)row)[i] = (uint16_t)(((dxb_store_int_rne(v.x * 31.0f, 31.0f) & 0x1Fu) << 11) | ((dxb_store_int_rne(v.y * 63.0f, 63.0f) & 0x3Fu) << 5) | (dxb_store_int_rne(v.z * 31.0f, 31.0f) & 0x1Fu));
        return;
    case DXB_FMT_B5G5R5A1_UNORM:         // * 31, XMStoreU555, alpha bit = (alpha > threshold) (:2116-2140)
        ((uint16_t*)row)[i] = (uint16_t)(((v.w > threshold) ? 0x8000u : 0u) | ((dxb_store_int_rne(v.x * 31.0f, 31.0f) & 0x1Fu) << 10)
                                         | ((dxb_store_int_rne(v.y * 31.0f, 31.0f) & 0x1Fu) << 5) | (dxb_store_int_rne(v.z * 31.0f, 31.0f) & 0x1Fu));
        return;
    case DXB_FMT_B4G4R4A4_UNORM:         // swizzle, * 15, XMStoreUNibble4 (:2399-2417)
        ((uint16_t*)row)[i] = (uint16_t)(((dxb_store_int_rne(v.w * 15.0f, 15.0f) & 0xFu) << 12) | ((dxb_store_int_rne(v.x * 15.0f, 15.0f) & 0xFu) << 8)
                                         | ((dxb_store_int_rne(v.y * 15.0f, 15.0f) & 0xFu) << 4) | (dxb_store_int_rne(v.z * 15.0f, 15.0f) & 0xFu));
        return;
    case DXB_FMT_R32G32B32A32_FLOAT:
    {
        float* p = (float*)row + i * 4; p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; return;
    }
    case DXB_FMT_R32G32B32_FLOAT:
    {
        float* p = (float*)row + i * 3; p[0] = v.x; p[1] = v.y; p[2] = v.z; return;
    }
    case DXB_FMT_R16G16B16A16_FLOAT:
    {
        uint16_t* p = (uint16_t*)row + i * 4;
        p[0] = dxb_float_to_half(dxb_clamp(v.x, -65504.0f, 65504.0f)); p[1] = dxb_float_to_half(dxb_clamp(v.y, -65504.0f, 65504.0f));
        p[2] = dxb_float_to_half(dxb_clamp(v.z, -65504.0f, 65504.0f)); p[3] = dxb_float_to_half(dxb_clamp(v.w, -65504.0f, 65504.0f));
        return;
    }
    case DXB_FMT_R16G16B16A16_UNORM:
    {
        uint16_t* p = (uint16_t*)row + i * 4;
        p[0] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.x, 0.0f, 1.0f) * 65535.0f); p[1] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.y, 0.0f, 1.0f) * 65535.0f);
        p[2] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.z, 0.0f, 1.0f) * 65535.0f); p[3] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.w, 0.0f, 1.0f) * 65535.0f);
        return;
    }
    case DXB_FMT_R16G16B16A16_SNORM:
    {
        int16_t* p = (int16_t*)row + i * 4;
        p[0] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.x, -1.0f, 1.0f) * 32767.0f); p[1] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.y, -1.0f, 1.0f) * 32767.0f);
        p[2] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.z, -1.0f, 1.0f) * 32767.0f); p[3] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.w, -1.0f, 1.0f) * 32767.0f);
        return;
    }
    case DXB_FMT_R32G32_FLOAT:
    {
        float* p = (float*)row + i * 2; p[0] = v.x; p[1] = v.y; return;
    }
    case DXB_FMT_R10G10B10A2_UNORM:
    {
        const uint32_t x = dxb_f2u_trunc_small(dxb_clamp(v.x, 0.0f, 1.0f) * 1023.0f), y = dxb_f2u_trunc_small(dxb_clamp(v.y, 0.0f, 1.0f) * 1023.0f);
        const uint32_t z = dxb_f2u_trunc_small(dxb_clamp(v.z, 0.0f, 1.0f) * 1023.0f), w = dxb_f2u_trunc_small(dxb_clamp(v.w, 0.0f, 1.0f) * 3.0f);
        ((uint32_t*)row)[i] = (w << 30) | ((z & 0x3FF) << 20) | ((y & 0x3FF) << 10) | (x & 0x3FF);
        return;
    }
    case DXB_FMT_R8G8B8A8_UNORM:
    case DXB_FMT_R8G8B8A8_UNORM_SRGB:
        ((uint32_t*)row)[i] = dxb_pack_unorm8x4(v.x, v.y, v.z, v.w);
        return;
    case DXB_FMT_B8G8R8A8_UNORM:
    case DXB_FMT_B8G8R8A8_UNORM_SRGB:
        ((uint32_t*)row)[i] = dxb_pack_unorm8x4(v.z, v.y, v.x, v.w);
        return;
    case DXB_FMT_B8G8R8X8_UNORM:
    case DXB_FMT_B8G8R8X8_UNORM_SRGB:
        ((uint32_t*)row)[i] = dxb_pack_unorm8x4(v.z, v.y, v.x, 1.0f);
        return;
    case DXB_FMT_R8G8B8A8_SNORM:
    {
        int8_t* p = (int8_t*)row + i * 4;
        p[0] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.x, -1.0f, 1.0f) * 127.0f); p[1] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.y, -1.0f, 1.0f) * 127.0f);
        p[2] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.z, -1.0f, 1.0f) * 127.0f); p[3] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.w, -1.0f, 1.0f) * 127.0f);
        return;
    }
    case DXB_FMT_R16G16_FLOAT:
    {
        uint16_t* p = (uint16_t*)row + i * 2;
        p[0] = dxb_float_to_half(dxb_clamp(v.x, -65504.0f, 65504.0f)); p[1] = dxb_float_to_half(dxb_clamp(v.y, -65504.0f, 65504.0f));
        return;
    }
    case DXB_FMT_R16G16_UNORM:
    {
        uint16_t* p = (uint16_t*)row + i * 2;
        p[0] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.x, 0.0f, 1.0f) * 65535.0f); p[1] = (uint16_t)dxb_f2i_rn_small(dxb_clamp(v.y, 0.0f, 1.0f) * 65535.0f);
        return;
    }
    case DXB_FMT_R16G16_SNORM:
    {
        int16_t* p = (int16_t*)row + i * 2;
        p[0] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.x, -1.0f, 1.0f) * 32767.0f); p[1] = (int16_t)dxb_f2i_rn_small(dxb_clamp(v.y, -1.0f, 1.0f) * 32767.0f);
        return;
    }
    case DXB_FMT_R32_FLOAT:
        ((float*)row)[i] = v.x; return;
    case DXB_FMT_R8G8_UNORM:
    {
        uint8_t* p = row + i * 2;
        p[0] = (uint8_t)dxb_f2i_rn_small(dxb_clamp(v.x, 0.0f, 1.0f) * 255.0f); p[1] = (uint8_t)dxb_f2i_rn_small(dxb_clamp(v.y, 0.0f, 1.0f) * 255.0f);
        return;
    }
    case DXB_FMT_R8G8_SNORM:
    {
        int8_t* p = (int8_t*)row + i * 2;
        p[0] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.x, -1.0f, 1.0f) * 127.0f); p[1] = (int8_t)dxb_f2i_rn_small(dxb_clamp(v.y, -1.0f, 1.0f) * 127.0f);
        return;
    }
    case DXB_FMT_R16_FLOAT:
        ((uint16_t*)row)[i] = dxb_float_to_half(dxb_stdclamp(v.x, -65504.0f, 65504.0f)); return;
    case DXB_FMT_R16_UNORM:
    {
        const float c = dxb_stdclamp(v.x, 0.0f, 1.0f);
        const float s = c * 65535.0f;
        ((uint16_t*)row)[i] = (uint16_t)dxb_f2i(s + 0.5f); return;
    }
    case DXB_FMT_R16_SNORM:
        ((int16_t*)row)[i] = (int16_t)dxb_lround(dxb_stdclamp(v.x, -1.0f, 1.0f) * 32767.0f); return;
    case DXB_FMT_R8_UNORM:
        row[i] = (uint8_t)dxb_unorm8_scalar(v.x); return;
    case DXB_FMT_R8_SNORM:
        ((int8_t*)row)[i] = (int8_t)dxb_lround(dxb_stdclamp(v.x, -1.0f, 1.0f) * 127.0f); return;
    case DXB_FMT_A8_UNORM:
        row[i] = (uint8_t)dxb_unorm8_scalar(v.w); return;
    default:
        return;
    }
}

// ---------------------------------------------------------------------------------------------
// Ordered-dither store (StoreScanlineDither with pDiffusionErrors == nullptr, DirectXTexConvert.cpp:4049-4567, macros
// STORE_SCANLINE / STORE_SCANLINE2 / STORE_SCANLINE1 :3895-4045): clamp, scale, add the 4x4 matrix entry of (x & 3, y & 3),
// round to nearest even, clamp to the code range, truncate to the integer type.  z = 0 for 2D images.
// Formats without a dither case fall through to the plain store (default: :4558-4559); returns nothing either way.
#if DXB_ON_DEVICE
static __device__ const float dxb_dither_matrix[32] =
#else
static const float dxb_dither_matrix[32] =
#endif
{   // index = (z & 3) + (y & 3) * 8 + (x & 3)   (:3863-3870)
    0.468750f, -0.031250f, 0.343750f, -0.156250f, 0.468750f, -0.031250f, 0.343750f, -0.156250f,
    -0.281250f, 0.218750f, -0.406250f, 0.093750f, -0.281250f, 0.218750f, -0.406250f, 0.093750f,
    0.281250f, -0.218750f, 0.406250f, -0.093750f, 0.281250f, -0.218750f, 0.406250f, -0.093750f,
    -0.468750f, 0.031250f, -0.343750f, 0.156250f, -0.468750f, 0.031250f, -0.343750f, 0.156250f,
};
DXB_DEV float dxb_round_even(float f)
{
#if DXB_ON_DEVICE
    return rintf(f);
#else
    return nearbyintf(f);
#endif
}
// per-format dither parameters: code range per (stored) channel, unsigned vs signed-normalised, BGR store order
struct dxb_dither_fmt { float sx, sy, sz, sw; bool ok, clampzero, bgr; };
DXB_DEV dxb_dither_fmt dxb_dither_format(uint32_t fmt)
{
    dxb_dither_fmt f; f.ok = true; f.clampzero = true; f.bgr = false; f.sx = f.sy = f.sz = f.sw = 255.0f;
    switch (fmt)
    {
    case DXB_FMT_R16G16B16A16_UNORM: case DXB_FMT_R16G16_UNORM: case DXB_FMT_R16_UNORM: f.sx = f.sy = f.sz = f.sw = 65535.0f; break;
    case DXB_FMT_R16G16B16A16_SNORM: case DXB_FMT_R16G16_SNORM: case DXB_FMT_R16_SNORM: f.sx = f.sy = f.sz = f.sw = 32767.0f; f.clampzero = false; break;
    case DXB_FMT_R10G10B10A2_UNORM: f.sx = f.sy = f.sz = 1023.0f; f.sw = 3.0f; break;
    case DXB_FMT_R8G8B8A8_UNORM: case DXB_FMT_R8G8B8A8_UNORM_SRGB: case DXB_FMT_R8G8_UNORM: case DXB_FMT_R8_UNORM: case DXB_FMT_A8_UNORM: break;
    case DXB_FMT_R8G8B8A8_SNORM: case DXB_FMT_R8G8_SNORM: case DXB_FMT_R8_SNORM: f.sx = f.sy = f.sz = f.sw = 127.0f; f.clampzero = false; break;
    case DXB_FMT_B8G8R8A8_UNORM: case DXB_FMT_B8G8R8A8_UNORM_SRGB: case DXB_FMT_B8G8R8X8_UNORM: case DXB_FMT_B8G8R8X8_UNORM_SRGB: f.bgr = true; break;
    default: f.ok = false; break;
    }
    return f;
}
// integer codes c (already in STORE order: for BGR formats c.x = blue) -> memory
DXB_DEV void dxb_store_codes(uint32_t fmt, uint8_t* row, size_t i, int32_t cx, int32_t cy, int32_t cz, int32_t cw)
{
    switch (fmt)
    {
    case DXB_FMT_R16G16B16A16_UNORM: case DXB_FMT_R16G16B16A16_SNORM:
    { uint16_t* p = (uint16_t*)row + 4 * i; p[0] = (uint16_t)cx; p[1] = (uint16_t)cy; p[2] = (uint16_t)cz; p[3] = (uint16_t)cw; return; }
    case DXB_FMT_R10G10B10A2_UNORM:
        ((uint32_t*)row)[i] = ((uint32_t)cx & 0x3FFu) | (((uint32_t)cy & 0x3FFu) << 10) | (((uint32_t)cz & 0x3FFu) << 20) | (((uint32_t)cw & 0x3u) << 30); return;
    case DXB_FMT_R8G8B8A8_UNORM: case DXB_FMT_R8G8B8A8_UNORM_SRGB: case DXB_FMT_R8G8B8A8_SNORM:
    case DXB_FMT_B8G8R8A8_UNORM: case DXB_FMT_B8G8R8A8_UNORM_SRGB:
        ((uint32_t*)row)[i] = ((uint32_t)cx & 0xFFu) | (((uint32_t)cy & 0xFFu) << 8) | (((uint32_t)cz & 0xFFu) << 16) | (((uint32_t)cw & 0xFFu) << 24); return;
    case DXB_FMT_B8G8R8X8_UNORM: case DXB_FMT_B8G8R8X8_UNORM_SRGB:          // the X byte is written as 0 on the dither paths (:4446)
        ((uint32_t*)row)[i] = ((uint32_t)cx & 0xFFu) | (((uint32_t)cy & 0xFFu) << 8) | (((uint32_t)cz & 0xFFu) << 16); return;
    case DXB_FMT_R16G16_UNORM: case DXB_FMT_R16G16_SNORM: { uint16_t* p = (uint16_t*)row + 2 * i; p[0] = (uint16_t)cx; p[1] = (uint16_t)cy; return; }
    case DXB_FMT_R8G8_UNORM: case DXB_FMT_R8G8_SNORM: row[2 * i] = (uint8_t)cx; row[2 * i + 1] = (uint8_t)cy; return;
    case DXB_FMT_R16_UNORM: case DXB_FMT_R16_SNORM: ((uint16_t*)row)[i] = (uint16_t)cx; return;
    case DXB_FMT_R8_UNORM: case DXB_FMT_R8_SNORM: row[i] = (uint8_t)cx; return;
    case DXB_FMT_A8_UNORM: row[i] = (uint8_t)cw; return;
    default: return;
    }
}
// one channel of the ordered path: returns the integer code (two's complement for the signed formats)
DXB_DEV int32_t dxb_dither_code(float v, float scale, bool clampzero, float d)
{
    v = clampzero ? dxb_clamp(v, 0.0f, 1.0f) : dxb_clamp(v, -1.0f, 1.0f);
    v = v + 0.0f;                                    // + vError (zero without error diffusion)
    v = v * scale;
    float t = dxb_round_even(v + d);
    t = dxb_ssemin(scale, t);
    t = dxb_ssemax(clampzero ? 0.0f : (-scale + 1.0f), t);
    return dxb_f2i(t);
}
DXB_DEV void dxb_store_pixel_dither(uint32_t fmt, uint8_t* row, size_t i, uint32_t y, dxb_px v)
{
    const dxb_dither_fmt f = dxb_dither_format(fmt);
    if (!f.ok) { dxb_store_pixel(fmt, row, i, v); return; }
    const float d = dxb_dither_matrix[((y & 3u) << 3) + (uint32_t)(i & 3u)];
    if (f.bgr) { const float t = v.x; v.x = v.z; v.z = t; }
    dxb_store_codes(fmt, row, i, dxb_dither_code(v.x, f.sx, f.clampzero, d), dxb_dither_code(v.y, f.sy, f.clampzero, d),
                    dxb_dither_code(v.z, f.sz, f.clampzero, d), dxb_dither_code(v.w, f.sw, f.clampzero, d));
}

// ---------------------------------------------------------------------------------------------
// Error-diffusion (Floyd-Steinberg) store of one whole image (StoreScanlineDither with pDiffusionErrors, driven by the row
// loop of ConvertCustom :4815-4858).  Inherently serial: the quantisation error of a pixel goes to its successor in a
// serpentine scan and to three pixels of the next row, so ONE thread walks the image; the four channels are independent
// dependency chains.  E0 / E1 = two error rows of (width + 2) pixels (this row's incoming / the next row's outgoing errors).
// Reference quirks kept: the incoming errors (held in store order) are added to the source before the BGR swizzle.
DXB_DEV float dxb_dd1(float& carry, float s, float scale, bool clampzero, float& e3, float& e5, float& e1)
{
    float v = clampzero ? dxb_clamp(s, 0.0f, 1.0f) : dxb_clamp(s, -1.0f, 1.0f);
    v = v + carry;
    v = v * scale;
    float t = dxb_round_even(v);
    float err = v - t;
    err = err / scale;
    e3 = (3.0f / 16.0f) * err; e5 = (5.0f / 16.0f) * err; e1 = (1.0f / 16.0f) * err;
    carry = err * (7.0f / 16.0f);
    t = dxb_ssemin(scale, t);
    t = dxb_ssemax(clampzero ? 0.0f : (-scale + 1.0f), t);
    return t;
}
DXB_DEV void dxb_convert_diffuse_image(uint32_t srcFmt, uint32_t dstFmt, uint32_t inF, uint32_t outF, uint32_t flags,
                                       const uint8_t* src, size_t srcPitch, uint8_t* dst, size_t dstPitch, uint32_t width, uint32_t height,
                                       dxb_px* E0, dxb_px* E1)
{
    const dxb_dither_fmt f = dxb_dither_format(dstFmt);
    for (uint32_t i = 0; i < width + 2u; ++i) { E0[i] = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f); E1[i] = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f); }
    for (uint32_t y = 0; y < height; ++y)
    {
        dxb_px* Ein = (y & 1u) ? E1 : E0;            // errors handed down by the previous row
        dxb_px* Eout = (y & 1u) ? E0 : E1;           // errors for the next row (zero on entry)
        const uint8_t* srow = src + (size_t)y * srcPitch;
        uint8_t* drow = dst + (size_t)y * dstPitch;
        dxb_px carry = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f);
        const int delta = (y & 1u) ? -2 : 0;
        for (uint32_t k = 0; k < width; ++k)
        {
            const uint32_t index = (y & 1u) ? (width - 1u - k) : k;
            dxb_px v = dxb_convert_pixel(dxb_load_pixel(srcFmt, srow, index), inF, outF, flags);
            if (!f.ok) { dxb_store_pixel(dstFmt, drow, index, v); continue; }      // formats without a dither case: plain store (:4558)
            const dxb_px e = Ein[index + 1u];
            v = dxb_make_px(v.x + e.x, v.y + e.y, v.z + e.z, v.w + e.w);
            if (f.bgr) { const float t = v.x; v.x = v.z; v.z = t; }
            dxb_px a, b, c;                                 // 3/16, 5/16, 1/16 shares
            const float tx = dxb_dd1(carry.x, v.x, f.sx, f.clampzero, a.x, b.x, c.x);
            const float ty = dxb_dd1(carry.y, v.y, f.sy, f.clampzero, a.y, b.y, c.y);
            const float tz = dxb_dd1(carry.z, v.z, f.sz, f.clampzero, a.z, b.z, c.z);
            const float tw = dxb_dd1(carry.w, v.w, f.sw, f.clampzero, a.w, b.w, c.w);
            dxb_px* p3 = Eout + ((int)index - delta); dxb_px* p5 = Eout + (index + 1u); dxb_px* p1 = Eout + ((int)index + 2 + delta);
            *p3 = dxb_make_px(a.x + p3->x, a.y + p3->y, a.z + p3->z, a.w + p3->w);
            *p5 = dxb_make_px(b.x + p5->x, b.y + p5->y, b.z + p5->z, b.w + p5->w);
            *p1 = dxb_make_px(c.x + p1->x, c.y + p1->y, c.z + p1->z, c.w + p1->w);
            dxb_store_codes(dstFmt, drow, index, dxb_f2i(tx), dxb_f2i(ty), dxb_f2i(tz), dxb_f2i(tw));
        }
        for (uint32_t i = 0; i < width + 2u; ++i) Ein[i] = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f);
    }
}
