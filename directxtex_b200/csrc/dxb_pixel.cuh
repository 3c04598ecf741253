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
DXB_DEV dxb_px dxb_load_pixel(uint32_t fmt, const uint8_t* row, size_t i)
{
    switch (fmt)
    {
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
                if (flags & DXB_FILTER_FLOAT_X2BIAS)
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
                v.x = dxb_clamp(v.x, -1.0f, 1.0f); v.y = dxb_clamp(v.y, -1.0f, 1.0f); v.z = dxb_clamp(v.z, -1.0f, 1.0f); v.w = dxb_clamp(v.w, -1.0f, 1.0f);
            }
        }
        else if (diff & DXB_CONVF_UNORM)
        {
            if ((outF & DXB_CONVF_FLOAT) && (flags & DXB_FILTER_FLOAT_X2BIAS))
            {
                v.x = dxb_madd(v.x, 2.0f, -1.0f); v.y = dxb_madd(v.y, 2.0f, -1.0f); v.z = dxb_madd(v.z, 2.0f, -1.0f); v.w = dxb_madd(v.w, 2.0f, -1.0f);
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

// Store pixel `i` of a row starting at `row`.
DXB_DEV void dxb_store_pixel(uint32_t fmt, uint8_t* row, size_t i, dxb_px v)
{
    switch (fmt)
    {
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
// one channel: returns the integer code (as int32, two's complement for the signed formats)
DXB_DEV int32_t dxb_dither_code(float v, float scale, bool clampzero, float d)
{
    // norm is true for every format implemented here
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
    const float d = dxb_dither_matrix[((y & 3u) << 3) + (uint32_t)(i & 3u)];
    switch (fmt)
    {
    case DXB_FMT_R16G16B16A16_UNORM:
    {
        uint16_t* p = (uint16_t*)row + 4 * i;
        p[0] = (uint16_t)dxb_dither_code(v.x, 65535.0f, true, d); p[1] = (uint16_t)dxb_dither_code(v.y, 65535.0f, true, d);
        p[2] = (uint16_t)dxb_dither_code(v.z, 65535.0f, true, d); p[3] = (uint16_t)dxb_dither_code(v.w, 65535.0f, true, d);
        return;
    }
    case DXB_FMT_R16G16B16A16_SNORM:
    {
        int16_t* p = (int16_t*)row + 4 * i;
        p[0] = (int16_t)dxb_dither_code(v.x, 32767.0f, false, d); p[1] = (int16_t)dxb_dither_code(v.y, 32767.0f, false, d);
        p[2] = (int16_t)dxb_dither_code(v.z, 32767.0f, false, d); p[3] = (int16_t)dxb_dither_code(v.w, 32767.0f, false, d);
        return;
    }
    case DXB_FMT_R10G10B10A2_UNORM:
    {
        const uint32_t x = (uint32_t)dxb_dither_code(v.x, 1023.0f, true, d) & 0x3FFu, yy = (uint32_t)dxb_dither_code(v.y, 1023.0f, true, d) & 0x3FFu;
        const uint32_t z = (uint32_t)dxb_dither_code(v.z, 1023.0f, true, d) & 0x3FFu, w = (uint32_t)dxb_dither_code(v.w, 3.0f, true, d) & 0x3u;
        ((uint32_t*)row)[i] = x | (yy << 10) | (z << 20) | (w << 30);
        return;
    }
    case DXB_FMT_R8G8B8A8_UNORM: case DXB_FMT_R8G8B8A8_UNORM_SRGB:
        ((uint32_t*)row)[i] = ((uint32_t)dxb_dither_code(v.x, 255.0f, true, d) & 0xFFu) | (((uint32_t)dxb_dither_code(v.y, 255.0f, true, d) & 0xFFu) << 8) |
                              (((uint32_t)dxb_dither_code(v.z, 255.0f, true, d) & 0xFFu) << 16) | (((uint32_t)dxb_dither_code(v.w, 255.0f, true, d) & 0xFFu) << 24);
        return;
    case DXB_FMT_B8G8R8A8_UNORM: case DXB_FMT_B8G8R8A8_UNORM_SRGB:
        ((uint32_t*)row)[i] = ((uint32_t)dxb_dither_code(v.z, 255.0f, true, d) & 0xFFu) | (((uint32_t)dxb_dither_code(v.y, 255.0f, true, d) & 0xFFu) << 8) |
                              (((uint32_t)dxb_dither_code(v.x, 255.0f, true, d) & 0xFFu) << 16) | (((uint32_t)dxb_dither_code(v.w, 255.0f, true, d) & 0xFFu) << 24);
        return;
    case DXB_FMT_B8G8R8X8_UNORM: case DXB_FMT_B8G8R8X8_UNORM_SRGB:          // the X byte is written as 0 on this path (:4446)
        ((uint32_t*)row)[i] = ((uint32_t)dxb_dither_code(v.z, 255.0f, true, d) & 0xFFu) | (((uint32_t)dxb_dither_code(v.y, 255.0f, true, d) & 0xFFu) << 8) |
                              (((uint32_t)dxb_dither_code(v.x, 255.0f, true, d) & 0xFFu) << 16);
        return;
    case DXB_FMT_R8G8B8A8_SNORM:
        ((uint32_t*)row)[i] = ((uint32_t)dxb_dither_code(v.x, 127.0f, false, d) & 0xFFu) | (((uint32_t)dxb_dither_code(v.y, 127.0f, false, d) & 0xFFu) << 8) |
                              (((uint32_t)dxb_dither_code(v.z, 127.0f, false, d) & 0xFFu) << 16) | (((uint32_t)dxb_dither_code(v.w, 127.0f, false, d) & 0xFFu) << 24);
        return;
    case DXB_FMT_R16G16_UNORM:
    {
        uint16_t* p = (uint16_t*)row + 2 * i;
        p[0] = (uint16_t)dxb_dither_code(v.x, 65535.0f, true, d); p[1] = (uint16_t)dxb_dither_code(v.y, 65535.0f, true, d);
        return;
    }
    case DXB_FMT_R16G16_SNORM:
    {
        int16_t* p = (int16_t*)row + 2 * i;
        p[0] = (int16_t)dxb_dither_code(v.x, 32767.0f, false, d); p[1] = (int16_t)dxb_dither_code(v.y, 32767.0f, false, d);
        return;
    }
    case DXB_FMT_R8G8_UNORM:
        row[2 * i] = (uint8_t)dxb_dither_code(v.x, 255.0f, true, d); row[2 * i + 1] = (uint8_t)dxb_dither_code(v.y, 255.0f, true, d);
        return;
    case DXB_FMT_R8G8_SNORM:
        row[2 * i] = (uint8_t)(int8_t)dxb_dither_code(v.x, 127.0f, false, d); row[2 * i + 1] = (uint8_t)(int8_t)dxb_dither_code(v.y, 127.0f, false, d);
        return;
    case DXB_FMT_R16_UNORM: ((uint16_t*)row)[i] = (uint16_t)dxb_dither_code(v.x, 65535.0f, true, d); return;
    case DXB_FMT_R16_SNORM: ((int16_t*)row)[i] = (int16_t)dxb_dither_code(v.x, 32767.0f, false, d); return;
    case DXB_FMT_R8_UNORM: row[i] = (uint8_t)dxb_dither_code(v.x, 255.0f, true, d); return;
    case DXB_FMT_R8_SNORM: row[i] = (uint8_t)(int8_t)dxb_dither_code(v.x, 127.0f, false, d); return;
    case DXB_FMT_A8_UNORM: row[i] = (uint8_t)dxb_dither_code(v.w, 255.0f, true, d); return;
    default: dxb_store_pixel(fmt, row, i, v); return;
    }
}
