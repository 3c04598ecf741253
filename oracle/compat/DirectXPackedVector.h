// oracle/compat/DirectXPackedVector.h — TEST INFRASTRUCTURE (oracle build only).
// Packed-pixel load/store surface used by DirectXTexConvert.cpp's LoadScanline /
// StoreScanline switch (DirectXTexConvert.cpp:798-1628, 1671-2530).  Scalar
// restatement of DirectXMath's x64 SSE2 path; (M) = restated from memory, see
// DirectXMath.h in this directory.  Summary of the rounding rules restated here:
//   *N loads   : float(int) * (1.0f / MAX)  (multiply by the fp32 reciprocal)
//   UByteN4 / UDecN4 / XDecN4 stores : saturate, scale, TRUNCATE  (cvttps + mask trick)
//   all other normalised / integer stores : clamp, scale, round-to-nearest-even (cvtps)
//   half : IEEE binary16, round-to-nearest-even, denormals kept, overflow -> Inf
#pragma once
#include "DirectXMath.h"

namespace DirectX
{
namespace PackedVector
{
    typedef uint16_t HALF;

    struct XMHALF2 { HALF x, y; };
    struct XMHALF4 { HALF x, y, z, w; };
    struct XMSHORTN2 { int16_t x, y; };
    struct XMSHORT2 { int16_t x, y; };
    struct XMUSHORTN2 { uint16_t x, y; };
    struct XMUSHORT2 { uint16_t x, y; };
    struct XMBYTEN2 { int8_t x, y; };
    struct XMBYTE2 { int8_t x, y; };
    struct XMUBYTEN2 { uint8_t x, y; };
    struct XMUBYTE2 { uint8_t x, y; };
    struct XMSHORTN4 { int16_t x, y, z, w; };
    struct XMSHORT4 { int16_t x, y, z, w; };
    struct XMUSHORTN4 { uint16_t x, y, z, w; };
    struct XMUSHORT4 { uint16_t x, y, z, w; };
    struct XMBYTEN4 { int8_t x, y, z, w; };
    struct XMBYTE4 { int8_t x, y, z, w; };
    struct XMUBYTEN4 { union { struct { uint8_t x, y, z, w; }; uint32_t v; }; };
    struct XMUBYTE4 { union { struct { uint8_t x, y, z, w; }; uint32_t v; }; };
    struct XMU565 { union { struct { uint16_t x : 5; uint16_t y : 6; uint16_t z : 5; }; uint16_t v; }; };
    struct XMU555 { union { struct { uint16_t x : 5; uint16_t y : 5; uint16_t z : 5; uint16_t w : 1; }; uint16_t v; }; };
    struct XMUNIBBLE4 { union { struct { uint16_t x : 4; uint16_t y : 4; uint16_t z : 4; uint16_t w : 4; }; uint16_t v; }; };
    struct XMUDECN4 { union { struct { uint32_t x : 10; uint32_t y : 10; uint32_t z : 10; uint32_t w : 2; }; uint32_t v; }; };
    struct XMUDEC4 { union { struct { uint32_t x : 10; uint32_t y : 10; uint32_t z : 10; uint32_t w : 2; }; uint32_t v; }; };
    struct XMXDECN4 { union { struct { int32_t x : 10; int32_t y : 10; int32_t z : 10; uint32_t w : 2; }; uint32_t v; }; };
    struct XMFLOAT3PK { union { struct { uint32_t xm : 6; uint32_t xe : 5; uint32_t ym : 6; uint32_t ye : 5; uint32_t zm : 5; uint32_t ze : 5; }; uint32_t v; }; };
    struct XMFLOAT3SE { union { struct { uint32_t xm : 9; uint32_t ym : 9; uint32_t zm : 9; uint32_t e : 5; }; uint32_t v; }; };

    //---------------------------------------------------------------------------------
    // half <-> float: IEEE binary16 (software path of DirectXMath; no F16C)
    inline float XMConvertHalfToFloat(HALF h) noexcept
    {
        uint32_t mant = h & 0x03FFu;
        uint32_t exp = (h & 0x7C00u);
        uint32_t out;
        if (exp == 0x7C00u)          // Inf / NaN
            exp = 0x8Fu;
        else if (exp != 0)
            exp = (h >> 10) & 0x1Fu;
        else if (mant != 0)          // denormal -> normalise
        {
            exp = 1;
            do { exp--; mant <<= 1; } while ((mant & 0x0400u) == 0);
            mant &= 0x03FFu;
        }
        else
            exp = static_cast<uint32_t>(-112);
        out = ((h & 0x8000u) << 16) | ((exp + 112u) << 23) | (mant << 13);
        float f; memcpy(&f, &out, 4); return f;
    }
    inline HALF XMConvertFloatToHalf(float f) noexcept
    {
        uint32_t iv; memcpy(&iv, &f, 4);
        const uint32_t sign = (iv & 0x80000000u) >> 16;
        iv &= 0x7FFFFFFFu;
        uint32_t r;
        if (iv >= 0x47800000u)       // too large: Inf or NaN
            r = 0x7C00u | ((iv > 0x7F800000u) ? (0x200u | ((iv >> 13) & 0x3FFu)) : 0u);
        else if (iv <= 0x33000000u)  // too small: 0
            r = 0;
        else if (iv < 0x38800000u)   // denormal half
        {
            const uint32_t shift = 125u - (iv >> 23);
            iv = 0x800000u | (iv & 0x7FFFFFu);
            r = iv >> (shift + 1);
            const uint32_t s = (iv & ((1u << shift) - 1)) != 0;
            r += (r | s) & ((iv >> shift) & 1u);
        }
        else
        {
            iv += 0xC8000000u;
            r = ((iv + 0x0FFFu + ((iv >> 13) & 1u)) >> 13) & 0x7FFFu;
        }
        return static_cast<HALF>(r | sign);
    }
    inline float* XMConvertHalfToFloatStream(float* out, size_t outStride, const HALF* in, size_t inStride, size_t count) noexcept
    {
        auto pi = reinterpret_cast<const uint8_t*>(in); auto po = reinterpret_cast<uint8_t*>(out);
        for (size_t i = 0; i < count; ++i) { HALF h; memcpy(&h, pi, 2); float f = XMConvertHalfToFloat(h); memcpy(po, &f, 4); pi += inStride; po += outStride; }
        return out;
    }
    inline HALF* XMConvertFloatToHalfStream(HALF* out, size_t outStride, const float* in, size_t inStride, size_t count) noexcept
    {
        auto pi = reinterpret_cast<const uint8_t*>(in); auto po = reinterpret_cast<uint8_t*>(out);
        for (size_t i = 0; i < count; ++i) { float f; memcpy(&f, pi, 4); HALF h = XMConvertFloatToHalf(f); memcpy(po, &h, 2); pi += inStride; po += outStride; }
        return out;
    }

    inline XMVECTOR XMLoadHalf2(const XMHALF2* p) noexcept { return shim::mk(XMConvertHalfToFloat(p->x), XMConvertHalfToFloat(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadHalf4(const XMHALF4* p) noexcept { return shim::mk(XMConvertHalfToFloat(p->x), XMConvertHalfToFloat(p->y), XMConvertHalfToFloat(p->z), XMConvertHalfToFloat(p->w)); }
    inline void XMStoreHalf2(XMHALF2* p, FXMVECTOR V) noexcept { p->x = XMConvertFloatToHalf(V.vector4_f32[0]); p->y = XMConvertFloatToHalf(V.vector4_f32[1]); }
    inline void XMStoreHalf4(XMHALF4* p, FXMVECTOR V) noexcept { p->x = XMConvertFloatToHalf(V.vector4_f32[0]); p->y = XMConvertFloatToHalf(V.vector4_f32[1]); p->z = XMConvertFloatToHalf(V.vector4_f32[2]); p->w = XMConvertFloatToHalf(V.vector4_f32[3]); }

    //---------------------------------------------------------------------------------
    namespace pk
    {
        inline float clampf(float v, float lo, float hi) noexcept { return shim::ssemin(shim::ssemax(v, lo), hi); }
        inline float snorm(int v, float rcp) noexcept { const float f = static_cast<float>(v) * rcp; return shim::ssemax(f, -1.0f); }
        inline int32_t st_norm(float v, float lo, float scale) noexcept { return shim::rne(clampf(v, lo, 1.0f) * scale); }
        inline int32_t st_int(float v, float lo, float hi) noexcept { return shim::rne(clampf(v, lo, hi)); }
    }

    // 16-bit
    inline XMVECTOR XMLoadShortN2(const XMSHORTN2* p) noexcept { return shim::mk(pk::snorm(p->x, 1.0f / 32767.0f), pk::snorm(p->y, 1.0f / 32767.0f), 0.f, 0.f); }
    inline XMVECTOR XMLoadShortN4(const XMSHORTN4* p) noexcept { return shim::mk(pk::snorm(p->x, 1.0f / 32767.0f), pk::snorm(p->y, 1.0f / 32767.0f), pk::snorm(p->z, 1.0f / 32767.0f), pk::snorm(p->w, 1.0f / 32767.0f)); }
    inline XMVECTOR XMLoadShort2(const XMSHORT2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadShort4(const XMSHORT4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }
    inline XMVECTOR XMLoadUShortN2(const XMUSHORTN2* p) noexcept { return shim::mk(float(p->x) * (1.0f / 65535.0f), float(p->y) * (1.0f / 65535.0f), 0.f, 0.f); }
    inline XMVECTOR XMLoadUShortN4(const XMUSHORTN4* p) noexcept { return shim::mk(float(p->x) * (1.0f / 65535.0f), float(p->y) * (1.0f / 65535.0f), float(p->z) * (1.0f / 65535.0f), float(p->w) * (1.0f / 65535.0f)); }
    inline XMVECTOR XMLoadUShort2(const XMUSHORT2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadUShort4(const XMUSHORT4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }

    inline void XMStoreShortN2(XMSHORTN2* p, FXMVECTOR V) noexcept { p->x = int16_t(pk::st_norm(V.vector4_f32[0], -1.f, 32767.f)); p->y = int16_t(pk::st_norm(V.vector4_f32[1], -1.f, 32767.f)); }
    inline void XMStoreShortN4(XMSHORTN4* p, FXMVECTOR V) noexcept { p->x = int16_t(pk::st_norm(V.vector4_f32[0], -1.f, 32767.f)); p->y = int16_t(pk::st_norm(V.vector4_f32[1], -1.f, 32767.f)); p->z = int16_t(pk::st_norm(V.vector4_f32[2], -1.f, 32767.f)); p->w = int16_t(pk::st_norm(V.vector4_f32[3], -1.f, 32767.f)); }
    inline void XMStoreShort2(XMSHORT2* p, FXMVECTOR V) noexcept { p->x = int16_t(pk::st_int(V.vector4_f32[0], -32767.f, 32767.f)); p->y = int16_t(pk::st_int(V.vector4_f32[1], -32767.f, 32767.f)); }
    inline void XMStoreShort4(XMSHORT4* p, FXMVECTOR V) noexcept { p->x = int16_t(pk::st_int(V.vector4_f32[0], -32767.f, 32767.f)); p->y = int16_t(pk::st_int(V.vector4_f32[1], -32767.f, 32767.f)); p->z = int16_t(pk::st_int(V.vector4_f32[2], -32767.f, 32767.f)); p->w = int16_t(pk::st_int(V.vector4_f32[3], -32767.f, 32767.f)); }
    inline void XMStoreUShortN2(XMUSHORTN2* p, FXMVECTOR V) noexcept { p->x = uint16_t(pk::st_norm(V.vector4_f32[0], 0.f, 65535.f)); p->y = uint16_t(pk::st_norm(V.vector4_f32[1], 0.f, 65535.f)); }
    inline void XMStoreUShortN4(XMUSHORTN4* p, FXMVECTOR V) noexcept { p->x = uint16_t(pk::st_norm(V.vector4_f32[0], 0.f, 65535.f)); p->y = uint16_t(pk::st_norm(V.vector4_f32[1], 0.f, 65535.f)); p->z = uint16_t(pk::st_norm(V.vector4_f32[2], 0.f, 65535.f)); p->w = uint16_t(pk::st_norm(V.vector4_f32[3], 0.f, 65535.f)); }
    inline void XMStoreUShort2(XMUSHORT2* p, FXMVECTOR V) noexcept { p->x = uint16_t(pk::st_int(V.vector4_f32[0], 0.f, 65535.f)); p->y = uint16_t(pk::st_int(V.vector4_f32[1], 0.f, 65535.f)); }
    inline void XMStoreUShort4(XMUSHORT4* p, FXMVECTOR V) noexcept { p->x = uint16_t(pk::st_int(V.vector4_f32[0], 0.f, 65535.f)); p->y = uint16_t(pk::st_int(V.vector4_f32[1], 0.f, 65535.f)); p->z = uint16_t(pk::st_int(V.vector4_f32[2], 0.f, 65535.f)); p->w = uint16_t(pk::st_int(V.vector4_f32[3], 0.f, 65535.f)); }

    // 8-bit
    inline XMVECTOR XMLoadByteN2(const XMBYTEN2* p) noexcept { return shim::mk(pk::snorm(p->x, 1.0f / 127.0f), pk::snorm(p->y, 1.0f / 127.0f), 0.f, 0.f); }
    inline XMVECTOR XMLoadByteN4(const XMBYTEN4* p) noexcept { return shim::mk(pk::snorm(p->x, 1.0f / 127.0f), pk::snorm(p->y, 1.0f / 127.0f), pk::snorm(p->z, 1.0f / 127.0f), pk::snorm(p->w, 1.0f / 127.0f)); }
    inline XMVECTOR XMLoadByte2(const XMBYTE2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadByte4(const XMBYTE4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }
    inline XMVECTOR XMLoadUByteN2(const XMUBYTEN2* p) noexcept { return shim::mk(float(p->x) * (1.0f / 255.0f), float(p->y) * (1.0f / 255.0f), 0.f, 0.f); }
    // x64 SSE2 path: cvtepi32_ps then multiply by {1/255, 1/(255*256), ...}: == float(b) * (1.0f/255.0f) per channel (M)
    inline XMVECTOR XMLoadUByteN4(const XMUBYTEN4* p) noexcept { return shim::mk(float(p->x) * (1.0f / 255.0f), float(p->y) * (1.0f / 255.0f), float(p->z) * (1.0f / 255.0f), float(p->w) * (1.0f / 255.0f)); }
    inline XMVECTOR XMLoadUByte2(const XMUBYTE2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadUByte4(const XMUBYTE4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }

    inline void XMStoreByteN2(XMBYTEN2* p, FXMVECTOR V) noexcept { p->x = int8_t(pk::st_norm(V.vector4_f32[0], -1.f, 127.f)); p->y = int8_t(pk::st_norm(V.vector4_f32[1], -1.f, 127.f)); }
    inline void XMStoreByteN4(XMBYTEN4* p, FXMVECTOR V) noexcept { p->x = int8_t(pk::st_norm(V.vector4_f32[0], -1.f, 127.f)); p->y = int8_t(pk::st_norm(V.vector4_f32[1], -1.f, 127.f)); p->z = int8_t(pk::st_norm(V.vector4_f32[2], -1.f, 127.f)); p->w = int8_t(pk::st_norm(V.vector4_f32[3], -1.f, 127.f)); }
    inline void XMStoreByte2(XMBYTE2* p, FXMVECTOR V) noexcept { p->x = int8_t(pk::st_int(V.vector4_f32[0], -127.f, 127.f)); p->y = int8_t(pk::st_int(V.vector4_f32[1], -127.f, 127.f)); }
    inline void XMStoreByte4(XMBYTE4* p, FXMVECTOR V) noexcept { p->x = int8_t(pk::st_int(V.vector4_f32[0], -127.f, 127.f)); p->y = int8_t(pk::st_int(V.vector4_f32[1], -127.f, 127.f)); p->z = int8_t(pk::st_int(V.vector4_f32[2], -127.f, 127.f)); p->w = int8_t(pk::st_int(V.vector4_f32[3], -127.f, 127.f)); }
    inline void XMStoreUByteN2(XMUBYTEN2* p, FXMVECTOR V) noexcept { p->x = uint8_t(pk::st_norm(V.vector4_f32[0], 0.f, 255.f)); p->y = uint8_t(pk::st_norm(V.vector4_f32[1], 0.f, 255.f)); }
    // x64 SSE2 path: saturate, scale, cvttps (TRUNCATE) + mask (M).  DirectXTex adds its own
    // +0.5/255 bias before calling this (DirectXTexConvert.cpp:198-199, 1766-1767).
    inline void XMStoreUByteN4(XMUBYTEN4* p, FXMVECTOR V) noexcept
    {
        p->x = uint8_t(int32_t(pk::clampf(V.vector4_f32[0], 0.f, 1.f) * 255.0f));
        p->y = uint8_t(int32_t(pk::clampf(V.vector4_f32[1], 0.f, 1.f) * 255.0f));
        p->z = uint8_t(int32_t(pk::clampf(V.vector4_f32[2], 0.f, 1.f) * 255.0f));
        p->w = uint8_t(int32_t(pk::clampf(V.vector4_f32[3], 0.f, 1.f) * 255.0f));
    }
    inline void XMStoreUByte2(XMUBYTE2* p, FXMVECTOR V) noexcept { p->x = uint8_t(pk::st_int(V.vector4_f32[0], 0.f, 255.f)); p->y = uint8_t(pk::st_int(V.vector4_f32[1], 0.f, 255.f)); }
    inline void XMStoreUByte4(XMUBYTE4* p, FXMVECTOR V) noexcept { p->x = uint8_t(pk::st_int(V.vector4_f32[0], 0.f, 255.f)); p->y = uint8_t(pk::st_int(V.vector4_f32[1], 0.f, 255.f)); p->z = uint8_t(pk::st_int(V.vector4_f32[2], 0.f, 255.f)); p->w = uint8_t(pk::st_int(V.vector4_f32[3], 0.f, 255.f)); }

    // 5:6:5 / 5:5:5:1 / 4:4:4:4 — unnormalised field values
    inline XMVECTOR XMLoadU565(const XMU565* p) noexcept { return shim::mk(float(p->v & 0x1F), float((p->v >> 5) & 0x3F), float((p->v >> 11) & 0x1F), 0.f); }
    inline XMVECTOR XMLoadU555(const XMU555* p) noexcept { return shim::mk(float(p->v & 0x1F), float((p->v >> 5) & 0x1F), float((p->v >> 10) & 0x1F), float((p->v >> 15) & 0x1)); }
    inline XMVECTOR XMLoadUNibble4(const XMUNIBBLE4* p) noexcept { return shim::mk(float(p->v & 0xF), float((p->v >> 4) & 0xF), float((p->v >> 8) & 0xF), float((p->v >> 12) & 0xF)); }
    inline void XMStoreU565(XMU565* p, FXMVECTOR V) noexcept
    {
        const uint32_t x = uint32_t(pk::st_int(V.vector4_f32[0], 0.f, 31.f)), y = uint32_t(pk::st_int(V.vector4_f32[1], 0.f, 63.f)), z = uint32_t(pk::st_int(V.vector4_f32[2], 0.f, 31.f));
        p->v = uint16_t(((z & 0x1F) << 11) | ((y & 0x3F) << 5) | (x & 0x1F));
    }
    inline void XMStoreU555(XMU555* p, FXMVECTOR V) noexcept
    {
        const uint32_t x = uint32_t(pk::st_int(V.vector4_f32[0], 0.f, 31.f)), y = uint32_t(pk::st_int(V.vector4_f32[1], 0.f, 31.f)), z = uint32_t(pk::st_int(V.vector4_f32[2], 0.f, 31.f)), w = uint32_t(pk::st_int(V.vector4_f32[3], 0.f, 1.f));
        p->v = uint16_t((w ? 0x8000u : 0u) | ((z & 0x1F) << 10) | ((y & 0x1F) << 5) | (x & 0x1F));
    }
    inline void XMStoreUNibble4(XMUNIBBLE4* p, FXMVECTOR V) noexcept
    {
        const uint32_t x = uint32_t(pk::st_int(V.vector4_f32[0], 0.f, 15.f)), y = uint32_t(pk::st_int(V.vector4_f32[1], 0.f, 15.f)), z = uint32_t(pk::st_int(V.vector4_f32[2], 0.f, 15.f)), w = uint32_t(pk::st_int(V.vector4_f32[3], 0.f, 15.f));
        p->v = uint16_t(((w & 0xF) << 12) | ((z & 0xF) << 8) | ((y & 0xF) << 4) | (x & 0xF));
    }

    // 10:10:10:2
    inline XMVECTOR XMLoadUDecN4(const XMUDECN4* p) noexcept { return shim::mk(float(p->v & 0x3FF) * (1.0f / 1023.0f), float((p->v >> 10) & 0x3FF) * (1.0f / 1023.0f), float((p->v >> 20) & 0x3FF) * (1.0f / 1023.0f), float(p->v >> 30) * (1.0f / 3.0f)); }
    inline XMVECTOR XMLoadUDec4(const XMUDEC4* p) noexcept { return shim::mk(float(p->v & 0x3FF), float((p->v >> 10) & 0x3FF), float((p->v >> 20) & 0x3FF), float(p->v >> 30)); }
    inline XMVECTOR XMLoadUDecN4_XR(const XMUDECN4* p) noexcept
    {
        const int32_t x = int32_t(p->v & 0x3FF) - 0x180, y = int32_t((p->v >> 10) & 0x3FF) - 0x180, z = int32_t((p->v >> 20) & 0x3FF) - 0x180;
        return shim::mk(float(x) / 510.0f, float(y) / 510.0f, float(z) / 510.0f, float(p->v >> 30) / 3.0f);
    }
    inline XMVECTOR XMLoadXDecN4(const XMXDECN4* p) noexcept
    {
        auto sx = [](uint32_t e) noexcept -> float { int32_t v = int32_t(e & 0x3FF); if (v & 0x200) v -= 0x400; return (v == -512) ? -1.f : float(v) * (1.0f / 511.0f); };
        return shim::mk(sx(p->v), sx(p->v >> 10), sx(p->v >> 20), float(p->v >> 30) * (1.0f / 3.0f));
    }
    inline void XMStoreUDecN4(XMUDECN4* p, FXMVECTOR V) noexcept
    {
        const uint32_t x = uint32_t(int32_t(pk::clampf(V.vector4_f32[0], 0.f, 1.f) * 1023.0f)), y = uint32_t(int32_t(pk::clampf(V.vector4_f32[1], 0.f, 1.f) * 1023.0f));
        const uint32_t z = uint32_t(int32_t(pk::clampf(V.vector4_f32[2], 0.f, 1.f) * 1023.0f)), w = uint32_t(int32_t(pk::clampf(V.vector4_f32[3], 0.f, 1.f) * 3.0f));
        p->v = (w << 30) | ((z & 0x3FF) << 20) | ((y & 0x3FF) << 10) | (x & 0x3FF);
    }
    inline void XMStoreUDec4(XMUDEC4* p, FXMVECTOR V) noexcept
    {
        const uint32_t x = uint32_t(pk::st_int(V.vector4_f32[0], 0.f, 1023.f)), y = uint32_t(pk::st_int(V.vector4_f32[1], 0.f, 1023.f)), z = uint32_t(pk::st_int(V.vector4_f32[2], 0.f, 1023.f)), w = uint32_t(pk::st_int(V.vector4_f32[3], 0.f, 3.f));
        p->v = (w << 30) | ((z & 0x3FF) << 20) | ((y & 0x3FF) << 10) | (x & 0x3FF);
    }
    inline void XMStoreUDecN4_XR(XMUDECN4* p, FXMVECTOR V) noexcept
    {
        auto c = [](float v) noexcept -> uint32_t { const float t = v * 510.0f + 384.0f; return uint32_t(shim::rne(pk::clampf(t, 0.f, 1023.f))); };
        const uint32_t w = uint32_t(shim::rne(pk::clampf(V.vector4_f32[3], 0.f, 1.f) * 3.0f));
        p->v = (w << 30) | ((c(V.vector4_f32[2]) & 0x3FF) << 20) | ((c(V.vector4_f32[1]) & 0x3FF) << 10) | (c(V.vector4_f32[0]) & 0x3FF);
    }
    inline void XMStoreXDecN4(XMXDECN4* p, FXMVECTOR V) noexcept
    {
        auto c = [](float v) noexcept -> uint32_t { return uint32_t(int32_t(pk::clampf(v, -1.f, 1.f) * 511.0f)) & 0x3FF; };
        const uint32_t w = uint32_t(int32_t(pk::clampf(V.vector4_f32[3], 0.f, 1.f) * 3.0f));
        p->v = (w << 30) | (c(V.vector4_f32[2]) << 20) | (c(V.vector4_f32[1]) << 10) | c(V.vector4_f32[0]);
    }

    // R11G11B10_FLOAT
    inline XMVECTOR XMLoadFloat3PK(const XMFLOAT3PK* p) noexcept
    {
        auto dec = [](uint32_t m, uint32_t e, uint32_t mbits) noexcept -> float
        {
            uint32_t bits;
            if (e == 0x1f) bits = 0x7f800000u | (m << (23 - mbits));
            else
            {
                int32_t ex;
                if (e != 0) ex = int32_t(e);
                else if (m != 0)
                {
                    ex = 1;
                    do { ex--; m <<= 1; } while ((m & (1u << mbits)) == 0);
                    m &= (1u << mbits) - 1;
                }
                else ex = -112;
                bits = (uint32_t(ex + 112) << 23) | (m << (23 - mbits));
            }
            float f; memcpy(&f, &bits, 4); return f;
        };
        return shim::mk(dec(p->xm, p->xe, 6), dec(p->ym, p->ye, 6), dec(p->zm, p->ze, 5), 0.f);
    }
    inline void XMStoreFloat3PK(XMFLOAT3PK* p, FXMVECTOR V) noexcept
    {
        auto enc = [](float f, uint32_t mbits) noexcept -> uint32_t
        {
            uint32_t I; memcpy(&I, &f, 4);
            const uint32_t sign = I & 0x80000000u;
            I &= 0x7FFFFFFFu;
            const uint32_t maxv = (mbits == 6) ? 0x7C0u : 0x3E0u;   // Inf pattern
            const uint32_t shift = 23 - mbits;
            if ((I & 0x7F800000u) == 0x7F800000u)
            {
                uint32_t r = maxv;
                if (I & 0x7FFFFFu) r = maxv | (((I >> shift) | (I >> (shift - 6)) | (I >> (shift - 12)) | I) & ((1u << mbits) - 1));
                else if (sign) r = 0;
                return r;
            }
            if (sign) return 0;
            if (I > 0x477E0000u && mbits == 6) return 0x7BFu;
            if (I > 0x477C0000u && mbits == 5) return 0x3DFu;
            if (I < 0x38800000u)
            {
                const uint32_t sh = 113u - (I >> 23);
                I = (sh < 32) ? ((0x800000u | (I & 0x7FFFFFu)) >> sh) : 0u;
            }
            else I += 0xC8000000u;
            const uint32_t half = (1u << (shift - 1)) - 1;
            return ((I + half + ((I >> shift) & 1u)) >> shift) & ((mbits == 6) ? 0x7FFu : 0x3FFu);
        };
        const uint32_t x = enc(V.vector4_f32[0], 6), y = enc(V.vector4_f32[1], 6), z = enc(V.vector4_f32[2], 5);
        p->v = (x & 0x7FF) | ((y & 0x7FF) << 11) | ((z & 0x3FF) << 22);
    }

    // R9G9B9E5_SHAREDEXP
    inline XMVECTOR XMLoadFloat3SE(const XMFLOAT3SE* p) noexcept
    {
        uint32_t bits = 0x33800000u + (uint32_t(p->e) << 23);
        float scale; memcpy(&scale, &bits, 4);
        return shim::mk(scale * float(p->xm), scale * float(p->ym), scale * float(p->zm), 1.0f);
    }
    inline void XMStoreFloat3SE(XMFLOAT3SE* p, FXMVECTOR V) noexcept
    {
        const float maxf9 = float(0x1FF << 7);
        const float minf9 = 1.f / float(1 << 16);
        const float x = (V.vector4_f32[0] >= 0.f) ? ((V.vector4_f32[0] > maxf9) ? maxf9 : V.vector4_f32[0]) : 0.f;
        const float y = (V.vector4_f32[1] >= 0.f) ? ((V.vector4_f32[1] > maxf9) ? maxf9 : V.vector4_f32[1]) : 0.f;
        const float z = (V.vector4_f32[2] >= 0.f) ? ((V.vector4_f32[2] > maxf9) ? maxf9 : V.vector4_f32[2]) : 0.f;
        const float max_xy = (x > y) ? x : y;
        const float max_xyz = (max_xy > z) ? max_xy : z;
        const float maxColor = (max_xyz > minf9) ? max_xyz : minf9;
        uint32_t mi; memcpy(&mi, &maxColor, 4);
        mi += 0x00004000u;
        const uint32_t exp = mi >> 23;
        p->e = exp - 0x6f;
        uint32_t si = 0x83000000u - (exp << 23);
        float scaleR; memcpy(&scaleR, &si, 4);
        p->xm = uint32_t(lroundf(x * scaleR));
        p->ym = uint32_t(lroundf(y * scaleR));
        p->zm = uint32_t(lroundf(z * scaleR));
    }
} // namespace PackedVector
} // namespace DirectX
