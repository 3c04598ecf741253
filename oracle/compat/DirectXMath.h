// oracle/compat/DirectXMath.h — TEST INFRASTRUCTURE (oracle build only; never shipped,
// never linked into the product library).
//
// Stand-in for the DirectXMath package that the reference requires
// (/root/reference/CMakeLists.txt:383-389, build/vcpkg.json:4, version unpinned) and
// that is absent from this container.  It implements ONLY the surface the hot-path
// sources touch with scalar, strictly-IEEE, unfused fp32
// semantics that restate the x64 SSE2 code path of DirectXMath (the path GCC builds
// of the reference take: -msse2, no SSE3/SSE4/AVX/F16C/FMA,
// build/CompilerAndLinker.cmake:110-121).  Behaviours marked (M) are restated from
// memory of DirectXMath and cannot be verified here: "parity unpinned" at this
// boundary.  The GPU kernels restate exactly these choices.
#pragma once

#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <cassert>
#include <algorithm>

#define DIRECTX_MATH_VERSION 320

#define XM_CALLCONV
#define XM_ALIGNED_DATA(x) alignas(x)
#define XM_ALIGNED_STRUCT(x) struct alignas(x)
#define XMGLOBALCONST extern const __attribute__((weak))
#define XM_CONSTEXPR constexpr
#define XM_DEPRECATED

namespace DirectX
{
    constexpr uint32_t XM_SELECT_0 = 0x00000000;
    constexpr uint32_t XM_SELECT_1 = 0xFFFFFFFF;

    constexpr uint32_t XM_PERMUTE_0X = 0, XM_PERMUTE_0Y = 1, XM_PERMUTE_0Z = 2, XM_PERMUTE_0W = 3;
    constexpr uint32_t XM_PERMUTE_1X = 4, XM_PERMUTE_1Y = 5, XM_PERMUTE_1Z = 6, XM_PERMUTE_1W = 7;
    constexpr uint32_t XM_SWIZZLE_X = 0, XM_SWIZZLE_Y = 1, XM_SWIZZLE_Z = 2, XM_SWIZZLE_W = 3;

    struct alignas(16) XMVECTOR
    {
        union
        {
            float vector4_f32[4];
            uint32_t vector4_u32[4];
        };
    };

    typedef const XMVECTOR& FXMVECTOR;
    typedef const XMVECTOR& GXMVECTOR;
    typedef const XMVECTOR& HXMVECTOR;
    typedef const XMVECTOR& CXMVECTOR;

    struct alignas(16) XMVECTORF32
    {
        union { float f[4]; XMVECTOR v; };
        operator XMVECTOR() const noexcept { return v; }
        operator const float*() const noexcept { return f; }
    };
    struct alignas(16) XMVECTORI32
    {
        union { int32_t i[4]; XMVECTOR v; };
        operator XMVECTOR() const noexcept { return v; }
    };
    struct alignas(16) XMVECTORU32
    {
        union { uint32_t u[4]; XMVECTOR v; };
        operator XMVECTOR() const noexcept { return v; }
    };

    struct XMFLOAT2 { float x, y; XMFLOAT2() = default; constexpr XMFLOAT2(float _x, float _y) noexcept : x(_x), y(_y) {} };
    struct XMFLOAT3 { float x, y, z; XMFLOAT3() = default; constexpr XMFLOAT3(float _x, float _y, float _z) noexcept : x(_x), y(_y), z(_z) {} };
    struct XMFLOAT4 { float x, y, z, w; XMFLOAT4() = default; constexpr XMFLOAT4(float _x, float _y, float _z, float _w) noexcept : x(_x), y(_y), z(_z), w(_w) {} };
    struct alignas(16) XMFLOAT3A : public XMFLOAT3 { using XMFLOAT3::XMFLOAT3; };
    struct alignas(16) XMFLOAT4A : public XMFLOAT4 { using XMFLOAT4::XMFLOAT4; };
    struct XMINT2 { int32_t x, y; };
    struct XMINT3 { int32_t x, y, z; };
    struct XMINT4 { int32_t x, y, z, w; };
    struct XMUINT2 { uint32_t x, y; };
    struct XMUINT3 { uint32_t x, y, z; };
    struct XMUINT4 { uint32_t x, y, z, w; };

    //---------------------------------------------------------------------------------
    // helpers
    namespace shim
    {
        inline XMVECTOR mk(float x, float y, float z, float w) noexcept
        {
            XMVECTOR r; r.vector4_f32[0] = x; r.vector4_f32[1] = y; r.vector4_f32[2] = z; r.vector4_f32[3] = w; return r;
        }
        inline XMVECTOR mku(uint32_t x, uint32_t y, uint32_t z, uint32_t w) noexcept
        {
            XMVECTOR r; r.vector4_u32[0] = x; r.vector4_u32[1] = y; r.vector4_u32[2] = z; r.vector4_u32[3] = w; return r;
        }
        // SSE semantics: _mm_max_ps(a,b) = (a > b) ? a : b ; _mm_min_ps(a,b) = (a < b) ? a : b
        inline float ssemax(float a, float b) noexcept { return (a > b) ? a : b; }
        inline float ssemin(float a, float b) noexcept { return (a < b) ? a : b; }
        // cvtps_epi32 under the default MXCSR (round to nearest even); inputs are pre-clamped
        inline int32_t rne(float f) noexcept { return static_cast<int32_t>(nearbyintf(f)); }
        // cvttps_epi32 with DirectXMath's positive-overflow fix-up (M)
        inline int32_t f2i_trunc_sat(float f) noexcept
        {
            if (f != f) return static_cast<int32_t>(0x80000000u);
            if (f > 2147483520.0f) return 0x7FFFFFFF;
            if (f < -2147483648.0f) return static_cast<int32_t>(0x80000000u);
            return static_cast<int32_t>(f);
        }
        inline uint32_t f2u_trunc_sat(float f) noexcept
        {
            if (!(f > 0.0f)) return 0u;
            if (f >= 4294967296.0f) return 0xFFFFFFFFu;
            return static_cast<uint32_t>(f);
        }
    }

    //---------------------------------------------------------------------------------
    // globals (values per DirectXMath.h "Globals" section)
    XMGLOBALCONST XMVECTORF32 g_XMZero = { { { 0.0f, 0.0f, 0.0f, 0.0f } } };
    XMGLOBALCONST XMVECTORF32 g_XMOne = { { { 1.0f, 1.0f, 1.0f, 1.0f } } };
    XMGLOBALCONST XMVECTORF32 g_XMOneHalf = { { { 0.5f, 0.5f, 0.5f, 0.5f } } };
    XMGLOBALCONST XMVECTORF32 g_XMTwo = { { { 2.0f, 2.0f, 2.0f, 2.0f } } };
    XMGLOBALCONST XMVECTORF32 g_XMNegativeOne = { { { -1.0f, -1.0f, -1.0f, -1.0f } } };
    XMGLOBALCONST XMVECTORF32 g_XMIdentityR3 = { { { 0.0f, 0.0f, 0.0f, 1.0f } } };
    XMGLOBALCONST XMVECTORU32 g_XMSelect1000 = { { { XM_SELECT_1, XM_SELECT_0, XM_SELECT_0, XM_SELECT_0 } } };
    XMGLOBALCONST XMVECTORU32 g_XMSelect1100 = { { { XM_SELECT_1, XM_SELECT_1, XM_SELECT_0, XM_SELECT_0 } } };
    XMGLOBALCONST XMVECTORU32 g_XMSelect1110 = { { { XM_SELECT_1, XM_SELECT_1, XM_SELECT_1, XM_SELECT_0 } } };
    XMGLOBALCONST XMVECTORU32 g_XMMaskX = { { { 0xFFFFFFFF, 0x00000000, 0x00000000, 0x00000000 } } };
    XMGLOBALCONST XMVECTORU32 g_XMMaskY = { { { 0x00000000, 0xFFFFFFFF, 0x00000000, 0x00000000 } } };
    XMGLOBALCONST XMVECTORU32 g_XMMaskZ = { { { 0x00000000, 0x00000000, 0xFFFFFFFF, 0x00000000 } } };
    XMGLOBALCONST XMVECTORU32 g_XMMaskW = { { { 0x00000000, 0x00000000, 0x00000000, 0xFFFFFFFF } } };

    //---------------------------------------------------------------------------------
    // construction / access
    inline XMVECTOR XMVectorZero() noexcept { return shim::mk(0.f, 0.f, 0.f, 0.f); }
    inline XMVECTOR XMVectorSet(float x, float y, float z, float w) noexcept { return shim::mk(x, y, z, w); }
    inline XMVECTOR XMVectorReplicate(float v) noexcept { return shim::mk(v, v, v, v); }
    inline XMVECTOR XMVectorSplatX(FXMVECTOR V) noexcept { return XMVectorReplicate(V.vector4_f32[0]); }
    inline XMVECTOR XMVectorSplatY(FXMVECTOR V) noexcept { return XMVectorReplicate(V.vector4_f32[1]); }
    inline XMVECTOR XMVectorSplatZ(FXMVECTOR V) noexcept { return XMVectorReplicate(V.vector4_f32[2]); }
    inline XMVECTOR XMVectorSplatW(FXMVECTOR V) noexcept { return XMVectorReplicate(V.vector4_f32[3]); }
    inline float XMVectorGetX(FXMVECTOR V) noexcept { return V.vector4_f32[0]; }
    inline float XMVectorGetY(FXMVECTOR V) noexcept { return V.vector4_f32[1]; }
    inline float XMVectorGetZ(FXMVECTOR V) noexcept { return V.vector4_f32[2]; }
    inline float XMVectorGetW(FXMVECTOR V) noexcept { return V.vector4_f32[3]; }
    inline XMVECTOR XMVectorSetX(FXMVECTOR V, float x) noexcept { XMVECTOR r = V; r.vector4_f32[0] = x; return r; }
    inline XMVECTOR XMVectorSetY(FXMVECTOR V, float y) noexcept { XMVECTOR r = V; r.vector4_f32[1] = y; return r; }
    inline XMVECTOR XMVectorSetZ(FXMVECTOR V, float z) noexcept { XMVECTOR r = V; r.vector4_f32[2] = z; return r; }
    inline XMVECTOR XMVectorSetW(FXMVECTOR V, float w) noexcept { XMVECTOR r = V; r.vector4_f32[3] = w; return r; }

    template<uint32_t E0, uint32_t E1, uint32_t E2, uint32_t E3>
    inline XMVECTOR XMVectorSwizzle(FXMVECTOR V) noexcept
    {
        static_assert(E0 < 4 && E1 < 4 && E2 < 4 && E3 < 4, "swizzle index");
        return shim::mku(V.vector4_u32[E0], V.vector4_u32[E1], V.vector4_u32[E2], V.vector4_u32[E3]);
    }
    template<uint32_t P0, uint32_t P1, uint32_t P2, uint32_t P3>
    inline XMVECTOR XMVectorPermute(FXMVECTOR V1, FXMVECTOR V2) noexcept
    {
        static_assert(P0 < 8 && P1 < 8 && P2 < 8 && P3 < 8, "permute index");
        const uint32_t* a[2] = { V1.vector4_u32, V2.vector4_u32 };
        return shim::mku(a[P0 >> 2][P0 & 3], a[P1 >> 2][P1 & 3], a[P2 >> 2][P2 & 3], a[P3 >> 2][P3 & 3]);
    }
    // (V1 & ~Control) | (V2 & Control)
    inline XMVECTOR XMVectorSelect(FXMVECTOR V1, FXMVECTOR V2, FXMVECTOR C) noexcept
    {
        XMVECTOR r;
        for (int i = 0; i < 4; ++i)
            r.vector4_u32[i] = (V1.vector4_u32[i] & ~C.vector4_u32[i]) | (V2.vector4_u32[i] & C.vector4_u32[i]);
        return r;
    }
    inline XMVECTOR XMVectorMergeXY(FXMVECTOR V1, FXMVECTOR V2) noexcept
    {
        return shim::mku(V1.vector4_u32[0], V2.vector4_u32[0], V1.vector4_u32[1], V2.vector4_u32[1]);
    }

    //---------------------------------------------------------------------------------
    // arithmetic: four independent scalar fp32 ops, never fused
#define DXM_SHIM_BINOP(name, expr) \
    inline XMVECTOR name(FXMVECTOR A, FXMVECTOR B) noexcept { XMVECTOR r; for (int i = 0; i < 4; ++i) { const float a = A.vector4_f32[i]; const float b = B.vector4_f32[i]; r.vector4_f32[i] = (expr); } return r; }
    DXM_SHIM_BINOP(XMVectorAdd, a + b)
    DXM_SHIM_BINOP(XMVectorSubtract, a - b)
    DXM_SHIM_BINOP(XMVectorMultiply, a * b)
    DXM_SHIM_BINOP(XMVectorDivide, a / b)
    DXM_SHIM_BINOP(XMVectorMin, shim::ssemin(a, b))
    DXM_SHIM_BINOP(XMVectorMax, shim::ssemax(a, b))
    DXM_SHIM_BINOP(XMVectorPow, powf(a, b))
#undef DXM_SHIM_BINOP

    // (V1*V2)+V3 with two roundings: SSE path without _XM_FMA3_INTRINSICS_ (M)
    inline XMVECTOR XMVectorMultiplyAdd(FXMVECTOR V1, FXMVECTOR V2, FXMVECTOR V3) noexcept
    {
        XMVECTOR r;
        for (int i = 0; i < 4; ++i)
        {
            const float p = V1.vector4_f32[i] * V2.vector4_f32[i];   // build uses -ffp-contract=off
            r.vector4_f32[i] = p + V3.vector4_f32[i];
        }
        return r;
    }
    inline XMVECTOR XMVectorNegate(FXMVECTOR V) noexcept
    {
        return shim::mk(0.f - V.vector4_f32[0], 0.f - V.vector4_f32[1], 0.f - V.vector4_f32[2], 0.f - V.vector4_f32[3]);   // SSE: _mm_sub_ps(zero, V)
    }
    inline XMVECTOR XMVectorScale(FXMVECTOR V, float s) noexcept
    {
        return shim::mk(V.vector4_f32[0] * s, V.vector4_f32[1] * s, V.vector4_f32[2] * s, V.vector4_f32[3] * s);
    }
    // SSE: min(max(Min, V), Max)
    inline XMVECTOR XMVectorClamp(FXMVECTOR V, FXMVECTOR Min, FXMVECTOR Max) noexcept
    {
        XMVECTOR r;
        for (int i = 0; i < 4; ++i)
            r.vector4_f32[i] = shim::ssemin(shim::ssemax(Min.vector4_f32[i], V.vector4_f32[i]), Max.vector4_f32[i]);
        return r;
    }
    // SSE: min(max(V, 0), 1)
    inline XMVECTOR XMVectorSaturate(FXMVECTOR V) noexcept
    {
        XMVECTOR r;
        for (int i = 0; i < 4; ++i)
            r.vector4_f32[i] = shim::ssemin(shim::ssemax(V.vector4_f32[i], 0.0f), 1.0f);
        return r;
    }
    // round half to even (magic-number add in the SSE2 path) (M)
    inline XMVECTOR XMVectorRound(FXMVECTOR V) noexcept
    {
        return shim::mk(nearbyintf(V.vector4_f32[0]), nearbyintf(V.vector4_f32[1]), nearbyintf(V.vector4_f32[2]), nearbyintf(V.vector4_f32[3]));
    }
    inline XMVECTOR XMVectorTruncate(FXMVECTOR V) noexcept
    {
        return shim::mk(truncf(V.vector4_f32[0]), truncf(V.vector4_f32[1]), truncf(V.vector4_f32[2]), truncf(V.vector4_f32[3]));
    }
    // V0 + (V1 - V0) * t, unfused
    inline XMVECTOR XMVectorLerp(FXMVECTOR V0, FXMVECTOR V1, float t) noexcept
    {
        return XMVectorMultiplyAdd(XMVectorSubtract(V1, V0), XMVectorReplicate(t), V0);
    }
    // SSE2 (no SSE3): (x+y)+(z+w) replicated
    inline XMVECTOR XMVectorSum(FXMVECTOR V) noexcept
    {
        const float a = V.vector4_f32[0] + V.vector4_f32[1];
        const float b = V.vector4_f32[2] + V.vector4_f32[3];
        return XMVectorReplicate(a + b);
    }
    // SSE2: (x*x' + y*y') + z*z' replicated
    inline XMVECTOR XMVector3Dot(FXMVECTOR V1, FXMVECTOR V2) noexcept
    {
        const float x = V1.vector4_f32[0] * V2.vector4_f32[0];
        const float y = V1.vector4_f32[1] * V2.vector4_f32[1];
        const float z = V1.vector4_f32[2] * V2.vector4_f32[2];
        const float xy = x + y;
        return XMVectorReplicate(xy + z);
    }
    // SSE2: (x+z)+(y+w) replicated (M)
    inline XMVECTOR XMVector4Dot(FXMVECTOR V1, FXMVECTOR V2) noexcept
    {
        const float x = V1.vector4_f32[0] * V2.vector4_f32[0];
        const float y = V1.vector4_f32[1] * V2.vector4_f32[1];
        const float z = V1.vector4_f32[2] * V2.vector4_f32[2];
        const float w = V1.vector4_f32[3] * V2.vector4_f32[3];
        const float xz = x + z;
        const float yw = y + w;
        return XMVectorReplicate(xz + yw);
    }
    inline bool XMVector4Less(FXMVECTOR V1, FXMVECTOR V2) noexcept
    {
        return V1.vector4_f32[0] < V2.vector4_f32[0] && V1.vector4_f32[1] < V2.vector4_f32[1]
            && V1.vector4_f32[2] < V2.vector4_f32[2] && V1.vector4_f32[3] < V2.vector4_f32[3];
    }
    inline XMVECTOR XMVectorGreater(FXMVECTOR A, FXMVECTOR B) noexcept
    {
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_u32[i] = (A.vector4_f32[i] > B.vector4_f32[i]) ? 0xFFFFFFFFu : 0u; return r;
    }
    inline XMVECTOR XMVectorLess(FXMVECTOR A, FXMVECTOR B) noexcept
    {
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_u32[i] = (A.vector4_f32[i] < B.vector4_f32[i]) ? 0xFFFFFFFFu : 0u; return r;
    }

    //---------------------------------------------------------------------------------
    // int <-> float vector conversion
    inline XMVECTOR XMConvertVectorIntToFloat(FXMVECTOR V, uint32_t DivExponent) noexcept
    {
        const float s = 1.0f / static_cast<float>(1u << DivExponent);
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_f32[i] = static_cast<float>(static_cast<int32_t>(V.vector4_u32[i])) * s; return r;
    }
    inline XMVECTOR XMConvertVectorUIntToFloat(FXMVECTOR V, uint32_t DivExponent) noexcept
    {
        const float s = 1.0f / static_cast<float>(1u << DivExponent);
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_f32[i] = static_cast<float>(V.vector4_u32[i]) * s; return r;
    }
    inline XMVECTOR XMConvertVectorFloatToInt(FXMVECTOR V, uint32_t MulExponent) noexcept
    {
        const float s = static_cast<float>(1u << MulExponent);
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_u32[i] = static_cast<uint32_t>(shim::f2i_trunc_sat(V.vector4_f32[i] * s)); return r;
    }
    inline XMVECTOR XMConvertVectorFloatToUInt(FXMVECTOR V, uint32_t MulExponent) noexcept
    {
        const float s = static_cast<float>(1u << MulExponent);
        XMVECTOR r; for (int i = 0; i < 4; ++i) r.vector4_u32[i] = shim::f2u_trunc_sat(V.vector4_f32[i] * s); return r;
    }

    //---------------------------------------------------------------------------------
    // loads / stores
    inline XMVECTOR XMLoadInt(const uint32_t* p) noexcept { return shim::mku(*p, 0, 0, 0); }
    inline XMVECTOR XMLoadFloat(const float* p) noexcept { return shim::mk(*p, 0.f, 0.f, 0.f); }
    inline XMVECTOR XMLoadFloat2(const XMFLOAT2* p) noexcept { return shim::mk(p->x, p->y, 0.f, 0.f); }
    inline XMVECTOR XMLoadFloat3(const XMFLOAT3* p) noexcept { return shim::mk(p->x, p->y, p->z, 0.f); }
    inline XMVECTOR XMLoadFloat4(const XMFLOAT4* p) noexcept { return shim::mk(p->x, p->y, p->z, p->w); }
    inline XMVECTOR XMLoadFloat4A(const XMFLOAT4A* p) noexcept { return shim::mk(p->x, p->y, p->z, p->w); }
    inline XMVECTOR XMLoadSInt2(const XMINT2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadSInt3(const XMINT3* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), 0.f); }
    inline XMVECTOR XMLoadSInt4(const XMINT4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }
    inline XMVECTOR XMLoadUInt2(const XMUINT2* p) noexcept { return shim::mk(float(p->x), float(p->y), 0.f, 0.f); }
    inline XMVECTOR XMLoadUInt3(const XMUINT3* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), 0.f); }
    inline XMVECTOR XMLoadUInt4(const XMUINT4* p) noexcept { return shim::mk(float(p->x), float(p->y), float(p->z), float(p->w)); }

    inline void XMStoreInt(uint32_t* p, FXMVECTOR V) noexcept { *p = V.vector4_u32[0]; }
    inline void XMStoreFloat(float* p, FXMVECTOR V) noexcept { *p = V.vector4_f32[0]; }
    inline void XMStoreFloat2(XMFLOAT2* p, FXMVECTOR V) noexcept { p->x = V.vector4_f32[0]; p->y = V.vector4_f32[1]; }
    inline void XMStoreFloat3(XMFLOAT3* p, FXMVECTOR V) noexcept { p->x = V.vector4_f32[0]; p->y = V.vector4_f32[1]; p->z = V.vector4_f32[2]; }
    inline void XMStoreFloat3A(XMFLOAT3A* p, FXMVECTOR V) noexcept { p->x = V.vector4_f32[0]; p->y = V.vector4_f32[1]; p->z = V.vector4_f32[2]; }
    inline void XMStoreFloat4(XMFLOAT4* p, FXMVECTOR V) noexcept { p->x = V.vector4_f32[0]; p->y = V.vector4_f32[1]; p->z = V.vector4_f32[2]; p->w = V.vector4_f32[3]; }
    inline void XMStoreFloat4A(XMFLOAT4A* p, FXMVECTOR V) noexcept { p->x = V.vector4_f32[0]; p->y = V.vector4_f32[1]; p->z = V.vector4_f32[2]; p->w = V.vector4_f32[3]; }
    inline void XMStoreSInt2(XMINT2* p, FXMVECTOR V) noexcept { p->x = shim::f2i_trunc_sat(V.vector4_f32[0]); p->y = shim::f2i_trunc_sat(V.vector4_f32[1]); }
    inline void XMStoreSInt3(XMINT3* p, FXMVECTOR V) noexcept { p->x = shim::f2i_trunc_sat(V.vector4_f32[0]); p->y = shim::f2i_trunc_sat(V.vector4_f32[1]); p->z = shim::f2i_trunc_sat(V.vector4_f32[2]); }
    inline void XMStoreSInt4(XMINT4* p, FXMVECTOR V) noexcept { p->x = shim::f2i_trunc_sat(V.vector4_f32[0]); p->y = shim::f2i_trunc_sat(V.vector4_f32[1]); p->z = shim::f2i_trunc_sat(V.vector4_f32[2]); p->w = shim::f2i_trunc_sat(V.vector4_f32[3]); }
    inline void XMStoreUInt2(XMUINT2* p, FXMVECTOR V) noexcept { p->x = shim::f2u_trunc_sat(V.vector4_f32[0]); p->y = shim::f2u_trunc_sat(V.vector4_f32[1]); }
    inline void XMStoreUInt3(XMUINT3* p, FXMVECTOR V) noexcept { p->x = shim::f2u_trunc_sat(V.vector4_f32[0]); p->y = shim::f2u_trunc_sat(V.vector4_f32[1]); p->z = shim::f2u_trunc_sat(V.vector4_f32[2]); }
    inline void XMStoreUInt4(XMUINT4* p, FXMVECTOR V) noexcept { p->x = shim::f2u_trunc_sat(V.vector4_f32[0]); p->y = shim::f2u_trunc_sat(V.vector4_f32[1]); p->z = shim::f2u_trunc_sat(V.vector4_f32[2]); p->w = shim::f2u_trunc_sat(V.vector4_f32[3]); }

    //---------------------------------------------------------------------------------
    // sRGB <-> linear (DirectXMathMisc.inl XMColorSRGBToRGB / XMColorRGBToSRGB) (M)
    inline XMVECTOR XMColorSRGBToRGB(FXMVECTOR srgb) noexcept
    {
        XMVECTOR r = srgb;
        for (int i = 0; i < 3; ++i)
        {
            const float v = shim::ssemin(shim::ssemax(srgb.vector4_f32[i], 0.0f), 1.0f);
            const float v0 = v * (1.0f / 12.92f);
            const float t = (v + 0.055f) * (1.0f / 1.055f);
            const float v1 = powf(t, 2.4f);
            r.vector4_f32[i] = (v > 0.04045f) ? v1 : v0;
        }
        return r;
    }
    inline XMVECTOR XMColorRGBToSRGB(FXMVECTOR rgb) noexcept
    {
        XMVECTOR r = rgb;
        for (int i = 0; i < 3; ++i)
        {
            const float v = shim::ssemin(shim::ssemax(rgb.vector4_f32[i], 0.0f), 1.0f);
            const float v0 = v * 12.92f;
            const float p = powf(v, 1.0f / 2.4f);
            const float v1 = 1.055f * p - 0.055f;
            r.vector4_f32[i] = (v < 0.0031308f) ? v0 : v1;
        }
        return r;
    }
} // namespace DirectX
