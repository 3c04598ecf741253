// dxb_bc15.cuh — BC1 / BC2 / BC3 / BC4 / BC5 block encoders, one thread per 4x4 block.
// Restates (same fp32 operation order, no FMA contraction, IEEE division, C-cast truncation):
//   OptimizeRGB            BC.cpp:65-314
//   EncodeBC1              BC.cpp:370-685      D3DXEncodeBC1 BC.cpp:738-795
//   D3DXEncodeBC2          BC.cpp:828-895
//   D3DXEncodeBC3          BC.cpp:944-1141
//   OptimizeAlpha<bRange>  BC.h:187-311
//   FindEndPointsBC4U/S, FindClosestUNORM/SNORM, FloatToSNorm, D3DXEncodeBC4/5   BC4BC5.cpp:158-562
// Bit-exactness contract: outputs are memcmp-identical to the reference built by oracle/Makefile.
// The fp32 sums are accumulated in pixel order 0..15 exactly as the reference loops do.
#pragma once
#include "dxb_portable.h"
#include "dxb_formats.h"
#include "dxb_pixel.cuh"
#include "dxb_block.cuh"

struct dxb_rgb { float r, g, b; };

// g_Luminance / g_LuminanceInv (BC.cpp:30-31): quotients folded in fp32
#define DXB_LUM_R (0.2125f / 0.7154f)
#define DXB_LUM_B (0.0721f / 0.7154f)
#define DXB_LUMINV_R (0.7154f / 0.2125f)
#define DXB_LUMINV_B (0.7154f / 0.0721f)

// Floyd–Steinberg propagation inside the 4x4 block (BC.cpp:451-481 and the identical copies)
DXB_DEV void dxb_diffuse1(float* err, int i, float d)
{
    if (3 != (i & 3)) err[i + 1] += d * (7.0f / 16.0f);
    if (i < 12)
    {
        if (i & 3) err[i + 3] += d * (3.0f / 16.0f);
        err[i + 4] += d * (5.0f / 16.0f);
        if (3 != (i & 3)) err[i + 5] += d * (1.0f / 16.0f);
    }
}

// Encode565 / Decode565 (BC.cpp:36-61)
DXB_DEV uint32_t dxb_encode565(float r, float g, float b)
{
    r = (r < 0.0f) ? 0.0f : (r > 1.0f) ? 1.0f : r;
    g = (g < 0.0f) ? 0.0f : (g > 1.0f) ? 1.0f : g;
    b = (b < 0.0f) ? 0.0f : (b > 1.0f) ? 1.0f : b;
    const float fr = r * 31.0f, fg = g * 63.0f, fb = b * 31.0f;
    return (uint32_t)(((dxb_f2i(fr + 0.5f) << 11) | (dxb_f2i(fg + 0.5f) << 5) | (dxb_f2i(fb + 0.5f) << 0)) & 0xFFFF);
}
DXB_DEV dxb_rgb dxb_decode565(uint32_t w)
{
    dxb_rgb c;
    c.r = (float)((w >> 11) & 31) * (1.0f / 31.0f);
    c.g = (float)((w >> 5) & 63) * (1.0f / 63.0f);
    c.b = (float)((w >> 0) & 31) * (1.0f / 31.0f);
    return c;
}

// OptimizeRGB (BC.cpp:65-314).  pts = 16 colours (already weighted by g_Luminance unless UNIFORM).
DXB_DEV void dxb_optimize_rgb(dxb_rgb* pX, dxb_rgb* pY, const dxb_rgb* pts, uint32_t cSteps, uint32_t flags)
{
    const float fEpsilon = (0.25f / 64.0f) * (0.25f / 64.0f);
    // pC3/pD3/pC4/pD4 (BC.cpp:73-76)
    float pC[4], pD[4];
    if (3 == cSteps)
    {
        pC[0] = 2.0f / 2.0f; pC[1] = 1.0f / 2.0f; pC[2] = 0.0f / 2.0f; pC[3] = 0.0f;
        pD[0] = 0.0f / 2.0f; pD[1] = 1.0f / 2.0f; pD[2] = 2.0f / 2.0f; pD[3] = 0.0f;
    }
    else
    {
        pC[0] = 3.0f / 3.0f; pC[1] = 2.0f / 3.0f; pC[2] = 1.0f / 3.0f; pC[3] = 0.0f / 3.0f;
        pD[0] = 0.0f / 3.0f; pD[1] = 1.0f / 3.0f; pD[2] = 2.0f / 3.0f; pD[3] = 3.0f / 3.0f;
    }

    dxb_rgb X, Y;
    if (flags & DXB_BC_FLAGS_UNIFORM) { X.r = 1.0f; X.g = 1.0f; X.b = 1.0f; }
    else { X.r = DXB_LUM_R; X.g = 1.0f; X.b = DXB_LUM_B; }
    Y.r = 0.0f; Y.g = 0.0f; Y.b = 0.0f;

    for (int i = 0; i < 16; ++i)
    {
        if (pts[i].r < X.r) X.r = pts[i].r;
        if (pts[i].g < X.g) X.g = pts[i].g;
        if (pts[i].b < X.b) X.b = pts[i].b;
        if (pts[i].r > Y.r) Y.r = pts[i].r;
        if (pts[i].g > Y.g) Y.g = pts[i].g;
        if (pts[i].b > Y.b) Y.b = pts[i].b;
    }

    const float ABr = Y.r - X.r, ABg = Y.g - X.g, ABb = Y.b - X.b;
    float fAB;
    { const float a = ABr * ABr, b = ABg * ABg, c = ABb * ABb; const float ab = a + b; fAB = ab + c; }

    if (fAB < 1.175494351e-38f)   // FLT_MIN
    {
        *pX = X; *pY = Y;
        return;
    }

    const float fABInv = 1.0f / fAB;
    dxb_rgb Dir; Dir.r = ABr * fABInv; Dir.g = ABg * fABInv; Dir.b = ABb * fABInv;
    dxb_rgb Mid; Mid.r = (X.r + Y.r) * 0.5f; Mid.g = (X.g + Y.g) * 0.5f; Mid.b = (X.b + Y.b) * 0.5f;

    float fDir0 = 0.0f, fDir1 = 0.0f, fDir2 = 0.0f, fDir3 = 0.0f;
    for (int i = 0; i < 16; ++i)
    {
        const float Ptr = (pts[i].r - Mid.r) * Dir.r;
        const float Ptg = (pts[i].g - Mid.g) * Dir.g;
        const float Ptb = (pts[i].b - Mid.b) * Dir.b;
        float f;
        { const float s = Ptr + Ptg; f = s + Ptb; } fDir0 += f * f;
        { const float s = Ptr + Ptg; f = s - Ptb; } fDir1 += f * f;
        { const float s = Ptr - Ptg; f = s + Ptb; } fDir2 += f * f;
        { const float s = Ptr - Ptg; f = s - Ptb; } fDir3 += f * f;
    }

    float fDirMax = fDir0; int iDirMax = 0;
    if (fDir1 > fDirMax) { fDirMax = fDir1; iDirMax = 1; }
    if (fDir2 > fDirMax) { fDirMax = fDir2; iDirMax = 2; }
    if (fDir3 > fDirMax) { fDirMax = fDir3; iDirMax = 3; }

    if (iDirMax & 2) { const float f = X.g; X.g = Y.g; Y.g = f; }
    if (iDirMax & 1) { const float f = X.b; X.b = Y.b; Y.b = f; }

    if (fAB < 1.0f / 4096.0f)
    {
        *pX = X; *pY = Y;
        return;
    }

    const float fSteps = (float)(cSteps - 1);

#if DXB_ON_DEVICE
    #pragma unroll 1      // keeps the kernel small: straight-line code this long is instruction-fetch bound
#endif
    for (int iter = 0; iter < 8; ++iter)
    {
        dxb_rgb pSteps[4];
        for (uint32_t s = 0; s < cSteps; ++s)
        {
            { const float a = X.r * pC[s], b = Y.r * pD[s]; pSteps[s].r = a + b; }
            { const float a = X.g * pC[s], b = Y.g * pD[s]; pSteps[s].g = a + b; }
            { const float a = X.b * pC[s], b = Y.b * pD[s]; pSteps[s].b = a + b; }
        }

        Dir.r = Y.r - X.r; Dir.g = Y.g - X.g; Dir.b = Y.b - X.b;
        float fLen;
        { const float a = Dir.r * Dir.r, b = Dir.g * Dir.g, c = Dir.b * Dir.b; const float ab = a + b; fLen = ab + c; }
        if (fLen < (1.0f / 4096.0f)) break;

        const float fScale = fSteps / fLen;
        Dir.r *= fScale; Dir.g *= fScale; Dir.b *= fScale;

        float d2X = 0.0f, d2Y = 0.0f;
        dxb_rgb dX, dY; dX.r = dX.g = dX.b = 0.0f; dY.r = dY.g = dY.b = 0.0f;

        for (int i = 0; i < 16; ++i)
        {
            float fDot;
            {
                const float a = (pts[i].r - X.r) * Dir.r, b = (pts[i].g - X.g) * Dir.g, c = (pts[i].b - X.b) * Dir.b;
                const float ab = a + b; fDot = ab + c;
            }
            uint32_t iStep;
            if (fDot <= 0.0f) iStep = 0;
            else if (fDot >= fSteps) iStep = cSteps - 1;
            else iStep = dxb_f2u(fDot + 0.5f);

            const float Dr = pSteps[iStep].r - pts[i].r;
            const float Dg = pSteps[iStep].g - pts[i].g;
            const float Db = pSteps[iStep].b - pts[i].b;

            const float fC = pC[iStep] * (1.0f / 8.0f);
            const float fD = pD[iStep] * (1.0f / 8.0f);

            d2X += fC * pC[iStep];
            dX.r += fC * Dr; dX.g += fC * Dg; dX.b += fC * Db;

            d2Y += fD * pD[iStep];
            dY.r += fD * Dr; dY.g += fD * Dg; dY.b += fD * Db;
        }

        if (d2X > 0.0f)
        {
            const float f = -1.0f / d2X;
            X.r += dX.r * f; X.g += dX.g * f; X.b += dX.b * f;
        }
        if (d2Y > 0.0f)
        {
            const float f = -1.0f / d2Y;
            Y.r += dY.r * f; Y.g += dY.g * f; Y.b += dY.b * f;
        }

        if ((dX.r * dX.r < fEpsilon) && (dX.g * dX.g < fEpsilon) && (dX.b * dX.b < fEpsilon) &&
            (dY.r * dY.r < fEpsilon) && (dY.g * dY.g < fEpsilon) && (dY.b * dY.b < fEpsilon))
            break;
    }

    *pX = X; *pY = Y;
}

// EncodeBC1 (BC.cpp:370-685).  px = 16 RGBA pixels as handed to D3DXEncodeBC1/2/3 (alpha possibly
// already dithered by the caller).  Writes rgb[0], rgb[1] (16-bit each) and the 32-bit bitmap.
DXB_DEV void dxb_encode_bc1_core(const dxb_px* px, bool bColorKey, float threshold, uint32_t flags,
                                 uint32_t* outRgb0, uint32_t* outRgb1, uint32_t* outBitmap)
{
    uint32_t uSteps;
    if (bColorKey)
    {
        uint32_t uColorKey = 0;
        for (int i = 0; i < 16; ++i) if (px[i].w < threshold) uColorKey++;
        if (16 == uColorKey)
        {
            *outRgb0 = 0x0000; *outRgb1 = 0xffff; *outBitmap = 0xffffffffu;
            return;
        }
        uSteps = (uColorKey > 0) ? 3u : 4u;
    }
    else uSteps = 4u;

    const bool dither = (flags & DXB_BC_FLAGS_DITHER_RGB) != 0;
    const bool uniform = (flags & DXB_BC_FLAGS_UNIFORM) != 0;

    dxb_rgb Color[16];
    float ErrR[16], ErrG[16], ErrB[16];
    if (dither) for (int i = 0; i < 16; ++i) { ErrR[i] = 0.0f; ErrG[i] = 0.0f; ErrB[i] = 0.0f; }

    for (int i = 0; i < 16; ++i)
    {
        float cr = px[i].x, cg = px[i].y, cb = px[i].z;
        if (dither) { cr += ErrR[i]; cg += ErrG[i]; cb += ErrB[i]; }

        { const float t = cr * 31.0f; Color[i].r = (float)dxb_f2i(t + 0.5f) * (1.0f / 31.0f); }
        { const float t = cg * 63.0f; Color[i].g = (float)dxb_f2i(t + 0.5f) * (1.0f / 63.0f); }
        { const float t = cb * 31.0f; Color[i].b = (float)dxb_f2i(t + 0.5f) * (1.0f / 31.0f); }

        if (dither)
        {
            // Color[i].a == 1.0f (BC.cpp:440)
            const float dr = 1.0f * (cr - Color[i].r), dg = 1.0f * (cg - Color[i].g), db = 1.0f * (cb - Color[i].b);
            dxb_diffuse1(ErrR, i, dr); dxb_diffuse1(ErrG, i, dg); dxb_diffuse1(ErrB, i, db);
        }

        if (!uniform) { Color[i].r *= DXB_LUM_R; Color[i].g *= 1.0f; Color[i].b *= DXB_LUM_B; }
    }

    dxb_rgb ColorA, ColorB, ColorC, ColorD;
    dxb_optimize_rgb(&ColorA, &ColorB, Color, uSteps, flags);

    if (uniform) { ColorC = ColorA; ColorD = ColorB; }
    else
    {
        ColorC.r = ColorA.r * DXB_LUMINV_R; ColorC.g = ColorA.g * 1.0f; ColorC.b = ColorA.b * DXB_LUMINV_B;
        ColorD.r = ColorB.r * DXB_LUMINV_R; ColorD.g = ColorB.g * 1.0f; ColorD.b = ColorB.b * DXB_LUMINV_B;
    }

    const uint32_t wColorA = dxb_encode565(ColorC.r, ColorC.g, ColorC.b);
    const uint32_t wColorB = dxb_encode565(ColorD.r, ColorD.g, ColorD.b);

    if ((uSteps == 4) && (wColorA == wColorB))
    {
        *outRgb0 = wColorA; *outRgb1 = wColorB; *outBitmap = 0;
        return;
    }

    ColorC = dxb_decode565(wColorA);
    ColorD = dxb_decode565(wColorB);

    if (uniform) { ColorA = ColorC; ColorB = ColorD; }
    else
    {
        ColorA.r = ColorC.r * DXB_LUM_R; ColorA.g = ColorC.g * 1.0f; ColorA.b = ColorC.b * DXB_LUM_B;
        ColorB.r = ColorD.r * DXB_LUM_R; ColorB.g = ColorD.g * 1.0f; ColorB.b = ColorD.b * DXB_LUM_B;
    }

    dxb_rgb Step[4];
    if ((3 == uSteps) == (wColorA <= wColorB))
    {
        *outRgb0 = wColorA; *outRgb1 = wColorB;
        Step[0] = ColorA; Step[1] = ColorB;
    }
    else
    {
        *outRgb0 = wColorB; *outRgb1 = wColorA;
        Step[0] = ColorB; Step[1] = ColorA;
    }

    // pSteps3 = {0,2,1}; pSteps4 = {0,2,3,1}   (BC.cpp:566-567); HDRColorALerp BC.h:149-156
    uint32_t pStepsLut;   // 2 bits per entry
    if (3 == uSteps)
    {
        pStepsLut = 0u | (2u << 2) | (1u << 4);
        { const float d = Step[1].r - Step[0].r; const float m = 0.5f * d; Step[2].r = Step[0].r + m; }
        { const float d = Step[1].g - Step[0].g; const float m = 0.5f * d; Step[2].g = Step[0].g + m; }
        { const float d = Step[1].b - Step[0].b; const float m = 0.5f * d; Step[2].b = Step[0].b + m; }
        Step[3].r = 0.0f; Step[3].g = 0.0f; Step[3].b = 0.0f;
    }
    else
    {
        pStepsLut = 0u | (2u << 2) | (3u << 4) | (1u << 6);
        { const float d = Step[1].r - Step[0].r; const float m1 = (1.0f / 3.0f) * d; Step[2].r = Step[0].r + m1; const float m2 = (2.0f / 3.0f) * d; Step[3].r = Step[0].r + m2; }
        { const float d = Step[1].g - Step[0].g; const float m1 = (1.0f / 3.0f) * d; Step[2].g = Step[0].g + m1; const float m2 = (2.0f / 3.0f) * d; Step[3].g = Step[0].g + m2; }
        { const float d = Step[1].b - Step[0].b; const float m1 = (1.0f / 3.0f) * d; Step[2].b = Step[0].b + m1; const float m2 = (2.0f / 3.0f) * d; Step[3].b = Step[0].b + m2; }
    }

    dxb_rgb Dir;
    Dir.r = Step[1].r - Step[0].r; Dir.g = Step[1].g - Step[0].g; Dir.b = Step[1].b - Step[0].b;

    const float fSteps = (float)(uSteps - 1);
    float fScale = 0.0f;
    if (wColorA != wColorB)
    {
        const float a = Dir.r * Dir.r, b = Dir.g * Dir.g, c = Dir.b * Dir.b;
        const float ab = a + b;
        fScale = fSteps / (ab + c);
    }
    Dir.r *= fScale; Dir.g *= fScale; Dir.b *= fScale;

    uint32_t dw = 0;
    if (dither) for (int i = 0; i < 16; ++i) { ErrR[i] = 0.0f; ErrG[i] = 0.0f; ErrB[i] = 0.0f; }

    for (int i = 0; i < 16; ++i)
    {
        if ((3 == uSteps) && (px[i].w < threshold))
        {
            dw = (3u << 30) | (dw >> 2);
        }
        else
        {
            float cr, cg, cb;
            if (uniform) { cr = px[i].x; cg = px[i].y; cb = px[i].z; }
            else { cr = px[i].x * DXB_LUM_R; cg = px[i].y * 1.0f; cb = px[i].z * DXB_LUM_B; }

            if (dither) { cr += ErrR[i]; cg += ErrG[i]; cb += ErrB[i]; }

            float fDot;
            {
                const float a = (cr - Step[0].r) * Dir.r, b = (cg - Step[0].g) * Dir.g, c = (cb - Step[0].b) * Dir.b;
                const float ab = a + b; fDot = ab + c;
            }

            uint32_t iStep;
            if (fDot <= 0.0f) iStep = 0;
            else if (fDot >= fSteps) iStep = 1;
            else iStep = (pStepsLut >> (2 * dxb_f2u(fDot + 0.5f))) & 3u;

            dw = (iStep << 30) | (dw >> 2);

            if (dither)
            {
                const float dr = 1.0f * (cr - Step[iStep].r), dg = 1.0f * (cg - Step[iStep].g), db = 1.0f * (cb - Step[iStep].b);
                dxb_diffuse1(ErrR, i, dr); dxb_diffuse1(ErrG, i, dg); dxb_diffuse1(ErrB, i, db);
            }
        }
    }
    *outBitmap = dw;
}

// D3DXEncodeBC1 (BC.cpp:738-795): optional alpha dithering, then EncodeBC1(colour-key on)
DXB_DEV void dxb_encode_bc1(const dxb_px* in, float threshold, uint32_t flags, uint8_t* out)
{
    dxb_px px[16];
    if (flags & DXB_BC_FLAGS_DITHER_A)
    {
        float fError[16];
        for (int i = 0; i < 16; ++i) fError[i] = 0.0f;
        for (int i = 0; i < 16; ++i)
        {
            const float fAlph = in[i].w + fError[i];
            px[i].x = in[i].x; px[i].y = in[i].y; px[i].z = in[i].z;
            { const float s = in[i].w + fError[i]; px[i].w = (float)dxb_f2i(s + 0.5f); }
            const float fDiff = fAlph - px[i].w;
            dxb_diffuse1(fError, i, fDiff);
        }
    }
    else
    {
        for (int i = 0; i < 16; ++i) px[i] = in[i];
    }
    uint32_t c0, c1, bm;
    dxb_encode_bc1_core(px, true, threshold, flags, &c0, &c1, &bm);
    uint32_t* o = (uint32_t*)out;
    o[0] = c0 | (c1 << 16);
    o[1] = bm;
}

// D3DXEncodeBC2 (BC.cpp:828-895)
DXB_DEV void dxb_encode_bc2(const dxb_px* in, uint32_t flags, uint8_t* out)
{
    uint32_t bitmap0 = 0, bitmap1 = 0;
    float fError[16];
    for (int i = 0; i < 16; ++i) fError[i] = 0.0f;
    for (int i = 0; i < 16; ++i)
    {
        float fAlph = in[i].w;
        if (flags & DXB_BC_FLAGS_DITHER_A) fAlph += fError[i];
        const float t = fAlph * 15.0f;
        const uint32_t u = dxb_f2u(t + 0.5f);
        if (i < 8) { bitmap0 >>= 4; bitmap0 |= (u << 28); }
        else { bitmap1 >>= 4; bitmap1 |= (u << 28); }
        if (flags & DXB_BC_FLAGS_DITHER_A)
        {
            const float q = (float)u * (1.0f / 15.0f);
            const float fDiff = fAlph - q;
            dxb_diffuse1(fError, i, fDiff);
        }
    }
    uint32_t c0, c1, bm;
    dxb_encode_bc1_core(in, false, 0.0f, flags, &c0, &c1, &bm);
    uint32_t* o = (uint32_t*)out;
    o[0] = bitmap0; o[1] = bitmap1; o[2] = c0 | (c1 << 16); o[3] = bm;
}

// OptimizeAlpha<bRange> (BC.h:187-311)
template <bool bRange>
DXB_DEV void dxb_optimize_alpha(float* pX, float* pY, const float* pPoints, uint32_t cSteps)
{
    float pC[8], pD[8];
    if (6 == cSteps)
    {
        pC[0] = 5.0f / 5.0f; pC[1] = 4.0f / 5.0f; pC[2] = 3.0f / 5.0f; pC[3] = 2.0f / 5.0f; pC[4] = 1.0f / 5.0f; pC[5] = 0.0f / 5.0f; pC[6] = 0.0f; pC[7] = 0.0f;
        pD[0] = 0.0f / 5.0f; pD[1] = 1.0f / 5.0f; pD[2] = 2.0f / 5.0f; pD[3] = 3.0f / 5.0f; pD[4] = 4.0f / 5.0f; pD[5] = 5.0f / 5.0f; pD[6] = 0.0f; pD[7] = 0.0f;
    }
    else
    {
        pC[0] = 7.0f / 7.0f; pC[1] = 6.0f / 7.0f; pC[2] = 5.0f / 7.0f; pC[3] = 4.0f / 7.0f; pC[4] = 3.0f / 7.0f; pC[5] = 2.0f / 7.0f; pC[6] = 1.0f / 7.0f; pC[7] = 0.0f / 7.0f;
        pD[0] = 0.0f / 7.0f; pD[1] = 1.0f / 7.0f; pD[2] = 2.0f / 7.0f; pD[3] = 3.0f / 7.0f; pD[4] = 4.0f / 7.0f; pD[5] = 5.0f / 7.0f; pD[6] = 6.0f / 7.0f; pD[7] = 7.0f / 7.0f;
    }

    const float MAX_VALUE = 1.0f;
    const float MIN_VALUE = bRange ? -1.0f : 0.0f;

    float fX = MAX_VALUE;
    float fY = MIN_VALUE;

    if (8 == cSteps)
    {
        for (int i = 0; i < 16; ++i)
        {
            if (pPoints[i] < fX) fX = pPoints[i];
            if (pPoints[i] > fY) fY = pPoints[i];
        }
    }
    else
    {
        for (int i = 0; i < 16; ++i)
        {
            if (pPoints[i] < fX && pPoints[i] > MIN_VALUE) fX = pPoints[i];
            if (pPoints[i] > fY && pPoints[i] < MAX_VALUE) fY = pPoints[i];
        }
        if (fX == fY) fY = MAX_VALUE;
    }

    const float fSteps = (float)(cSteps - 1);

#if DXB_ON_DEVICE
    #pragma unroll 1      // keeps the kernel small: straight-line code this long is instruction-fetch bound
#endif
    for (int iter = 0; iter < 8; ++iter)
    {
        if ((fY - fX) < (1.0f / 256.0f)) break;

        const float fScale = fSteps / (fY - fX);

        float pSteps[8];
        for (uint32_t s = 0; s < cSteps; ++s)
        {
            const float a = pC[s] * fX, b = pD[s] * fY;
            pSteps[s] = a + b;
        }
        if (6 == cSteps) { pSteps[6] = MIN_VALUE; pSteps[7] = MAX_VALUE; }

        float dX = 0.0f, dY = 0.0f, d2X = 0.0f, d2Y = 0.0f;

        for (int i = 0; i < 16; ++i)
        {
            const float fDot = (pPoints[i] - fX) * fScale;
            uint32_t iStep;
            if (fDot <= 0.0f)
                iStep = ((6 == cSteps) && (pPoints[i] <= (fX + MIN_VALUE) * 0.5f)) ? 6u : 0u;
            else if (fDot >= fSteps)
                iStep = ((6 == cSteps) && (pPoints[i] >= (fY + MAX_VALUE) * 0.5f)) ? 7u : (cSteps - 1);
            else
                iStep = dxb_f2u(fDot + 0.5f);

            if (iStep < cSteps)
            {
                const float fDiff = pSteps[iStep] - pPoints[i];
                dX += pC[iStep] * fDiff;
                d2X += pC[iStep] * pC[iStep];
                dY += pD[iStep] * fDiff;
                d2Y += pD[iStep] * pD[iStep];
            }
        }

        if (d2X > 0.0f) fX -= dX / d2X;
        if (d2Y > 0.0f) fY -= dY / d2Y;

        if (fX > fY) { const float f = fX; fX = fY; fY = f; }

        if ((dX * dX < (1.0f / 64.0f)) && (dY * dY < (1.0f / 64.0f))) break;
    }

    *pX = (fX < MIN_VALUE) ? MIN_VALUE : (fX > MAX_VALUE) ? MAX_VALUE : fX;
    *pY = (fY < MIN_VALUE) ? MIN_VALUE : (fY > MAX_VALUE) ? MAX_VALUE : fY;
}

// D3DXEncodeBC3 (BC.cpp:944-1141)
DXB_DEV void dxb_encode_bc3(const dxb_px* in, uint32_t flags, uint8_t* out)
{
    const bool ditherA = (flags & DXB_BC_FLAGS_DITHER_A) != 0;
    float fAlpha[16];
    float fError[16];
    for (int i = 0; i < 16; ++i) fError[i] = 0.0f;

    float fMinAlpha = in[0].w;
    float fMaxAlpha = in[0].w;

    for (int i = 0; i < 16; ++i)
    {
        float fAlph = in[i].w;
        if (ditherA) fAlph += fError[i];
        { const float t = fAlph * 255.0f; fAlpha[i] = (float)dxb_f2i(t + 0.5f) * (1.0f / 255.0f); }

        if (fAlpha[i] < fMinAlpha) fMinAlpha = fAlpha[i];
        else if (fAlpha[i] > fMaxAlpha) fMaxAlpha = fAlpha[i];

        if (ditherA)
        {
            const float fDiff = fAlph - fAlpha[i];
            dxb_diffuse1(fError, i, fDiff);
        }
    }

    uint32_t c0, c1, bm;
    dxb_encode_bc1_core(in, false, 0.0f, flags, &c0, &c1, &bm);
    uint32_t* o = (uint32_t*)out;
    o[2] = c0 | (c1 << 16);
    o[3] = bm;

    if (1.0f == fMinAlpha)
    {
        o[0] = 0x0000ffffu; o[1] = 0;
        return;
    }

    const uint32_t uSteps = ((0.0f == fMinAlpha) || (1.0f == fMaxAlpha)) ? 6u : 8u;

    float fAlphaA, fAlphaB;
    dxb_optimize_alpha<false>(&fAlphaA, &fAlphaB, fAlpha, uSteps);

    uint32_t bAlphaA, bAlphaB;
    { const float t = fAlphaA * 255.0f; bAlphaA = (uint32_t)dxb_f2i(t + 0.5f) & 0xFF; }
    { const float t = fAlphaB * 255.0f; bAlphaB = (uint32_t)dxb_f2i(t + 0.5f) & 0xFF; }

    fAlphaA = (float)bAlphaA * (1.0f / 255.0f);
    fAlphaB = (float)bAlphaB * (1.0f / 255.0f);

    if ((8 == uSteps) && (bAlphaA == bAlphaB))
    {
        o[0] = bAlphaA | (bAlphaB << 8); o[1] = 0;
        return;
    }

    // pSteps6 = {0,2,3,4,5,1}; pSteps8 = {0,2,3,4,5,6,7,1}   (BC.cpp:1050-1051): 3 bits per entry
    uint32_t lut;
    float fStep[8];
    for (int i = 0; i < 8; ++i) fStep[i] = 0.0f;
    uint32_t a0, a1;

    if (6 == uSteps)
    {
        a0 = bAlphaA; a1 = bAlphaB;
        fStep[0] = fAlphaA; fStep[1] = fAlphaB;
        for (int i = 1; i < 5; ++i)
        {
            const float a = fStep[0] * (float)(5 - i), b = fStep[1] * (float)i;
            fStep[i + 1] = (a + b) * (1.0f / 5.0f);
        }
        fStep[6] = 0.0f; fStep[7] = 1.0f;
        lut = 0u | (2u << 3) | (3u << 6) | (4u << 9) | (5u << 12) | (1u << 15);
    }
    else
    {
        a0 = bAlphaB; a1 = bAlphaA;
        fStep[0] = fAlphaB; fStep[1] = fAlphaA;
        for (int i = 1; i < 7; ++i)
        {
            const float a = fStep[0] * (float)(7 - i), b = fStep[1] * (float)i;
            fStep[i + 1] = (a + b) * (1.0f / 7.0f);
        }
        lut = 0u | (2u << 3) | (3u << 6) | (4u << 9) | (5u << 12) | (6u << 15) | (7u << 18) | (1u << 21);
    }

    const float fSteps = (float)(uSteps - 1);
    const float fScale = (fStep[0] != fStep[1]) ? (fSteps / (fStep[1] - fStep[0])) : 0.0f;

    if (ditherA) for (int i = 0; i < 16; ++i) fError[i] = 0.0f;

    uint32_t halves[2];
    for (int iSet = 0; iSet < 2; ++iSet)
    {
        uint32_t dw = 0;
        const int iMin = iSet * 8, iLim = iMin + 8;
        for (int i = iMin; i < iLim; ++i)
        {
            float fAlph = in[i].w;
            if (ditherA) fAlph += fError[i];
            const float fDot = (fAlph - fStep[0]) * fScale;

            uint32_t iStep;
            if (fDot <= 0.0f)
                iStep = ((6 == uSteps) && (fAlph <= fStep[0] * 0.5f)) ? 6u : 0u;
            else if (fDot >= fSteps)
                iStep = ((6 == uSteps) && (fAlph >= (fStep[1] + 1.0f) * 0.5f)) ? 7u : 1u;
            else
                iStep = (lut >> (3 * dxb_f2u(fDot + 0.5f))) & 7u;

            dw = (iStep << 21) | (dw >> 3);

            if (ditherA)
            {
                const float fDiff = (fAlph - fStep[iStep]);
                dxb_diffuse1(fError, i, fDiff);
            }
        }
        halves[iSet] = dw & 0xFFFFFFu;
    }
    // alpha[0], alpha[1], bitmap[0..5]
    o[0] = a0 | (a1 << 8) | ((halves[0] & 0xFFFF) << 16);
    o[1] = (halves[0] >> 16) | (halves[1] << 8);
}

// ---------------------------------------------------------------------------------------------
// BC4 / BC5 (BC4BC5.cpp)

// BC4_UNORM::DecodeFromIndex (BC4BC5.cpp:47-69)
DXB_DEV float dxb_bc4u_decode(uint32_t red_0, uint32_t red_1, uint32_t uIndex)
{
    if (uIndex == 0) return (float)red_0 / 255.0f;
    if (uIndex == 1) return (float)red_1 / 255.0f;
    const float fred_0 = (float)red_0 / 255.0f;
    const float fred_1 = (float)red_1 / 255.0f;
    if (red_0 > red_1)
    {
        uIndex -= 1;
        const float a = fred_0 * (float)(7u - uIndex), b = fred_1 * (float)uIndex;
        return (a + b) / 7.0f;
    }
    else
    {
        if (uIndex == 6) return 0.0f;
        if (uIndex == 7) return 1.0f;
        uIndex -= 1;
        const float a = fred_0 * (float)(5u - uIndex), b = fred_1 * (float)uIndex;
        return (a + b) / 5.0f;
    }
}
// BC4_SNORM::DecodeFromIndex (BC4BC5.cpp:103-128); red_0/red_1 are int8 values
DXB_DEV float dxb_bc4s_decode(int32_t red_0, int32_t red_1, uint32_t uIndex)
{
    const int32_t sred_0 = (red_0 == -128) ? -127 : red_0;
    const int32_t sred_1 = (red_1 == -128) ? -127 : red_1;
    if (uIndex == 0) return (float)sred_0 / 127.0f;
    if (uIndex == 1) return (float)sred_1 / 127.0f;
    const float fred_0 = (float)sred_0 / 127.0f;
    const float fred_1 = (float)sred_1 / 127.0f;
    if (red_0 > red_1)
    {
        uIndex -= 1;
        const float a = fred_0 * (float)(7u - uIndex), b = fred_1 * (float)uIndex;
        return (a + b) / 7.0f;
    }
    else
    {
        if (uIndex == 6) return -1.0f;
        if (uIndex == 7) return 1.0f;
        uIndex -= 1;
        const float a = fred_0 * (float)(5u - uIndex), b = fred_1 * (float)uIndex;
        return (a + b) / 5.0f;
    }
}

// FloatToSNorm (BC4BC5.cpp:158-179)
DXB_DEV int32_t dxb_float_to_snorm8(float fVal)
{
    if (fVal != fVal) fVal = 0.0f;
    else if (fVal > 1.0f) fVal = 1.0f;
    else if (fVal < -1.0f) fVal = -1.0f;
    fVal = fVal * 127.0f;
    if (fVal >= 0.0f) fVal += 0.5f; else fVal -= 0.5f;
    return (int32_t)(int8_t)dxb_f2i(fVal);
}

// 48-bit index search shared by FindClosestUNORM / FindClosestSNORM (BC4BC5.cpp:325-377)
DXB_DEV uint64_t dxb_bc4_indices(const float* rGradient, const float* texels)
{
    uint64_t data = 0;
    for (int i = 0; i < 16; ++i)
    {
        uint32_t uBestIndex = 0;
        float fBestDelta = 100000.0f;
        for (uint32_t k = 0; k < 8; ++k)
        {
            const float fCurrentDelta = fabsf(rGradient[k] - texels[i]);
            if (fCurrentDelta < fBestDelta) { uBestIndex = k; fBestDelta = fCurrentDelta; }
        }
        data |= ((uint64_t)uBestIndex) << (3 * i + 16);
    }
    return data;
}

// D3DXEncodeBC4U: FindEndPointsBC4U (BC4BC5.cpp:183-236) + FindClosestUNORM
DXB_DEV uint64_t dxb_encode_bc4u(const float* t)
{
    float fBlockMax = t[0], fBlockMin = t[0];
    for (int i = 0; i < 16; ++i)
    {
        if (t[i] < fBlockMin) fBlockMin = t[i];
        else if (t[i] > fBlockMax) fBlockMax = t[i];
    }
    const bool bUsing4BlockCodec = (0.0f == fBlockMin || 1.0f == fBlockMax);

    float fStart, fEnd;
    uint32_t e0, e1;
    if (!bUsing4BlockCodec)
    {
        dxb_optimize_alpha<false>(&fStart, &fEnd, t, 8);
        const uint32_t iStart = (uint32_t)dxb_f2i(fStart * 255.0f) & 0xFF;
        const uint32_t iEnd = (uint32_t)dxb_f2i(fEnd * 255.0f) & 0xFF;
        e0 = iEnd; e1 = iStart;
    }
    else
    {
        dxb_optimize_alpha<false>(&fStart, &fEnd, t, 6);
        const uint32_t iStart = (uint32_t)dxb_f2i(fStart * 255.0f) & 0xFF;
        const uint32_t iEnd = (uint32_t)dxb_f2i(fEnd * 255.0f) & 0xFF;
        e1 = iEnd; e0 = iStart;
    }
    float rGradient[8];
    for (uint32_t k = 0; k < 8; ++k) rGradient[k] = dxb_bc4u_decode(e0, e1, k);
    return (uint64_t)e0 | ((uint64_t)e1 << 8) | dxb_bc4_indices(rGradient, t);
}

// D3DXEncodeBC4S: FindEndPointsBC4S (BC4BC5.cpp:238-293) + FindClosestSNORM
DXB_DEV uint64_t dxb_encode_bc4s(const float* t)
{
    float fBlockMax = t[0], fBlockMin = t[0];
    for (int i = 0; i < 16; ++i)
    {
        if (t[i] < fBlockMin) fBlockMin = t[i];
        else if (t[i] > fBlockMax) fBlockMax = t[i];
    }
    const bool bUsing4BlockCodec = (-1.0f == fBlockMin || 1.0f == fBlockMax);

    float fStart, fEnd;
    int32_t e0, e1;
    if (!bUsing4BlockCodec)
    {
        dxb_optimize_alpha<true>(&fStart, &fEnd, t, 8);
        const int32_t iStart = dxb_float_to_snorm8(fStart), iEnd = dxb_float_to_snorm8(fEnd);
        e0 = iEnd; e1 = iStart;
    }
    else
    {
        dxb_optimize_alpha<true>(&fStart, &fEnd, t, 6);
        const int32_t iStart = dxb_float_to_snorm8(fStart), iEnd = dxb_float_to_snorm8(fEnd);
        e1 = iEnd; e0 = iStart;
    }
    float rGradient[8];
    for (uint32_t k = 0; k < 8; ++k) rGradient[k] = dxb_bc4s_decode(e0, e1, k);
    return (uint64_t)((uint32_t)e0 & 0xFF) | ((uint64_t)((uint32_t)e1 & 0xFF) << 8) | dxb_bc4_indices(rGradient, t);
}

// ---------------------------------------------------------------------------------------------
// Format dispatch for one block (DetermineEncoderSettings, DirectXTexCompress.cpp:46-68)
DXB_DEV void dxb_encode_block_bc15(uint32_t dstFmt, const dxb_px* px, uint32_t bcflags, float threshold, uint8_t* out)
{
    switch (dstFmt)
    {
    case DXB_FMT_BC1_UNORM: case DXB_FMT_BC1_UNORM_SRGB:
        dxb_encode_bc1(px, threshold, bcflags, out); break;
    case DXB_FMT_BC2_UNORM: case DXB_FMT_BC2_UNORM_SRGB:
        dxb_encode_bc2(px, bcflags, out); break;
    case DXB_FMT_BC3_UNORM: case DXB_FMT_BC3_UNORM_SRGB:
        dxb_encode_bc3(px, bcflags, out); break;
    case DXB_FMT_BC4_UNORM:
    {
        float t[16]; for (int i = 0; i < 16; ++i) t[i] = px[i].x;
        *(uint64_t*)out = dxb_encode_bc4u(t); break;
    }
    case DXB_FMT_BC4_SNORM:
    {
        float t[16]; for (int i = 0; i < 16; ++i) t[i] = px[i].x;
        *(uint64_t*)out = dxb_encode_bc4s(t); break;
    }
    case DXB_FMT_BC5_UNORM:
    {
        float t[16];
        for (int i = 0; i < 16; ++i) t[i] = px[i].x;
        ((uint64_t*)out)[0] = dxb_encode_bc4u(t);
        for (int i = 0; i < 16; ++i) t[i] = px[i].y;
        ((uint64_t*)out)[1] = dxb_encode_bc4u(t); break;
    }
    case DXB_FMT_BC5_SNORM:
    {
        float t[16];
        for (int i = 0; i < 16; ++i) t[i] = px[i].x;
        ((uint64_t*)out)[0] = dxb_encode_bc4s(t);
        for (int i = 0; i < 16; ++i) t[i] = px[i].y;
        ((uint64_t*)out)[1] = dxb_encode_bc4s(t); break;
    }
    default: break;
    }
}
