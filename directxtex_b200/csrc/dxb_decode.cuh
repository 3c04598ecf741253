// dxb_decode.cuh — BC1..BC7 block DEcoders, one thread per 4x4 block (DirectX::Decompress).
// Restates bit-exactly (same fp32 statements, unfused):
//   DecodeBC1 / D3DXDecodeBC1/2/3          BC.cpp:318-366, 731-735, 802-825, 902-941
//   D3DXDecodeBC4U/S, BC5U/S               BC4BC5.cpp:389-416, 465-494 (DecodeFromIndex :47-69, 103-128)
//   D3DX_BC7::Decode                       BC6HBC7.cpp:2566-2780   (LDRColorA -> HDRColorA :427-433)
//   D3DX_BC6H::Decode                      BC6HBC7.cpp:1658-1813
// Output = 16 RGBA fp32 pixels exactly as the reference hands them to ConvertScanline / StoreScanline
// (DecompressBC, DirectXTexCompress.cpp:488-528).
#pragma once
#include "dxb_pixel.cuh"
#include "dxb_bc15.cuh"            // dxb_bc4u_decode / dxb_bc4s_decode
#include "dxb_bc7.cuh"             // dxb_bc7_weight, dxb_bc7_unq
#include "dxb_bc6h.cuh"            // unquantize / finish helpers, mode tables

DXB_DEV float dxb_lerp1(float a, float b, float t) { const float l = b - a; const float m = l * t; return m + a; }   // XMVectorLerp

// DecodeBC1 (BC.cpp:318-366); blk = 8 bytes.  The four colours of the block ...
DXB_DEV void dxb_bc1_palette(const uint8_t* blk, bool isbc1, dxb_px* clr)
{
    const uint32_t c01 = ((const uint32_t*)blk)[0];
    const uint32_t rgb0 = c01 & 0xFFFFu, rgb1 = c01 >> 16;
    // XMLoadU565: x = bits 0-4, y = 5-10, z = 11-15; * {1/31, 1/63, 1/31, 1}; swizzle <2,1,0,3>; w = 1
    clr[0] = dxb_make_px((float)((rgb0 >> 11) & 31u) * (1.0f / 31.0f), (float)((rgb0 >> 5) & 63u) * (1.0f / 63.0f), (float)(rgb0 & 31u) * (1.0f / 31.0f), 1.0f);
    clr[1] = dxb_make_px((float)((rgb1 >> 11) & 31u) * (1.0f / 31.0f), (float)((rgb1 >> 5) & 63u) * (1.0f / 63.0f), (float)(rgb1 & 31u) * (1.0f / 31.0f), 1.0f);
    const dxb_px clr0 = clr[0], clr1 = clr[1];
    if (isbc1 && (rgb0 <= rgb1))
    {
        clr[2] = dxb_make_px(dxb_lerp1(clr0.x, clr1.x, 0.5f), dxb_lerp1(clr0.y, clr1.y, 0.5f), dxb_lerp1(clr0.z, clr1.z, 0.5f), dxb_lerp1(clr0.w, clr1.w, 0.5f));
        clr[3] = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f);
    }
    else
    {
        const float t1 = 1.0f / 3.0f, t2 = 2.0f / 3.0f;
        clr[2] = dxb_make_px(dxb_lerp1(clr0.x, clr1.x, t1), dxb_lerp1(clr0.y, clr1.y, t1), dxb_lerp1(clr0.z, clr1.z, t1), dxb_lerp1(clr0.w, clr1.w, t1));
        clr[3] = dxb_make_px(dxb_lerp1(clr0.x, clr1.x, t2), dxb_lerp1(clr0.y, clr1.y, t2), dxb_lerp1(clr0.z, clr1.z, t2), dxb_lerp1(clr0.w, clr1.w, t2));
    }
}
// ... and the 16 pixels they are assigned to
DXB_DEV void dxb_decode_bc1(const uint8_t* blk, bool isbc1, dxb_px* out)
{
    dxb_px clr[4];
    dxb_bc1_palette(blk, isbc1, clr);
    const dxb_px clr0 = clr[0], clr1 = clr[1], clr2 = clr[2], clr3 = clr[3];
    uint32_t dw = ((const uint32_t*)blk)[1];
    for (int i = 0; i < 16; ++i, dw >>= 2)
    {
        const uint32_t k = dw & 3u;
        out[i] = (k == 0) ? clr0 : (k == 1) ? clr1 : (k == 2) ? clr2 : clr3;
    }
}

DXB_DEV void dxb_decode_bc2(const uint8_t* blk, dxb_px* out)
{
    dxb_decode_bc1(blk + 8, false, out);
    uint32_t dw = ((const uint32_t*)blk)[0];
    for (int i = 0; i < 8; ++i, dw >>= 4) out[i].w = (float)(dw & 0xFu) * (1.0f / 15.0f);
    dw = ((const uint32_t*)blk)[1];
    for (int i = 8; i < 16; ++i, dw >>= 4) out[i].w = (float)(dw & 0xFu) * (1.0f / 15.0f);
}

// the eight alpha values of a BC3 block (BC.cpp:902-941)
DXB_DEV void dxb_bc3_alpha_table(const uint8_t* blk, float* fAlpha)
{
    const uint32_t a0 = blk[0], a1 = blk[1];
    fAlpha[0] = (float)a0 * (1.0f / 255.0f);
    fAlpha[1] = (float)a1 * (1.0f / 255.0f);
    if (a0 > a1)
    {
        for (int i = 1; i < 7; ++i)
        {
            const float x = fAlpha[0] * (float)(7 - i), y = fAlpha[1] * (float)i;
            fAlpha[i + 1] = (x + y) * (1.0f / 7.0f);
        }
    }
    else
    {
        for (int i = 1; i < 5; ++i)
        {
            const float x = fAlpha[0] * (float)(5 - i), y = fAlpha[1] * (float)i;
            fAlpha[i + 1] = (x + y) * (1.0f / 5.0f);
        }
        fAlpha[6] = 0.0f; fAlpha[7] = 1.0f;
    }
}
DXB_DEV void dxb_decode_bc3(const uint8_t* blk, dxb_px* out)
{
    dxb_decode_bc1(blk + 8, false, out);
    float fAlpha[8];
    const uint32_t a0 = blk[0], a1 = blk[1];
    fAlpha[0] = (float)a0 * (1.0f / 255.0f);
    fAlpha[1] = (float)a1 * (1.0f / 255.0f);
    if (a0 > a1)
    {
        for (int i = 1; i < 7; ++i)
        {
            const float x = fAlpha[0] * (float)(7 - i), y = fAlpha[1] * (float)i;
            fAlpha[i + 1] = (x + y) * (1.0f / 7.0f);
        }
    }
    else
    {
        for (int i = 1; i < 5; ++i)
        {
            const float x = fAlpha[0] * (float)(5 - i), y = fAlpha[1] * (float)i;
            fAlpha[i + 1] = (x + y) * (1.0f / 5.0f);
        }
        fAlpha[6] = 0.0f; fAlpha[7] = 1.0f;
    }
    uint32_t dw = (uint32_t)blk[2] | ((uint32_t)blk[3] << 8) | ((uint32_t)blk[4] << 16);
    for (int i = 0; i < 8; ++i, dw >>= 3) out[i].w = fAlpha[dw & 7u];
    dw = (uint32_t)blk[5] | ((uint32_t)blk[6] << 8) | ((uint32_t)blk[7] << 16);
    for (int i = 8; i < 16; ++i, dw >>= 3) out[i].w = fAlpha[dw & 7u];
}

// the eight values of one BC4 channel
DXB_DEV void dxb_bc4_table(const uint8_t* blk, bool bSigned, float* grad)
{
    if (bSigned)
    {
        const int32_t r0 = (int32_t)(int8_t)blk[0], r1 = (int32_t)(int8_t)blk[1];
        for (uint32_t k = 0; k < 8; ++k) grad[k] = dxb_bc4s_decode(r0, r1, k);
    }
    else
    {
        const uint32_t r0 = blk[0], r1 = blk[1];
        for (uint32_t k = 0; k < 8; ++k) grad[k] = dxb_bc4u_decode(r0, r1, k);
    }
}
// one BC4 channel: 16 values
DXB_DEV void dxb_decode_bc4_channel(const uint8_t* blk, bool bSigned, float* v)
{
    const uint64_t data = *(const uint64_t*)blk;
    float grad[8];
    if (bSigned)
    {
        const int32_t r0 = (int32_t)(int8_t)blk[0], r1 = (int32_t)(int8_t)blk[1];
        for (uint32_t k = 0; k < 8; ++k) grad[k] = dxb_bc4s_decode(r0, r1, k);
    }
    else
    {
        const uint32_t r0 = blk[0], r1 = blk[1];
        for (uint32_t k = 0; k < 8; ++k) grad[k] = dxb_bc4u_decode(r0, r1, k);
    }
    for (int i = 0; i < 16; ++i) v[i] = grad[(data >> (3 * i + 16)) & 7u];
}

// 128-bit little-endian bit reader
struct dxb_bits { uint64_t lo, hi; uint32_t pos; };
DXB_DEV uint32_t dxb_get_bits(dxb_bits* b, uint32_t n)
{
    if (n == 0) return 0;
    uint64_t v;
    const uint32_t p = b->pos;
    if (p < 64)
    {
        v = b->lo >> p;
        if (p + n > 64) v |= b->hi << (64 - p);
    }
    else v = b->hi >> (p - 64);
    b->pos = p + n;
    return (uint32_t)(v & ((1ull << n) - 1ull));
}

// D3DX_BC7::Decode (BC6HBC7.cpp:2566-2780)
DXB_DEV void dxb_decode_bc7(const uint8_t* blk, dxb_px* out)
{
    dxb_bits B; B.lo = ((const uint64_t*)blk)[0]; B.hi = ((const uint64_t*)blk)[1]; B.pos = 0;
    uint32_t mode = 0;
    while (mode < 8 && !((B.lo >> mode) & 1ull)) ++mode;          // first set bit of the low byte (bit 8+ irrelevant: mode < 8)
    if (mode >= 8)
    {
        for (int i = 0; i < 16; ++i) out[i] = dxb_make_px(0.0f, 0.0f, 0.0f, 0.0f);    // reserved mode: transparent black (:2773-2779)
        return;
    }
    B.pos = mode + 1;
    // mode table (BC6HBC7.cpp:1106-1124): partitions-1, partition bits, p bits, rotation bits, idx mode bits, idx prec, idx prec 2, rgb prec, a prec
    const uint32_t nsM1 = (mode == 0 || mode == 2) ? 2u : ((mode == 1 || mode == 3 || mode == 7) ? 1u : 0u);
    const uint32_t partBits = (mode == 0) ? 4u : ((nsM1 > 0) ? 6u : 0u);
    const uint32_t nPB = (mode == 0) ? 6u : (mode == 1) ? 2u : (mode == 3 || mode == 7) ? 4u : (mode == 6) ? 2u : 0u;
    const uint32_t rotBits = (mode == 4 || mode == 5) ? 2u : 0u;
    const uint32_t imBits = (mode == 4) ? 1u : 0u;
    const uint32_t ip = (mode == 0 || mode == 1) ? 3u : (mode == 6) ? 4u : 2u;
    const uint32_t ip2 = (mode == 4) ? 3u : (mode == 5) ? 2u : 0u;
    const uint32_t cp = (mode == 0) ? 4u : (mode == 1) ? 6u : (mode == 2) ? 5u : (mode == 3) ? 7u : (mode == 4) ? 5u : (mode == 5) ? 7u : (mode == 6) ? 7u : 5u;
    const uint32_t ap = (mode == 4) ? 6u : (mode == 5) ? 8u : (mode == 6) ? 7u : (mode == 7) ? 5u : 0u;
    const uint32_t shape = dxb_get_bits(&B, partBits);
    const uint32_t rot = dxb_get_bits(&B, rotBits);
    const uint32_t im = dxb_get_bits(&B, imBits);
    const uint32_t nep = (nsM1 + 1u) * 2u;
    uint32_t c[6][4];
    for (uint32_t ch = 0; ch < 3; ++ch) for (uint32_t e = 0; e < nep; ++e) c[e][ch] = dxb_get_bits(&B, cp);
    for (uint32_t e = 0; e < nep; ++e) c[e][3] = ap ? dxb_get_bits(&B, ap) : 255u;
    uint32_t P[6] = { 0, 0, 0, 0, 0, 0 };
    for (uint32_t i = 0; i < nPB; ++i) P[i] = dxb_get_bits(&B, 1);
    const uint32_t cpp = cp + (nPB ? 1u : 0u), app = ap ? (ap + (nPB ? 1u : 0u)) : 0u;
    for (uint32_t e = 0; e < nep; ++e)
    {
        if (nPB)
        {
            const uint32_t pi = e * nPB / nep;
            for (uint32_t ch = 0; ch < 3; ++ch) c[e][ch] = (c[e][ch] << 1) | P[pi];
            if (ap) c[e][3] = (c[e][3] << 1) | P[pi];
        }
        for (uint32_t ch = 0; ch < 3; ++ch) c[e][ch] = dxb_bc7_unq(c[e][ch], cpp);
        c[e][3] = app ? dxb_bc7_unq(c[e][3], app) : 255u;
    }
    // partition / anchors
    uint32_t a1 = 0, a2 = 0;
    if (nsM1 == 1) a1 = dxb_anchor2[shape];
    else if (nsM1 == 2) { a1 = dxb_anchor3a[shape]; a2 = dxb_anchor3b[shape]; }
    uint32_t w1[16], w2[16];
    for (uint32_t i = 0; i < 16; ++i)
    {
        const bool fix = (i == 0) || (nsM1 >= 1 && i == a1) || (nsM1 == 2 && i == a2);
        w1[i] = dxb_get_bits(&B, fix ? ip - 1u : ip);
    }
    if (ip2) for (uint32_t i = 0; i < 16; ++i) w2[i] = dxb_get_bits(&B, i ? ip2 : ip2 - 1u);
    for (uint32_t i = 0; i < 16; ++i)
    {
        const uint32_t region = (nsM1 == 0) ? 0u : (nsM1 == 1) ? ((dxb_part2[shape] >> i) & 1u) : ((dxb_part3[shape] >> (2 * i)) & 3u);
        const uint32_t* e0 = c[region * 2], *e1 = c[region * 2 + 1];
        uint32_t wc, wa, pc, pa;
        if (ip2 == 0) { wc = w1[i]; wa = w1[i]; pc = ip; pa = ip; }
        else if (im == 0) { wc = w1[i]; wa = w2[i]; pc = ip; pa = ip2; }
        else { wc = w2[i]; wa = w1[i]; pc = ip2; pa = ip; }
        const uint32_t kc = dxb_bc7_weight(pc, wc), ka = dxb_bc7_weight(pa, wa);
        uint32_t r = (e0[0] * (64u - kc) + e1[0] * kc + 32u) >> 6;
        uint32_t g = (e0[1] * (64u - kc) + e1[1] * kc + 32u) >> 6;
        uint32_t b = (e0[2] * (64u - kc) + e1[2] * kc + 32u) >> 6;
        uint32_t a = (e0[3] * (64u - ka) + e1[3] * ka + 32u) >> 6;
        r &= 0xFF; g &= 0xFF; b &= 0xFF; a &= 0xFF;
        if (rot == 1) { const uint32_t t = r; r = a; a = t; }
        else if (rot == 2) { const uint32_t t = g; g = a; a = t; }
        else if (rot == 3) { const uint32_t t = b; b = a; a = t; }
        out[i] = dxb_make_px((float)r * (1.0f / 255.0f), (float)g * (1.0f / 255.0f), (float)b * (1.0f / 255.0f), (float)a * (1.0f / 255.0f));
    }
}

// D3DX_BC6H::Decode (BC6HBC7.cpp:1658-1813)
DXB_DEV void dxb_decode_bc6h(const uint8_t* blk, bool bSigned, dxb_px* out)
{
    dxb_bits B; B.lo = ((const uint64_t*)blk)[0]; B.hi = ((const uint64_t*)blk)[1]; B.pos = 0;
    uint32_t m = dxb_get_bits(&B, 2);
    if (m != 0 && m != 1) m = (dxb_get_bits(&B, 3) << 2) | m;
    int mi = -1;
    for (int k = 0; k < 14; ++k) if ((dxb_bc6h_info[k] & 31u) == m) mi = k;
    if (mi < 0)
    {
        for (int i = 0; i < 16; ++i) out[i] = dxb_make_px(0.0f, 0.0f, 0.0f, 1.0f);      // reserved / invalid mode: opaque black
        return;
    }
    const uint32_t info = dxb_bc6h_info[mi];
    const bool two = ((info >> 5) & 1u) != 0u, transformed = ((info >> 6) & 1u) != 0u;
    const int32_t prec = (int32_t)((info >> 8) & 31u);
    const int32_t db[3] = { (int32_t)((info >> 16) & 15u), (int32_t)((info >> 20) & 15u), (int32_t)((info >> 24) & 15u) };
    int32_t ep[4][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    uint32_t shape = 0;
    const uint32_t hdr = two ? 82u : 65u;
    for (uint32_t b = B.pos; b < hdr; ++b)
    {
        const uint32_t bit = (b < 64) ? (uint32_t)((B.lo >> b) & 1ull) : (uint32_t)((B.hi >> (b - 64)) & 1ull);
        if (!bit) continue;
        const uint32_t d = dxb_bc6h_desc[mi][b];
        const uint32_t f = d >> 4, fb = d & 15u;
        if (f == 2) shape |= 1u << fb;
        else if (f >= 3) { const uint32_t ch = (f - 3u) >> 2, e = (f - 3u) & 3u; ep[e][ch] |= (int32_t)(1u << fb); }
    }
    B.pos = hdr;
    #define DXB_SEXT(x, nb) (((x) & (1 << ((nb) - 1))) ? ((~0) ^ ((1 << (nb)) - 1)) | (x) : (x))
    if (bSigned) for (int c = 0; c < 3; ++c) ep[0][c] = DXB_SEXT(ep[0][c], prec);
    if (bSigned || transformed)
    {
        const int np = two ? 2 : 1;
        for (int p = 0; p < np; ++p)
            for (int c = 0; c < 3; ++c)
            {
                const int32_t nb = transformed ? db[c] : prec;
                if (p != 0) ep[2][c] = DXB_SEXT(ep[2][c], nb);
                ep[p * 2 + 1][c] = DXB_SEXT(ep[p * 2 + 1][c], nb);
            }
    }
    if (transformed)
    {
        for (int c = 0; c < 3; ++c)
        {
            const int32_t mask = (1 << prec) - 1;
            ep[1][c] = (ep[1][c] + ep[0][c]) & mask;
            ep[2][c] = (ep[2][c] + ep[0][c]) & mask;
            ep[3][c] = (ep[3][c] + ep[0][c]) & mask;
            if (bSigned) { ep[1][c] = DXB_SEXT(ep[1][c], prec); ep[2][c] = DXB_SEXT(ep[2][c], prec); ep[3][c] = DXB_SEXT(ep[3][c], prec); }
        }
    }
    #undef DXB_SEXT
    const uint32_t ib = two ? 3u : 4u;
    const uint32_t anchor1 = two ? dxb_anchor2[shape & 31u] : 0u;
    for (uint32_t i = 0; i < 16; ++i)
    {
        const bool fix = (i == 0) || (two && i == anchor1);
        const uint32_t idx = dxb_get_bits(&B, fix ? ib - 1u : ib);
        const uint32_t region = two ? ((dxb_part2[shape & 31u] >> i) & 1u) : 0u;
        const int32_t w = (int32_t)dxb_bc7_weight(ib, idx);
        int32_t fc[3];
        for (int c = 0; c < 3; ++c)
        {
            const int32_t u1 = dxb_bc6h_unquantize(ep[region * 2][c], prec, bSigned), u2 = dxb_bc6h_unquantize(ep[region * 2 + 1][c], prec, bSigned);
            fc[c] = dxb_bc6h_finish((u1 * (64 - w) + u2 * w + 32) >> 6, bSigned);
        }
        float rgb[3];
        for (int c = 0; c < 3; ++c)
        {
            uint16_t h;
            if (bSigned) { int32_t v = fc[c]; uint32_t s = 0; if (v < 0) { s = 0x8000u; v = -v; } h = (uint16_t)(s | (uint32_t)v); }
            else h = (uint16_t)fc[c];
            rgb[c] = dxb_half_to_float(h);
        }
        out[i] = dxb_make_px(rgb[0], rgb[1], rgb[2], 1.0f);
    }
}

// decode one block of any BC format into 16 RGBA fp32 pixels
DXB_DEV void dxb_decode_block(uint32_t fmt, const uint8_t* blk, dxb_px* out)
{
    switch (fmt)
    {
    case DXB_FMT_BC1_UNORM: case DXB_FMT_BC1_UNORM_SRGB: dxb_decode_bc1(blk, true, out); break;
    case DXB_FMT_BC2_UNORM: case DXB_FMT_BC2_UNORM_SRGB: dxb_decode_bc2(blk, out); break;
    case DXB_FMT_BC3_UNORM: case DXB_FMT_BC3_UNORM_SRGB: dxb_decode_bc3(blk, out); break;
    case DXB_FMT_BC4_UNORM: case DXB_FMT_BC4_SNORM:
    {
        float v[16];
        dxb_decode_bc4_channel(blk, fmt == DXB_FMT_BC4_SNORM, v);
        for (int i = 0; i < 16; ++i) out[i] = dxb_make_px(v[i], 0.0f, 0.0f, 1.0f);
        break;
    }
    case DXB_FMT_BC5_UNORM: case DXB_FMT_BC5_SNORM:
    {
        float u[16], v[16];
        dxb_decode_bc4_channel(blk, fmt == DXB_FMT_BC5_SNORM, u);
        dxb_decode_bc4_channel(blk + 8, fmt == DXB_FMT_BC5_SNORM, v);
        for (int i = 0; i < 16; ++i) out[i] = dxb_make_px(u[i], v[i], 0.0f, 1.0f);
        break;
    }
    case DXB_FMT_BC6H_UF16: dxb_decode_bc6h(blk, false, out); break;
    case DXB_FMT_BC6H_SF16: dxb_decode_bc6h(blk, true, out); break;
    default: dxb_decode_bc7(blk, out); break;
    }
}
