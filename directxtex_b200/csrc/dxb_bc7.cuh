// dxb_bc7.cuh — BC7 block encoder, ONE WARP PER 4x4 BLOCK (single-source SPMD, see dxb_warp.cuh).
//
// What it replaces: D3DXEncodeBC7 -> D3DX_BC7::Encode (BC6HBC7.cpp:3654-3659, 2783-2889).
// Parity contract (north_star / SURVEY.md 8(d)): NOT bit-exact; the decoded result must be a valid BC7
// stream for the reference decoder (D3DX_BC7::Decode, BC6HBC7.cpp:2566-2780) and its RGBA MSE against the
// source must stay within the tolerance stated in DESIGN.md of the MSE the reference CPU encoder
// reaches on the same input.  The reference's search (Newton fit + rank 64 shapes + refine 16 with
// +-5 exhaustive perturbation, ~7 ms/block on one CPU core) is replaced by a search shaped for a warp:
//
//   stage 0  LDR pixels exactly as the reference quantises them: uint8(clamp(c*255 + 0.01))   (:2792-2799)
//   stage 1  all 64 two-subset shapes ranked by a closed-form line-fit residual from per-subset
//            second moments (2 shapes per lane), candidates kept in registers, integer-key warp min
//   stage 2  32 lane tasks evaluated concurrently, one (mode, shape, subset | rotation | p-bits) each:
//              opaque block : 7 best shapes x 2 subsets x {mode 1, mode 3}  +  mode 6 x 4 p-bit pairs
//              alpha block  : 8 best shapes x 2 subsets x mode 7, mode 6 x 4 p-bit pairs,
//                             mode 5 x 4 rotations, mode 4 x 4 rotations x 2 index selectors
//            each task: PCA axis (power iteration) -> endpoints -> quantise (+p-bit choice) ->
//            index assignment with exact integer palette error -> least-squares endpoint refit -> repeat
//   stage 3  subset errors combined with __shfl_xor, winner by integer-key warp min (ties: lowest lane)
//   stage 4  16 lanes = 16 pixels: exhaustive nearest palette entry, anchor fix-up, every lane shifts
//            its field into a 128-bit word, warp OR-reduction, one 128-bit store
// Modes tried with default flags equal the reference's (1,3,4,5,6 and 7 when alpha != 255, :2803-2821);
// BC7_QUICK keeps only mode 6 (:2811); USE_3SUBSETS is accepted and ignored (modes 0/2 are never emitted).
// Error metric = the reference's: sum of squared 8-bit differences over R,G,B,A (ComputeError :1559-1596).
#pragma once
#include "dxb_warp.cuh"
#include "dxb_pixel.cuh"
#include "dxb_bc67_tables.h"

#ifndef DXB_BC7_KSHAPES_DBG
#define DXB_BC7_KSHAPES_DBG 8       // experiment knob: evaluate only the K best-ranked shapes (8 = all lane slots distinct)
#endif
#ifndef DXB_BC7_ROUNDS
#define DXB_BC7_ROUNDS 2          // endpoint evaluation rounds per task (1 = PCA only, each extra = one LS refit)
#endif

struct dxb_bc7_res { float err; uint32_t q0, q1, pbits; };

// interpolation weight of index k at `ib` index bits: {0,21,43,64} {0,9,..,64} {0,4,..,64}  (BC6HBC7.cpp:327-329)
DXB_DEV uint32_t dxb_bc7_weight(uint32_t ib, uint32_t k)
{
    const uint32_t n = (1u << ib) - 1u;
    const uint32_t M = (ib == 2) ? 21846u : (ib == 3) ? 9363u : 4370u;      // ceil(65536 / n)
    return ((64u * k + (n >> 1)) * M) >> 16;
}

// bit-replicating unquantise of a `B`-bit field (D3DX_BC7::Unquantize, BC6HBC7.cpp:827-832)
DXB_DEV uint32_t dxb_bc7_unq(uint32_t f, uint32_t B)
{
    const uint32_t c = (f << (8u - B)) & 0xFFu;
    return c | (c >> B);
}

// stage 0: the reference's LDR conversion (BC6HBC7.cpp:2794-2797), result as float 0..255
DXB_DEV float dxb_bc7_ldr(float c)
{
    const float t = c * 255.0f;
    float u = t + 0.01f;
    u = (u < 255.0f) ? u : 255.0f;       // std::min<float>(255, u)
    u = (0.0f < u) ? u : 0.0f;           // std::max<float>(0, u)
    return (float)(dxb_f2i(u) & 0xFF);
}

DXB_DEV dxb_px dxb_bc7_rotate(dxb_px p, int rot)
{
    if (rot == 1) { const float t = p.x; p.x = p.w; p.w = t; }
    else if (rot == 2) { const float t = p.y; p.y = p.w; p.w = t; }
    else if (rot == 3) { const float t = p.z; p.z = p.w; p.w = t; }
    return p;
}

// ---------------------------------------------------------------------------------------------------
// stage 1: residual of the best line through one subset, from its moments (n*covariance form).
// est = (trace - lambda_max) + lambda_max * qf  where qf models the index quantisation along the axis.
DXB_DEV float dxb_bc7_subset_estimate(float n, const float* s, const float* m, float qf)
{
    if (n < 1.5f) return 0.0f;
    const float inv = 1.0f / n;
    const float c00 = dxb_fma(-s[0] * inv, s[0], m[0]), c01 = dxb_fma(-s[0] * inv, s[1], m[1]);
    const float c02 = dxb_fma(-s[0] * inv, s[2], m[2]), c03 = dxb_fma(-s[0] * inv, s[3], m[3]);
    const float c11 = dxb_fma(-s[1] * inv, s[1], m[4]), c12 = dxb_fma(-s[1] * inv, s[2], m[5]);
    const float c13 = dxb_fma(-s[1] * inv, s[3], m[6]), c22 = dxb_fma(-s[2] * inv, s[2], m[7]);
    const float c23 = dxb_fma(-s[2] * inv, s[3], m[8]), c33 = dxb_fma(-s[3] * inv, s[3], m[9]);
    const float tr = (c00 + c11) + (c22 + c33);
    if (!(tr > 1e-3f)) return 0.0f;
    // power iteration from the row with the largest diagonal
    float v0, v1, v2, v3;
    if (c00 >= c11 && c00 >= c22 && c00 >= c33) { v0 = c00; v1 = c01; v2 = c02; v3 = c03; }
    else if (c11 >= c22 && c11 >= c33) { v0 = c01; v1 = c11; v2 = c12; v3 = c13; }
    else if (c22 >= c33) { v0 = c02; v1 = c12; v2 = c22; v3 = c23; }
    else { v0 = c03; v1 = c13; v2 = c23; v3 = c33; }
    float lam = 0.0f;
    for (int it = 0; it < 3; ++it)
    {
        const float w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, dxb_fma(c02, v2, c03 * v3)));
        const float w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, dxb_fma(c12, v2, c13 * v3)));
        const float w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, dxb_fma(c22, v2, c23 * v3)));
        const float w3 = dxb_fma(c03, v0, dxb_fma(c13, v1, dxb_fma(c23, v2, c33 * v3)));
        const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, dxb_fma(v2, v2, v3 * v3)));
        const float vw = dxb_fma(v0, w0, dxb_fma(v1, w1, dxb_fma(v2, w2, v3 * w3)));
        lam = (vv > 0.0f) ? vw / vv : 0.0f;
        const float mx = fmaxf(fmaxf(fabsf(w0), fabsf(w1)), fmaxf(fabsf(w2), fabsf(w3)));
        if (!(mx > 0.0f)) break;
        const float r = 1.0f / mx;
        v0 = w0 * r; v1 = w1 * r; v2 = w2 * r; v3 = w3 * r;
    }
    lam = fminf(lam, tr);
    const float resid = fmaxf(tr - lam, 0.0f);
    return dxb_fma(lam, qf, resid);
}

// moments of the pixels selected by `mask`: s[4] sums, m[10] upper-triangular products, returns count
DXB_DEV float dxb_bc7_moments(const dxb_px* px, uint32_t mask, float* s, float* m)
{
    float n = 0.0f;
    for (int k = 0; k < 4; ++k) s[k] = 0.0f;
    for (int k = 0; k < 10; ++k) m[k] = 0.0f;
    for (int i = 0; i < 16; ++i)
    {
        const float f = dxb_uint_as_float((0u - ((mask >> i) & 1u)) & 0x3F800000u);
        const dxb_px p = px[i];
        const float x = p.x * f, y = p.y * f, z = p.z * f, w = p.w * f;
        n += f;
        s[0] += x; s[1] += y; s[2] += z; s[3] += w;
        m[0] = dxb_fma(x, p.x, m[0]); m[1] = dxb_fma(x, p.y, m[1]); m[2] = dxb_fma(x, p.z, m[2]); m[3] = dxb_fma(x, p.w, m[3]);
        m[4] = dxb_fma(y, p.y, m[4]); m[5] = dxb_fma(y, p.z, m[5]); m[6] = dxb_fma(y, p.w, m[6]);
        m[7] = dxb_fma(z, p.z, m[7]); m[8] = dxb_fma(z, p.w, m[8]); m[9] = dxb_fma(w, p.w, m[9]);
    }
    return n;
}

// estimate for a whole 2-subset shape; tot* = moments of all 16 pixels
DXB_DEV float dxb_bc7_shape_estimate(const dxb_px* px, uint32_t shape, float qf, const float* totS, const float* totM)
{
    const uint32_t mask1 = dxb_part2[shape];
    float s1[4], m1[10], s0[4], m0[10];
    const float n1 = dxb_bc7_moments(px, mask1, s1, m1);
    for (int k = 0; k < 4; ++k) s0[k] = totS[k] - s1[k];
    for (int k = 0; k < 10; ++k) m0[k] = totM[k] - m1[k];
    return dxb_bc7_subset_estimate(16.0f - n1, s0, m0, qf) + dxb_bc7_subset_estimate(n1, s1, m1, qf);
}

// ---------------------------------------------------------------------------------------------------
// branch-free helpers (every lane of the warp runs the same instruction stream whatever its mode)
#define DXB_MAGIC 12582912.0f                      // 1.5 * 2^23: (x + MAGIC) - MAGIC == round-to-nearest-even(x), |x| < 2^22

DXB_DEV float dxb_rne(float x) { const float t = x + DXB_MAGIC; return t - DXB_MAGIC; }
DXB_DEV float dxb_bit_as_float(uint32_t mask, int i)       // (mask >> i) & 1 as 0.0f / 1.0f without an I2F
{
    return dxb_uint_as_float((0u - ((mask >> i) & 1u)) & 0x3F800000u);
}
// interpolation weight / 64 of (float) index k at nmax = 2^ib - 1: RNE(k * 64 / nmax) / 64 reproduces the
// BC7 weight tables {0,21,43,64} {0,9,18,27,37,46,55,64} {0,4,9,...,60,64} exactly (no product is a tie)
DXB_DEV float dxb_bc7_weightf(float k, float c64 /* 64 / nmax */) { return dxb_rne(k * c64) * (1.0f / 64.0f); }

// endpoint quantisation for one channel value e (0..255 float), branch-free
//   bits : field bits without p-bit; hasP (0/1): field is followed by a p-bit; p (0/1): its value
// returns the field (without p); *deq = the 8-bit value the decoder reconstructs
DXB_DEV uint32_t dxb_bc7_quant1(float e, uint32_t bits, uint32_t hasP, uint32_t p, float* deq)
{
    const uint32_t B = bits + hasP;
    const uint32_t qmax = (1u << bits) - 1u;
    const float fmaxv = (float)((1u << B) - 1u);
    const float f = e * (fmaxv * (1.0f / 255.0f));
    // no p-bit: q = floor(f + 0.5);  p-bit: q = floor((f - p) / 2 + 0.5)
    const float half = hasP ? 0.5f : 1.0f;
    const float h = dxb_fma(f - (float)(p & hasP), half, 0.5f);
    int32_t qi = dxb_f2i(floorf(h));
    qi = qi < 0 ? 0 : (qi > (int32_t)qmax ? (int32_t)qmax : qi);
    const uint32_t q = (uint32_t)qi;
    const uint32_t full = hasP ? ((q << 1) | p) : q;
    *deq = (float)dxb_bc7_unq(full, B);
    return q;
}

struct dxb_bc7_modecfg { uint32_t cbits, abits, ptype /*0 none,1 unique,2 shared*/, ib, ib2; };

DXB_DEV dxb_bc7_modecfg dxb_bc7_cfg(int mode)
{
    // packed per mode: cbits | abits<<4 | ptype<<8 | ib<<12 | ib2<<16   (mode table BC6HBC7.cpp:1106-1124)
    const uint32_t t1 = 6u | (0u << 4) | (2u << 8) | (3u << 12) | (0u << 16);
    const uint32_t t3 = 7u | (0u << 4) | (1u << 8) | (2u << 12) | (0u << 16);
    const uint32_t t4 = 5u | (6u << 4) | (0u << 8) | (2u << 12) | (3u << 16);
    const uint32_t t5 = 7u | (8u << 4) | (0u << 8) | (2u << 12) | (2u << 16);
    const uint32_t t6 = 7u | (7u << 4) | (1u << 8) | (4u << 12) | (0u << 16);
    const uint32_t t7 = 5u | (5u << 4) | (1u << 8) | (2u << 12) | (0u << 16);
    const uint32_t t = (mode == 1) ? t1 : (mode == 3) ? t3 : (mode == 4) ? t4 : (mode == 5) ? t5 : (mode == 6) ? t6 : t7;
    dxb_bc7_modecfg c;
    c.cbits = t & 15u; c.abits = (t >> 4) & 15u; c.ptype = (t >> 8) & 15u; c.ib = (t >> 12) & 15u; c.ib2 = (t >> 16) & 15u;
    return c;
}

// Quantise both endpoints of a subset (vector part), choosing p-bits; branch-free over modes.
//   use3 = 1.0f when channel 3 is part of the vector (modes 6/7) else 0.0f (its field then stays 0)
//   pforce < 0 : choose p-bits by endpoint reconstruction error; else bit0/bit1 = forced p of endpoint 0/1
DXB_DEV void dxb_bc7_quant_endpoints(const float* E0, const float* E1, float use3, uint32_t cbits, uint32_t abits,
                                     uint32_t ptype, int pforce, uint32_t* q0, uint32_t* q1, uint32_t* pbits, float* D0, float* D1)
{
    const uint32_t hasP = (ptype != 0u) ? 1u : 0u;
    uint32_t Q0[2] = { 0, 0 }, Q1[2] = { 0, 0 };
    float d0[2][4], d1[2][4];
    float err0[2] = { 0.0f, 0.0f }, err1[2] = { 0.0f, 0.0f };
    for (int p = 0; p < 2; ++p)
    {
        for (uint32_t c = 0; c < 4; ++c)
        {
            const uint32_t bits = (c == 3) ? (abits ? abits : 1u) : cbits;
            const float wgt = (c == 3) ? use3 : 1.0f;
            float a, b;
            uint32_t f0 = dxb_bc7_quant1(E0[c], bits, hasP, (uint32_t)p, &a);
            uint32_t f1 = dxb_bc7_quant1(E1[c], bits, hasP, (uint32_t)p, &b);
            if (c == 3) { const uint32_t keep = (use3 != 0.0f) ? 0xFFu : 0u; f0 &= keep; f1 &= keep; a *= wgt; b *= wgt; }
            Q0[p] |= f0 << (8 * c); Q1[p] |= f1 << (8 * c);
            d0[p][c] = a; d1[p][c] = b;
            const float ea = (a - E0[c]) * wgt, eb = (b - E1[c]) * wgt;
            err0[p] = dxb_fma(ea, ea, err0[p]); err1[p] = dxb_fma(eb, eb, err1[p]);
        }
    }
    // heuristic choice, then overrides; all selects
    uint32_t p0 = (err0[1] < err0[0]) ? 1u : 0u;
    uint32_t p1 = (err1[1] < err1[0]) ? 1u : 0u;
    const uint32_t ps = ((err0[1] + err1[1]) < (err0[0] + err1[0])) ? 1u : 0u;
    if (ptype == 2u) { p0 = ps; p1 = ps; }
    if (pforce >= 0) { p0 = (uint32_t)pforce & 1u; p1 = (ptype == 2u) ? p0 : (((uint32_t)pforce >> 1) & 1u); }
    if (ptype == 0u) { p0 = 0u; p1 = 0u; }
    *q0 = p0 ? Q0[1] : Q0[0]; *q1 = p1 ? Q1[1] : Q1[0]; *pbits = p0 | (p1 << 1);
    for (int c = 0; c < 4; ++c) { D0[c] = p0 ? d0[1][c] : d0[0][c]; D1[c] = p1 ? d1[1][c] : d1[0][c]; }
}

// ---------------------------------------------------------------------------------------------------
// stage 2: one lane task.  px = 16 LDR pixels (floats 0..255).
//   mode 1/3/7: subset `mask` of a 2-subset shape;  mode 6: whole block, forced p-bit pair;
//   mode 4/5 : whole block, rotation `rot`, index selector `idxMode` (mode 4)
// Written so that all 32 lanes execute ONE instruction stream for the vector part (mode differences are
// data: bit counts, channel weight, p-bit type); only the separate-alpha part of modes 4/5 is a
// divergent section.  Idle lanes (mode < 0) run the same code on dummy parameters.
DXB_DEV dxb_bc7_res dxb_bc7_eval(const dxb_px* px, uint32_t mask, int mode, int rot, int idxMode, int pforce)
{
    const bool idle = (mode < 0);
    if (idle) { mode = 6; mask = 0xFFFFu; }
    const dxb_bc7_modecfg cfg = dxb_bc7_cfg(mode);
    const bool sep = (mode == 4 || mode == 5);
    const float use3 = (mode == 6 || mode == 7) ? 1.0f : 0.0f;
    const uint32_t ibc = (mode == 4 && idxMode) ? 3u : cfg.ib;           // colour index bits
    const uint32_t iba = (mode == 4) ? (idxMode ? 2u : 3u) : cfg.ib2;    // alpha index bits (modes 4/5)
    const bool r1 = (rot == 1), r2 = (rot == 2), r3 = (rot == 3);

    // rotated pixel: the colour vector is (x,y,z, use3*w); `a` is the rotated alpha slot
#define DXB_BC7_FETCH(i, X, Y, Z, Wv, A) \
    float X, Y, Z, Wv, A; \
    { const dxb_px p_ = px[i]; \
      X = r1 ? p_.w : p_.x; Y = r2 ? p_.w : p_.y; Z = r3 ? p_.w : p_.z; \
      A = r1 ? p_.x : (r2 ? p_.y : (r3 ? p_.z : p_.w)); Wv = A * use3; }

    // ---- vector part: moments
    float n = 0.0f, s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    float m00 = 0.0f, m01 = 0.0f, m02 = 0.0f, m03 = 0.0f, m11 = 0.0f, m12 = 0.0f, m13 = 0.0f, m22 = 0.0f, m23 = 0.0f, m33 = 0.0f;
    for (int i = 0; i < 16; ++i)
    {
        const float f = dxb_bit_as_float(mask, i);
        DXB_BC7_FETCH(i, X, Y, Z, Wv, A)
        (void)A;
        const float x = X * f, y = Y * f, z = Z * f, w = Wv * f;
        n += f;
        s0 += x; s1 += y; s2 += z; s3 += w;
        m00 = dxb_fma(x, X, m00); m01 = dxb_fma(x, Y, m01); m02 = dxb_fma(x, Z, m02); m03 = dxb_fma(x, Wv, m03);
        m11 = dxb_fma(y, Y, m11); m12 = dxb_fma(y, Z, m12); m13 = dxb_fma(y, Wv, m13);
        m22 = dxb_fma(z, Z, m22); m23 = dxb_fma(z, Wv, m23); m33 = dxb_fma(w, Wv, m33);
    }
    const float inv = 1.0f / fmaxf(n, 1.0f);
    const float mean[4] = { s0 * inv, s1 * inv, s2 * inv, s3 * inv };
    const float c00 = dxb_fma(-mean[0], s0, m00), c01 = dxb_fma(-mean[0], s1, m01), c02 = dxb_fma(-mean[0], s2, m02), c03 = dxb_fma(-mean[0], s3, m03);
    const float c11 = dxb_fma(-mean[1], s1, m11), c12 = dxb_fma(-mean[1], s2, m12), c13 = dxb_fma(-mean[1], s3, m13);
    const float c22 = dxb_fma(-mean[2], s2, m22), c23 = dxb_fma(-mean[2], s3, m23), c33 = dxb_fma(-mean[3], s3, m33);
    const float tr = (c00 + c11) + (c22 + c33);

    // principal axis: power iteration from the row with the largest diagonal (selects, no branches)
    float ax[4];
    {
        const bool b0 = (c00 >= c11 && c00 >= c22 && c00 >= c33);
        const bool b1 = !b0 && (c11 >= c22 && c11 >= c33);
        const bool b2 = !b0 && !b1 && (c22 >= c33);
        float v0 = b0 ? c00 : (b1 ? c01 : (b2 ? c02 : c03));
        float v1 = b0 ? c01 : (b1 ? c11 : (b2 ? c12 : c13));
        float v2 = b0 ? c02 : (b1 ? c12 : (b2 ? c22 : c23));
        float v3 = b0 ? c03 : (b1 ? c13 : (b2 ? c23 : c33));
        for (int it = 0; it < 4; ++it)
        {
            const float w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, dxb_fma(c02, v2, c03 * v3)));
            const float w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, dxb_fma(c12, v2, c13 * v3)));
            const float w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, dxb_fma(c22, v2, c23 * v3)));
            const float w3 = dxb_fma(c03, v0, dxb_fma(c13, v1, dxb_fma(c23, v2, c33 * v3)));
            const float mx = fmaxf(fmaxf(fabsf(w0), fabsf(w1)), fmaxf(fabsf(w2), fabsf(w3)));
            const float r = (mx > 1e-30f) ? 1.0f / mx : 0.0f;
            v0 = w0 * r; v1 = w1 * r; v2 = w2 * r; v3 = w3 * r;
        }
        const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, dxb_fma(v2, v2, v3 * v3)));
        const float r = (vv > 1e-30f && tr > 1e-3f) ? 1.0f / sqrtf(vv) : 0.0f;
        ax[0] = v0 * r; ax[1] = v1 * r; ax[2] = v2 * r; ax[3] = v3 * r;
    }

    // ---- projection extents -> initial endpoints
    float tmin = 3.0e38f, tmax = -3.0e38f;
    for (int i = 0; i < 16; ++i)
    {
        DXB_BC7_FETCH(i, X, Y, Z, Wv, A)
        (void)A;
        const float t = dxb_fma(X - mean[0], ax[0], dxb_fma(Y - mean[1], ax[1], dxb_fma(Z - mean[2], ax[2], (Wv - mean[3]) * ax[3])));
        const bool in = ((mask >> i) & 1u) != 0u;
        tmin = in ? fminf(tmin, t) : tmin; tmax = in ? fmaxf(tmax, t) : tmax;
    }
    if (!(tmin <= tmax)) { tmin = 0.0f; tmax = 0.0f; }          // empty subset (cannot happen for valid shapes)
    float E0[4], E1[4];
    for (int c = 0; c < 4; ++c)
    {
        E0[c] = fminf(fmaxf(dxb_fma(tmin, ax[c], mean[c]), 0.0f), 255.0f);
        E1[c] = fminf(fmaxf(dxb_fma(tmax, ax[c], mean[c]), 0.0f), 255.0f);
    }

    // ---- evaluation rounds (vector part)
    float bestErr = 3.0e38f; uint32_t bq0 = 0, bq1 = 0, bpb = 0;
    const float nmaxc = (float)((1u << ibc) - 1u);
    const float c64c = 64.0f / nmaxc;
    bool live = true;                                            // false once this lane has converged (keeps running, results ignored)
#if DXB_ON_DEVICE
    #pragma unroll
#endif
    for (int round = 0; round < DXB_BC7_ROUNDS; ++round)
    {
        const bool last = (round + 1 == DXB_BC7_ROUNDS);       // compile-time after unrolling: the refit sums vanish from the last round
        dxb_warp_sync();
        uint32_t q0, q1, pb; float D0[4], D1[4];
        dxb_bc7_quant_endpoints(E0, E1, use3, cfg.cbits, cfg.abits, cfg.ptype, pforce, &q0, &q1, &pb, D0, D1);
        const float dx = D1[0] - D0[0], dy = D1[1] - D0[1], dz = D1[2] - D0[2], dw = D1[3] - D0[3];
        const float dd = dxb_fma(dx, dx, dxb_fma(dy, dy, dxb_fma(dz, dz, dw * dw)));
        const float idd = (dd > 0.0f) ? 1.0f / dd : 0.0f;
        const float B0 = D0[0] + (1.0f / 128.0f), B1 = D0[1] + (1.0f / 128.0f), B2 = D0[2] + (1.0f / 128.0f), B3 = D0[3] + (1.0f / 128.0f);
        float err = 0.0f;
        float la = 0.0f, lb = 0.0f, lc = 0.0f;                     // sum (1-s)^2, s(1-s), s^2
        float u0 = 0, u1 = 0, u2 = 0, u3 = 0, v0 = 0, v1 = 0, v2 = 0, v3 = 0;     // sum (1-s) p, sum s p
        for (int i = 0; i < 16; ++i)
        {
            if ((mask >> i) & 1u)
            {
                DXB_BC7_FETCH(i, X, Y, Z, Wv, A)
                (void)A;
                const float t = dxb_fma(X - D0[0], dx, dxb_fma(Y - D0[1], dy, dxb_fma(Z - D0[2], dz, (Wv - D0[3]) * dw))) * idd;
                const float xk = fminf(fmaxf(t * nmaxc, 0.0f), nmaxc - 1.0f);
                const float k0 = dxb_rne(xk - 0.5f);            // lower bracket index without an XU-pipe FRND
                const float w0 = dxb_bc7_weightf(k0, c64c), w1 = dxb_bc7_weightf(k0 + 1.0f, c64c);
                const float sk = ((t - w0) > (w1 - t)) ? w1 : w0;
                // palette entry: floor(v/64 + 0.5) == RNE(v/64 + 1/128) because v/64 is a multiple of 1/64
                const float cx = dxb_rne(dxb_fma(dx, sk, B0)), cy = dxb_rne(dxb_fma(dy, sk, B1));
                const float cz = dxb_rne(dxb_fma(dz, sk, B2)), cw = dxb_rne(dxb_fma(dw, sk, B3));
                const float ex = X - cx, ey = Y - cy, ez = Z - cz, ew = Wv - cw;
                err += dxb_fma(ex, ex, dxb_fma(ey, ey, dxb_fma(ez, ez, ew * ew)));
                if (!last)
                {
                    const float os = 1.0f - sk;
                    la = dxb_fma(os, os, la); lb = dxb_fma(os, sk, lb); lc = dxb_fma(sk, sk, lc);
                    u0 = dxb_fma(os, X, u0); u1 = dxb_fma(os, Y, u1); u2 = dxb_fma(os, Z, u2); u3 = dxb_fma(os, Wv, u3);
                    v0 = dxb_fma(sk, X, v0); v1 = dxb_fma(sk, Y, v1); v2 = dxb_fma(sk, Z, v2); v3 = dxb_fma(sk, Wv, v3);
                }
            }
        }
        const bool better = live && (err < bestErr);
        bestErr = better ? err : bestErr; bq0 = better ? q0 : bq0; bq1 = better ? q1 : bq1; bpb = better ? pb : bpb;
        if (last) break;
        // least-squares refit for the next round (skipped lanes keep their endpoints)
        const float det = dxb_fma(la, lc, -(lb * lb));
        live = live && (det > 1e-4f) && (bestErr > 0.0f);
        const float id = live ? 1.0f / det : 0.0f;
        const float uu[4] = { u0, u1, u2, u3 }, vv[4] = { v0, v1, v2, v3 };
        for (int c = 0; c < 4; ++c)
        {
            const float a = dxb_fma(lc, uu[c], -(lb * vv[c])) * id;
            const float b = dxb_fma(la, vv[c], -(lb * uu[c])) * id;
            E0[c] = live ? fminf(fmaxf(a, 0.0f), 255.0f) : E0[c];
            E1[c] = live ? fminf(fmaxf(b, 0.0f), 255.0f) : E1[c];
        }
    }

    // ---- scalar part (modes 4/5): the rotated alpha slot with its own endpoints and indices
    if (sep)
    {
        float amin = 3.0e38f, amax = -3.0e38f;
        for (int i = 0; i < 16; ++i)
        {
            DXB_BC7_FETCH(i, X, Y, Z, Wv, A)
            (void)X; (void)Y; (void)Z; (void)Wv;
            amin = fminf(amin, A); amax = fmaxf(amax, A);
        }
        float A0 = amin, A1 = amax;
        float bestA = 3.0e38f; uint32_t ba0 = 0, ba1 = 0;
        const float nmaxa = (float)((1u << iba) - 1u);
        const float c64a = 64.0f / nmaxa;
        bool liveA = true;
        for (int round = 0; round < DXB_BC7_ROUNDS; ++round)
        {
            float d0, d1;
            const uint32_t f0 = dxb_bc7_quant1(A0, cfg.abits, 0u, 0u, &d0);
            const uint32_t f1 = dxb_bc7_quant1(A1, cfg.abits, 0u, 0u, &d1);
            const float da = d1 - d0;
            const float ida = (da != 0.0f) ? 1.0f / da : 0.0f;
            const float Ba = d0 + (1.0f / 128.0f);
            float err = 0.0f, la = 0.0f, lb = 0.0f, lc = 0.0f, ua = 0.0f, va = 0.0f;
            for (int i = 0; i < 16; ++i)
            {
                DXB_BC7_FETCH(i, X, Y, Z, Wv, A)
                (void)X; (void)Y; (void)Z; (void)Wv;
                const float t = (A - d0) * ida;
                const float xk = fminf(fmaxf(t * nmaxa, 0.0f), nmaxa - 1.0f);
                const float k0 = dxb_rne(xk - 0.5f);            // lower bracket index without an XU-pipe FRND
                const float w0 = dxb_bc7_weightf(k0, c64a), w1 = dxb_bc7_weightf(k0 + 1.0f, c64a);
                const float sk = ((t - w0) > (w1 - t)) ? w1 : w0;
                const float ca = dxb_rne(dxb_fma(da, sk, Ba));
                const float ea = A - ca;
                err = dxb_fma(ea, ea, err);
                const float os = 1.0f - sk;
                la = dxb_fma(os, os, la); lb = dxb_fma(os, sk, lb); lc = dxb_fma(sk, sk, lc);
                ua = dxb_fma(os, A, ua); va = dxb_fma(sk, A, va);
            }
            const bool better = liveA && (err < bestA);
            bestA = better ? err : bestA; ba0 = better ? f0 : ba0; ba1 = better ? f1 : ba1;
            const float det = dxb_fma(la, lc, -(lb * lb));
            liveA = liveA && (det > 1e-4f) && (bestA > 0.0f);
            const float id = liveA ? 1.0f / det : 0.0f;
            const float na = dxb_fma(lc, ua, -(lb * va)) * id, nb = dxb_fma(la, va, -(lb * ua)) * id;
            A0 = liveA ? fminf(fmaxf(na, 0.0f), 255.0f) : A0;
            A1 = liveA ? fminf(fmaxf(nb, 0.0f), 255.0f) : A1;
        }
        bestErr += bestA;
        bq0 = (bq0 & 0x00FFFFFFu) | (ba0 << 24);
        bq1 = (bq1 & 0x00FFFFFFu) | (ba1 << 24);
    }
#undef DXB_BC7_FETCH

    dxb_bc7_res R;
    R.err = idle ? 3.0e38f : bestErr; R.q0 = bq0; R.q1 = bq1; R.pbits = bpb;
    return R;
}

// ---------------------------------------------------------------------------------------------------
// stage 4 helpers

// dequantised 8-bit endpoint channel from its field (and p-bit if the mode has one)
DXB_DEV uint32_t dxb_bc7_deq_field(uint32_t field, uint32_t bits, uint32_t ptype, uint32_t p)
{
    return (ptype != 0) ? dxb_bc7_unq((field << 1) | p, bits + 1u) : dxb_bc7_unq(field, bits);
}

// exhaustive nearest palette entry over channels [c0, c1) ; returns index (ties -> lowest)
DXB_DEV uint32_t dxb_bc7_nearest(const int32_t* p, const int32_t* e0, const int32_t* e1, int c0, int c1, uint32_t ib)
{
    uint32_t best = 0; int32_t bestErr = 0x7fffffff;
    const uint32_t n = 1u << ib;
    for (uint32_t k = 0; k < n; ++k)
    {
        const int32_t w = (int32_t)dxb_bc7_weight(ib, k);
        int32_t err = 0;
        for (int c = c0; c < c1; ++c)
        {
            const int32_t col = (e0[c] * (64 - w) + e1[c] * w + 32) >> 6;
            const int32_t d = p[c] - col;
            err += d * d;
        }
        if (err < bestErr) { bestErr = err; best = k; }
    }
    return best;
}

// 128-bit little-endian bit field helper
struct dxb_u128 { uint64_t lo, hi; };
DXB_DEV void dxb_put_bits(dxb_u128* b, uint32_t pos, uint32_t nbits, uint32_t value)
{
    if (nbits == 0) return;
    const uint64_t v = (uint64_t)(value & ((nbits >= 32) ? 0xFFFFFFFFu : ((1u << nbits) - 1u)));
    if (pos < 64)
    {
        b->lo |= v << pos;
        if (pos + nbits > 64) b->hi |= v >> (64 - pos);
    }
    else b->hi |= v << (pos - 64);
}

// ---------------------------------------------------------------------------------------------------
// The whole-block encoder, SPMD over the 32 lanes of one warp.
//   spx   : 16 LDR pixels (floats 0..255), warp-shared (device: shared memory; emulator: plain array)
//   out   : 16 output bytes (written by lane 0)
DXB_DEV void dxb_bc7_encode_warp(const dxb_px* spx, uint32_t bcflags, uint8_t* out)
{
    const bool quick = (bcflags & DXB_BC_FLAGS_FORCE_BC7_MODE6) != 0;

    // ---- block-wide facts (every lane computes them redundantly from the shared pixels)
    bool hasAlpha = false;
    float totS[4], totM[10];
    dxb_bc7_moments(spx, 0xFFFFu, totS, totM);
    for (int i = 0; i < 16; ++i) hasAlpha = hasAlpha || (spx[i].w != 255.0f);

    // ---- stage 1: rank the 64 two-subset shapes (2 per lane), select the 8 best
    uint32_t keyA[DXB_NL], keyB[DXB_NL];
    uint32_t sel[8];
    if (!quick)
    {
        // index quantisation factor 1/(2^b-1)^2: 3-bit for mode 1 (opaque), 2-bit for mode 7 (alpha)
        const float qf = hasAlpha ? (1.0f / 9.0f) : (1.0f / 49.0f);
        DXB_LANES_BEGIN
            const float ea = dxb_bc7_shape_estimate(spx, (uint32_t)lane, qf, totS, totM);
            const float eb = dxb_bc7_shape_estimate(spx, (uint32_t)lane + 32u, qf, totS, totM);
            keyA[L] = (dxb_float_as_uint(ea) & 0xFFFFFFC0u) | (uint32_t)lane;
            keyB[L] = (dxb_float_as_uint(eb) & 0xFFFFFFC0u) | ((uint32_t)lane + 32u);
        DXB_LANES_END
        for (int r = 0; r < 8; ++r)
        {
            uint32_t cand[DXB_NL];
            DXB_LANES_BEGIN
                cand[L] = (keyA[L] < keyB[L]) ? keyA[L] : keyB[L];
            DXB_LANES_END
            const uint32_t win = dxb_warp_min_u32(cand);
            sel[r] = win & 63u;
            DXB_LANES_BEGIN
                if (keyA[L] == win) keyA[L] = 0xFFFFFFFFu;
                if (keyB[L] == win) keyB[L] = 0xFFFFFFFFu;
            DXB_LANES_END
        }
    }
    else
    {
        for (int r = 0; r < 8; ++r) sel[r] = 0;
    }

    // ---- stage 2: one task per lane
    uint32_t tMode[DXB_NL], tShape[DXB_NL], tRot[DXB_NL], tIdx[DXB_NL];
    uint32_t rErr[DXB_NL], rQ0[DXB_NL], rQ1[DXB_NL], rPb[DXB_NL];
    DXB_LANES_BEGIN
        int mode = -1, rot = 0, idxMode = 0, pforce = -1;
        uint32_t mask = 0xFFFFu, shape = 0;
        if (!hasAlpha)
        {
            if (lane < 28)
            {
                shape = sel[(lane >> 2) % DXB_BC7_KSHAPES_DBG];
                const uint32_t m1 = dxb_part2[shape];
                mask = ((lane >> 1) & 1) ? m1 : (~m1 & 0xFFFFu);
                mode = (lane & 1) ? 3 : 1;
                if (quick) mode = -1;
            }
            else { mode = 6; pforce = lane & 3; }
        }
        else
        {
            if (lane < 16)
            {
                shape = sel[(lane >> 1) % DXB_BC7_KSHAPES_DBG];
                const uint32_t m1 = dxb_part2[shape];
                mask = (lane & 1) ? m1 : (~m1 & 0xFFFFu);
                mode = quick ? -1 : 7;
            }
            else if (lane < 20) { mode = 6; pforce = lane & 3; }
            else if (lane < 24) { mode = quick ? -1 : 5; rot = lane & 3; }
            else { mode = quick ? -1 : 4; rot = lane & 3; idxMode = (lane >> 2) & 1; }
        }
        const dxb_bc7_res res = dxb_bc7_eval(spx, mask, mode, rot, idxMode, pforce);
        tMode[L] = (uint32_t)mode; tShape[L] = shape; tRot[L] = (uint32_t)rot; tIdx[L] = (uint32_t)idxMode;
        rErr[L] = (mode < 0) ? 0x03FFFFFFu : (uint32_t)dxb_f2i(fminf(res.err, 6.0e7f));
        rQ0[L] = res.q0; rQ1[L] = res.q1; rPb[L] = res.pbits;
    DXB_LANES_END

    // ---- stage 3: combine subset errors, pick the winner
    uint32_t partner[DXB_NL];
    dxb_xchg_xor_u32(rErr, partner, hasAlpha ? 1 : 2);
    uint32_t key[DXB_NL];
    DXB_LANES_BEGIN
        uint32_t e = rErr[L];
        const uint32_t md = tMode[L];
        if (md == 1u || md == 3u || md == 7u) e += partner[L];
        e = (e > 0x03FFFFFFu) ? 0x03FFFFFFu : e;
        key[L] = (e << 5) | (uint32_t)lane;
    DXB_LANES_END
    const uint32_t wkey = dxb_warp_min_u32(key);
    const int wl = (int)(wkey & 31u);
    const uint32_t wMode = dxb_bcast_u32(tMode, wl);
    const uint32_t wShape = dxb_bcast_u32(tShape, wl);
    const uint32_t wRot = dxb_bcast_u32(tRot, wl);
    const uint32_t wIdx = dxb_bcast_u32(tIdx, wl);
    const bool two = (wMode == 1u || wMode == 3u || wMode == 7u);
    const int subBit = hasAlpha ? 1 : 2;
    const int l0 = two ? (wl & ~subBit) : wl;
    const int l1 = two ? (wl | subBit) : wl;
    // endpoints of subset 0 and subset 1 (single-subset modes: both = the winner lane)
    uint32_t q0s[2], q1s[2], pbs[2];
    q0s[0] = dxb_bcast_u32(rQ0, l0); q1s[0] = dxb_bcast_u32(rQ1, l0); pbs[0] = dxb_bcast_u32(rPb, l0);
    q0s[1] = dxb_bcast_u32(rQ0, l1); q1s[1] = dxb_bcast_u32(rQ1, l1); pbs[1] = dxb_bcast_u32(rPb, l1);

    // ---- stage 4: indices, anchor fix-up, packing
    const dxb_bc7_modecfg cfg = dxb_bc7_cfg((int)wMode);
    const uint32_t ibc = (wMode == 4u && wIdx) ? 3u : cfg.ib;
    const uint32_t iba = (wMode == 4u) ? (wIdx ? 2u : 3u) : cfg.ib2;
    const uint32_t part = two ? dxb_part2[wShape] : 0u;
    const uint32_t anchor1 = two ? dxb_anchor2[wShape] : 0u;
    const bool vec4 = (wMode == 6u || wMode == 7u);

    // dequantised endpoints per subset
    int32_t e0[2][4], e1[2][4];
    for (int sb = 0; sb < 2; ++sb)
        for (uint32_t c = 0; c < 4; ++c)
        {
            const uint32_t bits = (c == 3) ? cfg.abits : cfg.cbits;
            const bool coded = (c < 3) || (cfg.abits != 0);
            const bool hasP = (cfg.ptype != 0) && !(wMode == 4u || wMode == 5u);
            e0[sb][c] = coded ? (int32_t)dxb_bc7_deq_field((q0s[sb] >> (8 * c)) & 0xFF, bits, hasP ? 1u : 0u, pbs[sb] & 1u) : 255;
            e1[sb][c] = coded ? (int32_t)dxb_bc7_deq_field((q1s[sb] >> (8 * c)) & 0xFF, bits, hasP ? 1u : 0u, (pbs[sb] >> 1) & 1u) : 255;
        }

    uint32_t idxC[DXB_NL], idxA[DXB_NL];
    DXB_LANES_BEGIN
        idxC[L] = 0; idxA[L] = 0;
        if (lane < 16)
        {
            const dxb_px pr = dxb_bc7_rotate(spx[lane], (int)wRot);
            int32_t p[4] = { dxb_f2i(pr.x), dxb_f2i(pr.y), dxb_f2i(pr.z), dxb_f2i(pr.w) };
            const int sb = (int)((part >> lane) & 1u);
            if (wMode == 4u || wMode == 5u)
            {
                idxC[L] = dxb_bc7_nearest(p, e0[sb], e1[sb], 0, 3, ibc);
                idxA[L] = dxb_bc7_nearest(p, e0[sb], e1[sb], 3, 4, iba);
            }
            else
                idxC[L] = dxb_bc7_nearest(p, e0[sb], e1[sb], 0, vec4 ? 4 : 3, ibc);
        }
    DXB_LANES_END

    // anchor fix-up: the anchor index of each subset must have its MSB clear; otherwise swap that
    // subset's endpoints and mirror its indices (weights are symmetric: w[n-k] = 64 - w[k])
    const uint32_t aC0 = dxb_bcast_u32(idxC, 0);
    const uint32_t aC1 = dxb_bcast_u32(idxC, (int)anchor1);
    const uint32_t aA0 = dxb_bcast_u32(idxA, 0);
    const bool flipC[2] = { ((aC0 >> (ibc - 1u)) & 1u) != 0, two && (((aC1 >> (ibc - 1u)) & 1u) != 0) };
    const bool flipA = (iba != 0) && (((aA0 >> (iba - 1u)) & 1u) != 0);
    DXB_LANES_BEGIN
        if (lane < 16)
        {
            const int sb = (int)((part >> lane) & 1u);
            if (flipC[sb]) idxC[L] = ((1u << ibc) - 1u) - idxC[L];
            if (flipA) idxA[L] = ((1u << iba) - 1u) - idxA[L];
        }
    DXB_LANES_END
    // endpoint fields after the swaps.  Colour channels follow flipC[subset]; in modes 4/5 the alpha
    // channel has its own index set and follows flipA.
    uint32_t f0[2], f1[2], pb0[2], pb1[2];
    for (int sb = 0; sb < 2; ++sb)
    {
        uint32_t a = q0s[sb], b = q1s[sb];
        uint32_t pa = pbs[sb] & 1u, pbv = (pbs[sb] >> 1) & 1u;
        if (wMode == 4u || wMode == 5u)
        {
            uint32_t ca = a & 0x00FFFFFFu, cb = b & 0x00FFFFFFu, aa = a >> 24, ab = b >> 24;
            if (flipC[sb]) { const uint32_t t = ca; ca = cb; cb = t; }
            if (flipA) { const uint32_t t = aa; aa = ab; ab = t; }
            a = ca | (aa << 24); b = cb | (ab << 24);
        }
        else if (flipC[sb])
        {
            const uint32_t t = a; a = b; b = t;
            const uint32_t tp = pa; pa = pbv; pbv = tp;
        }
        f0[sb] = a; f1[sb] = b; pb0[sb] = pa; pb1[sb] = pbv;
    }

    // bit layout (D3DX_BC7::Decode, BC6HBC7.cpp:2566-2780): mode (unary), partition, rotation, index
    // selector, then R of every endpoint, G, B, A, p-bits, colour indices, alpha indices.
    const uint32_t nsub = two ? 2u : 1u;
    const uint32_t partBits = two ? 6u : 0u;
    const uint32_t rotBits = (wMode == 4u || wMode == 5u) ? 2u : 0u;
    const uint32_t imBits = (wMode == 4u) ? 1u : 0u;
    const uint32_t hdr = (wMode + 1u) + partBits + rotBits + imBits;
    const uint32_t epBits = nsub * 2u * (3u * cfg.cbits + cfg.abits);
    const uint32_t npb = (cfg.ptype == 1) ? nsub * 2u : (cfg.ptype == 2) ? nsub : 0u;
    const uint32_t idxStart = hdr + epBits + npb;
    // Mode 4: the first index block is always the 2-bit set, the second the 3-bit set (:2727-2757)
    const bool swapSets = (wMode == 4u) && (wIdx != 0u);
    const uint32_t ib1 = swapSets ? iba : ibc;                 // bits of the first index block
    const uint32_t ib2v = swapSets ? ibc : iba;                // bits of the second index block
    const uint32_t firstLen = 16u * ib1 - nsub;
    const uint32_t secondStart = idxStart + firstLen;

    uint32_t w0[DXB_NL], w1[DXB_NL], w2[DXB_NL], w3[DXB_NL];
    DXB_LANES_BEGIN
        dxb_u128 bits; bits.lo = 0; bits.hi = 0;
        if (lane < 16)
        {
            const uint32_t i = (uint32_t)lane;
            const uint32_t first = swapSets ? idxA[L] : idxC[L];
            const uint32_t second = swapSets ? idxC[L] : idxA[L];
            // number of anchors strictly before pixel i in the first index block (anchors: 0 and anchor1)
            const uint32_t before = (i > 0 ? 1u : 0u) + ((two && i > anchor1) ? 1u : 0u);
            const bool isAnchor = (i == 0) || (two && i == anchor1);
            dxb_put_bits(&bits, idxStart + i * ib1 - before, isAnchor ? ib1 - 1u : ib1, first);
            if (ib2v)
                dxb_put_bits(&bits, secondStart + (i ? i * ib2v - 1u : 0u), i ? ib2v : ib2v - 1u, second);
        }
        else if (lane == 16)
        {
            uint32_t pos = 0;
            dxb_put_bits(&bits, wMode, 1, 1u); pos = wMode + 1u;
            dxb_put_bits(&bits, pos, partBits, wShape); pos += partBits;
            dxb_put_bits(&bits, pos, rotBits, wRot); pos += rotBits;
            dxb_put_bits(&bits, pos, imBits, wIdx); pos += imBits;
            for (uint32_t c = 0; c < 4; ++c)
            {
                const uint32_t nb = (c == 3) ? cfg.abits : cfg.cbits;
                if (nb == 0) continue;
                for (uint32_t sb = 0; sb < nsub; ++sb)
                {
                    dxb_put_bits(&bits, pos, nb, (f0[sb] >> (8 * c)) & 0xFF); pos += nb;
                    dxb_put_bits(&bits, pos, nb, (f1[sb] >> (8 * c)) & 0xFF); pos += nb;
                }
            }
            if (cfg.ptype == 1)
                for (uint32_t sb = 0; sb < nsub; ++sb)
                {
                    dxb_put_bits(&bits, pos, 1, pb0[sb]); pos += 1;
                    dxb_put_bits(&bits, pos, 1, pb1[sb]); pos += 1;
                }
            else if (cfg.ptype == 2)
                for (uint32_t sb = 0; sb < nsub; ++sb) { dxb_put_bits(&bits, pos, 1, pb0[sb]); pos += 1; }
        }
        w0[L] = (uint32_t)bits.lo; w1[L] = (uint32_t)(bits.lo >> 32); w2[L] = (uint32_t)bits.hi; w3[L] = (uint32_t)(bits.hi >> 32);
    DXB_LANES_END
    const uint32_t o0 = dxb_warp_or_u32(w0), o1 = dxb_warp_or_u32(w1), o2 = dxb_warp_or_u32(w2), o3 = dxb_warp_or_u32(w3);
    DXB_LANES_BEGIN
        if (lane == 0)
        {
            uint32_t* o = (uint32_t*)out;
            o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
        }
    DXB_LANES_END
}

#if !DXB_ON_DEVICE
// emulator entry: px = 16 RGBA fp32 pixels after ConvertScanline (values clamped to [0,1])
static inline void dxb_bc7_encode_block_emul(const dxb_px* px, uint32_t bcflags, uint8_t* out)
{
    dxb_px ldr[16];
    for (int i = 0; i < 16; ++i)
        ldr[i] = dxb_make_px(dxb_bc7_ldr(px[i].x), dxb_bc7_ldr(px[i].y), dxb_bc7_ldr(px[i].z), dxb_bc7_ldr(px[i].w));
    dxb_bc7_encode_warp(ldr, bcflags, out);
}
#endif
