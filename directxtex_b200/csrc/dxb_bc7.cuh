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

#ifndef DXB_BC7_ROUNDS
#define DXB_BC7_ROUNDS 3          // endpoint evaluation rounds per task (1 = PCA only, each extra = one LS refit)
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
        const float f = (float)((mask >> i) & 1u);
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
// endpoint quantisation for one channel value e (0..255 float)
//   bits  : field bits without p-bit;  hasP : field is followed by a p-bit;  p : its value
// returns the field (without p); *deq = the 8-bit value the decoder reconstructs
DXB_DEV uint32_t dxb_bc7_quant1(float e, uint32_t bits, bool hasP, uint32_t p, float* deq)
{
    const uint32_t qmax = (1u << bits) - 1u;
    uint32_t q;
    if (!hasP)
    {
        const float f = dxb_fma(e, (float)qmax * (1.0f / 255.0f), 0.5f);
        int32_t qi = dxb_f2i(f);
        qi = qi < 0 ? 0 : (qi > (int32_t)qmax ? (int32_t)qmax : qi);
        q = (uint32_t)qi;
        *deq = (float)dxb_bc7_unq(q, bits);
    }
    else
    {
        const uint32_t B = bits + 1u;
        const float fmaxv = (float)((1u << B) - 1u);
        const float f = e * (fmaxv * (1.0f / 255.0f));
        const float h = dxb_fma(f - (float)p, 0.5f, 0.5f);
        int32_t qi = dxb_f2i(floorf(h));
        qi = qi < 0 ? 0 : (qi > (int32_t)qmax ? (int32_t)qmax : qi);
        q = (uint32_t)qi;
        *deq = (float)dxb_bc7_unq((q << 1) | p, B);
    }
    return q;
}

struct dxb_bc7_modecfg { uint32_t cbits, abits, ptype /*0 none,1 unique,2 shared*/, ib, ib2; };

DXB_DEV dxb_bc7_modecfg dxb_bc7_cfg(int mode)
{
    dxb_bc7_modecfg c;
    switch (mode)
    {
    case 1: c.cbits = 6; c.abits = 0; c.ptype = 2; c.ib = 3; c.ib2 = 0; break;
    case 3: c.cbits = 7; c.abits = 0; c.ptype = 1; c.ib = 2; c.ib2 = 0; break;
    case 4: c.cbits = 5; c.abits = 6; c.ptype = 0; c.ib = 2; c.ib2 = 3; break;
    case 5: c.cbits = 7; c.abits = 8; c.ptype = 0; c.ib = 2; c.ib2 = 2; break;
    case 6: c.cbits = 7; c.abits = 7; c.ptype = 1; c.ib = 4; c.ib2 = 0; break;
    default: c.cbits = 5; c.abits = 5; c.ptype = 1; c.ib = 2; c.ib2 = 0; break;   // mode 7
    }
    return c;
}

// Quantise both endpoints of a subset (vector part, channels 0..nch-1), choosing p-bits.
//   pforce < 0 : choose p-bits by endpoint reconstruction error; else bit0/bit1 = forced p of endpoint 0/1
// outputs: q0/q1 packed fields (8 bits per channel), pbits, D0/D1 dequantised floats
DXB_DEV void dxb_bc7_quant_endpoints(const float* E0, const float* E1, uint32_t nch, uint32_t cbits, uint32_t abits,
                                     uint32_t ptype, int pforce, uint32_t* q0, uint32_t* q1, uint32_t* pbits, float* D0, float* D1)
{
    uint32_t Q0[2] = { 0, 0 }, Q1[2] = { 0, 0 };
    float d0[2][4], d1[2][4];
    float err0[2] = { 0.0f, 0.0f }, err1[2] = { 0.0f, 0.0f };
    const int np = (ptype == 0) ? 1 : 2;
    for (int p = 0; p < np; ++p)
    {
        for (uint32_t c = 0; c < 4; ++c)
        {
            if (c < nch)
            {
                const uint32_t bits = (c == 3) ? abits : cbits;
                float a, b;
                const uint32_t f0 = dxb_bc7_quant1(E0[c], bits, ptype != 0, (uint32_t)p, &a);
                const uint32_t f1 = dxb_bc7_quant1(E1[c], bits, ptype != 0, (uint32_t)p, &b);
                Q0[p] |= f0 << (8 * c); Q1[p] |= f1 << (8 * c);
                d0[p][c] = a; d1[p][c] = b;
                const float ea = a - E0[c], eb = b - E1[c];
                err0[p] = dxb_fma(ea, ea, err0[p]); err1[p] = dxb_fma(eb, eb, err1[p]);
            }
            else { d0[p][c] = 0.0f; d1[p][c] = 0.0f; }
        }
    }
    uint32_t p0 = 0, p1 = 0;
    if (ptype == 1)
    {
        if (pforce >= 0) { p0 = (uint32_t)pforce & 1u; p1 = ((uint32_t)pforce >> 1) & 1u; }
        else { p0 = (err0[1] < err0[0]) ? 1u : 0u; p1 = (err1[1] < err1[0]) ? 1u : 0u; }
    }
    else if (ptype == 2)
    {
        if (pforce >= 0) { p0 = p1 = (uint32_t)pforce & 1u; }
        else { p0 = p1 = ((err0[1] + err1[1]) < (err0[0] + err1[0])) ? 1u : 0u; }
    }
    *q0 = Q0[p0]; *q1 = Q1[p1]; *pbits = p0 | (p1 << 1);
    for (int c = 0; c < 4; ++c) { D0[c] = d0[p0][c]; D1[c] = d1[p1][c]; }
}

// ---------------------------------------------------------------------------------------------------
// stage 2: one lane task.  px = 16 LDR pixels (floats 0..255).
//   mode 1/3/7: subset `mask` of a 2-subset shape;  mode 6: whole block, forced p-bit pair;
//   mode 4/5 : whole block, rotation `rot`, index selector `idxMode` (mode 4)
DXB_DEV dxb_bc7_res dxb_bc7_eval(const dxb_px* px, uint32_t mask, int mode, int rot, int idxMode, int pforce)
{
    dxb_bc7_res R;
    R.err = 3.0e38f; R.q0 = 0; R.q1 = 0; R.pbits = 0;
    if (mode < 0) return R;

    const dxb_bc7_modecfg cfg = dxb_bc7_cfg(mode);
    const bool sep = (mode == 4 || mode == 5);
    const uint32_t nch = (mode == 6 || mode == 7) ? 4u : 3u;
    const uint32_t ibc = (mode == 4 && idxMode) ? 3u : cfg.ib;           // colour index bits
    const uint32_t iba = (mode == 4) ? (idxMode ? 2u : 3u) : cfg.ib2;    // alpha index bits (modes 4/5)
    const float wch3 = (nch == 4) ? 1.0f : 0.0f;

    // ---- vector part: moments
    float n = 0.0f, s[4] = { 0, 0, 0, 0 }, m[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 16; ++i)
    {
        const float f = (float)((mask >> i) & 1u);
        dxb_px p = dxb_bc7_rotate(px[i], rot);
        p.w *= wch3;
        const float x = p.x * f, y = p.y * f, z = p.z * f, w = p.w * f;
        n += f;
        s[0] += x; s[1] += y; s[2] += z; s[3] += w;
        m[0] = dxb_fma(x, p.x, m[0]); m[1] = dxb_fma(x, p.y, m[1]); m[2] = dxb_fma(x, p.z, m[2]); m[3] = dxb_fma(x, p.w, m[3]);
        m[4] = dxb_fma(y, p.y, m[4]); m[5] = dxb_fma(y, p.z, m[5]); m[6] = dxb_fma(y, p.w, m[6]);
        m[7] = dxb_fma(z, p.z, m[7]); m[8] = dxb_fma(z, p.w, m[8]); m[9] = dxb_fma(w, p.w, m[9]);
    }
    if (n < 0.5f) { R.err = 0.0f; return R; }
    const float inv = 1.0f / n;
    float mean[4] = { s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv };
    const float c00 = dxb_fma(-mean[0], s[0], m[0]), c01 = dxb_fma(-mean[0], s[1], m[1]), c02 = dxb_fma(-mean[0], s[2], m[2]), c03 = dxb_fma(-mean[0], s[3], m[3]);
    const float c11 = dxb_fma(-mean[1], s[1], m[4]), c12 = dxb_fma(-mean[1], s[2], m[5]), c13 = dxb_fma(-mean[1], s[3], m[6]);
    const float c22 = dxb_fma(-mean[2], s[2], m[7]), c23 = dxb_fma(-mean[2], s[3], m[8]), c33 = dxb_fma(-mean[3], s[3], m[9]);
    const float tr = (c00 + c11) + (c22 + c33);

    float ax[4] = { 0, 0, 0, 0 };
    if (tr > 1e-3f)
    {
        float v0, v1, v2, v3;
        if (c00 >= c11 && c00 >= c22 && c00 >= c33) { v0 = c00; v1 = c01; v2 = c02; v3 = c03; }
        else if (c11 >= c22 && c11 >= c33) { v0 = c01; v1 = c11; v2 = c12; v3 = c13; }
        else if (c22 >= c33) { v0 = c02; v1 = c12; v2 = c22; v3 = c23; }
        else { v0 = c03; v1 = c13; v2 = c23; v3 = c33; }
        for (int it = 0; it < 4; ++it)
        {
            const float w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, dxb_fma(c02, v2, c03 * v3)));
            const float w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, dxb_fma(c12, v2, c13 * v3)));
            const float w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, dxb_fma(c22, v2, c23 * v3)));
            const float w3 = dxb_fma(c03, v0, dxb_fma(c13, v1, dxb_fma(c23, v2, c33 * v3)));
            const float mx = fmaxf(fmaxf(fabsf(w0), fabsf(w1)), fmaxf(fabsf(w2), fabsf(w3)));
            if (!(mx > 0.0f)) break;
            const float r = 1.0f / mx;
            v0 = w0 * r; v1 = w1 * r; v2 = w2 * r; v3 = w3 * r;
        }
        const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, dxb_fma(v2, v2, v3 * v3)));
        if (vv > 0.0f)
        {
            const float r = 1.0f / sqrtf(vv);
            ax[0] = v0 * r; ax[1] = v1 * r; ax[2] = v2 * r; ax[3] = v3 * r;
        }
    }

    // ---- projection extents -> initial endpoints
    float tmin = 3.0e38f, tmax = -3.0e38f;
    for (int i = 0; i < 16; ++i)
    {
        if ((mask >> i) & 1u)
        {
            dxb_px p = dxb_bc7_rotate(px[i], rot);
            p.w *= wch3;
            const float t = dxb_fma(p.x - mean[0], ax[0], dxb_fma(p.y - mean[1], ax[1], dxb_fma(p.z - mean[2], ax[2], (p.w - mean[3]) * ax[3])));
            tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
        }
    }
    float E0[4], E1[4];
    for (int c = 0; c < 4; ++c)
    {
        E0[c] = fminf(fmaxf(dxb_fma(tmin, ax[c], mean[c]), 0.0f), 255.0f);
        E1[c] = fminf(fmaxf(dxb_fma(tmax, ax[c], mean[c]), 0.0f), 255.0f);
    }

    // ---- evaluation rounds (vector part)
    float bestErr = 3.0e38f; uint32_t bq0 = 0, bq1 = 0, bpb = 0;
    const float nmaxc = (float)((1u << ibc) - 1u);
    for (int round = 0; round < DXB_BC7_ROUNDS; ++round)
    {
        uint32_t q0, q1, pb; float D0[4], D1[4];
        dxb_bc7_quant_endpoints(E0, E1, nch, cfg.cbits, cfg.abits, cfg.ptype, pforce, &q0, &q1, &pb, D0, D1);
        const float dx = D1[0] - D0[0], dy = D1[1] - D0[1], dz = D1[2] - D0[2], dw = D1[3] - D0[3];
        const float dd = dxb_fma(dx, dx, dxb_fma(dy, dy, dxb_fma(dz, dz, dw * dw)));
        const float idd = (dd > 0.0f) ? 1.0f / dd : 0.0f;
        float err = 0.0f;
        float la = 0.0f, lb = 0.0f, lc = 0.0f;                     // sum (1-s)^2, s(1-s), s^2
        float u[4] = { 0, 0, 0, 0 }, v[4] = { 0, 0, 0, 0 };         // sum (1-s) p, sum s p
        for (int i = 0; i < 16; ++i)
        {
            if ((mask >> i) & 1u)
            {
                dxb_px p = dxb_bc7_rotate(px[i], rot);
                p.w *= wch3;
                const float t = dxb_fma(p.x - D0[0], dx, dxb_fma(p.y - D0[1], dy, dxb_fma(p.z - D0[2], dz, (p.w - D0[3]) * dw))) * idd;
                float xk = fminf(fmaxf(t * nmaxc, 0.0f), nmaxc - 1.0f);
                const uint32_t k0 = (uint32_t)dxb_f2i(xk);
                const float s0 = (float)dxb_bc7_weight(ibc, k0) * (1.0f / 64.0f);
                const float s1 = (float)dxb_bc7_weight(ibc, k0 + 1u) * (1.0f / 64.0f);
                const float sk = ((t - s0) > (s1 - t)) ? s1 : s0;
                const float cx = floorf(dxb_fma(dx, sk, D0[0]) + 0.5f), cy = floorf(dxb_fma(dy, sk, D0[1]) + 0.5f);
                const float cz = floorf(dxb_fma(dz, sk, D0[2]) + 0.5f), cw = floorf(dxb_fma(dw, sk, D0[3]) + 0.5f);
                const float ex = p.x - cx, ey = p.y - cy, ez = p.z - cz, ew = p.w - cw;
                err += dxb_fma(ex, ex, dxb_fma(ey, ey, dxb_fma(ez, ez, ew * ew)));
                const float os = 1.0f - sk;
                la = dxb_fma(os, os, la); lb = dxb_fma(os, sk, lb); lc = dxb_fma(sk, sk, lc);
                u[0] = dxb_fma(os, p.x, u[0]); u[1] = dxb_fma(os, p.y, u[1]); u[2] = dxb_fma(os, p.z, u[2]); u[3] = dxb_fma(os, p.w, u[3]);
                v[0] = dxb_fma(sk, p.x, v[0]); v[1] = dxb_fma(sk, p.y, v[1]); v[2] = dxb_fma(sk, p.z, v[2]); v[3] = dxb_fma(sk, p.w, v[3]);
            }
        }
        if (err < bestErr) { bestErr = err; bq0 = q0; bq1 = q1; bpb = pb; }
        if (round + 1 < DXB_BC7_ROUNDS)
        {
            const float det = dxb_fma(la, lc, -(lb * lb));
            if (!(det > 1e-4f) || bestErr <= 0.0f) break;
            const float id = 1.0f / det;
            for (int c = 0; c < 4; ++c)
            {
                const float a = dxb_fma(lc, u[c], -(lb * v[c])) * id;
                const float b = dxb_fma(la, v[c], -(lb * u[c])) * id;
                E0[c] = fminf(fmaxf(a, 0.0f), 255.0f);
                E1[c] = fminf(fmaxf(b, 0.0f), 255.0f);
            }
        }
    }

    // ---- scalar part (modes 4/5): the rotated alpha slot with its own endpoints and indices
    if (sep)
    {
        float amin = 3.0e38f, amax = -3.0e38f;
        for (int i = 0; i < 16; ++i)
        {
            const float a = dxb_bc7_rotate(px[i], rot).w;
            amin = fminf(amin, a); amax = fmaxf(amax, a);
        }
        float A0 = amin, A1 = amax;
        float bestA = 3.0e38f; uint32_t ba0 = 0, ba1 = 0;
        const float nmaxa = (float)((1u << iba) - 1u);
        for (int round = 0; round < DXB_BC7_ROUNDS; ++round)
        {
            float d0, d1;
            const uint32_t f0 = dxb_bc7_quant1(A0, cfg.abits, false, 0, &d0);
            const uint32_t f1 = dxb_bc7_quant1(A1, cfg.abits, false, 0, &d1);
            const float da = d1 - d0;
            const float ida = (da != 0.0f) ? 1.0f / da : 0.0f;
            float err = 0.0f, la = 0.0f, lb = 0.0f, lc = 0.0f, ua = 0.0f, va = 0.0f;
            for (int i = 0; i < 16; ++i)
            {
                const float a = dxb_bc7_rotate(px[i], rot).w;
                const float t = (a - d0) * ida;
                float xk = fminf(fmaxf(t * nmaxa, 0.0f), nmaxa - 1.0f);
                const uint32_t k0 = (uint32_t)dxb_f2i(xk);
                const float s0 = (float)dxb_bc7_weight(iba, k0) * (1.0f / 64.0f);
                const float s1 = (float)dxb_bc7_weight(iba, k0 + 1u) * (1.0f / 64.0f);
                const float sk = ((t - s0) > (s1 - t)) ? s1 : s0;
                const float ca = floorf(dxb_fma(da, sk, d0) + 0.5f);
                const float ea = a - ca;
                err = dxb_fma(ea, ea, err);
                const float os = 1.0f - sk;
                la = dxb_fma(os, os, la); lb = dxb_fma(os, sk, lb); lc = dxb_fma(sk, sk, lc);
                ua = dxb_fma(os, a, ua); va = dxb_fma(sk, a, va);
            }
            if (err < bestA) { bestA = err; ba0 = f0; ba1 = f1; }
            if (round + 1 < DXB_BC7_ROUNDS)
            {
                const float det = dxb_fma(la, lc, -(lb * lb));
                if (!(det > 1e-4f) || bestA <= 0.0f) break;
                const float id = 1.0f / det;
                A0 = fminf(fmaxf(dxb_fma(lc, ua, -(lb * va)) * id, 0.0f), 255.0f);
                A1 = fminf(fmaxf(dxb_fma(la, va, -(lb * ua)) * id, 0.0f), 255.0f);
            }
        }
        bestErr += bestA;
        bq0 = (bq0 & 0x00FFFFFFu) | (ba0 << 24);
        bq1 = (bq1 & 0x00FFFFFFu) | (ba1 << 24);
    }

    R.err = bestErr; R.q0 = bq0; R.q1 = bq1; R.pbits = bpb;
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
                shape = sel[lane >> 2];
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
                shape = sel[lane >> 1];
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
