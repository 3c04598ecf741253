// dxb_bc7.cuh — BC7 block encoder, ONE HALF-WARP PER 4x4 BLOCK, two blocks per warp (single-source SPMD, see dxb_warp.cuh).
//
// What it replaces: D3DXEncodeBC7 -> D3DX_BC7::Encode (BC6HBC7.cpp:3654-3659, 2783-2889).
// Parity contract (north_star / SURVEY.md 8(d)): NOT bit-exact; the decoded result must be a valid BC7
// stream for the reference decoder (D3DX_BC7::Decode, BC6HBC7.cpp:2566-2780) and its RGBA MSE against the
// source must stay within the tolerance stated in DESIGN.md of the MSE the reference CPU encoder
// reaches on the same input.  The reference's search (Newton fit + rank 64 shapes + refine 16 with
// +-5 exhaustive perturbation, ~7 ms/block on one CPU core) is replaced by a search shaped for the machine
// (per block = per 16-lane half of a warp):
//
//   stage 0  LDR pixels exactly as the reference quantises them: uint8(clamp(c*255 + 0.01))   (:2792-2799)
//   stage 1  the moments of every two-subset shape as ONE exact matrix product on the tensor cores
//            (dxb_bc7_build_moments), 4 shapes per lane ranked by a closed-form line-fit residual, the 3 best
//            kept by an integer-key half-warp min
//   stage 2  16 lane tasks evaluated concurrently, one (mode, shape, subset | rotation | p-bits) each:
//              opaque block : 3 best shapes x 2 subsets x {mode 1, mode 3}  +  mode 6 x 4 p-bit pairs
//              alpha block  : 3 best shapes x 2 subsets x mode 7, mode 6 x 4 p-bit pairs,
//                             mode 5 x 4 rotations, mode 4 x 2 index selectors
//            each task: covariance from the moment table -> PCA axis (power iteration) -> endpoints ->
//            float-only quantisation (+p-bit choice) -> index assignment -> least-squares endpoint refit -> repeat
//   stage 3  subset errors combined with __shfl_xor, winner by integer-key half-warp min (ties: lowest lane)
//   stage 4  16 lanes = 16 pixels: exhaustive nearest palette entry against the exact integer palette, anchor
//            fix-up, every lane shifts its index fields and one endpoint field into a 128-bit word, half-warp
//            OR-reduction, one 128-bit store per block
// Modes tried with default flags equal the reference's (1,3,4,5,6 and 7 when alpha != 255, :2803-2821);
// BC7_QUICK keeps only mode 6 (:2811); with USE_3SUBSETS a second pass of lane tasks tries the three-subset modes 0 and 2 (:2807).
// Error metric = the reference's: sum of squared 8-bit differences over R,G,B,A (ComputeError :1559-1596).
#pragma once
#include "dxb_warp.cuh"
#include "dxb_pixel.cuh"
#include "dxb_bc67_tables.h"

#ifndef DXB_BC7_PIXUNROLL
#define DXB_BC7_PIXUNROLL 2       // unroll factor of the 16-pixel loops of a lane task (code size vs loop overhead)
#endif
static constexpr int dxb_bc7_pixunroll = DXB_BC7_PIXUNROLL;
#ifndef DXB_BC7_PCA_ITERS
#define DXB_BC7_PCA_ITERS 2       // power-iteration steps for a task's principal axis
#endif
#ifndef DXB_BC7_EST_ITERS
#define DXB_BC7_EST_ITERS 2       // power-iteration steps inside the stage-1 shape estimate
#endif
#ifndef DXB_BC7_ROUNDS
#define DXB_BC7_ROUNDS 2          // endpoint evaluation rounds per task (1 = PCA only, each extra = one LS refit)
#endif


// packed-fp32 regions (R<n>_*): -DDXB_SCALAR_REGION=<n> issues region n as scalar instructions (bisecting tool)
#ifndef DXB_SCALAR_REGION
#define DXB_SCALAR_REGION 0
#endif
#define DXB_RDEF(N) \
    DXB_DEV dxb_f2 R##N##_fma2(dxb_f2 a, dxb_f2 b, dxb_f2 c) { return (DXB_SCALAR_REGION == N) ? dxb_fma2s(a, b, c) : dxb_fma2(a, b, c); } \
    DXB_DEV dxb_f2 R##N##_add2(dxb_f2 a, dxb_f2 b) { return (DXB_SCALAR_REGION == N) ? dxb_add2s(a, b) : dxb_add2(a, b); } \
    DXB_DEV dxb_f2 R##N##_mul2(dxb_f2 a, dxb_f2 b) { return (DXB_SCALAR_REGION == N) ? dxb_mul2s(a, b) : dxb_mul2(a, b); } \
    DXB_DEV dxb_f2 R##N##_sub2(dxb_f2 a, dxb_f2 b) { return (DXB_SCALAR_REGION == N) ? dxb_sub2s(a, b) : dxb_sub2(a, b); }
DXB_RDEF(1) DXB_RDEF(2) DXB_RDEF(3) DXB_RDEF(4) DXB_RDEF(5) DXB_RDEF(6)

struct dxb_bc7_res { float err; uint32_t q0, q1, pbits; };

#define DXB_MAGIC 12582912.0f                      // 1.5 * 2^23: (x + MAGIC) - MAGIC == round-to-nearest-even(x), |x| < 2^22
DXB_DEV float dxb_rne(float x) { const float t = x + DXB_MAGIC; return t - DXB_MAGIC; }

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
// stage 1: residual of the best line through one subset, from its moments v[14] (4 sums, 10 products).
// est = (trace - lambda_max) + lambda_max * qf  where qf models the index quantisation along the axis.
// lambda_max: DXB_BC7_EST_ITERS un-normalised power-iteration steps, Rayleigh quotient at the end.
// n = pixel count of the subset (0..16).
DXB_TABLE float dxb_rcp16[17] = { 1.0f, 1.0f, 1.0f / 2.0f, 1.0f / 3.0f, 1.0f / 4.0f, 1.0f / 5.0f, 1.0f / 6.0f, 1.0f / 7.0f, 1.0f / 8.0f,
                                  1.0f / 9.0f, 1.0f / 10.0f, 1.0f / 11.0f, 1.0f / 12.0f, 1.0f / 13.0f, 1.0f / 14.0f, 1.0f / 15.0f, 1.0f / 16.0f };

DXB_DEV float dxb_bc7_subset_estimate(uint32_t n, const float* v, float qf)
{
    const float inv = dxb_rcp16[n];
    const float* s = v; const float* m = v + 4;
    const float c00 = dxb_fma(-s[0] * inv, s[0], m[0]), c01 = dxb_fma(-s[0] * inv, s[1], m[1]);
    const float c02 = dxb_fma(-s[0] * inv, s[2], m[2]), c03 = dxb_fma(-s[0] * inv, s[3], m[3]);
    const float c11 = dxb_fma(-s[1] * inv, s[1], m[4]), c12 = dxb_fma(-s[1] * inv, s[2], m[5]);
    const float c13 = dxb_fma(-s[1] * inv, s[3], m[6]), c22 = dxb_fma(-s[2] * inv, s[2], m[7]);
    const float c23 = dxb_fma(-s[2] * inv, s[3], m[8]), c33 = dxb_fma(-s[3] * inv, s[3], m[9]);
    const float tr = (c00 + c11) + (c22 + c33);
    const bool flat = !(tr > 1e-3f) || (n < 2u);
    // start from the row with the largest diagonal.  Entries are at most 16 * 255^2 * 4 = 4.2e6 (and 1.6e10 for the
    // half-float domain of BC6H), so two un-normalised steps stay inside fp32: |v| <= c^2 * 4, |w| <= c^3 * 16,
    // v.w <= 4.5e32 (BC6H: scaled by the caller's centring, see dxb_bc6h.cuh)
    const bool b0 = (c00 >= c11 && c00 >= c22 && c00 >= c33);
    const bool b1 = !b0 && (c11 >= c22 && c11 >= c33);
    const bool b2 = !b0 && !b1 && (c22 >= c33);
    float v0 = b0 ? c00 : (b1 ? c01 : (b2 ? c02 : c03));
    float v1 = b0 ? c01 : (b1 ? c11 : (b2 ? c12 : c13));
    float v2 = b0 ? c02 : (b1 ? c12 : (b2 ? c22 : c23));
    float v3 = b0 ? c03 : (b1 ? c13 : (b2 ? c23 : c33));
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
    for (int it = 0; it < DXB_BC7_EST_ITERS; ++it)
    {
        w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, dxb_fma(c02, v2, c03 * v3)));
        w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, dxb_fma(c12, v2, c13 * v3)));
        w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, dxb_fma(c22, v2, c23 * v3)));
        w3 = dxb_fma(c03, v0, dxb_fma(c13, v1, dxb_fma(c23, v2, c33 * v3)));
        if (it < DXB_BC7_EST_ITERS - 1) { v0 = w0; v1 = w1; v2 = w2; v3 = w3; }
    }
    const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, dxb_fma(v2, v2, v3 * v3)));
    const float vw = dxb_fma(v0, w0, dxb_fma(v1, w1, dxb_fma(v2, w2, v3 * w3)));
    const float lam = (vv > 0.0f) ? fminf(vw / vv, tr) : 0.0f;
    // (tr - lam) + lam * qf
    const float e = dxb_fma(lam, qf, fmaxf(tr - lam, 0.0f));
    return flat ? 0.0f : e;
}

// ---- stage 1 moment table -------------------------------------------------------------------------
// For every two-subset shape s the 14 moments of subset 1 (sum of x,y,z,w and of the 10 products
// xx,xy,xz,xw,yy,yz,yw,zz,zw,ww over the pixels whose bit is set in dxb_part2[s]) are one matrix product
//     M[65 x 14] = S[65 x 16] * F[16 x 14],   S = 0/1 membership (row 64 = all ones -> whole-block totals).
// LDR pixel values are integers 0..255, so every entry is an integer < 2^24: exact in fp32 in any order.
// On sm_100a the product runs on the tensor cores (mma.sync m16n8k16, bf16 in / fp32 out): products up to
// 255^2 are split into two 8-bit halves (hi*256 + lo), each exactly representable in bf16, giving 24
// feature columns = three n-tiles; 5 m-tiles x 3 n-tiles = 15 MMAs per block instead of ~1100 FFMA per lane.
// The host emulator computes the same integers with plain loops, so device and emulator agree bit for bit.
#define DXB_BC7_MT_ROWS 65
#define DXB_BC7_MT_FLOATS (DXB_BC7_MT_ROWS * 16)

struct dxb_bc7_scratch
{
    dxb_px   px[32];                          // LDR pixels (floats 0..255) of the warp's two blocks: half h -> px[16h ..]
    uint32_t pq[32];                          // the same pixels packed as bytes R | G << 8 | B << 16 | A << 24
    float    mt[2][DXB_BC7_MT_FLOATS];        // moment tables: row = shape, 16-float rows, 16-byte chunks XOR-swizzled
    // (device: the bf16 feature matrix F^T, uint16_t[2][24][16], lives in the first 1536 bytes of mt until the
    //  MMA B fragments have been read into registers)
};

// float offset of 16-byte chunk c (0..3) of table row `row`; the swizzle makes both the MMA-fragment
// stores and the row-per-lane loads bank-conflict free
DXB_DEV int dxb_bc7_mt_chunk(int row, int c) { return row * 16 + (((c ^ (row >> 1)) & 3) << 2); }

// columns: 0..3 = sums, 4..13 = products.  v[14] <- row
DXB_DEV void dxb_bc7_mt_load(const float* mt, int row, float* v)
{
#if DXB_ON_DEVICE
    const float4 a = *(const float4*)(mt + dxb_bc7_mt_chunk(row, 0));
    const float4 b = *(const float4*)(mt + dxb_bc7_mt_chunk(row, 1));
    const float4 c = *(const float4*)(mt + dxb_bc7_mt_chunk(row, 2));
    const float2 d = *(const float2*)(mt + dxb_bc7_mt_chunk(row, 3));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    v[8] = c.x; v[9] = c.y; v[10] = c.z; v[11] = c.w; v[12] = d.x; v[13] = d.y;
#else
    for (int k = 0; k < 14; ++k) v[k] = mt[dxb_bc7_mt_chunk(row, k >> 2) + (k & 3)];
#endif
}

#if DXB_ON_DEVICE
DXB_DEV uint16_t dxb_bc7_bf16_of_byte(uint32_t n)          // bf16 bits of the integer n (0..255), no I2F
{
    const float f = __uint_as_float(0x4B000000u | n) - 8388608.0f;
    return (uint16_t)(__float_as_uint(f) >> 16);
}
#endif

// Fills S->mt[0..1] from S->px (both blocks of the warp).  Collective over the warp.
DXB_DEV uint32_t dxb_bc7_pack_px(const dxb_px p)
{
    return (uint32_t)dxb_f2i(p.x) | ((uint32_t)dxb_f2i(p.y) << 8) | ((uint32_t)dxb_f2i(p.z) << 16) | ((uint32_t)dxb_f2i(p.w) << 24);
}

DXB_DEV void dxb_bc7_build_moments(dxb_bc7_scratch* S)
{
#if DXB_ON_DEVICE
    const uint32_t lane = threadIdx.x & 31u, h = lane >> 4, hl = lane & 15u;
    {
        const dxb_px p = S->px[lane];
        S->pq[lane] = dxb_bc7_pack_px(p);
        uint16_t* F = (uint16_t*)S->mt + (h * 24 * 16 + hl);   // feature n of this pixel = F[16 * n]
        F[0] = (uint16_t)(__float_as_uint(p.x) >> 16); F[16] = (uint16_t)(__float_as_uint(p.y) >> 16);
        F[32] = (uint16_t)(__float_as_uint(p.z) >> 16); F[48] = (uint16_t)(__float_as_uint(p.w) >> 16);
        const float P[10] = { p.x * p.x, p.x * p.y, p.x * p.z, p.x * p.w, p.y * p.y, p.y * p.z, p.y * p.w, p.z * p.z, p.z * p.w, p.w * p.w };
        #pragma unroll
        for (int k = 0; k < 10; ++k)
        {
            const uint32_t b = __float_as_uint(P[k] + 8388608.0f);            // low 23 bits = the integer product
            const uint16_t hi = dxb_bc7_bf16_of_byte((b >> 8) & 0xFFu), lo = dxb_bc7_bf16_of_byte(b & 0xFFu);
            // n-tile 0 = {s0,s1,s2,s3, hi8,lo8, hi9,lo9}; n-tile 1 = hi0..7; n-tile 2 = lo0..7
            const int nh = (k < 8) ? 8 + k : 4 + 2 * (k - 8), nl = (k < 8) ? 16 + k : 5 + 2 * (k - 8);
            F[16 * nh] = hi; F[16 * nl] = lo;
        }
    }
    __syncwarp();
    const uint32_t g = lane >> 2, q = lane & 3u;
    uint32_t B[2][3][2];
    #pragma unroll
    for (int hb = 0; hb < 2; ++hb)
        #pragma unroll
        for (int t = 0; t < 3; ++t)
        {
            const uint32_t* w = (const uint32_t*)((const uint16_t*)S->mt + ((hb * 24 + t * 8 + (int)g) * 16));
            B[hb][t][0] = w[q]; B[hb][t][1] = w[q + 4];
        }
    __syncwarp();                                             // the feature bytes are dead from here: mt may be written
    #pragma unroll
    for (int tile = 0; tile < 5; ++tile)
    {
        const uint4 A = ((const uint4*)dxb_bc7_afrag)[tile * 32 + lane];
        #pragma unroll
        for (int hb = 0; hb < 2; ++hb)
        {
            float d[3][4];
            #pragma unroll
            for (int t = 0; t < 3; ++t)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                             : "=f"(d[t][0]), "=f"(d[t][1]), "=f"(d[t][2]), "=f"(d[t][3])
                             : "r"(A.x), "r"(A.y), "r"(A.z), "r"(A.w), "r"(B[hb][t][0]), "r"(B[hb][t][1]), "f"(0.0f));
            float* mt = S->mt[hb];
            #pragma unroll
            for (int rr = 0; rr < 2; ++rr)                // fragment rows g and g + 8
            {
                if (tile == 4 && (rr == 1 || g != 0)) continue;
                const int row = tile * 16 + (int)g + 8 * rr;
                // products 2q, 2q+1 (columns 4+2q, 5+2q)
                const float2 pp = make_float2(fmaf(d[1][2 * rr], 256.0f, d[2][2 * rr]), fmaf(d[1][2 * rr + 1], 256.0f, d[2][2 * rr + 1]));
                *(float2*)(mt + dxb_bc7_mt_chunk(row, 1 + (int)(q >> 1)) + 2 * (int)(q & 1u)) = pp;
                if (q < 2) *(float2*)(mt + dxb_bc7_mt_chunk(row, 0) + 2 * (int)q) = make_float2(d[0][2 * rr], d[0][2 * rr + 1]);
                else mt[dxb_bc7_mt_chunk(row, 3) + (int)q - 2] = fmaf(d[0][2 * rr], 256.0f, d[0][2 * rr + 1]);
            }
        }
    }
    __syncwarp();
#else
    for (int i = 0; i < 32; ++i) S->pq[i] = dxb_bc7_pack_px(S->px[i]);
    for (int hb = 0; hb < 2; ++hb)
        for (int row = 0; row < DXB_BC7_MT_ROWS; ++row)
        {
            const uint32_t mask = (row < 64) ? dxb_part2[row] : 0xFFFFu;
            float v[14];
            for (int k = 0; k < 14; ++k) v[k] = 0.0f;
            for (int i = 0; i < 16; ++i)
                if ((mask >> i) & 1u)
                {
                    const dxb_px p = S->px[16 * hb + i];
                    v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w;
                    v[4] += p.x * p.x; v[5] += p.x * p.y; v[6] += p.x * p.z; v[7] += p.x * p.w; v[8] += p.y * p.y;
                    v[9] += p.y * p.z; v[10] += p.y * p.w; v[11] += p.z * p.z; v[12] += p.z * p.w; v[13] += p.w * p.w;
                }
            for (int k = 0; k < 14; ++k) S->mt[hb][dxb_bc7_mt_chunk(row, k >> 2) + (k & 3)] = v[k];
        }
#endif
}

// three-channel variant for opaque blocks (alpha is the constant 255: its covariance row is zero up to rounding):
// 6 covariance entries and 3x3 power-iteration steps instead of 10 and 4x4.  v = the same 14 moments.
DXB_DEV float dxb_bc7_subset_estimate3(uint32_t n, const float* v, float qf)
{
    const float inv = dxb_rcp16[n];
    const float c00 = dxb_fma(-v[0] * inv, v[0], v[4]), c01 = dxb_fma(-v[0] * inv, v[1], v[5]), c02 = dxb_fma(-v[0] * inv, v[2], v[6]);
    const float c11 = dxb_fma(-v[1] * inv, v[1], v[8]), c12 = dxb_fma(-v[1] * inv, v[2], v[9]), c22 = dxb_fma(-v[2] * inv, v[2], v[11]);
    const float tr = (c00 + c11) + c22;
    const bool flat = !(tr > 1e-3f) || (n < 2u);
    const bool b0 = (c00 >= c11 && c00 >= c22);
    const bool b1 = !b0 && (c11 >= c22);
    float v0 = b0 ? c00 : (b1 ? c01 : c02);
    float v1 = b0 ? c01 : (b1 ? c11 : c12);
    float v2 = b0 ? c02 : (b1 ? c12 : c22);
    float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f;
    for (int it = 0; it < DXB_BC7_EST_ITERS; ++it)
    {
        w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, c02 * v2));
        w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, c12 * v2));
        w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, c22 * v2));
        if (it < DXB_BC7_EST_ITERS - 1) { v0 = w0; v1 = w1; v2 = w2; }
    }
    const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, v2 * v2));
    const float vw = dxb_fma(v0, w0, dxb_fma(v1, w1, v2 * w2));
    const float lam = (vv > 0.0f) ? fminf(vw / vv, tr) : 0.0f;
    const float e = dxb_fma(lam, qf, fmaxf(tr - lam, 0.0f));
    return flat ? 0.0f : e;
}

// estimate for a whole 2-subset shape from the moment table; tot = row 64
DXB_DEV float dxb_bc7_shape_estimate(const float* mt, uint32_t shape, float qf, const float* tot, bool opaque)
{
    float v1[14], v0[14];
    dxb_bc7_mt_load(mt, (int)shape, v1);
    for (int k = 0; k < 14; ++k) v0[k] = tot[k] - v1[k];
    const uint32_t n1 = dxb_popc16(dxb_part2[shape]);
#ifndef DXB_BC7_NO_EST3
    if (opaque) return dxb_bc7_subset_estimate3(16u - n1, v0, qf) + dxb_bc7_subset_estimate3(n1, v1, qf);
#endif
    return dxb_bc7_subset_estimate(16u - n1, v0, qf) + dxb_bc7_subset_estimate(n1, v1, qf);
}


// ---------------------------------------------------------------------------------------------------
// stage 1 shape estimate "h1": what the block would cost with this two-subset shape when every subset is coded as
// nl + 1 evenly spaced points between the extreme projections on its principal axis (the un-quantised endpoints a
// task of stage 2 starts from).  It follows the reference's ranking pass (RoughMSE of every shape, BC6HBC7.cpp:3045-3110)
// in spirit: shapes are ranked by an actual index-quantisation error, not by a model of it -- the closed-form
// line-fit residual (dxb_bc7_subset_estimate) ranks collinear content (text, two-colour edges with blends) badly,
// because there every shape has residual zero and only the position of the points along the line matters.
// Exact-integer formulation (device == host emulator bit for bit, whatever the evaluation order of the dot products):
// the axis is quantised to 8-bit integers, pixels are bytes, so a projection is one u8 x s8 dot product (dp4a).
DXB_DEV int32_t dxb_dp4a_u8s8(uint32_t pix, uint32_t axis, int32_t acc)
{
#if DXB_ON_DEVICE
    int32_t d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(axis), "r"(acc));
    return d;
#else
    int32_t d = acc;
    for (int c = 0; c < 4; ++c) d += (int32_t)((pix >> (8 * c)) & 0xFFu) * (int32_t)(int8_t)((axis >> (8 * c)) & 0xFFu);
    return d;
#endif
}
DXB_DEV int32_t dxb_min3_s32(int32_t a, int32_t b, int32_t c)
{
#if DXB_ON_DEVICE
    return __vimin3_s32(a, b, c);
#else
    const int32_t m = (a < b) ? a : b; return (m < c) ? m : c;
#endif
}
DXB_DEV int32_t dxb_max3_s32(int32_t a, int32_t b, int32_t c)
{
#if DXB_ON_DEVICE
    return __vimax3_s32(a, b, c);
#else
    const int32_t m = (a > b) ? a : b; return (m > c) ? m : c;
#endif
}

struct dxb_bc7_axis { uint32_t packed; float resid, inv_aa; };   // s8x4 axis, off-axis residual tr - a'Ca/|a|^2, 1/|a|^2

// axes of BOTH subsets of a shape from their moments (V[k] = (subset 0, subset 1) as packed pairs; n0 / n1 pixels): per subset
// the covariance row with the largest diagonal (= one power-iteration step from that unit vector; more steps do not change the
// ranking measurably), scaled by a power of two to integers of magnitude <= 64.  The arithmetic of the two subsets runs as
// packed fp32 pairs; the selects are per subset.
DXB_DEV void dxb_bc7_subset_axes(uint32_t n0, uint32_t n1, const dxb_f2* V, bool opaque, dxb_bc7_axis* A0, dxb_bc7_axis* A1)
{
    const dxb_f2 inv = dxb_mk2(dxb_rcp16[n0], dxb_rcp16[n1]);
    const dxb_f2 m0 = R1_mul2(dxb_mk2(-V[0].x, -V[0].y), inv), m1 = R1_mul2(dxb_mk2(-V[1].x, -V[1].y), inv);
    const dxb_f2 m2 = R1_mul2(dxb_mk2(-V[2].x, -V[2].y), inv), m3 = R1_mul2(dxb_mk2(-V[3].x, -V[3].y), inv);
    const dxb_f2 c00 = R1_fma2(m0, V[0], V[4]), c01 = R1_fma2(m0, V[1], V[5]), c02 = R1_fma2(m0, V[2], V[6]);
    const dxb_f2 c11 = R1_fma2(m1, V[1], V[8]), c12 = R1_fma2(m1, V[2], V[9]), c22 = R1_fma2(m2, V[2], V[11]);
    // opaque blocks: alpha is the constant 255, its covariance row is zero up to rounding: forced to zero (branch-free)
    const dxb_f2 z = dxb_bc2(opaque ? 0.0f : 1.0f);
    const dxb_f2 c03 = R1_mul2(z, R1_fma2(m0, V[3], V[7])), c13 = R1_mul2(z, R1_fma2(m1, V[3], V[10]));
    const dxb_f2 c23 = R1_mul2(z, R1_fma2(m2, V[3], V[12])), c33 = R1_mul2(z, R1_fma2(m3, V[3], V[13]));
    const dxb_f2 tr = R1_add2(R1_add2(c00, c11), R1_add2(c22, c33));
    dxb_f2 v0, v1, v2, v3, sc;
    {
        // per subset: the row with the largest diagonal; its diagonal entry is the largest component (|c_ij| <= max(c_ii, c_jj)):
        // scale it into [32, 64)
#define DXB_AX_ROW(H, N) { \
        const bool flat = !(tr.H > 1e-3f) || ((N) < 2u); \
        const bool b0 = (c00.H >= c11.H && c00.H >= c22.H && c00.H >= c33.H); \
        const bool b1 = !b0 && (c11.H >= c22.H && c11.H >= c33.H); \
        const bool b2 = !b0 && !b1 && (c22.H >= c33.H); \
        v0.H = b0 ? c00.H : (b1 ? c01.H : (b2 ? c02.H : c03.H)); \
        v1.H = b0 ? c01.H : (b1 ? c11.H : (b2 ? c12.H : c13.H)); \
        v2.H = b0 ? c02.H : (b1 ? c12.H : (b2 ? c22.H : c23.H)); \
        v3.H = b0 ? c03.H : (b1 ? c13.H : (b2 ? c23.H : c33.H)); \
        const float mx = b0 ? c00.H : (b1 ? c11.H : (b2 ? c22.H : c33.H)); \
        const uint32_t E = dxb_float_as_uint(mx) >> 23;                      /* biased exponent (mx > 0 unless flat) */ \
        sc.H = flat ? 0.0f : dxb_uint_as_float((259u - E) << 23); }         /* 2^(5 - (E - 127)) */
        DXB_AX_ROW(x, n0)
        DXB_AX_ROW(y, n1)
#undef DXB_AX_ROW
    }
    // round to integers with the magic constant: the sum's low mantissa byte is the two's complement byte of the integer
    const dxb_f2 MG = dxb_bc2(DXB_MAGIC), nMG = dxb_bc2(-DXB_MAGIC);
    const dxb_f2 t0 = R1_fma2(v0, sc, MG), t1 = R1_fma2(v1, sc, MG), t2 = R1_fma2(v2, sc, MG), t3 = R1_fma2(v3, sc, MG);
    const dxb_f2 a0 = R1_add2(t0, nMG), a1 = R1_add2(t1, nMG), a2 = R1_add2(t2, nMG), a3 = R1_add2(t3, nMG);
    A0->packed = (dxb_float_as_uint(t0.x) & 0xFFu) | ((dxb_float_as_uint(t1.x) & 0xFFu) << 8) | ((dxb_float_as_uint(t2.x) & 0xFFu) << 16) | (dxb_float_as_uint(t3.x) << 24);
    A1->packed = (dxb_float_as_uint(t0.y) & 0xFFu) | ((dxb_float_as_uint(t1.y) & 0xFFu) << 8) | ((dxb_float_as_uint(t2.y) & 0xFFu) << 16) | (dxb_float_as_uint(t3.y) << 24);
    const dxb_f2 aa = R1_fma2(a0, a0, R1_fma2(a1, a1, R1_fma2(a2, a2, R1_mul2(a3, a3))));
    const dxb_f2 q0 = R1_fma2(c00, a0, R1_fma2(c01, a1, R1_fma2(c02, a2, R1_mul2(c03, a3))));
    const dxb_f2 q1 = R1_fma2(c01, a0, R1_fma2(c11, a1, R1_fma2(c12, a2, R1_mul2(c13, a3))));
    const dxb_f2 q2 = R1_fma2(c02, a0, R1_fma2(c12, a1, R1_fma2(c22, a2, R1_mul2(c23, a3))));
    const dxb_f2 q3 = R1_fma2(c03, a0, R1_fma2(c13, a1, R1_fma2(c23, a2, R1_mul2(c33, a3))));
    const dxb_f2 aCa = R1_fma2(a0, q0, R1_fma2(a1, q1, R1_fma2(a2, q2, R1_mul2(a3, q3))));
    A0->inv_aa = (aa.x > 0.0f) ? 1.0f / aa.x : 0.0f;
    A1->inv_aa = (aa.y > 0.0f) ? 1.0f / aa.y : 0.0f;
    const dxb_f2 rs = R1_fma2(dxb_mk2(-aCa.x, -aCa.y), dxb_mk2(A0->inv_aa, A1->inv_aa), tr);
    A0->resid = (sc.x == 0.0f) ? 0.0f : fmaxf(rs.x, 0.0f);
    A1->resid = (sc.y == 0.0f) ? 0.0f : fmaxf(rs.y, 0.0f);
}

#define DXB_BC7_H1_OFF (1 << 20)          // separates the two subsets' projections (|t| <= 4 * 255 * 64 < 2^17)

// pq = the block's 16 LDR pixels packed as bytes (R | G << 8 | B << 16 | A << 24).
// opaque blocks: the better of 3-bit indices (mode 1) and 2-bit indices (mode 3); alpha blocks: 2-bit (mode 7).
DXB_DEV float dxb_bc7_shape_h1(const uint32_t* pq, const float* mt, uint32_t shape, const float* tot, bool opaque)
{
    float v1[14];
    dxb_f2 V[14];
    dxb_bc7_mt_load(mt, (int)shape, v1);
    for (int k = 0; k < 14; ++k) V[k] = dxb_mk2(tot[k] - v1[k], v1[k]);
    const uint32_t mask = dxb_part2[shape];
    const uint32_t n1 = dxb_popc16(mask);
    dxb_bc7_axis A0, A1;
    dxb_bc7_subset_axes(16u - n1, n1, V, opaque, &A0, &A1);
    int32_t T[16];                         // projection + (subset 1 ? OFF : 0)
    int32_t mnv = 0x7fffffff, mxv = -0x7fffffff, mny = 0x7fffffff, mxy = -0x7fffffff;
#if DXB_ON_DEVICE
    #pragma unroll
#endif
    for (int p = 0; p < 16; p += 2)
    {
        const bool ma = ((mask >> p) & 1u) != 0u, mb = ((mask >> (p + 1)) & 1u) != 0u;
        const int32_t va = dxb_dp4a_u8s8(pq[p], ma ? A1.packed : A0.packed, ma ? DXB_BC7_H1_OFF : 0);
        const int32_t vb = dxb_dp4a_u8s8(pq[p + 1], mb ? A1.packed : A0.packed, mb ? DXB_BC7_H1_OFF : 0);
        const int32_t ya = ma ? va - 2 * DXB_BC7_H1_OFF : va, yb = mb ? vb - 2 * DXB_BC7_H1_OFF : vb;
        T[p] = va; T[p + 1] = vb;
        mnv = dxb_min3_s32(mnv, va, vb); mxv = dxb_max3_s32(mxv, va, vb);
        mny = dxb_min3_s32(mny, ya, yb); mxy = dxb_max3_s32(mxy, ya, yb);
    }
    // subset 0 = the small keys of v and the large keys of y; both subsets of a valid shape are non-empty
    const int32_t tmin0 = mnv, tmax1 = mxv - DXB_BC7_H1_OFF, tmax0 = mxy, tmin1 = mny + DXB_BC7_H1_OFF;
    const float r0 = (float)(tmax0 - tmin0), r1 = (float)(tmax1 - tmin1);
    const float i0 = (r0 > 0.0f) ? 1.0f / r0 : 0.0f, i1 = (r1 > 0.0f) ? 1.0f / r1 : 0.0f;
    const int32_t base1 = tmin1 + DXB_BC7_H1_OFF;
    // squared index rounding errors as packed pairs (units of (range / 7)^2, units of (range / 3)^2), one accumulator per subset
    dxb_f2 E0 = dxb_bc2(0.0f), E1 = dxb_bc2(0.0f);
    const dxb_f2 NL = dxb_mk2(7.0f, 3.0f), MG = dxb_bc2(DXB_MAGIC), nMG = dxb_bc2(-DXB_MAGIC);
#if DXB_ON_DEVICE
    #pragma unroll
#endif
    for (int p = 0; p < 16; ++p)
    {
        const bool m = (T[p] >= (DXB_BC7_H1_OFF >> 1));
        const float x = (float)(T[p] - (m ? base1 : tmin0)) * (m ? i1 : i0);        // position in [0, 1]
        // u = x * (7, 3):  k = rne(u) = fma(x, nl, MAGIC) - MAGIC,  d = fma(x, nl, -k).  Written as explicit fused operations:
        // ptxas contracts a packed multiply that feeds a packed add into FFMA2 even for the .rn forms and with -fmad=false
        // (dxb_portable.h), so an unfused formulation would not be what runs.
        const dxb_f2 x2 = dxb_bc2(x);
        const dxb_f2 K = R2_add2(R2_fma2(x2, NL, MG), nMG);
        const dxb_f2 D = R2_fma2(x2, NL, dxb_mk2(-K.x, -K.y));
        if (m) E1 = R2_fma2(D, D, E1); else E0 = R2_fma2(D, D, E0);
    }
    const float e0a = E0.x, e0b = E0.y, e1a = E1.x, e1b = E1.y;
    // index-quantisation error in pixel units: e * (range / nl)^2 / |a|^2
    const float w0 = (r0 * r0) * A0.inv_aa, w1 = (r1 * r1) * A1.inv_aa;
    const float qb = dxb_fma(e0b, w0, e1b * w1) * (1.0f / 9.0f);
    const float qa = opaque ? dxb_fma(e0a, w0, e1a * w1) * (1.0f / 49.0f) : qb;
    return (A0.resid + A1.resid) + fminf(qa, qb);
}

// rotation heuristic for modes 4/5: cost of coding channel c as the separate scalar and the remaining channels as the
// vector, from the block totals (moments row 64): line-fit residual of the rest + index-quantisation models.
// Returns est[4] for scalar channel c = 0..3 (c = 3: alpha, rotation 0).
DXB_DEV float dxb_bc7_rotation_estimate1(float c00, float c01, float c02, float c11, float c12, float c22, float css, float qfv, float qfs)
{
    // remaining channels: largest eigenvalue by 2 power-iteration steps from the largest-diagonal row + Rayleigh quotient
    const float tr = (c00 + c11) + c22;
    const bool b0 = (c00 >= c11 && c00 >= c22), b1 = !b0 && (c11 >= c22);
    float v0 = b0 ? c00 : (b1 ? c01 : c02), v1 = b0 ? c01 : (b1 ? c11 : c12), v2 = b0 ? c02 : (b1 ? c12 : c22);
    float w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, c02 * v2)), w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, c12 * v2)), w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, c22 * v2));
    v0 = w0; v1 = w1; v2 = w2;
    w0 = dxb_fma(c00, v0, dxb_fma(c01, v1, c02 * v2)); w1 = dxb_fma(c01, v0, dxb_fma(c11, v1, c12 * v2)); w2 = dxb_fma(c02, v0, dxb_fma(c12, v1, c22 * v2));
    const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, v2 * v2)), vw = dxb_fma(v0, w0, dxb_fma(v1, w1, v2 * w2));
    const float lam = (vv > 0.0f) ? fminf(vw / vv, tr) : 0.0f;
    return dxb_fma(lam, qfv, fmaxf(tr - lam, 0.0f)) + css * qfs;
}
// estimate for ONE candidate scalar channel c (0..3; 3 = alpha = rotation 0): the four candidates of a block are evaluated by four
// lanes and exchanged (each lane used to evaluate all four)
DXB_DEV float dxb_bc7_rotation_estimate_c(const float* tot, bool opaque, float qfv, float qfs, uint32_t c)
{
    const float inv = 1.0f / 16.0f, z = opaque ? 0.0f : 1.0f;
    const float c00 = dxb_fma(-tot[0] * inv, tot[0], tot[4]), c01 = dxb_fma(-tot[0] * inv, tot[1], tot[5]), c02 = dxb_fma(-tot[0] * inv, tot[2], tot[6]);
    const float c11 = dxb_fma(-tot[1] * inv, tot[1], tot[8]), c12 = dxb_fma(-tot[1] * inv, tot[2], tot[9]), c22 = dxb_fma(-tot[2] * inv, tot[2], tot[11]);
    const float c03 = z * dxb_fma(-tot[0] * inv, tot[3], tot[7]), c13 = z * dxb_fma(-tot[1] * inv, tot[3], tot[10]);
    const float c23 = z * dxb_fma(-tot[2] * inv, tot[3], tot[12]), c33 = z * dxb_fma(-tot[3] * inv, tot[3], tot[13]);
    // covariance of the three remaining channels (a, b, d) and the variance of the scalar channel
    const float aa = (c == 0u) ? c11 : c00, ab = (c == 0u) ? c12 : ((c == 1u) ? c02 : c01), ad = (c == 0u) ? c13 : ((c == 3u) ? c02 : c03);
    const float bb = (c <= 1u) ? c22 : c11, bd = (c <= 1u) ? c23 : ((c == 2u) ? c13 : c12), dd = (c == 3u) ? c22 : c33;
    const float css = (c == 0u) ? c00 : ((c == 1u) ? c11 : ((c == 2u) ? c22 : c33));
    return dxb_bc7_rotation_estimate1(aa, ab, ad, bb, bd, dd, css, qfv, qfs);
}

// ---------------------------------------------------------------------------------------------------
// branch-free helpers (every lane of the warp runs the same instruction stream whatever its mode)
DXB_DEV float dxb_bit_as_float(uint32_t mask, int i)       // (mask >> i) & 1 as 0.0f / 1.0f without an I2F
{
    return dxb_uint_as_float((0u - ((mask >> i) & 1u)) & 0x3F800000u);
}
// interpolation weight / 64 of (float) index k at nmax = 2^ib - 1: RNE(k * 64 / nmax) / 64 reproduces the
// BC7 weight tables {0,21,43,64} {0,9,18,27,37,46,55,64} {0,4,9,...,60,64} exactly (no product is a tie)
DXB_DEV float dxb_bc7_weightf(float k, float c64 /* 64 / nmax */) { return dxb_rne(k * c64) * (1.0f / 64.0f); }

// ---- endpoint quantisation, float-only (no F2I / I2F / integer shifts in the per-round path) -----------
// A field of `bits` bits, optionally followed by a p-bit (hasP), B = bits + hasP total bits:
//   f = e * (2^B - 1) / 255;  no p-bit: q = round(f);  p-bit p: q = round((f - p) / 2);  q clamped to the field
//   full = 2 q + p (or q);  the decoder reconstructs  deq = (full << (8 - B)) | (full >> (2 B - 8))
// round() is the magic-number RNE; the ">>" is floor(full * 2^(8-2B)) = RNE(full * 2^(8-2B) - (1/2 - 2^-9)),
// exact because the product is a multiple of 2^-8.  All per-mode numbers are lane constants.
struct dxb_bc7_qconst { float scaleH, half, qmax, mul, c8, c2; };

DXB_DEV dxb_bc7_qconst dxb_bc7_make_qconst(uint32_t bits, uint32_t hasP)
{
    const uint32_t B = bits + hasP;
    dxb_bc7_qconst k;
    k.half = hasP ? 0.5f : 1.0f;
    k.scaleH = (float)((1u << B) - 1u) * (1.0f / 255.0f) * k.half;
    k.qmax = (float)((1u << bits) - 1u);
    k.mul = hasP ? 2.0f : 1.0f;
    k.c8 = dxb_uint_as_float((127u + 8u - B) << 23);            // 2^(8-B)
    k.c2 = dxb_uint_as_float((127u + 8u - 2u * B) << 23);       // 2^(8-2B)
    return k;
}
// pE = p * hasP as float.  Returns the field q (float integer); *deq = reconstructed 8-bit value (float integer)
DXB_DEV dxb_f2 dxb_bc7_quant2f(dxb_f2 e, const dxb_bc7_qconst& k, float pE, dxb_f2* deq)
{
    const dxb_f2 MG = dxb_bc2(DXB_MAGIC), nMG = dxb_bc2(-DXB_MAGIC);
    // rne(e * scaleH - pE * half).  Without a p-bit the addend is zero, the FFMA2 degenerates to a packed multiply and ptxas
    // contracts it with the packed add of the rounding constant (dxb_portable.h): that case is therefore WRITTEN as the fused
    // operation, so that the source says what runs (the host emulator executes the same branch).
    dxb_f2 hm;
    if (pE == 0.0f) hm = R3_fma2(e, dxb_bc2(k.scaleH), MG);
    else hm = R3_add2(R3_fma2(e, dxb_bc2(k.scaleH), dxb_bc2(-(pE * k.half))), MG);
    const dxb_f2 hr = R3_add2(hm, nMG);
    const dxb_f2 q = dxb_mk2(fminf(fmaxf(hr.x, 0.0f), k.qmax), fminf(fmaxf(hr.y, 0.0f), k.qmax));
    const dxb_f2 full = R3_fma2(q, dxb_bc2(k.mul), dxb_bc2(pE));
    const dxb_f2 r = R3_add2(R3_add2(R3_fma2(full, dxb_bc2(k.c2), dxb_bc2(-(0.5f - 1.0f / 512.0f))), MG), nMG);
    *deq = R3_fma2(full, dxb_bc2(k.c8), r);
    return q;
}

struct dxb_bc7_modecfg { uint32_t cbits, abits, ptype /*0 none,1 unique,2 shared*/, ib, ib2; };

DXB_DEV dxb_bc7_modecfg dxb_bc7_cfg(int mode)
{
    // packed per mode: cbits | abits<<4 | ptype<<8 | ib<<12 | ib2<<16   (mode table BC6HBC7.cpp:1106-1124)
    const uint32_t t0 = 4u | (0u << 4) | (1u << 8) | (3u << 12) | (0u << 16);
    const uint32_t t2 = 5u | (0u << 4) | (0u << 8) | (2u << 12) | (0u << 16);
    const uint32_t t1 = 6u | (0u << 4) | (2u << 8) | (3u << 12) | (0u << 16);
    const uint32_t t3 = 7u | (0u << 4) | (1u << 8) | (2u << 12) | (0u << 16);
    const uint32_t t4 = 5u | (6u << 4) | (0u << 8) | (2u << 12) | (3u << 16);
    const uint32_t t5 = 7u | (8u << 4) | (0u << 8) | (2u << 12) | (2u << 16);
    const uint32_t t6 = 7u | (7u << 4) | (1u << 8) | (4u << 12) | (0u << 16);
    const uint32_t t7 = 5u | (5u << 4) | (1u << 8) | (2u << 12) | (0u << 16);
    const uint32_t t = (mode == 0) ? t0 : (mode == 2) ? t2 : (mode == 1) ? t1 : (mode == 3) ? t3 : (mode == 4) ? t4 : (mode == 5) ? t5 : (mode == 6) ? t6 : t7;
    dxb_bc7_modecfg c;
    c.cbits = t & 15u; c.abits = (t >> 4) & 15u; c.ptype = (t >> 8) & 15u; c.ib = (t >> 12) & 15u; c.ib2 = (t >> 16) & 15u;
    return c;
}

// ---------------------------------------------------------------------------------------------------
// stage 2: one lane task = one endpoint-pair fit: the pixels of `mask` (a subset of a two-subset shape, or the whole
// block), the channels of `chmask` (bit c = natural channel c), endpoints of `bits` bits (+ a p-bit of type `ptype`:
// 0 none, 1 one per endpoint, 2 one shared by both endpoints), `ib` index bits.  A mode-1/3/7 candidate is two tasks
// (the two subsets), a mode-4/5 candidate is two tasks (the vector channels and the separately coded scalar channel,
// each with its own index set), mode 6 is one task.  px = the block's 16 LDR pixels (floats 0..255), mt = its moment
// table.  Channels outside chmask have zero moments, axis and endpoints, so one instruction stream serves every task.
// Idle lanes (idle = true) run the same code on dummy parameters.  The result's q0/q1 byte c = the field of natural
// channel c (0 for channels outside chmask).
struct dxb_bc7_task { uint32_t shape, mask, chmask, bits, ptype, ib; bool idle, direct; };   // direct: moments summed from the pixels (masks without a row in the stage-1 table: three-subset shapes)

DXB_DEV dxb_bc7_res dxb_bc7_eval(const dxb_px* px, const float* mt, const dxb_bc7_task& T)
{
    const bool idle = T.idle;
    const uint32_t mask = T.mask, shape = T.shape;
    const uint32_t ibc = T.ib;
    float vm[4];
    for (int c = 0; c < 4; ++c) vm[c] = ((T.chmask >> c) & 1u) ? 1.0f : 0.0f;
    // ---- moments of the subset from the stage-1 table (exact integers): subset 1 = row `shape`,
    // subset 0 = totals - row, whole block = totals; masked channels zeroed
    const float n = (float)dxb_popc16(mask);
    float s[4], m00, m01, m02, m03, m11, m12, m13, m22, m23, m33;
    {
        const bool whole = (mask == 0xFFFFu);
        const bool sub0 = !whole && ((mask & 1u) != 0u);        // pixel 0 always belongs to subset 0
        float R[14], TT[14];
        if (T.direct)
        {
            // exact integers < 2^24 in any order, like the table's entries
            for (int k = 0; k < 14; ++k) R[k] = 0.0f;
            for (int i = 0; i < 16; ++i)
            {
                const float f = dxb_bit_as_float(mask, i);
                const dxb_px p = px[i];
                const float x = p.x * f, y = p.y * f, z = p.z * f, w = p.w * f;
                R[0] += x; R[1] += y; R[2] += z; R[3] += w;
                R[4] = dxb_fma(x, p.x, R[4]); R[5] = dxb_fma(x, p.y, R[5]); R[6] = dxb_fma(x, p.z, R[6]); R[7] = dxb_fma(x, p.w, R[7]);
                R[8] = dxb_fma(y, p.y, R[8]); R[9] = dxb_fma(y, p.z, R[9]); R[10] = dxb_fma(y, p.w, R[10]);
                R[11] = dxb_fma(z, p.z, R[11]); R[12] = dxb_fma(z, p.w, R[12]); R[13] = dxb_fma(w, p.w, R[13]);
            }
        }
        else
        {
            dxb_bc7_mt_load(mt, whole ? 64 : (int)shape, R);
            dxb_bc7_mt_load(mt, 64, TT);
            for (int k = 0; k < 14; ++k) R[k] = sub0 ? TT[k] - R[k] : R[k];
        }
        for (int c = 0; c < 4; ++c) s[c] = R[c] * vm[c];
        m00 = R[4] * vm[0]; m01 = R[5] * (vm[0] * vm[1]); m02 = R[6] * (vm[0] * vm[2]); m03 = R[7] * (vm[0] * vm[3]);
        m11 = R[8] * vm[1]; m12 = R[9] * (vm[1] * vm[2]); m13 = R[10] * (vm[1] * vm[3]);
        m22 = R[11] * vm[2]; m23 = R[12] * (vm[2] * vm[3]); m33 = R[13] * vm[3];
    }
    const float inv = 1.0f / fmaxf(n, 1.0f);
    const float mean[4] = { s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv };
    const float c00 = dxb_fma(-mean[0], s[0], m00), c01 = dxb_fma(-mean[0], s[1], m01), c02 = dxb_fma(-mean[0], s[2], m02), c03 = dxb_fma(-mean[0], s[3], m03);
    const float c11 = dxb_fma(-mean[1], s[1], m11), c12 = dxb_fma(-mean[1], s[2], m12), c13 = dxb_fma(-mean[1], s[3], m13);
    const float c22 = dxb_fma(-mean[2], s[2], m22), c23 = dxb_fma(-mean[2], s[3], m23), c33 = dxb_fma(-mean[3], s[3], m33);
    const float tr = (c00 + c11) + (c22 + c33);

    // principal axis: power iteration on the covariance scaled to unit trace (largest eigenvalue >= 1/4, so four
    // un-normalised steps stay far inside the fp32 range), from the row with the largest diagonal; selects, no branches
    float ax[4];
    {
        const float sc = (tr > 1e-3f) ? 1.0f / tr : 0.0f;
        const float k00 = c00 * sc, k01 = c01 * sc, k02 = c02 * sc, k03 = c03 * sc, k11 = c11 * sc;
        const float k12 = c12 * sc, k13 = c13 * sc, k22 = c22 * sc, k23 = c23 * sc, k33 = c33 * sc;
        const bool b0 = (k00 >= k11 && k00 >= k22 && k00 >= k33);
        const bool b1 = !b0 && (k11 >= k22 && k11 >= k33);
        const bool b2 = !b0 && !b1 && (k22 >= k33);
        float v0 = b0 ? k00 : (b1 ? k01 : (b2 ? k02 : k03));
        float v1 = b0 ? k01 : (b1 ? k11 : (b2 ? k12 : k13));
        float v2 = b0 ? k02 : (b1 ? k12 : (b2 ? k22 : k23));
        float v3 = b0 ? k03 : (b1 ? k13 : (b2 ? k23 : k33));
        // matrix columns as packed row pairs: (w0, w1) and (w2, w3) of K v, each component with the scalar association
        const dxb_f2 K0a = dxb_mk2(k00, k01), K1a = dxb_mk2(k01, k11), K2a = dxb_mk2(k02, k12), K3a = dxb_mk2(k03, k13);
        const dxb_f2 K0b = dxb_mk2(k02, k03), K1b = dxb_mk2(k12, k13), K2b = dxb_mk2(k22, k23), K3b = dxb_mk2(k23, k33);
        for (int it = 0; it < DXB_BC7_PCA_ITERS; ++it)
        {
            const dxb_f2 V0 = dxb_bc2(v0), V1 = dxb_bc2(v1), V2 = dxb_bc2(v2), V3 = dxb_bc2(v3);
            const dxb_f2 Wa = R4_fma2(K0a, V0, R4_fma2(K1a, V1, R4_fma2(K2a, V2, R4_mul2(K3a, V3))));
            const dxb_f2 Wb = R4_fma2(K0b, V0, R4_fma2(K1b, V1, R4_fma2(K2b, V2, R4_mul2(K3b, V3))));
            v0 = Wa.x; v1 = Wa.y; v2 = Wb.x; v3 = Wb.y;
        }
        const float vv = dxb_fma(v0, v0, dxb_fma(v1, v1, dxb_fma(v2, v2, v3 * v3)));
        const float r = (vv > 1e-30f) ? 1.0f / sqrtf(vv) : 0.0f;
        ax[0] = v0 * r; ax[1] = v1 * r; ax[2] = v2 * r; ax[3] = v3 * r;
    }

    // ---- projection extents -> initial endpoints (masked channels: axis 0, mean 0 -> endpoints 0)
    float tmin = 3.0e38f, tmax = -3.0e38f;
#if DXB_ON_DEVICE
    #pragma unroll dxb_bc7_pixunroll
#endif
    for (int i = 0; i < 16; ++i)
    {
        const dxb_px p = px[i];
        const dxb_f2 T2 = R4_fma2(R4_add2(dxb_mk2(p.z, p.w), dxb_mk2(-mean[2], -mean[3])), dxb_mk2(ax[2], ax[3]),
                                   R4_mul2(R4_add2(dxb_mk2(p.x, p.y), dxb_mk2(-mean[0], -mean[1])), dxb_mk2(ax[0], ax[1])));
        const float t = T2.x + T2.y;
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

    // ---- evaluation rounds (vector part).  Every vector channel has cbits bits (modes 6/7: abits == cbits).
    const uint32_t hasP = (T.ptype != 0u) ? 1u : 0u;
    const dxb_bc7_qconst qk = dxb_bc7_make_qconst(T.bits, hasP);
    float bestErr = 3.0e38f, bqa0 = 0.0f, bqa1 = 0.0f, bqb0 = 0.0f, bqb1 = 0.0f; uint32_t bpb = 0;
    const float nmaxc = (float)((1u << ibc) - 1u);
    const float c64c = 64.0f / nmaxc;
    bool live = true;                                            // false once this lane has converged (keeps running, results ignored)
#if DXB_ON_DEVICE && defined(DXB_BC7_ROLLROUNDS)
    #pragma unroll 1
#elif DXB_ON_DEVICE
    #pragma unroll
#endif
    for (int round = 0; round < DXB_BC7_ROUNDS; ++round)
    {
        const bool last = (round + 1 == DXB_BC7_ROUNDS);       // compile-time after unrolling: the refit sums vanish from the last round
        dxb_warp_sync();
        dxb_phase_sync();
        // quantise both endpoints for p = 0 and p = 1; fields packed as float integers q0 + 256 q1 + 65536 q2, q3 apart.
        // The two endpoints of a channel travel as one packed pair (x = endpoint 0, y = endpoint 1).
        float qa[2][2], qb[2][2], d0[2][4], d1[2][4], err0[2], err1[2];
        for (int p = 0; p < 2; ++p)
        {
            const float pE = (p && hasP) ? 1.0f : 0.0f;
            dxb_f2 F[4], E2 = dxb_bc2(0.0f);
            for (int c = 0; c < 4; ++c)
            {
                dxb_f2 A;
                const dxb_f2 Ec = dxb_mk2(E0[c], E1[c]), vmc = dxb_bc2(vm[c]);
                F[c] = R5_mul2(dxb_bc7_quant2f(Ec, qk, pE, &A), vmc);
                A = R5_mul2(A, vmc);
                d0[p][c] = A.x; d1[p][c] = A.y;
                const dxb_f2 ea = R5_sub2(A, Ec);                  // masked channels: E = 0 and a = b = 0
                E2 = R5_fma2(ea, ea, E2);
            }
            err0[p] = E2.x; err1[p] = E2.y;
            const dxb_f2 QA = R5_fma2(F[2], dxb_bc2(65536.0f), R5_fma2(F[1], dxb_bc2(256.0f), F[0]));
            qa[p][0] = QA.x; qa[p][1] = QA.y; qb[p][0] = F[3].x; qb[p][1] = F[3].y;
        }
        // p-bit choice: by endpoint reconstruction error, then the overrides; all selects
        uint32_t p0 = (err0[1] < err0[0]) ? 1u : 0u;
        uint32_t p1 = (err1[1] < err1[0]) ? 1u : 0u;
        const uint32_t ps = ((err0[1] + err1[1]) < (err0[0] + err1[0])) ? 1u : 0u;
        if (T.ptype == 2u) { p0 = ps; p1 = ps; }
        if (T.ptype == 0u) { p0 = 0u; p1 = 0u; }
        const float qa0 = p0 ? qa[1][0] : qa[0][0], qb0 = p0 ? qb[1][0] : qb[0][0];
        const float qa1 = p1 ? qa[1][1] : qa[0][1], qb1 = p1 ? qb[1][1] : qb[0][1];
        float D0[4], D1[4];
        for (int c = 0; c < 4; ++c) { D0[c] = p0 ? d0[1][c] : d0[0][c]; D1[c] = p1 ? d1[1][c] : d1[0][c]; }

        const float dx = D1[0] - D0[0], dy = D1[1] - D0[1], dz = D1[2] - D0[2], dw = D1[3] - D0[3];
        const float dd = dxb_fma(dx, dx, dxb_fma(dy, dy, dxb_fma(dz, dz, dw * dw)));
        const float idd = (dd > 0.0f) ? nmaxc / dd : 0.0f;          // index scale folded in
        float err = 0.0f;
        float la = 0.0f, lb = 0.0f, lc = 0.0f;                     // sum (1-s)^2, s(1-s), s^2
        float u0 = 0, u1 = 0, u2 = 0, u3 = 0, v0 = 0, v1 = 0, v2 = 0, v3 = 0;     // sum (1-s) p, sum s p
        // channel pairs (x, y) and (z, w) as packed fp32 (dxb_portable.h): same IEEE operations, half the issue slots
        const dxb_f2 vm01 = dxb_mk2(vm[0], vm[1]), vm23 = dxb_mk2(vm[2], vm[3]);
        const dxb_f2 nD01 = dxb_mk2(-D0[0], -D0[1]), nD23 = dxb_mk2(-D0[2], -D0[3]);
        const dxb_f2 d01 = dxb_mk2(dx, dy), d23 = dxb_mk2(dz, dw);
        const dxb_f2 Dc01 = dxb_mk2(D0[0] + (1.0f / 128.0f), D0[1] + (1.0f / 128.0f)), Dc23 = dxb_mk2(D0[2] + (1.0f / 128.0f), D0[3] + (1.0f / 128.0f));
        const dxb_f2 MG = dxb_bc2(DXB_MAGIC), nMG = dxb_bc2(-DXB_MAGIC);
        dxb_f2 V01 = dxb_bc2(0.0f), V23 = dxb_bc2(0.0f);
#if DXB_ON_DEVICE
        #pragma unroll dxb_bc7_pixunroll
#endif
        for (int i = 0; i < 16; ++i)
        {
            const float f = dxb_bit_as_float(mask, i);          // 1 if the pixel belongs to this lane's subset
            const dxb_px p = px[i];
            const dxb_f2 P01 = R6_mul2(dxb_mk2(p.x, p.y), vm01), P23 = R6_mul2(dxb_mk2(p.z, p.w), vm23);
            const dxb_f2 A01 = R6_add2(P01, nD01), A23 = R6_add2(P23, nD23);
            const dxb_f2 T = R6_fma2(A23, d23, R6_mul2(A01, d01));
            const float pr = T.x + T.y;                           // (P - D0) . d
            const float tk = pr * idd;
            // index = nearest of the uniformly spaced positions (stage 4 assigns the winner's final indices exhaustively)
            const float kk = dxb_rne(fminf(fmaxf(tk, 0.0f), nmaxc));
            const float sk = dxb_bc7_weightf(kk, c64c);
            // candidate error against the decoder's palette entry (e0 (64 - w) + e1 w + 32) >> 6 = round-half-up of D0 + sk d
            // (a multiple of 1/64, so adding 1/128 before the RNE never ties).  An unrounded model mis-ranks near-lossless
            // candidates: the rounding noise (1/12 per value) is half of the error of a smooth 8-bit gradient.
            const dxb_f2 sk2 = dxb_bc2(sk);
            const dxb_f2 q01 = R6_add2(R6_add2(R6_fma2(d01, sk2, Dc01), MG), nMG), q23 = R6_add2(R6_add2(R6_fma2(d23, sk2, Dc23), MG), nMG);
            const dxb_f2 e01 = R6_sub2(P01, q01), e23 = R6_sub2(P23, q23);
            const dxb_f2 sq = R6_fma2(e23, e23, R6_mul2(e01, e01));
            err = dxb_fma(f, sq.x + sq.y, err);
            if (!last)
            {
                // refit sums: only sum f s, sum f s^2 and sum f s P are accumulated; the (1 - s) sums follow from the
                // subset's pixel count and channel sums (stage-1 moments) after the loop
                const float skf = sk * f;
                lb += skf; lc = dxb_fma(skf, sk, lc);
                const dxb_f2 skf2 = dxb_bc2(skf);
                V01 = R6_fma2(skf2, P01, V01); V23 = R6_fma2(skf2, P23, V23);
            }
        }
        v0 = V01.x; v1 = V01.y; v2 = V23.x; v3 = V23.y;
        if (!last)
        {
            const float fs = lb;                                   // sum f s
            lb = fs - lc;                                          // sum f s (1 - s)
            la = (n - fs) - lb;                                    // sum f (1 - s)^2 = n - 2 sum f s + sum f s^2
            u0 = s[0] - v0; u1 = s[1] - v1; u2 = s[2] - v2; u3 = s[3] - v3;
        }
        const bool better = live && (err < bestErr);
        bestErr = better ? err : bestErr; bpb = better ? (p0 | (p1 << 1)) : bpb;
        bqa0 = better ? qa0 : bqa0; bqa1 = better ? qa1 : bqa1; bqb0 = better ? qb0 : bqb0; bqb1 = better ? qb1 : bqb1;
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
    // natural-order fields as integers: n0 / n1 byte c = field of natural channel c (bqa = f0 + 256 f1 + 65536 f2 < 2^24: exact)
    const uint32_t n0 = (uint32_t)dxb_f2i(bqa0) | ((uint32_t)dxb_f2i(bqb0) << 24);
    const uint32_t n1 = (uint32_t)dxb_f2i(bqa1) | ((uint32_t)dxb_f2i(bqb1) << 24);
    dxb_bc7_res R;
    R.err = idle ? 3.0e38f : bestErr; R.q0 = n0; R.q1 = n1; R.pbits = bpb;
    return R;
}

// natural channel order -> the bit stream's slot order: a rotation swaps bytes rot-1 and 3
DXB_DEV uint32_t dxb_bc7_rotate_fields(uint32_t n, uint32_t rot)
{
    if (rot == 0u) return n;
    const uint32_t sh = 8u * (rot - 1u);
    const uint32_t a = (n >> sh) & 0xFFu, b = n >> 24;
    return (n & ~((0xFFu << sh) | 0xFF000000u)) | (b << sh) | (a << 24);
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
// The encoder proper, SPMD over the 32 lanes of one warp: TWO blocks per warp, one per 16-lane half.
//   S->px : LDR pixels of both blocks (floats 0..255): S->px[16 h + i] = pixel i of the half-h block
//   out0/out1 : 16 output bytes of the half-0 / half-1 block (nullptr = that half carries no block)
// Everything "per block" below is a lane-private value that is uniform inside a half.
struct dxb_bc7_win { uint32_t mode, shape, rot, idx, q0[3], q1[3], pb[3]; };
#if !DXB_ON_DEVICE
// test-infrastructure hook of the host emulator only (tools/bc_quality.py experiments): force the first candidate
// shape of the half-0 / half-1 block (-1 = none)
static thread_local int dxb_bc7_dbg_force_shape[2] = { -1, -1 };
#endif

// THREE: compile the three-subset pass (TEX_COMPRESS_BC7_USE_3SUBSETS); the default kernel is instantiated without it so that
// its instruction footprint (the kernel is sensitive to instruction-cache misses) does not grow for a non-default flag.
template <bool THREE>
DXB_DEV void dxb_bc7_encode_pair(dxb_bc7_scratch* S, uint32_t bcflags, uint8_t* out0, uint8_t* out1)
{
    const bool quick = (bcflags & DXB_BC_FLAGS_FORCE_BC7_MODE6) != 0;

    // ---- stage 1: moment table (tensor cores), 4 shapes per lane ranked, 3 best kept per block
    dxb_bc7_build_moments(S);
    uint32_t hasA[DXB_NL], sel[3][DXB_NL];
    {
        uint32_t k0[DXB_NL], k1[DXB_NL], k2[DXB_NL];             // each lane's three best keys, ascending
        DXB_LANES_BEGIN
            const float* mt = S->mt[lane >> 4];
            float tot[14];
            dxb_bc7_mt_load(mt, 64, tot);
            hasA[L] = (tot[3] != 4080.0f) ? 1u : 0u;             // 16 * 255: every alpha is 255
            // index quantisation factor 1/(2^b-1)^2: 3-bit for mode 1 (opaque), 2-bit for mode 7 (alpha)
            const float qf = hasA[L] ? (1.0f / 9.0f) : (1.0f / 49.0f);
            uint32_t a = 0xFFFFFFFFu, b = 0xFFFFFFFFu, c = 0xFFFFFFFFu;
#if defined(DXB_BC7_PRUNE)
            // per-lane pruning: the closed-form estimate keeps DXB_BC7_PRUNE of this lane's 4 shapes for the h1 estimate
            uint32_t pk[4];
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t shape = (uint32_t)(lane & 15) + 16u * (uint32_t)j;
                const float e0 = quick ? 0.0f : dxb_bc7_shape_estimate(mt, shape, qf, tot, hasA[L] == 0u);
                pk[j] = (dxb_float_as_uint(e0) & 0xFFFFFFC0u) | shape;
            }
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3 - i; ++j) if (pk[j + 1] < pk[j]) { const uint32_t t = pk[j]; pk[j] = pk[j + 1]; pk[j + 1] = t; }
            for (int j = 0; j < DXB_BC7_PRUNE; ++j)
            {
                const uint32_t shape = pk[j] & 63u;
                const float e = quick ? 0.0f : dxb_bc7_shape_h1(S->pq + (lane & 16), mt, shape, tot, hasA[L] == 0u);
                const uint32_t x = (dxb_float_as_uint(e) & 0xFFFFFFC0u) | shape;
                const uint32_t lo = (x < a) ? x : a, hi = (x < a) ? a : x;               // sorted insert
                const uint32_t lo2 = (hi < b) ? hi : b, hi2 = (hi < b) ? b : hi;
                a = lo; b = lo2; c = (hi2 < c) ? hi2 : c;
            }
#else
#if DXB_ON_DEVICE
            #pragma unroll 1
#endif
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t shape = (uint32_t)(lane & 15) + 16u * (uint32_t)j;
#ifdef DXB_BC7_EST_H0
                const float e = quick ? 0.0f : dxb_bc7_shape_estimate(mt, shape, qf, tot, hasA[L] == 0u);
#else
                (void)qf;
                const float e = quick ? 0.0f : dxb_bc7_shape_h1(S->pq + (lane & 16), mt, shape, tot, hasA[L] == 0u);
#endif
                const uint32_t x = (dxb_float_as_uint(e) & 0xFFFFFFC0u) | shape;
                const uint32_t lo = (x < a) ? x : a, hi = (x < a) ? a : x;               // sorted insert
                const uint32_t lo2 = (hi < b) ? hi : b, hi2 = (hi < b) ? b : hi;
                a = lo; b = lo2; c = (hi2 < c) ? hi2 : c;
            }
#endif
            k0[L] = a; k1[L] = b; k2[L] = c;
        DXB_LANES_END
        for (int r = 0; r < 3; ++r)
        {
            uint32_t win[DXB_NL];
            dxb_half_min_u32(k0, win);
            DXB_LANES_BEGIN
                sel[r][L] = win[L] & 63u;
                if (k0[L] == win[L]) { k0[L] = k1[L]; k1[L] = k2[L]; k2[L] = 0xFFFFFFFFu; }
            DXB_LANES_END
        }
    }

#if !DXB_ON_DEVICE
    DXB_LANES_BEGIN
        if (dxb_bc7_dbg_force_shape[lane >> 4] >= 0) sel[0][L] = (uint32_t)dxb_bc7_dbg_force_shape[lane >> 4];
    DXB_LANES_END
#endif
    dxb_phase_sync();
    // ---- stage 2: one fit task per lane (dxb_bc7_eval); a candidate encoding = one task or the sum of two.
    //   opaque block (the modes the reference tries, BC6HBC7.cpp:2803-2821: 1, 3, 4, 5, 6):
    //     0-7   2 best shapes x 2 subsets x {mode 1, mode 3}          (lane = 4 shape + 2 subset + modebit, partner lane ^ 2)
    //     8     mode 6
    //     9-12  mode 4, rotation r*: vector 2-bit, vector 3-bit, scalar 3-bit, scalar 2-bit   (index selector 0 = 9 + 11, 1 = 10 + 12)
    //     13-14 mode 5, rotation r*: vector, scalar
    //   alpha block (modes 4, 5, 6, 7):
    //     0-5   3 best shapes x 2 subsets x mode 7                     (lane = 2 shape + subset, partner lane ^ 1)
    //     6     mode 6
    //     7-8   mode 5, rotation r*: vector, scalar;   9-10  mode 5, second-best rotation
    //     11-14 mode 4, rotation r*: vector 2-bit, vector 3-bit, scalar 3-bit, scalar 2-bit
    //   r* = the rotation (scalar channel) with the smallest modelled cost (dxb_bc7_rotation_estimates); the reference
    //   tries every rotation (:2823-2829).  Lane 15 idles.
    uint32_t tMeta[DXB_NL], rErr[DXB_NL], rQ0[DXB_NL], rQ1[DXB_NL], partner[DXB_NL];
    // rotation ranking from the block totals: lane c (mod 4) evaluates scalar channel c, the four keys are exchanged.
    // keys: estimate bits | rotation (channel c = 3 is rotation 0, channel c < 3 rotation c + 1); opaque blocks never rotate alpha
    uint32_t rk[DXB_NL], rk0[DXB_NL], rk1[DXB_NL], rk2[DXB_NL], rk3[DXB_NL], li0[DXB_NL], li1[DXB_NL], li2[DXB_NL], li3[DXB_NL];
    DXB_LANES_BEGIN
        float tot[14];
        dxb_bc7_mt_load(S->mt[lane >> 4], 64, tot);
        const uint32_t c = (uint32_t)lane & 3u;
        const float est = dxb_bc7_rotation_estimate_c(tot, hasA[L] == 0u, 1.0f / 9.0f, 1.0f / 49.0f, c);
        rk[L] = (c == 3u && !hasA[L]) ? 0xFFFFFFFFu : ((dxb_float_as_uint(est) & 0xFFFFFFFCu) | ((c + 1u) & 3u));
        li0[L] = 0u; li1[L] = 1u; li2[L] = 2u; li3[L] = 3u;
    DXB_LANES_END
    dxb_half_gather_u32(rk, li0, rk0); dxb_half_gather_u32(rk, li1, rk1); dxb_half_gather_u32(rk, li2, rk2); dxb_half_gather_u32(rk, li3, rk3);
    DXB_LANES_BEGIN
        const int hl = lane & 15;
        int mode = -1, idxMode = 0, part = hl;
        uint32_t rot = 0;
        dxb_bc7_task T;
        T.shape = 0; T.mask = 0xFFFFu; T.chmask = 0xFu; T.bits = 7u; T.ptype = 1u; T.ib = 4u; T.idle = false; T.direct = false;
        uint32_t r1 = 0, r2 = 0, r4 = 0;     // mode 5 rotations (best, second), mode 4 rotation
        {
            const uint32_t keys[4] = { rk0[L], rk1[L], rk2[L], rk3[L] };
            uint32_t ka = 0xFFFFFFFFu, kb = 0xFFFFFFFFu;
            for (uint32_t c = 0; c < 4; ++c)
            {
                const uint32_t x = keys[c];
                const uint32_t lo = (x < ka) ? x : ka, hi = (x < ka) ? ka : x;
                ka = lo; kb = (hi < kb) ? hi : kb;
            }
            r1 = ka & 3u; r2 = kb & 3u;
            // alpha blocks: alpha is nearly always the channel to separate (rotation 0); the model decides the second rotation
            r4 = r1;
            if (hasA[L]) { r2 = (r1 != 0u) ? r1 : r2; r1 = 0u; }
        }
        int kind = -1;                     // 0 two-subset, 1 mode 6, 2 mode-4 part (sub = 0..3), 3 mode-5 part (sub = 0..1)
        int sub = 0;
        if (!hasA[L])
        {
            if (hl < 8)
            {
                kind = 0;
                T.shape = (hl >> 2) ? sel[1][L] : sel[0][L];
                const uint32_t m1 = dxb_part2[T.shape];
                T.mask = ((hl >> 1) & 1) ? m1 : (~m1 & 0xFFFFu);
                mode = (hl & 1) ? 3 : 1;
                T.chmask = 0x7u; T.bits = (hl & 1) ? 7u : 6u; T.ptype = (hl & 1) ? 1u : 2u; T.ib = (hl & 1) ? 2u : 3u;
                part = hl ^ 2;
            }
            else if (hl == 8) { kind = 1; mode = 6; }
            else if (hl < 13) { kind = 2; sub = hl - 9; rot = r4; part = (sub < 2) ? hl + 2 : hl - 2; }
            else if (hl < 15) { kind = 3; sub = hl - 13; rot = r1; part = (sub == 0) ? 14 : 13; }
        }
        else
        {
            if (hl < 6)
            {
                kind = 0;
                const int k = hl >> 1;
                T.shape = (k == 0) ? sel[0][L] : (k == 1) ? sel[1][L] : sel[2][L];
                const uint32_t m1 = dxb_part2[T.shape];
                T.mask = (hl & 1) ? m1 : (~m1 & 0xFFFFu);
                mode = 7;
                T.chmask = 0xFu; T.bits = 5u; T.ptype = 1u; T.ib = 2u;
                part = hl ^ 1;
            }
            else if (hl == 6) { kind = 1; mode = 6; }
            else if (hl < 11) { kind = 3; sub = (hl - 7) & 1; rot = (hl < 9) ? r1 : r2; part = sub ? hl - 1 : hl + 1; }
            else if (hl < 15) { kind = 2; sub = hl - 11; rot = r4; part = (sub < 2) ? hl + 2 : hl - 2; }
        }
        if (kind == 2 || kind == 3)
        {
            const uint32_t sc = rot ? rot - 1u : 3u;                      // natural channel of the separately coded scalar
            const bool scalar = (kind == 2) ? (sub >= 2) : (sub == 1);
            T.chmask = scalar ? (1u << sc) : (0xFu & ~(1u << sc));
            T.ptype = 0u;
            if (kind == 2)
            {
                mode = 4; idxMode = sub & 1;                               // candidates: vector 2-bit + scalar 3-bit (selector 0), vector 3-bit + scalar 2-bit (1)
                T.bits = scalar ? 6u : 5u;
                T.ib = (sub == 0 || sub == 3) ? 2u : 3u;
                idxMode = (sub == 1 || sub == 3) ? 1 : 0;
            }
            else { mode = 5; T.bits = scalar ? 8u : 7u; T.ib = 2u; }
        }
        if (kind < 0 || (quick && mode != 6)) { T.idle = true; mode = -1; part = hl; }
        const dxb_bc7_res res = dxb_bc7_eval(S->px + (lane & 16), S->mt[lane >> 4], T);
        // meta word: mode(3) | shape(6) << 3 | rot(2) << 9 | idx(1) << 11 | pbits(2) << 12
        tMeta[L] = ((uint32_t)mode & 7u) | (T.shape << 3) | (rot << 9) | ((uint32_t)idxMode << 11) | (res.pbits << 12);
        rErr[L] = (mode < 0) ? 0x03FFFFFFu : (uint32_t)dxb_f2i(fminf(res.err, 6.0e7f));
        rQ0[L] = res.q0; rQ1[L] = res.q1; partner[L] = (uint32_t)part;
    DXB_LANES_END

    dxb_phase_sync();
    // ---- stage 3: candidate error = the task's error + its partner's; winner of each half by integer key (ties: lowest lane,
    // which is the subset-0 / vector lane of its pair)
    dxb_bc7_win W[DXB_NL];
    uint32_t wkeyA[DXB_NL];
    {
        uint32_t pe[DXB_NL], key[DXB_NL], wkey[DXB_NL], src[DXB_NL], src1[DXB_NL], wm[DXB_NL];
        uint32_t g0[DXB_NL], g1[DXB_NL], g2[DXB_NL], h0[DXB_NL], h1[DXB_NL], h2[DXB_NL];
        dxb_half_gather_u32(rErr, partner, pe);
        DXB_LANES_BEGIN
            uint32_t e = rErr[L];
            if (partner[L] != (uint32_t)(lane & 15)) e += pe[L];
            e = (e > 0x03FFFFFFu) ? 0x03FFFFFFu : e;
            key[L] = (e << 5) | (uint32_t)(lane & 15);
        DXB_LANES_END
        dxb_half_min_u32(key, wkey);
        DXB_LANES_BEGIN
            src[L] = wkey[L] & 15u; wkeyA[L] = wkey[L];
        DXB_LANES_END
        dxb_half_gather_u32(partner, src, src1);
        dxb_half_gather_u32(tMeta, src, wm);
        dxb_half_gather_u32(rQ0, src, g0); dxb_half_gather_u32(rQ1, src, g1); dxb_half_gather_u32(tMeta, src, g2);
        dxb_half_gather_u32(rQ0, src1, h0); dxb_half_gather_u32(rQ1, src1, h1); dxb_half_gather_u32(tMeta, src1, h2);
        DXB_LANES_BEGIN
            W[L].mode = wm[L] & 7u; W[L].shape = (wm[L] >> 3) & 63u; W[L].rot = (wm[L] >> 9) & 3u; W[L].idx = (wm[L] >> 11) & 1u;
            const bool sepA = (W[L].mode == 4u || W[L].mode == 5u);
            // modes 4/5: vector fields | scalar field (disjoint bytes, natural order) -> slot order
            const uint32_t a0 = sepA ? dxb_bc7_rotate_fields(g0[L] | h0[L], W[L].rot) : g0[L];
            const uint32_t a1 = sepA ? dxb_bc7_rotate_fields(g1[L] | h1[L], W[L].rot) : g1[L];
            W[L].q0[0] = a0; W[L].q1[0] = a1; W[L].pb[0] = (g2[L] >> 12) & 3u;
            W[L].q0[1] = sepA ? a0 : h0[L]; W[L].q1[1] = sepA ? a1 : h1[L]; W[L].pb[1] = (h2[L] >> 12) & 3u;
            W[L].q0[2] = 0u; W[L].q1[2] = 0u; W[L].pb[2] = 0u;
        DXB_LANES_END
    }

    // ---- three-subset modes 0 and 2 (TEX_COMPRESS_BC7_USE_3SUBSETS; the reference tries them only with the flag,
    // BC6HBC7.cpp:2807): a second pass of lane tasks, taken only with the flag.  Both modes are RGB-only (alpha decodes as 255), so
    // blocks with alpha skip the pass.
    //   ranking   every three-subset shape (4 per lane) by the closed-form line-fit estimate of its three subsets; moments of
    //             subsets 1 and 2 are summed from the pixels, subset 0 = block totals - the others.  Mode 0 has 4 partition
    //             bits: shapes 0..15 = the first shape of every lane.
    //   tasks     lanes 0-8 = mode 2, 3 best shapes x 3 subsets;  lanes 9-14 = mode 0, 2 best of shapes 0..15 x 3 subsets
    //   winner    candidate = three consecutive lanes; replaces the winner of the first pass when its error is smaller
    if (THREE && (bcflags & DXB_BC_FLAGS_USE_3SUBSETS) != 0u && !quick)
    {
        uint32_t k0[DXB_NL], k1[DXB_NL], k2[DXB_NL], z0[DXB_NL], selB[3][DXB_NL], selZ[2][DXB_NL];
        DXB_LANES_BEGIN
            const dxb_px* px = S->px + (lane & 16);
            float tot[14];
            dxb_bc7_mt_load(S->mt[lane >> 4], 64, tot);
            uint32_t a = 0xFFFFFFFFu, b = 0xFFFFFFFFu, c = 0xFFFFFFFFu;
            z0[L] = 0xFFFFFFFFu;
#if DXB_ON_DEVICE
            #pragma unroll 1
#endif
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t shape = (uint32_t)(lane & 15) + 16u * (uint32_t)j;
                const uint32_t part = dxb_part3[shape];
                float s1[14], s2[14], s0[14];
                for (int k = 0; k < 14; ++k) { s1[k] = 0.0f; s2[k] = 0.0f; }
                uint32_t n1 = 0, n2 = 0;
                for (int i = 0; i < 16; ++i)
                {
                    const uint32_t sub = (part >> (2 * i)) & 3u;
                    const float f1 = (sub == 1u) ? 1.0f : 0.0f, f2 = (sub == 2u) ? 1.0f : 0.0f;
                    n1 += (sub == 1u) ? 1u : 0u; n2 += (sub == 2u) ? 1u : 0u;
                    const dxb_px p = px[i];
                    const float xx = p.x * p.x, xy = p.x * p.y, xz = p.x * p.z, yy = p.y * p.y, yz = p.y * p.z, zz = p.z * p.z;
                    s1[0] = dxb_fma(f1, p.x, s1[0]); s1[1] = dxb_fma(f1, p.y, s1[1]); s1[2] = dxb_fma(f1, p.z, s1[2]);
                    s1[4] = dxb_fma(f1, xx, s1[4]); s1[5] = dxb_fma(f1, xy, s1[5]); s1[6] = dxb_fma(f1, xz, s1[6]);
                    s1[8] = dxb_fma(f1, yy, s1[8]); s1[9] = dxb_fma(f1, yz, s1[9]); s1[11] = dxb_fma(f1, zz, s1[11]);
                    s2[0] = dxb_fma(f2, p.x, s2[0]); s2[1] = dxb_fma(f2, p.y, s2[1]); s2[2] = dxb_fma(f2, p.z, s2[2]);
                    s2[4] = dxb_fma(f2, xx, s2[4]); s2[5] = dxb_fma(f2, xy, s2[5]); s2[6] = dxb_fma(f2, xz, s2[6]);
                    s2[8] = dxb_fma(f2, yy, s2[8]); s2[9] = dxb_fma(f2, yz, s2[9]); s2[11] = dxb_fma(f2, zz, s2[11]);
                }
                for (int k = 0; k < 14; ++k) s0[k] = (tot[k] - s1[k]) - s2[k];
                const uint32_t n0 = 16u - n1 - n2;
                // mode 2: 2-bit indices; mode 0: 3-bit indices
                const float e2 = (dxb_bc7_subset_estimate3(n0, s0, 1.0f / 9.0f) + dxb_bc7_subset_estimate3(n1, s1, 1.0f / 9.0f)) + dxb_bc7_subset_estimate3(n2, s2, 1.0f / 9.0f);
                const uint32_t x = (dxb_float_as_uint(e2) & 0xFFFFFFC0u) | shape;
                const uint32_t lo = (x < a) ? x : a, hi = (x < a) ? a : x;               // sorted insert
                const uint32_t lo2 = (hi < b) ? hi : b, hi2 = (hi < b) ? b : hi;
                a = lo; b = lo2; c = (hi2 < c) ? hi2 : c;
                if (j == 0)
                {
                    const float e0 = (dxb_bc7_subset_estimate3(n0, s0, 1.0f / 49.0f) + dxb_bc7_subset_estimate3(n1, s1, 1.0f / 49.0f)) + dxb_bc7_subset_estimate3(n2, s2, 1.0f / 49.0f);
                    z0[L] = (dxb_float_as_uint(e0) & 0xFFFFFFC0u) | shape;
                }
            }
            k0[L] = a; k1[L] = b; k2[L] = c;
        DXB_LANES_END
        for (int r = 0; r < 3; ++r)
        {
            uint32_t win[DXB_NL];
            dxb_half_min_u32(k0, win);
            DXB_LANES_BEGIN
                selB[r][L] = win[L] & 63u;
                if (k0[L] == win[L]) { k0[L] = k1[L]; k1[L] = k2[L]; k2[L] = 0xFFFFFFFFu; }
            DXB_LANES_END
        }
        for (int r = 0; r < 2; ++r)
        {
            uint32_t win[DXB_NL];
            dxb_half_min_u32(z0, win);
            DXB_LANES_BEGIN
                selZ[r][L] = win[L] & 63u;
                if (z0[L] == win[L]) z0[L] = 0xFFFFFFFFu;
            DXB_LANES_END
        }
        dxb_phase_sync();
        uint32_t bMeta[DXB_NL], bErr[DXB_NL], bQ0[DXB_NL], bQ1[DXB_NL], n1i[DXB_NL], n2i[DXB_NL];
        DXB_LANES_BEGIN
            const int hl = lane & 15;
            const bool m2 = (hl < 9);
            const int t = m2 ? hl : hl - 9;
            const int k = t / 3, sub = t - 3 * k;
            dxb_bc7_task T;
            T.shape = m2 ? ((k == 0) ? selB[0][L] : (k == 1) ? selB[1][L] : selB[2][L]) : ((k == 0) ? selZ[0][L] : selZ[1][L]);
            const uint32_t part = dxb_part3[T.shape];
            uint32_t mask = 0;
            for (int i = 0; i < 16; ++i) mask |= (((part >> (2 * i)) & 3u) == (uint32_t)sub) ? (1u << i) : 0u;
            T.mask = mask; T.chmask = 0x7u; T.bits = m2 ? 5u : 4u; T.ptype = m2 ? 0u : 1u; T.ib = m2 ? 2u : 3u;
            T.direct = true; T.idle = (hl == 15) || (hasA[L] != 0u);
            const dxb_bc7_res res = dxb_bc7_eval(S->px + (lane & 16), S->mt[lane >> 4], T);
            bMeta[L] = (m2 ? 2u : 0u) | (T.shape << 3) | (res.pbits << 12) | ((uint32_t)sub << 16);
            bErr[L] = T.idle ? 0x03FFFFFFu : (uint32_t)dxb_f2i(fminf(res.err, 6.0e7f));
            bQ0[L] = res.q0; bQ1[L] = res.q1;
            n1i[L] = (uint32_t)((hl + 1) & 15); n2i[L] = (uint32_t)((hl + 2) & 15);
        DXB_LANES_END
        dxb_phase_sync();
        uint32_t e1[DXB_NL], e2[DXB_NL], key[DXB_NL], wkey[DXB_NL], src[DXB_NL], src1[DXB_NL], src2[DXB_NL];
        dxb_half_gather_u32(bErr, n1i, e1); dxb_half_gather_u32(bErr, n2i, e2);
        DXB_LANES_BEGIN
            uint32_t e = bErr[L] + e1[L] + e2[L];
            e = (e > 0x03FFFFFFu || ((bMeta[L] >> 16) & 3u) != 0u || (lane & 15) == 15) ? 0x03FFFFFFu : e;
            key[L] = (e << 5) | (uint32_t)(lane & 15);
        DXB_LANES_END
        dxb_half_min_u32(key, wkey);
        DXB_LANES_BEGIN
            src[L] = wkey[L] & 15u; src1[L] = (src[L] + 1u) & 15u; src2[L] = (src[L] + 2u) & 15u;
        DXB_LANES_END
        uint32_t gm[DXB_NL], g0[DXB_NL], g1[DXB_NL], h0[DXB_NL], h1[DXB_NL], hm[DXB_NL], i0[DXB_NL], i1[DXB_NL], im[DXB_NL];
        dxb_half_gather_u32(bMeta, src, gm); dxb_half_gather_u32(bQ0, src, g0); dxb_half_gather_u32(bQ1, src, g1);
        dxb_half_gather_u32(bMeta, src1, hm); dxb_half_gather_u32(bQ0, src1, h0); dxb_half_gather_u32(bQ1, src1, h1);
        dxb_half_gather_u32(bMeta, src2, im); dxb_half_gather_u32(bQ0, src2, i0); dxb_half_gather_u32(bQ1, src2, i1);
        DXB_LANES_BEGIN
            if ((wkey[L] >> 5) < (wkeyA[L] >> 5))
            {
                W[L].mode = gm[L] & 7u; W[L].shape = (gm[L] >> 3) & 63u; W[L].rot = 0u; W[L].idx = 0u;
                W[L].q0[0] = g0[L]; W[L].q1[0] = g1[L]; W[L].pb[0] = (gm[L] >> 12) & 3u;
                W[L].q0[1] = h0[L]; W[L].q1[1] = h1[L]; W[L].pb[1] = (hm[L] >> 12) & 3u;
                W[L].q0[2] = i0[L]; W[L].q1[2] = i1[L]; W[L].pb[2] = (im[L] >> 12) & 3u;
            }
        DXB_LANES_END
    }

    // ---- stage 4: every lane = one pixel of its block: exhaustive nearest palette entry
    uint32_t idxC[DXB_NL], idxA[DXB_NL], anchor1Src[DXB_NL], anchor2Src[DXB_NL], zeroSrc[DXB_NL];
    DXB_LANES_BEGIN
        const int hl = lane & 15;
        const uint32_t wMode = W[L].mode;
        const dxb_bc7_modecfg cfg = dxb_bc7_cfg((int)wMode);
        const bool two = (wMode == 1u || wMode == 3u || wMode == 7u);
        const bool three = (wMode == 0u || wMode == 2u);
        const bool sepA = (wMode == 4u || wMode == 5u);
        const uint32_t ibc = (wMode == 4u && W[L].idx) ? 3u : cfg.ib;
        const uint32_t iba = (wMode == 4u) ? (W[L].idx ? 2u : 3u) : cfg.ib2;
        // subset of this lane's pixel
        const uint32_t sb = three ? ((dxb_part3[W[L].shape] >> (2 * hl)) & 3u) : (two ? ((dxb_part2[W[L].shape] >> hl) & 1u) : 0u);
        const uint32_t q0 = (sb == 2u) ? W[L].q0[2] : (sb == 1u) ? W[L].q0[1] : W[L].q0[0];
        const uint32_t q1 = (sb == 2u) ? W[L].q1[2] : (sb == 1u) ? W[L].q1[1] : W[L].q1[0];
        const uint32_t pb = (sb == 2u) ? W[L].pb[2] : (sb == 1u) ? W[L].pb[1] : W[L].pb[0];
        const uint32_t hasP = (cfg.ptype != 0 && !sepA) ? 1u : 0u;
        int32_t e0[4], e1[4];
        for (uint32_t c = 0; c < 4; ++c)
        {
            const uint32_t bits = (c == 3) ? cfg.abits : cfg.cbits;
            const bool coded = (c < 3) || (cfg.abits != 0);
            e0[c] = coded ? (int32_t)dxb_bc7_deq_field((q0 >> (8 * c)) & 0xFF, bits, hasP, pb & 1u) : 255;
            e1[c] = coded ? (int32_t)dxb_bc7_deq_field((q1 >> (8 * c)) & 0xFF, bits, hasP, (pb >> 1) & 1u) : 255;
        }
        const dxb_px pr = dxb_bc7_rotate(S->px[lane], (int)W[L].rot);
        const int32_t p[4] = { dxb_f2i(pr.x), dxb_f2i(pr.y), dxb_f2i(pr.z), dxb_f2i(pr.w) };
        idxA[L] = 0;
        if (sepA)
        {
            idxC[L] = dxb_bc7_nearest(p, e0, e1, 0, 3, ibc);
            idxA[L] = dxb_bc7_nearest(p, e0, e1, 3, 4, iba);
        }
        else
            idxC[L] = dxb_bc7_nearest(p, e0, e1, 0, (wMode == 6u || wMode == 7u) ? 4 : 3, ibc);
        anchor1Src[L] = three ? dxb_anchor3a[W[L].shape] : (two ? dxb_anchor2[W[L].shape] : 0u);
        anchor2Src[L] = three ? dxb_anchor3b[W[L].shape] : 0u;
        zeroSrc[L] = 0u;
    DXB_LANES_END

    // anchor fix-up: the anchor index of each subset must have its MSB clear; otherwise swap that
    // subset's endpoints and mirror its indices (weights are symmetric: w[n-k] = 64 - w[k])
    uint32_t aC0[DXB_NL], aC1[DXB_NL], aC2[DXB_NL], aA0[DXB_NL];
    dxb_half_gather_u32(idxC, zeroSrc, aC0);
    dxb_half_gather_u32(idxC, anchor1Src, aC1);
    dxb_half_gather_u32(idxC, anchor2Src, aC2);
    dxb_half_gather_u32(idxA, zeroSrc, aA0);

    // bit layout (D3DX_BC7::Decode, BC6HBC7.cpp:2566-2780): mode (unary), partition, rotation, index
    // selector, then R of every endpoint, G, B, A, p-bits, colour indices, alpha indices.
    // Every lane contributes its pixel's index fields AND the endpoint fields hl and hl + 16 (field = channel x endpoint, up to
    // 18 for the three-subset modes); lanes 0..5 add the p-bits, lane 0 the header.
    uint32_t w0[DXB_NL], w1[DXB_NL], w2[DXB_NL], w3[DXB_NL];
    DXB_LANES_BEGIN
        const uint32_t hl = (uint32_t)(lane & 15);
        const uint32_t wMode = W[L].mode, wIdx = W[L].idx;
        const dxb_bc7_modecfg cfg = dxb_bc7_cfg((int)wMode);
        const bool two = (wMode == 1u || wMode == 3u || wMode == 7u);
        const bool three = (wMode == 0u || wMode == 2u);
        const bool sepA = (wMode == 4u || wMode == 5u);
        const uint32_t ibc = (wMode == 4u && wIdx) ? 3u : cfg.ib;
        const uint32_t iba = (wMode == 4u) ? (wIdx ? 2u : 3u) : cfg.ib2;
        const uint32_t nsub = three ? 3u : (two ? 2u : 1u);
        const uint32_t sb = three ? ((dxb_part3[W[L].shape] >> (2 * hl)) & 3u) : (two ? ((dxb_part2[W[L].shape] >> hl) & 1u) : 0u);
        const uint32_t anchor1 = anchor1Src[L], anchor2 = anchor2Src[L];
        const bool flipC0 = ((aC0[L] >> (ibc - 1u)) & 1u) != 0;
        const bool flipC1 = (nsub >= 2u) && (((aC1[L] >> (ibc - 1u)) & 1u) != 0);
        const bool flipC2 = (nsub == 3u) && (((aC2[L] >> (ibc - 1u)) & 1u) != 0);
        const bool flipA = (iba != 0) && (((aA0[L] >> (iba - 1u)) & 1u) != 0);
        uint32_t iC = idxC[L], iA = idxA[L];
        {
            const bool fl = (sb == 2u) ? flipC2 : (sb == 1u) ? flipC1 : flipC0;
            if (fl) iC = ((1u << ibc) - 1u) - iC;
            if (flipA) iA = ((1u << iba) - 1u) - iA;
        }
        const uint32_t partBits = three ? ((wMode == 0u) ? 4u : 6u) : (two ? 6u : 0u);
        const uint32_t rotBits = sepA ? 2u : 0u;
        const uint32_t imBits = (wMode == 4u) ? 1u : 0u;
        const uint32_t hdr = (wMode + 1u) + partBits + rotBits + imBits;
        const uint32_t epBits = nsub * 2u * (3u * cfg.cbits + cfg.abits);
        const uint32_t npb = (cfg.ptype == 1) ? nsub * 2u : (cfg.ptype == 2) ? nsub : 0u;
        const uint32_t idxStart = hdr + epBits + npb;
        // Mode 4: the first index block is always the 2-bit set, the second the 3-bit set (:2727-2757)
        const bool swapSets = (wMode == 4u) && (wIdx != 0u);
        const uint32_t ib1 = swapSets ? iba : ibc;                 // bits of the first index block
        const uint32_t ib2v = swapSets ? ibc : iba;                // bits of the second index block
        const uint32_t secondStart = idxStart + 16u * ib1 - nsub;

        dxb_u128 bits; bits.lo = 0; bits.hi = 0;
        {
            // index fields of pixel hl
            const uint32_t first = swapSets ? iA : iC, second = swapSets ? iC : iA;
            const uint32_t before = (hl > 0 ? 1u : 0u) + ((nsub >= 2u && hl > anchor1) ? 1u : 0u) + ((nsub == 3u && hl > anchor2) ? 1u : 0u);   // anchors before this pixel
            const bool isAnchor = (hl == 0) || (nsub >= 2u && hl == anchor1) || (nsub == 3u && hl == anchor2);
            dxb_put_bits(&bits, idxStart + hl * ib1 - before, isAnchor ? ib1 - 1u : ib1, first);
            dxb_put_bits(&bits, secondStart + (hl ? hl * ib2v - 1u : 0u), ib2v ? (hl ? ib2v : ib2v - 1u) : 0u, second);
        }
        for (uint32_t fi = hl; fi < 2u * nsub * 4u; fi += 16u)
        {
            // endpoint field fi: channel c = fi / (2 nsub), endpoint e = 2 * subset + which.  Colour channels follow the colour
            // flip of their subset; in modes 4/5 the alpha channel has its own index set and follows flipA.
            const uint32_t per = 2u * nsub;
            const uint32_t c = (per == 2u) ? (fi >> 1) : (per == 4u) ? (fi >> 2) : (fi / 6u);
            const uint32_t e = fi - c * per, sub = e >> 1, which = e & 1u;
            const bool fl = (sepA && c == 3u) ? flipA : ((sub == 2u) ? flipC2 : (sub == 1u) ? flipC1 : flipC0);
            const uint32_t qa = (sub == 2u) ? W[L].q0[2] : (sub == 1u) ? W[L].q0[1] : W[L].q0[0];
            const uint32_t qb = (sub == 2u) ? W[L].q1[2] : (sub == 1u) ? W[L].q1[1] : W[L].q1[0];
            const uint32_t field = (((which != 0u) != fl) ? qb : qa) >> (8u * c);
            const uint32_t nb = (c == 3u) ? cfg.abits : cfg.cbits;
            const uint32_t pos = hdr + ((c == 3u) ? 3u * per * cfg.cbits + e * cfg.abits : (c * per + e) * cfg.cbits);
            dxb_put_bits(&bits, pos, nb, field & 0xFFu);
        }
        if (hl < npb)
        {
            // p-bits: unique (ptype 1): endpoint order; shared (ptype 2): one per subset
            const uint32_t sub = (cfg.ptype == 2) ? hl : (hl >> 1), which = (cfg.ptype == 2) ? 0u : (hl & 1u);
            const bool fl = (sub == 2u) ? flipC2 : (sub == 1u) ? flipC1 : flipC0;
            const uint32_t pbv = (sub == 2u) ? W[L].pb[2] : (sub == 1u) ? W[L].pb[1] : W[L].pb[0];
            const uint32_t bit = (cfg.ptype == 2) ? (pbv & 1u) : ((pbv >> (((which != 0u) != fl) ? 1u : 0u)) & 1u);
            dxb_put_bits(&bits, hdr + epBits + hl, 1, bit);
        }
        if (hl == 0)
        {
            dxb_put_bits(&bits, wMode, 1, 1u);
            dxb_put_bits(&bits, wMode + 1u, partBits, W[L].shape);
            dxb_put_bits(&bits, wMode + 1u + partBits, rotBits, W[L].rot);
            dxb_put_bits(&bits, wMode + 1u + partBits + rotBits, imBits, wIdx);
        }
        w0[L] = (uint32_t)bits.lo; w1[L] = (uint32_t)(bits.lo >> 32); w2[L] = (uint32_t)bits.hi; w3[L] = (uint32_t)(bits.hi >> 32);
    DXB_LANES_END
    uint32_t o0[DXB_NL], o1[DXB_NL], o2[DXB_NL], o3[DXB_NL];
    dxb_half_or_u32(w0, o0); dxb_half_or_u32(w1, o1); dxb_half_or_u32(w2, o2); dxb_half_or_u32(w3, o3);
    DXB_LANES_BEGIN
        uint8_t* out = (lane & 16) ? out1 : out0;
        if ((lane & 15) == 0 && out)
        {
            uint32_t* o = (uint32_t*)out;
            o[0] = o0[L]; o[1] = o1[L]; o[2] = o2[L]; o[3] = o3[L];
        }
    DXB_LANES_END
}

#if !DXB_ON_DEVICE
// emulator entry: pxA / pxB = 16 RGBA fp32 pixels each after ConvertScanline (values clamped to [0,1]);
// pxB / outB may be null (odd block count)
static inline void dxb_bc7_encode_pair_emul(const dxb_px* pxA, const dxb_px* pxB, uint32_t bcflags, uint8_t* outA, uint8_t* outB)
{
    static thread_local dxb_bc7_scratch S;
    for (int i = 0; i < 16; ++i)
    {
        S.px[i] = dxb_make_px(dxb_bc7_ldr(pxA[i].x), dxb_bc7_ldr(pxA[i].y), dxb_bc7_ldr(pxA[i].z), dxb_bc7_ldr(pxA[i].w));
        S.px[16 + i] = pxB ? dxb_make_px(dxb_bc7_ldr(pxB[i].x), dxb_bc7_ldr(pxB[i].y), dxb_bc7_ldr(pxB[i].z), dxb_bc7_ldr(pxB[i].w))
                           : dxb_make_px(0.0f, 0.0f, 0.0f, 255.0f);
    }
    dxb_bc7_encode_pair<true>(&S, bcflags, outA, pxB ? outB : nullptr);
}
#endif
