// dxb_k_decode.cu — DecompressBC (DirectXTexCompress.cpp:425-535) for a batch of images: one THREAD per 4x4 block:
// decode (dxb_decode.cuh) -> ConvertScanline -> StoreScanline.
//   k_decompress            generic: any BC source, any implemented target format
//   k_decompress_t<SF,DF>   the default (source, target) pairs with no sRGB step: compile-time formats (one decoder, one
//                           store path per kernel) and one vector store per block row
#include "dxb_launch.h"
#include "dxb_decode.cuh"

template <bool GENERIC, uint32_t SF, uint32_t DF>
__device__ __forceinline__ void decode_body(const dxb_job* __restrict__ jobs, const dxb_job& single, const dxb_compress_params& P)
{
    const uint32_t srcFormat = GENERIC ? P.srcFormat : SF, dstFormat = GENERIC ? P.dstFormat : DF;
    const uint32_t inF = GENERIC ? P.inF : dxb_convert_flags(SF), outF = GENERIC ? P.outF : dxb_convert_flags(DF);
    const uint32_t cflags = GENERIC ? P.cflags : 0u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t bs = dxb_bc_block_bytes(srcFormat);
    for (uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x; unit < P.totalUnits; unit += stride)
    {
        const dxb_job& j = dxb_find_job(jobs, P.njobs, single, unit);
        const uint32_t local = unit - j.firstUnit;
        const uint32_t by = local / j.nbx, bx = local - by * j.nbx;
        __align__(16) uint8_t blk[16];
        const uint8_t* src = j.src + (size_t)by * j.srcPitch + (size_t)bx * bs;
        if (bs == 8) *reinterpret_cast<uint2*>(blk) = *reinterpret_cast<const uint2*>(src);
        else *reinterpret_cast<uint4*>(blk) = *reinterpret_cast<const uint4*>(src);
        dxb_px px[16];
        dxb_decode_block(srcFormat, blk, px);
        const uint32_t x0 = bx * 4, y0 = by * 4;
        const uint32_t pw = (j.width - x0 < 4u) ? (j.width - x0) : 4u;
        const uint32_t ph = (j.height - y0 < 4u) ? (j.height - y0) : 4u;
        if (!GENERIC)
        {
            constexpr uint32_t B = dxb_bytes_per_pixel(GENERIC ? 2u : DF), ROWB = 4u * B, V = (ROWB >= 16u) ? 16u : ROWB;
            uint8_t* d0 = j.dst + (size_t)y0 * j.dstPitch + (size_t)x0 * B;
            if (pw == 4u && ph == 4u && ((((uintptr_t)d0 | j.dstPitch) & (V - 1u)) == 0u))
            {
                #pragma unroll
                for (uint32_t t = 0; t < 4; ++t)
                {
                    __align__(16) uint8_t row[ROWB];
                    #pragma unroll
                    for (uint32_t s2 = 0; s2 < 4; ++s2) dxb_store_pixel(dstFormat, row, s2, dxb_convert_pixel(px[(t << 2) | s2], inF, outF, cflags));
                    uint8_t* d = d0 + (size_t)t * j.dstPitch;
                    #pragma unroll
                    for (uint32_t k = 0; k < ROWB; k += V)
                    {
                        if (V == 16u) *reinterpret_cast<uint4*>(d + k) = *reinterpret_cast<const uint4*>(row + k);
                        else if (V == 8u) *reinterpret_cast<uint2*>(d + k) = *reinterpret_cast<const uint2*>(row + k);
                        else *reinterpret_cast<uint32_t*>(d + k) = *reinterpret_cast<const uint32_t*>(row + k);
                    }
                }
                continue;
            }
        }
        for (uint32_t t = 0; t < ph; ++t)
        {
            uint8_t* row = j.dst + (size_t)(y0 + t) * j.dstPitch;
            for (uint32_t s2 = 0; s2 < pw; ++s2)
                dxb_store_pixel(dstFormat, row, x0 + s2, dxb_convert_pixel(px[(t << 2) | s2], inF, outF, cflags));
        }
    }
}

__global__ void __launch_bounds__(128) k_decompress(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    decode_body<true, 0, 0>(jobs, single, P);
}
template <uint32_t SF, uint32_t DF>
__global__ void __launch_bounds__(128) k_decompress_t(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    decode_body<false, SF, DF>(jobs, single, P);
}

// default targets (DirectXTexCompress.cpp:552-579): BC1/2/3/7 -> RGBA8, BC4 -> R8, BC5 -> R8G8, BC6H -> RGBA32F
#define DXB_DEC_PAIRS(X) X(71, 28) X(74, 28) X(77, 28) X(98, 28) X(80, 61) X(81, 63) X(83, 49) X(84, 51) X(95, 2) X(96, 2)

void dxb_launch_decompress(unsigned grid, cudaStream_t stream, const dxb_job* jobs, const dxb_job& single, const dxb_compress_params& P)
{
    uint32_t sf = P.srcFormat, df = P.dstFormat;
    if (sf == DXB_FMT_BC1_UNORM_SRGB) sf = DXB_FMT_BC1_UNORM;
    if (sf == DXB_FMT_BC2_UNORM_SRGB) sf = DXB_FMT_BC2_UNORM;
    if (sf == DXB_FMT_BC3_UNORM_SRGB) sf = DXB_FMT_BC3_UNORM;
    if (sf == DXB_FMT_BC7_UNORM_SRGB) sf = DXB_FMT_BC7_UNORM;
    if (df == DXB_FMT_R8G8B8A8_UNORM_SRGB) df = DXB_FMT_R8G8B8A8_UNORM;
#ifndef DXB_DEC_GENERIC_ONLY
    if (P.cflags == 0)
    {
#define DXB_X(SF, DF) if (sf == SF && df == DF) { k_decompress_t<SF, DF><<<grid, 128, 0, stream>>>(jobs, single, P); return; }
        DXB_DEC_PAIRS(DXB_X)
#undef DXB_X
    }
#endif
    k_decompress<<<grid, 128, 0, stream>>>(jobs, single, P);
}
