// dxb_k_decode.cu — k_decompress: one THREAD per 4x4 block: decode (dxb_decode.cuh) -> ConvertScanline -> StoreScanline,
// i.e. DecompressBC (DirectXTexCompress.cpp:425-535) for a batch of images.
#include "dxb_launch.h"
#include "dxb_decode.cuh"

__global__ void __launch_bounds__(128) k_decompress(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t bs = dxb_bc_block_bytes(P.srcFormat);
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
        dxb_decode_block(P.srcFormat, blk, px);
        const uint32_t x0 = bx * 4, y0 = by * 4;
        const uint32_t pw = (j.width - x0 < 4u) ? (j.width - x0) : 4u;
        const uint32_t ph = (j.height - y0 < 4u) ? (j.height - y0) : 4u;
        for (uint32_t t = 0; t < ph; ++t)
        {
            uint8_t* row = j.dst + (size_t)(y0 + t) * j.dstPitch;
            for (uint32_t s = 0; s < pw; ++s)
                dxb_store_pixel(P.dstFormat, row, x0 + s, dxb_convert_pixel(px[(t << 2) | s], P.inF, P.outF, P.cflags));
        }
    }
}

void dxb_launch_decompress(unsigned grid, cudaStream_t stream, const dxb_job* jobs, const dxb_job& single, const dxb_compress_params& P)
{
    k_decompress<<<grid, 128, 0, stream>>>(jobs, single, P);
}
