// dxb_k_bc15.cu — k_compress_bc15: one THREAD per 4x4 block, BC1/2/3/4/5 (bit-exact fp32 restatement, dxb_bc15.cuh)
#include "dxb_launch.h"
#include "dxb_bc15.cuh"

__global__ void __launch_bounds__(128) k_compress_bc15(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x; unit < P.totalUnits; unit += stride)
    {
        const dxb_job& j = dxb_find_job(jobs, P.njobs, single, unit);
        const uint32_t local = unit - j.firstUnit;
        const uint32_t by = local / j.nbx, bx = local - by * j.nbx;
        dxb_image_desc img; img.pixels = j.src; img.rowPitch = j.srcPitch; img.width = j.width; img.height = j.height; img.format = P.srcFormat;
        dxb_px px[16];
        dxb_gather_block(img, bx, by, P.inF, P.outF, P.cflags, px);
        const uint32_t bs = dxb_bc_block_bytes(P.dstFormat);
        uint8_t* out = j.dst + (size_t)by * j.dstPitch + (size_t)bx * bs;
        __align__(16) uint8_t blk[16];
        dxb_encode_block_bc15(P.dstFormat, px, P.bcflags, P.threshold, blk);
        if (bs == 8) *reinterpret_cast<uint2*>(out) = *reinterpret_cast<const uint2*>(blk);
        else *reinterpret_cast<uint4*>(out) = *reinterpret_cast<const uint4*>(blk);
    }
}


void dxb_launch_bc15(unsigned grid, cudaStream_t stream, const dxb_job* jobs, const dxb_job& single, const dxb_compress_params& P)
{
    k_compress_bc15<<<grid, 128, 0, stream>>>(jobs, single, P);
}
int dxb_occupancy_bc15()
{
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_compress_bc15, 128, 0) != cudaSuccess) { (void)cudaGetLastError(); b = 1; }
    return b > 0 ? b : 1;
}
