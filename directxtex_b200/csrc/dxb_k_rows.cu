// dxb_k_rows.cu — row kernels: k_convert (Load -> Convert -> Store per pixel) and k_mip_level (one mip level of a batch)
#include "dxb_launch.h"
#include "dxb_pixel.cuh"
#include "dxb_mips.cuh"

__global__ void __launch_bounds__(256) k_convert(const dxb_job* __restrict__ jobs, dxb_job single, dxb_convert_params P)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x; unit < P.totalUnits; unit += stride)
    {
        const dxb_job& j = dxb_find_job(jobs, P.njobs, single, unit);
        const uint32_t local = unit - j.firstUnit;
        const uint32_t y = local / j.width, x = local - y * j.width;
        dxb_px v = dxb_load_pixel(P.srcFormat, j.src + (size_t)y * j.srcPitch, x);
        v = dxb_convert_pixel(v, P.inF, P.outF, P.flags);
        dxb_store_pixel(P.dstFormat, j.dst + (size_t)y * j.dstPitch, x, v);
    }
}


__global__ void __launch_bounds__(256) k_mip_level(const dxb_mip_job* __restrict__ jobs, dxb_mip_job single, dxb_mip_params P)
{
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t unit = blockIdx.x * blockDim.x + threadIdx.x; unit < P.totalUnits; unit += stride)
    {
        const dxb_mip_job& j = dxb_find_job(jobs, P.njobs, single, unit);
        const uint32_t local = unit - j.firstUnit;
        const uint32_t y = local / j.dw, x = local - y * j.dw;
        dxb_px v;
        switch (P.mode)
        {
        case DXB_FILTER_POINT:
            v = dxb_mip_point(P.format, j, x, y);
            dxb_store_pixel(P.format, j.dst + (size_t)y * j.dstPitch, x, v);
            continue;
        case DXB_FILTER_BOX: v = dxb_mip_box(P.format, j, x, y, P.lflags); break;
        case DXB_FILTER_LINEAR: v = dxb_mip_linear(P.format, j, x, y, P.filter, P.lflags); break;
        case DXB_FILTER_CUBIC: v = dxb_mip_cubic(P.format, j, x, y, P.filter, P.lflags); break;
        default: v = dxb_mip_triangle(P.format, j, x, y, P.lflags, P.triX, P.triY); break;
        }
        dxb_store_linear(P.format, j.dst, j.dstPitch, x, y, v, P.lflags);
    }
}

void dxb_launch_convert(unsigned grid, cudaStream_t stream, const dxb_job* jobs, const dxb_job& single, const dxb_convert_params& P)
{
    k_convert<<<grid, 256, 0, stream>>>(jobs, single, P);
}
void dxb_launch_mip(unsigned grid, cudaStream_t stream, const dxb_mip_job* jobs, const dxb_mip_job& single, const dxb_mip_params& P)
{
    k_mip_level<<<grid, 256, 0, stream>>>(jobs, single, P);
}
