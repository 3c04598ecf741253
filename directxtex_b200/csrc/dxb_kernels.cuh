// dxb_kernels.cuh — __global__ kernels of libdxtex_b200 (sm_100a).
//   k_compress_bc15   one THREAD per 4x4 block   BC1/2/3/4/5, bit-exact fp32 restatement (dxb_bc15.cuh)
//   k_compress_bc7    one WARP per 4x4 block     BC7 mode/partition search (dxb_bc7.cuh)
//   k_convert         one thread per pixel       Load -> Convert -> Store (dxb_pixel.cuh)
//   k_mip_level       one thread per dest pixel  one mip level of a batch of images (dxb_mips.cuh)
// A launch covers a whole batch: jobs[] describes the images, `firstUnit` is the prefix sum of work
// units (blocks or pixels); a unit finds its job by binary search.  Single-image calls pass the job
// by value (no device-side descriptor).
#pragma once
#include "dxb_bc15.cuh"
#include "dxb_bc7.cuh"
#include "dxb_mips.cuh"

struct dxb_job
{
    const uint8_t* src; uint8_t* dst;
    size_t srcPitch, dstPitch;
    uint32_t width, height;
    uint32_t nbx, nby;
    uint32_t firstUnit;
    uint32_t pad;
};

template <typename J>
__device__ __forceinline__ const J& dxb_find_job(const J* jobs, uint32_t njobs, const J& single, uint32_t unit)
{
    if (jobs == nullptr) return single;
    uint32_t lo = 0, hi = njobs;            // last job with firstUnit <= unit
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (jobs[mid].firstUnit <= unit) lo = mid; else hi = mid;
    }
    return jobs[lo];
}

struct dxb_compress_params
{
    uint32_t srcFormat, dstFormat;
    uint32_t inF, outF, cflags, bcflags;
    float threshold;
    uint32_t totalUnits, njobs;
};

// ------------------------------------------------------------------------------------------------
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

// ------------------------------------------------------------------------------------------------
#define DXB_BC7_WARPS 8
__global__ void __launch_bounds__(DXB_BC7_WARPS * 32) k_compress_bc7(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    __shared__ dxb_px spx[DXB_BC7_WARPS][16];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * DXB_BC7_WARPS;
    for (uint32_t unit = blockIdx.x * DXB_BC7_WARPS + warp; unit < P.totalUnits; unit += stride)
    {
        const dxb_job& j = dxb_find_job(jobs, P.njobs, single, unit);
        const uint32_t local = unit - j.firstUnit;
        const uint32_t by = local / j.nbx, bx = local - by * j.nbx;
        if (lane < 16)
        {
            // CompressBC's partial-block replication with source map {0,0,0,1} (DirectXTexCompress.cpp:159-187)
            const uint32_t x0 = bx * 4, y0 = by * 4;
            const uint32_t pw = (j.width - x0 < 4u) ? (j.width - x0) : 4u;
            const uint32_t ph = (j.height - y0 < 4u) ? (j.height - y0) : 4u;
            const uint32_t s = lane & 3u, t = lane >> 2;
            const uint32_t sc = (s < pw) ? s : ((s == 3u && pw > 1u) ? 1u : 0u);
            const uint32_t tr = (t < ph) ? t : ((t == 3u && ph > 1u) ? 1u : 0u);
            dxb_px v = dxb_load_pixel(P.srcFormat, j.src + (size_t)(y0 + tr) * j.srcPitch, x0 + sc);
            v = dxb_convert_pixel(v, P.inF, P.outF, P.cflags);
            spx[warp][lane] = dxb_make_px(dxb_bc7_ldr(v.x), dxb_bc7_ldr(v.y), dxb_bc7_ldr(v.z), dxb_bc7_ldr(v.w));
        }
        __syncwarp();
        uint8_t* out = j.dst + (size_t)by * j.dstPitch + (size_t)bx * 16u;
        dxb_bc7_encode_warp(spx[warp], P.bcflags, out);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
struct dxb_convert_params
{
    uint32_t srcFormat, dstFormat, inF, outF, flags;
    uint32_t totalUnits, njobs;
};

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

// ------------------------------------------------------------------------------------------------
struct dxb_mip_params
{
    uint32_t format, mode /*DXB_FILTER_* mode bits*/, filter, lflags;
    uint32_t totalUnits, njobs;
    dxb_tri_axis triX, triY;       // triangle filter only (single job per launch)
};

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
