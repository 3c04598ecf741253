// dxb_k_bc7.cu — k_compress_bc7: one HALF-WARP per 4x4 block (two blocks per warp), BC7 mode/partition search (dxb_bc7.cuh)
#include "dxb_launch.h"
#include "dxb_bc7.cuh"

template <bool THREE>
__global__ void __launch_bounds__(DXB_BC7_WARPS * 32, DXB_BC7_MINB) k_compress_bc7(const dxb_job* __restrict__ jobs, dxb_job single, dxb_compress_params P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    dxb_bc7_scratch* scratch = (dxb_bc7_scratch*)smem_raw;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, hl = lane & 15u;
    dxb_bc7_scratch* S = &scratch[warp];
    const uint32_t stride = gridDim.x * DXB_BC7_WARPS;
    const uint32_t npairs = (P.totalUnits + 1u) >> 1;
    // every warp of the CTA runs the same number of iterations (a warp without a pair encodes dummy pixels and
    // stores nothing), so CTA-wide barriers inside the encoder are legal
    for (uint32_t base = blockIdx.x * DXB_BC7_WARPS; base < npairs; base += stride)
    {
        // lanes 0-15 stage block 2*pair, lanes 16-31 block 2*pair+1; lane = pixel
        const uint32_t unit = 2u * (base + warp) + (lane >> 4);
        uint8_t* out = nullptr;
        dxb_px ldr = dxb_make_px(0.0f, 0.0f, 0.0f, 255.0f);
        if (unit < P.totalUnits)
        {
            const dxb_job& j = dxb_find_job(jobs, P.njobs, single, unit, P.periodUnits, P.periodJobs);
            const uint32_t local = unit - j.firstUnit;
            const uint32_t by = local / j.nbx, bx = local - by * j.nbx;
            // CompressBC's partial-block replication with source map {0,0,0,1} (DirectXTexCompress.cpp:159-187)
            const uint32_t x0 = bx * 4, y0 = by * 4;
            const uint32_t pw = (j.width - x0 < 4u) ? (j.width - x0) : 4u;
            const uint32_t ph = (j.height - y0 < 4u) ? (j.height - y0) : 4u;
            const uint32_t s = hl & 3u, t = hl >> 2;
            const uint32_t sc = (s < pw) ? s : ((s == 3u && pw > 1u) ? 1u : 0u);
            const uint32_t tr = (t < ph) ? t : ((t == 3u && ph > 1u) ? 1u : 0u);
            dxb_px v = dxb_load_pixel(P.srcFormat, j.src + (size_t)(y0 + tr) * j.srcPitch, x0 + sc);
            v = dxb_convert_pixel(v, P.inF, P.outF, P.cflags);
            ldr = dxb_make_px(dxb_bc7_ldr(v.x), dxb_bc7_ldr(v.y), dxb_bc7_ldr(v.z), dxb_bc7_ldr(v.w));
            out = j.dst + (size_t)by * j.dstPitch + (size_t)bx * 16u;
        }
        S->px[lane] = ldr;
        __syncwarp();
        // the encoder takes one output pointer per half; every lane passes its own half's pointer in both slots
        dxb_bc7_encode_pair<THREE>(S, P.bcflags, out, out);
        __syncwarp();
    }
}

static const size_t kBC7Smem = sizeof(dxb_bc7_scratch) * DXB_BC7_WARPS;
static bool bc7_attr_set()
{
    static const bool ok = (cudaFuncSetAttribute(k_compress_bc7<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7Smem) == cudaSuccess) &&
                           (cudaFuncSetAttribute(k_compress_bc7<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBC7Smem) == cudaSuccess);
    return ok;
}

void dxb_launch_bc7(unsigned grid, cudaStream_t stream, const dxb_job* jobs, const dxb_job& single, const dxb_compress_params& P)
{
    bc7_attr_set();
    // the three-subset pass (a non-default flag) lives in its own instantiation
    if (P.bcflags & DXB_BC_FLAGS_USE_3SUBSETS) k_compress_bc7<true><<<grid, DXB_BC7_WARPS * 32, kBC7Smem, stream>>>(jobs, single, P);
    else k_compress_bc7<false><<<grid, DXB_BC7_WARPS * 32, kBC7Smem, stream>>>(jobs, single, P);
}
int dxb_occupancy_bc7()
{
    int b = 0;
    if (!bc7_attr_set() || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_compress_bc7<false>, DXB_BC7_WARPS * 32, kBC7Smem) != cudaSuccess) { (void)cudaGetLastError(); b = 1; }
    return b > 0 ? b : 1;
}
