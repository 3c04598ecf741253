// dxb_mipjob.h — plain structs describing one mip-level job (host + device)
#pragma once
#include <stdint.h>
#include <stddef.h>

struct dxb_mip_job
{
    const uint8_t* src; uint8_t* dst;
    size_t srcPitch, dstPitch;
    uint32_t sw, sh, dw, dh;          // source / destination size
    uint32_t firstUnit;               // prefix sum of destination pixels over the batch
    const uint8_t* stale; size_t stalePitch;   // box filter only: see dxb_mip_box
};

// triangle filter gather lists for one axis (CSR): contributions of destination index d are
// entries [off[d], off[d+1]) with ascending source index
struct dxb_tri_axis { const uint32_t* off; const uint32_t* src; const float* w; };
