// unpack.cuh — decode of the packed meshlet format (swr_meshlet_packed, include/swr_types.h) into the reference's Meshlet layout.
//
// The reference streams 1,728-byte meshlets and its author notes the loads of ShadeMeshlet as the place for "meshlet
// compression" (Shading.cpp:292-294). Here compression is an import-time / transport format: 16-bit positions inside the
// meshlet's bounding box halve the position bytes (1,376 bytes per meshlet over PCIe and on disk), and this kernel expands
// them at upload — one block per meshlet, HBM to HBM at copy speed — so the per-frame kernels keep one meshlet layout.
// (Decoding inside the mesh kernel instead would save it 352 of its 1,216 bytes per meshlet, but that kernel runs at 0.10 of
// the HBM roof and is bound by its instruction count: nothing to gain there, profiles/r02_summary.md.)
// The decode is one fused multiply-add per coordinate, fmaf((float)Q, Scale, Origin), bit-identical with orc_unpack_meshlets.
#pragma once

#include "common.cuh"

namespace swrb {

__global__ void __launch_bounds__(128)
k_unpack_meshlets(const swr_meshlet_packed* __restrict__ src, swr_meshlet* __restrict__ dst, uint32_t count) {
    const uint32_t t = threadIdx.x;
    for (uint32_t m = blockIdx.x; m < count; m += gridDim.x) {
        const swr_meshlet_packed* s = src + m;
        uint32_t* out = reinterpret_cast<uint32_t*>(dst + m);
        const uint32_t* in = reinterpret_cast<const uint32_t*>(s);
        if (t < 16) out[t] = in[t];                                                      // header, 64 bytes verbatim
        if (t < 96) {                                                                    // 3 x 64 positions, two per thread
            const uint32_t axis = t >> 5, pair = t & 31u;
            const uint32_t q2 = *reinterpret_cast<const uint32_t*>(&s->Q[axis][2 * pair]);
            const float scale = s->Scale[axis], origin = s->Origin[axis];
            float2 p;
            p.x = __fmaf_rn((float)(q2 & 0xFFFFu), scale, origin);
            p.y = __fmaf_rn((float)(q2 >> 16), scale, origin);
            *reinterpret_cast<float2*>(&dst[m].Positions[axis][2 * pair]) = p;
        }
        // TexCoords (256 B) + NormalTangents (256 B) + Indices (384 B) = 224 words, contiguous in both layouts
        const uint32_t* tail = reinterpret_cast<const uint32_t*>(s->TexCoords);
        uint32_t* tailOut = reinterpret_cast<uint32_t*>(dst[m].TexCoords);
        for (uint32_t w = t; w < 224u; w += 128u) tailOut[w] = tail[w];
    }
}

}  // namespace swrb
