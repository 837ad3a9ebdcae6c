// cull.cuh — ShadingContext::CullMeshlets, frustum part (Shading.cpp:795-809, :865-867).
//
// One thread per meshlet reads the 16-byte bounding sphere at the head of the 1728-byte Meshlet
// (one 32-byte sector per meshlet), tests it against the five normalised Gribb-Hartmann planes and
// the warp assembles the reference's bitmap (1 bit per meshlet, u16 words) with a ballot.
#pragma once

#include "common.cuh"

namespace swrb {

struct CullPlanes { float p[5][4]; };

__global__ void __launch_bounds__(256)
k_cull_meshlets(const swr_meshlet* __restrict__ meshlets, uint32_t count, CullPlanes planes,
                uint32_t* __restrict__ bitmap32, uint32_t* __restrict__ visibleCount) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool visible = false;
    if (i < count) {
        float4 s = __ldg(reinterpret_cast<const float4*>(meshlets + i));
        visible = true;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            // simd::dot(center, plane.xyz) + plane.w   (SIMD.h:437, Shading.cpp:807)
            float dist = __fadd_rn(__fmaf_rn(s.x, planes.p[k][0], __fmaf_rn(s.y, planes.p[k][1], __fmul_rn(s.z, planes.p[k][2]))), planes.p[k][3]);
            visible = visible && (dist > -s.w);
        }
    }
    uint32_t bits = __ballot_sync(0xFFFFFFFFu, visible);
    if ((threadIdx.x & 31u) == 0 && i < count) {
        bitmap32[i >> 5] = bits;     // two consecutive u16 words of the reference bitmap (little endian)
        if (bits) atomicAdd(visibleCount, (uint32_t)__popc(bits));
    }
}

}  // namespace swrb
