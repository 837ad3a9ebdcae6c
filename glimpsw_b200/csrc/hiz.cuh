// hiz.cuh — HiZ occlusion culling: depth pyramid + sphere test.
//
//   k_hiz_level          texutil::DownsampleDepth (ImageHelpers.cpp:150-247): the framebuffer's depth layer ->
//                        half-resolution R32f min pyramid in the reference's TiledY8 texture layout. Level m texel
//                        = min of depth over its 2^(m+1) x 2^(m+1) pixel footprint, FLT_MAX outside the frame.
//                        One launch per level (level 0 reads the 4x4-tiled depth layer, level m reads level m-1);
//                        only the 8x8-texel blocks the reference's recursion writes are written.
//   k_cull_meshlets_hiz  ShadingContext::CullMeshlets incl. ProjectSphere and the pyramid taps
//                        (Shading.cpp:264-279, :775-869), one thread per meshlet, ballot -> u16 bitmap.
// Canonical arithmetic like oracle_hiz.cpp (approx_rcp -> 1/x, approx_sqrt(x) -> (1/sqrt(x))*x, all _rn), so the
// visibility bitmap is bit-exact against the oracle.
#pragma once

#include <cfloat>

#include "common.cuh"
#include "cull.cuh"

namespace swrb {

struct HizDesc {             // Texture2D<R32f, TiledY8> (Texture.h:314-329)
    float* data;
    uint32_t width, height, mipLevels, rowShift;
    uint32_t mipOffsets[16];
};

__device__ __forceinline__ uint32_t hiz_texel_offset(uint32_t x, uint32_t y, uint32_t stride) {   // Texture.h:494-501
    return (y & 7u) | (x << 3) | ((y & ~7u) << stride);
}

// Level `m` of the pyramid. texelsX/Y = extent of the written region (whole 8x8 blocks, or the 4x4 top tile).
__global__ void __launch_bounds__(256)
k_hiz_level(const float* __restrict__ depthLayer, HizDesc hz, uint32_t m, uint32_t fbWidth, uint32_t fbHeight,
            uint32_t texelsX, uint32_t texelsY) {
    const uint32_t X = blockIdx.x * 32u + (threadIdx.x & 31u), Y = blockIdx.y * 8u + (threadIdx.x >> 5);
    if (X >= texelsX || Y >= texelsY) return;
    float v = FLT_MAX;
    if (m == 0) {
        const uint32_t px = X * 2u, py = Y * 2u;
        if (px < fbWidth && py < fbHeight) {      // width/height are multiples of 4: the 2x2 footprint is all in or all out
            const float2 a = *reinterpret_cast<const float2*>(depthLayer + fb_pixel_offset(px, py, fbWidth));
            const float2 b = *reinterpret_cast<const float2*>(depthLayer + fb_pixel_offset(px, py + 1u, fbWidth));
            v = fminf(fminf(a.x, a.y), fminf(b.x, b.y));
        }
    } else {
        const float* src = hz.data + hz.mipOffsets[m - 1];
        const uint32_t sstride = hz.rowShift - (m - 1), texel = 1u << m;      // footprint of a level m-1 texel in pixels
#pragma unroll
        for (uint32_t dy = 0; dy < 2; dy++)
#pragma unroll
            for (uint32_t dx = 0; dx < 2; dx++) {
                const uint32_t sx = X * 2u + dx, sy = Y * 2u + dy;
                if (sx * texel < fbWidth && sy * texel < fbHeight) v = fminf(v, src[hiz_texel_offset(sx, sy, sstride)]);
            }
    }
    hz.data[hz.mipOffsets[m] + hiz_texel_offset(X, Y, hz.rowShift - m)] = v;
}

struct CullHizParams {
    float planes[5][4];
    float objectToPrevView[16];
    float scale;              // length(vec3(modelMat[0]))           Shading.cpp:794
    float znear, p00, p11;    // projMat[3][2], [0][0], [1][1]        :795
    float frameW, frameH;
    int useHiz;
};

__global__ void __launch_bounds__(256)
k_cull_meshlets_hiz(const swr_meshlet* __restrict__ meshlets, uint32_t count, CullHizParams cp, HizDesc hz,
                    uint32_t* __restrict__ bitmap32, uint32_t* __restrict__ visibleCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool visible = false;
    if (i < count) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(meshlets + i));
        visible = true;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            float dist = __fadd_rn(__fmaf_rn(s.x, cp.planes[k][0], __fmaf_rn(s.y, cp.planes[k][1], __fmul_rn(s.z, cp.planes[k][2]))), cp.planes[k][3]);
            visible = visible && (dist > -s.w);
        }
        if (visible && cp.useHiz) {
            const float* M = cp.objectToPrevView;
            const float vx_ = __fmaf_rn(s.x, M[0], __fmaf_rn(s.y, M[4], __fmaf_rn(s.z, M[8], M[12])));
            const float vy_ = __fmaf_rn(s.x, M[1], __fmaf_rn(s.y, M[5], __fmaf_rn(s.z, M[9], M[13])));
            const float vz_ = __fmaf_rn(s.x, M[2], __fmaf_rn(s.y, M[6], __fmaf_rn(s.z, M[10], M[14])));
            const float rv = __fmul_rn(s.w, cp.scale);
            // ProjectSphere (Shading.cpp:264-279)
            const float cx = vx_, cy = vy_, cz = __fmul_rn(vz_, -1.0f);
            if (cz >= __fadd_rn(rv, cp.znear)) {
                const float crx = __fmul_rn(cx, rv), cry = __fmul_rn(cy, rv), crz = __fmul_rn(cz, rv);
                const float tx = __fsub_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cz, cz)), __fmul_rn(rv, rv));
                const float ty = __fsub_rn(__fadd_rn(__fmul_rn(cy, cy), __fmul_rn(cz, cz)), __fmul_rn(rv, rv));
                const float vx = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(tx)), tx), vy = __fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(ty)), ty);
                const float hx = __fmul_rn(cp.p00, 0.5f), hy = __fmul_rn(cp.p11, 0.5f);
                const float bbx = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(vx, cx), crz), __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(vx, cz), crx))), hx), 0.5f);
                const float bby = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(vy, cy), crz), __fdiv_rn(1.0f, __fsub_rn(__fmul_rn(vy, cz), cry))), hy), 0.5f);
                const float bbz = __fadd_rn(__fmul_rn(__fmul_rn(__fadd_rn(__fmul_rn(vx, cx), crz), __fdiv_rn(1.0f, __fsub_rn(__fmul_rn(vx, cz), crx))), hx), 0.5f);
                const float bbw = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(vy, cy), crz), __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(vy, cz), cry))), hy), 0.5f);
                const float sizeX = __fmul_rn(__fsub_rn(bbz, bbx), cp.frameW), sizeY = __fmul_rn(__fsub_rn(bbw, bby), cp.frameH);
                int32_t mip = ((int32_t)__float_as_uint(fmaxf(sizeX, sizeY)) - (127 << 23)) >> 23;     // ilog2 (:821)
                mip = max(1, min(mip, (int32_t)hz.mipLevels - 1));                                      // :822
                int32_t x0 = max(__float2int_rz(__fmul_rn(bbx, cp.frameW)), 0) >> mip, y0 = max(__float2int_rz(__fmul_rn(bby, cp.frameH)), 0) >> mip;
                int32_t x1 = min(__float2int_rz(__fmul_rn(bbz, cp.frameW)), (int32_t)cp.frameW - 1) >> mip;
                int32_t y1 = min(__float2int_rz(__fmul_rn(bbw, cp.frameH)), (int32_t)cp.frameH - 1) >> mip;
                const float depthSphere = __fdiv_rn(cp.znear, __fsub_rn(-vz_, rv));                     // :829
                float depthVisible = FLT_MAX;
                const float* lvl = hz.data + hz.mipOffsets[mip - 1];
                const uint32_t stride = hz.rowShift - (uint32_t)(mip - 1);
                for (int32_t y = y0; y <= y1; y++)
                    for (int32_t x = x0; x <= x1; x++) depthVisible = fminf(depthVisible, __ldg(lvl + hiz_texel_offset((uint32_t)x, (uint32_t)y, stride)));
                visible = depthSphere > depthVisible;                                                   // :839
            }
        }
    }
    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, visible);
    if ((threadIdx.x & 31u) == 0 && i < count) {
        bitmap32[i >> 5] = bits;
        if (bits) atomicAdd(visibleCount, (uint32_t)__popc(bits));
    }
}

}  // namespace swrb
