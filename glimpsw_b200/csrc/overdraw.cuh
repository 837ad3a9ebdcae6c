// overdraw.cuh — the OverdrawShader program (ShadingContext::OverdrawShader, Shading.cpp:656) and the colour-word
// debug layers of ShadingContext::ResolveDebug (Shading.cpp:734-773).
//
// FS_Overdraw (Shading.cpp:333-342) runs once per 4x4 fragment a triangle touches (DrawTriangle<>, Rasterizer.h:284),
// on all 16 lanes: covered lanes add 1 to the high u16 of the colour word, helper lanes add 1 to the low u16, both
// saturating; the depth layer takes max(Depth, stored) on all 16 lanes, Depth being the triangle's plane equation
// (also outside the triangle). There is no depth test, so every operation commutes and the result is independent of
// triangle order: counts are accumulated with 64-bit atomic adds (pixel count in the high word, helper count in the
// low word of the framebuffer's key buffer — two 32-bit counters cannot overflow into each other for any scene that
// fits in memory), depth with an unsigned atomicMax on the float bits (Depth <= 0 never beats the stored value,
// which starts >= 0); k_overdraw_finish folds the counters into layer 0 with the reference's u16 saturation.
#pragma once

#include "common.cuh"

namespace swrb {

__global__ void __launch_bounds__(256)
k_overdraw_begin(ulonglong2* __restrict__ counters, uint32_t numVec, DevCtl* __restrict__ ctl) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (gid == 0) { ctl->triCount = 0; ctl->bigCount = 0; ctl->binTotal = 0; ctl->numActiveTiles = 0; ctl->alphaCount = 0; ctl->clipCount = 0; ctl->workCursor = 0; ctl->superTotal = 0; ctl->sparseTiles = 0; ctl->denseTiles = 0; }
    for (uint32_t i = gid; i < 2u * numVec; i += stride) counters[i] = make_ulonglong2(0ull, 0ull);
}

// One warp per triangle record, one 4x4 fragment per half warp and iteration (lane & 15 = pixel inside the
// fragment, Rasterizer.h:247-248).
__global__ void __launch_bounds__(256)
k_raster_overdraw(const TriRecord* __restrict__ tris, FrameParams fp, unsigned long long* __restrict__ counters,
                  uint32_t* __restrict__ depthLayer, DevCtl* __restrict__ ctl) {
    const uint32_t n = ctl->overflow ? 0u : ctl->triCount;
    const uint32_t lane = lane_id(), frag = lane >> 4, i = lane & 15u;
    const uint32_t half = 0xFFFFu << (lane & 16u);
    const uint32_t warpsTotal = gridDim.x * (blockDim.x >> 5);
    for (uint32_t rec = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); rec < n; rec += warpsTotal) {
        const uint4* src = reinterpret_cast<const uint4*>(tris + rec);
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        TriRecord t;
        t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
        t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
        BBox r;
        if (!raster_region(t.pos0, t.pos1, t.pos2, fp, r)) continue;
        Edges e;
        edge_setup(t, fp.halfW, fp.halfH, e);
        // fragments overlapping the region; every fragment with a covered pixel is among them, and all of them lie
        // inside the reference's 4-aligned traversal box
        const int32_t fx0 = r.minX & ~3, fy0 = r.minY & ~3;
        const int32_t nfx = ((r.maxX + 3) >> 2) - (fx0 >> 2), nfy = ((r.maxY + 3) >> 2) - (fy0 >> 2);
        const int32_t total = nfx * nfy;
        for (int32_t f0 = 0; f0 < total; f0 += 2) {
            const int32_t f = f0 + (int32_t)frag;
            const bool active = f < total;
            const int32_t fyI = f / nfx, fxI = f - fyI * nfx;
            const uint32_t x = (uint32_t)(fx0 + fxI * 4) + (i & 3u), y = (uint32_t)(fy0 + fyI * 4) + (i >> 2);
            const uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * x + (uint32_t)e.b12 * y;
            const uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * x + (uint32_t)e.b20 * y;
            const uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * x + (uint32_t)e.b01 * y;
            const bool covered = active && (int32_t)(e0 | e1 | e2) >= 0;                          // Rasterizer.h:281-282
            const uint32_t tileMask = __ballot_sync(0xFFFFFFFFu, covered) & half;
            if (!active || tileMask == 0) continue;                                              // :284
            const uint32_t off = fb_pixel_offset(x, y, fp.width);
            atomicAdd(counters + off, covered ? (1ull << 32) : 1ull);                            // Shading.cpp:336-337
            const float d = pixel_depth(e, (int32_t)e1, (int32_t)e2);                            // Rasterizer.h:296
            if (d > 0.0f) atomicMax(depthLayer + off, __float_as_uint(d));                       // Shading.cpp:341 (NaN fails d > 0)
        }
    }
}

// layer 0 <- saturating u16 add of this draw's counters (_mm512_adds_epu16 composes: sat(sat(a + b) + c) = sat(a + b + c))
__global__ void __launch_bounds__(256)
k_overdraw_finish(const unsigned long long* __restrict__ counters, uint32_t* __restrict__ color, uint32_t numPixels, const DevCtl* __restrict__ ctl) {
    if (ctl->overflow) return;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < numPixels; p += gridDim.x * blockDim.x) {
        const unsigned long long c = counters[p];
        if (c == 0) continue;
        const uint32_t old = color[p];
        const unsigned long long hi = (unsigned long long)(old >> 16) + (c >> 32), lo = (unsigned long long)(old & 0xFFFFu) + (c & 0xFFFFFFFFull);
        color[p] = ((uint32_t)(hi < 0xFFFFull ? hi : 0xFFFFull) << 16) | (uint32_t)(lo < 0xFFFFull ? lo : 0xFFFFull);
    }
}

__constant__ float kTurboCoeffs[18] = {
    0.13572138f, 4.61539260f,  -42.66032258f, 132.13108234f, -152.94239396f, 59.28637943f,
    0.09140261f, 2.19418839f,  4.84296658f,   -14.18503333f, 4.27729857f,    2.82956604f,
    0.10667330f, 12.64194608f, -60.58204836f, 110.36276771f, -89.90310912f,  27.34824973f };

// ---- ResolveDebug layers that only read the colour word (Shading.cpp:755-766) ----------------------------------------
__device__ __forceinline__ float od_turbo(float x, const float* p) {                             // ColormapTurbo, Shading.cpp:249-260
    return __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(x, p[5]), p[4])), p[3])), p[2])), p[1])), p[0]);
}
// RGBA8u::Pack (Texture.h:55-67): round2i(v * 255) (vcvtps2dq: 0x80000000 on overflow / NaN), then the two saturating packs.
// The Turbo polynomial leaves [0, 1] by orders of magnitude for large counts, so the overflow case is real here.
__device__ __forceinline__ uint32_t od_pack(float v) {
    const float s = __fmul_rn(v, 255.0f);
    if (!(s < 2147483648.0f)) return 0u;           // INT_MIN -> packs_epi32 -> -32768 -> packus_epi16 -> 0
    return (uint32_t)min(max(__float2int_rn(s), 0), 255);
}

// layer: 4 MeshletId, 5 TriangleId, 6 OverdrawPixel, 7 OverdrawQuad (enum class DebugLayer, Shading.h:8)
__global__ void __launch_bounds__(256)
k_resolve_debug_ids(uint32_t* __restrict__ color, const uint32_t* __restrict__ depthLayer, uint32_t width, uint32_t height, int layer) {
    const uint32_t x = blockIdx.x * 32u + threadIdx.x, y = blockIdx.y * 8u + threadIdx.y;
    if (x >= width || y >= height) return;
    const uint32_t off = fb_pixel_offset(x, y, width);
    const float depth = __uint_as_float(depthLayer[off]);
    uint32_t outc;
    if (depth <= 0.0f) {
        outc = (((x & ~3u) ^ (y & ~3u)) & 4u) ? 0xFFA0A0A0u : 0xFFFFFFFFu;                       // :768
    } else {
        const uint32_t d = color[off];
        float c[3];
        if (layer <= 5) {
            const uint32_t h = (layer == 4 ? d / SWR_MAX_PRIMS : d) * 123456789u;                // :756, :758
            const float s = 1.0f / 255;                                                          // RGBA8u::Unpack
            c[0] = (float)(h & 255u) * s; c[1] = (float)((h >> 8) & 255u) * s; c[2] = (float)((h >> 16) & 255u) * s;
        } else {
            const float* k = kTurboCoeffs;
            const float countPix = (float)(d >> 16), countFrag = (float)(d & 0xFFFFu);
            const float v = layer == 6 ? __fdiv_rn(countPix, 30.0f) : __fdiv_rn(__fadd_rn(countPix, __fmul_rn(countFrag, 0.5f)), 30.0f);   // :760-765
            c[0] = od_turbo(v, k); c[1] = od_turbo(v, k + 6); c[2] = od_turbo(v, k + 12);
        }
        outc = od_pack(c[0]) | (od_pack(c[1]) << 8) | (od_pack(c[2]) << 16) | 0xFF000000u;
    }
    color[off] = outc;
}

}  // namespace swrb
