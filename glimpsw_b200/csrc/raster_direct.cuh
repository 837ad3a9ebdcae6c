// raster_direct.cuh — the unbinned path (Rasterizer::EnableBinning == false).
//
// Replaces Rasterizer::DrawMeshletsST's per-triangle work (Rasterizer.cpp:186-206): late setup
// (TriangleEdgeVars::Setup :296-329), DrawTriangle<> traversal (Rasterizer.h:250-328) and the
// FS_EncodeSurfaceId<false> fragment program (Shading.cpp:309-331). Instead of serialising the depth
// test per pixel it reduces 64-bit depth|id keys with atomicMax in an HBM/L2-resident key buffer
// (REDG.MAX.64), which gives the single-worker reference order independent of scheduling.
//
//   k_raster_direct : one thread per triangle record; triangles whose pixel region is small are
//                     rasterized by that thread, the rest are split into 128x128-px work items.
//   k_raster_big    : one warp per (triangle, 128x128 bin) work item: a coarse pass tests 32 8x4-px
//                     blocks at once (one per lane, trivial reject against each edge), then the warp
//                     visits surviving blocks with one pixel per lane.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kBigBinShift = 7;   // big-triangle work items are 128 x 128 px

struct BigItem { uint32_t tri; uint32_t bin; };   // bin = x | y << 16 in 128-px units

__global__ void __launch_bounds__(256)
k_raster_direct(const TriRecord* __restrict__ tris, FrameParams fp, unsigned long long* __restrict__ keys,
                BigItem* __restrict__ bigItems, uint32_t bigCapacity, DevCtl* __restrict__ ctl) {
    const uint32_t n = ctl->overflow ? 0u : ctl->triCount;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(tris + i);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        TriRecord t;
        t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
        t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;

        BBox r;
        if (!raster_region(t.pos0, t.pos1, t.pos2, fp, r)) continue;
        int32_t w = r.maxX - r.minX, h = r.maxY - r.minY;
        if (w * h > kMaxDirectSmallArea) {
            // split into 128x128 work items for the warp-cooperative kernel
            int32_t bx0 = r.minX >> kBigBinShift, bx1 = (r.maxX - 1) >> kBigBinShift;
            int32_t by0 = r.minY >> kBigBinShift, by1 = (r.maxY - 1) >> kBigBinShift;
            uint32_t cnt = (uint32_t)((bx1 - bx0 + 1) * (by1 - by0 + 1));
            uint32_t base = atomicAdd(&ctl->bigCount, cnt);
            if (base + cnt > bigCapacity) { atomicExch(&ctl->overflow, 2u); continue; }
            for (int32_t by = by0; by <= by1; by++)
                for (int32_t bx = bx0; bx <= bx1; bx++) bigItems[base++] = BigItem{ i, (uint32_t)bx | ((uint32_t)by << 16) };
            continue;
        }
        Edges e;
        edge_setup(t, fp.halfW, fp.halfH, e);
        uint32_t rowE0 = (uint32_t)e.e0 + (uint32_t)e.a12 * (uint32_t)r.minX + (uint32_t)e.b12 * (uint32_t)r.minY;
        uint32_t rowE1 = (uint32_t)e.e1 + (uint32_t)e.a20 * (uint32_t)r.minX + (uint32_t)e.b20 * (uint32_t)r.minY;
        uint32_t rowE2 = (uint32_t)e.e2 + (uint32_t)e.a01 * (uint32_t)r.minX + (uint32_t)e.b01 * (uint32_t)r.minY;
        for (int32_t y = r.minY; y < r.maxY; y++) {
            uint32_t e0 = rowE0, e1 = rowE1, e2 = rowE2;
            for (int32_t x = r.minX; x < r.maxX; x++) {
                if ((int32_t)(e0 | e1 | e2) >= 0) {                         // Rasterizer.h:289-290
                    float d = pixel_depth(e, (int32_t)e1, (int32_t)e2);
                    if (d > 0.0f) key_max(keys, fb_pixel_offset((uint32_t)x, (uint32_t)y, fp.width), make_key(d, t.id));
                }
                e0 += (uint32_t)e.a12; e1 += (uint32_t)e.a20; e2 += (uint32_t)e.a01;
            }
            rowE0 += (uint32_t)e.b12; rowE1 += (uint32_t)e.b20; rowE2 += (uint32_t)e.b01;
        }
    }
}

__global__ void __launch_bounds__(256)
k_raster_big(const TriRecord* __restrict__ tris, const BigItem* __restrict__ items, FrameParams fp,
             unsigned long long* __restrict__ keys, DevCtl* __restrict__ ctl) {
    const uint32_t n = ctl->overflow ? 0u : ctl->bigCount;
    const uint32_t lane = lane_id();
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t it = warp; it < n; it += warps) {
        BigItem item = items[it];
        const uint4* src = reinterpret_cast<const uint4*>(tris + item.tri);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        TriRecord t;
        t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
        t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
        BBox r;
        raster_region(t.pos0, t.pos1, t.pos2, fp, r);
        int32_t binX = (int32_t)(item.bin & 0xFFFFu) << kBigBinShift, binY = (int32_t)(item.bin >> 16) << kBigBinShift;
        r.minX = max(r.minX, binX); r.minY = max(r.minY, binY);
        r.maxX = min(r.maxX, binX + (1 << kBigBinShift)); r.maxY = min(r.maxY, binY + (1 << kBigBinShift));
        Edges e;
        edge_setup(t, fp.halfW, fp.halfH, e);
        // Can the int32 edge values wrap for this triangle? Then coarse rejection is not sound and
        // every block is visited (bit-exact with the reference's wrapped arithmetic).
        const bool mayWrap = !edges_wrap_free(t, e, fp);

        warp_raster_region<true>(t, e, r, mayWrap, fp, keys);
    }
}

}  // namespace swrb
