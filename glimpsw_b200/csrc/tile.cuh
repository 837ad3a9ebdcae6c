// tile.cuh — K3: the tile rasterizer of the binned path. One CTA owns one 32x32-px screen tile.
//
// Replaces Rasterizer::RasterizeBin (Rasterizer.cpp:696-739), TriangleEdgeVars::Setup (:296-329),
// Rasterizer::DrawTriangle<> (Rasterizer.h:250-328) and FS_EncodeSurfaceId<false> (Shading.cpp:309-331).
//
// The tile's depth (and, implicitly, surface id) is staged in shared memory as 1024 64-bit
// depth|id keys laid out exactly like the framebuffer's 4x4 fragments, so the epilogue is eight
// 512-byte coalesced stores per layer. Triangles are taken from the tile's list 256 at a time:
//   * a triangle whose pixel region inside the tile is small is rasterized by the thread that loaded it;
//   * the others are queued in shared memory and rasterized one per warp: a coarse pass evaluates the
//     tile's 32 8x4-px blocks at once (one block per lane, trivial reject per edge), then the warp
//     visits each surviving block with one pixel per lane (fine test);
//   * the wide triangles (more than kBigTriTileLimit tiles) listed for the tile's 256 x 256-px super-tile
//     are walked after the tile's own list and take the same two routes.
// Depth resolution is an order-independent max on the keys (CAS on shared memory, almost always
// skipped by a plain-load pre-check), which reproduces the reference's one-worker order.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kTileThreads = 256;
constexpr int kTileSmallArea = 32;   // pixels (inside the tile) a single thread rasterizes itself

__device__ __forceinline__ void smem_key_max(unsigned long long* s, uint32_t off, unsigned long long key) {
    unsigned long long cur = s[off];
    while (key > cur) {
        unsigned long long old = atomicCAS(&s[off], cur, key);
        if (old == cur) break;
        cur = old;
    }
}
// index of pixel (lx, ly) of the tile in the fragment-tiled shared array
__device__ __forceinline__ uint32_t tile_smem_index(uint32_t lx, uint32_t ly) {
    return (ly >> 2) * 128u + (lx >> 2) * 16u + (ly & 3u) * 4u + (lx & 3u);
}

__device__ __forceinline__ TriRecord load_record(const TriRecord* tris, uint32_t i) {
    const uint4* src = reinterpret_cast<const uint4*>(tris + i);
    uint4 a = __ldg(src), b = __ldg(src + 1);
    TriRecord t;
    t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
    t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
    return t;
}

// Rasterize `t` into the tile with one thread (small regions).
__device__ __forceinline__ void tile_raster_thread(const TriRecord& t, const BBox& r, const FrameParams& fp,
                                                   int32_t tileX0, int32_t tileY0, unsigned long long* keys) {
    Edges e;
    edge_setup(t, fp.halfW, fp.halfH, e);
    uint32_t rowE0 = (uint32_t)e.e0 + (uint32_t)e.a12 * (uint32_t)r.minX + (uint32_t)e.b12 * (uint32_t)r.minY;
    uint32_t rowE1 = (uint32_t)e.e1 + (uint32_t)e.a20 * (uint32_t)r.minX + (uint32_t)e.b20 * (uint32_t)r.minY;
    uint32_t rowE2 = (uint32_t)e.e2 + (uint32_t)e.a01 * (uint32_t)r.minX + (uint32_t)e.b01 * (uint32_t)r.minY;
    for (int32_t y = r.minY; y < r.maxY; y++) {
        uint32_t e0 = rowE0, e1 = rowE1, e2 = rowE2;
        for (int32_t x = r.minX; x < r.maxX; x++) {
            if ((int32_t)(e0 | e1 | e2) >= 0) {
                float d = pixel_depth(e, (int32_t)e1, (int32_t)e2);
                if (d > 0.0f) smem_key_max(keys, tile_smem_index((uint32_t)(x - tileX0), (uint32_t)(y - tileY0)), make_key(d, t.id));
            }
            e0 += (uint32_t)e.a12; e1 += (uint32_t)e.a20; e2 += (uint32_t)e.a01;
        }
        rowE0 += (uint32_t)e.b12; rowE1 += (uint32_t)e.b20; rowE2 += (uint32_t)e.b01;
    }
}

// Rasterize `t` into the tile with one warp: coarse (32 blocks of 8x4 px) then fine.
__device__ __forceinline__ void tile_raster_warp(const TriRecord& t, const FrameParams& fp, int32_t tileX0, int32_t tileY0,
                                                 unsigned long long* keys) {
    const uint32_t lane = lane_id();
    BBox r;
    raster_region<false>(t.pos0, t.pos1, t.pos2, fp, r);
    r.minX = max(r.minX, tileX0); r.minY = max(r.minY, tileY0);
    r.maxX = min(r.maxX, tileX0 + kTileSize); r.maxY = min(r.maxY, tileY0 + kTileSize);
    Edges e;
    edge_setup(t, fp.halfW, fp.halfH, e);
    const bool wrapFree = edges_wrap_free(t, e, fp);

    // coarse: lane owns block (lane & 3, lane >> 2) of the tile's 4 x 8 grid of 8x4-px blocks
    int32_t bxp = tileX0 + (int32_t)(lane & 3u) * 8, byp = tileY0 + (int32_t)(lane >> 2) * 4;
    bool alive = bxp < r.maxX && bxp + 8 > r.minX && byp < r.maxY && byp + 4 > r.minY;
    if (alive && wrapFree) {
        int32_t v0 = e.e0 + e.a12 * bxp + e.b12 * byp;
        int32_t v1 = e.e1 + e.a20 * bxp + e.b20 * byp;
        int32_t v2 = e.e2 + e.a01 * bxp + e.b01 * byp;
        v0 += (e.a12 > 0 ? e.a12 * 7 : 0) + (e.b12 > 0 ? e.b12 * 3 : 0);
        v1 += (e.a20 > 0 ? e.a20 * 7 : 0) + (e.b20 > 0 ? e.b20 * 3 : 0);
        v2 += (e.a01 > 0 ? e.a01 * 7 : 0) + (e.b01 > 0 ? e.b01 * 3 : 0);
        alive = (v0 | v1 | v2) >= 0;
    }
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, alive);
    // fine: lane owns pixel (lane & 7, lane >> 3) of the current block
    const uint32_t fx = lane & 7u, fy = lane >> 3;
    while (todo) {
        uint32_t blk = (uint32_t)__ffs(todo) - 1u;
        todo &= todo - 1u;
        int32_t px = tileX0 + (int32_t)((blk & 3u) * 8u + fx), py = tileY0 + (int32_t)((blk >> 2) * 4u + fy);
        if (px >= r.minX && px < r.maxX && py >= r.minY && py < r.maxY) {
            uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * (uint32_t)px + (uint32_t)e.b12 * (uint32_t)py;
            uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * (uint32_t)px + (uint32_t)e.b20 * (uint32_t)py;
            uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * (uint32_t)px + (uint32_t)e.b01 * (uint32_t)py;
            if ((int32_t)(e0 | e1 | e2) >= 0) {
                float d = pixel_depth(e, (int32_t)e1, (int32_t)e2);
                if (d > 0.0f) smem_key_max(keys, tile_smem_index((uint32_t)(px - tileX0), (uint32_t)(py - tileY0)), make_key(d, t.id));
            }
        }
    }
}

// The tile is staged from the frame's 64-bit key buffer, which already holds the seeds (pre-draw depth)
// and every small triangle the mesh kernel rasterized inline; this kernel adds the tile's binned
// triangles and stores the tile back with 128-bit coalesced writes (the key buffer shares the
// framebuffer's 4x4-tiled pixel order). Tiles without binned triangles exit at once. Keys become the
// depth / surface-id layers lazily (k_keys_unpack) or are consumed directly by the resolve pass.
__global__ void __launch_bounds__(kTileThreads)
k_tile_raster(const TriRecord* __restrict__ tris, const uint32_t* __restrict__ tileOffset, const uint32_t* __restrict__ activeTiles,
              const uint32_t* __restrict__ binEntries, const uint32_t* __restrict__ superOffset, const uint32_t* __restrict__ superEntries,
              FrameParams fp, unsigned long long* __restrict__ keysGlobal, DevCtl* __restrict__ ctl) {
    __shared__ __align__(16) unsigned long long keys[kTilePixels];
    __shared__ uint4 wideRecs[kTileThreads][2];     // records of the chunk's triangles that need a whole warp
    __shared__ uint32_t wideCount;

    if (ctl->overflow) return;
    const uint32_t tid = threadIdx.x;
    // persistent CTAs over the compacted list of tiles that have work (the launch order of a plain
    // one-CTA-per-tile grid would park the busy tiles behind thousands of empty ones)
    const uint32_t numActive = ctl->numActiveTiles;
  for (uint32_t at = blockIdx.x; at < numActive; at += gridDim.x) {
    const uint32_t tile = activeTiles[at];
    const uint32_t tx = tile % fp.tilesX, ty = tile / fp.tilesX;
    const int32_t tileX0 = (int32_t)(tx << kTileShift), tileY0 = (int32_t)(ty << kTileShift);
    const uint32_t listBegin = tileOffset[tile], listEnd = tileOffset[tile + 1];
    const uint32_t sh = kSuperShift - kTileShift, superX = (fp.tilesX + (1u << sh) - 1u) >> sh;
    const uint32_t superTile = (ty >> sh) * superX + (tx >> sh);
    const uint32_t bigBegin = superOffset[superTile], numBig = superOffset[superTile + 1] - bigBegin;
    if (listBegin == listEnd && numBig == 0) continue;    // nothing binned here: the keys are already final

    // ---- stage the tile: thread owns 4 consecutive pixels (one row of a 4x4 fragment)
    const uint32_t fr = tid >> 5, l4 = (tid & 31u) * 4u;                 // fragment row, first of 4 pixels in it
    const uint32_t gx = (uint32_t)tileX0 + (l4 >> 4) * 4u, gy = (uint32_t)tileY0 + fr * 4u + ((l4 >> 2) & 3u);
    const bool inFb = gx < fp.width && gy < fp.height;
    const uint32_t gOff = fb_pixel_offset(gx, gy, fp.width);
    {
        ulonglong2 a = make_ulonglong2(kKeySeed, kKeySeed), b = a;
        if (inFb) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(keysGlobal + gOff);
            a = __ldcg(src); b = __ldcg(src + 1);      // written with L2 atomics by the mesh kernel
        }
        ulonglong2* k = reinterpret_cast<ulonglong2*>(keys + fr * 128u + l4);
        k[0] = a; k[1] = b;
    }
    if (tid == 0) wideCount = 0;
    __syncthreads();

    // ---- triangles: the tile's list, then its super-tile's list of wide triangles
    const uint32_t total = (listEnd - listBegin) + numBig;
    for (uint32_t base = 0; base < total; base += kTileThreads) {
        uint32_t j = base + tid;
        if (j < total) {
            uint32_t triIdx = j < (listEnd - listBegin) ? binEntries[listBegin + j] : superEntries[bigBegin + (j - (listEnd - listBegin))];
            TriRecord t = load_record(tris, triIdx);
            BBox r;
            if (raster_region<false>(t.pos0, t.pos1, t.pos2, fp, r)) {
                r.minX = max(r.minX, tileX0); r.minY = max(r.minY, tileY0);
                r.maxX = min(r.maxX, tileX0 + kTileSize); r.maxY = min(r.maxY, tileY0 + kTileSize);
                int32_t w = r.maxX - r.minX, h = r.maxY - r.minY;
                if (w > 0 && h > 0) {
                    if (w * h <= kTileSmallArea) {
                        tile_raster_thread(t, r, fp, tileX0, tileY0, keys);
                    } else {   // park the record in shared memory for the warp-cooperative pass (no second global fetch)
                        uint32_t slot = atomicAdd(&wideCount, 1u);
                        wideRecs[slot][0] = make_uint4(t.pos0, t.pos1, t.pos2, __float_as_uint(t.z0));
                        wideRecs[slot][1] = make_uint4(__float_as_uint(t.z1), __float_as_uint(t.z2), t.id, t.aux);
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t nWide = wideCount;
        for (uint32_t w = tid >> 5; w < nWide; w += kTileThreads / 32) {
            uint4 a = wideRecs[w][0], b = wideRecs[w][1];
            TriRecord t;
            t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
            t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
            tile_raster_warp(t, fp, tileX0, tileY0, keys);
        }
        __syncthreads();
        if (tid == 0) wideCount = 0;      // ordered before the next chunk's atomics by the barrier after its loads
        if (base + kTileThreads < total) __syncthreads();
    }

    // ---- epilogue: the tile goes back to the key buffer, 2 x 128-bit stores per thread, 1 KB per warp
    if (inFb) {
        const ulonglong2* k = reinterpret_cast<const ulonglong2*>(keys + fr * 128u + l4);
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(keysGlobal + gOff);
        dst[0] = k[0]; dst[1] = k[1];
    }
    if (tid == 0 && listBegin != listEnd) atomicAdd(&ctl->perf[3], 1ull);   // BinQueueFlushes (Rasterizer.cpp:622)
    __syncthreads();      // the shared tile is reused by the next active tile
  }
}

}  // namespace swrb
