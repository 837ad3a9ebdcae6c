// bin.cuh — K2: the binner, a two-pass count / prefix-scan / scatter into 32x32-px screen tiles.
//
// Replaces Rasterizer::DistributeToBins + BinQueue::{InsertBin,Commit} (Rasterizer.cpp:664-695,
// :768-820). The CPU keeps per-worker fixed-size bins and flushes batches behind spin barriers; here
//   pass 1 (count)   rides along in the mesh kernel: per-tile counters, warp-aggregated atomics;
//   pass 2 (scan)    is the prologue of k_bin_scatter: every block scans the (at most 8281) counters
//                    into shared memory by itself — redundant but free of a launch and of a global sync;
//                    block 0 also publishes the offsets and the list of non-empty tiles;
//   pass 3 (scatter) writes compact per-tile triangle lists with warp ballots / match_any.
// Triangles overlapping more than kBigTriTileLimit tiles are not expanded per tile: they are listed per
// 256 x 256-px SUPER-TILE (8 x 8 tiles; at most 12 x 12 = 144 of them at the 2896-px limit) with the same
// count / scan / scatter, and a tile walks its own list plus its super-tile's. A screen-filling triangle then
// costs 40 list entries at 1080p instead of 2,040, and a tile only meets the wide triangles near it (round 1
// kept ONE frame-wide big list that every tile walked: any wide triangle made every tile active).
// Only triangles too large for the mesh kernel's inline raster get here.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kScatterThreads = 256;

// Tile range of a record, identical to the one the count pass used.
__device__ __forceinline__ uint32_t tile_range(const TriRecord& t, const FrameParams& fp,
                                               uint32_t& tx0, uint32_t& ty0, uint32_t& tx1, uint32_t& ty1) {
    BBox r;
    if (!raster_region(t.pos0, t.pos1, t.pos2, fp, r)) return 0;
    tx0 = (uint32_t)(r.minX >> kTileShift); ty0 = (uint32_t)(r.minY >> kTileShift);
    tx1 = (uint32_t)((r.maxX - 1) >> kTileShift); ty1 = (uint32_t)((r.maxY - 1) >> kTileShift);
    return (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
}

// Exclusive scan of `n` counters sitting in shared memory, in place; s[n] receives the total.
__device__ __forceinline__ void block_exclusive_scan_smem(uint32_t* s, uint32_t n, uint32_t* warpSums /* smem[32] */) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t per = (n + kScatterThreads - 1) / kScatterThreads;
    const uint32_t begin = min(tid * per, n), end = min(begin + per, n);
    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) local += s[i];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < kScatterThreads / 32 ? warpSums[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += v;
        }
        warpSums[lane] = wi - w;   // exclusive
    }
    __syncthreads();
    uint32_t run = warpSums[warp] + incl - local;
    for (uint32_t i = begin; i < end; i++) {
        uint32_t c = s[i];
        s[i] = run;
        run += c;
    }
    if (tid == kScatterThreads - 1) s[n] = run;
    __syncthreads();
}

constexpr uint32_t kMaxSuperTiles = 144;     // ceil(2896 / 256)^2

struct BinBuffers {                          // per device, sized for the framebuffer's tile grid
    uint32_t* tileCount;  uint32_t* tileOffset;  uint32_t* tileCursor;  uint32_t* activeTiles;
    uint32_t* superCount; uint32_t* superOffset; uint32_t* superCursor;   // kMaxSuperTiles (+1) words each
    uint32_t* binEntries; uint32_t binCapacity;
    uint32_t* superEntries; uint32_t superCapacity;
};

// Passes 2 + 3. Dynamic shared memory: (numTiles + 1) * 4 bytes.
__global__ void __launch_bounds__(kScatterThreads)
k_bin_scatter(const TriRecord* __restrict__ tris, FrameParams fp, BinBuffers bb, DevCtl* __restrict__ ctl) {
    extern __shared__ uint32_t sOffset[];
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t sSuper[kMaxSuperTiles + 1];
    const uint32_t* __restrict__ tileCount = bb.tileCount;
    uint32_t* __restrict__ tileOffset = bb.tileOffset; uint32_t* __restrict__ tileCursor = bb.tileCursor;
    uint32_t* __restrict__ activeTiles = bb.activeTiles; uint32_t* __restrict__ binEntries = bb.binEntries;
    const uint32_t binCapacity = bb.binCapacity;
    const uint32_t numTiles = fp.tilesX * fp.tilesY;
    const uint32_t sh = kSuperShift - kTileShift, superX = (fp.tilesX + (1u << sh) - 1u) >> sh;
    const uint32_t numSuper = superX * ((fp.tilesY + (1u << sh) - 1u) >> sh);
    const uint32_t n = ctl->overflow ? 0u : ctl->triCount;
    if (n == 0) return;     // no records (every triangle was rasterized inline): k_frame_begin left binTotal / numActiveTiles at 0,
                            // which is all the tile rasterizer looks at before it returns

    // ---- pass 2: counts (L2, produced by the mesh kernel's atomics) -> exclusive offsets in shared memory
    for (uint32_t i = threadIdx.x; i < numTiles; i += kScatterThreads) sOffset[i] = __ldcg(tileCount + i);
    for (uint32_t i = threadIdx.x; i < numSuper; i += kScatterThreads) sSuper[i] = __ldcg(bb.superCount + i);
    __syncthreads();
    if (blockIdx.x == 0) {     // publish the tiles the tile rasterizer has to visit (before the counts are overwritten):
        for (uint32_t i = threadIdx.x; i < numTiles; i += kScatterThreads) {   // its own list or its super-tile's is non-empty
            const uint32_t tx = i % fp.tilesX, ty = i / fp.tilesX;
            if (sOffset[i] != 0 || sSuper[(ty >> sh) * superX + (tx >> sh)] != 0) activeTiles[atomicAdd(&ctl->numActiveTiles, 1u)] = i;
        }
        __syncthreads();
    }
    block_exclusive_scan_smem(sOffset, numTiles, warpSums);
    if (threadIdx.x == 0) {    // at most 144 super-tiles: a serial scan is a few hundred cycles
        uint32_t run = 0;
        for (uint32_t i = 0; i < numSuper; i++) { const uint32_t c = sSuper[i]; sSuper[i] = run; run += c; }
        sSuper[numSuper] = run;
    }
    __syncthreads();
    const bool overflow = sOffset[numTiles] > binCapacity || sSuper[numSuper] > bb.superCapacity;
    if (blockIdx.x == 0) {
        for (uint32_t i = threadIdx.x; i <= numTiles; i += kScatterThreads) tileOffset[i] = sOffset[i];
        for (uint32_t i = threadIdx.x; i <= numSuper; i += kScatterThreads) bb.superOffset[i] = sSuper[i];
        if (threadIdx.x == 0) {
            ctl->binTotal = sOffset[numTiles];
            ctl->superTotal = sSuper[numSuper];
            if (overflow) atomicExch(&ctl->overflow, 3u);
        }
    }
    if (overflow) return;

    // ---- pass 3: scatter. Records are dealt to the warps of the grid round-robin (record i -> warp i mod W), so a few
    // thousand records keep a thousand warps busy with 3-4 each instead of a hundred with 32 each. A lane takes one record;
    // single-tile records are appended with one atomic per distinct tile of the warp, multi-tile ones are expanded by the
    // whole warp, a lane per tile — the cursor atomics return values (slots), so a lane walking 64 tiles alone would be
    // 64 dependent L2 round trips.
    const uint32_t lane = lane_id();
    const uint32_t warp = (blockIdx.x * kScatterThreads + threadIdx.x) >> 5, numWarps = (gridDim.x * kScatterThreads) >> 5;
    for (uint32_t slot0 = 0; slot0 * numWarps < n; slot0 += 32) {       // warp-uniform trip count
        const uint64_t i64 = (uint64_t)(slot0 + lane) * numWarps + warp;
        const uint32_t i = i64 < 0xFFFFFFFFull ? (uint32_t)i64 : 0xFFFFFFFFu;
        uint32_t tx0 = 0, ty0 = 0, tx1 = 0, ty1 = 0, nTiles = 0;
        if (i < n) {
            uint4 a = __ldg(reinterpret_cast<const uint4*>(tris + i));
            TriRecord t;
            t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z;
            nTiles = tile_range(t, fp, tx0, ty0, tx1, ty1);
        }
        // single-tile triangles: one atomic per distinct tile in the warp
        const bool single = nTiles == 1;
        const uint32_t tile = ty0 * fp.tilesX + tx0;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, single);
        if (single) {
            uint32_t peers = __match_any_sync(mask, tile);
            uint32_t leader = (uint32_t)__ffs(peers) - 1u;
            uint32_t slot = 0;
            if (lane == leader) slot = atomicAdd(&tileCursor[tile], (uint32_t)__popc(peers));
            slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1u));
            binEntries[sOffset[tile] + slot] = i;
        }
        uint32_t multi = __ballot_sync(0xFFFFFFFFu, nTiles > 1);
        while (multi) {
            const uint32_t src = (uint32_t)__ffs(multi) - 1u;
            multi &= multi - 1u;
            const uint32_t bx0 = __shfl_sync(0xFFFFFFFFu, tx0, src), by0 = __shfl_sync(0xFFFFFFFFu, ty0, src);
            const uint32_t bx1 = __shfl_sync(0xFFFFFFFFu, tx1, src), by1 = __shfl_sync(0xFFFFFFFFu, ty1, src);
            const uint32_t rec = __shfl_sync(0xFFFFFFFFu, i, src), cnt = __shfl_sync(0xFFFFFFFFu, nTiles, src);
            if (cnt > (uint32_t)kBigTriTileLimit) {          // wide: one entry per 256-px super-tile
                const uint32_t sx0 = bx0 >> sh, sy0 = by0 >> sh, sw = (bx1 >> sh) - sx0 + 1, total = sw * ((by1 >> sh) - sy0 + 1);
                for (uint32_t t = lane; t < total; t += 32) {
                    const uint32_t st = (sy0 + t / sw) * superX + sx0 + t % sw;
                    const uint32_t slot = atomicAdd(&bb.superCursor[st], 1u);
                    bb.superEntries[sSuper[st] + slot] = rec;
                }
            } else {
                const uint32_t w = bx1 - bx0 + 1;
                for (uint32_t t = lane; t < cnt; t += 32) {
                    const uint32_t tl = (by0 + t / w) * fp.tilesX + bx0 + t % w;
                    const uint32_t slot = atomicAdd(&tileCursor[tl], 1u);
                    binEntries[sOffset[tl] + slot] = rec;
                }
            }
        }
    }
}

}  // namespace swrb
