// bin.cuh — K2: the binner, a two-pass count / prefix-scan / scatter into 32x32-px screen tiles.
//
// Replaces Rasterizer::DistributeToBins + BinQueue::{InsertBin,Commit} (Rasterizer.cpp:664-695,
// :768-820). The CPU keeps per-worker fixed-size bins and flushes batches behind spin barriers; here
//   pass 1 (count)   rides along in the mesh kernel: per-tile counters, warp-aggregated atomics;
//   pass 2 (scan)    is the prologue of k_bin_scatter: every block scans the (at most 8281) counters
//                    into shared memory by itself — redundant but free of a launch and of a global sync;
//                    block 0 also publishes the offsets and the list of non-empty tiles;
//   pass 3 (scatter) writes compact per-tile triangle lists with warp ballots / match_any.
// Triangles overlapping more than kBigTriTileLimit tiles are not expanded: they sit in a short "big"
// list every active tile walks. Only triangles too large for the mesh kernel's inline raster get here.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kScatterThreads = 256;

// Tile range of a record, identical to the one the count pass used.
__device__ __forceinline__ uint32_t tile_range(const TriRecord& t, const FrameParams& fp,
                                               uint32_t& tx0, uint32_t& ty0, uint32_t& tx1, uint32_t& ty1) {
    BBox r;
    if (!raster_region(t.pos0, t.pos1, t.pos2, fp.halfW, fp.halfH, r)) return 0;
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

// Passes 2 + 3. Dynamic shared memory: (numTiles + 1) * 4 bytes.
__global__ void __launch_bounds__(kScatterThreads)
k_bin_scatter(const TriRecord* __restrict__ tris, FrameParams fp, const uint32_t* __restrict__ tileCount,
              uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ tileCursor, uint32_t* __restrict__ activeTiles,
              uint32_t* __restrict__ binEntries, uint32_t binCapacity, DevCtl* __restrict__ ctl) {
    extern __shared__ uint32_t sOffset[];
    __shared__ uint32_t warpSums[32];
    const uint32_t numTiles = fp.tilesX * fp.tilesY;
    const uint32_t n = ctl->overflow ? 0u : ctl->triCount;
    if (n == 0) return;     // no records (every triangle was rasterized inline): k_frame_begin left binTotal / numActiveTiles at 0,
                            // which is all the tile rasterizer looks at before it returns

    // ---- pass 2: counts (L2, produced by the mesh kernel's atomics) -> exclusive offsets in shared memory
    for (uint32_t i = threadIdx.x; i < numTiles; i += kScatterThreads) sOffset[i] = __ldcg(tileCount + i);
    __syncthreads();
    if (blockIdx.x == 0) {     // publish the tiles the tile rasterizer has to visit (before the counts are overwritten):
        const bool allActive = ctl->bigCount != 0;   // big-list triangles are walked by every tile
        for (uint32_t i = threadIdx.x; i < numTiles; i += kScatterThreads)
            if (allActive || sOffset[i] != 0) activeTiles[atomicAdd(&ctl->numActiveTiles, 1u)] = i;
        __syncthreads();
    }
    block_exclusive_scan_smem(sOffset, numTiles, warpSums);
    const bool overflow = sOffset[numTiles] > binCapacity;
    if (blockIdx.x == 0) {
        for (uint32_t i = threadIdx.x; i <= numTiles; i += kScatterThreads) tileOffset[i] = sOffset[i];
        if (threadIdx.x == 0) {
            ctl->binTotal = sOffset[numTiles];
            if (overflow) atomicExch(&ctl->overflow, 3u);
        }
    }
    if (overflow) return;

    // ---- pass 3: one thread per triangle record writes its index into every tile list it overlaps
    const uint32_t lane = lane_id();
    const uint32_t stride = gridDim.x * kScatterThreads;
    const uint32_t first = (blockIdx.x * kScatterThreads + threadIdx.x) & ~31u;   // warp-uniform trip count
    for (uint32_t base = first; base < n; base += stride) {
        uint32_t i = base + lane;
        uint32_t tx0 = 0, ty0 = 0, tx1 = 0, ty1 = 0, nTiles = 0;
        if (i < n) {
            uint4 a = __ldg(reinterpret_cast<const uint4*>(tris + i));
            TriRecord t;
            t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z;
            nTiles = tile_range(t, fp, tx0, ty0, tx1, ty1);
        }
        // single-tile triangles: one atomic per distinct tile in the warp
        bool single = nTiles == 1;
        uint32_t tile = ty0 * fp.tilesX + tx0;
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, single);
        if (single) {
            uint32_t peers = __match_any_sync(mask, tile);
            uint32_t leader = (uint32_t)__ffs(peers) - 1u;
            uint32_t slot = 0;
            if (lane == leader) slot = atomicAdd(&tileCursor[tile], (uint32_t)__popc(peers));
            slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1u));
            binEntries[sOffset[tile] + slot] = i;
        } else if (nTiles > 1 && nTiles <= (uint32_t)kBigTriTileLimit) {
            for (uint32_t ty = ty0; ty <= ty1; ty++)
                for (uint32_t tx = tx0; tx <= tx1; tx++) {
                    uint32_t tl = ty * fp.tilesX + tx;
                    uint32_t slot = atomicAdd(&tileCursor[tl], 1u);
                    binEntries[sOffset[tl] + slot] = i;
                }
        }
    }
}

}  // namespace swrb
