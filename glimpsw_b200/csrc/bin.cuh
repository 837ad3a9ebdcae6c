// bin.cuh — K2: the binner, a two-pass count / prefix-scan / scatter into 32x32-px screen tiles.
//
// Replaces Rasterizer::DistributeToBins + BinQueue::{InsertBin,Commit} (Rasterizer.cpp:664-695,
// :768-820). The CPU keeps per-worker fixed-size bins and flushes batches behind spin barriers; on the
// GPU the count pass rides along in the mesh kernel (mesh.cuh), a single block scans the per-tile
// counts, and the scatter pass writes compact per-tile triangle lists. Triangles overlapping more than
// kBigTriTileLimit tiles are not expanded: they sit in a short "big" list every tile walks.
#pragma once

#include "common.cuh"

namespace swrb {

// Tile range of a record, identical to the one the count pass used.
__device__ __forceinline__ uint32_t tile_range(const TriRecord& t, const FrameParams& fp,
                                               uint32_t& tx0, uint32_t& ty0, uint32_t& tx1, uint32_t& ty1) {
    BBox r;
    if (!raster_region(t.pos0, t.pos1, t.pos2, fp.halfW, fp.halfH, r)) return 0;
    tx0 = (uint32_t)(r.minX >> kTileShift); ty0 = (uint32_t)(r.minY >> kTileShift);
    tx1 = (uint32_t)((r.maxX - 1) >> kTileShift); ty1 = (uint32_t)((r.maxY - 1) >> kTileShift);
    return (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
}

// Exclusive scan of per-tile counts -> list offsets; also zeroes the scatter cursors. One block.
__global__ void __launch_bounds__(1024)
k_tile_scan(const uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tileOffset, uint32_t* __restrict__ tileCursor,
            uint32_t numTiles, uint32_t binCapacity, DevCtl* __restrict__ ctl) {
    __shared__ uint32_t warpSums[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t per = (numTiles + 1023u) / 1024u;
    const uint32_t begin = tid * per, end = min(begin + per, numTiles);
    uint32_t local = 0;
    for (uint32_t i = begin; i < end; i++) local += tileCount[i];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warpSums[lane], wi = w;
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
        tileOffset[i] = run;
        tileCursor[i] = 0;
        run += tileCount[i];
    }
    if (tid == 1023) {
        tileOffset[numTiles] = run;
        ctl->binTotal = run;
        if (run > binCapacity) atomicExch(&ctl->overflow, 3u);
    }
}

// Scatter pass: one thread per triangle record writes its index into every tile list it overlaps.
__global__ void __launch_bounds__(256)
k_bin_scatter(const TriRecord* __restrict__ tris, FrameParams fp, const uint32_t* __restrict__ tileOffset,
              uint32_t* __restrict__ tileCursor, uint32_t* __restrict__ binEntries, DevCtl* __restrict__ ctl) {
    const uint32_t n = ctl->overflow ? 0u : ctl->triCount;
    const uint32_t lane = lane_id();
    const uint32_t stride = gridDim.x * blockDim.x;
    // warp-uniform trip count so the ballots below stay convergent
    const uint32_t first = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
    for (uint32_t base = first; base < n; base += stride) {
        uint32_t i = base + lane;
        uint32_t tx0 = 0, ty0 = 0, tx1 = 0, ty1 = 0, nTiles = 0;
        if (i < n) {
            const uint4* src = reinterpret_cast<const uint4*>(tris + i);
            uint4 a = __ldg(src);
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
            binEntries[tileOffset[tile] + slot] = i;
        } else if (nTiles > 1 && nTiles <= (uint32_t)kBigTriTileLimit) {
            for (uint32_t ty = ty0; ty <= ty1; ty++)
                for (uint32_t tx = tx0; tx <= tx1; tx++) {
                    uint32_t tl = ty * fp.tilesX + tx;
                    uint32_t slot = atomicAdd(&tileCursor[tl], 1u);
                    binEntries[tileOffset[tl] + slot] = i;
                }
        }
    }
}

}  // namespace swrb
