// fbops.cuh — framebuffer kernels: clear, key-buffer init/unpack (direct path), de-tile.
//
// Framebuffer layout is the reference's (Rasterizer.h:10-78): layers of u32, 4x4-pixel tiles,
// layer 0 = colour / surface id, layer 1 = depth. All kernels move 128-bit words.
#pragma once

#include "common.cuh"

namespace swrb {

// Framebuffer::Clear / ClearLayer (Rasterizer.h:35-48): both layers in one pass when `two` is set.
__global__ void __launch_bounds__(256)
k_fb_clear(uint4* __restrict__ layerA, uint32_t valueA, uint4* __restrict__ layerB, uint32_t valueB, uint32_t numVec) {
    uint4 va = make_uint4(valueA, valueA, valueA, valueA), vb = make_uint4(valueB, valueB, valueB, valueB);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        layerA[i] = va;
        if (layerB != nullptr) layerB[i] = vb;
    }
}

// The draw's transient device state: per-tile / per-super-tile counters and cursors and the work counters in
// DevCtl (everything but `overflow`, which stays sticky until the host reads it, and the perf counters). The sizes
// of the draw that just finished are kept in ctl->last* for swrb_get_draw_stats. Called by thread `gid` of `stride`.
__device__ __forceinline__ void reset_draw_state(uint32_t gid, uint32_t stride, uint32_t* tileCount, uint32_t* tileCursor, uint32_t numTiles,
                                                 uint32_t* superCount, uint32_t* superCursor, DevCtl* ctl) {
    if (gid == 0) {
        ctl->lastTriCount = ctl->triCount; ctl->lastBigCount = ctl->bigCount; ctl->lastBinTotal = ctl->binTotal + ctl->superTotal;
        ctl->triCount = 0; ctl->bigCount = 0; ctl->binTotal = 0; ctl->numActiveTiles = 0; ctl->alphaCount = 0; ctl->clipCount = 0;
        ctl->workCursor = 0; ctl->superTotal = 0; ctl->sparseTiles = 0; ctl->denseTiles = 0;
    }
    for (uint32_t i = gid; i < numTiles; i += stride) { tileCount[i] = 0; tileCursor[i] = 0; }
    for (uint32_t i = gid; i < 160u; i += stride) { superCount[i] = 0; superCursor[i] = 0; }
}

// Start of a draw: seed the key buffer and reset the draw's transient device state in one launch.
// keys <- (depth << 32 | seed), where depth is the pixel's stored depth (depthLayer != null) or the
// pending clear's depth. A pixel keeps its seed unless a fragment of THIS draw passes the strict depth
// test, so earlier draws win ties like in the reference. numVec == 0: the keys already hold the seeds (the
// last resolve pass left them behind, k_resolve kReseed), only the state reset is needed — and not even that
// when the resolve pass did it too (the host then skips this launch).
__global__ void __launch_bounds__(256)
k_frame_begin(const uint4* __restrict__ depthLayer, uint32_t clearDepthBits, ulonglong2* __restrict__ keys, uint32_t numVec,
              uint32_t* __restrict__ tileCount, uint32_t* __restrict__ tileCursor, uint32_t numTiles,
              uint32_t* __restrict__ superCount, uint32_t* __restrict__ superCursor, DevCtl* __restrict__ ctl) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    reset_draw_state(gid, stride, tileCount, tileCursor, numTiles, superCount, superCursor, ctl);
    for (uint32_t i = gid; i < numVec; i += stride) {
        uint4 d = make_uint4(clearDepthBits, clearDepthBits, clearDepthBits, clearDepthBits);
        if (depthLayer != nullptr) d = depthLayer[i];
        keys[2 * i + 0] = make_ulonglong2(((unsigned long long)d.x << 32) | kKeySeed, ((unsigned long long)d.y << 32) | kKeySeed);
        keys[2 * i + 1] = make_ulonglong2(((unsigned long long)d.z << 32) | kKeySeed, ((unsigned long long)d.w << 32) | kKeySeed);
    }
}

// Keys -> framebuffer layers: depth + surface id of every pixel the draw won (FS_EncodeSurfaceId's
// masked stores, Shading.cpp:328-330). With clearAll the framebuffer was logically cleared before the
// draw, so every pixel is written (lost pixels get clearColor and the seed's depth = the clear depth).
// With depthOnly layer 0 is left alone (the resolve pass already replaced the ids by colour).
__global__ void __launch_bounds__(256)
k_keys_unpack(const ulonglong2* __restrict__ keys, uint4* __restrict__ color, uint4* __restrict__ depth, uint32_t numVec,
              int clearAll, uint32_t clearColor, int depthOnly, const DevCtl* __restrict__ ctl) {
    if (ctl->overflow) return;                                // the draw that produced these keys was aborted
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        ulonglong2 k0 = keys[2 * i], k1 = keys[2 * i + 1];
        uint32_t l0 = (uint32_t)k0.x, l1 = (uint32_t)k0.y, l2 = (uint32_t)k1.x, l3 = (uint32_t)k1.y;
        bool anyWon = (l0 & l1 & l2 & l3) != kKeySeed;
        if (!anyWon && !clearAll) continue;                  // nothing won in these 4 pixels
        depth[i] = make_uint4((uint32_t)(k0.x >> 32), (uint32_t)(k0.y >> 32), (uint32_t)(k1.x >> 32), (uint32_t)(k1.y >> 32));
        if (depthOnly) continue;
        bool allWon = l0 != kKeySeed && l1 != kKeySeed && l2 != kKeySeed && l3 != kKeySeed;
        uint4 c = make_uint4(clearColor, clearColor, clearColor, clearColor);
        if (!clearAll && !allWon) c = color[i];
        if (l0 != kKeySeed) c.x = rank_surface_id(kKeyIdBase - l0);
        if (l1 != kKeySeed) c.y = rank_surface_id(kKeyIdBase - l1);
        if (l2 != kKeySeed) c.z = rank_surface_id(kKeyIdBase - l2);
        if (l3 != kKeySeed) c.w = rank_surface_id(kKeyIdBase - l3);
        color[i] = c;
    }
}

// L2 eviction helper for benchmarks: streams `numVec` 128-bit words through the cache (read only).
__global__ void __launch_bounds__(256) k_l2_read(const uint4* __restrict__ src, uint32_t numVec, uint32_t* __restrict__ sink) {
    uint32_t acc = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        uint4 v = __ldcg(src + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345679u) sink[0] = acc;   // never true for the 0x5A fill pattern; keeps the loads alive
}

// Framebuffer::GetPixels (ImageHelpers.cpp:109-147): 4x4-tiled layer -> row-major. One thread moves
// one 4-pixel row of a tile (a 128-bit word) so both sides are 16-byte accesses; a warp covers 8 tiles
// x 4 rows = 512 contiguous source bytes.
__global__ void __launch_bounds__(256)
k_fb_detile(const uint4* __restrict__ layer, uint32_t* __restrict__ dst, uint32_t width, uint32_t height, uint32_t stride) {
    const uint32_t numVec = width * height / 4, tilesPerRow = width >> 2;
    const uint32_t step = gridDim.x * blockDim.x;
    // four independent loads in flight per thread: with a small grid (the peer-memory gather runs beside the
    // render kernels and must not take their SMs) the copy still keeps enough bytes in flight for NVLink
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < numVec; i0 += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t i = i0 + k * step;
            if (i < numVec) v[k] = layer[i];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t i = i0 + k * step;
            if (i < numVec) {
                uint32_t row = i & 3u, tile = i >> 2;
                uint32_t tx = tile % tilesPerRow, ty = tile / tilesPerRow;
                *reinterpret_cast<uint4*>(dst + (size_t)(ty * 4 + row) * stride + tx * 4) = v[k];
            }
        }
    }
}

// ---- multi-GPU composite exchange over NVLink peer memory ------------------------------------------------
// The de-tile kernel IS the transfer: it stores the finished view straight into the consumer GPU's memory, and the
// flow control that a collective library would add as extra kernels is folded into it:
//   * before the first store every block waits (one spinning thread) until the consumer has released the slot
//     (`waitFlag`, in this GPU's memory, written by the consumer's k_peer_collect);
//   * after the last store of the last block the producer raises the slot's ready flag in the CONSUMER's memory
//     (system-scope fence before it, so the pixels are visible there first).
// Flags carry monotonically increasing use counts, so nothing is ever reset across the link.
struct PeerSync {
    const unsigned long long* waitFlag;   // null: nothing to wait for
    unsigned long long waitValue;
    unsigned long long* signalFlag;       // null: no signal
    unsigned long long signalValue;
    uint32_t* blockCounter;               // scratch counter of THIS send (one of a ring per device), zero between uses
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256)
k_fb_detile_send(const uint4* __restrict__ layer, uint32_t* __restrict__ dst, uint32_t width, uint32_t height, uint32_t stride, PeerSync ps) {
    if (ps.waitFlag != nullptr) {
        if (threadIdx.x == 0) while (ld_acquire_sys(ps.waitFlag) < ps.waitValue) __nanosleep(100);
        __syncthreads();
    }
    const uint32_t numVec = width * height / 4, tilesPerRow = width >> 2;
    const uint32_t step = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < numVec; i0 += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t i = i0 + k * step;
            if (i < numVec) v[k] = layer[i];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t i = i0 + k * step;
            if (i < numVec) {
                uint32_t row = i & 3u, tile = i >> 2;
                uint32_t tx = tile % tilesPerRow, ty = tile / tilesPerRow;
                *reinterpret_cast<uint4*>(dst + (size_t)(ty * 4 + row) * stride + tx * 4) = v[k];
            }
        }
    }
    if (ps.signalFlag != nullptr) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t done = atomicAdd(ps.blockCounter, 1u) + 1u;
            if (done == gridDim.x) {
                *ps.blockCounter = 0;
                __threadfence_system();
                st_release_sys(ps.signalFlag, ps.signalValue);
            }
        }
    }
}

// Consumer side, one launch per slot: wait until every producer has raised its ready flag for this use of the slot
// (flags are in this GPU's memory), then release the slot on every producer (their ack flags, over NVLink).
struct PeerAcks { unsigned long long* flag[15]; };
__global__ void __launch_bounds__(32)
k_peer_collect(const unsigned long long* __restrict__ readyFlags, uint32_t n, unsigned long long expected, PeerAcks acks, unsigned long long ackValue) {
    const uint32_t t = threadIdx.x;
    if (t < n) while (ld_acquire_sys(readyFlags + t) < expected) __nanosleep(100);
    __syncwarp();
    if (t < n) st_release_sys(acks.flag[t], ackValue);
}

}  // namespace swrb
