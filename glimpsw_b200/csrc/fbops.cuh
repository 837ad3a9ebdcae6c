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

// Direct path, start of a draw: keys <- (stored depth << 32 | seed). A pixel keeps its seed unless a
// fragment of THIS draw passes the strict depth test, so earlier draws win ties like in the reference.
__global__ void __launch_bounds__(256)
k_keys_init(const uint4* __restrict__ depth, ulonglong2* __restrict__ keys, uint32_t numVec) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        uint4 d = depth[i];
        keys[2 * i + 0] = make_ulonglong2(((unsigned long long)d.x << 32) | kKeySeed, ((unsigned long long)d.y << 32) | kKeySeed);
        keys[2 * i + 1] = make_ulonglong2(((unsigned long long)d.z << 32) | kKeySeed, ((unsigned long long)d.w << 32) | kKeySeed);
    }
}
// Same, when the framebuffer was just cleared: no read.
__global__ void __launch_bounds__(256)
k_keys_fill(ulonglong2* __restrict__ keys, uint32_t depthBits, uint32_t numVec2) {
    unsigned long long k = ((unsigned long long)depthBits << 32) | kKeySeed;
    ulonglong2 v = make_ulonglong2(k, k);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec2; i += gridDim.x * blockDim.x) keys[i] = v;
}

// Direct path, end of a draw: write depth + surface id of every pixel this draw won (FS_EncodeSurfaceId's
// masked stores, Shading.cpp:328-330). With clearAll the framebuffer was logically cleared before the
// draw, so every pixel is written (lost pixels get clearColor and the seed's depth = the clear depth).
__global__ void __launch_bounds__(256)
k_keys_unpack(const ulonglong2* __restrict__ keys, uint4* __restrict__ color, uint4* __restrict__ depth, uint32_t numVec,
              int clearAll, uint32_t clearColor) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        ulonglong2 k0 = keys[2 * i], k1 = keys[2 * i + 1];
        uint32_t l0 = (uint32_t)k0.x, l1 = (uint32_t)k0.y, l2 = (uint32_t)k1.x, l3 = (uint32_t)k1.y;
        bool anyWon = (l0 & l1 & l2 & l3) != kKeySeed;
        if (!anyWon && !clearAll) continue;                  // nothing won in these 4 pixels
        bool allWon = l0 != kKeySeed && l1 != kKeySeed && l2 != kKeySeed && l3 != kKeySeed;
        uint4 c = make_uint4(clearColor, clearColor, clearColor, clearColor);
        if (!clearAll && !allWon) c = color[i];
        uint4 d = make_uint4((uint32_t)(k0.x >> 32), (uint32_t)(k0.y >> 32), (uint32_t)(k1.x >> 32), (uint32_t)(k1.y >> 32));
        if (l0 != kKeySeed) c.x = kKeyIdBase - l0;
        if (l1 != kKeySeed) c.y = kKeyIdBase - l1;
        if (l2 != kKeySeed) c.z = kKeyIdBase - l2;
        if (l3 != kKeySeed) c.w = kKeyIdBase - l3;
        color[i] = c; depth[i] = d;
    }
}

// Framebuffer::GetPixels (ImageHelpers.cpp:109-147): 4x4-tiled layer -> row-major. One thread moves
// one 4-pixel row of a tile (a 128-bit word) so both sides are 16-byte accesses; a warp covers 8 tiles
// x 4 rows = 512 contiguous source bytes.
__global__ void __launch_bounds__(256)
k_fb_detile(const uint4* __restrict__ layer, uint32_t* __restrict__ dst, uint32_t width, uint32_t height, uint32_t stride) {
    uint32_t numVec = width * height / 4;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numVec; i += gridDim.x * blockDim.x) {
        uint32_t row = i & 3u, tile = i >> 2;
        uint32_t tilesPerRow = width >> 2;
        uint32_t tx = tile % tilesPerRow, ty = tile / tilesPerRow;
        uint4 v = layer[i];
        *reinterpret_cast<uint4*>(dst + (size_t)(ty * 4 + row) * stride + tx * 4) = v;
    }
}

}  // namespace swrb
