// mesh.cuh — K1: meshlet cull-bit/frustum test + mesh shading + clip classification + early
// triangle setup, one warp per meshlet.
//
// Replaces (reference, one worker iteration of Rasterizer::DrawMeshlets, Rasterizer.cpp:535-594):
//   ShadeMeshlet                 Shading.cpp:281-307   (cull bit, SoA position transform, index copy)
//   CullMeshlets frustum test    Shading.cpp:803-809   (optional fused form)
//   GatherPos / index widen      Rasterizer.cpp:143-151, :553-558
//   Clipper::ComputeClipCodes    Rasterizer.cpp:353-397
//   TrianglePacket::Setup + bbox Rasterizer.cpp:257-289, :331-351
//
// Data flow per warp: 6 coalesced 128-byte position loads (lane L owns vertices L and L+32) and 24
// 128-bit index loads; per-VERTEX perspective divide / snap / outcodes are computed once and parked
// in shared memory (the CPU recomputes them per triangle corner; same inputs -> same bits); each lane
// then sets up triangles L, L+32, L+64, L+96 through a shared-memory 64-entry remap. Survivors are
// compacted with warp ballots and written as 32-byte records; in binned mode the same pass also
// counts triangles per 32x32 screen tile.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kMeshWarps = 8;   // warps (= meshlets in flight) per block

struct MeshWarpSmem {
    float nx[64], ny[64];       // NDC x, y (for the float determinant)
    float z[64], rw[64];        // z/w and 1/w
    uint32_t pos[64];           // packed 28.4 x | y << 16
    uint32_t flags[64];         // bits 0-5 Cohen-Sutherland outcodes, bit 6 inside guard band
    uint32_t idx[96];           // Indices[3][128] as bytes
};

__device__ __forceinline__ const DrawItem& find_draw(const DrawItem* draws, uint32_t numDraws, uint32_t work) {
    uint32_t lo = 0, hi = numDraws;   // last draw with firstWork <= work
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (draws[mid].firstWork <= work) lo = mid; else hi = mid;
    }
    return draws[lo];
}

// Warp-aggregated increment of per-tile counters for triangles that fall in one tile.
__device__ __forceinline__ void count_single_tile(uint32_t* tileCount, bool active, uint32_t tile) {
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, active);
    if (!active) return;
    uint32_t peers = __match_any_sync(mask, tile);
    if ((uint32_t)(__ffs(peers) - 1) == lane_id()) atomicAdd(&tileCount[tile], (uint32_t)__popc(peers));
}

template <bool kBinned>
__global__ void __launch_bounds__(kMeshWarps * 32)
k_mesh_setup(const swr_meshlet* __restrict__ meshlets, const swr_material* __restrict__ materials,
             const DrawItem* __restrict__ draws, uint32_t numDraws, uint32_t totalWork, FrameParams fp,
             TriRecord* __restrict__ tris, TriRecordW* __restrict__ trisW, uint32_t triCapacity,
             uint32_t* __restrict__ tileCount, uint32_t* __restrict__ bigList, DevCtl* __restrict__ ctl) {
    __shared__ MeshWarpSmem smem[kMeshWarps];
    MeshWarpSmem& s = smem[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    const uint32_t warpsTotal = gridDim.x * kMeshWarps;
    uint32_t nProcessed = 0, nRasterized = 0, nClipped = 0;

    for (uint32_t work = blockIdx.x * kMeshWarps + (threadIdx.x >> 5); work < totalWork; work += warpsTotal) {
        const DrawItem& d = find_draw(draws, numDraws, work);
        const uint32_t meshIdx = work - d.firstWork;

        // ---- ShadeMeshlet: cull bit (Shading.cpp:282-289)
        if (d.cullBitmap != nullptr) {
            uint32_t word = d.cullBitmap[meshIdx >> 4];
            if (((word >> (meshIdx & 15u)) & 1u) == 0) continue;
        }
        const swr_meshlet* m = meshlets + (d.meshletOffset + meshIdx);
        const uint4 hdrA = __ldg(reinterpret_cast<const uint4*>(m));            // BoundCenter, BoundRadius
        if (d.fusedCull) {                                                      // Shading.cpp:803-809
            float cx = __uint_as_float(hdrA.x), cy = __uint_as_float(hdrA.y), cz = __uint_as_float(hdrA.z);
            float rad = __uint_as_float(hdrA.w);
            bool vis = true;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                float dist = __fadd_rn(__fmaf_rn(cx, d.planes[i][0], __fmaf_rn(cy, d.planes[i][1], __fmul_rn(cz, d.planes[i][2]))), d.planes[i][3]);
                vis = vis && (dist > -rad);
            }
            if (!vis) continue;
        }
        const uint4 hdrB = __ldg(reinterpret_cast<const uint4*>(m) + 2);        // bytes 32..47: ConeAxis.yz, ConeCutoff, counts
        const uint32_t materialId = __ldg(reinterpret_cast<const uint32_t*>(m) + 12);
        const uint32_t numVerts = hdrB.w & 0xFFu, numTris = (hdrB.w >> 8) & 0xFFu;
        const uint32_t primCount = min(numTris, 128u);
        if (primCount == 0) continue;
        if (lane == 0) nProcessed += primCount;                                 // Rasterizer.cpp:545

        uint32_t cullMode = SWR_CULL_FRONT_CCW, fsId = 0;                       // Shading.cpp:302-306 (+ SURVEY App. B.4)
        if (materialId != SWR_NO_MATERIAL && materials != nullptr) {
            swr_material mat = materials[materialId];
            cullMode = mat.IsDoubleSided ? SWR_CULL_NONE : SWR_CULL_FRONT_CCW;
            fsId = mat.AlphaCutoff < 255 ? 1u : 0u;
        }

        // ---- index copy: 24 x 128-bit loads (Shading.cpp:300)
        if (lane < 24) reinterpret_cast<uint4*>(s.idx)[lane] = __ldg(reinterpret_cast<const uint4*>(m->Indices) + lane);

        // ---- transform + per-vertex setup: lane owns vertices lane and lane+32
        const uint32_t vertSlots = min((numVerts + 15u) & ~15u, 64u);           // reference walks 16-wide vectors
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t v = lane + h * 32;
            if (v < vertSlots) {
                float x = __ldg(&m->Positions[0][v]), y = __ldg(&m->Positions[1][v]), z = __ldg(&m->Positions[2][v]);
                // simd::mul(mat4, (pos,1)) — SIMD.h:457-464
                float cx = __fmaf_rn(x, d.M[0], __fmaf_rn(y, d.M[4], __fmaf_rn(z, d.M[8], d.M[12])));
                float cy = __fmaf_rn(x, d.M[1], __fmaf_rn(y, d.M[5], __fmaf_rn(z, d.M[9], d.M[13])));
                float cz = __fmaf_rn(x, d.M[2], __fmaf_rn(y, d.M[6], __fmaf_rn(z, d.M[10], d.M[14])));
                float cw = __fmaf_rn(x, d.M[3], __fmaf_rn(y, d.M[7], __fmaf_rn(z, d.M[11], d.M[15])));
                // ComputeClipCodes per vertex (Rasterizer.cpp:375-386)
                uint32_t f = 0;
                f |= (cx < -cw) ? 1u : 0u;
                f |= (cx > cw) ? 2u : 0u;
                f |= (cy < -cw) ? 4u : 0u;
                f |= (cy > cw) ? 8u : 0u;
                f |= (cz < -cw) ? 16u : 0u;
                f |= (cz > cw) ? 32u : 0u;
                f |= (fabsf(cx) < __fmul_rn(cw, fp.bx) && fabsf(cy) < __fmul_rn(cw, fp.by)) ? 64u : 0u;
                // perspective_div (SIMD.h:473-476) + snap (Rasterizer.cpp:272-279)
                float rw = __fdiv_rn(1.0f, cw);
                float nx = __fmul_rn(cx, rw), ny = __fmul_rn(cy, rw), nz = __fmul_rn(cz, rw);
                int32_t X = __float2int_rn(__fmul_rn(nx, fp.fixX)), Y = __float2int_rn(__fmul_rn(ny, fp.fixY));
                s.nx[v] = nx; s.ny[v] = ny; s.z[v] = nz; s.rw[v] = rw;
                s.pos[v] = ((uint32_t)X & 0xFFFFu) | ((uint32_t)Y << 16);
                s.flags[v] = f;
            }
        }
        __syncwarp();

        // ---- triangles: lane owns prims lane + 32k
        const uint8_t* idx = reinterpret_cast<const uint8_t*>(s.idx);
        TriRecord rec[4];
        float recW[4][3];
        uint32_t tileLo[4] = {1, 1, 1, 1}, tileHi[4] = {0, 0, 0, 0};   // empty range unless set below
        uint32_t keepBits = 0;      // bit k: this lane's k-th triangle survives
        uint32_t total = 0, myOffset[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t prim = lane + k * 32;
            bool keep = false, nonTrivial = false;
            if (prim < primCount) {
                uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                uint32_t f0 = s.flags[i0], f1 = s.flags[i1], f2 = s.flags[i2];
                uint32_t partial = f0 | f1 | f2, combined = f0 & f1 & f2;
                bool visible = (combined & 63u) == 0;                                   // Rasterizer.cpp:389
                bool trivial = (combined & 64u) != 0 && (partial & 48u) == 0;           // :386-388
                nonTrivial = visible && !trivial;                                       // :393
                if (visible && trivial) {
                    // TrianglePacket::Setup (Rasterizer.cpp:257-289)
                    float x0 = s.nx[i0], y0 = s.ny[i0], x1 = s.nx[i1], y1 = s.ny[i1], x2 = s.nx[i2], y2 = s.ny[i2];
                    float det = __fsub_rn(__fmul_rn(__fsub_rn(x2, x0), __fsub_rn(y1, y0)),
                                          __fmul_rn(__fsub_rn(x0, x1), __fsub_rn(y0, y2)));
                    if (cullMode != SWR_CULL_FRONT_CCW) {
                        bool flip = (cullMode == SWR_CULL_FRONT_CW) ? true : (det < 0.0f);
                        det = flip ? -det : det;
                    }
                    uint32_t p0 = s.pos[i0], p1 = s.pos[i1], p2 = s.pos[i2];
                    uint32_t bbMin, bbMax;
                    ref_render_bbox(p0, p1, p2, fp.halfW, fp.halfH, bbMin, bbMax);
                    keep = det > 0.0f && lo16(bbMin) < lo16(bbMax) && hi16(bbMin) < hi16(bbMax);   // :269, :283
                    if (keep) {
                        rec[k].pos0 = p0; rec[k].pos1 = p1; rec[k].pos2 = p2;
                        rec[k].z0 = s.z[i0]; rec[k].z1 = s.z[i1]; rec[k].z2 = s.z[i2];
                        rec[k].id = (d.meshletOffset + meshIdx) * SWR_MAX_PRIMS + prim;
                        rec[k].aux = fsId;
                        if (fsId) { recW[k][0] = s.rw[i0]; recW[k][1] = s.rw[i1]; recW[k][2] = s.rw[i2]; }
                        if (kBinned) {
                            BBox r;
                            if (raster_region(p0, p1, p2, fp.halfW, fp.halfH, r)) {
                                tileLo[k] = (uint32_t)(r.minX >> kTileShift) | ((uint32_t)(r.minY >> kTileShift) << 16);
                                tileHi[k] = (uint32_t)((r.maxX - 1) >> kTileShift) | ((uint32_t)((r.maxY - 1) >> kTileShift) << 16);
                            }   // else: counted as rasterized (the reference does) but touches no pixel
                        }
                    }
                }
            }
            uint32_t keepMask = __ballot_sync(0xFFFFFFFFu, keep);
            uint32_t clipMask = __ballot_sync(0xFFFFFFFFu, nonTrivial);
            myOffset[k] = total + __popc(keepMask & ((1u << lane) - 1u));
            total += __popc(keepMask);
            if (keep) keepBits |= 1u << k;
            if (lane == 0) { nRasterized += __popc(keepMask); nClipped += __popc(clipMask); }   // :579, :568
        }

        // ---- compact + write records (one atomic per meshlet)
        uint32_t base = 0;
        if (lane == 0 && total > 0) base = atomicAdd(&ctl->triCount, total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        bool fits = base + total <= triCapacity;
        if (!fits && lane == 0) atomicExch(&ctl->overflow, 1u);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            bool keep = (keepBits >> k) & 1u;
            uint32_t slot = base + myOffset[k];
            if (keep && fits) {
                uint4* dst = reinterpret_cast<uint4*>(tris + slot);
                dst[0] = make_uint4(rec[k].pos0, rec[k].pos1, rec[k].pos2, __float_as_uint(rec[k].z0));
                dst[1] = make_uint4(__float_as_uint(rec[k].z1), __float_as_uint(rec[k].z2), rec[k].id, rec[k].aux);
                if (rec[k].aux & 1u) *reinterpret_cast<float4*>(trisW + slot) = make_float4(recW[k][0], recW[k][1], recW[k][2], 0.0f);
            }
            if (kBinned) {
                // per-tile counting: one-tile triangles are warp-aggregated, the rest loop over their tiles
                uint32_t tx0 = tileLo[k] & 0xFFFFu, ty0 = tileLo[k] >> 16, tx1 = tileHi[k] & 0xFFFFu, ty1 = tileHi[k] >> 16;
                bool valid = keep && fits && tx0 <= tx1 && ty0 <= ty1;
                uint32_t nTiles = valid ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
                count_single_tile(tileCount, nTiles == 1, ty0 * fp.tilesX + tx0);
                if (nTiles > (uint32_t)kBigTriTileLimit) {
                    uint32_t b = atomicAdd(&ctl->bigCount, 1u);
                    bigList[b] = slot;          // capacity == triCapacity
                } else if (nTiles > 1) {
                    for (uint32_t ty = ty0; ty <= ty1; ty++)
                        for (uint32_t tx = tx0; tx <= tx1; tx++) atomicAdd(&tileCount[ty * fp.tilesX + tx], 1u);
                }
            }
        }
        __syncwarp();
    }

    // ---- perf counters: one atomic per warp per counter (Rasterizer.cpp:927-932 FlushThreadCounters)
    if (lane == 0) {
        if (nProcessed) atomicAdd(&ctl->perf[0], (unsigned long long)nProcessed);
        if (nRasterized) atomicAdd(&ctl->perf[1], (unsigned long long)nRasterized);
        if (nClipped) atomicAdd(&ctl->perf[2], (unsigned long long)nClipped);
    }
}

}  // namespace swrb
