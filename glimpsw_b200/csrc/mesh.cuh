// mesh.cuh — K1: meshlet cull-bit/frustum test + mesh shading + clip classification + triangle setup
// + rasterization of every triangle that is not "big", one warp per meshlet.
//
// Replaces (reference, one worker iteration of Rasterizer::DrawMeshlets, Rasterizer.cpp:535-594):
//   ShadeMeshlet                 Shading.cpp:281-307   (cull bit, SoA position transform, index copy)
//   CullMeshlets frustum test    Shading.cpp:803-809   (optional fused form)
//   GatherPos / index widen      Rasterizer.cpp:143-151, :553-558
//   Clipper::ComputeClipCodes    Rasterizer.cpp:353-397
//   TrianglePacket::Setup + bbox Rasterizer.cpp:257-289, :331-351
// and, for triangles whose pixel region is <= FrameParams::inlineMaxArea, also
//   TriangleEdgeVars::Setup      Rasterizer.cpp:296-329
//   DrawTriangle<> + FS_EncodeSurfaceId<false>   Rasterizer.h:250-328, Shading.cpp:309-331
//
// The warps of a persistent grid walk the work items (meshlets of the batch, in submission order) round-robin. The whole
// warp tests one meshlet (cull bit, bound sphere against the five frustum planes) and, if it survives, loads its 1,216 hot
// bytes straight into registers — 6 coalesced 128-byte position loads and 24 x 16 bytes of indices, all issued before
// anything depends on them. Per-VERTEX perspective divide / snap / outcodes are computed once (lane owns vertices L and
// L+32; the CPU recomputes them per triangle corner; same inputs -> same bits) and parked in 2 KB of shared memory; then
// lane L owns prims L + 32k: classification, determinant, bounding box and
//   * small (pixel region <= inlineMaxArea): rasterized right there by the lane, 64-bit depth|rank keys reduced straight
//     into the frame's key buffer with REDG.MAX.64 — no record, no bin entry, no second kernel;
//   * big: one slice of the global record array per meshlet (one atomic), 32-byte records, and on the binned path a
//     per-tile (or, over kBigTriTileLimit tiles, per-super-tile) count = pass 1 of the binner.
//
// Measured and dropped (profiles/r02_summary.md): a staged variant — the meshlet brought into shared memory by two bulk
// async copies (cp.async.bulk / UBLKCP on an mbarrier, double-buffered per warp), a cull phase with one LANE per meshlet
// feeding a visible list, survivors compacted by ballot before the inline raster — reached 27 instead of 23 active lanes
// per instruction and took the HBM latency off the warp's critical path, but needed 41 KB of shared memory per block, a
// grid-wide hand-over between its phases and ~15 % more instructions: C2 50 vs 35 us, a C4 view 207-233 vs 156-161 us.
// A warp-cooperative middle class (one warp working through a meshlet's mid-size triangles one after another) was a
// 100+ us serial tail in round 1; lanes in parallel or tiles in parallel win.
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kMeshWarps = 8;           // warps (= meshlets in flight) per block
constexpr int kInlineMaxArea = 256;     // largest pixel region a lane may rasterize itself (FrameParams::inlineMaxArea <= this; else: record + binner)

// Warp-aggregated increment of per-tile counters for triangles that fall in one tile.
__device__ __forceinline__ void count_single_tile(uint32_t* tileCount, bool active, uint32_t tile) {
    uint32_t mask = __ballot_sync(0xFFFFFFFFu, active);
    if (!active) return;
    uint32_t peers = __match_any_sync(mask, tile);
    if ((uint32_t)(__ffs(peers) - 1) == lane_id()) atomicAdd(&tileCount[tile], (uint32_t)__popc(peers));
}

// Rasterize a small triangle from registers into the key buffer (one lane).
__device__ __forceinline__ void raster_inline(const TriRecord& t, const BBox& r, const FrameParams& fp, unsigned long long* keys) {
    Edges e;
    edge_setup(t, fp.halfW, fp.halfH, e);
    uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * (uint32_t)r.minX + (uint32_t)e.b12 * (uint32_t)r.minY;
    uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * (uint32_t)r.minX + (uint32_t)e.b20 * (uint32_t)r.minY;
    uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * (uint32_t)r.minX + (uint32_t)e.b01 * (uint32_t)r.minY;
    // One flat loop over the region's pixels instead of nested row/column loops: the lanes of a warp walk regions
    // of different shapes, and a flat loop costs max(area) iterations per warp where nested loops cost
    // sum over rows of max(width). The edge values step by A inside a row and by B - (w-1)*A at the row's end,
    // all in wrapping 32-bit arithmetic — the same values the reference's incremental adds produce.
    const int32_t w = r.maxX - r.minX, n = w * (r.maxY - r.minY);
    const uint32_t wrap0 = (uint32_t)e.b12 - (uint32_t)(w - 1) * (uint32_t)e.a12;
    const uint32_t wrap1 = (uint32_t)e.b20 - (uint32_t)(w - 1) * (uint32_t)e.a20;
    const uint32_t wrap2 = (uint32_t)e.b01 - (uint32_t)(w - 1) * (uint32_t)e.a01;
    const unsigned long long keyLow = (unsigned long long)(kKeyIdBase - t.id);
    int32_t x = r.minX, y = r.minY;
    for (int32_t i = 0; i < n; i++) {
        if ((int32_t)(e0 | e1 | e2) >= 0) {                                       // Rasterizer.h:289-290
            float d = pixel_depth(e, (int32_t)e1, (int32_t)e2);                   // :296
            if (d > 0.0f) atomicMax(keys + fb_pixel_offset((uint32_t)x, (uint32_t)y, fp.width),
                                    ((unsigned long long)__float_as_uint(d) << 32) | keyLow);
        }
        const bool rowEnd = x + 1 == r.maxX;
        e0 += rowEnd ? wrap0 : (uint32_t)e.a12;
        e1 += rowEnd ? wrap1 : (uint32_t)e.a20;
        e2 += rowEnd ? wrap2 : (uint32_t)e.a01;
        x = rowEnd ? r.minX : x + 1;
        y += rowEnd ? 1 : 0;
    }
}

struct MeshOut {                         // where the kernel leaves what it does not rasterize itself
    TriRecord* tris; TriRecord* alphaTris; TriRecordW* alphaW; uint32_t triCapacity;
    uint32_t* tileCount; uint32_t* superCount;    // binned: pass 1 of the binner
    uint2* clipList;                              // unbinned + EnableClipping
    float4* clipCache;                            // per-vertex {x/w, y/w, 1/w, z/w} for this frame's resolve pass, or null
};


struct MeshWarpSmem {
    float nx[64], ny[64];       // NDC x, y (for the float determinant)
    float z[64], rw[64];        // z/w and 1/w
    uint32_t pos[64];           // packed 28.4 x | y << 16
    uint32_t flags[64];         // bits 0-5 Cohen-Sutherland outcodes, bit 6 inside guard band
    uint32_t idx[96];           // Indices[3][128] as bytes
    uint8_t big[128];           // prims that need a record
};

__device__ __forceinline__ const DrawItem& find_draw(const DrawItem* draws, uint32_t numDraws, uint32_t work) {
    uint32_t lo = 0, hi = numDraws;   // last draw with firstWork <= work
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (draws[mid].firstWork <= work) lo = mid; else hi = mid;
    }
    return draws[lo];
}

template <bool kBinned, bool kBand>   // kBand: the framebuffer has scissor rows (swrb_fb_set_scissor_rows); the plain frame pays nothing for them
// Persistent grid of 1..4 blocks per SM (swrb_device_set_mesh_occupancy): 4 = the whole register file for a lone frame,
// 1 leaves room for other render contexts' resolve blocks, whose issue-bound warps fill what these latency-bound ones leave idle.
__global__ void __launch_bounds__(kMeshWarps * 32, 4)      // 64 registers: up to 4 blocks = 32 warps per SM
k_mesh_setup(const swr_meshlet* __restrict__ meshlets, const swr_material* __restrict__ materials,
              const DrawItem* __restrict__ draws, uint32_t numDraws, uint32_t totalWork, FrameParams fp,
              unsigned long long* __restrict__ keys, MeshOut out, DevCtl* __restrict__ ctl) {
    __shared__ MeshWarpSmem smem[kMeshWarps];
    MeshWarpSmem& s = smem[threadIdx.x >> 5];
    (void)totalWork;
    TriRecord* const tris = out.tris; TriRecord* const alphaTris = out.alphaTris; TriRecordW* const alphaW = out.alphaW;
    const uint32_t triCapacity = out.triCapacity;
    uint32_t* const tileCount = out.tileCount;
    uint2* const clipList = out.clipList;
    float4* const clipCache = out.clipCache;
    const uint32_t lane = lane_id();
    const uint32_t warpsTotal = gridDim.x * kMeshWarps;
    uint32_t nProcessed = 0, nRasterized = 0, nClipped = 0;

    // The warp's work items are base, base + W, base + 2W, ... (W = warps in the grid). 32 of them are TESTED at a time, one per
    // lane — draw lookup, ShadeMeshlet's cull bit, CullMeshlets' frustum test — so a culled meshlet costs a lane, not a warp
    // iteration; the survivors (a ballot) are then shaded one after another by the whole warp.
    for (uint32_t base = fp.workBegin + blockIdx.x * kMeshWarps + (threadIdx.x >> 5); base < fp.workEnd; base += 32u * warpsTotal) {
      uint32_t myDraw = 0, myMeshIdx = 0;
      bool vis = false;
      {
        const uint64_t cand64 = (uint64_t)base + (uint64_t)lane * warpsTotal;
        if (cand64 < fp.workEnd) {
            const uint32_t cand = (uint32_t)cand64;
            const DrawItem& dc = find_draw(draws, numDraws, cand);
            myDraw = (uint32_t)(&dc - draws);
            myMeshIdx = cand - dc.firstWork;
            vis = true;
            if (dc.cullBitmap != nullptr) {                                         // ShadeMeshlet: cull bit (Shading.cpp:282-289)
                const uint32_t word = dc.cullBitmap[myMeshIdx >> 4];
                vis = ((word >> (myMeshIdx & 15u)) & 1u) != 0;
            }
            if (vis && dc.fusedCull) {                                              // Shading.cpp:803-809
                const uint4 hdrA = __ldg(reinterpret_cast<const uint4*>(meshlets + (dc.meshletOffset + myMeshIdx)));   // BoundCenter, BoundRadius
                const float cx = __uint_as_float(hdrA.x), cy = __uint_as_float(hdrA.y), cz = __uint_as_float(hdrA.z);
                const float rad = __uint_as_float(hdrA.w);
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const float dist = __fadd_rn(__fmaf_rn(cx, dc.planes[i][0], __fmaf_rn(cy, dc.planes[i][1], __fmul_rn(cz, dc.planes[i][2]))), dc.planes[i][3]);
                    vis = vis && (dist > -rad);
                }
            }
            if (kBand && vis) {
                // scissor rows: the band is the slab bandNdcLo <= y/w <= bandNdcHi, i.e. the object-space half-spaces
                // (row_y - lo * row_w) . p >= 0 and (hi * row_w - row_y) . p >= 0 of this draw's ObjectToClip; a meshlet whose
                // bound sphere lies wholly outside one of them has no pixel in the band (conservative: the slab is a pixel row wider)
                const float* Mc = fp.uniformMatrix ? fp.M : dc.M;
                const uint4 hdrA = __ldg(reinterpret_cast<const uint4*>(meshlets + (dc.meshletOffset + myMeshIdx)));
                const float cx = __uint_as_float(hdrA.x), cy = __uint_as_float(hdrA.y), cz = __uint_as_float(hdrA.z);
                const float rad = __uint_as_float(hdrA.w) * 1.0001f;
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    const float k = side ? fp.bandNdcHi : fp.bandNdcLo, sgn = side ? -1.0f : 1.0f;
                    const float a = sgn * (Mc[1] - k * Mc[3]), b = sgn * (Mc[5] - k * Mc[7]), c = sgn * (Mc[9] - k * Mc[11]), dd = sgn * (Mc[13] - k * Mc[15]);
                    const float len = sqrtf(a * a + b * b + c * c);
                    vis = vis && (a * cx + b * cy + c * cz + dd > -rad * len);
                }
            }
        }
      }
      uint32_t alive = __ballot_sync(0xFFFFFFFFu, vis);
      while (alive) {
        const uint32_t src = (uint32_t)__ffs(alive) - 1u;
        alive &= alive - 1u;
        const DrawItem& d = draws[__shfl_sync(0xFFFFFFFFu, myDraw, src)];
        const uint32_t meshIdx = __shfl_sync(0xFFFFFFFFu, myMeshIdx, src);
        const swr_meshlet* m = meshlets + (d.meshletOffset + meshIdx);
        // issue every load of the meshlet before anything depends on them
        const uint4 hdrB = __ldg(reinterpret_cast<const uint4*>(m) + 2);        // bytes 32..47: ..., NumVertices, NumTriangles, AlphaCutoff
        const uint32_t materialId = __ldg(reinterpret_cast<const uint32_t*>(m) + 12);
        uint4 idxWord = make_uint4(0, 0, 0, 0);
        if (lane < 24) idxWord = __ldg(reinterpret_cast<const uint4*>(m->Indices) + lane);   // Shading.cpp:300
        float px[2], py[2], pz[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t v = lane + h * 32;
            px[h] = __ldg(&m->Positions[0][v]); py[h] = __ldg(&m->Positions[1][v]); pz[h] = __ldg(&m->Positions[2][v]);
        }
        float M[16];                    // one matrix for the whole batch rides in the kernel parameters (constant bank)
        if (fp.uniformMatrix) {
#pragma unroll
            for (int i = 0; i < 16; i++) M[i] = fp.M[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) M[i] = d.M[i];
        }

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
        if (fp.program != 0u) fsId = fp.program == SWRB_PROGRAM_DEFERRED ? 1u : 0u;   // one fragment program in every slot; FS_EncodeGBuffer's records carry 1/w like the alpha program's
        if (lane < 24) reinterpret_cast<uint4*>(s.idx)[lane] = idxWord;

        // ---- transform + per-vertex setup: lane owns vertices lane and lane+32
        const uint32_t vertSlots = min((numVerts + 15u) & ~15u, 64u);           // reference walks 16-wide vectors
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t v = lane + h * 32;
            if (v < vertSlots) {
                float x = px[h], y = py[h], z = pz[h];
                // simd::mul(mat4, (pos,1)) — SIMD.h:457-464
                float cx = __fmaf_rn(x, M[0], __fmaf_rn(y, M[4], __fmaf_rn(z, M[8], M[12])));
                float cy = __fmaf_rn(x, M[1], __fmaf_rn(y, M[5], __fmaf_rn(z, M[9], M[13])));
                float cz = __fmaf_rn(x, M[2], __fmaf_rn(y, M[6], __fmaf_rn(z, M[10], M[14])));
                float cw = __fmaf_rn(x, M[3], __fmaf_rn(y, M[7], __fmaf_rn(z, M[11], M[15])));
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
                // per-vertex x/w, y/w, 1/w for this frame's resolve pass (IntersectTriangle re-derives exactly these)
                if (clipCache != nullptr) clipCache[(size_t)(d.meshletOffset + meshIdx) * SWR_MAX_VERTICES + v] = make_float4(nx, ny, rw, nz);
            }
        }
        __syncwarp();

        // ---- triangles: lane owns prims lane + 32k; small ones are rasterized right here
        const uint8_t* idx = reinterpret_cast<const uint8_t*>(s.idx);
        const uint32_t curMeshlet = d.meshletOffset + meshIdx, curDraw = (uint32_t)(&d - draws);
        const uint32_t rankBase = curMeshlet << 8;            // key_rank(meshlet, prim, accepted lane): draw order inside the batch
        uint32_t numBig = 0;
#pragma unroll 1
        for (uint32_t k = 0; k < 4 && k * 32 < primCount; k++) {
            const uint32_t prim = lane + k * 32;
            bool keep = false, nonTrivial = false, big = false;
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
                        BBox r;
                        if (raster_region<kBand>(p0, p1, p2, fp, r)) {     // else: counted, touches no pixel
                            const int32_t area = (r.maxX - r.minX) * (r.maxY - r.minY);
                            if (fsId == 0 && (uint32_t)area <= fp.inlineMaxArea && fp.program == 0u) {
                                TriRecord t;
                                t.pos0 = p0; t.pos1 = p1; t.pos2 = p2;
                                t.z0 = s.z[i0]; t.z1 = s.z[i1]; t.z2 = s.z[i2];
                                t.id = rankBase | ((prim >> 4) << 5) | (prim & 15u); t.aux = 0;
                                raster_inline(t, r, fp, keys);
                            } else {
                                big = true;
                            }
                        }
                    }
                }
            }
            const uint32_t keepMask = __ballot_sync(0xFFFFFFFFu, keep);
            const uint32_t clipMask = __ballot_sync(0xFFFFFFFFu, nonTrivial);
            const uint32_t bigMask = __ballot_sync(0xFFFFFFFFu, big);
            const uint32_t lt = (1u << lane) - 1u;
            if (big) s.big[numBig + __popc(bigMask & lt)] = (uint8_t)prim;
            numBig += __popc(bigMask);
            if (lane == 0) { nRasterized += __popc(keepMask); nClipped += fp.clipMode != 1u ? __popc(clipMask) : 0; }   // :579, :568 / :210
            if (!kBinned && fp.clipMode == 2u && clipMask) {
                // EnableClipping on the unbinned path (Rasterizer.cpp:209-249): name the triangle in the clip list;
                // k_clip_triangles re-derives its clip-space vertices, clips and appends the pieces as records
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&ctl->clipCount, (uint32_t)__popc(clipMask));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (nonTrivial) {
                    const uint32_t slot = base + __popc(clipMask & lt);
                    if (slot < triCapacity) clipList[slot] = make_uint2(curDraw, (meshIdx << 7) | prim);
                    else atomicExch(&ctl->overflow, 1u);
                }
            }
        }
        if (numBig) __syncwarp();

        // ---- big triangles: one slice of the record array per meshlet
        if (numBig && fsId) {
            // alpha-tested meshlet (FragmentShaderId 1): every surviving triangle goes to the alpha list with the
            // 1/w of its vertices; k_raster_alpha runs the textured fragment program on them
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->alphaCount, numBig);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            const bool fits = base + numBig <= triCapacity && alphaTris != nullptr;
            if (!fits && lane == 0) atomicExch(&ctl->overflow, 1u);
            uint32_t matWord = 0x00FFFFFFu;                       // TextureId | AlphaCutoff << 24 for k_raster_alpha (FS_EncodeGBuffer reads the material itself)
            if (materialId != SWR_NO_MATERIAL && materials != nullptr) {
                const swr_material mat = materials[materialId];
                matWord = ((uint32_t)mat.TextureId & 0x00FFFFFFu) | ((uint32_t)mat.AlphaCutoff << 24);
            }
            for (uint32_t j = lane; fits && j < numBig; j += 32) {
                const uint32_t prim = s.big[j];
                uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                uint4* dst = reinterpret_cast<uint4*>(alphaTris + base + j);
                dst[0] = make_uint4(s.pos[i0], s.pos[i1], s.pos[i2], __float_as_uint(s.z[i0]));
                dst[1] = make_uint4(__float_as_uint(s.z[i1]), __float_as_uint(s.z[i2]), rankBase | ((prim >> 4) << 5) | (prim & 15u), 1u);
                uint4* dw = reinterpret_cast<uint4*>(alphaW + base + j);
                dw[0] = make_uint4(__float_as_uint(s.rw[i0]), __float_as_uint(s.rw[i1]), __float_as_uint(s.rw[i2]), curDraw);   // .w: the draw (DeferredShader reads its ObjectToWorld)
                dw[1] = make_uint4(__ldg(&m->TexCoords[i0]), __ldg(&m->TexCoords[i1]), __ldg(&m->TexCoords[i2]), matWord);
            }
        } else if (numBig) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&ctl->triCount, numBig);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            const bool fits = base + numBig <= triCapacity;
            if (!fits && lane == 0) atomicExch(&ctl->overflow, 1u);
            for (uint32_t j0 = 0; j0 < numBig; j0 += 32) {
                const uint32_t j = j0 + lane;
                const bool active = fits && j < numBig;
                uint32_t tx0 = 1, ty0 = 1, tx1 = 0, ty1 = 0;
                if (active) {
                    const uint32_t prim = s.big[j];
                    uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                    uint32_t p0 = s.pos[i0], p1 = s.pos[i1], p2 = s.pos[i2];
                    uint4* dst = reinterpret_cast<uint4*>(tris + base + j);
                    dst[0] = make_uint4(p0, p1, p2, __float_as_uint(s.z[i0]));
                    dst[1] = make_uint4(__float_as_uint(s.z[i1]), __float_as_uint(s.z[i2]), rankBase | ((prim >> 4) << 5) | (prim & 15u), 0u);
                    if (kBinned) {
                        BBox r;
                        raster_region<kBand>(p0, p1, p2, fp, r);
                        tx0 = (uint32_t)(r.minX >> kTileShift); ty0 = (uint32_t)(r.minY >> kTileShift);
                        tx1 = (uint32_t)((r.maxX - 1) >> kTileShift); ty1 = (uint32_t)((r.maxY - 1) >> kTileShift);
                    }
                }
                if (kBinned) {   // pass 1 of the binner: per-tile counts
                    const uint32_t nTiles = (active && tx0 <= tx1 && ty0 <= ty1) ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
                    count_single_tile(tileCount, nTiles == 1, ty0 * fp.tilesX + tx0);
                    if (nTiles > (uint32_t)kBigTriTileLimit) {     // wide triangles are counted per 256-px super-tile
                        const uint32_t sh = kSuperShift - kTileShift, superX = (fp.tilesX + (1u << sh) - 1u) >> sh;
                        for (uint32_t sy = ty0 >> sh; sy <= (ty1 >> sh); sy++)
                            for (uint32_t sx = tx0 >> sh; sx <= (tx1 >> sh); sx++) atomicAdd(&out.superCount[sy * superX + sx], 1u);
                        atomicAdd(&ctl->bigCount, 1u);
                    } else if (nTiles > 1) {
                        for (uint32_t ty = ty0; ty <= ty1; ty++)
                            for (uint32_t tx = tx0; tx <= tx1; tx++) atomicAdd(&tileCount[ty * fp.tilesX + tx], 1u);
                    }
                }
            }
        }
        __syncwarp();
      }
    }

    // ---- perf counters: one atomic per warp per counter (Rasterizer.cpp:927-932 FlushThreadCounters)
    if (lane == 0) {
        if (nProcessed) atomicAdd(&ctl->perf[0], (unsigned long long)nProcessed);
        if (nRasterized) atomicAdd(&ctl->perf[1], (unsigned long long)nRasterized);
        if (nClipped) atomicAdd(&ctl->perf[2], (unsigned long long)nClipped);
    }
}

}  // namespace swrb
