// mesh.cuh — K1: meshlet cull-bit/frustum test + mesh shading + clip classification + triangle setup
// + rasterization of every triangle that is not "big", one warp per meshlet.
//
// Replaces (reference, one worker iteration of Rasterizer::DrawMeshlets, Rasterizer.cpp:535-594):
//   ShadeMeshlet                 Shading.cpp:281-307   (cull bit, SoA position transform, index copy)
//   CullMeshlets frustum test    Shading.cpp:803-809   (optional fused form)
//   GatherPos / index widen      Rasterizer.cpp:143-151, :553-558
//   Clipper::ComputeClipCodes    Rasterizer.cpp:353-397
//   TrianglePacket::Setup + bbox Rasterizer.cpp:257-289, :331-351
// and, for triangles whose pixel region is <= kInlineMaxArea, also
//   TriangleEdgeVars::Setup      Rasterizer.cpp:296-329
//   DrawTriangle<> + FS_EncodeSurfaceId<false>   Rasterizer.h:250-328, Shading.cpp:309-331
//
// Work distribution: a persistent grid, two phases in one launch (see k_mesh_setup): a cull phase in which each LANE tests
// one meshlet (cull bit, bound sphere against the five frustum planes) and the survivors are appended to a global visible
// list, then a shade phase in which every warp takes one visible meshlet at a time from a device-side cursor.
//
// Staging: the 1,216 hot bytes of a surviving meshlet (header + Positions[3][64], Indices[3][128]) are brought into
// shared memory by two bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) issued by one lane;
// the copy of survivor k+1 is in flight while the warp works on survivor k (two stages per warp), so HBM latency is
// off the warp's critical path without costing a single register.
//
// Per meshlet: per-VERTEX perspective divide / snap / outcodes are computed once (lane owns vertices L and L+32; the CPU
// recomputes them per triangle corner; same inputs -> same bits) and parked in shared memory as one 16-byte record;
// the triangles are then CLASSIFIED (lane owns prims L + 32k: visibility, guard band, determinant, bounding box) and
// the survivors compacted by ballot into two lists:
//   * small (pixel region <= kInlineMaxArea): rasterized right here, 32 survivors at a time with every lane busy
//     (a meshlet that keeps 40 of 98 triangles walks 2 rounds, not 4 with holes), 64-bit depth|rank keys reduced
//     straight into the frame's key buffer with REDG.MAX.64 — no record, no bin entry, no second kernel;
//   * big: one slice of the global record array per meshlet (one atomic), 32-byte records, and on the binned path a
//     per-tile (or, over kBigTriTileLimit tiles, per-super-tile) count = pass 1 of the binner.
// (A warp-cooperative middle class was measured and dropped in round 1: one warp working through a meshlet's ~100
// mid-size triangles one after another is a 100+ us serial tail; lanes in parallel or tiles in parallel win.)
#pragma once

#include "common.cuh"

namespace swrb {

constexpr int kMeshWarps = 8;           // warps (= meshlets in flight) per block
constexpr int kInlineMaxArea = 256;     // largest pixel region a lane may rasterize itself (FrameParams::inlineMaxArea <= this; else: record + binner)
constexpr uint32_t kMeshStageBytes = 1216;
constexpr float kLargeMeshletPx = 24.0f; // projected bound-sphere radius from which a meshlet is scheduled first (see phase A)

struct __align__(16) MeshStage {        // the hot bytes of one swr_meshlet, as the bulk copies land them
    uint32_t hdr[16];                   // bytes 0..63: bounds, cone, NumVertices/NumTriangles/AlphaCutoff @44, MaterialId @48
    float pos[3][64];                   // bytes 64..831: Positions
    uint32_t idx[96];                   // bytes 1344..1727: Indices[3][128]
};
static_assert(sizeof(MeshStage) == kMeshStageBytes, "stage layout");

struct __align__(16) MeshWarpSmem {
    MeshStage stage[2];
    float4 vert[64];                    // { x/w, y/w, z/w, bits(packed 28.4 x | y << 16) }
    float rw[64];                       // 1/w
    uint32_t flags[64];                 // bits 0-5 Cohen-Sutherland outcodes, bit 6 inside guard band
    uint2 small[128];                   // inline-raster survivors: { minX | minY << 16, (w-1) | (h-1) << 8 | prim << 16 }
    uint8_t big[128];                   // prims that need a record
    unsigned long long bar[2];          // one mbarrier per stage
};

// ---- async bulk copy (TMA unit, 1-D) + mbarrier, sm_90+ PTX ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// One lane starts the copy of meshlet `m` into `st`; everything the warp read from `st` before is ordered first.
__device__ __forceinline__ void stage_issue(MeshStage* st, unsigned long long* bar, const swr_meshlet* m) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, kMeshStageBytes);
    bulk_g2s(st->hdr, m, 832u, bar);                                                        // header + Positions
    bulk_g2s(st->idx, reinterpret_cast<const char*>(m) + 1344, 384u, bar);                  // Indices
}

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

// Spin until *p >= want (device scope). Used for the one grid-wide hand-over of the kernel (cull phase -> shade phase).
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool kBinned>
// Persistent grid of 1..4 blocks per SM (swrb_device_set_mesh_occupancy): 4 = the whole register file for a lone frame,
// 1 leaves room for other render contexts' resolve blocks, whose issue-bound warps fill what these latency-bound ones leave idle.
//
// Two phases inside the one launch:
//   A  cull   (only when the batch has cull bitmaps or fused frustum planes): the warps walk the work items 32 at a time,
//             one LANE per meshlet (cull bit, bound sphere against the five planes — a culled meshlet costs a lane, not a
//             warp), and append the survivors {meshlet, draw} to a global visible list (ballot + one atomic per warp step);
//   B  shade  every warp takes ONE visible meshlet at a time from a device-side cursor — a meshlet of large triangles can
//             cost 10x the average, so anything coarser leaves a tail — and always has the NEXT one's bulk copy in flight.
// The phases meet at a grid-wide hand-over: a warp that finished its share of A waits until all of A is published
// (ctl->cullDone). A is a few microseconds and every block of the persistent grid is resident or becomes resident without
// our help (other kernels never wait for this one), so the wait cannot deadlock.
__global__ void __launch_bounds__(kMeshWarps * 32, 4)      // 64 registers: up to 4 blocks = 32 warps per SM
k_mesh_setup(const swr_meshlet* __restrict__ meshlets, const swr_material* __restrict__ materials,
             const DrawItem* __restrict__ draws, uint32_t numDraws, uint32_t totalWork, uint2* __restrict__ visList, FrameParams fp,
             unsigned long long* __restrict__ keys, MeshOut out, DevCtl* __restrict__ ctl) {
    __shared__ MeshWarpSmem smem[kMeshWarps];
    MeshWarpSmem& s = smem[threadIdx.x >> 5];
    const uint32_t lane = lane_id();
    if (ctl->overflow) return;           // an earlier draw since the host last looked has aborted: every later draw is predicated off
    if (lane == 0) {
        mbar_init(&s.bar[0], 1u);
        mbar_init(&s.bar[1], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t nProcessed = 0, nRasterized = 0, nClipped = 0;
    const uint32_t warpGlobal = blockIdx.x * kMeshWarps + (threadIdx.x >> 5), warpsTotal = gridDim.x * kMeshWarps;
    uint32_t numItems = fp.workEnd - fp.workBegin, numFront = 0;   // visList == null: no culling anywhere in the batch, item i IS work item i

    if (visList != nullptr) {
        // ---- phase A: cull. ShadeMeshlet's cull bit (Shading.cpp:282-289) + CullMeshlets' frustum test (:803-809)
        const uint32_t numChunks = (fp.workEnd - fp.workBegin + 31u) >> 5;
        uint32_t myChunks = 0;
        for (uint32_t chunk = warpGlobal; chunk < numChunks; chunk += warpsTotal, myChunks++) {
            const uint32_t work = fp.workBegin + chunk * 32u + lane;
            bool vis = work < fp.workEnd, large = false;
            uint32_t dIdx = 0, meshletId = 0;
            uint32_t lo = 0, hi = numDraws;                     // warp-uniform: last draw with firstWork <= the chunk's first item
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (draws[mid].firstWork <= fp.workBegin + chunk * 32u) lo = mid; else hi = mid;
            }
            dIdx = lo;
            if (vis) {
                while (dIdx + 1 < numDraws && draws[dIdx + 1].firstWork <= work) dIdx++;
                const DrawItem& d = draws[dIdx];
                const uint32_t meshIdx = work - d.firstWork;
                meshletId = d.meshletOffset + meshIdx;
                if (d.cullBitmap != nullptr) {
                    uint32_t word = d.cullBitmap[meshIdx >> 4];
                    vis = ((word >> (meshIdx & 15u)) & 1u) != 0;
                }
                float cx = 0.0f, cy = 0.0f, cz = 0.0f, rad = 0.0f;
                if (vis) {
                    const uint4 hdrA = __ldg(reinterpret_cast<const uint4*>(meshlets + meshletId));        // BoundCenter, BoundRadius
                    cx = __uint_as_float(hdrA.x); cy = __uint_as_float(hdrA.y); cz = __uint_as_float(hdrA.z);
                    rad = __uint_as_float(hdrA.w);
                }
                if (vis && d.fusedCull) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        float dist = __fadd_rn(__fmaf_rn(cx, d.planes[i][0], __fmaf_rn(cy, d.planes[i][1], __fmul_rn(cz, d.planes[i][2]))), d.planes[i][3]);
                        vis = vis && (dist > -rad);
                    }
                }
                if (vis) {
                    // Scheduling hint only (no effect on any result): does the bound sphere project to more than ~kLargeMeshletPx
                    // pixels of radius? Such a meshlet's triangles have pixel regions of tens of pixels and the meshlet costs
                    // several times the average, so it goes to the front of the list and is started first.
                    const float* M = fp.uniformMatrix ? fp.M : d.M;
                    const float cw = M[3] * cx + M[7] * cy + M[11] * cz + M[15];
                    const float sy = sqrtf(M[1] * M[1] + M[5] * M[5] + M[9] * M[9]);
                    large = !(rad * sy * (float)fp.halfH <= kLargeMeshletPx * cw);
                }
            }
            const uint32_t alive = __ballot_sync(0xFFFFFFFFu, vis);
            if (alive) {
                const uint32_t front = __ballot_sync(0xFFFFFFFFu, vis && large), back = alive & ~front;
                uint32_t baseF = 0, baseB = 0;
                if (lane == 0) {
                    if (front) baseF = atomicAdd(&ctl->visCount, (uint32_t)__popc(front));
                    if (back) baseB = atomicAdd(&ctl->visCountBack, (uint32_t)__popc(back));
                }
                baseF = __shfl_sync(0xFFFFFFFFu, baseF, 0);
                baseB = __shfl_sync(0xFFFFFFFFu, baseB, 0);
                const uint32_t lt = (1u << lane) - 1u;
                if (vis) visList[large ? baseF + __popc(front & lt) : totalWork - 1u - (baseB + __popc(back & lt))] = make_uint2(meshletId, dIdx);
            }
        }
        // publish: this warp's entries, then its share of the chunk count; wait until every chunk has been published
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            if (myChunks) atomicAdd(&ctl->cullDone, myChunks);
            while (ld_acquire_gpu(&ctl->cullDone) < numChunks) __nanosleep(200);
        }
        __syncwarp();
        numFront = ld_acquire_gpu(&ctl->visCount);
        numItems = numFront + ld_acquire_gpu(&ctl->visCountBack);
    }

    // ---- phase B: shade. The warps take the visible meshlets in small runs of consecutive items: the first run is the
    // warp's own (no atomic), the later ones come from a device-side cursor and shrink to single meshlets towards the end of
    // the list (a meshlet of large triangles costs 10x the average; the list starts with those), and the next item's bytes
    // are always in flight.
    uint32_t slot = 0, parity = 0;       // bit k of `parity`: phase the next wait on stage k expects
    const uint32_t firstRun = min(max(numItems / (warpsTotal * 4u), 1u), 4u);
    uint32_t itemNext = warpGlobal * firstRun, itemEnd = min(itemNext + firstRun, numItems), seen = warpsTotal * firstRun;
    auto take_item = [&](uint32_t& meshletId, uint32_t& drawIdx) -> bool {       // warp-uniform
        if (itemNext >= itemEnd) {
            const uint32_t left = numItems > seen ? numItems - seen : 0u;      // as of this warp's last look at the cursor
            const uint32_t run = min(max(left / (warpsTotal * 2u), 1u), 4u);
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(&ctl->workCursor, run);
            b = __shfl_sync(0xFFFFFFFFu, b, 0) + warpsTotal * firstRun;
            seen = b + run;
            if (b >= numItems) return false;
            itemNext = b; itemEnd = min(b + run, numItems);
        }
        const uint32_t item = itemNext++;
        if (visList != nullptr) {
            const uint2 e = __ldcg(visList + (item < numFront ? item : totalWork - 1u - (item - numFront)));
            meshletId = e.x; drawIdx = e.y;
        } else {
            const uint32_t work = fp.workBegin + item;
            uint32_t lo = 0, hi = numDraws;
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (draws[mid].firstWork <= work) lo = mid; else hi = mid;
            }
            drawIdx = lo;
            meshletId = draws[lo].meshletOffset + (work - draws[lo].firstWork);
        }
        return true;
    };
    uint32_t curMeshlet = 0, curDraw = 0, nxtMeshlet = 0, nxtDraw = 0;
    bool have = take_item(curMeshlet, curDraw);
    if (have && lane == 0) stage_issue(&s.stage[slot], &s.bar[slot], meshlets + curMeshlet);
    while (have) {
        {
            const bool haveNext = take_item(nxtMeshlet, nxtDraw);
            if (haveNext && lane == 0) stage_issue(&s.stage[slot ^ 1u], &s.bar[slot ^ 1u], meshlets + nxtMeshlet);   // travels while this one is shaded
            mbar_wait(&s.bar[slot], (parity >> slot) & 1u);
            parity ^= 1u << slot;
            have = haveNext;
        }
        {
            const MeshStage& st = s.stage[slot];
            slot ^= 1u;
            const DrawItem& d = draws[curDraw];

            const uint32_t hdrW = st.hdr[11], materialId = st.hdr[12];
            const uint32_t numVerts = hdrW & 0xFFu, numTris = (hdrW >> 8) & 0xFFu;
            const uint32_t primCount = min(numTris, 128u);
            if (primCount != 0) {
                if (lane == 0) nProcessed += primCount;                                 // Rasterizer.cpp:545
                uint32_t cullMode = SWR_CULL_FRONT_CCW, fsId = 0;                       // Shading.cpp:302-306 (+ SURVEY App. B.4)
                if (materialId != SWR_NO_MATERIAL && materials != nullptr) {
                    swr_material mat = materials[materialId];
                    cullMode = mat.IsDoubleSided ? SWR_CULL_NONE : SWR_CULL_FRONT_CCW;
                    fsId = mat.AlphaCutoff < 255 ? 1u : 0u;
                }
                if (fp.program != 0u) fsId = fp.program == SWRB_PROGRAM_DEFERRED ? 1u : 0u;   // one fragment program in every slot; FS_EncodeGBuffer needs 1/w per vertex like the alpha program: same record list

                // ---- transform + per-vertex setup: lane owns vertices lane and lane+32
                {
                    float M[16];            // one matrix for the whole batch rides in the kernel parameters (constant bank)
                    if (fp.uniformMatrix) {
#pragma unroll
                        for (int i = 0; i < 16; i++) M[i] = fp.M[i];
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++) M[i] = d.M[i];
                    }
                    const uint32_t vertSlots = min((numVerts + 15u) & ~15u, 64u);       // reference walks 16-wide vectors
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t v = lane + h * 32;
                        if (v < vertSlots) {
                            const float x = st.pos[0][v], y = st.pos[1][v], z = st.pos[2][v];
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
                            s.vert[v] = make_float4(nx, ny, nz, __uint_as_float(((uint32_t)X & 0xFFFFu) | ((uint32_t)Y << 16)));
                            s.rw[v] = rw;
                            s.flags[v] = f;
                            // per-vertex x/w, y/w, 1/w for this frame's resolve pass (IntersectTriangle re-derives exactly these)
                            if (out.clipCache != nullptr) out.clipCache[(size_t)curMeshlet * SWR_MAX_VERTICES + v] = make_float4(nx, ny, rw, nz);
                        }
                    }
                }
                __syncwarp();

                // ---- classify: lane owns prims lane + 32k; survivors are compacted into the small / big lists
                const uint8_t* idx = reinterpret_cast<const uint8_t*>(st.idx);
                const uint32_t rankBase = curMeshlet << 8;
                const uint32_t lt = (1u << lane) - 1u;
                uint32_t numSmall = 0, numBig = 0;
#pragma unroll 1
                for (uint32_t k = 0; k * 32 < primCount; k++) {
                    const uint32_t prim = lane + k * 32;
                    bool keep = false, nonTrivial = false, big = false, small = false;
                    uint2 ent = make_uint2(0u, 0u);
                    if (prim < primCount) {
                        const uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                        const uint32_t f0 = s.flags[i0], f1 = s.flags[i1], f2 = s.flags[i2];
                        const uint32_t partial = f0 | f1 | f2, combined = f0 & f1 & f2;
                        const bool visible = (combined & 63u) == 0;                             // Rasterizer.cpp:389
                        const bool trivial = (combined & 64u) != 0 && (partial & 48u) == 0;     // :386-388
                        nonTrivial = visible && !trivial;                                       // :393
                        if (visible && trivial) {
                            // TrianglePacket::Setup (Rasterizer.cpp:257-289)
                            const float4 a = s.vert[i0], b = s.vert[i1], c = s.vert[i2];
                            float det = __fsub_rn(__fmul_rn(__fsub_rn(c.x, a.x), __fsub_rn(b.y, a.y)),
                                                  __fmul_rn(__fsub_rn(a.x, b.x), __fsub_rn(a.y, c.y)));
                            if (cullMode != SWR_CULL_FRONT_CCW) {
                                bool flip = (cullMode == SWR_CULL_FRONT_CW) ? true : (det < 0.0f);
                                det = flip ? -det : det;
                            }
                            const uint32_t p0 = __float_as_uint(a.w), p1 = __float_as_uint(b.w), p2 = __float_as_uint(c.w);
                            uint32_t bbMin, bbMax;
                            ref_render_bbox(p0, p1, p2, fp.halfW, fp.halfH, bbMin, bbMax);
                            keep = det > 0.0f && lo16(bbMin) < lo16(bbMax) && hi16(bbMin) < hi16(bbMax);   // :269, :283
                            if (keep) {
                                BBox r;
                                if (raster_region(p0, p1, p2, fp.halfW, fp.halfH, r)) {     // else: counted, touches no pixel
                                    const int32_t w = r.maxX - r.minX, h = r.maxY - r.minY;
                                    if (fsId == 0 && (uint32_t)(w * h) <= fp.inlineMaxArea && fp.program == 0u) {
                                        small = true;
                                        ent = make_uint2((uint32_t)r.minX | ((uint32_t)r.minY << 16),
                                                         (uint32_t)(w - 1) | ((uint32_t)(h - 1) << 8) | (prim << 16));
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
                    const uint32_t smallMask = __ballot_sync(0xFFFFFFFFu, small);
                    if (small) s.small[numSmall + __popc(smallMask & lt)] = ent;
                    numSmall += __popc(smallMask);
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
                            const uint32_t slotC = base + __popc(clipMask & lt);
                            if (slotC < out.triCapacity) out.clipList[slotC] = make_uint2(curDraw, ((curMeshlet - d.meshletOffset) << 7) | prim);
                            else atomicExch(&ctl->overflow, 1u);
                        }
                    }
                }
                __syncwarp();

                // ---- inline raster: every lane takes one small survivor per round
                for (uint32_t j = lane; j < numSmall; j += 32) {
                    const uint2 ent = s.small[j];
                    const uint32_t prim = ent.y >> 16;
                    const uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                    const float4 a = s.vert[i0], b = s.vert[i1], c = s.vert[i2];
                    TriRecord t;
                    t.pos0 = __float_as_uint(a.w); t.pos1 = __float_as_uint(b.w); t.pos2 = __float_as_uint(c.w);
                    t.z0 = a.z; t.z1 = b.z; t.z2 = c.z;
                    t.id = rankBase | ((prim >> 4) << 5) | (prim & 15u); t.aux = 0;
                    BBox r;
                    r.minX = (int32_t)(ent.x & 0xFFFFu); r.minY = (int32_t)(ent.x >> 16);
                    r.maxX = r.minX + (int32_t)(ent.y & 0xFFu) + 1; r.maxY = r.minY + (int32_t)((ent.y >> 8) & 0xFFu) + 1;
                    raster_inline(t, r, fp, keys);
                }

                // ---- big triangles: one slice of the record array per meshlet
                if (numBig && fsId) {
                    // alpha-tested meshlet (FragmentShaderId 1): every surviving triangle goes to the alpha list with the
                    // 1/w of its vertices; k_raster_alpha runs the textured fragment program on them
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&ctl->alphaCount, numBig);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    const bool fits = base + numBig <= out.triCapacity && out.alphaTris != nullptr;
                    if (!fits && lane == 0) atomicExch(&ctl->overflow, 1u);
                    for (uint32_t j = lane; fits && j < numBig; j += 32) {
                        const uint32_t prim = s.big[j];
                        const uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                        const float4 a = s.vert[i0], b = s.vert[i1], c = s.vert[i2];
                        uint4* dst = reinterpret_cast<uint4*>(out.alphaTris + base + j);
                        dst[0] = make_uint4(__float_as_uint(a.w), __float_as_uint(b.w), __float_as_uint(c.w), __float_as_uint(a.z));
                        dst[1] = make_uint4(__float_as_uint(b.z), __float_as_uint(c.z), rankBase | ((prim >> 4) << 5) | (prim & 15u), 1u);
                        *reinterpret_cast<float4*>(out.alphaW + base + j) = make_float4(s.rw[i0], s.rw[i1], s.rw[i2], __uint_as_float(curDraw));   // .w: the draw (DeferredShader reads its ObjectToWorld)
                    }
                } else if (numBig) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&ctl->triCount, numBig);
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    const bool fits = base + numBig <= out.triCapacity;
                    if (!fits && lane == 0) atomicExch(&ctl->overflow, 1u);
                    for (uint32_t j0 = 0; j0 < numBig; j0 += 32) {
                        const uint32_t j = j0 + lane;
                        const bool active = fits && j < numBig;
                        uint32_t tx0 = 1, ty0 = 1, tx1 = 0, ty1 = 0;
                        if (active) {
                            const uint32_t prim = s.big[j];
                            const uint32_t i0 = idx[prim] & 63u, i1 = idx[128 + prim] & 63u, i2 = idx[256 + prim] & 63u;
                            const float4 a = s.vert[i0], b = s.vert[i1], c = s.vert[i2];
                            const uint32_t p0 = __float_as_uint(a.w), p1 = __float_as_uint(b.w), p2 = __float_as_uint(c.w);
                            uint4* dst = reinterpret_cast<uint4*>(out.tris + base + j);
                            dst[0] = make_uint4(p0, p1, p2, __float_as_uint(a.z));
                            dst[1] = make_uint4(__float_as_uint(b.z), __float_as_uint(c.z), rankBase | ((prim >> 4) << 5) | (prim & 15u), 0u);
                            if (kBinned) {
                                BBox r;
                                raster_region(p0, p1, p2, fp.halfW, fp.halfH, r);
                                tx0 = (uint32_t)(r.minX >> kTileShift); ty0 = (uint32_t)(r.minY >> kTileShift);
                                tx1 = (uint32_t)((r.maxX - 1) >> kTileShift); ty1 = (uint32_t)((r.maxY - 1) >> kTileShift);
                            }
                        }
                        if (kBinned) {   // pass 1 of the binner: per-tile counts; wide triangles are counted per 256-px super-tile
                            const uint32_t nTiles = (active && tx0 <= tx1 && ty0 <= ty1) ? (tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
                            count_single_tile(out.tileCount, nTiles == 1, ty0 * fp.tilesX + tx0);
                            if (nTiles > (uint32_t)kBigTriTileLimit) {
                                const uint32_t sh = kSuperShift - kTileShift, superX = (fp.tilesX + (1u << sh) - 1u) >> sh;
                                for (uint32_t sy = ty0 >> sh; sy <= (ty1 >> sh); sy++)
                                    for (uint32_t sx = tx0 >> sh; sx <= (tx1 >> sh); sx++) atomicAdd(&out.superCount[sy * superX + sx], 1u);
                                atomicAdd(&ctl->bigCount, 1u);
                            } else if (nTiles > 1) {
                                for (uint32_t ty = ty0; ty <= ty1; ty++)
                                    for (uint32_t tx = tx0; tx <= tx1; tx++) atomicAdd(&out.tileCount[ty * fp.tilesX + tx], 1u);
                            }
                        }
                    }
                }
            }
            __syncwarp();    // every lane is done with this stage and the per-vertex records before either is overwritten
            curMeshlet = nxtMeshlet; curDraw = nxtDraw;
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
