// clip.cuh — polygon clipper of the unbinned path: Clipper::ClipTriangles + FlushPacket
// (Rasterizer.cpp:398-491), one thread per non-trivial triangle.
//
// The mesh kernel only names non-trivial triangles ({draw, meshlet, prim}, 8 bytes); this kernel re-derives
// their clip-space vertices with the same FMA chain (same inputs -> same bits), runs Sutherland-Hodgman
// against the planes in the triangle's partial outcodes (x/y planes on the guard band, z planes on the
// frustum, ascending plane id like BitIter), fan-triangulates, and pushes every piece through
// TrianglePacket::Setup (perspective divide, determinant cull, 28.4 snap, empty-bbox cull). Survivors are
// appended to the ordinary record list — k_raster_direct / k_raster_big rasterize them with the original
// primitive's surface id (DrawTriangle<FS, true> takes PrimId from ClipData, Rasterizer.h:272) — or, for
// alpha-tested materials, to the alpha list together with 1/w and the ClippedU/ClippedV barycentric remap
// (Rasterizer.h:312-318). Non-trivial triangles are a fraction of a percent of a scene (those crossing the
// near plane or leaving the 2896-px guard band), so a scalar thread each is plenty.
// Arithmetic: like oracle.cpp::clip_triangle — unfused a + w*scale for the plane distance, IEEE division for
// t, the source's fmaf(b, t, fmaf(-t, a, a)) for the new vertex.
#pragma once

#include "common.cuh"

namespace swrb {

struct ClipVert { float a[6]; };   // x y z w + barycentric weights of v1, v2 (Clipper::Vertices[6][..])

__device__ __forceinline__ float clip_dist(const ClipVert& v, uint32_t planeId, float scale) {   // GetIntersectDist, :76-81
    float a = v.a[planeId >> 1];
    if (planeId & 1u) a = -a;
    return __fadd_rn(a, __fmul_rn(v.a[3], scale));
}

__global__ void __launch_bounds__(128)
k_clip_triangles(const uint2* __restrict__ clipList, const swr_meshlet* __restrict__ meshlets,
                 const swr_material* __restrict__ materials, const DrawItem* __restrict__ draws, FrameParams fp,
                 TriRecord* __restrict__ tris, TriRecord* __restrict__ alphaTris, TriRecordW* __restrict__ alphaW,
                 float4* __restrict__ clipRemap, uint32_t triCapacity, DevCtl* __restrict__ ctl) {
    const uint32_t n = ctl->overflow ? 0u : min(ctl->clipCount, triCapacity);
    uint32_t nRasterized = 0;
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n; it += gridDim.x * blockDim.x) {
        const uint2 ent = clipList[it];
        const DrawItem& d = draws[ent.x];
        const uint32_t meshIdx = ent.y >> 7, prim = ent.y & 127u;
        const uint32_t meshletId = d.meshletOffset + meshIdx;
        const swr_meshlet* m = meshlets + meshletId;
        uint32_t cullMode = SWR_CULL_FRONT_CCW, fsId = 0;                       // Shading.cpp:302-306
        const uint32_t materialId = m->MaterialId;
        if (materialId != SWR_NO_MATERIAL && materials != nullptr) {
            swr_material mat = materials[materialId];
            cullMode = mat.IsDoubleSided ? SWR_CULL_NONE : SWR_CULL_FRONT_CCW;
            fsId = (mat.AlphaCutoff < 255 && fp.program == 0u) ? 1u : 0u;
        }
        if (fp.program == SWRB_PROGRAM_DEFERRED) fsId = 1u;                   // every piece needs 1/w and the barycentric remap (FS_EncodeGBuffer interpolates)

        ClipVert verts[16];          // 3 + at most 2 new vertices per plane (Clipper::Vertices, nextIdx <= 64 there)
        uint8_t indices[12], outIndices[12];
        uint32_t outcodes = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint32_t vid = m->Indices[k][prim] & 63u;
            const float x = m->Positions[0][vid], y = m->Positions[1][vid], z = m->Positions[2][vid];
            const float cx = __fmaf_rn(x, d.M[0], __fmaf_rn(y, d.M[4], __fmaf_rn(z, d.M[8], d.M[12])));
            const float cy = __fmaf_rn(x, d.M[1], __fmaf_rn(y, d.M[5], __fmaf_rn(z, d.M[9], d.M[13])));
            const float cz = __fmaf_rn(x, d.M[2], __fmaf_rn(y, d.M[6], __fmaf_rn(z, d.M[10], d.M[14])));
            const float cw = __fmaf_rn(x, d.M[3], __fmaf_rn(y, d.M[7], __fmaf_rn(z, d.M[11], d.M[15])));
            outcodes |= (cx < -cw) ? 1u : 0u;                                   // Rasterizer.cpp:375-383
            outcodes |= (cx > cw) ? 2u : 0u;
            outcodes |= (cy < -cw) ? 4u : 0u;
            outcodes |= (cy > cw) ? 8u : 0u;
            outcodes |= (cz < -cw) ? 16u : 0u;
            outcodes |= (cz > cw) ? 32u : 0u;
            verts[k].a[0] = cx; verts[k].a[1] = cy; verts[k].a[2] = cz; verts[k].a[3] = cw;
            verts[k].a[4] = k == 1 ? 1.0f : 0.0f; verts[k].a[5] = k == 2 ? 1.0f : 0.0f;
            indices[k] = (uint8_t)k;
        }

        uint32_t vertCount = 3, nextIdx = 3;
        for (uint32_t planeId = 0; planeId < 6; planeId++) {                    // :419
            if (!((outcodes >> planeId) & 1u)) continue;
            const float planeScale = planeId < 4 ? (planeId < 2 ? fp.bx : fp.by) : 1.0f;   // :422
            uint32_t outCount = 0;
            for (uint32_t vi = 0; vi < vertCount; vi++) {
                const uint32_t ia = indices[vi], ib = indices[vi + 1 == vertCount ? 0 : vi + 1];
                const float da = clip_dist(verts[ia], planeId, planeScale), db = clip_dist(verts[ib], planeId, planeScale);
                if (da >= 0.0f) outIndices[outCount++] = (uint8_t)ia;
                if ((da >= 0.0f) != (db >= 0.0f)) {                             // :432-439
                    const float t = __fdiv_rn(da, __fsub_rn(da, db));
#pragma unroll
                    for (int k = 0; k < 6; k++) verts[nextIdx].a[k] = __fmaf_rn(verts[ib].a[k], t, __fmaf_rn(-t, verts[ia].a[k], verts[ia].a[k]));
                    outIndices[outCount++] = (uint8_t)nextIdx++;
                }
            }
            if (vertCount < 3) break;                                           // :443
            vertCount = outCount;
            for (uint32_t k = 0; k < outCount; k++) indices[k] = outIndices[k];
        }
        if (vertCount < 3) continue;                                            // :448 (+ empty fan for 2 vertices)

        const uint32_t id = key_rank(meshletId, prim, 1u);      // pieces are drawn after the packet's accepted lanes (Rasterizer.cpp:209-249)
        for (uint32_t vi = 0; vi + 2 < vertCount; vi++) {                       // :451-466, then FlushPacket -> Setup
            const uint32_t ix[3] = { indices[0], indices[vi + 1], indices[vi + 2] };
            float nx[3], ny[3], nz[3], rw[3];
            uint32_t pos[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {                                       // perspective_div + snap (:258-260, :272-279)
                const ClipVert& v = verts[ix[k]];
                rw[k] = __fdiv_rn(1.0f, v.a[3]);
                nx[k] = __fmul_rn(v.a[0], rw[k]); ny[k] = __fmul_rn(v.a[1], rw[k]); nz[k] = __fmul_rn(v.a[2], rw[k]);
                const int32_t X = __float2int_rn(__fmul_rn(nx[k], fp.fixX)), Y = __float2int_rn(__fmul_rn(ny[k], fp.fixY));
                pos[k] = ((uint32_t)X & 0xFFFFu) | ((uint32_t)Y << 16);
            }
            float det = __fsub_rn(__fmul_rn(__fsub_rn(nx[2], nx[0]), __fsub_rn(ny[1], ny[0])),
                                  __fmul_rn(__fsub_rn(nx[0], nx[1]), __fsub_rn(ny[0], ny[2])));
            if (cullMode != SWR_CULL_FRONT_CCW) {
                const bool flip = (cullMode == SWR_CULL_FRONT_CW) ? true : (det < 0.0f);
                det = flip ? -det : det;
            }
            uint32_t bbMin, bbMax;
            ref_render_bbox(pos[0], pos[1], pos[2], fp.halfW, fp.halfH, bbMin, bbMax);
            if (!(det > 0.0f && lo16(bbMin) < lo16(bbMax) && hi16(bbMin) < hi16(bbMax))) continue;   // :269, :283
            nRasterized++;                                                      // :247
            if (hi16(bbMax) <= fp.bandY0 || hi16(bbMin) >= fp.bandY1) continue;   // no pixel inside the scissor rows

            const uint4 recA = make_uint4(pos[0], pos[1], pos[2], __float_as_uint(nz[0]));
            if (fsId && alphaTris != nullptr) {
                const uint32_t slot = atomicAdd(&ctl->alphaCount, 1u);
                if (slot >= triCapacity) { atomicExch(&ctl->overflow, 1u); continue; }
                uint4* dst = reinterpret_cast<uint4*>(alphaTris + slot);
                dst[0] = recA;
                dst[1] = make_uint4(__float_as_uint(nz[1]), __float_as_uint(nz[2]), id, 2u);   // aux 2: remap follows
                uint4* dw = reinterpret_cast<uint4*>(alphaW + slot);
                dw[0] = make_uint4(__float_as_uint(rw[0]), __float_as_uint(rw[1]), __float_as_uint(rw[2]), ent.x);
                uint32_t matWord = 0x00FFFFFFu;                                 // the ORIGINAL triangle's TexCoords: the remap below maps the piece's barycentrics back to them
                if (materialId != SWR_NO_MATERIAL && materials != nullptr) {
                    const swr_material mat = materials[materialId];
                    matWord = ((uint32_t)mat.TextureId & 0x00FFFFFFu) | ((uint32_t)mat.AlphaCutoff << 24);
                }
                dw[1] = make_uint4(m->TexCoords[m->Indices[0][prim] & 63u], m->TexCoords[m->Indices[1][prim] & 63u], m->TexCoords[m->Indices[2][prim] & 63u], matWord);
                const ClipVert &p0 = verts[ix[0]], &p1 = verts[ix[1]], &p2 = verts[ix[2]];
                clipRemap[2 * slot + 0] = make_float4(p0.a[4], __fsub_rn(p1.a[4], p0.a[4]), __fsub_rn(p2.a[4], p0.a[4]), p0.a[5]);   // ClippedU, ClippedV (:462-464)
                clipRemap[2 * slot + 1] = make_float4(__fsub_rn(p1.a[5], p0.a[5]), __fsub_rn(p2.a[5], p0.a[5]), 0.0f, 0.0f);
            } else {
                const uint32_t slot = atomicAdd(&ctl->triCount, 1u);
                if (slot >= triCapacity) { atomicExch(&ctl->overflow, 1u); continue; }
                uint4* dst = reinterpret_cast<uint4*>(tris + slot);
                dst[0] = recA;
                dst[1] = make_uint4(__float_as_uint(nz[1]), __float_as_uint(nz[2]), id, 0u);
            }
        }
    }
    if (nRasterized) atomicAdd(&ctl->perf[1], (unsigned long long)nRasterized);
}

}  // namespace swrb
