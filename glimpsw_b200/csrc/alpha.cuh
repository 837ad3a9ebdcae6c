// alpha.cuh — the alpha-tested fragment program: Rasterizer::DrawTriangle<FS_EncodeSurfaceId<true>>
// (Rasterizer.h:250-328 + Shading.cpp:309-331), one warp per triangle.
//
// The reference evaluates this program per 4x4 fragment with 16 SIMD lanes: perspective-correct
// barycentrics (Rasterizer.h:302-319), fp16 UVs interpolated with BaryLerp, SampleImplicitLod (texture LOD
// from 2x2 finite differences over ALL 16 lanes of the fragment, Texture.h:260-275, :403-410; nearest-vs-
// bilinear chosen by a fragment-wide vote, :432) and `texel >= AlphaCutoff << 24` (Shading.cpp:326).
// Here a warp walks the triangle's tile-aligned bounding box two fragments at a time (lanes 0-15 and 16-31),
// so the finite differences are warp shuffles (xor 1 / xor 4) and the vote a half-warp ballot.
// Whether a fragment passes the alpha test does not depend on the depth buffer, so the usual 64-bit
// atomicMax on depth|id keys still reproduces the reference's sequential result.
// Arithmetic: IEEE, op for op like oracle.cpp::draw_triangle_alpha (canonical approx_rcp = 1/w), so the
// vis-buffer stays bit-exact. Alpha-tested triangles of any size come here (never to the inline raster or the
// binner); the mesh kernel writes their records to a separate list together with 1/w of the three vertices, the three
// TexCoords words and the material's TextureId | AlphaCutoff (TriRecordW), so a warp here goes record -> texture header ->
// texels instead of record -> meshlet indices -> TexCoords -> material -> texture header -> texels.
#pragma once

#include "common.cuh"
#include "resolve.cuh"

namespace swrb {

__global__ void __launch_bounds__(256)
k_raster_alpha(const TriRecord* __restrict__ tris, const TriRecordW* __restrict__ trisW, FrameParams fp,
               const swr_meshlet* __restrict__ meshlets, const swr_material* __restrict__ materials,
               const ResolveTexture* __restrict__ textures, const float4* __restrict__ clipRemap,
               unsigned long long* __restrict__ keys, DevCtl* __restrict__ ctl) {
    (void)meshlets; (void)materials;
    const uint32_t n = ctl->overflow ? 0u : ctl->alphaCount;
    const uint32_t lane = lane_id(), i = lane & 15u, half = 0xFFFFu << (lane & 16u);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t it = warp; it < n; it += warps) {
        const uint4* src = reinterpret_cast<const uint4*>(tris + it);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        TriRecord t;
        t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
        t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(trisW + it));
        const uint4 tcm = __ldg(reinterpret_cast<const uint4*>(trisW + it) + 1);      // TexCoords[VertexId[0..2]], TextureId | AlphaCutoff << 24 (written with the record)
        const bool clipped = t.aux == 2u;                                             // a piece from k_clip_triangles: DrawTriangle<FS, true>
        float4 ruv0 = make_float4(0.0f, 1.0f, 0.0f, 0.0f), ruv1 = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
        if (clipped) { ruv0 = __ldg(clipRemap + 2 * it); ruv1 = __ldg(clipRemap + 2 * it + 1); }

        float uv[3][2];
        {                                                                             // UnpackHalf2x16 of the three TexCoords words
            const uint32_t tc[3] = { tcm.x, tcm.y, tcm.z };
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float2 f = __half22float2(*reinterpret_cast<const __half2*>(&tc[k]));
                uv[k][0] = f.x; uv[k][1] = f.y;
            }
        }
        const ResolveTexture& tex = textures[tcm.w & 0x00FFFFFFu];
        const uint32_t cutoff = tcm.w & 0xFF000000u;                                  // AlphaCutoff << 24
        const float scaleLerpU = (float)(tex.width << 8), scaleLerpV = (float)(tex.height << 8);

        uint32_t bbMin, bbMax;
        ref_render_bbox(t.pos0, t.pos1, t.pos2, fp.halfW, fp.halfH, bbMin, bbMax);    // the reference walks exactly this box
        const int32_t minX = lo16(bbMin), minY = hi16(bbMin), maxX = lo16(bbMax), maxY = hi16(bbMax);
        Edges e;
        const float rcpArea = edge_setup(t, fp.halfW, fp.halfH, e);
        const float W0 = w4.x, W0S = __fmul_rn(w4.x, rcpArea), W1S = __fmul_rn(w4.y, rcpArea), W2S = __fmul_rn(w4.z, rcpArea);   // Rasterizer.cpp:325-328

        const int32_t fragsX = (maxX - minX) >> 2, fragsY = (maxY - minY) >> 2;
        const int32_t numFrags = fragsX * fragsY;
        // fragment f0 = (fx0, fy0) in row-major order over the box, advanced by two per step without a division
        int32_t fx0 = 0, fy0 = 0;
        for (int32_t f0 = 0; f0 < numFrags; f0 += 2) {
            const int32_t frag = f0 + (int32_t)(lane >> 4);
            const bool valid = frag < numFrags;
            int32_t fx = fx0 + (int32_t)(lane >> 4), fy = fy0;
            if (fx >= fragsX) { fx -= fragsX; fy++; }             // the upper half warp's fragment may start the next row (fragsX >= 1)
            if (!valid) { fx = 0; fy = 0; }
            fx0 += 2;
            if (fx0 >= fragsX) { fx0 -= fragsX; fy0++; if (fx0 >= fragsX) { fx0 -= fragsX; fy0++; } }   // fragsX == 1: two rows per step
            const uint32_t px = (uint32_t)(minX + fx * 4) + (i & 3u), py = (uint32_t)(minY + fy * 4) + (i >> 2);
            const uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * px + (uint32_t)e.b12 * py;
            const uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * px + (uint32_t)e.b20 * py;
            const uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * px + (uint32_t)e.b01 * py;
            const bool covered = valid && (int32_t)(e0 | e1 | e2) >= 0;               // Rasterizer.h:289-290
            if ((__ballot_sync(0xFFFFFFFFu, covered)) == 0) continue;                 // :292 (warp-uniform)

            float u = __int2float_rn((int32_t)e1), v = __int2float_rn((int32_t)e2);
            const float depth = __fmaf_rn(u, e.z10, __fmaf_rn(v, e.z20, e.z0));       // :296
            // depth test first, like Shading.cpp:311-313 — and when no pixel of the two fragments can win, nobody needs the texture
            const uint32_t off = fb_pixel_offset(px, py, fp.width);
            const unsigned long long key = make_key(depth, t.id);
            const bool wins = covered && depth > 0.0f && key > __ldcg(keys + off);
            if (__ballot_sync(0xFFFFFFFFu, wins) == 0) continue;
            // perspective correction (:302-310, :319), canonical rcp = 1/w
            const float pw0 = __fmaf_rn(__fadd_rn(u, v), -W0S, W0);
            const float w = __fmaf_rn(u, W1S, __fmaf_rn(v, W2S, pw0));
            float rcpW = __fdiv_rn(1.0f, w);
            rcpW = __fmul_rn(rcpW, __fmaf_rn(-w, rcpW, 2.0f));
            u = __fmul_rn(u, __fmul_rn(W1S, rcpW));
            v = __fmul_rn(v, __fmul_rn(W2S, rcpW));
            if (clipped) {                                                            // ClippedU / ClippedV remap (Rasterizer.h:312-318)
                const float cu = __fmaf_rn(u, ruv0.y, __fmaf_rn(v, ruv0.z, ruv0.x));
                const float cv = __fmaf_rn(u, ruv1.x, __fmaf_rn(v, ruv1.y, ruv0.w));
                u = cu; v = cv;
            }
            const float b0 = __fsub_rn(__fsub_rn(1.0f, u), v);
            const float tu = __fmaf_rn(uv[0][0], b0, __fmaf_rn(uv[1][0], u, __fmul_rn(uv[2][0], v)));   // BaryLerp, Rasterizer.h:101-104
            const float tv = __fmaf_rn(uv[0][1], b0, __fmaf_rn(uv[1][1], u, __fmul_rn(uv[2][1], v)));

            // SampleImplicitLod: 2x2 finite differences inside the fragment (Texture.h:260-269, :403-410)
            const float su = __fmul_rn(tu, scaleLerpU), sv = __fmul_rn(tv, scaleLerpV);
            const float suX = __shfl_xor_sync(0xFFFFFFFFu, su, 1), svX = __shfl_xor_sync(0xFFFFFFFFu, sv, 1);
            const float suY = __shfl_xor_sync(0xFFFFFFFFu, su, 4), svY = __shfl_xor_sync(0xFFFFFFFFu, sv, 4);
            const bool oddX = (i & 1u) != 0, oddY = (i & 4u) != 0;
            const float gxu = oddX ? __fsub_rn(su, suX) : __fsub_rn(suX, su), gxv = oddX ? __fsub_rn(sv, svX) : __fsub_rn(svX, sv);
            const float gyu = oddY ? __fsub_rn(su, suY) : __fsub_rn(suY, su), gyv = oddY ? __fsub_rn(sv, svY) : __fsub_rn(svY, sv);
            const float dx = __fmaf_rn(gxu, gxu, __fmul_rn(gxv, gxv)), dy = __fmaf_rn(gyu, gyu, __fmul_rn(gyv, gyv));
            const int32_t mip = ((((int32_t)__float_as_uint(fmaxf(dx, dy)) - (127 << 23)) >> 23) >> 1) - 8;   // CalcMipLevel - LerpFracBits
            const bool useNearest = (__ballot_sync(0xFFFFFFFFu, valid && mip > 0) & half) != 0;               // Texture.h:432

            if (wins) {
                const uint32_t texel = r_sample_level(tex, tu, tv, 0, mip, useNearest);
                if (texel >= cutoff) atomicMax(keys + off, key);                      // :326-330
            }
        }
    }
}

}  // namespace swrb
