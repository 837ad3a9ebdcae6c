// alpha.cuh — the alpha-tested fragment program: Rasterizer::DrawTriangle<FS_EncodeSurfaceId<true>>
// (Rasterizer.h:250-328 + Shading.cpp:309-331), 32 triangles per warp.
//
// The reference evaluates this program per 4x4 fragment with 16 SIMD lanes: perspective-correct
// barycentrics (Rasterizer.h:302-319), fp16 UVs interpolated with BaryLerp, SampleImplicitLod (texture LOD
// from 2x2 finite differences over ALL 16 lanes of the fragment, Texture.h:260-275, :403-410; nearest-vs-
// bilinear chosen by a fragment-wide vote, :432) and `texel >= AlphaCutoff << 24` (Shading.cpp:326).
// Whether a fragment passes the alpha test does not depend on the depth buffer, so the usual 64-bit
// atomicMax on depth|id keys still reproduces the reference's sequential result.
// Arithmetic: IEEE, op for op like oracle.cpp::draw_triangle_alpha (canonical approx_rcp = 1/w), so the
// vis-buffer stays bit-exact. Alpha-tested triangles of any size come here (never to the inline raster or the
// binner); the mesh kernel writes their records to a separate list together with 1/w of the three vertices, the three
// TexCoords words and the material's TextureId | AlphaCutoff (TriRecordW), so nothing here has to chase
// record -> meshlet indices -> TexCoords -> material.
//
// Shape of the kernel. A warp takes 32 records — an interleaved sample of the list (record = warp + lane * warps):
// consecutive records are the triangles of one meshlet, neighbours on screen and all large or all small, and warps that
// drew 32 consecutive ones from a cursor finished between 0.2x and 4x the average time (measured on the C3 frame).
//  1. Every lane sets up ONE triangle (bounding box, edge functions, W terms, UVs) into shared memory — once, not 32
//     times over as with a warp per triangle (which cost 155 of ~650 warp instructions per triangle).
//  2. A warp scan of the boxes' fragment counts turns the 32 triangles into one list of (triangle, fragment) items.
//  3. 32 items at a time, a lane per fragment evaluates the edge functions at its 16 pixels; only fragments with a
//     covered pixel survive (Rasterizer.h:292; ballot + compaction into shared memory).
//  4. The survivors are shaded two per step, a half warp each, a lane per pixel: depth test first (a step none of whose
//     pixels can win is dropped before the UV / LOD arithmetic), the LOD finite differences are warp shuffles
//     (xor 1 / xor 4), the fragment-wide filter vote a half-warp ballot.
// C3 frame: 58 M warp instructions instead of 99 M, raster stage 214 -> 142 us (profiles/r02_summary.md). A lane per
// fragment for step 4 as well (16 pixels in sequence per lane) was built and measured 7x slower: the step's dependent
// loads (key, texel) then line up 16 deep per lane instead of running side by side across the lanes.
#pragma once

#include "common.cuh"
#include "resolve.cuh"

namespace swrb {

constexpr int kAlphaWarps = 8;

struct AlphaTri {                        // one triangle's setup
    int32_t e0, e1, e2, a12, a20, a01, b12, b20, b01;
    int32_t minX, minY, fragsX;
    float z0, z10, z20, W0, W0S, W1S, W2S;
    float uv00, uv01, uv10, uv11, uv20, uv21;
    float r0x, r0y, r0z, r0w, r1x, r1y;  // ClippedU / ClippedV remap
    uint32_t id, mat, clipped;
};

__global__ void __launch_bounds__(kAlphaWarps * 32, 4)
k_raster_alpha(const TriRecord* __restrict__ tris, const TriRecordW* __restrict__ trisW, FrameParams fp,
                    const ResolveTexture* __restrict__ textures, const float4* __restrict__ clipRemap,
                    unsigned long long* __restrict__ keys, DevCtl* __restrict__ ctl) {
    __shared__ AlphaTri sTri[kAlphaWarps][32];
    __shared__ uint32_t sPrefix[kAlphaWarps][33];
    __shared__ uint32_t sLive[kAlphaWarps][32];
    const uint32_t n = ctl->overflow ? 0u : ctl->alphaCount;
    const uint32_t lane = lane_id(), wib = threadIdx.x >> 5;
    const uint32_t i = lane & 15u, hw = lane >> 4, half = 0xFFFFu << (lane & 16u);
    AlphaTri* const myTris = sTri[wib];
    uint32_t* const prefix = sPrefix[wib];
    uint32_t* const liveList = sLive[wib];
    // The warp's records are base, base + W, base + 2W, ... (W = warps in the grid), 32 per trip: consecutive records are the
    // triangles of one meshlet — neighbours on screen, all large or all small — and a warp that drew 32 consecutive ones from a
    // cursor took four times as long as the average (measured); the interleaved deal hands every warp a sample of the whole list.
    const uint32_t warpsTotal = gridDim.x * kAlphaWarps;
    for (uint32_t base = blockIdx.x * kAlphaWarps + wib; base < n; base += 32u * warpsTotal) {
        const uint64_t it64 = (uint64_t)base + (uint64_t)lane * warpsTotal;
        const uint32_t it = it64 < n ? (uint32_t)it64 : n;
        uint32_t numFrags = 0;
        if (it < n) {
            const uint4* src = reinterpret_cast<const uint4*>(tris + it);
            const uint4 a = __ldg(src), b = __ldg(src + 1);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(trisW + it));
            const uint4 tcm = __ldg(reinterpret_cast<const uint4*>(trisW + it) + 1);
            TriRecord t;
            t.pos0 = a.x; t.pos1 = a.y; t.pos2 = a.z; t.z0 = __uint_as_float(a.w);
            t.z1 = __uint_as_float(b.x); t.z2 = __uint_as_float(b.y); t.id = b.z; t.aux = b.w;
            AlphaTri s;
            s.clipped = t.aux == 2u ? 1u : 0u;                                        // a piece from k_clip_triangles: DrawTriangle<FS, true>
            float4 ruv0 = make_float4(0.0f, 1.0f, 0.0f, 0.0f), ruv1 = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
            if (s.clipped) { ruv0 = __ldg(clipRemap + 2 * it); ruv1 = __ldg(clipRemap + 2 * it + 1); }
            s.r0x = ruv0.x; s.r0y = ruv0.y; s.r0z = ruv0.z; s.r0w = ruv0.w; s.r1x = ruv1.x; s.r1y = ruv1.y;
            const uint32_t tc[3] = { tcm.x, tcm.y, tcm.z };
            float uv[3][2];
#pragma unroll
            for (int k = 0; k < 3; k++) {                                             // UnpackHalf2x16 of the three TexCoords words
                float2 f = __half22float2(*reinterpret_cast<const __half2*>(&tc[k]));
                uv[k][0] = f.x; uv[k][1] = f.y;
            }
            s.uv00 = uv[0][0]; s.uv01 = uv[0][1]; s.uv10 = uv[1][0]; s.uv11 = uv[1][1]; s.uv20 = uv[2][0]; s.uv21 = uv[2][1];
            s.id = t.id; s.mat = tcm.w;
            uint32_t bbMin, bbMax;
            ref_render_bbox(t.pos0, t.pos1, t.pos2, fp.halfW, fp.halfH, bbMin, bbMax);    // the reference walks exactly this box
            s.minX = lo16(bbMin); s.minY = hi16(bbMin);
            const int32_t fragsY = (hi16(bbMax) - s.minY) >> 2;
            s.fragsX = (lo16(bbMax) - s.minX) >> 2;
            numFrags = (s.fragsX > 0 && fragsY > 0) ? (uint32_t)(s.fragsX * fragsY) : 0u;
            Edges e;
            const float rcpArea = edge_setup(t, fp.halfW, fp.halfH, e);
            s.e0 = e.e0; s.e1 = e.e1; s.e2 = e.e2; s.a12 = e.a12; s.a20 = e.a20; s.a01 = e.a01; s.b12 = e.b12; s.b20 = e.b20; s.b01 = e.b01;
            s.z0 = e.z0; s.z10 = e.z10; s.z20 = e.z20;
            s.W0 = w4.x; s.W0S = __fmul_rn(w4.x, rcpArea); s.W1S = __fmul_rn(w4.y, rcpArea); s.W2S = __fmul_rn(w4.z, rcpArea);   // Rasterizer.cpp:325-328
            myTris[lane] = s;
        }
        // exclusive scan of the fragment counts -> prefix[0..32]
        uint32_t incl = numFrags;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        prefix[lane] = incl - numFrags;
        if (lane == 31) prefix[32] = incl;
        __syncwarp();
        const uint32_t total = prefix[32];

        for (uint32_t g0 = 0; g0 < total; g0 += 32) {
            // ---- a lane per fragment: does the triangle cover any of its 16 pixels?
            const uint32_t g = g0 + lane;
            bool live = false;
            uint32_t packed = 0;
            if (g < total) {
                uint32_t lo = 0, hi = 32;                       // last triangle with prefix <= g (empty boxes share their successor's prefix)
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= g) lo = mid; else hi = mid; }
                const AlphaTri& t = myTris[lo];
                const uint32_t f = g - prefix[lo], fragsX = (uint32_t)t.fragsX;
                const uint32_t fy = f / fragsX, fx = f - fy * fragsX;
                const uint32_t x0 = (uint32_t)(t.minX + (int32_t)fx * 4), y0 = (uint32_t)(t.minY + (int32_t)fy * 4);
                const uint32_t a12 = (uint32_t)t.a12, a20 = (uint32_t)t.a20, a01 = (uint32_t)t.a01;
                const uint32_t b12 = (uint32_t)t.b12, b20 = (uint32_t)t.b20, b01 = (uint32_t)t.b01;
                uint32_t r0 = (uint32_t)t.e0 + a12 * x0 + b12 * y0, r1 = (uint32_t)t.e1 + a20 * x0 + b20 * y0, r2 = (uint32_t)t.e2 + a01 * x0 + b01 * y0;
#pragma unroll
                for (int yy = 0; yy < 4; yy++) {                // wrapping uint32 sums: the same values as e + a * px + b * py
                    uint32_t c0 = r0, c1 = r1, c2 = r2;
#pragma unroll
                    for (int xx = 0; xx < 4; xx++) {
                        live = live || (int32_t)(c0 | c1 | c2) >= 0;                // Rasterizer.h:289-290
                        c0 += a12; c1 += a20; c2 += a01;
                    }
                    r0 += b12; r1 += b20; r2 += b01;
                }
                packed = lo | ((x0 >> 2) << 5) | ((y0 >> 2) << 18);                  // x0, y0 < 8192, multiples of 4
            }
            const uint32_t liveMask = __ballot_sync(0xFFFFFFFFu, live);              // :292 — the rest of the box is skipped
            if (live) liveList[__popc(liveMask & ((1u << lane) - 1u))] = packed;
            __syncwarp();
            const uint32_t numLive = (uint32_t)__popc(liveMask);

            // ---- two surviving fragments per step, a half warp each, a lane per pixel
            for (uint32_t k0 = 0; k0 < numLive; k0 += 2) {
                const bool valid = k0 + hw < numLive;
                const uint32_t pk = liveList[valid ? k0 + hw : k0];
                const AlphaTri& t = myTris[pk & 31u];
                const uint32_t px = (((pk >> 5) & 0x1FFFu) << 2) + (i & 3u), py = ((pk >> 18) << 2) + (i >> 2);
                const uint32_t e0 = (uint32_t)t.e0 + (uint32_t)t.a12 * px + (uint32_t)t.b12 * py;
                const uint32_t e1 = (uint32_t)t.e1 + (uint32_t)t.a20 * px + (uint32_t)t.b20 * py;
                const uint32_t e2 = (uint32_t)t.e2 + (uint32_t)t.a01 * px + (uint32_t)t.b01 * py;
                const bool covered = valid && (int32_t)(e0 | e1 | e2) >= 0;           // Rasterizer.h:289-290
                float u = __int2float_rn((int32_t)e1), v = __int2float_rn((int32_t)e2);
                const float depth = __fmaf_rn(u, t.z10, __fmaf_rn(v, t.z20, t.z0));   // :296
                // depth test first, like Shading.cpp:311-313 — and when no pixel of the two fragments can win, nobody needs the texture
                const uint32_t id = t.id;
                const uint32_t off = fb_pixel_offset(px, py, fp.width);
                const unsigned long long key = make_key(depth, id);
                const bool wins = covered && depth > 0.0f && key > __ldcg(keys + off);
                if (__ballot_sync(0xFFFFFFFFu, wins) == 0) continue;
                // perspective correction (:302-310, :319), canonical rcp = 1/w
                const float W1S = t.W1S, W2S = t.W2S;
                const float pw0 = __fmaf_rn(__fadd_rn(u, v), -t.W0S, t.W0);
                const float w = __fmaf_rn(u, W1S, __fmaf_rn(v, W2S, pw0));
                float rcpW = __fdiv_rn(1.0f, w);
                rcpW = __fmul_rn(rcpW, __fmaf_rn(-w, rcpW, 2.0f));
                u = __fmul_rn(u, __fmul_rn(W1S, rcpW));
                v = __fmul_rn(v, __fmul_rn(W2S, rcpW));
                if (t.clipped) {                                                      // ClippedU / ClippedV remap (Rasterizer.h:312-318)
                    const float cu = __fmaf_rn(u, t.r0y, __fmaf_rn(v, t.r0z, t.r0x));
                    const float cv = __fmaf_rn(u, t.r1x, __fmaf_rn(v, t.r1y, t.r0w));
                    u = cu; v = cv;
                }
                const float b0 = __fsub_rn(__fsub_rn(1.0f, u), v);
                const float tu = __fmaf_rn(t.uv00, b0, __fmaf_rn(t.uv10, u, __fmul_rn(t.uv20, v)));   // BaryLerp, Rasterizer.h:101-104
                const float tv = __fmaf_rn(t.uv01, b0, __fmaf_rn(t.uv11, u, __fmul_rn(t.uv21, v)));
                const uint32_t mat = t.mat;
                const ResolveTexture& tex = textures[mat & 0x00FFFFFFu];
                // SampleImplicitLod: 2x2 finite differences inside the fragment (Texture.h:260-269, :403-410)
                const float su = __fmul_rn(tu, (float)(tex.width << 8)), sv = __fmul_rn(tv, (float)(tex.height << 8));
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
                    if (texel >= (mat & 0xFF000000u)) atomicMax(keys + off, key);     // :326-330 (AlphaCutoff << 24)
                }
            }
            __syncwarp();                                       // the list of survivors is rewritten by the next 32 items
        }
        __syncwarp();                                           // everyone is done with this batch's shared records
    }
}

}  // namespace swrb
