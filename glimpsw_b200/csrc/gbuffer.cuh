// gbuffer.cuh — ShadingContext::DeferredShader: Rasterizer::DrawTriangle<FS_EncodeGBuffer, IsClipped>
// (Rasterizer.h:250-328 + Shading.cpp:344-414, bound at :655), one warp per triangle.
//
// FS_EncodeGBuffer per 4x4 fragment: depth test against layer 1; material-less meshlets store the depth and nothing else;
// otherwise the base colour is sampled (SampleImplicitLod on layer 0, LOD from finite differences over the fragment's 16
// lanes, fragment-wide filter vote), fragments whose texel is below AlphaCutoff << 24 are discarded, and the survivor
// stores depth -> layer 1, base colour -> layer 0 and, for textures with a second layer, the normal-mapped world normal
// (octahedron, 2 x 10 bits) with the top 6 bits of metallic and roughness -> layer 2 (0 otherwise).
//
// The reference reaches the final G-buffer by read-modify-write in submission order. Here it takes two passes over the
// triangle records of a RUN of meshlets that are either all textured or all material-less (swrb.cu::draw_deferred splits
// a batch into such runs, in submission order):
//   pass 1  k_raster_gbuffer<false>: every fragment that passes its own alpha test competes with a 64-bit atomicMax on
//           depth | rank keys seeded with the depth layer — whether a fragment passes the alpha test does not depend on the
//           framebuffer, so the maximum is the fragment the sequential depth test would have kept (ties: first submitted);
//   pass 2  k_raster_gbuffer<true>: the same walk; the lane whose key IS the pixel's final key shades and stores.
// Inside a homogeneous run "the last fragment that passed the depth test" is also the one that wrote colour and normal
// last, which is what makes two passes enough; across runs the layers carry the state like they do between draws.
//
// Arithmetic: every operation that reaches a stored bit is the IEEE operation the canonical oracle performs (explicit _rn
// intrinsics, approx_rcp = 1/x, approx_rsqrt = 1/sqrt(x)): all three layers are bit-exact against
// oracle.cpp::draw_triangle_gbuffer, which is bit-exact against the reference's own code (tests/test_ref_pin.py).
#pragma once

#include "common.cuh"
#include "resolve.cuh"

namespace swrb {

// ---- exact (IEEE) versions of the small vector helpers; the resolve pass uses approximate ones under a tolerance gate
__device__ __forceinline__ float g_rsqrt(float x) { return __fdiv_rn(1.0f, __fsqrt_rn(x)); }                      // canonical approx_rsqrt
__device__ __forceinline__ F3 g_normalize(F3 a) {                                                                  // SIMD.h:439
    const float r = g_rsqrt(r_dot3(a, a));
    return { __fmul_rn(a.x, r), __fmul_rn(a.y, r), __fmul_rn(a.z, r) };
}
__device__ __forceinline__ F3 g_cross(F3 a, F3 b) {                                                                // SIMD.h:431-437
    return { __fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y)), __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z)), __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x)) };
}
__device__ __forceinline__ F3 g_mul_mat3(const float* m, F3 n) {                                                   // SIMD.h:465-471
    return { __fmaf_rn(n.x, m[0], __fmaf_rn(n.y, m[3], __fmul_rn(n.z, m[6]))),
             __fmaf_rn(n.x, m[1], __fmaf_rn(n.y, m[4], __fmul_rn(n.z, m[7]))),
             __fmaf_rn(n.x, m[2], __fmaf_rn(n.y, m[5], __fmul_rn(n.z, m[8]))) };
}
__device__ __forceinline__ F3 g_unmap_oct(float u, float v) {                                                      // Texture.h:289-296
    u = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f); v = __fsub_rn(__fmul_rn(v, 2.0f), 1.0f);
    F3 n = { u, v, __fsub_rn(__fsub_rn(1.0f, fabsf(u)), fabsf(v)) };
    const float t = fmaxf(-n.z, 0.0f);
    n.x = __fsub_rn(n.x, r_mulsign(t, n.x));
    n.y = __fsub_rn(n.y, r_mulsign(t, n.y));
    return g_normalize(n);
}
__device__ __forceinline__ void g_unpack_normal_tangent(uint32_t p, F3& n, F3& t) {                                // Shading.cpp:232-236
    const float s = 1.0f / 255;
    n = g_unmap_oct(__fmul_rn((float)(p & 255u), s), __fmul_rn((float)((p >> 8) & 255u), s));
    t = g_unmap_oct(__fmul_rn((float)((p >> 16) & 255u), s), __fmul_rn((float)(p >> 24), s));
}
__device__ __forceinline__ float g_bary(float b0, float b1, float b2, float v0, float v1, float v2) {             // Rasterizer.h:101-104
    return __fmaf_rn(v0, b0, __fmaf_rn(v1, b1, __fmul_rn(v2, b2)));
}

// Seeds the key buffer with the depth layer (low word 0xFFFFFFFF: "nobody of this run has won the pixel") and resets the
// per-draw work counters.
__global__ void __launch_bounds__(256)
k_gbuffer_begin(ulonglong2* __restrict__ keys, const uint4* __restrict__ depthLayer, uint32_t numVec, DevCtl* __restrict__ ctl) {
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (gid == 0) {
        ctl->triCount = 0; ctl->bigCount = 0; ctl->binTotal = 0; ctl->numActiveTiles = 0; ctl->alphaCount = 0; ctl->clipCount = 0;
        ctl->workCursor = 0; ctl->superTotal = 0; ctl->sparseTiles = 0; ctl->denseTiles = 0;
    }
    for (uint32_t i = gid; i < numVec; i += stride) {
        const uint4 d = __ldg(depthLayer + i);
        keys[2 * i + 0] = make_ulonglong2(((unsigned long long)d.x << 32) | 0xFFFFFFFFull, ((unsigned long long)d.y << 32) | 0xFFFFFFFFull);
        keys[2 * i + 1] = make_ulonglong2(((unsigned long long)d.z << 32) | 0xFFFFFFFFull, ((unsigned long long)d.w << 32) | 0xFFFFFFFFull);
    }
}

template <bool kWrite>
__global__ void __launch_bounds__(256)
k_raster_gbuffer(const TriRecord* __restrict__ tris, const TriRecordW* __restrict__ trisW, FrameParams fp,
                 const swr_meshlet* __restrict__ meshlets, const swr_material* __restrict__ materials,
                 const ResolveTexture* __restrict__ textures, const float4* __restrict__ clipRemap, const DrawItem* __restrict__ draws,
                 unsigned long long* __restrict__ keys, uint32_t* __restrict__ layer0, uint32_t* __restrict__ layer1, uint32_t* __restrict__ layer2,
                 DevCtl* __restrict__ ctl) {
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
        const bool clipped = t.aux == 2u;                                             // a piece from k_clip_triangles: DrawTriangle<FS, true>
        float4 ruv0 = make_float4(0.0f, 1.0f, 0.0f, 0.0f), ruv1 = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
        if (clipped) { ruv0 = __ldg(clipRemap + 2 * it); ruv1 = __ldg(clipRemap + 2 * it + 1); }

        const uint32_t meshletId = rank_meshlet(t.id);
        const swr_meshlet* mesh = meshlets + meshletId;
        const uint32_t prim = rank_prim(t.id);
        const uint32_t materialId = __ldg(&mesh->MaterialId);
        const bool textured = materialId != SWR_NO_MATERIAL;                          // Shading.cpp:352
        uint32_t vid[3];
        float uv[3][2];
#pragma unroll
        for (int k = 0; k < 3; k++) {                                                 // UnpackHalf2x16 of TexCoords[VertexId[k]]
            vid[k] = __ldg(&mesh->Indices[k][prim]) & 63u;
            uint32_t tc = __ldg(&mesh->TexCoords[vid[k]]);
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&tc));
            uv[k][0] = f.x; uv[k][1] = f.y;
        }
        swr_material mat;
        mat.TextureId = 0; mat.IsDoubleSided = 0; mat.AlphaCutoff = 0;
        if (textured) mat = materials[materialId];
        const ResolveTexture& tex = textures[textured ? mat.TextureId : 0];
        const uint32_t cutoff = (uint32_t)mat.AlphaCutoff << 24;
        const float scaleLerpU = textured ? (float)(tex.width << 8) : 1.0f, scaleLerpV = textured ? (float)(tex.height << 8) : 1.0f;

        uint32_t bbMin, bbMax;
        ref_render_bbox(t.pos0, t.pos1, t.pos2, fp.halfW, fp.halfH, bbMin, bbMax);    // the reference walks exactly this box
        const int32_t minX = lo16(bbMin), minY = hi16(bbMin), maxX = lo16(bbMax), maxY = hi16(bbMax);
        Edges e;
        const float rcpArea = edge_setup(t, fp.halfW, fp.halfH, e);
        const float W0 = w4.x, W0S = __fmul_rn(w4.x, rcpArea), W1S = __fmul_rn(w4.y, rcpArea), W2S = __fmul_rn(w4.z, rcpArea);   // Rasterizer.cpp:325-328

        const int32_t fragsX = (maxX - minX) >> 2, fragsY = (maxY - minY) >> 2;
        const int32_t numFrags = fragsX * fragsY;
        int32_t fx0 = 0, fy0 = 0;                                 // fragment f0 in row-major order over the box, advanced without a division (alpha.cuh)
        for (int32_t f0 = 0; f0 < numFrags; f0 += 2) {
            const int32_t frag = f0 + (int32_t)(lane >> 4);
            const bool valid = frag < numFrags;
            int32_t fx = fx0 + (int32_t)(lane >> 4), fy = fy0;
            if (fx >= fragsX) { fx -= fragsX; fy++; }
            if (!valid) { fx = 0; fy = 0; }
            fx0 += 2;
            if (fx0 >= fragsX) { fx0 -= fragsX; fy0++; if (fx0 >= fragsX) { fx0 -= fragsX; fy0++; } }
            const uint32_t px = (uint32_t)(minX + fx * 4) + (i & 3u), py = (uint32_t)(minY + fy * 4) + (i >> 2);
            const uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * px + (uint32_t)e.b12 * py;
            const uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * px + (uint32_t)e.b20 * py;
            const uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * px + (uint32_t)e.b01 * py;
            const bool covered = valid && (int32_t)(e0 | e1 | e2) >= 0;               // Rasterizer.h:289-290
            if ((__ballot_sync(0xFFFFFFFFu, covered)) == 0) continue;                 // :292 (warp-uniform)

            float u = __int2float_rn((int32_t)e1), v = __int2float_rn((int32_t)e2);
            const float depth = __fmaf_rn(u, e.z10, __fmaf_rn(v, e.z20, e.z0));       // :296
            const uint32_t off = fb_pixel_offset(px, py, fp.width);
            const unsigned long long key = make_key(depth, t.id);
            if (!textured) {                                                          // Shading.cpp:352-355: depth only
                if (covered && depth > 0.0f) {
                    if (!kWrite) { if (key > __ldcg(keys + off)) atomicMax(keys + off, key); }
                    else if (key == __ldcg(keys + off)) layer1[off] = __float_as_uint(depth);
                }
                continue;
            }
            // perspective correction (Rasterizer.h:302-310, :319), canonical rcp = 1/w
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
            const float tu = g_bary(b0, u, v, uv[0][0], uv[1][0], uv[2][0]);          // vars.Interpolate(uv0, uv1, uv2), Shading.cpp:362
            const float tv = g_bary(b0, u, v, uv[0][1], uv[1][1], uv[2][1]);

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

            if (!(covered && depth > 0.0f)) continue;
            const unsigned long long cur = __ldcg(keys + off);
            if (!kWrite) {
                if (key > cur) {                                                      // depth test first, like Shading.cpp:346-348
                    const uint32_t texel = r_sample_level(tex, tu, tv, 0, mip, useNearest);
                    if (texel >= cutoff) atomicMax(keys + off, key);                  // :365
                }
                continue;
            }
            if (key != cur) continue;                                                 // somebody else's pixel (or this fragment failed its alpha test)
            const uint32_t baseColor = r_sample_level(tex, tu, tv, 0, mip, useNearest);
            uint32_t packedCh2 = 0;                                                   // :368
            if (tex.numLayers >= 2u) {                                                // :370-403
                const uint32_t nmr = r_sample_level(tex, tu, tv, 1, mip, useNearest);
                const float nx = __fsub_rn(__fmul_rn((float)(nmr & 255u), 1.0f / 127.5f), 1.0f);
                const float ny = __fsub_rn(__fmul_rn((float)((nmr >> 8) & 255u), 1.0f / 127.5f), 1.0f);
                const float nzArg = __fsub_rn(1.0f, __fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)));
                const float nz = __fmul_rn(g_rsqrt(nzArg), nzArg);                    // approx_sqrt = approx_rsqrt(x) * x (SIMD.h:296)
                F3 n0, n1, n2, t0, t1, t2;
                g_unpack_normal_tangent(__ldg(&mesh->NormalTangents[vid[0]]), n0, t0);
                g_unpack_normal_tangent(__ldg(&mesh->NormalTangents[vid[1]]), n1, t1);
                g_unpack_normal_tangent(__ldg(&mesh->NormalTangents[vid[2]]), n2, t2);
                const float* o2w = draws[__float_as_uint(w4.w)].objectToWorld;        // ShadingContext::ObjectToWorldMat of the record's draw
                const F3 nl = { g_bary(b0, u, v, n0.x, n1.x, n2.x), g_bary(b0, u, v, n0.y, n1.y, n2.y), g_bary(b0, u, v, n0.z, n1.z, n2.z) };
                const F3 tl = { g_bary(b0, u, v, t0.x, t1.x, t2.x), g_bary(b0, u, v, t0.y, t1.y, t2.y), g_bary(b0, u, v, t0.z, t1.z, t2.z) };
                const F3 normalWS = g_normalize(g_mul_mat3(o2w, nl));
                const F3 tangentWS = g_normalize(g_mul_mat3(o2w, tl));
                const uint32_t handed = (uint32_t)((__ldg(&mesh->TangentHandedness) >> vid[0]) & 1ull) << 31;   // :386
                F3 bit = g_cross(normalWS, tangentWS);
                bit = { __uint_as_float(__float_as_uint(bit.x) ^ handed), __uint_as_float(__float_as_uint(bit.y) ^ handed), __uint_as_float(__float_as_uint(bit.z) ^ handed) };
                const F3 nr = g_normalize({ __fadd_rn(__fadd_rn(__fmul_rn(nx, tangentWS.x), __fmul_rn(ny, bit.x)), __fmul_rn(nz, normalWS.x)),
                                            __fadd_rn(__fadd_rn(__fmul_rn(nx, tangentWS.y), __fmul_rn(ny, bit.y)), __fmul_rn(nz, normalWS.y)),
                                            __fadd_rn(__fadd_rn(__fmul_rn(nx, tangentWS.z), __fmul_rn(ny, bit.z)), __fmul_rn(nz, normalWS.z)) });
                // texutil::MapOctahedron (Texture.h:282-288)
                const float ow = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(fabsf(nr.x), fabsf(nr.y)), fabsf(nr.z)));
                const float ot = fmaxf(__fmul_rn(-nr.z, ow), 0.0f);
                const float ou = __fadd_rn(__fmul_rn(__fmaf_rn(nr.x, ow, r_mulsign(ot, nr.x)), 0.5f), 0.5f);
                const float ov = __fadd_rn(__fmul_rn(__fmaf_rn(nr.y, ow, r_mulsign(ot, nr.y)), 0.5f), 0.5f);
                packedCh2 = __float2uint_rz(__fadd_rn(__fmul_rn(ou, 1023.0f), 0.5f));                          // :399
                packedCh2 |= __float2uint_rz(__fadd_rn(__fmul_rn(ov, 1023.0f), 0.5f)) << 10;                   // :400
                packedCh2 |= ((nmr >> 18) & 0x3Fu) << 20;                                                      // :401
                packedCh2 |= ((nmr >> 26) & 0x3Fu) << 26;                                                      // :402
            }
            layer1[off] = __float_as_uint(depth);                                     // :404-406
            layer0[off] = baseColor;
            layer2[off] = packedCh2;
        }
    }
}

}  // namespace swrb
