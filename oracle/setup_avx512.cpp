// setup_avx512.cpp — the mesh-shading / classification / early-setup half of the TIMED CPU baseline on 16-lane vectors, the
// way the reference processes it: ShadeMeshlet transforms 16 vertices per instruction (Shading.cpp:281-307, SIMD.h:457-464)
// and Rasterizer::DrawMeshlets classifies and sets up 16-triangle packets (ComputeClipCodes Rasterizer.cpp:353-397,
// TrianglePacket::Setup :257-289, GetRenderBoundingBox :331-351). TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
//
// Same IEEE operations in the same order as the scalar loop it replaces in baseline_mt.cpp, so the results are bit-identical
// (tests/test_baseline_cpu.py compares the whole frame with the scalar spec). One deliberate difference to the reference
// kept from the scalar baseline: perspective divide, snapping and outcodes are computed once per VERTEX and gathered
// per triangle corner (64-entry tables, the reference's GatherPos idea, Rasterizer.cpp:143-151) instead of once per corner.
// Compiled on its own with -march=x86-64-v4 (oracle/Makefile); callers check the CPU before calling in.
#include <immintrin.h>

#include <cstdint>

#include "../include/swr_types.h"

namespace {

inline __m512i gather_u32(const uint32_t* table, __m512i idx) { return _mm512_i32gather_epi32(idx, table, 4); }
inline __m512 gather_f32(const float* table, __m512i idx) { return _mm512_i32gather_ps(idx, table, 4); }

}  // namespace

// Per-vertex tables nx / ny / nz / pos / fl (64 entries each, owned by the calling worker and deliberately NOT cleared between
// meshlets, like the scalar path) are refreshed for the meshlet's vertex slots; then every 16-triangle packet is classified.
// Outputs: keep[4] = one bit per primitive that survives classification, culling and the bounding-box test;
// bbMin / bbMax[128] = GetRenderBoundingBox of those primitives; returns the number of non-trivial (clipped) primitives.
extern "C" uint32_t orc_meshlet_setup_avx512(const swr_meshlet* meshPtr, const float* M, float bx, float by, float fixX, float fixY,
                                             int halfW, int halfH, int cullMode, float* nx, float* ny, float* nz, uint32_t* pos,
                                             uint32_t* fl, uint32_t* keep, uint32_t* bbMinOut, uint32_t* bbMaxOut) {
    const swr_meshlet& mesh = *meshPtr;
    const uint32_t primCount = mesh.NumTriangles;
    uint32_t nv = ((uint32_t)mesh.NumVertices + 15u) & ~15u;
    if (nv > 64) nv = 64;

    // ---- ShadeMeshlet + the per-vertex part of ComputeClipCodes / TrianglePacket::Setup, 16 vertices at a time
    for (uint32_t v = 0; v < nv; v += 16) {
        const __m512 x = _mm512_loadu_ps(&mesh.Positions[0][v]), y = _mm512_loadu_ps(&mesh.Positions[1][v]), z = _mm512_loadu_ps(&mesh.Positions[2][v]);
#define ROW(r) _mm512_fmadd_ps(x, _mm512_set1_ps(M[0 + r]), _mm512_fmadd_ps(y, _mm512_set1_ps(M[4 + r]), _mm512_fmadd_ps(z, _mm512_set1_ps(M[8 + r]), _mm512_set1_ps(1.0f * M[12 + r]))))
        const __m512 cx = ROW(0), cy = ROW(1), cz = ROW(2), cw = ROW(3);
#undef ROW
        const __m512 ncw = _mm512_xor_ps(cw, _mm512_set1_ps(-0.0f));
        __m512i f = _mm512_setzero_si512();
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cx, ncw, _CMP_LT_OQ), f, _mm512_set1_epi32(1));
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cx, cw, _CMP_GT_OQ), f, _mm512_set1_epi32(2));
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cy, ncw, _CMP_LT_OQ), f, _mm512_set1_epi32(4));
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cy, cw, _CMP_GT_OQ), f, _mm512_set1_epi32(8));
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cz, ncw, _CMP_LT_OQ), f, _mm512_set1_epi32(16));
        f = _mm512_mask_or_epi32(f, _mm512_cmp_ps_mask(cz, cw, _CMP_GT_OQ), f, _mm512_set1_epi32(32));
        const __m512 absMask = _mm512_castsi512_ps(_mm512_set1_epi32(0x7FFFFFFF));
        const __mmask16 inGuard = _mm512_cmp_ps_mask(_mm512_and_ps(cx, absMask), _mm512_mul_ps(cw, _mm512_set1_ps(bx)), _CMP_LT_OQ) &
                                  _mm512_cmp_ps_mask(_mm512_and_ps(cy, absMask), _mm512_mul_ps(cw, _mm512_set1_ps(by)), _CMP_LT_OQ);
        f = _mm512_mask_or_epi32(f, inGuard, f, _mm512_set1_epi32(64));
        const __m512 rw = _mm512_div_ps(_mm512_set1_ps(1.0f), cw);
        const __m512 vnx = _mm512_mul_ps(cx, rw), vny = _mm512_mul_ps(cy, rw), vnz = _mm512_mul_ps(cz, rw);
        const __m512i X = _mm512_cvtps_epi32(_mm512_mul_ps(vnx, _mm512_set1_ps(fixX)));      // RNE; 0x80000000 out of range, like the scalar path
        const __m512i Y = _mm512_cvtps_epi32(_mm512_mul_ps(vny, _mm512_set1_ps(fixY)));
        _mm512_storeu_ps(nx + v, vnx); _mm512_storeu_ps(ny + v, vny); _mm512_storeu_ps(nz + v, vnz);
        _mm512_storeu_si512(pos + v, _mm512_or_si512(_mm512_and_si512(X, _mm512_set1_epi32(0xFFFF)), _mm512_slli_epi32(Y, 16)));
        _mm512_storeu_si512(fl + v, f);
    }

    // ---- 16-triangle packets
    uint32_t nClip = 0;
    keep[0] = keep[1] = keep[2] = keep[3] = 0;
    const __m512i lane = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i vp = _mm512_set1_epi32((int)((uint32_t)halfW | ((uint32_t)halfH << 16)));
    const __m512i vp2 = _mm512_set1_epi32((int)(((uint32_t)halfW | ((uint32_t)halfH << 16)) * 2u));
    for (uint32_t prim0 = 0; prim0 < primCount; prim0 += 16) {
        const __mmask16 valid = _mm512_cmplt_epu32_mask(_mm512_add_epi32(_mm512_set1_epi32((int)prim0), lane), _mm512_set1_epi32((int)primCount));
        const __m512i i0 = _mm512_and_si512(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)&mesh.Indices[0][prim0])), _mm512_set1_epi32(63));
        const __m512i i1 = _mm512_and_si512(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)&mesh.Indices[1][prim0])), _mm512_set1_epi32(63));
        const __m512i i2 = _mm512_and_si512(_mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i*)&mesh.Indices[2][prim0])), _mm512_set1_epi32(63));
        const __m512i f0 = gather_u32(fl, i0), f1 = gather_u32(fl, i1), f2 = gather_u32(fl, i2);
        const __m512i partial = _mm512_or_si512(_mm512_or_si512(f0, f1), f2), combined = _mm512_and_si512(_mm512_and_si512(f0, f1), f2);
        const __mmask16 visible = _mm512_testn_epi32_mask(combined, _mm512_set1_epi32(63)) & valid;                       // Rasterizer.cpp:389
        const __mmask16 trivial = _mm512_test_epi32_mask(combined, _mm512_set1_epi32(64)) & _mm512_testn_epi32_mask(partial, _mm512_set1_epi32(48));   // :386-388
        nClip += (uint32_t)_mm_popcnt_u32((uint32_t)(visible & (__mmask16)~trivial));                                      // :393
        __mmask16 accept = visible & trivial;
        if (accept == 0) continue;

        // TrianglePacket::Setup (Rasterizer.cpp:257-289)
        const __m512 x0 = gather_f32(nx, i0), y0 = gather_f32(ny, i0), x1 = gather_f32(nx, i1), y1 = gather_f32(ny, i1), x2 = gather_f32(nx, i2), y2 = gather_f32(ny, i2);
        __m512 det = _mm512_sub_ps(_mm512_mul_ps(_mm512_sub_ps(x2, x0), _mm512_sub_ps(y1, y0)), _mm512_mul_ps(_mm512_sub_ps(x0, x1), _mm512_sub_ps(y0, y2)));
        if (cullMode != SWR_CULL_FRONT_CCW) {
            const __mmask16 flip = cullMode == SWR_CULL_FRONT_CW ? (__mmask16)0xFFFF : _mm512_cmp_ps_mask(det, _mm512_setzero_ps(), _CMP_LT_OQ);
            det = _mm512_mask_xor_ps(det, flip, det, _mm512_set1_ps(-0.0f));
        }
        accept &= _mm512_cmp_ps_mask(det, _mm512_setzero_ps(), _CMP_GT_OQ);                                               // :269
        if (accept == 0) continue;

        // GetRenderBoundingBox on packed s16 pairs (:331-351), including the 32-bit carry of the +7
        const __m512i p0 = gather_u32(pos, i0), p1 = gather_u32(pos, i1), p2 = gather_u32(pos, i2);
        __m512i mn = _mm512_min_epi16(_mm512_min_epi16(p0, p1), p2), mx = _mm512_max_epi16(_mm512_max_epi16(p0, p1), p2);
        mn = _mm512_srai_epi16(_mm512_add_epi32(mn, _mm512_set1_epi32(0x00070007)), 4);
        mx = _mm512_srai_epi16(_mm512_add_epi32(mx, _mm512_set1_epi32(0x00070007)), 4);
        mn = _mm512_min_epi16(_mm512_max_epi16(_mm512_add_epi16(mn, vp), _mm512_setzero_si512()), vp2);
        mx = _mm512_min_epi16(_mm512_max_epi16(_mm512_add_epi16(mx, vp), _mm512_setzero_si512()), vp2);
        const __m512i bbMin = _mm512_andnot_si512(_mm512_set1_epi32(0x00030003), mn);
        const __m512i bbMax = _mm512_andnot_si512(_mm512_set1_epi32(0x00030003), _mm512_add_epi32(mx, _mm512_set1_epi32(0x00030003)));
        const uint32_t ge = (uint32_t)_mm512_cmpge_epi16_mask(bbMin, bbMax);        // 2 bits per triangle: x half, y half (:283)
        const __mmask16 empty = (__mmask16)_pext_u32(ge | (ge >> 1), 0x55555555u);
        accept &= (__mmask16)~empty;
        _mm512_storeu_si512(bbMinOut + prim0, bbMin);
        _mm512_storeu_si512(bbMaxOut + prim0, bbMax);
        keep[prim0 >> 5] |= (uint32_t)accept << (prim0 & 31);
    }
    return nClip;
}
