// resolve_avx512.cpp — ShadingContext::Resolve restated with the reference's own granularity: one 4x4 fragment =
// 16 AVX-512 lanes (Shading.cpp:658-689, DispatchPass Rasterizer.h:225-242). This is the resolve half of the TIMED CPU
// baseline (oracle/baseline_mt.cpp, bench.py cpu_baseline / --impl reference). TEST INFRASTRUCTURE ONLY.
//
// It evaluates exactly the operation sequence of the scalar spec (oracle_resolve.cpp: same IEEE operations in the same
// order, an FMA where the spec has std::fmaf, IEEE divisions and square roots for approx_rcp / approx_rsqrt), so its
// output must equal the spec's bit for bit (tests/test_baseline_cpu.py) — only 16 pixels at a time, which is what
// makes it a fair stand-in for the reference's speed. Like the reference, the per-surface-id work that cannot be
// expressed on vectors stays a scalar "waterfall": the vertex fetch per lane (Shading.cpp:482-507) and the texture
// sampling per lane inside the per-material loop (:532-545).
// Compiled on its own with -march=x86-64-v4 (oracle/Makefile); callers check the CPU before calling in.
#include <immintrin.h>

#include "resolve_common.h"

using namespace orc_detail;

namespace {

struct F {                       // 16 floats
    __m512 v;
    F() = default;
    F(__m512 x) : v(x) {}
    explicit F(float x) : v(_mm512_set1_ps(x)) {}
};
inline F operator+(F a, F b) { return _mm512_add_ps(a.v, b.v); }
inline F operator-(F a, F b) { return _mm512_sub_ps(a.v, b.v); }
inline F operator*(F a, F b) { return _mm512_mul_ps(a.v, b.v); }
inline F operator/(F a, F b) { return _mm512_div_ps(a.v, b.v); }
inline F operator-(F a) { return _mm512_xor_ps(a.v, _mm512_set1_ps(-0.0f)); }
inline F operator+(F a, float b) { return a + F(b); }
inline F operator-(F a, float b) { return a - F(b); }
inline F operator*(F a, float b) { return a * F(b); }
inline F operator-(float a, F b) { return F(a) - b; }
inline F operator*(float a, F b) { return F(a) * b; }
inline F operator/(float a, F b) { return F(a) / b; }
inline F fma(F a, F b, F c) { return _mm512_fmadd_ps(a.v, b.v, c.v); }
inline F fma(F a, float b, F c) { return fma(a, F(b), c); }
inline F vmax(F a, F b) { return _mm512_max_ps(a.v, b.v); }     // (differs from std::fmax only in the sign of a zero result)
inline F vmin(F a, F b) { return _mm512_min_ps(a.v, b.v); }
inline F vabs(F a) { return _mm512_andnot_ps(_mm512_set1_ps(-0.0f), a.v); }
inline F vsqrt(F a) { return _mm512_sqrt_ps(a.v); }
inline F from_int(__m512i i) { return _mm512_cvtepi32_ps(i); }
inline __m512i bits(F a) { return _mm512_castps_si512(a.v); }
inline F from_bits(__m512i i) { return _mm512_castsi512_ps(i); }
inline __m512i and_i(__m512i a, uint32_t m) { return _mm512_and_si512(a, _mm512_set1_epi32((int)m)); }
inline F load(const float* p) { return _mm512_loadu_ps(p); }
inline __m512i loadi(const uint32_t* p) { return _mm512_loadu_si512(p); }

inline F v_rcp(F x) { return 1.0f / x; }                                    // approx_rcp   (canonical: IEEE)
inline F v_rsqrt(F x) { return 1.0f / vsqrt(x); }                           // approx_rsqrt
inline F v_lerp(F a, F b, F t) { return fma(t, b, fma(-t, a, a)); }         // SIMD.h:445
inline F v_clamp01(F x) { return vmin(vmax(x, F(0.0f)), F(1.0f)); }
inline F v_mulsign(F x, F y) { return from_bits(_mm512_xor_si512(bits(x), and_i(bits(y), 0x80000000u))); }   // SIMD.h:341-344

struct F3 { F x, y, z; };
inline F v_dot3(F3 a, F3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }                 // SIMD.h:437
inline F3 v_normalize3(F3 a) { F r = v_rsqrt(v_dot3(a, a)); return { a.x * r, a.y * r, a.z * r }; }
inline F3 v_cross3(F3 a, F3 b) { return { fma(a.y, b.z, -a.z * b.y), fma(a.z, b.x, -a.x * b.z), fma(a.x, b.y, -a.y * b.x) }; }
inline F3 v_mul_mat3(const float* m, F3 n) {                                                   // SIMD.h:465-471
    return { fma(n.x, m[0], fma(n.y, m[3], n.z * m[6])), fma(n.x, m[1], fma(n.y, m[4], n.z * m[7])), fma(n.x, m[2], fma(n.y, m[5], n.z * m[8])) };
}
inline F v_bary_lerp(const F b[3], F v0, F v1, F v2) { return fma(v0, b[0], fma(v1, b[1], v2 * b[2])); }   // Rasterizer.h:101-104
inline F3 v_unmap_octahedron(F u, F v) {                                                        // Texture.h:289-296
    u = u * 2.0f - 1.0f; v = v * 2.0f - 1.0f;
    F3 n = { u, v, 1.0f - vabs(u) - vabs(v) };
    F t = vmax(-n.z, F(0.0f));
    n.x = n.x - v_mulsign(t, n.x);
    n.y = n.y - v_mulsign(t, n.y);
    return v_normalize3(n);
}
inline void v_unpack_normal_tangent(__m512i p, F3& n, F3& t) {                                  // Shading.cpp:232-236
    const float s = 1.0f / 255;
    F a = from_int(and_i(p, 255)) * s, b = from_int(and_i(_mm512_srli_epi32(p, 8), 255)) * s;
    F c = from_int(and_i(_mm512_srli_epi32(p, 16), 255)) * s, d = from_int(_mm512_srli_epi32(p, 24)) * s;
    n = v_unmap_octahedron(a, b);
    t = v_unmap_octahedron(c, d);
}
inline F v_pow5(F x) { return (x * x) * (x * x) * x; }
inline __m512i v_pack_channel(F v) {                                                            // RGBA8u::Pack, Texture.h:55-67
    __m512i i = _mm512_cvtps_epi32((v * 255.0f).v);                                             // RNE, 0x80000000 on overflow / NaN
    i = _mm512_min_epi32(_mm512_max_epi32(i, _mm512_set1_epi32(-32768)), _mm512_set1_epi32(32767));
    return _mm512_min_epi32(_mm512_max_epi32(i, _mm512_setzero_si512()), _mm512_set1_epi32(255));
}

}  // namespace

extern "C" void orc_resolve_rows_avx512(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                                        const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                                        const swr_light* lights, uint32_t numLights,
                                        const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                                        const float* viewPos, float exposure, uint32_t yBegin, uint32_t yEnd) {
    const float scaleU = 2.0f / (float)width, scaleV = 2.0f / (float)height;       // Rasterizer.h:226
    const float centerU = 0.5f * scaleU - 1.0f, centerV = 0.5f * scaleV - 1.0f;    // :227
    const float lightExposure = exposure * 0.001f;                                  // Shading.cpp:674
    const float kx = 2.0f / (float)width, ky = -(2.0f / (float)height);             // :454-457
    const __m512i laneX = _mm512_setr_epi32(0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3);   // Rasterizer.h:247-248
    const __m512i laneY = _mm512_setr_epi32(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3);
    const float* M = objectToClip;
    const float* P = invScreenProj;

    for (uint32_t y0 = yBegin; y0 < yEnd && y0 < height; y0 += 4) {
        for (uint32_t x0 = 0; x0 < width; x0 += 4) {
            const uint32_t tileOffset = ((x0 & ~3u) << 2) + (y0 & ~3u) * width;     // Rasterizer.h:50-56
            uint32_t* tileData = color + tileOffset;
            const F tileDepth = load(depth + tileOffset);
            const __mmask16 skyMask = _mm512_cmp_ps_mask(tileDepth.v, _mm512_setzero_ps(), _CMP_LE_OQ);   // Shading.cpp:664
            const __mmask16 surf = (__mmask16)~skyMask;
            if (surf == 0) {                                                        // tonemap(0) packs to opaque black
                _mm512_storeu_si512(tileData, _mm512_set1_epi32((int)0xFF000000u));
                continue;
            }
            const F px = from_int(_mm512_add_epi32(_mm512_set1_epi32((int)x0), laneX)), py = from_int(_mm512_add_epi32(_mm512_set1_epi32((int)y0), laneY));
            const F screenU = px * scaleU + centerU, screenV = py * scaleV + centerV;                     // Rasterizer.h:237
            F3 worldPos;
            {
                const F z = _mm512_mask_blend_ps(skyMask, tileDepth.v, _mm512_set1_ps(1.0f));            // :666
                F h[4];
                for (int r = 0; r < 4; r++) h[r] = fma(px, P[r], fma(py, P[4 + r], fma(z, P[8 + r], F(P[12 + r]))));
                const F rw = 1.0f / h[3];
                worldPos = { h[0] * rw, h[1] * rw, h[2] * rw };
            }

            // ---- ResolveSurface: per-lane vertex fetch (the reference's waterfall over surface ids, Shading.cpp:482-507)
            alignas(64) float pos[3][3][N] = {};
            alignas(64) uint32_t packedTC[3][N] = {}, packedNT[3][N] = {}, handedA[N] = {}, materialId[N];
            for (int i = 0; i < N; i++) {
                materialId[i] = SWR_NO_MATERIAL;
                if (!((surf >> i) & 1)) continue;
                const uint32_t sid = tileData[i];
                const swr_meshlet& mesh = meshlets[sid / SWR_MAX_PRIMS];
                const uint32_t tri = sid % SWR_MAX_PRIMS;
                for (int vi = 0; vi < 3; vi++) {
                    const uint32_t idx = mesh.Indices[vi][tri] & 63;
                    pos[vi][0][i] = mesh.Positions[0][idx]; pos[vi][1][i] = mesh.Positions[1][idx]; pos[vi][2][i] = mesh.Positions[2][idx];
                    packedTC[vi][i] = mesh.TexCoords[idx];
                    packedNT[vi][i] = mesh.NormalTangents[idx];
                    if (vi == 0 && ((mesh.TangentHandedness >> idx) & 1)) handedA[i] = 1u << 31;          // :500-502
                }
                materialId[i] = mesh.MaterialId;
            }

            // ---- clip transform (:509-511) + IntersectTriangle (:417-464), 16 pixels at a time
            F cx[3], cy[3], invW[3];
            for (int vi = 0; vi < 3; vi++) {
                const F x = load(pos[vi][0]), y = load(pos[vi][1]), z = load(pos[vi][2]);
                cx[vi] = fma(x, M[0], fma(y, M[4], fma(z, M[8], F(M[12]))));
                cy[vi] = fma(x, M[1], fma(y, M[5], fma(z, M[9], F(M[13]))));
                const F cw = fma(x, M[3], fma(y, M[7], fma(z, M[11], F(M[15]))));
                invW[vi] = 1.0f / cw;
            }
            const F p0x = cx[0] * invW[0], p0y = cy[0] * invW[0], p1x = cx[1] * invW[1], p1y = cy[1] * invW[1], p2x = cx[2] * invW[2], p2y = cy[2] * invW[2];
            const F m0x = p2x - p1x, m0y = p2y - p1y, m1x = p0x - p1x, m1y = p0y - p1y;
            const F invDet = 1.0f / (m0x * m1y - m1x * m0y);
            const F dxv[3] = { p1y - p2y, p2y - p0y, p0y - p1y }, dyv[3] = { p2x - p1x, p0x - p2x, p1x - p0x };
            F sx[3], sy[3];
            for (int k = 0; k < 3; k++) { sx[k] = dxv[k] * (invDet * invW[k]); sy[k] = dyv[k] * (invDet * invW[k]); }
            F dsum = sx[0] + sx[1] + sx[2], esum = sy[0] + sy[1] + sy[2];
            const F rel0x = screenU - p0x, rel0y = screenV - p0y;
            const F interpInvW = invW[0] + rel0x * dsum + rel0y * esum;
            const F interpW = 1.0f / interpInvW;
            F bary[3];
            bary[1] = interpW * (rel0x * sx[1] + rel0y * sy[1]);
            bary[2] = interpW * (rel0x * sx[2] + rel0y * sy[2]);
            bary[0] = 1.0f - bary[1] - bary[2];
            for (int k = 0; k < 3; k++) { sx[k] = sx[k] * kx; sy[k] = sy[k] * ky; }
            dsum = dsum * kx; esum = esum * ky;
            const F interpW_ddx = 1.0f / (interpInvW + dsum), interpW_ddy = 1.0f / (interpInvW + esum);
            F ddx[3], ddy[3];
            for (int k = 0; k < 3; k++) {
                ddx[k] = interpW_ddx * (bary[k] * interpInvW + sx[k]) - bary[k];
                ddy[k] = interpW_ddy * (bary[k] * interpInvW + sy[k]) - bary[k];
            }

            // ---- UVs and UV gradients (:516-527); RG16f::Unpack = vcvtph2ps (Texture.h:109-124)
            alignas(64) float texU[N], texV[N], texGrad[4][N];
            {
                F tu[3], tv[3];
                for (int vi = 0; vi < 3; vi++) {
                    const __m512i tc = loadi(packedTC[vi]);
                    tu[vi] = _mm512_cvtph_ps(_mm512_cvtepi32_epi16(tc));
                    tv[vi] = _mm512_cvtph_ps(_mm512_cvtepi32_epi16(_mm512_srli_epi32(tc, 16)));
                }
                const F t10u = tu[1] - tu[0], t10v = tv[1] - tv[0], t20u = tu[2] - tu[0], t20v = tv[2] - tv[0];
                _mm512_store_ps(texU, (tu[0] + t10u * bary[1] + t20u * bary[2]).v);
                _mm512_store_ps(texV, (tv[0] + t10v * bary[1] + t20v * bary[2]).v);
                _mm512_store_ps(texGrad[0], (t10u * ddx[1] + t20u * ddx[2]).v);
                _mm512_store_ps(texGrad[1], (t10v * ddx[1] + t20v * ddx[2]).v);
                _mm512_store_ps(texGrad[2], (t10u * ddy[1] + t20u * ddy[2]).v);
                _mm512_store_ps(texGrad[3], (t10v * ddy[1] + t20v * ddy[2]).v);
            }

            // ---- material waterfall (:532-545): scalar per lane, like the reference's loop over distinct materials
            alignas(64) uint32_t packedAlbedo[N] = {}, packedNMR[N] = {};
            {
                bool done[N];
                for (int i = 0; i < N; i++) done[i] = !((surf >> i) & 1) || materialId[i] == SWR_NO_MATERIAL;
                for (int i = 0; i < N; i++) {
                    if (done[i]) continue;
                    const uint32_t id = materialId[i];
                    const swr_texture_desc& tex = textures[materials[id].TextureId];
                    int32_t mip[N];
                    bool anyMin = false;
                    for (int l = 0; l < N; l++) {
                        if (!((surf >> l) & 1)) continue;                      // pinned: see oracle_resolve.cpp header
                        const float g[4] = { texGrad[0][l], texGrad[1][l], texGrad[2][l], texGrad[3][l] };
                        mip[l] = calc_mip_level(g, (float)tex.Width, (float)tex.Height);
                        anyMin = anyMin || mip[l] > 0;                         // Texture.h:432
                    }
                    for (int l = 0; l < N; l++) {
                        if (!((surf >> l) & 1) || materialId[l] != id) continue;
                        packedAlbedo[l] = sample_level(tex, texU[l], texV[l], 0, mip[l], anyMin);
                        if (tex.NumLayers >= 2) packedNMR[l] = sample_level(tex, texU[l], texV[l], 1, mip[l], anyMin);
                        done[l] = true;
                    }
                }
            }
            const __m512i nmr = loadi(packedNMR), albedo = loadi(packedAlbedo);
            const bool anyNormalMap = (_mm512_test_epi32_mask(nmr, _mm512_set1_epi32(0xFFFF)) & surf) != 0;   // :554

            // ---- normals (:547-577)
            F3 normal;
            {
                F3 n0, n1, n2, t0, t1, t2;
                v_unpack_normal_tangent(loadi(packedNT[0]), n0, t0);
                v_unpack_normal_tangent(loadi(packedNT[1]), n1, t1);
                v_unpack_normal_tangent(loadi(packedNT[2]), n2, t2);
                const F3 nl = { v_bary_lerp(bary, n0.x, n1.x, n2.x), v_bary_lerp(bary, n0.y, n1.y, n2.y), v_bary_lerp(bary, n0.z, n1.z, n2.z) };
                const F3 normalWS = v_normalize3(v_mul_mat3(objectToWorld3, nl));
                normal = normalWS;
                if (anyNormalMap) {
                    const F3 tl = { v_bary_lerp(bary, t0.x, t1.x, t2.x), v_bary_lerp(bary, t0.y, t1.y, t2.y), v_bary_lerp(bary, t0.z, t1.z, t2.z) };
                    const F3 tangentWS = v_normalize3(v_mul_mat3(objectToWorld3, tl));
                    F3 bit = v_cross3(normalWS, tangentWS);
                    const __m512i handed = loadi(handedA);
                    bit = { from_bits(_mm512_xor_si512(bits(bit.x), handed)), from_bits(_mm512_xor_si512(bits(bit.y), handed)), from_bits(_mm512_xor_si512(bits(bit.z), handed)) };
                    const F nx = from_int(and_i(nmr, 255)) * (1.0f / 127.5f) - 1.0f;
                    const F ny = from_int(and_i(_mm512_srli_epi32(nmr, 8), 255)) * (1.0f / 127.5f) - 1.0f;
                    const F nz2 = 1.0f - (nx * nx + ny * ny);
                    const F nz = v_rsqrt(nz2) * nz2;                                                   // approx_sqrt (SIMD.h:296)
                    normal = v_normalize3({ nx * tangentWS.x + ny * bit.x + nz * normalWS.x,
                                            nx * tangentWS.y + ny * bit.y + nz * normalWS.y,
                                            nx * tangentWS.z + ny * bit.z + nz * normalWS.z });
                }
            }
            const F metallic = from_int(and_i(_mm512_srli_epi32(nmr, 16), 255)) * (1.0f / 255);
            const F roughness = from_int(and_i(_mm512_srli_epi32(nmr, 24), 255)) * (1.0f / 255);

            // ---- EvalLighting (:602-645)
            F base[3];
            {   // RGBA8u::UnpackSrgb (Texture.h:37-54)
                const __m512i rb1 = _mm512_add_epi32(and_i(_mm512_slli_epi32(albedo, 8), 0xFF00FF00u), _mm512_set1_epi32(0x00FF00FF));
                const __m512i ag1 = _mm512_add_epi32(and_i(albedo, 0xFF00FF00u), _mm512_set1_epi32(0x00FF00FF));
                const __m512i rb2 = _mm512_mulhi_epu16(rb1, rb1), ag2 = _mm512_mulhi_epu16(ag1, ag1);
                const float scale = 1.0f / 65535;
                base[0] = from_int(and_i(rb2, 65535)) * scale;
                base[1] = from_int(and_i(ag2, 65535)) * scale;
                base[2] = from_int(_mm512_srli_epi32(rb2, 16)) * scale;
            }
            const float reflectance = 0.5f;
            const float f0c = 0.16f * reflectance * reflectance;
            const F alphaRoughness = vmax(roughness * roughness, F(1e-4f));
            F f0[3], diffuse[3], c[3];
            for (int k = 0; k < 3; k++) {
                f0[k] = v_lerp(F(f0c), base[k], metallic);
                diffuse[k] = base[k] * (1.0f - metallic);
                c[k] = F(0.0f);
            }
            const F3 viewDir = v_normalize3({ viewPos[0] - worldPos.x, viewPos[1] - worldPos.y, viewPos[2] - worldPos.z });
            const F NoV = vabs(v_dot3(normal, viewDir)) + 1e-5f;
            for (uint32_t li = 0; li < numLights; li++) {
                const swr_light& light = lights[li];
                F3 lightDir;
                if (light.Type == 0) lightDir = { F(-light.Direction[0]), F(-light.Direction[1]), F(-light.Direction[2]) };
                else lightDir = v_normalize3({ light.Position[0] - worldPos.x, light.Position[1] - worldPos.y, light.Position[2] - worldPos.z });
                const F NoL = v_dot3(normal, lightDir);
                if (((__mmask16)~_mm512_cmp_ps_mask(NoL.v, _mm512_set1_ps(1e-4f), _CMP_LT_OQ) & surf) == 0) continue;     // :620
                F attenuation(1.0f);                                                               // GetLightAttenuation :581-600
                if (light.Type != 0) {
                    const F3 ptl = { light.Position[0] - worldPos.x, light.Position[1] - worldPos.y, light.Position[2] - worldPos.z };
                    const F d2 = v_dot3(ptl, ptl);
                    const F factor = d2 * light.InvRadiusSq;
                    const F smooth = vmax(1.0f - factor * factor, F(0.0f));
                    attenuation = (smooth * smooth) * v_rcp(vmax(d2, F(1e-4f)));
                    if (light.Type == 2) {
                        const F3 nl = v_normalize3(ptl);
                        const F cd = v_dot3({ F(-light.Direction[0]), F(-light.Direction[1]), F(-light.Direction[2]) }, nl);
                        const F spot = v_clamp01(cd * light.SpotScale + light.SpotOffset);
                        attenuation = attenuation * (spot * spot);
                    }
                }
                attenuation = attenuation * light.Intensity * lightExposure;
                if (((__mmask16)~_mm512_cmp_ps_mask((NoL * attenuation).v, _mm512_set1_ps(1e-4f), _CMP_LT_OQ) & surf) == 0) continue;   // :623
                const F3 halfway = v_normalize3({ viewDir.x + lightDir.x, viewDir.y + lightDir.y, viewDir.z + lightDir.z });
                const F NoH = v_clamp01(v_dot3(normal, halfway)), LoH = v_clamp01(v_dot3(lightDir, halfway));
                F D, V;
                {   // D_GGX, V_SmithGGXCorrelatedFast (Shading.cpp:19-28)
                    const F a = NoH * alphaRoughness;
                    const F k = alphaRoughness * v_rcp(1.0f - NoH * NoH + a * a);
                    D = k * k * kInvPi;
                    const F a2 = 2.0f * NoL * NoV, b2 = NoL + NoV;
                    V = 0.5f / v_lerp(a2, b2, alphaRoughness);
                }
                const F weight = vmax(NoL * attenuation, F(0.0f));
                const F f = v_pow5(1.0f - LoH);                                                    // F_Schlick :29-32
                for (int k = 0; k < 3; k++) {
                    const F Fk = f + f0[k] * (1.0f - f);
                    const F Fr = (D * V) * Fk;
                    const F Fd = diffuse[k] * kInvPi;
                    c[k] = c[k] + (Fd + Fr) * light.Color[k] * weight;
                }
            }

            // ---- ambient (:642), Tonemap_Unreal (:221-226), RGBA8u::Pack; sky lanes resolve to colour 0
            __m512i packed = _mm512_set1_epi32((int)0xFF000000u);
            for (int k = 0; k < 3; k++) {
                F out = c[k] + base[k] * 0.05f;
                out = _mm512_maskz_mov_ps(surf, out.v);
                const F x = out * exposure;
                const F o = x / (x + 0.155f) * 1.019f;
                packed = _mm512_or_si512(packed, _mm512_slli_epi32(v_pack_channel(o), 8 * k));
            }
            _mm512_storeu_si512(tileData, packed);
        }
    }
}
