// oracle_resolve.cpp — CPU restatement of the vis-buffer resolve pass (ShadingContext::Resolve).
//
// TEST INFRASTRUCTURE ONLY (see oracle.cpp header). PARITY PINNED: bit-identical colour with the reference's
// own Shading.cpp built by oracle/ref_build.py (tests/test_ref_pin.py). Follows, per 4x4 fragment with 16 lanes like the reference:
//   ShadingContext::Resolve            Shading.cpp:658-689   (+ DispatchPass Rasterizer.h:225-242)
//   ResolveSurface / IntersectTriangle Shading.cpp:472-579 / :417-464
//   EvalLighting / GetLightAttenuation Shading.cpp:602-645 / :581-600, BRDF helpers :17-33
//   Texture2D::SampleLevel/SampleLinear Texture.h:412-459, :506-575; CalcMipLevel :276-280
//   pixfmt::RGBA8u::{Unpack,UnpackSrgb,Pack} Texture.h:28-67; RG16f::Unpack :109-124
//   texutil::UnmapOctahedron           Texture.h:289-296
// Canonical arithmetic: approx_rcp -> 1/x, approx_rsqrt -> 1/sqrt(x) (the AVX-512 14-bit tables are
// not reproducible elsewhere; the colour gate is a tolerance: max abs error <= 2/255, PSNR >= 50 dB).
// Two reference behaviours that read uninitialised lanes are pinned:
//   * the nearest-vs-bilinear choice `any(mipLevel > 0)` spans the NON-SKY lanes of the tile;
//   * sky lanes (depth <= 0) always resolve to colour 0 (there is no skybox on this path).
#include <cmath>
#include "resolve_common.h"

using namespace orc_detail;

static void resolve_rows_impl(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                              const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                              const swr_light* lights, uint32_t numLights,
                              const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                              const float* viewPos, float exposure, uint32_t yBegin, uint32_t yEnd, int debugLayer,
                              const swr_texture_desc* skybox = nullptr);

extern "C" {

// ShadingContext::Resolve over the whole framebuffer. color/depth are the 4x4-tiled layers 0/1;
// layer 0 holds surface ids on entry and RGBA8 colour on exit (Shading.cpp:688).
// Rows [yBegin, yEnd) (multiples of 4) only, so the threaded baseline can split the pass like
// Rasterizer::Dispatch does (Rasterizer.cpp:822-846).
void orc_resolve_rows(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                      const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                      const swr_light* lights, uint32_t numLights,
                      const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                      const float* viewPos, float exposure, uint32_t yBegin, uint32_t yEnd) {
    resolve_rows_impl(color, depth, width, height, meshlets, materials, textures, lights, numLights, objectToClip,
                      objectToWorld3, invScreenProj, viewPos, exposure, yBegin, yEnd, kLayerNone);
}

// ShadingContext::ResolveDebug (Shading.cpp:734-773): layer 0 is overwritten with the visualisation of `layer`
// (DebugLayer, Shading.h:8); sky pixels (depth <= 0) get a 4x4 checkerboard of 0xFFA0A0A0 / 0xFFFFFFFF.
void orc_resolve_debug(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                       const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                       const float* objectToClip, const float* objectToWorld3, const float* invScreenProj, int layer) {
    const float zero[3] = { 0, 0, 0 };
    resolve_rows_impl(color, depth, width, height, meshlets, materials, textures, nullptr, 0, objectToClip,
                      objectToWorld3, invScreenProj, zero, 1.0f, 0, height, layer);
}

}  // extern "C"

static void resolve_rows_impl(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                              const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                              const swr_light* lights, uint32_t numLights,
                              const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                              const float* viewPos, float exposure, uint32_t yBegin, uint32_t yEnd, int debugLayer,
                              const swr_texture_desc* skybox) {
    const float scaleU = 2.0f / (float)width, scaleV = 2.0f / (float)height;       // Rasterizer.h:226
    const float centerU = 0.5f * scaleU - 1.0f, centerV = 0.5f * scaleV - 1.0f;    // :227
    const V3 view = { viewPos[0], viewPos[1], viewPos[2] };
    const float lightExposure = exposure * 0.001f;                                  // Shading.cpp:674

    for (uint32_t y0 = yBegin; y0 < yEnd && y0 < height; y0 += 4) {
        for (uint32_t x0 = 0; x0 < width; x0 += 4) {
            uint32_t tileOffset = ((x0 & ~3u) << 2) + (y0 & ~3u) * width;           // Rasterizer.h:50-56
            const float* tileDepth = depth + tileOffset;
            uint32_t* tileData = color + tileOffset;
            bool sky[N];
            bool anySurface = false;
            V3 worldPos[N];
            float screenU[N], screenV[N];
            for (int i = 0; i < N; i++) {
                sky[i] = tileDepth[i] <= 0.0f;                                      // :664
                anySurface = anySurface || !sky[i];
                float px = (float)(int32_t)(x0 + (i & 3)), py = (float)(int32_t)(y0 + (i >> 2));
                screenU[i] = px * scaleU + centerU;                                 // Rasterizer.h:237
                screenV[i] = py * scaleV + centerV;
                float h[4];
                mul_mat4(invScreenProj, px, py, sky[i] ? 1.0f : tileDepth[i], 1.0f, h);   // :666-667
                float rw = 1.0f / h[3];
                worldPos[i] = { h[0] * rw, h[1] * rw, h[2] * rw };
            }
            float outColor[N][3];
            for (int i = 0; i < N; i++) outColor[i][0] = outColor[i][1] = outColor[i][2] = 0.0f;

            if (debugLayer >= kLayerMeshletId) {                                    // Shading.cpp:755-766: straight from the colour word
                for (int i = 0; i < N; i++) {
                    uint32_t d = tileData[i];
                    float* o = outColor[i];
                    if (debugLayer == kLayerMeshletId) unpack_rgba8((d / SWR_MAX_PRIMS) * 123456789u, o);
                    else if (debugLayer == kLayerTriangleId) unpack_rgba8(d * 123456789u, o);
                    else if (debugLayer == kLayerOverdrawPixel) colormap_turbo((float)(d >> 16) / 30.0f, o);
                    else colormap_turbo(((float)(d >> 16) + (float)(d & 0xFFFF) * 0.5f) / 30.0f, o);
                }
            } else if (anySurface) {
                // ---- ResolveSurface (Shading.cpp:472-579)
                float bary[N][3], ddx[N][3], ddy[N][3];
                uint32_t packedTC[N][3], packedNT[N][3], handed[N], materialId[N];
                float texU[N], texV[N], texGrad[N][4];
                for (int i = 0; i < N; i++) {
                    if (sky[i]) continue;
                    uint32_t sid = tileData[i];
                    const swr_meshlet& mesh = meshlets[sid / SWR_MAX_PRIMS];
                    uint32_t tri = sid % SWR_MAX_PRIMS;
                    float clip[3][4];
                    handed[i] = 0;
                    for (int vi = 0; vi < 3; vi++) {
                        uint32_t idx = mesh.Indices[vi][tri];
                        mul_mat4(objectToClip, mesh.Positions[0][idx & 63], mesh.Positions[1][idx & 63], mesh.Positions[2][idx & 63], 1.0f, clip[vi]);   // :509-511
                        packedTC[i][vi] = mesh.TexCoords[idx & 63];
                        packedNT[i][vi] = mesh.NormalTangents[idx & 63];
                        if (vi == 0 && ((mesh.TangentHandedness >> (idx & 63)) & 1)) handed[i] = 1u << 31;   // :500-502
                    }
                    materialId[i] = mesh.MaterialId;

                    // IntersectTriangle (Shading.cpp:417-464)
                    float invW[3] = { 1.0f / clip[0][3], 1.0f / clip[1][3], 1.0f / clip[2][3] };
                    float p0x = clip[0][0] * invW[0], p0y = clip[0][1] * invW[0];
                    float p1x = clip[1][0] * invW[1], p1y = clip[1][1] * invW[1];
                    float p2x = clip[2][0] * invW[2], p2y = clip[2][1] * invW[2];
                    float m0x = p2x - p1x, m0y = p2y - p1y, m1x = p0x - p1x, m1y = p0y - p1y;
                    float invDet = 1.0f / (m0x * m1y - m1x * m0y);
                    float sx[3], sy[3], dsum = 0, esum = 0;
                    float dxv[3] = { p1y - p2y, p2y - p0y, p0y - p1y }, dyv[3] = { p2x - p1x, p0x - p2x, p1x - p0x };
                    for (int k = 0; k < 3; k++) { sx[k] = dxv[k] * (invDet * invW[k]); sy[k] = dyv[k] * (invDet * invW[k]); }
                    dsum = sx[0] + sx[1] + sx[2];
                    esum = sy[0] + sy[1] + sy[2];
                    float rel0x = screenU[i] - p0x, rel0y = screenV[i] - p0y;
                    float interpInvW = invW[0] + rel0x * dsum + rel0y * esum;
                    float interpW = 1.0f / interpInvW;
                    bary[i][1] = interpW * (rel0x * sx[1] + rel0y * sy[1]);
                    bary[i][2] = interpW * (rel0x * sx[2] + rel0y * sy[2]);
                    bary[i][0] = 1 - bary[i][1] - bary[i][2];
                    float kx = 2.0f / (float)width, ky = -(2.0f / (float)height);   // :454-457
                    for (int k = 0; k < 3; k++) { sx[k] *= kx; sy[k] *= ky; }
                    dsum *= kx; esum *= ky;
                    float interpW_ddx = 1.0f / (interpInvW + dsum), interpW_ddy = 1.0f / (interpInvW + esum);
                    for (int k = 0; k < 3; k++) {
                        ddx[i][k] = interpW_ddx * (bary[i][k] * interpInvW + sx[k]) - bary[i][k];
                        ddy[i][k] = interpW_ddy * (bary[i][k] * interpInvW + sy[k]) - bary[i][k];
                    }

                    // UVs and UV gradients (:516-527); RG16f::Unpack (Texture.h:109-124)
                    float t0u = half2float((uint16_t)packedTC[i][0]), t0v = half2float((uint16_t)(packedTC[i][0] >> 16));
                    float t10u = half2float((uint16_t)packedTC[i][1]) - t0u, t10v = half2float((uint16_t)(packedTC[i][1] >> 16)) - t0v;
                    float t20u = half2float((uint16_t)packedTC[i][2]) - t0u, t20v = half2float((uint16_t)(packedTC[i][2] >> 16)) - t0v;
                    texU[i] = t0u + t10u * bary[i][1] + t20u * bary[i][2];
                    texV[i] = t0v + t10v * bary[i][1] + t20v * bary[i][2];
                    texGrad[i][0] = t10u * ddx[i][1] + t20u * ddx[i][2];
                    texGrad[i][1] = t10v * ddx[i][1] + t20v * ddx[i][2];
                    texGrad[i][2] = t10u * ddy[i][1] + t20u * ddy[i][2];
                    texGrad[i][3] = t10v * ddy[i][1] + t20v * ddy[i][2];
                }

                // material waterfall (:532-545)
                uint32_t packedAlbedo[N] = {}, packedNMR[N] = {};
                bool done[N];
                for (int i = 0; i < N; i++) done[i] = sky[i] || materialId[i] == SWR_NO_MATERIAL;
                for (int i = 0; i < N; i++) {
                    if (done[i]) continue;
                    uint32_t id = materialId[i];
                    const swr_texture_desc& tex = textures[materials[id].TextureId];
                    int32_t mip[N];
                    bool anyMin = false;
                    for (int l = 0; l < N; l++) {
                        if (sky[l]) continue;                                  // pinned: see header
                        mip[l] = calc_mip_level(texGrad[l], (float)tex.Width, (float)tex.Height);
                        anyMin = anyMin || mip[l] > 0;                         // Texture.h:432
                    }
                    for (int l = 0; l < N; l++) {
                        if (sky[l] || materialId[l] != id) continue;
                        packedAlbedo[l] = sample_level(tex, texU[l], texV[l], 0, mip[l], anyMin);
                        if (tex.NumLayers >= 2) packedNMR[l] = sample_level(tex, texU[l], texV[l], 1, mip[l], anyMin);
                        done[l] = true;
                    }
                }

                bool anyNormalMap = false;
                for (int i = 0; i < N; i++) anyNormalMap = anyNormalMap || (!sky[i] && (packedNMR[i] & 0xFFFF) != 0);   // :554

                Surface surf[N];
                for (int i = 0; i < N; i++) {
                    if (sky[i]) continue;
                    V3 n0, n1, n2, t0, t1, t2;
                    unpack_normal_tangent(packedNT[i][0], n0, t0);
                    unpack_normal_tangent(packedNT[i][1], n1, t1);
                    unpack_normal_tangent(packedNT[i][2], n2, t2);
                    V3 nl = { bary_lerp(bary[i], n0.x, n1.x, n2.x), bary_lerp(bary[i], n0.y, n1.y, n2.y), bary_lerp(bary[i], n0.z, n1.z, n2.z) };
                    V3 normalWS = normalize3(mul_mat3(objectToWorld3, nl));
                    V3 normal = normalWS;
                    if (anyNormalMap) {
                        V3 tl = { bary_lerp(bary[i], t0.x, t1.x, t2.x), bary_lerp(bary[i], t0.y, t1.y, t2.y), bary_lerp(bary[i], t0.z, t1.z, t2.z) };
                        V3 tangentWS = normalize3(mul_mat3(objectToWorld3, tl));
                        V3 bit = cross3(normalWS, tangentWS);
                        bit = { u2f(f2u(bit.x) ^ handed[i]), u2f(f2u(bit.y) ^ handed[i]), u2f(f2u(bit.z) ^ handed[i]) };
                        float nx = (float)(packedNMR[i] & 255) * (1.0f / 127.5f) - 1.0f;
                        float ny = (float)((packedNMR[i] >> 8) & 255) * (1.0f / 127.5f) - 1.0f;
                        float nz = approx_sqrt(1.0f - (nx * nx + ny * ny));
                        normal = normalize3({ nx * tangentWS.x + ny * bit.x + nz * normalWS.x,
                                              nx * tangentWS.y + ny * bit.y + nz * normalWS.y,
                                              nx * tangentWS.z + ny * bit.z + nz * normalWS.z });
                    }
                    surf[i] = { packedAlbedo[i], normal, (float)((packedNMR[i] >> 16) & 255) * (1.0f / 255),
                                (float)((packedNMR[i] >> 24) & 255) * (1.0f / 255) };
                }

                if (debugLayer != kLayerNone) {                                     // Shading.cpp:749-754
                    for (int i = 0; i < N; i++) {
                        if (sky[i]) continue;
                        float* o = outColor[i];
                        if (debugLayer == kLayerBaseColor) unpack_rgba8(surf[i].albedo, o);
                        else if (debugLayer == kLayerNormals) { o[0] = surf[i].normal.x * 0.5f + 0.5f; o[1] = surf[i].normal.y * 0.5f + 0.5f; o[2] = surf[i].normal.z * 0.5f + 0.5f; }
                        else { o[0] = surf[i].metallic; o[1] = surf[i].roughness; o[2] = 0.0f; }
                    }
                } else {
                // ---- EvalLighting (Shading.cpp:602-645), tile-wide like the reference: the two
                // `simd::all(...) continue` early-outs (:620, :623) span the non-sky lanes (pinned, see header).
                const float reflectance = 0.5f;
                float base[N][4], alphaRoughness[N], f0[N][3], diffuse[N][3], NoV[N], c[N][3];
                V3 viewDir[N];
                for (int i = 0; i < N; i++) {
                    if (sky[i]) continue;
                    unpack_srgb(surf[i].albedo, base[i]);
                    alphaRoughness[i] = std::fmax(surf[i].roughness * surf[i].roughness, 1e-4f);
                    float f0c = 0.16f * reflectance * reflectance;
                    for (int k = 0; k < 3; k++) {
                        f0[i][k] = lerpf(f0c, base[i][k], surf[i].metallic);
                        diffuse[i][k] = base[i][k] * (1.0f - surf[i].metallic);
                        c[i][k] = 0.0f;
                    }
                    viewDir[i] = normalize3({ view.x - worldPos[i].x, view.y - worldPos[i].y, view.z - worldPos[i].z });
                    NoV[i] = std::fabs(dot3(surf[i].normal, viewDir[i])) + 1e-5f;
                }
                for (uint32_t li = 0; li < numLights; li++) {
                    const swr_light& light = lights[li];
                    V3 lightDir[N];
                    float NoL[N], attenuation[N];
                    bool allDark = true;
                    for (int i = 0; i < N; i++) {
                        if (sky[i]) continue;
                        lightDir[i] = light.Type == 0 ? V3{ -light.Direction[0], -light.Direction[1], -light.Direction[2] }
                                                      : normalize3({ light.Position[0] - worldPos[i].x, light.Position[1] - worldPos[i].y, light.Position[2] - worldPos[i].z });
                        NoL[i] = dot3(surf[i].normal, lightDir[i]);
                        allDark = allDark && (NoL[i] < 1e-4f);
                    }
                    if (allDark) continue;                                                   // :620
                    bool allWeak = true;
                    for (int i = 0; i < N; i++) {
                        if (sky[i]) continue;
                        attenuation[i] = light_attenuation(light, worldPos[i]) * light.Intensity * lightExposure;
                        allWeak = allWeak && (NoL[i] * attenuation[i] < 1e-4f);
                    }
                    if (allWeak) continue;                                                   // :623
                    for (int i = 0; i < N; i++) {
                        if (sky[i]) continue;
                        V3 halfway = normalize3({ viewDir[i].x + lightDir[i].x, viewDir[i].y + lightDir[i].y, viewDir[i].z + lightDir[i].z });
                        float NoH = clampf(dot3(surf[i].normal, halfway), 0.0f, 1.0f);
                        float LoH = clampf(dot3(lightDir[i], halfway), 0.0f, 1.0f);
                        float D = D_GGX(NoH, alphaRoughness[i]);
                        float V = V_SmithGGXCorrelatedFast(NoV[i], NoL[i], alphaRoughness[i]);
                        float weight = std::fmax(NoL[i] * attenuation[i], 0.0f);
                        for (int k = 0; k < 3; k++) {
                            float F = F_Schlick1(LoH, f0[i][k]);
                            float Fr = (D * V) * F;
                            float Fd = diffuse[i][k] * kInvPi;
                            c[i][k] += (Fd + Fr) * light.Color[k] * weight;
                        }
                    }
                }
                for (int i = 0; i < N; i++) {
                    if (sky[i]) continue;
                    for (int k = 0; k < 3; k++) outColor[i][k] = c[i][k] + base[i][k] * 0.05f;   // :642
                }
                }
            }
            if (debugLayer != kLayerNone) {                                          // Shading.cpp:768-771
                const uint32_t background = ((x0 ^ y0) & 4) ? 0xFFA0A0A0u : 0xFFFFFFFFu;
                for (int i = 0; i < N; i++)
                    tileData[i] = sky[i] ? background : pack_rgba8(outColor[i][0], outColor[i][1], outColor[i][2], 1.0f);
                continue;
            }
            if (skybox != nullptr) {                                                 // Shading.cpp:676-679
                for (int i = 0; i < N; i++) {
                    if (!sky[i]) continue;
                    sample_skybox(*skybox, { worldPos[i].x - view.x, worldPos[i].y - view.y, worldPos[i].z - view.z }, outColor[i]);
                }
            }
            // tonemap + pack (:680-688, Tonemap_Unreal :221-226)
            for (int i = 0; i < N; i++) {
                float o[3];
                for (int k = 0; k < 3; k++) {
                    float x = outColor[i][k] * exposure;
                    o[k] = x / (x + 0.155f) * 1.019f;
                }
                tileData[i] = pack_rgba8(o[0], o[1], o[2], 1.0f);
            }
        }
    }
}

extern "C" {

// ShadingContext::Resolve with ShadingContext::SkyboxTex set (Shading.h:29, Shading.cpp:676-679): sky pixels take
// SkyboxTex->SampleOctLevel<EnvSampler>(worldPos - ViewPos, 1); `skybox` is a Texture2D<R11G11B10f, TiledY8>.
void orc_resolve_sky(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                     const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                     const swr_light* lights, uint32_t numLights,
                     const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                     const float* viewPos, float exposure, const swr_texture_desc* skybox) {
    resolve_rows_impl(color, depth, width, height, meshlets, materials, textures, lights, numLights, objectToClip,
                      objectToWorld3, invScreenProj, viewPos, exposure, 0, height, kLayerNone, skybox);
}

// Probes for the known-answer tests.
void orc_map_octahedron(const float* dir, float* uv) { map_octahedron({ dir[0], dir[1], dir[2] }, uv[0], uv[1]); }
void orc_sample_skybox(const swr_texture_desc* tex, const float* dir, float* rgb) { sample_skybox(*tex, { dir[0], dir[1], dir[2] }, rgb); }

void orc_resolve(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                 const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                 const swr_light* lights, uint32_t numLights,
                 const float* objectToClip, const float* objectToWorld3, const float* invScreenProj,
                 const float* viewPos, float exposure) {
    orc_resolve_rows(color, depth, width, height, meshlets, materials, textures, lights, numLights, objectToClip,
                     objectToWorld3, invScreenProj, viewPos, exposure, 0, height);
}

// Tail of ShadingContext::Resolve (Shading.cpp:690-731): every point/spot light inside the frustum is drawn as a
// soft disc over the resolved colour where it is not occluded (light depth > stored depth), lights in order.
// glm's mat4 * vec4 is (m0*x + m1*y) + (m2*z + m3*w); vec / scalar is a true division. The reference skips 4x4
// tiles without a pixel inside the disc; pixels outside the disc blend with alpha <= 0 -> 0, which AlphaBlendU8
// (:239-246) leaves unchanged (|fg - bg| * 64 / 32768 rounds to 0), so the pass is written per pixel.
void orc_draw_light_markers(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                            const swr_light* lights, uint32_t numLights, const float* worldToClip) {
    const float* m = worldToClip;
    for (uint32_t li = 0; li < numLights; li++) {
        const swr_light& light = lights[li];
        if (light.Type == 0) continue;                                                          // :693
        float clip[4];
        for (int r = 0; r < 4; r++)
            clip[r] = (m[0 * 4 + r] * light.Position[0] + m[1 * 4 + r] * light.Position[1]) + (m[2 * 4 + r] * light.Position[2] + m[3 * 4 + r] * 1.0f);
        if (std::fmax(std::fmax(std::fabs(clip[0]), std::fabs(clip[1])), std::fabs(clip[2])) > clip[3]) continue;   // :696
        float lightDepth = clip[2] / clip[3];
        float sx = (clip[0] / clip[3]) * 0.5f + 0.5f, sy = (clip[1] / clip[3]) * 0.5f + 0.5f;   // :699-701
        float radius = ((float)(width > height ? width : height) / 30.0f) / clip[3];            // :702
        float cx = sx * (float)width, cy = sy * (float)height;                                  // :704
        int32_t startX = (int32_t)(cx - radius), startY = (int32_t)(cy - radius);               // int2() truncates
        startX = (startX > 0 ? startX : 0) & ~3; startY = (startY > 0 ? startY : 0) & ~3;       // :705
        int32_t endX = (int32_t)(cx + radius), endY = (int32_t)(cy + radius);
        endX = endX < (int32_t)width ? endX : (int32_t)width; endY = endY < (int32_t)height ? endY : (int32_t)height;   // :706
        for (int32_t ty = startY; ty < endY; ty += 4) {
            for (int32_t tx = startX; tx < endX; tx += 4) {
                for (int i = 0; i < 16; i++) {
                    uint32_t x = (uint32_t)tx + (i & 3), y = (uint32_t)ty + (i >> 2);
                    uint32_t off = ((x & ~3u) << 2) + (y & ~3u) * width + (x & 3) + (y & 3) * 4;
                    if (!(lightDepth > depth[off])) continue;                                   // :712, :728
                    float rx = ((float)(int32_t)x + 0.5f) - cx, ry = ((float)(int32_t)y + 0.5f) - cy;   // :715
                    float distSq = std::fmaf(rx, rx, ry * ry) - radius * radius;                // :716
                    float a = 1.0f - (-distSq / (radius * radius));                             // :724
                    uint32_t fg = pack_rgba8(light.Color[0], light.Color[1], light.Color[2], 1.0f - a * a);
                    uint32_t bg = color[off];
                    uint32_t t = (fg >> 1) & 0x7F800000u;                                       // AlphaBlendU8 :239-246
                    t |= (t >> 16) | 0x00400040u;
                    uint32_t rb = lerp16(bg & 0x00FF00FFu, fg & 0x00FF00FFu, t);
                    uint32_t ag = lerp16((bg >> 8) & 0x00FF00FFu, (fg >> 8) & 0x00FF00FFu, t);
                    color[off] = rb | (ag << 8);
                }
            }
        }
    }
}

// Texture2D::SampleImplicitLod<SurfaceSampler>(u, v, layer 0) over one 4x4 fragment (Texture.h:403-410):
// LOD from 2x2 finite differences inside the fragment (dFdx/dFdy, :260-269), CalcMipLevel(grad) - LerpFracBits
// (:271-275, :408), fragment-wide filter vote in SampleLevel (:432). Used by the alpha-tested fragment program.
void orc_sample_implicit_lod_4x4_layer(const swr_texture_desc* tex, const float* u, const float* v, uint32_t layer, uint32_t* out);
void orc_sample_implicit_lod_4x4(const swr_texture_desc* tex, const float* u, const float* v, uint32_t* out) {
    orc_sample_implicit_lod_4x4_layer(tex, u, v, 0, out);
}
void orc_sample_implicit_lod_4x4_layer(const swr_texture_desc* tex, const float* u, const float* v, uint32_t layer, uint32_t* out) {
    const float scaleLerpU = (float)(tex->Width << 8), scaleLerpV = (float)(tex->Height << 8);
    float su[N], sv[N];
    for (int i = 0; i < N; i++) { su[i] = u[i] * scaleLerpU; sv[i] = v[i] * scaleLerpV; }
    int32_t mip[N];
    bool anyMin = false;
    for (int i = 0; i < N; i++) {
        int row = i >> 2, x = i & 3;
        int ax = row * 4 + (x & ~1), bx = row * 4 + (x | 1);          // dFdx: [1 1 3 3] - [0 0 2 2] within a row
        int ay = (row & ~1) * 4 + x, by = (row | 1) * 4 + x;          // dFdy: rows [1 1 3 3] - rows [0 0 2 2]
        float g[4] = { su[bx] - su[ax], sv[bx] - sv[ax], su[by] - su[ay], sv[by] - sv[ay] };
        float dx = std::fmaf(g[0], g[0], g[1] * g[1]);
        float dy = std::fmaf(g[2], g[2], g[3] * g[3]);
        mip[i] = (ilog2(std::fmax(dx, dy)) >> 1) - 8;
        anyMin = anyMin || mip[i] > 0;
    }
    for (int i = 0; i < N; i++) out[i] = sample_level(*tex, u[i], v[i], layer, mip[i], anyMin);
}

// The shading half of FS_EncodeGBuffer (Shading.cpp:358-402) for one 4x4 fragment of one triangle: base colour = layer 0
// through SampleImplicitLod; with a second texture layer, the tangent-space normal (Nx Ny from the texture, Z
// reconstructed) rotated into world space by the interpolated vertex normal / tangent frame, octahedron-mapped and packed
// 10 + 10 bits beside the top 6 bits of metallic and roughness. `bary` = FragmentVars::Bary per lane.
void orc_gbuffer_fragment(const swr_texture_desc* tex, const float* u, const float* v, const float* bary /*[16][3]*/,
                          const uint32_t packedNT[3], uint32_t handedness, const float* objectToWorld3,
                          uint32_t* baseColor, uint32_t* packedCh2) {
    orc_sample_implicit_lod_4x4_layer(tex, u, v, 0, baseColor);                             // :364
    for (int i = 0; i < N; i++) packedCh2[i] = 0;                                         // :368
    if (tex->NumLayers < 2) return;                                                       // :370
    uint32_t nmr[N];
    orc_sample_implicit_lod_4x4_layer(tex, u, v, 1, nmr);                                 // :372
    V3 n0, n1, n2, t0, t1, t2;
    unpack_normal_tangent(packedNT[0], n0, t0);                                           // :380-382
    unpack_normal_tangent(packedNT[1], n1, t1);
    unpack_normal_tangent(packedNT[2], n2, t2);
    for (int i = 0; i < N; i++) {
        const float* b = bary + i * 3;
        float nx = (float)(nmr[i] & 255) * (1.0f / 127.5f) - 1.0f;                        // :375
        float ny = (float)((nmr[i] >> 8) & 255) * (1.0f / 127.5f) - 1.0f;
        float nz = approx_sqrt(1.0f - (nx * nx + ny * ny));                               // :376
        V3 nl = { bary_lerp(b, n0.x, n1.x, n2.x), bary_lerp(b, n0.y, n1.y, n2.y), bary_lerp(b, n0.z, n1.z, n2.z) };
        V3 tl = { bary_lerp(b, t0.x, t1.x, t2.x), bary_lerp(b, t0.y, t1.y, t2.y), bary_lerp(b, t0.z, t1.z, t2.z) };
        V3 normalWS = normalize3(mul_mat3(objectToWorld3, nl));                           // :384
        V3 tangentWS = normalize3(mul_mat3(objectToWorld3, tl));                          // :385
        V3 bit = cross3(normalWS, tangentWS);                                             // :388
        bit = { u2f(f2u(bit.x) ^ handedness), u2f(f2u(bit.y) ^ handedness), u2f(f2u(bit.z) ^ handedness) };   // :389
        V3 nr = normalize3({ nx * tangentWS.x + ny * bit.x + nz * normalWS.x,             // :392-396
                             nx * tangentWS.y + ny * bit.y + nz * normalWS.y,
                             nx * tangentWS.z + ny * bit.z + nz * normalWS.z });
        float ou, ov;
        map_octahedron(nr, ou, ov);                                                       // :398
        uint32_t c = (uint32_t)(ou * 1023.0f + 0.5f);                                     // :399 (conv<uint32_t> truncates)
        c |= (uint32_t)(ov * 1023.0f + 0.5f) << 10;                                       // :400
        c |= ((nmr[i] >> 18) & 0x3F) << 20;                                               // :401
        c |= ((nmr[i] >> 26) & 0x3F) << 26;                                               // :402
        packedCh2[i] = c;
    }
}

// Texture2D::GenerateMip for one layer/level (Texture.h:577-596): 2x2 box filter in float, RNE pack.
void orc_generate_mip(uint32_t* data, const swr_texture_desc* t, uint32_t layer, uint32_t level) {
    uint32_t w = t->Width >> level, h = t->Height >> level;
    const uint32_t* src = data + layer * t->LayerStride + t->MipOffsets[level - 1];
    uint32_t* dst = data + layer * t->LayerStride + t->MipOffsets[level];
    uint32_t sstride = t->RowShift - level + 1, dstride = t->RowShift - level;
    const float s = 1.0f / 255;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint32_t c[4] = { src[texel_offset(2 * x, 2 * y, sstride)], src[texel_offset(2 * x + 1, 2 * y, sstride)],
                              src[texel_offset(2 * x, 2 * y + 1, sstride)], src[texel_offset(2 * x + 1, 2 * y + 1, sstride)] };
            float avg[4];
            for (int k = 0; k < 4; k++) {
                float a = (float)((c[0] >> (8 * k)) & 255) * s, b = (float)((c[1] >> (8 * k)) & 255) * s;
                float cc = (float)((c[2] >> (8 * k)) & 255) * s, d = (float)((c[3] >> (8 * k)) & 255) * s;
                avg[k] = (((a + b) + cc) + d) * 0.25f;
            }
            dst[texel_offset(x, y, dstride)] = pack_rgba8(avg[0], avg[1], avg[2], avg[3]);
        }
}

}  // extern "C"
