// resolve_common.h — the scalar building blocks of the resolve-pass restatement (texture sampling, pixel formats,
// BRDF helpers, octahedron mapping), shared by oracle_resolve.cpp (the scalar spec) and resolve_avx512.cpp (the
// 16-lane version the timed CPU baseline uses). TEST INFRASTRUCTURE ONLY (see oracle.cpp header). Every function cites
// the reference file:line it follows; paths are relative to /root/reference/.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/swr_types.h"

namespace orc_detail {

constexpr int N = 16;
constexpr float kInvPi = 0.3183098861837907f;   // SIMD.h:386

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline int32_t round2i(float x) {
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)std::nearbyintf(x);
}
inline float approx_rcp(float x) { return 1.0f / x; }
inline float approx_rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline float approx_sqrt(float x) { return approx_rsqrt(x) * x; }        // SIMD.h:296
inline float lerpf(float a, float b, float t) { return std::fmaf(t, b, std::fmaf(-t, a, a)); }   // SIMD.h:445
inline float clampf(float x, float a, float b) { return std::fmin(std::fmax(x, a), b); }
inline float mulsign(float x, float y) { return u2f(f2u(x) ^ (f2u(y) & 0x80000000u)); }          // SIMD.h:341-344
inline int32_t ilog2(float x) { return ((int32_t)f2u(x) - (127 << 23)) >> 23; }                  // SIMD.h:426

struct V3 { float x, y, z; };
inline float dot3(V3 a, V3 b) { return std::fmaf(a.x, b.x, std::fmaf(a.y, b.y, a.z * b.z)); }   // SIMD.h:437
inline V3 normalize3(V3 a) { float r = approx_rsqrt(dot3(a, a)); return { a.x * r, a.y * r, a.z * r }; }   // SIMD.h:443
inline V3 cross3(V3 a, V3 b) {                                                                    // SIMD.h:435-441
    return { std::fmaf(a.y, b.z, -a.z * b.y), std::fmaf(a.z, b.x, -a.x * b.z), std::fmaf(a.x, b.y, -a.y * b.x) };
}
inline V3 mul_mat3(const float* m, V3 n) {                                                        // SIMD.h:465-471
    return { std::fmaf(n.x, m[0], std::fmaf(n.y, m[3], n.z * m[6])),
             std::fmaf(n.x, m[1], std::fmaf(n.y, m[4], n.z * m[7])),
             std::fmaf(n.x, m[2], std::fmaf(n.y, m[5], n.z * m[8])) };
}
inline void mul_mat4(const float* m, float x, float y, float z, float w, float out[4]) {          // SIMD.h:457-464
    for (int r = 0; r < 4; r++)
        out[r] = std::fmaf(x, m[0 * 4 + r], std::fmaf(y, m[1 * 4 + r], std::fmaf(z, m[2 * 4 + r], w * m[3 * 4 + r])));
}
inline float bary_lerp(const float b[3], float v0, float v1, float v2) {                          // Rasterizer.h:101-104
    return std::fmaf(v0, b[0], std::fmaf(v1, b[1], v2 * b[2]));
}

inline float half2float(uint16_t h) {   // _mm512_cvtph_ps (exact)
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 31, man = h & 1023;
    if (exp == 0) {
        if (man == 0) return u2f(sign);
        float f = (float)man * (1.0f / 16777216.0f);   // man * 2^-24
        return u2f(f2u(f) | sign);
    }
    if (exp == 31) return u2f(sign | 0x7F800000u | (man << 13));
    return u2f(sign | ((exp + 112) << 23) | (man << 13));
}

// texutil::UnmapOctahedron — Texture.h:289-296
inline V3 unmap_octahedron(float u, float v) {
    u = u * 2.0f - 1.0f; v = v * 2.0f - 1.0f;
    V3 n = { u, v, 1.0f - std::fabs(u) - std::fabs(v) };
    float t = std::fmax(-n.z, 0.0f);
    n.x -= mulsign(t, n.x);
    n.y -= mulsign(t, n.y);
    return normalize3(n);
}
// UnpackNormalTangent — Shading.cpp:232-236 (RGBA8u::Unpack, Texture.h:28-36)
inline void unpack_normal_tangent(uint32_t p, V3& n, V3& t) {
    const float s = 1.0f / 255;
    float a = (float)(p & 255) * s, b = (float)((p >> 8) & 255) * s, c = (float)((p >> 16) & 255) * s, d = (float)((p >> 24) & 255) * s;
    n = unmap_octahedron(a, b);
    t = unmap_octahedron(c, d);
}

// ---- texture sampling -----------------------------------------------------------------------
inline uint32_t texel_offset(uint32_t x, uint32_t y, uint32_t stride) {        // Texture.h:494-501 (TiledY8)
    return (y & 7u) | (x << 3) | ((y & ~7u) << stride);
}
inline uint32_t lerp16(uint32_t a, uint32_t b, uint32_t t) {                    // SIMD.h:448-450 on both s16 halves
    uint32_t r = 0;
    for (int h = 0; h < 2; h++) {
        int16_t ah = (int16_t)(a >> (16 * h)), bh = (int16_t)(b >> (16 * h)), th = (int16_t)(t >> (16 * h));
        int16_t diff = (int16_t)(bh - ah);
        int16_t m = (int16_t)((((int32_t)diff * (int32_t)th) + (1 << 14)) >> 15);   // vpmulhrsw
        r |= (uint32_t)(uint16_t)(int16_t)(ah + m) << (16 * h);
    }
    return r;
}
// CalcMipLevel(grad, scale) — Texture.h:276-280
inline int32_t calc_mip_level(const float g[4], float scaleU, float scaleV) {
    float dx = std::fmaf(g[0], g[0], g[1] * g[1]) * (scaleU * scaleU);
    float dy = std::fmaf(g[2], g[2], g[3] * g[3]) * (scaleV * scaleV);
    return ilog2(std::fmax(dx, dy)) >> 1;
}
// Texture2D::SampleLevel<Repeat, mag Linear, min Nearest> for one lane — Texture.h:412-459, :506-575.
// `useNearest` is the tile-wide filter decision (`simd::any(mipLevel > 0)`, :432).
inline uint32_t sample_level(const swr_texture_desc& t, float u, float v, uint32_t layer, int32_t mipLevel, bool useNearest) {
    const int32_t maskLerpU = (int32_t)(t.Width << 8) - 1, maskLerpV = (int32_t)(t.Height << 8) - 1;   // :627-628
    const float scaleLerpU = (float)(maskLerpU + 1), scaleLerpV = (float)(maskLerpV + 1);
    float su = u * scaleLerpU, sv = v * scaleLerpV;
    int32_t ix = round2i(su) & maskLerpU, iy = round2i(sv) & maskLerpV;       // Repeat (:424-426)
    int32_t maxLevel = (int32_t)t.MipLevels - 1;
    mipLevel = mipLevel < 0 ? 0 : (mipLevel > maxLevel ? maxLevel : mipLevel);
    uint32_t offset = layer * t.LayerStride;
    uint32_t stride = t.RowShift;
    if (mipLevel > 0) {                                                       // :443-447 (no-op for level 0)
        ix >>= mipLevel; iy >>= mipLevel;
        stride -= (uint32_t)mipLevel;
        offset += t.MipOffsets[mipLevel];
    }
    if (useNearest) return t.Data[offset + texel_offset((uint32_t)(ix >> 8), (uint32_t)(iy >> 8), stride)];   // :450-451

    // SampleLinear (:506-575)
    int32_t ixf = ix - 127 > 0 ? ix - 127 : 0, iyf = iy - 127 > 0 ? iy - 127 : 0;
    int32_t tx = ixf >> 8, ty = iyf >> 8;
    bool inboundX = ((tx + 1) << mipLevel) < (int32_t)t.Width;
    bool inboundY = ((ty + 1) << mipLevel) < (int32_t)t.Height;
    uint32_t i00 = offset + texel_offset((uint32_t)tx, (uint32_t)ty, stride);
    uint32_t d00 = t.Data[i00], d10 = t.Data[i00 + 8];
    uint32_t i01 = offset + texel_offset((uint32_t)tx, (uint32_t)(ty + (inboundY ? 1 : 0)), stride);
    uint32_t d01 = t.Data[i01], d11 = t.Data[i01 + 8];
    uint32_t fx = (uint32_t)(ixf & 255) << 7, fy = (uint32_t)(iyf & 255) << 7;
    fx = (fx << 16) | fx; fy = (fy << 16) | fy;
    if (!inboundX) fx = 0;
    uint32_t rbRow1 = lerp16(d00 & 0x00FF00FFu, d10 & 0x00FF00FFu, fx);
    uint32_t gaRow1 = lerp16((d00 >> 8) & 0x00FF00FFu, (d10 >> 8) & 0x00FF00FFu, fx);
    uint32_t rbRow2 = lerp16(d01 & 0x00FF00FFu, d11 & 0x00FF00FFu, fx);
    uint32_t gaRow2 = lerp16((d01 >> 8) & 0x00FF00FFu, (d11 >> 8) & 0x00FF00FFu, fx);
    uint32_t rbCol = lerp16(rbRow1, rbRow2, fy);
    uint32_t gaCol = lerp16(gaRow1, gaRow2, fy);
    return rbCol | (gaCol << 8);
}

// pixfmt::R11G11B10f::Unpack — Texture.h:138-144, :170-182 (5-bit exponent, 6 / 5-bit mantissa, no denormals)
inline void unpack_r11g11b10f(uint32_t p, float out[3]) {
    out[0] = u2f((((p >> 21) << 17) & 0x0FFE0000u) + 0x38000000u);
    out[1] = u2f((((p >> 10) << 17) & 0x0FFE0000u) + 0x38000000u);
    out[2] = u2f(((p << 18) & 0x0FFC0000u) + 0x38000000u);
}

// texutil::MapOctahedron — Texture.h:282-288
inline void map_octahedron(V3 n, float& u, float& v) {
    float w = approx_rcp(std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z));
    float t = std::fmax(-n.z * w, 0.0f);
    u = std::fmaf(n.x, w, mulsign(t, n.x)) * 0.5f + 0.5f;
    v = std::fmaf(n.y, w, mulsign(t, n.y)) * 0.5f + 0.5f;
}

// HdrTexture2D::SampleOctLevel<EnvSampler>(dir, 1) — Texture.h:467-480 with SampleLevel<ClampToEdge, Linear, Linear>
// (:412-459) and the float branch of SampleLinear (:557-573). mipLevel = 1.0 exactly: baseMip = 1, mipFrac = 0, so one
// bilinear sample of level 1 (`any(mipLevel > 0)` selects MinFilter = Linear; the level is clamped to the chain).
inline void sample_skybox(const swr_texture_desc& t, V3 dir, float out[3]) {
    float u, v;
    map_octahedron(dir, u, v);
    const int32_t maskLerpU = (int32_t)(t.Width << 8) - 1, maskLerpV = (int32_t)(t.Height << 8) - 1;
    int32_t ix = round2i(u * (float)(maskLerpU + 1)), iy = round2i(v * (float)(maskLerpV + 1));
    ix = ix < 0 ? 0 : (ix > maskLerpU ? maskLerpU : ix);                      // ClampToEdge (:421-423)
    iy = iy < 0 ? 0 : (iy > maskLerpV ? maskLerpV : iy);
    int32_t maxLevel = (int32_t)t.MipLevels - 1, mipLevel = 1 > maxLevel ? maxLevel : 1;
    uint32_t offset = 0, stride = t.RowShift;
    if (mipLevel > 0) { ix >>= mipLevel; iy >>= mipLevel; stride -= (uint32_t)mipLevel; offset += t.MipOffsets[mipLevel]; }
    int32_t ixf = ix - 127 > 0 ? ix - 127 : 0, iyf = iy - 127 > 0 ? iy - 127 : 0;
    int32_t tx = ixf >> 8, ty = iyf >> 8;
    bool inboundX = ((tx + 1) << mipLevel) < (int32_t)t.Width, inboundY = ((ty + 1) << mipLevel) < (int32_t)t.Height;
    uint32_t i00 = offset + texel_offset((uint32_t)tx, (uint32_t)ty, stride);
    uint32_t i01 = offset + texel_offset((uint32_t)tx, (uint32_t)(ty + (inboundY ? 1 : 0)), stride);
    float c00[3], c10[3], c01[3], c11[3];
    unpack_r11g11b10f(t.Data[i00], c00); unpack_r11g11b10f(t.Data[i00 + 8], c10);
    unpack_r11g11b10f(t.Data[i01], c01); unpack_r11g11b10f(t.Data[i01 + 8], c11);
    const float fracScale = 1.0f / 256;
    float fx = inboundX ? (float)(ixf & 255) * fracScale : 0.0f, fy = (float)(iyf & 255) * fracScale;
    for (int k = 0; k < 3; k++) {
        float rowA = c00[k] + (c10[k] - c00[k]) * fx, rowB = c01[k] + (c11[k] - c01[k]) * fx;
        out[k] = rowA + (rowB - rowA) * fy;
    }
}

// RGBA8u::UnpackSrgb — Texture.h:37-54
inline void unpack_srgb(uint32_t packed, float out[4]) {
    uint32_t rb1 = ((packed << 8) & 0xFF00FF00u) + 0x00FF00FFu;
    uint32_t ag1 = (packed & 0xFF00FF00u) + 0x00FF00FFu;
    auto mulhi16 = [](uint32_t a) {
        uint32_t lo = a & 0xFFFF, hi = a >> 16;
        return ((lo * lo) >> 16) | (((hi * hi) >> 16) << 16);
    };
    uint32_t rb2 = mulhi16(rb1), ag2 = mulhi16(ag1);
    const float scale = 1.0f / 65535;
    out[0] = (float)(rb2 & 65535) * scale;
    out[1] = (float)(ag2 & 65535) * scale;
    out[2] = (float)(rb2 >> 16) * scale;
    out[3] = (float)(ag1 >> 16) * scale;
}
// RGBA8u::Pack — Texture.h:55-67 (vcvtps2dq RNE, packssdw, packuswb saturation)
inline uint32_t pack_rgba8(float r, float g, float b, float a) {
    auto ch = [](float v) -> uint32_t {
        int32_t i = round2i(v * 255.0f);
        i = i < -32768 ? -32768 : (i > 32767 ? 32767 : i);
        i = i < 0 ? 0 : (i > 255 ? 255 : i);
        return (uint32_t)i;
    };
    return ch(r) | (ch(g) << 8) | (ch(b) << 16) | (ch(a) << 24);
}

// BRDF helpers — Shading.cpp:17-33
inline float pow5(float x) { return (x * x) * (x * x) * x; }
inline float D_GGX(float NoH, float roughness) {
    float a = NoH * roughness;
    float k = roughness * approx_rcp(1.0f - NoH * NoH + a * a);
    return k * k * kInvPi;
}
inline float V_SmithGGXCorrelatedFast(float NoV, float NoL, float roughness) {
    float a = 2.0f * NoL * NoV;
    float b = NoL + NoV;
    return 0.5f / lerpf(a, b, roughness);
}
inline float F_Schlick1(float u, float f0) {
    float f = pow5(1.0f - u);
    return f + f0 * (1.0f - f);
}

// GetLightAttenuation — Shading.cpp:581-600
inline float light_attenuation(const swr_light& light, V3 p) {
    if (light.Type == 0) return 1.0f;
    V3 posToLight = { light.Position[0] - p.x, light.Position[1] - p.y, light.Position[2] - p.z };
    float distanceSquare = dot3(posToLight, posToLight);
    float factor = distanceSquare * light.InvRadiusSq;
    float smoothFactor = std::fmax(1.0f - factor * factor, 0.0f);
    float attenuation = (smoothFactor * smoothFactor) * approx_rcp(std::fmax(distanceSquare, 1e-4f));
    if (light.Type == 2) {
        V3 nl = normalize3(posToLight);
        V3 nd = { -light.Direction[0], -light.Direction[1], -light.Direction[2] };
        float cd = dot3(nd, nl);
        float spot = clampf(cd * light.SpotScale + light.SpotOffset, 0.0f, 1.0f);
        attenuation *= spot * spot;
    }
    return attenuation;
}

struct Surface { uint32_t albedo; V3 normal; float metallic, roughness; };

// ColormapTurbo (Shading.cpp:249-260): three degree-5 polynomials in Horner form
inline void colormap_turbo(float x, float out[3]) {
    static const float c[] = {
        0.13572138f, 4.61539260f,  -42.66032258f, 132.13108234f, -152.94239396f, 59.28637943f,
        0.09140261f, 2.19418839f,  4.84296658f,   -14.18503333f, 4.27729857f,    2.82956604f,
        0.10667330f, 12.64194608f, -60.58204836f, 110.36276771f, -89.90310912f,  27.34824973f,
    };
    for (int k = 0; k < 3; k++) {
        const float* p = c + 6 * k;
        out[k] = x * (x * (x * (x * (x * p[5] + p[4]) + p[3]) + p[2]) + p[1]) + p[0];
    }
}
inline void unpack_rgba8(uint32_t p, float out[3]) {                      // RGBA8u::Unpack, Texture.h:28-36
    const float scale = 1.0f / 255;
    out[0] = (float)(p & 255) * scale; out[1] = (float)((p >> 8) & 255) * scale; out[2] = (float)((p >> 16) & 255) * scale;
}

// enum class DebugLayer (Shading.h:8)
enum { kLayerNone = 0, kLayerBaseColor, kLayerNormals, kLayerMetallicRoughness, kLayerMeshletId, kLayerTriangleId,
       kLayerOverdrawPixel, kLayerOverdrawQuad };

}  // namespace orc_detail
