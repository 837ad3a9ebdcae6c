// oracle_hiz.cpp — CPU restatement of the HiZ occlusion-culling row.
//
// TEST INFRASTRUCTURE ONLY (see oracle.cpp header); PARITY PINNED: same pyramid texels and cull bitmaps as the
// reference's own ImageHelpers.cpp / Shading.cpp built by oracle/ref_build.py (tests/test_ref_pin.py).
//   texutil::DownsampleDepth      ImageHelpers.cpp:150-247   depth layer -> half-res R32f min pyramid (TiledY8)
//   ProjectSphere                 Shading.cpp:264-279
//   ShadingContext::CullMeshlets  Shading.cpp:775-869        frustum test + HiZ test against the pyramid
// Canonical arithmetic: approx_rcp -> 1/x, approx_sqrt(x) -> (1/sqrt(x)) * x, IEEE elsewhere, FMA only where
// the source writes simd::dot / simd::mul. conv<int> truncates (cvttps2dq).
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../include/swr_types.h"

namespace {

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t texel_offset(uint32_t x, uint32_t y, uint32_t stride) { return (y & 7u) | (x << 3) | ((y & ~7u) << stride); }   // Texture.h:494-501
inline uint32_t fb_pixel_offset(uint32_t x, uint32_t y, uint32_t width) { return ((x & ~3u) << 2) + (y & ~3u) * width + (x & 3) + (y & 3) * 4; }
inline int32_t conv_int(float x) {   // cvttps2dq
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)x;
}
inline int32_t ilog2(float x) { return ((int32_t)f2u(x) - (127 << 23)) >> 23; }   // SIMD.h:426

static void mat4_mul(const float* a, const float* b, float* r) {   // glm mat4 * mat4, column-major, left-to-right sums
    float t[16];
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 4; k++)
            t[c * 4 + k] = ((a[0 * 4 + k] * b[c * 4 + 0] + a[1 * 4 + k] * b[c * 4 + 1]) + a[2 * 4 + k] * b[c * 4 + 2]) + a[3 * 4 + k] * b[c * 4 + 3];
    memcpy(r, t, sizeof(t));
}

}  // namespace

extern "C" {

// texutil::DownsampleDepth (ImageHelpers.cpp:243-247). `pyramid` is the Texture2D<R32f, TiledY8> the
// Playground creates with CreateTexture2D(halfW, halfH, 16) (Main.cpp:54-56). Level m texel (X,Y) = min of the
// depth layer over its 2^(m+1) x 2^(m+1) pixel footprint, FLT_MAX outside the framebuffer. Like the reference's
// recursive ReduceNxN, only the 8x8-texel blocks whose origin lies inside the framebuffer are written (:233-240,
// :224); everything else keeps its previous contents (never read: lookups clamp to the frame, Shading.cpp:826-827).
void orc_downsample_depth(const float* depthLayer, uint32_t width, uint32_t height, const swr_texture_desc* pyramid, float* data) {
    uint32_t maxDim = width > height ? width : height;
    uint32_t rootLevel = 0;
    while ((1u << rootLevel) <= maxDim) rootLevel++;            // 32 - lzcnt(max(W,H))   (:244)
    for (uint32_t m = 0; m + 3 <= rootLevel; m++) {             // levels 0 .. rootLevel-3
        uint32_t texel = 1u << (m + 1);                         // footprint in pixels
        uint32_t stride = pyramid->RowShift - m;
        float* dst = data + pyramid->MipOffsets[m];
        // blocks of 8x8 texels whose origin pixel is inside the frame; the top level is one 4x4 tile (:246)
        bool top = (m + 3 == rootLevel);
        uint32_t blockTexels = top ? 4 : 8;
        for (uint32_t by = 0; by * blockTexels * texel < height && (!top || by == 0); by++)
            for (uint32_t bx = 0; bx * blockTexels * texel < width && (!top || bx == 0); bx++)
                for (uint32_t ty = 0; ty < blockTexels; ty++)
                    for (uint32_t tx = 0; tx < blockTexels; tx++) {
                        uint32_t X = bx * blockTexels + tx, Y = by * blockTexels + ty;
                        float v = FLT_MAX;
                        for (uint32_t py = Y * texel; py < (Y + 1) * texel && py < height; py++)
                            for (uint32_t px = X * texel; px < (X + 1) * texel && px < width; px++) {
                                float d = depthLayer[fb_pixel_offset(px, py, width)];
                                v = d < v ? d : v;
                            }
                        dst[texel_offset(X, Y, stride)] = v;
                    }
    }
}

// ShadingContext::CullMeshlets with the HiZ part (Shading.cpp:775-869). pyramid == NULL: frustum only.
uint32_t orc_cull_meshlets_hiz(uint16_t* bitmap, const swr_meshlet* meshlets, uint32_t count,
                               const float* proj, const float* view, const float* model, const float* prevView,
                               float frameW, float frameH, const swr_texture_desc* pyramid, const float* pyramidData) {
    float pv[16], pvm[16], objectToPrevView[16];
    mat4_mul(proj, view, pv);
    mat4_mul(pv, model, pvm);                                   // :781
    mat4_mul(prevView, model, objectToPrevView);                // :780
    float planes[6][4];
    for (int i = 0; i < 3; i++) {                               // :784-791
        float a[4], b[4];
        for (int c = 0; c < 4; c++) { a[c] = pvm[c * 4 + 3] + pvm[c * 4 + i]; b[c] = pvm[c * 4 + 3] - pvm[c * 4 + i]; }
        float la = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]), lb = std::sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
        for (int c = 0; c < 4; c++) { planes[i * 2][c] = a[c] / la; planes[i * 2 + 1][c] = b[c] / lb; }
    }
    const float scale = std::sqrt((model[0] * model[0] + model[1] * model[1]) + model[2] * model[2]);   // :794 length(vec3(modelMat[0]))
    const float pz = proj[3 * 4 + 2], px = proj[0], py = proj[1 * 4 + 1];                              // :795 (znear, f/ar, -f)
    uint32_t visibleCount = 0;
    for (uint32_t offset = 0; offset < count; offset += 16) {
        uint16_t bits = 0;
        for (uint32_t l = 0; l < 16 && offset + l < count; l++) {
            const swr_meshlet& m = meshlets[offset + l];
            const float cx = m.BoundCenter[0], cy = m.BoundCenter[1], cz = m.BoundCenter[2], r = m.BoundRadius;
            bool visible = true;
            for (int i = 0; i < 5; i++) {                       // :806-809
                float dist = std::fmaf(cx, planes[i][0], std::fmaf(cy, planes[i][1], cz * planes[i][2])) + planes[i][3];
                visible = visible && (dist > -r);
            }
            if (visible && pyramid != nullptr) {                // :811-838 (per lane)
                const float* M = objectToPrevView;
                float vx_ = std::fmaf(cx, M[0], std::fmaf(cy, M[4], std::fmaf(cz, M[8], 1.0f * M[12])));   // simd::mul(mat4, (c,1))
                float vy_ = std::fmaf(cx, M[1], std::fmaf(cy, M[5], std::fmaf(cz, M[9], 1.0f * M[13])));
                float vz_ = std::fmaf(cx, M[2], std::fmaf(cy, M[6], std::fmaf(cz, M[10], 1.0f * M[14])));
                float rv = r * scale;
                // ProjectSphere (Shading.cpp:264-279)
                float c_x = vx_, c_y = vy_, c_z = vz_ * -1.0f;
                bool projMask = c_z >= rv + pz;
                if (projMask) {
                    float crx = c_x * rv, cry = c_y * rv, crz = c_z * rv;
                    float tx = c_x * c_x + c_z * c_z - rv * rv, ty = c_y * c_y + c_z * c_z - rv * rv;
                    float vx = (1.0f / std::sqrt(tx)) * tx, vy = (1.0f / std::sqrt(ty)) * ty;             // approx_sqrt
                    float bbx = (vx * c_x - crz) * (1.0f / (vx * c_z + crx)) * (px * 0.5f) + 0.5f;
                    float bby = (vy * c_y + crz) * (1.0f / (vy * c_z - cry)) * (py * 0.5f) + 0.5f;
                    float bbz = (vx * c_x + crz) * (1.0f / (vx * c_z - crx)) * (px * 0.5f) + 0.5f;
                    float bbw = (vy * c_y - crz) * (1.0f / (vy * c_z + cry)) * (py * 0.5f) + 0.5f;
                    float sizeX = (bbz - bbx) * frameW, sizeY = (bbw - bby) * frameH;                      // :820
                    int32_t mip = ilog2(sizeX > sizeY ? sizeX : sizeY);                                    // :821 (simd::max = maxnum)
                    int32_t maxMip = (int32_t)pyramid->MipLevels - 1;
                    mip = mip < 1 ? 1 : (mip > maxMip ? maxMip : mip);                                     // :822
                    int32_t x0 = conv_int(bbx * frameW), y0 = conv_int(bby * frameH);
                    int32_t x1 = conv_int(bbz * frameW), y1 = conv_int(bbw * frameH);
                    x0 = (x0 > 0 ? x0 : 0) >> mip; y0 = (y0 > 0 ? y0 : 0) >> mip;                          // :824-825
                    int32_t fw = (int32_t)frameW - 1, fh = (int32_t)frameH - 1;
                    x1 = (x1 < fw ? x1 : fw) >> mip; y1 = (y1 < fh ? y1 : fh) >> mip;                      // :826-827
                    float depthSphere = pz / (-vz_ - rv);                                                  // :829
                    float depthVisible = FLT_MAX;
                    const float* lvl = pyramidData + pyramid->MipOffsets[mip - 1];
                    uint32_t stride = pyramid->RowShift - (uint32_t)(mip - 1);
                    for (int32_t y = y0; y <= y1; y++)                                                     // :832-838
                        for (int32_t x = x0; x <= x1; x++) {
                            float v = lvl[texel_offset((uint32_t)x, (uint32_t)y, stride)];
                            depthVisible = v < depthVisible ? v : depthVisible;
                        }
                    visible = depthSphere > depthVisible;                                                  // :839
                }
            }
            if (visible) { bits |= (uint16_t)(1u << l); visibleCount++; }
        }
        bitmap[offset / 16] = bits;
    }
    return visibleCount;
}

}  // extern "C"
