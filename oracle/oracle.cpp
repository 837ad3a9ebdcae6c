// oracle.cpp — CPU restatement of GLimpSW's meshlet raster path (vis-buffer half).
//
// TEST INFRASTRUCTURE ONLY. Nothing under glimpsw_b200/ may import, link or execute
// this file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs can check and time the reference algorithm on the host.
//
// PARITY PINNED against output of the reference itself: the reference ships no tests, golden
// images or known-answer vectors for this path, but its own Rasterizer.cpp /
// Shading.cpp / ImageHelpers.cpp compile with g++ from where they lie under /root/reference
// (oracle/ref_build.py -> oracle/_ref/libswr_ref.so; SIMD.h and GLM replaced by stand-ins, 26
// one-line syntax edits), and tests/test_ref_pin.py checks that this restatement and that library
// produce the same words — vis-buffer, counters, cull bitmaps, resolved colour — on every scene it runs.
// This file is a line-by-line restatement of the reference SOURCE, each function
// citing the file:line it follows (paths relative to /root/reference/), in the canonical
// arithmetic of SURVEY.md Appendix A: IEEE-754 binary32 round-to-nearest-even, an FMA
// exactly where the source writes simd::fma / simd::mul / simd::dot, separately rounded
// * + - elsewhere, correctly rounded '/', float->int by RNE (vcvtps2dq), wrapping int32.
// (The upstream build uses -ffast-math, under which clang may turn 1.0f/x into
// vrcp14ps+Newton; that is not reproducible off x86 — SURVEY.md App. B.1. orc_set_reciprocal_mode(1)
// emulates it on AVX-512 hosts, for the sensitivity study of tools/rcp14_sensitivity.py only.)
//
// Build: g++ -O2 -march=x86-64-v3 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
// Scalar on purpose: this file is the spec. oracle/baseline_mt.cpp is the threaded
// restatement used as the timed CPU baseline, and must equal this file bit for bit.

#include <cmath>
#include <immintrin.h>
#include <cstdint>
#include <cstring>
#include <cfenv>

#include "../include/swr_types.h"

namespace {

// ---- SIMD.h op semantics -------------------------------------------------------------
inline int32_t round2i(float x) {  // SIMD.h:304 (_mm512_cvtps_epi32: RNE, 0x80000000 on overflow/NaN)
    if (!(x > -2147483904.0f && x < 2147483648.0f)) return INT32_MIN;
    return (int32_t)std::nearbyintf(x);  // default rounding mode = nearest-even
}
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct Vec4 { float x, y, z, w; };

// simd::mul(mat4, v4) with v.w == 1 — SIMD.h:457-464. m is column-major (glm): m[c*4+r].
inline Vec4 mul_mat4_pos(const float* m, float x, float y, float z) {
    Vec4 r;
    r.x = std::fmaf(x, m[0 * 4 + 0], std::fmaf(y, m[1 * 4 + 0], std::fmaf(z, m[2 * 4 + 0], 1.0f * m[3 * 4 + 0])));
    r.y = std::fmaf(x, m[0 * 4 + 1], std::fmaf(y, m[1 * 4 + 1], std::fmaf(z, m[2 * 4 + 1], 1.0f * m[3 * 4 + 1])));
    r.z = std::fmaf(x, m[0 * 4 + 2], std::fmaf(y, m[1 * 4 + 2], std::fmaf(z, m[2 * 4 + 2], 1.0f * m[3 * 4 + 2])));
    r.w = std::fmaf(x, m[0 * 4 + 3], std::fmaf(y, m[1 * 4 + 3], std::fmaf(z, m[2 * 4 + 3], 1.0f * m[3 * 4 + 3])));
    return r;
}
// ---- sensitivity switch: what the upstream BINARY most likely computes for `1.0f / x` ----------------
// The reference is built with -ffast-math (src/SwRast/CMakeLists.txt:6); for 512-bit vector divisions clang then emits
// vrcp14ps plus one Newton-Raphson step instead of vdivps. Mode 1 emulates that (rcp14 estimate e, e + e * (1 - x * e)
// with FMAs, numerator multiplied afterwards) so tools/rcp14_sensitivity.py can measure how far such a binary would be
// from the canonical IEEE arithmetic (mode 0, the default, the only mode parity is defined against).
int g_rcpMode = 0;
int g_contractDet = 0;       // sensitivity only: the cull determinant as fma(a, b, -(c * d)), the contraction -ffast-math allows (App. B.1b)
__attribute__((target("avx512f"))) inline float rcp14_nr(float x) {
    float e = _mm_cvtss_f32(_mm_rcp14_ss(_mm_setzero_ps(), _mm_set_ss(x)));
    return std::fmaf(std::fmaf(-x, e, 1.0f), e, e);
}
inline float oracle_rcp(float x) { return g_rcpMode ? rcp14_nr(x) : 1.0f / x; }

// simd::perspective_div — SIMD.h:473-476
inline Vec4 perspective_div(Vec4 v) {
    float rw = oracle_rcp(v.w);
    return { v.x * rw, v.y * rw, v.z * rw, rw };
}

inline int16_t lo16(uint32_t p) { return (int16_t)(p & 0xFFFF); }
inline int16_t hi16(uint32_t p) { return (int16_t)(p >> 16); }
inline uint32_t pack16(int16_t lo, int16_t hi) { return (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16); }
inline int16_t min16(int16_t a, int16_t b) { return a < b ? a : b; }
inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }
// per-s16 helpers on packed pairs (vpminsw / vpmaxsw / vpsraw / vpaddw)
inline uint32_t pmin16(uint32_t a, uint32_t b) { return pack16(min16(lo16(a), lo16(b)), min16(hi16(a), hi16(b))); }
inline uint32_t pmax16(uint32_t a, uint32_t b) { return pack16(max16(lo16(a), lo16(b)), max16(hi16(a), hi16(b))); }
inline uint32_t psra16(uint32_t a, int s) { return pack16((int16_t)(lo16(a) >> s), (int16_t)(hi16(a) >> s)); }
inline uint32_t padd16(uint32_t a, uint32_t b) { return pack16((int16_t)(lo16(a) + lo16(b)), (int16_t)(hi16(a) + hi16(b))); }

// ---- Triangle setup ------------------------------------------------------------------
struct TriSetup {           // one lane of TrianglePacket (Rasterizer.h:145-161)
    uint32_t Pos0, Pos1, Pos2;
    float Z0, Z1, Z2, W0, W1, W2;
};
struct TriEdges {           // one lane of TriangleEdgeVars (Rasterizer.h:162-176)
    int32_t Edge0, Edge1, Edge2, A01, A12, A20, B01, B12, B20;
    float Z0, Z10, Z20, W0, W0S, W1S, W2S;
};

// TrianglePacket::GetBoundingBox + GetRenderBoundingBox — Rasterizer.cpp:331-351
inline void render_bbox(const TriSetup& t, int halfW, int halfH, uint32_t& bbMin, uint32_t& bbMax) {
    uint32_t minPos = pmin16(pmin16(t.Pos0, t.Pos1), t.Pos2);
    uint32_t maxPos = pmax16(pmax16(t.Pos0, t.Pos1), t.Pos2);
    // :335-336 — a 32-bit add on packed s16 pairs (the carry from x into y is intended to be replicated)
    minPos = psra16(minPos + 0x00070007u, 4);
    maxPos = psra16(maxPos + 0x00070007u, 4);
    uint32_t vpSize = (uint32_t)halfW | ((uint32_t)halfH << 16);            // :343
    minPos = pmin16(pmax16(padd16(minPos, vpSize), 0), vpSize * 2);          // :344
    maxPos = pmin16(pmax16(padd16(maxPos, vpSize), 0), vpSize * 2);          // :345
    bbMin = (minPos + 0x00000000u) & ~0x00030003u;                           // :347
    bbMax = (maxPos + 0x00030003u) & ~0x00030003u;                           // :348
}

// TrianglePacket::Setup for one lane — Rasterizer.cpp:257-289. Returns 1 if the lane survives.
inline int tri_setup(Vec4 v0, Vec4 v1, Vec4 v2, int halfW, int halfH, int cullMode, TriSetup& out) {
    v0 = perspective_div(v0);
    v1 = perspective_div(v1);
    v2 = perspective_div(v2);

    // :262 (canonical: every product rounded, then the subtraction)
    float det = (v2.x - v0.x) * (v1.y - v0.y) - (v0.x - v1.x) * (v0.y - v2.y);
    if (g_contractDet) det = std::fmaf(v2.x - v0.x, v1.y - v0.y, -((v0.x - v1.x) * (v0.y - v2.y)));
    if (cullMode != SWR_CULL_FRONT_CCW) {                                    // :264-267
        bool flip = (cullMode == SWR_CULL_FRONT_CW) ? true : (det < 0);
        det = flip ? -det : det;
    }
    int keep = det > 0;                                                      // :269

    float fixX = (float)(halfW * 16), fixY = (float)(halfH * 16);            // :272
    int32_t x0 = round2i(v0.x * fixX), y0 = round2i(v0.y * fixY);
    int32_t x1 = round2i(v1.x * fixX), y1 = round2i(v1.y * fixY);
    int32_t x2 = round2i(v2.x * fixX), y2 = round2i(v2.y * fixY);
    out.Pos0 = ((uint32_t)x0 & 0xFFFF) | ((uint32_t)y0 << 16);               // :277-279
    out.Pos1 = ((uint32_t)x1 & 0xFFFF) | ((uint32_t)y1 << 16);
    out.Pos2 = ((uint32_t)x2 & 0xFFFF) | ((uint32_t)y2 << 16);

    uint32_t bbMin, bbMax;
    render_bbox(out, halfW, halfH, bbMin, bbMax);
    // :283 — drop when minX >= maxX or minY >= maxY (signed 16-bit compares)
    if (lo16(bbMin) >= lo16(bbMax) || hi16(bbMin) >= hi16(bbMax)) keep = 0;

    out.Z0 = v0.z; out.Z1 = v1.z; out.Z2 = v2.z;                             // :285-286
    out.W0 = v0.w; out.W1 = v1.w; out.W2 = v2.w;
    return keep;
}

// ComputeEdge — Rasterizer.cpp:291-295 (wrapping 32-bit arithmetic, arithmetic shift)
inline int32_t compute_edge(int32_t a, int32_t x, int32_t b, int32_t y) {
    uint32_t w = (uint32_t)a * (uint32_t)x + (uint32_t)b * (uint32_t)y;
    w += (a > 0 || (a == 0 && b > 0)) ? 0u : 0xFFFFFFFFu;
    return (int32_t)w >> 4;
}
// TriangleEdgeVars::Setup for one lane — Rasterizer.cpp:296-329
inline void edge_setup(const TriSetup& t, int halfW, int halfH, TriEdges& e) {
    int32_t x0 = lo16(t.Pos0), y0 = hi16(t.Pos0);
    int32_t x1 = lo16(t.Pos1), y1 = hi16(t.Pos1);
    int32_t x2 = lo16(t.Pos2), y2 = hi16(t.Pos2);
    e.A01 = y1 - y0; e.B01 = x0 - x1;
    e.A12 = y2 - y1; e.B12 = x1 - x2;
    e.A20 = y0 - y2; e.B20 = x2 - x0;
    int32_t det = (int32_t)((uint32_t)e.B20 * (uint32_t)e.A01 - (uint32_t)e.B01 * (uint32_t)e.A20);
    if (det < 0) {
        e.A01 = -e.A01; e.B01 = -e.B01; e.A12 = -e.A12; e.B12 = -e.B12; e.A20 = -e.A20; e.B20 = -e.B20;
        det = (int32_t)(0u - (uint32_t)det);
    }
    int32_t sampleX = (int32_t)((uint32_t)(-halfW) << 4) + 8, sampleY = (int32_t)((uint32_t)(-halfH) << 4) + 8;  // :315
    e.Edge0 = compute_edge(e.A12, sampleX - x1, e.B12, sampleY - y1);
    e.Edge1 = compute_edge(e.A20, sampleX - x2, e.B20, sampleY - y2);
    e.Edge2 = compute_edge(e.A01, sampleX - x0, e.B01, sampleY - y0);

    float rcpArea = g_rcpMode ? 16.0f * rcp14_nr((float)det) : 16.0f / (float)det;   // :320
    e.Z0 = t.Z0;
    e.Z10 = (t.Z1 - t.Z0) * rcpArea;
    e.Z20 = (t.Z2 - t.Z0) * rcpArea;
    e.W0 = t.W0;
    e.W0S = t.W0 * rcpArea; e.W1S = t.W1 * rcpArea; e.W2S = t.W2 * rcpArea;
}

// Clipper::ComputeClipCodes for one lane — Rasterizer.cpp:353-397.
// returns bit0 = accept, bit1 = non-trivial; *outcodes = partial outcodes
inline int clip_codes(const Vec4 v[3], float bx, float by, uint8_t* outcodes) {
    uint8_t partial = 0, combined = 255;
    bool trivial = true;
    for (int i = 0; i < 3; i++) {
        const Vec4& vi = v[i];
        uint8_t oc = 0;
        if (vi.x < -vi.w) oc |= 1 << 0;   // Left
        if (vi.x > +vi.w) oc |= 1 << 1;   // Right
        if (vi.y < -vi.w) oc |= 1 << 2;   // Top
        if (vi.y > +vi.w) oc |= 1 << 3;   // Bottom
        if (vi.z < -vi.w) oc |= 1 << 4;   // Near
        if (vi.z > +vi.w) oc |= 1 << 5;   // Far
        partial |= oc;
        combined &= oc;
        trivial = trivial && (std::fabs(vi.x) < vi.w * bx && std::fabs(vi.y) < vi.w * by);   // :386
    }
    trivial = trivial && ((partial & ((1 << 4) | (1 << 5))) == 0);                           // :388
    bool visible = combined == 0;                                                             // :389
    if (outcodes) *outcodes = partial;
    return (visible && trivial ? 1 : 0) | (visible && !trivial ? 2 : 0);
}

// Clipper::ClipTriangles for one lane — Rasterizer.cpp:398-478 (Sutherland-Hodgman against the planes named
// by the triangle's partial outcodes, ascending plane id; x/y planes sit on the guard band, z planes on the
// frustum; the result polygon is fan-triangulated). attr[4]/attr[5] carry the barycentric weights of v1/v2.
// Returns the number of output triangles (<= 7); tris[n][3] in clip space, remapU/V[n] = ClippedU/ClippedV.
struct ClipVert { float a[6]; };
inline float clip_dist(const ClipVert& v, uint32_t planeId, float scale) {   // GetIntersectDist, :76-81
    float a = v.a[planeId / 2];
    if (planeId % 2) a = -a;
    return a + v.a[3] * scale;
}
inline int clip_triangle(const Vec4 v[3], uint8_t outcodes, float bx, float by,
                         Vec4 tris[7][3], float remapU[7][3], float remapV[7][3]) {
    ClipVert verts[64];
    uint8_t indices[32];
    for (int i = 0; i < 3; i++) {
        verts[i] = { { v[i].x, v[i].y, v[i].z, v[i].w, i == 1 ? 1.0f : 0.0f, i == 2 ? 1.0f : 0.0f } };
        indices[i] = (uint8_t)i;
    }
    uint32_t vertCount = 3, nextIdx = 3;
    for (uint32_t planeId = 0; planeId < 6; planeId++) {                     // BitIter(OutCodes) — ascending bits
        if (!((outcodes >> planeId) & 1)) continue;
        uint8_t outIndices[32];
        uint32_t outCount = 0;
        float planeScale = planeId < 4 ? (planeId / 2 == 0 ? bx : by) : 1.0f;      // :422
        for (uint32_t vi = 0; vi < vertCount; vi++) {
            uint8_t ia = indices[vi], ib = indices[(vi + 1) % vertCount];
            float da = clip_dist(verts[ia], planeId, planeScale);
            float db = clip_dist(verts[ib], planeId, planeScale);
            if (da >= 0) outIndices[outCount++] = ia;
            if ((da >= 0) != (db >= 0)) {                                    // :432-439
                float t = da / (da - db);
                for (int k = 0; k < 6; k++) verts[nextIdx].a[k] = std::fmaf(verts[ib].a[k], t, std::fmaf(-t, verts[ia].a[k], verts[ia].a[k]));
                outIndices[outCount++] = (uint8_t)nextIdx++;
            }
        }
        if (vertCount < 3) break;                                            // :443 (tests the count before this plane)
        vertCount = outCount;
        memcpy(indices, outIndices, sizeof(indices));
    }
    if (vertCount < 3) return 0;                                             // :448 + the empty fan loop for 2 vertices
    int n = 0;
    for (uint32_t vi = 0; vi < vertCount - 2; vi++, n++) {                   // :451-466
        const ClipVert &p0 = verts[indices[0]], &p1 = verts[indices[vi + 1]], &p2 = verts[indices[vi + 2]];
        tris[n][0] = { p0.a[0], p0.a[1], p0.a[2], p0.a[3] };
        tris[n][1] = { p1.a[0], p1.a[1], p1.a[2], p1.a[3] };
        tris[n][2] = { p2.a[0], p2.a[1], p2.a[2], p2.a[3] };
        remapU[n][0] = p0.a[4]; remapU[n][1] = p1.a[4] - p0.a[4]; remapU[n][2] = p2.a[4] - p0.a[4];
        remapV[n][0] = p0.a[5]; remapV[n][1] = p1.a[5] - p0.a[5]; remapV[n][2] = p2.a[5] - p0.a[5];
    }
    return n;
}

inline uint32_t fb_pixel_offset(uint32_t x, uint32_t y, uint32_t width) {   // Rasterizer.h:50-56
    return ((x & ~3u) << 2) + (y & ~3u) * width + (x & 3) + (y & 3) * 4;
}

// Rasterizer::DrawTriangle<FS_EncodeSurfaceId<false>, false> — Rasterizer.h:250-328 + Shading.cpp:309-331,
// written per pixel: the reference walks 4x4 fragments over [bbMin, bbMax) and tests sign(e0|e1|e2).
inline void draw_triangle(uint32_t* color, float* depth, uint32_t width, const TriEdges& e,
                          uint32_t bbMin, uint32_t bbMax, uint32_t surfaceId) {
    uint32_t minX = bbMin & 0xFFFF, minY = bbMin >> 16, maxX = bbMax & 0xFFFF, maxY = bbMax >> 16;
    for (uint32_t y = minY; y < maxY; y++) {
        for (uint32_t x = minX; x < maxX; x++) {
            uint32_t e0 = (uint32_t)e.Edge0 + (uint32_t)e.A12 * x + (uint32_t)e.B12 * y;
            uint32_t e1 = (uint32_t)e.Edge1 + (uint32_t)e.A20 * x + (uint32_t)e.B20 * y;
            uint32_t e2 = (uint32_t)e.Edge2 + (uint32_t)e.A01 * x + (uint32_t)e.B01 * y;
            if ((int32_t)(e0 | e1 | e2) < 0) continue;                       // Rasterizer.h:289-290
            float u = (float)(int32_t)e1, v = (float)(int32_t)e2;            // :293-294
            float d = std::fmaf(u, e.Z10, std::fmaf(v, e.Z20, e.Z0));        // :296
            uint32_t off = fb_pixel_offset(x, y, width);
            if (d > depth[off]) {                                            // Shading.cpp:311-313
                depth[off] = d;                                              // :329
                color[off] = surfaceId;                                      // :328,:330
            }
        }
    }
}

// Rasterizer::DrawTriangle<FS_Overdraw, *> — Rasterizer.h:250-328 + Shading.cpp:333-342. The fragment program runs
// once per 4x4 fragment of the triangle's (4-aligned) bounding box that has at least one covered lane (:284), on
// all 16 lanes (it sets TileMask = 0xFFFF): covered lanes add 1 to the high u16 of the colour word, the others
// add 1 to the low u16, both saturating (_mm512_adds_epu16); the depth layer takes max(Depth, stored) on all 16
// lanes, Depth being the plane equation extrapolated to lanes outside the triangle. No depth test.
inline void draw_triangle_overdraw(uint32_t* color, float* depth, uint32_t width, const TriEdges& e, uint32_t bbMin, uint32_t bbMax) {
    uint32_t minX = bbMin & 0xFFFF, minY = bbMin >> 16, maxX = bbMax & 0xFFFF, maxY = bbMax >> 16;
    for (uint32_t fy = minY; fy < maxY; fy += 4) {
        for (uint32_t fx = minX; fx < maxX; fx += 4) {
            uint32_t e1v[16], e2v[16], mask = 0;
            for (uint32_t lane = 0; lane < 16; lane++) {
                uint32_t x = fx + (lane & 3), y = fy + (lane >> 2);                     // Rasterizer.h:247-248
                uint32_t e0 = (uint32_t)e.Edge0 + (uint32_t)e.A12 * x + (uint32_t)e.B12 * y;
                e1v[lane] = (uint32_t)e.Edge1 + (uint32_t)e.A20 * x + (uint32_t)e.B20 * y;
                e2v[lane] = (uint32_t)e.Edge2 + (uint32_t)e.A01 * x + (uint32_t)e.B01 * y;
                if ((int32_t)(e0 | e1v[lane] | e2v[lane]) >= 0) mask |= 1u << lane;
            }
            if (mask == 0) continue;                                                     // :284
            for (uint32_t lane = 0; lane < 16; lane++) {
                uint32_t off = fb_pixel_offset(fx + (lane & 3), fy + (lane >> 2), width);
                uint32_t c = color[off], hi = c >> 16, lo = c & 0xFFFF;
                if ((mask >> lane) & 1) hi = hi < 0xFFFF ? hi + 1 : 0xFFFF;              // Shading.cpp:336-337
                else lo = lo < 0xFFFF ? lo + 1 : 0xFFFF;
                color[off] = (hi << 16) | lo;
                float u = (float)(int32_t)e1v[lane], v = (float)(int32_t)e2v[lane];
                float d = std::fmaf(u, e.Z10, std::fmaf(v, e.Z20, e.Z0));                // Rasterizer.h:296
                depth[off] = std::fmax(d, depth[off]);                                   // Shading.cpp:341
            }
        }
    }
}

}  // namespace

// Texture2D::SampleImplicitLod<SurfaceSampler> over one 4x4 fragment (oracle_resolve.cpp).
extern "C" void orc_sample_implicit_lod_4x4(const swr_texture_desc* tex, const float* u, const float* v, uint32_t* out);

namespace {

inline float half_to_float(uint16_t h) {   // _mm_cvtph_ps (exact)
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 31, man = h & 1023;
    if (exp == 0) {
        if (man == 0) return u2f(sign);
        return u2f(f2u((float)man * (1.0f / 16777216.0f)) | sign);
    }
    if (exp == 31) return u2f(sign | 0x7F800000u | (man << 13));
    return u2f(sign | ((exp + 112) << 23) | (man << 13));
}

// Rasterizer::DrawTriangle<FS_EncodeSurfaceId<true>, false> — Rasterizer.h:250-328 + Shading.cpp:309-331.
// Per 4x4 fragment like the reference, because the texture LOD comes from finite differences over all 16
// lanes of the fragment (Texture.h:260-269, :403-410) and the filter choice is a fragment-wide vote (:432).
// Canonical arithmetic for the perspective correction: approx_rcp -> 1/w, then the source's Newton step.
inline void draw_triangle_alpha(uint32_t* color, float* depth, uint32_t width, const TriEdges& e,
                                uint32_t bbMin, uint32_t bbMax, uint32_t surfaceId, const uint32_t packedTC[3],
                                const swr_texture_desc* tex, uint32_t alphaCutoff,
                                const float* clipU = nullptr, const float* clipV = nullptr) {
    uint32_t minX = bbMin & 0xFFFF, minY = bbMin >> 16, maxX = bbMax & 0xFFFF, maxY = bbMax >> 16;
    float uv[3][2];
    for (int k = 0; k < 3; k++) {                                                     // UnpackHalf2x16 (Shading.cpp:227-230)
        uv[k][0] = half_to_float((uint16_t)(packedTC[k] & 0xFFFF));
        uv[k][1] = half_to_float((uint16_t)(packedTC[k] >> 16));
    }
    for (uint32_t y0 = minY; y0 < maxY; y0 += 4) {
        for (uint32_t x0 = minX; x0 < maxX; x0 += 4) {
            bool mask[16];
            bool any = false;
            float d[16], tu[16], tv[16];
            uint32_t off0 = ((x0 & ~3u) << 2) + (y0 & ~3u) * width;
            for (int i = 0; i < 16; i++) {
                uint32_t x = x0 + (i & 3), y = y0 + (i >> 2);
                uint32_t e0 = (uint32_t)e.Edge0 + (uint32_t)e.A12 * x + (uint32_t)e.B12 * y;
                uint32_t e1 = (uint32_t)e.Edge1 + (uint32_t)e.A20 * x + (uint32_t)e.B20 * y;
                uint32_t e2 = (uint32_t)e.Edge2 + (uint32_t)e.A01 * x + (uint32_t)e.B01 * y;
                mask[i] = (int32_t)(e0 | e1 | e2) >= 0;                                // Rasterizer.h:289-290
                float u = (float)(int32_t)e1, v = (float)(int32_t)e2;
                d[i] = std::fmaf(u, e.Z10, std::fmaf(v, e.Z20, e.Z0));                 // :296
                // perspective correction (:302-310, :319)
                float pw0 = std::fmaf(u + v, -e.W0S, e.W0);
                float w = std::fmaf(u, e.W1S, std::fmaf(v, e.W2S, pw0));
                float rcpW = 1.0f / w;
                rcpW *= std::fmaf(-w, rcpW, 2.0f);
                u *= e.W1S * rcpW;
                v *= e.W2S * rcpW;
                if (clipU != nullptr) {                                                // IsClipped, Rasterizer.h:312-318
                    float cu = std::fmaf(u, clipU[1], std::fmaf(v, clipU[2], clipU[0]));
                    float cv = std::fmaf(u, clipV[1], std::fmaf(v, clipV[2], clipV[0]));
                    u = cu, v = cv;
                }
                float b0 = 1 - u - v;
                // vars.Interpolate = BaryLerp (Rasterizer.h:101-104), v0,v1,v2 = VertexId[0..2]
                tu[i] = std::fmaf(uv[0][0], b0, std::fmaf(uv[1][0], u, uv[2][0] * v));
                tv[i] = std::fmaf(uv[0][1], b0, std::fmaf(uv[1][1], u, uv[2][1] * v));
                any = any || mask[i];
            }
            if (!any) continue;                                                        // Rasterizer.h:292
            any = false;
            for (int i = 0; i < 16; i++) {                                             // Shading.cpp:311-313
                mask[i] = mask[i] && (d[i] > depth[off0 + i]);
                any = any || mask[i];
            }
            if (!any) continue;
            uint32_t texel[16];
            orc_sample_implicit_lod_4x4(tex, tu, tv, texel);                           // :324
            for (int i = 0; i < 16; i++) {
                if (mask[i] && texel[i] >= (alphaCutoff << 24)) {                      // :326
                    depth[off0 + i] = d[i];
                    color[off0 + i] = surfaceId;
                }
            }
        }
    }
}

}  // namespace

extern "C" void orc_gbuffer_fragment(const swr_texture_desc* tex, const float* u, const float* v, const float* bary,
                                     const uint32_t packedNT[3], uint32_t handedness, const float* objectToWorld3,
                                     uint32_t* baseColor, uint32_t* packedCh2);

namespace {

// Rasterizer::DrawTriangle<FS_EncodeGBuffer, IsClipped> — Rasterizer.h:250-328 + Shading.cpp:344-414 (ShadingContext::
// DeferredShader, :655): depth test; material-less meshlets store depth only; otherwise base colour -> layer 0 (fragments
// whose texel is below AlphaCutoff << 24 are discarded — with the default cutoff of 255 that is every texel that is not fully
// opaque), packed normal / metallic / roughness -> layer 2, depth -> layer 1. `tex` == nullptr: material-less.
inline void draw_triangle_gbuffer(uint32_t* color, float* depth, uint32_t* ch2, uint32_t width, const TriEdges& e,
                                  uint32_t bbMin, uint32_t bbMax, const uint32_t packedTC[3], const uint32_t packedNT[3], uint32_t handedness,
                                  const swr_texture_desc* tex, uint32_t alphaCutoff, const float* objectToWorld3,
                                  const float* clipU = nullptr, const float* clipV = nullptr) {
    uint32_t minX = bbMin & 0xFFFF, minY = bbMin >> 16, maxX = bbMax & 0xFFFF, maxY = bbMax >> 16;
    float uv[3][2];
    for (int k = 0; k < 3; k++) {
        uv[k][0] = half_to_float((uint16_t)(packedTC[k] & 0xFFFF));
        uv[k][1] = half_to_float((uint16_t)(packedTC[k] >> 16));
    }
    for (uint32_t y0 = minY; y0 < maxY; y0 += 4) {
        for (uint32_t x0 = minX; x0 < maxX; x0 += 4) {
            bool mask[16];
            bool any = false;
            float d[16], tu[16], tv[16], bary[16][3];
            uint32_t off0 = ((x0 & ~3u) << 2) + (y0 & ~3u) * width;
            for (int i = 0; i < 16; i++) {
                uint32_t x = x0 + (i & 3), y = y0 + (i >> 2);
                uint32_t e0 = (uint32_t)e.Edge0 + (uint32_t)e.A12 * x + (uint32_t)e.B12 * y;
                uint32_t e1 = (uint32_t)e.Edge1 + (uint32_t)e.A20 * x + (uint32_t)e.B20 * y;
                uint32_t e2 = (uint32_t)e.Edge2 + (uint32_t)e.A01 * x + (uint32_t)e.B01 * y;
                mask[i] = (int32_t)(e0 | e1 | e2) >= 0;                                // Rasterizer.h:289-290
                float u = (float)(int32_t)e1, v = (float)(int32_t)e2;
                d[i] = std::fmaf(u, e.Z10, std::fmaf(v, e.Z20, e.Z0));                 // :296
                float pw0 = std::fmaf(u + v, -e.W0S, e.W0);                            // :302-310
                float w = std::fmaf(u, e.W1S, std::fmaf(v, e.W2S, pw0));
                float rcpW = 1.0f / w;
                rcpW *= std::fmaf(-w, rcpW, 2.0f);
                u *= e.W1S * rcpW;
                v *= e.W2S * rcpW;
                if (clipU != nullptr) {                                                // :312-318
                    float cu = std::fmaf(u, clipU[1], std::fmaf(v, clipU[2], clipU[0]));
                    float cv = std::fmaf(u, clipV[1], std::fmaf(v, clipV[2], clipV[0]));
                    u = cu, v = cv;
                }
                bary[i][0] = 1 - u - v; bary[i][1] = u; bary[i][2] = v;                // :319
                tu[i] = std::fmaf(uv[0][0], bary[i][0], std::fmaf(uv[1][0], u, uv[2][0] * v));
                tv[i] = std::fmaf(uv[0][1], bary[i][0], std::fmaf(uv[1][1], u, uv[2][1] * v));
                any = any || mask[i];
            }
            if (!any) continue;                                                        // Rasterizer.h:292
            any = false;
            for (int i = 0; i < 16; i++) {                                             // Shading.cpp:346-348
                mask[i] = mask[i] && (d[i] > depth[off0 + i]);
                any = any || mask[i];
            }
            if (!any) continue;
            if (tex == nullptr) {                                                      // :352-355
                for (int i = 0; i < 16; i++) if (mask[i]) depth[off0 + i] = d[i];
                continue;
            }
            uint32_t baseColor[16], packedCh2[16];
            orc_gbuffer_fragment(tex, tu, tv, &bary[0][0], packedNT, handedness, objectToWorld3, baseColor, packedCh2);
            for (int i = 0; i < 16; i++) {
                if (mask[i] && baseColor[i] >= (alphaCutoff << 24)) {                  // :365, :404-406
                    depth[off0 + i] = d[i];
                    color[off0 + i] = baseColor[i];
                    ch2[off0 + i] = packedCh2[i];
                }
            }
        }
    }
}

}  // namespace

extern "C" {

// ShadeMeshlet — Shading.cpp:281-307. cullBitmap may be NULL. `index` is the meshlet index
// within the draw; meshlets points at element MeshletOffset already.
// Returns PrimCount. Canonical rule for material-less meshlets: FrontCCW, FS 0.
int orc_shade_meshlet(const swr_meshlet* meshlets, uint32_t index, const uint16_t* cullBitmap,
                      const float* objectToClip, const swr_material* materials, swr_shaded_meshlet* out) {
    out->PrimCount = 0;
    out->CullMode = SWR_CULL_FRONT_CCW;
    out->FragmentShaderId = 0;
    if (cullBitmap != nullptr) {
        uint16_t mask = cullBitmap[index / 16];
        if (((mask >> (index % 16)) & 1) == 0) return 0;
    }
    const swr_meshlet& mesh = meshlets[index];
    uint32_t nv = (mesh.NumVertices + 15u) & ~15u;   // the reference walks 16-wide vectors (:295)
    if (nv > 64) nv = 64;
    for (uint32_t i = 0; i < nv; i++) {
        Vec4 c = mul_mat4_pos(objectToClip, mesh.Positions[0][i], mesh.Positions[1][i], mesh.Positions[2][i]);
        out->Position[0][i] = c.x; out->Position[1][i] = c.y; out->Position[2][i] = c.z; out->Position[3][i] = c.w;
    }
    out->PrimCount = mesh.NumTriangles;
    memcpy(out->Indices, mesh.Indices, sizeof(mesh.Indices));
    if (mesh.MaterialId != SWR_NO_MATERIAL && materials != nullptr) {
        const swr_material& mat = materials[mesh.MaterialId];
        out->CullMode = mat.IsDoubleSided ? SWR_CULL_NONE : SWR_CULL_FRONT_CCW;
        out->FragmentShaderId = mat.AlphaCutoff < 255 ? 1 : 0;
    }
    return out->PrimCount;
}

// Single-triangle probes for known-answer tests ------------------------------------------
// clip-space verts v[3][4]; returns accept bits (bit0 accept, bit1 nontrivial) in *cc, the lane-survives flag
// as return value; packed positions, render bbox and edge equations in out arrays.
int orc_probe_triangle(const float* v, uint32_t width, uint32_t height, int cullMode, int guardband,
                       int* cc, uint32_t pos[3], uint32_t bbox[2], int32_t edges[9], float zw[7]) {
    Vec4 vv[3];
    for (int i = 0; i < 3; i++) vv[i] = { v[i * 4 + 0], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3] };
    float bx = guardband ? (float)SWR_MAX_RENDER_SIZE / (float)width : 1.0f;
    float by = guardband ? (float)SWR_MAX_RENDER_SIZE / (float)height : 1.0f;
    *cc = clip_codes(vv, bx, by, nullptr);
    TriSetup t;
    int keep = tri_setup(vv[0], vv[1], vv[2], (int)width / 2, (int)height / 2, cullMode, t);
    pos[0] = t.Pos0; pos[1] = t.Pos1; pos[2] = t.Pos2;
    render_bbox(t, (int)width / 2, (int)height / 2, bbox[0], bbox[1]);
    TriEdges e;
    edge_setup(t, (int)width / 2, (int)height / 2, e);
    int32_t ee[9] = { e.Edge0, e.Edge1, e.Edge2, e.A12, e.A20, e.A01, e.B12, e.B20, e.B01 };
    memcpy(edges, ee, sizeof(ee));
    float f[7] = { e.Z0, e.Z10, e.Z20, e.W0, e.W0S, e.W1S, e.W2S };
    memcpy(zw, f, sizeof(f));
    return keep;
}

// Rasterizer::DrawMeshlets (binned semantics: Rasterizer.cpp:493-644 with 1 worker, i.e. meshlet-ascending,
// primitive-ascending order; non-trivial triangles are counted then dropped :567-569) with the
// VisBufferShader table, opaque fragment program only.
//   color/depth : layer 0 / layer 1 of a 4x4-tiled framebuffer (Rasterizer.h:10-63)
//   meshlets    : scene base pointer; the draw covers [meshletOffset, meshletOffset+count)
//   flags bit0  : guard band enabled (the binned path always enables it, Rasterizer.cpp:509)
//   flags bit1  : EnableClipping on the unbinned path (DrawMeshletsST, Rasterizer.cpp:209-249): non-trivial
//                 triangles go through Clipper::ClipTriangles and DrawTriangle<FS, true>
//   flags bit2  : unbinned path with clipping off — non-trivial triangles are dropped WITHOUT being counted
//                 (the TrianglesClipped increment sits inside the EnableClipping branch, :209-210)
//   counters[4] : TrianglesProcessed, TrianglesRasterized, TrianglesClipped, (unused) — accumulated
// `textures` == NULL: alpha-tested materials (FragmentShaderId 1) are drawn with the opaque program.
// Otherwise they run FS_EncodeSurfaceId<true> (draw_triangle_alpha above); the sampler itself lives in
// oracle_resolve.cpp (orc_sample_implicit_lod_4x4).
void orc_draw_meshlets_ex(uint32_t* color, float* depth, uint32_t width, uint32_t height,
                          const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                          const float* objectToClip, const uint16_t* cullBitmap,
                          const swr_material* materials, const swr_texture_desc* textures,
                          uint32_t flags, uint64_t* counters);

void orc_draw_meshlets(uint32_t* color, float* depth, uint32_t width, uint32_t height,
                       const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                       const float* objectToClip, const uint16_t* cullBitmap,
                       const swr_material* materials, uint32_t flags, uint64_t* counters) {
    orc_draw_meshlets_ex(color, depth, width, height, meshlets, meshletOffset, count, objectToClip, cullBitmap, materials,
                         nullptr, flags, counters);
}

static void draw_meshlets_impl(uint32_t* color, float* depth, uint32_t* ch2, uint32_t width, uint32_t height,
                               const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                               const float* objectToClip, const float* objectToWorld3, const uint16_t* cullBitmap,
                               const swr_material* materials, const swr_texture_desc* textures,
                               uint32_t flags, uint64_t* counters);

void orc_draw_meshlets_ex(uint32_t* color, float* depth, uint32_t width, uint32_t height,
                          const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                          const float* objectToClip, const uint16_t* cullBitmap,
                          const swr_material* materials, const swr_texture_desc* textures,
                          uint32_t flags, uint64_t* counters) {
    draw_meshlets_impl(color, depth, nullptr, width, height, meshlets, meshletOffset, count, objectToClip, nullptr, cullBitmap, materials,
                       textures, flags & ~16u, counters);
}

// Rasterizer::DrawMeshlets with ShadingContext::DeferredShader (Shading.cpp:655): FS_EncodeGBuffer in every fragment slot.
// layers = the three layers of a ShadingContext::NumFbLayers framebuffer (base colour, depth, packed normal/metal/rough);
// flags bits 0-2 as for orc_draw_meshlets_ex.
void orc_draw_meshlets_gbuffer(uint32_t* color, float* depth, uint32_t* ch2, uint32_t width, uint32_t height,
                               const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                               const float* objectToClip, const float* objectToWorld3, const uint16_t* cullBitmap,
                               const swr_material* materials, const swr_texture_desc* textures,
                               uint32_t flags, uint64_t* counters) {
    draw_meshlets_impl(color, depth, ch2, width, height, meshlets, meshletOffset, count, objectToClip, objectToWorld3, cullBitmap, materials,
                       textures, (flags & 7u) | 16u, counters);
}

static void draw_meshlets_impl(uint32_t* color, float* depth, uint32_t* ch2, uint32_t width, uint32_t height,
                               const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count,
                               const float* objectToClip, const float* objectToWorld3, const uint16_t* cullBitmap,
                               const swr_material* materials, const swr_texture_desc* textures,
                               uint32_t flags, uint64_t* counters) {
    int halfW = (int)width / 2, halfH = (int)height / 2;                                   // :508
    float bx = (flags & 1) ? (float)SWR_MAX_RENDER_SIZE / (float)width : 1.0f;             // :509
    float by = (flags & 1) ? (float)SWR_MAX_RENDER_SIZE / (float)height : 1.0f;
    static swr_shaded_meshlet mesh;   // (single-threaded test oracle)
    const bool clipping = (flags & 2) != 0;

    for (uint32_t meshIdx = 0; meshIdx < count; meshIdx++) {
        if (orc_shade_meshlet(meshlets + meshletOffset, meshIdx, cullBitmap, objectToClip, materials, &mesh) == 0) continue;
        counters[0] += mesh.PrimCount;                                                     // :545

        const swr_meshlet& src = meshlets[meshletOffset + meshIdx];
        const bool alpha = mesh.FragmentShaderId == 1 && textures != nullptr;              // Rasterizer.cpp:729 (DrawTriangle[FragmentShaderId])
        const swr_material* mat = alpha ? &materials[src.MaterialId] : nullptr;
        auto draw = [&](const TriSetup& t, uint32_t prim, const float* clipU, const float* clipV) {
            uint32_t bbMin, bbMax;
            render_bbox(t, halfW, halfH, bbMin, bbMax);                                    // :588 / :714
            TriEdges e;
            edge_setup(t, halfW, halfH, e);                                                // :721
            uint32_t surfaceId = (meshletOffset + meshIdx) * SWR_MAX_PRIMS + prim;         // Shading.cpp:328
            if (flags & 8) {                                                               // ShadingContext::OverdrawShader: every slot is FS_Overdraw (Shading.cpp:656)
                draw_triangle_overdraw(color, depth, width, e, bbMin, bbMax);
            } else if (flags & 16) {                                                       // ShadingContext::DeferredShader: every slot is FS_EncodeGBuffer (:655)
                uint32_t i0 = mesh.Indices[0][prim] & 63, i1 = mesh.Indices[1][prim] & 63, i2 = mesh.Indices[2][prim] & 63;
                uint32_t tc[3] = { src.TexCoords[i0], src.TexCoords[i1], src.TexCoords[i2] };
                uint32_t nt[3] = { src.NormalTangents[i0], src.NormalTangents[i1], src.NormalTangents[i2] };
                uint32_t handed = (uint32_t)((src.TangentHandedness >> i0) & 1) << 31;    // Shading.cpp:386
                const swr_material* m = src.MaterialId != SWR_NO_MATERIAL ? &materials[src.MaterialId] : nullptr;
                draw_triangle_gbuffer(color, depth, ch2, width, e, bbMin, bbMax, tc, nt, handed, m ? &textures[m->TextureId] : nullptr,
                                      m ? m->AlphaCutoff : 0, objectToWorld3, clipU, clipV);
            } else if (alpha) {
                uint32_t tc[3] = { src.TexCoords[mesh.Indices[0][prim] & 63], src.TexCoords[mesh.Indices[1][prim] & 63],
                                   src.TexCoords[mesh.Indices[2][prim] & 63] };
                draw_triangle_alpha(color, depth, width, e, bbMin, bbMax, surfaceId, tc, &textures[mat->TextureId], mat->AlphaCutoff, clipU, clipV);
            } else {
                draw_triangle(color, depth, width, e, bbMin, bbMax, surfaceId);
            }
        };

        // 16-wide packets like the reference: a packet's accepted lanes are drawn first, then its clipped ones
        // (Rasterizer.cpp:181-249) — the order only matters for exact depth ties.
        for (uint32_t primOffset = 0; primOffset < mesh.PrimCount; primOffset += 16) {
            uint32_t end = primOffset + 16 < mesh.PrimCount ? primOffset + 16 : mesh.PrimCount;
            uint8_t outcodes[16];
            int codes[16];
            Vec4 pv[16][3];
            for (uint32_t prim = primOffset; prim < end; prim++) {                         // lanes of :550-594
                Vec4* v = pv[prim - primOffset];
                for (int k = 0; k < 3; k++) {
                    uint32_t idx = mesh.Indices[k][prim] & 63;                             // 64-entry permute (SIMD.h:219-230)
                    v[k] = { mesh.Position[0][idx], mesh.Position[1][idx], mesh.Position[2][idx], mesh.Position[3][idx] };
                }
                int cc = codes[prim - primOffset] = clip_codes(v, bx, by, &outcodes[prim - primOffset]);
                if ((cc & 2) && (clipping || !(flags & 4))) counters[2] += 1;              // :567-569 binned, :210 unbinned
                if (!(cc & 1)) continue;
                TriSetup t;
                if (!tri_setup(v[0], v[1], v[2], halfW, halfH, mesh.CullMode, t)) continue;    // :574-577
                counters[1] += 1;                                                          // :579
                draw(t, prim, nullptr, nullptr);
            }
            if (!clipping) continue;
            for (uint32_t prim = primOffset; prim < end; prim++) {                         // Rasterizer.cpp:209-249
                if (!(codes[prim - primOffset] & 2)) continue;
                Vec4 tris[7][3];
                float remapU[7][3], remapV[7][3];
                int n = clip_triangle(pv[prim - primOffset], outcodes[prim - primOffset], bx, by, tris, remapU, remapV);
                for (int j = 0; j < n; j++) {
                    TriSetup t;
                    if (!tri_setup(tris[j][0], tris[j][1], tris[j][2], halfW, halfH, mesh.CullMode, t)) continue;   // FlushPacket :487
                    counters[1] += 1;                                                      // :247
                    draw(t, prim, remapU[j], remapV[j]);
                }
            }
        }
    }
}

// Decode of the packed meshlet format (include/swr_types.h: swr_meshlet_packed), the transport format this library adds for
// the compression planned at Shading.cpp:292-294: every byte but the positions verbatim, position = fmaf(q, Scale, Origin).
// Parity of a packed scene is defined on these decoded floats: the CUDA decode kernel must reproduce them bit for bit.
void orc_unpack_meshlets(const swr_meshlet_packed* src, uint32_t count, swr_meshlet* dst) {
    for (uint32_t m = 0; m < count; m++) {
        memcpy(&dst[m], src[m].Header, 64);
        for (int a = 0; a < 3; a++)
            for (int v = 0; v < SWR_MAX_VERTICES; v++)
                dst[m].Positions[a][v] = std::fmaf((float)src[m].Q[a][v], src[m].Scale[a], src[m].Origin[a]);
        memcpy(dst[m].TexCoords, src[m].TexCoords, sizeof(dst[m].TexCoords));
        memcpy(dst[m].NormalTangents, src[m].NormalTangents, sizeof(dst[m].NormalTangents));
        memcpy(dst[m].Indices, src[m].Indices, sizeof(dst[m].Indices));
    }
}

// Framebuffer::GetPixels — ImageHelpers.cpp:109-147 (de-tile one layer into row-major)
void orc_fb_get_pixels(const uint32_t* layer, uint32_t width, uint32_t height, uint32_t* dest, uint32_t stride) {
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) dest[y * stride + x] = layer[fb_pixel_offset(x, y, width)];
}

// glm helpers needed by CullMeshlets' host-side plane derivation (glm 1.0.1 scalar code paths:
// mat4*mat4 is four column combinations a0*b0 + a1*b1 + a2*b2 + a3*b3 evaluated left to right).
static void mat4_mul(const float* a, const float* b, float* r) {
    float t[16];
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 4; k++)
            t[c * 4 + k] = ((a[0 * 4 + k] * b[c * 4 + 0] + a[1 * 4 + k] * b[c * 4 + 1]) + a[2 * 4 + k] * b[c * 4 + 2]) + a[3 * 4 + k] * b[c * 4 + 3];
    memcpy(r, t, sizeof(t));
}

// Frustum planes of CullMeshlets — Shading.cpp:779-792. planes[i] for i<5 are the ones tested (:806).
void orc_frustum_planes(const float* proj, const float* view, const float* model, float* planes /*[6][4]*/) {
    float pv[16], pvm[16];
    mat4_mul(proj, view, pv);
    mat4_mul(pv, model, pvm);
    // transpose: row i of pvm = (pvm[0*4+i], pvm[1*4+i], pvm[2*4+i], pvm[3*4+i])
    for (int i = 0; i < 3; i++) {
        float a[4], b[4];
        for (int c = 0; c < 4; c++) {
            a[c] = pvm[c * 4 + 3] + pvm[c * 4 + i];
            b[c] = pvm[c * 4 + 3] - pvm[c * 4 + i];
        }
        float la = std::sqrt((a[0] * a[0] + a[1] * a[1]) + a[2] * a[2]);   // glm::length = sqrt(dot)
        float lb = std::sqrt((b[0] * b[0] + b[1] * b[1]) + b[2] * b[2]);
        for (int c = 0; c < 4; c++) {
            planes[(i * 2 + 0) * 4 + c] = a[c] / la;
            planes[(i * 2 + 1) * 4 + c] = b[c] / lb;
        }
    }
}

// CullMeshlets, frustum part — Shading.cpp:795-809, :865-867. Returns the visible count.
uint32_t orc_cull_meshlets(uint16_t* bitmap, const swr_meshlet* meshlets, uint32_t count, const float* planes /*[5][4]*/) {
    uint32_t visibleCount = 0;
    for (uint32_t offset = 0; offset < count; offset += 16) {
        uint16_t bits = 0;
        for (uint32_t l = 0; l < 16 && offset + l < count; l++) {
            const swr_meshlet& m = meshlets[offset + l];
            bool visible = true;
            for (int i = 0; i < 5; i++) {
                const float* p = planes + i * 4;
                // simd::dot(v3,v3) = fma(a.x,b.x, fma(a.y,b.y, a.z*b.z)) — SIMD.h:437
                float dist = std::fmaf(m.BoundCenter[0], p[0], std::fmaf(m.BoundCenter[1], p[1], m.BoundCenter[2] * p[2])) + p[3];
                visible = visible && (dist > -m.BoundRadius);
            }
            if (visible) { bits |= (uint16_t)(1u << l); visibleCount++; }
        }
        bitmap[offset / 16] = bits;
    }
    return visibleCount;
}

// 0 = canonical IEEE division (default); 1 = vrcp14ps + one Newton-Raphson step (sensitivity study only).
// Returns -1 (mode unchanged) when the host CPU has no AVX-512F.
int orc_set_reciprocal_mode(int mode) {
    if ((mode & 1) && !__builtin_cpu_supports("avx512f")) return -1;
    g_rcpMode = mode & 1;
    g_contractDet = (mode >> 1) & 1;        // bit 1: contract the cull determinant into one FMA
    return 0;
}

const char* orc_build_info() {
    return "oracle: scalar canonical restatement; parity pinned against oracle/_ref (the reference's own sources built with g++, tests/test_ref_pin.py)"
#if defined(__FMA__)
           "; hw-fma"
#endif
        ;
}

}  // extern "C"
