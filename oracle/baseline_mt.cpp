// baseline_mt.cpp — multithreaded AVX-512 restatement of GLimpSW's binned CPU path, used ONLY as the
// timed CPU baseline (bench.py cpu_baseline / --impl reference) and cross-checked against oracle.cpp.
//
// TEST INFRASTRUCTURE ONLY (see oracle.cpp header); pinned through oracle.cpp, which it must equal bit for bit
// (tests/test_baseline_cpu.py) and which is itself pinned to the reference's own code (tests/test_ref_pin.py).
// The reference's sources do compile here with g++ (oracle/ref_build.py), but only on a lane-array stand-in for its
// Clang vector types, which is several times slower than an upstream Clang build would be; timing THAT as "the
// reference" would flatter the GPU. This file is the faster, fairer denominator and is reported as
// "restatement of GLimpSW's AVX-512 path", kind "port", never as the upstream binary.
//
// What it keeps from the reference (Rasterizer.cpp:493-739, Rasterizer.h:250-328, Shading.cpp:281-331):
//   * worker threads over meshlets, 128x128-px bins (BinShift 7, :15), bins rasterized by one owner each;
//   * per triangle: clip-code classification, 28.4 snap, s16 bbox, integer edge equations, top-left rule;
//   * the inner loop walks 4x4-pixel fragments as one 16-lane AVX-512 vector: edge sign test ->
//     mask, depth = fma(u, Z10, fma(v, Z20, Z0)), masked compare against the stored depth tile, masked
//     stores of depth and surface id (FS_EncodeSurfaceId<false>).
// What it does differently, in the CPU's favour: one barrier per draw instead of one per 256-packet
// batch (the reference reports up to 1/3 of its draw time in that sync, README.md:96), per-vertex
// instead of per-corner perspective divide, and contiguous meshlet ranges per worker so bin lists stay
// in meshlet order (= the reference's one-worker tie-break order).
// Its vis-buffer must equal oracle.cpp's bit for bit (tests/test_baseline_cpu.py).
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/swr_types.h"

extern "C" void orc_resolve_rows(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                                 const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                                 const swr_light* lights, uint32_t numLights, const float* objectToClip,
                                 const float* objectToWorld3, const float* invScreenProj, const float* viewPos, float exposure,
                                 uint32_t yBegin, uint32_t yEnd);
extern "C" void orc_resolve_rows_avx512(uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                                 const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                                 const swr_light* lights, uint32_t numLights, const float* objectToClip,
                                 const float* objectToWorld3, const float* invScreenProj, const float* viewPos, float exposure,
                                 uint32_t yBegin, uint32_t yEnd);

extern "C" uint32_t orc_meshlet_setup_avx512(const swr_meshlet* mesh, const float* M, float bx, float by, float fixX, float fixY,
                                             int halfW, int halfH, int cullMode, float* nx, float* ny, float* nz, uint32_t* pos,
                                             uint32_t* fl, uint32_t* keep, uint32_t* bbMin, uint32_t* bbMax);

namespace {

double g_phaseMs[2] = { 0, 0 };
constexpr uint32_t kBinShift = 7, kBinSize = 1u << kBinShift;   // Rasterizer.cpp:15

struct Tri {                 // TrianglePacket lane (Rasterizer.h:145-161)
    uint32_t pos0, pos1, pos2;
    float z0, z1, z2;
    uint32_t id;
    uint32_t bbMin, bbMax;   // GetRenderBoundingBox
};

inline int32_t lo16(uint32_t p) { return (int16_t)(p & 0xFFFF); }
inline int32_t hi16(uint32_t p) { return (int32_t)p >> 16; }
inline uint32_t pack16(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
inline uint32_t pmin16(uint32_t a, uint32_t b) { return pack16(std::min(lo16(a), lo16(b)), std::min(hi16(a), hi16(b))); }
inline uint32_t pmax16(uint32_t a, uint32_t b) { return pack16(std::max(lo16(a), lo16(b)), std::max(hi16(a), hi16(b))); }
inline uint32_t psra16(uint32_t a) { return pack16(lo16(a) >> 4, hi16(a) >> 4); }
inline uint32_t padd16(uint32_t a, uint32_t b) { return pack16(lo16(a) + lo16(b), hi16(a) + hi16(b)); }

inline void render_bbox(uint32_t p0, uint32_t p1, uint32_t p2, int halfW, int halfH, uint32_t& bbMin, uint32_t& bbMax) {   // Rasterizer.cpp:331-351
    uint32_t mn = pmin16(pmin16(p0, p1), p2), mx = pmax16(pmax16(p0, p1), p2);
    mn = psra16(mn + 0x00070007u);
    mx = psra16(mx + 0x00070007u);
    uint32_t vp = (uint32_t)halfW | ((uint32_t)halfH << 16);
    mn = pmin16(pmax16(padd16(mn, vp), 0), vp * 2);
    mx = pmin16(pmax16(padd16(mx, vp), 0), vp * 2);
    bbMin = mn & ~0x00030003u;
    bbMax = (mx + 0x00030003u) & ~0x00030003u;
}

inline int32_t compute_edge(int32_t a, int32_t x, int32_t b, int32_t y) {   // Rasterizer.cpp:291-295
    uint32_t w = (uint32_t)a * (uint32_t)x + (uint32_t)b * (uint32_t)y;
    w += (a > 0 || (a == 0 && b > 0)) ? 0u : 0xFFFFFFFFu;
    return (int32_t)w >> 4;
}

struct Edges { int32_t e0, e1, e2, a12, a20, a01, b12, b20, b01; float z0, z10, z20; };

inline void edge_setup(const Tri& t, int halfW, int halfH, Edges& e) {   // Rasterizer.cpp:296-329
    int32_t x0 = lo16(t.pos0), y0 = hi16(t.pos0), x1 = lo16(t.pos1), y1 = hi16(t.pos1), x2 = lo16(t.pos2), y2 = hi16(t.pos2);
    int32_t A01 = y1 - y0, B01 = x0 - x1, A12 = y2 - y1, B12 = x1 - x2, A20 = y0 - y2, B20 = x2 - x0;
    int32_t det = (int32_t)((uint32_t)B20 * (uint32_t)A01 - (uint32_t)B01 * (uint32_t)A20);
    if (det < 0) { A01 = -A01; B01 = -B01; A12 = -A12; B12 = -B12; A20 = -A20; B20 = -B20; det = (int32_t)(0u - (uint32_t)det); }
    int32_t sx = (int32_t)((uint32_t)(-halfW) << 4) + 8, sy = (int32_t)((uint32_t)(-halfH) << 4) + 8;
    e.e0 = compute_edge(A12, sx - x1, B12, sy - y1);
    e.e1 = compute_edge(A20, sx - x2, B20, sy - y2);
    e.e2 = compute_edge(A01, sx - x0, B01, sy - y0);
    e.a12 = A12; e.a20 = A20; e.a01 = A01; e.b12 = B12; e.b20 = B20; e.b01 = B01;
    float rcpArea = 16.0f / (float)det;
    e.z0 = t.z0; e.z10 = (t.z1 - t.z0) * rcpArea; e.z20 = (t.z2 - t.z0) * rcpArea;
}

// DrawTriangle<FS_EncodeSurfaceId<false>> over [minX,maxX) x [minY,maxY) in 4x4 fragments, 16 lanes.
__attribute__((target("avx512f,avx512bw,avx512dq,avx512vl")))
void draw_triangle_avx512(uint32_t* color, float* depth, uint32_t width, const Edges& e, uint32_t minX, uint32_t minY,
                          uint32_t maxX, uint32_t maxY, uint32_t id) {
    const __m512i laneX = _mm512_setr_epi32(0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3);   // Rasterizer.h:247-248
    const __m512i laneY = _mm512_setr_epi32(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3);
    __m512i px = _mm512_add_epi32(_mm512_set1_epi32((int)minX), laneX), py = _mm512_add_epi32(_mm512_set1_epi32((int)minY), laneY);
#define SWR_ORIGIN(E, A, B) \
    _mm512_add_epi32(_mm512_set1_epi32(E), _mm512_add_epi32(_mm512_mullo_epi32(_mm512_set1_epi32(A), px), _mm512_mullo_epi32(_mm512_set1_epi32(B), py)))
    __m512i row0 = SWR_ORIGIN(e.e0, e.a12, e.b12), row1 = SWR_ORIGIN(e.e1, e.a20, e.b20), row2 = SWR_ORIGIN(e.e2, e.a01, e.b01);
#undef SWR_ORIGIN
    const __m512i sx0 = _mm512_set1_epi32(e.a12 * 4), sx1 = _mm512_set1_epi32(e.a20 * 4), sx2 = _mm512_set1_epi32(e.a01 * 4);
    const __m512i sy0 = _mm512_set1_epi32(e.b12 * 4), sy1 = _mm512_set1_epi32(e.b20 * 4), sy2 = _mm512_set1_epi32(e.b01 * 4);
    const __m512 z0 = _mm512_set1_ps(e.z0), z10 = _mm512_set1_ps(e.z10), z20 = _mm512_set1_ps(e.z20);
    const __m512i vid = _mm512_set1_epi32((int)id);
    for (uint32_t y = minY; y < maxY; y += 4) {
        __m512i e0 = row0, e1 = row1, e2 = row2;
        size_t off = (size_t)(minX << 2) + (size_t)y * width;                        // Rasterizer.h:258
        for (uint32_t x = minX; x < maxX; x += 4, off += 16) {
            __mmask16 m = _mm512_movepi32_mask(_mm512_ternarylogic_epi32(e0, e1, e2, 0x01));   // ~(a|b|c): sign clear in all (Rasterizer.h:289-290)
            if (m) {
                __m512 u = _mm512_cvtepi32_ps(e1), v = _mm512_cvtepi32_ps(e2);
                __m512 d = _mm512_fmadd_ps(u, z10, _mm512_fmadd_ps(v, z20, z0));      // :296
                __m512 old = _mm512_loadu_ps(depth + off);
                m &= _mm512_cmp_ps_mask(d, old, _CMP_GT_OQ);                          // Shading.cpp:311-312
                if (m) {
                    _mm512_mask_storeu_ps(depth + off, m, d);                          // :329
                    _mm512_mask_storeu_epi32(color + off, m, vid);                     // :330
                }
            }
            e0 = _mm512_add_epi32(e0, sx0); e1 = _mm512_add_epi32(e1, sx1); e2 = _mm512_add_epi32(e2, sx2);
        }
        row0 = _mm512_add_epi32(row0, sy0); row1 = _mm512_add_epi32(row1, sy1); row2 = _mm512_add_epi32(row2, sy2);
    }
}

void draw_triangle_scalar(uint32_t* color, float* depth, uint32_t width, const Edges& e, uint32_t minX, uint32_t minY,
                          uint32_t maxX, uint32_t maxY, uint32_t id) {
    for (uint32_t y = minY; y < maxY; y++)
        for (uint32_t x = minX; x < maxX; x++) {
            uint32_t e0 = (uint32_t)e.e0 + (uint32_t)e.a12 * x + (uint32_t)e.b12 * y;
            uint32_t e1 = (uint32_t)e.e1 + (uint32_t)e.a20 * x + (uint32_t)e.b20 * y;
            uint32_t e2 = (uint32_t)e.e2 + (uint32_t)e.a01 * x + (uint32_t)e.b01 * y;
            if ((int32_t)(e0 | e1 | e2) < 0) continue;
            float d = std::fmaf((float)(int32_t)e1, e.z10, std::fmaf((float)(int32_t)e2, e.z20, e.z0));
            uint32_t off = ((x & ~3u) << 2) + (y & ~3u) * width + (x & 3) + (y & 3) * 4;
            if (d > depth[off]) { depth[off] = d; color[off] = id; }
        }
}

struct Pool {
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cvDone;
    std::function<void(uint32_t)> job;
    uint64_t generation = 0;
    uint32_t pending = 0;
    bool stop = false;
    uint32_t n;
    bool avx512;

    explicit Pool(uint32_t n_) : n(n_) {
        avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl");
        for (uint32_t i = 1; i < n; i++) threads.emplace_back([this, i] { worker(i); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> l(mu); stop = true; generation++; }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
    void worker(uint32_t id) {
        uint64_t seen = 0;
        for (;;) {
            std::function<void(uint32_t)> j;
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [&] { return generation != seen; });
                seen = generation;
                if (stop) return;
                j = job;
            }
            j(id);
            { std::lock_guard<std::mutex> l(mu); if (--pending == 0) cvDone.notify_all(); }
        }
    }
    void run(std::function<void(uint32_t)> f) {          // ThreadedRunner::Dispatch (Rasterizer.cpp:848-869)
        if (n == 1) { f(0); return; }
        { std::lock_guard<std::mutex> l(mu); job = f; pending = n - 1; generation++; }
        cv.notify_all();
        f(0);
        std::unique_lock<std::mutex> l(mu);
        cvDone.wait(l, [&] { return pending == 0; });
    }
    // per-worker scratch, reused across draws
    std::vector<std::vector<Tri>> tris;
    std::vector<std::vector<std::vector<uint32_t>>> bins;   // [worker][bin] -> indices into tris[worker]
};

}  // namespace

extern "C" {

void* orc_mt_create(int threads) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    return new Pool((uint32_t)threads);
}
void orc_mt_destroy(void* p) { delete (Pool*)p; }
int orc_mt_threads(void* p) { return (int)((Pool*)p)->n; }
int orc_mt_uses_avx512(void* p) { return ((Pool*)p)->avx512 ? 1 : 0; }

// Framebuffer::Clear (Rasterizer.h:35-48), split over the workers.
void orc_mt_clear(void* p, uint32_t* color, uint32_t* depth, uint32_t numPixels, uint32_t colorValue, uint32_t depthBits) {
    Pool& pool = *(Pool*)p;
    pool.run([&](uint32_t w) {
        size_t b = (size_t)numPixels * w / pool.n, e = (size_t)numPixels * (w + 1) / pool.n;
        for (size_t i = b; i < e; i++) { color[i] = colorValue; depth[i] = depthBits; }
    });
}

// Rasterizer::DrawMeshlets, binned, VisBufferShader opaque program. Same contract as orc_draw_meshlets.
void orc_mt_draw_meshlets(void* p, uint32_t* color, float* depth, uint32_t width, uint32_t height,
                          const swr_meshlet* meshlets, uint32_t meshletOffset, uint32_t count, const float* M,
                          const uint16_t* cullBitmap, const swr_material* materials, uint32_t flags, uint64_t* counters) {
    Pool& pool = *(Pool*)p;
    const int halfW = (int)width / 2, halfH = (int)height / 2;
    const float bx = (flags & 1) ? (float)SWR_MAX_RENDER_SIZE / (float)width : 1.0f;
    const float by = (flags & 1) ? (float)SWR_MAX_RENDER_SIZE / (float)height : 1.0f;
    const float fixX = (float)(halfW * 16), fixY = (float)(halfH * 16);
    const uint32_t binsX = (width + kBinSize - 1) >> kBinShift, binsY = (height + kBinSize - 1) >> kBinShift;
    const uint32_t numBins = binsX * binsY;
    pool.tris.resize(pool.n);
    pool.bins.resize(pool.n);
    std::atomic<uint64_t> cProcessed{0}, cRasterized{0}, cClipped{0};

    // ---- phase 1: mesh shading + setup + binning; worker w owns a contiguous meshlet range
    const auto tPhase0 = std::chrono::steady_clock::now();
    pool.run([&](uint32_t w) {
        auto& tris = pool.tris[w];
        auto& bins = pool.bins[w];
        tris.clear();
        bins.resize(numBins);
        for (auto& b : bins) b.clear();
        uint64_t nProc = 0, nRast = 0, nClip = 0;
        uint32_t mBegin = (uint32_t)((uint64_t)count * w / pool.n), mEnd = (uint32_t)((uint64_t)count * (w + 1) / pool.n);
        float nx[64], ny[64], nz[64];
        uint32_t pos[64], fl[64];
        for (uint32_t meshIdx = mBegin; meshIdx < mEnd; meshIdx++) {
            if (cullBitmap && ((cullBitmap[meshIdx / 16] >> (meshIdx % 16)) & 1) == 0) continue;   // Shading.cpp:282-289
            const swr_meshlet& mesh = meshlets[meshletOffset + meshIdx];
            uint32_t primCount = mesh.NumTriangles;
            if (primCount == 0) continue;
            nProc += primCount;
            int cullMode = SWR_CULL_FRONT_CCW;
            if (mesh.MaterialId != SWR_NO_MATERIAL && materials) cullMode = materials[mesh.MaterialId].IsDoubleSided ? SWR_CULL_NONE : SWR_CULL_FRONT_CCW;
            auto emit = [&](uint32_t prim, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t bbMin, uint32_t bbMax) {
                Tri t;
                t.pos0 = pos[i0]; t.pos1 = pos[i1]; t.pos2 = pos[i2];
                t.bbMin = bbMin; t.bbMax = bbMax;
                t.z0 = nz[i0]; t.z1 = nz[i1]; t.z2 = nz[i2];
                t.id = (meshletOffset + meshIdx) * SWR_MAX_PRIMS + prim;
                uint32_t ti = (uint32_t)tris.size();
                tris.push_back(t);
                // DistributeToBins (Rasterizer.cpp:664-695)
                uint32_t bx0 = (t.bbMin & 0xFFFF) >> kBinShift, by0 = (t.bbMin >> 16) >> kBinShift;
                uint32_t bx1 = ((t.bbMax & 0xFFFF) - 1) >> kBinShift, by1 = ((t.bbMax >> 16) - 1) >> kBinShift;
                for (uint32_t byy = by0; byy <= by1; byy++)
                    for (uint32_t bxx = bx0; bxx <= bx1; bxx++) bins[byy * binsX + bxx].push_back(ti);
            };
            if (pool.avx512) {           // 16 vertices / 16-triangle packets per step, like the reference (oracle/setup_avx512.cpp)
                uint32_t keep[4];
                alignas(64) uint32_t bbMinA[128], bbMaxA[128];
                nClip += orc_meshlet_setup_avx512(&mesh, M, bx, by, fixX, fixY, halfW, halfH, cullMode, nx, ny, nz, pos, fl, keep, bbMinA, bbMaxA);
                for (uint32_t w32 = 0; w32 < 4; w32++)
                    for (uint32_t bitsLeft = keep[w32]; bitsLeft; bitsLeft &= bitsLeft - 1) {
                        const uint32_t prim = w32 * 32 + (uint32_t)__builtin_ctz(bitsLeft);
                        nRast++;
                        emit(prim, mesh.Indices[0][prim] & 63, mesh.Indices[1][prim] & 63, mesh.Indices[2][prim] & 63, bbMinA[prim], bbMaxA[prim]);
                    }
                continue;
            }
            uint32_t nv = std::min(((uint32_t)mesh.NumVertices + 15u) & ~15u, 64u);
            for (uint32_t v = 0; v < nv; v++) {          // ShadeMeshlet + per-vertex part of ComputeClipCodes / Setup
                float x = mesh.Positions[0][v], y = mesh.Positions[1][v], z = mesh.Positions[2][v];
                float cx = std::fmaf(x, M[0], std::fmaf(y, M[4], std::fmaf(z, M[8], 1.0f * M[12])));
                float cy = std::fmaf(x, M[1], std::fmaf(y, M[5], std::fmaf(z, M[9], 1.0f * M[13])));
                float cz = std::fmaf(x, M[2], std::fmaf(y, M[6], std::fmaf(z, M[10], 1.0f * M[14])));
                float cw = std::fmaf(x, M[3], std::fmaf(y, M[7], std::fmaf(z, M[11], 1.0f * M[15])));
                uint32_t f = 0;
                f |= (cx < -cw) ? 1u : 0u; f |= (cx > cw) ? 2u : 0u; f |= (cy < -cw) ? 4u : 0u; f |= (cy > cw) ? 8u : 0u;
                f |= (cz < -cw) ? 16u : 0u; f |= (cz > cw) ? 32u : 0u;
                f |= (std::fabs(cx) < cw * bx && std::fabs(cy) < cw * by) ? 64u : 0u;
                float rw = 1.0f / cw;
                nx[v] = cx * rw; ny[v] = cy * rw; nz[v] = cz * rw;
                float fx = nx[v] * fixX, fy = ny[v] * fixY;
                int32_t X = (fx >= -2147483648.0f && fx < 2147483648.0f) ? (int32_t)std::nearbyintf(fx) : INT32_MIN;
                int32_t Y = (fy >= -2147483648.0f && fy < 2147483648.0f) ? (int32_t)std::nearbyintf(fy) : INT32_MIN;
                pos[v] = ((uint32_t)X & 0xFFFFu) | ((uint32_t)Y << 16);
                fl[v] = f;
            }
            for (uint32_t prim = 0; prim < primCount; prim++) {
                uint32_t i0 = mesh.Indices[0][prim] & 63, i1 = mesh.Indices[1][prim] & 63, i2 = mesh.Indices[2][prim] & 63;
                uint32_t partial = fl[i0] | fl[i1] | fl[i2], combined = fl[i0] & fl[i1] & fl[i2];
                bool visible = (combined & 63u) == 0, trivial = (combined & 64u) && !(partial & 48u);
                if (visible && !trivial) nClip++;
                if (!(visible && trivial)) continue;
                float det = (nx[i2] - nx[i0]) * (ny[i1] - ny[i0]) - (nx[i0] - nx[i1]) * (ny[i0] - ny[i2]);
                if (cullMode != SWR_CULL_FRONT_CCW) { bool flip = cullMode == SWR_CULL_FRONT_CW ? true : det < 0; det = flip ? -det : det; }
                if (!(det > 0)) continue;
                uint32_t bbMin, bbMax;
                render_bbox(pos[i0], pos[i1], pos[i2], halfW, halfH, bbMin, bbMax);
                if (lo16(bbMin) >= lo16(bbMax) || hi16(bbMin) >= hi16(bbMax)) continue;
                nRast++;
                emit(prim, i0, i1, i2, bbMin, bbMax);
            }
        }
        cProcessed += nProc; cRasterized += nRast; cClipped += nClip;
    });

    // ---- phase 2: RasterizeBin (Rasterizer.cpp:696-739); bins handed out dynamically, worker lists in order
    const auto tPhase1 = std::chrono::steady_clock::now();
    std::atomic<uint32_t> nextBin{0};
    pool.run([&](uint32_t) {
        for (;;) {
            uint32_t bin = nextBin.fetch_add(1, std::memory_order_relaxed);
            if (bin >= numBins) break;
            uint32_t binX = (bin % binsX) << kBinShift, binY = (bin / binsX) << kBinShift;
            for (uint32_t w = 0; w < pool.n; w++) {
                const auto& list = pool.bins[w][bin];
                const auto& tris = pool.tris[w];
                for (uint32_t ti : list) {
                    const Tri& t = tris[ti];
                    uint32_t minX = std::max(t.bbMin & 0xFFFF, binX), minY = std::max(t.bbMin >> 16, binY);
                    uint32_t maxX = std::min(t.bbMax & 0xFFFF, binX + kBinSize), maxY = std::min(t.bbMax >> 16, binY + kBinSize);
                    Edges e;
                    edge_setup(t, halfW, halfH, e);
                    if (pool.avx512) draw_triangle_avx512(color, depth, width, e, minX, minY, maxX, maxY, t.id);
                    else draw_triangle_scalar(color, depth, width, e, minX, minY, maxX, maxY, t.id);
                }
            }
        }
    });
    counters[0] += cProcessed; counters[1] += cRasterized; counters[2] += cClipped;
    const auto tPhase2 = std::chrono::steady_clock::now();
    g_phaseMs[0] = std::chrono::duration<double, std::milli>(tPhase1 - tPhase0).count();
    g_phaseMs[1] = std::chrono::duration<double, std::milli>(tPhase2 - tPhase1).count();
}

// Wall time of the two phases of the last orc_mt_draw_meshlets call (setup + binning, bin rasterization), for profiling the baseline.
void orc_mt_phase_ms(double out[2]) { out[0] = g_phaseMs[0]; out[1] = g_phaseMs[1]; }

// ShadingContext::Resolve split by rows of 32 px (Rasterizer::DispatchPass, Rasterizer.h:225-242).
void orc_mt_resolve(void* p, uint32_t* color, const float* depth, uint32_t width, uint32_t height,
                    const swr_meshlet* meshlets, const swr_material* materials, const swr_texture_desc* textures,
                    const swr_light* lights, uint32_t numLights, const float* objectToClip, const float* objectToWorld3,
                    const float* invScreenProj, const float* viewPos, float exposure) {
    Pool& pool = *(Pool*)p;
    std::atomic<uint32_t> nextRow{0};
    pool.run([&](uint32_t) {
        for (;;) {
            uint32_t y = nextRow.fetch_add(32, std::memory_order_relaxed);
            if (y >= height) break;
            // 16 pixels per step like the reference (one 4x4 fragment = one AVX-512 vector) where the CPU has AVX-512;
            // the scalar spec otherwise. Both produce the same bits (tests/test_baseline_cpu.py).
            if (pool.avx512)
                orc_resolve_rows_avx512(color, depth, width, height, meshlets, materials, textures, lights, numLights, objectToClip,
                                        objectToWorld3, invScreenProj, viewPos, exposure, y, std::min(y + 32, height));
            else
                orc_resolve_rows(color, depth, width, height, meshlets, materials, textures, lights, numLights, objectToClip,
                                 objectToWorld3, invScreenProj, viewPos, exposure, y, std::min(y + 32, height));
        }
    });
}

}  // extern "C"
