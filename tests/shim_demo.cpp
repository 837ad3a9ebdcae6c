// Test program for include/swr_b200.hpp: the host code of a GLimpSW frame (Main.cpp:213-252) written against
// the shim with a stand-in ShadingContext (same field names as Shading.h:12-39, plain float arrays instead of glm).
// Prints "NO_DEVICE <reason>" when no GPU / driver is present, else "OK <covered> <id at centre> <rasterized>".
#include <cstdio>
#include <vector>

#include "swr_b200.hpp"

struct Mat4 { float m[4][4]; float* operator[](int c) { return m[c]; } const float* operator[](int c) const { return m[c]; } };
struct Mat3 { float m[3][3]; float* operator[](int c) { return m[c]; } const float* operator[](int c) const { return m[c]; } };

struct ShadingContext {            // the fields the shim reads
    uint32_t MeshletOffset = 0;
    const uint8_t* MeshletCullBitmap = nullptr;
    Mat4 WorldToClipMat{}, ObjectToClipMat{};
    Mat3 ObjectToWorldMat{};
    float ViewPos[3] = { 0, 0, 0 };
    float Exposure = 1.0f;
};

static Mat4 identity4() { Mat4 r{}; for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f; return r; }

int main() {
    try {
        swrb200::Rasterizer rast(0);
        rast.EnableBinning = true;

        std::vector<swr_meshlet> meshlets(1);
        std::memset(meshlets.data(), 0, sizeof(swr_meshlet));
        swr_meshlet& m = meshlets[0];
        m.NumVertices = 3; m.NumTriangles = 1; m.AlphaCutoff = 255; m.MaterialId = SWR_NO_MATERIAL;
        const float pos[3][3] = { { -0.5f, -0.5f, 0.5f }, { -0.5f, 0.5f, 0.5f }, { 0.5f, -0.5f, 0.5f } };   // front-facing: det > 0
        for (int v = 0; v < 3; v++) for (int k = 0; k < 3; k++) m.Positions[k][v] = pos[v][k];
        m.Indices[0][0] = 0; m.Indices[1][0] = 1; m.Indices[2][0] = 2;
        m.BoundRadius = 2.0f;
        rast.UploadScene(meshlets.data(), 1, {}, {}, nullptr, 0);

        const uint32_t W = 64, H = 64;
        swrb200::Framebuffer fb = rast.CreateFramebuffer(W, H);
        fb.Clear(0xFFFFFFFFu, 0.0f);

        ShadingContext ctx;
        ctx.ObjectToClipMat = identity4();
        ctx.WorldToClipMat = identity4();
        uint16_t bitmap[1] = { 0 };
        Mat4 I = identity4();
        uint32_t visible = rast.CullMeshlets(bitmap, 0, 1, I, I, I, I, (float)W, (float)H, nullptr);
        ctx.MeshletCullBitmap = reinterpret_cast<const uint8_t*>(bitmap);
        rast.DrawMeshlets(fb, 1, ctx);

        std::vector<uint32_t> ids(W * H);
        fb.GetPixels(0, ids.data(), W);
        uint32_t covered = 0;
        for (uint32_t v : ids) covered += v != 0xFFFFFFFFu;
        const unsigned long long rasterized = rast.GetCounter(SWR_PERF_TrianglesRasterized);

        // the Playground's overdraw view (Main.cpp:204-209, :251): OverdrawShader, then ResolveDebug(OverdrawPixel)
        swrb200::Framebuffer od = rast.CreateFramebuffer(W, H);
        od.Clear(0u, 0.0f);
        rast.DrawMeshlets(od, 1, ctx, SWRB_PROGRAM_OVERDRAW);
        rast.DrawMeshlets(od, 1, ctx, SWRB_PROGRAM_OVERDRAW);
        std::vector<uint32_t> counts(W * H);
        od.GetPixels(0, counts.data(), W);
        uint32_t twice = 0;
        for (uint32_t v : counts) twice += (v >> 16) == 2u;
        rast.ResolveDebug(od, ctx, I, SWRB_LAYER_OVERDRAW_PIXEL);
        od.GetPixels(0, counts.data(), W);
        std::printf("OK %u %u %llu %u %u %08x\n", covered, ids[(H / 2 - 4) * W + (W / 2 - 4)], rasterized, visible, twice, counts[0]);
    } catch (const std::exception& e) {
        std::printf("NO_DEVICE %s\n", e.what());
    }
    return 0;
}
