// swr_b200.hpp — header-only C++ shim over the C ABI (swrb.h) that keeps the reference's own names
// (SURVEY.md §8b): swr::Rasterizer / swr::Framebuffer / ShadingContext::{CullMeshlets, Resolve} /
// texutil::DownsampleDepth call sites port 1:1.
//
// The shim does not include any reference header. The calls that take a ShadingContext are templates over the
// context type and read the reference's field names (Shading.h:12-39): MeshletOffset, MeshletCullBitmap,
// ObjectToClipMat, WorldToClipMat, ObjectToWorldMat, ViewPos, Exposure. Matrices are anything whose first element
// can be addressed as `&m[0][0]` with column-major floats (glm::mat4 / glm::mat3).
// Errors: every non-zero status becomes a std::runtime_error carrying swrb_last_error().
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <utility>
#include <vector>

#include "swrb.h"

namespace swrb200 {

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(swrb_last_error());
}

// swr::Framebuffer (Rasterizer.h:10-63) living in HBM.
class Framebuffer {
public:
    Framebuffer() = default;
    Framebuffer(swrb_device* dev, uint32_t width, uint32_t height, uint32_t numLayers) : Width(width), Height(height), NumLayers(numLayers) {
        check(swrb_fb_create(dev, width, height, numLayers, &_h));
    }
    Framebuffer(Framebuffer&& o) noexcept { *this = std::move(o); }
    Framebuffer& operator=(Framebuffer&& o) noexcept {
        if (this != &o) { reset(); _h = o._h; Width = o.Width; Height = o.Height; NumLayers = o.NumLayers; o._h = nullptr; }
        return *this;
    }
    Framebuffer(const Framebuffer&) = delete;
    Framebuffer& operator=(const Framebuffer&) = delete;
    ~Framebuffer() { reset(); }

    uint32_t Width = 0, Height = 0, NumLayers = 0;

    void Clear(uint32_t color, float depth) { check(swrb_fb_clear(_h, color, depth)); }                       // Rasterizer.h:35-38
    void ClearLayer(uint32_t layer, uint32_t value) { check(swrb_fb_clear_layer(_h, layer, value)); }         // :40-48
    void GetPixels(uint32_t layer, uint32_t* dest, uint32_t stride) { check(swrb_fb_get_pixels(_h, layer, dest, stride)); }   // ImageHelpers.cpp:109
    // no reference counterpart: sort-first split of one view over several GPUs — only rows [y0, y1) are drawn, resolved and read back
    void SetScissorRows(uint32_t y0, uint32_t y1) { check(swrb_fb_set_scissor_rows(_h, y0, y1)); }
    // GetLayerData(layer) copies: the raw 4x4-tiled words of one layer
    void DownloadLayer(uint32_t layer, uint32_t* destTiled) { check(swrb_fb_download_tiled(_h, layer, destTiled)); }
    void UploadLayer(uint32_t layer, const uint32_t* srcTiled) { check(swrb_fb_upload_tiled(_h, layer, srcTiled)); }
    static uint32_t GetPixelOffset(uint32_t x, uint32_t y, uint32_t width) {                                   // Rasterizer.h:50-56
        return ((x & ~3u) << 2) + (y & ~3u) * width + (x & 3u) + (y & 3u) * 4u;
    }
    swrb_fb* handle() const { return _h; }

private:
    void reset() { if (_h) swrb_fb_destroy(_h); _h = nullptr; }
    swrb_fb* _h = nullptr;
};

// Texture2D<R32f> depth pyramid + texutil::DownsampleDepth (ImageHelpers.cpp:150-247).
class DepthPyramid {
public:
    DepthPyramid() = default;
    DepthPyramid(swrb_device* dev, uint32_t fbWidth, uint32_t fbHeight) { check(swrb_hiz_create(dev, fbWidth, fbHeight, &_h)); }
    DepthPyramid(DepthPyramid&& o) noexcept : _h(o._h) { o._h = nullptr; }
    DepthPyramid& operator=(DepthPyramid&& o) noexcept { if (this != &o) { reset(); _h = o._h; o._h = nullptr; } return *this; }
    DepthPyramid(const DepthPyramid&) = delete;
    DepthPyramid& operator=(const DepthPyramid&) = delete;
    ~DepthPyramid() { reset(); }

    void Downsample(Framebuffer& fb) { check(swrb_hiz_build(_h, fb.handle())); }                               // texutil::DownsampleDepth(fb, depthMap)
    swr_texture_desc Info() const { swr_texture_desc d{}; check(swrb_hiz_info(_h, &d)); return d; }
    swrb_hiz* handle() const { return _h; }

private:
    void reset() { if (_h) swrb_hiz_destroy(_h); _h = nullptr; }
    swrb_hiz* _h = nullptr;
};

// swr::Rasterizer (Rasterizer.h:198-248) + the scene-resident half of ShadingContext.
class Rasterizer {
public:
    bool EnableBinning = true, EnableClipping = true, EnableGuardband = true;                                 // Rasterizer.h:206-208

    explicit Rasterizer(int cudaDevice = 0) { check(swrb_device_create(cudaDevice, &_dev)); }
    Rasterizer(const Rasterizer&) = delete;
    Rasterizer& operator=(const Rasterizer&) = delete;
    ~Rasterizer() {
        if (_scene) swrb_scene_destroy(_scene);
        if (_dev) swrb_device_destroy(_dev);
    }

    Framebuffer CreateFramebuffer(uint32_t width, uint32_t height, uint32_t numLayers = 2) {                  // swr::CreateFramebuffer, Rasterizer.h:66-78
        return Framebuffer(_dev, width, height, numLayers);
    }
    DepthPyramid CreateDepthPyramid(uint32_t fbWidth, uint32_t fbHeight) { return DepthPyramid(_dev, fbWidth, fbHeight); }   // Main.cpp:54-56

    // Scene::{Meshlets, Materials, Textures, Lights} (Scene.h:117-123), once after import. Material::Texture host
    // pointers become indices into `textures`; swr_texture_desc::Data points at Texture2D::Data as is.
    void UploadScene(const swr_meshlet* meshlets, uint32_t numMeshlets, const std::vector<swr_material>& materials,
                     const std::vector<swr_texture_desc>& textures, const swr_light* lights, uint32_t numLights) {
        if (_scene) { swrb_scene_destroy(_scene); _scene = nullptr; }
        check(swrb_scene_create(_dev, meshlets, numMeshlets, materials.data(), (uint32_t)materials.size(), textures.data(),
                                (uint32_t)textures.size(), lights, numLights, &_scene));
    }

    // ShadingContext::SkyboxTex = tex (Shading.h:29, Main.cpp:186): an octahedron-mapped HdrTexture2D; nullptr removes it
    void SetSkybox(const swr_texture_desc* hdrTexture) { check(swrb_scene_set_skybox(_scene, hdrTexture)); }

    // ShadingContext::CullMeshlets(bitmap, meshlets + offset, count, P, V, M, prevV, frameSize, depthMap) — Shading.cpp:775
    template <class Mat4>
    uint32_t CullMeshlets(uint16_t* bitmap, uint32_t meshletOffset, uint32_t count, const Mat4& proj, const Mat4& view, const Mat4& model,
                          const Mat4& prevView, float frameW, float frameH, DepthPyramid* depthMap = nullptr) {
        uint32_t visible = 0;
        check(swrb_cull_meshlets_hiz(_scene, meshletOffset, count, &proj[0][0], &view[0][0], &model[0][0], &prevView[0][0], frameW, frameH,
                                     depthMap ? depthMap->handle() : nullptr, bitmap, &visible));
        return visible;
    }

    // Rasterizer::DrawMeshlets(fb, count, {table, &ctx}) — Rasterizer.cpp:493. `table` names the reference's shader table:
    // SWRB_PROGRAM_VISBUFFER = ShadingContext::VisBufferShader, SWRB_PROGRAM_OVERDRAW = ShadingContext::OverdrawShader,
    // SWRB_PROGRAM_DEFERRED = ShadingContext::DeferredShader (3-layer framebuffer; ctx.ObjectToWorldMat is passed along)
    // (the choice the Playground makes per frame, Main.cpp:204-209).
    template <class Ctx>
    void DrawMeshlets(Framebuffer& fb, uint32_t count, const Ctx& ctx, swrb_program table = SWRB_PROGRAM_VISBUFFER) {
        apply_flags();
        swrb_draw_desc d{};
        d.MeshletOffset = ctx.MeshletOffset;
        d.MeshletCount = count;
        std::memcpy(d.ObjectToClip, &ctx.ObjectToClipMat[0][0], sizeof d.ObjectToClip);
        std::memcpy(d.ObjectToWorld, &ctx.ObjectToWorldMat[0][0], sizeof d.ObjectToWorld);
        d.CullBitmapHost = reinterpret_cast<const uint16_t*>(ctx.MeshletCullBitmap);
        check(swrb_draw_batch_program(fb.handle(), _scene, &d, 1, table));
    }
    // The per-node loop of Main.cpp:216-240 as one submission.
    void DrawBatch(Framebuffer& fb, const std::vector<swrb_draw_desc>& draws, swrb_program table = SWRB_PROGRAM_VISBUFFER) {
        apply_flags();
        check(swrb_draw_batch_program(fb.handle(), _scene, draws.data(), (uint32_t)draws.size(), table));
    }

    // ShadingContext::Resolve(rast, fb) — Shading.cpp:658. invScreenProj = GetInverseScreenProjMatrix(ctx.WorldToClipMat,
    // {fb.Width, fb.Height}) (Camera.h:140-146), computed by the caller exactly as today.
    template <class Ctx, class Mat4>
    void Resolve(Framebuffer& fb, const Ctx& ctx, const Mat4& invScreenProj) {
        const swrb_shading_uniforms u = uniforms(ctx, invScreenProj);
        check(swrb_resolve(fb.handle(), _scene, &u));
    }
    // ShadingContext::ResolveDebug(rast, fb, layer) — Shading.cpp:734; layer = enum class DebugLayer (Shading.h:8)
    template <class Ctx, class Mat4>
    void ResolveDebug(Framebuffer& fb, const Ctx& ctx, const Mat4& invScreenProj, swrb_debug_layer layer) {
        const swrb_shading_uniforms u = uniforms(ctx, invScreenProj);
        check(swrb_resolve_debug(fb.handle(), _scene, &u, layer));
    }

    // perf::GetCurrent(PerfCounter::…) — Rasterizer.h:381-395
    uint64_t GetCounter(swr_perf_counter which) {
        uint64_t c[SWR_PERF_Count_] = {};
        check(swrb_get_counters(_dev, c));
        return c[which];
    }
    void ResetCounters() { check(swrb_reset_counters(_dev)); }
    void Sync() { check(swrb_sync(_dev)); }

    swrb_device* device() const { return _dev; }
    swrb_scene* scene() const { return _scene; }

private:
    template <class Ctx, class Mat4>
    static swrb_shading_uniforms uniforms(const Ctx& ctx, const Mat4& invScreenProj) {
        swrb_shading_uniforms u{};
        std::memcpy(u.WorldToClip, &ctx.WorldToClipMat[0][0], sizeof u.WorldToClip);
        std::memcpy(u.ObjectToClip, &ctx.ObjectToClipMat[0][0], sizeof u.ObjectToClip);
        std::memcpy(u.ObjectToWorld, &ctx.ObjectToWorldMat[0][0], sizeof u.ObjectToWorld);
        std::memcpy(u.InvScreenProj, &invScreenProj[0][0], sizeof u.InvScreenProj);
        std::memcpy(u.ViewPos, &ctx.ViewPos[0], sizeof u.ViewPos);
        u.Exposure = ctx.Exposure;
        return u;
    }
    void apply_flags() {
        check(swrb_device_set_flags(_dev, (EnableBinning ? SWRB_FLAG_BINNING : 0u) | (EnableClipping ? SWRB_FLAG_CLIPPING : 0u) |
                                              (EnableGuardband ? SWRB_FLAG_GUARDBAND : 0u)));
    }
    swrb_device* _dev = nullptr;
    swrb_scene* _scene = nullptr;
};

}  // namespace swrb200
