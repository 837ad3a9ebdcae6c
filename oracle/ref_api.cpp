// ref_api.cpp — a C wrapper around the REFERENCE's own classes (swr::Rasterizer, swr::Framebuffer, ShadingContext,
// swr::Texture2D), compiled by oracle/ref_build.py together with the reference's Rasterizer.cpp / Shading.cpp /
// ImageHelpers.cpp into oracle/_ref/libswr_ref.so.
//
// TEST INFRASTRUCTURE. Nothing here computes anything: every function marshals flat buffers into the reference's types,
// calls the reference's public entry point and copies the result back, so that tests can pin oracle/oracle*.cpp (the
// restatement) against output of the reference's code itself. The argument lists mirror the orc_* functions of the
// restatement one for one (oracle/orc.py <-> oracle/ref.py).
//
//   ref_draw_meshlets     Rasterizer::DrawMeshlets(fb, count, {VisBufferShader | OverdrawShader | DeferredShader, &ctx})
//                         Rasterizer.h:213, Shading.cpp:649-656
//   ref_resolve           ShadingContext::Resolve                 Shading.cpp:658-732
//   ref_resolve_debug     ShadingContext::ResolveDebug            Shading.cpp:734-773
//   ref_cull_meshlets     ShadingContext::CullMeshlets            Shading.cpp:775-869
//   ref_downsample_depth  texutil::DownsampleDepth                ImageHelpers.cpp:150-247
//   ref_fb_get_pixels     Framebuffer::GetPixels                  ImageHelpers.cpp:109-147
//   ref_generate_mips     Texture2D::GenerateMips / GenerateMip   Texture.h:391-397, :577-597
//   ref_octahedron_from_panorama  texutil::LoadOctahedronFromPanoramaHDR  ImageHelpers.cpp:73-104 (loop re-typed around the reference's pieces)
//   ref_probe_triangle    Clipper::ComputeClipCodes, TrianglePacket::Setup, GetRenderBoundingBox, TriangleEdgeVars::Setup
//                         on a packet whose lane 0 holds the probe triangle (Rasterizer.cpp:257-397)
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include "Shading.h"
#include "compat_camera.h"

#include "swr_types.h"

static_assert(sizeof(Meshlet) == sizeof(swr_meshlet), "Meshlet (Scene.h:15-30)");
static_assert(sizeof(Light) == sizeof(swr_light), "Light (Scene.h:52-75)");
static_assert(sizeof(swr::ShadedMeshlet) == sizeof(swr_shaded_meshlet), "ShadedMeshlet (Rasterizer.h:82-99)");

namespace swr {      // internal to Rasterizer.cpp (:43-80); declared here only for the single-triangle probe
enum class ClipPlane;
struct ClipCodes {
    v_mask AcceptMask;
    v_mask NonTrivialMask;
    uint8_t OutCodes[16];
};
struct Clipper {
    static ClipCodes ComputeClipCodes(v_float4 v0, v_float4 v1, v_float4 v2, glm::vec2 guardBandFactor);
};
}  // namespace swr

namespace {

glm::mat4 to_mat4(const float* m) { glm::mat4 r; memcpy(&r, m, 64); return r; }

struct SceneView {           // reference-typed views of the flat scene arrays; owns the textures it had to create
    std::vector<swr::RgbaTexture2D::Ptr> textures;
    std::vector<Material> materials;
    swr::HdrTexture2D::Ptr skybox;
};

template<typename Tex>
typename Tex::Ptr make_texture(const swr_texture_desc& d) {
    auto tex = swr::CreateTexture<Tex>(d.Width, d.Height, d.MipLevels, d.NumLayers);
    // CreateTexture2D derives RowShift / LayerStride / MipOffsets itself; the caller's descriptor must agree (it is the
    // layout our importer and the CUDA library use).
    bool same = tex->MipLevels == d.MipLevels && tex->RowShift == d.RowShift && tex->LayerStride == d.LayerStride &&
                memcmp(tex->MipOffsets, d.MipOffsets, sizeof(uint32_t) * d.MipLevels) == 0;
    if (!same) return nullptr;
    memcpy(tex->Data, d.Data, (size_t)d.LayerStride * d.NumLayers * 4);
    return tex;
}

bool build_scene(SceneView& sv, const swr_material* materials, uint32_t numMaterials, const swr_texture_desc* textures, uint32_t numTextures) {
    for (uint32_t i = 0; i < numTextures; i++) {
        sv.textures.push_back(make_texture<swr::RgbaTexture2D>(textures[i]));
        if (!sv.textures.back()) return false;
    }
    for (uint32_t i = 0; i < numMaterials; i++) {
        const swr_material& m = materials[i];
        const swr::RgbaTexture2D* tex = (m.TextureId >= 0 && (uint32_t)m.TextureId < numTextures) ? sv.textures[(size_t)m.TextureId].get() : nullptr;
        sv.materials.push_back(Material{ tex, m.IsDoubleSided != 0, m.AlphaCutoff });
    }
    return true;
}

struct FbCopy {              // a reference Framebuffer holding a copy of the caller's layers; written back on scope exit
    swr::FramebufferPtr fb;
    uint32_t* layers;
    FbCopy(uint32_t* layers_, uint32_t width, uint32_t height, uint32_t numLayers) : fb(swr::CreateFramebuffer(width, height, numLayers)), layers(layers_) {
        memcpy(fb->Data, layers, (size_t)fb->LayerStride * numLayers * 4);
    }
    ~FbCopy() { memcpy(layers, fb->Data, (size_t)fb->LayerStride * fb->NumLayers * 4); }
};

}  // namespace

extern "C" {

const char* ref_build_info() {
    return "GLimpSW src/SwRast {Rasterizer,Shading,ImageHelpers}.cpp compiled by g++ " __VERSION__
#ifdef SWR_COMPAT_RCP14
           " [fast: -Ofast -mrecip, vrcp14ps/vrsqrt14ps]";
#else
           " [canonical: IEEE, no contraction, approx_rcp = 1/x]";
#endif
}

void* ref_rasterizer_create(uint32_t numThreads) { return new swr::Rasterizer(numThreads); }
void ref_rasterizer_destroy(void* r) { delete (swr::Rasterizer*)r; }
uint32_t ref_rasterizer_threads(void* r) { return ((swr::Rasterizer*)r)->GetThreadCount(); }

// flags: bit0 EnableGuardband, bit1 EnableClipping, bit2 !EnableBinning; bits 4-5 dispatch table (0 VisBufferShader,
// 1 OverdrawShader, 2 DeferredShader). counters[4] += TrianglesProcessed, TrianglesRasterized, TrianglesClipped, BinQueueFlushes.
int ref_draw_meshlets(void* rasterizer, uint32_t* layers, uint32_t width, uint32_t height, uint32_t numLayers,
                      const void* meshlets, uint32_t meshletOffset, uint32_t count, const float* objectToClip, const float* objectToWorld3,
                      const uint16_t* cullBitmap, const swr_material* materials, uint32_t numMaterials,
                      const swr_texture_desc* textures, uint32_t numTextures, uint32_t flags, uint64_t* counters) {
    auto& raster = *(swr::Rasterizer*)rasterizer;
    SceneView sv;
    if (!build_scene(sv, materials, numMaterials, textures, numTextures)) return -1;
    FbCopy fb(layers, width, height, numLayers);

    ShadingContext ctx = {};
    ctx.Meshlets = (const Meshlet*)meshlets;
    ctx.Materials = sv.materials.data();
    ctx.MeshletOffset = meshletOffset;
    ctx.MeshletCullBitmap = (const v_mask*)cullBitmap;
    ctx.ObjectToClipMat = to_mat4(objectToClip);
    ctx.WorldToClipMat = ctx.ObjectToClipMat;
    ctx.ObjectToWorldMat = glm::mat3(1.0f);
    if (objectToWorld3 != nullptr) memcpy(&ctx.ObjectToWorldMat, objectToWorld3, 36);

    raster.EnableGuardband = (flags & 1) != 0;
    raster.EnableClipping = (flags & 2) != 0;
    raster.EnableBinning = (flags & 4) == 0;
    uint32_t program = (flags >> 4) & 3;
    const swr::ShaderDispatchTable& table = program == 1 ? ShadingContext::OverdrawShader : (program == 2 ? ShadingContext::DeferredShader : ShadingContext::VisBufferShader);

    swr::perf::FlushThreadCounters();
    uint64_t before[4];
    for (int i = 0; i < 4; i++) before[i] = swr::perf::g_GlobalAccum[i];
    raster.DrawMeshlets(*fb.fb, count, { table, &ctx });
    swr::perf::FlushThreadCounters();
    if (counters != nullptr) for (int i = 0; i < 4; i++) counters[i] += swr::perf::g_GlobalAccum[i] - before[i];
    return 0;
}

// invScreenProj == NULL: Resolve derives it from worldToClip (GetInverseScreenProjMatrix on the GLM stand-in); otherwise the
// given matrix is used (see compat_camera.h).
int ref_resolve(void* rasterizer, uint32_t* layers, uint32_t width, uint32_t height, uint32_t numLayers, const void* meshlets, uint32_t meshletOffset,
                const swr_material* materials, uint32_t numMaterials, const swr_texture_desc* textures, uint32_t numTextures,
                const swr_light* lights, uint32_t numLights, const float* objectToClip, const float* objectToWorld3, const float* worldToClip,
                const float* invScreenProj, const float* viewPos, float exposure, const swr_texture_desc* skybox, int debugLayer) {
    auto& raster = *(swr::Rasterizer*)rasterizer;
    SceneView sv;
    if (!build_scene(sv, materials, numMaterials, textures, numTextures)) return -1;
    if (skybox != nullptr) {
        sv.skybox = make_texture<swr::HdrTexture2D>(*skybox);
        if (!sv.skybox) return -1;
    }
    FbCopy fb(layers, width, height, numLayers);

    ShadingContext ctx = {};
    ctx.Meshlets = (const Meshlet*)meshlets;
    ctx.Materials = sv.materials.data();
    ctx.Lights = (const Light*)lights;
    ctx.NumLights = numLights;
    ctx.MeshletOffset = meshletOffset;
    ctx.ObjectToClipMat = to_mat4(objectToClip);
    ctx.WorldToClipMat = to_mat4(worldToClip != nullptr ? worldToClip : objectToClip);
    memcpy(&ctx.ObjectToWorldMat, objectToWorld3, 36);
    ctx.SkyboxTex = sv.skybox.get();
    ctx.ViewPos = glm::vec3(viewPos[0], viewPos[1], viewPos[2]);
    ctx.Exposure = exposure;
    ctx.ShowPerfHeatmap = false;

    swr_compat::g_invScreenProj = invScreenProj;
    if (debugLayer > 0) ctx.ResolveDebug(raster, *fb.fb, (DebugLayer)debugLayer);
    else ctx.Resolve(raster, *fb.fb);
    swr_compat::g_invScreenProj = nullptr;
    return 0;
}

uint32_t ref_cull_meshlets(uint16_t* bitmap, const void* meshlets, uint32_t count, const float* proj, const float* view, const float* model,
                           const float* prevView, float frameW, float frameH, const swr_texture_desc* pyramid) {
    swr::TexturePtr2D<swr::pixfmt::R32f> depthMap;
    if (pyramid != nullptr) {
        depthMap = make_texture<swr::Texture2D<swr::pixfmt::R32f>>(*pyramid);
        if (!depthMap) return 0xFFFFFFFFu;
    }
    return ShadingContext::CullMeshlets((v_mask*)bitmap, (const Meshlet*)meshlets, count, to_mat4(proj), to_mat4(view), to_mat4(model),
                                        to_mat4(prevView), float2(frameW, frameH), depthMap.get());
}

// Builds the depth pyramid of `depthLayer` (layer 1 of a width x height framebuffer) into `data`, laid out as `pyramid` says.
int ref_downsample_depth(const float* depthLayer, uint32_t width, uint32_t height, const swr_texture_desc* pyramid, float* data) {
    auto fb = swr::CreateFramebuffer(width, height, 2);
    memcpy(fb->GetDepthBuffer(), depthLayer, (size_t)width * height * 4);
    swr_texture_desc d = *pyramid;
    std::vector<uint32_t> zeros((size_t)d.LayerStride * d.NumLayers, 0);
    d.Data = zeros.data();
    auto tex = make_texture<swr::Texture2D<swr::pixfmt::R32f>>(d);
    if (!tex) return -1;
    swr::texutil::DownsampleDepth(*fb, *tex);
    memcpy(data, tex->Data, (size_t)d.LayerStride * d.NumLayers * 4);
    return 0;
}

void ref_fb_clear(uint32_t* layers, uint32_t width, uint32_t height, uint32_t numLayers, uint32_t color, float depth) {
    FbCopy fb(layers, width, height, numLayers);
    fb.fb->Clear(color, depth);
}

void ref_fb_get_pixels(const uint32_t* layer, uint32_t width, uint32_t height, uint32_t* dest, uint32_t stride) {
    auto fb = swr::CreateFramebuffer(width, height, 1);
    memcpy(fb->Data, layer, (size_t)width * height * 4);
    std::unique_ptr<uint32_t[], simd::AlignedDeleter> tmp((uint32_t*)_mm_malloc((size_t)stride * height * 4 + 64, 64));   // GetPixels streams to 64-byte aligned rows
    fb->GetPixels(0, tmp.get(), stride);
    memcpy(dest, tmp.get(), (size_t)stride * height * 4);
}

int ref_generate_mips(uint32_t* data, const swr_texture_desc* t) {
    swr_texture_desc d = *t;
    d.Data = data;
    auto tex = make_texture<swr::RgbaTexture2D>(d);
    if (!tex) return -1;
    tex->GenerateMips();       // every level of every layer from level 0 (GenerateMip itself is private)
    memcpy(data, tex->Data, (size_t)d.LayerStride * d.NumLayers * 4);
    return 0;
}

// One 4x4 fragment through Texture2D::SampleImplicitLod<SurfaceSampler> (Texture.h:404-411), SurfaceSampler = Shading.cpp:6-10.
int ref_sample_implicit_lod_4x4(const swr_texture_desc* t, const float* u, const float* v, uint32_t* out) {
    auto tex = make_texture<swr::RgbaTexture2D>(*t);
    if (!tex) return -1;
    constexpr swr::SamplerDesc sd = { .Wrap = swr::WrapMode::Repeat, .MagFilter = swr::FilterMode::Linear, .MinFilter = swr::FilterMode::Nearest };
    v_uint c = tex->SampleImplicitLod<sd>(simd::load<v_float>(u), simd::load<v_float>(v));
    memcpy(out, &c, 64);
    return 0;
}

// texutil::LoadOctahedronFromPanoramaHDR (ImageHelpers.cpp:73-104) on a panorama that is already an HdrTexture2D in memory (the
// file decoder is not in this image): the loop below is that function's, calling the reference's own UnmapOctahedron,
// SampleLevel<PanoSampler>, WriteTile (Pack) and GenerateMips. `cube` describes the output texture (face x face, its mip chain).
int ref_octahedron_from_panorama(const swr_texture_desc* pano, const swr_texture_desc* cube, uint32_t* cubeData) {
    auto panoTex = make_texture<swr::HdrTexture2D>(*pano);
    if (!panoTex) return -1;
    swr_texture_desc d = *cube;
    std::vector<uint32_t> zeros((size_t)d.LayerStride * d.NumLayers, 0);
    d.Data = zeros.data();
    auto cubeTex = make_texture<swr::HdrTexture2D>(d);
    if (!cubeTex) return -1;
    const uint32_t faceSize = panoTex->Width;
    constexpr swr::SamplerDesc PanoSampler = { .Wrap = swr::WrapMode::Repeat, .MagFilter = swr::FilterMode::Linear, .MinFilter = swr::FilterMode::Linear };
    float scaleUV = 1.0f / (faceSize - 1);
    float centerUV = 0.5f * scaleUV;
    for (uint32_t y = 0; y < faceSize; y += 4) {
        for (uint32_t x = 0; x < faceSize; x += 4) {
            v_float u = simd::conv<float>((int32_t)x + swr::TilePixelOffsetsX) * scaleUV + centerUV;
            v_float v = simd::conv<float>((int32_t)y + swr::TilePixelOffsetsY) * scaleUV + centerUV;
            v_float3 dir = swr::texutil::UnmapOctahedron({ u, v });
            for (uint32_t i = 0; i < simd::vec_width; i++) {
                u[i] = atan2f(dir.z[i], dir.x[i]) / simd::tau + 0.5f;
                v[i] = asinf(-dir.y[i]) / simd::pi + 0.5f;
            }
            v_float3 tile = panoTex->SampleLevel<PanoSampler>(u, v, 0, 0);
            cubeTex->WriteTile(tile, x, y);
        }
    }
    cubeTex->GenerateMips();
    memcpy(cubeData, cubeTex->Data, (size_t)d.LayerStride * d.NumLayers * 4);
    return 0;
}

// Same outputs as orc_probe_triangle; the triangle sits in lane 0, the other lanes repeat it.
int ref_probe_triangle(const float* v, uint32_t width, uint32_t height, int cullMode, int guardband,
                       int* cc, uint32_t pos[3], uint32_t bbox[2], int32_t edges[9], float zw[7]) {
    v_float4 p[3];
    for (int i = 0; i < 3; i++) p[i] = v_float4(v_float(v[i * 4 + 0]), v_float(v[i * 4 + 1]), v_float(v[i * 4 + 2]), v_float(v[i * 4 + 3]));
    glm::vec2 gb = guardband ? glm::vec2(swr::Rasterizer::MaxRenderSize / (float)width, swr::Rasterizer::MaxRenderSize / (float)height) : glm::vec2(1.0f);
    swr::ClipCodes codes = swr::Clipper::ComputeClipCodes(p[0], p[1], p[2], gb);
    *cc = (codes.AcceptMask & 1) | ((codes.NonTrivialMask & 1) << 1);

    glm::ivec2 half = glm::ivec2((int)width, (int)height) / 2;
    swr::TrianglePacket tris;
    tris.Pos0 = tris.Pos1 = tris.Pos2 = 0;
    v_mask keep = tris.Setup(p[0], p[1], p[2], half, (swr::FaceCullMode)cullMode);
    pos[0] = (uint32_t)tris.Pos0[0]; pos[1] = (uint32_t)tris.Pos1[0]; pos[2] = (uint32_t)tris.Pos2[0];
    v_uint2 box = tris.GetRenderBoundingBox(half);
    bbox[0] = box.x[0]; bbox[1] = box.y[0];
    swr::TriangleEdgeVars e;
    e.Setup(tris, half);
    int32_t ee[9] = { e.Edge0[0], e.Edge1[0], e.Edge2[0], e.A12[0], e.A20[0], e.A01[0], e.B12[0], e.B20[0], e.B01[0] };
    memcpy(edges, ee, sizeof(ee));
    float f[7] = { e.Z0[0], e.Z10[0], e.Z20[0], e.W0[0], e.W0S[0], e.W1S[0], e.W2S[0] };
    memcpy(zw, f, sizeof(f));
    return keep & 1;
}

}  // extern "C"
