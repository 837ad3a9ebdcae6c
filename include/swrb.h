/* swrb.h — C ABI of the B200-native meshlet raster path (libswrb.so).
 *
 * This is the drop-in boundary for ONE path of GLimpSW: meshlet cull -> mesh shade ->
 * triangle setup -> bin -> tile raster -> vis-buffer -> resolve. Each entry point names
 * the reference interface it replaces (paths relative to /root/reference/). Plain
 * pointers and sizes only; no C++/torch types. All functions return 0 on success or a
 * negative swrb_status; swrb_last_error() gives a thread-local message. Nothing here
 * has a CPU fallback: without a CUDA device every call fails with SWRB_E_CUDA.
 *
 * Threading (Rasterizer.cpp:852 — the reference Rasterizer is single-caller): one
 * swrb_device = one CUDA device + one stream; calls enqueue asynchronously on that
 * stream; swrb_sync() and the download calls synchronise. Use one device object per GPU.
 */
#ifndef SWRB_H
#define SWRB_H

#include "swr_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SWRB_API
#else
#define SWRB_API __attribute__((visibility("default")))
#endif

typedef enum swrb_status {
    SWRB_OK = 0,
    SWRB_E_INVALID = -1,       /* bad argument (null handle, w/h not multiple of 4, > 2896, ...) */
    SWRB_E_CUDA = -2,          /* CUDA runtime error, message in swrb_last_error() */
    SWRB_E_OOM = -3,
    SWRB_E_BIN_OVERFLOW = -4,  /* device-side work list overflowed; nothing was dropped silently:
                                  the draw was aborted before touching the framebuffer */
    SWRB_E_UNSUPPORTED = -5
} swrb_status;

typedef struct swrb_device swrb_device;
typedef struct swrb_scene swrb_scene;
typedef struct swrb_fb swrb_fb;
struct swrb_peer_sync;

/* Rasterizer public toggles — Rasterizer.h:206-208 */
enum {
    SWRB_FLAG_BINNING   = 1u << 0,  /* EnableBinning: screen-tile binner + shared-memory tile raster.
                                       Off = direct path (per-triangle 64-bit atomicMax into HBM keys). */
    SWRB_FLAG_CLIPPING  = 1u << 1,  /* EnableClipping (only meaningful with binning off, Rasterizer.cpp:209) */
    SWRB_FLAG_GUARDBAND = 1u << 2,  /* EnableGuardband (ignored by the binned path, Rasterizer.cpp:509) */
    SWRB_FLAG_FUSED_FRUSTUM_CULL = 1u << 3, /* evaluate CullMeshlets' frustum test inside the mesh kernel
                                               (planes from swrb_draw_desc.FrustumPlanes) */
    SWRB_FLAG_NO_RESOLVE_CACHE = 1u << 4,   /* the mesh kernel does not keep per-vertex x/w, y/w, 1/w for the resolve pass;
                                               swrb_resolve then re-transforms every pixel's three corners itself
                                               (same bits either way — saves 1 KB of writes per drawn meshlet for
                                               callers that never resolve) */
    SWRB_FLAGS_DEFAULT = SWRB_FLAG_BINNING | SWRB_FLAG_CLIPPING | SWRB_FLAG_GUARDBAND
};

/* One DrawMeshlets call: Rasterizer::DrawMeshlets(fb, count, {VisBufferShader, &ctx})
 * (Rasterizer.h:213, Main.cpp:236-240) with the ShadingContext per-instance uniforms
 * (Shading.h:21-26) it reads. */
typedef struct swrb_draw_desc {
    uint32_t MeshletOffset;        /* ShadingContext::MeshletOffset */
    uint32_t MeshletCount;         /* `count` */
    float    ObjectToClip[16];     /* ShadingContext::ObjectToClipMat, column-major (glm::mat4) */
    const uint16_t* CullBitmapHost;/* ShadingContext::MeshletCullBitmap (host, 1 bit/meshlet) or NULL */
    int32_t  UseDeviceCullBitmap;  /* 1: use the bitmap the last swrb_cull_meshlets left on the device */
    float    FrustumPlanes[5][4];  /* only read with SWRB_FLAG_FUSED_FRUSTUM_CULL */
    float    ObjectToWorld[9];     /* ShadingContext::ObjectToWorldMat (glm::mat3, column-major): only read by
                                      SWRB_PROGRAM_DEFERRED (FS_EncodeGBuffer rotates the normals with it) */
} swrb_draw_desc;

/* ShadingContext resolve-pass uniforms — Shading.h:21-33 */
typedef struct swrb_shading_uniforms {
    float WorldToClip[16];         /* column-major */
    float ObjectToClip[16];
    float ObjectToWorld[9];        /* glm::mat3, column-major */
    float InvScreenProj[16];       /* GetInverseScreenProjMatrix(WorldToClip, size) (Camera.h:140-146), host-computed */
    float ViewPos[3];
    float Exposure;
} swrb_shading_uniforms;

/* ---- device ------------------------------------------------------------------------ */
SWRB_API int swrb_device_create(int cuda_device, swrb_device** out);   /* Rasterizer::Rasterizer (Rasterizer.cpp:125) */
SWRB_API void swrb_device_destroy(swrb_device* dev);
SWRB_API int swrb_device_set_stream(swrb_device* dev, void* cuda_stream); /* borrow an external cudaStream_t (NULL = own) */
SWRB_API int swrb_device_set_flags(swrb_device* dev, uint32_t flags);  /* EnableBinning/Clipping/Guardband */
/* Size of the mesh kernel's persistent grid in blocks (of 8 warps) per SM, 1..4, default 4. A lone frame is fastest with
 * the whole register file (4); a caller that keeps several render contexts in flight on one GPU gets more frames per
 * second with 2: the mesh kernel then leaves half of every SM to the other contexts' resolve blocks, and the
 * latency-bound mesh warps and the issue-bound resolve warps fill each other's idle issue slots. */
SWRB_API int swrb_device_set_mesh_occupancy(swrb_device* dev, uint32_t blocks_per_sm);
SWRB_API int swrb_device_reserve(swrb_device* dev, uint64_t max_triangles, uint64_t max_bin_entries);
SWRB_API int swrb_sync(swrb_device* dev);
SWRB_API const char* swrb_last_error(void);
SWRB_API const char* swrb_version(void);

/* perf::GetCurrent / perf::Reset (Rasterizer.h:381-395). Integer counters come from device atomics. */
SWRB_API int swrb_get_counters(swrb_device* dev, uint64_t out[SWR_PERF_Count_]);
SWRB_API int swrb_reset_counters(swrb_device* dev);

/* ---- scene (Scene::Meshlets/Materials/Textures/Lights, Scene.h:117-123) -------------- */
SWRB_API int swrb_scene_create(swrb_device* dev,
                               const swr_meshlet* meshlets, uint32_t num_meshlets,
                               const swr_material* materials, uint32_t num_materials,
                               const swr_texture_desc* textures, uint32_t num_textures,
                               const swr_light* lights, uint32_t num_lights,
                               swrb_scene** out);
SWRB_API int swrb_scene_update_meshlets(swrb_scene* scene, const swr_meshlet* meshlets, uint32_t first, uint32_t count);
/* The same from PACKED meshlets (swr_meshlet_packed, include/swr_types.h: 16-bit positions in the meshlet's bounding box, 1376
 * instead of 1728 bytes — the "meshlet compression" the reference's author planned at Shading.cpp:292-294, as an import-time /
 * transport format): the packed bytes cross PCIe, a decode kernel expands them into the reference layout in device memory,
 * every other call is unchanged. The decoded positions are fmaf((float)Q, Scale, Origin) exactly (oracle: orc_unpack_meshlets),
 * so results equal those of swrb_scene_create on the host-decoded meshlets bit for bit. */
SWRB_API int swrb_scene_create_packed(swrb_device* dev,
                                      const swr_meshlet_packed* meshlets, uint32_t num_meshlets,
                                      const swr_material* materials, uint32_t num_materials,
                                      const swr_texture_desc* textures, uint32_t num_textures,
                                      const swr_light* lights, uint32_t num_lights,
                                      swrb_scene** out);
SWRB_API int swrb_scene_update_packed(swrb_scene* scene, const swr_meshlet_packed* meshlets, uint32_t first, uint32_t count);
/* Reads meshlets [first, first + count) of the scene back from device memory (e.g. to inspect what a packed upload decoded to). */
SWRB_API int swrb_scene_download_meshlets(swrb_scene* scene, swr_meshlet* dst_host, uint32_t first, uint32_t count);
/* The scene's meshlet array in device memory (num_meshlets x 1728 bytes), for callers that fill it on the device — e.g.
 * each rank of a multi-GPU job uploads 1/N of the scene from the host and the ranks all-gather the rest over NVLink.
 * After writing, swrb_scene_touch(first, count) tells the library that derived data of that range is stale; the writes
 * must be ordered before later swrb calls by the caller (events / stream order on the device's stream). */
SWRB_API int swrb_scene_meshlets_device(swrb_scene* scene, void** device_ptr_out);
SWRB_API int swrb_scene_touch(swrb_scene* scene, uint32_t first, uint32_t count);
/* ShadingContext::SkyboxTex (Shading.h:29, Main.cpp:186): an HdrTexture2D = Texture2D<pixfmt::R11G11B10f> in the reference's
 * TiledY8 storage with its mip chain (Texture.h:133-200, :216), octahedron-mapped. With a skybox set, swrb_resolve gives sky
 * pixels SkyboxTex->SampleOctLevel<EnvSampler>(worldPos - ViewPos, 1) (Shading.cpp:676-679) instead of colour 0.
 * NULL removes it. The texels are copied; the pointer is only borrowed for the call. */
SWRB_API int swrb_scene_set_skybox(swrb_scene* scene, const swr_texture_desc* hdr_texture);
SWRB_API void swrb_scene_destroy(swrb_scene* scene);

/* ---- framebuffer (CreateFramebuffer Rasterizer.h:66-78; Clear/ClearLayer :35-48;
 *      GetPixels ImageHelpers.cpp:109-147) -------------------------------------------- */
SWRB_API int swrb_fb_create(swrb_device* dev, uint32_t width, uint32_t height, uint32_t num_layers, swrb_fb** out);
SWRB_API void swrb_fb_destroy(swrb_fb* fb);
SWRB_API int swrb_fb_info(const swrb_fb* fb, swr_fb_info* out);
/* Sort-last composition: after a draw the framebuffer's result lives in its 64-bit key buffer — one word per
 * pixel in the 4x4-tiled order, depth bits << 32 | (0xFFFFFFFE - draw-order rank), 0x00000000FFFFFFFF-style seeds where nothing
 * was drawn — and the maximum of two such words is the fragment the reference's sequential depth test would keep when both
 * triangles were submitted in one scene. GPUs that each drew a DISJOINT subset of the same scene's meshlets (same framebuffer
 * size, same clear) therefore composite bit-exactly with an element-wise max of their key buffers (ncclMax on int64: every key
 * is a positive int64). swrb_fb_keys_device hands out the buffer (width * height words; SWRB_E_INVALID unless a vis-buffer
 * draw is pending in the keys); after writing the combined keys back, swrb_fb_keys_touched tells the library that the keys no
 * longer come from this device's last batch alone (the resolve pass then re-derives vertices instead of using its cache).
 * The caller orders its writes against the device's stream. */
SWRB_API int swrb_fb_keys_device(swrb_fb* fb, void** device_ptr_out, uint64_t* num_words_out);
SWRB_API int swrb_fb_keys_touched(swrb_fb* fb);
SWRB_API int swrb_fb_clear(swrb_fb* fb, uint32_t color, float depth);
SWRB_API int swrb_fb_clear_layer(swrb_fb* fb, uint32_t layer, uint32_t value);
SWRB_API int swrb_fb_download_tiled(swrb_fb* fb, uint32_t layer, uint32_t* dst_host);  /* raw GetLayerData copy */
SWRB_API int swrb_fb_upload_tiled(swrb_fb* fb, uint32_t layer, const uint32_t* src_host);
SWRB_API int swrb_fb_get_pixels(swrb_fb* fb, uint32_t layer, uint32_t* dst_host, uint32_t stride); /* Framebuffer::GetPixels */
/* Same without the final synchronisation: dst_host should be pinned; the pixels are valid after swrb_sync(). The de-tile
 * runs on the device's stream, the PCIe copy on a copy stream of the device (three images in flight), so later frames render
 * while this one's pixels travel. */
SWRB_API int swrb_fb_get_pixels_async(swrb_fb* fb, uint32_t layer, uint32_t* dst_host, uint32_t stride);
SWRB_API int swrb_fb_get_pixels_device(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride); /* same, device dst */
/* Same, launched on a caller-provided stream (the caller orders it after the producing call with events);
 * dst may be peer memory of another GPU: the de-tile kernel then stores straight over NVLink. */
SWRB_API int swrb_fb_get_pixels_device_on_stream(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride, void* cuda_stream);

/* Scissor rows (no reference counterpart; the sort-first split of ONE view over several GPUs): from now on only
 * pixel rows [y0, y1) of the framebuffer have to come out right. Draws drop the meshlets and triangles that cannot touch
 * them (conservatively: bound sphere against the band's two planes, then the triangle's pixel box), the resolve pass shades
 * those rows only, and the GetPixels family (host, device, on-stream, swrb_fb_send_pixels) moves those rows only — to rows
 * [y0, y1) of the destination image, so N GPUs with disjoint bands fill one image without a depth compare. Inside the band
 * depth, ids and colour are bit-identical to the unscissored frame; rows outside it are unspecified; the perf counters
 * count the band's work (a triangle that straddles two bands counts in both). y0 must be a multiple of 8 and y1 a multiple
 * of 8 or the height; (0, height) or (0, 0) removes the scissor. Set it before the frame's swrb_fb_clear. */
SWRB_API int swrb_fb_set_scissor_rows(swrb_fb* fb, uint32_t y0, uint32_t y1);
SWRB_API int swrb_fb_get_scissor_rows(swrb_fb* fb, uint32_t* y0_out, uint32_t* y1_out);

/* Multi-GPU composite exchange (no reference counterpart: the reference is one process). GetPixels whose
 * destination is another GPU's memory (NVLink peer mapping) with the flow control folded into the same kernel:
 * it first waits until *WaitFlag >= WaitValue (a flag in THIS GPU's memory that the consumer raises when the
 * destination slot may be overwritten; NULL = no wait) and, after its last store, sets *SignalFlag = SignalValue
 * (a flag in the CONSUMER's memory; NULL = no signal) behind a system-scope fence. Runs on `cuda_stream`. */
typedef struct swrb_peer_sync {
    const uint64_t* WaitFlag;   uint64_t WaitValue;
    uint64_t*       SignalFlag; uint64_t SignalValue;
} swrb_peer_sync;
SWRB_API int swrb_fb_send_pixels(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride, void* cuda_stream, const swrb_peer_sync* sync);
/* Consumer side: one small kernel on `cuda_stream` that waits until ready_flags[0..n) (this GPU's memory) are all
 * >= expected, then stores ack_value to each of ack_flags[0..n) (the producers' memories). n <= 15. */
SWRB_API int swrb_peer_collect(swrb_device* dev, void* cuda_stream, const uint64_t* ready_flags, uint32_t n, uint64_t expected,
                               uint64_t* const* ack_flags, uint64_t ack_value);

/* ---- hot path ------------------------------------------------------------------------ */
/* ShadingContext::CullMeshlets frustum part (Shading.cpp:775-809, :865-867). Planes are
 * derived on the host from P,V,M exactly as the reference does (the HiZ half is
 * swrb_cull_meshlets_hiz below). Writes 1 bit/meshlet to bitmap_out_host (if non-NULL, ceil(count/16)
 * uint16) and keeps a device copy for UseDeviceCullBitmap. Returns the visible count. */
SWRB_API int swrb_cull_meshlets(swrb_scene* scene, uint32_t meshlet_offset, uint32_t count,
                                const float proj[16], const float view[16], const float model[16],
                                uint16_t* bitmap_out_host, uint32_t* visible_out);
/* HiZ occlusion culling (the reference's second half of CullMeshlets):
 *   swrb_hiz_create   the R32f TiledY8 depth pyramid the Playground allocates, CreateTexture2D<R32f>(halfW, halfH, 16)
 *                     with halfW/halfH from Main.cpp:54-56
 *   swrb_hiz_build    texutil::DownsampleDepth(fb, depthMap)  (ImageHelpers.cpp:150-247)
 *   swrb_cull_meshlets_hiz   ShadingContext::CullMeshlets(bitmap, meshlets, count, P, V, M, prevV, frameSize, depthMap)
 *                     (Shading.cpp:775-869) — hiz == NULL gives the frustum-only result of swrb_cull_meshlets. */
typedef struct swrb_hiz swrb_hiz;
SWRB_API int swrb_hiz_create(swrb_device* dev, uint32_t fb_width, uint32_t fb_height, swrb_hiz** out);
SWRB_API void swrb_hiz_destroy(swrb_hiz* hiz);
SWRB_API int swrb_hiz_info(const swrb_hiz* hiz, swr_texture_desc* out);    /* layout (Data = NULL) */
SWRB_API int swrb_hiz_build(swrb_hiz* hiz, swrb_fb* fb);
SWRB_API int swrb_hiz_download(swrb_hiz* hiz, float* dst_host);            /* LayerStride floats, raw TiledY8 */
SWRB_API int swrb_cull_meshlets_hiz(swrb_scene* scene, uint32_t meshlet_offset, uint32_t count,
                                    const float proj[16], const float view[16], const float model[16], const float prev_view[16],
                                    float frame_w, float frame_h, swrb_hiz* hiz,
                                    uint16_t* bitmap_out_host, uint32_t* visible_out);
/* The five normalised Gribb-Hartmann planes CullMeshlets tests (Shading.cpp:783-791, :806). */
SWRB_API int swrb_frustum_planes(const float proj[16], const float view[16], const float model[16], float planes_out[5][4]);

/* Rasterizer::DrawMeshlets with the VisBufferShader table (mesh program = ShadeMeshlet
 * Shading.cpp:281-307; fragment programs = FS_EncodeSurfaceId<false/true> :309-331). */
SWRB_API int swrb_draw_meshlets(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draw);
/* Several DrawMeshlets calls (one per glTF node, Main.cpp:216-240) submitted as one batch.
 * Results equal issuing them one after another when MeshletOffset is non-decreasing. */
SWRB_API int swrb_draw_batch(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws);
/* Literal drop-in form: ctx.Meshlets is a HOST pointer, uploaded on every call. */
SWRB_API int swrb_draw_meshlets_host(swrb_fb* fb, const swr_meshlet* meshlets_host, uint32_t count,
                                     const float object_to_clip[16], const uint16_t* cull_bitmap_host);

/* ShadingContext::Resolve (Shading.cpp:658-689): overwrites layer 0 with RGBA8 colour. */
SWRB_API int swrb_resolve(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* uniforms);

/* ---- prepared draws and whole frames ---------------------------------------------------------
 * The reference's frame loop (Main.cpp:213-252, RasterBench.cpp:92-106) rebuilds the per-node uniforms and calls
 * DrawMeshlets once per glTF node every frame. A swrb_batch is that list of DrawMeshlets calls validated, laid out and
 * uploaded ONCE (matrices, frustum planes, copies of host cull bitmaps): drawing it again costs no host staging and no
 * H2D copy. It captures SWRB_FLAG_FUSED_FRUSTUM_CULL as set at creation. Results equal swrb_draw_batch(draws). */
typedef struct swrb_batch swrb_batch;
SWRB_API int swrb_batch_create(swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws, swrb_batch** out);
SWRB_API void swrb_batch_destroy(swrb_batch* batch);
SWRB_API int swrb_draw_prepared(swrb_fb* fb, const swrb_batch* batch, uint32_t program);
/* One iteration of that loop in ONE call: Framebuffer::Clear -> DrawMeshlets per node (the batch) -> [Resolve] ->
 * [GetPixels(layer 0)]. Everything is enqueued on the device's stream except the optional GetPixels, which may run on
 * `PixelsStream` (NULL = the device stream): the library orders it after the resolve pass and orders the next writer of
 * layer 0 after it, so callers pipeline frames without any event code of their own. With PeerSync the copy is
 * swrb_fb_send_pixels (destination in another GPU's memory, flow control folded in). */
typedef struct swrb_frame_desc {
    uint32_t ClearColor; float ClearDepth;          /* Framebuffer::Clear(color, depth), Main.cpp:213 */
    const swrb_batch* Batch;                        /* the frame's DrawMeshlets calls */
    const swrb_shading_uniforms* Uniforms;          /* ShadingContext::Resolve; NULL = vis-buffer only */
    void* PixelsDevice; uint32_t PixelsStride;      /* GetPixels(0) into device / peer memory; NULL = none; stride 0 = width */
    void* PixelsStream;                             /* cudaStream_t of that copy, NULL = device stream */
    const struct swrb_peer_sync* PeerSync;          /* NULL = plain copy */
    uint32_t* PixelsHost;                           /* GetPixels(0) into (pinned) host memory, async; NULL = none */
} swrb_frame_desc;
SWRB_API int swrb_frame_submit(swrb_fb* fb, const swrb_frame_desc* frame);

/* The shader tables a DrawMeshlets call can be bound to (ShadingContext::VisBufferShader / DeferredShader /
 * OverdrawShader, Shading.h:49, Shading.cpp:648-656; the Playground picks one per frame, Main.cpp:204-209). */
typedef enum swrb_program {
    SWRB_PROGRAM_VISBUFFER = 0,   /* ShadeMeshlet + FS_EncodeSurfaceId<false/true> */
    SWRB_PROGRAM_OVERDRAW  = 1,   /* ShadeMeshlet + FS_Overdraw (Shading.cpp:333-342) in every fragment slot: layer 0 counts
                                     covered pixels (high u16) and helper lanes of touched 4x4 fragments (low u16), saturating;
                                     layer 1 keeps max(depth); no depth test */
    SWRB_PROGRAM_DEFERRED  = 2    /* ShadeMeshlet + FS_EncodeGBuffer (Shading.cpp:344-414) in every fragment slot: depth test,
                                     alpha test on the sampled base colour, layer 0 = base colour, layer 1 = depth, layer 2 =
                                     octahedron-packed world normal (2 x 10 bit) + metallic / roughness (2 x 6 bit). Needs a
                                     framebuffer with >= 3 layers; the per-draw ObjectToWorld is taken as the upper 3x3 of
                                     swrb_draw_desc.ObjectToWorld. */
} swrb_program;
/* Rasterizer::DrawMeshlets(fb, count, {table, &ctx}) with the table chosen by `program` (Main.cpp:204-209, :236-240). */
SWRB_API int swrb_draw_batch_program(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws, uint32_t program);

/* enum class DebugLayer (Shading.h:8) */
typedef enum swrb_debug_layer {
    SWRB_LAYER_NONE = 0, SWRB_LAYER_BASE_COLOR, SWRB_LAYER_NORMALS, SWRB_LAYER_METALLIC_ROUGHNESS, SWRB_LAYER_MESHLET_ID,
    SWRB_LAYER_TRIANGLE_ID, SWRB_LAYER_OVERDRAW_PIXEL, SWRB_LAYER_OVERDRAW_QUAD
} swrb_debug_layer;
/* ShadingContext::ResolveDebug (Shading.cpp:734-773): overwrites layer 0 with the visualisation of `layer`
 * (the overdraw layers expect a framebuffer drawn with SWRB_PROGRAM_OVERDRAW). SWRB_LAYER_NONE is invalid here
 * (the Playground calls Resolve instead, Main.cpp:248-252). */
SWRB_API int swrb_resolve_debug(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* uniforms, uint32_t layer);

/* ---- pinned host staging (cudaMallocHost) for callers that want full-rate PCIe copies ------ */
SWRB_API int swrb_alloc_pinned(swrb_device* dev, uint64_t bytes, void** out);
SWRB_API int swrb_free_pinned(swrb_device* dev, void* ptr);

/* ---- timing helpers (CUDA events on the device's stream; used by bench.py) ------------- */
SWRB_API int swrb_timer_begin(swrb_device* dev);
SWRB_API int swrb_timer_end(swrb_device* dev, float* elapsed_ms);    /* synchronises */
SWRB_API int swrb_flush_l2(swrb_device* dev);                          /* writes a >L2-sized scratch buffer */
/* Per-stage device time of the last frame in microseconds (stage ids below); needs
 * swrb_device_enable_stage_timing(dev,1), which inserts events around every kernel. */
enum { SWRB_STAGE_CLEAR = 0, SWRB_STAGE_CULL, SWRB_STAGE_MESH, SWRB_STAGE_BIN, SWRB_STAGE_RASTER,
       SWRB_STAGE_RESOLVE, SWRB_STAGE_COUNT_ };
SWRB_API int swrb_device_enable_stage_timing(swrb_device* dev, int enable);
SWRB_API int swrb_get_stage_times(swrb_device* dev, float out_us[SWRB_STAGE_COUNT_], uint32_t launches_out[SWRB_STAGE_COUNT_]);
SWRB_API int swrb_get_launch_count(swrb_device* dev, uint64_t* out);  /* kernels launched since create */
/* Work-list sizes of the last draw (synchronises): out = { triangle records written (triangles too large for
 * the mesh kernel's inline raster), big-list entries, tile-list entries, reserved }. */
SWRB_API int swrb_get_draw_stats(swrb_device* dev, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* SWRB_H */
