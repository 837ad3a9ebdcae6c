/* swr_types.h — plain-old-data layouts at the drop-in boundary.
 *
 * These mirror, byte for byte, the structs the reference's meshlet raster path
 * exchanges with its callers. Layout only: no algorithms live here, so the header
 * is shared by the CUDA product (glimpsw_b200/csrc), the C++ shim
 * (include/swr_b200.hpp) and the CPU oracle (oracle/).
 *
 * Reference citations are relative to /root/reference/.
 */
#ifndef SWR_TYPES_H
#define SWR_TYPES_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SWR_MAX_VERTICES = 64, SWR_MAX_PRIMS = 128 };  /* Rasterizer.h:83 */
enum { SWR_MAX_RENDER_SIZE = 2896 };                   /* Rasterizer.h:203 */
#define SWR_NO_MATERIAL 0xFFFFFFFFu                    /* Shading.cpp:302 (UINT_MAX) */

/* struct Meshlet — src/SwRast/Scene.h:15-30. 1728 bytes, 64-byte aligned.
 * Positions are SoA [xyz][vertex]; Indices are SoA [corner][triangle]. */
typedef struct swr_meshlet {
    float    BoundCenter[3];
    float    BoundRadius;
    float    ConeApex[3];
    float    ConeAxis[3];
    float    ConeCutoff;
    uint8_t  NumVertices, NumTriangles;
    uint8_t  AlphaCutoff;
    uint8_t  _pad0;
    uint32_t MaterialId;
    uint32_t _pad1;
    uint64_t TangentHandedness;
    float    Positions[3][SWR_MAX_VERTICES];
    uint32_t TexCoords[SWR_MAX_VERTICES];       /* 2 x fp16  */
    uint32_t NormalTangents[SWR_MAX_VERTICES];  /* 2 x oct16 (4 x unorm8) */
    uint8_t  Indices[3][SWR_MAX_PRIMS];
} swr_meshlet;

/* Packed meshlet — the import-time / transport format of this library for the compression the reference's author
 * planned (src/SwRast/Shading.cpp:292-294 "TODO: meshlet compression"). Everything but the positions is the Meshlet's bytes
 * verbatim; the positions are 16-bit fixed point inside the meshlet's own bounding box:
 *     Positions[a][v] = fmaf((float)Q[a][v], Scale[a], Origin[a])          (one fused multiply-add, IEEE binary32)
 * 1376 bytes instead of 1728 (positions 384 instead of 768). swrb_scene_create_packed / swrb_scene_update_packed decode it
 * on the device into the reference layout at upload, so every kernel keeps reading swr_meshlet. */
typedef struct swr_meshlet_packed {
    uint8_t  Header[64];                        /* swr_meshlet bytes 0..63: bounds, cone, counts, MaterialId, TangentHandedness */
    float    Origin[3];
    float    Scale[3];
    uint32_t _pad[2];
    uint16_t Q[3][SWR_MAX_VERTICES];
    uint32_t TexCoords[SWR_MAX_VERTICES];
    uint32_t NormalTangents[SWR_MAX_VERTICES];
    uint8_t  Indices[3][SWR_MAX_PRIMS];
} swr_meshlet_packed;

/* struct ShadedMeshlet — src/SwRast/Rasterizer.h:82-99 (mesh-shader output). 1472 bytes. */
typedef struct swr_shaded_meshlet {
    uint8_t PrimCount;
    uint8_t CullMode;          /* swr_cull_mode */
    uint8_t FragmentShaderId;
    uint8_t _pad[61];
    uint8_t Indices[3][SWR_MAX_PRIMS];
    float   Position[4][SWR_MAX_VERTICES];
} swr_shaded_meshlet;

typedef enum swr_cull_mode { SWR_CULL_NONE = 0, SWR_CULL_FRONT_CCW = 1, SWR_CULL_FRONT_CW = 2 } swr_cull_mode; /* Rasterizer.h:80 */

/* struct Material — src/SwRast/Scene.h:7-14. The host pointer becomes a texture index. */
typedef struct swr_material {
    int32_t TextureId;         /* index into the scene's texture table, -1 = none */
    uint8_t IsDoubleSided;
    uint8_t AlphaCutoff;
    uint8_t _pad[2];
} swr_material;

/* struct Light — src/SwRast/Scene.h:52-75 (same field order, 68 bytes). */
typedef struct swr_light {
    uint32_t Type;             /* 0 directional, 1 point, 2 spot */
    float Position[3];
    float Direction[3];
    float Color[3];
    float Intensity;
    float Radius;
    float SpotInnerAngle, SpotOuterAngle;
    float InvRadiusSq, SpotScale, SpotOffset;
} swr_light;

/* Texture2D<RGBA8u, TiledY8> — src/SwRast/Texture.h:314-329 + CreateTexture2D :600-636.
 * `Data` is the reference's texel array verbatim (all layers, all mips, TiledY8-swizzled). */
typedef struct swr_texture_desc {
    uint32_t Width, Height, MipLevels, NumLayers;
    uint32_t RowShift, LayerStride;           /* in texels */
    uint32_t MipOffsets[16];                  /* in texels */
    const uint32_t* Data;                     /* LayerStride * NumLayers texels */
} swr_texture_desc;

/* Framebuffer header — src/SwRast/Rasterizer.h:10-15. Data is 4x4-tiled u32 per layer:
 * offset(x,y) = ((x&~3)<<2) + (y&~3)*Width + (x&3) + (y&3)*4   (Rasterizer.h:50-56) */
typedef struct swr_fb_info {
    uint32_t Width, Height, TileStride, LayerStride, NumLayers;
} swr_fb_info;

/* PerfCounter — src/SwRast/Rasterizer.h:366-379 */
typedef enum swr_perf_counter {
    SWR_PERF_TrianglesProcessed = 0,
    SWR_PERF_TrianglesRasterized,
    SWR_PERF_TrianglesClipped,
    SWR_PERF_BinQueueFlushes,
    SWR_PERF_DrawTime,
    SWR_PERF_ResolveTime,
    SWR_PERF_ShadowTime,
    SWR_PERF_FrameTime,
    SWR_PERF_Count_
} swr_perf_counter;

#ifdef __cplusplus
}
static_assert(sizeof(swr_meshlet) == 1728, "Meshlet layout (Scene.h:15-30)");
static_assert(offsetof(swr_meshlet, NumVertices) == 44, "");
static_assert(offsetof(swr_meshlet, MaterialId) == 48, "");
static_assert(offsetof(swr_meshlet, TangentHandedness) == 56, "");
static_assert(offsetof(swr_meshlet, Positions) == 64, "");
static_assert(offsetof(swr_meshlet, TexCoords) == 832, "");
static_assert(offsetof(swr_meshlet, NormalTangents) == 1088, "");
static_assert(offsetof(swr_meshlet, Indices) == 1344, "");
static_assert(sizeof(swr_meshlet_packed) == 1376, "packed meshlet layout");
static_assert(offsetof(swr_meshlet_packed, Q) == 96 && offsetof(swr_meshlet_packed, TexCoords) == 480 && offsetof(swr_meshlet_packed, Indices) == 992, "");
static_assert(sizeof(swr_shaded_meshlet) == 1472, "ShadedMeshlet layout (Rasterizer.h:82-99)");
static_assert(offsetof(swr_shaded_meshlet, Indices) == 64, "");
static_assert(offsetof(swr_shaded_meshlet, Position) == 448, "");
static_assert(sizeof(swr_light) == 68, "Light layout (Scene.h:52-75)");
#endif

#endif /* SWR_TYPES_H */
