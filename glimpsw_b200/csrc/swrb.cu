// swrb.cu — host side of libswrb.so: the C ABI declared in include/swrb.h.
//
// One swrb_device owns a CUDA stream and the transient work buffers of the pipeline
//   [cull] -> K1 mesh/setup (+tile count) -> K2 scan -> K2 scatter -> K3 tile raster      (binned)
//   [cull] -> K1 mesh/setup -> key init -> direct raster (+big) -> key unpack             (unbinned)
//   -> K5 resolve
// Every call only enqueues work; nothing here computes pixels on the CPU.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "cull.cuh"
#include "hiz.cuh"
#include "bin.cuh"
#include "mesh.cuh"
#include "tile.cuh"
#include "raster_direct.cuh"
#include "fbops.cuh"
#include "resolve.cuh"
#include "alpha.cuh"
#include "clip.cuh"
#include "overdraw.cuh"
#include "gbuffer.cuh"
#include "unpack.cuh"

using namespace swrb;

// ---------------------------------------------------------------------------------------------
static thread_local std::string g_lastError;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return code;
}
#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(_e == cudaErrorMemoryAllocation ? SWRB_E_OOM : SWRB_E_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                               \
    } while (0)

struct StageTimer {
    cudaEvent_t begin[SWRB_STAGE_COUNT_][8], end[SWRB_STAGE_COUNT_][8];
    uint32_t used[SWRB_STAGE_COUNT_];
};

struct swrb_device {
    int cudaDevice = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    uint32_t flags = SWRB_FLAGS_DEFAULT;

    DevCtl* ctl = nullptr;            // device
    DevCtl* ctlHost = nullptr;        // pinned mirror

    TriRecord* tris = nullptr;
    TriRecord* alphaTris = nullptr;   // alpha-tested triangles (only allocated for scenes with AlphaCutoff < 255 materials)
    TriRecordW* trisW = nullptr;      // their 1/w
    float4* clipRemap = nullptr;      // ClippedU/ClippedV of clipped alpha-tested pieces (2 float4 per alpha record)
    uint32_t* superEntries = nullptr; // binned: per-super-tile lists of wide triangles (capacity superCap)
    uint64_t superCap = 0;
    uint32_t inlineMaxArea = 128;     // FrameParams::inlineMaxArea; measured on B200 (profiles/r02_summary.md): 64 / 128 / 256 px give 57.3 / 59.5 / 60.4 Gtri/s on the
                                      // 64-view C4 batch, 175 / - / 190 us on the Sponza frame, 121 / - / 124 us on the knot (SWRB_INLINE_AREA overrides, for experiments)
    BigItem* bigItems = nullptr;      // direct: (tri, bin) work items
    uint64_t triCap = 0, bigItemCap = 0;
    uint32_t* binEntries = nullptr;
    uint64_t binCap = 0;
    uint64_t reserveTris = 0, reserveBins = 0;
    uint64_t growTris = 0, growBins = 0;    // capacities the next draw must have at least (doubled after an overflow)

    uint32_t* tileCount = nullptr;    // [numTiles] + offsets [numTiles+1] + cursors [numTiles] + active list [numTiles] + 3 x 160 super-tile words
    uint32_t* tileOffset = nullptr;
    uint32_t* tileCursor = nullptr;
    uint32_t* activeTiles = nullptr;
    uint32_t* superCount = nullptr;
    uint32_t* superOffset = nullptr;
    uint32_t* superCursor = nullptr;
    uint32_t tileCap = 0;
    bool workClean = false;           // the draw's transient device state (counters, cursors) is already reset (the last resolve pass did it)

    DrawItem* drawItems = nullptr;    // device
    uint32_t drawItemCap = 0;
    static constexpr int kStagingSlots = 4;
    DrawItem* drawStaging[kStagingSlots] = {};
    uint32_t drawStagingCap[kStagingSlots] = {};
    cudaEvent_t drawStagingDone[kStagingSlots] = {};
    int drawStagingNext = 0;

    uint16_t* cullBitmapDev = nullptr;   // result of the last swrb_cull_meshlets
    uint32_t cullBitmapCap = 0;          // in meshlets
    uint16_t* cullUpload = nullptr;      // device copy of a host-provided bitmap
    uint32_t cullUploadCap = 0;
    uint32_t* visibleDev = nullptr;

    swr_meshlet* hostDrawMeshlets = nullptr;   // scratch scene for swrb_draw_meshlets_host
    uint32_t hostDrawCap = 0;

    // GetPixels to host memory: the de-tile kernel runs on the device stream into one of kDetileSlots row-major scratch images,
    // the PCIe copy on a copy stream of its own, so the next frame renders while this one's pixels travel.
    static const int kDetileSlots = 3;
    uint32_t* detileScratch[kDetileSlots] = {};
    size_t detileCap = 0;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t detiled[kDetileSlots] = {}, copied[kDetileSlots] = {};
    bool copyInFlight[kDetileSlots] = {};
    int detileSlot = 0;
    bool hostCopiesPending = false;   // swrb_sync also waits for the copy stream

    void* l2Scratch = nullptr;
    size_t l2ScratchBytes = 0;

    cudaEvent_t timerBegin = nullptr, timerEnd = nullptr;
    bool stageTiming = false;
    StageTimer* st = nullptr;
    uint64_t launches = 0;
    uint64_t hostTimeNs[SWR_PERF_Count_] = {};
    uint32_t lastTriCount = 0;            // records written by the most recent draw the host has read back
    // Framebuffers drawn into since the host last saw the overflow flag clear, each with the clear it had pending before its
    // first such draw: all of them are rolled back when a draw of that window turns out to have aborted.
    std::vector<std::pair<swrb_fb*, bool>> drawnSinceCheck;

    // Per-vertex {x/w, y/w, 1/w, z/w} of the last batch, written by the mesh kernel for the resolve pass.
    // swrb_resolve may read it only if every surface id in the framebuffer provably comes from that batch
    // and the batch used the one matrix the resolve is handed (see clip_cache_usable).
    uint32_t meshBlocksPerSM = 4;         // persistent grid of the mesh kernel (swrb_device_set_mesh_occupancy)
    uint32_t* peerCounter = nullptr;      // ring of block counters of k_fb_detile_send (one per in-flight send)
    uint32_t peerCounterNext = 0;
    static constexpr uint32_t kPeerCounters = 64;
    float4* clipCache = nullptr;
    uint64_t clipCacheCap = 0;            // in meshlets
    swrb_fb* clipCacheFb = nullptr;       // framebuffer the last batch drew into (null = cache unusable)
    const swr_meshlet* clipCacheMeshlets = nullptr;
    bool clipCacheUniform = false;        // all draws of that batch shared one ObjectToClip
    float clipCacheM[16] = {};
};

struct DeviceTexture {     // Texture2D<RGBA8u, TiledY8> (Texture.h:314-329)
    swr_texture_desc desc;
    uint32_t* data;
};

struct swrb_scene {
    swrb_device* dev;
    swr_meshlet* meshlets = nullptr;
    uint32_t numMeshlets = 0;
    swr_material* materials = nullptr;
    uint32_t numMaterials = 0;
    ResolveTexture* textures = nullptr;   // device table
    std::vector<uint32_t*> textureData;
    uint32_t numTextures = 0;
    ResolveTexture sky = {};              // ShadingContext::SkyboxTex (swrb_scene_set_skybox); skyData == nullptr: none
    uint32_t* skyData = nullptr;
    float4* attr = nullptr;               // resolve-pass attribute table (k_decode_attributes), 2 float4 per vertex
    uint32_t attrDirtyLo = 0, attrDirtyHi = 0;   // meshlet range whose attributes must be (re)decoded before the next resolve
    swr_light* lights = nullptr;
    std::vector<swr_light> lightsHost;    // for the light markers of Resolve (projected on the host)
    uint32_t numLights = 0;
    bool hasAlphaTest = false;
    // Host copy of "MaterialId == UINT_MAX" per meshlet: DeferredShader treats such meshlets differently (depth only,
    // Shading.cpp:352-355) and a batch that mixes both kinds is drawn as runs of one kind (draw_deferred).
    swr_meshlet_packed* packedStaging = nullptr;   // device buffer the packed bytes land in before k_unpack_meshlets (swrb_scene_*_packed)
    uint32_t packedStagingCap = 0;
    std::vector<uint8_t> materialless;
    std::vector<int32_t> materialTextureIds;   // host copy of Material::TextureId
    uint32_t numMaterialless = 0;
};

struct swrb_fb {
    swrb_device* dev;
    uint32_t width, height, layers, layerStride;
    uint32_t* data = nullptr;
    unsigned long long* keys = nullptr;   // direct path only (lazily allocated)
    bool pendingClear = false;            // Clear() recorded but not yet materialised (fused into the next draw)
    // Lazy vis-buffer: after a draw the result lives in `keys`; the depth / id layers are produced on demand.
    bool visInKeys = false;               // keys hold the latest vis-buffer, layers are stale for pixels the draw won
    bool keysClearMode = false;           // that draw started from a logically cleared framebuffer (seed = clear value)
    uint32_t keysClearColor = 0;
    bool layer0IsColor = false;           // resolve already consumed the keys and wrote colour to layer 0
    uint32_t clearColor = 0, clearDepthBits = 0;
    bool keysSeeded = false;              // the last resolve pass left the next frame's seeds in `keys` (depth = seedDepthBits)
    uint32_t seedDepthBits = 0;
    // GetPixels on a side stream: the copy is ordered after the device stream's work (evReady) and everything that
    // overwrites the layer it reads is ordered after the copy (evCopied).
    cudaEvent_t evReady = nullptr, evCopied = nullptr;
    bool copyPending = false;
    // scissor rows (swrb_fb_set_scissor_rows): draws, the resolve pass and the GetPixels family only deal with rows [bandY0, bandY1)
    uint32_t bandY0 = 0, bandY1 = 0;      // bandY1 == 0: the whole framebuffer
    uint32_t row0() const { return bandY1 ? bandY0 : 0u; }
    uint32_t row1() const { return bandY1 ? bandY1 : height; }
};

// An aborted draw (device work list overflow) never touched the depth / id layers; every kernel after it
// was predicated off by the sticky device flag. Drop its keys and restore the recorded clear, if any.
static void rollback_aborted_draw(swrb_device* d) {
    d->clipCacheFb = nullptr;
    d->workClean = false;
    for (auto& e : d->drawnSinceCheck) {
        e.first->visInKeys = false;
        e.first->layer0IsColor = false;
        e.first->keysSeeded = false;
        e.first->pendingClear = e.second;
    }
    d->drawnSinceCheck.clear();
}
static void note_draw_target(swrb_device* d, swrb_fb* fb, bool pendingClearBefore) {
    for (auto& e : d->drawnSinceCheck) if (e.first == fb) return;       // keep the state before the FIRST draw of the window
    d->drawnSinceCheck.emplace_back(fb, pendingClearBefore);
}
static void forget_draw_target(swrb_device* d, swrb_fb* fb) {
    for (size_t i = 0; i < d->drawnSinceCheck.size(); i++)
        if (d->drawnSinceCheck[i].first == fb) { d->drawnSinceCheck.erase(d->drawnSinceCheck.begin() + i); return; }
}

// NVTX ranges around the stages (the reference marks them with Tracy zones, Rasterizer.cpp:494,518,599,612).
// libnvToolsExt is only looked for when SWRB_NVTX is set in the environment; without it a range costs one branch.
struct nvtx_range {
    typedef int (*push_fn)(const char*);
    typedef int (*pop_fn)(void);
    static void resolve(push_fn& push, pop_fn& pop) {
        static push_fn s_push = nullptr; static pop_fn s_pop = nullptr; static bool tried = false;
        if (!tried) {
            tried = true;
            if (getenv("SWRB_NVTX")) {
                void* h = dlopen("libnvToolsExt.so.1", RTLD_NOW | RTLD_GLOBAL);
                if (!h) h = dlopen("libnvToolsExt.so", RTLD_NOW | RTLD_GLOBAL);
                if (h) { s_push = (push_fn)dlsym(h, "nvtxRangePushA"); s_pop = (pop_fn)dlsym(h, "nvtxRangePop"); }
            }
        }
        push = s_push; pop = s_pop;
    }
    pop_fn popFn = nullptr;
    explicit nvtx_range(const char* name) {
        push_fn push; pop_fn pop;
        resolve(push, pop);
        if (push && pop) { push(name); popFn = pop; }
    }
    ~nvtx_range() { if (popFn) popFn(); }
};

// ---------------------------------------------------------------------------------------------
struct StageScope {
    swrb_device* d; int stage; int slot = -1;
    StageScope(swrb_device* dev, int s) : d(dev), stage(s) {
        if (d->stageTiming && d->st->used[s] < 8) {
            slot = (int)d->st->used[s]++;
            cudaEventRecord(d->st->begin[s][slot], d->stream);
        }
    }
    ~StageScope() {
        if (slot >= 0) cudaEventRecord(d->st->end[stage][slot], d->stream);
    }
};

static inline uint32_t grid_for(const swrb_device* d, uint64_t items, uint32_t perBlock, uint32_t blocksPerSM) {
    uint64_t need = (items + perBlock - 1) / perBlock;
    uint64_t cap = (uint64_t)d->numSMs * blocksPerSM;
    return (uint32_t)std::max<uint64_t>(1, std::min(need, cap));
}

// Grid of a kernel whose blocks split a list by a fixed stride: exactly the blocks the GPU keeps resident (registers / shared
// memory decide how many per SM). A larger grid would run in waves — the first wave's blocks finish their 1/grid of the list,
// then the rest runs at a fraction of the occupancy.
template <typename K>
static uint32_t resident_grid(const swrb_device* d, K kernel, int blockThreads, uint32_t maxPerSM = 8) {
    int perSM = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, blockThreads, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
    return (uint32_t)d->numSMs * std::min<uint32_t>((uint32_t)perSM, maxPerSM);
}

static int ensure_buffer(void** ptr, uint64_t* cap, uint64_t need, size_t elem) {
    if (need <= *cap && *ptr != nullptr) return SWRB_OK;
    if (*ptr) CU(cudaFree(*ptr));
    *ptr = nullptr; *cap = 0;
    CU(cudaMalloc(ptr, need * elem));
    *cap = need;
    return SWRB_OK;
}

static void rollback_aborted_draw(swrb_device* d);

static int check_overflow(swrb_device* d) {
    // caller has synchronised the stream
    CU(cudaMemcpy(d->ctlHost, d->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost));
    d->lastTriCount = d->workClean ? d->ctlHost->lastTriCount : d->ctlHost->triCount;
    if (!d->ctlHost->overflow) d->drawnSinceCheck.clear();
    if (d->ctlHost->overflow) {
        uint32_t which = d->ctlHost->overflow;
        uint32_t zero = 0;
        cudaMemcpy(&d->ctl->overflow, &zero, 4, cudaMemcpyHostToDevice);
        rollback_aborted_draw(d);
        if (which == 1) d->growTris = std::max<uint64_t>(d->growTris, 2 * d->triCap);
        else d->growBins = std::max<uint64_t>(d->growBins, 2 * d->binCap);
        return fail(SWRB_E_BIN_OVERFLOW,
                    "device work list overflowed (%s); every draw since the last synchronising call was aborted before touching "
                    "its framebuffer — redraw: the lists are twice as large now (or call swrb_device_reserve() with your own limits)",
                    which == 1 ? "triangle records" : which == 2 ? "big-triangle work items" : "tile / super-tile bin entries");
    }
    return SWRB_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" {

static int fb_materialize_for_read(swrb_fb* fb, uint32_t layer);
void swrb_device_destroy(swrb_device* d);
void swrb_scene_destroy(swrb_scene* s);
void swrb_fb_destroy(swrb_fb* fb);
void swrb_hiz_destroy(swrb_hiz* z);
}

// A half-built object is torn down by its own destroy function when a create call fails part-way
// (every CU(...) inside returns early): the destroy functions tolerate null members.
template <class T, void (*Destroy)(T*)>
struct CreateGuard {
    T* p;
    ~CreateGuard() { if (p) Destroy(p); }
    T* release() { T* r = p; p = nullptr; return r; }
};

extern "C" {

const char* swrb_last_error(void) { return g_lastError.c_str(); }
const char* swrb_version(void) { return "swrb 0.1 (sm_100a)"; }

int swrb_device_create(int cuda_device, swrb_device** out) {
    if (!out) return fail(SWRB_E_INVALID, "out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SWRB_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cuda_device < 0 || cuda_device >= n) return fail(SWRB_E_INVALID, "cuda_device %d out of range (%d devices)", cuda_device, n);
    CU(cudaSetDevice(cuda_device));
    CreateGuard<swrb_device, swrb_device_destroy> guard{ new swrb_device() };
    swrb_device* d = guard.p;
    d->cudaDevice = cuda_device;
    CU(cudaDeviceGetAttribute(&d->numSMs, cudaDevAttrMultiProcessorCount, cuda_device));
    if (const char* e = getenv("SWRB_INLINE_AREA")) d->inlineMaxArea = (uint32_t)std::min(std::max(atoi(e), 1), kInlineMaxArea);
    CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&d->ctl, sizeof(DevCtl)));
    CU(cudaMemset(d->ctl, 0, sizeof(DevCtl)));
    CU(cudaMallocHost(&d->ctlHost, sizeof(DevCtl)));
    CU(cudaMalloc(&d->visibleDev, 4));
    CU(cudaEventCreate(&d->timerBegin));
    CU(cudaEventCreate(&d->timerEnd));
    for (int i = 0; i < swrb_device::kStagingSlots; i++) CU(cudaEventCreateWithFlags(&d->drawStagingDone[i], cudaEventDisableTiming));
    *out = guard.release();
    return SWRB_OK;
}

void swrb_device_destroy(swrb_device* d) {
    if (!d) return;
    cudaSetDevice(d->cudaDevice);
    cudaStreamSynchronize(d->stream);
    cudaFree(d->ctl); cudaFreeHost(d->ctlHost); cudaFree(d->tris); cudaFree(d->trisW); cudaFree(d->clipRemap); cudaFree(d->alphaTris); cudaFree(d->superEntries);
    cudaFree(d->bigItems); cudaFree(d->binEntries); cudaFree(d->tileCount); cudaFree(d->drawItems);
    for (int i = 0; i < swrb_device::kStagingSlots; i++) { cudaFreeHost(d->drawStaging[i]); cudaEventDestroy(d->drawStagingDone[i]); }
    cudaFree(d->cullBitmapDev); cudaFree(d->cullUpload); cudaFree(d->visibleDev); cudaFree(d->hostDrawMeshlets);
    for (int i = 0; i < swrb_device::kDetileSlots; i++) { cudaFree(d->detileScratch[i]); if (d->detiled[i]) cudaEventDestroy(d->detiled[i]); if (d->copied[i]) cudaEventDestroy(d->copied[i]); }
    if (d->copyStream) { cudaStreamSynchronize(d->copyStream); cudaStreamDestroy(d->copyStream); }
    cudaFree(d->clipCache); cudaFree(d->peerCounter); cudaFree(d->l2Scratch);
    cudaEventDestroy(d->timerBegin); cudaEventDestroy(d->timerEnd);
    if (d->st) {
        for (int s = 0; s < SWRB_STAGE_COUNT_; s++)
            for (int k = 0; k < 8; k++) { cudaEventDestroy(d->st->begin[s][k]); cudaEventDestroy(d->st->end[s][k]); }
        delete d->st;
    }
    if (d->ownStream) cudaStreamDestroy(d->stream);
    delete d;
}

int swrb_device_set_stream(swrb_device* d, void* cuda_stream) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));
    if (d->ownStream) { CU(cudaStreamDestroy(d->stream)); }
    if (cuda_stream) { d->stream = (cudaStream_t)cuda_stream; d->ownStream = false; }
    else { CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking)); d->ownStream = true; }
    return SWRB_OK;
}

int swrb_device_set_flags(swrb_device* d, uint32_t flags) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    d->flags = flags;
    return SWRB_OK;
}

int swrb_device_reserve(swrb_device* d, uint64_t max_triangles, uint64_t max_bin_entries) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    d->reserveTris = max_triangles;
    d->reserveBins = max_bin_entries;
    return SWRB_OK;
}

int swrb_sync(swrb_device* d) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));
    if (d->hostCopiesPending) { CU(cudaStreamSynchronize(d->copyStream)); d->hostCopiesPending = false; }
    return check_overflow(d);
}

int swrb_get_counters(swrb_device* d, uint64_t out[SWR_PERF_Count_]) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));
    int rc = check_overflow(d);
    if (rc) return rc;
    for (int i = 0; i < 4; i++) out[i] = d->ctlHost->perf[i];
    for (int i = 4; i < SWR_PERF_Count_; i++) out[i] = d->hostTimeNs[i];
    return SWRB_OK;
}

int swrb_device_set_mesh_occupancy(swrb_device* d, uint32_t blocks_per_sm) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    if (blocks_per_sm < 1 || blocks_per_sm > 4) return fail(SWRB_E_INVALID, "blocks_per_sm = %u: must be in [1, 4]", blocks_per_sm);
    d->meshBlocksPerSM = blocks_per_sm;
    return SWRB_OK;
}

int swrb_reset_counters(swrb_device* d) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaMemsetAsync(d->ctl->perf, 0, sizeof(d->ctl->perf), d->stream));
    memset(d->hostTimeNs, 0, sizeof(d->hostTimeNs));
    return SWRB_OK;
}

// ---- scene -------------------------------------------------------------------------------------
static uint32_t packed_material_id(const swr_meshlet_packed& p) { uint32_t id; memcpy(&id, p.Header + offsetof(swr_meshlet, MaterialId), 4); return id; }

// H2D of packed meshlets into the scene's staging buffer + decode into meshlets [first, first + count) (unpack.cuh).
static int upload_packed(swrb_scene* s, const swr_meshlet_packed* packed, uint32_t first, uint32_t count) {
    if (count == 0) return SWRB_OK;
    swrb_device* d = s->dev;
    if (count > s->packedStagingCap) {
        CU(cudaStreamSynchronize(d->stream));
        if (s->packedStaging) CU(cudaFree(s->packedStaging));
        s->packedStaging = nullptr; s->packedStagingCap = 0;
        CU(cudaMalloc(&s->packedStaging, (size_t)count * sizeof(swr_meshlet_packed)));
        s->packedStagingCap = count;
    }
    CU(cudaMemcpyAsync(s->packedStaging, packed, (size_t)count * sizeof(swr_meshlet_packed), cudaMemcpyHostToDevice, d->stream));
    k_unpack_meshlets<<<std::min<uint32_t>(count, (uint32_t)d->numSMs * 16u), 128, 0, d->stream>>>(s->packedStaging, s->meshlets + first, count);
    d->launches++;
    CU(cudaGetLastError());
    return SWRB_OK;
}

static int scene_create_common(swrb_device* d, const swr_meshlet* meshlets, const swr_meshlet_packed* packed, uint32_t num_meshlets,
                               const swr_material* materials, uint32_t num_materials, const swr_texture_desc* textures, uint32_t num_textures,
                               const swr_light* lights, uint32_t num_lights, swrb_scene** out);

int swrb_scene_create(swrb_device* d, const swr_meshlet* meshlets, uint32_t num_meshlets, const swr_material* materials,
                      uint32_t num_materials, const swr_texture_desc* textures, uint32_t num_textures,
                      const swr_light* lights, uint32_t num_lights, swrb_scene** out) {
    if (num_meshlets && !meshlets) return fail(SWRB_E_INVALID, "meshlets is null");
    return scene_create_common(d, meshlets, nullptr, num_meshlets, materials, num_materials, textures, num_textures, lights, num_lights, out);
}

int swrb_scene_create_packed(swrb_device* d, const swr_meshlet_packed* meshlets, uint32_t num_meshlets, const swr_material* materials,
                             uint32_t num_materials, const swr_texture_desc* textures, uint32_t num_textures,
                             const swr_light* lights, uint32_t num_lights, swrb_scene** out) {
    if (num_meshlets && !meshlets) return fail(SWRB_E_INVALID, "meshlets is null");
    return scene_create_common(d, nullptr, meshlets, num_meshlets, materials, num_materials, textures, num_textures, lights, num_lights, out);
}

static int scene_create_common(swrb_device* d, const swr_meshlet* meshlets, const swr_meshlet_packed* packed, uint32_t num_meshlets,
                               const swr_material* materials, uint32_t num_materials, const swr_texture_desc* textures, uint32_t num_textures,
                               const swr_light* lights, uint32_t num_lights, swrb_scene** out) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    if (num_meshlets >= (1u << 24) - 1u) return fail(SWRB_E_INVALID, "%u meshlets: the visibility key orders at most 2^24 - 2 meshlets per scene", num_meshlets);
    CU(cudaSetDevice(d->cudaDevice));
    CreateGuard<swrb_scene, swrb_scene_destroy> guard{ new swrb_scene() };
    swrb_scene* s = guard.p;
    s->dev = d;
    s->numMeshlets = num_meshlets;
    s->attrDirtyLo = 0; s->attrDirtyHi = num_meshlets;
    s->materialless.resize(num_meshlets);
    for (uint32_t i = 0; i < num_meshlets; i++) {
        s->materialless[i] = (meshlets ? meshlets[i].MaterialId : packed_material_id(packed[i])) == SWR_NO_MATERIAL;
        s->numMaterialless += s->materialless[i];
    }
    if (num_meshlets) {
        CU(cudaMalloc(&s->meshlets, (size_t)num_meshlets * sizeof(swr_meshlet)));
        if (meshlets) CU(cudaMemcpyAsync(s->meshlets, meshlets, (size_t)num_meshlets * sizeof(swr_meshlet), cudaMemcpyHostToDevice, d->stream));
        else { int rc = upload_packed(s, packed, 0, num_meshlets); if (rc) return rc; }
    }
    s->numMaterials = num_materials;
    if (num_materials) {
        for (uint32_t i = 0; i < num_materials; i++) {
            s->materialTextureIds.push_back(materials[i].TextureId);
            if (materials[i].AlphaCutoff < 255) s->hasAlphaTest = true;
            if (materials[i].TextureId >= (int32_t)num_textures) return fail(SWRB_E_INVALID, "material %u references texture %d of %u", i, materials[i].TextureId, num_textures);
            // FS_EncodeSurfaceId<true> samples Material::Texture unconditionally (Shading.cpp:319-326): an alpha-tested material
            // without a texture is a null dereference upstream, an error here
            if (materials[i].AlphaCutoff < 255 && materials[i].TextureId < 0)
                return fail(SWRB_E_INVALID, "material %u is alpha-tested (AlphaCutoff %u < 255) but has no texture", i, materials[i].AlphaCutoff);
        }
        CU(cudaMalloc(&s->materials, num_materials * sizeof(swr_material)));
        CU(cudaMemcpyAsync(s->materials, materials, num_materials * sizeof(swr_material), cudaMemcpyHostToDevice, d->stream));
    }
    s->numTextures = num_textures;
    if (num_textures) {
        std::vector<ResolveTexture> table(num_textures);
        for (uint32_t i = 0; i < num_textures; i++) {
            const swr_texture_desc& t = textures[i];
            size_t texels = (size_t)t.LayerStride * t.NumLayers;
            uint32_t* dev = nullptr;
            CU(cudaMalloc(&dev, texels * 4 + 256));
            CU(cudaMemcpyAsync(dev, t.Data, texels * 4, cudaMemcpyHostToDevice, d->stream));
            s->textureData.push_back(dev);
            ResolveTexture& r = table[i];
            r.data = dev;
            r.width = t.Width; r.height = t.Height; r.mipLevels = t.MipLevels; r.numLayers = t.NumLayers;
            r.rowShift = t.RowShift; r.layerStride = t.LayerStride;
            for (int m = 0; m < 16; m++) r.mipOffsets[m] = t.MipOffsets[m];
        }
        CU(cudaMalloc(&s->textures, num_textures * sizeof(ResolveTexture)));
        CU(cudaMemcpy(s->textures, table.data(), num_textures * sizeof(ResolveTexture), cudaMemcpyHostToDevice));
    }
    s->numLights = num_lights;
    if (num_lights) {
        CU(cudaMalloc(&s->lights, num_lights * sizeof(swr_light)));
        CU(cudaMemcpyAsync(s->lights, lights, num_lights * sizeof(swr_light), cudaMemcpyHostToDevice, d->stream));
        s->lightsHost.assign(lights, lights + num_lights);
    }
    CU(cudaStreamSynchronize(d->stream));   // host inputs are only borrowed for the duration of the call
    *out = guard.release();
    return SWRB_OK;
}

int swrb_scene_update_meshlets(swrb_scene* s, const swr_meshlet* meshlets, uint32_t first, uint32_t count) {
    if (!s || !meshlets) return fail(SWRB_E_INVALID, "null argument");
    if ((uint64_t)first + count > s->numMeshlets) return fail(SWRB_E_INVALID, "range [%u,%u) exceeds %u meshlets", first, first + count, s->numMeshlets);
    CU(cudaSetDevice(s->dev->cudaDevice));
    CU(cudaMemcpyAsync(s->meshlets + first, meshlets, (size_t)count * sizeof(swr_meshlet), cudaMemcpyHostToDevice, s->dev->stream));
    for (uint32_t i = 0; i < count; i++) {
        const uint8_t ml = meshlets[i].MaterialId == SWR_NO_MATERIAL;
        s->numMaterialless += (uint32_t)ml - (uint32_t)s->materialless[first + i];
        s->materialless[first + i] = ml;
    }
    if (count) {
        if (s->attrDirtyLo >= s->attrDirtyHi) { s->attrDirtyLo = first; s->attrDirtyHi = first + count; }
        else { s->attrDirtyLo = std::min(s->attrDirtyLo, first); s->attrDirtyHi = std::max(s->attrDirtyHi, first + count); }
        if (s->dev->clipCacheMeshlets == s->meshlets) s->dev->clipCacheFb = nullptr;   // cached vertices belong to the old positions
    }
    return SWRB_OK;
}

int swrb_scene_update_packed(swrb_scene* s, const swr_meshlet_packed* meshlets, uint32_t first, uint32_t count) {
    if (!s || !meshlets) return fail(SWRB_E_INVALID, "null argument");
    if ((uint64_t)first + count > s->numMeshlets) return fail(SWRB_E_INVALID, "range [%u,%u) exceeds %u meshlets", first, first + count, s->numMeshlets);
    CU(cudaSetDevice(s->dev->cudaDevice));
    int rc = upload_packed(s, meshlets, first, count);
    if (rc) return rc;
    for (uint32_t i = 0; i < count; i++) {
        const uint8_t ml = packed_material_id(meshlets[i]) == SWR_NO_MATERIAL;
        s->numMaterialless += (uint32_t)ml - (uint32_t)s->materialless[first + i];
        s->materialless[first + i] = ml;
    }
    CU(cudaStreamSynchronize(s->dev->stream));         // the host array is only borrowed for the call; the staging buffer is reused
    return swrb_scene_touch(s, first, count);
}

int swrb_scene_download_meshlets(swrb_scene* s, swr_meshlet* dst_host, uint32_t first, uint32_t count) {
    if (!s || !dst_host) return fail(SWRB_E_INVALID, "null argument");
    if ((uint64_t)first + count > s->numMeshlets) return fail(SWRB_E_INVALID, "range [%u,%u) exceeds %u meshlets", first, first + count, s->numMeshlets);
    CU(cudaSetDevice(s->dev->cudaDevice));
    CU(cudaMemcpyAsync(dst_host, s->meshlets + first, (size_t)count * sizeof(swr_meshlet), cudaMemcpyDeviceToHost, s->dev->stream));
    CU(cudaStreamSynchronize(s->dev->stream));
    return SWRB_OK;
}

int swrb_scene_meshlets_device(swrb_scene* s, void** out) {
    if (!s || !out) return fail(SWRB_E_INVALID, "null argument");
    *out = s->meshlets;
    return SWRB_OK;
}

int swrb_scene_touch(swrb_scene* s, uint32_t first, uint32_t count) {
    if (!s) return fail(SWRB_E_INVALID, "scene is null");
    if ((uint64_t)first + count > s->numMeshlets) return fail(SWRB_E_INVALID, "range [%u,%u) exceeds %u meshlets", first, first + count, s->numMeshlets);
    if (count) {
        if (s->attrDirtyLo >= s->attrDirtyHi) { s->attrDirtyLo = first; s->attrDirtyHi = first + count; }
        else { s->attrDirtyLo = std::min(s->attrDirtyLo, first); s->attrDirtyHi = std::max(s->attrDirtyHi, first + count); }
        if (s->dev->clipCacheMeshlets == s->meshlets) s->dev->clipCacheFb = nullptr;   // cached vertices belong to the old positions
    }
    return SWRB_OK;
}

int swrb_scene_set_skybox(swrb_scene* s, const swr_texture_desc* hdr) {
    if (!s) return fail(SWRB_E_INVALID, "scene is null");
    swrb_device* d = s->dev;
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));            // a resolve in flight may still be sampling the old skybox
    if (s->skyData) { CU(cudaFree(s->skyData)); s->skyData = nullptr; }
    if (!hdr) return SWRB_OK;
    if (!hdr->Data || hdr->Width < 8 || hdr->Height < 8 || (hdr->Width & (hdr->Width - 1)) || (hdr->Height & (hdr->Height - 1)) || hdr->MipLevels < 1 || hdr->MipLevels > 16)
        return fail(SWRB_E_INVALID, "skybox must be a power-of-two Texture2D<R11G11B10f> of at least 8x8 texels with 1..16 mip levels");
    const size_t texels = (size_t)hdr->LayerStride * std::max(1u, hdr->NumLayers);
    uint32_t* data = nullptr;
    CU(cudaMalloc(&data, texels * 4 + 256));
    cudaError_t e = cudaMemcpyAsync(data, hdr->Data, texels * 4, cudaMemcpyHostToDevice, d->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d->stream);      // the host texels are only borrowed for the call
    if (e != cudaSuccess) { cudaFree(data); return fail(SWRB_E_CUDA, "skybox upload: %s", cudaGetErrorString(e)); }
    s->skyData = data;                               // only a completely uploaded skybox becomes visible to swrb_resolve
    ResolveTexture& r = s->sky;
    r.data = s->skyData;
    r.width = hdr->Width; r.height = hdr->Height; r.mipLevels = hdr->MipLevels; r.numLayers = hdr->NumLayers;
    r.rowShift = hdr->RowShift; r.layerStride = hdr->LayerStride;
    for (int m = 0; m < 16; m++) r.mipOffsets[m] = hdr->MipOffsets[m];
    return SWRB_OK;
}

void swrb_scene_destroy(swrb_scene* s) {
    if (!s) return;
    cudaSetDevice(s->dev->cudaDevice);
    cudaStreamSynchronize(s->dev->stream);
    if (s->dev->clipCacheMeshlets == s->meshlets) s->dev->clipCacheFb = nullptr;
    cudaFree(s->meshlets); cudaFree(s->packedStaging); cudaFree(s->materials); cudaFree(s->textures); cudaFree(s->lights); cudaFree(s->attr); cudaFree(s->skyData);
    for (uint32_t* p : s->textureData) cudaFree(p);
    delete s;
}

// ---- framebuffer -------------------------------------------------------------------------------
int swrb_fb_create(swrb_device* d, uint32_t width, uint32_t height, uint32_t num_layers, swrb_fb** out) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    if (width == 0 || height == 0 || width % 4 || height % 4) return fail(SWRB_E_INVALID, "framebuffer size %ux%u must be a non-zero multiple of 4 (Rasterizer.h:67)", width, height);
    if (width > SWR_MAX_RENDER_SIZE || height > SWR_MAX_RENDER_SIZE) return fail(SWRB_E_INVALID, "framebuffer size %ux%u exceeds MaxRenderSize %d (Rasterizer.h:203)", width, height, SWR_MAX_RENDER_SIZE);
    if (num_layers < 2) return fail(SWRB_E_INVALID, "need at least 2 layers (colour/id + depth)");
    CU(cudaSetDevice(d->cudaDevice));
    CreateGuard<swrb_fb, swrb_fb_destroy> guard{ new swrb_fb() };
    swrb_fb* fb = guard.p;
    fb->dev = d; fb->width = width; fb->height = height; fb->layers = num_layers;
    fb->layerStride = (width * height + 63u) & ~63u;                      // Rasterizer.h:69
    CU(cudaMalloc(&fb->data, (size_t)fb->layerStride * num_layers * 4 + 256));
    CU(cudaMemsetAsync(fb->data, 0, (size_t)fb->layerStride * num_layers * 4, d->stream));
    CU(cudaEventCreateWithFlags(&fb->evReady, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&fb->evCopied, cudaEventDisableTiming));
    *out = guard.release();
    return SWRB_OK;
}

void swrb_fb_destroy(swrb_fb* fb) {
    if (!fb) return;
    cudaSetDevice(fb->dev->cudaDevice);
    cudaStreamSynchronize(fb->dev->stream);
    if (fb->copyPending) cudaEventSynchronize(fb->evCopied);      // a GetPixels on a side stream may still be reading the layers
    forget_draw_target(fb->dev, fb);
    if (fb->dev->clipCacheFb == fb) fb->dev->clipCacheFb = nullptr;
    cudaFree(fb->data); cudaFree(fb->keys);
    if (fb->evReady) cudaEventDestroy(fb->evReady);
    if (fb->evCopied) cudaEventDestroy(fb->evCopied);
    delete fb;
}

int swrb_fb_info(const swrb_fb* fb, swr_fb_info* out) {
    if (!fb || !out) return fail(SWRB_E_INVALID, "null argument");
    out->Width = fb->width; out->Height = fb->height; out->TileStride = fb->width / 4;
    out->LayerStride = fb->layerStride; out->NumLayers = fb->layers;
    return SWRB_OK;
}

int swrb_fb_keys_device(swrb_fb* fb, void** out, uint64_t* numWords) {
    if (!fb || !out) return fail(SWRB_E_INVALID, "null argument");
    if (!fb->keys || !fb->visInKeys || fb->layer0IsColor)
        return fail(SWRB_E_INVALID, "the framebuffer's result is not in its key buffer (draw a vis-buffer batch first, before any resolve or read-back)");
    *out = fb->keys;
    if (numWords) *numWords = (uint64_t)fb->width * fb->height;
    return SWRB_OK;
}

int swrb_fb_keys_touched(swrb_fb* fb) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    if (fb->dev->clipCacheFb == fb) fb->dev->clipCacheFb = nullptr;     // the keys may name meshlets this device never shaded
    return SWRB_OK;
}

// A GetPixels / send on a side stream may still be reading a layer: whatever overwrites the layers on the device's
// stream waits for it first.
static int fb_wait_readers(swrb_fb* fb) {
    if (fb->copyPending) {
        CU(cudaStreamWaitEvent(fb->dev->stream, fb->evCopied, 0));
        fb->copyPending = false;
    }
    return SWRB_OK;
}

static int fb_clear_layer_now(swrb_fb* fb, uint32_t layerA, uint32_t valueA, int layerB, uint32_t valueB) {
    swrb_device* d = fb->dev;
    { int rcw = fb_wait_readers(fb); if (rcw) return rcw; }
    uint32_t numVec = fb->width * fb->height / 4;
    StageScope ss(d, SWRB_STAGE_CLEAR);
    k_fb_clear<<<grid_for(d, numVec, 256, 8), 256, 0, d->stream>>>(
        reinterpret_cast<uint4*>(fb->data + (size_t)layerA * fb->layerStride), valueA,
        layerB >= 0 ? reinterpret_cast<uint4*>(fb->data + (size_t)layerB * fb->layerStride) : nullptr, valueB, numVec);
    d->launches++;
    CU(cudaGetLastError());
    return SWRB_OK;
}

// Brings layers 0/1 up to date: performs a recorded clear, or unpacks the last draw's keys.
static int fb_materialize(swrb_fb* fb) {
    swrb_device* d = fb->dev;
    if (fb->pendingClear) {
        fb->pendingClear = false;
        int rc = fb_clear_layer_now(fb, 0, fb->clearColor, 1, fb->clearDepthBits);
        if (rc) return rc;
    }
    if (fb->visInKeys) {
        { int rcw = fb_wait_readers(fb); if (rcw) return rcw; }
        uint32_t numVec = fb->width * fb->height / 4;
        StageScope ss(d, SWRB_STAGE_RASTER);
        k_keys_unpack<<<grid_for(d, numVec, 256, 8), 256, 0, d->stream>>>(
            reinterpret_cast<const ulonglong2*>(fb->keys), reinterpret_cast<uint4*>(fb->data),
            reinterpret_cast<uint4*>(fb->data + fb->layerStride), numVec, fb->keysClearMode ? 1 : 0, fb->keysClearColor,
            fb->layer0IsColor ? 1 : 0, d->ctl);
        d->launches++;
        CU(cudaGetLastError());
        fb->visInKeys = false;
        fb->layer0IsColor = false;
    }
    return SWRB_OK;
}

int swrb_fb_clear(swrb_fb* fb, uint32_t color, float depth) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    if (!(depth >= 0.0f)) return fail(SWRB_E_INVALID, "clear depth must be >= 0 (reverse-Z, Main.cpp:213)");
    // Recorded, not executed: the next draw performs the clear while it writes the tiles
    // (or any other access materialises it first).
    uint32_t bits; memcpy(&bits, &depth, 4);
    if (bits == 0x80000000u) bits = 0;
    forget_draw_target(fb->dev, fb);      // a clear recorded after an aborted draw must survive that draw's rollback
    fb->pendingClear = true;
    fb->visInKeys = false;
    fb->layer0IsColor = false;
    fb->clearColor = color;
    fb->clearDepthBits = bits;
    return SWRB_OK;
}

int swrb_fb_clear_layer(swrb_fb* fb, uint32_t layer, uint32_t value) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    if (layer >= fb->layers) return fail(SWRB_E_INVALID, "layer %u out of range", layer);
    CU(cudaSetDevice(fb->dev->cudaDevice));
    int rc = fb_materialize(fb);
    if (rc) return rc;
    return fb_clear_layer_now(fb, layer, value, -1, 0);
}

int swrb_fb_download_tiled(swrb_fb* fb, uint32_t layer, uint32_t* dst_host) {
    if (!fb || !dst_host) return fail(SWRB_E_INVALID, "null argument");
    if (layer >= fb->layers) return fail(SWRB_E_INVALID, "layer %u out of range", layer);
    CU(cudaSetDevice(fb->dev->cudaDevice));
    int rc = fb_materialize_for_read(fb, layer);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst_host, fb->data + (size_t)layer * fb->layerStride, (size_t)fb->width * fb->height * 4, cudaMemcpyDeviceToHost, fb->dev->stream));
    CU(cudaStreamSynchronize(fb->dev->stream));
    return check_overflow(fb->dev);
}

int swrb_fb_upload_tiled(swrb_fb* fb, uint32_t layer, const uint32_t* src_host) {
    if (!fb || !src_host) return fail(SWRB_E_INVALID, "null argument");
    if (layer >= fb->layers) return fail(SWRB_E_INVALID, "layer %u out of range", layer);
    CU(cudaSetDevice(fb->dev->cudaDevice));
    int rc = fb_materialize(fb);
    if (rc) return rc;
    CU(cudaMemcpyAsync(fb->data + (size_t)layer * fb->layerStride, src_host, (size_t)fb->width * fb->height * 4, cudaMemcpyHostToDevice, fb->dev->stream));
    CU(cudaStreamSynchronize(fb->dev->stream));
    return SWRB_OK;
}

// Layer 0 is already current once the resolve pass has written colour over it, and layers >= 2 never
// take part in the lazy vis-buffer; only depth / id reads force the key unpack.
static int fb_materialize_for_read(swrb_fb* fb, uint32_t layer) {
    if (fb->pendingClear && layer >= 2) return SWRB_OK;
    if (!fb->pendingClear && fb->visInKeys && (layer >= 2 || (layer == 0 && fb->layer0IsColor))) return SWRB_OK;
    return fb_materialize(fb);
}

// A copy that runs on another stream than the device's: it must see everything the device stream has enqueued for this
// framebuffer so far (the resolve pass, a key unpack this very call may have added), and later writers of the layers
// must wait for it (fb_wait_readers).
static int fb_order_side_stream(swrb_fb* fb, cudaStream_t stream) {
    if (stream == fb->dev->stream) return SWRB_OK;
    CU(cudaEventRecord(fb->evReady, fb->dev->stream));
    CU(cudaStreamWaitEvent(stream, fb->evReady, 0));
    return SWRB_OK;
}
static int fb_side_stream_done(swrb_fb* fb, cudaStream_t stream) {
    if (stream == fb->dev->stream) return SWRB_OK;
    CU(cudaEventRecord(fb->evCopied, stream));
    fb->copyPending = true;
    return SWRB_OK;
}

static int get_pixels_device_on(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride, cudaStream_t stream) {
    if (!fb || !dst_device) return fail(SWRB_E_INVALID, "null argument");
    if (layer >= fb->layers) return fail(SWRB_E_INVALID, "layer %u out of range", layer);
    if (stride < fb->width || stride % 4) return fail(SWRB_E_INVALID, "stride %u must be >= width and a multiple of 4", stride);
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    int rc = fb_materialize_for_read(fb, layer);
    if (rc) return rc;
    rc = fb_order_side_stream(fb, stream);
    if (rc) return rc;
    // rows [y0, y1) of the 4x4-tiled layer are one contiguous run of it, and land in rows [y0, y1) of the destination image
    const uint32_t y0 = fb->row0(), rows = fb->row1() - y0;
    uint32_t numVec = fb->width * rows / 4;
    // On the device's own stream the copy is on the critical path: fill the machine. On a caller's side
    // stream it runs beside the render kernels (typically storing to a peer GPU over NVLink): one block per
    // SM, four loads in flight per thread.
    const uint32_t grid = stream == d->stream ? grid_for(d, numVec, 256, 8) : std::max(1u, (uint32_t)d->numSMs);
    k_fb_detile<<<grid, 256, 0, stream>>>(
        reinterpret_cast<const uint4*>(fb->data + (size_t)layer * fb->layerStride + (size_t)y0 * fb->width),
        (uint32_t*)dst_device + (size_t)y0 * stride, fb->width, rows, stride);
    d->launches++;
    CU(cudaGetLastError());
    return fb_side_stream_done(fb, stream);
}

int swrb_fb_set_scissor_rows(swrb_fb* fb, uint32_t y0, uint32_t y1) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    if (y0 == 0 && (y1 == 0 || y1 == fb->height)) y1 = 0;                      // the whole framebuffer
    else if (y0 >= y1 || y1 > fb->height || y0 % 8 || (y1 % 8 && y1 != fb->height))
        return fail(SWRB_E_INVALID, "scissor rows [%u,%u): need y0 < y1 <= %u, y0 a multiple of 8, y1 a multiple of 8 or the height", y0, y1, fb->height);
    if (y0 == fb->bandY0 && y1 == fb->bandY1) return SWRB_OK;
    fb->bandY0 = y0; fb->bandY1 = y1;
    // rows outside the old scissor hold neither the next frame's seeds nor a vis-buffer: the next draw seeds the key buffer afresh,
    // and cached vertices of meshlets the old scissor dropped are missing
    fb->keysSeeded = false;
    if (fb->dev->clipCacheFb == fb) fb->dev->clipCacheFb = nullptr;
    return SWRB_OK;
}

int swrb_fb_get_scissor_rows(swrb_fb* fb, uint32_t* y0, uint32_t* y1) {
    if (!fb || !y0 || !y1) return fail(SWRB_E_INVALID, "null argument");
    *y0 = fb->row0(); *y1 = fb->row1();
    return SWRB_OK;
}

int swrb_fb_get_pixels_device(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    return get_pixels_device_on(fb, layer, dst_device, stride, fb->dev->stream);
}

int swrb_fb_get_pixels_device_on_stream(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride, void* cuda_stream) {
    if (!fb) return fail(SWRB_E_INVALID, "fb is null");
    return get_pixels_device_on(fb, layer, dst_device, stride, (cudaStream_t)cuda_stream);
}

int swrb_fb_send_pixels(swrb_fb* fb, uint32_t layer, void* dst_device, uint32_t stride, void* cuda_stream, const swrb_peer_sync* sync) {
    if (!fb || !dst_device || !sync) return fail(SWRB_E_INVALID, "null argument");
    if (layer >= fb->layers) return fail(SWRB_E_INVALID, "layer %u out of range", layer);
    if (stride < fb->width || stride % 4) return fail(SWRB_E_INVALID, "stride %u must be >= width and a multiple of 4", stride);
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    int rc = fb_materialize_for_read(fb, layer);
    if (rc) return rc;
    if (!d->peerCounter) {
        CU(cudaMalloc(&d->peerCounter, swrb_device::kPeerCounters * 4));
        CU(cudaMemsetAsync(d->peerCounter, 0, swrb_device::kPeerCounters * 4, d->stream));
        CU(cudaStreamSynchronize(d->stream));
    }
    cudaStream_t stream = cuda_stream ? (cudaStream_t)cuda_stream : d->stream;
    rc = fb_order_side_stream(fb, stream);
    if (rc) return rc;
    PeerSync ps;
    ps.waitFlag = reinterpret_cast<const unsigned long long*>(sync->WaitFlag); ps.waitValue = sync->WaitValue;
    ps.signalFlag = reinterpret_cast<unsigned long long*>(sync->SignalFlag); ps.signalValue = sync->SignalValue;
    // every send in flight counts its finished blocks in its own word (sends of one device may run on several streams);
    // a word is reused 64 sends later, long after its launch has zeroed it again
    ps.blockCounter = d->peerCounter + (d->peerCounterNext++ % swrb_device::kPeerCounters);
    // beside the render kernels: one block per SM; measured insensitive to the grid size between 32 and 148 blocks
    // (stores over NVLink are posted, a few hundred KB in flight keep the link busy)
    k_fb_detile_send<<<std::max(1u, (uint32_t)d->numSMs), 256, 0, stream>>>(
        reinterpret_cast<const uint4*>(fb->data + (size_t)layer * fb->layerStride + (size_t)fb->row0() * fb->width),
        (uint32_t*)dst_device + (size_t)fb->row0() * stride, fb->width, fb->row1() - fb->row0(), stride, ps);
    d->launches++;
    CU(cudaGetLastError());
    return fb_side_stream_done(fb, stream);
}

int swrb_peer_collect(swrb_device* d, void* cuda_stream, const uint64_t* ready_flags, uint32_t n, uint64_t expected,
                      uint64_t* const* ack_flags, uint64_t ack_value) {
    if (!d || !ready_flags || !ack_flags) return fail(SWRB_E_INVALID, "null argument");
    if (n == 0 || n > 15) return fail(SWRB_E_INVALID, "n = %u producers: must be in [1, 15]", n);
    CU(cudaSetDevice(d->cudaDevice));
    PeerAcks acks;
    for (uint32_t i = 0; i < 15; i++) acks.flag[i] = i < n ? reinterpret_cast<unsigned long long*>(ack_flags[i]) : nullptr;
    k_peer_collect<<<1, 32, 0, cuda_stream ? (cudaStream_t)cuda_stream : d->stream>>>(
        reinterpret_cast<const unsigned long long*>(ready_flags), n, expected, acks, ack_value);
    d->launches++;
    CU(cudaGetLastError());
    return SWRB_OK;
}

int swrb_fb_get_pixels_async(swrb_fb* fb, uint32_t layer, uint32_t* dst_host, uint32_t stride) {
    if (!fb || !dst_host) return fail(SWRB_E_INVALID, "null argument");
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    size_t need = (size_t)fb->width * fb->height;
    if (!d->copyStream) {
        CU(cudaStreamCreateWithFlags(&d->copyStream, cudaStreamNonBlocking));
        for (int i = 0; i < swrb_device::kDetileSlots; i++) {
            CU(cudaEventCreateWithFlags(&d->detiled[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&d->copied[i], cudaEventDisableTiming));
        }
    }
    if (d->detileCap < need) {
        CU(cudaStreamSynchronize(d->copyStream));
        for (int i = 0; i < swrb_device::kDetileSlots; i++) {
            if (d->detileScratch[i]) CU(cudaFree(d->detileScratch[i]));
            d->detileScratch[i] = nullptr; d->copyInFlight[i] = false;
        }
        d->detileCap = 0;
        for (int i = 0; i < swrb_device::kDetileSlots; i++) CU(cudaMalloc(&d->detileScratch[i], need * 4));
        d->detileCap = need;
    }
    const int slot = d->detileSlot;
    d->detileSlot = (slot + 1) % swrb_device::kDetileSlots;
    if (d->copyInFlight[slot]) CU(cudaStreamWaitEvent(d->stream, d->copied[slot], 0));     // the scratch image is free again once its copy has left
    int rc = swrb_fb_get_pixels_device(fb, layer, d->detileScratch[slot], fb->width);
    if (rc) return rc;
    CU(cudaEventRecord(d->detiled[slot], d->stream));
    CU(cudaStreamWaitEvent(d->copyStream, d->detiled[slot], 0));
    const uint32_t y0 = fb->row0(), rows = fb->row1() - y0;          // scissor rows only, to their place in the host image
    const uint32_t* src = d->detileScratch[slot] + (size_t)y0 * fb->width;
    if (stride == fb->width) CU(cudaMemcpyAsync(dst_host + (size_t)y0 * stride, src, (size_t)rows * fb->width * 4, cudaMemcpyDeviceToHost, d->copyStream));
    else CU(cudaMemcpy2DAsync(dst_host + (size_t)y0 * stride, (size_t)stride * 4, src, (size_t)fb->width * 4, (size_t)fb->width * 4, rows, cudaMemcpyDeviceToHost, d->copyStream));
    CU(cudaEventRecord(d->copied[slot], d->copyStream));
    d->copyInFlight[slot] = true;
    d->hostCopiesPending = true;
    return SWRB_OK;
}

int swrb_fb_get_pixels(swrb_fb* fb, uint32_t layer, uint32_t* dst_host, uint32_t stride) {
    int rc = swrb_fb_get_pixels_async(fb, layer, dst_host, stride);
    if (rc) return rc;
    return swrb_sync(fb->dev);
}

// ---- culling -----------------------------------------------------------------------------------
// glm::mat4 * glm::mat4 on column-major arrays: column c of the result is the combination of a's
// columns weighted by b's column c, summed left to right.
static void mat4_mul(const float* a, const float* b, float* r) {
    float t[16];
    for (int c = 0; c < 4; c++)
        for (int k = 0; k < 4; k++) {
            float acc = a[0 * 4 + k] * b[c * 4 + 0];
            acc = acc + a[1 * 4 + k] * b[c * 4 + 1];
            acc = acc + a[2 * 4 + k] * b[c * 4 + 2];
            acc = acc + a[3 * 4 + k] * b[c * 4 + 3];
            t[c * 4 + k] = acc;
        }
    memcpy(r, t, sizeof(t));
}

int swrb_frustum_planes(const float proj[16], const float view[16], const float model[16], float planes_out[5][4]) {
    if (!proj || !view || !model || !planes_out) return fail(SWRB_E_INVALID, "null argument");
    float pv[16], m[16];
    mat4_mul(proj, view, pv);
    mat4_mul(pv, model, m);                                       // Shading.cpp:781
    float planes[6][4];
    for (int i = 0; i < 3; i++) {                                 // Shading.cpp:784-791 (rows of the transposed matrix)
        float pa[4], pb[4];
        for (int c = 0; c < 4; c++) { pa[c] = m[c * 4 + 3] + m[c * 4 + i]; pb[c] = m[c * 4 + 3] - m[c * 4 + i]; }
        float la = sqrtf((pa[0] * pa[0] + pa[1] * pa[1]) + pa[2] * pa[2]);
        float lb = sqrtf((pb[0] * pb[0] + pb[1] * pb[1]) + pb[2] * pb[2]);
        for (int c = 0; c < 4; c++) { planes[i * 2][c] = pa[c] / la; planes[i * 2 + 1][c] = pb[c] / lb; }
    }
    memcpy(planes_out, planes, 5 * 4 * sizeof(float));            // plane 5 is never tested (Shading.cpp:806)
    return SWRB_OK;
}

int swrb_cull_meshlets(swrb_scene* s, uint32_t meshlet_offset, uint32_t count, const float proj[16], const float view[16],
                       const float model[16], uint16_t* bitmap_out_host, uint32_t* visible_out) {
    if (!s) return fail(SWRB_E_INVALID, "scene is null");
    if ((uint64_t)meshlet_offset + count > s->numMeshlets) return fail(SWRB_E_INVALID, "meshlet range out of bounds");
    swrb_device* d = s->dev;
    CU(cudaSetDevice(d->cudaDevice));
    CullPlanes planes;
    int rc = swrb_frustum_planes(proj, view, model, planes.p);
    if (rc) return rc;
    uint32_t words32 = (count + 31) / 32;
    if (d->cullBitmapCap < words32 * 32 || !d->cullBitmapDev) {
        if (d->cullBitmapDev) CU(cudaFree(d->cullBitmapDev));
        d->cullBitmapDev = nullptr;
        CU(cudaMalloc(&d->cullBitmapDev, (size_t)std::max(words32, 1u) * 4));
        d->cullBitmapCap = words32 * 32;
    }
    CU(cudaMemsetAsync(d->visibleDev, 0, 4, d->stream));
    if (count) {
        StageScope ss(d, SWRB_STAGE_CULL);
        k_cull_meshlets<<<(count + 255) / 256, 256, 0, d->stream>>>(s->meshlets + meshlet_offset, count, planes,
                                                                    reinterpret_cast<uint32_t*>(d->cullBitmapDev), d->visibleDev);
        d->launches++;
        CU(cudaGetLastError());
    }
    if (bitmap_out_host || visible_out) {
        if (bitmap_out_host && count) {
            // the kernel writes whole u32 words; copy only the u16 words the caller's array has
            CU(cudaMemcpyAsync(bitmap_out_host, d->cullBitmapDev, (size_t)((count + 15) / 16) * 2, cudaMemcpyDeviceToHost, d->stream));
        }
        uint32_t vis = 0;
        CU(cudaMemcpyAsync(&vis, d->visibleDev, 4, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        if (visible_out) *visible_out = vis;
    }
    return SWRB_OK;
}

// ---- HiZ occlusion culling ---------------------------------------------------------
struct swrb_hiz {
    swrb_device* dev;
    HizDesc desc;
    uint32_t layerStride;     // texels
};

int swrb_hiz_create(swrb_device* d, uint32_t fb_width, uint32_t fb_height, swrb_hiz** out) {
    if (!d || !out || fb_width < 16 || fb_height < 16) return fail(SWRB_E_INVALID, "bad argument");
    CU(cudaSetDevice(d->cudaDevice));
    auto bit_length = [](uint32_t v) { uint32_t n = 0; while (v) { n++; v >>= 1; } return n; };
    // Main.cpp:54-56: halfW = 1 << (32 - lzcnt((width - 1) / 2)); CreateTexture2D<R32f>(halfW, halfH, 16)
    uint32_t w = 1u << bit_length((fb_width - 1) / 2), h = 1u << bit_length((fb_height - 1) / 2);
    CreateGuard<swrb_hiz, swrb_hiz_destroy> guard{ new swrb_hiz() };
    swrb_hiz* z = guard.p;
    z->dev = d;
    z->desc.width = w; z->desc.height = h;
    z->desc.rowShift = 0;
    while ((1u << z->desc.rowShift) < std::max(w, 8u)) z->desc.rowShift++;                 // Texture.h:602
    uint32_t stride = 0, mip = 0;
    for (; mip < 16; mip++) {                                                             // Texture.h:607-612
        if ((w >> mip) < 4 || (h >> mip) < 4) break;
        z->desc.mipOffsets[mip] = stride;
        stride += ((w >> mip) * (h >> mip) + 63u) & ~63u;
    }
    for (uint32_t k = mip; k < 16; k++) z->desc.mipOffsets[k] = 0;
    z->desc.mipLevels = mip;
    z->layerStride = stride;
    CU(cudaMalloc(&z->desc.data, (size_t)stride * 4 + 256));
    CU(cudaMemsetAsync(z->desc.data, 0, (size_t)stride * 4, d->stream));
    *out = guard.release();
    return SWRB_OK;
}

void swrb_hiz_destroy(swrb_hiz* z) {
    if (!z) return;
    cudaSetDevice(z->dev->cudaDevice);
    cudaStreamSynchronize(z->dev->stream);
    cudaFree(z->desc.data);
    delete z;
}

int swrb_hiz_info(const swrb_hiz* z, swr_texture_desc* out) {
    if (!z || !out) return fail(SWRB_E_INVALID, "null argument");
    out->Width = z->desc.width; out->Height = z->desc.height; out->MipLevels = z->desc.mipLevels; out->NumLayers = 1;
    out->RowShift = z->desc.rowShift; out->LayerStride = z->layerStride; out->Data = nullptr;
    for (int k = 0; k < 16; k++) out->MipOffsets[k] = z->desc.mipOffsets[k];
    return SWRB_OK;
}

// texutil::DownsampleDepth(fb, depthMap) (ImageHelpers.cpp:243-247)
int swrb_hiz_build(swrb_hiz* z, swrb_fb* fb) {
    if (!z || !fb) return fail(SWRB_E_INVALID, "null argument");
    if (z->dev != fb->dev) return fail(SWRB_E_INVALID, "pyramid and framebuffer belong to different devices");
    swrb_device* d = z->dev;
    CU(cudaSetDevice(d->cudaDevice));
    int rc = fb_materialize_for_read(fb, 1);
    if (rc) return rc;
    uint32_t maxDim = std::max(fb->width, fb->height), rootLevel = 0;
    while ((1u << rootLevel) <= maxDim) rootLevel++;                                      // 32 - lzcnt(max(W,H))
    StageScope ss(d, SWRB_STAGE_CULL);
    for (uint32_t m = 0; m + 3 <= rootLevel && m < z->desc.mipLevels; m++) {
        const bool top = (m + 3 == rootLevel);
        const uint32_t texel = 1u << (m + 1), blockPx = (top ? 4u : 8u) * texel;
        uint32_t tx = top ? 4u : ((fb->width + blockPx - 1) / blockPx) * 8u, ty = top ? 4u : ((fb->height + blockPx - 1) / blockPx) * 8u;
        dim3 grid((tx + 31) / 32, (ty + 7) / 8);
        k_hiz_level<<<grid, 256, 0, d->stream>>>(reinterpret_cast<const float*>(fb->data + fb->layerStride), z->desc, m, fb->width, fb->height, tx, ty);
        d->launches++;
    }
    CU(cudaGetLastError());
    return SWRB_OK;
}

int swrb_hiz_download(swrb_hiz* z, float* dst_host) {
    if (!z || !dst_host) return fail(SWRB_E_INVALID, "null argument");
    CU(cudaSetDevice(z->dev->cudaDevice));
    CU(cudaMemcpyAsync(dst_host, z->desc.data, (size_t)z->layerStride * 4, cudaMemcpyDeviceToHost, z->dev->stream));
    CU(cudaStreamSynchronize(z->dev->stream));
    return SWRB_OK;
}

// ShadingContext::CullMeshlets(bitmap, meshlets, count, P, V, M, prevV, frameSize, depthMap) — Shading.cpp:775-869
int swrb_cull_meshlets_hiz(swrb_scene* s, uint32_t meshlet_offset, uint32_t count, const float proj[16], const float view[16],
                           const float model[16], const float prev_view[16], float frame_w, float frame_h, swrb_hiz* hiz,
                           uint16_t* bitmap_out_host, uint32_t* visible_out) {
    if (!s || !proj || !view || !model) return fail(SWRB_E_INVALID, "null argument");
    if ((uint64_t)meshlet_offset + count > s->numMeshlets) return fail(SWRB_E_INVALID, "meshlet range out of bounds");
    if (hiz && (!prev_view || hiz->dev != s->dev)) return fail(SWRB_E_INVALID, "HiZ culling needs prev_view and a pyramid of the same device");
    swrb_device* d = s->dev;
    CU(cudaSetDevice(d->cudaDevice));
    CullHizParams cp;
    int rc = swrb_frustum_planes(proj, view, model, cp.planes);
    if (rc) return rc;
    HizDesc hz{};
    cp.useHiz = hiz ? 1 : 0;
    cp.frameW = frame_w; cp.frameH = frame_h;
    cp.scale = sqrtf((model[0] * model[0] + model[1] * model[1]) + model[2] * model[2]);   // glm::length(vec3(modelMat[0]))
    cp.znear = proj[3 * 4 + 2]; cp.p00 = proj[0]; cp.p11 = proj[1 * 4 + 1];
    if (hiz) { mat4_mul(prev_view, model, cp.objectToPrevView); hz = hiz->desc; }
    else memset(cp.objectToPrevView, 0, sizeof(cp.objectToPrevView));
    uint32_t words32 = (count + 31) / 32;
    if (d->cullBitmapCap < words32 * 32 || !d->cullBitmapDev) {
        if (d->cullBitmapDev) CU(cudaFree(d->cullBitmapDev));
        d->cullBitmapDev = nullptr;
        CU(cudaMalloc(&d->cullBitmapDev, (size_t)std::max(words32, 1u) * 4));
        d->cullBitmapCap = words32 * 32;
    }
    CU(cudaMemsetAsync(d->visibleDev, 0, 4, d->stream));
    if (count) {
        StageScope ss(d, SWRB_STAGE_CULL);
        k_cull_meshlets_hiz<<<(count + 255) / 256, 256, 0, d->stream>>>(s->meshlets + meshlet_offset, count, cp, hz,
                                                                        reinterpret_cast<uint32_t*>(d->cullBitmapDev), d->visibleDev);
        d->launches++;
        CU(cudaGetLastError());
    }
    if (bitmap_out_host || visible_out) {
        if (bitmap_out_host && count)
            CU(cudaMemcpyAsync(bitmap_out_host, d->cullBitmapDev, (size_t)((count + 15) / 16) * 2, cudaMemcpyDeviceToHost, d->stream));
        uint32_t vis = 0;
        CU(cudaMemcpyAsync(&vis, d->visibleDev, 4, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
        if (visible_out) *visible_out = vis;
    }
    return SWRB_OK;
}

// ---- draw --------------------------------------------------------------------------------------
// Work-list capacities. Only triangles too large for the mesh kernel's inline raster (and alpha-tested / clipped ones) become
// records, so the lists are sized from what draws actually produce, not from meshlets x 128: start at min(worst case, 2 M
// records) — or what swrb_device_reserve asked for — and, when a draw does overflow (it is aborted on the device and reported
// as SWRB_E_BIN_OVERFLOW, never truncated), the next draw finds the capacities doubled.
static int ensure_work_buffers(swrb_device* d, swrb_fb* fb, uint64_t maxTris, bool alphaTest) {
    uint64_t needTris = std::max<uint64_t>(std::max<uint64_t>(std::min<uint64_t>(maxTris, 2u << 20), d->reserveTris), 1024);
    needTris = std::max(needTris, std::min<uint64_t>(d->growTris, std::max<uint64_t>(maxTris, 1024)));
    if (needTris > d->triCap) {
        uint64_t cap = d->triCap;
        int rc = ensure_buffer((void**)&d->tris, &cap, needTris, sizeof(TriRecord));
        if (rc) return rc;
        if (d->trisW) { CU(cudaFree(d->trisW)); d->trisW = nullptr; }
        if (d->clipRemap) { CU(cudaFree(d->clipRemap)); d->clipRemap = nullptr; }
        if (d->alphaTris) { CU(cudaFree(d->alphaTris)); d->alphaTris = nullptr; }
        d->triCap = needTris;
    }
    if (alphaTest && !d->trisW) {      // alpha-tested triangles have their own record list (+ 1/w per vertex)
        CU(cudaMalloc(&d->trisW, d->triCap * sizeof(TriRecordW)));
        CU(cudaMalloc(&d->alphaTris, d->triCap * sizeof(TriRecord)));
    }
    const bool clipping = !(d->flags & SWRB_FLAG_BINNING) && (d->flags & SWRB_FLAG_CLIPPING);
    if (alphaTest && clipping && !d->clipRemap) CU(cudaMalloc(&d->clipRemap, d->triCap * 2 * sizeof(float4)));
    uint64_t needBins = std::max<uint64_t>(std::max<uint64_t>(d->reserveBins, d->growBins), 2 * d->triCap + (1u << 20));   // (also the clip list: 2 words per entry)
    if (needBins > d->binCap) {
        int rc = ensure_buffer((void**)&d->binEntries, &d->binCap, needBins, 4);
        if (rc) return rc;
    }
    uint64_t needSuper = std::max<uint64_t>(std::max<uint64_t>(1u << 18, d->triCap), d->growBins / 4);
    if (needSuper > d->superCap) {
        int rc = ensure_buffer((void**)&d->superEntries, &d->superCap, needSuper, 4);
        if (rc) return rc;
    }
    uint64_t needBig = std::max<uint64_t>(std::max<uint64_t>(1u << 20, d->triCap / 2), d->growBins / 2);
    if (needBig > d->bigItemCap) {
        int rc = ensure_buffer((void**)&d->bigItems, &d->bigItemCap, needBig, sizeof(BigItem));
        if (rc) return rc;
    }
    uint32_t tilesX = (fb->width + kTileSize - 1) >> kTileShift, tilesY = (fb->height + kTileSize - 1) >> kTileShift;
    uint32_t numTiles = tilesX * tilesY;
    if (numTiles > d->tileCap) {
        if (d->tileCount) CU(cudaFree(d->tileCount));
        d->tileCount = nullptr;
        CU(cudaMalloc(&d->tileCount, ((size_t)numTiles * 4 + 4 + 3 * 160) * 4));
        d->tileCap = numTiles;
        d->workClean = false;
    }
    d->tileOffset = d->tileCount + d->tileCap;
    d->tileCursor = d->tileOffset + d->tileCap + 1;
    d->activeTiles = d->tileCursor + d->tileCap;
    d->superCount = d->activeTiles + d->tileCap;
    d->superOffset = d->superCount + 160;
    d->superCursor = d->superOffset + 160;
    return SWRB_OK;
}

static FrameParams frame_params(const swrb_device* d, const swrb_fb* fb, bool binned) {
    FrameParams fp;
    fp.program = SWRB_PROGRAM_VISBUFFER;
    fp.inlineMaxArea = d->inlineMaxArea;
    fp.uniformMatrix = 0;
    memset(fp.M, 0, sizeof(fp.M));
    fp.workBegin = 0; fp.workEnd = 0;
    fp.width = fb->width; fp.height = fb->height;
    fp.halfW = (int32_t)fb->width / 2; fp.halfH = (int32_t)fb->height / 2;          // Rasterizer.cpp:508
    fp.fixX = (float)(fp.halfW * 16); fp.fixY = (float)(fp.halfH * 16);               // :272
    bool guard = binned || (d->flags & SWRB_FLAG_GUARDBAND);                          // :509 vs :155-156
    fp.bx = guard ? (float)SWR_MAX_RENDER_SIZE / (float)fb->width : 1.0f;
    fp.by = guard ? (float)SWR_MAX_RENDER_SIZE / (float)fb->height : 1.0f;
    fp.tilesX = (fb->width + kTileSize - 1) >> kTileShift;
    fp.tilesY = (fb->height + kTileSize - 1) >> kTileShift;
    fp.layerStride = fb->layerStride;
    fp.clipMode = binned ? 0u : ((d->flags & SWRB_FLAG_CLIPPING) ? 2u : 1u);          // :567-569 vs :209
    fp.bandY0 = (int32_t)fb->row0(); fp.bandY1 = (int32_t)fb->row1();
    fp.bandCull = (fb->row0() != 0 || fb->row1() != fb->height) ? 1u : 0u;
    // pixel row y holds y/w in [(y - halfH) / halfH, (y + 1 - halfH) / halfH): the band, one row wider on both sides
    fp.bandNdcLo = (float)(fp.bandY0 - 1 - fp.halfH) / (float)fp.halfH;
    fp.bandNdcHi = (float)(fp.bandY1 + 1 - fp.halfH) / (float)fp.halfH;
    return fp;
}

// What a draw call hands the kernels: the device-resident per-draw items and what the host knows about them.
struct DrawList {
    const DrawItem* items;       // device
    uint32_t numDraws;
    uint64_t totalWork;          // meshlets over all draws
    bool uniformMatrix;          // every draw uses matrix M0 (the clip cache may stand in for the resolve pass's transform)
    const float* M0;
    bool anyCull;                // some draw carries a cull bitmap or fused frustum planes
    const uint32_t* runStarts = nullptr;   // DeferredShader on a scene that mixes textured and material-less meshlets: first work
    uint32_t numRuns = 0;                  // item of every run of one kind (runStarts[0] == 0); null = the batch is one run
};

// Work-item indices at which a batch switches between textured and material-less meshlets (see gbuffer.cuh).
static void compute_runs(const swrb_scene* scene, const swrb_draw_desc* draws, uint32_t numDraws, std::vector<uint32_t>& runStarts) {
    runStarts.clear();
    if (!scene || scene->numMaterialless == 0 || scene->numMaterialless == scene->numMeshlets) return;
    uint32_t work = 0;
    int kind = -1;
    for (uint32_t i = 0; i < numDraws; i++)
        for (uint32_t m = 0; m < draws[i].MeshletCount; m++, work++) {
            const int k = scene->materialless[draws[i].MeshletOffset + m];
            if (k != kind) { runStarts.push_back(work); kind = k; }
        }
}

// Host draws -> DrawItem array (firstWork prefix, planes, cull bitmap pointers).
static int fill_draw_items(swrb_device* d, uint32_t numMeshletsDev, const swrb_draw_desc* draws, uint32_t numDraws, DrawItem* items,
                           uint16_t* cullDev, uint64_t* totalWorkOut, bool* uniformOut, bool* anyCullOut, cudaStream_t uploadStream) {
    uint32_t firstWork = 0;
    size_t cullCursor = 0;
    bool uniform = true, anyCull = false;
    for (uint32_t i = 0; i < numDraws; i++) {
        if ((uint64_t)draws[i].MeshletOffset + draws[i].MeshletCount > numMeshletsDev)
            return fail(SWRB_E_INVALID, "draw %u: meshlet range [%u,+%u) exceeds the scene's %u meshlets", i, draws[i].MeshletOffset, draws[i].MeshletCount, numMeshletsDev);
        DrawItem& it = items[i];
        memcpy(it.M, draws[i].ObjectToClip, sizeof(it.M));
        memcpy(it.planes, draws[i].FrustumPlanes, sizeof(it.planes));
        memcpy(it.objectToWorld, draws[i].ObjectToWorld, sizeof(it.objectToWorld));
        it.meshletOffset = draws[i].MeshletOffset;
        it.count = draws[i].MeshletCount;
        it.firstWork = firstWork;
        it.fusedCull = (d->flags & SWRB_FLAG_FUSED_FRUSTUM_CULL) ? 1u : 0u;
        it.cullBitmap = nullptr;
        if (draws[i].CullBitmapHost) {
            size_t words = (draws[i].MeshletCount + 15) / 16;
            CU(cudaMemcpyAsync(cullDev + cullCursor, draws[i].CullBitmapHost, words * 2, cudaMemcpyHostToDevice, uploadStream));
            it.cullBitmap = cullDev + cullCursor;
            cullCursor += words + (words & 1);
        } else if (draws[i].UseDeviceCullBitmap) {
            if (!d->cullBitmapDev || d->cullBitmapCap < draws[i].MeshletCount)
                return fail(SWRB_E_INVALID, "draw %u: UseDeviceCullBitmap without a preceding swrb_cull_meshlets of >= %u meshlets", i, draws[i].MeshletCount);
            it.cullBitmap = d->cullBitmapDev;
        }
        if (i && memcmp(draws[i].ObjectToClip, draws[0].ObjectToClip, sizeof(it.M)) != 0) uniform = false;
        if (it.cullBitmap != nullptr || it.fusedCull) anyCull = true;
        if ((uint64_t)firstWork + draws[i].MeshletCount >= (1ull << 25)) return fail(SWRB_E_INVALID, "too many meshlets in one batch");
        firstWork += draws[i].MeshletCount;
    }
    *totalWorkOut = firstWork;
    *uniformOut = uniform;
    *anyCullOut = anyCull;
    return SWRB_OK;
}

static size_t cull_words_needed(const swrb_draw_desc* draws, uint32_t numDraws) {
    size_t cullWords = 0;
    for (uint32_t i = 0; i < numDraws; i++) if (draws[i].CullBitmapHost) cullWords += (draws[i].MeshletCount + 15) / 16 + 1;
    return cullWords;
}

static void launch_mesh(swrb_device* d, bool binned, uint32_t meshGrid, const swr_meshlet* meshlets, const swr_material* materials, const DrawItem* draws, uint32_t numDraws,
                        uint32_t totalWork, const FrameParams& fp, unsigned long long* keys, const MeshOut& mo, DevCtl* ctl) {
#define SWRB_MESH(B, S) k_mesh_setup<B, S><<<meshGrid, kMeshWarps * 32, 0, d->stream>>>(meshlets, materials, draws, numDraws, totalWork, fp, keys, mo, ctl)
    if (fp.bandCull) { if (binned) SWRB_MESH(true, true); else SWRB_MESH(false, true); }
    else { if (binned) SWRB_MESH(true, false); else SWRB_MESH(false, false); }
#undef SWRB_MESH
}

static int draw_deferred(swrb_fb* fb, const swr_meshlet* meshletsDev, const swr_material* materialsDev, const ResolveTexture* texturesDev,
                         const DrawList& dl, FrameParams fp, MeshOut mo, uint32_t meshGrid);

static int draw_list(swrb_fb* fb, const swr_meshlet* meshletsDev, uint32_t numMeshletsDev, const swr_material* materialsDev,
                     const ResolveTexture* texturesDev, bool alphaTest, const DrawList& dl, bool forResolve, uint32_t program) {
    swrb_device* d = fb->dev;
    if (program > SWRB_PROGRAM_DEFERRED) return fail(SWRB_E_INVALID, "unknown program %u", program);
    if (program == SWRB_PROGRAM_DEFERRED && fb->layers < 3) return fail(SWRB_E_INVALID, "DeferredShader writes three layers (Shading.cpp:344-414); the framebuffer has %u", fb->layers);
    const bool binned = (d->flags & SWRB_FLAG_BINNING) != 0;
    const uint32_t numDraws = dl.numDraws;
    const uint64_t totalWork = dl.totalWork;
    if (totalWork == 0) return SWRB_OK;
    int rc = ensure_work_buffers(d, fb, totalWork * SWR_MAX_PRIMS, alphaTest || program == SWRB_PROGRAM_DEFERRED);   // FS_EncodeGBuffer's records carry 1/w like the alpha program's
    if (rc) return rc;
    nvtx_range nv("swrb.draw");

    FrameParams fp = frame_params(d, fb, binned);
    fp.uniformMatrix = dl.uniformMatrix ? 1u : 0u;
    memcpy(fp.M, dl.M0, sizeof(fp.M));
    fp.workBegin = 0; fp.workEnd = (uint32_t)totalWork;
    const uint32_t numTiles = fp.tilesX * fp.tilesY;
    const uint32_t numVec = fb->width * fb->height / 4;
    uint32_t* depthLayer = fb->data + fb->layerStride;
    MeshOut mo;
    mo.tris = d->tris; mo.alphaTris = d->alphaTris; mo.alphaW = d->trisW; mo.triCapacity = (uint32_t)std::min<uint64_t>(d->triCap, 0xFFFFFFFFu);
    mo.tileCount = nullptr; mo.superCount = nullptr; mo.clipList = reinterpret_cast<uint2*>(d->binEntries); mo.clipCache = nullptr;
    // persistent mesh grid (64 registers -> at most 4 resident blocks per SM)
    const uint32_t meshGrid = grid_for(d, totalWork, kMeshWarps, d->meshBlocksPerSM);

    // ---- OverdrawShader / DeferredShader: no lazy vis-buffer — straight into the layers
    if (program != SWRB_PROGRAM_VISBUFFER) {
        fp.program = program;
        if (!fb->keys) CU(cudaMalloc(&fb->keys, (size_t)fb->width * fb->height * 8));
        rc = fb_materialize(fb);                // layers current (a recorded clear is executed now)
        if (rc) return rc;
        rc = fb_wait_readers(fb);
        if (rc) return rc;
        d->clipCacheFb = nullptr;
        d->workClean = false;
        fb->keysSeeded = false;                 // the key buffer doubles as this program's scratch
        if (program == SWRB_PROGRAM_OVERDRAW) {
            {
                StageScope ss(d, SWRB_STAGE_CLEAR);
                k_overdraw_begin<<<grid_for(d, numVec, 256, 8), 256, 0, d->stream>>>(reinterpret_cast<ulonglong2*>(fb->keys), numVec, d->ctl);
                d->launches++;
            }
            {
                StageScope ss(d, SWRB_STAGE_MESH);
                launch_mesh(d, false, meshGrid, meshletsDev, materialsDev, dl.items, numDraws, (uint32_t)totalWork, fp, fb->keys, mo, d->ctl);
                d->launches++;
                if (fp.clipMode == 2u) {
                    k_clip_triangles<<<d->numSMs, 128, 0, d->stream>>>(reinterpret_cast<const uint2*>(d->binEntries), meshletsDev, materialsDev, dl.items, fp,
                                                                      d->tris, d->alphaTris, d->trisW, d->clipRemap, mo.triCapacity, d->ctl);
                    d->launches++;
                }
            }
            {
                StageScope ss(d, SWRB_STAGE_RASTER);
                k_raster_overdraw<<<resident_grid(d, k_raster_overdraw, 256), 256, 0, d->stream>>>(d->tris, fp, fb->keys, depthLayer, d->ctl);
                k_overdraw_finish<<<grid_for(d, (uint64_t)fb->width * fb->height, 256, 8), 256, 0, d->stream>>>(fb->keys, fb->data, fb->width * fb->height, d->ctl);
                d->launches += 2;
            }
        } else {
            mo.alphaTris = d->alphaTris; mo.alphaW = d->trisW;
            rc = draw_deferred(fb, meshletsDev, materialsDev, texturesDev, dl, fp, mo, meshGrid);
            if (rc) return rc;
        }
        CU(cudaGetLastError());
        return SWRB_OK;
    }

    // ---- per-vertex clip cache for the resolve pass (any batch overwrites it, so it always describes the last one)
    d->clipCacheFb = nullptr;
    if (forResolve && !(d->flags & SWRB_FLAG_NO_RESOLVE_CACHE)) {
        if (numMeshletsDev > d->clipCacheCap) {
            rc = ensure_buffer((void**)&d->clipCache, &d->clipCacheCap, numMeshletsDev, sizeof(float4) * SWR_MAX_VERTICES);
            if (rc) return rc;
        }
        mo.clipCache = d->clipCache;
        d->clipCacheFb = fb;
        d->clipCacheMeshlets = meshletsDev;
        memcpy(d->clipCacheM, dl.M0, sizeof(d->clipCacheM));
        d->clipCacheUniform = dl.uniformMatrix;
    }

    // ---- key buffer: seeds = the depth every pixel has before this draw (or the pending clear's depth)
    if (!fb->keys) CU(cudaMalloc(&fb->keys, (size_t)fb->width * fb->height * 8));
    if (fb->visInKeys) {           // an earlier draw's result is still only in the keys: the layers must be current to re-seed
        rc = fb_materialize(fb);
        if (rc) return rc;
    }
    const int clearMode = fb->pendingClear ? 1 : 0;
    // the last resolve pass may have left exactly these seeds behind (and reset the work counters)
    const bool seeded = clearMode && fb->keysSeeded && fb->seedDepthBits == fb->clearDepthBits;
    if (!seeded || !d->workClean) {
        StageScope ss(d, SWRB_STAGE_CLEAR);
        const uint32_t vec = seeded ? 0u : numVec;
        k_frame_begin<<<seeded ? 8u : grid_for(d, numVec, 256, 8), 256, 0, d->stream>>>(clearMode ? nullptr : reinterpret_cast<const uint4*>(depthLayer), fb->clearDepthBits,
                                                                         reinterpret_cast<ulonglong2*>(fb->keys), vec, d->tileCount, d->tileCursor, numTiles,
                                                                         d->superCount, d->superCursor, d->ctl);
        d->launches++;
    }
    d->workClean = false;
    fb->keysSeeded = false;

    // record consumers are grid-stride loops over a device-side count; size their grids from the last count
    // the host has seen (any grid is correct, a fitting one avoids launching a thousand idle blocks)
    const uint64_t recEstimate = std::min<uint64_t>(d->triCap, std::max<uint64_t>(4096, 4 * (uint64_t)d->lastTriCount));
    const uint32_t scatterGrid = std::max<uint32_t>(grid_for(d, recEstimate, 256, 8), (uint32_t)d->numSMs);
    if (binned) {
        mo.tileCount = d->tileCount; mo.superCount = d->superCount;
        {
            StageScope ss(d, SWRB_STAGE_MESH);
            launch_mesh(d, true, meshGrid, meshletsDev, materialsDev, dl.items, numDraws, (uint32_t)totalWork, fp, fb->keys, mo, d->ctl);
            d->launches++;
        }
        {
            StageScope ss(d, SWRB_STAGE_BIN);
            BinBuffers bb;
            bb.tileCount = d->tileCount; bb.tileOffset = d->tileOffset; bb.tileCursor = d->tileCursor; bb.activeTiles = d->activeTiles;
            bb.superCount = d->superCount; bb.superOffset = d->superOffset; bb.superCursor = d->superCursor;
            bb.binEntries = d->binEntries; bb.binCapacity = (uint32_t)std::min<uint64_t>(d->binCap, 0xFFFFFFFFu);
            bb.superEntries = d->superEntries; bb.superCapacity = (uint32_t)std::min<uint64_t>(d->superCap, 0xFFFFFFFFu);
            k_bin_scatter<<<scatterGrid, kScatterThreads, (numTiles + 1) * sizeof(uint32_t), d->stream>>>(d->tris, fp, bb, d->ctl);
            d->launches++;
        }
        {
            StageScope ss(d, SWRB_STAGE_RASTER);
            k_tile_raster<<<std::min<uint32_t>(numTiles, (uint32_t)d->numSMs * 4u), kTileThreads, 0, d->stream>>>(
                d->tris, d->tileOffset, d->activeTiles, d->binEntries, d->superOffset, d->superEntries, fp, fb->keys, d->ctl);
            d->launches++;
        }
    } else {
        {
            StageScope ss(d, SWRB_STAGE_MESH);
            launch_mesh(d, false, meshGrid, meshletsDev, materialsDev, dl.items, numDraws, (uint32_t)totalWork, fp, fb->keys, mo, d->ctl);   // (the bin-entry buffer is idle on this path: it holds the clip list)
            d->launches++;
            if (fp.clipMode == 2u) {       // Clipper::ClipTriangles: pieces join the record / alpha lists before they are consumed
                k_clip_triangles<<<d->numSMs, 128, 0, d->stream>>>(reinterpret_cast<const uint2*>(d->binEntries), meshletsDev, materialsDev, dl.items, fp,
                                                                  d->tris, d->alphaTris, d->trisW, d->clipRemap, mo.triCapacity, d->ctl);
                d->launches++;
            }
        }
        {
            StageScope ss(d, SWRB_STAGE_RASTER);
            k_raster_direct<<<scatterGrid, 256, 0, d->stream>>>(d->tris, fp, fb->keys, d->bigItems, (uint32_t)std::min<uint64_t>(d->bigItemCap, 0xFFFFFFFFu), d->ctl);
            k_raster_big<<<resident_grid(d, k_raster_big, 256), 256, 0, d->stream>>>(d->tris, d->bigItems, fp, fb->keys, d->ctl);
            d->launches += 2;
        }
    }
    if (alphaTest && texturesDev != nullptr) {     // FS_EncodeSurfaceId<true> for the alpha list (both raster modes)
        StageScope ss(d, SWRB_STAGE_RASTER);
        k_raster_alpha<<<resident_grid(d, k_raster_alpha, kAlphaWarps * 32), kAlphaWarps * 32, 0, d->stream>>>(d->alphaTris, d->trisW, fp, texturesDev, d->clipRemap, fb->keys, d->ctl);
        d->launches++;
    }
    // The vis-buffer now lives in the key buffer; layers 0/1 are produced on demand (fb_materialize) or the
    // resolve pass reads the keys directly. If a work list overflowed, the draw was aborted on the device
    // and the next synchronising call reports SWRB_E_BIN_OVERFLOW and rolls this state back.
    note_draw_target(d, fb, fb->pendingClear);
    fb->visInKeys = true;
    fb->keysClearMode = clearMode != 0;
    fb->keysClearColor = fb->clearColor;
    fb->layer0IsColor = false;
    fb->pendingClear = false;
    CU(cudaGetLastError());
    return SWRB_OK;
}

// ShadingContext::DeferredShader (gbuffer.cuh). The layers are current on entry (draw_list materialized them). Per run of
// meshlets of one kind: seed the keys with the depth layer, mesh kernel (every surviving triangle -> a record with 1/w
// and its draw), [clipper], pass 1 (keys), pass 2 (the winners store depth / base colour / packed normal).
static int draw_deferred(swrb_fb* fb, const swr_meshlet* meshletsDev, const swr_material* materialsDev, const ResolveTexture* texturesDev,
                         const DrawList& dl, FrameParams fp, MeshOut mo, uint32_t meshGrid) {
    swrb_device* d = fb->dev;
    if (!mo.alphaTris || !mo.alphaW) return fail(SWRB_E_CUDA, "DeferredShader: record lists missing");
    const uint32_t numVec = fb->width * fb->height / 4;
    uint32_t* layer0 = fb->data;
    uint32_t* layer1 = fb->data + fb->layerStride;
    uint32_t* layer2 = fb->data + 2 * (size_t)fb->layerStride;
    const uint32_t numRuns = dl.runStarts ? dl.numRuns : 1u;
    for (uint32_t r = 0; r < numRuns; r++) {
        fp.workBegin = dl.runStarts ? dl.runStarts[r] : 0u;
        fp.workEnd = (dl.runStarts && r + 1 < numRuns) ? dl.runStarts[r + 1] : (uint32_t)dl.totalWork;
        {
            StageScope ss(d, SWRB_STAGE_CLEAR);
            k_gbuffer_begin<<<grid_for(d, numVec, 256, 8), 256, 0, d->stream>>>(reinterpret_cast<ulonglong2*>(fb->keys), reinterpret_cast<const uint4*>(layer1), numVec, d->ctl);
            d->launches++;
        }
        {
            StageScope ss(d, SWRB_STAGE_MESH);
            launch_mesh(d, false, meshGrid, meshletsDev, materialsDev, dl.items, dl.numDraws, (uint32_t)dl.totalWork, fp, fb->keys, mo, d->ctl);
            d->launches++;
            if (fp.clipMode == 2u) {
                k_clip_triangles<<<d->numSMs, 128, 0, d->stream>>>(reinterpret_cast<const uint2*>(d->binEntries), meshletsDev, materialsDev, dl.items, fp,
                                                                  d->tris, d->alphaTris, d->trisW, d->clipRemap, mo.triCapacity, d->ctl);
                d->launches++;
            }
        }
        {
            StageScope ss(d, SWRB_STAGE_RASTER);
            k_raster_gbuffer<false><<<resident_grid(d, k_raster_gbuffer<false>, 256), 256, 0, d->stream>>>(d->alphaTris, d->trisW, fp, meshletsDev, materialsDev, texturesDev, d->clipRemap, dl.items,
                                                                     fb->keys, layer0, layer1, layer2, d->ctl);
            k_raster_gbuffer<true><<<resident_grid(d, k_raster_gbuffer<true>, 256), 256, 0, d->stream>>>(d->alphaTris, d->trisW, fp, meshletsDev, materialsDev, texturesDev, d->clipRemap, dl.items,
                                                                    fb->keys, layer0, layer1, layer2, d->ctl);
            d->launches += 2;
        }
    }
    return SWRB_OK;
}

// Host-described draws: stage the items in pinned memory, upload, draw.
static int draw_internal(swrb_fb* fb, const swr_meshlet* meshletsDev, uint32_t numMeshletsDev, const swr_material* materialsDev,
                         const ResolveTexture* texturesDev, bool alphaTest, const swrb_draw_desc* draws, uint32_t numDraws,
                         bool forResolve = true, uint32_t program = SWRB_PROGRAM_VISBUFFER, const swrb_scene* scene = nullptr) {
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    if (numDraws == 0) return SWRB_OK;
    if (numDraws > d->drawItemCap) {
        if (d->drawItems) CU(cudaFree(d->drawItems));
        d->drawItems = nullptr;
        CU(cudaMalloc(&d->drawItems, (size_t)numDraws * sizeof(DrawItem)));
        d->drawItemCap = numDraws;
    }
    int slot = d->drawStagingNext;
    d->drawStagingNext = (slot + 1) % swrb_device::kStagingSlots;
    CU(cudaEventSynchronize(d->drawStagingDone[slot]));
    if (d->drawStagingCap[slot] < numDraws) {
        if (d->drawStaging[slot]) CU(cudaFreeHost(d->drawStaging[slot]));
        d->drawStaging[slot] = nullptr;
        CU(cudaMallocHost(&d->drawStaging[slot], (size_t)numDraws * sizeof(DrawItem)));
        d->drawStagingCap[slot] = numDraws;
    }
    DrawItem* items = d->drawStaging[slot];
    const size_t cullWords = cull_words_needed(draws, numDraws);
    if (cullWords > d->cullUploadCap) {
        if (d->cullUpload) CU(cudaFree(d->cullUpload));
        d->cullUpload = nullptr;
        CU(cudaMalloc(&d->cullUpload, cullWords * 2));
        d->cullUploadCap = (uint32_t)cullWords;
    }
    DrawList dl;
    int rc = fill_draw_items(d, numMeshletsDev, draws, numDraws, items, d->cullUpload, &dl.totalWork, &dl.uniformMatrix, &dl.anyCull, d->stream);
    if (rc) return rc;
    CU(cudaMemcpyAsync(d->drawItems, items, (size_t)numDraws * sizeof(DrawItem), cudaMemcpyHostToDevice, d->stream));
    CU(cudaEventRecord(d->drawStagingDone[slot], d->stream));
    dl.items = d->drawItems; dl.numDraws = numDraws; dl.M0 = draws[0].ObjectToClip;
    std::vector<uint32_t> runs;
    if (program == SWRB_PROGRAM_DEFERRED) {
        compute_runs(scene, draws, numDraws, runs);
        if (!runs.empty()) { dl.runStarts = runs.data(); dl.numRuns = (uint32_t)runs.size(); }
    }
    return draw_list(fb, meshletsDev, numMeshletsDev, materialsDev, texturesDev, alphaTest, dl, forResolve, program);
}

int swrb_draw_batch(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws) {
    if (!fb || !scene || (!draws && num_draws)) return fail(SWRB_E_INVALID, "null argument");
    if (fb->dev != scene->dev) return fail(SWRB_E_INVALID, "framebuffer and scene belong to different devices");
    return draw_internal(fb, scene->meshlets, scene->numMeshlets, scene->materials, scene->textures, scene->hasAlphaTest, draws, num_draws);
}

int swrb_draw_batch_program(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws, uint32_t program) {
    if (!fb || !scene || (!draws && num_draws)) return fail(SWRB_E_INVALID, "null argument");
    if (fb->dev != scene->dev) return fail(SWRB_E_INVALID, "framebuffer and scene belong to different devices");
    if (program == SWRB_PROGRAM_DEFERRED) {
        // FS_EncodeGBuffer samples Material::Texture of every meshlet that has a material (Shading.cpp:364): a null dereference
        // upstream, an error here
        for (uint32_t i = 0; i < scene->numMaterials; i++)
            if (scene->materialTextureIds[i] < 0) return fail(SWRB_E_INVALID, "DeferredShader: material %u has no texture", i);
    }
    return draw_internal(fb, scene->meshlets, scene->numMeshlets, scene->materials, scene->textures, scene->hasAlphaTest, draws, num_draws,
                         true, program, scene);
}

int swrb_draw_meshlets(swrb_fb* fb, swrb_scene* scene, const swrb_draw_desc* draw) {
    return swrb_draw_batch(fb, scene, draw, 1);
}

int swrb_draw_meshlets_host(swrb_fb* fb, const swr_meshlet* meshlets_host, uint32_t count, const float object_to_clip[16],
                            const uint16_t* cull_bitmap_host) {
    if (!fb || !meshlets_host || !object_to_clip) return fail(SWRB_E_INVALID, "null argument");
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    if (count > d->hostDrawCap) {
        if (d->hostDrawMeshlets) CU(cudaFree(d->hostDrawMeshlets));
        d->hostDrawMeshlets = nullptr; d->hostDrawCap = 0;
        CU(cudaMalloc(&d->hostDrawMeshlets, (size_t)count * sizeof(swr_meshlet)));
        d->hostDrawCap = count;
    }
    CU(cudaMemcpyAsync(d->hostDrawMeshlets, meshlets_host, (size_t)count * sizeof(swr_meshlet), cudaMemcpyHostToDevice, d->stream));
    swrb_draw_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.MeshletOffset = 0;
    desc.MeshletCount = count;
    memcpy(desc.ObjectToClip, object_to_clip, sizeof(desc.ObjectToClip));
    desc.CullBitmapHost = cull_bitmap_host;
    return draw_internal(fb, d->hostDrawMeshlets, count, nullptr, nullptr, false, &desc, 1, /*forResolve=*/false);
}

// ---- batches and frames: the per-frame host work of the reference's loop, done once ---------------------------------
struct swrb_batch {
    swrb_device* dev = nullptr;
    swrb_scene* scene = nullptr;
    DrawItem* items = nullptr;       // device
    uint16_t* cullBitmaps = nullptr; // device copies of the host bitmaps the draws named
    uint32_t numDraws = 0;
    uint64_t totalWork = 0;
    bool uniformMatrix = true, anyCull = false;
    float M0[16] = {};
    std::vector<uint32_t> runs;      // compute_runs of the draws (DeferredShader)
};

void swrb_batch_destroy(swrb_batch* b) {
    if (!b) return;
    if (b->dev) {
        cudaSetDevice(b->dev->cudaDevice);
        cudaStreamSynchronize(b->dev->stream);
    }
    cudaFree(b->items); cudaFree(b->cullBitmaps);
    delete b;
}

int swrb_batch_create(swrb_scene* scene, const swrb_draw_desc* draws, uint32_t num_draws, swrb_batch** out) {
    if (!scene || !draws || !num_draws || !out) return fail(SWRB_E_INVALID, "null argument");
    swrb_device* d = scene->dev;
    CU(cudaSetDevice(d->cudaDevice));
    for (uint32_t i = 0; i < num_draws; i++)
        if (draws[i].UseDeviceCullBitmap) return fail(SWRB_E_INVALID, "draw %u: a batch cannot capture the transient device cull bitmap (use FrustumPlanes or CullBitmapHost)", i);
    CreateGuard<swrb_batch, swrb_batch_destroy> guard{ new swrb_batch() };
    swrb_batch* b = guard.p;
    b->dev = d; b->scene = scene; b->numDraws = num_draws;
    std::vector<DrawItem> items(num_draws);
    const size_t cullWords = cull_words_needed(draws, num_draws);
    if (cullWords) CU(cudaMalloc(&b->cullBitmaps, cullWords * 2));
    int rc = fill_draw_items(d, scene->numMeshlets, draws, num_draws, items.data(), b->cullBitmaps, &b->totalWork, &b->uniformMatrix, &b->anyCull, d->stream);
    if (rc) return rc;
    memcpy(b->M0, draws[0].ObjectToClip, sizeof(b->M0));
    compute_runs(scene, draws, num_draws, b->runs);
    CU(cudaMalloc(&b->items, (size_t)num_draws * sizeof(DrawItem)));
    CU(cudaMemcpyAsync(b->items, items.data(), (size_t)num_draws * sizeof(DrawItem), cudaMemcpyHostToDevice, d->stream));
    CU(cudaStreamSynchronize(d->stream));     // the host arrays are only borrowed for the call
    *out = guard.release();
    return SWRB_OK;
}

int swrb_draw_prepared(swrb_fb* fb, const swrb_batch* batch, uint32_t program) {
    if (!fb || !batch) return fail(SWRB_E_INVALID, "null argument");
    swrb_scene* scene = batch->scene;
    if (fb->dev != scene->dev) return fail(SWRB_E_INVALID, "framebuffer and batch belong to different devices");
    CU(cudaSetDevice(fb->dev->cudaDevice));
    DrawList dl;
    dl.items = batch->items; dl.numDraws = batch->numDraws; dl.totalWork = batch->totalWork; dl.uniformMatrix = batch->uniformMatrix; dl.M0 = batch->M0;
    dl.anyCull = batch->anyCull;
    if (program == SWRB_PROGRAM_DEFERRED) {
        for (uint32_t i = 0; i < scene->numMaterials; i++)
            if (scene->materialTextureIds[i] < 0) return fail(SWRB_E_INVALID, "DeferredShader: material %u has no texture", i);
        if (!batch->runs.empty()) { dl.runStarts = batch->runs.data(); dl.numRuns = (uint32_t)batch->runs.size(); }
    }
    return draw_list(fb, scene->meshlets, scene->numMeshlets, scene->materials, scene->textures, scene->hasAlphaTest, dl, true, program);
}

static int resolve_internal(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* u, int debugLayer);

int swrb_frame_submit(swrb_fb* fb, const swrb_frame_desc* f) {
    if (!fb || !f || !f->Batch) return fail(SWRB_E_INVALID, "null argument");
    nvtx_range nv("swrb.frame");
    int rc = swrb_fb_clear(fb, f->ClearColor, f->ClearDepth);                               // Main.cpp:213
    if (rc) return rc;
    rc = swrb_draw_prepared(fb, f->Batch, SWRB_PROGRAM_VISBUFFER);                          // :216-240
    if (rc) return rc;
    if (f->Uniforms) {
        rc = resolve_internal(fb, f->Batch->scene, f->Uniforms, SWRB_LAYER_NONE);            // :249
        if (rc) return rc;
    }
    if (f->PixelsDevice) {                                                                  // :270 (GetPixels)
        const uint32_t stride = f->PixelsStride ? f->PixelsStride : fb->width;
        if (f->PeerSync) rc = swrb_fb_send_pixels(fb, 0, f->PixelsDevice, stride, f->PixelsStream, f->PeerSync);
        else rc = swrb_fb_get_pixels_device_on_stream(fb, 0, f->PixelsDevice, stride, f->PixelsStream ? f->PixelsStream : (void*)fb->dev->stream);
        if (rc) return rc;
    }
    if (f->PixelsHost) {
        rc = swrb_fb_get_pixels_async(fb, 0, f->PixelsHost, f->PixelsStride ? f->PixelsStride : fb->width);
        if (rc) return rc;
    }
    return SWRB_OK;
}

// ---- resolve -----------------------------------------------------------------------------------
int swrb_resolve(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* u) {
    return resolve_internal(fb, scene, u, SWRB_LAYER_NONE);
}

int swrb_resolve_debug(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* u, uint32_t layer) {
    if (layer < SWRB_LAYER_BASE_COLOR || layer > SWRB_LAYER_OVERDRAW_QUAD) return fail(SWRB_E_INVALID, "debug layer %u out of range [1, 7]", layer);
    if (!fb || !scene || !u) return fail(SWRB_E_INVALID, "null argument");
    if (layer <= SWRB_LAYER_METALLIC_ROUGHNESS) return resolve_internal(fb, scene, u, (int)layer);
    // the remaining layers are functions of the colour word alone (Shading.cpp:755-766)
    if (fb->dev != scene->dev) return fail(SWRB_E_INVALID, "framebuffer and scene belong to different devices");
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    int rc = fb_materialize(fb);
    if (rc) return rc;
    rc = fb_wait_readers(fb);
    if (rc) return rc;
    StageScope ss(d, SWRB_STAGE_RESOLVE);
    dim3 grid((fb->width + 31) / 32, (fb->height + 7) / 8);
    k_resolve_debug_ids<<<grid, dim3(32, 8), 0, d->stream>>>(fb->data, fb->data + fb->layerStride, fb->width, fb->height, (int)layer);
    d->launches++;
    CU(cudaGetLastError());
    return SWRB_OK;
}

static int resolve_internal(swrb_fb* fb, swrb_scene* scene, const swrb_shading_uniforms* u, int debugLayer) {
    if (!fb || !scene || !u) return fail(SWRB_E_INVALID, "null argument");
    if (fb->dev != scene->dev) return fail(SWRB_E_INVALID, "framebuffer and scene belong to different devices");
    swrb_device* d = fb->dev;
    CU(cudaSetDevice(d->cudaDevice));
    nvtx_range nv("swrb.resolve");
    const bool fromKeys = fb->visInKeys && !fb->layer0IsColor;
    if (!fromKeys) {
        int rc = fb_materialize(fb);
        if (rc) return rc;
    }
    { int rc = fb_wait_readers(fb); if (rc) return rc; }     // the pass overwrites layer 0
    ResolveParams rp;
    memset(&rp, 0, sizeof(rp));
    memcpy(rp.objectToClip, u->ObjectToClip, sizeof(rp.objectToClip));
    memcpy(rp.objectToWorld, u->ObjectToWorld, sizeof(rp.objectToWorld));
    memcpy(rp.invScreenProj, u->InvScreenProj, sizeof(rp.invScreenProj));
    memcpy(rp.viewPos, u->ViewPos, sizeof(rp.viewPos));
    rp.exposure = u->Exposure;
    rp.width = fb->width; rp.height = fb->height;
    rp.pixScaleX = 2.0f / (float)fb->width; rp.pixScaleY = 2.0f / (float)fb->height;
    rp.pixBiasX = 0.5f * rp.pixScaleX - 1.0f; rp.pixBiasY = 0.5f * rp.pixScaleY - 1.0f;
    rp.meshlets = scene->meshlets; rp.materials = scene->materials; rp.textures = scene->textures;
    rp.lights = scene->lights; rp.numLights = scene->numLights; rp.numMeshlets = scene->numMeshlets;
    for (uint32_t i = 0; i < std::min<uint32_t>(kInlineLights, (uint32_t)scene->lightsHost.size()); i++) rp.lightsInline[i] = scene->lightsHost[i];
    rp.color = fb->data; rp.depth = fb->data + fb->layerStride;
    rp.keys = fb->keys; rp.keysClearMode = fb->keysClearMode ? 1u : 0u; rp.clearColor = fb->keysClearColor;
    // attribute table: decode what changed since the last resolve (scene upload / swrb_scene_update_meshlets)
    if (!scene->attr && scene->numMeshlets) CU(cudaMalloc(&scene->attr, (size_t)scene->numMeshlets * SWR_MAX_VERTICES * 2 * sizeof(float4)));
    if (scene->attrDirtyLo < scene->attrDirtyHi) {
        StageScope ss(d, SWRB_STAGE_RESOLVE);
        const uint32_t count = scene->attrDirtyHi - scene->attrDirtyLo;
        k_decode_attributes<<<(count * SWR_MAX_VERTICES + 255) / 256, 256, 0, d->stream>>>(scene->meshlets, scene->attrDirtyLo, count, scene->attr);
        d->launches++;
        scene->attrDirtyLo = scene->attrDirtyHi = 0;
    }
    rp.attr = scene->attr;
    // The clip cache may stand in for the per-pixel corner transform only if every surface id the pass can meet
    // was written by the last batch (vis-buffer still in that batch's keys, drawn from a cleared framebuffer whose
    // clear depth marks sky), that batch used ONE matrix, and it is bit for bit the matrix of this resolve
    // (ShadingContext::Resolve transforms with the context's current ObjectToClipMat, Shading.cpp:509-511).
    const bool cached = fromKeys && fb->keysClearMode && fb->clearDepthBits == 0 && d->clipCacheFb == fb && d->clipCacheUniform &&
                        d->clipCacheMeshlets == scene->meshlets && memcmp(d->clipCacheM, u->ObjectToClip, sizeof(d->clipCacheM)) == 0;
    rp.clipCache = cached ? d->clipCache : nullptr;
    rp.debugLayer = debugLayer;
    rp.blockY0 = fb->row0() / 8;                     // scissor rows: only those blocks run (multiples of 8 by construction)
    if (debugLayer != SWRB_LAYER_NONE) {            // ResolveDebug: surface only, no lighting, no light markers
        StageScope ss(d, SWRB_STAGE_RESOLVE);
        dim3 grid((fb->width + 15) / 16, (fb->row1() - fb->row0() + 7) / 8), block(32, kResolveWarps);
        if (fromKeys) k_resolve<true, false, true><<<grid, block, 0, d->stream>>>(rp, d->ctl);
        else k_resolve<false, false, true><<<grid, block, 0, d->stream>>>(rp, d->ctl);
        d->launches++;
        if (fromKeys) fb->layer0IsColor = true;
        CU(cudaGetLastError());
        return SWRB_OK;
    }
    // Reading the keys retires them: the pass stores the depth layer (the layers are current afterwards) and leaves the next
    // frame's seeds — the framebuffer's last clear depth — and, on the binned path's buffers, a reset draw state behind.
    if (fromKeys) {
        rp.reseed = 1u; rp.reseedDepthBits = fb->clearDepthBits;
        rp.depthOut = fb->data + fb->layerStride; rp.keysOut = fb->keys;
        if (d->tileCount && !d->workClean) {
            rp.resetTileCount = d->tileCount; rp.resetTileCursor = d->tileCursor; rp.resetNumTiles = d->tileCap;
            rp.resetSuperCount = d->superCount; rp.resetSuperCursor = d->superCursor;
        }
    }
    {
        StageScope ss(d, SWRB_STAGE_RESOLVE);
        // 4 warps = 16 x 8 pixels per block: small blocks pack better beside other contexts' mesh blocks (measured +2 %
        // frames/s over 8-warp blocks, same single-frame time)
        dim3 grid((fb->width + 15) / 16, (fb->row1() - fb->row0() + 7) / 8), block(32, kResolveWarps);
        const bool sky = scene->skyData != nullptr;
        if (sky) rp.sky = scene->sky;
        if (cached) { if (sky) k_resolve<true, true, false, true><<<grid, block, 0, d->stream>>>(rp, d->ctl); else k_resolve<true, true><<<grid, block, 0, d->stream>>>(rp, d->ctl); }
        else if (fromKeys) { if (sky) k_resolve<true, false, false, true><<<grid, block, 0, d->stream>>>(rp, d->ctl); else k_resolve<true, false><<<grid, block, 0, d->stream>>>(rp, d->ctl); }
        else { if (sky) k_resolve<false, false, false, true><<<grid, block, 0, d->stream>>>(rp, d->ctl); else k_resolve<false, false><<<grid, block, 0, d->stream>>>(rp, d->ctl); }
        d->launches++;
    }
    if (fromKeys) {
        fb->visInKeys = false;                       // layers 0 (colour) and 1 (depth) are current
        fb->layer0IsColor = false;
        fb->keysSeeded = true; fb->seedDepthBits = fb->clearDepthBits;
        if (rp.resetTileCount) d->workClean = true;
        if (d->clipCacheFb == fb) d->clipCacheFb = nullptr;
    }
    // Tail of Resolve (Shading.cpp:690-731): point / spot lights inside the frustum become soft discs, in light order.
    for (const swr_light& light : scene->lightsHost) {
        if (light.Type == 0) continue;
        const float* m = u->WorldToClip;
        float clip[4];
        for (int r = 0; r < 4; r++)     // glm mat4 * vec4: (m0*x + m1*y) + (m2*z + m3*w)
            clip[r] = (m[0 * 4 + r] * light.Position[0] + m[1 * 4 + r] * light.Position[1]) + (m[2 * 4 + r] * light.Position[2] + m[3 * 4 + r] * 1.0f);
        if (fmaxf(fmaxf(fabsf(clip[0]), fabsf(clip[1])), fabsf(clip[2])) > clip[3]) continue;
        LightDisc ld;
        ld.depth = clip[2] / clip[3];
        const float sx = (clip[0] / clip[3]) * 0.5f + 0.5f, sy = (clip[1] / clip[3]) * 0.5f + 0.5f;
        ld.radius = ((float)std::max(fb->width, fb->height) / 30.0f) / clip[3];
        ld.cx = sx * (float)fb->width; ld.cy = sy * (float)fb->height;
        ld.startX = std::max((int32_t)(ld.cx - ld.radius), 0) & ~3; ld.startY = std::max((int32_t)(ld.cy - ld.radius), 0) & ~3;
        ld.endX = std::min((int32_t)(ld.cx + ld.radius), (int32_t)fb->width); ld.endY = std::min((int32_t)(ld.cy + ld.radius), (int32_t)fb->height);
        memcpy(ld.color, light.Color, sizeof(ld.color));
        if (ld.startX >= ld.endX || ld.startY >= ld.endY) continue;
        StageScope ss(d, SWRB_STAGE_RESOLVE);
        dim3 grid((uint32_t)(((ld.endX + 3) & ~3) - ld.startX + 31) / 32, (uint32_t)(((ld.endY + 3) & ~3) - ld.startY + 7) / 8);
        k_light_marker<<<grid, dim3(32, 8), 0, d->stream>>>(ld, fb->data, fb->data + fb->layerStride, nullptr, fb->width, d->ctl);
        d->launches++;
    }
    CU(cudaGetLastError());
    return SWRB_OK;
}


// ---- pinned host memory ------------------------------------------------------------------------
int swrb_alloc_pinned(swrb_device* d, uint64_t bytes, void** out) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaMallocHost(out, bytes));
    return SWRB_OK;
}
int swrb_free_pinned(swrb_device* d, void* ptr) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaFreeHost(ptr));
    return SWRB_OK;
}

// ---- timing ------------------------------------------------------------------------------------
int swrb_timer_begin(swrb_device* d) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaEventRecord(d->timerBegin, d->stream));
    return SWRB_OK;
}

int swrb_timer_end(swrb_device* d, float* elapsed_ms) {
    if (!d || !elapsed_ms) return fail(SWRB_E_INVALID, "null argument");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaEventRecord(d->timerEnd, d->stream));
    CU(cudaEventSynchronize(d->timerEnd));
    CU(cudaEventElapsedTime(elapsed_ms, d->timerBegin, d->timerEnd));
    return check_overflow(d);
}

int swrb_flush_l2(swrb_device* d) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    if (!d->l2Scratch) {
        d->l2ScratchBytes = (size_t)512 << 20;   // two halves of 256 MB, each 2x the 126 MB L2
        // (contents are irrelevant)
        CU(cudaMalloc(&d->l2Scratch, d->l2ScratchBytes));
    }
    // write 256 MB (evicts everything, leaves dirty scratch lines), then read a second 256 MB so that what
    // stays in L2 is clean: the first timed kernel then pays for its own traffic, not for our write-backs
    CU(cudaMemsetAsync(d->l2Scratch, 0x5A, d->l2ScratchBytes / 2, d->stream));
    k_l2_read<<<d->numSMs * 8, 256, 0, d->stream>>>(reinterpret_cast<const uint4*>((const char*)d->l2Scratch + d->l2ScratchBytes / 2),
                                                      (uint32_t)(d->l2ScratchBytes / 2 / 16), reinterpret_cast<uint32_t*>(d->l2Scratch));
    CU(cudaGetLastError());
    return SWRB_OK;
}

int swrb_device_enable_stage_timing(swrb_device* d, int enable) {
    if (!d) return fail(SWRB_E_INVALID, "device is null");
    CU(cudaSetDevice(d->cudaDevice));
    if (enable && !d->st) {
        d->st = new StageTimer();
        memset(d->st->used, 0, sizeof(d->st->used));
        for (int s = 0; s < SWRB_STAGE_COUNT_; s++)
            for (int k = 0; k < 8; k++) { CU(cudaEventCreate(&d->st->begin[s][k])); CU(cudaEventCreate(&d->st->end[s][k])); }
    }
    d->stageTiming = enable != 0;
    if (d->st) memset(d->st->used, 0, sizeof(d->st->used));
    return SWRB_OK;
}

int swrb_get_stage_times(swrb_device* d, float out_us[SWRB_STAGE_COUNT_], uint32_t launches_out[SWRB_STAGE_COUNT_]) {
    if (!d || !out_us) return fail(SWRB_E_INVALID, "null argument");
    if (!d->st) return fail(SWRB_E_INVALID, "stage timing is not enabled");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));
    for (int s = 0; s < SWRB_STAGE_COUNT_; s++) {
        float total = 0;
        for (uint32_t k = 0; k < d->st->used[s]; k++) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, d->st->begin[s][k], d->st->end[s][k]));
            total += ms;
        }
        out_us[s] = total * 1000.0f;
        if (launches_out) launches_out[s] = d->st->used[s];
    }
    memset(d->st->used, 0, sizeof(d->st->used));
    return check_overflow(d);
}

int swrb_get_launch_count(swrb_device* d, uint64_t* out) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    *out = d->launches;
    return SWRB_OK;
}

int swrb_get_draw_stats(swrb_device* d, uint32_t out[4]) {
    if (!d || !out) return fail(SWRB_E_INVALID, "null argument");
    CU(cudaSetDevice(d->cudaDevice));
    CU(cudaStreamSynchronize(d->stream));
    int rc = check_overflow(d);
    if (rc) return rc;
    const DevCtl& c = *d->ctlHost;      // after a resolve pass the counters of the finished draw live in last*
    out[0] = d->workClean ? c.lastTriCount : c.triCount; out[1] = d->workClean ? c.lastBigCount : c.bigCount;
    out[2] = d->workClean ? c.lastBinTotal : c.binTotal + c.superTotal; out[3] = 0;
    return SWRB_OK;
}

}  // extern "C"
