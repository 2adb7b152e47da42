// common.cuh — shared definitions of the CUDA implementation behind include/atlas_rt.h (sm_100a only).
//
// Arithmetic contract of every kernel in this directory: IEEE fp32, round-to-nearest, NO fused multiply-add on any
// value that feeds a split decision, a node box, a visit-order decision or a reported hit. The library is compiled
// with -fmad=false and without fast-math; where it matters the __f*_rn intrinsics are used as well so that a future
// flag change cannot silently contract them. min/max follow glm 0.9.8 / GLSL comparison forms (see gl_min/gl_max).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/atlas_rt.h"
#include "layouts.h"

namespace atlas {

constexpr float kFltMax = 3.402823466e+38f;

// glm::min / glm::max (glm 0.9.8 func_common.inl) and GLSL min/max: result is the FIRST argument unless the
// comparison holds, which also pins what happens with NaN and with -0 vs +0.
__host__ __device__ __forceinline__ float gl_min(float x, float y) { return (y < x) ? y : x; }
__host__ __device__ __forceinline__ float gl_max(float x, float y) { return (x < y) ? y : x; }
__host__ __device__ __forceinline__ float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }

// Order-preserving map float -> int32 so that integer atomicMin/atomicMax reduce floats exactly and independently of
// the order of operations (-0 sorts below +0; NaN is outside the contract).
__host__ __device__ __forceinline__ int ord_from_float(float f) {
#ifdef __CUDA_ARCH__
    int i = __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float float_from_ord(int i) {
    int b = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

struct Box3 {
    float lo[3], hi[3];
};

__host__ __device__ __forceinline__ Box3 empty_box() {
    Box3 b;
    b.lo[0] = b.lo[1] = b.lo[2] = kFltMax;
    b.hi[0] = b.hi[1] = b.hi[2] = -kFltMax;
    return b;
}

// AABB::GetSurfaceArea — src/engine/volume/AABB.cpp:109-115: 2 * ((dx*dy + dy*dz) + dz*dx), no clamping.
__device__ __forceinline__ float surface_area(const float lo[3], const float hi[3]) {
    const float dx = __fsub_rn(hi[0], lo[0]), dy = __fsub_rn(hi[1], lo[1]), dz = __fsub_rn(hi[2], lo[2]);
    return __fmul_rn(2.0f, __fadd_rn(__fadd_rn(__fmul_rn(dx, dy), __fmul_rn(dy, dz)), __fmul_rn(dz, dx)));
}
__device__ __forceinline__ float surface_area(const Box3& b) { return surface_area(b.lo, b.hi); }

}   // namespace atlas

// ------------------------------------------------------------------------------------------------ host objects
struct atlas_rt_context {
    // Objects (bvh / mesh / scene) outlive calls and may be freed from another thread after the creating thread has
    // destroyed "its" context (engine meshes are built on job-system workers and assembled elsewhere), so the context is
    // reference counted: atlas_rt_context_destroy drops the creator's reference, every live object holds one more.
    std::atomic<int> refs{1};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int smCount = 148;
    uint64_t launches = 0;
    std::string error;
    unsigned long long* dCounters = nullptr;   // 16 x u64: [0..5] traversal counters / overflow flag, [8..15] ray-queue heads (one per compute stream)
    unsigned int* dStreamState = nullptr;      // 64 x u32: [0] upload watermark of a streaming host-buffer trace, [1..] per-chunk completion counts
    void* waitValue32 = nullptr;               // cuStreamWaitValue32 / cuStreamWriteValue32 (driver entry points), or null: chunked pipeline instead of streaming
    void* writeValue32 = nullptr;
    int streamSpinLog2 = 22;                   // bound of one watermark wait in the kernel: 2^22 x 256 ns, about a second
    int traceStreaming = 0;                    // ATLAS_RT_TRACE_STREAMING=1: host-buffer traces as ONE persistent launch that consumes rays while they arrive
                                               // (measured slower than the chunked pipeline on C2: arrival order forfeits the longest-first fetch order)
    int l2PersistMB = 0;                       // ATLAS_RT_L2_PERSIST_MB: size of the persisting-L2 carve-out used for a scene's hottest node array (0 = off)
    float l2HitRatio = 1.0f;
    size_t l2WindowMax = 0;
    void* pinned = nullptr;                    // small pinned staging area for read-backs
    size_t pinnedBytes = 0;
    // copy engines used to overlap H2D / trace / D2H when a trace call is given host buffers (api.cu)
    cudaStream_t copyIn = nullptr, copyOut = nullptr;
    cudaStream_t computeExtra[7] = {};   // extra compute streams: the chunks of a pipelined host-buffer trace run side by side
    int pipeStreams = 8;                 // compute streams such a call uses (the context stream + computeExtra)
    cudaEvent_t pipeEvents[38] = {};   // pipelined host-buffer trace: [0] staging ready, [1+c] chunk c uploaded, [17+c] chunk c traced, [33] all done;
                                       // streaming trace: [1+c] chunk c uploaded (c < 32), [33] all downloaded, [34] ordering stream done; [35] join scratch, [36], [37] pipelined call on staging set 0 / 1 done
    void* levelSlots = nullptr;                // pinned: per-level flags the builder's kernels write for the host (build.cu)
    // scheduling knobs of the persistent traversal kernel (trace.cu); ATLAS_RT_TRACE_* environment variables override
    int traceLeafThreshold = 8;     // lanes waiting at a leaf before the warp runs a leaf round
    int traceRefillThreshold = 10;  // idle lanes before the warp fetches new rays (swept again with the final kernel: 8-12 is best, 16 costs 1.5-3 %)
    int traceBlocksPerSM = 9;
    int chainLaunch = 1;         // builder level loop as a chain of programmatic dependent launches
    int buildWide = 0;           // ATLAS_RT_BUILD_WIDE=1: the tree below the big levels level by level over global memory instead of per-subtree in shared memory
    int binCtasPerSM = 2;        // CTAs per SM of the builder's binning kernel (each merges its shared bins into global ones)
    int traceLongestFirst = 1;      // fetch rays longest-estimated-path first (hides the drain of the longest rays)
    int traceLongestFirstMin = 65536;
    int traceRaysPerWarp = 96;      // small batches use fewer persistent warps so each warp still sees this many rays
    int traceMinBlocksPerSM = 2;    // ... but never fewer CTAs than this per SM
    int ptLanes = 4;                // ATLAS_RT_PT_LANES: sample passes of one atlas_rt_pathtrace_bounces call that run side by side (1..4)
    int streamBlocksPerSM = 7;      // CTAs per SM of a streaming launch (the per-chunk ordering kernels need the rest of the SM)
    void* stageIn[2] = {nullptr, nullptr};   // persistent staging of ATLAS_RT_PIPELINED host-buffer traces, two sets used in turn
    void* stageOut[2] = {nullptr, nullptr};
    size_t stageInBytes[2] = {0, 0}, stageOutBytes[2] = {0, 0};
    bool stageUsed[2] = {false, false};
    int stageNext = 0;
    bool pendingJoin = false;       // ATLAS_RT_PIPELINED host-buffer traces are in flight that the context stream has not been ordered after
    bool pipeTimeline = false;      // ATLAS_RT_PIPE_TIMELINE: per-chunk upload / trace completion times of the chunked pipeline on stderr
    cudaStream_t sortStream = nullptr;   // high-priority stream of a streaming launch's per-chunk ordering kernels
    // worker contexts (own stream + own pinned level flags each) that atlas_rt_build_blas_batch builds on side by side
    atlas_rt_context* workers[16] = {};
    int batchWorkers = 8;
};

struct atlas_rt_bvh {
    atlas_rt_context* ctx = nullptr;
    uint64_t nodeCount = 0, refCount = 0;
    float4* nodes = nullptr;      // GPUBVHNode, 4 x float4 each
    uint32_t* order = nullptr;    // source index per flattened slot
    uint8_t* endOfNode = nullptr;
    uint64_t stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

struct atlas_rt_mesh {
    atlas_rt_context* ctx = nullptr;
    const atlas_rt_bvh* blas = nullptr;
    uint64_t triCount = 0;
    float4* tris = nullptr;       // GPUBVHTriangle, 3 x float4 each
    float4* tris96 = nullptr;     // GPUTriangle, 6 x float4 each (only after atlas_rt_mesh_pack_shading)
};

// R8 texture on the device (opacity maps): texels row-major, width * height bytes.
struct TextureDev {
    const uint8_t* texels;
    uint32_t width, height;
};

// textureLod(sampler2D(...), uv, 0).r for an R8 texture: bilinear filter, repeat addressing, texel centres at (i + 0.5) / size.
// Weights are kept in full fp32 (graphics hardware filters with 8 fractional bits; the difference is below 1/256 of a texel
// step and documented in DESIGN.md).
__device__ __forceinline__ float sample_r8(const TextureDev& t, float u, float v) {
    const float x = __fsub_rn(__fmul_rn(u, float(t.width)), 0.5f), y = __fsub_rn(__fmul_rn(v, float(t.height)), 0.5f);
    const float fx = floorf(x), fy = floorf(y);
    const float wx = __fsub_rn(x, fx), wy = __fsub_rn(y, fy);
    auto wrap = [](float f, uint32_t n) { long long i = (long long)f % (long long)n; if (i < 0) i += n; return uint32_t(i); };
    const uint32_t x0 = wrap(fx, t.width), x1 = wrap(__fadd_rn(fx, 1.0f), t.width), y0 = wrap(fy, t.height), y1 = wrap(__fadd_rn(fy, 1.0f), t.height);
    auto tx = [&](uint32_t xx, uint32_t yy) { return __fdiv_rn(float(t.texels[size_t(yy) * t.width + xx]), 255.0f); };
    const float top = __fadd_rn(__fmul_rn(tx(x0, y0), __fsub_rn(1.0f, wx)), __fmul_rn(tx(x1, y0), wx));
    const float bot = __fadd_rn(__fmul_rn(tx(x0, y1), __fsub_rn(1.0f, wx)), __fmul_rn(tx(x1, y1), wx));
    return __fadd_rn(__fmul_rn(top, __fsub_rn(1.0f, wy)), __fmul_rn(bot, wy));
}

struct atlas_rt_scene {
    atlas_rt_context* ctx = nullptr;
    const atlas_rt_bvh* tlas = nullptr;
    uint64_t instanceCount = 0;   // == tlas->refCount
    uint32_t meshCount = 0;
    float4* instances = nullptr;             // reordered GPUBVHInstance, 4 x float4 each
    const float4** blasNodes = nullptr;      // device array [meshCount]
    const float4** bvhTris = nullptr;        // device array [meshCount]
    const float4** triangles = nullptr;      // device array [meshCount] of 96-byte triangle arrays (entries may be null)
    bool allShading = false;                 // every mesh has its 96-byte array: the opacity-aware variants may run
    int fastDivision = 0;                    // bit 0: all scene coordinates below 2^60 (slab tests may use div_by_rcp, trace.cu); bit 1: no non-zero box coordinate below 2^-60
    // material / texture tables (atlas_rt_scene_set_materials): textured opacity in traversal, shading in the path tracer
    uint32_t* materials = nullptr;           // RaytraceMaterial, 23 words each
    uint32_t materialCount = 0;
    TextureDev* textures = nullptr;
    uint32_t textureCount = 0;
    uint8_t* texelStorage = nullptr;
    // the meshes the scene was assembled from (borrowed unless also listed as owned): atlas_rt_scene_replicate walks them
    // every triangle of every 96-byte array has opacity exactly 1 (no textured opacity either): HitClosestTransparency then
    // accepts exactly what HitClosest accepts, so the path tracer may run the plain 48-byte variant with identical results;
    // if in addition every instance carries the shadow bit, HitAnyTransparency's result is 1 - hit of plain HitAny
    bool allOpaque = false, allShadowBit = false;
    const void* hotNodes = nullptr;          // the largest node array of the scene (L2 access-policy window of the trace launches)
    size_t hotBytes = 0;
    std::vector<const atlas_rt_mesh*> partMeshes;
    // objects created on the scene's behalf (atlas_rt_build_scene_sharded, atlas_rt_scene_replicate): freed with the scene
    std::vector<atlas_rt_mesh*> ownedMeshes;
    std::vector<atlas_rt_bvh*> ownedBvhs;
};

namespace atlas {

int fail(atlas_rt_context* ctx, int status, const char* what, cudaError_t e = cudaSuccess);
void ctx_retain(atlas_rt_context* ctx);
void ctx_release(atlas_rt_context* ctx);   // destroys the context when the last reference goes

#define ATLAS_CUDA(ctx, call)                                                          \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) return atlas::fail((ctx), ATLAS_RT_ERR_CUDA, #call, e__); \
    } while (0)

#define ATLAS_LAUNCH_CHECK(ctx)                                                        \
    do {                                                                               \
        (ctx)->launches++;                                                             \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) return atlas::fail((ctx), ATLAS_RT_ERR_CUDA, "kernel launch", e__); \
    } while (0)

// Stream-ordered allocation from the device's default pool (kept warm: release threshold = max).
template <typename T>
inline cudaError_t dev_alloc(atlas_rt_context* ctx, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMallocAsync(reinterpret_cast<void**>(p), count * sizeof(T), ctx->stream);
}
inline void dev_free(atlas_rt_context* ctx, const void* p) {
    if (p) cudaFreeAsync(const_cast<void*>(p), ctx->stream);
}
template <typename T>
inline cudaError_t dev_alloc_on(cudaStream_t st, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    return cudaMallocAsync(reinterpret_cast<void**>(p), count * sizeof(T), st);
}
inline void dev_free_on(cudaStream_t st, const void* p) {
    if (p) cudaFreeAsync(const_cast<void*>(p), st);
}

// Kernels that follow each other on a stream (the builder's level loop, the ray sort + trace) form a chain of
// programmatic dependent launches: each one lets its successor be
// scheduled right away and then waits until everything before it has completed (griddepcontrol.wait returns once the
// prerequisite grids have finished and their memory is visible), so launch latency overlaps the predecessor's work.
// Every kernel of the chain executes chain_begin() first, on every path, because a kernel that skipped the wait could
// finish before ITS predecessor and release the kernel after it too early. Launched normally, both are no-ops.
__device__ __forceinline__ void chain_begin() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// Optional L2 access-policy window of a launch (the hot node array of a scene): hits inside the window are kept as
// persisting lines, everything else streams.
struct L2Window {
    const void* base = nullptr;
    size_t bytes = 0;
    float hitRatio = 1.0f;
};

template <typename... P, typename... A>
cudaError_t launch_chain_w(bool pdl, const L2Window& win, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attrs[2]{};
    unsigned n = 0;
    if (pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        n++;
    }
    if (win.base && win.bytes) {
        attrs[n].id = cudaLaunchAttributeAccessPolicyWindow;
        attrs[n].val.accessPolicyWindow.base_ptr = const_cast<void*>(win.base);
        attrs[n].val.accessPolicyWindow.num_bytes = win.bytes;
        attrs[n].val.accessPolicyWindow.hitRatio = win.hitRatio;
        attrs[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attrs[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        n++;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

template <typename... P, typename... A>
cudaError_t launch_chain(bool pdl, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
    return launch_chain_w(pdl, L2Window{}, kernel, grid, block, smem, st, std::forward<A>(args)...);
}

int launch_release_chunks(atlas_rt_context* ctx, unsigned int* chunkDone, uint32_t chunkRays, uint32_t count, uint32_t chunks);

// Copy count*bytes from src (host or device according to `device`) into device memory on the context stream.
cudaError_t copy_in(atlas_rt_context* ctx, void* dst, const void* src, size_t bytes, bool srcDevice);
cudaError_t copy_out(atlas_rt_context* ctx, void* dst, const void* src, size_t bytes, bool dstDevice);

int ensure_node_storage(atlas_rt_context* ctx, atlas_rt_bvh* bvh);

// builder entry points (build.cu)
int build_init_device(atlas_rt_context* ctx);   // per-device kernel attributes (opt-in shared memory); called by context_create
int build_bvh(atlas_rt_context* ctx, const float* dAabbs, const float* dTris, uint64_t count, bool tlas, atlas_rt_bvh* out);

// traversal entry points (trace.cu)
int scene_fast_flag(atlas_rt_context* ctx, atlas_rt_scene* scene, const uint32_t* dNodeCounts);
int scene_opacity_flags(atlas_rt_context* ctx, atlas_rt_scene* scene, const uint64_t* triCounts);
int launch_trace(atlas_rt_context* ctx, const atlas_rt_scene* scene, const float4* dIn, float4* dOut, uint64_t count,
                 uint32_t cullMask, float tMin, float tMax, bool any, bool perRayTMax, bool counters, bool resetCounters = true,
                 bool opacity = false, cudaStream_t st = nullptr /* context stream */, int queueSlot = 0,
                 const uint32_t* dCount = nullptr /* batch size on the device (<= count) */, bool hitsOnly = false /* dOut = 16-byte hit records */,
                 const unsigned int* watermark = nullptr /* streaming input: rays uploaded so far */, unsigned int* chunkDone = nullptr,
                 uint32_t chunkRays = 0, const uint32_t* streamPerm = nullptr /* streaming: fetch order, written chunk by chunk behind the watermark */,
                 int chain = -1 /* kernels of this call as programmatic dependent launches: 1 / 0, -1 = the context's setting */);
int launch_chunk_sort(atlas_rt_context* ctx, const atlas_rt_scene* scene, cudaStream_t st, const float4* rays, uint32_t n, uint32_t indexBase,
                      uint8_t* bucketOf, unsigned int* hist, uint32_t* perm, unsigned int* watermark);

}   // namespace atlas
