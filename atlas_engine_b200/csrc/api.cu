// api.cu — C ABI glue (include/atlas_rt.h): contexts, object lifetime, host<->device staging, pack and scene kernels.
#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

namespace atlas {

int fail(atlas_rt_context* ctx, int status, const char* what, cudaError_t e) {
    if (ctx) {
        ctx->error = what ? what : "error";
        if (e != cudaSuccess) {
            ctx->error += ": ";
            ctx->error += cudaGetErrorString(e);
        }
    }
    return status;
}

void ctx_retain(atlas_rt_context* ctx) { ctx->refs.fetch_add(1, std::memory_order_relaxed); }

void ctx_release(atlas_rt_context* ctx) {
    if (ctx->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    for (auto& w : ctx->workers) if (w) { atlas_rt_context_destroy(w); w = nullptr; }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // calls that were not ordered into the context stream (pipelined host-buffer traces) may still be using the other streams
    if (ctx->copyIn) cudaStreamSynchronize(ctx->copyIn);
    if (ctx->copyOut) cudaStreamSynchronize(ctx->copyOut);
    for (auto& cs : ctx->computeExtra) if (cs) cudaStreamSynchronize(cs);
    if (ctx->sortStream) cudaStreamSynchronize(ctx->sortStream);
    for (auto& ev : ctx->pipeEvents) if (ev) cudaEventDestroy(ev);
    if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
    if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
    for (auto& cs : ctx->computeExtra) if (cs) cudaStreamDestroy(cs);
    if (ctx->sortStream) cudaStreamDestroy(ctx->sortStream);
    for (int k = 0; k < 2; k++) { cudaFree(ctx->stageIn[k]); cudaFree(ctx->stageOut[k]); }
    cudaFree(ctx->dCounters);
    cudaFree(ctx->dStreamState);
    cudaFreeHost(ctx->pinned);
    if (ctx->ownStream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

cudaError_t copy_in(atlas_rt_context* ctx, void* dst, const void* src, size_t bytes, bool srcDevice) {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, srcDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
}

cudaError_t copy_out(atlas_rt_context* ctx, void* dst, const void* src, size_t bytes, bool dstDevice) {
    if (bytes == 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, dstDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream);
}

namespace {

constexpr int kBlock = 256;
inline uint32_t grid_for(uint64_t n, int block = kBlock) { return uint32_t((n + block - 1) / block); }

// BVHNode (56 B, 14 words) <-> GPUBVHNode (64 B, 16 words) — the field-by-field copy of mesh/MeshData.cpp:256-268 and
// raytracing/RayTracingWorld.cpp:272-284.
__global__ void nodes56_to_64(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t nodes) {
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= nodes * 16) return;
    const uint64_t n = i >> 4;
    const uint32_t w = uint32_t(i & 15);
    out[i] = w < 14 ? in[n * 14 + w] : 0u;
}
__global__ void nodes64_to_56(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t nodes) {
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= nodes * 14) return;
    const uint64_t n = i / 14;
    const uint32_t w = uint32_t(i - n * 14);
    out[i] = in[n * 16 + w];
}

// GPUBVHTriangle records in flattened order — mesh/MeshData.cpp:242-247.
__global__ void pack_bvh_triangles(const float* __restrict__ tris, const uint32_t* __restrict__ order,
                                   const uint8_t* __restrict__ endOfNode, const int32_t* __restrict__ material,
                                   const float* __restrict__ opacity, float4* __restrict__ out, uint64_t refs) {
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= refs) return;
    const uint32_t src = order[i];
    const float* t = tris + 9 * size_t(src);
    const int32_t mat = material ? material[src] : 0;
    const float op = opacity ? opacity[src] : 1.0f;
    out[3 * i + 0] = make_float4(t[0], t[1], t[2], endOfNode[i] ? 1.0f : -1.0f);
    out[3 * i + 1] = make_float4(t[3], t[4], t[5], __int_as_float(mat));
    out[3 * i + 2] = make_float4(t[6], t[7], t[8], op);
}

// GPUTriangle records in flattened order — mesh/MeshData.cpp:230-239; the 11 shading words per triangle are the
// engine's own packed values and are only moved.
__global__ void pack_shading_triangles(const float* __restrict__ tris, const uint32_t* __restrict__ order,
                                       const uint8_t* __restrict__ endOfNode, const int32_t* __restrict__ material,
                                       const float* __restrict__ opacity, const uint32_t* __restrict__ payload,
                                       float4* __restrict__ out, uint64_t refs) {
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= refs) return;
    const uint32_t src = order[i];
    const float* t = tris + 9 * size_t(src);
    uint32_t p[11];
#pragma unroll
    for (int k = 0; k < 11; k++) p[k] = payload ? payload[11 * size_t(src) + k] : 0u;
    const int32_t mat = material ? material[src] : 0;
    const float op = opacity ? opacity[src] : 1.0f;
    out[6 * i + 0] = make_float4(t[0], t[1], t[2], __uint_as_float(p[0]));
    out[6 * i + 1] = make_float4(t[3], t[4], t[5], __uint_as_float(p[1]));
    out[6 * i + 2] = make_float4(t[6], t[7], t[8], __uint_as_float(p[2]));
    out[6 * i + 3] = make_float4(__uint_as_float(p[3]), __uint_as_float(p[4]), __uint_as_float(p[5]), __int_as_float(mat));
    out[6 * i + 4] = make_float4(__uint_as_float(p[6]), __uint_as_float(p[7]), endOfNode[i] ? 1.0f : -1.0f, 0.0f);
    out[6 * i + 5] = make_float4(__uint_as_float(p[8]), __uint_as_float(p[9]), __uint_as_float(p[10]), op);
}

// Instance permutation of RayTracingWorld.cpp:287-295.
__global__ void reorder_instances(const float4* __restrict__ src, const uint32_t* __restrict__ order,
                                  const uint8_t* __restrict__ endOfNode, float4* __restrict__ dst, uint64_t refs) {
    const uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (i >= refs) return;
    const float4* s = src + 4 * size_t(order[i]);
    float4 last = s[3];
    last.z = __int_as_float(endOfNode[i] ? -1 : int(i) + 1);
    dst[4 * i + 0] = s[0];
    dst[4 * i + 1] = s[1];
    dst[4 * i + 2] = s[2];
    dst[4 * i + 3] = last;
}

// A BLAS whose root became a leaf has no nodes at all (Flatten emits none, BVH.cpp:411-417); the reference shader would
// then read blasNodes[..].data[0] out of bounds. Device storage for such a BLAS gets one synthetic node that is NOT part
// of the reported tree (nodeCount stays 0): its left child is the leaf starting at slot 0 inside a box that every ray
// hits, its right child a point box at +FLT_MAX that no ray hits. Rays entering the instance therefore test the leaf's
// triangles (up to endOfNode) and return correct hits instead of undefined behaviour.
__global__ void write_root_leaf_node(float4* node) {
    node[0] = make_float4(-kFltMax, -kFltMax, -kFltMax, kFltMax);
    node[1] = make_float4(kFltMax, kFltMax, kFltMax, kFltMax);
    node[2] = make_float4(kFltMax, kFltMax, kFltMax, kFltMax);
    node[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.0f, 0.0f);
}

int sync_unless_async(atlas_rt_context* ctx, uint32_t flags) {
    if (flags & ATLAS_RT_ASYNC) return ATLAS_RT_OK;
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

}   // namespace

int ensure_node_storage(atlas_rt_context* ctx, atlas_rt_bvh* bvh) {
    if (bvh->nodeCount != 0) return ATLAS_RT_OK;
    if (!bvh->nodes) ATLAS_CUDA(ctx, dev_alloc(ctx, &bvh->nodes, 4));
    write_root_leaf_node<<<1, 1, 0, ctx->stream>>>(bvh->nodes);
    ATLAS_LAUNCH_CHECK(ctx);
    return ATLAS_RT_OK;
}

}   // namespace atlas

using namespace atlas;

extern "C" {

int atlas_rt_version(void) { return ATLAS_RT_VERSION; }

int atlas_rt_context_create(int device, void* stream, atlas_rt_context** out_ctx) {
    if (!out_ctx) return ATLAS_RT_ERR_INVALID;
    *out_ctx = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ATLAS_RT_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return ATLAS_RT_ERR_CUDA;
    auto* ctx = new (std::nothrow) atlas_rt_context;
    if (!ctx) return ATLAS_RT_ERR_OOM;
    ctx->device = device;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ATLAS_RT_ERR_CUDA; }
        ctx->ownStream = true;
    }
    cudaDeviceGetAttribute(&ctx->smCount, cudaDevAttrMultiProcessorCount, device);
    if (const char* e = getenv("ATLAS_RT_TRACE_LEAF_THRESHOLD")) ctx->traceLeafThreshold = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_TRACE_REFILL_THRESHOLD")) ctx->traceRefillThreshold = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_TRACE_RAYS_PER_WARP")) ctx->traceRaysPerWarp = std::max(1, atoi(e));
    if (const char* e = getenv("ATLAS_RT_TRACE_LONGEST_FIRST")) ctx->traceLongestFirst = atoi(e);
    if (const char* e = getenv("ATLAS_RT_TRACE_BLOCKS_PER_SM")) ctx->traceBlocksPerSM = std::max(1, std::min(9, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_BIN_CTAS_PER_SM")) ctx->binCtasPerSM = std::max(1, std::min(8, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_CHAIN_LAUNCH")) ctx->chainLaunch = atoi(e);
    if (const char* e = getenv("ATLAS_RT_BUILD_WIDE")) ctx->buildWide = atoi(e);
    if (const char* e = getenv("ATLAS_RT_BATCH_WORKERS")) ctx->batchWorkers = std::max(1, std::min(16, atoi(e)));
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (cudaMalloc(&ctx->dCounters, 16 * sizeof(unsigned long long)) != cudaSuccess) { delete ctx; return ATLAS_RT_ERR_OOM; }
    cudaMemset(ctx->dCounters, 0, 16 * sizeof(unsigned long long));
    if (cudaMalloc(&ctx->dStreamState, 64 * sizeof(unsigned int)) != cudaSuccess) { cudaFree(ctx->dCounters); delete ctx; return ATLAS_RT_ERR_OOM; }
    {   // cuStreamWaitValue32 lets the download stream wait for a chunk's completion count without the host (streaming trace)
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) ctx->waitValue32 = fn;
        else cudaGetLastError();
        fn = nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) ctx->writeValue32 = fn;
        else cudaGetLastError();
    }
    if (const char* e = getenv("ATLAS_RT_TRACE_STREAMING")) ctx->traceStreaming = atoi(e);
    if (const char* e = getenv("ATLAS_RT_STREAM_SPIN_LOG2")) ctx->streamSpinLog2 = std::max(8, std::min(28, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_L2_PERSIST_MB")) ctx->l2PersistMB = std::max(0, atoi(e));
    if (const char* e = getenv("ATLAS_RT_L2_HIT_RATIO")) ctx->l2HitRatio = float(atof(e));
    if (ctx->l2PersistMB > 0) {   // carve persisting lines out of the L2 for the node array a trace launch declares hot
        int maxPersist = 0, maxWindow = 0;
        cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, device);
        cudaDeviceGetAttribute(&maxWindow, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t want = std::min<size_t>(size_t(ctx->l2PersistMB) << 20, size_t(maxPersist));
        if (want == 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) { cudaGetLastError(); ctx->l2PersistMB = 0; }
        ctx->l2WindowMax = size_t(maxWindow);
    }
    ctx->pinnedBytes = 8192;
    if (cudaMallocHost(&ctx->pinned, ctx->pinnedBytes) != cudaSuccess) { cudaFree(ctx->dCounters); delete ctx; return ATLAS_RT_ERR_OOM; }
    ctx->levelSlots = static_cast<char*>(ctx->pinned) + 4096;
    if (cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking) != cudaSuccess) ctx->copyIn = nullptr;
    if (cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking) != cudaSuccess) ctx->copyOut = nullptr;
    for (auto& cs : ctx->computeExtra)
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) cs = nullptr;
    {
        int prLow = 0, prHigh = 0;
        cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
        if (cudaStreamCreateWithPriority(&ctx->sortStream, cudaStreamNonBlocking, prHigh) != cudaSuccess) ctx->sortStream = nullptr;
    }
    if (const char* e = getenv("ATLAS_RT_PT_LANES")) ctx->ptLanes = std::max(1, std::min(8, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_STREAM_BLOCKS_PER_SM")) ctx->streamBlocksPerSM = std::max(1, std::min(8, atoi(e)));
    if (const char* e = getenv("ATLAS_RT_TRACE_MIN_BLOCKS_PER_SM")) ctx->traceMinBlocksPerSM = std::max(1, std::min(9, atoi(e)));
    ctx->pipeTimeline = getenv("ATLAS_RT_PIPE_TIMELINE") != nullptr;
    if (const char* e = getenv("ATLAS_RT_PIPE_STREAMS")) ctx->pipeStreams = std::max(1, std::min(8, atoi(e)));
    for (auto& ev : ctx->pipeEvents)
        if (cudaEventCreateWithFlags(&ev, ctx->pipeTimeline ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; ctx->copyIn = nullptr; }
    if (build_init_device(ctx) != ATLAS_RT_OK) { ctx_release(ctx); return ATLAS_RT_ERR_CUDA; }
    *out_ctx = ctx;
    return ATLAS_RT_OK;
}

void atlas_rt_context_destroy(atlas_rt_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx_release(ctx);   // objects that are still alive keep the context (and its stream) until they are freed
}

int atlas_rt_context_synchronize(atlas_rt_context* ctx) {
    if (!ctx) return ATLAS_RT_ERR_INVALID;
    if (ctx->pendingJoin) { const int rc = atlas_rt_trace_join(ctx); if (rc != ATLAS_RT_OK) return rc; }   // pipelined host-buffer traces in flight
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

const char* atlas_rt_last_error(const atlas_rt_context* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
uint64_t atlas_rt_kernel_launches(const atlas_rt_context* ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------------------------- build
static atlas_rt_bvh* new_bvh(atlas_rt_context* ctx) {
    auto* bvh = new (std::nothrow) atlas_rt_bvh;
    if (bvh) { bvh->ctx = ctx; ctx_retain(ctx); }
    return bvh;
}

static int build_common(atlas_rt_context* ctx, const float* aabbs, const float* tris, uint64_t count, uint32_t flags,
                        bool tlas, atlas_rt_bvh** out_bvh) {
    if (!ctx || !out_bvh || (count && !aabbs) || (!tlas && count && !tris)) return fail(ctx, ATLAS_RT_ERR_INVALID, "null argument");
    *out_bvh = nullptr;
    if (count > 0x3fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^30-1 primitives");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_INPUT;
    float *dA = nullptr, *dT = nullptr;
    atlas_rt_bvh* bvh = nullptr;
    auto done = [&](int rc) {
        dev_free(ctx, dA);
        dev_free(ctx, dT);
        if (rc != ATLAS_RT_OK) atlas_rt_bvh_free(bvh); else *out_bvh = bvh;
        return rc;
    };
#define ATLAS_TRY_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, #call, e__)); } while (0)
    if (!dev) {
        ATLAS_TRY_CUDA(dev_alloc(ctx, &dA, count * 6));
        ATLAS_TRY_CUDA(copy_in(ctx, dA, aabbs, count * 24, false));
        if (!tlas) {
            ATLAS_TRY_CUDA(dev_alloc(ctx, &dT, count * 9));
            ATLAS_TRY_CUDA(copy_in(ctx, dT, tris, count * 36, false));
        }
    }
    bvh = new_bvh(ctx);
    if (!bvh) return done(fail(ctx, ATLAS_RT_ERR_OOM, "host allocation"));
    int rc = build_bvh(ctx, dev ? aabbs : dA, tlas ? nullptr : (dev ? tris : dT), count, tlas, bvh);
    if (rc == ATLAS_RT_OK) rc = ensure_node_storage(ctx, bvh);
    if (rc == ATLAS_RT_OK) rc = sync_unless_async(ctx, flags);
    return done(rc);
}

int atlas_rt_build_blas(atlas_rt_context* ctx, const float* aabbs, const float* tris, uint64_t count, uint32_t flags,
                        atlas_rt_bvh** out_bvh) {
    return build_common(ctx, aabbs, tris, count, flags, false, out_bvh);
}

int atlas_rt_build_tlas(atlas_rt_context* ctx, const float* aabbs, uint64_t count, uint32_t flags, atlas_rt_bvh** out_bvh) {
    return build_common(ctx, aabbs, nullptr, count, flags, true, out_bvh);
}

// Many BLASes at once. The reference builds a scene's meshes concurrently from job-system workers (src/tests/App.cpp:362-370,
// MeshData::BuildBVH per mesh); one GPU build of a small mesh is a chain of short, latency-bound launches that leaves most
// of the device idle, so the batch runs up to 8 builds side by side, each on its own worker context (stream, pinned level
// flags) driven by its own host thread, largest mesh first. Results are the same objects atlas_rt_build_blas returns.
int atlas_rt_build_blas_batch(atlas_rt_context* ctx, uint32_t mesh_count, const float* const* aabbs, const float* const* tris,
                              const uint64_t* counts, uint32_t flags, atlas_rt_bvh** out_bvhs) {
    if (!ctx || !out_bvhs || (mesh_count && (!aabbs || !tris || !counts))) return fail(ctx, ATLAS_RT_ERR_INVALID, "null argument");
    for (uint32_t m = 0; m < mesh_count; m++) out_bvhs[m] = nullptr;
    if (mesh_count == 0) return ATLAS_RT_OK;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t nWorkers = std::min<uint32_t>(uint32_t(ctx->batchWorkers), mesh_count);
    for (uint32_t w = 0; w < nWorkers; w++) {
        if (!ctx->workers[w]) {
            const int rc = atlas_rt_context_create(ctx->device, nullptr, &ctx->workers[w]);
            if (rc != ATLAS_RT_OK) return fail(ctx, rc, "worker context");
        }
    }
    // inputs produced on the context's stream (device pointers) must be complete before the workers read them
    cudaEvent_t ready = ctx->pipeEvents[0];
    ATLAS_CUDA(ctx, cudaEventRecord(ready, ctx->stream));
    std::vector<uint32_t> bySize(mesh_count);
    std::iota(bySize.begin(), bySize.end(), 0u);
    std::stable_sort(bySize.begin(), bySize.end(), [&](uint32_t a, uint32_t b) { return counts[a] > counts[b]; });
    std::atomic<uint32_t> next{0};
    std::vector<int> status(nWorkers, ATLAS_RT_OK);
    auto work = [&](uint32_t w) {
        atlas_rt_context* wc = ctx->workers[w];
        if (cudaSetDevice(wc->device) != cudaSuccess || cudaStreamWaitEvent(wc->stream, ready, 0) != cudaSuccess) { status[w] = ATLAS_RT_ERR_CUDA; return; }
        for (;;) {
            const uint32_t k = next.fetch_add(1);
            if (k >= mesh_count) break;
            const uint32_t m = bySize[k];
            const int rc = build_common(wc, aabbs[m], tris[m], counts[m], flags | ATLAS_RT_ASYNC, false, &out_bvhs[m]);
            if (rc != ATLAS_RT_OK && status[w] == ATLAS_RT_OK) status[w] = rc;
        }
    };
    std::vector<std::thread> pool;
    for (uint32_t w = 1; w < nWorkers; w++) pool.emplace_back(work, w);
    work(0);
    for (auto& t : pool) t.join();
    int rc = ATLAS_RT_OK;
    for (uint32_t w = 0; w < nWorkers; w++) {
        atlas_rt_context* wc = ctx->workers[w];
        ctx->launches += wc->launches;
        wc->launches = 0;
        if (status[w] != ATLAS_RT_OK && rc == ATLAS_RT_OK) { rc = status[w]; ctx->error = "batch build: " + wc->error; }
        // the context's stream continues after every worker's builds
        cudaEvent_t ev = ctx->pipeEvents[1 + w];
        if (cudaEventRecord(ev, wc->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ev, 0) != cudaSuccess) { if (rc == ATLAS_RT_OK) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "batch join"); }
    }
    if (rc == ATLAS_RT_OK) rc = sync_unless_async(ctx, flags);
    if (rc != ATLAS_RT_OK) {
        for (uint32_t m = 0; m < mesh_count; m++) { atlas_rt_bvh_free(out_bvhs[m]); out_bvhs[m] = nullptr; }
    }
    return rc;
}

int atlas_rt_bvh_import(atlas_rt_context* ctx, const void* nodes56, uint64_t node_count, const uint32_t* order,
                        const uint8_t* end_of_node, uint64_t ref_count, uint32_t flags, atlas_rt_bvh** out_bvh) {
    if (!ctx || !out_bvh || (node_count && !nodes56) || (ref_count && (!order || !end_of_node))) return fail(ctx, ATLAS_RT_ERR_INVALID, "null argument");
    *out_bvh = nullptr;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_INPUT;
    atlas_rt_bvh* bvh = new_bvh(ctx);
    if (!bvh) return fail(ctx, ATLAS_RT_ERR_OOM, "host allocation");
    bvh->nodeCount = node_count;
    bvh->refCount = ref_count;
    uint32_t* staging = nullptr;
    auto done = [&](int rc) {
        dev_free(ctx, staging);
        if (rc != ATLAS_RT_OK) atlas_rt_bvh_free(bvh); else *out_bvh = bvh;
        return rc;
    };
    ATLAS_TRY_CUDA(dev_alloc(ctx, &bvh->nodes, (node_count ? node_count : 1) * 4));
    ATLAS_TRY_CUDA(dev_alloc(ctx, &bvh->order, ref_count));
    ATLAS_TRY_CUDA(dev_alloc(ctx, &bvh->endOfNode, ref_count));
    ATLAS_TRY_CUDA(dev_alloc(ctx, &staging, node_count * 14));
    ATLAS_TRY_CUDA(copy_in(ctx, staging, nodes56, node_count * 56, dev));
    ATLAS_TRY_CUDA(copy_in(ctx, bvh->order, order, ref_count * 4, dev));
    ATLAS_TRY_CUDA(copy_in(ctx, bvh->endOfNode, end_of_node, ref_count, dev));
    if (node_count) {
        nodes56_to_64<<<grid_for(node_count * 16), kBlock, 0, ctx->stream>>>(staging, reinterpret_cast<uint32_t*>(bvh->nodes), node_count);
        ctx->launches++;
        ATLAS_TRY_CUDA(cudaGetLastError());
    }
    int rc = ensure_node_storage(ctx, bvh);
    if (rc == ATLAS_RT_OK) ATLAS_TRY_CUDA(cudaStreamSynchronize(ctx->stream));
    return done(rc);
}

int atlas_rt_bvh_upload(atlas_rt_context* ctx, const void* nodes56, uint64_t node_count, const uint32_t* order,
                        const uint8_t* end_of_node, uint64_t ref_count, atlas_rt_bvh** out_bvh) {
    return atlas_rt_bvh_import(ctx, nodes56, node_count, order, end_of_node, ref_count, 0, out_bvh);
}

int atlas_rt_bvh_counts(const atlas_rt_bvh* bvh, uint64_t* node_count, uint64_t* ref_count) {
    if (!bvh) return ATLAS_RT_ERR_INVALID;
    if (node_count) *node_count = bvh->nodeCount;
    if (ref_count) *ref_count = bvh->refCount;
    return ATLAS_RT_OK;
}

int atlas_rt_bvh_download(const atlas_rt_bvh* bvh, void* nodes56, uint32_t* order, uint8_t* end_of_node, uint32_t flags) {
    if (!bvh) return ATLAS_RT_ERR_INVALID;
    atlas_rt_context* ctx = bvh->ctx;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_OUTPUT;
    if (nodes56 && bvh->nodeCount) {
        uint32_t* staging = nullptr;
        ATLAS_CUDA(ctx, dev_alloc(ctx, &staging, bvh->nodeCount * 14));
        nodes64_to_56<<<grid_for(bvh->nodeCount * 14), kBlock, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(bvh->nodes), staging, bvh->nodeCount);
        ATLAS_LAUNCH_CHECK(ctx);
        ATLAS_CUDA(ctx, copy_out(ctx, nodes56, staging, bvh->nodeCount * 56, dev));
        dev_free(ctx, staging);
    }
    if (order) ATLAS_CUDA(ctx, copy_out(ctx, order, bvh->order, bvh->refCount * 4, dev));
    if (end_of_node) ATLAS_CUDA(ctx, copy_out(ctx, end_of_node, bvh->endOfNode, bvh->refCount, dev));
    return sync_unless_async(ctx, flags);
}

int atlas_rt_bvh_device_ptrs(const atlas_rt_bvh* bvh, const void** gpu_nodes64, const uint32_t** order, const uint8_t** end_of_node) {
    if (!bvh) return ATLAS_RT_ERR_INVALID;
    if (gpu_nodes64) *gpu_nodes64 = bvh->nodes;
    if (order) *order = bvh->order;
    if (end_of_node) *end_of_node = bvh->endOfNode;
    return ATLAS_RT_OK;
}

int atlas_rt_bvh_stats(const atlas_rt_bvh* bvh, uint64_t out[8]) {
    if (!bvh || !out) return ATLAS_RT_ERR_INVALID;
    memcpy(out, bvh->stats, sizeof(bvh->stats));
    return ATLAS_RT_OK;
}

void atlas_rt_bvh_free(atlas_rt_bvh* bvh) {
    if (!bvh) return;
    atlas_rt_context* ctx = bvh->ctx;
    cudaSetDevice(ctx->device);
    if (ctx->pendingJoin) atlas_rt_trace_join(ctx);   // pipelined traces in flight may still read it: the frees below are ordered on the context stream
    dev_free(ctx, bvh->nodes);
    dev_free(ctx, bvh->order);
    dev_free(ctx, bvh->endOfNode);
    delete bvh;
    ctx_release(ctx);
}

// ----------------------------------------------------------------------------------------------------- pack
// Objects may be used from any context on the same device (device memory is shared; only stream order differs): a mesh
// built on a worker thread's context can be packed into a scene on the main thread's. The object must be complete, i.e.
// created without ATLAS_RT_ASYNC or its context synchronised since.
static inline bool same_device(const atlas_rt_context* a, const atlas_rt_context* b) { return a && b && a->device == b->device; }

// Host arrays staged on the device for one call; freed (stream-ordered) when the helper goes out of scope.
struct Staged {
    atlas_rt_context* ctx;
    std::vector<void*> ptrs;
    explicit Staged(atlas_rt_context* c) : ctx(c) {}
    ~Staged() { for (void* p : ptrs) dev_free(ctx, p); }
    // returns the device pointer for `src` (host unless dev), or nullptr with *err set
    const void* in(const void* src, size_t bytes, bool dev, cudaError_t* err) {
        if (!src || dev) return src;
        void* d = nullptr;
        *err = cudaMallocAsync(&d, bytes ? bytes : 1, ctx->stream);
        if (*err != cudaSuccess) return nullptr;
        ptrs.push_back(d);
        *err = copy_in(ctx, d, src, bytes, false);
        return *err == cudaSuccess ? d : nullptr;
    }
};

int atlas_rt_pack_mesh(atlas_rt_context* ctx, const atlas_rt_bvh* blas, const float* tris, uint64_t count,
                       const int32_t* material_idx, const float* opacity, uint32_t flags, atlas_rt_mesh** out_mesh) {
    if (!ctx || !blas || !out_mesh || (count && !tris) || !same_device(blas->ctx, ctx)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    *out_mesh = nullptr;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_INPUT;
    Staged st(ctx);
    cudaError_t e = cudaSuccess;
    const float* dT = static_cast<const float*>(st.in(tris, count * 36, dev, &e));
    const int32_t* dM = e == cudaSuccess ? static_cast<const int32_t*>(st.in(material_idx, count * 4, dev, &e)) : nullptr;
    const float* dO = e == cudaSuccess ? static_cast<const float*>(st.in(opacity, count * 4, dev, &e)) : nullptr;
    if (e != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "staging the triangles", e);
    auto* mesh = new (std::nothrow) atlas_rt_mesh;
    if (!mesh) return fail(ctx, ATLAS_RT_ERR_OOM, "host allocation");
    mesh->ctx = ctx;
    ctx_retain(ctx);
    mesh->blas = blas;
    mesh->triCount = blas->refCount;
    e = dev_alloc(ctx, &mesh->tris, mesh->triCount * 3);
    if (e == cudaSuccess && mesh->triCount) {
        pack_bvh_triangles<<<grid_for(mesh->triCount), kBlock, 0, ctx->stream>>>(dT, blas->order, blas->endOfNode, dM, dO, mesh->tris, mesh->triCount);
        ctx->launches++;
        e = cudaGetLastError();
    }
    int rc = e == cudaSuccess ? sync_unless_async(ctx, flags) : fail(ctx, ATLAS_RT_ERR_CUDA, "pack_bvh_triangles", e);
    if (rc != ATLAS_RT_OK) { atlas_rt_mesh_free(mesh); return rc; }
    *out_mesh = mesh;
    return ATLAS_RT_OK;
}

int atlas_rt_mesh_pack_shading(atlas_rt_context* ctx, atlas_rt_mesh* mesh, const float* tris, uint64_t count,
                               const int32_t* material_idx, const float* opacity, const uint32_t* payload11, uint32_t flags) {
    if (!ctx || !mesh || !same_device(mesh->ctx, ctx) || (count && !tris)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_INPUT;
    Staged st(ctx);
    cudaError_t e = cudaSuccess;
    const float* dT = static_cast<const float*>(st.in(tris, count * 36, dev, &e));
    const int32_t* dM = e == cudaSuccess ? static_cast<const int32_t*>(st.in(material_idx, count * 4, dev, &e)) : nullptr;
    const float* dO = e == cudaSuccess ? static_cast<const float*>(st.in(opacity, count * 4, dev, &e)) : nullptr;
    const uint32_t* dP = e == cudaSuccess ? static_cast<const uint32_t*>(st.in(payload11, count * 44, dev, &e)) : nullptr;
    if (e != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "staging the triangles", e);
    if (!mesh->tris96) ATLAS_CUDA(ctx, dev_alloc(ctx, &mesh->tris96, mesh->triCount * 6));
    if (mesh->triCount) {
        pack_shading_triangles<<<grid_for(mesh->triCount), kBlock, 0, ctx->stream>>>(dT, mesh->blas->order, mesh->blas->endOfNode, dM, dO, dP,
                                                                                      mesh->tris96, mesh->triCount);
        ATLAS_LAUNCH_CHECK(ctx);
    }
    return sync_unless_async(ctx, flags);
}

int atlas_rt_mesh_download_shading(const atlas_rt_mesh* mesh, void* gpu_triangles96, uint32_t flags) {
    if (!mesh || !gpu_triangles96) return ATLAS_RT_ERR_INVALID;
    atlas_rt_context* ctx = mesh->ctx;
    if (!mesh->tris96) return fail(ctx, ATLAS_RT_ERR_INVALID, "atlas_rt_mesh_pack_shading has not been called for this mesh");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    ATLAS_CUDA(ctx, copy_out(ctx, gpu_triangles96, mesh->tris96, mesh->triCount * 96, (flags & ATLAS_RT_DEVICE_OUTPUT) != 0));
    return sync_unless_async(ctx, flags);
}

int atlas_rt_mesh_counts(const atlas_rt_mesh* mesh, uint64_t* node_count, uint64_t* triangle_count) {
    if (!mesh) return ATLAS_RT_ERR_INVALID;
    if (node_count) *node_count = mesh->blas->nodeCount;
    if (triangle_count) *triangle_count = mesh->triCount;
    return ATLAS_RT_OK;
}

int atlas_rt_mesh_download(const atlas_rt_mesh* mesh, void* gpu_nodes64, void* gpu_bvh_triangles48, uint32_t flags) {
    if (!mesh) return ATLAS_RT_ERR_INVALID;
    atlas_rt_context* ctx = mesh->ctx;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_OUTPUT;
    if (gpu_nodes64) ATLAS_CUDA(ctx, copy_out(ctx, gpu_nodes64, mesh->blas->nodes, mesh->blas->nodeCount * 64, dev));
    if (gpu_bvh_triangles48) ATLAS_CUDA(ctx, copy_out(ctx, gpu_bvh_triangles48, mesh->tris, mesh->triCount * 48, dev));
    return sync_unless_async(ctx, flags);
}

void atlas_rt_mesh_free(atlas_rt_mesh* mesh) {
    if (!mesh) return;
    cudaSetDevice(mesh->ctx->device);
    atlas_rt_context* ctx = mesh->ctx;
    if (ctx->pendingJoin) atlas_rt_trace_join(ctx);
    dev_free(ctx, mesh->tris);
    dev_free(ctx, mesh->tris96);
    delete mesh;
    ctx_release(ctx);
}

// ---------------------------------------------------------------------------------------------------- scene
int atlas_rt_scene_create(atlas_rt_context* ctx, const atlas_rt_mesh* const* meshes, uint32_t mesh_count,
                          const void* instances64, uint64_t instance_count, const atlas_rt_bvh* tlas, uint32_t flags,
                          atlas_rt_scene** out_scene) {
    if (!ctx || !meshes || !mesh_count || !instances64 || !tlas || !out_scene || !same_device(tlas->ctx, ctx)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    *out_scene = nullptr;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_INPUT;
    std::vector<const float4*> nodePtrs(mesh_count), triPtrs(mesh_count), tri96Ptrs(mesh_count);
    std::vector<uint32_t> nodeCounts(mesh_count);
    bool allShading = true;
    for (uint32_t m = 0; m < mesh_count; m++) {
        if (!meshes[m] || !same_device(meshes[m]->ctx, ctx)) return fail(ctx, ATLAS_RT_ERR_INVALID, "mesh missing or from another device");
        nodePtrs[m] = meshes[m]->blas->nodes;
        triPtrs[m] = meshes[m]->tris;
        tri96Ptrs[m] = meshes[m]->tris96;
        allShading = allShading && meshes[m]->tris96 != nullptr;
        nodeCounts[m] = uint32_t(meshes[m]->blas->nodeCount);
    }
    auto* scene = new (std::nothrow) atlas_rt_scene;
    if (!scene) return fail(ctx, ATLAS_RT_ERR_OOM, "host allocation");
    scene->ctx = ctx;
    ctx_retain(ctx);
    scene->tlas = tlas;
    scene->meshCount = mesh_count;
    scene->instanceCount = tlas->refCount;
    scene->allShading = allShading;
    scene->partMeshes.assign(meshes, meshes + mesh_count);
    scene->hotNodes = tlas->nodes;
    scene->hotBytes = size_t(tlas->nodeCount) * 64;
    for (uint32_t m = 0; m < mesh_count; m++)
        if (size_t(meshes[m]->blas->nodeCount) * 64 > scene->hotBytes) { scene->hotNodes = meshes[m]->blas->nodes; scene->hotBytes = size_t(meshes[m]->blas->nodeCount) * 64; }
    Staged st(ctx);
    cudaError_t e = cudaSuccess;
    const uint32_t* dNodeCounts = static_cast<const uint32_t*>(st.in(nodeCounts.data(), mesh_count * sizeof(uint32_t), false, &e));
    if (e == cudaSuccess) e = dev_alloc(ctx, &scene->blasNodes, mesh_count);
    if (e == cudaSuccess) e = dev_alloc(ctx, &scene->bvhTris, mesh_count);
    if (e == cudaSuccess) e = dev_alloc(ctx, &scene->triangles, mesh_count);
    if (e == cudaSuccess) e = dev_alloc(ctx, &scene->instances, scene->instanceCount * 4);
    // pointer tables are tiny: plain copies from pageable memory are fine here
    if (e == cudaSuccess) e = cudaMemcpyAsync(scene->blasNodes, nodePtrs.data(), mesh_count * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scene->bvhTris, triPtrs.data(), mesh_count * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scene->triangles, tri96Ptrs.data(), mesh_count * sizeof(void*), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // the std::vectors die at scope exit
    int rc = e == cudaSuccess ? scene_fast_flag(ctx, scene, dNodeCounts) : fail(ctx, ATLAS_RT_ERR_CUDA, "scene tables", e);
    if (rc == ATLAS_RT_OK) {
        const float4* src = static_cast<const float4*>(st.in(instances64, instance_count * 64, dev, &e));
        if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "staging the instances", e);
        else if (scene->instanceCount) {
            reorder_instances<<<grid_for(scene->instanceCount), kBlock, 0, ctx->stream>>>(src, tlas->order, tlas->endOfNode, scene->instances, scene->instanceCount);
            ctx->launches++;
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "reorder_instances", e);
        }
    }
    if (rc == ATLAS_RT_OK) {   // (synchronises) which of the opacity-aware traces of the path tracer can run as plain ones
        std::vector<uint64_t> triCounts(mesh_count);
        for (uint32_t m = 0; m < mesh_count; m++) triCounts[m] = meshes[m]->triCount;
        rc = scene_opacity_flags(ctx, scene, triCounts.data());
    }
    if (rc == ATLAS_RT_OK) rc = sync_unless_async(ctx, flags);
    if (rc != ATLAS_RT_OK) { atlas_rt_scene_free(scene); return rc; }
    *out_scene = scene;
    return ATLAS_RT_OK;
}

int atlas_rt_scene_set_materials(atlas_rt_context* ctx, atlas_rt_scene* scene, const atlas_rt_material* materials, uint32_t material_count,
                                 const atlas_rt_texture* textures, uint32_t texture_count) {
    static_assert(sizeof(atlas_rt_material) == 92, "RaytraceMaterial is 23 words");
    if (!ctx || !scene || !same_device(scene->ctx, ctx) || (material_count && !materials) || (texture_count && !textures)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // nothing may still be reading the old tables
    dev_free(ctx, scene->materials); dev_free(ctx, scene->textures); dev_free(ctx, scene->texelStorage);
    scene->materials = nullptr; scene->textures = nullptr; scene->texelStorage = nullptr;
    scene->materialCount = scene->textureCount = 0;
    size_t texelBytes = 0;
    for (uint32_t t = 0; t < texture_count; t++) {
        if (!textures[t].texels || !textures[t].width || !textures[t].height) return fail(ctx, ATLAS_RT_ERR_INVALID, "empty texture");
        texelBytes += (size_t(textures[t].width) * textures[t].height + 15) & ~size_t(15);
    }
    if (material_count) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &scene->materials, size_t(material_count) * 23));
        ATLAS_CUDA(ctx, cudaMemcpyAsync(scene->materials, materials, size_t(material_count) * 92, cudaMemcpyHostToDevice, ctx->stream));
        scene->materialCount = material_count;
    }
    std::vector<TextureDev> table(texture_count);
    if (texture_count) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &scene->texelStorage, texelBytes));
        ATLAS_CUDA(ctx, dev_alloc(ctx, &scene->textures, texture_count));
        size_t off = 0;
        for (uint32_t t = 0; t < texture_count; t++) {
            const size_t bytes = size_t(textures[t].width) * textures[t].height;
            ATLAS_CUDA(ctx, cudaMemcpyAsync(scene->texelStorage + off, textures[t].texels, bytes, cudaMemcpyHostToDevice, ctx->stream));
            table[t] = TextureDev{scene->texelStorage + off, textures[t].width, textures[t].height};
            off += (bytes + 15) & ~size_t(15);
        }
        ATLAS_CUDA(ctx, cudaMemcpyAsync(scene->textures, table.data(), texture_count * sizeof(TextureDev), cudaMemcpyHostToDevice, ctx->stream));
        scene->textureCount = texture_count;
    }
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host arrays (caller's and `table`) are free to go
    return ATLAS_RT_OK;
}

int atlas_rt_scene_download(const atlas_rt_scene* scene, void* instances64, void* tlas_nodes64, uint32_t flags) {
    if (!scene) return ATLAS_RT_ERR_INVALID;
    atlas_rt_context* ctx = scene->ctx;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool dev = flags & ATLAS_RT_DEVICE_OUTPUT;
    if (instances64) ATLAS_CUDA(ctx, copy_out(ctx, instances64, scene->instances, scene->instanceCount * 64, dev));
    if (tlas_nodes64) ATLAS_CUDA(ctx, copy_out(ctx, tlas_nodes64, scene->tlas->nodes, scene->tlas->nodeCount * 64, dev));
    return sync_unless_async(ctx, flags);
}

void atlas_rt_scene_free(atlas_rt_scene* scene) {
    if (!scene) return;
    cudaSetDevice(scene->ctx->device);
    atlas_rt_context* ctx = scene->ctx;
    if (ctx->pendingJoin) atlas_rt_trace_join(ctx);
    dev_free(ctx, scene->instances);
    dev_free(ctx, scene->blasNodes);
    dev_free(ctx, scene->bvhTris);
    dev_free(ctx, scene->triangles);
    dev_free(ctx, scene->materials);
    dev_free(ctx, scene->textures);
    dev_free(ctx, scene->texelStorage);
    for (atlas_rt_mesh* m : scene->ownedMeshes) atlas_rt_mesh_free(m);
    for (atlas_rt_bvh* b : scene->ownedBvhs) atlas_rt_bvh_free(b);
    delete scene;
    ctx_release(ctx);
}

// ---------------------------------------------------------------------------------------------------- trace
static int trace_common(atlas_rt_context* ctx, const atlas_rt_scene* scene, const void* rays_in, uint64_t count,
                        uint32_t cull_mask, float t_min, float t_max, void* rays_out, uint32_t flags, bool any) {
    if (!ctx || !scene || !same_device(scene->ctx, ctx) || (count && (!rays_in || !rays_out))) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool devIn = flags & ATLAS_RT_DEVICE_INPUT, devOut = flags & ATLAS_RT_DEVICE_OUTPUT;
    float4 *dIn = nullptr, *dOut = nullptr;
    bool pipelinedCall = false;
    const float4* in = static_cast<const float4*>(rays_in);
    float4* out = static_cast<float4*>(rays_out);
    // ATLAS_RT_HITS_ONLY: the output is one 16-byte hit record per ray (never in place on the 48-byte rays)
    const bool hitsOnly = (flags & ATLAS_RT_HITS_ONLY) != 0;
    const size_t outStride = hitsOnly ? 1 : 3;   // float4s per ray in the output
    // ATLAS_RT_PIPELINED calls use two persistent staging sets in turn (a stream-ordered allocation would tie this call's first
    // operation to the previous call's last download, through the pool, and undo the overlap)
    const bool wantPipelined = (flags & ATLAS_RT_PIPELINED) && (flags & ATLAS_RT_ASYNC) && !devIn && !devOut && count >= 262144 && ctx->copyIn && ctx->copyOut &&
                               !ctx->traceStreaming && !(flags & ATLAS_RT_COUNTERS);
    int stageSet = -1;
    if (wantPipelined) {
        stageSet = ctx->stageNext;
        ctx->stageNext ^= 1;
        const size_t needIn = size_t(count) * 48, needOut = hitsOnly ? size_t(count) * 16 : 0;
        if (ctx->stageInBytes[stageSet] < needIn || ctx->stageOutBytes[stageSet] < needOut) {   // (re)size: rare, synchronises
            ATLAS_CUDA(ctx, cudaDeviceSynchronize());
            if (ctx->stageInBytes[stageSet] < needIn) {
                cudaFree(ctx->stageIn[stageSet]); ctx->stageIn[stageSet] = nullptr; ctx->stageInBytes[stageSet] = 0;
                ATLAS_CUDA(ctx, cudaMalloc(&ctx->stageIn[stageSet], needIn));
                ctx->stageInBytes[stageSet] = needIn;
            }
            if (ctx->stageOutBytes[stageSet] < needOut) {
                cudaFree(ctx->stageOut[stageSet]); ctx->stageOut[stageSet] = nullptr; ctx->stageOutBytes[stageSet] = 0;
                ATLAS_CUDA(ctx, cudaMalloc(&ctx->stageOut[stageSet], needOut));
                ctx->stageOutBytes[stageSet] = needOut;
            }
        }
        // the call that used this set two calls ago must be through with it before the uploads overwrite it
        if (ctx->stageUsed[stageSet]) ATLAS_CUDA(ctx, cudaStreamWaitEvent(ctx->copyIn, ctx->pipeEvents[36 + stageSet], 0));
        in = static_cast<const float4*>(ctx->stageIn[stageSet]);
        out = hitsOnly ? static_cast<float4*>(ctx->stageOut[stageSet]) : static_cast<float4*>(ctx->stageIn[stageSet]);
    }
    float4* const stagedIn = stageSet >= 0 ? static_cast<float4*>(ctx->stageIn[stageSet]) : nullptr;
    if (!devIn && stageSet < 0) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &dIn, count * 3));
        in = dIn;
    }
    if (stagedIn) dIn = stagedIn;   // (not owned by this call: see the release below)
    if (!devOut && stageSet < 0) {
        if (dIn && !hitsOnly) out = dIn;   // in-place on the staging buffer
        else {
            const cudaError_t ea = dev_alloc(ctx, &dOut, count * outStride);
            if (ea != cudaSuccess) { dev_free(ctx, dIn); return fail(ctx, ATLAS_RT_ERR_CUDA, "output staging", ea); }
            out = dOut;
        }
    }
    const bool perRay = (flags & ATLAS_RT_PER_RAY_TMAX) != 0, counters = (flags & ATLAS_RT_COUNTERS) != 0;
    const bool opacity = (flags & ATLAS_RT_OPACITY) != 0;
    if (opacity && !scene->allShading) {
        if (stageSet < 0) { dev_free(ctx, dIn); dev_free(ctx, dOut); }
        return fail(ctx, ATLAS_RT_ERR_INVALID, "ATLAS_RT_OPACITY needs atlas_rt_mesh_pack_shading on every mesh of the scene (before atlas_rt_scene_create)");
    }
    int rc = ATLAS_RT_OK;
    const uint64_t kPipeMin = 262144;
    if (!devIn && count >= kPipeMin && count < 0x7fffffffull && ctx->copyIn && ctx->copyOut && ctx->sortStream && ctx->waitValue32 && ctx->traceStreaming &&
        !(flags & ATLAS_RT_COUNTERS) && scene->tlas->nodeCount > 0) {
        // Host input, streaming: ONE persistent launch traces the batch while it is still being uploaded. The upload stream
        // copies the rays in chunks; behind each chunk a high-priority stream puts the chunk's rays in longest-first order (the
        // three small ordering kernels of a resident batch) and then moves a watermark in device memory to the end of the
        // chunk; the persistent kernel (one CTA per SM fewer than a resident launch, so those small kernels find room) only
        // fetches rays below the watermark, through the permutation. Every warp reports the rays it has finished per chunk;
        // the download stream waits on each chunk's count with cuStreamWaitValue32 and sends that chunk's results home while
        // later chunks are still being traced. Against the chunked pipeline below (one launch per chunk, each with its own
        // ramp-up and drain) the lanes of one launch refill across chunk boundaries, and only the last chunk drains.
        // Submission order matters when streams share a hardware queue: nothing that can block (the value waits, the
        // release kernel that depends on the trace kernel) is submitted before the work it could hold up.
        typedef int (*StreamValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
        const StreamValue32 waitValue = reinterpret_cast<StreamValue32>(ctx->waitValue32);
        uint32_t chunks = uint32_t(std::max<uint64_t>(4, std::min<uint64_t>(16, (count + 62500) / 125000)));
        if (const char* e = getenv("ATLAS_RT_STREAM_CHUNKS")) chunks = uint32_t(std::max(1, std::min(32, atoi(e))));
        const uint32_t chunkRays = uint32_t(((count + chunks - 1) / chunks + 31) & ~uint64_t(31));
        chunks = uint32_t((count + chunkRays - 1) / chunkRays);
        cudaEvent_t* ev = ctx->pipeEvents;   // [0] state reset, [1+c] chunk c uploaded, [33] all downloaded
        const unsigned long long stateAddr = reinterpret_cast<unsigned long long>(ctx->dStreamState);
        uint32_t* dPerm = nullptr;
        uint8_t* dBucket = nullptr;
        unsigned int* dHist = nullptr;
        // 1. reset the watermark and the per-chunk counts; the other streams start after that
        cudaError_t e = cudaMemsetAsync(ctx->dStreamState, 0, 64 * sizeof(unsigned int), ctx->stream);
        if (e == cudaSuccess) e = dev_alloc(ctx, &dPerm, count);
        if (e == cudaSuccess) e = dev_alloc(ctx, &dBucket, count);
        if (e == cudaSuccess) e = dev_alloc(ctx, &dHist, size_t(chunks) * 128);
        if (e == cudaSuccess) e = cudaEventRecord(ev[0], ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyIn, ev[0], 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyOut, ev[0], 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->sortStream, ev[0], 0);
        // 2. uploads, each followed (on the ordering stream) by the chunk's sort and the watermark it justifies
        bool started = false;
        for (uint32_t c = 0; c < chunks && e == cudaSuccess && rc == ATLAS_RT_OK; c++) {
            const uint64_t b = uint64_t(c) * chunkRays, end = std::min<uint64_t>(count, b + chunkRays);
            cudaStream_t up = ctx->copyIn;   // (two upload streams were tried: the copy engine serves one stream's copies first, which only delays every second chunk)
            e = cudaMemcpyAsync(dIn + 3 * b, static_cast<const char*>(rays_in) + 48 * b, 48 * (end - b), cudaMemcpyHostToDevice, up);
            if (e == cudaSuccess) e = cudaEventRecord(ev[1 + c], up);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->sortStream, ev[1 + c], 0);
            if (e == cudaSuccess) rc = launch_chunk_sort(ctx, scene, ctx->sortStream, dIn + 3 * b, uint32_t(end - b), uint32_t(b), dBucket + b, dHist + size_t(c) * 128,
                                                         dPerm + b, ctx->dStreamState);
            if (ctx->pipeTimeline && chunks <= 16 && e == cudaSuccess) e = cudaEventRecord(ev[17 + c], ctx->sortStream);
        }
        // 3. the persistent trace kernel (it may already find the first chunks in place)
        if (e == cudaSuccess && rc == ATLAS_RT_OK) {
            rc = launch_trace(ctx, scene, dIn, out, count, cull_mask, t_min, t_max, any, perRay, counters, true, opacity, nullptr, 0, nullptr, hitsOnly,
                              ctx->dStreamState, ctx->dStreamState + 1, chunkRays, dPerm);
            started = rc == ATLAS_RT_OK;
        }
        // 4. downloads, each behind its chunk's completion count
        for (uint32_t c = 0; c < chunks && e == cudaSuccess && rc == ATLAS_RT_OK && !devOut; c++) {
            const uint64_t b = uint64_t(c) * chunkRays, end = std::min<uint64_t>(count, b + chunkRays);
            if (waitValue(ctx->copyOut, stateAddr + 4ull * (1 + c), unsigned(end - b), 0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0) { e = cudaErrorUnknown; break; }
            e = cudaMemcpyAsync(static_cast<char*>(rays_out) + 16 * outStride * b, out + outStride * b, 16 * outStride * (end - b), cudaMemcpyDeviceToHost, ctx->copyOut);
        }
        // 5. behind the trace kernel: every chunk's count reaches its target no matter what (the waits above cannot be left hanging)
        if (started) { const int rr = launch_release_chunks(ctx, ctx->dStreamState + 1, chunkRays, uint32_t(count), chunks); if (rc == ATLAS_RT_OK) rc = rr; }
        if (started && (e != cudaSuccess || rc != ATLAS_RT_OK)) {   // never leave the kernel waiting for rays that will not come
            const unsigned int all = unsigned(count);
            cudaMemcpyAsync(ctx->dStreamState, &all, sizeof(all), cudaMemcpyHostToDevice, ctx->copyIn);
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev[33], ctx->copyOut);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev[33], 0);   // the context stream now orders after the downloads
        // the ordering stream's scratch is released behind the trace kernel (the context stream has just been ordered after it)
        if (started) { cudaEventRecord(ev[34], ctx->sortStream); cudaStreamWaitEvent(ctx->stream, ev[34], 0); }
        else cudaStreamSynchronize(ctx->sortStream);
        dev_free(ctx, dPerm);
        dev_free(ctx, dBucket);
        dev_free(ctx, dHist);
        if (e != cudaSuccess && rc == ATLAS_RT_OK) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "streaming trace", e);
        if (ctx->pipeTimeline && chunks <= 16 && rc == ATLAS_RT_OK && cudaStreamSynchronize(ctx->stream) == cudaSuccess) {
            fprintf(stderr, "[atlas_rt streaming timeline] %u chunks:", chunks);
            for (uint32_t c = 0; c < chunks; c++) {
                float up = 0.0f, pub = 0.0f;
                cudaEventElapsedTime(&up, ev[0], ev[1 + c]);
                cudaEventElapsedTime(&pub, ev[0], ev[17 + c]);
                fprintf(stderr, " [%u up %.3f sorted %.3f]", c, up, pub);
            }
            float all = 0.0f;
            cudaEventElapsedTime(&all, ev[0], ev[33]);
            fprintf(stderr, " all %.3f ms\n", all);
        }
    } else if (!devIn && count >= kPipeMin && ctx->copyIn && ctx->copyOut) {
        // Host input: split the batch and overlap H2D of chunk i+1, the trace of chunk i and (host output) D2H of
        // chunk i-1 on the two copy engines (pays off with pinned host memory; pageable memory still works). With
        // ATLAS_RT_DEVICE_OUTPUT the hits stay on the device, e.g. for an NCCL gather.
        // Chunks of about 125 k rays, spread over up to 8 compute streams: a launch that small occupies a fraction of the
        // GPU for the latency of its longest ray, and several of them run side by side while later chunks are still
        // arriving and earlier ones are going back (swept on C2: 8 chunks x 8 streams 1.89 ms, 4 x 2 2.04 ms, 1 x 1 2.75 ms)
        uint32_t chunks = uint32_t(std::max<uint64_t>(2, std::min<uint64_t>(16, (count + 62500) / 125000)));
        if (const char* e = getenv("ATLAS_RT_PIPE_CHUNKS")) chunks = uint32_t(std::max(1, std::min(16, atoi(e))));
        cudaEvent_t* ev = ctx->pipeEvents;   // [0] staging ready, [1+c] chunk c uploaded, [17+c] chunk c traced, [33] all downloaded
        // Chunks alternate between the context stream and a second compute stream (each with its own ray-queue head), so
        // the thin tail of one chunk's persistent kernel overlaps the start of the next chunk's.
        cudaError_t e = cudaMemsetAsync(ctx->dCounters, 0, 6 * sizeof(unsigned long long), ctx->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[0], ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyIn, ev[0], 0);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyOut, ev[0], 0);
        uint32_t nStreams = 1;   // usable compute streams: the context stream + the extra ones that exist
        while (nStreams < uint32_t(ctx->pipeStreams) && ctx->computeExtra[nStreams - 1]) nStreams++;
        for (uint32_t k = 1; k < nStreams && e == cudaSuccess; k++) e = cudaStreamWaitEvent(ctx->computeExtra[k - 1], ev[0], 0);
        // chunk boundaries as fractions of the batch; ATLAS_RT_PIPE_SPLIT="0.15,0.5,0.85" overrides the equal split
        double cut[17];
        for (uint32_t c = 0; c <= chunks; c++) cut[c] = double(c) / chunks;
        if (const char* sp = getenv("ATLAS_RT_PIPE_SPLIT")) {
            uint32_t k = 1;
            for (const char* q = sp; *q && k < 16; k++) {
                cut[k] = atof(q);
                const char* comma = strchr(q, ',');
                if (!comma) { k++; break; }
                q = comma + 1;
            }
            chunks = k;
            cut[chunks] = 1.0;
        }
        for (uint32_t c = 0; c < chunks && e == cudaSuccess && rc == ATLAS_RT_OK; c++) {
            const uint64_t b = uint64_t(count * cut[c]) & ~uint64_t(31), end = c + 1 == chunks ? count : (uint64_t(count * cut[c + 1]) & ~uint64_t(31));
            const char* hIn = static_cast<const char*>(rays_in) + 48 * b;
            char* hOut = static_cast<char*>(rays_out) + 16 * outStride * b;
            const int slot = int(c % nStreams);
            cudaStream_t cs = slot ? ctx->computeExtra[slot - 1] : ctx->stream;
            cudaStream_t up = ctx->copyIn;   // (two upload streams were tried: the copy engine serves one stream's copies first, which only delays every second chunk)
            e = cudaMemcpyAsync(dIn + 3 * b, hIn, 48 * (end - b), cudaMemcpyHostToDevice, up);
            if (e == cudaSuccess) e = cudaEventRecord(ev[1 + c], up);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ev[1 + c], 0);
            if (e != cudaSuccess) break;
            float4* dst = out + outStride * b;   // host output of whole rays: in place on the staging buffer (out == dIn)
            rc = launch_trace(ctx, scene, dIn + 3 * b, dst, end - b, cull_mask, t_min, t_max, any, perRay, counters, false, opacity, cs, slot, nullptr, hitsOnly);
            if (rc != ATLAS_RT_OK) break;
            e = cudaEventRecord(ev[17 + c], cs);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyOut, ev[17 + c], 0);
            if (e == cudaSuccess && !devOut) e = cudaMemcpyAsync(hOut, dst, 16 * outStride * (end - b), cudaMemcpyDeviceToHost, ctx->copyOut);
        }
        if (e == cudaSuccess) e = cudaEventRecord(ev[33], ctx->copyOut);
        // ATLAS_RT_PIPELINED (with ATLAS_RT_ASYNC, host output): the call's completion is NOT ordered into the context stream, so the
        // next such call starts uploading while this one's last chunks are still being traced and sent home (the calls share the
        // copy / compute streams, which keep each of them in order). atlas_rt_trace_join orders the context stream after the last one.
        pipelinedCall = stageSet >= 0 && rc == ATLAS_RT_OK && e == cudaSuccess;
        if (pipelinedCall) {
            ctx->pendingJoin = true;
            e = cudaEventRecord(ctx->pipeEvents[36 + stageSet], ctx->copyOut);
            ctx->stageUsed[stageSet] = true;
        } else if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev[33], 0);   // the context stream now orders after the downloads
        if (e != cudaSuccess && rc == ATLAS_RT_OK) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "pipelined trace", e);
        if (ctx->pipeTimeline && rc == ATLAS_RT_OK && cudaEventSynchronize(ev[33]) == cudaSuccess) {
            fprintf(stderr, "[atlas_rt timeline] %u chunks:", chunks);
            for (uint32_t c = 0; c < chunks; c++) {
                float up = 0.0f, tr = 0.0f;
                cudaEventElapsedTime(&up, ev[0], ev[1 + c]);
                cudaEventElapsedTime(&tr, ev[0], ev[17 + c]);
                fprintf(stderr, " [%u up %.3f traced %.3f]", c, up, tr);
            }
            float all = 0.0f;
            cudaEventElapsedTime(&all, ev[0], ev[33]);
            fprintf(stderr, " all %.3f ms\n", all);
        }
    } else {
        if (!devIn) {
            cudaError_t e = copy_in(ctx, dIn, rays_in, count * 48, false);
            if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "copy_in", e);
        }
        if (rc == ATLAS_RT_OK) rc = launch_trace(ctx, scene, in, out, count, cull_mask, t_min, t_max, any, perRay, counters, true, opacity, nullptr, 0, nullptr, hitsOnly);
        if (rc == ATLAS_RT_OK && !devOut) {
            cudaError_t e = copy_out(ctx, rays_out, out, count * 16 * outStride, false);
            if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "copy_out", e);
        }
    }
    if (rc != ATLAS_RT_OK) {   // the copy / extra compute streams may still be touching the staging buffer
        if (ctx->copyIn) cudaStreamSynchronize(ctx->copyIn);
        if (ctx->copyOut) cudaStreamSynchronize(ctx->copyOut);
        for (auto& cs : ctx->computeExtra) if (cs) cudaStreamSynchronize(cs);
        if (ctx->sortStream) cudaStreamSynchronize(ctx->sortStream);
    }
    if (stageSet >= 0) {
        // persistent staging: nothing to release; a failed or fallen-back call must not leave the set in flight unaccounted
        if (!pipelinedCall) { cudaStreamSynchronize(ctx->copyIn); cudaStreamSynchronize(ctx->copyOut); for (auto& cs : ctx->computeExtra) if (cs) cudaStreamSynchronize(cs); cudaStreamSynchronize(ctx->stream); }
    } else {
        dev_free(ctx, dIn);
        dev_free(ctx, dOut);
    }
    if (rc != ATLAS_RT_OK) return rc;
    if (flags & ATLAS_RT_ASYNC) return ATLAS_RT_OK;
    // synchronous call: also report rays that ran out of the reference's 32-entry stack
    ATLAS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->dCounters + 5, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t overflowed = 0;
    memcpy(&overflowed, ctx->pinned, sizeof(overflowed));
    if (overflowed) return fail(ctx, ATLAS_RT_ERR_STACK, "rays exceeded the 32-entry traversal stack (undefined behaviour in the reference)");
    return ATLAS_RT_OK;
}

int atlas_rt_trace_closest(atlas_rt_context* ctx, const atlas_rt_scene* scene, const void* rays_in, uint64_t count,
                           uint32_t cull_mask, float t_min, float t_max, void* rays_out, uint32_t flags) {
    return trace_common(ctx, scene, rays_in, count, cull_mask, t_min, t_max, rays_out, flags, false);
}

int atlas_rt_trace_any(atlas_rt_context* ctx, const atlas_rt_scene* scene, const void* rays_in, uint64_t count,
                       uint32_t cull_mask, float t_min, float t_max, void* rays_out, uint32_t flags) {
    return trace_common(ctx, scene, rays_in, count, cull_mask, t_min, t_max, rays_out, flags, true);
}

int atlas_rt_trace_join(atlas_rt_context* ctx) {
    if (!ctx) return ATLAS_RT_ERR_INVALID;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->pendingJoin) {
        ATLAS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipeEvents[33], 0));
        for (auto& cs : ctx->computeExtra) {   // (the chunk traces of those calls; implied by the downloads, made explicit for device-side consumers)
            if (!cs) continue;
            ATLAS_CUDA(ctx, cudaEventRecord(ctx->pipeEvents[35], cs));
            ATLAS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pipeEvents[35], 0));
        }
        ctx->pendingJoin = false;
    }
    return ATLAS_RT_OK;
}

int atlas_rt_trace_counters(atlas_rt_context* ctx, uint64_t out[6]) {
    if (!ctx || !out) return ATLAS_RT_ERR_INVALID;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    ATLAS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, ctx->dCounters, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(out, ctx->pinned, 6 * sizeof(uint64_t));
    return ATLAS_RT_OK;
}

int atlas_rt_shard_range(uint64_t count, uint32_t rank, uint32_t world, uint32_t align, uint64_t* begin, uint64_t* end) {
    if (!begin || !end || world == 0 || rank >= world) return ATLAS_RT_ERR_INVALID;
    if (align == 0) align = 1;
    const uint64_t units = (count + align - 1) / align;
    const uint64_t base = units / world, extra = units % world;
    const uint64_t b = (base * rank + (rank < extra ? rank : extra)) * align;
    const uint64_t e = b + (base + (rank < extra ? 1 : 0)) * align;
    *begin = b < count ? b : count;
    *end = e < count ? e : count;
    return ATLAS_RT_OK;
}

}   // extern "C"
