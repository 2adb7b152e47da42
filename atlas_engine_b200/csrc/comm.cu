// comm.cu — the multi-GPU half of the C ABI: one process and one context per GPU of a node, NCCL over
// NVLink / NVSwitch as the transport. Nothing here sits inside the traversal: the scene is replicated (broadcast once),
// every rank traces its own contiguous share of the rays, and ONE gather of 16-byte hit records per batch brings the
// result to the root — the shape the north star names. The reference has no multi-GPU code; the consumers are
// RayTracingWorld::UpdateForSoftwareRayTracing (scene assembly, src/engine/raytracing/RayTracingWorld.cpp:267-307) and
// RayTracingHelper::DispatchHitClosest (the trace batch, src/engine/renderer/helper/RayTracingHelper.cpp:346-364).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, a process that already
// carries NCCL (e.g. through PyTorch) shares that copy, and single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

namespace atlas {
namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) api.why = std::string("missing NCCL symbol ") + n; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.why.empty();
    });
    return api;
}

// ---- peer-memory window (atlas_rt_trace_sharded with ATLAS_RT_PEER_OUTPUT) -------------------------------------------
constexpr size_t kWindowFlagBytes = 4096;   // done[slot][rank] counters (u32), at the start of the root's window allocation

// Runs behind a rank's trace launch on the same stream (the launch, and with it every hit record it stored into the root's
// memory, is complete): publish "rank r has finished call k" in the root's memory.
__global__ void peer_signal(unsigned int* flag, unsigned int value) {
    chain_begin();
    __threadfence_system();
    *reinterpret_cast<volatile unsigned int*>(flag) = value;
    __threadfence_system();
}
struct PeerFlagPtrs { unsigned int* f[64]; };
// Root: tell every rank (in ITS memory, where it polls) that the slot of call k has been consumed and may be written again.
__global__ void peer_release(PeerFlagPtrs p, uint32_t world, unsigned int value) {
    if (threadIdx.x < world) {
        __threadfence_system();
        *reinterpret_cast<volatile unsigned int*>(p.f[threadIdx.x]) = value;
    }
}

}   // namespace
}   // namespace atlas

struct atlas_rt_comm {
    atlas_rt_context* ctx = nullptr;
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
    cudaStream_t stream = nullptr;          // collectives run here, beside the context's compute stream
    cudaEvent_t traced[2] = {}, gathered[2] = {};
    float4* hits[2] = {nullptr, nullptr};   // this rank's hit records, double buffered: gather k overlaps trace k + 1
    uint64_t hitsCapacity = 0;
    uint64_t calls = 0;
    void* pinned = nullptr;                 // header exchange
    // peer-memory window of atlas_rt_trace_sharded(ATLAS_RT_PEER_OUTPUT): the traversal kernels of all ranks store their
    // hit records straight into the root's buffer over NVLink; flags written the same way tell the root when a rank is done
    bool peerReady = false;
    uint32_t peerRoot = 0;
    uint64_t peerCapacity = 0;              // records per slot
    uint64_t peerCalls = 0;
    char* window = nullptr;                 // root: own allocation; elsewhere the root's, opened through CUDA IPC: [4096 B flags][2 slots][capacity] records
    unsigned int* localFlag = nullptr;      // this rank's "root has consumed call k" counter (written by the root over NVLink)
    unsigned int* peerFlags[64] = {};       // root: every rank's localFlag
};

using namespace atlas;

namespace {

int nccl_fail(atlas_rt_comm* c, const char* what, ncclResult_t r) {
    atlas_rt_context* ctx = c ? c->ctx : nullptr;
    if (ctx) {
        ctx->error = std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error");
    }
    return ATLAS_RT_ERR_CUDA;
}
#define ATLAS_NCCL(c, call)                                              \
    do {                                                                 \
        ncclResult_t r__ = (call);                                       \
        if (r__ != ncclSuccess) return nccl_fail((c), #call, r__);      \
    } while (0)

// Broadcast `bytes` bytes of device memory from `root` on the comm stream.
int bcast(atlas_rt_comm* c, void* dev, size_t bytes, uint32_t root) {
    if (bytes == 0) return ATLAS_RT_OK;
    ATLAS_NCCL(c, nccl().Broadcast(dev, dev, bytes, ncclUint8, int(root), c->comm, c->stream));
    return ATLAS_RT_OK;
}

// A few host words from root to everybody (sizes of what follows).
int bcast_header(atlas_rt_comm* c, uint64_t* words, size_t n, uint32_t root) {
    atlas_rt_context* ctx = c->ctx;
    uint64_t* dev = nullptr;
    ATLAS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&dev), n * 8, c->stream));
    if (c->rank == root) ATLAS_CUDA(ctx, cudaMemcpyAsync(dev, words, n * 8, cudaMemcpyHostToDevice, c->stream));
    int rc = bcast(c, dev, n * 8, root);
    if (rc == ATLAS_RT_OK) {
        cudaError_t e = cudaMemcpyAsync(words, dev, n * 8, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "header broadcast", e);
    }
    cudaFreeAsync(dev, c->stream);
    return rc;
}

// comm stream waits for everything issued so far on the context stream, and vice versa
int comm_after_ctx(atlas_rt_comm* c) {
    ATLAS_CUDA(c->ctx, cudaEventRecord(c->traced[0], c->ctx->stream));
    ATLAS_CUDA(c->ctx, cudaStreamWaitEvent(c->stream, c->traced[0], 0));
    return ATLAS_RT_OK;
}
int ctx_after_comm(atlas_rt_comm* c) {
    ATLAS_CUDA(c->ctx, cudaEventRecord(c->gathered[0], c->stream));
    ATLAS_CUDA(c->ctx, cudaStreamWaitEvent(c->ctx->stream, c->gathered[0], 0));
    return ATLAS_RT_OK;
}

// ---- peer-memory window ---------------------------------------------------------------------------------------------
struct PeerHandles { cudaIpcMemHandle_t flag, window; };
static_assert(sizeof(PeerHandles) == 128, "two 64-byte IPC handles");

void peer_teardown(atlas_rt_comm* c) {
    if (c->window) { if (c->rank == c->peerRoot) cudaFree(c->window); else cudaIpcCloseMemHandle(c->window); }
    for (uint32_t r = 0; r < 64; r++) if (c->peerFlags[r] && c->peerFlags[r] != c->localFlag) cudaIpcCloseMemHandle(c->peerFlags[r]);
    if (c->localFlag) cudaFree(c->localFlag);
    c->window = nullptr;
    c->localFlag = nullptr;
    memset(c->peerFlags, 0, sizeof(c->peerFlags));
    c->peerReady = false;
    c->peerCapacity = 0;
    c->peerCalls = 0;
}

// Collective: (re)create the window for `capacity` records per slot on `root` and open it everywhere else.
int peer_setup(atlas_rt_comm* c, uint64_t capacity, uint32_t root) {
    atlas_rt_context* ctx = c->ctx;
    if (c->world > 64) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "peer window: more than 64 ranks");
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(c->stream));
    peer_teardown(c);
    c->peerRoot = root;
    ATLAS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&c->localFlag), 256));
    ATLAS_CUDA(ctx, cudaMemset(c->localFlag, 0, 256));
    PeerHandles mine;
    memset(&mine, 0, sizeof(mine));
    ATLAS_CUDA(ctx, cudaIpcGetMemHandle(&mine.flag, c->localFlag));
    if (c->rank == root) {
        ATLAS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&c->window), kWindowFlagBytes + 2 * capacity * 16));
        ATLAS_CUDA(ctx, cudaMemset(c->window, 0, kWindowFlagBytes));
        ATLAS_CUDA(ctx, cudaIpcGetMemHandle(&mine.window, c->window));
    }
    // exchange: one 128-byte broadcast per rank
    std::vector<PeerHandles> all(c->world);
    PeerHandles* dev = nullptr;
    ATLAS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&dev), sizeof(PeerHandles) * c->world));
    cudaError_t e = cudaMemcpy(dev + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    int rc = e == cudaSuccess ? ATLAS_RT_OK : fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: handle upload", e);
    for (uint32_t r = 0; r < c->world && rc == ATLAS_RT_OK; r++) rc = bcast(c, dev + r, sizeof(PeerHandles), r);
    if (rc == ATLAS_RT_OK) {
        e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaMemcpy(all.data(), dev, sizeof(PeerHandles) * c->world, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: handle exchange", e);
    }
    cudaFree(dev);
    if (rc != ATLAS_RT_OK) return rc;
    if (c->rank == root) {
        for (uint32_t r = 0; r < c->world; r++) {
            if (r == root) { c->peerFlags[r] = c->localFlag; continue; }
            void* p = nullptr;
            e = cudaIpcOpenMemHandle(&p, all[r].flag, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: cudaIpcOpenMemHandle (a rank's flag)", e);
            c->peerFlags[r] = static_cast<unsigned int*>(p);
        }
    } else {
        void* p = nullptr;
        e = cudaIpcOpenMemHandle(&p, all[root].window, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: cudaIpcOpenMemHandle (the root's window; needs NVLink / PCIe peer access between the GPUs)", e);
        c->window = static_cast<char*>(p);
    }
    c->peerCapacity = capacity;
    c->peerReady = true;
    return ATLAS_RT_OK;
}

// The sharded trace batch with the gather fused into the traversal: every rank's kernel stores each hit record, as the ray
// finishes, at its global position in the ROOT's window (P2P stores over NVLink / NVSwitch; the root's own kernel writes
// locally), so the transfer rides along with the traversal and no collective, copy kernel or copy engine touches the data.
// Flow control is two counters per rank, both written over NVLink and polled locally with cuStreamWaitValue32: done[slot][r]
// in the root's memory (rank r has finished call k), consumed in rank r's memory (the root has used call k's slot, which
// call k + 2 writes again).
int trace_sharded_peer(atlas_rt_comm* c, const atlas_rt_scene* scene, const void* rays_in, uint64_t total_count, uint32_t cull_mask, float t_min,
                       float t_max, void* hits_out, uint32_t root, uint32_t flags, int any_hit) {
    atlas_rt_context* ctx = c->ctx;
    if (!ctx->waitValue32) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "ATLAS_RT_PEER_OUTPUT needs cuStreamWaitValue32");
    typedef int (*StreamValue32)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
    const StreamValue32 waitValue = reinterpret_cast<StreamValue32>(ctx->waitValue32);
    if (!c->peerReady || c->peerRoot != root || c->peerCapacity < total_count) {
        const int rc = peer_setup(c, std::max<uint64_t>(total_count, 64), root);
        if (rc != ATLAS_RT_OK) return rc;
    }
    uint64_t b = 0, e = 0;
    atlas_rt_shard_range(total_count, c->rank, c->world, 64, &b, &e);
    const uint64_t local = e - b;
    const uint64_t k = c->peerCalls;
    const uint32_t slot = uint32_t(k & 1);
    float4* slotBase = reinterpret_cast<float4*>(c->window + kWindowFlagBytes) + size_t(slot) * c->peerCapacity;
    unsigned int* done = reinterpret_cast<unsigned int*>(c->window) + slot * 64u;
    // the slot was last written by call k - 2: wait (locally) until the root has consumed that
    if (k >= 2 && waitValue(ctx->stream, reinterpret_cast<unsigned long long>(c->localFlag), unsigned(k - 1), 0u /* GEQ */) != 0)
        return fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: cuStreamWaitValue32");
    if (local) {
        const uint32_t tf = (flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_PER_RAY_TMAX | ATLAS_RT_OPACITY)) | ATLAS_RT_DEVICE_OUTPUT | ATLAS_RT_HITS_ONLY | ATLAS_RT_ASYNC;
        const int rc = any_hit ? atlas_rt_trace_any(ctx, scene, rays_in, local, cull_mask, t_min, t_max, slotBase + b, tf)
                               : atlas_rt_trace_closest(ctx, scene, rays_in, local, cull_mask, t_min, t_max, slotBase + b, tf);
        if (rc != ATLAS_RT_OK) return rc;
    }
    ATLAS_CUDA(ctx, launch_chain(ctx->chainLaunch != 0, peer_signal, 1, 1, 0, ctx->stream, done + c->rank, unsigned(k + 1)));
    ctx->launches++;
    if (c->rank == root) {
        // the root's communicator stream: wait for every rank, hand the records on, release the slot
        for (uint32_t r = 0; r < c->world; r++)
            if (waitValue(c->stream, reinterpret_cast<unsigned long long>(done + r), unsigned(k + 1), 0u) != 0) return fail(ctx, ATLAS_RT_ERR_CUDA, "peer window: cuStreamWaitValue32");
        if (hits_out && total_count)
            ATLAS_CUDA(ctx, cudaMemcpyAsync(hits_out, slotBase, total_count * 16, (flags & ATLAS_RT_DEVICE_OUTPUT) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        PeerFlagPtrs ptrs;
        memcpy(ptrs.f, c->peerFlags, sizeof(ptrs.f));
        peer_release<<<1, 64, 0, c->stream>>>(ptrs, c->world, unsigned(k + 1));
        ctx->launches++;
        ATLAS_CUDA(ctx, cudaGetLastError());
        ATLAS_CUDA(ctx, cudaEventRecord(c->gathered[slot], c->stream));
    }
    c->peerCalls++;
    if (!(flags & ATLAS_RT_ASYNC)) {
        ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ATLAS_CUDA(ctx, cudaStreamSynchronize(c->stream));
    }
    return ATLAS_RT_OK;
}

}   // namespace

extern "C" {

int atlas_rt_comm_unique_id(void* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
    if (!id128) return ATLAS_RT_ERR_INVALID;
    if (!nccl().ok) return ATLAS_RT_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return ATLAS_RT_ERR_CUDA;
    memcpy(id128, &id, 128);
    return ATLAS_RT_OK;
}

int atlas_rt_comm_init(atlas_rt_context* ctx, const void* id128, uint32_t rank, uint32_t world, atlas_rt_comm** out_comm) {
    if (!ctx || !id128 || !out_comm || world == 0 || rank >= world) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    *out_comm = nullptr;
    if (!nccl().ok) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, nccl().why.c_str());
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    auto* c = new (std::nothrow) atlas_rt_comm;
    if (!c) return fail(ctx, ATLAS_RT_ERR_OOM, "host allocation");
    c->ctx = ctx;
    ctx_retain(ctx);
    c->rank = rank;
    c->world = world;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&c->traced[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->gathered[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { atlas_rt_comm_destroy(c); return fail(ctx, ATLAS_RT_ERR_CUDA, "comm streams", e); }
    const ncclResult_t r = nccl().CommInitRank(&c->comm, int(world), id, int(rank));
    if (r != ncclSuccess) { c->comm = nullptr; const int rc = nccl_fail(c, "ncclCommInitRank", r); atlas_rt_comm_destroy(c); return rc; }
    *out_comm = c;
    return ATLAS_RT_OK;
}

void atlas_rt_comm_destroy(atlas_rt_comm* c) {
    if (!c) return;
    atlas_rt_context* ctx = c->ctx;
    cudaSetDevice(ctx->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(ctx->stream);
    peer_teardown(c);
    if (c->comm) nccl().CommDestroy(c->comm);
    for (int k = 0; k < 2; k++) {
        if (c->hits[k]) cudaFree(c->hits[k]);
        if (c->traced[k]) cudaEventDestroy(c->traced[k]);
        if (c->gathered[k]) cudaEventDestroy(c->gathered[k]);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    ctx_release(ctx);
}

int atlas_rt_comm_info(const atlas_rt_comm* c, uint32_t* rank, uint32_t* world) {
    if (!c) return ATLAS_RT_ERR_INVALID;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return ATLAS_RT_OK;
}

int atlas_rt_comm_synchronize(atlas_rt_comm* c) {
    if (!c) return ATLAS_RT_ERR_INVALID;
    ATLAS_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    ATLAS_CUDA(c->ctx, cudaStreamSynchronize(c->ctx->stream));
    ATLAS_CUDA(c->ctx, cudaStreamSynchronize(c->stream));
    return ATLAS_RT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// One flattened tree from the rank that built it to everybody else. On `root` the tree is returned as is; elsewhere a new
// object is created from the received arrays (GPUBVHNode / order / endOfNode are position independent).
int atlas_rt_bvh_broadcast(atlas_rt_comm* c, const atlas_rt_bvh* src, uint32_t root, atlas_rt_bvh** out_bvh) {
    if (!c || !out_bvh || root >= c->world || (c->rank == root && !src)) return fail(c ? c->ctx : nullptr, ATLAS_RT_ERR_INVALID, "bad argument");
    atlas_rt_context* ctx = c->ctx;
    *out_bvh = nullptr;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t header[2 + 8] = {0};
    if (c->rank == root) { header[0] = src->nodeCount; header[1] = src->refCount; memcpy(header + 2, src->stats, sizeof(src->stats)); }
    int rc = comm_after_ctx(c);   // the tree was built on the context stream
    if (rc == ATLAS_RT_OK) rc = bcast_header(c, header, 10, root);
    if (rc != ATLAS_RT_OK) return rc;
    atlas_rt_bvh* bvh = nullptr;
    if (c->rank == root) {
        bvh = const_cast<atlas_rt_bvh*>(src);
    } else {
        bvh = new (std::nothrow) atlas_rt_bvh;
        if (!bvh) return fail(ctx, ATLAS_RT_ERR_OOM, "host allocation");
        bvh->ctx = ctx;
        ctx_retain(ctx);
        bvh->nodeCount = header[0];
        bvh->refCount = header[1];
        memcpy(bvh->stats, header + 2, sizeof(bvh->stats));
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&bvh->nodes), std::max<uint64_t>(1, bvh->nodeCount) * 64, c->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&bvh->order), std::max<uint64_t>(1, bvh->refCount) * 4, c->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&bvh->endOfNode), std::max<uint64_t>(1, bvh->refCount), c->stream);
        if (e != cudaSuccess) { atlas_rt_bvh_free(bvh); return fail(ctx, ATLAS_RT_ERR_CUDA, "tree buffers", e); }
    }
    // a root-leaf BLAS (no nodes) carries one synthetic node in its storage: send that too
    const uint64_t nodeBytes = std::max<uint64_t>(1, header[0]) * 64;
    ATLAS_NCCL(c, nccl().GroupStart());
    rc = bcast(c, bvh->nodes, nodeBytes, root);
    if (rc == ATLAS_RT_OK) rc = bcast(c, bvh->order, header[1] * 4, root);
    if (rc == ATLAS_RT_OK) rc = bcast(c, bvh->endOfNode, header[1], root);
    ATLAS_NCCL(c, nccl().GroupEnd());
    if (rc == ATLAS_RT_OK) rc = ctx_after_comm(c);
    if (rc != ATLAS_RT_OK) { if (c->rank != root) atlas_rt_bvh_free(bvh); return rc; }
    *out_bvh = bvh;
    return ATLAS_RT_OK;
}

// Scene assembly with the BLAS builds dealt across the GPUs ("TLAS instance sets are split across GPUs"):
// rank r builds meshes r, r + world, ... as one batch, every tree is broadcast from its owner, the TLAS is built on rank 0
// and broadcast, and every rank packs and assembles the same scene. The scene owns everything it creates.
int atlas_rt_build_scene_sharded(atlas_rt_comm* c, uint32_t mesh_count, const float* const* aabbs, const float* const* tris, const uint64_t* counts,
                                 const void* instances64, const float* instance_aabbs, uint64_t instance_count, uint32_t flags,
                                 atlas_rt_scene** out_scene) {
    if (!c || !out_scene || !mesh_count || !aabbs || !tris || !counts || !instances64 || !instance_aabbs || !instance_count)
        return fail(c ? c->ctx : nullptr, ATLAS_RT_ERR_INVALID, "bad argument");
    atlas_rt_context* ctx = c->ctx;
    *out_scene = nullptr;
    if (flags & ATLAS_RT_DEVICE_INPUT) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_build_scene_sharded takes host arrays");
    std::vector<atlas_rt_bvh*> blas(mesh_count, nullptr);
    std::vector<atlas_rt_mesh*> meshes(mesh_count, nullptr);
    atlas_rt_bvh* tlas = nullptr;
    auto cleanup = [&](int rc) {
        for (auto* m : meshes) atlas_rt_mesh_free(m);
        for (auto* b : blas) atlas_rt_bvh_free(b);
        atlas_rt_bvh_free(tlas);
        return rc;
    };
    // 1. this rank's share of the BLAS builds, as one batch
    std::vector<uint32_t> mine;
    for (uint32_t k = c->rank; k < mesh_count; k += c->world) mine.push_back(k);
    if (!mine.empty()) {
        std::vector<const float*> a(mine.size()), t(mine.size());
        std::vector<uint64_t> n(mine.size());
        std::vector<atlas_rt_bvh*> out(mine.size(), nullptr);
        for (size_t i = 0; i < mine.size(); i++) { a[i] = aabbs[mine[i]]; t[i] = tris[mine[i]]; n[i] = counts[mine[i]]; }
        const int rc = atlas_rt_build_blas_batch(ctx, uint32_t(mine.size()), a.data(), t.data(), n.data(), 0, out.data());
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        for (size_t i = 0; i < mine.size(); i++) blas[mine[i]] = out[i];
    }
    // 2. every tree from its owner to everybody
    for (uint32_t k = 0; k < mesh_count; k++) {
        atlas_rt_bvh* got = nullptr;
        const int rc = atlas_rt_bvh_broadcast(c, blas[k], k % c->world, &got);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        blas[k] = got;
    }
    // 3. TLAS on rank 0, broadcast
    if (c->rank == 0) {
        const int rc = atlas_rt_build_tlas(ctx, instance_aabbs, instance_count, 0, &tlas);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
    }
    {
        atlas_rt_bvh* got = nullptr;
        const int rc = atlas_rt_bvh_broadcast(c, tlas, 0, &got);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        tlas = got;
    }
    // 4. pack + assemble locally (identical on every rank)
    for (uint32_t k = 0; k < mesh_count; k++) {
        const int rc = atlas_rt_pack_mesh(ctx, blas[k], tris[k], counts[k], nullptr, nullptr, 0, &meshes[k]);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
    }
    atlas_rt_scene* scene = nullptr;
    const int rc = atlas_rt_scene_create(ctx, meshes.data(), mesh_count, instances64, instance_count, tlas, 0, &scene);
    if (rc != ATLAS_RT_OK) return cleanup(rc);
    scene->ownedMeshes = meshes;
    scene->ownedBvhs = blas;
    scene->ownedBvhs.push_back(tlas);
    *out_scene = scene;
    return ATLAS_RT_OK;
}

// The scene of rank `root` on every rank: one broadcast per array. On root `src` is returned as is.
int atlas_rt_scene_replicate(atlas_rt_comm* c, const atlas_rt_scene* src, uint32_t root, atlas_rt_scene** out_scene) {
    if (!c || !out_scene || root >= c->world || (c->rank == root && !src)) return fail(c ? c->ctx : nullptr, ATLAS_RT_ERR_INVALID, "bad argument");
    atlas_rt_context* ctx = c->ctx;
    *out_scene = nullptr;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    // header: meshCount, instanceCount(source records = tlas refs), allShading, fastDivision, materialCount, textureCount, then per mesh triCount
    uint64_t head[6] = {0};
    std::vector<uint64_t> triCounts;
    std::vector<atlas_rt_mesh> hostMeshes;
    if (c->rank == root) {
        head[0] = src->meshCount; head[1] = src->instanceCount; head[2] = src->allShading ? 1 : 0; head[3] = uint64_t(src->fastDivision);
        head[4] = src->materialCount; head[5] = src->textureCount;
    }
    int rc = comm_after_ctx(c);
    if (rc == ATLAS_RT_OK) rc = bcast_header(c, head, 6, root);
    if (rc != ATLAS_RT_OK) return rc;
    if (c->rank == root) {
        if (src->partMeshes.size() != src->meshCount) return fail(ctx, ATLAS_RT_ERR_INVALID, "scene does not know its meshes");
    }
    const uint32_t meshCount = uint32_t(head[0]);
    std::vector<atlas_rt_bvh*> bvhs;
    std::vector<atlas_rt_mesh*> meshes(meshCount, nullptr);
    atlas_rt_bvh* tlas = nullptr;
    atlas_rt_scene* scene = nullptr;
    auto cleanup = [&](int code) {
        if (c->rank != root) {
            atlas_rt_scene_free(scene);
            for (auto* m : meshes) atlas_rt_mesh_free(m);
            for (auto* b : bvhs) atlas_rt_bvh_free(b);
            atlas_rt_bvh_free(tlas);
        }
        return code;
    };
    for (uint32_t k = 0; k < meshCount; k++) {
        const atlas_rt_mesh* sm = c->rank == root ? src->partMeshes[k] : nullptr;
        atlas_rt_bvh* b = nullptr;
        rc = atlas_rt_bvh_broadcast(c, sm ? sm->blas : nullptr, root, &b);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        if (c->rank != root) bvhs.push_back(b);
        uint64_t mh[2] = {sm ? sm->triCount : 0, (sm && sm->tris96) ? 1u : 0u};
        rc = bcast_header(c, mh, 2, root);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        atlas_rt_mesh* m = nullptr;
        if (c->rank == root) m = const_cast<atlas_rt_mesh*>(sm);
        else {
            m = new (std::nothrow) atlas_rt_mesh;
            if (!m) return cleanup(fail(ctx, ATLAS_RT_ERR_OOM, "host allocation"));
            m->ctx = ctx;
            ctx_retain(ctx);
            m->blas = b;
            m->triCount = mh[0];
            meshes[k] = m;
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&m->tris), std::max<uint64_t>(1, mh[0]) * 48, c->stream);
            if (e == cudaSuccess && mh[1]) e = cudaMallocAsync(reinterpret_cast<void**>(&m->tris96), std::max<uint64_t>(1, mh[0]) * 96, c->stream);
            if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "mesh buffers", e));
        }
        rc = bcast(c, m->tris, mh[0] * 48, root);
        if (rc == ATLAS_RT_OK && mh[1]) rc = bcast(c, m->tris96, mh[0] * 96, root);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
    }
    rc = atlas_rt_bvh_broadcast(c, c->rank == root ? src->tlas : nullptr, root, &tlas);
    if (rc != ATLAS_RT_OK) return cleanup(rc);
    if (c->rank == root) { *out_scene = const_cast<atlas_rt_scene*>(src); }
    // instances (already permuted into TLAS order on root) + pointer tables + materials / textures
    if (c->rank != root) {
        scene = new (std::nothrow) atlas_rt_scene;
        if (!scene) return cleanup(fail(ctx, ATLAS_RT_ERR_OOM, "host allocation"));
        scene->ctx = ctx;
        ctx_retain(ctx);
        scene->tlas = tlas;
        scene->meshCount = meshCount;
        scene->instanceCount = head[1];
        scene->allShading = head[2] != 0;
        scene->fastDivision = int(head[3]);
        std::vector<const float4*> nodePtrs(meshCount), triPtrs(meshCount), tri96Ptrs(meshCount);
        for (uint32_t k = 0; k < meshCount; k++) { nodePtrs[k] = meshes[k]->blas->nodes; triPtrs[k] = meshes[k]->tris; tri96Ptrs[k] = meshes[k]->tris96; }
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scene->blasNodes), meshCount * sizeof(void*), c->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&scene->bvhTris), meshCount * sizeof(void*), c->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&scene->triangles), meshCount * sizeof(void*), c->stream);
        if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&scene->instances), std::max<uint64_t>(1, head[1]) * 64, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scene->blasNodes, nodePtrs.data(), meshCount * sizeof(void*), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scene->bvhTris, triPtrs.data(), meshCount * sizeof(void*), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(scene->triangles, tri96Ptrs.data(), meshCount * sizeof(void*), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "scene tables", e));
        if (head[4]) { e = cudaMallocAsync(reinterpret_cast<void**>(&scene->materials), head[4] * 92, c->stream); scene->materialCount = uint32_t(head[4]); }
        if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "material table", e));
    }
    atlas_rt_scene* dst = c->rank == root ? const_cast<atlas_rt_scene*>(src) : scene;
    rc = bcast(c, dst->instances, head[1] * 64, root);
    if (rc == ATLAS_RT_OK && head[4]) rc = bcast(c, dst->materials, head[4] * 92, root);
    if (rc != ATLAS_RT_OK) return cleanup(rc);
    if (head[5]) {   // textures: dimensions first, then the texel storage; the receiver rebuilds its own pointer table
        std::vector<uint64_t> dims(2 * head[5] + 1, 0);
        std::vector<TextureDev> table(head[5]);
        if (c->rank == root) {
            cudaError_t e = cudaMemcpyAsync(table.data(), src->textures, head[5] * sizeof(TextureDev), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "texture table", e));
            for (uint64_t t = 0; t < head[5]; t++) { dims[2 * t] = table[t].width; dims[2 * t + 1] = table[t].height; }
        }
        rc = bcast_header(c, dims.data(), 2 * head[5], root);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
        size_t bytes = 0;
        for (uint64_t t = 0; t < head[5]; t++) bytes += (dims[2 * t] * dims[2 * t + 1] + 15) & ~uint64_t(15);
        if (c->rank != root) {
            cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scene->texelStorage), bytes, c->stream);
            if (e == cudaSuccess) e = cudaMallocAsync(reinterpret_cast<void**>(&scene->textures), head[5] * sizeof(TextureDev), c->stream);
            if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "texture storage", e));
            size_t off = 0;
            for (uint64_t t = 0; t < head[5]; t++) {
                table[t] = TextureDev{scene->texelStorage + off, uint32_t(dims[2 * t]), uint32_t(dims[2 * t + 1])};
                off += (dims[2 * t] * dims[2 * t + 1] + 15) & ~uint64_t(15);
            }
            e = cudaMemcpyAsync(scene->textures, table.data(), head[5] * sizeof(TextureDev), cudaMemcpyHostToDevice, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return cleanup(fail(ctx, ATLAS_RT_ERR_CUDA, "texture table", e));
            scene->textureCount = uint32_t(head[5]);
        }
        rc = bcast(c, dst->texelStorage, bytes, root);
        if (rc != ATLAS_RT_OK) return cleanup(rc);
    }
    rc = ctx_after_comm(c);
    if (rc == ATLAS_RT_OK) { cudaError_t e = cudaStreamSynchronize(c->stream); if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "replicate", e); }
    if (rc != ATLAS_RT_OK) return cleanup(rc);
    if (c->rank != root) {
        scene->ownedMeshes = meshes;
        scene->partMeshes.assign(meshes.begin(), meshes.end());
        scene->ownedBvhs = bvhs;
        scene->ownedBvhs.push_back(tlas);
        *out_scene = scene;
    }
    return ATLAS_RT_OK;
}

// Gather `bytes` bytes of device memory from every rank into `recv` on root at byte offsets `offsets[r]` (known to every
// rank: shares are made by atlas_rt_shard_range). One NCCL group of sends / receives; root copies its own part.
int atlas_rt_comm_gather(atlas_rt_comm* c, const void* send, uint64_t bytes, void* recv, const uint64_t* sizes, const uint64_t* offsets, uint32_t root,
                         uint32_t flags) {
    if (!c || root >= c->world || (bytes && !send) || (c->rank == root && (!recv || !sizes || !offsets))) return fail(c ? c->ctx : nullptr, ATLAS_RT_ERR_INVALID, "bad argument");
    atlas_rt_context* ctx = c->ctx;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = comm_after_ctx(c);
    if (rc != ATLAS_RT_OK) return rc;
    ATLAS_NCCL(c, nccl().GroupStart());
    if (c->rank == root) {
        for (uint32_t r = 0; r < c->world; r++) {
            if (r == root || sizes[r] == 0) continue;
            ATLAS_NCCL(c, nccl().Recv(static_cast<char*>(recv) + offsets[r], sizes[r], ncclUint8, int(r), c->comm, c->stream));
        }
    } else if (bytes) {
        ATLAS_NCCL(c, nccl().Send(send, bytes, ncclUint8, int(root), c->comm, c->stream));
    }
    ATLAS_NCCL(c, nccl().GroupEnd());
    if (c->rank == root && bytes && static_cast<char*>(recv) + offsets[root] != send)
        ATLAS_CUDA(ctx, cudaMemcpyAsync(static_cast<char*>(recv) + offsets[root], send, bytes, cudaMemcpyDeviceToDevice, c->stream));
    rc = ctx_after_comm(c);
    if (rc != ATLAS_RT_OK) return rc;
    if (!(flags & ATLAS_RT_ASYNC)) ATLAS_CUDA(ctx, cudaStreamSynchronize(c->stream));
    return ATLAS_RT_OK;
}

// The sharded trace batch. Every rank passes ITS share of the rays (share r of `total_count` by atlas_rt_shard_range with
// align 64; host or device memory); the scene is the rank's replica. Each rank traces its share into compact hit records
// and the records are gathered on `root` in global ray order: hits_out (root only; total_count x 16 B, device memory if
// ATLAS_RT_DEVICE_OUTPUT). With ATLAS_RT_ASYNC the gather of call k runs on the communicator's stream while call k + 1 is
// already tracing (the local records are double buffered); atlas_rt_comm_synchronize joins everything.
int atlas_rt_trace_sharded(atlas_rt_comm* c, const atlas_rt_scene* scene, const void* rays_in, uint64_t total_count, uint32_t cull_mask, float t_min,
                           float t_max, void* hits_out, uint32_t root, uint32_t flags, int any_hit) {
    if (!c || !scene || root >= c->world) return fail(c ? c->ctx : nullptr, ATLAS_RT_ERR_INVALID, "bad argument");
    atlas_rt_context* ctx = c->ctx;
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t b = 0, e = 0;
    atlas_rt_shard_range(total_count, c->rank, c->world, 64, &b, &e);
    const uint64_t local = e - b;
    if (local && !rays_in) return fail(ctx, ATLAS_RT_ERR_INVALID, "rays_in is null");
    if (flags & ATLAS_RT_PEER_OUTPUT) return trace_sharded_peer(c, scene, rays_in, total_count, cull_mask, t_min, t_max, hits_out, root, flags, any_hit);
    if (c->rank == root && total_count && !hits_out) return fail(ctx, ATLAS_RT_ERR_INVALID, "hits_out is null on the root");
    const int slot = int(c->calls & 1);
    if (local > c->hitsCapacity) {   // (re)allocate both local buffers; rare
        ATLAS_CUDA(ctx, cudaStreamSynchronize(c->stream));
        ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 2; k++) {
            if (c->hits[k]) cudaFree(c->hits[k]);
            c->hits[k] = nullptr;
            ATLAS_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&c->hits[k]), local * 16));
        }
        c->hitsCapacity = local;
    }
    const bool devOut = (flags & ATLAS_RT_DEVICE_OUTPUT) != 0;
    // the gather that last read this local buffer (two calls ago) must be done before the trace overwrites it
    if (c->calls >= 2) ATLAS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->gathered[slot], 0));
    int rc = ATLAS_RT_OK;
    if (local) {
        const uint32_t tf = (flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_PER_RAY_TMAX | ATLAS_RT_OPACITY)) | ATLAS_RT_DEVICE_OUTPUT | ATLAS_RT_HITS_ONLY | ATLAS_RT_ASYNC;
        rc = any_hit ? atlas_rt_trace_any(ctx, scene, rays_in, local, cull_mask, t_min, t_max, c->hits[slot], tf)
                     : atlas_rt_trace_closest(ctx, scene, rays_in, local, cull_mask, t_min, t_max, c->hits[slot], tf);
        if (rc != ATLAS_RT_OK) return rc;
    }
    ATLAS_CUDA(ctx, cudaEventRecord(c->traced[slot], ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->traced[slot], 0));
    // gather on the comm stream
    float4* gatherTo = nullptr;
    float4* staging = nullptr;
    if (c->rank == root) {
        if (devOut) gatherTo = static_cast<float4*>(hits_out);
        else { ATLAS_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&staging), std::max<uint64_t>(1, total_count) * 16, c->stream)); gatherTo = staging; }
    }
    ATLAS_NCCL(c, nccl().GroupStart());
    if (c->rank == root) {
        for (uint32_t r = 0; r < c->world; r++) {
            if (r == root) continue;
            uint64_t rb = 0, re = 0;
            atlas_rt_shard_range(total_count, r, c->world, 64, &rb, &re);
            if (re > rb) ATLAS_NCCL(c, nccl().Recv(gatherTo + rb, (re - rb) * 16, ncclUint8, int(r), c->comm, c->stream));
        }
    } else if (local) {
        ATLAS_NCCL(c, nccl().Send(c->hits[slot], local * 16, ncclUint8, int(root), c->comm, c->stream));
    }
    ATLAS_NCCL(c, nccl().GroupEnd());
    if (c->rank == root && local) ATLAS_CUDA(ctx, cudaMemcpyAsync(gatherTo + b, c->hits[slot], local * 16, cudaMemcpyDeviceToDevice, c->stream));
    if (staging) {
        ATLAS_CUDA(ctx, cudaMemcpyAsync(hits_out, staging, total_count * 16, cudaMemcpyDeviceToHost, c->stream));
        cudaFreeAsync(staging, c->stream);
    }
    ATLAS_CUDA(ctx, cudaEventRecord(c->gathered[slot], c->stream));
    c->calls++;
    if (!(flags & ATLAS_RT_ASYNC)) {
        ATLAS_CUDA(ctx, cudaStreamSynchronize(c->stream));
        ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return ATLAS_RT_OK;
}

int atlas_rt_comm_peer_hits(atlas_rt_comm* c, const void** hits) {
    if (!c || !hits) return ATLAS_RT_ERR_INVALID;
    *hits = nullptr;
    if (!c->peerReady || c->rank != c->peerRoot || c->peerCalls == 0) return ATLAS_RT_OK;
    *hits = reinterpret_cast<const float4*>(c->window + kWindowFlagBytes) + size_t((c->peerCalls - 1) & 1) * c->peerCapacity;
    return ATLAS_RT_OK;
}

}   // extern "C"
