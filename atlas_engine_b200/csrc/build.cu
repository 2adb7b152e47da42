// build.cu — BLAS / TLAS construction on the GPU, reproducing the split decisions, primitive order and flattened
// numbering of src/engine/volume/BVH.cpp bit for bit.
//
// Reference behaviour being reproduced (paths relative to the reference root):
//   BVH::BVH(aabbs, data, parallel)  BVH.cpp:14-56     root = Build #1 (:249-341): object split with 256 bins, SBVH
//                                                      spatial split attempt, median fallback; below the root Build #2.
//   BVH::BVH(aabbs, parallel)        BVH.cpp:58-101    TLAS: Build #2 (:343-406) everywhere, 64 bins.
//   Build #2                                           leaf iff one ref; FindObjectSplit (:444-528) else
//                                                      PerformMedianSplit (:800-855, std::sort fallback); no depth cap.
//   Flatten                          BVH.cpp:408-441   DFS pre-order, larger-area child first.
//
// GPU formulation. Every leaf holds exactly one reference, so a subtree over n refs has n-1 inner nodes and the
// reference's pre-order numbering can be computed top-down while building: a node with flat index i whose refs occupy
// final slots [b, b+n) gives its first child (the one with the larger surface area) index i+1 and slots
// [b, b+n1), its second child index i+n1 and slots [b+n1, b+n); a child with one ref becomes the leaf pointer ~slot.
// The GPUBVHNode array and the final primitive order are therefore written directly by the build — there is no
// separate flatten pass. Refs are kept as two float4 arrays (min.xyz|idx, max.xyz|-) and every node owns a contiguous
// segment; partitioning is stable (scan based) because ref order inside a node is observable through the median
// fallback's sort (SURVEY.md §7 "hard parts").
//
// Two regimes:
//   big nodes   (more than kSubtreeMax refs, or still in the first levels where more than 32 bins are used): processed
//               level by level, many CTAs per node; per-CTA bin histograms in shared memory (integer min/max/add
//               atomics, exact under any order) merged with global atomics; one warp per node does the SAH sweep;
//               three-kernel stable partition (count, scan, scatter) with coalesced float4 traffic.
//   subtrees    (<= kSubtreeMax refs and <= 32 bins): one CTA builds the whole subtree in shared memory, one warp per
//               node, level-synchronous inside the CTA; only node records and final slots go back to HBM.
// The root of a BLAS additionally runs the spatial-split search in parallel and, if that split wins, the
// order-dependent straddler loop of PerformSpatialSplit (:699-756) sequentially on one thread fed from shared memory.
#include <algorithm>
#include <cstring>
#include <vector>

#include "build_common.cuh"

namespace atlas {
namespace {

#ifndef ATLAS_SUBTREE_MAX
#define ATLAS_SUBTREE_MAX 1024
#define ATLAS_SUBTREE_BLOCK 256
#define ATLAS_SUBTREE_CTAS 2
#endif
constexpr uint32_t kSubtreeMax = ATLAS_SUBTREE_MAX;   // refs a shared-memory subtree CTA can hold
constexpr int kSubCtasPerSM = ATLAS_SUBTREE_CTAS;
constexpr uint32_t kSubtreeBins = 32;    // ... and the bin count it supports (one bin per lane)
constexpr int kSmemBin = 9;              // stride of a bin record in SHARED memory: odd, so lanes on different bins hit different banks
constexpr uint32_t kChunk = 1024;        // refs per CTA pass in the big-node kernels
constexpr int kBigBlock = 256;
constexpr int kSubBlock = ATLAS_SUBTREE_BLOCK;
constexpr int kSubWarps = kSubBlock / 32;

enum TaskKind : int32_t { kObject = 0, kMedian = 1, kDone = 2, kSpatial = 3, kPending = 4 };

struct Task {   // one big node of the current level (64 B)
    float lo[3], hi[3];
    uint32_t start, count, flatIdx, depth;
    int32_t kind, axis;
    uint32_t bin;
    float cutoff;
    uint32_t nFirst, leftIsSecond;
};

struct SmallTask {   // root of a shared-memory subtree (48 B)
    float lo[3], hi[3];
    uint32_t start, count, flatIdx, depth, buf, pad;
};

struct LevelInfo {   // device-resident state of the build; read back after the root decision and at the end
    uint32_t nTasks, nChunks, nNext, nSmall, nMedian, rootNeedSpatial, rootKind, rootLeaf;
    uint32_t totalRefs, negZero, nStraddle, nL0, nR0, levels, overflow, subTicket;
    unsigned long long stats[8];   // [2] duplicates [3] median splits [4] sort fallbacks [5] largest sort fallback
};

struct RootSplit {   // state of the BLAS root decision (Build #1)
    float objCost, spaCost;
    int32_t objAxis, spaAxis;
    uint32_t objBin, spaBin;
    float spaPos;
    float minOverlap;
    int splitBox[12];   // ord: left.lo, left.hi, right.lo, right.hi of the split being performed (grown by atomics)
    float finalBox[12];
    uint32_t nL, nR;
};

// chunk -> (task, first ref, refs in the chunk, offset of the chunk inside the task); written once per level so that the
// chunk kernels start their loads after a single dependent read
struct ChunkInfo {
    uint32_t task, refStart, nValid, off;
};

struct BuildBuffers {
    float4* lo[2];
    float4* hi[2];
    float4* nodes;
    uint32_t* order;
    uint8_t* eon;
    Task* tasks[2];
    SmallTask* small;
    LevelInfo* info;
    RootSplit* root;
    int* bins;          // [task][axis][bin][8]
    int* spaBins;       // [axis][256][8] root spatial bins
    int* medAcc;        // [task][16]
    uint32_t* chunkBase;    // [task]
    uint32_t* chunkFirst;   // [chunk]
    ChunkInfo* chunkInfo;   // [chunk]
    int* rootBox;       // 6 ord ints
    const float* tris;
    uint32_t budget;
    uint32_t cap;
};

__device__ __forceinline__ float comp(const float4& v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// ------------------------------------------------------------------------------------------------ level 0 set-up
__global__ void __launch_bounds__(256)
init_refs(const float* __restrict__ aabbs, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi,
          int* __restrict__ rootBox, LevelInfo* __restrict__ info) {
    chain_begin();
    // refs[i] = {idx = i, aabb = aabbs[i]} and the root box (BVH.cpp:21-31); grid-stride so that only one set of
    // atomics per CTA reaches the six root-box words.
    __shared__ int sBox[6];
    __shared__ int sNeg;
    if (threadIdx.x < 6) sBox[threadIdx.x] = threadIdx.x < 3 ? kOrdEmptyLo : kOrdEmptyHi;
    if (threadIdx.x == 6) sNeg = 0;
    __syncthreads();
    int o[6] = {kOrdEmptyLo, kOrdEmptyLo, kOrdEmptyLo, kOrdEmptyHi, kOrdEmptyHi, kOrdEmptyHi};
    bool neg = false;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* a = aabbs + 6 * size_t(i);
        const float v[6] = {a[0], a[1], a[2], a[3], a[4], a[5]};
        lo[i] = make_float4(v[0], v[1], v[2], __uint_as_float(i));
        hi[i] = make_float4(v[3], v[4], v[5], 0.0f);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            o[k] = min(o[k], ord_from_float(v[k]));
            o[3 + k] = max(o[3 + k], ord_from_float(v[3 + k]));
        }
#pragma unroll
        for (int k = 0; k < 6; k++) neg |= (__float_as_uint(v[k]) == 0x80000000u);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        o[k] = __reduce_min_sync(kFullMask, o[k]);
        o[3 + k] = __reduce_max_sync(kFullMask, o[3 + k]);
    }
    const bool anyNeg = __any_sync(kFullMask, neg);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { atomicMin(&sBox[k], o[k]); atomicMax(&sBox[3 + k], o[3 + k]); }
        if (anyNeg) sNeg = 1;
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&rootBox[threadIdx.x], sBox[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(&rootBox[threadIdx.x], sBox[threadIdx.x]);
    else if (threadIdx.x == 6 && sNeg) atomicOr(&info->negZero, 1u);
}

__global__ void make_root(const int* __restrict__ rootBox, uint32_t n, Task* __restrict__ tasks, LevelInfo* __restrict__ info,
                          RootSplit* __restrict__ root) {
    chain_begin();
    Task t;
    for (int k = 0; k < 3; k++) { t.lo[k] = float_from_ord(rootBox[k]); t.hi[k] = float_from_ord(rootBox[3 + k]); }
    t.start = 0; t.count = n; t.flatIdx = 0; t.depth = 0;
    t.kind = kPending; t.axis = -1; t.bin = 0; t.cutoff = 0.0f; t.nFirst = 0; t.leftIsSecond = 0;
    tasks[0] = t;
    info->nTasks = 1;
    info->totalRefs = n;
    root->minOverlap = __fmul_rn(surface_area(t.lo, t.hi), 10e-6f);   // BVH.cpp:33
}

// TLAS over a single instance — BVH.cpp:80-88 plus the two-entry refs quirk (:346-349 then :415).
__global__ void tlas_single(const float* __restrict__ aabbs, float4* __restrict__ nodes, uint32_t* __restrict__ order,
                            uint8_t* __restrict__ eon) {
    nodes[0] = make_float4(aabbs[0], aabbs[1], aabbs[2], aabbs[3]);
    nodes[1] = make_float4(aabbs[4], aabbs[5], 0.0f, 0.0f);
    nodes[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    nodes[3] = make_float4(__int_as_float(~0), __int_as_float(~0), 0.0f, 0.0f);
    order[0] = 0; order[1] = 0;
    eon[0] = 0; eon[1] = 1;
}

// Root turned into a leaf ("last resort", BVH.cpp:303-305; Flatten then emits no node at all).
__global__ void root_leaf_output(uint32_t n, uint32_t* __restrict__ order, uint8_t* __restrict__ eon) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[i] = i;
    eon[i] = (i == n - 1) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- per-level set-up
__device__ __forceinline__ void fill_chunk(ChunkInfo* __restrict__ chunkInfo, uint32_t* __restrict__ chunkFirst, uint32_t c, uint32_t task,
                                           uint32_t start, uint32_t count, uint32_t k) {
    ChunkInfo ci;
    ci.task = task;
    ci.off = k * kChunk;
    ci.refStart = start + ci.off;
    ci.nValid = min(kChunk, count - ci.off);
    chunkInfo[c] = ci;
    chunkFirst[c] = 0u;
}

__global__ void prepare_level(Task* __restrict__ tasks, LevelInfo* __restrict__ info, uint32_t* __restrict__ chunkBase,
                              ChunkInfo* __restrict__ chunkInfo, uint32_t* __restrict__ chunkFirst, int advance,
                              volatile uint32_t* hostFlag) {
    chain_begin();
    // single CTA: the level's task count (the previous level's nNext when `advance`), then an exclusive scan of
    // ceil(count / kChunk) over the level's tasks, and the chunk table
    __shared__ uint32_t carry;
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t sN, sBigN;
    __shared__ uint32_t sBig[1024];   // tasks of this pass with many chunks: filled by the whole CTA
    if (threadIdx.x == 0) {
        if (advance) info->nTasks = min(info->nNext, 0x7fffffffu);
        sN = info->nTasks;
        // tell the host (pinned memory) that this level has started and how many nodes it has: 1 = none, the tree is done
        *hostFlag = sN + 1u;
        __threadfence_system();
        if (sN) info->levels++;
        carry = 0;
        sBigN = 0;
    }
    __syncthreads();
    const uint32_t n = sN;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t tStart = 0, tCount = 0;
        if (i < n) { tStart = tasks[i].start; tCount = tasks[i].count; }
        const uint32_t v = (tCount + kChunk - 1) / kChunk;
        uint32_t s = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(kFullMask, s, off);
            if ((threadIdx.x & 31) >= off) s += o;
        }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? warpSums[threadIdx.x] : 0u;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(kFullMask, w, off);
                if (threadIdx.x >= off) w += o;
            }
            warpSums[threadIdx.x] = w;   // inclusive
        }
        __syncthreads();
        const uint32_t warpOff = (threadIdx.x >> 5) ? warpSums[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t myBase = carry + warpOff + s - v;
        if (i < n) {
            chunkBase[i] = myBase;
            if (v <= 8u) {
                for (uint32_t k = 0; k < v; k++) fill_chunk(chunkInfo, chunkFirst, myBase + k, i, tStart, tCount, k);
            } else {
                sBig[atomicAdd(&sBigN, 1u)] = i;
            }
        }
        __syncthreads();
        const uint32_t nBig = sBigN;
        for (uint32_t b = 0; b < nBig; b++) {
            const uint32_t t = sBig[b];
            const uint32_t bStart = tasks[t].start, bCount = tasks[t].count, bBase = chunkBase[t];
            const uint32_t bv = (bCount + kChunk - 1) / kChunk;
            for (uint32_t k = threadIdx.x; k < bv; k += blockDim.x) fill_chunk(chunkInfo, chunkFirst, bBase + k, t, bStart, bCount, k);
        }
        if (threadIdx.x == blockDim.x - 1) carry += warpOff + s;
        __syncthreads();
        if (threadIdx.x == 0) sBigN = 0;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        info->nChunks = carry;
        info->nNext = 0;
        info->nMedian = 0;
    }
}

__global__ void init_bins(int* __restrict__ bins, const LevelInfo* __restrict__ info, uint32_t binsPerTask3) {
    chain_begin();
    const uint64_t total = uint64_t(info->nTasks) * binsPerTask3;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x)
        bin_init(bins + i * kBinWords);
}

// ------------------------------------------------------------------------------------------ object-split binning
// FindObjectSplit's hot loop (BVH.cpp:474-482) for all big nodes of a level, all three axes in one pass.
__global__ void __launch_bounds__(kBigBlock)
bin_big(const Task* __restrict__ tasks, const LevelInfo* __restrict__ info, const ChunkInfo* __restrict__ chunkInfo,
        const float4* __restrict__ rlo, const float4* __restrict__ rhi, int* __restrict__ gbins, uint32_t nb) {
    chain_begin();
    extern __shared__ int sb[];   // [3][nb][kSmemBin]
    const uint32_t nChunks = info->nChunks;
    // every CTA takes a contiguous run of chunks and keeps accumulating in shared memory while the run stays inside one
    // node, so the shared bins are merged into the node's global bins once per (CTA, node) instead of once per chunk
    const uint32_t perCta = (nChunks + gridDim.x - 1) / gridDim.x;
    const uint32_t c0 = blockIdx.x * perCta, c1 = min(nChunks, c0 + perCta);
    uint32_t cur = 0xffffffffu;
    AxisBins ab[3];
    auto flush = [&](uint32_t t) {
        __syncthreads();
        int* g = gbins + size_t(t) * 3 * nb * kBinWords;
        for (uint32_t e = threadIdx.x; e < 3 * nb; e += kBigBlock) {
            const int* rec = sb + e * kSmemBin;
            if (rec[6] > 0) {
                int* d = g + e * kBinWords;
                // same monotone-value shortcut as in shared memory, reading past L1 (this level's values only)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (rec[k] < __ldcg(d + k)) atomicMin(d + k, rec[k]);
                    if (rec[3 + k] > __ldcg(d + 3 + k)) atomicMax(d + 3 + k, rec[3 + k]);
                }
                atomicAdd(d + 6, rec[6]);   // exit == enter == primitiveCount for the object split: the selection kernels mirror it
            }
        }
        __syncthreads();
    };
    for (uint32_t c = c0; c < c1; c++) {
        const ChunkInfo ci = chunkInfo[c];
        // all of the thread's refs are requested before anything else happens to them
        float4 l[kChunk / kBigBlock], h[kChunk / kBigBlock];
#pragma unroll
        for (uint32_t r = 0; r < kChunk / kBigBlock; r++) {
            const uint32_t i = r * kBigBlock + threadIdx.x;
            if (i < ci.nValid) { l[r] = rlo[ci.refStart + i]; h[r] = rhi[ci.refStart + i]; }
        }
        if (ci.task != cur) {
            if (cur != 0xffffffffu) flush(cur);
            for (uint32_t e = threadIdx.x; e < 3 * nb; e += kBigBlock) bin_init(sb + e * kSmemBin);
            cur = ci.task;
            const Task& tk = tasks[cur];
#pragma unroll
            for (int a = 0; a < 3; a++) ab[a] = axis_bins(tk.lo[a], tk.hi[a], nb);
            __syncthreads();
        }
#pragma unroll
        for (uint32_t r = 0; r < kChunk / kBigBlock; r++) {
            if (r * kBigBlock + threadIdx.x < ci.nValid) {
                const int ol[3] = {ord_from_float(l[r].x), ord_from_float(l[r].y), ord_from_float(l[r].z)};
                const int oh[3] = {ord_from_float(h[r].x), ord_from_float(h[r].y), ord_from_float(h[r].z)};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (!ab[a].active) continue;
                    const uint32_t b = bin_of(bin_centre(comp(l[r], a), comp(h[r], a)), ab[a].start, ab[a].inv, nb);
                    int* rec = sb + (a * nb + b) * kSmemBin;
                    // min/max only move one way, so a plain read that already covers the value makes the atomic
                    // unnecessary; after the first few refs of a bin almost every atomic is skipped
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if (ol[k] < rec[k]) atomicMin(rec + k, ol[k]);
                        if (oh[k] > rec[3 + k]) atomicMax(rec + 3 + k, oh[k]);
                    }
                    atomicAdd(rec + 6, 1);
                }
            }
        }
    }
    if (cur != 0xffffffffu) flush(cur);
}

// ------------------------------------------------------------------------------------------------ child creation
// ---- wide levels: the part of the tree below the big levels, level by level over the WHOLE tree (ATLAS_RT_BUILD_WIDE=1) ----
// Instead of one CTA per <= 1024-ref subtree with the refs in shared memory (build_subtrees), every node of a level — whatever
// subtree it belongs to — is one entry of a per-size-class list in global memory, the refs stay in the (L2-resident)
// ping-pong arrays, and one kernel per level hands the entries of each class to CTAs / warps / half warps / lanes. No barrier
// ever waits for another node; the only synchronisation is the launch boundary between two levels.
constexpr int kWideClasses = 5;   // 0: 2..4 refs (one lane), 1: 5..8 (one lane), 2: 9..16 with <= 16 bins (half warp), 3: up to 256 (warp), 4: more (CTA)
struct WideLists {
    SmallTask* list[kWideClasses];
    uint32_t* count;              // [kWideClasses], device
    uint32_t cap[kWideClasses];
};
__device__ __forceinline__ int wide_class(uint32_t count, uint32_t bins) {
    if (count <= 4u) return 0;
    if (count <= 8u) return 1;
    if (count <= 16u && bins <= 16u) return 2;
    return count <= 256u ? 3 : 4;
}
struct WideNode {   // what the node builders read: absolute first slot, ref count, absolute flattened index, box
    uint32_t start, count, rel;
    float lo[3], hi[3];
};
struct WideEmit {
    WideLists next;
    LevelInfo* info;
    uint32_t budget, depth, buf;   // of the children: binning budget, depth, which ref buffer holds them
    __device__ __forceinline__ void push(uint32_t start, uint32_t count, uint32_t rel, const Box3& b) const {
        const int c = wide_class(count, bins_at_depth(budget, depth));
        const uint32_t i = atomicAdd(&next.count[c], 1u);
        if (i >= next.cap[c]) { info->overflow = 1u; return; }
        SmallTask st;
#pragma unroll
        for (int k = 0; k < 3; k++) { st.lo[k] = b.lo[k]; st.hi[k] = b.hi[k]; }
        st.start = start; st.count = count; st.flatIdx = rel; st.depth = depth; st.buf = buf; st.pad = 0;
        next.list[c][i] = st;
    }
};

// The same, for the level kernels, where every node of the tree pushes: one same-address atomic per child serialises in the L2
// (measured: 0.8 ns each, 230 us for the 300 k pushes of one level). Lane-per-node code aggregates the lanes of a warp that push to
// the same class (one atomic per warp and class); group code (one pushing lane per warp / half warp / CTA) reserves kWideChunk
// slots at a time and marks what it did not use as holes (count 0) at the end of the kernel.
struct WideEmitAgg {
    WideEmit e;
    __device__ __forceinline__ void push(uint32_t start, uint32_t count, uint32_t rel, const Box3& b) const {
        const int c = wide_class(count, bins_at_depth(e.budget, e.depth));
        const unsigned lane = threadIdx.x & 31u;
        const unsigned peers = __match_any_sync(__activemask(), c);
        const int leader = __ffs(int(peers)) - 1;
        uint32_t base = 0;
        if (int(lane) == leader) base = atomicAdd(&e.next.count[c], uint32_t(__popc(peers)));
        base = __shfl_sync(peers, base, leader);
        const uint32_t i = base + uint32_t(__popc(peers & ((1u << lane) - 1u)));
        if (i >= e.next.cap[c]) { e.info->overflow = 1u; return; }
        SmallTask st;
#pragma unroll
        for (int k = 0; k < 3; k++) { st.lo[k] = b.lo[k]; st.hi[k] = b.hi[k]; }
        st.start = start; st.count = count; st.flatIdx = rel; st.depth = e.depth; st.buf = e.buf; st.pad = 0;
        e.next.list[c][i] = st;
    }
};
constexpr uint32_t kWideChunk = 8;
struct ChunkState { uint32_t base[kWideClasses], used[kWideClasses]; };   // shared memory, one per pushing lane
struct WideEmitChunk {
    WideEmit e;
    ChunkState* cs;
    __device__ __forceinline__ void push(uint32_t start, uint32_t count, uint32_t rel, const Box3& b) const {
        const int c = wide_class(count, bins_at_depth(e.budget, e.depth));
        if (cs->used[c] >= kWideChunk) { cs->base[c] = atomicAdd(&e.next.count[c], kWideChunk); cs->used[c] = 0u; }
        const uint32_t i = cs->base[c] + cs->used[c]++;
        if (i >= e.next.cap[c]) { e.info->overflow = 1u; return; }
        SmallTask st;
#pragma unroll
        for (int k = 0; k < 3; k++) { st.lo[k] = b.lo[k]; st.hi[k] = b.hi[k]; }
        st.start = start; st.count = count; st.flatIdx = rel; st.depth = e.depth; st.buf = e.buf; st.pad = 0;
        e.next.list[c][i] = st;
    }
};
__device__ __forceinline__ void chunk_flush(const WideLists& next, ChunkState* cs) {   // by the lane that pushed through cs
    for (int c = 0; c < kWideClasses; c++)
        for (uint32_t k = cs->used[c]; k < kWideChunk; k++) {
            const uint32_t i = cs->base[c] + k;
            if (i < next.cap[c]) next.list[c][i].count = 0u;
        }
}

struct Lists {
    WideLists wide;
    uint32_t wideOn;
    Task* next;
    SmallTask* small;
    LevelInfo* info;
    float4* nodes;
    uint32_t budget;
    uint32_t nextBuf;
    uint32_t maxTasks, maxSmall;
};

__device__ __forceinline__ void enqueue_child(const Lists& L, const Box3& box, uint32_t start, uint32_t count,
                                              uint32_t flatIdx, uint32_t depth) {
    if (count <= 1u) return;
    if (count <= kSubtreeMax && bins_at_depth(L.budget, depth) <= kSubtreeBins) {
        if (L.wideOn) {
            WideEmit{L.wide, L.info, L.budget, depth, L.nextBuf}.push(start, count, flatIdx, box);
            return;
        }
        const uint32_t s = atomicAdd(&L.info->nSmall, 1u);
        if (s >= L.maxSmall) { L.info->overflow = 1u; return; }
        SmallTask st;
#pragma unroll
        for (int k = 0; k < 3; k++) { st.lo[k] = box.lo[k]; st.hi[k] = box.hi[k]; }
        st.start = start; st.count = count; st.flatIdx = flatIdx; st.depth = depth; st.buf = L.nextBuf; st.pad = 0;
        L.small[s] = st;
    } else {
        const uint32_t s = atomicAdd(&L.info->nNext, 1u);
        if (s >= L.maxTasks) { L.info->overflow = 1u; return; }
        Task t;
#pragma unroll
        for (int k = 0; k < 3; k++) { t.lo[k] = box.lo[k]; t.hi[k] = box.hi[k]; }
        t.start = start; t.count = count; t.flatIdx = flatIdx; t.depth = depth;
        t.kind = kPending; t.axis = -1; t.bin = 0; t.cutoff = 0.0f; t.nFirst = 0; t.leftIsSecond = 0;
        L.next[s] = t;
    }
}

// Writes the flattened node of `t` and queues its children. Called by ONE thread. Flatten (BVH.cpp:419-438): the child
// with the larger surface area goes first (swap iff SA(left) < SA(right)); leaf pointer = ~slot.
__device__ __forceinline__ void emit_children(const Lists& L, Task& t, const Box3& left, const Box3& right, uint32_t nLeft) {
    const uint32_t nRight = t.count - nLeft;
    const bool swapped = surface_area(left) < surface_area(right);
    const Box3& first = swapped ? right : left;
    const Box3& second = swapped ? left : right;
    const uint32_t nFirst = swapped ? nRight : nLeft, nSecond = t.count - nFirst;
    const int32_t ptr1 = nFirst > 1u ? int32_t(t.flatIdx + 1u) : ~int32_t(t.start);
    const int32_t ptr2 = nSecond > 1u ? int32_t(t.flatIdx + nFirst) : ~int32_t(t.start + nFirst);
    float4* N = L.nodes + 4 * size_t(t.flatIdx);
    N[0] = make_float4(first.lo[0], first.lo[1], first.lo[2], first.hi[0]);
    N[1] = make_float4(first.hi[1], first.hi[2], second.lo[0], second.lo[1]);
    N[2] = make_float4(second.lo[2], second.hi[0], second.hi[1], second.hi[2]);
    N[3] = make_float4(__int_as_float(ptr1), __int_as_float(ptr2), 0.0f, 0.0f);
    t.nFirst = nFirst;
    t.leftIsSecond = swapped ? 1u : 0u;
    enqueue_child(L, first, t.start, nFirst, t.flatIdx + 1u, t.depth + 1u);
    enqueue_child(L, second, t.start + nFirst, nSecond, t.flatIdx + nFirst, t.depth + 1u);
}

__device__ __forceinline__ void start_median(Task& t, int* acc, LevelInfo* info) {
    int axis;
    float cutoff;
    median_plane(t.lo, t.hi, axis, cutoff);
    t.kind = kMedian;
    t.axis = axis;
    t.cutoff = cutoff;
#pragma unroll
    for (int k = 0; k < 3; k++) { acc[k] = kOrdEmptyLo; acc[3 + k] = kOrdEmptyHi; acc[6 + k] = kOrdEmptyLo; acc[9 + k] = kOrdEmptyHi; }
    acc[12] = 0;
    atomicAdd(&info->nMedian, 1u);
    atomicAdd(&info->stats[3], 1ull);
}

// ------------------------------------------------------------------------------------------- split selection
constexpr int kSelectBlock = 128;
__host__ __device__ constexpr size_t select_smem(uint32_t nb) { return size_t(3) * nb * (kSmemBin + 6) * sizeof(int); }

// Best object / spatial split of one node from its global bins, by a CTA of kSelectBlock threads: the bins are staged
// in shared memory (stride kSmemBin; they stay there for warp_split_boxes), warps 0..2 sweep one axis each, and the
// results are combined in axis order with a strict <, i.e. the lowest axis wins ties like the loop of BVH.cpp:463-523.
__device__ inline BestSplit cta_best_split(const int* __restrict__ gbins, uint32_t nb, const Task& tk, int* sbins, int* ssfx,
                                           BestSplit* sBest, bool objectBins) {
    // (object bins carry one count: exit == enter, word 7 is not maintained in global memory)
    for (uint32_t i = threadIdx.x; i < 3u * nb * kBinWords; i += kSelectBlock)
        sbins[(i >> 3) * kSmemBin + (i & 7u)] = gbins[(objectBins && (i & 7u) == 7u) ? i - 1u : i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5;
    if (warp < 3u) {
        BestSplit b = best_none();
        const AxisBins ab = axis_bins(tk.lo[warp], tk.hi[warp], nb);
        if (ab.active) warp_sweep_axis(sbins + warp * nb * kSmemBin, nb, ssfx + warp * nb * 6, tk.count, int(warp), b, kSmemBin);
        if ((threadIdx.x & 31u) == 0u) sBest[warp] = b;
    }
    __syncthreads();
    BestSplit best = best_none();
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const BestSplit b = sBest[a];
        if (b.axis >= 0 && b.cost < best.cost) best = b;
    }
    return best;
}

// Build #2 decision (BVH.cpp:351-370) for every big node of the level: one CTA per node.
__global__ void __launch_bounds__(kSelectBlock)
select_big(Task* __restrict__ tasks, LevelInfo* __restrict__ info, const int* __restrict__ gbins, int* __restrict__ medAcc, Lists L,
           uint32_t nb) {
    chain_begin();
    extern __shared__ int ss[];   // bins [3][nb][kSmemBin], suffix boxes [3][nb][6]
    __shared__ BestSplit sBest[3];
    const uint32_t t = blockIdx.x;
    if (t >= info->nTasks) return;
    const uint32_t lane = threadIdx.x & 31u;
    Task tk = tasks[t];
    int* sbins = ss;
    const BestSplit best = cta_best_split(gbins + size_t(t) * 3 * nb * kBinWords, nb, tk, sbins, ss + 3 * nb * kSmemBin, sBest, true);
    if (threadIdx.x >= 32u) return;
    const float nodeCost = __fmul_rn(__uint2float_rn(tk.count), surface_area(tk.lo, tk.hi));   // BVH.cpp:238
    if (best.axis < 0 || best.cost >= nodeCost) {
        if (lane == 0) {
            start_median(tk, medAcc + size_t(t) * 16, info);
            tasks[t] = tk;
        }
        return;
    }
    OBox l, r;
    uint32_t nLeft, nExit;
    warp_split_boxes(sbins + size_t(best.axis) * nb * kSmemBin, nb, best.bin, l, r, nLeft, nExit, kSmemBin);
    if (lane == 0) {
        tk.kind = kObject;
        tk.axis = best.axis;
        tk.bin = best.bin;
        emit_children(L, tk, obox_to_box(l), obox_to_box(r), nLeft);
        tasks[t] = tk;
    }
}

// Build #1, first half (BVH.cpp:258-268): object split of the BLAS root and the "try a spatial split?" test.
__global__ void __launch_bounds__(kSelectBlock)
select_root_object(Task* __restrict__ tasks, LevelInfo* __restrict__ info, const int* __restrict__ gbins, RootSplit* __restrict__ root,
                   uint32_t nb) {
    chain_begin();
    extern __shared__ int ss[];
    __shared__ BestSplit sBest[3];
    const uint32_t lane = threadIdx.x & 31u;
    Task tk = tasks[0];
    int* sbins = ss;
    const BestSplit best = cta_best_split(gbins, nb, tk, sbins, ss + 3 * nb * kSmemBin, sBest, true);
    if (threadIdx.x >= 32u) return;
    Box3 l = empty_box(), r = empty_box();
    if (best.axis >= 0) {
        OBox ol, orr;
        uint32_t nLeft, nExit;
        warp_split_boxes(sbins + size_t(best.axis) * nb * kSmemBin, nb, best.bin, ol, orr, nLeft, nExit, kSmemBin);
        l = obox_to_box(ol);
        r = obox_to_box(orr);
    }
    if (lane == 0) {
        root->objCost = best.cost;
        root->objAxis = best.axis;
        root->objBin = best.bin;
        root->spaCost = kFltMax;
        root->spaAxis = -1;
        root->spaBin = 0;
        // overlap = left; overlap.Intersect(right) — AABB.cpp:102-107
        float olo[3], ohi[3];
        for (int k = 0; k < 3; k++) { olo[k] = gl_max(l.lo[k], r.lo[k]); ohi[k] = gl_min(l.hi[k], r.hi[k]); }
        info->rootNeedSpatial = (surface_area(olo, ohi) >= root->minOverlap) ? 1u : 0u;
    }
}

// SplitReference — BVH.cpp:760-798. v = the triangle's vertices; cur = box of the reference being split.
template <int AXIS>
__device__ __forceinline__ void split_reference_t(const float (&v)[3][3], const Box3& cur, Box3& L, Box3& R, float plane) {
    L = empty_box();
    R = empty_box();
#pragma unroll
    for (int e = 0; e < 3; e++) {
        const float* a = v[e];
        const float* b = v[(e + 1) % 3];
        const float av = a[AXIS], bv = b[AXIS];
        if ((av < plane && bv > plane) || (av > plane && bv < plane)) {
            const float off = gl_clamp(__fdiv_rn(__fsub_rn(plane, av), __fsub_rn(bv, av)), 0.0f, 1.0f);
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const float p = __fadd_rn(a[k], __fmul_rn(off, __fsub_rn(b[k], a[k])));   // glm 0.9.8 mix
                L.hi[k] = gl_max(p, L.hi[k]); L.lo[k] = gl_min(p, L.lo[k]);               // Grow(vec3): point first
                R.hi[k] = gl_max(p, R.hi[k]); R.lo[k] = gl_min(p, R.lo[k]);
            }
        }
        if (av <= plane) {
#pragma unroll
            for (int k = 0; k < 3; k++) { L.hi[k] = gl_max(a[k], L.hi[k]); L.lo[k] = gl_min(a[k], L.lo[k]); }
        }
        if (av >= plane) {
#pragma unroll
            for (int k = 0; k < 3; k++) { R.hi[k] = gl_max(a[k], R.hi[k]); R.lo[k] = gl_min(a[k], R.lo[k]); }
        }
    }
    L.hi[AXIS] = plane;
    R.lo[AXIS] = plane;
#pragma unroll
    for (int k = 0; k < 3; k++) {   // Intersect(currentRef.aabb)
        L.lo[k] = gl_max(L.lo[k], cur.lo[k]); L.hi[k] = gl_min(L.hi[k], cur.hi[k]);
        R.lo[k] = gl_max(R.lo[k], cur.lo[k]); R.hi[k] = gl_min(R.hi[k], cur.hi[k]);
    }
}

__device__ inline void split_reference(const float* __restrict__ tri, const Box3& cur, Box3& L, Box3& R, float plane, int axis) {
    float v[3][3];
#pragma unroll
    for (int k = 0; k < 9; k++) v[k / 3][k % 3] = tri[k];
    if (axis == 0) split_reference_t<0>(v, cur, L, R, plane);
    else if (axis == 1) split_reference_t<1>(v, cur, L, R, plane);
    else split_reference_t<2>(v, cur, L, R, plane);
}

// FindSpatialSplit's binning over the root's refs (BVH.cpp:589-619): on every axis a ref's box is chopped at each bin
// boundary it straddles (SplitReference chain), each piece grows its bin, the first / last bin count an entry / an exit.
// Bins in shared memory. One thread per (ref, axis) chain — three neighbouring lanes share a ref, so its loads are
// broadcast — and the chain's axis is moved to component 0 by rotating triangle and box (x,y,z -> y,z,x once or twice).
// SplitReference treats the two other components alike, so the arithmetic is the reference's bit for bit, every lane
// runs the same code whatever its axis, and results are rotated back when they grow a bin.
// Falls through when the object split's children do not overlap enough for a spatial split to be tried (BVH.cpp:282).
constexpr uint32_t kChainsPerBlock = (kBigBlock / 3) * 3;   // 255: the last thread of a block idles

__global__ void __launch_bounds__(kBigBlock, 3)
spatial_bin_root(const Task* __restrict__ tasks, const LevelInfo* __restrict__ info, const float4* __restrict__ rlo,
                 const float4* __restrict__ rhi, const float* __restrict__ tris, int* __restrict__ gbins, uint32_t nb) {
    chain_begin();
    extern __shared__ int sb[];   // [3][nb][kSmemBin]
    if (!info->rootNeedSpatial) return;
    for (uint32_t e = threadIdx.x; e < 3 * nb; e += kBigBlock) bin_init(sb + e * kSmemBin);
    __syncthreads();
    const Task& tk = tasks[0];
    const uint64_t chains = uint64_t(tk.count) * 3u;
    const int axis = int(threadIdx.x % 3u);
    const AxisBins ab = axis_bins(axis == 0 ? tk.lo[0] : (axis == 1 ? tk.lo[1] : tk.lo[2]),
                                  axis == 0 ? tk.hi[0] : (axis == 1 ? tk.hi[1] : tk.hi[2]), nb);
    int* base = sb + axis * nb * kSmemBin;
    // rotated component k is the original component (k + axis) % 3
    const int c0 = axis, c1 = axis == 2 ? 0 : axis + 1, c2 = axis == 0 ? 2 : axis - 1;
    if (threadIdx.x < kChainsPerBlock && ab.active) {
        for (uint64_t c = uint64_t(blockIdx.x) * kChainsPerBlock + threadIdx.x; c < chains; c += uint64_t(gridDim.x) * kChainsPerBlock) {
            const uint32_t p = uint32_t(c / 3u);
            const float4 l = rlo[p], h = rhi[p];
            const float* tri = tris + 9 * size_t(__float_as_uint(l.w));
            float t[9];
#pragma unroll
            for (int k = 0; k < 9; k++) t[k] = tri[k];
            float v[3][3];
            Box3 rest;
#pragma unroll
            for (int e = 0; e < 3; e++) {
                v[e][0] = axis == 0 ? t[3 * e] : (axis == 1 ? t[3 * e + 1] : t[3 * e + 2]);
                v[e][1] = axis == 0 ? t[3 * e + 1] : (axis == 1 ? t[3 * e + 2] : t[3 * e]);
                v[e][2] = axis == 0 ? t[3 * e + 2] : (axis == 1 ? t[3 * e] : t[3 * e + 1]);
            }
            rest.lo[0] = axis == 0 ? l.x : (axis == 1 ? l.y : l.z);
            rest.lo[1] = axis == 0 ? l.y : (axis == 1 ? l.z : l.x);
            rest.lo[2] = axis == 0 ? l.z : (axis == 1 ? l.x : l.y);
            rest.hi[0] = axis == 0 ? h.x : (axis == 1 ? h.y : h.z);
            rest.hi[1] = axis == 0 ? h.y : (axis == 1 ? h.z : h.x);
            rest.hi[2] = axis == 0 ? h.z : (axis == 1 ? h.x : h.y);
            const uint32_t b0 = bin_of(rest.lo[0], ab.start, ab.inv, nb);
            const uint32_t b1 = bin_of(rest.hi[0], ab.start, ab.inv, nb);
            atomicAdd(base + b0 * kSmemBin + 6, 1);
            for (uint32_t j = b0;; j++) {
                Box3 piece = rest;
                if (j < b1) {
                    Box3 cr;
                    const float plane = __fadd_rn(ab.start, __fmul_rn(__uint2float_rn(j + 1u), ab.width));
                    split_reference_t<0>(v, rest, piece, cr, plane);
                    rest = cr;
                }
                int* rec = base + j * kSmemBin;
                const int cc[3] = {c0, c1, c2};
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int lo = ord_from_float(piece.lo[k]), hi = ord_from_float(piece.hi[k]);
                    if (lo < rec[cc[k]]) atomicMin(rec + cc[k], lo);   // monotone values: a covering plain read makes the atomic unnecessary
                    if (hi > rec[3 + cc[k]]) atomicMax(rec + 3 + cc[k], hi);
                }
                if (j >= b1) break;
            }
            atomicAdd(base + b1 * kSmemBin + 7, 1);
        }
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < 3 * nb; e += kBigBlock) {
        const int* rec = sb + e * kSmemBin;
        int* d = gbins + e * kBinWords;
        // a bin's box can be grown without its counters changing (chopped interior pieces), so test the box too
        if (rec[6] | rec[7] | (rec[0] != kOrdEmptyLo) | (rec[3] != kOrdEmptyHi)) {
#pragma unroll
            for (int k = 0; k < 3; k++) { atomicMin(d + k, rec[k]); atomicMax(d + 3 + k, rec[3 + k]); }
            if (rec[6]) atomicAdd(d + 6, rec[6]);
            if (rec[7]) atomicAdd(d + 7, rec[7]);
        }
    }
}

// Build #1, second half (BVH.cpp:276-301): choose between median, spatial and object split for the BLAS root.
__global__ void __launch_bounds__(kSelectBlock)
select_root_final(Task* __restrict__ tasks, LevelInfo* __restrict__ info, const int* __restrict__ objBins,
                  const int* __restrict__ spaBins, int* __restrict__ medAcc, RootSplit* __restrict__ root, Lists L, uint32_t nb) {
    chain_begin();
    extern __shared__ int ss[];
    __shared__ BestSplit sBest[3];
    const uint32_t lane = threadIdx.x & 31u;
    const bool trySpatial = info->rootNeedSpatial != 0u;
    Task tk = tasks[0];
    int* sbins = ss;   // the spatial bins when a spatial split is tried
    BestSplit spa = best_none();
    if (trySpatial) spa = cta_best_split(spaBins, nb, tk, sbins, ss + 3 * nb * kSmemBin, sBest, false);
    if (threadIdx.x >= 32u) return;
    const float objCost = root->objCost;
    const int objAxis = root->objAxis;
    const float nodeCost = __fmul_rn(__uint2float_rn(tk.count), surface_area(tk.lo, tk.hi));   // BVH.cpp:230
    if ((objAxis < 0 || objCost >= nodeCost) && (spa.axis < 0 || spa.cost >= nodeCost)) {
        if (lane == 0) {
            start_median(tk, medAcc, info);
            tasks[0] = tk;
            info->rootKind = kMedian;
        }
        return;
    }
    if (spa.cost < objCost) {
        OBox l, r;
        uint32_t nEnter, nExit;
        warp_split_boxes(sbins + size_t(spa.axis) * nb * kSmemBin, nb, spa.bin, l, r, nEnter, nExit, kSmemBin);
        if (lane == 0) {
            const AxisBins ab = axis_bins(tk.lo[spa.axis], tk.hi[spa.axis], nb);
            tk.kind = kSpatial;
            tk.axis = spa.axis;
            tk.bin = spa.bin;
            tasks[0] = tk;
            root->spaCost = spa.cost;
            root->spaAxis = spa.axis;
            root->spaBin = spa.bin;
            root->spaPos = __fadd_rn(ab.start, __fmul_rn(__uint2float_rn(spa.bin), ab.width));   // BVH.cpp:656
            for (int k = 0; k < 3; k++) {
                root->splitBox[k] = l.lo[k]; root->splitBox[3 + k] = l.hi[k];
                root->splitBox[6 + k] = r.lo[k]; root->splitBox[9 + k] = r.hi[k];
            }
            info->rootKind = kSpatial;
            atomicAdd(&info->stats[1], 1ull);
        }
        return;
    }
    // objCost <= spaCost
    OBox l, r;
    uint32_t nLeft, nExit;
    warp_split_boxes(objBins + size_t(objAxis) * nb * kBinWords, nb, root->objBin, l, r, nLeft, nExit);
    if (lane == 0) {
        tk.kind = kObject;
        tk.axis = objAxis;
        tk.bin = root->objBin;
        emit_children(L, tk, obox_to_box(l), obox_to_box(r), nLeft);
        tasks[0] = tk;
        info->rootKind = kObject;
    }
}

// --------------------------------------------------------------------------------------------- median split
// First loop of PerformMedianSplit (BVH.cpp:813-823) for big nodes flagged kMedian: side counts and side boxes.
__global__ void __launch_bounds__(kBigBlock)
median_reduce_big(const Task* __restrict__ tasks, const LevelInfo* __restrict__ info, const ChunkInfo* __restrict__ chunkInfo,
                  const float4* __restrict__ rlo, const float4* __restrict__ rhi, int* __restrict__ medAcc) {
    chain_begin();
    if (info->nMedian == 0) return;
    const uint32_t nChunks = info->nChunks;
    for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x) {
        const ChunkInfo ci = chunkInfo[c];
        const uint32_t t = ci.task;
        const Task& tk = tasks[t];
        if (tk.kind != kMedian) continue;
        const uint32_t off = ci.off;
        const uint32_t end = tk.start + tk.count;
        OBox L = obox_empty(), R = obox_empty();
        uint32_t nL = 0;
#pragma unroll
        for (uint32_t r = 0; r < kChunk / kBigBlock; r++) {
            const uint32_t p = tk.start + off + r * kBigBlock + threadIdx.x;
            if (p < end) {
                const float4 l = rlo[p], h = rhi[p];
                OBox b;
                b.lo[0] = ord_from_float(l.x); b.lo[1] = ord_from_float(l.y); b.lo[2] = ord_from_float(l.z);
                b.hi[0] = ord_from_float(h.x); b.hi[1] = ord_from_float(h.y); b.hi[2] = ord_from_float(h.z);
                if (median_centre(comp(l, tk.axis), comp(h, tk.axis)) < tk.cutoff) { obox_grow(L, b); nL++; }
                else obox_grow(R, b);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            L.lo[k] = __reduce_min_sync(kFullMask, L.lo[k]); L.hi[k] = __reduce_max_sync(kFullMask, L.hi[k]);
            R.lo[k] = __reduce_min_sync(kFullMask, R.lo[k]); R.hi[k] = __reduce_max_sync(kFullMask, R.hi[k]);
        }
        nL = __reduce_add_sync(kFullMask, nL);
        if ((threadIdx.x & 31) == 0) {
            int* acc = medAcc + size_t(t) * 16;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicMin(acc + k, L.lo[k]); atomicMax(acc + 3 + k, L.hi[k]);
                atomicMin(acc + 6 + k, R.lo[k]); atomicMax(acc + 9 + k, R.hi[k]);
            }
            if (nL) atomicAdd(acc + 12, int(nL));
        }
    }
}

// Rest of PerformMedianSplit for big nodes: either accept the cutoff split, or run the std::sort fallback
// (BVH.cpp:826-851) on the node's refs. One warp per task; the sort itself is the single-thread libstdc++ restatement.
__global__ void median_finalize_big(Task* __restrict__ tasks, LevelInfo* __restrict__ info, const int* __restrict__ medAcc,
                                    float4* __restrict__ curLo, float4* __restrict__ curHi, float4* __restrict__ nxtLo,
                                    float4* __restrict__ nxtHi, uint32_t* __restrict__ order, uint8_t* __restrict__ eon,
                                    Lists L, int blasRoot) {
    chain_begin();
    if (info->nMedian == 0) return;
    const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= info->nTasks) return;
    const uint32_t lane = threadIdx.x & 31u;
    Task tk = tasks[t];
    if (tk.kind != kMedian) return;
    const int* acc = medAcc + size_t(t) * 16;
    const uint32_t nL = uint32_t(acc[12]);
    if (nL != 0u && nL != tk.count) {
        if (lane == 0) {
            OBox l, r;
            for (int k = 0; k < 3; k++) { l.lo[k] = acc[k]; l.hi[k] = acc[3 + k]; r.lo[k] = acc[6 + k]; r.hi[k] = acc[9 + k]; }
            emit_children(L, tk, obox_to_box(l), obox_to_box(r), nL);
            tasks[t] = tk;
        }
        return;
    }
    // ---- fallback: sort by extent along the axis, split in the middle
    if (lane == 0) {
        atomicAdd(&info->stats[4], 1ull);
        atomicMax(&info->stats[5], (unsigned long long)tk.count);
        std_sort_refs(RefArray{curLo + tk.start, curHi + tk.start, tk.axis}, int(tk.count));
    }
    __syncwarp();
    const uint32_t half = tk.count / 2u;
    if (blasRoot && tk.depth == 0u && half == 0u) {   // Build #1 "last resort": an empty side makes the root a leaf
        if (lane == 0) { info->rootLeaf = 1u; tk.kind = kDone; tasks[t] = tk; }
        return;
    }
    OBox l = obox_empty(), r = obox_empty();
    for (uint32_t i = lane; i < tk.count; i += 32u) {
        const float4 a = curLo[tk.start + i], b = curHi[tk.start + i];
        OBox x;
        x.lo[0] = ord_from_float(a.x); x.lo[1] = ord_from_float(a.y); x.lo[2] = ord_from_float(a.z);
        x.hi[0] = ord_from_float(b.x); x.hi[1] = ord_from_float(b.y); x.hi[2] = ord_from_float(b.z);
        if (i < half) obox_grow(l, x); else obox_grow(r, x);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        l.lo[k] = __reduce_min_sync(kFullMask, l.lo[k]); l.hi[k] = __reduce_max_sync(kFullMask, l.hi[k]);
        r.lo[k] = __reduce_min_sync(kFullMask, r.lo[k]); r.hi[k] = __reduce_max_sync(kFullMask, r.hi[k]);
    }
    if (lane == 0) {
        emit_children(L, tk, obox_to_box(l), obox_to_box(r), half);
        tk.kind = kDone;
        tasks[t] = tk;
    }
    const uint32_t nFirst = __shfl_sync(kFullMask, tk.nFirst, 0);
    const uint32_t swapped = __shfl_sync(kFullMask, tk.leftIsSecond, 0);
    // copy the sorted refs into the next buffer in flattened arrangement: [first child | second child]
    for (uint32_t i = lane; i < tk.count; i += 32u) {
        const bool isLeft = i < half;
        const bool first = isLeft != (swapped != 0u);
        const uint32_t rank = isLeft ? i : i - half;
        const uint32_t dst = tk.start + (first ? rank : nFirst + rank);
        const float4 a = curLo[tk.start + i], b = curHi[tk.start + i];
        nxtLo[dst] = a;
        nxtHi[dst] = b;
        const uint32_t childCount = first ? nFirst : tk.count - nFirst;
        if (childCount == 1u) { order[dst] = __float_as_uint(a.w); eon[dst] = 1; }
    }
}

// ------------------------------------------------------------------------------------------------ partition
// PerformObjectSplit (BVH.cpp:541-554) / the cutoff loop of PerformMedianSplit (:813-823): which side a ref goes to.
__device__ __forceinline__ bool goes_left(const Task& tk, const AxisBins& ab, uint32_t nb, const float4& l, const float4& h) {
    if (tk.kind == kObject) return bin_of(bin_centre(comp(l, tk.axis), comp(h, tk.axis)), ab.start, ab.inv, nb) < tk.bin;
    return median_centre(comp(l, tk.axis), comp(h, tk.axis)) < tk.cutoff;
}

// The partition kernels give every warp a contiguous 128-ref slice of the chunk (4 coalesced rounds of 32 refs).
constexpr uint32_t kWarpSlice = kChunk / (kBigBlock / 32);
static_assert(kWarpSlice == 128, "partition kernels assume 4 rounds of 32 refs per warp");

__global__ void __launch_bounds__(kBigBlock)
partition_count(const Task* __restrict__ tasks, const LevelInfo* __restrict__ info, const ChunkInfo* __restrict__ chunkInfo,
                const float4* __restrict__ rlo, const float4* __restrict__ rhi, uint32_t* __restrict__ chunkFirst, uint32_t nb) {
    chain_begin();
    const uint32_t nChunks = info->nChunks;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x) {
        const ChunkInfo ci = chunkInfo[c];
        float4 l[4], h[4];
#pragma unroll
        for (uint32_t r = 0; r < 4; r++) {
            const uint32_t i = warp * kWarpSlice + r * 32u + lane;
            if (i < ci.nValid) { l[r] = rlo[ci.refStart + i]; h[r] = rhi[ci.refStart + i]; }
        }
        const Task& tk = tasks[ci.task];
        if (tk.kind != kObject && tk.kind != kMedian) continue;
        const AxisBins ab = axis_bins(tk.lo[tk.axis], tk.hi[tk.axis], nb);
        uint32_t mine = 0;
#pragma unroll
        for (uint32_t r = 0; r < 4; r++) {
            const uint32_t i = warp * kWarpSlice + r * 32u + lane;
            if (i < ci.nValid) mine += (goes_left(tk, ab, nb, l[r], h[r]) != (tk.leftIsSecond != 0u)) ? 1u : 0u;
        }
        mine = __reduce_add_sync(kFullMask, mine);
        if (lane == 0 && mine) atomicAdd(&chunkFirst[c], mine);   // zeroed by prepare_level
    }
}

// Stable partition of every big node into its children's final segments. A chunk's base = the firsts of the chunks of
// the same node before it, summed here from the per-chunk counts (at most a few KB, L2 resident) — no separate scan.
// Also resets the global bins of the NEXT level's nodes, which nothing reads at this point.
__global__ void __launch_bounds__(kBigBlock)
partition_scatter(const Task* __restrict__ tasks, const LevelInfo* __restrict__ info, const ChunkInfo* __restrict__ chunkInfo,
                  const uint32_t* __restrict__ chunkFirst, const float4* __restrict__ rlo, const float4* __restrict__ rhi,
                  float4* __restrict__ wlo, float4* __restrict__ whi, uint32_t* __restrict__ order, uint8_t* __restrict__ eon,
                  uint32_t nb, int* __restrict__ nextBins, uint32_t nextBinsPerTask3, uint32_t nextTaskCap) {
    chain_begin();
    __shared__ uint32_t sWarpFirst[2][kBigBlock / 32];   // double-buffered by iteration: one barrier per chunk
    __shared__ uint32_t sWarpBase[2][kBigBlock / 32];
    {
        const uint64_t total = uint64_t(min(info->nNext, nextTaskCap)) * nextBinsPerTask3;   // (the cap only matters for a build that is failing)
        for (uint64_t i = blockIdx.x * uint64_t(kBigBlock) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * kBigBlock)
            bin_init(nextBins + i * kBinWords);
    }
    const uint32_t nChunks = info->nChunks;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t it = 0;
    for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x) {
        const ChunkInfo ci = chunkInfo[c];
        float4 l[4], h[4];
#pragma unroll
        for (uint32_t r = 0; r < 4; r++) {
            const uint32_t i = warp * kWarpSlice + r * 32u + lane;
            if (i < ci.nValid) { l[r] = rlo[ci.refStart + i]; h[r] = rhi[ci.refStart + i]; }
        }
        const Task& tk = tasks[ci.task];
        if (tk.kind != kObject && tk.kind != kMedian) continue;   // uniform for the CTA
        // firsts of this node's chunks before this one
        const uint32_t before = ci.off / kChunk;
        uint32_t myBase = 0;
        for (uint32_t i = threadIdx.x; i < before; i += kBigBlock) myBase += chunkFirst[c - before + i];
        myBase = __reduce_add_sync(kFullMask, myBase);
        const AxisBins ab = axis_bins(tk.lo[tk.axis], tk.hi[tk.axis], nb);
        const uint32_t nFirst = tk.nFirst, nSecond = tk.count - tk.nFirst;
        unsigned bf[4], bv[4];
        uint32_t myFirst = 0;
#pragma unroll
        for (uint32_t r = 0; r < 4; r++) {
            const uint32_t i = warp * kWarpSlice + r * 32u + lane;
            const bool valid = i < ci.nValid;
            const bool first = valid && (goes_left(tk, ab, nb, l[r], h[r]) != (tk.leftIsSecond != 0u));
            bf[r] = __ballot_sync(kFullMask, first);
            bv[r] = __ballot_sync(kFullMask, valid);
            myFirst += __popc(bf[r]);
        }
        if (lane == 0) { sWarpFirst[it][warp] = myFirst; sWarpBase[it][warp] = myBase; }
        __syncthreads();
        uint32_t wf = 0, baseFirst = 0;   // firsts in the slices of the warps before this one; firsts in the chunks before this one
#pragma unroll
        for (int w = 0; w < kBigBlock / 32; w++) {
            wf += uint32_t(w) < warp ? sWarpFirst[it][w] : 0u;
            baseFirst += sWarpBase[it][w];
        }
        // position inside the task of this warp's first ref = ci.off + warp * kWarpSlice; seconds before = that - firsts before
        uint32_t doneFirst = baseFirst + wf;
        uint32_t doneSecond = ci.off + min(warp * kWarpSlice, ci.nValid) - doneFirst;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (uint32_t r = 0; r < 4; r++) {
            const bool valid = (bv[r] >> lane) & 1u, first = (bf[r] >> lane) & 1u;
            const unsigned bs = bv[r] & ~bf[r];
            if (valid) {
                const uint32_t dst = tk.start + (first ? doneFirst + __popc(bf[r] & lt) : nFirst + doneSecond + __popc(bs & lt));
                wlo[dst] = l[r];
                whi[dst] = h[r];
                if ((first ? nFirst : nSecond) == 1u) { order[dst] = __float_as_uint(l[r].w); eon[dst] = 1; }
            }
            doneFirst += __popc(bf[r]);
            doneSecond += __popc(bs);
        }
        it ^= 1u;   // only iterations that passed the barrier alternate the buffer
    }
}

// --------------------------------------------------------------------------------------- root spatial split
// PerformSpatialSplit, first loop (BVH.cpp:680-697): class of every root ref (0 = left, 1 = right, 2 = straddler),
// per-chunk class counts, and growth of split.leftAABB / rightAABB by the refs that are not straddlers.
__device__ __forceinline__ int spatial_class(const AxisBins& ab, uint32_t nb, uint32_t splitBin, int axis, const float4& l, const float4& h) {
    const uint32_t b0 = bin_of(comp(l, axis), ab.start, ab.inv, nb);
    const uint32_t b1 = bin_of(comp(h, axis), ab.start, ab.inv, nb);
    if (b1 < splitBin) return 0;
    if (b0 >= splitBin) return 1;
    return 2;
}

__global__ void __launch_bounds__(kBigBlock)
spatial_count(const Task* __restrict__ tasks, const float4* __restrict__ rlo, const float4* __restrict__ rhi,
              RootSplit* __restrict__ root, uint32_t* __restrict__ chunkCounts /*[3][nChunks]*/, uint32_t nChunks, uint32_t nb) {
    __shared__ uint32_t sCount[3];
    const Task& tk = tasks[0];
    const AxisBins ab = axis_bins(tk.lo[tk.axis], tk.hi[tk.axis], nb);
    for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 3) sCount[threadIdx.x] = 0;
        __syncthreads();
        OBox L = obox_empty(), R = obox_empty();
        uint32_t n[3] = {0, 0, 0};
#pragma unroll
        for (uint32_t r = 0; r < kChunk / kBigBlock; r++) {
            const uint32_t p = c * kChunk + r * kBigBlock + threadIdx.x;
            if (p < tk.count) {
                const float4 l = rlo[p], h = rhi[p];
                const int cls = spatial_class(ab, nb, tk.bin, tk.axis, l, h);
                OBox b;
                b.lo[0] = ord_from_float(l.x); b.lo[1] = ord_from_float(l.y); b.lo[2] = ord_from_float(l.z);
                b.hi[0] = ord_from_float(h.x); b.hi[1] = ord_from_float(h.y); b.hi[2] = ord_from_float(h.z);
                if (cls == 0) { obox_grow(L, b); n[0]++; }
                else if (cls == 1) { obox_grow(R, b); n[1]++; }
                else n[2]++;
            }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
            L.lo[k] = __reduce_min_sync(kFullMask, L.lo[k]); L.hi[k] = __reduce_max_sync(kFullMask, L.hi[k]);
            R.lo[k] = __reduce_min_sync(kFullMask, R.lo[k]); R.hi[k] = __reduce_max_sync(kFullMask, R.hi[k]);
            n[k] = __reduce_add_sync(kFullMask, n[k]);
        }
        if ((threadIdx.x & 31) == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                atomicMin(&root->splitBox[k], L.lo[k]); atomicMax(&root->splitBox[3 + k], L.hi[k]);
                atomicMin(&root->splitBox[6 + k], R.lo[k]); atomicMax(&root->splitBox[9 + k], R.hi[k]);
                if (n[k]) atomicAdd(&sCount[k], n[k]);
            }
        }
        __syncthreads();
        if (threadIdx.x < 3) chunkCounts[threadIdx.x * nChunks + c] = sCount[threadIdx.x];
    }
}

// exclusive scans of the three class-count arrays (single CTA, three warps)
__global__ void spatial_scan(uint32_t* __restrict__ chunkCounts, uint32_t nChunks, LevelInfo* __restrict__ info) {
    const uint32_t lane = threadIdx.x & 31u, cls = threadIdx.x >> 5;
    if (cls >= 3) return;
    uint32_t* a = chunkCounts + cls * nChunks;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < nChunks; b += 32u) {
        const uint32_t i = b + lane;
        const uint32_t v = i < nChunks ? a[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(kFullMask, s, off);
            if (lane >= uint32_t(off)) s += o;
        }
        if (i < nChunks) a[i] = carry + s - v;
        carry += __shfl_sync(kFullMask, s, 31);
    }
    if (lane == 0) {
        if (cls == 0) info->nL0 = carry;
        if (cls == 1) info->nR0 = carry;
        if (cls == 2) info->nStraddle = carry;
    }
}

// Stable three-way scatter: left refs -> tmpL, right refs -> tmpR, straddlers -> straddle arrays together with the
// two clipped boxes SplitReference gives them at the split plane (BVH.cpp:709-711), precomputed in parallel.
__global__ void __launch_bounds__(kBigBlock)
spatial_scatter(const Task* __restrict__ tasks, const float4* __restrict__ rlo, const float4* __restrict__ rhi,
                const float* __restrict__ tris, const RootSplit* __restrict__ root, const uint32_t* __restrict__ chunkCounts,
                uint32_t nChunks, uint32_t nb, float4* __restrict__ tmpLlo, float4* __restrict__ tmpLhi,
                float4* __restrict__ tmpRlo, float4* __restrict__ tmpRhi, float4* __restrict__ strad /*[6][cap]*/, uint32_t cap) {
    __shared__ uint32_t sWarp[3][kBigBlock / 32];
    const Task& tk = tasks[0];
    const AxisBins ab = axis_bins(tk.lo[tk.axis], tk.hi[tk.axis], nb);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const float plane = root->spaPos;
    for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x) {
        uint32_t base[3] = {chunkCounts[c], chunkCounts[nChunks + c], chunkCounts[2 * nChunks + c]};
        for (uint32_t r = 0; r < kChunk / kBigBlock; r++) {
            const uint32_t p = c * kChunk + r * kBigBlock + threadIdx.x;
            const bool valid = p < tk.count;
            float4 l = make_float4(0, 0, 0, 0), h = l;
            int cls = -1;
            if (valid) { l = rlo[p]; h = rhi[p]; cls = spatial_class(ab, nb, tk.bin, tk.axis, l, h); }
            unsigned bal[3];
#pragma unroll
            for (int k = 0; k < 3; k++) bal[k] = __ballot_sync(kFullMask, cls == k);
            __syncthreads();
            if (lane == 0) { sWarp[0][warp] = __popc(bal[0]); sWarp[1][warp] = __popc(bal[1]); sWarp[2][warp] = __popc(bal[2]); }
            __syncthreads();
            uint32_t before[3] = {0, 0, 0}, total[3] = {0, 0, 0};
#pragma unroll
            for (int w = 0; w < kBigBlock / 32; w++) {
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const uint32_t v = sWarp[k][w];
                    if (uint32_t(w) < warp) before[k] += v;
                    total[k] += v;
                }
            }
            if (valid) {
                const unsigned lt = (1u << lane) - 1u;
                if (cls == 0) { const uint32_t d = base[0] + before[0] + __popc(bal[0] & lt); tmpLlo[d] = l; tmpLhi[d] = h; }
                else if (cls == 1) { const uint32_t d = base[1] + before[1] + __popc(bal[1] & lt); tmpRlo[d] = l; tmpRhi[d] = h; }
                else {
                    const uint32_t d = base[2] + before[2] + __popc(bal[2] & lt);
                    Box3 box, cl, cr;
                    box.lo[0] = l.x; box.lo[1] = l.y; box.lo[2] = l.z;
                    box.hi[0] = h.x; box.hi[1] = h.y; box.hi[2] = h.z;
                    split_reference(tris + 9 * size_t(__float_as_uint(l.w)), box, cl, cr, plane, tk.axis);
                    strad[d] = l;
                    strad[cap + d] = h;
                    strad[2 * size_t(cap) + d] = make_float4(cl.lo[0], cl.lo[1], cl.lo[2], l.w);
                    strad[3 * size_t(cap) + d] = make_float4(cl.hi[0], cl.hi[1], cl.hi[2], 0.0f);
                    strad[4 * size_t(cap) + d] = make_float4(cr.lo[0], cr.lo[1], cr.lo[2], l.w);
                    strad[5 * size_t(cap) + d] = make_float4(cr.hi[0], cr.hi[1], cr.hi[2], 0.0f);
                }
            }
#pragma unroll
            for (int k = 0; k < 3; k++) base[k] += total[k];
        }
    }
}

// PerformSpatialSplit, second loop (BVH.cpp:699-756): order-dependent, so one thread decides while the CTA stages the
// straddlers' precomputed boxes through shared memory, 128 at a time.
constexpr int kSeqTile = 128;
__global__ void __launch_bounds__(kSeqTile)
spatial_sequential(LevelInfo* __restrict__ info, RootSplit* __restrict__ root, const float4* __restrict__ strad, uint32_t cap,
                   float4* __restrict__ tmpLlo, float4* __restrict__ tmpLhi, float4* __restrict__ tmpRlo,
                   float4* __restrict__ tmpRhi) {
    __shared__ float4 tile[6][kSeqTile];
    const uint32_t nS = info->nStraddle;
    uint32_t nL = info->nL0, nR = info->nR0;
    float L[6], R[6];   // lo.xyz, hi.xyz of split.leftAABB / split.rightAABB
    unsigned long long dups = 0;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; k++) { L[k] = float_from_ord(root->splitBox[k]); R[k] = float_from_ord(root->splitBox[6 + k]); }
    }
    for (uint32_t base = 0; base < nS; base += kSeqTile) {
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        if (i < nS) {
#pragma unroll
            for (int k = 0; k < 6; k++) tile[k][threadIdx.x] = strad[size_t(k) * cap + i];
        }
        __syncthreads();
        if (threadIdx.x != 0) continue;
        const uint32_t m = min(uint32_t(kSeqTile), nS - base);
        for (uint32_t s = 0; s < m; s++) {
            const float4 rl = tile[0][s], rh = tile[1][s], cll = tile[2][s], clh = tile[3][s], crl = tile[4][s], crh = tile[5][s];
            const float ref[6] = {rl.x, rl.y, rl.z, rh.x, rh.y, rh.z};
            const float cl[6] = {cll.x, cll.y, cll.z, clh.x, clh.y, clh.z};
            const float cr[6] = {crl.x, crl.y, crl.z, crh.x, crh.y, crh.z};
            float uL[6], uR[6], dL[6], dR[6];
#pragma unroll
            for (int k = 0; k < 3; k++) {   // Grow(AABB): accumulated value first — AABB.cpp:88-93
                uL[k] = gl_min(L[k], ref[k]); uL[3 + k] = gl_max(L[3 + k], ref[3 + k]);
                uR[k] = gl_min(R[k], ref[k]); uR[3 + k] = gl_max(R[3 + k], ref[3 + k]);
                dL[k] = gl_min(L[k], cl[k]); dL[3 + k] = gl_max(L[3 + k], cl[3 + k]);
                dR[k] = gl_min(R[k], cr[k]); dR[3 + k] = gl_max(R[3 + k], cr[3 + k]);
            }
            const float fl = __uint2float_rn(nL), fr = __uint2float_rn(nR);
            const float fl1 = __uint2float_rn(nL + 1u), fr1 = __uint2float_rn(nR + 1u);
            const float sahUL = __fadd_rn(__fmul_rn(surface_area(uL, uL + 3), fl1), __fmul_rn(surface_area(R, R + 3), fr));
            const float sahUR = __fadd_rn(__fmul_rn(surface_area(L, L + 3), fl), __fmul_rn(surface_area(uR, uR + 3), fr1));
            const float sahD = __fadd_rn(__fmul_rn(surface_area(dL, dL + 3), fl1), __fmul_rn(surface_area(dR, dR + 3), fr1));
            const float best = gl_min(sahD, gl_min(sahUR, sahUL));
            if (best == sahUL) {
#pragma unroll
                for (int k = 0; k < 6; k++) L[k] = uL[k];
                tmpLlo[nL] = rl; tmpLhi[nL] = rh; nL++;
            } else if (best == sahUR) {
#pragma unroll
                for (int k = 0; k < 6; k++) R[k] = uR[k];
                tmpRlo[nR] = rl; tmpRhi[nR] = rh; nR++;
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) { L[k] = dL[k]; R[k] = dR[k]; }
                tmpLlo[nL] = cll; tmpLhi[nL] = clh; nL++;
                tmpRlo[nR] = crl; tmpRhi[nR] = crh; nR++;
                dups++;
            }
        }
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; k++) { root->finalBox[k] = L[k]; root->finalBox[6 + k] = R[k]; }
        root->nL = nL;
        root->nR = nR;
        info->totalRefs = nL + nR;
        info->stats[2] = dups;
    }
}

// Children of the spatially split root (or the last-resort leaf when a side ended up empty).
__global__ void spatial_emit(Task* __restrict__ tasks, LevelInfo* __restrict__ info, const RootSplit* __restrict__ root, Lists L) {
    Task tk = tasks[0];
    if (root->nL == 0u || root->nR == 0u) {
        info->rootLeaf = 1u;
        info->totalRefs = tk.count;
        return;
    }
    Box3 l, r;
    for (int k = 0; k < 3; k++) { l.lo[k] = root->finalBox[k]; l.hi[k] = root->finalBox[3 + k]; r.lo[k] = root->finalBox[6 + k]; r.hi[k] = root->finalBox[9 + k]; }
    tk.count = root->nL + root->nR;
    emit_children(L, tk, l, r, root->nL);
    tasks[0] = tk;
}

// copy tmpL / tmpR into the ref buffer in flattened arrangement
__global__ void spatial_place(const Task* __restrict__ tasks, const RootSplit* __restrict__ root, const float4* __restrict__ tmpLlo,
                              const float4* __restrict__ tmpLhi, const float4* __restrict__ tmpRlo, const float4* __restrict__ tmpRhi,
                              float4* __restrict__ wlo, float4* __restrict__ whi, uint32_t* __restrict__ order, uint8_t* __restrict__ eon) {
    const Task& tk = tasks[0];
    const uint32_t nL = root->nL, nR = root->nR, total = nL + nR;
    const bool swapped = tk.leftIsSecond != 0u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const bool isLeft = i < nL;
        const uint32_t rank = isLeft ? i : i - nL;
        const bool first = isLeft != swapped;
        const uint32_t dst = first ? rank : tk.nFirst + rank;
        const float4 a = isLeft ? tmpLlo[rank] : tmpRlo[rank];
        const float4 b = isLeft ? tmpLhi[rank] : tmpRhi[rank];
        wlo[dst] = a;
        whi[dst] = b;
        const uint32_t childCount = first ? tk.nFirst : total - tk.nFirst;
        if (childCount == 1u) { order[dst] = __float_as_uint(a.w); eon[dst] = 1; }
    }
}

// ------------------------------------------------------------------------------------------- subtree kernel
struct alignas(16) SubNode {   // a node of the current level inside a subtree CTA (32 B)
    uint16_t start, count;     // ref range inside the CTA's buffers
    uint16_t rel, pad;         // flattened index relative to the subtree root's
    float lo[3], hi[3];        // the node's box = the child box its parent recorded (BVH.cpp:399-400)
};

__device__ __forceinline__ SubNode make_subnode(uint32_t start, uint32_t count, uint32_t rel, const Box3& b) {
    SubNode n;
    n.start = uint16_t(start); n.count = uint16_t(count); n.rel = uint16_t(rel); n.pad = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { n.lo[k] = b.lo[k]; n.hi[k] = b.hi[k]; }
    return n;
}

// Where a builder puts the children (with more than one ref) of the node it has just split. The shared-memory subtree kernel
// appends them to the CTA's list of the next level; the wide (whole-tree, global-memory) level kernel sorts them into the
// next level's per-size-class lists (WideEmit, below).
struct SmemEmit {
    SubNode* list;
    uint32_t* counter;
    __device__ __forceinline__ void push(uint32_t start, uint32_t count, uint32_t rel, const Box3& b) const {
        list[atomicAdd(counter, 1u)] = make_subnode(start, count, rel, b);
    }
};

constexpr uint32_t kTinyMax = 4;    // nodes with at most this many refs are handled by ONE lane (32 nodes per warp)
constexpr uint32_t kSmallMax = 8;   // ... and so are nodes with up to this many, in packs of their own (longer unrolled code)

// A ref held by one lane: the box as plain floats plus the primitive index.
struct LaneRef {
    float lo[3], hi[3];
    uint32_t idx;
};

// One lane builds one node with 2..MAXN refs. Same decisions as the warp path, evaluated without bins: with n refs at
// most n-1 bin boundaries separate them, and every other candidate j of BVH.cpp:496-522 repeats the partition (and hence
// the cost) of the nearest boundary below it, so scanning only the boundaries in ascending j and keeping strictly
// smaller costs selects the same (axis, j). Boxes are grown with float min/max here: no atomics are involved, and on
// inputs without -0.0 (the documented exception, flagged by the builder) fminf/fmaxf and the ordered-int reductions of
// the other paths give the same bits.
template <uint32_t MAXN, typename NodeT, typename Emit>
__device__ inline void build_tiny_node(const NodeT nd, const SmallTask& task, uint32_t depth, uint32_t budget,
                                       const float4* __restrict__ cLo, const float4* __restrict__ cHi, float4* __restrict__ nLo,
                                       float4* __restrict__ nHi, float4* nodes, uint32_t* __restrict__ order, uint8_t* __restrict__ eon,
                                       const Emit em, unsigned long long* sStats) {
    const uint32_t n = nd.count, s = nd.start;
    const uint32_t nb = bins_at_depth(budget, depth);
    const uint32_t flatIdx = task.flatIdx + nd.rel;
    float blo[3], bhi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { blo[k] = nd.lo[k]; bhi[k] = nd.hi[k]; }
    LaneRef r[MAXN];
#pragma unroll
    for (uint32_t i = 0; i < MAXN; i++) {
        if (i < n) {
            const float4 l = cLo[s + i], h = cHi[s + i];
            r[i].lo[0] = l.x; r[i].lo[1] = l.y; r[i].lo[2] = l.z;
            r[i].hi[0] = h.x; r[i].hi[1] = h.y; r[i].hi[2] = h.z;
            r[i].idx = __float_as_uint(l.w);
        } else {
#pragma unroll
            for (int k = 0; k < 3; k++) { r[i].lo[k] = kFltMax; r[i].hi[k] = -kFltMax; }
            r[i].idx = 0;
        }
    }
    auto grow = [](Box3& b, const LaneRef& x) {
#pragma unroll
        for (int k = 0; k < 3; k++) { b.lo[k] = fminf(b.lo[k], x.lo[k]); b.hi[k] = fmaxf(b.hi[k], x.hi[k]); }
    };
    // ---- FindObjectSplit over the bin boundaries that actually separate refs
    float bestCost = kFltMax;
    int bestAxis = -1;
    uint32_t bestJ = 0;
#pragma unroll 1
    for (int a = 0; a < 3; a++) {
        const AxisBins ab = axis_bins(blo[a], bhi[a], nb);
        if (!ab.active) continue;
        uint32_t b[MAXN];
#pragma unroll
        for (uint32_t i = 0; i < MAXN; i++) {
            const float lo = a == 0 ? r[i].lo[0] : (a == 1 ? r[i].lo[1] : r[i].lo[2]);
            const float hi = a == 0 ? r[i].hi[0] : (a == 1 ? r[i].hi[1] : r[i].hi[2]);
            b[i] = i < n ? bin_of(bin_centre(lo, hi), ab.start, ab.inv, nb) : 0xffffffffu;
        }
        uint32_t prevJ = 0;
#pragma unroll 1
        for (uint32_t c = 0; c + 1 < n; c++) {
            uint32_t j = 0xffffffffu;
#pragma unroll
            for (uint32_t i = 0; i < MAXN; i++) if (i < n && b[i] + 1u > prevJ && b[i] + 1u < j) j = b[i] + 1u;
            if (j == 0xffffffffu) break;
            prevJ = j;
            Box3 L = empty_box(), R = empty_box();
            uint32_t nL = 0, nR = 0;
#pragma unroll
            for (uint32_t i = 0; i < MAXN; i++) {
                if (i < n) { if (b[i] < j) { grow(L, r[i]); nL++; } else { grow(R, r[i]); nR++; } }
            }
            if (nL == 0u || nR == 0u) continue;
            const float cost = __fadd_rn(__fmul_rn(surface_area(L), __uint2float_rn(nL)), __fmul_rn(surface_area(R), __uint2float_rn(nR)));
            if (cost < bestCost) { bestCost = cost; bestAxis = a; bestJ = j; }
        }
    }
    const float nodeCost = __fmul_rn(__uint2float_rn(n), surface_area(blo, bhi));
    bool isLeft[MAXN];
    if (!(bestAxis < 0 || bestCost >= nodeCost)) {
        const AxisBins ab = axis_bins(blo[bestAxis], bhi[bestAxis], nb);
#pragma unroll
        for (uint32_t i = 0; i < MAXN; i++) {
            const float lo = bestAxis == 0 ? r[i].lo[0] : (bestAxis == 1 ? r[i].lo[1] : r[i].lo[2]);
            const float hi = bestAxis == 0 ? r[i].hi[0] : (bestAxis == 1 ? r[i].hi[1] : r[i].hi[2]);
            isLeft[i] = i < n && bin_of(bin_centre(lo, hi), ab.start, ab.inv, nb) < bestJ;
        }
    } else {
        // ---- PerformMedianSplit
        int axis;
        float cutoff;
        median_plane(blo, bhi, axis, cutoff);
        atomicAdd(&sStats[0], 1ull);
        uint32_t nL = 0;
        float key[MAXN];   // extent on the axis, the sort key of the fallback
#pragma unroll
        for (uint32_t i = 0; i < MAXN; i++) {
            const float lo = axis == 0 ? r[i].lo[0] : (axis == 1 ? r[i].lo[1] : r[i].lo[2]);
            const float hi = axis == 0 ? r[i].hi[0] : (axis == 1 ? r[i].hi[1] : r[i].hi[2]);
            isLeft[i] = i < n && median_centre(lo, hi) < cutoff;
            nL += isLeft[i] ? 1u : 0u;
            key[i] = i < n ? __fsub_rn(hi, lo) : 0.0f;
        }
        if (nL == 0u || nL == n) {
            // std::sort on <= 16 elements is a plain insertion sort, i.e. THE stable order by key (extent on the axis)
            atomicAdd(&sStats[1], 1ull);
            atomicMax(&sStats[2], (unsigned long long)n);
#pragma unroll
            for (uint32_t i = 1; i < MAXN; i++) {
#pragma unroll
                for (uint32_t k = i; k > 0; k--) {
                    if (i < n && key[k] < key[k - 1]) {
                        const float tk = key[k]; key[k] = key[k - 1]; key[k - 1] = tk;
                        const LaneRef tr = r[k]; r[k] = r[k - 1]; r[k - 1] = tr;
                    }
                }
            }
            const uint32_t half = n / 2u;
#pragma unroll
            for (uint32_t i = 0; i < MAXN; i++) isLeft[i] = i < half;
        }
    }
    Box3 lb = empty_box(), rbx = empty_box();
    uint32_t nLeft = 0;
#pragma unroll
    for (uint32_t i = 0; i < MAXN; i++) {
        if (i < n) { if (isLeft[i]) { grow(lb, r[i]); nLeft++; } else grow(rbx, r[i]); }
    }
    // ---- Flatten bookkeeping + stable placement
    const bool swapped = surface_area(lb) < surface_area(rbx);
    const uint32_t nFirst = swapped ? n - nLeft : nLeft, nSecond = n - nFirst;
    const uint32_t slot = task.start + s;
    const Box3& f = swapped ? rbx : lb;
    const Box3& g = swapped ? lb : rbx;
    const int32_t ptr1 = nFirst > 1u ? int32_t(flatIdx + 1u) : ~int32_t(slot);
    const int32_t ptr2 = nSecond > 1u ? int32_t(flatIdx + nFirst) : ~int32_t(slot + nFirst);
    float4* N = nodes + 4 * size_t(flatIdx);
    N[0] = make_float4(f.lo[0], f.lo[1], f.lo[2], f.hi[0]);
    N[1] = make_float4(f.hi[1], f.hi[2], g.lo[0], g.lo[1]);
    N[2] = make_float4(g.lo[2], g.hi[0], g.hi[1], g.hi[2]);
    N[3] = make_float4(__int_as_float(ptr1), __int_as_float(ptr2), 0.0f, 0.0f);
    if (nFirst > 1u) em.push(s, nFirst, nd.rel + 1u, f);
    if (nSecond > 1u) em.push(s + nFirst, nSecond, nd.rel + nFirst, g);
    uint32_t doneFirst = 0, doneSecond = 0;
#pragma unroll
    for (uint32_t i = 0; i < MAXN; i++) {
        if (i < n) {
            const bool first = isLeft[i] != swapped;
            const uint32_t dst = first ? doneFirst++ : nFirst + doneSecond++;
            if ((first ? nFirst : nSecond) == 1u) {
                order[slot + dst] = r[i].idx;
                eon[slot + dst] = 1;
            } else {
                nLo[s + dst] = make_float4(r[i].lo[0], r[i].lo[1], r[i].lo[2], __uint_as_float(r[i].idx));
                nHi[s + dst] = make_float4(r[i].hi[0], r[i].hi[1], r[i].hi[2], 0.0f);
            }
        }
    }
}

__device__ __forceinline__ OBox obox_of_ref(const float4& l, const float4& h) {
    OBox b;
    b.lo[0] = ord_from_float(l.x); b.lo[1] = ord_from_float(l.y); b.lo[2] = ord_from_float(l.z);
    b.hi[0] = ord_from_float(h.x); b.hi[1] = ord_from_float(h.y); b.hi[2] = ord_from_float(h.z);
    return b;
}

// A group of G lanes (a warp, or a half warp for nodes with at most 16 refs once the level has at most 16 bins) builds
// one node of a subtree: Build #2 (BVH.cpp:343-406) = FindObjectSplit else PerformMedianSplit, then the flattened node
// record and the stable partition. The three axes are binned one after the other in the group's single-axis bin array
// (shared-memory atomics on the ordered-int image) and swept with one bin per lane (group_sweep_single).
template <int G, typename NodeT, typename Emit>
__device__ inline void build_group_node(const NodeT nd, const SmallTask& task, uint32_t depth, uint32_t budget, float4* cLo, float4* cHi,
                                        float4* __restrict__ nLo, float4* __restrict__ nHi, float4* nodes, uint32_t* __restrict__ order,
                                        uint8_t* __restrict__ eon, const Emit em, unsigned long long* sStats,
                                        int* bins /*[32][kSubBinWords]*/) {
    const LaneGroup<G> g;
    const uint32_t lane = g.lane;
    const uint32_t n = nd.count, s = nd.start;
    const uint32_t nb = bins_at_depth(budget, depth);   // <= kSubtreeBins by construction, <= 16 for a half warp
    const uint32_t flatIdx = task.flatIdx + nd.rel;
    float blo[3], bhi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { blo[k] = nd.lo[k]; bhi[k] = nd.hi[k]; }
    AxisBins ab[3];
#pragma unroll
    for (int a = 0; a < 3; a++) ab[a] = axis_bins(blo[a], bhi[a], nb);
    // a node with at most one ref per lane (the common case) keeps it in registers for the three axes and the partition
    const bool oneEach = n <= uint32_t(G);
    float4 l0 = make_float4(0, 0, 0, 0), h0 = l0;
    OBox o0 = obox_empty();
    if (oneEach && lane < n) {
        l0 = cLo[s + lane];
        h0 = cHi[s + lane];
        o0 = obox_of_ref(l0, h0);
    }
    // ---- FindObjectSplit
    BestSplit best = best_none();
    OBox objL = obox_empty(), objR = obox_empty();
    uint32_t objLeft = 0;
    bool swept = false;
    if constexpr (G == 32) {
        if (nb <= 16u) {
            // a whole warp and at most 16 bins: the three axes are binned in one pass over the refs and swept together
            swept = true;
            const uint32_t active = (ab[0].active ? 1u : 0u) | (ab[1].active ? 2u : 0u) | (ab[2].active ? 4u : 0u);
            for (uint32_t e = lane; e < 48u; e += 32u) sub_bin_init(bins + e * kSubBinWords);
            g.sync();
            for (uint32_t i = lane; i < n; i += 32u) {
                float4 l = l0, h = h0;
                OBox o = o0;
                if (!oneEach) { l = cLo[s + i]; h = cHi[s + i]; o = obox_of_ref(l, h); }
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (!ab[a].active) continue;
                    const uint32_t b = bin_of(bin_centre(comp(l, a), comp(h, a)), ab[a].start, ab[a].inv, nb);
                    int* rec = bins + (a * 16 + b) * kSubBinWords;
                    atomicMin(rec + 0, o.lo[0]); atomicMin(rec + 1, o.lo[1]); atomicMin(rec + 2, o.lo[2]);
                    atomicMax(rec + 3, o.hi[0]); atomicMax(rec + 4, o.hi[1]); atomicMax(rec + 5, o.hi[2]);
                    atomicAdd(rec + 6, 1);
                }
            }
            g.sync();
            warp_sweep3_rev(bins, nb, n, active, best, objL, objR, objLeft);
            g.sync();
        }
    }
#pragma unroll 1
    for (int a = 0; a < 3 && !swept; a++) {
        if (!ab[a].active) continue;
        for (uint32_t e = lane; e < nb; e += G) sub_bin_init(bins + e * kSubBinWords);
        g.sync();
        if (oneEach) {
            if (lane < n) {
                const uint32_t b = bin_of(bin_centre(comp(l0, a), comp(h0, a)), ab[a].start, ab[a].inv, nb);
                int* rec = bins + b * kSubBinWords;
                atomicMin(rec + 0, o0.lo[0]); atomicMin(rec + 1, o0.lo[1]); atomicMin(rec + 2, o0.lo[2]);
                atomicMax(rec + 3, o0.hi[0]); atomicMax(rec + 4, o0.hi[1]); atomicMax(rec + 5, o0.hi[2]);
                atomicAdd(rec + 6, 1);
            }
        } else {
            for (uint32_t i = lane; i < n; i += G) {
                const float4 l = cLo[s + i], h = cHi[s + i];
                const uint32_t b = bin_of(bin_centre(comp(l, a), comp(h, a)), ab[a].start, ab[a].inv, nb);
                int* rec = bins + b * kSubBinWords;
                atomicMin(rec + 0, ord_from_float(l.x)); atomicMin(rec + 1, ord_from_float(l.y)); atomicMin(rec + 2, ord_from_float(l.z));
                atomicMax(rec + 3, ord_from_float(h.x)); atomicMax(rec + 4, ord_from_float(h.y)); atomicMax(rec + 5, ord_from_float(h.z));
                atomicAdd(rec + 6, 1);
            }
        }
        g.sync();
        if constexpr (G == 32) {
            if (nb <= 16u) group_sweep_single<32, true>(g, bins, nb, n, a, best, objL, objR, objLeft);
            else group_sweep_single<32, false>(g, bins, nb, n, a, best, objL, objR, objLeft);
        } else {
            group_sweep_single<G, false>(g, bins, nb, n, a, best, objL, objR, objLeft);
        }
        g.sync();
    }
    const float nodeCost = __fmul_rn(__uint2float_rn(n), surface_area(blo, bhi));
    // mode: 0 = object split on (axis, bin), 1 = median cutoff, 2 = sorted, split by position
    int mode, axis = 0;
    uint32_t splitBin = 0, nLeft = 0;
    float cutoff = 0.0f;
    OBox L = obox_empty(), R = obox_empty();
    if (!(best.axis < 0 || best.cost >= nodeCost)) {
        mode = 0;
        axis = best.axis;
        splitBin = best.bin;
        L = objL;
        R = objR;
        nLeft = objLeft;
    } else {
        mode = 1;
        median_plane(blo, bhi, axis, cutoff);
        uint32_t cnt = 0;
        for (uint32_t i = lane; i < n; i += G) {
            const float4 l = cLo[s + i], h = cHi[s + i];
            if (median_centre(comp(l, axis), comp(h, axis)) < cutoff) { obox_grow(L, obox_of_ref(l, h)); cnt++; } else obox_grow(R, obox_of_ref(l, h));
        }
        g.reduce(L);
        g.reduce(R);
        nLeft = g.radd(cnt);
    }
    if (mode == 1) {
        if (lane == 0) atomicAdd(&sStats[0], 1ull);
        if (nLeft == 0u || nLeft == n) {
            mode = 2;
            if (lane == 0) {
                atomicAdd(&sStats[1], 1ull);
                atomicMax(&sStats[2], (unsigned long long)n);
                std_sort_refs(RefArray{cLo + s, cHi + s, axis}, int(n));
            }
            g.sync();
            nLeft = n / 2u;
            L = obox_empty();
            R = obox_empty();
            for (uint32_t i = lane; i < n; i += G) {
                const OBox b = obox_of_ref(cLo[s + i], cHi[s + i]);
                if (i < nLeft) obox_grow(L, b); else obox_grow(R, b);
            }
            g.reduce(L);
            g.reduce(R);
        }
    }
    // ---- Flatten bookkeeping: larger-area child first
    const Box3 lb = obox_to_box(L), rb = obox_to_box(R);
    const bool swapped = surface_area(lb) < surface_area(rb);
    const uint32_t nFirst = swapped ? n - nLeft : nLeft, nSecond = n - nFirst;
    const uint32_t slot = task.start + s;
    if (lane == 0) {
        const Box3& f = swapped ? rb : lb;
        const Box3& h2 = swapped ? lb : rb;
        const int32_t ptr1 = nFirst > 1u ? int32_t(flatIdx + 1u) : ~int32_t(slot);
        const int32_t ptr2 = nSecond > 1u ? int32_t(flatIdx + nFirst) : ~int32_t(slot + nFirst);
        float4* N = nodes + 4 * size_t(flatIdx);
        N[0] = make_float4(f.lo[0], f.lo[1], f.lo[2], f.hi[0]);
        N[1] = make_float4(f.hi[1], f.hi[2], h2.lo[0], h2.lo[1]);
        N[2] = make_float4(h2.lo[2], h2.hi[0], h2.hi[1], h2.hi[2]);
        N[3] = make_float4(__int_as_float(ptr1), __int_as_float(ptr2), 0.0f, 0.0f);
        if (nFirst > 1u) em.push(s, nFirst, nd.rel + 1u, f);
        if (nSecond > 1u) em.push(s + nFirst, nSecond, nd.rel + nFirst, h2);
    }
    // ---- stable partition into the other buffer (or straight to the final slot for one-ref children)
    uint32_t doneFirst = 0, doneSecond = 0;
    for (uint32_t b = 0; b < n; b += G) {
        const uint32_t i = b + lane;
        const bool valid = i < n;
        float4 l = make_float4(0, 0, 0, 0), h = l;
        bool first = false;
        if (valid) {
            // (mode 2 has re-sorted the refs in shared memory, so the registers are stale there)
            l = (oneEach && mode != 2) ? l0 : cLo[s + i];
            h = (oneEach && mode != 2) ? h0 : cHi[s + i];
            bool isLeft;
            if (mode == 0) isLeft = bin_of(bin_centre(comp(l, axis), comp(h, axis)), ab[axis].start, ab[axis].inv, nb) < splitBin;
            else if (mode == 1) isLeft = median_centre(comp(l, axis), comp(h, axis)) < cutoff;
            else isLeft = i < nLeft;
            first = isLeft != swapped;
        }
        const unsigned bf = g.ballot(valid && first);
        const unsigned bs = g.ballot(valid && !first);
        if (valid) {
            const unsigned lt = (1u << lane) - 1u;
            const uint32_t dst = first ? doneFirst + __popc(bf & lt) : nFirst + doneSecond + __popc(bs & lt);
            if ((first ? nFirst : nSecond) == 1u) {
                order[slot + dst] = __float_as_uint(l.w);
                eon[slot + dst] = 1;
            } else {
                nLo[s + dst] = l;
                nHi[s + dst] = h;
            }
        }
        doneFirst += __popc(bf);
        doneSecond += __popc(bs);
    }
}

constexpr uint32_t kCtaNodeMin = 256;   // nodes with more refs than this are built by the whole CTA (the top one or two levels)

struct CtaSplit {   // one axis' sweep result, handed from the sweeping warp to the CTA
    float cost;
    uint32_t bin, nLeft;
    int box[12];   // ord: left lo, left hi, right lo, right hi
};

// The whole CTA builds one node: the refs are binned on all three axes in one pass by all threads, warps 0..2 sweep one
// axis each, and the partition gives every warp a contiguous slice. Same decisions as build_group_node; the median
// fallback (rare) is handed to warp 0's group path. Contains barriers: every thread of the CTA must call it.
template <typename NodeT, typename Emit>
__device__ inline void build_cta_node(const NodeT nd, const SmallTask& task, uint32_t depth, uint32_t budget, float4* cLo, float4* cHi,
                                      float4* __restrict__ nLo, float4* __restrict__ nHi, float4* nodes, uint32_t* __restrict__ order,
                                      uint8_t* __restrict__ eon, const Emit em, unsigned long long* sStats,
                                      int* bins3 /*[3][32][kSubBinWords]*/, CtaSplit* sSplit /*[3]*/, uint32_t* sWarpFirst /*[kSubWarps]*/) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n = nd.count, s = nd.start;
    const uint32_t nb = bins_at_depth(budget, depth);
    const uint32_t flatIdx = task.flatIdx + nd.rel;
    float blo[3], bhi[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { blo[k] = nd.lo[k]; bhi[k] = nd.hi[k]; }
    AxisBins ab[3];
#pragma unroll
    for (int a = 0; a < 3; a++) ab[a] = axis_bins(blo[a], bhi[a], nb);
    for (uint32_t e = tid; e < 3u * kSubtreeBins; e += kSubBlock) sub_bin_init(bins3 + e * kSubBinWords);
    __syncthreads();
    // ---- binning, all three axes in one pass
    for (uint32_t i = tid; i < n; i += kSubBlock) {
        const float4 l = cLo[s + i], h = cHi[s + i];
        const OBox o = obox_of_ref(l, h);
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!ab[a].active) continue;
            const uint32_t b = bin_of(bin_centre(comp(l, a), comp(h, a)), ab[a].start, ab[a].inv, nb);
            int* rec = bins3 + (a * kSubtreeBins + b) * kSubBinWords;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (o.lo[k] < rec[k]) atomicMin(rec + k, o.lo[k]);   // monotone: a covering plain read makes the atomic unnecessary
                if (o.hi[k] > rec[3 + k]) atomicMax(rec + 3 + k, o.hi[k]);
            }
            atomicAdd(rec + 6, 1);
        }
    }
    __syncthreads();
    // ---- one warp per axis sweeps its bins
    if (warp < 3u) {
        const LaneGroup<32> g;
        BestSplit b = best_none();
        OBox L = obox_empty(), R = obox_empty();
        uint32_t nl = 0;
        if (ab[warp].active) {
            if (nb <= 16u) group_sweep_single<32, true>(g, bins3 + warp * kSubtreeBins * kSubBinWords, nb, n, int(warp), b, L, R, nl);
            else group_sweep_single<32, false>(g, bins3 + warp * kSubtreeBins * kSubBinWords, nb, n, int(warp), b, L, R, nl);
        }
        if (lane == 0) {
            CtaSplit& o = sSplit[warp];
            o.cost = b.axis >= 0 ? b.cost : kFltMax;
            o.bin = b.bin;
            o.nLeft = nl;
#pragma unroll
            for (int k = 0; k < 3; k++) { o.box[k] = L.lo[k]; o.box[3 + k] = L.hi[k]; o.box[6 + k] = R.lo[k]; o.box[9 + k] = R.hi[k]; }
        }
    }
    __syncthreads();
    // ---- the lowest axis wins ties (the reference's axis loop keeps strictly smaller costs)
    int axis = -1;
    float cost = kFltMax;
#pragma unroll
    for (int a = 0; a < 3; a++)
        if (sSplit[a].cost < cost) { cost = sSplit[a].cost; axis = a; }
    const float nodeCost = __fmul_rn(__uint2float_rn(n), surface_area(blo, bhi));
    if (axis < 0 || cost >= nodeCost) {
        // PerformMedianSplit: rare up here; warp 0 runs the group path on the node (it bins again on its own)
        if (warp == 0u) build_group_node<32>(nd, task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats, bins3);
        __syncthreads();
        return;
    }
    const uint32_t splitBin = sSplit[axis].bin, nLeft = sSplit[axis].nLeft;
    OBox L, R;
#pragma unroll
    for (int k = 0; k < 3; k++) { L.lo[k] = sSplit[axis].box[k]; L.hi[k] = sSplit[axis].box[3 + k]; R.lo[k] = sSplit[axis].box[6 + k]; R.hi[k] = sSplit[axis].box[9 + k]; }
    const Box3 lb = obox_to_box(L), rb = obox_to_box(R);
    const bool swapped = surface_area(lb) < surface_area(rb);
    const uint32_t nFirst = swapped ? n - nLeft : nLeft, nSecond = n - nFirst;
    const uint32_t slot = task.start + s;
    if (tid == 0) {
        const Box3& f = swapped ? rb : lb;
        const Box3& h2 = swapped ? lb : rb;
        const int32_t ptr1 = nFirst > 1u ? int32_t(flatIdx + 1u) : ~int32_t(slot);
        const int32_t ptr2 = nSecond > 1u ? int32_t(flatIdx + nFirst) : ~int32_t(slot + nFirst);
        float4* N = nodes + 4 * size_t(flatIdx);
        N[0] = make_float4(f.lo[0], f.lo[1], f.lo[2], f.hi[0]);
        N[1] = make_float4(f.hi[1], f.hi[2], h2.lo[0], h2.lo[1]);
        N[2] = make_float4(h2.lo[2], h2.hi[0], h2.hi[1], h2.hi[2]);
        N[3] = make_float4(__int_as_float(ptr1), __int_as_float(ptr2), 0.0f, 0.0f);
        if (nFirst > 1u) em.push(s, nFirst, nd.rel + 1u, f);
        if (nSecond > 1u) em.push(s + nFirst, nSecond, nd.rel + nFirst, h2);
    }
    // ---- stable partition: warp w owns refs [w * per, (w + 1) * per)
    const uint32_t per = ((n + kSubBlock - 1u) / kSubBlock) * 32u;
    const uint32_t begin = min(warp * per, n), end = min(begin + per, n);
    uint32_t myFirst = 0;
    for (uint32_t b = begin; b < end; b += 32u) {
        const uint32_t i = b + lane;
        bool first = false;
        if (i < end) {
            const float4 l = cLo[s + i], h = cHi[s + i];
            first = (bin_of(bin_centre(comp(l, axis), comp(h, axis)), ab[axis].start, ab[axis].inv, nb) < splitBin) != swapped;
        }
        myFirst += __popc(__ballot_sync(kFullMask, first));
    }
    if (lane == 0) sWarpFirst[warp] = myFirst;
    __syncthreads();
    uint32_t doneFirst = 0;
#pragma unroll
    for (uint32_t w = 0; w < uint32_t(kSubWarps); w++) doneFirst += w < warp ? sWarpFirst[w] : 0u;
    uint32_t doneSecond = begin - doneFirst;
    for (uint32_t b = begin; b < end; b += 32u) {
        const uint32_t i = b + lane;
        const bool valid = i < end;
        float4 l = make_float4(0, 0, 0, 0), h = l;
        bool first = false;
        if (valid) {
            l = cLo[s + i];
            h = cHi[s + i];
            first = (bin_of(bin_centre(comp(l, axis), comp(h, axis)), ab[axis].start, ab[axis].inv, nb) < splitBin) != swapped;
        }
        const unsigned bf = __ballot_sync(kFullMask, valid && first);
        const unsigned bs = __ballot_sync(kFullMask, valid && !first);
        if (valid) {
            const unsigned lt = (1u << lane) - 1u;
            const uint32_t dst = first ? doneFirst + __popc(bf & lt) : nFirst + doneSecond + __popc(bs & lt);
            if ((first ? nFirst : nSecond) == 1u) {
                order[slot + dst] = __float_as_uint(l.w);
                eon[slot + dst] = 1;
            } else {
                nLo[s + dst] = l;
                nHi[s + dst] = h;
            }
        }
        doneFirst += __popc(bf);
        doneSecond += __popc(bs);
    }
    __syncthreads();   // the scratch areas are free again
}

constexpr uint32_t kSubNodes = kSubtreeMax / 2;   // most nodes one level of a subtree can hold (each has >= 2 refs)
constexpr size_t kSubtreeSmem = size_t(4) * kSubtreeMax * sizeof(float4) + size_t(2) * kSubNodes * sizeof(SubNode) +
                                size_t(kSubWarps) * 2 * kSubtreeBins * kSubBinWords * sizeof(int) + size_t(2) * kSubNodes * sizeof(uint16_t);
static_assert(kSubCtasPerSM * (kSubtreeSmem + 1024 + 128) <= 228 * 1024, "the subtree CTAs must fit one SM");

// Persistent: the grid is sized for the machine (two CTAs per SM) and every CTA draws subtrees from a ticket counter until
// none are left, so the launch does not need the subtree count on the host (no round trip after the level loop).
__global__ void __launch_bounds__(kSubBlock, kSubCtasPerSM)
build_subtrees(const SmallTask* __restrict__ small, float4* const lo0, float4* const hi0, float4* const lo1,
               float4* const hi1, float4* nodes, uint32_t* __restrict__ order, uint8_t* __restrict__ eon,
               LevelInfo* __restrict__ info, uint32_t budget, uint32_t maxSmall) {
    chain_begin();
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float4* sLo = reinterpret_cast<float4*>(smemRaw);                      // [2][kSubtreeMax]
    float4* sHi = sLo + 2 * kSubtreeMax;                                   // [2][kSubtreeMax]
    SubNode* lists = reinterpret_cast<SubNode*>(sHi + 2 * kSubtreeMax);    // [2][kSubNodes]
    int* wBins = reinterpret_cast<int*>(lists + 2 * kSubNodes);            // [warps][2 half-warp groups][32][kSubBinWords]
    uint16_t* clsA = reinterpret_cast<uint16_t*>(wBins + kSubWarps * 2 * kSubtreeBins * kSubBinWords);   // tiny nodes from the front, warp nodes from the back
    uint16_t* clsB = clsA + kSubNodes;                                     // half-warp nodes
    __shared__ uint32_t sNext;
    __shared__ uint32_t sCls[5];               // nodes of the level by size class: <= kTinyMax, half warp, warp, <= kSmallMax, whole CTA
    __shared__ uint16_t sCtaIdx[kSubtreeMax / kCtaNodeMin];
    __shared__ CtaSplit sSplit[3];
    __shared__ uint32_t sWarpFirst[kSubWarps];
    __shared__ unsigned long long sStats[3];   // median splits, sort fallbacks, largest fallback

    __shared__ uint32_t sTicket;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t nSmall = min(info->nSmall, maxSmall);
  while (true) {
    __syncthreads();   // the previous subtree is finished with shared memory
    if (tid == 0) sTicket = atomicAdd(&info->subTicket, 1u);
    __syncthreads();
    if (sTicket >= nSmall) break;
    const SmallTask task = small[sTicket];
    const float4* gLo = task.buf ? lo1 : lo0;
    const float4* gHi = task.buf ? hi1 : hi0;
    for (uint32_t i = tid; i < task.count; i += kSubBlock) {
        sLo[i] = gLo[task.start + i];
        sHi[i] = gHi[task.start + i];
    }
    if (tid == 0) {
        Box3 rootBox;
#pragma unroll
        for (int k = 0; k < 3; k++) { rootBox.lo[k] = task.lo[k]; rootBox.hi[k] = task.hi[k]; }
        lists[0] = make_subnode(0, task.count, 0, rootBox);
        sStats[0] = sStats[1] = sStats[2] = 0;
    }
    uint32_t nCur = 1, cur = 0, level = 0;
    const uint32_t half = lane >> 4;
    int* binsW = wBins + (warp * 2) * kSubtreeBins * kSubBinWords;   // the warp's first group area (a whole-warp group uses it alone)
    int* binsH = binsW + half * kSubtreeBins * kSubBinWords;         // this lane's half-warp group area
    __syncthreads();

    while (nCur > 0) {
        const uint32_t depth = task.depth + level;
        const bool halfOk = bins_at_depth(budget, depth) <= 16u;   // a half warp sweeps one bin per lane
        if (tid == 0) sNext = 0;
        if (tid < 5) sCls[tid] = 0;
        __syncthreads();
        const SubNode* curList = lists + (level & 1u) * kSubNodes;
        const SmemEmit em{lists + ((level + 1u) & 1u) * kSubNodes, &sNext};
        float4* cLo = sLo + cur * kSubtreeMax;
        float4* cHi = sHi + cur * kSubtreeMax;
        float4* nLo = sLo + (cur ^ 1u) * kSubtreeMax;
        float4* nHi = sHi + (cur ^ 1u) * kSubtreeMax;

        // sort the level's nodes into size classes so that lanes / half warps / warps each get a dense run of work
        for (uint32_t ni = tid; ni < nCur; ni += kSubBlock) {
            const uint32_t c = curList[ni].count;
            if (c > kCtaNodeMin) sCtaIdx[atomicAdd(&sCls[4], 1u)] = uint16_t(ni);   // fewer than kSubtreeMax / kCtaNodeMin of them
            else if (c <= kTinyMax) clsA[atomicAdd(&sCls[0], 1u)] = uint16_t(ni);
            else if (c <= kSmallMax) clsB[kSubNodes - 1u - atomicAdd(&sCls[3], 1u)] = uint16_t(ni);
            else if (c <= 16u && halfOk) clsB[atomicAdd(&sCls[1], 1u)] = uint16_t(ni);
            else clsA[kSubNodes - 1u - atomicAdd(&sCls[2], 1u)] = uint16_t(ni);
        }
        __syncthreads();
        const uint32_t nTiny = sCls[0], nHalf = sCls[1], nWarp = sCls[2], nSmallN = sCls[3], nCta = sCls[4];
        // the biggest nodes (top of the subtree) by the whole CTA, one after the other
        for (uint32_t k = 0; k < nCta; k++)
            build_cta_node(curList[sCtaIdx[k]], task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats, wBins,
                           sSplit, sWarpFirst);
        // larger nodes first (they are the long poles of the level): one warp each
        for (uint32_t k = warp; k < nWarp; k += kSubWarps)
            build_group_node<32>(curList[clsA[kSubNodes - 1u - k]], task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats, binsW);
        // 9..16 refs: one half warp each
        for (uint32_t k = warp * 2u + half; k < nHalf; k += kSubWarps * 2u)
            build_group_node<16>(curList[clsB[k]], task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats, binsH);
        __syncwarp();
        // 5..8 refs and 2..4 refs: one lane each, in packs of 32 handed out from the last warp backwards so that they
        // land on the warps the classes above loaded least
        for (uint32_t k = (kSubBlock - 1u - tid); k < nSmallN; k += kSubBlock)
            build_tiny_node<kSmallMax>(curList[clsB[kSubNodes - 1u - k]], task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats);
        __syncwarp();
        for (uint32_t k = (kSubBlock - 1u - tid); k < nTiny; k += kSubBlock)
            build_tiny_node<kTinyMax>(curList[clsA[k]], task, depth, budget, cLo, cHi, nLo, nHi, nodes, order, eon, em, sStats);
        __syncthreads();
        nCur = sNext;
        cur ^= 1u;
        level++;
        __syncthreads();
    }
    if (tid == 0) {
        if (sStats[0]) atomicAdd(&info->stats[3], sStats[0]);
        if (sStats[1]) atomicAdd(&info->stats[4], sStats[1]);
        if (sStats[2]) atomicMax(&info->stats[5], sStats[2]);
    }
  }
}

// Between two wide levels: the level's node count goes to the host (pinned flag: count + 1; 1 = the tree is finished), and
// the counters the level will fill for its successor are cleared.
__global__ void wide_prepare(const uint32_t* __restrict__ curCount, uint32_t* __restrict__ nextCount, volatile uint32_t* hostFlag) {
    chain_begin();
    uint32_t total = 0;
    for (int c = 0; c < kWideClasses; c++) total += curCount[c];
    for (int c = 0; c < 8; c++) nextCount[c] = 0u;   // five counts + the three work tickets of the group classes
    *hostFlag = total + 1u;
    __threadfence_system();
}

#ifndef ATLAS_WIDE_GROUP_CTAS
#define ATLAS_WIDE_GROUP_CTAS 3
#define ATLAS_WIDE_TINY_CTAS 2
#endif
constexpr int kWideGroupCtas = ATLAS_WIDE_GROUP_CTAS, kWideTinyCtas = ATLAS_WIDE_TINY_CTAS;
constexpr size_t kWideSmem = size_t(kSubWarps) * 2 * kSubtreeBins * kSubBinWords * sizeof(int);

// One level of the whole tree below the big levels. Every entry of the level's lists is built by the unit its size class
// names — the same node builders as the shared-memory kernel, reading and writing the global ping-pong ref arrays — and its
// children go to the next level's lists.
template <int CLASSES, int CTAS>   // CLASSES: bit c set = this instantiation builds the entries of size class c
__global__ void __launch_bounds__(kSubBlock, CTAS)
wide_level(WideLists cur, WideLists next, float4* const lo0, float4* const hi0, float4* const lo1, float4* const hi1, float4* nodes,
           uint32_t* __restrict__ order, uint8_t* __restrict__ eon, LevelInfo* __restrict__ info, uint32_t budget) {
    chain_begin();
    extern __shared__ __align__(16) unsigned char smemRaw[];
    int* wBins = reinterpret_cast<int*>(smemRaw);   // [warps][2 half-warp groups][32][kSubBinWords]
    __shared__ CtaSplit sSplit[3];
    __shared__ uint32_t sWarpFirst[kSubWarps];
    __shared__ unsigned long long sStats[3];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, half = lane >> 4;
    __shared__ ChunkState sChunk[kSubWarps * 2];   // [warp][half]: the CTA builder pushes through [0], a warp through [warp][0], a half warp through its own
    if (tid < 3) sStats[tid] = 0;
    if (tid < kSubWarps * 2)
        for (int c = 0; c < kWideClasses; c++) { sChunk[tid].base[c] = 0xffffffffu; sChunk[tid].used[c] = kWideChunk; }
    __syncthreads();
    SmallTask zero;   // the builders add the node's position to its task's: absolute positions, so a task at the origin
    zero.start = 0; zero.flatIdx = 0; zero.count = 0; zero.depth = 0; zero.buf = 0; zero.pad = 0;
    int* binsW = wBins + (warp * 2) * kSubtreeBins * kSubBinWords;
    int* binsH = binsW + half * kSubtreeBins * kSubBinWords;
    auto node_of = [](const SmallTask& t) {
        WideNode nd;
        nd.start = t.start; nd.count = t.count; nd.rel = t.flatIdx;
#pragma unroll
        for (int k = 0; k < 3; k++) { nd.lo[k] = t.lo[k]; nd.hi[k] = t.hi[k]; }
        return nd;
    };
#define ATLAS_WIDE_ARGS(t, EM) node_of(t), zero, (t).depth, budget, (t).buf ? lo1 : lo0, (t).buf ? hi1 : hi0, (t).buf ? lo0 : lo1, (t).buf ? hi0 : hi1, nodes, order, eon, \
                               EM, sStats
#define ATLAS_WIDE_EMIT(t) WideEmit{next, info, budget, (t).depth + 1u, (t).buf ^ 1u}
    // the biggest nodes: the whole CTA (uniform trip count: the builder contains barriers)
    const uint32_t nCta = (CLASSES & 16) ? min(cur.count[4], cur.cap[4]) : 0u;
    // (entries are drawn from tickets, words 5..7 of the level's counters: the lists of the group classes have holes in a
    // pattern that a fixed stride would hand to the same units every time, and the nodes of a class differ in size)
    __shared__ uint32_t sTicket;
    while (nCta) {
        __syncthreads();
        if (tid == 0) sTicket = atomicAdd(&cur.count[7], 1u);
        __syncthreads();
        const uint32_t k = sTicket;
        if (k >= nCta) break;
        const SmallTask t = cur.list[4][k];
        if (t.count < 2u) continue;   // a hole (uniform for the CTA)
        build_cta_node(ATLAS_WIDE_ARGS(t, (WideEmitChunk{ATLAS_WIDE_EMIT(t), &sChunk[0]})), wBins, sSplit, sWarpFirst);
    }
    // one warp each, handed out from the far end of the grid so that they land on CTAs the class above did not load
    const uint32_t nWarp = (CLASSES & 8) ? min(cur.count[3], cur.cap[3]) : 0u;
    // (several entries per ticket only when every unit gets plenty of them: the first levels have a few big nodes)
    const uint32_t stepW = nWarp >= gridDim.x * kSubWarps * 8u ? kWideChunk : 1u;
    while (nWarp) {
        uint32_t first = 0;
        if (lane == 0) first = atomicAdd(&cur.count[6], stepW);
        first = __shfl_sync(kFullMask, first, 0);
        if (first >= nWarp) break;
        for (uint32_t k = first; k < min(first + stepW, nWarp); k++) {
            const SmallTask t = cur.list[3][k];
            if (t.count < 2u) continue;
            build_group_node<32>(ATLAS_WIDE_ARGS(t, (WideEmitChunk{ATLAS_WIDE_EMIT(t), &sChunk[warp * 2]})), binsW);
        }
    }
    const uint32_t nHalf = (CLASSES & 4) ? min(cur.count[2], cur.cap[2]) : 0u;
    {
        // the two half warps of a warp draw ONE ticket and take alternate entries of it, so that they run the builder in the same
        // instruction stream (it is written for that); an entry that is a hole idles its half for one iteration only
        const uint32_t stepH = nHalf >= gridDim.x * kSubWarps * 16u ? kWideChunk : 1u;
        while (nHalf) {
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(&cur.count[5], 2u * stepH);
            first = __shfl_sync(kFullMask, first, 0);
            if (first >= nHalf) break;
            for (uint32_t j = 0; j < stepH; j++) {
                const uint32_t k = first + 2u * j + half;
                if (k < nHalf) {
                    const SmallTask t = cur.list[2][k];
                    if (t.count >= 2u)
                        build_group_node<16>(ATLAS_WIDE_ARGS(t, (WideEmitChunk{ATLAS_WIDE_EMIT(t), &sChunk[warp * 2 + half]})), binsH);
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
    const uint32_t nSmallN = (CLASSES & 2) ? min(cur.count[1], cur.cap[1]) : 0u;
    for (uint32_t k = blockIdx.x * kSubBlock + tid; k < nSmallN; k += gridDim.x * kSubBlock) {
        const SmallTask t = cur.list[1][k];
        if (t.count < 2u) continue;
        build_tiny_node<kSmallMax>(ATLAS_WIDE_ARGS(t, (WideEmitAgg{ATLAS_WIDE_EMIT(t)})));
    }
    __syncwarp();
    const uint32_t nTiny = (CLASSES & 1) ? min(cur.count[0], cur.cap[0]) : 0u;
    for (uint32_t k = blockIdx.x * kSubBlock + tid; k < nTiny; k += gridDim.x * kSubBlock) {
        const SmallTask t = cur.list[0][k];
        if (t.count < 2u) continue;
        build_tiny_node<kTinyMax>(ATLAS_WIDE_ARGS(t, (WideEmitAgg{ATLAS_WIDE_EMIT(t)})));
    }
#undef ATLAS_WIDE_ARGS
#undef ATLAS_WIDE_EMIT
    __syncthreads();
    // slots reserved but not used become holes
    if ((CLASSES & 28) && (tid & 15u) == 0u) chunk_flush(next, &sChunk[tid >> 4]);
    if (tid == 0) {
        if (sStats[0]) atomicAdd(&info->stats[3], sStats[0]);
        if (sStats[1]) atomicAdd(&info->stats[4], sStats[1]);
        if (sStats[2]) atomicMax(&info->stats[5], sStats[2]);
    }
}

template <typename T>
int read_back(atlas_rt_context* ctx, const T* dev, T* host) {
    static_assert(sizeof(T) <= 4096, "pinned staging too small");
    ATLAS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host, ctx->pinned, sizeof(T));
    return ATLAS_RT_OK;
}

}   // namespace

#define ATLAS_TRY(expr)            \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != ATLAS_RT_OK) { cleanup(); return rc__; } \
    } while (0)
#define ATLAS_CUDA_C(ctx, call)                                                          \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) { cleanup(); return fail((ctx), ATLAS_RT_ERR_CUDA, #call, e__); } \
    } while (0)
#define ATLAS_LAUNCHED(ctx)                                                              \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) { cleanup(); return fail((ctx), ATLAS_RT_ERR_CUDA, "kernel launch", e__); } \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of the function, so it is set for every context at
// creation (after cudaSetDevice), not once per process: a second context on another GPU would otherwise launch
// build_subtrees with > 48 KB of dynamic shared memory it never opted in to.
int build_init_device(atlas_rt_context* ctx) {
    ATLAS_CUDA(ctx, cudaFuncSetAttribute(build_subtrees, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSubtreeSmem)));
    return ATLAS_RT_OK;
}

int build_bvh(atlas_rt_context* ctx, const float* dAabbs, const float* dTris, uint64_t count64, bool tlas, atlas_rt_bvh* out) {
    const uint32_t n = uint32_t(count64);
    cudaStream_t st = ctx->stream;
    out->nodeCount = 0;
    out->refCount = 0;
    if (n == 0) return ATLAS_RT_OK;   // the reference dereferences a null child for an empty input; we return an empty BVH

    if (tlas && n == 1) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &out->nodes, 4));
        ATLAS_CUDA(ctx, dev_alloc(ctx, &out->order, 2));
        ATLAS_CUDA(ctx, dev_alloc(ctx, &out->endOfNode, 2));
        tlas_single<<<1, 1, 0, st>>>(dAabbs, out->nodes, out->order, out->endOfNode);
        ATLAS_LAUNCH_CHECK(ctx);
        out->nodeCount = 1;
        out->refCount = 2;
        return ATLAS_RT_OK;
    }

    BuildBuffers B;
    memset(&B, 0, sizeof(B));
    B.budget = tlas ? 64u : 256u;
    B.cap = tlas ? n : 2u * n;
    B.tris = dTris;
    const uint32_t cap = B.cap;
    const uint32_t maxTasks = cap / kSubtreeMax + 256u + 2u;
    const uint32_t maxSmall = cap / 2u + 2u;
    const uint32_t maxChunks = cap / kChunk + maxTasks + 2u;
    // bins scratch: the widest level is bounded by min(2^d, maxTasks) * bins(d)
    uint64_t binRecords = 0;
    for (uint32_t d = 0; d < 40; d++) {
        const uint64_t tasksAtDepth = d < 31 ? std::min<uint64_t>(1ull << d, maxTasks) : maxTasks;
        binRecords = std::max<uint64_t>(binRecords, tasksAtDepth * bins_at_depth(B.budget, d));
    }
    float4 *tmpLlo = nullptr, *tmpLhi = nullptr, *tmpRlo = nullptr, *tmpRhi = nullptr, *strad = nullptr;
    uint32_t* spaCounts = nullptr;
    // wide levels (ATLAS_RT_BUILD_WIDE): two sets of per-class node lists + their counters
    const bool wide = ctx->buildWide != 0;
    WideLists W[2];
    memset(W, 0, sizeof(W));
    uint32_t* wideCounts = nullptr;
    auto cleanup = [&]() {
        for (int k = 0; k < 2; k++) for (int c = 0; c < kWideClasses; c++) dev_free(ctx, W[k].list[c]);
        dev_free(ctx, wideCounts);
        for (int k = 0; k < 2; k++) { dev_free(ctx, B.lo[k]); dev_free(ctx, B.hi[k]); dev_free(ctx, B.tasks[k]); }
        dev_free(ctx, B.small); dev_free(ctx, B.info); dev_free(ctx, B.root); dev_free(ctx, B.bins);
        dev_free(ctx, B.spaBins); dev_free(ctx, B.medAcc); dev_free(ctx, B.chunkBase); dev_free(ctx, B.chunkFirst); dev_free(ctx, B.chunkInfo);
        dev_free(ctx, B.rootBox); dev_free(ctx, tmpLlo); dev_free(ctx, tmpLhi); dev_free(ctx, tmpRlo); dev_free(ctx, tmpRhi);
        dev_free(ctx, strad); dev_free(ctx, spaCounts);
    };
    for (int k = 0; k < 2; k++) {
        ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.lo[k], cap));
        ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.hi[k], cap));
        ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.tasks[k], maxTasks));
    }
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &out->nodes, size_t(cap) * 4));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &out->order, cap));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &out->endOfNode, cap));
    B.nodes = out->nodes; B.order = out->order; B.eon = out->endOfNode;
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.small, maxSmall));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.info, 1));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.root, 1));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.bins, binRecords * 3 * kBinWords));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.spaBins, size_t(3) * 256 * kBinWords));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.medAcc, size_t(maxTasks) * 16));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.chunkBase, maxTasks));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.chunkFirst, maxChunks));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.chunkInfo, maxChunks));
    ATLAS_CUDA_C(ctx, dev_alloc(ctx, &B.rootBox, 8));
    if (wide) {
        const uint32_t div[kWideClasses] = {2u, 5u, 9u, 9u, 257u};   // fewest refs a node of the class holds
        ATLAS_CUDA_C(ctx, dev_alloc(ctx, &wideCounts, 16));
        ATLAS_CUDA_C(ctx, cudaMemsetAsync(wideCounts, 0, 16 * sizeof(uint32_t), st));
        for (int k = 0; k < 2; k++) {
            W[k].count = wideCounts + 8 * k;
            for (int c = 0; c < kWideClasses; c++) {
                W[k].cap[c] = cap / div[c] + 2u + uint32_t(ctx->smCount) * 4u * uint32_t(kSubWarps) * 2u * kWideChunk;   // + the holes a level can leave
                ATLAS_CUDA_C(ctx, dev_alloc(ctx, &W[k].list[c], W[k].cap[c]));
            }
        }
    }
    ATLAS_CUDA_C(ctx, cudaMemsetAsync(B.info, 0, sizeof(LevelInfo), st));
    ATLAS_CUDA_C(ctx, cudaMemsetAsync(B.root, 0, sizeof(RootSplit), st));
    ATLAS_CUDA_C(ctx, cudaMemsetAsync(out->endOfNode, 0, cap, st));
    {
        const int initBox[8] = {kOrdEmptyLo, kOrdEmptyLo, kOrdEmptyLo, kOrdEmptyHi, kOrdEmptyHi, kOrdEmptyHi, 0, 0};
        memcpy(ctx->pinned, initBox, sizeof(initBox));
        ATLAS_CUDA_C(ctx, cudaMemcpyAsync(B.rootBox, ctx->pinned, sizeof(initBox), cudaMemcpyHostToDevice, st));
    }
    const bool pdl = ctx->chainLaunch != 0;
    ATLAS_CUDA_C(ctx, launch_chain(pdl, init_refs, std::max(1u, std::min<uint32_t>((n + 255) / 256, uint32_t(ctx->smCount) * 8u)), 256, 0, st, dAabbs, n,
                                   B.lo[0], B.hi[0], B.rootBox, B.info));
    ctx->launches++;
    ATLAS_CUDA_C(ctx, launch_chain(pdl, make_root, 1, 1, 0, st, B.rootBox, n, B.tasks[0], B.info, B.root));
    ctx->launches++;

    const uint32_t persistent = uint32_t(ctx->smCount) * 4u;
    uint32_t cur = 0;
    LevelInfo info;
    memset(&info, 0, sizeof(info));
    bool rootLeaf = false;
    bool binsReady = false;
    uint32_t totalRefs = n;
    // The level loop is enqueued WITHOUT waiting for the device: grids are sized from upper bounds, the kernels read the
    // real task / chunk counts from device memory and fall through when a level is empty. After each level the 128-byte
    // level record is copied to a pinned slot; the host only looks at records that are kLookahead levels old (or already
    // complete) to learn that the tree is finished, so host latency never stalls the GPU. Only the BLAS root, whose
    // outcome decides which kernels run and what must be allocated, is read back synchronously.
    constexpr uint32_t kLookahead = 3;
    constexpr uint32_t kFlagRing = 64;
    volatile uint32_t* flags = static_cast<volatile uint32_t*>(ctx->levelSlots);   // [kFlagRing], pinned: level d's node count + 1
    for (uint32_t k = 0; k < kFlagRing; k++) flags[k] = 0u;
    auto wait_flag = [&](uint32_t level) -> int {   // spin until prepare_level(level) has run; watch for a dead stream
        for (uint64_t spins = 0; flags[level % kFlagRing] == 0u; spins++) {
            if ((spins & 0xfffu) == 0xfffu) {
                const cudaError_t q = cudaStreamQuery(st);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail(ctx, ATLAS_RT_ERR_CUDA, "level loop", q);
                if (q == cudaSuccess && flags[level % kFlagRing] == 0u) return fail(ctx, ATLAS_RT_ERR_CUDA, "level flag never written");
            }
        }
        return ATLAS_RT_OK;
    };

    for (uint32_t depth = 0; depth < 100000u; depth++) {
        if (depth > 0) {
            // the host runs at most kLookahead levels ahead of the device and stops at the first level found empty
            if (depth > kLookahead) {
                const int rc = wait_flag(depth - kLookahead);
                if (rc != ATLAS_RT_OK) { cleanup(); return rc; }
            }
            bool finished = false;
            for (uint32_t k = depth > kLookahead + 1u ? depth - kLookahead - 1u : 1u; k < depth && !finished; k++)
                if (flags[k % kFlagRing] == 1u) finished = true;
            if (finished) break;
            flags[depth % kFlagRing] = 0u;   // (level depth - kFlagRing was consumed long ago)
        }
        const uint32_t nb = bins_at_depth(B.budget, depth);
        const uint32_t tasksBound = depth == 0 ? 1u : std::min<uint32_t>(depth < 31u ? (1u << depth) : maxTasks, maxTasks);
        const uint32_t chunksBound = std::min<uint32_t>(totalRefs / kChunk + tasksBound + 1u, maxChunks);
        const uint32_t gridChunks = std::max(1u, std::min(chunksBound, persistent));
        const uint32_t warpGrid = (tasksBound * 32u + 127u) / 128u;
        const size_t binSmem = size_t(3) * nb * kSmemBin * sizeof(int);
        Task* tasks = B.tasks[cur];
        Lists L{W[0], wide ? 1u : 0u, B.tasks[cur ^ 1u], B.small, B.info, B.nodes, B.budget, cur ^ 1u, maxTasks, maxSmall};
        const float4 *rlo = B.lo[cur], *rhi = B.hi[cur];
        float4 *wlo = B.lo[cur ^ 1u], *whi = B.hi[cur ^ 1u];

        ATLAS_CUDA_C(ctx, launch_chain(pdl, prepare_level, 1, 1024, 0, st, tasks, B.info, B.chunkBase, B.chunkInfo, B.chunkFirst, depth > 0 ? 1 : 0,
                                       flags + depth % kFlagRing));
        ctx->launches++;
        if (!binsReady) {   // otherwise the previous level's partition_scatter has reset this level's bins
            ATLAS_CUDA_C(ctx, launch_chain(pdl, init_bins, std::max(1u, std::min<uint32_t>(persistent, (tasksBound * 3u * nb + 255u) / 256u)), 256, 0, st,
                                           B.bins, B.info, 3u * nb));
            ctx->launches++;
        }
        binsReady = false;
        const uint32_t gridBin = std::max(1u, std::min(chunksBound, uint32_t(ctx->smCount) * uint32_t(ctx->binCtasPerSM)));
        ATLAS_CUDA_C(ctx, launch_chain(pdl, bin_big, gridBin, kBigBlock, binSmem, st, tasks, B.info, B.chunkInfo, rlo, rhi, B.bins, nb));
        ctx->launches++;

        bool spatialPath = false;
        if (depth == 0 && !tlas) {
            ATLAS_CUDA_C(ctx, launch_chain(pdl, select_root_object, 1, kSelectBlock, select_smem(nb), st, tasks, B.info, B.bins, B.root, nb));
            ctx->launches++;
            // the spatial binning is enqueued unconditionally and falls through on the device when the object split's
            // children do not overlap enough (no host round trip for that decision)
            ATLAS_CUDA_C(ctx, launch_chain(pdl, init_bins, 3, 256, 0, st, B.spaBins, B.info, 3u * nb));   // nTasks == 1
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, spatial_bin_root, std::max(1u, std::min(chunksBound, uint32_t(ctx->smCount) * 3u)), kBigBlock, binSmem, st,
                                           tasks, B.info, rlo, rhi, B.tris, B.spaBins, nb));
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, select_root_final, 1, kSelectBlock, select_smem(nb), st, tasks, B.info, B.bins, B.spaBins, B.medAcc, B.root,
                                           L, nb));
            ctx->launches++;
            ATLAS_TRY(read_back(ctx, B.info, &info));
            out->stats[0] = info.rootNeedSpatial ? 1 : 0;
            spatialPath = info.rootKind == uint32_t(kSpatial);
        } else {
            ATLAS_CUDA_C(ctx, launch_chain(pdl, select_big, tasksBound, kSelectBlock, select_smem(nb), st, tasks, B.info, B.bins, B.medAcc, L, nb));
            ctx->launches++;
        }

        if (spatialPath) {
            const uint32_t nChunks = (n + kChunk - 1) / kChunk;
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &tmpLlo, n));
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &tmpLhi, n));
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &tmpRlo, n));
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &tmpRhi, n));
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &strad, size_t(6) * n));
            ATLAS_CUDA_C(ctx, dev_alloc(ctx, &spaCounts, size_t(3) * nChunks));
            spatial_count<<<gridChunks, kBigBlock, 0, st>>>(tasks, rlo, rhi, B.root, spaCounts, nChunks, nb);
            ATLAS_LAUNCHED(ctx);
            spatial_scan<<<1, 96, 0, st>>>(spaCounts, nChunks, B.info);
            ATLAS_LAUNCHED(ctx);
            spatial_scatter<<<gridChunks, kBigBlock, 0, st>>>(tasks, rlo, rhi, B.tris, B.root, spaCounts, nChunks, nb, tmpLlo, tmpLhi,
                                                              tmpRlo, tmpRhi, strad, n);
            ATLAS_LAUNCHED(ctx);
            spatial_sequential<<<1, kSeqTile, 0, st>>>(B.info, B.root, strad, n, tmpLlo, tmpLhi, tmpRlo, tmpRhi);
            ATLAS_LAUNCHED(ctx);
            spatial_emit<<<1, 1, 0, st>>>(tasks, B.info, B.root, L);
            ATLAS_LAUNCHED(ctx);
            spatial_place<<<gridChunks, 256, 0, st>>>(tasks, B.root, tmpLlo, tmpLhi, tmpRlo, tmpRhi, wlo, whi, B.order, B.eon);
            ATLAS_LAUNCHED(ctx);
        } else {
            ATLAS_CUDA_C(ctx, launch_chain(pdl, median_reduce_big, gridChunks, kBigBlock, 0, st, tasks, B.info, B.chunkInfo, rlo, rhi, B.medAcc));
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, median_finalize_big, warpGrid, 128, 0, st, tasks, B.info, B.medAcc, B.lo[cur], B.hi[cur], wlo, whi,
                                           B.order, B.eon, L, (!tlas && depth == 0) ? 1 : 0));
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, partition_count, gridChunks, kBigBlock, 0, st, tasks, B.info, B.chunkInfo, rlo, rhi, B.chunkFirst, nb));
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, partition_scatter, gridChunks, kBigBlock, 0, st, tasks, B.info, B.chunkInfo, B.chunkFirst, rlo, rhi, wlo,
                                           whi, B.order, B.eon, nb, B.bins, 3u * bins_at_depth(B.budget, depth + 1u),
                                           std::min<uint32_t>(depth < 30u ? (2u << depth) : maxTasks, maxTasks)));
            binsReady = true;
            ctx->launches++;
        }
        if (depth == 0 && !tlas && info.rootKind != uint32_t(kObject)) {
            // a spatial root adds duplicates and a median root may have stayed a leaf: both size what follows
            ATLAS_TRY(read_back(ctx, B.info, &info));
            totalRefs = info.totalRefs;
            if (info.rootLeaf) { rootLeaf = true; break; }
        }
        cur ^= 1u;
    }
    // the subtrees are enqueued straight behind the last level (the kernel reads their number on the device); the one
    // read-back that follows waits for the whole build
    if (!rootLeaf && wide) {
        // the rest of the tree level by level: list set 0 holds what the big levels handed down; the host stays kLookahead
        // levels ahead of the device and stops at the first level it learns was empty
        constexpr uint32_t kWideRing = 128;
        volatile uint32_t* wflags = flags + kFlagRing;
        for (uint32_t k = 0; k < kWideRing; k++) wflags[k] = 0u;
        const uint32_t wgrid = uint32_t(ctx->smCount) * uint32_t(kSubCtasPerSM);
        int cw = 0;
        for (uint32_t lvl = 0; lvl < 100000u; lvl++) {
            if (lvl >= kLookahead) {
                const uint32_t seen = lvl - kLookahead;
                for (uint64_t spins = 0; wflags[seen % kWideRing] == 0u; spins++) {
                    if ((spins & 0xfffu) == 0xfffu) {
                        const cudaError_t q = cudaStreamQuery(st);
                        if (q != cudaSuccess && q != cudaErrorNotReady) { cleanup(); return fail(ctx, ATLAS_RT_ERR_CUDA, "wide level loop", q); }
                        if (q == cudaSuccess && wflags[seen % kWideRing] == 0u) { cleanup(); return fail(ctx, ATLAS_RT_ERR_CUDA, "wide level flag never written"); }
                    }
                }
                if (wflags[seen % kWideRing] == 1u) break;   // that level had no node: the tree was complete before it
                if (lvl >= kWideRing) wflags[lvl % kWideRing] = 0u;   // (consumed kWideRing - kLookahead levels ago)
            }
            ATLAS_CUDA_C(ctx, launch_chain(pdl, wide_prepare, 1, 1, 0, st, static_cast<const uint32_t*>(W[cw].count), W[cw ^ 1].count, wflags + lvl % kWideRing));
            ctx->launches++;
            ATLAS_CUDA_C(ctx, launch_chain(pdl, wide_level<16, 2>, wgrid, kSubBlock, kWideSmem, st, W[cw], W[cw ^ 1], B.lo[0], B.hi[0], B.lo[1], B.hi[1], B.nodes,
                                           B.order, B.eon, B.info, B.budget));
            ATLAS_CUDA_C(ctx, launch_chain(pdl, wide_level<12, kWideGroupCtas>, uint32_t(ctx->smCount) * kWideGroupCtas, kSubBlock, kWideSmem, st, W[cw], W[cw ^ 1], B.lo[0],
                                           B.hi[0], B.lo[1], B.hi[1], B.nodes, B.order, B.eon, B.info, B.budget));
            ATLAS_CUDA_C(ctx, launch_chain(pdl, wide_level<3, kWideTinyCtas>, uint32_t(ctx->smCount) * kWideTinyCtas, kSubBlock, kWideSmem, st, W[cw], W[cw ^ 1], B.lo[0],
                                           B.hi[0], B.lo[1], B.hi[1], B.nodes, B.order, B.eon, B.info, B.budget));
            ctx->launches += 3;
            cw ^= 1;
        }
    } else if (!rootLeaf) {
        ATLAS_CUDA_C(ctx, launch_chain(pdl, build_subtrees, uint32_t(ctx->smCount) * uint32_t(kSubCtasPerSM), kSubBlock, kSubtreeSmem, st, B.small, B.lo[0], B.hi[0], B.lo[1],
                                       B.hi[1], B.nodes, B.order, B.eon, B.info, B.budget, maxSmall));
        ctx->launches++;
    }
    ATLAS_TRY(read_back(ctx, B.info, &info));
    if (info.overflow) { cleanup(); return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "task list overflow"); }
    totalRefs = info.totalRefs;
    const uint32_t levels = info.levels;

    if (rootLeaf) {
        root_leaf_output<<<(n + 255) / 256, 256, 0, st>>>(n, B.order, B.eon);
        ATLAS_LAUNCHED(ctx);
        out->nodeCount = 0;
        out->refCount = n;
    } else {
        out->refCount = totalRefs;
        out->nodeCount = totalRefs - 1u;
    }
    out->stats[1] = info.stats[1];
    out->stats[2] = info.stats[2];
    out->stats[3] = info.stats[3];
    out->stats[4] = info.stats[4];
    out->stats[5] = info.stats[5];
    out->stats[6] = levels;
    out->stats[7] = info.negZero;
    cleanup();
    return ATLAS_RT_OK;
}

}   // namespace atlas
