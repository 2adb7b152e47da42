// trace.cu — closest-hit / any-hit traversal of the two-level (TLAS -> BLAS) scene in the engine's flattened layout.
//
// Restates, per ray and in the same visit order, data/shader/raytracer/bvh.hsh:191-273 (HitClosest) and :359-441
// (HitAny) with CheckInstance (:172-189), CheckLeafClosest / CheckLeaf (:44-104), UnpackNode (:21-37),
// IntersectAABB / IntersectTriangle (intersections.hsh:19-58) and the batch wrapper traceClosest.csh:12-36.
// Visit order is part of the contract: ties in t are resolved by "first visited wins" (strict < in bvh.hsh:62-63), so
// one ray's node/triangle sequence must never be reordered; parallelism is across rays only.
//
// Data movement: a node is 64 B = 2 x LDG.256 through the read-only path (sm_100's 256-bit loads halve the L1TEX
// wavefronts per visit), an instance 64 B = 2 x LDG.256, a triangle 48 B = 3 x LDG.128 (96 B = 3 x LDG.256 for the
// opacity-aware variants); rays are read and written once as 3 x 128-bit each. The per-ray stack (32 node pointers, the
// reference's STACK_SIZE) lives in shared memory laid out [entry][lane] so a warp's accesses never bank-conflict.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace atlas {

namespace {

constexpr int kTraceBlock = 128;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kStack = ATLAS_RT_STACK_SIZE;
constexpr uint32_t kTlasInvalid = kStack + 2;   // TLAS_INVALID, bvh.hsh:17

struct SceneDev {
    const float4* tlasNodes;
    const float4* instances;
    const float4* const* blasNodes;
    const float4* const* bvhTris;
    const float4* const* triangles;   // 96-byte GPUTriangle arrays (opacity-aware variants)
    const uint32_t* materials;        // RaytraceMaterial records, 23 words each (null: textured opacity counts as 1)
    const TextureDev* textures;       // R8 opacity textures
    uint32_t materialCount, textureCount;
};

// GetOpacity, data/shader/raytracer/surface.hsh:147-160: texture coordinates interpolated from the triangle's half2 words
// (d0.xyz), flipped for invertUVs, the material's opacity texture sampled bilinearly at mip 0, times the material opacity.
__device__ __forceinline__ float get_opacity(const SceneDev& sc, const float4 d0, float s, float t, int materialOffset) {
    if (!sc.materials) return 1.0f;
    const uint32_t mi = uint32_t(__float_as_int(d0.w) + materialOffset);
    if (mi >= sc.materialCount) return 1.0f;
    const uint32_t* M = sc.materials + 23 * size_t(mi);
    const float r = __fsub_rn(__fsub_rn(1.0f, s), t);
    const uint32_t w0 = __float_as_uint(d0.x), w1 = __float_as_uint(d0.y), w2 = __float_as_uint(d0.z);
    const float u0 = __half2float(__ushort_as_half(uint16_t(w0))), v0 = __half2float(__ushort_as_half(uint16_t(w0 >> 16)));
    const float u1 = __half2float(__ushort_as_half(uint16_t(w1))), v1 = __half2float(__ushort_as_half(uint16_t(w1 >> 16)));
    const float u2 = __half2float(__ushort_as_half(uint16_t(w2))), v2 = __half2float(__ushort_as_half(uint16_t(w2 >> 16)));
    const float u = __fadd_rn(__fadd_rn(__fmul_rn(r, u0), __fmul_rn(s, u1)), __fmul_rn(t, u2));
    float v = __fadd_rn(__fadd_rn(__fmul_rn(r, v0), __fmul_rn(s, v1)), __fmul_rn(t, v2));
    if (int(M[13]) > 0) v = __fsub_rn(1.0f, v);                      // invertUVs
    const int tex = int(M[18]);                                      // opacityTexture
    const float texel = (tex < 0 || uint32_t(tex) >= sc.textureCount) ? 1.0f : sample_r8(sc.textures[tex], u, v);
    return __fmul_rn(texel, __uint_as_float(M[7]));                  // * rayMat.opacity
}

// 256-bit read-only load (sm_100: LDG.E.256). A 64-byte node or instance record is two of these instead of four
// 128-bit loads, which halves the L1TEX wavefronts per visit — the unit the traversal saturates (profiles/: l1tex
// throughput 90 % with 128-bit loads). p must be 32-byte aligned (nodes and instances are 64-byte records in
// cudaMalloc'ed arrays).
struct alignas(32) Float8 { float4 a, b; };
__device__ __forceinline__ Float8 ldg256(const float4* p) {
    Float8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}

// IntersectAABB (intersections.hsh:19-34): true division by the direction, GLSL min/max forms.
__device__ __forceinline__ bool slab(const float o[3], const float d[3], const float lo[3], const float hi[3],
                                     float tmin, float tmax, float& dist) {
    float ts[3], tb[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float t0 = __fdiv_rn(__fsub_rn(lo[a], o[a]), d[a]);
        const float t1 = __fdiv_rn(__fsub_rn(hi[a], o[a]), d[a]);
        ts[a] = gl_min(t0, t1);
        tb[a] = gl_max(t0, t1);
    }
    const float tminf = gl_max(gl_max(tmin, ts[0]), gl_max(ts[1], ts[2]));
    const float tmaxf = gl_min(gl_min(tmax, tb[0]), gl_min(tb[1], tb[2]));
    const bool hit = tminf <= tmaxf;
    dist = hit ? tminf : tmax;
    return hit;
}

__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// IntersectTriangle (intersections.hsh:36-58).
__device__ __forceinline__ bool tri_test(const float o[3], const float d[3], const float4 a, const float4 b,
                                         const float4 c, float sol[3]) {
    const float e0x = __fsub_rn(b.x, a.x), e0y = __fsub_rn(b.y, a.y), e0z = __fsub_rn(b.z, a.z);
    const float e1x = __fsub_rn(c.x, a.x), e1y = __fsub_rn(c.y, a.y), e1z = __fsub_rn(c.z, a.z);
    const float sx = __fsub_rn(o[0], a.x), sy = __fsub_rn(o[1], a.y), sz = __fsub_rn(o[2], a.z);
    // cross(s, e0), cross(d, e1): (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
    const float px = __fsub_rn(__fmul_rn(sy, e0z), __fmul_rn(e0y, sz));
    const float py = __fsub_rn(__fmul_rn(sz, e0x), __fmul_rn(e0z, sx));
    const float pz = __fsub_rn(__fmul_rn(sx, e0y), __fmul_rn(e0x, sy));
    const float qx = __fsub_rn(__fmul_rn(d[1], e1z), __fmul_rn(e1y, d[2]));
    const float qy = __fsub_rn(__fmul_rn(d[2], e1x), __fmul_rn(e1z, d[0]));
    const float qz = __fsub_rn(__fmul_rn(d[0], e1y), __fmul_rn(e1x, d[1]));
    const float den = dot3(qx, qy, qz, e0x, e0y, e0z);
    sol[0] = __fdiv_rn(dot3(px, py, pz, e1x, e1y, e1z), den);
    sol[1] = __fdiv_rn(dot3(qx, qy, qz, sx, sy, sz), den);
    sol[2] = __fdiv_rn(dot3(px, py, pz, d[0], d[1], d[2]), den);
    return sol[0] >= 0.0f && sol[1] >= 0.0f && sol[2] >= 0.0f && __fadd_rn(sol[1], sol[2]) <= 1.0f;
}

// ---------------------------------------------------------------------------------------------------------------
// Exact division by a per-ray precomputed reciprocal. rc = __frcp_rn(d) is the correctly rounded reciprocal; q0 =
// RN(x * rc) is then a faithful quotient and one FMA residual + one FMA correction (Markstein's theorem) give the
// correctly rounded x / d — three fma-pipe instructions and nothing on the XU (MUFU) pipe, which the 12 IEEE divisions
// per node otherwise saturate (profiles/trace_r1.md: XU pipe 58 % busy in the first version).
// Valid while nothing over- or UNDERflows — the quotient, and the residual e, which is about 2^-24 of x, must stay normal:
// the caller only takes this path (`fast` flag) for rays whose direction components lie in [2^-64, 2^30], whose origin
// components are below 2^60 and either at least 2^-40 in magnitude or exactly zero in a scene whose boxes hold no
// non-zero coordinate below 2^-60 (scene flag bit 1), in scenes with all coordinates below 2^60 (bit 0). Then every
// numerator x = box - origin is 0 (quotient exactly 0 either way) or at least 2^-64, the quotient at least 2^-94 and the
// residual at least 2^-118. All other rays use __fdiv_rn. tools/divcheck.cu compares this sequence with __fdiv_rn over
// the admitted domain (uniform and adversarial mantissas, exponents over the whole range, zero numerators): 0 mismatches
// in 3e11 operand pairs on a B200 — and shows that mismatches DO occur once the quotient or the residual goes subnormal.
__device__ __forceinline__ float div_by_rcp(float x, float d, float rc) {
    const float q = __fmul_rn(x, rc);
    const float e = __fmaf_rn(-d, q, x);
    return __fmaf_rn(e, rc, q);
}

// IntersectAABB on the fast path. No operand can be NaN here (finite boxes, finite non-zero direction), so
// fminf/fmaxf (one FMNMX each) agree with the GLSL comparison forms up to the sign of zero, which no later comparison
// can observe.
__device__ __forceinline__ bool slab_fast(const float o[3], const float d[3], const float rc[3], const float lo[3],
                                          const float hi[3], float tmin, float tmax, float& dist) {
    float ts[3], tb[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float t0 = div_by_rcp(__fsub_rn(lo[a], o[a]), d[a], rc[a]);
        const float t1 = div_by_rcp(__fsub_rn(hi[a], o[a]), d[a], rc[a]);
        ts[a] = fminf(t0, t1);
        tb[a] = fmaxf(t0, t1);
    }
    const float tminf = fmaxf(fmaxf(tmin, ts[0]), fmaxf(ts[1], ts[2]));
    const float tmaxf = fminf(fminf(tmax, tb[0]), fminf(tb[1], tb[2]));
    const bool hit = tminf <= tmaxf;
    dist = hit ? tminf : tmax;
    return hit;
}

constexpr float kDirLo = 5.421010862427522e-20f;   // 2^-64
constexpr float kDirHi = 1073741824.0f;            // 2^30
constexpr float kPosHi = 1.152921504606847e18f;    // 2^60
constexpr float kPosLo = 9.094947017729282e-13f;   // 2^-40
constexpr float kCoordLo = 8.673617379884035e-19f; // 2^-60: smallest non-zero box coordinate for which origin components of exactly 0 stay fast

__device__ __forceinline__ bool fast_ok(const float o[3], const float d[3], int sceneFlags) {
    bool ok = (sceneFlags & 1) != 0;
    const bool zeroOk = (sceneFlags & 2) != 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float ad = fabsf(d[a]), ao = fabsf(o[a]);
        ok = ok && (ad >= kDirLo) && (ad <= kDirHi) && (ao <= kPosHi) && (ao >= kPosLo || (ao == 0.0f && zeroOk));
    }
    return ok;
}

constexpr int kBlocksPerSM = 9;

// Streaming input (host rays): the batch is still being uploaded while the kernel runs. `watermark` (device memory) is
// the number of rays that have arrived: the upload stream bumps it with a 4-byte copy behind every chunk of `chunkRays`
// rays. `chunkDone[c]` counts finished rays of chunk c so that the download stream (cuStreamWaitValue32) can send a chunk's
// results home as soon as its last ray is done. Null watermark = the whole batch is resident.
struct StreamIn {
    const unsigned int* watermark;
    unsigned int* chunkDone;
    uint32_t chunkRays;
    uint32_t spinBound;
    // (every variant) batch size known only on the device: CTAs beyond what that many rays need leave at once, so a small
    // late bounce of the path tracer does not fill the machine with idle persistent warps while another lane has work
    uint32_t raysPerBlock, minBlocks;
};
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Persistent-thread traversal. Each warp owns 32 ray slots and refills finished slots from a global ray counter, so
// lanes do not idle while the longest ray of a static batch finishes (the one-thread-per-ray kernel ran at 8.5 of 32
// active lanes). Within a warp every round is warp-uniform: either the lanes standing at an inner node take one
// traversal step, or — once enough lanes wait at a leaf / instance — those lanes process it. A ray's own visit order is
// exactly the reference's; only the interleaving between different rays changes.
// The opacity-aware variants carry the material offset, the triangle's uv words and the texture sampling: one CTA per SM
// fewer (64 registers instead of 56) keeps them out of local memory.
// STREAM (host rays still arriving, see StreamIn) is a template parameter so that the resident-batch instantiations carry
// none of its state: at the 56-register cap even a dead flag costs moves in the hot loop (measured: 4 % on C2).
template <bool ANY, bool COUNT, bool OPACITY, bool STREAM>
__global__ void __launch_bounds__(kTraceBlock, (OPACITY || STREAM) ? kBlocksPerSM - 1 : kBlocksPerSM)
trace_kernel(SceneDev sc, const float4* in, float4* out, const uint32_t* __restrict__ permIn, const unsigned int* __restrict__ usePerm,
             uint32_t count, const uint32_t* __restrict__ countPtr, uint32_t cullMask, float tMin, float tMaxArg,
             int perRayTMax, int sceneFast, int hitsOnly, int kLeafThreshold, int kRefillThreshold, unsigned int* __restrict__ rayCounter,
             unsigned long long* __restrict__ counters, StreamIn streamIn) {
    chain_begin();
    if (countPtr) {   // batch size produced on the device (path-tracer bounces): no host round trip
        count = min(count, *countPtr);
        if (streamIn.raysPerBlock && blockIdx.x >= max(streamIn.minBlocks, count / streamIn.raysPerBlock + 1u)) return;
    }
    __shared__ int stack[kStack][kTraceBlock];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const unsigned ltMask = (1u << lane) - 1u;
    // (streaming: the permutation is written chunk by chunk while the kernel runs and is always used; usePerm is null)
    const uint32_t* __restrict__ perm = (permIn && (STREAM || *usePerm)) ? permIn : nullptr;
    // batches kept in their own (coherent) order refill eagerly; reordered ones do better refilling half a warp at a time
    if (!perm && !STREAM) kRefillThreshold = min(kRefillThreshold, 6);
    bool reported = true;   // streaming: has this lane's finished ray been counted in chunkDone yet?

    bool alive = false, fast = false, moreRays = true, overflow = false;
    uint32_t ray = 0, sp = 0, tlasIndex = kTlasInvalid;
    int nodePtr = 0, curInst = 0, hitID = -1, hitInst = 0, matOffset = 0;
    float o[3] = {0, 0, 0}, d[3] = {1, 1, 1}, rc[3] = {1, 1, 1};
    float tMax = tMaxArg, hitT = 0.0f, baryU = 0.0f, baryV = 0.0f;
    float transparency = 1.0f;   // HitAnyTransparency's accumulator; reported in direction.w
    const float4* nodes = sc.tlasNodes;
    const float4* tris = nullptr;
    uint32_t cTlas = 0, cInst = 0, cBlas = 0, cTri = 0, cMaxSp = 1;

    const bool inPlace = in == out || hitsOnly;   // origin, ID and direction are already where they belong (or not wanted)
    auto finish = [&]() {   // PackRay, common.hsh:61-73 (+ barycentrics in the two lanes GLSL leaves unwritten)
        // origin, ID and direction were passed through when the ray was fetched; only the hit fields are written here
        if (hitsOnly) {   // compact 16-byte hit stream (the ray's `hit` vec4) for the multi-GPU gather
            out[ray] = make_float4(hitT, __int_as_float(hitID), __int_as_float(hitInst), (ANY && OPACITY) ? transparency : baryV);
        } else {
            reinterpret_cast<float*>(out + 3 * size_t(ray) + 1)[3] = (ANY && OPACITY) ? transparency : baryU;
            out[3 * size_t(ray) + 2] = make_float4(hitT, __int_as_float(hitID), __int_as_float(hitInst), baryV);
        }
        alive = false;
    };
    auto set_ray = [&](const float oo[3], const float dd[3]) {
#pragma unroll
        for (int a = 0; a < 3; a++) { o[a] = oo[a]; d[a] = dd[a]; }
        fast = fast_ok(o, d, sceneFast);
#pragma unroll
        for (int a = 0; a < 3; a++) rc[a] = fast ? __frcp_rn(d[a]) : 0.0f;
    };
    auto back_to_tlas = [&]() {   // bvh.hsh:217-220: restore the world-space ray when the stack drops below tlasIndex
        const float4 r0 = in[3 * size_t(ray)], r1 = in[3 * size_t(ray) + 1];
        const float oo[3] = {r0.x, r0.y, r0.z}, dd[3] = {r1.x, r1.y, r1.z};
        set_ray(oo, dd);
        tlasIndex = kTlasInvalid;
        nodes = sc.tlasNodes;
    };
    auto pop = [&]() {
        nodePtr = stack[--sp][tid];
        if (sp == 0u) finish();
        else if (sp < tlasIndex && tlasIndex != kTlasInvalid) back_to_tlas();
    };

    while (true) {
        const bool inner = alive && nodePtr >= 0;
        const bool leafy = alive && nodePtr < 0;
        const unsigned mI = __ballot_sync(kFull, inner), mL = __ballot_sync(kFull, leafy);
        const int nI = __popc(mI), nL = __popc(mL), nDead = 32 - nI - nL;

        bool refill = moreRays && (nDead >= kRefillThreshold || nI + nL == 0);
        if (STREAM && refill && nI + nL > 0) {
            // rays still arriving: do not claim rays that are not here yet while this warp has work — peek (non-binding) and
            // keep traversing instead of stalling the live lanes behind the upload
            unsigned arrived = 0, claimed = 0;
            if (lane == 0) { arrived = ld_volatile_u32(streamIn.watermark); claimed = ld_volatile_u32(rayCounter); }
            arrived = __shfl_sync(kFull, arrived, 0);
            claimed = __shfl_sync(kFull, claimed, 0);
            if (claimed + unsigned(nDead) > arrived && arrived < count) refill = false;
        }
        if (refill) {
            // ---- fetch new rays for the idle lanes (traceClosest.csh:18-30)
            const unsigned mDead = ~(mI | mL);
            if (STREAM) {   // report the rays these lanes finished since the last refill, per chunk, one atomic per chunk and warp
                __threadfence();        // their results are written before the count that releases the chunk's download
                const bool mine = !alive && !reported;
                const unsigned chunk = mine ? ray / streamIn.chunkRays : 0xffffffffu;
                unsigned pending = __ballot_sync(kFull, mine);
                while (pending) {
                    const int leader = __ffs(pending) - 1;
                    const unsigned c = __shfl_sync(kFull, chunk, leader);
                    const unsigned same = __ballot_sync(kFull, mine && chunk == c);
                    if (int(lane) == leader) atomicAdd(streamIn.chunkDone + c, unsigned(__popc(same)));
                    pending &= ~same;
                }
                if (mine) reported = true;   // live lanes keep their flag: their ray is still on its way
            }
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(rayCounter, unsigned(nDead));
            base = __shfl_sync(kFull, base, 0);
            if (base >= count || base + unsigned(nDead) >= count) moreRays = false;
            if (STREAM && base < count) {   // wait until every ray this warp has just claimed has been uploaded
                const unsigned need = min(base + unsigned(nDead), count);
                unsigned arrived = 0;
                unsigned spins = 0;
                do {
                    if (lane == 0) arrived = ld_volatile_u32(streamIn.watermark);
                    arrived = __shfl_sync(kFull, arrived, 0);
                    if (arrived < need) __nanosleep(256);
                } while (arrived < need && ++spins < streamIn.spinBound);   // bounded: a broken upload must not hang the GPU
                if (arrived < need) overflow = true;                  // reported as a failed call
            }
            if (!alive) {
                const unsigned idx = base + __popc(mDead & ltMask);
                if (idx < count && idx >= base) {
                    ray = perm ? (STREAM ? __ldcg(perm + idx) : perm[idx]) : idx;   // longest-first fetch order; results still go to the ray's own slot
                    if (STREAM) reported = false;
                    float4 r0, r1, r2;
                    if (STREAM) {   // L1-bypassing loads: these bytes were written by the copy engine while the kernel runs
                        r0 = __ldcg(in + 3 * size_t(ray)); r1 = __ldcg(in + 3 * size_t(ray) + 1); r2 = __ldcg(in + 3 * size_t(ray) + 2);
                    } else {                    // the three quarters of a ray share cache lines: let L1 serve the second and third
                        r0 = in[3 * size_t(ray)]; r1 = in[3 * size_t(ray) + 1]; r2 = in[3 * size_t(ray) + 2];
                    }
                    if (!inPlace) { out[3 * size_t(ray)] = r0; out[3 * size_t(ray) + 1] = r1; }
                    const int id = __float_as_int(r0.w);
                    hitID = -1;
                    hitInst = __float_as_int(r2.z);
                    hitT = 0.0f;
                    baryU = baryV = 0.0f;
                    transparency = 1.0f;
                    alive = true;
                    if (id < 0) {
                        finish();
                    } else {
                        tMax = (ANY && perRayTMax) ? r2.x : tMaxArg;
                        hitT = tMax;   // HitClosest: ray.hitDistance = tMax (bvh.hsh:202); any-hit reports tMax on a miss
                        if ((r1.x != r1.x) || (r1.y != r1.y) || (r1.z != r1.z)) {   // isnan3, bvh.hsh:204
                            finish();
                        } else {
                            const float oo[3] = {r0.x, r0.y, r0.z}, dd[3] = {r1.x, r1.y, r1.z};
                            set_ray(oo, dd);
                            sp = 1u;
                            tlasIndex = kTlasInvalid;
                            nodePtr = 0;
                            curInst = 0;
                            nodes = sc.tlasNodes;
                            stack[0][tid] = 0;
                        }
                    }
                }
            }
            continue;
        }
        if (nI + nL == 0) break;

        if (nL >= kLeafThreshold || nI == 0) {
            if (leafy) {
                if (sp < tlasIndex) {
                    // CheckInstance, bvh.hsh:172-189: vec4(o,1) * M and vec4(d,0) * M, no renormalisation.
                    const int inst = ~nodePtr;
                    const float4* I = sc.instances + 4 * size_t(inst);
                    const Float8 iA = ldg256(I), iB = ldg256(I + 2);
                    const float4 c0 = iA.a, c1 = iA.b, c2 = iB.a, c3 = iB.b;
                    if (COUNT) cInst++;
                    float no[3], nd[3];
                    no[0] = __fadd_rn(dot3(o[0], o[1], o[2], c0.x, c0.y, c0.z), __fmul_rn(1.0f, c0.w));
                    no[1] = __fadd_rn(dot3(o[0], o[1], o[2], c1.x, c1.y, c1.z), __fmul_rn(1.0f, c1.w));
                    no[2] = __fadd_rn(dot3(o[0], o[1], o[2], c2.x, c2.y, c2.z), __fmul_rn(1.0f, c2.w));
                    nd[0] = __fadd_rn(dot3(d[0], d[1], d[2], c0.x, c0.y, c0.z), __fmul_rn(0.0f, c0.w));
                    nd[1] = __fadd_rn(dot3(d[0], d[1], d[2], c1.x, c1.y, c1.z), __fmul_rn(0.0f, c1.w));
                    nd[2] = __fadd_rn(dot3(d[0], d[1], d[2], c2.x, c2.y, c2.z), __fmul_rn(0.0f, c2.w));
                    curInst = inst;
                    const int meshPtr = __float_as_int(c3.x);
                    if (OPACITY) matOffset = __float_as_int(c3.y);
                    const uint32_t mask = uint32_t(__float_as_int(c3.w));
                    nodePtr = 0;
                    if ((mask & cullMask) > 0u) {
                        set_ray(no, nd);
                        tlasIndex = sp;
                        nodes = sc.blasNodes[meshPtr];
                        tris = OPACITY ? sc.triangles[meshPtr] : sc.bvhTris[meshPtr];
                    } else {
                        // Culled by the mask. HitClosest and the *Transparency variants restore the world-space ray at the
                        // top of their next TLAS iteration (bvh.hsh:218-220, :302-304, :469-471), so o/d simply stay as they
                        // are. Plain HitAny restores only after leaving a BLAS (:387-390): it walks on through the TLAS with
                        // the ray CheckInstance has already transformed, and a later culled instance transforms it again.
                        if (ANY && !OPACITY) set_ray(no, nd);
                        pop();
                    }
                } else {
                    // CheckLeafClosest (bvh.hsh:44-72) / CheckLeaf (:74-104); with OPACITY CheckLeafClosestTransparency
                    // (:106-135) / CheckLeafTransparency (:137-170) over the 96-byte triangles
                    int triPtr = ~nodePtr;
                    bool end = false, hit = false;
                    const float tmaxLeaf = ANY ? tMax : hitT;
                    float leafTransparency = transparency;
                    while (!end && !(ANY && !OPACITY && hit)) {
                        float4 a, b, c, d0 = make_float4(0, 0, 0, 0);
                        float triOpacity = 1.0f;
                        if (OPACITY) {
                            const float4* T = tris + 6 * size_t(triPtr);   // 96-byte records: three 256-bit loads
                            const Float8 tA = ldg256(T), tB = ldg256(T + 2), tC = ldg256(T + 4);
                            a = tA.a; b = tA.b; c = tB.a; d0 = tB.b;
                            const float4 d1 = tC.a, d2 = tC.b;
                            end = d1.z > 0.0f;
                            triOpacity = d2.w;   // < 0: textured, resolved with the hit's barycentrics below
                        } else {
                            const float4* T = tris + 3 * size_t(triPtr);
                            a = __ldg(T); b = __ldg(T + 1); c = __ldg(T + 2);
                            end = a.w > 0.0f;
                        }
                        if (COUNT) cTri++;
                        float sol[3];
                        const bool inside = tri_test(o, d, a, b, c, sol);
                        if (inside && sol[0] > tMin && sol[0] < tmaxLeaf) {
                            // tri.opacity < 0 ? GetOpacity(tri, sol.yz, materialOffset, 0) : tri.opacity (bvh.hsh:127, :163); the
                            // closest-hit variant only evaluates it for a candidate that is nearer than the current hit
                            if (OPACITY && triOpacity < 0.0f && (ANY || sol[0] < hitT)) triOpacity = get_opacity(sc, d0, sol[1], sol[2], matOffset);
                            if (OPACITY && ANY) {
                                hitT = sol[0];
                                hitID = triPtr;
                                hitInst = curInst;
                                leafTransparency = __fmul_rn(leafTransparency, __fsub_rn(1.0f, triOpacity));
                            } else if (ANY || (sol[0] < hitT && (!OPACITY || triOpacity > 0.0f))) {
                                hitT = sol[0];
                                hitID = triPtr;
                                hitInst = curInst;
                                baryU = sol[1];
                                baryV = sol[2];
                                if (ANY) hit = true;
                            }
                        }
                        triPtr++;
                    }
                    if (ANY && OPACITY) {   // bvh.hsh:489-492: transparency *= CheckLeafTransparency(..., transparency)
                        transparency = __fmul_rn(transparency, leafTransparency);
                        if (transparency < 0.000001f) transparency = 0.0f;
                        hit = !(transparency > 0.0f);
                    }
                    if (ANY && hit) finish(); else pop();
                }
            }
        } else if (inner) {
            // inner node of the TLAS or of the current BLAS — UnpackNode, bvh.hsh:21-37
            const float4* N = nodes + 4 * size_t(nodePtr);
            const Float8 nA = ldg256(N), nB = ldg256(N + 2);
            const float4 n0 = nA.a, n1 = nA.b, n2 = nB.a, n3 = nB.b;
            if (COUNT) { if (sp < tlasIndex) cTlas++; else cBlas++; }
            const float llo[3] = {n0.x, n0.y, n0.z}, lhi[3] = {n0.w, n1.x, n1.y};
            const float rlo[3] = {n1.z, n1.w, n2.x}, rhi[3] = {n2.y, n2.z, n2.w};
            const int leftPtr = __float_as_int(n3.x), rightPtr = __float_as_int(n3.y);
            const float tfar = ANY ? tMax : hitT;
            float hitL = 0.0f, hitR = 0.0f;
            bool iL, iR;
            if (fast) {
                iL = slab_fast(o, d, rc, llo, lhi, tMin, tfar, hitL);
                iR = slab_fast(o, d, rc, rlo, rhi, tMin, tfar, hitR);
            } else {
                iL = slab(o, d, llo, lhi, tMin, tfar, hitL);
                iR = slab(o, d, rlo, rhi, tMin, tfar, hitR);
            }
            int pushPtr;
            if (!ANY) {
                const bool leftFirst = hitL <= hitR;
                nodePtr = leftFirst ? leftPtr : rightPtr;
                pushPtr = leftFirst ? rightPtr : leftPtr;
            } else {
                nodePtr = iL ? leftPtr : rightPtr;
                pushPtr = rightPtr;
            }
            if (iL && iR) {
                if (sp < kStack) { stack[sp][tid] = pushPtr; sp++; } else overflow = true;   // entry dropped -> ATLAS_RT_ERR_STACK
                if (COUNT && sp > cMaxSp) cMaxSp = sp;
            } else if (!iL && !iR) {
                pop();
            }
        }
    }

    if (STREAM) {   // the rays finished after this warp's last refill
        __threadfence();
        const bool mine = !reported;
        const unsigned chunk = mine ? ray / streamIn.chunkRays : 0xffffffffu;
        unsigned pending = __ballot_sync(kFull, mine);
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const unsigned c = __shfl_sync(kFull, chunk, leader);
            const unsigned same = __ballot_sync(kFull, mine && chunk == c);
            if (int(lane) == leader) atomicAdd(streamIn.chunkDone + c, unsigned(__popc(same)));
            pending &= ~same;
        }
    }
    if (overflow) atomicAdd(&counters[5], 1ull);
    if (COUNT) {
        const unsigned v[4] = {cTlas, cInst, cBlas, cTri};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // per-lane counts fit 32 bits; the warp sum may not, so add in two halves
            const unsigned lo16 = __reduce_add_sync(kFull, v[k] & 0xffffu), hi16 = __reduce_add_sync(kFull, v[k] >> 16);
            if (lane == 0) atomicAdd(&counters[k], (unsigned long long)lo16 + ((unsigned long long)hi16 << 16));
        }
        const unsigned mx = __reduce_max_sync(kFull, cMaxSp);
        if (lane == 0) atomicMax(&counters[4], (unsigned long long)mx);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Longest-first fetch order. A launch cannot end before its longest ray has walked its (strictly sequential) chain of
// node fetches, so with a plain FIFO queue the last-started long rays leave the machine draining for ~0.5 ms of a
// 1.5 ms launch (measured: time = 0.51 ms + 1.01 ms per million rays). Rays are therefore bucketed by an ESTIMATE of
// their work — the length of their path inside the scene box, 64 buckets — and fetched longest bucket first, so the
// expensive rays start early and the short ones fill the tail. Only the fetch order changes: each ray's traversal,
// and so every output bit, is the same, and results are written to the ray's original slot. The estimate itself needs
// no exactness (plain fast arithmetic).
constexpr int kCostBuckets = 64;
constexpr int kSortBlock = 256, kSortPerThread = 8;

__device__ __forceinline__ uint32_t ray_cost_bucket(const float4 r0, const float4 r1, const float lo[3], const float hi[3], float invDiag) {
    const float o[3] = {r0.x, r0.y, r0.z}, d[3] = {r1.x, r1.y, r1.z};
    float tn = 0.0f, tf = 3.0e38f;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float inv = 1.0f / d[a];
        const float t0 = (lo[a] - o[a]) * inv, t1 = (hi[a] - o[a]) * inv;
        tn = fmaxf(tn, fminf(t0, t1));   // fminf/fmaxf drop NaNs
        tf = fminf(tf, fmaxf(t0, t1));
    }
    const float len = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * fmaxf(tf - tn, 0.0f);
    const float f = fminf(fmaxf(len * invDiag, 0.0f), 1.0f);   // NaN -> 0
    const uint32_t q = min(uint32_t(kCostBuckets - 1), uint32_t(f * float(kCostBuckets)));
    return uint32_t(kCostBuckets - 1) - q;   // bucket 0 = longest
}

__device__ __forceinline__ void scene_box(const float4* __restrict__ tlasNodes, float lo[3], float hi[3], float& invDiag) {
    const float4 a = __ldg(tlasNodes), b = __ldg(tlasNodes + 1), c = __ldg(tlasNodes + 2);
    lo[0] = fminf(a.x, b.z); lo[1] = fminf(a.y, b.w); lo[2] = fminf(a.z, c.x);
    hi[0] = fmaxf(a.w, c.y); hi[1] = fmaxf(b.x, c.z); hi[2] = fmaxf(b.y, c.w);
    const float e[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    const float diag = sqrtf(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    invDiag = diag > 0.0f ? 1.0f / diag : 0.0f;
}

__global__ void __launch_bounds__(kSortBlock)
ray_cost_histogram(const float4* __restrict__ rays, uint32_t count, const uint32_t* __restrict__ countPtr, const float4* __restrict__ tlasNodes,
                   uint8_t* __restrict__ bucketOf, unsigned int* __restrict__ hist) {
    chain_begin();
    if (countPtr) count = min(count, *countPtr);
    __shared__ unsigned int sh[kCostBuckets + 1];   // [kCostBuckets] = pairs of neighbouring rays that are coherent
    if (threadIdx.x <= kCostBuckets) sh[threadIdx.x] = 0;
    __syncthreads();
    float lo[3], hi[3], invDiag;
    scene_box(tlasNodes, lo, hi, invDiag);
    const uint32_t base = blockIdx.x * (kSortBlock * kSortPerThread);
    unsigned coherent = 0;
#pragma unroll
    for (int k = 0; k < kSortPerThread; k++) {
        const uint32_t i = base + k * kSortBlock + threadIdx.x;
        float4 r0 = make_float4(0, 0, 0, 0), r1 = r0;
        if (i < count) {
            r0 = rays[3 * size_t(i)];
            r1 = rays[3 * size_t(i) + 1];
            const uint32_t b = ray_cost_bucket(r0, r1, lo, hi, invDiag);
            bucketOf[i] = uint8_t(b);
            atomicAdd(&sh[b], 1u);
        }
        // is the next ray in the buffer (the next lane) a near copy of this one? (same origin within 1 % of the scene
        // diagonal, directions within ~18 degrees) — true for rayGen's pixel tiles, false for random or bounced rays
        const float nox = __shfl_down_sync(kFull, r0.x, 1), noy = __shfl_down_sync(kFull, r0.y, 1), noz = __shfl_down_sync(kFull, r0.z, 1);
        const float ndx = __shfl_down_sync(kFull, r1.x, 1), ndy = __shfl_down_sync(kFull, r1.y, 1), ndz = __shfl_down_sync(kFull, r1.z, 1);
        if ((threadIdx.x & 31) != 31 && i + 1 < count) {
            const float dx = nox - r0.x, dy = noy - r0.y, dz = noz - r0.z;
            const float dd = r1.x * ndx + r1.y * ndy + r1.z * ndz;
            const float l0 = r1.x * r1.x + r1.y * r1.y + r1.z * r1.z, l1 = ndx * ndx + ndy * ndy + ndz * ndz;
            const bool near = (dx * dx + dy * dy + dz * dz) * invDiag * invDiag <= 1.0e-4f;
            coherent += (near && dd > 0.0f && dd * dd >= 0.9f * l0 * l1) ? 1u : 0u;
        }
    }
    coherent = __reduce_add_sync(kFull, coherent);
    if ((threadIdx.x & 31) == 0 && coherent) atomicAdd(&sh[kCostBuckets], coherent);
    __syncthreads();
    if (threadIdx.x <= kCostBuckets && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void ray_cost_offsets(unsigned int* __restrict__ hist, uint32_t count, const uint32_t* __restrict__ countPtr) {   // exclusive prefix over 64 buckets, in place (one warp)
    chain_begin();
    if (countPtr) count = min(count, *countPtr);
    const unsigned lane = threadIdx.x;
    // hist[kCostBuckets + 1] = 1 when the batch should be reordered: coherent batches (most neighbours are near copies)
    // keep their own order, which is worth more than starting the long rays first
    if (lane == 0) hist[kCostBuckets + 1] = (2ull * hist[kCostBuckets] < uint64_t(count) * 31ull / 32ull) ? 1u : 0u;
    const unsigned a = hist[lane], b = hist[32 + lane];
    unsigned sa = a, sb = b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned ua = __shfl_up_sync(kFull, sa, off), ub = __shfl_up_sync(kFull, sb, off);
        if (lane >= unsigned(off)) { sa += ua; sb += ub; }
    }
    const unsigned totalA = __shfl_sync(kFull, sa, 31);
    hist[lane] = sa - a;
    hist[32 + lane] = totalA + sb - b;
}

__global__ void __launch_bounds__(kSortBlock)
ray_cost_scatter(const uint8_t* __restrict__ bucketOf, uint32_t count, const uint32_t* __restrict__ countPtr, unsigned int* __restrict__ offsets,
                 uint32_t* __restrict__ perm, uint32_t indexBase, int always) {
    chain_begin();
    if (countPtr) count = min(count, *countPtr);
    __shared__ unsigned int cnt[kCostBuckets], base[kCostBuckets];
    if (offsets[kCostBuckets + 1] == 0u) {   // coherent batch: the trace kernel ignores the permutation ...
        if (always) {                        // ... except in a streaming launch, which always reads it: buffer order
            for (uint32_t k = 0; k < uint32_t(kSortPerThread); k++) {
                const uint32_t i = blockIdx.x * (kSortBlock * kSortPerThread) + k * kSortBlock + threadIdx.x;
                if (i < count) perm[i] = i + indexBase;
            }
        }
        return;
    }
    if (threadIdx.x < kCostBuckets) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t first = blockIdx.x * (kSortBlock * kSortPerThread);
    uint32_t rank[kSortPerThread], bk[kSortPerThread];
#pragma unroll
    for (int k = 0; k < kSortPerThread; k++) {
        const uint32_t i = first + k * kSortBlock + threadIdx.x;
        bk[k] = i < count ? bucketOf[i] : 0xffu;
        rank[k] = bk[k] != 0xffu ? atomicAdd(&cnt[bk[k]], 1u) : 0u;
    }
    __syncthreads();
    if (threadIdx.x < kCostBuckets) base[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&offsets[threadIdx.x], cnt[threadIdx.x]) : 0u;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortPerThread; k++) {
        const uint32_t i = first + k * kSortBlock + threadIdx.x;
        if (bk[k] != 0xffu) perm[base[bk[k]] + rank[k]] = i + indexBase;
    }
}

// Largest coordinate magnitude of the scene: node 0 of the TLAS and of every BLAS bounds everything below it.
__global__ void scene_bounds(const float4* __restrict__ tlasNodes, uint32_t tlasNodeCount, const float4* const* __restrict__ blasNodes,
                             const uint32_t* __restrict__ blasNodeCounts, uint32_t meshCount, unsigned int* __restrict__ maxAbsBits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > meshCount) return;
    const float4* N = nullptr;
    if (i == meshCount) { if (tlasNodeCount) N = tlasNodes; }
    else if (blasNodeCounts[i]) N = blasNodes[i];
    else atomicMax(maxAbsBits, 0x7f800000u);   // root-leaf BLAS: its synthetic node holds +-FLT_MAX boxes, keep such scenes on __fdiv_rn
    if (!N) return;
    const float4 a = N[0], b = N[1], c = N[2];
    const float v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    float m = 0.0f;
#pragma unroll
    for (int k = 0; k < 12; k++) m = fmaxf(m, fabsf(v[k]));
    if (!(m == m)) m = __int_as_float(0x7f800000);
    atomicMax(maxAbsBits, __float_as_uint(m));   // non-negative floats order like their bit patterns
}

// Smallest NON-ZERO coordinate magnitude in any node box of the scene (TLAS = mesh index meshCount): decides whether rays
// with an origin component of exactly zero may take the fast division path (see div_by_rcp).
__global__ void __launch_bounds__(256)
scene_min_abs(const float4* __restrict__ tlasNodes, uint32_t tlasNodeCount, const float4* const* __restrict__ blasNodes,
              const uint32_t* __restrict__ blasNodeCounts, uint32_t meshCount, unsigned int* __restrict__ minAbsBits) {
    const uint32_t mesh = blockIdx.y;
    const float4* N = mesh == meshCount ? tlasNodes : blasNodes[mesh];
    const uint32_t nodes = mesh == meshCount ? tlasNodeCount : blasNodeCounts[mesh];
    unsigned int best = 0x7f800000u;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < uint64_t(nodes) * 3; i += uint64_t(gridDim.x) * blockDim.x) {
        const float4 v = N[(i / 3) * 4 + i % 3];   // the three float4s that hold the two boxes
        const float c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned int b = __float_as_uint(fabsf(c[k]));
            if (b != 0u && b < best) best = b;
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if ((threadIdx.x & 31u) == 0u && best != 0x7f800000u) atomicMin(minAbsBits, best);
}

// Runs behind the streaming trace kernel on the same stream: every ray is finished by then, so every chunk's completion
// count is set to its target. The per-warp counts inside the kernel only release chunks EARLY; this makes the release
// unconditional, so a download stream can never be left waiting.
__global__ void release_chunks(unsigned int* chunkDone, uint32_t chunkRays, uint32_t count) {
    chain_begin();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t b = uint64_t(c) * chunkRays;
    if (b < count) atomicMax(chunkDone + c, unsigned(min(uint64_t(count), b + chunkRays) - b));
}

}   // namespace

namespace {
// flags[0] |= 1 when some triangle's opacity (d2.w of the 96-byte record) is not exactly 1; flags[0] |= 2 when some instance
// lacks the shadow bit. blockIdx.y = mesh, or meshCount for the instance array.
__global__ void __launch_bounds__(256)
scene_opacity_scan(const float4* const* __restrict__ triangles, const unsigned long long* __restrict__ triCounts, uint32_t meshCount,
                   const float4* __restrict__ instances, unsigned long long instanceCount, unsigned int* __restrict__ flags) {
    const uint32_t mesh = blockIdx.y;
    unsigned int bad = 0;
    if (mesh == meshCount) {
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < instanceCount; i += (unsigned long long)gridDim.x * blockDim.x)
            if ((uint32_t(__float_as_int(instances[4 * i + 3].w)) & ATLAS_RT_MASK_SHADOW) == 0u) bad |= 2u;
    } else if (triangles[mesh]) {
        const float4* T = triangles[mesh];
        for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < triCounts[mesh]; i += (unsigned long long)gridDim.x * blockDim.x)
            if (T[6 * i + 5].w != 1.0f) bad |= 1u;
    } else {
        bad |= 1u;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31u) == 0u && bad) atomicOr(flags, bad);
}
}   // namespace

int scene_opacity_flags(atlas_rt_context* ctx, atlas_rt_scene* scene, const uint64_t* triCounts) {
    scene->allOpaque = scene->allShadowBit = false;
    if (!scene->allShading) return ATLAS_RT_OK;
    unsigned long long* dCounts = nullptr;
    ATLAS_CUDA(ctx, dev_alloc(ctx, &dCounts, scene->meshCount));
    unsigned int* dFlags = reinterpret_cast<unsigned int*>(ctx->dCounters + 7);
    cudaError_t e = cudaMemcpyAsync(dCounts, triCounts, scene->meshCount * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(dFlags, 0, sizeof(unsigned int), ctx->stream);
    if (e == cudaSuccess) {
        scene_opacity_scan<<<dim3(32, scene->meshCount + 1), 256, 0, ctx->stream>>>(scene->triangles, dCounts, scene->meshCount, scene->instances,
                                                                                    scene->instanceCount, dFlags);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->pinned, dFlags, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    dev_free(ctx, dCounts);
    if (e != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "scene opacity scan", e);
    unsigned int bits = 0;
    memcpy(&bits, ctx->pinned, sizeof(bits));
    scene->allOpaque = (bits & 1u) == 0u;
    scene->allShadowBit = (bits & 2u) == 0u;
    return ATLAS_RT_OK;
}

// Streaming launch: one chunk of the batch has arrived in `rays` (device memory). Its rays are put in longest-first order
// (the three ordering kernels of a resident batch, over the chunk only; indices written are batch-global) and the watermark
// the persistent trace kernel polls is then moved to the end of the chunk — by a one-thread kernel behind the scatter, so
// that the permutation is complete (and visible) before any warp can claim a ray of the chunk.
namespace {
__global__ void publish_watermark(unsigned int* watermark, unsigned int value) {
    chain_begin();
    __threadfence();
    atomicMax(watermark, value);
    __threadfence_system();
}
}   // namespace

int launch_chunk_sort(atlas_rt_context* ctx, const atlas_rt_scene* scene, cudaStream_t st, const float4* rays, uint32_t n, uint32_t indexBase,
                      uint8_t* bucketOf, unsigned int* hist, uint32_t* perm, unsigned int* watermark) {
    ATLAS_CUDA(ctx, cudaMemsetAsync(hist, 0, (kCostBuckets + 2) * sizeof(unsigned int), st));
    const uint32_t sortGrid = (n + kSortBlock * kSortPerThread - 1) / (kSortBlock * kSortPerThread);
    const bool pdl = ctx->chainLaunch != 0;
    ATLAS_CUDA(ctx, launch_chain(false, ray_cost_histogram, sortGrid, kSortBlock, 0, st, rays, n, (const uint32_t*)nullptr, scene->tlas->nodes, bucketOf, hist));
    ATLAS_CUDA(ctx, launch_chain(pdl, ray_cost_offsets, 1, 32, 0, st, hist, n, (const uint32_t*)nullptr));
    ATLAS_CUDA(ctx, launch_chain(pdl, ray_cost_scatter, sortGrid, kSortBlock, 0, st, bucketOf, n, (const uint32_t*)nullptr, hist, perm, indexBase, 1));
    ATLAS_CUDA(ctx, launch_chain(pdl, publish_watermark, 1, 1, 0, st, watermark, indexBase + n));
    ctx->launches += 4;
    return ATLAS_RT_OK;
}

int launch_release_chunks(atlas_rt_context* ctx, unsigned int* chunkDone, uint32_t chunkRays, uint32_t count, uint32_t chunks) {
    release_chunks<<<(chunks + 63) / 64, 64, 0, ctx->stream>>>(chunkDone, chunkRays, count);
    ATLAS_LAUNCH_CHECK(ctx);
    return ATLAS_RT_OK;
}

namespace {
}   // namespace

int scene_fast_flag(atlas_rt_context* ctx, atlas_rt_scene* scene, const uint32_t* dNodeCounts) {
    unsigned int* dMax = reinterpret_cast<unsigned int*>(ctx->dCounters + 7);
    ATLAS_CUDA(ctx, cudaMemsetAsync(dMax, 0, sizeof(unsigned int), ctx->stream));
    scene_bounds<<<(scene->meshCount + 1 + 127) / 128, 128, 0, ctx->stream>>>(scene->tlas->nodes, uint32_t(scene->tlas->nodeCount), scene->blasNodes,
                                                                               dNodeCounts, scene->meshCount, dMax);
    ATLAS_LAUNCH_CHECK(ctx);
    ATLAS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dMax, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned int bits = 0;
    memcpy(&bits, ctx->pinned, sizeof(bits));
    float m;
    memcpy(&m, &bits, sizeof(m));
    scene->fastDivision = (m <= kPosHi) ? 1 : 0;
    if (scene->fastDivision) {
        unsigned int* dMin = dMax;
        const unsigned int inf = 0x7f800000u;
        memcpy(ctx->pinned, &inf, sizeof(inf));
        ATLAS_CUDA(ctx, cudaMemcpyAsync(dMin, ctx->pinned, sizeof(unsigned int), cudaMemcpyHostToDevice, ctx->stream));
        scene_min_abs<<<dim3(64, scene->meshCount + 1), 256, 0, ctx->stream>>>(scene->tlas->nodes, uint32_t(scene->tlas->nodeCount), scene->blasNodes,
                                                                               dNodeCounts, scene->meshCount, dMin);
        ATLAS_LAUNCH_CHECK(ctx);
        ATLAS_CUDA(ctx, cudaMemcpyAsync(ctx->pinned, dMin, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        memcpy(&bits, ctx->pinned, sizeof(bits));
        memcpy(&m, &bits, sizeof(m));
        if (m >= kCoordLo) scene->fastDivision |= 2;
    }
    return ATLAS_RT_OK;
}

int launch_trace(atlas_rt_context* ctx, const atlas_rt_scene* scene, const float4* dIn, float4* dOut, uint64_t count,
                 uint32_t cullMask, float tMin, float tMax, bool any, bool perRayTMax, bool counters, bool resetCounters, bool opacity,
                 cudaStream_t st, int queueSlot, const uint32_t* dCount, bool hitsOnly, const unsigned int* watermark, unsigned int* chunkDone,
                 uint32_t chunkRays, const uint32_t* streamPerm, int chain) {
    if (!st) st = ctx->stream;
    const bool chained = chain < 0 ? ctx->chainLaunch != 0 : chain != 0;   // programmatic dependent launches between this call's kernels
    if (count == 0) return ATLAS_RT_OK;
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 rays in one batch");
    SceneDev sc{scene->tlas->nodes, scene->instances, scene->blasNodes, scene->bvhTris, scene->triangles,
                scene->materials, scene->textures, scene->materialCount, scene->textureCount};
    const uint32_t n = uint32_t(count);
    // small batches get fewer persistent warps so that each still refills its lanes many times (>= traceRaysPerWarp rays per warp)
    const uint32_t wantBlocks = std::max<uint32_t>(uint32_t(ctx->smCount) * uint32_t(ctx->traceMinBlocksPerSM), n / (uint32_t(ctx->traceRaysPerWarp) * (kTraceBlock / 32)));
    // a streaming launch leaves room on every SM for the small ordering kernels that run beside it on another stream
    const uint32_t blocksPerSM = watermark ? uint32_t(std::min(ctx->traceBlocksPerSM, ctx->streamBlocksPerSM))
                                           : (opacity ? std::min(ctx->traceBlocksPerSM, kBlocksPerSM - 1) : ctx->traceBlocksPerSM);
    const uint32_t grid = std::min<uint32_t>(std::min<uint32_t>((n + kTraceBlock - 1) / kTraceBlock, wantBlocks), uint32_t(ctx->smCount) * blocksPerSM);
    const int lt = ctx->traceLeafThreshold, rt = ctx->traceRefillThreshold;
    // words 0-5: visit counters + overflow flag (kept across the chunks of one pipelined call), word 6: the ray queue head
    // (words 8-15: ray-queue heads, one per compute stream, for launches that run side by side)
    if (resetCounters) ATLAS_CUDA(ctx, cudaMemsetAsync(ctx->dCounters, 0, 6 * sizeof(unsigned long long), st));
    ATLAS_CUDA(ctx, cudaMemsetAsync(ctx->dCounters + 8 + queueSlot, 0, sizeof(unsigned long long), st));
    unsigned int* rayCounter = reinterpret_cast<unsigned int*>(ctx->dCounters + 8 + queueSlot);
    // ---- longest-first fetch order for batches large enough to have a tail worth hiding
    uint32_t* perm = nullptr;
    uint8_t* bucketOf = nullptr;
    unsigned int* hist = nullptr;
    const bool shrink = dCount != nullptr && ctx->ptLanes > 1;
    const StreamIn streamIn{watermark, chunkDone, chunkRays ? chunkRays : 1u, 1u << ctx->streamSpinLog2,
                            shrink ? uint32_t(ctx->traceRaysPerWarp) * (kTraceBlock / 32) : 0u, uint32_t(ctx->smCount) * uint32_t(ctx->traceMinBlocksPerSM)};
    if (ctx->traceLongestFirst && n >= uint32_t(ctx->traceLongestFirstMin) && scene->tlas->nodeCount > 0 && !watermark) {
        ATLAS_CUDA(ctx, dev_alloc_on(st, &perm, n));
        ATLAS_CUDA(ctx, dev_alloc_on(st, &bucketOf, n));
        ATLAS_CUDA(ctx, dev_alloc_on(st, &hist, kCostBuckets + 2));
        ATLAS_CUDA(ctx, cudaMemsetAsync(hist, 0, (kCostBuckets + 2) * sizeof(unsigned int), st));
        const uint32_t sortGrid = (n + kSortBlock * kSortPerThread - 1) / (kSortBlock * kSortPerThread);
        const bool pdl = chained;
        ATLAS_CUDA(ctx, launch_chain(pdl, ray_cost_histogram, sortGrid, kSortBlock, 0, st, dIn, n, dCount, scene->tlas->nodes, bucketOf, hist));
        ctx->launches++;
        ATLAS_CUDA(ctx, launch_chain(pdl, ray_cost_offsets, 1, 32, 0, st, hist, n, dCount));
        ctx->launches++;
        ATLAS_CUDA(ctx, launch_chain(pdl, ray_cost_scatter, sortGrid, kSortBlock, 0, st, bucketOf, n, dCount, hist, perm, 0u, 0));
        ctx->launches++;
    }
    const int pr = perRayTMax ? 1 : 0, sf = scene->fastDivision, ho = hitsOnly ? 1 : 0;
    L2Window win;
    if (ctx->l2PersistMB > 0 && scene->hotNodes && scene->hotBytes) {
        win.base = scene->hotNodes;
        win.bytes = std::min(scene->hotBytes, ctx->l2WindowMax);
        // the carve-out holds hitRatio * window bytes: scale the ratio down when the window is larger than the carve-out
        win.hitRatio = std::min(ctx->l2HitRatio, float(double(size_t(ctx->l2PersistMB) << 20) / double(win.bytes)));
    }
    cudaError_t launchErr = cudaSuccess;
#define ATLAS_TRACE_LAUNCH(A, C, O) \
    launchErr = (watermark && !C) ? launch_chain_w(chained, win, trace_kernel<A, false, O, true>, grid, kTraceBlock, 0, st, sc, dIn, dOut, streamPerm, (const unsigned int*)nullptr, n, dCount, cullMask, tMin, tMax, pr, sf, ho, lt, rt, rayCounter, ctx->dCounters, streamIn) : \
                launch_chain_w(chained, win, trace_kernel<A, C, O, false>, grid, kTraceBlock, 0, st, sc, dIn, dOut, perm, hist ? hist + kCostBuckets + 1 : nullptr, n, dCount, cullMask, tMin, tMax, pr, sf, ho, lt, rt, rayCounter, ctx->dCounters, streamIn)
    if (opacity) {
        if (any) { if (counters) ATLAS_TRACE_LAUNCH(true, true, true); else ATLAS_TRACE_LAUNCH(true, false, true); }
        else { if (counters) ATLAS_TRACE_LAUNCH(false, true, true); else ATLAS_TRACE_LAUNCH(false, false, true); }
    } else {
        if (any) { if (counters) ATLAS_TRACE_LAUNCH(true, true, false); else ATLAS_TRACE_LAUNCH(true, false, false); }
        else { if (counters) ATLAS_TRACE_LAUNCH(false, true, false); else ATLAS_TRACE_LAUNCH(false, false, false); }
    }
#undef ATLAS_TRACE_LAUNCH
    dev_free_on(st, perm);
    dev_free_on(st, bucketOf);
    dev_free_on(st, hist);
    ctx->launches++;
    if (launchErr != cudaSuccess) return fail(ctx, ATLAS_RT_ERR_CUDA, "trace launch", launchErr);
    return ATLAS_RT_OK;
}

}   // namespace atlas
