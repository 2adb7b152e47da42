// trace.cu — closest-hit / any-hit traversal of the two-level (TLAS -> BLAS) scene in the engine's flattened layout.
//
// Restates, per ray and in the same visit order, data/shader/raytracer/bvh.hsh:191-273 (HitClosest) and :359-441
// (HitAny) with CheckInstance (:172-189), CheckLeafClosest / CheckLeaf (:44-104), UnpackNode (:21-37),
// IntersectAABB / IntersectTriangle (intersections.hsh:19-58) and the batch wrapper traceClosest.csh:12-36.
// Visit order is part of the contract: ties in t are resolved by "first visited wins" (strict < in bvh.hsh:62-63), so
// one ray's node/triangle sequence must never be reordered; parallelism is across rays only.
//
// Data movement: a node is 64 B = 4 x LDG.128 through the read-only path, a triangle 48 B = 3 x LDG.128, an instance
// 64 B = 4 x LDG.128; rays are read and written once as 3 x 128-bit each. The per-ray stack (32 node pointers, the
// reference's STACK_SIZE) lives in shared memory laid out [entry][lane] so a warp's accesses never bank-conflict.
#include "common.cuh"

namespace atlas {

namespace {

constexpr int kTraceBlock = 128;
constexpr uint32_t kStack = ATLAS_RT_STACK_SIZE;
constexpr uint32_t kTlasInvalid = kStack + 2;   // TLAS_INVALID, bvh.hsh:17

struct SceneDev {
    const float4* tlasNodes;
    const float4* instances;
    const float4* const* blasNodes;
    const float4* const* bvhTris;
};

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

template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(kTraceBlock)
trace_kernel(SceneDev sc, const float4* __restrict__ in, float4* __restrict__ out, uint32_t count, uint32_t cullMask,
             float tMin, float tMaxArg, int perRayTMax, unsigned long long* __restrict__ counters) {
    __shared__ int stack[kStack][kTraceBlock];
    const uint32_t tid = threadIdx.x;
    const uint32_t i = blockIdx.x * kTraceBlock + tid;
    if (i >= count) return;

    const float4 r0 = in[3 * size_t(i)], r1 = in[3 * size_t(i) + 1], r2 = in[3 * size_t(i) + 2];
    const int id = __float_as_int(r0.w);
    const float o0[3] = {r0.x, r0.y, r0.z}, d0[3] = {r1.x, r1.y, r1.z};

    // traceClosest.csh:24-25
    int hitID = -1, hitInst = __float_as_int(r2.z);
    float hitT = 0.0f, baryU = 0.0f, baryV = 0.0f;
    uint32_t cTlas = 0, cInst = 0, cBlas = 0, cTri = 0, cMaxSp = 1;
    bool overflow = false;

    if (id >= 0) {
        const float tMax = (ANY && perRayTMax) ? r2.x : tMaxArg;
        hitT = tMax;   // HitClosest: ray.hitDistance = tMax (bvh.hsh:202); any-hit reports tMax on a miss
        const bool nanDir = (d0[0] != d0[0]) || (d0[1] != d0[1]) || (d0[2] != d0[2]);   // isnan3, bvh.hsh:204
        if (!nanDir) {
            uint32_t sp = 1u, tlasIndex = kTlasInvalid;
            int nodePtr = 0, curInst = 0;
            float o[3] = {o0[0], o0[1], o0[2]}, d[3] = {d0[0], d0[1], d0[2]};
            const float4* __restrict__ nodes = sc.tlasNodes;
            const float4* __restrict__ tris = nullptr;
            bool hit = false;
            stack[0][tid] = 0;
            while (sp != 0u && !(ANY && hit)) {
                const bool inTlas = sp < tlasIndex;
                if (inTlas) {
                    o[0] = o0[0]; o[1] = o0[1]; o[2] = o0[2];
                    d[0] = d0[0]; d[1] = d0[1]; d[2] = d0[2];
                    tlasIndex = kTlasInvalid;
                    nodes = sc.tlasNodes;
                }
                if (nodePtr < 0) {
                    if (inTlas) {
                        // CheckInstance, bvh.hsh:172-189: vec4(o,1) * M and vec4(d,0) * M, no renormalisation.
                        const int inst = ~nodePtr;
                        const float4* I = sc.instances + 4 * size_t(inst);
                        const float4 c0 = __ldg(I), c1 = __ldg(I + 1), c2 = __ldg(I + 2), c3 = __ldg(I + 3);
                        if (COUNT) cInst++;
                        float no[3], nd[3];
                        no[0] = __fadd_rn(dot3(o[0], o[1], o[2], c0.x, c0.y, c0.z), __fmul_rn(1.0f, c0.w));
                        no[1] = __fadd_rn(dot3(o[0], o[1], o[2], c1.x, c1.y, c1.z), __fmul_rn(1.0f, c1.w));
                        no[2] = __fadd_rn(dot3(o[0], o[1], o[2], c2.x, c2.y, c2.z), __fmul_rn(1.0f, c2.w));
                        nd[0] = __fadd_rn(dot3(d[0], d[1], d[2], c0.x, c0.y, c0.z), __fmul_rn(0.0f, c0.w));
                        nd[1] = __fadd_rn(dot3(d[0], d[1], d[2], c1.x, c1.y, c1.z), __fmul_rn(0.0f, c1.w));
                        nd[2] = __fadd_rn(dot3(d[0], d[1], d[2], c2.x, c2.y, c2.z), __fmul_rn(0.0f, c2.w));
                        o[0] = no[0]; o[1] = no[1]; o[2] = no[2];
                        d[0] = nd[0]; d[1] = nd[1]; d[2] = nd[2];
                        curInst = inst;
                        const int meshPtr = __float_as_int(c3.x);
                        const uint32_t mask = uint32_t(__float_as_int(c3.w));
                        nodePtr = 0;
                        if ((mask & cullMask) > 0u) {
                            tlasIndex = sp;
                            nodes = sc.blasNodes[meshPtr];
                            tris = sc.bvhTris[meshPtr];
                        } else {
                            nodePtr = stack[--sp][tid];
                        }
                    } else {
                        // CheckLeafClosest (bvh.hsh:44-72) / CheckLeaf (:74-104)
                        int triPtr = ~nodePtr;
                        bool end = false;
                        const float tmaxLeaf = ANY ? tMax : hitT;
                        while (!end && !(ANY && hit)) {
                            const float4* T = tris + 3 * size_t(triPtr);
                            const float4 a = __ldg(T), b = __ldg(T + 1), c = __ldg(T + 2);
                            end = a.w > 0.0f;
                            if (COUNT) cTri++;
                            float sol[3];
                            const bool inside = tri_test(o, d, a, b, c, sol);
                            if (inside && sol[0] > tMin && sol[0] < tmaxLeaf) {
                                if (ANY || sol[0] < hitT) {
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
                        nodePtr = stack[--sp][tid];
                    }
                } else {
                    // inner node of the TLAS or the current BLAS — UnpackNode, bvh.hsh:21-37
                    const float4* N = nodes + 4 * size_t(nodePtr);
                    const float4 n0 = __ldg(N), n1 = __ldg(N + 1), n2 = __ldg(N + 2), n3 = __ldg(N + 3);
                    if (COUNT) { if (inTlas) cTlas++; else cBlas++; }
                    const float llo[3] = {n0.x, n0.y, n0.z}, lhi[3] = {n0.w, n1.x, n1.y};
                    const float rlo[3] = {n1.z, n1.w, n2.x}, rhi[3] = {n2.y, n2.z, n2.w};
                    const int leftPtr = __float_as_int(n3.x), rightPtr = __float_as_int(n3.y);
                    const float tfar = ANY ? tMax : hitT;
                    float hitL = 0.0f, hitR = 0.0f;
                    const bool iL = slab(o, d, llo, lhi, tMin, tfar, hitL);
                    const bool iR = slab(o, d, rlo, rhi, tMin, tfar, hitR);
                    int pushPtr;
                    if (!ANY) {
                        const bool leftFirst = hitL <= hitR;
                        nodePtr = leftFirst ? leftPtr : rightPtr;
                        pushPtr = leftFirst ? rightPtr : leftPtr;
                    } else {
                        nodePtr = iL ? leftPtr : rightPtr;
                        pushPtr = rightPtr;
                    }
                    if (!iL && !iR) nodePtr = stack[--sp][tid];
                    if (iL && iR) {
                        if (sp < kStack) stack[sp][tid] = pushPtr; else overflow = true;
                        sp++;
                        if (sp > kStack) { sp = kStack; }   // entry dropped; flagged as ATLAS_RT_ERR_STACK
                    }
                    if (COUNT && sp > cMaxSp) cMaxSp = sp;
                }
            }
        }
    }

    // PackRay, common.hsh:61-73 (+ barycentrics in the two lanes GLSL leaves unwritten)
    out[3 * size_t(i)] = r0;
    out[3 * size_t(i) + 1] = make_float4(r1.x, r1.y, r1.z, baryU);
    out[3 * size_t(i) + 2] = make_float4(hitT, __int_as_float(hitID), __int_as_float(hitInst), baryV);

    if (overflow) atomicAdd(&counters[5], 1ull);
    if (COUNT) {
        // warp-aggregate before the global atomics
        const unsigned m = __activemask();
        unsigned long long v[4] = {cTlas, cInst, cBlas, cTri};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned s = unsigned(v[k]);
            s = __reduce_add_sync(m, s);
            if ((tid & 31u) == (__ffs(m) - 1)) atomicAdd(&counters[k], (unsigned long long)s);
        }
        const unsigned mx = __reduce_max_sync(m, cMaxSp);
        if ((tid & 31u) == (__ffs(m) - 1)) atomicMax(&counters[4], (unsigned long long)mx);
    }
}

}   // namespace

int launch_trace(atlas_rt_context* ctx, const atlas_rt_scene* scene, const float4* dIn, float4* dOut, uint64_t count,
                 uint32_t cullMask, float tMin, float tMax, bool any, bool perRayTMax, bool counters) {
    if (count == 0) return ATLAS_RT_OK;
    if (count > 0xffffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^32-1 rays in one batch");
    SceneDev sc{scene->tlas->nodes, scene->instances, scene->blasNodes, scene->bvhTris};
    const uint32_t n = uint32_t(count);
    const uint32_t grid = (n + kTraceBlock - 1) / kTraceBlock;
    ATLAS_CUDA(ctx, cudaMemsetAsync(ctx->dCounters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    const int pr = perRayTMax ? 1 : 0;
    if (any) {
        if (counters) trace_kernel<true, true><<<grid, kTraceBlock, 0, ctx->stream>>>(sc, dIn, dOut, n, cullMask, tMin, tMax, pr, ctx->dCounters);
        else trace_kernel<true, false><<<grid, kTraceBlock, 0, ctx->stream>>>(sc, dIn, dOut, n, cullMask, tMin, tMax, pr, ctx->dCounters);
    } else {
        if (counters) trace_kernel<false, true><<<grid, kTraceBlock, 0, ctx->stream>>>(sc, dIn, dOut, n, cullMask, tMin, tMax, pr, ctx->dCounters);
        else trace_kernel<false, false><<<grid, kTraceBlock, 0, ctx->stream>>>(sc, dIn, dOut, n, cullMask, tMin, tMax, pr, ctx->dCounters);
    }
    ATLAS_LAUNCH_CHECK(ctx);
    return ATLAS_RT_OK;
}

}   // namespace atlas
