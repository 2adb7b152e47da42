// pathtrace.cu — primary-ray generation and the diffuse bounce of the path tracer (rayGen.csh / rayHit.csh).
#include <cstring>

#include "common.cuh"

namespace atlas {
namespace {

// rayGen.csh:25-91. One thread per (pixel, sample). Storage index: 8x8 pixel tiles are contiguous (64 rays x samples)
// so that a 32-lane warp traces neighbouring pixels; the right and bottom borders that do not fill a tile follow.
__global__ void raygen_kernel(atlas_rt_camera cam, uint32_t width, uint32_t height, uint32_t samples,
                              const float* __restrict__ jitter, float4* __restrict__ out) {
    const uint32_t x = blockIdx.x * 8u + (threadIdx.x & 7u), y = blockIdx.y * 8u + (threadIdx.x >> 3);
    const uint32_t s = blockIdx.z;
    if (x >= width || y >= height) return;
    const float jx = jitter ? jitter[2 * s] : 0.5f, jy = jitter ? jitter[2 * s + 1] : 0.5f;
    const float cu = __fdiv_rn(__fadd_rn(float(x), jx), float(width));
    const float cv = __fdiv_rn(__fadd_rn(float(y), jy), float(height));
    float d[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        d[k] = __fsub_rn(__fadd_rn(__fadd_rn(cam.origin[k], __fmul_rn(cam.right[k], cu)), __fmul_rn(cam.bottom[k], cv)), cam.eye[k]);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    const int id = int((y * width + x) * samples + s);   // Flatten2D(pixel, resolution) * samples + sample
    // tile-coherent storage order (rayGen.csh:53-80)
    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
    const uint32_t gx = blockIdx.x, gy = blockIdx.y, local = threadIdx.x;
    uint32_t index;
    if (gx < perfX && gy < perfY) {
        index = local + (gy * perfX + gx) * 64u;
    } else if (gx >= perfX && gy < perfY) {
        const uint32_t off = perfX * perfY * 64u;
        index = y * overX + (x - perfX * 8u) + off;
    } else {
        const uint32_t off = perfX * perfY * 64u + overX * perfY * 8u;
        index = x * overY + (y - perfY * 8u) + off;   // Flatten2D(localID.yx, overlappingPixels.yx)
    }
    const size_t slot = size_t(index) * samples + s;
    out[3 * slot + 0] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], __int_as_float(id));
    out[3 * slot + 1] = make_float4(__fdiv_rn(d[0], len), __fdiv_rn(d[1], len), __fdiv_rn(d[2], len), 0.0f);
    out[3 * slot + 2] = make_float4(0.0f, __int_as_float(0), 0.0f, 0.0f);
}


// ------------------------------------------------------------------------------------------------ hash RNG
// data/shader/common/random.hsh:5-48 — Bob Jenkins' one-at-a-time hash, floats built from the low 23 bits.
__device__ __forceinline__ uint32_t hash1(uint32_t x) {
    x += (x << 10u);
    x ^= (x >> 6u);
    x += (x << 3u);
    x ^= (x >> 11u);
    x += (x << 15u);
    return x;
}
__device__ __forceinline__ float float_construct(uint32_t m) { return __fsub_rn(__uint_as_float((m & 0x007FFFFFu) | 0x3F800000u), 1.0f); }
// float random(float x, inout float seed): random(vec2(x, seed)); seed += 1.0
__device__ __forceinline__ float random2(float x, float& seed) {
    const float r = float_construct(hash1(__float_as_uint(x) ^ hash1(__float_as_uint(seed))));
    seed = __fadd_rn(seed, 1.0f);
    return r;
}

constexpr float kEpsilon = 0.1f;          // EPSILON, raytracer/common.hsh:9
constexpr float kPi = 3.14159265358979f;  // common/PI.hsh

struct Surf {
    float P[3], N[3], G[3];   // hit point, shading normal (facing the viewer), geometry normal
};

// World-space hit point and geometric normal of a hit (surface.hsh:72-98 reduced to what a Lambertian, untextured,
// two-sided surface needs): P = origin + t * direction; the triangle normal cross(v0 - v1, v0 - v2) is taken in
// instance space and carried to world space with the inverse-transpose, i.e. the transpose of the instance's
// inverseMatrix rows; it is flipped towards the viewer like `flipNormal && twoSided`.
__device__ __forceinline__ Surf surface_at(const float4* __restrict__ instances, const float4* const* __restrict__ bvhTris,
                                            const float o[3], const float d[3], float t, int hitID, int hitInst) {
    Surf s;
    const float4* I = instances + 4 * size_t(hitInst);
    const float4 c0 = __ldg(I), c1 = __ldg(I + 1), c2 = __ldg(I + 2), c3 = __ldg(I + 3);
    const float4* T = bvhTris[__float_as_int(c3.x)] + 3 * size_t(hitID);
    const float4 a = __ldg(T), b = __ldg(T + 1), c = __ldg(T + 2);
    const float e0[3] = {a.x - b.x, a.y - b.y, a.z - b.z}, e1[3] = {a.x - c.x, a.y - c.y, a.z - c.z};
    const float n[3] = {e0[1] * e1[2] - e1[1] * e0[2], e0[2] * e1[0] - e1[2] * e0[0], e0[0] * e1[1] - e1[0] * e0[1]};
    float g[3] = {c0.x * n[0] + c1.x * n[1] + c2.x * n[2], c0.y * n[0] + c1.y * n[1] + c2.y * n[2], c0.z * n[0] + c1.z * n[1] + c2.z * n[2]};
    const float inv = 1.0f / sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    const bool flip = (g[0] * d[0] + g[1] * d[1] + g[2] * d[2]) > 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        s.G[k] = g[k] * inv * (flip ? -1.0f : 1.0f);
        s.N[k] = s.G[k];
        s.P[k] = o[k] + t * d[k];
    }
    return s;
}

// After the closest-hit trace: environment for misses, and one shadow ray per hit towards the directional light
// (rayHit.csh:165-171 and CheckVisibility :327-337: origin = P + N * EPSILON, direction = L, tMax = lightDistance -
// 2 EPSILON with lightDistance = INF for a directional light). Rays that need no shadow ray get ID -1 so the any-hit
// batch passes them through.
__global__ void shade_prepare(const float4* __restrict__ rays, uint32_t count, atlas_rt_bounce_params prm,
                              const float4* __restrict__ instances, const float4* const* __restrict__ bvhTris,
                              float4* __restrict__ shadowRays) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float4 r0 = rays[3 * size_t(i)], r1 = rays[3 * size_t(i) + 1], r2 = rays[3 * size_t(i) + 2];
    const int hitID = __float_as_int(r2.y);
    float4 s0 = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)), s1 = make_float4(0.0f, 1.0f, 0.0f, 0.0f);
    float4 s2 = make_float4(ATLAS_RT_INF, __int_as_float(-1), 0.0f, 0.0f);
    if (hitID >= 0 && __float_as_int(r0.w) >= 0) {
        const float o[3] = {r0.x, r0.y, r0.z}, d[3] = {r1.x, r1.y, r1.z};
        const Surf sf = surface_at(instances, bvhTris, o, d, r2.x, hitID, __float_as_int(r2.z));
        const float ndl = sf.N[0] * prm.light_dir[0] + sf.N[1] * prm.light_dir[1] + sf.N[2] * prm.light_dir[2];
        if (ndl > 0.0f) {
            s0 = make_float4(sf.P[0] + sf.N[0] * kEpsilon, sf.P[1] + sf.N[1] * kEpsilon, sf.P[2] + sf.N[2] * kEpsilon, r0.w);
            s1 = make_float4(prm.light_dir[0], prm.light_dir[1], prm.light_dir[2], 0.0f);
            s2.x = ATLAS_RT_INF - 2.0f * kEpsilon;
        }
    }
    shadowRays[3 * size_t(i)] = s0;
    shadowRays[3 * size_t(i) + 1] = s1;
    shadowRays[3 * size_t(i) + 2] = s2;
}

// Rest of rayHit.csh for a Lambertian surface: direct light with the shadow ray's visibility (EvaluateDirectLight
// :209-235), cosine-weighted bounce from the hash RNG keyed by (ray.ID, seed) (EvaluateIndirectLight :237-325 with the
// diffuse branch of brdfSample.hsh:8-32; the RNG draws are consumed in the reference's order), Russian roulette, then
// either accumulation of a finished path or a warp-aggregated append of the surviving ray (tracing.hsh:70-77).
__global__ void shade_finish(const float4* __restrict__ rays, const float4* __restrict__ payloadIn, const float4* __restrict__ shadowRays,
                             uint32_t count, atlas_rt_bounce_params prm, const float4* __restrict__ instances,
                             const float4* const* __restrict__ bvhTris, float4* __restrict__ raysOut, float4* __restrict__ payloadOut,
                             float* __restrict__ accum, unsigned int* __restrict__ outCount) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool survive = false;
    float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, n2 = n0, p0 = n0, p1 = n0;
    if (i < count) {
        const float4 r0 = rays[3 * size_t(i)], r1 = rays[3 * size_t(i) + 1], r2 = rays[3 * size_t(i) + 2];
        const int id = __float_as_int(r0.w), hitID = __float_as_int(r2.y);
        if (id >= 0) {
            float rad[3] = {0.0f, 0.0f, 0.0f}, thr[3] = {1.0f, 1.0f, 1.0f};
            if (prm.bounce > 0u) {
                const float4 a = payloadIn[2 * size_t(i)], b = payloadIn[2 * size_t(i) + 1];
                rad[0] = a.x; rad[1] = a.y; rad[2] = a.z;
                thr[0] = b.x; thr[1] = b.y; thr[2] = b.z;
            }
            float no[3] = {r0.x, r0.y, r0.z}, nd[3] = {r1.x, r1.y, r1.z};
            if (hitID < 0) {
                // environment (rayHit.csh:166-171): min(sky * throughput, 10), path ends
#pragma unroll
                for (int k = 0; k < 3; k++) { rad[k] += fminf(prm.sky_radiance[k] * thr[k], 10.0f); thr[k] = 0.0f; }
            } else {
                const float o[3] = {r0.x, r0.y, r0.z}, d[3] = {r1.x, r1.y, r1.z};
                const Surf sf = surface_at(instances, bvhTris, o, d, r2.x, hitID, __float_as_int(r2.z));
                // ---- direct light
                const float ndl = sf.N[0] * prm.light_dir[0] + sf.N[1] * prm.light_dir[1] + sf.N[2] * prm.light_dir[2];
                float direct[3] = {0.0f, 0.0f, 0.0f};
                if (ndl > 0.0f) {
                    const bool occluded = __float_as_int(shadowRays[3 * size_t(i) + 2].y) >= 0;
                    if (!occluded) {
#pragma unroll
                        for (int k = 0; k < 3; k++) direct[k] = thr[k] * (prm.albedo[k] / kPi) * prm.light_radiance[k] * ndl;
                    }
                }
                if (prm.bounce > 0u) {   // radiance clamp of indirect bounces (rayHit.csh:190-195)
                    const float mx = fmaxf(fmaxf(direct[0], fmaxf(direct[1], direct[2])), 10.0f);
#pragma unroll
                    for (int k = 0; k < 3; k++) direct[k] *= 10.0f / mx;
                }
#pragma unroll
                for (int k = 0; k < 3; k++) rad[k] += direct[k];
                // ---- indirect: RNG draws in the reference's order
                float curSeed = prm.seed;
                const float raySeed = float(id);
                (void)random2(raySeed, curSeed);            // refraction choice (opacity 1: never refracts)
                (void)random2(raySeed, curSeed);            // specular / diffuse choice (Lambertian: always diffuse)
                const float u0 = random2(raySeed, curSeed), u1 = random2(raySeed, curSeed);
                const float rr = sqrtf(u0), phi = 2.0f * kPi * u1;
                const float lx = rr * cosf(phi), ly = rr * sinf(phi), lz = sqrtf(1.0f - u0);
                const float* N = sf.N;
                const float up[3] = {fabsf(N[2]) < 0.999f ? 0.0f : 1.0f, 0.0f, fabsf(N[2]) < 0.999f ? 1.0f : 0.0f};
                float tg[3] = {up[1] * N[2] - N[1] * up[2], up[2] * N[0] - N[2] * up[0], up[0] * N[1] - N[0] * up[1]};
                const float ti = 1.0f / sqrtf(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
                tg[0] *= ti; tg[1] *= ti; tg[2] *= ti;
                const float bt[3] = {N[1] * tg[2] - tg[1] * N[2], N[2] * tg[0] - tg[2] * N[0], N[0] * tg[1] - tg[0] * N[1]};
                float L[3];
#pragma unroll
                for (int k = 0; k < 3; k++) L[k] = tg[k] * lx + bt[k] * ly + N[k] * lz;
                const float li = 1.0f / sqrtf(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    nd[k] = L[k] * li;
                    no[k] = sf.P[k] + (-d[k]) * kEpsilon;     // ray.origin = P + V * EPSILON, V = -direction
                    thr[k] *= prm.albedo[k];                  // reflectance * NdotL / pdf for cosine-weighted Lambert
                }
                // Russian roulette (rayHit.csh:307-323)
                float prob = fminf(fmaxf(fmaxf(thr[0], fmaxf(thr[1], thr[2])), 0.01f), 0.99f);
                prob = prm.bounce < 3u ? fminf(3.0f * prob, 1.0f) : prob;
                const bool killed = random2(raySeed, curSeed) > prob;
                const bool below = (nd[0] * sf.G[0] + nd[1] * sf.G[1] + nd[2] * sf.G[2]) <= 0.0f;
#pragma unroll
                for (int k = 0; k < 3; k++) thr[k] = (killed || below) ? 0.0f : thr[k] / prob;
            }
            const float energy = thr[0] + thr[1] + thr[2];
            if (energy == 0.0f || prm.bounce == prm.max_bounces) {
                float* px = accum + 4 * size_t(uint32_t(id) / prm.samples);
                atomicAdd(px + 0, rad[0]); atomicAdd(px + 1, rad[1]); atomicAdd(px + 2, rad[2]); atomicAdd(px + 3, 1.0f);
            } else {
                survive = true;
                n0 = make_float4(no[0], no[1], no[2], r0.w);
                n1 = make_float4(nd[0], nd[1], nd[2], 0.0f);
                n2 = make_float4(0.0f, __int_as_float(-1), 0.0f, 0.0f);
                p0 = make_float4(rad[0], rad[1], rad[2], 0.0f);
                p1 = make_float4(thr[0], thr[1], thr[2], 0.0f);
            }
        }
    }
    // ---- compaction: one atomic per warp, survivors keep their relative order inside the warp
    const unsigned m = __ballot_sync(0xffffffffu, survive);
    if (m) {
        const unsigned lane = threadIdx.x & 31u;
        unsigned base = 0;
        if (lane == (__ffs(m) - 1)) base = atomicAdd(outCount, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (survive) {
            const size_t dst = base + __popc(m & ((1u << lane) - 1u));
            raysOut[3 * dst] = n0; raysOut[3 * dst + 1] = n1; raysOut[3 * dst + 2] = n2;
            payloadOut[2 * dst] = p0; payloadOut[2 * dst + 1] = p1;
        }
    }
}

}   // namespace
}   // namespace atlas

using namespace atlas;

extern "C" {

int atlas_rt_generate_primary_rays(atlas_rt_context* ctx, const atlas_rt_camera* camera, uint32_t width, uint32_t height,
                                   uint32_t samples, const float* jitter, void* rays_out, uint32_t flags) {
    if (!ctx || !camera || !rays_out || !width || !height || !samples) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t count = uint64_t(width) * height * samples;
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 primary rays");
    const bool devOut = flags & ATLAS_RT_DEVICE_OUTPUT;
    float4* dOut = static_cast<float4*>(rays_out);
    float4* tmp = nullptr;
    float* dJit = nullptr;
    if (!devOut) { ATLAS_CUDA(ctx, dev_alloc(ctx, &tmp, count * 3)); dOut = tmp; }
    if (jitter) {
        ATLAS_CUDA(ctx, dev_alloc(ctx, &dJit, size_t(samples) * 2));
        ATLAS_CUDA(ctx, cudaMemcpyAsync(dJit, jitter, size_t(samples) * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    const dim3 grid((width + 7) / 8, (height + 7) / 8, samples);
    raygen_kernel<<<grid, 64, 0, ctx->stream>>>(*camera, width, height, samples, dJit, dOut);
    ATLAS_LAUNCH_CHECK(ctx);
    if (!devOut) ATLAS_CUDA(ctx, copy_out(ctx, rays_out, dOut, count * 48, false));
    dev_free(ctx, tmp);
    dev_free(ctx, dJit);
    if (!(flags & ATLAS_RT_ASYNC) || jitter) ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

int atlas_rt_pathtrace_bounce(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_bounce_params* params,
                              const void* rays_in, const void* payload_in, uint64_t count, void* rays_out,
                              void* payload_out, float* accum, uint64_t* out_count, uint32_t flags) {
    if (!ctx || !scene || scene->ctx->device != ctx->device || !params || !rays_out || !payload_out || !accum || !out_count || (count && !rays_in))
        return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    if ((flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT)) != (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT))
        return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_pathtrace_bounce works on device-resident ray / payload / accumulation buffers");
    if (params->bounce > 0 && !payload_in) return fail(ctx, ATLAS_RT_ERR_INVALID, "payload_in required after the first bounce");
    if (rays_in == rays_out) return fail(ctx, ATLAS_RT_ERR_INVALID, "rays_out must not alias rays_in (survivors are compacted)");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    *out_count = 0;
    if (count == 0) return ATLAS_RT_OK;
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 rays");
    const uint32_t n = uint32_t(count);
    float4* rays = const_cast<float4*>(static_cast<const float4*>(rays_in));
    float4* shadow = nullptr;
    unsigned int* dCount = nullptr;
    ATLAS_CUDA(ctx, dev_alloc(ctx, &shadow, size_t(n) * 3));
    ATLAS_CUDA(ctx, dev_alloc(ctx, &dCount, 1));
    ATLAS_CUDA(ctx, cudaMemsetAsync(dCount, 0, sizeof(unsigned int), ctx->stream));
    // 1. closest hit, in place (traceClosest.csh)
    int rc = launch_trace(ctx, scene, rays, rays, n, ATLAS_RT_MASK_ALL, 0.0f, ATLAS_RT_INF, false, false, false);
    if (rc == ATLAS_RT_OK) {
        // 2. shadow rays, 3. any-hit over them with the shadow mask, 4. shade / bounce / compact
        shade_prepare<<<(n + 127) / 128, 128, 0, ctx->stream>>>(rays, n, *params, scene->instances, scene->bvhTris, shadow);
        ctx->launches++;
        rc = launch_trace(ctx, scene, shadow, shadow, n, ATLAS_RT_MASK_SHADOW, 0.0f, ATLAS_RT_INF, true, true, false);
    }
    if (rc == ATLAS_RT_OK) {
        shade_finish<<<(n + 127) / 128, 128, 0, ctx->stream>>>(rays, static_cast<const float4*>(payload_in), shadow, n, *params, scene->instances,
                                                               scene->bvhTris, static_cast<float4*>(rays_out), static_cast<float4*>(payload_out),
                                                               accum, dCount);
        ctx->launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "kernel launch", e);
    }
    if (rc == ATLAS_RT_OK) {
        cudaError_t e = cudaMemcpyAsync(ctx->pinned, dCount, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, ATLAS_RT_ERR_CUDA, "read back survivor count", e);
        else { unsigned int c = 0; memcpy(&c, ctx->pinned, sizeof(c)); *out_count = c; }
    }
    dev_free(ctx, shadow);
    dev_free(ctx, dCount);
    return rc;
}

}
