// pathtrace.cu — the path tracer around the traversal: primary-ray generation, ray binning, the hit shader of one bounce
// and the whole bounce loop kept on the device.
//
// Restates data/shader/pathtracer/rayGen.csh:25-91, data/shader/pathtracer/rayHit.csh:56-337 with what it calls
// (raytracer/surface.hsh:44-138, raytracer/direct.hsh:13-87 for one directional light, brdf/brdfSample.hsh:8-92,
// brdf/brdfEval.hsh:8-32, brdf/brdf.hsh:9-53, brdf/surface.hsh:26-43, common/random.hsh:5-48, raytracer/common.hsh:75-98),
// raytracer/tracing.hsh:18-29 + binning.csh + binningOffset.csh, and the dispatch sequence of
// renderer/PathTracingRenderer.cpp:146-192 / renderer/helper/RayTracingHelper.cpp:262-405 (device-written counts instead of
// traceDispatch.csh + DispatchIndirect: the kernels read the ray count from device memory, the host never waits).
// Scope: untextured materials (RaytraceMaterial table) with textured OPACITY, interpolated vertex normals from the 96-byte
// triangles, one directional light, constant environment. The library is compiled with -fmad=false: the expressions below
// are evaluated exactly as written (GLSL operation order).
#include <cstring>
#include <utility>
#include <vector>

#include "common.cuh"

namespace atlas {
namespace {

constexpr float kEpsilon = 0.1f;                 // EPSILON, raytracer/common.hsh:9
constexpr float kPi = 3.14159265358979f;         // common/PI.hsh
constexpr float kInvPi = 0.31830988618f;
constexpr unsigned kFullMask = 0xffffffffu;
constexpr uint32_t kNoSlot = 0xffffffffu;

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
__device__ __forceinline__ V3 normalize(V3 v) { return v * (1.0f / sqrtf(dot(v, v))); }
__device__ __forceinline__ float saturate(float x) { return gl_clamp(x, 0.0f, 1.0f); }
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
__device__ __forceinline__ V3 mix3(V3 x, V3 y, float a) { return {mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)}; }

// ------------------------------------------------------------------------------------------------ hash RNG
// common/random.hsh:5-48 — Bob Jenkins' one-at-a-time hash, floats built from the low 23 bits.
__host__ __device__ inline uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
__device__ __forceinline__ float float_construct(uint32_t m) { return __uint_as_float((m & 0x007FFFFFu) | 0x3F800000u) - 1.0f; }
__device__ __forceinline__ float random_seeded(float x, float& seed) {   // float random(float x, inout float seed)
    const float r = float_construct(hash1(__float_as_uint(x) ^ hash1(__float_as_uint(seed))));
    seed = seed + 1.0f;
    return r;
}

// Which storage slots of a frame a call renders. Contiguous: [begin, end). Interleaved (parts > 0): the blocks of `block`
// slots whose number is congruent to `part` modulo `parts` — neighbouring blocks go to different GPUs, so sky and
// geometry, cheap and expensive parts of the image, are dealt evenly. local index <-> slot:
//   slot = ((local / block) * parts + part) * block + local % block.
struct SlotShard {
    unsigned long long begin, end;   // contiguous range (parts == 0); with parts > 0: end = total slots of the frame
    uint32_t part, parts, block;
};
__host__ __device__ inline unsigned long long shard_slot(const SlotShard& sh, unsigned long long local) {
    if (sh.parts == 0u) return sh.begin + local;
    return ((local / sh.block) * sh.parts + sh.part) * sh.block + local % sh.block;
}
__host__ __device__ inline unsigned long long shard_count(const SlotShard& sh) {
    if (sh.parts == 0u) return sh.end - sh.begin;
    const unsigned long long blocks = (sh.end + sh.block - 1) / sh.block;            // blocks of the frame, the last one may be short
    if (sh.part >= blocks) return 0;
    const unsigned long long mine = (blocks - sh.part + sh.parts - 1) / sh.parts;    // blocks part, part + parts, ...
    const unsigned long long lastBlock = sh.part + (mine - 1) * sh.parts;
    const unsigned long long lastLen = (lastBlock == blocks - 1) ? sh.end - lastBlock * sh.block : sh.block;
    return (mine - 1) * sh.block + lastLen;
}

// ------------------------------------------------------------------------------------------------ ray generation
// rayGen.csh:25-91, one thread per STORAGE slot (the shader's slot <-> pixel mapping is a bijection, so walking the slots
// gives the same buffer and lets a caller generate any contiguous range of it — the unit of multi-GPU sharding).
__global__ void raygen_kernel(atlas_rt_camera cam, uint32_t width, uint32_t height, uint32_t samples, const float* __restrict__ jitter,
                              float jx0, float jy0, SlotShard shard, unsigned long long count, float4* __restrict__ out) {
    chain_begin();
    const uint64_t o = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;   // local index = position in the output buffer
    if (o >= count) return;
    const uint64_t slot = shard_slot(shard, o);
    const uint32_t index = uint32_t(slot / samples), s = uint32_t(slot % samples);
    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
    const uint32_t full = perfX * perfY * 64u, rightStrip = overX * perfY * 8u;
    uint32_t x, y;
    if (index < full) {                       // whole 8x8 groups, 64 consecutive slots each (rayGen.csh:62-65)
        const uint32_t group = index / 64u, local = index % 64u;
        x = (group % perfX) * 8u + (local & 7u);
        y = (group / perfX) * 8u + (local >> 3);
    } else if (index < full + rightStrip) {   // ragged right border (:66-71): Flatten2D(localID, overlappingPixels)
        const uint32_t rem = index - full;
        y = rem / overX;
        x = perfX * 8u + rem % overX;
    } else {                                  // bottom border (:72-77): Flatten2D(localID.yx, overlappingPixels.yx)
        const uint32_t rem = index - full - rightStrip;
        x = rem / overY;
        y = perfY * 8u + rem % overY;
    }
    const float jx = jitter ? jitter[2 * s] : jx0, jy = jitter ? jitter[2 * s + 1] : jy0;
    const float cu = (float(x) + jx) / float(width), cv = (float(y) + jy) / float(height);
    V3 d;
    d.x = ((cam.origin[0] + cam.right[0] * cu) + cam.bottom[0] * cv) - cam.eye[0];
    d.y = ((cam.origin[1] + cam.right[1] * cu) + cam.bottom[1] * cv) - cam.eye[1];
    d.z = ((cam.origin[2] + cam.right[2] * cu) + cam.bottom[2] * cv) - cam.eye[2];
    d = normalize(d);
    const int id = int((y * width + x) * samples + s);   // Flatten2D(pixel, resolution) * samples + sample
    out[3 * o + 0] = make_float4(cam.eye[0], cam.eye[1], cam.eye[2], __int_as_float(id));
    out[3 * o + 1] = make_float4(d.x, d.y, d.z, 0.0f);
    out[3 * o + 2] = make_float4(0.0f, __int_as_float(0), 0.0f, 0.0f);
}

// ------------------------------------------------------------------------------------------------ surface
struct Surface {   // brdf/surface.hsh:4-24 + the Material fields the shader reads
    V3 P, V, N, L, H, geometryNormal, F0, baseColor, emissive;
    float NdotL, LdotH, NdotH, NdotV, F90;
    float opacity, roughness, metalness, ao, reflectance;
};

__device__ __forceinline__ V3 unpack_unit(uint32_t c) {   // common/packing.hsh:2-13
    return {float((c >> 0) & 1023u) / 1023.0f * 2.0f - 1.0f, float((c >> 10) & 1023u) / 1023.0f * 2.0f - 1.0f, float((c >> 20) & 1023u) / 1023.0f * 2.0f - 1.0f};
}
__device__ __forceinline__ float half_lo(uint32_t w) { return __half2float(__ushort_as_half(uint16_t(w))); }
__device__ __forceinline__ float half_hi(uint32_t w) { return __half2float(__ushort_as_half(uint16_t(w >> 16))); }

struct SceneTables {
    const float4* instances;
    const float4* const* triangles;
    const uint32_t* materials;
    const TextureDev* textures;
    uint32_t materialCount, textureCount;
};

__device__ __forceinline__ void update_surface(Surface& s) {   // brdf/surface.hsh:26-43
    s.L = normalize(s.L); s.V = normalize(s.V); s.N = normalize(s.N);
    s.H = normalize(s.L + s.V);
    s.NdotL = saturate(dot(s.N, s.L));
    s.LdotH = saturate(dot(s.L, s.H));
    s.NdotH = saturate(dot(s.N, s.H));
    s.NdotV = saturate(dot(s.N, s.V));
    const float f = 0.16f * sqr(s.reflectance);
    s.F0 = mix3(V3{f, f, f}, s.baseColor, s.metalness);
    s.F90 = saturate(50.0f * dot(s.F0, V3{0.333f, 0.333f, 0.333f}));
}

__device__ __forceinline__ V3 fresnel_schlick(V3 F0, float F90, float c) {   // brdf/brdf.hsh:9-13
    const float p = powf(1.0f - c, 5.0f);
    return F0 + (V3{F90, F90, F90} - F0) * p;
}
__device__ __forceinline__ float disney_diffuse(float NdotV, float NdotL, float LdotH, float lr) {   // brdf.hsh:15-25
    const float bias = mixf(0.0f, 0.5f, lr), factor = mixf(1.0f, 1.0f / 1.51f, lr);
    const float FD90 = bias + 2.0f * LdotH * LdotH * lr;
    const float ls = fresnel_schlick(V3{1, 1, 1}, FD90, NdotL).x, vs = fresnel_schlick(V3{1, 1, 1}, FD90, NdotV).x;
    return ls * vs * factor;
}
__device__ __forceinline__ float vis_separable(float c, float alpha) {   // brdf.hsh:27-32
    const float a2 = alpha * alpha;
    return 2.0f * c / (c + sqrtf(a2 + (1 - a2) * c * c));
}
__device__ __forceinline__ float vis_correlated(float NdotL, float NdotV, float alpha) {   // brdf.hsh:34-43
    const float a2 = alpha * alpha;
    const float GGXL = NdotV * sqrtf((-NdotL * a2 + NdotL) * NdotL + a2);
    const float GGXV = NdotL * sqrtf((-NdotV * a2 + NdotV) * NdotV + a2);
    return 0.5f / (GGXL + GGXV + 0.0000001f);
}
__device__ __forceinline__ float distribution_ggx(float NdotH, float alpha) {   // brdf.hsh:45-53
    const float a2 = alpha * alpha;
    const float f = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (f * f + 0.0000001f) * kInvPi;
}
__device__ __forceinline__ V3 eval_diffuse(const Surface& s) {   // brdf/brdfEval.hsh:8-19
    const float roughness = gl_max(sqr(s.roughness), 0.00001f);
    const float dd = disney_diffuse(s.NdotV, s.NdotL, s.LdotH, roughness);
    return s.baseColor * (1.0f - s.metalness) * dd * kInvPi;
}
__device__ __forceinline__ V3 eval_specular(const Surface& s) {   // brdfEval.hsh:21-32
    const float roughness = gl_max(sqr(s.roughness), 0.00001f);
    const V3 F = fresnel_schlick(s.F0, s.F90, s.LdotH);
    const float G = vis_correlated(s.NdotV, s.NdotL, roughness), D = distribution_ggx(s.NdotH, roughness);
    return F * D * G;
}

// GetSurfaceParameters, raytracer/surface.hsh:64-138 with TransformTriangle (:44-62: the forward matrix is the inverse of
// the instance's inverse matrix, here by cofactors) and GetTriangleMaterial (:22-42). Material textures other than the
// opacity map are outside this path (their Sample*Bilinear return 1 for a negative texture id).
__device__ __forceinline__ Surface surface_at(const SceneTables& sc, const float4 r0, const float4 r1, const float4 r2) {
    const int hitID = __float_as_int(r2.y), hitInst = __float_as_int(r2.z);
    const float4* I = sc.instances + 4 * size_t(hitInst);
    const float4 c0 = __ldg(I), c1 = __ldg(I + 1), c2 = __ldg(I + 2), c3 = __ldg(I + 3);
    const int meshOffset = __float_as_int(c3.x), materialOffset = __float_as_int(c3.y);
    const float4* T = sc.triangles[meshOffset] + 6 * size_t(hitID);
    const float4 t0 = __ldg(T), t1 = __ldg(T + 1), t2 = __ldg(T + 2), d0 = __ldg(T + 3), d2 = __ldg(T + 5);
    // forward matrix
    const float a = c0.x, b = c0.y, c = c0.z, d = c1.x, e = c1.y, f = c1.z, g = c2.x, h = c2.y, i = c2.z;
    const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const float det = a * A + b * B + c * C;
    const float inv = 1.0f / det;
    float m[3][3], mt[3];
    m[0][0] = A * inv; m[0][1] = -(b * i - c * h) * inv; m[0][2] = (b * f - c * e) * inv;
    m[1][0] = B * inv; m[1][1] = (a * i - c * g) * inv;  m[1][2] = -(a * f - c * d) * inv;
    m[2][0] = C * inv; m[2][1] = -(a * h - b * g) * inv; m[2][2] = (a * e - b * d) * inv;
    for (int k = 0; k < 3; k++) mt[k] = -((m[k][0] * c0.w + m[k][1] * c1.w) + m[k][2] * c2.w);
    auto point = [&](V3 p) { return V3{((m[0][0] * p.x + m[0][1] * p.y) + m[0][2] * p.z) + mt[0], ((m[1][0] * p.x + m[1][1] * p.y) + m[1][2] * p.z) + mt[1],
                                       ((m[2][0] * p.x + m[2][1] * p.y) + m[2][2] * p.z) + mt[2]}; };
    auto dir = [&](V3 p) { return V3{(m[0][0] * p.x + m[0][1] * p.y) + m[0][2] * p.z, (m[1][0] * p.x + m[1][1] * p.y) + m[1][2] * p.z,
                                     (m[2][0] * p.x + m[2][1] * p.y) + m[2][2] * p.z}; };
    const V3 v0 = point(V3{t0.x, t0.y, t0.z}), v1 = point(V3{t1.x, t1.y, t1.z}), v2 = point(V3{t2.x, t2.y, t2.z});
    const V3 n0 = normalize(dir(unpack_unit(__float_as_uint(t0.w)))), n1 = normalize(dir(unpack_unit(__float_as_uint(t1.w)))),
             n2 = normalize(dir(unpack_unit(__float_as_uint(t2.w))));
    // material
    float mat[13] = {0.0f, 0.8f, 0.8f, 0.8f, 0.0f, 0.0f, 0.0f, 1.0f, 1.0f, 0.0f, 1.0f, 0.5f, 0.0f};
    int invertUVs = 0, twoSided = 1, useVertexColors = 0, opacityTexture = -1;
    const uint32_t mi = uint32_t(__float_as_int(d0.w) + materialOffset);
    if (sc.materials && mi < sc.materialCount) {
        const uint32_t* M = sc.materials + 23 * size_t(mi);
        for (int k = 1; k < 13; k++) mat[k] = __uint_as_float(__ldg(M + k));
        invertUVs = int(__ldg(M + 13)); twoSided = int(__ldg(M + 14)); useVertexColors = int(__ldg(M + 16)); opacityTexture = int(__ldg(M + 18));
    }
    const V3 o{r0.x, r0.y, r0.z}, dd{r1.x, r1.y, r1.z};
    // IntersectTriangle again, in world space: the ray does not carry barycentrics (surface.hsh:72-78)
    const V3 e0 = v1 - v0, e1 = v2 - v0, sv = o - v0;
    const V3 p = cross(sv, e0), q = cross(dd, e1);
    const float den = dot(q, e0);
    const float dist = dot(p, e1) / den, s = dot(q, sv) / den, t = dot(p, dd) / den;
    const float r = 1.0f - s - t;
    Surface sf;
    sf.P = o + dd * dist;
    const uint32_t w0 = __float_as_uint(d0.x), w1 = __float_as_uint(d0.y), w2 = __float_as_uint(d0.z);
    const float u = r * half_lo(w0) + s * half_lo(w1) + t * half_lo(w2);
    float v = r * half_hi(w0) + s * half_hi(w1) + t * half_hi(w2);
    V3 normal = normalize((n0 * r + n1 * s) + n2 * t);
    if (invertUVs > 0) v = 1.0f - v;
    V3 tn = normalize(cross(v0 - v1, v0 - v2));
    const bool flip = dot(tn, dd) > 0.0f;
    if (flip && twoSided > 0) tn = neg(tn);
    sf.geometryNormal = tn;
    sf.baseColor = V3{mat[1], mat[2], mat[3]};
    if (useVertexColors > 0) {
        const uint32_t k0 = __float_as_uint(d2.x), k1 = __float_as_uint(d2.y), k2 = __float_as_uint(d2.z);
        auto ch = [](uint32_t cc, int ax) { return float((cc >> (8 * ax)) & 0xffu) / 255.0f; };   // unpackUnorm4x8
        const V3 vc{r * ch(k0, 0) + s * ch(k1, 0) + t * ch(k2, 0), r * ch(k0, 1) + s * ch(k1, 1) + t * ch(k2, 1), r * ch(k0, 2) + s * ch(k1, 2) + t * ch(k2, 2)};
        sf.baseColor = sf.baseColor * vc;
    }
    sf.emissive = V3{mat[4], mat[5], mat[6]};
    sf.opacity = mat[7] * ((opacityTexture < 0 || uint32_t(opacityTexture) >= sc.textureCount) ? 1.0f : sample_r8(sc.textures[opacityTexture], u, v));
    sf.roughness = mat[8]; sf.metalness = mat[9]; sf.ao = mat[10]; sf.reflectance = mat[11];
    if (dot(normal, tn) < 0.0f && twoSided > 0) normal = neg(normal);
    sf.V = neg(dd);
    sf.N = normalize(normal);
    sf.F0 = mix3(V3{0.04f, 0.04f, 0.04f}, sf.baseColor, sf.metalness);
    sf.F90 = 1.0f;
    sf.L = V3{0, 0, 0}; sf.H = V3{0, 0, 0};
    sf.NdotL = 0.0f; sf.LdotH = 0.0f; sf.NdotH = 0.0f; sf.NdotV = 0.0f;
    return sf;
}

// SampleLight for the directional light (raytracer/direct.hsh:79-86): L = normalize(-light.N), UpdateSurface.
__device__ __forceinline__ void light_surface(Surface& sf, const atlas_rt_pt_params& prm) {
    if (prm.light_count > 0) {
        sf.L = normalize(V3{prm.light_dir[0], prm.light_dir[1], prm.light_dir[2]});
        update_surface(sf);
    } else {   // without lights the shader reads NdotV uninitialised in EvaluateIndirectLight; defined here
        sf.NdotL = 0.0f;
        sf.NdotV = saturate(dot(sf.N, sf.V));
    }
}

__device__ __forceinline__ uint32_t batch_count(uint32_t n, const uint32_t* countPtr) { return countPtr ? min(n, *countPtr) : n; }

// After the closest-hit trace: the shadow rays of CheckVisibility (rayHit.csh:327-337), origin = P + N * EPSILON,
// direction = L, compacted so that the any-hit batch holds only rays that are really cast; slotOf[i] = the shadow ray of
// ray i or kNoSlot (miss, no light, NdotL <= 0).
__global__ void __launch_bounds__(128)
shade_prepare(const float4* __restrict__ rays, uint32_t n, const uint32_t* __restrict__ countPtr, atlas_rt_pt_params prm, SceneTables sc,
              float4* __restrict__ shadowRays, uint32_t* __restrict__ slotOf, uint32_t* __restrict__ shadowCount, unsigned long long* __restrict__ traced) {
    chain_begin();
    const uint32_t count = batch_count(n, countPtr);
    if (traced && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(traced, (unsigned long long)count);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool cast = false;
    float4 s0 = make_float4(0, 0, 0, 0), s1 = s0;
    if (i < count) {
        const float4 r0 = rays[3 * size_t(i)], r1 = rays[3 * size_t(i) + 1], r2 = rays[3 * size_t(i) + 2];
        if (__float_as_int(r0.w) >= 0 && __float_as_int(r2.y) >= 0 && prm.light_count > 0) {
            Surface sf = surface_at(sc, r0, r1, r2);
            light_surface(sf, prm);
            if (sf.NdotL > 0.0f) {
                cast = true;
                const V3 o = sf.P + sf.N * kEpsilon;
                s0 = make_float4(o.x, o.y, o.z, r0.w);
                s1 = make_float4(sf.L.x, sf.L.y, sf.L.z, 0.0f);
            }
        }
    }
    const unsigned m = __ballot_sync(kFullMask, cast);
    unsigned base = 0;
    const unsigned lane = threadIdx.x & 31u;
    if (m) {
        const int leader = __ffs(m) - 1;
        if (int(lane) == leader) base = atomicAdd(shadowCount, unsigned(__popc(m)));
        base = __shfl_sync(kFullMask, base, leader);
    }
    if (i < count) {
        uint32_t slot = kNoSlot;
        if (cast) {
            slot = base + __popc(m & ((1u << lane) - 1u));
            shadowRays[3 * size_t(slot)] = s0;
            shadowRays[3 * size_t(slot) + 1] = s1;
            shadowRays[3 * size_t(slot) + 2] = make_float4(0.0f, __int_as_float(-1), 0.0f, 0.0f);
        }
        slotOf[i] = slot;
    }
}

// The rest of rayHit.csh: main (:56-158), EvaluateBounce (:160-207), EvaluateDirectLight (:209-235), EvaluateIndirectLight
// (:237-325). Finished paths are accumulated (:119-123: accum += vec4(radiance, 1)); surviving rays are appended with
// their half-precision payload (WriteRay, tracing.hsh:70-77; PackRayPayload, common.hsh:88-98), one atomic per warp.
__global__ void __launch_bounds__(128)
shade_finish(const float4* __restrict__ rays, const float4* __restrict__ payloadIn, const float4* __restrict__ shadowRays, const uint32_t* __restrict__ slotOf,
             uint32_t n, const uint32_t* __restrict__ countPtr, atlas_rt_pt_params prm, float seed, uint32_t bounce, SceneTables sc,
             float4* __restrict__ raysOut, float4* __restrict__ payloadOut, float* __restrict__ accum, uint32_t accumTileOrder, uint32_t width, uint32_t height,
             uint32_t* __restrict__ outCount, int shadowIsPlainAnyHit, SlotShard shard) {
    chain_begin();
    const uint32_t count = batch_count(n, countPtr);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool survive = false;
    float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, pay = n0;
    if (i < count) {
        const float4 r0 = rays[3 * size_t(i)], r1 = rays[3 * size_t(i) + 1], r2 = rays[3 * size_t(i) + 2];
        const int id = __float_as_int(r0.w), hitID = __float_as_int(r2.y);
        if (id >= 0) {
            V3 radiance{0, 0, 0}, throughput{1, 1, 1};
            if (bounce > 0u) {   // UnpackRayPayload
                const float4 p = payloadIn[i];
                const uint32_t w0 = __float_as_uint(p.x), w1 = __float_as_uint(p.y), w2 = __float_as_uint(p.z);
                radiance = {half_lo(w0), half_hi(w0), half_lo(w2)};
                throughput = {half_lo(w1), half_hi(w1), half_hi(w2)};
            }
            V3 o{r0.x, r0.y, r0.z}, d{r1.x, r1.y, r1.z};
            if (hitID == -1) {
                const V3 env = V3{prm.sky_radiance[0], prm.sky_radiance[1], prm.sky_radiance[2]} * 1.0f * throughput;
                radiance = radiance + V3{gl_min(env.x, 10.0f), gl_min(env.y, 10.0f), gl_min(env.z, 10.0f)};
                throughput = {0, 0, 0};
            } else {
                Surface sf = surface_at(sc, r0, r1, r2);
                if (dot(sf.emissive, V3{1, 1, 1}) > 0.0f && bounce == 0u) radiance = radiance + sf.emissive;
                V3 direct{0, 0, 0};
                light_surface(sf, prm);
                if (prm.light_count > 0) {
                    const uint32_t slot = slotOf[i];
                    // HitAnyTransparency's result; in an all-opaque scene the shadow batch ran as plain HitAny: 1 - hit
                    float visibility = 0.0f;
                    if (sf.NdotL > 0.0f && slot != kNoSlot)
                        visibility = shadowIsPlainAnyHit ? (__float_as_int(shadowRays[3 * size_t(slot) + 2].y) >= 0 ? 0.0f : 1.0f) : shadowRays[3 * size_t(slot) + 1].w;
                    const V3 reflectance = (eval_diffuse(sf) + eval_specular(sf)) * sf.opacity;
                    V3 rad = V3{prm.light_radiance[0], prm.light_radiance[1], prm.light_radiance[2]} * 1.0f;
                    rad = rad * visibility;
                    direct = reflectance * rad * sf.NdotL / 1.0f;
                }
                V3 rad = throughput * sf.opacity * direct;
                if (bounce > 0u) {
                    const float limit = 10.0f;
                    const float mx = gl_max(gl_max(rad.x, gl_max(rad.y, rad.z)), limit);
                    rad = rad * (limit / mx);
                }
                radiance = radiance + rad;
                // ---- EvaluateIndirectLight
                o = sf.P;
                float curSeed = seed;
                const float raySeed = float(id);
                float refractChance = gl_clamp(1.0f - sf.opacity, 0.1f, 0.9f);
                refractChance = sf.opacity == 1.0f ? 0.0f : refractChance;
                float rnd = saturate(random_seeded(raySeed, curSeed));
                V3 L{0, 0, 0}, refl{0, 0, 0};
                float pdf = 0.0f;
                bool refracted = false;
                if (rnd >= refractChance) {
                    rnd = random_seeded(raySeed, curSeed);
                    const V3 F = fresnel_schlick(sf.F0, sf.F90, sf.NdotV);
                    const float specChance = gl_clamp(dot(F, V3{0.33333f, 0.33333f, 0.33333f}), 0.1f, 0.9f);
                    const float u0 = random_seeded(raySeed, curSeed), u1 = random_seeded(raySeed, curSeed);
                    if (rnd < specChance) {   // SampleSpecularBRDF with SampleGGXVNDF
                        const float alpha = sqr(sf.roughness);
                        sf.V = normalize(sf.V);
                        const V3 N = normalize(sf.N);
                        const V3 up = fabsf(N.z) < 0.999f ? V3{0, 0, 1} : V3{1, 0, 0};
                        const V3 tangent = normalize(cross(up, N)), bitangent = normalize(cross(N, tangent));
                        V3 Vt = normalize(V3{dot(sf.V, tangent), dot(sf.V, bitangent), dot(sf.V, N)});
                        Vt = normalize(V3{Vt.x * alpha, Vt.y * alpha, Vt.z});
                        const float phi = kPi * 2.0f * u0;
                        float cx = cosf(phi), cy = sinf(phi);
                        const float cz = (1.0f - u1) * (1.0f + Vt.z) + -Vt.z;
                        const float sc2 = sqrtf(gl_clamp(1.0f - cz * cz, 0.0f, 1.0f));
                        cx *= sc2; cy *= sc2;
                        const V3 H{cx + Vt.x, cy + Vt.y, cz + Vt.z};
                        const V3 Hs{H.x * alpha, H.y * alpha, gl_max(H.z, 0.0f)};
                        const V3 Mv = normalize((tangent * Hs.x + bitangent * Hs.y) + N * Hs.z);
                        sf.L = Mv * (2.0f * dot(sf.V, Mv)) - sf.V;
                        update_surface(sf);
                        pdf = 1.0f;
                        if (sf.NdotL > 0.0f && sf.LdotH > 0.0f) {
                            const V3 F2 = fresnel_schlick(sf.F0, sf.F90, sf.LdotH);
                            const float Vis = vis_correlated(sf.NdotV, sf.NdotL, alpha), G1 = vis_separable(sf.NdotV, alpha);
                            L = sf.L;
                            pdf = G1 / (4.0f * fabsf(dot(sf.V, sf.N)));
                            refl = F2 * Vis;
                        }
                        refl = refl * sf.opacity;
                        pdf *= specChance;
                    } else {   // SampleDiffuseBRDF
                        const float theta = sqrtf(u0), phi = 2.0f * kPi * u1;
                        const V3 Ll{theta * cosf(phi), theta * sinf(phi), sqrtf(1.0f - u0)};
                        const V3 N = sf.N;
                        const V3 up = fabsf(N.z) < 0.999f ? V3{0, 0, 1} : V3{1, 0, 0};
                        const V3 tangent = normalize(cross(up, N)), bitangent = cross(N, tangent);
                        sf.L = normalize((tangent * Ll.x + bitangent * Ll.y) + N * Ll.z);
                        update_surface(sf);
                        L = sf.L;
                        pdf = sf.NdotL / kPi;
                        refl = eval_diffuse(sf);
                        refl = refl * ((1.0f - sf.metalness) * sf.opacity);
                        pdf *= (1.0f - specChance);
                    }
                    pdf *= (1.0f - refractChance);
                    o = o + sf.V * kEpsilon;
                } else {
                    L = d;
                    pdf = refractChance;
                    const float k = 1.0f - sf.opacity;
                    refl = {k, k, k};
                    sf.NdotL = 1.0f;
                    o = o - sf.N * kEpsilon;
                    refracted = true;
                }
                if (pdf > 0.0f && dot(refl, V3{1, 1, 1}) > 0.0f) throughput = throughput * (refl * sf.NdotL / pdf);
                else throughput = {0, 0, 0};
                d = normalize(L);
                throughput = throughput * sf.ao;
                float probability = gl_clamp(gl_max(throughput.x, gl_max(throughput.y, throughput.z)), 0.01f, 0.99f);
                probability = bounce < 3u ? gl_min(3.0f * probability, 1.0f) : probability;
                if (random_seeded(raySeed, curSeed) > probability) throughput = {0, 0, 0};
                else if (dot(d, sf.geometryNormal) <= 0.0f && !refracted) throughput = {0, 0, 0};
                else throughput = throughput / probability;
            }
            const float energy = dot(throughput, V3{1, 1, 1});
            if (energy == 0.0f || bounce == prm.max_bounces) {
                const uint32_t pixel = uint32_t(id) / prm.samples_per_frame;   // Flatten2D(pixel, resolution)
                uint32_t at = pixel;
                if (accumTileOrder) {   // index of the pixel in rayGen's storage order (a rank's pixels are then contiguous)
                    const uint32_t x = pixel % width, y = pixel / width;
                    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
                    const uint32_t gx = x / 8u, gy = y / 8u;
                    if (gx < perfX && gy < perfY) at = ((y & 7u) * 8u + (x & 7u)) + (gy * perfX + gx) * 64u;
                    else if (gx >= perfX && gy < perfY) at = y * overX + (x - perfX * 8u) + perfX * perfY * 64u;
                    else at = x * overY + (y - perfY * 8u) + perfX * perfY * 64u + overX * perfY * 8u;
                    if (accumTileOrder == 2u) {   // compact index inside this call's interleaved shard (the buffer holds only its pixels)
                        const unsigned long long slot0 = (unsigned long long)at * prm.samples_per_frame;
                        const unsigned long long blk = slot0 / shard.block;
                        at = uint32_t(((blk / shard.parts) * shard.block + slot0 % shard.block) / prm.samples_per_frame);
                    }
                }
                float* px = accum + 4 * size_t(at);
                atomicAdd(px + 0, radiance.x); atomicAdd(px + 1, radiance.y); atomicAdd(px + 2, radiance.z); atomicAdd(px + 3, 1.0f);
            } else {
                survive = true;
                n0 = make_float4(o.x, o.y, o.z, r0.w);
                n1 = make_float4(d.x, d.y, d.z, 0.0f);
                const uint32_t p0 = uint32_t(__half_as_ushort(__float2half_rn(radiance.x))) | (uint32_t(__half_as_ushort(__float2half_rn(radiance.y))) << 16);
                const uint32_t p1 = uint32_t(__half_as_ushort(__float2half_rn(throughput.x))) | (uint32_t(__half_as_ushort(__float2half_rn(throughput.y))) << 16);
                const uint32_t p2 = uint32_t(__half_as_ushort(__float2half_rn(radiance.z))) | (uint32_t(__half_as_ushort(__float2half_rn(throughput.z))) << 16);
                pay = make_float4(__uint_as_float(p0), __uint_as_float(p1), __uint_as_float(p2), 0.0f);
            }
        }
    }
    const unsigned m = __ballot_sync(kFullMask, survive);
    if (m) {
        const unsigned lane = threadIdx.x & 31u;
        const int leader = __ffs(m) - 1;
        unsigned base = 0;
        if (int(lane) == leader) base = atomicAdd(outCount, unsigned(__popc(m)));
        base = __shfl_sync(kFullMask, base, leader);
        if (survive) {
            const size_t dst = base + __popc(m & ((1u << lane) - 1u));
            raysOut[3 * dst] = n0; raysOut[3 * dst + 1] = n1; raysOut[3 * dst + 2] = make_float4(0.0f, __int_as_float(-1), 0.0f, 0.0f);
            payloadOut[dst] = pay;
        }
    }
}

// ------------------------------------------------------------------------------------------------ ray binning
// DetermineRayBin (raytracer/tracing.hsh:18-21): ivec2(UnitVectorToOctahedron(direction) * 8.0) flattened over 8 columns.
// A coordinate that saturates to exactly 1.0 gives 8, so bins run up to 8 * 8 + 8 = 72 (the shader has the same range).
constexpr uint32_t kBins = 80;
__device__ __forceinline__ uint32_t ray_bin(float x, float y, float z) {
    const float l1 = (fabsf(x) + fabsf(y)) + fabsf(z);
    x = x / l1; z = z / l1;
    if (y < 0.0f) {
        const float ox = x, oz = z;
        x = (ox >= 0.0f ? 1.0f : -1.0f) * (1.0f - fabsf(oz));
        z = (oz >= 0.0f ? 1.0f : -1.0f) * (1.0f - fabsf(ox));
    }
    const float cx = saturate(0.5f * x + 0.5f), cz = saturate(0.5f * z + 0.5f);
    const float fx = cx * 8.0f, fz = cz * 8.0f;
    const int ix = (fx != fx) ? int(0x80000000u) : __float2int_rz(fx), iz = (fz != fz) ? int(0x80000000u) : __float2int_rz(fz);
    return min(uint32_t(iz * 8 + ix), kBins - 1u);   // NaN directions land in the last bin
}

constexpr uint32_t kBinChunk = 2048;   // rays per CTA: per-chunk histograms make the scatter a STABLE counting sort

__global__ void __launch_bounds__(256)
bin_count(const float4* __restrict__ rays, uint32_t n, const uint32_t* __restrict__ countPtr, uint32_t* __restrict__ chunkHist /* [chunks][kBins] */) {
    chain_begin();
    const uint32_t count = batch_count(n, countPtr);
    __shared__ uint32_t h[kBins];
    if (threadIdx.x < kBins) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t first = blockIdx.x * kBinChunk;
    for (uint32_t k = threadIdx.x; k < kBinChunk; k += blockDim.x) {
        const uint32_t i = first + k;
        if (i < count) {
            const float4 d = rays[3 * size_t(i) + 1];
            atomicAdd(&h[ray_bin(d.x, d.y, d.z)], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < kBins) chunkHist[size_t(blockIdx.x) * kBins + threadIdx.x] = h[threadIdx.x];
}

// binningOffset.csh: exclusive offsets of the bins; here also per chunk (column-wise scan over the chunk histograms).
__global__ void bin_offsets(uint32_t* __restrict__ chunkHist, uint32_t chunks) {
    chain_begin();
    __shared__ uint32_t total[kBins];
    const uint32_t b = threadIdx.x;
    if (b < kBins) {
        uint32_t run = 0;
        for (uint32_t c = 0; c < chunks; c++) { const uint32_t v = chunkHist[size_t(c) * kBins + b]; chunkHist[size_t(c) * kBins + b] = run; run += v; }
        total[b] = run;
    }
    __syncthreads();
    if (b < kBins) {
        uint32_t off = 0;
        for (uint32_t k = 0; k < b; k++) off += total[k];
        for (uint32_t c = 0; c < chunks; c++) chunkHist[size_t(c) * kBins + b] += off;
    }
}

// binning.csh: move every ray (and its payload) to its bin's segment. Stable: within a bin rays keep their order (the
// shader's atomic order is arbitrary, so any order inside a bin is a valid result of the reference).
__global__ void __launch_bounds__(256)
bin_scatter(const float4* __restrict__ rays, const float4* __restrict__ payload, uint32_t n, const uint32_t* __restrict__ countPtr,
            const uint32_t* __restrict__ chunkHist, float4* __restrict__ raysOut, float4* __restrict__ payloadOut) {
    chain_begin();
    const uint32_t count = batch_count(n, countPtr);
    __shared__ uint32_t base[kBins];
    __shared__ uint8_t binOf[kBinChunk];
    if (threadIdx.x < kBins) base[threadIdx.x] = chunkHist[size_t(blockIdx.x) * kBins + threadIdx.x];
    const uint32_t first = blockIdx.x * kBinChunk;
    for (uint32_t k = threadIdx.x; k < kBinChunk; k += blockDim.x) {
        const uint32_t i = first + k;
        uint8_t b = 0xff;
        if (i < count) { const float4 d = rays[3 * size_t(i) + 1]; b = uint8_t(ray_bin(d.x, d.y, d.z)); }
        binOf[k] = b;
    }
    __syncthreads();
    // one warp per group of bins walks the chunk in order: rank of ray k inside its bin = number of earlier rays of that bin
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, warps = blockDim.x >> 5;
    for (uint32_t b = warp; b < kBins; b += warps) {
        uint32_t run = base[b];
        for (uint32_t k0 = 0; k0 < kBinChunk; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool mine = binOf[k] == b;
            const unsigned m = __ballot_sync(kFullMask, mine);
            if (mine) {
                const size_t dst = run + __popc(m & ((1u << lane) - 1u)), src = size_t(first) + k;
                raysOut[3 * dst] = rays[3 * src]; raysOut[3 * dst + 1] = rays[3 * src + 1]; raysOut[3 * dst + 2] = rays[3 * src + 2];
                if (payload) payloadOut[dst] = payload[src];
            }
            run += __popc(m);
        }
    }
}

__global__ void set_words(uint32_t* p, uint32_t a, uint32_t b, uint32_t c) {
    chain_begin();
    p[0] = a; p[1] = b; p[2] = c;
}
__global__ void next_bounce_counts(uint32_t* p) {   // survivors become the next batch; survivor and shadow counters restart
    chain_begin();
    p[0] = p[1]; p[1] = 0u; p[2] = 0u;
}

// dst += src (float4 records): the accumulation buffer of a second lane of sample passes joins the caller's
__global__ void __launch_bounds__(256) add_accum(float4* __restrict__ dst, const float4* __restrict__ src, unsigned long long n) {
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += gridDim.x * 256ull) {
        const float4 a = dst[i], b = src[i];
        dst[i] = make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    }
}

SceneTables tables_of(const atlas_rt_scene* scene) {
    return SceneTables{scene->instances, scene->triangles, scene->materials, scene->textures, scene->materialCount, scene->textureCount};
}

int enqueue_binning(atlas_rt_context* ctx, const float4* rays, const float4* payload, uint32_t n, const uint32_t* dCount, float4* raysOut,
                    float4* payloadOut, uint32_t* chunkHist, cudaStream_t st = nullptr, int chain = -1) {
    if (!st) st = ctx->stream;
    const uint32_t chunks = (n + kBinChunk - 1) / kBinChunk;
    const bool pdl = chain < 0 ? ctx->chainLaunch != 0 : chain != 0;
    ATLAS_CUDA(ctx, launch_chain(pdl, bin_count, chunks, 256, 0, st, rays, n, dCount, chunkHist));
    ATLAS_CUDA(ctx, launch_chain(pdl, bin_offsets, 1, 96, 0, st, chunkHist, chunks));
    ATLAS_CUDA(ctx, launch_chain(pdl, bin_scatter, chunks, 256, 0, st, rays, payload, n, dCount, static_cast<const uint32_t*>(chunkHist), raysOut, payloadOut));
    ctx->launches += 3;
    return ATLAS_RT_OK;
}

// One bounce on device buffers with the batch size on the device: closest hit in place (OPACITY_CHECK like
// PathTracingRenderer.cpp:186), shadow rays, any-hit with transparency over them, hit shader.
// dCounts: [0] = rays in this batch (read), [1] = survivors (accumulated), [2] = shadow rays (accumulated; both start at 0).
int enqueue_bounce(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_pt_params& prm, float seed, uint32_t bounce, float4* rays,
                   const float4* payloadIn, uint32_t n, uint32_t* dCounts, float4* raysOut, float4* payloadOut, float4* shadow, uint32_t* slotOf,
                   float* accum, uint32_t accumTileOrder, uint32_t width, uint32_t height, unsigned long long* dTraced, const SlotShard& shard,
                   cudaStream_t st = nullptr, int lane = 0, int chain = -1) {
    if (!st) st = ctx->stream;
    const bool pdl = chain < 0 ? ctx->chainLaunch != 0 : chain != 0;
    // OPACITY_CHECK traces (PathTracingRenderer.cpp:186, rayHit.csh:331). Where every triangle is fully opaque the
    // *Transparency variants accept exactly what the plain ones accept (and both closest-hit loops restore the ray the same
    // way), so the cheaper 48-byte kernels run instead, bit for bit the same; the shadow batch additionally needs every
    // instance to carry the shadow bit, because plain HitAny treats culled instances differently (bvh.hsh:387-390).
    const bool closestOpacity = !scene->allOpaque, shadowOpacity = !(scene->allOpaque && scene->allShadowBit);
    // (lanes: sample passes running side by side, each on its own stream with its own ray-queue head)
    int rc = launch_trace(ctx, scene, rays, rays, n, ATLAS_RT_MASK_ALL, 0.0f, ATLAS_RT_INF, false, false, false, lane == 0, closestOpacity, st, lane, dCounts, false, nullptr, nullptr, 0, nullptr, chain);
    if (rc != ATLAS_RT_OK) return rc;
    const SceneTables sc = tables_of(scene);
    const uint32_t grid = (n + 127) / 128;
    ATLAS_CUDA(ctx, launch_chain(pdl, shade_prepare, grid, 128, 0, st, static_cast<const float4*>(rays), n, static_cast<const uint32_t*>(dCounts), prm, sc,
                                 shadow, slotOf, dCounts + 2, dTraced));
    ctx->launches++;
    // HitAnyTransparency(ray, INSTANCE_MASK_SHADOW, 0.0, lightDistance - 2.0 * EPSILON), lightDistance = INF
    rc = launch_trace(ctx, scene, shadow, shadow, n, ATLAS_RT_MASK_SHADOW, 0.0f, ATLAS_RT_INF - 2.0f * kEpsilon, true, false, false, lane == 0, shadowOpacity, st, lane,
                      dCounts + 2, false, nullptr, nullptr, 0, nullptr, chain);
    if (rc != ATLAS_RT_OK) return rc;
    ATLAS_CUDA(ctx, launch_chain(pdl, shade_finish, grid, 128, 0, st, static_cast<const float4*>(rays), payloadIn, static_cast<const float4*>(shadow),
                                 static_cast<const uint32_t*>(slotOf), n, static_cast<const uint32_t*>(dCounts), prm, seed, bounce, sc, raysOut, payloadOut, accum,
                                 accumTileOrder, width, height, dCounts + 1, shadowOpacity ? 0 : 1, shard));
    ctx->launches++;
    return ATLAS_RT_OK;
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
    auto done = [&](int rc) { dev_free(ctx, tmp); dev_free(ctx, dJit); return rc; };
    cudaError_t e = cudaSuccess;
    if (!devOut) { e = dev_alloc(ctx, &tmp, count * 3); dOut = tmp; }
    if (e == cudaSuccess && jitter) {
        e = dev_alloc(ctx, &dJit, size_t(samples) * 2);
        if (e == cudaSuccess) e = cudaMemcpyAsync(dJit, jitter, size_t(samples) * 8, cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "primary rays: staging", e));
    raygen_kernel<<<uint32_t((count + 127) / 128), 128, 0, ctx->stream>>>(*camera, width, height, samples, dJit, 0.5f, 0.5f, SlotShard{0ull, count, 0u, 0u, 1u},
                                                                          count, dOut);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess && !devOut) e = copy_out(ctx, rays_out, dOut, count * 48, false);
    if (e == cudaSuccess && (!(flags & ATLAS_RT_ASYNC) || jitter || !devOut)) e = cudaStreamSynchronize(ctx->stream);
    return done(e == cudaSuccess ? ATLAS_RT_OK : fail(ctx, ATLAS_RT_ERR_CUDA, "primary rays", e));
}

void atlas_rt_sample_jitter(int32_t sample_count, float jitter_xy[2]) {
    // rayGen.csh:33-34: random(vec2(float(sampleCount), 0.0)), random(vec2(float(sampleCount), 1.0)) — integer hash, exact on the host
    if (!jitter_xy) return;
    for (int k = 0; k < 2; k++) {
        const float a = float(sample_count), b = float(k);
        uint32_t ua, ub;
        memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
        const uint32_t m = (hash1(ua ^ hash1(ub)) & 0x007FFFFFu) | 0x3F800000u;
        float f;
        memcpy(&f, &m, 4);
        jitter_xy[k] = f - 1.0f;
    }
}

int atlas_rt_bin_rays(atlas_rt_context* ctx, const void* rays_in, const void* payload_in, uint64_t count, void* rays_out, void* payload_out, uint32_t flags) {
    if (!ctx || (count && (!rays_in || !rays_out)) || (payload_in && !payload_out) || (count && rays_in == rays_out)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    if ((flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT)) != (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT))
        return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_bin_rays works on device-resident buffers");
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 rays");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return ATLAS_RT_OK;
    const uint32_t n = uint32_t(count);
    uint32_t* hist = nullptr;
    ATLAS_CUDA(ctx, dev_alloc(ctx, &hist, size_t((n + kBinChunk - 1) / kBinChunk) * kBins));
    int rc = enqueue_binning(ctx, static_cast<const float4*>(rays_in), static_cast<const float4*>(payload_in), n, nullptr, static_cast<float4*>(rays_out),
                             static_cast<float4*>(payload_out), hist);
    dev_free(ctx, hist);
    if (rc != ATLAS_RT_OK) return rc;
    if (!(flags & ATLAS_RT_ASYNC)) ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

int atlas_rt_pathtrace_bounce(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_pt_params* params, float seed, uint32_t bounce,
                              const void* rays_in, const void* payload_in, uint64_t count, void* rays_out, void* payload_out, float* accum,
                              uint32_t width, uint32_t height, uint64_t* out_count, uint32_t flags) {
    if (!ctx || !scene || scene->ctx->device != ctx->device || !params || !rays_out || !payload_out || !accum || !out_count || (count && !rays_in) ||
        !params->samples_per_frame)
        return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    if ((flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT)) != (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT))
        return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_pathtrace_bounce works on device-resident ray / payload / accumulation buffers");
    if (!scene->allShading) return fail(ctx, ATLAS_RT_ERR_INVALID, "the path tracer shades from the 96-byte triangles: call atlas_rt_mesh_pack_shading on every mesh");
    if (bounce > 0 && !payload_in) return fail(ctx, ATLAS_RT_ERR_INVALID, "payload_in required after the first bounce");
    if (rays_in == rays_out) return fail(ctx, ATLAS_RT_ERR_INVALID, "rays_out must not alias rays_in (survivors are compacted)");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    *out_count = 0;
    if (count == 0) return ATLAS_RT_OK;
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 rays");
    const uint32_t n = uint32_t(count);
    float4* shadow = nullptr;
    uint32_t *slotOf = nullptr, *dCounts = nullptr;
    auto done = [&](int rc) { dev_free(ctx, shadow); dev_free(ctx, slotOf); dev_free(ctx, dCounts); return rc; };
    cudaError_t e = dev_alloc(ctx, &shadow, size_t(n) * 3);
    if (e == cudaSuccess) e = dev_alloc(ctx, &slotOf, n);
    if (e == cudaSuccess) e = dev_alloc(ctx, &dCounts, 4);
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "bounce scratch", e));
    set_words<<<1, 1, 0, ctx->stream>>>(dCounts, n, 0u, 0u);
    ctx->launches++;
    int rc = enqueue_bounce(ctx, scene, *params, seed, bounce, const_cast<float4*>(static_cast<const float4*>(rays_in)), static_cast<const float4*>(payload_in), n,
                            dCounts, static_cast<float4*>(rays_out), static_cast<float4*>(payload_out), shadow, slotOf, accum,
                            (flags & ATLAS_RT_ACCUM_TILE_ORDER) ? 1u : 0u, width, height, nullptr, SlotShard{0ull, 0ull, 0u, 0u, 1u});
    if (rc != ATLAS_RT_OK) return done(rc);
    e = cudaMemcpyAsync(ctx->pinned, dCounts + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "read back survivor count", e));
    uint32_t c = 0;
    memcpy(&c, ctx->pinned, sizeof(c));
    *out_count = c;
    return done(ATLAS_RT_OK);
}

static int pathtrace_frames(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_camera* camera, uint32_t width, uint32_t height,
                            const atlas_rt_pt_params* params, uint32_t frames, int32_t first_sample_count, const float* seeds, const SlotShard& shard,
                            uint32_t accumMode, float* accum, uint64_t* rays_traced, uint32_t flags) {
    const uint64_t count = shard_count(shard);
    if (count > 0x7fffffffull) return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "more than 2^31-1 rays per frame");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (rays_traced) *rays_traced = 0;
    if (count == 0 || frames == 0) return ATLAS_RT_OK;
    const uint32_t n = uint32_t(count);
    const bool binning = (flags & ATLAS_RT_RAY_BINNING) != 0;
    // Sample passes are independent until they add into the image, and a pass ends in small, latency-bound launches (the
    // later bounces: a few hundred thousand rays, each launch as long as its longest ray). With more than one pass to do
    // they therefore run on up to ctx->ptLanes LANES — pass f on lane f % lanes, each lane its own stream, ray / payload /
    // shadow buffers, counters and ray-queue head — so that one pass's tail overlaps another's big first bounces. Lane 0
    // accumulates into the caller's buffer, the others into zeroed buffers of their own that are added to it in lane order
    // at the end: the result is deterministic (it differs from the one-lane result only in the order of the float adds).
    constexpr int kMaxLanes = 8;
    const int lanes = int(std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(frames, uint32_t(ctx->ptLanes)), kMaxLanes)));
    const unsigned long long accumRecords = accumMode == 2u ? count / params->samples_per_frame : (unsigned long long)width * height;
    struct Lane {
        cudaStream_t st = nullptr;
        float4 *rays[3] = {nullptr, nullptr, nullptr}, *pay[3] = {nullptr, nullptr, nullptr}, *shadow = nullptr;
        uint32_t *slotOf = nullptr, *dCounts = nullptr, *hist = nullptr;
        float* accum = nullptr;
    } L[kMaxLanes];
    unsigned long long* dTraced = nullptr;
    auto done = [&](int rc) {
        if (rc != ATLAS_RT_OK) for (int l = 1; l < lanes; l++) if (L[l].st) cudaStreamSynchronize(L[l].st);   // nothing may still use what is freed below
        for (int l = 0; l < lanes; l++) {
            for (int k = 0; k < 3; k++) { dev_free(ctx, L[l].rays[k]); dev_free(ctx, L[l].pay[k]); }
            dev_free(ctx, L[l].shadow); dev_free(ctx, L[l].slotOf); dev_free(ctx, L[l].dCounts); dev_free(ctx, L[l].hist);
            if (l > 0) dev_free(ctx, L[l].accum);
        }
        dev_free(ctx, dTraced);
        return rc;
    };
    cudaError_t e = cudaSuccess;
    int usable = 1;
    for (int l = 1; l < lanes; l++) if (ctx->computeExtra[l - 1]) usable = l + 1; else break;
    const int nl = std::min(lanes, usable);
    for (int l = 0; l < nl && e == cudaSuccess; l++) {
        L[l].st = l == 0 ? ctx->stream : ctx->computeExtra[l - 1];
        for (int k = 0; k < (binning ? 3 : 2) && e == cudaSuccess; k++) { e = dev_alloc(ctx, &L[l].rays[k], size_t(n) * 3); if (e == cudaSuccess) e = dev_alloc(ctx, &L[l].pay[k], n); }
        if (e == cudaSuccess) e = dev_alloc(ctx, &L[l].shadow, size_t(n) * 3);
        if (e == cudaSuccess) e = dev_alloc(ctx, &L[l].slotOf, n);
        if (e == cudaSuccess) e = dev_alloc(ctx, &L[l].dCounts, 4);
        if (e == cudaSuccess && binning) e = dev_alloc(ctx, &L[l].hist, size_t((n + kBinChunk - 1) / kBinChunk) * kBins);
        if (l == 0) L[l].accum = accum;
        else {
            if (e == cudaSuccess) e = dev_alloc(ctx, &L[l].accum, size_t(accumRecords) * 4);
            if (e == cudaSuccess) e = cudaMemsetAsync(L[l].accum, 0, size_t(accumRecords) * 16, ctx->stream);
        }
    }
    if (e == cudaSuccess) e = dev_alloc(ctx, &dTraced, 1);
    if (e == cudaSuccess) e = cudaMemsetAsync(dTraced, 0, sizeof(unsigned long long), ctx->stream);
    // the other lanes start behind everything the context stream has done so far (the caller's inputs, the buffers above)
    cudaEvent_t* ev = ctx->pipeEvents;
    if (e == cudaSuccess && nl > 1) e = cudaEventRecord(ev[0], ctx->stream);
    for (int l = 1; l < nl && e == cudaSuccess; l++) e = cudaStreamWaitEvent(L[l].st, ev[0], 0);
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "path tracer buffers", e));
    // With several lanes the launches are NOT chained as programmatic dependent launches: a chained kernel's CTAs become
    // resident as soon as slots free up and wait there for their predecessor, i.e. the tail of one lane's trace kernel would fill
    // with the waiting CTAs of the same lane's next kernel instead of the runnable CTAs of another lane (measured on a 1/8 shard
    // of C5: 4 lanes 3.55 ms per pass chained, 2.10 ms unchained; one lane 3.94 / 4.10 ms).
    const int chain = nl > 1 ? 0 : -1;
    const bool pdl = nl > 1 ? false : ctx->chainLaunch != 0;
    const uint32_t bounces = params->max_bounces;
    for (uint32_t f = 0; f < frames; f++) {
        const int l = int(f % uint32_t(nl));
        Lane& ln = L[l];
        float jit[2];
        atlas_rt_sample_jitter(first_sample_count + int32_t(f), jit);
        int cur = 0;
        e = launch_chain(pdl, raygen_kernel, (n + 127) / 128, 128, 0, ln.st, *camera, width, height, params->samples_per_frame,
                         static_cast<const float*>(nullptr), jit[0], jit[1], shard, (unsigned long long)count, ln.rays[cur]);
        ctx->launches++;
        if (e == cudaSuccess) e = launch_chain(pdl, set_words, 1, 1, 0, ln.st, ln.dCounts, n, 0u, 0u);
        ctx->launches++;
        if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "raygen", e));
        for (uint32_t b = 0; b <= bounces; b++) {
            if (binning && b > 0) {   // RayTracingHelper.cpp:304-344 (dormant in the reference): order the rays by direction bin
                const int rc = enqueue_binning(ctx, ln.rays[cur], ln.pay[cur], n, ln.dCounts, ln.rays[2], ln.pay[2], ln.hist, ln.st, chain);
                if (rc != ATLAS_RT_OK) return done(rc);
                std::swap(ln.rays[cur], ln.rays[2]);
                std::swap(ln.pay[cur], ln.pay[2]);
            }
            const int rc = enqueue_bounce(ctx, scene, *params, seeds[size_t(f) * (bounces + 1) + b], b, ln.rays[cur], ln.pay[cur], n, ln.dCounts, ln.rays[cur ^ 1],
                                          ln.pay[cur ^ 1], ln.shadow, ln.slotOf, ln.accum, accumMode, width, height, dTraced, shard, ln.st, l, chain);
            if (rc != ATLAS_RT_OK) return done(rc);
            e = launch_chain(pdl, next_bounce_counts, 1, 1, 0, ln.st, ln.dCounts);
            ctx->launches++;
            if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "counters", e));
            cur ^= 1;
        }
    }
    // join: the context stream continues behind every lane, then the lanes' images are added in lane order
    for (int l = 1; l < nl && e == cudaSuccess; l++) {
        e = cudaEventRecord(ev[l], L[l].st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev[l], 0);
    }
    for (int l = 1; l < nl && e == cudaSuccess; l++) {
        add_accum<<<uint32_t(std::min<unsigned long long>((accumRecords + 255) / 256, 148ull * 8)), 256, 0, ctx->stream>>>(reinterpret_cast<float4*>(accum),
                                                                                                     reinterpret_cast<const float4*>(L[l].accum), accumRecords);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "path tracer lanes", e));
    if (rays_traced || !(flags & ATLAS_RT_ASYNC)) {
        if (rays_traced) e = cudaMemcpyAsync(ctx->pinned, dTraced, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "path tracer", e));
        if (rays_traced) memcpy(rays_traced, ctx->pinned, sizeof(uint64_t));
    }
    return done(ATLAS_RT_OK);
}

int atlas_rt_pathtrace_bounces(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_camera* camera, uint32_t width, uint32_t height,
                               const atlas_rt_pt_params* params, uint32_t frames, int32_t first_sample_count, const float* seeds, uint64_t slot_begin,
                               uint64_t slot_end, float* accum, uint64_t* rays_traced, uint32_t flags) {
    if (!ctx || !scene || scene->ctx->device != ctx->device || !camera || !params || !seeds || !accum || !width || !height || !params->samples_per_frame)
        return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    if (!scene->allShading) return fail(ctx, ATLAS_RT_ERR_INVALID, "the path tracer shades from the 96-byte triangles: call atlas_rt_mesh_pack_shading on every mesh");
    const uint64_t total = uint64_t(width) * height * params->samples_per_frame;
    if (slot_end == 0) slot_end = total;
    if (slot_begin > slot_end || slot_end > total) return fail(ctx, ATLAS_RT_ERR_INVALID, "slot range outside the frame");
    return pathtrace_frames(ctx, scene, camera, width, height, params, frames, first_sample_count, seeds, SlotShard{slot_begin, slot_end, 0u, 0u, 1u},
                            (flags & ATLAS_RT_ACCUM_TILE_ORDER) ? 1u : 0u, accum, rays_traced, flags);
}

int atlas_rt_pathtrace_bounces_interleaved(atlas_rt_context* ctx, const atlas_rt_scene* scene, const atlas_rt_camera* camera, uint32_t width, uint32_t height,
                                           const atlas_rt_pt_params* params, uint32_t frames, int32_t first_sample_count, const float* seeds, uint32_t part,
                                           uint32_t parts, uint32_t block_pixels, float* accum_local, uint64_t* local_pixels, uint64_t* rays_traced,
                                           uint32_t flags) {
    if (!ctx || !scene || scene->ctx->device != ctx->device || !camera || !params || !seeds || !width || !height || !params->samples_per_frame || !parts ||
        part >= parts || !block_pixels || (block_pixels % 64u) != 0u)
        return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument (block_pixels must be a multiple of 64: whole rayGen tiles)");
    if (!scene->allShading) return fail(ctx, ATLAS_RT_ERR_INVALID, "the path tracer shades from the 96-byte triangles: call atlas_rt_mesh_pack_shading on every mesh");
    const uint64_t total = uint64_t(width) * height * params->samples_per_frame;
    const SlotShard shard{0ull, total, part, parts, block_pixels * params->samples_per_frame};
    if (local_pixels) *local_pixels = shard_count(shard) / params->samples_per_frame;
    if (!accum_local) return ATLAS_RT_OK;   // size query
    return pathtrace_frames(ctx, scene, camera, width, height, params, frames, first_sample_count, seeds, shard, 2u, accum_local, rays_traced, flags);
}

// gathered: the shards' compact accumulation buffers one after the other (shard p first); image: y * width + x.
static __global__ void assemble_image(const float4* __restrict__ gathered, uint32_t width, uint32_t height, uint32_t parts, uint32_t blockPixels,
                               float4* __restrict__ image) {
    const uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= width * height) return;
    const uint32_t x = pixel % width, y = pixel / width;
    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
    const uint32_t gx = x / 8u, gy = y / 8u;
    uint32_t at;
    if (gx < perfX && gy < perfY) at = ((y & 7u) * 8u + (x & 7u)) + (gy * perfX + gx) * 64u;
    else if (gx >= perfX && gy < perfY) at = y * overX + (x - perfX * 8u) + perfX * perfY * 64u;
    else at = x * overY + (y - perfY * 8u) + perfX * perfY * 64u + overX * perfY * 8u;
    const uint32_t blk = at / blockPixels, part = blk % parts;
    // pixels of the shards before `part`
    unsigned long long base = 0;
    const SlotShard all{0ull, (unsigned long long)width * height, 0u, parts, blockPixels};
    for (uint32_t p = 0; p < part; p++) { SlotShard sh = all; sh.part = p; base += shard_count(sh); }
    image[pixel] = gathered[base + (unsigned long long)(blk / parts) * blockPixels + at % blockPixels];
}

int atlas_rt_image_from_shards(atlas_rt_context* ctx, const float* gathered, uint32_t width, uint32_t height, uint32_t parts, uint32_t block_pixels,
                               float* image, uint32_t flags) {
    if (!ctx || !gathered || !image || !width || !height || !parts || !block_pixels) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    if ((flags & (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT)) != (ATLAS_RT_DEVICE_INPUT | ATLAS_RT_DEVICE_OUTPUT))
        return fail(ctx, ATLAS_RT_ERR_UNSUPPORTED, "atlas_rt_image_from_shards works on device-resident buffers");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = width * height;
    assemble_image<<<(n + 255) / 256, 256, 0, ctx->stream>>>(reinterpret_cast<const float4*>(gathered), width, height, parts, block_pixels, reinterpret_cast<float4*>(image));
    ATLAS_LAUNCH_CHECK(ctx);
    if (!(flags & ATLAS_RT_ASYNC)) ATLAS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ATLAS_RT_OK;
}

}
