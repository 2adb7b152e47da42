// shading.cu — the packed shading words of GPUTriangle, computed on the device.
//
// Restates the second loop of Atlas::Mesh::MeshData::BuildBVH, src/engine/mesh/MeshData.cpp:176-228: per triangle the
// tangent frame from positions + texture coordinates, then
//   pn0..2, pt, pbt   Common::Packing::PackSignedVector3x10_1x2(vec4(v, 0))       src/engine/common/Packing.cpp:24-35
//   puv0..2           glm::packHalf2x16                                           glm 0.9.8 gtc/packing + detail/type_half.inl
//   pc0..2            glm::packUnorm4x8 = round(clamp(c, 0, 1) * 255)             glm 0.9.8 func_packing.inl
// glm (pinned 0.9.8.0 in vcpkg.json, not vendored under /root/reference) is restated from its published algorithm: parity
// of these words is "unpinned" by a reference-made golden and is checked in tests/ against an independent CPU restatement
// of the same lines. Plain IEEE fp32, no FMA, glm's operation order.
//
// The reference runs on x86-64, where float -> int32 conversion of NaN or out-of-range values (cvttss2si) yields
// 0x80000000; degenerate texture coordinates make the tangent NaN routinely (r = 1 / 0), so that behaviour is part of the
// words a real mesh gets and is reproduced here (CUDA's own conversion would give 0 / saturate).
#include "common.cuh"

namespace atlas {
namespace {

__device__ __forceinline__ int32_t cvt_x86(float x) {
    return (x != x || x >= 2147483648.0f || x < -2147483648.0f) ? int32_t(0x80000000u) : __float2int_rz(x);
}

// Common::Packing::PackSignedVector3x10_1x2 — Packing.cpp:24-35 (int32 shifts wrap like x86 shl).
__device__ __forceinline__ uint32_t pack_signed_3x10_1x2(float x, float y, float z, float w) {
    uint32_t packed = 0;
    packed |= uint32_t(cvt_x86(__fmul_rn(__fadd_rn(__fmul_rn(x, 0.5f), 0.5f), 1023.0f))) << 0;
    packed |= uint32_t(cvt_x86(__fmul_rn(__fadd_rn(__fmul_rn(y, 0.5f), 0.5f), 1023.0f))) << 10;
    packed |= uint32_t(cvt_x86(__fmul_rn(__fadd_rn(__fmul_rn(z, 0.5f), 0.5f), 1023.0f))) << 20;
    packed |= uint32_t(cvt_x86(__fmul_rn(__fadd_rn(__fmul_rn(w, 0.5f), 0.5f), 2.0f))) << 30;
    return packed;
}

// glm::detail::toFloat16 (glm 0.9.8 detail/type_half.inl): round-half-up on the 13 dropped mantissa bits.
__device__ __forceinline__ uint32_t to_float16(float f) {
    const int i = __float_as_int(f);
    const int s = (i >> 16) & 0x00008000;
    int e = ((i >> 23) & 0x000000ff) - (127 - 15);
    int m = i & 0x007fffff;
    if (e <= 0) {
        if (e < -10) return uint32_t(s);
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x00001000) m += 0x00002000;
        return uint32_t(s | (m >> 13));
    }
    if (e == 0xff - (127 - 15)) {
        if (m == 0) return uint32_t(s | 0x7c00);
        m >>= 13;
        return uint32_t(s | 0x7c00 | m | (m == 0));
    }
    if (m & 0x00001000) {
        m += 0x00002000;
        if (m & 0x00800000) { m = 0; e += 1; }
    }
    if (e > 30) return uint32_t(s | 0x7c00);
    return uint32_t(s | (e << 10) | (m >> 13));
}
__device__ __forceinline__ uint32_t pack_half2x16(float x, float y) { return to_float16(x) | (to_float16(y) << 16); }

// glm::packUnorm4x8: u8vec4(round(clamp(v, 0, 1) * 255)); round = std::round (half away from zero); uint8 from float
// truncates, NaN -> cvttss2si -> low byte 0.
__device__ __forceinline__ uint32_t unorm8(float c) {
    const float cl = gl_clamp(c, 0.0f, 1.0f);                  // glm::clamp = min(max(x, lo), hi)
    const float r = roundf(__fmul_rn(cl, 255.0f));
    return uint32_t(cvt_x86(r)) & 0xffu;
}
__device__ __forceinline__ uint32_t pack_unorm4x8(const float* c) { return unorm8(c[0]) | (unorm8(c[1]) << 8) | (unorm8(c[2]) << 16) | (unorm8(c[3]) << 24); }

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 mul(V3 a, float s) { return {__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return __fadd_rn(__fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fmul_rn(a.z, b.z)); }
__device__ __forceinline__ V3 cross(V3 x, V3 y) {   // glm::cross
    return {__fsub_rn(__fmul_rn(x.y, y.z), __fmul_rn(y.y, x.z)), __fsub_rn(__fmul_rn(x.z, y.x), __fmul_rn(y.z, x.x)),
            __fsub_rn(__fmul_rn(x.x, y.y), __fmul_rn(y.x, x.y))};
}
__device__ __forceinline__ V3 normalize(V3 v) { return mul(v, __fdiv_rn(1.0f, __fsqrt_rn(dot(v, v)))); }   // v * inversesqrt(dot(v, v))

__global__ void pack_shading_words_kernel(const float* __restrict__ tris, const float* __restrict__ normals, const float* __restrict__ uvs,
                                          const float* __restrict__ colors, uint32_t* __restrict__ out, uint64_t count) {
    const uint64_t k = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    if (k >= count) return;
    const float* t = tris + 9 * k;
    const V3 v0{t[0], t[1], t[2]}, v1{t[3], t[4], t[5]}, v2{t[6], t[7], t[8]};
    V3 n0{0, 0, 0}, n1 = n0, n2 = n0;
    if (normals) { const float* n = normals + 9 * k; n0 = {n[0], n[1], n[2]}; n1 = {n[3], n[4], n[5]}; n2 = {n[6], n[7], n[8]}; }
    float uv[6] = {0, 0, 0, 0, 0, 0};
    if (uvs) for (int a = 0; a < 6; a++) uv[a] = uvs[6 * k + a];
    float col[12];
    for (int a = 0; a < 12; a++) col[a] = colors ? colors[12 * k + a] : 1.0f;

    const V3 v0v1 = sub(v1, v0), v0v2 = sub(v2, v0);
    const float u01x = __fsub_rn(uv[2], uv[0]), u01y = __fsub_rn(uv[3], uv[1]);
    const float u02x = __fsub_rn(uv[4], uv[0]), u02y = __fsub_rn(uv[5], uv[1]);
    const float r = __fdiv_rn(1.0f, __fsub_rn(__fmul_rn(u01x, u02y), __fmul_rn(u02x, u01y)));
    const V3 s = mul(V3{__fsub_rn(__fmul_rn(u02y, v0v1.x), __fmul_rn(u01y, v0v2.x)), __fsub_rn(__fmul_rn(u02y, v0v1.y), __fmul_rn(u01y, v0v2.y)),
                        __fsub_rn(__fmul_rn(u02y, v0v1.z), __fmul_rn(u01y, v0v2.z))}, r);
    const V3 tt = mul(V3{__fsub_rn(__fmul_rn(u01x, v0v2.x), __fmul_rn(u02x, v0v1.x)), __fsub_rn(__fmul_rn(u01x, v0v2.y), __fmul_rn(u02x, v0v1.y)),
                         __fsub_rn(__fmul_rn(u01x, v0v2.z), __fmul_rn(u02x, v0v1.z))}, r);
    const V3 nsum{__fadd_rn(__fadd_rn(n0.x, n1.x), n2.x), __fadd_rn(__fadd_rn(n0.y, n1.y), n2.y), __fadd_rn(__fadd_rn(n0.z, n1.z), n2.z)};
    const V3 normal = normalize(nsum);
    const V3 tangent = normalize(sub(s, mul(normal, dot(normal, s))));
    const float handedness = dot(cross(tangent, normal), tt) < 0.0f ? 1.0f : -1.0f;
    const V3 bitangent = mul(normalize(cross(tangent, normal)), handedness);   // handedness * normalize(..): commutative

    uint32_t* o = out + 11 * k;
    o[0] = pack_signed_3x10_1x2(n0.x, n0.y, n0.z, 0.0f);
    o[1] = pack_signed_3x10_1x2(n1.x, n1.y, n1.z, 0.0f);
    o[2] = pack_signed_3x10_1x2(n2.x, n2.y, n2.z, 0.0f);
    o[3] = pack_half2x16(uv[0], uv[1]);
    o[4] = pack_half2x16(uv[2], uv[3]);
    o[5] = pack_half2x16(uv[4], uv[5]);
    o[6] = pack_signed_3x10_1x2(tangent.x, tangent.y, tangent.z, 0.0f);
    o[7] = pack_signed_3x10_1x2(bitangent.x, bitangent.y, bitangent.z, 0.0f);
    o[8] = pack_unorm4x8(col);
    o[9] = pack_unorm4x8(col + 4);
    o[10] = pack_unorm4x8(col + 8);
}

}   // namespace
}   // namespace atlas

using namespace atlas;

extern "C" int atlas_rt_pack_shading_words(atlas_rt_context* ctx, const float* tris, const float* normals9, const float* uvs6, const float* colors12,
                                           uint64_t count, uint32_t* payload11, uint32_t flags) {
    if (!ctx || !payload11 || (count && !tris)) return fail(ctx, ATLAS_RT_ERR_INVALID, "bad argument");
    ATLAS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return ATLAS_RT_OK;
    const bool devIn = flags & ATLAS_RT_DEVICE_INPUT, devOut = flags & ATLAS_RT_DEVICE_OUTPUT;
    float *dT = nullptr, *dN = nullptr, *dU = nullptr, *dC = nullptr;
    uint32_t* dOut = nullptr;
    auto done = [&](int rc) { dev_free(ctx, dT); dev_free(ctx, dN); dev_free(ctx, dU); dev_free(ctx, dC); dev_free(ctx, dOut); return rc; };
    auto stage = [&](const float* src, size_t floats, float** slot) -> cudaError_t {
        if (!src || devIn) return cudaSuccess;
        cudaError_t e = dev_alloc(ctx, slot, floats);
        if (e == cudaSuccess) e = copy_in(ctx, *slot, src, floats * 4, false);
        return e;
    };
    cudaError_t e = stage(tris, count * 9, &dT);
    if (e == cudaSuccess) e = stage(normals9, count * 9, &dN);
    if (e == cudaSuccess) e = stage(uvs6, count * 6, &dU);
    if (e == cudaSuccess) e = stage(colors12, count * 12, &dC);
    if (e == cudaSuccess && !devOut) e = dev_alloc(ctx, &dOut, count * 11);
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "staging", e));
    pack_shading_words_kernel<<<uint32_t((count + 127) / 128), 128, 0, ctx->stream>>>(devIn ? tris : dT, devIn ? normals9 : dN, devIn ? uvs6 : dU,
                                                                                      devIn ? colors12 : dC, devOut ? payload11 : dOut, count);
    ctx->launches++;
    e = cudaGetLastError();
    if (e == cudaSuccess && !devOut) e = copy_out(ctx, payload11, dOut, count * 44, false);
    if (e != cudaSuccess) return done(fail(ctx, ATLAS_RT_ERR_CUDA, "pack_shading_words", e));
    if (!(flags & ATLAS_RT_ASYNC) || !devOut || !devIn) e = cudaStreamSynchronize(ctx->stream);
    return done(e == cudaSuccess ? ATLAS_RT_OK : fail(ctx, ATLAS_RT_ERR_CUDA, "pack_shading_words", e));
}
