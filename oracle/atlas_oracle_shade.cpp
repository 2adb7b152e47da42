// atlas_oracle_shade.cpp — CPU restatement (plain C++, no FMA: -ffp-contract=off) of the reference code AROUND the
// traversal: the packed shading words of GPUTriangle, textured opacity, ray binning and the path tracer's bounce.
// TEST INFRASTRUCTURE ONLY: linked into oracle/libatlas_oracle.so, used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg as the checker — never by the product.
//
// Every function cites the reference lines it follows (paths relative to /root/reference). Third-party arithmetic:
// glm 0.9.8.0 (vcpkg.json:45-48; header-only, NOT under /root/reference) — normalize / dot / cross / packHalf2x16 /
// packUnorm4x8 are restated from glm's published source; the GLSL built-ins (unpackHalf2x16, unpackUnorm4x8, normalize,
// ...) from the GLSL 4.60 specification. Nothing here is pinned by a golden vector of the reference (it has none for these
// functions, and neither its shaders nor MeshData::BuildBVH can run in this container: no Vulkan device): PARITY UNPINNED
// for this file; the CUDA kernels are compared against it.
#include <cmath>
#include <cstdint>
#include <cstring>

namespace {

struct V3 { float x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }                       // glm: tmp.x + tmp.y + tmp.z
inline V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }   // glm::cross
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }                             // glm: x * inversesqrt(dot(x, x))
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }

// float -> int32 as x86-64 does it (cvttss2si): NaN and out-of-range give 0x80000000. The C++ conversion is undefined
// there, but the reference's binaries run on x86 and real meshes hit it (NaN tangents from degenerate uvs).
inline int32_t cvt_x86(float x) {
    if (x != x || x >= 2147483648.0f || x < -2147483648.0f) return int32_t(0x80000000u);
    return int32_t(x);
}

// Common::Packing::PackSignedVector3x10_1x2 — src/engine/common/Packing.cpp:24-35.
inline uint32_t pack_signed_3x10_1x2(float x, float y, float z, float w) {
    uint32_t packed = 0;
    packed |= uint32_t(cvt_x86((x * 0.5f + 0.5f) * 1023.0f)) << 0;
    packed |= uint32_t(cvt_x86((y * 0.5f + 0.5f) * 1023.0f)) << 10;
    packed |= uint32_t(cvt_x86((z * 0.5f + 0.5f) * 1023.0f)) << 20;
    packed |= uint32_t(cvt_x86((w * 0.5f + 0.5f) * 2.0f)) << 30;
    return packed;
}

// glm::detail::toFloat16 — glm 0.9.8 detail/type_half.inl.
inline uint32_t to_float16(float f) {
    int32_t i;
    std::memcpy(&i, &f, 4);
    const int s = (i >> 16) & 0x00008000;
    int e = ((i >> 23) & 0x000000ff) - (127 - 15);
    int m = i & 0x007fffff;
    if (e <= 0) {
        if (e < -10) return uint32_t(s);
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x00001000) m += 0x00002000;
        return uint32_t(s | (m >> 13));
    } else if (e == 0xff - (127 - 15)) {
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
inline uint32_t pack_half2x16(float x, float y) { return to_float16(x) | (to_float16(y) << 16); }

// glm::packUnorm4x8 — glm 0.9.8 detail/func_packing.inl: u8vec4(round(clamp(v, 0, 1) * 255)).
inline uint32_t unorm8(float c) { return uint32_t(cvt_x86(std::round(gclamp(c, 0.0f, 1.0f) * 255.0f))) & 0xffu; }
inline uint32_t pack_unorm4x8(const float* c) { return unorm8(c[0]) | (unorm8(c[1]) << 8) | (unorm8(c[2]) << 16) | (unorm8(c[3]) << 24); }

}   // namespace

extern "C" {

// MeshData::BuildBVH, second loop — src/engine/mesh/MeshData.cpp:176-228. Inputs per SOURCE triangle (as the first loop,
// :102-133, leaves them): tris n x 9, normals n x 9 (may be null = 0), uvs n x 6 (null = 0), colors n x 12 (null = 1).
// out: n x 11 words pn0 pn1 pn2 puv0 puv1 puv2 pt pbt pc0 pc1 pc2.
void oracle_pack_shading_words(const float* tris, const float* normals, const float* uvs, const float* colors, uint64_t n, uint32_t* out) {
    for (uint64_t k = 0; k < n; k++) {
        const float* t = tris + 9 * k;
        const V3 v0{t[0], t[1], t[2]}, v1{t[3], t[4], t[5]}, v2{t[6], t[7], t[8]};
        V3 n0{0, 0, 0}, n1 = n0, n2 = n0;
        if (normals) { const float* q = normals + 9 * k; n0 = {q[0], q[1], q[2]}; n1 = {q[3], q[4], q[5]}; n2 = {q[6], q[7], q[8]}; }
        float uv[6] = {0, 0, 0, 0, 0, 0};
        if (uvs) std::memcpy(uv, uvs + 6 * k, 24);
        float col[12];
        for (int a = 0; a < 12; a++) col[a] = colors ? colors[12 * k + a] : 1.0f;
        const V3 v0v1 = v1 - v0, v0v2 = v2 - v0;                                           // :181-182
        const float u01x = uv[2] - uv[0], u01y = uv[3] - uv[1], u02x = uv[4] - uv[0], u02y = uv[5] - uv[1];   // :184-185
        const float r = 1.0f / (u01x * u02y - u02x * u01y);                               // :187
        const V3 s = V3{u02y * v0v1.x - u01y * v0v2.x, u02y * v0v1.y - u01y * v0v2.y, u02y * v0v1.z - u01y * v0v2.z} * r;   // :189-191
        const V3 tt = V3{u01x * v0v2.x - u02x * v0v1.x, u01x * v0v2.y - u02x * v0v1.y, u01x * v0v2.z - u02x * v0v1.z} * r;  // :193-195
        const V3 normal = normalize((n0 + n1) + n2);                                       // :197
        const V3 tangent = normalize(s - normal * dot(normal, s));                         // :199
        const float handedness = dot(cross(tangent, normal), tt) < 0.0f ? 1.0f : -1.0f;    // :200
        const V3 bitangent = normalize(cross(tangent, normal)) * handedness;               // :202
        uint32_t* o = out + 11 * k;
        o[0] = pack_signed_3x10_1x2(n0.x, n0.y, n0.z, 0.0f);                               // :205-207
        o[1] = pack_signed_3x10_1x2(n1.x, n1.y, n1.z, 0.0f);
        o[2] = pack_signed_3x10_1x2(n2.x, n2.y, n2.z, 0.0f);
        o[3] = pack_half2x16(uv[0], uv[1]);                                                // :212-214
        o[4] = pack_half2x16(uv[2], uv[3]);
        o[5] = pack_half2x16(uv[4], uv[5]);
        o[6] = pack_signed_3x10_1x2(tangent.x, tangent.y, tangent.z, 0.0f);                // :209-210
        o[7] = pack_signed_3x10_1x2(bitangent.x, bitangent.y, bitangent.z, 0.0f);
        o[8] = pack_unorm4x8(col);                                                         // :216-218
        o[9] = pack_unorm4x8(col + 4);
        o[10] = pack_unorm4x8(col + 8);
    }
}

}   // extern "C"
