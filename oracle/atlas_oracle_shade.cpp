// atlas_oracle_shade.cpp — CPU restatement (plain C++, no FMA: -ffp-contract=off) of the reference code AROUND the
// traversal: the packed shading words of GPUTriangle, textured opacity, ray binning and the path tracer's bounce.
// TEST INFRASTRUCTURE ONLY: linked into oracle/libatlas_oracle.so, used by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg as the checker — never by the product.
//
// Every function cites the reference lines it follows (paths relative to /root/reference). Third-party arithmetic:
// glm 0.9.8.0 (vcpkg.json:45-48; header-only, NOT under /root/reference) — normalize / dot / cross / packHalf2x16 /
// packUnorm4x8 are restated from glm's published source; the GLSL built-ins (unpackHalf2x16, unpackUnorm4x8, normalize,
// ...) from the GLSL 4.60 specification. Pinned against the reference's own compiled code: pack_signed_3x10_1x2
// (Common::Packing::PackSignedVector3x10_1x2 in oracle/_ref, tests/test_oracle_vs_ref.py). Nothing else here is pinned by a
// golden vector of the reference (it has none for these functions, and neither its shaders nor MeshData::BuildBVH can run in
// this container: no Vulkan device): PARITY UNPINNED for the rest of this file; the CUDA kernels are compared against it.
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

// The restatement of Common::Packing::PackSignedVector3x10_1x2 by itself, so that tests/test_oracle_vs_ref.py can pin it
// against the reference's compiled function (oracle/_ref: common/Packing.cpp).
void oracle_pack_signed_3x10_1x2(const float* vec4s, uint64_t n, int32_t* out) {
    for (uint64_t i = 0; i < n; i++) out[i] = int32_t(pack_signed_3x10_1x2(vec4s[4 * i], vec4s[4 * i + 1], vec4s[4 * i + 2], vec4s[4 * i + 3]));
}

}   // extern "C"


// =================================================================================================================
// Path tracer: pathtracer/rayGen.csh, pathtracer/rayHit.csh and what they call. GLSL built-ins as the specification
// defines them: mix(x, y, a) = x * (1 - a) + y * a, clamp = min(max(x, lo), hi), normalize(v) = v * (1 / sqrt(dot(v, v)))
// (the spec leaves the rounding of normalize / inversesqrt / pow / sin / cos to the implementation: float terms are
// compared with a tolerance, integer and decision parts exactly).
namespace {

constexpr float PI = 3.14159265358979f;   // common/PI.hsh
constexpr float INV_PI = 0.31830988618f;
constexpr float EPSILON = 0.1f;           // raytracer/common.hsh:9
constexpr float INF = 1000000000000.0f;   // raytracer/common.hsh:8

inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline float saturate(float x) { return gclamp(x, 0.0f, 1.0f); }
inline float sqr(float x) { return x * x; }
inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline V3 mix3(V3 x, V3 y, float a) { return {mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)}; }

// common/random.hsh:5-48
inline uint32_t hash1(uint32_t x) {
    x += (x << 10u); x ^= (x >> 6u); x += (x << 3u); x ^= (x >> 11u); x += (x << 15u);
    return x;
}
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bitsf(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float float_construct(uint32_t m) { return bitsf((m & 0x007FFFFFu) | 0x3F800000u) - 1.0f; }
inline float random2(float x, float y) { return float_construct(hash1(fbits(x) ^ hash1(fbits(y)))); }       // random(vec2)
inline float random_seeded(float x, float& seed) { const float r = random2(x, seed); seed += 1.0f; return r; }   // random(float x, inout float seed)

// IEEE half <-> float, round to nearest even (what packHalf2x16 / unpackHalf2x16 do on the GPUs the engine targets).
inline uint32_t half_from_float(float f) {
    const uint32_t x = fbits(f), sign = (x >> 16) & 0x8000u;
    const uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u);
    if (ax >= 0x477ff000u) return sign | 0x7c00u;                      // rounds to >= 65520 -> inf
    if (ax < 0x33000001u) return sign;                                  // below half of the smallest subnormal
    int e = int(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;
    int shift = e < -14 ? (13 + (-14 - e)) : 13;
    uint32_t r = m >> shift;
    const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (r & 1u))) r++;
    if (e < -14) return sign | r;                                       // subnormal (r may carry into the exponent: correct)
    return sign | (uint32_t((e + 15) << 10) + (r - 0x400u));
}
inline float float_from_half(uint32_t h) {
    const uint32_t sign = (h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0) {
        if (m == 0) return bitsf(sign);
        const float v = float(m) * 5.9604644775390625e-8f;             // m * 2^-24
        return (sign ? -v : v);
    }
    if (e == 31) return bitsf(sign | 0x7f800000u | (m << 13));
    return bitsf(sign | ((e + 112u) << 23) | (m << 13));
}

struct Material {   // RaytraceMaterial, raytracer/structures.hsh:108-141 (23 words)
    int32_t ID; float baseR, baseG, baseB, emissR, emissG, emissB, opacity, roughness, metalness, ao, reflectance, normalScale;
    int32_t invertUVs, twoSided, cullBackFaces, useVertexColors, baseColorTexture, opacityTexture, normalTexture, roughnessTexture,
        metalnessTexture, aoTexture;
};
static_assert(sizeof(Material) == 92, "RaytraceMaterial");

struct Tex { uint32_t width, height; const uint8_t* texels; };

// textureLod(sampler2D(..), uv, 0).r of an R8 texture: bilinear, repeat addressing, texel centres at (i + 0.5) / size,
// weights in full fp32 (hardware filters with 8-bit weights: the product documents the same choice).
inline float sample_r8(const Tex& t, float u, float v) {
    const float x = u * float(t.width) - 0.5f, y = v * float(t.height) - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y);
    const float wx = x - fx, wy = y - fy;
    auto wrap = [](float f, uint32_t n) { int64_t i = int64_t(f) % int64_t(n); if (i < 0) i += n; return uint32_t(i); };
    const uint32_t x0 = wrap(fx, t.width), x1 = wrap(fx + 1.0f, t.width), y0 = wrap(fy, t.height), y1 = wrap(fy + 1.0f, t.height);
    auto tx = [&](uint32_t xx, uint32_t yy) { return float(t.texels[size_t(yy) * t.width + xx]) / 255.0f; };
    const float top = tx(x0, y0) * (1.0f - wx) + tx(x1, y0) * wx, bot = tx(x0, y1) * (1.0f - wx) + tx(x1, y1) * wx;
    return top * (1.0f - wy) + bot * wy;
}

struct Tri {   // Triangle after UnpackTriangle, raytracer/common.hsh:17-44
    V3 v0, v1, v2, n0, n1, n2;
    float uv[3][2];
    float col[3][4];
    int32_t materialIndex;
    float opacity;
};
inline V3 unpack_unit(uint32_t c) {   // common/packing.hsh:2-13 (xyz)
    return {float((c >> 0) & 1023u) / 1023.0f * 2.0f - 1.0f, float((c >> 10) & 1023u) / 1023.0f * 2.0f - 1.0f, float((c >> 20) & 1023u) / 1023.0f * 2.0f - 1.0f};
}
inline Tri unpack_triangle(const float* T) {   // 24 floats
    Tri t;
    t.v0 = {T[0], T[1], T[2]}; t.v1 = {T[4], T[5], T[6]}; t.v2 = {T[8], T[9], T[10]};
    t.n0 = unpack_unit(fbits(T[3])); t.n1 = unpack_unit(fbits(T[7])); t.n2 = unpack_unit(fbits(T[11]));
    for (int k = 0; k < 3; k++) {
        const uint32_t w = fbits(T[12 + k]);
        t.uv[k][0] = float_from_half(w & 0xffffu); t.uv[k][1] = float_from_half(w >> 16);
        const uint32_t c = fbits(T[20 + k]);
        for (int a = 0; a < 4; a++) t.col[k][a] = float((c >> (8 * a)) & 0xffu) / 255.0f;   // unpackUnorm4x8
    }
    t.materialIndex = int32_t(fbits(T[15]));
    t.opacity = T[23];
    return t;
}

// GetOpacity, raytracer/surface.hsh:147-160.
inline float get_opacity(const Tri& tri, float s, float t, const Material& m, const Tex* textures, uint32_t textureCount) {
    const float r = 1.0f - s - t;
    float u = r * tri.uv[0][0] + s * tri.uv[1][0] + t * tri.uv[2][0];
    float v = r * tri.uv[0][1] + s * tri.uv[1][1] + t * tri.uv[2][1];
    if (m.invertUVs > 0) v = 1.0f - v;
    const float tex = (m.opacityTexture < 0 || uint32_t(m.opacityTexture) >= textureCount) ? 1.0f : sample_r8(textures[m.opacityTexture], u, v);
    return tex * m.opacity;
}

struct Surface {   // brdf/surface.hsh:4-24
    V3 P, V, N, L, H, geometryNormal, F0;
    float NdotL, LdotH, NdotH, NdotV, F90;
    V3 baseColor, emissive;
    float opacity, roughness, metalness, ao, reflectance;
};

inline void update_surface(Surface& s) {   // brdf/surface.hsh:26-43
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

inline V3 fresnel_schlick(V3 F0, float F90, float c) {   // brdf/brdf.hsh:9-13
    const float p = std::pow(1.0f - c, 5.0f);
    return F0 + (V3{F90, F90, F90} - F0) * p;
}
inline float disney_diffuse(float NdotV, float NdotL, float LdotH, float lr) {   // brdf.hsh:15-25
    const float bias = mixf(0.0f, 0.5f, lr), factor = mixf(1.0f, 1.0f / 1.51f, lr);
    const float FD90 = bias + 2.0f * LdotH * LdotH * lr;
    const float ls = fresnel_schlick(V3{1, 1, 1}, FD90, NdotL).x, vs = fresnel_schlick(V3{1, 1, 1}, FD90, NdotV).x;
    return ls * vs * factor;
}
inline float vis_separable(float c, float alpha) {   // brdf.hsh:27-32
    const float a2 = alpha * alpha;
    return 2.0f * c / (c + std::sqrt(a2 + (1 - a2) * c * c));
}
inline float vis_correlated(float NdotL, float NdotV, float alpha) {   // brdf.hsh:34-43
    const float a2 = alpha * alpha;
    const float GGXL = NdotV * std::sqrt((-NdotL * a2 + NdotL) * NdotL + a2);
    const float GGXV = NdotL * std::sqrt((-NdotV * a2 + NdotV) * NdotV + a2);
    return 0.5f / (GGXL + GGXV + 0.0000001f);
}
inline float distribution_ggx(float NdotH, float alpha) {   // brdf.hsh:45-53
    const float a2 = alpha * alpha;
    const float f = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (f * f + 0.0000001f) * INV_PI;
}
inline V3 eval_diffuse(const Surface& s) {   // brdf/brdfEval.hsh:8-19
    const float roughness = gmax(sqr(s.roughness), 0.00001f);
    const float dd = disney_diffuse(s.NdotV, s.NdotL, s.LdotH, roughness);
    return s.baseColor * (1.0f - s.metalness) * dd * INV_PI;
}
inline V3 eval_specular(const Surface& s) {   // brdfEval.hsh:21-32
    const float roughness = gmax(sqr(s.roughness), 0.00001f);
    const V3 F = fresnel_schlick(s.F0, s.F90, s.LdotH);
    const float G = vis_correlated(s.NdotV, s.NdotL, roughness), D = distribution_ggx(s.NdotH, roughness);
    return F * D * G;
}

// Forward matrix of an instance: inverse(mat4(transpose(instance.inverseMatrix))) — surface.hsh:46-48. The affine part is
// inverted by cofactors in fp32 (the GLSL inverse() is implementation-defined in precision).
struct Affine { float m[3][3]; float t[3]; };
inline Affine forward_matrix(const float* I) {   // I: 12 floats = rows of the inverse matrix
    const float a = I[0], b = I[1], c = I[2], d = I[4], e = I[5], f = I[6], g = I[8], h = I[9], i = I[10];
    const float A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    const float det = a * A + b * B + c * C;
    const float inv = 1.0f / det;
    Affine r;
    r.m[0][0] = A * inv; r.m[0][1] = -(b * i - c * h) * inv; r.m[0][2] = (b * f - c * e) * inv;
    r.m[1][0] = B * inv; r.m[1][1] = (a * i - c * g) * inv;  r.m[1][2] = -(a * f - c * d) * inv;
    r.m[2][0] = C * inv; r.m[2][1] = -(a * h - b * g) * inv; r.m[2][2] = (a * e - b * d) * inv;
    const float tx = I[3], ty = I[7], tz = I[11];
    for (int k = 0; k < 3; k++) r.t[k] = -((r.m[k][0] * tx + r.m[k][1] * ty) + r.m[k][2] * tz);
    return r;
}
inline V3 xform_point(const Affine& M, V3 p) { return {((M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z) + M.t[0], ((M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z) + M.t[1], ((M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z) + M.t[2]}; }
inline V3 xform_dir(const Affine& M, V3 p) { return {(M.m[0][0] * p.x + M.m[0][1] * p.y) + M.m[0][2] * p.z, (M.m[1][0] * p.x + M.m[1][1] * p.y) + M.m[1][2] * p.z, (M.m[2][0] * p.x + M.m[2][1] * p.y) + M.m[2][2] * p.z}; }

struct PtScene {
    const float* instances; const float* const* triangles; const Material* materials; uint32_t materialCount; const Tex* textures; uint32_t textureCount;
};
struct PtParams {   // mirrors atlas_rt_pt_params (include/atlas_rt.h)
    float light_dir[3]; float light_radiance[3]; int32_t light_count; float sky_radiance[3]; uint32_t max_bounces; uint32_t samples_per_frame;
};
const Material kDefaultMaterial = {0, 0.8f, 0.8f, 0.8f, 0, 0, 0, 1.0f, 1.0f, 0.0f, 1.0f, 0.5f, 0.0f, 0, 1, 0, 0, -1, -1, -1, -1, -1, -1};

// GetSurfaceParameters, raytracer/surface.hsh:64-138 (+ TransformTriangle :44-62, GetTriangleMaterial :22-42); material
// textures other than opacity are not part of this path (treated as absent: the Sample*Bilinear functions return 1).
inline Surface surface_at(const PtScene& sc, const float* ray /*12 floats*/) {
    const int32_t hitID = int32_t(fbits(ray[9])), hitInst = int32_t(fbits(ray[10]));
    const float* I = sc.instances + 16 * size_t(hitInst);
    const int32_t meshOffset = int32_t(fbits(I[12])), materialOffset = int32_t(fbits(I[13]));
    Tri tri = unpack_triangle(sc.triangles[meshOffset] + 24 * size_t(hitID));
    const Affine M = forward_matrix(I);
    tri.v0 = xform_point(M, tri.v0); tri.v1 = xform_point(M, tri.v1); tri.v2 = xform_point(M, tri.v2);
    tri.n0 = normalize(xform_dir(M, tri.n0)); tri.n1 = normalize(xform_dir(M, tri.n1)); tri.n2 = normalize(xform_dir(M, tri.n2));
    const uint32_t mi = uint32_t(tri.materialIndex + materialOffset);
    const Material& rm = (sc.materials && mi < sc.materialCount) ? sc.materials[mi] : kDefaultMaterial;
    const V3 o{ray[0], ray[1], ray[2]}, d{ray[4], ray[5], ray[6]};
    // IntersectTriangle again, in world space (surface.hsh:72-78)
    const V3 e0 = tri.v1 - tri.v0, e1 = tri.v2 - tri.v0, sv = o - tri.v0;
    const V3 p = cross(sv, e0), q = cross(d, e1);
    const float den = dot(q, e0);
    const float dist = dot(p, e1) / den, s = dot(q, sv) / den, t = dot(p, d) / den;
    const float r = 1.0f - s - t;
    Surface sf{};
    sf.P = o + d * dist;
    float u = r * tri.uv[0][0] + s * tri.uv[1][0] + t * tri.uv[2][0], v = r * tri.uv[0][1] + s * tri.uv[1][1] + t * tri.uv[2][1];
    V3 normal = normalize((tri.n0 * r + tri.n1 * s) + tri.n2 * t);
    const V3 vc{r * tri.col[0][0] + s * tri.col[1][0] + t * tri.col[2][0], r * tri.col[0][1] + s * tri.col[1][1] + t * tri.col[2][1],
                r * tri.col[0][2] + s * tri.col[1][2] + t * tri.col[2][2]};
    if (rm.invertUVs > 0) v = 1.0f - v;
    V3 tn = normalize(cross(tri.v0 - tri.v1, tri.v0 - tri.v2));
    const bool flip = dot(tn, d) > 0.0f;
    if (flip && rm.twoSided > 0) tn = neg(tn);
    sf.geometryNormal = tn;
    sf.baseColor = V3{rm.baseR, rm.baseG, rm.baseB};
    if (rm.useVertexColors > 0) sf.baseColor = sf.baseColor * vc;
    sf.emissive = V3{rm.emissR, rm.emissG, rm.emissB};
    sf.opacity = rm.opacity * ((rm.opacityTexture < 0 || uint32_t(rm.opacityTexture) >= sc.textureCount) ? 1.0f : sample_r8(sc.textures[rm.opacityTexture], u, v));
    sf.roughness = rm.roughness; sf.metalness = rm.metalness; sf.ao = rm.ao; sf.reflectance = rm.reflectance;
    if (dot(normal, tn) < 0.0f && rm.twoSided > 0) normal = neg(normal);
    sf.V = neg(d);
    sf.N = normalize(normal);
    sf.F0 = mix3(V3{0.04f, 0.04f, 0.04f}, sf.baseColor, sf.metalness);
    sf.F90 = 1.0f;
    return sf;
}

inline void light_surface(Surface& sf, const PtParams& prm) {   // SampleLight for a directional light, raytracer/direct.hsh:79-86
    if (prm.light_count > 0) {
        sf.L = normalize(V3{prm.light_dir[0], prm.light_dir[1], prm.light_dir[2]});
        update_surface(sf);
    } else {   // no light: the shader reads NdotV uninitialised; we define it
        sf.NdotL = 0.0f;
        sf.NdotV = saturate(dot(sf.N, sf.V));
    }
}

}   // namespace

extern "C" {

// Primary rays, pathtracer/rayGen.csh:25-91 (non-REALTIME branch): jitter from the hash of the sample count, ID =
// Flatten2D(pixel) * samples + sample, 8x8-tile storage order with the ragged right / bottom borders.
void oracle_raygen(const float* eye, const float* origin, const float* right, const float* bottom, uint32_t width, uint32_t height, uint32_t samples,
                   int32_t sampleCount, float* out) {
    const float jx = random2(float(sampleCount), 0.0f), jy = random2(float(sampleCount), 1.0f);
    const uint32_t perfX = width / 8u, perfY = height / 8u, overX = width % 8u, overY = height % 8u;
    for (uint32_t s = 0; s < samples; s++)
        for (uint32_t y = 0; y < height; y++)
            for (uint32_t x = 0; x < width; x++) {
                const float cu = (float(x) + jx) / float(width), cv = (float(y) + jy) / float(height);
                V3 d;
                float* dd = &d.x;
                for (int k = 0; k < 3; k++) dd[k] = ((origin[k] + right[k] * cu) + bottom[k] * cv) - eye[k];
                d = normalize(d);
                const uint32_t gx = x / 8u, gy = y / 8u, local = (y % 8u) * 8u + (x % 8u);
                uint32_t index;
                if (gx < perfX && gy < perfY) index = local + (gy * perfX + gx) * 64u;
                else if (gx >= perfX && gy < perfY) index = y * overX + (x - perfX * 8u) + perfX * perfY * 64u;
                else index = x * overY + (y - perfY * 8u) + perfX * perfY * 64u + overX * perfY * 8u;
                float* o = out + 12 * (size_t(index) * samples + s);
                const int32_t id = int32_t((y * width + x) * samples + s);
                o[0] = eye[0]; o[1] = eye[1]; o[2] = eye[2]; std::memcpy(o + 3, &id, 4);
                o[4] = d.x; o[5] = d.y; o[6] = d.z; o[7] = 0.0f;
                o[8] = 0.0f; o[9] = 0.0f; o[10] = 0.0f; o[11] = 0.0f;     // ray.hitID = 0 (rayGen.csh:51)
            }
}

// Phase 1 of rayHit.csh for rays whose closest hits are known: the shadow ray of CheckVisibility (:327-337) per ray, or a
// dead ray (ID = -1) when none is cast (miss, no light, NdotL <= 0). shadow: n x 12.
void oracle_pt_shadow_rays(const float* instances, const float* const* triangles, const uint32_t* materials, uint32_t materialCount,
                           const uint32_t* texDims, const uint8_t* const* texels, uint32_t textureCount, const float* rays, uint64_t n,
                           const PtParams* prm, float* shadow) {
    Tex tex[64];
    for (uint32_t k = 0; k < textureCount && k < 64; k++) tex[k] = Tex{texDims[2 * k], texDims[2 * k + 1], texels[k]};
    const PtScene sc{instances, triangles, reinterpret_cast<const Material*>(materials), materialCount, tex, textureCount};
    for (uint64_t i = 0; i < n; i++) {
        const float* r = rays + 12 * i;
        float* s = shadow + 12 * i;
        const int32_t none = -1;
        for (int k = 0; k < 12; k++) s[k] = 0.0f;
        std::memcpy(s + 3, &none, 4); std::memcpy(s + 9, &none, 4);
        s[5] = 1.0f;
        const int32_t id = int32_t(fbits(r[3])), hitID = int32_t(fbits(r[9]));
        if (id < 0 || hitID < 0 || prm->light_count <= 0) continue;
        Surface sf = surface_at(sc, r);
        light_surface(sf, *prm);
        if (!(sf.NdotL > 0.0f)) continue;
        const V3 o = sf.P + sf.N * EPSILON;
        s[0] = o.x; s[1] = o.y; s[2] = o.z; std::memcpy(s + 3, &id, 4);
        s[4] = sf.L.x; s[5] = sf.L.y; s[6] = sf.L.z;
    }
}

// Phase 2: everything else rayHit.csh does (:56-158 main, EvaluateBounce :160-207, EvaluateDirectLight :209-235,
// EvaluateIndirectLight :237-325). visibility[i] = the transparency HitAnyTransparency returned for ray i's shadow ray.
// Outputs per INPUT ray i: alive[i]; if alive the next ray (12 floats, hit fields cleared) and its packed payload (4 words:
// half2 radiance.xy, half2 throughput.xy, half2 (radiance.z, throughput.z), 0 — PackRayPayload, common.hsh:88-98);
// otherwise finished[i] = the radiance the path adds to its pixel (main :119-123). rr[i] = (random, probability) of the
// Russian-roulette draw (for the test's borderline bookkeeping).
void oracle_pt_shade(const float* instances, const float* const* triangles, const uint32_t* materials, uint32_t materialCount,
                     const uint32_t* texDims, const uint8_t* const* texels, uint32_t textureCount, const float* rays, const uint32_t* payloadIn,
                     const float* visibility, uint64_t n, const PtParams* prm, float seed, uint32_t bounce, uint8_t* alive, float* raysOut,
                     uint32_t* payloadOut, float* finished, float* rr) {
    Tex tex[64];
    for (uint32_t k = 0; k < textureCount && k < 64; k++) tex[k] = Tex{texDims[2 * k], texDims[2 * k + 1], texels[k]};
    const PtScene sc{instances, triangles, reinterpret_cast<const Material*>(materials), materialCount, tex, textureCount};
    for (uint64_t i = 0; i < n; i++) {
        const float* r = rays + 12 * i;
        alive[i] = 0;
        for (int k = 0; k < 3; k++) finished[3 * i + k] = 0.0f;
        rr[2 * i] = rr[2 * i + 1] = 0.0f;
        const int32_t id = int32_t(fbits(r[3])), hitID = int32_t(fbits(r[9]));
        if (id < 0) continue;
        V3 radiance{0, 0, 0}, throughput{1, 1, 1};
        if (bounce > 0) {   // UnpackRayPayload, common.hsh:75-86
            const uint32_t* p = payloadIn + 4 * i;
            radiance = {float_from_half(p[0] & 0xffffu), float_from_half(p[0] >> 16), float_from_half(p[2] & 0xffffu)};
            throughput = {float_from_half(p[1] & 0xffffu), float_from_half(p[1] >> 16), float_from_half(p[2] >> 16)};
        }
        V3 o{r[0], r[1], r[2]}, d{r[4], r[5], r[6]};
        if (hitID == -1) {   // EvaluateBounce :166-171
            const V3 env = V3{prm->sky_radiance[0], prm->sky_radiance[1], prm->sky_radiance[2]} * 1.0f * throughput;
            radiance = radiance + V3{gmin(env.x, 10.0f), gmin(env.y, 10.0f), gmin(env.z, 10.0f)};
            throughput = {0, 0, 0};
        } else {
            Surface sf = surface_at(sc, r);
            if (dot(sf.emissive, V3{1, 1, 1}) > 0.0f && bounce == 0) radiance = radiance + sf.emissive;   // :180-183
            // EvaluateDirectLight :209-235 (one directional light: lightPdf = 1, solidAngle = 1)
            V3 direct{0, 0, 0};
            light_surface(sf, *prm);
            if (prm->light_count > 0) {
                V3 reflectance = (eval_diffuse(sf) + eval_specular(sf)) * sf.opacity;
                V3 rad = V3{prm->light_radiance[0], prm->light_radiance[1], prm->light_radiance[2]} * 1.0f;
                rad = rad * (sf.NdotL > 0.0f ? visibility[i] : 0.0f);
                direct = reflectance * rad * sf.NdotL / 1.0f;
            }
            V3 rad = throughput * sf.opacity * direct;   // :186-187
            if (bounce > 0) {   // :196-201
                const float limit = 10.0f;
                const float mx = gmax(gmax(rad.x, gmax(rad.y, rad.z)), limit);
                rad = rad * (limit / mx);
            }
            radiance = radiance + rad;
            // EvaluateIndirectLight :237-325
            o = sf.P;
            float curSeed = seed;
            const float raySeed = float(id);
            float refractChance = gclamp(1.0f - sf.opacity, 0.1f, 0.9f);
            refractChance = sf.opacity == 1.0f ? 0.0f : refractChance;
            float rnd = saturate(random_seeded(raySeed, curSeed));
            V3 L{0, 0, 0}, refl{0, 0, 0};
            float pdf = 0.0f;
            bool refracted = false;
            if (rnd >= refractChance) {
                rnd = random_seeded(raySeed, curSeed);
                const V3 F = fresnel_schlick(sf.F0, sf.F90, sf.NdotV);
                const float specChance = gclamp(dot(F, V3{0.33333f, 0.33333f, 0.33333f}), 0.1f, 0.9f);
                const float u0 = random_seeded(raySeed, curSeed), u1 = random_seeded(raySeed, curSeed);
                if (rnd < specChance) {   // SampleSpecularBRDF, brdf/brdfSample.hsh:57-92
                    const float alpha = sqr(sf.roughness);
                    sf.V = normalize(sf.V);
                    const V3 N = normalize(sf.N);
                    const V3 up = std::fabs(N.z) < 0.999f ? V3{0, 0, 1} : V3{1, 0, 0};
                    const V3 tangent = normalize(cross(up, N)), bitangent = normalize(cross(N, tangent));
                    // V * TBN = (dot(V, tangent), dot(V, bitangent), dot(V, N))
                    V3 Vt = normalize(V3{dot(sf.V, tangent), dot(sf.V, bitangent), dot(sf.V, N)});
                    // SampleGGXVNDF :34-55 with Xi = (u0, u1)
                    Vt = normalize(V3{Vt.x * alpha, Vt.y * alpha, Vt.z});
                    const float phi = PI * 2.0f * u0;
                    float cx = std::cos(phi), cy = std::sin(phi);
                    const float cz = (1.0f - u1) * (1.0f + Vt.z) + -Vt.z;
                    const float sc2 = std::sqrt(gclamp(1.0f - cz * cz, 0.0f, 1.0f));
                    cx *= sc2; cy *= sc2;
                    const V3 H{cx + Vt.x, cy + Vt.y, cz + Vt.z};
                    const V3 Hs{H.x * alpha, H.y * alpha, gmax(H.z, 0.0f)};
                    // TBN * v = tangent * v.x + bitangent * v.y + N * v.z
                    const V3 Mv = normalize((tangent * Hs.x + bitangent * Hs.y) + N * Hs.z);
                    sf.L = Mv * (2.0f * dot(sf.V, Mv)) - sf.V;
                    update_surface(sf);
                    pdf = 1.0f;
                    refl = {0, 0, 0};
                    if (sf.NdotL > 0.0f && sf.LdotH > 0.0f) {
                        const V3 F2 = fresnel_schlick(sf.F0, sf.F90, sf.LdotH);
                        const float Vis = vis_correlated(sf.NdotV, sf.NdotL, alpha), G1 = vis_separable(sf.NdotV, alpha);
                        L = sf.L;
                        pdf = G1 / (4.0f * std::fabs(dot(sf.V, sf.N)));
                        refl = F2 * Vis;
                    }
                    refl = refl * sf.opacity;
                    pdf *= specChance;
                } else {   // SampleDiffuseBRDF :8-32
                    const float theta = std::sqrt(u0), phi = 2.0f * PI * u1;
                    const V3 Ll{theta * std::cos(phi), theta * std::sin(phi), std::sqrt(1.0f - u0)};
                    const V3 N = sf.N;
                    const V3 up = std::fabs(N.z) < 0.999f ? V3{0, 0, 1} : V3{1, 0, 0};
                    const V3 tangent = normalize(cross(up, N)), bitangent = cross(N, tangent);
                    sf.L = normalize((tangent * Ll.x + bitangent * Ll.y) + N * Ll.z);
                    update_surface(sf);
                    L = sf.L;
                    pdf = sf.NdotL / PI;
                    refl = eval_diffuse(sf);
                    refl = refl * ((1.0f - sf.metalness) * sf.opacity);
                    pdf *= (1.0f - specChance);
                }
                pdf *= (1.0f - refractChance);
                o = o + sf.V * EPSILON;
            } else {
                L = d;
                pdf = refractChance;
                const float k = 1.0f - sf.opacity;
                refl = {k, k, k};
                sf.NdotL = 1.0f;
                o = o - sf.N * EPSILON;
                refracted = true;
            }
            if (pdf > 0.0f && dot(refl, V3{1, 1, 1}) > 0.0f) throughput = throughput * (refl * sf.NdotL / pdf);
            else throughput = {0, 0, 0};
            d = normalize(L);
            throughput = throughput * sf.ao;
            float probability = gclamp(gmax(throughput.x, gmax(throughput.y, throughput.z)), 0.01f, 0.99f);
            probability = bounce < 3 ? gmin(3.0f * probability, 1.0f) : probability;
            const float draw = random_seeded(raySeed, curSeed);
            rr[2 * i] = draw; rr[2 * i + 1] = probability;
            if (draw > probability) throughput = {0, 0, 0};
            else if (dot(d, sf.geometryNormal) <= 0.0f && !refracted) throughput = {0, 0, 0};
            else throughput = throughput / probability;
        }
        const float energy = dot(throughput, V3{1, 1, 1});
        if (energy == 0.0f || bounce == prm->max_bounces) {
            finished[3 * i] = radiance.x; finished[3 * i + 1] = radiance.y; finished[3 * i + 2] = radiance.z;
        } else {
            alive[i] = 1;
            float* q = raysOut + 12 * i;
            const int32_t none = -1;
            q[0] = o.x; q[1] = o.y; q[2] = o.z; std::memcpy(q + 3, &id, 4);
            q[4] = d.x; q[5] = d.y; q[6] = d.z; q[7] = 0.0f;
            q[8] = 0.0f; std::memcpy(q + 9, &none, 4); q[10] = 0.0f; q[11] = 0.0f;
            uint32_t* p = payloadOut + 4 * i;
            p[0] = half_from_float(radiance.x) | (half_from_float(radiance.y) << 16);
            p[1] = half_from_float(throughput.x) | (half_from_float(throughput.y) << 16);
            p[2] = half_from_float(radiance.z) | (half_from_float(throughput.z) << 16);
            p[3] = 0;
        }
    }
}

// DetermineRayBin, raytracer/tracing.hsh:18-21 with UnitVectorToOctahedron (common/octahedron.hsh:22-35) and Flatten2D.
// ivec2(x * 8.0) reaches 8 when a coordinate saturates to exactly 1.0, so bins run up to 8 * 8 + 8 = 72.
void oracle_ray_bins(const float* rays, uint64_t n, uint32_t* bins) {
    for (uint64_t i = 0; i < n; i++) {
        float x = rays[12 * i + 4], y = rays[12 * i + 5], z = rays[12 * i + 6];
        const float l1 = (std::fabs(x) + std::fabs(y)) + std::fabs(z);
        x /= l1; z /= l1;
        if (y < 0.0f) {
            const float ox = x, oz = z;
            x = (ox >= 0.0f ? 1.0f : -1.0f) * (1.0f - std::fabs(oz));
            z = (oz >= 0.0f ? 1.0f : -1.0f) * (1.0f - std::fabs(ox));
        }
        const float cx = saturate(0.5f * x + 0.5f), cz = saturate(0.5f * z + 0.5f);
        const int32_t ix = cvt_x86(cx * 8.0f), iz = cvt_x86(cz * 8.0f);
        bins[i] = uint32_t(iz * 8 + ix);
    }
}

// GetOpacity (surface.hsh:147-160) for triangle `tri` (24 floats) at barycentrics (s, t): exposed for the traversal tests.
float oracle_get_opacity(const float* tri, float s, float t, const uint32_t* material, const uint32_t* texDims, const uint8_t* const* texels,
                         uint32_t textureCount) {
    Tex tex[64];
    for (uint32_t k = 0; k < textureCount && k < 64; k++) tex[k] = Tex{texDims[2 * k], texDims[2 * k + 1], texels[k]};
    return get_opacity(unpack_triangle(tri), s, t, *reinterpret_cast<const Material*>(material), tex, textureCount);
}

}   // extern "C"
