// Stand-in for <glm/gtc/packing.hpp> (glm 0.9.8, not vendored by the reference): only what src/engine/common/Packing.cpp
// needs to COMPILE. The two snorm functions are restated from glm's published definition
// (packSnorm3x10_1x2: round(clamp(v, -1, 1) * vec4(511, 511, 511, 1)) into 10/10/10/2-bit signed fields); nothing on the
// accelerated path calls them - the path uses Packing::PackSignedVector3x10_1x2, which is plain arithmetic in the reference.
#pragma once
#include <cmath>
#include <cstdint>
#include "../glm.hpp"

namespace glm {
inline uint32_t packSnorm3x10_1x2(const vec4& v) {
    auto q = [](float x, float s) { const float c = x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); return int32_t(std::round(c * s)); };
    const uint32_t x = uint32_t(q(v.x, 511.0f)) & 1023u, y = uint32_t(q(v.y, 511.0f)) & 1023u, z = uint32_t(q(v.z, 511.0f)) & 1023u;
    const uint32_t w = uint32_t(q(v.w, 1.0f)) & 3u;
    return x | (y << 10) | (z << 20) | (w << 30);
}
inline vec4 unpackSnorm3x10_1x2(uint32_t p) {
    auto s = [](uint32_t f, int bits) { const int32_t v = int32_t(f << (32 - bits)) >> (32 - bits); return float(v); };
    auto c = [](float x) { return x < -1.0f ? -1.0f : (x > 1.0f ? 1.0f : x); };
    return vec4(c(s(p & 1023u, 10) / 511.0f), c(s((p >> 10) & 1023u, 10) / 511.0f), c(s((p >> 20) & 1023u, 10) / 511.0f), c(s(p >> 30, 2)));
}
}   // namespace glm
