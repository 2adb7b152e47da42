// Minimal stand-in for glm 0.9.8 (the version the reference pins in vcpkg.json:45-48).
// TEST INFRASTRUCTURE ONLY: lets /root/reference/src/engine/volume/*.cpp and jobsystem/*.cpp compile
// unmodified into oracle/_ref/. Only the operations those files use are provided, with glm 0.9.8 semantics:
//   min(x,y) = (y<x)?y:x   max(x,y) = (x<y)?y:x   clamp = min(max(x,lo),hi)   mix = x + a*(y-x)
//   dot = ((x*x'+y*y')+z*z')   cross = (a.y*b.z-b.y*a.z, a.z*b.x-b.z*a.x, a.x*b.y-b.x*a.y)
// plain fp32, no FMA (compile with -ffp-contract=off).
#pragma once
#include <cassert>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstddef>

namespace glm {

template<typename T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    explicit tvec2(T s) : x(s), y(s) {}
    tvec2(T x, T y) : x(x), y(y) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template<typename T> struct tvec4;

template<typename T> struct tvec3 {
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    tvec3(T x, T y, T z) : x(x), y(y), z(z) {}
    tvec3(const tvec4<T>& v);
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template<typename T> struct tvec4 {
    T x, y, z, w;
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    tvec4(T x, T y, T z, T w) : x(x), y(y), z(z), w(w) {}
    tvec4(const tvec3<T>& v, T w) : x(v.x), y(v.y), z(v.z), w(w) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template<typename T> tvec3<T>::tvec3(const tvec4<T>& v) : x(v.x), y(v.y), z(v.z) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int32_t> ivec2;
typedef tvec3<int32_t> ivec3;
typedef tvec4<int32_t> ivec4;

// ---- vec3 arithmetic (component-wise) ----
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { a = a + b; return a; }
inline vec3& operator-=(vec3& a, const vec3& b) { a = a - b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }

// ---- scalar helpers with glm's comparison forms ----
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3 clamp(const vec3& v, const vec3& lo, const vec3& hi) { return min(max(v, lo), hi); }
inline float mix(float x, float y, float a) { return x + a * (y - x); }
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x + a * (y - x); }

inline float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline float distance(const vec3& a, const vec3& b) { return length(b - a); }
inline vec3 normalize(const vec3& a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline float pow(float a, float b) { return std::pow(a, b); }

// ---- matrices (column-major, only what the compiled files touch) ----
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() { c[0] = vec4(1, 0, 0, 0); c[1] = vec4(0, 1, 0, 0); c[2] = vec4(0, 0, 1, 0); c[3] = vec4(0, 0, 0, 1); }
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
struct mat3x4 { vec4 c[3]; vec4& operator[](int i) { return c[i]; } const vec4& operator[](int i) const { return c[i]; } };
struct mat4x3 { vec3 c[4]; vec3& operator[](int i) { return c[i]; } const vec3& operator[](int i) const { return c[i]; } };
struct quat { float x, y, z, w; quat() : x(0), y(0), z(0), w(1) {} };

inline float determinant(const mat3& m) {
    return m[0].x * (m[1].y * m[2].z - m[2].y * m[1].z)
         - m[1].x * (m[0].y * m[2].z - m[2].y * m[0].z)
         + m[2].x * (m[0].y * m[1].z - m[1].y * m[0].z);
}
// glm 0.9.8 mat4 * vec4: Mov0*m[0] + Mov1*m[1] then Mov2*m[2] + Mov3*m[3], summed.
inline vec4 operator*(const mat4& m, const vec4& v) {
    vec4 r;
    for (int i = 0; i < 4; i++) {
        float a0 = m[0][i] * v.x, a1 = m[1][i] * v.y, a2 = m[2][i] * v.z, a3 = m[3][i] * v.w;
        r[i] = (a0 + a1) + (a2 + a3);
    }
    return r;
}

}
