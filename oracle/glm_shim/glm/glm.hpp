// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Minimal scalar stand-in for the subset of g-truc/glm that the reference's
// svo_builder sources use (glm is neither vendored under /root/reference nor
// installed in this image; the reference pins no glm version -- its CI installs
// whatever `libglm-dev` provides, .github/workflows/build_cmake.yml:16-18).
//
// The arithmetic below restates glm's published scalar (non-SIMD) definitions:
//   dot       products first, then summed left to right        (glm/detail/func_geometric.inl, compute_dot)
//   cross     (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)   (compute_cross)
//   length    sqrt(dot(v,v))
//   normalize v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   inverse   (mat3) 1/det then nine cofactors times 1/det      (glm/detail/func_matrix.inl, compute_inverse<3,3>)
//   mat*vec   m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z           (glm/detail/type_mat3x3.inl)
// Every float expression keeps exactly that operation order; the oracle is
// built without FMA contraction so each operation rounds once to binary32.
//
// Call sites served: voxelizer.cpp:207-254,266-287; BarycentricCoords.h:6,12-15,26-30;
// OctreeBuilder.cpp:88-93; main.cpp:380-381; TriReader.h:76; intersection.h.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>

namespace glm {

template <typename T>
struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    tvec2(T a, T b) : x(a), y(b) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

// A fourth, unused lane keeps the reference's 8-byte `(uint_fast32_t&)` stores
// into 4-byte uvec3 elements (voxelizer.cpp:149-150, partitioner.cpp:52-53)
// inside the object: the store to element 2 lands in `pad_` instead of in the
// neighbouring variable. Values read back are unchanged.
template <typename T>
struct tvec3 {
    T x, y, z;
    T pad_;
    tvec3() : x(0), y(0), z(0), pad_(0) {}
    tvec3(T a, T b, T c) : x(a), y(b), z(c), pad_(0) {}
    template <typename U>
    explicit tvec3(const tvec3<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), pad_(0) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

// The reference memcpy's triangles as packed 3-float vertices (tri_util.h:29-55,
// tri_tools.h:41-55), so the float vec3 must be exactly 12 bytes: specialise
// without the pad lane.
template <>
struct tvec3<float> {
    float x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    tvec3(float a, float b, float c) : x(a), y(b), z(c) {}
    tvec3(double a, double b, double c) : x(float(a)), y(float(b)), z(float(c)) {}
    tvec3(int a, int b, int c) : x(float(a)), y(float(b)), z(float(c)) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec3<int> ivec3;
typedef tvec3<unsigned int> uvec3;

// ---- vec2 ----
inline vec2 operator*(float s, const vec2& v) { return vec2(s * v.x, s * v.y); }
inline vec2 operator*(const vec2& v, float s) { return vec2(v.x * s, v.y * s); }
inline vec2 operator-(const vec2& v) { return vec2(-v.x, -v.y); }
inline float dot(const vec2& a, const vec2& b) {
    float px = a.x * b.x, py = a.y * b.y;
    return px + py;
}

// ---- vec3 ----
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(float s, const vec3& v) { return vec3(s * v.x, s * v.y, s * v.z); }
inline vec3 operator*(const vec3& v, float s) { return vec3(v.x * s, v.y * s, v.z * s); }
inline vec3 operator/(const vec3& v, float s) { return vec3(v.x / s, v.y / s, v.z / s); }
inline bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline float dot(const vec3& a, const vec3& b) {
    float px = a.x * b.x, py = a.y * b.y, pz = a.z * b.z;
    return px + py + pz;
}
inline vec3 cross(const vec3& x, const vec3& y) {
    return vec3(x.y * y.z - y.y * x.z,
                x.z * y.x - y.z * x.x,
                x.x * y.y - y.x * x.y);
}
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }

// ---- mat3 (column major: m[c][r]) ----
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};

inline mat3 inverse(const mat3& m) {
    float OneOverDeterminant = 1.0f / (
        + m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    mat3 I;
    I[0][0] = + (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * OneOverDeterminant;
    I[1][0] = - (m[1][0] * m[2][2] - m[2][0] * m[1][2]) * OneOverDeterminant;
    I[2][0] = + (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * OneOverDeterminant;
    I[0][1] = - (m[0][1] * m[2][2] - m[2][1] * m[0][2]) * OneOverDeterminant;
    I[1][1] = + (m[0][0] * m[2][2] - m[2][0] * m[0][2]) * OneOverDeterminant;
    I[2][1] = - (m[0][0] * m[2][1] - m[2][0] * m[0][1]) * OneOverDeterminant;
    I[0][2] = + (m[0][1] * m[1][2] - m[1][1] * m[0][2]) * OneOverDeterminant;
    I[1][2] = - (m[0][0] * m[1][2] - m[1][0] * m[0][2]) * OneOverDeterminant;
    I[2][2] = + (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * OneOverDeterminant;
    return I;
}

inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

// ---- scalar helpers (templates so abs(float) resolves to std::abs) ----
template <typename T> inline T min(T a, T b) { return (b < a) ? b : a; }
template <typename T> inline T max(T a, T b) { return (a < b) ? b : a; }
template <typename T> inline T abs(T a) { return std::abs(a); }

}  // namespace glm
