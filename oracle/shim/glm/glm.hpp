// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Scalar stand-in for the subset of glm (header-only math library, an
// un-vendored, version-unpinned dependency of the reference: CMakeLists.txt:136
// `find_package(glm REQUIRED)`) that the reference's collision hot path names.
// glm is NOT present in this image, so its published scalar (non-SIMD)
// algorithms for glm 0.9.9.x are restated here.  PARITY UNPINNED for this
// header: the reference has no test that pins results at the glm boundary
// (SURVEY.md section 8c).  The operation order below *defines* the oracle and
// is mirrored one-to-one by prototype2_b200/csrc/prt_math.cuh on the device.
//
// Every function notes the glm source it restates.  This file contains no
// reference (prototype2) code.
#ifndef PRT_ORACLE_GLM_STANDIN_HPP
#define PRT_ORACLE_GLM_STANDIN_HPP

#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>
#include <limits>

namespace glm {

struct vec2 { float x, y; };

struct vec3 {
    float x, y, z;
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    float & operator[](size_t i) { return (&x)[i]; }
    float const & operator[](size_t i) const { return (&x)[i]; }
    // glm/detail/type_vec3.inl: component-wise compound ops
    vec3 & operator+=(vec3 const & v) { x += v.x; y += v.y; z += v.z; return *this; }
    vec3 & operator-=(vec3 const & v) { x -= v.x; y -= v.y; z -= v.z; return *this; }
    vec3 & operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
    vec3 & operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 const & v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float & operator[](size_t i) { return (&x)[i]; }
    float const & operator[](size_t i) const { return (&x)[i]; }
    // glm: vec<3>(vec<4>) conversion constructor (implicit unless
    // GLM_FORCE_EXPLICIT_CTOR); drops w.
    operator vec3() const { return vec3(x, y, z); }
};

// --- vec3 operators (glm/detail/type_vec3.inl) -----------------------------
inline vec3 operator+(vec3 const & a, vec3 const & b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 const & a, vec3 const & b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator-(vec3 const & a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 const & a, vec3 const & b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator*(vec3 const & a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 const & a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 const & a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator+(vec3 const & a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(vec3 const & a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

// --- vec4 operators (glm/detail/type_vec4.inl) -----------------------------
inline vec4 operator*(vec4 const & a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator+(vec4 const & a, vec4 const & b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// --- scalar functions (glm/detail/func_common.inl, func_exponential.inl) ----
inline float sqrt(float x) { return std::sqrt(x); }
inline float abs(float x) { return std::fabs(x); }
// glm::inversesqrt: `static_cast<T>(1) / sqrt(x)` (two roundings)
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
// glm::min / glm::max: `(y < x) ? y : x` / `(x < y) ? y : x`
inline float min(float x, float y) { return (y < x) ? y : x; }
inline float max(float x, float y) { return (x < y) ? y : x; }
inline vec3 min(vec3 const & a, vec3 const & b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(vec3 const & a, vec3 const & b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
// glm::mix(x, y, a) for floating a: `x * (1 - a) + y * a`
inline vec3 mix(vec3 const & x, vec3 const & y, float a) { return x * (1.0f - a) + y * a; }

// --- geometric functions (glm/detail/func_geometric.inl) -------------------
// compute_dot<vec<3>>: `tmp = a * b; return tmp.x + tmp.y + tmp.z;`
inline float dot(vec3 const & a, vec3 const & b) { vec3 t(a * b); return t.x + t.y + t.z; }
// compute_cross
inline vec3 cross(vec3 const & x, vec3 const & y) {
    return vec3(x.y * y.z - y.y * x.z,
                x.z * y.x - y.z * x.x,
                x.x * y.y - y.x * x.y);
}
inline float length(vec3 const & v) { return sqrt(dot(v, v)); }
inline float distance(vec3 const & p0, vec3 const & p1) { return length(p1 - p0); }
// glm/gtx/norm.inl
inline float length2(vec3 const & v) { return dot(v, v); }
inline float distance2(vec3 const & p0, vec3 const & p1) { return length2(p1 - p0); }
// compute_normalize: `v * inversesqrt(dot(v, v))`
inline vec3 normalize(vec3 const & v) { return v * inversesqrt(dot(v, v)); }

// --- quaternion (glm/detail/type_quat.inl, glm/ext/quaternion_geometric.inl)
struct quat {
    float x, y, z, w;            // glm 0.9.9 storage order x,y,z,w
    quat() : x(0.0f), y(0.0f), z(0.0f), w(1.0f) {}
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
};
// compute_dot<qua>: `tmp(a.w*b.w, a.x*b.x, a.y*b.y, a.z*b.z); (tmp.x+tmp.y)+(tmp.z+tmp.w)`
inline float dot(quat const & a, quat const & b) {
    float t0 = a.w * b.w, t1 = a.x * b.x, t2 = a.y * b.y, t3 = a.z * b.z;
    return (t0 + t1) + (t2 + t3);
}
inline float length(quat const & q) { return sqrt(dot(q, q)); }
inline quat normalize(quat const & q) {
    float len = length(q);
    if (len <= 0.0f) return quat(1.0f, 0.0f, 0.0f, 0.0f);
    float oneOverLen = 1.0f / len;
    return quat(q.w * oneOverLen, q.x * oneOverLen, q.y * oneOverLen, q.z * oneOverLen);
}

// --- matrices (glm/detail/type_mat4x4.inl) ---------------------------------
struct mat3 { vec3 c[3]; vec3 & operator[](size_t i) { return c[i]; } vec3 const & operator[](size_t i) const { return c[i]; } };

struct mat4 {
    vec4 c[4];                   // column-major
    mat4() {}
    explicit mat4(float s) { c[0] = vec4(s, 0, 0, 0); c[1] = vec4(0, s, 0, 0); c[2] = vec4(0, 0, s, 0); c[3] = vec4(0, 0, 0, s); }
    vec4 & operator[](size_t i) { return c[i]; }
    vec4 const & operator[](size_t i) const { return c[i]; }
};
// operator*(mat4, vec4): Add0 = m[0]*v[0] + m[1]*v[1]; Add1 = m[2]*v[2] + m[3]*v[3]; Add0 + Add1
inline vec4 operator*(mat4 const & m, vec4 const & v) {
    vec4 add0 = m[0] * v.x + m[1] * v.y;
    vec4 add1 = m[2] * v.z + m[3] * v.w;
    return add0 + add1;
}
// operator*(mat4, mat4): Result[i] = SrcA0*SrcBi[0] + SrcA1*SrcBi[1] + SrcA2*SrcBi[2] + SrcA3*SrcBi[3]
inline mat4 operator*(mat4 const & a, mat4 const & b) {
    mat4 r;
    for (int i = 0; i < 4; ++i) {
        r[i] = ((a[0] * b[i].x + a[1] * b[i].y) + a[2] * b[i].z) + a[3] * b[i].w;
    }
    return r;
}
// glm/ext/matrix_transform.inl translate: Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]
inline mat4 translate(mat4 const & m, vec3 const & v) {
    mat4 r(m);
    r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
    return r;
}
// glm/gtx/transform.inl scale(v) = scale(mat4(1), v): Result[i] = m[i] * v[i]
inline mat4 scale(vec3 const & v) {
    mat4 m(1.0f);
    mat4 r;
    r[0] = m[0] * v.x; r[1] = m[1] * v.y; r[2] = m[2] * v.z; r[3] = m[3];
    return r;
}
// glm/gtc/quaternion.inl mat3_cast / mat4_cast (toMat4)
inline mat4 toMat4(quat const & q) {
    mat4 r(1.0f);
    float qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z);
    float qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
    float qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
    r[0][0] = 1.0f - 2.0f * (qyy + qzz);
    r[0][1] = 2.0f * (qxy + qwz);
    r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz);
    r[1][1] = 1.0f - 2.0f * (qxx + qzz);
    r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy);
    r[2][1] = 2.0f * (qyz - qwx);
    r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}

} // namespace glm
#endif
