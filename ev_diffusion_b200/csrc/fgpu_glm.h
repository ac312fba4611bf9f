/*
 * fgpu_glm.h -- the slice of the glm API that FLAME GPU 1.5 agent scripts use.
 *
 * The reference's header.h includes the vendored glm 0.9.9 with
 * GLM_FORCE_NO_CTOR_INIT + GLM_FORCE_PURE (header.xslt:19-21) and typedefs
 * vec2..dvec4 from it (header.xslt:46-63).  Scripts use glm::vec2/3/4,
 * ivec*, uvec*, dvec* with arithmetic operators and length / normalize / dot /
 * cross / distance.  This header provides exactly that surface, written from
 * scratch, as templates -- so a script's own non-template `dot(glm::vec3,
 * glm::vec3)` / `length(glm::vec3)` overloads (Boids_Partitioning
 * functions.c:40-48) still win overload resolution under ADL, as they do
 * against real glm.
 *
 * Default constructors leave members uninitialised (GLM_FORCE_NO_CTOR_INIT).
 * All arithmetic is component-wise in the component type, in the order the
 * expression is written, so results are bit-identical to glm's pure path.
 */
#ifndef FGPU_GLM_H
#define FGPU_GLM_H

#include <math.h>

#if defined(__CUDACC__)
#define FGPU_GLM_FN __host__ __device__ inline
#else
#define FGPU_GLM_FN inline
#endif

namespace glm {

template <typename T> struct tvec2;
template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T>
struct tvec2 {
    T x, y;
    tvec2() = default;
    FGPU_GLM_FN explicit tvec2(T s) : x(s), y(s) {}
    FGPU_GLM_FN tvec2(T a, T b) : x(a), y(b) {}
    template <typename U> FGPU_GLM_FN explicit tvec2(const tvec2<U> &v) : x((T)v.x), y((T)v.y) {}
    template <typename U> FGPU_GLM_FN explicit tvec2(const tvec3<U> &v) : x((T)v.x), y((T)v.y) {}
    FGPU_GLM_FN T &operator[](int i) { return (&x)[i]; }
    FGPU_GLM_FN const T &operator[](int i) const { return (&x)[i]; }
    FGPU_GLM_FN tvec2 &operator+=(const tvec2 &o) { x += o.x; y += o.y; return *this; }
    FGPU_GLM_FN tvec2 &operator-=(const tvec2 &o) { x -= o.x; y -= o.y; return *this; }
    FGPU_GLM_FN tvec2 &operator*=(const tvec2 &o) { x *= o.x; y *= o.y; return *this; }
    FGPU_GLM_FN tvec2 &operator/=(const tvec2 &o) { x /= o.x; y /= o.y; return *this; }
    template <typename U> FGPU_GLM_FN tvec2 &operator+=(U s) { x += (T)s; y += (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec2 &operator-=(U s) { x -= (T)s; y -= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec2 &operator*=(U s) { x *= (T)s; y *= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec2 &operator/=(U s) { x /= (T)s; y /= (T)s; return *this; }
};

template <typename T>
struct tvec3 {
    T x, y, z;
    tvec3() = default;
    FGPU_GLM_FN explicit tvec3(T s) : x(s), y(s), z(s) {}
    FGPU_GLM_FN tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    template <typename A, typename B, typename C>
    FGPU_GLM_FN tvec3(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
    template <typename U> FGPU_GLM_FN explicit tvec3(const tvec3<U> &v) : x((T)v.x), y((T)v.y), z((T)v.z) {}
    template <typename U> FGPU_GLM_FN tvec3(const tvec2<U> &v, T c) : x((T)v.x), y((T)v.y), z(c) {}
    template <typename U> FGPU_GLM_FN explicit tvec3(const tvec4<U> &v) : x((T)v.x), y((T)v.y), z((T)v.z) {}
    FGPU_GLM_FN T &operator[](int i) { return (&x)[i]; }
    FGPU_GLM_FN const T &operator[](int i) const { return (&x)[i]; }
    FGPU_GLM_FN tvec3 &operator+=(const tvec3 &o) { x += o.x; y += o.y; z += o.z; return *this; }
    FGPU_GLM_FN tvec3 &operator-=(const tvec3 &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    FGPU_GLM_FN tvec3 &operator*=(const tvec3 &o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    FGPU_GLM_FN tvec3 &operator/=(const tvec3 &o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
    template <typename U> FGPU_GLM_FN tvec3 &operator+=(U s) { x += (T)s; y += (T)s; z += (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec3 &operator-=(U s) { x -= (T)s; y -= (T)s; z -= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec3 &operator*=(U s) { x *= (T)s; y *= (T)s; z *= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec3 &operator/=(U s) { x /= (T)s; y /= (T)s; z /= (T)s; return *this; }
};

template <typename T>
struct tvec4 {
    T x, y, z, w;
    tvec4() = default;
    FGPU_GLM_FN explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    FGPU_GLM_FN tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    template <typename U> FGPU_GLM_FN explicit tvec4(const tvec4<U> &v) : x((T)v.x), y((T)v.y), z((T)v.z), w((T)v.w) {}
    template <typename U> FGPU_GLM_FN tvec4(const tvec3<U> &v, T d) : x((T)v.x), y((T)v.y), z((T)v.z), w(d) {}
    FGPU_GLM_FN T &operator[](int i) { return (&x)[i]; }
    FGPU_GLM_FN const T &operator[](int i) const { return (&x)[i]; }
    FGPU_GLM_FN tvec4 &operator+=(const tvec4 &o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    FGPU_GLM_FN tvec4 &operator-=(const tvec4 &o) { x -= o.x; y -= o.y; z -= o.z; w -= o.w; return *this; }
    FGPU_GLM_FN tvec4 &operator*=(const tvec4 &o) { x *= o.x; y *= o.y; z *= o.z; w *= o.w; return *this; }
    FGPU_GLM_FN tvec4 &operator/=(const tvec4 &o) { x /= o.x; y /= o.y; z /= o.z; w /= o.w; return *this; }
    template <typename U> FGPU_GLM_FN tvec4 &operator+=(U s) { x += (T)s; y += (T)s; z += (T)s; w += (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec4 &operator-=(U s) { x -= (T)s; y -= (T)s; z -= (T)s; w -= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec4 &operator*=(U s) { x *= (T)s; y *= (T)s; z *= (T)s; w *= (T)s; return *this; }
    template <typename U> FGPU_GLM_FN tvec4 &operator/=(U s) { x /= (T)s; y /= (T)s; z /= (T)s; w /= (T)s; return *this; }
};

/* ---- binary operators ----------------------------------------------------- */
#define FGPU_GLM_BINOPS(OP)                                                                                      \
    template <typename T> FGPU_GLM_FN tvec2<T> operator OP(const tvec2<T> &a, const tvec2<T> &b) { return tvec2<T>(a.x OP b.x, a.y OP b.y); } \
    template <typename T> FGPU_GLM_FN tvec2<T> operator OP(const tvec2<T> &a, T s) { return tvec2<T>(a.x OP s, a.y OP s); }                 \
    template <typename T> FGPU_GLM_FN tvec2<T> operator OP(T s, const tvec2<T> &a) { return tvec2<T>(s OP a.x, s OP a.y); }                 \
    template <typename T> FGPU_GLM_FN tvec3<T> operator OP(const tvec3<T> &a, const tvec3<T> &b) { return tvec3<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z); } \
    template <typename T> FGPU_GLM_FN tvec3<T> operator OP(const tvec3<T> &a, T s) { return tvec3<T>(a.x OP s, a.y OP s, a.z OP s); }       \
    template <typename T> FGPU_GLM_FN tvec3<T> operator OP(T s, const tvec3<T> &a) { return tvec3<T>(s OP a.x, s OP a.y, s OP a.z); }       \
    template <typename T> FGPU_GLM_FN tvec4<T> operator OP(const tvec4<T> &a, const tvec4<T> &b) { return tvec4<T>(a.x OP b.x, a.y OP b.y, a.z OP b.z, a.w OP b.w); } \
    template <typename T> FGPU_GLM_FN tvec4<T> operator OP(const tvec4<T> &a, T s) { return tvec4<T>(a.x OP s, a.y OP s, a.z OP s, a.w OP s); } \
    template <typename T> FGPU_GLM_FN tvec4<T> operator OP(T s, const tvec4<T> &a) { return tvec4<T>(s OP a.x, s OP a.y, s OP a.z, s OP a.w); }
FGPU_GLM_BINOPS(+)
FGPU_GLM_BINOPS(-)
FGPU_GLM_BINOPS(*)
FGPU_GLM_BINOPS(/)
#undef FGPU_GLM_BINOPS

template <typename T> FGPU_GLM_FN tvec2<T> operator-(const tvec2<T> &a) { return tvec2<T>(-a.x, -a.y); }
template <typename T> FGPU_GLM_FN tvec3<T> operator-(const tvec3<T> &a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> FGPU_GLM_FN tvec4<T> operator-(const tvec4<T> &a) { return tvec4<T>(-a.x, -a.y, -a.z, -a.w); }
template <typename T> FGPU_GLM_FN bool operator==(const tvec2<T> &a, const tvec2<T> &b) { return a.x == b.x && a.y == b.y; }
template <typename T> FGPU_GLM_FN bool operator==(const tvec3<T> &a, const tvec3<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> FGPU_GLM_FN bool operator==(const tvec4<T> &a, const tvec4<T> &b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
template <typename T> FGPU_GLM_FN bool operator!=(const tvec2<T> &a, const tvec2<T> &b) { return !(a == b); }
template <typename T> FGPU_GLM_FN bool operator!=(const tvec3<T> &a, const tvec3<T> &b) { return !(a == b); }
template <typename T> FGPU_GLM_FN bool operator!=(const tvec4<T> &a, const tvec4<T> &b) { return !(a == b); }

/* ---- geometric functions (glm/detail/func_geometric.inl semantics) -------- */
namespace detail {
FGPU_GLM_FN float fsqrt(float v) { return sqrtf(v); }
FGPU_GLM_FN double fsqrt(double v) { return sqrt(v); }
}  // namespace detail

/* dot: glm computes tmp = a*b then sums left to right: (x + y) + z (+ w) */
template <typename T> FGPU_GLM_FN T dot(const tvec2<T> &a, const tvec2<T> &b) { tvec2<T> t(a * b); return t.x + t.y; }
template <typename T> FGPU_GLM_FN T dot(const tvec3<T> &a, const tvec3<T> &b) { tvec3<T> t(a * b); return t.x + t.y + t.z; }
template <typename T> FGPU_GLM_FN T dot(const tvec4<T> &a, const tvec4<T> &b) { tvec4<T> t(a * b); return (t.x + t.y) + (t.z + t.w); }
template <typename T> FGPU_GLM_FN T length(const tvec2<T> &v) { return detail::fsqrt(dot(v, v)); }
template <typename T> FGPU_GLM_FN T length(const tvec3<T> &v) { return detail::fsqrt(dot(v, v)); }
template <typename T> FGPU_GLM_FN T length(const tvec4<T> &v) { return detail::fsqrt(dot(v, v)); }
template <typename T> FGPU_GLM_FN T distance(const tvec2<T> &a, const tvec2<T> &b) { return length(b - a); }
template <typename T> FGPU_GLM_FN T distance(const tvec3<T> &a, const tvec3<T> &b) { return length(b - a); }
/* normalize: v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt */
template <typename T> FGPU_GLM_FN tvec2<T> normalize(const tvec2<T> &v) { return v * ((T)1 / detail::fsqrt(dot(v, v))); }
template <typename T> FGPU_GLM_FN tvec3<T> normalize(const tvec3<T> &v) { return v * ((T)1 / detail::fsqrt(dot(v, v))); }
template <typename T> FGPU_GLM_FN tvec4<T> normalize(const tvec4<T> &v) { return v * ((T)1 / detail::fsqrt(dot(v, v))); }
template <typename T> FGPU_GLM_FN tvec3<T> cross(const tvec3<T> &a, const tvec3<T> &b) {
    return tvec3<T>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<float> fvec2;
typedef tvec3<float> fvec3;
typedef tvec4<float> fvec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec4<int> ivec4;
typedef tvec2<unsigned int> uvec2;
typedef tvec3<unsigned int> uvec3;
typedef tvec4<unsigned int> uvec4;
typedef tvec2<double> dvec2;
typedef tvec3<double> dvec3;
typedef tvec4<double> dvec4;

}  // namespace glm

#endif
