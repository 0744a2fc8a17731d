// mini_glm.hpp — just enough of glm's vector surface to compile the reference's portable hot-path sources (NaiveFracturer.cpp,
// Seeder.cpp, RandomUtilities.h, FractureParameters.h, the Möller SAT of Intersections3D.h) in this container, where glm
// 0.9.9.8 (README badge) is not installed.  Plain IEEE component-wise arithmetic, same expression shapes as glm
// (dot = x*x' + y*y' + z*z', distance = sqrt(dot(d,d)), cross as in glm/detail/func_geometric.inl).  TEST TOOLING ONLY.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace glm {
template <int N, class T> struct vec;
template <class T> struct vec<1, T> { T x; vec() : x(0) {} vec(T a) : x(a) {} };
template <class T> struct vec<2, T> { T x, y; vec() : x(0), y(0) {} vec(T a) : x(a), y(a) {} vec(T a, T b) : x(a), y(b) {} };
template <class T> struct vec<3, T> {
    T x, y, z;
    vec() : x(0), y(0), z(0) {}
    vec(T a) : x(a), y(a), z(a) {}
    template <class A, class B, class C> vec(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
    template <class U> vec(const vec<3, U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    template <class U> vec(const vec<4, U>& o);
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    vec& operator-=(const vec& o) { x -= o.x, y -= o.y, z -= o.z; return *this; }
    vec& operator+=(const vec& o) { x += o.x, y += o.y, z += o.z; return *this; }
};
template <class T> struct vec<4, T> {
    T x, y, z, w;
    vec() : x(0), y(0), z(0), w(0) {}
    vec(T a) : x(a), y(a), z(a), w(a) {}
    template <class A, class B, class C, class D> vec(A a, B b, C c, D d) : x((T)a), y((T)b), z((T)c), w((T)d) {}
    template <class U, class D> vec(const vec<3, U>& o, D d) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)d) {}
    template <class U> vec(const vec<4, U>& o) : x((T)o.x), y((T)o.y), z((T)o.z), w((T)o.w) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <class T> template <class U> vec<3, T>::vec(const vec<4, U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}

#define MINI_GLM_OP(op)                                                                                                            \
    template <class T> vec<3, T> operator op(const vec<3, T>& a, const vec<3, T>& b) { return vec<3, T>(a.x op b.x, a.y op b.y, a.z op b.z); } \
    template <class T> vec<3, T> operator op(const vec<3, T>& a, T s) { return vec<3, T>(a.x op s, a.y op s, a.z op s); }         \
    template <class T> vec<3, T> operator op(T s, const vec<3, T>& a) { return vec<3, T>(s op a.x, s op a.y, s op a.z); }         \
    template <class T> vec<4, T> operator op(const vec<4, T>& a, const vec<4, T>& b) { return vec<4, T>(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }
MINI_GLM_OP(+) MINI_GLM_OP(-) MINI_GLM_OP(*) MINI_GLM_OP(/)
#undef MINI_GLM_OP
template <class T> vec<3, T> operator-(const vec<3, T>& a) { return vec<3, T>(-a.x, -a.y, -a.z); }
template <class T> bool operator==(const vec<3, T>& a, const vec<3, T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

typedef vec<1, float> vec1; typedef vec<2, float> vec2; typedef vec<3, float> vec3; typedef vec<4, float> vec4;
typedef vec<1, int> ivec1; typedef vec<2, int> ivec2; typedef vec<3, int> ivec3; typedef vec<4, int> ivec4;
typedef vec<1, unsigned> uvec1; typedef vec<2, unsigned> uvec2; typedef vec<3, unsigned> uvec3; typedef vec<4, unsigned> uvec4;
typedef unsigned uint;
struct mat3 { float m[9]; }; struct mat4 { float m[16]; };

template <class T> T min(T a, T b) { return b < a ? b : a; }
template <class T> T max(T a, T b) { return a < b ? b : a; }
template <class T> vec<2, T> min(vec<2, T> a, vec<2, T> b) { return vec<2, T>(min(a.x, b.x), min(a.y, b.y)); }  // component-wise, as glm (GStack::updateBoundaries)
template <class T> vec<2, T> max(vec<2, T> a, vec<2, T> b) { return vec<2, T>(max(a.x, b.x), max(a.y, b.y)); }
template <class T> T abs(T a) { return a < 0 ? -a : a; }
template <class T> T clamp(T v, T lo, T hi) { return min(max(v, lo), hi); }
template <class T> T dot(const vec<3, T>& a, const vec<3, T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> T length2(const vec<3, T>& a) { return dot(a, a); }
template <class T> T length(const vec<3, T>& a) { return std::sqrt(dot(a, a)); }
template <class T> T distance(const vec<3, T>& a, const vec<3, T>& b) { return length(b - a); }
template <class T> vec<3, T> cross(const vec<3, T>& x, const vec<3, T>& y) { return vec<3, T>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
template <class T> T pi() { return (T)3.14159265358979323846264338327950288; }
template <class T> T epsilon() { return std::numeric_limits<T>::epsilon(); }
}  // namespace glm
namespace glm {
inline vec4 operator*(const mat4& m, const vec4& v)  // column-major, as glm (AABB::dot only; unused by the checker)
{
    return vec4(m.m[0] * v.x + m.m[4] * v.y + m.m[8] * v.z + m.m[12] * v.w, m.m[1] * v.x + m.m[5] * v.y + m.m[9] * v.z + m.m[13] * v.w,
                m.m[2] * v.x + m.m[6] * v.y + m.m[10] * v.z + m.m[14] * v.w, m.m[3] * v.x + m.m[7] * v.y + m.m[11] * v.z + m.m[15] * v.w);
}
}  // namespace glm
