// glsl_rt.h — TEST INFRASTRUCTURE (oracle/): just enough of the GLSL 4.50 compute language, as C++, for the reference's fracturer
// shaders to compile unmodified where they lie (oracle/ref_shim/glsl2cpp.py turns each `*-comp.glsl` under
// /root/reference/MeshFragments/Assets/Shaders/Compute/Fracturer into an include file: `#include <...>` expanded the way
// ShaderProgram::includeLibraries does (ShaderProgram.cpp:362-404), buffer blocks turned into pointers, `main` renamed; nothing in a
// shader's body is rewritten except `.xyz` -> `.xyz()`).  Only what those shaders use is here.  Arithmetic follows the GLSL rules the
// shaders rely on: uint wraps modulo 2^32, float is IEEE binary32 (the file is compiled with -ffp-contract=off), uvec + ivec adds in uint.
#pragma once

#include <stdint.h>

#include <stddef.h>

#include <cmath>

typedef unsigned int uint;

struct uvec3;
struct ivec3;

struct vec2 {
    float x, y;
};

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(const uvec3& v);  // implicit, as GLSL converts uvec3 -> vec3 in a call
    vec3(const ivec3& v);  // implicit: `vec3 p = cell + neighbour` (marchingCubes-comp.glsl:70-71)
};
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(double s, const vec3& a) { return vec3((float)s / a.x, (float)s / a.y, (float)s / a.z); }  // `1.0 / vec3(...)`: a float literal in GLSL
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

struct uvec3 {
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(const ivec3& v);  // implicit int -> uint, component-wise (getPositionIndex(cell + neighbour), marchingCubes-comp.glsl:110)
    template <typename A, typename B, typename C>
    uvec3(A a, B b, C c) : x((uint)a), y((uint)b), z((uint)c) {}  // uvec3(float, float, float) truncates towards zero
    explicit uvec3(uint s) : x(s), y(s), z(s) {}
};
inline vec3::vec3(const uvec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}

struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit ivec3(int s) : x(s), y(s), z(s) {}
    explicit ivec3(uint s) : x((int)s), y((int)s), z((int)s) {}
    explicit ivec3(const uvec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
    explicit ivec3(const vec3& v);  // truncation towards zero
};
inline ivec3::ivec3(const vec3& v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
inline uvec3::uvec3(const ivec3& v) : x((uint)v.x), y((uint)v.y), z((uint)v.z) {}
inline vec3::vec3(const ivec3& v) : x((float)v.x), y((float)v.y), z((float)v.z) {}
inline uvec3 operator+(const uvec3& a, const uvec3& b) { return uvec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator+(const ivec3& a, const ivec3& b) { return ivec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline ivec3 operator-(const ivec3& a, const ivec3& b) { return ivec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline int glsl_clamp1(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }  // min(max(x, minVal), maxVal)
inline ivec3 clamp(const ivec3& v, const ivec3& lo, const ivec3& hi) { return ivec3(glsl_clamp1(v.x, lo.x, hi.x), glsl_clamp1(v.y, lo.y, hi.y), glsl_clamp1(v.z, lo.z, hi.z)); }

struct ivec2 {
    int x, y;
};
struct mat4 {
    float m[16];  // column-major, as GLSL stores it: m[4 * column + row]
};
// `v.xyz` of a vec4 as an lvalue: the generator writes `.xyz()`, which yields this view (reads convert to vec3)
struct vec3ref {
    float &x, &y, &z;
    vec3ref& operator=(const vec3& v) { x = v.x, y = v.y, z = v.z; return *this; }
    vec3ref& operator=(const vec3ref& v) { const float a = v.x, b = v.y, c = v.z; x = a, y = b, z = c; return *this; }
    operator vec3() const { return vec3(x, y, z); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec3ref xyz() { return vec3ref{ x, y, z }; }
    vec3 xyz() const { return vec3(x, y, z); }
};
// matrix * column vector, the four products of a row added left to right (a scale + translation matrix gives s * x + t exactly as written)
inline vec4 operator*(const mat4& a, const vec4& v)
{
    vec4 r;
    r.x = a.m[0] * v.x + a.m[4] * v.y + a.m[8] * v.z + a.m[12] * v.w;
    r.y = a.m[1] * v.x + a.m[5] * v.y + a.m[9] * v.z + a.m[13] * v.w;
    r.z = a.m[2] * v.x + a.m[6] * v.y + a.m[10] * v.z + a.m[14] * v.w;
    r.w = a.m[3] * v.x + a.m[7] * v.y + a.m[11] * v.z + a.m[15] * v.w;
    return r;
}
inline vec3 mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }  // x * (1 - a) + y * a (GLSL 4.50 section 8.3)
inline int clamp(int v, int lo, int hi) { return glsl_clamp1(v, lo, hi); }
struct ivec4 {
    int x, y, z, w;
    ivec4() : x(0), y(0), z(0), w(0) {}
    ivec4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
    ivec4(const ivec3& v, int d) : x(v.x), y(v.y), z(v.z), w(d) {}
    int& operator[](int i) { return i == 0 ? x : i == 1 ? y : i == 2 ? z : w; }
};
struct uvec4 {
    uint x, y, z, w;
    uvec4() : x(0), y(0), z(0), w(0) {}
    uvec4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
    template <typename W>
    uvec4(const uvec3& v, W d) : x(v.x), y(v.y), z(v.z), w((uint)d) {}
    uvec3 xyz() const { return uvec3(x, y, z); }
    uint& operator[](uint i) { return i == 0 ? x : i == 1 ? y : i == 2 ? z : w; }
};
// uvec4 + ivec4: the signed operand converts to unsigned, the sum wraps (floodFracturer-comp.glsl:33 relies on it for "-1")
inline uvec4 operator+(const uvec4& a, const ivec4& b) { return uvec4(a.x + (uint)b.x, a.y + (uint)b.y, a.z + (uint)b.z, a.w + (uint)b.w); }

inline float distance(const vec3& p, const vec3& q)
{
    const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
    return sqrtf(dx * dx + dy * dy + dz * dz);
}
// abs / max / floor: the including file brings std::abs, std::max and std::floor into each shader's namespace

// A shader storage block `buffer B { T name[]; }`.  Out-of-range accesses behave as under robust buffer access (reads give zero, writes are
// dropped): floodFracturer-comp.glsl:36 reads grid[] at a wrapped neighbour index BEFORE its bounds test.
template <typename T>
struct glsl_buffer {
    T* p = nullptr;
    size_t n = 0;
    T sink;
    void bind(T* ptr, size_t count) { p = ptr, n = count; }
    T& operator[](size_t i)
    {
        if (i < n) return p[i];
        sink = T();
        return sink;
    }
};

inline uint atomicAdd(uint& mem, uint data) { return __atomic_fetch_add(&mem, data, __ATOMIC_RELAXED); }
inline int atomicAdd(int& mem, int data) { return __atomic_fetch_add(&mem, data, __ATOMIC_RELAXED); }
inline uint atomicMin(uint& mem, uint data)
{
    uint old = __atomic_load_n(&mem, __ATOMIC_RELAXED);
    while (data < old && !__atomic_compare_exchange_n(&mem, &old, data, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// gl_GlobalInvocationID of the invocation running on this host thread
extern thread_local uvec3 gl_GlobalInvocationID;

#define uniform      /* a plain namespace-scope variable the driver assigns */
#define in           /* parameter qualifier */
