/*
 * vf_oracle.cpp — CPU ORACLE (test infrastructure, never shipped, never on the product path).
 * See vf_oracle.h for scope and pin status.  Every function cites the reference file:line it restates.
 * Build: make -C oracle   (g++ -O2 -fopenmp -ffp-contract=off: NO fused multiply-add, the SAT test is
 * float32 operation-order sensitive, SURVEY §7 "SAT epsilon").
 */
#include "vf_oracle.h"

#include <algorithm>
#include <array>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <queue>
#include <random>
#include <set>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct U3 { uint32_t x, y, z; };

inline uint64_t lin(uint32_t x, uint32_t y, uint32_t z, const uint32_t d[3])
{
    /* RegularGrid.cpp:839-842, voxel.glsl:16-19: x slowest, z fastest */
    return (uint64_t)x * d[1] * d[2] + (uint64_t)y * d[2] + z;
}

/* FloodFracturer.cpp:8-27 — neighbour tables, reference order */
const int VON_NEUMANN[6][3] = { {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1} };
const int MOORE[26][3] = {
    {-1, -1, -1}, {1, 1, 1}, {-1, -1, 0}, {1, 1, 0}, {-1, -1, 1}, {1, 1, -1}, {0, -1, -1}, {0, 1, 1}, {0, -1, 0},
    {0, 1, 0},    {0, -1, 1}, {0, 1, -1}, {1, -1, -1}, {-1, 1, 1}, {1, -1, 0}, {-1, 1, 0}, {1, -1, 1}, {-1, 1, -1},
    {-1, 0, -1},  {1, 0, 1},  {-1, 0, 0}, {1, 0, 0},   {-1, 0, 1}, {1, 0, -1}, {0, 0, -1}, {0, 0, 1}
};

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

/* ===================================================================================== RNG */

struct orc_rng {
    std::mt19937 gen; /* RandomUtilities.h:11 typedef std::mt19937 RandomNumberGenerator */
};

extern "C" orc_rng* orc_rng_create(uint32_t seed)
{
    orc_rng* r = new orc_rng();
    r->gen.seed(seed); /* RandomUtilities.h:86-89 initSeed */
    return r;
}
extern "C" void orc_rng_destroy(orc_rng* r) { delete r; }
extern "C" void orc_rng_seed(orc_rng* r, uint32_t seed) { r->gen.seed(seed); }
extern "C" uint32_t orc_rng_raw(orc_rng* r) { return (uint32_t)r->gen(); }

/* SURVEY finding 9: std::uniform_real_distribution<float>(0,1) under libstdc++ (g++ 13.3) is exactly
 * u = float(mt()) * 2^-32, and a result that rounds to 1.0f is replaced by nextafter(1,0).
 * Hard-coded so the stream does not depend on the standard library (MSVC's differs). */
extern "C" float orc_rng_uniform(orc_rng* r)
{
    const uint32_t raw = (uint32_t)r->gen();
    float u = (float)raw * 2.3283064365386963e-10f; /* 2^-32, exact power of two => same as the division */
    if (u >= 1.0f) u = 0.99999994f;                 /* nextafterf(1,0) */
    return u;
}
extern "C" float orc_rng_uniform_range(orc_rng* r, float lo, float hi)
{
    return lo + (hi - lo) * orc_rng_uniform(r); /* RandomUtilities.h:108-111 */
}
extern "C" int orc_rng_uniform_int(orc_rng* r, int lo, int hi)
{
    return (int)orc_rng_uniform_range(r, (float)lo, (float)hi); /* RandomUtilities.h:141-144 */
}
extern "C" int orc_selfcheck_rng(uint32_t seed, int ndraws)
{
    orc_rng a;
    a.gen.seed(seed);
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> dist(.0f, 1.0f); /* RandomUtilities.h:12,18 */
    int bad = 0;
    for (int i = 0; i < ndraws; ++i) {
        const float x = orc_rng_uniform(&a), y = dist(g);
        if (std::memcmp(&x, &y, 4) != 0) ++bad;
    }
    return bad;
}

/* ===================================================================================== grid model */

extern "C" void orc_decode_position(uint32_t index, const uint32_t dims[3], int decode_mode, uint32_t out[3])
{
    const uint32_t yz = dims[1] * dims[2];
    if (decode_mode == 0) {
        out[0] = index / yz;
        const uint32_t w = index % yz;
        out[1] = w / dims[2];
        out[2] = w % dims[2];
    } else {
        /* SH/Fracturer/voxel.glsl:6-14, float32 arithmetic as written (finding 6) */
        const float x = (float)index / (float)yz;
        const float w = (float)(index % yz);
        const float y = w / (float)dims[2];
        const float z = (float)((uint32_t)w % dims[2]);
        out[0] = (uint32_t)x;
        out[1] = (uint32_t)y;
        out[2] = (uint32_t)z;
    }
}

extern "C" void orc_dims_rule(const float mn[3], const float mx[3], uint32_t maxVoxels, uint32_t out[3])
{
    /* CADScene.cpp:545-556 */
    float size[3], maxSize = 0.0f;
    for (int i = 0; i < 3; ++i) {
        size[i] = mx[i] - mn[i];
        maxSize = std::max(maxSize, size[i]);
    }
    for (int i = 0; i < 3; ++i) {
        int v = (int)std::floor((float)maxVoxels * size[i] / maxSize);
        while (v % 4 != 0) ++v;
        while (v % 4 != 0 || v > (int)maxVoxels) --v;
        out[i] = (uint32_t)v;
    }
}

/* ===================================================================================== V2: SAT */

namespace {

struct SatData {
    float c[3], r[3];     /* box centre, half extent */
    float v0[3], v1[3], v2[3];
    float e0[3], e1[3], e2[3];
};

/* returns pass/fail of one edge-axis test and (optionally) accumulates the smallest relative slack */
inline bool axis_pass(float pa, float pb, float rad, float* margin)
{
    float mn, mx;
    if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
    if (margin) {
        const float scale = std::max(std::max(std::fabs(rad), std::fabs(mn)), std::max(std::fabs(mx), 1e-30f));
        const float s = std::min(std::fabs(rad - mn), std::fabs(mx + rad)) / scale;
        if (s < *margin) *margin = s;
    }
    return !(mn > rad || mx < -rad); /* Intersections3D.h:311,330,... strict compares */
}

/* Intersections3D.h:204-258 with helpers :260-420.  Operation order preserved; margin==nullptr => early exit. */
bool tri_box(const float p1[3], const float p2[3], const float p3[3], const float bmin[3], const float bmax[3], float* margin)
{
    SatData d;
    for (int i = 0; i < 3; ++i) {
        d.c[i] = (bmax[i] + bmin[i]) / 2.0f; /* AABB.h:41 center() */
        d.r[i] = bmax[i] - d.c[i];           /* AABB.h:51 extent() */
    }
    for (int i = 0; i < 3; ++i) {
        d.v0[i] = p1[i] - d.c[i];
        d.v1[i] = p2[i] - d.c[i];
        d.v2[i] = p3[i] - d.c[i];
    }
    for (int i = 0; i < 3; ++i) {
        d.e0[i] = d.v1[i] - d.v0[i];
        d.e1[i] = d.v2[i] - d.v1[i];
        d.e2[i] = d.v0[i] - d.v2[i];
    }
    bool ok = true;
    float fx, fy, fz, a, b;
#define FAIL_IF(cond) do { if (cond) { if (!margin) return false; ok = false; } } while (0)
    /* edge0: X01, Y02, Z12   (:223-228) */
    fx = std::fabs(d.e0[0]); fy = std::fabs(d.e0[1]); fz = std::fabs(d.e0[2]);
    a = d.e0[2]; b = d.e0[1]; /* xAxisTest_01 :296-315: p0 = a*v0.y - b*v0.z ; p2 = a*v2.y - b*v2.z ; rad = fa*r.y + fb*r.z */
    FAIL_IF(!axis_pass(a * d.v0[1] - b * d.v0[2], a * d.v2[1] - b * d.v2[2], fz * d.r[1] + fy * d.r[2], margin));
    a = d.e0[2]; b = d.e0[0]; /* yAxisTest_02 :338-357: p0 = -a*v0.x + b*v0.z ; p2 = -a*v2.x + b*v2.z ; rad = fa*r.x + fb*r.z */
    FAIL_IF(!axis_pass(-a * d.v0[0] + b * d.v0[2], -a * d.v2[0] + b * d.v2[2], fz * d.r[0] + fx * d.r[2], margin));
    a = d.e0[1]; b = d.e0[0]; /* zAxisTest_12 :380-399: p1 = a*v1.x - b*v1.y ; p2 = a*v2.x - b*v2.y ; rad = fa*r.x + fb*r.y */
    FAIL_IF(!axis_pass(a * d.v2[0] - b * d.v2[1], a * d.v1[0] - b * d.v1[1], fy * d.r[0] + fx * d.r[1], margin));
    /* edge1: X01, Y02, Z0    (:230-235) */
    fx = std::fabs(d.e1[0]); fy = std::fabs(d.e1[1]); fz = std::fabs(d.e1[2]);
    a = d.e1[2]; b = d.e1[1];
    FAIL_IF(!axis_pass(a * d.v0[1] - b * d.v0[2], a * d.v2[1] - b * d.v2[2], fz * d.r[1] + fy * d.r[2], margin));
    a = d.e1[2]; b = d.e1[0];
    FAIL_IF(!axis_pass(-a * d.v0[0] + b * d.v0[2], -a * d.v2[0] + b * d.v2[2], fz * d.r[0] + fx * d.r[2], margin));
    a = d.e1[1]; b = d.e1[0]; /* zAxisTest_0 :401-420: p0 = a*v0.x - b*v0.y ; p1 = a*v1.x - b*v1.y */
    FAIL_IF(!axis_pass(a * d.v0[0] - b * d.v0[1], a * d.v1[0] - b * d.v1[1], fy * d.r[0] + fx * d.r[1], margin));
    /* edge2: X2, Y1, Z12     (:237-242) */
    fx = std::fabs(d.e2[0]); fy = std::fabs(d.e2[1]); fz = std::fabs(d.e2[2]);
    a = d.e2[2]; b = d.e2[1]; /* xAxisTest_2 :317-336: p0 = a*v0.y - b*v0.z ; p1 = a*v1.y - b*v1.z */
    FAIL_IF(!axis_pass(a * d.v0[1] - b * d.v0[2], a * d.v1[1] - b * d.v1[2], fz * d.r[1] + fy * d.r[2], margin));
    a = d.e2[2]; b = d.e2[0]; /* yAxisTest_1 :359-378: p0 = -a*v0.x + b*v0.z ; p1 = -a*v1.x + b*v1.z */
    FAIL_IF(!axis_pass(-a * d.v0[0] + b * d.v0[2], -a * d.v1[0] + b * d.v1[2], fz * d.r[0] + fx * d.r[2], margin));
    a = d.e2[1]; b = d.e2[0];
    FAIL_IF(!axis_pass(a * d.v2[0] - b * d.v2[1], a * d.v1[0] - b * d.v1[1], fy * d.r[0] + fx * d.r[1], margin));
    /* box axes (:246-253, findMinMax :260-267) */
    for (int q = 0; q < 3; ++q) {
        float mn = d.v0[q], mx = d.v0[q];
        if (d.v1[q] < mn) mn = d.v1[q];
        if (d.v1[q] > mx) mx = d.v1[q];
        if (d.v2[q] < mn) mn = d.v2[q];
        if (d.v2[q] > mx) mx = d.v2[q];
        if (margin) {
            const float scale = std::max(std::max(std::fabs(d.r[q]), std::fabs(mn)), std::max(std::fabs(mx), 1e-30f));
            const float s = std::min(std::fabs(d.r[q] - mn), std::fabs(mx + d.r[q])) / scale;
            if (s < *margin) *margin = s;
        }
        FAIL_IF(mn > d.r[q] || mx < -d.r[q]);
    }
    /* plane/box (:256-257, planeBoxOverlap :269-294); normal = cross(edge0, edge1) (glm::cross order) */
    float n[3];
    n[0] = d.e0[1] * d.e1[2] - d.e1[1] * d.e0[2];
    n[1] = d.e0[2] * d.e1[0] - d.e1[2] * d.e0[0];
    n[2] = d.e0[0] * d.e1[1] - d.e1[0] * d.e0[1];
    float vmin[3], vmax[3];
    for (int q = 0; q < 3; ++q) {
        const float v = d.v0[q];
        if (n[q] > 0.0f) { vmin[q] = -d.r[q] - v; vmax[q] = d.r[q] - v; }
        else             { vmin[q] = d.r[q] - v;  vmax[q] = -d.r[q] - v; }
    }
    /* glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z */
    const float dmin = n[0] * vmin[0] + n[1] * vmin[1] + n[2] * vmin[2];
    const float dmax = n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2];
    if (margin) {
        const float scale = std::max(std::fabs(n[0] * d.r[0]) + std::fabs(n[1] * d.r[1]) + std::fabs(n[2] * d.r[2]), 1e-30f);
        const float s = std::min(std::fabs(dmin), std::fabs(dmax)) / scale;
        if (s < *margin) *margin = s;
    }
    FAIL_IF(dmin > 0.0f);
    FAIL_IF(!(dmax >= 0.0f));
#undef FAIL_IF
    return ok;
}

}  // namespace

extern "C" int orc_tri_box_intersect(const float p1[3], const float p2[3], const float p3[3], const float bmin[3], const float bmax[3])
{
    return tri_box(p1, p2, p3, bmin, bmax, nullptr) ? 1 : 0;
}

extern "C" int orc_voxelize_sat(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, const float amin[3],
                                const float amax[3], const uint32_t dims[3], uint16_t* grid, int clear, float* margin)
{
    const uint64_t N = (uint64_t)dims[0] * dims[1] * dims[2];
    if (clear) std::memset(grid, 0, N * sizeof(uint16_t));
    if (margin) for (uint64_t i = 0; i < N; ++i) margin[i] = 1e30f;
    /* RegularGrid.cpp:438 _cellSize = aabb.size() / vec3(numDivs) */
    float cell[3];
    for (int i = 0; i < 3; ++i) cell[i] = (amax[i] - amin[i]) / (float)dims[i];
    /* serial over triangles (writes are idempotent but margin is a min-reduction) */
    for (uint32_t f = 0; f < nf; ++f) {
        const uint32_t ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
        if (ia >= nv || ib >= nv || ic >= nv) return ORC_ERR_CAPACITY;
        const float* p1 = verts + 3 * ia;
        const float* p2 = verts + 3 * ib;
        const float* p3 = verts + 3 * ic;
        int lo[3], hi[3];
        for (int q = 0; q < 3; ++q) {
            const float tmn = std::min(p1[q], std::min(p2[q], p3[q])), tmx = std::max(p1[q], std::max(p2[q], p3[q]));
            /* conservative candidate range: one extra cell each side; the SAT predicate itself decides */
            lo[q] = clampi((int)std::floor((tmn - amin[q]) / cell[q]) - 1, 0, (int)dims[q] - 1);
            hi[q] = clampi((int)std::floor((tmx - amin[q]) / cell[q]) + 1, 0, (int)dims[q] - 1);
        }
        for (int x = lo[0]; x <= hi[0]; ++x)
            for (int y = lo[1]; y <= hi[1]; ++y)
                for (int z = lo[2]; z <= hi[2]; ++z) {
                    /* RegularGrid.cpp:258-259: min = aabb.min + cellSize * vec3(x,y,z); max = min + cellSize */
                    float bmin[3], bmax[3];
                    bmin[0] = amin[0] + cell[0] * (float)x; bmin[1] = amin[1] + cell[1] * (float)y; bmin[2] = amin[2] + cell[2] * (float)z;
                    for (int q = 0; q < 3; ++q) bmax[q] = bmin[q] + cell[q];
                    const uint64_t i = lin(x, y, z, dims);
                    if (margin) {
                        float m = margin[i];
                        if (tri_box(p1, p2, p3, bmin, bmax, &m)) grid[i] = ORC_VOXEL_FREE;
                        margin[i] = m;
                    } else if (grid[i] == ORC_VOXEL_EMPTY && tri_box(p1, p2, p3, bmin, bmax, nullptr)) {
                        grid[i] = ORC_VOXEL_FREE;
                    }
                }
    }
    return ORC_OK;
}

/* ---- V1: the live occupancy of RegularGrid::fill(Model3D*) (RegularGrid.cpp:173-212) = Tetravoxelizer (SRC/Graphics/Core/
 * Tetravoxelizer.{h,cpp}): every face + the vertex-average centroid is a tetrahedron in the grid AABB's NDC space
 * (initializeModel :198-247, scaleToNDC Tetravoxelizer.h:70-72, 4-element sort :75-81); for slice s the geometry shader (:42-92)
 * cuts each tetrahedron with the plane y = ySlice (ySlice starts at -1 and is ACCUMULATED in float32, compute :282-299) into one
 * or two triangles in (x, z), which the GL rasteriser draws into an X x Z R8UI target under glLogicOp(GL_XOR) (:271-272);
 * the slice is read back as uint8[y][z][x] (:303) and cells holding 1 become VOXEL_FREE (RegularGrid.cpp:186-199).
 * The FILTER_TETRAHEDRA_BY_Y path only drops tetrahedra that lie wholly below the next slice (its qsort comparator truncates
 * float differences to int, Tetravoxelizer.h:84-89, but dropping is conservative whatever the order), so it changes nothing.
 *
 * PARITY UNPINNED BY CONSTRUCTION: pixel coverage is decided by the GPU's rasteriser (sub-pixel snapping, fill rule on shared
 * edges, mix() contraction are implementation-defined in OpenGL).  The rule fixed here: shader arithmetic in float32 with
 * mix(a, b, t) = a*(1-t) + b*t un-fused; window coordinates xw = x*(X/2) + X/2 snapped to 1/256 pixel (round half to even);
 * a pixel centre is covered iff it is strictly inside, or on an edge whose direction (dx, dy) in the counter-clockwise
 * orientation has dy > 0 or (dy == 0 and dx < 0); zero-area triangles cover nothing.  Because neighbouring tetrahedra
 * interpolate a shared edge from the same (lower-y, higher-y) vertex pair, their cross-sections share exact vertices and the
 * XOR is the parity of the tetrahedra containing the sample point (cell centre in x and z, slice plane in y). */
namespace {
struct SolidPt { int64_t x, y; };
inline int64_t solid_snap(float w)
{
    float f = std::nearbyintf(w * 256.0f); /* default rounding mode: to nearest even */
    if (!(f > -268435456.0f)) f = -268435456.0f; /* also catches NaN */
    if (f > 268435456.0f) f = 268435456.0f;
    return (int64_t)f;
}
inline int64_t floor_shift8(int64_t v) { return v >> 8; } /* arithmetic shift = floor(v / 256) */
/* XOR the pixels of one triangle into slice[z * X + x] */
void solid_raster(uint8_t* slice, int X, int Z, SolidPt a, SolidPt b, SolidPt c)
{
    int64_t area2 = (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
    if (area2 == 0) return;
    if (area2 < 0) std::swap(b, c);
    const SolidPt v[3] = { a, b, c };
    int64_t dx[3], dy[3];
    int own[3];
    for (int e = 0; e < 3; ++e) {
        dx[e] = v[(e + 1) % 3].x - v[e].x, dy[e] = v[(e + 1) % 3].y - v[e].y;
        own[e] = (dy[e] > 0 || (dy[e] == 0 && dx[e] < 0)) ? 1 : 0;
    }
    const int64_t minx = std::min(a.x, std::min(b.x, c.x)), maxx = std::max(a.x, std::max(b.x, c.x));
    const int64_t miny = std::min(a.y, std::min(b.y, c.y)), maxy = std::max(a.y, std::max(b.y, c.y));
    /* pixel i has its centre at 256 i + 128 */
    const int i0 = (int)std::max<int64_t>(0, floor_shift8(minx - 128 + 255)), i1 = (int)std::min<int64_t>(X - 1, floor_shift8(maxx - 128));
    const int k0 = (int)std::max<int64_t>(0, floor_shift8(miny - 128 + 255)), k1 = (int)std::min<int64_t>(Z - 1, floor_shift8(maxy - 128));
    for (int k = k0; k <= k1; ++k)
        for (int i = i0; i <= i1; ++i) {
            const int64_t px = 256 * (int64_t)i + 128, py = 256 * (int64_t)k + 128;
            bool in = true;
            for (int e = 0; e < 3 && in; ++e) in = dx[e] * (py - v[e].y) - dy[e] * (px - v[e].x) + own[e] > 0;
            if (in) slice[(size_t)k * X + i] ^= 1;
        }
}
}  // namespace

extern "C" int orc_voxelize_solid(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, const float amin[3],
                                  const float amax[3], const uint32_t dims[3], uint16_t* grid, int clear)
{
    const int X = (int)dims[0], Y = (int)dims[1], Z = (int)dims[2];
    const uint64_t N = (uint64_t)X * Y * Z;
    if (clear) std::memset(grid, 0, N * sizeof(uint16_t));
    if (!nv || !nf) return ORC_OK;
    /* initializeModel :204-217 */
    float cen[3] = { 0.0f, 0.0f, 0.0f };
    for (uint32_t i = 0; i < nv; ++i)
        for (int q = 0; q < 3; ++q) cen[q] += verts[3 * i + q];
    for (int q = 0; q < 3; ++q) cen[q] /= (float)nv;
    float dim[3], ctr[3];
    for (int q = 0; q < 3; ++q) dim[q] = amax[q] - amin[q], ctr[q] = 0.5f * (amin[q] + amax[q]);
    auto ndc = [&](const float* v, float* o) { /* Tetravoxelizer.h:70-72: 2.0f * (v - centre) / dim */
        for (int q = 0; q < 3; ++q) o[q] = (2.0f * (v[q] - ctr[q])) / dim[q];
    };
    float cn[3];
    ndc(cen, cn);
    struct Tet { float v[4][3]; };
    std::vector<Tet> tets(nf);
    for (uint32_t f = 0; f < nf; ++f) {
        Tet& t = tets[f];
        for (int k = 0; k < 3; ++k) {
            if (faces[3 * f + k] >= nv) return ORC_ERR_CAPACITY;
            ndc(verts + 3 * (size_t)faces[3 * f + k], t.v[k]);
        }
        for (int q = 0; q < 3; ++q) t.v[3][q] = cn[q];
        auto sw = [&](int a, int b) { /* sort4Vec3ByLowerY, Tetravoxelizer.h:75-81 */
            if (t.v[a][1] > t.v[b][1])
                for (int q = 0; q < 3; ++q) std::swap(t.v[a][q], t.v[b][q]);
        };
        sw(0, 1), sw(2, 3), sw(0, 2), sw(1, 3), sw(1, 2);
    }
    /* compute :282-304 */
    std::vector<float> ys(Y);
    {
        float ySlice = -1.0f;
        const float yStep = 2.0f / (float)Y;
        for (int s = 0; s < Y; ++s) ys[s] = ySlice, ySlice += yStep;
    }
    const float hx = (float)X * 0.5f, hz = (float)Z * 0.5f;
#pragma omp parallel
    {
        std::vector<uint8_t> slice((size_t)X * Z);
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s < Y; ++s) {
            std::fill(slice.begin(), slice.end(), 0);
            const float sl = ys[s];
            for (const Tet& t : tets) {
                const float *A = t.v[0], *B = t.v[1], *C = t.v[2], *D = t.v[3];
                if (!(A[1] < sl && sl <= D[1])) continue;
                auto interp = [&](const float* p, const float* q) { /* INTERP, :46: mix(p, q, (s - p.y) / (q.y - p.y)).xz -> window, snapped */
                    const float w = (sl - p[1]) / (q[1] - p[1]);
                    const float x = p[0] * (1.0f - w) + q[0] * w, z = p[2] * (1.0f - w) + q[2] * w;
                    return SolidPt{ solid_snap(x * hx + hx), solid_snap(z * hz + hz) };
                };
                const SolidPt p0 = interp(A, D);
                const SolidPt v1 = (sl <= B[1]) ? interp(A, B) : interp(B, D);
                const SolidPt v2 = (sl <= C[1]) ? interp(A, C) : interp(C, D);
                solid_raster(slice.data(), X, Z, p0, v1, v2);
                if (B[1] < sl && sl <= C[1]) solid_raster(slice.data(), X, Z, interp(B, C), v2, v1);
            }
            /* RegularGrid.cpp:186-199: result[y][z][x] == 1 -> set(x, y, z, VOXEL_FREE) */
            for (int z = 0; z < Z; ++z)
                for (int x = 0; x < X; ++x)
                    if (slice[(size_t)z * X + x] == 1) grid[lin(x, s, z, dims)] = ORC_VOXEL_FREE;
        }
    }
    return ORC_OK;
}

/* ===================================================================================== S1/S2: seeding */

namespace {
/* RegularGrid.cpp:543-559 isBoundary(x,y,z,neighbourhoodSize=1) */
bool grid_is_boundary(const uint16_t* g, const uint32_t d[3], int x, int y, int z)
{
    const int nb = 1;
    const int x0 = clampi(x - nb, 0, (int)d[0] - 1), x1 = clampi(x + nb, 0, (int)d[0] - 1);
    const int y0 = clampi(y - nb, 0, (int)d[1] - 1), y1 = clampi(y + nb, 0, (int)d[1] - 1);
    const int z0 = clampi(z - nb, 0, (int)d[2] - 1), z1 = clampi(z + nb, 0, (int)d[2] - 1);
    for (int a = x0; a <= x1; ++a)
        for (int b = y0; b <= y1; ++b)
            for (int c = z0; c <= z1; ++c)
                if (g[lin(a, b, c, d)] == ORC_VOXEL_EMPTY) return true;
    return false;
}
struct U3Less {
    bool operator()(const U3& l, const U3& r) const
    {
        if (l.x != r.x) return l.x < r.x;
        if (l.y != r.y) return l.y < r.y;
        return l.z < r.z;
    }
};
}  // namespace

/* Halton_sampler::sample for the three dimensions Seeder::uniform uses (HaltonSampler.h:625-632), after init_faure (:572-602):
 * dimension 0 = base-2 radical inverse written into a float mantissa (:1416-1430); dimensions 1, 2 = radical inverses in base 3 / 5
 * over 20 / 12 digits with the Faure digit permutation (tables m_perm3 / m_perm5 of :903-906 hold 5 / 3 digits each and are
 * combined most-significant first, :1432-1446), scaled by float(0x1.fffffcp-1 / base^digits).  The permutation is built by
 * Faure's recursion (:576-600): identity up to base 3; even b: [2 s(b/2), 2 s(b/2) + 1]; odd b: s(b-1) with the values >= b/2
 * shifted up and b/2 inserted in the middle. */
static std::vector<unsigned> faure_permutation(unsigned base)
{
    std::vector<unsigned> p(base);
    if (base <= 3) {
        for (unsigned i = 0; i < base; ++i) p[i] = i;
        return p;
    }
    const unsigned half = base / 2;
    if (base % 2 == 0) {
        const std::vector<unsigned> q = faure_permutation(half);
        for (unsigned i = 0; i < half; ++i) p[i] = 2 * q[i], p[half + i] = 2 * q[i] + 1;
    } else {
        const std::vector<unsigned> q = faure_permutation(base - 1);
        for (unsigned i = 0; i + 1 < base; ++i) p[i + (i >= half ? 1 : 0)] = q[i] + (q[i] >= half ? 1 : 0);
        p[half] = half;
    }
    return p;
}

extern "C" float orc_halton(unsigned dimension, unsigned index)
{
    if (dimension == 0) {
        unsigned rev = 0;
        for (int b = 0; b < 32; ++b) rev |= ((index >> b) & 1u) << (31 - b);
        const unsigned bits = 0x3f800000u | (rev >> 9);
        float f;
        std::memcpy(&f, &bits, 4);
        return f - 1.f;
    }
    const unsigned base = dimension == 1 ? 3u : 5u, digits = dimension == 1 ? 20u : 12u;
    static const std::vector<unsigned> perm3 = faure_permutation(3), perm5 = faure_permutation(5);
    const std::vector<unsigned>& perm = dimension == 1 ? perm3 : perm5;
    unsigned numerator = 0; /* digit-reversed, permuted; < base^digits < 2^32 */
    for (unsigned d = 0; d < digits; ++d) {
        numerator = numerator * base + perm[index % base];
        index /= base;
    }
    const float scale = dimension == 1 ? float(0x1.fffffcp-1 / 3486784401u) : float(0x1.fffffcp-1 / 244140625u);
    return numerator * scale;
}

extern "C" int orc_seed_uniform(orc_rng* rng, const uint16_t* grid, const uint32_t dims[3], uint32_t n, int random_mode,
                                int location, uint32_t* out, uint32_t* attempts_out)
{
    /* Seeder.cpp:154-208 */
    if (random_mode != ORC_STD_UNIFORM && random_mode != ORC_HALTON) return ORC_ERR_UNSUPPORTED; /* BOOST_NORMAL: parity unpinned, SURVEY §8a S3 */
    std::set<U3, U3Less> seeds;
    const int nd[3] = { (int)dims[0] - 2, (int)dims[1] - 2, (int)dims[2] - 2 }; /* :165 numDivs - 2 */
    const uint32_t MAX_TRIES = 1000000;                                         /* Seeder.h:48 */
    uint32_t attempt = 0;
    while (seeds.size() != n) {
        if (attempt == MAX_TRIES) {
            if (attempts_out) *attempts_out = attempt;
            return ORC_ERR_SEEDER_EXHAUSTED; /* :173-174 SeederSearchError */
        }
        /* :177-179 — three draws per attempt, consumed even when rejected */
        int x, y, z;
        if (random_mode == ORC_HALTON) { /* Seeder.cpp:28: int(sample(coord, attempt) * (max - min) + min), float32; no generator state */
            x = (int)(orc_halton(0, attempt) * (float)(nd[0] + 1) + 0.0f);
            y = (int)(orc_halton(1, attempt) * (float)(nd[1] + 1) + 0.0f);
            z = (int)(orc_halton(2, attempt) * (float)(nd[2] + 1) + 0.0f);
        } else {
            x = orc_rng_uniform_int(rng, 0, nd[0] + 1);
            y = orc_rng_uniform_int(rng, 0, nd[1] + 1);
            z = orc_rng_uniform_int(rng, 0, nd[2] + 1);
        }
        const U3 v = { (uint32_t)x, (uint32_t)y, (uint32_t)z };
        const bool occupied = grid[lin(x, y, z, dims)] != ORC_VOXEL_EMPTY; /* RegularGrid.cpp:561-564 */
        const bool boundary = grid_is_boundary(grid, dims, x, y, z);
        const bool isFree = seeds.find(v) == seeds.end();
        if (occupied && isFree)
            if ((location == ORC_OUTER && boundary) || (location == ORC_INNER && !boundary) || location == ORC_BOTH) seeds.insert(v);
        ++attempt;
    }
    uint32_t label = ORC_VOXEL_FREE + 1, k = 0; /* :202 ids start at 2 */
    for (const U3& s : seeds) {              /* std::set order: lexicographic x,y,z */
        out[4 * k + 0] = s.x; out[4 * k + 1] = s.y; out[4 * k + 2] = s.z; out[4 * k + 3] = label++;
        ++k;
    }
    if (attempts_out) *attempts_out = attempt;
    return ORC_OK;
}

extern "C" void orc_merge_seeds(const uint32_t* frags, uint32_t nfrags, uint32_t* seeds, uint32_t nseeds, int dfunc)
{
    /* Seeder.cpp:115-152 */
    std::vector<int> idFragment(1u << ORC_ID_POSITION, 0);
    for (uint32_t si = 0; si < nseeds; ++si) {
        uint32_t* seed = seeds + 4 * si;
        float mn = FLT_MAX;
        int nearest = -1;
        for (uint32_t i = 0; i < nfrags; ++i) {
            const float fx = (float)frags[4 * i], fy = (float)frags[4 * i + 1], fz = (float)frags[4 * i + 2];
            const float sx = (float)seed[0], sy = (float)seed[1], sz = (float)seed[2];
            float dist = .0f;
            switch (dfunc) {
            case ORC_EUCLIDEAN: {
                const float dx = sx - fx, dy = sy - fy, dz = sz - fz;
                dist = std::sqrt(dx * dx + dy * dy + dz * dz); /* glm::distance = sqrt(dot(d,d)) */
                break;
            }
            case ORC_MANHATTAN: dist = std::fabs(sx - fx) + std::fabs(sy - fy) + std::fabs(sz - fz); break;
            case ORC_CHEBYSHEV: dist = std::max(std::fabs(sx - fx), std::max(std::fabs(sy - fy), std::fabs(sz - fz))); break;
            }
            if (dist < mn) { mn = dist; nearest = (int)i; }
        }
        const uint32_t fw = frags[4 * nearest + 3];
        seed[3] = fw | ((uint32_t)(++idFragment[fw & 0xFFu]) << ORC_ID_POSITION); /* :150 */
    }
}

extern "C" int orc_make_seeds(orc_rng* rng, const uint16_t* grid, const uint32_t dims[3], uint32_t n, uint32_t n_extra,
                              int random_mode, int merge_dfunc, uint32_t* out, uint32_t cap)
{
    /* CADScene.cpp:626-655 (numImpacts == 0 branch) */
    const uint32_t total = n + (n_extra ? n + n_extra : 0);
    if (total > cap) return ORC_ERR_CAPACITY;
    int rc = orc_seed_uniform(rng, grid, dims, n, random_mode, ORC_OUTER, out, nullptr);
    if (rc != ORC_OK) return rc;
    if (n_extra > 0) {
        std::vector<uint32_t> extra(4 * (size_t)(n + n_extra));
        rc = orc_seed_uniform(rng, grid, dims, n_extra, random_mode, ORC_BOTH, extra.data() + 4 * n, nullptr);
        if (rc != ORC_OK) return rc;
        std::memcpy(extra.data(), out, 16 * (size_t)n);              /* :651 originals prepended */
        orc_merge_seeds(out, n, extra.data(), n + n_extra, merge_dfunc); /* :653 */
        std::memcpy(out + 4 * n, extra.data(), 16 * (size_t)(n + n_extra)); /* :654 */
    }
    return (int)total;
}

/* ---- S3: Seeder::nearSeeds (Seeder.cpp:49-113): impact-biased seeds.  Two generators are involved: RandomUtilities' mt19937 picks the
 * impacted fragment and the seed count per impact (:67-68), and C rand() drives getBiasedRandomInt (RandomUtilities.h:146-156),
 * seeded by srand(_fractParameters._seed) at CADScene.cpp:36.  rand() is the C runtime's: the reference is an MSVC program, whose
 * rand() is the documented LCG  state = state * 214013 + 2531011; return (state >> 16) & 0x7fff  (crand_mode 0, *crand_state is
 * the LCG word).  crand_mode 1 calls this process's ::rand() instead, which lets the test pin every other line of the function
 * against the reference's own Seeder.cpp compiled in place (both then consume the same glibc stream).
 * The reference loops without bound when no candidate qualifies; the restatement gives up after 1e6 candidates per impact. */
namespace {
inline int crand_next(int mode, uint32_t* state)
{
    if (mode == 1) return std::rand();
    *state = *state * 214013u + 2531011u;
    return (int)((*state >> 16) & 0x7fffu);
}
}  // namespace

extern "C" int orc_near_seeds(orc_rng* rng, int crand_mode, uint32_t* crand_state, const uint16_t* grid, const uint32_t dims[3], const uint32_t* frags,
                              uint32_t nfrags, uint32_t numImpacts, uint32_t numSeeds, uint32_t spreading, uint32_t* out, uint32_t cap)
{
    if (!nfrags || !spreading) return ORC_ERR_UNSUPPORTED;
    for (int q = 0; q < 3; ++q)
        if ((int)dims[q] / (int)spreading == 0) return ORC_ERR_UNSUPPORTED; /* rand() % 0 in getBiasedRandomInt */
    std::set<std::array<uint32_t, 3>> seeds; /* lexicographic comparator, :52-59 */
    unsigned numPendingSeeds = numSeeds;
    const uint32_t half[3] = { dims[0] / 2, dims[1] / 2, dims[2] / 2 };
    const unsigned minDiv = std::min(dims[0], std::min(dims[1], dims[2])) / 2;
    auto biased = [&](int mx) { /* RandomUtilities.h:146-156 with min = 0 */
        int number = 0;
        mx /= (int)spreading;
        for (uint32_t i = 0; i < spreading; ++i) number += crand_next(crand_mode, crand_state) % mx;
        return number;
    };
    for (uint32_t idx = 0; idx < numImpacts; ++idx) {
        const uint32_t* frag = frags + 4 * (size_t)orc_rng_uniform_int(rng, 0, (int)nfrags - 1);
        const unsigned nseeds = (unsigned)orc_rng_uniform_int(rng, 1, (int)numPendingSeeds);
        unsigned currentSeeds = 0;
        uint64_t tries = 0;
        while (currentSeeds != nseeds) {
            if (++tries > 1000000) return ORC_ERR_SEEDER_EXHAUSTED;
            int x = (int)(half[0] - (uint32_t)biased((int)dims[0]));
            int y = (int)(half[1] - (uint32_t)biased((int)dims[1]));
            int z = (int)(half[2] - (uint32_t)biased((int)dims[2]));
            x = (int)((frag[0] + (uint32_t)x + dims[0]) % dims[0]);
            y = (int)((frag[1] + (uint32_t)y + dims[1]) % dims[1]);
            z = (int)((frag[2] + (uint32_t)z + dims[2]) % dims[2]);
            const float dx = (float)x - (float)frag[0], dy = (float)y - (float)frag[1], dz = (float)z - (float)frag[2];
            if (std::sqrt(dx * dx + dy * dy + dz * dz) > (float)minDiv) continue; /* :87 */
            const bool occupied = grid[lin(x, y, z, dims)] != ORC_VOXEL_EMPTY;
            const bool isFree = seeds.find({ (uint32_t)x, (uint32_t)y, (uint32_t)z }) == seeds.end();
            const bool isBoundary = grid_is_boundary(grid, dims, x, y, z);
            if (occupied && isFree && isBoundary) seeds.insert({ (uint32_t)x, (uint32_t)y, (uint32_t)z }), ++currentSeeds;
        }
        numPendingSeeds -= nseeds;
    }
    const uint32_t total = nfrags + (uint32_t)seeds.size();
    if (total > cap) return ORC_ERR_CAPACITY;
    std::memcpy(out, frags, 16 * (size_t)nfrags); /* :103 result = frags */
    uint32_t nseed = frags[4 * (size_t)(nfrags - 1) + 3], k = nfrags;
    for (const auto& sd : seeds) out[4 * k] = sd[0], out[4 * k + 1] = sd[1], out[4 * k + 2] = sd[2], out[4 * k + 3] = ++nseed, ++k;
    return (int)total;
}

/* ===================================================================================== F1: naive */

extern "C" void orc_naive(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc, int decode_mode)
{
    /* NaiveFracturer.cpp:26-68 (buildCPU) == naiveFracturer-comp.glsl:19-43; distances NaiveFracturer.cpp:12-23 */
    const int64_t N = (int64_t)dims[0] * dims[1] * dims[2];
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        if (grid[idx] == ORC_VOXEL_EMPTY) continue;
        uint32_t p[3];
        orc_decode_position((uint32_t)idx, dims, decode_mode, p);
        float minDistance = FLT_MAX;
        uint16_t value = grid[idx];
        for (uint32_t s = 0; s < nseeds; ++s) {
            const int dx = (int)p[0] - (int)seeds[4 * s], dy = (int)p[1] - (int)seeds[4 * s + 1], dz = (int)p[2] - (int)seeds[4 * s + 2];
            float dist;
            if (dfunc == ORC_EUCLIDEAN) {
                const float x = (float)dx, y = (float)dy, z = (float)dz;
                dist = std::sqrt(x * x + y * y + z * z);
            } else if (dfunc == ORC_MANHATTAN) {
                dist = (float)(std::abs(dx) + std::abs(dy) + std::abs(dz));
            } else {
                dist = (float)std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz)));
            }
            if (dist < minDistance) { /* strict: lowest seed index wins ties */
                minDistance = dist;
                value = (uint16_t)seeds[4 * s + 3];
            }
        }
        grid[idx] = value;
    }
}

/* ===================================================================================== F2/F3: flood */

namespace {

const uint32_t KEY_WALL = 0xFFFFFFFFu, KEY_UNREACHED = 0xFFFFFFFEu;
const int KEY_DIST_SHIFT = 15;
const uint32_t KEY_MAX_DIST = (1u << 17) - 2;

struct FloodCtx {
    const uint32_t* dims;
    int nneigh;
    const int (*nb)[3];
    uint64_t N;
};

inline bool inside(const uint32_t d[3], int x, int y, int z)
{
    return x >= 0 && y >= 0 && z >= 0 && x < (int)d[0] && y < (int)d[1] && z < (int)d[2];
}
inline void delin(uint64_t i, const uint32_t d[3], int& x, int& y, int& z)
{
    const uint64_t yz = (uint64_t)d[1] * d[2];
    x = (int)(i / yz);
    const uint64_t w = i % yz;
    y = (int)(w / d[2]);
    z = (int)(w % d[2]);
}

/* One flood phase (inner `while (stackSize > 0)` loop, FloodFracturer.cpp:143-158) under the deterministic rule of
 * every FREE cell takes the word of the source minimising (geodesic distance, order).
 * ord[] holds the order of each labelled cell (index into the seeds vector), word_of_order[] the 16-bit word.
 * dist_out (optional) receives the level at which each cell was claimed. */
uint32_t flood_phase_levels(const FloodCtx& c, uint16_t* g, int32_t* ord, const std::vector<uint16_t>& word_of_order,
                            std::vector<uint64_t>& frontier, uint32_t* dist_out)
{
    std::vector<int32_t> cand(c.N, INT32_MAX);
    std::vector<uint64_t> next;
    uint32_t level = 0;
    while (!frontier.empty()) {
        next.clear();
        for (uint64_t v : frontier) {
            int x, y, z;
            delin(v, c.dims, x, y, z);
            for (int k = 0; k < c.nneigh; ++k) {
                const int nx = x + c.nb[k][0], ny = y + c.nb[k][1], nz = z + c.nb[k][2];
                if (!inside(c.dims, nx, ny, nz)) continue; /* floodFracturer-comp.glsl:37-39 */
                const uint64_t n = lin(nx, ny, nz, c.dims);
                if (g[n] != ORC_VOXEL_FREE) continue;       /* :36,41 */
                if (cand[n] == INT32_MAX) next.push_back(n);
                cand[n] = std::min(cand[n], ord[v]);        /* all same-level parents compete; lowest order wins */
            }
        }
        ++level;
        for (uint64_t n : next) { /* commit the level */
            ord[n] = cand[n];
            g[n] = word_of_order[cand[n]];
            cand[n] = INT32_MAX;
            if (dist_out) dist_out[n] = level;
        }
        frontier.swap(next);
    }
    return level;
}

/* Same fixed point computed as Dijkstra over (dist, order) keys — SURVEY finding 5 equivalence. */
uint32_t flood_phase_dijkstra(const FloodCtx& c, uint16_t* g, int32_t* ord, const std::vector<uint16_t>& word_of_order,
                              std::vector<uint64_t>& frontier, uint32_t* dist_out)
{
    typedef std::pair<uint64_t, uint64_t> QE; /* (dist<<32 | order, cell) */
    std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
    std::vector<uint64_t> key(c.N, UINT64_MAX);
    for (uint64_t v : frontier) {
        key[v] = (uint64_t)(uint32_t)ord[v];
        pq.push(QE(key[v], v));
    }
    uint32_t maxd = 0;
    while (!pq.empty()) {
        const QE e = pq.top();
        pq.pop();
        if (e.first != key[e.second]) continue;
        int x, y, z;
        delin(e.second, c.dims, x, y, z);
        for (int k = 0; k < c.nneigh; ++k) {
            const int nx = x + c.nb[k][0], ny = y + c.nb[k][1], nz = z + c.nb[k][2];
            if (!inside(c.dims, nx, ny, nz)) continue;
            const uint64_t n = lin(nx, ny, nz, c.dims);
            if (g[n] != ORC_VOXEL_FREE && key[n] == UINT64_MAX) continue; /* walls / pre-labelled non-sources */
            const uint64_t nk = e.first + (1ull << 32);
            if (nk < key[n]) {
                key[n] = nk;
                pq.push(QE(nk, n));
            }
        }
    }
    for (uint64_t i = 0; i < c.N; ++i) {
        if (key[i] == UINT64_MAX || (key[i] >> 32) == 0) continue;
        const int32_t o = (int32_t)(key[i] & 0xFFFFFFFFu);
        ord[i] = o;
        g[i] = word_of_order[o];
        const uint32_t dd = (uint32_t)(key[i] >> 32);
        if (dist_out) dist_out[i] = dd;
        maxd = std::max(maxd, dd);
    }
    frontier.clear();
    return maxd + 1;
}

/* prefix merge inside one fragment id (floodFracturer-comp.glsl:49-63), taken to its fixed point:
 * every connected (under the flood neighbourhood) set of cells with equal fragment id converges to the
 * lowest prefix present in it. */
void merge_prefixes(const FloodCtx& c, uint16_t* g)
{
    std::vector<uint8_t> seen(c.N, 0);
    std::vector<uint64_t> comp, stack;
    for (uint64_t s = 0; s < c.N; ++s) {
        if (seen[s] || g[s] <= ORC_VOXEL_FREE) continue;
        const uint16_t frag = g[s] & 0xFF;
        uint16_t minp = 0xFFFF;
        comp.clear();
        stack.clear();
        stack.push_back(s);
        seen[s] = 1;
        while (!stack.empty()) {
            const uint64_t v = stack.back();
            stack.pop_back();
            comp.push_back(v);
            minp = std::min<uint16_t>(minp, g[v] >> ORC_ID_POSITION);
            int x, y, z;
            delin(v, c.dims, x, y, z);
            for (int k = 0; k < c.nneigh; ++k) {
                const int nx = x + c.nb[k][0], ny = y + c.nb[k][1], nz = z + c.nb[k][2];
                if (!inside(c.dims, nx, ny, nz)) continue;
                const uint64_t n = lin(nx, ny, nz, c.dims);
                if (seen[n] || g[n] <= ORC_VOXEL_FREE || (g[n] & 0xFF) != frag) continue;
                seen[n] = 1;
                stack.push_back(n);
            }
        }
        const uint16_t w = (uint16_t)(frag | (minp << ORC_ID_POSITION));
        for (uint64_t v : comp) g[v] = w;
    }
}

int flood_impl(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc, int id_bits, int algo,
               orc_flood_stats* stats, uint32_t* keys_out)
{
    FloodCtx c;
    c.dims = dims;
    c.N = (uint64_t)dims[0] * dims[1] * dims[2];
    c.nneigh = (dfunc == ORC_MANHATTAN) ? 6 : 26; /* FloodFracturer.cpp:114 */
    c.nb = (dfunc == ORC_MANHATTAN) ? VON_NEUMANN : MOORE;
    if (nseeds > 32767) return ORC_ERR_CAPACITY;

    orc_homogenize(grid, c.N); /* :99 */
    std::vector<int32_t> ord(c.N, -1);
    std::vector<uint16_t> word_of_order(nseeds);
    std::vector<uint64_t> frontier;
    for (uint32_t s = 0; s < nseeds; ++s) { /* :102-103 later seeds overwrite earlier ones on the same cell */
        const uint64_t i = lin(seeds[4 * s], seeds[4 * s + 1], seeds[4 * s + 2], dims);
        word_of_order[s] = (uint16_t)seeds[4 * s + 3];
        if (ord[i] < 0) frontier.push_back(i);
        ord[i] = (int32_t)s;
        grid[i] = word_of_order[s];
    }
    std::vector<uint32_t> dist;
    if (keys_out || stats) dist.assign(c.N, 0);
    orc_flood_stats st = { 0, 0, 0, 0 };

    uint32_t numDisjoint = nseeds; /* :133 */
    bool first = true;
    while (numDisjoint != 0) {      /* :135 */
        uint32_t* dptr = (first && !dist.empty()) ? dist.data() : nullptr;
        st.levels += algo == 0 ? flood_phase_levels(c, grid, ord.data(), word_of_order, frontier, dptr)
                               : flood_phase_dijkstra(c, grid, ord.data(), word_of_order, frontier, dptr);
        ++st.rounds;
        if (first && keys_out) {
            for (uint64_t i = 0; i < c.N; ++i) {
                if (grid[i] == ORC_VOXEL_EMPTY) keys_out[i] = KEY_WALL;
                else if (ord[i] < 0) keys_out[i] = KEY_UNREACHED;
                else {
                    if (dist[i] > KEY_MAX_DIST) return ORC_ERR_CAPACITY;
                    keys_out[i] = (dist[i] << KEY_DIST_SHIFT) | (uint32_t)ord[i];
                }
            }
            return ORC_OK;
        }
        if (first && !dist.empty()) st.max_dist = *std::max_element(dist.begin(), dist.end());
        first = false;
        if (id_bits != 8) break; /* extension (finding 7): no prefix field, nothing to dissolve */

        merge_prefixes(c, grid);
        /* disjointSet-comp.glsl:17-24 */
        uint32_t minPrefix[1u << ORC_ID_POSITION];
        for (uint32_t& m : minPrefix) m = UINT32_MAX;
        for (uint64_t i = 0; i < c.N; ++i)
            if (grid[i] > ORC_VOXEL_FREE) minPrefix[grid[i] & 0xFF] = std::min<uint32_t>(minPrefix[grid[i] & 0xFF], grid[i] >> ORC_ID_POSITION);
        /* disjointSetStack-comp.glsl:20-37 */
        numDisjoint = 0;
        frontier.clear();
        /* order of a surviving word = lowest seed index carrying that word ("lowest seed index" rule, round >= 2) */
        std::vector<int32_t> order_of_word(65536, -1);
        for (int32_t s = (int32_t)nseeds - 1; s >= 0; --s) order_of_word[word_of_order[s]] = s;
        for (uint64_t i = 0; i < c.N; ++i) {
            if (grid[i] <= ORC_VOXEL_FREE) continue;
            if ((uint32_t)(grid[i] >> ORC_ID_POSITION) != minPrefix[grid[i] & 0xFF]) {
                grid[i] = ORC_VOXEL_FREE;
                ord[i] = -1;
                ++numDisjoint;
            } else {
                ord[i] = order_of_word[grid[i]];
                frontier.push_back(i);
            }
        }
        st.freed_voxels += numDisjoint;
    }
    if (id_bits == 8) orc_undo_mask(grid, c.N, ORC_ID_POSITION, 1); /* :180-186 unmaskRightMost(8) */
    if (stats) *stats = st;
    return ORC_OK;
}

}  // namespace

extern "C" int orc_flood(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc, int id_bits,
                         int algo, orc_flood_stats* stats)
{
    return flood_impl(grid, dims, seeds, nseeds, dfunc, id_bits, algo, stats, nullptr);
}

extern "C" int orc_flood_keys(const uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc,
                              uint32_t* keys)
{
    const uint64_t N = (uint64_t)dims[0] * dims[1] * dims[2];
    std::vector<uint16_t> g(grid, grid + N);
    return flood_impl(g.data(), dims, seeds, nseeds, dfunc, 15, 0, nullptr, keys);
}

extern "C" uint64_t orc_relax_keys_slab(uint32_t* keys, uint32_t XS, uint32_t Y, uint32_t Z, int nneigh)
{
    /* chaotic relaxation key(v) = min(key(v), min_nbr key + 1 level) on the interior planes 1..XS-2; halo planes fixed.
     * Converges to the unique least fixed point (monotone, bounded below). */
    const uint32_t d[3] = { XS, Y, Z };
    const int (*nb)[3] = nneigh == 6 ? VON_NEUMANN : MOORE;
    std::vector<uint8_t> changed_any((size_t)XS * Y * Z, 0);
    uint64_t changed_cells = 0;
    bool changed = true;
    while (changed) {
        changed = false;
        for (int pass = 0; pass < 2; ++pass) {
            for (int64_t t = 0; t < (int64_t)(XS - 2) * Y * Z; ++t) {
                const int64_t u = pass == 0 ? t : (int64_t)(XS - 2) * Y * Z - 1 - t;
                const uint64_t i = (uint64_t)Y * Z + (uint64_t)u;
                if (keys[i] == KEY_WALL) continue;
                int x, y, z;
                delin(i, d, x, y, z);
                uint32_t best = keys[i];
                for (int k = 0; k < nneigh; ++k) {
                    const int nx = x + nb[k][0], ny = y + nb[k][1], nz = z + nb[k][2];
                    if (!inside(d, nx, ny, nz)) continue;
                    const uint32_t nk = keys[lin(nx, ny, nz, d)];
                    if (nk >= KEY_UNREACHED) continue;
                    const uint32_t cnd = nk + (1u << KEY_DIST_SHIFT);
                    if (cnd < best) best = cnd;
                }
                if (best < keys[i]) {
                    keys[i] = best;
                    changed = true;
                    if (!changed_any[i]) { changed_any[i] = 1; ++changed_cells; }
                }
            }
        }
    }
    return changed_cells;
}

/* ===================================================================================== C1..C4 */

extern "C" void orc_remove_isolated_regions_cpu(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds)
{
    /* NaiveFracturer.cpp:111-150 */
    const uint64_t N = (uint64_t)dims[0] * dims[1] * dims[2];
    std::vector<uint16_t> newGrid(N, ORC_VOXEL_EMPTY);
    struct V4 { uint32_t x, y, z, w; };
    std::deque<V4> front;
    for (uint32_t s = 0; s < nseeds; ++s) {
        front.push_back({ seeds[4 * s], seeds[4 * s + 1], seeds[4 * s + 2], seeds[4 * s + 3] });
        newGrid[lin(seeds[4 * s], seeds[4 * s + 1], seeds[4 * s + 2], dims)] = (uint16_t)seeds[4 * s + 3];
    }
    while (!front.empty()) {
        const V4 v = front.front();
        front.pop_front();
        auto expand = [&](int dx, int dy, int dz) {
            const uint64_t ci = lin(v.x + dx, v.y + dy, v.z + dz, dims);
            if (grid[ci] == v.w && newGrid[ci] == ORC_VOXEL_EMPTY) {
                front.push_back({ v.x + dx, v.y + dy, v.z + dz, v.w });
                newGrid[ci] = (uint16_t)v.w;
            }
        };
        if (v.x < dims[0] - 1) expand(+1, 0, 0);
        if (v.x > 0) expand(-1, 0, 0);
        if (v.y < dims[1] - 1) expand(0, +1, 0);
        if (v.y > 0) expand(0, -1, 0);
        if (v.z < dims[2] - 1) expand(0, 0, +1);
        if (v.z > 0) expand(0, 0, -1);
    }
    std::memcpy(grid, newGrid.data(), N * sizeof(uint16_t)); /* grid.swap(newGrid) :149 */
}

extern "C" void orc_detect_boundaries(uint16_t* grid, const uint32_t dims[3], int bs)
{
    /* detectBoundaries-comp.glsl:18-43.  In place is race-free by value: neighbours are read with bit 15 cleared,
     * the centre's raw word is only written by its own invocation. */
    const int64_t N = (int64_t)dims[0] * dims[1] * dims[2];
    std::vector<uint8_t> flag((size_t)N, 0);
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        const uint16_t own = grid[idx];
        if (own <= ORC_VOXEL_FREE) continue;
        int x, y, z;
        delin((uint64_t)idx, dims, x, y, z);
        const int x0 = clampi(x - bs, 0, (int)dims[0] - 1), x1 = clampi(x + bs, 0, (int)dims[0] - 1);
        const int y0 = clampi(y - bs, 0, (int)dims[1] - 1), y1 = clampi(y + bs, 0, (int)dims[1] - 1);
        const int z0 = clampi(z - bs, 0, (int)dims[2] - 1), z1 = clampi(z + bs, 0, (int)dims[2] - 1);
        bool boundary = false;
        for (int a = x0; a <= x1 && !boundary; ++a)
            for (int b = y0; b <= y1 && !boundary; ++b)
                for (int cz = z0; cz <= z1 && !boundary; ++cz) {
                    const uint16_t v = grid[lin(a, b, cz, dims)] & (uint16_t)~(1u << ORC_MASK_POSITION);
                    boundary = boundary || (v > ORC_VOXEL_FREE && v != own);
                }
        flag[idx] = boundary;
    }
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx)
        if (flag[idx]) grid[idx] |= (uint16_t)(1u << ORC_MASK_POSITION);
}

extern "C" void orc_fill_noise(orc_rng* rng, float* noise, uint32_t n)
{
    /* RegularGrid.cpp:238-244, serialised (finding 8: the reference's omp loop over a shared mt19937 is not reproducible) */
    for (uint32_t i = 0; i < n; ++i) noise[i] = orc_rng_uniform_range(rng, .0f, 1.0f);
}

extern "C" void orc_erode_mask(int type, uint32_t size, float* mask, float* activations_out, uint32_t* k_out)
{
    /* RegularGrid.cpp:84-122 */
    if (!(size % 2)) ++size;
    const uint32_t k = size, maskSize = k * k * k, cc = (uint32_t)std::floor(k / 2.0f);
    float activations = 0;
    std::fill(mask, mask + maskSize, 0.0f);
    if (type == ORC_SQUARE) {
        std::fill(mask, mask + maskSize, 1.0f);
        activations = (float)maskSize;
    } else if (type == ORC_CROSS) {
        for (uint32_t x = 0; x < k; ++x) mask[x * k * k + cc * k + cc] = 1.0f;
        for (uint32_t y = 0; y < k; ++y) mask[cc * k * k + y * k + cc] = 1.0f;
        for (uint32_t z = 0; z < k; ++z) mask[cc * k * k + cc * k + z] = 1.0f;
        activations = 1.0f / 3.0f * maskSize;
    } else if (type == ORC_ELLIPSE) {
        for (uint32_t x = 0; x < k; ++x)
            for (uint32_t y = 0; y < k; ++y)
                for (uint32_t z = 0; z < k; ++z) {
                    const float dx = (float)x - (float)cc, dy = (float)y - (float)cc, dz = (float)z - (float)cc;
                    if (std::sqrt(dx * dx + dy * dy + dz * dz) < (float)cc + FLT_EPSILON) { /* glm::epsilon<float>() */
                        mask[x * k * k + y * k + z] = 1.0f;
                        ++activations;
                    }
                }
    }
    activations /= maskSize;
    *activations_out = activations;
    *k_out = k;
}

extern "C" void orc_remove_isolated_regions_grid(uint16_t* grid, const uint32_t dims[3])
{
    /* removeIsolatedRegionsGrid-comp.glsl:16-39.  The reference runs in place with racy neighbour reads; the
     * deterministic rule here reads every neighbour from the pre-pass snapshot. */
    const int64_t N = (int64_t)dims[0] * dims[1] * dims[2];
    std::vector<uint16_t> src(grid, grid + N);
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < N; ++idx) {
        int x, y, z;
        delin((uint64_t)idx, dims, x, y, z);
        const int x0 = clampi(x - 1, 0, (int)dims[0] - 1), x1 = clampi(x + 1, 0, (int)dims[0] - 1);
        const int y0 = clampi(y - 1, 0, (int)dims[1] - 1), y1 = clampi(y + 1, 0, (int)dims[1] - 1);
        const int z0 = clampi(z - 1, 0, (int)dims[2] - 1), z1 = clampi(z + 1, 0, (int)dims[2] - 1);
        int count = -1;
        for (int a = x0; a <= x1; ++a)
            for (int b = y0; b <= y1; ++b)
                for (int c = z0; c <= z1; ++c) count += (int)(src[lin(a, b, c, dims)] == src[idx]);
        if (count < 6) grid[idx] = ORC_VOXEL_EMPTY;
    }
}

extern "C" void orc_erode(uint16_t* grid, const uint32_t dims[3], int type, uint32_t size, uint32_t iters, float prob, float thr,
                          const float* noise, uint32_t nnoise, int boundary_mode)
{
    /* RegularGrid.cpp:82-159 + erodeGrid-comp.glsl:26-59 + copyGrid-comp.glsl */
    const int64_t N = (int64_t)dims[0] * dims[1] * dims[2];
    uint32_t k;
    float activations;
    std::vector<float> mask((size_t)(size + 1) * (size + 1) * (size + 1));
    orc_erode_mask(type, size, mask.data(), &activations, &k);
    const int k2 = (int)std::floor(k / 2.0f); /* maskSize2 :143 */
    std::vector<uint16_t> dest((size_t)N);
    for (uint32_t it = 0; it < iters; ++it) {
        orc_detect_boundaries(grid, dims, 1); /* :135 */
#pragma omp parallel for schedule(static)
        for (int64_t idx = 0; idx < N; ++idx) {
            const uint16_t own = grid[idx];
            const bool isBoundary = boundary_mode == 0 ? (own & (uint16_t)~(1u << ORC_MASK_POSITION)) != 0 /* :31 as written */
                                                       : (own >> ORC_MASK_POSITION) != 0;
            dest[idx] = own;
            if (own > ORC_VOXEL_FREE && isBoundary && noise[(uint64_t)idx % nnoise] < prob) {
                int x, y, z;
                delin((uint64_t)idx, dims, x, y, z);
                const int mnx = x - k2, mny = y - k2, mnz = z - k2;
                const int x0 = clampi(mnx, 0, (int)dims[0] - 1), x1 = clampi(x + k2, 0, (int)dims[0] - 1);
                const int y0 = clampi(mny, 0, (int)dims[1] - 1), y1 = clampi(y + k2, 0, (int)dims[1] - 1);
                const int z0 = clampi(mnz, 0, (int)dims[2] - 1), z1 = clampi(z + k2, 0, (int)dims[2] - 1);
                uint32_t count = 0, globalCount = 0;
                for (int a = x0; a <= x1; ++a)
                    for (int b = y0; b <= y1; ++b)
                        for (int c = z0; c <= z1; ++c) {
                            const float w = mask[(size_t)(a - mnx) * k * k + (size_t)(b - mny) * k + (size_t)(c - mnz)];
                            count += (uint32_t)((float)(uint32_t)(grid[lin(a, b, c, dims)] == own) * w); /* :49 */
                            ++globalCount;
                        }
                const float activation = (float)count / (float)globalCount;
                if (activation < activations * thr) dest[idx] = ORC_VOXEL_EMPTY; /* :55-57 */
            }
        }
        std::memcpy(grid, dest.data(), (size_t)N * sizeof(uint16_t)); /* copyGrid :149-152 */
    }
    orc_remove_isolated_regions_grid(grid, dims); /* :155 */
}

extern "C" void orc_undo_mask(uint16_t* grid, uint64_t n, uint32_t position, int rightmost)
{
    /* voxelMask.glsl:9-17, undoMask-comp.glsl:19-37 */
    const uint16_t m = rightmost ? (uint16_t)((1u << position) - 1) : (uint16_t)~(1u << position);
    for (uint64_t i = 0; i < n; ++i) grid[i] &= m;
}
extern "C" void orc_reset_filling(uint16_t* grid, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) grid[i] = std::min<uint16_t>(grid[i], ORC_VOXEL_FREE + 1); /* RegularGrid.cpp:417 */
}
extern "C" void orc_homogenize(uint16_t* grid, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i)
        if (grid[i] != ORC_VOXEL_EMPTY) grid[i] = ORC_VOXEL_FREE; /* RegularGrid.cpp:533-541 */
}

/* ===================================================================================== H1 */

extern "C" uint64_t orc_count_values(const uint16_t* grid, uint64_t n, uint32_t* counts)
{
    /* RegularGrid.cpp:601-625 (countValues), :280-287 (numOccupiedVoxels) */
    std::memset(counts, 0, 32768 * sizeof(uint32_t));
    uint64_t occupied = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (grid[i] > ORC_VOXEL_FREE) {
            ++counts[grid[i] & 0x7FFF]; /* unmask() :1026-1029 */
            ++occupied;
        }
    return occupied;
}

/* ===================================================================================== X1 */

extern "C" uint64_t orc_encode_rle(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    /* RegularGrid.cpp:672-714: uvec3 dims, then {uint16 value, uint32 repetitions} packed, over the flat x-major array */
    const uint64_t size = (uint64_t)dims[0] * dims[1] * dims[2];
    uint64_t pos = 12;
    if (out && cap >= 12) std::memcpy(out, dims, 12);
    uint64_t idx = 0;
    while (idx < size) {
        const uint16_t value = grid[idx];
        uint32_t rep = 0;
        while (idx < size && grid[idx] == value) { ++rep; ++idx; }
        if (out && pos + 6 <= cap) {
            std::memcpy(out + pos, &value, 2);
            std::memcpy(out + pos + 2, &rep, 4);
        }
        pos += 6;
    }
    return pos;
}

extern "C" int orc_decode_rle(const uint8_t* data, uint64_t len, uint32_t dims[3], uint16_t* grid, uint64_t cap)
{
    /* inverse of the above == docs/decompress/decompress_grid.py:16-33 (without its dataset-specific flip/pad) */
    if (len < 12) return ORC_ERR_IO;
    std::memcpy(dims, data, 12);
    const uint64_t N = (uint64_t)dims[0] * dims[1] * dims[2];
    if (N > cap) return ORC_ERR_CAPACITY;
    uint64_t off = 0;
    for (uint64_t p = 12; p + 6 <= len; p += 6) {
        uint16_t v;
        uint32_t r;
        std::memcpy(&v, data + p, 2);
        std::memcpy(&r, data + p + 2, 4);
        if (off + r > N) return ORC_ERR_IO;
        std::fill(grid + off, grid + off + r, v);
        off += r;
    }
    return off == N ? ORC_OK : ORC_ERR_IO;
}

extern "C" uint64_t orc_encode_bing_squared(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    /* RegularGrid.cpp:638-666 */
    const uint32_t M = std::max(dims[0], std::max(dims[1], dims[2]));
    const uint64_t need = 12 + (uint64_t)M * M * M * 2;
    if (!out || cap < need) return need;
    const uint32_t end[3] = { M, M, M };
    std::memcpy(out, end, 12);
    const int start[3] = { (int)((M - dims[0]) / 2), (int)((M - dims[1]) / 2), (int)((M - dims[2]) / 2) };
    uint64_t pos = 12;
    for (int x = 0; x < (int)M; ++x)
        for (int y = 0; y < (int)M; ++y)
            for (int z = 0; z < (int)M; ++z) {
                const int cx = x - start[0], cy = y - start[1], cz = z - start[2];
                uint16_t v = ORC_VOXEL_EMPTY;
                if (inside(dims, cx, cy, cz)) v = grid[lin(cx, cy, cz, dims)];
                std::memcpy(out + pos, &v, 2);
                pos += 2;
            }
    return pos;
}

/* ---- .vox: RegularGrid::exportVox (RegularGrid.cpp:740-798) over VoxWriter (Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp).
 * A call-by-call restatement: an AddVoxel per cell in the reference's order, cubes in a map in order of first appearance,
 * SaveToFile's chunk sequence.  Pinned byte-for-byte against the reference's own VoxWriter.cpp (tests/test_oracle_vs_ref.py). */
namespace {
struct VoxCubeO {
    int tx, ty, tz;
    std::vector<uint8_t> voxels;
};
struct VoxWriterO {
    static constexpr size_t L = 126;             /* VoxWriter.h:451 */
    size_t minX = 10000000, minY = 10000000, minZ = 10000000; /* VoxWriter.h:428-430 */
    double lo[3] = { 1e7, 1e7, 1e7 }, hi[3] = { -1e7, -1e7, -1e7 }; /* maxVolume, VoxWriter.h:432 */
    std::map<std::array<size_t, 3>, size_t> ids; /* cubesId */
    std::vector<VoxCubeO> cubes;
    void add(size_t vx, size_t vy, size_t vz, uint8_t colour)
    {
        /* AddVoxel, VoxWriter.cpp:449-461 (positions are never repeated by exportVox, so the voxelId lookup is dropped) */
        const size_t ox = vx / L, oy = vy / L, oz = vz / L;
        minX = std::min(minX, ox);
        minY = std::min(minX, oy); /* sic: against minCubeX */
        minZ = std::min(minX, oz); /* sic */
        auto it = ids.find({ ox, oy, oz });
        if (it == ids.end()) {
            it = ids.emplace(std::array<size_t, 3>{ ox, oy, oz }, cubes.size()).first;
            cubes.push_back({ (int)ox, (int)oy, (int)oz, {} });
        }
        const double p[3] = { (double)vx, (double)vy, (double)vz };
        for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], p[a]), hi[a] = std::max(hi[a], p[a]);
        auto& v = cubes[it->second].voxels;
        v.push_back((uint8_t)(vx % L)), v.push_back((uint8_t)(vy % L)), v.push_back((uint8_t)(vz % L)), v.push_back(colour);
    }
};
struct ByteSink {
    uint8_t* out;
    uint64_t cap, pos;
    void put(const void* p, uint64_t n)
    {
        if (out && pos + n <= cap) std::memcpy(out + pos, p, n);
        pos += n;
    }
    void i32(int32_t v) { put(&v, 4); }
    void tag(const char* t) { put(t, 4); }
    void dict(const std::vector<std::pair<std::string, std::string>>& kv)
    {
        i32((int32_t)kv.size());
        for (auto& e : kv) i32((int32_t)e.first.size()), put(e.first.data(), e.first.size()), i32((int32_t)e.second.size()), put(e.second.data(), e.second.size());
    }
};
uint64_t dict_size(const std::vector<std::pair<std::string, std::string>>& kv)
{
    uint64_t s = 4;
    for (auto& e : kv) s += 8 + e.first.size() + e.second.size();
    return s;
}
int to_int_x86(double v) { return (v >= -2147483648.0 && v < 2147483648.0) ? (int)v : INT_MIN; }
}  // namespace

extern "C" uint64_t orc_encode_vox(const uint16_t* grid, const uint32_t dims[3], int squared, uint8_t* out, uint64_t cap)
{
    VoxWriterO vw;
    if (squared) { /* RegularGrid.cpp:747-768 */
        const int M = (int)std::max(dims[0], std::max(dims[1], dims[2]));
        const int start[3] = { (int)((M - dims[0]) / 2), (int)((M - dims[1]) / 2), (int)((M - dims[2]) / 2) };
        for (int x = 0; x < M; ++x)
            for (int y = 0; y < M; ++y)
                for (int z = 0; z < M; ++z) {
                    const int cx = x - start[0], cy = y - start[1], cz = z - start[2];
                    vw.add(x, z, y, inside(dims, cx, cy, cz) ? (uint8_t)grid[lin(cx, cy, cz, dims)] : (uint8_t)ORC_VOXEL_EMPTY);
                }
    } else { /* :770-785 */
        for (int x = 0; x < (int)dims[0]; ++x)
            for (int y = 0; y < (int)dims[1]; ++y)
                for (int z = 0; z < (int)dims[2]; ++z) {
                    const uint16_t v = grid[lin(x, y, z, dims)];
                    if (v > ORC_VOXEL_FREE) vw.add(x, z, y, (uint8_t)(v - ORC_VOXEL_FREE));
                }
    }
    /* SaveToFile, VoxWriter.cpp:462-540 */
    ByteSink w{ out, cap, 0 };
    w.tag("VOX "), w.i32(150), w.tag("MAIN"), w.i32(0);
    const uint64_t patch = w.pos;
    w.i32(0);
    const uint64_t header = w.pos;
    typedef std::vector<std::pair<std::string, std::string>> Dict;
    std::vector<Dict> frames;
    for (auto& c : vw.cubes) {
        w.tag("SIZE"), w.i32(12), w.i32(0), w.i32(126), w.i32(126), w.i32(126);
        const int32_t nvox = (int32_t)c.voxels.size() / 4;
        w.tag("XYZI"), w.i32(4 * (1 + nvox)), w.i32(0), w.i32(nvox), w.put(c.voxels.data(), c.voxels.size());
        /* :489-491, with minCube* being size_t: the subtraction is unsigned, then float, then double */
        c.tx = to_int_x86(std::floor(((size_t)c.tx - vw.minX + 0.5f) * VoxWriterO::L - vw.lo[0] - (vw.hi[0] - vw.lo[0]) * 0.5));
        c.ty = to_int_x86(std::floor(((size_t)c.ty - vw.minY + 0.5f) * VoxWriterO::L - vw.lo[1] - (vw.hi[1] - vw.lo[1]) * 0.5));
        c.tz = to_int_x86(std::floor(((size_t)c.tz - vw.minZ + 0.5f) * VoxWriterO::L));
        frames.push_back({ { "_t", std::to_string(c.tx) + " " + std::to_string(c.ty) + " " + std::to_string(c.tz) } });
    }
    const int32_t n = (int32_t)vw.cubes.size();
    const Dict none, keyframe = { { "_f", "0" } };
    w.tag("nTRN"), w.i32((int32_t)(20 + dict_size(none) + dict_size(none))), w.i32(0);
    w.i32(0), w.dict(none), w.i32(1), w.i32(-1), w.i32(-1), w.i32(1), w.dict(none);
    w.tag("nGRP"), w.i32((int32_t)(4 * (2 + n) + dict_size(none))), w.i32(0);
    w.i32(1), w.dict(none), w.i32(n);
    for (int32_t i = 0; i < n; ++i) w.i32(2 + 2 * i);
    for (int32_t i = 0; i < n; ++i) {
        w.tag("nTRN"), w.i32((int32_t)(20 + dict_size(none) + dict_size(frames[i]))), w.i32(0);
        w.i32(2 + 2 * i), w.dict(none), w.i32(3 + 2 * i), w.i32(-1), w.i32(0), w.i32(1), w.dict(frames[i]);
        w.tag("nSHP"), w.i32((int32_t)(8 + dict_size(none) + 4 + dict_size(keyframe))), w.i32(0);
        w.i32(3 + 2 * i), w.dict(none), w.i32(1), w.i32(i), w.dict(keyframe);
    }
    const uint32_t children = (uint32_t)(w.pos - header);
    if (out && w.pos <= cap) std::memcpy(out + patch, &children, 4);
    return w.pos;
}

/* ---- .qstack: RegularGrid::exportQuadStack (RegularGrid.cpp:716-725) = QuadStack<uint16_t>::loadCube, compress_y, compress_x,
 * saveCheckpoint (SRC/DataStructures/QuadStack.h:99-140,188-226,247-310,343-433) over GStack<uint16_t> (SRC/DataStructures/GStack.h).
 * A call-by-call restatement that keeps the structure's observable quirks: interval counts are read through a uint8_t
 * (GStack.h:50), columns are "identical" when their value sequences agree whatever the run lengths (GStack.h:202-205,104-111),
 * leaf height fields are cumulative (QuadStack.h:289-303: the matrix is not reset between layers), and mergeStacks erases the
 * merged interval from the children while its depth-0 loop keeps counting (GStack.h:278-319).  loadCube's function-static
 * materialMatrix only ever grows (QuadStack.h:125-126, RegularGrid.h:370-391); the restatement is the first export of a process.
 * Pinned byte-for-byte against the reference's own headers compiled in place (tests/test_oracle_vs_ref.py). */
namespace {
struct QsInterval {
    uint16_t value;
    std::vector<std::vector<uint16_t>> length;
};
struct QsStack {
    QsStack* children[2][2] = { { nullptr, nullptr }, { nullptr, nullptr } };
    std::vector<QsInterval> intervals;
    uint32_t maxp[2] = { 0, 0 }, minp[2] = { UINT_MAX, UINT_MAX }; /* GStack.h:57 */
    ~QsStack()
    {
        for (auto& r : children)
            for (auto*& c : r) delete c;
    }
    uint8_t numIntervals() const { return (uint8_t)intervals.size(); } /* GStack.h:50 */
    uint8_t numChildren() const
    {
        uint8_t n = 0;
        for (auto& r : children)
            for (auto* c : r) n += c != nullptr;
        return n;
    }
    bool isWildcard(uint8_t i) const { return intervals[i].value == 0xFFFF; } /* GStack.h:171-175 */
    void bound(uint32_t x, uint32_t y) /* GStack.h:240-244 */
    {
        maxp[0] = std::max(maxp[0], x), maxp[1] = std::max(maxp[1], y);
        minp[0] = std::min(minp[0], x), minp[1] = std::min(minp[1], y);
    }
    bool sameColours(const QsStack& o) const /* operator== + areColorEqual, GStack.h:202-205,104-111 */
    {
        if (o.intervals.size() != intervals.size()) return false;
        for (size_t i = 0; i < o.intervals.size(); ++i)
            if (o.intervals[i].value != intervals[i].value) return false;
        return true;
    }
};

struct QsTree {
    uint16_t W, H, D;
    std::vector<std::vector<QsStack>> cols; /* _gStacks after compress_y */
    QsStack* root = nullptr;
    ~QsTree() { delete root; }

    QsStack* create(uint16_t x0, uint16_t x1, uint16_t y0, uint16_t y1, bool leaf = false, QsStack* node = nullptr) /* QuadStack.h:281-310 */
    {
        QsStack* g = node ? node : new QsStack;
        g->bound(x0, y0), g->bound(x1, y1);
        if (leaf) {
            const size_t layers = cols[x0][y0].intervals.size();
            std::vector<std::vector<uint16_t>> m(x1 - x0, std::vector<uint16_t>(y1 - y0, 0));
            for (size_t l = 0; l < layers; ++l) {
                for (uint16_t x = x0; x < x1; ++x)
                    for (uint16_t y = y0; y < y1; ++y) m[x - x0][y - y0] += cols[x][y].intervals[l].length[0][0];
                g->intervals.push_back({ cols[x0][y0].intervals[l].value, m });
            }
        }
        return g;
    }
    void split(uint16_t x0, uint16_t x1, uint16_t y0, uint16_t y1, QsStack* node) /* recursiveSplitQuadtree, QuadStack.h:376-433 */
    {
        if ((x1 - x0) <= 1 && (y1 - y0) <= 1) {
            create(x0, x1, y0, y1, true, node);
            return;
        }
        bool same = true;
        for (uint16_t x = x0; x < x1 && same; ++x)
            for (uint16_t y = y0; y < y1 && same; ++y) same &= cols[x0][y0].sameColours(cols[x][y]);
        if (same) {
            create(x0, x1, y0, y1, true, node);
            return;
        }
        const size_t sx = x1 - x0, ex = (sx + 1) / 2, sy = y1 - y0, ey = (sy + 1) / 2;
        node->children[0][0] = create(x0, x0 + ex, y0, y0 + ey);
        split(x0, x0 + ex, y0, y0 + ey, node->children[0][0]);
        if (sy > 1) {
            node->children[0][1] = create(x0, x0 + ex, y0 + ey, y1);
            split(x0, x0 + ex, y0 + ey, y1, node->children[0][1]);
        }
        if (sx > 1) {
            node->children[1][0] = create(x0 + ex, x1, y0, y0 + ey);
            split(x0 + ex, x1, y0, y0 + ey, node->children[1][0]);
        }
        if (sx > 1 && sy > 1) {
            node->children[1][1] = create(x0 + ex, x1, y0 + ey, y1);
            split(x0 + ex, x1, y0 + ey, y1, node->children[1][1]);
        }
    }
    static bool merge(QsStack* rootn, std::vector<QsStack*>& st, std::vector<uint8_t>& it, uint8_t depth) /* GStack.h:278-319 */
    {
        if (depth == it.size()) {
            bool same = true, wildcard = st[0]->isWildcard(it[0]);
            for (size_t k = 1; k < st.size() && same && !wildcard; ++k) {
                same &= st[k]->intervals[it[k]].value == st[k - 1]->intervals[it[k - 1]].value;
                wildcard |= st[k]->isWildcard(it[k]);
            }
            if (same && !wildcard) {
                QsInterval iv; /* mergeInterval, GStack.h:185-196 */
                iv.value = st[0]->intervals[it[0]].value;
                iv.length.assign(rootn->maxp[0] - rootn->minp[0], std::vector<uint16_t>(rootn->maxp[1] - rootn->minp[1], 0));
                for (size_t k = 0; k < st.size(); ++k)
                    for (unsigned x = st[k]->minp[0]; x < st[k]->maxp[0]; ++x)
                        for (unsigned y = st[k]->minp[1]; y < st[k]->maxp[1]; ++y)
                            iv.length[x - rootn->minp[0]][y - rootn->minp[1]] = st[k]->intervals[it[k]].length[x - st[k]->minp[0]][y - st[k]->minp[1]];
                rootn->intervals.push_back(iv);
                for (size_t k = 0; k < st.size(); ++k) st[k]->intervals.erase(st[k]->intervals.begin() + it[k]);
            }
            return true;
        }
        const size_t start = depth > 0 ? it[depth - 1] : 0;
        for (size_t i = start; i < st[depth]->numIntervals(); ++i) {
            it[depth] = (uint8_t)i;
            if (merge(rootn, st, it, depth + 1) && depth > 0) return true;
        }
        return false;
    }
    void compress(QsStack* node) /* compressQuadStack(root, depth, true), QuadStack.h:260-279 */
    {
        if (!node || !node->numChildren()) return;
        for (auto& r : node->children)
            for (auto* c : r) compress(c);
        std::vector<QsStack*> kids;
        for (auto& r : node->children)
            for (auto* c : r)
                if (c) kids.push_back(c);
        std::vector<uint8_t> it(kids.size(), 0);
        merge(node, kids, it, 0);
    }
    void leaves(QsStack* node, std::vector<QsStack*>& out) /* QuadStack::getLeaves, QuadStack.h:33-44 */
    {
        if (!node) return;
        if (node->numIntervals()) out.push_back(node);
        for (auto& r : node->children)
            for (auto* c : r) leaves(c, out);
    }
};
}  // namespace

extern "C" uint64_t orc_encode_qstack(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    QsTree t;
    t.W = (uint16_t)dims[0], t.H = (uint16_t)dims[1], t.D = (uint16_t)dims[2];
    if (!t.W || !t.H || !t.D) return 0; /* loadCube fails; compress_x would dereference an empty structure */
    /* loadCube + compress_y: one unit interval per cell (GStack.h:177-183), then addColor per cell (GStack.h:89-101,127-138) */
    t.cols.assign(t.W, std::vector<QsStack>(t.H));
    for (uint32_t x = 0; x < t.W; ++x)
        for (uint32_t y = 0; y < t.H; ++y) {
            QsStack& c = t.cols[x][y];
            for (uint32_t z = 0; z < t.D; ++z) {
                const uint16_t v = grid[lin(x, y, z, dims)];
                if (c.intervals.empty() || c.intervals.back().value != v)
                    c.intervals.push_back({ v, { { 1 } } });
                else
                    ++c.intervals.back().length[0][0];
            }
            c.bound(x, y);
        }
    /* compress_x: buildQuadStack + compressQuadStack (QuadStack.h:91-96,228-234) */
    t.root = t.create(0, t.W, 0, t.H);
    t.split(0, t.W, 0, t.H, t.root);
    t.compress(t.root);
    /* saveCheckpoint, QuadStack.h:188-226 (size_t is 8 bytes on the reference's x64 target) */
    ByteSink w{ out, cap, 0 };
    std::vector<QsStack*> nodes;
    t.leaves(t.root, nodes);
    const uint64_t tSize = sizeof(uint16_t), numNodes = nodes.size();
    w.put(&tSize, 8), w.put(&t.W, 2), w.put(&t.H, 2), w.put(&t.D, 2), w.put(&numNodes, 8);
    for (QsStack* n : nodes) {
        const uint64_t ni = n->numIntervals();
        w.put(&ni, 8), w.put(n->maxp, 8), w.put(n->minp, 8);
        for (uint64_t i = 0; i < ni; ++i) {
            const QsInterval& iv = n->intervals[i];
            const uint8_t lw = (uint8_t)iv.length.size(), lh = (uint8_t)iv.length[0].size();
            w.put(&lw, 1), w.put(&lh, 1), w.put(&iv.value, 2);
            for (auto& row : iv.length)
                for (uint16_t l : row) w.put(&l, 2);
        }
    }
    return w.pos;
}

/* ---- f2: per-fragment marching cubes — RegularGrid::toTriangleMesh's mesh side (RegularGrid.cpp:473-486) = MarchingCubes::setGrid +
 * triangulateFieldGPU (SRC/Graphics/Core/MarchingCubes.cpp:523-540, 364-432) over the shaders marchingCubes-comp.glsl:96-168,
 * computeMortonCodes-comp.glsl, findSameVertices_01/02-comp.glsl, buildMarchingCubesFaces-comp.glsl, markBoundaryTriangles-comp.glsl,
 * resetLaplacianBuffer / laplacianSmoothing / finishLaplacianSmoothing-comp.glsl, for _marchingCubesSubdivisions == 1 (the default).
 * The reference hands out vertex slots and fused-vertex numbers with atomicAdd (marchingCubes-comp.glsl:139, findSameVertices_01:28),
 * so the ORDER of its vertices and faces is a race; positions, connectivity and flags are not.  The deterministic rule fixed here
 * ("what the shaders give when their threads run in index order"): triangles in (cell index, row position) order; vertices sorted by
 * (30-bit Morton code as computed by the shader, then x, y, z) — the reference sorts by the Morton code alone, stably — and numbered in
 * that order; a fused vertex takes position and boundary flag from its first copy.  float32 throughout, mix(a, b, w) = a*(1-w) + b*w,
 * model matrix applied as x*scale + (min + (-scale)).  PARITY UNPINNED BY CONSTRUCTION (GLSL only, racy order). */
namespace {
const uint64_t kMcRows[256] = {
#include "../voxelfragmentml_b200/csrc/mc_tritable.inc"
};
const int kMcCorner[8][3] = { { 0, 0, 0 }, { 0, 0, 1 }, { -1, 0, 1 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, 1, 1 }, { -1, 1, 1 }, { -1, 1, 0 } }; /* marchingCubes-comp.glsl:28-38 */
const int kMcEdge[12][2] = { { 0, 1 }, { 1, 2 }, { 2, 3 }, { 3, 0 }, { 4, 5 }, { 5, 6 }, { 6, 7 }, { 7, 4 }, { 0, 4 }, { 1, 5 }, { 2, 6 }, { 3, 7 } }; /* :40-54 */
inline uint32_t mc_expand_bits(uint32_t v) /* computeMortonCodes-comp.glsl expandBits */
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
struct McVert {
    float p[3];
    float w;
    uint32_t morton;
};
}  // namespace

/* setGrid + the march dispatch + computeMortonCodes: the triangle soup of one fragment, three vertices per triangle, cells in index order */
static void mc_soup(const uint16_t* grid, const uint32_t dims[3], uint32_t target, std::vector<McVert>& soup)
{
    /* setGrid, MarchingCubes.cpp:523-540: dims + 2, a ring of VOXEL_FREE */
    const int PX = (int)dims[0] + 2, PY = (int)dims[1] + 2, PZ = (int)dims[2] + 2;
    std::vector<uint16_t> P((size_t)PX * PY * PZ, (uint16_t)ORC_VOXEL_FREE);
    for (int x = 1; x < PX - 1; ++x)
        for (int y = 1; y < PY - 1; ++y)
            for (int z = 1; z < PZ - 1; ++z) P[((size_t)x * PY + y) * PZ + z] = grid[lin(x - 1, y - 1, z - 1, dims)];
    /* march, marchingCubes-comp.glsl:96-156, cells in index order */
    const float pd[3] = { (float)PX, (float)PY, (float)PZ };
    for (int x = 0; x < PX; ++x)
        for (int y = 0; y < PY; ++y)
            for (int z = 0; z < PZ; ++z) {
                if (x == 0 || y == PY - 1 || z == PZ - 1) continue; /* :102-103 */
                float values[8];
                int configuration = 0;
                for (int i = 0; i < 8; ++i) {
                    const uint16_t v = P[((size_t)(x + kMcCorner[i][0]) * PY + (y + kMcCorner[i][1])) * PZ + (z + kMcCorner[i][2])];
                    values[i] = (uint16_t)(v & 0x7FFFu) == (uint16_t)target ? 1.0f : 0.0f;
                    if (values[i] < 0.5f) configuration |= 1 << i;
                }
                const uint64_t row = kMcRows[configuration];
                if ((row & 0xF) == 0xF) continue;
                const float wflag = (P[((size_t)x * PY + y) * PZ + z] >> 15) != 0 ? 1.0f : 0.0f; /* :147 */
                auto edge_vertex = [&](int e, float out[3]) { /* findVertex :67-87 with isolevel 0.5 and values in {0, 1} */
                    const int a = kMcEdge[e][0], b = kMcEdge[e][1];
                    const float mu = (0.5f - values[a]) / (values[b] - values[a]);
                    const int c[3] = { x, y, z };
                    for (int q = 0; q < 3; ++q) {
                        const float p1 = (float)(c[q] + kMcCorner[a][q]), p2 = (float)(c[q] + kMcCorner[b][q]);
                        out[q] = p1 + mu * (p2 - p1);
                    }
                };
                for (int t = 0; t < 5; ++t) {
                    const int e0 = (int)((row >> (12 * t)) & 0xF);
                    if (e0 == 0xF) break; /* rows are packed front to back */
                    const int e1 = (int)((row >> (12 * t + 4)) & 0xF), e2 = (int)((row >> (12 * t + 8)) & 0xF);
                    const int order[3] = { e0, e2, e1 }; /* :141-143 */
                    for (int k = 0; k < 3; ++k) {
                        McVert v;
                        edge_vertex(order[k], v.p);
                        v.w = wflag;
                        /* computeMortonCodes-comp.glsl: (p - 0) / (numDivs - 0) * 1024, truncated */
                        const uint32_t xx = mc_expand_bits((uint32_t)((v.p[0] / pd[0]) * 1024.0f)), yy = mc_expand_bits((uint32_t)((v.p[1] / pd[1]) * 1024.0f)),
                                       zz = mc_expand_bits((uint32_t)((v.p[2] / pd[2]) * 1024.0f));
                        v.morton = xx * 4 + yy * 2 + zz;
                        soup.push_back(v);
                    }
                }
            }
}

/* the soup alone (tests: the reference's marchingCubes-comp.glsl and computeMortonCodes-comp.glsl compiled in place produce the same vertices
 * and codes, up to the order the shader's atomic counter hands out): verts[4 * i] = x, y, z (padded-grid units), boundary flag */
extern "C" uint32_t orc_mc_soup(const uint16_t* grid, const uint32_t dims[3], uint32_t target, float* verts, uint32_t* morton, uint32_t cap)
{
    std::vector<McVert> soup;
    mc_soup(grid, dims, target, soup);
    for (uint32_t i = 0; i < soup.size() && i < cap; ++i) {
        verts[4 * i] = soup[i].p[0], verts[4 * i + 1] = soup[i].p[1], verts[4 * i + 2] = soup[i].p[2], verts[4 * i + 3] = soup[i].w;
        morton[i] = soup[i].morton;
    }
    return (uint32_t)soup.size();
}

extern "C" int orc_marching_cubes(const uint16_t* grid, const uint32_t dims[3], uint32_t target, const float amin[3], const float amax[3], uint32_t nb_iters,
                                  float nb_weight, uint32_t b_iters, float b_weight, float* verts, uint32_t cap_v, uint32_t* faces, uint32_t cap_f,
                                  uint32_t counts[2])
{
    std::vector<McVert> soup;
    mc_soup(grid, dims, target, soup);
    const uint32_t nsoup = (uint32_t)soup.size(), nf = nsoup / 3;
    /* sortMortonCodes + findSameVertices_01/02 under the deterministic order */
    std::vector<uint32_t> idx(nsoup);
    for (uint32_t i = 0; i < nsoup; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
        const McVert &A = soup[a], &B = soup[b];
        if (A.morton != B.morton) return A.morton < B.morton;
        for (int q = 0; q < 3; ++q)
            if (A.p[q] != B.p[q]) return A.p[q] < B.p[q];
        return false;
    });
    /* model matrix, RegularGrid.cpp:478-480 */
    float scale[3], shift[3];
    for (int q = 0; q < 3; ++q) scale[q] = (amax[q] - amin[q]) / (float)dims[q], shift[q] = amin[q] + (-scale[q]);
    std::vector<float> V; /* xyzw */
    std::vector<uint32_t> fused(nsoup);
    for (uint32_t k = 0; k < nsoup; ++k) {
        const McVert& v = soup[idx[k]];
        bool fresh = k == 0;
        if (!fresh) {
            const McVert& u = soup[idx[k - 1]];
            fresh = !(u.p[0] == v.p[0] && u.p[1] == v.p[1] && u.p[2] == v.p[2]); /* distance > 1e-8 on half-integer coordinates */
        }
        if (fresh) {
            for (int q = 0; q < 3; ++q) V.push_back(v.p[q] * scale[q] + shift[q]);
            V.push_back(v.w);
        }
        fused[idx[k]] = (uint32_t)(V.size() / 4 - 1);
    }
    const uint32_t nv = (uint32_t)(V.size() / 4);
    counts[0] = nv, counts[1] = nf;
    if (!verts || !faces || cap_v < nv || cap_f < nf) return ORC_OK;
    std::vector<uint32_t> F((size_t)nf * 4);
    for (uint32_t f = 0; f < nf; ++f) {
        for (int k = 0; k < 3; ++k) F[4 * (size_t)f + k] = fused[3 * f + k]; /* buildMarchingCubesFaces-comp.glsl */
        const float m = std::max(V[4 * (size_t)F[4 * (size_t)f] + 3], std::max(V[4 * (size_t)F[4 * (size_t)f + 1] + 3], V[4 * (size_t)F[4 * (size_t)f + 2] + 3]));
        F[4 * (size_t)f + 3] = (uint32_t)m; /* markBoundaryTriangles-comp.glsl */
    }
    /* smoothSurface, MarchingCubes.cpp:498-521: non-boundary pass, then boundary pass */
    std::vector<int32_t> L((size_t)nv * 4);
    auto smooth = [&](uint32_t iters, float weight, bool boundary) {
        const float tgt = boundary ? 1.0f : 0.0f;
        for (uint32_t it = 0; it < iters; ++it) {
            std::fill(L.begin(), L.end(), 0);
            for (uint32_t f = 0; f < nf; ++f) { /* laplacianSmoothing-comp.glsl */
                const uint32_t* fc = &F[4 * (size_t)f];
                bool valid = true;
                if (!boundary)
                    for (int i = 0; i < 3 && valid; ++i) valid = std::fabs(V[4 * (size_t)fc[i] + 3] - tgt) < 0.00000001f;
                if (!valid) continue;
                for (int i = 0; i < 3; ++i) {
                    const float* p = &V[4 * (size_t)fc[i]];
                    const int32_t q[4] = { (int32_t)(p[0] * 10000.0f), (int32_t)(p[1] * 10000.0f), (int32_t)(p[2] * 10000.0f), 1 };
                    for (int n = 1; n <= 2; ++n)
                        for (int c = 0; c < 4; ++c) L[4 * (size_t)fc[(i + n) % 3] + c] += q[c];
                }
            }
            for (uint32_t v = 0; v < nv; ++v) { /* finishLaplacianSmoothing-comp.glsl */
                float* p = &V[4 * (size_t)v];
                const int32_t* l = &L[4 * (size_t)v];
                if (std::fabs(p[3] - tgt) < 0.00000001f && l[3] > 0)
                    for (int c = 0; c < 3; ++c) {
                        const float avg = (float)l[c] / (float)l[3] / 10000.0f;
                        p[c] = p[c] * (1.0f - weight) + avg * weight;
                    }
            }
        }
    };
    smooth(nb_iters, nb_weight, false);
    smooth(b_iters, b_weight, true);
    std::memcpy(verts, V.data(), V.size() * sizeof(float));
    std::memcpy(faces, F.data(), F.size() * sizeof(uint32_t));
    return ORC_OK;
}

extern "C" void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

extern "C" int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
