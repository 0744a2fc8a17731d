// voxelize.cu — V2: mesh -> occupancy by per-voxel triangle/box separating-axis tests over a binned triangle list.
//
// Replaces RegularGrid::fill(Model3D*) (SRC/DataStructures/RegularGrid.cpp:173-212) with the occupancy predicate the
// north star names: voxel (x,y,z) becomes VOXEL_FREE iff some triangle passes Intersections3D::intersect(Triangle3D&, AABB&)
// (SRC/Geometry/3D/Intersections3D.h:204-258, helpers :260-420) against the voxel box of RegularGrid.cpp:258-259
// (min = aabb.min + cellSize * (x,y,z), max = min + cellSize; centre / half extent per AABB.h:41,51).
//
// The predicate is float32 and operation-order sensitive; this file is compiled with -fmad=false and spells every product
// and sum in the reference's order, so occupancy is bit-identical to the CPU restatement (the 13 axis tests are a pure
// conjunction, so evaluating the three cheap box-axis tests first changes nothing but the early-out).
//
// B200 design: triangles are binned to 4 x 4 x 32 voxel bricks (count -> scan -> fill, ids only); one CTA per non-empty brick
// stages its bin's vertices in shared memory in chunks and every thread tests its two voxels against the chunk; a brick row is
// 32 z-cells = one 64-byte run, written once.  Bricks without triangles are covered by the initial memset (2 B/voxel, the
// algorithmic traffic of this stage: B = 2N + 36T).
#include <vector>

#include "vf_internal.h"

namespace {

constexpr int BX = 8, BY = 8, BZ = 32;  // brick: 64 z-rows of 32 cells
constexpr int kChunk = 64;  // triangles staged per iteration
constexpr unsigned kFull = 0xFFFFFFFFu;

struct VoxGeom {
    int X, Y, Z;
    int nbx, nby, nbz;
    float amin[3], cell[3];
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// candidate voxel range of a triangle: one extra cell each side of its AABB (same rule as the CPU restatement)
__device__ __forceinline__ void tri_range(const VoxGeom& g, const float* p1, const float* p2, const float* p3, int lo[3], int hi[3])
{
    const int dims[3] = { g.X, g.Y, g.Z };
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float tmn = fminf(p1[q], fminf(p2[q], p3[q])), tmx = fmaxf(p1[q], fmaxf(p2[q], p3[q]));
        lo[q] = clampi((int)floorf((tmn - g.amin[q]) / g.cell[q]) - 1, 0, dims[q] - 1);
        hi[q] = clampi((int)floorf((tmx - g.amin[q]) / g.cell[q]) + 1, 0, dims[q] - 1);
    }
}

// One warp per triangle; its lanes take the bricks of the triangle's candidate range (a triangle at 512^3 touches a few dozen), so the
// atomics of one triangle are in flight together instead of one round trip after the other.  The fill pass writes, per (brick, triangle)
// pair, a 64-byte record the voxel kernel consumes without any further indirection:
//   words 0-8 the three vertices, word 9 the candidate range relative to the brick and cut to it (x0 | x1 << 3 | y0 << 6 | y1 << 9 |
//   z0 << 12 | z1 << 17), words 10-14 a conservative plane pre-test (see kPlaneMargin): nx, ny, nz, d, R — cells with |n . c - d| > R lie
//   clearly off the triangle's plane and skip the exact test.
constexpr int kRecWords = 16;
constexpr float kPlaneMargin = 1.05f;  // the exact plane test rejects when the distance exceeds the box's projected radius; 5 % + rounding slack on top
template <bool FILL>
__global__ void __launch_bounds__(256) bin_triangles_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ faces, uint32_t nf, VoxGeom g,
                                                            uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint4* __restrict__ list)
{
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= nf) return;
    const float* p1 = verts + 3 * (size_t)faces[3 * f];
    const float* p2 = verts + 3 * (size_t)faces[3 * f + 1];
    const float* p3 = verts + 3 * (size_t)faces[3 * f + 2];
    int lo[3], hi[3];
    tri_range(g, p1, p2, p3, lo, hi);
    float n[3] = { 0, 0, 0 }, d = 0, R = 3.0e38f;
    if (FILL) {
        const float e0[3] = { p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2] }, e1[3] = { p3[0] - p2[0], p3[1] - p2[1], p3[2] - p2[2] };
        n[0] = e0[1] * e1[2] - e1[1] * e0[2], n[1] = e0[2] * e1[0] - e1[2] * e0[0], n[2] = e0[0] * e1[1] - e1[0] * e0[1];
        const float n1 = fabsf(n[0]) + fabsf(n[1]) + fabsf(n[2]);
        const float l0 = fabsf(e0[0]) + fabsf(e0[1]) + fabsf(e0[2]), l1 = fabsf(e1[0]) + fabsf(e1[1]) + fabsf(e1[2]);
        // slivers (edges parallel to within 1e-3) and specks (edges below a hundredth of a cell): the cross product is mostly rounding, and the
        // exact test's own normal — built from vertex-minus-centre differences — may point elsewhere: no pre-test
        const float speck = 1e-2f * fminf(g.cell[0], fminf(g.cell[1], g.cell[2]));
        if (n1 > 1e-3f * l0 * l1 && n1 < 1e30f && l0 > speck && l1 > speck) {
            d = n[0] * p1[0] + n[1] * p1[1] + n[2] * p1[2];
            float rad = 0, mag = 0;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float ext = g.cell[q] * (float)(q == 0 ? g.X : q == 1 ? g.Y : g.Z);
                rad += fabsf(n[q]) * (0.5f * g.cell[q]);
                mag += fabsf(n[q]) * (fabsf(g.amin[q]) + fabsf(ext));  // bounds |c| and |p| along q
            }
            R = kPlaneMargin * rad + 1e-5f * mag;
        }
    }
    const int bx0 = lo[0] / BX, by0 = lo[1] / BY, bz0 = lo[2] / BZ;
    const int ny = hi[1] / BY - by0 + 1, nz = hi[2] / BZ - bz0 + 1, nn = (hi[0] / BX - bx0 + 1) * ny * nz;
    for (int i = lane; i < nn; i += 32) {
        const int iz = i % nz, iy = (i / nz) % ny, ix = i / (nz * ny);
        const int bx = bx0 + ix, by = by0 + iy, bz = bz0 + iz;
        const uint32_t b = ((uint32_t)bx * g.nby + by) * g.nbz + bz;
        const uint32_t pos = atomicAdd(&counts[b], 1u);
        if (FILL) {
            const int x0 = max(lo[0] - bx * BX, 0), x1 = min(hi[0] - bx * BX, BX - 1), y0 = max(lo[1] - by * BY, 0), y1 = min(hi[1] - by * BY, BY - 1);
            const int z0 = max(lo[2] - bz * BZ, 0), z1 = min(hi[2] - bz * BZ, BZ - 1);
            const uint32_t rng = (uint32_t)(x0 | x1 << 3 | y0 << 6 | y1 << 9 | z0 << 12 | z1 << 17);
            uint4* rec = list + (size_t)(offsets[b] + pos) * (kRecWords / 4);
            rec[0] = make_uint4(__float_as_uint(p1[0]), __float_as_uint(p1[1]), __float_as_uint(p1[2]), __float_as_uint(p2[0]));
            rec[1] = make_uint4(__float_as_uint(p2[1]), __float_as_uint(p2[2]), __float_as_uint(p3[0]), __float_as_uint(p3[1]));
            rec[2] = make_uint4(__float_as_uint(p3[2]), rng, __float_as_uint(n[0]), __float_as_uint(n[1]));
            rec[3] = make_uint4(__float_as_uint(n[2]), __float_as_uint(d), __float_as_uint(R), 0u);
        }
    }
}

// Space for every non-empty brick's records (any order: a brick's list is a set) and the list of the non-empty bricks themselves { brick,
// first record }, which is what the voxel kernel runs over — most bricks of a surface mesh's grid hold no triangle.  One atomic pair per
// warp.  Also resets the counts, which become the fill cursors.  header[0] = records in total, header[1] = non-empty bricks.
__global__ void __launch_bounds__(256) reserve_bricks_kernel(uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t nb, uint2* __restrict__ bricks,
                                                             uint32_t* __restrict__ header)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const uint32_t c = b < nb ? counts[b] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    const unsigned some = __ballot_sync(kFull, c != 0);
    if (some == 0) return;
    uint32_t base = 0, bbase = 0;
    if (lane == 31) base = atomicAdd(&header[0], incl), bbase = atomicAdd(&header[1], (uint32_t)__popc(some));
    base = __shfl_sync(kFull, base, 31), bbase = __shfl_sync(kFull, bbase, 31);
    if (c) {
        offsets[b] = base + incl - c;
        bricks[bbase + __popc(some & ((1u << lane) - 1u))] = make_uint2(b, base + incl - c);
        counts[b] = 0;
    }
}

// Intersections3D.h:204-258 for one voxel box (centre c, half extent r) and one triangle; float32 ops in reference order.
__device__ __forceinline__ bool tri_box_sat(const float c[3], const float r[3], const float* p1, const float* p2, const float* p3)
{
    float v0[3], v1[3], v2[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        v0[q] = p1[q] - c[q];
        v1[q] = p2[q] - c[q];
        v2[q] = p3[q] - c[q];
    }
    // box axes first (:246-253): cheapest rejection, same conjunction
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        float mn = v0[q], mx = v0[q];
        if (v1[q] < mn) mn = v1[q];
        if (v1[q] > mx) mx = v1[q];
        if (v2[q] < mn) mn = v2[q];
        if (v2[q] > mx) mx = v2[q];
        if (mn > r[q] || mx < -r[q]) return false;
    }
    float e0[3], e1[3], e2[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        e0[q] = v1[q] - v0[q];
        e1[q] = v2[q] - v1[q];
        e2[q] = v0[q] - v2[q];
    }
    // plane / box next (:256-257, planeBoxOverlap :269-294; normal = cross(edge0, edge1)): the test that rejects the voxels of a triangle's
    // bounding box that lie off its plane — most of them — before the nine edge axes; the verdict is a conjunction, its order is free
    float n[3];
    n[0] = e0[1] * e1[2] - e1[1] * e0[2];
    n[1] = e0[2] * e1[0] - e1[2] * e0[0];
    n[2] = e0[0] * e1[1] - e1[0] * e0[1];
    float vmin[3], vmax[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float v = v0[q];
        if (n[q] > 0.0f) {
            vmin[q] = -r[q] - v;
            vmax[q] = r[q] - v;
        } else {
            vmin[q] = r[q] - v;
            vmax[q] = -r[q] - v;
        }
    }
    if (n[0] * vmin[0] + n[1] * vmin[1] + n[2] * vmin[2] > 0.0f) return false;
    if (!(n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2] >= 0.0f)) return false;
#define VF_AXIS(pa, pb, rad)                              \
    {                                                     \
        const float pa_ = (pa), pb_ = (pb), rad_ = (rad); \
        const float mn_ = fminf(pa_, pb_), mx_ = fmaxf(pa_, pb_); \
        if (mn_ > rad_ || mx_ < -rad_) return false;      \
    }
    float fx, fy, fz, a, b;
    // edge0: X01 (:296-315), Y02 (:338-357), Z12 (:380-399)
    fx = fabsf(e0[0]), fy = fabsf(e0[1]), fz = fabsf(e0[2]);
    a = e0[2], b = e0[1];
    VF_AXIS(a * v0[1] - b * v0[2], a * v2[1] - b * v2[2], fz * r[1] + fy * r[2]);
    a = e0[2], b = e0[0];
    VF_AXIS(-a * v0[0] + b * v0[2], -a * v2[0] + b * v2[2], fz * r[0] + fx * r[2]);
    a = e0[1], b = e0[0];
    VF_AXIS(a * v1[0] - b * v1[1], a * v2[0] - b * v2[1], fy * r[0] + fx * r[1]);
    // edge1: X01, Y02, Z0 (:401-420)
    fx = fabsf(e1[0]), fy = fabsf(e1[1]), fz = fabsf(e1[2]);
    a = e1[2], b = e1[1];
    VF_AXIS(a * v0[1] - b * v0[2], a * v2[1] - b * v2[2], fz * r[1] + fy * r[2]);
    a = e1[2], b = e1[0];
    VF_AXIS(-a * v0[0] + b * v0[2], -a * v2[0] + b * v2[2], fz * r[0] + fx * r[2]);
    a = e1[1], b = e1[0];
    VF_AXIS(a * v0[0] - b * v0[1], a * v1[0] - b * v1[1], fy * r[0] + fx * r[1]);
    // edge2: X2 (:317-336), Y1 (:359-378), Z12
    fx = fabsf(e2[0]), fy = fabsf(e2[1]), fz = fabsf(e2[2]);
    a = e2[2], b = e2[1];
    VF_AXIS(a * v0[1] - b * v0[2], a * v1[1] - b * v1[2], fz * r[1] + fy * r[2]);
    a = e2[2], b = e2[0];
    VF_AXIS(-a * v0[0] + b * v0[2], -a * v1[0] + b * v1[2], fz * r[0] + fx * r[2]);
    a = e2[1], b = e2[0];
    VF_AXIS(a * v1[0] - b * v1[1], a * v2[0] - b * v2[1], fy * r[0] + fx * r[1]);
#undef VF_AXIS
    return true;
}

// fminf/fmaxf above pick min/max of two finite values exactly like the reference's `if (a < b)` swap (ties give equal values).

// Persistent CTAs (four warps) over the NON-EMPTY bricks.  A brick holds a handful of triangles, so a brick's visit is short and what it
// would wait for is memory latency: the brick after next's header and the next brick's first records are requested before the current brick
// is computed.
//
// Inside a brick the work is dealt by TRIANGLE: a warp takes a record and walks the cells of the triangle's candidate range inside the brick
// (the range the triangle was binned with, tri_range: cells outside it lie a full cell away from the triangle's bounding box and are never
// tested by the restated CPU path either), a power-of-two stretch of z per row so that several rows fill the warp.  Every cell gets two
// cheap conservative tests — its box against the triangle's bounding box, its centre against the triangle's plane (the record's pre-test) —
// and the survivors, a few per row, are queued per warp; the exact separating-axis test then runs 32 queued cells at a time, i.e. with full
// warps, and sets the cell's bit in the brick's bitmap.  The brick is written once, as 64-byte rows.
constexpr int kVoxWarps = 4, kVoxQueue = 96;  // queue: up to 31 waiting + 32 pushed per step, with room to spare
__global__ void __launch_bounds__(kVoxWarps * 32) voxelize_brick_kernel(uint16_t* __restrict__ grid, VoxGeom g, const uint32_t* __restrict__ counts,
                                                                         const uint2* __restrict__ bricks, uint32_t nbricks, const uint4* __restrict__ list)
{
    constexpr int kStage = kVoxWarps * 32 / (kRecWords / 4);  // records staged per chunk: one 16-byte quarter per thread
    __shared__ uint4 rec4[kStage * (kRecWords / 4)];
    __shared__ uint32_t hitbits[BX * BY];  // one word per z-row of the brick
    __shared__ uint32_t queue[kVoxWarps][kVoxQueue];  // row | z << 6 | record << 11
    const float* rec = reinterpret_cast<const float*>(rec4);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t G = gridDim.x;
    uint32_t i = blockIdx.x;
    if (i >= nbricks) return;
    const uint2 none = make_uint2(0u, 0u);
    uint2 h0 = bricks[i], h1 = i + G < nbricks ? bricks[i + G] : none;
    // the list is padded by one chunk: the first chunk of a brick is fetched whole before its length is known
    uint4 pv = list[(size_t)h0.y * (kRecWords / 4) + t];
    // box of a cell: RegularGrid.cpp:258-259 + AABB.h:41,51 (min = aabbMin + cell * index, max = min + cell, centre = (max + min) / 2, half extent = max - centre)
    auto cell_box = [&](int x, int y, int z, float c[3], float r[3]) {
        const float bmin[3] = { g.amin[0] + g.cell[0] * (float)x, g.amin[1] + g.cell[1] * (float)y, g.amin[2] + g.cell[2] * (float)z };
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float hi = bmin[q] + g.cell[q];
            c[q] = (hi + bmin[q]) / 2.0f;
            r[q] = hi - c[q];
        }
    };
    // slack of the bounding-box pre-test: a cell passes when its box, grown by 2 % of a cell plus rounding room, reaches the triangle's box
    float grow[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) grow[q] = 0.51f * g.cell[q] + 1e-5f * (fabsf(g.amin[q]) + g.cell[q] * (float)(q == 0 ? g.X : q == 1 ? g.Y : g.Z));
    for (; i < nbricks; i += G) {
        const uint2 h2 = i + 2 * G < nbricks ? bricks[i + 2 * G] : none;
        uint4 pvn = make_uint4(0u, 0u, 0u, 0u);
        if (i + G < nbricks) pvn = list[(size_t)h1.y * (kRecWords / 4) + t];
        const uint32_t b = h0.x, cnt = counts[b];  // after the fill pass: the number of records of the brick
        const int bz = b % g.nbz, by = (b / g.nbz) % g.nby, bx = b / (g.nbz * g.nby);
        const int gx0 = bx * BX, gy0 = by * BY, gz0 = bz * BZ;
        if (t < BX * BY) hitbits[t] = 0;
        uint32_t* q = queue[warp];
        int qn = 0;
        // exact test of the queued cells 32 at a time (`all`: also the last, partial batch)
        auto drain = [&](bool all) {
            while (qn >= 32 || (all && qn > 0)) {
                const int take = min(qn, 32);
                qn -= take;
                if (lane < take) {
                    const uint32_t e = q[qn + lane];
                    const int row = e & 63, z = e >> 6 & 31, k = e >> 11;
                    if (!(hitbits[row] >> z & 1u)) {
                        float c[3], r[3];
                        cell_box(gx0 + (row >> 3), gy0 + (row & 7), gz0 + z, c, r);
                        const float* tr = rec + k * kRecWords;
                        if (tri_box_sat(c, r, tr, tr + 3, tr + 6)) atomicOr(&hitbits[row], 1u << z);
                    }
                }
                __syncwarp();
            }
        };
        for (uint32_t base = 0; base < cnt; base += kStage) {
            const int m = (int)min((uint32_t)kStage, cnt - base);
            __syncthreads();  // the previous chunk has been read (and hitbits is cleared)
            if (t < m * (kRecWords / 4)) rec4[t] = base == 0 ? pv : list[(size_t)(h0.y + base) * (kRecWords / 4) + t];
            __syncthreads();
            for (int k = warp; k < m; k += kVoxWarps) {
                const float* tr = rec + k * kRecWords;
                const uint32_t rr = __float_as_uint(tr[9]);
                const int x0 = rr & 7, x1 = rr >> 3 & 7, y0 = rr >> 6 & 7, y1 = rr >> 9 & 7, z0 = rr >> 12 & 31, z1 = rr >> 17 & 31;
                const int ey = y1 - y0 + 1, ez = z1 - z0 + 1, rows = (x1 - x0 + 1) * ey;
                const int zsh = ez <= 1 ? 0 : 32 - __clz(ez - 1);  // log2 of the z stretch (ez rounded up to a power of two)
                const int rps = 32 >> zsh, zl = lane & ((1 << zsh) - 1), rl = lane >> zsh;
                const float inv_ey = 1.0f / (float)ey;
                // the triangle's bounding box, for the first pre-test
                float tmin[3], tmax[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    tmin[a] = fminf(tr[a], fminf(tr[3 + a], tr[6 + a]));
                    tmax[a] = fmaxf(tr[a], fmaxf(tr[3 + a], tr[6 + a]));
                }
                const float nx = tr[10], ny = tr[11], nz = tr[12], nd = tr[13], R = tr[14];
                for (int r0 = 0; r0 < rows; r0 += rps) {
                    const int rho = r0 + rl;
                    const int ix = (int)(((float)rho + 0.5f) * inv_ey), iy = rho - ix * ey;  // rho < 64, ey <= 8: exact
                    bool cand = rho < rows && zl < ez;
                    const int cx = x0 + ix, cy = y0 + iy, cz = z0 + zl;
                    if (cand) {
                        const int x = gx0 + cx, y = gy0 + cy, z = gz0 + cz;
                        // centres as the exact test computes them, to within an ulp (the pre-tests carry their own slack)
                        const float px = g.amin[0] + g.cell[0] * ((float)x + 0.5f), py = g.amin[1] + g.cell[1] * ((float)y + 0.5f), pz = g.amin[2] + g.cell[2] * ((float)z + 0.5f);
                        cand = px + grow[0] >= tmin[0] && px - grow[0] <= tmax[0] && py + grow[1] >= tmin[1] && py - grow[1] <= tmax[1] &&
                               pz + grow[2] >= tmin[2] && pz - grow[2] <= tmax[2] && fabsf(nx * px + ny * py + nz * pz - nd) <= R &&
                               x < g.X && y < g.Y && z < g.Z;
                    }
                    const unsigned bal = __ballot_sync(kFull, cand);
                    if (bal) {
                        if (cand) q[qn + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)(cx * BY + cy) | (uint32_t)cz << 6 | (uint32_t)k << 11;
                        qn += __popc(bal);
                        __syncwarp();
                        drain(false);
                    }
                }
            }
            drain(true);  // before the records of this chunk are overwritten
        }
        __syncthreads();
        // the brick's rows: 64 rows x four 16-byte quarters
        for (int it = t; it < BX * BY * 4; it += kVoxWarps * 32) {
            const int row = it >> 2, qz = (it & 3) * 8;
            const int x = gx0 + (row >> 3), y = gy0 + (row & 7), z = gz0 + qz;
            if (x >= g.X || y >= g.Y || z >= g.Z) continue;
            const uint32_t bits = hitbits[row] >> qz;
            uint32_t w[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) w[a] = (bits >> (2 * a) & 1u) | (bits >> (2 * a + 1) & 1u) << 16;  // VF_VOXEL_FREE == 1
            uint16_t* p = grid + ((size_t)x * g.Y + y) * g.Z + z;
            if (z + 8 <= g.Z && ((uintptr_t)p & 15) == 0) {
                *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
                for (int a = 0; a < 8 && z + a < g.Z; ++a) p[a] = (uint16_t)(bits >> a & 1u);
            }
        }
        h0 = h1, h1 = h2, pv = pvn;
        __syncthreads();  // everybody has read hitbits before the next brick clears it
    }
}

}  // namespace

extern "C" vf_status vf_voxelize(vf_grid* grid, const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf)
{
    VF_REQUIRE(grid && verts && faces && nv > 0, VF_ERR_INVALID_ARGUMENT, "voxelize: null or empty mesh");
    vf_ctx* c = grid->ctx;
    VF_TRY(vf_enter(c));
    for (uint32_t i = 0; i < 3 * nf; ++i) VF_REQUIRE(faces[i] < nv, VF_ERR_INVALID_ARGUMENT, "face %u references vertex %u >= %u", i / 3, faces[i], nv);
    VoxGeom g;
    g.X = (int)grid->X, g.Y = (int)grid->Y, g.Z = (int)grid->Z;
    g.nbx = (g.X + BX - 1) / BX, g.nby = (g.Y + BY - 1) / BY, g.nbz = (g.Z + BZ - 1) / BZ;
    const uint32_t dims[3] = { grid->X, grid->Y, grid->Z };
    for (int q = 0; q < 3; ++q) {
        g.amin[q] = grid->aabb_min[q];
        g.cell[q] = (grid->aabb_max[q] - grid->aabb_min[q]) / (float)dims[q];  // RegularGrid.cpp:438
    }
    const size_t nb = (size_t)g.nbx * g.nby * g.nbz;
    VF_CUDA(cudaMemsetAsync(grid->d, 0, grid->n() * sizeof(uint16_t), c->stream));  // cleanGrid + every brick without triangles
    if (nf == 0) return VF_OK;

    // arena: verts | faces | counts[nb] | offsets[nb] | non-empty bricks[nb] | header | records (+ one chunk of padding)
    const size_t vbytes = ((size_t)nv * 12 + 255) & ~(size_t)255, fbytes = ((size_t)nf * 12 + 255) & ~(size_t)255;
    const size_t cbytes = (nb * 4 + 255) & ~(size_t)255;
    const size_t need = vbytes + fbytes + 4 * cbytes + 256, rec_bytes = kRecWords * 4, pad = (size_t)64 * rec_bytes;
    VF_TRY(vf_scratch_reserve(c, c->mesh, need + (size_t)nf * 16 * rec_bytes + pad));
    char* base = (char*)c->mesh.ptr;
    float* d_verts = (float*)base;
    uint32_t* d_faces = (uint32_t*)(base + vbytes);
    uint32_t* d_counts = (uint32_t*)(base + vbytes + fbytes);
    uint32_t* d_offsets = (uint32_t*)(base + vbytes + fbytes + cbytes);
    uint2* d_bricks = (uint2*)(base + vbytes + fbytes + 2 * cbytes);
    uint32_t* d_header = (uint32_t*)(base + vbytes + fbytes + 4 * cbytes);
    VF_CUDA(cudaMemcpyAsync(d_verts, verts, (size_t)nv * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemcpyAsync(d_faces, faces, (size_t)nf * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemsetAsync(d_counts, 0, cbytes, c->stream));
    VF_CUDA(cudaMemsetAsync(d_header, 0, 8, c->stream));
    const int tb = (int)((nf + 7) / 8);  // a warp per triangle
    bin_triangles_kernel<false><<<tb, 256, 0, c->stream>>>(d_verts, d_faces, nf, g, d_counts, nullptr, nullptr);
    VF_LAUNCHED(c);
    reserve_bricks_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(d_counts, d_offsets, (uint32_t)nb, d_bricks, d_header);
    VF_LAUNCHED(c);
    uint32_t* h_header = (uint32_t*)((char*)c->pinned + 65536);
    VF_CUDA(cudaMemcpyAsync(h_header, d_header, 8, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));  // also covers the pageable verts/faces uploads
    const size_t total = h_header[0];
    const uint32_t nonempty = h_header[1];
    VF_REQUIRE(total < (1ull << 27), VF_ERR_CAPACITY, "voxelize: triangle/brick list too long (%zu records)", total);
    if (nonempty == 0) return VF_OK;
    if (need + total * rec_bytes + pad > c->mesh.bytes) {
        // grow the arena, keeping what the passes above left in it (rare: only when the mesh has very large triangles)
        VfScratch old = c->mesh;
        c->mesh = VfScratch();
        const vf_status grown = vf_scratch_reserve(c, c->mesh, need + total * rec_bytes + pad + 256);
        if (grown != VF_OK) {  // keep the old arena: nothing is lost, the call fails
            c->mesh = old;
            return grown;
        }
        cudaError_t ce = cudaMemcpyAsync(c->mesh.ptr, old.ptr, need, cudaMemcpyDeviceToDevice, c->stream);
        if (ce == cudaSuccess) ce = vf_sync(c);
        cudaFree(old.ptr);  // the new arena is the context's either way
        VF_CUDA(ce);
        base = (char*)c->mesh.ptr;
        d_verts = (float*)base;
        d_faces = (uint32_t*)(base + vbytes);
        d_counts = (uint32_t*)(base + vbytes + fbytes);
        d_offsets = (uint32_t*)(base + vbytes + fbytes + cbytes);
        d_bricks = (uint2*)(base + vbytes + fbytes + 2 * cbytes);
    }
    uint4* d_list = (uint4*)(base + need);
    bin_triangles_kernel<true><<<tb, 256, 0, c->stream>>>(d_verts, d_faces, nf, g, d_counts, d_offsets, d_list);
    VF_LAUNCHED(c);
    int per_sm = 0;
    VF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, voxelize_brick_kernel, kVoxWarps * 32, 0));
    const unsigned ctas = std::min<unsigned>(nonempty, (unsigned)(c->num_sms * std::max(per_sm, 1)));
    voxelize_brick_kernel<<<ctas, kVoxWarps * 32, 0, c->stream>>>(grid->d, g, d_counts, d_bricks, nonempty, d_list);
    VF_LAUNCHED(c);
    return VF_OK;
}

// =========================================================================================================================
// V1: solid occupancy — the LIVE path of RegularGrid::fill(Model3D*) (RegularGrid.cpp:173-212): Tetravoxelizer
// (SRC/Graphics/Core/Tetravoxelizer.{h,cpp}).  Every face and the vertex-average centroid form a tetrahedron in the grid AABB's
// NDC space (initializeModel :198-247); for every y-slice the geometry shader (:42-92) cuts it with the plane y = ySlice
// (accumulated float32, :282-299) into one or two (x, z) triangles that are drawn with glLogicOp(GL_XOR) (:271-272), so a cell
// ends up FREE iff an odd number of tetrahedra cover (cell centre x, slice plane y, cell centre z).  The reference needs one draw +
// glReadPixels + glFinish per slice (:288-304) and a host pass over the uint8[y][z][x] result (RegularGrid.cpp:186-199).
//
// Pixel coverage in the reference is whatever the GL rasteriser of the machine decides (implementation-defined snapping and
// fill rule): parity is unpinned by construction.  The rule implemented here — identical in the CPU checker — is: shader
// arithmetic in float32 with mix(a,b,t) = a*(1-t) + b*t (this file is built with -fmad=false), window coordinates snapped to
// 1/256 pixel (round half even), a centre on an edge belongs to the triangle iff the edge's counter-clockwise direction has
// dy > 0 or (dy == 0 and dx < 0), 64-bit integer edge functions.
//
// B200 design: the unit of work is a (tetrahedron, slice) pair.  One warp owns a tetrahedron (its 12 floats are loaded once,
// its slice range comes from a binary search in the table of accumulated slice planes) and its lanes take the slices 32 at a
// time, so every lane cuts and draws its own cross-section; a face is a few cells wide, so a cross-section is a handful of
// pixels, walked row by row with incremental 64-bit edge functions and XOR-ed into a bit grid (1 bit/voxel, L2 resident) with
// one atomic per touched 32-bit word.  Cross-sections with a large box are voted out (ballot) and drawn by the whole warp, one
// lane per row.  A last pass expands the bit grid to uint16 labels, coalesced, and counts the FREE cells: 2 B/voxel written,
// N/8 B of bits written and read, 48 B/tetrahedron read — no per-slice draw call, read-back or host pass.
namespace {

struct SolidGeom {
    int X, Y, Z;
    int zw;       // 32-bit words per (x, slice) row of the bit grid
    float ctr[3], dim[3], cen[3];  // AABB centre and size, centroid in NDC
    float hx, hz; // X/2, Z/2
};

// per face: the tetrahedron (face + centroid) in NDC sorted by y, and the slices s with A.y < ys[s] <= D.y (geometry shader :58)
__global__ void __launch_bounds__(256) tetra_setup_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ faces, uint32_t nf, SolidGeom g,
                                                          const float* __restrict__ ys, float4* __restrict__ tets, int2* __restrict__ srange)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    float v[4][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* p = verts + 3 * (size_t)faces[3 * f + k];
#pragma unroll
        for (int q = 0; q < 3; ++q) v[k][q] = (2.0f * (p[q] - g.ctr[q])) / g.dim[q];  // scaleToNDC, Tetravoxelizer.h:70-72
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) v[3][q] = g.cen[q];
#define VF_CSWAP(a, b)                                    \
    if (v[a][1] > v[b][1]) {                              \
        _Pragma("unroll") for (int q = 0; q < 3; ++q) {   \
            const float t_ = v[a][q];                     \
            v[a][q] = v[b][q], v[b][q] = t_;              \
        }                                                 \
    }
    VF_CSWAP(0, 1) VF_CSWAP(2, 3) VF_CSWAP(0, 2) VF_CSWAP(1, 3) VF_CSWAP(1, 2)  // sort4Vec3ByLowerY, Tetravoxelizer.h:75-81
#undef VF_CSWAP
    tets[3 * (size_t)f + 0] = make_float4(v[0][0], v[0][1], v[0][2], v[1][0]);
    tets[3 * (size_t)f + 1] = make_float4(v[1][1], v[1][2], v[2][0], v[2][1]);
    tets[3 * (size_t)f + 2] = make_float4(v[2][2], v[3][0], v[3][1], v[3][2]);
    // ys is strictly increasing: first slice above A.y, last slice not above D.y
    int lo = 0, hi = g.Y;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (ys[m] > v[0][1]) hi = m; else lo = m + 1;
    }
    const int s0 = lo;
    lo = 0, hi = g.Y;
    while (lo < hi) {
        const int m = (lo + hi) >> 1;
        if (ys[m] <= v[3][1]) lo = m + 1; else hi = m;
    }
    srange[f] = make_int2(s0, lo - 1);
}

__device__ __forceinline__ int solid_snap(float w)
{
    float f = rintf(w * 256.0f);
    if (!(f > -268435456.0f)) f = -268435456.0f;
    if (f > 268435456.0f) f = 268435456.0f;
    return (int)f;
}

// a cross-section triangle in 24.8 fixed-point window coordinates, made counter-clockwise, with its clipped pixel box
struct SolidTri {
    int x0, y0, x1, y1, x2, y2;
    int i0, i1, k0, k1;  // rows (x) and columns (z) of the box; i1 < i0 = nothing to draw
};

__device__ __forceinline__ void solid_tri_finish(SolidTri& t, const SolidGeom& g)
{
    const long long area2 = (long long)(t.x1 - t.x0) * (t.y2 - t.y0) - (long long)(t.y1 - t.y0) * (t.x2 - t.x0);
    t.i0 = 0, t.i1 = -1, t.k0 = 0, t.k1 = -1;
    if (area2 == 0) return;
    if (area2 < 0) {
        const int tx = t.x1, ty = t.y1;
        t.x1 = t.x2, t.y1 = t.y2, t.x2 = tx, t.y2 = ty;
    }
    const int minx = min(t.x0, min(t.x1, t.x2)), maxx = max(t.x0, max(t.x1, t.x2));
    const int miny = min(t.y0, min(t.y1, t.y2)), maxy = max(t.y0, max(t.y1, t.y2));
    const int i0 = max(0, (minx + 127) >> 8), i1 = min(g.X - 1, (maxx - 128) >> 8);  // pixel centre = 256 i + 128
    const int k0 = max(0, (miny + 127) >> 8), k1 = min(g.Z - 1, (maxy - 128) >> 8);
    if (i0 > i1 || k0 > k1) return;
    t.i0 = i0, t.i1 = i1, t.k0 = k0, t.k1 = k1;
}

__device__ __forceinline__ int solid_rows(const SolidTri& t) { return t.i1 - t.i0 + 1; }

// XOR rows i_first, i_first + i_step, ... of the triangle's box into the bit rows of slice s.  A centre is covered iff all three
// edge functions (+1 on the edges the triangle owns) are positive.  Along a row each of them is linear in the column, so it
// bounds the covered columns from one side: the bound is estimated in float32 and then moved to the exact integer answer with
// the 64-bit function itself, and the row costs three such solves whatever the width of the box (cross-sections are thin
// diagonal slivers: their boxes hold ten times more pixels than they cover).
__device__ __forceinline__ void solid_draw(uint32_t* __restrict__ bits, const SolidGeom& g, int s, const SolidTri& t, int i_first, int i_step)
{
    if (i_first > t.i1) return;
    const int dx[3] = { t.x1 - t.x0, t.x2 - t.x1, t.x0 - t.x2 }, dy[3] = { t.y1 - t.y0, t.y2 - t.y1, t.y0 - t.y2 };
    const int vx[3] = { t.x0, t.x1, t.x2 }, vy[3] = { t.y0, t.y1, t.y2 };
    const int px = 256 * i_first + 128, py = 256 * t.k0 + 128, bw = t.k1 - t.k0 + 1;
    long long r[3], c[3], n[3];
    float rc[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        r[j] = (long long)dx[j] * (py - vy[j]) - (long long)dy[j] * (px - vx[j]) + ((dy[j] > 0 || (dy[j] == 0 && dx[j] < 0)) ? 1 : 0);
        c[j] = 256ll * dx[j];                  // one column (z) further
        n[j] = -256ll * dy[j] * i_step;        // this lane's next row
        rc[j] = dx[j] ? -1.0f / (float)c[j] : 0.0f;
    }
    for (int i = i_first; i <= t.i1; i += i_step) {
        int lo = 0, hi = bw - 1;  // covered columns, relative to k0
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const long long R = r[j], Cc = c[j];
            if (Cc == 0) {
                if (R <= 0) hi = -1;
                continue;
            }
            const float q = fminf(fmaxf((float)R * rc[j], -2.0f), (float)bw + 1.0f);  // zero crossing, approximately
            if (Cc > 0) {  // smallest column with R + k Cc > 0
                int k = (int)floorf(q) + 1;
                k = max(0, min(k, bw));
                while (k > 0 && R + (long long)(k - 1) * Cc > 0) --k;
                while (k < bw && R + (long long)k * Cc <= 0) ++k;
                lo = max(lo, k);
            } else {       // largest column with R + k Cc > 0
                int k = (int)ceilf(q) - 1;
                k = max(-1, min(k, bw - 1));
                while (k < bw - 1 && R + (long long)(k + 1) * Cc > 0) ++k;
                while (k >= 0 && R + (long long)k * Cc <= 0) --k;
                hi = min(hi, k);
            }
        }
        if (lo <= hi) {
            uint32_t* row = bits + ((size_t)i * g.Y + s) * g.zw;
            const int ka = t.k0 + lo, kb = t.k0 + hi;
            for (int w = ka >> 5; w <= (kb >> 5); ++w) {
                const int a = max(ka, 32 * w) & 31, b = min(kb, 32 * w + 31) & 31;
                atomicXor(row + w, (0xFFFFFFFFu >> (31 - b + a)) << a);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) r[j] += n[j];
    }
}

__device__ __forceinline__ SolidTri solid_bcast(const SolidTri& t, int src)
{
    SolidTri o;
    o.x0 = __shfl_sync(kFull, t.x0, src), o.y0 = __shfl_sync(kFull, t.y0, src);
    o.x1 = __shfl_sync(kFull, t.x1, src), o.y1 = __shfl_sync(kFull, t.y1, src);
    o.x2 = __shfl_sync(kFull, t.x2, src), o.y2 = __shfl_sync(kFull, t.y2, src);
    o.i0 = __shfl_sync(kFull, t.i0, src), o.i1 = __shfl_sync(kFull, t.i1, src);
    o.k0 = __shfl_sync(kFull, t.k0, src), o.k1 = __shfl_sync(kFull, t.k1, src);
    return o;
}

constexpr int kSolidSelf = 48;  // cross-sections whose boxes have at most this many rows are drawn by the lane that cut them
                               // (the slices of one tetrahedron have similar cross-sections, so the lanes of a warp stay in step)

__global__ void __launch_bounds__(256) solid_scatter_kernel(uint32_t* __restrict__ bits, const float4* __restrict__ tets, const int2* __restrict__ srange,
                                                            uint32_t nf, const float* __restrict__ ys, SolidGeom g)
{
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= nf) return;  // whole warps leave together
    const int lane = threadIdx.x & 31;
    const int2 sr = srange[f];
    if (sr.y < sr.x) return;
    const float4 q0 = tets[3 * (size_t)f], q1 = tets[3 * (size_t)f + 1], q2 = tets[3 * (size_t)f + 2];
    const float A[3] = { q0.x, q0.y, q0.z }, B[3] = { q0.w, q1.x, q1.y }, C[3] = { q1.z, q1.w, q2.x }, D[3] = { q2.y, q2.z, q2.w };
    for (int base = sr.x; base <= sr.y; base += 32) {
        const int s = base + lane;
        SolidTri t1, t2;
        t1.i0 = t2.i0 = 0, t1.i1 = t2.i1 = -1;
        if (s <= sr.y) {
            const float sl = ys[s];
            // INTERP(P, Q, s).xz (:46) -> window coordinates of the X x Z viewport (:177) -> 1/256 pixel
            auto cut = [&](const float* P, const float* Q, int& wx, int& wy) {
                const float w = (sl - P[1]) / (Q[1] - P[1]);
                const float x = P[0] * (1.0f - w) + Q[0] * w, z = P[2] * (1.0f - w) + Q[2] * w;
                wx = solid_snap(x * g.hx + g.hx), wy = solid_snap(z * g.hz + g.hz);
            };
            cut(A, D, t1.x0, t1.y0);
            if (sl <= B[1]) cut(A, B, t1.x1, t1.y1); else cut(B, D, t1.x1, t1.y1);  // v1 (:62-64)
            if (sl <= C[1]) cut(A, C, t1.x2, t1.y2); else cut(C, D, t1.x2, t1.y2);  // v2 (:68-70)
            if (B[1] < sl && sl <= C[1]) {                                          // extra triangle (:77)
                cut(B, C, t2.x0, t2.y0);
                t2.x1 = t1.x2, t2.y1 = t1.y2, t2.x2 = t1.x1, t2.y2 = t1.y1;
                solid_tri_finish(t2, g);
            }
            solid_tri_finish(t1, g);
        }
        const bool self = max(solid_rows(t1), solid_rows(t2)) <= kSolidSelf;
        if (self) {
            solid_draw(bits, g, s, t1, t1.i0, 1);
            solid_draw(bits, g, s, t2, t2.i0, 1);
        }
        unsigned big = __ballot_sync(kFull, !self);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const SolidTri a = solid_bcast(t1, src), b = solid_bcast(t2, src);
            solid_draw(bits, g, base + src, a, a.i0 + lane, 32);
            solid_draw(bits, g, base + src, b, b.i0 + lane, 32);
        }
    }
}

// bit grid -> labels (RegularGrid.cpp:186-199 after cleanGrid :591-599) + FREE count; one thread per 8 cells when Z % 8 == 0
__global__ void __launch_bounds__(256) solid_expand_kernel(uint16_t* __restrict__ grid, const uint32_t* __restrict__ bits, SolidGeom g, unsigned long long* __restrict__ occupied)
{
    const size_t rows = (size_t)g.X * g.Y;
    unsigned cnt = 0;
    if ((g.Z & 7) == 0) {
        const int vec_per_row = g.Z >> 3;
        const size_t total = rows * vec_per_row;
        for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
            const size_t r = e / vec_per_row;
            const int c = (int)(e - r * vec_per_row);
            const uint32_t b = (bits[r * g.zw + (c >> 2)] >> ((c & 3) * 8)) & 0xFFu;
            cnt += __popc(b);
            uint4 o;
            o.x = (b & 1u) | ((b & 2u) << 15), o.y = ((b >> 2) & 1u) | ((b & 8u) << 13);
            o.z = ((b >> 4) & 1u) | ((b & 32u) << 11), o.w = ((b >> 6) & 1u) | ((b & 128u) << 9);
            *reinterpret_cast<uint4*>(grid + r * g.Z + (size_t)c * 8) = o;
        }
    } else {
        const size_t total = rows * g.Z;
        for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
            const size_t r = e / g.Z;
            const int k = (int)(e - r * g.Z);
            const uint32_t b = (bits[r * g.zw + (k >> 5)] >> (k & 31)) & 1u;
            cnt += b;
            grid[e] = (uint16_t)b;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(occupied, (unsigned long long)cnt);
}

}  // namespace

extern "C" vf_status vf_voxelize_solid(vf_grid* grid, const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, uint64_t* occupied_out)
{
    VF_REQUIRE(grid && verts && faces && nv > 0, VF_ERR_INVALID_ARGUMENT, "voxelize_solid: null or empty mesh");
    vf_ctx* c = grid->ctx;
    VF_TRY(vf_enter(c));
    for (uint32_t i = 0; i < 3 * nf; ++i) VF_REQUIRE(faces[i] < nv, VF_ERR_INVALID_ARGUMENT, "face %u references vertex %u >= %u", i / 3, faces[i], nv);
    if (occupied_out) *occupied_out = 0;
    if (nf == 0) {
        VF_CUDA(cudaMemsetAsync(grid->d, 0, grid->n() * sizeof(uint16_t), c->stream));
        return VF_OK;
    }
    SolidGeom g;
    g.X = (int)grid->X, g.Y = (int)grid->Y, g.Z = (int)grid->Z;
    g.zw = (g.Z + 31) / 32;
    // initializeModel :204-217: centroid = float32 sum in vertex order / count, then scaleToNDC
    float cen[3] = { 0.0f, 0.0f, 0.0f };
    for (uint32_t i = 0; i < nv; ++i)
        for (int q = 0; q < 3; ++q) cen[q] += verts[3 * (size_t)i + q];
    for (int q = 0; q < 3; ++q) {
        cen[q] /= (float)nv;
        g.dim[q] = grid->aabb_max[q] - grid->aabb_min[q];
        g.ctr[q] = 0.5f * (grid->aabb_min[q] + grid->aabb_max[q]);
        g.cen[q] = (2.0f * (cen[q] - g.ctr[q])) / g.dim[q];
    }
    g.hx = (float)g.X * 0.5f, g.hz = (float)g.Z * 0.5f;

    // arena: verts | faces | tets[nf][3] float4 | srange[nf] int2 | ys[Y] | occupied | bit grid
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t vbytes = up((size_t)nv * 12), fbytes = up((size_t)nf * 12), tbytes = up((size_t)nf * 48), rbytes = up((size_t)nf * 8), ybytes = up((size_t)g.Y * 4);
    const size_t bbytes = up((size_t)g.X * g.Y * g.zw * 4);
    VF_TRY(vf_scratch_reserve(c, c->mesh, vbytes + fbytes + tbytes + rbytes + ybytes + 256 + bbytes));
    char* base = (char*)c->mesh.ptr;
    float* d_verts = (float*)base;
    uint32_t* d_faces = (uint32_t*)(base + vbytes);
    float4* d_tets = (float4*)(base + vbytes + fbytes);
    int2* d_sr = (int2*)(base + vbytes + fbytes + tbytes);
    float* d_ys = (float*)(base + vbytes + fbytes + tbytes + rbytes);
    unsigned long long* d_occ = (unsigned long long*)(base + vbytes + fbytes + tbytes + rbytes + ybytes);
    uint32_t* d_bits = (uint32_t*)(base + vbytes + fbytes + tbytes + rbytes + ybytes + 256);
    // slice planes: ySlice starts at -1 and is accumulated (compute :282-299)
    std::vector<float> ys((size_t)g.Y);
    {
        float y = -1.0f;
        const float step = 2.0f / (float)g.Y;
        for (int s = 0; s < g.Y; ++s) ys[s] = y, y += step;
    }
    VF_CUDA(cudaMemcpyAsync(d_verts, verts, (size_t)nv * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemcpyAsync(d_faces, faces, (size_t)nf * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemcpyAsync(d_ys, ys.data(), (size_t)g.Y * 4, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemsetAsync(d_occ, 0, 256 + bbytes, c->stream));  // counter + bit grid
    tetra_setup_kernel<<<(nf + 255) / 256, 256, 0, c->stream>>>(d_verts, d_faces, nf, g, d_ys, d_tets, d_sr);
    VF_LAUNCHED(c);
    solid_scatter_kernel<<<(nf + 7) / 8, 256, 0, c->stream>>>(d_bits, d_tets, d_sr, nf, d_ys, g);  // one warp per tetrahedron
    VF_LAUNCHED(c);
    {
        const size_t units = (g.Z & 7) == 0 ? grid->n() / 8 : grid->n();
        const unsigned blocks = (unsigned)std::min<size_t>((units + 255) / 256, (size_t)c->num_sms * 16);
        solid_expand_kernel<<<blocks, 256, 0, c->stream>>>(grid->d, d_bits, g, d_occ);
        VF_LAUNCHED(c);
    }
    unsigned long long* h_occ = (unsigned long long*)((char*)c->pinned + 65536 + 64);
    VF_CUDA(cudaMemcpyAsync(h_occ, d_occ, 8, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));  // also covers the pageable uploads (verts, faces, ys)
    if (occupied_out) *occupied_out = *h_occ;
    return VF_OK;
}
