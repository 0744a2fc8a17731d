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

constexpr int BX = 4, BY = 4, BZ = 32;
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

template <bool FILL>
__global__ void __launch_bounds__(256) bin_triangles_kernel(const float* __restrict__ verts, const uint32_t* __restrict__ faces, uint32_t nf, VoxGeom g,
                                                            uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ list)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const float* p1 = verts + 3 * (size_t)faces[3 * f];
    const float* p2 = verts + 3 * (size_t)faces[3 * f + 1];
    const float* p3 = verts + 3 * (size_t)faces[3 * f + 2];
    int lo[3], hi[3];
    tri_range(g, p1, p2, p3, lo, hi);
    for (int bx = lo[0] / BX; bx <= hi[0] / BX; ++bx)
        for (int by = lo[1] / BY; by <= hi[1] / BY; ++by)
            for (int bz = lo[2] / BZ; bz <= hi[2] / BZ; ++bz) {
                const uint32_t b = ((uint32_t)bx * g.nby + by) * g.nbz + bz;
                const uint32_t pos = atomicAdd(&counts[b], 1u);
                if (FILL) list[offsets[b] + pos] = f;
            }
}

// exclusive scan of the brick counts by one CTA (a few hundred thousand entries at most); also resets the fill cursors
__global__ void __launch_bounds__(1024) scan_counts_kernel(uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(kFull, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(kFull, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + s - v;
        if (i < n) {
            offsets[i] = before;
            counts[i] = 0;  // becomes the fill cursor
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
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
    // plane / box (:256-257, planeBoxOverlap :269-294); normal = cross(edge0, edge1)
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
    return n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2] >= 0.0f;
}

// fminf/fmaxf above pick min/max of two finite values exactly like the reference's `if (a < b)` swap (ties give equal values).

__global__ void __launch_bounds__(256) voxelize_brick_kernel(uint16_t* __restrict__ grid, const float* __restrict__ verts, const uint32_t* __restrict__ faces,
                                                             VoxGeom g, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
                                                             const uint32_t* __restrict__ list)
{
    __shared__ float tri[kChunk][9];
    const uint32_t b = blockIdx.x;
    // after the fill pass `counts` holds the number of ids written per brick
    const uint32_t cnt = counts[b];
    if (cnt == 0) return;
    const int bz = b % g.nbz, by = (b / g.nbz) % g.nby, bx = b / (g.nbz * g.nby);
    // thread -> two z-adjacent voxels: 16 threads per 32-cell row, 16 rows
    const int t = threadIdx.x, row = t >> 4, zq = (t & 15) * 2;
    const int x = bx * BX + row / BY, y = by * BY + row % BY, z = bz * BZ + zq;
    const bool in0 = x < g.X && y < g.Y && z < g.Z, in1 = in0 && z + 1 < g.Z;
    float c0[3], c1[3], r0[3], r1[3];
    {
        // RegularGrid.cpp:258-259 + AABB.h:41,51
        const float bmin[3] = { g.amin[0] + g.cell[0] * (float)x, g.amin[1] + g.cell[1] * (float)y, g.amin[2] + g.cell[2] * (float)z };
        const float bmin1z = g.amin[2] + g.cell[2] * (float)(z + 1);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float lo = bmin[q];
            const float hi = lo + g.cell[q];
            c0[q] = (hi + lo) / 2.0f;
            r0[q] = hi - c0[q];
            c1[q] = c0[q];
            r1[q] = r0[q];
        }
        const float hi1 = bmin1z + g.cell[2];
        c1[2] = (hi1 + bmin1z) / 2.0f;
        r1[2] = hi1 - c1[2];
    }
    bool hit0 = false, hit1 = false;
    const uint32_t off = offsets[b];
    for (uint32_t base = 0; base < cnt; base += kChunk) {
        const int m = (int)min((uint32_t)kChunk, cnt - base);
        __syncthreads();
        for (int i = t; i < m * 9; i += blockDim.x) {
            const int k = i / 9, e = i % 9;
            const uint32_t f = list[off + base + k];
            tri[k][e] = verts[3 * (size_t)faces[3 * f + e / 3] + e % 3];
        }
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            if (in0 && !hit0) hit0 = tri_box_sat(c0, r0, &tri[k][0], &tri[k][3], &tri[k][6]);
            if (in1 && !hit1) hit1 = tri_box_sat(c1, r1, &tri[k][0], &tri[k][3], &tri[k][6]);
        }
    }
    if (in0) {
        const size_t gi = ((size_t)x * g.Y + y) * g.Z + z;
        if (in1 && (gi & 1) == 0) {
            *reinterpret_cast<uint32_t*>(grid + gi) = (hit0 ? 1u : 0u) | (hit1 ? 0x10000u : 0u);
        } else {
            grid[gi] = hit0 ? VF_VOXEL_FREE : VF_VOXEL_EMPTY;
            if (in1) grid[gi + 1] = hit1 ? VF_VOXEL_FREE : VF_VOXEL_EMPTY;
        }
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

    // arena: verts | faces | counts[nb] | offsets[nb] | total | list
    const size_t vbytes = ((size_t)nv * 12 + 255) & ~(size_t)255, fbytes = ((size_t)nf * 12 + 255) & ~(size_t)255;
    const size_t cbytes = (nb * 4 + 255) & ~(size_t)255;
    size_t need = vbytes + fbytes + 2 * cbytes + 256;
    VF_TRY(vf_scratch_reserve(c, c->mesh, need + ((size_t)nf * 16 * 4)));
    char* base = (char*)c->mesh.ptr;
    float* d_verts = (float*)base;
    uint32_t* d_faces = (uint32_t*)(base + vbytes);
    uint32_t* d_counts = (uint32_t*)(base + vbytes + fbytes);
    uint32_t* d_offsets = (uint32_t*)(base + vbytes + fbytes + cbytes);
    uint32_t* d_total = (uint32_t*)(base + vbytes + fbytes + 2 * cbytes);
    VF_CUDA(cudaMemcpyAsync(d_verts, verts, (size_t)nv * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemcpyAsync(d_faces, faces, (size_t)nf * 12, cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemsetAsync(d_counts, 0, cbytes, c->stream));
    const int tb = (int)((nf + 255) / 256);
    bin_triangles_kernel<false><<<tb, 256, 0, c->stream>>>(d_verts, d_faces, nf, g, d_counts, nullptr, nullptr);
    VF_LAUNCHED(c);
    scan_counts_kernel<<<1, 1024, 0, c->stream>>>(d_counts, d_offsets, (uint32_t)nb, d_total);
    VF_LAUNCHED(c);
    uint32_t* h_total = (uint32_t*)((char*)c->pinned + 65536);
    VF_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(cudaStreamSynchronize(c->stream));  // also covers the pageable verts/faces uploads
    const size_t total = *h_total;
    VF_REQUIRE(total < (1ull << 31), VF_ERR_CAPACITY, "voxelize: triangle/brick list too long (%zu)", total);
    if (need + total * 4 > c->mesh.bytes) {
        // grow the arena and redo the uploads + count/scan (rare: only when the mesh has very large triangles)
        VfScratch old = c->mesh;
        c->mesh = VfScratch();
        VF_TRY(vf_scratch_reserve(c, c->mesh, need + total * 4 + 256));
        VF_CUDA(cudaMemcpyAsync(c->mesh.ptr, old.ptr, need, cudaMemcpyDeviceToDevice, c->stream));
        VF_CUDA(cudaStreamSynchronize(c->stream));
        VF_CUDA(cudaFree(old.ptr));
        base = (char*)c->mesh.ptr;
        d_verts = (float*)base;
        d_faces = (uint32_t*)(base + vbytes);
        d_counts = (uint32_t*)(base + vbytes + fbytes);
        d_offsets = (uint32_t*)(base + vbytes + fbytes + cbytes);
    }
    uint32_t* d_list = (uint32_t*)(base + need);
    bin_triangles_kernel<true><<<tb, 256, 0, c->stream>>>(d_verts, d_faces, nf, g, d_counts, d_offsets, d_list);
    VF_LAUNCHED(c);
    voxelize_brick_kernel<<<(unsigned)nb, 256, 0, c->stream>>>(grid->d, d_verts, d_faces, g, d_counts, d_offsets, d_list);
    VF_LAUNCHED(c);
    return VF_OK;
}
