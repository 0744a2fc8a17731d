// mesh.cu — f2: per-fragment marching cubes on the label grid, vertex fusion and two-pass Laplacian smoothing.
//
// Replaces the mesh side of RegularGrid::toTriangleMesh (SRC/DataStructures/RegularGrid.cpp:473-486): MarchingCubes::setGrid
// (SRC/Graphics/Core/MarchingCubes.cpp:523-540) + triangulateFieldGPU (:364-432) and its shaders marchingCubes-comp.glsl:96-168,
// computeMortonCodes-comp.glsl, the 30-pass one-bit radix sort (sortMortonCodes, MarchingCubes.cpp:542-609), findSameVertices_01/02,
// buildMarchingCubesFaces, markBoundaryTriangles, resetLaplacianBuffer / laplacianSmoothing / finishLaplacianSmoothing-comp.glsl, for one
// block over the whole grid (the only way the reference runs it: RegularGrid.cpp:423 constructs MarchingCubes with subdivisions = 1, and the
// parameter FractureParameters::_marchingCubesSubdivisions is read nowhere).
//
// What the reference computes: the grid is padded by one cell of VOXEL_FREE; the field is 1 where (label without bit 15) == target and
// 0 elsewhere, isolevel 0.5, so every surface vertex is the midpoint of a cell edge (half-integer coordinates, exact in float32); a
// triangle carries the boundary flag (bit 15) of its cell; vertices are sorted by a 30-bit Morton code, equal neighbours in the sorted
// order are fused, moved by the grid's model matrix, and smoothed: first the vertices of non-boundary faces (weight 0.9), then the
// boundary vertices (weight 0.2), each pass unsigned(max dim * 0.048) iterations of a Laplacian accumulated in int32 at 1e-4 units.
// The reference numbers vertices and faces with atomicAdd, so their ORDER is a race; the deterministic order used here (and in the
// CPU checker) is "threads run in index order": triangles by (cell index, position in the case row), vertices by (Morton code, x, y, z).
// The case table is the classic public-domain one (mc_tritable.inc, see tools/make_mc_table.py).
//
// B200 design: no padded copy of the grid (the ring is synthesised by the corner fetch); case counts per cell in one byte, block sums
// scanned once, so extraction is two streaming passes that write only the triangle soup; a vertex is one 64-bit key (Morton code |
// doubled coordinates) so that ONE radix sort (cub::DeviceRadixSort, library code like the reference's own sort shaders) replaces the
// reference's 30 passes x 5 dispatches; fusion is a flag + scan; the Laplacian uses the same int32 atomics as the shader, which makes
// the smoothing deterministic.  The mesh stays on the device until vf_mesh_download.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>

#include "vf_internal.h"

struct vf_mesh {
    vf_ctx* ctx = nullptr;
    float4* verts = nullptr;   // [nv] xyz + boundary flag
    uint4* faces = nullptr;    // [nf] three vertex numbers + boundary flag
    uint32_t nv = 0, nf = 0;
};

namespace {

__constant__ unsigned long long c_mc_rows[256];
const unsigned long long h_mc_rows[256] = {
#include "mc_tritable.inc"
};

struct McGeom {
    int X, Y, Z;     // grid
    int PX, PY, PZ;  // padded grid
    float scale[3], shift[3];
};

// corner i of cell (x, y, z) in the padded grid (marchingCubes-comp.glsl:28-38): (0,0,0) (0,0,1) (-1,0,1) (-1,0,0) (0,1,0) (0,1,1) (-1,1,1) (-1,1,0)
__device__ __forceinline__ void mc_corner(int i, int& dx, int& dy, int& dz)
{
    dx = -((i & 3) >> 1), dy = i >> 2, dz = ((i & 3) == 1 || (i & 3) == 2) ? 1 : 0;
}
// padded-grid fetch: the ring holds VOXEL_FREE (MarchingCubes.cpp:526-531)
__device__ __forceinline__ uint32_t mc_fetch(const uint16_t* __restrict__ grid, const McGeom& g, int x, int y, int z)
{
    const int gx = x - 1, gy = y - 1, gz = z - 1;
    if ((unsigned)gx >= (unsigned)g.X || (unsigned)gy >= (unsigned)g.Y || (unsigned)gz >= (unsigned)g.Z) return VF_VOXEL_FREE;
    return grid[((size_t)gx * g.Y + gy) * g.Z + gz];
}
__device__ __forceinline__ int mc_configuration(const uint16_t* __restrict__ grid, const McGeom& g, int x, int y, int z, uint32_t target)
{
    if (x == 0 || y == g.PY - 1 || z == g.PZ - 1) return 0;  // :102-103; case 0 has no triangles either
    int configuration = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int dx, dy, dz;
        mc_corner(i, dx, dy, dz);
        if ((mc_fetch(grid, g, x + dx, y + dy, z + dz) & 0x7FFFu) != target) configuration |= 1 << i;  // value < isolevel
    }
    return configuration;
}
__device__ __forceinline__ int mc_row_triangles(unsigned long long row)
{
    // nibbles are packed front to back and triangles take three: count the nibbles before the first 0xF
    const unsigned long long ends = row & (row >> 1) & (row >> 2) & (row >> 3) & 0x1111111111111111ull;
    return (__ffsll((long long)ends) - 1) / 12;  // every row ends with 0xF at nibble 15 at the latest
}
__device__ __forceinline__ uint32_t mc_expand_bits(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

constexpr int kMcBlock = 256;

__device__ __forceinline__ void mc_cell_of(const McGeom& g, size_t cell, int& x, int& y, int& z)
{
    z = (int)(cell % g.PZ);
    const size_t r = cell / g.PZ;
    y = (int)(r % g.PY), x = (int)(r / g.PY);
}

// pass 1: triangles per cell (one byte) and per block of 256 cells
__global__ void __launch_bounds__(kMcBlock) mc_count_kernel(const uint16_t* __restrict__ grid, McGeom g, uint32_t target, size_t ncells, uint8_t* __restrict__ tri_count,
                                                             uint32_t* __restrict__ block_sums)
{
    __shared__ uint32_t warp_sums[kMcBlock / 32];
    const size_t cell = (size_t)blockIdx.x * kMcBlock + threadIdx.x;
    uint32_t n = 0;
    if (cell < ncells) {
        int x, y, z;
        mc_cell_of(g, cell, x, y, z);
        n = (uint32_t)mc_row_triangles(c_mc_rows[mc_configuration(grid, g, x, y, z, target)]);
        tri_count[cell] = (uint8_t)n;
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kMcBlock / 32; ++w) s += warp_sums[w];
        block_sums[blockIdx.x] = s;
    }
}

// pass 2: every cell writes its triangles at its rank: one 64-bit key per vertex = Morton code (computeMortonCodes-comp.glsl) << 33 |
// doubled padded-grid coordinates (11 bits each), and the cell's boundary flag
__global__ void __launch_bounds__(kMcBlock) mc_emit_kernel(const uint16_t* __restrict__ grid, McGeom g, uint32_t target, size_t ncells, const uint8_t* __restrict__ tri_count,
                                                            const uint32_t* __restrict__ block_offsets, unsigned long long* __restrict__ keys, uint8_t* __restrict__ wflag)
{
    __shared__ uint32_t warp_sums[kMcBlock / 32];
    const size_t cell = (size_t)blockIdx.x * kMcBlock + threadIdx.x;
    const uint32_t mine = cell < ncells ? tri_count[cell] : 0;
    uint32_t s = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (!mine) return;
    uint32_t tri = block_offsets[blockIdx.x] + s - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) tri += warp_sums[w];
    int x, y, z;
    mc_cell_of(g, cell, x, y, z);
    const unsigned long long row = c_mc_rows[mc_configuration(grid, g, x, y, z, target)];
    const uint8_t wf = (mc_fetch(grid, g, x, y, z) >> 15) != 0 ? 1 : 0;  // :147 isBoundary(grid[cell])
    const float pd[3] = { (float)g.PX, (float)g.PY, (float)g.PZ };
    for (uint32_t t = 0; t < mine; ++t, ++tri) {
        const int e[3] = { (int)((row >> (12 * t)) & 0xF), (int)((row >> (12 * t + 8)) & 0xF), (int)((row >> (12 * t + 4)) & 0xF) };  // (0, 2, 1), :141-143
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // edge e joins corners a and b (marchingCubes-comp.glsl:40-54); with values in {0, 1} and isolevel 0.5 the vertex is the midpoint
            const int a = e[k] < 8 ? e[k] : e[k] - 8, b = e[k] < 8 ? ((e[k] & 4) | ((e[k] + 1) & 3)) : e[k] - 4;
            int ax, ay, az, bx, by, bz;
            mc_corner(a, ax, ay, az), mc_corner(b, bx, by, bz);
            const int c2[3] = { 2 * x + ax + bx, 2 * y + ay + by, 2 * z + az + bz };  // doubled coordinates
            const float p[3] = { 0.5f * (float)c2[0], 0.5f * (float)c2[1], 0.5f * (float)c2[2] };
            const uint32_t morton = mc_expand_bits((uint32_t)(__fdiv_rn(p[0], pd[0]) * 1024.0f)) * 4 + mc_expand_bits((uint32_t)(__fdiv_rn(p[1], pd[1]) * 1024.0f)) * 2 +
                                    mc_expand_bits((uint32_t)(__fdiv_rn(p[2], pd[2]) * 1024.0f));
            keys[3 * (size_t)tri + k] = (unsigned long long)morton << 33 | (unsigned long long)c2[0] << 22 | (unsigned long long)c2[1] << 11 | (unsigned long long)c2[2];
            wflag[3 * (size_t)tri + k] = wf;
        }
    }
}

__global__ void __launch_bounds__(256) mc_iota_kernel(uint32_t* __restrict__ v, uint32_t n)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// findSameVertices_01: a sorted vertex that differs from its predecessor opens a new fused vertex
__global__ void __launch_bounds__(256) mc_flag_kernel(const unsigned long long* __restrict__ sorted_keys, uint32_t n, uint32_t* __restrict__ flag)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flag[k] = (k == 0 || sorted_keys[k] != sorted_keys[k - 1]) ? 1u : 0u;
}

// findSameVertices_01/02 + the model matrix: fused numbers back to the soup; the first copy of a group gives position and flag
__global__ void __launch_bounds__(256) mc_fuse_kernel(const unsigned long long* __restrict__ sorted_keys, const uint32_t* __restrict__ sorted_soup,
                                                      const uint32_t* __restrict__ rank_incl, const uint8_t* __restrict__ wflag, uint32_t n, McGeom g,
                                                      uint32_t* __restrict__ fused_of_soup, float4* __restrict__ verts)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = rank_incl[k] - 1, soup = sorted_soup[k];
    fused_of_soup[soup] = id;
    if (k == 0 || sorted_keys[k] != sorted_keys[k - 1]) {
        const unsigned long long key = sorted_keys[k];
        const float p[3] = { 0.5f * (float)((key >> 22) & 0x7FF), 0.5f * (float)((key >> 11) & 0x7FF), 0.5f * (float)(key & 0x7FF) };
        verts[id] = make_float4(__fadd_rn(__fmul_rn(p[0], g.scale[0]), g.shift[0]), __fadd_rn(__fmul_rn(p[1], g.scale[1]), g.shift[1]),
                                __fadd_rn(__fmul_rn(p[2], g.scale[2]), g.shift[2]), (float)wflag[soup]);
    }
}

// buildMarchingCubesFaces + markBoundaryTriangles
__global__ void __launch_bounds__(256) mc_faces_kernel(const uint32_t* __restrict__ fused_of_soup, const float4* __restrict__ verts, uint32_t nf, uint4* __restrict__ faces)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const uint32_t a = fused_of_soup[3 * (size_t)f], b = fused_of_soup[3 * (size_t)f + 1], c = fused_of_soup[3 * (size_t)f + 2];
    faces[f] = make_uint4(a, b, c, (uint32_t)fmaxf(verts[a].w, fmaxf(verts[b].w, verts[c].w)));
}

// laplacianSmoothing-comp.glsl: every face adds each of its vertices (at 1e-4 units, truncated) to the other two
__global__ void __launch_bounds__(256) mc_laplacian_kernel(const float4* __restrict__ verts, const uint4* __restrict__ faces, uint32_t nf, int check_validity, float target,
                                                           int4* __restrict__ lap)
{
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const uint4 fc = faces[f];
    const uint32_t id[3] = { fc.x, fc.y, fc.z };
    float4 v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) v[i] = verts[id[i]];
    if (check_validity)
        for (int i = 0; i < 3; ++i)
            if (!(fabsf(v[i].w - target) < 0.00000001f)) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int qx = (int)__fmul_rn(v[i].x, 10000.0f), qy = (int)__fmul_rn(v[i].y, 10000.0f), qz = (int)__fmul_rn(v[i].z, 10000.0f);
#pragma unroll
        for (int n = 1; n <= 2; ++n) {
            int* l = reinterpret_cast<int*>(&lap[id[(i + n) % 3]]);
            atomicAdd(l, qx), atomicAdd(l + 1, qy), atomicAdd(l + 2, qz), atomicAdd(l + 3, 1);
        }
    }
}

// finishLaplacianSmoothing-comp.glsl, and the reset of the accumulators for the next iteration
__global__ void __launch_bounds__(256) mc_finish_kernel(float4* __restrict__ verts, int4* __restrict__ lap, uint32_t nv, float target, float weight)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int4 l = lap[i];
    lap[i] = make_int4(0, 0, 0, 0);
    float4 v = verts[i];
    if (fabsf(v.w - target) < 0.00000001f && l.w > 0) {
        const float d = (float)l.w, keep = __fsub_rn(1.0f, weight);
        const float ax = __fdiv_rn(__fdiv_rn((float)l.x, d), 10000.0f), ay = __fdiv_rn(__fdiv_rn((float)l.y, d), 10000.0f), az = __fdiv_rn(__fdiv_rn((float)l.z, d), 10000.0f);
        v.x = __fadd_rn(__fmul_rn(v.x, keep), __fmul_rn(ax, weight));
        v.y = __fadd_rn(__fmul_rn(v.y, keep), __fmul_rn(ay, weight));
        v.z = __fadd_rn(__fmul_rn(v.z, keep), __fmul_rn(az, weight));
        verts[i] = v;
    }
}

// exclusive scan of the block sums by one CTA (a few hundred thousand entries at most)
__global__ void __launch_bounds__(1024) mc_scan_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? in[i] : 0;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + s - v;
        if (i < n) out[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

bool g_rows_uploaded[64] = {};

}  // namespace

extern "C" void vf_mc_params_default(vf_mc_params* p)
{
    if (!p) return;
    p->boundaryMCIterations = 0.048f;     // FractureParameters.h:93
    p->boundaryMCWeight = 0.2f;           // :94
    p->nonBoundaryMCIterations = 0.048f;  // :109
    p->nonBoundaryMCWeight = 0.9f;        // :110
    p->marchingCubesSubdivisions = 1;     // :105
}

extern "C" vf_status vf_marching_cubes(vf_grid* grid, uint32_t target_value, const vf_mc_params* params, vf_mesh** out)
{
    VF_REQUIRE(grid && out, VF_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    vf_ctx* c = grid->ctx;
    VF_TRY(vf_enter(c));
    vf_mc_params mp;
    vf_mc_params_default(&mp);
    if (params) mp = *params;
    // mp.marchingCubesSubdivisions is carried for the parameter surface only, exactly as in the reference: FractureParameters.h:56 declares it and
    // nothing reads it — RegularGrid::resetMarchingCubes builds MarchingCubes(*this, 1, _numDivs, 5) (RegularGrid.cpp:423), one block whatever it says.
    VF_REQUIRE(grid->X + 2 <= 1023 && grid->Y + 2 <= 1023 && grid->Z + 2 <= 1023, VF_ERR_CAPACITY, "marching cubes keys hold 11 bits per doubled coordinate (<= 1021 cells per axis)");
    VF_REQUIRE(target_value > VF_VOXEL_FREE && target_value < 0x8000u, VF_ERR_INVALID_ARGUMENT, "target value %u is not a fragment label", target_value);
    if (c->device < 64 && !g_rows_uploaded[c->device]) {
        VF_CUDA(cudaMemcpyToSymbolAsync(c_mc_rows, h_mc_rows, sizeof(h_mc_rows), 0, cudaMemcpyHostToDevice, c->stream));
        g_rows_uploaded[c->device] = true;
    }
    McGeom g;
    g.X = (int)grid->X, g.Y = (int)grid->Y, g.Z = (int)grid->Z;
    g.PX = g.X + 2, g.PY = g.Y + 2, g.PZ = g.Z + 2;
    const uint32_t dims[3] = { grid->X, grid->Y, grid->Z };
    for (int q = 0; q < 3; ++q) {  // RegularGrid.cpp:478-480: translate(-scale) * translate(min) * scale(scale)
        g.scale[q] = (grid->aabb_max[q] - grid->aabb_min[q]) / (float)dims[q];
        g.shift[q] = grid->aabb_min[q] + (-g.scale[q]);
    }
    const size_t ncells = (size_t)g.PX * g.PY * g.PZ;
    const uint32_t nblocks = (uint32_t)((ncells + kMcBlock - 1) / kMcBlock);
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // arena (the flood key scratch is free between fragmentations): tri_count | block_sums | block_offsets | total
    const size_t cb = up(ncells), bb = up((size_t)nblocks * 4);
    VF_TRY(vf_scratch_reserve(c, c->keys, cb + 2 * bb + 256));
    uint8_t* d_count = (uint8_t*)c->keys.ptr;
    uint32_t* d_bsum = (uint32_t*)((char*)c->keys.ptr + cb);
    uint32_t* d_boff = (uint32_t*)((char*)c->keys.ptr + cb + bb);
    uint32_t* d_total = (uint32_t*)((char*)c->keys.ptr + cb + 2 * bb);
    mc_count_kernel<<<nblocks, kMcBlock, 0, c->stream>>>(grid->d, g, target_value, ncells, d_count, d_bsum);
    VF_LAUNCHED(c);
    mc_scan_kernel<<<1, 1024, 0, c->stream>>>(d_bsum, d_boff, nblocks, d_total);
    VF_LAUNCHED(c);
    uint32_t* h_total = (uint32_t*)((char*)c->pinned + 65536 + 192);
    VF_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    const uint32_t nf = *h_total;
    VF_REQUIRE((uint64_t)nf * 3 < (1ull << 31), VF_ERR_CAPACITY, "marching cubes: %u triangles exceed the 32-bit soup", nf);
    vf_mesh* m = new (std::nothrow) vf_mesh();
    VF_REQUIRE(m != nullptr, VF_ERR_CAPACITY, "out of host memory");
    m->ctx = c, m->nf = nf;
    *out = m;
    if (nf == 0) return VF_OK;
    const uint32_t ns = 3 * nf;  // soup vertices

    // working set of the sort / fusion in the mesh arena: keys x2 | soup ids x2 | flags/ranks | fused_of_soup | wflag | cub temp
    size_t temp_sort = 0, temp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_sort, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)ns, 0, 63,
                                    c->stream);
    cub::DeviceScan::InclusiveSum(nullptr, temp_scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)ns, c->stream);
    const size_t kb = up((size_t)ns * 8), ib = up((size_t)ns * 4), wb = up(ns), tb = up(std::max(temp_sort, temp_scan));
    VF_TRY(vf_scratch_reserve(c, c->mesh, 2 * kb + 5 * ib + wb + tb + 256));
    char* base = (char*)c->mesh.ptr;
    unsigned long long* d_keys = (unsigned long long*)base;
    unsigned long long* d_keys2 = (unsigned long long*)(base + kb);
    uint32_t* d_soup = (uint32_t*)(base + 2 * kb);
    uint32_t* d_soup2 = (uint32_t*)(base + 2 * kb + ib);
    uint32_t* d_flag = (uint32_t*)(base + 2 * kb + 2 * ib);
    uint32_t* d_rank = (uint32_t*)(base + 2 * kb + 3 * ib);
    uint32_t* d_fused = (uint32_t*)(base + 2 * kb + 4 * ib);
    uint8_t* d_w = (uint8_t*)(base + 2 * kb + 5 * ib);
    void* d_temp = base + 2 * kb + 5 * ib + wb;
    mc_emit_kernel<<<nblocks, kMcBlock, 0, c->stream>>>(grid->d, g, target_value, ncells, d_count, d_boff, d_keys, d_w);
    VF_LAUNCHED(c);
    mc_iota_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(d_soup, ns);
    VF_LAUNCHED(c);
    size_t tbytes = tb;
    VF_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, tbytes, d_keys, d_keys2, d_soup, d_soup2, (int)ns, 0, 63, c->stream));  // stable: equal keys keep soup order
    ++c->launches;
    mc_flag_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(d_keys2, ns, d_flag);
    VF_LAUNCHED(c);
    tbytes = tb;
    VF_CUDA(cub::DeviceScan::InclusiveSum(d_temp, tbytes, d_flag, d_rank, (int)ns, c->stream));
    ++c->launches;
    VF_CUDA(cudaMemcpyAsync(h_total, d_rank + (ns - 1), 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    const uint32_t nv = *h_total;
    m->nv = nv;
    VF_CUDA(cudaMalloc(&m->verts, (size_t)nv * sizeof(float4)));
    VF_CUDA(cudaMalloc(&m->faces, (size_t)nf * sizeof(uint4)));
    mc_fuse_kernel<<<(ns + 255) / 256, 256, 0, c->stream>>>(d_keys2, d_soup2, d_rank, d_w, ns, g, d_fused, m->verts);
    VF_LAUNCHED(c);
    mc_faces_kernel<<<(nf + 255) / 256, 256, 0, c->stream>>>(d_fused, m->verts, nf, m->faces);
    VF_LAUNCHED(c);

    // smoothSurface (MarchingCubes.cpp:498-521): iterations = unsigned(maxVoxels * _{non}boundaryMCIterations) (:399-400)
    const uint32_t max_voxels = std::max(grid->X, std::max(grid->Y, grid->Z));
    const uint32_t nb_iters = (uint32_t)((float)max_voxels * mp.nonBoundaryMCIterations), b_iters = (uint32_t)((float)max_voxels * mp.boundaryMCIterations);
    if (nb_iters + b_iters) {
        int4* d_lap = (int4*)base;  // the sort buffers are free now
        VF_REQUIRE((size_t)nv * sizeof(int4) <= 2 * kb, VF_ERR_CAPACITY, "laplacian accumulators do not fit");
        VF_TRY(vf_k_zero(c, d_lap, (size_t)nv * sizeof(int4)));
        for (int pass = 0; pass < 2; ++pass) {
            const bool boundary = pass == 1;
            const uint32_t iters = boundary ? b_iters : nb_iters;
            const float weight = boundary ? mp.boundaryMCWeight : mp.nonBoundaryMCWeight, target = boundary ? 1.0f : 0.0f;
            for (uint32_t it = 0; it < iters; ++it) {
                mc_laplacian_kernel<<<(nf + 255) / 256, 256, 0, c->stream>>>(m->verts, m->faces, nf, boundary ? 0 : 1, target, d_lap);
                VF_LAUNCHED(c);
                mc_finish_kernel<<<(nv + 255) / 256, 256, 0, c->stream>>>(m->verts, d_lap, nv, target, weight);
                VF_LAUNCHED(c);
            }
        }
    }
    return VF_OK;
}

extern "C" vf_status vf_mesh_counts(const vf_mesh* m, uint32_t* nv, uint32_t* nf)
{
    VF_REQUIRE(m != nullptr, VF_ERR_INVALID_ARGUMENT, "null mesh");
    if (nv) *nv = m->nv;
    if (nf) *nf = m->nf;
    return VF_OK;
}

extern "C" vf_status vf_mesh_download(vf_mesh* m, float* verts, uint32_t* faces)
{
    VF_REQUIRE(m != nullptr, VF_ERR_INVALID_ARGUMENT, "null mesh");
    vf_ctx* c = m->ctx;
    VF_TRY(vf_enter(c));
    if (verts && m->nv) VF_CUDA(cudaMemcpyAsync(verts, m->verts, (size_t)m->nv * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (faces && m->nf) VF_CUDA(cudaMemcpyAsync(faces, m->faces, (size_t)m->nf * sizeof(uint4), cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    return VF_OK;
}

extern "C" void vf_mesh_destroy(vf_mesh* m)
{
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    if (m->verts) cudaFree(m->verts);
    if (m->faces) cudaFree(m->faces);
    delete m;
}
