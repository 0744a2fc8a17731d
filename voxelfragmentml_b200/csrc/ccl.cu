// ccl.cu — connected-to-seed component selection by run-based union-find (C1, and the disjoint step of F3).
//
// C1 replaces the intended semantics of NaiveFracturer::removeIsolatedRegions (removeIsolatedRegionsCPU,
// SRC/Fracturer/NaiveFracturer.cpp:111-150: keep only cells 6-connected to their own seed through same-label cells, rebuild the
// grid from an all-EMPTY one).  F3 is the net effect of floodFracturer-comp.glsl:49-63 + disjointSet-comp.glsl:17-24 +
// disjointSetStack-comp.glsl:20-37: per fragment id keep the component (under the flood neighbourhood, through equal-fragId
// cells) that holds the fragment's lowest-prefix source, return the rest to FREE.
//
// Connectivity has no notion of distance, so it does not need the wavefront iterations of the flood.  A lock-free union-find
// (parent links always point to a lower index, roots are linked with atomicMin) labels every component:
//   * the unit of work is a 32-cell z-segment = one row of a 16 x 16 x 32 tile.  A warp ballot over "same region as my
//     z-predecessor" turns a segment into two 32-bit masks (active cells, run starts); runs, not cells, are the union-find nodes;
//   * two rows have to be united only at "key positions": cells where one of the two rows starts a run while both are active —
//     if both rows merely continue their runs the pair one cell earlier already did the job.  Key positions are bit tricks on
//     the masks, so a thread serves a whole row against a neighbouring row in a handful of instructions;
//   * stage 1 resolves the components inside every tile in shared memory (thread per row), writes each cell's tile-local root
//     to the global parent array (depth-1 forest) and the segment masks next to it; stage 2 (thread per segment) unites only
//     pairs that straddle a tile border; stage 3 marks the roots of the start cells and selects.
// The parent array reuses the 4 B/voxel flood key scratch, the masks (N/4 bytes) the second-grid scratch.
// Requires N < 2^31 (bit 31 of a root's parent word carries "keep").
#include <algorithm>

#include "vf_internal.h"

namespace vfccl {

constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t KEEP = 0x80000000u;
constexpr unsigned kFull = 0xFFFFFFFFu;
enum { MODE_C1 = 0, MODE_F3 = 1 };
constexpr int LX = 16, LY = 16, LZ = 32, LROWS = LX * LY, LCELLS = LROWS * LZ;

template <int MODE>
__device__ __forceinline__ bool active(uint32_t v)
{
    return (v & 0x7FFFu) > VF_VOXEL_FREE;
}
template <int MODE>
__device__ __forceinline__ bool same(uint32_t a, uint32_t b)
{
    return MODE == MODE_C1 ? a == b : ((a ^ b) & 0xFFu) == 0;
}

// position of the run start that covers bit z (S has bit 0 set)
__device__ __forceinline__ int run_start(unsigned S, int z) { return 31 - __clz(S & (0xFFFFFFFFu >> (31 - z))); }

struct Geo {
    int X, Y, Z, segs;  // segs = 32-cell segments per z-row
    int ntx, nty, ntz;
    uint32_t nsegs;
};

// ------------------------------------------------------------------------------------------------ shared-memory union-find
__device__ __forceinline__ uint32_t find_local(volatile uint32_t* par, uint32_t i)
{
    uint32_t p;
    while ((p = par[i]) != i) i = p;
    return i;
}
__device__ __forceinline__ void unite_local(uint32_t* par, uint32_t a, uint32_t b)
{
    while (true) {
        a = find_local(par, a);
        b = find_local(par, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(&par[a], b);
        if (old == a) return;
        a = old;
    }
}

// ------------------------------------------------------------------------------------------------ global union-find
__device__ __forceinline__ uint32_t find_root(uint32_t* __restrict__ P, uint32_t i)
{
    // path halving; every store writes an ancestor, so concurrent finds/unions stay consistent
    uint32_t p = __ldcg(&P[i]);
    while (p != i) {
        const uint32_t gp = __ldcg(&P[p]);
        if (gp != p) P[i] = gp;
        i = p;
        p = gp;
    }
    return i;
}
__device__ __forceinline__ void unite(uint32_t* __restrict__ P, uint32_t a, uint32_t b)
{
    while (true) {
        a = find_root(P, a);
        b = find_root(P, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        const uint32_t old = atomicMin(&P[a], b);  // link the higher root under the lower one
        if (old == a) return;
        a = old;  // somebody linked `a` meanwhile: continue from where it points now
    }
}

// C1 only: newGrid[seed] = seed.w whatever the cell held (NaiveFracturer.cpp:120-123); a later seed on the same cell wins
__global__ void ccl_plant_seeds_kernel(uint16_t* __restrict__ grid, Geo g, const ushort4* __restrict__ seeds, int S)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const ushort4 sd = seeds[s];
    for (int t = s + 1; t < S; ++t)
        if (seeds[t].x == sd.x && seeds[t].y == sd.y && seeds[t].z == sd.z) return;
    grid[((size_t)sd.x * g.Y + sd.y) * g.Z + sd.z] = sd.w;
}

// ------------------------------------------------------------------------------------------------ stage 1: inside a tile
template <int MODE, int NNEIGH>
__global__ void __launch_bounds__(256) ccl_tile_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ P, uint2* __restrict__ masks, Geo g)
{
    extern __shared__ uint32_t smem_ccl[];
    uint32_t* par = smem_ccl;                                       // [LCELLS] only run-start entries are nodes
    uint32_t* sA = par + LCELLS;                                    // [LROWS] active mask per row
    uint32_t* sS = sA + LROWS;                                      // [LROWS] run-start mask per row (bit 0 always set)
    uint16_t* lab = reinterpret_cast<uint16_t*>(sS + LROWS);        // [LCELLS] local index = row * 32 + z, row = x * LY + y
    const int tile = blockIdx.x;
    const int tz = tile % g.ntz, ty = (tile / g.ntz) % g.nty, tx = tile / (g.ntz * g.nty);
    const int gx0 = tx * LX, gy0 = ty * LY, gz0 = tz * LZ;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // (a1) stage the tile's labels: 16-byte asynchronous copies when rows are 16-byte aligned (Z % 8 == 0), so that all of a
    //      thread's loads are in flight at once; scalar loads otherwise.  Cells outside the grid read as EMPTY.
    const bool vec_ok = (g.Z % 8 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    if (vec_ok) {
        for (int q = threadIdx.x; q < LROWS * 4; q += 256) {  // 4 chunks of 8 cells per row
            const int r = q >> 2, ch = q & 3;
            const int gx = gx0 + r / LY, gy = gy0 + r % LY, gz = gz0 + ch * 8;
            uint16_t* dst = &lab[r * LZ + ch * 8];
            if (gx < g.X && gy < g.Y && gz < g.Z) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(grid + ((size_t)gx * g.Y + gy) * g.Z + gz) : "memory");
            } else {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int r = warp; r < LROWS; r += 8) {
            const int gx = gx0 + r / LY, gy = gy0 + r % LY, gz = gz0 + lane;
            lab[r * LZ + lane] = (gx < g.X && gy < g.Y && gz < g.Z) ? grid[((size_t)gx * g.Y + gy) * g.Z + gz] : (uint16_t)0;
        }
    }
    __syncthreads();
    // (a2) warp per row, lane = z: masks, run-start parents
    for (int r = warp; r < LROWS; r += 8) {
        const int gx = gx0 + r / LY, gy = gy0 + r % LY;
        const uint32_t v = lab[r * LZ + lane];
        const bool act = active<MODE>(v);
        const uint32_t vp = __shfl_up_sync(kFull, v, 1);
        const bool cont = lane > 0 && act && active<MODE>(vp) && same<MODE>(v, vp);
        const unsigned A = __ballot_sync(kFull, act), S = ~__ballot_sync(kFull, cont);
        par[r * LZ + lane] = r * LZ + lane;
        if (lane == 0) {
            sA[r] = A;
            sS[r] = S;
            if (gx < g.X && gy < g.Y && gz0 < g.Z) masks[((size_t)gx * g.Y + gy) * g.segs + tz] = make_uint2(A, S);
        }
    }
    __syncthreads();

    // (b) thread per row: unite my runs with the runs of the backward neighbour rows at the key positions
    {
        const int r = threadIdx.x, x = r / LY, y = r % LY;
        const unsigned Am = sA[r], Sm = sS[r];
        auto against = [&](int nx, int ny, int dz) {
            if (nx < 0 || ny < 0 || ny >= LY) return;  // other tiles: stage 2
            const int rn = nx * LY + ny;
            unsigned An = sA[rn], Sn = sS[rn];
            if (dz < 0) An <<= 1, Sn <<= 1;   // position z of the shifted row is neighbour cell z-1
            if (dz > 0) An >>= 1, Sn >>= 1;   // ... neighbour cell z+1
            unsigned m = (Sm | Sn) & Am & An;
            while (m) {
                const int z = __ffs(m) - 1;
                m &= m - 1;
                const int zn = z + dz;
                if (same<MODE>(lab[r * LZ + z], lab[rn * LZ + zn]))
                    unite_local(par, r * LZ + run_start(Sm, z), rn * LZ + run_start(sS[rn], zn));
            }
        };
        if (Am) {
            against(x, y - 1, 0), against(x - 1, y, 0);
            if (NNEIGH == 26) {
                against(x, y - 1, -1), against(x, y - 1, 1);
                against(x - 1, y, -1), against(x - 1, y, 1);
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz) against(x - 1, y - 1, dz), against(x - 1, y + 1, dz);
            }
        }
    }
    __syncthreads();

    // (b2) thread per row: flatten the run starts of my row, so that (c) is one shared-memory read per cell
    {
        const int r = threadIdx.x;
        unsigned st = sS[r] & sA[r];
        while (st) {
            const int z = __ffs(st) - 1;
            st &= st - 1;
            par[r * LZ + z] = find_local(par, r * LZ + z);  // writes an ancestor: safe against concurrent finds
        }
    }
    __syncthreads();

    // (c) warp per row: every active cell points at the global index of its tile-local root
    for (int r = warp; r < LROWS; r += 8) {
        const int gx = gx0 + r / LY, gy = gy0 + r % LY, gz = gz0 + lane;
        if (!(gx < g.X && gy < g.Y && gz < g.Z)) continue;
        uint32_t out = NONE;
        if (sA[r] >> lane & 1u) {
            const uint32_t root = par[r * LZ + run_start(sS[r], lane)];
            const int rz = root % LZ, rr = root / LZ;
            out = ((uint32_t)(gx0 + rr / LY) * g.Y + gy0 + rr % LY) * g.Z + gz0 + rz;
        }
        P[((size_t)gx * g.Y + gy) * g.Z + gz] = out;
    }
}

// ------------------------------------------------------------------------------------------------ stage 2: across tile borders
// thread per 32-cell segment.  A pair (my cell z, neighbour cell z+dz of row (nx,ny)) is this stage's business iff the two cells
// lie in different tiles.
template <int MODE, int NNEIGH>
__global__ void __launch_bounds__(256) ccl_border_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ P, const uint2* __restrict__ masks, Geo g)
{
    for (uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x; sg < g.nsegs; sg += gridDim.x * blockDim.x) {
        const int seg = sg % g.segs;
        const uint32_t row = sg / g.segs;
        const int y = (int)(row % g.Y), x = (int)(row / g.Y);
        const uint2 mm = masks[sg];
        const unsigned Am = mm.x, Sm = mm.y;
        if (!Am) continue;
        const uint32_t base = row * (uint32_t)g.Z + seg * 32;
        // A pair (i, n) that straddles a tile border need not be united when a "witness" pair one row (or plane) back, inside the
        // same two tiles, carries the same labels: stage 1 united i with its witness and n with its witness, and the witness pair
        // is handled by its own thread (induction on (x, y)).  Only rows on a tile edge, or rows where the labels change, pay.
        const int YZ = g.Y * g.Z;
        auto witnessed = [&](uint32_t i, uint32_t n, uint32_t vi, uint32_t vn, int back) {
            const uint32_t wi = grid[i - back], wn = grid[n - back];
            return active<MODE>(wi) && active<MODE>(wn) && same<MODE>(vi, wi) && same<MODE>(vn, wn);
        };
        // same row, previous segment (always another tile because LZ == 32)
        if (seg > 0 && (Am & 1u)) {
            const uint2 pm = masks[sg - 1];
            if (pm.x >> 31) {
                const uint32_t vi = grid[base], vn = grid[base - 1];
                if (same<MODE>(vi, vn)) {
                    const bool skip = (y % LY != 0 && witnessed(base, base - 1, vi, vn, g.Z)) || (x % LX != 0 && witnessed(base, base - 1, vi, vn, YZ));
                    if (!skip) unite(P, base, base - 1);
                }
            }
        }
        auto against = [&](int nx, int ny, int dz) {
            if (nx < 0 || ny < 0 || ny >= g.Y) return;
            const bool other_tile = (nx / LX != x / LX) || (ny / LY != y / LY);
            if (!other_tile && dz == 0) return;
            const uint32_t nrow = (uint32_t)nx * g.Y + ny;
            const uint32_t nsg = nrow * g.segs + seg;
            const uint2 cm = masks[nsg];
            unsigned An = cm.x, Sn = cm.y;
            if (dz < 0) {
                const uint2 pm = seg > 0 ? masks[nsg - 1] : make_uint2(0u, 0u);
                An = (An << 1) | (pm.x >> 31), Sn = (Sn << 1) | 1u;  // neighbour cell -1 sits in the previous segment: treat as a start
            } else if (dz > 0) {
                const uint2 qm = seg + 1 < g.segs ? masks[nsg + 1] : make_uint2(0u, 0u);
                An = (An >> 1) | (qm.x << 31), Sn = (Sn >> 1) | 0x80000000u;
            }
            unsigned m = (Sm | Sn) & Am & An;
            if (!other_tile) m &= dz < 0 ? 1u : 0x80000000u;  // same tile row: only the cell that leaves the segment crosses a border
            const uint32_t nbase = nrow * (uint32_t)g.Z + seg * 32;
            while (m) {
                const int z = __ffs(m) - 1;
                m &= m - 1;
                const uint32_t i = base + z, n = nbase + z + dz;
                const uint32_t vi = grid[i], vn = grid[n];
                if (!same<MODE>(vi, vn)) continue;
                // witness one plane back (pair crosses a y border only) or one row back (pair crosses an x border only)
                bool skip = false;
                if (NNEIGH == 6) {
                    if (ny != y && x % LX != 0) skip = witnessed(i, n, vi, vn, YZ);
                    else if (nx != x && y % LY != 0) skip = witnessed(i, n, vi, vn, g.Z);
                }
                if (!skip) unite(P, i, n);  // P[cell] is the cell's tile-local root: depth-1 entry points
            }
        };
        against(x, y - 1, 0), against(x - 1, y, 0);
        if (NNEIGH == 26) {
            against(x, y - 1, -1), against(x, y - 1, 1);
            against(x - 1, y, -1), against(x - 1, y, 1);
#pragma unroll
            for (int dz = -1; dz <= 1; ++dz) against(x - 1, y - 1, dz), against(x - 1, y + 1, dz);
        }
    }
}

// mark the root of each start cell's component
__global__ void ccl_mark_kernel(uint32_t* __restrict__ P, Geo g, const ushort4* __restrict__ starts, int S)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const ushort4 sd = starts[s];
    const uint32_t c = ((uint32_t)sd.x * g.Y + sd.y) * g.Z + sd.z;
    uint32_t r = P[c];
    if (r == NONE) return;
    r &= ~KEEP;  // the start cell may itself be a root that another start already marked
    while (true) {
        const uint32_t q = __ldcg(&P[r]) & ~KEEP;
        if (q == r) break;
        r = q;
    }
    atomicOr(&P[r], KEEP);
}

// C1: everything that is not in a kept component becomes EMPTY (the reference rebuilds from an all-EMPTY grid);
// F3: labelled cells outside the kept components return to FREE and are counted (disjointSetStack-comp.glsl:27-31).
// The root lookup walks the (shallow) parent chain; cells of one segment share it, so it is done once per run start.
template <int MODE>
__global__ void __launch_bounds__(256) ccl_select_kernel(uint16_t* __restrict__ grid, const uint32_t* __restrict__ P, const uint2* __restrict__ masks, Geo g,
                                                         uint32_t* __restrict__ freed_out)
{
    // thread per 32-cell segment: resolve each run's root once, build the mask of cells to clear; only segments that lose
    // cells touch the grid (2 B written per removed cell)
    unsigned freed = 0;
    for (uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x; sg < g.nsegs; sg += gridDim.x * blockDim.x) {
        const uint2 mm = masks[sg];
        const int seg = sg % g.segs;
        const uint32_t base = (sg / g.segs) * (uint32_t)g.Z + seg * 32;
        const int ncell = min(32, g.Z - seg * 32);
        const unsigned valid = ncell == 32 ? 0xFFFFFFFFu : ((1u << ncell) - 1);
        unsigned dead = MODE == MODE_C1 ? valid & ~mm.x : 0u;  // C1: inactive non-EMPTY cells (FREE) are dropped as well
        unsigned starts = mm.y & mm.x;
        // run starts among the active cells: a start bit on an inactive cell opens no run
        unsigned todo = mm.x;
        while (todo) {
            const int z = __ffs(todo) - 1;
            // the run containing z: from z up to (excluding) the next start bit or the next inactive cell
            unsigned after = (mm.y | ~mm.x) & ~((2u << z) - 1);
            const int end = after ? __ffs(after) - 1 : 32;
            const unsigned runmask = (end == 32 ? 0xFFFFFFFFu : ((1u << end) - 1)) & ~((1u << z) - 1);
            todo &= ~runmask;
            uint32_t r = P[base + z] & ~KEEP;
            bool keep;
            while (true) {
                const uint32_t q = __ldg(&P[r]);
                if ((q & ~KEEP) == r) {
                    keep = (q & KEEP) != 0;
                    break;
                }
                r = q & ~KEEP;
            }
            if (!keep) dead |= runmask;
        }
        (void)starts;
        while (dead) {
            const int z = __ffs(dead) - 1;
            dead &= dead - 1;
            const uint16_t v = grid[base + z];
            if (MODE == MODE_C1 ? v != VF_VOXEL_EMPTY : v > VF_VOXEL_FREE) {
                grid[base + z] = MODE == MODE_C1 ? VF_VOXEL_EMPTY : VF_VOXEL_FREE;
                ++freed;
            }
        }
    }
    freed = __reduce_add_sync(kFull, freed);
    if ((threadIdx.x & 31) == 0 && freed) atomicAdd(freed_out, freed);
}

constexpr size_t kTileSmem = (size_t)LCELLS * 4 + (size_t)LROWS * 8 + (size_t)LCELLS * 2;

template <int MODE, int NNEIGH>
vf_status run_ccl(vf_grid* grid, const Geo& g, uint32_t* P, uint2* masks, int blocks_lin)
{
    vf_ctx* c = grid->ctx;
    auto tk = ccl_tile_kernel<MODE, NNEIGH>;
    VF_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem));
    tk<<<g.ntx * g.nty * g.ntz, 256, kTileSmem, c->stream>>>(grid->d, P, masks, g);
    VF_LAUNCHED(c);
    ccl_border_kernel<MODE, NNEIGH><<<blocks_lin, 256, 0, c->stream>>>(grid->d, P, masks, g);
    VF_LAUNCHED(c);
    return VF_OK;
}

}  // namespace vfccl

using namespace vfccl;

// Keeps only the components that contain a start cell.  mode 0 = C1 (6-neighbourhood, whole-word equality, the rest -> EMPTY,
// start cells are first overwritten with their seed word); mode 1 = F3 (nneigh 6|26, fragment-id equality, the rest -> FREE).
// *d_freed (device, may be null) accumulates the number of cells removed.
vf_status vf_k_keep_seed_components(vf_grid* grid, const ushort4* d_starts, int nstarts, int mode, int nneigh, uint32_t* d_freed)
{
    vf_ctx* c = grid->ctx;
    const size_t n = grid->n();
    VF_REQUIRE(n < (1ull << 31), VF_ERR_CAPACITY, "connected components: grid has %zu cells (limit 2^31 - 1)", n);
    Geo g;
    g.X = (int)grid->X, g.Y = (int)grid->Y, g.Z = (int)grid->Z;
    g.segs = (g.Z + 31) / 32;
    g.ntx = (g.X + LX - 1) / LX, g.nty = (g.Y + LY - 1) / LY, g.ntz = g.segs;
    g.nsegs = (uint32_t)((size_t)g.X * g.Y * g.segs);
    VF_TRY(vf_scratch_reserve(c, c->keys, n * 4));
    VF_TRY(vf_scratch_reserve(c, c->grid2, std::max(n * 2, (size_t)g.nsegs * 8)));
    uint32_t* P = (uint32_t*)c->keys.ptr;
    uint2* masks = (uint2*)c->grid2.ptr;
    const int blocks_lin = c->num_sms * 8;
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    if (!d_freed) {
        d_freed = (uint32_t*)((char*)c->small.ptr + (700 << 10));
        VF_CUDA(cudaMemsetAsync(d_freed, 0, 4, c->stream));
    }
    if (mode == MODE_C1) {
        ccl_plant_seeds_kernel<<<(nstarts + 127) / 128, 128, 0, c->stream>>>(grid->d, g, d_starts, nstarts);
        VF_LAUNCHED(c);
        VF_TRY((run_ccl<MODE_C1, 6>(grid, g, P, masks, blocks_lin)));
    } else if (nneigh == 6) {
        VF_TRY((run_ccl<MODE_F3, 6>(grid, g, P, masks, blocks_lin)));
    } else {
        VF_TRY((run_ccl<MODE_F3, 26>(grid, g, P, masks, blocks_lin)));
    }
    ccl_mark_kernel<<<(nstarts + 127) / 128, 128, 0, c->stream>>>(P, g, d_starts, nstarts);
    VF_LAUNCHED(c);
    if (mode == MODE_C1) ccl_select_kernel<MODE_C1><<<blocks_lin, 256, 0, c->stream>>>(grid->d, P, masks, g, d_freed);
    else ccl_select_kernel<MODE_F3><<<blocks_lin, 256, 0, c->stream>>>(grid->d, P, masks, g, d_freed);
    VF_LAUNCHED(c);
    return VF_OK;
}
