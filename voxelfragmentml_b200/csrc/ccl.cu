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
//   * the unit of work is a 32-cell z-segment = one row of a 16 x 16 x 32 tile.  Packed 16-bit compares of "same region as my
//     z-predecessor" turn a segment into two 32-bit masks (active cells, run starts); runs, not cells, are the union-find nodes;
//   * two rows have to be united only at "key positions": cells where one of the two rows starts a run while both are active —
//     if both rows merely continue their runs the pair one cell earlier already did the job.  Key positions are bit tricks on
//     the masks, so a thread serves a whole row against a neighbouring row in a handful of instructions;
//   * stage 1 resolves the components inside every tile in shared memory (thread per row), writes each run start's tile-local
//     root to the global parent array (depth-1 forest; only run starts are ever written or read there) and the segment masks
//     next to it; stage 2 (thread per segment) unites only pairs that straddle a tile border; stage 3 marks the roots of the
//     start cells and selects.
// The parent array reuses the 4 B/voxel flood key scratch; the segment records (masks N/4 bytes + first/last labels N/8 bytes)
// the second-grid scratch.
// Requires N < 2^31 (bit 31 of a root's parent word carries "keep").
#include <algorithm>

#include "vf_internal.h"

namespace vfccl {

constexpr uint32_t KEEP = 0x80000000u;
constexpr unsigned kFull = 0xFFFFFFFFu;
enum { MODE_C1 = 0, MODE_F3 = 1 };
constexpr int LX = 16, LY = 16, LZ = 32, LROWS = LX * LY, LCELLS = LROWS * LZ;
constexpr int LABW = LZ / 2 + 1;  // staged label row stride in 32-bit words: odd, so that equal z in consecutive rows hits distinct banks

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

// 26-neighbourhood: the diagonal pair (my cell z, neighbour row's cell z + dz) joins nothing new when a straight pair joins the
// same two runs: the neighbour's cells z and z + dz are active in one run (then the pair at z does it), or my cells z and z + dz are
// (then the pair at z + dz does it) — runs carry one label, and the straight pairs of the same two rows are examined by the dz = 0
// call.  Masks are the unshifted (active, run-start) words of the two rows; cells of other segments never count as continuing.
__device__ __forceinline__ unsigned diagonal_redundant(unsigned Am, unsigned Sm, unsigned An, unsigned Sn, int dz)
{
    if (dz > 0) return (An & (An >> 1) & ~(Sn >> 1)) | ((Am >> 1) & ~(Sm >> 1));
    return (An & (An << 1) & ~Sn) | ((Am << 1) & ~Sm);
}

// position of the run start that covers bit z (S has bit 0 set)
__device__ __forceinline__ int run_start(unsigned S, int z) { return 31 - __clz(S & (0xFFFFFFFFu >> (31 - z))); }

struct Geo {
    int X, Y, Z, segs;  // segs = 32-cell segments per z-row
    int ntx, nty, ntz;
    uint32_t nsegs;
};

// ------------------------------------------------------------------------------------------------ shared-memory union-find
// Node id = row * 32 + z.  Thread-per-row phases touch the same z in 32 consecutive rows at once, which would be one bank; the
// storage slot of a node is therefore rotated by its row: slot(id) = row * 32 + ((z + row) & 31).
__device__ __forceinline__ uint32_t slot(uint32_t id) { return (id & ~31u) | ((id + (id >> 5)) & 31u); }
__device__ __forceinline__ uint32_t find_local(volatile uint32_t* par, uint32_t i)
{
    uint32_t p;
    while ((p = par[slot(i)]) != i) i = p;
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
        const uint32_t old = atomicMin(&par[slot(a)], b);
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
__global__ void __launch_bounds__(256) ccl_tile_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ P, uint2* __restrict__ masks,
                                                       uint32_t* __restrict__ ends, Geo g)
{
    extern __shared__ uint32_t smem_ccl[];
    uint32_t* par = smem_ccl;                                       // [LCELLS] only run-start entries are nodes
    uint32_t* sA = par + LCELLS;                                    // [LROWS] active mask per row
    uint32_t* sS = sA + LROWS;                                      // [LROWS] run-start mask per row (bit 0 always set)
    uint32_t* labw = sS + LROWS;                                    // [LROWS][LABW] staged labels, two per word; row = x * LY + y
    auto lab = [&](int r, int z) -> uint32_t { return (labw[r * LABW + (z >> 1)] >> ((z & 1) * 16)) & 0xFFFFu; };
    const int tz = blockIdx.x;  // 3-D launch: no index divisions
    const int gx0 = blockIdx.z * LX, gy0 = blockIdx.y * LY, gz0 = tz * LZ;

    // (a) thread per 8-cell chunk (4 threads = one row): 128-bit loads straight from the grid, all of a thread's 4 chunks in flight
    //     at once (scalar loads when rows are not 16-byte aligned).  "active" and "differs from the z-predecessor" are evaluated
    //     on packed 16-bit lanes (VIMNMX.U16x2) and gathered into byte masks; two shuffles assemble the row's 32-bit masks.
    //     Cells outside the grid read as EMPTY.
    const bool vec_ok = (g.Z % 8 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 15) == 0);
    const bool half_ok = (g.Z % 4 == 0) && ((reinterpret_cast<uintptr_t>(grid) & 7) == 0);
    constexpr int kIts = LROWS * 4 / 256;  // 8-cell chunks per thread
    uint4 chunk[kIts];
#pragma unroll
    for (int it = 0; it < kIts; ++it) {
        const int q = it * 256 + threadIdx.x, r = q >> 2, ch = q & 3;
        const int gx = gx0 + r / LY, gy = gy0 + r % LY, gz = gz0 + ch * 8;
        chunk[it] = make_uint4(0, 0, 0, 0);
        if (gx < g.X && gy < g.Y && gz < g.Z) {
            const uint16_t* src = grid + ((size_t)gx * g.Y + gy) * g.Z + gz;
            if (vec_ok) {
                chunk[it] = __ldg(reinterpret_cast<const uint4*>(src));
            } else if (half_ok) {  // Z % 8 == 4: rows are 8-byte aligned and the second half of a row's last chunk lies outside the grid
                const uint2 lo = __ldg(reinterpret_cast<const uint2*>(src));
                const uint2 hi = gz + 4 < g.Z ? __ldg(reinterpret_cast<const uint2*>(src + 4)) : make_uint2(0u, 0u);
                chunk[it] = make_uint4(lo.x, lo.y, hi.x, hi.y);
            } else {
                uint32_t w[4] = { 0, 0, 0, 0 };
                for (int k = 0; k < 8 && gz + k < g.Z; ++k) w[k >> 1] |= (uint32_t)src[k] << ((k & 1) * 16);
                chunk[it] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    auto gather8 = [](uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3) -> uint32_t {  // lanes hold 0/1: cell 2k at bit 2k, cell 2k+1 at bit 2k+1
        const uint32_t b = x0 | x1 << 2 | x2 << 4 | x3 << 6;
        return (b | b >> 15) & 0xFFu;
    };
#pragma unroll
    for (int it = 0; it < kIts; ++it) {
        const int q = it * 256 + threadIdx.x, r = q >> 2, ch = q & 3;
        const uint4 v = chunk[it];
        const uint32_t pw = __shfl_up_sync(kFull, v.w, 1);  // last word of the chunk before mine (another row's when ch == 0: masked below)
        const uint32_t one = 0x00010001u, idm = MODE == MODE_C1 ? 0xFFFFFFFFu : 0x00FF00FFu;
        const uint32_t a8 = gather8(__vminu2(v.x & 0x7FFE7FFEu, one), __vminu2(v.y & 0x7FFE7FFEu, one), __vminu2(v.z & 0x7FFE7FFEu, one),
                                    __vminu2(v.w & 0x7FFE7FFEu, one));
        const uint32_t n8 = gather8(__vminu2((v.x ^ __byte_perm(pw, v.x, 0x5432)) & idm, one), __vminu2((v.y ^ __byte_perm(v.x, v.y, 0x5432)) & idm, one),
                                    __vminu2((v.z ^ __byte_perm(v.y, v.z, 0x5432)) & idm, one), __vminu2((v.w ^ __byte_perm(v.z, v.w, 0x5432)) & idm, one));
        uint32_t A = a8 << (8 * ch), Nq = n8 << (8 * ch);
        A |= __shfl_xor_sync(kFull, A, 1), Nq |= __shfl_xor_sync(kFull, Nq, 1);
        A |= __shfl_xor_sync(kFull, A, 2), Nq |= __shfl_xor_sync(kFull, Nq, 2);
        const uint32_t S = ~(~Nq & A & (A << 1));  // a run continues where the cell and its predecessor are active and alike
        uint32_t* lw = &labw[r * LABW + ch * 4];
        lw[0] = v.x, lw[1] = v.y, lw[2] = v.z, lw[3] = v.w;
        // every node starts as its own root.  Only the starts of active runs are nodes (unite_local / find_local are entered through
        // run_start() of an active cell), so only those entries are initialised: one store per row inside a fragment instead of 32
        const uint32_t rb = r * LZ;
        for (uint32_t st = ((S & A) >> (8 * ch)) & 0xFFu; st; st &= st - 1) {
            const uint32_t id = rb + 8 * ch + (__ffs(st) - 1);
            par[slot(id)] = id;
        }
        const uint32_t lastw = __shfl_down_sync(kFull, v.w, 3);  // for ch == 0: the row's last word (cells 30, 31)
        if (ch == 0) {
            const int gx = gx0 + r / LY, gy = gy0 + r % LY;
            sA[r] = A;
            sS[r] = S;
            if (gx < g.X && gy < g.Y) {
                const size_t sg = ((size_t)gx * g.Y + gy) * g.segs + tz;
                masks[sg] = make_uint2(A, S);
                ends[sg] = (v.x & 0xFFFFu) | (lastw & 0xFFFF0000u);  // first and last label of the segment: stage 2's z-border test
            }
        }
    }
    __syncthreads();

    // (a') a tile that lies wholly inside one region — every row one full run, all rows alike: the interior of a fragment, i.e.
    //      most tiles of a solid grid — is one component whatever the neighbourhood: every row's run start points at the tile's
    //      first cell and the union-find phases are skipped.  (A full active mask also means the row lies inside the grid.)
    {
        const int r = threadIdx.x;
        const bool whole = sA[r] == 0xFFFFFFFFu && sS[r] == 1u && same<MODE>(lab(r, 0), lab(0, 0));
        if (__syncthreads_and(whole)) {
            P[((size_t)(gx0 + r / LY) * g.Y + gy0 + r % LY) * g.Z + gz0] = ((uint32_t)gx0 * g.Y + gy0) * g.Z + gz0;
            return;
        }
    }

    // (b) thread per row: unite my runs with the runs of the backward neighbour rows at the key positions
    {
        const int r = threadIdx.x, x = r / LY, y = r % LY;
        const unsigned Am = sA[r], Sm = sS[r];
        auto against = [&](int nx, int ny, int dz) {
            if (nx < 0 || ny < 0 || ny >= LY) return;  // other tiles: stage 2
            const int rn = nx * LY + ny;
            const unsigned An0 = sA[rn], Sn0 = sS[rn];
            unsigned An = An0, Sn = Sn0;
            if (dz < 0) An <<= 1, Sn <<= 1;   // position z of the shifted row is neighbour cell z-1
            if (dz > 0) An >>= 1, Sn >>= 1;   // ... neighbour cell z+1
            unsigned m = (Sm | Sn) & Am & An;
            if (dz != 0) m &= ~diagonal_redundant(Am, Sm, An0, Sn0, dz);
            while (m) {
                const int z = __ffs(m) - 1;
                m &= m - 1;
                const int zn = z + dz;
                const uint32_t mine = lab(r, z);
                if (!same<MODE>(mine, lab(rn, zn))) continue;
                if (nx != x && ny != y) {
                    // diagonal in (x, y): nothing new when (nx, y, z) or (x, ny, z) carries the same label (see stage 2)
                    const int r1 = nx * LY + y, r2 = x * LY + ny;
                    if (((sA[r1] >> z & 1u) && same<MODE>(lab(r1, z), mine)) || ((sA[r2] >> z & 1u) && same<MODE>(lab(r2, z), mine))) continue;
                }
                unite_local(par, r * LZ + run_start(Sm, z), rn * LZ + run_start(sS[rn], zn));
            }
        };
        // (b1) rows of one x-plane first.  The warps move in lockstep, so nearly every thread finds both runs still their own
        //      roots and links them with one atomic: chains along y, at most LY deep.
        if (Am) {
            against(x, y - 1, 0);
            if (NNEIGH == 26) against(x, y - 1, -1), against(x, y - 1, 1);
        }
        __syncthreads();
        // (b2) flatten, so that the finds of the plane-to-plane unions below start one step from a root
        {
            unsigned st = Sm & Am;
            while (st) {
                const int z = __ffs(st) - 1;
                st &= st - 1;
                par[slot(r * LZ + z)] = find_local(par, r * LZ + z);  // writes an ancestor: safe against concurrent finds
            }
        }
        __syncthreads();
        // (b3) plane x against plane x-1
        if (Am) {
            against(x - 1, y, 0);
            if (NNEIGH == 26) {
                against(x - 1, y, -1), against(x - 1, y, 1);
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz) against(x - 1, y - 1, dz), against(x - 1, y + 1, dz);
            }
        }
    }
    __syncthreads();

    // (c) thread per row: every run start gets the global index of its tile-local root.  Only run starts are union-find
    //     nodes: the parent array is written (and later read) at run starts only.
    {
        const int r = threadIdx.x;
        const int gx = gx0 + r / LY, gy = gy0 + r % LY;
        const unsigned Sr = sS[r];
        unsigned st = Sr & sA[r];
        while (st) {
            const int z = __ffs(st) - 1;
            st &= st - 1;
            const uint32_t root = find_local(par, r * LZ + z);
            const int rz = root % LZ, rr = root / LZ;
            P[((size_t)gx * g.Y + gy) * g.Z + gz0 + z] = ((uint32_t)(gx0 + rr / LY) * g.Y + gy0 + rr % LY) * g.Z + gz0 + rz;
        }
    }
}

// ------------------------------------------------------------------------------------------------ stage 2: across tile borders
// thread per 32-cell segment.  A pair (my cell z, neighbour cell z+dz of row (nx,ny)) is this stage's business iff the two cells
// lie in different tiles.  One launch per axis, z then y then x (the order of stage 1's in-tile phases): the unions of a launch
// join trees that the previous launches left at most one tile row / tile plane deep, instead of all borders of a fragment
// contending for the same roots at once.
enum { AXIS_Z = 0, AXIS_Y = 1, AXIS_X = 2 };
template <int MODE, int NNEIGH, int AXIS>
__global__ void __launch_bounds__(256) ccl_border_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ P, const uint2* __restrict__ masks,
                                                         const uint32_t* __restrict__ ends, Geo g)
{
    {
        const uint32_t sg = blockIdx.x * blockDim.x + threadIdx.x;
        if (sg >= g.nsegs) return;
        const int seg = sg % g.segs;
        const uint32_t row = sg / g.segs;
        const int y = (int)(row % g.Y), x = (int)(row / g.Y);
        if (AXIS == AXIS_Z && seg == 0) return;
        if (NNEIGH == 6 && ((AXIS == AXIS_Y && y % LY != 0) || (AXIS == AXIS_X && x % LX != 0))) return;  // rows inside a tile: stage 1
        const uint2 mm = masks[sg];
        const unsigned Am = mm.x, Sm = mm.y;
        if (!Am) return;
        const uint32_t base = row * (uint32_t)g.Z + seg * 32;
        // A pair (i, n) that straddles a tile border need not be united when a "witness" pair one row (or plane) back, inside the
        // same two tiles, carries the same labels: stage 1 united i with its witness and n with its witness, and the witness pair
        // is handled by its own thread (induction on (x, y)).  Only rows on a tile edge, or rows where the labels change, pay.
        const int YZ = g.Y * g.Z;
        auto witnessed = [&](uint32_t i, uint32_t n, uint32_t vi, uint32_t vn, int back) {
            const uint32_t wi = grid[i - back], wn = grid[n - back];
            return active<MODE>(wi) && active<MODE>(wn) && same<MODE>(vi, wi) && same<MODE>(vn, wn);
        };
        // same row, previous segment (always another tile because LZ == 32): decided from the segment records alone
        if (AXIS == AXIS_Z && (Am & 1u)) {
            const uint2 pm = masks[sg - 1];
            if (pm.x >> 31) {
                const uint32_t vi = ends[sg] & 0xFFFFu, vn = ends[sg - 1] >> 16;
                if (same<MODE>(vi, vn)) {
                    auto wit = [&](uint32_t wsg) {  // the same pair one row / one plane back
                        return (masks[wsg].x & 1u) && (masks[wsg - 1].x >> 31) && same<MODE>(vi, ends[wsg] & 0xFFFFu) && same<MODE>(vn, ends[wsg - 1] >> 16);
                    };
                    const bool skip = (y % LY != 0 && wit(sg - g.segs)) || (x % LX != 0 && wit(sg - (uint32_t)g.Y * g.segs));
                    if (!skip) unite(P, base, base - 32 + run_start(pm.y, 31));  // nodes are run starts; cell 0 of a segment always is one
                }
            }
        }
        auto against = [&](int nx, int ny, int dz) {
            if (nx < 0 || ny < 0 || ny >= g.Y) return;
            const bool other_tile = (nx / LX != x / LX) || (ny / LY != y / LY);
            if (!other_tile && (dz == 0 || !(Am & (dz < 0 ? 1u : 0x80000000u)))) return;  // in-tile pairs: stage 1, except the cell that leaves the segment
            const uint32_t nrow = (uint32_t)nx * g.Y + ny;
            const uint32_t nsg = nrow * g.segs + seg;
            const uint2 cm = masks[nsg];
            unsigned An = cm.x, Sn = cm.y;
            uint2 pm = make_uint2(0u, 0u);
            if (dz < 0) {
                if (seg > 0) pm = masks[nsg - 1];
                An = (An << 1) | (pm.x >> 31), Sn = (Sn << 1) | 1u;  // neighbour cell -1 sits in the previous segment: treat as a start
            } else if (dz > 0) {
                const uint2 qm = seg + 1 < g.segs ? masks[nsg + 1] : make_uint2(0u, 0u);
                An = (An >> 1) | (qm.x << 31), Sn = (Sn >> 1) | 0x80000000u;
            }
            unsigned m = (Sm | Sn) & Am & An;
            if (dz != 0) m &= ~diagonal_redundant(Am, Sm, cm.x, cm.y, dz);
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
                if (nx == x) {  // crosses a y border only
                    if (x % LX != 0) skip = witnessed(i, n, vi, vn, YZ);
                } else if (ny == y) {  // crosses an x border only
                    if (y % LY != 0) skip = witnessed(i, n, vi, vn, g.Z);
                } else {
                    // diagonal in (x, y): nothing new when a cell next to both — (nx, y, z) or (x, ny, z) — carries the same label:
                    // it is joined to my cell and to the neighbour by pairs that are straight in (x, y)
                    const uint32_t c1 = grid[((size_t)nx * g.Y + y) * g.Z + seg * 32 + z], c2 = grid[((size_t)x * g.Y + ny) * g.Z + seg * 32 + z];
                    skip = (active<MODE>(c1) && same<MODE>(c1, vi)) || (active<MODE>(c2) && same<MODE>(c2, vi));
                }
                if (skip) continue;
                // union-find nodes are run starts: translate both cells (P[run start] is its tile-local root: depth-1 entry points)
                const int zn = z + dz;
                const uint32_t nrs = zn < 0 ? nbase - 32 + run_start(pm.y, 31) : zn > 31 ? nbase + 32 : nbase + run_start(cm.y, zn);
                unite(P, base + run_start(Sm, z), nrs);
            }
        };
        if (AXIS == AXIS_Y) {
            against(x, y - 1, 0);
            if (NNEIGH == 26) against(x, y - 1, -1), against(x, y - 1, 1);
        }
        if (AXIS == AXIS_X) {
            against(x - 1, y, 0);
            if (NNEIGH == 26) {
                against(x - 1, y, -1), against(x - 1, y, 1);
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz) against(x - 1, y - 1, dz), against(x - 1, y + 1, dz);
            }
        }
    }
}

// mark the root of each start cell's component.
// C1 (c1 != 0): every seed is searched from with its OWN label (NaiveFracturer.cpp:116-146: the front is seeded with all seeds, a neighbour is
// entered when the grid holds the label the front element carries).  A seed whose cell was taken by a later seed — CADScene's copies of the
// original seeds sit on the originals' cells (CADScene.cpp:651) — therefore still keeps the components of ITS label that touch the cell.
__global__ void ccl_mark_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ P, const uint2* __restrict__ masks, Geo g,
                                const ushort4* __restrict__ starts, int S, int c1)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const ushort4 sd = starts[s];
    auto mark = [&](int x, int y, int zz) {
        const uint32_t row = (uint32_t)x * g.Y + y;
        const int seg = zz / 32, z = zz % 32;
        const uint2 mm = masks[(size_t)row * g.segs + seg];
        if (!(mm.x >> z & 1u)) return;  // the cell is not part of any region
        uint32_t r = P[row * (uint32_t)g.Z + seg * 32 + run_start(mm.y, z)];
        r &= ~KEEP;  // the run start may itself be a root that another start already marked
        while (true) {
            const uint32_t q = __ldcg(&P[r]) & ~KEEP;
            if (q == r) break;
            r = q;
        }
        atomicOr(&P[r], KEEP);
    };
    if (c1 && grid[((size_t)sd.x * g.Y + sd.y) * g.Z + sd.z] != sd.w) {
        // the cell carries a later seed's label: this seed's search enters the 6-neighbours that hold its own label
        const int x = sd.x, y = sd.y, z = sd.z;
        if (x + 1 < g.X && grid[((size_t)(x + 1) * g.Y + y) * g.Z + z] == sd.w) mark(x + 1, y, z);
        if (x > 0 && grid[((size_t)(x - 1) * g.Y + y) * g.Z + z] == sd.w) mark(x - 1, y, z);
        if (y + 1 < g.Y && grid[((size_t)x * g.Y + y + 1) * g.Z + z] == sd.w) mark(x, y + 1, z);
        if (y > 0 && grid[((size_t)x * g.Y + y - 1) * g.Z + z] == sd.w) mark(x, y - 1, z);
        if (z + 1 < g.Z && grid[((size_t)x * g.Y + y) * g.Z + z + 1] == sd.w) mark(x, y, z + 1);
        if (z > 0 && grid[((size_t)x * g.Y + y) * g.Z + z - 1] == sd.w) mark(x, y, z - 1);
        return;
    }
    mark(sd.x, sd.y, sd.z);
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
        if (MODE == MODE_C1 && dead && ncell == 32 && ((uintptr_t)(grid + base) & 15) == 0) {
            // most inactive cells are EMPTY already (all of them on a surface model's grid): find the others with four 128-bit loads instead of
            // one 16-bit load per cell
            const uint4* g4 = reinterpret_cast<const uint4*>(grid + base);
            unsigned nonzero = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 v = g4[k];
                const unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int j = 0; j < 4; ++j) nonzero |= ((w[j] & 0xFFFFu) ? 1u : 0u) << (8 * k + 2 * j) | ((w[j] >> 16) ? 1u : 0u) << (8 * k + 2 * j + 1);
            }
            dead &= nonzero;
        }
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

constexpr size_t kTileSmem = (size_t)LCELLS * 4 + (size_t)LROWS * 8 + (size_t)LROWS * LABW * 4;

template <int MODE, int NNEIGH>
vf_status run_ccl(vf_grid* grid, const Geo& g, uint32_t* P, uint2* masks, uint32_t* ends, int blocks_lin)
{
    vf_ctx* c = grid->ctx;
    auto tk = ccl_tile_kernel<MODE, NNEIGH>;
    VF_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTileSmem));
    tk<<<dim3(g.ntz, g.nty, g.ntx), 256, kTileSmem, c->stream>>>(grid->d, P, masks, ends, g);
    VF_LAUNCHED(c);
    const int blocks_seg = (int)((g.nsegs + 255) / 256);
    ccl_border_kernel<MODE, NNEIGH, AXIS_Z><<<blocks_seg, 256, 0, c->stream>>>(grid->d, P, masks, ends, g);
    VF_LAUNCHED(c);
    ccl_border_kernel<MODE, NNEIGH, AXIS_Y><<<blocks_seg, 256, 0, c->stream>>>(grid->d, P, masks, ends, g);
    VF_LAUNCHED(c);
    ccl_border_kernel<MODE, NNEIGH, AXIS_X><<<blocks_seg, 256, 0, c->stream>>>(grid->d, P, masks, ends, g);
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
    VF_TRY(vf_scratch_reserve(c, c->grid2, std::max(n * 2, (size_t)g.nsegs * 12)));
    uint32_t* P = (uint32_t*)c->keys.ptr;
    uint2* masks = (uint2*)c->grid2.ptr;
    uint32_t* ends = (uint32_t*)(masks + g.nsegs);
    const int blocks_lin = c->num_sms * 8;
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    if (!d_freed) {
        d_freed = (uint32_t*)((char*)c->small.ptr + (700 << 10));
        VF_TRY(vf_k_zero(c, d_freed, 4));
    }
    if (mode == MODE_C1) {
        ccl_plant_seeds_kernel<<<(nstarts + 127) / 128, 128, 0, c->stream>>>(grid->d, g, d_starts, nstarts);
        VF_LAUNCHED(c);
        VF_TRY((run_ccl<MODE_C1, 6>(grid, g, P, masks, ends, blocks_lin)));
    } else if (nneigh == 6) {
        VF_TRY((run_ccl<MODE_F3, 6>(grid, g, P, masks, ends, blocks_lin)));
    } else {
        VF_TRY((run_ccl<MODE_F3, 26>(grid, g, P, masks, ends, blocks_lin)));
    }
    ccl_mark_kernel<<<(nstarts + 127) / 128, 128, 0, c->stream>>>(grid->d, P, masks, g, d_starts, nstarts, mode == MODE_C1 ? 1 : 0);
    VF_LAUNCHED(c);
    if (mode == MODE_C1) ccl_select_kernel<MODE_C1><<<blocks_lin, 256, 0, c->stream>>>(grid->d, P, masks, g, d_freed);
    else ccl_select_kernel<MODE_F3><<<blocks_lin, 256, 0, c->stream>>>(grid->d, P, masks, g, d_freed);
    VF_LAUNCHED(c);
    return VF_OK;
}
