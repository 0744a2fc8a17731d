// stencil.cu — C2/C3: boundary marking, erosion ("boundary noise") and the 3^3 isolated-voxel sweep; H1: histogram.
//
// Replaces RegularGrid::detectBoundaries (SRC/DataStructures/RegularGrid.cpp:64-80, detectBoundaries-comp.glsl:18-43),
// RegularGrid::erode (:82-159, erodeGrid-comp.glsl:26-59, copyGrid-comp.glsl), RegularGrid::removeIsolatedRegions
// (:1006-1015, removeIsolatedRegionsGrid-comp.glsl:16-39) and RegularGrid::countValues / numOccupiedVoxels (:601-625, 280-287).
//
// All three stencils are 3^3 neighbourhoods over 2-byte labels: a CTA stages an 8 x 8 x 64 tile plus a 1-cell halo in shared
// memory (coalesced z-rows), so HBM sees each label once per pass (2 B read + 2 B written per voxel) instead of up to 27
// scattered 2-byte loads per voxel as in the shaders.  Erosion ping-pongs between the grid and one scratch grid, which
// removes the reference's copyGrid pass; the final sweep reads a snapshot and writes the other buffer (the reference's
// in-place sweep is racy; DESIGN.md defines snapshot semantics).  Erosion masks larger than 3^3 take a direct global-memory path.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)

#include <cmath>
#include <cstring>
#include <vector>

#include "vf_internal.h"

namespace {

constexpr int SX = 8, SYT = 8, SZT = 64;  // tile
constexpr int HX = SX + 2, HY = SYT + 2, HZ = SZT + 2;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Dims {
    int X, Y, Z;
};

__device__ __forceinline__ int s_at(int x, int y, int z) { return ((x + 1) * HY + (y + 1)) * HZ + (z + 1); }

// stage tile + halo; cells outside the grid read as `outside`
__device__ __forceinline__ void stage(uint16_t* s, const uint16_t* __restrict__ src, Dims d, int gx0, int gy0, int gz0, uint16_t outside)
{
    for (int i = threadIdx.x; i < HX * HY * HZ; i += blockDim.x) {
        const int z = i % HZ - 1, r = i / HZ;
        const int y = r % HY - 1, x = r / HY - 1;
        const int gx = gx0 + x, gy = gy0 + y, gz = gz0 + z;
        uint16_t v = outside;
        if (gx >= 0 && gx < d.X && gy >= 0 && gy < d.Y && gz >= 0 && gz < d.Z) v = src[((size_t)gx * d.Y + gy) * d.Z + gz];
        s[i] = v;
    }
}

enum { OP_DETECT = 0, OP_ERODE3 = 1, OP_SWEEP = 2 };

struct ErodeArgs {
    const float* noise;
    unsigned nnoise;
    float prob, thr, activations;
    int boundary_mode;
    unsigned maskbits;  // 27-bit mask of the 3^3 convolution (bit = (dx+1)*9 + (dy+1)*3 + (dz+1))
    // erodes[visited] bit `count` = (float(count) / float(visited) < activations * thr), evaluated on the host with the same
    // IEEE float32 expression as erodeGrid-comp.glsl:55-56; visited <= 27 cells of the clamped 3^3 box
    unsigned erodes[28];
    unsigned long long nn_magic;  // 2^64 / nnoise + 1: exact 32-bit modulo by multiplication when the grid has < 2^32 cells
    int idx32;
    unsigned long long cell_offset;  // index of this grid's first cell in the grid the noise is indexed by (a slab of a larger grid; else 0)
    // Change tracking for the sparse follow-up passes (erode_sparse_kernel, sweep_sparse_kernel); all null when nothing is recorded.
    // `ever`: one bit per cell, set when some pass of this call empties the cell (a cell is emptied at most once).  The thread that sets the bit
    // appends the cell to this pass's list and to the call's list, so the lists hold no duplicates; a list that outgrows kListCap is ignored by
    // its reader, which then scans `ever` (a superset of any pass's cells: re-deciding more cells than necessary is harmless).
    uint32_t* ever;
    uint32_t* pass_list;   // [kListCap] cells this pass empties
    uint32_t* pass_count;  // entries appended (may exceed kListCap)
    uint32_t* ever_list;   // [kListCap] cells any pass empties
    uint32_t* ever_count;
    // first erosion pass only: cells of ITS INPUT grid that the final 3^3 sweep would empty (fewer than 6 equal neighbours); null: not asked for
    uint32_t* s0kill;
};

constexpr uint32_t kListCap = 1u << 20;

__device__ __forceinline__ void note_eroded(const ErodeArgs& ea, size_t gi)
{
    if (!ea.ever) return;
    const uint32_t bit = 1u << (gi & 31u);
    if (atomicOr(&ea.ever[gi >> 5], bit) & bit) return;
    if (ea.pass_list) {
        const uint32_t j = atomicAdd(ea.pass_count, 1u);
        if (j < kListCap) ea.pass_list[j] = (uint32_t)gi;
    }
    if (ea.ever_list) {
        const uint32_t j = atomicAdd(ea.ever_count, 1u);
        if (j < kListCap) ea.ever_list[j] = (uint32_t)gi;
    }
}

__device__ __forceinline__ unsigned noise_index(const ErodeArgs& ea, size_t gi)
{
    gi += ea.cell_offset;
    if (ea.idx32) return (unsigned)__umul64hi(ea.nn_magic * (unsigned)gi, (unsigned long long)ea.nnoise);
    return (unsigned)(gi % ea.nnoise);
}

// Halo cells outside the grid are staged as EMPTY.  That is exact for all three stencils: detect only looks for labels
// (> FREE); erosion only acts on labelled voxels, which never equal EMPTY, and takes "visited" from the coordinates; the sweep
// compares with `own`, and a voxel whose own word is EMPTY comes out EMPTY whatever it counts.
constexpr uint16_t kOutside = 0u;

template <int OP>
__global__ void __launch_bounds__(256) stencil_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, Dims d, int ntx, int nty, int ntz,
                                                      ErodeArgs ea)
{
    __shared__ uint16_t s[HX * HY * HZ];
    const int tile = blockIdx.x;
    const int tz = tile % ntz, ty = (tile / ntz) % nty, tx = tile / (ntz * nty);
    const int gx0 = tx * SX, gy0 = ty * SYT, gz0 = tz * SZT;
    stage(s, src, d, gx0, gy0, gz0, kOutside);
    __syncthreads();
    const int z = threadIdx.x % SZT, grp = threadIdx.x / SZT;  // 4 groups of 64 lanes; each walks 16 (x,y) columns
    const int gz = gz0 + z;
    if (gz >= d.Z) return;
    for (int c = grp; c < SX * SYT; c += 4) {
        const int x = c / SYT, y = c % SYT;
        const int gx = gx0 + x, gy = gy0 + y;
        if (gx >= d.X || gy >= d.Y) continue;
        const size_t gi = ((size_t)gx * d.Y + gy) * d.Z + gz;
        const uint16_t own = s[s_at(x, y, z)];
        if (OP == OP_DETECT) {
            // detectBoundaries-comp.glsl:21-42 (in place: neighbours are read with bit 15 cleared, so the result does not
            // depend on which neighbours were already tagged by this pass)
            if (own <= VF_VOXEL_FREE) continue;
            bool boundary = false;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) {
                        const uint16_t raw = s[s_at(x + dx, y + dy, z + dz)];
                        const uint16_t v = raw & 0x7FFFu;
                        boundary = boundary || (v > VF_VOXEL_FREE && v != own);
                    }
            if (boundary) dst[gi] = own | 0x8000u;
        } else if (OP == OP_ERODE3) {
            // erodeGrid-comp.glsl:31-58 for maskSize 3
            uint16_t out = own;
            const bool isB = ea.boundary_mode == 0 ? (own & 0x7FFFu) != 0 : (own >> 15) != 0;
            if (own > VF_VOXEL_FREE && isB && ea.noise[noise_index(ea, gi)] < ea.prob) {
                unsigned count = 0, visited = 0;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                        for (int dz = -1; dz <= 1; ++dz) {
                            const uint16_t raw = s[s_at(x + dx, y + dy, z + dz)];
                            const bool inside = gx + dx >= 0 && gx + dx < d.X && gy + dy >= 0 && gy + dy < d.Y && gz + dz >= 0 && gz + dz < d.Z;
                            visited += inside;
                            const unsigned bit = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1);
                            count += (inside && raw == own && (ea.maskbits >> bit & 1u));
                        }
                const float activation = __fdiv_rn((float)count, (float)visited);
                if (activation < __fmul_rn(ea.activations, ea.thr)) out = VF_VOXEL_EMPTY, note_eroded(ea, gi);
            }
            dst[gi] = out;
        } else {
            // removeIsolatedRegionsGrid-comp.glsl:24-38, snapshot semantics
            int count = -1;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dz = -1; dz <= 1; ++dz) {
                        const bool inside = gx + dx >= 0 && gx + dx < d.X && gy + dy >= 0 && gy + dy < d.Y && gz + dz >= 0 && gz + dz < d.Z;
                        count += (inside && s[s_at(x + dx, y + dy, z + dz)] == own);
                    }
            dst[gi] = count < 6 ? (uint16_t)VF_VOXEL_EMPTY : own;
        }
    }
}

// erosion with an arbitrary odd mask size k > 3: direct global reads (rare path; GUI lets the user pick larger kernels)
__global__ void __launch_bounds__(256) erode_generic_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, Dims d,
                                                            const float* __restrict__ mask, int k, ErodeArgs ea)
{
    const size_t n = (size_t)d.X * d.Y * d.Z;
    const int k2 = k / 2;
    for (size_t gi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gi < n; gi += (size_t)gridDim.x * blockDim.x) {
        const uint16_t own = src[gi];
        uint16_t out = own;
        const bool isB = ea.boundary_mode == 0 ? (own & 0x7FFFu) != 0 : (own >> 15) != 0;
        if (own > VF_VOXEL_FREE && isB && ea.noise[noise_index(ea, gi)] < ea.prob) {
            const int z = (int)(gi % d.Z);
            const size_t r = gi / d.Z;
            const int y = (int)(r % d.Y), x = (int)(r / d.Y);
            const int mnx = x - k2, mny = y - k2, mnz = z - k2;
            const int x0 = max(mnx, 0), x1 = min(x + k2, d.X - 1), y0 = max(mny, 0), y1 = min(y + k2, d.Y - 1), z0 = max(mnz, 0), z1 = min(z + k2, d.Z - 1);
            unsigned count = 0, visited = 0;
            for (int a = x0; a <= x1; ++a)
                for (int b = y0; b <= y1; ++b)
                    for (int c = z0; c <= z1; ++c) {
                        const float w = mask[((a - mnx) * k + (b - mny)) * k + (c - mnz)];
                        count += (unsigned)__fmul_rn((float)(unsigned)(src[((size_t)a * d.Y + b) * d.Z + c] == own), w);
                        ++visited;
                    }
            const float activation = __fdiv_rn((float)count, (float)visited);
            if (activation < __fmul_rn(ea.activations, ea.thr)) out = VF_VOXEL_EMPTY;
        }
        dst[gi] = out;
    }
}

// Sparse follow-up pass of the 3^3 erosion (iterations 2..K).  Erosion is a function of a cell's word, its 3^3 window and its noise value: a
// cell whose window did not change since the previous pass decides as it did then, i.e. it stays.  So pass k only has to look at the cells
// whose window holds a cell that pass k-1 emptied — the 27 positions around every bit of that pass's bitmap — and `dst`, which still holds
// the grid of two passes ago, already carries the right word everywhere else (it differs from `src` exactly at the marked cells, which are
// among the 27).  Each candidate is evaluated exactly as in the full pass (clamped box, mask, noise, float32 comparison through the same
// table); a candidate reached from several marked neighbours gets the same answer from each.  The bitmaps have a fixed size (one bit per
// cell), so nothing here depends on how many cells a pass empties and the host never has to look: the whole erode call is asynchronous.
// A warp scans 32 bitmap words per step and hands the set bits out to its lanes.
// one candidate of the sparse erosion pass: the cell at position p (0..26) of the 3^3 box around the marked cell u
__device__ __forceinline__ void erode_candidate(const uint16_t* __restrict__ src, uint16_t* dst, const Dims& d, const ErodeArgs& ea, uint32_t u, int p)
{
    const int uz = (int)(u % (uint32_t)d.Z);
    const uint32_t r = u / (uint32_t)d.Z;
    const int uy = (int)(r % (uint32_t)d.Y), ux = (int)(r / (uint32_t)d.Y);
    const int x = ux + p / 9 - 1, y = uy + (p / 3) % 3 - 1, z = uz + p % 3 - 1;
    if ((unsigned)x >= (unsigned)d.X || (unsigned)y >= (unsigned)d.Y || (unsigned)z >= (unsigned)d.Z) return;
    const size_t gi = ((size_t)x * d.Y + y) * d.Z + z;
    const uint16_t own = src[gi];
    bool erodes = false;
    const bool isB = ea.boundary_mode == 0 ? (own & 0x7FFFu) != 0 : (own >> 15) != 0;
    if (own > VF_VOXEL_FREE && isB && ea.noise[noise_index(ea, gi)] < ea.prob) {
        // visited = cells of the clamped box; only the cells the mask selects are read (7 of 27 for ELLIPSE / CROSS)
        const unsigned visited = (1 + (x > 0) + (x < d.X - 1)) * (1 + (y > 0) + (y < d.Y - 1)) * (1 + (z > 0) + (z < d.Z - 1));
        unsigned count = 0;
        for (unsigned m = ea.maskbits; m; m &= m - 1) {
            const int bit = __ffs(m) - 1;
            const int a = x + bit / 9 - 1, b = y + (bit / 3) % 3 - 1, c = z + bit % 3 - 1;
            if ((unsigned)a >= (unsigned)d.X || (unsigned)b >= (unsigned)d.Y || (unsigned)c >= (unsigned)d.Z) continue;
            count += src[((size_t)a * d.Y + b) * d.Z + c] == own;
        }
        erodes = ea.erodes[visited] >> count & 1u;
    }
    dst[gi] = erodes ? (uint16_t)VF_VOXEL_EMPTY : own;  // equal to what dst holds unless this or the previous pass empties the cell
    if (erodes) note_eroded(ea, gi);
}

__global__ void __launch_bounds__(256) erode_sparse_kernel(const uint16_t* __restrict__ src, uint16_t* dst, Dims d, ErodeArgs ea, const uint32_t* __restrict__ marked,
                                                           uint32_t nwords, const uint32_t* __restrict__ in_list, const uint32_t* __restrict__ in_count)
{
    const uint32_t listed = *in_count;
    if (listed <= kListCap) {  // the usual case: the previous pass's cells as a list, 27 candidates each, dealt evenly to all threads
        const uint64_t total = (uint64_t)listed * 27u;
        for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total; w += (uint64_t)gridDim.x * blockDim.x)
            erode_candidate(src, dst, d, ea, in_list[w / 27u], (int)(w % 27u));
        return;
    }
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x / 32), wid = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    for (uint32_t base = wid * 32u; base < nwords; base += nwarps * 32u) {
        const uint32_t mine = base + lane < nwords ? marked[base + lane] : 0u;
        for (unsigned have = __ballot_sync(0xFFFFFFFFu, mine != 0); have; have &= have - 1) {
            const int src_lane = __ffs(have) - 1;
            uint32_t bits = __shfl_sync(0xFFFFFFFFu, mine, src_lane);
            const uint32_t word = base + src_lane;
            while (bits) {  // one marked cell at a time: its 27 neighbourhood positions go to lanes 0..26
                const uint32_t u = word * 32u + (__ffs(bits) - 1);
                bits &= bits - 1;
                if (lane < 27) erode_candidate(src, dst, d, ea, u, lane);
            }
        }
    }
}

// Sparse final sweep (removeIsolatedRegionsGrid, snapshot semantics).  The first erosion pass has already decided the sweep for its input grid
// (s0kill, see stencil_pair); that verdict holds for every cell whose 3^3 window no erosion pass touched.  sweep_sparse_kernel re-decides the 27
// cells around every cell some pass emptied (`ever`) on the final erosion output `src`, setting or clearing their bits; sweep_apply_kernel then
// empties the marked cells of the caller's grid — and, when that grid holds the output of the pass before the last, the last pass's cells too.
__device__ __forceinline__ void sweep_candidate(const uint16_t* __restrict__ src, const Dims& d, uint32_t* kill, uint32_t u, int p)
{
    const int uz = (int)(u % (uint32_t)d.Z);
    const uint32_t r = u / (uint32_t)d.Z;
    const int uy = (int)(r % (uint32_t)d.Y), ux = (int)(r / (uint32_t)d.Y);
    const int x = ux + p / 9 - 1, y = uy + (p / 3) % 3 - 1, z = uz + p % 3 - 1;
    if ((unsigned)x >= (unsigned)d.X || (unsigned)y >= (unsigned)d.Y || (unsigned)z >= (unsigned)d.Z) return;
    const size_t gi = ((size_t)x * d.Y + y) * d.Z + z;
    const uint16_t own = src[gi];
    int count = -1;
    for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz) {
                const int a = x + dx, b = y + dy, c = z + dz;
                if ((unsigned)a >= (unsigned)d.X || (unsigned)b >= (unsigned)d.Y || (unsigned)c >= (unsigned)d.Z) continue;
                count += src[((size_t)a * d.Y + b) * d.Z + c] == own;
            }
    const uint32_t bit = 1u << (gi & 31u);
    if (own != VF_VOXEL_EMPTY && count < 6) atomicOr(&kill[gi >> 5], bit);
    else atomicAnd(&kill[gi >> 5], ~bit);
}

__global__ void __launch_bounds__(256) sweep_sparse_kernel(const uint16_t* __restrict__ src, Dims d, const uint32_t* __restrict__ ever, uint32_t* kill, uint32_t nwords,
                                                           const uint32_t* __restrict__ ever_list, const uint32_t* __restrict__ ever_count)
{
    const uint32_t listed = *ever_count;
    if (listed <= kListCap) {
        const uint64_t total = (uint64_t)listed * 27u;
        for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total; w += (uint64_t)gridDim.x * blockDim.x)
            sweep_candidate(src, d, kill, ever_list[w / 27u], (int)(w % 27u));
        return;
    }
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (blockDim.x / 32), wid = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    for (uint32_t base = wid * 32u; base < nwords; base += nwarps * 32u) {
        const uint32_t mine = base + lane < nwords ? ever[base + lane] : 0u;
        for (unsigned have = __ballot_sync(0xFFFFFFFFu, mine != 0); have; have &= have - 1) {
            const int src_lane = __ffs(have) - 1;
            uint32_t bits = __shfl_sync(0xFFFFFFFFu, mine, src_lane);
            const uint32_t word = base + src_lane;
            while (bits) {
                const uint32_t u = word * 32u + (__ffs(bits) - 1);
                bits &= bits - 1;
                if (lane < 27) sweep_candidate(src, d, kill, u, lane);
            }
        }
    }
}

// `last_list` / `last_count`: the cells the last erosion pass emptied, when `grid` holds the output of the pass before it (null otherwise); a list
// that overflowed is replaced by `ever` (every cell in it is EMPTY in the final erosion output)
__global__ void __launch_bounds__(256) sweep_apply_kernel(uint16_t* __restrict__ grid, const uint32_t* __restrict__ kill, const uint32_t* __restrict__ ever,
                                                          const uint32_t* __restrict__ last_list, const uint32_t* __restrict__ last_count, uint32_t nwords, size_t n)
{
    const bool use_ever = last_count != nullptr && *last_count > kListCap;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        uint32_t bits = kill[w] | (use_ever ? ever[w] : 0u);
        while (bits) {
            const size_t gi = (size_t)w * 32u + (__ffs(bits) - 1);
            bits &= bits - 1;
            if (gi < n) grid[gi] = VF_VOXEL_EMPTY;
        }
    }
    if (last_count != nullptr && !use_ever) {
        const uint32_t cnt = *last_count;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) grid[last_list[i]] = VF_VOXEL_EMPTY;
    }
}

// detectBoundaries with boundarySize b > 1 (the reference itself only ever passes 1: CADScene.cpp:687, RegularGrid.cpp:135): direct global
// reads over the clamped (2b + 1)^3 box, as detectBoundaries-comp.glsl:24-40.  In place like the shader: neighbours are compared with bit 15
// cleared against the centre's raw word, which this pass alone may change, so the result does not depend on the order of the threads.
__global__ void __launch_bounds__(256) detect_generic_kernel(uint16_t* grid, Dims d, int b)
{
    const size_t n = (size_t)d.X * d.Y * d.Z;
    for (size_t gi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; gi < n; gi += (size_t)gridDim.x * blockDim.x) {
        const uint16_t own = grid[gi];
        if (own <= VF_VOXEL_FREE) continue;
        const int z = (int)(gi % d.Z);
        const size_t r = gi / d.Z;
        const int y = (int)(r % d.Y), x = (int)(r / d.Y);
        const int x0 = max(x - b, 0), x1 = min(x + b, d.X - 1), y0 = max(y - b, 0), y1 = min(y + b, d.Y - 1), z0 = max(z - b, 0), z1 = min(z + b, d.Z - 1);
        bool boundary = false;
        for (int a = x0; a <= x1 && !boundary; ++a)
            for (int bb = y0; bb <= y1 && !boundary; ++bb)
                for (int c = z0; c <= z1 && !boundary; ++c) {
                    const uint16_t v = ((volatile uint16_t*)grid)[((size_t)a * d.Y + bb) * d.Z + c] & 0x7FFFu;
                    boundary = v > VF_VOXEL_FREE && v != own;
                }
        if (boundary) grid[gi] = own | 0x8000u;
    }
}

// ------------------------------------------------------------------------------------------------ H1 histogram
// One pass over 2 B/voxel.  A lane reads four consecutive 8-cell vectors (64 bytes; a warp covers 2 KB) and keeps a (label, count) run in
// registers: a vector that holds one label — 85 % of them on the cfg3 grid — costs a compare and an add, and a shared-memory atomic only
// when the lane's label changes.  Vectors that straddle a label change are queued in shared memory and counted 32 at a time, one vector per
// lane, run by run, so that the warp does not walk the slow path for the sake of a few lanes.  Shared bins for ids < 4096, global atomics
// beyond.  UNMASK: the same pass also clears bit 15 (undoMask) — vectors that carry a tag are written back, the others are only read.
constexpr int kSmemBins = 4096;
constexpr int kHistBatch = 4, kHistWarps = 8, kHistQueue = 160;  // queue: up to 31 waiting + 4 x 32 pushed per step

template <bool UNMASK>
__global__ void __launch_bounds__(kHistWarps * 32) histogram_kernel(uint16_t* __restrict__ grid, size_t n, uint32_t* __restrict__ counts,
                                                                    unsigned long long* __restrict__ occupied, uint32_t* __restrict__ maxbin)
{
    __shared__ uint32_t bins[kSmemBins];
    __shared__ uint4 queue[kHistWarps][kHistQueue];
    for (int i = threadIdx.x; i < kSmemBins; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint4* q = queue[threadIdx.x >> 5];
    int qn = 0;
    unsigned occ = 0;
    const size_t nvec = n / 8;
    uint4* g4 = reinterpret_cast<uint4*>(grid);
    uint32_t top = 0;  // highest bin this thread touched in global memory
    auto add = [&](uint32_t label, uint32_t c) {
        if (label < kSmemBins) atomicAdd(&bins[label], c);
        else atomicAdd(&counts[label], c), top = max(top, label);
    };
    // the cells of one queued vector, run by run (RegularGrid.cpp:612: raw value > FREE, then unmask)
    auto count_runs = [&](const uint4& v) {
        const uint32_t one = 0x00010001u;  // run starts: bit k set <=> cell k differs from cell k - 1
        const uint32_t n0 = __vminu2(v.x ^ __byte_perm(v.x, 0, 0x1010), one), n1 = __vminu2(v.y ^ __byte_perm(v.x, v.y, 0x5432), one);
        const uint32_t n2 = __vminu2(v.z ^ __byte_perm(v.y, v.z, 0x5432), one), n3 = __vminu2(v.w ^ __byte_perm(v.z, v.w, 0x5432), one);
        const uint32_t bb = n0 | n1 << 2 | n2 << 4 | n3 << 6;
        const unsigned long long lo = (unsigned long long)v.y << 32 | v.x, hi = (unsigned long long)v.w << 32 | v.z;
        for (uint32_t st = ((bb | bb >> 15) & 0xFFu) | 1u; st;) {
            const int a = __ffs(st) - 1;
            st &= st - 1;
            const int len = (st ? __ffs(st) - 1 : 8) - a;
            const uint32_t raw = (uint32_t)((a < 4 ? lo >> (16 * a) : hi >> (16 * (a - 4))) & 0xFFFFu);
            if (raw > VF_VOXEL_FREE) {
                occ += len;
                add(raw & 0x7FFFu, (uint32_t)len);
            }
        }
    };
    uint32_t run_raw = 0, run = 0;  // the lane's current run: `run` cells holding the word run_raw (> FREE)
    const size_t span = 32 * kHistBatch;
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x / 32);
    for (size_t base = ((size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * span; base < nvec; base += nwarps * span) {
        uint4 v[kHistBatch];
        const size_t i0 = base + (size_t)lane * kHistBatch;
#pragma unroll
        for (int k = 0; k < kHistBatch; ++k) v[k] = i0 + k < nvec ? g4[i0 + k] : make_uint4(0, 0, 0, 0);  // EMPTY counts nowhere
        if (UNMASK) {
#pragma unroll
            for (int k = 0; k < kHistBatch; ++k)
                if ((v[k].x | v[k].y | v[k].z | v[k].w) & 0x80008000u)  // out-of-range vectors were loaded as zeros
                    g4[i0 + k] = make_uint4(v[k].x & 0x7FFF7FFFu, v[k].y & 0x7FFF7FFFu, v[k].z & 0x7FFF7FFFu, v[k].w & 0x7FFF7FFFu);
        }
#pragma unroll
        for (int k = 0; k < kHistBatch; ++k) {
            const uint32_t first = v[k].x & 0xFFFFu;
            const bool uni = v[k].x == first * 0x10001u && v[k].y == v[k].x && v[k].z == v[k].x && v[k].w == v[k].x;
            if (uni && first > VF_VOXEL_FREE) {
                if (first == run_raw) {
                    run += 8;
                } else {
                    if (run) occ += run, add(run_raw & 0x7FFFu, run);
                    run_raw = first, run = 8;
                }
            }
            const unsigned m = __ballot_sync(kFull, !uni);
            if (m) {
                if (!uni) q[qn + __popc(m & ((1u << lane) - 1u))] = v[k];
                qn += __popc(m);
            }
        }
        __syncwarp();
        while (qn >= 32) {
            qn -= 32;
            count_runs(q[qn + lane]);
            __syncwarp();
        }
    }
    if (lane < qn) count_runs(q[lane]);
    if (blockIdx.x == 0 && threadIdx.x < n % 8) {
        const uint16_t raw = grid[nvec * 8 + threadIdx.x];
        if (raw > VF_VOXEL_FREE) occ += 1, add(raw & 0x7FFFu, 1u);
        if (UNMASK && (raw & 0x8000u)) grid[nvec * 8 + threadIdx.x] = raw & 0x7FFFu;
    }
    if (run) occ += run, add(run_raw & 0x7FFFu, run);
    occ = __reduce_add_sync(kFull, occ);
    if (lane == 0 && occ) atomicAdd(occupied, (unsigned long long)occ);
    __syncthreads();
    for (int i = threadIdx.x; i < kSmemBins; i += blockDim.x)
        if (bins[i]) atomicAdd(&counts[i], bins[i]), top = max(top, (uint32_t)i);
    top = __reduce_max_sync(kFull, top);
    if (lane == 0 && top) atomicMax(maxbin, top);  // the host reads back (and re-zeroes) the bins up to here only
}

// ------------------------------------------------------------------------------------------------ fast 3^3 stencils
// Same tile (8 x 8 x 64 + halo), same exact per-voxel code, but the per-voxel code only runs where it can matter.
// Every 8-cell chunk of the staged region gets a summary word (first value | "the 10 cells z-1..z+8 of this row all hold it");
// an interior chunk whose nine rows carry the same summary has a uniform 3 x 3 x 10 window, and for a uniform window all three
// stencils are the identity (detect: no differing neighbour; erode: count == mask population, decided once per launch;
// sweep: 26 equal neighbours).  Those chunks are copied as 128-bit vectors (detect writes nothing at all).  The remaining
// chunks — label boundaries, tagged shells, grid faces — are queued in shared memory and their voxels dealt evenly to all 256
// threads, so a tile with a thin boundary costs a few balanced rounds of the 27-cell code instead of stalling whole warps.
constexpr int FRS = 80;                 // staged row stride in cells: z = -1 at slot 7, z = 0..63 at 8..71, z = 64 at 72
constexpr int FROWS = HX * HY;          // 100 staged rows
__device__ __forceinline__ int f_at(int x, int y, int z) { return ((x + 1) * HY + (y + 1)) * FRS + z + 8; }

// shared-memory offset of window cell `bit` = (dx+1)*9 + (dy+1)*3 + (dz+1) relative to the centre
__constant__ int c_win_off[27] = {
    (-1 * HY - 1) * FRS - 1, (-1 * HY - 1) * FRS, (-1 * HY - 1) * FRS + 1, (-1 * HY) * FRS - 1, (-1 * HY) * FRS, (-1 * HY) * FRS + 1,
    (-1 * HY + 1) * FRS - 1, (-1 * HY + 1) * FRS, (-1 * HY + 1) * FRS + 1,
    (-1) * FRS - 1, (-1) * FRS, (-1) * FRS + 1, -1, 0, 1, FRS - 1, FRS, FRS + 1,
    (HY - 1) * FRS - 1, (HY - 1) * FRS, (HY - 1) * FRS + 1, HY * FRS - 1, HY * FRS, HY * FRS + 1,
    (HY + 1) * FRS - 1, (HY + 1) * FRS, (HY + 1) * FRS + 1
};

// Exact stencils for a PAIR of z-adjacent voxels A = (x, y, z), B = (x, y, z + 1), z even, on packed 16-bit lanes.
// Per staged row three aligned words W0 = cells (z-2, z-1), W1 = (z, z+1), W2 = (z+2, z+3) give the three compare words
//   C1 = (z-1, z)   both lanes neighbours of A,      C2 = (z+1, z+2)  both lanes neighbours of B,
//   C3 = W1         lane z is a neighbour of B, lane z+1 a neighbour of A,
// so the 2 x 27 comparisons of the pair cost 27 packed ones.  "differs from own" per lane is min(word ^ own, 1) (VIMNMX.U16x2),
// accumulated with a plain add (<= 27 per lane).  Cells outside the grid hold EMPTY (see kOutside).
__device__ __forceinline__ unsigned ne2(unsigned w, unsigned own2) { return __vminu2(w ^ own2, 0x00010001u); }
// detect: lane != 0 iff the cell holds a label (> FREE, tag ignored) different from own (own2m = own without tags)
__device__ __forceinline__ unsigned other2(unsigned w, unsigned own2m) { return __vminu2(w & 0x7FFE7FFEu, (w & 0x7FFF7FFFu) ^ own2m); }

struct PairAcc {
    unsigned a1 = 0, a2 = 0, a3 = 0, a4 = 0;  // C1 (A|A), C2 (B|B), C3 (B|A), face rows (A|B)
};

constexpr unsigned kStarMask = 0x00417410u;  // centre + 6 face neighbours (ELLIPSE / CROSS of size 3)
constexpr unsigned kFullMask = 0x07FFFFFFu;

template <int OP>
__device__ __forceinline__ void pair_row(const uint16_t* p, unsigned AA, unsigned BB, unsigned BA, PairAcc& acc)
{
    const unsigned W0 = *reinterpret_cast<const unsigned*>(p - 2), W1 = *reinterpret_cast<const unsigned*>(p), W2 = *reinterpret_cast<const unsigned*>(p + 2);
    const unsigned C1 = __byte_perm(W0, W1, 0x5432), C2 = __byte_perm(W1, W2, 0x5432);
    if (OP == OP_DETECT) {
        acc.a1 |= other2(C1, AA), acc.a2 |= other2(C2, BB), acc.a3 |= other2(W1, BA);
    } else {
        acc.a1 += ne2(C1, AA), acc.a2 += ne2(C2, BB), acc.a3 += ne2(W1, BA);
    }
}

// generic 3^3 erosion mask (anything but the star and the full cube): per-voxel count over the set bits
__device__ __forceinline__ unsigned masked_count(const uint16_t* ctr, unsigned maskbits)
{
    const unsigned own = *ctr;
    unsigned count = 0;
    for (unsigned m = maskbits; m; m &= m - 1) count += ctr[c_win_off[__ffs(m) - 1]] == own;  // mask bits are warp-uniform
    return count;
}

template <int OP>
__device__ __forceinline__ void stencil_pair(const uint16_t* s, const Dims& d, int x, int y, int z, int gx, int gy, int gz, size_t gi, const ErodeArgs& ea,
                                             uint16_t* __restrict__ dst)
{
    const uint16_t* ctr = s + f_at(x, y, z);
    const unsigned own2 = *reinterpret_cast<const unsigned*>(ctr);
    const unsigned ownA = own2 & 0xFFFFu, ownB = own2 >> 16;
    PairAcc acc;
    if (OP == OP_DETECT) {
        // detectBoundaries-comp.glsl:21-42; in place: neighbours are compared with bit 15 cleared, so the result does not depend
        // on which neighbours this pass has already tagged
        const unsigned m2 = own2 & 0x7FFF7FFFu;
        const unsigned AA = __byte_perm(m2, 0, 0x1010), BB = __byte_perm(m2, 0, 0x3232), BA = __byte_perm(m2, 0, 0x1032);
#pragma unroll
        for (int r = 0; r < 9; ++r) pair_row<OP>(ctr + ((r / 3 - 1) * HY + (r % 3 - 1)) * FRS, AA, BB, BA, acc);
        const bool otherA = ((acc.a1 | (acc.a1 >> 16) | (acc.a3 >> 16)) & 0xFFFFu) != 0, otherB = ((acc.a2 | (acc.a2 >> 16) | acc.a3) & 0xFFFFu) != 0;
        // unlabelled or already tagged voxels stay as they are
        const bool tagA = otherA && ownA > VF_VOXEL_FREE && !(ownA & 0x8000u), tagB = otherB && ownB > VF_VOXEL_FREE && !(ownB & 0x8000u);
        if (tagA || tagB) *reinterpret_cast<unsigned*>(dst + gi) = own2 | (tagA ? 0x8000u : 0u) | (tagB ? 0x80000000u : 0u);
        return;
    }
    const unsigned AA = __byte_perm(own2, 0, 0x1010), BB = __byte_perm(own2, 0, 0x3232), BA = __byte_perm(own2, 0, 0x1032);
    if (OP == OP_SWEEP) {
        // removeIsolatedRegionsGrid-comp.glsl:24-38, snapshot semantics: fewer than 6 equal neighbours -> EMPTY
#pragma unroll
        for (int r = 0; r < 9; ++r) pair_row<OP>(ctr + ((r / 3 - 1) * HY + (r % 3 - 1)) * FRS, AA, BB, BA, acc);
        const unsigned neA = (acc.a1 & 0xFFFFu) + (acc.a1 >> 16) + (acc.a3 >> 16), neB = (acc.a2 & 0xFFFFu) + (acc.a2 >> 16) + (acc.a3 & 0xFFFFu);
        const unsigned outA = 26u - neA < 6u ? (unsigned)VF_VOXEL_EMPTY : ownA, outB = 26u - neB < 6u ? (unsigned)VF_VOXEL_EMPTY : ownB;
        *reinterpret_cast<unsigned*>(dst + gi) = outA | outB << 16;
        return;
    }
    // erodeGrid-comp.glsl:31-58 for maskSize 3
    unsigned countA, countB;
    if (ea.s0kill) {
        // First pass of a call whose final sweep runs sparsely: one walk over the nine rows gives the sweep's verdict on this pass's INPUT
        // (removeIsolatedRegionsGrid-comp.glsl:24-38; it stays valid for every cell whose 3^3 window no erosion pass touches) AND the erosion
        // count: the 27-cell count is the walk's total, the 7-cell count of ELLIPSE / CROSS is the part of it that the six face cells contribute.
        PairAcc full;
        unsigned faceA = 0, faceB = 0;  // differing cells among the six face neighbours
        auto walk = [&](int r) {
            const uint16_t* p = ctr + ((r / 3 - 1) * HY + (r % 3 - 1)) * FRS;
            const unsigned W0 = *reinterpret_cast<const unsigned*>(p - 2), W1 = *reinterpret_cast<const unsigned*>(p), W2 = *reinterpret_cast<const unsigned*>(p + 2);
            const unsigned n1 = ne2(__byte_perm(W0, W1, 0x5432), AA), n2 = ne2(__byte_perm(W1, W2, 0x5432), BB), n3 = ne2(W1, BA);
            full.a1 += n1, full.a2 += n2, full.a3 += n3;
            if (r == 4) faceA += (n1 & 0xFFFFu) + (n3 >> 16), faceB += (n3 & 0xFFFFu) + (n2 >> 16);        // z - 1 and z + 1 of either cell
            else if (r == 1 || r == 3 || r == 5 || r == 7) faceA += n1 >> 16, faceB += n2 & 0xFFFFu;  // the cell's own z in the x / y face rows
        };
        // the centre row and the four face rows first: 14 of a cell's 26 neighbours.  Six equal ones among them already clear the cell for the
        // sweep, which is the case nearly everywhere (a cell on a flat region border still has 11); the four corner rows are walked only for
        // pairs that are not cleared yet, or when the erosion count itself needs all 27 cells (SQUARE)
        walk(1), walk(3), walk(4), walk(5), walk(7);
        unsigned neA = (full.a1 & 0xFFFFu) + (full.a1 >> 16) + (full.a3 >> 16), neB = (full.a2 & 0xFFFFu) + (full.a2 >> 16) + (full.a3 & 0xFFFFu);
        const bool clearedA = ownA == VF_VOXEL_EMPTY || 14u - neA >= 6u, clearedB = ownB == VF_VOXEL_EMPTY || 14u - neB >= 6u;
        if (!(clearedA && clearedB) || ea.maskbits != kStarMask) {
            walk(0), walk(2), walk(6), walk(8);
            neA = (full.a1 & 0xFFFFu) + (full.a1 >> 16) + (full.a3 >> 16), neB = (full.a2 & 0xFFFFu) + (full.a2 >> 16) + (full.a3 & 0xFFFFu);
            const unsigned bits = (ownA != VF_VOXEL_EMPTY && 26u - neA < 6u ? 1u : 0u) | (ownB != VF_VOXEL_EMPTY && 26u - neB < 6u ? 2u : 0u);
            if (bits) atomicOr(&ea.s0kill[gi >> 5], bits << (gi & 31u));  // gi is even: both bits land in one word
        }
        if (ea.maskbits == kStarMask) countA = 7u - faceA, countB = 7u - faceB;
        else if (ea.maskbits == kFullMask) countA = 27u - neA, countB = 27u - neB;
        else countA = masked_count(ctr, ea.maskbits), countB = masked_count(ctr + 1, ea.maskbits);
    } else if (ea.maskbits == kStarMask) {
        pair_row<OP>(ctr, AA, BB, BA, acc);
        acc.a4 = ne2(*reinterpret_cast<const unsigned*>(ctr - HY * FRS), own2) + ne2(*reinterpret_cast<const unsigned*>(ctr + HY * FRS), own2) +
                 ne2(*reinterpret_cast<const unsigned*>(ctr - FRS), own2) + ne2(*reinterpret_cast<const unsigned*>(ctr + FRS), own2);
        countA = 7u - ((acc.a1 & 0xFFFFu) + (acc.a1 >> 16) + (acc.a3 >> 16) + (acc.a4 & 0xFFFFu));
        countB = 7u - ((acc.a2 & 0xFFFFu) + (acc.a2 >> 16) + (acc.a3 & 0xFFFFu) + (acc.a4 >> 16));
    } else if (ea.maskbits == kFullMask) {
#pragma unroll
        for (int r = 0; r < 9; ++r) pair_row<OP>(ctr + ((r / 3 - 1) * HY + (r % 3 - 1)) * FRS, AA, BB, BA, acc);
        countA = 27u - ((acc.a1 & 0xFFFFu) + (acc.a1 >> 16) + (acc.a3 >> 16));
        countB = 27u - ((acc.a2 & 0xFFFFu) + (acc.a2 >> 16) + (acc.a3 & 0xFFFFu));
    } else {
        countA = masked_count(ctr, ea.maskbits), countB = masked_count(ctr + 1, ea.maskbits);
    }
    const int vxy = (1 + (gx > 0) + (gx < d.X - 1)) * (1 + (gy > 0) + (gy < d.Y - 1));
    const unsigned ia = noise_index(ea, gi), ib = ia + 1 == ea.nnoise ? 0u : ia + 1;
    auto eroded = [&](unsigned own, unsigned count, int zz, unsigned ni, size_t cell) -> unsigned {
        const bool isB = ea.boundary_mode == 0 ? (own & 0x7FFFu) != 0 : (own >> 15) != 0;
        if (own > VF_VOXEL_FREE && isB && ea.noise[ni] < ea.prob) {
            const int visited = vxy * (1 + (zz > 0) + (zz < d.Z - 1));
            if (ea.erodes[visited] >> count & 1u) {
                note_eroded(ea, cell);
                return VF_VOXEL_EMPTY;
            }
        }
        return own;
    };
    *reinterpret_cast<unsigned*>(dst + gi) = eroded(ownA, countA, gz, ia, gi) | eroded(ownB, countB, gz + 1, ib, gi + 1) << 16;
}

// 8 cells of a z-row of `dst`: one 128-bit store where rows are 16-byte aligned (Z % 8 == 0), otherwise two 64-bit halves, the second
// of which lies outside the grid in the last chunk of a row when Z % 8 == 4
template <bool ALIGNED16>
__device__ __forceinline__ void store_chunk(uint16_t* __restrict__ p, const uint4& v, int gz, int Z)
{
    if (ALIGNED16) {
        *reinterpret_cast<uint4*>(p) = v;
    } else {
        *reinterpret_cast<uint2*>(p) = make_uint2(v.x, v.y);
        if (gz + 4 < Z) *reinterpret_cast<uint2*>(p + 4) = make_uint2(v.z, v.w);
    }
}

// TMA: the tile is staged by one tensor-map box load, which needs 16-byte row pitches (Z % 8 == 0).  !TMA: grids whose Z is a multiple
// of 4 only (the reference's dataset dims rule rounds x and z to multiples of 4, CADScene.cpp:272-273) stage the same box with 64-bit
// loads; everything after the staging is shared.
template <int OP, bool TMA>
__global__ void __launch_bounds__(256, 8) stencil_fast_kernel(const __grid_constant__ CUtensorMap src_map, const uint16_t* src, uint16_t* dst, Dims d, ErodeArgs ea,
                                                           int uniform_erodes)  // src == dst for the in-place detect pass: no __restrict__
{
    __shared__ __align__(128) uint16_t s[FROWS * FRS];
    __shared__ __align__(8) unsigned long long tma_bar;
    __shared__ uint32_t summ[FROWS * 8];
    __shared__ uint16_t tasks[SX * SYT * 8];
    __shared__ int ntasks;
    const int gx0 = blockIdx.z * SX, gy0 = blockIdx.y * SYT, gz0 = blockIdx.x * SZT;  // 3-D launch: no index divisions
    const int t = threadIdx.x;
    if (t == 0) ntasks = 0;

    // ---- stage: one TMA box of 10 x 10 rows x 80 cells (z = -8 .. 71 of the tile: 160-byte rows keep every 8-cell chunk 16-byte
    //      aligned and bring the z halo cells -1 and 64 along).  The copy engine fills cells outside the grid with zeros = kOutside,
    //      so staging costs the CTA one instruction instead of a thousand address computations.
    const int ch = t & 7;
    const int gzc = gz0 + ch * 8;
    if (!TMA) {
        // 100 rows x 20 half-chunks of 4 cells; Z % 4 == 0 puts every half wholly inside or wholly outside the grid
        for (int h = t; h < FROWS * (FRS / 4); h += 256) {
            const int row = h / (FRS / 4), hh = h - row * (FRS / 4);
            const int gx = gx0 + row / HY - 1, gy = gy0 + row % HY - 1, gz = gz0 - 8 + 4 * hh;
            uint2 v = make_uint2(0u, 0u);  // kOutside
            if ((unsigned)gx < (unsigned)d.X && (unsigned)gy < (unsigned)d.Y && gz >= 0 && gz < d.Z)
                v = __ldcg(reinterpret_cast<const uint2*>(src + ((size_t)((unsigned)gx * (unsigned)d.Y + gy)) * d.Z + gz));  // coherent: detect runs in place
            *reinterpret_cast<uint2*>(&s[row * FRS + 4 * hh]) = v;
        }
        __syncthreads();
    } else {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&tma_bar);
        if (t == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(FROWS * FRS * 2)) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(s)),
                         "l"(reinterpret_cast<unsigned long long>(&src_map)), "r"(bar), "r"(gz0 - 8), "r"(gy0 - 1), "r"(gx0 - 1)
                         : "memory");
        }
        __syncthreads();  // the barrier is initialised for everybody
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    }

    // ---- chunk summaries; along the way: does the whole staged region hold one value?
    const unsigned ref = s[f_at(0, 0, 0)];
    bool all_ref = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int row = (t >> 3) + 32 * k;
        if (row < FROWS) {
            const uint16_t* p = &s[row * FRS + 8 + ch * 8];
            const uint4 v = *reinterpret_cast<const uint4*>(p);
            const unsigned v0 = v.x & 0xFFFFu;
            const bool u8 = v.x == v0 * 0x10001u && v.y == v.x && v.z == v.x && v.w == v.x;
            const bool uz = u8 && p[-1] == v0 && p[8] == v0;
            summ[row * 8 + ch] = v0 | (uz ? 0x10000u : 0u);
            all_ref = all_ref && uz && v0 == ref;
        }
    }
    if (__syncthreads_and(all_ref) && !(OP == OP_ERODE3 && uniform_erodes)) {
        // one value everywhere, halo included: all three stencils are the identity on this tile
        if (OP != OP_DETECT) {
            const uint4 fill = make_uint4(ref * 0x10001u, ref * 0x10001u, ref * 0x10001u, ref * 0x10001u);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int y = (t >> 3) & 7, x = (t >> 6) + 4 * k;
                const int gx = gx0 + x, gy = gy0 + y;
                if (gx < d.X && gy < d.Y && gzc < d.Z) store_chunk<TMA>(dst + ((size_t)((unsigned)gx * (unsigned)d.Y + gy)) * d.Z + gzc, fill, gzc, d.Z);
            }
        }
        return;
    }

    // ---- interior chunks: uniform window -> identity (vector copy / nothing), otherwise queue
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int y = (t >> 3) & 7, x = (t >> 6) + 4 * k;
        const int gx = gx0 + x, gy = gy0 + y;
        if (gx >= d.X || gy >= d.Y || gzc >= d.Z) continue;
        const uint32_t* sp = &summ[(x * HY + y) * 8 + ch];  // summary of staged row (x-1, y-1)
        const unsigned su = sp[(HY + 1) * 8];
        unsigned diff = ~su & 0x10000u;  // window not uniform along z
#pragma unroll
        for (int dx = 0; dx <= 2; ++dx)
#pragma unroll
            for (int dy = 0; dy <= 2; ++dy) diff |= sp[(dx * HY + dy) * 8] ^ su;
        if (diff == 0 && !(OP == OP_ERODE3 && uniform_erodes)) {
            if (OP != OP_DETECT)
                store_chunk<TMA>(dst + ((size_t)((unsigned)gx * (unsigned)d.Y + gy)) * d.Z + gzc, *reinterpret_cast<const uint4*>(&s[f_at(x, y, ch * 8)]), gzc, d.Z);
        } else {
            tasks[atomicAdd(&ntasks, 1)] = (uint16_t)((x << 6) | (y << 3) | ch);
        }
    }
    __syncthreads();

    // ---- queued chunks: exact stencil, one voxel pair per thread per round (balanced across the CTA)
    const int nq = ntasks;
    for (int q = t; q < nq * 4; q += 256) {
        const int c = tasks[q >> 2], k = q & 3;
        const int zc = c & 7, y = (c >> 3) & 7, x = c >> 6, z = zc * 8 + k * 2;
        const int gx = gx0 + x, gy = gy0 + y, gz = gz0 + z;
        if (!TMA && gz >= d.Z) continue;  // second half of a row's last chunk when Z % 8 == 4
        const size_t gi = ((size_t)gx * d.Y + gy) * d.Z + gz;
        stencil_pair<OP>(s, d, x, y, z, gx, gy, gz, gi, ea, dst);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

vf_status make_grid_map(CUtensorMap* map, const uint16_t* src, const Dims& d)
{
    static EncodeTiledFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    VF_REQUIRE(encode != nullptr, VF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = { (cuuint64_t)d.Z, (cuuint64_t)d.Y, (cuuint64_t)d.X };
    const cuuint64_t strides[2] = { (cuuint64_t)d.Z * 2, (cuuint64_t)d.Z * d.Y * 2 };  // bytes; multiples of 16 because Z % 8 == 0
    const cuuint32_t box[3] = { FRS, HY, HX }, estr[3] = { 1, 1, 1 };
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t*>(src), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VF_REQUIRE(r == CUDA_SUCCESS, VF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%dx%d grid", (int)r, d.X, d.Y, d.Z);
    return VF_OK;
}

vf_status launch_stencil(vf_grid* g, int op, const uint16_t* src, uint16_t* dst, const ErodeArgs& ea)
{
    vf_ctx* c = g->ctx;
    Dims d = { (int)g->X, (int)g->Y, (int)g->Z };
    const int ntx = (d.X + SX - 1) / SX, nty = (d.Y + SYT - 1) / SYT, ntz = (d.Z + SZT - 1) / SZT;
    const int blocks = ntx * nty * ntz;
    const bool tma_ok = d.Z % 8 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    if ((tma_ok || (d.Z % 4 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 7) == 0)) && nty <= 65535 && ntx <= 65535) {
        // would a voxel with a completely uniform neighbourhood erode?  (count = mask population, visited = 27; true only for
        // unusual thresholds, e.g. CROSS with threshold > 0.78).  Evaluated with the kernel's float32 expression.
        int uniform_erodes = 0;
        if (op == OP_ERODE3) {
            const float activation = (float)__builtin_popcount(ea.maskbits) / 27.0f;
            uniform_erodes = activation < ea.activations * ea.thr;
        }
        // tensor map over the source grid as (Z, Y, X) uint16 with a box of (FRS, HY, HX) cells
        CUtensorMap map = {};
        if (tma_ok) VF_TRY(make_grid_map(&map, src, d));
        const dim3 grid3(ntz, nty, ntx);
        if (tma_ok) {
            switch (op) {
            case OP_DETECT: stencil_fast_kernel<OP_DETECT, true><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, 0); break;
            case OP_ERODE3: stencil_fast_kernel<OP_ERODE3, true><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, uniform_erodes); break;
            default: stencil_fast_kernel<OP_SWEEP, true><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, 0); break;
            }
        } else {
            switch (op) {
            case OP_DETECT: stencil_fast_kernel<OP_DETECT, false><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, 0); break;
            case OP_ERODE3: stencil_fast_kernel<OP_ERODE3, false><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, uniform_erodes); break;
            default: stencil_fast_kernel<OP_SWEEP, false><<<grid3, 256, 0, c->stream>>>(map, src, dst, d, ea, 0); break;
            }
        }
        VF_LAUNCHED(c);
        return VF_OK;
    }
    switch (op) {
    case OP_DETECT: stencil_kernel<OP_DETECT><<<blocks, 256, 0, c->stream>>>(src, dst, d, ntx, nty, ntz, ea); break;
    case OP_ERODE3: stencil_kernel<OP_ERODE3><<<blocks, 256, 0, c->stream>>>(src, dst, d, ntx, nty, ntz, ea); break;
    default: stencil_kernel<OP_SWEEP><<<blocks, 256, 0, c->stream>>>(src, dst, d, ntx, nty, ntz, ea); break;
    }
    VF_LAUNCHED(c);
    return VF_OK;
}

// RegularGrid.cpp:84-122 — erosion mask and the "activations" normalisation quirks, float32 as written
void build_mask(int type, uint32_t& size, std::vector<float>& mask, float& activations)
{
    if (!(size % 2)) ++size;
    const uint32_t k = size, maskSize = k * k * k, cc = (uint32_t)floorf(k / 2.0f);
    mask.assign(maskSize, 0.0f);
    activations = 0;
    if (type == VF_SQUARE) {
        std::fill(mask.begin(), mask.end(), 1.0f);
        activations = (float)maskSize;
    } else if (type == VF_CROSS) {
        for (uint32_t x = 0; x < k; ++x) mask[x * k * k + cc * k + cc] = 1.0f;
        for (uint32_t y = 0; y < k; ++y) mask[cc * k * k + y * k + cc] = 1.0f;
        for (uint32_t z = 0; z < k; ++z) mask[cc * k * k + cc * k + z] = 1.0f;
        activations = 1.0f / 3.0f * maskSize;
    } else if (type == VF_ELLIPSE) {
        for (uint32_t x = 0; x < k; ++x)
            for (uint32_t y = 0; y < k; ++y)
                for (uint32_t z = 0; z < k; ++z) {
                    const float dx = (float)x - (float)cc, dy = (float)y - (float)cc, dz = (float)z - (float)cc;
                    if (sqrtf(dx * dx + dy * dy + dz * dz) < (float)cc + 1.1920929e-07f) {  // glm::epsilon<float>()
                        mask[x * k * k + y * k + z] = 1.0f;
                        ++activations;
                    }
                }
    }
    activations /= maskSize;
}

}  // namespace

extern "C" vf_status vf_detect_boundaries(vf_grid* g, int boundary_size)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    VF_REQUIRE(boundary_size >= 0 && boundary_size <= 64, VF_ERR_INVALID_ARGUMENT, "detectBoundaries: boundarySize %d out of range [0, 64]", boundary_size);
    if (boundary_size != 1) {  // the reference only ever passes 1 (CADScene.cpp:687, RegularGrid.cpp:135); other sizes take the direct path
        Dims d = { (int)g->X, (int)g->Y, (int)g->Z };
        detect_generic_kernel<<<g->ctx->num_sms * 8, 256, 0, g->ctx->stream>>>(g->d, d, boundary_size);
        VF_LAUNCHED(g->ctx);
        return VF_OK;
    }
    ErodeArgs ea = {};
    return launch_stencil(g, OP_DETECT, g->d, g->d, ea);
}

extern "C" vf_status vf_remove_isolated_regions_grid(vf_grid* g)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    VF_TRY(vf_scratch_reserve(c, c->grid2, g->n() * 2));
    ErodeArgs ea = {};
    VF_TRY(launch_stencil(g, OP_SWEEP, g->d, (uint16_t*)c->grid2.ptr, ea));
    VF_CUDA(cudaMemcpyAsync(g->d, c->grid2.ptr, g->n() * 2, cudaMemcpyDeviceToDevice, c->stream));
    return VF_OK;
}

// skip_detect / skip_sweep: the caller runs detectBoundaries / the final sweep itself (a slab of a larger grid exchanges its halo planes between
// the passes); cell_offset: index of the grid's first cell in the grid the noise table is indexed by.
static vf_status erode_impl(vf_grid* g, int type, uint32_t size, uint32_t iterations, float prob, float thr, const float* noise, uint32_t nnoise,
                            int boundary_mode, bool skip_detect, bool skip_sweep, unsigned long long cell_offset)
{
    VF_REQUIRE(g != nullptr && noise != nullptr && nnoise > 0, VF_ERR_INVALID_ARGUMENT, "erode: null grid or noise table");
    VF_REQUIRE(type >= 0 && type <= 2, VF_ERR_INVALID_ARGUMENT, "erode: bad convolution type %d", type);
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    std::vector<float> mask;
    float activations;
    build_mask(type, size, mask, activations);
    const size_t n = g->n();
    VF_TRY(vf_scratch_reserve(c, c->grid2, n * 2));
    VF_TRY(vf_scratch_reserve(c, c->noise, (size_t)nnoise * 4 + mask.size() * 4 + 256));
    float* d_noise = (float*)c->noise.ptr;
    float* d_mask = d_noise + nnoise;
    if (size != 3)  // the 3^3 kernels take the mask as 27 bits; only the generic kernel reads the float mask (pageable: staged before return)
        VF_CUDA(cudaMemcpyAsync(d_mask, mask.data(), mask.size() * 4, cudaMemcpyHostToDevice, c->stream));
    ErodeArgs ea = {};
    ea.noise = d_noise, ea.nnoise = nnoise, ea.prob = prob, ea.thr = thr, ea.activations = activations, ea.boundary_mode = boundary_mode;
    for (int visited = 1; visited <= 27; ++visited)
        for (int count = 0; count <= 27; ++count) {
            const volatile float activation = (float)count / (float)visited;  // float32, as erodeGrid-comp.glsl:55
            const volatile float limit = activations * thr;
            if (activation < limit) ea.erodes[visited] |= 1u << count;
        }
    ea.nn_magic = ~0ull / nnoise + 1;
    ea.idx32 = cell_offset + n <= 0xFFFFFFFFull;
    ea.cell_offset = cell_offset;
    if (size == 3)
        for (int i = 0; i < 27; ++i)
            if (mask[i] != 0.0f) ea.maskbits |= 1u << i;
    Dims d = { (int)g->X, (int)g->Y, (int)g->Z };
    uint16_t* a = g->d;
    uint16_t* b = (uint16_t*)c->grid2.ptr;
    vf_grid view = *g;  // shallow view used to run passes on either buffer

    // Iterations 2..K of the 3^3 erosion run sparsely (erode_sparse_kernel): every pass marks the cells it empties in a bitmap (one bit per
    // cell, two bitmaps in the key arena), the next one only re-decides the 27 cells around each of them.  cfg3 (512^3, 64 regions): the first
    // pass empties ~3 * 10^4 cells, the second ~4 * 10^3, the third ~7 * 10^2.
    const bool sparse_ok = size == 3 && iterations >= 2 && n <= 0xFFFFFFFFull;
    // The final 3^3 sweep runs sparsely too when the first pass takes the tiled kernel: that pass also decides the sweep for ITS input (one more
    // 27-cell count per evaluated pair, `s0kill`), a verdict that holds wherever no erosion pass touches the 3^3 window; afterwards only the cells
    // around eroded cells (`ever`) are re-decided on the final erosion output, and the marked cells are emptied in the caller's grid — instead
    // of one more full pass over the grid (0.20 ms of the 0.68 ms erode stage at 512^3).
    const int ntx_ = (d.X + SX - 1) / SX, nty_ = (d.Y + SYT - 1) / SYT;
    const bool tiled = ((d.Z % 8 == 0 && (((uintptr_t)a | (uintptr_t)b) & 15) == 0) || (d.Z % 4 == 0 && (((uintptr_t)a | (uintptr_t)b) & 7) == 0)) && nty_ <= 65535 &&
                       ntx_ <= 65535;
    const bool s0_ok = size == 3 && iterations >= 1 && n <= 0xFFFFFFFFull && tiled && !skip_sweep;
    const uint32_t nwords = (uint32_t)((n + 31) / 32);
    constexpr uint32_t kMaxTracked = 60;  // per-pass counters in the header; longer calls take full passes
    const bool track = (sparse_ok || s0_ok) && iterations <= kMaxTracked;
    uint32_t* lists[2] = { nullptr, nullptr };
    uint32_t *ever = nullptr, *s0kill = nullptr, *counters = nullptr, *ever_list = nullptr;
    if (track) {
        // key arena: ever | s0kill (one bit per cell each) | 64 counters (per pass, [63] = the call's) | two per-pass lists | the call's list
        VF_TRY(vf_scratch_reserve(c, c->keys, (2 * (size_t)nwords + 64 + 3 * (size_t)kListCap) * 4));
        ever = (uint32_t*)c->keys.ptr, s0kill = ever + nwords, counters = s0kill + nwords;
        lists[0] = counters + 64, lists[1] = lists[0] + kListCap, ever_list = lists[1] + kListCap;
        VF_TRY(vf_k_zero(c, ever, (2 * (size_t)nwords + 64) * 4));  // one launch for everything this call tracks
    }
    for (uint32_t it = 0; it < iterations; ++it) {
        view.d = a;
        const int out = (int)(it & 1u), in = out ^ 1;  // list written / read by this pass (the counters are per pass: nothing is reset mid-call)
        ea.ever = track ? ever : nullptr;
        ea.pass_list = track ? lists[out] : nullptr, ea.pass_count = track ? counters + it : nullptr;
        ea.ever_list = track && s0_ok ? ever_list : nullptr, ea.ever_count = track && s0_ok ? counters + 63 : nullptr;
        ea.s0kill = (track && s0_ok && it == 0) ? s0kill : nullptr;
        if (it == 0) {
            if (!skip_detect) VF_TRY(launch_stencil(&view, OP_DETECT, a, a, ea));  // only in the first iteration: see below
            // The noise table is handled AFTER the detect pass is queued: comparing a 4 MB table with its shadow takes the host longer than a
            // kernel launch, and the GPU would sit idle meanwhile.  The table travels on the stream: from pageable memory the call returns once the
            // driver has staged it; a pinned table must stay valid until the context is synchronised (documented in voxfrag.h).  A table identical
            // to the one already on the device (compared against a host shadow) is not sent again.
            if (c->noise_shadow.size() != nnoise || std::memcmp(c->noise_shadow.data(), noise, (size_t)nnoise * 4) != 0) {
                c->noise_shadow.clear();
                VF_CUDA(cudaMemcpyAsync(d_noise, noise, (size_t)nnoise * 4, cudaMemcpyHostToDevice, c->stream));
                c->noise_shadow.assign(noise, noise + nnoise);
            }
        }
        // RegularGrid.cpp:135 calls detectBoundaries(1) in every iteration; only the first call can change the grid.  The pass tags an
        // untagged labelled cell iff its clamped 3^3 box holds another label (> FREE, tag ignored) and never clears a tag; erosion copies
        // words, tags included, or writes EMPTY.  After the first pass every cell whose box holds another label is tagged, and a later
        // pass could only tag a cell whose box GAINED another label — no pass creates labels.  (Checked against the literal shader
        // transcription with and without the repeated passes: tests/test_oracle_literal_shaders.py.)
        if (it > 0 && sparse_ok && track) {
            erode_sparse_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(a, b, d, ea, ever, nwords, lists[in], counters + it - 1);
            VF_LAUNCHED(c);
        } else if (size == 3) {
            VF_TRY(launch_stencil(&view, OP_ERODE3, a, b, ea));
        } else {
            erode_generic_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(a, b, d, d_mask, (int)size, ea);
            VF_LAUNCHED(c);
        }
        std::swap(a, b);  // replaces copyGrid (:149-152): the eroded grid becomes the current one
    }
    view.d = a;
    if (track && s0_ok) {  // RegularGrid.cpp:155, sparsely: `a` holds the final erosion output, the other buffer the output of the pass before
        sweep_sparse_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(a, d, ever, s0kill, nwords, ever_list, counters + 63);
        VF_LAUNCHED(c);
        const bool stale = a != g->d;  // the caller's grid misses the last pass
        sweep_apply_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>(g->d, s0kill, ever, stale ? lists[(iterations - 1) & 1u] : nullptr,
                                                                  stale ? counters + iterations - 1 : nullptr, nwords, n);
        VF_LAUNCHED(c);
        return VF_OK;
    }
    if (skip_sweep) {
        if (a != g->d) VF_CUDA(cudaMemcpyAsync(g->d, a, n * 2, cudaMemcpyDeviceToDevice, c->stream));
        return VF_OK;
    }
    VF_TRY(launch_stencil(&view, OP_SWEEP, a, b, ea));  // RegularGrid.cpp:155
    if (b != g->d) VF_CUDA(cudaMemcpyAsync(g->d, b, n * 2, cudaMemcpyDeviceToDevice, c->stream));
    return VF_OK;
}

extern "C" vf_status vf_erode(vf_grid* g, int type, uint32_t size, uint32_t iterations, float prob, float thr, const float* noise, uint32_t nnoise,
                              int boundary_mode)
{
    return erode_impl(g, type, size, iterations, prob, thr, noise, nnoise, boundary_mode, false, false, 0);
}

// ONE erosion pass (erodeGrid-comp.glsl + copyGrid, RegularGrid.cpp:137-152) without detectBoundaries before it and without the sweep after it,
// for a slab of a grid cut along x (multi-GPU): the caller tags the boundaries once (vf_detect_boundaries), exchanges the halo
// planes, runs this pass + an exchange per iteration, then vf_remove_isolated_regions_grid.  cell_offset = (first plane of the slab, halo
// included) * Y * Z: the noise is indexed by the cell's position in the whole grid.
extern "C" vf_status vf_erode_pass(vf_grid* g, int type, uint32_t size, float prob, float thr, const float* noise, uint32_t nnoise, int boundary_mode,
                                   uint64_t cell_offset)
{
    return erode_impl(g, type, size, 1, prob, thr, noise, nnoise, boundary_mode, true, true, cell_offset);
}

static vf_status histogram_impl(vf_grid* g, uint32_t* counts, uint64_t* occupied, bool unmask)
{
    VF_REQUIRE(g != nullptr && counts != nullptr, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    // device side, after the seed area of the small arena: { occupied (8 B), highest non-empty bin (4 B), pad } | bins.  The bins are zero on
    // entry: the first call on this arena zeroes all of them, every call re-zeroes what it used AFTER its read-back, off the caller's path.
    char* d_head = (char*)c->small.ptr + (512 << 10);
    unsigned long long* d_occ = (unsigned long long*)d_head;
    uint32_t* d_max = (uint32_t*)(d_head + 8);
    uint32_t* d_counts = (uint32_t*)(d_head + 16);
    if (c->hist_clean != d_head) VF_TRY(vf_k_zero(c, d_head, 16 + VF_HISTOGRAM_BINS * 4));
    c->hist_clean = nullptr;
    // 38-40 registers: six CTAs are resident per SM, so the grid is one full wave (measured equal to eight per SM within noise)
    const int blocks = (int)std::min((size_t)c->num_sms * 6, (g->n() / 8 + 255) / 256 + 1);
    if (unmask) histogram_kernel<true><<<blocks, 256, 0, c->stream>>>(g->d, g->n(), d_counts, d_occ, d_max);
    else histogram_kernel<false><<<blocks, 256, 0, c->stream>>>(g->d, g->n(), d_counts, d_occ, d_max);
    VF_LAUNCHED(c);
    // one copy of the head + the first kHead bins into the context's pinned read-back area (a copy into the caller's pageable buffer is staged
    // by the runtime and costs two extra synchronisations); labels beyond kHead (more than a thousand fragments) cost a second copy
    constexpr uint32_t kHead = 1024;
    char* h = (char*)c->pinned + (1 << 17);
    VF_CUDA(cudaMemcpyAsync(h, d_head, 16 + kHead * 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    const uint32_t top = std::min(*(const uint32_t*)(h + 8), (uint32_t)VF_HISTOGRAM_BINS - 1);
    if (top >= kHead) {
        VF_CUDA(cudaMemcpyAsync(h + 16 + kHead * 4, d_counts + kHead, (size_t)(top + 1 - kHead) * 4, cudaMemcpyDeviceToHost, c->stream));
        VF_CUDA(vf_sync(c));
    }
    const uint32_t have = std::max(top + 1, kHead);
    std::memcpy(counts, h + 16, (size_t)have * 4);
    std::memset(counts + have, 0, (size_t)(VF_HISTOGRAM_BINS - have) * 4);
    if (occupied) std::memcpy(occupied, h, 8);
    if (vf_k_zero(c, d_head, 16 + (size_t)(top + 1) * 4) == VF_OK) c->hist_clean = d_head;  // enqueued, not waited for
    return VF_OK;
}

extern "C" vf_status vf_histogram(vf_grid* g, uint32_t* counts, uint64_t* occupied) { return histogram_impl(g, counts, occupied, false); }
extern "C" vf_status vf_histogram_undo_mask(vf_grid* g, uint32_t* counts, uint64_t* occupied) { return histogram_impl(g, counts, occupied, true); }
