// flood.cu — F2/F3 flood-fill fragmentation (the connected-component step of F3 and C1 live in ccl.cu).
//
// Replaces FloodFracturer::build (SRC/Fracturer/FloodFracturer.cpp:98-191; shaders floodFracturer-comp.glsl:24-64,
// disjointSet-comp.glsl:17-24, disjointSetStack-comp.glsl:20-37, undoMask-comp.glsl) and the intended semantics of
// NaiveFracturer::removeIsolatedRegions (removeIsolatedRegionsCPU, NaiveFracturer.cpp:111-150).
//
// Deterministic rule (the reference's plain-store claims are racy): every FREE cell takes the word of the
// source minimising (geodesic distance through non-EMPTY cells under the 6/26 neighbourhood, index in the seeds vector).
// That is the unique least fixed point of  key(v) = min(key(v), min_{u ~ v} key(u) + 1 level)  over 32-bit keys
// (dist << 15 | order), which makes the result independent of the evaluation schedule: tiles can be relaxed in any order,
// concurrently, on any number of GPUs, and still give bit-identical labels.
//
// Schedule: worklist rounds over 16 x 16 x 32 tiles (tiles.cuh).  A round only lowers cells to distance levels below
// lo + kLevelsPerRound, where lo is the lowest level a candidate was deferred at in the previous round (delta-stepping with
// unit weights): fronts from different seeds then reach a cell in distance order, so almost every cell is lowered once, instead
// of being claimed by whichever front arrives first in tile order and corrected later.  Deferred candidates stay marked in a
// per-tile mask and the tile re-enqueues itself.
//
// Thin fronts first: on voxelized surfaces (shells a few cells thick — what the reference floods) a BFS level holds a few thousand cells,
// far too few to keep tiles busy, and a flood is a chain of hundreds of dependent levels.  A phase therefore starts at cell granularity
// (flood_front_kernel: ONE thread-block cluster, the front as lists of (cell, key) pairs in shared memory, the same keys lowered by
// atomicMin, a cluster barrier every 16 steps) and hands over to the tile worklist as soon as more than vf_ctx_set_flood_front pairs are
// pending (solid interiors); the keys it leaves are upper bounds realised by real paths, which is all the tiles' relaxation needs.
//
// F3 (extra seeds, words = fragId | prefix << 8): the reference's in-flood prefix merge converges, inside every connected set
// of equal-fragId cells, to the lowest prefix present; the disjoint step then frees every cell whose prefix is not its
// fragment's minimum and re-floods from all labelled cells.  Restated: keep, per fragment id, only the component (under the
// flood neighbourhood, through equal-fragId cells) that contains the fragment's lowest-prefix source; free the rest; re-flood
// with order(cell) = index of that source.  One such round reaches the reference loop's fixed point (numDisjointVoxels == 0).
#include <cooperative_groups.h>
#include <cuda.h>  // CUtensorMap (types only; the encoder comes through cudaGetDriverEntryPoint)

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nccl_dyn.h"
#include "tiles.cuh"

using namespace vft;

namespace {

constexpr uint32_t KEY_WALL = 0xFFFFFFFFu, KEY_UNREACHED = 0xFFFFFFFEu;
constexpr int KEY_SHIFT = 15;
constexpr uint32_t KEY_LEVEL = 1u << KEY_SHIFT;
constexpr uint32_t KEY_LIMIT = 0xFFFF0000u;  // keys at or above this cannot take another level (dist >= 2^17 - 2)
constexpr uint32_t kLevelsPerRound = 16;     // width of a round's distance window (one tile edge)
constexpr uint32_t kNoLevel = 0xFFFFFFFFu;
constexpr unsigned kOwnRowMax = 2;           // rows with at most this many candidates in a step are relaxed by their owner lane

enum { ST_VISITS = 0, ST_ROUNDS = 1, ST_ERROR = 2, ST_FREED = 3, ST_MAXDIST = 4, ST_CHANGED = 5, ST_STEPS = 6, ST_MAXSTEPS = 7 };

#ifdef VF_FLOOD_TIMING  // tools/ only (VF_NVCC_EXTRA=-DVF_FLOOD_TIMING): cycles per phase of a tile visit, summed over visits
__device__ unsigned long long g_flood_cycles[8];  // 0 load, 1 masks, 2 relaxation, 3 write-back + wake, 4 visits, 5 steps
__device__ unsigned long long g_front_cycles[12];  // thin-front solver, thread 0 of CTA 0: 0 level set-up, 1 key loads, 2 claims, 3 pushes, 4 end of level, 5 barrier, 7 levels, 8 CTAs, 9 cells, 10 passes over the own list
#define VF_TICK(slot)                                                   \
    do {                                                                \
        if (threadIdx.x == 0) {                                         \
            const long long now__ = clock64();                          \
            atomicAdd(&g_flood_cycles[slot], (unsigned long long)(now__ - tick__)); \
            tick__ = now__;                                             \
        }                                                               \
    } while (0)
#else
#define VF_TICK(slot) do { } while (0)
#endif
#ifdef VF_FLOOD_TIMING
// thread 0 of CTA 0 accumulates in shared memory (no global traffic of its own); `dep` makes the clock read wait for the value it names
#define VF_FTICK(slot, dep)                                                                  \
    do {                                                                                     \
        if (gtid == 0) {                                                                     \
            long long now__;                                                                 \
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(now__) : "r"((uint32_t)(dep)) : "memory"); \
            s_ft[slot] += (unsigned long long)(now__ - ftick__);                             \
            ftick__ = now__;                                                                 \
        }                                                                                    \
    } while (0)
#else
#define VF_FTICK(slot, dep) do { } while (0)
#endif

constexpr int kRoundWord = 14;  // worklist header layout: stats[8] count[3] lo[3] | word 14: device-resident round id | word 15: steps run by the thin-front solver
constexpr int kFrontWord = 15;
constexpr int kFrontThreads = 512;             // threads per CTA of the thin-front solver
constexpr int kFrontLocal = 4096;              // cells of a level one CTA can hold (two lists of this size in shared memory)
constexpr uint32_t kFrontCap = (uint32_t)kVfFrontCap;  // entries per front list
constexpr int kFrontSublevels = 16;            // steps a CTA of the thin-front solver runs on its own list between two cluster barriers
constexpr uint32_t kFrontMaxLevel = 131000;    // the key field holds 17 bits of distance: beyond this the tiles take over (and report the overflow)
constexpr size_t kSmemBytes = (size_t)(kCells + 5 * kThreads + 8 + 2) * sizeof(uint32_t);  // tile | 5 row-mask arrays | misc | mbarrier

// ------------------------------------------------------------------------------------------------ key field set-up
// phase 1: homogenize (FloodFracturer.cpp:99) folded in: EMPTY -> WALL, anything else -> UNREACHED.
// phase 2: surviving words -> (0, order of their fragment's principal source); FREE -> UNREACHED.
template <bool PHASE2>
__global__ void __launch_bounds__(256) flood_init_keys_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ keys, TileGeom g,
                                                              uint8_t* __restrict__ occ, const uint16_t* __restrict__ order_of_frag)
{
    const size_t n = (size_t)g.X * g.Y * g.Z;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t v = grid[i];
        uint32_t k;
        if (v == VF_VOXEL_EMPTY) k = KEY_WALL;
        else if (!PHASE2 || v == VF_VOXEL_FREE) k = KEY_UNREACHED;
        else k = order_of_frag[v & 0xFFu];
        keys[i] = k;
        if (v != VF_VOXEL_EMPTY) {
            const int z = (int)(i % g.Z);
            const size_t r = i / g.Z;
            const int y = (int)(r % g.Y), x = (int)(r / g.Y);
            const uint32_t tile = ((uint32_t)(x / TX) * g.nty + y / TY) * g.ntz + z / TZ;
            if (!occ[tile]) occ[tile] = 1;
        }
    }
}

// The same, CH = 8 or 4 voxels per thread (Z % CH == 0): one 128- or 64-bit label load, 128-bit key stores, no per-voxel index
// arithmetic.  A warp walks the chunks of one z-row, so (x, y) are decomposed once per row; a chunk never straddles a tile (CH | TZ).
template <bool PHASE2, int CH>
__global__ void __launch_bounds__(256) flood_init_keys_vec_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ keys, TileGeom g,
                                                                  uint8_t* __restrict__ occ, const uint16_t* __restrict__ order_of_frag)
{
    const int lane = threadIdx.x & 31, cpr = g.Z / CH;  // chunks per row
    const size_t rows = (size_t)g.X * g.Y, nwarps = (size_t)gridDim.x * (blockDim.x / 32);
    for (size_t row = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; row < rows; row += nwarps) {
        const int y = (int)(row % g.Y), x = (int)(row / g.Y);
        const uint32_t tile_row = ((uint32_t)(x / TX) * g.nty + y / TY) * g.ntz;
        for (int ch = lane; ch < cpr; ch += 32) {
            const size_t i = row * g.Z + (size_t)ch * CH;
            uint32_t w[4] = { 0, 0, 0, 0 };
            if (CH == 8) {
                const uint4 v = vf_ldg_stream(reinterpret_cast<const uint4*>(grid + i));
                w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
            } else {
                const uint2 v = vf_ldg_stream(reinterpret_cast<const uint2*>(grid + i));
                w[0] = v.x, w[1] = v.y;
            }
            uint32_t k[8];
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint32_t lab = (w[c >> 1] >> ((c & 1) * 16)) & 0xFFFFu;
                k[c] = lab == VF_VOXEL_EMPTY ? KEY_WALL : (!PHASE2 || lab == VF_VOXEL_FREE) ? KEY_UNREACHED : (uint32_t)order_of_frag[lab & 0xFFu];
            }
            vf_stg_stream(reinterpret_cast<uint4*>(keys + i), make_uint4(k[0], k[1], k[2], k[3]));
            if (CH == 8) vf_stg_stream(reinterpret_cast<uint4*>(keys + i + 4), make_uint4(k[4], k[5], k[6], k[7]));
            if ((w[0] | w[1] | w[2] | w[3]) != 0) {
                const uint32_t tile = tile_row + (uint32_t)(ch * CH) / TZ;
                if (!occ[tile]) occ[tile] = 1;
            }
        }
    }
}

template <bool PHASE2>
vf_status launch_init_keys(vf_ctx* c, const uint16_t* grid, uint32_t* keys, const TileGeom& g, uint8_t* occ, const uint16_t* order, int blocks)
{
    if (g.Z % 8 == 0 && (((uintptr_t)grid | (uintptr_t)keys) & 15) == 0)
        flood_init_keys_vec_kernel<PHASE2, 8><<<blocks, 256, 0, c->stream>>>(grid, keys, g, occ, order);
    else if (g.Z % 4 == 0 && ((uintptr_t)grid & 7) == 0 && ((uintptr_t)keys & 15) == 0)  // the reference's dataset dims: x and z multiples of 4
        flood_init_keys_vec_kernel<PHASE2, 4><<<blocks, 256, 0, c->stream>>>(grid, keys, g, occ, order);
    else
        flood_init_keys_kernel<PHASE2><<<blocks, 256, 0, c->stream>>>(grid, keys, g, occ, order);
    VF_LAUNCHED(c);
    return VF_OK;
}

// FloodFracturer.cpp:102-103,127-132: seeds written in order (a later seed on the same cell overwrites), their cells pushed.
__global__ void flood_seed_kernel(uint32_t* __restrict__ keys, TileGeom g, Worklist wl, const ushort4* __restrict__ seeds, int S, uint32_t round)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    wl.lo[round % 3] = 0, wl.lo[(round + 1) % 3] = kNoLevel;  // the first round's window starts at level 0
    for (int s = 0; s < S; ++s) {
        const ushort4 sd = seeds[s];
        keys[((size_t)sd.x * g.Y + sd.y) * g.Z + sd.z] = (uint32_t)s;  // dist 0, order s
        const uint32_t tile = ((uint32_t)(sd.x / TX) * g.nty + sd.y / TY) * g.ntz + sd.z / TZ;
        wl.occ[tile] = 1;
        enqueue_tile(wl, tile, round);
    }
}

// slab mode: the seed's order in the GLOBAL seed list travels in .w
__global__ void flood_seed_order_kernel(uint32_t* __restrict__ keys, TileGeom g, Worklist wl, const ushort4* __restrict__ seeds, int S, uint32_t round)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    wl.lo[round % 3] = 0, wl.lo[(round + 1) % 3] = kNoLevel;
    for (int s = 0; s < S; ++s) {
        const ushort4 sd = seeds[s];
        keys[((size_t)sd.x * g.Y + sd.y) * g.Z + sd.z] = (uint32_t)sd.w;
        const uint32_t tile = ((uint32_t)(sd.x / TX) * g.nty + sd.y / TY) * g.ntz + sd.z / TZ;
        wl.occ[tile] = 1;
        enqueue_tile(wl, tile, round);
    }
}

// phase 2 worklist: every tile that holds a freed (FREE) cell.
__global__ void __launch_bounds__(256) enqueue_tiles_with_free_kernel(const uint16_t* __restrict__ grid, TileGeom g, Worklist wl, uint32_t round)
{
    const size_t n = (size_t)g.X * g.Y * g.Z;
    if (blockIdx.x == 0 && threadIdx.x == 0) wl.lo[round % 3] = 0, wl.lo[(round + 1) % 3] = kNoLevel;  // every source sits at level 0
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (grid[i] != VF_VOXEL_FREE) continue;
        const int z = (int)(i % g.Z);
        const size_t r = i / g.Z;
        const int y = (int)(r % g.Y), x = (int)(r / g.Y);
        const uint32_t tile = ((uint32_t)(x / TX) * g.nty + y / TY) * g.ntz + z / TZ;
        if (wl.stamp[tile] != round) enqueue_tile(wl, tile, round);
    }
}

// ------------------------------------------------------------------------------------------------ tile load helpers
// Stages tile + halo keys with 4-byte cp.async: a warp covers one 128-byte z-row per instruction and all ~40 rows of a warp
// are in flight at once (a plain load -> store loop serialises one global round trip per row).  Caller waits + syncs.
template <int NNEIGH>
__device__ __forceinline__ void load_tile_async(uint32_t* sk, const uint32_t* __restrict__ src, const TileGeom& g, int gx0, int gy0, int gz0, uint32_t outside)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < (TX + 2) * SY; row += kThreads / 32) {
        const int x = row / SY - 1, y = row % SY - 1;
        const int gx = gx0 + x, gy = gy0 + y;
        bool rowin = gx >= 0 && gx < g.X && gy >= 0 && gy < g.Y;
        if (NNEIGH == 6 && (x < 0 || x >= TX) && (y < 0 || y >= TY)) rowin = false;  // corner columns are never read
        const uint32_t* base = src + ((size_t)(rowin ? gx : 0) * g.Y + (rowin ? gy : 0)) * g.Z;
        const int gz = gz0 + lane;
        uint32_t* dst = &sk[sidx(x, y, lane)];
        if (rowin && gz < g.Z) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(base + gz) : "memory");
        } else {
            *dst = outside;
        }
        if (lane < 2) {
            const int hz = lane ? gz0 + TZ : gz0 - 1;
            uint32_t* hd = &sk[sidx(x, y, lane ? TZ : -1)];
            if (rowin && hz >= 0 && hz < g.Z) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(hd);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(base + hz) : "memory");
            } else {
                *hd = outside;
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// The same staging as ONE TMA box load (18 x 18 rows x 40 keys; the box starts ZPAD keys before the tile in z because its first
// element must sit on a 16-byte boundary, and one row / plane before it in y / x).  The copy engine zero-fills what lies outside
// the grid, and a zero key would be a seed, so tiles on the grid border patch those cells to WALL.
__device__ __forceinline__ void load_tile_tma(uint32_t* sk, const CUtensorMap* map, unsigned bar, unsigned& parity, const TileGeom& g, int gx0, int gy0,
                                              int gz0)
{
    const int t = threadIdx.x;
    if (t == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the previous visit's accesses to the tile come first
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(kCells * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(sk)),
                     "l"(reinterpret_cast<unsigned long long>(map)), "r"(bar), "r"(gz0 - ZPAD), "r"(gy0 - 1), "r"(gx0 - 1)
                     : "memory");
    }
    unsigned done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
    parity ^= 1u;
    if (gx0 == 0 || gy0 == 0 || gz0 == 0 || gx0 + TX >= g.X || gy0 + TY >= g.Y || gz0 + TZ >= g.Z) {
        for (int row = t; row < (TX + 2) * SY; row += kThreads) {
            const int gx = gx0 + row / SY - 1, gy = gy0 + row % SY - 1;
            uint32_t* p = &sk[row * SZ];
            if ((unsigned)gx >= (unsigned)g.X || (unsigned)gy >= (unsigned)g.Y) {
#pragma unroll
                for (int k = 0; k < SZ; k += 4) *reinterpret_cast<uint4*>(p + k) = make_uint4(KEY_WALL, KEY_WALL, KEY_WALL, KEY_WALL);
            } else {
                if (gz0 == 0) *reinterpret_cast<uint4*>(p) = make_uint4(KEY_WALL, KEY_WALL, KEY_WALL, KEY_WALL);
                for (int k = max(0, g.Z - gz0 + ZPAD); k < SZ; ++k) p[k] = KEY_WALL;  // staged cell k holds gz = gz0 - ZPAD + k
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ F2: key relaxation round
template <int NNEIGH>
__device__ __forceinline__ uint32_t min_neighbour_key(const uint32_t* sk, int x, int y, int z)
{
    uint32_t m;
    if (NNEIGH == 6) {
        m = min(min(sk[sidx(x - 1, y, z)], sk[sidx(x + 1, y, z)]), min(sk[sidx(x, y - 1, z)], sk[sidx(x, y + 1, z)]));
        m = min(m, min(sk[sidx(x, y, z - 1)], sk[sidx(x, y, z + 1)]));
    } else {
        m = KEY_WALL;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dz = -1; dz <= 1; ++dz)
                    if (dx | dy | dz) m = min(m, sk[sidx(x + dx, y + dy, z + dz)]);
    }
    return m;
}

// COOP: ONE cooperative launch runs all rounds of a flood phase — the round loop lives on the device, rounds are separated by a grid barrier
// instead of a kernel boundary, and the host is not involved until the phase has converged (it reads the id of the next round with the
// flood's statistics).  Data that other CTAs wrote in an earlier round of the same launch (lists, counters, pending masks, keys) is read
// past the L1 (__ldcg, TMA; the host only launches this variant when the keys qualify for TMA staging).  !COOP: one launch per round (several jobs sharing a GPU interleave their rounds that way).
template <int NNEIGH, bool COOP>
__global__ void __launch_bounds__(kThreads, 4) flood_round_kernel(uint32_t* __restrict__ keys, const __grid_constant__ CUtensorMap keys_map, int use_tma,
                                                                  TileGeom g, Worklist wl, uint32_t round_arg, uint32_t last_round)
{
    extern __shared__ __align__(128) uint32_t sm[];
    uint32_t* sk = sm;
    uint32_t* act = sm + kCells;        // [2][kThreads] wavefront bitmasks, one word per z-row
    uint32_t* chg = act + 2 * kThreads;  // [kThreads]    cells whose key was lowered during this visit
    uint32_t* nw = chg + kThreads;       // [kThreads]    cells that are not walls
    uint32_t* pnd = nw + kThreads;       // [kThreads]    candidates whose new level lies beyond this round's window
    uint32_t* misc = pnd + kThreads;     // [0..3] neighbour-tile mask, [4] lowest deferred level
    const unsigned bar = (unsigned)__cvta_generic_to_shared(misc + 8);  // 8-byte aligned mbarrier for the TMA loads
    unsigned parity = 0;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (use_tma && (COOP || blockIdx.x < __ldcg(&wl.count[round_arg % 3]))) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
  for (uint32_t round = round_arg;; ++round) {
    const uint32_t count = __ldcg(&wl.count[round % 3]);
    if (blockIdx.x == 0 && t == 0) {
        wl.count[(round + 2) % 3] = 0;
        wl.lo[(round + 2) % 3] = kNoLevel;
        if (count) atomicAdd(&wl.stats[ST_ROUNDS], 1u);
    }
    // !COOP: rounds are launched in batches without knowing how long the worklists are: CTAs beyond the list — all of them once the flood
    // has converged — leave before they set anything up, so that an empty round costs the GPU (and the other jobs sharing it) nothing.
    if (!COOP && blockIdx.x >= count) return;
    const uint32_t lo = __ldcg(&wl.lo[round % 3]);
    const uint32_t hi = lo >= kNoLevel - wl.levels ? kNoLevel : lo + wl.levels;  // levels this round may assign: < hi
    for (uint32_t wi = blockIdx.x; wi < count; wi += gridDim.x) {
        const uint32_t tile = __ldcg(&wl.list[round & 1][wi]);
        const int tz = tile % g.ntz, ty = (tile / g.ntz) % g.nty, tx = tile / (g.ntz * g.nty);
        const int gx0 = tx * TX, gy0 = ty * TY, gz0 = tz * TZ;

#ifdef VF_FLOOD_TIMING
        long long tick__ = clock64();
#endif
        if (use_tma) load_tile_tma(sk, &keys_map, bar, parity, g, gx0, gy0, gz0);
        else load_tile_async<NNEIGH>(sk, keys, g, gx0, gy0, gz0, KEY_WALL);
        if (t < 4) misc[t] = 0;
        if (t == 4) misc[4] = kNoLevel;
        __syncthreads();
        VF_TICK(0);

        // ---- per-row masks: non-wall cells and the entry candidates.
        //      A tile that was already relaxed in this phase left its own edges repaired, and its cells have not changed since
        //      (only this tile writes them): new violations can only end in the layer of cells next to the halo.  The first visit
        //      (seeds or phase-2 sources inside) and slab tiles that contain a neighbour GPU's plane check every cell.
        const uint32_t sflag = __ldcg(&wl.seen[tile]);
        const bool revisit = (sflag & 0x7Fu) == wl.epoch, has_pending = revisit && (sflag & 0x80u);
        const bool full_entry = !revisit || (g.fix_lo && tx == 0) || (g.fix_hi && tx == g.ntx - 1);
        {
            // warp per row, lane = z (conflict-free whatever the row stride); eight rows' loads in flight at a time, lane i keeps the mask of
            // row warp * 32 + i, then every lane writes its row's words
            unsigned myw = 0;
#pragma unroll
            for (int i0 = 0; i0 < 32; i0 += 8) {
                uint32_t v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = warp * 32 + i0 + k;
                    v[k] = sk[sidx(r / TY, r % TY, lane)];
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = warp * 32 + i0 + k, x = r / TY;
                    // slab mode: halo planes are copies of a neighbour GPU's cells — sources only
                    const bool fixed = (g.fix_lo && gx0 + x == 0) || (g.fix_hi && gx0 + x == g.X - 1);
                    const unsigned w = __ballot_sync(kFull, !fixed && v[k] != KEY_WALL);
                    if (lane == i0 + k) myw = w;
                }
            }
            const int r = warp * 32 + lane, x = r / TY, y = r % TY;
            const bool face = x == 0 || x == TX - 1 || y == 0 || y == TY - 1;
            unsigned entry = (full_entry || face) ? 0xFFFFFFFFu : 0x80000001u;
            if (has_pending) entry |= __ldcg(&wl.pend[(size_t)tile * kThreads + r]);  // candidates deferred by the previous visit
            nw[r] = myw;
            chg[r] = 0;
            pnd[r] = 0;
            act[r] = myw & entry;
        }
        __syncthreads();

        VF_TICK(1);
        // ---- relaxation steps in shared memory, pull style, one barrier per step.  Thread t owns row (lane * 8 + warp): rows are
        //      dealt round-robin so that a flat front spreads over all warps.  Step 0 takes the entry candidates; step k > 0 takes
        //      the cells next to the cells lowered in step k-1 (wavefront masks shifted by +-1 in z | masks of the neighbouring
        //      rows) & not-wall.  A warp serves its 32 rows either row by row (lane = z; cheapest when the rows are well filled) or
        //      by dealing the candidate CELLS to its lanes 32 at a time (prefix sum of the popcounts, owner by binary search, n-th
        //      set bit; ceil(candidates / 32) neighbourhood evaluations whatever the orientation of the front).  Every candidate takes min(neighbour keys) + 1 level;
        //      the cells that got lower are the next wavefront.  Keys only decrease and every written value is the key of a real
        //      path, so reading a neighbour while another warp lowers it is harmless; the fixed point is the same.
        {
            const int myrow = lane * 8 + warp, mx = myrow / TY, my = myrow % TY;
            uint32_t dmin = kNoLevel;
            for (int it = 0; it < TX * TY * TZ; ++it) {
#ifdef VF_FLOOD_TIMING
                const long long step_t0 = clock64();
#endif
                const uint32_t* cur = act + (it & 1) * kThreads;
                uint32_t* nxt = act + ((it & 1) ^ 1) * kThreads;
                unsigned cand;
                if (it == 0) {
                    cand = cur[myrow];
                } else {
                    const unsigned a = cur[myrow];
                    cand = (a << 1) | (a >> 1);
                    if (NNEIGH == 26) cand |= a;
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy) {
                            if ((dx | dy) == 0) continue;
                            if (NNEIGH == 6 && dx != 0 && dy != 0) continue;
                            const int nx = mx + dx, ny = my + dy;
                            if (nx < 0 || nx >= TX || ny < 0 || ny >= TY) continue;
                            const unsigned an = cur[nx * TY + ny];
                            cand |= NNEIGH == 26 ? (an | (an << 1) | (an >> 1)) : an;
                        }
                    cand &= nw[myrow];
                }
                nxt[myrow] = 0;
                __syncwarp();
                bool any = false, overflow = false;
                // relax one candidate cell; its fate: 0 unchanged, 1 lowered, 2 deferred (beyond this round's window)
                auto relax = [&](int x, int y, int z) -> int {
                    const uint32_t m = min_neighbour_key<NNEIGH>(sk, x, y, z);
                    if (m < KEY_LIMIT) {
                        const uint32_t c = m + KEY_LEVEL;
                        if (c < sk[sidx(x, y, z)]) {
                            if ((c >> KEY_SHIFT) < hi) {
                                sk[sidx(x, y, z)] = c;
                                return 1;
                            }
                            dmin = min(dmin, c >> KEY_SHIFT);
                            return 2;
                        }
                    } else if (m < KEY_UNREACHED) {
                        overflow = true;
                    }
                    return 0;
                };
                const unsigned rows = __ballot_sync(kFull, cand != 0);  // bit l <-> row l * 8 + warp
                const int total = (int)__reduce_add_sync(kFull, (unsigned)__popc(cand));
                if (__reduce_max_sync(kFull, (unsigned)__popc(cand)) <= kOwnRowMax) {
                    // thin fronts (the diagonal planes of a Manhattan front, any front moving along z): a candidate or two per row.  Every lane
                    // relaxes the candidates of its OWN row one after the other — no dealing, no shuffles, no atomics: only the owner writes
                    // its row's words
                    unsigned low = 0, def = 0;
                    for (unsigned cb = cand; cb; cb &= cb - 1) {
                        const int z = __ffs(cb) - 1;
                        const int f = relax(mx, my, z);
                        if (f == 1) low |= 1u << z;
                        else if (f == 2) def |= 1u << z;
                    }
                    if (low) nxt[myrow] = low, chg[myrow] |= low, any = true;
                    if (def) pnd[myrow] |= def;
                } else if (2 * __popc(rows) <= 3 * ((total + 31) / 32) + 1) {
                    // well-filled rows (a thick front in a solid region): row by row, lane = z, results by ballot
                    for (unsigned rr = rows; rr; rr &= rr - 1) {
                        const int l = __ffs(rr) - 1;
                        const unsigned cw = __shfl_sync(kFull, cand, l);
                        const int r = l * 8 + warp;
                        const int f = (cw >> lane & 1u) ? relax(r / TY, r % TY, lane) : 0;
                        const unsigned low = __ballot_sync(kFull, f == 1), def = __ballot_sync(kFull, f == 2);
                        if (lane == 0) {
                            if (low) nxt[r] = low, chg[r] |= low;
                            if (def) pnd[r] |= def;
                        }
                        any = any || low != 0;
                    }
                } else {
                    // sparse rows (thin shells, fronts moving along z): deal the candidate cells to the lanes 32 at a time
                    int incl = __popc(cand);
#pragma unroll
                    for (int dlt = 1; dlt < 32; dlt <<= 1) {
                        const int up = __shfl_up_sync(kFull, incl, dlt);
                        if (lane >= dlt) incl += up;
                    }
                    for (int base = 0; base < total; base += 32) {
                        const int j = base + lane;
                        int owner = 0;  // first lane whose inclusive prefix exceeds j
#pragma unroll
                        for (int sft = 16; sft >= 1; sft >>= 1) {
                            const int v = __shfl_sync(kFull, incl, owner + sft - 1);
                            if (v <= j) owner += sft;
                        }
                        owner = min(owner, 31);
                        const unsigned cw = __shfl_sync(kFull, cand, owner);
                        const int before = __shfl_sync(kFull, incl, owner) - __popc(cw);
                        if (j < total) {
                            const int z = __fns(cw, 0, j - before + 1);
                            const int r = owner * 8 + warp;
                            const int f = relax(r / TY, r % TY, z);
                            if (f == 1) {
                                atomicOr(&nxt[r], 1u << z);
                                atomicOr(&chg[r], 1u << z);
                                any = true;
                            } else if (f == 2) {
                                atomicOr(&pnd[r], 1u << z);
                            }
                        }
                    }
                }
                if (overflow) atomicOr(&wl.stats[ST_ERROR], 1u);
#ifdef VF_FLOOD_TIMING
                const long long step_t1 = clock64();
                const int more = __syncthreads_or(any);
                if (t == 0) {
                    atomicAdd(&g_flood_cycles[7], (unsigned long long)(clock64() - step_t1));  // warp 0's wait at the step barrier
                    if (it == 0) atomicAdd(&g_flood_cycles[6], (unsigned long long)(clock64() - step_t0));  // the entry step
                }
                if (!more) {
#else
                if (!__syncthreads_or(any)) {
#endif
                    if (t == 0) atomicAdd(&wl.stats[ST_STEPS], (uint32_t)it + 1), atomicMax(&wl.stats[ST_MAXSTEPS], (uint32_t)it + 1);
#ifdef VF_FLOOD_TIMING
                    if (t == 0) atomicAdd(&g_flood_cycles[5], (unsigned long long)it + 1), atomicAdd(&g_flood_cycles[4], 1ull);
#endif
                    break;
                }
            }
            dmin = __reduce_min_sync(kFull, dmin);
            if (lane == 0 && dmin != kNoLevel) atomicMin(&misc[4], dmin);
        }
        __syncthreads();
        VF_TICK(2);
        {
            // deferred candidates: remember them, come back next round, and tell the next round where its window starts
            const uint32_t tile_dmin = misc[4];
            if (tile_dmin != kNoLevel) wl.pend[(size_t)tile * kThreads + t] = pnd[t];
            if (t == 0) {
                wl.seen[tile] = (uint8_t)(wl.epoch | (tile_dmin != kNoLevel ? 0x80u : 0u));
                if (tile_dmin != kNoLevel) {
                    atomicMin(&wl.lo[(round + 1) % 3], tile_dmin);
                    enqueue_tile(wl, tile, round + 1);
                }
            }
        }

        // ---- write back the rows that changed (warp per row: one 128-byte line), wake the neighbours that saw them change
        unsigned nchanged = 0;
        {
            const unsigned mine = chg[warp * 32 + lane];
            nchanged = __popc(mine);
            for (unsigned rows = __ballot_sync(kFull, mine != 0); rows; rows &= rows - 1) {  // only the rows that changed
                const int r = warp * 32 + __ffs(rows) - 1;
                const int x = r / TY, y = r % TY, gz = gz0 + lane;
                if (gz < g.Z && gx0 + x < g.X && gy0 + y < g.Y) keys[((size_t)(gx0 + x) * g.Y + gy0 + y) * g.Z + gz] = sk[sidx(x, y, lane)];
            }
            nchanged = __reduce_add_sync(kFull, nchanged);
        }
        if (lane == 0 && nchanged) atomicAdd(&wl.stats[ST_CHANGED], nchanged);
        if (t == 0) atomicAdd(&wl.stats[ST_VISITS], 1u);
        enqueue_neighbours<NNEIGH>(g, wl, tx, ty, tz, chg[t], &misc[3], round + 1);
        __syncthreads();
        VF_TICK(3);
    }
    if (!COOP) return;
    // ---- next round: everybody's key stores, list appends and pending masks are visible (also to the copy engine's reads) after the barrier
    // (a hand-rolled release / acquire counter barrier was measured against grid.sync(): no difference — the round is bound by its slowest
    // tile visit, not by the barrier)
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
    cooperative_groups::this_grid().sync();
    if (__ldcg(&wl.count[(round + 1) % 3]) == 0 || round + 1 >= last_round) {
        if (blockIdx.x == 0 && t == 0) wl.stats[kRoundWord] = round + 1;  // the id the next phase starts with
        return;
    }
  }
}

// ------------------------------------------------------------------------------------------------ F2: thin-front solver
// State in the `tiles` arena behind the worklists: the first front of a phase as the set-up kernel leaves it (hdr[0] cells in `list`).
struct Front {
    uint32_t* list;
    uint32_t* hdr;
};

// Phase 1 sources.  FloodFracturer.cpp:102-103,127-132: seeds written in order (a later seed on the same cell overwrites); the level-0
// front holds every seeded cell once.  One CTA.
__global__ void __launch_bounds__(256) flood_front_seed_kernel(uint32_t* __restrict__ keys, TileGeom g, Worklist wl, Front fr, const ushort4* __restrict__ seeds, int S,
                                                               uint32_t round)
{
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) {
        s_n = 0;
        wl.lo[round % 3] = 0, wl.lo[(round + 1) % 3] = kNoLevel;  // where the tiles' first window starts, should they take over
        for (int s = 0; s < S; ++s) {
            const ushort4 sd = seeds[s];
            keys[((size_t)sd.x * g.Y + sd.y) * g.Z + sd.z] = (uint32_t)s;  // dist 0, order s
            wl.occ[((uint32_t)(sd.x / TX) * g.nty + sd.y / TY) * g.ntz + sd.z / TZ] = 1;
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const ushort4 sd = seeds[s];
        const uint32_t cell = ((uint32_t)sd.x * g.Y + sd.y) * g.Z + sd.z;
        if (__ldcg(keys + cell) == (uint32_t)s) fr.list[atomicAdd(&s_n, 1u)] = cell;  // S <= 4096 <= kFrontCap
    }
    __syncthreads();
    if (threadIdx.x == 0) fr.hdr[0] = s_n;
}

// Phase 2 sources are all labelled cells (keys of level 0 after launch_init_keys<true>): a FREE cell next to one is a level-1 cell and its key
// is 1 level | the lowest order among those neighbours.  Level-1 cells are the first front (hdr[0] zeroed by the host).
template <int NNEIGH>
__global__ void __launch_bounds__(256) flood_front_level1_kernel(const uint16_t* __restrict__ grid, uint32_t* __restrict__ keys, TileGeom g, Worklist wl, Front fr,
                                                                 uint32_t round)
{
    const uint32_t n = (uint32_t)g.X * g.Y * g.Z, Z = g.Z, YZ = (uint32_t)g.Y * g.Z;
    if (blockIdx.x == 0 && threadIdx.x == 0) wl.lo[round % 3] = 0, wl.lo[(round + 1) % 3] = kNoLevel;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (grid[i] != VF_VOXEL_FREE) continue;
        const int x = (int)(i / YZ), r = (int)(i - (uint32_t)x * YZ), y = r / (int)Z, z = r - y * (int)Z;
        uint32_t m = KEY_WALL;  // a neighbour that became a level-1 cell meanwhile (or is seen stale as unreached) is not a source either way
        if (NNEIGH == 6) {
            if (x > 0) m = min(m, keys[i - YZ]);
            if (x + 1 < g.X) m = min(m, keys[i + YZ]);
            if (y > 0) m = min(m, keys[i - Z]);
            if (y + 1 < g.Y) m = min(m, keys[i + Z]);
            if (z > 0) m = min(m, keys[i - 1]);
            if (z + 1 < g.Z) m = min(m, keys[i + 1]);
        } else {
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dz = -1; dz <= 1; ++dz) {
                        if (!(dx | dy | dz) || x + dx < 0 || x + dx >= g.X || y + dy < 0 || y + dy >= g.Y || z + dz < 0 || z + dz >= g.Z) continue;
                        m = min(m, keys[(int64_t)i + (int64_t)dx * (int)YZ + dy * (int)Z + dz]);
                    }
        }
        if (m < KEY_LEVEL) {
            keys[i] = m + KEY_LEVEL;
            const uint32_t q = atomicAdd(&fr.hdr[0], 1u);
            if (q < kFrontCap) fr.list[q] = i;  // a longer list makes the solver hand over at once: the tiles check every cell on their first visit
        }
    }
}

// atomicMin under a predicate instead of a branch: the claims of one cell (up to 26) are then in flight together; returns the old value, 0 when not executed
__device__ __forceinline__ uint32_t atom_min_if(uint32_t* p, uint32_t v, bool doit)
{
    uint32_t old;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.u32 %0, 0;\n\t@p atom.global.min.u32 %0, [%1], %2;\n\t}"
        : "=r"(old)
        : "l"(p), "r"(v), "r"((uint32_t)doit)
        : "memory");
    return old;
}

// the same for a counter that may live in another CTA's shared memory (generic address): slot reservations of one thread are in flight together
__device__ __forceinline__ uint32_t atom_inc_if(uint32_t* p, bool doit)
{
    uint32_t old;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0;\n\t@p atom.add.u32 %0, [%1], 1;\n\t}"
        : "=r"(old)
        : "l"(p), "r"((uint32_t)doit)
        : "memory");
    return old;
}

// One launch = one thread-block cluster (every CTA of the grid).  Label-correcting relaxation of the same keys, at cell granularity: the
// front is a list of (cell, key) pairs in SHARED memory; expanding a pair lowers the neighbours to key + 1 level with atomicMin, and whoever
// lowers a cell — unreached, or claimed at a higher (level, order) — lists it with the key it wrote.  A pair that is out of date (its cell
// was lowered again meanwhile) only repeats work: the thread that lowered it listed the better pair.  The least fixed point is the tiles'.
//
// A CTA expands its own list for `sublevels` steps between two cluster barriers (a __syncthreads per step; a flood is a chain of hundreds
// of dependent levels, so what a level costs is memory round trips and barriers, not bandwidth): a front that grew out of one seed stays
// with one CTA and advances level by level exactly as a BFS; where fronts of different CTAs meet they may run ahead of each other by up to
// `sublevels` levels and correct each other.  A CTA that holds more than twice its share of the front deals its claims out instead — into
// the inbox of the CTA a hash of the cell names, through distributed shared memory; inboxes are emptied after each cluster barrier.
// The neighbour keys are read first (plain loads: a line loaded earlier in the same barrier interval may lack the latest claims, which only
// costs a redundant atomic; the barrier's acquire empties the L1) and only cells that are not walls and lie above are lowered.  (Measured
// alternative: atomicMin on every neighbour without looking, walls written back — one round trip less, but slower: 6 atomics per cell.)
// The launch ends when no pair is pending anywhere (the phase has converged: the tile worklist stays empty), or when more than `limit`
// pairs are pending or a list overflowed: then every occupied tile is enqueued for round `tile_round` and the tiles' full first-visit check
// takes it from whatever upper bounds the keys hold.  spread: 0 = lists stay local, 1 = as described, 2 = always dealt out (tools/ only).
constexpr int kFrontInbox = 2048;  // pairs a CTA can receive between two cluster barriers
constexpr size_t kFrontSmemBytes = (size_t)(2 * kFrontLocal + 2 * kFrontInbox) * 8;

template <int NNEIGH>
__global__ void __launch_bounds__(kFrontThreads, 1) flood_front_kernel(uint32_t* __restrict__ keys, TileGeom g, Worklist wl, Front fr,
                                                                       uint32_t limit, uint32_t tile_round, int spread, int sublevels)
{
    namespace cg = cooperative_groups;
    constexpr uint32_t kLost = 0x80000000u;  // in a pending count: a list overflowed (or a level does not fit the key), some lowered cells are not listed
    extern __shared__ __align__(16) uint32_t s_dyn[];
    // two lists (the step's / the next step's) and two inboxes (this / the next barrier interval) of (cell, key) pairs
    auto list_cell = [&](int b) { return s_dyn + b * kFrontLocal; };
    auto list_key = [&](int b) { return s_dyn + (2 + b) * kFrontLocal; };
    auto inbox_cell = [&](int b) { return s_dyn + 4 * kFrontLocal + b * kFrontInbox; };
    auto inbox_key = [&](int b) { return s_dyn + 4 * kFrontLocal + (2 + b) * kFrontInbox; };
    __shared__ uint32_t s_cnt[2], s_icnt[2];  // entries (may count past the capacity)
    __shared__ uint32_t s_sent;               // pairs this CTA dealt out in this interval | kLost
    __shared__ uint32_t s_tot[2][16];         // pending pairs of every CTA of the cluster, written by its owner before the barrier
    const cg::cluster_group cluster = cg::this_cluster();
    const uint32_t t = threadIdx.x, T = blockDim.x, C = gridDim.x, rank = blockIdx.x, nthreads = C * T, gtid = rank * T + t;
    const uint32_t Z = g.Z, YZ = (uint32_t)g.Y * g.Z;
#ifdef VF_FLOOD_TIMING
    __shared__ unsigned long long s_ft[12];
    if (t < 12) s_ft[t] = 0;
#endif
    if (t < 2) s_cnt[t] = 0, s_icnt[t] = 0;
    if (t == 0) s_sent = 0;
    __syncthreads();
    // the first front: the cells the set-up kernel listed in global memory, dealt round-robin to the CTAs, with the keys they hold
    const uint32_t init_total = fr.hdr[0];
    const bool skip = init_total > limit;  // already too wide: nothing is listed, the tiles do everything
    if (!skip)
        for (uint32_t i = t * C + rank; i < init_total; i += nthreads) {
            const uint32_t cell = fr.list[i], p = atomicAdd(&s_cnt[0], 1u);
            if (p < (uint32_t)kFrontLocal) list_cell(0)[p] = cell, list_key(0)[p] = keys[cell];
        }
    const bool lost0 = !skip && (init_total + C - 1) / C > (uint32_t)kFrontLocal;  // a small cluster cannot hold a front that wide
    cluster.sync();  // also: nobody writes into a CTA whose counters are not zeroed yet
#ifdef VF_FLOOD_TIMING
    long long ftick__ = clock64();
#endif
    uint32_t total = init_total | (lost0 ? kLost : 0u), steps = 0;
    int cur = 0;
    bool handover = false;
    for (uint32_t interval = 0; !skip; ++interval) {
        if (total == 0) break;  // converged (the same word in every thread of the cluster)
        if ((total & ~kLost) > limit || (total & kLost) || steps >= kFrontMaxLevel) {
            handover = true;
            break;
        }
        const uint32_t fair = total / C;
        const int ib = interval & 1;
        for (int sl = 0; sl < sublevels; ++sl, cur ^= 1) {
            const uint32_t mine = min(s_cnt[cur], (uint32_t)kFrontLocal);
            const bool far = spread == 2 || (spread == 1 && mine > 2 * fair + 64);  // this CTA holds too much of the front: deal its claims out
            const uint32_t* ccell = list_cell(cur);
            const uint32_t* ckey = list_key(cur);
            uint32_t* ncell = list_cell(cur ^ 1);
            uint32_t* nkey = list_key(cur ^ 1);
            VF_FTICK(0, mine);
            // 26 neighbours: three threads per pair (one per x-plane of the neighbourhood), so that a thin front still fills the CTA
            constexpr int LANES = NNEIGH == 6 ? 1 : 3, NB = NNEIGH == 6 ? 6 : 9;
            for (uint32_t it = t; it < mine * LANES; it += T) {
                const uint32_t i = it / LANES;
                const int dx = NNEIGH == 6 ? 0 : (int)(it - i * LANES) - 1;
                const uint32_t cell = ccell[i], own = ckey[i];
                if (own >= KEY_LIMIT) {  // no room for another level in the key: the tiles report it
                    atomicOr(&s_sent, kLost);
                    continue;
                }
                const uint32_t nk = own + KEY_LEVEL;
                const int x = (int)(cell / YZ), r = (int)(cell - (uint32_t)x * YZ), y = r / (int)Z, z = r - y * (int)Z;
                uint32_t nb[NB], kv[NB];
                uint32_t inside = 0;
                if (NNEIGH == 6) {
                    nb[0] = cell - YZ, nb[1] = cell + YZ, nb[2] = cell - Z, nb[3] = cell + Z, nb[4] = cell - 1, nb[5] = cell + 1;
                    inside = (x > 0 ? 1u : 0u) | (x + 1 < g.X ? 2u : 0u) | (y > 0 ? 4u : 0u) | (y + 1 < g.Y ? 8u : 0u) | (z > 0 ? 16u : 0u) | (z + 1 < g.Z ? 32u : 0u);
                } else {
                    const bool xin = x + dx >= 0 && x + dx < g.X;
                    const uint32_t plane = cell + (uint32_t)(dx * (int)YZ);  // modulo 2^32; only used when the neighbour lies inside the grid
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int dy = q / 3 - 1, dz = q % 3 - 1;
                        nb[q] = plane + (uint32_t)(dy * (int)Z + dz);
                        const bool in = xin && (dx != 0 || q != 4) && y + dy >= 0 && y + dy < g.Y && z + dz >= 0 && z + dz < g.Z;
                        inside |= (in ? 1u : 0u) << q;
                    }
                }
                // look first (plain loads: a stale line only costs a redundant atomic), then lower what is not a wall and lies above
#pragma unroll
                for (int q = 0; q < NB; ++q) kv[q] = (inside >> q & 1u) ? keys[nb[q]] : KEY_WALL;
#ifdef VF_FLOOD_TIMING
                {
                    uint32_t all = nk;
#pragma unroll
                    for (int q = 0; q < NB; ++q) all ^= kv[q];
                    VF_FTICK(1, all);
                }
#endif
#pragma unroll
                for (int q = 0; q < NB; ++q) kv[q] = atom_min_if(keys + nb[q], nk, kv[q] != KEY_WALL && kv[q] > nk);  // old value; 0 = not executed
                uint32_t won = 0;  // bit q: this thread lowered neighbour q (unreached, or a correction: claimed at a higher (level, order))
#pragma unroll
                for (int q = 0; q < NB; ++q) won |= (kv[q] > nk && kv[q] != KEY_WALL ? 1u : 0u) << q;
                VF_FTICK(2, won);
                if (won) {
                    const uint32_t k = __popc(won);
                    bool lost = false;
                    if (!far) {
                        uint32_t p = atomicAdd(&s_cnt[cur ^ 1], k);
#pragma unroll
                        for (int q = 0; q < NB; ++q)
                            if (won >> q & 1u) {
                                if (p < (uint32_t)kFrontLocal) ncell[p] = nb[q], nkey[p] = nk;
                                else lost = true;
                                ++p;
                            }
                    } else {
#pragma unroll
                        for (int q = 0; q < NB; ++q) {  // reserve all slots, then fill them
                            const unsigned to = ((nb[q] * 2654435761u) >> 20) & (C - 1);  // C is a power of two
                            kv[q] = atom_inc_if(cluster.map_shared_rank(&s_icnt[ib], to), won >> q & 1u);
                        }
#pragma unroll
                        for (int q = 0; q < NB; ++q)
                            if (won >> q & 1u) {
                                const unsigned to = ((nb[q] * 2654435761u) >> 20) & (C - 1);
                                if (kv[q] < (uint32_t)kFrontInbox) {
                                    cluster.map_shared_rank(inbox_cell(ib), to)[kv[q]] = nb[q];
                                    cluster.map_shared_rank(inbox_key(ib), to)[kv[q]] = nk;
                                } else {
                                    lost = true;
                                }
                            }
                        atomicAdd(&s_sent, k);
                    }
                    if (lost) atomicOr(&s_sent, kLost);
                }
                VF_FTICK(3, won);
            }
            __syncthreads();  // the step's list is consumed, the next one complete
            if (t == 0) s_cnt[cur] = 0;
#ifdef VF_FLOOD_TIMING
            if (gtid == 0) s_ft[7] += 1, s_ft[10] += (mine * LANES + T - 1) / T;
#endif
            __syncthreads();
            VF_FTICK(4, 0);
        }
        steps += (uint32_t)sublevels;
        // pending pairs of this CTA: its next list + what it dealt out; told to every CTA of the cluster
        if (t < C) *cluster.map_shared_rank(&s_tot[ib][rank], t) = (min(s_cnt[cur], (uint32_t)kFrontLocal) + (s_sent & ~kLost)) | (s_sent & kLost) | (s_cnt[cur] > (uint32_t)kFrontLocal ? kLost : 0u);
#ifdef VF_FLOOD_TIMING
        if (gtid == 0) s_ft[9] += total & ~kLost, s_ft[11] += 1;
#endif
        cluster.sync();  // release / acquire at cluster scope: keys, inbox entries and counts of this interval are visible to every CTA
        uint32_t sum = 0, flags = 0;
        for (uint32_t r = 0; r < C; ++r) sum += s_tot[ib][r] & ~kLost, flags |= s_tot[ib][r] & kLost;
        total = sum | flags;  // the same in every CTA
        // empty the inbox into the list of the next step (the senders of the next interval use the other inbox)
        const uint32_t have = min(s_cnt[cur], (uint32_t)kFrontLocal), got = min(s_icnt[ib], (uint32_t)kFrontInbox);
        const bool spilled = have + got > (uint32_t)kFrontLocal;  // known to this CTA only: it stops everybody through the next exchange
        for (uint32_t i = t; i < got; i += T)
            if (have + i < (uint32_t)kFrontLocal) list_cell(cur)[have + i] = inbox_cell(ib)[i], list_key(cur)[have + i] = inbox_key(ib)[i];
        __syncthreads();
        if (t == 0) s_cnt[cur] = min(have + got, (uint32_t)kFrontLocal), s_icnt[ib] = 0, s_sent = spilled ? kLost : 0u;
        __syncthreads();
        VF_FTICK(5, total);
    }
    if (handover || skip) {
        const uint32_t nt = (uint32_t)g.ntiles();
        for (uint32_t tile = gtid; tile < nt; tile += nthreads)
            if (wl.occ[tile]) enqueue_tile(wl, tile, tile_round);
        if (gtid == 0) wl.lo[tile_round % 3] = steps > 2u * (uint32_t)sublevels ? steps - 2u * (uint32_t)sublevels : 0u;  // where the first window starts (any value is correct)
    }
    if (gtid == 0) wl.stats[kFrontWord] += steps;
    cluster.sync();  // no CTA may exit while another may still write into its shared memory
#ifdef VF_FLOOD_TIMING
    if (gtid == 0) {
        for (int q = 0; q < 12; ++q) g_front_cycles[q] += s_ft[q];
        g_front_cycles[8] = gridDim.x;
    }
#endif
}

// keys -> label words.  order -> seeds[order].w & mask; unreached non-empty cells stay FREE (never claimed in the reference).
__global__ void __launch_bounds__(256) flood_finalize_kernel(const uint32_t* __restrict__ keys, uint16_t* __restrict__ grid, size_t n,
                                                             const ushort4* __restrict__ seeds, uint32_t mask, uint32_t* __restrict__ stats)
{
    uint32_t maxd = 0;
    auto label_of = [&](uint32_t k) -> uint32_t {
        if (k == KEY_WALL) return VF_VOXEL_EMPTY;
        if (k == KEY_UNREACHED) return VF_VOXEL_FREE;
        maxd = max(maxd, k >> KEY_SHIFT);
        return seeds[k & (KEY_LEVEL - 1)].w & mask;
    };
    if (n % 8 == 0 && (((uintptr_t)grid | (uintptr_t)keys) & 15) == 0) {  // 8 voxels per thread: two 128-bit key loads, one 128-bit label store
        for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n / 8; c += (size_t)gridDim.x * blockDim.x) {
            const uint4 a = vf_ldg_stream(reinterpret_cast<const uint4*>(keys + c * 8)), b = vf_ldg_stream(reinterpret_cast<const uint4*>(keys + c * 8 + 4));
            uint4 o;
            o.x = label_of(a.x) | label_of(a.y) << 16, o.y = label_of(a.z) | label_of(a.w) << 16;
            o.z = label_of(b.x) | label_of(b.y) << 16, o.w = label_of(b.z) | label_of(b.w) << 16;
            vf_stg_stream(reinterpret_cast<uint4*>(grid + c * 8), o);
        }
    } else {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) grid[i] = (uint16_t)label_of(keys[i]);
    }
    maxd = __reduce_max_sync(kFull, maxd);
    if ((threadIdx.x & 31) == 0 && maxd) atomicMax(&stats[ST_MAXDIST], maxd);
}

// ------------------------------------------------------------------------------------------------ host side
struct Job {
    vf_ctx* c;
    TileGeom g;
    Worklist wl;
    uint32_t round;       // next round id
    uint32_t* h_mail;     // pinned mailbox
    int blocks_stream;    // grid for streaming kernels
    int blocks_tiles;     // grid for tile kernels
    bool round_on_device = false;  // a cooperative phase advanced the round id on the device; read_stats brings it back
    Front fr;             // thin-front solver state
    uint32_t front_levels = 0;  // BFS levels it ran so far (read_stats)
};

vf_status job_begin(vf_grid* grid, Job& j)
{
    vf_ctx* c = grid->ctx;
    j.c = c;
    j.g = make_geom(grid->X, grid->Y, grid->Z);
    const size_t nt = (size_t)j.g.ntiles();
    // layout: stats[8] count[3] lo[3] round id, pad | pad[16] | list0 | list1 | stamp | occ | seen | (aligned) pend[nt][kThreads]
    const size_t words = 32 + 3 * nt;
    const size_t pend_off = (words * 4 + 2 * nt + 255) & ~(size_t)255;
    const size_t front_off = (pend_off + nt * kThreads * 4 + 255) & ~(size_t)255;
    VF_TRY(vf_scratch_reserve(c, c->tiles, front_off + kVfFrontBytes));
    uint32_t* base = (uint32_t*)c->tiles.ptr;
    j.fr.list = (uint32_t*)((char*)base + front_off);
    j.fr.hdr = j.fr.list + kFrontCap;
    j.wl.stats = base;
    j.wl.count = base + 8;
    j.wl.list[0] = base + 32;
    j.wl.list[1] = base + 32 + nt;
    j.wl.stamp = base + 32 + 2 * nt;
    j.wl.occ = (uint8_t*)(base + 32 + 3 * nt);
    j.wl.seen = j.wl.occ + nt;
    j.wl.epoch = 1;
    j.wl.lo = base + 11;
    j.wl.levels = c->flood_levels ? c->flood_levels : kLevelsPerRound;
    j.wl.pend = (uint32_t*)((char*)base + pend_off);
    VF_CUDA(cudaMemsetAsync(base, 0, 32 * 4, c->stream));
    VF_CUDA(cudaMemsetAsync(j.wl.stamp, 0, nt * 4 + 2 * nt, c->stream));
    j.round = 1;
    j.h_mail = (uint32_t*)((char*)c->pinned + 65536);  // upper half of the mailbox; the lower half carries seeds
    j.blocks_stream = c->num_sms * 8;
    j.blocks_tiles = c->num_sms * 4;
    return VF_OK;
}

// run worklist rounds until the list for the next round is empty; `launch(round)` issues one round kernel
template <typename LaunchF>
vf_status run_rounds(Job& j, LaunchF launch)
{
    vf_ctx* c = j.c;
    int batch = 4;
    for (int guard = 0; guard < 100000; ++guard) {
        for (int b = 0; b < batch; ++b) {
            launch(j.round + b);
            VF_LAUNCHED(c);
        }
        j.round += batch;
        VF_CUDA(cudaMemcpyAsync(j.h_mail, j.wl.count + (j.round % 3), 4, cudaMemcpyDeviceToHost, c->stream));
        VF_CUDA(vf_sync(c));
        if (j.h_mail[0] == 0) return VF_OK;
        batch = std::min(batch * 2, 32);
    }
    return vf_set_error(VF_ERR_CAPACITY, "tile worklist did not drain");
}

vf_status read_stats(Job& j, uint32_t out[8])
{
    VF_CUDA(cudaMemcpyAsync(j.h_mail, j.wl.stats, 64, cudaMemcpyDeviceToHost, j.c->stream));
    VF_CUDA(vf_sync(j.c));
    for (int i = 0; i < 8; ++i) out[i] = j.h_mail[i];
    j.front_levels = j.h_mail[kFrontWord];
    if (j.round_on_device) {  // a cooperative phase ran: the id of the next round is in the header
        VF_REQUIRE(j.h_mail[kRoundWord] < j.round + 100000, VF_ERR_CAPACITY, "tile worklist did not drain");
        j.round = j.h_mail[kRoundWord];
        j.round_on_device = false;
    }
    return VF_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map over the key field as (Z, Y, X) uint32 with a box of one staged tile; false when the layout does not qualify
bool make_keys_map(CUtensorMap* map, const uint32_t* keys, const TileGeom& g)
{
    static EncodeTiledFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    std::memset(map, 0, sizeof(*map));
    if (!encode || g.Z % 4 != 0 || ((uintptr_t)keys & 15) != 0) return false;  // global strides must be multiples of 16 bytes
    const cuuint64_t dims[3] = { (cuuint64_t)g.Z, (cuuint64_t)g.Y, (cuuint64_t)g.X };
    const cuuint64_t strides[2] = { (cuuint64_t)g.Z * 4, (cuuint64_t)g.Z * g.Y * 4 };
    const cuuint32_t box[3] = { SZ, SY, TX + 2 }, estr[3] = { 1, 1, 1 };
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(keys), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


template <int NNEIGH>
vf_status flood_phase(Job& j, uint32_t* keys)
{
    CUtensorMap map;
    int use_tma = make_keys_map(&map, keys, j.g) ? 1 : 0;  // otherwise: per-row cp.async staging
    vf_ctx* c = j.c;
    if (use_tma && c->flood_coop) {
        // one cooperative launch for the whole phase; the grid must be resident at once
        auto kern = flood_round_kernel<NNEIGH, true>;
        VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        int per_sm = 0;
        VF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, kSmemBytes));
        if (per_sm >= 1) {
            int blocks = c->num_sms * std::min(per_sm, c->flood_coop);
            uint32_t round = j.round, last_round = j.round + 100000;  // the guard of run_rounds
            void* args[] = { &keys, &map, &use_tma, &j.g, &j.wl, &round, &last_round };
            VF_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)blocks), dim3(kThreads), args, kSmemBytes, c->stream));
            ++c->launches;
            j.round_on_device = true;
            return VF_OK;
        }
    }
    auto kern = flood_round_kernel<NNEIGH, false>;
    VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    return run_rounds(j, [&](uint32_t r) { kern<<<j.blocks_tiles, kThreads, kSmemBytes, j.c->stream>>>(keys, map, use_tma, j.g, j.wl, r, 0u); });
}

// Largest cluster the thin-front kernel can run as on this device: 16 CTAs (non-portable size, one GPC), else 8 … else a single CTA.
template <int NNEIGH>
int front_cluster_size(vf_ctx* c)
{
    static int cached[64];  // per device; 0 = not probed yet (a benign race: every thread computes the same value)
    const int dev = c->device & 63;
    if (cached[dev]) return cached[dev];
    auto kern = flood_front_kernel<NNEIGH>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFrontSmemBytes) != cudaSuccess) {
        cudaGetLastError();
        return cached[dev] = -1;  // no thin-front solver on this device
    }
    int best = 1;
#ifdef VF_FLOOD_TIMING  // tools/ only
    const int want = std::getenv("VF_FRONT_CLUSTER") ? std::atoi(std::getenv("VF_FRONT_CLUSTER")) : 16;
#else
    const int want = 16;
#endif
    for (int size : { 16, 8, 4, 2 }) {
        if (size > want) continue;
        if (size > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            continue;
        }
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = (unsigned)size, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)size), cfg.blockDim = dim3(kFrontThreads), cfg.dynamicSmemBytes = kFrontSmemBytes, cfg.stream = c->stream, cfg.attrs = &at, cfg.numAttrs = 1;
        int clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg) == cudaSuccess && clusters >= 1) {
            best = size;
            break;
        }
        cudaGetLastError();
    }
    return cached[dev] = best;
}

inline int front_env(const char* name, int dflt)
{
#ifdef VF_FLOOD_TIMING  // tools/ only
    return std::getenv(name) ? std::atoi(std::getenv(name)) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// the cell-granular start of a flood phase (front list and count set up in j.fr); whatever it leaves undone is on the tile worklist afterwards
template <int NNEIGH>
vf_status front_phase(Job& j, uint32_t* keys)
{
    vf_ctx* c = j.c;
    const int size = front_cluster_size<NNEIGH>(c);
    VF_REQUIRE(size >= 1, VF_ERR_CUDA, "the thin-front flood kernel cannot run on this device");  // (vf_fracture_flood asks front_available first)
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = (unsigned)size, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)size), cfg.blockDim = dim3(kFrontThreads), cfg.dynamicSmemBytes = kFrontSmemBytes, cfg.stream = c->stream, cfg.attrs = &at, cfg.numAttrs = 1;
    static const int spread = front_env("VF_FRONT_SPREAD", 1), sublevels = std::max(1, front_env("VF_FRONT_SUBLEVELS", kFrontSublevels));
    VF_CUDA(cudaLaunchKernelEx(&cfg, flood_front_kernel<NNEIGH>, keys, j.g, j.wl, j.fr, c->flood_front, j.round, spread, sublevels));
    ++c->launches;
    return VF_OK;
}

// the front kernel needs 96 KB of dynamic shared memory per CTA; a device that cannot give it floods on the tiles alone
bool front_available(vf_ctx* c, int nneigh) { return (nneigh == 6 ? front_cluster_size<6>(c) : front_cluster_size<26>(c)) >= 1; }

}  // namespace

extern "C" vf_status vf_fracture_flood(vf_grid* grid, const uint32_t* seeds, uint32_t nseeds, int dfunc, int id_bits, vf_flood_stats* stats_out)
{
    VF_REQUIRE(grid != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    vf_ctx* c = grid->ctx;
    VF_TRY(vf_enter(c));
    VF_REQUIRE(dfunc >= 0 && dfunc <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    if (id_bits == 0) id_bits = 8;
    VF_REQUIRE(id_bits == 8 || id_bits == 15, VF_ERR_INVALID_ARGUMENT, "id_bits must be 0/8 or 15");
    VF_REQUIRE(nseeds >= 1 && nseeds < 32768 && nseeds <= 4096, VF_ERR_CAPACITY, "flood: seed count %u out of range [1, 4096]", nseeds);
    const int nneigh = dfunc == VF_MANHATTAN ? 6 : 26;  // FloodFracturer.cpp:114
    const size_t n = grid->n();

    ushort4* d_seeds = nullptr;
    VF_TRY(vf_upload_seeds(c, seeds, nseeds, grid->X, grid->Y, grid->Z, &d_seeds));
    VF_TRY(vf_scratch_reserve(c, c->keys, n * 4));
    uint32_t* keys = (uint32_t*)c->keys.ptr;
    Job j;
    VF_TRY(job_begin(grid, j));
    vf_flood_stats st = { 0, 0, 0, 0, 0, 0 };
    uint32_t hs[8];

    // F3 bookkeeping on the host: effective source per cell (later seed wins), principal (lowest-prefix) source per fragment id
    bool prefixes = false;
    std::vector<int> principal(256, -1);
    if (id_bits == 8) {
        std::vector<std::pair<uint64_t, int>> cells(nseeds);
        for (uint32_t s = 0; s < nseeds; ++s) {
            const uint64_t lin = ((uint64_t)seeds[4 * s] * grid->Y + seeds[4 * s + 1]) * grid->Z + seeds[4 * s + 2];
            cells[s] = std::make_pair(lin, (int)s);
        }
        std::sort(cells.begin(), cells.end());
        for (uint32_t i = 0; i < nseeds; ++i) {
            if (i + 1 < nseeds && cells[i + 1].first == cells[i].first) continue;  // overwritten by a later seed on the same cell
            const int s = cells[i].second;
            const uint32_t w = seeds[4 * s + 3], frag = w & 0xFFu, pre = w >> VF_ID_POSITION;
            if (pre) prefixes = true;
            const int cur = principal[frag];
            if (cur < 0 || pre < (seeds[4 * cur + 3] >> VF_ID_POSITION) || (pre == (seeds[4 * cur + 3] >> VF_ID_POSITION) && s < cur)) principal[frag] = s;
        }
    }

    // ---- phase 1
    VF_TRY(launch_init_keys<false>(c, grid->d, keys, j.g, j.wl.occ, nullptr, j.blocks_stream));
    const bool front = c->flood_front != 0 && n < ((size_t)1 << 32) && front_available(c, nneigh);  // thin-front solver first (cell indices are 32-bit there)
    if (front) {
        flood_front_seed_kernel<<<1, 256, 0, c->stream>>>(keys, j.g, j.wl, j.fr, d_seeds, (int)nseeds, j.round);
        VF_LAUNCHED(c);
        VF_TRY(nneigh == 6 ? front_phase<6>(j, keys) : front_phase<26>(j, keys));
    } else {
        flood_seed_kernel<<<1, 32, 0, c->stream>>>(keys, j.g, j.wl, d_seeds, (int)nseeds, j.round);
        VF_LAUNCHED(c);
    }
    VF_TRY(nneigh == 6 ? flood_phase<6>(j, keys) : flood_phase<26>(j, keys));  // nothing to do when the front solver finished the phase
    const bool need_f3 = id_bits == 8 && prefixes;
    flood_finalize_kernel<<<j.blocks_stream, 256, 0, c->stream>>>(keys, grid->d, n, d_seeds, (id_bits == 8 && !need_f3) ? 0xFFu : 0xFFFFu, j.wl.stats);
    VF_LAUNCHED(c);
    st.disjoint_rounds = 1;

    if (need_f3) {
        // ---- disjoint step: keep the component of each fragment's principal source, free the rest
        ushort4* h = (ushort4*)c->pinned;
        uint16_t* h_order = (uint16_t*)(h + 256);
        int nstart = 0;
        VF_CUDA(vf_sync(c));  // the mailbox still carries the seed upload
        for (int f = 0; f < 256; ++f) {
            h_order[f] = 0;
            if (principal[f] < 0) continue;
            const int s = principal[f];
            h[nstart++] = make_ushort4((unsigned short)seeds[4 * s], (unsigned short)seeds[4 * s + 1], (unsigned short)seeds[4 * s + 2],
                                       (unsigned short)seeds[4 * s + 3]);
            h_order[f] = (uint16_t)s;
        }
        ushort4* d_starts = d_seeds + ((nseeds + 255) & ~255u);
        uint16_t* d_order = (uint16_t*)(d_starts + 256);
        VF_CUDA(cudaMemcpyAsync(d_starts, h, 256 * sizeof(ushort4) + 256 * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
        VF_TRY(vf_k_keep_seed_components(grid, d_starts, nstart, 1, nneigh, j.wl.stats + ST_FREED));  // union-find (ccl.cu); reuses the key scratch
        VF_TRY(read_stats(j, hs));
        st.freed_voxels = hs[ST_FREED];
        if (hs[ST_FREED] != 0) {
            // ---- phase 2: re-flood from every labelled cell (FloodFracturer.cpp:135-177, second trip of the loop)
            j.wl.epoch = 2;  // every labelled cell is a source now: tiles start over with a full entry check
            VF_TRY(launch_init_keys<true>(c, grid->d, keys, j.g, j.wl.occ, d_order, j.blocks_stream));
            if (front) {
                VF_CUDA(cudaMemsetAsync(j.fr.hdr, 0, 4, c->stream));
                if (nneigh == 6) flood_front_level1_kernel<6><<<j.blocks_stream, 256, 0, c->stream>>>(grid->d, keys, j.g, j.wl, j.fr, j.round);
                else flood_front_level1_kernel<26><<<j.blocks_stream, 256, 0, c->stream>>>(grid->d, keys, j.g, j.wl, j.fr, j.round);
                VF_LAUNCHED(c);
                VF_TRY(nneigh == 6 ? front_phase<6>(j, keys) : front_phase<26>(j, keys));
            } else {
                enqueue_tiles_with_free_kernel<<<j.blocks_stream, 256, 0, c->stream>>>(grid->d, j.g, j.wl, j.round);
                VF_LAUNCHED(c);
            }
            VF_TRY(nneigh == 6 ? flood_phase<6>(j, keys) : flood_phase<26>(j, keys));
            flood_finalize_kernel<<<j.blocks_stream, 256, 0, c->stream>>>(keys, grid->d, n, d_seeds, 0xFFu, j.wl.stats);
            VF_LAUNCHED(c);
            st.disjoint_rounds = 2;
        } else {
            VF_TRY(vf_k_pointwise(grid, VF_PW_RIGHTMOST8));  // FloodFracturer.cpp:180-186
        }
    }
    VF_TRY(read_stats(j, hs));
    VF_REQUIRE(hs[ST_ERROR] == 0, VF_ERR_CAPACITY, "flood: geodesic distance exceeds the 17-bit key field");
    st.tile_visits = hs[ST_VISITS];
    st.tile_rounds = hs[ST_ROUNDS];
    st.max_dist = hs[ST_MAXDIST];
    st.front_levels = j.front_levels;
    if (stats_out) *stats_out = st;
    return VF_OK;
}

extern "C" vf_status vf_remove_isolated_regions(vf_grid* grid, const uint32_t* seeds, uint32_t nseeds)
{
    VF_REQUIRE(grid != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    vf_ctx* c = grid->ctx;
    VF_TRY(vf_enter(c));
    ushort4* d_seeds = nullptr;
    VF_TRY(vf_upload_seeds(c, seeds, nseeds, grid->X, grid->Y, grid->Z, &d_seeds));
    // descent certificate (c1_descent.cu); the union-find below takes over when it declines.  A context whose last grid was declined (thin
    // shells: batch producers work through similar shapes) goes straight to the union-find and tries the certificate again every 16th call.
    // Mode 0 (default) tries it on grids of at least 2^26 cells only: that is where it pays (512^3: 0.19 against 0.60 ms), while on a small
    // thin shell the certificate keeps fewer than its 4096-cell bail-out uncertified yet leaves one CTA list work over most of the shell —
    // measured on BASELINE cfg1 (88 x 128 x 88 vessel, 8 seeds): 16.7 ms against 0.09 ms for the union-find (tools/prof_cfg1_stages.py).
    constexpr size_t kCertificateMinCells = (size_t)1 << 26;
    const bool certificate = c->c1_mode == 2 || (c->c1_mode == 0 && grid->n() >= kCertificateMinCells);
    if (certificate && (c->c1_declined == 0 || (++c->c1_declined & 15u) == 0)) {
        int handled = 0;
        uint32_t max_label = 0;
        for (uint32_t i = 0; i < nseeds; ++i) max_label = std::max(max_label, seeds[4 * i + 3]);
        VF_TRY(vf_k_c1_descent(grid, d_seeds, (int)nseeds, max_label, &handled));
        c->c1_declined = handled ? 0u : 1u;
        if (handled) return VF_OK;
    }
    return vf_k_keep_seed_components(grid, d_seeds, (int)nseeds, 0, 6, nullptr);
}


// ================================================================================================ slab-partitioned flood (multi-GPU)
// One very large grid is cut into slabs of the slowest axis x (contiguous in the reference layout; BASELINE.json calls them
// "z-slabs").  Each GPU holds its slab plus one halo plane on either side as an ordinary grid of (xs + 2) x Y x Z
// cells and a key field of the same shape (caller-owned device memory, so the host layer can hand the boundary planes to NCCL).
// Because the key field's fixed point is schedule independent, the protocol is simply: relax the slab to a local fixed point with
// the halo planes held fixed, send the two owned boundary planes to the neighbours, ingest the planes received (lower the halo
// copy, wake the tiles next to it), repeat until an all-reduced change counter is zero.  Collectives stay outside this library:
// vf_flood_slab_boundary_ptr / vf_flood_slab_ingest take and return raw device pointers.
struct vf_slab {
    vf_grid* grid;
    uint32_t* keys;
    Job job;
    int nneigh;
    uint32_t* d_changed;  // ingest counter
    uint32_t* recv = nullptr;            // vf_flood_slab_run: two receive planes + the packed change flags (owned, freed by destroy)
    unsigned long long* d_flags = nullptr;
};

namespace {

__global__ void __launch_bounds__(256) slab_ingest_kernel(uint32_t* __restrict__ halo, const uint32_t* __restrict__ recv, TileGeom g, Worklist wl, int tx,
                                                          uint32_t round, uint32_t* __restrict__ changed, int round_on_device)
{
    if (round_on_device) round = __ldcg(wl.stats + kRoundWord);  // a cooperative phase left the id of the next round in the header
    const size_t n = (size_t)g.Y * g.Z;
    unsigned c = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = recv[i];
        if (r < halo[i]) {
            halo[i] = r;
            ++c;
            atomicMin(&wl.lo[round % 3], r >> KEY_SHIFT);  // the next round's window starts at the lowest level that came in
            const int z = (int)(i % g.Z), y = (int)(i / g.Z);
            const uint32_t tile = ((uint32_t)tx * g.nty + y / TY) * g.ntz + z / TZ;
            if (wl.stamp[tile] != round) {
                wl.occ[tile] = 1;
                enqueue_tile(wl, tile, round);
            }
        }
    }
    c = __reduce_add_sync(kFull, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(changed, c);
}

// labels of the halo planes decide wall / not wall there: the caller uploads the neighbours' boundary label planes with the slab
}  // namespace

extern "C" vf_status vf_flood_slab_init(vf_grid* slab_grid, uint32_t* keys_dev, const uint32_t* seeds_local, uint32_t nseeds, int dfunc, int has_lo, int has_hi,
                                        vf_slab** out)
{
    VF_REQUIRE(slab_grid && keys_dev && out, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = slab_grid->ctx;
    VF_TRY(vf_enter(c));
    VF_REQUIRE(dfunc >= 0 && dfunc <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    VF_REQUIRE(nseeds < 32768, VF_ERR_CAPACITY, "flood: too many seeds");
    VF_REQUIRE(slab_grid->X >= 3, VF_ERR_INVALID_ARGUMENT, "slab needs at least one owned plane between its halo planes");
    vf_slab* s = new vf_slab();
    s->grid = slab_grid;
    s->keys = keys_dev;
    s->nneigh = dfunc == VF_MANHATTAN ? 6 : 26;
    VF_TRY(job_begin(slab_grid, s->job));
    s->job.g.fix_lo = has_lo ? 1 : 0;
    s->job.g.fix_hi = has_hi ? 1 : 0;
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    s->d_changed = (uint32_t*)((char*)c->small.ptr + (704 << 10));
    VF_TRY(launch_init_keys<false>(c, slab_grid->d, keys_dev, s->job.g, s->job.wl.occ, nullptr, s->job.blocks_stream));
    if (nseeds) {
        // seeds_local: {x (slab-local, halo planes included), y, z, GLOBAL order}.  The order goes into the key, labels are
        // looked up at finalize time from the global seed list.
        VF_REQUIRE((size_t)nseeds * sizeof(ushort4) <= 65536, VF_ERR_CAPACITY, "too many seeds in one slab");
        VF_CUDA(vf_sync(c));
        ushort4* h = (ushort4*)c->pinned;
        for (uint32_t i = 0; i < nseeds; ++i) {
            VF_REQUIRE(seeds_local[4 * i] < slab_grid->X && seeds_local[4 * i + 1] < slab_grid->Y && seeds_local[4 * i + 2] < slab_grid->Z, VF_ERR_INVALID_ARGUMENT,
                       "slab seed %u outside the slab", i);
            h[i] = make_ushort4((unsigned short)seeds_local[4 * i], (unsigned short)seeds_local[4 * i + 1], (unsigned short)seeds_local[4 * i + 2],
                                (unsigned short)seeds_local[4 * i + 3]);
        }
        ushort4* d_seeds = (ushort4*)c->small.ptr;
        c->seed_shadow.clear();  // the seed area is overwritten behind vf_upload_seeds' back
        VF_CUDA(cudaMemcpyAsync(d_seeds, h, (size_t)nseeds * sizeof(ushort4), cudaMemcpyHostToDevice, c->stream));
        flood_seed_order_kernel<<<1, 32, 0, c->stream>>>(keys_dev, s->job.g, s->job.wl, d_seeds, (int)nseeds, s->job.round);
        VF_LAUNCHED(c);
    }
    *out = s;
    return VF_OK;
}

extern "C" vf_status vf_flood_slab_relax(vf_slab* s, uint64_t* changed)
{
    VF_REQUIRE(s != nullptr, VF_ERR_INVALID_ARGUMENT, "null slab");
    vf_ctx* c = s->grid->ctx;
    VF_TRY(vf_enter(c));
    VF_CUDA(cudaMemsetAsync(s->job.wl.stats + ST_CHANGED, 0, 4, c->stream));
    VF_TRY(s->nneigh == 6 ? flood_phase<6>(s->job, s->keys) : flood_phase<26>(s->job, s->keys));
    uint32_t hs[8];
    VF_TRY(read_stats(s->job, hs));
    VF_REQUIRE(hs[ST_ERROR] == 0, VF_ERR_CAPACITY, "flood: geodesic distance exceeds the 17-bit key field");
    if (changed) *changed = hs[ST_CHANGED];
    return VF_OK;
}

extern "C" void* vf_flood_slab_boundary_ptr(vf_slab* s, int side)
{
    if (!s) return nullptr;
    const size_t plane = (size_t)s->grid->Y * s->grid->Z;
    return side == 0 ? (void*)(s->keys + plane) : (void*)(s->keys + plane * (s->grid->X - 2));  // owned plane next to the lo / hi halo
}

extern "C" vf_status vf_flood_slab_ingest(vf_slab* s, int side, const uint32_t* plane_dev, uint64_t* changed)
{
    VF_REQUIRE(s && plane_dev, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = s->grid->ctx;
    VF_TRY(vf_enter(c));
    const size_t plane = (size_t)s->grid->Y * s->grid->Z;
    uint32_t* halo = side == 0 ? s->keys : s->keys + plane * (s->grid->X - 1);
    VF_CUDA(cudaMemsetAsync(s->d_changed, 0, 4, c->stream));
    slab_ingest_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(halo, plane_dev, s->job.g, s->job.wl, side == 0 ? 0 : s->job.g.ntx - 1, s->job.round, s->d_changed, 0);
    VF_LAUNCHED(c);
    VF_CUDA(cudaMemcpyAsync(s->job.h_mail, s->d_changed, 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    if (changed) *changed = s->job.h_mail[0];
    return VF_OK;
}

extern "C" vf_status vf_flood_slab_finalize(vf_slab* s, const uint32_t* seeds_global, uint32_t nseeds_total, uint32_t* max_dist)
{
    VF_REQUIRE(s && seeds_global, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = s->grid->ctx;
    VF_TRY(vf_enter(c));
    ushort4* d_seeds = nullptr;
    // only .w is read by the finalize kernel; coordinates are global and may exceed the slab, so upload without range checks
    VF_REQUIRE((size_t)nseeds_total * sizeof(ushort4) <= 65536, VF_ERR_CAPACITY, "too many seeds");
    VF_CUDA(vf_sync(c));
    ushort4* h = (ushort4*)c->pinned;
    for (uint32_t i = 0; i < nseeds_total; ++i) h[i] = make_ushort4(0, 0, 0, (unsigned short)seeds_global[4 * i + 3]);
    d_seeds = (ushort4*)c->small.ptr;
    c->seed_shadow.clear();
    VF_CUDA(cudaMemcpyAsync(d_seeds, h, (size_t)nseeds_total * sizeof(ushort4), cudaMemcpyHostToDevice, c->stream));
    VF_CUDA(cudaMemsetAsync(s->job.wl.stats + ST_MAXDIST, 0, 4, c->stream));
    flood_finalize_kernel<<<s->job.blocks_stream, 256, 0, c->stream>>>(s->keys, s->grid->d, s->grid->n(), d_seeds, 0xFFFFu, s->job.wl.stats);
    VF_LAUNCHED(c);
    uint32_t hs[8];
    VF_TRY(read_stats(s->job, hs));
    if (max_dist) *max_dist = hs[ST_MAXDIST];
    return VF_OK;
}

extern "C" void vf_flood_slab_destroy(vf_slab* s)
{
    if (!s) return;
    if (s->recv) {
        cudaSetDevice(s->grid->ctx->device);
        cudaFree(s->recv);
    }
    delete s;
}

// ------------------------------------------------------------------------------------------------ the exchange loop in C++ over NCCL
namespace {
__global__ void slab_pack_flags_kernel(const uint32_t* __restrict__ stats, const uint32_t* __restrict__ ingested, unsigned long long* __restrict__ flags)
{
    flags[0] = (unsigned long long)stats[ST_CHANGED] + *ingested;  // cells this rank lowered in this iteration (relaxation + ingested halo cells)
    flags[1] = stats[ST_ERROR];
}
}  // namespace

#define VF_NCCL(call)                                                                                                            \
    do {                                                                                                                         \
        ncclResult_t r__ = (call);                                                                                               \
        if (r__ != ncclSuccess) return vf_set_error(VF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl->GetErrorString(r__)); \
    } while (0)

extern "C" vf_status vf_nccl_unique_id(void* id128)
{
    const VfNcclApi* nccl = vf_nccl_api();
    VF_REQUIRE(nccl != nullptr, VF_ERR_UNSUPPORTED, "libnccl.so.2 is not available in this process");
    VF_REQUIRE(id128 != nullptr, VF_ERR_INVALID_ARGUMENT, "null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    VF_NCCL(nccl->GetUniqueId((ncclUniqueId*)id128));
    return VF_OK;
}

extern "C" vf_status vf_nccl_comm_create(vf_ctx* ctx, const void* id128, int world, int rank, void** comm_out)
{
    const VfNcclApi* nccl = vf_nccl_api();
    VF_REQUIRE(nccl != nullptr, VF_ERR_UNSUPPORTED, "libnccl.so.2 is not available in this process");
    VF_REQUIRE(ctx && id128 && comm_out && world >= 1 && rank >= 0 && rank < world, VF_ERR_INVALID_ARGUMENT, "bad communicator arguments");
    VF_TRY(vf_enter(ctx));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    VF_NCCL(nccl->CommInitRank(&comm, world, id, rank));
    *comm_out = comm;
    return VF_OK;
}

extern "C" void vf_nccl_comm_destroy(void* comm)
{
    const VfNcclApi* nccl = vf_nccl_api();
    if (nccl && comm) nccl->CommDestroy((ncclComm_t)comm);
}

// { relax to the local fixed point; send the two owned boundary planes, receive the neighbours' (grouped ncclSend / ncclRecv on the
// context's stream); ingest; all-reduce the change count } until no rank changed a cell.  Everything between two iterations' single host
// wait is enqueued on the stream: the relaxation is one cooperative launch, the round id and the change counters stay on the device.
// comm == NULL with world == 1 runs the loop without NCCL (one slab = the whole grid).
extern "C" vf_status vf_flood_slab_run(vf_slab* s, void* comm_, int rank, int world, uint32_t* iterations, uint64_t* halo_bytes)
{
    VF_REQUIRE(s != nullptr && world >= 1 && rank >= 0 && rank < world, VF_ERR_INVALID_ARGUMENT, "bad slab run arguments");
    const VfNcclApi* nccl = world > 1 ? vf_nccl_api() : nullptr;
    VF_REQUIRE(world == 1 || (nccl != nullptr && comm_ != nullptr), VF_ERR_UNSUPPORTED, "a multi-rank slab run needs NCCL and a communicator");
    ncclComm_t comm = (ncclComm_t)comm_;
    vf_ctx* c = s->grid->ctx;
    VF_TRY(vf_enter(c));
    Job& j = s->job;
    const size_t plane = (size_t)s->grid->Y * s->grid->Z;
    if (!s->recv) {
        VF_CUDA(cudaMalloc(&s->recv, 2 * plane * 4 + 256));
        s->d_flags = (unsigned long long*)(s->recv + 2 * plane);
    }
    const bool lo = rank > 0, hi = rank + 1 < world;  // neighbours; must agree with has_lo / has_hi of vf_flood_slab_init
    uint32_t iters = 0;
    uint64_t moved = 0;
    for (;;) {
        ++iters;
        VF_CUDA(cudaMemsetAsync(j.wl.stats + ST_CHANGED, 0, 4, c->stream));
        VF_CUDA(cudaMemsetAsync(s->d_changed, 0, 4, c->stream));
        VF_TRY(s->nneigh == 6 ? flood_phase<6>(j, s->keys) : flood_phase<26>(j, s->keys));
        if (world > 1) {
            VF_NCCL(nccl->GroupStart());
            if (lo) {
                VF_NCCL(nccl->Send(s->keys + plane, plane, ncclUint32, rank - 1, comm, c->stream));
                VF_NCCL(nccl->Recv(s->recv, plane, ncclUint32, rank - 1, comm, c->stream));
            }
            if (hi) {
                VF_NCCL(nccl->Send(s->keys + plane * (s->grid->X - 2), plane, ncclUint32, rank + 1, comm, c->stream));
                VF_NCCL(nccl->Recv(s->recv + plane, plane, ncclUint32, rank + 1, comm, c->stream));
            }
            VF_NCCL(nccl->GroupEnd());
            const int on_dev = j.round_on_device ? 1 : 0;
            if (lo) {
                slab_ingest_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(s->keys, s->recv, j.g, j.wl, 0, j.round, s->d_changed, on_dev);
                VF_LAUNCHED(c);
                moved += plane * 4;
            }
            if (hi) {
                slab_ingest_kernel<<<c->num_sms * 4, 256, 0, c->stream>>>(s->keys + plane * (s->grid->X - 1), s->recv + plane, j.g, j.wl, j.g.ntx - 1, j.round,
                                                                          s->d_changed, on_dev);
                VF_LAUNCHED(c);
                moved += plane * 4;
            }
        }
        slab_pack_flags_kernel<<<1, 1, 0, c->stream>>>(j.wl.stats, s->d_changed, s->d_flags);
        VF_LAUNCHED(c);
        if (world > 1) VF_NCCL(nccl->AllReduce(s->d_flags, s->d_flags, 2, ncclUint64, ncclSum, comm, c->stream));
        unsigned long long* h_flags = (unsigned long long*)((char*)j.h_mail + 256);
        VF_CUDA(cudaMemcpyAsync(h_flags, s->d_flags, 16, cudaMemcpyDeviceToHost, c->stream));
        uint32_t hs[8];
        VF_TRY(read_stats(j, hs));  // the iteration's one host wait; also brings the round id back
        VF_REQUIRE(h_flags[1] == 0, VF_ERR_CAPACITY, "flood: geodesic distance exceeds the 17-bit key field");
        if (h_flags[0] == 0) break;
        VF_REQUIRE(iters < 100000, VF_ERR_CAPACITY, "slab exchange did not converge");
    }
    if (iterations) *iterations = iters;
    if (halo_bytes) *halo_bytes = moved;
    return VF_OK;
}

#ifdef VF_FLOOD_TIMING
extern "C" void vf_debug_flood_cycles(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_flood_cycles, sizeof(g_flood_cycles));
    if (reset) {
        unsigned long long z[8] = { 0 };
        cudaMemcpyToSymbol(g_flood_cycles, z, sizeof(z));
    }
}
extern "C" void vf_debug_front_cycles(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_front_cycles, sizeof(g_front_cycles));
    if (reset) {
        unsigned long long z[12] = { 0 };
        cudaMemcpyToSymbol(g_front_cycles, z, sizeof(z));
    }
}
#endif
