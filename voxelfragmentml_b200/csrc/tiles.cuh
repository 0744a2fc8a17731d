// tiles.cuh — tile worklist machinery shared by the flood (F2/F3) and connected-to-seed (C1) kernels.
//
// The grid is cut into TX x TY x TZ = 16 x 16 x 32 tiles (TZ = 32 so that one z-row of a tile is one 32-bit mask word and
// one 128-byte line of uint32 keys).  A CTA of 256 threads stages one tile plus its 1-cell halo in shared memory, iterates
// it to a local fixed point there (one __syncthreads per wavefront step instead of one kernel launch + host readback per
// BFS level as in FloodFracturer.cpp:143-158), writes back the rows that changed and enqueues the neighbour tiles whose halo
// it changed.  Rounds of the global worklist run back to back without host synchronisation; the host only reads one
// counter per batch of rounds.
#pragma once

#include "vf_internal.h"

namespace vft {

constexpr int TX = 16, TY = 16, TZ = 32;
constexpr int kThreads = TX * TY;              // one thread per z-row in the thread-per-row phases
constexpr int SY = TY + 2;                     // rows incl. halo
constexpr int ZPAD = 4;                        // staged rows start 4 keys before the tile: a TMA box must start on a 16-byte boundary
constexpr int SZ = TZ + 2 * ZPAD;              // 160-byte rows = the inner extent of the TMA box; z = -1 sits at slot 3, z = TZ at slot 36
constexpr int kCells = (TX + 2) * SY * SZ;     // shared-memory words per tile
constexpr unsigned kFull = 0xFFFFFFFFu;

struct TileGeom {
    int X, Y, Z;        // grid dims
    int ntx, nty, ntz;  // tiles per axis
    int fix_lo, fix_hi; // slab mode: plane x = 0 / x = X-1 is a halo copy owned by a neighbour GPU: readable, never relaxed here
    __host__ __device__ int ntiles() const { return ntx * nty * ntz; }
};

inline TileGeom make_geom(uint32_t X, uint32_t Y, uint32_t Z)
{
    TileGeom g;
    g.X = (int)X, g.Y = (int)Y, g.Z = (int)Z;
    g.ntx = (g.X + TX - 1) / TX, g.nty = (g.Y + TY - 1) / TY, g.ntz = (g.Z + TZ - 1) / TZ;
    g.fix_lo = g.fix_hi = 0;
    return g;
}

// worklist state in device memory (ctx->tiles scratch)
struct Worklist {
    uint32_t* list[2];   // ping-pong tile lists
    uint32_t* count;     // [3] rotating counters: round r reads count[r%3], appends to count[(r+1)%3], zeroes count[(r+2)%3]
    uint32_t* stamp;     // [ntiles] round id at which the tile was last enqueued (dedupe)
    uint8_t* occ;        // [ntiles] tile holds at least one non-wall cell
    uint32_t* stats;     // [8]: 0 tile visits, 1 non-empty rounds, 2 error flags, 3 freed voxels, 4 max dist, 5 scratch
    uint8_t* seen;       // [ntiles] flood: phase id (epoch) of the tile's last relaxation visit | 0x80 when candidates were deferred
    uint32_t epoch;      // flood: current phase id (1, or 2 for the re-flood of F3)
    uint32_t* lo;        // [3] rotating like `count`: lowest distance level among the candidates deferred to the next round
    uint32_t* pend;      // [ntiles][kThreads] flood: per-row masks of the deferred candidate cells of a tile
    uint32_t levels;     // flood: width of a round's distance window
};

#ifdef __CUDACC__
__device__ __forceinline__ int sidx(int x, int y, int z)  // x,y,z in [-1, T]
{
    return ((x + 1) * SY + (y + 1)) * SZ + (z + ZPAD);
}

__device__ __forceinline__ void enqueue_tile(const Worklist& wl, uint32_t tile, uint32_t round)
{
    if (atomicExch(&wl.stamp[tile], round) != round) {
        const uint32_t i = atomicAdd(&wl.count[round % 3], 1u);
        wl.list[round & 1][i] = tile;
    }
}

// After a tile converged: from the per-row change masks derive which of the 26 neighbour tiles saw their halo change and
// enqueue them for `next_round`.  Must be called by all kThreads threads; `s_nbmask` is a shared word zeroed beforehand.
template <int NNEIGH>
__device__ __forceinline__ void enqueue_neighbours(const TileGeom& g, const Worklist& wl, int tx, int ty, int tz, unsigned rowchg,
                                                   unsigned* s_nbmask, uint32_t next_round)
{
    const int t = threadIdx.x, x = t / TY, y = t % TY;
    unsigned m = 0;
    if (rowchg) {
        const unsigned zm = (rowchg & 1u ? 1u : 0u) | 2u | (rowchg >> 31 ? 4u : 0u);
        const unsigned xm = (x == 0 ? 1u : 0u) | 2u | (x == TX - 1 ? 4u : 0u);
        const unsigned ym = (y == 0 ? 1u : 0u) | 2u | (y == TY - 1 ? 4u : 0u);
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    if ((xm >> a & 1u) && (ym >> b & 1u) && (zm >> c & 1u)) m |= 1u << ((a * 3 + b) * 3 + c);
    }
    m = __reduce_or_sync(kFull, m);
    if ((t & 31) == 0 && m) atomicOr(s_nbmask, m);
    __syncthreads();
    if (t < 27 && t != 13 && (*s_nbmask >> t & 1u)) {
        const int dx = t / 9 - 1, dy = (t / 3) % 3 - 1, dz = t % 3 - 1;
        const int nz = (dx != 0) + (dy != 0) + (dz != 0);
        if (NNEIGH == 26 || nz == 1) {
            const int ax = tx + dx, ay = ty + dy, az = tz + dz;
            if (ax >= 0 && ay >= 0 && az >= 0 && ax < g.ntx && ay < g.nty && az < g.ntz) {
                const uint32_t nt = ((uint32_t)ax * g.nty + ay) * g.ntz + az;
                if (wl.occ[nt]) enqueue_tile(wl, nt, next_round);
            }
        }
    }
}
#endif

}  // namespace vft
