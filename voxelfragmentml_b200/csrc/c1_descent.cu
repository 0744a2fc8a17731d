// c1_descent.cu — C1 (connected-to-seed cleanup) as a "descent certificate": one streaming pass over the labels plus list work on the few
// cells the pass cannot certify.  This is the default path of vf_remove_isolated_regions; the union-find of ccl.cu takes over when the
// certificate declines (see the end of this comment).
//
// Semantics (NaiveFracturer::removeIsolatedRegionsCPU, SRC/Fracturer/NaiveFracturer.cpp:111-150): every seed cell is overwritten with its
// seed's label (a later seed on the same cell wins), then only cells 6-connected to their own seed's cell through same-label cells survive
// (the seed's cell is the start even when a later seed took it); everything else, FREE cells included, becomes EMPTY.
//
// Idea: a labelled cell that has a same-label 6-neighbour ONE MANHATTAN STEP CLOSER to its own seed ("descent neighbour") is connected to the
// seed if that neighbour is, and the distance strictly decreases, so a cell with a descent neighbour outside the set D below is connected.
//   F = labelled cells, not seeds, without any descent neighbour                      (one streaming pass, 2 B read per voxel)
//   D = least set that holds F and every cell ALL of whose descent neighbours are in D (frontier propagation away from the seeds)
//   kept = cells outside D  +  cells of D reachable inside D from a D-cell that touches a same-label cell outside D
// On the cfg3 grid (512^3, 64 Voronoi regions) F and D are a few dozen cells: the union-find over all runs of the grid is replaced by one
// pass plus list work on D.  On thin curved shells F runs into the 10^4s; the pass notices (kBailF), stops and leaves the grid to the union-find.
//
// The pass: a warp walks 32 consecutive 8-cell chunks of a z-row (one 128-bit load per lane, the next item's load already in flight).  The
// z-neighbours of a chunk's end cells come from the adjacent lanes by shuffle; a chunk that holds one label certifies seven of its cells by
// construction (their descent neighbour is inside the chunk) and the eighth with the shuffled neighbour, so a solid region costs no extra
// load at all.  Only cells whose z-neighbour towards the seed carries another label look at their y / x neighbours in global memory.
//
// D lives in a list + an open-addressing hash set (cell -> list position), both sized for 65 536 cells.  Anything unusual — lists that
// outgrow that, two seeds with one label, rows that are not 16-byte aligned, more than 4096 seeds — reports "not handled" and the caller
// runs ccl.cu on the grid (which by then carries the planted seed labels and has lost its FREE cells, both of which ccl.cu does as well).
#include <algorithm>
#include <cstring>

#include "vf_internal.h"

namespace vfc1 {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kCap = 1u << 16;       // list capacity in cells
constexpr uint32_t kBailF = 4096;         // more uncertified cells than this: the pass stops and the union-find takes the grid (thin shells)
constexpr uint32_t kHashBits = 18;        // hash set: 4 x the list capacity
constexpr uint32_t kHashSize = 1u << kHashBits;
constexpr uint32_t kMaxSeeds = 4096;      // seed positions live in shared memory (32 KB)
constexpr int kResolveThreads = 1024;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Dims {
    int X, Y, Z;
};
struct Ctl {  // device control block
    uint32_t tail, overflow, freed, dup, pad[4];
};
struct Set {
    uint32_t* keys;  // [kHashSize] cell index or kNone
    uint32_t* vals;  // [kHashSize] position in the list
};

__device__ __forceinline__ uint32_t cell(const Dims& d, int x, int y, int z) { return ((uint32_t)x * d.Y + y) * d.Z + z; }
__device__ __forceinline__ uint32_t hash_of(uint32_t c) { return (c * 2654435761u) >> (32 - kHashBits); }

// position of `c` in the list, kNone when it is not in D.  The set never fills up (inserts stop at kCap entries).
__device__ __forceinline__ uint32_t set_find(const Set& s, uint32_t c)
{
    for (uint32_t h = hash_of(c);; h = (h + 1) & (kHashSize - 1)) {
        const uint32_t k = s.keys[h];
        if (k == c) return h;
        if (k == kNone) return kNone;
    }
}
__device__ __forceinline__ bool in_set(const Set& s, uint32_t c) { return set_find(s, c) != kNone; }

// insert `c`; returns its slot when this call inserted it, kNone when it was there already
__device__ __forceinline__ uint32_t set_insert(const Set& s, uint32_t c)
{
    for (uint32_t h = hash_of(c);; h = (h + 1) & (kHashSize - 1)) {
        const uint32_t k = atomicCAS(&s.keys[h], kNone, c);
        if (k == kNone) return h;
        if (k == c) return kNone;
    }
}

__global__ void fill_kernel(uint32_t* p, uint32_t words, uint32_t value, uint32_t* zeros, uint32_t zero_words)
{
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = t0; i < words; i += stride) p[i] = value;
    for (uint32_t i = t0; i < zero_words; i += stride) zeros[i] = 0;
}

// seed s registers as the start of its label, and plants the label unless a later seed sits on the same cell (NaiveFracturer.cpp:120-123)
__global__ void plant_kernel(uint16_t* __restrict__ grid, Dims d, const ushort4* __restrict__ seeds, int S, uint32_t* __restrict__ table, Ctl* ctl)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const ushort4 sd = seeds[s];
    // every seed is a start for its own label, also one whose cell a later seed takes (the search of NaiveFracturer.cpp:116-146 is seeded
    // with all seeds and enters the neighbours that hold the label the front element carries)
    const uint32_t old = atomicCAS(&table[sd.w], kNone, (uint32_t)s);
    if (old != kNone) ctl->dup = 1;  // two seeds with one label: "own seed" is ambiguous, the union-find decides
    for (int t = s + 1; t < S; ++t)
        if (seeds[t].x == sd.x && seeds[t].y == sd.y && seeds[t].z == sd.z) return;
    grid[cell(d, sd.x, sd.y, sd.z)] = sd.w;
}

__device__ __forceinline__ void append_f(uint32_t i, uint32_t* __restrict__ list, Ctl* ctl)
{
    const uint32_t j = atomicAdd(&ctl->tail, 1u);
    if (j < kCap) list[j] = i;
    else ctl->overflow = 1;
}

// pass 1: FREE -> EMPTY, and F (cells without a descent neighbour) into the list.
//
// A warp owns a contiguous range of items (item = 32 consecutive chunks of one z-row), so (x, y, segment) advance by carry instead of by
// division.  A chunk is examined run by run (maximal stretches of one word inside the chunk; one run when the chunk is uniform): every cell of a
// run but the one nearest to the seed's z has its descent neighbour inside the run, so a run costs ONE examination — of the cell
// zc = clamp(seed.z, run) — whatever its length.  zc's z-neighbour towards the seed lies outside the run: in the adjacent chunk when the run
// touches the chunk's end (then it may carry the same label), otherwise it is the next run and differs.
//
// Most chunks lie inside a region: one label, the same label on either side, the seed's z elsewhere.  A lane recognises that in a dozen
// instructions (the seed's z of the label it saw last is cached in a register) and is done.  The other chunks — region borders, the chunk
// of each row that holds the seed's z — are queued in shared memory and examined 32 at a time, one chunk per lane, so that the run loop and
// the y / x look-ups in global memory run with full warps instead of stalling the warp for one lane.
// TSM: label -> seed table in shared memory (all seed labels < kSmemLabels), else in global memory.
constexpr uint32_t kSmemLabels = 4096;
constexpr int kCertWarps = 8, kQueue = 64;

struct QEntry {
    uint4 v;        // the chunk
    uint32_t base;  // linear index of its first cell
    uint32_t xy;    // x << 16 | y
    uint32_t nb;    // cell before the chunk | cell after it << 16 (EMPTY outside the row)
    uint32_t pad;
};

template <bool TSM>
__global__ void __launch_bounds__(kCertWarps * 32) certificate_kernel(uint16_t* grid, Dims d, const ushort4* __restrict__ seeds, int S,
                                                                      const uint32_t* __restrict__ table, uint32_t* __restrict__ list, Ctl* ctl)
{
    extern __shared__ ushort4 sp[];
    __shared__ uint16_t stab[TSM ? kSmemLabels : 1];
    __shared__ QEntry queue[kCertWarps][kQueue];
    for (int i = threadIdx.x; i < S; i += blockDim.x) sp[i] = seeds[i];
    if (TSM)
        for (int i = threadIdx.x; i < (int)kSmemLabels; i += blockDim.x) stab[i] = (uint16_t)min(table[i], 0xFFFFu);  // seed indices are < 4096
    __syncthreads();
    auto seed_of = [&](uint32_t L) -> uint32_t {
        if (TSM) {
            const uint32_t s = L < kSmemLabels ? stab[L] : 0xFFFFu;
            return s == 0xFFFFu ? kNone : s;
        }
        return table[L];
    };
    const int lane = threadIdx.x & 31;
    QEntry* q = queue[threadIdx.x >> 5];
    int qn = 0;
    const uint32_t Zu = (uint32_t)d.Z, YZ = (uint32_t)d.Y * Zu;

    // one queued chunk: every run of it (see above)
    auto examine = [&](const QEntry& e) {
        const uint4 v = e.v;
        const uint32_t base = e.base, prev_last = e.nb & 0xFFFFu, next_first = e.nb >> 16;
        const int x = (int)(e.xy >> 16), y = (int)(e.xy & 0xFFFFu), z0 = (int)(base - ((uint32_t)x * d.Y + y) * Zu);
        const uint32_t one = 0x00010001u;  // run starts: bit k set <=> cell k differs from cell k - 1
        const uint32_t n0 = __vminu2(v.x ^ __byte_perm(v.x, 0, 0x1010), one), n1 = __vminu2(v.y ^ __byte_perm(v.x, v.y, 0x5432), one);
        const uint32_t n2 = __vminu2(v.z ^ __byte_perm(v.y, v.z, 0x5432), one), n3 = __vminu2(v.w ^ __byte_perm(v.z, v.w, 0x5432), one);
        const uint32_t bb = n0 | n1 << 2 | n2 << 4 | n3 << 6;
        const unsigned long long lo = (unsigned long long)v.y << 32 | v.x, hi = (unsigned long long)v.w << 32 | v.z;
        bool has_free = false;
        for (uint32_t st = ((bb | bb >> 15) & 0xFFu) | 1u; st;) {
            const int a = __ffs(st) - 1;
            st &= st - 1;
            const int b = (st ? __ffs(st) - 1 : 8) - 1;  // the run covers cells a..b of the chunk
            const uint32_t L = (uint32_t)((a < 4 ? lo >> (16 * a) : hi >> (16 * (a - 4))) & 0xFFFFu);
            if (L <= VF_VOXEL_FREE) {
                has_free = has_free || L == VF_VOXEL_FREE;
                continue;
            }
            const uint32_t s = seed_of(L);
            if (s == kNone) {  // a label without a seed: nothing of it survives
                for (int k = a; k <= b; ++k) append_f(base + k, list, ctl);
                continue;
            }
            const ushort4 p = sp[s];
            const int zc = min(max((int)p.z, z0 + a), z0 + b);
            if (x == p.x && y == p.y && zc == p.z) continue;                                   // the seed's own cell
            if (abs(x - (int)p.x) + abs(y - (int)p.y) + abs(zc - (int)p.z) == 1) continue;  // entered from the start itself
            if (zc > p.z && a == 0 && prev_last == L) continue;
            if (zc < p.z && b == 7 && next_first == L) continue;
            // y / x neighbour one step closer to the seed (between the cell and the seed: inside the grid).  The y neighbour lives in a row this
            // warp has just streamed (an L2 hit); the x neighbour is a plane away — a 32-byte sector from DRAM for 2 bytes — and is only
            // looked at when y does not settle it
            const uint32_t u = base + (uint32_t)(zc - z0);
            if (y != p.y && grid[y > p.y ? u - Zu : u + Zu] == L) continue;
            if (x != p.x && grid[x > p.x ? u - YZ : u + YZ] == L) continue;
            append_f(u, list, ctl);
        }
        if (has_free) {
            // dropped: the reference rebuilds the grid from an all-EMPTY one.  Neighbours may read this chunk while it is rewritten:
            // FREE and EMPTY both differ from every label, so their verdicts do not change.
            auto drop = [](uint32_t w) {
                if ((w & 0xFFFFu) == VF_VOXEL_FREE) w &= 0xFFFF0000u;
                if ((w >> 16) == VF_VOXEL_FREE) w &= 0x0000FFFFu;
                return w;
            };
            *reinterpret_cast<uint4*>(grid + base) = make_uint4(drop(v.x), drop(v.y), drop(v.z), drop(v.w));
        }
    };

    const uint32_t cpr = Zu / 8, segs = (cpr + 31) / 32, items = (uint32_t)d.X * d.Y * segs;
    const uint32_t nwarps = gridDim.x * (blockDim.x / 32), wid = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const uint32_t ipw = (items + nwarps - 1) / nwarps, begin = wid * ipw, end = min(items, begin + ipw);
    if (begin >= end) return;
    uint32_t row = begin / segs, seg = begin - row * segs;
    uint32_t y = row % (uint32_t)d.Y, x = row / (uint32_t)d.Y;
    auto load = [&](uint32_t r, uint32_t sg) -> uint4 {
        const uint32_t ch = sg * 32 + lane;
        if (ch >= cpr) return make_uint4(0u, 0u, 0u, 0u);
        return __ldcg(reinterpret_cast<const uint4*>(grid + (size_t)r * Zu + ch * 8));  // coherent: seeds were planted by the previous kernel
    };
    uint32_t cL = 0;      // label this lane saw last in a uniform chunk ...
    int cz = -1 << 20;    // ... and its seed's z (none: far away)
    uint4 vnext = load(row, seg);
    for (uint32_t item = begin; item < end; ++item) {
        // a grid the certificate does not suit (thin curved shells: the straight way to the seed leaves the label) is recognised early: once
        // more than kBailF cells are listed every warp stops, the list work is skipped and the caller runs the union-find
        if ((item & 7u) == 0 && *(volatile uint32_t*)&ctl->tail > kBailF) {
            ctl->overflow = 1;
            return;
        }
        const uint4 v = vnext;
        uint32_t nrow = row, nseg = seg + 1;
        if (nseg == segs) nseg = 0, ++nrow;
        if (item + 1 < end) vnext = load(nrow, nseg);
        const uint32_t ch = seg * 32 + lane;
        const bool active = ch < cpr;
        const uint32_t base = row * Zu + ch * 8;
        // the cells next to the chunk's ends; EMPTY (never equal to a label) outside the row
        uint32_t prev_last = __shfl_up_sync(kFull, v.w >> 16, 1), next_first = __shfl_down_sync(kFull, v.x & 0xFFFFu, 1);
        if (lane == 0) prev_last = (active && ch > 0) ? grid[base - 1] : 0u;
        if (lane == 31) next_first = (active && ch + 1 < cpr) ? grid[base + 8] : 0u;
        const uint32_t first = v.x & 0xFFFFu;
        const bool uni = first > VF_VOXEL_FREE && v.x == first * 0x10001u && v.y == v.x && v.z == v.x && v.w == v.x;
        if (uni && first != cL) {
            const uint32_t s = seed_of(first);
            cL = first, cz = s == kNone ? -1 << 20 : (int)sp[s].z;
            if (s == kNone) cL = 0;  // never matches: the chunk goes to the queue
        }
        const int z0 = (int)ch * 8;
        const bool quiet = uni && first == cL && prev_last == first && next_first == first && (cz < z0 || cz > z0 + 7);
        const bool slow = active && !quiet && (v.x | v.y | v.z | v.w) != 0u;
        const unsigned m = __ballot_sync(kFull, slow);
        if (m) {
            if (slow) {
                QEntry& e = q[qn + __popc(m & ((1u << lane) - 1u))];
                e.v = v, e.base = base, e.xy = x << 16 | y, e.nb = prev_last | next_first << 16;
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                examine(q[lane]);
                __syncwarp();
                qn -= 32;
                if (lane < qn) {
                    const QEntry t = q[32 + lane];
                    q[lane] = t;
                }
                __syncwarp();
            }
        }
        row = nrow, seg = nseg;
        if (seg == 0 && ++y == (uint32_t)d.Y) y = 0, ++x;
    }
    if (lane < qn) examine(q[lane]);
}

// pass 2 (one CTA; D is small): closure of F, liveness inside D, removal.  Leaves at once when the list is empty.
__device__ __forceinline__ void resolve_cells(uint16_t* grid, Dims d, const ushort4* __restrict__ seeds, int S, const uint32_t* __restrict__ table, Set set, uint32_t* alive,
                                              uint32_t* list, Ctl* ctl)
{
    extern __shared__ ushort4 sp[];
    __shared__ uint32_t s_head, s_tail, s_flag;
    const int t = threadIdx.x;
    if (t == 0) {
        if (ctl->tail > kBailF) ctl->overflow = 1;  // the pass stopped, or ended, with more cells than the list work is meant for
        s_head = 0, s_tail = min(ctl->tail, kCap), s_flag = ctl->overflow | ctl->dup;
    }
    __syncthreads();
    if (s_flag || s_tail == 0) return;
    for (int i = t; i < S; i += blockDim.x) sp[i] = seeds[i];
    for (uint32_t i = t; i < s_tail; i += blockDim.x) {  // the cells of F are distinct
        const uint32_t h = set_insert(set, list[i]);
        if (h != kNone) set.vals[h] = i;
    }
    __syncthreads();
    const int dx[6] = { 1, -1, 0, 0, 0, 0 }, dy[6] = { 0, 0, 1, -1, 0, 0 }, dz[6] = { 0, 0, 0, 0, 1, -1 };
    auto decode = [&](uint32_t u, int& x, int& y, int& z) {
        z = (int)(u % (uint32_t)d.Z);
        const uint32_t r = u / (uint32_t)d.Z;
        y = (int)(r % (uint32_t)d.Y), x = (int)(r / (uint32_t)d.Y);
    };
    auto inside = [&](int x, int y, int z) { return (unsigned)x < (unsigned)d.X && (unsigned)y < (unsigned)d.Y && (unsigned)z < (unsigned)d.Z; };
    auto man = [](int x, int y, int z, const ushort4& p) { return abs(x - (int)p.x) + abs(y - (int)p.y) + abs(z - (int)p.z); };

    // ---- closure: a cell all of whose descent neighbours are in D joins D; examined when one of them joins
    for (;;) {
        const uint32_t head = s_head, tail = s_tail;
        __syncthreads();
        if (head == tail) break;
        for (uint32_t i = head + t; i < tail; i += blockDim.x) {
            const uint32_t u = list[i];
            const uint32_t L = grid[u];
            const uint32_t s = table[L];
            if (s == kNone) continue;  // a label without a seed: all of its cells are in F already
            const ushort4 p = sp[s];
            int x, y, z;
            decode(u, x, y, z);
            const int du = man(x, y, z, p);
            for (int k = 0; k < 6; ++k) {
                const int vx = x + dx[k], vy = y + dy[k], vz = z + dz[k];
                if (!inside(vx, vy, vz)) continue;
                const uint32_t v = cell(d, vx, vy, vz);
                if (grid[v] != L || man(vx, vy, vz, p) != du + 1 || in_set(set, v)) continue;
                bool all_dead = true;  // v is one step farther than u, so it is not the seed and differs from p on at least one axis
                if (vz != p.z) {
                    const uint32_t n = cell(d, vx, vy, vz + (vz > p.z ? -1 : 1));
                    all_dead = all_dead && !(grid[n] == L && !in_set(set, n));
                }
                if (all_dead && vy != p.y) {
                    const uint32_t n = cell(d, vx, vy + (vy > p.y ? -1 : 1), vz);
                    all_dead = all_dead && !(grid[n] == L && !in_set(set, n));
                }
                if (all_dead && vx != p.x) {
                    const uint32_t n = cell(d, vx + (vx > p.x ? -1 : 1), vy, vz);
                    all_dead = all_dead && !(grid[n] == L && !in_set(set, n));
                }
                if (all_dead) {
                    const uint32_t h = set_insert(set, v);
                    if (h != kNone) {
                        const uint32_t j = atomicAdd(&ctl->tail, 1u);
                        if (j < kCap) list[j] = v, set.vals[h] = j;
                        else ctl->overflow = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (t == 0) s_head = tail, s_tail = min(ctl->tail, kCap), s_flag = ctl->overflow;
        __syncthreads();
        if (s_flag) return;  // the caller falls back to the union-find; nothing has been removed yet
    }
    const uint32_t nD = s_tail;

    // ---- liveness: a D-cell next to a same-label cell outside D is connected (cells outside D are); liveness spreads inside D
    for (uint32_t i = t; i < nD; i += blockDim.x) {
        const uint32_t u = list[i];
        const uint32_t L = grid[u];
        int x, y, z;
        decode(u, x, y, z);
        uint32_t a = 0;
        for (int k = 0; k < 6 && !a; ++k) {
            const int vx = x + dx[k], vy = y + dy[k], vz = z + dz[k];
            if (!inside(vx, vy, vz)) continue;
            const uint32_t v = cell(d, vx, vy, vz);
            if (grid[v] == L && !in_set(set, v)) a = 1;
        }
        alive[i] = a;
    }
    for (;;) {
        __syncthreads();
        if (t == 0) s_flag = 0;
        __syncthreads();
        for (uint32_t i = t; i < nD; i += blockDim.x) {
            if (alive[i]) continue;
            const uint32_t u = list[i];
            const uint32_t L = grid[u];
            int x, y, z;
            decode(u, x, y, z);
            for (int k = 0; k < 6; ++k) {
                const int vx = x + dx[k], vy = y + dy[k], vz = z + dz[k];
                if (!inside(vx, vy, vz)) continue;
                const uint32_t v = cell(d, vx, vy, vz);
                if (grid[v] != L) continue;
                const uint32_t h = set_find(set, v);
                if (h != kNone && ((volatile uint32_t*)alive)[set.vals[h]]) {
                    alive[i] = 1;
                    s_flag = 1;
                    break;
                }
            }
        }
        __syncthreads();
        const uint32_t changed = s_flag;
        if (!changed) break;
    }

    // ---- removal.  Labels are read above and written only here, after the last barrier of the loop.
    uint32_t freed = 0;
    for (uint32_t i = t; i < nD; i += blockDim.x) {
        if (!alive[i]) {
            grid[list[i]] = VF_VOXEL_EMPTY;
            ++freed;
        }
    }
    if (freed) atomicAdd(&ctl->freed, freed);
}

// The list work, then the verdict of the certificate straight into (mapped) pinned host memory when `host_word` is given: the host polls the
// word instead of going through a copy + a stream synchronisation (two driver calls and their wake-up latency, which grows when several
// processes drive GPUs from one box).
__global__ void __launch_bounds__(kResolveThreads) resolve_kernel(uint16_t* grid, Dims d, const ushort4* __restrict__ seeds, int S, const uint32_t* __restrict__ table,
                                                                  Set set, uint32_t* alive, uint32_t* list, Ctl* ctl, volatile uint32_t* host_word)
{
    resolve_cells(grid, d, seeds, S, table, set, alive, list, ctl);
    if (host_word == nullptr) return;
    __syncthreads();  // the removals and the flags of every thread are in place
    if (threadIdx.x == 0) {
        __threadfence_system();  // the grid before the word: the host launches the next stage as soon as it reads it
        const volatile Ctl* vc = ctl;
        *host_word = 1u | (vc->overflow ? 2u : 0u) | (vc->dup ? 4u : 0u);
        __threadfence_system();
    }
}

}  // namespace vfc1

// Returns VF_OK with *handled = 1 when the grid now holds C1's result; *handled = 0 when the caller must run the union-find (the grid then
// carries the planted seed labels and has lost its FREE cells, both of which the union-find path does as well).
vf_status vf_k_c1_descent(vf_grid* grid, const ushort4* d_seeds, int nseeds, uint32_t max_label, int* handled)
{
    using namespace vfc1;
    *handled = 0;
    vf_ctx* c = grid->ctx;
    const size_t n = grid->n();
    if (grid->Z % 8 != 0 || ((uintptr_t)grid->d & 15) != 0 || n >= 0xFFFFFFFFull || nseeds < 1 || (uint32_t)nseeds > kMaxSeeds) return VF_OK;
    Dims d = { (int)grid->X, (int)grid->Y, (int)grid->Z };
    // scratch (the flood key arena): label -> seed table | hash keys | hash values | liveness flags | control block | list
    const size_t words = 65536 + 2 * (size_t)kHashSize + kCap + 64 + kCap;
    VF_TRY(vf_scratch_reserve(c, c->keys, words * 4));
    uint32_t* table = (uint32_t*)c->keys.ptr;
    Set set = { table + 65536, table + 65536 + kHashSize };
    uint32_t* alive = set.vals + kHashSize;
    Ctl* ctl = (Ctl*)(alive + kCap);
    uint32_t* list = (uint32_t*)ctl + 64;
    fill_kernel<<<c->num_sms, 256, 0, c->stream>>>(table, 65536 + kHashSize, kNone, (uint32_t*)ctl, 64);  // table + hash keys <- none; control block <- 0
    VF_LAUNCHED(c);
    plant_kernel<<<(nseeds + 127) / 128, 128, 0, c->stream>>>(grid->d, d, d_seeds, nseeds, table, ctl);
    VF_LAUNCHED(c);
    const size_t smem = (size_t)nseeds * sizeof(ushort4);
    // one resident wave: every warp owns a contiguous share of the rows
    auto kern = max_label < kSmemLabels ? certificate_kernel<true> : certificate_kernel<false>;
    int per_sm = 0;
    VF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCertWarps * 32, smem));
    kern<<<c->num_sms * std::max(per_sm, 1), kCertWarps * 32, smem, c->stream>>>(grid->d, d, d_seeds, nseeds, table, list, ctl);
    VF_LAUNCHED(c);
    volatile uint32_t* h_word = (volatile uint32_t*)((char*)c->pinned + 65536 + 256);
    *h_word = 0;
    resolve_kernel<<<1, kResolveThreads, smem, c->stream>>>(grid->d, d, d_seeds, nseeds, table, set, alive, list, ctl, c->blocking_sync ? nullptr : h_word);
    VF_LAUNCHED(c);
    if (c->blocking_sync) {  // producers that share cores: sleep on the event instead of polling
        Ctl* h = (Ctl*)((char*)c->pinned + 65536 + 320);
        VF_CUDA(cudaMemcpyAsync(h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream));
        VF_CUDA(vf_sync(c));
        if (h->overflow || h->dup) return VF_OK;  // too much for one CTA's list work, or ambiguous starts: the union-find is the right tool
        *handled = 1;
        return VF_OK;
    }
    uint32_t verdict = 0;
    for (uint64_t spins = 0; (verdict = *h_word) == 0; ++spins) {
        if ((spins & 0xFFFFu) == 0xFFFFu && cudaStreamQuery(c->stream) != cudaErrorNotReady) {  // the stream drained or failed: stop polling
            VF_CUDA(vf_sync(c));
            verdict = *h_word;
            VF_REQUIRE(verdict != 0, VF_ERR_CUDA, "C1: the certificate's verdict never arrived");
            break;
        }
        if (c->yield_wait) sched_yield();
#if defined(__x86_64__)
        else __builtin_ia32_pause();
#endif
    }
    if (verdict & 6u) return VF_OK;  // too much for one CTA's list work, or ambiguous starts: the union-find is the right tool
    *handled = 1;
    return VF_OK;
}
