// c1_descent.cu — EXPERIMENTAL alternative for C1 (connected-to-seed cleanup), selected at run time with VF_C1_DESCENT=1; the default is
// the union-find of ccl.cu.  Not yet run on a GPU: the algorithm is validated on the CPU (tools/c1_descent_prototype.py: identical to the
// CPU checker on dense Voronoi labels under all three metrics, the reference's vessel grid and porous blobs), the kernels are not.
//
// Semantics (NaiveFracturer::removeIsolatedRegionsCPU, SRC/Fracturer/NaiveFracturer.cpp:111-150): every seed cell is overwritten with its
// seed's label (a later seed on the same cell wins), then only cells 6-connected to their own seed's cell through same-label cells survive
// (the seed's cell is the start even when a later seed took it); everything else, FREE cells included, becomes EMPTY.
//
// Idea: a labelled cell that has a same-label 6-neighbour ONE MANHATTAN STEP CLOSER to its own seed ("descent neighbour") is connected to the
// seed if that neighbour is, and the distance strictly decreases, so a cell with a descent neighbour outside the set D below is connected.
//   F = labelled cells, not seeds, without any descent neighbour                      (one streaming pass with a 6-point stencil)
//   D = least set that holds F and every cell ALL of whose descent neighbours are in D (frontier propagation away from the seeds)
//   kept = cells outside D  +  cells of D reachable inside D from a D-cell that touches a same-label cell outside D
// On the cfg3 grid (512^3, 64 Voronoi regions) F and D are a few dozen cells, on the reference's vessel shell ~10^4: the union-find over
// all runs of the grid is replaced by one pass plus list work on D.  Anything unusual — lists that outgrow their capacity, two surviving
// seeds with one label, rows that are not 16-byte aligned, more than 1024 seeds — reports "not handled" and the caller runs ccl.cu.
#include <algorithm>
#include <cstring>

#include "vf_internal.h"

namespace vfc1 {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr uint32_t kCap = 1u << 20;       // list capacity in cells
constexpr uint32_t kMaxSeeds = 1024;      // seed positions live in shared memory
constexpr int kResolveThreads = 1024;

struct Dims {
    int X, Y, Z;
};
struct Ctl {  // device control block
    uint32_t tail, overflow, freed, dup, pad[4];
};

__device__ __forceinline__ bool bit(const uint32_t* b, uint32_t i) { return (b[i >> 5] >> (i & 31u)) & 1u; }
__device__ __forceinline__ uint32_t cell(const Dims& d, int x, int y, int z) { return ((uint32_t)x * d.Y + y) * d.Z + z; }

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

// same-label 6-neighbour one Manhattan step closer to p?  The neighbour lies between the cell and p: always inside the grid.
// A cell next to p is entered from the start itself, whatever label p's cell carries (a later seed may have taken it).
__device__ __forceinline__ bool descends(const uint16_t* g, const Dims& d, int x, int y, int z, uint32_t L, const ushort4& p)
{
    if (abs(x - (int)p.x) + abs(y - (int)p.y) + abs(z - (int)p.z) == 1) return true;
    if (z != p.z && g[cell(d, x, y, z + (z > p.z ? -1 : 1))] == L) return true;
    if (y != p.y && g[cell(d, x, y + (y > p.y ? -1 : 1), z)] == L) return true;
    if (x != p.x && g[cell(d, x + (x > p.x ? -1 : 1), y, z)] == L) return true;
    return false;
}

__device__ __forceinline__ void append(uint32_t i, uint32_t* __restrict__ Dbits, uint32_t* __restrict__ list, Ctl* ctl)
{
    const uint32_t old = atomicOr(&Dbits[i >> 5], 1u << (i & 31u));
    if (old >> (i & 31u) & 1u) return;
    const uint32_t j = atomicAdd(&ctl->tail, 1u);
    if (j < kCap) list[j] = i;
    else ctl->overflow = 1;
}

// pass 1: FREE -> EMPTY, and F (cells without a descent neighbour) into the list / bitmap.  Thread per 8-cell chunk of a z-row.
__global__ void __launch_bounds__(256) certificate_kernel(uint16_t* grid, Dims d, const ushort4* __restrict__ seeds, int S, const uint32_t* __restrict__ table,
                                                          uint32_t* __restrict__ Dbits, uint32_t* __restrict__ list, Ctl* ctl)
{
    extern __shared__ ushort4 sp[];
    for (int i = threadIdx.x; i < S; i += blockDim.x) sp[i] = seeds[i];
    __syncthreads();
    const uint32_t cpr = (uint32_t)d.Z / 8, nchunks = (uint32_t)d.X * d.Y * cpr;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += gridDim.x * blockDim.x) {
        const uint32_t row = c / cpr;
        const int z0 = (int)(c - row * cpr) * 8, y = (int)(row % (uint32_t)d.Y), x = (int)(row / (uint32_t)d.Y);
        const uint32_t base = row * (uint32_t)d.Z + z0;
        uint4 v = *reinterpret_cast<const uint4*>(grid + base);
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        const uint32_t first = v.x & 0xFFFFu;
        if (first > VF_VOXEL_FREE && v.x == first * 0x10001u && v.y == v.x && v.z == v.x && v.w == v.x) {
            // one label over the chunk: every cell but the one nearest to the seed's z has its descent neighbour inside the chunk
            const uint32_t s = table[first];
            if (s == kNone) {
                for (int k = 0; k < 8; ++k) append(base + k, Dbits, list, ctl);
                continue;
            }
            const ushort4 p = sp[s];
            const int zc = min(max((int)p.z, z0), z0 + 7);
            const bool is_seed = x == p.x && y == p.y && zc == p.z;
            if (!is_seed && !descends(grid, d, x, y, zc, first, p)) append(base + (zc - z0), Dbits, list, ctl);
            continue;
        }
        uint32_t w[4] = { v.x, v.y, v.z, v.w };
        bool rewrite = false;
        for (int k = 0; k < 8; ++k) {
            const uint32_t L = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
            if (L == VF_VOXEL_FREE) {  // dropped: the reference rebuilds the grid from an all-EMPTY one
                w[k >> 1] &= ~(0xFFFFu << ((k & 1) * 16));
                rewrite = true;
            } else if (L > VF_VOXEL_FREE) {
                const uint32_t s = table[L];
                if (s == kNone) {
                    append(base + k, Dbits, list, ctl);
                    continue;
                }
                const ushort4 p = sp[s];
                const int z = z0 + k;
                if (!(x == p.x && y == p.y && z == p.z) && !descends(grid, d, x, y, z, L, p)) append(base + k, Dbits, list, ctl);
            }
        }
        // neighbours may read this chunk while it is rewritten: FREE and EMPTY both differ from every label, so their verdicts do not change
        if (rewrite) *reinterpret_cast<uint4*>(grid + base) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// pass 2 (one CTA; D is small): closure of F, liveness inside D, removal
__global__ void __launch_bounds__(kResolveThreads) resolve_kernel(uint16_t* grid, Dims d, const ushort4* __restrict__ seeds, int S, const uint32_t* __restrict__ table,
                                                                  uint32_t* Dbits, uint32_t* alive, uint32_t* list, Ctl* ctl)
{
    extern __shared__ ushort4 sp[];
    __shared__ uint32_t s_head, s_tail, s_flag;
    const int t = threadIdx.x;
    for (int i = t; i < S; i += blockDim.x) sp[i] = seeds[i];
    if (t == 0) s_head = 0, s_tail = min(ctl->tail, kCap), s_flag = ctl->overflow;
    __syncthreads();
    if (s_flag) return;
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
                if (grid[v] != L || bit(Dbits, v) || man(vx, vy, vz, p) != du + 1) continue;
                bool all_dead = true;  // v is one step farther than u, so it is not the seed and differs from p on at least one axis
                if (vz != p.z) {
                    const uint32_t n = cell(d, vx, vy, vz + (vz > p.z ? -1 : 1));
                    all_dead = all_dead && !(grid[n] == L && !bit(Dbits, n));
                }
                if (vy != p.y) {
                    const uint32_t n = cell(d, vx, vy + (vy > p.y ? -1 : 1), vz);
                    all_dead = all_dead && !(grid[n] == L && !bit(Dbits, n));
                }
                if (vx != p.x) {
                    const uint32_t n = cell(d, vx + (vx > p.x ? -1 : 1), vy, vz);
                    all_dead = all_dead && !(grid[n] == L && !bit(Dbits, n));
                }
                if (all_dead) append(v, Dbits, list, ctl);
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
        for (int k = 0; k < 6; ++k) {
            const int vx = x + dx[k], vy = y + dy[k], vz = z + dz[k];
            if (!inside(vx, vy, vz)) continue;
            const uint32_t v = cell(d, vx, vy, vz);
            if (grid[v] == L && !bit(Dbits, v)) {
                atomicOr(&alive[u >> 5], 1u << (u & 31u));
                break;
            }
        }
    }
    for (;;) {
        __syncthreads();
        if (t == 0) s_flag = 0;
        __syncthreads();
        for (uint32_t i = t; i < nD; i += blockDim.x) {
            const uint32_t u = list[i];
            if (bit(alive, u)) continue;
            const uint32_t L = grid[u];
            int x, y, z;
            decode(u, x, y, z);
            for (int k = 0; k < 6; ++k) {
                const int vx = x + dx[k], vy = y + dy[k], vz = z + dz[k];
                if (!inside(vx, vy, vz)) continue;
                const uint32_t v = cell(d, vx, vy, vz);
                if (grid[v] == L && bit(Dbits, v) && bit(alive, v)) {
                    atomicOr(&alive[u >> 5], 1u << (u & 31u));
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
        const uint32_t u = list[i];
        if (!bit(alive, u)) {
            grid[u] = VF_VOXEL_EMPTY;
            ++freed;
        }
    }
    if (freed) atomicAdd(&ctl->freed, freed);
}

}  // namespace vfc1

// Returns VF_OK with *handled = 1 when the grid now holds C1's result; *handled = 0 when the caller must run the union-find (the grid then
// carries the planted seed labels and has lost its FREE cells, both of which the union-find path does as well).
vf_status vf_k_c1_descent(vf_grid* grid, const ushort4* d_seeds, int nseeds, int* handled)
{
    using namespace vfc1;
    *handled = 0;
    vf_ctx* c = grid->ctx;
    const size_t n = grid->n();
    if (grid->Z % 8 != 0 || ((uintptr_t)grid->d & 15) != 0 || n >= (1ull << 31) || nseeds < 1 || (uint32_t)nseeds > kMaxSeeds) return VF_OK;
    Dims d = { (int)grid->X, (int)grid->Y, (int)grid->Z };
    // scratch (the flood key arena): D bitmap | liveness bitmap | label -> seed table | control block | list
    const size_t bm_words = (n + 31) / 32;
    const size_t need = bm_words * 8 + 65536 * 4 + 256 + (size_t)kCap * 4;
    VF_TRY(vf_scratch_reserve(c, c->keys, std::max(need, n * 4)));
    uint32_t* Dbits = (uint32_t*)c->keys.ptr;
    uint32_t* alive = Dbits + bm_words;
    uint32_t* table = alive + bm_words;
    Ctl* ctl = (Ctl*)(table + 65536);
    uint32_t* list = (uint32_t*)((char*)ctl + 256);
    VF_TRY(vf_k_zero(c, Dbits, bm_words * 8));
    VF_CUDA(cudaMemsetAsync(table, 0xFF, 65536 * 4, c->stream));
    VF_TRY(vf_k_zero(c, ctl, 256));
    plant_kernel<<<(nseeds + 127) / 128, 128, 0, c->stream>>>(grid->d, d, d_seeds, nseeds, table, ctl);
    VF_LAUNCHED(c);
    const size_t smem = (size_t)nseeds * sizeof(ushort4);
    certificate_kernel<<<c->num_sms * 8, 256, smem, c->stream>>>(grid->d, d, d_seeds, nseeds, table, Dbits, list, ctl);
    VF_LAUNCHED(c);
    Ctl* h = (Ctl*)((char*)c->pinned + 65536 + 256);
    VF_CUDA(cudaMemcpyAsync(h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    if (h->overflow || h->dup || h->tail > (1u << 16)) return VF_OK;  // too much for one CTA's list work: the union-find is the right tool
    if (h->tail == 0) {
        *handled = 1;  // every labelled cell has a descent chain to its seed
        return VF_OK;
    }
    resolve_kernel<<<1, kResolveThreads, smem, c->stream>>>(grid->d, d, d_seeds, nseeds, table, Dbits, alive, list, ctl);
    VF_LAUNCHED(c);
    VF_CUDA(cudaMemcpyAsync(h, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    if (h->overflow) return VF_OK;
    *handled = 1;
    return VF_OK;
}
