// vf_internal.h — shared declarations of libvoxfrag (not part of the public ABI; see include/voxfrag.h).
#pragma once
#include <sched.h>

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../include/voxfrag.h"

// ---------------------------------------------------------------------------------------------- errors
vf_status vf_set_error(vf_status code, const char* fmt, ...);

#define VF_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t vf_e__ = (call);                                                                    \
        if (vf_e__ != cudaSuccess)                                                                      \
            return vf_set_error(VF_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(vf_e__)); \
    } while (0)

#define VF_TRY(call)                        \
    do {                                    \
        vf_status vf_s__ = (call);          \
        if (vf_s__ != VF_OK) return vf_s__; \
    } while (0)

#define VF_REQUIRE(cond, code, ...)                        \
    do {                                                   \
        if (!(cond)) return vf_set_error(code, __VA_ARGS__); \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define VF_LAUNCHED(ctx)                \
    do {                                \
        ++(ctx)->launches;              \
        VF_CUDA(cudaGetLastError());    \
    } while (0)

// ---------------------------------------------------------------------------------------------- RNG
// std::mt19937 restated (32-bit Mersenne twister, init_genrand seeding) + libstdc++'s float recipe
// (SURVEY finding 9): u = float(raw) * 2^-32, a result that rounds to 1.0f becomes nextafter(1,0).
struct VfMt19937 {
    uint32_t mt[624];
    int idx;
    void seed(uint32_t s)
    {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void twist()
    {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7FFFFFFFu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
        }
        idx = 0;
    }
    uint32_t next()
    {
        if (idx >= 624) twist();
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9D2C5680u;
        y ^= (y << 15) & 0xEFC60000u;
        y ^= y >> 18;
        return y;
    }
    float uniform()
    {
        float u = (float)next() * 2.3283064365386963e-10f;
        return u >= 1.0f ? 0.99999994f : u;
    }
    float uniform(float lo, float hi) { return lo + (hi - lo) * uniform(); }     // RandomUtilities.h:108-111
    int uniform_int(int lo, int hi) { return (int)uniform((float)lo, (float)hi); } // RandomUtilities.h:141-144
};

// ---------------------------------------------------------------------------------------------- objects
// thin-front flood solver (flood.cu): the list of cells a phase starts from + a small header, kept behind the tile worklists in the `tiles` arena
constexpr size_t kVfFrontCap = (size_t)1 << 16;
constexpr size_t kVfFrontBytes = kVfFrontCap * 4 + 512;

struct VfScratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct vf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    int smem_optin = 0;
    uint64_t launches = 0;
    uint64_t host_waits = 0;  // times the host waited for the stream (vf_sync): what a producer with few cores pays per call
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    cudaEvent_t ev_block = nullptr;  // cudaEventBlockingSync: waits that give the host core back (vf_ctx_set_blocking_sync)
    bool blocking_sync = false;
    bool yield_wait = false;  // waits poll and give the core to any runnable thread between polls (vf_ctx_set_blocking_sync(ctx, 2))
    uint32_t flood_levels = 0;  // width of a flood round's distance window; 0 = the library default (vf_ctx_set_flood_levels)
    int flood_coop = 4;         // flood: CTAs per SM of the cooperative round loop, 0 = one launch per round (vf_ctx_set_flood_mode)
    uint32_t flood_front = 8192;   // flood: most (cell, key) pairs the thin-front solver keeps pending before it hands over to the tiles; 0 = tiles only (vf_ctx_set_flood_front)
    const void* hist_clean = nullptr;  // histogram: device bins at this address are known to be zero (the previous call left them so)
    uint32_t c1_declined = 0;   // C1: calls since the certificate last declined a grid of this context (0: it handled the last one)
    int c1_mode = 0;            // C1: 0 = descent certificate on grids of >= 2^26 cells with the union-find as its fallback, 1 = union-find only, 2 = certificate at any size (vf_ctx_set_c1_mode)
    VfMt19937 rng;
    uint32_t crand = 80;  // state of the C runtime's rand() as the reference's platform implements it (MSVC LCG); srand(_seed), CADScene.cpp:36
    // scratch arenas, grown on demand (FloodFracturer.cpp:116-120 "grown on demand")
    VfScratch keys;      // flood: 4 B / voxel (dist<<15 | order)
    VfScratch grid2;     // second label grid (erode destination, snapshot sweeps): 2 B / voxel
    VfScratch tiles;     // tile worklists + flags
    VfScratch small;     // seeds, counters, histogram bins, masks
    VfScratch noise;     // erosion noise table
    VfScratch mesh;      // voxelizer: vertices, faces, bins
    VfScratch bits;      // vf_grid_upload_bits: the occupancy bitmap before it is expanded
    VfScratch codec;     // device-side .rle encoder: tile counts, run starts / values, packed records
    void* pinned = nullptr;  // small pinned host mailbox for counters
    size_t pinned_bytes = 0;
    // Host shadows of what the seed area of `small` and the `noise` arena hold on the device.  A call that brings the same
    // seeds / noise table again (every step of a batch producer does) skips the transfer: copy engines serve transfers in
    // submission order, so even a 1 KiB upload would wait for a neighbouring context's whole-grid copy to drain.
    std::vector<ushort4> seed_shadow;
    std::vector<float> noise_shadow;
};

struct vf_grid {
    vf_ctx* ctx = nullptr;
    uint16_t* d = nullptr;
    bool own = false;
    size_t capacity = 0;  // in voxels
    uint32_t X = 0, Y = 0, Z = 0;
    float aabb_min[3] = { -0.5f, -0.5f, -0.5f };
    float aabb_max[3] = { 0.5f, 0.5f, 0.5f };
    size_t n() const { return (size_t)X * Y * Z; }
};

// Every host wait on the context's stream goes through here.  Spinning (the runtime's default) has the lowest latency and is right
// for one job per host core; a producer that drives more contexts than it has cores (batch mode: the SMs are filled by overlapping
// the latency-bound rounds of several jobs) asks for blocking waits, which sleep on an event instead of burning the core.
inline cudaError_t vf_sync(vf_ctx* c)
{
    ++c->host_waits;
    if (!c->blocking_sync && !c->yield_wait) return cudaStreamSynchronize(c->stream);
    cudaError_t e = cudaEventRecord(c->ev_block, c->stream);
    if (e != cudaSuccess) return e;
    if (c->blocking_sync) return cudaEventSynchronize(c->ev_block);
    while ((e = cudaEventQuery(c->ev_block)) == cudaErrorNotReady) sched_yield();
    return e;
}

vf_status vf_scratch_reserve(vf_ctx* ctx, VfScratch& s, size_t bytes);
vf_status vf_enter(vf_ctx* ctx);  // cudaSetDevice
vf_status vf_upload_seeds(vf_ctx* ctx, const uint32_t* seeds, uint32_t n, uint32_t X, uint32_t Y, uint32_t Z, ushort4** d_out);
vf_status vf_k_zero(vf_ctx* ctx, void* d, size_t bytes);  // zero-fill by kernel (bytes % 4 == 0): never queues behind a copy engine

// ---------------------------------------------------------------------------------------------- kernels (one per file)
vf_status vf_k_naive(vf_grid* g, const ushort4* d_seeds, uint32_t nseeds, int dfunc);
vf_status vf_k_keep_seed_components(vf_grid* grid, const ushort4* d_starts, int nstarts, int mode, int nneigh, uint32_t* d_freed);  // ccl.cu
vf_status vf_k_c1_descent(vf_grid* grid, const ushort4* d_seeds, int nseeds, uint32_t max_label, int* handled);  // c1_descent.cu
vf_status vf_k_pointwise(vf_grid* g, int op);  // 0 undoMask(bit15) 1 undoMask(rightmost 8) 2 resetFilling 3 homogenize
enum { VF_PW_UNMASK15 = 0, VF_PW_RIGHTMOST8 = 1, VF_PW_RESET_FILLING = 2, VF_PW_HOMOGENIZE = 3 };

// ---------------------------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint4 vf_ldg_stream(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 vf_ldg_stream(const uint2* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void vf_stg_stream(uint4* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void vf_stg_stream(uint2* p, const uint2& v)
{
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
#endif
