// context.cu — contexts, grids, RNG, errors, host<->device copies and the trivially HBM-bound pointwise passes.
// Replaces the reference's GL runtime plumbing (SRC/Graphics/Core/ComputeShader.cpp) and RegularGrid's
// host/SSBO mirroring (SRC/DataStructures/RegularGrid.cpp:505-514, 583-599).
#include <cmath>
#include <cstring>
#include <new>

#include <algorithm>

#include "vf_internal.h"

// ---------------------------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

vf_status vf_set_error(vf_status code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* vf_last_error(void) { return g_err; }
extern "C" const char* vf_version(void) { return "voxfrag-b200 0.1 (sm_100a)"; }

extern "C" int vf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" void vf_params_default(vf_params* p)
{
    // FractureParameters::FractureParameters(), FractureParameters.h:91-145
    std::memset(p, 0, sizeof(*p));
    p->biasFocus = 5;
    p->biasSeeds = 32;
    p->clampVoxelMetricUnit = 200;
    p->erode = 0;
    p->erosionConvolution = VF_ELLIPSE;
    p->erosionIterations = 3;
    p->erosionProbability = .5f;
    p->erosionSize = 3;
    p->erosionThreshold = .5f;
    p->fractureAlgorithm = VF_FLOOD;
    p->distanceFunction = VF_CHEBYSHEV;
    p->launchGPU = 1;
    p->mergeSeedsDistanceFunction = VF_EUCLIDEAN;
    p->neighbourhoodType = VF_VON_NEUMANN;
    p->numExtraSeeds = 16;
    p->numImpacts = 0;
    p->numSeeds = 8;
    p->removeIsolatedRegions = 1;
    p->seed = 80;
    p->seedingRandom = VF_STD_UNIFORM;
    p->voxelPerMetricUnit = 20;
    p->voxelizationSize[0] = p->voxelizationSize[1] = p->voxelizationSize[2] = 128;
    p->exportGridExtension = VF_VOX;
    p->floodIdBits = 0;
    p->erodeBoundaryMode = 0;
}

// ---------------------------------------------------------------------------------------------- context
vf_status vf_enter(vf_ctx* ctx)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    VF_CUDA(cudaSetDevice(ctx->device));
    return VF_OK;
}

static vf_status ctx_create(int device, void* stream, bool borrow, vf_ctx** out)
{
    VF_REQUIRE(out != nullptr, VF_ERR_INVALID_ARGUMENT, "out == NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return vf_set_error(VF_ERR_CUDA, "no CUDA device: libvoxfrag has no CPU fallback");
    }
    VF_REQUIRE(device >= 0 && device < n, VF_ERR_INVALID_ARGUMENT, "device %d out of range (have %d)", device, n);
    VF_CUDA(cudaSetDevice(device));
    vf_ctx* c = new (std::nothrow) vf_ctx();
    VF_REQUIRE(c != nullptr, VF_ERR_CAPACITY, "out of host memory");
    c->device = device;
    // any failure below releases what has been created so far (vf_ctx_destroy copes with a half-built context)
    auto fail = [&](cudaError_t e, const char* what) {
        vf_ctx_destroy(c);
        return vf_set_error(VF_ERR_CUDA, "vf_ctx_create: %s -> %s", what, cudaGetErrorString(e));
    };
    cudaError_t e;
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(e, "cudaGetDeviceProperties");
    c->num_sms = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (borrow) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreateWithFlags");
        c->own_stream = true;
    }
    if ((e = cudaEventCreate(&c->ev_start)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&c->ev_stop)) != cudaSuccess) return fail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&c->ev_block, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreateWithFlags");
    c->pinned_bytes = (1 << 17) + (VF_HISTOGRAM_BINS * 4 + 64);  // [0, 64K) seed staging, [64K, 128K) counter mailbox, then the histogram read-back
    if ((e = cudaMallocHost(&c->pinned, c->pinned_bytes)) != cudaSuccess) return fail(e, "cudaMallocHost");
    c->rng.seed(80);  // FractureParameters::_seed default (FractureParameters.h:116), applied at CADScene.cpp:36-37
    *out = c;
    return VF_OK;
}

extern "C" vf_status vf_ctx_set_blocking_sync(vf_ctx* ctx, int on)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    ctx->blocking_sync = on == 1;
    ctx->yield_wait = on == 2;
    return VF_OK;
}

extern "C" vf_status vf_ctx_set_flood_levels(vf_ctx* ctx, uint32_t levels)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    VF_REQUIRE(levels < (1u << 17), VF_ERR_INVALID_ARGUMENT, "flood levels %u: the key field holds 17 bits of distance (0 = library default)", levels);
    ctx->flood_levels = levels;
    return VF_OK;
}

extern "C" vf_status vf_ctx_set_flood_front(vf_ctx* ctx, uint32_t max_front_cells)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    VF_REQUIRE(max_front_cells <= (uint32_t)kVfFrontCap, VF_ERR_INVALID_ARGUMENT, "flood front limit %u: the front lists hold %u cells (0 = tiles only)", max_front_cells,
               (uint32_t)kVfFrontCap);
    ctx->flood_front = max_front_cells;
    return VF_OK;
}

extern "C" vf_status vf_ctx_set_flood_mode(vf_ctx* ctx, int ctas_per_sm)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    VF_REQUIRE(ctas_per_sm >= 0 && ctas_per_sm <= 4, VF_ERR_INVALID_ARGUMENT, "flood mode %d: 0 = one launch per round, 1..4 = CTAs per SM of the cooperative loop", ctas_per_sm);
    ctx->flood_coop = ctas_per_sm;
    return VF_OK;
}

extern "C" vf_status vf_ctx_set_c1_mode(vf_ctx* ctx, int mode)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    VF_REQUIRE(mode >= 0 && mode <= 2, VF_ERR_INVALID_ARGUMENT, "c1 mode %d (0 = descent certificate on large grids + fallback, 1 = union-find only, 2 = certificate on any grid + fallback)", mode);
    ctx->c1_mode = mode;
    return VF_OK;
}

extern "C" vf_status vf_ctx_create(int device, vf_ctx** out) { return ctx_create(device, nullptr, false, out); }
extern "C" vf_status vf_ctx_create_on_stream(int device, void* s, vf_ctx** out) { return ctx_create(device, s, true, out); }

extern "C" void vf_ctx_destroy(vf_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream && c->ev_block) vf_sync(c);
    VfScratch* all[] = { &c->keys, &c->grid2, &c->tiles, &c->small, &c->noise, &c->mesh, &c->codec, &c->bits };
    for (VfScratch* s : all)
        if (s->ptr) cudaFree(s->ptr);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->ev_start) cudaEventDestroy(c->ev_start);
    if (c->ev_stop) cudaEventDestroy(c->ev_stop);
    if (c->ev_block) cudaEventDestroy(c->ev_block);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

vf_status vf_scratch_reserve(vf_ctx* ctx, VfScratch& s, size_t bytes)
{
    if (s.bytes >= bytes) return VF_OK;
    if (&s == &ctx->small) ctx->seed_shadow.clear();   // the arena moves: what the shadows describe is gone
    if (&s == &ctx->noise) ctx->noise_shadow.clear();
    const size_t old_bytes = s.bytes;
    if (s.ptr) {
        VF_CUDA(vf_sync(ctx));
        VF_CUDA(cudaFree(s.ptr));
        s.ptr = nullptr;
        s.bytes = 0;
    }
    // An arena that has to move grows by at least half: cudaFree / cudaMalloc synchronise the whole device, and a producer that works through
    // models of slightly different sizes (tile counts, triangle bins) would otherwise pay that once per new maximum, stalling every other
    // context on the GPU each time.
    if (old_bytes) bytes = std::max(bytes, old_bytes + old_bytes / 2);
    bytes = (bytes + 255) & ~(size_t)255;
    VF_CUDA(cudaMalloc(&s.ptr, bytes));
    s.bytes = bytes;
    return VF_OK;
}

extern "C" vf_status vf_ctx_reserve(vf_ctx* ctx, uint32_t X, uint32_t Y, uint32_t Z)
{
    VF_TRY(vf_enter(ctx));
    const size_t n = (size_t)X * Y * Z;
    VF_TRY(vf_scratch_reserve(ctx, ctx->keys, n * 4));
    VF_TRY(vf_scratch_reserve(ctx, ctx->grid2, n * 2));
    // the smaller arenas a fragmentation of a grid of this size touches, so that no call has to move one later: tile worklists of the flood
    // (16 x 16 x 32 tiles: 14 B + 1 KiB of pending masks each), seeds / counters / histogram bins, brick bins of the voxelizer (4 x 4 x 32
    // bricks: 8 B each, plus room for a 64k-triangle mesh and its brick lists)
    const size_t nt = (size_t)((X + 15) / 16) * ((Y + 15) / 16) * ((Z + 31) / 32);
    VF_TRY(vf_scratch_reserve(ctx, ctx->tiles, nt * (16 + 1024) + 4096 + kVfFrontBytes));
    VF_TRY(vf_scratch_reserve(ctx, ctx->small, 1 << 20));
    const size_t nb = (size_t)((X + 3) / 4) * ((Y + 3) / 4) * ((Z + 31) / 32);
    VF_TRY(vf_scratch_reserve(ctx, ctx->mesh, nb * 8 + ((size_t)65536 * (24 + 64)) + 4096));
    return VF_OK;
}

extern "C" vf_status vf_ctx_synchronize(vf_ctx* ctx)
{
    VF_TRY(vf_enter(ctx));
    VF_CUDA(vf_sync(ctx));
    return VF_OK;
}

extern "C" void* vf_ctx_stream(vf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t vf_ctx_kernel_launches(vf_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" uint64_t vf_ctx_host_waits(vf_ctx* ctx) { return ctx ? ctx->host_waits : 0; }

extern "C" vf_status vf_ctx_timer_start(vf_ctx* ctx)
{
    VF_TRY(vf_enter(ctx));
    VF_CUDA(cudaEventRecord(ctx->ev_start, ctx->stream));
    return VF_OK;
}
extern "C" vf_status vf_ctx_timer_stop(vf_ctx* ctx, float* ms)
{
    VF_TRY(vf_enter(ctx));
    VF_CUDA(cudaEventRecord(ctx->ev_stop, ctx->stream));
    VF_CUDA(cudaEventSynchronize(ctx->ev_stop));
    VF_CUDA(cudaEventElapsedTime(ms, ctx->ev_start, ctx->ev_stop));
    return VF_OK;
}

// ---------------------------------------------------------------------------------------------- RNG
extern "C" vf_status vf_rng_seed(vf_ctx* ctx, uint32_t seed)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    ctx->rng.seed(seed);
    ctx->crand = seed;  // CADScene.cpp:36 srand(_fractParameters._seed) next to :37 RandomUtilities::initSeed
    return VF_OK;
}
extern "C" vf_status vf_crand_seed(vf_ctx* ctx, uint32_t seed)
{
    VF_REQUIRE(ctx != nullptr, VF_ERR_INVALID_ARGUMENT, "null context");
    ctx->crand = seed;
    return VF_OK;
}
extern "C" int vf_crand_next(vf_ctx* ctx)
{
    ctx->crand = ctx->crand * 214013u + 2531011u;  // the MSVC runtime's rand()
    return (int)((ctx->crand >> 16) & 0x7fffu);
}
extern "C" float vf_rng_uniform(vf_ctx* ctx) { return ctx->rng.uniform(); }
extern "C" uint32_t vf_rng_raw(vf_ctx* ctx) { return ctx->rng.next(); }
extern "C" vf_status vf_fill_noise(vf_ctx* ctx, float* noise, uint32_t n)
{
    VF_REQUIRE(ctx && noise, VF_ERR_INVALID_ARGUMENT, "null argument");
    for (uint32_t i = 0; i < n; ++i) noise[i] = ctx->rng.uniform(.0f, 1.0f);  // RegularGrid.cpp:238-244, serial order
    return VF_OK;
}

// ---------------------------------------------------------------------------------------------- grid
extern "C" vf_status vf_grid_create(vf_ctx* ctx, uint32_t X, uint32_t Y, uint32_t Z, vf_grid** out)
{
    VF_TRY(vf_enter(ctx));
    VF_REQUIRE(out && X && Y && Z, VF_ERR_INVALID_ARGUMENT, "bad grid dims %ux%ux%u", X, Y, Z);
    VF_REQUIRE(X <= 65535 && Y <= 65535 && Z <= 65535, VF_ERR_CAPACITY, "grid axis > 65535");
    vf_grid* g = new (std::nothrow) vf_grid();
    VF_REQUIRE(g != nullptr, VF_ERR_CAPACITY, "out of host memory");
    g->ctx = ctx;
    g->X = X, g->Y = Y, g->Z = Z;
    g->capacity = g->n();
    g->own = true;
    cudaError_t e = cudaMalloc(&g->d, g->capacity * sizeof(uint16_t) + 64);
    if (e != cudaSuccess) {
        const size_t want = g->capacity * 2;
        delete g;
        return vf_set_error(VF_ERR_CUDA, "cudaMalloc(%zu B): %s", want, cudaGetErrorString(e));
    }
    e = cudaMemsetAsync(g->d, 0, g->capacity * sizeof(uint16_t), ctx->stream);  // RegularGrid::buildGrid: all EMPTY
    if (e != cudaSuccess) {
        cudaFree(g->d);
        delete g;
        return vf_set_error(VF_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    }
    *out = g;
    return VF_OK;
}

extern "C" vf_status vf_grid_wrap(vf_ctx* ctx, void* dptr, uint32_t X, uint32_t Y, uint32_t Z, vf_grid** out)
{
    VF_TRY(vf_enter(ctx));
    VF_REQUIRE(out && dptr && X && Y && Z, VF_ERR_INVALID_ARGUMENT, "bad arguments");
    VF_REQUIRE(((uintptr_t)dptr & 15) == 0, VF_ERR_INVALID_ARGUMENT, "device pointer must be 16-byte aligned");
    VF_REQUIRE(X <= 65535 && Y <= 65535 && Z <= 65535, VF_ERR_CAPACITY, "grid axis > 65535");
    vf_grid* g = new (std::nothrow) vf_grid();
    VF_REQUIRE(g != nullptr, VF_ERR_CAPACITY, "out of host memory");
    g->ctx = ctx;
    g->d = (uint16_t*)dptr;
    g->X = X, g->Y = Y, g->Z = Z;
    g->capacity = g->n();
    g->own = false;
    *out = g;
    return VF_OK;
}

extern "C" void vf_grid_destroy(vf_grid* g)
{
    if (!g) return;
    if (g->own && g->d) {
        cudaSetDevice(g->ctx->device);
        vf_sync(g->ctx);
        cudaFree(g->d);
    }
    delete g;
}

extern "C" vf_status vf_grid_set_aabb(vf_grid* g, const float mn[3], const float mx[3], uint32_t X, uint32_t Y, uint32_t Z)
{
    // RegularGrid::setAABB (:426-441): re-dimension inside the existing allocation (dataset mode allocates once at the
    // clamp size, CADScene.cpp:529-543) and cleanGrid (:591-599)
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    VF_REQUIRE(X && Y && Z && (size_t)X * Y * Z <= g->capacity, VF_ERR_CAPACITY, "setAABB dims %ux%ux%u exceed the allocation (%zu voxels)",
               X, Y, Z, g->capacity);
    for (int i = 0; i < 3; ++i) g->aabb_min[i] = mn[i], g->aabb_max[i] = mx[i];
    g->X = X, g->Y = Y, g->Z = Z;
    VF_CUDA(cudaMemsetAsync(g->d, 0, g->n() * sizeof(uint16_t), g->ctx->stream));
    return VF_OK;
}

extern "C" vf_status vf_grid_dims(const vf_grid* g, uint32_t dims[3])
{
    VF_REQUIRE(g && dims, VF_ERR_INVALID_ARGUMENT, "null argument");
    dims[0] = g->X, dims[1] = g->Y, dims[2] = g->Z;
    return VF_OK;
}

extern "C" void* vf_grid_device_ptr(vf_grid* g) { return g ? g->d : nullptr; }

// Whole-grid copies are issued in 8 MiB pieces: a copy engine serves its queue in order, so one 256 MiB transfer of another
// context (the next model's upload in a pipelined producer) would otherwise hold back this context's small transfers — the seed
// list, the noise table, the histogram read-back — for its full duration.
static vf_status copy_in_pieces(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream)
{
    constexpr size_t kPiece = 8u << 20;
    for (size_t off = 0; off < bytes; off += kPiece)
        VF_CUDA(cudaMemcpyAsync((char*)dst + off, (const char*)src + off, std::min(kPiece, bytes - off), kind, stream));
    return VF_OK;
}

extern "C" vf_status vf_grid_upload_async(vf_grid* g, const uint16_t* host)
{
    VF_REQUIRE(g && host, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(g->ctx));
    return copy_in_pieces(g->d, host, g->n() * sizeof(uint16_t), cudaMemcpyHostToDevice, g->ctx->stream);
}
extern "C" vf_status vf_grid_download_async(vf_grid* g, uint16_t* host)
{
    VF_REQUIRE(g && host, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(g->ctx));
    return copy_in_pieces(host, g->d, g->n() * sizeof(uint16_t), cudaMemcpyDeviceToHost, g->ctx->stream);
}
// occupancy bit -> label word: one byte of the bitmap = one 128-bit store of eight cells
__global__ void __launch_bounds__(256) expand_bits_kernel(const uint8_t* __restrict__ bits, uint16_t* __restrict__ grid, size_t n)
{
    const size_t nvec = n / 8;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = bits[i];
        uint4 o;
        o.x = (b & 1u) | (b & 2u) << 15, o.y = (b >> 2 & 1u) | (b >> 2 & 2u) << 15;
        o.z = (b >> 4 & 1u) | (b >> 4 & 2u) << 15, o.w = (b >> 6 & 1u) | (b >> 6 & 2u) << 15;
        *reinterpret_cast<uint4*>(grid + i * 8) = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < n % 8) grid[nvec * 8 + threadIdx.x] = (bits[nvec] >> threadIdx.x) & 1u;
}

extern "C" vf_status vf_grid_upload_bits(vf_grid* g, const uint8_t* host_bits)
{
    VF_REQUIRE(g && host_bits, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    const size_t n = g->n(), nbytes = (n + 7) / 8;
    VF_REQUIRE(((uintptr_t)g->d & 15) == 0, VF_ERR_INVALID_ARGUMENT, "upload_bits needs a 16-byte aligned grid");
    VF_TRY(vf_scratch_reserve(c, c->bits, nbytes + 16));
    VF_TRY(copy_in_pieces(c->bits.ptr, host_bits, nbytes, cudaMemcpyHostToDevice, c->stream));
    expand_bits_kernel<<<c->num_sms * 8, 256, 0, c->stream>>>((const uint8_t*)c->bits.ptr, g->d, n);
    VF_LAUNCHED(c);
    return VF_OK;
}

extern "C" vf_status vf_grid_upload(vf_grid* g, const uint16_t* host)
{
    VF_TRY(vf_grid_upload_async(g, host));
    VF_CUDA(vf_sync(g->ctx));
    return VF_OK;
}
extern "C" vf_status vf_grid_download(vf_grid* g, uint16_t* host)
{
    VF_TRY(vf_grid_download_async(g, host));
    VF_CUDA(vf_sync(g->ctx));
    return VF_OK;
}

extern "C" void vf_dims_rule(const float mn[3], const float mx[3], uint32_t maxVoxels, uint32_t out[3])
{
    // CADScene::allocateMeshGrid, CADScene.cpp:545-556 (float32 arithmetic as written)
    float size[3], maxSize = 0.0f;
    for (int i = 0; i < 3; ++i) {
        size[i] = mx[i] - mn[i];
        maxSize = fmaxf(maxSize, size[i]);
    }
    for (int i = 0; i < 3; ++i) {
        int v = (int)floorf((float)maxVoxels * size[i] / maxSize);
        while (v % 4 != 0) ++v;
        while (v % 4 != 0 || v > (int)maxVoxels) --v;
        out[i] = (uint32_t)v;
    }
}

// ---------------------------------------------------------------------------------------------- seeds
vf_status vf_upload_seeds(vf_ctx* ctx, const uint32_t* seeds, uint32_t n, uint32_t X, uint32_t Y, uint32_t Z, ushort4** d_out)
{
    VF_REQUIRE(seeds && n > 0, VF_ERR_INVALID_ARGUMENT, "no seeds");
    VF_REQUIRE((size_t)n * sizeof(ushort4) <= 65536, VF_ERR_CAPACITY, "too many seeds (%u)", n);
    VF_TRY(vf_scratch_reserve(ctx, ctx->small, 1 << 20));
    std::vector<ushort4> packed(n);
    for (uint32_t i = 0; i < n; ++i) {
        VF_REQUIRE(seeds[4 * i] < X && seeds[4 * i + 1] < Y && seeds[4 * i + 2] < Z, VF_ERR_INVALID_ARGUMENT,
                   "seed %u (%u,%u,%u) outside the %ux%ux%u grid", i, seeds[4 * i], seeds[4 * i + 1], seeds[4 * i + 2], X, Y, Z);
        VF_REQUIRE(seeds[4 * i + 3] <= 0xFFFFu, VF_ERR_CAPACITY, "seed %u label %u does not fit the uint16 cell", i, seeds[4 * i + 3]);
        packed[i] = make_ushort4((unsigned short)seeds[4 * i], (unsigned short)seeds[4 * i + 1], (unsigned short)seeds[4 * i + 2],
                                 (unsigned short)seeds[4 * i + 3]);
    }
    *d_out = (ushort4*)ctx->small.ptr;
    if (ctx->seed_shadow.size() == n && std::memcmp(ctx->seed_shadow.data(), packed.data(), (size_t)n * sizeof(ushort4)) == 0) return VF_OK;  // already there
    // the pinned mailbox may still be in flight from a previous call on this stream
    VF_CUDA(vf_sync(ctx));
    std::memcpy(ctx->pinned, packed.data(), (size_t)n * sizeof(ushort4));
    ctx->seed_shadow.clear();
    VF_CUDA(cudaMemcpyAsync(ctx->small.ptr, ctx->pinned, (size_t)n * sizeof(ushort4), cudaMemcpyHostToDevice, ctx->stream));
    ctx->seed_shadow = std::move(packed);
    return VF_OK;
}

__global__ void zero_kernel(uint32_t* __restrict__ p, size_t nwords)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
vf_status vf_k_zero(vf_ctx* ctx, void* d, size_t bytes)
{
    const size_t nwords = bytes / 4;
    const int blocks = (int)std::min<size_t>((nwords + 255) / 256, (size_t)ctx->num_sms * 4);
    zero_kernel<<<blocks, 256, 0, ctx->stream>>>((uint32_t*)d, nwords);
    VF_LAUNCHED(ctx);
    return VF_OK;
}

// ---------------------------------------------------------------------------------------------- pointwise passes (C4)
// undoMask (undoMask-comp.glsl:19-37), resetFilling (RegularGrid.cpp:412-418), homogenize (:533-541), fill.
// 2 B read + 2 B write per voxel, 128-bit vectors, grid-stride; nothing to tile.
template <int OP>
__device__ __forceinline__ uint32_t pw2(uint32_t v, uint32_t fillv)
{
    if (OP == VF_PW_UNMASK15) return v & 0x7FFF7FFFu;
    if (OP == VF_PW_RIGHTMOST8) return v & 0x00FF00FFu;
    if (OP == VF_PW_RESET_FILLING) {
        const uint32_t lo = min(v & 0xFFFFu, 2u), hi = min(v >> 16, 2u);
        return lo | (hi << 16);
    }
    if (OP == VF_PW_HOMOGENIZE) return ((v & 0xFFFFu) ? 1u : 0u) | ((v >> 16) ? 0x10000u : 0u);
    return fillv;
}

template <int OP>
__global__ void __launch_bounds__(256) pointwise_kernel(uint16_t* __restrict__ grid, size_t n, uint32_t fillv)
{
    const size_t nvec = n / 8;
    uint4* g4 = reinterpret_cast<uint4*>(grid);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = OP == 4 ? make_uint4(0, 0, 0, 0) : vf_ldg_stream(g4 + i);
        v.x = pw2<OP>(v.x, fillv), v.y = pw2<OP>(v.y, fillv), v.z = pw2<OP>(v.z, fillv), v.w = pw2<OP>(v.w, fillv);
        vf_stg_stream(g4 + i, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < n % 8) {
        const size_t i = nvec * 8 + threadIdx.x;
        uint32_t v = grid[i];
        v = pw2<OP>(v, fillv) & 0xFFFFu;
        grid[i] = (uint16_t)v;
    }
}

vf_status vf_k_pointwise(vf_grid* g, int op)
{
    vf_ctx* c = g->ctx;
    const size_t n = g->n();
    const int blocks = (int)min((size_t)c->num_sms * 8, (n / 8 + 255) / 256 + 1);
    switch (op) {
    case VF_PW_UNMASK15: pointwise_kernel<VF_PW_UNMASK15><<<blocks, 256, 0, c->stream>>>(g->d, n, 0); break;
    case VF_PW_RIGHTMOST8: pointwise_kernel<VF_PW_RIGHTMOST8><<<blocks, 256, 0, c->stream>>>(g->d, n, 0); break;
    case VF_PW_RESET_FILLING: pointwise_kernel<VF_PW_RESET_FILLING><<<blocks, 256, 0, c->stream>>>(g->d, n, 0); break;
    case VF_PW_HOMOGENIZE: pointwise_kernel<VF_PW_HOMOGENIZE><<<blocks, 256, 0, c->stream>>>(g->d, n, 0); break;
    default: return vf_set_error(VF_ERR_INVALID_ARGUMENT, "bad pointwise op %d", op);
    }
    VF_LAUNCHED(c);
    return VF_OK;
}

extern "C" vf_status vf_grid_fill(vf_grid* g, uint16_t value)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    vf_ctx* c = g->ctx;
    const size_t n = g->n();
    const int blocks = (int)min((size_t)c->num_sms * 8, (n / 8 + 255) / 256 + 1);
    pointwise_kernel<4><<<blocks, 256, 0, c->stream>>>(g->d, n, (uint32_t)value | ((uint32_t)value << 16));
    VF_LAUNCHED(c);
    return VF_OK;
}

extern "C" vf_status vf_undo_mask(vf_grid* g)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    return vf_k_pointwise(g, VF_PW_UNMASK15);
}
extern "C" vf_status vf_reset_filling(vf_grid* g)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    return vf_k_pointwise(g, VF_PW_RESET_FILLING);
}
extern "C" vf_status vf_homogenize(vf_grid* g)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    return vf_k_pointwise(g, VF_PW_HOMOGENIZE);
}
