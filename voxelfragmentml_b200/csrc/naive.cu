// naive.cu — F1: nearest-seed Voronoi labelling of the occupied grid.
//
// Replaces NaiveFracturer::build (SRC/Fracturer/NaiveFracturer.cpp:215-225); the executable spec is buildCPU
// (:26-68) == naiveFracturer-comp.glsl:19-43 with the metrics of distance.glsl:5-21:
//     label(v) = seeds[argmin_i d(v, seed_i)].w   over non-EMPTY cells, strict '<' with i ascending (lowest index wins ties).
//
// B200 design (HBM bound: 2 B read + 2 B written per voxel; SURVEY §7 "naive kernel is ALU-bound unless seeds are culled"):
//   * one warp owns a 4(x) x 4(y) x 8*VEC(z) brick; each lane moves 128-bit (VEC=8) or 64-bit (VEC=4) vectors of labels,
//     eight lanes cover one 128-byte line of a z-row, four rows per instruction, four x-planes in flight per lane;
//   * while those loads are in flight the warp culls the seed set against the brick: seed s survives iff
//     minDist(s, brick) <= min_s' maxDist(s', brick)   (exact: a culled seed is strictly worse than s' for every voxel of
//     the brick), survivors are compacted with ballot/popc in ascending seed order into a per-warp shared-memory list;
//   * distances are evaluated in integers, which is exact for the reference's float32 compare: d^2 < 2^24 and float sqrt is
//     injective on integer d^2 up to 3*1181^2, Manhattan/Chebyshev are integers outright.  The running best is one
//     32-bit key  (d << 8 | slot)  so that a single min keeps the lowest-index seed on ties;
//   * seeds live in shared memory (loaded once per CTA), CTAs are persistent over bricks.
// Grids that do not meet the fast path's preconditions (Z % 4 != 0, or Euclidean with an axis > 1182 where float sqrt stops
// being injective) take the generic kernel, which compares float32 distances exactly like buildCPU.
#include "vf_internal.h"

namespace {

constexpr int kWarps = 8;
constexpr int kCMax = 64;  // candidate slots per warp brick (slot index must fit the key's low 8 bits)
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// distance from coordinate range [lo, hi] to p: smallest and largest |c - p|
__device__ __forceinline__ void axis_range(int lo, int hi, int p, int& dmin, int& dmax)
{
    dmin = max(0, max(lo - p, p - hi));
    dmax = max(iabs(p - lo), iabs(p - hi));
}

template <int DF>
__device__ __forceinline__ unsigned combine(int a, int b, int c)
{
    if (DF == VF_EUCLIDEAN) return (unsigned)(a * a + b * b + c * c);
    if (DF == VF_MANHATTAN) return (unsigned)(a + b + c);
    return (unsigned)max(a, max(b, c));
}

// exact float32 restatement of NaiveFracturer.cpp:12-23 for one voxel against every seed (slow path)
template <int DF>
__device__ __forceinline__ unsigned short scan_all_seeds(int x, int y, int z, const ushort4* __restrict__ seeds, int S, unsigned short own)
{
    float best = 3.402823466e+38f;  // FLT_MAX
    unsigned short lab = own;
    for (int s = 0; s < S; ++s) {
        const ushort4 sd = seeds[s];
        const int dx = x - (int)sd.x, dy = y - (int)sd.y, dz = z - (int)sd.z;
        float d;
        if (DF == VF_EUCLIDEAN) {
            const float fx = (float)dx, fy = (float)dy, fz = (float)dz;
            d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy)), __fmul_rn(fz, fz)));
        } else if (DF == VF_MANHATTAN) {
            d = (float)(iabs(dx) + iabs(dy) + iabs(dz));
        } else {
            d = (float)max(iabs(dx), max(iabs(dy), iabs(dz)));
        }
        if (d < best) {
            best = d;
            lab = sd.w;
        }
    }
    return lab;
}

template <int VEC>
struct VecT;
template <>
struct VecT<8> {
    typedef uint4 type;
};
template <>
struct VecT<4> {
    typedef uint2 type;
};

template <int VEC>
__device__ __forceinline__ void unpack(const typename VecT<VEC>::type& v, unsigned (&w)[VEC / 2]);
template <>
__device__ __forceinline__ void unpack<8>(const uint4& v, unsigned (&w)[4])
{
    w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
}
template <>
__device__ __forceinline__ void unpack<4>(const uint2& v, unsigned (&w)[2])
{
    w[0] = v.x, w[1] = v.y;
}
template <int VEC>
__device__ __forceinline__ typename VecT<VEC>::type pack(const unsigned (&w)[VEC / 2]);
template <>
__device__ __forceinline__ uint4 pack<8>(const unsigned (&w)[4])
{
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ uint2 pack<4>(const unsigned (&w)[2])
{
    return make_uint2(w[0], w[1]);
}

template <int DF, int VEC>
__global__ void __launch_bounds__(kWarps * 32, 4)
naive_brick_kernel(uint16_t* __restrict__ grid, int X, int Y, int Z, const ushort4* __restrict__ seeds_g, int S, int nby, int nbz,
                   unsigned total_bricks)
{
    typedef typename VecT<VEC>::type V;
    constexpr int BZ = 8 * VEC;  // brick extent along z
    extern __shared__ ushort4 smem[];
    ushort4* sseeds = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ushort4* cand = smem + ((S + 3) & ~3) + warp * kCMax;

    for (int i = threadIdx.x; i < S; i += blockDim.x) sseeds[i] = seeds_g[i];
    __syncthreads();

    const int ly = lane >> 3, lz = lane & 7;
    for (unsigned brick = blockIdx.x * kWarps + warp; brick < total_bricks; brick += gridDim.x * kWarps) {
        const int bz = brick % nbz;
        const unsigned t = brick / nbz;
        const int by = t % nby, bx = t / nby;
        const int x0 = bx * 4, y0 = by * 4, z0 = bz * BZ;
        const int y = y0 + ly, z = z0 + lz * VEC;
        const bool rowvalid = (y < Y) && (z < Z);  // Z % VEC == 0, so a valid chunk is entirely inside the row

        // ---- 1. put four 128-bit (64-bit) loads in flight
        V raw[4];
        bool valid[4];
        V* ptr[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int x = x0 + it;
            valid[it] = rowvalid && (x < X);
            ptr[it] = reinterpret_cast<V*>(grid + ((size_t)x * Y + y) * Z + z);
            if (valid[it]) raw[it] = vf_ldg_stream(ptr[it]);
        }

        // ---- 2. cull the seed set against the brick while the loads fly
        const int x1 = min(x0 + 3, X - 1), y1 = min(y0 + 3, Y - 1), z1 = min(z0 + BZ - 1, Z - 1);
        unsigned bound = 0xFFFFFFFFu;
        for (int base = 0; base < S; base += 32) {
            const int s = base + lane;
            if (s < S) {
                const ushort4 sd = sseeds[s];
                int a0, a1, b0, b1, c0, c1;
                axis_range(x0, x1, sd.x, a0, a1);
                axis_range(y0, y1, sd.y, b0, b1);
                axis_range(z0, z1, sd.z, c0, c1);
                bound = min(bound, combine<DF>(a1, b1, c1));
            }
        }
        bound = __reduce_min_sync(kFull, bound);
        int C = 0;
        for (int base = 0; base < S; base += 32) {
            const int s = base + lane;
            bool keep = false;
            ushort4 sd = make_ushort4(0, 0, 0, 0);
            if (s < S) {
                sd = sseeds[s];
                int a0, a1, b0, b1, c0, c1;
                axis_range(x0, x1, sd.x, a0, a1);
                axis_range(y0, y1, sd.y, b0, b1);
                axis_range(z0, z1, sd.z, c0, c1);
                keep = combine<DF>(a0, b0, c0) <= bound;
            }
            const unsigned m = __ballot_sync(kFull, keep);
            if (keep) {
                const int slot = C + __popc(m & ((1u << lane) - 1));
                if (slot < kCMax) cand[slot] = sd;
            }
            C += __popc(m);
        }
        __syncwarp();

        // ---- 3. label the four x-planes of the brick
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            unsigned w[VEC / 2];
            bool any = false;
            if (valid[it]) {
                unpack<VEC>(raw[it], w);
#pragma unroll
                for (int k = 0; k < VEC / 2; ++k) any = any || (w[k] != 0);
            }
            if (!any) continue;  // nothing occupied in this chunk: no store (2*N_occ write bytes)
            const int x = x0 + it;
            unsigned short lab[VEC];
            if (C <= kCMax) {
                // For EUCLIDEAN and MANHATTAN the difference of two seeds' distances along an axis-parallel line is monotone,
                // so the set of z where one seed is the (lowest-index) winner is an interval: when both ends of the chunk have
                // the same winner the whole chunk has it, and only chunks that straddle a cell boundary evaluate every voxel.
                // (Not true for CHEBYSHEV: max(c, |dz|) plateaus let a lower-index seed win two disjoint tie ranges.)
                unsigned key[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) key[k] = 0xFFFFFFFFu;
                if (DF != VF_CHEBYSHEV) {
                    unsigned k0 = 0xFFFFFFFFu, k1 = 0xFFFFFFFFu;
                    for (int slot = 0; slot < C; ++slot) {
                        const ushort4 sd = cand[slot];
                        const int dx = x - (int)sd.x, dy = y - (int)sd.y;
                        if (DF == VF_EUCLIDEAN) {
                            const unsigned base = ((unsigned)(dx * dx + dy * dy) << 8) | (unsigned)slot;
                            const int zs = (z - (int)sd.z) * 16, ze = zs + 16 * (VEC - 1);
                            k0 = min(k0, base + (unsigned)(zs * zs));
                            k1 = min(k1, base + (unsigned)(ze * ze));
                        } else {
                            const unsigned base = ((unsigned)(iabs(dx) + iabs(dy)) << 8) | (unsigned)slot;
                            const int zs = (z - (int)sd.z) * 256;
                            k0 = min(k0, base + (unsigned)iabs(zs));
                            k1 = min(k1, base + (unsigned)iabs(zs + 256 * (VEC - 1)));
                        }
                    }
                    if ((k0 & 0xFFu) == (k1 & 0xFFu)) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k) key[k] = k0;
                    } else {
                        key[0] = k0, key[VEC - 1] = k1;
                        for (int slot = 0; slot < C; ++slot) {
                            const ushort4 sd = cand[slot];
                            const int dx = x - (int)sd.x, dy = y - (int)sd.y;
                            if (DF == VF_EUCLIDEAN) {
                                const unsigned base = ((unsigned)(dx * dx + dy * dy) << 8) | (unsigned)slot;
                                const int zs = (z - (int)sd.z) * 16;  // (16*dz)^2 = dz^2 << 8
#pragma unroll
                                for (int k = 1; k < VEC - 1; ++k) {
                                    const int dzs = zs + 16 * k;
                                    key[k] = min(key[k], base + (unsigned)(dzs * dzs));
                                }
                            } else {
                                const unsigned base = ((unsigned)(iabs(dx) + iabs(dy)) << 8) | (unsigned)slot;
                                const int zs = (z - (int)sd.z) * 256;
#pragma unroll
                                for (int k = 1; k < VEC - 1; ++k) key[k] = min(key[k], base + (unsigned)iabs(zs + 256 * k));
                            }
                        }
                    }
                } else {
                    for (int slot = 0; slot < C; ++slot) {
                        const ushort4 sd = cand[slot];
                        const int dx = x - (int)sd.x, dy = y - (int)sd.y;
                        const unsigned base = ((unsigned)max(iabs(dx), iabs(dy)) << 8) | (unsigned)slot;
                        const int zs = (z - (int)sd.z) * 256;
#pragma unroll
                        for (int k = 0; k < VEC; ++k) key[k] = min(key[k], max(base, (unsigned)iabs(zs + 256 * k) | (unsigned)slot));
                    }
                }
#pragma unroll
                for (int k = 0; k < VEC; ++k) lab[k] = cand[key[k] & 0xFFu].w;
            } else {
                // more survivors than slots (brick equidistant to many seeds): exact scan of the whole seed set
#pragma unroll
                for (int k = 0; k < VEC; ++k) lab[k] = scan_all_seeds<DF>(x, y, z + k, sseeds, S, 0);
            }
#pragma unroll
            for (int k = 0; k < VEC / 2; ++k) {
                const unsigned lo = (w[k] & 0xFFFFu) ? (unsigned)lab[2 * k] : 0u;
                const unsigned hi = (w[k] >> 16) ? (unsigned)lab[2 * k + 1] : 0u;
                w[k] = lo | (hi << 16);
            }
            vf_stg_stream(ptr[it], pack<VEC>(w));
        }
        __syncwarp();  // cand[] is rewritten by the next brick
    }
}

// generic path: any dims, float32 compare exactly as buildCPU.  One thread per voxel.
template <int DF>
__global__ void __launch_bounds__(256) naive_generic_kernel(uint16_t* __restrict__ grid, int X, int Y, int Z, const ushort4* __restrict__ seeds_g, int S)
{
    extern __shared__ ushort4 smem[];
    for (int i = threadIdx.x; i < S; i += blockDim.x) smem[i] = seeds_g[i];
    __syncthreads();
    const size_t n = (size_t)X * Y * Z;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned short own = grid[i];
        if (own == VF_VOXEL_EMPTY) continue;
        const int z = (int)(i % Z);
        const size_t r = i / Z;
        const int y = (int)(r % Y), x = (int)(r / Y);
        grid[i] = scan_all_seeds<DF>(x, y, z, smem, S, own);
    }
}

template <int DF, int VEC>
vf_status launch_brick(vf_grid* g, const ushort4* d_seeds, int S)
{
    vf_ctx* c = g->ctx;
    const int nbx = (g->X + 3) / 4, nby = (g->Y + 3) / 4, nbz = (g->Z + 8 * VEC - 1) / (8 * VEC);
    const unsigned total = (unsigned)nbx * nby * nbz;
    const size_t smem = ((size_t)((S + 3) & ~3) + kWarps * kCMax) * sizeof(ushort4);
    auto kern = naive_brick_kernel<DF, VEC>;
    if (smem > 48 * 1024) VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = min((total + kWarps - 1) / kWarps, (unsigned)c->num_sms * 8u);
    kern<<<blocks, kWarps * 32, smem, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_seeds, S, nby, nbz, total);
    VF_LAUNCHED(c);
    return VF_OK;
}

template <int DF>
vf_status launch_generic(vf_grid* g, const ushort4* d_seeds, int S)
{
    vf_ctx* c = g->ctx;
    const size_t smem = (size_t)S * sizeof(ushort4);
    auto kern = naive_generic_kernel<DF>;
    if (smem > 48 * 1024) VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)min((g->n() + 255) / 256, (size_t)c->num_sms * 16);
    kern<<<blocks, 256, smem, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_seeds, S);
    VF_LAUNCHED(c);
    return VF_OK;
}

template <int DF>
vf_status dispatch(vf_grid* g, const ushort4* d_seeds, int S)
{
    const uint32_t maxdim = max(g->X, max(g->Y, g->Z));
    const bool int_exact = DF != VF_EUCLIDEAN || maxdim <= 1182;  // float sqrt injective on integer d^2
    const bool aligned = ((uintptr_t)g->d & 15) == 0;
    if (int_exact && aligned && g->Z % 8 == 0) return launch_brick<DF, 8>(g, d_seeds, S);
    if (int_exact && aligned && g->Z % 4 == 0) return launch_brick<DF, 4>(g, d_seeds, S);
    return launch_generic<DF>(g, d_seeds, S);
}

}  // namespace

vf_status vf_k_naive(vf_grid* g, const ushort4* d_seeds, uint32_t nseeds, int dfunc)
{
    VF_REQUIRE(nseeds <= 16384, VF_ERR_CAPACITY, "naive: %u seeds exceed the shared-memory seed table (16384)", nseeds);
    switch (dfunc) {
    case VF_EUCLIDEAN: return dispatch<VF_EUCLIDEAN>(g, d_seeds, (int)nseeds);
    case VF_MANHATTAN: return dispatch<VF_MANHATTAN>(g, d_seeds, (int)nseeds);
    case VF_CHEBYSHEV: return dispatch<VF_CHEBYSHEV>(g, d_seeds, (int)nseeds);
    default: return vf_set_error(VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    }
}

extern "C" vf_status vf_fracture_naive(vf_grid* g, const uint32_t* seeds, uint32_t nseeds, int dfunc)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    VF_REQUIRE(dfunc >= 0 && dfunc <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    ushort4* d_seeds = nullptr;
    VF_TRY(vf_upload_seeds(g->ctx, seeds, nseeds, g->X, g->Y, g->Z, &d_seeds));
    return vf_k_naive(g, d_seeds, nseeds, dfunc);
}
