// naive.cu — F1: nearest-seed Voronoi labelling of the occupied grid.
//
// Replaces NaiveFracturer::build (SRC/Fracturer/NaiveFracturer.cpp:215-225); the executable spec is buildCPU
// (:26-68) == naiveFracturer-comp.glsl:19-43 with the metrics of distance.glsl:5-21:
//     label(v) = seeds[argmin_i d(v, seed_i)].w   over non-EMPTY cells, strict '<' with i ascending (lowest index wins ties).
//
// B200 design (HBM bound: 2 B read + 2 B written per voxel; a per-voxel loop over all seeds is ALU bound):
//   * one warp owns an 8(x) x 4(y) x 8*VEC(z) brick; each lane moves 128-bit (VEC=8) or 64-bit (VEC=4) vectors of labels,
//     eight lanes cover one 128-byte line of a z-row, four rows per instruction, eight x-planes in flight per lane;
//   * while those loads are in flight the warp culls the seed set against the brick with an exact pairwise DOMINANCE test:
//     let s* be the seed nearest to the brick centre; seed s is dropped iff s* beats s (smaller distance, or equal distance and
//     lower index) at every point of the brick.  For EUCLIDEAN d^2(c,s) - d^2(c,s*) is linear in c, for MANHATTAN it is a sum of
//     per-axis monotone functions, so its minimum over the box is attained at box ends per axis and the test costs ~30 integer
//     ops per seed.  Bricks inside one Voronoi cell end up with exactly one candidate and need no per-voxel arithmetic at all;
//     bricks on a cell boundary typically keep 2-3.  (CHEBYSHEV is not separable: it keeps the weaker but still exact
//     minDist(s) <= min_s' maxDist(s') rule.)  Survivors are compacted with ballot/popc in ascending seed order;
//   * distances are evaluated in integers, which is exact for the reference's float32 compare: d^2 < 2^24 and float sqrt is
//     injective on integer d^2 up to 3*1181^2, Manhattan/Chebyshev are integers outright.  The running best is one
//     32-bit key  (d << 8 | slot)  so that a single min keeps the lowest-index seed on ties;
//   * the brick's label vectors are staged in shared memory with cp.async (no registers held while the warp culls); in bricks cut
//     by a cell boundary every lane repeats the dominance test on its own 8 x 1 x VEC box, so only lanes whose box is cut evaluate
//     voxels, and their chunks are compacted with ballot/popc and dealt evenly to the lanes in pass 2;
//   * seeds live in shared memory (loaded once per CTA), CTAs are persistent over bricks.
// Grids that do not meet the fast path's preconditions (Z % 4 != 0, or Euclidean with an axis > 1182 where float sqrt stops
// being injective) take the generic kernel, which compares float32 distances exactly like buildCPU.
#include "vf_internal.h"

namespace {

constexpr int kWarps = 8;
constexpr int kIts = 8;    // x-planes per brick
constexpr int kCMax = 64;  // candidate slots per warp brick (slot index must fit the key's low 8 bits)
constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }

// distance from coordinate range [lo, hi] to p: smallest and largest |c - p|
__device__ __forceinline__ void axis_range(int lo, int hi, int p, int& dmin, int& dmax)
{
    dmin = max(0, max(lo - p, p - hi));
    dmax = max(iabs(p - lo), iabs(p - hi));
}

// min over c in {lo, hi} of  f(c, s) - f(c, t)   with f = squared difference (EUCLIDEAN) or absolute difference (MANHATTAN).
// Both are monotone in c (linear resp. clamp-shaped), so the minimum over the whole interval [lo, hi] is at an end.
template <int DF>
__device__ __forceinline__ int axis_gap_min(int lo, int hi, int s, int t)
{
    if (DF == VF_EUCLIDEAN) {
        const int a = (lo - s) * (lo - s) - (lo - t) * (lo - t);
        const int b = (hi - s) * (hi - s) - (hi - t) * (hi - t);
        return min(a, b);
    }
    const int a = iabs(lo - s) - iabs(lo - t);
    const int b = iabs(hi - s) - iabs(hi - t);
    return min(a, b);
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

// 0xFFFF in each 16-bit half of w that is non-zero
__device__ __forceinline__ unsigned nonzero_halves(unsigned w)
{
    return ((w & 0xFFFFu) ? 0xFFFFu : 0u) | ((w >> 16) ? 0xFFFF0000u : 0u);
}

// exact division by a runtime constant for the brick index decode (d < 2^16, n < 2^31)
struct FastDiv {
    unsigned d, m;
    __host__ FastDiv(unsigned div = 1) : d(div), m((unsigned)(((1ull << 32) + div - 1) / div)) {}
    __device__ __forceinline__ unsigned div(unsigned n) const { return d == 1 ? n : __umulhi(n, m); }
};

__device__ __forceinline__ void cp_async(void* smem, const void* gmem, int bytes16)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if (bytes16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}

// evaluate the candidates for the VEC voxels of one chunk (x, y fixed, z .. z+VEC-1); returns per-voxel labels packed 2 per word
template <int DF, int VEC>
__device__ __forceinline__ void eval_chunk(const ushort4* __restrict__ cand, int C, int x, int y, int z, unsigned (&lab2)[VEC / 2])
{
    unsigned key[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) key[k] = 0xFFFFFFFFu;
    for (int slot = 0; slot < C; ++slot) {
        const ushort4 sd = cand[slot];
        const int dx = x - (int)sd.x, dy = y - (int)sd.y;
        if (DF == VF_EUCLIDEAN) {
            const unsigned base = ((unsigned)(dx * dx + dy * dy) << 8) | (unsigned)slot;
            const int zs = (z - (int)sd.z) * 16;  // (16*dz)^2 = dz^2 << 8
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                const int dzs = zs + 16 * k;
                key[k] = min(key[k], base + (unsigned)(dzs * dzs));
            }
        } else if (DF == VF_MANHATTAN) {
            const unsigned base = ((unsigned)(iabs(dx) + iabs(dy)) << 8) | (unsigned)slot;
            const int zs = (z - (int)sd.z) * 256;
#pragma unroll
            for (int k = 0; k < VEC; ++k) key[k] = min(key[k], base + (unsigned)iabs(zs + 256 * k));
        } else {
            const unsigned base = ((unsigned)max(iabs(dx), iabs(dy)) << 8) | (unsigned)slot;
            const int zs = (z - (int)sd.z) * 256;
#pragma unroll
            for (int k = 0; k < VEC; ++k) key[k] = min(key[k], max(base, (unsigned)iabs(zs + 256 * k) | (unsigned)slot));
        }
    }
#pragma unroll
    for (int k = 0; k < VEC / 2; ++k) lab2[k] = (unsigned)cand[key[2 * k] & 0xFFu].w | ((unsigned)cand[key[2 * k + 1] & 0xFFu].w << 16);
}

template <int DF, int VEC>
__device__ __noinline__ void scan_chunk(const ushort4* __restrict__ seeds, int S, int x, int y, int z, unsigned (&lab2)[VEC / 2])
{
    for (int k = 0; k < VEC / 2; ++k)
        lab2[k] = (unsigned)scan_all_seeds<DF>(x, y, z + 2 * k, seeds, S, 0) | ((unsigned)scan_all_seeds<DF>(x, y, z + 2 * k + 1, seeds, S, 0) << 16);
}

template <int DF, int VEC>
__global__ void __launch_bounds__(kWarps * 32, 4)
naive_brick_kernel(uint16_t* __restrict__ grid, int X, int Y, int Z, const ushort4* __restrict__ seeds_g, int S, FastDiv div_nbz, FastDiv div_nby,
                   unsigned total_bricks, int xb)  // xb: first plane to label (0, or a slab's first plane: `grid` then points at the virtual plane 0)
{
    typedef typename VecT<VEC>::type V;
    constexpr int BZ = 8 * VEC;  // brick extent along z
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: staging [kWarps][kIts][32] vectors | seeds [S] | candidates [kWarps][kCMax] | task lists [kWarps][kIts*32] bytes
    V* stage_all = reinterpret_cast<V*>(smem_raw);
    ushort4* sseeds = reinterpret_cast<ushort4*>(smem_raw + sizeof(V) * kWarps * kIts * 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ushort4* cand = sseeds + ((S + 3) & ~3) + warp * kCMax;
    unsigned char* tasks = reinterpret_cast<unsigned char*>(sseeds + ((S + 3) & ~3) + kWarps * kCMax) + warp * kIts * 32;
    V* stage = stage_all + warp * kIts * 32;

    for (int i = threadIdx.x; i < S; i += blockDim.x) sseeds[i] = seeds_g[i];
    __syncthreads();

    const int ly = lane >> 3, lz = lane & 7;
    const size_t xstride = (size_t)Y * Z / VEC;  // in vectors
    for (unsigned brick = blockIdx.x * kWarps + warp; brick < total_bricks; brick += gridDim.x * kWarps) {
        const unsigned t = div_nbz.div(brick);
        const int bz = brick - t * div_nbz.d;
        const unsigned bx = div_nby.div(t);
        const int by = t - bx * div_nby.d;
        const int x0 = xb + bx * kIts, y0 = by * 4, z0 = bz * BZ;
        const int y = y0 + ly, z = z0 + lz * VEC;
        const bool rowvalid = (y < Y) && (z < Z);  // Z % VEC == 0, so a valid chunk is entirely inside the row
        const int nits = min(kIts, X - x0);

        // ---- 1. eight asynchronous 128-bit (64-bit) copies per lane, global -> shared, no registers held
        V* ptr0 = reinterpret_cast<V*>(grid + ((size_t)x0 * Y + y) * Z + z);
        if (rowvalid) {
            const V* gq = ptr0;
            for (int it = 0; it < nits; ++it, gq += xstride) cp_async(&stage[it * 32 + lane], gq, VEC == 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");

        // ---- 2. cull the seed set against the brick while the copies fly
        const int x1 = min(x0 + kIts - 1, X - 1), y1 = min(y0 + 3, Y - 1), z1 = min(z0 + BZ - 1, Z - 1);
        int C = 0;
        if (DF != VF_CHEBYSHEV) {
            // s* = seed nearest (L1, doubled coordinates) to the brick centre, lowest index on ties.  The choice of s* affects
            // only how many seeds survive, never the result.
            const int cx2 = x0 + x1, cy2 = y0 + y1, cz2 = z0 + z1;
            unsigned best = 0xFFFFFFFFu;
            for (int base = 0; base < S; base += 32) {
                const int s = base + lane;
                if (s < S) {
                    const ushort4 sd = sseeds[s];
                    const unsigned d = (unsigned)(iabs(2 * (int)sd.x - cx2) + iabs(2 * (int)sd.y - cy2) + iabs(2 * (int)sd.z - cz2));
                    best = min(best, (d << 13) | (unsigned)s);  // d < 2^19 (3 * 2 * 65535), s < 2^13
                }
            }
            const int star = (int)(__reduce_min_sync(kFull, best) & 0x1FFFu);
            const ushort4 st = sseeds[star];
            for (int base = 0; base < S; base += 32) {
                const int s = base + lane;
                bool keep = false;
                ushort4 sd = make_ushort4(0, 0, 0, 0);
                if (s < S) {
                    sd = sseeds[s];
                    // g(c) = d(c, s) - d(c, s*); s* beats s at c iff g > 0, or g == 0 and star < s
                    const int g = axis_gap_min<DF>(x0, x1, sd.x, st.x) + axis_gap_min<DF>(y0, y1, sd.y, st.y) + axis_gap_min<DF>(z0, z1, sd.z, st.z);
                    keep = (s == star) || !(g > 0 || (g == 0 && star < s));
                }
                const unsigned m = __ballot_sync(kFull, keep);
                if (keep) {
                    const int slot = C + __popc(m & ((1u << lane) - 1));
                    if (slot < kCMax) cand[slot] = sd;
                }
                C += __popc(m);
            }
        } else {
            unsigned bound = 0xFFFFFFFFu;
            for (int base = 0; base < S; base += 32) {
                const int s = base + lane;
                if (s < S) {
                    const ushort4 sd = sseeds[s];
                    int a0, a1, b0, b1, c0, c1;
                    axis_range(x0, x1, sd.x, a0, a1);
                    axis_range(y0, y1, sd.y, b0, b1);
                    axis_range(z0, z1, sd.z, c0, c1);
                    bound = min(bound, (unsigned)max(a1, max(b1, c1)));
                }
            }
            bound = __reduce_min_sync(kFull, bound);
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
                    keep = (unsigned)max(a0, max(b0, c0)) <= bound;
                }
                const unsigned m = __ballot_sync(kFull, keep);
                if (keep) {
                    const int slot = C + __popc(m & ((1u << lane) - 1));
                    if (slot < kCMax) cand[slot] = sd;
                }
                C += __popc(m);
            }
        }
        __syncwarp();

        // ---- 3. per-lane refinement (C >= 2): a lane owns the box  x0..x1 x {y} x z..z+VEC-1  (its chunk in every plane).  One point
        //      evaluation picks the winner w at the box corner; if w dominates every other candidate over the box (same separable
        //      minimum as above) the lane's chunks all take w's label without further arithmetic.  Lanes whose box is cut by a cell
        //      boundary queue their chunks for pass 2.  (CHEBYSHEV is not separable: with C >= 2 every chunk goes to pass 2.)
        unsigned lab = (unsigned)cand[0].w * 0x10001u;  // C == 1: the whole brick lies in one Voronoi cell
        bool lane_uniform = C == 1;
        if (C > 1 && C <= kCMax && DF != VF_CHEBYSHEV && rowvalid) {
            unsigned kbest = 0xFFFFFFFFu;
            for (int slot = 0; slot < C; ++slot) {
                const ushort4 sd = cand[slot];
                const int dx = x0 - (int)sd.x, dy = y - (int)sd.y, dz = z - (int)sd.z;
                const unsigned d = DF == VF_EUCLIDEAN ? (unsigned)(dx * dx + dy * dy + dz * dz) : (unsigned)(iabs(dx) + iabs(dy) + iabs(dz));
                kbest = min(kbest, (d << 8) | (unsigned)slot);
            }
            const int ws = (int)(kbest & 0xFFu);
            const ushort4 wn = cand[ws];
            bool all_dominated = true;
            for (int slot = 0; slot < C; ++slot) {
                const ushort4 sd = cand[slot];
                const int g = axis_gap_min<DF>(x0, x1, sd.x, wn.x) + axis_gap_min<DF>(y, y, sd.y, wn.y) + axis_gap_min<DF>(z, z + VEC - 1, sd.z, wn.z);
                all_dominated = all_dominated && (slot == ws || g > 0 || (g == 0 && ws < slot));  // slots are in seed-index order
            }
            lane_uniform = all_dominated;
            lab = (unsigned)wn.w * 0x10001u;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();

        // ---- pass 1: label the chunks whose lane box is uniform; queue the rest
        int ntasks = 0;
        V* gp = ptr0;
        for (int it = 0; it < nits; ++it, gp += xstride) {
            bool slow = false;
            if (rowvalid) {
                unsigned w[VEC / 2];
                unpack<VEC>(stage[it * 32 + lane], w);
                // SWAR: bit 15 / 31 of t is set iff the corresponding 16-bit half of w is non-zero
                unsigned all = 0x80008000u, any = 0;
#pragma unroll
                for (int k = 0; k < VEC / 2; ++k) {
                    const unsigned t = ((w[k] & 0x7FFF7FFFu) + 0x7FFF7FFFu) | w[k];
                    all &= t;
                    any |= w[k];
                }
                if (any) {  // nothing occupied in this chunk: no store (2*N_occ write bytes)
                    if (lane_uniform) {
                        if (all == 0x80008000u) {
#pragma unroll
                            for (int k = 0; k < VEC / 2; ++k) w[k] = lab;
                        } else {
#pragma unroll
                            for (int k = 0; k < VEC / 2; ++k) w[k] = nonzero_halves(w[k]) & lab;
                        }
                        vf_stg_stream(gp, pack<VEC>(w));
                    } else {
                        slow = true;
                    }
                }
            }
            if (C > 1) {  // warp-uniform
                const unsigned m = __ballot_sync(kFull, slow);
                if (slow) tasks[ntasks + __popc(m & ((1u << lane) - 1))] = (unsigned char)(it * 32 + lane);
                ntasks += __popc(m);
            }
        }

        // ---- 4. pass 2: the queued chunks are dealt round-robin to the lanes (any lane can serve any chunk: the data sits in
        //      shared memory), so a brick with a few boundary chunks costs a few balanced rounds instead of stalling whole planes
        if (ntasks) {
            __syncwarp();
            for (int q = lane; q < ntasks; q += 32) {
                const int id = tasks[q], it = id >> 5, src = id & 31;
                const int tx = x0 + it, ty = y0 + (src >> 3), tz = z0 + (src & 7) * VEC;
                unsigned w[VEC / 2], lab2[VEC / 2];
                unpack<VEC>(stage[id], w);
                if (C <= kCMax) eval_chunk<DF, VEC>(cand, C, tx, ty, tz, lab2);
                else scan_chunk<DF, VEC>(sseeds, S, tx, ty, tz, lab2);  // more survivors than slots: exact scan of the whole seed set
#pragma unroll
                for (int k = 0; k < VEC / 2; ++k) w[k] = nonzero_halves(w[k]) & lab2[k];
                vf_stg_stream(reinterpret_cast<V*>(grid + ((size_t)tx * Y + ty) * Z + tz), pack<VEC>(w));
            }
        }
        __syncwarp();  // cand[], tasks[] and stage[] are rewritten by the next brick
    }
}

// generic path: any dims, float32 compare exactly as buildCPU.  One thread per voxel.
template <int DF>
__global__ void __launch_bounds__(256) naive_generic_kernel(uint16_t* __restrict__ grid, int X, int Y, int Z, const ushort4* __restrict__ seeds_g, int S, int xb)
{
    extern __shared__ ushort4 smem[];
    for (int i = threadIdx.x; i < S; i += blockDim.x) smem[i] = seeds_g[i];
    __syncthreads();
    const size_t n = (size_t)X * Y * Z;
    for (size_t i = (size_t)xb * Y * Z + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned short own = grid[i];
        if (own == VF_VOXEL_EMPTY) continue;
        const int z = (int)(i % Z);
        const size_t r = i / Z;
        const int y = (int)(r % Y), x = (int)(r / Y);
        grid[i] = scan_all_seeds<DF>(x, y, z, smem, S, own);
    }
}

template <int DF, int VEC>
vf_status launch_brick(vf_grid* g, const ushort4* d_seeds, int S, int xb)
{
    vf_ctx* c = g->ctx;
    const int nbx = (g->X - xb + kIts - 1) / kIts, nby = (g->Y + 3) / 4, nbz = (g->Z + 8 * VEC - 1) / (8 * VEC);
    const unsigned total = (unsigned)nbx * nby * nbz;
    VF_REQUIRE((uint64_t)total * (uint64_t)max(nbz, nby) < (1ull << 32), VF_ERR_CAPACITY, "naive: grid too large for the brick index decode");
    const size_t smem = (size_t)kWarps * kIts * 32 * (VEC * 2) + ((size_t)((S + 3) & ~3) + kWarps * kCMax) * sizeof(ushort4) + (size_t)kWarps * kIts * 32;
    auto kern = naive_brick_kernel<DF, VEC>;
    if (smem > 48 * 1024) VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = min((total + kWarps - 1) / kWarps, (unsigned)c->num_sms * 4u);
    kern<<<blocks, kWarps * 32, smem, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_seeds, S, FastDiv((unsigned)nbz), FastDiv((unsigned)nby), total, xb);
    VF_LAUNCHED(c);
    return VF_OK;
}

template <int DF>
vf_status launch_generic(vf_grid* g, const ushort4* d_seeds, int S, int xb)
{
    vf_ctx* c = g->ctx;
    const size_t smem = (size_t)S * sizeof(ushort4);
    auto kern = naive_generic_kernel<DF>;
    if (smem > 48 * 1024) VF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)min(((size_t)(g->X - xb) * g->Y * g->Z + 255) / 256, (size_t)c->num_sms * 16);
    kern<<<blocks, 256, smem, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_seeds, S, xb);
    VF_LAUNCHED(c);
    return VF_OK;
}

template <int DF>
vf_status dispatch(vf_grid* g, const ushort4* d_seeds, int S, int xb, uint32_t Xfull)
{
    const uint32_t maxdim = max(Xfull, max(g->Y, g->Z));
    const bool int_exact = DF != VF_EUCLIDEAN || maxdim <= 1182;  // float sqrt injective on integer d^2
    const bool aligned = ((uintptr_t)g->d & 15) == 0;
    if (int_exact && aligned && g->Z % 8 == 0) return launch_brick<DF, 8>(g, d_seeds, S, xb);
    if (int_exact && aligned && g->Z % 4 == 0) return launch_brick<DF, 4>(g, d_seeds, S, xb);
    return launch_generic<DF>(g, d_seeds, S, xb);
}

}  // namespace

// planes xb .. g->X - 1 of the grid `g` describes (xb = 0: all of it); Xfull: the extent of the whole grid along x, which decides the arithmetic
static vf_status naive_planes(vf_grid* g, const ushort4* d_seeds, uint32_t nseeds, int dfunc, int xb, uint32_t Xfull)
{
    VF_REQUIRE(nseeds <= 8192, VF_ERR_CAPACITY, "naive: %u seeds exceed the shared-memory seed table (8192)", nseeds);
    switch (dfunc) {
    case VF_EUCLIDEAN: return dispatch<VF_EUCLIDEAN>(g, d_seeds, (int)nseeds, xb, Xfull);
    case VF_MANHATTAN: return dispatch<VF_MANHATTAN>(g, d_seeds, (int)nseeds, xb, Xfull);
    case VF_CHEBYSHEV: return dispatch<VF_CHEBYSHEV>(g, d_seeds, (int)nseeds, xb, Xfull);
    default: return vf_set_error(VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    }
}

vf_status vf_k_naive(vf_grid* g, const ushort4* d_seeds, uint32_t nseeds, int dfunc) { return naive_planes(g, d_seeds, nseeds, dfunc, 0, g->X); }

// F1 on one slab of a grid cut along x (multi-GPU): `slab` holds planes x_origin .. x_origin + slab->X - 1 of a grid of X_full planes
// (halo planes included — nearest-seed labels are pointwise, so the halo is computed, not exchanged); seeds in the coordinates of the whole grid.
extern "C" vf_status vf_fracture_naive_slab(vf_grid* slab, const uint32_t* seeds, uint32_t nseeds, int dfunc, uint32_t x_origin, uint32_t X_full)
{
    VF_REQUIRE(slab != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(slab->ctx));
    VF_REQUIRE(dfunc >= 0 && dfunc <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    VF_REQUIRE((uint64_t)x_origin + slab->X <= X_full, VF_ERR_INVALID_ARGUMENT, "naive: slab planes %u..%u lie outside a grid of %u planes", x_origin,
               x_origin + slab->X - 1, X_full);
    ushort4* d_seeds = nullptr;
    VF_TRY(vf_upload_seeds(slab->ctx, seeds, nseeds, X_full, slab->Y, slab->Z, &d_seeds));
    vf_grid view = *slab;  // the slab seen as planes x_origin .. of a grid that starts x_origin planes earlier (those planes are never touched)
    view.X = x_origin + slab->X;
    view.d = slab->d - (size_t)x_origin * slab->Y * slab->Z;
    return naive_planes(&view, d_seeds, nseeds, dfunc, (int)x_origin, X_full);
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
