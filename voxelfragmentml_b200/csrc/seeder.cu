// seeder.cu — S1/S2 seed placement and the caller's block CADScene::fractureModel.
//
// Replaces fracturer::Seeder::uniform / mergeSeeds (SRC/Fracturer/Seeder.cpp:154-208, 115-152) and the seed + build + cleanup
// sequence of CADScene::fractureModel (SRC/Graphics/Application/CADScene.cpp:624-691).
//
// The reference samples on the host against the host copy of the grid, which forces a full-grid readback after voxelization.
// Here the grid stays on the device: the host draws candidate cells in exactly the reference's order (three draws per attempt,
// consumed even when rejected), a small kernel evaluates isOccupied / isBoundary (RegularGrid.cpp:543-564) for a batch of
// candidates, and the host replays the acceptance rule on the flags.  When the n-th seed is found mid-batch the generator is
// rewound to the state the reference's would have, so everything drawn afterwards (extra seeds, erosion noise) matches.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <set>
#include <vector>

#include "vf_internal.h"

namespace {

constexpr int kBatch = 4096;
constexpr uint32_t kMaxTries = 1000000;  // Seeder.h:48

__global__ void __launch_bounds__(128) seed_probe_kernel(const uint16_t* __restrict__ grid, int X, int Y, int Z, const ushort4* __restrict__ cand,
                                                         int n, uint8_t* __restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ushort4 c = cand[i];
    const int x = c.x, y = c.y, z = c.z;
    const bool occupied = grid[((size_t)x * Y + y) * Z + z] != VF_VOXEL_EMPTY;  // RegularGrid.cpp:561-564
    bool boundary = false;                                                      // RegularGrid.cpp:543-559, neighbourhoodSize 1
    const int x0 = max(x - 1, 0), x1 = min(x + 1, X - 1), y0 = max(y - 1, 0), y1 = min(y + 1, Y - 1), z0 = max(z - 1, 0), z1 = min(z + 1, Z - 1);
    for (int a = x0; a <= x1; ++a)
        for (int b = y0; b <= y1; ++b)
            for (int d = z0; d <= z1; ++d) boundary = boundary || grid[((size_t)a * Y + b) * Z + d] == VF_VOXEL_EMPTY;
    flags[i] = (occupied ? 1 : 0) | (boundary ? 2 : 0);
}

struct U3 {
    uint32_t x, y, z;
    bool operator<(const U3& r) const
    {
        if (x != r.x) return x < r.x;
        if (y != r.y) return y < r.y;
        return z < r.z;
    }
};

}  // namespace

// Halton_sampler::sample for dimensions 0..2 with Faure permutations (Utilities/HaltonSampler.h:572-632, 1416-1446): the radical
// inverse of the attempt index in base 2 (bits mirrored into a float mantissa), base 3 (20 digits) and base 5 (12 digits).  Faure's
// permutation is the identity for base 3 and (0 3 2 1 4) for base 5; the scale constants are the reference's float32 literals.
static float halton_sample(int dim, uint32_t index)
{
    if (dim == 0) {
        uint32_t mirrored = 0;
        for (uint32_t v = index, k = 0; k < 32; ++k, v >>= 1) mirrored = (mirrored << 1) | (v & 1u);
        const uint32_t word = 0x3f800000u | (mirrored >> 9);
        float f;
        std::memcpy(&f, &word, sizeof f);
        return f - 1.f;
    }
    static const uint32_t kDigit5[5] = { 0, 3, 2, 1, 4 };
    uint32_t acc = 0;
    if (dim == 1) {
        for (int d = 0; d < 20; ++d, index /= 3) acc = acc * 3 + index % 3;
        return acc * float(0x1.fffffcp-1 / 3486784401u);
    }
    for (int d = 0; d < 12; ++d, index /= 5) acc = acc * 5 + kDigit5[index % 5];
    return acc * float(0x1.fffffcp-1 / 244140625u);
}

extern "C" vf_status vf_seed_uniform(vf_grid* g, uint32_t n, int random_mode, int location, uint32_t* out, uint32_t* attempts_out)
{
    VF_REQUIRE(g != nullptr && out != nullptr, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    VF_REQUIRE(random_mode == VF_STD_UNIFORM || random_mode == VF_HALTON, VF_ERR_UNSUPPORTED,
               "seeding mode %d: STD_UNIFORM and HALTON are implemented (BOOST_NORMAL is parity-unpinned, SURVEY §8a S3)", random_mode);
    VF_REQUIRE(location >= 0 && location <= 2, VF_ERR_INVALID_ARGUMENT, "bad seed location %d", location);
    VF_REQUIRE(g->X >= 2 && g->Y >= 2 && g->Z >= 2, VF_ERR_INVALID_ARGUMENT, "grid too small to seed");
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    ushort4* d_cand = (ushort4*)((char*)c->small.ptr + (768 << 10));
    uint8_t* d_flags = (uint8_t*)(d_cand + kBatch);
    ushort4* h_cand = (ushort4*)c->pinned;
    uint8_t* h_flags = (uint8_t*)c->pinned + 65536;

    std::set<U3> seeds;  // Seeder.cpp:163 std::set with the lexicographic comparator
    const int ndx = (int)g->X - 2, ndy = (int)g->Y - 2, ndz = (int)g->Z - 2;  // :165
    uint32_t attempt = 0;
    while (seeds.size() != n) {
        VF_CUDA(vf_sync(c));  // the pinned staging areas may still be in flight
        const VfMt19937 saved = c->rng;
        const int batch = (int)std::min<uint32_t>(kBatch, kMaxTries - attempt);
        if (batch == 0) {
            if (attempts_out) *attempts_out = attempt;
            return vf_set_error(VF_ERR_SEEDER_EXHAUSTED, "Max. number of tries surpassed (%u)", kMaxTries);  // :173-174
        }
        for (int i = 0; i < batch; ++i) {  // :177-179
            int x, y, z;
            if (random_mode == VF_HALTON) {  // Seeder.cpp:28: int(sample(coord, attempt) * (max - min) + min) in float32; stateless
                x = (int)(halton_sample(0, attempt + i) * (float)(ndx + 1) + 0.0f);
                y = (int)(halton_sample(1, attempt + i) * (float)(ndy + 1) + 0.0f);
                z = (int)(halton_sample(2, attempt + i) * (float)(ndz + 1) + 0.0f);
            } else {
                x = c->rng.uniform_int(0, ndx + 1), y = c->rng.uniform_int(0, ndy + 1), z = c->rng.uniform_int(0, ndz + 1);
            }
            h_cand[i] = make_ushort4((unsigned short)x, (unsigned short)y, (unsigned short)z, 0);
        }
        VF_CUDA(cudaMemcpyAsync(d_cand, h_cand, batch * sizeof(ushort4), cudaMemcpyHostToDevice, c->stream));
        seed_probe_kernel<<<(batch + 127) / 128, 128, 0, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_cand, batch, d_flags);
        VF_LAUNCHED(c);
        VF_CUDA(cudaMemcpyAsync(h_flags, d_flags, batch, cudaMemcpyDeviceToHost, c->stream));
        VF_CUDA(vf_sync(c));
        int used = batch;
        for (int i = 0; i < batch; ++i) {
            const U3 v = { h_cand[i].x, h_cand[i].y, h_cand[i].z };
            const bool occupied = h_flags[i] & 1, boundary = h_flags[i] & 2;
            const bool isFree = seeds.find(v) == seeds.end();
            if (occupied && isFree)
                if ((location == VF_OUTER && boundary) || (location == VF_INNER && !boundary) || location == VF_BOTH) seeds.insert(v);  // :190-192
            if (seeds.size() == n) {
                used = i + 1;
                break;
            }
        }
        attempt += used;
        if (used != batch && random_mode == VF_STD_UNIFORM) {  // rewind to the reference's RNG position: exactly 3 draws per attempt
            c->rng = saved;
            for (int i = 0; i < 3 * used; ++i) c->rng.next();
        }
    }
    uint32_t label = VF_VOXEL_FREE + 1, k = 0;  // :202
    for (const U3& s : seeds) {
        out[4 * k] = s.x, out[4 * k + 1] = s.y, out[4 * k + 2] = s.z, out[4 * k + 3] = label++;
        ++k;
    }
    if (attempts_out) *attempts_out = attempt;
    return VF_OK;
}

// Seeder::nearSeeds (Seeder.cpp:49-113): numImpacts impacts, each on a fragment drawn from `frags`, each adding a random share of
// numSeeds (= _biasSeeds) boundary cells around it; the offsets come from RandomUtilities::getBiasedRandomInt (RandomUtilities.h:
// 146-156), i.e. from the C runtime's rand() — the MSVC LCG on the reference's platform, kept per context (vf_crand_seed).
// As in vf_seed_uniform the host draws the candidates in the reference's order, a kernel evaluates isOccupied / isBoundary for a
// batch, and the host replays the acceptance rule; the LCG is put back to the state it had after the accepted candidate.
// The reference's search loop has no bound; this one gives up after 1e6 candidates per impact (VF_ERR_SEEDER_EXHAUSTED).
extern "C" vf_status vf_seed_near(vf_grid* g, const uint32_t* frags, uint32_t nfrags, uint32_t num_impacts, uint32_t num_seeds, uint32_t spreading,
                                  uint32_t* out, uint32_t capacity, uint32_t* count_out)
{
    VF_REQUIRE(g && frags && out && count_out, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_REQUIRE(nfrags > 0 && spreading > 0, VF_ERR_INVALID_ARGUMENT, "nearSeeds needs fragments and a positive spreading");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    const uint32_t dims[3] = { g->X, g->Y, g->Z };
    for (int q = 0; q < 3; ++q) VF_REQUIRE(dims[q] / spreading > 0, VF_ERR_INVALID_ARGUMENT, "spreading %u exceeds a grid dimension (rand() %% 0 in the reference)", spreading);
    VF_TRY(vf_scratch_reserve(c, c->small, 1 << 20));
    ushort4* d_cand = (ushort4*)((char*)c->small.ptr + (768 << 10));
    uint8_t* d_flags = (uint8_t*)(d_cand + kBatch);
    ushort4* h_cand = (ushort4*)c->pinned;
    uint8_t* h_flags = (uint8_t*)c->pinned + 65536;
    std::vector<uint32_t> after(kBatch);  // LCG word after each candidate of the batch

    std::set<U3> seeds;
    unsigned numPendingSeeds = num_seeds;
    const uint32_t half[3] = { dims[0] / 2, dims[1] / 2, dims[2] / 2 };
    const float minDiv = (float)(std::min(dims[0], std::min(dims[1], dims[2])) / 2);
    auto biased = [&](uint32_t dim) {
        int number = 0;
        const int mx = (int)dim / (int)spreading;
        for (uint32_t i = 0; i < spreading; ++i) number += vf_crand_next(c) % mx;
        return (uint32_t)number;
    };
    for (uint32_t impact = 0; impact < num_impacts; ++impact) {
        const uint32_t* frag = frags + 4 * (size_t)c->rng.uniform_int(0, (int)nfrags - 1);  // :67
        const unsigned nseeds = (unsigned)c->rng.uniform_int(1, (int)numPendingSeeds);       // :68
        unsigned currentSeeds = 0;
        uint64_t tries = 0;
        while (currentSeeds != nseeds) {
            VF_REQUIRE(tries < kMaxTries, VF_ERR_SEEDER_EXHAUSTED, "nearSeeds: no boundary cell near fragment (%u, %u, %u) after %u candidates", frag[0], frag[1],
                       frag[2], kMaxTries);
            VF_CUDA(vf_sync(c));
            for (int i = 0; i < kBatch; ++i) {  // :74-83
                const uint32_t x = (frag[0] + (half[0] - biased(dims[0])) + dims[0]) % dims[0];
                const uint32_t y = (frag[1] + (half[1] - biased(dims[1])) + dims[1]) % dims[1];
                const uint32_t z = (frag[2] + (half[2] - biased(dims[2])) + dims[2]) % dims[2];
                h_cand[i] = make_ushort4((unsigned short)x, (unsigned short)y, (unsigned short)z, 0);
                after[i] = c->crand;
            }
            VF_CUDA(cudaMemcpyAsync(d_cand, h_cand, kBatch * sizeof(ushort4), cudaMemcpyHostToDevice, c->stream));
            seed_probe_kernel<<<(kBatch + 127) / 128, 128, 0, c->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, d_cand, kBatch, d_flags);
            VF_LAUNCHED(c);
            VF_CUDA(cudaMemcpyAsync(h_flags, d_flags, kBatch, cudaMemcpyDeviceToHost, c->stream));
            VF_CUDA(vf_sync(c));
            for (int i = 0; i < kBatch; ++i) {
                const U3 v = { h_cand[i].x, h_cand[i].y, h_cand[i].z };
                const float dx = (float)v.x - (float)frag[0], dy = (float)v.y - (float)frag[1], dz = (float)v.z - (float)frag[2];
                if (sqrtf(dx * dx + dy * dy + dz * dz) > minDiv) continue;  // :87
                if ((h_flags[i] & 1) && (h_flags[i] & 2) && seeds.find(v) == seeds.end()) {
                    seeds.insert(v);
                    if (++currentSeeds == nseeds) {
                        c->crand = after[i];
                        break;
                    }
                }
            }
            tries += kBatch;
        }
        numPendingSeeds -= nseeds;
    }
    const uint32_t total = nfrags + (uint32_t)seeds.size();
    VF_REQUIRE(total <= capacity, VF_ERR_CAPACITY, "nearSeeds: %u seeds do not fit the output (%u)", total, capacity);
    if (out != frags) std::memmove(out, frags, 16 * (size_t)nfrags);  // :103 result = frags
    uint32_t nseed = out[4 * (size_t)(nfrags - 1) + 3], k = nfrags;
    for (const U3& s : seeds) out[4 * k] = s.x, out[4 * k + 1] = s.y, out[4 * k + 2] = s.z, out[4 * k + 3] = ++nseed, ++k;
    *count_out = total;
    return VF_OK;
}

extern "C" vf_status vf_merge_seeds(const uint32_t* frags, uint32_t nfrags, uint32_t* seeds, uint32_t nseeds, int dfunc)
{
    // Seeder.cpp:115-152, float32 distances on integer coordinates, strict '<' (lowest fragment index wins ties)
    VF_REQUIRE(frags && seeds && nfrags > 0, VF_ERR_INVALID_ARGUMENT, "mergeSeeds: null or empty input");
    VF_REQUIRE(dfunc >= 0 && dfunc <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function %d", dfunc);
    int idFragment[1 << VF_ID_POSITION];
    std::memset(idFragment, 0, sizeof(idFragment));
    for (uint32_t si = 0; si < nseeds; ++si) {
        uint32_t* seed = seeds + 4 * si;
        float mn = FLT_MAX;
        int nearest = -1;
        for (uint32_t i = 0; i < nfrags; ++i) {
            const float dx = (float)seed[0] - (float)frags[4 * i], dy = (float)seed[1] - (float)frags[4 * i + 1], dz = (float)seed[2] - (float)frags[4 * i + 2];
            float dist;
            if (dfunc == VF_EUCLIDEAN) dist = sqrtf(dx * dx + dy * dy + dz * dz);
            else if (dfunc == VF_MANHATTAN) dist = fabsf(dx) + fabsf(dy) + fabsf(dz);
            else dist = fmaxf(fabsf(dx), fmaxf(fabsf(dy), fabsf(dz)));
            if (dist < mn) mn = dist, nearest = (int)i;
        }
        const uint32_t fw = frags[4 * nearest + 3];
        seed[3] = fw | ((uint32_t)(++idFragment[fw & 0xFFu]) << VF_ID_POSITION);  // :150
    }
    return VF_OK;
}

extern "C" vf_status vf_make_seeds(vf_grid* g, uint32_t n, uint32_t n_extra, int random_mode, int merge_dfunc, uint32_t* out, uint32_t cap,
                                   uint32_t* count_out)
{
    // CADScene.cpp:626-655, numImpacts == 0 branch
    VF_REQUIRE(g && out && count_out, VF_ERR_INVALID_ARGUMENT, "null argument");
    const uint32_t total = n + (n_extra ? n + n_extra : 0);
    VF_REQUIRE(total <= cap, VF_ERR_CAPACITY, "seed buffer too small (%u < %u)", cap, total);
    VF_TRY(vf_seed_uniform(g, n, random_mode, VF_OUTER, out, nullptr));
    if (n_extra > 0) {
        std::vector<uint32_t> extra(4 * (size_t)(n + n_extra));
        VF_TRY(vf_seed_uniform(g, n_extra, random_mode, VF_BOTH, extra.data() + 4 * n, nullptr));
        std::memcpy(extra.data(), out, 16 * (size_t)n);                              // :651
        VF_TRY(vf_merge_seeds(out, n, extra.data(), n + n_extra, merge_dfunc));       // :653
        std::memcpy(out + 4 * (size_t)n, extra.data(), 16 * (size_t)(n + n_extra));  // :654
    }
    *count_out = total;
    return VF_OK;
}

extern "C" vf_status vf_fracture_model(vf_grid* g, const vf_params* p, uint32_t* seeds_out, uint32_t* nseeds_out, vf_flood_stats* stats)
{
    VF_REQUIRE(g && p, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    VF_REQUIRE(p->numSeeds > 0 && p->numExtraSeeds >= 0 && p->numImpacts >= 0, VF_ERR_INVALID_ARGUMENT, "bad seed counts");
    // seeds = uniform(OUTER) [-> nearSeeds] (:629-639); with extra seeds: [originals..., copies of the originals..., extras...] (:647-655)
    const uint32_t nmax = (uint32_t)p->numSeeds + (p->numImpacts > 0 ? (uint32_t)std::max(p->biasSeeds, 0) : 0u);
    const uint32_t cap = nmax * 2 + (uint32_t)p->numExtraSeeds;
    std::vector<uint32_t> seeds(4 * (size_t)cap);
    uint32_t ns = (uint32_t)p->numSeeds;
    VF_TRY(vf_seed_uniform(g, ns, p->seedingRandom, VF_OUTER, seeds.data(), nullptr));
    if (p->numImpacts > 0) VF_TRY(vf_seed_near(g, seeds.data(), ns, (uint32_t)p->numImpacts, (uint32_t)p->biasSeeds, (uint32_t)p->biasFocus, seeds.data(), nmax, &ns));
    if (p->numExtraSeeds > 0) {
        const uint32_t ne = (uint32_t)p->numExtraSeeds;
        std::vector<uint32_t> extra(4 * (size_t)(ns + ne));
        VF_TRY(vf_seed_uniform(g, ne, p->seedingRandom, VF_BOTH, extra.data() + 4 * (size_t)ns, nullptr));
        std::memcpy(extra.data(), seeds.data(), 16 * (size_t)ns);
        VF_TRY(vf_merge_seeds(seeds.data(), ns, extra.data(), ns + ne, p->mergeSeedsDistanceFunction));
        std::memcpy(seeds.data() + 4 * (size_t)ns, extra.data(), 16 * (size_t)(ns + ne));
        ns += ns + ne;
    }
    if (p->fractureAlgorithm == VF_NAIVE) {
        VF_REQUIRE(p->distanceFunction >= 0 && p->distanceFunction <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function");  // CADScene.cpp:665
        VF_TRY(vf_fracture_naive(g, seeds.data(), ns, p->distanceFunction));
        if (p->removeIsolatedRegions) VF_TRY(vf_remove_isolated_regions(g, seeds.data(), ns));  // NaiveFracturer.cpp:100-103 (CPU semantics)
    } else if (p->fractureAlgorithm == VF_FLOOD) {
        VF_REQUIRE(p->distanceFunction >= 0 && p->distanceFunction <= 2, VF_ERR_INVALID_DISTANCE, "Invalid distance function");
        VF_TRY(vf_fracture_flood(g, seeds.data(), ns, p->distanceFunction, p->floodIdBits, stats));
    } else {
        return vf_set_error(VF_ERR_UNSUPPORTED, "VORONOI (CGAL Delaunay, SRC/Graphics/Core/Voronoi.cpp) is outside the hot path");
    }
    if (p->erode) {  // CADScene.cpp:679-684
        std::vector<float> noise(1000000);  // RegularGrid.cpp:126
        VF_TRY(vf_fill_noise(c, noise.data(), (uint32_t)noise.size()));
        VF_TRY(vf_erode(g, p->erosionConvolution, (uint32_t)p->erosionSize, (uint32_t)p->erosionIterations, p->erosionProbability, p->erosionThreshold,
                        noise.data(), (uint32_t)noise.size(), p->erodeBoundaryMode));
    } else {
        VF_TRY(vf_detect_boundaries(g, 1));  // :687
    }
    if (seeds_out) std::memcpy(seeds_out, seeds.data(), 16 * (size_t)ns);
    if (nseeds_out) *nseeds_out = ns;
    return VF_OK;
}
