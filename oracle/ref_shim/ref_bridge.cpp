// ref_bridge.cpp — C entry points over the REFERENCE'S OWN hot-path code, compiled in place from /root/reference (see Makefile):
//   Fracturer/NaiveFracturer.cpp   buildCPU (:26-68), removeIsolatedRegionsCPU (:111-150)
//   Fracturer/Seeder.cpp           uniform (:154-208), mergeSeeds (:115-152)  (+ Utilities/RandomUtilities.h, HaltonSampler.h as they are)
//   Geometry/3D/Intersections3D.h  intersect(Triangle3D&, AABB&) + helpers (:204-420; the lines are extracted by the Makefile into
//                                  oracle/_ref/*.inc because the header's other functions need the whole geometry/GL tree)
//   Geometry/3D/AABB.{h,cpp}
// Used only by tests/test_oracle_vs_ref.py to pin the oracle.  No reference source is copied into the repository; the absent
// dependencies (glm, GL, boost, the RegularGrid / ShaderList classes) are replaced by the shims under include/.
#include "stdafx.h"
#define protected public  // buildCPU / removeIsolatedRegionsCPU are protected members of the singleton
#define private public
#include "Fracturer/NaiveFracturer.h"
#undef protected
#undef private
#include "Fracturer/Seeder.h"
#include "Geometry/3D/AABB.h"
#include "Geometry/3D/Triangle3D.h"
#include "Utilities/RandomUtilities.h"

namespace Intersections3D {
bool intersect(Triangle3D& triangle, AABB& aabb);
#include "sat_decl.inc"
}  // namespace Intersections3D
#include "sat_impl.inc"

static std::vector<glm::uvec4> to_seeds(const uint32_t* s, uint32_t n)
{
    std::vector<glm::uvec4> v(n);
    for (uint32_t i = 0; i < n; ++i) v[i] = glm::uvec4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]);
    return v;
}

extern "C" {

void ref_naive_build_cpu(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t n, int dfunc)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    FractureParameters fp;
    fp._distanceFunction = dfunc;
    fp._launchGPU = false;
    fp._removeIsolatedRegions = false;
    fracturer::NaiveFracturer::getInstance()->build(g, to_seeds(seeds, n), &fp);  // -> buildCPU
}

void ref_remove_isolated_regions_cpu(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t n)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    fracturer::NaiveFracturer::getInstance()->removeIsolatedRegionsCPU(g, to_seeds(seeds, n));
}

// returns 0, or -1 on SeederSearchError; *next_uniform receives the generator's next draw (to compare RNG positions)
int ref_seed_uniform(uint16_t* grid, const uint32_t dims[3], uint32_t n, int location, int rng_seed, uint32_t* out, float* next_uniform)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    RandomUtilities::initSeed(rng_seed);
    try {
        std::vector<glm::uvec4> s = fracturer::Seeder::uniform(g, n, FractureParameters::STD_UNIFORM, static_cast<fracturer::Seeder::Location>(location));
        for (uint32_t i = 0; i < s.size(); ++i) out[4 * i] = s[i].x, out[4 * i + 1] = s[i].y, out[4 * i + 2] = s[i].z, out[4 * i + 3] = s[i].w;
    } catch (const fracturer::Seeder::SeederSearchError&) {
        return -1;
    }
    if (next_uniform) *next_uniform = RandomUtilities::getUniformRandom();
    return 0;
}

// Seeder::uniform with the HALTON generator (Seeder.cpp:18,28 + Utilities/HaltonSampler.h as it is)
int ref_seed_halton(uint16_t* grid, const uint32_t dims[3], uint32_t n, int location, uint32_t* out)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    try {
        std::vector<glm::uvec4> s = fracturer::Seeder::uniform(g, n, FractureParameters::HALTON, static_cast<fracturer::Seeder::Location>(location));
        for (uint32_t i = 0; i < s.size(); ++i) out[4 * i] = s[i].x, out[4 * i + 1] = s[i].y, out[4 * i + 2] = s[i].z, out[4 * i + 3] = s[i].w;
    } catch (const fracturer::Seeder::SeederSearchError&) {
        return -1;
    }
    return 0;
}

// Halton_sampler::sample after init_faure, straight from the reference header
float ref_halton(unsigned dimension, unsigned index)
{
    static Halton_sampler sampler = [] {
        Halton_sampler h;
        h.init_faure();
        return h;
    }();
    return sampler.sample(dimension, index);
}

// Seeder::nearSeeds (Seeder.cpp:49-113) as it is; consumes RandomUtilities' generator and this process's ::rand()
int ref_near_seeds(uint16_t* grid, const uint32_t dims[3], const uint32_t* frags, uint32_t nfrags, unsigned numImpacts, unsigned numSeeds, unsigned spreading,
                   int rng_seed, unsigned crand_seed, uint32_t* out)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    RandomUtilities::initSeed(rng_seed);
    srand(crand_seed);
    std::vector<glm::uvec4> s = fracturer::Seeder::nearSeeds(g, to_seeds(frags, nfrags), numImpacts, numSeeds, spreading);
    for (uint32_t i = 0; i < s.size(); ++i) out[4 * i] = s[i].x, out[4 * i + 1] = s[i].y, out[4 * i + 2] = s[i].z, out[4 * i + 3] = s[i].w;
    return (int)s.size();
}

void ref_merge_seeds(const uint32_t* frags, uint32_t nfrags, uint32_t* seeds, uint32_t nseeds, int dfunc)
{
    std::vector<glm::uvec4> s = to_seeds(seeds, nseeds);
    fracturer::Seeder::mergeSeeds(to_seeds(frags, nfrags), s, static_cast<fracturer::DistanceFunction>(dfunc));
    for (uint32_t i = 0; i < nseeds; ++i) seeds[4 * i + 3] = s[i].w;
}

int ref_tri_box_intersect(const float* p1, const float* p2, const float* p3, const float* bmin, const float* bmax)
{
    Triangle3D t(vec3(p1[0], p1[1], p1[2]), vec3(p2[0], p2[1], p2[2]), vec3(p3[0], p3[1], p3[2]));
    AABB box(vec3(bmin[0], bmin[1], bmin[2]), vec3(bmax[0], bmax[1], bmax[2]));
    return Intersections3D::intersect(t, box) ? 1 : 0;
}

void ref_uniform_draws(int rng_seed, int n, float* out)
{
    RandomUtilities::initSeed(rng_seed);
    for (int i = 0; i < n; ++i) out[i] = RandomUtilities::getUniformRandom();
}
}
