// mirror_demo.cpp — exercises include/voxfrag.hpp exactly the way CADScene::fractureModel drives the reference classes
// (CADScene.cpp:624-691): Seeder::uniform -> [uniform(BOTH) + mergeSeeds] -> Fracturer::build -> detectBoundaries -> undoMask ->
// exportGrid(RLE).  Input: an .rle occupancy grid.  Prints an FNV-1a checksum of the labelled grid and the seed list so that
// tests/test_cpp_mirror_gpu.py can compare it with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

#include "voxfrag.hpp"

using namespace voxfrag;

int main(int argc, char** argv)
{
    if (argc == 5 && std::strcmp(argv[1], "dataset") == 0) {
        // main.cpp:37-39 in GENERATE_DATASET mode: FragmentationProcedure procedure; scene->generateDataset(procedure, folder, ext, dest)
        Context ctx(0);
        FragmentationProcedure procedure;
        procedure._fragmentInterval = { 2, 3 }, procedure._iterationInterval = { 2, 1 };
        procedure._fractureParameters._clampVoxelMetricUnit = procedure._fractureParameters._voxelPerMetricUnit = std::atoi(argv[4]);
        ctx.initSeed(procedure._fractureParameters._seed);
        const vf_dataset_stats st = generateDataset(ctx, procedure, argv[2], procedure._searchExtension, argv[3]);
        std::printf("models %llu fragmentations %llu fragments %llu files %llu\n", (unsigned long long)st.models, (unsigned long long)st.fragmentations,
                    (unsigned long long)st.fragments, (unsigned long long)st.files);
        return 0;
    }
    if (argc < 4) {
        std::fprintf(stderr, "usage: mirror_demo in.rle out_basename naive|flood\n");
        return 2;
    }
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<unsigned char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    uint32_t dims[3];
    std::memcpy(dims, bytes.data(), 12);
    std::vector<RegularGrid::CellGrid> occ((size_t)dims[0] * dims[1] * dims[2]);
    size_t off = 0;
    for (size_t p = 12; p + 6 <= bytes.size(); p += 6) {
        uint16_t v;
        uint32_t r;
        std::memcpy(&v, &bytes[p], 2);
        std::memcpy(&r, &bytes[p + 2], 4);
        for (uint32_t i = 0; i < r; ++i) occ[off++]._value = v;
    }
    try {
        Context ctx(0);
        FractureParameters params;  // reference defaults: FLOOD, CHEBYSHEV, 8 seeds, 16 extra seeds, seed 80
        const bool naive = std::strcmp(argv[3], "naive") == 0;
        if (naive) {
            params._fractureAlgorithm = FractureParameters::NAIVE;
            params._distanceFunction = FractureParameters::EUCLIDEAN;
            params._numExtraSeeds = 0;
        }
        ctx.initSeed(params._seed);  // CADScene.cpp:36-37
        RegularGrid grid(ctx, ivec3{ (int)dims[0], (int)dims[1], (int)dims[2] });
        grid.swap(occ.data(), occ.size());

        // --- CADScene::fractureModel, spelled out with the reference's own call sequence
        auto dfunc = static_cast<fracturer::DistanceFunction>(params._distanceFunction);
        std::vector<uvec4> seeds = fracturer::Seeder::uniform(grid, params._numSeeds, params._seedingRandom, fracturer::Seeder::OUTER);
        if (params._numExtraSeeds > 0) {
            auto mergeDFunc = static_cast<fracturer::DistanceFunction>(params._mergeSeedsDistanceFunction);
            auto extraSeeds = fracturer::Seeder::uniform(grid, params._numExtraSeeds, params._seedingRandom, fracturer::Seeder::BOTH);
            extraSeeds.insert(extraSeeds.begin(), seeds.begin(), seeds.end());
            fracturer::Seeder::mergeSeeds(seeds, extraSeeds, mergeDFunc);
            seeds.insert(seeds.end(), extraSeeds.begin(), extraSeeds.end());
        }
        // the host-side cell accessors of the reference (Seeder::uniform's OUTER rule, RegularGrid.cpp:543-564), on the host copy
        for (int i = 0; i < params._numSeeds; ++i)
            if (!grid.isOccupied(seeds[i].x, seeds[i].y, seeds[i].z) || grid.isEmpty(seeds[i].x, seeds[i].y, seeds[i].z) ||
                !grid.isBoundary(seeds[i].x, seeds[i].y, seeds[i].z)) {
                std::fprintf(stderr, "seed %d is not an occupied OUTER cell\n", i);
                return 4;
            }
        fracturer::Fracturer* fracturer = naive ? static_cast<fracturer::Fracturer*>(fracturer::NaiveFracturer::getInstance())
                                                : static_cast<fracturer::Fracturer*>(fracturer::FloodFracturer::getInstance());
        if (!fracturer->setDistanceFunction(dfunc)) return 3;
        fracturer->build(grid, seeds, &params);
        grid.detectBoundaries(1);
        grid.undoMask();
        grid.exportGrid(argv[2], true, FractureParameters::RLE);

        grid.updateGrid();
        uint64_t h = 1469598103934665603ull;
        const RegularGrid::CellGrid* d = grid.data();
        for (size_t i = 0; i < grid.length(); ++i) {
            h = (h ^ (d[i]._value & 0xFF)) * 1099511628211ull;
            h = (h ^ (d[i]._value >> 8)) * 1099511628211ull;
        }
        std::printf("grid_fnv1a %016llx\nseeds %zu\noccupied %u\n", (unsigned long long)h, seeds.size(), grid.numOccupiedVoxels());
        for (const uvec4& s : seeds) std::printf("seed %u %u %u %u\n", s.x, s.y, s.z, s.w);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
