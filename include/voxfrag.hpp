// voxfrag.hpp — header-only C++ host mirror of the reference's operator surface for the hot path, over the C ABI of voxfrag.h.
//
// Same class names, method names, argument meaning and error behaviour as the reference, so that the call sites in
// CADScene::allocateMeshGrid / fractureModel / generateDataset (SRC/Graphics/Application/CADScene.cpp:545-561, 624-691, 209-507)
// compile against it with `using namespace voxfrag;`.  No arithmetic happens here: every method is one call
// into libvoxfrag.so (hand-written CUDA, sm_100a).  There is no CPU fallback; errors surface as exceptions on this side of the ABI.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "voxfrag.h"

namespace voxfrag {

// glm::uvec4 / glm::ivec3 / glm::vec3 stand-ins with identical memory layout (the reference passes seeds as std::vector<glm::uvec4>)
struct uvec4 { uint32_t x, y, z, w; };
struct uvec3 { uint32_t x, y, z; };
struct ivec3 { int32_t x, y, z; };
struct ivec2 { int32_t x, y; };
struct vec3 { float x, y, z; };

struct AABB {  // SRC/Geometry/3D/AABB.h: only min()/max()/size() are used on the path
    vec3 _min{ -0.5f, -0.5f, -0.5f }, _max{ 0.5f, 0.5f, 0.5f };
    vec3 min() const { return _min; }
    vec3 max() const { return _max; }
    vec3 size() const { return { _max.x - _min.x, _max.y - _min.y, _max.z - _min.z }; }
};

class Error : public std::runtime_error {
public:
    Error(vf_status s, const std::string& m) : std::runtime_error(m), status(s) {}
    vf_status status;
};

inline void check(vf_status s)
{
    if (s != VF_OK) throw Error(s, vf_last_error());
}

// SRC/Graphics/Core/FractureParameters.h:14-146 — same field names (leading underscore kept), same defaults.
struct FractureParameters {
    enum FractureAlgorithm : uint8_t { NAIVE, FLOOD, VORONOI, BASE_ALGORITHMS };
    enum DistanceFunction : uint8_t { EUCLIDEAN, MANHATTAN, CHEBYSHEV, DISTANCE_FUNCTIONS };
    enum RandomUniformType { STD_UNIFORM, HALTON, BOOST_NORMAL_DISTRIBUTION, NUM_RANDOM_FUNCTIONS };
    enum ErosionType { SQUARE, ELLIPSE, CROSS, NUM_EROSION_CONVOLUTIONS };
    enum NeighbourhoodType { VON_NEUMANN, MOORE, NUM_NEIGHBOURHOODS };
    enum ExportGrid { RLE, QUADSTACK, VOX, UNCOMPRESSED_BINARY, NUM_GRID_EXTENSIONS };

    int _biasFocus, _biasSeeds, _clampVoxelMetricUnit;
    bool _erode;
    int _erosionConvolution, _erosionIterations;
    float _erosionProbability;
    int _erosionSize;
    float _erosionThreshold;
    int _fractureAlgorithm, _distanceFunction;
    bool _launchGPU;
    int _mergeSeedsDistanceFunction, _neighbourhoodType, _numExtraSeeds, _numImpacts, _numSeeds;
    bool _removeIsolatedRegions;
    int _seed, _seedingRandom, _voxelPerMetricUnit;
    ivec3 _voxelizationSize;
    int _exportGridExtension;
    int _floodIdBits = 0, _erodeBoundaryMode = 0;  // extensions, 0 = reference behaviour

    FractureParameters()
    {
        vf_params p;
        vf_params_default(&p);
        from_c(p);
    }
    void from_c(const vf_params& p)
    {
        _biasFocus = p.biasFocus, _biasSeeds = p.biasSeeds, _clampVoxelMetricUnit = p.clampVoxelMetricUnit, _erode = p.erode != 0;
        _erosionConvolution = p.erosionConvolution, _erosionIterations = p.erosionIterations, _erosionProbability = p.erosionProbability;
        _erosionSize = p.erosionSize, _erosionThreshold = p.erosionThreshold, _fractureAlgorithm = p.fractureAlgorithm;
        _distanceFunction = p.distanceFunction, _launchGPU = p.launchGPU != 0, _mergeSeedsDistanceFunction = p.mergeSeedsDistanceFunction;
        _neighbourhoodType = p.neighbourhoodType, _numExtraSeeds = p.numExtraSeeds, _numImpacts = p.numImpacts, _numSeeds = p.numSeeds;
        _removeIsolatedRegions = p.removeIsolatedRegions != 0, _seed = p.seed, _seedingRandom = p.seedingRandom;
        _voxelPerMetricUnit = p.voxelPerMetricUnit;
        _voxelizationSize = { p.voxelizationSize[0], p.voxelizationSize[1], p.voxelizationSize[2] };
        _exportGridExtension = p.exportGridExtension, _floodIdBits = p.floodIdBits, _erodeBoundaryMode = p.erodeBoundaryMode;
    }
    vf_params to_c() const
    {
        vf_params p;
        vf_params_default(&p);
        p.biasFocus = _biasFocus, p.biasSeeds = _biasSeeds, p.clampVoxelMetricUnit = _clampVoxelMetricUnit, p.erode = _erode;
        p.erosionConvolution = _erosionConvolution, p.erosionIterations = _erosionIterations, p.erosionProbability = _erosionProbability;
        p.erosionSize = _erosionSize, p.erosionThreshold = _erosionThreshold, p.fractureAlgorithm = _fractureAlgorithm;
        p.distanceFunction = _distanceFunction, p.launchGPU = _launchGPU, p.mergeSeedsDistanceFunction = _mergeSeedsDistanceFunction;
        p.neighbourhoodType = _neighbourhoodType, p.numExtraSeeds = _numExtraSeeds, p.numImpacts = _numImpacts, p.numSeeds = _numSeeds;
        p.removeIsolatedRegions = _removeIsolatedRegions, p.seed = _seed, p.seedingRandom = _seedingRandom;
        p.voxelPerMetricUnit = _voxelPerMetricUnit;
        p.voxelizationSize[0] = _voxelizationSize.x, p.voxelizationSize[1] = _voxelizationSize.y, p.voxelizationSize[2] = _voxelizationSize.z;
        p.exportGridExtension = _exportGridExtension, p.floodIdBits = _floodIdBits, p.erodeBoundaryMode = _erodeBoundaryMode;
        return p;
    }
};

// One GPU + stream + scratch (stands in for the GL context / ShaderList singletons).  RandomUtilities' process-global
// generator (SRC/Utilities/RandomUtilities.h:17-18) lives here.
class Context {
public:
    explicit Context(int device = 0) { check(vf_ctx_create(device, &_h)); }
    ~Context() { vf_ctx_destroy(_h); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    vf_ctx* handle() const { return _h; }
    void initSeed(int seed) { check(vf_rng_seed(_h, (uint32_t)seed)); }  // RandomUtilities::initSeed
    float getUniformRandom() { return vf_rng_uniform(_h); }
    void synchronize() { check(vf_ctx_synchronize(_h)); }

private:
    vf_ctx* _h = nullptr;
};

// SRC/DataStructures/RegularGrid.h:15-367 (hot-path subset).  The CPU vector + SSBO mirror pair of the reference collapses
// into one device-resident grid; data() downloads on demand.
class RegularGrid {
public:
    struct CellGrid { uint16_t _value; };

    RegularGrid(Context& ctx, const ivec3& subdivisions) : _ctx(ctx)
    {
        check(vf_grid_create(ctx.handle(), (uint32_t)subdivisions.x, (uint32_t)subdivisions.y, (uint32_t)subdivisions.z, &_h));
    }
    RegularGrid(Context& ctx, const AABB& aabb, const ivec3& subdivisions) : RegularGrid(ctx, subdivisions) { setAABB(aabb, subdivisions); }
    ~RegularGrid() { vf_grid_destroy(_h); }
    RegularGrid(const RegularGrid&) = delete;  // RegularGrid.h: copy constructor deleted in the reference too
    vf_grid* handle() const { return _h; }
    Context& context() const { return _ctx; }

    void setAABB(const AABB& aabb, const ivec3& gridDims)
    {
        const float mn[3] = { aabb._min.x, aabb._min.y, aabb._min.z }, mx[3] = { aabb._max.x, aabb._max.y, aabb._max.z };
        check(vf_grid_set_aabb(_h, mn, mx, (uint32_t)gridDims.x, (uint32_t)gridDims.y, (uint32_t)gridDims.z));
    }
    uvec3 getNumSubdivisions() const
    {
        uint32_t d[3];
        check(vf_grid_dims(_h, d));
        return { d[0], d[1], d[2] };
    }
    size_t length() const
    {
        const uvec3 d = getNumSubdivisions();
        return (size_t)d.x * d.y * d.z;
    }
    // fill(Model3D*): vertices float[nv][3], faces uint32[nf][3]
    void fill(const float* vertices, uint32_t numVertices, const uint32_t* faces, uint32_t numFaces) { check(vf_voxelize(_h, vertices, numVertices, faces, numFaces)); }
    // fill(Model3D*) with the Tetravoxelizer occupancy the reference computes today (solid interiors); returns the FREE cell count
    uint64_t fillSolid(const float* vertices, uint32_t numVertices, const uint32_t* faces, uint32_t numFaces)
    {
        uint64_t occupied = 0;
        check(vf_voxelize_solid(_h, vertices, numVertices, faces, numFaces, &occupied));
        return occupied;
    }
    // MarchingCubes::triangulateFieldGPU for one fragment label (RegularGrid::toTriangleMesh runs it per value, RegularGrid.cpp:482-483):
    // vertices = xyz + boundary flag, faces = three vertex numbers + boundary flag
    void triangulateField(uint16_t targetValue, std::vector<float>& vertices4, std::vector<uint32_t>& faces4, const vf_mc_params* params = nullptr)
    {
        vf_mesh* m = nullptr;
        check(vf_marching_cubes(_h, targetValue, params, &m));
        uint32_t nv = 0, nf = 0;
        vf_mesh_counts(m, &nv, &nf);
        vertices4.resize(4 * (size_t)nv), faces4.resize(4 * (size_t)nf);
        const vf_status s = vf_mesh_download(m, vertices4.data(), faces4.data());
        vf_mesh_destroy(m);
        check(s);
    }
    void detectBoundaries(int boundarySize) { check(vf_detect_boundaries(_h, boundarySize)); }
    void erode(FractureParameters::ErosionType type, uint32_t convolutionSize, uint16_t numIterations, float erosionProbability, float erosionThreshold,
               int boundaryMode = 0)
    {
        std::vector<float> noise(1000000);  // RegularGrid.cpp:126 fillNoiseBuffer(noiseBuffer, 1e6)
        check(vf_fill_noise(_ctx.handle(), noise.data(), (uint32_t)noise.size()));
        check(vf_erode(_h, (int)type, convolutionSize, numIterations, erosionProbability, erosionThreshold, noise.data(), (uint32_t)noise.size(), boundaryMode));
    }
    void removeIsolatedRegions() { check(vf_remove_isolated_regions_grid(_h)); }
    void undoMask() { check(vf_undo_mask(_h)); }
    void resetFilling() { check(vf_reset_filling(_h)); }
    void homogenize() { check(vf_homogenize(_h)); }
    // the `.rle` byte stream (exportRLE) with the runs found on the device
    std::vector<uint8_t> encodeRLE()
    {
        uint64_t need = 0;
        check(vf_grid_encode_rle(_h, nullptr, 0, &need));
        std::vector<uint8_t> bytes(need);
        check(vf_grid_encode_rle(_h, bytes.data(), bytes.size(), &need));
        return bytes;
    }
    void exportGrid(const std::string& filename, bool squared, FractureParameters::ExportGrid exportType) { check(vf_export(_h, filename.c_str(), (int)exportType, squared)); }
    // host access: updateGrid() downloads, updateSSBO() uploads, data() is the last downloaded copy
    void updateGrid()
    {
        _host.resize(length());
        check(vf_grid_download(_h, reinterpret_cast<uint16_t*>(_host.data())));
    }
    void updateSSBO() { check(vf_grid_upload(_h, reinterpret_cast<const uint16_t*>(_host.data()))); }
    CellGrid* data()
    {
        if (_host.size() != length()) updateGrid();
        return _host.data();
    }
    void swap(const CellGrid* src, size_t n)
    {
        _host.assign(src, src + n);
        updateSSBO();
    }
    uint16_t at(int x, int y, int z)
    {
        const uvec3 d = getNumSubdivisions();
        return data()[((size_t)x * d.y + y) * d.z + z]._value;
    }
    static unsigned getPositionIndex(int x, int y, int z, const uvec3& numDivs) { return x * numDivs.y * numDivs.z + y * numDivs.z + z; }
    // Host-side cell accessors of the reference (RegularGrid.cpp:523-531, 543-569, 1031-1039): they act on the host copy, as `_grid`
    // does there; call updateSSBO() to publish set() calls to the device and updateGrid() to refresh the copy after device work.
    void set(int x, int y, int z, uint16_t i)
    {
        const uvec3 d = getNumSubdivisions();
        data()[((size_t)x * d.y + y) * d.z + z]._value = i;
    }
    bool isOccupied(int x, int y, int z) { return at(x, y, z) != VF_VOXEL_EMPTY; }
    bool isEmpty(int x, int y, int z) { return at(x, y, z) == VF_VOXEL_EMPTY; }
    bool isBoundary(int x, int y, int z, int neighbourhoodSize = 1)  // "some cell of the clamped box is EMPTY" (RegularGrid.cpp:543-559)
    {
        if (neighbourhoodSize % 2 == 0) ++neighbourhoodSize;
        const uvec3 d = getNumSubdivisions();
        auto clampi = [](int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); };
        const int x0 = clampi(x - neighbourhoodSize, (int)d.x - 1), x1 = clampi(x + neighbourhoodSize, (int)d.x - 1);
        const int y0 = clampi(y - neighbourhoodSize, (int)d.y - 1), y1 = clampi(y + neighbourhoodSize, (int)d.y - 1);
        const int z0 = clampi(z - neighbourhoodSize, (int)d.z - 1), z1 = clampi(z + neighbourhoodSize, (int)d.z - 1);
        const CellGrid* g = data();
        for (int a = x0; a <= x1; ++a)
            for (int b = y0; b <= y1; ++b)
                for (int c = z0; c <= z1; ++c)
                    if (g[((size_t)a * d.y + b) * d.z + c]._value == VF_VOXEL_EMPTY) return true;
        return false;
    }
    // countValues + numOccupiedVoxels
    size_t countValues(std::vector<uint32_t>& countsByLabel)
    {
        countsByLabel.assign(VF_HISTOGRAM_BINS, 0);
        uint64_t occ = 0;
        check(vf_histogram(_h, countsByLabel.data(), &occ));
        size_t distinct = 0;
        for (uint32_t c : countsByLabel) distinct += c != 0;
        return distinct;
    }
    // countValues followed by undoMask (the grid side of CADScene::prepareScene, CADScene.cpp:813-832) in one pass over the grid
    size_t countValuesUndoMask(std::vector<uint32_t>& countsByLabel)
    {
        countsByLabel.assign(VF_HISTOGRAM_BINS, 0);
        uint64_t occ = 0;
        check(vf_histogram_undo_mask(_h, countsByLabel.data(), &occ));
        size_t distinct = 0;
        for (uint32_t c : countsByLabel) distinct += c != 0;
        return distinct;
    }
    unsigned numOccupiedVoxels()
    {
        std::vector<uint32_t> c(VF_HISTOGRAM_BINS);
        uint64_t occ = 0;
        check(vf_histogram(_h, c.data(), &occ));
        return (unsigned)occ;
    }

private:
    Context& _ctx;
    vf_grid* _h = nullptr;
    std::vector<CellGrid> _host;
};

namespace fracturer {

enum DistanceFunction : uint32_t { EUCLIDEAN_DISTANCE = 0, MANHATTAN_DISTANCE = 1, CHEBYSHEV_DISTANCE = 2 };  // Fracturer.h:10-14

class Fracturer {  // Fracturer.h:22-51
public:
    virtual ~Fracturer() {}
    virtual void build(RegularGrid& grid, const std::vector<uvec4>& seeds, FractureParameters* fractParameters) = 0;
    virtual void destroy() {}
    virtual void init(FractureParameters*) {}
    virtual void prepareSSBOs(FractureParameters*) {}
    virtual bool setDistanceFunction(DistanceFunction dfunc) = 0;
};

class NaiveFracturer : public Fracturer {  // NaiveFracturer.cpp
public:
    static NaiveFracturer* getInstance()
    {
        static NaiveFracturer inst;
        return &inst;
    }
    void build(RegularGrid& grid, const std::vector<uvec4>& seeds, FractureParameters* fp) override
    {
        check(vf_fracture_naive(grid.handle(), reinterpret_cast<const uint32_t*>(seeds.data()), (uint32_t)seeds.size(), (int)_dfunc));
        if (fp && fp->_removeIsolatedRegions) check(vf_remove_isolated_regions(grid.handle(), reinterpret_cast<const uint32_t*>(seeds.data()), (uint32_t)seeds.size()));
    }
    bool setDistanceFunction(DistanceFunction dfunc) override
    {
        _dfunc = dfunc;
        return true;
    }

private:
    DistanceFunction _dfunc = EUCLIDEAN_DISTANCE;
};

class FloodFracturer : public Fracturer {  // FloodFracturer.cpp
public:
    static FloodFracturer* getInstance()
    {
        static FloodFracturer inst;
        return &inst;
    }
    void build(RegularGrid& grid, const std::vector<uvec4>& seeds, FractureParameters* fp) override
    {
        check(vf_fracture_flood(grid.handle(), reinterpret_cast<const uint32_t*>(seeds.data()), (uint32_t)seeds.size(), (int)_dfunc, fp ? fp->_floodIdBits : 0, &lastStats));
    }
    bool setDistanceFunction(DistanceFunction dfunc) override
    {
        _dfunc = dfunc;
        return true;
    }
    vf_flood_stats lastStats{};

private:
    DistanceFunction _dfunc = MANHATTAN_DISTANCE;
};

class Seeder {  // Seeder.h:54-76 (all static)
public:
    static const uint32_t VOXEL_ID_POSITION = 8;
    class SeederSearchError : public std::runtime_error {
    public:
        explicit SeederSearchError(const std::string& msg) : std::runtime_error(msg) {}
    };
    enum Location { INNER, OUTER, BOTH };

    static std::vector<uvec4> uniform(RegularGrid& grid, unsigned int nseeds, int randomSeedFunction, Location location = OUTER)
    {
        std::vector<uvec4> out(nseeds);
        const vf_status s = vf_seed_uniform(grid.handle(), nseeds, randomSeedFunction, (int)location, reinterpret_cast<uint32_t*>(out.data()), nullptr);
        if (s == VF_ERR_SEEDER_EXHAUSTED) throw SeederSearchError(vf_last_error());
        check(s);
        return out;
    }
    // nearSeeds(grid, frags, numImpacts, numSeeds, spreading), Seeder.h:62-64
    static std::vector<uvec4> nearSeeds(const RegularGrid& grid, const std::vector<uvec4>& frags, unsigned numImpacts, unsigned numSeeds, unsigned spreading)
    {
        std::vector<uvec4> out(frags.size() + numSeeds);
        uint32_t n = 0;
        const vf_status s = vf_seed_near(grid.handle(), reinterpret_cast<const uint32_t*>(frags.data()), (uint32_t)frags.size(), numImpacts, numSeeds, spreading,
                                         reinterpret_cast<uint32_t*>(out.data()), (uint32_t)out.size(), &n);
        if (s == VF_ERR_SEEDER_EXHAUSTED) throw SeederSearchError(vf_last_error());
        check(s);
        out.resize(n);
        return out;
    }
    static void mergeSeeds(const std::vector<uvec4>& frags, std::vector<uvec4>& seeds, DistanceFunction dfunc)
    {
        check(vf_merge_seeds(reinterpret_cast<const uint32_t*>(frags.data()), (uint32_t)frags.size(), reinterpret_cast<uint32_t*>(seeds.data()), (uint32_t)seeds.size(), (int)dfunc));
    }
};

}  // namespace fracturer

// CADScene::fractureModel (CADScene.cpp:624-691) as one call; returns "" or the reference's error strings.
inline std::string fractureModel(RegularGrid& grid, FractureParameters& fractParameters, std::vector<uvec4>* seedsOut = nullptr)
{
    const vf_params p = fractParameters.to_c();
    std::vector<uvec4> seeds((size_t)p.numSeeds * 2 + p.numExtraSeeds);
    uint32_t n = 0;
    vf_flood_stats st;
    const vf_status s = vf_fracture_model(grid.handle(), &p, reinterpret_cast<uint32_t*>(seeds.data()), &n, &st);
    if (s == VF_ERR_INVALID_DISTANCE) return "Invalid distance function";  // CADScene.cpp:665
    if (s == VF_ERR_SEEDER_EXHAUSTED) throw fracturer::Seeder::SeederSearchError(vf_last_error());
    check(s);
    seeds.resize(n);
    if (seedsOut) *seedsOut = seeds;
    return "";
}

// SRC/Graphics/Core/FragmentationProcedure.h:6-60 — the voxel-path fields, same names and defaults (the constructor's overrides of
// FractureParameters included).
struct FragmentationProcedure {
    FractureParameters _fractureParameters;
    ivec2 _fragmentInterval, _iterationInterval;
    size_t _maxFragmentsModel;
    std::string _startVessel = "", _searchExtension = ".obj";
    bool _exportGrid;
    bool _exportMesh = false;         // fragment meshes as .binm + mesh metadata rows (the reference's _exportMesh with _targetTriangles empty)
    bool _solidVoxelization = false;  // extension: Tetravoxelizer occupancy instead of the SAT surface occupancy
    int _writerThreads = 2;           // extension: file writers beside the GPU loop

    FragmentationProcedure()
    {
        vf_procedure p;
        vf_procedure_default(&p);
        _fractureParameters.from_c(p.fractureParameters);
        _fragmentInterval = { p.fragmentInterval[0], p.fragmentInterval[1] };
        _iterationInterval = { p.iterationInterval[0], p.iterationInterval[1] };
        _maxFragmentsModel = p.maxFragmentsModel, _exportGrid = p.exportGrid != 0;
    }
    vf_procedure to_c() const
    {
        vf_procedure p;
        vf_procedure_default(&p);
        p.fractureParameters = _fractureParameters.to_c();
        p.fragmentInterval[0] = _fragmentInterval.x, p.fragmentInterval[1] = _fragmentInterval.y;
        p.iterationInterval[0] = _iterationInterval.x, p.iterationInterval[1] = _iterationInterval.y;
        p.maxFragmentsModel = _maxFragmentsModel, p.exportGrid = _exportGrid, p.solidVoxelization = _solidVoxelization, p.writerThreads = _writerThreads;
        p.exportMesh = _exportMesh;
        return p;
    }
};

// CADScene::generateDataset(fractureProcedure, folder, extension, destinationFolder) (CADScene.cpp:209-507), voxel path
inline vf_dataset_stats generateDataset(Context& ctx, FragmentationProcedure& fractureProcedure, const std::string& folder, const std::string& extension,
                                        const std::string& destinationFolder)
{
    vf_dataset_stats st{};
    const vf_procedure p = fractureProcedure.to_c();
    const vf_status s = vf_dataset_generate(ctx.handle(), &p, folder.c_str(), extension.c_str(), fractureProcedure._startVessel.c_str(), destinationFolder.c_str(), &st);
    if (s == VF_ERR_SEEDER_EXHAUSTED) throw fracturer::Seeder::SeederSearchError(vf_last_error());
    if (s == VF_ERR_IO && std::string(vf_last_error()).rfind("No files found", 0) == 0) throw std::runtime_error(vf_last_error());  // :213-214
    check(s);
    return st;
}

}  // namespace voxfrag
