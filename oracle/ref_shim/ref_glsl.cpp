// ref_glsl.cpp — TEST INFRASTRUCTURE (oracle/): the reference's OWN compute shaders, compiled as C++ from the text where it lies under
// /root/reference/MeshFragments/Assets/Shaders/Compute/Fracturer (glsl2cpp.py writes the include files into the git-ignored oracle/_ref/glsl/
// at build time), run on the host cores, one loop iteration per shader invocation.  It is the second pin of the oracle for the stages that
// exist only as GLSL in the reference (detectBoundaries, erodeGrid, copyGrid, removeIsolatedRegionsGrid, undoMask, naiveFracturer,
// floodFracturer, disjointSet, disjointSetStack) and the "reference shaders on the box's host cores" arm of bench.py --impl reference.
//
// What is restated here is the HOST side that binds buffers, sets uniforms and dispatches — the reference does that through OpenGL
// (ComputeShader / ShaderProgram), which this image does not have.  Every driver below cites the host loop it follows.
//
// Invocation order.  A GPU runs the invocations of a dispatch in an unspecified order; shaders whose result does not depend on it
// (detectBoundaries: neighbours are read with bit 15 cleared; erodeGrid / copyGrid / undoMask / naiveFracturer / disjointSet: every
// invocation writes only its own cell or uses atomics that commute) run under OpenMP.  The racy ones run under a stated schedule:
//   floodFracturer, disjointSetStack   ascending invocation index, one at a time (one of the schedules a GPU may produce);
//   removeIsolatedRegionsGrid          schedule 0 = ascending in place; schedule 1 = "every invocation reads the grid as it was before the
//                                      dispatch" (all reads before all writes), which is the snapshot rule DESIGN.md adopts; schedule 2 =
//                                      concurrently in place under OpenMP, as a GPU runs it (bench timing only).
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "glsl_rt.h"

thread_local uvec3 gl_GlobalInvocationID;

#define NCELLS ((size_t)dims[0] * dims[1] * dims[2])
#define GLSL_NAMES using std::abs; using std::max; using std::floor;

// clang-format off
namespace sh_detect   { GLSL_NAMES
#include "glsl/detectBoundaries-comp.inc"
}
namespace sh_erode    { GLSL_NAMES
#include "glsl/erodeGrid-comp.inc"
}
namespace sh_copy     { GLSL_NAMES
#include "glsl/copyGrid-comp.inc"
}
namespace sh_sweep    { GLSL_NAMES
#include "glsl/removeIsolatedRegionsGrid-comp.inc"
}
namespace sh_unmask   { GLSL_NAMES
#include "glsl/undoMask-comp.inc"
}
namespace sh_naive    { GLSL_NAMES
#include "glsl/naiveFracturer-comp.inc"
}
namespace sh_flood    { GLSL_NAMES
#include "glsl/floodFracturer-comp.inc"
}
namespace sh_djset    { GLSL_NAMES
#include "glsl/disjointSet-comp.inc"
}
namespace sh_djstack  { GLSL_NAMES
#include "glsl/disjointSetStack-comp.inc"
}
namespace sh_march    { GLSL_NAMES
#include "glsl/marchingCubes-comp.inc"
}
namespace sh_morton   { GLSL_NAMES
#include "glsl/computeMortonCodes-comp.inc"
}
namespace sh_fuse1    { GLSL_NAMES
#include "glsl/findSameVertices_01-comp.inc"
}
namespace sh_fuse2    { GLSL_NAMES
#include "glsl/findSameVertices_02-comp.inc"
}
namespace sh_faces    { GLSL_NAMES
#include "glsl/buildMarchingCubesFaces-comp.inc"
}
namespace sh_markb    { GLSL_NAMES
#include "glsl/markBoundaryTriangles-comp.inc"
}
namespace sh_lapreset { GLSL_NAMES
#include "glsl/resetLaplacianBuffer-comp.inc"
}
namespace sh_lap      { GLSL_NAMES
#include "glsl/laplacianSmoothing-comp.inc"
}
namespace sh_lapfin   { GLSL_NAMES
#include "glsl/finishLaplacianSmoothing-comp.inc"
}
// clang-format on
#undef uniform
#undef in
#undef UINT_MAX  // constraints.glsl's own (28-bit) value, not <climits>'

namespace {

// MarchingCubes::_triangleTable / _edgeTable (MarchingCubes.cpp:6-298), cut out of the reference source at build time
const int mc_triangle_table[256 * 16] = {
#include "mc_triangle_table.inc"
};
const int mc_edge_table[256] = {
#include "mc_edge_table.inc"
};

// ComputeShader::execute(numGroups, ...) launches numGroups * groupSize invocations; the ones at or beyond numCells return at once, so only
// [0, n) is run.  `parallel` only for the order-independent shaders (see the header comment).
template <typename Main>
void dispatch(uint n, bool parallel, Main shader_main)
{
    if (parallel) {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) {
            gl_GlobalInvocationID.x = (uint)i;
            shader_main();
        }
    } else {
        for (uint i = 0; i < n; ++i) {
            gl_GlobalInvocationID.x = i;
            shader_main();
        }
    }
}

uvec3 dims3(const uint32_t* d) { return uvec3(d[0], d[1], d[2]); }

// RegularGrid::detectBoundaries (RegularGrid.cpp:64-80)
void detect_boundaries(uint16_t* grid, const uint32_t* dims, int boundarySize)
{
    using namespace sh_detect;
    sh_detect::grid.bind(reinterpret_cast<CellGrid*>(grid), NCELLS);
    sh_detect::boundarySize = boundarySize;
    sh_detect::gridDims = dims3(dims);
    sh_detect::numCells = dims[0] * dims[1] * dims[2];
    dispatch(sh_detect::numCells, true, sh_detect::shader_main);
}

// RegularGrid::removeIsolatedRegions (RegularGrid.cpp:1006-1015)
void sweep(uint16_t* grid, const uint32_t* dims, int schedule)
{
    const uint n = dims[0] * dims[1] * dims[2];
    sh_sweep::gridDims = dims3(dims);
    sh_sweep::numCells = n;
    if (schedule == 0 || schedule == 2) {  // 2: concurrently in place, as a GPU runs it (timing only: the outcome depends on the interleaving)
        sh_sweep::grid.bind(reinterpret_cast<sh_sweep::CellGrid*>(grid), NCELLS);
        dispatch(n, schedule == 2, sh_sweep::shader_main);
        return;
    }
    // all reads before all writes: every invocation runs on the grid as it was before the dispatch (its own store is collected and undone)
    std::vector<uint16_t> before(grid, grid + n);
    sh_sweep::grid.bind(reinterpret_cast<sh_sweep::CellGrid*>(before.data()), NCELLS);
    for (uint i = 0; i < n; ++i) {
        const uint16_t own = before[i];
        gl_GlobalInvocationID.x = i;
        sh_sweep::shader_main();
        grid[i] = before[i];
        before[i] = own;
    }
}

}  // namespace

extern "C" {

void glsl_detect_boundaries(uint16_t* grid, const uint32_t* dims, int boundarySize) { detect_boundaries(grid, dims, boundarySize); }

void glsl_remove_isolated_regions_grid(uint16_t* grid, const uint32_t* dims, int schedule) { sweep(grid, dims, schedule); }

// RegularGrid::undoMask (RegularGrid.cpp:488-503: position 15, unmaskBit) and the tail of FloodFracturer::build (FloodFracturer.cpp:180-186:
// position 8, unmaskRightMost)
void glsl_undo_mask(uint16_t* grid, const uint32_t* dims, uint32_t position, int rightmost)
{
    sh_unmask::grid.bind(reinterpret_cast<sh_unmask::CellGrid*>(grid), NCELLS);
    sh_unmask::numCells = dims[0] * dims[1] * dims[2];
    sh_unmask::position = position;
    sh_unmask::unmaskUniform = rightmost ? sh_unmask::unmaskRightMost : sh_unmask::unmaskBit;
    dispatch(sh_unmask::numCells, true, sh_unmask::shader_main);
}

// RegularGrid::erode (RegularGrid.cpp:82-159).  The noise table is the caller's (fillNoiseBuffer, :238-244, draws it from the global RNG).
// sweep_schedule: see removeIsolatedRegionsGrid above.
void glsl_erode(uint16_t* grid, const uint32_t* dims, int type, uint32_t convolutionSize, uint32_t numIterations, float erosionProbability, float erosionThreshold,
                const float* noise, uint32_t nnoise, int sweep_schedule)
{
    if (!(convolutionSize % 2)) ++convolutionSize;                                                          // :84-85
    const uint32_t maskSize = convolutionSize * convolutionSize * convolutionSize;                          // :87
    const uint32_t convolutionCenter = (uint32_t)std::floor(convolutionSize / 2.0f);
    float activations = 0;
    std::vector<float> erosionMask(maskSize, 0.0f);
    if (type == 0) {  // SQUARE :91-95
        std::fill(erosionMask.begin(), erosionMask.end(), 1.0f);
        activations = (float)maskSize;
    } else if (type == 2) {  // CROSS :96-105
        for (uint32_t x = 0; x < convolutionSize; ++x) erosionMask[x * convolutionSize * convolutionSize + convolutionCenter * convolutionSize + convolutionCenter] = 1.0f;
        for (uint32_t y = 0; y < convolutionSize; ++y) erosionMask[convolutionCenter * convolutionSize * convolutionSize + y * convolutionSize + convolutionCenter] = 1.0f;
        for (uint32_t z = 0; z < convolutionSize; ++z) erosionMask[convolutionCenter * convolutionSize * convolutionSize + convolutionCenter * convolutionSize + z] = 1.0f;
        activations = 1.0f / 3.0f * maskSize;
    } else if (type == 1) {  // ELLIPSE :106-119 (glm::distance on float vec3, glm::epsilon<float>() = FLT_EPSILON)
        for (uint32_t x = 0; x < convolutionSize; ++x)
            for (uint32_t y = 0; y < convolutionSize; ++y)
                for (uint32_t z = 0; z < convolutionSize; ++z)
                    if (distance(vec3((float)x, (float)y, (float)z), vec3((float)convolutionCenter, (float)convolutionCenter, (float)convolutionCenter)) <
                        convolutionCenter + 1.1920928955078125e-07f) {
                        erosionMask[x * convolutionSize * convolutionSize + y * convolutionSize + z] = 1.0f;
                        ++activations;
                    }
    }
    activations /= maskSize;  // :122

    const uint n = dims[0] * dims[1] * dims[2];
    std::vector<uint16_t> dest(n);  // _marchingCubes->getGridSSBO() plays the destination (:137)
    for (uint32_t idx = 0; idx < numIterations; ++idx) {  // :133-153
        detect_boundaries(grid, dims, 1);
        sh_erode::grid.bind(reinterpret_cast<sh_erode::CellGrid*>(grid), NCELLS);
        sh_erode::destGrid.bind(reinterpret_cast<sh_erode::CellGrid*>(dest.data()), n);
        sh_erode::convolution.bind(erosionMask.data(), erosionMask.size());
        sh_erode::noise.bind(const_cast<float*>(noise), nnoise);
        sh_erode::numActivationsFloat = activations;
        sh_erode::gridDims = dims3(dims);
        sh_erode::maskSize = convolutionSize;
        sh_erode::maskSize2 = (unsigned)std::floor(convolutionSize / 2.0f);
        sh_erode::numCells = n;
        sh_erode::noiseBufferSize = nnoise;
        sh_erode::erosionProbability = erosionProbability;
        sh_erode::erosionThreshold = erosionThreshold;
        dispatch(n, true, sh_erode::shader_main);
        sh_copy::grid.bind(reinterpret_cast<sh_copy::CellGrid*>(grid), NCELLS);
        sh_copy::destGrid.bind(reinterpret_cast<sh_copy::CellGrid*>(dest.data()), n);
        sh_copy::numCells = n;
        dispatch(n, true, sh_copy::shader_main);
    }
    sweep(grid, dims, sweep_schedule);  // :155
}

// NaiveFracturer::buildGPU without the removeIsolatedRegions branch (NaiveFracturer.cpp:71-109); seeds = uvec4 {x, y, z, label}
void glsl_naive(uint16_t* grid, const uint32_t* dims, const uint32_t* seeds, uint32_t nseeds, int dfunc)
{
    sh_naive::seed.bind(reinterpret_cast<uvec4*>(const_cast<uint32_t*>(seeds)), nseeds);
    sh_naive::grid.bind(reinterpret_cast<sh_naive::CellGrid*>(grid), NCELLS);
    sh_naive::gridDims = dims3(dims);
    sh_naive::numSeeds = nseeds;
    sh_naive::distanceUniform = dfunc == 0 ? sh_naive::euclideanDistance : dfunc == 1 ? sh_naive::manhattanDistance : sh_naive::chebyshevDistance;
    dispatch(dims[0] * dims[1] * dims[2], true, sh_naive::shader_main);
}

// FloodFracturer::build (FloodFracturer.cpp:98-191).  stats[0] = BFS dispatches, [1] = trips of the outer (disjoint) loop, [2] = cells freed
// by disjointSetStack over all trips.  Returns 0, or -1 when a stack outgrew `stack_capacity` entries (the reference's stacks hold
// _voxelizationSize cells, :53-55, and overflow silently on the GPU).
int glsl_flood(uint16_t* grid, const uint32_t* dims, const uint32_t* seeds, uint32_t nseeds, int dfunc, uint32_t stack_capacity, uint32_t* stats)
{
    const uint numCells = dims[0] * dims[1] * dims[2];
    for (uint i = 0; i < numCells; ++i)  // grid.homogenize() :99 (RegularGrid.cpp:533-541)
        if (grid[i] != 0) grid[i] = 1;
    for (uint32_t s = 0; s < nseeds; ++s)  // :102-103, in order: a later seed on the same cell overwrites
        grid[(seeds[4 * s] * dims[1] + seeds[4 * s + 1]) * dims[2] + seeds[4 * s + 2]] = (uint16_t)seeds[4 * s + 3];

    // FloodFracturer.cpp:8-27
    static const ivec4 VON_NEUMANN[6] = { { 1, 0, 0, 0 }, { -1, 0, 0, 0 }, { 0, 1, 0, 0 }, { 0, -1, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, -1, 0 } };
    static const ivec4 MOORE[26] = { { 1, 0, 0, 0 },   { -1, 0, 0, 0 },  { 0, 1, 0, 0 },   { 0, -1, 0, 0 },   { 0, 0, 1, 0 },    { 0, 0, -1, 0 },  { 1, 1, 0, 0 },
                                     { 1, -1, 0, 0 },  { -1, 1, 0, 0 },  { -1, -1, 0, 0 }, { 0, 1, 1, 0 },    { 0, 1, -1, 0 },   { 0, -1, 1, 0 },  { 0, -1, -1, 0 },
                                     { 1, 0, 1, 0 },   { -1, 0, 1, 0 },  { 1, 0, -1, 0 },  { -1, 0, -1, 0 },  { 1, 1, 1, 0 },    { 1, 1, -1, 0 },  { 1, -1, 1, 0 },
                                     { -1, 1, 1, 0 },  { 1, -1, -1, 0 }, { -1, 1, -1, 0 }, { -1, -1, 1, 0 },  { -1, -1, -1, 0 } };
    const uint numNeigh = dfunc == 1 ? 6u : 26u;  // :114
    std::vector<ivec4> neigh(dfunc == 1 ? VON_NEUMANN : MOORE, (dfunc == 1 ? VON_NEUMANN : MOORE) + numNeigh);

    std::vector<uint> stack1((size_t)stack_capacity + 64), stack2((size_t)stack_capacity + 64);
    uint stackSize = nseeds;
    if (nseeds > stack_capacity) return -1;
    for (uint32_t s = 0; s < nseeds; ++s) stack1[s] = (seeds[4 * s] * dims[1] + seeds[4 * s + 1]) * dims[2] + seeds[4 * s + 2];  // :127-132

    // _disjointSetZero (:48-49): 256 uints meant to be UINT_MAX.  As written, std::iota runs over a uint16_t* and leaves pairs of counting
    // 16-bit values in the first 128 words (0x0000FFFF, 0x00020001, ...) and the second half of the malloc'ed block untouched; every prefix the
    // flood produces (< 256) is below all of those, so atomicMin behaves as if they were UINT_MAX.  The untouched half is taken as UINT_MAX here.
    std::vector<uint> disjointZero(256, 0xFFFFFFFFu);
    {
        uint16_t* h = reinterpret_cast<uint16_t*>(disjointZero.data());
        uint16_t v = 0xFFFFu;
        for (int i = 0; i < 256; ++i) h[i] = v++;
    }
    std::vector<uint> disjointSet(256);
    uint stackCounter = 0, disjointVoxels = 0;
    uint iteration = 0, trips = 0, freed_total = 0;
    uint numDisjointVoxels = stackSize;
    while (numDisjointVoxels != 0) {  // :135
        sh_flood::gridDims = dims3(dims);
        sh_flood::numNeighbors = numNeigh;
        while (stackSize > 0) {  // :143-158
            stackCounter = 0;
            sh_flood::grid.bind(reinterpret_cast<sh_flood::CellGrid*>(grid), NCELLS);
            sh_flood::stack01.bind(stack1.data(), stack1.size());
            sh_flood::stack02.bind(stack2.data(), stack2.size());
            sh_flood::p_stackCounter = &stackCounter;
            sh_flood::neighborOffset.bind(neigh.data(), neigh.size());
            sh_flood::stackSize = stackSize;
            const uint n = stackSize * numNeigh;
            for (uint i = 0; i < n; ++i) {
                gl_GlobalInvocationID.x = i;
                sh_flood::shader_main();
                if (stackCounter > stack_capacity) return -1;
            }
            stackSize = stackCounter;
            std::swap(stack1, stack2);
            ++iteration;
        }
        // :161-176
        disjointSet = disjointZero;
        disjointVoxels = 0;
        stackCounter = 0;
        sh_djset::grid.bind(reinterpret_cast<sh_djset::CellGrid*>(grid), NCELLS);
        sh_djset::disjointSet.bind(disjointSet.data(), disjointSet.size());
        sh_djset::gridDims = dims3(dims);
        dispatch(numCells, true, sh_djset::shader_main);
        sh_djstack::grid.bind(reinterpret_cast<sh_djstack::CellGrid*>(grid), NCELLS);
        sh_djstack::disjointSet.bind(disjointSet.data(), disjointSet.size());
        sh_djstack::stack.bind(stack1.data(), stack1.size());
        sh_djstack::p_stackSize = &stackCounter;
        sh_djstack::p_disjointVoxels = &disjointVoxels;
        sh_djstack::gridDims = dims3(dims);
        if (numCells > stack_capacity) return -1;
        dispatch(numCells, false, sh_djstack::shader_main);
        stackSize = stackCounter;
        numDisjointVoxels = disjointVoxels;
        freed_total += disjointVoxels;
        ++trips;
    }
    glsl_undo_mask(grid, dims, 8u, 1);  // :180-186
    if (stats) stats[0] = iteration, stats[1] = trips, stats[2] = freed_total;
    return 0;
}

// MarchingCubes::triangulateFieldGPU's first two dispatches for _marchingCubesSubdivisions == 1 (MarchingCubes.cpp:364-388, 445-453):
// marchingCubes-comp.glsl over every cell of the PADDED grid (`grid` = MarchingCubes::setGrid's copy: dims + 2, a ring of VOXEL_FREE,
// :523-534 — restated by the caller), then computeMortonCodes-comp.glsl over the vertices it appended.  The shader hands out vertex slots
// with an atomic counter, so the order of the triangles is the schedule's (ascending invocation index here); callers compare as sets.
// Returns the number of vertices (3 per triangle); verts = vec4 per vertex (x, y, z in padded-grid cells, boundary flag), morton = its code.
uint32_t glsl_mc_soup(const uint16_t* grid, const uint32_t* dims, uint32_t target, float* verts, uint32_t* morton, uint32_t cap)
{
    const uint n = dims[0] * dims[1] * dims[2];
    std::vector<vec4> vertexData(cap + 16), support((size_t)n * 12);
    std::vector<int> tri(mc_triangle_table, mc_triangle_table + 256 * 16), cfg(mc_edge_table, mc_edge_table + 256);
    uint counter = 0;  // resetCounter(_numVerticesSSBO), :379
    sh_march::grid.bind(const_cast<uint16_t*>(grid), n);
    sh_march::vertexData.bind(vertexData.data(), cap);
    sh_march::p_numVertices = &counter;
    sh_march::triangleTable.bind(tri.data(), tri.size());
    sh_march::configurationTable.bind(cfg.data(), cfg.size());
    sh_march::vertexList.bind(support.data(), support.size());
    sh_march::gridDims = dims3(dims);
    sh_march::isolevel = 0.5f;
    sh_march::localSize = dims3(dims);
    sh_march::start = uvec3(0, 0, 0);
    sh_march::targetValue = (int)target;
    dispatch(n, false, sh_march::shader_main);
    const uint nv = std::min(counter, cap);  // :390 (the buffer's capacity)
    std::vector<uint> codes(nv);
    sh_morton::vertices.bind(vertexData.data(), nv);
    sh_morton::mortonCode.bind(codes.data(), nv);
    sh_morton::numPoints = nv;
    sh_morton::sceneMaxBoundary = vec3(dims3(dims));  // :450-451
    sh_morton::sceneMinBoundary = vec3(.0f);
    dispatch(nv, true, sh_morton::shader_main);
    for (uint i = 0; i < nv; ++i) {
        verts[4 * i] = vertexData[i].x, verts[4 * i + 1] = vertexData[i].y, verts[4 * i + 2] = vertexData[i].z, verts[4 * i + 3] = vertexData[i].w;
        morton[i] = codes[i];
    }
    return counter;
}

// The whole of MarchingCubes::triangulateFieldGPU for one fragment and _marchingCubesSubdivisions == 1 (MarchingCubes.cpp:364-407): march,
// Morton codes, the sort, findSameVertices_01 / _02, buildMarchingCubesFaces, markBoundaryTriangles, smoothSurface twice (:498-521).
// `grid` is the padded copy (see glsl_mc_soup), `dims` the UNPADDED extent (RegularGrid::_numDivs, which scales the model matrix,
// RegularGrid.cpp:478-480).  The sort: the reference runs a stable 30-pass binary radix sort of the codes (sortMortonCodes, :537-609, generic
// RadixSort shaders); its result — indices ordered by code, ties in buffer order — is produced here by std::stable_sort.  Every dispatch runs
// in ascending invocation order (the atomic counters of march and findSameVertices_01 then number triangles and fused vertices in that order).
// counts = { fused vertices, faces }; nothing is written when the capacities are too small.
int glsl_marching_cubes(const uint16_t* grid, const uint32_t* dims, uint32_t target, const float* amin, const float* amax, uint32_t nb_iters,
                        float nb_weight, uint32_t b_iters, float b_weight, float* verts, uint32_t cap_v, uint32_t* faces, uint32_t cap_f, uint32_t* counts)
{
    const uint32_t pd[3] = { dims[0] + 2, dims[1] + 2, dims[2] + 2 };
    const uint n = pd[0] * pd[1] * pd[2];
    const uint cap = 15u * n;  // five triangles per cell at most
    std::vector<vec4> soup(cap), support((size_t)n * 12);
    std::vector<int> tri(mc_triangle_table, mc_triangle_table + 256 * 16), cfg(mc_edge_table, mc_edge_table + 256);
    uint counter = 0;
    sh_march::grid.bind(const_cast<uint16_t*>(grid), n);
    sh_march::vertexData.bind(soup.data(), cap);
    sh_march::p_numVertices = &counter;
    sh_march::triangleTable.bind(tri.data(), tri.size());
    sh_march::configurationTable.bind(cfg.data(), cfg.size());
    sh_march::vertexList.bind(support.data(), support.size());
    sh_march::gridDims = dims3(pd);
    sh_march::isolevel = 0.5f;
    sh_march::localSize = dims3(pd);
    sh_march::start = uvec3(0, 0, 0);
    sh_march::targetValue = (int)target;
    dispatch(n, false, sh_march::shader_main);
    const uint nv = counter;
    counts[0] = counts[1] = 0;
    if (nv == 0) return 0;
    // calculateMortonCodes :445-453
    std::vector<uint> codes(nv);
    sh_morton::vertices.bind(soup.data(), nv);
    sh_morton::mortonCode.bind(codes.data(), nv);
    sh_morton::numPoints = nv;
    sh_morton::sceneMaxBoundary = vec3(dims3(pd));
    sh_morton::sceneMinBoundary = vec3(.0f);
    dispatch(nv, true, sh_morton::shader_main);
    // sortMortonCodes: indices by code, stable
    std::vector<uint> indices(nv), indices2(nv, 0u);
    for (uint i = 0; i < nv; ++i) indices[i] = i;
    std::stable_sort(indices.begin(), indices.end(), [&](uint a, uint b) { return codes[a] < codes[b]; });
    // fuseSimilarVertices :455-474 with the model matrix of RegularGrid.cpp:478-480: translate(-scale) * translate(aabb.min) * scale(scale)
    mat4 model = {};
    for (int q = 0; q < 3; ++q) {
        const float scale = (amax[q] - amin[q]) / (float)dims[q];
        model.m[5 * q] = scale;
        model.m[12 + q] = amin[q] + (-scale);
    }
    model.m[15] = 1.0f;
    std::vector<vec4> fusedv(nv);
    uint vertexCount = 0;
    sh_fuse1::indices.bind(indices.data(), nv);
    sh_fuse1::indices2.bind(indices2.data(), nv);
    sh_fuse1::points.bind(soup.data(), nv);
    sh_fuse1::vertexData.bind(fusedv.data(), nv);
    sh_fuse1::p_vertexCount = &vertexCount;
    sh_fuse1::defaultValue = 0xFFFFFFFFu;  // UINT_MAX of <climits> (:457), not the shaders' own 28-bit constant
    sh_fuse1::modelMatrix = model;
    sh_fuse1::numPoints = nv;
    dispatch(nv, false, sh_fuse1::shader_main);
    sh_fuse2::indices.bind(indices.data(), nv);
    sh_fuse2::indices2.bind(indices2.data(), nv);
    sh_fuse2::defaultValue = 0xFFFFFFFFu;
    sh_fuse2::numPoints = nv;
    dispatch(nv, false, sh_fuse2::shader_main);
    const uint nfused = vertexCount, nf = nv / 3;
    counts[0] = nfused, counts[1] = nf;
    if (!verts || !faces || cap_v < nfused || cap_f < nf) return 0;
    // buildMarchingCubesFaces :436-442, markBoundaryTriangles :484-490
    std::vector<uvec4> face(nf);
    sh_faces::indices.bind(indices.data(), nv);
    sh_faces::indices2.bind(indices2.data(), nv);
    sh_faces::faceData.bind(face.data(), nf);
    sh_faces::numPoints = nv;
    dispatch(nv, true, sh_faces::shader_main);
    sh_markb::points.bind(fusedv.data(), nfused);
    sh_markb::faceData.bind(face.data(), nf);
    sh_markb::numFaces = nf;
    dispatch(nf, true, sh_markb::shader_main);
    // smoothSurface :498-521
    std::vector<int> l1(nfused), l2(nfused), l3(nfused), l4(nfused);
    auto smooth = [&](uint iterations, float weight, bool boundary) {
        for (uint it = 0; it < iterations; ++it) {
            sh_lapreset::laplacian01.bind(l1.data(), nfused), sh_lapreset::laplacian02.bind(l2.data(), nfused);
            sh_lapreset::laplacian03.bind(l3.data(), nfused), sh_lapreset::laplacian04.bind(l4.data(), nfused);
            sh_lapreset::numVertices = nfused;
            dispatch(nfused, true, sh_lapreset::shader_main);
            sh_lap::laplacian01.bind(l1.data(), nfused), sh_lap::laplacian02.bind(l2.data(), nfused);
            sh_lap::laplacian03.bind(l3.data(), nfused), sh_lap::laplacian04.bind(l4.data(), nfused);
            sh_lap::vertices.bind(fusedv.data(), nfused);
            sh_lap::face.bind(face.data(), nf);
            sh_lap::checkValidity = boundary ? 0u : 1u;
            sh_lap::numFaces = nf;
            sh_lap::targetVertexType = boundary ? 1.0f : .0f;
            dispatch(nf, true, sh_lap::shader_main);  // integer atomic adds commute
            sh_lapfin::laplacian01.bind(l1.data(), nfused), sh_lapfin::laplacian02.bind(l2.data(), nfused);
            sh_lapfin::laplacian03.bind(l3.data(), nfused), sh_lapfin::laplacian04.bind(l4.data(), nfused);
            sh_lapfin::vertices.bind(fusedv.data(), nfused);
            sh_lapfin::numVertices = nfused;
            sh_lapfin::targetVertexType = boundary ? 1.0f : .0f;
            sh_lapfin::weight = weight;
            dispatch(nfused, true, sh_lapfin::shader_main);
        }
    };
    smooth(nb_iters, nb_weight, false);
    smooth(b_iters, b_weight, true);
    for (uint i = 0; i < nfused; ++i) verts[4 * i] = fusedv[i].x, verts[4 * i + 1] = fusedv[i].y, verts[4 * i + 2] = fusedv[i].z, verts[4 * i + 3] = fusedv[i].w;
    for (uint i = 0; i < nf; ++i) faces[4 * i] = face[i].x, faces[4 * i + 1] = face[i].y, faces[4 * i + 2] = face[i].z, faces[4 * i + 3] = face[i].w;
    return 0;
}

}  // extern "C"
