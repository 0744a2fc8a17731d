/*
 * voxfrag.h — C ABI of libvoxfrag.so: the B200-native (sm_100a) implementation of VoxelFragmentML's
 * hot path  mesh -> uint16 voxel grid -> seeded fragmentation -> small-fragment cleanup.
 *
 * This is the drop-in boundary.  Every entry point names the reference interface it
 * replaces.  Shorthand: SRC/ = MeshFragments/Source/ of AlfonsoLRz/VoxelFragmentML.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary; no exception crosses it.
 *   - every call returns a vf_status (0 = ok); vf_last_error() gives a thread-local message.
 *   - a vf_ctx owns one CUDA device + one stream + scratch memory.  Calls on one context are issued on its
 *     stream and must be serialised by the caller (mirrors the reference's single-GL-thread rule,
 *     SRC/Graphics/Core/ComputeShader.cpp); distinct contexts may be driven from distinct host threads.
 *   - grids are device resident; host copies happen only in vf_grid_upload / vf_grid_download / vf_export
 *     (the reference's updateSSBO / updateGrid, SRC/DataStructures/RegularGrid.cpp:505-514).
 *   - label word layout is the reference's (SRC/DataStructures/RegularGrid.h:12-27): uint16 per voxel,
 *     0 = EMPTY, 1 = FREE, >= 2 fragment id, bit 15 = boundary mask, bits 8..14 = sub-seed prefix while flooding.
 *     Linear index = x*Y*Z + y*Z + z (z fastest; RegularGrid.cpp:839-842).
 *   - seeds are uint32[n][4] = {x, y, z, label} exactly like std::vector<glm::uvec4> (SRC/Fracturer/Fracturer.h:35).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns VF_ERR_CUDA.
 */
#ifndef VOXFRAG_H
#define VOXFRAG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VF_VOXEL_EMPTY 0u      /* RegularGrid.h:12 */
#define VF_VOXEL_FREE 1u       /* RegularGrid.h:13 */
#define VF_MASK_POSITION 15u   /* RegularGrid.h:18 */
#define VF_ID_POSITION 8u      /* SRC/Fracturer/Seeder.h:12 */
#define VF_HISTOGRAM_BINS 32768u

typedef enum vf_status {
    VF_OK = 0,
    VF_ERR_INVALID_ARGUMENT = 1,
    VF_ERR_SEEDER_EXHAUSTED = 2,   /* Seeder::SeederSearchError, SRC/Fracturer/Seeder.cpp:173-174 */
    VF_ERR_INVALID_DISTANCE = 3,   /* "Invalid distance function", SRC/Graphics/Application/CADScene.cpp:665 */
    VF_ERR_CAPACITY = 4,           /* label / distance / seed-count capacity exceeded (FloodFracturer::StackOverflowError's role) */
    VF_ERR_CUDA = 5,
    VF_ERR_IO = 6,
    VF_ERR_UNSUPPORTED = 7
} vf_status;

/* SRC/Fracturer/Fracturer.h:10-14 == FractureParameters::DistanceFunction (FractureParameters.h:20) */
typedef enum { VF_EUCLIDEAN = 0, VF_MANHATTAN = 1, VF_CHEBYSHEV = 2 } vf_distance;
/* FractureParameters.h:17 */
typedef enum { VF_NAIVE = 0, VF_FLOOD = 1, VF_VORONOI = 2 } vf_algorithm;
/* FractureParameters.h:23 */
typedef enum { VF_STD_UNIFORM = 0, VF_HALTON = 1, VF_BOOST_NORMAL_DISTRIBUTION = 2 } vf_random_type;
/* FractureParameters.h:26 */
typedef enum { VF_SQUARE = 0, VF_ELLIPSE = 1, VF_CROSS = 2 } vf_erosion_type;
/* FractureParameters.h:29 */
typedef enum { VF_VON_NEUMANN = 0, VF_MOORE = 1 } vf_neighbourhood;
/* FractureParameters.h:35 */
typedef enum { VF_RLE = 0, VF_QUADSTACK = 1, VF_VOX = 2, VF_UNCOMPRESSED_BINARY = 3 } vf_export_grid;
/* SRC/Fracturer/Seeder.h:24 */
typedef enum { VF_INNER = 0, VF_OUTER = 1, VF_BOTH = 2 } vf_seed_location;

/* The unchanged fragmentation-parameter surface: same field names (minus the leading underscore), same
 * enumerator values and same defaults (vf_params_default) as struct FractureParameters
 * (SRC/Graphics/Core/FractureParameters.h:43-145).  Rendering / mesh-export fields are not on this path. */
typedef struct vf_params {
    int32_t biasFocus;                   /* :43  (nearSeeds only)            default 5   */
    int32_t biasSeeds;                   /* :44                              default 32  */
    int32_t clampVoxelMetricUnit;        /* :46                              default 200 */
    int32_t erode;                       /* :47  bool                        default 0   */
    int32_t erosionConvolution;          /* :48  vf_erosion_type             default ELLIPSE */
    int32_t erosionIterations;           /* :49                              default 3   */
    float   erosionProbability;          /* :50                              default .5  */
    int32_t erosionSize;                 /* :51                              default 3   */
    float   erosionThreshold;            /* :52                              default .5  */
    int32_t fractureAlgorithm;           /* :53  vf_algorithm                default FLOOD */
    int32_t distanceFunction;            /* :54  vf_distance                 default CHEBYSHEV */
    int32_t launchGPU;                   /* :55  kept for layout parity; the CUDA path always runs */
    int32_t mergeSeedsDistanceFunction;  /* :57                              default EUCLIDEAN */
    int32_t neighbourhoodType;           /* :59                              default VON_NEUMANN (unused by FloodFracturer::build, which keys on the distance function, FloodFracturer.cpp:114) */
    int32_t numExtraSeeds;               /* :61                              default 16  */
    int32_t numImpacts;                  /* :62                              default 0   */
    int32_t numSeeds;                    /* :63                              default 8   */
    int32_t removeIsolatedRegions;       /* :65  bool                        default 1   */
    int32_t seed;                        /* :66                              default 80  */
    int32_t seedingRandom;               /* :67  vf_random_type              default STD_UNIFORM */
    int32_t voxelPerMetricUnit;          /* :70                              default 20  */
    int32_t voxelizationSize[3];         /* :71                              default 128^3 */
    int32_t exportGridExtension;         /* :79  vf_export_grid              default VOX */
    /* -- extensions (not in the reference; 0 = reference behaviour) -- */
    int32_t floodIdBits;                 /* 0/8: fragId|prefix<<8 words + disjoint rounds; 15: ids up to 32767, no prefixes (SURVEY finding 7) */
    int32_t erodeBoundaryMode;           /* 0: erodeGrid-comp.glsl:31 as written; 1: test bit 15 (SURVEY finding 8) */
} vf_params;

typedef struct vf_flood_stats {
    uint32_t tile_rounds;    /* global relaxation rounds (tile worklist generations), summed over flood phases */
    uint32_t tile_visits;    /* tiles processed, summed */
    uint32_t disjoint_rounds;/* outer `while (numDisjointVoxels != 0)` iterations, FloodFracturer.cpp:135 */
    uint32_t freed_voxels;   /* voxels returned to FREE by the disjoint step, summed */
    uint32_t max_dist;       /* largest geodesic distance reached in phase 1 */
    uint32_t front_levels;   /* relaxation steps (~ BFS levels) the thin-front solver ran before the phase converged or the tiles took over, summed over phases */
} vf_flood_stats;

typedef struct vf_ctx vf_ctx;
typedef struct vf_grid vf_grid;

/* ------------------------------------------------------------------ library / context */
const char* vf_last_error(void);
const char* vf_version(void);
int vf_device_count(void);
void vf_params_default(vf_params* p);                                   /* FractureParameters::FractureParameters(), FractureParameters.h:91-145 */

vf_status vf_ctx_create(int device, vf_ctx** out);                       /* replaces the GL context + ShaderList singletons */
vf_status vf_ctx_create_on_stream(int device, void* cuda_stream, vf_ctx** out); /* borrow an existing cudaStream_t (e.g. torch's current stream) */
void      vf_ctx_destroy(vf_ctx* ctx);
vf_status vf_ctx_reserve(vf_ctx* ctx, uint32_t X, uint32_t Y, uint32_t Z); /* Fracturer::prepareSSBOs / init, Fracturer.h:43-48; FloodFracturer.cpp:47-59 */
vf_status vf_ctx_synchronize(vf_ctx* ctx);
/* on = 1: host waits of this context sleep on a blocking event instead of spinning — for producers that drive more contexts than
 * they have host cores (batch generation overlaps several jobs per GPU); on = 2: they poll and yield the core between polls (sched_yield);
 * 0 (default): spin, lowest latency */
vf_status vf_ctx_set_blocking_sync(vf_ctx* ctx, int on);
/* Width, in distance levels, of the window a flood round may assign (the result does not depend on it; 0 = default 16, the best
 * latency for one job).  A narrower window orders the fronts better at the price of more rounds: 8 gives the highest throughput
 * when several jobs share the GPU (measured: 182 -> 200 models/s in batch generation). */
vf_status vf_ctx_set_flood_levels(vf_ctx* ctx, uint32_t levels);
/* Every flood phase starts at cell granularity on one thread-block cluster (the front as lists of (cell, key) pairs in shared memory; right for
 * voxelized surfaces, whose BFS levels hold a few thousand cells: 0.93 -> 0.70-0.80 ms at 176x256x176 with 16 seeds, 1.86 -> 1.21 ms with extra
 * seeds) and moves to the tile worklist when more than max_front_cells pairs are pending (solid interiors).  Default 8192; 0 = tiles only; at most
 * 65536.  Same labels.  It is the latency path of ONE flood: producers that drive many contexts per GPU (batch generation) should pass 0 — a
 * 16-CTA cluster per phase and job is the wrong grain when jobs share the SMs (bench.py does). */
vf_status vf_ctx_set_flood_front(vf_ctx* ctx, uint32_t max_front_cells);
/* How the tile rounds of a flood phase are driven; the labels do not depend on it.  1..4 (default 4): ONE cooperative launch per phase with that many CTAs per SM,
 * the round loop on the device (no host read-back until the phase has converged): lowest latency for one job, and with 1 or 2 several jobs
 * fit on the GPU side by side.  0: one launch per round, read-backs every few rounds (rounds of many jobs interleave freely). */
vf_status vf_ctx_set_flood_mode(vf_ctx* ctx, int ctas_per_sm);
/* How vf_remove_isolated_regions (C1) is computed; the result is the same.  0 (default): on grids of at least 2^26 cells one streaming "descent
 * certificate" pass plus list work on the cells it cannot certify (512^3: 0.19 ms against 0.60), falling back to the union-find when those lists
 * outgrow 65 536 cells, two seeds share a label or the rows are not 16-byte aligned; smaller grids go straight to the union-find (<= 0.1 ms there,
 * while the certificate's one-CTA list work can take milliseconds on a small thin shell).  1: the union-find only.  2: the certificate on a grid
 * of any size (tests, tools). */
vf_status vf_ctx_set_c1_mode(vf_ctx* ctx, int mode);
void*     vf_ctx_stream(vf_ctx* ctx);                                    /* the cudaStream_t every call of this context is issued on */
uint64_t  vf_ctx_kernel_launches(vf_ctx* ctx);                           /* kernels launched by this context so far (bench "gpu_launches") */
uint64_t  vf_ctx_host_waits(vf_ctx* ctx);                                /* times a call of this context made the host wait for the stream so far */
/* CUDA-event timing on the context's stream (ResourceTracker's role, SRC/Utilities/ResourceTracker.cpp:58-72) */
vf_status vf_ctx_timer_start(vf_ctx* ctx);
vf_status vf_ctx_timer_stop(vf_ctx* ctx, float* elapsed_ms);             /* synchronises on the stop event */

/* ------------------------------------------------------------------ RNG (process-global in the reference, per context here) */
vf_status vf_rng_seed(vf_ctx* ctx, uint32_t seed);                       /* RandomUtilities::initSeed, SRC/Utilities/RandomUtilities.h:86-89; CADScene.cpp:36-37 */
vf_status vf_crand_seed(vf_ctx* ctx, uint32_t seed);                     /* srand(), CADScene.cpp:36 (vf_rng_seed does both, as :36-37 do) */
int       vf_crand_next(vf_ctx* ctx);                                    /* rand() of the reference's C runtime (MSVC): state * 214013 + 2531011, bits 16..30 */
float     vf_rng_uniform(vf_ctx* ctx);                                   /* RandomUtilities::getUniformRandom, :103-106 (libstdc++ float recipe, SURVEY finding 9) */
uint32_t  vf_rng_raw(vf_ctx* ctx);
vf_status vf_fill_noise(vf_ctx* ctx, float* noise, uint32_t n);          /* RegularGrid::fillNoiseBuffer, RegularGrid.cpp:238-244 (serial draw order) */

/* ------------------------------------------------------------------ grid (class RegularGrid) */
vf_status vf_grid_create(vf_ctx* ctx, uint32_t X, uint32_t Y, uint32_t Z, vf_grid** out);  /* RegularGrid(ivec3), RegularGrid.cpp:28-32 */
vf_status vf_grid_wrap(vf_ctx* ctx, void* device_u16, uint32_t X, uint32_t Y, uint32_t Z, vf_grid** out); /* borrow caller-owned device memory (16-byte aligned) */
void      vf_grid_destroy(vf_grid* g);
vf_status vf_grid_set_aabb(vf_grid* g, const float aabb_min[3], const float aabb_max[3], uint32_t X, uint32_t Y, uint32_t Z); /* RegularGrid::setAABB + cleanGrid, :426-441,591-599 */
vf_status vf_grid_dims(const vf_grid* g, uint32_t dims[3]);              /* getNumSubdivisions, :528-531 */
void*     vf_grid_device_ptr(vf_grid* g);                                /* RegularGrid::ssbo() */
vf_status vf_grid_upload(vf_grid* g, const uint16_t* host);              /* updateSSBO, :511-514 */
vf_status vf_grid_download(vf_grid* g, uint16_t* host);                  /* updateGrid, :505-509 (synchronises) */
vf_status vf_grid_upload_async(vf_grid* g, const uint16_t* pinned_host);
/* Occupancy as ONE BIT per cell instead of a 16-bit word (new: 16x fewer bytes over PCIe for a grid that only holds EMPTY / FREE, which is
 * what every fragmentation starts from after RegularGrid::fill + homogenize): bit (i & 7) of byte i >> 3 is cell i of the linear x-major
 * order; set -> VOXEL_FREE, clear -> VOXEL_EMPTY.  Enqueued on the context's stream like vf_grid_upload_async (pinned memory: asynchronous). */
vf_status vf_grid_upload_bits(vf_grid* g, const uint8_t* host_bits);
vf_status vf_grid_download_async(vf_grid* g, uint16_t* pinned_host);
vf_status vf_grid_fill(vf_grid* g, uint16_t value);
/* interactive dims rule of CADScene::allocateMeshGrid, CADScene.cpp:545-556 (host arithmetic only) */
void      vf_dims_rule(const float aabb_min[3], const float aabb_max[3], uint32_t max_voxels, uint32_t dims_out[3]);

/* ------------------------------------------------------------------ V2: voxelization */
/* RegularGrid::fill(Model3D*) (RegularGrid.cpp:173-212) with the north-star occupancy predicate: voxel = FREE iff some
 * triangle passes Intersections3D::intersect(Triangle3D&, AABB&) (SRC/Geometry/3D/Intersections3D.h:204-420) against the
 * voxel box of RegularGrid.cpp:258-259.  verts/faces are HOST pointers: float[nv][3], uint32[nf][3]. */
vf_status vf_voxelize(vf_grid* g, const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf);
/* V1: RegularGrid::fill(Model3D*) as the reference runs it today: Tetravoxelizer (SRC/Graphics/Core/Tetravoxelizer.cpp:198-315,
 * geometry shader :42-92) — a cell is FREE iff an odd number of (face, centroid) tetrahedra cover (cell centre x, slice plane
 * y = -1 + slice * 2/Y accumulated in float32, cell centre z) of the grid AABB's NDC space: solid interiors, one model component
 * per call.  The reference's pixel coverage is decided by the GL rasteriser; the rule used here is stated in csrc/voxelize.cu.
 * The grid is cleared first (setAABB -> cleanGrid, RegularGrid.cpp:426-441).  *occupied_out (optional) = FREE cells written; the
 * reference falls back to random surface sampling (fillNaive, :800-816) when that is 0 — that fallback is not on this path. */
vf_status vf_voxelize_solid(vf_grid* g, const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, uint64_t* occupied_out);

/* ------------------------------------------------------------------ S1/S2: seeding (class fracturer::Seeder) */
/* Seeder::uniform (Seeder.cpp:154-208).  The grid stays on the device: candidate draws are made on the host in the
 * reference's order and tested on the device in batches; the RNG ends exactly where the reference's would. */
vf_status vf_seed_uniform(vf_grid* g, uint32_t n, int random_mode, int location, uint32_t* seeds_out, uint32_t* attempts_out);
/* Seeder::mergeSeeds (Seeder.cpp:115-152) — host only */
vf_status vf_merge_seeds(const uint32_t* frags, uint32_t nfrags, uint32_t* seeds, uint32_t nseeds, int dfunc);
/* Seeder::nearSeeds (Seeder.cpp:49-113): impact-biased boundary seeds around randomly chosen fragments; returns frags + the new seeds
 * (labels continue after the last fragment's).  `out` may alias `frags`.  The offsets come from C rand() — the MSVC LCG, per context. */
vf_status vf_seed_near(vf_grid* g, const uint32_t* frags, uint32_t nfrags, uint32_t num_impacts, uint32_t num_seeds, uint32_t spreading,
                       uint32_t* out, uint32_t capacity, uint32_t* count_out);
/* seed block of CADScene::fractureModel (CADScene.cpp:626-655, numImpacts == 0): returns n or n + n + n_extra seeds */
vf_status vf_make_seeds(vf_grid* g, uint32_t n, uint32_t n_extra, int random_mode, int merge_dfunc, uint32_t* seeds_out,
                        uint32_t capacity, uint32_t* count_out);

/* ------------------------------------------------------------------ F1..F3: fragmentation (class fracturer::Fracturer) */
/* NaiveFracturer::build (NaiveFracturer.cpp:215-225; spec = buildCPU :26-68 / naiveFracturer-comp.glsl:19-43) */
vf_status vf_fracture_naive(vf_grid* g, const uint32_t* seeds, uint32_t nseeds, int dfunc);
/* the same on ONE SLAB of a grid cut along x (multi-GPU; the reference is single-GPU): `slab` holds planes x_origin ..
 * x_origin + X(slab) - 1 of a grid of X_full planes, halo planes included (nearest-seed labels are pointwise: the halo is computed, not
 * exchanged); seeds in the coordinates of the whole grid.  Bit-identical to the same planes of vf_fracture_naive on the whole grid. */
vf_status vf_fracture_naive_slab(vf_grid* slab, const uint32_t* seeds, uint32_t nseeds, int dfunc, uint32_t x_origin, uint32_t X_full);
/* FloodFracturer::build (FloodFracturer.cpp:98-191) under the deterministic lowest-seed-index rule.
 * dfunc MANHATTAN -> 6-neighbourhood, otherwise 26 (FloodFracturer.cpp:114).  id_bits 0/8 or 15 (see vf_params). */
vf_status vf_fracture_flood(vf_grid* g, const uint32_t* seeds, uint32_t nseeds, int dfunc, int id_bits, vf_flood_stats* stats);

/* ------------------------------------------------------------------ F2 on one grid split into x-slabs over several GPUs (new: the reference is single-GPU) */
/* A slab is an ordinary vf_grid of (xs + 2) x Y x Z cells: xs owned planes plus one halo plane on each side that mirrors the
 * neighbour GPU's boundary plane (EMPTY where the global grid ends).  keys_dev is caller-owned device memory of the same shape
 * (uint32 per cell) so that the host layer can hand its boundary planes to NCCL.  seeds_local = uint32[n][4] {slab-local x
 * (halo planes included), y, z, ORDER in the global seed list}.  Protocol: init; repeat { relax; send the two owned boundary
 * planes (vf_flood_slab_boundary_ptr) to the neighbours; ingest what they sent; all-reduce the change counts } until zero;
 * finalize writes labels (seeds_global[order].w) into the slab grid.  Only the 15-bit id layout is supported (no extra seeds).
 * A slab session uses its context's tile scratch: do not interleave other flood / cleanup calls on the same context. */
typedef struct vf_slab vf_slab;
vf_status vf_flood_slab_init(vf_grid* slab_grid, uint32_t* keys_dev, const uint32_t* seeds_local, uint32_t nseeds, int dfunc, int has_lo, int has_hi,
                             vf_slab** out);
vf_status vf_flood_slab_relax(vf_slab* s, uint64_t* changed_cells);
void*     vf_flood_slab_boundary_ptr(vf_slab* s, int side /* 0 = towards lower x, 1 = towards higher x */);
vf_status vf_flood_slab_ingest(vf_slab* s, int side, const uint32_t* plane_dev, uint64_t* changed_cells);
vf_status vf_flood_slab_finalize(vf_slab* s, const uint32_t* seeds_global, uint32_t nseeds_total, uint32_t* max_dist);
void      vf_flood_slab_destroy(vf_slab* s);
/* The whole exchange loop on the host side of this library, over NCCL: { relax; grouped ncclSend / ncclRecv of the owned boundary planes
 * with ranks rank-1 / rank+1 on the context's stream; ingest; ncclAllReduce of the change count } until no rank changed a cell — one
 * host wait per iteration, change counters and round id stay on the device.  NCCL is taken from the libnccl.so.2 already in the process
 * (no link-time dependency).  A host application creates the communicator itself (vf_nccl_unique_id on one rank, the 128 bytes carried to
 * the others by whatever it uses for bootstrap, vf_nccl_comm_create on each) or passes its own ncclComm_t.  world == 1: comm may be NULL. */
vf_status vf_nccl_unique_id(void* id128);
vf_status vf_nccl_comm_create(vf_ctx* ctx, const void* id128, int world, int rank, void** comm_out);
void      vf_nccl_comm_destroy(void* comm);
vf_status vf_flood_slab_run(vf_slab* s, void* nccl_comm, int rank, int world, uint32_t* iterations, uint64_t* halo_bytes);

/* ------------------------------------------------------------------ C1..C4: cleanup */
vf_status vf_remove_isolated_regions(vf_grid* g, const uint32_t* seeds, uint32_t nseeds); /* NaiveFracturer::removeIsolatedRegionsCPU semantics, NaiveFracturer.cpp:111-150 */
vf_status vf_detect_boundaries(vf_grid* g, int boundary_size);           /* RegularGrid::detectBoundaries, RegularGrid.cpp:64-80 */
/* RegularGrid::erode (RegularGrid.cpp:82-159); noise = HOST table (vf_fill_noise), boundary_mode see vf_params.
 * The table is copied on the context's stream: a pageable table may be reused on return, a pinned one must stay untouched until
 * the context is synchronised. */
vf_status vf_erode(vf_grid* g, int erosion_type, uint32_t size, uint32_t iterations, float probability, float threshold,
                   const float* noise, uint32_t nnoise, int boundary_mode);
/* ONE erosion pass of RegularGrid::erode's loop (erodeGrid-comp.glsl + copyGrid, RegularGrid.cpp:137-152) without detectBoundaries before
 * and without the sweep after it — for a slab of a grid cut along x: vf_detect_boundaries, halo exchange, { vf_erode_pass, halo exchange }
 * per iteration, vf_remove_isolated_regions_grid (voxelfragmentml_b200/slab.py: LabelSlab).  cell_offset = index of the slab's first cell
 * (halo included) in the whole grid: the noise is indexed by the cell's position there. */
vf_status vf_erode_pass(vf_grid* g, int erosion_type, uint32_t size, float probability, float threshold, const float* noise, uint32_t nnoise,
                        int boundary_mode, uint64_t cell_offset);
vf_status vf_remove_isolated_regions_grid(vf_grid* g);                   /* RegularGrid::removeIsolatedRegions, RegularGrid.cpp:1006-1015 (snapshot semantics) */
vf_status vf_undo_mask(vf_grid* g);                                      /* RegularGrid::undoMask, RegularGrid.cpp:488-503 */
vf_status vf_reset_filling(vf_grid* g);                                  /* RegularGrid::resetFilling, :412-418 */
vf_status vf_homogenize(vf_grid* g);                                     /* RegularGrid::homogenize, :533-541 */

/* ------------------------------------------------------------------ H1: histogram */
/* RegularGrid::countValues + numOccupiedVoxels (RegularGrid.cpp:601-625, 280-287): counts[VF_HISTOGRAM_BINS] indexed by
 * (value & 0x7FFF) over cells with value > FREE; *occupied = number of such cells. */
vf_status vf_histogram(vf_grid* g, uint32_t* counts, uint64_t* occupied);
/* The grid side of CADScene::prepareScene (CADScene.cpp:813-832) in one pass over the grid: toTriangleMesh's countValues
 * (RegularGrid.cpp:443-471) followed by undoMask (:832) — same counts, bit 15 cleared afterwards. */
vf_status vf_histogram_undo_mask(vf_grid* g, uint32_t* counts, uint64_t* occupied);

/* ------------------------------------------------------------------ X1: export */
vf_status vf_export(vf_grid* g, const char* path_without_extension, int export_type, int squared); /* RegularGrid::exportGrid, :161-171 */
/* exportRLE (:672-714) with the runs found on the device: *bytes_out = 12 + 6 * runs; the byte stream is written to the HOST
 * buffer `out` when out != NULL && cap >= *bytes_out (otherwise the call is a size query).  Only the stream crosses PCIe. */
vf_status vf_grid_encode_rle(vf_grid* g, uint8_t* out, uint64_t cap, uint64_t* bytes_out);
/* in-memory encoders (host): return bytes needed; write when out != NULL && cap is large enough */
uint64_t  vf_encode_rle(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap);          /* exportRLE :672-714 */
uint64_t  vf_encode_bing_squared(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap); /* exportRawCompressed squared :638-666 */
uint64_t  vf_encode_vox(const uint16_t* grid, const uint32_t dims[3], int squared, uint8_t* out, uint64_t cap); /* exportVox :740-798 + VoxWriter.cpp:449-540 */
uint64_t  vf_encode_qstack(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap);       /* exportQuadStack :716-725 + SRC/DataStructures/QuadStack.h:91-226, GStack.h:278-319 */

/* ------------------------------------------------------------------ benchmark input synthesis (no reference counterpart) */
/* analytic solid vessel of revolution sampled at cell centres of an n^3 grid; the grid holds planes x_offset .. x_offset + X - 1 */
vf_status vf_synth_solid_vessel(vf_grid* g, int x_offset, uint32_t n, float base, float a1, float a2);

/* ------------------------------------------------------------------ the caller's block: CADScene::fractureModel */
/* CADScene.cpp:624-691: seeds -> Fracturer::build -> erode | detectBoundaries(1).  seeds_out (optional, capacity
 * (numSeeds [+ biasSeeds when numImpacts > 0]) * 2 + numExtraSeeds entries of uint32[4]) receives the seed list used. */
vf_status vf_fracture_model(vf_grid* g, const vf_params* p, uint32_t* seeds_out, uint32_t* nseeds_out, vf_flood_stats* stats);

/* ------------------------------------------------------------------ f2: per-fragment marching cubes (RegularGrid::toTriangleMesh's mesh side) */
/* FractureParameters' marching-cubes fields (FractureParameters.h:45,56,60; defaults :93-94,105,109-110) */
typedef struct vf_mc_params {
    float   boundaryMCIterations;     /* 0.048: iterations = unsigned(max grid dim * this), MarchingCubes.cpp:399-400 */
    float   boundaryMCWeight;         /* 0.2 */
    float   nonBoundaryMCIterations;  /* 0.048 */
    float   nonBoundaryMCWeight;      /* 0.9 */
    int32_t marchingCubesSubdivisions;/* 1; carried only: the reference never reads it (RegularGrid.cpp:423 passes a literal 1) */
} vf_mc_params;
typedef struct vf_mesh vf_mesh;       /* device-resident result: vertices float[nv][4] = xyz + boundary flag, faces uint32[nf][4] = 3 vertex numbers + boundary flag */
void      vf_mc_params_default(vf_mc_params* p);
/* MarchingCubes::setGrid + triangulateFieldGPU (SRC/Graphics/Core/MarchingCubes.cpp:523-540, 364-432) for one fragment label: marching cubes
 * on the padded grid (isolevel 0.5 on label == target_value, bit 15 ignored), Morton-sorted vertex fusion, model matrix
 * (RegularGrid.cpp:478-480), boundary marking, two-pass Laplacian smoothing.  Call between fractureModel and undoMask, as
 * prepareScene does (CADScene.cpp:813-832).  The reference's vertex / face ORDER is a race; the order here is deterministic (csrc/mesh.cu). */
vf_status vf_marching_cubes(vf_grid* g, uint32_t target_value, const vf_mc_params* params /* NULL = defaults */, vf_mesh** out);
vf_status vf_mesh_counts(const vf_mesh* m, uint32_t* num_vertices, uint32_t* num_faces);
vf_status vf_mesh_download(vf_mesh* m, float* vertices /* [nv][4] */, uint32_t* faces /* [nf][4] */);
void      vf_mesh_destroy(vf_mesh* m);

/* ------------------------------------------------------------------ f1: the dataset driver (CADScene::generateDataset, CADScene.cpp:209-507) */
/* struct FragmentationProcedure (SRC/Graphics/Core/FragmentationProcedure.h:6-60), voxel-path fields only */
typedef struct vf_procedure {
    vf_params fractureParameters;   /* :10, with the constructor's overrides (:41-60): biasSeeds 0, erode 0, voxelPerMetricUnit = clamp, RLE grids */
    int32_t   fragmentInterval[2];  /* :12  (2, 10) */
    int32_t   iterationInterval[2]; /* :13  (25, 15) */
    uint64_t  maxFragmentsModel;    /* :16  1000 */
    int32_t   exportGrid;           /* FractureParameters::_exportGrid, :52: 1 */
    /* -- extensions (0 = reference behaviour / north-star occupancy) -- */
    int32_t   solidVoxelization;    /* 0: SAT surface occupancy (vf_voxelize); 1: Tetravoxelizer occupancy (vf_voxelize_solid) */
    int32_t   exportMesh;           /* 1: per fragment marching cubes -> <itFile>_<idx>.binm (CADModel::saveBinary) + mesh metadata rows, i.e. the reference's
                                       _exportMesh with an empty _targetTriangles list (CADScene.cpp:344-420; simplification is not on this path); default 0 */
    int32_t   writerThreads;        /* file writers running beside the GPU loop (0 = write synchronously like the reference); default 2 */
} vf_procedure;

typedef struct vf_dataset_stats {
    uint64_t models, fragmentations, fragments, files, bytes_written, bytes_downloaded, voxels;
    double   seconds_voxelize, seconds_fracture, seconds_export;   /* host wall clock per phase (ResourceTracker's events, CADScene.cpp:236-238) */
} vf_dataset_stats;

void      vf_procedure_default(vf_procedure* p);                               /* FragmentationProcedure::FragmentationProcedure() */
void      vf_dataset_dims_rule(const float aabb_min[3], const float aabb_max[3], int32_t voxelPerMetricUnit, int32_t clampVoxelMetricUnit,
                               uint32_t dims_out[3]);                        /* CADScene.cpp:262-273 */
/* glm::mix of the iteration interval, CADScene.cpp:304-306.  A degenerate fragment interval (x == y) makes the reference divide by zero;
 * this function then returns iterationInterval.x (documented divergence: the reference's result is undefined there). */
int32_t   vf_dataset_iterations(const vf_procedure* p, int32_t numFragments);
/* the body of generateDataset's model loop (:240-466) for one already-loaded model: dims rule, setAABB + fill, starting-grid export,
 * every (numFragments, iteration) fragmentation with its grid export and metadata rows, exportMetadata.  `g` must have capacity for
 * the dims rule's result (allocate (clamp+3) x clamp x (clamp+3)).  The context's RNG stream is continued, not re-seeded. */
vf_status vf_dataset_model(vf_grid* g, const vf_procedure* proc, const char* model_name, const float* verts, uint32_t nv, const uint32_t* faces,
                           uint32_t nf, const char* destination_folder, vf_dataset_stats* stats);
/* generateDataset itself: searchFiles(folder, extension) + _startVessel skip + one grid for the whole run.  Models are read with a
 * minimal Wavefront .obj reader (Assimp is not on this path) and normalised as CADModel::load does (CADModel.cpp:148-152). */
vf_status vf_dataset_generate(vf_ctx* ctx, const vf_procedure* proc, const char* folder, const char* extension, const char* start_vessel,
                              const char* destination_folder, vf_dataset_stats* stats);
vf_status vf_load_obj(const char* path, float** verts_out, uint32_t* nv_out, uint32_t** faces_out, uint32_t* nf_out); /* free with vf_free_host */
void      vf_free_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* VOXFRAG_H */
