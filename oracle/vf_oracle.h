/*
 * vf_oracle.h — CPU ORACLE for the VoxelFragmentML hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a restatement, function by function, of the reference's algorithm for
 * mesh -> voxel grid -> seeded fragmentation -> cleanup.  It is the
 * checker the CUDA path is compared against; it is never linked into, imported by or
 * called from the product (voxelfragmentml_b200/).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pin status:
 *   - .rle layout + a real occupancy grid: pinned by the reference fixture
 *     docs/decompress/samples/AL_12B_grid_128r.{rle,npy} (tests/golden/).
 *   - RNG recipe: pinned by libstdc++'s std::mt19937 + uniform_real_distribution<float>
 *     (self-check in orc_selfcheck_rng) and the SURVEY finding-9 draws for seed 80.
 *   - naive / C1 / SAT / seeder: pinned against the reference's own sources compiled in place
 *     (oracle/_ref, see oracle/ref_shim/) when that build is available; otherwise restatement only.
 *   - flood tie-break, erosion noise order, in-place isolated-region sweep: the reference is racy /
 *     non-reproducible there (SURVEY findings 5, 8); the deterministic rule is defined HERE
 *     ("parity unpinned" for those three, by construction).
 *
 * Shorthand: SRC/ = /root/reference/MeshFragments/Source/, SH/ = /root/reference/MeshFragments/Assets/Shaders/Compute/
 */
#ifndef VF_ORACLE_H
#define VF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* label words — SRC/DataStructures/RegularGrid.h:12-27 */
#define ORC_VOXEL_EMPTY 0
#define ORC_VOXEL_FREE 1
#define ORC_MASK_POSITION 15 /* RegularGrid.h:18, SH/Fracturer/voxelMask.glsl:1 */
#define ORC_ID_POSITION 8    /* SRC/Fracturer/Seeder.h:12, voxelMask.glsl:2 */

/* SRC/Fracturer/Fracturer.h:10-14 */
enum { ORC_EUCLIDEAN = 0, ORC_MANHATTAN = 1, ORC_CHEBYSHEV = 2 };
/* SRC/Fracturer/Seeder.h:24 */
enum { ORC_INNER = 0, ORC_OUTER = 1, ORC_BOTH = 2 };
/* SRC/Graphics/Core/FractureParameters.h:23,26 */
enum { ORC_STD_UNIFORM = 0, ORC_HALTON = 1, ORC_BOOST_NORMAL = 2 };
enum { ORC_SQUARE = 0, ORC_ELLIPSE = 1, ORC_CROSS = 2 };

/* error codes */
enum { ORC_OK = 0, ORC_ERR_SEEDER_EXHAUSTED = -1, ORC_ERR_UNSUPPORTED = -2, ORC_ERR_CAPACITY = -3, ORC_ERR_IO = -4 };

/* ---- RNG: SRC/Utilities/RandomUtilities.h:11-12,86-111,141-144 + SURVEY finding 9 ---- */
typedef struct orc_rng orc_rng;
orc_rng* orc_rng_create(uint32_t seed);
void     orc_rng_destroy(orc_rng*);
void     orc_rng_seed(orc_rng*, uint32_t seed);
uint32_t orc_rng_raw(orc_rng*);                       /* one mt19937 draw */
float    orc_rng_uniform(orc_rng*);                   /* getUniformRandom()            :103-106 */
float    orc_rng_uniform_range(orc_rng*, float lo, float hi); /* getUniformRandom(min,max) :108-111 */
int      orc_rng_uniform_int(orc_rng*, int lo, int hi);       /* getUniformRandomInt       :141-144 */
/* returns number of mismatches between the hard-coded recipe and libstdc++'s uniform_real_distribution<float> */
int      orc_selfcheck_rng(uint32_t seed, int ndraws);

/* ---- grid model: RegularGrid.cpp:426-441,831-842; SH/Fracturer/voxel.glsl:6-19 ---- */
/* decode_mode 0 = exact integer, 1 = reference float decode (finding 6) */
void orc_decode_position(uint32_t index, const uint32_t dims[3], int decode_mode, uint32_t out_xyz[3]);
/* CADScene.cpp:545-556 dims rule (interactive) */
void orc_dims_rule(const float aabb_min[3], const float aabb_max[3], uint32_t max_voxels, uint32_t out_dims[3]);

/* ---- V2: SAT occupancy — SRC/Geometry/3D/Intersections3D.h:204-420, RegularGrid.cpp:258-259, AABB.h:41,51 ---- */
/* returns 1 if the triangle intersects the box [bmin,bmax] */
int orc_tri_box_intersect(const float p1[3], const float p2[3], const float p3[3], const float bmin[3], const float bmax[3]);
/* grid must be zeroed by caller or clear=1.  margin (optional, N floats) receives the smallest relative
 * axis-test gap seen for each voxel (1e30 where no candidate triangle) for the 1e-6 epsilon report. */
int orc_voxelize_sat(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf,
                     const float aabb_min[3], const float aabb_max[3], const uint32_t dims[3],
                     uint16_t* grid, int clear, float* margin);
/* V1: Tetravoxelizer occupancy (SRC/Graphics/Core/Tetravoxelizer.cpp:198-315), rasteriser rule fixed in vf_oracle.cpp */
int orc_voxelize_solid(const float* verts, uint32_t nv, const uint32_t* faces, uint32_t nf, const float amin[3],
                       const float amax[3], const uint32_t dims[3], uint16_t* grid, int clear);

/* ---- S1/S2: seeding — SRC/Fracturer/Seeder.cpp:154-208,115-152; RegularGrid.cpp:543-564 ---- */
/* out_seeds: n x {x,y,z,label}.  *attempts (optional) receives the number of attempts consumed. */
/* Halton_sampler::sample(dimension 0..2, index) after init_faure (Utilities/HaltonSampler.h:572-632,1416-1446) */
float orc_halton(unsigned dimension, unsigned index);
int orc_seed_uniform(orc_rng*, const uint16_t* grid, const uint32_t dims[3], uint32_t n, int random_mode,
                     int location, uint32_t* out_seeds, uint32_t* attempts);
/* seeds: nseeds x uvec4, w rewritten in place;  frags: nfrags x uvec4 */
void orc_merge_seeds(const uint32_t* frags, uint32_t nfrags, uint32_t* seeds, uint32_t nseeds, int dfunc);
/* CADScene.cpp:626-655 call pattern: returns total seed count written to out (n + (n_extra? n + n_extra : 0)) */
int orc_make_seeds(orc_rng*, const uint16_t* grid, const uint32_t dims[3], uint32_t n, uint32_t n_extra,
                   int random_mode, int merge_dfunc, uint32_t* out_seeds, uint32_t out_capacity);
/* S3: Seeder::nearSeeds (Seeder.cpp:49-113); crand_mode 0 = MSVC rand() LCG on *crand_state, 1 = this process's ::rand() */
int orc_near_seeds(orc_rng* rng, int crand_mode, uint32_t* crand_state, const uint16_t* grid, const uint32_t dims[3], const uint32_t* frags,
                   uint32_t nfrags, uint32_t numImpacts, uint32_t numSeeds, uint32_t spreading, uint32_t* out, uint32_t cap);

/* ---- F1: naive — SRC/Fracturer/NaiveFracturer.cpp:26-68; SH/Fracturer/naiveFracturer-comp.glsl:19-43 ---- */
void orc_naive(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc, int decode_mode);

/* ---- F2/F3: flood — SRC/Fracturer/FloodFracturer.cpp:98-191; SH floodFracturer/disjointSet/disjointSetStack/undoMask ---- */
typedef struct {
    uint32_t levels;        /* BFS levels summed over rounds */
    uint32_t rounds;        /* outer disjoint rounds executed */
    uint32_t freed_voxels;  /* cells returned to FREE by disjointSetStack, summed */
    uint32_t max_dist;      /* largest geodesic distance assigned in round 1 */
} orc_flood_stats;
/* id_bits = 8: reference word layout (fragId | prefix<<8) with the disjoint rounds.
 * id_bits = 15: extension for > 254 labels, no prefixes (finding 7): whole word is the id.
 * algo 0 = level-synchronous min-claim BFS, 1 = Dijkstra on (dist, order) keys (must agree). */
int orc_flood(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc,
              int id_bits, int algo, orc_flood_stats* stats);
/* the geodesic (dist, order) key field of round 1, for slab tests: keys[N] = dist<<15|order, 0xFFFFFFFF wall, 0xFFFFFFFE unreached */
int orc_flood_keys(const uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds, int dfunc,
                   uint32_t* keys);
/* slab-local relaxation to a fixed point on a key field with fixed halo planes (host logic tests of the slab protocol).
 * keys: (xs+2) x Y x Z including one halo plane each side.  returns number of cells changed. */
uint64_t orc_relax_keys_slab(uint32_t* keys, uint32_t xs_with_halo, uint32_t Y, uint32_t Z, int nneigh);

/* ---- C1..C4: cleanup ---- */
void orc_remove_isolated_regions_cpu(uint16_t* grid, const uint32_t dims[3], const uint32_t* seeds, uint32_t nseeds); /* NaiveFracturer.cpp:111-150 */
void orc_detect_boundaries(uint16_t* grid, const uint32_t dims[3], int boundary_size);  /* SH detectBoundaries-comp.glsl:18-43 */
void orc_fill_noise(orc_rng*, float* noise, uint32_t n);                                 /* RegularGrid.cpp:238-244, serial */
/* boundary_mode 0 = as written (erodeGrid-comp.glsl:31: any non-zero low 15 bits), 1 = intended (bit 15 set) */
void orc_erode(uint16_t* grid, const uint32_t dims[3], int type, uint32_t size, uint32_t iters, float prob, float thr,
               const float* noise, uint32_t nnoise, int boundary_mode);                  /* RegularGrid.cpp:82-159 */
void orc_erode_mask(int type, uint32_t size, float* mask /* k^3, k=odd(size) */, float* activations, uint32_t* k_out);
void orc_remove_isolated_regions_grid(uint16_t* grid, const uint32_t dims[3]);           /* removeIsolatedRegionsGrid-comp.glsl:16-39, snapshot semantics */
void orc_undo_mask(uint16_t* grid, uint64_t n, uint32_t position, int rightmost);        /* undoMask-comp.glsl:19-37 */
void orc_reset_filling(uint16_t* grid, uint64_t n);                                      /* RegularGrid.cpp:412-418 */
void orc_homogenize(uint16_t* grid, uint64_t n);                                         /* RegularGrid.cpp:533-541 */

/* ---- H1: histogram — RegularGrid.cpp:601-625, 280-287 ---- */
/* counts: 32768 entries indexed by (value & 0x7FFF); returns number of cells with value > FREE */
uint64_t orc_count_values(const uint16_t* grid, uint64_t n, uint32_t* counts);

/* ---- X1: formats — RegularGrid.cpp:627-714 ---- */
/* returns bytes written (or needed if out==NULL) */
uint64_t orc_encode_rle(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap);
int      orc_decode_rle(const uint8_t* data, uint64_t len, uint32_t dims_out[3], uint16_t* grid, uint64_t grid_cap);
uint64_t orc_encode_bing_squared(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap);
/* RegularGrid::exportVox (RegularGrid.cpp:740-798) through VoxWriter (Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp:449-540) */
uint64_t orc_encode_vox(const uint16_t* grid, const uint32_t dims[3], int squared, uint8_t* out, uint64_t cap);
/* f2: per-fragment marching cubes (MarchingCubes::triangulateFieldGPU, MarchingCubes.cpp:364-432 + shaders); verts float[nv][4] =
 * xyz + boundary flag, faces uint32[nf][4] = three vertex numbers + boundary flag; counts = {nv, nf}; buffers too small = size query */
int orc_marching_cubes(const uint16_t* grid, const uint32_t dims[3], uint32_t target, const float amin[3], const float amax[3], uint32_t nb_iters,
                       float nb_weight, uint32_t b_iters, float b_weight, float* verts, uint32_t cap_v, uint32_t* faces, uint32_t cap_f, uint32_t counts[2]);
/* exportQuadStack, RegularGrid.cpp:716-725 over SRC/DataStructures/QuadStack.h + GStack.h */
uint64_t orc_encode_qstack(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap);

/* ---- composite used by the CPU baseline (bench.py): cfg3 pipeline on one grid ---- */
int orc_num_threads(void);
void orc_set_num_threads(int n); /* torchrun exports OMP_NUM_THREADS=1; the CPU baseline asks for all host cores explicitly */

#ifdef __cplusplus
}
#endif
#endif
