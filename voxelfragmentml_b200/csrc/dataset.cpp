// dataset.cpp — the caller that turns the kernels into models/s: CADScene::generateDataset (SRC/Graphics/Application/CADScene.cpp:209-507)
// restricted to what lies on the voxel path: per model the metric dims rule (:262-273), setAABB + fill (:281-282),
// the export of the starting grid (:291-292 -> exportGrid(params, folder) :77-94), then for numFragments in _fragmentInterval
// (:298-306: numSeeds = n, numExtraSeeds = 2n, numIterations = glm::mix(_iterationInterval.x, .y, t)) and for every iteration
// fractureGrid(…, false) (:170-181 = resetFilling + fractureModel :624-691), prepareScene's grid side (:813-832: countValues ->
// fragment metadata, undoMask), exportGrid(itFile, true, ext) (:331) and the metadata rows; the _maxFragmentsModel cap (:298,313,
// :447); exportMetadata's three tab-separated files (:580-622).  File names are the reference's
// (<dest><model>/<model>_<n>f_<maxDim>r_<it>it.<ext>, <model>_grid_<maxDim>r.<ext>, <model>_<maxDim>_metadata_{grid,mesh,pointcloud}.txt).
//
// Not on this path (SURVEY §8 out of scope): Assimp loading (a minimal Wavefront .obj reader stands in: v / f records, fan
// triangulation, CADModel::load's normalisation :148-152), marching cubes -> fragment meshes, point clouds, zip post-processing.
// Fragment meshes (marching cubes, csrc/mesh.cu) are written as `.binm` with their metadata rows when the procedure asks for them
// (exportMesh: the reference's behaviour with an empty _targetTriangles list; mesh simplification is not on this path); the
// point-cloud metadata file holds its header only, as in the reference with _exportPointCloud switched off.
//
// Host design: the grid never leaves the GPU inside the loop — seeds are tested on the device, `.rle` runs are found on the
// device — and finished byte streams go to a small pool of writer threads so that file I/O overlaps the next fragmentation
// (the reference exports synchronously between two fragmentations, :331).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <filesystem>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "vf_internal.h"

namespace {

const char* kGridExt[4] = { "rle", "qstack", "vox", "bing" };  // FractureParameters::ExportGrid_STR (FractureParameters.h:36)

// writer pool: jobs are closures that encode (when the format is encoded on the host) and write one file
class Writers {
public:
    explicit Writers(int n)
    {
        for (int i = 0; i < n; ++i) threads_.emplace_back([this] { run(); });
    }
    ~Writers() { finish(); }
    void submit(std::function<bool()> job)
    {
        if (threads_.empty()) {
            if (!job()) failed_ = true;
            return;
        }
        std::unique_lock<std::mutex> lk(m_);
        // bounded queue: a producer far ahead of the disks would otherwise hold every grid copy in memory
        space_.wait(lk, [this] { return jobs_.size() < 4 * threads_.size(); });
        jobs_.push_back(std::move(job));
        work_.notify_one();
    }
    bool finish()
    {
        {
            std::unique_lock<std::mutex> lk(m_);
            done_ = true;
            work_.notify_all();
        }
        for (auto& t : threads_) t.join();
        threads_.clear();
        return !failed_;
    }

private:
    void run()
    {
        for (;;) {
            std::function<bool()> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                work_.wait(lk, [this] { return done_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                job = std::move(jobs_.front());
                jobs_.pop_front();
                space_.notify_one();
            }
            if (!job()) failed_ = true;
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::function<bool()>> jobs_;
    std::mutex m_;
    std::condition_variable work_, space_;
    bool done_ = false;
    std::atomic<bool> failed_{ false };
};

bool write_file(const std::string& path, const std::vector<uint8_t>& bytes)
{
    std::ofstream f(path, std::ios::out | std::ios::binary);
    if (!f.is_open()) return false;
    f.write(reinterpret_cast<const char*>(bytes.data()), (std::streamsize)bytes.size());
    return f.good();
}

double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// RegularGrid::exportGrid(name, squared, type) (RegularGrid.cpp:161-171) with the write handed to the pool
vf_status export_async(vf_grid* g, const std::string& base, int type, int squared, Writers& pool, vf_dataset_stats* st)
{
    const std::string path = base + "." + kGridExt[type];
    const uint32_t dims[3] = { g->X, g->Y, g->Z };
    if (type == VF_RLE && g->n() < (1ull << 32)) {
        auto bytes = std::make_shared<std::vector<uint8_t>>();
        uint64_t need = 0;
        VF_TRY(vf_grid_encode_rle(g, nullptr, 0, &need));
        bytes->resize(need);
        VF_TRY(vf_grid_encode_rle(g, bytes->data(), bytes->size(), &need));
        if (st) st->bytes_written += need, st->bytes_downloaded += need;
        pool.submit([path, bytes] { return write_file(path, *bytes); });
        return VF_OK;
    }
    auto host = std::make_shared<std::vector<uint16_t>>(g->n());
    VF_TRY(vf_grid_download(g, host->data()));
    if (st) st->bytes_downloaded += host->size() * 2;
    const uint32_t d0 = dims[0], d1 = dims[1], d2 = dims[2];
    pool.submit([path, host, type, squared, d0, d1, d2] {
        const uint32_t d[3] = { d0, d1, d2 };
        std::vector<uint8_t> bytes;
        if (type == VF_RLE) {
            bytes.resize(vf_encode_rle(host->data(), d, nullptr, 0));
            vf_encode_rle(host->data(), d, bytes.data(), bytes.size());
        } else if (type == VF_QUADSTACK) {
            bytes.resize(vf_encode_qstack(host->data(), d, nullptr, 0));
            vf_encode_qstack(host->data(), d, bytes.data(), bytes.size());
        } else if (type == VF_VOX) {
            bytes.resize(vf_encode_vox(host->data(), d, squared, nullptr, 0));
            vf_encode_vox(host->data(), d, squared, bytes.data(), bytes.size());
        } else if (squared) {
            bytes.resize(vf_encode_bing_squared(host->data(), d, nullptr, 0));
            vf_encode_bing_squared(host->data(), d, bytes.data(), bytes.size());
        } else {
            bytes.resize(12 + host->size() * 2);
            std::memcpy(bytes.data(), d, 12);
            std::memcpy(bytes.data() + 12, host->data(), host->size() * 2);
        }
        return write_file(path, bytes);
    });
    return VF_OK;
}

// CADModel::saveBinary (SRC/Graphics/Core/CADModel.cpp:839-855): uint32 numVertices, numVertices x Model3D::VertexGPUData (64 bytes: position
// + padding, normal + padding, texture coordinate + padding, tangent + padding — everything but the position is zero in dataset mode,
// where CADModel::insert value-initialises the struct (:93-101) and endInsertionBatch skips computeMeshData (:66-83)), uint32 numTriangles,
// numTriangles x Model3D::FaceGPUData (uvec3 vertices + modelCompID = 0)
std::vector<uint8_t> encode_binm(const std::vector<float>& v4, const std::vector<uint32_t>& f4)
{
    const uint32_t nv = (uint32_t)(v4.size() / 4), nf = (uint32_t)(f4.size() / 4);
    std::vector<uint8_t> out(8 + (size_t)nv * 64 + (size_t)nf * 16, 0);
    uint8_t* p = out.data();
    std::memcpy(p, &nv, 4), p += 4;
    for (uint32_t i = 0; i < nv; ++i, p += 64) std::memcpy(p, &v4[4 * (size_t)i], 12);
    std::memcpy(p, &nf, 4), p += 4;
    for (uint32_t i = 0; i < nf; ++i, p += 16) std::memcpy(p, &f4[4 * (size_t)i], 12);
    return out;
}

std::string dims_str(const uint32_t d[3])
{
    return std::to_string(d[0]) + "x" + std::to_string(d[1]) + "x" + std::to_string(d[2]);
}

}  // namespace

extern "C" void vf_procedure_default(vf_procedure* p)
{
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    vf_params_default(&p->fractureParameters);
    // FragmentationProcedure::FragmentationProcedure (FragmentationProcedure.h:41-60)
    p->fractureParameters.biasSeeds = 0;
    p->fractureParameters.erode = 0;
    p->fractureParameters.voxelPerMetricUnit = p->fractureParameters.clampVoxelMetricUnit;
    p->fractureParameters.exportGridExtension = VF_RLE;
    p->fragmentInterval[0] = 2, p->fragmentInterval[1] = 10;    // :12
    p->iterationInterval[0] = 25, p->iterationInterval[1] = 15; // :13
    p->maxFragmentsModel = 1000;                                // :16
    p->exportGrid = 1;                                          // :52
    p->solidVoxelization = 0;
    p->exportMesh = 0;
    p->writerThreads = 2;
}

// CADScene.cpp:262-273: ceil(aabb.size() * voxelPerMetricUnit); clamp rule; x and z up to the next multiple of 4 (y is left alone)
extern "C" void vf_dataset_dims_rule(const float mn[3], const float mx[3], int32_t voxelPerMetricUnit, int32_t clampVoxelMetricUnit, uint32_t dims[3])
{
    float size[3];
    int v[3];
    for (int q = 0; q < 3; ++q) size[q] = mx[q] - mn[q], v[q] = (int)std::ceil(size[q] * (float)voxelPerMetricUnit);
    if (v[0] > clampVoxelMetricUnit || v[1] > clampVoxelMetricUnit || v[2] > clampVoxelMetricUnit) {
        const float m = std::max(size[0], std::max(size[1], size[2]));
        for (int q = 0; q < 3; ++q) v[q] = (int)std::floor(((float)clampVoxelMetricUnit * size[q]) / m);
    }
    while (v[0] % 4 != 0) ++v[0];
    while (v[2] % 4 != 0) ++v[2];
    for (int q = 0; q < 3; ++q) dims[q] = (uint32_t)std::max(v[q], 0);
}

// glm::mix(int x, int y, float a) = int(float(x) * (1 - a) + float(y) * a)  (CADScene.cpp:304-306)
extern "C" int32_t vf_dataset_iterations(const vf_procedure* p, int32_t numFragments)
{
    const int32_t span = p->fragmentInterval[1] - p->fragmentInterval[0];
    if (span == 0) return p->iterationInterval[0];  // the reference divides by zero here; one fragment count = the first iteration count
    const float a = (float)(numFragments - p->fragmentInterval[0]) / (float)span;
    return (int32_t)((float)p->iterationInterval[0] * (1.0f - a) + (float)p->iterationInterval[1] * a);
}

extern "C" vf_status vf_dataset_model(vf_grid* g, const vf_procedure* proc, const char* model_name, const float* verts, uint32_t nv, const uint32_t* faces,
                                      uint32_t nf, const char* destination_folder, vf_dataset_stats* stats)
{
    VF_REQUIRE(g && proc && model_name && verts && faces && destination_folder, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_REQUIRE(nv > 0 && nf > 0, VF_ERR_INVALID_ARGUMENT, "empty mesh");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    vf_params fp = proc->fractureParameters;
    const int ext = fp.exportGridExtension;
    VF_REQUIRE(ext >= 0 && ext < 4, VF_ERR_INVALID_ARGUMENT, "bad export type %d", ext);
    vf_dataset_stats local;
    std::memset(&local, 0, sizeof(local));
    namespace fs = std::filesystem;
    std::error_code ec;
    std::string dest = destination_folder;
    if (!dest.empty() && dest.back() != '/') dest += '/';
    const std::string name = model_name;
    const std::string meshFolder = dest + name + "/", meshFile = meshFolder + name + "_";  // :249-251
    fs::create_directories(meshFolder, ec);
    VF_REQUIRE(fs::is_directory(meshFolder), VF_ERR_IO, "cannot create %s", meshFolder.c_str());

    // model AABB (Model3D keeps it while loading) and the voxelization size (:262-273)
    float mn[3] = { verts[0], verts[1], verts[2] }, mx[3] = { verts[0], verts[1], verts[2] };
    for (uint32_t i = 1; i < nv; ++i)
        for (int q = 0; q < 3; ++q) mn[q] = std::min(mn[q], verts[3 * (size_t)i + q]), mx[q] = std::max(mx[q], verts[3 * (size_t)i + q]);
    uint32_t dims[3];
    vf_dataset_dims_rule(mn, mx, fp.voxelPerMetricUnit, fp.clampVoxelMetricUnit, dims);
    VF_REQUIRE(dims[0] && dims[1] && dims[2], VF_ERR_INVALID_ARGUMENT, "degenerate model AABB");
    const uint32_t maxDimension = std::max(dims[0], std::max(dims[1], dims[2]));  // :278
    for (int q = 0; q < 3; ++q) fp.voxelizationSize[q] = (int32_t)dims[q];

    Writers pool(std::max(0, proc->writerThreads));
    double t0 = now();
    VF_TRY(vf_grid_set_aabb(g, mn, mx, dims[0], dims[1], dims[2]));  // :281
    if (proc->solidVoxelization) {
        uint64_t occ = 0;
        VF_TRY(vf_voxelize_solid(g, verts, nv, faces, nf, &occ));
        // RegularGrid::fill, RegularGrid.cpp:205-211: when the solid voxelizer sets nothing (an open surface) the reference falls back to
        // fillNaive (:800-816), which marks the cells hit by ~10^4 RANDOM surface samples drawn under an OpenMP race — no two runs of the
        // reference agree on them.  The deterministic stand-in is the exact surface occupancy (every cell a triangle touches, vf_voxelize),
        // a superset of every sample set fillNaive can draw.
        if (occ == 0) VF_TRY(vf_voxelize(g, verts, nv, faces, nf));
    } else {
        VF_TRY(vf_voxelize(g, verts, nv, faces, nf));
    }
    VF_TRY(vf_ctx_synchronize(c));
    local.seconds_voxelize += now() - t0;
    std::vector<std::string> gridRows, meshRows;  // VOXEL / MESH metadata rows (:333-337, :413-418), in order
    if (proc->exportGrid) {  // :291-292 -> exportGrid(params, folder) :89
        t0 = now();
        VF_TRY(export_async(g, meshFolder + name + "_grid_" + std::to_string(maxDimension) + "r", ext, 1, pool, &local));
        ++local.files;
        local.seconds_export += now() - t0;
    }

    std::vector<uint32_t> counts(VF_HISTOGRAM_BINS);
    uint64_t numGeneratedFragments = 0;
    for (int numFragments = proc->fragmentInterval[0]; numFragments <= proc->fragmentInterval[1] && numGeneratedFragments < proc->maxFragmentsModel; ++numFragments) {
        const std::string fragmentFile = meshFile + std::to_string(numFragments) + "f_";
        fp.numExtraSeeds = numFragments * 2;  // :302-303
        fp.numSeeds = numFragments;
        const int numIterations = vf_dataset_iterations(proc, numFragments);
        for (int iteration = 0; iteration < numIterations && numGeneratedFragments < proc->maxFragmentsModel; ++iteration) {
            const std::string itFile = fragmentFile + std::to_string(maxDimension) + "r_" + std::to_string(iteration) + "it";  // :316
            t0 = now();
            VF_TRY(vf_reset_filling(g));                                   // fractureGrid -> rebuildGrid (:174, :835-838)
            VF_TRY(vf_fracture_model(g, &fp, nullptr, nullptr, nullptr));  // :175
            uint64_t occupied = 0;                                          // prepareScene -> toTriangleMesh: countValues (:813, RegularGrid.cpp:443-471)
            uint64_t fragments = 0;
            if (!proc->exportMesh) {
                VF_TRY(vf_histogram_undo_mask(g, counts.data(), &occupied));  // ... and undoMask (:832), one pass
            } else {
                // toTriangleMesh (RegularGrid.cpp:443-486): one mesh per value in ascending order, before undoMask clears the boundary tags
                VF_TRY(vf_histogram(g, counts.data(), &occupied));
                uint32_t idx = 0;
                for (uint32_t v = 2; v < VF_HISTOGRAM_BINS; ++v) {
                    if (!counts[v]) continue;
                    vf_mesh* mesh = nullptr;
                    VF_TRY(vf_marching_cubes(g, v, nullptr, &mesh));
                    uint32_t nv = 0, nf = 0;
                    vf_mesh_counts(mesh, &nv, &nf);
                    auto v4 = std::make_shared<std::vector<float>>(4 * (size_t)nv);
                    auto f4 = std::make_shared<std::vector<uint32_t>>(4 * (size_t)nf);
                    const vf_status ds = vf_mesh_download(mesh, v4->data(), f4->data());
                    vf_mesh_destroy(mesh);
                    VF_TRY(ds);
                    local.bytes_downloaded += 16ull * nv + 16ull * nf;
                    // CADScene.cpp:344-420 with _targetTriangles empty: <itFile>_<idx>.binm, and the fragment's metadata row (:413-418, exportMetadata :605-615)
                    const std::string filename = itFile + "_" + std::to_string(idx);
                    std::ostringstream row;
                    row << filename << ".binm\t" << idx << "\t" << dims_str(dims) << "\t" << counts[v] << "\t" << (uint32_t)occupied << "\t"
                        << (float)counts[v] / static_cast<float>((uint32_t)occupied) << "\t" << nv << "\t" << nf << "\t";
                    meshRows.push_back(row.str());
                    pool.submit([filename, v4, f4] { return write_file(filename + ".binm", encode_binm(*v4, *f4)); });
                    ++local.files;
                    ++idx;
                }
                VF_TRY(vf_undo_mask(g));  // :832
            }
            local.seconds_fracture += now() - t0;
            for (uint32_t v = 2; v < VF_HISTOGRAM_BINS; ++v) fragments += counts[v] != 0;
            if (proc->exportGrid) {  // :325-338
                t0 = now();
                VF_TRY(export_async(g, itFile, ext, 1, pool, &local));
                ++local.files;
                gridRows.push_back(itFile + "." + kGridExt[ext] + "\t" + dims_str(dims));
                local.seconds_export += now() - t0;
            }
            numGeneratedFragments += fragments;  // :444 (_fractureMeshes.size() = one mesh per fragment value)
            ++local.fragmentations;
            local.fragments += fragments;
            local.voxels += occupied;
        }
    }

    // exportMetadata (:580-622)
    {
        const std::string stem = meshFile + std::to_string(maxDimension);
        std::ofstream gridOut(stem + "_metadata_grid.txt"), meshOut(stem + "_metadata_mesh.txt"), pcOut(stem + "_metadata_pointcloud.txt");
        VF_REQUIRE(!gridOut.fail() && !meshOut.fail() && !pcOut.fail(), VF_ERR_IO, "cannot write metadata under %s", meshFolder.c_str());
        gridOut << "Filename\tVoxelization size" << std::endl;
        meshOut << "Filename\tFragment id\tVoxelization size\tVoxels\tOccupied voxels\tPercentage\tVertices\tFaces" << std::endl;
        pcOut << "Filename\tVoxelization size\tPoints" << std::endl;
        for (const std::string& r : gridRows) gridOut << r << std::endl;
        for (const std::string& r : meshRows) meshOut << r << std::endl;
        local.files += 3;
    }
    t0 = now();
    const bool ok = pool.finish();  // :455-466 "Waiting threads to finish..."
    local.seconds_export += now() - t0;
    ++local.models;
    if (stats) {
        stats->models += local.models, stats->fragmentations += local.fragmentations, stats->fragments += local.fragments;
        stats->files += local.files, stats->bytes_written += local.bytes_written, stats->bytes_downloaded += local.bytes_downloaded;
        stats->voxels += local.voxels;
        stats->seconds_voxelize += local.seconds_voxelize, stats->seconds_fracture += local.seconds_fracture, stats->seconds_export += local.seconds_export;
    }
    VF_REQUIRE(ok, VF_ERR_IO, "a writer thread could not write under %s", meshFolder.c_str());
    return VF_OK;
}

// ---- minimal Wavefront reader standing in for Assimp (CADModel::load, SRC/Graphics/Core/CADModel.cpp:113-175) ------------------
extern "C" vf_status vf_load_obj(const char* path, float** verts_out, uint32_t* nv_out, uint32_t** faces_out, uint32_t* nf_out)
{
    VF_REQUIRE(path && verts_out && nv_out && faces_out && nf_out, VF_ERR_INVALID_ARGUMENT, "null argument");
    std::ifstream in(path);
    VF_REQUIRE(in.is_open(), VF_ERR_IO, "cannot open %s", path);
    std::vector<float> v;
    std::vector<uint32_t> f;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        std::string tag;
        ss >> tag;
        if (tag == "v") {
            float x, y, z;
            if (ss >> x >> y >> z) v.push_back(x), v.push_back(y), v.push_back(z);
        } else if (tag == "f") {
            std::vector<int64_t> idx;
            std::string tok;
            while (ss >> tok) {
                const int64_t i = std::strtoll(tok.c_str(), nullptr, 10);  // "i", "i/t", "i//n", "i/t/n": the position index comes first
                if (i == 0) continue;
                idx.push_back(i > 0 ? i - 1 : (int64_t)(v.size() / 3) + i);  // negative = relative to the vertices read so far
            }
            for (size_t k = 2; k < idx.size(); ++k) f.push_back((uint32_t)idx[0]), f.push_back((uint32_t)idx[k - 1]), f.push_back((uint32_t)idx[k]);  // aiProcess_Triangulate
        }
    }
    const uint32_t nv = (uint32_t)(v.size() / 3), nf = (uint32_t)(f.size() / 3);
    VF_REQUIRE(nv > 0 && nf > 0, VF_ERR_IO, "%s holds no triangles", path);
    for (uint32_t i : f) VF_REQUIRE(i < nv, VF_ERR_IO, "%s: face index out of range", path);
    // normalisation of CADModel::load (:148-152): scale 0.499999 / max(extent) * 2 about the AABB centre, as the matrix product
    // glm::scale(s) * glm::translate(-centre) evaluates it: s * x + s * (-cx)
    float mn[3] = { v[0], v[1], v[2] }, mx[3] = { v[0], v[1], v[2] };
    for (uint32_t i = 1; i < nv; ++i)
        for (int q = 0; q < 3; ++q) mn[q] = std::min(mn[q], v[3 * (size_t)i + q]), mx[q] = std::max(mx[q], v[3 * (size_t)i + q]);
    float ctr[3], extent[3];
    for (int q = 0; q < 3; ++q) ctr[q] = (mx[q] + mn[q]) / 2.0f, extent[q] = mx[q] - ctr[q];  // AABB.h:41,51
    const float s = 0.499999f / std::max(extent[0], std::max(extent[1], extent[2])) * 2.0f;
    for (uint32_t i = 0; i < nv; ++i)
        for (int q = 0; q < 3; ++q) v[3 * (size_t)i + q] = s * v[3 * (size_t)i + q] + s * (-ctr[q]);
    *verts_out = (float*)std::malloc(v.size() * sizeof(float));
    *faces_out = (uint32_t*)std::malloc(f.size() * sizeof(uint32_t));
    VF_REQUIRE(*verts_out && *faces_out, VF_ERR_CAPACITY, "out of host memory");
    std::memcpy(*verts_out, v.data(), v.size() * sizeof(float));
    std::memcpy(*faces_out, f.data(), f.size() * sizeof(uint32_t));
    *nv_out = nv, *nf_out = nf;
    return VF_OK;
}

extern "C" void vf_free_host(void* p) { std::free(p); }

// generateDataset (:209-246, 497-506): searchFiles (FileManagement.h:93-102: recursive, substring match on the extension), the
// _startVessel skip (:223-230), one grid allocated at the clamp size for the whole run (allocateMemoryDataset :529-543)
extern "C" vf_status vf_dataset_generate(vf_ctx* ctx, const vf_procedure* proc, const char* folder, const char* extension, const char* start_vessel,
                                         const char* destination_folder, vf_dataset_stats* stats)
{
    VF_REQUIRE(ctx && proc && folder && extension && destination_folder, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(ctx));
    namespace fs = std::filesystem;
    std::error_code ec;
    std::vector<std::string> files;
    for (fs::recursive_directory_iterator it(folder, ec), end; !ec && it != end; it.increment(ec))
        if (!it->is_directory() && it->path().generic_string().find(extension) != std::string::npos) files.push_back(it->path().generic_string());
    VF_REQUIRE(!files.empty(), VF_ERR_IO, "No files found in %s", folder);  // :213-214
    std::sort(files.begin(), files.end());  // directory order is unspecified; sorted here so that the RNG stream maps to models reproducibly
    if (start_vessel && *start_vessel) {
        do {
            files.erase(files.begin());
        } while (!files.empty() && files[0].find(start_vessel) == std::string::npos);
    }
    fs::create_directories(destination_folder, ec);
    const int32_t clamp = proc->fractureParameters.clampVoxelMetricUnit;
    VF_REQUIRE(clamp > 0, VF_ERR_INVALID_ARGUMENT, "clampVoxelMetricUnit must be positive");
    vf_grid* grid = nullptr;
    // x and z may be rounded up past the clamp by the multiple-of-4 rule
    VF_TRY(vf_grid_create(ctx, (uint32_t)clamp + 3, (uint32_t)clamp, (uint32_t)clamp + 3, &grid));
    vf_status rc = vf_ctx_reserve(ctx, (uint32_t)clamp + 3, (uint32_t)clamp, (uint32_t)clamp + 3);
    for (size_t i = 0; rc == VF_OK && i < files.size(); ++i) {
        float* v = nullptr;
        uint32_t *f = nullptr, nv = 0, nf = 0;
        rc = vf_load_obj(files[i].c_str(), &v, &nv, &f, &nf);
        if (rc == VF_OK) {
            const std::string shortName = fs::path(files[i]).stem().string();
            rc = vf_dataset_model(grid, proc, shortName.c_str(), v, nv, f, nf, destination_folder, stats);
        }
        std::free(v), std::free(f);
    }
    vf_grid_destroy(grid);
    return rc;
}
