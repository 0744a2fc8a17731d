// export.cu — X1: on-disk grid formats (host side; the grid is downloaded once per export).
//
// Replaces RegularGrid::exportGrid (SRC/DataStructures/RegularGrid.cpp:161-171): exportRLE (:672-714) and the squared
// layout of exportRawCompressed (:638-666).  The byte layouts are unchanged: `.rle` = uvec3 dims + packed {uint16 value,
// uint32 repetitions} runs over the x-major array (decoder: docs/decompress/decompress_grid.py:16-33); `.bing` squared =
// uvec3(M,M,M) + M^3 uint16 with the grid centred at (M - dims) / 2 and EMPTY padding.
// The reference's non-squared `.bing` writes the std::vector object instead of its data (:636, SURVEY finding 10); here it
// writes the intended dims + raw cells.
// `.vox` = exportVox (:740-798) over the vendored MagicaVoxel writer (Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp): see
// vf_encode_vox below.  `.qstack` is not on this round's path (VF_ERR_UNSUPPORTED).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "vf_internal.h"

extern "C" uint64_t vf_encode_rle(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    const uint64_t size = (uint64_t)dims[0] * dims[1] * dims[2];
    uint64_t pos = 12;
    if (out && cap >= 12) std::memcpy(out, dims, 12);
    uint64_t idx = 0;
    while (idx < size) {
        const uint16_t value = grid[idx];
        uint64_t end = idx + 1;
        while (end < size && grid[end] == value) ++end;
        const uint32_t rep = (uint32_t)(end - idx);  // size < 2^32 cells in the reference (uint32_t size, :681)
        if (out && pos + 6 <= cap) {
            std::memcpy(out + pos, &value, 2);
            std::memcpy(out + pos + 2, &rep, 4);
        }
        pos += 6;
        idx = end;
    }
    return pos;
}

extern "C" uint64_t vf_encode_bing_squared(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    const uint32_t M = std::max(dims[0], std::max(dims[1], dims[2]));
    const uint64_t need = 12 + (uint64_t)M * M * M * 2;
    if (!out || cap < need) return need;
    const uint32_t end[3] = { M, M, M };
    std::memcpy(out, end, 12);
    uint16_t* cube = reinterpret_cast<uint16_t*>(out + 12);  // 12-byte header keeps 2-byte alignment
    std::memset(cube, 0, (size_t)M * M * M * 2);
    const uint32_t sx = (M - dims[0]) / 2, sy = (M - dims[1]) / 2, sz = (M - dims[2]) / 2;  // :643
    for (uint32_t x = 0; x < dims[0]; ++x)
        for (uint32_t y = 0; y < dims[1]; ++y)
            std::memcpy(cube + ((size_t)(x + sx) * M + (y + sy)) * M + sz, grid + ((size_t)x * dims[1] + y) * dims[2], (size_t)dims[2] * 2);
    return need;
}

// ---- .vox ---------------------------------------------------------------------------------------------------------------
// File layout produced by VoxWriter::SaveToFile (VoxWriter.cpp:462-540) for the calls exportVox makes (no AddColor, one key
// frame): 'VOX ' 150 | MAIN 0 <children bytes> | per cube {SIZE, XYZI} | nTRN root | nGRP | per cube {nTRN, nSHP}.
// exportVox feeds AddVoxel(x, z, y, colour) in x,y,z order (RegularGrid.cpp:762,781); the writer cuts voxel space into 126^3
// cubes keyed floor(v / 126), numbers them in order of first appearance and appends {v % 126, colour} to the cube's XYZI
// payload (:449-461, :631-663).  The same bytes are produced here by two passes over the z-rows (count per cube, then fill at
// each cube's payload offset) instead of one std::map node per voxel.
namespace {

constexpr uint32_t kVoxCube = 126;  // VoxWriter's default per-cube limit (VoxWriter.h:451)

constexpr uint32_t vox_tag(char a, char b, char c, char d)
{
    return (uint32_t)(uint8_t)a | (uint32_t)(uint8_t)b << 8 | (uint32_t)(uint8_t)c << 16 | (uint32_t)(uint8_t)d << 24;
}

struct VoxBytes {
    uint8_t* out;
    uint64_t cap, pos;
    void raw(const void* p, uint64_t n)
    {
        if (out && pos + n <= cap) std::memcpy(out + pos, p, n);
        pos += n;
    }
    void i32(int32_t v) { raw(&v, 4); }
    void str(const std::string& s) { i32((int32_t)s.size()), raw(s.data(), s.size()); }          // DICTstring::write (:37-41)
    void chunk(uint32_t id, uint64_t content) { i32((int32_t)id), i32((int32_t)content), i32(0); }  // size_t sizes go out as 4 bytes
};

// (int)std::floor(v) as the reference's x86-64 build evaluates it: out-of-range doubles become INT_MIN (cvttsd2si)
inline int32_t vox_to_int(double v) { return (v >= -2147483648.0 && v < 2147483648.0) ? (int32_t)v : INT_MIN; }

}  // namespace

extern "C" uint64_t vf_encode_vox(const uint16_t* grid, const uint32_t dims[3], int squared, uint8_t* out, uint64_t cap)
{
    // iteration box, and where the grid sits inside it (squared: centred in the M^3 cube, RegularGrid.cpp:749-752)
    const uint32_t M = std::max(dims[0], std::max(dims[1], dims[2]));
    const uint32_t E[3] = { squared ? M : dims[0], squared ? M : dims[1], squared ? M : dims[2] };
    const uint32_t S[3] = { squared ? (M - dims[0]) / 2 : 0, squared ? (M - dims[1]) / 2 : 0, squared ? (M - dims[2]) / 2 : 0 };
    // grid row under iteration row (x, y); null when the row is padding (unsigned wrap == the reference's ">= 0" test failing)
    auto row = [&](uint32_t x, uint32_t y) -> const uint16_t* {
        const uint32_t gx = x - S[0], gy = y - S[1];
        return (gx < dims[0] && gy < dims[1]) ? grid + ((size_t)gx * dims[1] + gy) * dims[2] : nullptr;
    };
    // colour index AddVoxel receives for iteration cell z of that row, or -1 when the cell is not added.  The uint16 cell goes
    // through a `const uint8_t&` parameter, i.e. it is truncated to its low byte.
    auto colour = [&](const uint16_t* r, uint32_t z) -> int {
        if (squared) return (r && z - S[2] < dims[2]) ? (uint8_t)r[z - S[2]] : 0;  // EMPTY padding is written too (:762-764)
        return r[z] > 1 ? (uint8_t)(r[z] - 1) : -1;                                // value > VOXEL_FREE, value - VOXEL_FREE (:780-781)
    };

    // writer coordinates: vX = x, vY = z, vZ = y
    const uint32_t ncy = (E[2] + kVoxCube - 1) / kVoxCube, ncz = (E[1] + kVoxCube - 1) / kVoxCube;
    const size_t ncubes = (size_t)((E[0] + kVoxCube - 1) / kVoxCube) * ncy * ncz;
    std::vector<uint64_t> count(ncubes, 0);
    std::vector<uint32_t> order;                                                    // cube keys by first appearance (:617-629)
    uint64_t lo[3] = { UINT64_MAX, UINT64_MAX, UINT64_MAX }, hi[3] = { 0, 0, 0 };  // maxVolume (:632)
    uint64_t min_ox = 10000000, last_oy = 0, last_oz = 0, total = 0;               // minCube* start at 1e7 (VoxWriter.h:428-430)
    for (uint32_t x = 0; x < E[0]; ++x)
        for (uint32_t y = 0; y < E[1]; ++y) {
            const uint16_t* r = row(x, y);
            for (uint32_t z0 = 0; z0 < E[2]; z0 += kVoxCube) {
                const uint32_t z1 = std::min(E[2], z0 + kVoxCube);
                uint32_t n = 0, zmin = 0, zmax = 0;
                for (uint32_t z = z0; z < z1; ++z)
                    if (colour(r, z) >= 0) {
                        if (!n) zmin = z;
                        zmax = z, ++n;
                    }
                if (!n) continue;
                const uint32_t ox = x / kVoxCube, oy = z0 / kVoxCube, oz = y / kVoxCube;
                const size_t key = ((size_t)ox * ncy + oy) * ncz + oz;
                if (!count[key]) order.push_back((uint32_t)key);
                count[key] += n, total += n;
                lo[0] = std::min<uint64_t>(lo[0], x), hi[0] = std::max<uint64_t>(hi[0], x);
                lo[1] = std::min<uint64_t>(lo[1], zmin), hi[1] = std::max<uint64_t>(hi[1], zmax);
                lo[2] = std::min<uint64_t>(lo[2], y), hi[2] = std::max<uint64_t>(hi[2], y);
                min_ox = std::min<uint64_t>(min_ox, ox);
                last_oy = oy, last_oz = oz;
            }
        }
    // :457-458 assign minCubeY = mini(minCubeX, oy) and minCubeZ = mini(minCubeX, oz) on every call — against minCubeX and
    // without accumulating — so what SaveToFile sees is the last voxel's cube against the running minimum of ox.
    const uint64_t min_oy = std::min(min_ox, last_oy), min_oz = std::min(min_ox, last_oz);

    VoxBytes w{ out, cap, 0 };
    w.i32((int32_t)vox_tag('V', 'O', 'X', ' ')), w.i32(150);
    w.i32((int32_t)vox_tag('M', 'A', 'I', 'N')), w.i32(0);
    const uint64_t main_size_pos = w.pos;
    w.i32(0);
    const uint64_t header = w.pos;

    // per cube: SIZE + XYZI headers now, payload offsets remembered for the fill pass
    std::vector<uint64_t> cursor(ncubes, 0);
    for (uint32_t key : order) {
        w.chunk(vox_tag('S', 'I', 'Z', 'E'), 12);
        w.i32(kVoxCube), w.i32(kVoxCube), w.i32(kVoxCube);               // the cube limit, not the occupied extent (:675-677)
        const int32_t nvox = (int32_t)(uint32_t)(4 * count[key]) / 4;    // (int32_t)voxels.size() / 4 (:236)
        w.chunk(vox_tag('X', 'Y', 'Z', 'I'), 4ull * (uint64_t)(1 + (int64_t)nvox));
        w.i32(nvox);
        cursor[key] = w.pos;
        w.pos += 4 * count[key];
    }
    const bool writing = out && w.pos <= cap;  // payloads are filled only when they fit; later chunks are bounds-checked by raw()
    if (writing)
        for (uint32_t x = 0; x < E[0]; ++x)
            for (uint32_t y = 0; y < E[1]; ++y) {
                const uint16_t* r = row(x, y);
                const uint8_t bx = (uint8_t)(x % kVoxCube), bz = (uint8_t)(y % kVoxCube);
                for (uint32_t z0 = 0; z0 < E[2]; z0 += kVoxCube) {
                    const uint32_t z1 = std::min(E[2], z0 + kVoxCube);
                    const size_t key = ((size_t)(x / kVoxCube) * ncy + z0 / kVoxCube) * ncz + y / kVoxCube;
                    uint8_t* p = out + cursor[key];
                    for (uint32_t z = z0; z < z1; ++z) {
                        const int c = colour(r, z);
                        if (c < 0) continue;
                        p[0] = bx, p[1] = (uint8_t)(z - z0), p[2] = bz, p[3] = (uint8_t)c;
                        p += 4;
                    }
                    cursor[key] = (uint64_t)(p - out);
                }
            }

    // scene graph (:468-519): nTRN 0 -> nGRP 1 -> per cube nTRN (2, 4, ...) -> nSHP (3, 5, ...)
    const uint32_t kNTRN = vox_tag('n', 'T', 'R', 'N'), kNGRP = vox_tag('n', 'G', 'R', 'P'), kNSHP = vox_tag('n', 'S', 'H', 'P');
    const int32_t ncube = (int32_t)order.size();
    w.chunk(kNTRN, 4 * 5 + 4 + 4);
    w.i32(0), w.i32(0), w.i32(1), w.i32(-1), w.i32(-1), w.i32(1), w.i32(0);  // node, attribs{}, child, reserved, layer -1, 1 frame {}
    w.chunk(kNGRP, 4 * (2 + (uint64_t)ncube) + 4);
    w.i32(1), w.i32(0), w.i32(ncube);
    for (int32_t i = 0; i < ncube; ++i) w.i32(2 + 2 * i);
    const double lox = (double)lo[0], loy = (double)lo[1], sizex = (double)hi[0] - lox, sizey = (double)hi[1] - loy;
    for (int32_t i = 0; i < ncube; ++i) {
        const uint32_t key = order[i];
        const int32_t cx = (int32_t)(key / (ncy * ncz)), cy = (int32_t)(key / ncz % ncy), cz = (int32_t)(key % ncz);
        // :489-491 — `cube.tx - minCubeX` is size_t arithmetic (wraps if the minimum is larger), then float, then double
        const float fx = ((float)((uint64_t)(int64_t)cx - min_ox) + 0.5f) * (float)kVoxCube;
        const float fy = ((float)((uint64_t)(int64_t)cy - min_oy) + 0.5f) * (float)kVoxCube;
        const float fz = ((float)((uint64_t)(int64_t)cz - min_oz) + 0.5f) * (float)kVoxCube;
        const int32_t tx = vox_to_int(std::floor((double)fx - lox - sizex * 0.5));
        const int32_t ty = vox_to_int(std::floor((double)fy - loy - sizey * 0.5));
        const int32_t tz = vox_to_int((double)std::floor(fz));
        const std::string t = std::to_string(tx) + " " + std::to_string(ty) + " " + std::to_string(tz);
        w.chunk(kNTRN, 4 * 5 + 4 + 4 + (4 + 2) + (4 + t.size()));
        w.i32(2 + 2 * i), w.i32(0), w.i32(3 + 2 * i), w.i32(-1), w.i32(0), w.i32(1);  // layer 0 (:488)
        w.i32(1), w.str("_t"), w.str(t);
        w.chunk(kNSHP, 4 * 2 + 4 + 4 + 4 + (4 + 2) + (4 + 1));
        w.i32(3 + 2 * i), w.i32(0), w.i32(1);      // one model per cube: a single key frame
        w.i32(i), w.i32(1), w.str("_f"), w.str("0");  // modelId, {"_f": "0"} (:500-501)
    }
    // no RGBA chunk: exportVox never calls AddColor, so `colors` is empty (:524)
    const uint32_t children = (uint32_t)(w.pos - header);
    if (out && w.pos <= cap) std::memcpy(out + main_size_pos, &children, 4);
    return w.pos;
}

extern "C" vf_status vf_export(vf_grid* g, const char* path, int type, int squared)
{
    VF_REQUIRE(g && path, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(g->ctx));
    static const char* ext[4] = { "rle", "qstack", "vox", "bing" };  // FractureParameters::ExportGrid_STR, FractureParameters.h:36
    VF_REQUIRE(type >= 0 && type < 4, VF_ERR_INVALID_ARGUMENT, "bad export type %d", type);
    VF_REQUIRE(type != VF_QUADSTACK, VF_ERR_UNSUPPORTED, ".%s export is not implemented in this round", ext[type]);
    std::vector<uint16_t> host(g->n());
    VF_TRY(vf_grid_download(g, host.data()));
    const uint32_t dims[3] = { g->X, g->Y, g->Z };
    std::vector<uint8_t> bytes;
    if (type == VF_RLE) {
        bytes.resize(vf_encode_rle(host.data(), dims, nullptr, 0));
        vf_encode_rle(host.data(), dims, bytes.data(), bytes.size());
    } else if (type == VF_VOX) {
        bytes.resize(vf_encode_vox(host.data(), dims, squared, nullptr, 0));
        vf_encode_vox(host.data(), dims, squared, bytes.data(), bytes.size());
    } else if (squared) {
        bytes.resize(vf_encode_bing_squared(host.data(), dims, nullptr, 0));
        vf_encode_bing_squared(host.data(), dims, bytes.data(), bytes.size());
    } else {
        bytes.resize(12 + host.size() * 2);
        std::memcpy(bytes.data(), dims, 12);
        std::memcpy(bytes.data() + 12, host.data(), host.size() * 2);
    }
    const std::string file = std::string(path) + "." + ext[type];
    FILE* f = std::fopen(file.c_str(), "wb");
    VF_REQUIRE(f != nullptr, VF_ERR_IO, "cannot open %s", file.c_str());
    const size_t w = std::fwrite(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    VF_REQUIRE(w == bytes.size(), VF_ERR_IO, "short write to %s", file.c_str());
    return VF_OK;
}
