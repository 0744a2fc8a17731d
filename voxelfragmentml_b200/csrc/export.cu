// export.cu — X1: on-disk grid formats (host side; the grid is downloaded once per export).
//
// Replaces RegularGrid::exportGrid (SRC/DataStructures/RegularGrid.cpp:161-171): exportRLE (:672-714) and the squared
// layout of exportRawCompressed (:638-666).  The byte layouts are unchanged: `.rle` = uvec3 dims + packed {uint16 value,
// uint32 repetitions} runs over the x-major array (decoder: docs/decompress/decompress_grid.py:16-33); `.bing` squared =
// uvec3(M,M,M) + M^3 uint16 with the grid centred at (M - dims) / 2 and EMPTY padding.
// The reference's non-squared `.bing` writes the std::vector object instead of its data (:636, SURVEY finding 10); here it
// writes the intended dims + raw cells.  `.vox` / `.qstack` are not on this round's path (VF_ERR_UNSUPPORTED).
#include <algorithm>
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

extern "C" vf_status vf_export(vf_grid* g, const char* path, int type, int squared)
{
    VF_REQUIRE(g && path, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(g->ctx));
    static const char* ext[4] = { "rle", "qstack", "vox", "bing" };  // FractureParameters::ExportGrid_STR, FractureParameters.h:36
    VF_REQUIRE(type >= 0 && type < 4, VF_ERR_INVALID_ARGUMENT, "bad export type %d", type);
    VF_REQUIRE(type == VF_RLE || type == VF_UNCOMPRESSED_BINARY, VF_ERR_UNSUPPORTED, ".%s export is not implemented in this round", ext[type]);
    std::vector<uint16_t> host(g->n());
    VF_TRY(vf_grid_download(g, host.data()));
    const uint32_t dims[3] = { g->X, g->Y, g->Z };
    std::vector<uint8_t> bytes;
    if (type == VF_RLE) {
        bytes.resize(vf_encode_rle(host.data(), dims, nullptr, 0));
        vf_encode_rle(host.data(), dims, bytes.data(), bytes.size());
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
