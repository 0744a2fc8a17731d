// export.cu — X1: on-disk grid formats (host side; the grid is downloaded once per export).
//
// Replaces RegularGrid::exportGrid (SRC/DataStructures/RegularGrid.cpp:161-171): exportRLE (:672-714) and the squared
// layout of exportRawCompressed (:638-666).  The byte layouts are unchanged: `.rle` = uvec3 dims + packed {uint16 value,
// uint32 repetitions} runs over the x-major array (decoder: docs/decompress/decompress_grid.py:16-33); `.bing` squared =
// uvec3(M,M,M) + M^3 uint16 with the grid centred at (M - dims) / 2 and EMPTY padding.
// The reference's non-squared `.bing` writes the std::vector object instead of its data (:636, SURVEY finding 10); here it
// writes the intended dims + raw cells.
// `.vox` = exportVox (:740-798) over the vendored MagicaVoxel writer (Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp): see
// vf_encode_vox below.  `.qstack` = exportQuadStack (:716-725) over DataStructures/QuadStack.h + GStack.h: see vf_encode_qstack.
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

// ---- .qstack ------------------------------------------------------------------------------------------------------------
// exportQuadStack (RegularGrid.cpp:716-725): QuadStack<uint16_t>::loadCube -> compress_y -> compress_x -> saveCheckpoint
// (DataStructures/QuadStack.h:99-140,91-96,188-226) over GStack<uint16_t> (DataStructures/GStack.h).  What the file holds:
//   z-columns are run-length coded (compress_y); a quadtree over (x, y) stops where every column of a region has the same value
//   sequence — run lengths are not compared (GStack.h:202-205) — and stores per layer the CUMULATIVE height field of the region
//   (QuadStack.h:289-303); bottom-up, a parent takes over layer i from all its children when they agree on its value
//   (GStack::mergeStacks, GStack.h:278-319), erasing it from the children while the layer counter keeps running, so the layer
//   after a merged one is skipped; every count is read through a uint8_t (GStack.h:50); nodes that keep at least one layer
//   are written in pre-order.
// The same bytes are produced here from flat arrays: columns in CSR form, nodes in creation (= pre-) order so that a reverse
// sweep is a post-order merge, height fields in one pool.  mergeStacks' recursion over per-child iterators collapses to one
// loop: below depth 0 every iterator starts at the previous one and the first candidate returns, so the only index tuples ever
// tried are (i, i, ..., i), and a tuple is tried iff every child still has a layer i.
namespace {

struct QsLayer {
    uint16_t value;
    uint64_t field;  // offset of the node-sized height field in the pool
};
struct QsNode {
    uint32_t x0, y0, x1, y1;  // _minPoints / _maxPoints
    int32_t child[4];         // [0][0], [0][1], [1][0], [1][1]; -1 = none
    std::vector<QsLayer> layers;
    uint32_t w() const { return x1 - x0; }
    uint32_t h() const { return y1 - y0; }
    uint8_t count() const { return (uint8_t)layers.size(); }
};

}  // namespace

// the quadtree stages over the run-length coded z-columns (CSR: column c owns runs first[c] .. first[c + 1])
static uint64_t qstack_from_columns(const std::vector<uint64_t>& first, const std::vector<uint16_t>& rval, const std::vector<uint16_t>& rlen, const uint32_t dims[3],
                                    uint8_t* out, uint64_t cap);

extern "C" uint64_t vf_encode_qstack(const uint16_t* grid, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    const uint32_t W = dims[0], H = dims[1], D = dims[2];
    if (!W || !H || !D || W > 0xFFFF || H > 0xFFFF || D > 0xFFFF) return 0;  // _width/_height/_depth are uint16_t (QuadStack.h:9)
    // compress_y: runs per column
    std::vector<uint64_t> first((size_t)W * H + 1, 0);
    std::vector<uint16_t> rval, rlen;
    for (size_t c = 0; c < (size_t)W * H; ++c) {
        const uint16_t* col = grid + c * D;
        for (uint32_t z = 0; z < D;) {
            uint32_t e = z + 1;
            while (e < D && col[e] == col[z]) ++e;
            rval.push_back(col[z]), rlen.push_back((uint16_t)(e - z));
            z = e;
        }
        first[c + 1] = rval.size();
    }
    return qstack_from_columns(first, rval, rlen, dims, out, cap);
}

// The same file from the `.rle` stream of the grid (12-byte header, then {uint16 value, uint32 repetitions} records over the x-major array,
// RegularGrid.cpp:672-714): a z-column is D consecutive cells of that array, so compress_y's column runs are the stream's runs cut at the
// multiples of D.  With the stream produced on the device (vf_grid_encode_rle) the `.qstack` export downloads runs, not 2 B per voxel.
static uint64_t qstack_from_rle_stream(const uint8_t* stream, uint64_t bytes, const uint32_t dims[3], uint8_t* out, uint64_t cap)
{
    const uint32_t W = dims[0], H = dims[1], D = dims[2];
    if (!W || !H || !D || W > 0xFFFF || H > 0xFFFF || D > 0xFFFF || bytes < 12) return 0;
    const uint64_t nrec = (bytes - 12) / 6, ncol = (uint64_t)W * H;
    std::vector<uint64_t> first(ncol + 1, 0);
    std::vector<uint16_t> rval, rlen;
    rval.reserve(nrec + ncol), rlen.reserve(nrec + ncol);
    uint64_t col = 0, z = 0;  // position of the next cell
    for (uint64_t r = 0; r < nrec; ++r) {
        uint16_t v;
        uint32_t rep;
        std::memcpy(&v, stream + 12 + 6 * r, 2), std::memcpy(&rep, stream + 12 + 6 * r + 2, 4);
        while (rep) {
            const uint32_t take = (uint32_t)std::min<uint64_t>(rep, D - z);
            rval.push_back(v), rlen.push_back((uint16_t)take);
            rep -= take, z += take;
            if (z == D) z = 0, first[++col] = rval.size();
        }
    }
    if (col != ncol) return 0;  // the stream does not cover the grid
    return qstack_from_columns(first, rval, rlen, dims, out, cap);
}

static uint64_t qstack_from_columns(const std::vector<uint64_t>& first, const std::vector<uint16_t>& rval, const std::vector<uint16_t>& rlen, const uint32_t dims[3],
                                    uint8_t* out, uint64_t cap)
{
    const uint32_t W = dims[0], H = dims[1], D = dims[2];
    auto same_values = [&](size_t a, size_t b) {
        const uint64_t na = first[a + 1] - first[a];
        return na == first[b + 1] - first[b] && std::memcmp(&rval[first[a]], &rval[first[b]], na * 2) == 0;
    };

    // buildQuadStack: explicit stack, children pushed in reverse so that nodes are numbered in the reference's recursion order
    std::vector<QsNode> nodes;
    std::vector<uint16_t> pool;
    nodes.push_back({ 0, 0, W, H, { -1, -1, -1, -1 }, {} });
    struct Pending { int32_t parent, slot; uint32_t x0, y0, x1, y1; };
    std::vector<Pending> pending;  // regions waiting to become nodes, LIFO
    auto visit = [&](int32_t id) {
        const uint32_t x0 = nodes[id].x0, y0 = nodes[id].y0, x1 = nodes[id].x1, y1 = nodes[id].y1;
        const size_t c0 = (size_t)x0 * H + y0;
        bool uniform = true;
        if (x1 - x0 > 1 || y1 - y0 > 1)
            for (uint32_t x = x0; x < x1 && uniform; ++x)
                for (uint32_t y = y0; y < y1 && uniform; ++y) uniform = same_values(c0, (size_t)x * H + y);
        if (uniform) {
            const uint32_t w = x1 - x0, h = y1 - y0;
            const uint64_t nl = first[c0 + 1] - first[c0];
            const uint64_t base = pool.size();
            pool.resize(base + nl * w * h);
            for (uint32_t x = 0; x < w; ++x)
                for (uint32_t y = 0; y < h; ++y) {
                    const uint64_t r = first[(size_t)(x0 + x) * H + y0 + y];
                    uint16_t top = 0;
                    for (uint64_t l = 0; l < nl; ++l) top = (uint16_t)(top + rlen[r + l]), pool[base + l * w * h + (size_t)x * h + y] = top;
                }
            for (uint64_t l = 0; l < nl; ++l) nodes[id].layers.push_back({ rval[first[c0] + l], base + l * w * h });
            return;
        }
        const uint32_t mx = x0 + (x1 - x0 + 1) / 2, my = y0 + (y1 - y0 + 1) / 2;
        const bool sx = x1 - x0 > 1, sy = y1 - y0 > 1;
        if (sx && sy) pending.push_back({ id, 3, mx, my, x1, y1 });
        if (sx) pending.push_back({ id, 2, mx, y0, x1, my });
        if (sy) pending.push_back({ id, 1, x0, my, mx, y1 });
        pending.push_back({ id, 0, x0, y0, mx, my });
    };
    visit(0);
    while (!pending.empty()) {
        const Pending p = pending.back();
        pending.pop_back();
        const int32_t id = (int32_t)nodes.size();
        nodes.push_back({ p.x0, p.y0, p.x1, p.y1, { -1, -1, -1, -1 }, {} });
        nodes[p.parent].child[p.slot] = id;
        visit(id);
    }

    // compressQuadStack: descendants have larger numbers than their ancestors
    for (int32_t id = (int32_t)nodes.size() - 1; id >= 0; --id) {
        int32_t kids[4], nk = 0;
        for (int32_t c : nodes[id].child)
            if (c >= 0) kids[nk++] = c;
        if (!nk) continue;
        const uint32_t w = nodes[id].w(), h = nodes[id].h();
        for (size_t i = 0; i < nodes[kids[0]].count(); ++i) {
            bool all = true;
            for (int32_t k = 1; k < nk; ++k) all = all && i < nodes[kids[k]].count();
            if (!all) continue;
            const uint16_t v = nodes[kids[0]].layers[i].value;
            for (int32_t k = 1; k < nk && all; ++k) all = nodes[kids[k]].layers[i].value == v;
            if (!all || v == 0xFFFF) continue;  // 0xFFFF is the wildcard value (GStack.h:171-175)
            const uint64_t base = pool.size();
            pool.resize(base + (uint64_t)w * h, 0);
            for (int32_t k = 0; k < nk; ++k) {
                QsNode& c = nodes[kids[k]];
                const uint16_t* src = &pool[c.layers[i].field];
                for (uint32_t x = 0; x < c.w(); ++x)
                    std::memcpy(&pool[base + (uint64_t)(c.x0 - nodes[id].x0 + x) * h + (c.y0 - nodes[id].y0)], src + (size_t)x * c.h(), (size_t)c.h() * 2);
                c.layers.erase(c.layers.begin() + i);
            }
            nodes[id].layers.push_back({ v, base });
        }
    }

    // saveCheckpoint (QuadStack.h:188-226); size_t fields are 8 bytes on the reference's x64 target
    VoxBytes f{ out, cap, 0 };
    uint64_t kept = 0;
    for (const QsNode& n : nodes) kept += n.count() != 0;
    const uint64_t tsize = 2;
    const uint16_t whd[3] = { (uint16_t)W, (uint16_t)H, (uint16_t)D };
    f.raw(&tsize, 8), f.raw(whd, 6), f.raw(&kept, 8);
    for (const QsNode& n : nodes) {
        const uint64_t nl = n.count();
        if (!nl) continue;
        const uint32_t mx[2] = { n.x1, n.y1 }, mn[2] = { n.x0, n.y0 };
        f.raw(&nl, 8), f.raw(mx, 8), f.raw(mn, 8);
        const uint8_t lw = (uint8_t)n.w(), lh = (uint8_t)n.h();  // written through a uint8_t (QuadStack.h:212-213)
        for (uint64_t l = 0; l < nl; ++l) {
            f.raw(&lw, 1), f.raw(&lh, 1), f.raw(&n.layers[l].value, 2);
            f.raw(&pool[n.layers[l].field], (uint64_t)n.w() * n.h() * 2);
        }
    }
    return f.pos;
}

// ---- .rle on the device -------------------------------------------------------------------------------------------------
// exportRLE (RegularGrid.cpp:672-714) walks the host copy of the grid cell by cell; with the grid resident in HBM that costs a
// 2 B/voxel download per export.  Here the runs are found where the grid lives: a cell starts a run iff it
// differs from its predecessor in the x-major array (cell 0 always does), so   count starts per 8192-cell tile -> exclusive scan
// of the tile counts -> every start writes {cell index, value} at its rank -> repetitions = next start - this start, packed
// into the file's 6-byte records {uint16 value, uint32 repetitions}.  Only the finished byte stream (12 + 6 R bytes, R = runs)
// crosses PCIe.  Reads 2 x 2 B/voxel, writes 12 B/run.
namespace {

constexpr int kRleThreads = 256, kRleCells = 32, kRleTile = kRleThreads * kRleCells;  // a thread owns 32 consecutive cells (four 128-bit loads)

// bit j of the result = cell (first + j) starts a run; w[] receives the 32 values, two per word (cells past the end read as the last valid one)
__device__ __forceinline__ uint32_t rle_flags(const uint16_t* __restrict__ grid, uint64_t n, uint64_t first, uint32_t w[kRleCells / 2])
{
    if (first >= n) return 0;
    uint32_t prev = first ? grid[first - 1] : (uint32_t)(uint16_t)~grid[0];
    if (first + kRleCells <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(grid + first);  // first % 32 == 0 and the grid is 16-byte aligned
#pragma unroll
        for (int k = 0; k < kRleCells / 8; ++k) {
            const uint4 v = p[k];
            w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRleCells / 2; ++j) {
            const uint64_t c = first + 2 * j;
            const uint32_t lo = c < n ? grid[c] : 0u, hi = c + 1 < n ? grid[c + 1] : 0u;
            w[j] = lo | hi << 16;
        }
    }
    uint32_t flags = 0;
#pragma unroll
    for (int j = 0; j < kRleCells; ++j) {
        const uint32_t cell = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
        if (first + j < n && cell != prev) flags |= 1u << j;
        prev = cell;
    }
    return flags;
}

__global__ void __launch_bounds__(kRleThreads) rle_count_kernel(const uint16_t* __restrict__ grid, uint64_t n, uint32_t* __restrict__ counts)
{
    __shared__ uint32_t warp_sums[kRleThreads / 32];
    uint32_t w[kRleCells / 2];
    const uint64_t first = ((uint64_t)blockIdx.x * kRleThreads + threadIdx.x) * kRleCells;
    uint32_t c = __popc(rle_flags(grid, n, first, w));
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int k = 0; k < kRleThreads / 32; ++k) t += warp_sums[k];
        counts[blockIdx.x] = t;
    }
}

// exclusive scan of the tile counts by one CTA, four counts per thread and step; *total = number of runs
__global__ void __launch_bounds__(1024) rle_scan_kernel(const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += 4096) {
        const uint32_t i = base + 4 * threadIdx.x;
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = i + k < n ? counts[i + k] : 0;
        const uint32_t mine = v[0] + v[1] + v[2] + v[3];
        uint32_t s = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if ((threadIdx.x & 31) >= o) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warp_sums[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (threadIdx.x >= o) w += t;
            }
            warp_sums[threadIdx.x] = w;
        }
        __syncthreads();
        uint32_t before = carry + (threadIdx.x >= 32 ? warp_sums[(threadIdx.x >> 5) - 1] : 0) + s - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < n) offsets[i + k] = before;
            before += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = before;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kRleThreads) rle_emit_kernel(const uint16_t* __restrict__ grid, uint64_t n, const uint32_t* __restrict__ offsets,
                                                                uint32_t* __restrict__ starts, uint16_t* __restrict__ values)
{
    __shared__ uint32_t warp_sums[kRleThreads / 32];
    uint32_t w[kRleCells / 2];
    const uint64_t first = ((uint64_t)blockIdx.x * kRleThreads + threadIdx.x) * kRleCells;
    const uint32_t flags = rle_flags(grid, n, first, w);
    const uint32_t mine = __popc(flags);
    uint32_t s = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if ((threadIdx.x & 31) >= o) s += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    uint32_t rank = offsets[blockIdx.x] + s - mine;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) rank += warp_sums[k];
    if (!flags) return;
#pragma unroll
    for (int j = 0; j < kRleCells; ++j) {  // unrolled: the values stay in registers
        if (flags >> j & 1u) {
            starts[rank] = (uint32_t)(first + j);
            values[rank] = (uint16_t)(w[j >> 1] >> ((j & 1) * 16));
            ++rank;
        }
    }
}

// records {uint16 value, uint32 repetitions}, 6 bytes each, after the 12-byte header (all 2-byte aligned)
__global__ void __launch_bounds__(256) rle_pack_kernel(const uint32_t* __restrict__ starts, const uint16_t* __restrict__ values, uint32_t runs, uint64_t n,
                                                       uint3 dims, uint16_t* __restrict__ out)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) {
        out[0] = (uint16_t)dims.x, out[1] = (uint16_t)(dims.x >> 16), out[2] = (uint16_t)dims.y;
        out[3] = (uint16_t)(dims.y >> 16), out[4] = (uint16_t)dims.z, out[5] = (uint16_t)(dims.z >> 16);
    }
    if (r >= runs) return;
    const uint32_t rep = (uint32_t)((r + 1 < runs ? (uint64_t)starts[r + 1] : n) - starts[r]);
    uint16_t* rec = out + 6 + 3 * (size_t)r;
    rec[0] = values[r], rec[1] = (uint16_t)rep, rec[2] = (uint16_t)(rep >> 16);
}

}  // namespace

// `grow` (optional) is resized to the stream's size and receives it; otherwise `out`/`cap` as in the public entry point
static vf_status rle_encode_device(vf_grid* g, std::vector<uint8_t>* grow, uint8_t* out, uint64_t cap, uint64_t* bytes_out)
{
    VF_REQUIRE(g && bytes_out, VF_ERR_INVALID_ARGUMENT, "null argument");
    vf_ctx* c = g->ctx;
    VF_TRY(vf_enter(c));
    const uint64_t n = g->n();
    VF_REQUIRE(n > 0 && n < (1ull << 32), VF_ERR_CAPACITY, "the .rle layout holds uint32 cell counts (RegularGrid.cpp:681)");
    const uint32_t ntiles = (uint32_t)((n + kRleTile - 1) / kRleTile);
    const size_t tb = ((size_t)ntiles * 4 + 255) & ~(size_t)255;
    VF_TRY(vf_scratch_reserve(c, c->codec, 2 * tb + 256));
    uint32_t* d_counts = (uint32_t*)c->codec.ptr;
    uint32_t* d_offsets = (uint32_t*)((char*)c->codec.ptr + tb);
    uint32_t* d_total = (uint32_t*)((char*)c->codec.ptr + 2 * tb);
    rle_count_kernel<<<ntiles, kRleThreads, 0, c->stream>>>(g->d, n, d_counts);
    VF_LAUNCHED(c);
    rle_scan_kernel<<<1, 1024, 0, c->stream>>>(d_counts, d_offsets, ntiles, d_total);
    VF_LAUNCHED(c);
    uint32_t* h_total = (uint32_t*)((char*)c->pinned + 65536 + 128);
    VF_CUDA(cudaMemcpyAsync(h_total, d_total, 4, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    const uint64_t runs = *h_total, bytes = 12 + 6 * runs;
    *bytes_out = bytes;
    if (grow) {
        grow->resize(bytes);
        out = grow->data(), cap = bytes;
    }
    if (!out || cap < bytes) return VF_OK;  // size query, like the host encoders
    const size_t sb = ((size_t)runs * 4 + 255) & ~(size_t)255, vb = ((size_t)runs * 2 + 255) & ~(size_t)255;
    if (c->codec.bytes < 2 * tb + 256 + sb + vb + bytes) {
        // the arena moves: keep the tile offsets (cheaper to copy than to recount)
        VfScratch bigger;  // allocated before the old arena is let go, so that no error path leaves device memory behind
        VF_TRY(vf_scratch_reserve(c, bigger, 2 * tb + 256 + sb + vb + bytes + 256));
        cudaError_t e = cudaMemcpyAsync(bigger.ptr, c->codec.ptr, 2 * tb + 256, cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = vf_sync(c);
        if (e != cudaSuccess) {
            cudaFree(bigger.ptr);
            return vf_set_error(VF_ERR_CUDA, "%s:%d growing the codec arena -> %s", __FILE__, __LINE__, cudaGetErrorString(e));
        }
        cudaFree(c->codec.ptr);
        c->codec = bigger;
        d_offsets = (uint32_t*)((char*)c->codec.ptr + tb);
    }
    uint32_t* d_starts = (uint32_t*)((char*)c->codec.ptr + 2 * tb + 256);
    uint16_t* d_values = (uint16_t*)((char*)d_starts + sb);
    uint16_t* d_out = (uint16_t*)((char*)d_values + vb);
    rle_emit_kernel<<<ntiles, kRleThreads, 0, c->stream>>>(g->d, n, d_offsets, d_starts, d_values);
    VF_LAUNCHED(c);
    rle_pack_kernel<<<(unsigned)((runs + 255) / 256), 256, 0, c->stream>>>(d_starts, d_values, (uint32_t)runs, n, make_uint3(g->X, g->Y, g->Z), d_out);
    VF_LAUNCHED(c);
    VF_CUDA(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream));
    VF_CUDA(vf_sync(c));
    return VF_OK;
}

extern "C" vf_status vf_grid_encode_rle(vf_grid* g, uint8_t* out, uint64_t cap, uint64_t* bytes_out) { return rle_encode_device(g, nullptr, out, cap, bytes_out); }

extern "C" vf_status vf_export(vf_grid* g, const char* path, int type, int squared)
{
    VF_REQUIRE(g && path, VF_ERR_INVALID_ARGUMENT, "null argument");
    VF_TRY(vf_enter(g->ctx));
    static const char* ext[4] = { "rle", "qstack", "vox", "bing" };  // FractureParameters::ExportGrid_STR, FractureParameters.h:36
    VF_REQUIRE(type >= 0 && type < 4, VF_ERR_INVALID_ARGUMENT, "bad export type %d", type);
    const uint32_t dims[3] = { g->X, g->Y, g->Z };
    std::vector<uint8_t> bytes;
    if ((type == VF_RLE || type == VF_QUADSTACK) && g->n() < (1ull << 32)) {
        // runs are found on the device; only the finished run stream is downloaded (for .qstack it is then cut into z-columns on the host)
        uint64_t need = 0;
        VF_TRY(rle_encode_device(g, &bytes, nullptr, 0, &need));
        if (type == VF_QUADSTACK) {
            std::vector<uint8_t> q(qstack_from_rle_stream(bytes.data(), bytes.size(), dims, nullptr, 0));
            VF_REQUIRE(!q.empty(), VF_ERR_CAPACITY, "exportQuadStack: dimensions must fit uint16_t (QuadStack.h:9)");
            qstack_from_rle_stream(bytes.data(), bytes.size(), dims, q.data(), q.size());
            bytes.swap(q);
        }
    } else {
        std::vector<uint16_t> host(g->n());
        VF_TRY(vf_grid_download(g, host.data()));
        if (type == VF_RLE) {
            bytes.resize(vf_encode_rle(host.data(), dims, nullptr, 0));
            vf_encode_rle(host.data(), dims, bytes.data(), bytes.size());
        } else if (type == VF_QUADSTACK) {
            bytes.resize(vf_encode_qstack(host.data(), dims, nullptr, 0));
            vf_encode_qstack(host.data(), dims, bytes.data(), bytes.size());
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
    }
    const std::string file = std::string(path) + "." + ext[type];
    FILE* f = std::fopen(file.c_str(), "wb");
    VF_REQUIRE(f != nullptr, VF_ERR_IO, "cannot open %s", file.c_str());
    const size_t w = std::fwrite(bytes.data(), 1, bytes.size(), f);
    std::fclose(f);
    VF_REQUIRE(w == bytes.size(), VF_ERR_IO, "short write to %s", file.c_str());
    return VF_OK;
}
