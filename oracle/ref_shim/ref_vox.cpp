// ref_vox.cpp — drives the REFERENCE'S OWN MagicaVoxel writer (Libraries/MagicaVoxel_File_Writer/VoxWriter.cpp, compiled in
// place by the Makefile) with the AddVoxel call sequence RegularGrid::exportVox issues (RegularGrid.cpp:740-798; that file
// itself needs the GL tree, so its two loops are driven from here).  Used only by tests/test_oracle_vs_ref.py to pin the
// `.vox` byte layout of orc_encode_vox / vf_encode_vox.
#include <algorithm>
#include <cstdint>
#include <string>

#include "VoxWriter.h"

extern "C" void ref_export_vox(const uint16_t* grid, const uint32_t dims[3], int squared, const char* path)
{
    vox::VoxWriter vox;
    vox.ClearVoxels();
    vox.ClearColors();
    const int X = (int)dims[0], Y = (int)dims[1], Z = (int)dims[2];
    auto at = [&](int x, int y, int z) { return grid[((size_t)x * Y + y) * Z + z]; };
    if (squared) {
        const int M = std::max(X, std::max(Y, Z));
        const int sx = (M - X) / 2, sy = (M - Y) / 2, sz = (M - Z) / 2;
        for (int x = 0; x < M; ++x)
            for (int y = 0; y < M; ++y)
                for (int z = 0; z < M; ++z) {
                    const int cx = x - sx, cy = y - sy, cz = z - sz;
                    if (cx >= 0 && cy >= 0 && cz >= 0 && cx < X && cy < Y && cz < Z)
                        vox.AddVoxel(x, z, y, at(cx, cy, cz));
                    else
                        vox.AddVoxel(x, z, y, 0);
                }
    } else {
        for (int x = 0; x < X; ++x)
            for (int y = 0; y < Y; ++y)
                for (int z = 0; z < Z; ++z)
                    if (at(x, y, z) > 1) vox.AddVoxel(x, z, y, at(x, y, z) - 1);
    }
    vox.SaveToFile(path);
}
