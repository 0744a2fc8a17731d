// RegularGrid shim: the host-side data model of SRC/DataStructures/RegularGrid.h (CellGrid, index rule, isOccupied / isBoundary /
// at / set / swap) without the GL, marching-cubes and export machinery its .cpp pulls in.  Lets NaiveFracturer.cpp and Seeder.cpp
// compile unmodified.  The bodies restate RegularGrid.cpp:523-526, 543-564, 566-569, 831-842.  TEST TOOLING ONLY.
#pragma once
#include "stdafx.h"
#include "Utilities/RandomUtilities.h"  // reaches Seeder.cpp through RegularGrid.h in the reference too
#define VOXEL_EMPTY 0
#define VOXEL_FREE 1
class RegularGrid {
public:
    struct CellGrid {
        uint16_t _value;
        CellGrid() : _value(VOXEL_EMPTY) {}
        CellGrid(uint16_t v) : _value(v) {}
    };
    RegularGrid(uint16_t* cells, const uvec3& dims) : _cells(reinterpret_cast<CellGrid*>(cells)), _numDivs(dims) {}
    CellGrid* data() { return _cells; }
    uvec3 getNumSubdivisions() const { return _numDivs; }
    GLuint ssbo() const { return 0; }
    uint16_t at(int x, int y, int z) const { return _cells[getPositionIndex(x, y, z, _numDivs)]._value; }
    void set(int x, int y, int z, uint16_t i) { _cells[getPositionIndex(x, y, z, _numDivs)]._value = i; }
    bool isOccupied(int x, int y, int z) const { return at(x, y, z) != VOXEL_EMPTY; }
    bool isBoundary(int x, int y, int z, int neighbourhoodSize = 1) const
    {
        if (neighbourhoodSize % 2 == 0) ++neighbourhoodSize;
        ivec3 mn(glm::clamp(x - neighbourhoodSize, 0, int(_numDivs.x) - 1), glm::clamp(y - neighbourhoodSize, 0, int(_numDivs.y) - 1), glm::clamp(z - neighbourhoodSize, 0, int(_numDivs.z) - 1));
        ivec3 mx(glm::clamp(x + neighbourhoodSize, 0, int(_numDivs.x) - 1), glm::clamp(y + neighbourhoodSize, 0, int(_numDivs.y) - 1), glm::clamp(z + neighbourhoodSize, 0, int(_numDivs.z) - 1));
        for (int a = mn.x; a <= mx.x; ++a)
            for (int b = mn.y; b <= mx.y; ++b)
                for (int c = mn.z; c <= mx.z; ++c)
                    if (at(a, b, c) == VOXEL_EMPTY) return true;
        return false;
    }
    void swap(CellGrid* src, size_t n) { std::copy(src, src + n, _cells); }
    static unsigned getPositionIndex(int x, int y, int z, const uvec3& numDivs) { return x * numDivs.y * numDivs.z + y * numDivs.z + z; }
    // RegularGrid.h:370-391 (QuadStack::loadCube reads the grid through this): vectors only ever grow
    template <typename T>
    void getData(std::vector<std::vector<std::vector<T>>>& data)
    {
        if (data.size() < _numDivs.x) data.resize(_numDivs.x);
        for (unsigned x = 0; x < _numDivs.x; ++x) {
            if (data[x].size() < _numDivs.y) data[x].resize(_numDivs.y);
            for (unsigned y = 0; y < _numDivs.y; ++y) {
                if (data[x][y].size() < _numDivs.z) data[x][y].resize(_numDivs.z);
                for (unsigned z = 0; z < _numDivs.z; ++z) data[x][y][z] = static_cast<T>(at(x, y, z));
            }
        }
    }

private:
    CellGrid* _cells;
    uvec3 _numDivs;
};
