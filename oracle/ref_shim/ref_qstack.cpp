// ref_qstack.cpp — drives the REFERENCE'S OWN QuadStack<uint16_t> (SRC/DataStructures/QuadStack.h + GStack.h, compiled in place) with
// the call sequence of RegularGrid::exportQuadStack (RegularGrid.cpp:716-725).  GStack.h declares a nested template whose parameter
// shadows the outer one (accepted by MSVC, rejected by g++), so the Makefile writes a copy with that one parameter renamed into the
// git-ignored oracle/_ref/DataStructures/ at build time; nothing of the reference is stored in the repository.  TEST TOOLING ONLY.
#include "stdafx.h"
namespace lodepng {  // QuadStack::backgroundWriteImage names lodepng; it is never called on the export path
inline unsigned encode(std::vector<unsigned char>&, std::vector<unsigned char>&, unsigned, unsigned, int) { return 1; }
inline void save_file(std::vector<unsigned char>&, const std::string&) {}
}  // namespace lodepng
enum LodePNGColorType { LCT_GREY = 0 };
#include "DataStructures/QuadStack.h"

extern "C" int ref_export_qstack(uint16_t* grid, const uint32_t dims[3], const char* filename)
{
    RegularGrid g(grid, uvec3(dims[0], dims[1], dims[2]));
    QuadStack<uint16_t>* quadStack = new QuadStack<uint16_t>();
    if (!quadStack->loadCube(&g)) {
        delete quadStack;
        return -1;
    }
    quadStack->compress_y();
    quadStack->compress_x();
    quadStack->saveCheckpoint(filename);
    delete quadStack;
    return 0;
}
