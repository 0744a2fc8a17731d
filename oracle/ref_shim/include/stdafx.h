// stdafx.h shim — stands in for SRC/PrecompiledHeaders/stdafx.h (which needs <windows.h>, GLEW, GLFW, glm, lodepng, CGAL) when the
// reference's portable hot-path sources are compiled in place as a checker for the oracle.  TEST TOOLING ONLY.
#pragma once
#define GENERATE_DATASET false
#define TESTING_FORMAT_MODE false
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include "glm/mini_glm.hpp"
typedef unsigned int GLuint;
typedef int GLint;
typedef unsigned int GLenum;
typedef float GLfloat;
#define GL_DYNAMIC_DRAW 0x88E8
#define GL_STATIC_DRAW 0x88E4
#define GL_COMPUTE_SHADER 0x91B9
#include "Geometry/General/Adapter.h"  // the reference's own typedefs (vec3 = glm::vec3, ...)
#include "boost/random.hpp"  // Seeder.h names boost types before Seeder.cpp includes boost
namespace std {
inline float fabsf(float x) { return ::fabsf(x); }  // MSVC exposes std::fabsf; libstdc++ only ::fabsf / std::fabs
}
