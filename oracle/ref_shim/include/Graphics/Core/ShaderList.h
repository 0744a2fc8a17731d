// ShaderList / ComputeShader shim: inert stand-ins so that the GPU halves of NaiveFracturer.cpp (buildGPU,
// removeIsolatedRegionsGPU) compile; they are never executed by the checker, which drives buildCPU / removeIsolatedRegionsCPU.
#pragma once
#include "stdafx.h"
namespace RendEnum { enum CompShaderTypes { NAIVE_FRACTURER, REMOVE_ISOLATED_REGIONS }; }
class ComputeShader {
public:
    template <class... A> static GLuint setWriteBuffer(A&&...) { return 0; }
    template <class... A> static GLuint setReadBuffer(A&&...) { return 0; }
    template <class... A> static void updateReadBuffer(A&&...) {}
    template <class... A> static void updateReadBufferSubset(A&&...) {}
    template <class T> static T* readData(GLuint, const T&) { return nullptr; }
    static unsigned getNumGroups(unsigned n) { return (n + 1023) / 1024; }
    static unsigned getMaxGroupSize() { return 1024; }
    static void deleteBuffer(GLuint) {}
    static void deleteBuffers(const std::vector<GLuint>&) {}
    void bindBuffers(const std::vector<GLuint>&) {}
    void use() {}
    template <class... A> void setUniform(A&&...) {}
    template <class... A> void setSubroutineUniform(A&&...) {}
    void applyActiveSubroutines() {}
    template <class... A> void execute(A&&...) {}
};
class ShaderList {
public:
    static ShaderList* getInstance() { static ShaderList s; return &s; }
    ComputeShader* getComputeShader(int) { static ComputeShader c; return &c; }
};
