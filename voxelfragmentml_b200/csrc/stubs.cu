// stubs.cu — entry points of include/voxfrag.h that are not implemented yet return VF_ERR_UNSUPPORTED (never a CPU fallback).
#include "vf_internal.h"

#define VF_STUB(name, ...) \
    extern "C" vf_status name(__VA_ARGS__) { return vf_set_error(VF_ERR_UNSUPPORTED, #name ": not implemented yet"); }

VF_STUB(vf_voxelize, vf_grid*, const float*, uint32_t, const uint32_t*, uint32_t)
