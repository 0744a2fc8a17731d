// stubs.cu — entry points of include/voxfrag.h that are not implemented yet return VF_ERR_UNSUPPORTED (never a CPU fallback).
#include "vf_internal.h"

#define VF_STUB(name, ...) \
    extern "C" vf_status name(__VA_ARGS__) { return vf_set_error(VF_ERR_UNSUPPORTED, #name ": not implemented yet"); }

VF_STUB(vf_voxelize, vf_grid*, const float*, uint32_t, const uint32_t*, uint32_t)
VF_STUB(vf_seed_uniform, vf_grid*, uint32_t, int, int, uint32_t*, uint32_t*)
VF_STUB(vf_merge_seeds, const uint32_t*, uint32_t, uint32_t*, uint32_t, int)
VF_STUB(vf_make_seeds, vf_grid*, uint32_t, uint32_t, int, int, uint32_t*, uint32_t, uint32_t*)
VF_STUB(vf_detect_boundaries, vf_grid*, int)
VF_STUB(vf_erode, vf_grid*, int, uint32_t, uint32_t, float, float, const float*, uint32_t, int)
VF_STUB(vf_remove_isolated_regions_grid, vf_grid*)
VF_STUB(vf_histogram, vf_grid*, uint32_t*, uint64_t*)
VF_STUB(vf_export, vf_grid*, const char*, int, int)
VF_STUB(vf_fracture_model, vf_grid*, const vf_params*, uint32_t*, uint32_t*, vf_flood_stats*)
extern "C" uint64_t vf_encode_rle(const uint16_t*, const uint32_t*, uint8_t*, uint64_t) { return 0; }
extern "C" uint64_t vf_encode_bing_squared(const uint16_t*, const uint32_t*, uint8_t*, uint64_t) { return 0; }
