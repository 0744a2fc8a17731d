// synth.cu — input synthesis for benchmarks (not part of the reference's path): fills a grid (or an x-slab of a larger grid)
// with the analytic solid "vessel of revolution" used for BASELINE config 5, where voxelizing a 20k-triangle mesh at
// 2048^3 would only measure the synthetic input, not the flood.
#include "vf_internal.h"

namespace {
__global__ void __launch_bounds__(256) solid_vessel_kernel(uint16_t* __restrict__ grid, int XS, int Y, int Z, int x_offset, int n, float base, float a1, float a2)
{
    const size_t total = (size_t)XS * Y * Z;
    const float inv = 1.0f / (float)n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(i % Z);
        const size_t r = i / Z;
        const int y = (int)(r % Y), xl = (int)(r / Y);
        const int gx = xl + x_offset;
        uint16_t v = VF_VOXEL_EMPTY;
        if (gx >= 0 && gx < n) {
            const float cx = ((float)gx + 0.5f) * inv - 0.5f, cy = ((float)y + 0.5f) * inv, cz = ((float)z + 0.5f) * inv - 0.5f;
            const float rad = sqrtf(cx * cx + cz * cz) / 0.6f;
            const float s1 = sinf(3.14159265358979f * cy);
            const float ro = base + a1 * powf(fmaxf(s1, 0.0f), 0.8f) + a2 * sinf(3.0f * 3.14159265358979f * cy);
            if (rad <= ro && (rad >= ro - 0.06f || cy < 0.08f)) v = VF_VOXEL_FREE;
        }
        grid[i] = v;
    }
}
}  // namespace

// grid dims = (XS, n, n) where plane xl corresponds to global plane xl + x_offset of an n^3 grid (x_offset may be -1 for a
// slab's lower halo plane; planes outside [0, n) are EMPTY).
extern "C" vf_status vf_synth_solid_vessel(vf_grid* g, int x_offset, uint32_t n, float base, float a1, float a2)
{
    VF_REQUIRE(g != nullptr, VF_ERR_INVALID_ARGUMENT, "null grid");
    VF_TRY(vf_enter(g->ctx));
    solid_vessel_kernel<<<g->ctx->num_sms * 8, 256, 0, g->ctx->stream>>>(g->d, (int)g->X, (int)g->Y, (int)g->Z, x_offset, (int)n, base, a1, a2);
    VF_LAUNCHED(g->ctx);
    return VF_OK;
}
