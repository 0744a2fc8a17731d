// standalone probe: which TMA box shapes / smem placements load correctly on this GPU (not part of the product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BX, int BY, int BZ, bool DYN>
__global__ void probe(const __grid_constant__ CUtensorMap map, uint32_t* out, int cz, int cy, int cx)
{
    extern __shared__ __align__(128) uint32_t dyn[];
    __shared__ __align__(128) uint32_t stat[DYN ? 1 : BX * BY * BZ];
    __shared__ __align__(8) unsigned long long barv;
    uint32_t* s = DYN ? dyn : stat;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&barv);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)(BX * BY * BZ * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                         (unsigned)__cvta_generic_to_shared(s)),
                     "l"(reinterpret_cast<unsigned long long>(&map)), "r"(bar), "r"(cz), "r"(cy), "r"(cx)
                     : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    for (int i = threadIdx.x; i < BX * BY * BZ; i += blockDim.x) out[i] = s[i];
}
static int g_u16 = 0, g_cz = -1;
template <int BX, int BY, int BZ, bool DYN>
void run(EncodeTiledFn enc, uint32_t* d, int X, int Y, int Z, uint32_t* out)
{
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = { (cuuint64_t)Z * (g_u16 ? 2 : 1), (cuuint64_t)Y, (cuuint64_t)X };
    const cuuint64_t strides[2] = { (cuuint64_t)Z * 4, (cuuint64_t)Z * Y * 4 };
    const cuuint32_t box[3] = { (cuuint32_t)(BZ * (g_u16 ? 2 : 1)), BY, BX }, es[3] = { 1, 1, 1 };
    CUresult r = enc(&map, g_u16 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const size_t bytes = (size_t)BX * BY * BZ * 4;
    if (DYN) cudaFuncSetAttribute(probe<BX, BY, BZ, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes + 1024);
    probe<BX, BY, BZ, DYN><<<1, 256, DYN ? bytes + 1024 : 0>>>(map, out, g_u16 ? 2 * g_cz : g_cz, -1, -1);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<uint32_t> h(BX * BY * BZ);
    cudaMemcpy(h.data(), out, bytes, cudaMemcpyDeviceToHost);
    // expect element (x,y,z) of box = d[(x-1),(y-1),(z-1)] = linear index + 1 or 0 outside
    int bad = 0;
    for (int x = 0; x < BX; ++x) for (int y = 0; y < BY; ++y) for (int z = 0; z < BZ; ++z) {
        const int gx = x - 1, gy = y - 1, gz = z + g_cz;
        const uint32_t want = (gx < 0 || gy < 0 || gz < 0 || gx >= X || gy >= Y || gz >= Z) ? 0u : (uint32_t)((gx * Y + gy) * Z + gz + 1);
        bad += h[(x * BY + y) * BZ + z] != want;
    }
    std::printf("box %dx%dx%d %s: encode=%d run=%s mismatches=%d\n", BX, BY, BZ, DYN ? "dynamic" : "static", (int)r, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) { cudaDeviceReset(); }
}
int main(int argc, char** argv)
{
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int X = 40, Y = 48, Z = 64;
    std::vector<uint32_t> h((size_t)X * Y * Z);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint32_t)i + 1;
    uint32_t *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 1 << 20);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which >= 10) g_u16 = 1, which -= 10;
    if (argc > 2) g_cz = atoi(argv[2]);
    switch (which) {
    case 0: run<10, 10, 36, false>(enc, d, X, Y, Z, out); break;
    case 1: run<10, 10, 32, false>(enc, d, X, Y, Z, out); break;
    case 2: run<10, 10, 40, false>(enc, d, X, Y, Z, out); break;
    case 3: run<18, 18, 40, true>(enc, d, X, Y, Z, out); break;
    case 4: run<18, 18, 36, true>(enc, d, X, Y, Z, out); break;
    case 5: run<10, 10, 44, false>(enc, d, X, Y, Z, out); break;
    case 6: run<10, 10, 48, false>(enc, d, X, Y, Z, out); break;
    }
    return 0;
}
