#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int x, int y) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(256 * 96) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem)),
                     "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&mbar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
    }
    for (int i = threadIdx.x; i < 256 * 96; i += blockDim.x) out[i] = smem[i];
}
int main() {
    const int pitch = 1280, h = 376, S = 2;
    std::vector<uint8_t> img((size_t)pitch * h * S);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 7 + i / pitch);
    uint8_t *d, *o;
    cudaMalloc(&d, img.size() + 4096); cudaMalloc(&o, 256 * 96);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    printf("entry %d\n", (int)cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)h * S}; cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {256, 96}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    const int xs[7] = {0, 48, 1104, 13, 1107, 3, 1030}, ys[7] = {0, 381, 700, 7, 690, 375, 1};
    for (int t = 0; t < 7; ++t) {
        int x = xs[t], y = ys[t];
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 96 + 1024);
        k<<<1, 256, 256 * 96 + 1024>>>(tm, o, x, y);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> got(256 * 96);
        cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < 96; ++r2) for (int c = 0; c < 256; ++c) {
            uint8_t exp = (x + c < pitch && y + r2 < h * S) ? img[(size_t)(y + r2) * pitch + x + c] : 0;
            bad += got[r2 * 256 + c] != exp;
        }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
