// imma.cu -- issue rate of the legacy int8 tensor-core instruction (mma.sync.m16n8k32.u8.u8.s32 -> IMMA.16832.U8.U8) on B200,
// per SM sub-partition, with 1 / 2 / 4 warps per sub-partition and 4 independent accumulator sets per warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 4096
__global__ void k(int* out, uint32_t a, uint32_t b, long long* cyc) {
    int c[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x + i + j;
    uint32_t a0 = a + threadIdx.x, a1 = a * 3, a2 = a ^ 5, a3 = a + 7, b0 = b + threadIdx.x, b1 = b * 5;
    long long t0 = clock64();
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    int* out; long long* cyc; cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 16; warps *= 2) {
        k<<<148, warps * 32>>>(out, 3, 0x01010101u, cyc);
        k<<<148, warps * 32>>>(out, 3, 0x01010101u, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_cycle = (double)N_IT * 4 * (warps / 4) / (double)h;
        printf("IMMA.16832.U8 warps/SMSP=%d cycles=%lld IMMA/cycle/SMSP=%.4f -> %.1f int8 TOPS per GPU at 1.965 GHz\n", warps / 4, h, per_cycle,
               per_cycle * 16 * 8 * 32 * 2 * 148 * 4 * 1.965e9 / 1e12);
    }
    return 0;
}
