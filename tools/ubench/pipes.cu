// pipes.cu -- issue-rate microbenchmark for the integer ops the LK / FAST kernels are built from (B200, sm_100a).
// Each kernel runs ILP-8 chains of one op; prints warp-instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 4096
template <int OP> __global__ void k(uint32_t* out, uint32_t a, uint32_t b, long long* cyc) {
    uint32_t r[8];
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * a;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * (i + 1) + a;
    long long t0 = clock64();
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) r[i] = r[i] * a + b;                                  // IMAD
            if (OP == 1) r[i] = __dp2a_lo(r[i], b, r[i]);                      // IDP.2A
            if (OP == 2) r[i] = __dp4a(r[i], b, r[i]);                         // IDP.4A
            if (OP == 3) r[i] = (r[i] + b) ^ a;                                // LOP3/IADD3 mix (alu)
            if (OP == 4) r[i] = __shfl_down_sync(0xFFFFFFFFu, r[i], 1);        // SHFL
            if (OP == 5) r[i] = sm[(r[i] & 2047)];                             // LDS.32 (dependent)
            if (OP == 6) r[i] = __vabsdiffu4(r[i], b);                         // VABSDIFF4
            if (OP == 7) r[i] = __vimin3_s16x2(r[i], a, b);                    // VIMNMX3
            if (OP == 8) r[i] = __funnelshift_r(r[i], b, 8);                   // SHF
            if (OP == 9) r[i] = __byte_perm(r[i], b, 0x5140);                  // PRMT
            if (OP == 10) { r[i] = r[i] * a + b; r[i] = (r[i] >> 3) ^ b; }     // IMAD + alu pair
            if (OP == 11) r[i] = __popc(r[i]) + b;                             // POPC
            if (OP == 12) r[i] = *((volatile uint16_t*)((uint8_t*)sm + ((r[i] & 2047) << 1)));  // LDS.U16
            if (OP == 13) r[i] = __reduce_add_sync(0xFFFFFFFFu, r[i]);         // REDUX
            if (OP == 14) r[i] = __umulhi(r[i], a) + b;                        // IMAD.HI
            if (OP == 15) r[i] = (r[i] & a) | b;                               // LOP3 only
            if (OP == 16) r[i] = r[i] + a + b;                                 // IADD3 only
            if (OP == 17) { r[i] = r[i] * a + b; r[i] = __vimin3_s16x2(r[i], a, b); }   // IMAD + VIMNMX3 (two pipes?)
            if (OP == 18) { r[i] = r[i] * a + b; r[i] = (r[i] & a) | b; }      // IMAD + LOP3
            if (OP == 19) { r[i] = __vimin3_s16x2(r[i], a, b); r[i] = (r[i] & a) | b; } // VIMNMX3 + LOP3
            if (OP == 20) { r[i] = __vabsdiffu4(r[i], b); r[i] = __funnelshift_r(r[i], b, 8); }  // VABSDIFF4 + SHF (same pipe?)
            if (OP == 21) { uint4 v = *reinterpret_cast<const uint4*>(sm + ((threadIdx.x * 4 + i * 128 + it) & 2044)); r[i] += v.x ^ v.y ^ v.z ^ v.w; }  // LDS.128 independent + 4 alu
            if (OP == 22) r[i] += *((volatile uint8_t*)sm + ((threadIdx.x + i * 37 + it) & 8191));  // LDS.U8 independent + IADD
            if (OP == 23) r[i] += atomicAdd(&sm[(threadIdx.x >> 5) * 8 + i], 1u);  // ATOMS.ADD with return, one address per warp
            if (OP == 24) atomicOr(&sm[(threadIdx.x + i) & 2047], r[i] | it);     // ATOMS.OR without return, distinct addresses
            if (OP == 25) r[i] = __ffs(r[i]) + b;                              // BREV + FLO
            if (OP == 26) r[i] = __ballot_sync(0xFFFFFFFFu, r[i] & 1) + r[i];   // VOTE + IADD
            if (OP == 27) { sm[(threadIdx.x + i * 32 + it) & 2047] = r[i]; r[i] += b; }  // STS.32 + IADD
            if (OP == 28) r[i] = __umulhi(r[i], a);                             // IMAD.HI no addend
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name, int per) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 4 * 512 * 4); cudaMalloc(&cyc, 8);
    for (int warps = 4; warps <= 16; warps *= 2) {   // warps per SM (1, 2, 4 per sub-partition)
        k<OP><<<148, warps * 32>>>(out, 3, 0x01020304u, cyc);
        k<OP><<<148, warps * 32>>>(out, 3, 0x01020304u, cyc);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double ipc = (double)N_IT * 8 * per * (warps / 4) / (double)h;
        printf("%-14s warps/SMSP=%d  cycles=%lld  warp-instr/cycle/SMSP=%.3f\n", name, warps / 4, h, ipc);
    }
}
int main() {
    run<0>("IMAD", 1); run<1>("IDP.2A", 1); run<2>("IDP.4A", 1); run<3>("IADD+LOP3", 2); run<4>("SHFL", 1); run<5>("LDS.32", 2);
    run<6>("VABSDIFF4", 1); run<7>("VIMNMX3.16x2", 1); run<8>("SHF", 1); run<9>("PRMT", 1); run<10>("IMAD+SHF+LOP", 3); run<11>("POPC+IADD", 2);
    run<12>("LDS.U16", 3); run<13>("REDUX", 1);
    run<14>("IMAD.HI+c", 1); run<15>("LOP3", 1); run<16>("IADD3", 1); run<17>("IMAD+VIMNMX3", 2); run<18>("IMAD+LOP3", 2); run<19>("VIMNMX3+LOP3", 2);
    run<20>("VABSDIFF4+SHF", 2); run<21>("LDS.128+4alu", 1); run<22>("LDS.U8+IADD", 1); run<23>("ATOMS.ADD ret", 1); run<24>("ATOMS.OR noret", 1);
    run<25>("FFS", 1); run<26>("VOTE+IADD", 1); run<27>("STS.32+IADD", 1); run<28>("IMAD.HI", 1);
    return 0;
}
