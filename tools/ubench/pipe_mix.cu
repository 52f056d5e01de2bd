// pipe_mix.cu -- microbenchmark: how many LOP3 (ALU pipe) + IMAD (FMA pipe) instructions per clock an SM
// sustains at the occupancy of the sliced kernel (12 warps / SM) when every instruction reads DISTINCT registers
// (no operand reuse), for several LOP3 : IMAD mixes and IMAD operand forms.  Exploration tool, not product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// MODE: 0 = LOP3 only; 1 = NL LOP3 + NI IMAD(mult uniform); 2 = same with the multiplier in a vector register
// SHARE: the LOP3s of a group read the same two registers in the same operand slots (operand reuse cache)
template <int NL, int NI, int MODE, int SHARE = 0>
__global__ void __launch_bounds__(128, 3) mix_kernel(uint32_t *out, int iters, uint32_t mul_u, uint32_t seed) {
    constexpr int NR = 48;
    uint32_t r[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = seed * (i + 1) + threadIdx.x;
    uint32_t mul = mul_u;
    if (MODE == 2) mul = mul_u | (threadIdx.x >> 20);  // thread dependent as far as the compiler knows
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < 16; ++g) {  // 16 "cells" per iteration
            const int b = g * 3;
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const int d = (b + l) % NR;
                if (SHARE == 0) r[d] = lop3<0xE0 + 0>(r[d], r[(d + 17) % NR], r[(d + 29) % NR]);
                else if (SHARE == 1) r[d] = lop3<0xE0 + 0>(r[d], r[(b + 17) % NR], r[(b + 29) % NR]);
                else r[d] = lop3<0xE0 + 0>(r[d], r[(d + 17) % NR], r[(b + 29) % NR]);
            }
#pragma unroll
            for (int l = 0; l < NI; ++l) {
                const int d = (b + 5 + l) % NR;
                r[d] = imad(r[(d + 11) % NR], mul, r[d]);
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < NR; ++i) x ^= r[i];
    if (x == 0x12345678u) out[0] = x;
}

template <int NL, int NI, int MODE, int SHARE = 0>
void run(const char *name) {
    uint32_t *d;
    cudaMalloc(&d, 4);
    const int iters = 20000;
    auto k = mix_kernel<NL, NI, MODE, SHARE>;
    k<<<148 * 3, 128>>>(d, 100, 0xFFFFFFFFu, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<148 * 3, 128>>>(d, iters, 0xFFFFFFFFu, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SMSP: 3 warps, each executes iters*16 groups
    const double groups = (double)iters * 16 * 3;  // per SMSP
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s NL=%d NI=%d  %.3f ms  clk/group/SMSP=%.2f  (ALU-bound %.1f, issue-bound %.1f)\n", name, NL, NI, ms,
           clk / groups, 2.0 * NL, (double)(NL + NI));
    cudaFree(d);
}

int main() {
    run<5, 0, 0>("lop3 only");
    run<4, 0, 0>("lop3 only");
    run<4, 3, 1>("lop3 + imad(uniform mult)");
    run<4, 3, 2>("lop3 + imad(vector mult)");
    run<4, 2, 1>("lop3 + imad(uniform mult)");
    run<4, 2, 2>("lop3 + imad(vector mult)");
    run<4, 1, 1>("lop3 + imad(uniform mult)");
    run<4, 4, 1>("lop3 + imad(uniform mult)");
    run<4, 4, 2>("lop3 + imad(vector mult)");
    run<5, 0, 0, 1>("lop3 only, 2 shared operands");
    run<4, 3, 1, 1>("lop3(2 shared) + imad(uniform)");
    run<4, 2, 1, 1>("lop3(2 shared) + imad(uniform)");
    run<4, 4, 1, 1>("lop3(2 shared) + imad(uniform)");
    run<5, 0, 0, 2>("lop3 only, 1 shared operand");
    run<4, 3, 1, 2>("lop3(1 shared) + imad(uniform)");
    run<4, 2, 1, 2>("lop3(1 shared) + imad(uniform)");
    run<0, 4, 1>("imad only (uniform mult)");
    run<0, 4, 2>("imad only (vector mult)");
    return 0;
}
