// issue_order.cu -- microbenchmark: does the ORDER in which a warp's LOP3 (ALU pipe) and IMAD (FMA pipe)
// instructions appear matter for how well the two pipes overlap at the occupancy of the sliced kernel?
// Every instruction reads registers written >= 8 instructions earlier (no RAW stalls at 3 warps / SMSP); the
// IMAD multiplier is a uniform register as in the product kernel.  A "group" is one pass of the pattern string
// ('L' = lop3, 'I' = imad with uniform multiplier, 'J' = imad with register multiplier, 'F' = ffma, 'A' = iadd3).
// Exploration tool, not product.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xE0;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t ffma(uint32_t a, uint32_t b, uint32_t c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
    return __float_as_uint(d);
}
__device__ __forceinline__ uint32_t iadd3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ uint32_t shfw(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) {
    uint32_t d;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
struct Pat { char s[24]; };
template <int N> constexpr Pat mk(const char (&x)[N]) { Pat p{}; for (int i = 0; i < N; ++i) p.s[i] = x[i]; return p; }
constexpr int plen(const Pat &p) { int n = 0; while (p.s[n]) ++n; return n; }

template <int ID> struct P;
#define DEFPAT(id, str) template <> struct P<id> { static constexpr Pat pat = mk(str); };
DEFPAT(0, "LLLLL")
DEFPAT(1, "LLLL")
DEFPAT(2, "IIII")
DEFPAT(3, "LLLLIII")
DEFPAT(4, "LILILIL")
DEFPAT(5, "LLLLII")
DEFPAT(6, "LILILL")
DEFPAT(7, "LLILLI")
DEFPAT(8, "LI")
DEFPAT(9, "LLI")
DEFPAT(10, "LLLLJJJ")
DEFPAT(11, "LJLJLJL")
DEFPAT(12, "LLLLFFF")
DEFPAT(13, "LFLFLFL")
DEFPAT(14, "LF")
DEFPAT(15, "FFFF")
DEFPAT(16, "LLLLFFFF")
DEFPAT(17, "LFLFLFLF")
DEFPAT(18, "LLLLIIII")
DEFPAT(19, "LILILILI")
DEFPAT(20, "LLLLLLLLIIIIII")
DEFPAT(21, "LLLLLLLLIIII")
DEFPAT(22, "LLLLI")
DEFPAT(23, "IF")
DEFPAT(24, "LIF")
DEFPAT(25, "LLIF")
DEFPAT(26, "SSSS")
DEFPAT(27, "HHHH")
DEFPAT(28, "SH")
DEFPAT(29, "SSH")
DEFPAT(30, "SSSSH")
DEFPAT(31, "SSSSHI")
DEFPAT(32, "LLLLH")

// O1, O2: register-index offsets of the 2nd / 3rd source; LD = distinct source registers of a LOP3 (1..3);
// IDR = distinct source registers of an IMAD (1..2, plus the uniform multiplier)
template <int ID, int WPS, int O1 = 13, int O2 = 26, int LD = 3, int IDR = 2>
__global__ void __launch_bounds__(128, WPS) order_kernel(uint32_t *out, int iters, uint32_t mul_u, uint32_t seed) {
    constexpr Pat pat = P<ID>::pat;
    constexpr int L = plen(pat);
    constexpr int NR = 40;
    constexpr int REP = NR;  // NR * L instructions per loop trip: every trip sees the same register pattern
    uint32_t r[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) r[i] = seed * (i + 1) + threadIdx.x;
    const uint32_t mulv = mul_u | (threadIdx.x >> 20);
    const uint32_t fx = seed ^ (threadIdx.x * 77u), fy = seed + threadIdx.x * 1234567u;  // fixed source registers
    int n = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int g = 0; g < REP; ++g) {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int idx = g * L + l;
                const int d = idx % NR;  // d last written NR instrs ago
                const int s1 = LD >= 2 ? (idx + O1) % NR : d, s2 = LD >= 3 ? (idx + O2) % NR : s1;
                (void)s2;
                const int i1 = IDR >= 2 ? (idx + O1) % NR : d;
                const char c = pat.s[l];
                if (c == 'L' && LD == 4) r[d] = lop3(fx, fy, r[d]);
                else if (c == 'L' && LD == 5) r[d] = lop3(fx, r[d], r[s1]);
                else if (c == 'L') r[d] = lop3(r[d], r[s1], r[s2]);
                else if (c == 'I') r[d] = imad(r[i1], mul_u, r[d]);
                else if (c == 'J') r[d] = imad(r[i1], mulv, r[d]);
                else if (c == 'F') r[d] = ffma(r[s1], r[s2], r[d]);
                else if (c == 'A') r[d] = iadd3(r[d], r[s1], r[s2]);
                else if (c == 'S') r[d] = shfw(r[d], r[s1], r[s1]);
                else if (c == 'H') r[d] = mulhi(r[d], mul_u);
            }
        }
        ++n;
    }
    uint32_t x = n;
#pragma unroll
    for (int i = 0; i < NR; ++i) x ^= r[i];
    if (x == 0x12345678u) out[0] = x;
}

template <int ID, int WPS, int O1 = 13, int O2 = 26, int LD = 3, int IDR = 2>
void run() {
    constexpr Pat pat = P<ID>::pat;
    constexpr int L = plen(pat);
    constexpr int REP = 40;
    uint32_t *d;
    cudaMalloc(&d, 4);
    const int iters = 3000;
    auto k = order_kernel<ID, WPS, O1, O2, LD, IDR>;
    k<<<148 * WPS, 128>>>(d, 100, 0xFFFFFFFFu, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * WPS, 128>>>(d, iters, 0xFFFFFFFFu, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int nl = 0, ni = 0;
    for (int i = 0; i < L; ++i) { if (pat.s[i] == 'L' || pat.s[i] == 'A' || pat.s[i] == 'S') ++nl; else ++ni; }
    const double groups = (double)iters * REP * WPS;  // per SMSP
    const double clk = best * 1e-3 * 1.965e9;
    printf("o1=%2d o2=%2d LD=%d ID=%d  ", O1, O2, LD, IDR);
    printf("%-16s warps/SMSP=%d  %.3f ms  clk/group/SMSP=%6.2f  (ALU-bound %4.1f, FMA-bound %4.1f, issue-bound %4.1f)  clk/instr=%.3f\n",
           pat.s, WPS, best, clk / groups, 2.0 * nl, 2.0 * ni, (double)L, clk / groups / L);
    cudaFree(d);
}

template <int ID> void run_all() { run<ID, 3>(); }
template <int ID> void sweep() { run<ID, 3, 13, 26, 2, 2>(); run<ID, 8, 13, 26, 2, 2>(); }
int main() {
    sweep<26>(); sweep<27>(); sweep<28>(); sweep<29>(); sweep<30>(); sweep<31>(); sweep<32>(); sweep<2>();
    return 0;
}
