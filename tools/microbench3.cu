// tools/microbench3.cu -- what can issue next to FFMA2 (fma.rn.f32x2)?  Not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/microbench3.cu -o tools/microbench3
// Each kernel runs a loop of 16 independent FFMA2 chains per thread, alone or interleaved one-to-one with an independent
// instruction of another class; prints cycles per loop body slot (SM cycles * 4 sub-partitions / warp instructions).
// If "FFMA2 + X" takes the same cycles per FFMA2 as FFMA2 alone (2.0), X issues in the cycle the FMA pipe is still busy.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__constant__ unsigned c_tab[1024];
template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, u64* out, u64 seed) {
    u64 a[8];
    unsigned b[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = (unsigned)(seed >> 3) + i * 13 + threadIdx.x; d[i] = 1.0 + i + threadIdx.x; }
    const u64 m = 0x3f8000003f800000ull ^ (seed & 1), c = 0x3a8000003a800000ull;
    const unsigned bm = (unsigned)seed | 1u;
    const double dm = 1.0000001;
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE != 9 && MODE < 10) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
                if (MODE == 10) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(bm), "r"(b[(i + 1) & 7]));    // LOP3 alone
                if (MODE == 11) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(bm));                                 // IMAD alone
                if (MODE == 12) asm volatile("max.f32 %0, %0, %1;" : "+r"(b[i]) : "r"(b[(i + 1) & 7]));                          // FMNMX alone
                if (MODE == 13) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(bm));                                 // FFMA alone
                if (MODE == 14) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dm));                               // DFMA alone
                if (MODE == 8) { unsigned v; asm volatile("ld.const.u32 %0, [%1];" : "=r"(v) : "l"(c_tab + ((it * 32 + u * 8 + i) & 1023))); b[i] ^= v; }   // FFMA2 + LDCU (+ LOP3 to consume)
                if (MODE == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(bm), "r"(b[(i + 1) & 7]));     // ALU
                if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(dm));                                // FP64
                if (MODE == 3) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(bm));                                  // IMAD (FMA pipe)
                if (MODE == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b[i]) : "r"((unsigned)__cvta_generic_to_shared(sh + ((b[i] + i) & 255))));   // LSU
                if (MODE == 5) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(bm), "r"(b[(i + 1) & 7]));
                                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x69;" : "+r"(b[(i + 3) & 7]) : "r"(bm), "r"(b[(i + 5) & 7])); }   // 2 ALU per FFMA2
                if (MODE == 6) asm volatile("max.f32 %0, %0, %1;" : "+r"(b[i]) : "r"(b[(i + 1) & 7]));                           // FMNMX (ALU)
                if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(bm));                                  // scalar FFMA next to FFMA2
                if (MODE == 9) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+r"(b[i]) : "r"(bm));
                                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[(i + 3) & 7]) : "r"(bm), "r"(b[(i + 5) & 7])); }   // scalar FFMA + ALU
            }
    }
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i] + (u64)d[i];
    if (s == 0x123456789ull) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_ffma2_extra, int sms) {
    const int iters = 4000;
    u64* out; CK(cudaMalloc(&out, 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<sms * 8, 256>>>(100, out, 12345);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        k<MODE><<<sms * 8, 256>>>(iters, out, 12345);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    // warps per sub-partition: 8 blocks * 8 warps / 4 = 16; main instructions per warp: iters * 32
    const double cyc = best * 1e-3 * 1.965e9;   // SM cycles (assumes 1965 MHz)
    const double main_per_smsp = 16.0 * iters * 32;
    printf("%-34s %8.3f ms  %.3f cycles per loop slot (FFMA2%s)\n", name, best, cyc / main_per_smsp, per_ffma2_extra ? " + extra" : "");
    cudaFree(out);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA2 alone", 0, sms);
    run<1>("FFMA2 + LOP3 (ALU)", 1, sms);
    run<6>("FFMA2 + FMNMX (ALU)", 1, sms);
    run<5>("FFMA2 + 2 LOP3 (ALU)", 2, sms);
    run<2>("FFMA2 + DFMA (FP64)", 1, sms);
    run<3>("FFMA2 + IMAD (FMA pipe)", 1, sms);
    run<7>("FFMA2 + FFMA (FMA pipe)", 1, sms);
    run<4>("FFMA2 + LDS (LSU)", 1, sms);
    run<9>("FFMA + LOP3 (no FFMA2)", 1, sms);
    run<8>("FFMA2 + LDCU + LOP3", 2, sms);
    run<10>("LOP3 alone", 0, sms);
    run<11>("IMAD alone", 0, sms);
    run<12>("FMNMX alone", 0, sms);
    run<13>("FFMA alone", 0, sms);
    run<14>("DFMA alone", 0, sms);
    return 0;
}
