// tools/microbench.cu -- design probes for the cull scan (not part of the product library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false tools/microbench.cu -o tools/microbench
// Prints, for several scan-loop shapes, Gtests/s and issue slots per test at the measured SM clock,
// plus FFMA peak with register vs constant operands.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct CullRay { float dx, dy, dz, ndo, mx, my, mz, o2; };

__device__ __forceinline__ float cull_D(const CullRay& f, const float4 s) {
    float b = fmaf(f.dz, s.z, f.ndo);
    b = fmaf(f.dy, s.y, b);
    b = fmaf(f.dx, s.x, b);
    float q = s.w + f.o2;
    q = fmaf(f.mx, s.x, q);
    q = fmaf(f.my, s.y, q);
    q = fmaf(f.mz, s.z, q);
    return fmaf(b, b, -q);
}

__constant__ float4 c_filt[4096];

__device__ __forceinline__ CullRay make_ray(int seed) {
    CullRay f;
    const float a = 0.001f * (float)(seed % 1000), b = 0.002f * (float)(seed % 777);
    f.dx = __sinf(a) * __cosf(b); f.dy = __cosf(a); f.dz = __sinf(a) * __sinf(b);
    const float ox = 13.f + a, oy = 2.f + b, oz = 3.f + a * b;
    f.ndo = -(f.dx * ox + f.dy * oy + f.dz * oz);
    f.mx = -2.f * ox; f.my = -2.f * oy; f.mz = -2.f * oz;
    f.o2 = ox * ox + oy * oy + oz * oz;
    return f;
}

// SRC: 0 = shared memory (LDS.128), 1 = __constant__ (uniform loads), 2 = global (LDG, L1-resident)
template <int R, int U, int SRC>
__global__ void __launch_bounds__(256) scan_kernel(const float4* __restrict__ g_filt, int npad, int reps, unsigned* out) {
    extern __shared__ float4 s_filt[];
    if (SRC == 0) {
        for (int i = threadIdx.x; i < npad; i += blockDim.x) s_filt[i] = g_filt[i];
        __syncthreads();
    }
    CullRay f[R];
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = make_ray(threadIdx.x * 7 + blockIdx.x * 13 + r * 101);
    unsigned hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int k = 0; k < npad; k += U) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = SRC == 0 ? s_filt[k + u] : (SRC == 1 ? c_filt[k + u] : __ldg(g_filt + k + u));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float D[U];
                unsigned all_neg = 0x80000000u;
#pragma unroll
                for (int u = 0; u < U; ++u) { D[u] = cull_D(f[r], s[u]); all_neg &= __float_as_uint(D[u]); }
                if ((int)all_neg >= 0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (!(D[u] < 0.f)) hits += k + u;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) f[r].o2 += 1e-3f;  // keep reps from being hoisted
    }
    if (hits == 0xdeadbeef) out[0] = hits;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = hits;
}

// FFMA peak: MODE 0 = constant-bank operands, 1 = three distinct register operands, 2 = regs w/ shared multiplier
template <int MODE>
__global__ void __launch_bounds__(256) ffma_kernel(float* out, const float* in, int iters, float ca, float cb) {
    float x[8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; a[i] = in[(threadIdx.x + i) & 63]; b[i] = in[(threadIdx.x + 2 * i + 1) & 63]; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fmaf(x[i], ca, cb);
                else if (MODE == 1) x[i] = fmaf(x[i], a[i], b[i]);
                else x[i] = fmaf(x[i], a[0], b[i]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <typename F>
float time_ms(F launch, int n = 3) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int i = 0; i < n; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i > 0 && ms < best) best = ms;
    }
    return best;
}

template <int R, int U, int SRC>
void run_scan(const char* name, const float4* d_filt, int npad, unsigned* d_out, int sms, int blocks_per_sm) {
    const int reps = 400;
    const int grid = sms * blocks_per_sm;
    const size_t smem = SRC == 0 ? (size_t)npad * 16 : 0;
    CK(cudaFuncSetAttribute(scan_kernel<R, U, SRC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_kernel<R, U, SRC>, 256, smem));
    const float ms = time_ms([&] { scan_kernel<R, U, SRC><<<grid, 256, smem>>>(d_filt, npad, reps, d_out); });
    CK(cudaGetLastError());
    const double tests = (double)grid * 256 * R * (double)npad * reps;
    const double gts = tests / (ms * 1e-3) / 1e9;
    // slots per test at 1.965 GHz nominal: SM issue slots/s = sms*4*clk warps-instr/s = sms*128*clk lane-slots/s
    const double lane_slots = (double)sms * 128 * 1.965e9;
    printf("%-28s R=%d U=%d blocks/SM=%d (occ %d)  %8.3f ms  %8.1f Gtests/s  %.2f lane-slots/test @1.965GHz\n", name, R, U,
           blocks_per_sm, occ, ms, gts, lane_slots / (gts * 1e9));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    const int n = 485, npad = 488;
    std::vector<float4> h(4096);
    for (int k = 0; k < 4096; ++k) {
        float cx = (float)((k * 37) % 23) - 11.f, cy = 0.2f, cz = (float)((k * 53) % 23) - 11.f, r = 0.2f;
        if (k == 0) { cx = 0; cy = -1000.f; cz = 0; r = 1000.f; }
        h[k] = make_float4(cx, cy, cz, k < n ? cx * cx + cy * cy + cz * cz - r * r : INFINITY);
    }
    float4* d_filt; unsigned* d_out; float* d_in; float* d_fo;
    CK(cudaMalloc(&d_filt, 4096 * 16)); CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_in, 256)); CK(cudaMalloc(&d_fo, 64));
    CK(cudaMemcpy(d_filt, h.data(), 4096 * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpyToSymbol(c_filt, h.data(), 4096 * 16));
    std::vector<float> hin(64);
    for (int i = 0; i < 64; ++i) hin[i] = 1.0f + 1e-7f * i;
    CK(cudaMemcpy(d_in, hin.data(), 256, cudaMemcpyHostToDevice));

    {
        const int grid = sms * 8, iters = 4096;
        const double fmas = (double)grid * 256 * iters * 16.0 * 8.0;
        float ms0 = time_ms([&] { ffma_kernel<0><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        float ms1 = time_ms([&] { ffma_kernel<1><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        float ms2 = time_ms([&] { ffma_kernel<2><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        printf("FFMA const operands : %.2f T FMA/s\nFFMA 3 reg operands : %.2f T FMA/s\nFFMA shared mult reg: %.2f T FMA/s   (nominal %.2f)\n",
               fmas / ms0 / 1e9, fmas / ms1 / 1e9, fmas / ms2 / 1e9, sms * 128 * 1.965e-3);
    }
    for (int bps : {2, 3, 4, 6, 8}) {
        run_scan<1, 4, 0>("smem LDS.128", d_filt, npad, d_out, sms, bps);
        run_scan<2, 4, 0>("smem LDS.128", d_filt, npad, d_out, sms, bps);
        run_scan<1, 8, 0>("smem LDS.128", d_filt, npad, d_out, sms, bps);
        run_scan<2, 8, 0>("smem LDS.128", d_filt, npad, d_out, sms, bps);
        run_scan<4, 4, 0>("smem LDS.128", d_filt, npad, d_out, sms, bps);
        run_scan<1, 8, 1>("constant bank", d_filt, npad, d_out, sms, bps);
        run_scan<2, 4, 1>("constant bank", d_filt, npad, d_out, sms, bps);
        run_scan<2, 8, 1>("constant bank", d_filt, npad, d_out, sms, bps);
        run_scan<4, 4, 1>("constant bank", d_filt, npad, d_out, sms, bps);
        run_scan<2, 8, 2>("global LDG", d_filt, npad, d_out, sms, bps);
    }
    return 0;
}
