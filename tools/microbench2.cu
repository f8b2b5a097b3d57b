// tools/microbench2.cu -- FFMA2 (fma.rn.f32x2, sm_100 packed FP32) probes for the cull scan (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false tools/microbench2.cu -o tools/microbench2
// Prints FFMA2 peak with uniform / register operands and, for packed scan-loop shapes, Gtests/s and lane-slots
// per test (SMs x 128 x 1.965 GHz / tests per second) next to the scalar constant-bank loop of microbench.cu.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void upk(u64 v, float& x, float& y) { asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
__device__ __forceinline__ u64 neg2(u64 v) { return v ^ 0x8000000080000000ull; }

struct CullRay { float dx, dy, dz, ndo, mx, my, mz, o2; };
__device__ __forceinline__ CullRay make_ray(int seed) {
    CullRay f;
    const float a = 0.001f * (float)(seed % 1000), b = 0.002f * (float)(seed % 777);
    f.dx = __sinf(a) * __cosf(b); f.dy = __cosf(a); f.dz = __sinf(a) * __sinf(b);
    const float ox = 13.f + a, oy = 2.f + b, oz = 3.f + a * b;
    f.ndo = -(f.dx * ox + f.dy * oy + f.dz * oz);
    f.mx = 2.f * ox; f.my = 2.f * oy; f.mz = 2.f * oz;
    f.o2 = ox * ox + oy * oy + oz * oz;
    return f;
}

// layout A ("ray pairs"): per sphere 4 x u64 = {cx,cx},{cy,cy},{cz,cz},{w,w}
// layout B ("sphere pairs"): per 2 spheres 4 x u64 = {cx0,cx1},{cy0,cy1},{cz0,cz1},{w0,w1}
__constant__ u64 c_pk[4 * 1024];
__constant__ float4 c_filt[1024];

// ---- scalar reference loop (as microbench.cu, constant bank)
template <int R, int U>
__global__ void __launch_bounds__(256) scan_scalar(int npad, int reps, unsigned* out) {
    CullRay f[R];
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = make_ray(threadIdx.x * 7 + blockIdx.x * 13 + r * 101);
    unsigned hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 1
        for (int k = 0; k < npad; k += U) {
            float4 s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) s[u] = c_filt[k + u];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float m = -INFINITY;
                float D[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float b = fmaf(f[r].dz, s[u].z, f[r].ndo);
                    b = fmaf(f[r].dy, s[u].y, b);
                    b = fmaf(f[r].dx, s[u].x, b);
                    float P = fmaf(f[r].mx, s[u].x, s[u].w);
                    P = fmaf(f[r].my, s[u].y, P);
                    P = fmaf(f[r].mz, s[u].z, P);
                    D[u] = fmaf(b, b, P);
                    m = fmaxf(m, D[u]);
                }
                if (!(m < f[r].o2)) {
#pragma unroll
                    for (int u = 0; u < U; ++u) if (!(D[u] < f[r].o2)) hits += k + u;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) f[r].o2 += 1e-3f;
    }
    if (hits == 0xdeadbeef) out[0] = hits;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = hits;
}

// ---- layout A: RP ray pairs per lane (2*RP rays), U spheres per step.
// VAR 0: P = fma2(m, c, w) (two sphere operands in the first P op; compiler's choice how to feed w)
// VAR 1: threshold folded into the chain (q seeded with o2), + add2 of w: 8 packed ops, pass = D >= 0
template <int RP, int U, int VAR>
__global__ void __launch_bounds__(256) scan_raypairs(int npad, int reps, unsigned* out) {
    u64 dx[RP], dy[RP], dz[RP], ndo[RP], mx[RP], my[RP], mz[RP], o2[RP];
    float o2a[RP], o2b[RP];
#pragma unroll
    for (int r = 0; r < RP; ++r) {
        CullRay a = make_ray(threadIdx.x * 7 + blockIdx.x * 13 + r * 101), b = make_ray(threadIdx.x * 7 + blockIdx.x * 13 + r * 101 + 50);
        dx[r] = pk(a.dx, b.dx); dy[r] = pk(a.dy, b.dy); dz[r] = pk(a.dz, b.dz); ndo[r] = pk(a.ndo, b.ndo);
        mx[r] = pk(a.mx, b.mx); my[r] = pk(a.my, b.my); mz[r] = pk(a.mz, b.mz); o2[r] = pk(-a.o2, -b.o2);
        o2a[r] = a.o2; o2b[r] = b.o2;
    }
    unsigned hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        int kv;
        kv = min((int)threadIdx.x, reps >> 20);  // 0, but lane-dependent as far as ptxas can tell
#pragma unroll 1
        for (int k = 0; k < npad; k += U, kv += U) {
            u64 cx[U], cy[U], cz[U], w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { cx[u] = c_pk[4 * (k + u)]; cy[u] = c_pk[4 * (k + u) + 1]; cz[u] = c_pk[4 * (k + u) + 2]; w[u] = c_pk[4 * ((VAR == 2 ? kv : k) + u) + 3]; }
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                float ma = -INFINITY, mb = -INFINITY;
                float Da[U], Db[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    u64 b = fma2(dz[r], cz[u], ndo[r]);
                    b = fma2(dy[r], cy[u], b);
                    b = fma2(dx[r], cx[u], b);
                    u64 D;
                    if (VAR != 1) {   // mx.. = +2o, w = -(|c|^2 - r^2): P is built negated, no sign flips
                        u64 P = fma2(mx[r], cx[u], w[u]);
                        P = fma2(my[r], cy[u], P);
                        P = fma2(mz[r], cz[u], P);
                        D = fma2(b, b, P);
                    } else {
                        u64 q = fma2(mx[r], cx[u], o2[r]);
                        q = fma2(my[r], cy[u], q);
                        q = fma2(mz[r], cz[u], q);
                        q = add2(q, w[u]);
                        D = fma2(b, b, q);
                    }
                    upk(D, Da[u], Db[u]);
                    ma = fmaxf(ma, Da[u]); mb = fmaxf(mb, Db[u]);
                }
                const float ta = VAR == 0 ? o2a[r] : 0.f, tb = VAR == 0 ? o2b[r] : 0.f;
                if (!(ma < ta) || !(mb < tb)) {
#pragma unroll
                    for (int u = 0; u < U; ++u) { if (!(Da[u] < ta)) hits += k + u; if (!(Db[u] < tb)) hits += 3 * (k + u); }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RP; ++r) { o2a[r] += 1e-3f; o2b[r] += 1e-3f; o2[r] = pk(-o2a[r], -o2b[r]); }
    }
    if (hits == 0xdeadbeef) out[0] = hits;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = hits;
}

// ---- layout B: sphere pairs; R rays per lane with duplicated constants; U sphere PAIRS per step
template <int R, int U, int VAR>
__global__ void __launch_bounds__(256) scan_spherepairs(int npad, int reps, unsigned* out) {
    u64 dx[R], dy[R], dz[R], ndo[R], mx[R], my[R], mz[R], o2[R];
    float o2s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        CullRay a = make_ray(threadIdx.x * 7 + blockIdx.x * 13 + r * 101);
        dx[r] = pk(a.dx, a.dx); dy[r] = pk(a.dy, a.dy); dz[r] = pk(a.dz, a.dz); ndo[r] = pk(a.ndo, a.ndo);
        mx[r] = pk(a.mx, a.mx); my[r] = pk(a.my, a.my); mz[r] = pk(a.mz, a.mz); o2[r] = pk(-a.o2, -a.o2);
        o2s[r] = a.o2;
    }
    unsigned hits = 0;
    for (int rep = 0; rep < reps; ++rep) {
        int kv;
        kv = min((int)threadIdx.x, reps >> 20);  // 0, but lane-dependent as far as ptxas can tell
#pragma unroll 1
        for (int k = 0; k < npad / 2; k += U, kv += U) {
            u64 cx[U], cy[U], cz[U], w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { cx[u] = c_pk[4 * (k + u)]; cy[u] = c_pk[4 * (k + u) + 1]; cz[u] = c_pk[4 * (k + u) + 2]; w[u] = c_pk[4 * ((VAR == 2 ? kv : k) + u) + 3]; }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float m = -INFINITY;
                float Da[U], Db[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    u64 b = fma2(dz[r], cz[u], ndo[r]);
                    b = fma2(dy[r], cy[u], b);
                    b = fma2(dx[r], cx[u], b);
                    u64 D;
                    if (VAR != 1) {   // mx.. = +2o, w = -(|c|^2 - r^2): P is built negated, no sign flips
                        u64 P = fma2(mx[r], cx[u], w[u]);
                        P = fma2(my[r], cy[u], P);
                        P = fma2(mz[r], cz[u], P);
                        D = fma2(b, b, P);
                    } else {
                        u64 q = fma2(mx[r], cx[u], o2[r]);
                        q = fma2(my[r], cy[u], q);
                        q = fma2(mz[r], cz[u], q);
                        q = add2(q, w[u]);
                        D = fma2(b, b, q);
                    }
                    upk(D, Da[u], Db[u]);
                    m = fmaxf(m, fmaxf(Da[u], Db[u]));
                }
                const float t = VAR == 0 ? o2s[r] : 0.f;
                if (!(m < t)) {
#pragma unroll
                    for (int u = 0; u < U; ++u) { if (!(Da[u] < t)) hits += k + u; if (!(Db[u] < t)) hits += 3 * (k + u); }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { o2s[r] += 1e-3f; o2[r] = pk(-o2s[r], -o2s[r]); }
    }
    if (hits == 0xdeadbeef) out[0] = hits;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[1] = hits;
}

// ---- peaks.  MODE 0: FFMA R,R,c,c   1: FFMA2 x = x*UR + y (one uniform pair)   2: FFMA2 three register pairs
//              3: FFMA2 x = x*y0 + y (shared multiplier register pair)
__constant__ u64 c_one[16];
template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(float* out, const float* in, int iters, float ca, float cb) {
    u64 x[8], a[8], b[8];
    float xs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        xs[i] = threadIdx.x + i;
        x[i] = pk(threadIdx.x + i, threadIdx.x - i);
        a[i] = pk(in[(threadIdx.x + i) & 63], in[(threadIdx.x + i + 7) & 63]);
        b[i] = pk(in[(threadIdx.x + 2 * i + 1) & 63], in[(threadIdx.x + 2 * i + 9) & 63]);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) xs[i] = fmaf(xs[i], ca, cb);
                else if (MODE == 1) x[i] = fma2(x[i], c_one[u], b[i]);
                else if (MODE == 2) x[i] = fma2(x[i], a[i], b[i]);
                else x[i] = fma2(x[i], a[0], b[i]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float p, q; upk(x[i], p, q); s += p + q + xs[i]; }
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

static int g_sms = 148;
template <typename K>
void report(const char* name, K kernel, int rays_per_lane, int U, int npad, int bps, unsigned* d_out) {
    const int reps = 400, grid = g_sms * bps;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0));
    const float ms = time_ms([&] { kernel<<<grid, 256>>>(npad, reps, d_out); });
    CK(cudaGetLastError());
    unsigned h[2]; CK(cudaMemcpy(h, d_out, 8, cudaMemcpyDeviceToHost));
    const double tests = (double)grid * 256 * rays_per_lane * (double)npad * reps;
    const double gts = tests / (ms * 1e-3) / 1e9;
    printf("%-34s rays/lane=%d U=%d blocks/SM=%d (occ %d) %8.3f ms %8.1f Gtests/s  %.2f lane-slots/test  [chk %u]\n", name,
           rays_per_lane, U, bps, occ, ms, gts, (double)g_sms * 128 * 1.965e9 / (gts * 1e9), h[1]);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    g_sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, g_sms);
    const int n = 485, npad = 496;  // multiple of 16
    std::vector<float4> h(1024);
    std::vector<u64> ha(4096), hb(4096);
    auto bits = [](float a, float b) { uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return (u64)x | ((u64)y << 32); };
    for (int k = 0; k < 1024; ++k) {
        float cx = (float)((k * 37) % 23) - 11.f, cy = 0.2f, cz = (float)((k * 53) % 23) - 11.f, r = 0.2f;
        if (k == 0) { cx = 0; cy = -1000.f; cz = 0; r = 1000.f; }
        h[k] = make_float4(cx, cy, cz, k < n ? -(cx * cx + cy * cy + cz * cz - r * r) : -INFINITY);
    }
    for (int k = 0; k < 1024; ++k) {
        ha[4 * k] = bits(h[k].x, h[k].x); ha[4 * k + 1] = bits(h[k].y, h[k].y); ha[4 * k + 2] = bits(h[k].z, h[k].z); ha[4 * k + 3] = bits(h[k].w, h[k].w);
    }
    for (int k = 0; k < 512; ++k) {
        hb[4 * k] = bits(h[2 * k].x, h[2 * k + 1].x); hb[4 * k + 1] = bits(h[2 * k].y, h[2 * k + 1].y);
        hb[4 * k + 2] = bits(h[2 * k].z, h[2 * k + 1].z); hb[4 * k + 3] = bits(h[2 * k].w, h[2 * k + 1].w);
    }
    unsigned* d_out; float* d_in; float* d_fo;
    CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_in, 256)); CK(cudaMalloc(&d_fo, 64));
    CK(cudaMemcpyToSymbol(c_filt, h.data(), 1024 * 16));
    std::vector<float> hin(64);
    for (int i = 0; i < 64; ++i) hin[i] = 1.0f + 1e-7f * i;
    CK(cudaMemcpy(d_in, hin.data(), 256, cudaMemcpyHostToDevice));
    std::vector<u64> hone(16);
    for (int i = 0; i < 16; ++i) hone[i] = bits(1.0f + 1e-7f * i, 1.0f - 1e-7f * i);
    CK(cudaMemcpyToSymbol(c_one, hone.data(), 128));
    {
        const int grid = g_sms * 8, iters = 4096;
        const double ops = (double)grid * 256 * iters * 16.0 * 8.0;
        float ms0 = time_ms([&] { peak_kernel<0><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        float ms1 = time_ms([&] { peak_kernel<1><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        float ms2 = time_ms([&] { peak_kernel<2><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        float ms3 = time_ms([&] { peak_kernel<3><<<grid, 256>>>(d_fo, d_in, iters, 1.0000001f, 1e-9f); });
        printf("FFMA  R,R,c,c          : %.2f T FMA/s\n", ops / ms0 / 1e9);
        printf("FFMA2 R,R,UR,R         : %.2f T FMA/s (%.2f T inst-lanes/s)\n", 2 * ops / ms1 / 1e9, ops / ms1 / 1e9);
        printf("FFMA2 three reg pairs  : %.2f T FMA/s (%.2f T inst-lanes/s)\n", 2 * ops / ms2 / 1e9, ops / ms2 / 1e9);
        printf("FFMA2 shared multiplier: %.2f T FMA/s (%.2f T inst-lanes/s)   nominal scalar %.2f\n", 2 * ops / ms3 / 1e9, ops / ms3 / 1e9,
               g_sms * 128 * 1.965e-3);
    }
    CK(cudaMemcpyToSymbol(c_pk, ha.data(), 4096 * 8));
    for (int bps : {2, 4, 8}) {
        report("scalar const bank", scan_scalar<2, 8>, 2, 8, npad, bps, d_out);
        report("scalar const bank", scan_scalar<4, 4>, 4, 4, npad, bps, d_out);
        report("ray pairs VAR0 (P seeded w)", scan_raypairs<1, 8, 0>, 2, 8, npad, bps, d_out);
        report("ray pairs VAR1 (add2 w)", scan_raypairs<1, 8, 1>, 2, 8, npad, bps, d_out);
        report("ray pairs VAR0 (P seeded w)", scan_raypairs<2, 4, 0>, 4, 4, npad, bps, d_out);
        report("ray pairs VAR1 (add2 w)", scan_raypairs<2, 4, 1>, 4, 4, npad, bps, d_out);
        report("ray pairs VAR0 (P seeded w)", scan_raypairs<2, 8, 0>, 4, 8, npad, bps, d_out);
        report("ray pairs VAR2 (w via LDC)", scan_raypairs<1, 8, 2>, 2, 8, npad, bps, d_out);
        report("ray pairs VAR2 (w via LDC)", scan_raypairs<2, 4, 2>, 4, 4, npad, bps, d_out);
        report("ray pairs VAR2 (w via LDC)", scan_raypairs<2, 8, 2>, 4, 8, npad, bps, d_out);
        report("ray pairs VAR1 (add2 w)", scan_raypairs<2, 8, 1>, 4, 8, npad, bps, d_out);
    }
    CK(cudaMemcpyToSymbol(c_pk, hb.data(), 4096 * 8));
    for (int bps : {2, 4, 8}) {
        report("sphere pairs VAR0", scan_spherepairs<1, 8, 0>, 1, 16, npad, bps, d_out);
        report("sphere pairs VAR1", scan_spherepairs<1, 8, 1>, 1, 16, npad, bps, d_out);
        report("sphere pairs VAR0", scan_spherepairs<2, 4, 0>, 2, 8, npad, bps, d_out);
        report("sphere pairs VAR1", scan_spherepairs<2, 4, 1>, 2, 8, npad, bps, d_out);
        report("sphere pairs VAR0", scan_spherepairs<2, 8, 0>, 2, 16, npad, bps, d_out);
        report("sphere pairs VAR2", scan_spherepairs<1, 8, 2>, 1, 16, npad, bps, d_out);
        report("sphere pairs VAR2", scan_spherepairs<2, 4, 2>, 2, 8, npad, bps, d_out);
        report("sphere pairs VAR2", scan_spherepairs<2, 8, 2>, 2, 16, npad, bps, d_out);
        report("sphere pairs VAR1", scan_spherepairs<2, 8, 1>, 2, 16, npad, bps, d_out);
    }
    return 0;
}
