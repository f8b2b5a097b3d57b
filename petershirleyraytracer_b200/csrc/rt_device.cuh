// rt_device.cuh -- device-side building blocks of the sm_100a path-tracing hot path.
//
// Reference semantics (fengye/PeterShirleyRaytracer, cited as programs/<file>:<line>) are reproduced
// EXACTLY: every quantity that feeds a decision or the image is computed in FP64 with one rounding per
// operation in the reference's evaluation order (__dmul_rn/__dadd_rn/... never contract into FMA).
// What makes it fast is that the O(N) part of hittable_list::hit (programs/hittable_list.cc:9-17) is a
// *conservative FP32 cull*: 7 FP32-pipe instructions per (ray, sphere) decide "cannot be hit" with a
// rigorous error bound; only the few survivors run the FP64 sphere::hit (programs/sphere.cc:3-40), in
// list order with the shrinking tmax, so index / t / p / normal are bit-identical to the reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

// ---------------------------------------------------------------- constants
constexpr int kTileW = 8, kTileH = 8, kTilePix = kTileW * kTileH;  // one warp renders one tile at a time
constexpr int kThreads = 128;                                       // 4 warps per CTA
constexpr int kWarps = kThreads / 32;
constexpr int kCandCap = 12;        // survivors per cast kept in smem; more -> full FP64 scan for that cast
constexpr int kScanStep = 12;       // cull entries per scan step; arrays are padded to a multiple of this ...
constexpr int kScanPad = 8;         // ... plus this many never-pass entries (prefetch runs two groups ahead)
constexpr int kMaxLinear = 4080;    // cull entries that fit the 64 KB constant bank (with the prefetch pad)
constexpr int kMaxLinearSmem = 11520;  // ... that fit one CTA's shared memory next to its path state (TMA-staged variant)
constexpr int kFixShift = 44;       // radiance accumulates as 20.44 fixed point (order-independent sums)
constexpr int kNumStats = 10;

enum StatSlot { ST_SAMPLES = 0, ST_CASTS, ST_SPHERE_TESTS, ST_NODE_TESTS, ST_EXACT_TESTS, ST_BLACK, ST_EARLY_OUTS,
                ST_PRIMARY_HITS, ST_OVERFLOWS, ST_SELF_RESOLVED };

// error-bound constants of the FP32 cull (see DESIGN.md "cull error bound"): with S = |c| + |o| and u = 2^-24,
// |D_fp32 - D| <= u * (19 S^2 + 4 r^2) <= u * (38 |c|^2 + 38 |o|^2 + 4 r^2)  [b: 5uS -> b^2: 10uS^2; inputs of
// 2o.c: uS^2; four FFMA roundings of magnitudes <= 2S^2 + r^2]; we fold u * (48 |c|^2 + 8 r^2) into the
// per-sphere constant and u * 48 |o|^2 into the per-ray constant.
constexpr double kCullEps = 5.9604644775390625e-08;  // 2^-24
constexpr double kCullKc = 48.0, kCullKr = 8.0, kCullKo = 48.0;
// the FP32 cull is used only while |c|^2, r^2 and |o|^2 stay below this (no overflow / NaN in FP32)
constexpr double kCullMaxMag2 = 1e30;

// ---------------------------------------------------------------- FP64, reference evaluation order
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
// num / A for the hit distance (programs/sphere.cc:24,29), correctly rounded (== __ddiv_rn) but cheap.
// All candidates of a cast divide by the same A = dot(dir, dir), so the cast computes rA = RN(1/A) once
// (RcpA) and every quotient is  q0 = RN(num*rA); q1 = RN(q0 + (num - A*q0)*rA); q = RN(q1 + (num - A*q1)*rA)
// with the residuals exact in one FMA each.  q1 is a faithful rounding of num/A (|q1 - num/A| <= half an ulp
// plus 1.5 ulp * 2^-53), so by Markstein's theorem (Muller et al., Handbook of Floating-Point Arithmetic,
// "division by Newton-Raphson": y = RN(1/b), q faithful, r = a - b*q exact => RN(q + r*y) = RN(a/b)) the last
// step is the IEEE quotient.  The theorem needs no over/underflow, so the short form is used only when the
// exponents of A and num are moderate; everything else takes __ddiv_rn.  (tools/div_check.c compares the two on
// 4e8 random and adversarial operand pairs.)  With tmin = 0 about half of all hits are self-hits whose
// numerator is exactly +-0: +-0 / A = +-0 for finite A > 0 is answered directly.
struct RcpA {
    double A, rA;
    bool fast;  // 2^-200 < A < 2^200: the short division is valid for moderate numerators
};
__device__ __forceinline__ RcpA make_rcp(double A) {
    RcpA d;
    d.A = A;
    const int e = (__double2hiint(A) >> 20) & 0x7ff;   // sign bit excluded; A <= 0, inf, NaN fail the range test
    d.fast = A > 0.0 && e > 1023 - 200 && e < 1023 + 200;
    d.rA = __drcp_rn(d.fast ? A : 1.0);
    return d;
}
// (out of line: the IEEE division sequence is needed once in a blue moon; inlined at every call site it bloats
// the hit test with its own branch structure)
__device__ __noinline__ double ddiv_ieee(double num, double A) { return __ddiv_rn(num, A); }

__device__ __forceinline__ double ddiv_t(double num, const RcpA& d) {
    const unsigned e = ((unsigned)__double2hiint(num) >> 20) & 0x7ffu;
    const bool zero = num == 0.0;
    // the short form runs for every lane (no branch); only operands outside its range take the subroutine
    const double q0 = __dmul_rn(num, d.rA);
    const double q1 = __fma_rn(__fma_rn(-d.A, q0, num), d.rA, q0);
    double q = __fma_rn(__fma_rn(-d.A, q1, num), d.rA, q1);
    if (!(d.fast && (zero || (e - 324u) < 1399u))) q = ddiv_ieee(num, d.A);  // exponent outside (-700, 700)
    return (zero && d.fast) ? num : q;  // +-0 / A = +-0 (the short form would lose the sign of -0)
}
// programs/vec3.h:156-159: (u0*v0 + u1*v1) + u2*v2
__device__ __forceinline__ double ddot(double ax, double ay, double az, double bx, double by, double bz) {
    return dadd(dadd(dmul(ax, bx), dmul(ay, by)), dmul(az, bz));
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}
// 32-bit word -> [0,1): exact in double
__device__ __forceinline__ double u32_unit(uint32_t w) { return dmul((double)w, 1.0 / 4294967296.0); }

// ---------------------------------------------------------------- scene / launch arguments
struct SceneDev {
    const float4* filt;    // npad entries {cx, cy, cz, -(|c|^2 - r^2 - E_k)} (FP32 cull), padded with never-pass entries
    const double4* exact;  // n entries {cx, cy, cz, r} (FP64, list order)
    const double* inv_r;   // n entries RN(1.0 / r): the reciprocal of programs/vec3.h:151-154, tabulated at upload
    int n, npad;
    const float4* bvh_nodes;   // 2 float4 per child, kBvhW children per node (rt_bvh.h: Bvh4Node), root = node 0; NULL if not built
    const int32_t* bvh_leaf;   // sphere list indices, leaf by leaf
    // tie grid (rt_bvh.h: TieGridHost): O(1) answer to "which spheres' surfaces pass through this point"
    const float4* sph32;       // n entries {cx, cy, cz, |r|} in FP32
    const int4* tie_cells;     // 4 sphere indices per cell (-1 unused; .x == -2: overfull, undecidable)
    float tie_g0[3], tie_g1[3], tie_inv_h, tie_rho_max;
    int tie_dimx, tie_dimy;
    int tie_ok, tie_ngiants;
    int tie_giants[4];
};

// The constants programs/main.cc hard-codes in ray_color, as launch arguments (rt_params).
// custom == 0 selects the reference's own values through the original code path (exact 0.5^k attenuation).
struct ShadeDev {
    double albedo;             // programs/main.cc:43
    double sky_a[3], sky_b[3]; // programs/main.cc:48: (1-t)*sky_a + t*sky_b
    int custom;                // 0: albedo 0.5, sky (1,1,1) -> (0.5,0.7,1.0), the fields above are not read
    int lambertian;            // 0: vec3::random_in_hemisphere (main.cc:42); 1: vec3::random_unit_vector (vec3.h:97-100)
};

struct RenderArgs {
    SceneDev sc;
    ShadeDev sh;
    double cam_org[3], cam_llc[3], cam_hor[3], cam_ver[3];
    double tmin;
    int W, H, spp, max_depth;
    uint32_t key0, key1;
    int jitter, early_out, scan_mode;
    int tiles_x, tiles_total, shard_rank, shard_count, tiles_local;
    // Work units = (tile, sample chunk).  A tile's samples are cut into GRADED chunks: lv_n[0] chunks of lv_spp[0] samples
    // per pixel, then lv_n[1] of lv_spp[1], then lv_n[2] of lv_spp[2] (the very last chunk may be shorter); unit ids are
    // level-major (every tile's long chunks first), so a launch ends on short units (rt_api.cu: launch_render).
    int chunks, units_local;       // chunks per tile over all levels; units = tiles_local * chunks
    int lv_n[3], lv_spp[3];
    int compact_out;
    uchar4* out;
    double* sum_out;               // optional W*H*3
    unsigned long long* frame_accum;  // optional W*H*3 fixed-point sums carried across passes (progressive rendering)
    int sample_base;               // first sample index of this pass (Philox streams are keyed on the absolute index)
    unsigned int* unit_counter;    // persistent-warp work queue head
    unsigned long long* accum;     // tiles_local * 192 fixed-point sums (used when chunks > 1)
    unsigned int* tile_done;       // tiles_local chunk-completion counters
    unsigned long long* stats;
};

// ---------------------------------------------------------------- TMA bulk staging of the cull array
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One thread arms an mbarrier with the byte count and issues cp.async.bulk (TMA, SASS UBLKCP) pieces;
// every thread then waits on the barrier's phase 0.  bytes is a multiple of 16.
__device__ __forceinline__ void stage_bulk(void* s_dst, const void* g_src, uint32_t bytes, uint64_t* s_mbar) {
    const uint32_t mbar = smem_u32(s_mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        const uint32_t kPiece = 16384;
        for (uint32_t off = 0; off < bytes; off += kPiece) {
            const uint32_t sz = (bytes - off < kPiece) ? (bytes - off) : kPiece;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32((const char*)s_dst + off)),
                "l"((const char*)g_src + off), "r"(sz), "r"(mbar)
                : "memory");
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar)
            : "memory");
    }
}

// ---------------------------------------------------------------- the FP32 conservative cull
// Cull entries live in the constant bank: a warp-uniform index makes the loads ULDC/LDCU into UNIFORM
// registers, so every FFMA of the scan reads two vector registers + one uniform register.  Measured on
// B200 (tools/microbench.cu): FFMA with three distinct vector-register sources sustains 22.8 T FMA/s,
// with a uniform/constant operand 36.2 T FMA/s; the scan costs 11.8 issue slots per test from the
// constant bank against 13.7 from shared memory (LDS.128 + three-register FFMAs).  The shared-memory
// variant (TMA-staged) is kept selectable for comparison (rt_params.reserved[2]); it is the only linear scan
// for scenes beyond the 64 KB constant bank (up to kMaxLinearSmem entries).
__constant__ float4 c_filt[kMaxLinear + kScanPad];

struct CullRay {  // per-cast constants, 8 registers
    float dx, dy, dz;  // unit direction
    float ndo;         // -(d̂ . o)
    float mx, my, mz;  // +2 o
    float o2;          // |o|^2 - E_o, rounded down: the pass threshold
};

// Builds the cull constants of one ray in FP64 and rounds once.  A dead slot gets constants for which no
// entry can pass (D = -inf).
__device__ __forceinline__ CullRay make_cull_ray(bool alive, double ox, double oy, double oz, double dx, double dy,
                                                 double dz, double A) {
    CullRay f;
    if (alive) {
        const double inv = rsqrt(A);
        const double ux = dx * inv, uy = dy * inv, uz = dz * inv;
        f.dx = (float)ux; f.dy = (float)uy; f.dz = (float)uz;
        f.ndo = (float)(-(ux * ox + uy * oy + uz * oz));
        f.mx = (float)(2.0 * ox); f.my = (float)(2.0 * oy); f.mz = (float)(2.0 * oz);
        const double g2 = ox * ox + oy * oy + oz * oz;
        f.o2 = __double2float_rd(g2 - kCullEps * kCullKo * g2);
    } else {
        f.dx = f.dy = f.dz = 0.f; f.ndo = 0.f; f.mx = f.my = f.mz = 0.f;
        f.o2 = __int_as_float(0x7f800000);
    }
    return f;
}

// The line through (o, d̂) can touch sphere (c, r) only if  (d̂.(c-o))^2 - |c-o|^2 + r^2 >= 0.  Expanded around
// the world origin this is  b^2 - w + 2 o.c >= |o|^2  with b = d̂.c - d̂.o and w = |c|^2 - r^2, i.e. 7 FP32-pipe
// instructions per (ray, sphere): 3 FFMA (b), 1 FFMA (b*b - w), 3 FFMA (+ 2 o.c); the per-sphere and per-ray
// constants carry the error bound, so "pass" is conservative.  The order is chosen so that EVERY FFMA has exactly
// one per-sphere operand: in the scan those live in uniform registers, an FFMA takes one uniform operand
// (multiplicand or addend), and a form with two of them (b^2 - (w + ...) seeded with w) costs one extra move of
// w into a vector register per sphere and lane.
__device__ __forceinline__ float cull_D(const CullRay& f, const float4 s) {
    float b = fmaf(f.dz, s.z, f.ndo);
    b = fmaf(f.dy, s.y, b);
    b = fmaf(f.dx, s.x, b);
    float D = fmaf(b, b, s.w);   // s.w = -(|c|^2 - r^2 - E_k)
    D = fmaf(f.mx, s.x, D);      // f.m = +2 o
    D = fmaf(f.my, s.y, D);
    D = fmaf(f.mz, s.z, D);
    return D;  // passes unless this is < f.o2
}

// Scans the npad cull entries (kConst: constant bank, else shared memory) for R rays at once, 12 entries per
// step in three groups of 4, software-pipelined two groups deep: while group g is evaluated, the loads of
// groups g+1 and g+2 are in flight (48 of the 63 uniform registers; the arrays carry 8 extra never-pass
// entries so the last prefetches stay in bounds).  Per (ray, step) one running maximum decides whether any
// of the 12 entries can pass; survivors (list order) go to the per-slot candidate lists
// cand[(e*R + r)*stride].  All lanes execute the same instruction stream; the only divergent code is the
// (rare) append.  Inputs are finite by construction (scene validated at upload, ray checked by the caller),
// so no value here is NaN.
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

template <int R, bool kConst>
__device__ __forceinline__ void cull_scan(const float4* __restrict__ s_filt, int npad, const CullRay (&f)[R],
                                          uint16_t* cand, int stride, int (&cnt)[R], bool (&ovf)[R]) {
    // kv mirrors k in a VECTOR register (opaque to the compiler).  If the stored index were derived from
    // k itself, ptxas would move the loop counter off the uniform datapath and load every cull entry into
    // vector registers (LDC instead of LDCU): three-register FFMAs at 63% rate instead of FFMA R,UR,R.
    int kv;
    asm volatile("mov.u32 %0, 0;" : "=r"(kv));
    float4 g0[4], g1[4], g2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        g0[u] = kConst ? c_filt[u] : s_filt[u];
        g1[u] = kConst ? c_filt[4 + u] : s_filt[4 + u];
    }
#pragma unroll 1
    for (int k = 0; k < npad; k += kScanStep, kv += kScanStep) {
        float D[R][kScanStep];
#pragma unroll
        for (int u = 0; u < 4; ++u) g2[u] = kConst ? c_filt[k + 8 + u] : s_filt[k + 8 + u];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 4; ++u) D[r][u] = cull_D(f[r], g0[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) g0[u] = kConst ? c_filt[k + 12 + u] : s_filt[k + 12 + u];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 4; ++u) D[r][4 + u] = cull_D(f[r], g1[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) g1[u] = kConst ? c_filt[k + 16 + u] : s_filt[k + 16 + u];
        bool any = false;
        float m[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) D[r][8 + u] = cull_D(f[r], g2[u]);
            // 12 -> 1 in six 3-input maxima (FMNMX3)
            m[r] = fmax3(fmax3(fmax3(D[r][0], D[r][1], D[r][2]), fmax3(D[r][3], D[r][4], D[r][5]), fmax3(D[r][6], D[r][7], D[r][8])),
                         fmax3(D[r][9], D[r][10], D[r][11]), D[r][0]);
            any |= !(m[r] < f[r].o2);
        }
        if (any) {  // some entry of this step may be hit by one of this lane's rays
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (m[r] < f[r].o2) continue;
                // 12-bit pass mask without branches (one compare + one select per entry), then one loop over
                // the set bits (list order).  The straightforward "if (pass) append" per entry compiles to 24
                // tiny divergent regions per step, which held 16 % of the kernel's warp-state samples.
                uint32_t pm = 0u;
#pragma unroll
                for (int u = 0; u < kScanStep; ++u) pm |= (D[r][u] < f[r].o2) ? 0u : (1u << u);
                while (pm) {
                    const int u = __ffs((int)pm) - 1;
                    pm &= pm - 1u;
                    if (cnt[r] < kCandCap) {
                        cand[(cnt[r] * R + r) * stride] = (uint16_t)(kv + u);
                        ++cnt[r];
                    } else {
                        ovf[r] = true;
                    }
                }
            }
        }
    }
}

// ---- packed FP32 (fma.rn.f32x2, SASS FFMA2) form of the constant-bank scan: one instruction evaluates one FMA of the
// test for TWO spheres.  The bank then holds the entries in pairs -- for spheres 2j, 2j+1 four 64-bit words
// {cx,cx'} {cy,cy'} {cz,cz'} {w,w'} (rt_api.cu: fill_scene) -- so that one LDCU.64 brings an operand pair into an aligned
// uniform-register pair, and the lane keeps its ray constants duplicated in both halves of vector-register pairs.
// Same arithmetic per half as cull_D (each half is an IEEE FP32 fma), same D, same survivors.
#ifndef RT_SCAN_PACKED
#define RT_SCAN_PACKED 1   // 0: the scalar FFMA scan (round 1), kept for A/B
#endif
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pack2(float x, float y) { u64 d; asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float& x, float& y) { asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(v)); }
struct CullRay2 { u64 dx, dy, dz, ndo, mx, my, mz; };
__device__ __forceinline__ CullRay2 dup_cull_ray(const CullRay& f) {
    CullRay2 g;
    g.dx = pack2(f.dx, f.dx); g.dy = pack2(f.dy, f.dy); g.dz = pack2(f.dz, f.dz); g.ndo = pack2(f.ndo, f.ndo);
    g.mx = pack2(f.mx, f.mx); g.my = pack2(f.my, f.my); g.mz = pack2(f.mz, f.mz);
    return g;
}
struct CullPair { u64 x, y, z, w; };
__device__ __forceinline__ void cull_D2(const CullRay2& f, const CullPair& s, float& D0, float& D1) {
    u64 b = fma2(f.dz, s.z, f.ndo);
    b = fma2(f.dy, s.y, b);
    b = fma2(f.dx, s.x, b);
    u64 D = fma2(b, b, s.w);
    D = fma2(f.mx, s.x, D);
    D = fma2(f.my, s.y, D);
    D = fma2(f.mz, s.z, D);
    unpack2(D, D0, D1);
}
__device__ __forceinline__ CullPair load_pair(int j) {   // pair j = entries 2j, 2j+1 of the (pair-packed) constant bank
    const u64* p = reinterpret_cast<const u64*>(c_filt) + 4 * j;
    CullPair s;
    s.x = p[0]; s.y = p[1]; s.z = p[2]; s.w = p[3];
    return s;
}
template <int R>
__device__ __forceinline__ void cull_scan_packed(int npad, const CullRay (&f)[R], uint16_t* cand, int stride, int (&cnt)[R],
                                                 bool (&ovf)[R]) {
    int kv;
    asm volatile("mov.u32 %0, 0;" : "=r"(kv));
    CullRay2 f2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) f2[r] = dup_cull_ray(f[r]);
    CullPair g0[2], g1[2], g2[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) { g0[u] = load_pair(u); g1[u] = load_pair(2 + u); }
#pragma unroll 1
    for (int k = 0; k < npad; k += kScanStep, kv += kScanStep) {
        float D[R][kScanStep];
#pragma unroll
        for (int u = 0; u < 2; ++u) g2[u] = load_pair(k / 2 + 4 + u);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 2; ++u) cull_D2(f2[r], g0[u], D[r][2 * u], D[r][2 * u + 1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) g0[u] = load_pair(k / 2 + 6 + u);
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 2; ++u) cull_D2(f2[r], g1[u], D[r][4 + 2 * u], D[r][4 + 2 * u + 1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) g1[u] = load_pair(k / 2 + 8 + u);
        bool any = false;
        float m[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int u = 0; u < 2; ++u) cull_D2(f2[r], g2[u], D[r][8 + 2 * u], D[r][8 + 2 * u + 1]);
            m[r] = fmax3(fmax3(fmax3(D[r][0], D[r][1], D[r][2]), fmax3(D[r][3], D[r][4], D[r][5]), fmax3(D[r][6], D[r][7], D[r][8])),
                         fmax3(D[r][9], D[r][10], D[r][11]), D[r][0]);
            any |= !(m[r] < f[r].o2);
        }
        if (any) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (m[r] < f[r].o2) continue;
                uint32_t pm = 0u;
#pragma unroll
                for (int u = 0; u < kScanStep; ++u) pm |= (D[r][u] < f[r].o2) ? 0u : (1u << u);
                while (pm) {
                    const int u = __ffs((int)pm) - 1;
                    pm &= pm - 1u;
                    if (cnt[r] < kCandCap) {
                        cand[(cnt[r] * R + r) * stride] = (uint16_t)(kv + u);
                        ++cnt[r];
                    } else {
                        ovf[r] = true;
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- FP64 exact sphere::hit
struct Best {
    double t;  // closest_so_far (programs/hittable_list.cc:7,14)
    double C;  // value of C (programs/sphere.cc:11) of the kept hit, for the exact early-out
    int k;     // list index of the kept hit, -1 = none
};

// programs/sphere.cc:3-32 for sphere k against (o, d) with A = dot(d,d) hoisted, tmax = best.t; on success
// the caller's record (t, k) is replaced, which is hittable_list.cc:13-15.  The hit record itself
// (p, normal) is only needed for the final winner and is built by make_record().
__device__ __forceinline__ void exact_test(const double4* __restrict__ exact, int k, double ox, double oy, double oz,
                                           double dx, double dy, double dz, const RcpA& dA, double tmin, Best& best) {
    const double A = dA.A;
    const double2 c01 = __ldg(reinterpret_cast<const double2*>(exact + k));
    const double2 c23 = __ldg(reinterpret_cast<const double2*>(exact + k) + 1);
    const double amx = dsub(ox, c01.x), amy = dsub(oy, c01.y), amz = dsub(oz, c23.x);  // sphere.cc:7
    const double HALF_B = ddot(dx, dy, dz, amx, amy, amz);                              // sphere.cc:10
    const double C = dsub(ddot(amx, amy, amz, amx, amy, amz), dmul(c23.y, c23.y));      // sphere.cc:11
    const double disc = dsub(dmul(HALF_B, HALF_B), dmul(A, C));                         // sphere.cc:14
    if (disc < 0) return;                                                                // sphere.cc:15-18
    const double sqrt_d = dsqrt(disc);
    double t = ddiv_t(dsub(-HALF_B, sqrt_d), dA);  // sphere.cc:24
    if (t < tmin || t > best.t) {                // sphere.cc:26 (closed interval)
        t = ddiv_t(dadd(-HALF_B, sqrt_d), dA);   // sphere.cc:29
        if (t < tmin || t > best.t) return;     // sphere.cc:30-31
    }
    best.t = t; best.C = C; best.k = k;
}

struct Record {  // programs/hittable.h:7-13
    double px, py, pz, nx, ny, nz;
    bool front_face;
};

// programs/sphere.cc:34-36 + programs/hittable.h:14-18 for the kept hit
__device__ __forceinline__ Record make_record(const SceneDev& sc, const Best& best, double ox, double oy,
                                              double oz, double dx, double dy, double dz) {
    const double2 c01 = __ldg(reinterpret_cast<const double2*>(sc.exact + best.k));
    const double2 c23 = __ldg(reinterpret_cast<const double2*>(sc.exact + best.k) + 1);
    Record rec;
    rec.px = dadd(ox, dmul(best.t, dx));  // programs/ray.h:27 orig + t*dir
    rec.py = dadd(oy, dmul(best.t, dy));
    rec.pz = dadd(oz, dmul(best.t, dz));
    const double inv_r = __ldg(sc.inv_r + best.k);  // programs/vec3.h:151-154: (1/t) * v, 1/r tabulated (same IEEE quotient)
    const double wx = dmul(inv_r, dsub(rec.px, c01.x));
    const double wy = dmul(inv_r, dsub(rec.py, c01.y));
    const double wz = dmul(inv_r, dsub(rec.pz, c23.x));
    rec.front_face = ddot(dx, dy, dz, wx, wy, wz) < 0;
    rec.nx = rec.front_face ? wx : -wx;
    rec.ny = rec.front_face ? wy : -wy;
    rec.nz = rec.front_face ? wz : -wz;
    return rec;
}

// sphere::hit for sphere k with the caller's fixed [tmin, tmax] (programs/sphere.cc:3-32), merged into `best` by
// the ORDER-INDEPENDENT form of hittable_list.cc:9-17: the list scan keeps the smallest accepted t and, on
// equal t, the later index.  (Equivalent because the far root is never smaller than the near root, so a sphere
// whose near root lies beyond closest_so_far is rejected either way.)  Used by the BVH traversal, which
// meets spheres in tree order.
__device__ __forceinline__ bool exact_test_unordered(const double4* __restrict__ exact, int k, double ox, double oy,
                                                     double oz, double dx, double dy, double dz, const RcpA& dA,
                                                     double tmin, double tmax, Best& best) {
    const double A = dA.A;
    const double2 c01 = __ldg(reinterpret_cast<const double2*>(exact + k));
    const double2 c23 = __ldg(reinterpret_cast<const double2*>(exact + k) + 1);
    const double amx = dsub(ox, c01.x), amy = dsub(oy, c01.y), amz = dsub(oz, c23.x);
    const double HALF_B = ddot(dx, dy, dz, amx, amy, amz);
    const double C = dsub(ddot(amx, amy, amz, amx, amy, amz), dmul(c23.y, c23.y));
    const double disc = dsub(dmul(HALF_B, HALF_B), dmul(A, C));
    if (disc < 0) return false;
    const double sqrt_d = dsqrt(disc);
    double t = ddiv_t(dsub(-HALF_B, sqrt_d), dA);
    if (t < tmin || t > tmax) {
        t = ddiv_t(dadd(-HALF_B, sqrt_d), dA);
        if (t < tmin || t > tmax) return false;
    }
    if (best.k < 0 || t < best.t || (t == best.t && k > best.k)) {
        best.t = t; best.C = C; best.k = k;
        return true;
    }
    return false;
}

// Closest hit through the flattened BVH (rt_bvh.h), exact semantics: box tests are conservative FP32 slab
// tests (boxes rounded outward and padded at build time; the ray origin is widened by its own FP32 rounding
// here), a subtree is skipped only if it is missed or its entry distance is strictly beyond the current
// best, and every sphere of a visited leaf runs the FP64 test above.  Requires tmin >= 0 and a finite,
// non-zero direction (callers route other rays to the sequential scan).  Nearer child first.
#ifndef RT_BVH_WIDTH
#define RT_BVH_WIDTH 4
#endif
#ifndef RT_BVH_LEAF
#define RT_BVH_LEAF 1   // spheres per leaf (== rt_bvh.h: kBvhLeafMax)
#endif
constexpr int kBvhW = RT_BVH_WIDTH;  // children per device BVH node (== rt_bvh.h: kBvhWidth)
static_assert(RT_BVH_WIDTH == 4 && RT_BVH_LEAF == 1, "the traversal is written for 4-wide nodes with single-sphere leaves (8-wide nodes and 2-6 sphere leaves were measured slower in round 1)");
constexpr int kBvhStack = 48;
// `best` is in/out: a hit already known (the start sphere's, see self_cast) bounds the traversal from the first
// node on; `skip` names a sphere that needs no further test (that start sphere; -1: none).
__device__ __forceinline__ void bvh_cast(const SceneDev& sc, double ox, double oy, double oz, double dx, double dy,
                                         double dz, double A, double tmin, double tmax, Best& best, int skip,
                                         uint32_t& n_exact, uint32_t& n_nodes, bool& overflow) {
    const RcpA dA = make_rcp(A);
    // (An FP32 line test per leaf sphere before the FP64 test was measured: exact tests/cast 4.96 -> 1.21,
    //  but 6 % slower overall -- the per-cast cull constants cost more than the FP64 tests they save.)
    const float kUp = 1.0f + 1.9073486328125e-06f;  // 1 + 2^-19
    const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
    const float padx = fabsf(fx) * 2.384185791015625e-07f + 1e-37f, pady = fabsf(fy) * 2.384185791015625e-07f + 1e-37f,
                padz = fabsf(fz) * 2.384185791015625e-07f + 1e-37f;  // 2^-22 |o|: covers the FP32 rounding of o
    const float opx = __fadd_ru(fx, padx), opy = __fadd_ru(fy, pady), opz = __fadd_ru(fz, padz);
    const float omx = __fadd_rd(fx, -padx), omy = __fadd_rd(fy, -pady), omz = __fadd_rd(fz, -padz);
    // 1/d in FP32 (__frcp_rn of the rounded component): 2^-23 relative instead of the 2^-24 of rounding the FP64
    // quotient -- the products below then carry 4 x 2^-24, inside the 1 +- 2^-19 slack -- and no FP64 division
    // A component that is exactly 0 makes the ray parallel to that slab pair: the axis then constrains no distance,
    // it only decides (below, per box) whether the origin's padded interval overlaps the slab at all.  With iv = 0
    // and constants -inf / +inf both plane distances come out as -inf / +inf without any inf * 0.  (iv = inf would
    // give fma(plane, inf, -+inf) = NaN for the plane on the far side of the origin, which fmaxf drops: a false
    // miss.)  A non-zero component whose FP32 value is zero or denormal (|d| < 2^-126) is not static -- the ray does
    // cross that slab at some huge t -- and no FP32 reciprocal can bound 1/d: such rays take the sequential scan.
    const bool zx = dx == 0.0, zy = dy == 0.0, zz = dz == 0.0;
    const bool zany = zx || zy || zz;
    {
        const float kMinNormal = 1.17549435e-38f;
        if ((!zx && fabsf((float)dx) < kMinNormal) || (!zy && fabsf((float)dy) < kMinNormal) ||
            (!zz && fabsf((float)dz) < kMinNormal)) {
            overflow = true;
            return;
        }
    }
    const float ivx = zx ? 0.f : __frcp_rn((float)dx), ivy = zy ? 0.f : __frcp_rn((float)dy),
                ivz = zz ? 0.f : __frcp_rn((float)dz);
    // Slab distances as ONE FFMA per plane: (plane - o) * iv = fma(plane, iv, -(o * iv)).  The constant -(o*iv) is
    // formed exactly in FP64 (24 x 24 bits) and rounded in the direction that keeps the plane's role conservative:
    // for iv > 0 the lo planes give the near distance (must not be overestimated: round down) and the hi planes the
    // far distance (round up); for iv < 0 the roles swap.  The FFMA's own rounding is relative to its result and is
    // covered by the 1 +- 2^-19 factors like before.  (iv is never +-inf here: zero components are static axes,
    // FP32-denormal ones left above.  iv = +-0 -- a component beyond the float range -- gives both planes the
    // distance 0, which only asks that the origin lie inside the other two slabs: conservative.)
    const double pxl = -((double)opx * (double)ivx), pyl = -((double)opy * (double)ivy), pzl = -((double)opz * (double)ivz);
    const double pxh = -((double)omx * (double)ivx), pyh = -((double)omy * (double)ivy), pzh = -((double)omz * (double)ivz);
    const float kFInf = __int_as_float(0x7f800000);
    const float clx = zx ? -kFInf : (ivx > 0.f ? __double2float_rd(pxl) : __double2float_ru(pxl)), chx = zx ? kFInf : (ivx > 0.f ? __double2float_ru(pxh) : __double2float_rd(pxh));
    const float cly = zy ? -kFInf : (ivy > 0.f ? __double2float_rd(pyl) : __double2float_ru(pyl)), chy = zy ? kFInf : (ivy > 0.f ? __double2float_ru(pyh) : __double2float_rd(pyh));
    const float clz = zz ? -kFInf : (ivz > 0.f ? __double2float_rd(pzl) : __double2float_ru(pzl)), chz = zz ? kFInf : (ivz > 0.f ? __double2float_ru(pzh) : __double2float_rd(pzh));
    // the pruning bound: entry distance tn may matter iff tn * (1 - 2^-19) <= RU(best t); kept pre-divided (rounded
    // up, so the test only gets looser) so that a box costs one compare against it
    const float kInvDn = 1.0f + 3.814697265625e-06f;  // > 1 / (1 - 2^-19)
    float bu = __fmul_ru(__double2float_ru(best.k >= 0 ? best.t : tmax), kInvDn);
    int stack_n[kBvhStack];
    float stack_t[kBvhStack];
    int sp = 0, node = 0;
    for (;;) {
        // one wide node (32 bytes per child), structure of arrays: the same bound of four children per float4
        constexpr int W = kBvhW, Q = kBvhW / 4;
        const float4* nb = sc.bvh_nodes + 8 * Q * node;
        float lx[W], ly[W], lz[W], hx[W], hy[W], hz[W];
        int ch[W];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float4 a0 = __ldg(nb + 0 * Q + q), a1 = __ldg(nb + 1 * Q + q), a2 = __ldg(nb + 2 * Q + q);
            const float4 a3 = __ldg(nb + 3 * Q + q), a4 = __ldg(nb + 4 * Q + q), a5 = __ldg(nb + 5 * Q + q);
            const float4 a6 = __ldg(nb + 6 * Q + q);
            lx[4 * q] = a0.x; lx[4 * q + 1] = a0.y; lx[4 * q + 2] = a0.z; lx[4 * q + 3] = a0.w;
            ly[4 * q] = a1.x; ly[4 * q + 1] = a1.y; ly[4 * q + 2] = a1.z; ly[4 * q + 3] = a1.w;
            lz[4 * q] = a2.x; lz[4 * q + 1] = a2.y; lz[4 * q + 2] = a2.z; lz[4 * q + 3] = a2.w;
            hx[4 * q] = a3.x; hx[4 * q + 1] = a3.y; hx[4 * q + 2] = a3.z; hx[4 * q + 3] = a3.w;
            hy[4 * q] = a4.x; hy[4 * q + 1] = a4.y; hy[4 * q + 2] = a4.z; hy[4 * q + 3] = a4.w;
            hz[4 * q] = a5.x; hz[4 * q + 1] = a5.y; hz[4 * q + 2] = a5.z; hz[4 * q + 3] = a5.w;
            ch[4 * q] = __float_as_int(a6.x); ch[4 * q + 1] = __float_as_int(a6.y);
            ch[4 * q + 2] = __float_as_int(a6.z); ch[4 * q + 3] = __float_as_int(a6.w);
        }
        ++n_nodes;
        float tn[W];
        bool hit[W];
#pragma unroll
        for (int i = 0; i < W; ++i) {
            // (the 24 slab FMAs as 12 FFMA2 -- child pairs from the float4 quads, 1/d and -o/d broadcast -- measured: 0..-3 %,
            //  profiles/r2_ab_wave2.txt)
            const float ax = fmaf(lx[i], ivx, clx), bx = fmaf(hx[i], ivx, chx);
            const float ay = fmaf(ly[i], ivy, cly), by = fmaf(hy[i], ivy, chy);
            const float az = fmaf(lz[i], ivz, clz), bz = fmaf(hz[i], ivz, chz);
            tn[i] = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            hit[i] = tn[i] <= tf * kUp && tn[i] <= bu;   // (empty slots carry NaN boxes: tf = NaN, never hit)
        }
        if (zany) {  // static axes: the (padded) origin coordinate must lie inside the slab
#pragma unroll
            for (int i = 0; i < W; ++i) {
                if (zx) hit[i] = hit[i] && lx[i] <= opx && hx[i] >= omx;
                if (zy) hit[i] = hit[i] && ly[i] <= opy && hy[i] >= omy;
                if (zz) hit[i] = hit[i] && lz[i] <= opz && hz[i] >= omz;
            }
        }
        // hit leaves are resolved on the spot, one after the other in ONE rolled loop (each re-checked against the best
        // found so far): a single copy of the FP64 test in the instruction stream -- with one copy per child slot the
        // kernel outgrew the 32 KB instruction cache and stalled on instruction fetch
        uint32_t leafm = 0u;
#pragma unroll
        for (int i = 0; i < W; ++i) leafm |= (hit[i] && ch[i] < 0) ? (1u << i) : 0u;
#pragma unroll 1
        while (leafm) {
            const int i = __ffs((int)leafm) - 1;
            leafm &= leafm - 1u;
            const float tni = i == 0 ? tn[0] : (i == 1 ? tn[1] : (i == 2 ? tn[2] : tn[3]));
            const int chi = i == 0 ? ch[0] : (i == 1 ? ch[1] : (i == 2 ? ch[2] : ch[3]));
            const int k = (int)(((unsigned)chi & 0x7fffffffu) >> 3);   // single-sphere leaves carry the sphere index itself (rt_bvh.h)
            if (tni <= bu && k != skip) {
                ++n_exact;
                if (exact_test_unordered(sc.exact, k, ox, oy, oz, dx, dy, dz, dA, tmin, tmax, best))
                    bu = __fmul_ru(__double2float_ru(best.t), kInvDn);
            }
        }
        // hit inner children: descend into the nearest, stack the others
        int next = -1;
        float next_t = 0.f;
#pragma unroll
        for (int i = 0; i < W; ++i) {
            if (hit[i] && ch[i] >= 0 && tn[i] <= bu) {
                int pn = ch[i];
                float pt = tn[i];
                if (next < 0 || pt < next_t) {   // new nearest: the old one (if any) goes to the stack
                    const int on = next; const float ot = next_t;
                    next = pn; next_t = pt;
                    pn = on; pt = ot;
                }
                if (pn >= 0) {
                    if (sp >= kBvhStack) { overflow = true; return; }
                    stack_n[sp] = pn; stack_t[sp] = pt; ++sp;
                }
            }
        }
        if (next >= 0) {
            node = next;
        } else {
            bool found = false;
            while (sp > 0) {
                --sp;
                if (stack_t[sp] <= bu) { node = stack_n[sp]; found = true; break; }
            }
            if (!found) break;
        }
    }
}

// ---------------------------------------------------------------- start-sphere test + tie grid (rt_bvh.h: TieGridHost)
// A cast that starts on the sphere the path just hit (`self`) and hits it again at t ~ 0 -- 86 % of the reference's
// casts on the book scene, the tmin = 0 artefact of programs/main.cc:40 -- is decided here without a traversal:
// another sphere j can only replace that hit if it yields an accepted root <= t (programs/hittable_list.cc:11-15),
// which needs its surface within reach = t * |dir| of the origin.  Conservative FP32 shell test
//     | |o - c_j|^2 - r_j^2 |  <=  2^-19 (|o|^2 + |c_j|^2 + r_j^2)  +  reach (2 r_j + reach)
// (the first term covers the FP32 evaluation error 2^-24 (14 (o^2 + c^2) + 4 r^2) with a factor 2 to spare and,
// by 30 binary orders, the band in which sphere::hit can round a root to 0) on the giants and on the <= 4 spheres
// listed in the grid cell of o; those that pass run the FP64 sphere::hit and are merged with the reference's tie rule.
// Returns false if the cast cannot be decided here (reach too long, overfull cell): the caller traverses with
// `best` as the initial bound.
__device__ __forceinline__ bool tie_candidate(const float4 s, float fx, float fy, float fz, float o2, float rho) {
    const float ex = fx - s.x, ey = fy - s.y, ez = fz - s.z;
    const float q = fmaf(-s.w, s.w, fmaf(ex, ex, fmaf(ey, ey, ez * ez)));
    const float w = fmaf(s.x, s.x, fmaf(s.y, s.y, fmaf(s.z, s.z, s.w * s.w)));
    const float tol = fmaf(1.9073486328125e-06f, o2 + w, rho * fmaf(2.0002f, s.w, rho));
    return !(fabsf(q) > tol);   // (NaN counts as a candidate)
}

// The part of a BVH-mode cast that needs no traversal: FP64 sphere::hit of the start sphere, then the tie grid.
// Returns true if `best` is the cast's final answer; false if the cast cannot be decided here (not a hit of the
// start sphere, reach too long, overfull cell): the caller traverses with `best` as the initial bound.  (tmax = +inf.)
// One rolled loop over a pending mask -- bit 0 is the start sphere, bits 1-4 the giants, bits 5-8 the cell's spheres --
// so that the FP64 test exists once in the instruction stream.  The first trip tests the start sphere and then runs the
// FP32 shell tests of ALL candidates in straight-line code (their loads are in flight together); only candidates that
// pass (0.3 % of the casts) set their bit and cost another trip.  (Round 2, first form: one trip per candidate, shell
// test or not -- 4-5 trips per warp and one dependent load each.)
__device__ __forceinline__ bool self_cast_flat(const SceneDev& sc, int self, double ox, double oy, double oz, double dx,
                                          double dy, double dz, double A, double tmin, Best& best, uint32_t& n_exact) {
    if (self < 0) return false;
    const RcpA dA = make_rcp(A);
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    bool decided = false;
    uint32_t pend = 1u;
    int4 cell = make_int4(-1, -1, -1, -1);
#pragma unroll 1
    while (pend) {
        const int e = __ffs((int)pend) - 1;
        pend &= pend - 1u;
        int j = self;
        if (e > 0) j = e <= 4 ? sc.tie_giants[e - 1] : (e == 5 ? cell.x : (e == 6 ? cell.y : (e == 7 ? cell.z : cell.w)));
        ++n_exact;
        exact_test_unordered(sc.exact, j, ox, oy, oz, dx, dy, dz, dA, tmin, kInf, best);
        if (e == 0) {
            if (!(sc.tie_ok && best.k == self && dA.fast)) break;
            float rho = 0.f;
            if (best.t != 0.0) {   // reach = t * |dir|, rounded up
                rho = __fmul_ru(__fmul_ru(__double2float_ru(best.t), __fsqrt_ru(__double2float_ru(A))), 1.0000019f);
                if (!(rho <= sc.tie_rho_max)) break;   // (also NaN / negative t)
            }
            const float fx = (float)ox, fy = (float)oy, fz = (float)oz;
            const float o2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
            const bool inside = fx >= sc.tie_g0[0] && fx <= sc.tie_g1[0] && fy >= sc.tie_g0[1] && fy <= sc.tie_g1[1] &&
                                fz >= sc.tie_g0[2] && fz <= sc.tie_g1[2];
            if (inside) {   // outside the grid no listed sphere has its (padded) box around o
                const int ix = (int)((fx - sc.tie_g0[0]) * sc.tie_inv_h), iy = (int)((fy - sc.tie_g0[1]) * sc.tie_inv_h),
                          iz = (int)((fz - sc.tie_g0[2]) * sc.tie_inv_h);
                cell = __ldg(sc.tie_cells + ((size_t)iz * sc.tie_dimy + iy) * sc.tie_dimx + ix);
                if (cell.x == -2) break;   // overfull cell
            }
            // the cell's spheres (entries fill from .x up; -1: unused): loads first, they fly while the giants are tested
            const bool c0 = cell.x >= 0 && cell.x != self, c1 = cell.y >= 0 && cell.y != self, c2 = cell.z >= 0 && cell.z != self,
                       c3 = cell.w >= 0 && cell.w != self;
            const float4 never = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);   // (|q| = inf: no candidate)
            const float4 s0 = c0 ? __ldg(sc.sph32 + cell.x) : never, s1 = c1 ? __ldg(sc.sph32 + cell.y) : never,
                         s2 = c2 ? __ldg(sc.sph32 + cell.z) : never, s3 = c3 ? __ldg(sc.sph32 + cell.w) : never;
#pragma unroll 1
            for (int g = 0; g < sc.tie_ngiants; ++g) {   // (warp-uniform trip count)
                const int jg = sc.tie_giants[g];
                if (jg != self && tie_candidate(__ldg(sc.sph32 + jg), fx, fy, fz, o2, rho)) pend |= 2u << g;
            }
            if (c0 && tie_candidate(s0, fx, fy, fz, o2, rho)) pend |= 32u;
            if (c1 && tie_candidate(s1, fx, fy, fz, o2, rho)) pend |= 64u;
            if (c2 && tie_candidate(s2, fx, fy, fz, o2, rho)) pend |= 128u;
            if (c3 && tie_candidate(s3, fx, fy, fz, o2, rho)) pend |= 256u;
            decided = true;
        }
    }
    return decided;
}

// The same step with one trip per candidate -- entry 0 is the start sphere, then the giants, then the cell's spheres,
// each with its own load and shell test.  Fewer instructions per trip, dependent loads: the faster form while the FP32
// sphere array stays in L1 (small scenes: C3 with the early-out +3.5 %, the default scene +7 %), the slower one when it
// does not (C4: -7 %); profiles/r2_ab_wave2.txt.  The wavefront kernel is instantiated with both (kTieFlat).
__device__ __forceinline__ bool self_cast_loop(const SceneDev& sc, int self, double ox, double oy, double oz, double dx,
                                          double dy, double dz, double A, double tmin, Best& best, uint32_t& n_exact) {
    if (self < 0) return false;
    const RcpA dA = make_rcp(A);
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    bool decided = false;
    int ncand = 1;
    int4 cell = make_int4(-1, -1, -1, -1);
    float fx = 0.f, fy = 0.f, fz = 0.f, o2 = 0.f, rho = 0.f;
#pragma unroll 1
    for (int e = 0; e < ncand; ++e) {
        int j = self;
        bool test = true;
        if (e > 0) {
            const int g = e - 1 - sc.tie_ngiants;
            if (g < 0) j = sc.tie_giants[e - 1];
            else j = g == 0 ? cell.x : (g == 1 ? cell.y : (g == 2 ? cell.z : cell.w));
            test = j != self && tie_candidate(__ldg(sc.sph32 + j), fx, fy, fz, o2, rho);
        }
        if (test) {
            ++n_exact;
            exact_test_unordered(sc.exact, j, ox, oy, oz, dx, dy, dz, dA, tmin, kInf, best);
        }
        if (e == 0) {
            if (!(sc.tie_ok && best.k == self && dA.fast)) break;
            if (best.t != 0.0) {   // reach = t * |dir|, rounded up
                rho = __fmul_ru(__fmul_ru(__double2float_ru(best.t), __fsqrt_ru(__double2float_ru(A))), 1.0000019f);
                if (!(rho <= sc.tie_rho_max)) break;   // (also NaN / negative t)
            }
            fx = (float)ox; fy = (float)oy; fz = (float)oz;
            o2 = fmaf(fx, fx, fmaf(fy, fy, fz * fz));
            int ncell = 0;
            const bool inside = fx >= sc.tie_g0[0] && fx <= sc.tie_g1[0] && fy >= sc.tie_g0[1] && fy <= sc.tie_g1[1] &&
                                fz >= sc.tie_g0[2] && fz <= sc.tie_g1[2];
            if (inside) {   // outside the grid no listed sphere has its (padded) box around o
                const int ix = (int)((fx - sc.tie_g0[0]) * sc.tie_inv_h), iy = (int)((fy - sc.tie_g0[1]) * sc.tie_inv_h),
                          iz = (int)((fz - sc.tie_g0[2]) * sc.tie_inv_h);
                cell = __ldg(sc.tie_cells + ((size_t)iz * sc.tie_dimy + iy) * sc.tie_dimx + ix);
                if (cell.x == -2) break;   // overfull cell
                ncell = (cell.x >= 0) + (cell.y >= 0) + (cell.z >= 0) + (cell.w >= 0);   // (entries fill from .x up)
            }
            ncand = 1 + sc.tie_ngiants + ncell;
            decided = true;
        }
    }
    return decided;
}

template <bool kFlat>
__device__ __forceinline__ bool self_cast(const SceneDev& sc, int self, double ox, double oy, double oz, double dx, double dy,
                                          double dz, double A, double tmin, Best& best, uint32_t& n_exact) {
    return kFlat ? self_cast_flat(sc, self, ox, oy, oz, dx, dy, dz, A, tmin, best, n_exact)
                 : self_cast_loop(sc, self, ox, oy, oz, dx, dy, dz, A, tmin, best, n_exact);
}

// hittable_list::hit for one ray given its survivors (or the full list when ovf): list order, shrinking tmax.
__device__ __forceinline__ Best resolve_hits(const SceneDev& sc, bool ovf, int cnt, const uint16_t* cand, int cand_step,
                                             double ox, double oy, double oz, double dx, double dy, double dz, double A,
                                             double tmin, double tmax, uint32_t& n_exact) {
    Best best;
    best.t = tmax; best.C = 1.0; best.k = -1;
    const int n_iter = ovf ? sc.n : cnt;  // candidate list, or every sphere when the list overflowed
    const RcpA dA = make_rcp(A);
#pragma unroll 1
    for (int e = 0; e < n_iter; ++e) {
        const int k = ovf ? e : (int)cand[e * cand_step];
        if (k < sc.n) exact_test(sc.exact, k, ox, oy, oz, dx, dy, dz, dA, tmin, best);
    }
    n_exact += (uint32_t)n_iter;
    return best;
}

// One rejection try of programs/vec3.h:87-94 from three 21-bit uniforms f: v = -1 + 2*(f / 2^21) (random.h:10-14)
// is exact in FP64, and so are v*v and the sum of the three squares (44 significant bits), so the reference's
// test len^2 > 1 (vec3.h:90) is decided EXACTLY by integers: with s = f - 2^20, len^2 <= 1 <=> sum s^2 <= 2^40.
// The tries therefore run on the integer pipe (three IMAD.WIDE, two 64-bit adds, one compare) instead of 17
// half-rate FP64 instructions each, and only the accepted point is converted to FP64.
__device__ __forceinline__ bool in_unit_ball(uint32_t fx, uint32_t fy, uint32_t fz) {
    const int sx = (int)fx - (1 << 20), sy = (int)fy - (1 << 20), sz = (int)fz - (1 << 20);
    const unsigned long long l2 = (unsigned long long)((long long)sx * sx) + (unsigned long long)((long long)sy * sy) +
                                  (unsigned long long)((long long)sz * sz);
    return l2 <= (1ull << 40);
}
__device__ __forceinline__ double cube_coord(uint32_t f) {  // -1 + 2 * (f * 2^-21), every step exact
    return dadd(-1.0, dmul(2.0, dmul((double)f, 1.0 / 2097152.0)));
}

// programs/vec3.h:83-109 random_in_hemisphere.  One Philox block per bounce: it carries the first TWO tries of
// the rejection loop as six 21-bit uniforms (the reference's rand() has 15 bits): try A = top 21 bits of words
// 0,1,2; try B = the low 11 bits of words 0,1,2 extended by 10-bit fields of word 3.  The 23 % of bounces that
// reject both (a try is kept with probability pi/6) continue with xorshift128 (Marsaglia 2003) seeded by the
// block's four words, three outputs per try (top 21 bits each): a cheap continuation instead of further
// 10-round blocks, since the warp runs as many loop iterations as its unluckiest lane.
// lambertian != 0 replaces the hemisphere flip of vec3.h:102-109 by vec3::random_unit_vector (vec3.h:97-100):
// unit_vector(v) = (1 / v.length()) * v (vec3.h:172-175, 151-154), same rejection tries.
__device__ __forceinline__ void random_scatter(uint32_t pix, uint32_t smp, uint32_t& blk, uint32_t k0, uint32_t k1,
                                               double nx, double ny, double nz, int lambertian, double& rx, double& ry,
                                               double& rz) {
    const uint4 w = philox4x32_10(pix, smp, blk, 0u, k0, k1);
    ++blk;
    uint32_t fx = w.x >> 11, fy = w.y >> 11, fz = w.z >> 11;
    bool ok = in_unit_ball(fx, fy, fz);
    {
        const uint32_t gx = ((w.x & 0x7ffu) << 10) | (w.w >> 22), gy = ((w.y & 0x7ffu) << 10) | ((w.w >> 12) & 0x3ffu),
                       gz = ((w.z & 0x7ffu) << 10) | ((w.w >> 2) & 0x3ffu);
        const bool okb = in_unit_ball(gx, gy, gz);
        fx = ok ? fx : gx; fy = ok ? fy : gy; fz = ok ? fz : gz;
        ok = ok || okb;
    }
    if (!ok) {
        uint32_t x0 = w.x, x1 = w.y, x2 = w.z, x3 = w.w;
        if ((x0 | x1 | x2 | x3) == 0u) x0 = 1u;
        do {
            uint32_t f[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                uint32_t t = x3;
                const uint32_t s0 = x0;
                x3 = x2; x2 = x1; x1 = s0;
                t ^= t << 11; t ^= t >> 8;
                x0 = t ^ s0 ^ (s0 >> 19);
                f[i] = x0 >> 11;
            }
            fx = f[0]; fy = f[1]; fz = f[2];
            ok = in_unit_ball(fx, fy, fz);
        } while (!ok);
    }
    rx = cube_coord(fx); ry = cube_coord(fy); rz = cube_coord(fz);
    if (lambertian) {
        const double inv_len = ddiv(1.0, dsqrt(ddot(rx, ry, rz, rx, ry, rz)));   // programs/vec3.h:172-175
        rx = dmul(inv_len, rx); ry = dmul(inv_len, ry); rz = dmul(inv_len, rz);
    } else if (!(ddot(rx, ry, rz, nx, ny, nz) > 0)) { rx = -rx; ry = -ry; rz = -rz; }  // programs/vec3.h:105-108
}

// programs/main.cc:46-48 sky colour of a ray that missed, times the attenuation of its `bounces` hits.  The
// recursion returns albedo * (albedo * (... * sky)) (main.cc:43): with the reference's albedo 0.5 every product
// is exact and equals 0.5^bounces * sky (no denormals: bounces <= 1001 and the reference's sky is >= 0.5); with
// custom shading the multiplications are applied one by one, innermost first, as the recursion unwinds.
__device__ __forceinline__ void sky_color(const ShadeDev& sh, double dy, double A, int bounces, double& r, double& g,
                                          double& b) {
    const double inv_len = ddiv(1.0, dsqrt(A));  // unit_vector = v / v.length() = (1/len) * v
    const double uy = dmul(inv_len, dy);
    const double t = dmul(0.5, dadd(uy, 1.0));
    const double omt = dsub(1.0, t);
    if (!sh.custom) {
        const double att = __longlong_as_double((long long)(1023 - bounces) << 52);  // 0.5^bounces, exact
        r = dmul(att, dadd(dmul(omt, 1.0), dmul(t, 0.5)));
        g = dmul(att, dadd(dmul(omt, 1.0), dmul(t, 0.7)));
        b = dmul(att, dadd(dmul(omt, 1.0), dmul(t, 1.0)));
    } else {
        r = dadd(dmul(omt, sh.sky_a[0]), dmul(t, sh.sky_b[0]));
        g = dadd(dmul(omt, sh.sky_a[1]), dmul(t, sh.sky_b[1]));
        b = dadd(dmul(omt, sh.sky_a[2]), dmul(t, sh.sky_b[2]));
#pragma unroll 1
        for (int k = 0; k < bounces; ++k) { r = dmul(sh.albedo, r); g = dmul(sh.albedo, g); b = dmul(sh.albedo, b); }
    }
}

// programs/color.h:16-23 for one channel
__device__ __forceinline__ int write_color_channel(double sum, double one_over_samples) {
    double x = dsqrt(dmul(sum, one_over_samples));
    x = (x < 0.0) ? 0.0 : x;        // std::max(x, 0.0)
    x = (0.999 < x) ? 0.999 : x;    // std::min(., 0.999)
    return (int)dmul(255.999, x);
}

// programs/camera.h:25-28
__device__ __forceinline__ void camera_ray(const double* org, const double* llc, const double* hor, const double* ver,
                                           double u, double v, double& dx, double& dy, double& dz) {
    dx = dsub(dadd(dadd(llc[0], dmul(u, hor[0])), dmul(v, ver[0])), org[0]);
    dy = dsub(dadd(dadd(llc[1], dmul(u, hor[1])), dmul(v, ver[1])), org[1]);
    dz = dsub(dadd(dadd(llc[2], dmul(u, hor[2])), dmul(v, ver[2])), org[2]);
}

}  // namespace rt
