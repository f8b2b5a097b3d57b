// rt_api.cu -- the C ABI of include/rt.h over the sm_100a kernels in rt_kernels.cuh.
// Host side only flattens, uploads, launches and copies; there is no CPU implementation of the path.
#include "../../include/rt.h"
#include "rt_bvh.h"
#include "rt_kernels.cuh"
#include "rt_units.h"

#include <algorithm>
#include <cmath>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define RT_CUDA(expr)                                                                                       \
    do {                                                                                                    \
        cudaError_t e__ = (expr);                                                                           \
        if (e__ != cudaSuccess)                                                                             \
            return fail(RT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
    } while (0)

struct DeviceGuard {  // the caller (e.g. torch) keeps its own current device
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <typename T>
struct DevBuf {  // scoped device allocation
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};

}  // namespace

// Device scratch one render launch works in.  A scene owns one; a progressive accumulator owns a second one so that
// consecutive passes can overlap on two streams (the tail of pass k hides behind the start of pass k+1).
struct Scratch {
    unsigned int* d_tile_counter = nullptr;   // work-queue head of the persistent warps
    unsigned long long* d_stats = nullptr;
    void* d_accum = nullptr; size_t accum_cap = 0;   // fixed-point radiance per tile pixel (+ chunk counters behind it)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaError_t create() {
        cudaError_t e;
        if ((e = cudaMalloc(&d_tile_counter, sizeof(unsigned int))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&d_stats, 16 * sizeof(unsigned long long))) != cudaSuccess) return e;
        if ((e = cudaEventCreate(&ev0)) != cudaSuccess) return e;
        return cudaEventCreate(&ev1);
    }
    void destroy() {
        cudaFree(d_tile_counter); cudaFree(d_stats); cudaFree(d_accum);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        *this = Scratch();
    }
};

struct rt_scene {
    int device = 0;
    int n = 0, npad = 0;
    int sm_count = 0;
    bool cull_ok = true;   // every sphere finite and of moderate magnitude: the FP32 cull is usable
    uint64_t filt_generation = 0;     // version of d_filt (the constant bank caches the last one it was given)
    float4* d_filt = nullptr;
    float4* d_filt_pk = nullptr;      // the same entries in pairs (rt_device.cuh: cull_scan_packed), what the constant bank holds
    double4* d_exact = nullptr;
    double* d_inv_r = nullptr;        // RN(1/r) per sphere
    float4* d_bvh_nodes = nullptr;    // flattened BVH (rt_bvh.h), built at upload
    int32_t* d_bvh_leaf = nullptr;
    int bvh_nodes = 0;
    rt::BvhHost bvh_host;             // topology kept for rt_update_scene's refit
    rt::TieGridHost tie;              // tie grid of the current sphere positions (rt_bvh.h)
    float4* d_sph32 = nullptr;        // FP32 {centre, |r|} per sphere (tie grid shell tests)
    int4* d_tie_cells = nullptr; size_t tie_cells_cap = 0;
    Scratch scr;                      // per-launch device scratch of this scene's renders
    // scratch for the host-buffer entry points (grow only)
    void* d_frame = nullptr; size_t frame_cap = 0;
    void* d_sum = nullptr;   size_t sum_cap = 0;
    // last asynchronous render
    bool pending = false;
    cudaStream_t last_stream = nullptr;
    int last_mode = 0;
    uint64_t last_launches = 0;
};

struct rt_accum {
    int device = 0;
    int W = 0, H = 0;
    int samples = 0;                          // samples per pixel accumulated so far
    unsigned long long* d_sums = nullptr;     // W*H*3, 20.44 fixed point, frame pixel order
    uchar4* d_frame = nullptr;                // write_color over all samples so far (made on demand by rt_accum_frame)
    Scratch scr[2];                           // passes alternate between two scratch sets / streams
    cudaStream_t stream[2] = {nullptr, nullptr};
    int passes = 0;
    const rt_scene* last_scene[2] = {nullptr, nullptr};
    int last_mode[2] = {0, 0};
};

namespace {

// The cull array of the scene being rendered lives in the (per-device) constant bank.  Renders issued on
// different streams of one device are therefore chained: the copy into the bank waits for the previous
// constant-bank render on that device.
struct ConstBankGuard {
    std::mutex mu;
    struct Dev {
        const void* owner = nullptr;     // scene whose cull array the bank holds ...
        uint64_t generation = 0;         // ... and which version of it (rt_update_scene bumps it)
        cudaEvent_t copied = nullptr;    // ... recorded behind the copy into the bank
        cudaStream_t copy_stream = nullptr;
        std::vector<std::pair<cudaStream_t, cudaEvent_t>> users;   // last constant-bank render per stream
    } dev[64];
} g_const_bank;

float round_down_f32(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -INFINITY);
    return f;
}

int check_params(const rt_params* p) {
    if (!p) return fail(RT_ERR_INVALID, "params is NULL");
    if (p->width < 2 || p->height < 2) return fail(RT_ERR_INVALID, "width and height must be >= 2 (u,v divide by W-1, H-1)");
    if ((int64_t)p->width * p->height > (int64_t)1 << 30) return fail(RT_ERR_INVALID, "frame too large");
    if (p->spp < 1 || p->spp > (1 << 20)) return fail(RT_ERR_INVALID, "spp must be in [1, 2^20] (radiance sums are 20.44 fixed point)");
    if (p->max_depth > 1000) return fail(RT_ERR_INVALID, "max_depth must be <= 1000");
    if (p->shard_count < 1 || p->shard_rank < 0 || p->shard_rank >= p->shard_count)
        return fail(RT_ERR_INVALID, "bad shard_rank / shard_count");
    if (p->scan_mode < 0 || p->scan_mode > RT_SCAN_AUTO) return fail(RT_ERR_INVALID, "bad scan_mode");
    if (p->custom_shading) {
        if (p->scatter_mode != RT_SCATTER_HEMISPHERE && p->scatter_mode != RT_SCATTER_LAMBERTIAN)
            return fail(RT_ERR_INVALID, "bad scatter_mode");
        if (!(p->albedo >= 0.0 && p->albedo <= 1.0)) return fail(RT_ERR_INVALID, "albedo must be in [0, 1]");
        for (int c = 0; c < 3; ++c)
            if (!(p->sky_a[c] >= 0.0 && p->sky_a[c] <= 1.0) || !(p->sky_b[c] >= 0.0 && p->sky_b[c] <= 1.0))
                return fail(RT_ERR_INVALID, "sky colour components must be in [0, 1] (radiance sums are 20.44 fixed point)");
    }
    return RT_OK;
}

rt::ShadeDev shade_dev(const rt_params* p) {
    rt::ShadeDev sh;
    std::memset(&sh, 0, sizeof sh);
    sh.custom = p->custom_shading ? 1 : 0;
    sh.albedo = sh.custom ? p->albedo : 0.5;                                       // programs/main.cc:43
    const double a[3] = {1.0, 1.0, 1.0}, b[3] = {0.5, 0.7, 1.0};                   // programs/main.cc:48
    for (int c = 0; c < 3; ++c) { sh.sky_a[c] = sh.custom ? p->sky_a[c] : a[c]; sh.sky_b[c] = sh.custom ? p->sky_b[c] : b[c]; }
    sh.lambertian = (sh.custom && p->scatter_mode == RT_SCATTER_LAMBERTIAN) ? 1 : 0;
    return sh;
}

void tile_layout(const rt_params* p, rt_tile_layout* L) {
    L->tile_w = rt::kTileW; L->tile_h = rt::kTileH;
    L->tiles_x = (p->width + rt::kTileW - 1) / rt::kTileW;
    L->tiles_y = (p->height + rt::kTileH - 1) / rt::kTileH;
    L->tiles_total = L->tiles_x * L->tiles_y;
    L->tiles_per_shard = (L->tiles_total + p->shard_count - 1) / p->shard_count;
    L->shard_bytes = (int64_t)L->tiles_per_shard * rt::kTilePix * 4;
}

// AUTO: the BVH mode (wavefront kernel: start-sphere test + tie grid, traversal for the rest) wherever a tree was built.
// Round 1 kept the linear cull scan below 16 spheres (the two tied there); with the round-2 kernel the BVH mode wins
// from the reference's own two-sphere scene up (1084 vs 1036 Msamples/s, profiles/r2_ab_wave.txt) and by 3.6x at 485
// spheres (778 vs 215), so AUTO takes it for every scene of two or more spheres.
int resolve_scan_mode(const rt_scene* sc, int mode, int* out) {
    if (mode == RT_SCAN_AUTO) mode = (!sc->cull_ok) ? RT_SCAN_EXACT : (sc->n < 2 ? RT_SCAN_FILTERED : RT_SCAN_BVH);
    if (mode == RT_SCAN_FILTERED && !sc->cull_ok)
        return fail(RT_ERR_UNSUPPORTED, "scene has non-finite or huge (>1e15) coordinates: use RT_SCAN_EXACT");
    if (mode == RT_SCAN_BVH && !sc->d_bvh_nodes)
        return fail(RT_ERR_UNSUPPORTED, "no BVH for this scene (non-finite or huge coordinates): use RT_SCAN_EXACT");
    if (mode == RT_SCAN_FILTERED && sc->npad > rt::kMaxLinearSmem)
        return fail(RT_ERR_UNSUPPORTED, "linear cull scan holds at most 11520 spheres (4080 in the constant bank, beyond that one CTA's "
                                        "shared memory); use RT_SCAN_EXACT or the BVH");
    *out = mode;
    return RT_OK;
}

rt::SceneDev scene_dev(const rt_scene* sc, int mode) {
    rt::SceneDev d;
    d.filt = sc->d_filt; d.exact = sc->d_exact; d.inv_r = sc->d_inv_r; d.n = sc->n;
    d.npad = (mode == RT_SCAN_FILTERED) ? sc->npad : 0;  // EXACT / BVH stage nothing
    d.bvh_nodes = sc->d_bvh_nodes; d.bvh_leaf = sc->d_bvh_leaf;
    const rt::TieGridHost& g = sc->tie;
    d.sph32 = sc->d_sph32; d.tie_cells = sc->d_tie_cells;
    d.tie_ok = (g.ok && sc->d_sph32 && sc->d_tie_cells) ? 1 : 0;
    for (int a = 0; a < 3; ++a) { d.tie_g0[a] = g.g0[a]; d.tie_g1[a] = g.g1[a]; }
    d.tie_inv_h = g.inv_h; d.tie_rho_max = g.rho_max;
    d.tie_dimx = g.dim[0]; d.tie_dimy = g.dim[1];
    d.tie_ngiants = g.n_giants;
    for (int i = 0; i < 4; ++i) d.tie_giants[i] = g.giants[i];
    return d;
}

int grow(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return RT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    RT_CUDA(cudaMalloc(p, need));
    *cap = need;
    return RT_OK;
}

int launch_render(rt_scene* sc, const rt_camera* cam, const rt_params* p, void* d_rgba, void* d_sum,
                  cudaStream_t stream, int sample_base = 0, void* d_frame_accum = nullptr, bool frame_order_out = false,
                  Scratch* scratch = nullptr) {
    Scratch& X = scratch ? *scratch : sc->scr;
    int mode = 0;
    int rc = resolve_scan_mode(sc, p->scan_mode, &mode);
    if (rc) return rc;
    rt_tile_layout L;
    tile_layout(p, &L);
    rt::RenderArgs a;
    std::memset(&a, 0, sizeof a);
    a.sc = scene_dev(sc, mode);
    if (p->reserved[2] == 3) a.sc.tie_ok = 0;   // A/B knob: BVH mode without the tie-grid fast path
    for (int c = 0; c < 3; ++c) {
        a.cam_org[c] = cam->origin[c]; a.cam_llc[c] = cam->lower_left_corner[c];
        a.cam_hor[c] = cam->horizontal[c]; a.cam_ver[c] = cam->vertical[c];
    }
    a.tmin = p->tmin;
    a.sh = shade_dev(p);
    a.W = p->width; a.H = p->height; a.spp = p->spp; a.max_depth = p->max_depth;
    a.key0 = (uint32_t)p->seed; a.key1 = (uint32_t)(p->seed >> 32);
    a.jitter = p->jitter; a.scan_mode = mode;
    a.early_out = (p->early_out && p->tmin == 0.0) ? 1 : 0;  // the cut is only exact for tmin == 0
    a.tiles_x = L.tiles_x; a.tiles_total = L.tiles_total;
    a.shard_rank = p->shard_rank; a.shard_count = p->shard_count;
    a.tiles_local = (L.tiles_total - p->shard_rank + p->shard_count - 1) / p->shard_count;
    a.compact_out = p->shard_count > 1 && !frame_order_out;   // (rt_render_multi: shards store straight into one frame)
    a.out = (uchar4*)d_rgba; a.sum_out = (double*)d_sum;
    a.unit_counter = X.d_tile_counter; a.stats = X.d_stats;
    a.sample_base = sample_base; a.frame_accum = (unsigned long long*)d_frame_accum;

    int R = p->reserved[0];
    if (R == 0) R = 2;
    if (R != 1 && R != 2 && R != 4) return fail(RT_ERR_INVALID, "reserved[0] (paths per lane) must be 0, 1, 2 or 4");
    if (mode == RT_SCAN_BVH) R = 1;  // traversal is divergent: one path per lane, more warps
    // cull array source: constant bank (default; FFMAs then read a uniform-register operand) or the
    // TMA-staged shared-memory copy (reserved[2] == 1), kept for A/B evidence
    // (scenes beyond the 64 KB bank take the shared-memory variant: the only linear scan that can hold them)
    const bool use_const = p->reserved[2] != 1 && mode == RT_SCAN_FILTERED && sc->npad <= rt::kMaxLinear;
    // One critical section from the wait on the previous constant-bank renders through this render's launch and
    // event record: a second host thread rendering another FILTERED scene on another stream of this device then
    // queues its copy into the bank strictly behind this kernel (and not between this copy and this launch).  Renders
    // of the scene whose array the bank already holds need neither the copy nor the wait (they may overlap).
    std::unique_lock<std::mutex> bank_lock;
    if (use_const) {
        bank_lock = std::unique_lock<std::mutex>(g_const_bank.mu);
        ConstBankGuard::Dev& B = g_const_bank.dev[sc->device & 63];
        if (B.owner != sc || B.generation != sc->filt_generation) {
            for (auto& u : B.users) RT_CUDA(cudaStreamWaitEvent(stream, u.second, 0));
            RT_CUDA(cudaMemcpyToSymbolAsync(rt::c_filt, RT_SCAN_PACKED ? sc->d_filt_pk : sc->d_filt, (size_t)(sc->npad + rt::kScanPad) * sizeof(float4), 0,
                                            cudaMemcpyDeviceToDevice, stream));
            // (the events stay in the list: a later copy by another scene must still wait for those kernels; entries of
            //  this stream are overwritten below)
            if (!B.copied) RT_CUDA(cudaEventCreateWithFlags(&B.copied, cudaEventDisableTiming));
            RT_CUDA(cudaEventRecord(B.copied, stream));
            B.owner = sc; B.generation = sc->filt_generation; B.copy_stream = stream;
        } else if (stream != B.copy_stream) {
            RT_CUDA(cudaStreamWaitEvent(stream, B.copied, 0));   // the bank is this scene's once that copy has run
        }
    }
    rt::RenderSmem S = rt::render_smem(use_const ? 0 : a.sc.npad, R);
    int threads = rt::kThreads;
    void (*kern)(const rt::RenderArgs) = nullptr;
    if (mode == RT_SCAN_BVH && p->reserved[2] == 2) kern = rt::render_kernel<1, 2>;   // round-1 traversal kernel (A/B evidence only)
    else if (mode == RT_SCAN_BVH) {
        // tie-grid candidates in straight-line code once the FP32 sphere array outgrows L1 (rt_device.cuh: self_cast_flat)
        kern = sc->n > 4096 ? rt::render_wave_kernel<true> : rt::render_wave_kernel<false>;
        S.total = rt::wave_smem().total; threads = rt::kWaveThreads;
    }
    else if (use_const) kern = R == 1 ? rt::render_kernel<1, 1> : (R == 2 ? rt::render_kernel<2, 1> : rt::render_kernel<4, 1>);
    else kern = R == 1 ? rt::render_kernel<1, 0> : (R == 2 ? rt::render_kernel<2, 0> : rt::render_kernel<4, 0>);
    if (const char* e = std::getenv("RT_SMEM_PAD")) S.total += (uint32_t)std::atoi(e);   // tuning experiments only: fewer CTAs per SM
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
    int per_sm = 0;
    RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, S.total));
    if (per_sm < 1) return fail(RT_ERR_CUDA, "render kernel does not fit on an SM");
    const int full_grid = sc->sm_count * per_sm;  // persistent: every resident warp pulls work units

    // work units: graded sample chunks per tile (rt_units.h)
    const int warps_per_cta = threads / 32;
    const int warps = full_grid * warps_per_cta;
    const bool wave = mode == RT_SCAN_BVH && p->reserved[2] != 2;
    const rt::UnitPlan up = rt::plan_units(p->spp, a.tiles_local, warps, wave, p->reserved[1], d_frame_accum != nullptr);
    for (int l = 0; l < 3; ++l) { a.lv_n[l] = up.lv_n[l]; a.lv_spp[l] = up.lv_spp[l]; }
    a.chunks = rt::plan_chunks(up);
    a.units_local = a.tiles_local * a.chunks;
    int grid = full_grid;
    const int need = (a.units_local + warps_per_cta - 1) / warps_per_cta;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;

    if (a.chunks > 1) {
        const size_t acc_bytes = (size_t)a.tiles_local * rt::kTilePix * 3 * sizeof(unsigned long long);
        const size_t done_bytes = (size_t)a.tiles_local * sizeof(unsigned int);
        int rc2 = grow(&X.d_accum, &X.accum_cap, acc_bytes + done_bytes);
        if (rc2) return rc2;
        a.accum = (unsigned long long*)X.d_accum;
        a.tile_done = (unsigned int*)((char*)X.d_accum + acc_bytes);
        RT_CUDA(cudaMemsetAsync(X.d_accum, 0, acc_bytes + done_bytes, stream));
    }
    RT_CUDA(cudaMemsetAsync(X.d_tile_counter, 0, sizeof(unsigned int), stream));
    RT_CUDA(cudaMemsetAsync(X.d_stats, 0, rt::kNumStats * sizeof(unsigned long long), stream));
    if (a.compact_out && d_rgba) RT_CUDA(cudaMemsetAsync(d_rgba, 0, (size_t)L.shard_bytes, stream));
    RT_CUDA(cudaEventRecord(X.ev0, stream));
    kern<<<grid, threads, S.total, stream>>>(a);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaEventRecord(X.ev1, stream));
    if (use_const) {
        ConstBankGuard::Dev& B = g_const_bank.dev[sc->device & 63];
        cudaEvent_t ev = nullptr;
        for (auto& u : B.users) if (u.first == stream) ev = u.second;
        if (!ev) {
            RT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            B.users.push_back({stream, ev});
        }
        RT_CUDA(cudaEventRecord(ev, stream));
        bank_lock.unlock();
    }
    if (!scratch) { sc->pending = true; sc->last_stream = stream; sc->last_mode = mode; sc->last_launches = 1; }
    return RT_OK;
}

int finish_render(rt_scene* sc, rt_stats* st) {
    if (!sc->pending) return fail(RT_ERR_INVALID, "no render in flight on this scene");
    RT_CUDA(cudaEventSynchronize(sc->scr.ev1));
    sc->pending = false;
    if (!st) return RT_OK;
    unsigned long long h[rt::kNumStats];
    RT_CUDA(cudaMemcpy(h, sc->scr.d_stats, sizeof h, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms, sc->scr.ev0, sc->scr.ev1));
    std::memset(st, 0, sizeof *st);
    st->kernel_ms = ms;
    st->samples = h[rt::ST_SAMPLES]; st->casts = h[rt::ST_CASTS];
    st->sphere_tests = (sc->last_mode == RT_SCAN_FILTERED) ? h[rt::ST_CASTS] * (uint64_t)sc->n : 0;
    st->node_tests = rt::kBvhWidth * h[rt::ST_NODE_TESTS];  // child boxes per visited wide node
    st->exact_tests = h[rt::ST_EXACT_TESTS];
    st->black = h[rt::ST_BLACK]; st->early_outs = h[rt::ST_EARLY_OUTS];
    st->primary_hits = h[rt::ST_PRIMARY_HITS]; st->overflows = h[rt::ST_OVERFLOWS];
    st->self_resolved = h[rt::ST_SELF_RESOLVED];
    st->launches = sc->last_launches;
    return RT_OK;
}

int batch_grid(const rt_scene* sc, int nrays) {
    int g = (nrays + rt::kThreads - 1) / rt::kThreads;
    const int cap = sc->sm_count * 4;
    return g < 1 ? 1 : (g > cap ? cap : g);
}

// Builds the scene's arrays from (centres, radii) and copies them into the already allocated device buffers:
// FP32 cull entries, FP64 exact array, 1/r table, and the flattened BVH (rebuilt, or refitted on the kept
// topology).  Shared by rt_upload_scene and rt_update_scene.
int fill_scene(rt_scene* sc, const double* centres_xyz, const double* radii, bool refit) {
    const int n = sc->n;
    // FP32 cull entries {c, -(|c|^2 - r^2 - E_k)}: computed in FP64, constant term rounded DOWN before the negation.
    std::vector<float4> filt((size_t)sc->npad + rt::kScanPad);
    std::vector<double4> exact((size_t)n > 0 ? n : 1);
    std::vector<double> inv_r((size_t)n > 0 ? n : 1);
    sc->cull_ok = true;
    static std::atomic<uint64_t> next_generation{1};
    sc->filt_generation = next_generation++;   // (unique across scenes: a freed scene's address may be reused)
    for (int k = 0; k < sc->npad + rt::kScanPad; ++k) {
        float4 f;
        if (k < n) {
            const double cx = centres_xyz[3 * k], cy = centres_xyz[3 * k + 1], cz = centres_xyz[3 * k + 2], r = radii[k];
            const double a2 = cx * cx + cy * cy + cz * cz, r2 = r * r;
            const double Ek = rt::kCullEps * (rt::kCullKc * a2 + rt::kCullKr * r2);
            f.x = (float)cx; f.y = (float)cy; f.z = (float)cz;
            f.w = -round_down_f32(a2 - r2 - Ek);  // stored negated (cull_D adds it)
            if (!(a2 < rt::kCullMaxMag2) || !(r2 < rt::kCullMaxMag2)) sc->cull_ok = false;  // also catches NaN / inf
            exact[k] = make_double4(cx, cy, cz, r);
            inv_r[k] = 1.0 / r;  // IEEE double division on the host == __ddiv_rn(1.0, r)
        } else {
            f.x = f.y = f.z = 0.f;
            f.w = -INFINITY;  // D = -inf: never passes
        }
        filt[k] = f;
    }
    std::vector<float4> sph((size_t)n > 0 ? n : 1);   // FP32 {centre, |r|}: shell pre-tests (rt_device.cuh: tie_candidate)
    for (int k = 0; k < n; ++k)
        sph[k] = make_float4((float)centres_xyz[3 * k], (float)centres_xyz[3 * k + 1], (float)centres_xyz[3 * k + 2], (float)std::fabs(radii[k]));
    if (cudaMemcpy(sc->d_sph32, sph.data(), sph.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return RT_ERR_CUDA;
    {   // pair-packed copy: spheres 2j, 2j+1 -> {cx,cx',cy,cy'} {cz,cz',w,w'}  (npad + kScanPad is even)
        std::vector<float4> pk(filt.size());
        for (size_t j = 0; j + 1 < filt.size(); j += 2) {
            pk[j] = make_float4(filt[j].x, filt[j + 1].x, filt[j].y, filt[j + 1].y);
            pk[j + 1] = make_float4(filt[j].z, filt[j + 1].z, filt[j].w, filt[j + 1].w);
        }
        if (cudaMemcpy(sc->d_filt_pk, pk.data(), pk.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return RT_ERR_CUDA;
    }
    if (cudaMemcpy(sc->d_filt, filt.data(), filt.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sc->d_exact, exact.data(), exact.size() * sizeof(double4), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(sc->d_inv_r, inv_r.data(), inv_r.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
        return RT_ERR_CUDA;
    // flattened BVH (exact closest-hit semantics, rt_bvh.h); finite scenes only
    if (!sc->cull_ok) {
        cudaFree(sc->d_bvh_nodes); cudaFree(sc->d_bvh_leaf);
        sc->d_bvh_nodes = nullptr; sc->d_bvh_leaf = nullptr; sc->bvh_nodes = 0;
        sc->bvh_host = rt::BvhHost();
        sc->tie = rt::TieGridHost();
        return RT_OK;
    }
    const bool have_tree = sc->d_bvh_nodes && !sc->bvh_host.nodes.empty();
    // tie grid (start-sphere fast path of the BVH mode), rebuilt for the current positions: independent of the tree, so
    // it is built on a thread of its own while this one (and, for large scenes, the builder's threads) does the tree
    std::thread tie_thread;
    int host_threads = 0;   // 0: what the host offers
    if (const char* e = std::getenv("RT_HOST_THREADS")) host_threads = std::atoi(e);   // tuning experiments only (1: sequential build)
    bool tie_inline = !(n >= 4096 && host_threads != 1);
    if (!tie_inline) {
        try { tie_thread = std::thread([&] { rt::build_tie_grid(centres_xyz, radii, n, &sc->tie); }); }
        catch (const std::system_error&) { tie_inline = true; }   // no thread to be had: build it here
    }
    if (tie_inline) rt::build_tie_grid(centres_xyz, radii, n, &sc->tie);
    struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{tie_thread};   // (also on the error returns)
    if (refit && have_tree) {
        rt::refit_bvh(centres_xyz, radii, &sc->bvh_host);
    } else {
        rt::build_bvh(centres_xyz, radii, n, &sc->bvh_host, host_threads);
    }
    const rt::BvhHost& bvh = sc->bvh_host;
    const size_t nb = bvh.nodes4.size() * sizeof(rt::Bvh4Node), lb = (bvh.leaf_idx.size() + 1) * sizeof(int32_t);
    if (!have_tree || (int)bvh.nodes4.size() != sc->bvh_nodes) {  // (a rebuild may change the node count)
        cudaFree(sc->d_bvh_nodes); cudaFree(sc->d_bvh_leaf);
        sc->d_bvh_nodes = nullptr; sc->d_bvh_leaf = nullptr;
        if (cudaMalloc(&sc->d_bvh_nodes, nb) != cudaSuccess || cudaMalloc(&sc->d_bvh_leaf, lb) != cudaSuccess) return RT_ERR_CUDA;
    }
    sc->bvh_nodes = (int)bvh.nodes4.size();
    if (cudaMemcpy(sc->d_bvh_nodes, bvh.nodes4.data(), nb, cudaMemcpyHostToDevice) != cudaSuccess) return RT_ERR_CUDA;
    if (!bvh.leaf_idx.empty() &&
        cudaMemcpy(sc->d_bvh_leaf, bvh.leaf_idx.data(), bvh.leaf_idx.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess)
        return RT_ERR_CUDA;
    if (tie_thread.joinable()) tie_thread.join();
    if (sc->tie.ok) {
        const size_t cb = sc->tie.cells.size() * sizeof(int32_t);
        if (cb > sc->tie_cells_cap) {
            cudaFree(sc->d_tie_cells); sc->d_tie_cells = nullptr; sc->tie_cells_cap = 0;
            if (cudaMalloc(&sc->d_tie_cells, cb) != cudaSuccess) return RT_ERR_CUDA;
            sc->tie_cells_cap = cb;
        }
        if (cudaMemcpy(sc->d_tie_cells, sc->tie.cells.data(), cb, cudaMemcpyHostToDevice) != cudaSuccess) return RT_ERR_CUDA;
    }
    return RT_OK;
}

}  // namespace

extern "C" {

int rt_abi_version(void) { return RT_ABI_VERSION; }
const char* rt_last_error(void) { return g_err.c_str(); }

int rt_params_init(rt_params* p, int32_t width, int32_t height, int32_t spp, int32_t max_depth) {
    if (!p) return fail(RT_ERR_INVALID, "params is NULL");
    std::memset(p, 0, sizeof *p);
    p->width = width; p->height = height; p->spp = spp; p->max_depth = max_depth;
    p->seed = 0; p->tmin = 0.0;                      // programs/main.cc:40
    p->jitter = 1; p->early_out = 1; p->scan_mode = RT_SCAN_AUTO; p->shard_rank = 0; p->shard_count = 1;
    p->custom_shading = 0; p->scatter_mode = RT_SCATTER_HEMISPHERE;   // programs/main.cc:42
    p->albedo = 0.5;                                 // programs/main.cc:43
    p->sky_a[0] = p->sky_a[1] = p->sky_a[2] = 1.0;   // programs/main.cc:48
    p->sky_b[0] = 0.5; p->sky_b[1] = 0.7; p->sky_b[2] = 1.0;
    return RT_OK;
}

int rt_device_info(int32_t device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char* name, int32_t name_cap) {
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_cap > 0) { std::strncpy(name, prop.name, (size_t)name_cap - 1); name[name_cap - 1] = 0; }
    return RT_OK;
}

int rt_upload_scene(const double* centres_xyz, const double* radii, int32_t n, int32_t device, rt_scene** out) {
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n < 0 || (n > 0 && (!centres_xyz || !radii))) return fail(RT_ERR_INVALID, "bad sphere arrays");
    if (n > (1 << 27)) return fail(RT_ERR_UNSUPPORTED, "at most 2^27 spheres");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed (no CUDA device?)");
    rt_scene* sc = new (std::nothrow) rt_scene();
    if (!sc) return fail(RT_ERR_NOMEM, "host allocation failed");
    sc->device = device; sc->n = n;
    sc->npad = (n + rt::kScanStep - 1) / rt::kScanStep * rt::kScanStep;  // the arrays carry kScanPad more never-pass entries
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete sc; return fail(RT_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    sc->sm_count = prop.multiProcessorCount;

    int rc = RT_OK;
    do {
        const size_t nf = (size_t)sc->npad + rt::kScanPad, ne = (size_t)(n > 0 ? n : 1);
        if (cudaMalloc(&sc->d_filt, nf * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&sc->d_filt_pk, nf * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&sc->d_exact, ne * sizeof(double4)) != cudaSuccess ||
            cudaMalloc(&sc->d_inv_r, ne * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&sc->d_sph32, ne * sizeof(float4)) != cudaSuccess ||
            sc->scr.create() != cudaSuccess) { rc = RT_ERR_CUDA; break; }
        rc = fill_scene(sc, centres_xyz, radii, /*refit=*/false);
    } while (0);
    if (rc != RT_OK) {
        const std::string msg = std::string("scene upload: ") + cudaGetErrorString(cudaGetLastError());
        rt_free_scene(sc);
        return fail(rc, msg);
    }
    *out = sc;
    return RT_OK;
}

int rt_update_scene(rt_scene* sc, const double* centres_xyz, const double* radii, int32_t n, int32_t refit) {
    if (!sc || (n > 0 && (!centres_xyz || !radii))) return fail(RT_ERR_INVALID, "NULL argument");
    if (n != sc->n) return fail(RT_ERR_INVALID, "rt_update_scene keeps the sphere count; upload a new scene to change it");
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));  // (rt_render_finish can still read that render's counters)
    RT_CUDA(cudaDeviceSynchronize());   // renders of the old arrays may still be queued on other streams of this device
    const int rc = fill_scene(sc, centres_xyz, radii, refit != 0);
    if (rc != RT_OK) return fail(rc, std::string("scene update: ") + cudaGetErrorString(cudaGetLastError()));
    return RT_OK;
}

void rt_free_scene(rt_scene* sc) {
    if (!sc) return;
    DeviceGuard guard(sc->device);
    if (sc->pending) cudaEventSynchronize(sc->scr.ev1);
    cudaFree(sc->d_filt); cudaFree(sc->d_filt_pk); cudaFree(sc->d_exact); cudaFree(sc->d_inv_r);
    cudaFree(sc->d_bvh_nodes); cudaFree(sc->d_bvh_leaf); cudaFree(sc->d_sph32); cudaFree(sc->d_tie_cells);
    cudaFree(sc->d_frame); cudaFree(sc->d_sum);
    sc->scr.destroy();
    delete sc;
}

int rt_scene_size(const rt_scene* sc) { return sc ? sc->n : RT_ERR_INVALID; }

int rt_get_tile_layout(const rt_params* p, rt_tile_layout* out) {
    int rc = check_params(p);
    if (rc) return rc;
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    tile_layout(p, out);
    return RT_OK;
}

int rt_render_device(const rt_scene* scene, const rt_camera* cam, const rt_params* p, void* d_rgba, void* d_sum,
                     void* stream) {
    if (!scene || !cam || !d_rgba) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));
    if (d_sum && p->shard_count > 1) return fail(RT_ERR_INVALID, "radiance sums are only available for shard_count == 1");
    return launch_render(sc, cam, p, d_rgba, d_sum, (cudaStream_t)stream);
}

int64_t rt_accum_bytes(const rt_params* p) {
    if (check_params(p)) return RT_ERR_INVALID;
    return (int64_t)p->width * p->height * 3 * (int64_t)sizeof(unsigned long long);
}

int rt_render_pass_device(const rt_scene* scene, const rt_camera* cam, const rt_params* p, int32_t sample_begin,
                          void* d_accum, void* d_rgba, void* stream) {
    if (!scene || !cam || !d_accum) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (sample_begin < 0 || (int64_t)sample_begin + p->spp > (1 << 20))
        return fail(RT_ERR_INVALID, "sample_begin + spp must stay within [0, 2^20] (radiance sums are 20.44 fixed point)");
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));
    return launch_render(sc, cam, p, d_rgba, nullptr, (cudaStream_t)stream, sample_begin, d_accum);
}

int rt_render_pass(const rt_scene* scene, const rt_camera* cam, const rt_params* p, int32_t sample_begin,
                   uint64_t* accum, uint8_t* rgba_out, rt_stats* st) {
    if (!scene || !cam || !accum) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->shard_count != 1) return fail(RT_ERR_INVALID, "rt_render_pass renders whole frames; use rt_render_pass_device for shards");
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    const size_t npix = (size_t)p->width * p->height;
    DevBuf<unsigned long long> d_acc;
    RT_CUDA(d_acc.alloc(npix * 3));
    RT_CUDA(cudaMemcpy(d_acc.p, accum, npix * 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    rc = grow(&sc->d_frame, &sc->frame_cap, npix * 4);
    if (rc) return rc;
    rc = rt_render_pass_device(scene, cam, p, sample_begin, d_acc.p, sc->d_frame, nullptr);
    if (rc) return rc;
    rc = finish_render(sc, st);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(accum, d_acc.p, npix * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (rgba_out) RT_CUDA(cudaMemcpy(rgba_out, sc->d_frame, npix * 4, cudaMemcpyDeviceToHost));
    return RT_OK;
}

// ---- device-resident progressive accumulator (no allocation, no host copy per pass)
static int read_stats(const Scratch& X, int n_spheres, int mode, rt_stats* st) {
    unsigned long long h[rt::kNumStats];
    RT_CUDA(cudaMemcpy(h, X.d_stats, sizeof h, cudaMemcpyDeviceToHost));
    float ms = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms, X.ev0, X.ev1));
    std::memset(st, 0, sizeof *st);
    st->kernel_ms = ms;
    st->samples = h[rt::ST_SAMPLES]; st->casts = h[rt::ST_CASTS];
    st->sphere_tests = (mode == RT_SCAN_FILTERED) ? h[rt::ST_CASTS] * (uint64_t)n_spheres : 0;
    st->node_tests = rt::kBvhWidth * h[rt::ST_NODE_TESTS];
    st->exact_tests = h[rt::ST_EXACT_TESTS];
    st->black = h[rt::ST_BLACK]; st->early_outs = h[rt::ST_EARLY_OUTS];
    st->primary_hits = h[rt::ST_PRIMARY_HITS]; st->overflows = h[rt::ST_OVERFLOWS];
    st->self_resolved = h[rt::ST_SELF_RESOLVED];
    st->launches = 1;
    return RT_OK;
}

static int accum_sync(const rt_accum* a) {   // every pass launched so far has finished
    for (int i = 0; i < 2; ++i) if (a->stream[i]) RT_CUDA(cudaStreamSynchronize(a->stream[i]));
    return RT_OK;
}

int rt_accum_create(int32_t width, int32_t height, int32_t device, rt_accum** out) {
    if (!out) return fail(RT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (width < 2 || height < 2 || (int64_t)width * height > (int64_t)1 << 30) return fail(RT_ERR_INVALID, "bad frame size");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed (no CUDA device?)");
    rt_accum* a = new (std::nothrow) rt_accum();
    if (!a) return fail(RT_ERR_NOMEM, "host allocation failed");
    a->device = device; a->W = width; a->H = height;
    const size_t npix = (size_t)width * height;
    bool ok = cudaMalloc(&a->d_sums, npix * 3 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMalloc(&a->d_frame, npix * sizeof(uchar4)) == cudaSuccess &&
              cudaMemset(a->d_sums, 0, npix * 3 * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(a->d_frame, 0, npix * sizeof(uchar4)) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i)
        ok = a->scr[i].create() == cudaSuccess && cudaStreamCreateWithFlags(&a->stream[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        const std::string msg = std::string("rt_accum_create: ") + cudaGetErrorString(cudaGetLastError());
        rt_accum_destroy(a);
        return fail(RT_ERR_CUDA, msg);
    }
    *out = a;
    return RT_OK;
}

void rt_accum_destroy(rt_accum* a) {
    if (!a) return;
    DeviceGuard guard(a->device);
    for (int i = 0; i < 2; ++i) {
        if (a->stream[i]) { cudaStreamSynchronize(a->stream[i]); cudaStreamDestroy(a->stream[i]); }
        a->scr[i].destroy();
    }
    cudaFree(a->d_sums); cudaFree(a->d_frame);
    delete a;
}

int rt_accum_samples(const rt_accum* a) { return a ? a->samples : RT_ERR_INVALID; }

int rt_accum_reset(rt_accum* a) {
    if (!a) return fail(RT_ERR_INVALID, "NULL accumulator");
    DeviceGuard guard(a->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int rc = accum_sync(a);
    if (rc) return rc;
    RT_CUDA(cudaMemset(a->d_sums, 0, (size_t)a->W * a->H * 3 * sizeof(unsigned long long)));
    a->samples = 0;
    return RT_OK;
}

// Consecutive passes alternate between two streams (each with its own scratch), so the ramp-down of pass k -- the last
// work units finishing on a few SMs -- overlaps the start of pass k+1; the integer sums are added atomically and do
// not depend on the order.  (In RT_SCAN_FILTERED mode both streams share the scene's cull array in the constant bank.)
int rt_accum_add(const rt_scene* scene, const rt_camera* cam, const rt_params* p, rt_accum* acc, rt_stats* st) {
    if (!scene || !cam || !acc) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->shard_count != 1) return fail(RT_ERR_INVALID, "rt_accum_add renders whole frames; use rt_render_pass_device for shards");
    if (p->width != acc->W || p->height != acc->H) return fail(RT_ERR_INVALID, "frame size differs from the accumulator's");
    rt_scene* sc = const_cast<rt_scene*>(scene);
    if (sc->device != acc->device) return fail(RT_ERR_INVALID, "scene and accumulator live on different devices");
    if ((int64_t)acc->samples + p->spp > (1 << 20))
        return fail(RT_ERR_INVALID, "at most 2^20 samples per pixel in one accumulator (radiance sums are 20.44 fixed point)");
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));   // (a render through the scene's own entry points)
    const int slot = acc->passes & 1;
    int mode = 0;
    rc = resolve_scan_mode(sc, p->scan_mode, &mode);
    if (rc) return rc;
    rc = launch_render(sc, cam, p, nullptr, nullptr, acc->stream[slot], acc->samples, acc->d_sums, false, &acc->scr[slot]);
    if (rc) return rc;
    acc->passes += 1;
    acc->samples += p->spp;
    acc->last_scene[slot] = sc; acc->last_mode[slot] = mode;
    if (!st) return RT_OK;                       // without stats the pass stays asynchronous
    RT_CUDA(cudaEventSynchronize(acc->scr[slot].ev1));
    return read_stats(acc->scr[slot], sc->n, mode, st);
}

int rt_accum_frame(const rt_accum* a, uint8_t* rgba_out) {
    if (!a || !rgba_out) return fail(RT_ERR_INVALID, "NULL argument");
    DeviceGuard guard(a->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int rc = accum_sync(a);
    if (rc) return rc;
    const size_t npix = (size_t)a->W * a->H;
    if (a->samples > 0) {   // write_color (programs/color.h:16-23) over all samples so far
        int grid = (int)((npix + 255) / 256);
        if (grid > 148 * 16) grid = 148 * 16;
        rt::accum_to_frame_kernel<<<grid, 256, 0, a->stream[0]>>>(a->d_sums, a->d_frame, npix, a->samples);
        RT_CUDA(cudaGetLastError());
        RT_CUDA(cudaStreamSynchronize(a->stream[0]));
    }
    RT_CUDA(cudaMemcpy(rgba_out, a->d_frame, npix * 4, cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_accum_read(const rt_accum* a, uint64_t* sums_out) {
    if (!a || !sums_out) return fail(RT_ERR_INVALID, "NULL argument");
    DeviceGuard guard(a->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int rc = accum_sync(a);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(sums_out, a->d_sums, (size_t)a->W * a->H * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_accum_write(rt_accum* a, const uint64_t* sums, int32_t samples_done) {
    if (!a || !sums || samples_done < 0 || samples_done > (1 << 20)) return fail(RT_ERR_INVALID, "bad argument");
    DeviceGuard guard(a->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int rc = accum_sync(a);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(a->d_sums, sums, (size_t)a->W * a->H * 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    a->samples = samples_done;
    return RT_OK;
}

int rt_accum_to_frame(const rt_params* p, const void* d_accum, int32_t total_samples, void* d_rgba, int32_t device, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (!d_accum || !d_rgba || total_samples < 1) return fail(RT_ERR_INVALID, "bad argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    const size_t npix = (size_t)p->width * p->height;
    int grid = (int)((npix + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    rt::accum_to_frame_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned long long*)d_accum, (uchar4*)d_rgba, npix, total_samples);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

// ---- the pixel loop on several GPUs of this process
int rt_render_multi(rt_scene* const* scenes, int32_t n_scenes, const rt_camera* cam, const rt_params* p, uint8_t* rgba_out,
                    rt_stats* st) {
    if (!scenes || n_scenes < 1 || !cam || !rgba_out) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->shard_count != 1) return fail(RT_ERR_INVALID, "rt_render_multi deals the shards itself: shard_count must be 1");
    for (int i = 0; i < n_scenes; ++i) {
        if (!scenes[i]) return fail(RT_ERR_INVALID, "NULL scene");
        if (scenes[i]->n != scenes[0]->n) return fail(RT_ERR_INVALID, "the scenes of a group hold the same spheres (one copy per device)");
        for (int j = 0; j < i; ++j) if (scenes[j] == scenes[i]) return fail(RT_ERR_INVALID, "a scene handle appears twice");
    }
    rt_scene* sc0 = scenes[0];
    const size_t npix = (size_t)p->width * p->height;
    // peer access from every device to the first one: its frame receives all stores
    bool peer_ok = true;
    for (int i = 1; i < n_scenes && peer_ok; ++i) {
        if (scenes[i]->device == sc0->device) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, scenes[i]->device, sc0->device) != cudaSuccess || !can) { cudaGetLastError(); peer_ok = false; break; }
        DeviceGuard g(scenes[i]->device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(sc0->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) peer_ok = false;
        cudaGetLastError();
    }
    {
        DeviceGuard g0(sc0->device);
        if (!g0.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
        rc = grow(&sc0->d_frame, &sc0->frame_cap, npix * 4);
        if (rc) return rc;
    }
    rt_tile_layout L;
    rt_params q = *p;
    q.shard_count = n_scenes;
    tile_layout(&q, &L);
    // launch every shard (asynchronous, one stream per device), then wait for all of them
    for (int i = 0; i < n_scenes; ++i) {
        rt_scene* sc = scenes[i];
        DeviceGuard g(sc->device);
        if (!g.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
        if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));
        q.shard_rank = i;
        if (peer_ok) {
            rc = launch_render(sc, cam, &q, sc0->d_frame, nullptr, nullptr, 0, nullptr, /*frame_order_out=*/true);
        } else {
            if (i > 0) { rc = grow(&sc->d_frame, &sc->frame_cap, (size_t)L.shard_bytes); if (rc) return rc; }
            rc = launch_render(sc, cam, &q, i == 0 ? (void*)((char*)sc0->d_frame) : sc->d_frame, nullptr, nullptr, 0, nullptr,
                               /*frame_order_out=*/i == 0);
        }
        if (rc) return rc;
    }
    rt_stats total;
    std::memset(&total, 0, sizeof total);
    for (int i = 0; i < n_scenes; ++i) {
        DeviceGuard g(scenes[i]->device);
        rt_stats one;
        rc = finish_render(scenes[i], &one);
        if (rc) return rc;
        total.kernel_ms = one.kernel_ms > total.kernel_ms ? one.kernel_ms : total.kernel_ms;   // the shards run concurrently
        total.samples += one.samples; total.casts += one.casts; total.sphere_tests += one.sphere_tests;
        total.node_tests += one.node_tests; total.exact_tests += one.exact_tests; total.black += one.black;
        total.early_outs += one.early_outs; total.primary_hits += one.primary_hits; total.overflows += one.overflows;
        total.launches += one.launches; total.self_resolved += one.self_resolved;
    }
    {
        DeviceGuard g0(sc0->device);
        RT_CUDA(cudaMemcpy(rgba_out, sc0->d_frame, npix * 4, cudaMemcpyDeviceToHost));
    }
    if (!peer_ok) {
        // no peer access between these devices: the other shards come back as compact tile buffers and are placed on the host
        std::vector<uint8_t> shard((size_t)L.shard_bytes);
        for (int i = 1; i < n_scenes; ++i) {
            DeviceGuard g(scenes[i]->device);
            RT_CUDA(cudaMemcpy(shard.data(), scenes[i]->d_frame, (size_t)L.shard_bytes, cudaMemcpyDeviceToHost));
            for (int l = 0; l < L.tiles_per_shard; ++l) {
                const int t = l * n_scenes + i;
                if (t >= L.tiles_total) break;
                const int ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
                for (int py = 0; py < rt::kTileH && ty * rt::kTileH + py < p->height; ++py)
                    for (int px = 0; px < rt::kTileW && tx * rt::kTileW + px < p->width; ++px)
                        std::memcpy(rgba_out + 4 * ((size_t)(ty * rt::kTileH + py) * p->width + tx * rt::kTileW + px),
                                    shard.data() + 4 * ((size_t)l * rt::kTilePix + py * rt::kTileW + px), 4);
            }
        }
    }
    if (st) *st = total;
    return RT_OK;
}

int rt_render_finish(const rt_scene* scene, rt_stats* st) {
    if (!scene) return fail(RT_ERR_INVALID, "NULL scene");
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    return finish_render(sc, st);
}

int rt_render(const rt_scene* scene, const rt_camera* cam, const rt_params* p, uint8_t* rgba_out, double* sum_out,
              rt_stats* st) {
    if (!scene || !cam || !rgba_out) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    if (p->shard_count != 1) return fail(RT_ERR_INVALID, "rt_render renders whole frames; use rt_render_device for shards");
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    const size_t npix = (size_t)p->width * p->height;
    rc = grow(&sc->d_frame, &sc->frame_cap, npix * 4);
    if (rc) return rc;
    if (sum_out) { rc = grow(&sc->d_sum, &sc->sum_cap, npix * 3 * sizeof(double)); if (rc) return rc; }
    if (sc->pending) RT_CUDA(cudaEventSynchronize(sc->scr.ev1));
    rc = launch_render(sc, cam, p, sc->d_frame, sum_out ? sc->d_sum : nullptr, nullptr);
    if (rc) return rc;
    rc = finish_render(sc, st);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(rgba_out, sc->d_frame, npix * 4, cudaMemcpyDeviceToHost));
    if (sum_out) RT_CUDA(cudaMemcpy(sum_out, sc->d_sum, npix * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_deinterleave(const rt_params* p, const void* d_gathered, void* d_rgba, int32_t device, void* stream) {
    int rc = check_params(p);
    if (rc) return rc;
    if (!d_gathered || !d_rgba) return fail(RT_ERR_INVALID, "NULL buffer");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    rt_tile_layout L;
    tile_layout(p, &L);
    const size_t npix = (size_t)p->width * p->height;
    int grid = (int)((npix + 255) / 256);
    if (grid > 148 * 16) grid = 148 * 16;
    rt::deinterleave_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uchar4*)d_gathered, (uchar4*)d_rgba, p->width,
                                                                    p->height, L.tiles_x, L.tiles_total, p->shard_count,
                                                                    L.tiles_per_shard);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

static int run_hit_kernel(rt_scene* sc, int mode_kernel, rt::RayBatchArgs& a, int nrays) {
    const rt::BatchSmem S = rt::batch_smem(a.sc.npad);
    auto kern = mode_kernel == 0 ? rt::hit_kernel<0> : rt::hit_kernel<1>;
    RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
    kern<<<batch_grid(sc, nrays), rt::kThreads, S.total>>>(a);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaDeviceSynchronize());
    return RT_OK;
}

int rt_primary_hits(const rt_scene* scene, const rt_camera* cam, const rt_params* p, int32_t* idx_out, double* t_out) {
    if (!scene || !cam || !idx_out || !t_out) return fail(RT_ERR_INVALID, "NULL argument");
    int rc = check_params(p);
    if (rc) return rc;
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int mode = 0;
    rc = resolve_scan_mode(sc, p->scan_mode, &mode);
    if (rc) return rc;
    const int npix = p->width * p->height;
    DevBuf<int32_t> d_idx; DevBuf<double> d_t;
    RT_CUDA(d_idx.alloc(npix)); RT_CUDA(d_t.alloc(npix));
    rt::RayBatchArgs a;
    std::memset(&a, 0, sizeof a);
    a.sc = scene_dev(sc, mode); a.nrays = npix; a.tmin = p->tmin; a.tmax = INFINITY; a.scan_mode = mode;
    for (int c = 0; c < 3; ++c) {
        a.cam_org[c] = cam->origin[c]; a.cam_llc[c] = cam->lower_left_corner[c];
        a.cam_hor[c] = cam->horizontal[c]; a.cam_ver[c] = cam->vertical[c];
    }
    a.W = p->width; a.H = p->height; a.idx_out = d_idx.p; a.t_out = d_t.p;
    rc = run_hit_kernel(sc, 1, a, npix);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(idx_out, d_idx.p, (size_t)npix * sizeof(int32_t), cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(t_out, d_t.p, (size_t)npix * sizeof(double), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_hit(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, double tmin, double tmax,
           int32_t scan_mode, int32_t* idx_out, double* rec_out) {
    if (!scene || !org || !dir || !idx_out || !rec_out || nrays < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (nrays == 0) return RT_OK;
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int mode = 0;
    int rc = resolve_scan_mode(sc, scan_mode, &mode);
    if (rc) return rc;
    DevBuf<double> d_org, d_dir, d_rec; DevBuf<int32_t> d_idx;
    RT_CUDA(d_org.alloc(3 * (size_t)nrays)); RT_CUDA(d_dir.alloc(3 * (size_t)nrays));
    RT_CUDA(d_rec.alloc(8 * (size_t)nrays)); RT_CUDA(d_idx.alloc(nrays));
    RT_CUDA(cudaMemcpy(d_org.p, org, 3 * (size_t)nrays * sizeof(double), cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemcpy(d_dir.p, dir, 3 * (size_t)nrays * sizeof(double), cudaMemcpyHostToDevice));
    rt::RayBatchArgs a;
    std::memset(&a, 0, sizeof a);
    a.sc = scene_dev(sc, mode); a.org = d_org.p; a.dir = d_dir.p; a.nrays = nrays; a.tmin = tmin; a.tmax = tmax;
    a.scan_mode = mode; a.idx_out = d_idx.p; a.rec_out = d_rec.p;
    rc = run_hit_kernel(sc, 0, a, nrays);
    if (rc) return rc;
    RT_CUDA(cudaMemcpy(idx_out, d_idx.p, (size_t)nrays * sizeof(int32_t), cudaMemcpyDeviceToHost));
    RT_CUDA(cudaMemcpy(rec_out, d_rec.p, 8 * (size_t)nrays * sizeof(double), cudaMemcpyDeviceToHost));
    return RT_OK;
}

static int ray_color_impl(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, int32_t depth,
                          uint64_t seed, int32_t early_out, int32_t scan_mode, double tmin, const rt::ShadeDev& sh,
                          double* rgb_out, rt_stats* st) {
    if (!scene || !org || !dir || !rgb_out || nrays < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (depth > 1000) return fail(RT_ERR_INVALID, "depth must be <= 1000 (as rt_params.max_depth)");
    if (st) std::memset(st, 0, sizeof *st);
    if (nrays == 0) return RT_OK;
    rt_scene* sc = const_cast<rt_scene*>(scene);
    DeviceGuard guard(sc->device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    int mode = 0;
    int rc = resolve_scan_mode(sc, scan_mode, &mode);
    if (rc) return rc;
    DevBuf<double> d_org, d_dir, d_rgb;
    RT_CUDA(d_org.alloc(3 * (size_t)nrays)); RT_CUDA(d_dir.alloc(3 * (size_t)nrays)); RT_CUDA(d_rgb.alloc(3 * (size_t)nrays));
    RT_CUDA(cudaMemcpy(d_org.p, org, 3 * (size_t)nrays * sizeof(double), cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemcpy(d_dir.p, dir, 3 * (size_t)nrays * sizeof(double), cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemset(sc->scr.d_stats, 0, rt::kNumStats * sizeof(unsigned long long)));
    rt::RayBatchArgs a;
    std::memset(&a, 0, sizeof a);
    a.sc = scene_dev(sc, mode); a.org = d_org.p; a.dir = d_dir.p; a.nrays = nrays; a.scan_mode = mode;
    a.tmin = tmin; a.sh = sh;
    a.depth = depth; a.early_out = (early_out && tmin == 0.0) ? 1 : 0;  // the cut is only exact for tmin == 0
    a.key0 = (uint32_t)seed; a.key1 = (uint32_t)(seed >> 32);
    a.rgb_out = d_rgb.p; a.stats = sc->scr.d_stats;
    const rt::BatchSmem S = rt::batch_smem(a.sc.npad);
    RT_CUDA(cudaFuncSetAttribute(rt::ray_color_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.total));
    rt::ray_color_kernel<<<batch_grid(sc, nrays), rt::kThreads, S.total>>>(a);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaDeviceSynchronize());
    RT_CUDA(cudaMemcpy(rgb_out, d_rgb.p, 3 * (size_t)nrays * sizeof(double), cudaMemcpyDeviceToHost));
    if (st) {
        unsigned long long h[rt::kNumStats];
        RT_CUDA(cudaMemcpy(h, sc->scr.d_stats, sizeof h, cudaMemcpyDeviceToHost));
        st->samples = h[rt::ST_SAMPLES]; st->casts = h[rt::ST_CASTS]; st->exact_tests = h[rt::ST_EXACT_TESTS];
        st->sphere_tests = mode == RT_SCAN_FILTERED ? h[rt::ST_CASTS] * (uint64_t)sc->n : 0;
        st->black = h[rt::ST_BLACK]; st->early_outs = h[rt::ST_EARLY_OUTS]; st->primary_hits = h[rt::ST_PRIMARY_HITS];
        st->overflows = h[rt::ST_OVERFLOWS]; st->self_resolved = h[rt::ST_SELF_RESOLVED]; st->launches = 1;
    }
    return RT_OK;
}

int rt_ray_color(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, int32_t depth, uint64_t seed,
                 int32_t early_out, int32_t scan_mode, double* rgb_out, rt_stats* st) {
    rt_params p;
    rt_params_init(&p, 2, 2, 1, depth);
    return ray_color_impl(scene, org, dir, nrays, depth, seed, early_out, scan_mode, 0.0, shade_dev(&p), rgb_out, st);
}

int rt_ray_color_params(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, const rt_params* p,
                        double* rgb_out, rt_stats* st) {
    if (!p) return fail(RT_ERR_INVALID, "params is NULL");
    rt_params q = *p;
    q.width = q.height = 2; q.spp = 1; q.shard_rank = 0; q.shard_count = 1;   // not used here: keep check_params quiet
    int rc = check_params(&q);
    if (rc) return rc;
    return ray_color_impl(scene, org, dir, nrays, p->max_depth, p->seed, p->early_out, p->scan_mode, p->tmin, shade_dev(p),
                          rgb_out, st);
}

int rt_write_color(const double* rgb_sum, int32_t npix, int32_t spp, int32_t device, int32_t* out) {
    if (!rgb_sum || !out || npix < 0 || spp < 1) return fail(RT_ERR_INVALID, "bad argument");
    if (npix == 0) return RT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    DevBuf<double> d_in; DevBuf<int32_t> d_out;
    RT_CUDA(d_in.alloc(3 * (size_t)npix)); RT_CUDA(d_out.alloc(3 * (size_t)npix));
    RT_CUDA(cudaMemcpy(d_in.p, rgb_sum, 3 * (size_t)npix * sizeof(double), cudaMemcpyHostToDevice));
    rt::write_color_kernel<<<(npix + 255) / 256 > 1024 ? 1024 : (npix + 255) / 256, 256>>>(d_in.p, npix, spp, d_out.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(out, d_out.p, 3 * (size_t)npix * sizeof(int32_t), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_get_ray(const rt_camera* cam, const double* uv, int32_t nq, int32_t device, double* out) {
    if (!cam || !uv || !out || nq < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (nq == 0) return RT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    DevBuf<double> d_uv, d_out;
    RT_CUDA(d_uv.alloc(2 * (size_t)nq)); RT_CUDA(d_out.alloc(6 * (size_t)nq));
    RT_CUDA(cudaMemcpy(d_uv.p, uv, 2 * (size_t)nq * sizeof(double), cudaMemcpyHostToDevice));
    rt::CamArgs c;
    for (int e = 0; e < 3; ++e) {
        c.org[e] = cam->origin[e]; c.llc[e] = cam->lower_left_corner[e]; c.hor[e] = cam->horizontal[e]; c.ver[e] = cam->vertical[e];
    }
    rt::get_ray_kernel<<<(nq + 255) / 256 > 1024 ? 1024 : (nq + 255) / 256, 256>>>(c, d_uv.p, nq, d_out.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(out, d_out.p, 6 * (size_t)nq * sizeof(double), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_philox(const uint32_t* ctr4, const uint32_t* key2, int32_t nblocks, int32_t device, uint32_t* out4) {
    if (!ctr4 || !key2 || !out4 || nblocks < 0) return fail(RT_ERR_INVALID, "bad argument");
    if (nblocks == 0) return RT_OK;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    DevBuf<uint32_t> d_ctr, d_key, d_out;
    RT_CUDA(d_ctr.alloc(4 * (size_t)nblocks)); RT_CUDA(d_key.alloc(2)); RT_CUDA(d_out.alloc(4 * (size_t)nblocks));
    RT_CUDA(cudaMemcpy(d_ctr.p, ctr4, 4 * (size_t)nblocks * sizeof(uint32_t), cudaMemcpyHostToDevice));
    RT_CUDA(cudaMemcpy(d_key.p, key2, 2 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    rt::philox_kernel<<<(nblocks + 255) / 256 > 1024 ? 1024 : (nblocks + 255) / 256, 256>>>(d_ctr.p, d_key.p, nblocks, d_out.p);
    RT_CUDA(cudaGetLastError());
    RT_CUDA(cudaMemcpy(out4, d_out.p, 4 * (size_t)nblocks * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RT_OK;
}

int rt_check_division(int32_t device, uint64_t n, uint64_t seed, uint64_t* mismatches_out) {
    if (!mismatches_out) return fail(RT_ERR_INVALID, "NULL out");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    DevBuf<unsigned long long> d_bad;
    RT_CUDA(d_bad.alloc(1));
    RT_CUDA(cudaMemset(d_bad.p, 0, sizeof(unsigned long long)));
    rt::div_check_kernel<<<1184, 256>>>((unsigned long long)n, (unsigned long long)seed, d_bad.p);
    RT_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    RT_CUDA(cudaMemcpy(&h, d_bad.p, sizeof h, cudaMemcpyDeviceToHost));
    *mismatches_out = h;
    return RT_OK;
}

int rt_measure_fp32_peak(int32_t device, double* fma_per_s_out, double* ms_out) {
    if (!fma_per_s_out) return fail(RT_ERR_INVALID, "NULL out");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RT_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    RT_CUDA(cudaGetDeviceProperties(&prop, device));
    DevBuf<float> d_out;
    RT_CUDA(d_out.alloc(1));
    cudaEvent_t e0, e1;
    RT_CUDA(cudaEventCreate(&e0)); RT_CUDA(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    const double fmas = (double)grid * 256.0 * (double)iters * 16.0 * 8.0;
    double best_ms = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        RT_CUDA(cudaEventRecord(e0));
        rt::ffma_peak_kernel<<<grid, 256>>>(d_out.p, iters, 1.0000001f, 1e-9f);
        RT_CUDA(cudaEventRecord(e1));
        RT_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        RT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *fma_per_s_out = fmas / (best_ms * 1e-3);
    if (ms_out) *ms_out = best_ms;
    return RT_OK;
}

}  // extern "C"
