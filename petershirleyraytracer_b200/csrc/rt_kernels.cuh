// rt_kernels.cuh -- the sm_100a kernels of the hot path.
//
//   render_kernel<R>   persistent CTAs of 4 warps; each WARP pulls work units (8x8-pixel tile x sample
//                      chunk) from a global counter.  Lanes hold R independent paths each and re-arm a
//                      finished path with the unit's next (pixel, sample) immediately (path regeneration),
//                      so the bimodal path length of the reference (1-4 casts or 51) does not idle lanes; two
//                      units overlap per warp so there is no per-unit tail.  Per cast: FP32 cull scan of
//                      the sphere array (constant bank -> uniform registers, or TMA-staged shared memory),
//                      FP64 exact tests of the survivors, FP64 shading.  Path state lives in shared memory
//                      so the FP64 code exists once.  Radiance is summed in 20.44 fixed point (shared-memory
//                      atomics per warp, global integer atomics across chunks) so the image does not depend
//                      on scheduling or GPU count; write_color's arithmetic (programs/color.h:16-23) runs in
//                      FP64 and stores coalesced uchar4.
//   primary_kernel, hit_kernel, ray_color_kernel   the same device functions behind the per-function
//                      entry points of include/rt.h (one thread per ray, R = 1).
#pragma once

#include "rt_device.cuh"

namespace rt {

// Shared-memory carve-up of the render kernel (R paths per lane).  Everything a path carries between casts
// lives here, so the FP64 code exists once (a rolled loop over the lane's slots) and the registers of the
// scan hold only the cull constants.
struct RenderSmem {
    // filt   : (npad + kScanPad) float4 cull entries        (shared-memory variant only; TMA bulk copy)
    // acc    : kWarps * 2 * kTilePix*3 u64                   per-warp radiance accumulators, two units in flight
    // state  : 6 * R * kThreads double                       ox,oy,oz,dx,dy,dz   [c][r][tid]
    // meta   : R * kThreads uint4                            pixel id, sample, next Philox block, depth | pixel-in-tile<<16
    // cand   : kCandCap * R * kThreads u16                   survivors of the cull scan  [e][r][tid]
    uint32_t filt_bytes, acc_off, state_off, meta_off, cand_off, total;
};
__host__ __device__ inline RenderSmem render_smem(int npad_smem, int R) {
    RenderSmem L;
    L.filt_bytes = npad_smem > 0 ? (uint32_t)(npad_smem + kScanPad) * 16u : 0u;
    L.acc_off = (L.filt_bytes + 127u) & ~127u;
    L.state_off = L.acc_off + kWarps * 2 * kTilePix * 3 * 8;
    L.meta_off = L.state_off + 6 * R * kThreads * 8;
    L.cand_off = L.meta_off + R * kThreads * 16;
    L.total = L.cand_off + kCandCap * R * kThreads * 2;
    return L;
}

// per-function kernels (one ray per thread): cull entries + candidate lists
struct BatchSmem { uint32_t filt_bytes, cand_off, total; };
__host__ __device__ inline BatchSmem batch_smem(int npad) {
    BatchSmem L;
    L.filt_bytes = npad > 0 ? (uint32_t)(npad + kScanPad) * 16u : 0u;
    L.cand_off = (L.filt_bytes + 127u) & ~127u;
    L.total = L.cand_off + kCandCap * kThreads * 2;
    return L;
}

// A work unit = (tile, sample chunk): the 8x8 tile `tile_l` (shard-local index) for the samples of the tile's
// chunk `chunk` (chunk_first_sample / unit_spp) of every pixel.  Warp-uniform.
struct Unit {
    int valid;
    int tile_l, chunk;
    uint32_t total, next;  // (pixel, sample) ids in the unit / already handed to lanes
};

struct TileGeom { int x0, y0, tw, th; };
__device__ __forceinline__ TileGeom tile_geom(const RenderArgs& a, int tile_l) {
    const int t = tile_l * a.shard_count + a.shard_rank;
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    TileGeom g;
    g.x0 = tx * kTileW; g.y0 = ty * kTileH;  // y0 counts rows from the TOP
    g.tw = min(kTileW, a.W - g.x0); g.th = min(kTileH, a.H - g.y0);
    return g;
}
// Graded chunks (RenderArgs::lv_n / lv_spp): unit id -> (tile, chunk of that tile), chunk -> first sample and length.
__device__ __forceinline__ void unit_of(const RenderArgs& a, unsigned id, int& tile_l, int& chunk) {
    const unsigned e0 = (unsigned)a.tiles_local * (unsigned)a.lv_n[0], e1 = e0 + (unsigned)a.tiles_local * (unsigned)a.lv_n[1];
    unsigned j = id, n = (unsigned)a.lv_n[0];
    int first = 0;
    if (id >= e0) { j = id - e0; n = (unsigned)a.lv_n[1]; first = a.lv_n[0]; }
    if (id >= e1) { j = id - e1; n = (unsigned)a.lv_n[2]; first = a.lv_n[0] + a.lv_n[1]; }
    const unsigned t = j / n;
    // Tiles are handed out from the LAST row of the frame up: a launch's ramp-down is set by the longest paths of its
    // last units, and the frames of this renderer have their sky (paths of one cast) at the top and the ground with its
    // trapped paths (51 casts) at the bottom.  (Any order gives the same frame.)
#ifndef RT_TILE_REVERSE
#define RT_TILE_REVERSE 1
#endif
    tile_l = RT_TILE_REVERSE ? a.tiles_local - 1 - (int)t : (int)t;
    chunk = first + (int)(j - t * n);
}
__device__ __forceinline__ int chunk_first_sample(const RenderArgs& a, int chunk) {
    const int k1 = chunk - a.lv_n[0], k2 = k1 - a.lv_n[1];
    if (k1 < 0) return chunk * a.lv_spp[0];
    if (k2 < 0) return a.lv_n[0] * a.lv_spp[0] + k1 * a.lv_spp[1];
    return a.lv_n[0] * a.lv_spp[0] + a.lv_n[1] * a.lv_spp[1] + k2 * a.lv_spp[2];
}
__device__ __forceinline__ int unit_spp(const RenderArgs& a, int chunk) {
    const int k1 = chunk - a.lv_n[0], k2 = k1 - a.lv_n[1];
    const int len = k1 < 0 ? a.lv_spp[0] : (k2 < 0 ? a.lv_spp[1] : a.lv_spp[2]);
    return min(len, a.spp - chunk_first_sample(a, chunk));
}

// write_color's arithmetic (programs/color.h:16-23) for the pixels of a finished tile; coalesced uchar4 rows.
template <bool kFromGlobal>
__device__ __forceinline__ void finalize_tile(const RenderArgs& a, int tile_l, const unsigned long long* src, int lane) {
    const TileGeom g = tile_geom(a, tile_l);
    const double one_over_samples = ddiv(1.0, (double)(a.sample_base + a.spp));  // programs/color.h:16 (all samples so far)
    const double inv_fs = 1.0 / (double)(1ull << kFixShift);
#pragma unroll 1
    for (int pj = 0; pj < kTilePix / 32; ++pj) {  // uniform trip count (lane-strided bounds would mark the warp divergent)
        const int p = pj * 32 + lane;
        const int ly = p / kTileW, lx = p - ly * kTileW;
        if (lx >= g.tw || ly >= g.th) continue;
        unsigned long long v0, v1, v2;
        if (kFromGlobal) { v0 = __ldcg(src + p * 3); v1 = __ldcg(src + p * 3 + 1); v2 = __ldcg(src + p * 3 + 2); }
        else { v0 = src[p * 3]; v1 = src[p * 3 + 1]; v2 = src[p * 3 + 2]; }
        const size_t frame_idx = (size_t)(g.y0 + ly) * a.W + (g.x0 + lx);
        if (a.frame_accum) {  // progressive pass: integer sums continue from the earlier passes (order-independent;
                              // atomic because two passes of one accumulator may be in flight, rt_accum_add)
            unsigned long long* fa = a.frame_accum + frame_idx * 3;
            v0 += atomicAdd(fa + 0, v0); v1 += atomicAdd(fa + 1, v1); v2 += atomicAdd(fa + 2, v2);
        }
        const double sr = __ull2double_rn(v0) * inv_fs, sg = __ull2double_rn(v1) * inv_fs, sb = __ull2double_rn(v2) * inv_fs;
        uchar4 q;
        q.x = (unsigned char)write_color_channel(sr, one_over_samples);
        q.y = (unsigned char)write_color_channel(sg, one_over_samples);
        q.z = (unsigned char)write_color_channel(sb, one_over_samples);
        q.w = 255;
        if (a.out) {
            if (a.compact_out) a.out[(size_t)tile_l * kTilePix + p] = q;
            else a.out[frame_idx] = q;
        }
        if (a.sum_out) {
            a.sum_out[frame_idx * 3 + 0] = sr; a.sum_out[frame_idx * 3 + 1] = sg; a.sum_out[frame_idx * 3 + 2] = sb;
        }
    }
}

// A unit whose samples are all traced: fold its fixed-point sums into the tile.  With one chunk per tile
// the warp finalizes straight from shared memory; otherwise sums go to the global integer accumulator
// (order-independent) and the warp that completes the tile's last chunk finalizes it.
__device__ __forceinline__ void flush_unit(const RenderArgs& a, const Unit& u, unsigned long long* accp, int lane) {
    // No __syncwarp() in here: a warp barrier inside this conditionally executed region makes ptxas treat
    // the main loop as divergent and drop the scan's uniform-datapath loads.  Ordering against the lanes'
    // shared-memory atomics comes from the unconditional __syncwarp()s of the main loop.
    if (a.chunks == 1) {
        finalize_tile<false>(a, u.tile_l, accp, lane);
    } else {
        unsigned long long* g = a.accum + (size_t)u.tile_l * (kTilePix * 3);
#pragma unroll
        for (int j = 0; j < kTilePix * 3 / 32; ++j) {
            const unsigned long long v = accp[j * 32 + lane];
            if (v) atomicAdd(g + j * 32 + lane, v);
        }
        __threadfence();  // this lane's sums are visible before lane 0 (below, after the vote) counts the chunk
        const bool all_here = __all_sync(0xffffffffu, true);
        unsigned int d = 0;
        if (lane == 0 && all_here) d = atomicAdd(a.tile_done + u.tile_l, 1u);
        d = __shfl_sync(0xffffffffu, d, 0);
        if (d == (unsigned)a.chunks - 1u) {
            __threadfence();
            finalize_tile<true>(a, u.tile_l, g, lane);
        }
    }
#pragma unroll
    for (int j = 0; j < kTilePix * 3 / 32; ++j) accp[j * 32 + lane] = 0ull;
}

#ifndef RT_COOP_OVERFLOW
#define RT_COOP_OVERFLOW 1
#endif
// hittable_list::hit for ONE ray by the whole warp: lane i runs the FP64 sphere::hit of spheres i, i + 32, ... on the
// fixed interval [tmin, tmax], then five butterfly steps keep the smallest accepted t and, on equal t, the later list
// index (hittable_list.cc:9-17 in its order-independent form).  Returns the winning list index to every lane, -1 = miss.
// (The candidate lists hold 16-bit indices: the linear scan serves at most kMaxLinearSmem spheres.)
__device__ __noinline__ int coop_list_scan(const SceneDev& sc, int lane, double ox, double oy, double oz, double dx, double dy,
                                           double dz, double A, double tmin, double tmax) {
    Best best;
    best.t = tmax; best.C = 1.0; best.k = -1;
    const RcpA dA = make_rcp(A);
#pragma unroll 1
    for (int k = lane; k < sc.n; k += 32) exact_test_unordered(sc.exact, k, ox, oy, oz, dx, dy, dz, dA, tmin, tmax, best);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double t2 = __shfl_xor_sync(0xffffffffu, best.t, o);
        const int k2 = __shfl_xor_sync(0xffffffffu, best.k, o);
        if (k2 >= 0 && (best.k < 0 || t2 < best.t || (t2 == best.t && k2 > best.k))) { best.t = t2; best.k = k2; }
    }
    return best.k;
}

template <int R, int kSrc>
struct RenderTraits {
#ifndef RT_MINB1
#define RT_MINB1 5
#endif
#ifndef RT_MINB2
#define RT_MINB2 5
#endif
#ifndef RT_MINB4
#define RT_MINB4 3
#endif
    // CTAs of 128 threads per SM (register budget).  The shared-memory variant (kSrc == 0) reads its cull entries with
    // LDS.128 into vector registers: at 96 registers it spilled 40-68 bytes, so it gets 128 (4 CTAs/SM).
    static constexpr int kMinBlocks = R >= 4 ? RT_MINB4 : (kSrc == 0 ? 4 : (R == 2 ? RT_MINB2 : RT_MINB1));
};

// kSrc: where the cast looks for hits: 0 = cull array in TMA-staged shared memory, 1 = cull array in the
// constant bank (default), 2 = flattened BVH (large scenes)
template <int R, int kSrc>
__global__ void __launch_bounds__(kThreads, RenderTraits<R, kSrc>::kMinBlocks) render_kernel(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_mbar;
    constexpr bool kConst = kSrc == 1, kBvh = kSrc == 2;
    const RenderSmem L = render_smem(kSrc == 0 ? a.sc.npad : 0, R);
    const float4* s_filt = reinterpret_cast<const float4*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // two accumulator buffers per warp: the next unit starts while the last paths of the previous one finish
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(smem + L.acc_off) + warp * (2 * kTilePix * 3);
    double* st = reinterpret_cast<double*>(smem + L.state_off) + tid;      // component c of slot r: st[(c*R + r)*kThreads]
    uint4* meta = reinterpret_cast<uint4*>(smem + L.meta_off) + tid;       // slot r: meta[r*kThreads]
    uint16_t* cand = reinterpret_cast<uint16_t*>(smem + L.cand_off) + tid;  // entry e of slot r: cand[(e*R + r)*kThreads]

#pragma unroll
    for (int j = 0; j < 2 * kTilePix * 3 / 32; ++j) acc[j * 32 + lane] = 0ull;
    if (kSrc == 0) stage_bulk(smem, a.sc.filt, L.filt_bytes, &s_mbar);

    const double wm1 = (double)(a.W - 1), hm1 = (double)(a.H - 1);
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);

    uint32_t alive_mask = 0, par_mask = 0;  // per lane, bit r: slot r holds a live path / its unit is buffer 1
    uint32_t cntpack = 0;                   // per slot 8 bits: survivors of the last scan (bit 7: list overflowed)
    uint32_t n_samples = 0, n_casts = 0, n_exact = 0, n_black = 0, n_early = 0, n_primary = 0, n_ovf = 0, n_nodes = 0;
    Best pending[R];  // BVH mode: closest hit found by this round's traversal, consumed by the next advance
#pragma unroll
    for (int r = 0; r < R; ++r) { pending[r].t = 0.0; pending[r].C = 1.0; pending[r].k = -1; }

    Unit u0, u1;
    u0.valid = u1.valid = 0; u0.total = u0.next = u1.total = u1.next = 0; u0.tile_l = u1.tile_l = u0.chunk = u1.chunk = 0;
    int cur = 0;
    bool no_more = false;

    for (;;) {
        __syncwarp();  // orders last round's shared-memory atomics before the sums are read below
        // ---------------- retire units whose samples are all traced
        {
            const bool in0 = (alive_mask & ~par_mask) != 0, in1 = (alive_mask & par_mask) != 0;
            if (__all_sync(0xffffffffu, u0.valid && u0.next >= u0.total && !in0)) { flush_unit(a, u0, acc, lane); u0.valid = 0; }
            if (__all_sync(0xffffffffu, u1.valid && u1.next >= u1.total && !in1)) { flush_unit(a, u1, acc + kTilePix * 3, lane); u1.valid = 0; }
        }

        // ---------------- advance every slot: finish the cast scanned last round, then re-arm dead slots
#pragma unroll 1
        for (int r = 0; r < R; ++r) {
            const uint32_t bit = 1u << r;
#if RT_COOP_OVERFLOW
            if (!kBvh) {
                // A cast whose survivor list overflowed (grazing rays along a sphere row: 0.016 % of the casts, but 1 % of
                // the rounds have one) needs the FP64 test of EVERY sphere.  One lane doing that alone holds the warp for
                // n trips; instead the warp scans the list for that lane's ray, 32 spheres per trip, and reduces to the
                // winner of hittable_list.cc:9-17 (smallest accepted t, later index on ties: order-independent, see
                // exact_test_unordered).  The winner becomes the lane's only candidate and takes the ordinary path below.
                unsigned om = __ballot_sync(0xffffffffu, (alive_mask & bit) && ((cntpack >> (8 * r + 7)) & 1u) && a.scan_mode == 0);
                while (om) {
                    const int L = __ffs((int)om) - 1;
                    om &= om - 1u;
                    const double* sl = st + (L - lane);   // lane L's path state (shared memory: a broadcast read)
                    const double ox = sl[(0 * R + r) * kThreads], oy = sl[(1 * R + r) * kThreads], oz = sl[(2 * R + r) * kThreads];
                    const double dx = sl[(3 * R + r) * kThreads], dy = sl[(4 * R + r) * kThreads], dz = sl[(5 * R + r) * kThreads];
                    const int kw = coop_list_scan(a.sc, lane, ox, oy, oz, dx, dy, dz, ddot(dx, dy, dz, dx, dy, dz), a.tmin, kInf);
                    if (lane == L) {
                        if (kw >= 0) cand[(0 * R + r) * kThreads] = (uint16_t)kw;
                        cntpack = (cntpack & ~(0xffu << (8 * r))) | ((kw >= 0 ? 1u : 0u) << (8 * r));
                        ++n_ovf;
                        n_exact += (uint32_t)a.sc.n;
                    }
                }
            }
#endif
            if (alive_mask & bit) {
                const double ox = st[(0 * R + r) * kThreads], oy = st[(1 * R + r) * kThreads], oz = st[(2 * R + r) * kThreads];
                const double dx = st[(3 * R + r) * kThreads], dy = st[(4 * R + r) * kThreads], dz = st[(5 * R + r) * kThreads];
                uint4 m = meta[r * kThreads];
                const int depth = (int)(m.w & 0xffffu), lp = (int)(m.w >> 16);  // lp: bits 0-5 pixel in tile, bit 6 buffer
                const int bounces = a.max_depth - depth;
                const double A = ddot(dx, dy, dz, dx, dy, dz);  // programs/sphere.cc:9
                const int cnt = (int)((cntpack >> (8 * r)) & 0x7fu);
                const bool ovf = ((cntpack >> (8 * r + 7)) & 1u) != 0;
                ++n_casts;
                if (ovf && a.scan_mode == 0 && (kBvh || !RT_COOP_OVERFLOW)) ++n_ovf;
                Best best = pending[0];
#pragma unroll
                for (int q = 1; q < R; ++q) if (r == q) best = pending[q];
                if (!kBvh || ovf)
                    best = resolve_hits(a.sc, ovf, cnt, cand + r * kThreads, R * kThreads, ox, oy, oz, dx, dy, dz, A, a.tmin,
                                        kInf, n_exact);
                if (best.k < 0) {
                    // miss: sky (programs/main.cc:46-48) * 0.5^bounces -> fixed-point accumulate
                    double cr, cg, cb;
                    sky_color(a.sh, dy, A, bounces, cr, cg, cb);
                    const double fs = (double)(1ull << kFixShift);
                    unsigned long long* ap = acc + ((lp >> 6) * kTilePix + (lp & 63)) * 3;
                    atomicAdd(ap + 0, __double2ull_rz(cr * fs));
                    atomicAdd(ap + 1, __double2ull_rz(cg * fs));
                    atomicAdd(ap + 2, __double2ull_rz(cb * fs));
                    alive_mask &= ~bit;
                } else {
                    if (bounces == 0) ++n_primary;
                    if (a.early_out && best.t == 0.0 && best.C == 0.0) {
                        // origin stays on this sphere with C == 0: every later cast hits at t == 0 -> black
                        ++n_early; ++n_black;
                        alive_mask &= ~bit;
                    } else {
                        const Record rec = make_record(a.sc, best, ox, oy, oz, dx, dy, dz);
                        double rx, ry, rz;
                        random_scatter(m.x, m.y, m.z, a.key0, a.key1, rec.nx, rec.ny, rec.nz, a.sh.lambertian, rx, ry, rz);
                        // programs/main.cc:42-43: target = (p + normal) + rv; next ray = (p, target - p)
                        const double tgx = dadd(dadd(rec.px, rec.nx), rx);
                        const double tgy = dadd(dadd(rec.py, rec.ny), ry);
                        const double tgz = dadd(dadd(rec.pz, rec.nz), rz);
                        st[(0 * R + r) * kThreads] = rec.px; st[(1 * R + r) * kThreads] = rec.py; st[(2 * R + r) * kThreads] = rec.pz;
                        st[(3 * R + r) * kThreads] = dsub(tgx, rec.px);
                        st[(4 * R + r) * kThreads] = dsub(tgy, rec.py);
                        st[(5 * R + r) * kThreads] = dsub(tgz, rec.pz);
                        if (depth == 0) {  // programs/main.cc:36-37: the next ray_color call has depth < 0
                            ++n_black;
                            alive_mask &= ~bit;
                        } else {
                            m.w = (uint32_t)(depth - 1) | ((uint32_t)lp << 16);
                            meta[r * kThreads] = m;
                        }
                    }
                }
            }

            // re-arm: hand the next (pixel, sample) ids to dead slots, pulling units as needed
            unsigned need = __ballot_sync(0xffffffffu, !(alive_mask & bit));
            while (need) {
                Unit c = cur ? u1 : u0;
                // (votes make the warp-uniform exits visible to ptxas, which then keeps the scan's loop
                //  counter and cull entries on the uniform datapath)
                if (__all_sync(0xffffffffu, !c.valid || c.next >= c.total)) {
                    const int np = c.valid ? (cur ^ 1) : cur;
                    const Unit o = np ? u1 : u0;
                    if (__any_sync(0xffffffffu, o.valid || no_more)) break;  // other buffer draining, or frame exhausted
                    unsigned int id = 0;
                    if (lane == 0) id = atomicAdd(a.unit_counter, 1u);
                    id = __shfl_sync(0xffffffffu, id, 0);
                    if (__any_sync(0xffffffffu, id >= (unsigned)a.units_local)) { no_more = true; break; }
                    c.valid = 1;
                    unit_of(a, id, c.tile_l, c.chunk);
                    const TileGeom g = tile_geom(a, c.tile_l);
                    c.total = (uint32_t)(g.tw * g.th) * (uint32_t)unit_spp(a, c.chunk);
                    c.next = 0;
                    cur = np;
                }
                const uint32_t avail = c.total - c.next;
                const uint32_t rank = __popc(need & ((1u << lane) - 1u));
                const bool take = ((need >> lane) & 1u) && rank < avail;
                if (take) {
                    const TileGeom g = tile_geom(a, c.tile_l);
                    const uint32_t id = c.next + rank;
                    const uint32_t ns = (uint32_t)unit_spp(a, c.chunk);
                    const uint32_t p = id / ns;
                    const uint32_t s = (uint32_t)(a.sample_base + chunk_first_sample(a, c.chunk)) + (id - p * ns);
                    const int ly = (int)p / g.tw, lx = (int)p - ly * g.tw;
                    const int i = g.x0 + lx, j = a.H - 1 - (g.y0 + ly);  // j from the bottom (programs/main.cc:72)
                    const uint32_t pixid = (uint32_t)(j * a.W + i);
                    double xu = 0.5, xv = 0.5;
                    if (a.jitter) {
                        const uint4 w = philox4x32_10(pixid, s, 0u, 0u, a.key0, a.key1);
                        xu = u32_unit(w.x); xv = u32_unit(w.y);
                    }
                    const double u = ddiv(dadd((double)i, xu), wm1);  // programs/main.cc:80
                    const double v = ddiv(dadd((double)j, xv), hm1);  // programs/main.cc:81
                    double dx, dy, dz;
                    camera_ray(a.cam_org, a.cam_llc, a.cam_hor, a.cam_ver, u, v, dx, dy, dz);
                    ++n_samples;
                    if (a.max_depth >= 0) {
                        st[(0 * R + r) * kThreads] = a.cam_org[0]; st[(1 * R + r) * kThreads] = a.cam_org[1];
                        st[(2 * R + r) * kThreads] = a.cam_org[2];
                        st[(3 * R + r) * kThreads] = dx; st[(4 * R + r) * kThreads] = dy; st[(5 * R + r) * kThreads] = dz;
                        meta[r * kThreads] = make_uint4(pixid, s, 1u,
                                                        (uint32_t)a.max_depth | ((uint32_t)((ly * kTileW + lx) | (cur << 6)) << 16));
                        alive_mask |= bit;
                        par_mask = (par_mask & ~bit) | ((uint32_t)cur << r);
                    } else {
                        ++n_black;  // ray_color(r, world, depth < 0) is black without a cast (main.cc:36)
                    }
                }
                const unsigned taken = __ballot_sync(0xffffffffu, take);
                c.next += __popc(taken);
                need &= ~taken;
                if (cur) u1 = c; else u0 = c;
                if (a.max_depth < 0) need = __ballot_sync(0xffffffffu, !(alive_mask & bit));  // nothing stays alive: keep draining ids
            }
        }
        if (!__any_sync(0xffffffffu, alive_mask != 0)) {
            if (__all_sync(0xffffffffu, no_more && !u0.valid && !u1.valid)) break;
            continue;  // units fully claimed and nothing in flight: they retire at the top of the loop
        }

        // ---------------- cast: cull constants of every slot, then the FP32 scan (uniform across the warp)
        __threadfence_block();  // accumulator zeroing (retire) is ordered before this round's atomics.  (Deliberately
                                // not a second __syncwarp(): see tests/test_build_sass.py -- ptxas keeps the scan on
                                // the uniform datapath only for this barrier arrangement.)
        CullRay f[R];
        int cnt[R];
        bool ovf[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool live = (alive_mask >> r) & 1u;
            const double ox = st[(0 * R + r) * kThreads], oy = st[(1 * R + r) * kThreads], oz = st[(2 * R + r) * kThreads];
            const double dx = st[(3 * R + r) * kThreads], dy = st[(4 * R + r) * kThreads], dz = st[(5 * R + r) * kThreads];
            const double A = dx * dx + dy * dy + dz * dz;
            // a direction of length 0 / inf / NaN or a far-away origin is left to the sequential FP64 scan
            const bool sane = A > 0.0 && A < kInf && (ox * ox + oy * oy + oz * oz) < kCullMaxMag2;
            cnt[r] = 0;
            if (kBvh) {
                ovf[r] = live && !(sane && a.tmin >= 0.0);  // such rays take the sequential FP64 scan
                if (live && !ovf[r]) {
                    bool deep = false;
                    pending[r].t = kInf; pending[r].C = 1.0; pending[r].k = -1;
                    bvh_cast(a.sc, ox, oy, oz, dx, dy, dz, A, a.tmin, kInf, pending[r], -1, n_exact, n_nodes, deep);
                    ovf[r] = deep;  // traversal stack exhausted (degenerate tree): sequential scan instead
                }
            } else {
                f[r] = make_cull_ray(live && sane && a.scan_mode == 0, ox, oy, oz, dx, dy, dz, A);
                ovf[r] = live && (a.scan_mode != 0 || !sane);
            }
        }
        if (!kBvh && a.scan_mode == 0) {
            if (RT_SCAN_PACKED && kConst) cull_scan_packed<R>(a.sc.npad, f, cand, kThreads, cnt, ovf);
            else cull_scan<R, kConst>(s_filt, a.sc.npad, f, cand, kThreads, cnt, ovf);
        }
        cntpack = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) cntpack |= ((uint32_t)cnt[r] | (ovf[r] ? 0x80u : 0u)) << (8 * r);
    }

    // ---------------- flush counters: warp-shuffle reduce, one atomic per warp and counter
    uint32_t vals[8] = {n_samples, n_casts, n_exact, n_black, n_early, n_primary, n_ovf, n_nodes};
    const int slots[8] = {ST_SAMPLES, ST_CASTS, ST_EXACT_TESTS, ST_BLACK, ST_EARLY_OUTS, ST_PRIMARY_HITS, ST_OVERFLOWS, ST_NODE_TESTS};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        unsigned long long v = vals[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&a.stats[slots[c]], v);
    }
}

// ------------------------------------------------------------------------------------------------
// render_wave_kernel: the BVH mode (RT_SCAN_BVH, the RT_SCAN_AUTO default from 16 spheres up).
//
// The linear-scan kernel above keeps one path per (lane, slot) and runs every lane through the same steps; with a
// tree that costs SIMT efficiency twice over: lanes leave the traversal loop after different numbers of node visits
// (14.5 of 32 threads active per instruction in round 1, profiles/r1p_bvh_kernel_summary.txt), and 86 % of the
// reference's casts do not need a traversal at all (self_cast above).  Here every WARP owns a pool of kPool path
// records in shared memory and three index queues; each round it runs ONE phase on up to 32 records of one queue:
//   S  consume the record's cast result -- sky + accumulate (programs/main.cc:46-48), early-out, or hit_record +
//      scatter (sphere.cc:34-36, main.cc:42-43) -- then start the next cast: FP64 sphere::hit of the sphere the new
//      ray starts on and the tie grid.  Decided casts (86-91 % on the book scene) stay in S, the rest go to T.
//   T  4-wide BVH traversal bounded by the start sphere's hit, then back to S.
//   regeneration: freed records take the warp's next (pixel, sample) ids and enter T as primary rays.
// The queues compact: each phase runs with (nearly) all 32 lanes whatever mix of trapped, escaping and new paths
// the pool holds.  Radiance sums, units, chunk bookkeeping and write_color are the linear-scan kernel's.
#ifndef RT_WAVE_POOL
#define RT_WAVE_POOL 96
#endif
#ifndef RT_WAVE_REGEN
#define RT_WAVE_REGEN 24
#endif
#ifndef RT_WAVE_WARPS
#define RT_WAVE_WARPS 4
#endif
constexpr int kWaveWarps = RT_WAVE_WARPS;             // warps per CTA of the wavefront kernel (they never synchronise)
constexpr int kWaveThreads = 32 * kWaveWarps;
constexpr int kPool = RT_WAVE_POOL;        // path records per warp
constexpr int kRegenMin = RT_WAVE_REGEN;   // free records are refilled in batches of at least this many (a regeneration
                                           // round costs the same for 2 lanes as for 32), so (kPool - kRegenMin) / 2 >= 32
                                           // keeps both phases at full width

// acc += v for a 64-bit fixed-point sum in shared memory, as two native 32-bit atomics with an explicit carry
// (a 64-bit shared-memory atomicAdd compiles to a compare-and-swap loop, ~65 instructions per call site).  All
// additions of a round are complete before anybody reads the sum (the __syncwarp() at the top of the main loop).
__device__ __forceinline__ void fixed_add(unsigned long long* acc, unsigned long long v) {
    unsigned int* w = reinterpret_cast<unsigned int*>(acc);   // little endian: w[0] = low word
    const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
    const unsigned int old = atomicAdd(w, lo);
    const unsigned int carry = (old + lo < old) ? 1u : 0u;
    if (hi | carry) atomicAdd(w + 1, hi + carry);
}

// Cold paths of the wavefront kernel, out of line: its two phases are long straight-line code that every warp streams
// through once per round, so the kernel's hot instruction footprint has to stay inside the 32 KB instruction cache
// (the first version, everything inlined, was 82 KB and spent 3 stall cycles per issued instruction on fetch).
__device__ __noinline__ void flush_unit_cold(const RenderArgs& a, Unit u, unsigned long long* accp, int lane) {
    flush_unit(a, u, accp, lane);
}
__device__ __noinline__ void full_scan_cold(const SceneDev& sc, double ox, double oy, double oz, double dx, double dy,
                                            double dz, double A, double tmin, Best* best, uint32_t* n_exact) {
    *best = resolve_hits(sc, true, 0, nullptr, 0, ox, oy, oz, dx, dy, dz, A, tmin, __longlong_as_double(0x7ff0000000000000ll),
                         *n_exact);
}

struct WaveSmem { uint32_t acc_off, state_off, meta_off, bt_off, bk_off, self_off, q_off, total; };
__host__ __device__ inline WaveSmem wave_smem() {
    WaveSmem L;
    L.acc_off = 0;
    L.state_off = L.acc_off + kWaveWarps * 2 * kTilePix * 3 * 8;
    L.meta_off = L.state_off + kWaveWarps * 6 * kPool * 8;
    L.bt_off = L.meta_off + kWaveWarps * kPool * 16;
    L.bk_off = L.bt_off + kWaveWarps * kPool * 8;
    L.self_off = L.bk_off + kWaveWarps * kPool * 4;
    L.q_off = L.self_off + kWaveWarps * kPool * 4;
    L.total = L.q_off + kWaveWarps * 3 * kPool;
    return L;
}

#ifndef RT_WAVE_MINB
#define RT_WAVE_MINB 4   // 128 registers: at 5 CTAs/SM (96) the two phases spill ~230 bytes and run 12 % slower
#endif
// kTieFlat: the start-sphere step tests all tie-grid candidates in straight-line code (large scenes), or one per trip
template <bool kTieFlat>
__global__ void __launch_bounds__(kWaveThreads, RT_WAVE_MINB) render_wave_kernel(const __grid_constant__ RenderArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const WaveSmem L = wave_smem();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(smem + L.acc_off) + warp * (2 * kTilePix * 3);
    double* st = reinterpret_cast<double*>(smem + L.state_off) + warp * (6 * kPool);   // component c of record i: st[c*kPool + i]
    uint4* meta = reinterpret_cast<uint4*>(smem + L.meta_off) + warp * kPool;          // pixel id, sample, next Philox block, depth | (pixel-in-tile | buffer << 6) << 16
    double* bt = reinterpret_cast<double*>(smem + L.bt_off) + warp * kPool;             // pending cast result: t ...
    int32_t* bk = reinterpret_cast<int32_t*>(smem + L.bk_off) + warp * kPool;           // ... sphere index | (C == 0) << 30, or -1 = miss
    int32_t* selfk = reinterpret_cast<int32_t*>(smem + L.self_off) + warp * kPool;      // sphere the record's ray starts on (-1: camera ray)
    uint8_t* qS = smem + L.q_off + warp * (3 * kPool);
    uint8_t* qT = qS + kPool;
    uint8_t* qF = qT + kPool;

#pragma unroll
    for (int j = 0; j < 2 * kTilePix * 3 / 32; ++j) acc[j * 32 + lane] = 0ull;
    for (int i = lane; i < kPool; i += 32) qF[i] = (uint8_t)i;
    int nS = 0, nT = 0, nF = kPool;   // queue lengths (warp-uniform)
    int infl0 = 0, infl1 = 0;         // records in flight per unit buffer (warp-uniform)

    const RcpA rcp_w = make_rcp((double)(a.W - 1)), rcp_h = make_rcp((double)(a.H - 1));   // divisors of programs/main.cc:80-81
    const double kInf = __longlong_as_double(0x7ff0000000000000ll);
    uint32_t n_samples = 0, n_casts = 0, n_exact = 0, n_black = 0, n_early = 0, n_primary = 0, n_ovf = 0, n_nodes = 0, n_self = 0;

    Unit u0, u1;
    u0.valid = u1.valid = 0; u0.total = u0.next = u1.total = u1.next = 0; u0.tile_l = u1.tile_l = u0.chunk = u1.chunk = 0;
    TileGeom geo;                 // of the unit that is handing out samples (cur)
    geo.x0 = geo.y0 = 0; geo.tw = geo.th = 1;
    uint32_t unit_ns = 1, unit_s0 = 0;   // its samples per pixel and first sample index
    int cur = 0;
    bool no_more = false;

    for (;;) {
        __syncwarp();   // queue / record writes of the last phase are visible to every lane
        // ---------------- retire units whose samples are all traced
        if (u0.valid && u0.next >= u0.total && infl0 == 0) { flush_unit_cold(a, u0, acc, lane); u0.valid = 0; }
        if (u1.valid && u1.next >= u1.total && infl1 == 0) { flush_unit_cold(a, u1, acc + kTilePix * 3, lane); u1.valid = 0; }

        // ---------------- regeneration: free records take the next (pixel, sample) ids
        while (nF >= kRegenMin || (nF > 0 && nS == 0 && nT == 0)) {
            Unit c = cur ? u1 : u0;
            if (!c.valid || c.next >= c.total) {
                const int np = c.valid ? (cur ^ 1) : cur;
                const Unit o = np ? u1 : u0;
                if (o.valid || no_more) break;   // other buffer draining, or frame exhausted
                unsigned int id = 0;
                if (lane == 0) id = atomicAdd(a.unit_counter, 1u);
                id = __shfl_sync(0xffffffffu, id, 0);
                if (id >= (unsigned)a.units_local) { no_more = true; break; }
                c.valid = 1;
                unit_of(a, id, c.tile_l, c.chunk);
                geo = tile_geom(a, c.tile_l);
                unit_ns = (uint32_t)unit_spp(a, c.chunk);
                unit_s0 = (uint32_t)(a.sample_base + chunk_first_sample(a, c.chunk));
                c.total = (uint32_t)(geo.tw * geo.th) * unit_ns;
                c.next = 0;
                cur = np;
            }
            const uint32_t avail = c.total - c.next;
            const int m = (int)min((uint32_t)min(nF, 32), avail);
            if (lane < m) {
                const TileGeom g = geo;
                const uint32_t id = c.next + (uint32_t)lane;
                const uint32_t ns = unit_ns;
                const uint32_t p = id / ns;
                const uint32_t s = unit_s0 + (id - p * ns);
                const int ly = (int)p / g.tw, lx = (int)p - ly * g.tw;
                const int i = g.x0 + lx, j = a.H - 1 - (g.y0 + ly);  // j from the bottom (programs/main.cc:72)
                const uint32_t pixid = (uint32_t)(j * a.W + i);
                double xu = 0.5, xv = 0.5;
                if (a.jitter) {
                    const uint4 w = philox4x32_10(pixid, s, 0u, 0u, a.key0, a.key1);
                    xu = u32_unit(w.x); xv = u32_unit(w.y);
                }
                const double u = ddiv_t(dadd((double)i, xu), rcp_w);  // programs/main.cc:80 (== IEEE quotient, ddiv_t)
                const double v = ddiv_t(dadd((double)j, xv), rcp_h);  // programs/main.cc:81
                double dx, dy, dz;
                camera_ray(a.cam_org, a.cam_llc, a.cam_hor, a.cam_ver, u, v, dx, dy, dz);
                ++n_samples;
                if (a.max_depth >= 0) {
                    const int rec = qF[nF - m + lane];
                    st[0 * kPool + rec] = a.cam_org[0]; st[1 * kPool + rec] = a.cam_org[1]; st[2 * kPool + rec] = a.cam_org[2];
                    st[3 * kPool + rec] = dx; st[4 * kPool + rec] = dy; st[5 * kPool + rec] = dz;
                    meta[rec] = make_uint4(pixid, s, 1u, (uint32_t)a.max_depth | ((uint32_t)((ly * kTileW + lx) | (cur << 6)) << 16));
                    bt[rec] = kInf; bk[rec] = -1; selfk[rec] = -1;
                    qT[nT + lane] = (uint8_t)rec;
                } else {
                    ++n_black;  // ray_color(r, world, depth < 0) is black without a cast (main.cc:36)
                }
            }
            c.next += (uint32_t)m;
            if (cur) u1 = c; else u0 = c;
            if (a.max_depth >= 0) {
                nF -= m; nT += m;
                if (cur) infl1 += m; else infl0 += m;
            }
        }
        if (nS == 0 && nT == 0) {
            if (no_more && !u0.valid && !u1.valid) break;
            continue;   // units fully claimed and nothing in flight: they retire at the top of the loop
        }
        __syncwarp();

        if (nS >= nT) {
            // ================= S: consume a cast result, shade, start the next cast (start sphere + tie grid)
            const int m = min(nS, 32);
            const bool active = lane < m;
            const int rec = active ? (int)qS[nS - m + lane] : 0;
            nS -= m;
            int dest = 0;        // 1: S again (next cast decided here), 2: T, 3: record is free
            int buf = 0;
            if (active) {
                const double ox = st[0 * kPool + rec], oy = st[1 * kPool + rec], oz = st[2 * kPool + rec];
                const double dx = st[3 * kPool + rec], dy = st[4 * kPool + rec], dz = st[5 * kPool + rec];
                uint4 mt = meta[rec];
                const int depth = (int)(mt.w & 0xffffu), lp = (int)(mt.w >> 16);  // lp: bits 0-5 pixel in tile, bit 6 buffer
                buf = lp >> 6;
                const int bounces = a.max_depth - depth;
                const double A = ddot(dx, dy, dz, dx, dy, dz);  // programs/sphere.cc:9
                const int kk = bk[rec];
                ++n_casts;
                if (kk < 0) {
                    // miss: sky (programs/main.cc:46-48) * attenuation -> fixed-point accumulate
                    double cr, cg, cb;
                    sky_color(a.sh, dy, A, bounces, cr, cg, cb);
                    const double fs = (double)(1ull << kFixShift);
                    unsigned long long* ap = acc + (buf * kTilePix + (lp & 63)) * 3;
                    fixed_add(ap + 0, __double2ull_rz(cr * fs));
                    fixed_add(ap + 1, __double2ull_rz(cg * fs));
                    fixed_add(ap + 2, __double2ull_rz(cb * fs));
                    dest = 3;
                } else {
                    Best hitb;
                    hitb.t = bt[rec]; hitb.k = kk & 0x3fffffff; hitb.C = (kk >> 30) & 1 ? 0.0 : 1.0;
                    if (bounces == 0) ++n_primary;
                    if (a.early_out && hitb.t == 0.0 && hitb.C == 0.0) {
                        // origin stays on this sphere with C == 0: every later cast hits at t == 0 -> black
                        ++n_early; ++n_black;
                        dest = 3;
                    } else {
                        const Record rc = make_record(a.sc, hitb, ox, oy, oz, dx, dy, dz);
                        double rx, ry, rz;
                        random_scatter(mt.x, mt.y, mt.z, a.key0, a.key1, rc.nx, rc.ny, rc.nz, a.sh.lambertian, rx, ry, rz);
                        // programs/main.cc:42-43: target = (p + normal) + rv; next ray = (p, target - p)
                        const double ndx = dsub(dadd(dadd(rc.px, rc.nx), rx), rc.px);
                        const double ndy = dsub(dadd(dadd(rc.py, rc.ny), ry), rc.py);
                        const double ndz = dsub(dadd(dadd(rc.pz, rc.nz), rz), rc.pz);
                        if (depth == 0) {  // programs/main.cc:36-37: the next ray_color call has depth < 0
                            ++n_black;
                            dest = 3;
                        } else {
                            st[0 * kPool + rec] = rc.px; st[1 * kPool + rec] = rc.py; st[2 * kPool + rec] = rc.pz;
                            st[3 * kPool + rec] = ndx; st[4 * kPool + rec] = ndy; st[5 * kPool + rec] = ndz;
                            mt.w = (uint32_t)(depth - 1) | ((uint32_t)lp << 16);
                            meta[rec] = mt;
                            // ---- the next cast: start sphere first (programs/sphere.cc:3-32), then the tie grid
                            const double A2 = ddot(ndx, ndy, ndz, ndx, ndy, ndz);
                            const bool sane = A2 > 0.0 && A2 < kInf && (rc.px * rc.px + rc.py * rc.py + rc.pz * rc.pz) < kCullMaxMag2;
                            Best nb;
                            nb.t = kInf; nb.C = 1.0; nb.k = -1;
                            bool decided = false;
                            if (sane && a.tmin >= 0.0)   // (other rays: the T phase runs the sequential FP64 scan)
                                decided = self_cast<kTieFlat>(a.sc, hitb.k, rc.px, rc.py, rc.pz, ndx, ndy, ndz, A2, a.tmin, nb, n_exact);
                            bt[rec] = nb.t;
                            bk[rec] = nb.k < 0 ? -1 : (nb.k | (nb.C == 0.0 ? (1 << 30) : 0));
                            selfk[rec] = hitb.k;
                            if (decided) ++n_self;
                            dest = decided ? 1 : 2;
                        }
                    }
                }
            }
            const unsigned mS = __ballot_sync(0xffffffffu, dest == 1), mT = __ballot_sync(0xffffffffu, dest == 2),
                           mF = __ballot_sync(0xffffffffu, dest == 3), mB = __ballot_sync(0xffffffffu, dest == 3 && buf);
            if (dest == 1) qS[nS + __popc(mS & lt_mask)] = (uint8_t)rec;
            if (dest == 2) qT[nT + __popc(mT & lt_mask)] = (uint8_t)rec;
            if (dest == 3) qF[nF + __popc(mF & lt_mask)] = (uint8_t)rec;
            nS += __popc(mS); nT += __popc(mT); nF += __popc(mF);
            infl1 -= __popc(mB); infl0 -= __popc(mF) - __popc(mB);
        } else {
            // ================= T: BVH traversal bounded by the start sphere's hit, every lane to completion.
            // (Measured and rejected, profiles/r2_ab_fetch.txt / r2_ab_wave.txt: handing finished lanes the next record
            //  inside the traversal loop, one node per iteration, at batch sizes 4..32: 5-25 % slower; shading missed rays
            //  here instead of in S: -1..-9 %; drawing a bounce's random point one bounce ahead: -2..-9 %.)
            const int m = min(nT, 32);
            const bool active = lane < m;
            const int rec = active ? (int)qT[nT - m + lane] : 0;
            nT -= m;
            if (active) {
                const double ox = st[0 * kPool + rec], oy = st[1 * kPool + rec], oz = st[2 * kPool + rec];
                const double dx = st[3 * kPool + rec], dy = st[4 * kPool + rec], dz = st[5 * kPool + rec];
                const double A = ddot(dx, dy, dz, dx, dy, dz);
                const int kk = bk[rec];
                Best best;
                best.t = bt[rec]; best.k = kk < 0 ? -1 : (kk & 0x3fffffff); best.C = (kk >= 0 && ((kk >> 30) & 1)) ? 0.0 : 1.0;
                const bool sane = A > 0.0 && A < kInf && (ox * ox + oy * oy + oz * oz) < kCullMaxMag2;
                bool seq = !(sane && a.tmin >= 0.0);   // such rays take the sequential FP64 scan
                if (!seq) {
                    bool deep = false;
                    bvh_cast(a.sc, ox, oy, oz, dx, dy, dz, A, a.tmin, kInf, best, selfk[rec], n_exact, n_nodes, deep);
                    seq = deep;   // traversal stack exhausted (degenerate tree) / FP32-denormal direction component
                }
                if (seq) {
                    ++n_ovf;
                    full_scan_cold(a.sc, ox, oy, oz, dx, dy, dz, A, a.tmin, &best, &n_exact);
                }
                bt[rec] = best.t;
                bk[rec] = best.k < 0 ? -1 : (best.k | (best.C == 0.0 ? (1 << 30) : 0));
                qS[nS + lane] = (uint8_t)rec;
            }
            nS += m;
        }
    }

    // ---------------- flush counters: warp-shuffle reduce, one atomic per warp and counter
    uint32_t vals[9] = {n_samples, n_casts, n_exact, n_black, n_early, n_primary, n_ovf, n_nodes, n_self};
    const int slots[9] = {ST_SAMPLES, ST_CASTS, ST_EXACT_TESTS, ST_BLACK, ST_EARLY_OUTS, ST_PRIMARY_HITS, ST_OVERFLOWS, ST_NODE_TESTS,
                          ST_SELF_RESOLVED};
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        // (two 16-bit halves: a lane's counter can approach 2^32, the warp sum of a half stays below 2^21)
        const unsigned lo = __reduce_add_sync(0xffffffffu, vals[c] & 0xffffu), hi = __reduce_add_sync(0xffffffffu, vals[c] >> 16);
        const unsigned long long v = (unsigned long long)lo + ((unsigned long long)hi << 16);
        if (lane == 0 && v) atomicAdd(&a.stats[slots[c]], v);
    }
}

// ------------------------------------------------------------------------------------------------
// Per-function entry points (one thread per ray).  They stage the cull array the same way and call the
// same device functions, so the parity tests of rt_hit / rt_primary_hits / rt_ray_color exercise the
// code the render kernel runs.

struct RayBatchArgs {
    SceneDev sc;
    const double* org;   // 3*n (ignored by primary_kernel)
    const double* dir;
    int nrays;
    double tmin, tmax;
    int scan_mode;
    // primary_kernel
    double cam_org[3], cam_llc[3], cam_hor[3], cam_ver[3];
    int W, H;
    // ray_color_kernel
    ShadeDev sh;
    int depth, early_out;
    uint32_t key0, key1;
    // outputs
    int32_t* idx_out;
    double* t_out;
    double* rec_out;  // 8 per ray
    double* rgb_out;  // 3 per ray
    unsigned long long* stats;
};

// `self`: the sphere the ray starts on (the path's previous hit), -1 for primary / explicit rays.  Only the BVH mode
// uses it (start-sphere test + tie grid before the traversal, the same device functions as the render kernel).
__device__ __forceinline__ Best cast_one(const SceneDev& sc, const float4* s_filt, uint16_t* cand, int scan_mode,
                                         bool alive, int self, double ox, double oy, double oz, double dx, double dy,
                                         double dz, double A, double tmin, double tmax, uint32_t& n_exact,
                                         uint32_t& n_ovf, uint32_t& n_self) {
    if (scan_mode == 2) {  // flattened BVH
        Best best;
        best.t = tmax; best.C = 1.0; best.k = -1;
        if (alive) {
            const double kInf = __longlong_as_double(0x7ff0000000000000ll);
            const bool ok = A > 0.0 && A < kInf && tmin >= 0.0 && (ox * ox + oy * oy + oz * oz) < kCullMaxMag2;
            bool deep = false;
            uint32_t n_nodes = 0;
            if (ok) {
                if (tmax == kInf && self_cast<false>(sc, self, ox, oy, oz, dx, dy, dz, A, tmin, best, n_exact)) { ++n_self; return best; }
                bvh_cast(sc, ox, oy, oz, dx, dy, dz, A, tmin, tmax, best, self, n_exact, n_nodes, deep);
            }
            if (!ok || deep) {
                ++n_ovf;
                best = resolve_hits(sc, true, 0, cand, kThreads, ox, oy, oz, dx, dy, dz, A, tmin, tmax, n_exact);
            }
        }
        return best;
    }
    CullRay f[1];
    int cnt[1] = {0};
    const bool sane = A > 0.0 && A < __longlong_as_double(0x7ff0000000000000ll) && (ox * ox + oy * oy + oz * oz) < kCullMaxMag2;
    bool ovf[1] = {alive && (scan_mode != 0 || !sane)};
    f[0] = make_cull_ray(alive && sane && scan_mode == 0, ox, oy, oz, dx, dy, dz, A);
    if (scan_mode == 0) cull_scan<1, false>(s_filt, sc.npad, f, cand, kThreads, cnt, ovf);
    Best best;
    best.t = tmax; best.C = 1.0; best.k = -1;
    if (alive) {
        if (ovf[0] && scan_mode == 0) ++n_ovf;
        best = resolve_hits(sc, ovf[0], cnt[0], cand, kThreads, ox, oy, oz, dx, dy, dz, A, tmin, tmax, n_exact);
    }
    return best;
}

// mode 0: explicit rays -> idx + record (rt_hit); mode 1: pixel-centre camera rays -> idx + t (rt_primary_hits)
template <int MODE>
__global__ void __launch_bounds__(kThreads) hit_kernel(const __grid_constant__ RayBatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_mbar;
    const BatchSmem L = batch_smem(a.sc.npad);
    const float4* s_filt = reinterpret_cast<const float4*>(smem);
    uint16_t* cand = reinterpret_cast<uint16_t*>(smem + L.cand_off) + threadIdx.x;
    stage_bulk(smem, a.sc.filt, L.filt_bytes, &s_mbar);
    uint32_t n_exact = 0, n_ovf = 0;
    const int nrounds = (a.nrays + (int)(gridDim.x * blockDim.x) - 1) / (int)(gridDim.x * blockDim.x);
    for (int round = 0; round < nrounds; ++round) {
        const int q = (round * (int)gridDim.x + (int)blockIdx.x) * (int)blockDim.x + (int)threadIdx.x;
        const bool alive = q < a.nrays;
        double ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1;
        if (alive) {
            if (MODE == 0) {
                ox = a.org[3 * q]; oy = a.org[3 * q + 1]; oz = a.org[3 * q + 2];
                dx = a.dir[3 * q]; dy = a.dir[3 * q + 1]; dz = a.dir[3 * q + 2];
            } else {
                const int row = q / a.W, i = q - row * a.W, j = a.H - 1 - row;
                const double u = ddiv(dadd((double)i, 0.5), (double)(a.W - 1));
                const double v = ddiv(dadd((double)j, 0.5), (double)(a.H - 1));
                ox = a.cam_org[0]; oy = a.cam_org[1]; oz = a.cam_org[2];
                camera_ray(a.cam_org, a.cam_llc, a.cam_hor, a.cam_ver, u, v, dx, dy, dz);
            }
        }
        const double A = ddot(dx, dy, dz, dx, dy, dz);
        uint32_t n_self = 0;
        const Best best = cast_one(a.sc, s_filt, cand, a.scan_mode, alive, -1, ox, oy, oz, dx, dy, dz, A, a.tmin, a.tmax,
                                   n_exact, n_ovf, n_self);
        if (!alive) continue;
        a.idx_out[q] = best.k;
        if (MODE == 1) {
            a.t_out[q] = best.k >= 0 ? best.t : __longlong_as_double(0x7ff0000000000000ll);
        } else {
            double* o = a.rec_out + 8 * (size_t)q;
            if (best.k >= 0) {
                const Record rec = make_record(a.sc, best, ox, oy, oz, dx, dy, dz);
                o[0] = best.t; o[1] = rec.px; o[2] = rec.py; o[3] = rec.pz;
                o[4] = rec.nx; o[5] = rec.ny; o[6] = rec.nz; o[7] = rec.front_face ? 1.0 : 0.0;
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = 0.0;
            }
        }
    }
    if (a.stats) {
        if (n_exact) atomicAdd(&a.stats[ST_EXACT_TESTS], (unsigned long long)n_exact);
        if (n_ovf) atomicAdd(&a.stats[ST_OVERFLOWS], (unsigned long long)n_ovf);
    }
}

// ray_color (programs/main.cc:34-49) on explicit rays, one thread per ray, no regeneration.
__global__ void __launch_bounds__(kThreads) ray_color_kernel(const __grid_constant__ RayBatchArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t s_mbar;
    const BatchSmem L = batch_smem(a.sc.npad);
    const float4* s_filt = reinterpret_cast<const float4*>(smem);
    uint16_t* cand = reinterpret_cast<uint16_t*>(smem + L.cand_off) + threadIdx.x;
    stage_bulk(smem, a.sc.filt, L.filt_bytes, &s_mbar);
    uint32_t n_exact = 0, n_ovf = 0, n_casts = 0, n_black = 0, n_early = 0, n_primary = 0, n_samples = 0, n_self = 0;
    const int nrounds = (a.nrays + (int)(gridDim.x * blockDim.x) - 1) / (int)(gridDim.x * blockDim.x);
    for (int round = 0; round < nrounds; ++round) {
        const int q = (round * (int)gridDim.x + (int)blockIdx.x) * (int)blockDim.x + (int)threadIdx.x;
        bool alive = q < a.nrays && a.depth >= 0;
        double ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 1;
        double cr = 0, cg = 0, cb = 0;
        int depth = a.depth, bounces = 0, self = -1;
        uint32_t blk = 1u;
        if (q < a.nrays) {
            ox = a.org[3 * q]; oy = a.org[3 * q + 1]; oz = a.org[3 * q + 2];
            dx = a.dir[3 * q]; dy = a.dir[3 * q + 1]; dz = a.dir[3 * q + 2];
            ++n_samples;
            if (!alive) ++n_black;
        }
        while (__any_sync(0xffffffffu, alive)) {
            const double A = ddot(dx, dy, dz, dx, dy, dz);
            const Best best = cast_one(a.sc, s_filt, cand, a.scan_mode, alive, self, ox, oy, oz, dx, dy, dz, A, a.tmin,
                                       __longlong_as_double(0x7ff0000000000000ll), n_exact, n_ovf, n_self);
            if (!alive) continue;
            ++n_casts;
            if (best.k < 0) { sky_color(a.sh, dy, A, bounces, cr, cg, cb); alive = false; continue; }
            if (bounces == 0) ++n_primary;
            if (a.early_out && best.t == 0.0 && best.C == 0.0) { ++n_early; ++n_black; alive = false; continue; }
            const Record rec = make_record(a.sc, best, ox, oy, oz, dx, dy, dz);
            double rx, ry, rz;
            random_scatter((uint32_t)q, 0u, blk, a.key0, a.key1, rec.nx, rec.ny, rec.nz, a.sh.lambertian, rx, ry, rz);
            const double tgx = dadd(dadd(rec.px, rec.nx), rx);
            const double tgy = dadd(dadd(rec.py, rec.ny), ry);
            const double tgz = dadd(dadd(rec.pz, rec.nz), rz);
            ox = rec.px; oy = rec.py; oz = rec.pz;
            dx = dsub(tgx, rec.px); dy = dsub(tgy, rec.py); dz = dsub(tgz, rec.pz);
            self = best.k;
            ++bounces;
            if (--depth < 0) { ++n_black; alive = false; }
        }
        if (q < a.nrays) { a.rgb_out[3 * q] = cr; a.rgb_out[3 * q + 1] = cg; a.rgb_out[3 * q + 2] = cb; }
    }
    if (a.stats) {
        const uint32_t vals[8] = {n_samples, n_casts, n_exact, n_black, n_early, n_primary, n_ovf, n_self};
        const int slots[8] = {ST_SAMPLES, ST_CASTS, ST_EXACT_TESTS, ST_BLACK, ST_EARLY_OUTS, ST_PRIMARY_HITS, ST_OVERFLOWS,
                              ST_SELF_RESOLVED};
        for (int c = 0; c < 8; ++c)
            if (vals[c]) atomicAdd(&a.stats[slots[c]], (unsigned long long)vals[c]);
    }
}

// Compact per-shard tile buffers (all-gather output) -> W*H frame.
__global__ void deinterleave_kernel(const uchar4* __restrict__ gathered, uchar4* __restrict__ frame, int W, int H,
                                    int tiles_x, int tiles_total, int shard_count, int tiles_per_shard) {
    const size_t n = (size_t)W * H;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(idx / W), x = (int)(idx - (size_t)y * W);
        const int t = (y / kTileH) * tiles_x + (x / kTileW);
        const int rank = t % shard_count, l = t / shard_count;
        const int p = (y % kTileH) * kTileW + (x % kTileW);
        frame[idx] = gathered[((size_t)rank * tiles_per_shard + l) * kTilePix + p];
    }
    (void)tiles_total;
}

// write_color (programs/color.h:16-23) over a frame-ordered fixed-point accumulator: the last step of a render whose
// SAMPLES were split over ranks (integer all-reduce of the sums first; order-independent, so the frame is the
// single-GPU frame bit for bit).
__global__ void accum_to_frame_kernel(const unsigned long long* __restrict__ accum, uchar4* __restrict__ frame, size_t npix,
                                      int total_samples) {
    const double one_over_samples = ddiv(1.0, (double)total_samples);
    const double inv_fs = 1.0 / (double)(1ull << kFixShift);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        uchar4 q;
        q.x = (unsigned char)write_color_channel(__ull2double_rn(accum[3 * i]) * inv_fs, one_over_samples);
        q.y = (unsigned char)write_color_channel(__ull2double_rn(accum[3 * i + 1]) * inv_fs, one_over_samples);
        q.z = (unsigned char)write_color_channel(__ull2double_rn(accum[3 * i + 2]) * inv_fs, one_over_samples);
        q.w = 255;
        frame[i] = q;
    }
}

// Small device-side mirrors for the per-function parity tests.
__global__ void write_color_kernel(const double* __restrict__ rgb_sum, int npix, int spp, int32_t* __restrict__ out) {
    const double one_over_samples = ddiv(1.0, (double)spp);
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < npix; q += gridDim.x * blockDim.x)
        for (int c = 0; c < 3; ++c) out[3 * q + c] = write_color_channel(rgb_sum[3 * q + c], one_over_samples);
}

struct CamArgs { double org[3], llc[3], hor[3], ver[3]; };
__global__ void get_ray_kernel(const __grid_constant__ CamArgs cam, const double* __restrict__ uv, int nq,
                               double* __restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        double dx, dy, dz;
        camera_ray(cam.org, cam.llc, cam.hor, cam.ver, uv[2 * q], uv[2 * q + 1], dx, dy, dz);
        double* o = out + 6 * (size_t)q;
        o[0] = cam.org[0]; o[1] = cam.org[1]; o[2] = cam.org[2]; o[3] = dx; o[4] = dy; o[5] = dz;
    }
}

__global__ void philox_kernel(const uint32_t* __restrict__ ctr, const uint32_t* __restrict__ key, int n,
                              uint32_t* __restrict__ out) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
        const uint4 w = philox4x32_10(ctr[4 * q], ctr[4 * q + 1], ctr[4 * q + 2], ctr[4 * q + 3], key[0], key[1]);
        out[4 * q] = w.x; out[4 * q + 1] = w.y; out[4 * q + 2] = w.z; out[4 * q + 3] = w.w;
    }
}

// Self-check of ddiv_t() (the short correctly-rounded division of the hit distance) against __ddiv_rn on random
// operand pairs: exponents of A spread over +-260 and of num over +-760 so that both the short form and the
// fallback ranges are exercised, plus zeros, denormals, infinities and all-ones / all-zeros mantissas.
__global__ void div_check_kernel(unsigned long long n, unsigned long long seed, unsigned long long* mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + i * 0x9E3779B97F4A7C15ull, w[3];
        for (int j = 0; j < 3; ++j) {  // splitmix64
            z += 0x9E3779B97F4A7C15ull;
            unsigned long long x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; w[j] = x ^ (x >> 31);
        }
        unsigned long long ma = w[0] & 0xfffffffffffffull, mn = w[1] & 0xfffffffffffffull;
        const unsigned sel = (unsigned)(w[2] & 15u);
        if (sel == 1) ma |= 0xffffffffff000ull;      // mantissa near all ones
        if (sel == 2) ma &= 0x0000000000fffull;      // mantissa near a power of two
        if (sel == 3) mn |= 0xffffffffff000ull;
        if (sel == 4) mn &= 0x0000000000fffull;
        const int ea = (int)((w[2] >> 8) % 521u) - 260, en = (int)((w[2] >> 20) % 1521u) - 760;
        double A = __longlong_as_double((long long)(((unsigned long long)(ea + 1023) << 52) | ma));
        double num = __longlong_as_double((long long)(((unsigned long long)(en + 1023) << 52) | mn | ((w[2] >> 40) << 63)));
        if (sel == 5) num = 0.0;
        if (sel == 6) num = -0.0;
        if (sel == 7) num = __longlong_as_double((long long)(mn >> 3));                    // denormal numerator
        if (sel == 8 && (w[2] >> 41) % 64 == 0) num = __longlong_as_double(0x7ff0000000000000ll);
        if (sel == 9 && (w[2] >> 41) % 64 == 0) A = __longlong_as_double(0x7ff0000000000000ll);
        if (sel == 10 && (w[2] >> 41) % 64 == 0) A = 0.0;
        const RcpA d = make_rcp(A);
        const double q = ddiv_t(num, d), ref = __ddiv_rn(num, A);
        const bool same = __double_as_longlong(q) == __double_as_longlong(ref) || (q != q && ref != ref);
        if (!same) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// FP32 roofline denominator: 8 independent FFMA chains per thread, register operands only.
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456f) out[0] = s;  // keep the chains alive
}

}  // namespace rt
