// rt_units.h -- work units of a render launch: how a tile's samples are cut into chunks (host), and how a unit id maps
// back to (tile, chunk, first sample, samples per pixel) (device).  Plain C++ so that tests/cpp can check it on the host.
//
// Work units = (tile, sample chunk).  A launch ends with about one unit's duration of ramp-down, so units should be
// short -- but a warp holds only two units at a time, and a unit whose last paths are still bouncing blocks the
// hand-out of the next one.  Measured on an eighth of the C3 frame (tools/chunk_probe.py, profiles/r2_chunk_probe.txt):
// the linear-scan kernel (64 paths per warp) is best at 4 samples per pixel and unit (8: -0.5 %, 2: -1.8 %), the
// wavefront kernel (96 records per warp) at 16 (8: -1 %, 4: -7 %).  So the chunks are GRADED: most of a tile's samples
// go out in chunks of that efficient length c0, and the launch ends on a level of 4x shorter ones that holds half a
// long unit of work per warp -- what it takes to even out the warps' last long units, which end spread over one
// long unit's duration.  (A second, again 4x shorter level was measured and costs more than it evens out.)
// Levels: lv 0 = the long chunks, lv 1 = one chunk with what does not fill a long one, lv 2 = the short chunks; only the
// very last chunk of a tile may be shorter than its level's length.  Unit ids are level-major (every tile's long chunks
// first), and within a level the tiles go out from the LAST row of the frame up: the ramp-down is set by the longest
// paths of the last units, and this renderer's frames have the sky (one cast per path) at the top and the ground with
// its trapped paths (51 casts) at the bottom.  Any plan and any order give the same frame (integer sums, one Philox
// stream per (pixel, sample)).
#pragma once

#if defined(__CUDACC__)
#define RT_HD __host__ __device__
#else
#define RT_HD
#endif

namespace rt {

struct UnitPlan {
    int lv_n[3];    // chunks per tile and level
    int lv_spp[3];  // samples per pixel of a chunk of that level
};

// The decode functions below are the host's statement of what the kernels do with RenderArgs::lv_n / lv_spp
// (rt_kernels.cuh: unit_of, chunk_first_sample, unit_spp -- kept there verbatim: the register allocation of the scan
// kernel, and with it 2.7 % of its speed, turned out to depend on where these few lines are compiled from,
// profiles/r2_ab_refactor.txt).  tests/cpp/units_test.cc sweeps this copy; the GPU suite checks the kernels' copy by
// rendering graded, ungraded and single-chunk frames that must agree bit for bit.
RT_HD inline int plan_chunks(const UnitPlan& u) { return u.lv_n[0] + u.lv_n[1] + u.lv_n[2]; }

RT_HD inline int chunk_first_sample(const UnitPlan& u, int chunk) {
    const int k1 = chunk - u.lv_n[0], k2 = k1 - u.lv_n[1];
    if (k1 < 0) return chunk * u.lv_spp[0];
    if (k2 < 0) return u.lv_n[0] * u.lv_spp[0] + k1 * u.lv_spp[1];
    return u.lv_n[0] * u.lv_spp[0] + u.lv_n[1] * u.lv_spp[1] + k2 * u.lv_spp[2];
}

RT_HD inline int chunk_spp(const UnitPlan& u, int spp, int chunk) {
    const int k1 = chunk - u.lv_n[0], k2 = k1 - u.lv_n[1];
    const int len = k1 < 0 ? u.lv_spp[0] : (k2 < 0 ? u.lv_spp[1] : u.lv_spp[2]);
    const int left = spp - chunk_first_sample(u, chunk);
    return len < left ? len : left;
}

// unit id -> (tile of this shard, chunk of that tile)
RT_HD inline void unit_of(const UnitPlan& u, int tiles_local, unsigned id, int& tile_l, int& chunk) {
    const unsigned e0 = (unsigned)tiles_local * (unsigned)u.lv_n[0], e1 = e0 + (unsigned)tiles_local * (unsigned)u.lv_n[1];
    unsigned j = id, n = (unsigned)u.lv_n[0];
    int first = 0;
    if (id >= e0) { j = id - e0; n = (unsigned)u.lv_n[1]; first = u.lv_n[0]; }
    if (id >= e1) { j = id - e1; n = (unsigned)u.lv_n[2]; first = u.lv_n[0] + u.lv_n[1]; }
    const unsigned t = j / n;
    tile_l = tiles_local - 1 - (int)t;   // last tile row first
    chunk = first + (int)(j - t * n);
}

// The plan of one launch.  `warps`: resident warps of the grid; `wave`: wavefront kernel (longer efficient chunk);
// `knob` = rt_params.reserved[1]: 0 automatic, n > 0 that many equal chunks per tile, -1 automatic length but ungraded,
// -(10 + F) short level sized at F/4 long units per warp (tests, A/B); `progressive`: a pass of an accumulator -- the
// next pass on the other stream fills its ramp-down, so it is not graded.
inline UnitPlan plan_units(int spp, int tiles_local, int warps, bool wave, int knob, bool progressive) {
    const int min_chunk_spp = wave ? 16 : 4, units_per_warp = wave ? 64 : 256;
    const long long tiles = tiles_local > 0 ? tiles_local : 1;
    long long chunks = ((long long)units_per_warp * warps + tiles - 1) / tiles;
    const long long max_chunks = (spp + min_chunk_spp - 1) / min_chunk_spp;
    if (chunks > max_chunks) chunks = max_chunks;
    if (knob > 0) chunks = knob < spp ? knob : spp;
    if (chunks < 1) chunks = 1;
    const int c0 = (int)((spp + chunks - 1) / chunks);
    UnitPlan u;
    for (int l = 0; l < 3; ++l) { u.lv_n[l] = 0; u.lv_spp[l] = 1; }
    u.lv_spp[0] = c0;
    int quarters = 2;
    if (knob <= -11) { quarters = -knob - 10; if (quarters > 9) quarters = 9; }
    if (knob > 0 || knob == -1 || c0 < 2 || (progressive && knob == 0)) {
        u.lv_n[0] = (spp + c0 - 1) / c0;  // no empty chunk
        return u;
    }
    const int c1 = c0 / 4 > 1 ? c0 / 4 : 1;
    long long t1 = ((long long)quarters * warps * c0 + 4 * tiles - 1) / (4 * tiles);   // samples per pixel of the short level
    t1 = (t1 + c1 - 1) / c1 * c1;
    if (t1 > spp) t1 = spp;
    const int long_total = spp - (int)t1, rem = long_total % c0;
    int short_total = (int)t1;
    u.lv_n[0] = long_total / c0;
    if (rem > c1) { u.lv_n[1] = 1; u.lv_spp[1] = rem; } else short_total += rem;
    u.lv_spp[2] = c1;
    u.lv_n[2] = (short_total + c1 - 1) / c1;   // (the very last chunk may be shorter)
    return u;
}

}  // namespace rt
