// Host check of rt_units.h: every plan covers each tile's samples exactly once -- the chunks of a tile are contiguous,
// none is empty, only the last may be short, and the unit ids of a launch enumerate every (tile, chunk) exactly once --
// whatever the frame size, spp, grid size, kernel and tuning knob.  (The GPU suite checks the frames; this checks the
// arithmetic where it is cheap to sweep.)
#include "rt_units.h"

#include <cstdio>
#include <vector>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s (line %d): spp %d tiles %d warps %d wave %d knob %d prog %d\n", #c, __LINE__, spp, tiles, warps, (int)wave, knob, (int)prog); return 1; } } while (0)

int main() {
    long plans = 0;
    for (int spp : {1, 2, 3, 4, 5, 7, 8, 16, 17, 31, 48, 64, 100, 256, 500, 1024})
        for (int tiles : {1, 2, 7, 64, 1500, 1875, 15000, 129600})
            for (int warps : {4, 148 * 16, 148 * 20})
                for (bool wave : {false, true})
                    for (int knob : {0, -1, -11, -12, -14, -19, 1, 3, 8, 100000})
                        for (bool prog : {false, true}) {
                            const rt::UnitPlan u = rt::plan_units(spp, tiles, warps, wave, knob, prog);
                            const int chunks = rt::plan_chunks(u);
                            CHECK(chunks >= 1 && chunks <= spp);
                            int next = 0;
                            for (int k = 0; k < chunks; ++k) {
                                CHECK(rt::chunk_first_sample(u, k) == next);
                                const int ns = rt::chunk_spp(u, spp, k);
                                CHECK(ns >= 1);
                                const int lvl = k < u.lv_n[0] ? 0 : (k < u.lv_n[0] + u.lv_n[1] ? 1 : 2);
                                CHECK(ns == u.lv_spp[lvl] || k == chunks - 1);   // only the last chunk may be short
                                next += ns;
                            }
                            CHECK(next == spp);
                            if (knob > 0) CHECK(u.lv_n[1] == 0 && u.lv_n[2] == 0 && chunks <= knob);
                            if (knob == -1 || (prog && knob == 0)) CHECK(u.lv_n[1] == 0 && u.lv_n[2] == 0);
                            if (tiles <= 1875) {   // unit ids <-> (tile, chunk): a bijection, level-major
                                std::vector<char> seen((size_t)tiles * chunks, 0);
                                int last_level = 0;
                                for (unsigned id = 0; id < (unsigned)tiles * (unsigned)chunks; ++id) {
                                    int t = -1, k = -1;
                                    rt::unit_of(u, tiles, id, t, k);
                                    CHECK(t >= 0 && t < tiles && k >= 0 && k < chunks);
                                    CHECK(!seen[(size_t)t * chunks + k]);
                                    seen[(size_t)t * chunks + k] = 1;
                                    const int lvl = k < u.lv_n[0] ? 0 : (k < u.lv_n[0] + u.lv_n[1] ? 1 : 2);
                                    CHECK(lvl >= last_level);
                                    last_level = lvl;
                                }
                            }
                            ++plans;
                        }
    // the shapes the design relies on (DESIGN.md "Graded work units")
    {
        int spp = 500, tiles = 1875, warps = 148 * 20, knob = 0; bool wave = false, prog = false;
        const rt::UnitPlan u = rt::plan_units(spp, tiles, warps, wave, knob, prog);   // 1/8 of C3, scan kernel
        CHECK(u.lv_spp[0] == 4 && u.lv_spp[2] == 1 && u.lv_n[2] >= 1 && u.lv_n[2] <= 8 && u.lv_n[0] >= 120);
        wave = true; warps = 148 * 16;
        const rt::UnitPlan w = rt::plan_units(spp, tiles, warps, wave, knob, prog);   // wavefront kernel
        CHECK(w.lv_spp[0] == 16 && w.lv_spp[2] == 4 && w.lv_n[2] >= 1 && w.lv_n[0] >= 28);
    }
    std::printf("%ld plans\nok\n", plans);
    return 0;
}
