// Host-only check of the tie grid (petershirleyraytracer_b200/csrc/rt_bvh.h: build_tie_grid): for points on, near and
// far from sphere surfaces, every sphere that passes the device's FP32 shell test (evaluated here with the same float
// arithmetic, at the largest reach the device accepts) must be a giant or be listed in the cell the device would
// look up -- unless that cell is marked overfull or the point lies outside the grid AND the sphere is not near it.
// That containment is what makes "no listed sphere passes" a proof that the start sphere's hit stands.
#include "rt_bvh.h"

#include <cstdio>
#include <random>

#define CHECK(x) do { if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

struct Scene { std::vector<double> c, r; };

static bool candidate(const float* s, float fx, float fy, float fz, float o2, float rho) {   // == rt_device.cuh: tie_candidate
    const float ex = fx - s[0], ey = fy - s[1], ez = fz - s[2];
    const float q = std::fmaf(-s[3], s[3], std::fmaf(ex, ex, std::fmaf(ey, ey, ez * ez)));
    const float w = std::fmaf(s[0], s[0], std::fmaf(s[1], s[1], std::fmaf(s[2], s[2], s[3] * s[3])));
    const float tol = std::fmaf(1.9073486328125e-06f, o2 + w, rho * std::fmaf(2.0002f, s[3], rho));
    return !(std::fabs(q) > tol);
}

static int check_scene(const char* name, const Scene& s, std::mt19937_64& g, int npoints, bool expect_decisive = false) {
    const int n = (int)s.r.size();
    rt::TieGridHost t;
    rt::build_tie_grid(s.c.data(), s.r.data(), n, &t);
    if (!t.ok) { std::printf("%s: n %d, no tie grid (fast path off)\n", name, n); return 0; }
    std::uniform_real_distribution<double> U(0, 1);
    std::normal_distribution<double> N(0, 1);
    long tested = 0, cands = 0, overfull = 0, outside = 0;
    for (int it = 0; it < npoints; ++it) {
        const int k = (int)(U(g) * n) % n;
        double u[3] = {N(g), N(g), N(g)};
        const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        const double rad = std::fabs(s.r[k]);
        const int kind = it % 6;
        const double dist = kind == 0 ? rad : kind == 1 ? rad * (1 + 1e-7) : kind == 2 ? rad + t.rho_max * U(g)
                          : kind == 3 ? rad * (1 - 1e-6) : kind == 4 ? rad * 0.5 : rad * (1 + 3 * U(g));
        double o[3];
        for (int a = 0; a < 3; ++a) o[a] = s.c[3 * k + a] + dist * u[a] / ul;
        const float fx = (float)o[0], fy = (float)o[1], fz = (float)o[2];
        const float o2 = std::fmaf(fx, fx, std::fmaf(fy, fy, fz * fz));
        const bool inside = fx >= t.g0[0] && fx <= t.g1[0] && fy >= t.g0[1] && fy <= t.g1[1] && fz >= t.g0[2] && fz <= t.g1[2];
        const int32_t* cell = nullptr;
        if (inside) {
            const int ix = (int)((fx - t.g0[0]) * t.inv_h), iy = (int)((fy - t.g0[1]) * t.inv_h), iz = (int)((fz - t.g0[2]) * t.inv_h);
            CHECK(ix >= 0 && ix < t.dim[0] && iy >= 0 && iy < t.dim[1] && iz >= 0 && iz < t.dim[2]);
            cell = &t.cells[4 * (((size_t)iz * t.dim[1] + iy) * t.dim[0] + ix)];
            if (cell[0] == rt::kTieOverfull) { ++overfull; continue; }
        } else {
            ++outside;
        }
        ++tested;
        for (int j = 0; j < n; ++j) {
            if (!candidate(&t.sph[4 * (size_t)j], fx, fy, fz, o2, t.rho_max)) continue;
            ++cands;
            bool listed = false;
            for (int gi = 0; gi < t.n_giants; ++gi) listed |= t.giants[gi] == j;
            if (cell) for (int e = 0; e < 4; ++e) listed |= cell[e] == j;
            if (!listed) {
                std::printf("%s: sphere %d passes the shell test at (%g %g %g) [inside %d] but is not listed\n", name, j, o[0], o[1], o[2], (int)inside);
                return 1;
            }
        }
    }
    size_t over_cells = 0;
    for (size_t i = 0; i < t.cells.size(); i += 4) over_cells += t.cells[i] == rt::kTieOverfull;
    std::printf("%s: n %d, grid %dx%dx%d (1/h %g), giants %d, overfull cells %zu; %ld points checked, %ld candidates, %ld in overfull cells, %ld outside\n",
                name, n, t.dim[0], t.dim[1], t.dim[2], (double)t.inv_h, t.n_giants, over_cells, tested, cands, overfull, outside);
    if (expect_decisive) { CHECK(overfull == 0 && cands >= tested / 2); }   // ordinary scenes: every point decidable, on-surface points list their own sphere
    return 0;
}

static Scene book(int G, std::mt19937_64& g) {
    std::uniform_real_distribution<double> U(0, 1);
    Scene s;
    auto add = [&](double x, double y, double z, double r) { s.c.push_back(x); s.c.push_back(y); s.c.push_back(z); s.r.push_back(r); };
    add(0, -1000, 0, 1000);
    for (int a = -G; a < G; ++a)
        for (int b = -G; b < G; ++b) add(a + 0.9 * U(g), 0.2, b + 0.9 * U(g), 0.2);
    add(0, 1, 0, 1); add(-4, 1, 0, 1); add(4, 1, 0, 1);
    return s;
}

int main() {
    std::mt19937_64 g(11);
    std::uniform_real_distribution<double> U(0, 1);
    std::normal_distribution<double> N(0, 1);
    { Scene s = book(11, g); if (check_scene("book11", s, g, 60000, true)) return 1; }
    { Scene s = book(40, g); if (check_scene("book40", s, g, 60000, true)) return 1; }
    {   // clusters with log-uniform radii, duplicates and a giant far off-centre
        Scene s;
        for (int i = 0; i < 3000; ++i) {
            const int cl = i % 12;
            for (int a = 0; a < 3; ++a) s.c.push_back(20.0 * std::sin(cl * (a + 1.7)) + N(g));
            s.r.push_back(std::exp(std::log(1e-3) + U(g) * (std::log(3.0) - std::log(1e-3))));
        }
        for (int i = 100; i < 120; ++i) { for (int a = 0; a < 3; ++a) s.c[3 * i + a] = s.c[300 + a]; s.r[i] = s.r[100]; }
        s.c.push_back(0); s.c.push_back(-5000); s.c.push_back(0); s.r.push_back(4990);
        if (check_scene("clusters", s, g, 60000)) return 1;
    }
    {   // concentric shells (all giants or overfull), far from the origin, negative radius
        Scene s;
        for (int i = 0; i < 40; ++i) { s.c.push_back(500); s.c.push_back(-300); s.c.push_back(100); s.r.push_back((i % 2 ? -1.0 : 1.0) * 0.01 * std::pow(1.4, i)); }
        for (int i = 0; i < 400; ++i) { s.c.push_back(500 + 0.3 * N(g)); s.c.push_back(-300 + 0.3 * N(g)); s.c.push_back(100 + 0.3 * N(g)); s.r.push_back(0.02); }
        if (check_scene("nested", s, g, 40000)) return 1;
    }
    {   // two spheres (the reference's own scene)
        Scene s;
        s.c = {0, 0, -1, 0, -100.5, 0}; s.r = {0.5, 100.0};
        if (check_scene("default", s, g, 20000, true)) return 1;
    }
    {   // one sphere; all coincident
        Scene s; s.c = {1, 2, 3}; s.r = {0.7};
        if (check_scene("single", s, g, 5000)) return 1;
        Scene q;
        for (int i = 0; i < 300; ++i) { q.c.push_back(1); q.c.push_back(2); q.c.push_back(-3); q.r.push_back(0.5 + 0.25 * (i % 3)); }
        if (check_scene("coincident", q, g, 20000)) return 1;
    }
    std::printf("ok\n");
    return 0;
}
