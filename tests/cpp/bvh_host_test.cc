// Host-only checks of the BVH builder (petershirleyraytracer_b200/csrc/rt_bvh.h): every sphere sits in exactly one
// leaf, every stored box (binary and 4-wide) contains the spheres below it, refit keeps those properties for moved
// spheres, and the depth stays inside the device traversal stack.
#include "rt_bvh.h"

#include <cstdio>
#include <cstring>
#include <random>
#include <set>

#define CHECK(x) do { if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

struct Scene { std::vector<double> c, r; };

static bool box_holds(const float* lo3, const float* hi3, const Scene& s, int k) {
    for (int a = 0; a < 3; ++a) {
        const double rad = std::fabs(s.r[k]);
        if (!((double)lo3[a] <= s.c[3 * k + a] - rad && (double)hi3[a] >= s.c[3 * k + a] + rad)) return false;
    }
    return true;
}

// spheres below a 4-wide child reference; checks containment on the way
static bool walk4(const rt::BvhHost& b, const Scene& s, int32_t ref, std::vector<int>& out, int depth, int& max_depth) {
    if (ref == rt::kBvhEmpty) return true;
    if (ref < 0) {
        const int first = (int)(((uint32_t)ref & 0x7fffffffu) >> 3), count = ref & 7;
        if (count < 1 || count > rt::kBvhLeafMax) return false;
        for (int j = 0; j < count; ++j) out.push_back(b.leaf_idx[(size_t)first + j]);
        return true;
    }
    max_depth = std::max(max_depth, depth + 1);
    const rt::Bvh4Node& n = b.nodes4[(size_t)ref];
    for (int i = 0; i < rt::kBvhWidth; ++i) {
        std::vector<int> below;
        if (!walk4(b, s, n.child[i], below, depth + 1, max_depth)) return false;
        const float lo[3] = {n.lox[i], n.loy[i], n.loz[i]}, hi[3] = {n.hix[i], n.hiy[i], n.hiz[i]};
        for (int k : below) if (!box_holds(lo, hi, s, k)) return false;
        out.insert(out.end(), below.begin(), below.end());
    }
    return true;
}

static int check_tree(const rt::BvhHost& b, const Scene& s, const char* name) {
    const int n = (int)s.r.size();
    std::vector<int> all;
    int max_depth = 0;
    CHECK(walk4(b, s, 0, all, 0, max_depth));
    CHECK((int)all.size() == n);
    std::set<int> uniq(all.begin(), all.end());
    CHECK((int)uniq.size() == n);
    CHECK(max_depth <= 40);   // device stack: 48 entries, up to 3 pushes per level are popped before descending further
    std::printf("%s: n %d, binary nodes %zu, 4-wide nodes %zu, depth %d\n", name, n, b.nodes.size(), b.nodes4.size(), max_depth);
    return 0;
}

int main() {
    std::mt19937_64 g(7);
    std::uniform_real_distribution<double> U(0, 1);
    std::normal_distribution<double> N(0, 1);
    std::vector<std::pair<const char*, Scene>> scenes;
    for (int n : {0, 1, 2, 3, 5, 9}) {
        Scene s;
        for (int k = 0; k < n; ++k) { s.c.insert(s.c.end(), {N(g), N(g), N(g)}); s.r.push_back(0.1 + U(g)); }
        scenes.push_back({"tiny", s});
    }
    { Scene s; s.c.insert(s.c.end(), {0, -1000, 0}); s.r.push_back(1000);
      for (int a = -11; a < 11; ++a) for (int b = -11; b < 11; ++b) { s.c.insert(s.c.end(), {a + 0.9 * U(g), 0.2, b + 0.9 * U(g)}); s.r.push_back(0.2); }
      scenes.push_back({"book", s}); }
    { Scene s; for (int k = 0; k < 300; ++k) { s.c.insert(s.c.end(), {1.0, 2.0, -3.0}); s.r.push_back(0.5 + 0.25 * (k % 3)); } scenes.push_back({"coincident", s}); }
    { Scene s; for (int k = 0; k < 600; ++k) { const double x = 1e-3 * std::pow(1e8, k / 599.0); s.c.insert(s.c.end(), {x, 0, 0}); s.r.push_back(0.3 * x); } scenes.push_back({"line", s}); }
    { Scene s; for (int k = 0; k < 40; ++k) { s.c.insert(s.c.end(), {0, 0, 0}); s.r.push_back(0.01 * std::pow(1e6, k / 39.0)); }
      for (int k = 0; k < 400; ++k) { s.c.insert(s.c.end(), {0.3 * N(g), 0.3 * N(g), 0.3 * N(g)}); s.r.push_back(0.02); } scenes.push_back({"nested", s}); }
    for (auto& [name, s] : scenes) {
        rt::BvhHost b;
        rt::build_bvh(s.c.data(), s.r.data(), (int)s.r.size(), &b);
        if (check_tree(b, s, name)) return 1;
        // move everything, refit on the kept topology: same guarantees for the NEW positions
        Scene m = s;
        for (size_t i = 0; i < m.c.size(); ++i) m.c[i] += 0.7 * N(g);
        for (size_t k = 0; k < m.r.size(); ++k) m.r[k] *= 0.5 + U(g);
        const size_t nodes_before = b.nodes.size();
        rt::refit_bvh(m.c.data(), m.r.data(), &b);
        CHECK(b.nodes.size() == nodes_before);
        if (check_tree(b, m, "  refit")) return 1;
    }
    {   // the threaded build (subtrees on several threads, spliced back in pre-order) is the sequential tree, index for index
        Scene s; s.c.insert(s.c.end(), {0, -1000, 0}); s.r.push_back(1000);
        for (int a = -70; a < 70; ++a) for (int b = -70; b < 70; ++b) { s.c.insert(s.c.end(), {a + 0.9 * U(g), 0.2, b + 0.9 * U(g)}); s.r.push_back(0.2); }
        rt::BvhHost seq, par;
        rt::build_bvh(s.c.data(), s.r.data(), (int)s.r.size(), &seq, 1);
        rt::build_bvh(s.c.data(), s.r.data(), (int)s.r.size(), &par, 8);
        CHECK(seq.nodes.size() == par.nodes.size() && seq.nodes4.size() == par.nodes4.size());
        CHECK(std::memcmp(seq.nodes.data(), par.nodes.data(), seq.nodes.size() * sizeof(rt::BvhNode)) == 0);
        CHECK(std::memcmp(seq.nodes4.data(), par.nodes4.data(), seq.nodes4.size() * sizeof(rt::Bvh4Node)) == 0);
        if (check_tree(par, s, "book 19601, 8 threads")) return 1;
    }
    std::printf("ok\n");
    return 0;
}
