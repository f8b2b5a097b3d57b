// rt_bvh.h -- host-side build of the flattened BVH used for scenes too large for the linear cull scan
// (BASELINE config 4: ~100k spheres).  The reference has no acceleration structure: hittable_list::hit
// (programs/hittable_list.cc:3-20) tests every object.  Its result -- the smallest accepted t, ties going to
// the LATER list index -- does not depend on the order in which objects are visited, so a BVH may replace
// the scan as long as (a) no box that the ray's hit sphere lies in is ever skipped and (b) a subtree is pruned
// only when its entry distance is strictly beyond the current best.  Boxes are therefore stored in FP32
// rounded OUTWARD and padded; the device test adds the per-ray padding (rt_device.cuh: bvh_cast).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <system_error>
#include <thread>
#include <vector>

namespace rt {

// Two children per node, both boxes inline: one 64-byte fetch decides both subtrees.
struct BvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t child0, child1;  // >= 0: node index; < 0: leaf = 0x80000000 | first << 3 | count (count <= 4)
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

#ifndef RT_BVH_LEAF
#define RT_BVH_LEAF 1
#endif
constexpr int kBvhLeafMax = RT_BVH_LEAF;  // spheres per leaf (the leaf reference holds a 3-bit count)

// The device tree: four children per node (the binary tree with every other level collapsed), boxes in
// structure-of-arrays form so that one float4 load brings the same bound of all four children.  128 bytes.
#ifndef RT_BVH_WIDTH
#define RT_BVH_WIDTH 4
#endif
constexpr int kBvhWidth = RT_BVH_WIDTH;  // children per device node (4 or 8)
struct Bvh4Node {
    float lox[kBvhWidth], loy[kBvhWidth], loz[kBvhWidth], hix[kBvhWidth], hiy[kBvhWidth], hiz[kBvhWidth];
    int32_t child[kBvhWidth];  // >= 0: node index; < 0: leaf = 0x80000000 | first << 3 | count; kBvhEmpty: no child
    int32_t pad[kBvhWidth];
};
static_assert(sizeof(Bvh4Node) == 32 * kBvhWidth, "Bvh4Node must be 32 bytes per child");
constexpr int32_t kBvhEmpty = (int32_t)0x80000000u;

struct BvhHost {
    std::vector<BvhNode> nodes;     // binary tree, nodes[0] = root (build / refit work on this one)
    std::vector<int32_t> leaf_idx;  // sphere list indices, leaf by leaf
    std::vector<Bvh4Node> nodes4;   // collapsed 4-wide tree, nodes4[0] = root (what the device traverses)
};

namespace bvh_detail {

struct Box { double lo[3], hi[3]; };

inline float down(double v) { float f = (float)v; if ((double)f > v) f = std::nextafterf(f, -INFINITY); return f; }
inline float up(double v) { float f = (float)v; if ((double)f < v) f = std::nextafterf(f, INFINITY); return f; }

inline void store_box(const Box& b, float* lo, float* hi) {
    for (int a = 0; a < 3; ++a) {
        // pad: 2^-20 of the coordinate magnitude covers FP32 rounding of the box, of the ray origin and of
        // the subtraction in the slab test; the reference's own FP64 slop is 2^-32 of that
        const double mag = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
        const double pad = mag * 9.5367431640625e-07 + 1e-30;
        lo[a] = down(b.lo[a] - pad);
        hi[a] = up(b.hi[a] + pad);
    }
}

inline void empty_box(float* lo, float* hi) {
    for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
}

struct Builder {
    const double* c; const double* r;
    std::vector<int32_t>& order;   // sphere indices; a (sub)tree permutes only its own range, so subtrees build in parallel
    BvhHost* out;

    Box bounds(int first, int count) const {
        Box b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; }
        for (int i = first; i < first + count; ++i) {
            const int k = order[i];
            const double rad = std::fabs(r[k]);
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = std::min(b.lo[a], c[3 * k + a] - rad);
                b.hi[a] = std::max(b.hi[a], c[3 * k + a] + rad);
            }
        }
        return b;
    }

    static double half_area(const Box& b) {
        const double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
        return dx * dy + dy * dz + dz * dx;
    }
    static void grow(Box& b, const Box& o) {
        for (int a = 0; a < 3; ++a) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); }
    }
    Box sphere_box(int k) const {
        Box b;
        const double rad = std::fabs(r[k]);
        for (int a = 0; a < 3; ++a) { b.lo[a] = c[3 * k + a] - rad; b.hi[a] = c[3 * k + a] + rad; }
        return b;
    }

    // Binned surface-area heuristic (16 bins per axis over the centroid bounds): returns the number of spheres
    // that go left after partitioning order[first, first+count), or 0 if no split beats an object-median split
    // (all centroids in one bin).  Besides the usual quality gain this isolates outsized spheres near the root:
    // the r = 1000 ground sphere of the book scenes would otherwise inflate every box on the way down to its
    // leaf, and every ray would walk that spine.
    int sah_split(int first, int count) {
        constexpr int kBins = 16;
        double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = first; i < first + count; ++i)
            for (int a = 0; a < 3; ++a) {
                clo[a] = std::min(clo[a], c[3 * order[i] + a]);
                chi[a] = std::max(chi[a], c[3 * order[i] + a]);
            }
        double best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        for (int axis = 0; axis < 3; ++axis) {
            const double ext = chi[axis] - clo[axis];
            if (!(ext > 0.0) || !std::isfinite(ext)) continue;
            const double scale = kBins / ext;
            Box bb[kBins];
            int bn[kBins];
            for (int b = 0; b < kBins; ++b) {
                bn[b] = 0;
                for (int a = 0; a < 3; ++a) { bb[b].lo[a] = INFINITY; bb[b].hi[a] = -INFINITY; }
            }
            for (int i = first; i < first + count; ++i) {
                const int k = order[i];
                int b = (int)((c[3 * k + axis] - clo[axis]) * scale);
                b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
                ++bn[b];
                grow(bb[b], sphere_box(k));
            }
            double right_area[kBins];
            int right_n[kBins];
            Box acc;
            for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
            int n_acc = 0;
            for (int b = kBins - 1; b > 0; --b) {
                if (bn[b]) grow(acc, bb[b]);
                n_acc += bn[b];
                right_area[b] = n_acc ? half_area(acc) : 0.0;
                right_n[b] = n_acc;
            }
            for (int a = 0; a < 3; ++a) { acc.lo[a] = INFINITY; acc.hi[a] = -INFINITY; }
            n_acc = 0;
            for (int b = 0; b < kBins - 1; ++b) {  // split between bin b and b+1
                if (bn[b]) grow(acc, bb[b]);
                n_acc += bn[b];
                if (n_acc == 0 || right_n[b + 1] == 0) continue;
                const double cost = half_area(acc) * n_acc + right_area[b + 1] * right_n[b + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        if (best_axis < 0) return 0;
        const double ext = chi[best_axis] - clo[best_axis], scale = kBins / ext;
        const double lo = clo[best_axis];
        const int axis = best_axis, bin = best_bin;
        auto mid = std::partition(order.begin() + first, order.begin() + first + count, [&](int32_t k) {
            int b = (int)((c[3 * k + axis] - lo) * scale);
            b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
            return b <= bin;
        });
        const int left = (int)(mid - (order.begin() + first));
        return (left > 0 && left < count) ? left : 0;
    }

    // returns the child reference for spheres order[first, first+count); nodes are appended to `nodes` in pre-order
    // (parent, left subtree, right subtree).  While fork > 0 the left subtree is built by another thread into a vector
    // of its own and spliced in behind the parent -- the same array, index for index, as the sequential build.
    int32_t build(std::vector<BvhNode>& nodes, int first, int count, int depth = 0, int fork = 0) {
        if (kBvhLeafMax == 1 && count == 1) {
            // single-sphere leaves name their sphere directly (leaf_idx is the identity): the device skips the
            // leaf-list load, a dependent fetch in front of every FP64 sphere test
            return (int32_t)(0x80000000u | ((uint32_t)order[first] << 3) | 1u);
        }
        if (count <= kBvhLeafMax) {
            const int32_t at = (int32_t)out->leaf_idx.size();
            // inside a leaf keep list order (not required for correctness; keeps tests readable)
            std::sort(order.begin() + first, order.begin() + first + count);
            for (int i = 0; i < count; ++i) out->leaf_idx.push_back(order[first + i]);
            return (int32_t)(0x80000000u | ((uint32_t)at << 3) | (uint32_t)count);
        }
        // below depth 24 only median splits: bounds the tree depth (device traversal stack: 48 entries)
        int mid = depth < 24 ? sah_split(first, count) : 0;
        if (mid == 0) {
            // degenerate centroids (coincident spheres): object median of the axis with the largest extent
            double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int i = first; i < first + count; ++i)
                for (int a = 0; a < 3; ++a) {
                    lo[a] = std::min(lo[a], c[3 * order[i] + a]);
                    hi[a] = std::max(hi[a], c[3 * order[i] + a]);
                }
            int axis = 0;
            for (int a = 1; a < 3; ++a) if (hi[a] - lo[a] > hi[axis] - lo[axis]) axis = a;
            mid = count / 2;
            std::nth_element(order.begin() + first, order.begin() + first + mid, order.begin() + first + count,
                             [&](int32_t x, int32_t y) { return c[3 * x + axis] < c[3 * y + axis]; });
        }
        const int32_t me = (int32_t)nodes.size();
        nodes.push_back(BvhNode{});
        int32_t c0, c1;
        Box b0, b1;
        if (fork > 0 && kBvhLeafMax == 1 && count >= 4096) {
            std::vector<BvhNode> left;
            int32_t c0_local = 0;
            auto build_left = [&] { b0 = bounds(first, mid); c0_local = build(left, first, mid, depth + 1, fork - 1); };
            std::thread th;
            try { th = std::thread(build_left); } catch (const std::system_error&) {}   // no thread to be had: build it here
            b1 = bounds(first + mid, count - mid);
            std::vector<BvhNode> right;
            const int32_t c1_local = build(right, first + mid, count - mid, depth + 1, fork - 1);
            if (th.joinable()) th.join(); else build_left();
            auto splice = [&](std::vector<BvhNode>& sub, int32_t ref) {
                const int32_t off = (int32_t)nodes.size();
                for (BvhNode& n : sub) {
                    if (n.child0 >= 0) n.child0 += off;
                    if (n.child1 >= 0) n.child1 += off;
                }
                nodes.insert(nodes.end(), sub.begin(), sub.end());
                return ref >= 0 ? ref + off : ref;
            };
            c0 = splice(left, c0_local);
            c1 = splice(right, c1_local);
        } else {
            b0 = bounds(first, mid); b1 = bounds(first + mid, count - mid);
            c0 = build(nodes, first, mid, depth + 1, 0);
            c1 = build(nodes, first + mid, count - mid, depth + 1, 0);
        }
        BvhNode& n = nodes[me];
        store_box(b0, n.lo0, n.hi0);
        store_box(b1, n.lo1, n.hi1);
        n.child0 = c0; n.child1 = c1;
        return me;
    }
};

}  // namespace bvh_detail

inline void collapse_bvh4(BvhHost* bvh);

// Refit: same topology (nodes, leaf lists), boxes recomputed for moved / resized spheres.  Children are
// stored after their parent, so one pass from the last node to the root sees every child box before it is needed.
// Exactness never depends on the tree's quality, only the traversal cost does; rebuild after large motion.
inline void refit_bvh(const double* centres, const double* radii, BvhHost* bvh) {
    using namespace bvh_detail;
    const int nn = (int)bvh->nodes.size();
    std::vector<Box> whole((size_t)nn);  // union of both child boxes of node i
    auto child_box = [&](int32_t ref) {
        Box b;
        for (int a = 0; a < 3; ++a) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; }
        if (ref >= 0) return whole[(size_t)ref];
        const int first = (int)(((uint32_t)ref & 0x7fffffffu) >> 3), count = ref & 7;
        for (int i = 0; i < count; ++i) {
            const int k = bvh->leaf_idx[(size_t)first + i];
            const double rad = std::fabs(radii[k]);
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = std::min(b.lo[a], centres[3 * k + a] - rad);
                b.hi[a] = std::max(b.hi[a], centres[3 * k + a] + rad);
            }
        }
        return b;
    };
    for (int i = nn - 1; i >= 0; --i) {
        BvhNode& n = bvh->nodes[(size_t)i];
        const Box b0 = child_box(n.child0), b1 = child_box(n.child1);
        const bool e0 = n.child0 < 0 && (n.child0 & 7) == 0, e1 = n.child1 < 0 && (n.child1 & 7) == 0;
        if (e0) empty_box(n.lo0, n.hi0); else store_box(b0, n.lo0, n.hi0);
        if (e1) empty_box(n.lo1, n.hi1); else store_box(b1, n.lo1, n.hi1);
        Box w = b0;
        for (int a = 0; a < 3; ++a) { w.lo[a] = std::min(w.lo[a], b1.lo[a]); w.hi[a] = std::max(w.hi[a], b1.hi[a]); }
        whole[(size_t)i] = w;
    }
    collapse_bvh4(bvh);  // (derived from the binary tree: same topology unless box areas changed the collapse order)
}

// Binary -> 4-wide: a node's children are replaced by their own children, largest box first, until four slots are
// filled or only leaves remain.  Works on the stored (already padded, FP32) child boxes, so the 4-wide boxes are
// exactly the binary tree's boxes and the conservativeness argument carries over unchanged.
inline void collapse_bvh4(BvhHost* bvh) {
    struct Slot { float lo[3], hi[3]; int32_t ref; };
    auto area = [](const Slot& s) {
        const double dx = (double)s.hi[0] - s.lo[0], dy = (double)s.hi[1] - s.lo[1], dz = (double)s.hi[2] - s.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    auto slots_of = [&](int32_t node, Slot* out) {
        const BvhNode& b = bvh->nodes[(size_t)node];
        for (int a = 0; a < 3; ++a) { out[0].lo[a] = b.lo0[a]; out[0].hi[a] = b.hi0[a]; out[1].lo[a] = b.lo1[a]; out[1].hi[a] = b.hi1[a]; }
        out[0].ref = b.child0; out[1].ref = b.child1;
    };
    bvh->nodes4.clear();
    if (bvh->nodes.empty()) return;
    std::vector<std::pair<int32_t, int32_t>> todo;  // (binary node, 4-wide node to fill)
    bvh->nodes4.push_back(Bvh4Node{});
    todo.push_back({0, 0});
    while (!todo.empty()) {
        const auto [bin, me] = todo.back();
        todo.pop_back();
        Slot s[kBvhWidth];
        int ns = 2;
        slots_of(bin, s);
        while (ns < kBvhWidth) {
            int pick = -1;
            double best = -1.0;
            for (int i = 0; i < ns; ++i)
                if (s[i].ref >= 0) {
                    const double a = area(s[i]);
                    if (!(a <= best)) { best = a; pick = i; }  // (NaN-safe: inf - inf areas still get picked)
                }
            if (pick < 0) break;
            Slot two[2];
            slots_of(s[pick].ref, two);
            s[pick] = two[0];
            s[ns++] = two[1];
        }
        Bvh4Node n4{};
        for (int i = 0; i < kBvhWidth; ++i) {
            if (i < ns && s[i].ref != kBvhEmpty) {
                n4.lox[i] = s[i].lo[0]; n4.loy[i] = s[i].lo[1]; n4.loz[i] = s[i].lo[2];
                n4.hix[i] = s[i].hi[0]; n4.hiy[i] = s[i].hi[1]; n4.hiz[i] = s[i].hi[2];
                if (s[i].ref >= 0) {
                    const int32_t child = (int32_t)bvh->nodes4.size();
                    bvh->nodes4.push_back(Bvh4Node{});
                    todo.push_back({s[i].ref, child});
                    n4.child[i] = child;
                } else {
                    n4.child[i] = s[i].ref;
                }
            } else {
                // an empty slot carries NaN bounds: its slab distances are NaN, the far distance stays NaN through
                // fminf, and `tn <= tf` is false -- the device needs no separate emptiness test
                n4.lox[i] = n4.loy[i] = n4.loz[i] = NAN; n4.hix[i] = n4.hiy[i] = n4.hiz[i] = NAN;
                n4.child[i] = kBvhEmpty;
            }
        }
        bvh->nodes4[(size_t)me] = n4;
    }
}

// `threads` <= 0: as many as the host offers (at most 16); the tree does not depend on the thread count.
inline void build_bvh(const double* centres, const double* radii, int n, BvhHost* out, int threads = 0) {
    using namespace bvh_detail;
    out->nodes.clear(); out->leaf_idx.clear();
    std::vector<int32_t> order((size_t)n);
    Builder b{centres, radii, order, out};
    for (int i = 0; i < n; ++i) order[i] = i;
    if (kBvhLeafMax == 1) {  // identity leaf list (see Builder::build)
        out->leaf_idx.resize(n);
        for (int i = 0; i < n; ++i) out->leaf_idx[i] = i;
    }
    if (n <= kBvhLeafMax) {  // root must be a node: one real leaf + one empty child
        out->nodes.push_back(BvhNode{});
        const int32_t leaf = n > 0 ? b.build(out->nodes, 0, n) : (int32_t)0x80000000u;
        BvhNode& root = out->nodes[0];
        if (n > 0) store_box(b.bounds(0, n), root.lo0, root.hi0); else empty_box(root.lo0, root.hi0);
        empty_box(root.lo1, root.hi1);
        root.child0 = leaf; root.child1 = (int32_t)0x80000000u;
        collapse_bvh4(out);
        return;
    }
    if (threads <= 0) threads = (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
    int fork = 0;
    while ((1 << fork) < threads) ++fork;   // 2^fork subtrees build concurrently
    out->nodes.reserve((size_t)n);
    b.build(out->nodes, 0, n, 0, fork);  // nodes[0] is the root because the first push_back happens at the top call
    collapse_bvh4(out);
}

// ------------------------------------------------------------------------------------------------------------
// Tie grid: "which spheres' surfaces pass (numerically) through this point?"  answered in O(1).
//
// With the reference's tmin = 0 (programs/main.cc:40) 86 % of all ray casts of the book scene start ON the sphere
// the path just hit and hit that sphere again at t == 0 exactly (the self-hit artefact).  For such a
// cast hittable_list::hit (programs/hittable_list.cc:3-20) can only return another sphere j if j ALSO yields an
// accepted root <= that t (ties go to the later index), which requires the ray origin o to lie on j's surface up to
// rounding: sphere::hit's root (programs/sphere.cc:24,29) is <= t only if dist(o, surface_j) <= t * |dir|, and a
// root of exactly 0 only if |C_j| = ||o - c_j|^2 - r_j^2| is below ~2^-49 (|o| + |c_j|)^2.  The device therefore
// runs the FP64 test of the start sphere first and, when it returns t ~ 0, decides the whole cast with a conservative
// FP32 shell test  | |o - c_j|^2 - r_j^2 | <= tol_j(o) + reach  on the few spheres whose surface can come near o
// (FP64 sphere::hit on those that pass) -- no traversal.  Candidates: up to kTieGiants "giant" spheres (tested for
// every such cast) plus the <= 4 spheres listed in the uniform-grid cell that contains o.  A sphere is listed in every
// cell its padded bounding box overlaps; the padding covers the shell tolerance at the grid's largest |o|, the largest
// reach the device accepts (rho_max), the FP32 rounding of o and of the cell arithmetic.  Cells that would need
// more than four entries carry kTieOverfull and send the cast to the BVH traversal instead (always correct).
constexpr int kTieGiants = 4;
constexpr double kTieCellRadii = 2.0;   // cell size in median sphere radii (doubled until giants / cell budget fit)
constexpr int32_t kTieOverfull = -2;
constexpr double kTieShell = 1.9073486328125e-06;   // 2^-19 = 32 * 2^-24: FP32 error of the shell quantity is <= 2^-24 (14 (o^2 + c^2) + 4 r^2)

struct TieGridHost {
    bool ok = false;             // false: no fast path for this scene (every cast takes the traversal)
    float g0[3] = {0, 0, 0}, g1[3] = {0, 0, 0};   // grid bounds (outside: no non-giant sphere can be a candidate)
    float inv_h = 0.f, rho_max = 0.f;
    int dim[3] = {0, 0, 0};
    int n_giants = 0;
    int32_t giants[kTieGiants] = {-1, -1, -1, -1};
    std::vector<int32_t> cells;  // 4 per cell: sphere indices, -1 = unused; cells[4 i] == kTieOverfull: undecidable here
    std::vector<float> sph;      // 4 per sphere: cx, cy, cz, |r| rounded to nearest
};

inline void build_tie_grid(const double* c, const double* r, int n, TieGridHost* out) {
    *out = TieGridHost();
    out->sph.resize((size_t)4 * (n > 0 ? n : 1));
    if (n <= 0) return;
    std::vector<double> rad((size_t)n);
    double cmax = 0.0;
    for (int k = 0; k < n; ++k) {
        rad[(size_t)k] = std::fabs(r[k]);
        for (int a = 0; a < 3; ++a) { out->sph[4 * (size_t)k + a] = (float)c[3 * k + a]; cmax = std::max(cmax, std::fabs(c[3 * k + a])); }
        out->sph[4 * (size_t)k + 3] = (float)rad[(size_t)k];
        if (!std::isfinite(c[3 * k]) || !std::isfinite(c[3 * k + 1]) || !std::isfinite(c[3 * k + 2]) || !std::isfinite(r[k])) return;
    }
    std::vector<double> sorted(rad);
    std::nth_element(sorted.begin(), sorted.begin() + n / 2, sorted.end());
    const double r_med = sorted[(size_t)n / 2];
    if (!(r_med >= 1e-12) || !(cmax <= 1e12)) return;   // FP32 shell arithmetic needs moderate magnitudes

    double h = kTieCellRadii * r_med;
    if (const char* e = std::getenv("RT_TIE_CELL")) { const double v = std::atof(e); if (v > 0.1 && v < 100.0) h = v * r_med; }   // tuning experiments only
    for (int attempt = 0; attempt < 24; ++attempt, h *= 2.0) {
        const double rho_max = h / 64.0;
        // pass 1: giants (by the number of cells their box would cover) and the bounds of everything else
        std::vector<char> giant((size_t)n, 0);
        int ng = 0;
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int k = 0; k < n; ++k) {
            const double cells_axis = 2.0 * rad[(size_t)k] / h + 2.0;
            if (cells_axis * cells_axis * cells_axis > 1000.0) { giant[(size_t)k] = 1; ++ng; continue; }   // > 8 cells across
            for (int a = 0; a < 3; ++a) {
                lo[a] = std::min(lo[a], c[3 * k + a] - rad[(size_t)k]);
                hi[a] = std::max(hi[a], c[3 * k + a] + rad[(size_t)k]);
            }
        }
        if (ng > kTieGiants) continue;
        TieGridHost g;
        g.sph = out->sph;
        g.n_giants = ng;
        for (int k = 0, i = 0; k < n; ++k) if (giant[(size_t)k]) g.giants[i++] = k;
        g.rho_max = bvh_detail::down(rho_max);
        if (ng == n) {   // nothing but giants: an empty grid (every o is "outside")
            g.dim[0] = g.dim[1] = g.dim[2] = 0; g.inv_h = 0.f;
            g.g0[0] = g.g0[1] = g.g0[2] = INFINITY; g.g1[0] = g.g1[1] = g.g1[2] = -INFINITY;
            g.cells.assign(4, -1);
            g.ok = true;
            *out = g;
            return;
        }
        // largest |o| inside the (generously padded) grid, for the shell tolerance the padding must cover
        double omax2 = 0.0;
        for (int a = 0; a < 3; ++a) { const double m = std::max(std::fabs(lo[a]), std::fabs(hi[a])) + 2.0 * h; omax2 += m * m; }
        const double omax = std::sqrt(omax2);
        const double delta = 0.02 * h + 4.76837158203125e-07 * (omax + h);   // FP32 rounding of o and of the cell arithmetic (2^-21 |o|)
        std::vector<double> pad((size_t)n, 0.0);
        double pad_max = 0.0;
        for (int k = 0; k < n; ++k) {
            if (giant[(size_t)k]) continue;
            const double c2 = c[3 * k] * c[3 * k] + c[3 * k + 1] * c[3 * k + 1] + c[3 * k + 2] * c[3 * k + 2], r2 = rad[(size_t)k] * rad[(size_t)k];
            const double T = 2.0 * kTieShell * (1.01 * omax2 + c2 + r2);                 // >= tol_j(o) for every o in the grid
            const double reach = rho_max * (2.0 * rad[(size_t)k] + rho_max) * 1.01;      // the reach term of the device test
            pad[(size_t)k] = std::sqrt(r2 + T + reach) - rad[(size_t)k] + rho_max + delta;
            pad_max = std::max(pad_max, pad[(size_t)k]);
        }
        double ext[3];
        long long total = 1;
        bool fits = true;
        for (int a = 0; a < 3; ++a) {
            lo[a] -= pad_max + 0.5 * h; hi[a] += pad_max + 0.5 * h;
            g.g0[a] = bvh_detail::down(lo[a]);
            ext[a] = hi[a] - (double)g.g0[a];
            const double cells_d = std::ceil(ext[a] / h) + 1.0;
            if (!(cells_d < 1e6)) { fits = false; break; }
            g.dim[a] = (int)cells_d;
            total *= g.dim[a];
            if (total > (1ll << 21)) { fits = false; break; }
        }
        if (!fits) continue;
        g.inv_h = (float)(1.0 / h);
        const double hf = 1.0 / (double)g.inv_h;   // the cell size the device effectively uses
        for (int a = 0; a < 3; ++a) g.g1[a] = bvh_detail::down((double)g.g0[a] + (g.dim[a] - 0.5) * hf);   // strictly inside the last cell
        g.cells.assign((size_t)total * 4, -1);
        size_t overfull = 0;
        for (int k = 0; k < n; ++k) {
            if (giant[(size_t)k]) continue;
            int i0[3], i1[3];
            for (int a = 0; a < 3; ++a) {
                const double x0 = (c[3 * k + a] - rad[(size_t)k] - pad[(size_t)k] - (double)g.g0[a]) / hf;
                const double x1 = (c[3 * k + a] + rad[(size_t)k] + pad[(size_t)k] - (double)g.g0[a]) / hf;
                i0[a] = std::max(0, (int)std::floor(x0)); i1[a] = std::min(g.dim[a] - 1, (int)std::floor(x1));
            }
            for (int z = i0[2]; z <= i1[2]; ++z)
                for (int y = i0[1]; y <= i1[1]; ++y)
                    for (int x = i0[0]; x <= i1[0]; ++x) {
                        int32_t* cell = &g.cells[4 * (((size_t)z * g.dim[1] + y) * g.dim[0] + x)];
                        if (cell[0] == kTieOverfull) continue;
                        int e = 0;
                        while (e < 4 && cell[e] >= 0) ++e;
                        if (e < 4) cell[e] = k;
                        else { cell[0] = kTieOverfull; ++overfull; }
                    }
        }
        // a grid where a large part of the occupied cells is overfull decides little: coarser cells do not help
        // (lists only grow), so accept it as it is -- overfull cells fall back to the traversal
        (void)overfull;
        g.ok = true;
        *out = g;
        return;
    }
}

}  // namespace rt
