/* oracle/rt_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference hot path; see rt_oracle.h for status
 * (parity PINNED against oracle/_ref/libref.so and tests/golden/).
 * Every function cites the reference lines it follows.  All arithmetic is
 * FP64 with one rounding per * and + in the reference's evaluation order
 * (built with -ffp-contract=off, x86-64 SSE2, no -march flags).
 */
#include "rt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ vec3 */
typedef struct { double e[3]; } v3;

static inline v3 v3_make(double a, double b, double c) { v3 r = {{a, b, c}}; return r; }
/* programs/vec3.h:126-129 */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.e[0] + b.e[0], a.e[1] + b.e[1], a.e[2] + b.e[2]); }
/* programs/vec3.h:131-134 */
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.e[0] - b.e[0], a.e[1] - b.e[1], a.e[2] - b.e[2]); }
/* programs/vec3.h:136-139 (v*t forwards to t*v, :146-149) */
static inline v3 v3_scale(double t, v3 a) { return v3_make(t * a.e[0], t * a.e[1], t * a.e[2]); }
/* programs/vec3.h:151-154: division is multiplication by the reciprocal */
static inline v3 v3_div(v3 a, double t) { return v3_scale(1 / t, a); }
/* programs/vec3.h:29-32 */
static inline v3 v3_neg(v3 a) { return v3_make(-a.e[0], -a.e[1], -a.e[2]); }
/* programs/vec3.h:156-159: (u0*v0 + u1*v1) + u2*v2 */
static inline double v3_dot(v3 a, v3 b) { return a.e[0] * b.e[0] + a.e[1] * b.e[1] + a.e[2] * b.e[2]; }
/* programs/vec3.h:63-71 */
static inline double v3_len2(v3 a) { return a.e[0] * a.e[0] + a.e[1] * a.e[1] + a.e[2] * a.e[2]; }
/* programs/vec3.h:172-175 */
static inline v3 v3_unit(v3 a) { return v3_div(a, sqrt(v3_len2(a))); }

/* ------------------------------------------------------------------- RNG */
static const uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
static const uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    int mode;
    uint64_t state;              /* RAND15: splitmix64 state (oracle/ref_rng.cc) */
    uint32_t key[2], pix, smp;   /* PHILOX: key = seed, counter = (pix, smp, blk, 0) */
    uint32_t blk;
} rng_t;

/* oracle/ref_rng.cc: the shim's rand() */
static inline int rng_rand15(rng_t* g) {
    uint64_t z = (g->state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (int)(z >> 49);
}
/* programs/random.h:4-8 in the RAND_MAX == 0x7fff environment it was written for */
static inline double rng_double15(rng_t* g) { return (double)rng_rand15(g) / (0x7fff + 1); }

static inline void rng_philox_block(rng_t* g, uint32_t w[4]) {
    uint32_t ctr[4] = {g->pix, g->smp, g->blk, 0u};
    orc_philox4x32_10(ctr, g->key, w);
    g->blk++;
}
static inline double u32_to_unit(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

/* the two jitter draws of programs/main.cc:80-81 (u first, then v) */
static inline void rng_jitter(rng_t* g, double* xu, double* xv) {
    if (g->mode == ORC_RNG_RAND15) { *xu = rng_double15(g); *xv = rng_double15(g); }
    else { uint32_t w[4]; rng_philox_block(g, w); *xu = u32_to_unit(w[0]); *xv = u32_to_unit(w[1]); }
}

/* programs/vec3.h:78-81 vec3::random(-1,1) with random_double(min,max) = min + (max-min)*xi
 * (programs/random.h:10-14), RAND15 mode: the three calls are evaluated in the order the reference build
 * evaluates them (g++ 13 x86-64: last argument first), which is what makes this file reproduce
 * libref.so bit for bit. */
static inline v3 rng_cube15(rng_t* g) {
    double z = -1.0 + (1.0 - -1.0) * rng_double15(g);
    double y = -1.0 + (1.0 - -1.0) * rng_double15(g);
    double x = -1.0 + (1.0 - -1.0) * rng_double15(g);
    return v3_make(x, y, z);
}

/* PHILOX mode (shared with the CUDA path, include/rt.h): each block carries two tries of the rejection
 * loop as six 21-bit uniforms (the reference's rand() has 15 bits): try A = top 21 bits of words 0,1,2;
 * try B = low 11 bits of words 0,1,2 extended by 10-bit fields of word 3. */
static inline v3 cube21(uint32_t fx, uint32_t fy, uint32_t fz) {
    const double s21 = 1.0 / 2097152.0;
    return v3_make(-1.0 + (1.0 - -1.0) * ((double)fx * s21), -1.0 + (1.0 - -1.0) * ((double)fy * s21),
                   -1.0 + (1.0 - -1.0) * ((double)fz * s21));
}

/* programs/vec3.h:83-95: reject only if len^2 > 1.  PHILOX mode: one block per bounce gives tries A and B;
 * bounces that reject both continue with xorshift128 seeded by the block's words, three outputs per try. */
static inline v3 random_in_unit_sphere(rng_t* g) {
    if (g->mode == ORC_RNG_RAND15) {
        for (;;) {
            v3 v = rng_cube15(g);
            if (v3_len2(v) > 1.0) continue;
            return v;
        }
    }
    uint32_t w[4];
    rng_philox_block(g, w);
    v3 a = cube21(w[0] >> 11, w[1] >> 11, w[2] >> 11);
    if (!(v3_len2(a) > 1.0)) return a;
    v3 b = cube21(((w[0] & 0x7ffu) << 10) | (w[3] >> 22), ((w[1] & 0x7ffu) << 10) | ((w[3] >> 12) & 0x3ffu),
                  ((w[2] & 0x7ffu) << 10) | ((w[3] >> 2) & 0x3ffu));
    if (!(v3_len2(b) > 1.0)) return b;
    uint32_t x0 = w[0], x1 = w[1], x2 = w[2], x3 = w[3];
    if ((x0 | x1 | x2 | x3) == 0u) x0 = 1u;
    for (;;) {
        uint32_t f[3];
        for (int i = 0; i < 3; ++i) {
            uint32_t t = x3;
            const uint32_t s0 = x0;
            x3 = x2; x2 = x1; x1 = s0;
            t ^= t << 11; t ^= t >> 8;
            x0 = t ^ s0 ^ (s0 >> 19);
            f[i] = x0 >> 11;
        }
        v3 c = cube21(f[0], f[1], f[2]);
        if (!(v3_len2(c) > 1.0)) return c;
    }
}
/* programs/vec3.h:102-109: keep if dot > 0, else negate (dot == 0 negates) */
static inline v3 random_in_hemisphere(rng_t* g, v3 normal) {
    v3 rv = random_in_unit_sphere(g);
    if (v3_dot(rv, normal) > 0) return rv;
    return v3_neg(rv);
}

/* programs/vec3.h:97-100: unit_vector(random_in_unit_sphere()) -- present but unused in the reference's
 * main.cc; the book's Lambertian scatter (orc_shading.scatter_mode == 1) */
static inline v3 random_unit_vector(rng_t* g) { return v3_unit(random_in_unit_sphere(g)); }

/* --------------------------------------------------------------- geometry */
typedef struct { v3 o, d; } ray_t;
typedef struct { v3 p, normal; double t; int front_face; double C; } rec_t;

/* programs/ray.h:25-28: orig + (t*dir) */
static inline v3 ray_at(const ray_t* r, double t) { return v3_add(r->o, v3_scale(t, r->d)); }

/* programs/sphere.cc:3-40.  rec->C additionally keeps the value of C (sphere.cc:11) for the
 * exact early-out; it does not influence the result. */
static inline int sphere_hit(v3 centre, double radius, const ray_t* r, double tmin, double tmax, rec_t* rec) {
    v3 b = r->d;
    v3 a_minus_c = v3_sub(r->o, centre);
    double A = v3_dot(b, b);
    double HALF_B = v3_dot(b, a_minus_c);
    double C = v3_dot(a_minus_c, a_minus_c) - radius * radius;
    double discriminant = HALF_B * HALF_B - A * C;
    if (discriminant < 0) return 0;
    double sqrt_d = sqrt(discriminant);
    double t = (-HALF_B - sqrt_d) / A;
    if (t < tmin || t > tmax) {
        t = (-HALF_B + sqrt_d) / A;
        if (t < tmin || t > tmax) return 0;
    }
    rec->p = ray_at(r, t);
    /* programs/hittable.h:14-18 set_face_normal with outward = (p - centre) / radius */
    v3 outward = v3_div(v3_sub(rec->p, centre), radius);
    rec->front_face = v3_dot(r->d, outward) < 0;
    rec->normal = rec->front_face ? outward : v3_neg(outward);
    rec->t = t;
    rec->C = C;
    return 1;
}

typedef struct { int n; const double* c; const double* r; } world_t;

/* programs/hittable_list.cc:3-20: list order, shrinking tmax, closed interval -> ties go to the later
 * object.  Returns the index of the object whose record was kept, or -1. */
static inline int list_hit(const world_t* w, const ray_t* r, double tmin, double tmax, rec_t* rec) {
    rec_t tmp;
    int hit_idx = -1;
    double closest_so_far = tmax;
    for (int k = 0; k < w->n; ++k) {
        if (sphere_hit(v3_make(w->c[3 * k], w->c[3 * k + 1], w->c[3 * k + 2]), w->r[k], r, tmin, closest_so_far, &tmp)) {
            hit_idx = k;
            closest_so_far = tmp.t;
            *rec = tmp;
        }
    }
    return hit_idx;
}

/* programs/main.cc:34-49 ray_color, recursion unrolled, with the constants the reference hard-codes as
 * parameters (orc_shading; the defaults are main.cc's: tmin 0 at :40, albedo 0.5 at :43, sky colours at :48,
 * vec3::random_in_hemisphere at :42).  The recursion returns albedo * (albedo * (... * sky)): one
 * multiplication per bounce applied to the sky colour from the innermost call outwards, which is what the
 * loop at the end does (for albedo = 0.5 every product is exact, so this equals the 0.5^k * sky the first
 * version of this file computed; checked bit for bit against libref.so).  depth < 0 -> black (main.cc:36).
 * early_out (tmin == 0 only): if the kept hit has t == 0 and C == 0 exactly, the next origin equals this
 * one, every later cast hits that sphere at t == 0 again, and the path must end black: return 0 now. */
static inline v3 ray_color(ray_t r, const world_t* w, int depth, rng_t* g, const orc_shading* sh, int early_out,
                           orc_stats* st, int* first_hit) {
    int bounces = 0;
    int first = 1;
    for (;;) {
        if (depth < 0) return v3_make(0, 0, 0);   /* albedo * ... * 0 == 0 for every finite albedo >= 0 */
        rec_t rec;
        rec.t = 0; rec.C = 1;
        st->casts += 1;
        int k = list_hit(w, &r, sh->tmin, INFINITY, &rec);
        if (first) { *first_hit = k; first = 0; }
        if (k < 0) break;
        if (early_out && sh->tmin == 0 && rec.t == 0 && rec.C == 0) { st->early_outs += 1; return v3_make(0, 0, 0); }
        /* main.cc:42-43: target = (p + normal) + scatter; next ray = (p, target - p) */
        v3 sc = sh->scatter_mode == ORC_SCATTER_LAMBERTIAN ? random_unit_vector(g) : random_in_hemisphere(g, rec.normal);
        v3 target = v3_add(v3_add(rec.p, rec.normal), sc);
        r.o = rec.p;
        r.d = v3_sub(target, rec.p);
        ++bounces;
        --depth;
    }
    /* main.cc:46-48 */
    v3 ud = v3_unit(r.d);
    double t = 0.5 * (ud.e[1] + 1.0);
    v3 c = v3_add(v3_scale(1.0 - t, v3_make(sh->sky_a[0], sh->sky_a[1], sh->sky_a[2])),
                  v3_scale(t, v3_make(sh->sky_b[0], sh->sky_b[1], sh->sky_b[2])));
    for (int b = 0; b < bounces; ++b) c = v3_scale(sh->albedo, c);   /* main.cc:43, once per unwound call */
    return c;
}

void orc_default_shading(orc_shading* sh) {
    sh->tmin = 0.0;                                            /* programs/main.cc:40 */
    sh->albedo = 0.5;                                          /* programs/main.cc:43 */
    sh->sky_a[0] = 1.0; sh->sky_a[1] = 1.0; sh->sky_a[2] = 1.0;  /* programs/main.cc:48 */
    sh->sky_b[0] = 0.5; sh->sky_b[1] = 0.7; sh->sky_b[2] = 1.0;
    sh->scatter_mode = ORC_SCATTER_HEMISPHERE;                 /* programs/main.cc:42 */
}

/* programs/camera.h:25-28: dir = ((llc + u*horizontal) + v*vertical) - origin */
static inline ray_t get_ray(const double* cam12, double u, double v) {
    v3 origin = v3_make(cam12[0], cam12[1], cam12[2]);
    v3 llc = v3_make(cam12[3], cam12[4], cam12[5]);
    v3 hor = v3_make(cam12[6], cam12[7], cam12[8]);
    v3 ver = v3_make(cam12[9], cam12[10], cam12[11]);
    ray_t r;
    r.o = origin;
    r.d = v3_sub(v3_add(v3_add(llc, v3_scale(u, hor)), v3_scale(v, ver)), origin);
    return r;
}

/* programs/raytracer.h:19-23: std::min(std::max(v, lo), hi) */
static inline double clampd(double v, double lo, double hi) {
    double m = (v < lo) ? lo : v;   /* std::max(v, lo) */
    return (hi < m) ? hi : m;       /* std::min(m, hi) */
}
/* programs/color.h:8-24 */
static inline void write_color(v3 pixel_color, int spp, int out[3]) {
    double one_over_samples = 1.0 / spp;
    for (int c = 0; c < 3; ++c) {
        double x = sqrt(pixel_color.e[c] * one_over_samples);
        out[c] = (int)(255.999 * clampd(x, 0.0, 0.999));
    }
}

static inline uint64_t row_seed(uint64_t seed, int j) { return seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(j + 1)); }

/* ------------------------------------------------------------- entry points */
void orc_default_camera(double* cam12, double* aspect) {
    /* programs/camera.h:11-23 */
    double aspect_ratio = 16.0 / 9.0;
    const double viewport_height = 2.0;
    const double viewport_width = viewport_height * aspect_ratio;
    const double focal_length = 1.0;
    v3 origin = v3_make(0, 0, 0);
    v3 horizontal = v3_make(viewport_width, 0, 0);
    v3 vertical = v3_make(0, viewport_height, 0);
    v3 llc = v3_add(v3_sub(v3_sub(origin, v3_div(horizontal, 2.0)), v3_div(vertical, 2.0)), v3_make(0, 0, -focal_length));
    memcpy(cam12 + 0, origin.e, 24); memcpy(cam12 + 3, llc.e, 24);
    memcpy(cam12 + 6, horizontal.e, 24); memcpy(cam12 + 9, vertical.e, 24);
    if (aspect) *aspect = aspect_ratio;
}

long orc_main_ppm(uint64_t seed, char* buf, long cap) {
    /* programs/main.cc:51-92 */
    double cam12[12], aspect;
    orc_default_camera(cam12, &aspect);
    const int W = 400, H = (int)(W / aspect);
    const double centres[6] = {0, 0, -1, 0, -100.5, 0};
    const double radii[2] = {0.5, 100.0};
    world_t w = {2, centres, radii};
    const int spp = 100, max_depth = 50;
    rng_t g; memset(&g, 0, sizeof g); g.mode = ORC_RNG_RAND15; g.state = seed;
    orc_stats st; memset(&st, 0, sizeof st);
    orc_shading sh; orc_default_shading(&sh);
    long len = 0;
    char line[64];
    int m = snprintf(line, sizeof line, "P3\n%d %d\n255\n", W, H);
    if (len + m <= cap) memcpy(buf + len, line, (size_t)m);
    len += m;
    for (int j = H - 1; j >= 0; --j)
        for (int i = 0; i < W; ++i) {
            v3 px = v3_make(0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                double xu, xv; rng_jitter(&g, &xu, &xv);
                double u = ((double)i + xu) / (W - 1);
                double v = ((double)j + xv) / (H - 1);
                int fh;
                px = v3_add(px, ray_color(get_ray(cam12, u, v), &w, max_depth, &g, &sh, 0, &st, &fh));
            }
            int q[3]; write_color(px, spp, q);
            m = snprintf(line, sizeof line, "%d %d %d\n", q[0], q[1], q[2]);
            if (len + m <= cap) memcpy(buf + len, line, (size_t)m);
            len += m;
        }
    return len <= cap ? len : -len;
}

void orc_render_rows_ex(const double* centres, const double* radii, int n, const double* cam12,
                        int W, int H, int spp, int max_depth, uint64_t seed, int rng_mode, int early_out,
                        const orc_shading* shading, int j0, int j1, int nthreads, uint8_t* rgb, double* rgb_sum,
                        orc_stats* stats) {
    world_t w = {n, centres, radii};
    orc_shading sh;
    if (shading) sh = *shading; else orc_default_shading(&sh);
    double t_casts = 0, t_black = 0, t_prim = 0, t_eo = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : t_casts, t_black, t_prim, t_eo)
#endif
    for (int j = j1 - 1; j >= j0; --j) {
        rng_t g; memset(&g, 0, sizeof g);
        g.mode = rng_mode;
        g.state = row_seed(seed, j);
        g.key[0] = (uint32_t)seed; g.key[1] = (uint32_t)(seed >> 32);
        orc_stats st; memset(&st, 0, sizeof st);
        for (int i = 0; i < W; ++i) {
            v3 px = v3_make(0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                g.pix = (uint32_t)(j * W + i); g.smp = (uint32_t)s; g.blk = 0;
                double xu, xv; rng_jitter(&g, &xu, &xv);
                double u = ((double)i + xu) / (W - 1);
                double v = ((double)j + xv) / (H - 1);
                int fh = -1;
                v3 c = ray_color(get_ray(cam12, u, v), &w, max_depth, &g, &sh, early_out, &st, &fh);
                if (c.e[0] == 0 && c.e[1] == 0 && c.e[2] == 0) st.black += 1;
                if (fh >= 0) st.primary_hits += 1;
                px = v3_add(px, c);   /* programs/vec3.h:42-48 operator+= */
            }
            size_t o = (size_t)(H - 1 - j) * W + i;
            if (rgb) {
                int q[3]; write_color(px, spp, q);
                rgb[3 * o] = (uint8_t)q[0]; rgb[3 * o + 1] = (uint8_t)q[1]; rgb[3 * o + 2] = (uint8_t)q[2];
            }
            if (rgb_sum) { rgb_sum[3 * o] = px.e[0]; rgb_sum[3 * o + 1] = px.e[1]; rgb_sum[3 * o + 2] = px.e[2]; }
        }
        t_casts += st.casts; t_black += st.black; t_prim += st.primary_hits; t_eo += st.early_outs;
    }
    if (stats) {
        stats->samples = (double)(j1 - j0) * W * spp;
        stats->casts = t_casts; stats->black = t_black; stats->primary_hits = t_prim; stats->early_outs = t_eo;
    }
}

void orc_render_rows(const double* centres, const double* radii, int n, const double* cam12,
                     int W, int H, int spp, int max_depth, uint64_t seed, int rng_mode, int early_out,
                     int j0, int j1, int nthreads, uint8_t* rgb, double* rgb_sum, orc_stats* stats) {
    orc_render_rows_ex(centres, radii, n, cam12, W, H, spp, max_depth, seed, rng_mode, early_out, NULL, j0, j1, nthreads,
                       rgb, rgb_sum, stats);
}

void orc_primary_hits(const double* centres, const double* radii, int n, const double* cam12,
                      int W, int H, int32_t* idx, double* t) {
    world_t w = {n, centres, radii};
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int j = H - 1; j >= 0; --j)
        for (int i = 0; i < W; ++i) {
            double u = ((double)i + 0.5) / (W - 1);
            double v = ((double)j + 0.5) / (H - 1);
            ray_t r = get_ray(cam12, u, v);
            rec_t rec;
            size_t o = (size_t)(H - 1 - j) * W + i;
            int k = list_hit(&w, &r, 0, INFINITY, &rec);
            idx[o] = k;
            t[o] = k >= 0 ? rec.t : INFINITY;
        }
}

static void rec_store(double* o, int hit, const rec_t* rec) {
    if (hit) {
        o[0] = rec->t; memcpy(o + 1, rec->p.e, 24); memcpy(o + 4, rec->normal.e, 24); o[7] = rec->front_face ? 1.0 : 0.0;
    } else {
        for (int e = 0; e < 8; ++e) o[e] = 0.0;
    }
}

void orc_hit_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                   int nrays, double tmin, double tmax, int32_t* idx, double* rec_out) {
    world_t w = {n, centres, radii};
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int q = 0; q < nrays; ++q) {
        ray_t r = {v3_make(org[3 * q], org[3 * q + 1], org[3 * q + 2]), v3_make(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2])};
        rec_t rec;
        int k = list_hit(&w, &r, tmin, tmax, &rec);
        idx[q] = k;
        rec_store(rec_out + 8 * (size_t)q, k >= 0, &rec);
    }
}

void orc_sphere_hit_batch(const double* centre, const double* radius, const double* org, const double* dir,
                          int nq, double tmin, double tmax, int32_t* hit, double* rec_out) {
    for (int q = 0; q < nq; ++q) {
        ray_t r = {v3_make(org[3 * q], org[3 * q + 1], org[3 * q + 2]), v3_make(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2])};
        rec_t rec;
        int h = sphere_hit(v3_make(centre[3 * q], centre[3 * q + 1], centre[3 * q + 2]), radius[q], &r, tmin, tmax, &rec);
        hit[q] = h;
        rec_store(rec_out + 8 * (size_t)q, h, &rec);
    }
}

void orc_ray_color_batch_ex(const double* centres, const double* radii, int n, const double* org, const double* dir,
                            const uint64_t* seeds, int rng_mode, int early_out, const orc_shading* shading, int nrays,
                            int depth, double* rgb_out, orc_stats* stats) {
    world_t w = {n, centres, radii};
    orc_shading sh;
    if (shading) sh = *shading; else orc_default_shading(&sh);
    orc_stats st; memset(&st, 0, sizeof st);
    for (int q = 0; q < nrays; ++q) {
        ray_t r = {v3_make(org[3 * q], org[3 * q + 1], org[3 * q + 2]), v3_make(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2])};
        rng_t g; memset(&g, 0, sizeof g);
        g.mode = rng_mode;
        if (rng_mode == ORC_RNG_RAND15) g.state = seeds[q];
        else { g.key[0] = (uint32_t)seeds[0]; g.key[1] = (uint32_t)(seeds[0] >> 32); g.pix = (uint32_t)q; g.smp = 0; g.blk = 1; }
        int fh = -1;
        v3 c = ray_color(r, &w, depth, &g, &sh, early_out, &st, &fh);
        st.samples += 1;
        if (c.e[0] == 0 && c.e[1] == 0 && c.e[2] == 0) st.black += 1;
        if (fh >= 0) st.primary_hits += 1;
        memcpy(rgb_out + 3 * (size_t)q, c.e, 24);
    }
    if (stats) *stats = st;
}

void orc_ray_color_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                         const uint64_t* seeds, int rng_mode, int early_out, int nrays, int depth, double* rgb_out,
                         orc_stats* stats) {
    orc_ray_color_batch_ex(centres, radii, n, org, dir, seeds, rng_mode, early_out, NULL, nrays, depth, rgb_out, stats);
}

void orc_get_ray_batch(const double* cam12, const double* uv, int nq, double* out) {
    for (int q = 0; q < nq; ++q) {
        ray_t r = get_ray(cam12, uv[2 * q], uv[2 * q + 1]);
        memcpy(out + 6 * (size_t)q, r.o.e, 24);
        memcpy(out + 6 * (size_t)q + 3, r.d.e, 24);
    }
}

void orc_write_color_batch(const double* rgb_sum, int nq, int spp, int32_t* out) {
    for (int q = 0; q < nq; ++q) {
        int v[3];
        write_color(v3_make(rgb_sum[3 * q], rgb_sum[3 * q + 1], rgb_sum[3 * q + 2]), spp, v);
        out[3 * q] = v[0]; out[3 * q + 1] = v[1]; out[3 * q + 2] = v[2];
    }
}

void orc_random_in_hemisphere_batch(const double* normal, const uint64_t* seeds, int nq, double* out) {
    for (int q = 0; q < nq; ++q) {
        rng_t g; memset(&g, 0, sizeof g); g.mode = ORC_RNG_RAND15; g.state = seeds[q];
        v3 v = random_in_hemisphere(&g, v3_make(normal[3 * q], normal[3 * q + 1], normal[3 * q + 2]));
        memcpy(out + 3 * (size_t)q, v.e, 24);
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
