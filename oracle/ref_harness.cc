/* oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin extern "C" driver around the UNMODIFIED reference classes, compiled
 * (with oracle/shim.h force-included) against the headers where they lie in
 * /root/reference/programs and linked with main.o / sphere.o /
 * hittable_list.o built from the reference .cc files (see oracle/Makefile).
 * The result, oracle/_ref/libref.so, is "the reference itself run here":
 *   - it pins oracle/rt_oracle.c (the C restatement) bit for bit, and
 *   - it is the `cpu_baseline.kind == "reference"` arm of bench.py.
 * Nothing in the product path (petershirleyraytracer_b200/) may load it.
 *
 * Every number is produced by the reference's own code:
 *   camera::get_ray            programs/camera.h:25-28
 *   hittable_list::hit         programs/hittable_list.cc:3-20
 *   sphere::hit                programs/sphere.cc:3-40
 *   ray_color                  programs/main.cc:34-49   (external linkage)
 *   write_color                programs/color.h:8-24    (defined in main.o)
 *   random_double / vec3::random_in_hemisphere  programs/random.h, vec3.h
 * The only code added here is the pixel loop of programs/main.cc:72-88,
 * re-stated so that scene, camera, size and spp become arguments, and rows
 * can run on several threads (one shim RNG stream per row).
 */
#include "raytracer.h"
#include "camera.h"
#include "hittable_list.h"
#include "sphere.h"

#include <cstring>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

/* defined in the reference's main.cc (compiled with -Dmain=reference_main) */
color ray_color(const ray& r, const hittable& world, int depth);
void write_color(std::ostream& out, const color& pixel_color, int samples_per_pixel);
int reference_main();

namespace {

/* records which child of the list produced the last successful hit: the
 * reference's own hittable_list::hit leaves the id of the closest one. */
struct tagged_sphere : public hittable {
    sphere inner;
    int id;
    int* last;
    tagged_sphere(const point3& c, double r, int id_, int* last_) : inner(c, r), id(id_), last(last_) {}
    bool hit(const ray& r, double tmin, double tmax, hit_record& rec) const override {
        if (inner.hit(r, tmin, tmax, rec)) { *last = id; return true; }
        return false;
    }
};

/* counts ray casts (calls of world.hit from ray_color) */
struct counting_world : public hittable {
    const hittable& w;
    mutable long long casts = 0;
    explicit counting_world(const hittable& w_) : w(w_) {}
    bool hit(const ray& r, double tmin, double tmax, hit_record& rec) const override {
        ++casts;
        return w.hit(r, tmin, tmax, rec);
    }
};

void build_world(hittable_list& world, const double* c, const double* rad, int n) {
    for (int k = 0; k < n; ++k)
        world.add(make_shared<sphere>(point3(c[3 * k], c[3 * k + 1], c[3 * k + 2]), rad[k]));
}

void set_camera(camera& cam, const double* f) {
    cam.origin = point3(f[0], f[1], f[2]);
    cam.lower_left_corner = point3(f[3], f[4], f[5]);
    cam.horizontal = vec3(f[6], f[7], f[8]);
    cam.vertical = vec3(f[9], f[10], f[11]);
}

inline uint64_t row_seed(uint64_t seed, int j) {
    return seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(j + 1));
}

/* The constants main.cc:34-49 hard-codes, as arguments (layout == orc_shading of oracle/rt_oracle.h). */
struct ref_shading {
    double tmin;
    double albedo;
    double sky_a[3], sky_b[3];
    int scatter_mode;   /* 0: vec3::random_in_hemisphere (main.cc:42); 1: vec3::random_unit_vector (vec3.h:97-100) */
};

/* ray_color with those arguments, every operation done by the reference's own classes and operators
 * (hittable::hit, vec3 arithmetic, unit_vector, vec3::random_*), in the shape of main.cc:34-49.  With the
 * reference's constants it returns what the reference's ray_color returns, bit for bit
 * (tests/test_oracle_vs_ref.py); it pins the parameterised oracle where main.cc's literals cannot. */
color ray_color_param(const ray& r, const hittable& world, int depth, const ref_shading& sh) {
    if (depth < 0) return color(0, 0, 0);
    hit_record rec;
    if (world.hit(r, sh.tmin, infinity, rec)) {
        vec3 bounce = sh.scatter_mode == 1 ? vec3::random_unit_vector() : vec3::random_in_hemisphere(rec.normal);
        vec3 target = rec.p + rec.normal + bounce;
        return sh.albedo * ray_color_param(ray(rec.p, target - rec.p), world, depth - 1, sh);
    }
    vec3 dir = unit_vector(r.direction());
    double t = 0.5 * (dir.y() + 1.0);
    return (1.0 - t) * color(sh.sky_a[0], sh.sky_a[1], sh.sky_a[2]) + t * color(sh.sky_b[0], sh.sky_b[1], sh.sky_b[2]);
}

}  // namespace

extern "C" {

/* Runs the reference's own main() (default scene, 400x225x100spp) with stdout
 * captured; returns the PPM length (or -needed if buf is too small). */
long ref_main_ppm(uint64_t seed, char* buf, long cap) {
    std::ostringstream cap_out, cap_err;
    std::streambuf* old_out = std::cout.rdbuf(cap_out.rdbuf());
    std::streambuf* old_err = std::cerr.rdbuf(cap_err.rdbuf());
    oracle_seed(seed);
    reference_main();
    std::cout.rdbuf(old_out);
    std::cerr.rdbuf(old_err);
    const std::string s = cap_out.str();
    if ((long)s.size() > cap) return -(long)s.size();
    std::memcpy(buf, s.data(), s.size());
    return (long)s.size();
}

/* The pixel loop of main.cc:72-88 over rows j in [j0, j1) (j counted from the
 * bottom, like the reference).  rgb is the full W*H*3 frame, row 0 = top.
 * stats[0]=samples, [1]=casts, [2]=samples that returned exactly black. */
void ref_render_rows(const double* centres, const double* radii, int n, const double* cam12,
                     int W, int H, int spp, int max_depth, uint64_t seed, int j0, int j1,
                     int nthreads, uint8_t* rgb, double* stats) {
    hittable_list world;
    build_world(world, centres, radii, n);
    camera cam;
    set_camera(cam, cam12);
    long long tot_casts = 0, tot_black = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_casts, tot_black)
#endif
    for (int j = j1 - 1; j >= j0; --j) {
        counting_world cw(world);
        oracle_seed(row_seed(seed, j));
        long long black = 0;
        for (int i = 0; i < W; ++i) {
            color pixel_color(0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                double u = ((double)i + random_double()) / (W - 1);
                double v = ((double)j + random_double()) / (H - 1);
                color c = ray_color(cam.get_ray(u, v), cw, max_depth);
                if (c.x() == 0 && c.y() == 0 && c.z() == 0) ++black;
                pixel_color += c;
            }
            std::ostringstream os;
            write_color(os, pixel_color, spp);
            int r = 0, g = 0, b = 0;
            std::istringstream is(os.str());
            is >> r >> g >> b;
            uint8_t* px = rgb + ((size_t)(H - 1 - j) * W + i) * 3;
            px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b;
        }
        tot_casts += cw.casts;
        tot_black += black;
    }
    if (stats) {
        stats[0] = (double)(j1 - j0) * W * spp;
        stats[1] = (double)tot_casts;
        stats[2] = (double)tot_black;
    }
}

/* Primary (camera) rays through pixel centres: u=(i+0.5)/(W-1), v=(j+0.5)/(H-1).
 * idx/t are W*H, row 0 = top; idx = -1 and t = +inf on a miss. */
void ref_primary_hits(const double* centres, const double* radii, int n, const double* cam12,
                      int W, int H, int32_t* idx, double* t) {
    int last = -1;
    hittable_list world;
    for (int k = 0; k < n; ++k)
        world.add(make_shared<tagged_sphere>(point3(centres[3 * k], centres[3 * k + 1], centres[3 * k + 2]),
                                             radii[k], k, &last));
    camera cam;
    set_camera(cam, cam12);
    for (int j = H - 1; j >= 0; --j)
        for (int i = 0; i < W; ++i) {
            double u = ((double)i + 0.5) / (W - 1);
            double v = ((double)j + 0.5) / (H - 1);
            hit_record rec;
            size_t o = (size_t)(H - 1 - j) * W + i;
            if (world.hit(cam.get_ray(u, v), 0, infinity, rec)) { idx[o] = last; t[o] = rec.t; }
            else { idx[o] = -1; t[o] = infinity; }
        }
}

/* hittable_list::hit on explicit rays.  rec_out per ray: t, p[3], normal[3], front_face (8 doubles). */
void ref_hit_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                   int nrays, double tmin, double tmax, int32_t* idx, double* rec_out) {
    int last = -1;
    hittable_list world;
    for (int k = 0; k < n; ++k)
        world.add(make_shared<tagged_sphere>(point3(centres[3 * k], centres[3 * k + 1], centres[3 * k + 2]),
                                             radii[k], k, &last));
    for (int q = 0; q < nrays; ++q) {
        ray r(point3(org[3 * q], org[3 * q + 1], org[3 * q + 2]), vec3(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2]));
        hit_record rec;
        double* o = rec_out + 8 * (size_t)q;
        if (world.hit(r, tmin, tmax, rec)) {
            idx[q] = last;
            o[0] = rec.t; o[1] = rec.p.x(); o[2] = rec.p.y(); o[3] = rec.p.z();
            o[4] = rec.normal.x(); o[5] = rec.normal.y(); o[6] = rec.normal.z(); o[7] = rec.front_face ? 1.0 : 0.0;
        } else {
            idx[q] = -1;
            for (int e = 0; e < 8; ++e) o[e] = 0.0;
        }
    }
}

/* sphere::hit, one (ray, sphere) pair per entry. */
void ref_sphere_hit_batch(const double* centre, const double* radius, const double* org, const double* dir,
                          int nq, double tmin, double tmax, int32_t* hit, double* rec_out) {
    for (int q = 0; q < nq; ++q) {
        sphere s(point3(centre[3 * q], centre[3 * q + 1], centre[3 * q + 2]), radius[q]);
        ray r(point3(org[3 * q], org[3 * q + 1], org[3 * q + 2]), vec3(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2]));
        hit_record rec;
        double* o = rec_out + 8 * (size_t)q;
        if (s.hit(r, tmin, tmax, rec)) {
            hit[q] = 1;
            o[0] = rec.t; o[1] = rec.p.x(); o[2] = rec.p.y(); o[3] = rec.p.z();
            o[4] = rec.normal.x(); o[5] = rec.normal.y(); o[6] = rec.normal.z(); o[7] = rec.front_face ? 1.0 : 0.0;
        } else {
            hit[q] = 0;
            for (int e = 0; e < 8; ++e) o[e] = 0.0;
        }
    }
}

/* ray_color on explicit rays; ray q draws from the shim stream seeded with seeds[q]. */
void ref_ray_color_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                         const uint64_t* seeds, int nrays, int depth, double* rgb_out) {
    hittable_list world;
    build_world(world, centres, radii, n);
    for (int q = 0; q < nrays; ++q) {
        ray r(point3(org[3 * q], org[3 * q + 1], org[3 * q + 2]), vec3(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2]));
        oracle_seed(seeds[q]);
        color c = ray_color(r, world, depth);
        rgb_out[3 * q] = c.x(); rgb_out[3 * q + 1] = c.y(); rgb_out[3 * q + 2] = c.z();
    }
}

/* ray_color_param on explicit rays (shading = ref_shading / orc_shading layout). */
void ref_ray_color_param_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                               const uint64_t* seeds, const void* shading, int nrays, int depth, double* rgb_out) {
    hittable_list world;
    build_world(world, centres, radii, n);
    const ref_shading sh = *static_cast<const ref_shading*>(shading);
    for (int q = 0; q < nrays; ++q) {
        ray r(point3(org[3 * q], org[3 * q + 1], org[3 * q + 2]), vec3(dir[3 * q], dir[3 * q + 1], dir[3 * q + 2]));
        oracle_seed(seeds[q]);
        color c = ray_color_param(r, world, depth, sh);
        rgb_out[3 * q] = c.x(); rgb_out[3 * q + 1] = c.y(); rgb_out[3 * q + 2] = c.z();
    }
}

/* ref_render_rows with ray_color_param in place of ray_color. */
void ref_render_rows_param(const double* centres, const double* radii, int n, const double* cam12,
                           int W, int H, int spp, int max_depth, uint64_t seed, const void* shading, int j0, int j1,
                           int nthreads, uint8_t* rgb, double* stats) {
    hittable_list world;
    build_world(world, centres, radii, n);
    camera cam;
    set_camera(cam, cam12);
    const ref_shading sh = *static_cast<const ref_shading*>(shading);
    long long tot_casts = 0, tot_black = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : tot_casts, tot_black)
#endif
    for (int j = j1 - 1; j >= j0; --j) {
        counting_world cw(world);
        oracle_seed(row_seed(seed, j));
        long long black = 0;
        for (int i = 0; i < W; ++i) {
            color pixel_color(0, 0, 0);
            for (int s = 0; s < spp; ++s) {
                double u = ((double)i + random_double()) / (W - 1);
                double v = ((double)j + random_double()) / (H - 1);
                color c = ray_color_param(cam.get_ray(u, v), cw, max_depth, sh);
                if (c.x() == 0 && c.y() == 0 && c.z() == 0) ++black;
                pixel_color += c;
            }
            std::ostringstream os;
            write_color(os, pixel_color, spp);
            int r = 0, g = 0, b = 0;
            std::istringstream is(os.str());
            is >> r >> g >> b;
            uint8_t* px = rgb + ((size_t)(H - 1 - j) * W + i) * 3;
            px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b;
        }
        tot_casts += cw.casts;
        tot_black += black;
    }
    if (stats) {
        stats[0] = (double)(j1 - j0) * W * spp;
        stats[1] = (double)tot_casts;
        stats[2] = (double)tot_black;
    }
}

/* camera::get_ray. out per query: origin[3], dir[3]. */
void ref_get_ray_batch(const double* cam12, const double* uv, int nq, double* out) {
    camera cam;
    set_camera(cam, cam12);
    for (int q = 0; q < nq; ++q) {
        ray r = cam.get_ray(uv[2 * q], uv[2 * q + 1]);
        double* o = out + 6 * (size_t)q;
        o[0] = r.orig.x(); o[1] = r.orig.y(); o[2] = r.orig.z();
        o[3] = r.dir.x(); o[4] = r.dir.y(); o[5] = r.dir.z();
    }
}

/* The default camera of camera.h:11-23 as the 12 doubles used above + aspect. */
void ref_default_camera(double* cam12, double* aspect) {
    camera cam;
    cam12[0] = cam.origin.x(); cam12[1] = cam.origin.y(); cam12[2] = cam.origin.z();
    cam12[3] = cam.lower_left_corner.x(); cam12[4] = cam.lower_left_corner.y(); cam12[5] = cam.lower_left_corner.z();
    cam12[6] = cam.horizontal.x(); cam12[7] = cam.horizontal.y(); cam12[8] = cam.horizontal.z();
    cam12[9] = cam.vertical.x(); cam12[10] = cam.vertical.y(); cam12[11] = cam.vertical.z();
    *aspect = cam.aspect_ratio;
}

/* write_color: summed pixel colour + spp -> the three integers it prints. */
void ref_write_color_batch(const double* rgb_sum, int nq, int spp, int32_t* out) {
    for (int q = 0; q < nq; ++q) {
        std::ostringstream os;
        write_color(os, color(rgb_sum[3 * q], rgb_sum[3 * q + 1], rgb_sum[3 * q + 2]), spp);
        std::istringstream is(os.str());
        int r = 0, g = 0, b = 0;
        is >> r >> g >> b;
        out[3 * q] = r; out[3 * q + 1] = g; out[3 * q + 2] = b;
    }
}

/* vec3::random_in_hemisphere from the shim stream seeded with seeds[q]. */
void ref_random_in_hemisphere_batch(const double* normal, const uint64_t* seeds, int nq, double* out) {
    for (int q = 0; q < nq; ++q) {
        oracle_seed(seeds[q]);
        vec3 v = vec3::random_in_hemisphere(vec3(normal[3 * q], normal[3 * q + 1], normal[3 * q + 2]));
        out[3 * q] = v.x(); out[3 * q + 1] = v.y(); out[3 * q + 2] = v.z();
    }
}

int ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  /* extern "C" */
