/* oracle/rt_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, FP64, contraction-free restatement of the reference's per-pixel
 * path-tracing hot path (fengye/PeterShirleyRaytracer, programs/).  It is the
 * checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this file
 * bit for bit against oracle/_ref/libref.so (the unmodified reference compiled
 * here) and against the fixtures in tests/golden/ generated from it
 * (tests/golden/make_golden.py), including the full PPM of the reference's own
 * main() (md5 46a8c6e4ae914126e69a67ccbe41250e with shim seed 0x9E3779B97F4A7C15).
 *
 * Conventions shared with include/rt.h:
 *   camera  = 12 doubles: origin[3], lower_left_corner[3], horizontal[3], vertical[3]
 *             (the public fields of programs/camera.h:31-35)
 *   frames  = row 0 is the TOP row (reference j = H-1), RGB or RGBA bytes
 *   pixel id for the counter-based RNG = j*W + i with j counted from the BOTTOM
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_RNG_RAND15 = 0, ORC_RNG_PHILOX = 1 };

typedef struct {
    double samples;       /* ray_color calls from the pixel loop */
    double casts;         /* world.hit calls */
    double black;         /* samples whose colour is exactly 0 (depth exhausted) */
    double primary_hits;  /* samples whose first cast hit something */
    double early_outs;    /* paths cut by the exact early-out (only if enabled) */
} orc_stats;

/* The constants programs/main.cc hard-codes in ray_color, as parameters.  orc_default_shading
 * fills the reference's values: tmin 0 (main.cc:40), albedo 0.5 (main.cc:43), sky (1,1,1) -> (0.5,0.7,1.0)
 * (main.cc:48), scatter = vec3::random_in_hemisphere (main.cc:42).  ORC_SCATTER_LAMBERTIAN swaps in
 * vec3::random_unit_vector (programs/vec3.h:97-100, present but unused by the reference's main.cc). */
enum { ORC_SCATTER_HEMISPHERE = 0, ORC_SCATTER_LAMBERTIAN = 1 };
typedef struct {
    double tmin;
    double albedo;
    double sky_a[3], sky_b[3];
    int scatter_mode;
} orc_shading;
void orc_default_shading(orc_shading* sh);

/* Philox4x32-10 (Salmon et al., SC'11), one block. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* main() of programs/main.cc:51-92 restated: default scene/camera, rand15 stream seeded with `seed`.
 * Writes the P3 text; returns its length or -needed. */
long orc_main_ppm(uint64_t seed, char* buf, long cap);

/* pixel loop main.cc:72-88 over rows [j0,j1) (j from the bottom).
 * rng_mode RAND15: one shim stream per row (same seeding as oracle/ref_harness.cc) -> equals ref_render_rows.
 * rng_mode PHILOX: stream keyed on (seed, pixel id, sample) -> equals the CUDA path.
 * early_out != 0 applies the exact `t == 0 && C == 0` cut (same image, fewer casts).
 * rgb: W*H*3, row 0 = top.  rgb_sum (optional): W*H*3 doubles, the summed radiance per pixel. */
void orc_render_rows(const double* centres, const double* radii, int n, const double* cam12,
                     int W, int H, int spp, int max_depth, uint64_t seed, int rng_mode, int early_out,
                     int j0, int j1, int nthreads, uint8_t* rgb, double* rgb_sum, orc_stats* stats);

/* the same with explicit shading parameters (NULL = the reference's constants) */
void orc_render_rows_ex(const double* centres, const double* radii, int n, const double* cam12,
                        int W, int H, int spp, int max_depth, uint64_t seed, int rng_mode, int early_out,
                        const orc_shading* shading, int j0, int j1, int nthreads, uint8_t* rgb, double* rgb_sum,
                        orc_stats* stats);

void orc_primary_hits(const double* centres, const double* radii, int n, const double* cam12,
                      int W, int H, int32_t* idx, double* t);

/* hittable_list::hit on explicit rays; rec_out per ray: t, p[3], normal[3], front_face */
void orc_hit_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                   int nrays, double tmin, double tmax, int32_t* idx, double* rec_out);

void orc_sphere_hit_batch(const double* centre, const double* radius, const double* org, const double* dir,
                          int nq, double tmin, double tmax, int32_t* hit, double* rec_out);

/* ray_color on explicit rays.  RAND15: ray q draws from the shim stream seeded with seeds[q].
 * PHILOX: ray q uses key seeds[0], pixel id q, sample 0 (bounce blocks start at 1). */
void orc_ray_color_batch(const double* centres, const double* radii, int n, const double* org, const double* dir,
                         const uint64_t* seeds, int rng_mode, int early_out, int nrays, int depth, double* rgb_out,
                         orc_stats* stats);

void orc_ray_color_batch_ex(const double* centres, const double* radii, int n, const double* org, const double* dir,
                            const uint64_t* seeds, int rng_mode, int early_out, const orc_shading* shading, int nrays,
                            int depth, double* rgb_out, orc_stats* stats);

void orc_get_ray_batch(const double* cam12, const double* uv, int nq, double* out);
void orc_default_camera(double* cam12, double* aspect);
void orc_write_color_batch(const double* rgb_sum, int nq, int spp, int32_t* out);
void orc_random_in_hemisphere_batch(const double* normal, const uint64_t* seeds, int nq, double* out);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif /* RT_ORACLE_H */
