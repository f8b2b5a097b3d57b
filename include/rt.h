/* include/rt.h -- C ABI of the B200-native path-tracing hot path.
 *
 * Drop-in boundary for fengye/PeterShirleyRaytracer (reference files are cited as programs/<file>:<line>).
 * The reference has no FFI; its extension point is the C++ virtual
 *     hittable::hit(const ray&, double tmin, double tmax, hit_record&) const     programs/hittable.h:24
 * plus the free functions ray_color (programs/main.cc:34) and write_color (programs/color.h:8), all
 * driven by the pixel loop of main() (programs/main.cc:72-88).  A scene built through the reference's
 * API (hittable_list of spheres, camera with public fields) is flattened by include/rt_host.hpp into the
 * plain arrays below; everything from the pixel loop down runs on the GPU.
 *
 * Conventions
 *   - every function returns RT_OK (0) or a negative rt_status; nothing throws across this boundary;
 *     rt_last_error() returns a thread-local message for the last failure.
 *   - the caller owns host buffers; the library owns device memory behind the opaque rt_scene.
 *   - there is NO CPU fallback: without a CUDA device every entry point that computes returns
 *     RT_ERR_CUDA.
 *   - frames are W*H RGBA8 (A = 255), row 0 = TOP row = the reference's j = H-1 (programs/main.cc:72),
 *     i.e. the order main() prints pixels.
 *   - spheres keep list order: index k here == position k in hittable_list::objects
 *     (programs/hittable_list.h:40); ties in t go to the LATER index, as in programs/hittable_list.cc:11-15.
 *   - random numbers: Philox4x32-10, key = seed, counter = (pixel id, sample, block, 0) with pixel id =
 *     j*W + i (j counted from the bottom like the reference).  Replaces the global rand() stream of
 *     programs/random.h:4-8.  The stream, exactly (an independent implementation must reproduce it to get
 *     the same frames; the CPU checker used by this repository's tests does):
 *       block 0 of a (pixel, sample): the jitter draws of programs/main.cc:80-81, xi_u = w0 * 2^-32,
 *         xi_v = w1 * 2^-32 (32-bit uniforms; w2, w3 unused);
 *       bounce b (b = 1, 2, ...) takes block b, whose words w0..w3 carry the first TWO tries of the
 *         rejection loop of vec3::random_in_unit_sphere (programs/vec3.h:83-95) as 21-bit uniforms
 *         f in [0, 2^21), coordinate = -1 + 2 * (f * 2^-21) (random_double(-1, 1), programs/random.h:10-14):
 *           try A: fx = w0 >> 11, fy = w1 >> 11, fz = w2 >> 11;
 *           try B: fx = (w0 & 0x7ff) << 10 | w3 >> 22, fy = (w1 & 0x7ff) << 10 | (w3 >> 12 & 0x3ff),
 *                  fz = (w2 & 0x7ff) << 10 | (w3 >> 2 & 0x3ff);
 *         a try is kept unless len^2 > 1 (vec3.h:90).  If both are rejected (23 % of bounces) the loop
 *         continues with xorshift128 (Marsaglia 2003: t = x3; x3 = x2; x2 = x1; x1 = x0; t ^= t << 11;
 *         t ^= t >> 8; x0 = t ^ x1 ^ (x1 >> 19), state (x0..x3) = (w0..w3), x0 = 1 if all are zero),
 *         three successive outputs per try, f = output >> 11, until a try is kept.
 *       rt_ray_color: ray q uses pixel id q, sample 0, bounce blocks from 1.
 *   - thread safety: distinct rt_scene handles may be used from different host threads concurrently (renders
 *     that share the per-device constant bank are serialised on the device by the library); one handle must
 *     not be used by two threads at the same time.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 3

typedef enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,     /* bad argument */
    RT_ERR_CUDA = -2,        /* CUDA runtime failure or no device */
    RT_ERR_UNSUPPORTED = -3, /* e.g. linear scan requested for a scene that needs the BVH */
    RT_ERR_NOMEM = -4
} rt_status;

typedef struct rt_scene rt_scene; /* opaque: device-resident sphere buffers (+ BVH) */

/* The four public vec3 fields of the reference camera that get_ray reads (programs/camera.h:25-35). */
typedef struct {
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
} rt_camera;

enum { RT_SCAN_FILTERED = 0, RT_SCAN_EXACT = 1, RT_SCAN_BVH = 2, RT_SCAN_AUTO = 3 };
/* what ray_color adds to p + normal (programs/main.cc:42): vec3::random_in_hemisphere(normal) as the reference
 * does, or vec3::random_unit_vector() (programs/vec3.h:97-100, shipped but unused: the book's Lambertian) */
enum { RT_SCATTER_HEMISPHERE = 0, RT_SCATTER_LAMBERTIAN = 1 };

/* The constants main() and ray_color hard-code, as parameters. */
typedef struct {
    int32_t width, height;       /* programs/main.cc:57-58 */
    int32_t spp;                 /* programs/main.cc:66; 1 .. 2^20 */
    int32_t max_depth;           /* programs/main.cc:68; depth < 0 ends a path (main.cc:36) -> max_depth+1 casts */
    uint64_t seed;               /* Philox key */
    double tmin;                 /* programs/main.cc:40 passes 0 */
    int32_t jitter;              /* 1: main.cc:80-81; 0: sample at the pixel centre (i+0.5, j+0.5) */
    int32_t early_out;           /* 1: cut paths pinned at t==0 && C==0 (bit-identical image, fewer casts) */
    int32_t scan_mode;           /* RT_SCAN_*: FILTERED = FP32 conservative cull + FP64 exact test (default);
                                    EXACT = FP64 test of every sphere (validation); BVH; AUTO */
    int32_t shard_rank;          /* multi-GPU: this call renders tiles t with t % shard_count == shard_rank */
    int32_t shard_count;         /* 1 = whole frame */
    int32_t reserved[3];         /* 0 = defaults.  Tuning / A-B knobs that never change the frame: [0] paths per lane of the
                                    linear scan (1, 2, 4); [1] work units: n > 0 = n equal sample chunks per tile, -1 = automatic
                                    length but ungraded, -(10 + F) = short level of F/4 long units per warp (csrc/rt_units.h);
                                    [2] 1 = cull array from TMA-staged shared memory, 2 = round-1 BVH kernel, 3 = no tie grid */
    /* ABI 3: the constants ray_color hard-codes.  custom_shading == 0 (a zero-filled struct)
     * renders with the reference's values and ignores the four fields below, so the default path stays the
     * reference bit for bit. */
    int32_t custom_shading;      /* 1: use scatter_mode / albedo / sky_a / sky_b */
    int32_t scatter_mode;        /* RT_SCATTER_*; programs/main.cc:42 */
    double albedo;               /* programs/main.cc:43 returns 0.5 * ray_color(...); 0 <= albedo <= 1 */
    double sky_a[3], sky_b[3];   /* programs/main.cc:48: (1-t)*sky_a + t*sky_b, reference (1,1,1) and (0.5,0.7,1.0);
                                    each component in [0, 1] (radiance sums are 20.44 fixed point) */
} rt_params;

typedef struct {
    double kernel_ms;            /* CUDA-event time of the render kernel on its stream */
    uint64_t samples;            /* ray_color calls from the pixel loop */
    uint64_t casts;              /* world.hit calls */
    uint64_t sphere_tests;       /* FP32 cull tests executed (casts * N for the linear scan) */
    uint64_t node_tests;         /* BVH node box tests executed */
    uint64_t exact_tests;        /* FP64 sphere::hit evaluations */
    uint64_t black;              /* samples that returned exactly 0 */
    uint64_t early_outs;
    uint64_t primary_hits;
    uint64_t overflows;          /* casts that fell back to the full FP64 scan (candidate list full) */
    uint64_t launches;           /* kernels launched by this call */
    uint64_t self_resolved;      /* BVH mode: casts decided by the start-sphere test + tie grid, without a traversal */
} rt_stats;

typedef struct {
    int32_t tile_w, tile_h;      /* pixels */
    int32_t tiles_x, tiles_y;    /* tiles per row / column of the frame */
    int32_t tiles_total;
    int32_t tiles_per_shard;     /* ceil(tiles_total / shard_count): every shard buffer has this many tile slots */
    int64_t shard_bytes;         /* tiles_per_shard * tile_w * tile_h * 4 */
} rt_tile_layout;

int rt_abi_version(void);
const char* rt_last_error(void);

/* Fills *params with what programs/main.cc hard-codes (tmin 0, jitter on, albedo 0.5, the sky colours, hemisphere
 * scatter; custom_shading = 0) plus this library's defaults (seed 0, exact early-out on, RT_SCAN_AUTO, one shard). */
int rt_params_init(rt_params* params, int32_t width, int32_t height, int32_t spp, int32_t max_depth);

/* Flatten of hittable_list::objects (programs/hittable_list.h:40) with sphere::centre / radius
 * (programs/sphere.h:18-19): centres_xyz = 3n doubles, radii = n doubles, list order.  Uploads to
 * `device`; builds the FP32 cull array, the FP64 exact array and the flattened BVH. */
int rt_upload_scene(const double* centres_xyz, const double* radii, int32_t n, int32_t device, rt_scene** out);
/* Moved / resized spheres, same count and list order (an animation step).  Device buffers are reused.
 * refit = 1 keeps the BVH topology and recomputes its boxes (O(n), exact for any motion, cheaper traversal only
 * while the motion is moderate); refit = 0 rebuilds the tree.  Waits for renders in flight on this scene. */
int rt_update_scene(rt_scene* scene, const double* centres_xyz, const double* radii, int32_t n, int32_t refit);
void rt_free_scene(rt_scene* scene);
int rt_scene_size(const rt_scene* scene);

/* The pixel loop of programs/main.cc:72-88 + ray_color + write_color's arithmetic.  Synchronous.
 * rgba_out: host, W*H*4 bytes (shard_count must be 1).  radiance_sum_out (optional, may be NULL):
 * host, W*H*3 doubles, the per-pixel sum over samples that write_color receives. */
int rt_render(const rt_scene* scene, const rt_camera* cam, const rt_params* params, uint8_t* rgba_out,
              double* radiance_sum_out, rt_stats* stats_out);

/* Same, device buffers, asynchronous on `stream` (a cudaStream_t; NULL = default stream).
 * shard_count == 1: d_rgba is the W*H*4 frame.  shard_count > 1: d_rgba is this shard's compact tile
 * buffer (rt_tile_layout.shard_bytes; tile slot l holds frame tile l*shard_count + shard_rank), ready
 * for an all-gather followed by rt_deinterleave.  d_radiance_sum may be NULL.  stats_out (optional)
 * is filled by rt_render_finish. */
int rt_render_device(const rt_scene* scene, const rt_camera* cam, const rt_params* params, void* d_rgba,
                     void* d_radiance_sum, void* stream);
/* Waits for the last rt_render_device on this scene and reads its counters. */
int rt_render_finish(const rt_scene* scene, rt_stats* stats_out);

/* Progressive / resumable rendering (replaces the per-scanline progress of programs/main.cc:74 with real
 * checkpoints).  One pass traces samples [sample_begin, sample_begin + params->spp) of every pixel -- the Philox
 * streams are keyed on (pixel, absolute sample index) -- and ADDS their colours to the accumulator: W*H*3
 * uint64, 20.44 fixed point, frame pixel order (row 0 = top), zeroed by the caller before the first pass.  Integer
 * sums do not depend on order, so ANY split of [0, S) into passes leaves the same accumulator, and the same
 * frame, bit for bit, as a single render of S spp; the accumulator can be saved and a render resumed later.
 * rgba (optional) = write_color over all sample_begin + spp samples so far.  sample_begin + spp <= 2^20.
 * With shard_count > 1 a rank touches only its own tiles of d_accum; d_rgba is then the compact shard buffer. */
int64_t rt_accum_bytes(const rt_params* params);
int rt_render_pass(const rt_scene* scene, const rt_camera* cam, const rt_params* params, int32_t sample_begin,
                   uint64_t* accum /* host, in/out */, uint8_t* rgba_out /* host, may be NULL */, rt_stats* stats_out);
int rt_render_pass_device(const rt_scene* scene, const rt_camera* cam, const rt_params* params, int32_t sample_begin,
                          void* d_accum, void* d_rgba /* may be NULL */, void* stream);

/* Device-resident progressive accumulator: the same sums as rt_render_pass, kept on the GPU between passes -- no
 * allocation and no host copy per pass (a 64-pass render costs what one render costs).  rt_accum_add traces the next
 * params->spp samples of every pixel ([samples so far, + spp); params->width/height must match) and is asynchronous
 * unless stats_out is given; rt_accum_frame waits and copies write_color over all samples so far; rt_accum_read /
 * rt_accum_write move the raw W*H*3 uint64 sums (checkpoint / resume, also across devices). */
typedef struct rt_accum rt_accum;
int rt_accum_create(int32_t width, int32_t height, int32_t device, rt_accum** out);
void rt_accum_destroy(rt_accum* accum);
int rt_accum_samples(const rt_accum* accum);
int rt_accum_reset(rt_accum* accum);
int rt_accum_add(const rt_scene* scene, const rt_camera* cam, const rt_params* params, rt_accum* accum, rt_stats* stats_out);
int rt_accum_frame(const rt_accum* accum, uint8_t* rgba_out);
int rt_accum_read(const rt_accum* accum, uint64_t* sums_out);
int rt_accum_write(rt_accum* accum, const uint64_t* sums, int32_t samples_done);
/* write_color over a frame-ordered device accumulator (W*H*3 uint64) holding total_samples samples per pixel -> d_rgba.
 * Last step of a render whose samples were split over ranks and whose sums were added up with an integer all-reduce. */
int rt_accum_to_frame(const rt_params* params, const void* d_accum, int32_t total_samples, void* d_rgba, int32_t device,
                      void* stream);

/* rt_render on several GPUs of this process: scenes[i] holds the same spheres on device i (one rt_upload_scene per
 * device); the frame's 8x8 tiles are dealt to the scenes (tile t -> scene t % n_scenes), all shards render
 * concurrently, and every device stores its pixels straight into the first device's frame over NVLink peer access
 * (no gather step; without peer access the shards come back through the host).  Same frame, bit for bit, as
 * rt_render on one device.  params->shard_count must be 1.  stats_out: counters summed, kernel_ms = the slowest shard. */
int rt_render_multi(rt_scene* const* scenes, int32_t n_scenes, const rt_camera* cam, const rt_params* params,
                    uint8_t* rgba_out, rt_stats* stats_out);

int rt_get_tile_layout(const rt_params* params, rt_tile_layout* out);
/* d_gathered: shard_count consecutive shard buffers (all-gather output) -> d_rgba frame. */
int rt_deinterleave(const rt_params* params, const void* d_gathered, void* d_rgba, int32_t device, void* stream);

/* hittable_list::hit (programs/hittable_list.cc:3-20) for the camera ray through every pixel centre
 * (u = (i+0.5)/(W-1), v = (j+0.5)/(H-1)): idx_out (-1 = miss) and t_out (+inf on a miss), W*H, row 0 = top. */
int rt_primary_hits(const rt_scene* scene, const rt_camera* cam, const rt_params* params, int32_t* idx_out,
                    double* t_out);

/* hittable_list::hit on explicit rays: org/dir = 3*nrays doubles; idx_out = nrays; rec_out = 8 doubles per
 * ray: t, p[3], normal[3], front_face (programs/hittable.h:7-13); zeros on a miss. */
int rt_hit(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, double tmin, double tmax,
           int32_t scan_mode, int32_t* idx_out, double* rec_out);

/* ray_color (programs/main.cc:34-49) on explicit rays: ray q uses Philox pixel id q, sample 0, bounce
 * blocks from 1.  rgb_out = 3*nrays doubles. */
int rt_ray_color(const rt_scene* scene, const double* org, const double* dir, int32_t nrays, int32_t depth,
                 uint64_t seed, int32_t early_out, int32_t scan_mode, double* rgb_out, rt_stats* stats_out);

/* The same with tmin, the shading constants, max_depth (= depth), seed, early_out and scan_mode taken from *params
 * (width / height / spp / shards are not used). */
int rt_ray_color_params(const rt_scene* scene, const double* org, const double* dir, int32_t nrays,
                        const rt_params* params, double* rgb_out, rt_stats* stats_out);

/* write_color's arithmetic (programs/color.h:16-23) on device: summed colour + spp -> 3 ints per pixel. */
int rt_write_color(const double* rgb_sum, int32_t npix, int32_t spp, int32_t device, int32_t* out);

/* camera::get_ray (programs/camera.h:25-28) on device: uv = 2*nq doubles -> out = 6*nq (origin, dir). */
int rt_get_ray(const rt_camera* cam, const double* uv, int32_t nq, int32_t device, double* out);

/* Philox4x32-10 block on device (known-answer tests). */
int rt_philox(const uint32_t* ctr4, const uint32_t* key2, int32_t nblocks, int32_t device, uint32_t* out4);

/* Sustained FP32 FFMA issue rate of `device` (warp-instruction slots * 32 lanes per second, i.e.
 * FMA/s) measured with a register-resident FFMA loop; the denominator of the FP32 roofline. */
int rt_measure_fp32_peak(int32_t device, double* fma_per_s_out, double* ms_out);

/* Self-check: the hit distance t = num / A (programs/sphere.cc:24,29) is computed from one reciprocal per cast
 * plus two exact-residual corrections (DESIGN.md "division"); this compares that form with the IEEE division on n
 * pseudo-random operand pairs (all exponent ranges, zeros, denormals, infinities) and returns the number of
 * pairs whose quotients differ in any bit (must be 0). */
int rt_check_division(int32_t device, uint64_t n, uint64_t seed, uint64_t* mismatches_out);

int rt_device_info(int32_t device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, char* name, int32_t name_cap);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
