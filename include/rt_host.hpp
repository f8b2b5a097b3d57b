// include/rt_host.hpp -- C++ host API above the C ABI (include/rt.h).
//
// Keeps the API surface of fengye/PeterShirleyRaytracer so that a program written against the reference
// (vec3 / ray / camera / hittable / hittable_list / sphere / ray_color / write_color) compiles unchanged,
// while the pixel loop of programs/main.cc:72-88 runs on the GPU:
//
//     camera cam;  hittable_list world;  world.add(make_shared<sphere>(point3(0,0,-1), 0.5)); ...
//     rt::frame img = rt::render(world, cam, img_width, img_height, samples_per_pixel, max_depth);
//     rt::write_ppm(std::cout, img);                       // the P3 text main() prints
//
// Names, argument meaning and behaviour follow the reference headers (cited per item).  The host-side
// hittable::hit implementations exist for API completeness (picking, tests); rendering never uses them --
// rt::render flattens the list and calls rt_upload_scene / rt_render, and fails if the scene holds anything the
// GPU path cannot represent (there is no CPU fallback).
#ifndef RT_HOST_HPP
#define RT_HOST_HPP

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rt.h"

using std::make_shared;
using std::shared_ptr;
using std::sqrt;

// ---- programs/raytracer.h:12-23
const double infinity = std::numeric_limits<double>::infinity();
const double pi = 3.1415926535897932385;
inline double degrees_to_radians(double degrees) { return degrees * pi / 180.0; }
template <class T>
inline T clamp(const T v, const T lo, const T hi) { return std::min(std::max(v, lo), hi); }

// ---- programs/random.h:4-14.  The reference divides by the int expression RAND_MAX + 1, which overflows
// where RAND_MAX == INT_MAX; the intended [0,1) value is produced here with a double divisor.
inline double random_double() { return std::rand() / (RAND_MAX + 1.0); }
inline double random_double(double min, double max) { return min + (max - min) * random_double(); }

// ---- programs/vec3.h
class vec3 {
public:
    double e[3];

    vec3() : e{0, 0, 0} {}
    vec3(double e0, double e1, double e2) : e{e0, e1, e2} {}

    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    double operator[](int i) const { return e[i]; }
    double& operator[](int i) { return e[i]; }

    vec3 operator-() const { return vec3(-e[0], -e[1], -e[2]); }
    vec3& operator+=(const vec3& o) { e[0] += o.e[0]; e[1] += o.e[1]; e[2] += o.e[2]; return *this; }
    vec3& operator*=(double t) { e[0] *= t; e[1] *= t; e[2] *= t; return *this; }
    vec3& operator/=(double t) { return *this *= 1 / t; }  // reciprocal multiply, like vec3.h:58-61

    double length_squared() const { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return sqrt(length_squared()); }

    static vec3 random() { return vec3(random_double(), random_double(), random_double()); }
    static vec3 random(double min, double max) {
        return vec3(random_double(min, max), random_double(min, max), random_double(min, max));
    }
    static vec3 random_in_unit_sphere() {  // vec3.h:83-95: keep the first cube point with len^2 <= 1
        for (;;) {
            const vec3 v = random(-1.0, 1.0);
            if (!(v.length_squared() > 1.0)) return v;
        }
    }
    static vec3 random_unit_vector();
    static vec3 random_in_hemisphere(const vec3& normal);
};
using point3 = vec3;
using color = vec3;

inline std::ostream& operator<<(std::ostream& out, const vec3& v) { return out << v[0] << ' ' << v[1] << ' ' << v[2]; }
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline vec3 operator*(double t, const vec3& v) { return vec3(t * v[0], t * v[1], t * v[2]); }
inline vec3 operator*(const vec3& v, double t) { return t * v; }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
inline vec3 operator/(const vec3& v, double t) { return (1 / t) * v; }  // vec3.h:151-154
inline double dot(const vec3& a, const vec3& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline vec3 cross(const vec3& a, const vec3& b) {
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline vec3 unit_vector(const vec3& v) { return v / v.length(); }
inline vec3 vec3::random_unit_vector() { return unit_vector(random_in_unit_sphere()); }
inline vec3 vec3::random_in_hemisphere(const vec3& normal) {  // vec3.h:102-109 (dot == 0 negates)
    const vec3 v = random_in_unit_sphere();
    return dot(v, normal) > 0 ? v : -v;
}

// ---- programs/ray.h
class ray {
public:
    point3 orig;
    vec3 dir;
    ray() {}
    ray(const point3& origin, const vec3& direction) : orig(origin), dir(direction) {}
    point3 origin() const { return orig; }
    vec3 direction() const { return dir; }
    point3 at(double t) const { return orig + dir * t; }
};

// ---- programs/camera.h (fixed 16:9 pinhole; the five fields are public and may be overwritten)
class camera {
public:
    double aspect_ratio;
    point3 origin;
    vec3 horizontal, vertical, lower_left_corner;
    camera() {
        aspect_ratio = 16.0 / 9.0;
        const double viewport_height = 2.0, viewport_width = viewport_height * aspect_ratio, focal_length = 1.0;
        origin = point3(0, 0, 0);
        horizontal = vec3(viewport_width, 0, 0);
        vertical = vec3(0, viewport_height, 0);
        lower_left_corner = origin - horizontal / 2.0 - vertical / 2.0 + vec3(0, 0, -focal_length);
    }
    ray get_ray(double u, double v) const { return ray(origin, lower_left_corner + horizontal * u + vertical * v - origin); }
};

// ---- programs/hittable.h
struct hit_record {
    point3 p;
    vec3 normal;
    double t;
    bool front_face;
    inline void set_face_normal(const ray& r, const vec3& outward_normal) {
        front_face = dot(r.direction(), outward_normal) < 0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};
class hittable {
public:
    virtual ~hittable() = default;
    virtual bool hit(const ray& r, double tmin, double tmax, hit_record& record) const = 0;
};

// ---- programs/sphere.h + sphere.cc
class sphere : public hittable {
public:
    point3 centre;
    double radius;
    sphere() : centre(0, 0, 0), radius(0) {}
    sphere(const point3& c, double r) : centre(c), radius(r) {}
    bool hit(const ray& r, double tmin, double tmax, hit_record& record) const override {
        const vec3 oc = r.origin() - centre;
        const double A = dot(r.direction(), r.direction()), half_b = dot(r.direction(), oc);
        const double C = dot(oc, oc) - radius * radius, disc = half_b * half_b - A * C;
        if (disc < 0) return false;
        const double sd = sqrt(disc);
        double t = (-half_b - sd) / A;
        if (t < tmin || t > tmax) {  // closed interval, near root first
            t = (-half_b + sd) / A;
            if (t < tmin || t > tmax) return false;
        }
        record.p = r.at(t);
        record.set_face_normal(r, (record.p - centre) / radius);
        record.t = t;
        return true;
    }
};

// ---- programs/hittable_list.h + hittable_list.cc
class hittable_list : public hittable {
public:
    std::vector<shared_ptr<hittable>> objects;
    hittable_list() {}
    explicit hittable_list(shared_ptr<hittable> object) { add(object); }
    void add(shared_ptr<hittable> object) { objects.push_back(object); }
    void clear() { objects.clear(); }
    bool hit(const ray& r, double tmin, double tmax, hit_record& record) const override {
        hit_record tmp;
        bool any = false;
        double closest = tmax;
        for (const auto& obj : objects)
            if (obj->hit(r, tmin, closest, tmp)) { any = true; closest = tmp.t; record = tmp; }  // ties -> later object
        return any;
    }
};

// ================================================================= GPU path
namespace rt {

struct error : std::runtime_error {
    int code;
    error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != RT_OK) throw error(rc, std::string("rt error ") + std::to_string(rc) + ": " + rt_last_error());
}

// Depth-first flatten of a hittable (nested hittable_lists allowed) into list-order sphere arrays.
inline void flatten(const hittable& h, std::vector<double>& centres, std::vector<double>& radii) {
    if (const auto* s = dynamic_cast<const sphere*>(&h)) {
        centres.push_back(s->centre.x()); centres.push_back(s->centre.y()); centres.push_back(s->centre.z());
        radii.push_back(s->radius);
    } else if (const auto* l = dynamic_cast<const hittable_list*>(&h)) {
        for (const auto& o : l->objects) {
            if (!o) throw error(RT_ERR_INVALID, "null object in hittable_list");
            flatten(*o, centres, radii);
        }
    } else {
        throw error(RT_ERR_UNSUPPORTED, "hittable is neither sphere nor hittable_list: no GPU representation, no CPU fallback");
    }
}

inline rt_camera to_abi(const camera& c) {
    rt_camera a;
    for (int i = 0; i < 3; ++i) {
        a.origin[i] = c.origin[i]; a.lower_left_corner[i] = c.lower_left_corner[i];
        a.horizontal[i] = c.horizontal[i]; a.vertical[i] = c.vertical[i];
    }
    return a;
}

// ---- Scene construction helpers.  The reference ships one hard-coded scene
// (programs/main.cc:62-63); these build the synthetic BASELINE scenes through the same public API, from a fixed
// splitmix64 stream, so that the C++ host, the Python binding (scenes.py) and the oracle see identical doubles.
struct splitmix64 {
    uint64_t s;
    explicit splitmix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double u01() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// Book-layout random spheres: ground r = 1000, a jittered grid of r = 0.2 spheres (cells [-g, g)^2), three r = 1
// spheres.  grid_half = 11 -> 485 spheres (BASELINE configs 3/5); 158 -> ~99.9k (config 4).
inline hittable_list book_scene(int grid_half = 11, uint64_t seed = 42) {
    splitmix64 rng(seed);
    hittable_list world;
    world.add(make_shared<sphere>(point3(0, -1000, 0), 1000));
    for (int a = -grid_half; a < grid_half; ++a)
        for (int b = -grid_half; b < grid_half; ++b) {
            rng.u01();  // mirrors the book's material draw
            const double x1 = rng.u01(), x2 = rng.u01();
            const point3 centre(a + 0.9 * x1, 0.2, b + 0.9 * x2);
            if ((centre - point3(4, 0.2, 0)).length() > 0.9) world.add(make_shared<sphere>(centre, 0.2));
        }
    world.add(make_shared<sphere>(point3(0, 1, 0), 1.0));
    world.add(make_shared<sphere>(point3(-4, 1, 0), 1.0));
    world.add(make_shared<sphere>(point3(4, 1, 0), 1.0));
    return world;
}

// A positionable pinhole camera expressed through the reference camera's public fields (programs/camera.h:31-35):
// the class itself is fixed at the origin looking down -z.
inline camera look_at_camera(const point3& lookfrom, const point3& lookat, const vec3& vup, double vfov_deg, double aspect) {
    const vec3 w = unit_vector(lookfrom - lookat), u = unit_vector(cross(vup, w)), v = cross(w, u);
    const double vh = 2.0 * std::tan(degrees_to_radians(vfov_deg) / 2.0), vw = vh * aspect;
    camera cam;
    cam.aspect_ratio = aspect;
    cam.origin = lookfrom;
    cam.horizontal = vw * u;
    cam.vertical = vh * v;
    cam.lower_left_corner = cam.origin - cam.horizontal / 2.0 - cam.vertical / 2.0 - w;
    return cam;
}
inline camera book_camera(int width, int height) {  // lookfrom (13,2,3) -> origin, vfov 20 degrees, no lens
    return look_at_camera(point3(13, 2, 3), point3(0, 0, 0), vec3(0, 1, 0), 20.0, (double)width / height);
}

// A world resident on one GPU (owns the rt_scene handle).
class device_world {
public:
    explicit device_world(const hittable& world, int device = 0) {
        std::vector<double> c, r;
        flatten(world, c, r);
        check(rt_upload_scene(c.data(), r.data(), (int32_t)r.size(), device, &scene_));
    }
    ~device_world() { rt_free_scene(scene_); }
    device_world(const device_world&) = delete;
    device_world& operator=(const device_world&) = delete;
    const rt_scene* handle() const { return scene_; }
    int size() const { return rt_scene_size(scene_); }
    // The same world after its spheres moved or were resized (same objects, same order): re-flattens and
    // updates the device buffers in place; refit keeps the BVH topology and only recomputes its boxes.
    void update(const hittable& world, bool refit = true) {
        std::vector<double> c, r;
        flatten(world, c, r);
        check(rt_update_scene(scene_, c.data(), r.data(), (int32_t)r.size(), refit ? 1 : 0));
    }

private:
    rt_scene* scene_ = nullptr;
};

// The same world on several GPUs of this process (one device_world per device): rt::render deals the frame's tiles to
// them and every GPU stores its pixels into the first one's frame over NVLink (rt_render_multi).
class device_world_group {
public:
    device_world_group(const hittable& world, const std::vector<int>& devices) {
        if (devices.empty()) throw error(RT_ERR_INVALID, "device_world_group needs at least one device");
        for (int d : devices) worlds_.push_back(std::make_unique<device_world>(world, d));
    }
    int devices() const { return (int)worlds_.size(); }
    const device_world& operator[](int i) const { return *worlds_[(size_t)i]; }
    void update(const hittable& world, bool refit = true) { for (auto& w : worlds_) w->update(world, refit); }
    std::vector<rt_scene*> handles() const {
        std::vector<rt_scene*> h;
        for (const auto& w : worlds_) h.push_back(const_cast<rt_scene*>(w->handle()));
        return h;
    }

private:
    std::vector<std::unique_ptr<device_world>> worlds_;
};

struct frame {
    int width = 0, height = 0;
    std::vector<uint8_t> rgba;  // row 0 = top row, the order main() prints
    rt_stats stats{};
};

inline rt_params default_params(int width, int height, int spp, int max_depth, uint64_t seed = 0) {
    rt_params p;
    check(rt_params_init(&p, width, height, spp, max_depth));  // tmin 0, albedo 0.5, the sky, hemisphere scatter: main.cc:40-48
    p.seed = seed;
    return p;
}
// The constants ray_color hard-codes (programs/main.cc:40,42,43,48) as arguments, e.g. the book's next chapter:
//     rt_params p = rt::with_shading(rt::default_params(w, h, spp, depth), 0.001, 0.5, RT_SCATTER_LAMBERTIAN);
inline rt_params with_shading(rt_params p, double tmin, double albedo, int scatter_mode = RT_SCATTER_HEMISPHERE,
                              const color& sky_a = color(1.0, 1.0, 1.0), const color& sky_b = color(0.5, 0.7, 1.0)) {
    p.tmin = tmin;
    p.custom_shading = 1; p.scatter_mode = scatter_mode; p.albedo = albedo;
    for (int i = 0; i < 3; ++i) { p.sky_a[i] = sky_a[i]; p.sky_b[i] = sky_b[i]; }
    return p;
}

// The triple loop of programs/main.cc:72-88 (+ ray_color + write_color's arithmetic) on the GPU.
inline frame render(const device_world& world, const camera& cam, const rt_params& p) {
    frame f;
    f.width = p.width; f.height = p.height;
    f.rgba.resize((size_t)p.width * p.height * 4);
    const rt_camera c = to_abi(cam);
    check(rt_render(world.handle(), &c, &p, f.rgba.data(), nullptr, &f.stats));
    return f;
}
inline frame render(const hittable& world, const camera& cam, int width, int height, int spp, int max_depth,
                    uint64_t seed = 0, int device = 0) {
    device_world dw(world, device);
    return render(dw, cam, default_params(width, height, spp, max_depth, seed));
}
// The same loop over several GPUs (tiles dealt to the devices; the frame is the single-GPU frame bit for bit).
inline frame render(const device_world_group& worlds, const camera& cam, const rt_params& p) {
    frame f;
    f.width = p.width; f.height = p.height;
    f.rgba.resize((size_t)p.width * p.height * 4);
    const rt_camera c = to_abi(cam);
    const std::vector<rt_scene*> h = worlds.handles();
    check(rt_render_multi(h.data(), (int32_t)h.size(), &c, &p, f.rgba.data(), &f.stats));
    return f;
}
inline frame render(const hittable& world, const camera& cam, int width, int height, int spp, int max_depth,
                    uint64_t seed, const std::vector<int>& devices) {
    device_world_group g(world, devices);
    return render(g, cam, default_params(width, height, spp, max_depth, seed));
}

// Progressive / resumable form of the same loop (where the reference can only report "Scanline remaining",
// programs/main.cc:74): add(n) traces the next n samples of every pixel and folds them into integer sums, so
// the frame after any sequence of add() calls equals a single render of that many samples bit for bit;
// sums() / resume() let a caller checkpoint the accumulator and continue later (or on another GPU).
class progressive_render {
public:
    progressive_render(const device_world& world, const camera& cam, const rt_params& p, int device = 0)
        : world_(world), cam_(to_abi(cam)), p_(p) {
        frame_.width = p.width; frame_.height = p.height;
        frame_.rgba.resize((size_t)p.width * p.height * 4);
        check(rt_accum_create(p.width, p.height, device, &accum_));
    }
    ~progressive_render() { rt_accum_destroy(accum_); }
    progressive_render(const progressive_render&) = delete;
    progressive_render& operator=(const progressive_render&) = delete;
    // Traces the next `samples` samples of every pixel into the device-resident sums (rt_accum_add: no allocation and no
    // host copy per pass; asynchronous).  current() fetches the frame when somebody wants to look at it.
    // want_stats waits for the pass and keeps its counters in current().stats.
    void add(int samples, bool want_stats = false) {
        rt_params q = p_;
        q.spp = samples;
        check(rt_accum_add(world_.handle(), &cam_, &q, accum_, want_stats ? &frame_.stats : nullptr));
        fresh_ = false;
    }
    int samples_done() const { return rt_accum_samples(accum_); }
    const frame& current() {
        if (!fresh_) {
            check(rt_accum_frame(accum_, frame_.rgba.data()));
            fresh_ = true;
        }
        return frame_;
    }
    std::vector<uint64_t> sums() const {  // 20.44 fixed point, W*H*3, row 0 = top: the checkpoint
        std::vector<uint64_t> s((size_t)p_.width * p_.height * 3);
        check(rt_accum_read(accum_, s.data()));
        return s;
    }
    void resume(const std::vector<uint64_t>& sums, int samples_done) {
        if (sums.size() != (size_t)p_.width * p_.height * 3 || samples_done < 0) throw error(RT_ERR_INVALID, "checkpoint does not match this frame");
        check(rt_accum_write(accum_, sums.data(), samples_done));
        fresh_ = false;
    }

private:
    const device_world& world_;
    rt_camera cam_;
    rt_params p_;
    rt_accum* accum_ = nullptr;
    frame frame_;
    bool fresh_ = false;
};

// programs/main.cc:70 + the per-pixel lines write_color emits (programs/color.h:21-23), from the 8-bit frame.
inline void write_ppm(std::ostream& out, const frame& f) {
    std::string s = "P3\n" + std::to_string(f.width) + ' ' + std::to_string(f.height) + "\n255\n";
    s.reserve(s.size() + (size_t)f.width * f.height * 12);
    for (size_t i = 0; i < (size_t)f.width * f.height; ++i) {
        s += std::to_string((int)f.rgba[4 * i]); s += ' ';
        s += std::to_string((int)f.rgba[4 * i + 1]); s += ' ';
        s += std::to_string((int)f.rgba[4 * i + 2]); s += '\n';
    }
    out << s;
}
inline void write_ppm_binary(std::ostream& out, const frame& f) {  // P6
    out << "P6\n" << f.width << ' ' << f.height << "\n255\n";
    for (size_t i = 0; i < (size_t)f.width * f.height; ++i) out.write(reinterpret_cast<const char*>(&f.rgba[4 * i]), 3);
}

}  // namespace rt

// ---- programs/main.cc:34-49 ray_color, GPU-backed.  One ray per call is the reference's signature; the
// batch form (rt_ray_color) is what a caller with many rays should use.  Random draws come from the Philox
// stream of include/rt.h (pixel id = `stream`, sample 0), not from rand().
inline color ray_color(const ray& r, const rt::device_world& world, int depth, uint64_t seed = 0) {
    const double o[3] = {r.orig.x(), r.orig.y(), r.orig.z()}, d[3] = {r.dir.x(), r.dir.y(), r.dir.z()};
    double rgb[3];
    rt::check(rt_ray_color(world.handle(), o, d, 1, depth, seed, 0, RT_SCAN_AUTO, rgb, nullptr));
    return color(rgb[0], rgb[1], rgb[2]);
}
inline color ray_color(const ray& r, const hittable& world, int depth) {
    const rt::device_world dw(world);  // uploads the scene: convenient, not fast
    return ray_color(r, dw, depth);
}

// ---- programs/color.h:8-24 write_color: one pixel's summed colour -> "r g b\n".  (The arithmetic is the
// device function write_color_channel; this host overload serves programs that still sum colours themselves.)
inline void write_color(std::ostream& out, const color& pixel_color, int samples_per_pixel) {
    const double scale = 1.0 / samples_per_pixel;
    const double r = sqrt(pixel_color.x() * scale), g = sqrt(pixel_color.y() * scale), b = sqrt(pixel_color.z() * scale);
    out << (int)(255.999 * clamp(r, 0.0, 0.999)) << ' ' << (int)(255.999 * clamp(g, 0.0, 0.999)) << ' '
        << (int)(255.999 * clamp(b, 0.0, 0.999)) << '\n';
}

#endif  // RT_HOST_HPP
