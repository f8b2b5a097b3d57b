// Host-only check of include/rt_host.hpp: flatten order through nested lists, the reference API surface
// compiles and behaves (vec3 ops, sphere::hit closed interval, list tie rule), unsupported hittables are refused.
#include "raytracer.h"
#include "camera.h"
#include "color.h"
#include "hittable_list.h"
#include "sphere.h"

#include <cstdio>
#include <sstream>
#include <string>

struct box_like : hittable {
    bool hit(const ray&, double, double, hit_record&) const override { return false; }
};

#define CHECK(x) do { if (!(x)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "book") {  // dump the generated scene + camera for the Python comparison
        const hittable_list w = rt::book_scene(11, 42);
        std::vector<double> c, r;
        rt::flatten(w, c, r);
        std::printf("%zu\n", r.size());
        for (size_t k = 0; k < r.size(); ++k) std::printf("%a %a %a %a\n", c[3 * k], c[3 * k + 1], c[3 * k + 2], r[k]);
        const camera cam = rt::book_camera(1200, 800);
        const vec3 f[4] = {cam.origin, cam.lower_left_corner, cam.horizontal, cam.vertical};
        for (const vec3& v : f) std::printf("%a %a %a\n", v.x(), v.y(), v.z());
        return 0;
    }
    hittable_list inner;
    inner.add(make_shared<sphere>(point3(1, 2, 3), 0.5));
    inner.add(make_shared<sphere>(point3(4, 5, 6), 1.5));
    hittable_list world;
    world.add(make_shared<sphere>(point3(0, -100.5, 0), 100.0));
    world.add(make_shared<hittable_list>(inner));
    world.add(make_shared<sphere>(point3(7, 8, 9), 2.5));
    std::vector<double> c, r;
    rt::flatten(world, c, r);
    CHECK(r.size() == 4 && c.size() == 12);
    CHECK(r[0] == 100.0 && r[1] == 0.5 && r[2] == 1.5 && r[3] == 2.5);   // depth-first, list order
    CHECK(c[3] == 1 && c[4] == 2 && c[5] == 3 && c[9] == 7);

    world.add(make_shared<box_like>());
    bool threw = false;
    try { rt::flatten(world, c, r); } catch (const rt::error& e) { threw = e.code == RT_ERR_UNSUPPORTED; }
    CHECK(threw);   // no CPU fallback for unknown hittables

    // reference API behaviour on the host
    camera cam;
    ray q = cam.get_ray(0.5, 0.5);
    CHECK(q.origin().x() == 0 && q.direction().z() == -1.0);
    sphere s(point3(0, 0, -3), 1.0);
    hit_record rec;
    CHECK(s.hit(ray(point3(0, 0, 0), vec3(0, 0, -1)), 0, infinity, rec) && rec.t == 2.0 && rec.front_face);
    CHECK(s.hit(ray(point3(0, 0, 0), vec3(0, 0, -1)), 0, 2.0, rec));        // closed interval: t == tmax accepted
    CHECK(!s.hit(ray(point3(0, 0, 0), vec3(0, 0, -1)), 0, 1.999, rec));
    hittable_list twins;
    twins.add(make_shared<sphere>(point3(0, 0, -3), 1.0));
    twins.add(make_shared<sphere>(point3(0, 0, -3), 1.0));
    CHECK(twins.hit(ray(point3(0, 0, 0), vec3(0, 0, -1)), 0, infinity, rec) && rec.t == 2.0);
    std::ostringstream os;
    write_color(os, color(100, 25, 0), 100);
    CHECK(os.str() == "255 127 0\n");
    CHECK(dot(cross(vec3(1, 0, 0), vec3(0, 1, 0)), vec3(0, 0, 1)) == 1.0);
    CHECK(random_double() >= 0.0 && random_double() < 1.0);
    std::printf("ok\n");
    return 0;
}
