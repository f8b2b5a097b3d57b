// Host driver with the shape of the reference's main() (programs/main.cc:51-92): same camera, same
// two-sphere world, same image size / spp / depth, same P3 text on stdout -- but the pixel loop runs on
// the GPU through include/rt_host.hpp.  Usage: rt_main [width [spp [max_depth [seed]]]]
#include "raytracer.h"

#include "camera.h"
#include "color.h"
#include "hittable_list.h"
#include "sphere.h"

int main(int argc, char** argv) {
    camera cam;

    const int img_width = argc > 1 ? std::atoi(argv[1]) : 400;
    const int img_height = (int)(img_width / cam.aspect_ratio);
    const int sample_per_pixel = argc > 2 ? std::atoi(argv[2]) : 100;
    const int max_depth = argc > 3 ? std::atoi(argv[3]) : 50;
    const uint64_t seed = argc > 4 ? std::strtoull(argv[4], nullptr, 0) : 0;

    hittable_list world;
    world.add(make_shared<sphere>(point3(0, 0, -1), 0.5));
    world.add(make_shared<sphere>(point3(0, -100.5, 0), 100.0));

    try {
        const rt::frame img = rt::render(world, cam, img_width, img_height, sample_per_pixel, max_depth, seed);
        rt::write_ppm(std::cout, img);
        std::cerr << "\nDone. " << img.stats.samples << " samples, " << img.stats.casts << " casts, kernel "
                  << img.stats.kernel_ms << " ms\n";
    } catch (const rt::error& e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
