// Host driver with the shape of the reference's main() (programs/main.cc:51-92): same camera, same
// two-sphere world, same image size / spp / depth, same P3 text on stdout -- but the pixel loop runs on
// the GPU through include/rt_host.hpp.  Usage: rt_main [width [spp [max_depth [seed [passes]]]]]
// passes > 1 renders progressively (rt::progressive_render), reporting "Samples done" where the reference
// reports "Scanline remaining" (programs/main.cc:74); the image does not depend on the number of passes.
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
    const int passes = argc > 5 ? std::atoi(argv[5]) : 1;

    hittable_list world;
    world.add(make_shared<sphere>(point3(0, 0, -1), 0.5));
    world.add(make_shared<sphere>(point3(0, -100.5, 0), 100.0));

    try {
        rt::frame img;
        if (passes > 1) {
            const rt::device_world dw(world);
            rt::progressive_render pr(dw, cam, rt::default_params(img_width, img_height, sample_per_pixel, max_depth, seed));
            for (int k = 0; k < passes; ++k) {
                const int upto = (int)((long long)sample_per_pixel * (k + 1) / passes);
                if (upto > pr.samples_done()) pr.add(upto - pr.samples_done());
                std::cerr << "\rSamples done: " << pr.samples_done() << ' ' << std::flush;
            }
            img = pr.current();
        } else {
            img = rt::render(world, cam, img_width, img_height, sample_per_pixel, max_depth, seed);
        }
        rt::write_ppm(std::cout, img);
        std::cerr << "\nDone. " << img.stats.samples << " samples, " << img.stats.casts << " casts, kernel "
                  << img.stats.kernel_ms << " ms\n";
    } catch (const rt::error& e) {
        std::cerr << e.what() << "\n";
        return 1;
    }
    return 0;
}
