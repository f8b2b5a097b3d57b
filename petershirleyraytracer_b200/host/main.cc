// Host driver with the shape of the reference's main() (programs/main.cc:51-92): same camera, same
// two-sphere world, same image size / spp / depth, same P3 text on stdout -- but the pixel loop runs on
// the GPU through include/rt_host.hpp.
// Usage: rt_main [width [spp [max_depth [seed [passes]]]]] [--gpus N]
// passes > 1 renders progressively (rt::progressive_render), reporting "Samples done" where the reference
// reports "Scanline remaining" (programs/main.cc:74); the image does not depend on the number of passes.
// --gpus N deals the frame's tiles to GPUs 0..N-1 of this process (rt_render_multi): same image bit for bit.
// RT_MAIN_DEVICES=0,0 (a comma list) names the devices explicitly; the same device may appear twice (tests).
#include "raytracer.h"

#include "camera.h"
#include "color.h"
#include "hittable_list.h"
#include "sphere.h"

int main(int argc, char** argv) {
    camera cam;

    std::vector<std::string> pos;   // positional arguments, --gpus N taken out
    int gpus = 1;
    for (int i = 1; i < argc; ++i) {
        if (std::string(argv[i]) == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
        else pos.push_back(argv[i]);
    }
    const int img_width = pos.size() > 0 ? std::atoi(pos[0].c_str()) : 400;
    const int img_height = (int)(img_width / cam.aspect_ratio);
    const int sample_per_pixel = pos.size() > 1 ? std::atoi(pos[1].c_str()) : 100;
    const int max_depth = pos.size() > 2 ? std::atoi(pos[2].c_str()) : 50;
    const uint64_t seed = pos.size() > 3 ? std::strtoull(pos[3].c_str(), nullptr, 0) : 0;
    const int passes = pos.size() > 4 ? std::atoi(pos[4].c_str()) : 1;
    std::vector<int> devices;
    if (const char* e = std::getenv("RT_MAIN_DEVICES")) {
        for (const char* q = e; *q;) {
            devices.push_back(std::atoi(q));
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
    } else {
        for (int d = 0; d < (gpus > 0 ? gpus : 1); ++d) devices.push_back(d);
    }

    hittable_list world;
    world.add(make_shared<sphere>(point3(0, 0, -1), 0.5));
    world.add(make_shared<sphere>(point3(0, -100.5, 0), 100.0));

    try {
        rt::frame img;
        const rt_params params = rt::default_params(img_width, img_height, sample_per_pixel, max_depth, seed);
        if (passes > 1) {
            const rt::device_world dw(world, devices[0]);
            rt::progressive_render pr(dw, cam, params, devices[0]);
            for (int k = 0; k < passes; ++k) {
                const int upto = (int)((long long)sample_per_pixel * (k + 1) / passes);
                if (upto > pr.samples_done()) pr.add(upto - pr.samples_done(), /*want_stats=*/k == passes - 1);
                std::cerr << "\rSamples done: " << pr.samples_done() << ' ' << std::flush;
            }
            img = pr.current();
        } else if (devices.size() > 1) {
            const rt::device_world_group group(world, devices);
            img = rt::render(group, cam, params);
            std::cerr << "Rendered on " << group.devices() << " device worlds\n";
        } else {
            const rt::device_world dw(world, devices[0]);
            img = rt::render(dw, cam, params);
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
