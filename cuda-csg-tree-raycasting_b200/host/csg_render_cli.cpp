// csg_render — headless harness replacing the reference's SDL/OpenGL/ImGui shell (Application.cpp, RenderManager.cpp)
// for the raycast path: load a scene file, render one frame (or a camera orbit), write a binary PPM, print timings.
//
//   csg_render scene.txt [--w 3840] [--h 2160] [--cam x y z pitch yaw] [--fov deg] [--light polar azimuth]
//              [--frames N] [--gpus N] [--no-optimize] [--ss K] [--pruning 0|1|2] [--out frame.ppm]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>

#include "csg_raycaster.hpp"

using namespace csg_b200;

int main(int argc, char** argv)
{
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s scene.txt [--w W] [--h H] [--cam x y z pitch yaw] [--fov deg] [--light polar azimuth] "
                             "[--frames N] [--gpus N] [--no-optimize] [--ss K] [--pruning 0|1|2] [--out frame.ppm]\n", argv[0]);
        return 2;
    }
    int w = 800, h = 600, frames = 1, gpus = 1, optimize = 1, samples = 1, pruning = 1;   // 800x600 is the reference's window size (Application.h:16-17)
    const char* out = nullptr;
    Camera cam;
    DirectionalLight light;
    for (int i = 2; i < argc; ++i) {
        auto need = [&](int k) { if (i + k >= argc) { std::fprintf(stderr, "missing value after %s\n", argv[i]); std::exit(2); } };
        if (!std::strcmp(argv[i], "--w")) { need(1); w = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--h")) { need(1); h = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--frames")) { need(1); frames = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--gpus")) { need(1); gpus = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--ss")) { need(1); samples = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--pruning")) { need(1); pruning = std::atoi(argv[++i]); }
        else if (!std::strcmp(argv[i], "--fov")) { need(1); cam.setFOV((float)std::atof(argv[++i])); }
        else if (!std::strcmp(argv[i], "--out")) { need(1); out = argv[++i]; }
        else if (!std::strcmp(argv[i], "--no-optimize")) optimize = 0;
        else if (!std::strcmp(argv[i], "--cam")) {
            need(5);
            cam.setPosition((float)std::atof(argv[i + 1]), (float)std::atof(argv[i + 2]), (float)std::atof(argv[i + 3]));
            cam.setRotation((float)std::atof(argv[i + 4]), (float)std::atof(argv[i + 5]));
            i += 5;
        } else if (!std::strcmp(argv[i], "--light")) { need(2); light.polar = (float)std::atof(argv[++i]); light.azimuth = (float)std::atof(argv[++i]); }
        else { std::fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
    }
    std::ifstream f(argv[1], std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        CSGTree tree = CSGTree::Parse(ss.str());   // Application::LoadCSGTree, Application.cpp:59-83
        csg_scene_set_optimize(tree.handle(), optimize);
        Raycaster rc;
        rc.ChangeSize(w, h, tree, gpus);
        if (samples != 1) rc.SetSupersampling(samples);
        if (pruning != 1) rc.SetPruning(pruning);
        std::vector<uint8_t> img((size_t)w * h * 4);
        for (int k = 0; k < frames; ++k) {
            auto t0 = std::chrono::steady_clock::now();
            rc.RaycastRGBA8(img.data(), cam, light);
            auto t1 = std::chrono::steady_clock::now();
            float ms = 0;
            csg_last_frame_ms(rc.context(), &ms);
            std::printf("frame %d: %.3f ms on the GPU, %.3f ms end to end, %.1f M rays/s\n", k, ms,
                        std::chrono::duration<double, std::milli>(t1 - t0).count(), w * (double)h / ms * 1e-3);
        }
        std::printf("%s\n", csg_context_info(rc.context()));
        if (out) {
            std::ofstream o(out, std::ios::binary);
            o << "P6\n" << w << " " << h << "\n255\n";
            for (int y = h - 1; y >= 0; --y)   // row 0 is the bottom scanline (GL order): flip for the PPM
                for (int x = 0; x < w; ++x) o.write(reinterpret_cast<const char*>(&img[((size_t)y * w + x) * 4]), 3);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "Cannot load tree: %s\n", e.what());   // Application.cpp:81
        return 1;
    }
    return 0;
}
