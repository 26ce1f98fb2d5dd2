// rtb_main.cpp — headless drop-in for src/main.cpp: `rtb_main [scene_id] [integrator_id]`
// renders on the GPU and writes output/sceneNN_integratorK_<unix-time>.png (main.cpp:135-151).
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>
#include <sys/stat.h>

extern "C" int rtbh_render_png(int, int, int, int, const char *, double *);
extern "C" const char *rtbh_last_error();

int main(int argc, char **argv) {
    const int scene_id = argc > 1 ? std::atoi(argv[1]) : 23;     // main.cpp:51-55
    const int integrator_id = argc > 2 ? std::atoi(argv[2]) : 4; // main.cpp:52,56-58
    mkdir("output", 0755);
    char name[128];
    std::snprintf(name, sizeof(name), "output/scene%02d_integrator%d_%ld.png", scene_id, integrator_id, long(std::time(nullptr)));
    double secs = 0;
    if (rtbh_render_png(scene_id, integrator_id, 0, 0, name, &secs) != 0) {
        std::fprintf(stderr, "rtb_main: %s\n", rtbh_last_error());
        return 1;
    }
    std::printf("Image saved successfully to %s (GPU render %.3f s)\n", name, secs);
    return 0;
}
